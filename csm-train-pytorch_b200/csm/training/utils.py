"""Training utilities — drop-in for /root/reference/src/csm/training/utils.py (PyTorch half).

``compute_loss`` keeps the reference signature and return contract (utils.py:56-63,119) and delegates to
``Model.forward`` (CUDA kernels).  Unlike the reference placeholder (utils.py:109-117) the acoustic term is real.
Checkpoint helpers keep the reference's file format and naming (utils.py:526-574,864-895).
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Optional, Tuple

import torch


def setup_logger(name: str, log_file: Optional[str] = None, level: int = logging.INFO) -> logging.Logger:
    logger = logging.getLogger(name)
    logger.setLevel(level)
    for h in logger.handlers[:]:
        logger.removeHandler(h)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    ch = logging.StreamHandler()
    ch.setLevel(level)
    ch.setFormatter(fmt)
    logger.addHandler(ch)
    if log_file:
        d = os.path.dirname(log_file)
        if d:
            os.makedirs(d, exist_ok=True)
        fh = logging.FileHandler(log_file)
        fh.setLevel(level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    return logger


def compute_loss(model, input_tokens: torch.Tensor, input_masks: torch.Tensor, target_audio_tokens: torch.Tensor,
                 semantic_weight: float = 100.0, acoustic_weight: float = 1.0, *,
                 frame_idx: Optional[torch.Tensor] = None, decoder_frame_fraction: float = 1.0 / 16,
                 target_lengths: Optional[torch.Tensor] = None, mask_padded_targets: bool = False,
                 speaker_ids: Optional[torch.Tensor] = None, segment_starts: Optional[torch.Tensor] = None,
                 segment_ends: Optional[torch.Tensor] = None,
                 target_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """(total loss, {"semantic_loss", "acoustic_loss", "per_codebook_loss"}) — utils.py:56-119.  The keyword-only
    arguments are extensions (explicit decoder frames, true target lengths); without them the call is the reference's."""
    return model(input_tokens, input_masks, target_audio_tokens, frame_idx=frame_idx,
                 decoder_frame_fraction=decoder_frame_fraction, semantic_weight=semantic_weight,
                 acoustic_weight=acoustic_weight, target_lengths=target_lengths,
                 mask_padded_targets=mask_padded_targets, speaker_ids=speaker_ids, segment_starts=segment_starts,
                 segment_ends=segment_ends, target_mask=target_mask)


def batch_loss(model, batch: Dict[str, torch.Tensor], semantic_weight: float = 100.0, acoustic_weight: float = 1.0,
               mask_padded_targets: bool = False) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """compute_loss on a batch dict as the trainers' collate / pack functions build it: the three reference keys plus
    the optional extensions (frame_idx, target_lengths, speaker_ids, segment_starts / segment_ends / target_mask)."""
    return compute_loss(model, batch["input_tokens"], batch["input_masks"], batch["target_audio_tokens"],
                        semantic_weight, acoustic_weight, frame_idx=batch.get("frame_idx"),
                        target_lengths=batch.get("target_lengths"), mask_padded_targets=mask_padded_targets,
                        speaker_ids=batch.get("speaker_ids"), segment_starts=batch.get("segment_starts"),
                        segment_ends=batch.get("segment_ends"), target_mask=batch.get("target_mask"))


def save_checkpoint(model, optimizer, epoch: int, global_step: int, loss: float, save_dir: str,
                    name: str = "checkpoint") -> str:
    os.makedirs(save_dir, exist_ok=True)
    payload = {"model": model.state_dict(), "optimizer": optimizer.state_dict() if optimizer is not None else None,
               "epoch": epoch, "global_step": global_step, "loss": loss}
    path = os.path.join(save_dir, f"{name}_epoch{epoch}_step{global_step}.pt")
    torch.save(payload, path)
    torch.save(payload, os.path.join(save_dir, f"{name}_latest.pt"))
    return path


def load_checkpoint(checkpoint_path: str, model, optimizer=None, device="cuda") -> Dict:
    ckpt = torch.load(checkpoint_path, map_location=device)
    model.load_state_dict(ckpt["model"])
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    return {"epoch": ckpt.get("epoch", 0), "global_step": ckpt.get("global_step", 0),
            "loss": ckpt.get("loss", float("inf"))}
