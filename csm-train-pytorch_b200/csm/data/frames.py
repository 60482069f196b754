"""Frame layout of the CSM training samples from PRE-TOKENISED ids.

The reference builds its [frames, 33] token / mask tensors inside ``Generator`` / ``CSMDataset`` with the Llama text
tokenizer and the Mimi codec in the loop (generator.py:77-130, training_data.py:245-302).  Neither tokenizer is part of
the training hot path; what the kernels consume is the layout, restated here on integer ids:

  * text frame  : column 32 = text token id, mask only column 32           (generator.py:91-95)
  * audio frame : columns 0..31 = the 32 codebook ids, mask columns 0..31   (generator.py:121-124)
  * every audio segment ends with an all-zero EOS frame (mask columns 0..31) (generator.py:117-119)
  * a sample = context segments (text + audio each), then the target's text (training_data.py:271-284); the target's audio
    codes are the prediction targets [T, 32]

plus the batch builders the trainers use: zero / False padding to the batch maximum (training_data.py:379-408) into
pinned host tensors, and a length-bucketed batch order that keeps padding (wasted backbone FLOPs) low.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

AUDIO_CODEBOOKS = 32


def text_frames(text_ids: Sequence[int], codebooks: int = AUDIO_CODEBOOKS) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n, C+1] int64 tokens and bool mask for n text token ids (generator.py:88-95)."""
    ids = torch.as_tensor(list(text_ids), dtype=torch.int64)
    tok = torch.zeros(ids.numel(), codebooks + 1, dtype=torch.int64)
    msk = torch.zeros(ids.numel(), codebooks + 1, dtype=torch.bool)
    tok[:, -1] = ids
    msk[:, -1] = True
    return tok, msk


def audio_frames(codes: torch.Tensor, add_eos: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """codes int64 [C, T] (Mimi layout: codebook-major) -> [T(+1), C+1] tokens / mask (generator.py:113-124)."""
    if codes.dim() != 2:
        raise ValueError("audio codes must be [codebooks, frames]")
    C, T = codes.shape
    codes = codes.to(torch.int64)
    if add_eos:
        codes = torch.cat([codes, torch.zeros(C, 1, dtype=torch.int64)], dim=1)
    tok = torch.zeros(codes.shape[1], C + 1, dtype=torch.int64)
    msk = torch.zeros(codes.shape[1], C + 1, dtype=torch.bool)
    tok[:, :-1] = codes.t()
    msk[:, :-1] = True
    return tok, msk


def build_sample(context: Iterable[Tuple[Sequence[int], torch.Tensor]], target_text_ids: Sequence[int],
                 target_codes: torch.Tensor, max_seq_len: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """One training sample: context segments (text ids, audio codes [C, T]) followed by the target's text frames as
    input; the target's audio codes [C, T] as ``target_audio_tokens`` [T, C] (training_data.py:245-302, including its
    truncation rule: cut from the beginning, the target text always survives)."""
    toks: List[torch.Tensor] = []
    msks: List[torch.Tensor] = []
    for text_ids, codes in context:
        for t, m in (text_frames(text_ids, codes.shape[0]), audio_frames(codes)):
            toks.append(t)
            msks.append(m)
    C = target_codes.shape[0]
    tt, tm = text_frames(target_text_ids, C)
    toks.append(tt)
    msks.append(tm)
    tokens, masks = torch.cat(toks, dim=0), torch.cat(msks, dim=0)
    if max_seq_len is not None and tokens.shape[0] > max_seq_len:
        keep = min(max_seq_len, tt.shape[0])                 # training_data.py:289-294
        tokens, masks = tokens[tokens.shape[0] - keep:], masks[masks.shape[0] - keep:]
    return {"input_tokens": tokens, "input_masks": masks, "target_audio_tokens": target_codes.t().contiguous().long()}


def collate_pinned(batch: List[Dict[str, torch.Tensor]], pin: bool = True,
                   pad_to_multiple: int = 1) -> Dict[str, torch.Tensor]:
    """Zero / False padding to the batch maximum (training_data.py:379-408) straight into pinned host tensors, so the
    trainer's H2D copies are asynchronous.  The target length is raised to the input length (the semantic term reads
    targets[:, :S-1]); ``pad_to_multiple`` rounds the frame count up (e.g. 128: whole attention tiles).  One key more
    than the reference's collate: ``target_lengths`` int64 [B], each sample's target frames before padding, so that the
    decoder is never trained on padded all-zero code frames (Model.select_frames)."""
    S = max(b["input_tokens"].shape[0] for b in batch)
    S = (S + pad_to_multiple - 1) // pad_to_multiple * pad_to_multiple
    T = max(max(b["target_audio_tokens"].shape[0] for b in batch), S)
    W = batch[0]["input_tokens"].shape[1]
    C = batch[0]["target_audio_tokens"].shape[1]
    pin = pin and torch.cuda.is_available()
    tok = torch.zeros(len(batch), S, W, dtype=torch.int64, pin_memory=pin)
    msk = torch.zeros(len(batch), S, W, dtype=torch.bool, pin_memory=pin)
    tgt = torch.zeros(len(batch), T, C, dtype=torch.int64, pin_memory=pin)
    lens = torch.zeros(len(batch), dtype=torch.int64, pin_memory=pin)
    for i, b in enumerate(batch):
        s, t = b["input_tokens"].shape[0], b["target_audio_tokens"].shape[0]
        tok[i, :s] = b["input_tokens"]
        msk[i, :s] = b["input_masks"].bool()
        tgt[i, :t] = b["target_audio_tokens"]
        lens[i] = t
    return {"input_tokens": tok, "input_masks": msk, "target_audio_tokens": tgt, "target_lengths": lens}


def length_bucketed_order(lengths: Sequence[int], batch_size: int, seed: int = 0, window: int = 50) -> List[List[int]]:
    """Batches of sample indices with similar lengths: shuffle, sort inside windows of ``window * batch_size`` samples,
    cut into batches, shuffle the batches.  Padding frames cost full backbone FLOPs; on a long-tailed length
    distribution this removes most of them while keeping the epoch order random."""
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(len(lengths), generator=g).tolist()
    span = max(1, window * batch_size)
    batches: List[List[int]] = []
    for i in range(0, len(order), span):
        chunk = sorted(order[i:i + span], key=lambda j: lengths[j])
        batches += [chunk[k:k + batch_size] for k in range(0, len(chunk), batch_size)]
    perm = torch.randperm(len(batches), generator=g).tolist()
    return [batches[k] for k in perm]


def padding_fraction(lengths: Sequence[int], batches: Sequence[Sequence[int]]) -> float:
    """Fraction of frames in the padded batches that are padding."""
    real = sum(lengths[j] for b in batches for j in b)
    padded = sum(max(lengths[j] for j in b) * len(b) for b in batches if b)
    return 1.0 - real / max(1, padded)
