"""``CSMLoRATrainer`` — API of /root/reference/src/csm/training/lora_trainer.py:29-633 (an MLX/Apple class in the
reference) re-homed on the PyTorch ``Model`` and the B200 kernels; the arithmetic follows the PyTorch model, not the
MLX restatement.  Kept: constructor arguments, ``prepare_optimizer``, ``train_step(batch)``,
``train(...)`` (signature of mlx_trainer.py:733-743), ``save_model(path, save_mode in {lora, full, both})`` with the
``_metadata.json`` side file (lora_trainer.py:459-528) and ``load_lora_weights``.  Errors propagate: no
swallow-and-placeholder losses (lora_trainer.py:448-457 is not reproduced).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional

import torch

from ..models import lora as lora_mod
from ..models.model import Model
from . import dp
from .graph import GraphedStep
from .trainer import clip_and_step, csm_1b_args, iterate_batches, make_optimizer
from .utils import batch_loss, compute_loss, setup_logger


class CSMLoRATrainer:
    def __init__(self, model_path: str, output_dir: str, log_file: Optional[str] = None,
                 learning_rate: float = 1e-4, semantic_weight: float = 100.0, acoustic_weight: float = 1.0,
                 weight_decay: float = 0.01, lora_r: int = 8, lora_alpha: float = 16.0, lora_dropout: float = 0.0,
                 target_modules: Optional[List[str]] = None, target_layers: Optional[List[int]] = None,
                 lora_use_bias: bool = False, *, model: Optional[Model] = None, device: str = "cuda",
                 num_adapters=1, target_decoder_layers: Optional[List[int]] = None):
        if not 0.0 <= lora_dropout < 1.0:
            raise ValueError("lora_dropout must be in [0, 1)")
        self.model_path = model_path
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.rank, self.world, self.local_rank = dp.init_distributed()
        self.device = f"cuda:{self.local_rank}" if (device == "cuda" and self.world > 1) else device
        self.logger = setup_logger("csm_lora_trainer", log_file or str(self.output_dir / "lora_training.log"))
        self.learning_rate, self.weight_decay = learning_rate, weight_decay
        self.semantic_weight, self.acoustic_weight = semantic_weight, acoustic_weight
        self.lora_r, self.lora_alpha, self.lora_dropout = lora_r, lora_alpha, lora_dropout
        self.target_modules = target_modules or list(lora_mod.DEFAULT_TARGETS)
        self.target_layers, self.lora_use_bias = target_layers, lora_use_bias
        # extensions used by MultiSpeakerLoRATrainer: several adapters per projection (one per speaker) in one model
        self.num_adapters, self.target_decoder_layers = num_adapters, target_decoder_layers
        self.decoder_frame_fraction = 1.0 / 16
        # batches from collate_pinned carry each sample's true target length: padded all-zero target frames are then
        # left out of the semantic mean as well (the reference averages over them, utils.py:101-105; set False for that)
        self.mask_padded_targets = True
        self.pack_sequences_to: Optional[int] = None       # sequence packing (see CSMTrainer)
        self.compact_tokens: bool = False                  # batches as int32 rows + mask words (see CSMTrainer)
        self.model = model
        self.optimizer = None
        self._sync = None
        self._graphed = None
        self._load_model_with_lora()
        self.epoch, self.global_step, self.best_loss = 0, 0, float("inf")

    def _load_model_with_lora(self):
        if self.model is None:
            if not self.model_path:
                self.logger.warning("Empty model path provided. Model will need to be set manually.")
                return
            self.model = Model(csm_1b_args()).to(torch.bfloat16)
            if self.model_path.endswith(".safetensors"):
                from safetensors.torch import load_file
                state = load_file(self.model_path)
            else:
                state = torch.load(self.model_path, map_location="cpu")
                if "model" in state and "audio_head" not in state:
                    state = state["model"]
            self.model.load_state_dict(state)
        self.model = self.model.to(torch.bfloat16).to(self.device)
        self.lora_names = lora_mod.apply_lora(self.model, self.lora_r, self.lora_alpha, self.target_modules,
                                              self.target_layers, num_adapters=self.num_adapters,
                                              target_decoder_layers=self.target_decoder_layers,
                                              dropout=self.lora_dropout, use_bias=self.lora_use_bias)

    def set_model(self, model: Model):
        self.model = model
        self._load_model_with_lora()

    def get_lora_params(self) -> Dict[str, torch.nn.Parameter]:
        return {n: p for n, p in self.model.named_parameters() if n.endswith(("lora_A", "lora_B", "lora_bias"))}

    def prepare_optimizer(self):
        params = list(self.get_lora_params().values())
        n = sum(p.numel() for p in params)
        self.logger.info(f"LoRA: {len(params)} tensors, {n:,} trainable parameters")
        self.optimizer = make_optimizer(params, self.learning_rate, self.weight_decay)
        self._sync = dp.GradSynchronizer(params)

    def enable_cuda_graph(self, warmup: int = 3, max_grad_norm: float = 1.0) -> None:
        """Replay the whole optimiser step as one CUDA graph once shapes are static (see training/graph.py)."""
        if self.optimizer is None:
            self.prepare_optimizer()
        self._graph_max_norm = max_grad_norm
        self._graphed = GraphedStep(lambda b: self._step_impl(b, max_grad_norm), self.device, warmup)

    def _to_device(self, batch):
        if "frame_idx" not in batch:
            batch = dict(batch)
            batch["frame_idx"] = Model.select_frames(batch["input_masks"], batch["target_audio_tokens"].shape[1],
                                                     self.decoder_frame_fraction,
                                                     target_lengths=batch.get("target_lengths"))
        return {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}

    def train_step(self, batch, max_grad_norm: float = 1.0) -> torch.Tensor:
        """One optimiser step on one batch {input_tokens, input_masks, target_audio_tokens[, frame_idx]} (host or
        device tensors).  Returns the detached loss tensor (on device; reading it is the caller's D2H)."""
        if self.optimizer is None:
            self.prepare_optimizer()
        if "frame_idx" not in batch:
            batch = dict(batch)
            batch["frame_idx"] = Model.select_frames(batch["input_masks"], batch["target_audio_tokens"].shape[1],
                                                     self.decoder_frame_fraction,
                                                     target_lengths=batch.get("target_lengths"))
        self.global_step += 1
        if self._graphed is not None and max_grad_norm == self._graph_max_norm:
            return self._graphed(batch)
        return self._step_impl({k: v.to(self.device, non_blocking=True) for k, v in batch.items()}, max_grad_norm)

    def _step_impl(self, b, max_grad_norm: float) -> torch.Tensor:
        loss, _ = batch_loss(self.model, b, self.semantic_weight, self.acoustic_weight, self.mask_padded_targets)
        loss.backward()
        self._sync.finish()
        clip_and_step(self.optimizer, list(self.get_lora_params().values()), max_grad_norm)
        self.optimizer.zero_grad(set_to_none=True)
        return loss.detach()

    def train(self, train_dataset, val_dataset=None, batch_size: int = 2, epochs: int = 5, val_every: int = 100,
              save_every: int = 500, max_grad_norm: float = 1.0, resume_from: Optional[str] = None):
        if self.optimizer is None:
            self.prepare_optimizer()
        if resume_from:
            self.load_lora_weights(resume_from)
        self.model.train()
        for epoch in range(self.epoch, self.epoch + epochs):
            losses = []
            for batch in iterate_batches(train_dataset, batch_size, True, self.rank, self.world, seed=epoch,
                                         pack_to=self.pack_sequences_to,
                                         compact_vocab=(int(self.model.args.audio_vocab_size)
                                                        if self.compact_tokens else None)):
                losses.append(self.train_step(batch, max_grad_norm))
                if val_dataset is not None and self.global_step % val_every == 0:
                    val = self._validate(val_dataset, batch_size)
                    self.logger.info(f"Epoch {epoch + 1}, Step {self.global_step}, Val Loss: {val:.6f}")
                    if val < self.best_loss:
                        self.best_loss = val
                        if self.rank == 0:
                            self.save_model(str(self.output_dir / "best.safetensors"), "lora")
                if self.global_step % save_every == 0 and self.rank == 0:
                    self.save_model(str(self.output_dir / f"checkpoint_step{self.global_step}.safetensors"), "lora")
            avg = float(torch.stack(losses).mean()) if losses else float("nan")
            self.logger.info(f"Epoch {epoch + 1} Avg Loss: {avg:.6f}")
            self.epoch = epoch + 1
        if self.rank == 0:
            self.save_model(str(self.output_dir / "final.safetensors"), "lora")
        return self.best_loss

    def _validate(self, val_dataset, batch_size: int = 2) -> float:
        self.model.eval()
        total, n = 0.0, 0
        with torch.no_grad():
            for batch in iterate_batches(val_dataset, batch_size, False):
                b = self._to_device(batch)
                loss, _ = batch_loss(self.model, b, self.semantic_weight, self.acoustic_weight, self.mask_padded_targets)
                total += float(loss)
                n += 1
        self.model.train()
        return total / max(1, n)

    # ---- persistence (lora_trainer.py:459-633) ------------------------------------------------------------
    def save_model(self, save_path: str, save_mode: str = "lora"):
        if save_mode not in ("lora", "full", "both"):
            raise ValueError(f"save_mode must be lora|full|both, got {save_mode!r}")
        from safetensors.torch import save_file
        d = os.path.dirname(save_path)
        if d:
            os.makedirs(d, exist_ok=True)
        if save_mode in ("lora", "both"):
            lora_path = save_path.replace(".safetensors", "_lora.safetensors") if save_mode == "both" else save_path
            tensors = {n: p.detach().float().cpu().contiguous() for n, p in self.get_lora_params().items()}
            save_file(tensors, lora_path)
            meta = {"lora_r": self.lora_r, "lora_alpha": self.lora_alpha, "lora_dropout": self.lora_dropout,
                    "target_modules": self.target_modules, "target_layers": self.target_layers,
                    "lora_use_bias": self.lora_use_bias, "params_count": len(tensors)}
            with open(lora_path.replace(".safetensors", "_metadata.json"), "w") as f:
                json.dump(meta, f, indent=2)
        if save_mode in ("full", "both"):
            full_path = save_path.replace(".safetensors", "_full.safetensors") if save_mode == "both" else save_path
            merged = {}
            adapters = {n.rsplit(".", 1)[0] for n in self.get_lora_params()}
            for n, p in self.model.named_parameters():
                if n.endswith(("lora_A", "lora_B", "lora_bias")):
                    continue
                t = p.detach()
                mod_name = n.rsplit(".", 1)[0]
                if n.endswith(".weight") and mod_name in adapters:
                    mod = self.model.get_submodule(mod_name)
                    t = t.clone()
                    from .. import ops
                    ops.gemm(mod.lora_B.data, mod.lora_A.data, trans_b=True, out=t, accumulate=True,
                             alpha=float(mod.lora_scaling))          # W0 + (alpha/r) B A  (lora.py:140-153)
                merged[n] = t.cpu().contiguous()
            save_file(merged, full_path)

    def _invalidate_optimizer_master(self):
        """Parameters were overwritten outside the optimiser (a loaded adapter): drop the fp32 master copies so the
        next step re-derives them from the new bf16 values."""
        if self.optimizer is not None:
            for st in self.optimizer.state.values():
                st.pop("master", None)
            if hasattr(self.optimizer, "_table_key"):
                self.optimizer._table_key = None

    def load_lora_weights(self, lora_path: str):
        from safetensors.torch import load_file
        tensors = load_file(lora_path)
        params = self.get_lora_params()
        missing = set(params) - set(tensors)
        if missing:
            raise KeyError(f"LoRA file lacks {len(missing)} tensors, e.g. {sorted(missing)[:3]}")
        with torch.no_grad():
            for n, p in params.items():
                p.copy_(tensors[n].to(p.dtype))
        self._invalidate_optimizer_master()
