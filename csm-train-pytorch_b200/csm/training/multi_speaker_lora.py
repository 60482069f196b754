"""``MultiSpeakerLoRATrainer`` — API of /root/reference/src/csm/training/multi_speaker_lora.py:29-437 as multi-adapter
BATCHING on one GPU-resident model.

The reference keeps one ``CSMLoRATrainer`` — i.e. one full model copy — per speaker and trains them one after the other
(multi_speaker_lora.py:137-213, 276-300), optionally starting each from a "shared" adapter file.  Here ONE base model
holds every speaker's adapters side by side (``lora_A`` [K*r, in], ``lora_B`` [out, K*r] per adapted projection,
csm/models/lora.py) and a training batch MIXES speakers: each sample carries ``speaker_ids[b]`` and its rows keep only
their own adapter's block of the low-rank term (``csm_lora_mask_rows``), so the frozen base GEMM — 99.9 % of the FLOPs —
runs once for all speakers.  A stack whose adapters are shared (``share_backbone`` / ``share_decoder``) holds a single
adapter that every speaker's rows train.

Kept: constructor arguments, ``initialize_trainers`` / ``prepare_optimizers`` / ``train(speaker_datasets, ...)`` /
``save_all_models`` / ``load_speaker_model`` / ``merge_speaker_models``, the ``shared/`` and ``speaker_{id}/`` output
layout and file names (multi_speaker_lora.py:326-343).  Every per-speaker file uses the single-adapter tensor names and
shapes, so ``CSMLoRATrainer.load_lora_weights`` reads it.  ``generate_sample`` needs the Mimi codec and the Llama
tokenizer (absent: no network) and raises.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch

from ..data.frames import collate_pinned
from ..models.model import Model
from .lora_trainer import CSMLoRATrainer
from .utils import compute_loss, setup_logger


class _SpeakerView:
    """What ``trainers[speaker_id]`` offers in the reference (a per-speaker CSMLoRATrainer), as a view of one speaker's
    adapter slices inside the shared engine."""

    def __init__(self, owner: "MultiSpeakerLoRATrainer", speaker_id: int):
        self.owner, self.speaker_id = owner, speaker_id
        self.best_loss = float("inf")

    @property
    def model(self):
        return self.owner.engine.model

    def get_lora_params(self) -> Dict[str, torch.Tensor]:
        return self.owner.speaker_state(self.speaker_id)

    def save_model(self, save_path: str, save_mode: str = "lora"):
        if save_mode != "lora":
            raise ValueError("a speaker view saves adapters only (save_mode='lora')")
        self.owner.save_speaker(self.speaker_id, save_path)

    def load_lora_weights(self, path: str):
        self.owner.load_speaker_model(self.speaker_id, path)


class MultiSpeakerLoRATrainer:
    def __init__(self, model_path: str, output_dir: str, speaker_ids: List[int], log_file: Optional[str] = None,
                 learning_rate: float = 1e-4, semantic_weight: float = 100.0, acoustic_weight: float = 1.0,
                 weight_decay: float = 0.01, lora_r: int = 8, lora_alpha: float = 16.0, lora_dropout: float = 0.0,
                 share_backbone: bool = True, share_decoder: bool = False,
                 target_modules: Optional[List[str]] = None, target_backbone_layers: Optional[List[int]] = None,
                 target_decoder_layers: Optional[List[int]] = None, lora_use_bias: bool = False, *,
                 model: Optional[Model] = None, device: str = "cuda"):
        if len(set(speaker_ids)) != len(speaker_ids) or not speaker_ids:
            raise ValueError("speaker_ids must be a non-empty list of distinct ids")
        self.model_path = model_path
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.speaker_ids = list(speaker_ids)
        self.index_of = {sid: k for k, sid in enumerate(self.speaker_ids)}       # speaker id -> adapter index
        self.logger = setup_logger("multi_speaker_lora_trainer",
                                   log_file or str(self.output_dir / "multi_speaker_training.log"))
        self.learning_rate, self.weight_decay = learning_rate, weight_decay
        self.semantic_weight, self.acoustic_weight = semantic_weight, acoustic_weight
        self.lora_r, self.lora_alpha, self.lora_dropout = lora_r, lora_alpha, lora_dropout
        self.target_modules = target_modules or ["q_proj", "v_proj"]
        self.target_backbone_layers, self.target_decoder_layers = target_backbone_layers, target_decoder_layers
        self.lora_use_bias = lora_use_bias
        self.share_backbone, self.share_decoder = share_backbone, share_decoder
        self._model, self._device = model, device
        self.logger.info(f"Initializing multi-speaker LoRA trainer for {len(speaker_ids)} speakers: {speaker_ids}; "
                         f"sharing backbone: {share_backbone}, decoder: {share_decoder}; r={lora_r}, alpha={lora_alpha}")
        self.trainers: Dict[int, _SpeakerView] = {}
        self.engine: Optional[CSMLoRATrainer] = None
        self.shared_trainer = None
        self.initialize_trainers()
        self.epoch, self.global_step, self.best_loss = 0, 0, float("inf")

    # ------------------------------------------------------------------ construction
    def initialize_trainers(self):
        """One engine (base model + all adapters) instead of one trainer per speaker (multi_speaker_lora.py:137-213)."""
        K = len(self.speaker_ids)
        adapters = {"backbone": 1 if self.share_backbone else K, "decoder": 1 if self.share_decoder else K}
        self.engine = CSMLoRATrainer(
            self.model_path, str(self.output_dir / "engine"), learning_rate=self.learning_rate,
            semantic_weight=self.semantic_weight, acoustic_weight=self.acoustic_weight, weight_decay=self.weight_decay,
            lora_r=self.lora_r, lora_alpha=self.lora_alpha, lora_dropout=self.lora_dropout,
            target_modules=self.target_modules, target_layers=self.target_backbone_layers,
            lora_use_bias=self.lora_use_bias, model=self._model, device=self._device, num_adapters=adapters,
            target_decoder_layers=self.target_decoder_layers)
        self.adapters = adapters
        for sid in self.speaker_ids:
            os.makedirs(self.output_dir / f"speaker_{sid}", exist_ok=True)
            self.trainers[sid] = _SpeakerView(self, sid)
        if self.share_backbone or self.share_decoder:
            os.makedirs(self.output_dir / "shared", exist_ok=True)
            self.shared_trainer = self.engine                   # the shared adapters live in the same engine
        self.logger.info(f"Created {len(self.trainers)} speaker adapter sets in one model")

    def set_model(self, model: Model):
        self._model = model
        self.engine.set_model(model)

    def prepare_optimizers(self):
        self.engine.prepare_optimizer()

    # ------------------------------------------------------------------ per-speaker adapter slices
    @staticmethod
    def _stack_of(name: str) -> str:
        return "backbone" if name.startswith("backbone.") else "decoder"

    def speaker_state(self, speaker_id: int, shared_scale: float = 1.0) -> Dict[str, torch.Tensor]:
        """The adapters speaker `speaker_id` runs through, in single-adapter names and shapes (lora_A [r, in],
        lora_B [out, r]): its own slices where a stack is per-speaker, the shared adapter elsewhere."""
        k, r = self.index_of[speaker_id], self.lora_r
        out = {}
        for name, p in self.engine.get_lora_params().items():
            K = self.adapters[self._stack_of(name)]
            t = p.detach()
            if K > 1:
                t = t[k * r:(k + 1) * r] if name.endswith("lora_A") else t[:, k * r:(k + 1) * r]
            elif shared_scale != 1.0 and name.endswith("lora_B"):
                t = t * shared_scale
            out[name] = t
        return out

    def shared_state(self) -> Dict[str, torch.Tensor]:
        return {n: p.detach() for n, p in self.engine.get_lora_params().items()
                if self.adapters[self._stack_of(n)] == 1}

    def _save(self, tensors: Dict[str, torch.Tensor], path: str, extra: Dict):
        from safetensors.torch import save_file
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        save_file({n: t.float().cpu().contiguous() for n, t in tensors.items()}, path)
        meta = {"lora_r": self.lora_r, "lora_alpha": self.lora_alpha, "lora_dropout": self.lora_dropout,
                "target_modules": self.target_modules, "target_layers": self.target_backbone_layers,
                "target_decoder_layers": self.target_decoder_layers, "lora_use_bias": self.lora_use_bias,
                "params_count": len(tensors), **extra}
        with open(path.replace(".safetensors", "_metadata.json"), "w") as f:
            json.dump(meta, f, indent=2)

    def save_speaker(self, speaker_id: int, path: str, shared_scale: float = 1.0):
        self._save(self.speaker_state(speaker_id, shared_scale), path,
                   {"speaker_id": speaker_id, "share_backbone": self.share_backbone,
                    "share_decoder": self.share_decoder})

    def save_all_models(self):
        """multi_speaker_lora.py:326-343: shared/shared_lora.safetensors + speaker_{id}/speaker_{id}_lora.safetensors."""
        if self.shared_trainer is not None:
            self._save(self.shared_state(), str(self.output_dir / "shared" / "shared_lora.safetensors"),
                       {"shared": True})
        for sid in self.speaker_ids:
            self.save_speaker(sid, str(self.output_dir / f"speaker_{sid}" / f"speaker_{sid}_lora.safetensors"))

    def load_speaker_model(self, speaker_id: int, checkpoint_path: str):
        """Loads a single-adapter file into speaker `speaker_id`'s slices (and into the shared adapters it contains)."""
        if speaker_id not in self.trainers:
            self.logger.error(f"No trainer found for speaker {speaker_id}")
            return
        from safetensors.torch import load_file
        tensors = load_file(checkpoint_path)
        k, r = self.index_of[speaker_id], self.lora_r
        params = self.engine.get_lora_params()
        missing = set(params) - set(tensors)
        if missing:
            raise KeyError(f"LoRA file lacks {len(missing)} tensors, e.g. {sorted(missing)[:3]}")
        with torch.no_grad():
            for name, p in params.items():
                src = tensors[name].to(device=p.device, dtype=p.dtype)
                if self.adapters[self._stack_of(name)] > 1:
                    (p[k * r:(k + 1) * r] if name.endswith("lora_A") else p[:, k * r:(k + 1) * r]).copy_(src)
                else:
                    p.copy_(src)
        self.engine._invalidate_optimizer_master()

    def merge_speaker_models(self, shared_weight: float = 0.5) -> Dict[int, str]:
        """multi_speaker_lora.py:378-437: per speaker, the shared components weighted by `shared_weight` next to the
        speaker's own (full weight); written as speaker_{id}_merged.safetensors.  Scaling an adapter's B factor scales
        its low-rank update."""
        merged: Dict[int, str] = {}
        if self.shared_trainer is None:
            self.logger.warning("No shared components to merge")
            return merged
        for sid in self.speaker_ids:
            path = str(self.output_dir / f"speaker_{sid}" / f"speaker_{sid}_merged.safetensors")
            self.save_speaker(sid, path, shared_scale=shared_weight)
            merged[sid] = path
        return merged

    def generate_sample(self, text: str, speaker_id: int, output_path: Optional[str] = None) -> str:
        raise NotImplementedError("generate_sample needs the Mimi codec and the Llama-3 tokenizer (moshi / HF hub: no "
                                  "network in this environment); Model.generate_frame produces the audio codes")

    # ------------------------------------------------------------------ training
    def mixed_batches(self, speaker_datasets: Dict[int, Tuple], batch_size: int, seed: int, split: int = 0,
                      shuffle: bool = True):
        """Batches that MIX speakers: the union of the speakers' datasets (split 0 = train, 1 = val), shuffled,
        partitioned over the data-parallel ranks, collated with a ``speaker_ids`` tensor (adapter index per sample)."""
        def pick(sets):
            if isinstance(sets, (tuple, list)) and len(sets) == 2 and not isinstance(sets[0], dict):
                return sets[split]
            return sets if split == 0 else None
        items = []
        for sid, sets in speaker_datasets.items():
            ds = pick(sets) if sid in self.index_of else None
            if ds is not None:
                items += [(sid, i) for i in range(len(ds))]
        g = torch.Generator().manual_seed(seed)
        order = torch.randperm(len(items), generator=g).tolist() if shuffle else list(range(len(items)))
        world, rank = self.engine.world, self.engine.rank
        if world > 1:
            order = order[:(len(order) // world) * world][rank::world]
        for i in range(0, len(order), batch_size):
            chunk = [items[j] for j in order[i:i + batch_size]]
            batch = collate_pinned([pick(speaker_datasets[sid])[j] for sid, j in chunk])
            batch["speaker_ids"] = torch.tensor([self.index_of[sid] for sid, _ in chunk], dtype=torch.int64)
            yield batch

    def train_step(self, batch, max_grad_norm: float = 1.0) -> torch.Tensor:
        """One optimiser step on a mixed-speaker batch (needs ``speaker_ids``: adapter index per sample)."""
        if "speaker_ids" not in batch:
            raise KeyError("a multi-speaker batch needs 'speaker_ids' (int64 [B], index into the trainer's speakers)")
        self.global_step += 1
        return self.engine.train_step(batch, max_grad_norm)

    def train(self, speaker_datasets: Dict[int, Tuple], batch_size: int = 2, epochs: int = 5, val_every: int = 100,
              save_every: int = 500, max_grad_norm: float = 1.0,
              resume_from: Optional[Dict[int, str]] = None) -> Dict[int, float]:
        for sid in self.speaker_ids:
            if sid not in speaker_datasets:
                self.logger.warning(f"No dataset provided for speaker {sid}")
        self.prepare_optimizers()
        if resume_from:
            for sid, path in resume_from.items():
                if sid in self.trainers:
                    self.load_speaker_model(sid, path)
        best: Dict[int, float] = {}

        def fold(vals):
            for sid, v in vals.items():
                best[sid] = min(best.get(sid, float("inf")), v)
                self.trainers[sid].best_loss = best[sid]
        self.engine.model.train()
        for epoch in range(self.epoch, self.epoch + epochs):
            losses = []
            for batch in self.mixed_batches(speaker_datasets, batch_size, seed=epoch):
                losses.append(self.train_step(batch, max_grad_norm))
                if self.global_step % val_every == 0:
                    fold(self.validate(speaker_datasets, batch_size))
                if self.global_step % save_every == 0 and self.engine.rank == 0:
                    self.save_all_models()
            avg = float(torch.stack(losses).mean()) if losses else float("nan")
            self.logger.info(f"Completed epoch {epoch + 1}: avg loss {avg:.6f}")
            self.epoch = epoch + 1
        fold(self.validate(speaker_datasets, batch_size))
        if self.engine.rank == 0:
            self.save_all_models()
        if best:
            self.best_loss = min(best.values())
        return best

    def validate(self, speaker_datasets: Dict[int, Tuple], batch_size: int = 2) -> Dict[int, float]:
        """Per-speaker validation loss (each speaker's val set through its own adapters)."""
        out: Dict[int, float] = {}
        eng = self.engine
        eng.model.eval()
        with torch.no_grad():
            for sid in self.speaker_ids:
                sets = speaker_datasets.get(sid)
                if not isinstance(sets, (tuple, list)) or len(sets) != 2 or sets[1] is None or len(sets[1]) == 0:
                    continue
                total, n = 0.0, 0
                for i in range(0, len(sets[1]), batch_size):
                    batch = collate_pinned([sets[1][j] for j in range(i, min(i + batch_size, len(sets[1])))])
                    batch["speaker_ids"] = torch.full((batch["input_tokens"].shape[0],), self.index_of[sid],
                                                      dtype=torch.int64)
                    b = eng._to_device(batch)
                    loss, _ = compute_loss(eng.model, b["input_tokens"], b["input_masks"], b["target_audio_tokens"],
                                           self.semantic_weight, self.acoustic_weight, frame_idx=b["frame_idx"],
                                           target_lengths=b.get("target_lengths"),
                                           mask_padded_targets=eng.mask_padded_targets, speaker_ids=b["speaker_ids"])
                    total += float(loss)
                    n += 1
                if n:
                    out[sid] = total / n
        eng.model.train()
        return out
