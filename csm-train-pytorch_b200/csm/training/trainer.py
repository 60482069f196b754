"""``CSMTrainer`` — drop-in for /root/reference/src/csm/training/trainer.py:26-394 on the B200 kernels.

Same constructor, ``prepare_optimizer`` (AdamW, lr x {backbone 0.1, decoder 1.0, embeddings 0.5, other 1.0} by name
substring, trainer.py:123-173), ``train`` loop (grad accumulation, clip_grad_norm_, checkpoints, trainer.py:175-357)
and ``_validate``.  Differences, all on purpose: no KV caches in training, the acoustic loss is real,
``loss.item()`` is read once per optimiser step instead of every micro-batch (trainer.py:266 forces a sync), and the
trainer is data-parallel when launched under torchrun (one process per GPU, NCCL all-reduce of gradients).
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Dict, Optional

import torch

from ..models.model import Model, ModelArgs
from . import dp
from .graph import GraphedStep
from .utils import batch_loss, compute_loss, load_checkpoint, save_checkpoint, setup_logger


def csm_1b_args() -> ModelArgs:                       # trainer.py:100-106
    return ModelArgs(backbone_flavor="llama-1B", decoder_flavor="llama-100M", text_vocab_size=128256,
                     audio_vocab_size=2051, audio_num_codebooks=32)


def collate_variable_length(batch):
    """Zero / False padding to the batch maximum — contract of data/training_data.py:379-408; the padded tensors are
    pinned on a GPU host so the trainer's H2D copies are asynchronous (csm/data/frames.py)."""
    from ..data.frames import collate_pinned
    return collate_pinned(batch)


def iterate_batches(dataset, batch_size: int, shuffle: bool, rank: int = 0, world: int = 1, seed: int = 0,
                    pack_to: Optional[int] = None, compact_vocab: Optional[int] = None):
    """`pack_to`: lay each batch's samples back to back in rows of at most that many frames (sequence packing,
    csm/data/frames.py::pack_samples) instead of zero-padding every sample to the batch maximum.
    `compact_vocab` (the model's audio_vocab_size): hand the inputs over in the compact device format
    (csm/data/frames.py::pack_tokens: int32 pre-offset rows + one mask word per frame, 140 instead of 297 bytes per
    frame) — results are bit-identical."""
    def out(b):
        if compact_vocab:
            from ..data.frames import compact_batch
            return compact_batch(b, compact_vocab)
        return b

    n = len(dataset)
    order = torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist() if shuffle else list(range(n))
    if world > 1:
        # every rank must run the same number of steps (each one ends in a collective): drop the ragged tail
        order = order[:(n // world) * world]
    order = order[rank::world]
    for i in range(0, len(order), batch_size):
        samples = [dataset[j] for j in order[i:i + batch_size]]
        if pack_to:
            from ..data.frames import pack_samples
            b = pack_samples(samples, pack_to, generator=torch.Generator().manual_seed(seed * 100003 + i))
            b.pop("sample_index")
            yield out(b)
        else:
            yield out(collate_variable_length(samples))


def make_optimizer(param_groups, lr: float, weight_decay: float):
    """AdamW of the reference trainers (trainer.py:166-173).  CUDA bf16 parameters get the two-pass multi-tensor
    clip + AdamW kernels (csm/training/optim.py); anything else (CPU unit tests of the host logic) stock torch AdamW."""
    params = [p for g in param_groups for p in g["params"]] if isinstance(param_groups[0], dict) else list(param_groups)
    if params and all(p.is_cuda and p.dtype == torch.bfloat16 for p in params):
        from .optim import FusedClipAdamW
        return FusedClipAdamW(param_groups, lr=lr, weight_decay=weight_decay)
    return torch.optim.AdamW(param_groups, lr=lr, weight_decay=weight_decay)


def clip_and_step(optimizer, params, max_grad_norm) -> None:
    """clip_grad_norm_(params, max_grad_norm) + optimizer.step() (trainer.py:271-276)."""
    from .optim import FusedClipAdamW
    if isinstance(optimizer, FusedClipAdamW):
        optimizer.step(max_grad_norm=max_grad_norm if max_grad_norm and max_grad_norm > 0 else 0.0)
        return
    if max_grad_norm and max_grad_norm > 0:
        torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
    optimizer.step()


class CSMTrainer:
    def __init__(self, model_path: str, output_dir: str, device: str = "cuda", log_file: Optional[str] = None,
                 learning_rate: float = 1e-5, backbone_lr_multiplier: float = 0.1,
                 decoder_lr_multiplier: float = 1.0, embedding_lr_multiplier: float = 0.5,
                 semantic_weight: float = 100.0, acoustic_weight: float = 1.0, weight_decay: float = 0.01):
        self.model_path = model_path
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.rank, self.world, self.local_rank = dp.init_distributed()
        self.device = f"cuda:{self.local_rank}" if (device == "cuda" and self.world > 1) else device
        self.logger = setup_logger("csm_trainer", log_file or str(self.output_dir / "training.log"))
        self.learning_rate = learning_rate
        self.backbone_lr_multiplier = backbone_lr_multiplier
        self.decoder_lr_multiplier = decoder_lr_multiplier
        self.embedding_lr_multiplier = embedding_lr_multiplier
        self.semantic_weight = semantic_weight
        self.acoustic_weight = acoustic_weight
        self.weight_decay = weight_decay
        self.decoder_frame_fraction = 1.0 / 16
        # batches from collate_pinned carry each sample's true target length: padded all-zero target frames are then
        # left out of the semantic mean as well (the reference averages over them, utils.py:101-105; set False for that)
        self.mask_padded_targets = True
        # sequence packing: rows of at most this many frames holding several samples each (None: pad like the reference)
        self.pack_sequences_to: Optional[int] = None
        self.compact_tokens: bool = False     # hand batches to the GPU as int32 rows + mask words (frames.pack_tokens)
        self.model = None
        self.optimizer = None
        self._sync = None
        self._graphed = None
        self._load_model()
        self.epoch = 0
        self.global_step = 0
        self.best_loss = float("inf")

    def _load_model(self):
        if not self.model_path:                       # trainer.py:92-95: model injected later (tests)
            self.logger.warning("Empty model path provided. Model will need to be set manually.")
            return
        if not self.model_path.endswith(".pt"):
            raise ValueError("csm_b200 loads .pt state dicts (trainer.py:97-111); hub loading needs network access")
        self.model = Model(csm_1b_args()).to(torch.bfloat16)
        state = torch.load(self.model_path, map_location="cpu")
        if "model" in state and "audio_head" not in state:
            state = state["model"]
        self.model.load_state_dict(state)
        self.model = self.model.to(self.device)

    def prepare_optimizer(self, freeze_backbone: bool = False, freeze_decoder: bool = False,
                          freeze_embeddings: bool = False):
        groups = {"backbone": [], "decoder": [], "embeddings": [], "other": []}
        for name, p in self.model.named_parameters():
            if freeze_backbone and "backbone" in name:
                p.requires_grad = False
            elif freeze_decoder and "decoder" in name:
                p.requires_grad = False
            elif freeze_embeddings and "embeddings" in name:
                p.requires_grad = False
            if p.requires_grad:
                key = "backbone" if "backbone" in name else "decoder" if "decoder" in name else \
                    "embeddings" if "embeddings" in name else "other"
                groups[key].append(p)
        total = sum(p.numel() for p in self.model.parameters() if p.requires_grad)
        self.logger.info(f"Training with {total:,} trainable parameters")
        lr = self.learning_rate
        param_groups = [g for g in (
            {"params": groups["backbone"], "lr": lr * self.backbone_lr_multiplier},
            {"params": groups["decoder"], "lr": lr * self.decoder_lr_multiplier},
            {"params": groups["embeddings"], "lr": lr * self.embedding_lr_multiplier},
            {"params": groups["other"], "lr": lr}) if g["params"]]
        self.optimizer = make_optimizer(param_groups, lr, self.weight_decay)
        trainable = [p for p in self.model.parameters() if p.requires_grad]
        self._sync = dp.GradSynchronizer(trainable, bucket_bytes=64 << 20 if total > (32 << 20) else None,
                                         sparse_rows=self.model.text_embeddings.weight,
                                         bucket_order=self._backward_order())
        self._sync.text_capacity_seq = self.model.backbone.max_seq_len
        # text-embedding gradient: gathered rows instead of a dense 525 MB all-reduce (dp.exchange_text_rows)
        self.model._text_grad_exchange = self._sync.exchange_text_rows if self._sync.sparse_param is not None else None
        sink = self._sync if self._sync.bucketed else None
        self.model.backbone._grad_sink = sink        # the stacks hand their layers' gradients over as they finish
        self.model.decoder._grad_sink = sink

    def _backward_order(self):
        """Parameters in the order the backward finishes their gradients: audio_head, decoder (final norm, layers
        last to first), projection, codebook0_head, backbone, then the embedding tables (last: the scatter at the very
        end of backward).  Inside a layer the projections of a fused GEMM are adjacent and in GEMM order (q|k|v,
        gate|up) so that their wgrad GEMM writes straight into the bucket."""
        m = self.model

        def stack(st):
            out = [st.norm.scale]
            for layer in reversed(list(st.layers)):
                a, f = layer.attn, layer.mlp
                out += [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.output_proj.weight, f.w1.weight,
                        f.w3.weight, f.w2.weight, layer.sa_norm.scale, layer.mlp_norm.scale]
            return out
        return [m.audio_head] + stack(m.decoder) + [m.projection.weight, m.codebook0_head.weight] + \
            stack(m.backbone) + [m.audio_embeddings.weight, m.text_embeddings.weight]

    def enable_cuda_graph(self, warmup: int = 3, max_grad_norm: float = 1.0) -> None:
        """``train_step`` (one micro-batch + optimiser step) replayed as one CUDA graph (training/graph.py)."""
        if self.optimizer is None:
            self.prepare_optimizer()

        def impl(b):
            loss, _ = batch_loss(self.model, b, self.semantic_weight, self.acoustic_weight, self.mask_padded_targets)
            loss.backward()
            self._sync.finish()
            clip_and_step(self.optimizer, [p for p in self.model.parameters() if p.requires_grad], max_grad_norm)
            self.optimizer.zero_grad(set_to_none=True)
            return loss.detach()
        self._graphed = GraphedStep(impl, self.device, warmup)

    def train_step(self, batch) -> torch.Tensor:
        """One micro-batch followed by one optimiser step (accumulation 1); graph-replayed when enabled."""
        if "frame_idx" not in batch:
            batch = dict(batch)
            batch["frame_idx"] = Model.select_frames(batch["input_masks"], batch["target_audio_tokens"].shape[1],
                                                     self.decoder_frame_fraction,
                                                     target_lengths=batch.get("target_lengths"))
        if getattr(self, "_graphed", None) is not None:
            self.global_step += 1
            return self._graphed(batch)
        loss = self.train_micro_batch(batch, 1)
        self.optimizer_step(1.0)
        return loss

    def _compact_vocab(self) -> Optional[int]:
        return int(self.model.args.audio_vocab_size) if self.compact_tokens else None

    def _to_device(self, batch) -> Dict[str, torch.Tensor]:
        if "frame_idx" not in batch:                  # A8: chosen on the host copy, before the H2D copy
            batch = dict(batch)
            batch["frame_idx"] = Model.select_frames(batch["input_masks"], batch["target_audio_tokens"].shape[1],
                                                     self.decoder_frame_fraction,
                                                     target_lengths=batch.get("target_lengths"))
        return {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}

    def train_micro_batch(self, batch, accumulation_steps: int = 1, last: bool = True) -> torch.Tensor:
        """`last` = this micro-batch closes the accumulation window (gradients are exchanged during its backward)."""
        self._sync.accumulating = not last
        b = self._to_device(batch)
        loss, _ = batch_loss(self.model, b, self.semantic_weight, self.acoustic_weight, self.mask_padded_targets)
        (loss / accumulation_steps).backward()
        return loss.detach()

    def optimizer_step(self, max_grad_norm: float = 1.0) -> None:
        self._sync.finish()
        clip_and_step(self.optimizer, [p for p in self.model.parameters() if p.requires_grad], max_grad_norm)
        self.optimizer.zero_grad(set_to_none=True)
        self.global_step += 1

    def train(self, train_dataset, val_dataset=None, batch_size: int = 2, accumulation_steps: int = 4,
              epochs: int = 5, val_every: int = 100, save_every: int = 500, max_grad_norm: float = 1.0,
              resume_from: Optional[str] = None):
        if self.optimizer is None:
            self.prepare_optimizer()
        if resume_from:
            meta = load_checkpoint(resume_from, self.model, self.optimizer, self.device)
            self.epoch, self.global_step, self.best_loss = meta["epoch"], meta["global_step"], meta["loss"]
        self.model.train()
        avg_loss = float("nan")
        for epoch in range(self.epoch, self.epoch + epochs):
            t0 = time.time()
            losses, window = [], []
            for bi, batch in enumerate(iterate_batches(train_dataset, batch_size, True, self.rank, self.world,
                                                       seed=epoch, pack_to=self.pack_sequences_to,
                                                       compact_vocab=self._compact_vocab())):
                closes = (bi + 1) % accumulation_steps == 0
                window.append(self.train_micro_batch(batch, accumulation_steps, last=closes))
                if closes:
                    self.optimizer_step(max_grad_norm)
                    step_loss = float(torch.stack(window).mean())      # one D2H read per optimiser step
                    losses.append(step_loss)
                    window = []
                    if val_dataset is not None and self.global_step % val_every == 0:
                        val = self._validate(val_dataset, batch_size)
                        self.logger.info(f"Epoch {epoch + 1}, Step {self.global_step}, Val Loss: {val:.6f}")
                        if val < self.best_loss and self.rank == 0:
                            self.best_loss = val
                            save_checkpoint(self.model, self.optimizer, epoch + 1, self.global_step, val,
                                            str(self.output_dir), "best")
                    if self.global_step % save_every == 0 and self.rank == 0:
                        save_checkpoint(self.model, self.optimizer, epoch + 1, self.global_step, step_loss,
                                        str(self.output_dir))
            avg_loss = sum(losses) / max(1, len(losses))
            self.logger.info(f"Epoch {epoch + 1} completed in {time.time() - t0:.2f}s, Avg Loss: {avg_loss:.6f}")
            if self.rank == 0:
                save_checkpoint(self.model, self.optimizer, epoch + 1, self.global_step, avg_loss,
                                str(self.output_dir), f"epoch_{epoch + 1}")
            self.epoch = epoch + 1
        if self.rank == 0:
            save_checkpoint(self.model, self.optimizer, self.epoch, self.global_step, avg_loss,
                            str(self.output_dir), "final")
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
