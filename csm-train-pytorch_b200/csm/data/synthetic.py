"""Synthetic training batches with the tensor contract of the reference's data pipeline
(/root/reference/src/csm/data/training_data.py:245-302 ``CSMDataset.__getitem__`` and :379-408
``collate_variable_length``): ``input_tokens`` int64 [B,S,33], ``input_masks`` bool [B,S,33],
``target_audio_tokens`` int64 [B,S,32]; frame layout of generator.py:77-130 ([text frames | audio frames], masks on
column 32 vs columns 0..31) plus zero/False padding frames.  Shapes and seeding follow SURVEY.md §8(d)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch


def synthetic_batch(text_vocab: int, audio_vocab: int, codebooks: int, B: int, S: int, seed: int = 1234,
                    fraction: float = 1.0 / 16, s_text: Optional[int] = None) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    C, V, Vt = codebooks, audio_vocab, text_vocab
    s_text = min(64, max(1, S // 4)) if s_text is None else s_text
    s_pad = max(1, S // 16)
    a0, a1 = s_text, S - s_pad
    tokens = torch.zeros(B, S, C + 1, dtype=torch.int64)
    mask = torch.zeros(B, S, C + 1, dtype=torch.bool)
    tokens[:, :a0, C] = torch.randint(0, Vt, (B, a0), generator=g)
    mask[:, :a0, C] = True
    tokens[:, a0:a1, :C] = torch.randint(0, V, (B, a1 - a0, C), generator=g)
    mask[:, a0:a1, :C] = True
    targets = torch.randint(0, V, (B, S, C), generator=g)
    n_keep = max(1, math.ceil((a1 - a0) * fraction))
    fi = []
    for b in range(B):
        perm = torch.randperm(a1 - a0, generator=g)[:n_keep].sort().values + a0
        fi.append(torch.stack([torch.full_like(perm, b), perm], dim=1))
    return {"input_tokens": tokens, "input_masks": mask, "target_audio_tokens": targets,
            "frame_idx": torch.cat(fi, 0)}
