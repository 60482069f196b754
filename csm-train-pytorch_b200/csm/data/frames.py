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

import math
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


def pack_tokens(tokens: torch.Tensor, masks: torch.Tensor, audio_vocab: int,
                pin: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """The compact device format: tokens int64 [..., C+1] + bool mask [..., C+1]  ->
    rows int32 [..., C+1] of PRE-OFFSET embedding-table rows (audio column c: id + c * audio_vocab — the index
    ``Model._embed_tokens`` computes, model.py:209-212 — text column: the id) and one int64 word per frame whose bit c is
    the mask of column c.  140 instead of 297 bytes per frame; ``Model.forward`` takes the pair in place of
    (tokens, tokens_mask) and produces bit-identical results (csm_embed_gather_sum_packed_fwd / _bwd)."""
    W = tokens.shape[-1]
    C = W - 1
    if W > 63:
        raise ValueError("pack_tokens: at most 63 columns fit the mask word")
    off = torch.arange(W, dtype=torch.int64) * int(audio_vocab)
    off[C] = 0
    rows64 = tokens.to(torch.int64) + off
    if int(rows64.max()) >= 2 ** 31 or int(rows64.min()) < -(2 ** 31):
        raise ValueError("pack_tokens: table rows do not fit int32")
    bits64 = (masks.to(torch.int64) << torch.arange(W, dtype=torch.int64)).sum(dim=-1)
    pin = pin and torch.cuda.is_available()
    rows = torch.empty(rows64.shape, dtype=torch.int32, pin_memory=pin)
    bits = torch.empty(bits64.shape, dtype=torch.int64, pin_memory=pin)
    rows.copy_(rows64)
    bits.copy_(bits64)
    return rows, bits


def unpack_tokens(rows: torch.Tensor, mask_bits: torch.Tensor, audio_vocab: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of pack_tokens: (tokens int64 [..., C+1], masks bool [..., C+1])."""
    W = rows.shape[-1]
    off = torch.arange(W, dtype=torch.int64, device=rows.device) * int(audio_vocab)
    off[W - 1] = 0
    tokens = rows.to(torch.int64) - off
    masks = ((mask_bits.unsqueeze(-1) >> torch.arange(W, dtype=torch.int64, device=rows.device)) & 1).bool()
    return tokens, masks


def compact_batch(batch: Dict[str, torch.Tensor], audio_vocab: int, pin: bool = True) -> Dict[str, torch.Tensor]:
    """A copy of a collated (or packed) host batch whose ``input_tokens`` / ``input_masks`` are in the compact device
    format (pack_tokens); every other key is passed through.  The trainers and ``Model.forward`` take it as is."""
    out = dict(batch)
    out["input_tokens"], out["input_masks"] = pack_tokens(batch["input_tokens"], batch["input_masks"], audio_vocab, pin)
    return out


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


def pack_samples(samples: List[Dict[str, torch.Tensor]], max_len: int, pad_to_multiple: int = 128,
                 fraction: float = 1.0 / 16, generator: Optional[torch.Generator] = None,
                 pin: bool = True) -> Dict[str, torch.Tensor]:
    """Sequence packing: variable-length samples laid back to back in rows of at most ``max_len``
    frames instead of one zero-padded row per sample (the reference's collate pads every sample to the batch maximum,
    training_data.py:379-408, and the padding frames cost full backbone FLOPs).  First-fit-decreasing bin packing; the
    row length is the longest row rounded up to ``pad_to_multiple`` (whole attention tiles).

    Returns the usual keys for R rows of S frames — ``input_tokens`` [R,S,C+1], ``input_masks`` [R,S,C+1],
    ``target_audio_tokens`` [R,S,C] (each sample's target rows at its own positions: position p of a sample pairs with
    its target row p, utils.py:101-104) — plus what the packed kernels need:
      ``segment_starts`` / ``segment_ends`` int32 [R,S]   first / last+1 position of the sample position i belongs to
                                                          (a padding frame is its own one-frame segment)
      ``target_mask``  bool [R,S]    positions of the semantic term: p < min(len - 1, target rows) of each sample
      ``frame_idx``    int64 [N,2]   decoder frames (row, position): ceil(n/16) of each sample's valid positions
      ``sample_index`` int64 [R,S]   which input sample a position came from (-1: padding)"""
    if not samples:
        raise ValueError("pack_samples: no samples")
    lens = [int(b["input_tokens"].shape[0]) for b in samples]
    if max(lens) > max_len:
        raise ValueError(f"pack_samples: a sample of {max(lens)} frames does not fit max_len={max_len}")
    order = sorted(range(len(samples)), key=lambda j: -lens[j])
    rows: List[List[int]] = []
    used: List[int] = []
    for j in order:
        for r in range(len(rows)):
            if used[r] + lens[j] <= max_len:
                rows[r].append(j)
                used[r] += lens[j]
                break
        else:
            rows.append([j])
            used.append(lens[j])
    S = (max(used) + pad_to_multiple - 1) // pad_to_multiple * pad_to_multiple
    R = len(rows)
    W = samples[0]["input_tokens"].shape[1]
    C = samples[0]["target_audio_tokens"].shape[1]
    pin = pin and torch.cuda.is_available()
    tok = torch.zeros(R, S, W, dtype=torch.int64, pin_memory=pin)
    msk = torch.zeros(R, S, W, dtype=torch.bool, pin_memory=pin)
    tgt = torch.zeros(R, S, C, dtype=torch.int64, pin_memory=pin)
    idx = torch.arange(S, dtype=torch.int32)
    ss = idx.repeat(R, 1).contiguous()
    se = (idx + 1).repeat(R, 1).contiguous()
    tmask = torch.zeros(R, S, dtype=torch.bool)
    owner = torch.full((R, S), -1, dtype=torch.int64)
    frames = []
    for r, members in enumerate(rows):
        off = 0
        for j in members:
            b, n = samples[j], lens[j]
            t = min(int(b["target_audio_tokens"].shape[0]), n)
            tok[r, off:off + n] = b["input_tokens"]
            msk[r, off:off + n] = b["input_masks"].bool()
            tgt[r, off:off + t] = b["target_audio_tokens"][:t]
            ss[r, off:off + n] = off
            se[r, off:off + n] = off + n
            owner[r, off:off + n] = j
            valid = max(0, min(n - 1, t))
            tmask[r, off:off + valid] = True
            if valid > 0:
                keep = max(1, math.ceil(valid * fraction))
                sel = torch.randperm(valid, generator=generator)[:keep].sort().values + off
                frames.append(torch.stack([torch.full_like(sel, r), sel], dim=1))
            off += n
    fidx = torch.cat(frames, 0) if frames else torch.zeros(0, 2, dtype=torch.int64)
    return {"input_tokens": tok, "input_masks": msk, "target_audio_tokens": tgt, "segment_starts": ss,
            "segment_ends": se, "target_mask": tmask, "frame_idx": fidx, "sample_index": owner}


def packing_efficiency(batch: Dict[str, torch.Tensor]) -> float:
    """Fraction of the packed frames that are real (not padding)."""
    return float((batch["sample_index"] >= 0).float().mean())


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
