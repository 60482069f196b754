"""CPU tests of the data contract helpers (csm/data/frames.py) against the reference's layout rules."""
import os

import pytest
import torch

from csm.data import frames as F

REF = "/root/reference/src/csm"


def test_text_and_audio_frame_layout():
    tok, msk = F.text_frames([5, 7, 9])
    assert tok.shape == (3, 33) and tok[:, 32].tolist() == [5, 7, 9] and int(tok[:, :32].abs().sum()) == 0
    assert msk[:, 32].all() and not msk[:, :32].any()                      # generator.py:91-95
    codes = torch.arange(32 * 4).view(32, 4) % 2051
    at, am = F.audio_frames(codes)
    assert at.shape == (5, 33)                                              # + EOS frame (generator.py:117-119)
    assert torch.equal(at[:4, :32], codes.t()) and int(at[4].sum()) == 0 and int(at[:, 32].sum()) == 0
    assert am[:, :32].all() and not am[:, 32].any()
    at2, _ = F.audio_frames(codes, add_eos=False)
    assert at2.shape == (4, 33)
    with pytest.raises(ValueError):
        F.audio_frames(torch.zeros(3, dtype=torch.long))


def test_build_sample_matches_reference_concatenation_and_truncation():
    ctx_codes = torch.randint(0, 2051, (32, 6), generator=torch.Generator().manual_seed(0))
    tgt_codes = torch.randint(0, 2051, (32, 9), generator=torch.Generator().manual_seed(1))
    s = F.build_sample([([1, 2, 3], ctx_codes)], [4, 5], tgt_codes)
    # context text (3) + context audio (6 + EOS) + target text (2)
    assert s["input_tokens"].shape == (3 + 7 + 2, 33) and s["input_masks"].shape == (12, 33)
    assert s["input_tokens"][-2:, 32].tolist() == [4, 5]
    assert s["target_audio_tokens"].shape == (9, 32) and torch.equal(s["target_audio_tokens"], tgt_codes.t())
    # truncation: cut from the beginning, keep min(max_seq_len, target text length) frames (training_data.py:289-294)
    t = F.build_sample([([1, 2, 3], ctx_codes)], [4, 5], tgt_codes, max_seq_len=8)
    assert t["input_tokens"].shape[0] == 2 and t["input_tokens"][:, 32].tolist() == [4, 5]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_collate_equals_reference_collate():
    import importlib.util
    import sys
    import types
    # the reference's collate lives in a module whose package imports need moshi etc.: load the function's source only
    src = open(os.path.join(REF, "data", "training_data.py")).read()
    start = src.index("def collate_variable_length(batch):")
    ns = {"torch": torch}
    exec(src[start:], ns)                                                   # noqa: S102 (reference code, test only)
    ref_collate = ns["collate_variable_length"]
    g = torch.Generator().manual_seed(3)
    batch = []
    for n, t in [(5, 7), (9, 4), (2, 9)]:
        tok = torch.randint(0, 100, (n, 33), generator=g)
        batch.append({"input_tokens": tok, "input_masks": torch.rand(n, 33, generator=g) > 0.5,
                      "target_audio_tokens": torch.randint(0, 100, (t, 32), generator=g)})
    ours, ref = F.collate_pinned(batch, pin=False), ref_collate(batch)
    assert torch.equal(ours["input_tokens"], ref["input_tokens"])
    assert torch.equal(ours["input_masks"], ref["input_masks"])
    T = ref["target_audio_tokens"].shape[1]
    assert torch.equal(ours["target_audio_tokens"][:, :T], ref["target_audio_tokens"])
    assert int(ours["target_audio_tokens"][:, T:].abs().sum()) == 0          # only zero padding beyond the reference's
    p = F.collate_pinned(batch, pin=False, pad_to_multiple=8)
    assert p["input_tokens"].shape[1] == 16 and not p["input_masks"][:, 9:].any()


def test_length_bucketing_cuts_padding_and_covers_every_sample():
    g = torch.Generator().manual_seed(0)
    lengths = (torch.rand(1000, generator=g) ** 3 * 2000 + 50).long().tolist()     # long-tailed
    naive = [list(range(i, min(i + 8, 1000))) for i in range(0, 1000, 8)]
    bucketed = F.length_bucketed_order(lengths, 8, seed=1)
    assert sorted(j for b in bucketed for j in b) == list(range(1000))
    assert all(1 <= len(b) <= 8 for b in bucketed)
    assert F.padding_fraction(lengths, bucketed) < 0.5 * F.padding_fraction(lengths, naive)
    assert bucketed != F.length_bucketed_order(lengths, 8, seed=2)                   # the epoch order stays random


def test_pack_samples_layout_segments_targets_and_frames():
    """Sequence packing host logic: first-fit-decreasing rows, per-position segment tables, targets aligned with their
    sample's positions, the semantic mask (p < min(len - 1, target rows)), per-sample decoder frames."""
    import math
    import torch
    from csm.data import frames as F
    lens, tlens = [100, 156, 60, 250, 30], [100, 120, 60, 250, 10]
    samples = [{"input_tokens": torch.full((n, 33), i + 1), "input_masks": torch.ones(n, 33, dtype=torch.bool),
                "target_audio_tokens": torch.full((t, 32), 10 * (i + 1))} for i, (n, t) in enumerate(zip(lens, tlens))]
    b = F.pack_samples(samples, 256, pad_to_multiple=128, generator=torch.Generator().manual_seed(0), pin=False)
    R, S = b["input_tokens"].shape[:2]
    assert S == 256 and R == 3 and b["segment_starts"].dtype == torch.int32
    own, ss, se = b["sample_index"], b["segment_starts"], b["segment_ends"]
    for j, (n, t) in enumerate(zip(lens, tlens)):
        where = torch.nonzero(own == j)
        assert where.shape[0] == n and len(set(where[:, 0].tolist())) == 1          # one contiguous run in one row
        r, p0 = int(where[0, 0]), int(where[0, 1])
        assert where[:, 1].tolist() == list(range(p0, p0 + n))
        assert (ss[r, p0:p0 + n] == p0).all() and (se[r, p0:p0 + n] == p0 + n).all()
        assert (b["input_tokens"][r, p0:p0 + n] == j + 1).all() and b["input_masks"][r, p0:p0 + n].all()
        tt = min(n, t)
        assert (b["target_audio_tokens"][r, p0:p0 + tt] == 10 * (j + 1)).all()
        valid = min(n - 1, tt)
        assert b["target_mask"][r, p0:p0 + valid].all() and not b["target_mask"][r, p0 + valid:p0 + n].any()
        mine = b["frame_idx"][(own[b["frame_idx"][:, 0], b["frame_idx"][:, 1]] == j)]
        assert mine.shape[0] == max(1, math.ceil(valid / 16)) and (mine[:, 1] < p0 + valid).all()
    pad = own < 0
    assert (ss[pad] == torch.arange(S).repeat(R, 1)[pad]).all() and (se[pad] == ss[pad] + 1).all()
    assert not b["input_masks"][pad].any() and not b["target_mask"][pad].any()
    assert abs(F.packing_efficiency(b) - sum(lens) / (R * S)) < 1e-6
    import pytest
    with pytest.raises(ValueError):
        F.pack_samples(samples, 200, pin=False)


def test_pack_tokens_compact_device_format_round_trip():
    """int32 pre-offset table rows + one 33-bit mask word per frame; the offsets are the ones
    Model._embed_tokens adds (model.py:209-212)."""
    g = torch.Generator().manual_seed(3)
    C, V, Vt = 32, 2051, 128256
    tok = torch.zeros(2, 7, C + 1, dtype=torch.int64)
    tok[..., :C] = torch.randint(0, V, (2, 7, C), generator=g)
    tok[..., C] = torch.randint(0, Vt, (2, 7), generator=g)
    msk = torch.rand(2, 7, C + 1, generator=g) < 0.5
    rows, bits = F.pack_tokens(tok, msk, V, pin=False)
    assert rows.dtype == torch.int32 and rows.shape == tok.shape
    assert bits.dtype == torch.int64 and bits.shape == (2, 7)
    assert torch.equal(rows[..., :C].long(), tok[..., :C] + V * torch.arange(C))
    assert torch.equal(rows[..., C].long(), tok[..., C])
    for c in (0, 5, 31, 32):
        assert torch.equal(((bits >> c) & 1).bool(), msk[..., c])
    assert int(bits.max()) < 2 ** 33
    t2, m2 = F.unpack_tokens(rows, bits, V)
    assert torch.equal(t2, tok) and torch.equal(m2, msk)
    # bytes per frame: 33 * 4 + 8 against 33 * 8 + 33
    assert rows[0, 0].numel() * 4 + 8 == 140
    b = {"input_tokens": tok, "input_masks": msk, "target_audio_tokens": torch.zeros(2, 7, C, dtype=torch.int64)}
    cb = F.compact_batch(b, V, pin=False)
    assert cb["input_tokens"].dtype == torch.int32 and cb["input_masks"].dtype == torch.int64
    assert cb["target_audio_tokens"] is b["target_audio_tokens"]
    with pytest.raises(ValueError):
        F.pack_tokens(torch.full((1, 1, 33), 2 ** 31, dtype=torch.int64), torch.ones(1, 1, 33, dtype=torch.bool), V)
