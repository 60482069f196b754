"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute calls),
the Python mirror keeps the reference's interface, data contract and persistence formats."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "csm_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from csm import _lib
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/csm_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert lib.csm_abi_version() == 2
    assert isinstance(lib.csm_last_error(), bytes)
    # sizing helpers are pure host functions
    assert lib.csm_attn_bwd_workspace_bytes(2, 64, 4, 1, 64) >= 2 * 4 * 64 * 4
    assert lib.csm_linear_ce_workspace_bytes(128, 2051, 1024, 31) > 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from csm import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_model_state_dict_keys_match_reference_contract():
    from csm.models.model import Model, ModelArgs
    from oracle import csm_oracle as O
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    om = O.OracleModel(O.cfg_tiny())             # key set verified against the reference in test_oracle.py
    assert set(m.state_dict()) == set(om.state_dict())
    for k, v in om.state_dict().items():
        assert m.state_dict()[k].shape == v.shape, k
    assert m.audio_head.shape == (31, 16, 200)
    # CSM-1B flavours: dims of model.py:11-42 without allocating them
    from csm.models import model as mm
    with torch.device("meta"):
        bb, dec = mm.llama3_2_1B(), mm.llama3_2_100M()
    assert (len(bb.layers), bb.num_heads, bb.num_kv_heads, bb.embed_dim, bb.head_dim) == (16, 32, 8, 2048, 64)
    assert (len(dec.layers), dec.num_heads, dec.num_kv_heads, dec.embed_dim, dec.head_dim) == (4, 8, 2, 1024, 128)
    assert bb.layers[0].mlp.w1.weight.shape == (8192, 2048)
    n = sum(p.numel() for p in bb.parameters()) + sum(p.numel() for p in dec.parameters())
    assert abs(n - 1.084e9) < 0.01e9


def test_reference_helpers_present():
    from csm.models import model as mm
    m = mm.Model(mm.ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    tok = torch.randint(0, 100, (2, 5, 33))
    e = m._embed_tokens(tok)
    assert e.shape == (2, 5, 33, 32)
    assert torch.equal(e[:, :, 3], m.audio_embeddings(tok[:, :, 3] + 3 * 200))
    assert torch.equal(m._embed_audio(3, tok[:, :, 3]), e[:, :, 3])
    mask = mm._create_causal_mask(6, torch.device("cpu"))
    pos = torch.arange(5).unsqueeze(0).repeat(2, 1)
    assert torch.equal(m._index_causal_mask(mask, pos), mask[pos])          # method form (fixes utils.py:90)
    assert torch.equal(mm._index_causal_mask(mask, pos)[0, 2], torch.tensor([1, 1, 1, 0, 0, 0], dtype=torch.bool))


def test_api_signatures_match_reference():
    from csm.training.lora_trainer import CSMLoRATrainer
    from csm.training.trainer import CSMTrainer
    from csm.training.utils import compute_loss, load_checkpoint, save_checkpoint
    p = list(inspect.signature(compute_loss).parameters)
    assert p[:6] == ["model", "input_tokens", "input_masks", "target_audio_tokens", "semantic_weight",
                     "acoustic_weight"]                                        # utils.py:56-63
    sig = inspect.signature(compute_loss)
    assert sig.parameters["semantic_weight"].default == 100.0 and sig.parameters["acoustic_weight"].default == 1.0
    p = list(inspect.signature(CSMTrainer.__init__).parameters)[1:]
    assert p == ["model_path", "output_dir", "device", "log_file", "learning_rate", "backbone_lr_multiplier",
                 "decoder_lr_multiplier", "embedding_lr_multiplier", "semantic_weight", "acoustic_weight",
                 "weight_decay"]                                               # trainer.py:29-42
    p = list(inspect.signature(CSMTrainer.train).parameters)[1:]
    assert p == ["train_dataset", "val_dataset", "batch_size", "accumulation_steps", "epochs", "val_every",
                 "save_every", "max_grad_norm", "resume_from"]                 # trainer.py:175-186
    p = list(inspect.signature(CSMLoRATrainer.__init__).parameters)[1:14]
    assert p == ["model_path", "output_dir", "log_file", "learning_rate", "semantic_weight", "acoustic_weight",
                 "weight_decay", "lora_r", "lora_alpha", "lora_dropout", "target_modules", "target_layers",
                 "lora_use_bias"]                                              # lora_trainer.py:32-48
    d = inspect.signature(CSMLoRATrainer.__init__).parameters
    assert (d["learning_rate"].default, d["lora_r"].default, d["lora_alpha"].default) == (1e-4, 8, 16.0)
    p = list(inspect.signature(CSMLoRATrainer.train).parameters)[1:]
    assert p == ["train_dataset", "val_dataset", "batch_size", "epochs", "val_every", "save_every", "max_grad_norm",
                 "resume_from"]                                                # mlx_trainer.py:733-743
    for name in ("prepare_optimizer", "train_step", "save_model", "load_lora_weights"):
        assert callable(getattr(CSMLoRATrainer, name))
    assert list(inspect.signature(save_checkpoint).parameters) == ["model", "optimizer", "epoch", "global_step",
                                                                   "loss", "save_dir", "name"]
    assert list(inspect.signature(load_checkpoint).parameters)[:3] == ["checkpoint_path", "model", "optimizer"]


def test_multi_speaker_trainer_api_and_adapter_layout():
    """MultiSpeakerLoRATrainer keeps the reference's constructor / method surface (multi_speaker_lora.py:36-60,
    233-243, 326-437); apply_lora(num_adapters=...) lays K adapters side by side per projection."""
    from csm.models import lora
    from csm.models.model import Model, ModelArgs
    from csm.training.multi_speaker_lora import MultiSpeakerLoRATrainer as T
    p = list(inspect.signature(T.__init__).parameters)[1:19]
    assert p == ["model_path", "output_dir", "speaker_ids", "log_file", "learning_rate", "semantic_weight",
                 "acoustic_weight", "weight_decay", "lora_r", "lora_alpha", "lora_dropout", "share_backbone",
                 "share_decoder", "target_modules", "target_backbone_layers", "target_decoder_layers",
                 "lora_use_bias", "model"]
    d = inspect.signature(T.__init__).parameters
    assert (d["share_backbone"].default, d["share_decoder"].default, d["lora_r"].default) == (True, False, 8)
    assert list(inspect.signature(T.train).parameters)[1:] == ["speaker_datasets", "batch_size", "epochs", "val_every",
                                                               "save_every", "max_grad_norm", "resume_from"]
    for name in ("initialize_trainers", "prepare_optimizers", "save_all_models", "load_speaker_model",
                 "generate_sample", "merge_speaker_models"):
        assert callable(getattr(T, name))
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    lora.apply_lora(m, r=4, num_adapters={"backbone": 1, "decoder": 3}, seed=0)
    q = m.decoder.layers[0].attn.q_proj
    assert q.lora_A.shape == (12, 16) and q.lora_B.shape == (16, 12) and q.lora_adapters == 3 and q.lora_r == 4
    assert m.backbone.layers[0].attn.q_proj.lora_A.shape == (4, 32) and m.backbone.lora_adapters == 1
    a1, b1 = lora.adapter_slices(q, 1)
    assert a1.shape == (4, 16) and b1.shape == (16, 4) and a1.data_ptr() == q.lora_A[4:8].data_ptr()
    with pytest.raises(RuntimeError):
        lora.merge_lora(m)                                    # no single merged weight with several adapters
    lora.apply_lora(m, r=4, target_layers=[0], target_decoder_layers=[])       # separate layer filters per stack
    assert not any("decoder" in n for n, p in m.named_parameters() if p.requires_grad)


def test_lora_names_counts_and_freezing():
    from csm.models import lora
    from csm.models.model import Model, ModelArgs
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    names = lora.apply_lora(m, r=8, alpha=16.0, seed=0)
    assert len(names) == 2 * (2 + 1)                                          # q,v x (2 backbone + 1 decoder layers)
    sd = lora.lora_state_dict(m)
    assert "backbone.layers.0.attn.q_proj.lora_A" in sd and "decoder.layers.0.attn.v_proj.lora_B" in sd
    assert sd["backbone.layers.0.attn.q_proj.lora_A"].shape == (8, 32)
    assert sd["backbone.layers.0.attn.q_proj.lora_B"].shape == (32, 8)
    assert float(sd["backbone.layers.0.attn.q_proj.lora_B"].abs().sum()) == 0.0   # B = 0 init (lora.py:66)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert sorted(trainable) == sorted(sd)
    assert m.backbone.layers[0].attn.q_proj.lora_scaling == 2.0               # alpha / r
    # CSM-1B r=8 q/v: 958 464 trainable parameters
    n = 16 * (8 * 2048 + 2048 * 8 + 8 * 2048 + 512 * 8) + 4 * (8 * 1024 + 1024 * 8 + 8 * 1024 + 256 * 8)
    assert n == 958_464
    with pytest.raises(ValueError):
        lora.apply_lora(m, target_modules=["nope"])
    m2 = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    names = lora.apply_lora(m2, r=4, target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj",
                                                     "down_proj"], target_layers=[0])
    assert len(names) == 7 * 2


def test_synthetic_batch_contract_and_frame_selection():
    from csm.data.synthetic import synthetic_batch
    from csm.models.model import Model
    from oracle import csm_oracle as O
    b = synthetic_batch(1000, 200, 32, 2, 64, seed=7)
    ob = O.synthetic_batch(O.cfg_tiny(), 2, 64, seed=7)
    for k in b:
        assert torch.equal(b[k], ob[k]), k                                    # product generator == oracle generator
    tok, msk, tgt, fi = b["input_tokens"], b["input_masks"], b["target_audio_tokens"], b["frame_idx"]
    assert tok.shape == (2, 64, 33) and tok.dtype == torch.int64
    assert msk.shape == (2, 64, 33) and msk.dtype == torch.bool
    assert tgt.shape == (2, 64, 32)
    assert not msk[:, -4:].any() and (tok[:, -4:] == 0).all()                 # padding frames: zeros / False
    assert msk[:, :16, 32].all() and not msk[:, :16, :32].any()               # text frames
    assert fi.shape[1] == 2 and int(fi[:, 1].max()) < 63
    # A8: ceil(T_b / 16) positions per sample out of the sample's REAL target frames, p < min(S-1, T_b) (ADVICE r1:
    # never a frame from the zero-padded tail of the targets, whatever the input mask says)
    sel = Model.select_frames(msk, 64, 1 / 16, torch.Generator().manual_seed(0))
    assert sel.shape == (2 * 4, 2) and int(sel[:, 1].max()) < 63              # ceil(63 / 16) = 4 per sample
    lens = torch.tensor([20, 3])
    sel = Model.select_frames(msk, 64, 1 / 16, torch.Generator().manual_seed(0), target_lengths=lens)
    assert sel[:, 0].tolist() == [0, 0, 1]                                    # ceil(20/16) = 2, ceil(3/16) = 1
    for bb, p in sel.tolist():
        assert p < int(lens[bb])
    sel = Model.select_frames(msk, 64, 1.0, target_lengths=torch.tensor([0, 100]))
    assert sel[:, 0].tolist() == [1] * 63 and sel[:, 1].tolist() == list(range(63))   # no targets -> no frames; p < S-1
    assert Model.select_frames(torch.zeros(1, 1, 33, dtype=torch.bool), 8).shape == (0, 2)


def test_collate_pads_like_reference():
    from csm.training.trainer import collate_variable_length, iterate_batches
    items = [{"input_tokens": torch.ones(s, 33, dtype=torch.int64), "input_masks": torch.ones(s, 33, dtype=torch.bool),
              "target_audio_tokens": torch.ones(s, 32, dtype=torch.int64)} for s in (3, 5)]
    out = collate_variable_length(items)
    assert out["input_tokens"].shape == (2, 5, 33)
    assert out["input_tokens"][0, 3:].sum() == 0 and not out["input_masks"][0, 3:].any()   # zero / False padding
    assert out["target_audio_tokens"].shape == (2, 5, 32)
    r0 = [b["input_tokens"].shape[0] for b in iterate_batches(items * 4, 2, True, rank=0, world=2, seed=1)]
    r1 = [b["input_tokens"].shape[0] for b in iterate_batches(items * 4, 2, True, rank=1, world=2, seed=1)]
    assert sum(r0) + sum(r1) == 8                                              # ranks partition the dataset


def test_checkpoint_roundtrip_format(tmp_path):
    from csm.models.model import Model, ModelArgs
    from csm.training.utils import load_checkpoint, save_checkpoint
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    with torch.no_grad():
        m.audio_head.normal_()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    path = save_checkpoint(m, opt, 2, 17, 1.5, str(tmp_path))
    assert os.path.basename(path) == "checkpoint_epoch2_step17.pt"             # utils.py:545
    assert os.path.exists(tmp_path / "checkpoint_latest.pt")
    ck = torch.load(path)
    assert set(ck) == {"model", "optimizer", "epoch", "global_step", "loss"}
    m2 = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    meta = load_checkpoint(path, m2, None, "cpu")
    assert meta == {"epoch": 2, "global_step": 17, "loss": 1.5}
    assert torch.equal(m2.audio_head, m.audio_head)


def test_lora_trainer_dropout_and_bias_options(tmp_path):
    """lora_dropout / lora_use_bias (lora_trainer.py:42-47, lora.py:87-90,101-102) are constructor arguments of the API:
    accepted, validated, and they shape the adapters (a ``lora_bias`` [out] per projection; the stack's dropout rate)."""
    from csm.models import lora
    from csm.models.model import Model, ModelArgs
    from csm.training.lora_trainer import CSMLoRATrainer
    with pytest.raises(ValueError):
        CSMLoRATrainer("", str(tmp_path), lora_dropout=1.5)
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32))
    t = CSMLoRATrainer("", str(tmp_path), lora_dropout=0.1, lora_use_bias=True, model=m, device="cpu")
    names = set(t.get_lora_params())
    assert "backbone.layers.0.attn.q_proj.lora_bias" in names and "decoder.layers.0.attn.v_proj.lora_bias" in names
    assert m.backbone.layers[0].attn.q_proj.lora_bias.shape == (32,) and m.backbone.lora_dropout == 0.1
    assert float(m.backbone.layers[0].attn.q_proj.lora_bias.abs().sum()) == 0.0          # zeros, lora.py:66
    assert set(lora.lora_state_dict(m)) == names
    with pytest.raises(ValueError):
        lora.apply_lora(m, use_bias=True, num_adapters=2)


def test_rope_table_matches_oracle():
    from csm.models.rope import build_rope_cache
    from oracle.torchtune_shim import Llama3ScaledRoPE
    for hd in (8, 64, 128):
        assert torch.equal(build_rope_cache(hd, 4096, 500000.0, 32.0), Llama3ScaledRoPE(hd, 4096, 500000.0, 32.0).cache)


def test_skinny_gemm_split_choice():
    """Which GEMMs take the split-reduction path (csm/ops.py::_splitk_choice): LoRA-shaped ones only."""
    from csm import ops
    # t = x A^T, dts = dy B, dB = dy^T t, dA = dts^T x  at CSM-1B c2 shapes
    assert ops._splitk_choice(4096, 16, 2048) == 4
    assert ops._splitk_choice(4096, 16, 3072) == 4
    assert ops._splitk_choice(3072, 16, 4096) == 4
    assert ops._splitk_choice(16, 2048, 4096) == 8
    for M, N, K in [(4096, 2048, 2048), (4096, 16384, 2048), (232, 2051, 1024), (4096, 16, 512), (7424, 8, 1024)]:
        assert ops._splitk_choice(M, N, K) == 0, (M, N, K)        # wide, short-K or many-tile problems: one GEMM
    for M, N, K in [(4096, 16, 2048), (16, 2048, 4096), (24, 1000, 7424)]:
        s = ops._splitk_choice(M, N, K)
        assert K % (s * 64) == 0 and K // s >= 256                # every group is whole 64-wide k-blocks


def test_optimizer_selection_and_loud_failure():
    """CPU parameters get stock AdamW (host-logic tests); the kernel optimiser refuses anything but CUDA bf16."""
    from csm.training.optim import FusedClipAdamW
    from csm.training.trainer import clip_and_step, make_optimizer
    p = torch.nn.Parameter(torch.ones(4, 4))
    opt = make_optimizer([{"params": [p], "lr": 1e-2}], 1e-3, 0.01)
    assert isinstance(opt, torch.optim.AdamW) and not isinstance(opt, FusedClipAdamW)
    p.grad = torch.full_like(p, 10.0)
    clip_and_step(opt, [p], 1.0)                                  # clip_grad_norm_ + step (trainer.py:271-276)
    assert float(p.grad.norm()) <= 1.0 + 1e-4 and not torch.equal(p.data, torch.ones(4, 4))
    q = torch.nn.Parameter(torch.ones(4, 4))
    q.grad = torch.ones(4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedClipAdamW([q], lr=1e-3).step()
    with pytest.raises(ValueError):
        FusedClipAdamW([{"params": [q], "betas": (0.9, 0.99)}, {"params": [p], "betas": (0.8, 0.99)}])


def test_sparse_text_exchange_only_for_cuda_data_parallel():
    """The row-sparse text-embedding exchange needs NCCL ranks and CUDA tensors; otherwise the table stays bucketed."""
    from csm.training import dp
    table = torch.nn.Parameter(torch.zeros(10, 4))
    other = torch.nn.Parameter(torch.zeros(3, 3))
    sync = dp.GradSynchronizer([table, other], bucket_bytes=1 << 20, sparse_rows=table)
    assert sync.sparse_param is None and len(sync.params) == 2 and not sync.bucketed


def _streamk_items(unit, n_units, num_tiles, kb_total):
    """Python mirror of csrc/gemm_tc.cu::next_item (stream-K branch) — keep in sync with the kernel."""
    U = num_tiles * kb_total
    u0, u1 = U * unit // n_units, U * (unit + 1) // n_units
    tA, ka = divmod(u0, kb_total)
    tC, kc = divmod(u1, kb_total)
    items = []
    if kc > 0:
        items.append((tC, 0, kc, 1))                        # head part of a tile the next unit finishes
    if ka > 0:
        items.append((tA, ka, kb_total, 2))                 # tail part: this unit finishes the tile
    f0 = tA if ka == 0 else tA + 1
    items += [(t, 0, kb_total, 0) for t in range(f0, tC)]
    return items


@pytest.mark.parametrize("tiles,units,kb", [(128, 74, 128), (128, 74, 33), (116, 74, 16), (75, 74, 9), (300, 70, 48)])
def test_stream_k_schedule_covers_every_k_block_once(tiles, units, kb):
    """Every (tile, k-block) is computed exactly once; a cut tile has exactly one head owner (unit p) and one tail
    owner (unit p + 1), the head is the first item of p and the tail comes before p + 1's whole tiles — the ordering the
    fix-up protocol in the epilogue relies on (no finisher ever waits on work scheduled after its own)."""
    seen = {}
    for p in range(units):
        items = _streamk_items(p, units, tiles, kb)
        kinds = [k for *_, k in items]
        assert kinds == sorted(kinds, key=lambda k: {1: 0, 2: 1, 0: 2}[k])          # head, tail, whole tiles
        assert kinds.count(1) <= 1 and kinds.count(2) <= 1
        for t, a, b, k in items:
            assert 0 <= t < tiles and 0 <= a < b <= kb
            for x in range(a, b):
                assert (t, x) not in seen
                seen[(t, x)] = (p, k)
        load = sum(b - a for _, a, b, _ in items)
        assert abs(load - tiles * kb / units) < 1.0 + 1e-9                           # balanced to one k-block
    assert len(seen) == tiles * kb
    for t in range(tiles):
        owners = sorted({seen[(t, x)] for x in range(kb)})
        if len(owners) == 2:
            (p0, k0), (p1, k1) = owners
            assert p1 == p0 + 1 and k0 == 1 and k1 == 2
        else:
            assert len(owners) == 1 and owners[0][1] == 0


def test_every_op_rejects_cpu_tensors_before_touching_the_library():
    """No CPU fallback and no host pointers handed to a kernel: each compute wrapper raises on CPU tensors."""
    from csm import ops
    bf = torch.bfloat16
    x, w = torch.zeros(4, 8, dtype=bf), torch.zeros(8, 8, dtype=bf)
    tok, msk = torch.zeros(1, 2, 3, dtype=torch.long), torch.ones(1, 2, 3, dtype=torch.bool)
    f32 = torch.zeros(4)
    calls = [
        lambda: ops.embed_gather_sum(tok, msk, w, w),
        lambda: ops.embed_gather_sum_bwd(tok, msk, x, w, w, 4, 8),
        lambda: ops.decoder_input(x.view(1, 4, 8), w, tok, tok[0, :, :2], 3, 2),
        lambda: ops.decoder_input_bwd(x, tok, tok[0, :, :2], x.view(1, 4, 8), None, 3, 2),
        lambda: ops.rmsnorm(x, w[0], 1e-5),
        lambda: ops.rmsnorm_bwd(x, x, w[0], f32, None, None),
        lambda: ops.rope_(x, f32, 4, 1, 8),
        lambda: ops.swiglu(x, x),
        lambda: ops.swiglu_bwd(x, x, x),
        lambda: ops.gemm(x, w),
        lambda: ops.gemm_rope(x, w, f32, 4, 8, 8),
        lambda: ops.gemm_swiglu_fwd(x, w),
        lambda: ops.gemm_swiglu_bwd(x, w, x),
        lambda: ops.attention_fwd(x, x, x, 1, 4, 1, 1, 8),
        lambda: ops.attention_bwd(x, x, x, x, f32, x, 1, 4, 1, 1, 8),
        lambda: ops.linear_ce_fwd(x, w, tok.view(-1)[:4]),
        lambda: ops.linear_ce_bwd(x, w, tok.view(-1)[:4], f32, 1.0, dh=x),
        lambda: ops.f32_to_bf16_(f32, x),
        lambda: ops.add_bf16(x, x),
    ]
    for i, c in enumerate(calls):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            c()
