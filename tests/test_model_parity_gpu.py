"""Parity of the whole training step (Model.forward + backward through libcsm_b200) against the CPU oracle
(oracle/csm_oracle.py, itself pinned bit-exactly to the reference's own code on the semantic path by
tests/golden/make_golden.py).  Gates from BASELINE.json north_star: gather indices/masks bit-exact,
per-codebook bf16 loss within 1e-2 relative, gradient cosine >= 0.999 per trainable tensor."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_tiny.pt")
LOSS_RTOL = 1e-2
GRAD_COS = 0.999


def _product_model(cfg_name):
    from csm.models.model import Model, ModelArgs
    from oracle import csm_oracle as O
    cfg = O.CONFIGS[cfg_name]()
    flav = {"tiny": ("tiny-backbone", "tiny-decoder"), "small": ("small-backbone", "small-decoder")}[cfg_name]
    m = Model(ModelArgs(flav[0], flav[1], cfg.text_vocab_size, cfg.audio_vocab_size, cfg.audio_num_codebooks))
    return m, cfg


def _pair(cfg_name, device, lora, targets=None, r=8):
    """(oracle bf16 on CPU, product bf16 on CUDA) with identical parameters."""
    from csm.models import lora as plora
    from oracle import csm_oracle as O
    prod, cfg = _product_model(cfg_name)
    orc = O.OracleModel(cfg)
    O.init_weights(orc, 0)
    orc = orc.to(torch.bfloat16)
    prod = prod.to(torch.bfloat16)
    if lora:
        O.apply_lora(orc, r=r, alpha=16.0, target_modules=targets, seed=1)
        plora.apply_lora(prod, r=r, alpha=16.0, target_modules=targets, seed=7)
    prod.load_state_dict(orc.state_dict(), strict=True)
    prod = prod.to(device)
    return orc, prod, cfg


def _run_both(orc, prod, cfg, B, S, device, seed=1234):
    from oracle import csm_oracle as O
    batch = O.synthetic_batch(cfg, B, S, seed=seed)
    tok, msk, tgt, fidx = (batch[k] for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"))
    ol, od = O.oracle_forward(orc, tok, msk, tgt, fidx)
    ol.backward()
    pl, pd = prod(tok.to(device), msk.to(device), tgt.to(device), frame_idx=fidx.to(device))
    pl.backward()
    torch.cuda.synchronize()
    return (ol.detach(), od), (pl.detach(), pd)


def _check(orc, prod, o, p, min_cos=GRAD_COS):
    (ol, od), (pl, pd) = o, p
    assert abs(float(pl) - float(ol)) <= LOSS_RTOL * abs(float(ol)), (float(pl), float(ol))
    ref = od["per_codebook_loss"].float()
    got = pd["per_codebook_loss"].float().cpu()
    rel = ((got - ref).abs() / ref.abs()).max().item()
    assert rel <= LOSS_RTOL, f"per-codebook loss rel err {rel}"
    og = {n: q.grad for n, q in orc.named_parameters() if q.grad is not None}
    pg = {n: q.grad for n, q in prod.named_parameters() if q.grad is not None}
    assert set(og) == set(pg), set(og) ^ set(pg)
    worst = (1.0, None)
    for n in og:
        a, b = og[n].float().flatten(), pg[n].float().cpu().flatten()
        if float(a.norm()) == 0.0 and float(b.norm()) == 0.0:
            continue
        c = float(F.cosine_similarity(a, b, dim=0))
        if c < worst[0]:
            worst = (c, n)
    assert worst[0] >= min_cos, f"gradient cosine {worst[0]:.5f} for {worst[1]}"
    return worst


def test_c1_tiny_lora_matches_oracle_and_golden(cuda):
    """BASELINE config 1: tiny model, LoRA r=8 q/v, batch 2."""
    orc, prod, cfg = _pair("tiny", cuda, lora=True)
    o, p = _run_both(orc, prod, cfg, 2, 32, cuda)
    _check(orc, prod, o, p)
    g = torch.load(GOLD)["lora_bf16"]
    # the committed golden was produced by the same oracle code: it guards the oracle against drift ...
    assert torch.allclose(o[1]["per_codebook_loss"], g["per_codebook_loss"], rtol=2e-3, atol=0)
    # ... and pins the CUDA path directly to the committed numbers
    got = p[1]["per_codebook_loss"].cpu()
    assert ((got - g["per_codebook_loss"]).abs() / g["per_codebook_loss"]).max() <= LOSS_RTOL
    named = dict(prod.named_parameters())
    for n, gg in g["grads"].items():
        mine = named[n].grad.float().cpu().flatten()
        assert float(F.cosine_similarity(mine, gg.float().flatten(), dim=0)) >= GRAD_COS, n


def test_c1_tiny_gather_indices_and_masks_bit_exact(cuda):
    from csm import ops
    gold = torch.load(GOLD)
    orc, prod, cfg = _pair("tiny", cuda, lora=False)
    tok, msk = gold["input_tokens"].to(cuda), gold["input_masks"].to(cuda)
    h, idx, eff, status = ops.embed_gather_sum(tok, msk, prod.audio_embeddings.weight, prod.text_embeddings.weight,
                                               debug=True)
    assert torch.equal(idx.cpu(), gold["gather_idx"])                       # tokens + c*V, reference model.py:210-212
    assert torch.equal(eff.bool().cpu(), gold["input_masks"])
    assert int(status.item()) == 0
    # causal mask helper == the reference's indexed tril mask (golden made by the reference's own function)
    prod.backbone.max_seq_len = tok.shape[1]
    prod.setup_caches(2)
    S = tok.shape[1]
    pos = torch.arange(S, device=cuda).unsqueeze(0).repeat(tok.shape[0], 1)
    cm = prod._index_causal_mask(prod.backbone_causal_mask, pos)
    assert torch.equal(cm.cpu(), gold["ref_causal_mask"])
    # values: bit-exact against the oracle's (== reference's) formula evaluated on the same bf16 tables
    oh = (orc._embed_tokens(gold["input_tokens"]) * gold["input_masks"].unsqueeze(-1)).sum(dim=2)
    assert torch.equal(h.cpu(), oh)


def test_tiny_full_finetune_grads(cuda):
    orc, prod, cfg = _pair("tiny", cuda, lora=False)
    o, p = _run_both(orc, prod, cfg, 2, 32, cuda)
    _check(orc, prod, o, p, min_cos=0.99)      # bf16 oracle vs bf16 kernels on 32-wide layers: rounding-noise floor


@pytest.mark.parametrize("targets,r", [(None, 8), (["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj",
                                                  "down_proj"], 16)])
def test_small_lora_tensor_core_path(cuda, targets, r):
    """head_dim 64/128, GQA 4:1, V=2051: runs the tcgen05 GEMM (+LoRA tail), fused CE and attention kernels.
    (None, 8) is config 2's adapter set; the second case is config 4's (r=16 on all seven projections)."""
    orc, prod, cfg = _pair("small", cuda, lora=True, targets=targets, r=r)
    o, p = _run_both(orc, prod, cfg, 2, 128, cuda)
    _check(orc, prod, o, p)


@pytest.mark.parametrize("lora", [True, False])
def test_small_model_on_cta_pair_and_fused_swiglu_kernels(cuda, lora):
    """The kernels CSM-1B shapes select automatically — tcgen05 cta_group::2 GEMMs (256-row tiles over a 2-CTA cluster)
    and the w1|w3 GEMM with the SwiGLU epilogue — forced on for the small model so the whole training step is checked
    against the oracle on them (LoRA on all seven projections: extra-K-block tail through both; and full fine-tune)."""
    from csm import autograd, ops
    targets = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"] if lora else None
    orc, prod, cfg = _pair("small", cuda, lora=lora, targets=targets, r=16)
    ops.set_gemm_cta_pair_mode(1)
    try:
        assert ops.swiglu_fusable(2 * 128, 512, 256)
        for fuse_bwd in (False, True):
            autograd.FUSE_SWIGLU_BWD = fuse_bwd
            o, p = _run_both(orc, prod, cfg, 2, 128, cuda)
            _check(orc, prod, o, p, min_cos=0.999 if lora else 0.99)
            for q in list(orc.parameters()) + list(prod.parameters()):
                q.grad = None
    finally:
        autograd.FUSE_SWIGLU_BWD = False
        ops.set_gemm_cta_pair_mode(-1)


def test_small_full_finetune_tensor_core_path(cuda):
    orc, prod, cfg = _pair("small", cuda, lora=False)
    o, p = _run_both(orc, prod, cfg, 2, 128, cuda)
    _check(orc, prod, o, p, min_cos=0.99)


def test_compute_loss_contract(cuda):
    """The reference's own assertion on this path (test_training.py:209-236): scalar, positive, has semantic_loss."""
    from csm.training.utils import compute_loss
    orc, prod, cfg = _pair("tiny", cuda, lora=False)
    B, S = 2, 5
    tok = torch.randint(0, 100, (B, S, 33), device=cuda)
    msk = torch.ones(B, S, 33, dtype=torch.bool, device=cuda)
    tgt = torch.randint(0, 100, (B, S, 32), device=cuda)
    loss, comp = compute_loss(prod, tok, msk, tgt)
    assert isinstance(loss, torch.Tensor) and loss.dim() == 0 and loss.item() > 0
    assert "semantic_loss" in comp and isinstance(comp["semantic_loss"], torch.Tensor)
    assert "acoustic_loss" in comp


def test_reference_test_shape_b2_s5_matches_oracle(cuda):
    """The reference's own fixture shape for this path (test_training.py:215-219: B=2, S=5, tokens < 100, all-ones
    mask) — not just "a positive scalar": loss, semantic / acoustic terms, per-codebook losses and every gradient
    against the oracle (whose semantic path is bit-equal to the reference's compute_loss, tests/golden/make_golden.py)."""
    from csm.training.utils import compute_loss
    from oracle import csm_oracle as O
    orc, prod, cfg = _pair("tiny", cuda, lora=False)
    B, S = 2, 5
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, 100, (B, S, 33), generator=g)
    msk = torch.ones(B, S, 33, dtype=torch.bool)
    tgt = torch.randint(0, 100, (B, S, 32), generator=g)
    fidx = torch.tensor([[0, 0], [0, 3], [1, 1], [1, 2], [1, 3]])              # p < S-1
    ol, od = O.oracle_forward(orc, tok, msk, tgt, fidx)
    ol.backward()
    pl, pd = compute_loss(prod, tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda))
    pl.backward()
    torch.cuda.synchronize()
    assert abs(float(pd["semantic_loss"]) - float(od["semantic_loss"])) <= LOSS_RTOL * float(od["semantic_loss"])
    assert abs(float(pd["acoustic_loss"]) - float(od["acoustic_loss"])) <= LOSS_RTOL * float(od["acoustic_loss"])
    _check(orc, prod, (ol.detach(), od), (pl.detach(), pd), min_cos=0.99)
    # without an explicit frame_idx the model picks ceil(4/16) = 1 target frame per sample itself; the semantic term
    # does not depend on that choice
    pl2, pd2 = compute_loss(prod, tok.to(cuda), msk.to(cuda), tgt.to(cuda))
    assert float(pd2["semantic_loss"]) == float(pd["semantic_loss"]) and float(pl2) > 0


@pytest.mark.parametrize("share_backbone", [True, False])
def test_multi_adapter_batching_matches_per_row_adapter_oracle(cuda, share_backbone):
    """three speakers' adapters side by side in every adapted projection, one mixed-speaker batch,
    ONE base GEMM — against the plain-torch restatement that selects each row's adapter explicitly
    (oracle.MultiLoRALinear): losses and every adapter gradient, including the zero gradient of a speaker that is
    absent from the batch.  r=16 on all seven projections: the fused q|k|v tail is 3 x 3 x 16 = 144 columns wide."""
    from csm.models import lora as plora
    from oracle import csm_oracle as O
    prod, cfg = _product_model("small")
    orc = O.OracleModel(cfg)
    O.init_weights(orc, 0)
    orc, prod = orc.to(torch.bfloat16), prod.to(torch.bfloat16)
    targets = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
    adapters = {"backbone": 1 if share_backbone else 3, "decoder": 3}
    O.apply_multi_lora(orc, 16, 16.0, adapters, target_modules=targets, seed=1)
    plora.apply_lora(prod, r=16, alpha=16.0, target_modules=targets, seed=9, num_adapters=adapters)
    prod.load_state_dict(orc.state_dict(), strict=True)
    prod = prod.to(cuda)
    B, S = 4, 128
    batch = O.synthetic_batch(cfg, B, S, seed=31)
    speakers = torch.tensor([2, 0, 2, 0])                                      # adapter 1 never appears
    tok, msk, tgt, fidx = (batch[k] for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"))
    O.set_adapter_rows(orc, speakers, fidx, S, cfg.audio_num_codebooks)
    ol, od = O.oracle_forward(orc, tok, msk, tgt, fidx)
    ol.backward()
    pl, pd = prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda), speaker_ids=speakers.to(cuda))
    pl.backward()
    torch.cuda.synchronize()
    assert abs(float(pl) - float(ol)) <= LOSS_RTOL * abs(float(ol))
    rel = ((pd["per_codebook_loss"].cpu() - od["per_codebook_loss"]).abs() / od["per_codebook_loss"]).max().item()
    assert rel <= LOSS_RTOL
    named = dict(prod.named_parameters())
    checked = 0
    for n, q in orc.named_parameters():
        if q.grad is None:
            continue
        a, b = q.grad.float(), named[n].grad.float().cpu()
        K = adapters["backbone" if n.startswith("backbone") else "decoder"]
        for k in range(K):
            sa = a[k * 16:(k + 1) * 16] if n.endswith("lora_A") else a[:, k * 16:(k + 1) * 16]
            sb = b[k * 16:(k + 1) * 16] if n.endswith("lora_A") else b[:, k * 16:(k + 1) * 16]
            if K > 1 and k == 1:
                assert float(sa.abs().max()) == 0.0 and float(sb.abs().max()) == 0.0, n     # absent speaker: no gradient
                continue
            c = float(F.cosine_similarity(sa.flatten(), sb.flatten(), dim=0))
            assert c >= GRAD_COS, (n, k, c)
            checked += 1
    assert checked > 50
    with pytest.raises(RuntimeError):
        prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda))     # several adapters: ids required


def _gen_prompt(cfg, B, S, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(B, S, 33, dtype=torch.int64)
    msk = torch.zeros(B, S, 33, dtype=torch.bool)
    st = S // 2
    tok[:, :st, 32] = torch.randint(0, cfg.text_vocab_size, (B, st), generator=g)
    msk[:, :st, 32] = True
    tok[:, st:, :32] = torch.randint(0, cfg.audio_vocab_size, (B, S - st, 32), generator=g)
    msk[:, st:, :32] = True
    return tok, msk


def _frame_as_input(codes):
    B = codes.shape[0]
    t = torch.cat([codes.long(), torch.zeros(B, 1, dtype=torch.int64, device=codes.device)], dim=1).unsqueeze(1)
    m = torch.cat([torch.ones(B, 32, dtype=torch.bool, device=codes.device),
                   torch.zeros(B, 1, dtype=torch.bool, device=codes.device)], dim=1).unsqueeze(1)
    return t, m


@pytest.mark.parametrize("cfg_name,S,steps", [("tiny", 12, 4), ("small", 160, 3)])
def test_generate_frame_with_kv_cache_matches_oracle(cuda, cfg_name, S, steps):
    """Model.generate_frame (model.py:140-195) on the training kernels + KV caches — prompt
    prefill, then single-frame steps through the decode-attention kernel, 31 cached depth-decoder steps per frame —
    against the oracle's restatement, which equals the REFERENCE's own generate_frame code for code
    (tests/golden/make_golden_generate.py, tests/test_oracle.py).  topk = 1: sample_topk is the argmax.  The oracle runs
    in fp32 on the same bf16-representable weights; its codes are fed back on both sides (teacher forcing) so that every
    one of the 32 logit rows of every frame is compared, and the kernels' own argmax must agree wherever the oracle's
    top-2 margin is decisive.  The small config prefills 160 positions through the tcgen05 attention kernel."""
    from oracle import csm_oracle as O
    prod, cfg = _product_model(cfg_name)
    orc = O.OracleModel(cfg)
    O.init_weights(orc, 0, std=0.3 if cfg_name == "tiny" else 0.08)
    with torch.no_grad():
        for q in orc.parameters():
            q.copy_(q.to(torch.bfloat16).float())
    prod = prod.to(torch.bfloat16)
    prod.load_state_dict(orc.state_dict())
    prod = prod.to(cuda)
    B = 2
    tok, msk = _gen_prompt(cfg, B, S, 5)
    orc.setup_caches(B)
    prod.setup_caches(B)
    assert prod.backbone.caches_are_enabled() and prod.decoder.caches_are_enabled()
    agree = decisive = total = 0
    t_o, m_o, pos = tok, msk, torch.arange(S).unsqueeze(0).repeat(B, 1)
    for step in range(steps + 1):
        o_codes, o_logits = orc.generate_frame(t_o, m_o, pos, 0.9, 1, return_logits=True)
        p_codes, p_logits = prod.generate_frame(t_o.to(cuda), m_o.to(cuda), pos.to(cuda), 0.9, 1, return_logits=True,
                                                forced_codes=o_codes.to(cuda))
        assert p_codes.shape == (B, 32) and p_codes.dtype == torch.int32
        assert torch.equal(p_codes.cpu(), o_codes)                       # forced: the fed-back codes are the oracle's
        for i, (lo, lp) in enumerate(zip(o_logits, p_logits)):
            lp = lp.float().cpu()
            c = float(F.cosine_similarity(lo.flatten(), lp.flatten(), dim=0))
            assert c >= 0.999, (step, i, c)
            assert float((lo - lp).abs().max()) <= 0.25 * float(lo.std()) + 0.05, (step, i)
            top2 = lo.topk(2, dim=-1).values
            margin = top2[:, 0] - top2[:, 1]
            mine = lp.argmax(-1)
            for b in range(B):
                total += 1
                ok = int(mine[b]) == int(o_codes[b, i])
                agree += ok
                if float(margin[b]) > 0.25 * float(lo.std()):
                    decisive += 1
                    assert ok, (step, i, b, float(margin[b]))
        t_o, m_o = _frame_as_input(o_codes)
        pos = torch.full((B, 1), S + step)
    assert decisive > total // 4 and agree >= 0.9 * total, (agree, decisive, total)
    # free-running (no teacher forcing): the API call a user makes; same codes as the forced run on the first frame
    prod.reset_caches()
    free = prod.generate_frame(tok.to(cuda), msk.to(cuda), torch.arange(S, device=cuda).unsqueeze(0).repeat(B, 1), 0.9, 1)
    assert free.shape == (B, 32)
    if cfg_name == "tiny":
        # the committed golden: codes of the REFERENCE's own generate_frame (fp32 weights there, their bf16 rounding
        # here).  Teacher-forced with the reference's codes (a flipped near-tie would otherwise cascade through the
        # rest of the frame), the kernels' argmax reproduces nearly every one of the 2 x 5 x 32 draws.
        gold = torch.load(os.path.join(os.path.dirname(GOLD), "c1_tiny_generate.pt"))
        assert torch.equal(gold["tokens"], tok) and torch.equal(gold["mask"], msk)
        prod.reset_caches()
        hits = n = 0
        t_g, m_g, pos_g = tok, msk, torch.arange(S).unsqueeze(0).repeat(B, 1)
        for f in range(gold["frames"].shape[1]):
            want = gold["frames"][:, f]
            _, lg = prod.generate_frame(t_g.to(cuda), m_g.to(cuda), pos_g.to(cuda), gold["temperature"], gold["topk"],
                                        return_logits=True, forced_codes=want.to(cuda))
            mine = torch.stack([x.float().argmax(-1) for x in lg], dim=1).cpu()
            hits += int((mine == want).sum())
            n += want.numel()
            t_g, m_g = _frame_as_input(want)
            pos_g = torch.full((B, 1), S + f)
        assert hits >= 0.9 * n, (hits, n)
    with pytest.raises(RuntimeError):
        prod.generate_frame(tok.to(cuda), msk.to(cuda), torch.arange(S, device=cuda).unsqueeze(0).repeat(B, 1), 0.9, 1)


def test_sequence_packing_equals_running_every_sample_alone(cuda):
    """five variable-length samples packed into rows of 384 frames (csm/data/frames.py::
    pack_samples; block-diagonal causal attention, per-sample RoPE positions) give the loss and the gradients of the
    same five samples run ONE BY ONE without any padding — the semantics of the reference's compute_loss on each
    sample (position p predicts target row p, last position excluded), which its zero-padding collate only blurs."""
    from csm.data.frames import pack_samples, packing_efficiency
    from csm.data.synthetic import synthetic_batch
    from oracle import csm_oracle as O
    orc, prod, cfg = _pair("small", cuda, lora=False)
    lens = [130, 200, 384, 150, 97]
    samples = []
    for i, n in enumerate(lens):
        b = synthetic_batch(cfg.text_vocab_size, cfg.audio_vocab_size, cfg.audio_num_codebooks, 1, n, seed=70 + i)
        samples.append({"input_tokens": b["input_tokens"][0], "input_masks": b["input_masks"][0],
                        "target_audio_tokens": b["target_audio_tokens"][0]})
    packed = pack_samples(samples, 384, pad_to_multiple=128, generator=torch.Generator().manual_seed(1), pin=False)
    R, S = packed["input_tokens"].shape[:2]
    assert (R, S) == (3, 384) and packing_efficiency(packed) > 0.8
    dev = {k: v.to(cuda) for k, v in packed.items()}
    loss, det = prod(dev["input_tokens"], dev["input_masks"], dev["target_audio_tokens"], frame_idx=dev["frame_idx"],
                     segment_starts=dev["segment_starts"], segment_ends=dev["segment_ends"],
                     target_mask=dev["target_mask"])
    loss.backward()
    torch.cuda.synchronize()
    g_packed = {n: p.grad.detach().float().clone() for n, p in prod.named_parameters() if p.grad is not None}
    for p in prod.parameters():
        p.grad = None
    # the same samples alone (B = 1, their own length, no padding), weighted as the packed means weight them
    n_sem = sum(n - 1 for n in lens)
    owner, fi = packed["sample_index"], packed["frame_idx"]
    n_ac = fi.shape[0]
    sem_sum = ac_sum = 0.0
    o_sem = 0.0
    for j, smp in enumerate(samples):
        rows = (owner[fi[:, 0], fi[:, 1]] == j)
        r = int(torch.nonzero((owner == j).any(1))[0])
        off = int(packed["segment_starts"][r][owner[r] == j][0])
        mine = fi[rows]
        fj = torch.stack([torch.zeros_like(mine[:, 1]), mine[:, 1] - off], dim=1)
        tok, msk, tgt = (smp[k].unsqueeze(0) for k in ("input_tokens", "input_masks", "target_audio_tokens"))
        lj, dj = prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fj.to(cuda),
                      semantic_weight=100.0 * (lens[j] - 1) / n_sem, acoustic_weight=1.0 * fj.shape[0] / n_ac)
        lj.backward()                                   # gradients accumulate over the samples
        sem_sum += float(dj["semantic_loss"]) * (lens[j] - 1)
        ac_sum += float(dj["acoustic_loss"]) * fj.shape[0]
        ol, od = O.oracle_forward(orc, tok, msk, tgt, fj)
        o_sem += float(od["semantic_loss"]) * (lens[j] - 1)
    torch.cuda.synchronize()
    assert abs(float(det["semantic_loss"]) - sem_sum / n_sem) <= 2e-3 * sem_sum / n_sem
    assert abs(float(det["acoustic_loss"]) - ac_sum / n_ac) <= 2e-3 * ac_sum / n_ac
    assert abs(float(det["semantic_loss"]) - o_sem / n_sem) <= LOSS_RTOL * o_sem / n_sem          # and the oracle's
    for n, p in prod.named_parameters():
        if p.grad is None:
            continue
        c = float(F.cosine_similarity(g_packed[n].flatten(), p.grad.float().flatten(), dim=0))
        assert c >= 0.999, (n, c)
        assert abs(float(g_packed[n].norm()) - float(p.grad.float().norm())) <= 3e-2 * float(p.grad.float().norm()), n
    # packed rows need their segment tables and their own frame_idx
    with pytest.raises(RuntimeError):
        prod(dev["input_tokens"], dev["input_masks"], dev["target_audio_tokens"], segment_starts=dev["segment_starts"],
             segment_ends=dev["segment_ends"])


@pytest.mark.parametrize("p_drop,use_bias", [(0.0, True), (0.25, False), (0.1, True)])
def test_lora_dropout_and_bias_match_oracle(cuda, p_drop, use_bias):
    """The two remaining options of the reference's LoRALinear (lora.py:87-90 dropout on the low-rank path's input,
    :66,101-102 a LoRA bias), r=8 on all seven projections of the small model.  Dropout masks are stateless hashes:
    the test regenerates the very masks the kernels will use (same seed / salt) and hands them to the oracle, so losses
    and every gradient — lora_A, lora_B, lora_bias — are compared exactly like in the other parity tests; in eval mode
    dropout is off."""
    from csm import ops
    from csm.models import lora as plora
    from oracle import csm_oracle as O
    prod, cfg = _product_model("small")
    orc = O.OracleModel(cfg)
    O.init_weights(orc, 0)
    orc, prod = orc.to(torch.bfloat16), prod.to(torch.bfloat16)
    targets = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
    O.apply_lora(orc, r=8, alpha=16.0, target_modules=targets, seed=1, use_bias=use_bias)
    plora.apply_lora(prod, r=8, alpha=16.0, target_modules=targets, seed=4, dropout=p_drop, use_bias=use_bias)
    prod.load_state_dict(orc.state_dict(), strict=True)
    prod = prod.to(cuda).train()
    B, S = 2, 128
    batch = O.synthetic_batch(cfg, B, S, seed=77)
    tok, msk, tgt, fidx = (batch[k] for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"))
    if p_drop > 0:
        seed = torch.ones(1, dtype=torch.int64, device=cuda)                  # the value of the first training forward
        C = cfg.audio_num_codebooks
        for stack, rows, shape3 in ((orc.backbone, B * S, (B, S)), (orc.decoder, fidx.shape[0] * C, (fidx.shape[0], C))):
            for li, layer in enumerate(stack.layers):
                a, f = layer.attn, layer.mlp
                for salt, mods in ((4 * li, (a.q_proj, a.k_proj, a.v_proj)), (4 * li + 1, (f.w1, f.w3)),
                                   (4 * li + 2, (a.output_proj,)), (4 * li + 3, (f.w2,))):
                    ones = torch.ones(rows, mods[0].weight.shape[1], dtype=torch.bfloat16, device=cuda)
                    keep = ops.lora_dropout(ones, p_drop, seed, salt).float().cpu()
                    frac = float((keep > 0).float().mean())
                    assert abs(frac - (1 - p_drop)) < 0.02 and abs(float(keep.max()) - 1 / (1 - p_drop)) < 1e-2
                    for m in mods:
                        m.keep = keep.view(*shape3, -1)
    ol, od = O.oracle_forward(orc, tok, msk, tgt, fidx)
    ol.backward()
    pl, pd = prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda))
    pl.backward()
    torch.cuda.synchronize()
    _check(orc, prod, (ol.detach(), od), (pl.detach(), pd))
    if use_bias:
        assert prod.backbone.layers[0].attn.q_proj.lora_bias.grad is not None
    if p_drop > 0:
        prod.eval()
        for m in orc.modules():
            if isinstance(m, O.LoRALinear):
                m.keep = None
        with torch.no_grad():
            pe, _ = prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda))
            oe, _ = O.oracle_forward(orc, tok, msk, tgt, fidx)
        assert abs(float(pe) - float(oe)) <= LOSS_RTOL * abs(float(oe)) and float(pe) != float(pl)


@pytest.mark.parametrize("lora", [True, False])
def test_compact_token_format_is_bit_identical(cuda, lora):
    """the int32-rows + mask-word batch gives the same loss and gradients as the reference's int64 /
    bool batch (same kernels downstream; the embedding tables' scatter-add uses bf16 atomics and the norm scales' gradient
    fp32 atomics, so those are compared with a tolerance)."""
    from csm.data.frames import compact_batch
    from oracle import csm_oracle as O
    _, prod, cfg = _pair("small", cuda, lora=lora)
    batch = O.synthetic_batch(cfg, 2, 128, seed=99)
    cb = compact_batch(batch, cfg.audio_vocab_size, pin=False)
    outs = []
    for b in (batch, cb):
        prod.zero_grad(set_to_none=True)
        loss, d = prod(b["input_tokens"].to(cuda), b["input_masks"].to(cuda), b["target_audio_tokens"].to(cuda),
                       frame_idx=b["frame_idx"].to(cuda))
        loss.backward()
        torch.cuda.synchronize()
        outs.append((loss.detach().clone(), d["per_codebook_loss"].clone(),
                     {n: p.grad.clone() for n, p in prod.named_parameters() if p.grad is not None}))
    (l0, c0, g0), (l1, c1, g1) = outs
    assert torch.equal(l0, l1) and torch.equal(c0, c1)
    assert set(g0) == set(g1)
    for n in g0:
        if "embeddings" in n or n.endswith(".scale"):
            assert float(F.cosine_similarity(g0[n].float().flatten(), g1[n].float().flatten(), dim=0)) > 0.9999, n
        else:
            assert torch.equal(g0[n], g1[n]), n


def test_no_cpu_fallback():
    from csm.models.model import Model, ModelArgs
    m = Model(ModelArgs("tiny-backbone", "tiny-decoder", 1000, 200, 32)).to(torch.bfloat16)
    tok = torch.zeros(1, 4, 33, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        m(tok, torch.ones(1, 4, 33, dtype=torch.bool), torch.zeros(1, 4, 32, dtype=torch.int64))
