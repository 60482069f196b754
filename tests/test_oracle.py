"""CPU tests of the oracle (test infrastructure): pinned against the committed golden vectors, against the
reference's own code when /root/reference is present (build container), and against independent restatements."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import csm_oracle as O
from oracle import reference_loader as R
from oracle import torchtune_shim as tt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_tiny.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.fixture(scope="module")
def tiny():
    torch.set_num_threads(1)
    cfg = O.cfg_tiny()
    m = O.OracleModel(cfg)
    O.init_weights(m, 0)
    return cfg, m


def test_synthetic_batch_is_reproducible_and_matches_golden(gold, tiny):
    cfg, _ = tiny
    b = O.synthetic_batch(cfg, gold["B"], gold["S"], seed=gold["batch_seed"])
    for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"):
        assert torch.equal(b[k], gold[k]), k
    assert torch.equal(O.gather_indices(b["input_tokens"], cfg.audio_vocab_size, cfg.audio_num_codebooks),
                       gold["gather_idx"])


def test_oracle_semantic_path_equals_reference_golden(gold, tiny):
    """ref_* entries were produced by the reference's own model.py/compute_loss (tests/golden/make_golden.py)."""
    cfg, m = tiny
    m.zero_grad()
    loss, d = O.oracle_forward(m, gold["input_tokens"], gold["input_masks"], gold["target_audio_tokens"], None)
    loss.backward()
    assert torch.equal(loss.detach(), gold["ref_loss"])
    assert torch.equal(d["semantic_loss"].detach(), gold["ref_semantic_loss"])
    assert torch.equal(m.codebook0_head.weight.grad, gold["ref_grad_codebook0_head"])
    assert torch.equal(m.backbone.layers[0].attn.q_proj.weight.grad, gold["ref_grad_q_proj_l0"])
    h = (m._embed_tokens(gold["input_tokens"]) * gold["input_masks"].unsqueeze(-1)).sum(dim=2)
    assert torch.equal(h, gold["ref_h_embed"])


def test_oracle_full_step_matches_golden(gold, tiny):
    cfg, m = tiny
    m.zero_grad()
    loss, d = O.oracle_forward(m, gold["input_tokens"], gold["input_masks"], gold["target_audio_tokens"],
                               gold["frame_idx"])
    loss.backward()
    g = gold["fullft_fp32"]
    assert torch.allclose(loss.detach(), g["loss"], rtol=1e-6)
    assert torch.allclose(d["per_codebook_loss"], g["per_codebook_loss"], rtol=1e-6)
    named = dict(m.named_parameters())
    for n, ref in g["grads"].items():
        assert torch.allclose(named[n].grad, ref, rtol=1e-4, atol=1e-7), n
    # the decoder term trains what the reference leaves untrained
    assert m.audio_head.grad is not None and float(m.audio_head.grad.abs().sum()) > 0


@pytest.mark.parametrize("tag,dtype", [("lora_fp32", torch.float32), ("lora_bf16", torch.bfloat16)])
def test_oracle_lora_matches_golden(gold, tiny, tag, dtype):
    cfg, base = tiny
    m = O.OracleModel(cfg)
    m.load_state_dict(base.state_dict())
    m = m.to(dtype)
    names = O.apply_lora(m, r=8, alpha=16.0, seed=1)
    assert len(names) == 2 * (cfg.backbone.num_layers + cfg.decoder.num_layers)
    loss, d = O.oracle_forward(m, gold["input_tokens"], gold["input_masks"], gold["target_audio_tokens"],
                               gold["frame_idx"])
    loss.backward()
    g = gold[tag]
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert torch.allclose(d["per_codebook_loss"], g["per_codebook_loss"], rtol=tol)
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    for n, ref in g["grads"].items():
        c = float(F.cosine_similarity(got[n].float().flatten(), ref.float().flatten(), dim=0))
        assert c > 0.9999, (n, c)


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_reference_code_live():
    torch.set_num_threads(1)
    rm, ru = R.load_reference()
    cfg = O.cfg_tiny()
    R.register_flavor(rm, "tiny-bb", cfg.backbone)
    R.register_flavor(rm, "tiny-dec", cfg.decoder)
    ref = rm.Model(rm.ModelArgs("tiny-bb", "tiny-dec", cfg.text_vocab_size, cfg.audio_vocab_size,
                                cfg.audio_num_codebooks))
    om = O.OracleModel(cfg)
    O.init_weights(om, 3)
    assert set(ref.state_dict()) == set(om.state_dict())
    ref.load_state_dict(om.state_dict())
    B, S = 2, 5                                  # the reference's own test shape (test_training.py:215-219)
    tok = torch.randint(0, 100, (B, S, 33))
    msk = torch.ones(B, S, 33, dtype=torch.bool)
    tgt = torch.randint(0, 100, (B, S, 32))
    ref.backbone_causal_mask = rm._create_causal_mask(S, torch.device("cpu"))
    ref._index_causal_mask = lambda m, p: rm._index_causal_mask(m, p)
    rl, rd = ru.compute_loss(ref, tok, msk, tgt)
    ol, od = O.oracle_forward(om, tok, msk, tgt, None)
    assert torch.equal(rl, ol) and torch.equal(rd["semantic_loss"], od["semantic_loss"])
    assert float(rd["acoustic_loss"]) == 0.0     # the reference's placeholder (utils.py:116)
    assert torch.equal(ref._embed_tokens(tok), om._embed_tokens(tok))
    assert torch.equal(ref._embed_audio(3, tok[:, :, 3]), om._embed_audio(3, tok[:, :, 3]))


def test_rope_scaling_matches_transformers_llama3():
    """Independent corroboration of the [recalled] torchtune spec: HF's Llama-3 rope init."""
    tr = pytest.importorskip("transformers.modeling_rope_utils")
    from transformers import LlamaConfig
    hd = 64
    cfg = LlamaConfig(hidden_size=hd * 4, num_attention_heads=4, max_position_embeddings=2048, rope_theta=500000.0,
                      rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                                    "high_freq_factor": 4.0, "original_max_position_embeddings": 8192})
    try:
        inv, _ = tr._compute_llama3_parameters(cfg, "cpu")
    except Exception as e:                       # config schema differs across transformers versions
        pytest.skip(f"transformers rope API changed: {e}")
    ours = tt.llama3_scaled_freqs(hd, 500000.0, 32.0)
    assert torch.allclose(ours, inv.float(), rtol=1e-6)


def test_llama_stack_restatement_matches_transformers_llama():
    """The whole restated torchtune stack (RMSNorm, scaled RoPE, GQA attention, SwiGLU, residual wiring, final norm)
    against transformers' independent Llama implementation on the same weights, forward and input gradient.  HF
    rotates (i, i + hd/2) pairs where torchtune rotates (2i, 2i + 1): the q/k projection rows are permuted per head,
    which leaves q.k unchanged."""
    pytest.importorskip("transformers")
    from transformers import LlamaConfig, LlamaModel
    torch.manual_seed(0)
    D, H, KV, L, I, S = 128, 4, 2, 3, 256, 40
    hd = D // H
    dec = tt.llama3_2(vocab_size=32, num_layers=L, num_heads=H, num_kv_heads=KV, embed_dim=D, max_seq_len=2048,
                      intermediate_dim=I, norm_eps=1e-5, scale_factor=32)
    dec.tok_embeddings = torch.nn.Identity()                 # as the reference does (model.py:51-56)
    for p_ in dec.parameters():
        torch.nn.init.normal_(p_, std=0.08)
    for m in dec.modules():
        if isinstance(m, tt.RMSNorm):
            torch.nn.init.normal_(m.scale, mean=1.0, std=0.1)
    try:
        cfg = LlamaConfig(vocab_size=32, hidden_size=D, intermediate_size=I, num_hidden_layers=L,
                          num_attention_heads=H, num_key_value_heads=KV, head_dim=hd, max_position_embeddings=131072,
                          rms_norm_eps=1e-5, rope_theta=500000.0, attention_bias=False, mlp_bias=False,
                          rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                                        "high_freq_factor": 4.0, "original_max_position_embeddings": 8192},
                          attn_implementation="eager")
        hf = LlamaModel(cfg).eval()
    except Exception as e:                                   # config schema differs across transformers versions
        pytest.skip(f"transformers Llama config API changed: {e}")

    def halves(w, nh):                                       # interleaved pair layout -> HF's two-halves layout
        w = w.view(nh, hd // 2, 2, -1)
        return torch.cat([w[:, :, 0], w[:, :, 1]], 1).reshape(nh * hd, -1)

    with torch.no_grad():
        for a, b in zip(dec.layers, hf.layers):
            b.self_attn.q_proj.weight.copy_(halves(a.attn.q_proj.weight, H))
            b.self_attn.k_proj.weight.copy_(halves(a.attn.k_proj.weight, KV))
            b.self_attn.v_proj.weight.copy_(a.attn.v_proj.weight)
            b.self_attn.o_proj.weight.copy_(a.attn.output_proj.weight)
            b.mlp.gate_proj.weight.copy_(a.mlp.w1.weight)
            b.mlp.down_proj.weight.copy_(a.mlp.w2.weight)
            b.mlp.up_proj.weight.copy_(a.mlp.w3.weight)
            b.input_layernorm.weight.copy_(a.sa_norm.scale)
            b.post_attention_layernorm.weight.copy_(a.mlp_norm.scale)
        hf.norm.weight.copy_(dec.norm.scale)
    x1 = torch.randn(2, S, D, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    pos = torch.arange(S)[None].expand(2, -1)
    mask = torch.tril(torch.ones(S, S, dtype=torch.bool))[None].expand(2, -1, -1)
    o1 = dec(x1, mask=mask, input_pos=pos)
    o2 = hf(inputs_embeds=x2, position_ids=pos).last_hidden_state
    assert torch.allclose(o1, o2, atol=2e-5, rtol=1e-5), float((o1 - o2).abs().max())
    probe = torch.randn_like(o1)
    (o1 * probe).sum().backward()
    (o2 * probe).sum().backward()
    assert torch.allclose(x1.grad, x2.grad, atol=2e-5, rtol=1e-4), float((x1.grad - x2.grad).abs().max())


def test_decoder_term_matches_transformers_csm_depth_decoder():
    """The acoustic term the reference leaves as a placeholder (utils.py:109-117), restated in the oracle from
    generate_frame (model.py:171-193), against the training loss of transformers' CSM port on the same weights:
    ``CsmDepthDecoderForCausalLM`` takes [placeholder, c_0 .. c_{C-2}] with the backbone state written over position 0,
    embeds position i with codebook i-1's table, projects, runs the depth decoder and scores position i with head i-1
    against c_i (modeling_csm.py:441-488, 522-539, 565-623).  Same alignment, same mean over frames x (C-1) codes."""
    pytest.importorskip("transformers")
    try:
        from transformers.models.csm.configuration_csm import CsmDepthDecoderConfig
        from transformers.models.csm.modeling_csm import CsmDepthDecoderForCausalLM
    except Exception as e:
        pytest.skip(f"transformers has no CSM port: {e}")
    torch.manual_seed(0)
    C, V, D, Dd, H, KV, L, I = 8, 50, 96, 64, 4, 2, 2, 128
    hd = Dd // H
    cfg = O.OracleCfg(backbone=O.StackCfg(1, 2, 2, D, 64, 2048), decoder=O.StackCfg(L, H, KV, Dd, I, 2048),
                      text_vocab_size=30, audio_vocab_size=V, audio_num_codebooks=C)
    om = O.OracleModel(cfg)
    O.init_weights(om, 3, std=0.08)
    for m in om.decoder.modules():
        if isinstance(m, tt.RMSNorm):
            torch.nn.init.normal_(m.scale, mean=1.0, std=0.1)
    try:
        hc = CsmDepthDecoderConfig(
            num_codebooks=C, backbone_hidden_size=D, vocab_size=V, hidden_size=Dd, intermediate_size=I,
            num_hidden_layers=L, num_attention_heads=H, num_key_value_heads=KV, max_position_embeddings=33,
            rms_norm_eps=1e-5,
            rope_parameters={"rope_type": "llama3", "rope_theta": 500000.0, "factor": 32.0, "low_freq_factor": 1.0,
                             "high_freq_factor": 4.0, "original_max_position_embeddings": 8192})
        hc._attn_implementation = "eager"
        hf = CsmDepthDecoderForCausalLM(hc).eval()
    except Exception as e:                                   # config schema differs across transformers versions
        pytest.skip(f"transformers CSM config API changed: {e}")

    def halves(w, nh):                                       # interleaved pair layout -> HF's two-halves layout
        w = w.view(nh, hd // 2, 2, -1)
        return torch.cat([w[:, :, 0], w[:, :, 1]], 1).reshape(nh * hd, -1)

    with torch.no_grad():
        hf.model.embed_tokens.weight.copy_(om.audio_embeddings.weight)
        hf.model.inputs_embeds_projector.weight.copy_(om.projection.weight)
        hf.codebooks_head.weight.copy_(om.audio_head)
        for a, b in zip(om.decoder.layers, hf.model.layers):
            b.self_attn.q_proj.weight.copy_(halves(a.attn.q_proj.weight, H))
            b.self_attn.k_proj.weight.copy_(halves(a.attn.k_proj.weight, KV))
            b.self_attn.v_proj.weight.copy_(a.attn.v_proj.weight)
            b.self_attn.o_proj.weight.copy_(a.attn.output_proj.weight)
            b.mlp.gate_proj.weight.copy_(a.mlp.w1.weight)
            b.mlp.down_proj.weight.copy_(a.mlp.w2.weight)
            b.mlp.up_proj.weight.copy_(a.mlp.w3.weight)
            b.input_layernorm.weight.copy_(a.sa_norm.scale)
            b.post_attention_layernorm.weight.copy_(a.mlp_norm.scale)
        hf.model.norm.weight.copy_(om.decoder.norm.scale)
    Ns = 5
    hrows = torch.randn(Ns, D)
    codes = torch.randint(0, V, (Ns, C))
    with torch.no_grad():
        ce = O.oracle_decoder_ce(om, hrows, codes)                                   # [Ns, C-1]
        ids = torch.nn.functional.pad(codes[:, : C - 1], (1, 0), value=0)
        out = hf(input_ids=ids, backbone_last_hidden_state=hrows.clone(), labels=codes, use_cache=False)
    assert out.logits.shape == (Ns, C - 1, V)
    assert abs(float(ce.mean()) - float(out.loss)) < 1e-5 * float(out.loss)
    per_code = torch.nn.functional.cross_entropy(out.logits.reshape(-1, V), codes[:, 1:].reshape(-1),
                                                 reduction="none").view(Ns, C - 1)
    assert torch.allclose(ce, per_code, atol=1e-4, rtol=1e-4)


def test_attention_restatement_matches_dense_formula():
    torch.manual_seed(0)
    b, s, H, KV, hd = 2, 9, 4, 2, 8
    rope = tt.Llama3ScaledRoPE(hd, 64, 500000.0, 32.0)
    attn = tt.MultiHeadAttention(H * hd, H, KV, hd, rope, 64)
    x = torch.randn(b, s, H * hd)
    out = attn(x, x)
    q = rope(attn.q_proj(x).view(b, s, H, hd))
    k = rope(attn.k_proj(x).view(b, s, KV, hd))
    v = attn.v_proj(x).view(b, s, KV, hd)
    ref = torch.zeros(b, s, H, hd)
    for h in range(H):
        kv = h // (H // KV)                       # adjacent q heads share a kv head
        sc = torch.einsum("bid,bjd->bij", q[:, :, h], k[:, :, kv]) / math.sqrt(hd)
        sc = sc.masked_fill(~torch.tril(torch.ones(s, s, dtype=torch.bool)), float("-inf"))
        ref[:, :, h] = torch.einsum("bij,bjd->bid", sc.softmax(-1), v[:, :, kv])
    assert torch.allclose(out, attn.output_proj(ref.reshape(b, s, -1)), atol=1e-5)


def test_rope_is_interleaved_pair_rotation():
    rope = tt.Llama3ScaledRoPE(8, 16, 500000.0, 32.0)
    x = torch.randn(1, 4, 2, 8)
    y = rope(x)
    c, s_ = rope.cache[:4, :, 0], rope.cache[:4, :, 1]
    for j in range(4):
        x0, x1 = x[0, :, :, 2 * j], x[0, :, :, 2 * j + 1]
        assert torch.allclose(y[0, :, :, 2 * j], x0 * c[:, j, None] - x1 * s_[:, j, None], atol=1e-6)
        assert torch.allclose(y[0, :, :, 2 * j + 1], x1 * c[:, j, None] + x0 * s_[:, j, None], atol=1e-6)


def test_lora_linear_math_and_zero_B_is_identity():
    base = torch.nn.Linear(16, 12, bias=False)
    lin = O.LoRALinear(base, r=4, alpha=8.0)
    x = torch.randn(3, 16)
    assert torch.equal(lin(x), F.linear(x, base.weight))            # B = 0 -> base output
    with torch.no_grad():
        lin.lora_A.normal_()
        lin.lora_B.normal_()
    ref = x @ (base.weight + 2.0 * lin.lora_B @ lin.lora_A).t()      # merge formula lora.py:140-153
    assert torch.allclose(lin(x), ref, atol=1e-5)


def test_decoder_alignment_position_i_predicts_code_i(tiny):
    """A7: changing code c_j must not change the loss of codebooks <= j (teacher forcing is causal)."""
    cfg, m = tiny
    b = O.synthetic_batch(cfg, 1, 16, seed=5)
    fidx = b["frame_idx"][:1]
    _, d0 = O.oracle_forward(m, b["input_tokens"], b["input_masks"], b["target_audio_tokens"], fidx)
    tgt = b["target_audio_tokens"].clone()
    j = 10
    bb, p = int(fidx[0, 0]), int(fidx[0, 1])
    tgt[bb, p, j] = (tgt[bb, p, j] + 1) % cfg.audio_vocab_size
    _, d1 = O.oracle_forward(m, b["input_tokens"], b["input_masks"], tgt, fidx)
    pc0, pc1 = d0["per_codebook_loss"], d1["per_codebook_loss"]
    assert torch.allclose(pc0[1:j], pc1[1:j], atol=1e-6)            # earlier codebooks unaffected
    assert not torch.allclose(pc0[j:], pc1[j:], atol=1e-6)          # own target + later inputs change


def _load_make_golden_generate():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_generate.py")
    spec = importlib.util.spec_from_file_location("_make_golden_generate", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_generate_frame_reproduces_reference_golden():
    """tests/golden/c1_tiny_generate.pt holds codes produced by the REFERENCE's own Model.generate_frame (KV caches,
    prompt + 4 single-frame steps, topk = 1); the oracle restatement (shim KV cache + OracleModel.generate_frame) must
    reproduce them exactly — everywhere, also where /root/reference is absent."""
    G = _load_make_golden_generate()
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_tiny_generate.pt"))
    torch.set_num_threads(1)
    cfg = O.cfg_tiny()
    om = O.OracleModel(cfg)
    O.init_weights(om, gold["weight_seed"], std=gold["weight_std"])
    tok, msk = G.prompt(cfg, gold["B"], gold["S"], gold["prompt_seed"])
    assert torch.equal(tok, gold["tokens"]) and torch.equal(msk, gold["mask"])
    got = G.run(om, tok, msk, gold["frames"].shape[1] - 1, gold["B"])
    assert torch.equal(got.long(), gold["frames"])


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_generate_frame_equals_reference_live():
    G = _load_make_golden_generate()
    torch.set_num_threads(1)
    rm, _ = R.load_reference()
    cfg = O.cfg_tiny()
    R.register_flavor(rm, "tiny-bb", cfg.backbone)
    R.register_flavor(rm, "tiny-dec", cfg.decoder)
    om = O.OracleModel(cfg)
    O.init_weights(om, 11, std=0.3)
    ref = rm.Model(rm.ModelArgs("tiny-bb", "tiny-dec", cfg.text_vocab_size, cfg.audio_vocab_size,
                                cfg.audio_num_codebooks))
    ref.load_state_dict(om.state_dict())
    tok, msk = G.prompt(cfg, 3, 9, 2)
    with torch.no_grad():
        want = G.run(ref, tok, msk, 3, 3)
    assert torch.equal(want.long(), G.run(om, tok, msk, 3, 3).long())
