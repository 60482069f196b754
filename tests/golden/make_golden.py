"""Mints tests/golden/c1_tiny.pt.  Run HERE (build container) only: it imports the reference's
own model.py / compute_loss verbatim from /root/reference (oracle/reference_loader.py) and refuses
to write anything unless the oracle restatement agrees with them bit-for-bit on the semantic path.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import csm_oracle as O, reference_loader as R   # noqa: E402


def grads_of(model, names=None):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()
            if p.grad is not None and (names is None or n in names)}


def main():
    assert R.available(), "reference tree not present"
    torch.manual_seed(0)
    torch.set_num_threads(1)          # deterministic reduction order
    rm, ru = R.load_reference()
    cfg = O.cfg_tiny()
    R.register_flavor(rm, "tiny-bb", cfg.backbone)
    R.register_flavor(rm, "tiny-dec", cfg.decoder)
    B, S = 2, 32
    batch = O.synthetic_batch(cfg, B, S, seed=1234)
    tok, msk, tgt, fidx = (batch[k] for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"))

    om = O.OracleModel(cfg)
    O.init_weights(om, 0)
    base_state = {k: v.clone() for k, v in om.state_dict().items()}

    # ---- the reference itself (fp32, full fine-tune, semantic term only: utils.py:98-119)
    ref = rm.Model(rm.ModelArgs("tiny-bb", "tiny-dec", cfg.text_vocab_size, cfg.audio_vocab_size,
                                cfg.audio_num_codebooks))
    ref.load_state_dict(base_state)
    ref.backbone_causal_mask = rm._create_causal_mask(S, torch.device("cpu"))       # model.py:137
    ref._index_causal_mask = lambda m, p: rm._index_causal_mask(m, p)              # test_training.py:99
    ref_loss, ref_d = ru.compute_loss(ref, tok, msk, tgt)
    ref_loss.backward()
    ref_embed = ref._embed_tokens(tok)
    ref_h = (ref_embed * msk.unsqueeze(-1)).sum(dim=2)
    ref_cmask = rm._index_causal_mask(ref.backbone_causal_mask, torch.arange(S).unsqueeze(0).repeat(B, 1))

    # ---- oracle, semantic path only, must equal the reference bit for bit
    ol, od = O.oracle_forward(om, tok, msk, tgt, None)
    ol.backward()
    assert torch.equal(ol, ref_loss) and torch.equal(od["semantic_loss"], ref_d["semantic_loss"])
    assert torch.equal(om._embed_tokens(tok), ref_embed)
    rg, og = grads_of(ref), grads_of(om)
    assert set(rg) == set(og)
    for n in rg:
        assert torch.equal(rg[n], og[n]), n
    assert ref.audio_head.grad is None          # decoder receives no gradient in the reference
    print("oracle == reference on the semantic path (loss, embeds, every gradient): bit-exact")

    out = {"cfg": "tiny", "B": B, "S": S, "batch_seed": 1234, "weight_seed": 0,
           "input_tokens": tok, "input_masks": msk, "target_audio_tokens": tgt, "frame_idx": fidx,
           "gather_idx": O.gather_indices(tok, cfg.audio_vocab_size, cfg.audio_num_codebooks),
           "ref_h_embed": ref_h, "ref_causal_mask": ref_cmask,
           "ref_loss": ref_loss.detach(), "ref_semantic_loss": ref_d["semantic_loss"].detach(),
           "ref_grad_codebook0_head": rg["codebook0_head.weight"],
           "ref_grad_q_proj_l0": rg["backbone.layers.0.attn.q_proj.weight"]}

    # ---- oracle with the acoustic (decoder) term, full fine-tune, fp32
    om.zero_grad()
    l, d = O.oracle_forward(om, tok, msk, tgt, fidx)
    l.backward()
    g = grads_of(om)
    out["fullft_fp32"] = {"loss": l.detach(), "semantic_loss": d["semantic_loss"].detach(),
                          "acoustic_loss": d["acoustic_loss"].detach(),
                          "per_codebook_loss": d["per_codebook_loss"],
                          "grads": {n: g[n] for n in ("audio_head", "projection.weight", "codebook0_head.weight",
                                                      "backbone.layers.1.mlp.w2.weight",
                                                      "decoder.layers.0.attn.k_proj.weight",
                                                      "backbone.norm.scale")}}

    # ---- LoRA r=8 q/v (BASELINE config 1), fp32 and bf16
    for tag, dtype in (("lora_fp32", torch.float32), ("lora_bf16", torch.bfloat16)):
        m = O.OracleModel(cfg)
        m.load_state_dict(base_state)
        m = m.to(dtype)
        O.apply_lora(m, r=8, alpha=16.0, seed=1)
        l, d = O.oracle_forward(m, tok, msk, tgt, fidx)
        l.backward()
        out[tag] = {"loss": l.detach().float(), "semantic_loss": d["semantic_loss"].detach().float(),
                    "acoustic_loss": d["acoustic_loss"].detach().float(),
                    "per_codebook_loss": d["per_codebook_loss"],
                    "grads": {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}}
        print(tag, float(l), "trainable tensors:", len(out[tag]["grads"]))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
