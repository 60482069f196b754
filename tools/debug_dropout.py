import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import test_model_parity_gpu as T
from csm import ops
from csm.models import lora as plora
from oracle import csm_oracle as O
cuda = torch.device("cuda:0")
p_drop = 0.25
targets = sys.argv[1].split(",") if len(sys.argv) > 1 else ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
prod, cfg = T._product_model("small")
orc = O.OracleModel(cfg); O.init_weights(orc, 0)
orc, prod = orc.to(torch.bfloat16), prod.to(torch.bfloat16)
O.apply_lora(orc, r=8, alpha=16.0, target_modules=targets, seed=1)
plora.apply_lora(prod, r=8, alpha=16.0, target_modules=targets, seed=4, dropout=p_drop)
prod.load_state_dict(orc.state_dict(), strict=True)
prod = prod.to(cuda).train()
B, S = 2, 128
batch = O.synthetic_batch(cfg, B, S, seed=77)
tok, msk, tgt, fidx = (batch[k] for k in ("input_tokens", "input_masks", "target_audio_tokens", "frame_idx"))
seed = torch.ones(1, dtype=torch.int64, device=cuda)
C = cfg.audio_num_codebooks
names = {"q_proj": "attn", "k_proj": "attn", "v_proj": "attn", "o_proj": "attn", "gate_proj": "mlp", "up_proj": "mlp", "down_proj": "mlp"}
for stack, rows, shape3 in ((orc.backbone, B * S, (B, S)), (orc.decoder, fidx.shape[0] * C, (fidx.shape[0], C))):
    for li, layer in enumerate(stack.layers):
        a, f = layer.attn, layer.mlp
        for salt, mods in ((4 * li, (a.q_proj, a.k_proj, a.v_proj)), (4 * li + 1, (f.w1, f.w3)), (4 * li + 2, (a.output_proj,)), (4 * li + 3, (f.w2,))):
            for m in mods:
                if isinstance(m, O.LoRALinear):
                    ones = torch.ones(rows, m.weight.shape[1], dtype=torch.bfloat16, device=cuda)
                    m.keep = ops.lora_dropout(ones, p_drop, seed, salt).float().cpu().view(*shape3, -1)
ol, od = O.oracle_forward(orc, tok, msk, tgt, fidx); ol.backward()
pl, pd = prod(tok.to(cuda), msk.to(cuda), tgt.to(cuda), frame_idx=fidx.to(cuda)); pl.backward(); torch.cuda.synchronize()
print("loss", float(pl), float(ol), "seeds", prod.backbone._lora_seed.item(), prod.decoder._lora_seed.item())
named = dict(prod.named_parameters())
rows = []
for n, q in orc.named_parameters():
    if q.grad is not None:
        rows.append((float(F.cosine_similarity(q.grad.float().flatten(), named[n].grad.float().cpu().flatten(), dim=0)), n))
for c, n in sorted(rows)[:12]: print(f"{c:.5f} {n}")
print("...", sorted(rows)[-1])
