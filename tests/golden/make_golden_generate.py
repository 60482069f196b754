"""Mints tests/golden/c1_tiny_generate.pt: outputs of the REFERENCE's own ``Model.generate_frame``
(/root/reference/src/csm/models/model.py:140-195, imported verbatim through oracle/torchtune_shim.py) on the tiny model:
a 12-frame prompt followed by four single-frame steps, topk = 1 (sample_topk then returns the argmax).  Refuses to write
unless the oracle restatement (oracle/csm_oracle.py::OracleModel.generate_frame) produces the same codes.
Run in the build container only:   python tests/golden/make_golden_generate.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import csm_oracle as O, reference_loader as R   # noqa: E402


def prompt(cfg, B, S, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(B, S, 33, dtype=torch.int64)
    msk = torch.zeros(B, S, 33, dtype=torch.bool)
    st = S // 2
    tok[:, :st, 32] = torch.randint(0, cfg.text_vocab_size, (B, st), generator=g)
    msk[:, :st, 32] = True
    tok[:, st:, :32] = torch.randint(0, cfg.audio_vocab_size, (B, S - st, 32), generator=g)
    msk[:, st:, :32] = True
    return tok, msk


def run(model, tok, msk, steps, B):
    """prompt -> frame, then `steps` more frames, each fed back as one audio frame (generator.py:150-170)."""
    model.setup_caches(B)
    S = tok.shape[1]
    pos = torch.arange(S).unsqueeze(0).repeat(B, 1)
    frames = [model.generate_frame(tok, msk, pos, 0.9, 1)]
    for i in range(steps):
        t = torch.cat([frames[-1].long(), torch.zeros(B, 1, dtype=torch.int64)], dim=1).unsqueeze(1)
        m = torch.cat([torch.ones(B, 32, dtype=torch.bool), torch.zeros(B, 1, dtype=torch.bool)], dim=1).unsqueeze(1)
        frames.append(model.generate_frame(t, m, torch.full((B, 1), S + i), 0.9, 1))
    return torch.stack(frames, 1)          # [B, steps + 1, 32]


def main():
    assert R.available(), "reference tree not present"
    torch.manual_seed(0)
    torch.set_num_threads(1)
    rm, _ = R.load_reference()
    cfg = O.cfg_tiny()
    R.register_flavor(rm, "tiny-bb", cfg.backbone)
    R.register_flavor(rm, "tiny-dec", cfg.decoder)
    om = O.OracleModel(cfg)
    O.init_weights(om, 0, std=0.3)          # wide logits: decisive argmax
    ref = rm.Model(rm.ModelArgs("tiny-bb", "tiny-dec", cfg.text_vocab_size, cfg.audio_vocab_size,
                                cfg.audio_num_codebooks))
    ref.load_state_dict(om.state_dict())
    B, S = 2, 12
    tok, msk = prompt(cfg, B, S, 5)
    with torch.no_grad():
        want = run(ref, tok, msk, 4, B)
    got = run(om, tok, msk, 4, B)
    assert torch.equal(want.long(), got.long()), "oracle generate_frame != reference generate_frame"
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_tiny_generate.pt")
    torch.save({"cfg": "tiny", "weight_seed": 0, "weight_std": 0.3, "B": B, "S": S, "prompt_seed": 5,
                "tokens": tok, "mask": msk, "frames": want.long(), "temperature": 0.9, "topk": 1}, path)
    print("oracle == reference generate_frame on", tuple(want.shape), "codes; wrote", path)


if __name__ == "__main__":
    main()
