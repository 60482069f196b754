"""ORACLE — test infrastructure only.

Loads the reference's own Python sources *verbatim, where they lie* under /root/reference
(never copied into this repo) so that the restatement in csm_oracle.py can be validated
against them and golden vectors can be minted (tests/golden/make_golden.py).
/root/reference exists only in the build container; on the GPU box ``available()`` is False
and nothing here is used.

  * src/csm/models/model.py       imported by file path after oracle.torchtune_shim.install()
  * src/csm/training/utils.py     imported by file path (its top-level imports are stdlib/torch/numpy)
"""
from __future__ import annotations

import importlib.util
import os
import sys

REF_ROOT = os.environ.get("CSM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/csm/models/model.py"))


def _load(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (model_module, training_utils_module) of the unmodified reference."""
    from . import torchtune_shim
    torchtune_shim.install()
    ref_model = _load("_csm_ref_model", "src/csm/models/model.py")
    ref_utils = _load("_csm_ref_training_utils", "src/csm/training/utils.py")
    return ref_model, ref_utils


def register_flavor(ref_model, name: str, c) -> None:
    """Adds an entry to the reference's FLAVORS registry (model.py:45-48) for a non-1B shape,
    built through the same torchtune builder the reference's own flavors call."""
    from torchtune.models import llama3_2

    def build():
        return llama3_2.llama3_2(vocab_size=8, num_layers=c.num_layers, num_heads=c.num_heads,
                                 num_kv_heads=c.num_kv_heads, embed_dim=c.embed_dim,
                                 max_seq_len=c.max_seq_len, intermediate_dim=c.intermediate_dim,
                                 attn_dropout=0.0, norm_eps=c.norm_eps, rope_base=c.rope_base,
                                 scale_factor=c.scale_factor)
    ref_model.FLAVORS[name] = build
