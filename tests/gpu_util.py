"""Helpers shared by the -m gpu parity tests."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from diag import gemv_call, rand_q4, ref_linear, relerr, tc_call, tile, tile_mma  # noqa: E402,F401


def build_tiny(dev, cfg, mode="gptq.int4", seed=1234, tile_cols=-1, exact_linears=False):
    import lit_llama_b200 as P
    from lit_llama_b200.utils import quantization
    from oracle import llama_oracle as O

    sd = O.synth_state_dict(cfg["n_layer"], cfg["n_head"], cfg["n_embd"], cfg["vocab_size"], None if mode == "llm.int8" else mode,
                            dtype=torch.bfloat16, seed=seed)  # llm.int8 loads a float checkpoint and quantises on load
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev), quantization(mode):
            model = P.LLaMA(P.LLaMAConfig(**cfg))
    finally:
        torch.set_default_dtype(prev)
    model.load_state_dict(sd)
    oracle = O.OracleLLaMA.from_state_dict(sd, cfg["n_layer"], cfg["n_head"], cfg["block_size"], mode, exact_linears=exact_linears)
    return model.eval(), oracle, sd
