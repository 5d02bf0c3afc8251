"""The pieces of lit_llama/utils.py that sit on the decode path."""
import functools
from contextlib import contextmanager

import torch

llama_model_sizes = {4096: "7B", 5120: "13B", 6656: "30B", 8192: "65B"}  # utils.py:20-25


def llama_model_lookup(checkpoint: dict) -> str:
    """utils.py:29-35: the model name from the embedding width."""
    embedding_size = checkpoint["transformer.wte.weight"].shape[1]
    return llama_model_sizes[embedding_size]


def find_multiple(n: int, k: int) -> int:
    """utils.py:38-41."""
    if n % k == 0:
        return n
    return n + k - (n % k)


@contextmanager
def quantization(mode: str = None):
    """utils.py:141-162: while active, `torch.nn.Linear` constructs the quantized class
    of `mode` ('llm.int8', 'gptq.int4', 'gptq.int8'); unknown modes raise ValueError."""
    quantized_linear_cls = None
    if mode == "llm.int8":
        from .int8 import Linear8bitLt

        quantized_linear_cls = Linear8bitLt
    elif mode == "gptq.int4":
        from .quantization import ColBlockQuantizedLinear

        quantized_linear_cls = functools.partial(ColBlockQuantizedLinear, bits=4, tile_cols=-1)
    elif mode == "gptq.int8":
        from .quantization import ColBlockQuantizedLinear

        quantized_linear_cls = functools.partial(ColBlockQuantizedLinear, bits=8, tile_cols=-1)
    elif mode is not None:
        raise ValueError(f"Unknown quantization mode: {mode}")

    enabled = mode is not None
    torch_linear_cls = torch.nn.Linear
    if enabled:
        torch.nn.Linear = quantized_linear_cls
    try:
        yield
    finally:
        if enabled:
            torch.nn.Linear = torch_linear_cls
