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


class EmptyInitOnDevice:
    """utils.py:73-138: inside the context, tensors are created on `device` with `dtype`, `torch.nn.init.*`
    does nothing (parameters stay uninitialised until `load_state_dict`) and, with `quantization_mode`,
    `torch.nn.Linear` constructs the quantized class.  Same constructor and errors as the reference
    (`llm.int8` off the GPU: ValueError; unknown mode: RuntimeError); built from torch's own default-device
    context instead of a TorchFunctionMode.

        with EmptyInitOnDevice(device="cuda", dtype=torch.bfloat16, quantization_mode="gptq.int4"):
            model = LLaMA.from_name("7B")
        model.load_state_dict(checkpoint)
    """

    def __init__(self, device=None, dtype=None, quantization_mode=None):
        if quantization_mode == "llm.int8" and (device is None or torch.device(device).type != "cuda"):
            raise ValueError("Quantization is only supported on the GPU.")
        if quantization_mode not in (None, "llm.int8", "gptq.int4", "gptq.int8"):
            raise RuntimeError(f"unknown quantization mode {quantization_mode}")
        self.device, self.dtype, self.quantization_mode = device, dtype, quantization_mode
        self._stack = None

    def __enter__(self):
        from contextlib import ExitStack

        stack = ExitStack()
        try:
            if self.quantization_mode is not None:
                stack.enter_context(quantization(self.quantization_mode))
            if self.device is not None:
                stack.enter_context(torch.device(self.device))
            if self.dtype is not None:
                previous = torch.get_default_dtype()
                torch.set_default_dtype(self.dtype)
                stack.callback(torch.set_default_dtype, previous)
            # initialisers become no-ops that hand back their tensor
            import torch.nn.init as init

            for name in [n for n in dir(init) if n.endswith("_") and not n.startswith("_") and callable(getattr(init, n))]:
                original = getattr(init, name)
                setattr(init, name, lambda tensor, *a, **k: tensor)
                stack.callback(setattr, init, name, original)
        except BaseException:
            stack.close()
            raise
        self._stack = stack
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        stack, self._stack = self._stack, None
        return stack.__exit__(exc_type, exc_val, exc_tb)


class lazy_load:
    """utils.py:332-344: `with lazy_load(path) as checkpoint:` yields the state dict of a `torch.save`d file
    without reading the tensors up front.  The reference unpickles into placeholder tensors; torch's
    memory-mapped load gives the same behaviour (file-backed storages, pages read on first access)."""

    def __init__(self, fn):
        self.sd = torch.load(str(fn), map_location="cpu", mmap=True, weights_only=True)

    def __enter__(self):
        return self.sd

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.sd = None
