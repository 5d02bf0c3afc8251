"""Plug the B200 modules into an UNMODIFIED reference checkout.

The reference resolves its classes through module globals at construction time
(lit_llama/model.py:61,152,154 and the torch.nn.Linear swap of utils.py:156-162 - the
same trick its own lora() context uses, lit_llama/lora.py:472-476).  After

    import lit_llama, lit_llama_b200
    lit_llama_b200.patch_reference(lit_llama)

the reference's own `generate.py` (`main()`, `--quantize gptq.int4`) builds B200
modules: `lit_llama.LLaMA`, `lit_llama.model.{LLaMA,Block,CausalSelfAttention,MLP,
RMSNorm,apply_rope,build_rope_cache}`, `lit_llama.quantization.{ColBlockQuantizedLinear,
Linear8bitLt}` and `lit_llama.utils.{quantization,EmptyInitOnDevice,lazy_load}`
all point at this package.
"""
import sys


def patch_reference(lit_llama_module=None):
    from . import model as m, quantization as q, utils as u

    if lit_llama_module is None:
        import lit_llama as lit_llama_module  # noqa: N813
    ref_model = sys.modules[lit_llama_module.__name__ + ".model"]
    ref_utils = sys.modules[lit_llama_module.__name__ + ".utils"]
    ref_quant = sys.modules.get(lit_llama_module.__name__ + ".quantization")
    if ref_quant is None:
        import importlib

        ref_quant = importlib.import_module(lit_llama_module.__name__ + ".quantization")
    saved = {}
    for name in ("LLaMA", "LLaMAConfig", "Block", "CausalSelfAttention", "MLP", "RMSNorm", "apply_rope", "build_rope_cache"):
        saved[("model", name)] = getattr(ref_model, name)
        setattr(ref_model, name, getattr(m, name))
        if hasattr(lit_llama_module, name):
            saved[("pkg", name)] = getattr(lit_llama_module, name)
            setattr(lit_llama_module, name, getattr(m, name))
    saved[("quant", "ColBlockQuantizedLinear")] = ref_quant.ColBlockQuantizedLinear
    ref_quant.ColBlockQuantizedLinear = q.ColBlockQuantizedLinear
    from . import int8 as i8

    saved[("quant", "Linear8bitLt")] = getattr(ref_quant, "Linear8bitLt", None)  # absent when bitsandbytes is not installed
    ref_quant.Linear8bitLt = i8.Linear8bitLt
    saved[("quant", "qlinear_4bit_weight")] = getattr(ref_quant, "qlinear_4bit_weight", None)
    ref_quant.qlinear_4bit_weight = q.qlinear_4bit_weight
    # The offline converter stays the reference's own (quantize/gptq.py + quantization.GPTQQuantizer, host-side torch,
    # out of the decode path): it builds `ColBlockQuantizedLinear` through the name patched above and calls
    # pack_weight(), so it converts straight into B200 modules.
    for name in ("EmptyInitOnDevice", "lazy_load"):
        saved[("utils", name)] = getattr(ref_utils, name, None)
        setattr(ref_utils, name, getattr(u, name))
    saved[("utils", "quantization")] = ref_utils.quantization
    ref_utils.quantization = u.quantization
    for mod in list(sys.modules.values()):  # scripts that did `from lit_llama.utils import quantization`
        if mod is not None and getattr(mod, "quantization", None) is saved[("utils", "quantization")]:
            setattr(mod, "quantization", u.quantization)
        if mod is not None and getattr(mod, "LLaMA", None) is saved[("model", "LLaMA")]:
            setattr(mod, "LLaMA", m.LLaMA)
    return saved
