"""CPU oracle for the lit-llama quantized decode path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the arithmetic of the reference hot path
 as plain functions over torch CPU tensors.  It exists to
*check* the CUDA path; nothing under `lit-llama_b200/` may import it.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` use it.

Pinning: the reference's own tests hold no golden vectors for quantization.py
(SURVEY.md section 8c); the pin is therefore the reference itself, executed in the
build container by `oracle/make_golden.py`, whose outputs are committed under
`tests/golden/` and compared against this restatement by
`tests/test_oracle_golden.py`.  The gptq.int4/int8 + model.py + generate.py part
is pinned that way.  The llm.int8 part restates the published bitsandbytes
LLM.int8() algorithm (bitsandbytes is unpinned in pyproject.toml:19, absent from
/root/reference and not installed) and is therefore **parity unpinned**.

Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# lit_llama/utils.py
# ----------------------------------------------------------------------------
def find_multiple(n: int, k: int) -> int:
    """lit_llama/utils.py:38-41."""
    r = n % k
    return n if r == 0 else n + (k - r)


MODEL_SIZES = {4096: "7B", 5120: "13B", 6656: "30B", 8192: "65B"}  # utils.py:20-25
CONFIGS = {  # model.py:42-47
    "7B": dict(n_layer=32, n_head=32, n_embd=4096),
    "13B": dict(n_layer=40, n_head=40, n_embd=5120),
    "30B": dict(n_layer=60, n_head=52, n_embd=6656),
    "65B": dict(n_layer=80, n_head=64, n_embd=8192),
}


def n_hidden_for(n_embd: int) -> int:
    """model.py:243-245: SwiGLU hidden width."""
    return find_multiple(int(2 * (4 * n_embd) / 3), 256)


# ----------------------------------------------------------------------------
# lit_llama/quantization.py : round-to-nearest parameters + packing
# ----------------------------------------------------------------------------
def rtn_params(w: Tensor, bits: int) -> Tuple[Tensor, Tensor]:
    """Per-output-row asymmetric min/max grid.  quantization.py:477-513
    (perchannel=True, sym=False).  Returns (scale, zero), both (out, 1)."""
    maxq = 2**bits - 1
    lo = torch.clamp(w.amin(dim=1), max=0.0)
    hi = torch.clamp(w.amax(dim=1), min=0.0)
    dead = (lo == 0) & (hi == 0)
    lo = torch.where(dead, torch.full_like(lo, -1.0), lo)
    hi = torch.where(dead, torch.full_like(hi, 1.0), hi)
    scale = (hi - lo) / maxq
    zero = torch.round(-lo / scale)
    return scale.reshape(-1, 1), zero.reshape(-1, 1)


def rtn_levels(w: Tensor, scale: Tensor, zero: Tensor, bits: int) -> Tensor:
    """Integer levels q = clamp(round(w/scale)+zero, 0, maxq).  quantization.py:471-475."""
    return torch.clamp(torch.round(w / scale) + zero, 0, 2**bits - 1)


def pack_levels(levels: Tensor, bits: int) -> Tensor:
    """levels (out, in) in [0, 2^bits) -> uint8 (out, in/epb) stored with strides
    (1, out), i.e. memory is row-major (in/epb, out).  Entry `nr` of a byte sits at
    bit nr*bits and holds column epb*j+nr.  quantization.py:350-359, 386-390."""
    epb = 8 // bits
    lv = levels.to(torch.uint8)
    out, inf = lv.shape
    packed = torch.zeros((out, inf // epb), dtype=torch.uint8)
    for nr in range(epb):
        packed |= lv[:, nr::epb] << (nr * bits)
    return packed.t().contiguous().t()


def pack_weight(w: Tensor, scales: Tensor, zeros: Tensor, bits: int, tile_cols: int) -> Tensor:
    """ColBlockQuantizedLinear.pack_weight, quantization.py:376-390: divide by the
    group's scale, add the zero, clamp, *truncate* to uint8 (no rounding: the
    caller passes scale*(q-zero) so the quotient is integral up to fp error)."""
    w = w.clone()
    for g in range(scales.size(1)):
        sl = slice(g * tile_cols, (g + 1) * tile_cols)
        w[:, sl] /= scales[:, g : g + 1]
        w[:, sl] += zeros[:, g : g + 1]
    return pack_levels(w.clamp_(0, 2**bits - 1).to(torch.uint8), bits)


def unpack_levels(qw: Tensor, bits: int) -> Tensor:
    """uint8 (out, in/epb) -> int levels (out, in).  quantization.py:398-402."""
    epb = 8 // bits
    mask = (1 << bits) - 1
    out, packed_cols = qw.shape
    lv = torch.empty((out, packed_cols * epb), dtype=torch.uint8)
    for nr in range(epb):
        lv[:, nr::epb] = (qw >> (nr * bits)) & mask
    return lv


def dequant(qw: Tensor, scales: Tensor, zeros: Tensor, bits: int, tile_cols: int, dtype=torch.float32) -> Tensor:
    """ColBlockQuantizedLinear.get_weight, quantization.py:392-411.  The level is
    written into a `dtype` tensor, the zero subtracted and the scale multiplied *in
    that dtype* (so in bf16 each weight carries one bf16 rounding)."""
    w = unpack_levels(qw, bits).float().to(dtype)
    for g in range(scales.size(1)):
        sl = slice(g * tile_cols, (g + 1) * tile_cols)
        w[:, sl] -= zeros[:, g : g + 1]
        w[:, sl] *= scales[:, g : g + 1]
    return w


def qlinear(x: Tensor, qw: Tensor, scales: Tensor, zeros: Tensor, bits: int, tile_cols: int,
            bias: Optional[Tensor] = None) -> Tensor:
    """ColBlockQuantizedLinear.forward dense branch, quantization.py:422-423: the
    weight is re-materialised in x.dtype on every call, then F.linear."""
    return torch.nn.functional.linear(x, dequant(qw, scales, zeros, bits, tile_cols, x.dtype), bias)


def qlinear_exact(x: Tensor, qw: Tensor, scales: Tensor, zeros: Tensor, bits: int, tile_cols: int,
                  bias: Optional[Tensor] = None) -> Tensor:
    """Same contraction with exact fp32 dequant (level-zero)*scale and fp64
    accumulation - the arithmetic of the reference's GPU kernel
    (quantization.py:259-269: fp32 dequant, fp32 accumulate) without its TF32 dot.
    Returned in fp32; used as the tight numerical target for the CUDA kernels."""
    w = dequant(qw, scales.float(), zeros.float(), bits, tile_cols, torch.float32).double()
    y = x.double() @ w.t()
    if bias is not None:
        y = y + bias.double()
    return y.float()


# ----------------------------------------------------------------------------
# LLM.int8()  (restatement of bitsandbytes; parity unpinned - see module docstring)
# ----------------------------------------------------------------------------
def int8_quantize_weight(w: Tensor) -> Tuple[Tensor, Tensor]:
    """quantization.py:69-77 -> bnb.functional.double_quant on W.half(): row-wise
    absmax scaling to int8.  Returns CB int8 (out,in) and SCB fp32 (out,)."""
    wh = w.half().float()
    scb = wh.abs().amax(dim=1)
    cb = torch.round(wh * (127.0 / scb.clamp_min(1e-30)).unsqueeze(1)).clamp_(-127, 127).to(torch.int8)
    return cb, scb


def int8_linear(x: Tensor, cb: Tensor, scb: Tensor, threshold: float = 6.0) -> Tensor:
    """bnb.matmul / MatMul8bitLt.forward with has_fp16_weights=False,
    threshold=6.0 (quantization.py:47): activations go to fp16; columns of A that
    hold any |a| >= threshold are handled in fp16 against the dequantised weight
    columns, the rest row-wise absmax-quantised to int8; int32 GEMM; dequant by
    SCA*SCB/127^2; sum; cast back to x.dtype."""
    shape = x.shape
    a = x.reshape(-1, shape[-1]).half().float()
    outlier_cols = (a.abs() >= threshold).any(dim=0)
    a_in = a.clone()
    a_in[:, outlier_cols] = 0
    sca = a_in.abs().amax(dim=1)
    ca = torch.round(a_in * (127.0 / sca.clamp_min(1e-30)).unsqueeze(1)).clamp_(-127, 127)
    acc = ca.double() @ cb.double().t()  # exact int32 contraction
    y = (acc * (sca.double().unsqueeze(1) * scb.double().unsqueeze(0) / (127.0 * 127.0))).float()
    y = y.half().float()
    if outlier_cols.any():
        w_sub = (cb[:, outlier_cols].float() * (scb / 127.0).unsqueeze(1)).half().float()
        y = (y + (a[:, outlier_cols] @ w_sub.t()).half().float()).half().float()
    return y.to(x.dtype).reshape(*shape[:-1], cb.shape[0])


# ----------------------------------------------------------------------------
# lit_llama/model.py
# ----------------------------------------------------------------------------
def rmsnorm(x: Tensor, scale: Tensor, eps: float = 1e-5) -> Tensor:
    """RMSNorm.forward, model.py:270-277, evaluated in x.dtype with no upcast:
    mean(x*x) -> +eps -> rsqrt -> x*that -> scale*that."""
    ms = torch.mean(x * x, dim=-1, keepdim=True)
    return scale * (x * torch.rsqrt(ms + eps))


def rope_table(seq_len: int, n_elem: int, base: int = 10000) -> Tensor:
    """build_rope_cache, model.py:280-303, for the integer-`idx` call made by
    LLaMA.forward (model.py:128-134): fp32 (seq_len, n_elem/2, 2) of (cos, sin)."""
    theta = 1.0 / (base ** (torch.arange(0, n_elem, 2, dtype=torch.float32) / n_elem))
    ang = torch.outer(torch.arange(seq_len, dtype=torch.float32), theta)
    return torch.stack((torch.cos(ang), torch.sin(ang)), dim=-1)


def rope_apply(x: Tensor, rows: Tensor) -> Tensor:
    """apply_rope, model.py:306-323.  x (B,T,nh,hs); rows (T,hs/2,2) fp32.
    Interleaved pairs rotate in fp32, result cast back to x.dtype."""
    B, T, nh, hs = x.shape
    xp = x.float().reshape(B, T, nh, hs // 2, 2)
    c = rows[:T, :, 0].reshape(1, T, 1, hs // 2)
    s = rows[:T, :, 1].reshape(1, T, 1, hs // 2)
    even = xp[..., 0] * c - xp[..., 1] * s
    odd = xp[..., 1] * c + xp[..., 0] * s
    return torch.stack((even, odd), dim=-1).reshape(B, T, nh, hs).to(x.dtype)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Tensor) -> Tensor:
    """model.py:230 semantics: softmax(q k^T / sqrt(hs) masked) v.  Evaluated in
    fp32 from the stored-dtype operands, one rounding to q.dtype at the end."""
    att = (q.float() @ k.float().transpose(-2, -1)) * (1.0 / math.sqrt(q.size(-1)))
    att = att.masked_fill(~mask, float("-inf"))
    return (torch.softmax(att, dim=-1) @ v.float()).to(q.dtype)


@dataclass
class QLin:
    """One linear layer of the model in whichever storage the mode uses."""
    kind: str  # "dense" | "gptq" | "gptq_exact" | "int8"
    weight: Optional[Tensor] = None  # dense
    qw: Optional[Tensor] = None
    scales: Optional[Tensor] = None
    zeros: Optional[Tensor] = None
    bits: int = 4
    tile_cols: int = -1
    cb: Optional[Tensor] = None
    scb: Optional[Tensor] = None

    def __call__(self, x: Tensor) -> Tensor:
        if self.kind == "dense":
            return torch.nn.functional.linear(x, self.weight)
        if self.kind == "gptq":
            return qlinear(x, self.qw, self.scales, self.zeros, self.bits, self.tile_cols)
        if self.kind == "gptq_exact":  # arithmetic of the reference's GPU kernel: fp32 dequant, fp32+ accumulate
            return qlinear_exact(x, self.qw, self.scales, self.zeros, self.bits, self.tile_cols).to(x.dtype)
        return int8_linear(x, self.cb, self.scb)


@dataclass
class OracleLLaMA:
    """Functional restatement of lit_llama.model.LLaMA for inference with a KV cache."""
    n_layer: int
    n_head: int
    n_embd: int
    block_size: int
    padded_vocab_size: int
    wte: Tensor
    lm_head: QLin
    ln_f: Tensor
    layers: List[Dict[str, object]]  # keys: rms_1, rms_2 (Tensor); c_attn, c_proj, c_fc1, c_fc2, mlp_proj (QLin)
    rope: Optional[Tensor] = None
    kv: List[Tuple[Tensor, Tensor]] = field(default_factory=list)

    @staticmethod
    def from_state_dict(sd: Dict[str, Tensor], n_layer: int, n_head: int, block_size: int,
                        mode: Optional[str] = None, exact_linears: bool = False) -> "OracleLLaMA":
        """Builds from a reference-format state dict (keys as produced by
        lit_llama.model.LLaMA under utils.quantization(mode)).  `exact_linears` selects the
        arithmetic of the reference's GPU branch (quantization.py:259-269: fp32 dequant and
        accumulate, no per-weight bf16 rounding) instead of its dense CPU branch (:392-423)."""

        def lin(prefix: str) -> QLin:
            if prefix + ".quant_weight" in sd:
                qw = sd[prefix + ".quant_weight"]
                sc = sd[prefix + ".scales"]
                bits = 4 if mode in (None, "gptq.int4") else 8
                epb = 8 // bits
                in_features = qw.shape[1] * epb
                n_groups = sc.shape[1]
                tile_cols = in_features if n_groups == 1 else -(-in_features // n_groups)
                return QLin("gptq_exact" if exact_linears else "gptq", qw=qw, scales=sc, zeros=sd[prefix + ".zeros"], bits=bits,
                            tile_cols=tile_cols)
            w = sd[prefix + ".weight"]
            if mode == "llm.int8":
                cb, scb = int8_quantize_weight(w)
                return QLin("int8", cb=cb, scb=scb)
            return QLin("dense", weight=w)

        wte = sd["transformer.wte.weight"]
        layers = []
        for i in range(n_layer):
            p = f"transformer.h.{i}."
            layers.append(dict(
                rms_1=sd[p + "rms_1.scale"], rms_2=sd[p + "rms_2.scale"],
                c_attn=lin(p + "attn.c_attn"), c_proj=lin(p + "attn.c_proj"),
                c_fc1=lin(p + "mlp.c_fc1"), c_fc2=lin(p + "mlp.c_fc2"), mlp_proj=lin(p + "mlp.c_proj"),
            ))
        return OracleLLaMA(n_layer=n_layer, n_head=n_head, n_embd=wte.shape[1], block_size=block_size,
                           padded_vocab_size=wte.shape[0], wte=wte, lm_head=lin("lm_head"),
                           ln_f=sd["transformer.ln_f.scale"], layers=layers)

    def reset_cache(self) -> None:
        """model.py:140-145."""
        self.kv = []

    def _attn(self, x: Tensor, lay, rope: Tensor, mask: Tensor, S: int, input_pos: Optional[Tensor], li: int) -> Tensor:
        """CausalSelfAttention.forward, model.py:185-237."""
        B, T, C = x.shape
        hs = C // self.n_head
        q, k, v = lay["c_attn"](x).split(C, dim=2)
        q = rope_apply(q.view(B, T, self.n_head, hs), rope).transpose(1, 2)
        k = rope_apply(k.view(B, T, self.n_head, hs), rope).transpose(1, 2)
        v = v.view(B, T, self.n_head, hs).transpose(1, 2)
        if input_pos is not None:
            ck, cv = self.kv[li]
            if int(input_pos[-1]) >= S:  # model.py:214-218: sliding window by one slot
                input_pos = torch.tensor(S - 1)
                ck = torch.roll(ck, -1, dims=2)
                cv = torch.roll(cv, -1, dims=2)
            k = ck.index_copy(2, input_pos.reshape(-1), k)
            v = cv.index_copy(2, input_pos.reshape(-1), v)
            self.kv[li] = (k, v)
        y = sdpa(q, k, v, mask)
        return lay["c_proj"](y.transpose(1, 2).contiguous().view(B, T, C))

    def forward(self, idx: Tensor, max_seq_length: Optional[int] = None, input_pos: Optional[Tensor] = None) -> Tensor:
        """LLaMA.forward, model.py:76-122 (+ Block.forward :156-168, MLP.forward :251-254)."""
        B, T = idx.shape
        S = self.block_size if max_seq_length is None else max_seq_length
        assert T <= S <= self.block_size
        hs = self.n_embd // self.n_head
        if self.rope is None:
            self.rope = rope_table(self.block_size, hs)
        tril = torch.tril(torch.ones(self.block_size, self.block_size, dtype=torch.bool))
        if input_pos is not None:
            rope = self.rope.index_select(0, input_pos)
            mask = tril.index_select(0, input_pos)[:, :S].reshape(1, 1, T, S)
        else:
            rope = self.rope[:T]
            mask = tril[:T, :T].reshape(1, 1, T, T)
        x = self.wte[idx]
        if input_pos is not None and not self.kv:
            shape = (B, self.n_head, S, hs)
            self.kv = [(torch.zeros(shape, dtype=x.dtype), torch.zeros(shape, dtype=x.dtype)) for _ in range(self.n_layer)]
        for li, lay in enumerate(self.layers):
            x = x + self._attn(rmsnorm(x, lay["rms_1"]), lay, rope, mask, S, input_pos, li)
            h = rmsnorm(x, lay["rms_2"])
            x = x + lay["mlp_proj"](torch.nn.functional.silu(lay["c_fc1"](h)) * lay["c_fc2"](h))
        return self.lm_head(rmsnorm(x, self.ln_f))


def generate(model, idx: Tensor, max_new_tokens: int, *, max_seq_length: Optional[int] = None,
             temperature: float = 1.0, top_k: Optional[int] = None, eos_id: Optional[int] = None) -> Tensor:
    """generate.py:20-91.  `model` needs .forward(idx, max_seq_length, input_pos)
    and .block_size.  Sampling uses torch.multinomial so the RNG stream is the
    reference's."""
    T = idx.size(0)
    T_new = T + max_new_tokens
    if max_seq_length is None:
        max_seq_length = min(T_new, model.block_size)
    out = torch.empty(T_new, dtype=idx.dtype)
    out[:T] = idx
    input_pos = torch.arange(0, T)
    for _ in range(max_new_tokens):
        x = out.index_select(0, input_pos).view(1, -1)
        logits = model.forward(x, max_seq_length, input_pos)[0, -1] / temperature
        if top_k is not None:
            v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
            logits = torch.where(logits < v[[-1]], -float("inf"), logits)
        probs = torch.softmax(logits, dim=-1)
        nxt = torch.multinomial(probs, num_samples=1).to(dtype=idx.dtype)
        input_pos = input_pos[-1:] + 1
        out = out.index_copy(0, input_pos, nxt)
        if eos_id is not None and int(nxt) == eos_id:
            return out[: int(input_pos)]
    return out


# ----------------------------------------------------------------------------
# Synthetic weights - shared by tests, smoke and bench
# ----------------------------------------------------------------------------
def synth_state_dict(n_layer: int, n_head: int, n_embd: int, vocab_size: int, mode: Optional[str],
                     dtype=torch.bfloat16, seed: int = 1234, tile_cols: int = -1) -> Dict[str, Tensor]:
    """Random-init weights with the reference initialiser's statistics
    (model.py:70-74: N(0, 0.02/sqrt(2 n_layer))), RMSNorm scales near 1, quantised
    round-to-nearest with the reference formulas (rtn_params / rtn_levels /
    pack_levels).  Keys, shapes, dtypes and strides match what the reference model
    holds under utils.quantization(mode)."""
    g = torch.Generator().manual_seed(seed)
    std = 0.02 / math.sqrt(2 * n_layer)
    V = find_multiple(vocab_size, 64)
    nh = n_hidden_for(n_embd)
    sd: Dict[str, Tensor] = {}

    def put_linear(prefix: str, out_f: int, in_f: int) -> None:
        w = torch.randn(out_f, in_f, generator=g) * std
        if mode in ("gptq.int4", "gptq.int8"):
            bits = 4 if mode == "gptq.int4" else 8
            tc = in_f if tile_cols == -1 else tile_cols
            ng = -(-in_f // tc)
            scales = torch.empty(out_f, ng)
            zeros = torch.empty(out_f, ng)
            lv = torch.empty(out_f, in_f)
            for gi in range(ng):
                sl = slice(gi * tc, (gi + 1) * tc)
                s, z = rtn_params(w[:, sl], bits)
                s = s.to(dtype).float()  # the stored scale is what every consumer sees
                scales[:, gi : gi + 1], zeros[:, gi : gi + 1] = s, z
                lv[:, sl] = rtn_levels(w[:, sl], s, z, bits)
            sd[prefix + ".quant_weight"] = pack_levels(lv, bits)
            sd[prefix + ".scales"] = scales.to(dtype)
            sd[prefix + ".zeros"] = zeros.to(dtype)
        else:
            sd[prefix + ".weight"] = w.to(dtype)

    put_linear("lm_head", V, n_embd)
    sd["transformer.wte.weight"] = (torch.randn(V, n_embd, generator=g) * 0.02).to(dtype)
    for i in range(n_layer):
        p = f"transformer.h.{i}."
        sd[p + "rms_1.scale"] = (1.0 + 0.1 * torch.randn(n_embd, generator=g)).to(dtype)
        put_linear(p + "attn.c_attn", 3 * n_embd, n_embd)
        put_linear(p + "attn.c_proj", n_embd, n_embd)
        sd[p + "rms_2.scale"] = (1.0 + 0.1 * torch.randn(n_embd, generator=g)).to(dtype)
        put_linear(p + "mlp.c_fc1", nh, n_embd)
        put_linear(p + "mlp.c_fc2", nh, n_embd)
        put_linear(p + "mlp.c_proj", n_embd, nh)
    sd["transformer.ln_f.scale"] = (1.0 + 0.1 * torch.randn(n_embd, generator=g)).to(dtype)
    return sd
