"""Offline GPTQ conversion (lit_llama_b200.gptq) against fixtures produced by the
unmodified reference's GPTQQuantizer (oracle/make_golden.py: golden_gptq) -- host-side code, runs on the CPU."""
import os
import sys

import pytest
import torch

from conftest import load_golden

import lit_llama_b200 as P
from lit_llama_b200.gptq import GPTQQuantizer, affine_grid, snap_to_grid
from oracle import llama_oracle as O


def _dense(qm):
    """CPU dequantisation of a converted module (the product's get_weight is CUDA-only)."""
    return O.dequant(qm.quant_weight, qm.scales, qm.zeros, qm.bits, qm.tile_cols, torch.float32)


def _run(case):
    lin = torch.nn.Linear(case["in_f"], case["out_f"], bias=False)
    lin.weight.data.copy_(case["w"])
    gq = GPTQQuantizer(lin, bits=case["bits"], groupsize=case["groupsize"], actorder=case["actorder"], blocksize=case["blocksize"])
    for x in case["xs"]:
        gq.collect_input_stats(lin, (x,), None)
    trace = float(torch.diag(gq.H).sum())
    qm, err = gq.quantize()
    return qm, err, trace


def test_gptq_reproduces_the_reference_integers():
    cases = load_golden("gptq_cases.pt")
    assert len(cases) == 4
    for c in cases:
        qm, err, trace = _run(c)
        assert isinstance(qm, P.ColBlockQuantizedLinear) and qm.bits == c["bits"]
        assert trace == pytest.approx(c["h_trace"], rel=1e-6)
        assert torch.equal(qm.scales, c["scales"]) and torch.equal(qm.zeros, c["zeros"])
        assert qm.quant_weight.stride() == c["quant_weight"].stride()
        assert torch.equal(qm.quant_weight, c["quant_weight"])          # every packed level identical
        assert err == pytest.approx(c["error"], rel=1e-5, abs=1e-9)


def test_gptq_improves_on_round_to_nearest():
    """The point of the algorithm: with correlated inputs the layer OUTPUT error is smaller than plain rounding."""
    g = torch.Generator().manual_seed(3)
    out_f, in_f = 32, 128
    w = torch.randn(out_f, in_f, generator=g) * 0.05
    mix = torch.randn(in_f, in_f, generator=g) * 0.3 + torch.eye(in_f)
    x = torch.randn(4, 64, in_f, generator=g) @ mix                       # correlated features
    lin = torch.nn.Linear(in_f, out_f, bias=False)
    lin.weight.data.copy_(w)
    gq = GPTQQuantizer(lin, bits=4, groupsize=-1, actorder=True)
    gq.collect_input_stats(lin, (x,), None)
    qm, _ = gq.quantize()
    scale, zero = affine_grid(w, 15)
    rtn = snap_to_grid(w, scale, zero, 15)
    ref = x @ w.t()
    e_gptq = float((x @ _dense(qm).t() - ref).norm())
    e_rtn = float((x @ rtn.t() - ref).norm())
    assert e_gptq < 0.9 * e_rtn, (e_gptq, e_rtn)


def test_gptq_grouped_grids_work():
    """groupsize > 0 raises inside the reference (quantization.py:578); here it produces one grid per group."""
    g = torch.Generator().manual_seed(4)
    lin = torch.nn.Linear(128, 16, bias=False)
    lin.weight.data.copy_(torch.randn(16, 128, generator=g) * 0.05)
    gq = GPTQQuantizer(lin, bits=4, groupsize=32, actorder=False)
    gq.collect_input_stats(lin, (torch.randn(2, 40, 128, generator=g),), None)
    qm, err = gq.quantize()
    assert qm.scales.shape == (16, 4) and qm.tile_cols == 32 and err >= 0
    dense = _dense(qm)
    assert float((dense - lin.weight).abs().max()) < float(qm.scales.max()) * 4   # stays within a few grid steps after error feedback


@pytest.mark.skipif(not os.path.isdir("/root/reference/lit_llama"), reason="needs the reference checkout (build container only)")
def test_gptq_matches_live_reference_on_random_layers():
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "oracle", "_shim"))
    sys.path.insert(0, "/root/reference")
    from lit_llama.quantization import GPTQQuantizer as RefQ

    for seed, (out_f, in_f, bits, act) in enumerate([(12, 96, 4, True), (20, 144, 8, False), (7, 64, 4, False)]):
        g = torch.Generator().manual_seed(50 + seed)
        w = torch.randn(out_f, in_f, generator=g) * 0.1
        xs = [torch.randn(2, 11, in_f, generator=g) for _ in range(2)]
        outs = []
        for cls in (RefQ, GPTQQuantizer):
            lin = torch.nn.Linear(in_f, out_f, bias=False)
            lin.weight.data.copy_(w)
            q = cls(lin, bits=bits, groupsize=-1, actorder=act, blocksize=64)
            for x in xs:
                q.collect_input_stats(lin, (x,), None)
            qm, err = q.quantize()
            outs.append((qm.quant_weight.clone(), qm.scales.clone(), qm.zeros.clone(), err))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
        assert outs[0][3] == pytest.approx(outs[1][3], rel=1e-6)
