"""Pins oracle/llama_oracle.py against vectors produced by the unmodified reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import torch

from conftest import load_golden
from oracle import llama_oracle as O


def test_find_multiple():
    for n, k, want in load_golden("ops.pt")["find_multiple"]:
        assert O.find_multiple(n, k) == want


def test_pack_unpack_dequant_bit_exact():
    for c in load_golden("quant_cases.pt"):
        bits, gs = c["bits"], c["groupsize"]
        in_f = c["w"].shape[1]
        tc = in_f if gs == -1 else gs
        # RTN parameters and levels follow quantization.py:471-513
        for g in range(c["scales"].shape[1]):
            sl = slice(g * tc, (g + 1) * tc)
            s, z = O.rtn_params(c["w"][:, sl], bits)
            assert torch.equal(s, c["scales"][:, g : g + 1]) and torch.equal(z, c["zeros"][:, g : g + 1])
        lv = torch.cat([O.rtn_levels(c["w"][:, g * tc : (g + 1) * tc], c["scales"][:, g : g + 1], c["zeros"][:, g : g + 1], bits)
                        for g in range(c["scales"].shape[1])], dim=1)
        # the reference reconstructs scale*(q-zero) (quantize_weight) and pack_weight divides
        # it back and TRUNCATES, so a level can drop by one through fp error: follow it exactly
        rec = torch.cat([c["scales"][:, g : g + 1] * (lv[:, g * tc : (g + 1) * tc] - c["zeros"][:, g : g + 1])
                         for g in range(c["scales"].shape[1])], dim=1)
        packed = O.pack_weight(rec, c["scales"], c["zeros"], bits, tc)
        assert packed.dtype == torch.uint8 and tuple(packed.stride()) == c["qw_stride"] == (1, packed.shape[0])
        assert torch.equal(packed, c["quant_weight"])
        got_lv = O.unpack_levels(c["quant_weight"], bits).float()
        assert (got_lv - lv).abs().max() <= 1
        assert torch.equal(O.unpack_levels(O.pack_levels(lv, bits), bits).float(), lv)
        assert torch.equal(O.dequant(c["quant_weight"], c["scales"], c["zeros"], bits, tc, torch.float32), c["deq_f32"])
        assert torch.equal(O.dequant(c["quant_weight"], c["scales"], c["zeros"], bits, tc, torch.bfloat16), c["deq_bf16"])


def test_qlinear_matches_reference_forward():
    for c in load_golden("quant_cases.pt"):
        bits, gs = c["bits"], c["groupsize"]
        tc = c["w"].shape[1] if gs == -1 else gs
        y = O.qlinear(c["x"], c["quant_weight"], c["scales"], c["zeros"], bits, tc)
        assert torch.equal(y, c["y_f32"])
        yb = O.qlinear(c["x"].bfloat16(), c["quant_weight"], c["scales"].bfloat16(), c["zeros"].bfloat16(), bits, tc)
        assert torch.equal(yb, c["y_bf16"])
        # the exact-arithmetic target agrees with the fp32 reference forward to fp32 noise
        ye = O.qlinear_exact(c["x"], c["quant_weight"], c["scales"], c["zeros"], bits, tc)
        torch.testing.assert_close(ye, c["y_f32"], rtol=1e-5, atol=1e-6)


def test_rmsnorm_rope():
    g = load_golden("ops.pt")
    assert torch.equal(O.rmsnorm(g["rms_x"], g["rms_scale"]), g["rms_y_f32"])
    assert torch.equal(O.rmsnorm(g["rms_x"].bfloat16(), g["rms_scale"].bfloat16()), g["rms_y_bf16"])
    assert torch.equal(O.rope_table(64, 32), g["rope_table_64x32"])
    assert torch.equal(O.rope_table(2048, 128)[[0, 1, 777, 2047]], g["rope_table_2048x128_rows"])
    assert torch.equal(O.rope_apply(g["rope_x"], g["rope_table_64x32"]), g["rope_y_f32"])
    assert torch.equal(O.rope_apply(g["rope_x"].bfloat16(), g["rope_table_64x32"]), g["rope_y_bf16"])


def _oracle_model(gd, dtype, mode="gptq.int4"):
    cfg = gd["cfg"]
    sd = O.synth_state_dict(cfg["n_layer"], cfg["n_head"], cfg["n_embd"], cfg["vocab_size"], mode, dtype=dtype, seed=gd["seed"])
    return O.OracleLLaMA.from_state_dict(sd, cfg["n_layer"], cfg["n_head"], cfg["block_size"], mode)


def _check_model(tag, dtype, rtol, atol):
    gd = load_golden(f"tiny_int4_{tag}.pt")
    m = _oracle_model(gd, dtype)
    prompt = gd["prompt"]
    S = 16
    got = [m.forward(prompt.view(1, -1), S, torch.arange(7))]
    for i, t in enumerate(gd["steps_tokens"]):
        got.append(m.forward(torch.tensor([[t]]), S, torch.tensor([7 + i])))
    for a, b in zip(got, gd["steps_logits"]):
        torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)
    torch.testing.assert_close(m.kv[0][0].float(), gd["kv0_k"].float(), rtol=rtol, atol=atol)
    torch.testing.assert_close(m.kv[0][1].float(), gd["kv0_v"].float(), rtol=rtol, atol=atol)
    m.reset_cache()
    torch.testing.assert_close(m.forward(prompt.view(1, -1)).float(), gd["nocache_logits"].float(), rtol=rtol, atol=atol)
    # roll branch
    m.reset_cache()
    S2 = 8
    got = [m.forward(prompt.view(1, -1), S2, torch.arange(7))[:, -1]]
    for i, t in enumerate(gd["roll_tokens"]):
        got.append(m.forward(torch.tensor([[t]]), S2, torch.tensor([7 + i]))[:, -1])
    for a, b in zip(got, gd["roll_logits"]):
        torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)
    torch.testing.assert_close(m.kv[1][0].float(), gd["roll_kv1_k"].float(), rtol=rtol, atol=atol)
    return gd, m


def test_tiny_model_fp32():
    gd, m = _check_model("f32", torch.float32, 1e-4, 1e-5)
    m.reset_cache()
    assert torch.equal(O.generate(m, gd["prompt"].to(torch.int32), 12, top_k=1), gd["gen_greedy"])
    m.reset_cache()
    torch.manual_seed(1234)
    assert torch.equal(O.generate(m, gd["prompt"].to(torch.int32), 12, temperature=0.8, top_k=20), gd["gen_sampled"])
    m.reset_cache()
    torch.manual_seed(99)
    assert torch.equal(O.generate(m, gd["prompt"].to(torch.int32), 12, max_seq_length=10, top_k=4), gd["gen_roll"])


def test_tiny_model_bf16():
    # the reference's own bf16 tolerance (tests/test_model.py:133)
    gd, m = _check_model("bf16", torch.bfloat16, 1e-3, 5e-3)
    m.reset_cache()
    assert torch.equal(O.generate(m, gd["prompt"].to(torch.int32), 12, top_k=1), gd["gen_greedy"])


def test_dense_generate_roll():
    """tests/test_generate.py:26-54 shape: unquantized fp32, head_size 2, roll branch."""
    gd = load_golden("tiny_dense_f32.pt")
    cfg = gd["cfg"]
    sd = O.synth_state_dict(1, 4, 8, 16, None, dtype=torch.float32, seed=gd["seed"])
    m = O.OracleLLaMA.from_state_dict(sd, 1, 4, cfg["block_size"], None)
    torch.manual_seed(4)
    y = O.generate(m, gd["prompt"], 20, max_seq_length=10, top_k=4)
    assert torch.equal(y, gd["gen"])


def test_int8_restatement_self_consistency():
    """llm.int8 is parity-unpinned (no bitsandbytes here); check the restatement's
    own invariants: no-outlier path is close to the fp product, outlier columns are
    routed through the fp16 branch."""
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 64, generator=g) * 0.05
    x = torch.randn(4, 64, generator=g)
    cb, scb = O.int8_quantize_weight(w)
    assert cb.dtype == torch.int8 and scb.shape == (32,)
    y = O.int8_linear(x, cb, scb)
    ref = x @ w.t()
    assert (y - ref).norm() / ref.norm() < 2e-2
    x2 = x.clone()
    x2[1, 5] = 9.0
    y2 = O.int8_linear(x2, cb, scb)
    ref2 = x2 @ w.t()
    assert (y2 - ref2).norm() / ref2.norm() < 2e-2
