"""-m gpu: model.py pieces, the whole tiny model, generate() and the 7B-shaped Block
against the oracle and the golden vectors produced by the unmodified reference."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from oracle import llama_oracle as O  # noqa: E402

CFG = dict(block_size=64, vocab_size=96, n_layer=2, n_head=4, n_embd=128)
RTOL, ATOL = 1e-3, 5e-3  # the reference's bf16 tolerance, tests/test_model.py:133


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def test_rmsnorm_and_rope_match_reference(dev):
    import lit_llama_b200 as P

    g = load_golden("ops.pt")
    n = P.RMSNorm(128).to(dev).bfloat16()
    n.scale.data = g["rms_scale"].bfloat16().to(dev)
    y = n(g["rms_x"].bfloat16().to(dev)).cpu()
    assert float((y == g["rms_y_bf16"]).float().mean()) > 0.995  # same rounding points; summation order may flip an ulp
    torch.testing.assert_close(y.float(), g["rms_y_bf16"].float(), rtol=2 ** -7, atol=1e-6)
    yr = P.apply_rope(g["rope_x"].bfloat16().to(dev), g["rope_table_64x32"].to(dev)).cpu()
    assert torch.equal(yr, g["rope_y_bf16"])  # fp32 products and sums in the reference's order: bit-exact
    tab = P.build_rope_cache(64, 32, torch.int64, dev)
    torch.testing.assert_close(tab.cpu(), g["rope_table_64x32"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("graph_after", [0, 2])
def test_tiny_model_matches_reference(dev, graph_after):
    from gpu_util import build_tiny

    gd = load_golden("tiny_int4_bf16.pt")
    model, _, _ = build_tiny(dev, CFG)
    model.graph_after = graph_after
    S = 16
    with torch.no_grad():
        got = [model(gd["prompt"].view(1, -1).to(dev), S, torch.arange(7, device=dev))]
        for i, t in enumerate(gd["steps_tokens"]):
            got.append(model(torch.tensor([[t]], device=dev), S, torch.tensor([7 + i], device=dev)))
    for a, b in zip(got, gd["steps_logits"]):
        torch.testing.assert_close(a.float().cpu(), b.float(), rtol=RTOL, atol=ATOL)
    k, v = model.kv_caches[0]
    torch.testing.assert_close(k.float().cpu(), gd["kv0_k"].float(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(v.float().cpu(), gd["kv0_v"].float(), rtol=RTOL, atol=ATOL)
    model.reset_cache()
    with torch.no_grad():
        lg = model(gd["prompt"].view(1, -1).to(dev))
    torch.testing.assert_close(lg.float().cpu(), gd["nocache_logits"].float(), rtol=RTOL, atol=ATOL)


def test_roll_branch_matches_reference(dev):
    from gpu_util import build_tiny

    gd = load_golden("tiny_int4_bf16.pt")
    model, _, _ = build_tiny(dev, CFG)
    S2 = 8
    with torch.no_grad():
        got = [model(gd["prompt"].view(1, -1).to(dev), S2, torch.arange(7, device=dev))[:, -1]]
        for i, t in enumerate(gd["roll_tokens"]):
            got.append(model(torch.tensor([[t]], device=dev), S2, torch.tensor([7 + i], device=dev))[:, -1])
    for a, b in zip(got, gd["roll_logits"]):
        torch.testing.assert_close(a.float().cpu(), b.float(), rtol=RTOL, atol=ATOL)
    kl = model.logical_kv_caches()[1][0]
    torch.testing.assert_close(kl.float().cpu(), gd["roll_kv1_k"].float(), rtol=RTOL, atol=ATOL)


def test_generate_matches_reference_tokens(dev):
    import lit_llama_b200 as P
    from gpu_util import build_tiny

    gd = load_golden("tiny_int4_bf16.pt")
    model, oracle, _ = build_tiny(dev, CFG)
    prompt = gd["prompt"].to(torch.int32).to(dev)
    y = P.generate(model, prompt, 12, top_k=1)
    assert y.shape == gd["gen_greedy"].shape and y.dtype == torch.int32
    assert torch.equal(y[:7].cpu(), gd["prompt"].to(torch.int32))
    # greedy tokens: equal to the reference's unless two logits tie within bf16 noise
    same = (y.cpu() == gd["gen_greedy"]).float().mean()
    assert same >= 0.9, (y.cpu().tolist(), gd["gen_greedy"].tolist())
    # loop semantics of tests/test_generate.py:26-54: length, order of the sampled tokens, roll branch
    model.reset_cache()
    from unittest import mock

    draws = []
    orig = torch.multinomial

    def spy(*a, **k):
        out = orig(*a, **k)
        draws.append(out)
        return out

    with mock.patch("torch.multinomial", spy):
        out = P.generate(model, prompt, 20, max_seq_length=10, top_k=4)
    assert out.size(0) == 7 + 20
    assert torch.equal(out.cpu(), torch.cat((prompt.cpu(), torch.hstack(draws).cpu().to(torch.int32))))
    # eos stops and includes the eos token (generate.py:88-89)
    model.reset_cache()
    first = int(P.generate(model, prompt, 1, top_k=1)[-1])
    model.reset_cache()
    out = P.generate(model, prompt, 5, top_k=1, eos_id=first)
    assert out.tolist() == prompt.tolist() + [first]


def test_fast_path_equals_module_path(dev):
    """b2l_decode_step (fused, graph) vs the module-by-module path on the same step."""
    from gpu_util import build_tiny

    model, _, _ = build_tiny(dev, CFG)
    prompt = torch.tensor([[3, 17, 40, 41, 2, 77, 5]], device=dev)
    with torch.no_grad():
        model(prompt, 16, torch.arange(7, device=dev))
        fast = model(torch.tensor([[9]], device=dev), 16, torch.tensor([7], device=dev)).clone()
        model.reset_cache()
        model._fast_ok = False
        model(prompt, 16, torch.arange(7, device=dev))
        slow = model(torch.tensor([[9]], device=dev), 16, torch.tensor([7], device=dev))
    torch.testing.assert_close(fast.float(), slow.float(), rtol=RTOL, atol=ATOL)


def test_7b_shaped_block_vs_oracle(dev):
    """One Block + lm_head at the BASELINE 7B widths (n_embd 4096, 32 heads, n_hidden
    11008, vocab 32000): prefill 5 tokens then 2 decode steps, against the oracle."""
    from gpu_util import build_tiny

    cfg = dict(block_size=32, vocab_size=32000, n_layer=1, n_head=32, n_embd=4096)
    model, oracle, _ = build_tiny(dev, cfg, seed=11)
    prompt = torch.tensor([[5, 100, 31999, 7, 2048]])
    S = 8
    with torch.no_grad():
        got = [model(prompt.to(dev), S, torch.arange(5, device=dev))]
        want = [oracle.forward(prompt, S, torch.arange(5))]
        for i, t in enumerate([77, 12345]):
            got.append(model(torch.tensor([[t]], device=dev), S, torch.tensor([5 + i], device=dev)))
            want.append(oracle.forward(torch.tensor([[t]]), S, torch.tensor([5 + i])))
    for a, b in zip(got, want):
        a, b = a.float().cpu(), b.float()
        assert (a - b).norm() / b.norm() < 2e-2  # the oracle's dense path rounds every weight to bf16
        torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)


def test_batched_decode_rows_are_independent(dev):
    from gpu_util import build_tiny

    model, _, _ = build_tiny(dev, CFG)
    idx = torch.tensor([[3, 17, 40], [9, 9, 1]], device=dev)
    with torch.no_grad():
        model(idx, 16, torch.arange(3, device=dev))
        both = model(torch.tensor([[5], [60]], device=dev), 16, torch.tensor([3], device=dev)).clone()
        model.reset_cache()
        model(idx[1:], 16, torch.arange(3, device=dev))
        one = model(torch.tensor([[60]], device=dev), 16, torch.tensor([3], device=dev))
    torch.testing.assert_close(both[1:].float(), one.float(), rtol=RTOL, atol=ATOL)
