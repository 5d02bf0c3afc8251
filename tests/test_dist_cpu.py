"""CPU, world_size 2, gloo: the N > 1 path of bench.py (independent replicas; the only
cross-rank step is the max-over-ranks of the timed region) and the rank-0-only rule of
the reference arm."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench

    t = bench.reduce_max([1.0 + rank, 5.0 - rank], torch.device("cpu"))
    dist.barrier()
    if rank == 0:
        json.dump({"t": t, "tp": bench.aggregate_throughput(world, 100, t[0])}, open(out, "w"))
    dist.destroy_process_group()


def test_replica_timing_is_max_over_ranks(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["t"] == [2.0, 5.0]
    assert r["tp"] == 2 * 100 / 2.0


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


# ---- tensor-parallel sharding (lit_llama_b200.tp): host logic on CPU, 2 ranks over gloo -------------------------
def _tp_oracle_forward(sd_l, cfg, world, idx, S, pos):
    """The TPLLaMA.forward dataflow restated with oracle arithmetic (fp32) on one rank's shard."""
    from oracle import llama_oracle as O

    C, nh = cfg["n_embd"], cfg["n_head"]
    hs, nh_l = C // nh, nh // world

    def lin(prefix, x):
        return O.qlinear_exact(x, sd_l[prefix + ".quant_weight"], sd_l[prefix + ".scales"], sd_l[prefix + ".zeros"], 4,
                               sd_l[prefix + ".quant_weight"].shape[1] * 2)

    B, T = idx.shape
    rope = O.rope_table(cfg["block_size"], hs).index_select(0, pos)
    mask = torch.tril(torch.ones(cfg["block_size"], cfg["block_size"], dtype=torch.bool)).index_select(0, pos)[:, :S].reshape(1, 1, T, S)
    x = sd_l["transformer.wte.weight"].float()[idx]
    for i in range(cfg["n_layer"]):
        p = f"transformer.h.{i}."
        qkv = lin(p + "attn.c_attn", O.rmsnorm(x, sd_l[p + "rms_1.scale"].float()))
        q, k, v = qkv.split(nh_l * hs, dim=2)
        q = O.rope_apply(q.view(B, T, nh_l, hs), rope).transpose(1, 2)
        k = O.rope_apply(k.view(B, T, nh_l, hs), rope).transpose(1, 2)
        v = v.view(B, T, nh_l, hs).transpose(1, 2)
        kc = torch.zeros(B, nh_l, S, hs).index_copy(2, pos, k)
        vc = torch.zeros(B, nh_l, S, hs).index_copy(2, pos, v)
        y = O.sdpa(q, kc, vc, mask).transpose(1, 2).contiguous().view(B, T, nh_l * hs)
        part = lin(p + "attn.c_proj", y)
        dist.all_reduce(part)
        x = x + part
        h = O.rmsnorm(x, sd_l[p + "rms_2.scale"].float())
        part = lin(p + "mlp.c_proj", torch.nn.functional.silu(lin(p + "mlp.c_fc1", h)) * lin(p + "mlp.c_fc2", h))
        dist.all_reduce(part)
        x = x + part
    logits_l = lin("lm_head", O.rmsnorm(x, sd_l["transformer.ln_f.scale"].float()))
    parts = [torch.empty_like(logits_l) for _ in range(world)]
    dist.all_gather(parts, logits_l)
    return torch.cat(parts, dim=-1)


def _tp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import lit_llama_b200 as P
    from oracle import llama_oracle as O

    cfg = dict(block_size=32, vocab_size=96, n_layer=2, n_head=4, n_embd=128)
    sd = O.synth_state_dict(2, 4, 128, 96, "gptq.int4", dtype=torch.float32, seed=3)
    sd_l = P.shard_state_dict(sd, rank, world, cfg["n_head"])
    # the shard loads into the sharded module tree with the reference's key names
    m = P.TPLLaMA(P.LLaMAConfig(**cfg), rank, world, O.n_hidden_for(128))
    res = m.load_state_dict(sd_l)
    assert not res.missing_keys and not res.unexpected_keys
    assert tuple(m.transformer.h[0].attn.c_attn.quant_weight.stride()) == (1, 3 * 128 // world)
    idx = torch.tensor([[3, 17, 40, 41, 2]])
    got = _tp_oracle_forward(sd_l, cfg, world, idx, 16, torch.arange(5))
    if rank == 0:
        full = O.OracleLLaMA.from_state_dict(sd, 2, 4, 32, "gptq.int4", exact_linears=True)
        want = full.forward(idx, 16, torch.arange(5))
        json.dump({"err": float((got - want).abs().max()), "scale": float(want.abs().max())}, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_tp_sharding_reproduces_the_unsharded_model(tmp_path):
    out = str(tmp_path / "tp.json")
    mp.spawn(_tp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["err"] < 1e-4 * max(1.0, r["scale"]), r
