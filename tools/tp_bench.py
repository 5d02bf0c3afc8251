"""Tensor-parallel batch-1 decode throughput on N GPUs of one node (BASELINE.json configs[4]):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/tp_bench.py --model 65B [--steps 64] [--warmup 8]

Every rank holds its Megatron shard of a random-init gptq.int4 model (synthesised shard by shard on its GPU: a full 65B
state dict never exists), decodes the same token stream (identical logits after the all-gather, identical sampler seed)
and replays its fused per-rank step as a CUDA graph with the NCCL all-reduces inside (lit_llama_b200/tp.py).  Timing:
CUDA events around K steps, max over ranks.  Rank 0 prints one JSON line.  Also imported by bench.py (`tp` block)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {"7B": (32, 32, 4096), "13B": (40, 40, 5120), "30B": (60, 52, 6656), "65B": (80, 64, 8192)}  # model.py:43-48


def n_hidden_for(n_embd):
    h = int(2 * 4 * n_embd / 3)
    return h if h % 256 == 0 else h + 256 - h % 256


def build_tp_model(name, rank, world, dev, group=None, seed=1234):
    import torch

    import lit_llama_b200 as P
    from lit_llama_b200.quantization import ColBlockQuantizedLinear

    L_, nh, C = SHAPES[name]
    cfg = P.LLaMAConfig(block_size=2048, vocab_size=32000, n_layer=L_, n_head=nh, n_embd=C)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            m = P.TPLLaMA(cfg, rank, world, n_hidden_for(C), group=group)
    finally:
        torch.set_default_dtype(prev)
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    std = 0.02 / (2 * L_) ** 0.5
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, ColBlockQuantizedLinear):
                mod.quant_weight.random_(0, 256, generator=g)
                mod.zeros.fill_(8.0)
                mod.scales.copy_(((0.75 + 0.5 * torch.rand(mod.scales.shape, device=dev, generator=g)) * (std / 4.61)).to(mod.scales.dtype))
            elif isinstance(mod, P.RMSNorm):
                mod.scale.fill_(1.0)
        gw = torch.Generator(device=dev).manual_seed(seed)   # replicated tensors: the same on every rank
        m.transformer.wte.weight.normal_(0.0, 0.02, generator=gw)
    return m.eval()


def run_tp(name, steps, warmup, dev, rank, world, group=None, pos0=16):
    """Decode `steps` tokens (positions spread evenly over pos0..2047) after `warmup` steps; returns the result dict."""
    import torch
    import torch.distributed as dist

    from lit_llama_b200 import sample_token

    model = build_tp_model(name, rank, world, dev, group)
    S = 2048
    L_, nh, C = SHAPES[name]
    nhid = n_hidden_for(C)
    w_total = (L_ * (4 * C * C + 3 * C * nhid) + 32000 * C) // 2
    w_rank = w_total // world
    kv_rank_per_pos = 2 * L_ * C * 2 // world
    span = S - pos0
    timed = [pos0 + (i * span) // steps for i in range(steps)]
    tok = torch.tensor([[7]], device=dev, dtype=torch.int32)
    torch.manual_seed(99)   # same sampler stream on every rank
    with torch.no_grad():
        for i in range(max(3, warmup)):
            logits = model(tok, S, torch.tensor([pos0 + i], device=dev))
            tok = sample_token(logits[0, -1], 0.8, 200).to(torch.int32).view(1, 1)
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pos_t = [torch.tensor([q], device=dev) for q in timed]
        e0.record()
        for i in range(steps):
            logits = model(tok, S, pos_t[i])
            tok = sample_token(logits[0, -1], 0.8, 200).to(torch.int32).view(1, 1)
        e1.record()
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    t = float(t)
    peak = 6573.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    mean_p = sum(timed) / steps
    bytes_rank = w_rank + kv_rank_per_pos * (mean_p + 2)
    st = model._decode
    model.tp_check()
    ll = st is not None and getattr(st, "comm", None) is not None
    res = {"model": f"LLaMA-{name} gptq.int4", "tp": world, "tokens_per_s": steps / t, "ms_per_token": t / steps * 1e3, "steps": steps,
           "positions": f"{steps} positions spread evenly over {pos0}..{S - 1}", "graph": bool(st is not None and st.graph is not None),
           "allreduces_per_token": 2 * L_, "collective": ("b2l_tp_allreduce: one-shot sum over peer memory (NVLink stores of {2 x bf16, epoch} words, fp32 sum in rank order) inside "
                          "the captured graph" if ll else "NCCL all-reduce (bf16 sum) inside the captured graph") + "; NCCL all-gather of the logits",
           "per_rank_weight_bytes": w_rank, "per_rank_hbm_frac": bytes_rank / (t / steps) / 1e9 / peak}
    del model
    torch.cuda.empty_cache()
    return res


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7B")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_tp(args.model, args.steps, args.warmup, dev, rank, world)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:   # see bench.py: teardown with peer-mapped buffers alive can block; everything is printed
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
