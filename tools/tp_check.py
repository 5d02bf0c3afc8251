"""Tensor-parallel check on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py

1. tiny gptq.int4 model: TPLLaMA logits (prefill + decode) vs the single-GPU LLaMA on rank 0's GPU;
2. LLaMA-7B shapes (random-init): decode tokens/s of the TP path, eager and CUDA-graph replayed.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lit_llama_b200 as P  # noqa: E402
from lit_llama_b200.utils import quantization  # noqa: E402
from oracle import llama_oracle as O  # noqa: E402


def build_tp(cfg, sd, rank, world, dev):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            m = P.TPLLaMA(P.LLaMAConfig(**cfg), rank, world, O.n_hidden_for(cfg["n_embd"]))
    finally:
        torch.set_default_dtype(prev)
    m.load_state_dict(P.shard_state_dict(sd, rank, world, cfg["n_head"]))
    return m.eval()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---- 1. parity on a tiny model
    cfg = dict(block_size=64, vocab_size=96, n_layer=2, n_head=4, n_embd=128)
    sd = O.synth_state_dict(2, 4, 128, 96, "gptq.int4", dtype=torch.bfloat16, seed=1234)
    tp = build_tp(cfg, sd, rank, world, dev)
    prompt = torch.tensor([[3, 17, 40, 41, 2, 77, 5]], device=dev)
    got = [tp(prompt, 16, torch.arange(7, device=dev))]
    for i, t in enumerate([11, 5, 90]):
        got.append(tp(torch.tensor([[t]], device=dev), 16, torch.tensor([7 + i], device=dev)))
    if rank == 0:
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        with torch.device(dev), quantization("gptq.int4"):
            full = P.LLaMA(P.LLaMAConfig(**cfg))
        torch.set_default_dtype(prev)
        full.load_state_dict(sd)
        with torch.no_grad():
            want = [full(prompt, 16, torch.arange(7, device=dev))]
            for i, t in enumerate([11, 5, 90]):
                want.append(full(torch.tensor([[t]], device=dev), 16, torch.tensor([7 + i], device=dev)).clone())
        for j, (a, b) in enumerate(zip(got, want)):
            err = float((a.float() - b.float()).abs().max())
            rel = float((a.float() - b.float()).norm() / b.float().norm())
            print(f"[tp={world}] tiny step {j}: max_abs={err:.4f} relerr={rel:.3e} {'OK' if rel < 1e-2 else 'MISMATCH'}", flush=True)
    dist.barrier()

    # ---- 2. 7B shapes: decode speed
    from bench import S_CTX
    c7 = dict(block_size=2048, vocab_size=32000, n_layer=32, n_head=32, n_embd=4096)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        m = P.TPLLaMA(P.LLaMAConfig(**c7), rank, world, O.n_hidden_for(4096))
    torch.set_default_dtype(prev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, P.ColBlockQuantizedLinear):
                mod.quant_weight.random_(0, 256, generator=g)
                mod.zeros.fill_(8.0)
                mod.scales.fill_(0.0005)
            elif isinstance(mod, P.RMSNorm):
                mod.scale.fill_(1.0)
        m.transformer.wte.weight.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(1))
    m.eval()
    tok = torch.randint(0, 32000, (1, 1), device=dev, dtype=torch.int32)
    dist.broadcast(tok, 0)
    m(torch.randint(0, 32000, (1, 16), device=dev, dtype=torch.int32), S_CTX, torch.arange(16, device=dev))
    pos = [torch.tensor([16 + i], device=dev) for i in range(40)]
    for i in range(4):
        m(tok, S_CTX, pos[i])
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(4, 36):
        m(tok, S_CTX, pos[i])
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / 32
    if rank == 0:
        print(f"[tp={world}] 7B decode, eager module path: {dt * 1e6:.0f} us/token  {1 / dt:.1f} tok/s", flush=True)
    # CUDA graph of one step (NCCL collectives captured) - opt-in: B2L_TP_GRAPH=1
    if os.environ.get("B2L_TP_GRAPH") != "1":
        dist.destroy_process_group()
        return
    try:
        st_tok, st_pos = tok.clone(), pos[36].clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                m(st_tok, S_CTX, st_pos)
        torch.cuda.current_stream().wait_stream(s)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = m(st_tok, S_CTX, st_pos)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(64):
            gr.replay()
        torch.cuda.synchronize(); dist.barrier()
        dt = (time.perf_counter() - t0) / 64
        if rank == 0:
            print(f"[tp={world}] 7B decode, CUDA graph: {dt * 1e6:.0f} us/token  {1 / dt:.1f} tok/s", flush=True)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(f"[tp={world}] graph capture of the TP step failed: {type(e).__name__}: {e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
