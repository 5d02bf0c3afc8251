"""bench.py - LLaMA-7B gptq.int4 batch-1 decode throughput on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one decoded token (one pass of generate()'s loop body, generate.py:63-89:
model forward + top-k/softmax/multinomial sampling) on random-init 7B gptq.int4
weights, KV cache S = 2048.  Prints ONE JSON line (rank 0).  See DESIGN.md section
"Measurement" for what every field means.

  value     tokens/s, device-timed (CUDA events), inputs resident in HBM, no host sync
  e2e       same loop driven from HOST buffers: per step a pinned H2D copy of the token
            and position, and a D2H read of the sampled token
  roofline  the tcgen05 int4 linear kernel: algorithmic bytes of all its launches in one
            token / their summed duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the oracle port of the reference CPU path
            (dense dequant + F.linear per call, quantization.py:392-423) on host cores
N > 1: independent replicas, one process per GPU (the reference has no multi-GPU
inference); weak scaling, no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "7B"
S_CTX = 2048
PROMPT_T = 16
TOP_K, TEMPERATURE = 200, 0.8  # generate.py:99-100 defaults


def model_bytes(cfg_name: str):
    """Algorithmic bytes of SURVEY.md section 8d for batch 1: W and KV bytes per position."""
    from oracle.llama_oracle import CONFIGS, n_hidden_for

    c = CONFIGS[cfg_name]
    C, L = c["n_embd"], c["n_layer"]
    nh, V = n_hidden_for(C), 32000
    lin_params = L * (3 * C * C + C * C + 3 * C * nh) + V * C
    lin_rows = L * (3 * C + C + 2 * nh + C) + V
    W = lin_params // 2 + lin_rows * 2 * 2 + (2 * L + 1) * C * 2 + C * 2  # packed + scales/zeros bf16 + norms + 1 wte row
    kv_per_pos = 2 * L * C * 2
    return W, kv_per_pos


def build_synthetic_model(name, dev, seed=1234):
    """Random-init gptq.int4 model of the named size, built directly on the GPU with the
    direct synthesis of SURVEY.md section 8d (uniform levels, zero 8, per-row scales)."""
    import torch

    import lit_llama_b200 as P
    from lit_llama_b200.utils import quantization
    from lit_llama_b200.quantization import ColBlockQuantizedLinear

    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev), quantization("gptq.int4"):
            model = P.LLaMA.from_name(name)
    finally:
        torch.set_default_dtype(prev)
    g = torch.Generator(device=dev).manual_seed(seed)
    std = 0.02 / (2 * model.config.n_layer) ** 0.5
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, ColBlockQuantizedLinear):
                m.quant_weight.random_(0, 256, generator=g)
                m.zeros.fill_(8.0)
                m.scales.copy_(((0.75 + 0.5 * torch.rand(m.scales.shape, device=dev, generator=g)) * (std / 4.61)).to(m.scales.dtype))
            elif isinstance(m, P.RMSNorm):
                m.scale.fill_(1.0)
        model.transformer.wte.weight.normal_(0.0, 0.02, generator=g)
    return model.eval()


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.15 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def sample_next(logits, top_k=TOP_K, temperature=TEMPERATURE):
    """generate.py:68-76 as lit_llama_b200.generate() runs it on the GPU (fused temperature /
    top-k / softmax / draw kernel on torch's Exp(1) noise = torch.multinomial's sample); the CPU baseline keeps the reference's torch ops."""
    import torch

    if logits.is_cuda:
        from lit_llama_b200 import sample_token

        return sample_token(logits[0, -1], temperature, top_k)
    logits = logits[0, -1] / temperature
    v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
    logits = torch.where(logits < v[[-1]], -float("Inf"), logits)
    probs = torch.nn.functional.softmax(logits, dim=-1)
    return torch.multinomial(probs, num_samples=1)


def cpu_baseline(n_blocks=4, threads=None):
    """The reference's CPU path (oracle port) on 7B shapes: one decoded token through
    `n_blocks` Blocks + lm_head, extrapolated to n_layer Blocks."""
    import torch

    from oracle import llama_oracle as O

    if threads is None:
        # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1, which would
        # silently turn the baseline into a single-core run)
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    c = O.CONFIGS[MODEL]
    C, nhd, L = c["n_embd"], c["n_head"], c["n_layer"]
    nh = O.n_hidden_for(C)
    g = torch.Generator().manual_seed(0)

    def lin(out_f, in_f):
        qw = torch.randint(0, 256, (in_f // 2, out_f), dtype=torch.uint8, generator=g).t()
        return O.QLin("gptq", qw=qw, scales=torch.full((out_f, 1), 0.0005, dtype=torch.bfloat16),
                      zeros=torch.full((out_f, 1), 8.0, dtype=torch.bfloat16), bits=4, tile_cols=in_f)

    layers = [dict(rms_1=torch.ones(C, dtype=torch.bfloat16), rms_2=torch.ones(C, dtype=torch.bfloat16), c_attn=lin(3 * C, C),
                   c_proj=lin(C, C), c_fc1=lin(nh, C), c_fc2=lin(nh, C), mlp_proj=lin(C, nh)) for _ in range(n_blocks)]
    m = O.OracleLLaMA(n_layer=n_blocks, n_head=nhd, n_embd=C, block_size=S_CTX, padded_vocab_size=32000,
                      wte=(torch.randn(32000, C, generator=g) * 0.02).bfloat16(), lm_head=lin(32000, C),
                      ln_f=torch.ones(C, dtype=torch.bfloat16), layers=layers)
    with torch.no_grad():
        m.forward(torch.randint(0, 32000, (1, PROMPT_T), generator=g), S_CTX, torch.arange(PROMPT_T))  # prefill, untimed
        tok = torch.randint(0, 32000, (1, 1), generator=g)
        # time the pieces separately so the extrapolation to n_layer Blocks is exact
        x = m.wte[tok]
        t0 = time.perf_counter()
        rope = m.rope.index_select(0, torch.tensor([PROMPT_T]))
        mask = torch.ones(1, 1, 1, S_CTX, dtype=torch.bool)
        mask[..., PROMPT_T + 1:] = False
        for li, lay in enumerate(m.layers):
            x = x + m._attn(O.rmsnorm(x, lay["rms_1"]), lay, rope, mask, S_CTX, torch.tensor([PROMPT_T]), li)
            h = O.rmsnorm(x, lay["rms_2"])
            x = x + lay["mlp_proj"](torch.nn.functional.silu(lay["c_fc1"](h)) * lay["c_fc2"](h))
        t_blocks = time.perf_counter() - t0
        t0 = time.perf_counter()
        logits = m.lm_head(O.rmsnorm(x, m.ln_f))
        sample_next(logits)
        t_head = time.perf_counter() - t0
    t_token = t_blocks / n_blocks * L + t_head
    return {"value": 1.0 / t_token, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 decoded token at pos {PROMPT_T}: {n_blocks} of {L} Blocks timed ({t_blocks:.2f} s) and scaled x{L}/{n_blocks}, "
                      f"+ ln_f/lm_head/sampling ({t_head:.2f} s); oracle port of the reference CPU path (dense dequant + F.linear per call)",
            "s_per_token": t_token}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for i in range(max(1, args.warmup > 0) + max(1, min(args.steps, 2))):
        cb = cpu_baseline(n_blocks=2)
        vals.append(cb)
        if time.perf_counter() - t0 > 150:
            break
    cb = vals[-1]
    line = {"impl": "reference", "metric": "LLaMA-7B gptq.int4 decode tokens/sec", "value": cb["value"], "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["s_per_token"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "LLaMA-7B gptq.int4 decode batch=1 ctx=2048 (random-init weights)", "where": "host CPU"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_q4_launches(model, dev):
    """Every launch of the batch-1 int4 linear kernel (q4_gemv_kernel) of one token, back to back (all
    layers' distinct weights: 3.3 GB, far beyond L2), timed with CUDA events."""
    import ctypes as C

    import torch

    from lit_llama_b200 import _lib as L

    st = model._decode
    a = st.args
    calls = []

    gemv = bool(a.lm_head.qw_mma)

    def mk(w, x, ldx, y, ldy, pro, ns, epi, res):
        return L.Q4LinearArgs(x=x, ldx=ldx, qw_tiled=w.qw_mma if gemv else w.qw_tiled, scales=w.scales, zeros=w.zeros, sz_dtype=a.sz_dtype, y=y, ldy=ldy,
                              M=1, N=w.N, K=w.K, prologue=pro, norm_scale=ns, eps=a.eps, epilogue=epi, res=res, ldres=ldy,
                              split_k=0, flags=1)  # PDL, as b2l_decode_step launches them

    Cd = a.n_embd
    for i in range(a.n_layer):
        ly = st.layers[i]
        calls.append(mk(ly.c_attn, a.x, Cd, a.qkv, 3 * Cd, 1, ly.rms_1, 0, None))
        calls.append(mk(ly.c_proj, a.att, Cd, a.x, Cd, 0, None, 1, a.x))
        calls.append(mk(ly.c_fc12, a.x, Cd, a.hid, a.n_hidden, 1, ly.rms_2, 2, None))
        calls.append(mk(ly.mlp_proj, a.hid, a.n_hidden, a.x, Cd, 0, None, 1, a.x))
    calls.append(mk(a.lm_head, a.x, Cd, a.logits, a.vocab, 1, a.ln_f, 0, None))
    lib, sp = L.lib(), L.stream_ptr()

    fn = lib.b2l_q4_gemv if gemv else lib.b2l_q4_linear_tc

    def run():
        for c in calls:
            rc = fn(C.byref(c), sp)
            if rc:
                raise RuntimeError(lib.b2l_last_error().decode())

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, len(calls), ("q4_gemv_kernel" if gemv else "q4_linear_tc_kernel")


def reduce_max(times, device):
    """Max over ranks of per-rank times (the N > 1 rule of the bench contract).  Replicas
    share nothing else: there is no data-path collective."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(times, device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_throughput(world, steps, t_max):
    """Whole-job tokens/s of `world` replicas that each decoded `steps` tokens."""
    return world * steps / t_max


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(3, args.warmup)
    K = args.steps

    model = build_synthetic_model(MODEL, dev, seed=1234 + rank)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    prompt = torch.randint(0, 32000, (PROMPT_T,), device=dev, dtype=torch.int32, generator=gen)
    lo, span = PROMPT_T, S_CTX - PROMPT_T  # decode positions cycle through [16, 2047]
    pos_all = [torch.tensor([lo + (i % span)], device=dev) for i in range(warm + K)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ---- setup: prefill + graph capture (untimed)
        logits = model(prompt.view(1, -1), S_CTX, torch.arange(PROMPT_T, device=dev))
        tok = sample_next(logits).to(torch.int32)

        # ---- value: device-resident loop, no host sync inside
        for i in range(warm):
            tok = sample_next(model(tok.view(1, 1), S_CTX, pos_all[i])).to(torch.int32)
        barrier()
        clocks = ClockSampler(local)
        clocks.start()
        time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.time()
        e0.record()
        for i in range(warm, warm + K):
            tok = sample_next(model(tok.view(1, 1), S_CTX, pos_all[i])).to(torch.int32)
        e1.record()
        barrier()
        tw1 = time.time()
        t_dev = e0.elapsed_time(e1) * 1e-3
        clk = clocks.stop(tw0, tw1)

        # ---- e2e: host buffers; per step H2D (token, position) from pinned memory, D2H sampled token
        h_tok = torch.empty(1, dtype=torch.int32).pin_memory()
        h_pos = torch.empty(1, dtype=torch.int64).pin_memory()
        h_out = torch.empty(1, dtype=torch.int64).pin_memory()
        d_tok = torch.empty(1, dtype=torch.int32, device=dev)
        d_pos = torch.empty(1, dtype=torch.int64, device=dev)
        h_tok[0] = int(tok)
        Ke = min(K, 512)
        for phase in ("warm", "timed"):
            n = 8 if phase == "warm" else Ke
            barrier()
            t0 = time.perf_counter()
            for i in range(n):
                h_pos[0] = lo + (i % span)
                d_tok.copy_(h_tok, non_blocking=True)
                d_pos.copy_(h_pos, non_blocking=True)
                nxt = sample_next(model(d_tok.view(1, 1), S_CTX, d_pos))
                h_out.copy_(nxt, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                h_tok[0] = int(h_out[0])
            barrier()
            t_e2e = time.perf_counter() - t0

        t_q4, n_q4, q4_name = time_q4_launches(model, dev)

    t_dev, t_e2e = reduce_max([t_dev, t_e2e], dev)

    if rank == 0:
        W, kv = model_bytes(MODEL)
        mean_p = sum(lo + (i % span) for i in range(warm, warm + K)) / K
        bytes_per_token = W + kv * (mean_p + 1) + kv
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ach = W / t_q4 / 1e9
        traffic = None
        try:  # dram bytes of the kernel's launches of one token, from the committed ncu capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["bytes_per_token"]
        except (OSError, KeyError, ValueError):
            pass
        from lit_llama_b200 import _lib as L
        import ctypes as C
        launches = L.lib().b2l_decode_step_launches(C.byref(model._decode.args))
        line = {
            "metric": "LLaMA-7B gptq.int4 decode tokens/sec", "value": aggregate_throughput(world, K, t_dev), "unit": "tokens/s", "n_gpus": world,
            "steps": K, "warmup": warm, "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "LLaMA-7B gptq.int4 decode batch=1 ctx=2048 (random-init weights)", "prompt_tokens": PROMPT_T,
                       "positions": f"cycle {lo}..{S_CTX - 1}, mean {mean_p:.0f}", "sampling": f"top_k={TOP_K} temperature={TEMPERATURE}",
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "l2": "weights 3.31 GB per token >> 126 MB L2 (inputs larger than L2)"},
            "clocks": clk,
            "e2e": {"value": aggregate_throughput(world, Ke, t_e2e), "unit": "tokens/s", "h2d_bytes_per_step": 12, "d2h_bytes_per_step": 8, "steps": Ke},
            "gpu_launches": (launches + 1) * K,  # b2l_decode_step's kernels + the fused sampling kernel, per token
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "kernel": q4_name, "launches_per_token": n_q4, "bytes_per_token_launches": W,
                         "peak_source": which,
                         "whole_token": {"bytes": bytes_per_token, "achieved": bytes_per_token * K / t_dev / 1e9,
                                         "frac": bytes_per_token * K / t_dev / 1e9 / peak}},
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(n_blocks=2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
