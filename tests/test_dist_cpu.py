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
