"""BASELINE config 5 strong-scaling point: 400M rows, 10 regressors | unit(4M) + time(500), IID, on N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        tests/scale_c5_multigpu.py [--rows-total 400000000] [--steps 3]

Every rank generates its shard on the device (units [r*G/N, (r+1)*G/N), rows_total/N rows), the fit is timed
device-resident with a barrier + synchronize on both sides, max over ranks.  Rank 0 prints one JSON line.
With N = 1 this is the single-GPU reference point of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from leanfe_b200 import leanfe  # noqa: E402
from leanfe_b200.engine import get_engine  # noqa: E402


def shard(n, units_lo, units_n, seed, k=10):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    unit = torch.randint(0, units_n, (n,), generator=g, device="cuda")
    tm = torch.randint(0, 500, (n,), generator=g, device="cuda")
    a = torch.randn(units_n, generator=g, device="cuda", dtype=torch.float64)
    gt = torch.sin(torch.arange(500, device="cuda", dtype=torch.float64))          # same period effects on all ranks
    d = dict(unit=unit + units_lo, time=tm)
    y = a[unit] + gt[tm] + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    for j in range(1, k + 1):
        x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.5 * a[unit]
        y += (1 + 0.1 * j) * x
        d[f"x{j}"] = x
    d["y"] = y
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-total", type=int, default=400_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    eng = get_engine(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        eng.comm_init_from_torch()
    n = a.rows_total // world
    units_total = a.rows_total // 100
    d = shard(n, rank * (units_total // world), units_total // world, 1234 + rank)
    xs = " + ".join(f"x{j}" for j in range(1, 11))
    kw = dict(formula=f"y ~ {xs} | unit + time", vcov="iid", strategy="alt_proj")
    if world > 1:
        kw["local_fe"] = "unit"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(a.warmup):
        res = leanfe(d, **kw)
    eng.set_profiling(True)
    eng.reset_timers()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = leanfe(d, **kw)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    timers = {k_: round(v / a.steps, 3) for k_, v in eng.timers().items() if v > 0}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        print(json.dumps(dict(config="c5", rows_total=a.rows_total, n_gpus=world, ms_per_fit=ms,
                              rows_per_s=a.rows_total / (ms / 1e3), sweeps=res.iterations, converged=res.converged,
                              n_obs=res.n_obs, coefs=res.coefs, phases_ms=timers,
                              mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
