"""BASELINE configs 3 and 4 at full size on N GPUs (config 5: tests/scale_c5_multigpu.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29545 \
        tests/scale_multigpu.py {c3|c4} [--rows-total R] [--steps 3]

  c3: AKM-style three-way FE, 100M rows | worker(10M) + firm(1M) + year(30), k = 2, HC1, rows sharded by worker
      (each rank owns a contiguous block of workers; firm and year tables are all-reduced every sweep).
  c4: YOCO, 200M rows, 23 treated x period dummies | group(1000) + period(24), cluster(group); rows sharded
      ARBITRARILY (every rank draws its own rows of the same DGP), partial tables merged by lfe_compress_merge.

Every rank generates its shard on the device; fits are device-resident, timed with CUDA events between
barriers, max over ranks.  Rank 0 prints one JSON line with the size-independent checks:
  * truth recovery |beta - truth| / se (max over regressors);
  * c4: the merged group counts sum to rows_total and n_compressed <= 24 000;
  * identical coefficients on all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from leanfe_b200 import leanfe  # noqa: E402
from leanfe_b200.engine import get_engine  # noqa: E402


def gen(seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def shard_c3(n, rank, world, rows_total):
    nw_total, nf = rows_total // 10, rows_total // 100
    nw = nw_total // world
    g = gen(4200 + rank)
    gs = gen(99)                                                       # shared by all ranks: firm effects
    af = torch.randn(nf, generator=gs, device="cuda", dtype=torch.float64)
    worker = torch.randint(0, nw, (n,), generator=g, device="cuda")
    home = torch.randint(0, nf, (nw,), generator=g, device="cuda")
    move = torch.rand(n, generator=g, device="cuda") >= 0.8
    firm = torch.where(move, torch.randint(0, nf, (n,), generator=g, device="cuda"), home[worker])
    year = torch.randint(0, 30, (n,), generator=g, device="cuda")
    aw = torch.randn(nw, generator=g, device="cuda", dtype=torch.float64)
    x1 = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.3 * aw[worker]
    x2 = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.2 * af[firm]
    y = 0.8 * x1 - 0.3 * x2 + aw[worker] + af[firm] + 0.02 * year + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    d = dict(y=y, x1=x1, x2=x2, worker=worker + rank * nw, firm=firm, year=year)
    return d, {"x1": 0.8, "x2": -0.3}


def shard_c4(n, rank):
    g = gen(777 + rank)
    gs = gen(5)
    ag = torch.randn(1000, generator=gs, device="cuda", dtype=torch.float64)
    ap = torch.randn(24, generator=gs, device="cuda", dtype=torch.float64)
    grp = torch.randint(0, 1000, (n,), generator=g, device="cuda")
    per = torch.randint(0, 24, (n,), generator=g, device="cuda")
    treated = grp < 400
    ar = torch.arange(24, device="cuda")
    eff = torch.where(ar >= 12, 0.5 + 0.05 * (ar - 12), 0.0).double()
    y = ag[grp] + ap[per] + treated * eff[per] + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    d, truth = dict(y=y, group=grp, period=per), {}
    for t in range(24):
        if t == 11:
            continue
        d[f"d{t}"] = (treated & (per == t)).to(torch.uint8)
        truth[f"d{t}"] = float(eff[t] - eff[11])
    return d, truth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4"])
    ap.add_argument("--rows-total", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--max-iter", type=int, default=50)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, 0 if k != "WORLD_SIZE" else 1)) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    eng = get_engine(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        eng.comm_init_from_torch()
    rows_total = a.rows_total or (100_000_000 if a.config == "c3" else 200_000_000)
    n = rows_total // world
    if a.config == "c3":
        d, truth = shard_c3(n, rank, world, rows_total)
        kw = dict(formula="y ~ x1 + x2 | worker + firm + year", vcov="hc1", strategy="alt_proj", max_iter=a.max_iter)
        if world > 1:
            kw["local_fe"] = "worker"
    else:
        d, truth = shard_c4(n, rank)
        kw = dict(formula="y ~ " + " + ".join(truth) + " | group + period", vcov="cluster", cluster_cols=["group"],
                  strategy="compress")

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = leanfe(d, **kw)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    timers = {k_: round(v / a.steps, 3) for k_, v in eng.timers().items() if v > 0}
    b = torch.tensor([res.coefs[c] for c in truth] + [res.std_errors[c] for c in truth], dtype=torch.float64, device="cuda")
    same = True
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        lo, hi = b.clone(), b.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    if rank == 0:
        z = max(abs(res.coefs[c] - truth[c]) / res.std_errors[c] for c in truth)
        out = dict(config=a.config, rows_total=rows_total, n_gpus=world, ms_per_fit=ms,
                   rows_per_s=rows_total / (ms / 1e3), iterations=res.iterations, n_obs=res.n_obs,
                   df_resid=res.df_resid, n_compressed=res.n_compressed, n_clusters=res.n_clusters,
                   fe_dims=list(res.fe_dims) if res.fe_dims else None, truth_max_z=z, identical_on_all_ranks=same,
                   phases_ms=timers, mem_gb=torch.cuda.max_memory_allocated() / 1e9)
        out["ok"] = bool(same and z < 6 and (a.config != "c4" or (res.n_obs == n * world and res.n_compressed <= 24_000)))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
