"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank holds the rows of a disjoint block of units (the shard key); the fit must equal the oracle's
single-process fit of the whole panel (beta 1e-8, SE 1e-6, equal iteration count) and be identical on
all ranks.  Exit code 0 on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_checks(rank: int, world: int) -> list:
    """All parity panels against the single-process oracle; needs torch.distributed (nccl) initialised and the
    engine's communicator attached.  -> list of failure records (empty = parity green).  Also called by bench.py
    before the timed region of every multi-GPU run (`mg_parity`)."""
    import torch
    import torch.distributed as dist

    from leanfe_b200 import leanfe
    from leanfe_b200.sharding import shard_rows_by_key
    from oracle import ref_numpy as O

    rng = np.random.default_rng(123)                      # same panel on every rank
    N, G1, G2 = 600_000, 12_000, 60
    unit = rng.integers(0, G1, N) * 3 + 7                 # non-dense ids
    time_ = rng.integers(0, G2, N)
    time_[unit < 40] = 1                                   # some periods exist on one shard only
    a, g = rng.normal(size=G1 * 3 + 8), rng.normal(size=G2)
    x1, x2 = rng.normal(size=N) + 0.5 * a[unit], rng.normal(size=N) - 0.3 * a[unit]
    y = 1.5 * x1 - 0.7 * x2 + a[unit] + g[time_] + rng.normal(size=N) * (1 + 0.3 * (unit % 5))
    unit[:50] = np.arange(900_000, 900_050)               # singletons -> dropped
    full = dict(y=y, x1=x1, x2=x2, unit=unit, time=time_)
    mine = shard_rows_by_key(unit, world)[rank]
    shard = {k: v[mine] for k, v in full.items()}

    failures = []
    for vc, cl in (("iid", None), ("hc1", None), ("cluster", ["unit"]), ("cluster", ["time"])):
        kw = dict(formula="y ~ x1 + x2 | unit + time", vcov=vc, cluster_cols=cl, strategy="alt_proj")
        res = leanfe(shard, local_fe="unit", **kw)
        ref = O.leanfe_oracle(full, **kw)
        b, eb = np.array(list(res.coefs.values())), np.array(list(ref["coefs"].values()))
        s, es = np.array(list(res.std_errors.values())), np.array(list(ref["std_errors"].values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.iterations == ref["iterations"] and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and tuple(res.fe_dims) == tuple(ref["fe_dims"]) and res.n_clusters == ref["n_clusters"])
        # identical on all ranks
        t = torch.tensor(list(b) + list(s), dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and bool(torch.equal(lo, hi))
        if not ok:
            failures.append((vc, cl, b, eb, s, es, res.iterations, ref["iterations"], res.n_obs, ref["n_obs"],
                             res.fe_dims, ref["fe_dims"], res.n_clusters, ref["n_clusters"]))
    # ARBITRARY row shards (no local_fe): the library repartitions by the highest-cardinality FE itself
    # (lfe_shard_plan / lfe_shard_exchange); shard sizes are deliberately unequal (rank 0 holds a third of its share)
    owner = (np.arange(N) * 7919) % world
    owner[(owner == 0) & (rng.random(N) < 0.67)] = world - 1
    mine_a = np.flatnonzero(owner == rank)
    shard_a = {k: v[mine_a] for k, v in full.items()}
    for vc, cl in (("iid", None), ("cluster", ["unit"]), ("cluster", ["time"])):
        kw = dict(formula="y ~ x1 + x2 | unit + time", vcov=vc, cluster_cols=cl, strategy="alt_proj")
        res = leanfe(shard_a, **kw)
        ref = O.leanfe_oracle(full, **kw)
        b, eb = np.array(list(res.coefs.values())), np.array(list(ref["coefs"].values()))
        s, es = np.array(list(res.std_errors.values())), np.array(list(ref["std_errors"].values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.iterations == ref["iterations"] and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and tuple(res.fe_dims) == tuple(ref["fe_dims"]) and res.n_clusters == ref["n_clusters"])
        t = torch.tensor(list(b) + list(s), dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and bool(torch.equal(lo, hi))
        if not ok:
            failures.append(("arbitrary-shards", vc, cl, b, eb, s, es, res.iterations, ref["iterations"], res.n_obs,
                             ref["n_obs"], res.fe_dims, ref["fe_dims"], res.n_clusters, ref["n_clusters"]))
    # single FE ('demean': the within transform), arbitrary shards: the library repartitions by that FE
    for vc, cl in (("iid", None), ("cluster", ["unit"])):
        kw = dict(formula="y ~ x1 + x2 | unit", vcov=vc, cluster_cols=cl, strategy="demean")
        res = leanfe(shard_a, **kw)
        ref = O.leanfe_oracle(full, **kw)
        b, eb = np.array(list(res.coefs.values())), np.array(list(ref["coefs"].values()))
        s, es = np.array(list(res.std_errors.values())), np.array(list(ref["std_errors"].values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and tuple(res.fe_dims) == tuple(ref["fe_dims"]) and res.n_clusters == ref["n_clusters"])
        if not ok:
            failures.append(("demean-arbitrary-shards", vc, cl, b, eb, s, es, res.n_obs, ref["n_obs"], res.df_resid,
                             ref["df_resid"], res.fe_dims, ref["fe_dims"]))
    # three-way FE (worker = shard key, firm and year shared by all ranks), weighted, max_iter reached
    nw, nf = 9000, 700
    worker = rng.integers(0, nw, N)
    firm = np.where(rng.random(N) < 0.8, rng.integers(0, nf, nw)[worker], rng.integers(0, nf, N))
    year = rng.integers(0, 12, N)
    wgt = rng.uniform(0.5, 2.0, N)
    z1 = rng.normal(size=N)
    yy = 0.4 * z1 + rng.normal(size=nw)[worker] + rng.normal(size=nf)[firm] + 0.1 * year + rng.normal(size=N)
    full3 = dict(y=yy, z1=z1, worker=worker, firm=firm, year=year, wgt=wgt)
    mine3 = shard_rows_by_key(worker, world)[rank]
    shard3 = {k: v[mine3] for k, v in full3.items()}
    for vc, cl, wname in (("hc1", None, None), ("iid", None, "wgt"), ("cluster", ["firm"], None)):
        kw = dict(formula="y ~ z1 | worker + firm + year", vcov=vc, cluster_cols=cl, weights=wname, strategy="alt_proj",
                  max_iter=12)
        res = leanfe(shard3, local_fe="worker", **kw)
        ref = O.leanfe_oracle(full3, **kw)
        b, eb = np.array(list(res.coefs.values())), np.array(list(ref["coefs"].values()))
        s, es = np.array(list(res.std_errors.values())), np.array(list(ref["std_errors"].values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.iterations == ref["iterations"] and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and tuple(res.fe_dims) == tuple(ref["fe_dims"]) and res.n_clusters == ref["n_clusters"])
        if not ok:
            failures.append(("3way", vc, cl, wname, b, eb, s, es, res.iterations, ref["iterations"], res.n_obs,
                             ref["n_obs"], res.fe_dims, ref["fe_dims"]))
        # the same fit from round-robin row shards, repartitioned by the library (weights travel with the rows)
        shard3a = {k: v[rank::world] for k, v in full3.items()}
        res = leanfe(shard3a, **kw)
        b = np.array(list(res.coefs.values()))
        s = np.array(list(res.std_errors.values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.iterations == ref["iterations"] and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and tuple(res.fe_dims) == tuple(ref["fe_dims"]) and res.n_clusters == ref["n_clusters"])
        if not ok:
            failures.append(("3way-arbitrary-shards", vc, cl, wname, b, eb, s, es, res.iterations, ref["iterations"],
                             res.n_obs, ref["n_obs"], res.fe_dims, ref["fe_dims"]))
    # YOCO (strategy='compress'): rows sharded ARBITRARILY (round-robin), every key spans ranks; one dummy is
    # non-zero on rank 0's rows only, one period exists on rank 1's rows only (shared code space needed)
    Ny, ng, npd = 400_000, 150, 10
    grp, per = rng.integers(0, ng, Ny), rng.integers(0, npd - 1, Ny)
    per[(np.arange(Ny) % world == min(1, world - 1)) & (rng.random(Ny) < 0.02)] = npd - 1
    treated = (grp % 5 < 2)
    d1 = (treated & (per >= 5)).astype(np.uint8)
    d2 = (treated & (per == 7)).astype(np.uint8)
    d3 = ((np.arange(Ny) % world == 0) & (rng.random(Ny) < 0.1)).astype(np.uint8)
    dose = rng.integers(0, 4, Ny).astype(np.float64)          # integer-valued float regressor
    wy = rng.integers(1, 4, Ny).astype(np.float64)
    yz = 0.8 * d1 - 0.3 * d2 + 0.2 * d3 + 0.1 * dose + rng.normal(size=ng)[grp] + 0.05 * per + rng.normal(size=Ny)
    fully = dict(y=yz, d1=d1, d2=d2, d3=d3, dose=dose, grp=grp, per=per, wy=wy)
    mine_y = np.arange(rank, Ny, world)
    shardy = {k: v[mine_y] for k, v in fully.items()}
    for vc, cl, wname in (("iid", None, None), ("hc1", None, None), ("cluster", ["grp"], None),
                          ("cluster", ["grp", "per"], None), ("iid", None, "wy")):
        kw = dict(formula="y ~ d1 + d2 + d3 + dose | grp + per", vcov=vc, cluster_cols=cl, weights=wname,
                  strategy="compress")
        res = leanfe(shardy, **kw)
        ref = O.leanfe_oracle(fully, **kw)
        b, eb = np.array(list(res.coefs.values())), np.array(list(ref["coefs"].values()))
        s, es = np.array(list(res.std_errors.values())), np.array(list(ref["std_errors"].values()))
        ok = (np.max(np.abs(b - eb) / np.abs(eb)) < 1e-8 and np.max(np.abs(s - es) / es) < 1e-6
              and res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
              and res.n_compressed == ref["n_compressed"] and res.n_clusters == ref["n_clusters"])
        t = torch.tensor(list(b) + list(s), dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and bool(torch.equal(lo, hi))
        if not ok:
            failures.append(("yoco", vc, cl, wname, b, eb, s, es, res.n_obs, ref["n_obs"], res.df_resid, ref["df_resid"],
                             res.n_compressed, ref["n_compressed"], res.n_clusters, ref["n_clusters"]))
    return failures


def main():
    import torch
    import torch.distributed as dist

    from leanfe_b200.engine import get_engine

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine(local)
    eng.comm_init_from_torch()
    failures = run_checks(rank, world)
    if rank == 0:
        print("multigpu_check:", "OK" if not failures else f"FAILED {failures}", f"(world={world})")
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
