#!/usr/bin/env python
"""bench.py -- fit rows/s on BASELINE.json config 2 (50M rows, 5 regressors | unit(1M) + time(500),
clustered by unit) and the demean-sweep roofline, with the reference algorithm timed on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R]

A "step" is one complete fit through the public API (`leanfe_b200.leanfe`): factorise the raw int64
ids, singleton filter, entry lists, alternating projections to the reference's tolerance, Gram,
Cholesky solve, cluster-robust sandwich.  `value` = rows/s with the inputs already resident in HBM;
`e2e` = the same call on pinned HOST tensors (H2D inside the timed region, result read back);
`e2e_pageable` / `e2e_pandas` = the same call on a plain NumPy dict / a pandas DataFrame (what the
reference API receives, polars_impl.py:342-347).
With N > 1 (torchrun) every rank owns a disjoint block of units (weak scaling: 50M rows per GPU); the
time-FE tables, Gram and meat are all-reduced over NCCL inside liblfe.  Before the timed region a multi-GPU
run checks the sharded fits against the single-process oracle (`mg_parity`), and every run adds a
`c5_strong` record: BASELINE config 5 (400M rows, 10 regressors, two-way FE) generated on the device by a
counter-based RNG keyed by the GLOBAL row index -- the same panel at every N -- split into arbitrary
(contiguous-index) row shards that the library repartitions by unit itself; `c3_akm` (config 3: 100M rows,
worker + firm + year, HC1) and `c4_yoco` (config 4: 200M rows, 23 dummies, strategy='compress', clustered) are
measured the same way, so every named shape has a driver-visible line at 1, 2, 4 and 8 GPUs.

The CPU arm (`cpu_baseline`, `--impl reference`) fits the SAME full panel on all host cores: the real
reference through Polars when `import polars` works on the box, else the OpenMP C port of its algorithm
(oracle/ref_port.c).  `parity` compares the 50M-row GPU fit with that CPU fit.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_REG = 5
N_TIME = 500
ROWS_PER_UNIT = 50
FORMULA = "y ~ x1 + x2 + x3 + x4 + x5 | unit + time"
C5_K = 10
C5_ROWS_PER_UNIT = 100
C5_FORMULA = "y ~ " + " + ".join(f"x{j}" for j in range(1, C5_K + 1)) + " | unit + time"


def make_panel(n_rows: int, seed: int, unit_offset: int = 0):
    """unit ~ U{0..N/50}, time ~ U{0..500}, x_j ~ N(0,1) + 0.5 a_unit,
    eps = s_unit * N(0,1) with s ~ U(0.5, 1.5), y = sum (1 + 0.1 j) x_j + a_unit + g_time + eps."""
    rng = np.random.default_rng(seed)
    G1 = max(n_rows // ROWS_PER_UNIT, 2)
    unit = rng.integers(0, G1, n_rows)
    tm = rng.integers(0, N_TIME, n_rows)
    a, g, s = rng.normal(size=G1), rng.normal(size=N_TIME), rng.uniform(0.5, 1.5, G1)
    au = a[unit]
    d = {}
    y = au + g[tm] + s[unit] * rng.normal(size=n_rows)
    for j in range(1, K_REG + 1):
        x = rng.normal(size=n_rows)
        x += 0.5 * au
        y += (1 + 0.1 * j) * x
        d[f"x{j}"] = x
    d["y"] = y
    d["unit"] = unit + unit_offset
    d["time"] = tm
    return d, G1


def sample_for_cpu(d: dict, G1: int, target_rows: int):
    """Bounded sample of the same workload: all rows of the first units (keeps ~50 rows per unit, all 500 periods)."""
    n = d["y"].shape[0]
    if target_rows >= n:
        return d, n
    cut = int(G1 * target_rows / n)
    base = int(d["unit"].min())
    m = (d["unit"] - base) < cut
    return {k: v[m] for k, v in d.items()}, int(m.sum())


# ---- c5: counter-based device generator (same panel for every world size) ---------------------------
_M64 = (1 << 64) - 1


def _i64(c: int) -> int:
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix(torch, z):
    """splitmix64 finaliser on an int64 tensor (two's-complement wrap-around arithmetic, logical shifts)."""
    z = z + _i64(0x9E3779B97F4A7C15)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _i64(0x94D049BB133111EB)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def _uniform(torch, idx, stream: int):
    """U(0,1) keyed by (stream, idx): 53 random bits."""
    h = _mix(torch, idx * _i64(0xD1342543DE82EF95) + _i64(stream * 0x632BE59BD9B4E019 + 0x1234567))
    return (((h >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) * (1.0 / (1 << 53))


def _normal(torch, idx, stream: int):
    u1, u2 = _uniform(torch, idx, 2 * stream), _uniform(torch, idx, 2 * stream + 1)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos((2.0 * np.pi) * u2)


def make_c5_rows(torch, dev, lo: int, hi: int, n_total: int, chunk: int = 25_000_000):
    """Rows [lo, hi) of the c5 panel: unit ~ U{0..N/100}, time ~ U{0..500}, 10 regressors
    x_j ~ N(0,1) + 0.5 a_unit, y = sum (1 + 0.05 j) x_j + a_unit + g_time + N(0,1).  Every value is a pure function
    of the GLOBAL row index (and of the unit / period id for a, g), so 1/2/4/8-GPU runs see identical data."""
    G1 = max(n_total // C5_ROWS_PER_UNIT, 2)
    n = hi - lo
    out = {nm: torch.empty(n, dtype=torch.float64, device=dev) for nm in ["y"] + [f"x{j}" for j in range(1, C5_K + 1)]}
    out["unit"] = torch.empty(n, dtype=torch.int64, device=dev)
    out["time"] = torch.empty(n, dtype=torch.int64, device=dev)
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        sl = slice(a - lo, b - lo)
        idx = torch.arange(a, b, dtype=torch.int64, device=dev)
        unit = (_uniform(torch, idx, 1) * G1).to(torch.int64).clamp_(max=G1 - 1)
        tm = (_uniform(torch, idx, 2) * N_TIME).to(torch.int64).clamp_(max=N_TIME - 1)
        au = _normal(torch, unit, 50)
        y = au + _normal(torch, tm, 51) + _normal(torch, idx, 52)
        for j in range(1, C5_K + 1):
            x = _normal(torch, idx, 100 + j) + 0.5 * au
            y += (1 + 0.05 * j) * x
            out[f"x{j}"][sl] = x
        out["y"][sl] = y
        out["unit"][sl] = unit
        out["time"][sl] = tm
    return out


def make_c3_rows(torch, dev, lo: int, hi: int, n_total: int, chunk: int = 25_000_000):
    """Rows [lo, hi) of the c3 AKM panel: worker ~ U{0..N/10}, each worker has a home firm, 80 % of its
    rows are at home and 20 % at a random firm (N/100 firms), year ~ U{0..30}; k = 2.  Pure function of the global
    row index (and of the worker / firm id for the effects)."""
    nw, nf = max(n_total // 10, 10), max(n_total // 100, 5)
    n = hi - lo
    out = {nm: torch.empty(n, dtype=torch.float64, device=dev) for nm in ("y", "x1", "x2")}
    for nm in ("worker", "firm", "year"):
        out[nm] = torch.empty(n, dtype=torch.int64, device=dev)
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        sl = slice(a - lo, b - lo)
        idx = torch.arange(a, b, dtype=torch.int64, device=dev)
        worker = (_uniform(torch, idx, 1) * nw).to(torch.int64).clamp_(max=nw - 1)
        home = (_uniform(torch, worker, 60) * nf).to(torch.int64).clamp_(max=nf - 1)
        firm = torch.where(_uniform(torch, idx, 2) >= 0.8, (_uniform(torch, idx, 3) * nf).to(torch.int64).clamp_(max=nf - 1), home)
        year = (_uniform(torch, idx, 4) * 30).to(torch.int64).clamp_(max=29)
        aw, af = _normal(torch, worker, 61), _normal(torch, firm, 62)
        x1 = _normal(torch, idx, 101) + 0.3 * aw
        x2 = _normal(torch, idx, 102) + 0.2 * af
        out["y"][sl] = 0.8 * x1 - 0.3 * x2 + aw + af + 0.05 * year + _normal(torch, idx, 52)
        out["x1"][sl], out["x2"][sl] = x1, x2
        out["worker"][sl], out["firm"][sl], out["year"][sl] = worker, firm, year
    return out


def make_c4_rows(torch, dev, lo: int, hi: int, chunk: int = 25_000_000):
    """Rows [lo, hi) of the c4 event-study panel (after docs/tutorials/did.qmd:25-103): group ~ U{0..1000},
    period ~ U{0..24}, 40 % of the groups treated from period 12; 23 uint8 dummies treated x 1[period = t], t != 11."""
    n = hi - lo
    out = {"y": torch.empty(n, dtype=torch.float64, device=dev), "group": torch.empty(n, dtype=torch.int64, device=dev),
           "period": torch.empty(n, dtype=torch.int64, device=dev)}
    names = [f"d{t}" for t in range(24) if t != 11]
    for nm in names:
        out[nm] = torch.empty(n, dtype=torch.uint8, device=dev)
    ar = torch.arange(24, device=dev)
    eff = torch.where(ar >= 12, 0.5 + 0.05 * (ar - 12), 0.0).double()
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        sl = slice(a - lo, b - lo)
        idx = torch.arange(a, b, dtype=torch.int64, device=dev)
        grp = (_uniform(torch, idx, 1) * 1000).to(torch.int64).clamp_(max=999)
        per = (_uniform(torch, idx, 2) * 24).to(torch.int64).clamp_(max=23)
        treated = grp < 400
        out["y"][sl] = _normal(torch, grp, 70) + _normal(torch, per, 71) + treated * eff[per] + _normal(torch, idx, 52)
        out["group"][sl], out["period"][sl] = grp, per
        for t in range(24):
            if t != 11:
                out[f"d{t}"][sl] = (treated & (per == t)).to(torch.uint8)
    return out, names


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  Started BEFORE the warm-up (nvidia-smi needs
    ~100 ms to come up); only samples whose timestamp falls inside [mark_start(), stop()] are reported."""

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = None
        q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def stop(self):
        t1 = time.time()
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]),
                             [nm for nm, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= t1 + 0.02]
        if not inside and rows:          # region shorter than the sampling period: take the closest samples
            mid = 0.5 * ((self.t0 or t1) + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
        reasons = sorted({x for r in inside for x in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "power_w_max": max(r[3] for r in inside) if inside else None,
                "samples": len(inside), "reasons": reasons}


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads (and therefore the first-touch placement of its pinned buffers) to the NUMA
    node its GPU hangs off, so the H2D copies of the end-to-end arm do not cross the socket interconnect when
    eight ranks share one host.  Best effort: returns a short description, never raises."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return {"node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:       # no sysfs / no permission / old torch: run unbound
        return {"bound": False, "why": type(e).__name__}


# ---- the CPU arm: the real reference when Polars is importable, else the OpenMP port ------------------
def live_reference():
    """Probe for the UNMODIFIED reference (python/leanfe/leanfe.py:14-184 -> polars_impl.py:468-537).  It needs
    Polars, which this image does not ship; `baseline/_ref` holds the reference package itself (pip --no-deps).
    -> (callable | None, why)"""
    try:
        import polars  # noqa: F401
    except Exception as e:
        return None, f"import polars failed ({type(e).__name__}): the reference's engine is not installed on this box"
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/python"):
        if os.path.isdir(os.path.join(cand, "leanfe")) and cand not in sys.path:
            sys.path.insert(0, cand)
    try:
        from leanfe import leanfe as ref_leanfe
        return ref_leanfe, "polars importable"
    except Exception as e:
        return None, f"polars importable but `import leanfe` failed ({type(e).__name__}: {e})"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(d: dict, steps: int, warmup: int):
    """Time the reference's algorithm on ALL host cores on the panel `d`.  -> dict(rate, cores, kind, note, sec,
    beta, se, iterations, n_obs)"""
    X = [d[f"x{j}"] for j in range(1, K_REG + 1)]
    n = d["y"].shape[0]
    ref_leanfe, why = live_reference()
    times = []
    if ref_leanfe is not None:
        import polars as pl
        df = pl.DataFrame({k: v for k, v in d.items()})
        res = None
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            res = ref_leanfe(df, formula=FORMULA, vcov="cluster", cluster_cols=["unit"], backend="polars", strategy="alt_proj")
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        names = [f"x{j}" for j in range(1, K_REG + 1)]
        return dict(rate=n * len(times) / sum(times), cores=host_cores(), kind="reference", note=why,
                    sec=float(np.mean(times)), beta=np.array([res.coefs[c] for c in names]),
                    se=np.array([res.std_errors[c] for c in names]), iterations=int(res.iterations),
                    n_obs=int(res.n_obs))
    from oracle import ref_port as P
    cores = P.set_threads(host_cores())      # torchrun exports OMP_NUM_THREADS=1 to its workers: set it explicitly
    res = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = P.fit(d["y"], X, [d["unit"], d["time"]], cluster=d["unit"], vcov="cluster", ssc=True, tol=1e-6, max_iter=50)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return dict(rate=n * len(times) / sum(times), cores=cores, kind="port",
                note=f"OpenMP C restatement of the reference algorithm (oracle/ref_port.c); live reference unavailable: {why}",
                sec=float(np.mean(times)), beta=res["beta"], se=res["se"], iterations=res["iterations"], n_obs=res["n_obs"])


def ncu_traffic(key: str):
    """dram__bytes_read + dram__bytes_write of one sweep's projection launches, from the committed ncu summary."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(key)
        return (float(e["bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=50_000_000, help="rows per GPU")
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the CPU arm (default: the full per-GPU panel)")
    ap.add_argument("--c5-rows", type=int, default=400_000_000, help="TOTAL rows of the c5 strong-scaling record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arms (profiling runs)")
    ap.add_argument("--c3-rows", type=int, default=100_000_000, help="TOTAL rows of the c3 (AKM, 3-way FE) record")
    ap.add_argument("--c4-rows", type=int, default=200_000_000, help="TOTAL rows of the c4 (YOCO) record")
    ap.add_argument("--no-c5", action="store_true", help="skip the c5 strong-scaling record")
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 records")
    ap.add_argument("--no-mg-parity", action="store_true", help="skip the multi-GPU oracle check (world > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_rows = args.rows if args.cpu_rows is None else args.cpu_rows
    workload = f"c2: {args.rows // 1_000_000}M rows/GPU, 5 regressors | unit({args.rows // ROWS_PER_UNIT}) + time(500), cluster(unit), ssc"

    if args.impl == "reference":
        if rank != 0:
            return
        d, G1 = make_panel(args.rows, 42)
        ds, ns = sample_for_cpu(d, G1, cpu_rows)
        del d
        r = cpu_reference_run(ds, args.steps, max(args.warmup, 1))
        sample = (f"the full {ns}-row panel of one GPU" if ns == args.rows else
                  f"{ns} rows = all rows of the first {ns // ROWS_PER_UNIT} units of the {args.rows}-row panel")
        if args.gpus > 1:
            sample += f" (the {args.gpus}-GPU arm fits {args.gpus} such shards, weak scaling)"
        print(json.dumps({
            "impl": "reference", "metric": "fit rows/s", "value": r["rate"], "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample, "sweeps": r["iterations"]},
            "cpu_baseline": {"value": r["rate"], "unit": "rows/s", "cores": r["cores"], "kind": r["kind"], "sample": sample,
                             "note": r["note"], "host_cpus": os.cpu_count()},
            "e2e": {"value": r["rate"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "coefs": {f"x{j + 1}": float(b) for j, b in enumerate(r["beta"])},
        }))
        return

    import torch
    import torch.distributed as dist
    from leanfe_b200 import leanfe
    from leanfe_b200.engine import get_engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = get_engine(local_rank)
    if world > 1:
        eng.comm_init_from_torch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- multi-GPU parity against the single-process oracle, before anything is timed ------------------
    mg_parity = None
    if world > 1 and not args.no_mg_parity:
        spec = importlib.util.spec_from_file_location("multigpu_check", os.path.join(ROOT, "tests", "multigpu_check.py"))
        mgc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mgc)
        t0 = time.perf_counter()
        fails = mgc.run_checks(rank, world)
        bad = torch.tensor([float(len(fails))], dtype=torch.float64, device=dev)
        dist.all_reduce(bad)
        mg_parity = {"ok": int(bad.item()) == 0, "world": world, "failed_cases_over_ranks": int(bad.item()),
                     "cases": "two-way iid/hc1/cluster(unit)/cluster(time) pre-partitioned + arbitrary shards "
                              "(library repartition); one-way 'demean' from arbitrary shards; three-way weighted / hc1 / "
                              "cluster(firm) both ways; YOCO x5",
                     "tolerance": "beta 1e-8, se 1e-6, iterations / n_obs / df_resid / fe_dims / n_clusters equal, "
                                  "results bit-identical on all ranks", "seconds": round(time.perf_counter() - t0, 1)}
        if fails and rank == 0:
            print(f"mg_parity FAILED: {fails[:2]}", file=sys.stderr)

    # ---- synthetic panel in pinned host memory (rank r owns units [r*G1, (r+1)*G1)) ---------------
    n = args.rows
    d, G1 = make_panel(n, 42 + rank, unit_offset=rank * max(n // ROWS_PER_UNIT, 2))
    host = {}
    for k_, v in d.items():
        t = torch.empty(v.shape, dtype=torch.from_numpy(v[:1]).dtype, pin_memory=True)
        t.numpy()[...] = v
        host[k_] = t
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    kw = dict(formula=FORMULA, vcov="cluster", cluster_cols=["unit"], strategy="alt_proj")
    if world > 1:
        kw["local_fe"] = "unit"      # rows are sharded by unit: its tables, counts and cluster scores stay rank-local

    # ---- device-resident arm ----------------------------------------------------------------------
    resident = {k_: t.to(dev) for k_, t in host.items()}
    res = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        res = leanfe(resident, **kw)
    eng.set_profiling(True)
    eng.reset_timers()
    barrier()
    if sampler:
        sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = leanfe(resident, **kw)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    timers = eng.timers()
    counters = eng.counters()
    eng.set_profiling(False)
    clocks = sampler.stop() if sampler else None
    del resident
    torch.cuda.empty_cache()

    # ---- end-to-end arms: host buffers -> public API -> result on the host ------------------------
    def timed_e2e(data, steps):
        for _ in range(max(1, min(args.warmup, 2))):
            leanfe(data, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r2 = leanfe(data, **kw)
            _ = float(r2.coefs["x1"])
        barrier()
        return (time.perf_counter() - t0) / steps

    e2e_s = float("nan")
    h2d_only_ms = None
    e2e_extra = {}
    if not args.no_e2e:
        # the PCIe floor of the end-to-end arm: the same host tensors copied to the device with no compute
        dst = {k_: torch.empty_like(t, device=dev) for k_, t in host.items()}
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            for k_, t in host.items():
                dst[k_].copy_(t, non_blocking=True)
        c1.record()
        barrier()
        h2d_only_ms = c0.elapsed_time(c1) / 3
        del dst
        e2e_s = timed_e2e(host, args.steps) * args.steps
        if world == 1:
            # what the reference API actually receives: pageable NumPy columns / a pandas frame
            sec = timed_e2e(d, max(2, args.steps // 2))
            e2e_extra["e2e_pageable"] = {"value": n / sec, "unit": "rows/s", "ms_per_step": sec * 1e3,
                                         "input": "dict of pageable NumPy arrays (int64 ids, float64 columns)",
                                         "h2d_bytes_per_step": h2d_bytes}
            try:
                import pandas as pd
                pdf = pd.DataFrame(d, copy=False)
                sec = timed_e2e(pdf, max(2, args.steps // 2))
                e2e_extra["e2e_pandas"] = {"value": n / sec, "unit": "rows/s", "ms_per_step": sec * 1e3,
                                           "input": "pandas DataFrame", "h2d_bytes_per_step": h2d_bytes}
                del pdf
            except Exception as e:       # pandas missing: report, do not fail the bench
                e2e_extra["e2e_pandas"] = {"unavailable": f"{type(e).__name__}: {e}"}

    if world > 1:
        tt = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    else:
        e2e_ms = e2e_s * 1e3
    del host
    if world > 1:
        del d

    # ---- the other named shapes, strong scaling: the SAME panel at every N, arbitrary row shards in ---------
    def strong_record(label, shard, kw_fit, n_tot, steps_s=2):
        """Fit `shard` (this rank's contiguous-global-row-index block of an n_tot-row panel) steps_s times; the
        library repartitions the rows by the highest-cardinality FE inside the timed region."""
        r = leanfe(shard, **kw_fit)       # warm-up (NCCL all-to-all buffers, pool growth)
        eng.set_profiling(True)
        eng.reset_timers()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps_s):
            r = leanfe(shard, **kw_fit)
        a1.record()
        barrier()
        ms_s = a0.elapsed_time(a1) / steps_s
        tm = eng.timers()
        cn = eng.counters()
        eng.set_profiling(False)
        if world > 1:
            tt_ = torch.tensor([ms_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
            ms_s = float(tt_[0])
        return {"workload": label + "; identical panel at every N (counter-based RNG keyed by the global row index)",
                "scaling": "strong", "n_gpus": world, "rows_total": n_tot, "ms_per_fit": ms_s, "value": n_tot / (ms_s / 1e3),
                "unit": "rows/s", "input": "device-resident; rank r holds global rows [r N/W, (r+1) N/W) -- arbitrary with "
                "respect to every fixed effect" + ("" if world == 1 else "; the library's repartition / partial-table merge "
                                                   "runs inside the timed region"),
                "sweeps": r.iterations, "converged": r.converged, "n_obs": r.n_obs, "strategy": r.strategy,
                "n_compressed": r.n_compressed, "collectives_per_fit": cn["collectives"] / steps_s,
                "phases_ms_per_fit": {k_: round(v / steps_s, 3) for k_, v in tm.items() if v > 0},
                "coefs": {k_: float(v) for k_, v in list(r.coefs.items())[:10]},
                "std_errors": {k_: float(v) for k_, v in list(r.std_errors.items())[:10]}}

    def guarded(fn):
        """A side record must never cost the headline line: a failure (e.g. out of memory on a smaller part) is
        reported in the record instead.  Every rank takes the same branch (the failure flag is all-reduced)."""
        rec, err = None, None
        try:
            rec = fn()
        except Exception as e:       # noqa: BLE001
            err = f"{type(e).__name__}: {e}"[:300]
            eng.set_profiling(False)
        if world > 1:
            bad_ = torch.tensor([1.0 if err else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(bad_)
            if bad_.item() > 0 and err is None:
                err = "failed on another rank"
        torch.cuda.empty_cache()
        return {"unavailable": err} if err else rec

    def run_c5():
        n5 = args.c5_rows
        shard5 = make_c5_rows(torch, dev, n5 * rank // world, n5 * (rank + 1) // world, n5)
        return strong_record(f"c5: {n5 // 1_000_000}M rows TOTAL, 10 regressors | unit({n5 // C5_ROWS_PER_UNIT}) + time(500), iid",
                             shard5, dict(formula=C5_FORMULA, vcov="iid", strategy="alt_proj"), n5)

    def run_c3():
        n3 = args.c3_rows
        shard3 = make_c3_rows(torch, dev, n3 * rank // world, n3 * (rank + 1) // world, n3)
        rec = strong_record(f"c3: AKM {n3 // 1_000_000}M rows TOTAL, 2 regressors | worker({n3 // 10}) + firm({n3 // 100}) + "
                            "year(30), HC1, 50 sweeps", shard3,
                            dict(formula="y ~ x1 + x2 | worker + firm + year", vcov="hc1", strategy="alt_proj"), n3)
        # in-place-equivalent sweep rate (N F (16 C + 8) bytes per sweep of the reference's formulation)
        sw_ms = (rec["phases_ms_per_fit"].get("demean_sweeps", 0) + rec["phases_ms_per_fit"].get("demean_check", 0)) / max(rec["sweeps"], 1)
        if sw_ms > 0:
            rec["sweep_ms"] = sw_ms
            rec["sweep_inplace_equivalent_gbs_all_gpus"] = n3 * 3 * (16 * 3 + 8) / (sw_ms / 1e3) / 1e9
        return rec

    def run_c4():
        n4 = args.c4_rows
        shard4, dnames = make_c4_rows(torch, dev, n4 * rank // world, n4 * (rank + 1) // world)
        return strong_record(f"c4: YOCO {n4 // 1_000_000}M rows TOTAL, 23 treatment x period dummies | group(1000) + period(24), "
                             "cluster(group)", shard4,
                             dict(formula="y ~ " + " + ".join(dnames) + " | group + period", vcov="cluster",
                                  cluster_cols=["group"], strategy="compress"), n4, steps_s=3)

    c5 = c3 = c4 = None
    torch.cuda.empty_cache()
    if not args.no_c5:
        c5 = guarded(run_c5)
    if not args.no_extra:
        c3 = guarded(run_c3)
        c4 = guarded(run_c4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_rows = n * world
    value = total_rows * args.steps / (ms / 1e3)
    e2e = total_rows * args.steps / (e2e_ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    # ---- roofline of the demean sweep (dominant iterative kernel k_proj) -----------------------------
    C, F = K_REG + 1, 2
    sweeps = res.iterations
    # every k_proj launch (regular sweeps + the speculative first projection of the stopping test, which is the
    # next sweep's first step) is timed by the library's CUDA events on the launching stream
    proj_launches = max(counters["sweep_launches"] / args.steps, 1)
    proj_ms = (timers["demean_sweeps"] + timers["demean_check"]) / args.steps / proj_launches
    sweep_ms = F * proj_ms
    # bytes the coefficient-space sweep must move: per projection one packed 32-bit (unit code, time code) word
    # per entry of the FE-sorted list, plus one read of S0 and one read-modify-write of alpha/acc per group row
    # (SURVEY 8d 'read-only sweep' form)
    G_unit, G_time = n // ROWS_PER_UNIT, N_TIME
    own_bytes = F * n * 4 + 5 * 8 * C * (G_unit + G_time)
    inplace_bytes = n * F * (16 * C + 8)                    # the reference's in-place formulation
    traffic, traffic_src = ncu_traffic("c2_50M_sweep") if (n == 50_000_000 and world == 1) else (None, None)
    roofline = {
        "bound": "hbm", "kernel": "k_proj_q<6,1> (time FE, unit table gathered from L2) + k_proj_rep<6,8> (unit FE, replicated period table "
                                  "in shared memory) + k_finalize (one sweep = 2 projections)",
        "achieved": own_bytes / (sweep_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": own_bytes / (sweep_ms / 1e3) / 1e9 / peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "basis": "achieved/frac = bytes this implementation's sweep must move (packed code words + tables) / time; "
                 "inplace_equivalent = SURVEY 8(d) bytes of the reference's in-place sweep / the same time",
        "limiter": "not HBM: the L1TEX data pipe (l1tex__data_pipe_lsu_wavefronts 76-79 % of peak in "
                   "profiles/r2_ncu_c2_kernels.md) -- table-row gathers from L2 / shared memory; the sweep reads only the "
                   "packed code words",
        "peak_source": peak_src, "bytes_per_sweep": own_bytes, "sweep_ms": sweep_ms, "sweeps_per_fit": sweeps,
        "proj_launches_per_fit": proj_launches,
        "inplace_equivalent": {"bytes_per_sweep": inplace_bytes, "achieved": inplace_bytes / (sweep_ms / 1e3) / 1e9,
                               "frac": inplace_bytes / (sweep_ms / 1e3) / 1e9 / peak,
                               "note": "SURVEY 8(d) figure N*F*(16C+8) for the reference's in-place sweeps; this "
                                       "implementation iterates in FE-coefficient space and never re-reads the columns"},
    }
    phases = {k_: round(v / args.steps, 4) for k_, v in timers.items() if v > 0}

    out = {
        "metric": "fit rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rows_total": total_rows, "sweeps": sweeps, "converged": bool(res.converged),
                   "l2": "inputs (3.2 GB/GPU) exceed the 126 MB L2; no explicit flush", "tol": 1e-6,
                   "parallelism": f"row-sharded by unit x{world}" if world > 1 else "single GPU",
                   "host_numa": numa},
        "e2e": {"value": e2e, "unit": "rows/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8 * (2 * K_REG + 3),
                "ms_per_step": e2e_ms / args.steps, "input": "pinned host torch tensors",
                # the same 3.2 GB copied host -> device with nothing else running: what PCIe alone costs per step
                "h2d_only_ms_per_step": h2d_only_ms,
                "h2d_aggregate_gbs": (h2d_bytes * world / (h2d_only_ms / 1e3) / 1e9) if h2d_only_ms else None},
        "gpu_launches": int(counters["launches"]),
        "collectives_per_fit": counters["collectives"] / args.steps,
        "clocks": clocks, "roofline": roofline, "phases_ms_per_step": phases,
        "coefs": {k_: float(v) for k_, v in res.coefs.items()},
    }
    out.update(e2e_extra)
    if mg_parity is not None:
        out["mg_parity"] = mg_parity
    if c5 is not None:
        out["c5_strong"] = c5
    if c3 is not None:
        out["c3_akm"] = c3
    if c4 is not None:
        out["c4_yoco"] = c4
    if world == 1 and not args.no_cpu_baseline:
        ds, ns = sample_for_cpu(d, G1, cpu_rows)
        r = cpu_reference_run(ds, 1, 1)
        sample = (f"the full {ns}-row panel (same arrays as the GPU fit)" if ns == n else
                  f"{ns} rows = all rows of the first {ns // ROWS_PER_UNIT} units of the same panel")
        out["cpu_baseline"] = {"value": r["rate"], "unit": "rows/s", "cores": r["cores"], "kind": r["kind"],
                               "seconds": r["sec"], "sample": sample, "sweeps": r["iterations"], "note": r["note"],
                               "host_cpus": os.cpu_count()}
        if ns == n:
            names = [f"x{j}" for j in range(1, K_REG + 1)]
            gb = np.array([res.coefs[c] for c in names])
            gs = np.array([res.std_errors[c] for c in names])
            out["parity"] = {"against": f"cpu_baseline ({r['kind']}) on the same {n} rows",
                             "beta_relerr": float(np.max(np.abs(gb - r["beta"]) / np.abs(r["beta"]))),
                             "se_relerr": float(np.max(np.abs(gs - r["se"]) / np.abs(r["se"]))),
                             "iterations_equal": bool(res.iterations == r["iterations"]),
                             "n_obs_equal": bool(res.n_obs == r["n_obs"]),
                             "tolerance": {"beta": 1e-8, "se": 1e-6}}
            out["parity"]["ok"] = bool(out["parity"]["beta_relerr"] < 1e-8 and out["parity"]["se_relerr"] < 1e-6
                                       and out["parity"]["iterations_equal"] and out["parity"]["n_obs_equal"])
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
