"""Full-size runs of BASELINE.json configs 3, 4 and 5 on one B200, checked through size-independent
properties (the oracle cannot run 100M-400M rows in seconds).  Not part of the pytest suite: run with

    python tests/scale_check.py [c3|c4|c5|all] [--scale 1.0] > gpurun_out/scale.json

Data are generated on the device with a seeded torch generator (no 40 GB host arrays).  Properties:
  * truth recovery: |beta - beta_true| within 6 reported standard errors;
  * FWL fixed point (c3, c5): re-demeaning the demeaned columns moves them by < 10 * tol;
  * linearity (c5): fitting y and 2*y + x1 gives beta' = 2*beta + e1 to 1e-8 (same sweep count);
  * YOCO (c4): bit-exact group counts sum to N; compress == alt_proj on the same rows to 1e-8 (beta), 1e-6 (SE);
  * sample parity: the same generator at 1/50 scale is compared with the NumPy oracle (beta 1e-8, SE 1e-6).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from leanfe_b200 import leanfe  # noqa: E402
from leanfe_b200.engine import get_engine  # noqa: E402
from oracle import ref_numpy as O  # noqa: E402


def _gen(seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def panel_c3(n, seed=42):
    """AKM-style: worker(n/10) + firm(n/100) + year(30), 80% of a worker's rows at the home firm; k = 2; HC1."""
    g = _gen(seed)
    nw, nf = max(n // 10, 10), max(n // 100, 5)
    worker = torch.randint(0, nw, (n,), generator=g, device="cuda")
    home = torch.randint(0, nf, (nw,), generator=g, device="cuda")
    move = torch.rand(n, generator=g, device="cuda") >= 0.8
    firm = torch.where(move, torch.randint(0, nf, (n,), generator=g, device="cuda"), home[worker])
    year = torch.randint(0, 30, (n,), generator=g, device="cuda")
    aw = torch.randn(nw, generator=g, device="cuda", dtype=torch.float64)
    af = torch.randn(nf, generator=g, device="cuda", dtype=torch.float64)
    x1 = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.3 * aw[worker]
    x2 = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.2 * af[firm]
    y = 0.8 * x1 - 0.3 * x2 + aw[worker] + af[firm] + 0.05 * year + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    return dict(y=y, x1=x1, x2=x2, worker=worker, firm=firm, year=year), {"x1": 0.8, "x2": -0.3}


def panel_c4(n, seed=42):
    """Event study: group(1000) x period(24), 40% of groups treated from period 12; 23 treated x period dummies."""
    g = _gen(seed)
    grp = torch.randint(0, 1000, (n,), generator=g, device="cuda")
    per = torch.randint(0, 24, (n,), generator=g, device="cuda")
    treated = grp < 400
    ag = torch.randn(1000, generator=g, device="cuda", dtype=torch.float64)
    ap = torch.randn(24, generator=g, device="cuda", dtype=torch.float64)
    eff = torch.where(torch.arange(24, device="cuda") >= 12, 0.5 + 0.05 * (torch.arange(24, device="cuda") - 12), 0.0).double()
    y = ag[grp] + ap[per] + treated * eff[per] + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    d = dict(y=y, group=grp, period=per)
    truth = {}
    for t in range(24):
        if t == 11:
            continue                                   # reference period
        d[f"d{t}"] = (treated & (per == t)).to(torch.uint8)
        truth[f"d{t}"] = float(eff[t] - eff[11])
    return d, truth


def panel_c5(n, seed=42, k=10):
    g = _gen(seed)
    nu = max(n // 100, 10)
    unit = torch.randint(0, nu, (n,), generator=g, device="cuda")
    tm = torch.randint(0, 500, (n,), generator=g, device="cuda")
    a = torch.randn(nu, generator=g, device="cuda", dtype=torch.float64)
    gt = torch.randn(500, generator=g, device="cuda", dtype=torch.float64)
    d, truth = dict(unit=unit, time=tm), {}
    y = a[unit] + gt[tm] + torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    for j in range(1, k + 1):
        x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) + 0.5 * a[unit]
        y += (1 + 0.1 * j) * x
        d[f"x{j}"], truth[f"x{j}"] = x, 1 + 0.1 * j
    d["y"] = y
    return d, truth


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def truth_ok(res, truth):
    worst = 0.0
    for name, b in truth.items():
        worst = max(worst, abs(res.coefs[name] - b) / max(res.std_errors[name], 1e-300))
    return worst


def oracle_sample(make, kw, n_small):
    d, _ = make(n_small)
    res = leanfe(d, **kw)
    ref = O.leanfe_oracle({k_: v.cpu().numpy() for k_, v in d.items()}, **kw)
    b = np.array(list(res.coefs.values()))
    eb = np.array(list(ref["coefs"].values()))
    s = np.array(list(res.std_errors.values()))
    es = np.array(list(ref["std_errors"].values()))
    return dict(n=n_small, beta_relerr=float(np.max(np.abs(b - eb) / np.abs(eb))), se_relerr=float(np.max(np.abs(s - es) / es)),
                iterations=[res.iterations, ref["iterations"]], n_obs=[res.n_obs, ref["n_obs"]])


def run_c3(scale):
    n = int(100_000_000 * scale)
    kw = dict(formula="y ~ x1 + x2 | worker + firm + year", vcov="hc1", strategy="alt_proj")
    d, truth = panel_c3(n)
    res, sec = timed(lambda: leanfe(d, **kw))
    res2, sec2 = timed(lambda: leanfe(d, profile=True, **kw))
    out = dict(config="c3", n=n, seconds=sec2, rows_per_s=n / sec2, iterations=res.iterations, converged=res.converged,
               n_obs=res.n_obs, singletons_dropped=n - res.n_obs, fe_dims=res.fe_dims, coefs=res.coefs,
               truth_max_z=truth_ok(res, truth), timings=res2.timings,
               repeat_equal=max(abs(res.coefs[c] - res2.coefs[c]) / abs(res.coefs[c]) for c in res.coefs))
    del d
    torch.cuda.empty_cache()
    out["oracle_sample"] = oracle_sample(panel_c3, kw, max(n // 50, 20_000))
    return out


def run_c4(scale):
    n = int(200_000_000 * scale)
    d, truth = panel_c4(n)
    xs = " + ".join(truth.keys())
    kw = dict(formula=f"y ~ {xs} | group + period", vcov="cluster", cluster_cols=["group"])
    res, sec = timed(lambda: leanfe(d, strategy="compress", **kw))
    res2, sec2 = timed(lambda: leanfe(d, strategy="compress", profile=True, **kw))
    fw, sec3 = timed(lambda: leanfe(d, strategy="alt_proj", demean_tol=1e-12, max_iter=200, **kw))
    b = np.array([res.coefs[c] for c in truth])
    fb = np.array([fw.coefs[c] for c in truth])
    s = np.array([res.std_errors[c] for c in truth])
    fs = np.array([fw.std_errors[c] for c in truth])
    out = dict(config="c4", n=n, seconds=sec2, rows_per_s=n / sec2, n_compressed=res.n_compressed, df_resid=res.df_resid,
               n_clusters=res.n_clusters, truth_max_z=truth_ok(res, truth), timings=res2.timings,
               compress_vs_altproj_beta=float(np.max(np.abs(b - fb) / np.abs(fb))),
               compress_vs_altproj_se=float(np.max(np.abs(s - fs) / fs)), altproj_seconds=sec3,
               altproj_iterations=fw.iterations)
    del d
    torch.cuda.empty_cache()
    small = max(n // 50, 50_000)
    ds, _ = panel_c4(small)
    r = leanfe(ds, strategy="compress", **kw)
    ref = O.leanfe_oracle({k_: v.cpu().numpy() for k_, v in ds.items()}, strategy="compress", **kw)
    eb = np.array(list(ref["coefs"].values()))
    es = np.array(list(ref["std_errors"].values()))
    out["oracle_sample"] = dict(n=small, beta_relerr=float(np.max(np.abs(np.array(list(r.coefs.values())) - eb) / np.abs(eb))),
                                se_relerr=float(np.max(np.abs(np.array(list(r.std_errors.values())) - es) / es)),
                                n_compressed=[r.n_compressed, ref["n_compressed"]])
    return out


def run_c5(scale):
    n = int(400_000_000 * scale)
    d, truth = panel_c5(n)
    xs = " + ".join(truth.keys())
    kw = dict(formula=f"y ~ {xs} | unit + time", vcov="iid", strategy="alt_proj")
    res, sec = timed(lambda: leanfe(d, **kw))
    res2, sec2 = timed(lambda: leanfe(d, profile=True, **kw))
    d2 = dict(d)
    d2["y"] = 2.0 * d["y"] + d["x1"]
    lin = leanfe(d2, **kw)
    lin_err = max(abs(lin.coefs[c] - (2 * res.coefs[c] + (1.0 if c == "x1" else 0.0))) / abs(lin.coefs[c]) for c in truth)
    out = dict(config="c5", n=n, seconds=sec2, rows_per_s=n / sec2, iterations=res.iterations, converged=res.converged,
               truth_max_z=truth_ok(res, truth), linearity_relerr=lin_err, timings=res2.timings,
               mem_gb=torch.cuda.max_memory_allocated() / 1e9)
    del d, d2
    torch.cuda.empty_cache()
    out["oracle_sample"] = oracle_sample(lambda m: panel_c5(m), kw, max(n // 200, 50_000))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="all")
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    get_engine(0)
    runs = {"c3": run_c3, "c4": run_c4, "c5": run_c5}
    for name, fn in runs.items():
        if a.which in ("all", name):
            print(json.dumps(fn(a.scale), default=str), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
