#!/usr/bin/env python
"""One fit of a named BASELINE shape on cuda:0, for ncu captures (launch lists and --set full):

    python profiles/profile_case.py c3 [--rows N] [--max-iter K] [--fits 1] [--vcov hc1]

Uses bench.py's counter-based generators, so the panel is the one the bench line reports."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from leanfe_b200 import leanfe  # noqa: E402
from leanfe_b200.engine import get_engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["c2", "c3", "c5"])
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--max-iter", type=int, default=50)
    ap.add_argument("--fits", type=int, default=1)
    ap.add_argument("--vcov", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = get_engine(0)
    if a.case == "c3":
        n = a.rows or 100_000_000
        d = bench.make_c3_rows(torch, dev, 0, n, n)
        kw = dict(formula="y ~ x1 + x2 | worker + firm + year", vcov=a.vcov or "hc1", strategy="alt_proj")
    elif a.case == "c5":
        n = a.rows or 400_000_000
        d = bench.make_c5_rows(torch, dev, 0, n, n)
        kw = dict(formula=bench.C5_FORMULA, vcov=a.vcov or "iid", strategy="alt_proj")
    else:
        n = a.rows or 50_000_000
        h, _ = bench.make_panel(n, 42)
        d = {k: torch.from_numpy(v).to(dev) for k, v in h.items()}
        kw = dict(formula=bench.FORMULA, vcov=a.vcov or "cluster", strategy="alt_proj")
    if kw["vcov"] == "cluster":
        kw["cluster_cols"] = ["unit"]
    torch.cuda.synchronize()
    for i in range(a.fits):
        t0 = time.perf_counter()
        r = leanfe(d, max_iter=a.max_iter, profile=True, **kw)
        torch.cuda.synchronize()
        print(json.dumps({"case": a.case, "rows": n, "fit": i, "seconds": time.perf_counter() - t0, "sweeps": r.iterations,
                          "timings": {k: round(v, 3) for k, v in r.timings.items() if v}}), flush=True)


if __name__ == "__main__":
    main()
