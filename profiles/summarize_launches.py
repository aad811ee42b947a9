#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--fits 2] [--title "..."] > profiles/x.md

`--fits` = number of fits captured in the list (per-fit numbers are the totals divided by it).
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--fits", type=int, default=1)
    ap.add_argument("--title", default="ncu launch list, aggregated per kernel")
    ap.add_argument("--cmd", default="")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        rows.append((name, ms))
    agg = collections.OrderedDict()
    for name, ms in rows:
        t, c = agg.get(name, (0.0, 0))
        agg[name] = (t + ms, c + 1)
    total = sum(t for t, _ in agg.values())
    print(f"# {a.title}\n")
    if a.cmd:
        print(f"`{a.cmd}`\n")
    print(f"{len(rows)} launches captured = {a.fits} fit(s); per-fit = totals / {a.fits}.  Times are cold-cache and "
          "serialised under ncu: compare shares, not absolutes.\n")
    print("| kernel | ms per fit | share | launches per fit |\n|---|---|---|---|")
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| `{name}` | {t / a.fits:.3f} | {100 * t / total:.1f}% | {c / a.fits:g} |")
    print(f"| total | {total / a.fits:.2f} | 100% | {len(rows) / a.fits:g} |")


if __name__ == "__main__":
    main()
