#!/usr/bin/env python
"""Summarise an `ncu --set full` report per launch (run where ncu is installed):

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--peak 6547.2] > profiles/x.md
"""
import argparse
import csv
import io
import re
import subprocess


def f(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return float("nan")


def to_base(v, unit, kind):
    """kind 'time' -> ms, 'bytes' -> GB"""
    unit = (unit or "").strip().lower()
    if kind == "time":
        return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3,
                    "nsecond": 1e-6}.get(unit, 1.0)
    return v * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0, "tbyte": 1e3}.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--peak", type=float, default=6547.2, help="measured HBM GB/s (MEASURED_PEAKS.json)")
    ap.add_argument("--title", default="ncu --set full, per launch")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h
             and "not_issued" not in h and "selected" not in h]

    def g(r, k):
        return f(r[ix[k]]) if k in ix else float("nan")

    print(f"# {a.title}\n")
    print("Per launch, cold-cache, serialised (ncu replays each kernel).  HBM GB/s = (dram read + write) / time; "
          f"peak = {a.peak:.0f} GB/s (MEASURED_PEAKS.json).  l1tex = l1tex__data_pipe_lsu_wavefronts % of peak "
          "(the shared-memory / gather wavefront pipe), lts = L2 throughput %.\n")
    print("| kernel | ms | DRAM rd GB | DRAM wr GB | HBM GB/s | % HBM peak | l1tex % | lts % | issue % | warps % | regs | "
          "grid x block | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").strip()
        ms = to_base(g(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]], "time")
        rd = to_base(g(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]], "bytes")
        wr = to_base(g(r, "dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]], "bytes")
        key = (name, round(ms, 2), r[ix["Grid Size"]])
        if key in seen:
            seen[key] += 1
            continue
        seen[key] = 1
        st = sorted(((g(r, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)[:3]
        sts = ", ".join(f"{n} {v:.1f}" for v, n in st)
        gbs = (rd + wr) / (ms / 1e3) if ms > 0 else float("nan")
        grid = re.sub(r"[^0-9]", " ", r[ix["Grid Size"]]).split()[0]
        blk = re.sub(r"[^0-9]", " ", r[ix["Block Size"]]).split()[0]
        print(f"| `{name}` | {ms:.3f} | {rd:.3f} | {wr:.3f} | {gbs:.0f} | {100 * gbs / a.peak:.1f} | "
              f"{g(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{g(r, 'launch__registers_per_thread'):.0f} | {grid} x {blk} | {sts} |")
    print("\n(launches of the same kernel with the same grid and duration to 0.01 ms are listed once.)")


if __name__ == "__main__":
    main()
