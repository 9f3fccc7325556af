#!/usr/bin/env python
"""Turn ncu exports brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv profiles/r01_launches_c2.md [first-K-launches-per-kernel]
  python tools/profile_summary.py kernel   gpurun_out/prof_fwd9.ncu-rep profiles/r01_ring_vit_forward_ws_ncu.md
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, first=0):
    """first > 0: only the first `first` launches of every kernel (bench.py runs its device-resident steps first;
    the later launches of the same kernels belong to the pipelined host-pointer decode, which works on segments)."""
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = agg.setdefault(row["Kernel Name"], [])
        if first <= 0 or len(v) < first:
            v.append(float(row["Metric Value"].replace(",", "")))
    groups = collections.OrderedDict([
        ("decode step (config 2, device-resident): one launch of each per step", lambda k: "ring_vit_" in k),
        ("Baum-Welch iteration (config 3): one launch of each per iteration", lambda k: "::em_" in k),
        ("not part of a step (peak probes of hmm_measure_peaks, the pipelined host-pointer decode's helpers)", lambda k: True)])
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: `{src}`\n")
        left = collections.OrderedDict(agg)
        for title, pred in groups.items():
            mine = collections.OrderedDict((k, v) for k, v in left.items() if pred(k))
            for k in mine:
                del left[k]
            if not mine:
                continue
            tot = sum(sum(v) / len(v) for v in mine.values())
            f.write(f"\n## {title}\n\n| kernel | launches | mean us | share |\n|---|---|---|---|\n")
            for k, v in mine.items():
                m = sum(v) / len(v)
                f.write(f"| `{k[:90]}` | {len(v)} | {m / 1e3:.1f} | {100 * m / tot:.1f}% |\n")
            f.write(f"\nsum of per-kernel means: {tot / 1e3:.1f} us\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled"]


def kernel(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\nsource: `{rep}` (kept in gpurun_out/, not tracked)\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"\n## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---|---|\n")
            for h, u, v in zip(hdr, units, vals):
                if any(h == w or (w.endswith("stalled") and w in h and "per_issue_active" in h) for w in WANT):
                    f.write(f"| {h} | {v} | {u} |\n")
            try:
                tr = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))
                f.write(f"\ntraffic (dram read + write) = {tr:.1f} {units[hdr.index('dram__bytes_read.sum')]}\n")
            except Exception:
                pass


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        kernel(sys.argv[2], sys.argv[3])
