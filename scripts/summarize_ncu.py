#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag>      # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep

Writes profiles/<tag>_launches.md (per-kernel share of a short bench run, from the
`--metrics gpu__time_duration.sum --clock-control none` pass) and profiles/<tag>_kernel.md (selected metrics
of the `--set full` capture of the dominant kernel).
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", x["Kernel Name"])
        name = re.sub(r"dm::<unnamed>::|void ", "", name)[:70]
        v = float(x["Metric Value"].replace(",", ""))
        if x.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = [f"# ncu launch list `{tag}` (per-kernel totals; serialised, cold-cache times: shares matter, not absolutes)", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
    out.append("")
    out.append(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    return "\n".join(out) + "\n"


def kernel(tag):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full capture `{tag}` (selected metrics per captured launch)", ""]
    for r in rows[2:]:
        out.append(f"## {r[hdr.index('Kernel Name')]}")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---:|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} | {units[i]} |")
        out.append("")
    return "\n".join(out) + "\n"


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for name, text in ((f"{tag}_launches.md", launches(tag)), (f"{tag}_kernel.md", kernel(tag))):
        if text:
            with open(os.path.join(ROOT, "profiles", name), "w") as f:
                f.write(text)
            print("wrote profiles/" + name)
    src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(src):
        import shutil
        shutil.copy(src, os.path.join(ROOT, "profiles", f"launches_{tag}.csv"))


if __name__ == "__main__":
    main()
