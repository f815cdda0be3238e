#!/usr/bin/env python
"""Turns an .ncu-rep brought back from gpurun into the text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_r01_k16.ncu-rep > profiles/r01_integrate_k16.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed_op_branch.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]} (per launch, cold-cache, serialised)")
    for r in rows[2:]:
        print(f"\n## launch {r[0]}: {r[hdr.index('Kernel Name')][:110]}")
        for k in KEYS:
            if k in hdr:
                print(f"{k:72s} {r[hdr.index(k)]:>22s} {units[hdr.index(k)]}")
        st = [(h, float(r[i])) for i, h in enumerate(hdr)
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i]]
        tot = sum(v for _, v in st) or 1.0
        print("warp stall samples (share of all samples):")
        for h, v in sorted(st, key=lambda t: -t[1])[:8]:
            print(f"    {h.split('stalled_')[1]:28s} {100 * v / tot:5.1f} %")


def launches(path):
    """Share of a step per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
    import collections
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        name = r[kn].split("(")[0].split("::")[-1].split("<")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    total = sum(v for _, v in agg.values()) or 1.0
    print(f"# {path}: per-launch durations under ncu (cold-cache, serialised); share of the summed kernel time")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:6d} launches {v / 1e3:12.3f} ms {100 * v / total:7.2f} %")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        main(sys.argv[1])
