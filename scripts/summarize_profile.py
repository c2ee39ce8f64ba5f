#!/usr/bin/env python
"""Summarise gpurun_out artefacts into profiles/ (tracked): launch list shares and an ncu --set full extract."""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
note = sys.argv[2] if len(sys.argv) > 2 else ""


def launches():
    lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        v = float(r.get("Metric Value", "0").replace(",", ""))
        agg[r.get("Kernel Name", "?")][0] += 1
        agg[r.get("Kernel Name", "?")][1] += v
    tot = sum(v[1] for v in agg.values())
    with open("profiles/%s_launches.txt" % tag, "w") as f:
        f.write("# %s %s\n# ncu --metrics gpu__time_duration.sum --clock-control none; python bench.py --samples 4e8 --steps 1 --warmup 1\n"
                "# per-launch times are cold-cache and serialised: compare shares, not absolutes\n" % (tag, note))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%4d %10.1f us %5.1f%%\n" % (k[:72], n, t / 1e3, 100 * t / tot))


def full():
    out = subprocess.run(["ncu", "-i", "gpurun_out/prof_fast.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, vals = rd[0], rd[-1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
            "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled",
            "sass__inst_executed_local", "sm__throughput.avg.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit",
            "lts__t_bytes.sum ", "sm__pipe_fp64_cycles_active.avg.pct", "smsp__thread_inst_executed.sum"]
    with open("profiles/%s_slicer_ncu.txt" % tag, "w") as f:
        f.write("# %s %s\n# ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1; "
                "python bench.py --samples 5.3e9 --steps 1 --warmup 1 (the timed step: one launch of the streaming slicer over five slabs)\n" % (tag, note))
        for h, v in zip(hdr, vals):
            if any(w in h for w in want) and "pcsamp" not in h:
                f.write("%s = %s\n" % (h, v))


if __name__ == "__main__":
    for fn in (launches, full):
        try:
            fn()
        except Exception as e:
            print(fn.__name__, "failed:", e)
    try:
        b = json.load(open("gpurun_out/bench.json"))
        json.dump(b, open("profiles/%s_bench.json" % tag, "w"), indent=1)
    except Exception as e:
        print("bench copy failed", e)
