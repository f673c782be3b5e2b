#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall reasons, opcode mix, hottest source lines."""
import collections
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    src_file = sys.argv[2] if len(sys.argv) > 2 else None
    h, units, vals = raw(rep)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]
    for v in vals:
        print("kernel:", v[h.index("Kernel Name")][:80])
        for k in keys:
            if k in h:
                print(f"  {k:80s} {v[h.index(k)]:>16s} {units[h.index(k)]}")
        st = [(float(v[i]), k) for i, k in enumerate(h) if "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct") and v[i]]
        for x, k in sorted(st, reverse=True)[:8]:
            print(f"  stall {k.split('stalled_')[1].split('_per_warp')[0]:28s} {x:6.1f}%")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    per, st, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    hist = collections.Counter()
    cur = -1
    for r in rows[3:]:
        ai = [i for i, x in enumerate(r) if x.startswith("0x7f")]
        ln = r[0].strip()
        if ln.isdigit() and (not ai or ai[0] > 1):
            cur = int(ln)
        if not ai or ai[0] + 5 >= len(r):
            continue
        a = ai[0]
        sass, sa, ex = r[a + 1], r[a + 2], r[a + 5]
        if not ex.isdigit():
            continue
        t = sass.split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        per[cur] += int(ex); st[cur] += int(sa) if sa.isdigit() else 0
        ops[cur][op] += int(ex); hist[op] += int(ex)
    tot = sum(per.values())
    print("opcode mix:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in hist.most_common(16)))
    if src_file:
        lines = open(src_file).read().split("\n")
        for ln, n in sorted(per.items(), key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
            top = ", ".join(f"{k}:{100*v/tot:.1f}" for k, v in ops[ln].most_common(4))
            print(f"{100*n/tot:5.1f}% st={st[ln]:5d} L{ln:4d} {lines[ln-1].strip()[:84]:84s} [{top}]")


if __name__ == "__main__":
    main()
