#!/usr/bin/env python
"""Per-SASS-instruction stall samples of an .ncu-rep (read here): the top instructions by stall samples, with the reasons
and shared-memory wavefront counts.   tools/ncu_sass_stalls.py rep [top] [lo hi]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    reasons = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data = []
    for n, r in enumerate(rows[2:]):
        if len(r) < len(h):
            continue
        try:
            samples = int(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        st = {k: int(r[ix[k]] or 0) for k in reasons}
        data.append((n, r[ix["Source"]].strip(), samples, int(r[ix["Instructions Executed"]] or 0), st, r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]]))
    tot = sum(d[2] for d in data)
    tots = {k: sum(d[4][k] for d in data) for k in reasons}
    print("total samples", tot, {k[6:]: v for k, v in sorted(tots.items(), key=lambda x: -x[1]) if v})
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3]), int(sys.argv[4])
        for d in data[lo:hi]:
            why = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(d[4].items(), key=lambda x: -x[1])[:3] if v)
            print(f"{d[0]:5d} {d[2]:6d} ex={d[3]:9d} wf={d[5]:>9s}/{d[6]:>9s} {d[1][:70]:70s} {why}")
        return
    for d in sorted(data, key=lambda x: -x[2])[:top]:
        why = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(d[4].items(), key=lambda x: -x[1])[:3] if v)
        print(f"{d[0]:5d} {100*d[2]/tot:5.1f}% ex={d[3]:9d} wf={d[5]:>9s}/{d[6]:>9s} {d[1][:70]:70s} {why}")


if __name__ == "__main__":
    main()
