#!/usr/bin/env python
"""Occupancy timeline of one k_treewalk launch from per-tile %globaltimer stamps (SB2_DEBUG_WALK_TIMES=file):
when CTAs start, wait for k_prepare, walk and end; how many walks are in flight over time."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    wl, n = sys.argv[1], int(sys.argv[2])
    path = os.environ.get("SB2_DEBUG_WALK_TIMES", "/tmp/walk_times.bin")
    if "SB2_DEBUG_WALK_TIMES" not in os.environ:
        env = dict(os.environ, SB2_DEBUG_WALK_TIMES=path)
        subprocess.check_call([sys.executable, os.path.join(os.path.dirname(__file__), "profile_step.py"), "--workload", wl, "--loci", str(n), "--steps", "3", "--warmup", "2"], env=env)
    t = np.fromfile(path, dtype=np.uint64).reshape(-1, 8).astype(np.float64)
    t0 = t[:, 0].min()
    us = (t[:, :6] - t0) * 1e-3
    names = ["start", "staged", "dep met", "walk begins", "walk ends", "CTA ends"]
    print(f"{wl} x {n}: {len(us)} tiles; launch spans {us[:, 5].max():.1f} us")
    for i, nm in enumerate(names):
        c = us[:, i]
        print(f"  {nm:12s} min {c.min():7.1f}  p10 {np.percentile(c, 10):7.1f}  median {np.median(c):7.1f}  p90 {np.percentile(c, 90):7.1f}  max {c.max():7.1f}")
    dur = us[:, 4] - us[:, 3]
    print(f"  walk duration: min {dur.min():.1f} median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f} us; epilogue median {np.median(us[:, 5] - us[:, 4]):.1f} us; "
          f"prologue (start -> walk begins) median {np.median(us[:, 3] - us[:, 0]):.1f} us")
    # who is slow?  by SM (the memory system does not serve every SM alike) and by tile (the last tile of a locus is partial)
    sm = t[:, 6].astype(int)
    first = us[:, 3] < 5.0   # first-wave walks: all start together
    if first.sum() > 100:
        per_sm = {k: dur[first & (sm == k)].mean() for k in np.unique(sm[first])}
        v = np.array(list(per_sm.values()))
        order = sorted(per_sm, key=per_sm.get)
        print(f"  first-wave walk duration by SM: min {v.min():.1f} median {np.median(v):.1f} max {v.max():.1f} us; fastest SMs {order[:8]}, slowest {order[-8:]}")
        within = np.mean([dur[first & (sm == k)].std() for k in np.unique(sm[first]) if (first & (sm == k)).sum() > 1])
        print(f"  spread: between SM means {v.std():.1f} us, within an SM {within:.1f} us")
    end = us[:, 5].max()
    edges = np.linspace(0, end, 21)
    print("  walks in flight (start of each 5 % slice of the launch):", [int(((us[:, 3] <= x) & (us[:, 4] > x)).sum()) for x in edges[:-1]])


if __name__ == "__main__":
    main()
