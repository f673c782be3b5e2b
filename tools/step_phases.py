#!/usr/bin/env python
"""Where one fused step (k_step) spends its time: %globaltimer stamps of the first cluster's leader at the phase boundaries
(sb2_step_phase_times), the CUDA-event time of the launch, and the host's wall time per call -- regime A (everything dirty)
and regime B (one gene-tree node moved), at a named shape."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from starbeast2_b200 import synth  # noqa: E402
from starbeast2_b200.engine import Engine  # noqa: E402

NAMES = ["start", "tree", "walk", "gather", "to-red", "publish", "red-local", "finish", "loaded", "embed", "coal", "dirty", "sched", "mats"]


def phases(e):
    us = (C.c_double * 14)()
    if e._L.sb2_step_phase_times(e._h, us, 14) != 0:
        return np.zeros(14)   # this evaluation took the three-kernel path
    return np.array(list(us))


def main():
    name, n_loci = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("C2", 50)
    pb = synth.make_problem(name, n_loci=n_loci)
    e = Engine.from_problem(pb, profile=True)
    e.eval_posterior()
    tot = np.zeros(3)
    acc, wall, kern = np.zeros(14), 0.0, 0.0
    n = 200
    for i in range(n + 20):
        e.mark_all_dirty()
        t0 = time.perf_counter()
        e.eval_posterior_total(tot)
        t1 = time.perf_counter()
        ms, _ = e.kernel_time()
        if i >= 20:
            acc += phases(e); wall += t1 - t0; kern += ms
    print(name, n_loci, "regime A:", dict(zip(NAMES, np.round(acc / n, 2))), 
          "event us", round(1e3 * kern / n, 2), "host wall us", round(1e6 * wall / n, 2))
    rng = np.random.default_rng(12345)
    moves = []
    for _ in range(64):
        l = int(rng.integers(0, pb.n_loci))
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        moves.append((l, t))
    acc, wall, kern = np.zeros(14), 0.0, 0.0
    for i in range(n + 20):
        l, t = moves[i % len(moves)]
        e.store()
        e.set_gene_tree(l, t)
        t0 = time.perf_counter()
        e.eval_posterior_total(tot)
        t1 = time.perf_counter()
        e.restore()
        ms, _ = e.kernel_time()
        if i >= 20:
            acc += phases(e); wall += t1 - t0; kern += ms
    print(name, n_loci, "regime B:", dict(zip(NAMES, np.round(acc / n, 2))), 
          "event us", round(1e3 * kern / n, 2), "host wall us (eval call)", round(1e6 * wall / n, 2))


if __name__ == "__main__":
    main()
