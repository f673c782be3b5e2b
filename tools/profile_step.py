#!/usr/bin/env python
"""Small driver for ncu: builds a workload, warms up, runs a few full-evaluation steps (all nodes dirty)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--loci", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--incremental", action="store_true", help="regime B: single-locus moves instead of full evaluations")
    a = ap.parse_args()
    from starbeast2_b200 import synth
    from starbeast2_b200.engine import Engine
    n = a.loci or synth.CONFIGS[a.workload]["n_loci"]
    pb = synth.make_problem(a.workload, n_loci=n)
    e = Engine.from_problem(pb, exact_order=a.exact)
    tot = np.zeros(3)
    rng = np.random.default_rng(0)
    for i in range(a.warmup + a.steps):
        if a.incremental and i > 0:
            l = int(rng.integers(0, pb.n_loci))
            t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
            e.store()
            e.set_gene_tree(l, t)
        else:
            e.mark_all_dirty()
        e.eval_posterior_total(tot)
    print("logP", tot, "units", pb.units, "launches", e.launch_count())


if __name__ == "__main__":
    main()
