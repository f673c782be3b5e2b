#!/usr/bin/env python
"""Host-side time of every C-ABI call of one regime-B MCMC step (store, set_gene_tree, eval_coalescent, eval_likelihood, restore)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from starbeast2_b200 import synth  # noqa: E402
from starbeast2_b200.engine import Engine  # noqa: E402


def main():
    name, n_loci = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("C2", 50)
    pb = synth.make_problem(name, n_loci=n_loci)
    e = Engine.from_problem(pb)
    e.eval_posterior()
    rng = np.random.default_rng(12345)
    moves = []
    for _ in range(64):
        l = int(rng.integers(0, pb.n_loci))
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        moves.append((l, t))
    acc = np.zeros(5)
    n = 600
    for i in range(n + 50):
        l, t = moves[i % len(moves)]
        t0 = time.perf_counter(); e.store()
        t1 = time.perf_counter(); e.set_gene_tree(l, t)
        t2 = time.perf_counter(); e.eval_coalescent()
        t3 = time.perf_counter(); e.eval_likelihood()
        t4 = time.perf_counter(); e.restore()
        t5 = time.perf_counter()
        if i >= 50:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
    us = 1e6 * acc / n
    print(name, n_loci, dict(zip(["store", "set_gene_tree", "eval_coalescent", "eval_likelihood", "restore"], np.round(us, 2))), "total", round(us.sum(), 2))


if __name__ == "__main__":
    main()
