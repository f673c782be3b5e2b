#!/usr/bin/env python
"""Writes tests/golden/*.json.

The reference is Java and cannot run in this image (no JVM), so its outputs cannot be regenerated here.  Two kinds of
fixtures are committed instead:
  reference_known_answers.json -- the inputs and expected values of the reference's own JUnit tests
      (/root/reference/src/sb2tests/*.java; file:line recorded per case), plus the full-precision values the oracle
      reproduces for them;
  oracle_cases.json -- small seeded posterior states with per-locus values computed by the CPU oracle, so the GPU parity
      tests also have committed numbers to hit that do not depend on rebuilding the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
import ref_fixtures as rf  # noqa: E402
from starbeast2_b200 import synth  # noqa: E402


def main():
    known = {
        "source": "genomescale/starbeast2 v1.0.0, src/sb2tests/*.java (tolerance 10e-6, ConstantPopulationTest.java:34)",
        "cases": {
            "ConstantPopulationTest": {"at": "ConstantPopulationTest.java:41,46-48", "species": rf.SPECIES_4, "genes": rf.GENES_4,
                                       "popSize": 0.3, "ploidy": 2.0, "expected": rf.EXPECTED["constant"], "oracle": 2.0784641097740977},
            "LinearWithConstantRootTest": {"at": "LinearWithConstantRootTest.java:42,48-50,68", "species": rf.SPECIES_4, "genes": rf.GENES_4,
                                           "tip": 0.3, "top": 0.15, "ploidy": 2.0, "expected": rf.EXPECTED["lwcr"], "oracle": 2.887976975975221},
            "ConstantIOTest": {"at": "ConstantIOTest.java:37-39,45,54-56", "species": rf.SPECIES_4, "genes": rf.GENES_4, "alpha": 1.5, "beta": 2.5,
                               "expected": rf.EXPECTED["analytic"], "oracle": -14.523398476213451},
            "MissingDataConstantIOTest": {"at": "MissingDataConstantIOTest.java:45,54-56", "species": rf.SPECIES_4, "genes": rf.GENES_4_MISSING,
                                          "alpha": 1.5, "beta": 2.5, "expected": rf.EXPECTED["analytic_missing"], "oracle": -10.956285249389675},
            "IncompatibleTreeTest": {"at": "IncompatibleTreeTest.java:41,46-48", "species": rf.SPECIES_8, "genes": rf.GENES_8, "expected": "-Infinity"},
            "StarbeastClockTest": {"at": "StarbeastClockTest.java:49-50,98-110,133-143", "species": rf.SPECIES_3, "gene": rf.GENE_CLOCK,
                                   "speciesBins": {",".join(k): v for k, v in rf.CLOCK_SPECIES_BINS.items()},
                                   "expectedRates": {",".join(k): v for k, v in rf.CLOCK_EXPECTED.items()}},
            "UncorrelatedRatesTest": {"at": "UncorrelatedRatesTest.java:19,70-79,107-117", "tree": rf.GENE_CLOCK,
                                      "bins": {",".join(k): v for k, v in rf.UCLN_BINS.items()},
                                      "expectedRates": {",".join(k): v for k, v in rf.UCLN_EXPECTED.items()}},
        },
    }
    json.dump(known, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)
    cases = {}
    for name, kw in (("C2", dict(n_loci=4, n_sites=200)), ("C3", dict(n_loci=3, n_sites=150)), ("C4", dict(n_loci=5, n_sites=100)),
                     ("C5", dict(n_loci=2, n_sites=100))):
        pb = synth.make_problem(name, missing_frac=0.02, **kw)
        r = oracle.posterior(pb)
        cases[name] = {"make_problem": dict(name=name, missing_frac=0.02, **kw), "per_locus_coal": r.per_locus_coal.tolist(),
                       "per_locus_lik": r.per_locus_lik.tolist(), "per_branch": r.per_branch.tolist(), "coalescent": r.coalescent,
                       "likelihood": r.likelihood, "total": r.total}
        # operator-side walks (sb2_connecting_nodes / sb2_max_height): every internal species node; one canonical order
        sp = pb.species
        q = {}
        for node in range(sp.n_leaves, sp.n_nodes):
            mask, tip, root, cnt = oracle.connecting_nodes(pb, node)
            q[str(node)] = {"connecting": np.flatnonzero(mask).tolist(), "tipward": None if np.isinf(tip) else tip,
                            "rootward": None if np.isinf(root) else root, "count": cnt}
        cmap = np.zeros(sp.n_nodes, np.int32)
        cmap[: sp.n_leaves] = np.arange(sp.n_leaves)                  # leaves in number order; centre after the first half
        mh = oracle.max_height(pb, cmap, sp.n_leaves // 2)
        cases[name]["operator_queries"] = {"connecting_nodes": q, "max_height_center": sp.n_leaves // 2, "max_height": None if np.isinf(mh) else mh}
    json.dump({"generator": "tests/golden/make_golden.py (oracle/sb2_oracle.c, gcc -O2 -ffp-contract=off)", "cases": cases},
              open(os.path.join(HERE, "oracle_cases.json"), "w"), indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
