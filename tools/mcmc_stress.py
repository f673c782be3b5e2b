#!/usr/bin/env python
"""One-off stress: the fixed-seed chain test (tests/test_mcmc_trace.py) with more steps and several seeds."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import test_mcmc_trace as T  # noqa: E402

orig = np.random.PCG64
for seed in (1, 2, 3):
    np.random.PCG64 = lambda s=777, _seed=seed: orig(s + 1000 * _seed) if isinstance(s, int) and s == 777 else orig(s)   # only the chain's own stream
    try:
        T.test_fixed_seed_chain_matches_oracle_step_by_step("C2", dict(n_loci=6, n_sites=100), 700)
        T.test_fixed_seed_chain_matches_oracle_step_by_step("C3", dict(n_loci=4, n_sites=80), 400)
        print("seed", seed, "ok", flush=True)
    finally:
        np.random.PCG64 = orig
