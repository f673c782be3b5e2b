"""Synthetic StarBEAST2 states of the BASELINE.json shapes (recipe: SURVEY.md 8d).

Species tree: Yule (lambda=1) rescaled to root height 1.  Gene trees: multispecies-coalescent
simulation in the manner of starbeast2.utils.SimulatedGeneTree.getSimulatedGeneTree
(/root/reference/src/starbeast2/utils/SimulatedGeneTree.java:57-152: exponential waiting times
with propensity k(k-1)/2 inside each species branch, lineages merged upward at speciation
times), with time scaled by N_b * ploidy.  Sequences are evolved down each gene tree under the
locus's substitution model and compressed to unique site patterns + weights.
Everything is seeded: numpy.random.Generator(PCG64(20260100 + config#)).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .problem import (CLOCK_SPECIES_RATES, CLOCK_STRICT, POP_ANALYTIC, POP_CONSTANT, POP_LWCR, SUBST_EIGEN, MorphPartition, lewis_mk_eigen,
                      SUBST_HKY, Locus, Problem, TreeArrays, gtr_eigen_params, hky_params)


def yule_species_tree(rng, n_species: int, n_fossils: int = 0, sa_fraction: float = 0.25) -> TreeArrays:
    """Leaves 0..n-1, internal nodes numbered by increasing height, root last.
    n_fossils extra dated leaves (C4): heights U(0.05, 0.8); a fraction attach as sampled ancestors
    (zero-length branch below a 'fake' internal node, as the SA package represents them)."""
    n = n_species
    heights = [0.0] * n
    left: List[int] = [-1] * n
    right: List[int] = [-1] * n
    active = list(range(n))
    t = 0.0
    while len(active) > 1:
        k = len(active)
        t += rng.exponential(1.0 / k)
        i, j = rng.choice(k, size=2, replace=False)
        a, b = active[i], active[j]
        nr = len(heights)
        heights.append(t)
        left.append(a)
        right.append(b)
        active = [x for x in active if x not in (a, b)] + [nr]
    h = np.array(heights) / heights[-1]
    if n_fossils:
        return _add_fossils(rng, h, left, right, n, n_fossils, sa_fraction)
    return _finish(h, left, right, n)


def _finish(h, left, right, n_leaves) -> TreeArrays:
    G = len(h)
    parent = np.full(G, -1, np.int32)
    for i in range(G):
        if left[i] >= 0:
            parent[left[i]] = i
            parent[right[i]] = i
    return TreeArrays(parent, np.array(left, np.int32), np.array(right, np.int32), np.array(h, dtype=np.float64))


def _add_fossils(rng, h, left, right, n_extant, n_fossils, sa_fraction) -> TreeArrays:
    """Graft fossil leaves onto the extant tree, then renumber: leaves (extant, then fossils) first,
    internal nodes by height with the root last."""
    nodes = [{"h": float(h[i]), "l": left[i], "r": right[i], "leaf": left[i] < 0, "fossil": False} for i in range(len(h))]
    root = len(nodes) - 1
    par = {}
    for i, nd in enumerate(nodes):
        if not nd["leaf"]:
            par[nd["l"]] = i
            par[nd["r"]] = i
    for _ in range(n_fossils):
        fh = rng.uniform(0.05, 0.8)
        # branches spanning fh
        cands = [i for i in range(len(nodes)) if i != root and nodes[i]["h"] <= fh < nodes[par[i]]["h"]]
        b = int(rng.choice(cands))
        is_sa = rng.uniform() < sa_fraction
        p = par[b]
        ah = fh if is_sa else rng.uniform(fh, nodes[p]["h"])  # attachment (fake node) height
        fossil = len(nodes)
        nodes.append({"h": fh, "l": -1, "r": -1, "leaf": True, "fossil": True})
        fake = len(nodes)
        nodes.append({"h": ah, "l": b, "r": fossil, "leaf": False, "fossil": False})
        if nodes[p]["l"] == b:
            nodes[p]["l"] = fake
        else:
            nodes[p]["r"] = fake
        par[b] = fake
        par[fossil] = fake
        par[fake] = p
    leaves = [i for i, nd in enumerate(nodes) if nd["leaf"] and not nd["fossil"]] + \
             [i for i, nd in enumerate(nodes) if nd["fossil"]]
    internal = sorted([i for i, nd in enumerate(nodes) if not nd["leaf"] and i != root], key=lambda i: nodes[i]["h"]) + [root]
    order = leaves + internal
    new = {old: k for k, old in enumerate(order)}
    hh = [nodes[o]["h"] for o in order]
    ll = [new[nodes[o]["l"]] if nodes[o]["l"] >= 0 else -1 for o in order]
    rr = [new[nodes[o]["r"]] if nodes[o]["r"] >= 0 else -1 for o in order]
    return _finish(np.array(hh), ll, rr, len(leaves))


def simulate_gene_tree(rng, sp: TreeArrays, pop_scale: np.ndarray, tips_per_leaf: np.ndarray) -> Tuple[TreeArrays, np.ndarray]:
    """MSC simulation.  pop_scale[b] = N_b * ploidy (coalescent rate per pair = 1/pop_scale[b]).
    Gene leaves are numbered species leaf by species leaf; returns (tree, tip_species)."""
    S = sp.n_nodes
    tip_species = np.repeat(np.arange(len(tips_per_leaf)), tips_per_leaf).astype(np.int32)
    n_tips = len(tip_species)
    G = 2 * n_tips - 1
    parent = np.full(G, -1, np.int32)
    left = np.full(G, -1, np.int32)
    right = np.full(G, -1, np.int32)
    height = np.zeros(G)
    nxt = n_tips
    lineages = {b: [] for b in range(S)}
    for g, s in enumerate(tip_species):
        lineages[int(s)].append(g)
    order = np.argsort(sp.height, kind="stable")
    # process species nodes bottom-up: coalesce inside the branch above each node
    for b in order:
        b = int(b)
        if sp.left[b] >= 0:
            lineages[b] = lineages[int(sp.left[b])] + lineages[int(sp.right[b])]
        t = sp.height[b]
        top = sp.height[sp.parent[b]] if sp.parent[b] >= 0 else np.inf
        cur = lineages[b]
        while len(cur) > 1:
            k = len(cur)
            t += rng.exponential(pop_scale[b] / (0.5 * k * (k - 1)))
            if t >= top:
                break
            i, j = rng.choice(k, size=2, replace=False)
            a, c = cur[i], cur[j]
            height[nxt] = t
            left[nxt], right[nxt] = a, c
            parent[a] = parent[c] = nxt
            cur = [x for x in cur if x not in (a, c)] + [nxt]
            nxt += 1
        lineages[b] = cur
    assert nxt == G
    return TreeArrays(parent, left, right, height), tip_species


def discrete_gamma_rates(shape: float, n_cat: int) -> np.ndarray:
    """Equal-probability Gamma(shape, mean 1) categories, mean-of-category rates normalised to mean 1.
    (BEAST's SiteModel uses a chi-square-quantile approximation, quirk Q5: the plugin takes the
    category rates from the host object; this generator only needs *some* valid rates.)"""
    from scipy.special import gammainc
    from scipy.stats import gamma
    if n_cat == 1:
        return np.ones(1)
    cuts = gamma.ppf(np.arange(1, n_cat) / n_cat, shape, scale=1.0 / shape)
    edges = np.concatenate([[0.0], cuts, [np.inf]])
    cdf1 = gammainc(shape + 1.0, edges * shape)
    r = (cdf1[1:] - cdf1[:-1]) * n_cat
    return r / r.mean()


def _transition_matrix(kind, sp, d):
    if kind == SUBST_HKY:
        kappa, pi = sp[0], sp[1:5]
        q = np.array([[0, 1, kappa, 1], [1, 0, 1, kappa], [kappa, 1, 0, 1], [1, kappa, 1, 0]], dtype=float) * pi[None, :]
        np.fill_diagonal(q, -q.sum(1))
        q /= -(pi * np.diag(q)).sum()
        w, v = np.linalg.eig(q)
        return np.real((v * np.exp(w * d)) @ np.linalg.inv(v))
    ev, V, Vi = sp[0:4], sp[4:20].reshape(4, 4), sp[20:36].reshape(4, 4)
    return np.abs((V * np.exp(ev * d)) @ Vi)


def simulate_alignment(rng, tree: TreeArrays, n_tips: int, n_sites: int, kind: int, sparams: np.ndarray,
                       cat_rates: np.ndarray, branch_rates: np.ndarray, missing_frac: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """Evolve n_sites down the tree; returns (unique patterns uint8 [n_tips, P], weights int32 [P])."""
    G = tree.n_nodes
    pi = sparams[1:5] if kind == SUBST_HKY else sparams[36:40]
    cats = rng.integers(0, len(cat_rates), n_sites)
    states = np.zeros((G, n_sites), np.uint8)
    root = tree.root
    states[root] = rng.choice(4, size=n_sites, p=pi / pi.sum())
    stack = [root]
    while stack:
        n = stack.pop()
        for c in (tree.left[n], tree.right[n]):
            if c < 0:
                continue
            c = int(c)
            blen = (tree.height[n] - tree.height[c]) * branch_rates[c]
            for ci, cr in enumerate(cat_rates):
                sel = np.nonzero(cats == ci)[0]
                if sel.size == 0:
                    continue
                P = _transition_matrix(kind, sparams, blen * cr)
                P = np.clip(P, 0, None)
                cum = np.cumsum(P / P.sum(1, keepdims=True), axis=1)
                u = rng.uniform(size=sel.size)
                states[c, sel] = (u[:, None] > cum[states[n, sel]]).sum(1).clip(0, 3)
            stack.append(c)
    tips = states[:n_tips].copy()
    if missing_frac > 0:
        tips[rng.uniform(size=tips.shape) < missing_frac] = 4
    pats, weights = np.unique(tips, axis=1, return_counts=True)
    return np.ascontiguousarray(pats, np.uint8), weights.astype(np.int32)


def simulate_morphology(rng, sp: TreeArrays, K: int, n_chars: int, clock_rate: float, missing_frac: float = 0.05):
    """K-state characters evolved down the SPECIES tree under Lewis's Mk, kept only if variable (Mkv ascertainment), compressed
    to patterns, with the K constant dummy patterns of an `ascertained="true"` alignment appended at weight 0 (the
    excludefrom / excludeto columns of tutorial/CanisFBD/Canis-FBD.xml:969-971).  Returns (patterns, weights, excluded)."""
    n_leaves, G = sp.n_leaves, sp.n_nodes

    def mk(t):
        e = np.exp(-K * t / (K - 1.0))
        p = np.full((K, K), 1.0 / K - e / K)
        np.fill_diagonal(p, 1.0 / K + (K - 1.0) * e / K)
        return p

    cols = []
    while len(cols) < n_chars:
        st = np.zeros(G, np.int64)
        st[sp.root] = rng.integers(0, K)
        stack = [sp.root]
        while stack:
            n = stack.pop()
            for c in (sp.left[n], sp.right[n]):
                if c < 0:
                    continue
                c = int(c)
                st[c] = rng.choice(K, p=mk((sp.height[n] - sp.height[c]) * clock_rate)[st[n]])
                stack.append(c)
        col = st[:n_leaves].astype(np.uint8)
        if len(np.unique(col)) > 1:
            cols.append(col)
    tips = np.stack(cols, axis=1)
    tips[rng.uniform(size=tips.shape) < missing_frac] = K    # '?'
    pats, weights = np.unique(tips, axis=1, return_counts=True)
    dummies = np.stack([np.full(n_leaves, s, np.uint8) for s in range(K)], axis=1)
    excluded = np.arange(pats.shape[1], pats.shape[1] + K, dtype=np.int32)
    pats = np.concatenate([pats, dummies], axis=1)
    weights = np.concatenate([weights, np.zeros(K, weights.dtype)])
    return np.ascontiguousarray(pats, np.uint8), weights.astype(np.int32), excluded


CONFIGS = {
    # name: loci, sites, species, tips/species, subst, n_cat, clock, pop model
    "C2": dict(n_loci=50, n_sites=1000, n_species=10, tips_per=3, subst="hky", n_cat=1, clock="strict", pop="constant", seed=2),
    "C3": dict(n_loci=200, n_sites=800, n_species=20, tips_per=3, subst="gtr", n_cat=4, clock="ucln", pop="lwcr", seed=3),
    # CanisFBD-style (tutorial/CanisFBD/Canis-FBD.xml): fossil tips in the species tree, analytic population sizes, and -- like the
    # reference's four morphTreeLikelihoods (:964-1033) -- four ascertained LewisMK partitions (K = 2..5) on the species tree
    "C4": dict(n_loci=500, n_sites=600, n_species=8, tips_per=2, subst="hky", n_cat=1, clock="strict", pop="analytic", seed=4,
               n_fossils=16, n_morph=4, n_chars=24),
    "C5": dict(n_loci=2000, n_sites=500, n_species=40, tips_per=3, subst="gtr", n_cat=4, clock="ucln", pop="constant", seed=5),
}


def morph_owner_locus(mi: int, n_loci: int, n_morph: int) -> int:
    """Sharded runs: morphological partition mi travels with the shard that holds this locus index."""
    return ((2 * mi + 1) * n_loci) // (2 * n_morph)


def make_problem(name: str = "C2", n_loci: Optional[int] = None, n_sites: Optional[int] = None,
                 seed_offset: int = 0, missing_frac: float = 0.0, locus_range: Optional[Tuple[int, int]] = None,
                 **override) -> Problem:
    """locus_range=(lo, hi): generate only loci lo..hi-1 of the n_loci-locus problem (species tree and shared
    parameters come from the config seed, each locus from its own (seed, index) stream, so every rank of a
    multi-GPU run can build just its shard of one consistent global problem)."""
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    if n_loci is not None:
        cfg["n_loci"] = n_loci
    if n_sites is not None:
        cfg["n_sites"] = n_sites
    base_seed = 20260100 + cfg["seed"] + 1000 * seed_offset
    rng = np.random.Generator(np.random.PCG64(base_seed))
    n_sp = cfg["n_species"]
    sp = yule_species_tree(rng, n_sp, cfg.get("n_fossils", 0))
    S = sp.n_nodes
    n_sp_leaves = sp.n_leaves
    tips_per_leaf = np.array([cfg["tips_per"]] * n_sp + [0] * (n_sp_leaves - n_sp))
    ploidy = 2.0
    pop_n = rng.gamma(2.0, 0.05 / 2.0, size=S)

    # species-tree clock
    if cfg["clock"] == "ucln":
        from scipy.special import ndtri
        sigma = 0.3
        n_bins = S - 1
        bins = np.exp(-(0.5 * sigma * sigma) + sigma * ndtri((np.arange(n_bins) + 0.5) / n_bins))
        idx = rng.integers(0, n_bins, size=S - 1)
        srates = np.ones(S)
        srates[: S - 1] = bins[idx]
        # root keeps the mean (estimateRoot=false, UncorrelatedRates.java:161); root is node S-1
    else:
        srates = np.ones(S)

    # population model parameters
    if cfg["pop"] == "constant":
        pop_kind, pop_a, pop_b, alpha, beta = POP_CONSTANT, pop_n, None, 0.0, 0.0
        scale_of = lambda b: pop_n[b] * ploidy
    elif cfg["pop"] == "lwcr":
        pop_kind = POP_LWCR
        top = pop_n[: S - 1] * 0.5
        tip = pop_n[:n_sp_leaves].copy()
        pop_a, pop_b, alpha, beta = tip, top, 0.0, 0.0
        scale_of = lambda b: pop_n[b] * ploidy
    else:
        pop_kind, pop_a, pop_b = POP_ANALYTIC, None, None
        alpha = 3.0
        beta = 0.05 * (alpha - 1.0)
        scale_of = lambda b: pop_n[b] * ploidy
    pop_scale = np.array([scale_of(b) for b in range(S)])

    pi_hky = np.array([0.30, 0.20, 0.20, 0.30])
    loci: List[Locus] = []
    lo, hi = locus_range if locus_range is not None else (0, cfg["n_loci"])
    for li in range(lo, hi):
        rng = np.random.Generator(np.random.PCG64([base_seed, li]))
        tree, tip_species = simulate_gene_tree(rng, sp, pop_scale, tips_per_leaf)
        n_tips = len(tip_species)
        mu = float(np.exp(rng.normal(0.0, 0.3)))
        if cfg["subst"] == "hky":
            kind, sparams = SUBST_HKY, hky_params(2.0, pi_hky)
        else:
            kind, sparams = SUBST_EIGEN, gtr_eigen_params((1, 3, 0.8, 1.2, 3.5, 1), np.array([0.28, 0.22, 0.21, 0.29]))
        cat_rates = discrete_gamma_rates(0.5, cfg["n_cat"]) * mu
        cat_props = np.full(cfg["n_cat"], 1.0 / cfg["n_cat"])
        root_h = tree.height[tree.root]
        clock_rate = 0.25 / max(root_h, 1e-6)  # expected root-to-tip 0.25 subs/site
        if cfg["clock"] == "ucln":
            clock_kind = CLOCK_SPECIES_RATES
            br = np.full(tree.n_nodes, clock_rate)  # sequences simulated at the mean rate
        else:
            clock_kind = CLOCK_STRICT
            br = np.full(tree.n_nodes, clock_rate)
        pats, w = simulate_alignment(rng, tree, n_tips, cfg["n_sites"], kind, sparams, cat_rates, br, missing_frac)
        loci.append(Locus(tree=tree, tip_states=pats, weights=w, tip_species=tip_species, ploidy=ploidy,
                          subst_kind=kind, subst_params=sparams, cat_rates=cat_rates, cat_props=cat_props,
                          clock_kind=clock_kind, clock_rate=clock_rate))
    # morphological partitions on the species tree (CanisFBD-style: one TreeLikelihood per state count, LewisMK, ascertained)
    morph = []
    # (sharded runs: partition mi travels with the shard that holds locus (2 mi + 1) n / (2 n_morph) -- spread evenly along the
    # locus axis, so that every rank sees each partition once and no rank carries all of them)
    n_morph = int(cfg.get("n_morph", 0))
    for mi in range(n_morph):
        if not (lo <= morph_owner_locus(mi, cfg["n_loci"], n_morph) < hi):
            continue
        rng = np.random.Generator(np.random.PCG64([base_seed, 900000 + mi]))
        K = 2 + mi % 4
        rate = 0.6 / max(float(sp.height[sp.root]), 1e-6)
        pats, w, excl = simulate_morphology(rng, sp, K, int(cfg.get("n_chars", 24)), rate)
        morph.append(MorphPartition(tip_states=pats, weights=w, n_states=K, subst_params=lewis_mk_eigen(K),
                                    clock_rate=rate * float(np.exp(rng.normal(0.0, 0.2))), excluded=excl))
    return Problem(species=sp, loci=loci, species_rates=srates, pop_kind=pop_kind, pop_a=pop_a, pop_b=pop_b,
                   alpha=alpha, beta=beta, morph=morph)


def make_problem_parallel(name: str, n_loci: Optional[int] = None, locus_range: Optional[Tuple[int, int]] = None,
                          procs: int = 1, **kw) -> Problem:
    """make_problem with the loci generated by `procs` worker processes (every locus has its own (seed, index) stream, so
    the result is identical to the serial one).  The workers are fresh interpreters (`python -m starbeast2_b200.synth`)
    that send their chunk back pickled: safe after CUDA initialisation and under torchrun, and independent of __main__."""
    import os
    import pickle
    import subprocess
    import sys
    total = n_loci if n_loci is not None else CONFIGS[name]["n_loci"]
    lo, hi = locus_range if locus_range is not None else (0, total)
    n = hi - lo
    procs = max(1, min(procs, n // 16))
    if procs <= 1:
        return make_problem(name, n_loci=total, locus_range=(lo, hi), **kw)
    cuts = [lo + (n * i) // procs for i in range(procs + 1)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    workers = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        job = pickle.dumps((name, total, a, b, kw))
        w = subprocess.Popen([sys.executable, "-m", "starbeast2_b200.synth"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, cwd=root, env=env)
        w.stdin.write(job)
        w.stdin.close()
        workers.append(w)
    parts = []
    for w in workers:
        data = w.stdout.read()
        if w.wait() != 0:
            raise RuntimeError("synthetic-problem worker failed")
        parts.append(pickle.loads(data))
    pb = parts[0]
    for q in parts[1:]:
        pb.loci.extend(q.loci)
        pb.morph.extend(q.morph)   # every partition travels with exactly one chunk, in partition order
    return pb


def perturb_gene_tree(rng, tree: TreeArrays, n_tips: int) -> Tuple[TreeArrays, int]:
    """A BEAST 'Uniform'-operator style move: slide one internal non-root node between its oldest
    child and its parent.  Returns (new tree, moved node)."""
    t = tree.copy()
    G = t.n_nodes
    cands = [i for i in range(n_tips, G) if t.parent[i] >= 0]
    i = int(rng.choice(cands))
    lo = max(t.height[t.left[i]], t.height[t.right[i]])
    hi = t.height[t.parent[i]]
    t.height[i] = lo + (hi - lo) * rng.uniform(0.05, 0.95)
    return t, i


if __name__ == "__main__":   # worker of make_problem_parallel: job on stdin, pickled Problem chunk on stdout
    import pickle
    import sys
    _name, _total, _a, _b, _kw = pickle.loads(sys.stdin.buffer.read())
    sys.stdout.buffer.write(pickle.dumps(make_problem(_name, n_loci=_total, locus_range=(_a, _b), **_kw)))
