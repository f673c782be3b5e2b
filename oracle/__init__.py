"""ctypes front-end of the CPU oracle (oracle/sb2_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of sb2_oracle.c.  Nothing under starbeast2_b200/
imports this package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do, and only as the checker / reported CPU baseline.

Parity status: the StarBEAST2-owned half (GeneTree, population models, MultispeciesCoalescent,
StarBeastClock, UncorrelatedRates) is pinned by the reference's JUnit known-answer values
(tests/test_oracle_golden.py).  The BEAST.base half (TreeLikelihood & co) is PARITY UNPINNED:
its source is not under /root/reference and no reference test touches it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsb2oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sb2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


class _Problem(C.Structure):
    _fields_ = [
        ("nLoci", C.c_int32), ("S", C.c_int32), ("nSpeciesLeaves", C.c_int32), ("popKind", C.c_int32),
        ("sParent", C.c_void_p), ("sLeft", C.c_void_p), ("sRight", C.c_void_p),
        ("sHeight", C.c_void_p), ("speciesRates", C.c_void_p), ("popA", C.c_void_p), ("popB", C.c_void_p),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("nTips", C.c_void_p), ("nPat", C.c_void_p), ("nCat", C.c_void_p), ("substKind", C.c_void_p), ("clockKind", C.c_void_p),
        ("nodeOff", C.c_void_p), ("tipOff", C.c_void_p), ("patOff", C.c_void_p), ("stateOff", C.c_void_p), ("catOff", C.c_void_p),
        ("gParent", C.c_void_p), ("gLeft", C.c_void_p), ("gRight", C.c_void_p), ("gHeight", C.c_void_p),
        ("tipSpecies", C.c_void_p), ("tipStates", C.c_void_p), ("weights", C.c_void_p), ("constMask", C.c_void_p),
        ("ploidy", C.c_void_p), ("substParams", C.c_void_p), ("catRates", C.c_void_p), ("catProps", C.c_void_p),
        ("pInv", C.c_void_p), ("clockRate", C.c_void_p), ("explicitRates", C.c_void_p),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.sb2o_gene_tree_update.restype = C.c_int
        L.sb2o_gene_tree_update.argtypes = [C.c_int, C.c_int, _i32p, _f64p, _i32p, C.c_int, C.c_int, _i32p, _f64p, C.c_int,
                                            _i32p, _i32p, _f64p, _i32p, C.c_void_p]
        L.sb2o_coalescent_times.restype = None
        L.sb2o_coalescent_times.argtypes = [C.c_int, _i32p, _f64p, _i32p, _f64p, C.c_int, _f64p]
        L.sb2o_constant_logp.restype = C.c_double
        L.sb2o_constant_logp.argtypes = [C.c_double, C.c_double, _f64p, C.c_int, C.c_int]
        L.sb2o_linear_logp.restype = C.c_double
        L.sb2o_linear_logp.argtypes = [C.c_double, C.c_double, C.c_double, _f64p, C.c_int, C.c_int]
        L.sb2o_lwcr_branch_logp.restype = C.c_double
        L.sb2o_lwcr_branch_logp.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, _f64p, C.c_double, _f64p, C.c_int, C.c_int]
        L.sb2o_analytical_logp.restype = C.c_double
        L.sb2o_analytical_logp.argtypes = [C.c_double, C.c_double, C.c_int, _f64p, _f64p, _i64p, _i32p, _i32p]
        L.sb2o_starbeast_clock.restype = None
        L.sb2o_starbeast_clock.argtypes = [C.c_int, C.c_int, C.c_double, _f64p, _f64p, _f64p]
        L.sb2o_ucln_bin_rates.restype = None
        L.sb2o_ucln_bin_rates.argtypes = [C.c_int, C.c_double, _f64p]
        L.sb2o_uced_bin_rates.restype = None
        L.sb2o_uced_bin_rates.argtypes = [C.c_int, _f64p]
        L.sb2o_uncorrelated_rates.restype = None
        L.sb2o_uncorrelated_rates.argtypes = [C.c_int, C.c_int, _i32p, _f64p, C.c_double, C.c_int, C.c_int, _f64p]
        L.sb2o_hky_matrix.restype = None
        L.sb2o_hky_matrix.argtypes = [C.c_double, _f64p, C.c_double, _f64p]
        L.sb2o_eigen_matrix.restype = None
        L.sb2o_eigen_matrix.argtypes = [_f64p, _f64p, _f64p, C.c_double, _f64p]
        L.sb2o_tree_likelihood.restype = C.c_double
        L.sb2o_tree_likelihood.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, _u8p, _i32p,
                                           C.c_int, _f64p, _f64p, _f64p, C.c_double, C.c_void_p, _f64p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sb2o_eigen_matrix_k.restype = None
        L.sb2o_eigen_matrix_k.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_double, _f64p]
        L.sb2o_tree_likelihood_k.restype = C.c_double
        L.sb2o_tree_likelihood_k.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, _u8p, _i32p,
                                             _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sb2o_posterior.restype = C.c_int
        L.sb2o_random_local_rates.restype = None
        L.sb2o_random_local_rates.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _u8p, _f64p, C.c_double, _f64p]
        L.sb2o_connecting_nodes.restype = C.c_int
        L.sb2o_connecting_nodes.argtypes = [C.c_int, _i32p, _i32p, C.c_int, C.c_int, _i64p, _i64p, _i32p, _i32p, _i32p, _f64p, _i32p, C.c_int,
                                            _u8p, _f64p]
        L.sb2o_max_height.restype = C.c_double
        L.sb2o_max_height.argtypes = [_i32p, C.c_int, C.c_double, C.c_int, _i64p, _i64p, _i32p, _i32p, _i32p, _f64p, _i32p]
        L.sb2o_posterior.argtypes = [C.POINTER(_Problem), C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.sb2o_max_threads.restype = C.c_int
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# reference-owned half
# ---------------------------------------------------------------------------------------------

class Embedding:
    """Result of GeneTree.update(): n, k, per-branch sorted times (getCoalescentTimes), occupancy."""

    def __init__(self, compatible, n, k, raw_times, blk, assignment, occupancy, s_parent, s_height):
        self.compatible = bool(compatible)
        self.n = n
        self.k = k
        self._raw = raw_times
        self._blk = blk
        self.assignment = assignment
        self.occupancy = occupancy
        self._sp = s_parent
        self._sh = s_height

    def coalescent_times(self, node: int) -> np.ndarray:
        out = np.zeros(int(self.k[node]) + 2)
        lib().sb2o_coalescent_times(node, self._sp, self._sh, self.k, self._raw, self._blk, out)
        return out


def gene_tree_update(gene, tip_species, species, want_occupancy: bool = True) -> Embedding:
    G = gene.n_nodes
    n_tips = (G + 1) // 2
    S = species.n_nodes
    blk = n_tips
    n = np.zeros(S, np.int32)
    k = np.zeros(S, np.int32)
    times = np.zeros(S * blk)
    assign = np.zeros(G, np.int32)
    occ = np.zeros((G, S)) if want_occupancy else None
    ok = lib().sb2o_gene_tree_update(G, n_tips, gene.parent, gene.height, np.ascontiguousarray(tip_species, np.int32),
                                     S, species.n_leaves, species.parent, species.height, blk, n, k, times, assign,
                                     occ.ctypes.data if occ is not None else None)
    return Embedding(ok, n, k, times, blk, assign, occ, species.parent, species.height)


def constant_logp(pop_size, ploidy, times, n, k) -> float:
    return lib().sb2o_constant_logp(pop_size, ploidy, np.ascontiguousarray(times, np.float64), int(n), int(k))


def linear_logp(top, tip, ploidy, times, n, k) -> float:
    return lib().sb2o_linear_logp(top, tip, ploidy, np.ascontiguousarray(times, np.float64), int(n), int(k))


def lwcr_branch_logp(node, species, tip_pop, top_pop, ploidy, times, n, k) -> float:
    return lib().sb2o_lwcr_branch_logp(int(node), species.n_leaves, species.parent, species.left, species.right,
                                       np.ascontiguousarray(tip_pop, np.float64), np.ascontiguousarray(top_pop, np.float64),
                                       ploidy, np.ascontiguousarray(times, np.float64), int(n), int(k))


def analytical_logp(alpha, beta, ploidies, times_list, n_list, k_list) -> float:
    off = np.zeros(len(times_list), np.int64)
    o = 0
    for i, t in enumerate(times_list):
        off[i] = o
        o += len(t)
    flat = np.ascontiguousarray(np.concatenate(times_list), np.float64)
    return lib().sb2o_analytical_logp(alpha, beta, len(times_list), np.ascontiguousarray(ploidies, np.float64), flat, off,
                                      np.ascontiguousarray(n_list, np.int32), np.ascontiguousarray(k_list, np.int32))


def starbeast_clock(clock_rate, species_rates, occupancy) -> np.ndarray:
    G, S = occupancy.shape
    out = np.zeros(G)
    lib().sb2o_starbeast_clock(G, S, clock_rate, np.ascontiguousarray(species_rates, np.float64),
                               np.ascontiguousarray(occupancy, np.float64), out)
    return out


def ucln_bin_rates(n_bins, stdev) -> np.ndarray:
    out = np.zeros(n_bins)
    lib().sb2o_ucln_bin_rates(n_bins, stdev, out)
    return out


def uced_bin_rates(n_bins) -> np.ndarray:
    out = np.zeros(n_bins)
    lib().sb2o_uced_bin_rates(n_bins, out)
    return out


def uncorrelated_rates(S, bin_index, bin_rates, mean, estimate_root, root) -> np.ndarray:
    out = np.zeros(S)
    bi = np.ascontiguousarray(bin_index, np.int32)
    lib().sb2o_uncorrelated_rates(S, len(bi), bi, np.ascontiguousarray(bin_rates, np.float64), mean, int(estimate_root), int(root), out)
    return out


# ---------------------------------------------------------------------------------------------
# BEAST.base half (parity unpinned)
# ---------------------------------------------------------------------------------------------

def hky_matrix(kappa, pi, distance) -> np.ndarray:
    out = np.zeros(16)
    lib().sb2o_hky_matrix(kappa, np.ascontiguousarray(pi, np.float64), distance, out)
    return out.reshape(4, 4)


def eigen_matrix(evals, evec, ievc, distance) -> np.ndarray:
    out = np.zeros(16)
    lib().sb2o_eigen_matrix(np.ascontiguousarray(evals, np.float64), np.ascontiguousarray(evec, np.float64).reshape(16),
                            np.ascontiguousarray(ievc, np.float64).reshape(16), distance, out)
    return out.reshape(4, 4)


def tree_likelihood(locus, branch_rates, use_scaling: int = 2, want_details: bool = False, const_mask=None):
    """TreeLikelihood.calculateLogP of one locus, everything dirty.  branch_rates[G] as returned by
    branchRateModel.getRateForBranch.  Returns logL, or (logL, dict) with want_details."""
    from starbeast2_b200.problem import const_pattern_mask
    G, nT, P, Cn = locus.n_nodes, locus.n_tips, locus.n_pat, locus.n_cat
    t = locus.tree
    pat = np.zeros(P) if want_details else None
    parts = np.zeros((G - nT, Cn, P, 4)) if want_details else None
    mats = np.zeros((G, Cn, 4, 4)) if want_details else None
    scal = np.zeros((G - nT, P)) if want_details else None
    cm = const_mask if const_mask is not None else const_pattern_mask(locus.tip_states)
    cm = np.ascontiguousarray(cm, np.uint8)
    ptr = lambda a: a.ctypes.data if a is not None else None
    ll = lib().sb2o_tree_likelihood(G, nT, P, Cn, t.parent, t.left, t.right, t.height, locus.tip_states, locus.weights,
                                    locus.subst_kind, locus.subst_params, locus.cat_rates, locus.cat_props, locus.p_inv,
                                    cm.ctypes.data, np.ascontiguousarray(branch_rates, np.float64), use_scaling,
                                    ptr(pat), ptr(parts), ptr(mats), ptr(scal))
    if want_details:
        return ll, {"pattern_logl": pat, "partials": parts, "matrices": mats, "scaling": scal}
    return ll


def split_eigen_k(K: int, sp):
    sp = np.ascontiguousarray(sp, np.float64)
    return sp[0:K].copy(), sp[K:K + K * K].copy(), sp[K + K * K:K + 2 * K * K].copy(), sp[K + 2 * K * K:K + 2 * K * K + K].copy()


def eigen_matrix_k(K, evals, evec, ievc, distance) -> np.ndarray:
    out = np.zeros(K * K)
    lib().sb2o_eigen_matrix_k(K, np.ascontiguousarray(evals, np.float64), np.ascontiguousarray(evec, np.float64).reshape(K * K),
                              np.ascontiguousarray(ievc, np.float64).reshape(K * K), distance, out)
    return out.reshape(K, K)


def tree_likelihood_k(part, tree, branch_rates=None, use_scaling: int = 2, want_patterns: bool = False):
    """TreeLikelihood.calculateLogP of a K-state partition (starbeast2_b200.problem.MorphPartition) on `tree`, everything
    dirty, with the ascertainment correction when part.excluded is not empty."""
    K, G, nT, P, Cn = part.n_states, tree.n_nodes, part.n_tips, part.n_pat, part.n_cat
    ev, vec, ivec, pi = split_eigen_k(K, part.subst_params)
    br = np.full(G, part.clock_rate) if branch_rates is None else np.ascontiguousarray(branch_rates, np.float64)
    pat = np.zeros(P) if want_patterns else None
    exc = np.ascontiguousarray(part.excluded, np.int32)
    ll = lib().sb2o_tree_likelihood_k(G, nT, P, Cn, K, tree.parent, tree.left, tree.right, tree.height, part.tip_states, part.weights,
                                      ev, vec, ivec, pi, part.cat_rates, part.cat_props, br, use_scaling,
                                      len(exc), exc.ctypes.data if len(exc) else None, pat.ctypes.data if pat is not None else None)
    return (ll, pat) if want_patterns else ll


class PosteriorResult:
    def __init__(self, per_locus_coal, per_locus_lik, per_branch, total):
        self.per_locus_coal = per_locus_coal
        self.per_locus_lik = per_locus_lik
        self.per_branch = per_branch
        self.coalescent = float(total[0])
        self.likelihood = float(total[1])
        self.total = float(total[2])


class FlatProblem:
    """Concatenated arrays + the C struct; keeps the numpy buffers alive."""

    def __init__(self, pb):
        from starbeast2_b200.problem import SUBST_STRIDE, const_pattern_mask
        L = pb.n_loci
        sp = pb.species
        self.keep = []

        def k(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self.keep.append(a)
            return a.ctypes.data

        def prefix(vals):
            o = np.zeros(L + 1, np.int64)
            o[1:] = np.cumsum(vals)
            return o

        nT = [l.n_tips for l in pb.loci]
        nP = [l.n_pat for l in pb.loci]
        nC = [l.n_cat for l in pb.loci]
        G = [l.n_nodes for l in pb.loci]
        st = _Problem()
        st.nLoci, st.S, st.nSpeciesLeaves, st.popKind = L, sp.n_nodes, sp.n_leaves, pb.pop_kind
        st.sParent, st.sLeft, st.sRight = k(sp.parent, np.int32), k(sp.left, np.int32), k(sp.right, np.int32)
        st.sHeight = k(sp.height, np.float64)
        st.speciesRates = k(pb.species_rates, np.float64)
        st.popA = k(pb.pop_a if pb.pop_a is not None else np.zeros(1), np.float64)
        st.popB = k(pb.pop_b if pb.pop_b is not None else np.zeros(1), np.float64)
        st.alpha, st.beta = pb.alpha, pb.beta
        st.nTips, st.nPat, st.nCat = k(nT, np.int32), k(nP, np.int32), k(nC, np.int32)
        st.substKind = k([l.subst_kind for l in pb.loci], np.int32)
        st.clockKind = k([l.clock_kind for l in pb.loci], np.int32)
        st.nodeOff, st.tipOff, st.patOff = k(prefix(G), np.int64), k(prefix(nT), np.int64), k(prefix(nP), np.int64)
        st.stateOff = k(prefix([a * b for a, b in zip(nT, nP)]), np.int64)
        st.catOff = k(prefix(nC), np.int64)
        cat = lambda f, dt: k(np.concatenate([np.asarray(f(l)).reshape(-1) for l in pb.loci]), dt)
        st.gParent, st.gLeft, st.gRight = cat(lambda l: l.tree.parent, np.int32), cat(lambda l: l.tree.left, np.int32), cat(lambda l: l.tree.right, np.int32)
        st.gHeight = cat(lambda l: l.tree.height, np.float64)
        self.g_parent, self.g_left, self.g_right, self.g_height = self.keep[-4:]   # writable views for set_gene_tree
        self.node_off = prefix(G)
        st.tipSpecies = cat(lambda l: l.tip_species, np.int32)
        st.tipStates = cat(lambda l: l.tip_states, np.uint8)
        st.weights = cat(lambda l: l.weights, np.int32)
        st.constMask = cat(lambda l: const_pattern_mask(l.tip_states), np.uint8)
        st.ploidy = k([l.ploidy for l in pb.loci], np.float64)
        st.substParams = k(np.stack([l.subst_params[:SUBST_STRIDE] for l in pb.loci]), np.float64)
        st.catRates, st.catProps = cat(lambda l: l.cat_rates, np.float64), cat(lambda l: l.cat_props, np.float64)
        st.pInv = k([l.p_inv for l in pb.loci], np.float64)
        st.clockRate = k([l.clock_rate for l in pb.loci], np.float64)
        st.explicitRates = cat(lambda l: l.explicit_rates if l.explicit_rates is not None else np.ones(l.n_nodes), np.float64)
        self.struct = st
        self.n_loci = L
        self.S = sp.n_nodes


def posterior(pb, n_threads: int = 1, use_scaling: int = 2, flat: Optional[FlatProblem] = None) -> PosteriorResult:
    """One full posterior evaluation (MSC + clock + likelihood) of a starbeast2_b200.problem.Problem."""
    fp = flat if flat is not None else FlatProblem(pb)
    coal = np.zeros(fp.n_loci)
    lik = np.zeros(fp.n_loci)
    br = np.zeros(fp.S)
    tot = np.zeros(3)
    lib().sb2o_posterior(C.byref(fp.struct), n_threads, use_scaling, coal, lik, br, tot)
    morph = getattr(pb, "morph", None) or []
    if morph:   # K-state likelihoods on the species tree: appended after the loci, no coalescent term
        extra = np.array([tree_likelihood_k(m, pb.species, use_scaling=use_scaling) for m in morph])
        coal = np.concatenate([coal, np.zeros(len(morph))])
        lik = np.concatenate([lik, extra])
        tot[1] += extra.sum()
        tot[2] = tot[0] + tot[1]
    return PosteriorResult(coal, lik, br, tot)


class IncrementalBaseline:
    """Regime-B CPU baseline (sb2o_incr_*): one MCMC step the way the reference evaluates it when ONE gene tree changed --
    embedding of every locus (quirk Q2), population terms + clock rates + incremental TreeLikelihood of the changed locus.
    Non-analytic population models.  Test / bench infrastructure like everything under oracle/."""

    def __init__(self, pb):
        L = lib()
        L.sb2o_incr_create.restype = C.c_void_p
        L.sb2o_incr_create.argtypes = [C.POINTER(_Problem)]
        L.sb2o_incr_step.restype = C.c_double
        L.sb2o_incr_step.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_int]
        L.sb2o_incr_destroy.restype = None
        L.sb2o_incr_destroy.argtypes = [C.c_void_p]
        self._L = L
        self.fp = FlatProblem(pb)
        self._h = L.sb2o_incr_create(C.byref(self.fp.struct))
        if not self._h:
            raise ValueError("the incremental baseline supports the constant / linear population models only")

    def set_gene_tree(self, locus: int, tree):
        o, n = int(self.fp.node_off[locus]), tree.n_nodes
        self.fp.g_parent[o:o + n] = tree.parent
        self.fp.g_left[o:o + n] = tree.left
        self.fp.g_right[o:o + n] = tree.right
        self.fp.g_height[o:o + n] = tree.height

    def step(self, locus: int, tree=None) -> float:
        if tree is not None:
            self.set_gene_tree(locus, tree)
        return self._L.sb2o_incr_step(self._h, C.byref(self.fp.struct), int(locus))

    def raw_step(self):
        """(function, handle, struct pointer) for a timing loop without Python argument handling."""
        return self._L.sb2o_incr_step, self._h, C.byref(self.fp.struct)

    def close(self):
        if self._h:
            self._L.sb2o_incr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_local_rates(species, indicators, branch_rates, mean: float = 1.0) -> np.ndarray:
    """RandomLocalRates.update(): rates by species node number (indicators / branch_rates have S-1 entries)."""
    S = species.n_nodes
    out = np.zeros(S)
    ind = np.zeros(S, np.uint8)
    ind[: S - 1] = np.asarray(indicators, bool)[: S - 1]
    br = np.zeros(S)
    br[: S - 1] = np.asarray(branch_rates, float)[: S - 1]
    lib().sb2o_random_local_rates(S, species.root, species.left, species.right, species.height, ind, br, float(mean), out)
    return out


def _concat_trees(pb):
    node_off = np.zeros(len(pb.loci) + 1, np.int64)
    tip_off = np.zeros(len(pb.loci) + 1, np.int64)
    for i, loc in enumerate(pb.loci):
        node_off[i + 1] = node_off[i] + loc.tree.n_nodes
        tip_off[i + 1] = tip_off[i] + loc.n_tips
    cat = lambda f, dt: np.ascontiguousarray(np.concatenate([f(loc) for loc in pb.loci]), dt)
    return (node_off, tip_off, cat(lambda l: l.tree.parent, np.int32), cat(lambda l: l.tree.left, np.int32),
            cat(lambda l: l.tree.right, np.int32), cat(lambda l: l.tree.height, np.float64), cat(lambda l: l.tip_species, np.int32))


def connecting_nodes(pb, species_node: int, exponential: bool = False):
    """getConnectingNodes of CoordinatedUniform / CoordinatedExponential over all gene trees of `pb`:
    (mask[sum G] uint8, tipwardFreedom, rootwardFreedom, count)."""
    node_off, tip_off, gp, gl, gr, gh, ts = _concat_trees(pb)
    mask = np.zeros(int(node_off[-1]), np.uint8)
    fr = np.zeros(2)
    sp = pb.species
    n = lib().sb2o_connecting_nodes(sp.n_nodes, sp.left, sp.right, int(species_node), len(pb.loci), node_off, tip_off, gp, gl, gr, gh, ts,
                                    int(exponential), mask, fr)
    return mask, float(fr[0]), float(fr[1]), n


def max_height(pb, canonical_map, center_index: int, origin: float = float("inf")) -> float:
    """NodeReheight2.recalculateMaxHeight."""
    node_off, tip_off, gp, gl, gr, gh, ts = _concat_trees(pb)
    return lib().sb2o_max_height(np.ascontiguousarray(canonical_map, np.int32), int(center_index), float(origin), len(pb.loci),
                                 node_off, tip_off, gp, gl, gr, gh, ts)


def max_threads() -> int:
    return int(lib().sb2o_max_threads())
