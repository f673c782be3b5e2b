"""Flat host-side description of one StarBEAST2 posterior state (plain numpy, no torch).

This is the data the reference keeps in BEAST2 objects, flattened the way the C ABI
(include/sb2gpu.h) wants it: every array is indexed by the *current* BEAST node number
(leaves 0..tips-1, root = last node).

Reference objects mirrored:
  SpeciesTree           -> starbeast2.SpeciesTree (src/starbeast2/SpeciesTree.java) node arrays
  Locus tree arrays     -> beast.base.evolution.tree.Tree of one GeneTree (GeneTree.java:22)
  Locus.tip_species     -> SpeciesTreeInterface.getTipNumberMap (SpeciesTreeInterface.java:25-58)
  Locus subst/site/clock-> TreeLikelihood inputs siteModel / branchRateModel (fxtemplates/StarBeast2.xml:158-163)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

SUBST_HKY = 0
SUBST_EIGEN = 1
SUBST_STRIDE = 40          # doubles of a 4-state model; K-state models take K + 2 K^2 + K (<= SUBST_STRIDE_MAX)
SUBST_STRIDE_MAX = 64

CLOCK_STRICT = 0
CLOCK_SPECIES_RATES = 1
CLOCK_EXPLICIT = 2

POP_CONSTANT = 0
POP_LWCR = 1
POP_ANALYTIC = 2

MISSING_STATE = 4


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class TreeArrays:
    """parent/left/right/height by node number; -1 for absent parent/children."""
    parent: np.ndarray
    left: np.ndarray
    right: np.ndarray
    height: np.ndarray

    def __post_init__(self):
        self.parent = _i32(self.parent)
        self.left = _i32(self.left)
        self.right = _i32(self.right)
        self.height = _f64(self.height)

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    @property
    def n_leaves(self) -> int:
        return int((self.left < 0).sum())

    @property
    def root(self) -> int:
        return int(np.nonzero(self.parent < 0)[0][0])

    def copy(self) -> "TreeArrays":
        return TreeArrays(self.parent.copy(), self.left.copy(), self.right.copy(), self.height.copy())


def hky_params(kappa: float, pi) -> np.ndarray:
    sp = np.zeros(SUBST_STRIDE)
    sp[0] = kappa
    sp[1:5] = pi
    return sp


def eigen_params(evals, evec, ievc, pi) -> np.ndarray:
    sp = np.zeros(SUBST_STRIDE)
    sp[0:4] = evals
    sp[4:20] = np.asarray(evec).reshape(16)
    sp[20:36] = np.asarray(ievc).reshape(16)
    sp[36:40] = pi
    return sp


def gtr_eigen_params(rates6, pi) -> np.ndarray:
    """GTR rate matrix -> eigen system the way BEAST's GeneralSubstitutionModel is set up:
    Q_ij = r_ij * pi_j, rows sum to zero, normalised so that -sum_i pi_i Q_ii = 1.
    rates6 = (AC, AG, AT, CG, CT, GT).  The decomposition itself is host work in the reference
    (DefaultEigenSystem on parameter change); any valid (eval, V, V^-1) triple gives the same P(t)."""
    pi = np.asarray(pi, dtype=np.float64)
    ac, ag, at, cg, ct, gt = rates6
    r = np.array([[0, ac, ag, at], [ac, 0, cg, ct], [ag, cg, 0, gt], [at, ct, gt, 0]], dtype=np.float64)
    q = r * pi[None, :]
    np.fill_diagonal(q, 0.0)
    np.fill_diagonal(q, -q.sum(axis=1))
    q /= -(pi * np.diag(q)).sum()
    # symmetrise for a real, well conditioned decomposition: S = D^1/2 Q D^-1/2
    d = np.sqrt(pi)
    s = (d[:, None] * q) / d[None, :]
    w, u = np.linalg.eigh((s + s.T) / 2.0)
    evec = u / d[:, None]
    ievc = u.T * d[None, :]
    return eigen_params(w, evec, ievc, pi)


def const_pattern_mask(tip_states: np.ndarray) -> np.ndarray:
    """TreeLikelihood.calcConstantPatternIndices with useAmbiguities=false: bit s is set when every
    unambiguous tip of the pattern shows state s (ambiguous / missing tips are ignored)."""
    ts = np.asarray(tip_states)
    mask = np.full(ts.shape[1], 0xF, dtype=np.uint8)
    for s in range(4):
        bad = ((ts < 4) & (ts != s)).any(axis=0)
        mask[bad] &= np.uint8(~(1 << s) & 0xF)
    return mask


@dataclass
class Locus:
    tree: TreeArrays
    tip_states: np.ndarray          # uint8 [n_tips, n_pat], 0..3, >=4 missing/ambiguous
    weights: np.ndarray             # int32 [n_pat]
    tip_species: np.ndarray         # int32 [n_tips] species-tree leaf node number of each gene tip
    ploidy: float = 2.0
    subst_kind: int = SUBST_HKY
    subst_params: np.ndarray = field(default_factory=lambda: hky_params(2.0, [0.25] * 4))
    cat_rates: np.ndarray = field(default_factory=lambda: np.ones(1))   # already x mutationRate
    cat_props: np.ndarray = field(default_factory=lambda: np.ones(1))
    p_inv: float = 0.0
    clock_kind: int = CLOCK_STRICT
    clock_rate: float = 1.0
    explicit_rates: Optional[np.ndarray] = None

    def __post_init__(self):
        self.tip_states = np.ascontiguousarray(self.tip_states, dtype=np.uint8)
        self.weights = _i32(self.weights)
        self.tip_species = _i32(self.tip_species)
        self.subst_params = _f64(self.subst_params)
        self.cat_rates = _f64(self.cat_rates)
        self.cat_props = _f64(self.cat_props)
        assert self.tip_states.shape == (self.n_tips, self.n_pat)
        assert self.tree.n_nodes == 2 * self.n_tips - 1

    @property
    def n_tips(self) -> int:
        return int(self.tip_species.shape[0])

    @property
    def n_pat(self) -> int:
        return int(self.weights.shape[0])

    @property
    def n_cat(self) -> int:
        return int(self.cat_rates.shape[0])

    @property
    def n_nodes(self) -> int:
        return 2 * self.n_tips - 1

    @property
    def units(self) -> int:
        """pattern x node x category units of one full evaluation (BASELINE.json metric)."""
        return self.n_pat * (self.n_tips - 1) * self.n_cat


def eigen_params_k(evals, evec, ievc, pi) -> np.ndarray:
    """{eval[K], evec[K*K], ievc[K*K], pi[K]} (row-major), the K-state layout of SB2_SUBST_EIGEN; K = 4 gives eigen_params."""
    K = len(evals)
    sp = np.zeros(K + 2 * K * K + K)
    sp[0:K] = evals
    sp[K:K + K * K] = np.asarray(evec).reshape(K * K)
    sp[K + K * K:K + 2 * K * K] = np.asarray(ievc).reshape(K * K)
    sp[K + 2 * K * K:] = pi
    return sp


def lewis_mk_eigen(K: int) -> np.ndarray:
    """LewisMK (morph-models package, named at tutorial/CanisFBD/Canis-FBD.xml:975): all K (K-1) rates equal, equal
    frequencies, normalised to one expected change per unit time as GeneralSubstitutionModel does:
    Q_ij = 1/(K-1) (i != j), Q_ii = -1.  Symmetric, so an orthonormal eigen system serves as (V, V^-1)."""
    q = np.full((K, K), 1.0 / (K - 1))
    np.fill_diagonal(q, -1.0)
    w, u = np.linalg.eigh(q)
    return eigen_params_k(w, u, u.T, np.full(K, 1.0 / K))


@dataclass
class MorphPartition:
    """One morphological TreeLikelihood of the FBD configs: K-state characters scored on the SPECIES tree's tips
    (tutorial/CanisFBD/Canis-FBD.xml:964-1033), a likelihood without a coalescent term.  `excluded` lists the ascertained
    dummy patterns of an `ascertained="true"` FilteredAlignment (their weight is 0)."""
    tip_states: np.ndarray          # uint8 [n species leaves, n_pat]; >= n_states: missing / ambiguous
    weights: np.ndarray             # int32 [n_pat]
    n_states: int
    subst_params: np.ndarray        # eigen_params_k layout
    cat_rates: np.ndarray = field(default_factory=lambda: np.ones(1))
    cat_props: np.ndarray = field(default_factory=lambda: np.ones(1))
    clock_rate: float = 1.0
    excluded: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    def __post_init__(self):
        self.tip_states = np.ascontiguousarray(self.tip_states, dtype=np.uint8)
        self.weights = _i32(self.weights)
        self.subst_params = _f64(self.subst_params)
        self.cat_rates = _f64(self.cat_rates)
        self.cat_props = _f64(self.cat_props)
        self.excluded = _i32(self.excluded)
        K = self.n_states
        assert 2 <= K <= 5 and self.subst_params.shape[0] == K + 2 * K * K + K

    @property
    def n_tips(self) -> int:
        return int(self.tip_states.shape[0])

    @property
    def n_pat(self) -> int:
        return int(self.weights.shape[0])

    @property
    def n_cat(self) -> int:
        return int(self.cat_rates.shape[0])


@dataclass
class Problem:
    species: TreeArrays
    loci: List[Locus]
    species_rates: Optional[np.ndarray] = None   # [S]; None -> all ones
    pop_kind: int = POP_CONSTANT
    pop_a: Optional[np.ndarray] = None           # CONSTANT: N[S]; LWCR: tip sizes [n species leaves]
    pop_b: Optional[np.ndarray] = None           # LWCR: top sizes [S-1]
    alpha: float = 0.0                           # ANALYTIC: populationShape
    beta: float = 0.0                            # ANALYTIC: mean*(alpha-1), by the reference rule (quirk Q1)
    morph: List[MorphPartition] = field(default_factory=list)   # K-state likelihoods on the species tree (after the loci in every per-locus output)

    def __post_init__(self):
        S = self.species.n_nodes
        if self.species_rates is None:
            self.species_rates = np.ones(S)
        self.species_rates = _f64(self.species_rates)
        if self.pop_a is not None:
            self.pop_a = _f64(self.pop_a)
        if self.pop_b is not None:
            self.pop_b = _f64(self.pop_b)

    @property
    def n_loci(self) -> int:
        return len(self.loci)

    @property
    def units(self) -> int:
        return sum(l.units for l in self.loci)

    def algorithmic_bytes_survey(self) -> float:
        """level-synchronous traffic, B/unit = [32(2 tips-3)+tips]/(tips-1) + 16/C."""
        return float(sum(l.units * ((32.0 * (2 * l.n_tips - 3) + l.n_tips) / (l.n_tips - 1) + 16.0 / l.n_cat)
                         for l in self.loci))
