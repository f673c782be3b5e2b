"""alignment -> unique site patterns + weights on the device (BEAST.base Alignment.calcPatterns).
Byte / integer work: bit-identical to numpy.unique(axis=1), which is what the host mirror (plugins.Alignment) uses."""
import os

import numpy as np
import pytest

from starbeast2_b200 import synth
from starbeast2_b200._lib import Sb2Error
from starbeast2_b200.io import compress_alignment, read_fasta

pytestmark = pytest.mark.gpu


def check(mat):
    pats, w, smap = compress_alignment(mat)
    want_p, want_inv, want_w = np.unique(mat, axis=1, return_inverse=True, return_counts=True)
    assert pats.shape == want_p.shape and np.array_equal(pats, want_p)
    assert np.array_equal(w, want_w.astype(np.int32)) and int(w.sum()) == mat.shape[1]
    assert np.array_equal(smap, np.asarray(want_inv).reshape(-1).astype(np.int32))
    assert np.array_equal(pats[:, smap], mat)                    # decode(encode(x)) == x
    p2, w2, none = compress_alignment(mat, want_map=False)
    assert none is None and np.array_equal(p2, pats) and np.array_equal(w2, w)


@pytest.mark.parametrize("n_tips,n_sites,alphabet", [(1, 1, 4), (2, 7, 2), (16, 1000, 5), (30, 1000, 5), (120, 500, 5), (120, 4097, 3),
                                                     (9, 16384, 2), (257, 300, 5),
                                                     (2, 16385, 4), (30, 40000, 5), (12, 100000, 3)])   # longer than one block: block sort + merge
def test_random_alignments_match_numpy_unique(n_tips, n_sites, alphabet):
    rng = np.random.default_rng(n_tips * 100003 + n_sites)
    mat = rng.integers(0, alphabet, size=(n_tips, n_sites)).astype(np.uint8)
    if n_sites > 50:   # realistic redundancy: most columns are copies of a few
        src = rng.integers(0, max(2, n_sites // 7), size=n_sites)
        mat = mat[:, src]
    check(np.ascontiguousarray(mat))


def test_degenerate_and_error_cases():
    check(np.zeros((5, 64), np.uint8))                           # every column identical: one pattern of weight 64
    check(np.arange(200, dtype=np.uint8)[None, :].repeat(3, 0))  # every column distinct, bytes above the state range
    check(np.zeros((2, 16385), np.uint8))                        # one pattern spread over two blocks of the sorter
    check(np.arange(40000, dtype=np.uint32).view(np.uint8).reshape(-1, 4).T.copy())   # 40,000 distinct columns, three blocks
    with pytest.raises(Sb2Error):
        compress_alignment(np.zeros((2, 0), np.uint8))           # no sites


def test_simulated_locus_and_engine_round_trip():
    """Sites simulated down a gene tree (C2 shape) -> device compression -> the engine evaluates the compressed locus to the
    same log-likelihood as the oracle on the expanded (site-by-site) alignment."""
    import oracle
    from starbeast2_b200.engine import Engine
    from starbeast2_b200.problem import Locus, Problem
    pb = synth.make_problem("C2", n_loci=2, n_sites=400)
    loci = []
    for loc in pb.loci:
        sites = np.repeat(loc.tip_states, loc.weights, axis=1)                  # expand patterns back to sites ...
        perm = np.random.default_rng(1).permutation(sites.shape[1])
        sites = np.ascontiguousarray(sites[:, perm])                            # ... in arbitrary site order
        pats, w, smap = compress_alignment(sites)
        assert np.array_equal(pats[:, smap], sites)
        loci.append(Locus(tree=loc.tree, tip_states=pats, weights=w, tip_species=loc.tip_species, ploidy=loc.ploidy, subst_kind=loc.subst_kind,
                          subst_params=loc.subst_params, cat_rates=loc.cat_rates, cat_props=loc.cat_props, clock_rate=loc.clock_rate))
    pb2 = Problem(species=pb.species, loci=loci, pop_kind=pb.pop_kind, pop_a=pb.pop_a, pop_b=pb.pop_b, species_rates=pb.species_rates)
    e = Engine.from_problem(pb2)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    np.testing.assert_allclose(got.per_locus_lik, want.per_locus_lik, rtol=1e-10)
    e.close()


def test_device_compression_equals_the_host_alignment_mirror(tmp_path):
    """FASTA -> plugins.Alignment (host numpy, what the Java side's BEAST Alignment does) and -> the device path give the
    same patterns and weights."""
    from starbeast2_b200.plugins import Alignment
    rng = np.random.default_rng(5)
    names = [f"t{i:02d}" for i in range(12)]
    base = rng.integers(0, 4, size=300)
    lines = []
    mats = {}
    for n in names:
        s = base.copy()
        flip = rng.random(300) < 0.04
        s[flip] = rng.integers(0, 4, size=int(flip.sum()))
        gaps = rng.random(300) < 0.02
        mats[n] = np.where(gaps, 4, s)
        lines.append(f">{n}\n" + "".join("ACGT-"[x] for x in mats[n]) + "\n")
    fa = tmp_path / "x.fasta"
    fa.write_text("".join(lines))
    aln = Alignment(sequences=read_fasta(str(fa)))
    mat = np.array([mats[n] for n in sorted(names)], np.uint8)
    pats, w, smap = compress_alignment(mat)
    assert np.array_equal(pats, aln.patterns) and np.array_equal(w, aln.weights)
