"""Reading alignments the way the StarBEAST2 tutorials ship them (FASTA; tutorial/data/*.fasta) -- the step before the hot
path.  Pattern compression itself is plugins.Alignment (BEAST's Alignment.calcPatterns)."""
from __future__ import annotations

from typing import Dict


def read_fasta(path: str) -> Dict[str, str]:
    """{taxon: sequence}; the taxon is the first word of the header line."""
    seqs: Dict[str, str] = {}
    name = None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                name = line[1:].split()[0]
                seqs[name] = ""
            elif name is not None:
                seqs[name] += line
    return seqs


def compress_alignment(matrix, device: int = 0, want_map: bool = True):
    """Site patterns of an alignment on the device (libsb2gpu sb2_compress_alignment): matrix uint8 [tips, sites] ->
    (patterns uint8 [tips, nPat] in lexicographic column order, weights int32 [nPat], site_to_pattern int32 [sites] | None).
    Same result as numpy.unique(matrix, axis=1, return_counts=True, return_inverse=True); no CPU fallback."""
    import ctypes as C

    import numpy as np

    from . import _lib
    L = _lib.load()
    m = np.ascontiguousarray(matrix, np.uint8)
    n_tips, n_sites = m.shape
    pats = np.zeros((n_tips, n_sites), np.uint8)
    w = np.zeros(n_sites, np.int32)
    smap = np.zeros(n_sites, np.int32) if want_map else None
    n = C.c_int32(0)
    rc = L.sb2_compress_alignment(int(device), n_tips, n_sites, m, pats.reshape(-1), w, smap.ctypes.data if want_map else None, C.byref(n))
    if rc != 0:
        raise _lib.Sb2Error(L.sb2_compress_last_error().decode())
    k = int(n.value)
    return np.ascontiguousarray(pats.reshape(-1)[: n_tips * k].reshape(n_tips, k)), w[:k].copy(), smap
