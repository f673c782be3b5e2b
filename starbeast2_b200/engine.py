"""Engine: a thin Python handle on libsb2gpu's C ABI (include/sb2gpu.h), plus the multi-GPU plumbing.

One Engine = one GPU = one shard of loci.  With torch.distributed initialised (one process per GPU), loci are
block-partitioned over ranks (shard_loci) and each evaluation all-reduces the small reduce vector
{coalescent, likelihood, nIncompatible, (q, gamma, logR) x species nodes} -- the only exchange the path has
(MultispeciesCoalescent.java:203-232 needs per-branch sums over all loci before its closed form).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .problem import (CLOCK_EXPLICIT, POP_ANALYTIC, POP_CONSTANT, POP_LWCR, Locus, Problem, TreeArrays)


def shard_loci(weights: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Static block partition of loci over ranks, balanced by weight (pattern x node x category units).
    Returns [(first, end)] per rank; contiguous so per-locus outputs concatenate in rank order."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world_size <= 1:
        return [(0, n)]
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        i = int(np.searchsorted(cum, target, side="left"))
        # pick the closer of the two neighbouring cut points
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        i = max(bounds[-1], min(i, n))
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class EvalResult:
    __slots__ = ("coalescent", "likelihood", "total", "per_locus_coal", "per_locus_lik", "per_branch")

    def __init__(self, tot, coal, lik, br):
        self.coalescent, self.likelihood, self.total = float(tot[0]), float(tot[1]), float(tot[2])
        self.per_locus_coal, self.per_locus_lik, self.per_branch = coal, lik, br


class Engine:
    def __init__(self, device: int = 0, exact_order: bool = False, profile: bool = False):
        self._L = _lib.load()
        cfg = _lib.Sb2Config(device=device, exact_order=int(exact_order), profile=int(profile))
        h = C.c_void_p()
        rc = self._L.sb2_create(C.byref(cfg), C.byref(h))
        self._h = h
        if rc != 0:
            msg = self._L.sb2_last_error(h).decode() if h else "sb2_create failed"
            if h:
                self._L.sb2_destroy(h)
            self._h = None
            raise _lib.Sb2Error(msg)
        self.n_loci = 0
        self.S = 0
        self._loci_shape: List[Tuple[int, int, int]] = []
        self._locus_is_k: List[bool] = []   # likelihood-only (K-state) loci: no gene tree, no part in the operator queries
        self._bound_tensor = None

    # -- helpers ---------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc < 0:
            raise _lib.Sb2Error(self._L.sb2_last_error(self._h).decode())
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._L.sb2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- build -----------------------------------------------------------------------------------
    def add_locus(self, tip_states, weights, n_cat, tip_species, ploidy=2.0) -> int:
        ts = np.ascontiguousarray(tip_states, np.uint8)
        w = np.ascontiguousarray(weights, np.int32)
        sp = np.ascontiguousarray(tip_species, np.int32)
        idx = self._ck(self._L.sb2_add_locus(self._h, ts.shape[0], ts.shape[1], ts, w, int(n_cat), sp, float(ploidy)))
        self._loci_shape.append((ts.shape[0], ts.shape[1], int(n_cat)))
        self._locus_is_k.append(False)
        self.n_loci = idx + 1
        return idx

    def add_locus_k(self, tip_states, weights, n_states, n_cat=1, excluded=()) -> int:
        """A K-state LIKELIHOOD-ONLY locus (sb2_add_locus_k with tipToSpecies = NULL): a TreeLikelihood whose tree is not a
        GeneTree, e.g. a morphological partition on the species tree.  Its tree is staged with set_gene_tree."""
        ts = np.ascontiguousarray(tip_states, np.uint8)
        w = np.ascontiguousarray(weights, np.int32)
        ex = np.ascontiguousarray(excluded, np.int32)
        idx = self._ck(self._L.sb2_add_locus_k(self._h, ts.shape[0], ts.shape[1], int(n_states), ts, w, int(n_cat), None, 2.0,
                                               len(ex), ex.ctypes.data if len(ex) else None))
        self._locus_is_k.append(True)
        self._loci_shape.append((ts.shape[0], ts.shape[1], int(n_cat)))
        self.n_loci = idx + 1
        return idx

    def finalize(self, n_species_nodes: int, n_species_leaves: int):
        self._ck(self._L.sb2_finalize(self._h, n_species_nodes, n_species_leaves))
        self.S = n_species_nodes
        self.n_species_leaves = n_species_leaves

    @classmethod
    def from_problem(cls, pb: Problem, device: int = 0, exact_order: bool = False, profile: bool = False,
                     loci: Optional[Tuple[int, int]] = None) -> "Engine":
        """Registers (a shard of) a Problem and stages its full state."""
        e = cls(device=device, exact_order=exact_order, profile=profile)
        lo, hi = loci if loci is not None else (0, pb.n_loci)
        e.shard = (lo, hi)
        for loc in pb.loci[lo:hi]:
            e.add_locus(loc.tip_states, loc.weights, loc.n_cat, loc.tip_species, loc.ploidy)
        # morphological partitions (K-state likelihoods on the species tree) follow the loci.  A shard built by
        # synth.make_problem(locus_range=...) already carries only the partitions that travel with it; when a range of a
        # whole Problem is registered here instead, the same rule applies (synth.morph_owner_locus)
        allm = list(getattr(pb, "morph", None) or [])
        if (lo, hi) == (0, pb.n_loci):
            e.morph = allm
        else:
            from .synth import morph_owner_locus
            e.morph = [m for mi, m in enumerate(allm) if lo <= morph_owner_locus(mi, pb.n_loci, len(allm)) < hi]
        for m in e.morph:
            e.add_locus_k(m.tip_states, m.weights, m.n_states, m.n_cat, m.excluded)
        e.finalize(pb.species.n_nodes, pb.species.n_leaves)
        e.set_problem_state(pb)
        return e

    def set_problem_state(self, pb: Problem):
        lo, hi = getattr(self, "shard", (0, pb.n_loci))
        self.set_species_tree(pb.species)
        self.set_species_rates(pb.species_rates)
        self.set_pop_model(pb.pop_kind, pb.pop_a, pb.pop_b, pb.alpha, pb.beta)
        self.set_gene_trees(0, [l.tree for l in pb.loci[lo:hi]])
        for i, loc in enumerate(pb.loci[lo:hi]):
            self.set_subst_model(i, loc.subst_kind, loc.subst_params)
            self.set_site_model(i, loc.cat_rates, loc.cat_props, loc.p_inv)
            self.set_clock(i, loc.clock_kind, loc.clock_rate, loc.explicit_rates)
        for j, m in enumerate(getattr(self, "morph", [])):
            i = hi - lo + j
            self.set_gene_tree(i, pb.species)            # tree="@Tree.t:Species"
            self.set_subst_model(i, 1, m.subst_params)   # SB2_SUBST_EIGEN, K-state layout
            self.set_site_model(i, m.cat_rates, m.cat_props, 0.0)
            self.set_clock(i, 0, m.clock_rate, None)     # StrictClockModel

    # -- state input -----------------------------------------------------------------------------
    def set_species_tree(self, t: TreeArrays):
        self._ck(self._L.sb2_set_species_tree(self._h, t.parent, t.left, t.right, t.height))

    def set_gene_tree(self, locus: int, t: TreeArrays):
        self._ck(self._L.sb2_set_gene_tree(self._h, locus, t.parent, t.left, t.right, t.height))

    def set_gene_trees(self, first: int, trees: Sequence[TreeArrays]):
        if not trees:
            return
        cat = lambda f, dt: np.ascontiguousarray(np.concatenate([f(t) for t in trees]), dt)
        self._ck(self._L.sb2_set_gene_trees(self._h, first, len(trees), cat(lambda t: t.parent, np.int32),
                                            cat(lambda t: t.left, np.int32), cat(lambda t: t.right, np.int32),
                                            cat(lambda t: t.height, np.float64)))

    def set_gene_trees_flat(self, first: int, n: int, parent, left, right, height):
        """Pre-concatenated arrays (the bench keeps them that way to time the C call, not numpy.concatenate)."""
        self._ck(self._L.sb2_set_gene_trees(self._h, first, n, parent, left, right, height))

    def set_subst_model(self, locus: int, kind: int, params):
        p = np.zeros(64)
        src = np.asarray(params, np.float64).reshape(-1)
        p[: len(src)] = src
        self._ck(self._L.sb2_set_subst_model(self._h, locus, int(kind), p))

    def set_site_model(self, locus: int, cat_rates, cat_props, p_inv: float = 0.0):
        self._ck(self._L.sb2_set_site_model(self._h, locus, np.ascontiguousarray(cat_rates, np.float64),
                                            np.ascontiguousarray(cat_props, np.float64), float(p_inv)))

    def set_clock(self, locus: int, kind: int, mean_rate: float, rates=None):
        r = np.ascontiguousarray(rates, np.float64) if (rates is not None and kind == CLOCK_EXPLICIT) else None
        self._ck(self._L.sb2_set_clock(self._h, locus, int(kind), float(mean_rate), r.ctypes.data if r is not None else None))

    def set_species_rates(self, rates):
        self._ck(self._L.sb2_set_species_rates(self._h, np.ascontiguousarray(rates, np.float64)))

    def set_pop_model(self, kind: int, a=None, b=None, alpha: float = 0.0, beta: float = 0.0):
        aa = np.ascontiguousarray(a, np.float64) if a is not None else None
        bb = np.ascontiguousarray(b, np.float64) if b is not None else None
        self._ck(self._L.sb2_set_pop_model(self._h, int(kind), aa.ctypes.data if aa is not None else None,
                                           bb.ctypes.data if bb is not None else None, float(alpha), float(beta)))

    # -- evaluation ------------------------------------------------------------------------------
    def _bufs(self):
        return np.zeros(self.n_loci), np.zeros(self.n_loci), np.zeros(self.S), np.zeros(3)

    def eval_posterior(self) -> EvalResult:
        coal, lik, br, tot = self._bufs()
        self._ck(self._L.sb2_eval_posterior(self._h, coal.ctypes.data, lik.ctypes.data, br.ctypes.data, tot.ctypes.data))
        return EvalResult(tot, coal, lik, br)

    def eval_posterior_total(self, tot: np.ndarray):
        """Totals only, into a caller-owned float64[3] (no allocations on the hot path)."""
        self._ck(self._L.sb2_eval_posterior(self._h, None, None, None, tot.ctypes.data))

    def eval_coalescent(self):
        coal, _, br, tot = self._bufs()
        t = C.c_double()
        self._ck(self._L.sb2_eval_coalescent(self._h, coal.ctypes.data, br.ctypes.data, C.addressof(t)))
        return t.value, coal, br

    def eval_likelihood(self):
        _, lik, _, _ = self._bufs()
        t = C.c_double()
        self._ck(self._L.sb2_eval_likelihood(self._h, lik.ctypes.data, C.addressof(t)))
        return t.value, lik

    def eval_begin(self):
        self._ck(self._L.sb2_eval_begin(self._h))

    def eval_end(self) -> EvalResult:
        coal, lik, br, tot = self._bufs()
        self._ck(self._L.sb2_eval_end(self._h, coal.ctypes.data, lik.ctypes.data, br.ctypes.data, tot.ctypes.data))
        return EvalResult(tot, coal, lik, br)

    def eval_end_total(self, tot: np.ndarray):
        self._ck(self._L.sb2_eval_end(self._h, None, None, None, tot.ctypes.data))

    def store(self):
        self._ck(self._L.sb2_store(self._h))

    def restore(self):
        self._ck(self._L.sb2_restore(self._h))

    def accept(self):
        self._ck(self._L.sb2_accept(self._h))

    def mark_all_dirty(self):
        self._ck(self._L.sb2_mark_all_dirty(self._h))

    # -- multi-GPU plumbing ----------------------------------------------------------------------
    def reduce_vector_len(self) -> int:
        p, n = C.c_void_p(), C.c_int32()
        self._ck(self._L.sb2_reduce_vector(self._h, C.byref(p), C.byref(n)))
        return n.value

    def bind_torch(self, stream=None):
        """Allocates the reduce vector as a torch tensor (so torch.distributed can all-reduce it in place) and
        runs the engine on torch's current stream."""
        import torch
        n = self.reduce_vector_len()
        self._bound_tensor = torch.zeros(n, dtype=torch.float64, device="cuda")
        self._ck(self._L.sb2_bind_reduce_vector(self._h, self._bound_tensor.data_ptr()))
        st = stream if stream is not None else torch.cuda.current_stream()
        self._ck(self._L.sb2_set_stream(self._h, st.cuda_stream))
        return self._bound_tensor

    def connect_peers(self, group=None) -> bool:
        """Collective: exchanges the CUDA-IPC handles of the engines' communication buffers (one all_gather) and switches
        the single-call evaluations to the in-kernel peer-memory all-reduce.  Returns False (and changes nothing) when
        there is a single rank."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return False
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_ubyte * 64)()
        self._ck(self._L.sb2_comm_handle(self._h, buf))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        blob = b"".join(bytes(t.cpu().tolist()) for t in allh)
        self._ck(self._L.sb2_comm_connect(self._h, rank, world, blob))
        dist.barrier(group=group)
        self._peers = True
        return True

    def read_reduce_vector(self) -> np.ndarray:
        out = np.zeros(self.reduce_vector_len())
        self._ck(self._L.sb2_read_reduce_vector(self._h, out))
        return out

    def write_reduce_vector(self, v):
        self._ck(self._L.sb2_write_reduce_vector(self._h, np.ascontiguousarray(v, np.float64)))

    @staticmethod
    def eval_posterior_multi(engines: Sequence["Engine"]) -> List[EvalResult]:
        """One host thread, several engines (one per GPU, loci sharded): the in-process twin of eval_posterior_distributed,
        what a single JVM does.  All engines run their kernels concurrently; the reduce vectors
        are summed on the host in engine order, so every engine finishes with the same global totals."""
        for e in engines:
            e.eval_begin()
        total = None
        for e in engines:
            v = e.read_reduce_vector()
            total = v if total is None else total + v
        for e in engines:
            e.write_reduce_vector(total)
        return [e.eval_end() for e in engines]

    def eval_posterior_distributed(self, group=None) -> EvalResult:
        """begin -> all-reduce(sum) of the reduce vector over ranks -> end.  Totals are global, per-locus arrays local."""
        import torch.distributed as dist
        if getattr(self, "_peers", False):
            return self.eval_posterior()      # the all-reduce happens inside the reduction kernel, over peer memory
        self.eval_begin()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self._bound_tensor is None:
                raise _lib.Sb2Error("call bind_torch() before distributed evaluation")
            dist.all_reduce(self._bound_tensor, op=dist.ReduceOp.SUM, group=group)
        return self.eval_end()

    # -- read-back -------------------------------------------------------------------------------
    # -- operator-side queries over the resident gene trees (CoordinatedUniform/Exponential, NodeReheight2) ----------
    def connecting_nodes(self, species_node: int, want_mask: bool = True):
        """(mask[sum G] uint8 or None, tipwardFreedom, rootwardFreedom, count) of the last evaluated / restored state."""
        total = sum(2 * sh[0] - 1 for sh in self._loci_shape)
        mask = np.zeros(total, np.uint8) if want_mask else None
        fr = np.zeros(2)
        cnt = C.c_int64(0)
        self._ck(self._L.sb2_connecting_nodes(self._h, int(species_node), mask.ctypes.data if want_mask else None, fr, C.byref(cnt)))
        if want_mask and any(self._locus_is_k):
            # the C ABI's mask runs over every locus in engine order (likelihood-only loci: zeros); the reference's walk only
            # knows gene trees, so their slots are cut out here
            keep = np.concatenate([np.full(2 * sh[0] - 1, not k) for sh, k in zip(self._loci_shape, self._locus_is_k)])
            mask = mask[keep]
        return mask, float(fr[0]), float(fr[1]), int(cnt.value)

    def connecting_nodes_distributed(self, species_node: int, group=None):
        """Loci sharded over ranks: the mask is this rank's, the freedoms are the minima over all ranks (the operator's draw
        must be the same everywhere), the count is global."""
        import torch
        import torch.distributed as dist
        mask, tip, root, cnt = self.connecting_nodes(species_node)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
            fr = torch.tensor([tip, root], dtype=torch.float64, device=dev)
            dist.all_reduce(fr, op=dist.ReduceOp.MIN, group=group)
            c = torch.tensor([cnt], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
            tip, root, cnt = float(fr[0]), float(fr[1]), int(c[0])
        return mask, tip, root, cnt

    def max_height_distributed(self, leaf_is_right, group=None) -> float:
        import torch
        import torch.distributed as dist
        h = self.max_height(leaf_is_right)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            t = torch.tensor([h], dtype=torch.float64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            h = float(t[0])
        return h

    def max_height(self, leaf_is_right) -> float:
        out = C.c_double(0.0)
        self._ck(self._L.sb2_max_height(self._h, np.ascontiguousarray(leaf_is_right, np.uint8), C.byref(out)))
        return float(out.value)

    def branch_rates(self, locus: int) -> np.ndarray:
        out = np.zeros(2 * self._loci_shape[locus][0] - 1)
        self._ck(self._L.sb2_get_branch_rates(self._h, locus, out))
        return out

    def embedding(self, locus: int):
        n, k = np.zeros(self.S, np.int32), np.zeros(self.S, np.int32)
        a = np.zeros(2 * self._loci_shape[locus][0] - 1, np.int32)
        br = np.zeros(self.S)
        self._ck(self._L.sb2_get_embedding(self._h, locus, n.ctypes.data, k.ctypes.data, a.ctypes.data, br.ctypes.data))
        return n, k, a, br

    def partials(self, locus: int, node: int):
        nt, npat, nc = self._loci_shape[locus]
        p = np.zeros((nc, npat, 4))
        ex = np.zeros((nc, npat), np.int32)
        self._ck(self._L.sb2_get_partials(self._h, locus, node, p.ctypes.data, ex.ctypes.data))
        return p, ex

    def matrices(self, locus: int, node: int) -> np.ndarray:
        nc = self._loci_shape[locus][2]
        m = np.zeros((nc, 4, 4))
        self._ck(self._L.sb2_get_matrices(self._h, locus, node, m.reshape(-1)))
        return m

    def pattern_logl(self, locus: int) -> np.ndarray:
        out = np.zeros(self._loci_shape[locus][1])
        self._ck(self._L.sb2_get_pattern_logl(self._h, locus, out))
        return out

    # -- introspection ---------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._L.sb2_launch_count(self._h))

    def last_node_updates(self) -> int:
        return int(self._L.sb2_last_node_updates(self._h))

    def kernel_time(self) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        self._ck(self._L.sb2_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_profile(self, on: bool):
        self._ck(self._L.sb2_set_profile(self._h, int(on)))

    def device_bytes(self) -> int:
        return int(self._L.sb2_device_bytes(self._h))

    def exchange_wait_us(self) -> float:
        return float(self._L.sb2_exchange_wait_us(self._h))
