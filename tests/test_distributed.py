"""The N>1 path: loci sharded over ranks, one all-reduce(sum) of the reduce vector
{coalescent, likelihood, nIncompatible, (q, gamma, logR) x species nodes}, closed forms afterwards
(MultispeciesCoalescent.java:203-232 needs the per-branch sums over ALL loci before log(beta + gamma)).

CPU test (gloo, world_size 2): the sharding + reduce-vector protocol, with the oracle standing in for the device.
GPU test (gloo over one GPU, world_size 2): Engine.eval_posterior_distributed end to end."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.engine import shard_loci
from starbeast2_b200.problem import POP_ANALYTIC, Problem


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_reduce_vector(pb_shard: Problem):
    """What k_reduce_local leaves in the reduce vector, restated with the oracle."""
    S = pb_shard.species.n_nodes
    res = oracle.posterior(pb_shard)
    vec = np.zeros(3 + 3 * S)
    # per-locus GeneTree logP (0 in analytic mode) and likelihoods
    vec[0] = float(np.sum(res.per_locus_coal))
    vec[1] = float(np.nansum(res.per_locus_lik))
    vec[2] = float(np.sum(~np.isfinite(res.per_locus_coal)))
    if pb_shard.pop_kind == POP_ANALYTIC:
        for loc in pb_shard.loci:
            emb = oracle.gene_tree_update(loc.tree, loc.tip_species, pb_shard.species)
            for b in range(S):
                t, n, k = emb.coalescent_times(b), int(emb.n[b]), int(emb.k[b])
                g = sum((t[i + 1] - t[i]) * (n - i) * (n - (i + 1.0)) / 2.0 for i in range(k))
                if n - k > 1:
                    g += (t[k + 1] - t[k]) * (n - k) * (n - (k + 1.0)) / 2.0
                vec[3 + 3 * b] += k
                vec[3 + 3 * b + 1] += g / loc.ploidy
                vec[3 + 3 * b + 2] += k * math.log(loc.ploidy)
    return vec, res


def _finish(vec, pb):
    """k_finish restated: closed form per branch on the globally reduced sums."""
    S = pb.species.n_nodes
    coal = vec[0]
    if pb.pop_kind == POP_ANALYTIC:
        for b in range(S):
            q, g, r = int(round(vec[3 + 3 * b])), vec[3 + 3 * b + 1], vec[3 + 3 * b + 2]
            lgr = sum(math.log(pb.alpha + i) for i in range(q))
            coal += -r + pb.alpha * math.log(pb.beta) - (pb.alpha + q) * math.log(pb.beta + g) + lgr
    if vec[2] > 0:
        coal = -math.inf
    return coal, vec[1]


def _cpu_worker(rank, world, port, name, n_loci, n_sites, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_loci([1.0] * n_loci, world)[rank]
    shard = synth.make_problem(name, n_loci=n_loci, n_sites=n_sites, locus_range=(lo, hi))
    vec, _ = _local_reduce_vector(shard)
    t = torch.from_numpy(vec)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    coal, lik = _finish(t.numpy(), shard)
    if rank == 0:
        out.put((coal, lik))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["C4", "C2"])
def test_sharded_reduce_vector_protocol_gloo_cpu(name):
    n_loci, n_sites, world = 6, 60, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, name, n_loci, n_sites, q)) for r in range(world)]
    [p.start() for p in procs]
    coal, lik = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    full = oracle.posterior(synth.make_problem(name, n_loci=n_loci, n_sites=n_sites))
    assert coal == pytest.approx(full.coalescent, rel=1e-12)
    assert lik == pytest.approx(full.likelihood, rel=1e-12)


def _shard_values(per_locus, pb, n_loci, lo, hi):
    """The shard's slice of a global per-locus vector: its loci, then the morphological partitions that travel with it."""
    n_morph = len(pb.morph)
    mine = [n_loci + mi for mi in range(n_morph) if lo <= synth.morph_owner_locus(mi, n_loci, n_morph) < hi]
    return np.concatenate([np.asarray(per_locus)[lo:hi], np.asarray(per_locus)[mine]])


def _gpu_worker(rank, world, port, name, n_loci, n_sites, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)   # both ranks share cuda:0, so gloo (NCCL wants 1 GPU per rank)
    from starbeast2_b200.engine import Engine
    torch.cuda.set_device(0)
    lo, hi = shard_loci([1.0] * n_loci, world)[rank]
    shard = synth.make_problem(name, n_loci=n_loci, n_sites=n_sites, locus_range=(lo, hi))
    e = Engine.from_problem(shard, device=0)
    e.bind_torch()
    r = e.eval_posterior_distributed()
    # the operators' gene-tree walks over sharded loci: local mask, global freedoms / count / max height
    sp = shard.species
    node = sp.n_leaves + 1
    mask, tip, root, cnt = e.connecting_nodes_distributed(node)
    side = (np.arange(sp.n_leaves) % 2).astype(np.uint8)
    mh = e.max_height_distributed(side)
    out.put((rank, r.coalescent, r.likelihood, r.total, r.per_locus_lik.tolist(), r.per_branch.tolist(), (mask.tolist(), tip, root, cnt, mh)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C4", "C3"])
def test_engine_distributed_two_ranks(name):
    n_loci, n_sites, world = 6, 80, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, name, n_loci, n_sites, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    pb = synth.make_problem(name, n_loci=n_loci, n_sites=n_sites)
    full = oracle.posterior(pb)
    sp = pb.species
    w_mask, w_tip, w_root, w_cnt = oracle.connecting_nodes(pb, sp.n_leaves + 1)
    cmap = np.zeros(sp.n_nodes, np.int32)
    cmap[: sp.n_leaves] = np.arange(sp.n_leaves) % 2            # "right" leaves get canonical index 1 >= centre 1
    w_mh = oracle.max_height(pb, cmap, 1)
    node_off = np.concatenate([[0], np.cumsum([l.tree.n_nodes for l in pb.loci])])
    for rank, coal, lik, total, per_locus, per_branch, q in got:
        lo, hi = shard_loci([1.0] * n_loci, world)[rank]
        n_gene = int(node_off[hi] - node_off[lo])   # the mask covers gene trees only (morphological partitions have none)
        assert q[0] == w_mask[node_off[lo]:node_off[hi]].tolist() and len(q[0]) == n_gene
        assert q[1:] == (w_tip, w_root, w_cnt, w_mh)
        assert coal == pytest.approx(full.coalescent, rel=1e-9)      # every rank ends with the global totals
        assert lik == pytest.approx(full.likelihood, rel=1e-9)
        assert total == pytest.approx(full.total, rel=1e-9)
        lo, hi = shard_loci([1.0] * n_loci, world)[rank]
        np.testing.assert_allclose(per_locus, _shard_values(full.per_locus_lik, pb, n_loci, lo, hi), rtol=1e-9)
        if pb.pop_kind == POP_ANALYTIC:
            np.testing.assert_allclose(per_branch, full.per_branch, rtol=1e-9)


def _peer_worker(rank, world, port, name, n_loci, n_sites, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from starbeast2_b200.engine import Engine
    lo, hi = shard_loci([1.0] * n_loci, world)[rank]
    shard = synth.make_problem(name, n_loci=n_loci, n_sites=n_sites, locus_range=(lo, hi))
    e = Engine.from_problem(shard, device=rank)
    e.bind_torch()
    assert e.connect_peers()
    totals = []
    for it in range(5):                       # several evaluations: both buffer halves, sequence numbers
        e.mark_all_dirty()
        r = e.eval_posterior_distributed()
        totals.append((r.coalescent, r.likelihood, r.total))
    # an incremental step on rank 1's first locus only; every rank still takes part in the evaluation
    e.store()
    if rank == 1:
        rng = np.random.default_rng(5)
        t, _ = synth.perturb_gene_tree(rng, shard.loci[0].tree, shard.loci[0].n_tips)
        e.set_gene_tree(0, t)
        moved = (t.parent.tolist(), t.left.tolist(), t.right.tolist(), t.height.tolist())
    else:
        moved = None
    r2 = e.eval_posterior_distributed()
    e.restore()
    r3 = e.eval_posterior_distributed()
    # the reference's two-phase protocol with an asymmetric proposal (only rank 1 stages a tree): every call is one exchange
    # on every rank, whatever the rank's local state, so the phases must reproduce r2's global totals on both ranks
    e.store()
    if rank == 1:
        e.set_gene_tree(0, t)
    c_tot, _, _ = e.eval_coalescent()
    l_tot, _ = e.eval_likelihood()
    e.restore()
    r4 = e.eval_posterior_distributed()
    assert (r4.coalescent, r4.likelihood) == (r3.coalescent, r3.likelihood)
    out.put((rank, totals, (r2.coalescent, r2.likelihood, r2.total), (r3.coalescent, r3.likelihood, r3.total), moved,
             r.per_branch.tolist(), (c_tot, l_tot)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C4", "C3"])
def test_peer_memory_allreduce_two_gpus(name):
    """The in-kernel all-reduce over NVLink peer memory (sb2_comm_connect): needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n_loci, n_sites, world = 6, 80, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, name, n_loci, n_sites, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    pb = synth.make_problem(name, n_loci=n_loci, n_sites=n_sites)
    full = oracle.posterior(pb)
    for rank, totals, r2, r3, moved, per_branch, two_phase in got:
        assert two_phase == (r2[0], r2[1])                    # coalescent / likelihood phases == the single-call evaluation
        for coal, lik, total in totals:
            assert coal == pytest.approx(full.coalescent, rel=1e-9) and lik == pytest.approx(full.likelihood, rel=1e-9)
            assert total == pytest.approx(full.total, rel=1e-9)
        assert totals[0] == totals[-1]                    # repeatable bit for bit
        assert r3 == totals[-1]                           # restore brings the global totals back exactly
        if pb.pop_kind == POP_ANALYTIC:
            np.testing.assert_allclose(per_branch, full.per_branch, rtol=1e-9)
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2]   # summed in rank order: both ranks hold identical bits
    # the moved state, evaluated by the oracle
    from starbeast2_b200.problem import TreeArrays
    m = got[1][4]
    lo, _ = shard_loci([1.0] * n_loci, world)[1]
    pb.loci[lo].tree = TreeArrays(*[np.array(x) for x in m])
    w2 = oracle.posterior(pb)
    if math.isfinite(w2.coalescent):
        assert got[0][2][2] == pytest.approx(w2.total, rel=1e-9)
    else:
        assert got[0][2][0] == -math.inf


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C4", "C2"])
def test_one_host_thread_several_engines(name):
    """one JVM thread drives one engine per GPU.  Here: three engines with the loci sharded over
    them -- on different GPUs when the box has them, else all on cuda:0 -- evaluated through begin / read / host sum / write /
    end.  Every engine ends with the global totals; per-locus values are its shard's."""
    from starbeast2_b200.engine import Engine
    n_loci, world = 9, 3
    n_dev = torch.cuda.device_count()
    pb = synth.make_problem(name, n_loci=n_loci, n_sites=80)
    full = oracle.posterior(pb)
    shards = shard_loci([1.0] * n_loci, world)
    engines = []
    for r, (lo, hi) in enumerate(shards):
        shard = synth.make_problem(name, n_loci=n_loci, n_sites=80, locus_range=(lo, hi))
        engines.append(Engine.from_problem(shard, device=r % n_dev))
    for rep in range(2):
        res = Engine.eval_posterior_multi(engines)
        for (lo, hi), r in zip(shards, res):
            assert r.coalescent == pytest.approx(full.coalescent, rel=1e-9)
            assert r.likelihood == pytest.approx(full.likelihood, rel=1e-9)
            assert r.total == pytest.approx(full.total, rel=1e-9)
            np.testing.assert_allclose(r.per_locus_lik, _shard_values(full.per_locus_lik, pb, n_loci, lo, hi), rtol=1e-9)
            if pb.pop_kind == POP_ANALYTIC:
                np.testing.assert_allclose(r.per_branch, full.per_branch, rtol=1e-9)
        assert res[0].total == res[1].total == res[2].total          # same summation order everywhere
        for e in engines:
            e.mark_all_dirty()
    for e in engines:
        e.close()
