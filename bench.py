#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200.

A "step" is one full posterior evaluation with every node dirty (SURVEY.md 8d regime A): multispecies-coalescent
embedding + per-branch terms + species-clock rates + transition matrices + Felsenstein pruning of every gene tree
+ root integration, for all loci of the workload.

Headline leg (value / e2e / roofline / cpu_baseline): BASELINE.json configs[1] ("C2": 50 loci x 1,000 bp x 30
individuals / 10 species, HKY, ConstantPopulations).  N>1: weak scaling, every rank holds its own 50-locus shard of one
50*N-locus posterior; the reduce vector is all-reduced inside the reduction kernel over peer memory (NVLink).

North-star legs, run after the headline leg in the same process and carried in keys the driver's record keeps:
  config.legs.C5_strong  BASELINE configs[4]: 2,000 loci x 500 bp x 120 tips, GTR+G4, UCLN -- STRONG scaling, 2,000/N loci per GPU
  config.legs.C4_strong  BASELINE configs[3]: 500 loci, 24-species FBD-style tree, analytic populations -- 500/N per GPU
  config.legs.C3         BASELINE configs[2]: 200 loci x 800 bp x 60 tips, GTR+G4, UCLN, LinearWithConstantRoot (N=1 only)
  roofline.shapes.{C2,C3,C5s,C5}  the pruning kernel's CUDA-event time on each shape, same accounting as `roofline`
Every leg is checked against the CPU oracle on the GLOBAL problem (rank 0, all host threads): the run FAILS (exit 1) when
|logP_gpu - logP_oracle| / |logP_oracle| > 1e-9.

  value : pattern x node x category partials per second, whole job, inputs resident in HBM (device-timed)
  e2e   : the same metric through the C ABI with HOST buffers: species tree, all gene trees and parameters are
          re-sent every step (H2D inside the timed region) and the logP vector is read back (D2H)
  roofline : the pruning kernel (k_treewalk) alone, timed live with CUDA events on its stream
  cpu_baseline : the C restatement of the reference's Java path (oracle/), 1 thread, same workload, rank 0 only

`--impl reference` times that CPU restatement on all host threads instead (no JVM exists in this image, so the
reference's own code cannot run), on the same global problem the GPU arm evaluates at this N.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pattern·node·cat partials/sec (full multi-locus posterior evaluation, all nodes dirty)"
UNIT = "partials/s"
WORKLOAD = "C2"
L2_FLUSH_BYTES = 512 << 20
PARITY_TOL = 1e-9
LEG_MIN_STEPS = 200


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def design_bytes(pb) -> float:
    """Compulsory HBM bytes of one full evaluation in THIS design (children are consumed from registers / shared memory, so
    partials are written once and never re-read): 32 B partials + 2 B exponent per unit, tip states once,
    pattern log-likelihoods once.  DESIGN.md section 5 and BASELINE.md derive it."""
    b = 0.0
    for l in pb.loci:
        n_int = l.n_tips - 1
        b += 34.0 * l.n_pat * l.n_cat * n_int + l.n_pat * l.n_tips + 8.0 * l.n_pat
    return b


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under the bench load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# problems
# ----------------------------------------------------------------------------------------------------------------------

def rank_range(total: int, rank: int, world: int):
    """Contiguous block of loci of rank `rank` (equal counts: the synthetic loci of one config are statistically alike)."""
    return (total * rank) // world, (total * (rank + 1)) // world


def build_problem(workload: str, total: int, lo: int, hi: int, procs: int):
    from starbeast2_b200 import synth
    return synth.make_problem_parallel(workload, n_loci=total, locus_range=(lo, hi), procs=procs)


def canis_problem():
    """BASELINE configs[0] (C1): the tutorial's 16 alignments (tests/golden/canis_c1.json, generated from the reference's FASTA
    files by tests/golden/make_canis.py), 16 individuals / 8 species, HKY, constant populations, strict clock; the trees are
    random MSC draws (the tutorial starts from random trees too)."""
    from starbeast2_b200 import synth
    from starbeast2_b200.problem import POP_CONSTANT, Locus, Problem, hky_params
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "canis_c1.json")))["loci"]
    species = sorted({t[:-2] for v in fx.values() for t in v["taxa"]})
    rng = np.random.Generator(np.random.PCG64(20260101))
    sp = synth.yule_species_tree(rng, len(species))
    pop = rng.gamma(2.0, 0.025, size=sp.n_nodes)
    code = {c: i for i, c in enumerate("ACGT-")}
    loci = []
    for v in fx.values():
        tree, tip_species = synth.simulate_gene_tree(rng, sp, pop * 2.0, np.array([2] * len(species)))
        order = {}
        for k, s in enumerate(tip_species):   # gene tip k belongs to species s: take that species' next individual of the alignment
            order.setdefault(int(s), []).append(k)
        rows = np.zeros((len(tip_species), len(v["weights"])), np.uint8)
        seen = {}
        for ti, taxon in enumerate(v["taxa"]):
            s = species.index(taxon[:-2])
            k = order[s][seen.get(s, 0)]
            seen[s] = seen.get(s, 0) + 1
            rows[k] = [code[c] for c in v["patterns"][ti]]
        loci.append(Locus(tree=tree, tip_states=rows, weights=np.array(v["weights"], np.int32), tip_species=tip_species,
                          subst_params=hky_params(2.0, [0.25] * 4), clock_rate=0.05))
    return Problem(species=sp, loci=loci, pop_kind=POP_CONSTANT, pop_a=pop, species_rates=np.ones(sp.n_nodes))


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle as the reported baseline, never as the thing measured on the GPU side)
# ----------------------------------------------------------------------------------------------------------------------

def cpu_baseline(pb, n_threads: int, budget_s: float):
    import oracle
    fp = oracle.FlatProblem(pb)
    oracle.posterior(pb, n_threads=n_threads, flat=fp)   # warm-up (thread pool, page faults)
    n, t0 = 0, time.perf_counter()
    while True:
        r = oracle.posterior(pb, n_threads=n_threads, flat=fp)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 100000:
            break
    return n, dt, r


def cpu_incremental(pb, synth, n_steps=300):
    """Regime B on the host: the C restatement of what the reference evaluates when one gene tree changed (oracle
    IncrementalBaseline: embedding of EVERY locus -- GeneTree.requiresRecalculation() is always true --, population terms,
    clock rates and incremental TreeLikelihood of the changed locus), 1 thread like the reference's MCMC."""
    import oracle
    if pb.pop_kind not in (0, 1):
        return None
    rng = np.random.default_rng(12345)
    inc = oracle.IncrementalBaseline(pb)
    moves = []
    for _ in range(64):
        l = int(rng.integers(0, pb.n_loci))
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        moves.append((l, t, pb.loci[l].tree))
    for l, t, t0 in moves[:8]:
        inc.step(l, t); inc.step(l, t0)
    t_start = time.perf_counter()
    for i in range(n_steps):
        l, t, t0 = moves[i % len(moves)]
        inc.step(l, t)          # propose + evaluate
        inc.step(l, t0)         # the move back is an equally small update; both are counted as steps
    dt = time.perf_counter() - t_start
    inc.close()
    return {"steps_per_sec": 2 * n_steps / dt, "us_per_step": 1e6 * dt / (2 * n_steps), "cores": 1, "kind": "port",
            "note": "C restatement of the reference's per-step work (all-loci embedding + one-locus incremental likelihood), "
                    "gcc -O2 -ffp-contract=off; not the JVM"}


def cpu_operator_queries(pb, reps=20):
    """The same walks by the C restatement of the Java recursions (1 thread; the Java adds HashSet<String> lookups per leaf)."""
    import oracle
    sp = pb.species
    nodes = list(range(sp.n_leaves, sp.n_nodes - 1))
    cmap = np.arange(sp.n_nodes, dtype=np.int32)
    oracle.connecting_nodes(pb, nodes[0])
    t0 = time.perf_counter()
    for i in range(reps):
        oracle.connecting_nodes(pb, nodes[i % len(nodes)])
    t1 = time.perf_counter()
    for i in range(reps):
        oracle.max_height(pb, cmap, sp.n_leaves // 2)
    t2 = time.perf_counter()
    return {"connecting_nodes_us": 1e6 * (t1 - t0) / reps, "max_height_us": 1e6 * (t2 - t1) / reps,
            "note": "ctypes call incl. concatenating the trees on the host; C port of the recursion, 1 thread"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU legs
# ----------------------------------------------------------------------------------------------------------------------

def bench_incremental(eng, pb, synth, n_steps=400):
    """Regime B of SURVEY.md 3.3: one MCMC step = store, move one node of one gene tree (BEAST's Uniform operator),
    evaluate the posterior through the reference's two-phase protocol (coalescent first, likelihood only if finite),
    then reject (restore) or accept.  Host-timed, host buffers: this is latency, not throughput."""
    import math
    rng = np.random.default_rng(12345)
    moves = []
    for _ in range(64):
        l = int(rng.integers(0, pb.n_loci))
        t, node = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        moves.append((l, t))
    eng.eval_posterior()
    updates = []
    # through the C ABI with host buffers, addresses resolved once (what a JNI / FFM caller does; the Python convenience
    # wrappers add ~15 us of numpy argument handling per step that is not the engine's)
    from starbeast2_b200 import _lib as sblib
    raw, h = sblib.load_raw(), eng._h
    keep = []
    ptrs = []
    for l, t in moves:
        arrs = [np.ascontiguousarray(t.parent, np.int32), np.ascontiguousarray(t.left, np.int32), np.ascontiguousarray(t.right, np.int32),
                np.ascontiguousarray(t.height, np.float64)]
        keep.append(arrs)
        ptrs.append((l, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data))
    coal_buf, lik_buf, br_buf = np.zeros(eng.n_loci), np.zeros(eng.n_loci), np.zeros(pb.species.n_nodes)   # eng.n_loci: + the morphological partitions
    c1, l1, tot = np.zeros(1), np.zeros(1), np.zeros(3)
    pc, pl, pbr, pc1, pl1, pt = (a.ctypes.data for a in (coal_buf, lik_buf, br_buf, c1, l1, tot))

    def one(i, count=False):
        l, a, b, c, d = ptrs[i % len(ptrs)]
        rc = raw.sb2_store(h)
        rc |= raw.sb2_set_gene_trees(h, l, 1, a, b, c, d)
        rc |= raw.sb2_eval_coalescent(h, pc, pbr, pc1)
        if math.isfinite(c1[0]):
            rc |= raw.sb2_eval_likelihood(h, pl, pl1)
        if count:
            updates.append(eng.last_node_updates())
        rc |= raw.sb2_restore(h)
        if rc:
            eng._ck(-1)

    def one_call(i):
        l, a, b, c, d = ptrs[i % len(ptrs)]
        rc = raw.sb2_store(h)
        rc |= raw.sb2_set_gene_trees(h, l, 1, a, b, c, d)
        rc |= raw.sb2_eval_posterior(h, pc, pl, pbr, pt)      # speculative single round trip: coalescent and likelihood together
        rc |= raw.sb2_restore(h)
        if rc:
            eng._ck(-1)

    for i in range(20):
        one(i, count=True)
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for i in range(n_steps):
        one(i)
    dt = time.perf_counter() - t0
    launches = (eng.launch_count() - l0) / n_steps
    for i in range(20):
        one_call(i)
    t0 = time.perf_counter()
    for i in range(n_steps):
        one_call(i)
    dt1 = time.perf_counter() - t0
    return {"steps_per_sec": n_steps / dt, "us_per_step": 1e6 * dt / n_steps, "mean_node_updates": float(np.mean(updates)),
            "kernel_launches_per_step": launches,
            "protocol": "store, set_gene_tree(1 locus), eval_coalescent (likelihood evaluated speculatively in the same device pass), "
                        "eval_likelihood (cached), restore; host-timed",
            "single_round_trip": {"steps_per_sec": n_steps / dt1, "us_per_step": 1e6 * dt1 / n_steps,
                                  "protocol": "store, set_gene_tree, eval_posterior, restore"}}


def bench_operator_queries(eng, pb, reps=200):
    """the all-gene-tree walks of CoordinatedUniform/Exponential and NodeReheight2 as device queries,
    host-timed per call (host buffers in and out; latency, one proposal = one query)."""
    sp = pb.species
    nodes = list(range(sp.n_leaves, sp.n_nodes - 1))
    side = (np.arange(sp.n_leaves) % 2).astype(np.uint8)
    eng.eval_posterior()
    out = {}
    for name, fn in (("connecting_nodes_with_mask", lambda i: eng.connecting_nodes(nodes[i % len(nodes)])),
                     ("connecting_nodes_freedoms_only", lambda i: eng.connecting_nodes(nodes[i % len(nodes)], want_mask=False)),
                     ("max_height", lambda i: eng.max_height(side))):
        for i in range(10):
            fn(i)
        t0 = time.perf_counter()
        for i in range(reps):
            fn(i)
        out[name + "_us"] = 1e6 * (time.perf_counter() - t0) / reps
    out["gene_nodes_walked_per_query"] = int(sum(l.tree.n_nodes for l in pb.loci))
    return out


class Ctx:
    """Process-wide bench context: rank / world / torch handles / the L2-flush buffer."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libsb2gpu has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.hbm_gbs, self.peak_src = load_peaks()
        self.procs = max(1, min(16, host_threads() // self.world))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


def stats(ms):
    ms = np.asarray(ms, dtype=np.float64)
    return {"min": float(ms.min()), "median": float(np.median(ms)), "mean": float(ms.mean()), "max": float(ms.max()),
            "p90": float(np.percentile(ms, 90)), "n": int(ms.size)}


class Leg:
    """One workload on this rank's shard: engine, device-timed steps, e2e steps through the raw C ABI, roofline pass."""

    def __init__(self, ctx: Ctx, workload: str, loci_total: int, lo: int, hi: int, pb=None):
        from starbeast2_b200.engine import Engine
        self.ctx, self.workload, self.loci_total = ctx, workload, loci_total
        self.pb = pb if pb is not None else build_problem(workload, loci_total, lo, hi, ctx.procs)
        pbs = self.pb
        self.eng = Engine.from_problem(pbs, device=ctx.local, exact_order=ctx.args.exact, profile=False)
        self.red = self.eng.bind_torch()
        self.comm = "none"
        if ctx.world > 1:
            self.comm = "nccl"
            if os.environ.get("SB2_COMM", "peer") == "peer" and self.eng.connect_peers():
                self.comm = "peer-memory all-reduce inside the reduction kernel (NVLink stores, flag-in-data words)"
        self.peers = ctx.world > 1 and self.comm != "nccl"
        self.tot = np.zeros(3)
        self.units_local = pbs.units
        self._prepare_e2e()

    def close(self):
        self.eng.close()

    def step_device(self):
        ctx, eng = self.ctx, self.eng
        eng.mark_all_dirty()
        if ctx.world == 1 or self.peers:
            eng.eval_posterior_total(self.tot)
            return
        eng.eval_begin()
        ctx.dist.all_reduce(self.red)
        eng.eval_end_total(self.tot)

    def _prepare_e2e(self):
        """Flat host copies of everything a step of the reference-facing API re-sends, addresses resolved once (as a JNI /
        FFM caller would): species tree, species rates, population sizes, all gene trees in; per-locus + total logP out."""
        from starbeast2_b200 import _lib as sblib
        pb, eng = self.pb, self.eng
        trees = [l.tree for l in pb.loci]
        self._keep = [np.ascontiguousarray(np.concatenate([t.parent for t in trees]), np.int32),
                      np.ascontiguousarray(np.concatenate([t.left for t in trees]), np.int32),
                      np.ascontiguousarray(np.concatenate([t.right for t in trees]), np.int32),
                      np.ascontiguousarray(np.concatenate([t.height for t in trees]), np.float64)]
        n_nodes = int(self._keep[0].shape[0])
        S = pb.species.n_nodes
        self.h2d = 20 * n_nodes + 20 * S + 8 * S * 3 + 16 + pb.n_loci          # trees + species block + flags
        self.n_morph = len(getattr(eng, "morph", []))
        self.h2d += 20 * S * self.n_morph
        self.d2h = 8 * (3 + 2 * eng.n_loci + S)
        self._out = [np.zeros(eng.n_loci), np.zeros(eng.n_loci), np.zeros(S)]   # per-locus outputs cover the morphological partitions too
        sp = pb.species
        self._keep2 = [np.ascontiguousarray(x) for x in (sp.parent, sp.left, sp.right, sp.height, pb.species_rates,
                                                        pb.pop_a if pb.pop_a is not None else np.zeros(1),
                                                        pb.pop_b if pb.pop_b is not None else np.zeros(1))]
        self._raw, self._h = sblib.load_raw(), eng._h

    def step_e2e(self):
        raw, h, ctx = self._raw, self._h, self.ctx
        ptr = [a.ctypes.data for a in self._keep2]
        tp, tl, tr, th = (a.ctypes.data for a in self._keep)
        oc, ol, ob = (a.ctypes.data for a in self._out)
        ot = self.tot.ctypes.data
        pb = self.pb
        rc = raw.sb2_set_species_tree(h, ptr[0], ptr[1], ptr[2], ptr[3])
        rc |= raw.sb2_set_species_rates(h, ptr[4])
        rc |= raw.sb2_set_pop_model(h, pb.pop_kind, ptr[5], ptr[6], pb.alpha, pb.beta)
        rc |= raw.sb2_set_gene_trees(h, 0, pb.n_loci, tp, tl, tr, th)
        for j in range(self.n_morph):   # the morphological partitions live on the species tree: re-stage it for each of them
            rc |= raw.sb2_set_gene_trees(h, pb.n_loci + j, 1, ptr[0], ptr[1], ptr[2], ptr[3])
        rc |= raw.sb2_mark_all_dirty(h)
        if ctx.world == 1 or self.peers:
            rc |= raw.sb2_eval_posterior(h, oc, ol, ob, ot)
        else:
            rc |= raw.sb2_eval_begin(h)
            ctx.dist.all_reduce(self.red)
            rc |= raw.sb2_eval_end(h, oc, ol, ob, ot)
        if rc:
            self.eng._ck(-1)

    def timed(self, fn, steps, flush_l2=True, waits=None):
        """Per-step device times (ms), CUDA events on the engine's stream, L2 flushed before every step (outside the events)."""
        torch, ctx = self.ctx.torch, self.ctx
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            if flush_l2:
                ctx.flush.fill_(1)          # evict the working set; outside the timed events
            a.record(ctx.stream)
            fn()
            b.record(ctx.stream)
            if waits is not None:
                waits.append(self.eng.exchange_wait_us())
        torch.cuda.synchronize()
        return np.array([a.elapsed_time(b) for a, b in ev])

    def measure(self, steps, warmup, e2e_steps=None, roofline_steps=50):
        """Returns (per-step max-over-ranks device ms, e2e ms array, kernel us, launches, exchange-wait stats)."""
        ctx, eng = self.ctx, self.eng
        for _ in range(max(warmup, 3)):
            ctx.flush.fill_(1)
            self.step_device()
        ctx.barrier()
        l0 = eng.launch_count()
        waits = [] if self.peers else None
        ms = self.timed(self.step_device, steps, waits=waits)
        launches = eng.launch_count() - l0
        self.total_logp = float(self.tot[2])
        ctx.barrier()
        # roofline pass: the same steps with the pruning kernel bracketed by CUDA events on its own stream
        eng.kernel_time()
        eng.set_profile(True)
        self.timed(self.step_device, min(steps, roofline_steps))
        k_ms, k_n = eng.kernel_time()
        eng.set_profile(False)
        ctx.barrier()
        e2e_steps = steps if e2e_steps is None else e2e_steps
        for _ in range(3):
            self.step_e2e()
        ctx.barrier()
        ms_e2e = self.timed(self.step_e2e, e2e_steps)
        ctx.barrier()
        ms = ctx.max_over_ranks(ms)
        ms_e2e = ctx.max_over_ranks(ms_e2e)
        wait = None
        if waits:
            w = ctx.max_over_ranks(waits)
            wait = {"us_per_step": stats(w), "share_of_step": float(np.median(w) * 1e-3 / np.median(ms)),
                    "note": "device time (globaltimer) the reduction kernel spends between publishing its words and having summed "
                            "every peer's: rank skew + one NVLink trip; max over ranks per step"}
        return ms, ms_e2e, 1e3 * k_ms / max(k_n, 1), int(k_n), launches, wait

    def roofline(self, k_us, k_n):
        ctx, pb = self.ctx, self.pb
        dbytes, sbytes = design_bytes(pb), pb.algorithmic_bytes_survey()
        achieved = dbytes / (k_us * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": "k_treewalk", "achieved": achieved, "peak": ctx.hbm_gbs, "unit": "GB/s", "frac": achieved / ctx.hbm_gbs,
                "us_per_launch": k_us, "launches_timed": k_n, "algorithmic_bytes_per_launch": dbytes, "bytes_per_unit": dbytes / pb.units,
                "survey_formula_bytes_per_launch": sbytes, "survey_formula_frac": sbytes / (k_us * 1e-6) / 1e9 / ctx.hbm_gbs,
                "kernel_partials_per_sec": pb.units / (k_us * 1e-6), "loci_on_this_gpu": pb.n_loci, "units_on_this_gpu": pb.units}


def oracle_total(workload, loci_total, pb_local, ctx):
    """log posterior of the GLOBAL problem by the CPU oracle (rank 0 only; all host threads)."""
    import oracle
    oracle.build()
    pb = pb_local if ctx.world == 1 else build_problem(workload, loci_total, 0, loci_total, host_threads())
    t0 = time.perf_counter()
    r = oracle.posterior(pb, n_threads=host_threads())
    return r, time.perf_counter() - t0, pb


def run_leg(ctx: Ctx, name: str, workload: str, loci_total: int, steps: int, warmup: int, scaling: str, pb=None, check=True,
            keep_open=False):
    """One north-star leg: loci_total loci split over the ranks, device-timed per-step stats (max over ranks), e2e, the
    pruning kernel's roofline on this rank's shard and the oracle check of the global log posterior."""
    lo, hi = rank_range(loci_total, ctx.rank, ctx.world)
    leg = Leg(ctx, workload, loci_total, lo, hi, pb=pb)
    ms, ms_e2e, k_us, k_n, launches, wait = leg.measure(steps, warmup, e2e_steps=min(steps, 50))
    units_total = float(ctx.sum_over_ranks([float(leg.units_local)])[0])
    out = None
    if ctx.rank == 0:
        med = float(np.median(ms))
        out = {"workload": workload, "loci_total": loci_total, "loci_per_gpu": hi - lo, "n_gpus": ctx.world, "scaling": scaling,
               "units_per_step": units_total, "steps": int(len(ms)), "ms_per_step": stats(ms),
               "partials_per_sec": units_total / (float(ms.mean()) * 1e-3), "posterior_evals_per_sec": 1e3 / float(ms.mean()),
               "partials_per_sec_median_step": units_total / (med * 1e-3),
               "e2e_ms_per_step": stats(ms_e2e), "e2e_partials_per_sec": units_total / (float(ms_e2e.mean()) * 1e-3),
               "gpu_launches_per_step": launches / len(ms), "log_posterior": leg.total_logp,
               "device_bytes_per_gpu": leg.eng.device_bytes(), "kernel": leg.roofline(k_us, k_n),
               "timing": "CUDA events per step on the engine's stream, L2 flushed (512 MiB write) before every step, per-step max over ranks"}
        if wait:
            out["exchange_wait"] = wait
        if check and not ctx.args.no_cpu:
            r, dt, _ = oracle_total(workload, loci_total, leg.pb, ctx)
            out["oracle_log_posterior"] = r.total
            out["log_posterior_rel_diff"] = abs(r.total - leg.total_logp) / abs(r.total)
            out["oracle_seconds"] = dt
            out["parity_ok"] = bool(out["log_posterior_rel_diff"] <= PARITY_TOL)
    if keep_open:
        return out, leg
    leg.close()
    return out, None


def run_reference(args):
    """`--impl reference`: the CPU restatement of the reference path (oracle/, kind "port"), all host threads, on the same
    global problem the GPU arm evaluates at this N (loci_per_gpu x N loci)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    total = args.loci * world
    threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to run on
    pb = build_problem(args.workload, total, 0, total, threads)
    fp = oracle.FlatProblem(pb)
    for _ in range(max(args.warmup, 1)):
        oracle.posterior(pb, n_threads=threads, flat=fp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = oracle.posterior(pb, n_threads=threads, flat=fp)
    dt = time.perf_counter() - t0
    units = pb.units
    v = units * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "loci_per_gpu": args.loci, "loci_total": total, "units_per_step": float(units),
                   "regime": "A (every node of every locus dirty)",
                   "note": "C restatement of the reference's Java path (no JVM in this image); one step = one full posterior "
                           "evaluation of the same global problem as the GPU arm at this N, on the host cores (rank 0 only)"},
        "posterior_evals_per_sec": args.steps / dt, "log_posterior": r.total,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full evaluations of {args.workload} ({total} loci), OpenMP over loci"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--loci", type=int, default=0, help="loci per GPU (default: the workload's own count)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip every CPU-oracle leg (baseline and parity checks)")
    ap.add_argument("--no-extra", action="store_true", help="headline leg only: skip the north-star legs, regime B and the operator queries")
    ap.add_argument("--legs", default="C5,C4,C3", help="north-star legs to run (comma separated subset of C5,C4,C3)")
    ap.add_argument("--exact", action="store_true")
    args = ap.parse_args()
    from starbeast2_b200 import synth
    if args.loci <= 0:
        args.loci = synth.CONFIGS[args.workload]["n_loci"]
    if args.impl == "reference":
        return run_reference(args)

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")            # both print "NCCL version ..." on stdout: keep stdout to the one JSON line
    from starbeast2_b200 import build as sbuild
    ctx = Ctx(args)
    sbuild.build()
    rank, world = ctx.rank, ctx.world

    # ---- headline leg: C2, weak scaling ----------------------------------------------------------------------------------
    total = args.loci * world
    lo, hi = rank * args.loci, (rank + 1) * args.loci
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    head = Leg(ctx, args.workload, total, lo, hi)
    t_wall0 = time.time()
    ms, ms_e2e, k_us, k_n, launches, wait = head.measure(args.steps, args.warmup)
    # keep the load up long enough for nvidia-smi to see it (100 ms sampling), same step, untimed
    t_probe = time.time()
    while time.time() - t_probe < 1.2:
        head.step_device()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    units_total = float(ctx.sum_over_ranks([float(head.units_local)])[0])
    launches_total = int(ctx.sum_over_ranks([float(launches)])[0])
    pb, eng = head.pb, head.eng

    line = None
    failed = []
    if rank == 0:
        roof = head.roofline(k_us, k_n)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        roof.update({"traffic": traffic, "peak_source": ctx.peak_src,
                     "accounting": "compulsory bytes of this design: 34 B per unit written once (32 B partials + 2 B exponent) + tip states + "
                                   "pattern logL; SURVEY 8(d)'s 68.5-79.9 B/unit charges child re-reads that do not happen (survey_formula_*; "
                                   "see BASELINE.md)"})
        ms_sum = float(ms.sum())
        line = {
            "metric": METRIC, "value": units_total * len(ms) / (ms_sum * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_sum / len(ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "loci_per_gpu": args.loci, "loci_total": total,
                       "units_per_step": units_total, "mean_patterns_per_locus": float(np.mean([l.n_pat for l in pb.loci])),
                       "regime": "A (every node of every locus dirty)", "device_bytes": eng.device_bytes(), "exact_order": bool(args.exact),
                       "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write; working set is L2-resident)",
                       "ms_per_step_stats": stats(ms),
                       "parallelism": f"loci sharded over {world} GPU(s), 1 all-reduce of {eng.reduce_vector_len()} doubles per step ({head.comm})" if world > 1 else "1 GPU"},
            "posterior_evals_per_sec": len(ms) / (ms_sum * 1e-3),
            "log_posterior": head.total_logp,
            "e2e": {"value": units_total * len(ms_e2e) / (float(ms_e2e.sum()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": head.h2d * world,
                    "d2h_bytes_per_step": head.d2h * world, "ms_per_step": float(ms_e2e.mean()), "ms_per_step_stats": stats(ms_e2e),
                    "posterior_evals_per_sec": len(ms_e2e) / (float(ms_e2e.sum()) * 1e-3)},
            "gpu_launches": launches_total,
            "roofline": roof,
            "clocks": clocks,
        }
        if wait:
            line["config"]["exchange_wait"] = wait
        # parity of the global log posterior at every N (the in-kernel all-reduce needs a driver-run check)
        if not args.no_cpu:
            r, dt, pb_glob = oracle_total(args.workload, total, pb, ctx)
            rel = abs(r.total - head.total_logp) / abs(r.total)
            line["e2e"]["log_posterior_rel_diff_vs_oracle"] = rel
            line["config"]["log_posterior_rel_diff_vs_oracle"] = rel
            line["config"]["oracle_log_posterior"] = r.total
            if rel > PARITY_TOL:
                failed.append(f"{args.workload} x{world}: rel diff {rel:.3e}")
            if world == 1:
                import oracle
                n, dt, r1 = cpu_baseline(pb, 1, args.cpu_seconds)
                line["cpu_baseline"] = {"value": pb.units * n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"{n} full evaluations of {args.workload} ({args.loci} loci) in {dt:.1f} s, single thread "
                                                  "(the reference runs threads: 1); C restatement of the Java path, gcc -O2 -ffp-contract=off",
                                        "posterior_evals_per_sec": n / dt,
                                        "log_posterior_rel_diff": abs(r1.total - head.total_logp) / abs(r1.total)}
            else:
                line["cpu_baseline"] = None
        line["roofline"]["shapes"] = {args.workload: {k: roof[k] for k in ("frac", "achieved", "us_per_launch", "algorithmic_bytes_per_launch",
                                                                            "bytes_per_unit", "loci_on_this_gpu", "kernel_partials_per_sec")}}

    # ---- regime B + operator queries on the headline shape (N=1) ---------------------------------------------------------------
    if world == 1 and not args.no_extra:
        line["incremental"] = bench_incremental(eng, pb, synth)
        line["operator_queries"] = bench_operator_queries(eng, pb)
        if not args.no_cpu:
            line["operator_queries"]["cpu_port"] = cpu_operator_queries(pb)
            line["incremental"]["cpu_port"] = cpu_incremental(pb, synth)
        line["config"]["regime_B"] = {args.workload: {k: line["incremental"][k] for k in ("us_per_step", "steps_per_sec", "mean_node_updates", "kernel_launches_per_step")}}
        if "cpu_port" in line["incremental"] and line["incremental"]["cpu_port"]:
            line["config"]["regime_B"][args.workload]["cpu_port_us_per_step"] = line["incremental"]["cpu_port"]["us_per_step"]
    head.close()

    # ---- north-star legs -----------------------------------------------------------------------------------------------------
    if not args.no_extra:
        legs = {}
        want = [x.strip() for x in args.legs.split(",") if x.strip()]
        leg_steps = max(args.steps, LEG_MIN_STEPS)
        if "C5" in want:
            n5 = synth.CONFIGS["C5"]["n_loci"]
            out, leg = run_leg(ctx, "C5_strong", "C5", n5, leg_steps, 5, "strong", keep_open=(world == 1))
            if rank == 0:
                legs["C5_strong"] = out
                line["roofline"]["shapes"]["C5" if world == 1 else f"C5/{world}"] = _shape(out["kernel"])
            if world == 1:
                # the per-GPU shard of the 8-GPU config (250 loci) on one GPU: the first 250 loci of the same problem
                from starbeast2_b200.problem import Problem
                pbs = leg.pb
                shard = Problem(species=pbs.species, loci=pbs.loci[:n5 // 8], species_rates=pbs.species_rates, pop_kind=pbs.pop_kind,
                                pop_a=pbs.pop_a, pop_b=pbs.pop_b, alpha=pbs.alpha, beta=pbs.beta)
                leg.close()
                out_s, leg_s = run_leg(ctx, "C5_shard", "C5", n5 // 8, leg_steps, 5, "n/a", pb=shard, keep_open=True)
                legs["C5_shard_250"] = out_s
                line["roofline"]["shapes"]["C5s"] = _shape(out_s["kernel"])
                inc = bench_incremental(leg_s.eng, shard, synth, n_steps=200)
                q = bench_operator_queries(leg_s.eng, shard, reps=100)
                if not args.no_cpu:
                    q["cpu_port"] = cpu_operator_queries(shard, reps=5)
                    inc["cpu_port"] = cpu_incremental(shard, synth, n_steps=60)
                line["roofline_streaming"] = {"incremental": inc, "operator_queries": q}
                line["config"]["regime_B"]["C5s"] = {k: inc[k] for k in ("us_per_step", "steps_per_sec", "mean_node_updates", "kernel_launches_per_step")}
                if inc.get("cpu_port"):
                    line["config"]["regime_B"]["C5s"]["cpu_port_us_per_step"] = inc["cpu_port"]["us_per_step"]
                leg_s.close()
        if "C4" in want:
            out, _ = run_leg(ctx, "C4_strong", "C4", synth.CONFIGS["C4"]["n_loci"], leg_steps, 5, "strong")
            if rank == 0:
                legs["C4_strong"] = out
        if "C3" in want and world == 1:
            out, _ = run_leg(ctx, "C3", "C3", synth.CONFIGS["C3"]["n_loci"], leg_steps, 5, "n/a")
            legs["C3"] = out
            line["roofline"]["shapes"]["C3"] = _shape(out["kernel"])
        if world == 1:
            # BASELINE configs[0]: the tutorial data, regime B (what the tutorial's 67-133 k steps/s on the JVM measure)
            c1 = canis_problem()
            from starbeast2_b200.engine import Engine
            e1 = Engine.from_problem(c1, device=ctx.local)
            inc = bench_incremental(e1, c1, synth, n_steps=400)
            if not args.no_cpu:
                inc["cpu_port"] = cpu_incremental(c1, synth, n_steps=300)
                import oracle
                inc["log_posterior_rel_diff"] = abs(oracle.posterior(c1).total - e1.eval_posterior().total) / abs(oracle.posterior(c1).total)
            e1.close()
            line["config"]["regime_B"]["C1"] = {k: inc[k] for k in ("us_per_step", "steps_per_sec", "mean_node_updates", "kernel_launches_per_step")}
            line["config"]["regime_B"]["C1"]["reference_jvm_steps_per_sec"] = "67,000-133,000 (tutorial/StarBEAST2-tutorial.tex:378,419-421; other hardware, not measured here)"
            if inc.get("cpu_port"):
                line["config"]["regime_B"]["C1"]["cpu_port_us_per_step"] = inc["cpu_port"]["us_per_step"]
            line["incremental_C1"] = inc
        if rank == 0:
            line["config"]["legs"] = legs
            for k, v in legs.items():
                if v and v.get("parity_ok") is False:
                    failed.append(f"{k} x{world}: rel diff {v['log_posterior_rel_diff']:.3e}")

    if rank == 0:
        line["config"]["parity"] = {"tolerance": PARITY_TOL, "failed": failed,
                                    "checked": "global log posterior of every leg vs the CPU oracle (rank 0)" if not args.no_cpu else "skipped (--no-cpu)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()
    if failed:
        raise SystemExit("parity check failed: " + "; ".join(failed))


def _shape(k):
    return {x: k[x] for x in ("frac", "achieved", "us_per_launch", "algorithmic_bytes_per_launch", "bytes_per_unit", "loci_on_this_gpu",
                              "kernel_partials_per_sec")}


if __name__ == "__main__":
    main()
