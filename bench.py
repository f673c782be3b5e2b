#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200.

A "step" is one full posterior evaluation with every node dirty (SURVEY.md 8d regime A): multispecies-coalescent
embedding + per-branch terms + species-clock rates + transition matrices + Felsenstein pruning of every gene tree
+ root integration, for all loci of the workload.  N=1 workload: BASELINE.json configs[1] ("C2": 50 loci x
1,000 bp x 30 individuals / 10 species, HKY, ConstantPopulations).  N>1: weak scaling, every rank holds its own
50-locus shard of one 50*N-locus posterior and each step all-reduces the reduce vector over NCCL.

  value : pattern x node x category partials per second, whole job, inputs resident in HBM (device-timed)
  e2e   : the same metric through the C ABI with HOST buffers: species tree, all gene trees and parameters are
          re-sent every step (H2D inside the timed region) and the logP vector is read back (D2H)
  roofline : the pruning kernel (k_treewalk) alone, timed live with CUDA events on its stream
  cpu_baseline : the C restatement of the reference's Java path (oracle/), 1 thread, same workload, rank 0 only

`--impl reference` times that CPU restatement on all host threads instead (no JVM exists in this image, so the
reference's own code cannot run).
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def design_bytes(pb) -> float:
    """Compulsory HBM bytes of one full evaluation in THIS design (children are consumed from shared memory, so
    partials are written once and never re-read): 32 B partials + 2 B exponent per unit, tip states once,
    pattern log-likelihoods once.  DESIGN.md derives it."""
    b = 0.0
    for l in pb.loci:
        n_int = l.n_tips - 1
        b += 34.0 * l.n_pat * l.n_cat * n_int + l.n_pat * l.n_tips + 8.0 * l.n_pat
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under the bench load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index
        self.marks = []

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


def build_problem(rank: int, world: int, loci_per_rank: int, workload: str):
    from starbeast2_b200 import synth
    total = loci_per_rank * world
    return synth.make_problem(workload, n_loci=total, locus_range=(rank * loci_per_rank, (rank + 1) * loci_per_rank))


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
    coal_buf, lik_buf, br_buf = np.zeros(pb.n_loci), np.zeros(pb.n_loci), np.zeros(pb.species.n_nodes)
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
    t0 = time.perf_counter()
    for i in range(n_steps):
        one(i)
    dt = time.perf_counter() - t0
    for i in range(20):
        one_call(i)
    t0 = time.perf_counter()
    for i in range(n_steps):
        one_call(i)
    dt1 = time.perf_counter() - t0
    return {"steps_per_sec": n_steps / dt, "us_per_step": 1e6 * dt / n_steps, "mean_node_updates": float(np.mean(updates)),
            "protocol": "store, set_gene_tree(1 locus), eval_coalescent (likelihood evaluated speculatively in the same device pass), "
                        "eval_likelihood (cached), restore; host-timed",
            "single_round_trip": {"steps_per_sec": n_steps / dt1, "us_per_step": 1e6 * dt1 / n_steps,
                                  "protocol": "store, set_gene_tree, eval_posterior, restore"}}


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


def bench_streaming(args, hbm_gbs, peak_src, flush, n_loci=250, steps=30):
    """The pruning kernel where it actually streams HBM: a C5-shaped shard (GTR+G4, 120 individuals / 40 species, ~450
    patterns per locus; 250 loci = the per-GPU shard of the 8-GPU config, 1.7 GB of partials per evaluation).  Same accounting as the headline `roofline`.
    Not a bench line of its own (the C2 workload above is): it documents what bounds the kernel once the GPU is full."""
    import torch
    from starbeast2_b200 import synth
    from starbeast2_b200.engine import Engine
    pb = synth.make_problem("C5", n_loci=n_loci)
    eng = Engine.from_problem(pb, device=torch.cuda.current_device(), exact_order=args.exact, profile=True)
    eng.bind_torch()
    tot = np.zeros(3)
    for _ in range(3):
        eng.mark_all_dirty(); eng.eval_posterior_total(tot)
    eng.kernel_time()
    stream = torch.cuda.current_stream()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        eng.mark_all_dirty(); eng.eval_posterior_total(tot)
        b.record(stream)
    torch.cuda.synchronize()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    k_ms, k_n = eng.kernel_time()
    k_us = 1e3 * k_ms / max(k_n, 1)
    dbytes, sbytes = design_bytes(pb), pb.algorithmic_bytes_survey()
    achieved = dbytes / (k_us * 1e-6) / 1e9
    eng.set_profile(False)
    incremental = bench_incremental(eng, pb, synth, n_steps=200)
    queries = bench_operator_queries(eng, pb, reps=100)
    if not args.no_cpu:
        queries["cpu_port"] = cpu_operator_queries(pb, reps=5)
        incremental["cpu_port"] = cpu_incremental(pb, synth, n_steps=60)
    dev_bytes = eng.device_bytes()
    eng.close()
    return {"operator_queries": queries, "incremental": incremental, "device_bytes": dev_bytes, "workload": f"C5 shard, {n_loci} loci of the 2,000 (GTR+G4, UCLN species clock, 120 tips)", "kernel": "k_treewalk",
            "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "peak_source": peak_src,
            "us_per_launch": k_us, "launches_timed": k_n, "algorithmic_bytes_per_launch": dbytes, "bytes_per_unit": dbytes / pb.units,
            "survey_formula_frac": sbytes / (k_us * 1e-6) / 1e9 / hbm_gbs, "kernel_partials_per_sec": pb.units / (k_us * 1e-6),
            "step_ms": step_ms, "step_partials_per_sec": pb.units / (step_ms * 1e-3)}


def run_reference(args):
    """`--impl reference`: the CPU restatement of the reference path (oracle/, kind "port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    pb = build_problem(0, 1, args.loci, args.workload)
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to run on
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    fp = oracle.FlatProblem(pb)
    for _ in range(max(args.warmup, 1)):
        oracle.posterior(pb, n_threads=threads, flat=fp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.posterior(pb, n_threads=threads, flat=fp)
    dt = time.perf_counter() - t0
    units = pb.units
    v = units * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "loci": args.loci, "units_per_step": units,
                   "note": "C restatement of the reference's Java path (no JVM in this image); one step = one full "
                           "posterior evaluation of the N=1 workload on the host cores (rank 0 only)"},
        "posterior_evals_per_sec": args.steps / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full evaluations of {args.workload} ({args.loci} loci), OpenMP over loci"},
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C5-shard streaming measurement of the pruning kernel")
    ap.add_argument("--exact", action="store_true")
    args = ap.parse_args()
    from starbeast2_b200 import synth
    if args.loci <= 0:
        args.loci = synth.CONFIGS[args.workload]["n_loci"]
    if args.impl == "reference":
        return run_reference(args)

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
    import torch
    from starbeast2_b200 import build as sbuild
    from starbeast2_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsb2gpu has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sbuild.build()

    pb = build_problem(rank, world, args.loci, args.workload)
    eng = Engine.from_problem(pb, device=local, exact_order=args.exact, profile=True)
    red = eng.bind_torch()
    comm = "none"
    if world > 1:
        comm = "nccl"
        if os.environ.get("SB2_COMM", "peer") == "peer" and eng.connect_peers():
            comm = "peer-memory all-reduce inside the reduction kernel (NVLink stores + flags)"
    peers = world > 1 and comm != "nccl"
    units_local = pb.units
    stream = torch.cuda.current_stream()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    tot = np.zeros(3)

    def step_device():
        eng.mark_all_dirty()
        if world == 1 or peers:
            eng.eval_posterior_total(tot)
            return
        eng.eval_begin()
        dist.all_reduce(red)
        eng.eval_end_total(tot)

    # flat host copies of everything a step of the reference-facing API re-sends
    trees = [l.tree for l in pb.loci]
    hp = np.ascontiguousarray(np.concatenate([t.parent for t in trees]), np.int32)
    hl = np.ascontiguousarray(np.concatenate([t.left for t in trees]), np.int32)
    hr = np.ascontiguousarray(np.concatenate([t.right for t in trees]), np.int32)
    hh = np.ascontiguousarray(np.concatenate([t.height for t in trees]), np.float64)
    n_nodes = int(hp.shape[0])
    S = pb.species.n_nodes
    h2d = 20 * n_nodes + 20 * S + 8 * S * 3 + 16 + pb.n_loci          # trees + species block + flags
    d2h = 8 * (3 + 2 * pb.n_loci + S)
    coal_buf, lik_buf, br_buf = np.zeros(pb.n_loci), np.zeros(pb.n_loci), np.zeros(S)

    # the reference-facing call sequence of one step, through the C ABI with HOST buffers (addresses resolved once, as a
    # JNI / FFM caller would): species tree, species rates, population sizes, all gene trees in; per-locus + total logP out
    from starbeast2_b200 import _lib as sblib
    raw, h = sblib.load_raw(), eng._h
    sp = pb.species
    keep = [np.ascontiguousarray(x) for x in (sp.parent, sp.left, sp.right, sp.height, pb.species_rates,
                                              pb.pop_a if pb.pop_a is not None else np.zeros(1), pb.pop_b if pb.pop_b is not None else np.zeros(1))]
    ptr = [a.ctypes.data for a in keep]
    tp, tl, tr, th = hp.ctypes.data, hl.ctypes.data, hr.ctypes.data, hh.ctypes.data
    oc, ol, ob, ot = coal_buf.ctypes.data, lik_buf.ctypes.data, br_buf.ctypes.data, tot.ctypes.data
    n_loci, pop_kind, alpha, beta = pb.n_loci, pb.pop_kind, pb.alpha, pb.beta

    def step_e2e():
        rc = raw.sb2_set_species_tree(h, ptr[0], ptr[1], ptr[2], ptr[3])
        rc |= raw.sb2_set_species_rates(h, ptr[4])
        rc |= raw.sb2_set_pop_model(h, pop_kind, ptr[5], ptr[6], alpha, beta)
        rc |= raw.sb2_set_gene_trees(h, 0, n_loci, tp, tl, tr, th)
        rc |= raw.sb2_mark_all_dirty(h)
        if world == 1 or peers:
            rc |= raw.sb2_eval_posterior(h, oc, ol, ob, ot)
        else:
            rc |= raw.sb2_eval_begin(h)
            dist.all_reduce(red)
            rc |= raw.sb2_eval_end(h, oc, ol, ob, ot)
        if rc:
            eng._ck(-1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_l2=True):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            if flush_l2:
                flush.fill_(1)          # evict the (L2-resident) working set; outside the timed events
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev)   # ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step_device()
    eng.kernel_time()
    eng.set_profile(False)        # no events between our launches in the timed region (they defeat dependent launch)
    barrier()
    l0 = eng.launch_count()
    t_wall0 = time.time()
    ms = timed(step_device, args.steps)
    launches = eng.launch_count() - l0
    total_ref = float(tot[2])
    barrier()
    # roofline pass: the same steps with the pruning kernel bracketed by CUDA events on its own stream
    eng.set_profile(True)
    timed(step_device, min(args.steps, 100))
    k_ms, k_n = eng.kernel_time()
    eng.set_profile(False)
    barrier()
    # e2e
    for _ in range(3):
        step_e2e()
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    # keep the load up long enough for nvidia-smi to see it (100 ms sampling), same step, untimed
    t_probe = time.time()
    while time.time() - t_probe < 1.2:
        step_device()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    t = torch.tensor([ms, ms_e2e, float(units_local), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(tmax[0]), float(tmax[1])
        units_total, launches_total = float(tsum[2]), int(tsum[3])
    else:
        units_total, launches_total = float(units_local), launches
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, peak_src = load_peaks()
    k_us = 1e3 * k_ms / max(k_n, 1)
    dbytes = design_bytes(pb)
    sbytes = pb.algorithmic_bytes_survey()
    achieved = dbytes / (k_us * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    value = units_total * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "loci_per_gpu": args.loci, "loci_total": args.loci * world,
                   "units_per_step": units_total, "mean_patterns_per_locus": float(np.mean([l.n_pat for l in pb.loci])),
                   "regime": "A (every node of every locus dirty)", "device_bytes": eng.device_bytes(), "exact_order": bool(args.exact),
                   "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write; working set is L2-resident)",
                   "parallelism": f"loci sharded over {world} GPU(s), 1 all-reduce of {eng.reduce_vector_len()} doubles per step ({comm})" if world > 1 else "1 GPU"},
        "posterior_evals_per_sec": args.steps / (ms * 1e-3),
        "log_posterior": total_ref,
        "e2e": {"value": units_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "posterior_evals_per_sec": args.steps / (ms_e2e * 1e-3)},
        "gpu_launches": launches_total,
        "roofline": {"bound": "hbm", "kernel": "k_treewalk", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                     "traffic": traffic, "peak_source": peak_src, "us_per_launch": k_us, "launches_timed": k_n,
                     "algorithmic_bytes_per_launch": dbytes, "bytes_per_unit": dbytes / pb.units,
                     "survey_formula_bytes_per_launch": sbytes, "survey_formula_GBps": sbytes / (k_us * 1e-6) / 1e9,
                     "survey_formula_frac": sbytes / (k_us * 1e-6) / 1e9 / hbm_gbs,
                     "kernel_partials_per_sec": pb.units / (k_us * 1e-6)},
        "clocks": clocks,
    }
    if world == 1:
        line["incremental"] = bench_incremental(eng, pb, synth)
        line["operator_queries"] = bench_operator_queries(eng, pb)
        if not args.no_cpu:
            line["operator_queries"]["cpu_port"] = cpu_operator_queries(pb)
            line["incremental"]["cpu_port"] = cpu_incremental(pb, synth)
        if not args.no_extra:
            eng.close()
            line["roofline_streaming"] = bench_streaming(args, hbm_gbs, peak_src, flush)
    if not args.no_cpu and world == 1:
        import oracle
        oracle.build()
        n, dt, r = cpu_baseline(pb, 1, args.cpu_seconds)
        line["cpu_baseline"] = {"value": pb.units * n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{n} full evaluations of {args.workload} ({args.loci} loci) in {dt:.1f} s, single thread "
                                          "(the reference runs threads: 1); C restatement of the Java path, gcc -O2 -ffp-contract=off",
                                "posterior_evals_per_sec": n / dt,
                                "log_posterior_rel_diff": abs(r.total - total_ref) / abs(r.total)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
