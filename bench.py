#!/usr/bin/env python
"""Benchmark of the force-evaluation hot path (contract: see the task statement / DESIGN.md "Measurement").

Workload (BASELINE.json configs[1], "C2"): bcc Fe 20x20x20 (a = 2.8553), one vacancy, one H -> 16 000 atoms; dimer saddle
searches started from Master::perturb-style states around the H atom, 64 of them resident on a GPU at any time.
One *step* = two environments' worth of searches on every GPU: a queue of 512 start states per GPU (the reference's find_n
launches batches until max_searches = 250 are done per environment, src/saddle/find.cpp:374-422, find.hpp:171, and
Master::find_mechs runs all new environments of a KMC step concurrently, find.cpp:249-259) streams through the 64 resident
slots; a slot whose search has finished claims the next start state on the device (ofly_dimer_find_sp_stream), so there is
no barrier after every 64. All of it -- every neighbour rebuild/update, EAM gradient, rotor and L-BFGS update -- is ONE
CUDA-graph launch whose body (a tick: vector algebra of every active replica, rebuilds, one gradient evaluation) repeats on
the device.

  value  = EAM atom-gradient evaluations / s inside the step, inputs resident in HBM (ofly_dimer_find_sp_stream_dev)
  e2e    = same metric through the host-pointer C-ABI call (ofly_dimer_find_sp_stream): pinned host inputs are copied
           to the device and positions/axes/exit codes are copied back inside the timed region
  roofline: the dominant kernels, k_eam_rho + k_eam_force (one gradient evaluation of the resident batch), timed
           live with CUDA events on the library's stream; algorithmic bytes (8 n + 64 per atom-gradient eval, n =
           measured mean stored neighbours) against MEASURED_PEAKS.json's HBM figure
  cpu_baseline / --impl reference: the CPU oracle built with the reference's flags, all host cores, one
           single-threaded search per core exactly like saddle::Master (src/saddle/find.cpp:698,990)

N > 1 (torchrun): ONE global queue of N x 512 searches is sharded by search id (global id g runs on rank g mod N,
openfly_b200/shard.py), every rank streams its shard through its own GPU, the finished records are merged on the host in id
order (shard.merge_records) -- no collective on the data path; per-GPU work is fixed as N grows (weak scaling) but every
rank runs DIFFERENT searches. A strong-scaling leg rides along: BASELINE configs[4], 2048 searches on the 431-atom cell
sharded over the N GPUs ("c5_strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEN = os.path.join(ROOT, "data", "wen.eam.fs")
CELLS, LATTICE, SLOTS, QUEUE = 20, 2.8553, 64, 512
C5_TOTAL, C5_SLOTS = 2048, 1024
METRIC, UNIT = "eam_atom_gradient_evals_per_s", "atom-evals/s"
WORKLOAD = ("C2: bcc Fe 20^3 (a=2.8553) - 1 vacancy + 1 H = 16000 atoms, Wen Fe-H EAM, dimer searches (Dimer::find_sp, default options) "
            "from Master::perturb start states: 64 concurrent searches per GPU, 512 queued per GPU and step")
CPU_SAMPLE_MAX_STEPS = 40  # translation steps per CPU search in the bounded sample (a full search averages ~45; shorter cuts flatter the CPU: the first steps rarely rebuild the list)


def config_dict(seed: int):
    """The same object in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "atoms": 16000, "resident_searches_per_gpu": SLOTS, "queued_searches_per_gpu_per_step": QUEUE,
            "multi_gpu": "one global queue of n_gpus x 512 searches sharded by id (g mod n_gpus), no data-path collective, records merged on the host",
            "l2": "inputs larger than L2: neighbour rows 64x16000x104x4 B = 426 MB + 1.1 GB of search vectors vs 126 MB L2; no explicit flush",
            "seed": seed}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(geometry: dict):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch pair, from the committed ncu --set full capture of the SAME
    launch geometry (profiles/pair_traffic.json); None when the geometry of this run differs from the captured one."""
    path = os.path.join(ROOT, "profiles", "pair_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None, "no capture committed"
    if all(rec.get("geometry", {}).get(k) == v for k, v in geometry.items()):
        return rec["dram_bytes_per_launch_pair"], rec.get("source", path)
    return None, f"capture geometry {rec.get('geometry')} differs from this run's {geometry}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu,clocks.mem")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, util, mem, pw, reasons = [], [], [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
                if len(f) > 8:
                    util.append(float(f[8]))
                if len(f) > 9:
                    mem.append(float(f[9]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "gpu_util_mean": (float(np.mean(util)) if util else None), "sm_mhz_min": float(min(sm)),
                "mem_mhz_min": (float(min(mem)) if mem else None), "power_w_max": (float(max(pw)) if pw else None)}


def c2_cell():
    from openfly_b200 import inputs

    return inputs.iron_hydrogen_cell(CELLS, LATTICE, remove=(1,), hydrogen=True)


def start_states(rmin, ex, centre, gids, seed):
    """Start state of global search g: the independent stream (seed, g) -- the same whatever rank draws it."""
    from openfly_b200 import inputs

    P = np.empty((len(gids),) + rmin.shape)
    A = np.empty_like(P)
    for k, g in enumerate(gids):
        P[k], A[k] = inputs.perturb(rmin, ex, centre, np.random.default_rng([seed, int(g)]))
    return P, A


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C

    import torch

    from openfly_b200 import capi, fly, inputs, shard

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    ctx = fly.Context(local)
    lib = ctx.lib
    data = fly.DataEAM(ctx, WEN)
    pot = fly.Generic(data, type_map=["Fe", "H"])
    pos, ex, ids = c2_cell()
    N, R, Q = len(ids), SLOTS, QUEUE
    box = fly.Box(tuple(ex))

    # set-up (untimed): relax the cell, draw this rank's shard of the global queue
    s1 = fly.System(data, box, ids, replicas=1)
    rmin, _, failed, _ = fly.LBFGS().minimise(s1, pos, pot)
    assert not failed[0], "set-up minimisation did not converge"
    rmin = rmin[0]
    s1.close()
    gids = shard.replica_ids(world * Q, world, rank)
    P, A = start_states(rmin, ex, N - 1, gids, args.seed)

    sysR = fly.System(data, box, ids, replicas=R)
    nbytes = P.nbytes
    d_pos, d_axis, d_outp, d_outa = (ctx.alloc(nbytes) for _ in range(4))
    d_min = ctx.alloc(rmin.nbytes)
    d_code, d_cnt = ctx.alloc(4 * Q), ctx.alloc(8 * 4 * Q)
    ctx.upload(d_pos, P), ctx.upload(d_axis, A), ctx.upload(d_min, rmin)
    dopt, ropt = fly.dimer_options(), fly.rotor_options()
    theta_tol = np.pi / 6  # empty SP cache, first batch of an environment (find.cpp:692 at tot = 0)

    def step_dev():
        capi.check(lib.ofly_dimer_find_sp_stream_dev(sysR.h, C.byref(dopt), C.byref(ropt), Q, d_pos, d_axis, d_min, None, 0, theta_tol, 0,
                                                     d_outp, d_outa, d_code, d_cnt))

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(*vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0, ticks0 = ctx.launch_count(), ctx.tick_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step_dev()
        ev1.record(stream)
        barrier()
    launches, ticks = ctx.launch_count() - launches0, ctx.tick_count() - ticks0
    ms_total = ev0.elapsed_time(ev1)
    my_ms = ms_total

    cnt = np.zeros((Q, 4), dtype=np.int64)
    code = np.zeros(Q, dtype=np.int32)
    ctx.download(cnt, d_cnt), ctx.download(code, d_code)
    evals_step = int(cnt[:, 0].sum())  # gradient evaluations of one step on this rank
    nbar = fly.List(sysR, 0).mean_count()

    # end-to-end leg: host buffers in, host buffers out, through the reference-facing call
    hP, hA = torch.from_numpy(P).pin_memory(), torch.from_numpy(A).pin_memory()
    oP, oA = torch.empty_like(hP).pin_memory(), torch.empty_like(hA).pin_memory()
    hcode, hcnt = np.zeros(Q, dtype=np.int32), np.zeros((Q, 4), dtype=np.int64)
    _dp, _ip, _lp = capi._dp, capi._ip, capi._lp

    def step_host():
        capi.check(lib.ofly_dimer_find_sp_stream(sysR.h, C.byref(dopt), C.byref(ropt), Q, C.cast(hP.data_ptr(), _dp), C.cast(hA.data_ptr(), _dp),
                                                 capi.ptr(rmin, _dp), None, 0, theta_tol, 0, C.cast(oP.data_ptr(), _dp),
                                                 C.cast(oA.data_ptr(), _dp), capi.ptr(hcode, _ip), capi.ptr(hcnt, _lp)))

    step_host()
    barrier()
    e2e_steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(hcode, code) and np.array_equal(hcnt, cnt), "host-buffer and device-resident calls disagree"

    # records of the global queue, merged on the host in id order (Master's merge after taskwait, find.cpp:737-752)
    merged = shard.merge_records(gids, {"code": hcode, "evals": hcnt[:, 0]}, world * Q)

    # the dominant kernels alone: gradient evaluations of the resident batch (every rank; rank 0 reports)
    r_list = data.rcut() + 2 * ropt.delta_r
    r_list += max(dopt.skin_frac ** (1 / 3) - 1, 0) * r_list
    d_slot = ctx.alloc(P[:R].nbytes)
    ctx.upload(d_slot, P[:R])
    capi.check(lib.ofly_neigh_rebuild_dev(sysR.h, r_list, d_slot))

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        ctx.synchronize()
        return a.elapsed_time(b) / n

    iters = 20
    gms = timed(lambda: capi.check(lib.ofly_eam_gradient_dev(sysR.h, d_outp)), iters)
    n2 = fly.List(sysR, 0).mean_count()
    grad_leg = {"atom_evals_per_s": R * N / (gms * 1e-3), "ms_per_launch_pair": gms, "mean_neighbours": n2, "launch_pairs_timed": iters}

    # neigh::List::rebuild of the same batch (six kernels per call)
    rb_ms = timed(lambda: capi.check(lib.ofly_neigh_rebuild_dev(sysR.h, r_list, d_slot)), 10, warm=2)
    rebuild_leg = {"atoms_per_s": R * N / (rb_ms * 1e-3), "ms_per_rebuild": rb_ms, "algorithmic_bytes_per_atom": 24 + 4 * n2 + 16,
                   "hbm_frac": R * N / (rb_ms * 1e-3) * (24 + 4 * n2 + 16) / 1e9 / peaks()[0]}
    ctx.free(d_slot)
    sysR.close()
    for d in (d_pos, d_axis, d_outp, d_outa, d_min, d_code, d_cnt):
        ctx.free(d)

    # strong-scaling leg, every rank: BASELINE configs[4]
    c5_leg = None
    if not args.no_c5:
        c5_leg = c5_strong(ctx, data, pot, dopt, ropt, args.seed, rank, world, barrier, reduce_max, torch, stream)

    single = {}
    if rank == 0:
        if not args.no_c4:
            single["c4_single_system"] = guarded_leg(lambda: c4_single(ctx, data, r_list, timed))
        if not args.no_c3:
            single["c3_environment"] = guarded_leg(lambda: c3_environment(ctx, data, pot, dopt, ropt, args.seed))
        if not args.no_c1:
            single["c1_small_system"] = guarded_leg(lambda: c1_small(ctx, data, pot, timed, with_cpu=(world == 1 and not args.no_cpu)))
        single["fp64_fma"] = guarded_leg(lambda: fp64_fma(ctx))

    # reduce over ranks: max time, sum of work
    ms_total, e2e_s = reduce_max(ms_total, e2e_s)
    if world > 1:
        w = torch.tensor([evals_step, Q, int((code == 0).sum())], dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        evals_all, searches_all, success_all = float(w[0]), float(w[1]), float(w[2])
        tl = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(tl, torch.tensor([my_ms / args.steps, float(cnt[:, 0].max())], dtype=torch.float64, device="cuda"))
        per_rank = [{"ms_per_step": float(t[0]), "longest_search_evals": int(t[1])} for t in tl]
    else:
        evals_all, searches_all, success_all = float(evals_step), float(Q), float((code == 0).sum())
        per_rank = [{"ms_per_step": my_ms / args.steps, "longest_search_evals": int(cnt[:, 0].max())}]

    if rank == 0:
        sec = ms_total * 1e-3
        value = evals_all * N * args.steps / sec
        e2e = evals_all * N * e2e_steps / e2e_s
        peak, peak_src = peaks()
        b_alg = 8 * grad_leg["mean_neighbours"] + 64
        achieved = grad_leg["atom_evals_per_s"] * b_alg / 1e9
        step_share = (evals_step * N / grad_leg["atom_evals_per_s"]) / (my_ms * 1e-3 / args.steps)
        traffic, traffic_src = ncu_traffic({"replicas": R, "atoms": N, "group": 4, "pair_ctas_per_sm": int(os.environ.get("OFLY_PAIR_CTAS", 32))})
        gcode = merged["code"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.seed),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(2 * nbytes + rmin.nbytes),
                    "d2h_bytes_per_step": int(2 * nbytes + 4 * Q + 32 * Q), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "ticks_per_step": ticks / args.steps,
            "mean_active_replicas": evals_step / (ticks / args.steps),
            "dimer_searches_per_s": searches_all * args.steps / sec,
            "successful_saddles_per_s": success_all * args.steps / sec,
            "gradient_evals_per_search": evals_all / searches_all,
            "exit_codes": {str(k): int((gcode == k).sum()) for k in sorted(set(gcode.tolist()))},
            "global_queue": {"searches": int(world * Q), "merged_records": int(len(gcode)), "longest_search_evals": int(merged["evals"].max())},
            "per_rank": per_rank,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "kernel": "k_eam_rho + k_eam_force: one EAM gradient evaluation of the resident batch (64 x 16000 atoms per launch pair)",
                         "launch_ms": grad_leg["ms_per_launch_pair"], "launches_timed": grad_leg["launch_pairs_timed"],
                         "peak_source": peak_src, "algorithmic_bytes_per_atom_eval": b_alg, "mean_neighbours": grad_leg["mean_neighbours"],
                         "algorithmic_bytes_per_launch_pair": b_alg * R * N,
                         "share_of_step_at_full_occupancy": step_share,
                         "fp64_co_limit": single.get("fp64_fma"),
                         "note": "co-limited by the FP64 pipe and the L1 data pipe (one wavefront per gathered 32 B sector), see DESIGN.md 4 and "
                                 "profiles/; the search step as a whole moves %.1f GB/s of algorithmic gradient bytes"
                                 % (evals_step * N * b_alg / (my_ms * 1e-3 / args.steps) / 1e9)},
            "clocks": clocks.summary(),
            "neighbour_rebuild": rebuild_leg,
        }
        if c5_leg:
            out["c5_strong"] = c5_leg
        out.update({k: v for k, v in single.items() if k != "fp64_fma"})
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(rmin, ex, ids, P, A, steps=1)
        print(json.dumps(out), flush=True)

    data.close(), ctx.close()
    if world > 1:
        dist.destroy_process_group()


def guarded_leg(fn):
    try:  # a secondary leg must never cost the headline line
        return fn()
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def c5_strong(ctx, data, pot, dopt, ropt, seed, rank, world, barrier, reduce_max, torch, stream):
    """BASELINE configs[4]: ONE global set of 2048 dimer searches on the apps/hydrogen-sized cell (6^3 Fe - 1 vacancy + 1 H =
    431 atoms), sharded by id over the GPUs (strong scaling); every rank streams its shard through min(shard, 1024) slots.
    Time = device time of the slowest rank; records merged on the host."""
    import ctypes as C

    from openfly_b200 import capi, fly, inputs, shard

    pos, ex, ids = inputs.iron_hydrogen_cell(6, 2.855, remove=(1,), hydrogen=True)
    box = fly.Box(tuple(ex))
    s1 = fly.System(data, box, ids)
    rmin, _, failed, _ = fly.LBFGS().minimise(s1, pos, pot)
    s1.close()
    rmin = rmin[0]
    gids = shard.replica_ids(C5_TOTAL, world, rank)
    P, A = start_states(rmin, ex, len(ids) - 1, gids, seed + 7)
    Qr = len(gids)
    slots = min(Qr, int(os.environ.get("OFLY_C5_SLOTS", C5_SLOTS)))
    s5 = fly.System(data, box, ids, replicas=slots)
    d = [ctx.alloc(P.nbytes) for _ in range(4)]
    d_min, d_code, d_cnt = ctx.alloc(rmin.nbytes), ctx.alloc(4 * Qr), ctx.alloc(32 * Qr)
    ctx.upload(d[0], P), ctx.upload(d[1], A), ctx.upload(d_min, rmin)
    lib = ctx.lib

    def go():
        capi.check(lib.ofly_dimer_find_sp_stream_dev(s5.h, C.byref(dopt), C.byref(ropt), Qr, d[0], d[1], d_min, None, 0, np.pi / 6, 0,
                                                     d[2], d[3], d_code, d_cnt))

    go()
    barrier()
    reps = 2
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        go()
    b.record(stream)
    barrier()
    mine = a.elapsed_time(b) * 1e-3 / reps
    (t5,) = reduce_max(mine)
    code, cnt = np.zeros(Qr, np.int32), np.zeros((Qr, 4), np.int64)
    ctx.download(code, d_code), ctx.download(cnt, d_cnt)
    merged = shard.merge_records(gids, {"code": code, "evals": cnt[:, 0]}, C5_TOTAL)
    for p in d + [d_min, d_code, d_cnt]:
        ctx.free(p)
    s5.close()
    ev = merged["evals"]
    return {"workload": "C5: 2048 dimer searches on the 431-atom Fe-H cell, ONE global set sharded by id over the GPUs (strong scaling)",
            "searches": C5_TOTAL, "slots_per_gpu": slots, "search_s": t5, "searches_per_s": C5_TOTAL / t5,
            "atom_evals_per_s": float(ev.sum()) * len(ids) / t5, "this_rank_s": mine,
            "longest_search_evals": int(ev.max()), "mean_search_evals": float(ev.mean()), "saddles": int((merged["code"] == 0).sum()),
            "exit_codes": {str(k): int((merged["code"] == k).sum()) for k in sorted(set(merged["code"].tolist()))},
            "critical_path_note": "a search is a serial chain of gradient evaluations: search_s >= longest_search_evals x tick latency on "
                                  "whatever rank owns that search, however many GPUs share the rest"}


def c4_single(ctx, data, r_list, timed):
    from openfly_b200 import capi, fly, inputs

    lib = ctx.lib
    p4, ex4, id4 = inputs.vacancy_cell(50, LATTICE, 0.01, 0.05, seed=4331)
    s4 = fly.System(data, fly.Box(tuple(ex4)), id4)
    d4, g4 = ctx.alloc(p4.nbytes), ctx.alloc(p4.nbytes)
    ctx.upload(d4, p4)
    rb4 = timed(lambda: capi.check(lib.ofly_neigh_rebuild_dev(s4.h, r_list, d4)), 10, warm=2)
    gr4 = timed(lambda: capi.check(lib.ofly_eam_gradient_dev(s4.h, g4)), 20)
    nb4 = fly.List(s4, 0).mean_count()
    out = {"workload": "C4: bcc Fe 50^3, 1 % vacancies, jitter N(0, 0.05 A): one 247 500-atom system, dimer list radius",
           "atoms": len(id4), "mean_neighbours": nb4,
           "gradient_atom_evals_per_s": len(id4) / (gr4 * 1e-3), "gradient_ms": gr4,
           "gradient_hbm_frac": len(id4) / (gr4 * 1e-3) * (8 * nb4 + 64) / 1e9 / peaks()[0],
           "rebuild_atoms_per_s": len(id4) / (rb4 * 1e-3), "rebuild_ms": rb4}
    ctx.free(d4), ctx.free(g4)
    s4.close()
    return out


def c1_small(ctx, data, pot, timed, with_cpu):
    """BASELINE configs[0]: bcc Fe 10^3 - 1 vacancy (1 999 atoms), jitter N(0, 0.05 A): List::rebuild(5.3) + energy + gradient,
    GPU (device-resident positions; one small system is a latency measurement, not a throughput one) beside the CPU oracle."""
    from openfly_b200 import capi, fly, inputs

    lib = ctx.lib
    pos, ex, ids = inputs.iron_hydrogen_cell(10, LATTICE, remove=(1,), hydrogen=False)
    pos = pos + np.random.default_rng(1234).normal(0, 0.05, pos.shape)
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    d, g = ctx.alloc(pos.nbytes), ctx.alloc(pos.nbytes)
    ctx.upload(d, pos)
    e = np.zeros(1)
    rb = timed(lambda: capi.check(lib.ofly_neigh_rebuild_dev(s.h, data.rcut(), d)), 20)
    gr = timed(lambda: capi.check(lib.ofly_eam_gradient_dev(s.h, g)), 50)
    en = timed(lambda: capi.check(lib.ofly_eam_energy(s.h, capi.ptr(e, capi._dp))), 20)
    out = {"workload": "C1: bcc Fe 10^3 - 1 vacancy = 1999 atoms, jitter N(0, 0.05 A): rebuild(5.3) + energy + gradient of ONE system",
           "atoms": len(ids), "gpu": {"rebuild_ms": rb, "gradient_ms": gr, "energy_ms_incl_readback": en,
                                      "gradient_atom_evals_per_s": len(ids) / (gr * 1e-3)}}
    ctx.free(d), ctx.free(g)
    s.close()
    if with_cpu:
        n_threads = host_threads()
        O, path, flags = oracle_fast_lib()
        tab = O.Tables(WEN, fast=True, libpath=path)
        osys = O.System(tab, ex)
        fz = np.zeros(len(ids), np.uint8)
        cores = max(tab.L.max_threads(), n_threads)
        osys.rebuild(tab.rcut, pos, cores)
        t_rb = osys.time_rebuild(tab.rcut, pos, cores, 20) / 20
        t_gr = osys.time_gradient(ids, fz, np.zeros_like(pos), cores, 50) / 50
        t0 = time.perf_counter()
        for _ in range(20):
            osys.energy(ids, cores)
        t_en = (time.perf_counter() - t0) / 20
        out["cpu"] = {"threads": int(cores), "rebuild_ms": t_rb * 1e3, "update_plus_gradient_ms": t_gr * 1e3, "energy_ms": t_en * 1e3,
                      "gradient_atom_evals_per_s": len(ids) / t_gr, "oracle_flags": flags}
    return out


def fp64_fma(ctx):
    """The co-limiter of the pair kernels: achievable FP64 FMA rate of this device, measured with a register-only
    kernel of independent FMA chains (ofly_bench_fp64_fma)."""
    import ctypes as C

    from openfly_b200 import capi

    rate = C.c_double()
    capi.check(ctx.lib.ofly_bench_fp64_fma(ctx.h, 1 << 14, C.byref(rate)))
    # one atom-gradient evaluation: n = 58 pairs x (~22 FP64 instructions in pass 1 + ~40 in pass 2), counted from the SASS
    per_atom = 58 * 62
    return {"fma_per_s": rate.value, "tflops": 2 * rate.value / 1e12,
            "pair_fp64_instructions_per_atom_eval": per_atom,
            "atom_evals_per_s_ceiling_if_fp64_bound": rate.value / per_atom}


def c3_environment(ctx, data, pot, dopt, ropt, seed):
    """BASELINE configs[2]: apps/hydrogen-sized cell (6^3 Fe - 1 vacancy + 1 H = 431 atoms), 256 dimer searches from one
    environment, then for every saddle found the mass-weighted Hessian and its eigenvalues on the device and the Vineyard
    prefactor (find.cpp:846-871); wall-clock, host buffers in and out."""
    from openfly_b200 import fly, inputs

    pos, ex, ids = inputs.iron_hydrogen_cell(6, 2.855, remove=(1,), hydrogen=True)
    box = fly.Box(tuple(ex))
    s1 = fly.System(data, box, ids)
    rmin, _, failed, _ = fly.LBFGS().minimise(s1, pos, pot)
    R = 256
    P, A = inputs.perturbed_batch(rmin[0], ex, len(ids) - 1, R, seed)
    sR = fly.System(data, box, ids, replicas=R)
    dimer = fly.Dimer(dopt, ropt)
    dimer.find_sp(sR, P, A, rmin[0], pot)  # warm-up (allocations, module load)
    t0 = time.perf_counter()
    out_r, out_ax, code, cnt = dimer.find_sp(sR, P, A, rmin[0], pot)
    t_search = time.perf_counter() - t0
    sR.close()
    s1.close()
    ok = np.nonzero(code == 0)[0]
    n_ok = len(ok)
    try:
        mech = fly.Mechanisms(data, box, ids, replicas=int(min(max(n_ok, 1), 192)))  # every saddle's Hessian in flight at once
        mech.prefactors(rmin[0], out_r[ok[:2]])  # warm-up
        with ClockSampler(ctx.device) as mech_clocks:
            t0 = time.perf_counter()
            res = mech.prefactors(rmin[0], out_r[ok])
            t_hess = time.perf_counter() - t0
            time.sleep(0.3)
        clk = mech_clocks.summary()
        nus = res["prefactor"][res["valid"] == 1]
        mech.close()
    except fly.OflyError as e:
        t_hess, nus, clk = None, str(e), None
    return {"workload": "C3: 6^3 Fe - vacancy + H (431 atoms), 256 dimer searches of one environment + Hessian / eigenvalues / prefactor per saddle",
            "searches": R, "saddles": n_ok, "search_s": t_search, "searches_per_s": R / t_search,
            "atom_evals_per_s": float(cnt[:, 0].sum()) * len(ids) / t_search, "longest_search_evals": int(cnt[:, 0].max()),
            "hessian_eigen_s": t_hess, "hessian_eigen_ms_per_saddle": (t_hess / max(n_ok, 1) * 1e3) if t_hess else None,
            "hessian_eigen_clocks": (clk if t_hess else None),
            "prefactors": (int(len(nus)) if not isinstance(nus, str) else nus),
            "median_prefactor_hz": (float(np.median(nus)) if not isinstance(nus, str) and len(nus) else None)}


# --------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which
    would otherwise turn the all-core baseline into a single-core one)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def oracle_fast_lib():
    """The oracle built with the reference's flags (-O3 -march=native) for THIS host when g++ is present,
    else the prebuilt x86-64-v3 file that travelled with the repository."""
    from oracle import oracle as O

    src = os.path.join(ROOT, "oracle", "capi.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    out = os.path.join(tempfile.gettempdir(), f"liboracle_native_{os.getuid()}.so")
    try:
        subprocess.run([cxx, "-std=c++17", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", out, src], check=True,
                       capture_output=True, timeout=300)
        return O, out, "-O3 -march=native -fopenmp (reference CMakePresets flags), built on this host"
    except Exception:  # noqa: BLE001
        return O, None, "-O3 -march=x86-64-v3 -fopenmp (prebuilt)"


def cpu_m1_sample(osys, cores, P, A, ids, rmin, O):
    """M1 on the CPU: `cores` of the queue's searches, one single-threaded search per core (saddle::Master), each cut at
    CPU_SAMPLE_MAX_STEPS translation steps so that a step of the sample stays a few seconds."""
    fz = np.zeros(len(ids), np.uint8)
    iopt = O.DIMER_IOPT.copy()
    iopt[1] = CPU_SAMPLE_MAX_STEPS
    t0 = time.perf_counter()
    _, _, code, cnt = osys.find_sp_batch(P[:cores], A[:cores], ids, fz, rmin, threads=cores, iopt=iopt)
    dt = time.perf_counter() - t0
    return float(cnt[:, 0].sum()) * len(ids) / dt, dt


def cpu_m2_sample(O, tab, cores, seed):
    """M2 on the CPU: FULL dimer searches per second. The first 2 x cores searches of the C5 set (431-atom cell, the same
    start states the GPU's c5_strong leg runs), one single-threaded search per core, run to their own exit."""
    from openfly_b200 import inputs

    pos, ex, ids = inputs.iron_hydrogen_cell(6, 2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = O.System(tab, ex)
    rmin, _, failed, _ = osys.minimise(pos, ids, fz, threads=cores)
    n = 2 * cores
    P, A = start_states(rmin, ex, len(ids) - 1, np.arange(n), seed + 7)
    t0 = time.perf_counter()
    _, _, code, cnt = osys.find_sp_batch(P, A, ids, fz, rmin, threads=cores)
    dt = time.perf_counter() - t0
    return {"searches_per_s": n / dt, "searches": n, "seconds": dt, "atoms": len(ids), "mean_search_evals": float(cnt[:, 0].mean()),
            "longest_search_evals": int(cnt[:, 0].max()), "atom_evals_per_s": float(cnt[:, 0].sum()) * len(ids) / dt,
            "exit_codes": {str(k): int((code == k).sum()) for k in sorted(set(code.tolist()))},
            "what": "full searches (no step cut) of the C5 set's first 2 x cores start states, one single-threaded search per core"}


def cpu_baseline(rmin, ex, ids, P, A, steps: int):
    n_threads = host_threads()
    O, path, flags = oracle_fast_lib()
    tab = O.Tables(WEN, fast=True, libpath=path)
    osys = O.System(tab, ex)
    cores = max(tab.L.max_threads(), n_threads)
    fz = np.zeros(len(ids), np.uint8)
    best = None
    for _ in range(steps):
        v, _ = cpu_m1_sample(osys, cores, P, A, ids, rmin, O)
        best = v if best is None else max(best, v)
    # the same kernels as the roofline pair, CPU side: update + gradient of ONE C2 system with every core inside the call,
    # exactly as EAM::gradient(..., threads) is used (eam.cpp:81,115); list at the dimer's radius
    r_list = 5.3 + 0.002
    r_list += (1.1 ** (1 / 3) - 1) * r_list
    osys.rebuild(r_list, P[0], cores)
    delta = np.random.default_rng(0).normal(0, 1e-3, P[0].shape)
    osys.time_gradient(ids, fz, delta, cores, 2)
    iters = 20
    g_s = osys.time_gradient(ids, fz, delta, cores, iters)
    m2 = guarded_leg(lambda: cpu_m2_sample(O, tab, cores, 1334))
    return {"value": best, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": f"{cores} of this GPU's queued C2 searches (one single-threaded search per core, as saddle::Master), each cut at "
                      f"max_steps={CPU_SAMPLE_MAX_STEPS}; oracle flags: {flags}",
            "gradient_only": {"atom_evals_per_s": len(ids) * iters / g_s, "ms_per_eval": g_s / iters * 1e3, "threads": int(cores),
                              "what": "List::update + EAM::gradient of one 16 000-atom system, all cores inside the call"},
            "m2_full_searches": m2}


def run_reference(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    n_threads = host_threads()
    O, path, flags = oracle_fast_lib()
    tab = O.Tables(WEN, fast=True, libpath=path)
    pos, ex, ids = c2_cell()
    N = len(ids)
    osys = O.System(tab, ex)
    cores = max(tab.L.max_threads(), n_threads)
    fz = np.zeros(N, np.uint8)
    rmin, _, failed, _ = osys.minimise(pos, ids, fz, threads=cores)
    assert not failed
    # the first `cores` searches of rank 0's queue: same seeds, same recipe as the GPU arm (each arm relaxes the cell
    # itself; the two minima agree to ~1e-6 A)
    from openfly_b200 import shard

    gids = shard.replica_ids(world * QUEUE, world, 0)[:cores]
    P, A = start_states(rmin, ex, N - 1, gids, args.seed)

    def step():
        v, dt = cpu_m1_sample(osys, cores, P, A, ids, rmin, O)
        return v * dt / N  # gradient evaluations of the step

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        evals += step()
    dt = time.perf_counter() - t0
    value = evals * N / dt
    sample = (f"{cores} of the queue's C2 searches per step (one single-threaded search per core, as saddle::Master), each cut at "
              f"max_steps={CPU_SAMPLE_MAX_STEPS}; oracle flags: {flags}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args.seed),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself cannot be built here (Eigen/fmt/cereal/nauty/KIM-API absent, DESIGN.md); this is the CPU oracle port",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=1334)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 small-system leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 single-system leg")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 environment leg (256 searches + Hessians / prefactors)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 strong-scaling leg (2048 searches sharded over the GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
