#!/usr/bin/env python
"""Benchmark of the FLIP time-step hot path (BASELINE.json metric: particle-steps/s and ms/frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one frame, FluidSimulation::update(1/30) (fluidsimulation.cpp:5755), i.e. all CFL
substeps of the whole hot path (liquid SDF, P2G, extrapolation, body force, pressure projection,
G2P, RK3 advection + collision + removal).  Workload (pick_workload): on a 1-GPU box the synthetic
sphere-drop scene of SURVEY §8d config 3 (256^3 grid, 16.1 M marker particles); for the scaling
series -- N > 1 under torchrun, and N = 1 on a box that shows more than one GPU -- the dam break of
config 5 (512^3, 127.9 M particles), the same fixed work at every N (strong scaling).  value = sum
over timed substeps of live particles / device time (CUDA events on the context's stream, max over
ranks).

Our arm keeps the state resident in HBM for `value`; the `e2e` leg repeats the same frames through
the C-ABI with HOST buffers: flip_set_particles (pinned host AoS -> device), flip_update,
flip_get_particles (device -> pinned host AoS), every step, inside the timed region.  With N > 1 the
run starts with a slab-vs-single-GPU parity check of the same scene (`parity_ok` in the line).

--impl reference times the UNMODIFIED reference engine (oracle/_ref/libflipref_fast.so, built by
oracle/Makefile from /root/reference) on the host cores, all threads, on the SAME scene at full
size; a 256^3 frame costs tens of CPU seconds, so the frame count is bounded by a wall-clock budget
(the metric is per particle-substep).  This file and tests/ are the only places that may execute
anything under oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAME_DT = 1.0 / 30.0
METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------------------------------
def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class quiet_stdout:
    """The reference engine prints a version banner and its log straight to fd 1; keep our stdout to
    the single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def workload(grid, name="spheredrop", krange=None):
    from flipengine3d_b200 import scenes
    if name == "dambreak":
        return scenes.dam_break(grid, krange=krange)
    return scenes.sphere_drop(grid)


def visible_gpus():
    try:
        import torch
        return int(torch.cuda.device_count())
    except Exception:
        return 0


def pick_workload(args, world):
    """The workload of this run.  BENCH (a 1-GPU box): the single-GPU headline, spheredrop256.  The scaling series
    (N = 1, 2, 4, 8 on ONE multi-GPU box) must be the same workload at every N, dambreak512 (BASELINE config 5),
    so on a box that shows more than one GPU the N=1 run is that series' base point."""
    if args.workload is not None:
        return args.workload, args.grid, "given on the command line"
    if world > 1:
        return "dambreak", 512, "z-slab strong-scaling series"
    if visible_gpus() > 1:
        return "dambreak", 512, "N=1 point of the strong-scaling series (this box shows more than one GPU)"
    return "spheredrop", 256, "single-GPU headline"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference engine on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(name, grid, steps, warmup, budget_s, threads=None):
    """The reference engine's own update(1/30) on the SAME scene the GPU arm runs, all host threads.  A frame of the
    256^3 scene costs tens of seconds of CPU, so the number of frames is bounded by `budget_s` of wall clock (at
    least one warm-up and one timed frame); the metric is normalised per particle-substep, so it does not depend on
    the frame count.  Returns (particle_steps_per_s, ms_per_frame, info)."""
    from oracle import refengine
    kind = "fast" if refengine.available("fast") else "golden"
    if not refengine.available(kind):
        raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    sc = workload(grid, name)
    with quiet_stdout():
        ref = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], kind=kind, threads=threads)
        cores = ref.L.ref_get_threads()
        w_done, t_frame = 0, None
        t_w0 = time.perf_counter()
        while w_done < max(warmup, 0):
            t0 = time.perf_counter()
            ref.update(FRAME_DT)
            t_frame = time.perf_counter() - t0
            w_done += 1
            # warm-up may use up to 30 % of the budget
            if time.perf_counter() - t_w0 + t_frame > 0.3 * budget_s:
                break
        psteps, frames, t0 = 0, 0, time.perf_counter()
        while frames < max(steps, 1):
            n_before = ref.num_particles
            t1 = time.perf_counter()
            ref.update(FRAME_DT)
            t_frame = time.perf_counter() - t1
            psteps += n_before * max(ref.substeps, 1)
            frames += 1
            if time.perf_counter() - t0 + t_frame > 0.7 * budget_s:
                break
        el = time.perf_counter() - t0
    info = dict(kind="reference", cores=int(cores), build=kind, frames_timed=frames, warmup_frames=w_done,
                substeps_last_frame=int(ref.substeps),
                sample=f"{name}{grid} ({sc['pos'].shape[0]} particles: the GPU arm's scene), {frames} frame(s) of update(1/30) "
                       f"after {w_done} warm-up frame(s) (bounded by a {budget_s:.0f} s wall-clock budget), surface reconstruction off")
    ref.close()
    return psteps / el, 1e3 * el / frames, info


def reference_arm(args, rank, world):
    if rank != 0:
        return
    wl, grid, why = pick_workload(args, world)
    v, ms, info = run_reference(wl, grid, args.steps, args.warmup, args.ref_budget)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 fields / f64 PCG", "data": "synthetic",
            "config": {"workload": f"{wl}{grid}", "workload_choice": why, "same_config": True, "dx": 0.125, "frame_dt": FRAME_DT,
                       "step": "one frame = FluidSimulation::update(1/30)", "frames_timed": info["frames_timed"],
                       "warmup_frames": info["warmup_frames"]},
            "cpu_baseline": dict(info, value=v, unit=UNIT),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def liquid_face_count(phi):
    """Nf_liq of SURVEY §8d: faces adjacent to at least one liquid cell."""
    liq = phi < 0
    K, J, I = liq.shape
    u = np.zeros((K, J, I + 1), dtype=bool); u[:, :, :-1] |= liq; u[:, :, 1:] |= liq
    v = np.zeros((K, J + 1, I), dtype=bool); v[:, :-1, :] |= liq; v[:, 1:, :] |= liq
    w = np.zeros((K + 1, J, I), dtype=bool); w[:-1, :, :] |= liq; w[1:, :, :] |= liq
    return int(u.sum() + v.sum() + w.sum())


def algorithmic_bytes(Np, dims, n_rows, nf_liq):
    """SURVEY §8d 'Algorithmic bytes' per launch of each kernel class."""
    I, J, K = dims
    Nc = I * J * K
    Nf = (I + 1) * J * K + I * (J + 1) * K + I * J * (K + 1)
    P = 16 * n_rows                               # preconditioner data of level 0: 1/diag + three face weights, fp32
    return {
        "sdf_p2g": 24 * Np + 5 * Nf + 4 * Nc,     # fused liquid SDF + P2G: particles read once
        "g2p": 36 * Np + 8 * nf_liq,
        "advance": 24 * Np + 4 * nf_liq,
        "g2p_advance": 48 * Np + 8 * nf_liq,      # fused G2P + RK3
        "extrapolate": 9 * Nf,                     # one extrapolateVelocityField call (3 components)
        "pcg_spmv": 36 * n_rows,
        "pcg_dir_spmv": 60 * n_rows,               # direction update (z, s in, s out: 24n) fused with the SpMV (36n)
        "precond": 16 * n_rows + P,                # V-cycle: r in, z out (fp64) + P; the coarse levels are not credited
        "pcg_iter": 124 * n_rows + P,              # SpMV+dot 36n; x,r update 48n; precond 16n+P; direction 24n
    }


def make_sim(args, wl_name, grid, rank, world, local_rank, dist, ids=False):
    """One context per rank; with world > 1 the domain is split into z-slabs (flip_set_slab)."""
    from flipengine3d_b200 import engine as fe
    I = J = K = grid
    sim = fe.FluidSimulation(I, J, K, 0.125, device=local_rank)
    sim.addBodyForce(0.0, -25.0, 0.0)
    if args.preconditioner:
        sim.setPreconditioner(args.preconditioner)
    krange = None
    if world > 1:
        ident = [fe.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sim.setSlab(rank, world, ident[0])
        krange = fe.slab_range(K, world, rank)
    sc = workload(grid, wl_name, krange=krange)
    if ids:     # global ids: a rank that generates only its own planes counts from the particles that precede them
        sim.enableParticleIds(True, base=sc.get("id_offset", 0))
    sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
    sim.initialize()
    return sim, sc


def timed_frames(sim, steps, stream, torch, barrier):
    """K frames with the state resident in HBM; returns (ms, particle-steps, substeps, pcg its, rows, stage_ms)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    psteps, substeps, pcg_iters, rows = 0, 0, 0, []
    stage_ms = {}
    for _ in range(steps):
        sim.update(FRAME_DT)
        for st in sim.substep_stats():
            # particles that went through the substep = survivors + those removed at its end (global counts)
            psteps += st["particles"] + st["removed_solid"] + st["removed_crowded"] + st["removed_fast"]
            substeps += 1
            pcg_iters += st["pcg_iterations"]
            rows.append(st["pressure_rows"])
        for k, v in sim.stage_times_ms().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1), psteps, substeps, pcg_iters, rows, stage_ms


def slab_parity_check(args, wl_name, grid, rank, world, local_rank, dist, torch, frames=2):
    """Before anything is timed on N > 1 GPUs: the z-slab run against the single-GPU run of the SAME scene (rank 0
    runs the whole domain on its own GPU next to its slab).  Integer bookkeeping (substeps, particle count, pressure
    rows, the set of particle ids) must be equal, per-particle positions rel-L2 <= 1e-4 (tests/parity_common.py).
    Returns the report dict on rank 0, None elsewhere."""
    dev = torch.device("cuda", local_rank)
    sim, _ = make_sim(args, wl_name, grid, rank, world, local_rank, dist, ids=True)
    single = None
    if rank == 0:
        single, _ = make_sim(args, wl_name, grid, 0, 1, local_rank, dist, ids=True)
    rep = {"workload": f"{wl_name}{grid}", "frames": frames, "ok": True} if rank == 0 else None
    for f in range(frames):
        sim.update(FRAME_DT)
        st = sim.substep_stats()
        if rank == 0:
            single.update(FRAME_DT)
            rs = single.substep_stats()
            same = (len(st) == len(rs) and [s["pressure_rows"] for s in st] == [s["pressure_rows"] for s in rs]
                    and st[-1]["particles"] == rs[-1]["particles"])
            rep["ok"] = bool(rep["ok"] and same)
            rep[f"frame{f}"] = {"substeps": [len(st), len(rs)], "particles": [st[-1]["particles"], rs[-1]["particles"]],
                                "rows": [[s["pressure_rows"] for s in st], [s["pressure_rows"] for s in rs]],
                                "pcg_iterations": [[s["pcg_iterations"] for s in st], [s["pcg_iterations"] for s in rs]]}
    # per-particle comparison after the last frame: every rank's (particles, ids) to rank 0 through NCCL
    p = torch.from_numpy(sim.getMarkerParticles().copy()).to(dev)
    ids = torch.from_numpy(sim.getParticleIds().astype(np.int64)).to(dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([p.shape[0]], dtype=torch.int64, device=dev))
    if rank == 0:
        P, IDS = [p], [ids]
        for r in range(1, world):
            n = int(counts[r].item())
            bp = torch.empty((n, 6), dtype=torch.float32, device=dev)
            bi = torch.empty((n,), dtype=torch.int64, device=dev)
            dist.recv(bp, src=r)
            dist.recv(bi, src=r)
            P.append(bp); IDS.append(bi)
        P, IDS = torch.cat(P), torch.cat(IDS)
        R = torch.from_numpy(single.getMarkerParticles().copy()).to(dev)
        RID = torch.from_numpy(single.getParticleIds().astype(np.int64)).to(dev)
        same_ids = P.shape[0] == R.shape[0]
        if same_ids:
            o, ro = torch.argsort(IDS), torch.argsort(RID)
            same_ids = bool(torch.equal(IDS[o], RID[ro]))
        rep["same_ids"] = bool(same_ids)
        rep["local_counts"] = [int(c.item()) for c in counts]
        if same_ids:
            a, b = P[o, :3].double(), R[ro, :3].double()
            rep["pos_rel_l2"] = float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
            rep["pos_max_abs"] = float((a - b).abs().max())
            rep["ok"] = bool(rep["ok"] and rep["pos_rel_l2"] <= 1e-4 and rep["pos_max_abs"] <= 1e-3 * 0.125)
        else:
            rep["ok"] = False
        single.close()
    else:
        dist.send(p, dst=0)
        dist.send(ids, dst=0)
    sim.close()
    del sim, p, ids
    torch.cuda.empty_cache()
    dist.barrier()
    return rep


def ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libflip_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl_name, grid, why = pick_workload(args, world)

    def maxreduce(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumreduce(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    parity = None
    if world > 1 and args.parity_check:
        parity = slab_parity_check(args, wl_name, grid, rank, world, local_rank, dist, torch, frames=args.parity_frames)

    sim, sc = make_sim(args, wl_name, grid, rank, world, local_rank, dist)
    I, J, K = sc["dims"]
    stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        sim.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        sim.update(FRAME_DT)
    barrier()

    # ---- timed region: K frames, state resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    sim.enable_kernel_timing(True)
    sim.reset_kernel_timing()
    launches0 = sim.kernel_launches()
    ms, psteps, substeps, pcg_iters, rows, stage_ms = timed_frames(sim, args.steps, stream, torch, barrier)
    launches = sim.kernel_launches() - launches0
    kt = sim.kernel_timing()
    sim.enable_kernel_timing(False)
    clocks = sampler.stop()
    ms_max = maxreduce(ms)
    launches_all = sumreduce(float(launches))
    value = psteps / (ms_max * 1e-3)      # particle counts in the step stats are global

    # ---- e2e leg: the same frames through the C-ABI with host buffers (each rank moves its own slab's particles)
    Np_local = sim.getNumMarkerParticles()
    host = torch.empty((Np_local + Np_local // 4 + 65536, 6), dtype=torch.float32).pin_memory().numpy()
    sim.getMarkerParticles(out=host)
    barrier()
    e2e_psteps, h2d, d2h = 0, 0, 0
    t0 = time.perf_counter()
    e2e_per_step, e2e_substeps = [], 0
    for _ in range(args.steps):
        t_it = time.perf_counter()
        n = sim.getNumMarkerParticles()
        sim.setMarkerParticles(host[:n])             # host -> device (pinned)
        h2d += n * 24
        sim.update(FRAME_DT)
        for st in sim.substep_stats():
            e2e_psteps += st["particles"] + st["removed_solid"] + st["removed_crowded"] + st["removed_fast"]
            e2e_substeps += 1
        n = sim.getNumMarkerParticles()
        if n > host.shape[0]:
            host = torch.empty((n + n // 4, 6), dtype=torch.float32).pin_memory().numpy()
        sim.getMarkerParticles(out=host)             # device -> host (pinned)
        d2h += n * 24
        e2e_per_step.append(round(1e3 * (time.perf_counter() - t_it), 3))
    sim.synchronize()
    e2e_s = maxreduce(time.perf_counter() - t0)
    e2e_value = e2e_psteps / e2e_s
    h2d_all, d2h_all = sumreduce(float(h2d)), sumreduce(float(d2h))
    Np = int(sumreduce(float(Np_local)))

    # ---- the same frames with the reference's literal arithmetic everywhere (FLIP_SAMPLING_EXACT: double-precision
    # trilinear blend in G2P / RK3, literal P2G weights) for the number beside the default single-precision blend
    exact = None
    if args.exact_steps > 0:
        sim.setSamplingMode("exact")
        sim.update(FRAME_DT)
        ms_x, ps_x, sub_x, _, _, _ = timed_frames(sim, args.exact_steps, stream, torch, barrier)
        ms_x = maxreduce(ms_x)
        exact = {"sampling_mode": "exact", "value": ps_x / (ms_x * 1e-3), "unit": UNIT, "ms_per_step": ms_x / args.exact_steps,
                 "steps": args.exact_steps, "substeps_timed": sub_x}
        sim.setSamplingMode("fast")

    line = None
    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        n_rows = int(np.mean(rows)) if rows else 0
        kernels, roof = {}, None
        if world == 1:
            nf_liq = liquid_face_count(sim.array("liquid_phi"))
            ab = algorithmic_bytes(Np, (I, J, K), n_rows, nf_liq)
            for name, (tot_ms, n) in kt.items():
                if n == 0:
                    continue
                avg_ms = tot_ms / n
                ent = {"launches": int(n), "avg_ms": avg_ms, "total_ms": tot_ms, "share_of_step": tot_ms / ms}
                if name in ab:
                    gbs = ab[name] / (avg_ms * 1e-3) / 1e9
                    ent.update(algorithmic_bytes=int(ab[name]), achieved_gbs=gbs, frac=gbs / peak)
                kernels[name] = ent
            # the dominant kernel class: the largest share of the step over ALL classes that have an algorithmic-bytes
            # figure (pcg_iter = one PCG iteration, i.e. SpMV + update + V-cycle + direction; it contains pcg_spmv and
            # precond, which are also listed on their own)
            cands = [k for k in kernels if "algorithmic_bytes" in kernels[k]]
            dom = max(cands, key=lambda k: kernels[k]["total_ms"])
            roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "peak_kind": peak_kind, "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": None,
                    "algorithmic_bytes": kernels[dom]["algorithmic_bytes"], "avg_launch_ms": kernels[dom]["avg_ms"],
                    "share_of_step": kernels[dom]["share_of_step"]}
            # DRAM bytes per launch come from an `ncu --set full` capture of THIS revision when one is committed
            # (profiles/traffic.json names the capture); never from the run itself (ncu-timed runs are not bench values)
            tp_file = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp_file):
                try:
                    tj = json.load(open(tp_file))
                    if tj.get("workload") == f"{wl_name}{grid}" and dom in tj.get("per_launch_bytes", {}):
                        roof["traffic"] = tj["per_launch_bytes"][dom]
                        roof["traffic_source"] = tj.get("source")
                except Exception:
                    pass
        else:
            for name, (tot_ms, n) in kt.items():
                if n:
                    kernels[name] = {"launches": int(n), "avg_ms": tot_ms / n, "total_ms": tot_ms, "share_of_step": tot_ms / ms,
                                     "note": "rank 0's slab"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 fields / f64 PCG vectors", "data": "synthetic",
                "config": {"workload": f"{wl_name}{grid}", "workload_choice": why, "grid": [I, J, K], "dx": 0.125, "particles": Np,
                           "frame_dt": FRAME_DT, "step": "one frame = flip_update(1/30), all CFL substeps",
                           "substeps_timed": substeps, "ms_per_substep": ms_max / max(substeps, 1),
                           "pcg_iterations_timed": pcg_iters, "pressure_rows": n_rows,
                           "sampling_mode": "fast (single-precision 8-point blend in G2P / RK3 and packed P2G weights; indices and "
                                            "weights exact; the reference blends in double) -- see exact_mode for the literal arithmetic",
                           "parallelism": "single GPU" if world == 1 else f"{world} z-slabs (one process per GPU, NVLink halo exchange, distributed PCG)",
                           "l2": "inputs larger than L2 (particles and each MAC field exceed the 126 MB L2)"},
                "roofline": roof, "kernels": kernels, "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all) // args.steps,
                        "d2h_bytes_per_step": int(d2h_all) // args.steps, "ms_per_step": 1e3 * e2e_s / args.steps,
                        "per_step_ms_rank0": e2e_per_step, "substeps_timed": e2e_substeps},
                "exact_mode": exact, "gpu_launches": int(launches_all), "clocks": clocks}
        if parity is not None:
            line["parity_ok"] = bool(parity["ok"])
            line["parity"] = parity
    sim.close()
    del sim

    # ---- N=1 only: the CPU baseline (the reference engine on this box's host cores), on the same scene where a frame
    # of it fits the bounded sample, else on the same scene at half the resolution (said so in `sample`)
    if world == 1 and rank == 0 and args.cpu_budget > 0:
        try:
            same = grid <= 256
            v, cms, info = run_reference(wl_name, grid if same else 256, 1, 1 if grid <= 128 else 0, args.cpu_budget)
            line["cpu_baseline"] = dict(info, value=v, unit=UNIT, ms_per_step=cms, same_config=same)
        except Exception as e:   # the oracle always exists on the GPU box; report loudly if not
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"FAILED: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "spheredrop", "dambreak"],
                    help="default: spheredrop256 on a 1-GPU box; dambreak512 for the scaling series (N > 1, or N = 1 on a multi-GPU box)")
    ap.add_argument("--grid", type=int, default=256, help="grid size when --workload is given")
    ap.add_argument("--ref-budget", type=float, default=240.0,
                    help="wall-clock budget (s) of the reference arm's frames (the scene is the GPU arm's, at full size)")
    ap.add_argument("--cpu-budget", type=float, default=30.0,
                    help="wall-clock budget (s) of the cpu_baseline sample inside our arm (0: skip)")
    ap.add_argument("--exact-steps", type=int, default=3, help="frames timed in FLIP_SAMPLING_EXACT beside the default mode (0: skip)")
    ap.add_argument("--no-parity-check", dest="parity_check", action="store_false",
                    help="N > 1: skip the slab-vs-single-GPU parity check that precedes the timing")
    ap.add_argument("--parity-frames", type=int, default=2)
    ap.add_argument("--preconditioner", default=None)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
