"""BASELINE config 4: pressure-solve-only stress.  256^3 domain, liquid in every cell of [2,n-2)^3 (one particle
per cell centre, uniform(-1,1) velocities), ONE substep run stage by stage; the pressure stage is timed alone.
Prints one JSON line per tolerance (1e-6 of BASELINE.json and the engine default 1e-9): rows, PCG iterations,
solve time, iterations/s, SpMV time and GB/s against the 36 B/row algorithmic traffic.
Usage: pressure_stress.py [n=256] [preconditioner]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import scenes, engine as fe

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else None
sc = scenes.pressure_stress(n)
peak = 6548.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for tol in (1e-6, 1e-9):
    sim = fe.FluidSimulation(n, n, n, sc["dx"])
    sim.addBodyForce(0, -25, 0)
    sim.setPressureSolver(tolerance=tol)
    sim.setPressureWarmStart(False)
    if prec:
        sim.setPreconditioner(prec)
    sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
    sim.initialize()
    best = None
    for rep in range(3):       # the same state every time: the grids are rebuilt from the unchanged particles
        sim.begin_frame(1 / 30)
        dt = sim.begin_substep()
        for st in ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force"):
            sim.stage(st, dt)
        sim.synchronize()
        sim.enable_kernel_timing(True)
        sim.reset_kernel_timing()
        sim.stage("pressure", dt)
        sim.synchronize()
        kt = sim.kernel_timing()
        sim.enable_kernel_timing(False)
        ms = sim.stage_times_ms()["pressure"]
        sim.end_substep()
        sim.end_frame()
        s = sim.substep_stats()[-1]
        rows, its = s["pressure_rows"], s["pcg_iterations"]
        spmv_ms = kt["pcg_spmv"][0] / max(kt["pcg_spmv"][1], 1)
        iter_ms = kt["pcg_iter"][0] / max(kt["pcg_iter"][1], 1)
        line = {"workload": sc["name"], "tolerance": tol, "preconditioner": prec or "multigrid", "rows": rows, "pcg_iterations": its,
                "pcg_error": s["pcg_error"], "rhs_max": s["rhs_max"], "converged": s["pcg_converged"],
                "pressure_stage_ms": ms, "iterations_per_s": its / (ms * 1e-3) if ms > 0 else None,
                "pcg_iteration_ms": iter_ms, "spmv_ms": spmv_ms, "spmv_gbs_36B_per_row": 36.0 * rows / (spmv_ms * 1e-3) / 1e9,
                "spmv_frac_of_hbm_peak": 36.0 * rows / (spmv_ms * 1e-3) / 1e9 / peak,
                "pcg_iteration_gbs_124B_per_row": 124.0 * rows / (iter_ms * 1e-3) / 1e9}
        if best is None or ms < best["pressure_stage_ms"]:
            best = line
    print(json.dumps(best), flush=True)
    sim.close()
