"""BASELINE configs 1-3 side by side: our engine on cuda:0 against the unmodified reference engine (fast build,
all host cores) on the same scene, frames of update(1/30), surface reconstruction off on both.  Writes a markdown
table to stdout.  Usage: config_table.py [frames_default=100] [frames_dam=10] [frames_cpu_dam=2]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flipengine3d_b200 import scenes, engine as fe
from oracle import refengine   # test infrastructure: the CPU arm only

FR_DEF = int(sys.argv[1]) if len(sys.argv) > 1 else 100
FR_DAM = int(sys.argv[2]) if len(sys.argv) > 2 else 10
FR_CPU = int(sys.argv[3]) if len(sys.argv) > 3 else 2


def gpu_run(sc, frames, warm=3):
    I, J, K = sc["dims"]
    sim = fe.FluidSimulation(I, J, K, sc["dx"])
    sim.addBodyForce(0, -25, 0)
    sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
    sim.initialize()
    for _ in range(warm):
        sim.update(1 / 30)
    sim.synchronize()
    t0, ps, sub = time.perf_counter(), 0, 0
    for _ in range(frames):
        sim.update(1 / 30)
        for st in sim.substep_stats():
            ps += st["particles"]
            sub += 1
    sim.synchronize()
    el = time.perf_counter() - t0
    n = sim.getNumMarkerParticles()
    sim.close()
    return 1e3 * el / frames, ps / el, sub / frames, n


def cpu_run(sc, frames, warm=3):
    kind = "fast" if refengine.available("fast") else "golden"
    ref = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], kind=kind)
    cores = ref.L.ref_get_threads()
    for _ in range(warm):
        ref.update(1 / 30)
    t0, ps, sub = time.perf_counter(), 0, 0
    for _ in range(frames):
        n0 = ref.num_particles
        ref.update(1 / 30)
        ps += n0 * max(ref.substeps, 1)
        sub += max(ref.substeps, 1)
    el = time.perf_counter() - t0
    ref.close()
    return 1e3 * el / frames, ps / el, sub / frames, cores


rows = []
for name, sc, fg, fc in (("1 default FluidManager scene 30^3", scenes.default_scene(30), FR_DEF, FR_DEF),
                         ("2 dam-break 128^3", scenes.dam_break(128), FR_DAM, FR_CPU),
                         ("3 sphere-drop 256^3 (headline)", scenes.sphere_drop(256), FR_DAM, 0)):
    g = gpu_run(sc, fg)
    c = cpu_run(sc, fc, warm=(3 if fc > 10 else 1)) if fc else None
    rows.append((name, sc["pos"].shape[0], fg, g, fc, c))
    print(name, g, c, file=sys.stderr, flush=True)
print("| config | particles | GPU frames | GPU ms/frame (wall, host included) | GPU particle-steps/s | substeps/frame | CPU frames | CPU ms/frame | CPU particle-steps/s | CPU cores | speed-up |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, n, fg, g, fc, c in rows:
    if c:
        print(f"| {name} | {n} | {fg} | {g[0]:.2f} | {g[1]:.3e} | {g[2]:.2f} | {fc} | {c[0]:.1f} | {c[1]:.3e} | {c[3]} | {c[0] / g[0]:.0f}x |")
    else:
        print(f"| {name} | {n} | {fg} | {g[0]:.2f} | {g[1]:.3e} | {g[2]:.2f} | – | – (see bench.py cpu_baseline: 128^3 sample) | – | – | – |")
