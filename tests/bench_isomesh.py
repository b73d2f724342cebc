"""(test infrastructure: it runs the oracle as the CPU arm) Surface reconstruction timed: getIsomesh() on the device against ParticleMesher of the unmodified
reference on the host cores, same particles.  Usage: bench_isomesh.py [scene ...]   scene = dam128:2 | spheredrop256:1 ...
(name:subdivision).  Prints one JSON line per scene."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flipengine3d_b200 import scenes, engine as fe
from oracle import refengine

specs = sys.argv[1:] or ["dam64:2", "dam128:2", "spheredrop256:1"]
for spec in specs:
    name, sub = spec.split(":"); sub = int(sub)
    n = int("".join(ch for ch in name if ch.isdigit()))
    sc = scenes.dam_break(n) if name.startswith("dam") else (scenes.default_scene(n) if name.startswith("default") else scenes.sphere_drop(n))
    I, J, K = sc["dims"]
    sim = fe.FluidSimulation(I, J, K, sc["dx"])
    sim.addBodyForce(0, -25, 0)
    sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
    sim.setSurfaceSubdivisionLevel(sub)
    sim.initialize()
    for _ in range(3):
        sim.update(1 / 30)
    sim.getIsomesh()                                    # warm-up (allocations)
    ts = []
    for _ in range(3):
        sim.update(1 / 30)                              # a new particle state: the mesh cache does not apply
        sim.synchronize()
        t = time.perf_counter(); v, tr = sim.getIsomesh(); ts.append(time.perf_counter() - t)
    line = {"scene": name, "grid": [I, J, K], "subdivision": sub, "particles": sim.getNumMarkerParticles(),
            "vertices": int(v.shape[0]), "triangles": int(tr.shape[0]), "gpu_ms": round(1e3 * min(ts), 3),
            "gpu_ms_all": [round(1e3 * x, 3) for x in ts], "note": "getIsomesh() wall time incl. the device->host copy of the mesh"}
    if refengine.available("fast") and os.environ.get("ISOMESH_REF", "1") != "0" and sim.getNumMarkerParticles() <= 3_000_000:
        P = sim.getMarkerParticles()
        ref = refengine.RefEngine(sc["dims"], sc["dx"], P[:, :3].copy(), P[:, 3:].copy(), kind="fast")
        t = time.perf_counter(); rv, rt = ref.isomesh(subdivisions=sub); el = time.perf_counter() - t
        line.update(ref_cpu_ms=round(1e3 * el, 1), ref_threads=os.cpu_count(), ref_vertices=int(rv.shape[0]), ref_triangles=int(rt.shape[0]),
                    speedup=round(el / min(ts), 1))
    print(json.dumps(line), flush=True)
