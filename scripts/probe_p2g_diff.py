"""Developer probe: P2G scatter against the gather on the same developed state (run on a GPU box).
Usage: probe_p2g_diff.py [scene] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flipengine3d_b200 import scenes, engine as fe

which = sys.argv[1] if len(sys.argv) > 1 else "sphere256"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = scenes.sphere_drop(int(which[6:])) if which.startswith("sphere") else scenes.dam_break(int(which[3:]))
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
os.environ["FLIP_P2G_GATHER"] = "1"
for f in range(frames):
    sim.update(1 / 30)
sim.begin_frame(1 / 30)
dt = sim.begin_substep()
sim.stage("liquid_sdf", dt)
G = {n: sim.array(n) for n in ("U", "V", "W", "validU", "validV", "validW", "liquid_phi")}
del os.environ["FLIP_P2G_GATHER"]
sim.stage("liquid_sdf", dt)
S = {n: sim.array(n) for n in G}
print("phi mismatch", int(np.count_nonzero(G["liquid_phi"] != S["liquid_phi"])))
P = sim.getMarkerParticles()
cells = np.floor(P[:, :3].astype(np.float64) / sc["dx"]).astype(np.int64)
lin = cells[:, 0] + I * (cells[:, 1] + J * cells[:, 2])
cnt = np.bincount(lin, minlength=I * J * K)
print("max particles per cell", cnt.max(), "cells over 56:", int((cnt > 56).sum()))
for n in "UVW":
    d = np.abs(G[n].astype(np.float64) - S[n].astype(np.float64))
    print(n, "valid hamming", int(np.count_nonzero(G["valid" + n] != S["valid" + n])), "max", d.max(),
          "rel_l2", np.sqrt((d ** 2).sum() / (G[n].astype(np.float64) ** 2).sum()))
    bad = np.argwhere(d > 1e-3)
    print("  faces with |diff| > 1e-3:", len(bad))
    for (k, j, i) in bad[:12]:
        ci, cj, ck = min(i, I - 1), min(j, J - 1), min(k, K - 1)
        nb = cnt.reshape(K, J, I)[max(ck - 1, 0):ck + 2, max(cj - 1, 0):cj + 2, max(ci - 1, 0):ci + 2]
        print("   ", n, (i, j, k), "gather", G[n][k, j, i], "scatter", S[n][k, j, i], "i%8,j%8,k%8", i % 8, j % 8, k % 8,
              "max cnt around", nb.max())
