"""Developer probe: lock-step parity report on small scenes (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from flipengine3d_b200 import scenes
import parity_common as pc

which = sys.argv[1] if len(sys.argv) > 1 else "default"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
isolate = (sys.argv[3] != "chain") if len(sys.argv) > 3 else True
if which == "default":
    sc = scenes.default_scene(30)
elif which.startswith("dam"):
    sc = scenes.dam_break(int(which[3:]))
elif which.startswith("sphere"):
    sc = scenes.sphere_drop(int(which[6:]))
t = time.time()
reps = pc.lockstep_frames(sc, frames=frames, isolate=isolate, verbose=True, preconditioner="jacobi")
print("elapsed", time.time() - t)
