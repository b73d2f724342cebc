"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/flip_b200.h declares, validates arguments before touching CUDA, and refuses to run without
a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flip_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flip_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("flip_create", "flip_destroy", "flip_initialize", "flip_update", "flip_load_particles",
                 "flip_add_marker_particle", "flip_get_particles", "flip_get_velocity_field", "flip_run_stage"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    from flipengine3d_b200 import engine
    L = engine.load_library()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_header_cites_reference_for_each_entry_point():
    src = open(HEADER).read()
    # every declaration is preceded by a comment that names a reference file:line (or says it is new)
    assert src.count("fluidsimulation.cpp:") >= 12
    assert "velocityadvector.cpp:38" in src and "pressuresolver.cpp:44" in src


def test_bad_arguments_are_rejected_before_cuda(built):
    from flipengine3d_b200 import engine
    L = engine.load_library()
    h = C.c_void_p()
    assert L.flip_create(C.byref(h), 0, 8, 8, 0.125, 0) == engine.FLIP_ERR_DOMAIN
    assert L.flip_create(C.byref(h), 8, 8, 8, -1.0, 0) == engine.FLIP_ERR_DOMAIN
    assert b"greater than 0" in L.flip_create_error()


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    from flipengine3d_b200 import engine
    with pytest.raises(engine.FlipCudaError) as e:
        engine.FluidSimulation(8, 8, 8, 0.125)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product package must never import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "flipengine3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("parity oracle", "").replace("the oracle's", "").replace("oracle is built", ""), (dirpath, f)
                assert "libflipref" not in txt, (dirpath, f)


def test_step_stats_struct_layout_matches_header(built):
    from flipengine3d_b200 import engine
    # 8 int32 + 3 double, naturally aligned
    assert C.sizeof(engine.StepStats) == 8 * 4 + 3 * 8


def test_cpp_facade_example_links_and_refuses_to_run_without_a_device(built):
    """include/fluidsimulation_b200.hpp (the FluidSimulation-named C++ façade) and examples/fluidmanager_headless.cpp
    (the reference's FluidManager scene without DXViewer) compile against the C-ABI; without a GPU the program
    exits with the no-CPU-fallback error instead of computing anything."""
    import shutil
    import subprocess
    import torch
    exe = os.path.join(ROOT, "build", "fluidmanager_headless")
    if not os.path.exists(exe):
        if shutil.which("g++") is None or shutil.which("make") is None:
            pytest.skip("no C++ toolchain and no prebuilt example")
        subprocess.run(["make", "-C", os.path.join(ROOT, "flipengine3d_b200", "csrc")], check=True, stdout=subprocess.DEVNULL)
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by tests/test_gpu_parity.py")
    r = subprocess.run([exe, "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.parametrize("fixture,n", [("dam24_stages.npz", 24), ("default30_static.npz", 30)])
def test_host_static_inputs_match_the_reference(built, fixture, n):
    """The host code that prepares the static inputs (csrc/static_host.cpp through flip_static_inputs, no device):
    from the reference's solid SDF the face weights and the near-solid mask come out bit for bit; the built-in box
    SDF has the reference's sign everywhere and its values in the 3dx band the step consumes."""
    from flipengine3d_b200 import engine as fe
    g = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    dx = 0.125
    s = fe.static_inputs(n, n, n, dx, solid_phi=g["solid_phi"])
    for c in "UVW":
        assert np.array_equal(s["weight" + c], g["weight" + c]), c
    assert np.array_equal(s["near_solid"].ravel(), g["near_solid"].ravel())
    own = fe.static_inputs(n, n, n, dx)
    assert np.array_equal(own["solid_phi"] < 0, g["solid_phi"] < 0)
    band = np.abs(g["solid_phi"]) < 3 * dx
    assert np.max(np.abs(own["solid_phi"][band] - g["solid_phi"][band])) < 1e-5
    for c in "UVW":
        assert np.max(np.abs(own["weight" + c] - g["weight" + c])) < 1e-4, c
    assert np.array_equal(own["near_solid"].ravel(), g["near_solid"].ravel())
