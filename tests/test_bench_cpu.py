"""bench.py's reference arm on CPU: one JSON line with the contract's keys, from rank 0 alone under torchrun."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refengine  # noqa: E402

pytestmark = pytest.mark.skipif(not (refengine.available("fast") or refengine.available("golden")), reason="oracle/_ref not built")


def _check(line, n):
    d = json.loads(line)
    assert d["impl"] == "reference" and d["n_gpus"] == n and d["steps"] == 1
    assert d["metric"] == "particle_steps_per_s" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["same_config"] is True and d["config"]["frames_timed"] >= 1
    assert d["cpu_baseline"]["frames_timed"] == d["config"]["frames_timed"]
    return d


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "spheredrop", "--grid", "32"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    assert _check(lines[0], 1)["config"]["workload"] == "spheredrop32"


def test_reference_arm_under_torchrun_prints_from_rank_zero_only():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29545", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--workload", "dambreak", "--grid", "32"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    assert _check(lines[0], 2)["config"]["workload"] == "dambreak32"


def test_default_workloads(monkeypatch):
    """BENCH (1-GPU box): spheredrop256.  The scaling series runs ONE workload at every N: dambreak512 for N > 1 and for
    N = 1 on a box that shows several GPUs, so the driver's v_N / (N v_1) compares like with like."""
    import argparse
    import bench
    a = argparse.Namespace(workload=None, grid=256)
    monkeypatch.setattr(bench, "visible_gpus", lambda: 1)
    assert bench.pick_workload(a, 1)[:2] == ("spheredrop", 256)
    assert bench.pick_workload(a, 4)[:2] == ("dambreak", 512)
    monkeypatch.setattr(bench, "visible_gpus", lambda: 8)
    assert bench.pick_workload(a, 1)[:2] == ("dambreak", 512)
    assert bench.pick_workload(argparse.Namespace(workload="spheredrop", grid=64), 1)[:2] == ("spheredrop", 64)


def test_algorithmic_bytes_follow_the_survey():
    import bench
    ab = bench.algorithmic_bytes(1000, (4, 4, 4), 10, 7)
    nf = 5 * 4 * 4 * 3
    assert ab["sdf_p2g"] == 24 * 1000 + 5 * nf + 4 * 64
    assert ab["g2p_advance"] == 48 * 1000 + 8 * 7
    assert ab["pcg_spmv"] == 360 and ab["pcg_iter"] == 1240 + 160 and ab["precond"] == 320
