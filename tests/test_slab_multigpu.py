"""z-slab decomposition on real GPUs: a 2-rank torchrun of scripts/slab_check.py, which compares the
slab run with the single-GPU run of the same scene AND with the unmodified reference engine on the host
(counts and pressure rows exact, per-particle positions rel-L2 <= 1e-4) while particles migrate across the slab
cut.  Skipped with < 2 GPUs; bench.py --gpus N runs the same kind of check on the scaling workload before it times
anything (`parity_ok` in its line), and profiles/ keeps the 2/4/8-rank logs of this script.  The scene stays under the
208 333 particles above which the reference aliases particle slots: with aliasing the reference
simulates a slightly different scene and free-running trajectories part chaotically within a few frames."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_slabs_match_single_gpu(peer):
    """peer=1: PCG / multigrid halo planes and scalars through CUDA-IPC peer memory (csrc/peer.cu);
    peer=0: the same exchanges as NCCL send/recv groups and all-reduces."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "slab_check.py"), "damz48", "12"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, FLIP_PEER=peer))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SLAB_CHECK OK" in r.stdout, r.stdout[-3000:]
    from oracle import refengine
    if refengine.available("golden"):
        assert "SLAB_CHECK OK (oracle compared)" in r.stdout, r.stdout[-3000:]
