"""Host-side logic of the z-slab decomposition on CPU: the slab split exported by the C-ABI, the
slab-restricted scene generator, and a world_size-2 gloo run of the rendezvous the multi-GPU path
uses (unique-id broadcast, per-rank slab ranges, per-rank particle ownership)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flipengine3d_b200 import engine as fe
from flipengine3d_b200 import scenes


def test_slab_range_is_a_balanced_partition(built):
    for K, n in ((512, 8), (512, 4), (256, 2), (96, 4), (30, 3), (7, 7)):
        edges = [fe.slab_range(K, n, r) for r in range(n)]
        assert edges[0][0] == 0 and edges[-1][1] == K
        for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
            assert a1 == b0 and a1 > a0
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(IndexError):
        fe.slab_range(64, 2, 2)


def test_slab_scene_parts_concatenate_to_the_full_scene():
    full = scenes.dam_break(32)
    parts = [scenes.dam_break(32, krange=fe.slab_range(32, 4, r) if False else (8 * r, 8 * r + 8)) for r in range(4)]
    assert np.array_equal(np.concatenate([p["pos"] for p in parts]), full["pos"])
    assert np.array_equal(scenes.lcg_uniform(1000, 5)[123:], scenes.lcg_uniform(877, 5, skip=123))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 owns the rendezvous blob (the ncclUniqueId on a GPU box; any 128 bytes here)
    ident = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    K = 64
    k0, k1 = fe.slab_range(K, world, rank)
    sc = scenes.dam_break(K, krange=(k0, k1))
    kcell = np.floor(sc["pos"][:, 2].astype(np.float64) / sc["dx"]).astype(np.int64)
    owned_ok = bool(((kcell >= k0) & (kcell < k1)).all())
    counts = [None] * world
    dist.all_gather_object(counts, (k0, k1, sc["pos"].shape[0], owned_ok, ident[0] == bytes(range(128))))
    if rank == 0:
        out.put(counts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_rendezvous_and_ownership_gloo(built):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    counts = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (a0, a1, na, oka, ida), (b0, b1, nb, okb, idb) = counts
    assert (a0, a1, b0, b1) == (0, 32, 32, 64)
    assert oka and okb and ida and idb
    assert na + nb == scenes.dam_break(64)["pos"].shape[0]
