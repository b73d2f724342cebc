"""Synthetic scenes of BASELINE.json / SURVEY.md §8d, generated deterministically with numpy.

Particles are seeded 8 per cell at the sub-cell offsets (+-dx/4)^3 around the cell centre (the
pattern of the reference's seeding, fluidsimulation.cpp:4528-4539) plus a small LCG jitter, and are
meant to be injected through loadMarkerParticleData / flip_load_particles, never through the
reference's racy addMeshFluid path.

All scenes use dx = 0.125 (dyadic) and gravity (0,-25,0).
"""
import numpy as np

DX = 0.125
GRAVITY = (0.0, -25.0, 0.0)
FRAME_DT = 1.0 / 30.0


def _lcg_jump(k):
    """(A, C) with s_{i+k} = A*s_i + C (mod 2^32) for the LCG below."""
    M = 0xFFFFFFFF
    A, C = 1, 0
    a, c = 1664525, 1013904223
    while k:
        if k & 1:
            A, C = (A * a) & M, (C * a + c) & M
        a, c = (a * a) & M, (c * a + c) & M
        k >>= 1
    return A, C


def lcg_uniform(n, seed, skip=0):
    """n numbers in [0,1): s <- s*1664525 + 1013904223 (mod 2^32), bits 8..23 of each state; `skip`
    leading numbers of the sequence are jumped over (so a z-slab can generate just its own part)."""
    a, c = np.uint64(1664525), np.uint64(1013904223)
    mask = np.uint64(0xFFFFFFFF)
    out = np.empty(max(n, 1), dtype=np.uint64)
    A0, C0 = _lcg_jump(int(skip) + 1)
    out[0] = np.uint64((A0 * int(seed) + C0) & 0xFFFFFFFF)
    filled = 1
    A, C = a, c  # jump by `filled` steps: s_{i+filled} = A*s_i + C
    while filled < n:
        m = min(filled, n - filled)
        out[filled:filled + m] = (out[:m] * A + C) & mask
        filled += m
        C = (A * C + C) & mask
        A = (A * A) & mask
    bits = (out[:n] >> np.uint64(8)) & np.uint64(0xFFFF)
    return bits.astype(np.float64) / 65536.0


def seed_cells(cells_ijk, dx, seed, velocity=None, skip_cells=0):
    """cells_ijk: (M,3) int array in the order the particles are emitted. Returns (pos, vel) float32.
    skip_cells: number of cells of the full scene that precede cells_ijk (jitter sequence offset)."""
    cells = np.asarray(cells_ijk, dtype=np.int64).reshape(-1, 3)
    m = cells.shape[0]
    q = 0.25 * dx
    sub = np.array([[sx, sy, sz] for sz in (-q, q) for sy in (-q, q) for sx in (-q, q)], dtype=np.float64)
    centre = (cells.astype(np.float64) + 0.5) * dx
    pos = centre[:, None, :] + sub[None, :, :]
    jit = lcg_uniform(m * 8 * 3, seed, skip=skip_cells * 24).reshape(m, 8, 3)
    pos = pos + 0.05 * q * (jit - 0.5)
    pos = pos.reshape(-1, 3).astype(np.float32)
    vel = np.zeros_like(pos)
    if velocity is not None:
        vel[:] = np.asarray(velocity, dtype=np.float32).reshape(-1, 3) if np.ndim(velocity) > 1 else np.asarray(velocity, dtype=np.float32)
    return pos, vel


def box_cells(i0, i1, j0, j1, k0, k1):
    k, j, i = np.meshgrid(np.arange(k0, k1), np.arange(j0, j1), np.arange(i0, i1), indexing="ij")
    return np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)


def default_scene(n=30, seed=12344):
    """Config 1: FluidManager's scene (src/FluidManager.cpp:47-69): fluid box = middle third."""
    a, b = n // 3, n - n // 3
    cells = box_cells(a, b, a, b, a, b)
    pos, vel = seed_cells(cells, DX, seed)
    return dict(name=f"default{n}", dims=(n, n, n), dx=DX, pos=pos, vel=vel)


def dam_break(n=128, seed=12345, krange=None, dx=DX):
    """Config 2 (n=128) / config 5 (n=512): column i∈[3s,35s) j∈[3s,67s) k∈[3s,n-3s), s=n/128.
    dx: cell width (every BASELINE config uses 0.125; the tests also run a non-power-of-two width).
    krange=(k0,k1): only the particles seeded in cell planes [k0,k1) (identical to that part of the full
    scene), for z-slab ranks that must not materialise 128 M particles each."""
    s = max(n // 128, 1)
    if n >= 128:
        box = (3 * s, 35 * s, 3 * s, 67 * s, 3 * s, n - 3 * s)
    else:  # small test sizes keep the same proportions
        box = (3, max(n // 4 + 3, 4), 3, max(n // 2 + 3, 4), 3, n - 3)
    i0, i1, j0, j1, k0, k1 = box
    skip = 0
    if krange is not None:
        ka, kb = max(k0, krange[0]), min(k1, krange[1])
        skip = max(ka - k0, 0) * (i1 - i0) * (j1 - j0)
        k0, k1 = ka, max(kb, ka)
    cells = box_cells(i0, i1, j0, j1, k0, k1)
    pos, vel = seed_cells(cells, dx, seed, skip_cells=skip)
    return dict(name=f"dambreak{n}", dims=(n, n, n), dx=dx, pos=pos, vel=vel, id_offset=8 * skip)


def dam_break_with_chamber(n=32, seed=12349):
    """The small dam break plus a closed chamber high above the floor, built from six wall boxes (none aligned with the
    grid) and filled to the brim: a liquid region enclosed by solids that touches no air -- what
    PressureSolver::_conditionSolidVelocityField (pressuresolver.cpp:124-244) looks for.  Returns the scene with the
    wall boxes under "obstacles" ((lo, hi) world coordinates) and the chamber's cells under "chamber_cells"."""
    sc = dam_break(n, seed=seed)
    dx = sc["dx"]
    o0, o1 = np.array([18.3, 15.3, 8.3]), np.array([27.7, 24.7, 20.7])      # outer box, in cells
    t = 2.4                                                                  # wall thickness
    walls = []
    for a in range(3):
        lo, hi = o0.copy(), o1.copy()
        hi[a] = o0[a] + t
        walls.append((tuple(lo * dx), tuple(hi * dx)))
        lo, hi = o0.copy(), o1.copy()
        lo[a] = o1[a] - t
        walls.append((tuple(lo * dx), tuple(hi * dx)))
    cells = box_cells(21, 25, 18, 22, 11, 18)          # the cells wholly inside the cavity
    pos, vel = seed_cells(cells, dx, seed + 1)
    sc = dict(sc)
    sc["name"] = f"dambreakchamber{n}"
    sc["pos"] = np.concatenate([sc["pos"], pos], axis=0)
    sc["vel"] = np.concatenate([sc["vel"], vel], axis=0)
    sc["obstacles"] = walls
    sc["chamber_cells"] = cells
    return sc


WEDGE_TRIANGLES = np.array([[0, 2, 1], [3, 4, 5], [0, 1, 4], [0, 4, 3], [1, 2, 5], [1, 5, 4], [2, 0, 3], [2, 3, 5]], dtype=np.int32)


def wedge_vertices(centre, angle, scale=1.0):
    """A triangular prism (closed mesh of 6 vertices, WEDGE_TRIANGLES) turned by `angle` about its z axis: the animated
    mesh obstacle of the moving-solid tests (every vertex has its own velocity when it turns)."""
    base = scale * np.array([[-0.45, -0.5, -0.9], [0.45, -0.5, -0.9], [0.0, 0.55, -0.9], [-0.45, -0.5, 0.9], [0.45, -0.5, 0.9], [0.0, 0.55, 0.9]])
    c, s = np.cos(angle), np.sin(angle)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return (base @ R.T + np.asarray(centre, dtype=np.float64)).astype(np.float32)


def dam_break_z(n=64, seed=12347):
    """A column against the low-z wall that collapses along +z: particles cross z-slab boundaries
    (migration test of the multi-GPU decomposition)."""
    cells = box_cells(3, n - 3, 3, max(n // 2 + 3, 4), 3, max(n // 4 + 3, 4))
    pos, vel = seed_cells(cells, DX, seed)
    return dict(name=f"dambreakz{n}", dims=(n, n, n), dx=DX, pos=pos, vel=vel)


def sphere_drop(n=256, seed=12346):
    """Config 3: pool i,k∈[3s,n-3s) j∈[2s,32s) + sphere radius 32s cells at (n/2, 160s, n/2), s=n/256."""
    s = n / 256.0
    lo, hi = 3, n - 3
    pool = box_cells(lo, hi, 2, max(int(round(32 * s)), 3), lo, hi)
    r = 32.0 * s
    cx, cy, cz = n / 2.0, 160.0 * s, n / 2.0
    b = box_cells(int(cx - r) - 1, int(cx + r) + 2, int(cy - r) - 1, int(cy + r) + 2, int(cz - r) - 1, int(cz + r) + 2)
    d2 = (b[:, 0] + 0.5 - cx) ** 2 + (b[:, 1] + 0.5 - cy) ** 2 + (b[:, 2] + 0.5 - cz) ** 2
    sph = b[d2 < r * r]
    cells = np.concatenate([pool, sph], axis=0)
    pos, vel = seed_cells(cells, DX, seed)
    vel[pool.shape[0] * 8:, 1] = -5.0
    return dict(name=f"spheredrop{n}", dims=(n, n, n), dx=DX, pos=pos, vel=vel)


def pressure_stress(n=256, seed=777):
    """Config 4: one particle at the centre of every cell of [2,n-2)^3 (phi_liquid<0 everywhere inside)
    and uniform(-1,1) particle velocities; the pressure stage alone is exercised on this state."""
    cells = box_cells(2, n - 2, 2, n - 2, 2, n - 2)
    pos = ((cells.astype(np.float64) + 0.5) * DX).astype(np.float32)
    vel = (2.0 * lcg_uniform(pos.size, seed) - 1.0).reshape(-1, 3).astype(np.float32)
    return dict(name=f"pressurestress{n}", dims=(n, n, n), dx=DX, pos=pos, vel=vel)


SCENES = {"default": default_scene, "dambreak": dam_break, "dambreakz": dam_break_z, "spheredrop": sphere_drop, "pressurestress": pressure_stress}
