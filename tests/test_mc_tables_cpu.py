"""The generated marching-cubes case table of the surface reconstruction (csrc/mc_tables.h, exported by the C-ABI as
flip_mc_case_table; host code, no GPU): every configuration uses exactly its crossed edges, every cube's patch is a
set of consistently oriented patches whose boundary edges lie in the faces of the cube (edges across its interior are
shared by two triangles), and the orientation is outward."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table():
    L = C.CDLL(os.path.join(ROOT, "flipengine3d_b200", "libflip_b200.so"))
    cnt = np.zeros(256, np.uint8)
    tri = np.zeros((256, 24), np.uint8)
    assert L.flip_mc_case_table(cnt.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p)) == 0
    return cnt, tri


def _edge_corners(e):
    a, q = e // 4, e % 4
    o = sorted([(a + 1) % 3, (a + 2) % 3])
    c0 = ((q & 1) << o[0]) | ((q >> 1) << o[1])
    return c0, c0 | (1 << a)


def test_case_table_properties():
    cnt, tri = _table()
    assert cnt[0] == 0 and cnt[255] == 0 and cnt.max() == 5
    pos = np.array([[c & 1, (c >> 1) & 1, c >> 2] for c in range(8)], dtype=np.float64)
    mid = np.array([(pos[_edge_corners(e)[0]] + pos[_edge_corners(e)[1]]) / 2 for e in range(12)])
    for cfg in range(256):
        crossed = {e for e in range(12) if ((cfg >> _edge_corners(e)[0]) & 1) != ((cfg >> _edge_corners(e)[1]) & 1)}
        t = tri[cfg, :3 * cnt[cfg]].reshape(-1, 3)
        assert set(t.ravel().tolist()) == crossed, cfg
        # directed edges: an interior edge of the patch appears once in each direction, a boundary edge (on a cube
        # face) once
        directed = {}
        for a, b, c in t.tolist():
            for u, v in ((a, b), (b, c), (c, a)):
                directed[(u, v)] = directed.get((u, v), 0) + 1
        for (u, v), n in directed.items():
            assert n == 1, (cfg, u, v)
            on_face = any(mid[u][x] == mid[v][x] and mid[u][x] in (0.0, 1.0) for x in range(3))
            if (v, u) not in directed:      # a boundary edge of the patch: it must lie in a face of the cube
                assert on_face, (cfg, u, v)
        # outward: normals point away from the inside corners
        if cnt[cfg]:
            inside = np.array([pos[c] for c in range(8) if (cfg >> c) & 1])
            outside = np.array([pos[c] for c in range(8) if not (cfg >> c) & 1])
            for a, b, c in t.tolist():
                n = np.cross(mid[b] - mid[a], mid[c] - mid[a])
                ctr = (mid[a] + mid[b] + mid[c]) / 3
                # the corners of the crossed edges of this triangle
                score = 0.0
                for e in (a, b, c):
                    c0, c1 = _edge_corners(e)
                    cin, cout = (c0, c1) if (cfg >> c0) & 1 else (c1, c0)
                    score += float(np.dot(n, pos[cout] - pos[cin]))
                assert score > 0.0, (cfg, a, b, c)
