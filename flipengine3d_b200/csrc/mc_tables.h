// The marching-cubes case table of csrc/mesher.cu, generated instead of tabulated (host code, no CUDA).
// Corner c of a cube: bit 0 = +x, bit 1 = +y, bit 2 = +z; a configuration has bit c set where corner c is INSIDE.
// Edge e = axis * 4 + q, q = the bits of its lower corner on the two other axes (lower axis first).
// For every configuration the crossed edges are joined into closed loops by walking the six faces (a face with two
// crossed edges joins them; an ambiguous face -- four crossed edges, inside corners on a diagonal -- cuts off each
// inside corner, a rule that depends on the face's corner signs only, so the two cubes sharing the face agree and the
// surface is watertight), each loop is oriented with its normal pointing from the inside corners to the outside ones
// and fan-triangulated.
#pragma once
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

namespace flip {

static constexpr int MC_MAX_TRIS = 8;

struct CubeTopology {
    int edgeCorner[12][2];
    CubeTopology() {
        for (int a = 0; a < 3; a++)
            for (int q = 0; q < 4; q++) {
                const int o1 = (a + 1) % 3, o2 = (a + 2) % 3;
                const int lo1 = std::min(o1, o2), lo2 = std::max(o1, o2);
                const int c0 = ((q & 1) << lo1) | ((q >> 1) << lo2);
                edgeCorner[a * 4 + q][0] = c0;
                edgeCorner[a * 4 + q][1] = c0 | (1 << a);
            }
    }
    int edgeBetween(int c0, int c1) const {
        for (int e = 0; e < 12; e++)
            if ((edgeCorner[e][0] == c0 && edgeCorner[e][1] == c1) || (edgeCorner[e][0] == c1 && edgeCorner[e][1] == c0)) return e;
        return -1;
    }
};

inline void build_mc_tables(unsigned char count[256], unsigned char tris[256][MC_MAX_TRIS * 3]) {
    CubeTopology T;
    const double cpos[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
    for (int cfg = 0; cfg < 256; cfg++) {
        count[cfg] = 0;
        memset(tris[cfg], 0, MC_MAX_TRIS * 3);
        auto inside = [&](int c) { return (cfg >> c) & 1; };
        // segments on the faces
        std::vector<std::pair<int, int>> seg;
        for (int a = 0; a < 3; a++)
            for (int side = 0; side < 2; side++) {
                const int o1 = (a + 1) % 3, o2 = (a + 2) % 3;
                // the four corners of the face in cyclic order
                int fc[4];
                const int base = side << a;
                fc[0] = base; fc[1] = base | (1 << o1); fc[2] = base | (1 << o1) | (1 << o2); fc[3] = base | (1 << o2);
                int fe[4], crossed[4], n = 0;
                for (int q = 0; q < 4; q++) {
                    fe[q] = T.edgeBetween(fc[q], fc[(q + 1) % 4]);
                    crossed[q] = inside(fc[q]) != inside(fc[(q + 1) % 4]);
                    n += crossed[q];
                }
                if (n == 2) {
                    int e[2], m = 0;
                    for (int q = 0; q < 4; q++) if (crossed[q]) e[m++] = fe[q];
                    seg.push_back({e[0], e[1]});
                } else if (n == 4) {
                    // ambiguous face: cut off each INSIDE corner (edge before it, edge after it)
                    for (int q = 0; q < 4; q++)
                        if (inside(fc[q])) seg.push_back({fe[(q + 3) % 4], fe[q]});
                }
            }
        // trace the loops
        std::vector<char> used(seg.size(), 0);
        for (size_t s0 = 0; s0 < seg.size(); s0++) {
            if (used[s0]) continue;
            std::vector<int> loop;
            int start = seg[s0].first, cur = seg[s0].second;
            used[s0] = 1;
            loop.push_back(start);
            while (cur != start) {
                loop.push_back(cur);
                bool found = false;
                for (size_t s = 0; s < seg.size() && !found; s++) {
                    if (used[s]) continue;
                    if (seg[s].first == cur) { cur = seg[s].second; used[s] = 1; found = true; }
                    else if (seg[s].second == cur) { cur = seg[s].first; used[s] = 1; found = true; }
                }
                if (!found) break;
            }
            if (loop.size() < 3) continue;
            // orientation: the normal points from the inside corners to the outside
            double mid[16][3], nrm[3] = {0, 0, 0};
            for (size_t v = 0; v < loop.size(); v++)
                for (int x = 0; x < 3; x++) mid[v][x] = 0.5 * (cpos[T.edgeCorner[loop[v]][0]][x] + cpos[T.edgeCorner[loop[v]][1]][x]);
            for (size_t v = 0; v < loop.size(); v++) {
                const double *p = mid[v], *q = mid[(v + 1) % loop.size()];
                nrm[0] += (p[1] - q[1]) * (p[2] + q[2]); nrm[1] += (p[2] - q[2]) * (p[0] + q[0]); nrm[2] += (p[0] - q[0]) * (p[1] + q[1]);
            }
            double dotsum = 0.0;
            for (size_t v = 0; v < loop.size(); v++) {
                const int e = loop[v];
                const int cin = inside(T.edgeCorner[e][0]) ? T.edgeCorner[e][0] : T.edgeCorner[e][1];
                const int cout = inside(T.edgeCorner[e][0]) ? T.edgeCorner[e][1] : T.edgeCorner[e][0];
                for (int x = 0; x < 3; x++) dotsum += nrm[x] * (cpos[cout][x] - cpos[cin][x]);
            }
            if (dotsum < 0.0) std::reverse(loop.begin(), loop.end());
            // fan apex: one that leaves no triangle lying inside a face of the cube (a loop that crosses an ambiguous
            // face twice has four of its vertices in that face)
            auto in_one_face = [&](int e0, int e1, int e2) {
                for (int x = 0; x < 3; x++)
                    for (int side = 0; side < 2; side++) {
                        bool all = true;
                        for (int e : {e0, e1, e2})
                            all = all && cpos[T.edgeCorner[e][0]][x] == side && cpos[T.edgeCorner[e][1]][x] == side;
                        if (all) return true;
                    }
                return false;
            };
            size_t apex = 0;
            for (size_t a = 0; a < loop.size(); a++) {
                bool ok = true;
                for (size_t v = 1; v + 1 < loop.size() && ok; v++)
                    ok = !in_one_face(loop[a], loop[(a + v) % loop.size()], loop[(a + v + 1) % loop.size()]);
                if (ok) { apex = a; break; }
            }
            for (size_t v = 1; v + 1 < loop.size(); v++) {
                if (count[cfg] >= MC_MAX_TRIS) break;
                unsigned char *t = tris[cfg] + 3 * count[cfg];
                t[0] = (unsigned char)loop[apex]; t[1] = (unsigned char)loop[(apex + v) % loop.size()];
                t[2] = (unsigned char)loop[(apex + v + 1) % loop.size()];
                count[cfg]++;
            }
        }
    }
}


}  // namespace flip
