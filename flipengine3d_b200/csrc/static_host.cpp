// Static inputs of the step, computed once on the host and uploaded: the solid SDF of
// the domain box, the variational face weights derived from it, and the coarse near-solid mask.
#include <algorithm>
#include <cmath>
#include <vector>
#include "flip_internal.h"
#include "mc_tables.h"

namespace flip {

// Solid SDF of the reference's domain: the simulation box inset by 1.5dx + 5e-5 per side
// (FluidSimulation::_getBoundaryAABB, fluidsimulation.cpp:2834-2839; AABB::expand aabb.cpp:122-128),
// negated so that the fluid side is positive (fluidsimulation.cpp:2927-2936).  The reference
// rasterises the 12 triangles of that box (meshlevelset.cpp:572-601) and stores the exact
// point-to-triangle distance inside a band around each triangle and an upper bound elsewhere;
// the distance to the closed box surface computed here is the same quantity (to float rounding)
// inside the band, and stays positive where the reference keeps its positive upper bound, which is
// all that the step consumes (signs and near-wall values).  Parity tests upload the oracle's array
// instead (flip_set_solid_sdf).
void build_box_solid_sdf(const Dims &d, std::vector<float> &phi) {
    double eps = 1e-4;
    double v = -3 * d.dx - eps;
    float lo = 0.0f - (float)(0.5 * v);
    double w = d.I * d.dx + v, h = d.J * d.dx + v, dp = d.K * d.dx + v;
    float hix = lo + (float)w, hiy = lo + (float)h, hiz = lo + (float)dp;
    phi.resize((size_t)d.nN);
    size_t idx = 0;
    for (int k = 0; k <= d.K; k++) {
        float z = (float)(k * d.dx);
        for (int j = 0; j <= d.J; j++) {
            float y = (float)(j * d.dx);
            for (int i = 0; i <= d.I; i++, idx++) {
                float x = (float)(i * d.dx);
                float ox = std::max(std::max(lo - x, x - hix), 0.0f);
                float oy = std::max(std::max(lo - y, y - hiy), 0.0f);
                float oz = std::max(std::max(lo - z, z - hiz), 0.0f);
                float val;
                if (ox > 0 || oy > 0 || oz > 0) {
                    val = -std::sqrt(ox * ox + oy * oy + oz * oz);
                } else {
                    float m = std::min(std::min(std::min(x - lo, hix - x), std::min(y - lo, hiy - y)), std::min(z - lo, hiz - z));
                    val = m;
                }
                phi[idx] = val;
            }
        }
    }
}

// LevelsetUtils::fractionInside(float,float)  levelsetutils.cpp:39-51
static float fractionInside2(float phiLeft, float phiRight) {
    if (phiLeft < 0 && phiRight < 0) return 1;
    if (phiLeft < 0 && phiRight >= 0) return phiLeft / (phiLeft - phiRight);
    if (phiLeft >= 0 && phiRight < 0) return phiRight / (phiRight - phiLeft);
    return 0;
}

static void cycle4(float *a) {
    float t = a[0];
    a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = t;
}

// LevelsetUtils::fractionInside(float bl, br, tl, tr)  levelsetutils.cpp:62-142 (marching-squares
// area of the negative region of a bilinear patch).
static float fractionInside4(float phibl, float phibr, float phitl, float phitr) {
    int insideCount = (phibl < 0 ? 1 : 0) + (phitl < 0 ? 1 : 0) + (phibr < 0 ? 1 : 0) + (phitr < 0 ? 1 : 0);
    float list[4] = {phibl, phibr, phitr, phitl};
    if (insideCount == 4) return 1;
    if (insideCount == 3) {
        while (list[0] < 0) cycle4(list);
        float side0 = 1 - fractionInside2(list[0], list[3]);
        float side1 = 1 - fractionInside2(list[0], list[1]);
        return 1.0f - 0.5f * side0 * side1;
    }
    if (insideCount == 2) {
        while (list[0] >= 0 || !(list[1] < 0 || list[2] < 0)) cycle4(list);
        if (list[1] < 0) {
            float sideLeft = fractionInside2(list[0], list[3]);
            float sideRight = fractionInside2(list[1], list[2]);
            return 0.5f * (sideLeft + sideRight);
        }
        float middlePoint = 0.25f * (list[0] + list[1] + list[2] + list[3]);
        if (middlePoint < 0) {
            float area = 0;
            float side1 = 1 - fractionInside2(list[0], list[3]);
            float side3 = 1 - fractionInside2(list[2], list[3]);
            area += 0.5f * side1 * side3;
            float side2 = 1 - fractionInside2(list[2], list[1]);
            float side0 = 1 - fractionInside2(list[0], list[1]);
            area += 0.5f * side0 * side2;
            return 1.0f - area;
        }
        float area = 0;
        float side0 = fractionInside2(list[0], list[1]);
        float side1 = fractionInside2(list[0], list[3]);
        area += 0.5f * side0 * side1;
        float side2 = fractionInside2(list[2], list[1]);
        float side3 = fractionInside2(list[2], list[3]);
        area += 0.5f * side2 * side3;
        return area;
    }
    if (insideCount == 1) {
        while (list[0] >= 0) cycle4(list);
        float side0 = fractionInside2(list[0], list[3]);
        float side1 = fractionInside2(list[0], list[1]);
        return 0.5f * side0 * side1;
    }
    return 0;
}

static inline float clamp01(float w) { return std::max(0.0f, std::min(w, 1.0f)); }

// FluidSimulation::_updateWeightGridThread (fluidsimulation.cpp:3690-3730) over
// MeshLevelSet::getFaceWeightU/V/W (meshlevelset.cpp:357-387).  The cell-centre weight multiplies
// the (zero) velocities of static solids only (pressuresolver.cpp:595-600) and is not built.
void build_weights(const Dims &d, const std::vector<float> &phi, std::vector<float> &wU, std::vector<float> &wV,
                   std::vector<float> &wW, std::vector<float> &wC) {
    const int ni = d.I + 1, nj = d.J + 1;
    auto P = [&](int i, int j, int k) { return phi[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)]; };
    wU.resize(d.nU); wV.resize(d.nV); wW.resize(d.nW); wC.clear();
    size_t idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i <= d.I; i++, idx++)
                wU[idx] = clamp01(1.0f - fractionInside4(P(i, j, k), P(i, j + 1, k), P(i, j, k + 1), P(i, j + 1, k + 1)));
    idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j <= d.J; j++)
            for (int i = 0; i < d.I; i++, idx++)
                wV[idx] = clamp01(1.0f - fractionInside4(P(i, j, k), P(i, j, k + 1), P(i + 1, j, k), P(i + 1, j, k + 1)));
    idx = 0;
    for (int k = 0; k <= d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++, idx++)
                wW[idx] = clamp01(1.0f - fractionInside4(P(i, j, k), P(i, j + 1, k), P(i + 1, j, k), P(i + 1, j + 1, k)));
}

// FluidSimulation::_updateNearSolidGrid (fluidsimulation.cpp:3083-3127): coarse cells (3dx) holding a
// node with |phi_solid| < 3dx, dilated ceil(CFL/3) times with the 6-neighbourhood (GridUtils::featherGrid6).
void build_near_solid(const Dims &d, const std::vector<float> &phi, int factor, int band, double cfl,
                      std::vector<unsigned char> &grid, int &gi, int &gj, int &gk) {
    double cell = factor * d.dx;
    gi = (int)std::ceil((d.I * d.dx) / cell);
    gj = (int)std::ceil((d.J * d.dx) / cell);
    gk = (int)std::ceil((d.K * d.dx) / cell);
    grid.assign((size_t)gi * gj * gk, 0);
    float maxd = (float)(band * d.dx);
    const int ni = d.I + 1, nj = d.J + 1;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++) {
                float v = phi[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)];
                if (std::abs(v) < maxd) grid[(size_t)(i / factor) + (size_t)gi * ((size_t)(j / factor) + (size_t)gj * (k / factor))] = 1;
            }
    int numlayers = (int)std::ceil((float)cfl / (float)factor);
    for (int l = 0; l < numlayers; l++) {
        std::vector<unsigned char> tmp = grid;
        for (int k = 0; k < gk; k++)
            for (int j = 0; j < gj; j++)
                for (int i = 0; i < gi; i++) {
                    if (!tmp[(size_t)i + (size_t)gi * ((size_t)j + (size_t)gj * k)]) continue;
                    const int nb[6][3] = {{i - 1, j, k}, {i + 1, j, k}, {i, j - 1, k}, {i, j + 1, k}, {i, j, k - 1}, {i, j, k + 1}};
                    for (auto &q : nb) {
                        if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= gi || q[1] >= gj || q[2] >= gk) continue;
                        grid[(size_t)q[0] + (size_t)gi * ((size_t)q[1] + (size_t)gj * q[2])] = 1;
                    }
                }
    }
}

}  // namespace flip

// the generated marching-cubes case table of the surface reconstruction (mc_tables.h), for inspection and tests
extern "C" int flip_mc_case_table(unsigned char counts[256], unsigned char edge_triples[256 * 24]) {
    unsigned char tris[256][flip::MC_MAX_TRIS * 3];
    unsigned char cnt[256];
    flip::build_mc_tables(cnt, tris);
    if (counts) memcpy(counts, cnt, 256);
    if (edge_triples) memcpy(edge_triples, tris, sizeof(tris));
    return FLIP_OK;
}
