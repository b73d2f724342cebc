// Static inputs of the step, computed once on the host and uploaded: the solid SDF of
// the domain box, the variational face weights derived from it, and the coarse near-solid mask.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
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

// Fraction of a segment on which a linear function with end values a, b is negative
// (LevelsetUtils::fractionInside(float,float), levelsetutils.cpp:39-51).
static float negative_fraction(float a, float b) {
    const bool na = a < 0, nb = b < 0;
    if (na && nb) return 1;
    if (na) return a / (a - b);
    if (nb) return b / (b - a);
    return 0;
}

// Area fraction of a unit square on which the bilinear patch through its corner values is negative, by the sixteen
// marching-squares cases (LevelsetUtils::fractionInside(bl, br, tl, tr), levelsetutils.cpp:62-142; the reference rotates
// the corner list in a loop until it reaches the canonical orientation of the case -- here the rotation of every
// sign mask is tabulated).  Corners in cyclic order q[0..3] = bl, br, tr, tl; bit c of the mask: corner c negative.
// The float operations and their order are the reference's: the weights derived from this are bit-identical.
namespace {
struct SquareCase {
    signed char kind;      // -1: none negative, 0: one, 1: two adjacent, 2: two opposite, 3: three, 4: all
    signed char first;     // index of the corner that comes first in the canonical orientation
};
SquareCase square_case(unsigned mask) {
    const int count = __builtin_popcount(mask);
    auto neg = [&](int c) { return (mask >> (c & 3)) & 1u; };
    if (count == 0) return {-1, 0};
    if (count == 4) return {4, 0};
    for (int s = 0; s < 4; s++) {
        if (count == 1 && neg(s)) return {0, (signed char)s};                       // the negative corner first
        if (count == 3 && !neg(s)) return {3, (signed char)s};                      // the positive corner first
        if (count == 2 && neg(s) && neg(s + 1)) return {1, (signed char)s};         // the negative edge first
    }
    for (int s = 0; s < 4; s++)
        if (neg(s) && neg(s + 2)) return {2, (signed char)s};                       // the diagonal, lowest index first
    return {-1, 0};
}
const struct SquareTable {
    SquareCase c[16];
    SquareTable() { for (unsigned m = 0; m < 16; m++) c[m] = square_case(m); }
} kSquare;
}  // namespace

static float negative_area(float bl, float br, float tl, float tr) {
    const float q[4] = {bl, br, tr, tl};
    const unsigned mask = (bl < 0 ? 1u : 0u) | (br < 0 ? 2u : 0u) | (tr < 0 ? 4u : 0u) | (tl < 0 ? 8u : 0u);
    const SquareCase sc = kSquare.c[mask];
    const float v0 = q[sc.first & 3], v1 = q[(sc.first + 1) & 3], v2 = q[(sc.first + 2) & 3], v3 = q[(sc.first + 3) & 3];
    switch (sc.kind) {
        case 4: return 1;
        case 3: {       // a positive corner (v0) cut off by its two edges
            const float cutA = 1 - negative_fraction(v0, v3), cutB = 1 - negative_fraction(v0, v1);
            return 1.0f - 0.5f * cutA * cutB;
        }
        case 1: {       // the edge v0-v1 is negative: a trapezoid
            const float along03 = negative_fraction(v0, v3), along12 = negative_fraction(v1, v2);
            return 0.5f * (along03 + along12);
        }
        case 2: {       // v0 and v2 negative: the sign of the centre decides whether they are joined
            const float centre = 0.25f * (v0 + v1 + v2 + v3);
            if (centre < 0) {
                float positive = 0;
                const float a3 = 1 - negative_fraction(v0, v3), c3 = 1 - negative_fraction(v2, v3);
                positive += 0.5f * a3 * c3;
                const float c1 = 1 - negative_fraction(v2, v1), a1 = 1 - negative_fraction(v0, v1);
                positive += 0.5f * a1 * c1;
                return 1.0f - positive;
            }
            float negative = 0;
            const float a1 = negative_fraction(v0, v1), a3 = negative_fraction(v0, v3);
            negative += 0.5f * a1 * a3;
            const float c1 = negative_fraction(v2, v1), c3 = negative_fraction(v2, v3);
            negative += 0.5f * c1 * c3;
            return negative;
        }
        case 0: {       // a negative corner (v0) cut off by its two edges
            const float cutA = negative_fraction(v0, v3), cutB = negative_fraction(v0, v1);
            return 0.5f * cutA * cutB;
        }
        default: return 0;
    }
}

static inline float clamp01(float w) { return std::max(0.0f, std::min(w, 1.0f)); }

// Volume fraction of a tetrahedron on which the linear function through its corner values is negative
// (LevelsetUtils::volumeFraction(float, float, float, float), levelsetutils.cpp:213-226, over the sorted-tetrahedron
// and sorted-prism formulas of levelsetutils.h:72-85).  Only the sorted values enter; they are put in order by the
// reference's five-comparator network, so that equal values (and signed zeros) end up where they do there.
static float negative_tet_volume(float a, float b, float c, float d) {
    float v[4] = {a, b, c, d};
    static const unsigned char net[5][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}, {1, 2}};
    for (auto &cmp : net)
        if (v[cmp[0]] > v[cmp[1]]) std::swap(v[cmp[0]], v[cmp[1]]);
    auto corner = [](float p, float q, float r, float s) { return p * p * p / ((p - q) * (p - r) * (p - s)); };
    if (v[3] <= 0) return 1;
    if (v[2] <= 0) return 1 - corner(v[3], v[2], v[1], v[0]);      // one positive corner cut off
    if (v[1] <= 0) {                                               // two and two: a prism
        const float e02 = v[0] / (v[0] - v[2]), e03 = v[0] / (v[0] - v[3]);
        const float e13 = v[1] / (v[1] - v[3]), e12 = v[1] / (v[1] - v[2]);
        return e02 * e03 * (1 - e12) + e03 * (1 - e13) * e12 + e13 * e12;
    }
    if (v[0] <= 0) return corner(v[0], v[1], v[2], v[3]);          // one negative corner cut off
    return 0;
}

// Solid fraction of a cell from the eight nodal values of the solid SDF (MeshLevelSet::_getCellWeight,
// meshlevelset.cpp:1490-1513): the mean of the two decompositions of the cube into five tetrahedra
// (LevelsetUtils::volumeFraction of eight values, levelsetutils.cpp:243-259), summed in the reference's order; the
// central tetrahedron of each decomposition counts twice.  q[x + 2 y + 4 z] = phi(i + x, j + y, k + z).
static float negative_cell_volume(const float q[8]) {
    int negatives = 0;
    for (int n = 0; n < 8; n++) negatives += q[n] < 0 ? 1 : 0;
    if (negatives == 8) return 1.0f;
    if (negatives == 0) return 0.0f;
    static const unsigned char tets[10][5] = {       // four corners and the multiplicity
        {0, 4, 5, 6, 1}, {0, 5, 1, 3, 1}, {0, 2, 6, 3, 1}, {5, 6, 7, 3, 1}, {0, 6, 5, 3, 2},
        {1, 5, 4, 7, 1}, {1, 4, 0, 2, 1}, {1, 3, 7, 2, 1}, {4, 7, 6, 2, 1}, {1, 7, 4, 2, 2}};
    float sum = 0.0f;
    bool first = true;
    for (auto &t : tets) {
        float part = negative_tet_volume(q[t[0]], q[t[1]], q[t[2]], q[t[3]]);
        if (t[4] == 2) part = 2 * part;
        sum = first ? part : sum + part;
        first = false;
    }
    return sum / 12.0f;
}

// The cell-centre entry of the weight grid (_updateWeightGridThread, CENTER branch, fluidsimulation.cpp:3720-3727): it
// multiplies the solid face velocities in the divergence (pressuresolver.cpp:595-613), so it is only built when solid
// velocities are in use.
void build_center_weights(const Dims &d, const std::vector<float> &phi, std::vector<float> &wC) {
    const int ni = d.I + 1, nj = d.J + 1;
    wC.resize(d.nC);
    size_t idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++, idx++) {
                float q[8];
                for (int n = 0; n < 8; n++)
                    q[n] = phi[(size_t)(i + (n & 1)) + (size_t)ni * ((size_t)(j + ((n >> 1) & 1)) + (size_t)nj * (k + (n >> 2)))];
                wC[idx] = clamp01(1.0f - negative_cell_volume(q));
            }
}

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
                wU[idx] = clamp01(1.0f - negative_area(P(i, j, k), P(i, j + 1, k), P(i, j, k + 1), P(i, j + 1, k + 1)));
    idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j <= d.J; j++)
            for (int i = 0; i < d.I; i++, idx++)
                wV[idx] = clamp01(1.0f - negative_area(P(i, j, k), P(i, j, k + 1), P(i + 1, j, k), P(i + 1, j, k + 1)));
    idx = 0;
    for (int k = 0; k <= d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++, idx++)
                wW[idx] = clamp01(1.0f - negative_area(P(i, j, k), P(i, j + 1, k), P(i + 1, j, k), P(i + 1, j + 1, k)));
}

// What one solid adds to the velocity data of the solid SDF (MeshLevelSet::_computeVelocityGridThread, meshlevelset.cpp:
// 1319-1372, summed by calculateUnion :1797-1828): on every face its solid fraction (getFaceWeightU/V/W of ITS OWN signed
// distance field, :357-387) as weight and fraction x velocity as value.  phi: the solid's nodal field on this grid;
// velocity: the velocity of a rigidly translating solid (null: at rest -- the domain walls and static obstacles add
// weight only).
void add_solid_fractions(const Dims &d, const std::vector<float> &phi, const float velocity[3], std::vector<float> weightSum[3],
                         std::vector<float> fieldSum[3]) {
    const int ni = d.I + 1, nj = d.J + 1;
    auto P = [&](int i, int j, int k) { return phi[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)]; };
    auto add = [&](int comp, size_t idx, float fraction) {
        if (!(fraction > 0.0f)) return;
        weightSum[comp][idx] += fraction;
        if (velocity) fieldSum[comp][idx] += fraction * velocity[comp];
    };
    size_t idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i <= d.I; i++, idx++) add(0, idx, negative_area(P(i, j, k), P(i, j + 1, k), P(i, j, k + 1), P(i, j + 1, k + 1)));
    idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j <= d.J; j++)
            for (int i = 0; i < d.I; i++, idx++) add(1, idx, negative_area(P(i, j, k), P(i, j, k + 1), P(i + 1, j, k), P(i + 1, j, k + 1)));
    idx = 0;
    for (int k = 0; k <= d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++, idx++) add(2, idx, negative_area(P(i, j, k), P(i, j + 1, k), P(i + 1, j, k), P(i + 1, j + 1, k)));
}

// Friction of the solids on the faces (_getFaceFrictionU/V/W, fluidsimulation.cpp:3785-3853): a quarter of the sum, over the
// four nodes of the face, of the friction of the mesh object the solid SDF names as closest at that node
// (MeshLevelSet::getClosestMeshObject, meshlevelset.cpp:140-150; none: 0).  The closest object of a node is settled by the
// order in which the solids are merged (_addStaticObjectsToSDF / _addAnimatedObjectsToSolidSDF, :2877-2975, over
// calculateUnion, meshlevelset.cpp:1782-1796): the domain first, wherever its boundary mesh lies within the exact band;
// a later solid takes a node over where its distance is below the merged one AND nearer to zero, if the node lies in that
// solid's band.  solids[0] is the domain (its field everywhere; in band where |phi| < (band + 1/2) dx), the others carry the
// largest float outside their band.
void build_face_friction(const Dims &d, int band, const std::vector<FrictionSolid> &solids, std::vector<float> &fU,
                         std::vector<float> &fV, std::vector<float> &fW) {
    const int ni = d.I + 1, nj = d.J + 1, nk = d.K + 1;
    const size_t nN = (size_t)ni * nj * nk;
    std::vector<float> merged(nN, std::numeric_limits<float>::max()), nodeFriction(nN, 0.0f);
    const float far = std::numeric_limits<float>::max();
    for (size_t s = 0; s < solids.size(); s++) {
        const std::vector<float> &phi = *solids[s].phi;
        const float domainBand = (float)((band + 0.5) * d.dx);
        for (size_t q = 0; q < nN; q++) {
            const float v = phi[q];
            if (!(v < merged[q])) continue;
            const bool inBand = s == 0 ? std::fabs(v) < domainBand : v != far;
            if (inBand && std::fabs(v) < std::fabs(merged[q])) nodeFriction[q] = solids[s].friction;
            merged[q] = v;
        }
    }
    auto F = [&](int i, int j, int k) { return nodeFriction[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)]; };
    fU.resize(d.nU); fV.resize(d.nV); fW.resize(d.nW);
    size_t idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i <= d.I; i++, idx++) fU[idx] = 0.25f * (((F(i, j, k) + F(i, j + 1, k)) + F(i, j, k + 1)) + F(i, j + 1, k + 1));
    idx = 0;
    for (int k = 0; k < d.K; k++)
        for (int j = 0; j <= d.J; j++)
            for (int i = 0; i < d.I; i++, idx++) fV[idx] = 0.25f * (((F(i, j, k) + F(i + 1, j, k)) + F(i, j, k + 1)) + F(i + 1, j, k + 1));
    idx = 0;
    for (int k = 0; k <= d.K; k++)
        for (int j = 0; j < d.J; j++)
            for (int i = 0; i < d.I; i++, idx++) fW[idx] = 0.25f * (((F(i, j, k) + F(i + 1, j, k)) + F(i, j + 1, k)) + F(i + 1, j + 1, k));
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

// ------------------------------------------------------------------------------------------------
// Nodal signed distance field of a closed triangle mesh on the simulation grid (what a MeshLevelSet holds after
// fastCalculateSignedDistanceField, meshlevelset.cpp:773-828): the exact point-to-triangle distance at the nodes of the
// mesh's index box grown by `band` cells (the reference computes exact distances in the band of every triangle's box,
// :572-601, and leaves an upper bound elsewhere), negative inside (parity of the crossings of a ray; the reference
// counts intersections along grid lines, :603-700 -- the same set for a closed mesh), `far` at all other nodes.
// Used by the façade for fluid objects, sources and obstacles that are not axis-aligned boxes.
namespace {
struct V3 { float x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// closest point on a triangle (Ericson, Real-Time Collision Detection 5.1.5), squared distance
float point_triangle_dist2(V3 p, V3 a, V3 b, V3 c) {
    const V3 ab = b - a, ac = c - a, ap = p - a;
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return dot3(ap, ap);
    const V3 bp = p - b;
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return dot3(bp, bp);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { const V3 q = p - (a + (d1 / (d1 - d3)) * ab); return dot3(q, q); }
    const V3 cp = p - c;
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return dot3(cp, cp);
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { const V3 q = p - (a + (d2 / (d2 - d6)) * ac); return dot3(q, q); }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const V3 q = p - (b + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * (c - b));
        return dot3(q, q);
    }
    const float denom = 1.0f / (va + vb + vc);
    const V3 q = p - (a + (vb * denom) * ab + (vc * denom) * ac);
    return dot3(q, q);
}
}  // namespace

// closest: (may be null) per node the triangle that gave the distance, -1 where none lies within the band -- what
// MeshLevelSet::_closestTriangles holds after the exact-band pass.
static int mesh_sdf_impl(int isize, int jsize, int ksize, double dx, const float *vertices_xyz, int num_vertices,
                         const int *triangles, int num_triangles, int band, float far_value, float *phi, int cell_lo[3],
                         int cell_hi[3], int *closest) {
    if (isize <= 0 || jsize <= 0 || ksize <= 0 || !(dx > 0.0) || !phi || band < 0) return FLIP_ERR_DOMAIN;
    for (int t = 0; t < 3 * num_triangles; t++)
        if (triangles[t] < 0 || triangles[t] >= num_vertices) return FLIP_ERR_OUT_OF_RANGE;
    const int ni = isize + 1, nj = jsize + 1, nk = ksize + 1;
    const float far = far_value > 0.0f ? far_value : (float)((band + 1) * dx);
    std::fill(phi, phi + (size_t)ni * nj * nk, far);
    if (closest) std::fill(closest, closest + (size_t)ni * nj * nk, -1);
    const V3 *vert = reinterpret_cast<const V3 *>(vertices_xyz);
    float blo[3] = {1e30f, 1e30f, 1e30f}, bhi[3] = {-1e30f, -1e30f, -1e30f};
    for (int v = 0; v < num_vertices; v++) {
        const float q[3] = {vert[v].x, vert[v].y, vert[v].z};
        for (int a = 0; a < 3; a++) { blo[a] = std::min(blo[a], q[a]); bhi[a] = std::max(bhi[a], q[a]); }
    }
    const int n[3] = {isize, jsize, ksize};
    int nlo[3] = {0, 0, 0}, nhi[3] = {-1, -1, -1};
    if (num_vertices > 0 && num_triangles > 0)
        for (int a = 0; a < 3; a++) {
            nlo[a] = std::max(0, (int)std::floor(blo[a] / dx) - band);
            nhi[a] = std::min(n[a], (int)std::ceil(bhi[a] / dx) + band);
        }
    for (int a = 0; a < 3; a++) {
        if (cell_lo) cell_lo[a] = nlo[a];
        if (cell_hi) cell_hi[a] = std::max(nhi[a], nlo[a]);
    }
    if (nhi[0] < nlo[0] || nhi[1] < nlo[1] || nhi[2] < nlo[2]) return FLIP_OK;
    // the region of interest: the mesh's index box grown by the band
    const int ri = nhi[0] - nlo[0] + 1, rj = nhi[1] - nlo[1] + 1, rk = nhi[2] - nlo[2] + 1;
    std::vector<float> best((size_t)ri * rj * rk, 1e30f);          // squared distance to the nearest triangle seen
    std::vector<int> bestTri;
    if (closest) bestTri.assign((size_t)ri * rj * rk, -1);
    std::vector<std::vector<float>> hits((size_t)rj * rk);          // x of the crossings of the +x ray of every (j, k) line
    for (int t = 0; t < num_triangles; t++) {
        const V3 &a = vert[triangles[3 * t]], &b = vert[triangles[3 * t + 1]], &c = vert[triangles[3 * t + 2]];
        const float tlo[3] = {std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z))};
        const float thi[3] = {std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z))};
        int q0[3], q1[3];
        for (int ax = 0; ax < 3; ax++) {
            q0[ax] = std::max(nlo[ax], (int)std::floor(tlo[ax] / dx) - band);
            q1[ax] = std::min(nhi[ax], (int)std::ceil(thi[ax] / dx) + band);
        }
        // exact distances in the band of the triangle's index box (meshlevelset.cpp:572-601)
        for (int k = q0[2]; k <= q1[2]; k++)
            for (int j = q0[1]; j <= q1[1]; j++)
                for (int i = q0[0]; i <= q1[0]; i++) {
                    const V3 p = {(float)(i * dx), (float)(j * dx), (float)(k * dx)};
                    const size_t q = (size_t)(i - nlo[0]) + (size_t)ri * ((j - nlo[1]) + (size_t)rj * (k - nlo[2]));
                    const float cand = point_triangle_dist2(p, a, b, c);
                    if (cand < best[q]) {           // the first of equally near triangles stays (meshlevelset.cpp:596)
                        best[q] = cand;
                        if (closest) bestTri[q] = t;
                    }
                }
        // crossings of the +x rays through the (j, k) lines the triangle's yz projection can cover; every ray starts from a
        // point nudged off the lattice (mesh vertices on grid lines)
        const int j0 = std::max(nlo[1], (int)std::floor(tlo[1] / dx) - 1), j1 = std::min(nhi[1], (int)std::ceil(thi[1] / dx));
        const int k0 = std::max(nlo[2], (int)std::floor(tlo[2] / dx) - 1), k1 = std::min(nhi[2], (int)std::ceil(thi[2] / dx));
        for (int k = k0; k <= k1; k++)
            for (int j = j0; j <= j1; j++) {
                const double ry = (double)(float)(j * dx) + 1.2345e-4 * dx, rz = (double)(float)(k * dx) + 2.3456e-4 * dx;
                const double ay = a.y - ry, az = a.z - rz, by = b.y - ry, bz = b.z - rz, cy = c.y - ry, cz = c.z - rz;
                const double w0 = by * cz - bz * cy, w1 = cy * az - cz * ay, w2 = ay * bz - az * by;
                if ((w0 > 0 && w1 > 0 && w2 > 0) || (w0 < 0 && w1 < 0 && w2 < 0)) {
                    const double s = w0 + w1 + w2;
                    hits[(size_t)(j - nlo[1]) + (size_t)rj * (k - nlo[2])].push_back((float)((w0 * a.x + w1 * b.x + w2 * c.x) / s));
                }
            }
    }
    for (int k = nlo[2]; k <= nhi[2]; k++)
        for (int j = nlo[1]; j <= nhi[1]; j++) {
            const std::vector<float> &xs = hits[(size_t)(j - nlo[1]) + (size_t)rj * (k - nlo[2])];
            for (int i = nlo[0]; i <= nhi[0]; i++) {
                const float px = (float)(i * dx);
                int crossings = 0;
                for (float x : xs) crossings += x > px ? 1 : 0;
                const float d2 = best[(size_t)(i - nlo[0]) + (size_t)ri * ((j - nlo[1]) + (size_t)rj * (k - nlo[2]))];
                const float d = d2 < 1e29f ? std::min(std::sqrt(d2), far) : far;       // no triangle within the band: the bound
                phi[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] = (crossings & 1) ? -d : d;
                if (closest) closest[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] = bestTri[(size_t)(i - nlo[0]) + (size_t)ri * ((j - nlo[1]) + (size_t)rj * (k - nlo[2]))];
            }
        }
    return FLIP_OK;
}

extern "C" int flip_mesh_sdf(int isize, int jsize, int ksize, double dx, const float *vertices_xyz, int num_vertices,
                             const int *triangles, int num_triangles, int band, float far_value, float *phi, int cell_lo[3],
                             int cell_hi[3]) {
    return mesh_sdf_impl(isize, jsize, ksize, dx, vertices_xyz, num_vertices, triangles, num_triangles, band, far_value, phi,
                         cell_lo, cell_hi, nullptr);
}

// ------------------------------------------------------------------------------------------------
// The velocity data of a moving mesh (MeshLevelSet::_computeVelocityGridThread, meshlevelset.cpp:1319-1372): on every face
// whose solid fraction (of the mesh's own signed distance field) is positive, fraction x the velocity of the mesh surface
// nearest to the face centre.  Nearest surface point as the reference finds it (getNearestVelocity :168-205): among the
// closest triangles of the eight nodes of the cell that holds the face centre, the one nearest to the centre; its
// velocity there (_pointToTriangleVelocity :1567-1640): the barycentric blend of the vertex velocities where the
// projection falls inside the triangle, else the blend along the nearer of the two candidate edges
// (_pointToSegmentVelocity :1696-1714).
namespace {
inline float length3(V3 a) { return std::sqrt(dot3(a, a)); }
// velocity at the point of the segment pa-pb nearest to p, blended from the end velocities; *distance: how far that is
V3 segment_velocity(V3 p, V3 pa, V3 pb, V3 va, V3 vb, float *distance) {
    const V3 along = pb - pa;
    const double len2 = dot3(along, along);
    float towardsA = (float)(dot3(pb - p, along) / len2);      // 1 at pa, 0 at pb
    towardsA = towardsA < 0 ? 0 : (towardsA > 1 ? 1 : towardsA);
    *distance = length3(p - (towardsA * pa + (1 - towardsA) * pb));
    return towardsA * va + (1 - towardsA) * vb;
}
// corner[3], cornerVel[3]: a triangle and the velocities of its corners
V3 triangle_velocity(V3 p, const V3 corner[3], const V3 cornerVel[3]) {
    const float tiny = 1e-6f;
    bool atRest = true;
    for (int c = 0; c < 3; c++)
        atRest = atRest && std::fabs(cornerVel[c].x) < tiny && std::fabs(cornerVel[c].y) < tiny && std::fabs(cornerVel[c].z) < tiny;
    if (atRest) return {0, 0, 0};
    // barycentric weights of the projection of p onto the triangle's plane, with corner 2 as the origin
    const V3 e0 = corner[0] - corner[2], e1 = corner[1] - corner[2], rel = p - corner[2];
    const float g00 = dot3(e0, e0), g11 = dot3(e1, e1), g01 = dot3(e0, e1);
    const float scale = 1.0f / std::fmax(g00 * g11 - g01 * g01, 1e-30f);
    const float r0 = dot3(e0, rel), r1 = dot3(e1, rel);
    float bary[3];
    bary[0] = scale * (g11 * r0 - g01 * r1);
    bary[1] = scale * (g00 * r1 - g01 * r0);
    bary[2] = 1 - bary[0] - bary[1];
    if (bary[0] >= 0 && bary[1] >= 0 && bary[2] >= 0) return bary[0] * cornerVel[0] + bary[1] * cornerVel[1] + bary[2] * cornerVel[2];
    // outside: the nearer of the two edges at the first corner with a positive weight (its opposite edge is ruled out);
    // with neither of the first two positive, the two edges at corner 2
    static const int edges[3][2][2] = {{{0, 1}, {0, 2}}, {{0, 1}, {1, 2}}, {{0, 2}, {1, 2}}};
    const int which = bary[0] > 0 ? 0 : (bary[1] > 0 ? 1 : 2);
    float dist[2];
    V3 cand[2];
    for (int e = 0; e < 2; e++) {
        const int a = edges[which][e][0], b = edges[which][e][1];
        cand[e] = segment_velocity(p, corner[a], corner[b], cornerVel[a], cornerVel[b], &dist[e]);
    }
    return dist[0] < dist[1] ? cand[0] : cand[1];
}
}  // namespace

namespace flip {
// phi: the mesh's nodal field (out); fraction[3] / field[3]: per face of U, V, W the solid fraction and fraction x
// velocity component (out, resized).
int mesh_velocity_data(const Dims &d, const float *vertices_xyz, int num_vertices, const int *triangles, int num_triangles,
                       const float *vertex_velocities_xyz, int band, float far_value, std::vector<float> &phi,
                       std::vector<float> fraction[3], std::vector<float> field[3]) {
    const int ni = d.I + 1, nj = d.J + 1, nk = d.K + 1;
    phi.resize((size_t)ni * nj * nk);
    std::vector<int> closest((size_t)ni * nj * nk);
    int clo[3], chi[3];
    const int rc = mesh_sdf_impl(d.I, d.J, d.K, d.dx, vertices_xyz, num_vertices, triangles, num_triangles, band, far_value,
                                 phi.data(), clo, chi, closest.data());
    if (rc != FLIP_OK) return rc;
    const V3 *vert = reinterpret_cast<const V3 *>(vertices_xyz), *vel = reinterpret_cast<const V3 *>(vertex_velocities_xyz);
    const int n[3] = {d.nU, d.nV, d.nW};
    for (int m = 0; m < 3; m++) { fraction[m].assign(n[m], 0.0f); field[m].assign(n[m], 0.0f); }
    auto P = [&](int i, int j, int k) { return phi[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * k)]; };
    const double invdx = 1.0 / d.dx;
    auto nearest_velocity = [&](V3 p) -> V3 {
        const int gi = (int)std::floor(p.x * invdx), gj = (int)std::floor(p.y * invdx), gk = (int)std::floor(p.z * invdx);
        int tri = -1;
        float nearest = 1e30f;
        static const int order[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}};
        for (auto &o : order) {
            const int a = gi + o[0], b = gj + o[1], c = gk + o[2];
            if (a < 0 || b < 0 || c < 0 || a >= ni || b >= nj || c >= nk) continue;
            const int t = closest[(size_t)a + (size_t)ni * ((size_t)b + (size_t)nj * c)];
            if (t < 0) continue;
            const float dist = std::sqrt(point_triangle_dist2(p, vert[triangles[3 * t]], vert[triangles[3 * t + 1]], vert[triangles[3 * t + 2]]));
            if (dist < nearest) { nearest = dist; tri = t; }
        }
        if (tri < 0) return {0, 0, 0};
        const int *q = triangles + 3 * tri;
        const V3 corner[3] = {vert[q[0]], vert[q[1]], vert[q[2]]}, cornerVel[3] = {vel[q[0]], vel[q[1]], vel[q[2]]};
        return triangle_velocity(p, corner, cornerVel);
    };
    // only the faces of the mesh's index box grown by the band can have a positive fraction
    const int i0 = clo[0], i1 = std::min(chi[0], d.I), j0 = clo[1], j1 = std::min(chi[1], d.J), k0 = clo[2], k1 = std::min(chi[2], d.K);
    for (int k = k0; k <= k1; k++)
        for (int j = j0; j <= j1; j++)
            for (int i = i0; i <= i1; i++) {
                if (j < d.J && k < d.K) {
                    const float f = negative_area(P(i, j, k), P(i, j + 1, k), P(i, j, k + 1), P(i, j + 1, k + 1));
                    if (f > 0.0f) {
                        const size_t idx = (size_t)i + (size_t)(d.I + 1) * ((size_t)j + (size_t)d.J * k);
                        const V3 v = nearest_velocity({(float)(i * d.dx), (float)((j + 0.5) * d.dx), (float)((k + 0.5) * d.dx)});
                        fraction[0][idx] = f; field[0][idx] = f * v.x;
                    }
                }
                if (i < d.I && k < d.K) {
                    const float f = negative_area(P(i, j, k), P(i, j, k + 1), P(i + 1, j, k), P(i + 1, j, k + 1));
                    if (f > 0.0f) {
                        const size_t idx = (size_t)i + (size_t)d.I * ((size_t)j + (size_t)(d.J + 1) * k);
                        const V3 v = nearest_velocity({(float)((i + 0.5) * d.dx), (float)(j * d.dx), (float)((k + 0.5) * d.dx)});
                        fraction[1][idx] = f; field[1][idx] = f * v.y;
                    }
                }
                if (i < d.I && j < d.J) {
                    const float f = negative_area(P(i, j, k), P(i, j + 1, k), P(i + 1, j, k), P(i + 1, j + 1, k));
                    if (f > 0.0f) {
                        const size_t idx = (size_t)i + (size_t)d.I * ((size_t)j + (size_t)d.J * k);
                        const V3 v = nearest_velocity({(float)((i + 0.5) * d.dx), (float)((j + 0.5) * d.dx), (float)(k * d.dx)});
                        fraction[2][idx] = f; field[2][idx] = f * v.z;
                    }
                }
            }
    return FLIP_OK;
}
}  // namespace flip

extern "C" int flip_mesh_velocity_data(int isize, int jsize, int ksize, double dx, const float *vertices_xyz, int num_vertices,
                                       const int *triangles, int num_triangles, const float *vertex_velocities_xyz, int band,
                                       float far_value, float *phi, float *fractionU, float *fractionV, float *fractionW,
                                       float *fieldU, float *fieldV, float *fieldW) {
    if (isize <= 0 || jsize <= 0 || ksize <= 0 || !(dx > 0.0) || !phi || !vertex_velocities_xyz) return FLIP_ERR_DOMAIN;
    try {
        flip::Dims d;
        d.I = isize; d.J = jsize; d.K = ksize; d.dx = dx; d.kOff = 0; d.Kg = ksize; d.kOwn0 = 0; d.kOwn1 = ksize;
        d.nU = (isize + 1) * jsize * ksize; d.nV = isize * (jsize + 1) * ksize; d.nW = isize * jsize * (ksize + 1);
        d.nC = isize * jsize * ksize; d.nN = (isize + 1) * (jsize + 1) * (ksize + 1);
        std::vector<float> p, fr[3], fl[3];
        const int rc = flip::mesh_velocity_data(d, vertices_xyz, num_vertices, triangles, num_triangles, vertex_velocities_xyz, band,
                                                far_value, p, fr, fl);
        if (rc != FLIP_OK) return rc;
        memcpy(phi, p.data(), sizeof(float) * p.size());
        float *outs[6] = {fractionU, fractionV, fractionW, fieldU, fieldV, fieldW};
        for (int m = 0; m < 3; m++) {
            if (outs[m]) memcpy(outs[m], fr[m].data(), sizeof(float) * fr[m].size());
            if (outs[3 + m]) memcpy(outs[3 + m], fl[m].data(), sizeof(float) * fl[m].size());
        }
        return FLIP_OK;
    } catch (const std::bad_alloc &) { return FLIP_ERR_RUNTIME; }
}

// the generated marching-cubes case table of the surface reconstruction (mc_tables.h), for inspection and tests
extern "C" int flip_mc_case_table(unsigned char counts[256], unsigned char edge_triples[256 * 24]) {
    unsigned char tris[256][flip::MC_MAX_TRIS * 3];
    unsigned char cnt[256];
    flip::build_mc_tables(cnt, tris);
    if (counts) memcpy(counts, cnt, 256);
    if (edge_triples) memcpy(edge_triples, tris, sizeof(tris));
    return FLIP_OK;
}
