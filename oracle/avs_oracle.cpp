// =============================================================================
// avs_oracle.cpp -- CPU ORACLE for the adaptive-octree viscosity solve.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (adaptiveviscositysolver_b200/,
// include/, the C-ABI library) may include, link or call this file.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / the timed CPU baseline.
//
// It restates, stage by stage, what rgoldade/AdaptiveViscositySolver does
// between field validation and the CG solve
// (Source/HDK_AdaptiveViscosity.cpp:233-653, "AV.cpp" below;
//  Source/HDK_OctreeGrid.cpp "OG.cpp"; Source/HDK_Utilities.h "UTIL.h").
//
// PARITY STATUS: pinned to the reference's OWN CODE, with Houdini's and Eigen's behaviour
// assumed.  The reference's three .cpp files compile UNCHANGED against the stand-ins of
// oracle/mock_hdk into oracle/_ref/libavs_ref.so (oracle/Makefile target `ref`), and
// tests/test_reference_pin.py holds this restatement to that library bit for bit on
// labels, weights, DOF numbering, matrix, rhs and restricted velocity, exactly on
// iteration counts, to 1e-9 on the solution and the regular-grid output.  What is NOT
// pinned is what the stand-ins assume about closed / absent code (top of mock_hdk.h):
// the reference ships no tests, golden vectors or fixtures, and Houdini's
// SIM_RawField / UT_VoxelArray / computeSDFWeightsSampled and Eigen itself are not here.
// Further pins: the reference's debug invariants restated as property tests, analytic
// known answers (tests/test_oracle_*.py) and a SciPy cross-check of the CG loop.
//
// Conventions the reference delegates to Houdini and that are FIXED HERE
// (SURVEY.md section 8c / Appendix D):
//   * flat arrays are x-fastest: idx = x + nx*(y + ny*z);
//   * a sampled field stores, per component, the world position of its
//     sample (0,0,0) ("org") and its spacing; getValue(pos) is trilinear with
//     clamp-to-edge, evaluated in fp64 as lerp(a,b,t)=a+t*(b-a), x then y then z,
//     without FMA contraction (compile with -ffp-contract=off);
//   * raw voxel reads with out-of-range indices clamp (UTIL.h:219-227);
//   * computeSDFWeightsSampled(sdf,n,...): fraction of the n^3 sub-samples at
//     offsets ((k+1/2)/n - 1/2)*dx around the sample whose interpolated sdf
//     (minus dilate) is < 0;
//   * voxel sizes are float32-rounded like UT_Vector3 (AV.cpp:242, 1733),
//     gradientDx accumulates in float32 (AV.cpp:1738);
//   * 16^3 voxel tiles (UT_VoxelArray) -- they matter because the reference
//     only classifies faces/edges inside "occupied" tiles (AV.cpp:886-1085).
// =============================================================================
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef int64_t exint;

// OG.h:33-39
enum CellLabel : uint8_t { INACTIVE = 0, ACTIVE = 1, UP = 2, DOWN = 3 };
// UTIL.h:18-21
constexpr exint FLUID = 0, UNASSIGNED = -1, SOLIDBOUNDARY = -2, OUTSIDE = -3;
constexpr int TILE = 16;  // UT_VoxelArray tile edge

struct I3 {
    int v[3];
    int &operator[](int a) { return v[a]; }
    int operator[](int a) const { return v[a]; }
};
static inline I3 mk(int x, int y, int z) { return I3{{x, y, z}}; }

template <class T>
struct Arr3 {
    I3 n{{0, 0, 0}};
    std::vector<T> d;
    void init(I3 r, T v) {
        n = r;
        d.assign((size_t)r[0] * r[1] * r[2], v);
    }
    size_t size() const { return d.size(); }
    size_t lin(int x, int y, int z) const { return (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * z); }
    T &at(const I3 &c) { return d[lin(c[0], c[1], c[2])]; }
    const T &at(const I3 &c) const { return d[lin(c[0], c[1], c[2])]; }
    // HDKgetFieldValue: out-of-range indices clamp (UTIL.h:219-227, Appendix D)
    T get(const I3 &c) const {
        int x = std::min(std::max(c[0], 0), n[0] - 1);
        int y = std::min(std::max(c[1], 0), n[1] - 1);
        int z = std::min(std::max(c[2], 0), n[2] - 1);
        return d[lin(x, y, z)];
    }
    bool inside(const I3 &c) const {
        return c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[0] < n[0] && c[1] < n[1] && c[2] < n[2];
    }
};

// A sampled scalar field as the caller hands it over (one component).
struct Field {
    const float *data = nullptr;  // null => constant
    int res[3] = {1, 1, 1};
    double org[3] = {0, 0, 0};    // world position of sample (0,0,0)
    double dx = 1;
    float constant = 0;
    bool is_constant() const { return data == nullptr; }
    float raw(int x, int y, int z) const {
        if (!data) return constant;
        x = std::min(std::max(x, 0), res[0] - 1);
        y = std::min(std::max(y, 0), res[1] - 1);
        z = std::min(std::max(z, 0), res[2] - 1);
        return data[(size_t)x + (size_t)res[0] * ((size_t)y + (size_t)res[1] * z)];
    }
    // SIM_RawField::getValue(pos): trilinear, clamp-to-edge (Appendix D)
    double value(const double p[3]) const {
        if (!data) return (double)constant;
        int i0[3], i1[3];
        double t[3];
        for (int a = 0; a < 3; ++a) {
            double g = (p[a] - org[a]) / dx;
            double hi = (double)(res[a] - 1);
            if (g < 0.0) g = 0.0;
            if (g > hi) g = hi;
            double f = std::floor(g);
            i0[a] = (int)f;
            i1[a] = std::min(i0[a] + 1, res[a] - 1);
            t[a] = g - f;
        }
        auto V = [&](int x, int y, int z) { return (double)data[(size_t)x + (size_t)res[0] * ((size_t)y + (size_t)res[1] * z)]; };
        auto lerp = [](double a, double b, double tt) { return a + tt * (b - a); };
        double c00 = lerp(V(i0[0], i0[1], i0[2]), V(i1[0], i0[1], i0[2]), t[0]);
        double c10 = lerp(V(i0[0], i1[1], i0[2]), V(i1[0], i1[1], i0[2]), t[0]);
        double c01 = lerp(V(i0[0], i0[1], i1[2]), V(i1[0], i0[1], i1[2]), t[0]);
        double c11 = lerp(V(i0[0], i1[1], i1[2]), V(i1[0], i1[1], i1[2]), t[0]);
        double c0 = lerp(c00, c10, t[1]);
        double c1 = lerp(c01, c11, t[1]);
        return lerp(c0, c1, t[2]);
    }
};

// ---- index algebra (UTIL.h:46-217, OG.h:53-142) ----------------------------
static inline I3 cellToFace(I3 c, int axis, int dir) { if (dir == 1) ++c[axis]; return c; }               // UTIL.h:46
static inline I3 cellToCell(I3 c, int axis, int dir) { if (dir == 0) --c[axis]; else ++c[axis]; return c; } // UTIL.h:56
static inline I3 cellToEdge(I3 c, int edgeAxis, int edgeIndex) {                                          // UTIL.h:70
    for (int o = 0; o < 2; ++o) if (edgeIndex & (1 << o)) ++c[(edgeAxis + 1 + o) % 3];
    return c;
}
static inline I3 faceToCell(I3 f, int axis, int dir) { if (dir == 0) --f[axis]; return f; }               // UTIL.h:102
static inline I3 faceToEdge(I3 f, int faceAxis, int edgeAxis, int dir) {                                  // UTIL.h:115
    if (dir == 1) ++f[3 - faceAxis - edgeAxis];
    return f;
}
static inline I3 edgeToFace(I3 e, int edgeAxis, int faceAxis, int dir) {                                  // UTIL.h:151
    if (dir == 0) --e[3 - faceAxis - edgeAxis];
    return e;
}
static inline I3 edgeToCell(I3 e, int edgeAxis, int cellIndex) {                                          // UTIL.h:169
    for (int o = 0; o < 2; ++o) if (!(cellIndex & (1 << o))) --e[(edgeAxis + 1 + o) % 3];
    return e;
}
static inline I3 parentOf(I3 c) { return mk(c[0] / 2, c[1] / 2, c[2] / 2); }                              // OG.h:53
static inline I3 childCell(I3 c, int child) {                                                             // OG.h:71
    I3 r = mk(c[0] * 2, c[1] * 2, c[2] * 2);
    for (int a = 0; a < 3; ++a) if (child & (1 << a)) ++r[a];
    return r;
}
static inline I3 childFace(I3 f, int axis, int child) {                                                   // OG.h:94
    I3 r = mk(f[0] * 2, f[1] * 2, f[2] * 2);
    if (child & 1) ++r[(axis + 1) % 3];
    if (child & 2) ++r[(axis + 2) % 3];
    return r;
}
static inline I3 childEdge(I3 e, int edgeAxis, int child) {                                               // OG.h:108
    I3 r = mk(e[0] * 2, e[1] * 2, e[2] * 2);
    if (child > 0) ++r[edgeAxis];
    return r;
}
static inline I3 childEdgeInFace(I3 f, int faceAxis, int edgeAxis, int child) {                           // OG.h:126
    I3 r = mk(f[0] * 2, f[1] * 2, f[2] * 2);
    if (child == 1) ++r[edgeAxis];
    ++r[3 - faceAxis - edgeAxis];
    return r;
}

static inline I3 nodeToFace(I3 n, int faceAxis, int faceIndex) {                                          // UTIL.h:187
    for (int o = 0; o < 2; ++o) if (!(faceIndex & (1 << o))) --n[(faceAxis + 1 + o) % 3];
    return n;
}
static inline I3 faceToNode(I3 f, int faceAxis, int nodeIndex) {                                          // UTIL.h:133
    for (int o = 0; o < 2; ++o) if (nodeIndex & (1 << o)) ++f[(faceAxis + 1 + o) % 3];
    return f;
}
static inline I3 cellToNode(I3 c, int nodeIndex) {                                                        // UTIL.h:88
    for (int a = 0; a < 3; ++a) if (nodeIndex & (1 << a)) ++c[a];
    return c;
}

struct Params {
    double dt = 1.0 / 24;
    double tolerance = 1e-3;       // AV.cpp:62-63
    int maxIterations = 2500;      // AV.cpp:65-66
    int numberSuperSamples = 3;    // AV.cpp:104
    int octreeLevels = 4;          // AV.cpp:106
    int fineBandwidth = 0;         // AV.h:33 reads an option no parm defines => 0
    int useEnhancedGradients = 1;  // AV.cpp:109
    int doApplySolidWeights = 0;   // AV.h:37 (name mismatch => false)
    double extrapolation = 0.5;    // AV.cpp:69 (in voxels; scaled by dx at AV.cpp:243)
    int singlePrecision = 0;       // USESINGLEPRECISION (UTIL.h:25-37)
};

struct Scene {
    int res[3];
    double origin[3];
    double dx;  // fine voxel size (cubic)
    Field surface;         // centre sampled, res
    Field vel[3];          // face sampled
    Field faceWeights[3];  // face sampled ("surfaceweights")
    Field viscosity;       // centre sampled or constant
    Field density;         // centre sampled or constant
    Field collision;       // any grid or constant (positive inside the solid)
    Field collisionVel[3]; // any grid or constant
};

struct Oracle {
    Scene S;
    Params P;
    std::vector<std::vector<float>> owned;  // copies of caller data

    double dx0 = 1;                 // float32-rounded voxel size (AV.cpp:242)
    I3 N{{0, 0, 0}};                // liquid surface resolution
    I3 Pad{{0, 0, 0}};              // power-of-two padded resolution (OG.cpp:18-24)
    int levels = 0;                 // built levels after capping (OG.cpp:198-211)
    int levelsAllocated = 0;

    bool useShortcut = true;        // exact early-out in sdfWeights (tests switch it off to compare)
    Arr3<float> centerW;            // AV.cpp:761
    Arr3<float> edgeW[3];           // AV.cpp:763-765
    Arr3<float> mask;               // AV.cpp:806-871
    std::vector<Arr3<uint8_t>> labels;             // OG.h:325
    Arr3<exint> regIdx[3];                         // AV.cpp:303
    exint regularDOFs = 0;
    std::vector<std::vector<Arr3<exint>>> faceIdx; // [level][axis] AV.cpp:337
    std::vector<std::vector<Arr3<exint>>> edgeIdx; // [level][axis] AV.cpp:340
    std::vector<Arr3<exint>> centerIdx;            // [level]       AV.cpp:343
    exint nFace = 0, nEdge = 0, nCenter = 0;

    std::vector<int32_t> faceKey;   // per face DOF: level, axis, i, j, k
    std::vector<double> x0;         // restricted u^n == initial guess (AV.cpp:507-529)
    std::vector<double> rhs;
    std::vector<exint> rowPtr;
    std::vector<int32_t> colIdx;
    std::vector<double> val;
    std::vector<double> sol;
    // stage 11 (AV.cpp:661-707, HDK_OctreeVectorFieldInterpolator.*)
    std::vector<std::vector<Arr3<float>>> octVel;    // [level][axis] setOctreeVelocity (fp32 fields)
    std::vector<Arr3<uint8_t>> nodeLabel;            // INACTIVENODE / ACTIVENODE / DEPENDENTNODE
    std::vector<std::vector<Arr3<float>>> nodeVal, nodeW;
    std::vector<Arr3<exint>> nodeFlag;
    Arr3<float> outVel[3];                           // regular-grid velocity after the solve
    exint interpolatedFaces = 0;
    int iterations = 0;
    double error = 0;
    int stage = 0;

    double levelDx(int level) const { return (double)(float)(dx0 * (double)(1 << level)); }
    I3 cellRes(int level) const { return mk(Pad[0] >> level, Pad[1] >> level, Pad[2] >> level); }

    // indexToPos for the sample types (Appendix D)
    void centerPos(const I3 &c, int level, double p[3]) const {
        double h = levelDx(level);
        for (int a = 0; a < 3; ++a) p[a] = S.origin[a] + (c[a] + 0.5) * h;
    }
    void facePos(const I3 &f, int axis, int level, double p[3]) const {
        double h = levelDx(level);
        for (int a = 0; a < 3; ++a) p[a] = S.origin[a] + (f[a] + (a == axis ? 0.0 : 0.5)) * h;
    }
    void edgePos(const I3 &e, int axis, int level, double p[3]) const {
        double h = levelDx(level);
        for (int a = 0; a < 3; ++a) p[a] = S.origin[a] + (e[a] + (a == axis ? 0.5 : 0.0)) * h;
    }

    // ---------------------------------------------------------------- stage 1
    // computeIntegrationWeights (AV.cpp:712-726) for one sample type.
    // off[a] = 0.5 where the sample is cell-centred on axis a, 0 where node-aligned.
    void sdfWeights(Arr3<float> &w, const Field &sdf, const double off[3], I3 res, int n, double dilate) {
        w.init(res, 0.f);
        const double inv = 1.0 / (double)n;
        const double total = (double)n * n * n;
#pragma omp parallel for schedule(dynamic, 1)
        for (int z = 0; z < res[2]; ++z)
            for (int y = 0; y < res[1]; ++y)
                for (int x = 0; x < res[0]; ++x) {
                    double c[3] = {S.origin[0] + (x + off[0]) * dx0, S.origin[1] + (y + off[1]) * dx0,
                                   S.origin[2] + (z + off[2]) * dx0};
                    // Exact shortcut (not in the definition, only faster): the interpolant is a convex
                    // combination of the voxels under the sample's box, so if they all have one sign the
                    // count is n^3 or 0.  Checked against brute force in tests/test_oracle_system.py.
                    if (sdf.data && useShortcut) {
                        const double h = (0.5 - 0.5 * inv) * dx0;
                        int lo[3], hi[3];
                        for (int a = 0; a < 3; ++a) {
                            double gl = (c[a] - h - sdf.org[a]) / sdf.dx - 1e-9, gh = (c[a] + h - sdf.org[a]) / sdf.dx + 1e-9;
                            double top = (double)(sdf.res[a] - 1);
                            gl = std::min(std::max(gl, 0.0), top);
                            gh = std::min(std::max(gh, 0.0), top);
                            lo[a] = (int)std::floor(gl);
                            hi[a] = std::min((int)std::floor(gh) + 1, sdf.res[a] - 1);
                        }
                        bool allNeg = true, allPos = true;
                        for (int kz = lo[2]; kz <= hi[2]; ++kz)
                            for (int ky = lo[1]; ky <= hi[1]; ++ky)
                                for (int kx = lo[0]; kx <= hi[0]; ++kx) {
                                    double v = (double)sdf.raw(kx, ky, kz) - dilate;
                                    allNeg = allNeg && (v < 0.0);
                                    allPos = allPos && (v >= 0.0);
                                }
                        if (allNeg) { w.d[w.lin(x, y, z)] = 1.f; continue; }
                        if (allPos) { w.d[w.lin(x, y, z)] = 0.f; continue; }
                    }
                    int count = 0;
                    for (int sz = 0; sz < n; ++sz)
                        for (int sy = 0; sy < n; ++sy)
                            for (int sx = 0; sx < n; ++sx) {
                                double p[3] = {c[0] + ((sx + 0.5) * inv - 0.5) * dx0,
                                               c[1] + ((sy + 0.5) * inv - 0.5) * dx0,
                                               c[2] + ((sz + 0.5) * inv - 0.5) * dx0};
                                if (sdf.value(p) - dilate < 0.0) ++count;
                            }
                    w.d[w.lin(x, y, z)] = (float)((double)count / total);
                }
    }

    // buildIntegrationWeights (AV.cpp:748-791)
    void buildIntegrationWeights() {
        const int n = P.numberSuperSamples;
        const double offC[3] = {0.5, 0.5, 0.5};
        sdfWeights(centerW, S.surface, offC, N, n, 0.0);
        for (int a = 0; a < 3; ++a) {
            double off[3] = {0, 0, 0};
            off[a] = 0.5;  // HDK_XEDGE = SIM_SAMPLE_EDGEYZ: centred along its own axis only (UTIL.h:13-15)
            I3 r = mk(N[0] + 1, N[1] + 1, N[2] + 1);
            r[a] -= 1;
            sdfWeights(edgeW[a], S.surface, off, r, n, 0.0);
        }
        if (P.doApplySolidWeights) {
            // AV.cpp:772-790: liquid weight /= solid weight where the latter is > 0
            // (setScaleDivideThreshold(1, nullptr, &b, 0)); solid weights use dilate = -extrapolation.
            const double extrap = dx0 * P.extrapolation;
            Arr3<float> sw;
            sdfWeights(sw, S.collision, offC, N, n, -extrap);
            for (size_t i = 0; i < centerW.size(); ++i)
                if (sw.d[i] > 0.f) centerW.d[i] = centerW.d[i] / sw.d[i];
            for (int a = 0; a < 3; ++a) {
                double off[3] = {0, 0, 0};
                off[a] = 0.5;
                sdfWeights(sw, S.collision, off, edgeW[a].n, n, -extrap);
                for (size_t i = 0; i < edgeW[a].size(); ++i)
                    if (sw.d[i] > 0.f) edgeW[a].d[i] = edgeW[a].d[i] / sw.d[i];
            }
        }
    }

    // ---------------------------------------------------------------- stage 2
    // buildOctree mask lambda (AV.cpp:793-871)
    void buildMask() {
        const double fineVoxelWidth = std::max(2.0, (double)P.fineBandwidth);  // AV.cpp:259
        const double inner = dx0 * fineVoxelWidth;                             // AV.cpp:261
        const double outer = 3.0 * dx0;                                        // AV.cpp:262
        const double extrap = dx0 * P.extrapolation;                           // AV.cpp:243
        mask.init(N, 1.f);
#pragma omp parallel for schedule(static)
        for (int z = 0; z < N[2]; ++z)
            for (int y = 0; y < N[1]; ++y)
                for (int x = 0; x < N[0]; ++x) {
                    double sdf = (double)S.surface.raw(x, y, z);
                    float m;
                    if (sdf > 0 && sdf < outer) m = 0;
                    else if (sdf <= 0.) {
                        if (sdf > -inner) m = 0;
                        else {
                            double p[3];
                            centerPos(mk(x, y, z), 0, p);
                            m = (S.collision.value(p) > (-inner - extrap)) ? 0.f : -1.f;
                        }
                    } else m = 1;
                    mask.d[mask.lin(x, y, z)] = m;
                }
    }

    // ---------------------------------------------------------------- stage 3
    // HDK_OctreeGrid::init (OG.cpp:4-243)
    void buildOctree() {
        for (int a = 0; a < 3; ++a) {
            double l = std::ceil(std::log2((double)N[a]));  // OG.cpp:18-24
            Pad[a] = (int)std::exp2(l);
        }
        int L = P.octreeLevels;  // OG.cpp:32-40
        if (Pad[0] > 0 && Pad[1] > 0 && Pad[2] > 0)
            for (int a = 0; a < 3; ++a)
                if (std::log2((double)Pad[a]) < L) L = (int)std::log2((double)Pad[a]);
        levelsAllocated = L;
        labels.assign(L, Arr3<uint8_t>());
        for (int l = 0; l < L; ++l) labels[l].init(cellRes(l), INACTIVE);

        // setBaseGridLabels (OG.cpp:310-392)
        for (int z = 0; z < N[2]; ++z)
            for (int y = 0; y < N[1]; ++y)
                for (int x = 0; x < N[0]; ++x) {
                    float m = mask.d[mask.lin(x, y, z)];
                    if (m == 0) labels[0].d[labels[0].lin(x, y, z)] = ACTIVE;
                    else if (m < 0) labels[0].d[labels[0].lin(x, y, z)] = UP;
                }

        for (int level = 0; level < L - 1; ++level) {
            Arr3<uint8_t> &cur = labels[level];
            Arr3<uint8_t> &par = labels[level + 1];
            const I3 r = cur.n;
            // pass 1: setActiveCellsAndParentList (OG.cpp:394-565) -- per 2x2x2 sibling block:
            // any ACTIVE sibling turns every UP sibling ACTIVE; ACTIVE => parent DOWN.
            std::vector<I3> downList, activeList, upList;
            for (int z = 0; z < r[2]; ++z)
                for (int y = 0; y < r[1]; ++y)
                    for (int x = 0; x < r[0]; ++x) {
                        I3 c = mk(x, y, z);
                        uint8_t v = cur.at(c);
                        if (v == UP) {
                            I3 p = parentOf(c);
                            for (int ch = 0; ch < 8; ++ch)
                                if (cur.get(childCell(p, ch)) == ACTIVE) { cur.at(c) = ACTIVE; break; }
                        } else if (v == ACTIVE)
                            downList.push_back(parentOf(c));
                    }
            for (auto &p : downList) par.at(p) = DOWN;  // setParentCellLabel (OG.cpp:597-654)
            downList.clear();
            // pass 2: setFaceGrading (OG.cpp:656-754)
            for (int z = 0; z < r[2]; ++z)
                for (int y = 0; y < r[1]; ++y)
                    for (int x = 0; x < r[0]; ++x) {
                        I3 c = mk(x, y, z);
                        uint8_t v = cur.at(c);
                        if (v == ACTIVE) {
                            for (int axis = 0; axis < 3; ++axis)
                                for (int dir = 0; dir < 2; ++dir) {
                                    I3 adj = cellToCell(c, axis, dir);
                                    if (adj[axis] < 0 || adj[axis] >= r[axis]) continue;
                                    if (cur.at(adj) == UP) activeList.push_back(parentOf(adj));
                                }
                        } else if (v == DOWN)
                            downList.push_back(parentOf(c));
                    }
            for (auto &p : downList) par.at(p) = DOWN;      // OG.cpp:145
            for (auto &p : activeList) par.at(p) = ACTIVE;  // OG.cpp:162
            // pass 3: setParentsUp (OG.cpp:756-840)
            for (int z = 0; z < r[2]; ++z)
                for (int y = 0; y < r[1]; ++y)
                    for (int x = 0; x < r[0]; ++x) {
                        I3 c = mk(x, y, z);
                        if (cur.at(c) == UP && par.at(parentOf(c)) == INACTIVE) upList.push_back(parentOf(c));
                    }
            for (auto &p : upList) par.at(p) = UP;  // OG.cpp:186
        }
        // setTopLevel (OG.cpp:843-875)
        for (auto &v : labels[L - 1].d)
            if (v == UP) v = ACTIVE;
        // level capping (OG.cpp:198-211)
        int capped = 0;
        for (; capped < L; ++capped) {
            bool has = false;
            for (auto v : labels[capped].d)
                if (v == ACTIVE) { has = true; break; }
            if (!has) break;
        }
        levels = capped;
    }

    // ---------------------------------------------------------------- stage 4/5
    template <class Pred>
    static void markTiles(std::vector<uint8_t> &occ, const I3 &tileRes, const I3 &idx) {
        occ[(size_t)(idx[0] / TILE) + (size_t)tileRes[0] * ((size_t)(idx[1] / TILE) + (size_t)tileRes[1] * (idx[2] / TILE))] = 1;
    }
    static I3 tilesOf(const I3 &r) { return mk((r[0] + TILE - 1) / TILE, (r[1] + TILE - 1) / TILE, (r[2] + TILE - 1) / TILE); }
    static bool tileOcc(const std::vector<uint8_t> &occ, const I3 &tr, int x, int y, int z) {
        return occ[(size_t)(x / TILE) + (size_t)tr[0] * ((size_t)(y / TILE) + (size_t)tr[1] * (z / TILE))] != 0;
    }
    static void setTile(std::vector<uint8_t> &occ, const I3 &tr, const I3 &c) {
        occ[(size_t)(c[0] / TILE) + (size_t)tr[0] * ((size_t)(c[1] / TILE) + (size_t)tr[1] * (c[2] / TILE))] = 1;
    }

    // findOccupiedRegularVelocityTiles (AV.cpp:886-943): faces of every cell with sdf < 2*dx
    void occupiedFromSurface(std::vector<uint8_t> &occ, const I3 &faceGridRes, int axis) const {
        I3 tr = tilesOf(faceGridRes);
        occ.assign((size_t)tr[0] * tr[1] * tr[2], 0);
        const double thr = 2.0 * dx0;
        for (int z = 0; z < N[2]; ++z)
            for (int y = 0; y < N[1]; ++y)
                for (int x = 0; x < N[0]; ++x)
                    if ((double)S.surface.raw(x, y, z) < thr)
                        for (int dir = 0; dir < 2; ++dir) setTile(occ, tr, cellToFace(mk(x, y, z), axis, dir));
    }

    // face activity test shared by AV.cpp:1127-1150 and AV.cpp:1235-1258
    bool faceHasWeight(const I3 &face, int axis) const {
        I3 b = faceToCell(face, axis, 0), f = faceToCell(face, axis, 1);
        if (centerW.get(b) > 0.f || centerW.get(f) > 0.f) return true;
        for (int ea = 0; ea < 3; ++ea) {
            if (ea == axis) continue;
            for (int dir = 0; dir < 2; ++dir)
                if (edgeW[ea].get(faceToEdge(face, axis, ea, dir)) > 0.f) return true;
        }
        return false;
    }

    // number FLUID entries in UT_VoxelArray tile order (AV.cpp:1563-1591 and friends)
    static exint numberInTileOrder(Arr3<exint> &g, exint start, std::vector<int32_t> *keys, int level, int axis) {
        I3 tr = tilesOf(g.n);
        exint idx = start;
        for (int tz = 0; tz < tr[2]; ++tz)
            for (int ty = 0; ty < tr[1]; ++ty)
                for (int tx = 0; tx < tr[0]; ++tx) {
                    int x1 = std::min((tx + 1) * TILE, g.n[0]), y1 = std::min((ty + 1) * TILE, g.n[1]), z1 = std::min((tz + 1) * TILE, g.n[2]);
                    for (int z = tz * TILE; z < z1; ++z)
                        for (int y = ty * TILE; y < y1; ++y)
                            for (int x = tx * TILE; x < x1; ++x) {
                                exint &v = g.d[g.lin(x, y, z)];
                                if (v == FLUID) {
                                    v = idx++;
                                    if (keys) {
                                        keys->push_back(level); keys->push_back(axis);
                                        keys->push_back(x); keys->push_back(y); keys->push_back(z);
                                    }
                                }
                            }
                }
        return idx;
    }

    // buildRegularVelocityIndices (AV.cpp:1445-1512) + classifyRegularVelocityFaces (AV.cpp:1087-1165)
    void buildRegularVelocityIndices() {
        const double extrap = dx0 * P.extrapolation;
        for (int axis = 0; axis < 3; ++axis) {
            I3 r = N;
            r[axis] += 1;
            regIdx[axis].init(r, UNASSIGNED);
            std::vector<uint8_t> occ;
            occupiedFromSurface(occ, r, axis);
            I3 tr = tilesOf(r);
            Arr3<exint> &g = regIdx[axis];
#pragma omp parallel for schedule(static)
            for (int z = 0; z < r[2]; ++z)
                for (int y = 0; y < r[1]; ++y)
                    for (int x = 0; x < r[0]; ++x) {
                        if (!tileOcc(occ, tr, x, y, z)) continue;
                        I3 face = mk(x, y, z);
                        I3 b = faceToCell(face, axis, 0), f = faceToCell(face, axis, 1);
                        if (b[axis] < 0 || f[axis] >= N[axis]) continue;
                        if (faceHasWeight(face, axis)) {
                            double p[3];
                            facePos(face, axis, 0, p);
                            g.at(face) = (S.collision.value(p) > -extrap) ? SOLIDBOUNDARY : FLUID;
                        }
                    }
        }
        exint idx = 0;
        for (int axis = 0; axis < 3; ++axis) idx = numberInTileOrder(regIdx[axis], idx, nullptr, 0, axis);
        regularDOFs = idx;
    }

    // buildOctreeVelocityIndices (AV.cpp:1514-1594) + classifyOctreeVelocityFaces (AV.cpp:1167-1323)
    void buildOctreeVelocityIndices() {
        const double extrap = dx0 * P.extrapolation;
        faceIdx.assign(levels, std::vector<Arr3<exint>>(3));
        for (int level = 0; level < levels; ++level)
            for (int axis = 0; axis < 3; ++axis) {
                I3 cr = cellRes(level);
                I3 r = cr;
                r[axis] += 1;
                Arr3<exint> &g = faceIdx[level][axis];
                g.init(r, UNASSIGNED);
                const Arr3<uint8_t> &lab = labels[level];
                I3 tr = tilesOf(r);
                std::vector<uint8_t> occ;
                if (level == 0) occupiedFromSurface(occ, r, axis);
                else {
                    // findOccupiedOctreeVelocityTiles (AV.cpp:945-1000): faces of ACTIVE cells
                    occ.assign((size_t)tr[0] * tr[1] * tr[2], 0);
                    for (int z = 0; z < cr[2]; ++z)
                        for (int y = 0; y < cr[1]; ++y)
                            for (int x = 0; x < cr[0]; ++x)
                                if (lab.d[lab.lin(x, y, z)] == ACTIVE)
                                    for (int dir = 0; dir < 2; ++dir) setTile(occ, tr, cellToFace(mk(x, y, z), axis, dir));
                }
                I3 vr = (level == 0) ? N : cr;  // AV.cpp:1185-1189
#pragma omp parallel for schedule(static)
                for (int z = 0; z < r[2]; ++z)
                    for (int y = 0; y < r[1]; ++y)
                        for (int x = 0; x < r[0]; ++x) {
                            if (!tileOcc(occ, tr, x, y, z)) continue;
                            I3 face = mk(x, y, z);
                            I3 b = faceToCell(face, axis, 0), f = faceToCell(face, axis, 1);
                            if (b[axis] < 0 || f[axis] >= vr[axis]) {  // AV.cpp:1210-1215
                                if (level == 0) g.at(face) = OUTSIDE;
                                continue;
                            }
                            const int bl = lab.get(b), fl = lab.get(f);
                            if (level == 0) {
                                if (bl == ACTIVE && fl == ACTIVE) {
                                    if (faceHasWeight(face, axis)) {
                                        double p[3];
                                        facePos(face, axis, 0, p);
                                        g.at(face) = (S.collision.value(p) > -extrap) ? SOLIDBOUNDARY : FLUID;
                                    } else g.at(face) = OUTSIDE;
                                } else if (bl == INACTIVE || fl == INACTIVE) g.at(face) = OUTSIDE;
                                else if ((bl == UP && fl == ACTIVE) || (bl == ACTIVE && fl == UP)) g.at(face) = FLUID;
                            } else {
                                if ((bl == ACTIVE && fl == ACTIVE) || (bl == UP && fl == ACTIVE) || (bl == ACTIVE && fl == UP))
                                    g.at(face) = FLUID;
                            }
                        }
            }
        exint idx = 0;
        faceKey.clear();
        for (int level = 0; level < levels; ++level)
            for (int axis = 0; axis < 3; ++axis) idx = numberInTileOrder(faceIdx[level][axis], idx, &faceKey, level, axis);
        nFace = idx;
    }

    // buildEdgeStressIndices (AV.cpp:1596-1663) + classifyEdgeStresses (AV.cpp:1325-1405)
    void buildEdgeStressIndices() {
        edgeIdx.assign(levels, std::vector<Arr3<exint>>(3));
        for (int level = 0; level < levels; ++level)
            for (int axis = 0; axis < 3; ++axis) {
                I3 cr = cellRes(level);
                I3 r = mk(cr[0] + 1, cr[1] + 1, cr[2] + 1);
                r[axis] -= 1;
                Arr3<exint> &g = edgeIdx[level][axis];
                g.init(r, UNASSIGNED);
                const Arr3<uint8_t> &lab = labels[level];
                I3 tr = tilesOf(r);
                // findOccupiedEdgeStressTiles (AV.cpp:1002-1057): edges of ACTIVE cells
                std::vector<uint8_t> occ((size_t)tr[0] * tr[1] * tr[2], 0);
                for (int z = 0; z < cr[2]; ++z)
                    for (int y = 0; y < cr[1]; ++y)
                        for (int x = 0; x < cr[0]; ++x)
                            if (lab.d[lab.lin(x, y, z)] == ACTIVE)
                                for (int e = 0; e < 4; ++e) setTile(occ, tr, cellToEdge(mk(x, y, z), axis, e));
                I3 vr = (level == 0) ? N : cr;  // AV.cpp:1340-1344
#pragma omp parallel for schedule(static)
                for (int z = 0; z < r[2]; ++z)
                    for (int y = 0; y < r[1]; ++y)
                        for (int x = 0; x < r[0]; ++x) {
                            if (!tileOcc(occ, tr, x, y, z)) continue;
                            I3 edge = mk(x, y, z);
                            bool active = false;
                            for (int ci = 0; ci < 4; ++ci) {
                                I3 c = edgeToCell(edge, axis, ci);
                                if (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[0] >= vr[0] || c[1] >= vr[1] || c[2] >= vr[2]) {
                                    g.at(edge) = OUTSIDE;  // AV.cpp:1365-1370 (note: isStressActive keeps its value)
                                    break;
                                }
                                uint8_t l = lab.get(c);
                                if (l == DOWN) { active = false; break; }
                                else if (l == ACTIVE) active = true;
                            }
                            if (active) {
                                if (level == 0) g.at(edge) = (edgeW[axis].get(edge) > 0.f) ? FLUID : OUTSIDE;
                                else g.at(edge) = FLUID;
                            }
                        }
            }
        exint idx = 0;
        for (int level = 0; level < levels; ++level)
            for (int axis = 0; axis < 3; ++axis) idx = numberInTileOrder(edgeIdx[level][axis], idx, nullptr, level, axis);
        nEdge = idx;
    }

    // buildCenterStressIndices (AV.cpp:1665-1715) + classifyCenterStresses (AV.cpp:1407-1443)
    void buildCenterStressIndices() {
        centerIdx.assign(levels, Arr3<exint>());
        exint idx = 0;
        for (int level = 0; level < levels; ++level) {
            I3 cr = cellRes(level);
            centerIdx[level].init(cr, UNASSIGNED);
            const Arr3<uint8_t> &lab = labels[level];
            for (int z = 0; z < cr[2]; ++z)
                for (int y = 0; y < cr[1]; ++y)
                    for (int x = 0; x < cr[0]; ++x)
                        if (lab.d[lab.lin(x, y, z)] == ACTIVE && (level != 0 || centerW.get(mk(x, y, z)) > 0.f))
                            centerIdx[level].d[centerIdx[level].lin(x, y, z)] = FLUID;
        }
        for (int level = 0; level < levels; ++level) idx = numberInTileOrder(centerIdx[level], idx, nullptr, level, 0);
        nCenter = idx;
    }

    // ---------------------------------------------------------------- stage 6/7
    struct Stencil {
        int n = 0, nb = 0;
        exint idx[40];
        double coef[40];
        double bnd[8];
        void add(exint i, double c) { assert(n < 40); idx[n] = i; coef[n] = c; ++n; }
        void addB(double b) { assert(nb < 8); bnd[nb++] = b; }
    };

    // getEdgeStressFaces (AV.cpp:1717-1908)
    void edgeStressFaces(Stencil &st, const I3 &edge, int axis, int level) const {
        st.n = st.nb = 0;
        const double dx = levelDx(level);  // AV.cpp:1733
        bool isAtTransition[3] = {false, false, false};
        bool isFaceOutside[3] = {false, false, false};
        float gradientDx[3] = {0.f, 0.f, 0.f};  // UT_Vector3 (float32) AV.cpp:1738
        const bool enhanced = P.useEnhancedGradients != 0;
        for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
            if (faceAxis == axis) continue;
            const Arr3<exint> &fg = faceIdx[level][faceAxis];
            for (int dir = 0; dir < 2; ++dir) {
                I3 face = edgeToFace(edge, axis, faceAxis, dir);
                const int g = 3 - faceAxis - axis;
                if (face[g] < 0 || face[g] >= fg.n[g]) {
                    gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                    isFaceOutside[g] = true;
                } else {
                    exint vi = fg.get(face);
                    if (vi >= 0) gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                    else if (vi == OUTSIDE || vi == SOLIDBOUNDARY) {
                        gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                        isFaceOutside[g] = true;
                    } else if (vi == UNASSIGNED) {
                        gradientDx[g] = (float)((double)gradientDx[g] + dx);
                        if (enhanced) isAtTransition[g] = true;
                    }
                }
            }
        }
        for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
            if (faceAxis == axis) continue;
            const Arr3<exint> &fg = faceIdx[level][faceAxis];
            for (int dir = 0; dir < 2; ++dir) {
                I3 face = edgeToFace(edge, axis, faceAxis, dir);
                const int g = 3 - faceAxis - axis;
                const double sign = (dir == 0) ? -1 : 1;
                const double gdx = (double)gradientDx[g];
                if (face[g] < 0 || face[g] >= fg.n[g]) continue;
                exint vi = fg.get(face);
                if (vi >= 0) {
                    if (isAtTransition[g] && !isFaceOutside[g]) {  // AV.cpp:1814-1824
                        I3 sib = face;
                        sib[axis] += (edge[axis] % 2 == 0) ? 1 : -1;
                        exint si = fg.get(sib);
                        st.add(si, .25 * sign / gdx);
                        st.add(vi, .25 * sign / gdx);
                    } else
                        st.add(vi, .5 * sign / gdx);
                } else if (vi == UNASSIGNED) {
                    if (edge[faceAxis] % 2 != 0) {  // dangling edge (AV.cpp:1835-1884)
                        for (int off = -1; off <= 1; off += 2) {
                            I3 of = face;
                            of[faceAxis] += off;
                            I3 pf = parentOf(of);
                            exint pi = (level + 1 < levels) ? faceIdx[level + 1][faceAxis].get(pf) : OUTSIDE;
                            if (pi >= 0) st.add(pi, .25 * sign / gdx);
                            else if (pi == UNASSIGNED) {
                                for (int ch = 0; ch < 4; ++ch) {
                                    exint ci = fg.get(childFace(pf, faceAxis, ch));
                                    if (ci >= 0) st.add(ci, .0625 * sign / gdx);
                                }
                            }
                        }
                    } else {  // AV.cpp:1886-1894
                        I3 pf = parentOf(face);
                        exint pi = (level + 1 < levels) ? faceIdx[level + 1][faceAxis].get(pf) : OUTSIDE;
                        st.add(pi, .5 * sign / gdx);
                    }
                } else if (vi == SOLIDBOUNDARY) {  // AV.cpp:1896-1905 -- component `axis` as written
                    double p[3];
                    facePos(face, faceAxis, level, p);
                    double lv = S.collisionVel[axis].value(p);
                    st.addB(.5 * sign * lv / gdx);
                }
            }
        }
    }

    // getCenterStressFaces (AV.cpp:1910-1963)
    void centerStressFaces(Stencil &st, const I3 &cell, int axis, int level) const {
        st.n = st.nb = 0;
        const double dx = levelDx(level);  // AV.cpp:1923
        for (int dir = 0; dir < 2; ++dir) {
            I3 face = cellToFace(cell, axis, dir);
            const double sign = (dir == 0) ? -1 : 1;
            exint vi = faceIdx[level][axis].get(face);
            if (vi >= 0) st.add(vi, sign / dx);
            else if (vi == UNASSIGNED) {
                if (level > 0)
                    for (int ch = 0; ch < 4; ++ch) {
                        exint ci = faceIdx[level - 1][axis].get(childFace(face, axis, ch));
                        st.add(ci, .25 * sign / dx);
                    }
            } else if (vi == SOLIDBOUNDARY) {
                double p[3];
                facePos(face, axis, level, p);
                st.addB(sign * S.collisionVel[axis].value(p) / dx);
            }
        }
    }

    // faceOctreeVolumes (AV.cpp:1965-2002)
    double faceOctreeVolume(const I3 &face, int axis, int level) const {
        const I3 vr = cellRes(level);
        const double dx = (double)(1 << level);
        double g = 0;
        for (int dir = 0; dir < 2; ++dir) {
            I3 c = faceToCell(face, axis, dir);
            if (c[axis] < 0 || c[axis] >= vr[axis]) g += .5 * dx;
            else {
                uint8_t l = labels[level].get(c);
                if (l == ACTIVE || l == INACTIVE) g += .5 * dx;
                else g += dx;  // parent is ACTIVE one level up (asserted in the reference)
            }
        }
        return dx * dx * g;
    }

    // edgeOctreeVolumes (AV.cpp:2004-2057)
    double edgeOctreeVolume(const I3 &edge, int axis, int level) const {
        const double dx = (double)(1 << level);
        float v[3] = {0.f, 0.f, 0.f};  // UT_Vector3 volumeDx
        v[axis] = (float)dx;
        for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
            if (faceAxis == axis) continue;
            const Arr3<exint> &fg = faceIdx[level][faceAxis];
            for (int dir = 0; dir < 2; ++dir) {
                I3 face = edgeToFace(edge, axis, faceAxis, dir);
                const int g = 3 - faceAxis - axis;
                if (face[g] < 0 || face[g] >= fg.n[g]) v[g] = (float)((double)v[g] + .5 * dx);
                else {
                    exint vi = fg.get(face);
                    if (vi >= 0 || vi == OUTSIDE || vi == SOLIDBOUNDARY) v[g] = (float)((double)v[g] + .5 * dx);
                    else if (vi == UNASSIGNED) v[g] = (float)((double)v[g] + dx);
                }
            }
        }
        return (double)(v[0] * v[1] * v[2]);  // float product, as UT_Vector3 components
    }

    // stress weights (AV.cpp:2124-2155 and AV.cpp:2223-2289)
    double edgeStressWeight(const I3 &edge, int axis, int level) const {
        double w;
        if (level == 0) {
            w = (double)edgeW[axis].get(edge);
            if (w == 1.) w = edgeOctreeVolume(edge, axis, level);
        } else w = edgeOctreeVolume(edge, axis, level);
        if (S.viscosity.is_constant()) w *= (double)S.viscosity.constant;
        else {
            double p[3];
            edgePos(edge, axis, level, p);
            w *= S.viscosity.value(p);
        }
        return 4. * P.dt * w;
    }
    double centerStressWeight(const I3 &cell, int level) const {
        double w;
        if (level == 0) w = (double)centerW.get(cell);
        else { double dx = (double)(1 << level); w = dx * dx * dx; }
        if (S.viscosity.is_constant()) w *= (double)S.viscosity.constant;
        else {
            double p[3];
            centerPos(cell, level, p);
            w *= S.viscosity.value(p);
        }
        return 2. * P.dt * w;
    }

    // ---------------------------------------------------------------- stage 8
    // buildVelocityMapping (AV.cpp:2291-2402): recursive 4 children x 3 in-axis offsets
    // The reference walks a FIFO queue and adds weight*leaf into one running sum (AV.cpp:2349-2395);
    // all leaves sit at depth `level`, so queue order == depth-first order of the leaves.
    void restrictFace(double &acc, const I3 &face, int axis, int level, double weight) const {
        if (level == 0) {
            acc += weight * (double)S.vel[axis].raw(face[0], face[1], face[2]);
            return;
        }
        static const double inAxis[3] = {1. / 16., 1. / 8., 1. / 16.};
        for (int ch = 0; ch < 4; ++ch) {
            I3 cf = childFace(face, axis, ch);
            for (int o = -1; o < 2; ++o) {
                I3 af = cf;
                af[axis] += o;
                restrictFace(acc, af, axis, level - 1, (double)(float)(inAxis[o + 1] * weight));  // weight stored as fpreal32 (AV.cpp:2318)
            }
        }
    }
    void buildVelocityMapping() {
        x0.assign((size_t)nFace, 0.0);
#pragma omp parallel for schedule(dynamic, 4096)
        for (exint i = 0; i < nFace; ++i) {
            const int32_t *k = &faceKey[(size_t)i * 5];
            double acc = 0;
            restrictFace(acc, mk(k[2], k[3], k[4]), k[1], k[0], 1.0);
            x0[(size_t)i] = acc;
        }
    }

    // ---------------------------------------------------------------- stage 9
    struct Row {
        std::vector<std::pair<exint, double>> e;
        void add(exint c, double v) {
            for (auto &p : e)
                if (p.first == c) { p.second += v; return; }  // setFromTriplets sums duplicates (AV.cpp:614)
            e.emplace_back(c, v);
        }
    };

    // applyToMatrix (AV.cpp:2404-2457)
    static void applyToMatrix(Row &row, double &rhsI, double &diag, double coefficient, exint vi, const Stencil &st) {
        for (int i = 0; i < st.n; ++i)
            if (st.idx[i] == vi) { coefficient *= st.coef[i]; break; }
        for (int i = 0; i < st.n; ++i) {
            double el = coefficient * st.coef[i];
            if (st.idx[i] == vi) diag += el;
            else row.add(st.idx[i], el);
        }
        for (int i = 0; i < st.nb; ++i) rhsI -= coefficient * st.bnd[i];
    }

    // buildOctreeSystemFromStencilsPartial (AV.cpp:2459-2777), one face row
    void buildRow(exint vi, Row &row, double &rhsI) const {
        const int32_t *k = &faceKey[(size_t)vi * 5];
        const int level = k[0], axis = k[1];
        const I3 face = mk(k[2], k[3], k[4]);
        const I3 vr = cellRes(level);
        const Arr3<exint> &fgrid = faceIdx[level][axis];
        double diag = 0;
        Stencil st;
        for (int dir = 0; dir < 2; ++dir) {
            I3 cell = faceToCell(face, axis, dir);
            if (cell[axis] < 0 || cell[axis] >= vr[axis]) continue;
            I3 sc;
            int sl;
            if (labels[level].get(cell) == ACTIVE) { sc = cell; sl = level; }
            else { sc = parentOf(cell); sl = level + 1; }
            if (sl >= levels) continue;  // cannot happen on a graded tree (asserted AV.cpp:2571)
            exint ci = centerIdx[sl].get(sc);
            if (ci >= 0) {
                centerStressFaces(st, sc, axis, sl);
                applyToMatrix(row, rhsI, diag, centerStressWeight(sc, sl), vi, st);
            }
            // T-junction ghost stresses (AV.cpp:2614-2649)
            for (int fa = 0; fa < 3; ++fa) {
                if (fa == axis) continue;
                for (int fd = 0; fd < 2; ++fd) {
                    I3 af = cellToFace(sc, fa, fd);
                    if (faceIdx[sl][fa].get(af) == UNASSIGNED && sl > 0) {
                        int ea = 3 - fa - axis;
                        for (int ins = 0; ins < 2; ++ins) {
                            I3 e = childEdgeInFace(af, fa, ea, ins);
                            exint ei = edgeIdx[sl - 1][ea].get(e);
                            if (ei >= 0) {
                                edgeStressFaces(st, e, ea, sl - 1);
                                applyToMatrix(row, rhsI, diag, edgeStressWeight(e, ea, sl - 1), vi, st);
                            }
                        }
                    }
                }
            }
        }
        for (int ea = 0; ea < 3; ++ea) {
            if (ea == axis) continue;
            for (int dir = 0; dir < 2; ++dir) {
                I3 e = faceToEdge(face, axis, ea, dir);
                exint ei = edgeIdx[level][ea].get(e);
                if (ei >= 0) {
                    if (P.useEnhancedGradients) {  // AV.cpp:2664-2697
                        const int ta = 3 - ea - axis;
                        I3 af = face;
                        af[ta] += (dir == 0) ? -1 : 1;
                        if (af[ta] >= 0 && af[ta] < fgrid.n[ta] && fgrid.get(af) == UNASSIGNED) {
                            I3 se = e;
                            se[ea] += (e[ea] % 2 == 0) ? 1 : -1;
                            exint ti = edgeIdx[level][ea].get(se);
                            if (ti >= 0) {  // asserted in the reference (AV.cpp:2680)
                                edgeStressFaces(st, se, ea, level);
                                applyToMatrix(row, rhsI, diag, edgeStressWeight(se, ea, level), vi, st);
                            }
                        }
                    }
                    edgeStressFaces(st, e, ea, level);
                    applyToMatrix(row, rhsI, diag, edgeStressWeight(e, ea, level), vi, st);
                } else if (ei == UNASSIGNED && level > 0) {  // AV.cpp:2714-2742
                    for (int ch = 0; ch < 2; ++ch) {
                        I3 ce = childEdge(e, ea, ch);
                        exint cei = edgeIdx[level - 1][ea].get(ce);
                        if (cei >= 0) {
                            edgeStressFaces(st, ce, ea, level - 1);
                            applyToMatrix(row, rhsI, diag, edgeStressWeight(ce, ea, level - 1), vi, st);
                        }
                    }
                }
            }
        }
        // velocity control volume (AV.cpp:2748-2772)
        double fw;
        if (level == 0) {
            fw = (double)S.faceWeights[axis].raw(face[0], face[1], face[2]);
            if (fw == 1.) fw = faceOctreeVolume(face, axis, level);
        } else fw = faceOctreeVolume(face, axis, level);
        if (S.density.is_constant()) fw *= (double)S.density.constant;
        else {
            double p[3];
            facePos(face, axis, level, p);
            fw *= S.density.value(p);
        }
        row.add(vi, fw + diag);
        rhsI += fw * x0[(size_t)vi];
    }

    void buildSystem() {
        rhs.assign((size_t)nFace, 0.0);
        std::vector<Row> rows((size_t)nFace);
#pragma omp parallel for schedule(dynamic, 1024)
        for (exint i = 0; i < nFace; ++i) {
            double r = 0;
            buildRow(i, rows[(size_t)i], r);
            std::sort(rows[(size_t)i].e.begin(), rows[(size_t)i].e.end(),
                      [](const std::pair<exint, double> &a, const std::pair<exint, double> &b) { return a.first < b.first; });
            rhs[(size_t)i] = r;
        }
        rowPtr.assign((size_t)nFace + 1, 0);
        for (exint i = 0; i < nFace; ++i) rowPtr[(size_t)i + 1] = rowPtr[(size_t)i] + (exint)rows[(size_t)i].e.size();
        colIdx.resize((size_t)rowPtr[(size_t)nFace]);
        val.resize((size_t)rowPtr[(size_t)nFace]);
#pragma omp parallel for schedule(static)
        for (exint i = 0; i < nFace; ++i) {
            exint o = rowPtr[(size_t)i];
            for (auto &p : rows[(size_t)i].e) {
                colIdx[(size_t)o] = (int32_t)p.first;
                val[(size_t)o] = p.second;
                ++o;
            }
        }
    }

    // ---------------------------------------------------------------- stage 11
    // HDK_OctreeVectorFieldInterpolator (VFI.h:30-138, VFI.cpp) + applyVelocitiesToRegularGrid (AV.cpp:2815-2894).
    // All node fields are SIM_RawField, i.e. float32 storage: every set rounds to float like the reference.
    enum { INACTIVENODE = 0, ACTIVENODE = 1, DEPENDENTNODE = 2 };

    float octVelAt(int level, int axis, const I3 &f) const { return octVel[level][axis].get(f); }

    void buildInterpolator() {
        const int L = levels;
        // setOctreeVelocity (AV.cpp:2779-2813): fp32 fields, 0 where there is no DOF
        octVel.assign(L, std::vector<Arr3<float>>(3));
        for (int l = 0; l < L; ++l)
            for (int a = 0; a < 3; ++a) {
                const Arr3<exint> &g = faceIdx[l][a];
                octVel[l][a].init(g.n, 0.f);
                for (size_t i = 0; i < g.d.size(); ++i)
                    if (g.d[i] >= 0) octVel[l][a].d[i] = (float)sol[(size_t)g.d[i]];
            }
        nodeLabel.assign(L, Arr3<uint8_t>());
        nodeFlag.assign(L, Arr3<exint>());
        nodeVal.assign(L, std::vector<Arr3<float>>(3));
        nodeW.assign(L, std::vector<Arr3<float>>(3));
        for (int l = 0; l < L; ++l) {
            I3 cr = cellRes(l);
            I3 nr = mk(cr[0] + 1, cr[1] + 1, cr[2] + 1);
            nodeLabel[l].init(nr, INACTIVENODE);
            nodeFlag[l].init(nr, 0);
            for (int a = 0; a < 3; ++a) { nodeVal[l][a].init(nr, 0.f); nodeW[l][a].init(nr, 0.f); }
        }
        // setActiveNodes (VFI.cpp:118-188) + sampleActiveNodes (VFI.cpp:190-286)
        for (int l = 0; l < L; ++l) {
            const I3 nr = nodeLabel[l].n;
            const double weight = (double)(1 << (L - l - 1));
#pragma omp parallel for schedule(static)
            for (int z = 0; z < nr[2]; ++z)
                for (int y = 0; y < nr[1]; ++y)
                    for (int x = 0; x < nr[0]; ++x) {
                        I3 node = mk(x, y, z);
                        bool act = false, inact = false;
                        for (int fa = 0; !inact && fa < 3; ++fa) {
                            const Arr3<exint> &fg = faceIdx[l][fa];
                            const int a1 = (fa + 1) % 3, a2 = (fa + 2) % 3;
                            for (int fi = 0; fi < 4; ++fi) {
                                I3 f = nodeToFace(node, fa, fi);
                                if (f[a1] < 0 || f[a2] < 0 || f[a1] >= fg.n[a1] || f[a2] >= fg.n[a2]) { inact = true; continue; }
                                exint vi = fg.get(f);
                                if (vi >= 0) act = true;
                                else if (vi == SOLIDBOUNDARY || vi == OUTSIDE) { inact = true; break; }
                            }
                        }
                        if (!(act && !inact)) continue;
                        nodeLabel[l].at(node) = ACTIVENODE;
                        exint flag = 0;
                        for (int fa = 0; fa < 3; ++fa) {
                            const Arr3<exint> &fg = faceIdx[l][fa];
                            const int a1 = (fa + 1) % 3, a2 = (fa + 2) % 3;
                            double av = 0, aw = 0;
                            for (int fi = 0; fi < 4; ++fi) {
                                I3 f = nodeToFace(node, fa, fi);
                                if (f[a1] < 0 || f[a2] < 0 || f[a1] >= fg.n[a1] || f[a2] >= fg.n[a2]) {
                                    flag += (exint)(1 << (fa * 4 + fi));
                                    aw += weight;
                                    continue;
                                }
                                exint vi = fg.get(f);
                                if (vi >= 0) { av += weight * (double)octVelAt(l, fa, f); aw += weight; flag += (exint)(1 << (fa * 4 + fi)); }
                                else if (vi != UNASSIGNED) { aw += weight; flag += (exint)(1 << (fa * 4 + fi)); }
                            }
                            nodeVal[l][fa].at(node) = (float)av;
                            nodeW[l][fa].at(node) = (float)aw;
                        }
                        nodeFlag[l].at(node) = flag;
                    }
        }
        // bubbleActiveNodeValues (VFI.cpp:288-355), level by level
        for (int l = 0; l < L - 1; ++l) {
            const I3 nr = nodeLabel[l].n;
            for (int z = 0; z < nr[2]; z += 2)
                for (int y = 0; y < nr[1]; y += 2)
                    for (int x = 0; x < nr[0]; x += 2) {
                        I3 node = mk(x, y, z);
                        if (nodeLabel[l].at(node) != ACTIVENODE) continue;
                        I3 pn = parentOf(node);
                        if (nodeLabel[l + 1].get(pn) != ACTIVENODE) continue;
                        nodeFlag[l + 1].at(pn) = nodeFlag[l].at(node) + nodeFlag[l + 1].at(pn);
                        for (int a = 0; a < 3; ++a) {
                            double w = nodeW[l][a].at(node), v = nodeVal[l][a].at(node);
                            double pw = nodeW[l + 1][a].at(pn), pv = nodeVal[l + 1][a].at(pn);
                            nodeW[l + 1][a].at(pn) = (float)(w + pw);
                            nodeVal[l + 1][a].at(pn) = (float)(v + pv);
                        }
                        nodeLabel[l].at(node) = DEPENDENTNODE;
                    }
        }
        // finishIncompleteNodes (VFI.cpp:357-567)
        for (int l = 0; l < L - 1; ++l) {
            const I3 nr = nodeLabel[l].n;
            const double lw = (double)(1 << (L - l - 1));
#pragma omp parallel for schedule(static)
            for (int z = 0; z < nr[2]; ++z)
                for (int y = 0; y < nr[1]; ++y)
                    for (int x = 0; x < nr[0]; ++x) {
                        I3 node = mk(x, y, z);
                        if (nodeLabel[l].at(node) != ACTIVENODE) continue;
                        exint flag = nodeFlag[l].at(node);
                        if (flag == 0xFFF) continue;
                        exint temp = flag;
                        for (int bit = 0; flag != 0xFFF && bit < 12; ++bit, temp >>= 1) {
                            if (temp & 1) continue;
                            const int fa = bit / 4, fi = bit % 4;
                            bool found = false;
                            if (node[fa] % 2 == 0) {
                                I3 f = nodeToFace(node, fa, fi);
                                I3 pf = parentOf(f);
                                if (faceIdx[l + 1][fa].get(pf) >= 0) {
                                    double ghost = (double)octVelAt(l + 1, fa, pf);
                                    double v = (double)nodeVal[l][fa].at(node);
                                    v += lw * ghost;
                                    nodeVal[l][fa].at(node) = (float)v;
                                    double w = (double)nodeW[l][fa].at(node);
                                    w += lw;
                                    nodeW[l][fa].at(node) = (float)w;
                                    flag += (exint)(1 << bit);
                                    found = true;
                                }
                            }
                            if (!found) {
                                I3 f = nodeToFace(node, fa, fi);
                                I3 cell = faceToCell(f, fa, 1);
                                int sl = l;
                                while (sl < L && labels[sl].get(cell) != ACTIVE) { cell = parentOf(cell); ++sl; }
                                if (sl >= L) { flag += (exint)(1 << bit); continue; }  // asserted impossible (VFI.cpp:490)
                                double fp[3];
                                facePos(f, fa, l, fp);
                                const double idxNode = (fp[fa] - S.origin[fa]) / levelDx(sl);
                                const double iw = idxNode - std::floor(idxNode);
                                double ghost = 0;
                                for (int dir = 0; dir < 2; ++dir) {
                                    I3 of = cellToFace(cell, fa, dir);
                                    exint ovi = faceIdx[sl][fa].get(of);
                                    const double liw = dir == 0 ? 1. - iw : iw;
                                    if (ovi >= 0) ghost += liw * (double)octVelAt(sl, fa, of);
                                    else if (ovi == UNASSIGNED && sl > 0)
                                        for (int ch = 0; ch < 4; ++ch) {
                                            I3 cf = childFace(of, fa, ch);
                                            if (faceIdx[sl - 1][fa].get(cf) >= 0) ghost += .25 * liw * (double)octVelAt(sl - 1, fa, cf);
                                        }
                                }
                                double v = (double)nodeVal[l][fa].at(node);
                                v += lw * ghost;
                                nodeVal[l][fa].at(node) = (float)v;
                                double w = (double)nodeW[l][fa].at(node);
                                w += lw;
                                nodeW[l][fa].at(node) = (float)w;
                                flag += (exint)(1 << bit);
                            }
                        }
                        nodeFlag[l].at(node) = flag;
                    }
        }
        // normalizeActiveNodes (VFI.cpp:569-613)
        for (int l = 0; l < L; ++l)
            for (size_t i = 0; i < nodeLabel[l].d.size(); ++i)
                if (nodeLabel[l].d[i] == ACTIVENODE)
                    for (int a = 0; a < 3; ++a) nodeVal[l][a].d[i] = (float)((double)nodeVal[l][a].d[i] / (double)nodeW[l][a].d[i]);
        // distributeNodeValuesDown (VFI.cpp:615-658), top down
        for (int l = L - 2; l >= 0; --l) {
            const I3 nr = nodeLabel[l].n;
            for (int z = 0; z < nr[2]; ++z)
                for (int y = 0; y < nr[1]; ++y)
                    for (int x = 0; x < nr[0]; ++x) {
                        I3 node = mk(x, y, z);
                        if (nodeLabel[l].at(node) != DEPENDENTNODE) continue;
                        I3 pn = parentOf(node);
                        for (int a = 0; a < 3; ++a) nodeVal[l][a].at(node) = nodeVal[l + 1][a].get(pn);
                        nodeLabel[l].at(node) = ACTIVENODE;
                    }
        }
    }

    // interpSPGrid (VFI.cpp:660-845)
    double interpSPGrid(const double pos[3], int axis) const {
        const int L = levels;
        I3 cell;
        for (int a = 0; a < 3; ++a) cell[a] = (int)std::floor((pos[a] - S.origin[a]) / levelDx(0));
        const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        for (int level = 0; level < L; ++level) {
            if (labels[level].get(cell) == ACTIVE) {
                const double h = levelDx(level);
                double ifp[3];
                I3 face;
                for (int a = 0; a < 3; ++a) {
                    ifp[a] = (pos[a] - S.origin[a]) / h - (a == axis ? 0.0 : 0.5);
                    face[a] = (int)std::floor(ifp[a]);
                }
                const Arr3<exint> &fg = faceIdx[level][axis];
                bool transition = false;
                for (int q = 0; q < 8; ++q)
                    if (fg.get(cellToNode(face, q)) == UNASSIGNED) { transition = true; break; }
                if (!transition) {
                    float iw[3];  // UT_Vector3 interpolationWeight (float32)
                    for (int a = 0; a < 3; ++a) {
                        iw[a] = (float)(ifp[a] - (double)face[a]);
                        iw[a] = std::min(std::max(iw[a], 0.f), 1.f);
                    }
                    double value = 0;
                    for (int q = 0; q < 8; ++q) {
                        I3 nf = cellToNode(face, q);
                        double w = 1.;
                        for (int a = 0; a < 3; ++a) w *= (nf[a] - face[a] == 0) ? (1. - (double)iw[a]) : (double)iw[a];
                        value += w * (double)octVelAt(level, axis, nf);
                    }
                    return value;
                }
                double ciw = (pos[axis] - S.origin[axis]) / h - (double)cell[axis];
                ciw = std::min(std::max(ciw, 0.), 1.);
                double fiv[2] = {0., 0.};
                for (int dir = 0; dir < 2; ++dir) {
                    I3 af = cellToFace(cell, axis, dir);
                    int fl = level;
                    if (fg.get(af) == UNASSIGNED && level > 0) {
                        const double hc = levelDx(level - 1);
                        const double c1 = (pos[a1] - S.origin[a1]) / hc, c2 = (pos[a2] - S.origin[a2]) / hc;
                        for (int ch = 0; ch < 4; ++ch) {
                            I3 cf = childFace(af, axis, ch);
                            if ((double)cf[a1] <= c1 && (double)cf[a2] <= c2 && (double)(cf[a1] + 1) >= c1 && (double)(cf[a2] + 1) >= c2) {
                                fl = level - 1;
                                af = cf;
                                break;
                            }
                        }
                    }
                    const double hf = levelDx(fl);
                    const double n1 = (pos[a1] - S.origin[a1]) / hf, n2 = (pos[a2] - S.origin[a2]) / hf;
                    const double w0 = n1 - std::floor(n1), w1 = n2 - std::floor(n2);
                    const double faceVelocity = (double)octVelAt(fl, axis, af);
                    double avg = 0;
                    for (int q = 0; q < 4; ++q) {
                        I3 node = faceToNode(af, axis, q);
                        double w = 1.;
                        w *= (node[a1] - af[a1] == 0) ? 1. - w0 : w0;
                        w *= (node[a2] - af[a2] == 0) ? 1. - w1 : w1;
                        const double nv = (double)nodeVal[fl][axis].get(node);
                        avg += nv;
                        fiv[dir] += nv * w;
                    }
                    fiv[dir] += 2. * (faceVelocity - .25 * avg) * std::min(w0, std::min(w1, std::min(1. - w0, 1. - w1)));
                }
                return (1. - ciw) * fiv[0] + ciw * fiv[1];
            }
            cell = parentOf(cell);
        }
        return 0.;  // asserted unreachable (VFI.cpp:843)
    }

    // applyVelocitiesToRegularGrid (AV.cpp:2815-2894)
    void applyToRegularGrid() {
        exint count = 0;
        for (int axis = 0; axis < 3; ++axis) {
            const Arr3<exint> &rg = regIdx[axis];
            outVel[axis].init(rg.n, 0.f);
            for (int z = 0; z < rg.n[2]; ++z)
                for (int y = 0; y < rg.n[1]; ++y)
                    for (int x = 0; x < rg.n[0]; ++x) outVel[axis].d[outVel[axis].lin(x, y, z)] = S.vel[axis].raw(x, y, z);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : count)
            for (int z = 0; z < rg.n[2]; ++z)
                for (int y = 0; y < rg.n[1]; ++y)
                    for (int x = 0; x < rg.n[0]; ++x) {
                        I3 face = mk(x, y, z);
                        const exint lab = rg.at(face);
                        double p[3];
                        if (lab >= 0) {
                            const exint oi = faceIdx[0][axis].get(face);
                            if (oi >= 0) outVel[axis].at(face) = (float)sol[(size_t)oi];
                            else if (oi == SOLIDBOUNDARY) {
                                facePos(face, axis, 0, p);
                                outVel[axis].at(face) = (float)S.collisionVel[axis].value(p);
                            } else if (oi == UNASSIGNED) {
                                facePos(face, axis, 0, p);
                                outVel[axis].at(face) = (float)interpSPGrid(p, axis);
                                ++count;
                            }
                        } else if (lab == SOLIDBOUNDARY) {
                            facePos(face, axis, 0, p);
                            outVel[axis].at(face) = (float)S.collisionVel[axis].value(p);
                        }
                    }
        }
        interpolatedFaces = count;
    }
};

// ---------------------------------------------------------------- stage 10
// Eigen::ConjugateGradient<SparseMatrix<T>, Lower|Upper, DiagonalPreconditioner>::solveWithGuess
// (call site AV.cpp:611-630).  Eigen is NOT in /root/reference (un-vendored, version unpinned:
// cmake/FindEIGEN3.cmake:19-31); the loop below restates upstream Eigen 3.3/3.4
// IterativeLinearSolvers/ConjugateGradient.h::conjugate_gradient and
// BasicPreconditioners.h::DiagonalPreconditioner.
template <class T>
static void spmv(exint n, const exint *ptr, const int32_t *col, const T *val, const T *x, T *y) {
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < n; ++i) {
        T s = 0;
        for (exint k = ptr[i]; k < ptr[i + 1]; ++k) s += val[k] * x[col[k]];
        y[i] = s;
    }
}
template <class T>
static T dot(exint n, const T *a, const T *b) {
    // blocked so that the result does not depend on the thread count
    const exint B = 4096;
    const exint nb = (n + B - 1) / B;
    std::vector<T> part((size_t)nb);
#pragma omp parallel for schedule(static)
    for (exint blk = 0; blk < nb; ++blk) {
        T s = 0;
        exint e = std::min(n, (blk + 1) * B);
        for (exint i = blk * B; i < e; ++i) s += a[i] * b[i];
        part[(size_t)blk] = s;
    }
    T s = 0;
    for (exint blk = 0; blk < nb; ++blk) s += part[(size_t)blk];
    return s;
}

template <class T>
static void eigenCG(exint n, const exint *ptr, const int32_t *col, const T *val, const T *b, T *x,
                    double tol_in, int maxIters, int *itersOut, double *errOut) {
    // work vectors: NOT value-initialised (five serial 80 MB memsets per call at 10 M rows); every entry is written before it is
    // read, first touched by the thread that owns it
    std::unique_ptr<T[]> work(new T[5 * (size_t)std::max<exint>(n, 1)]);
    T *invdiag = work.get(), *r = invdiag + n, *p = r + n, *z = p + n, *tmp = z + n;
    // DiagonalPreconditioner::factorize
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < n; ++i) {
        T d = 0;
        bool found = false;
        for (exint k = ptr[i]; k < ptr[i + 1]; ++k)
            if (col[k] == i) { d = val[k]; found = true; break; }
        invdiag[i] = (found && d != T(0)) ? T(1) / d : T(1);
    }
    const T tol = (T)tol_in;
    spmv(n, ptr, col, val, x, tmp);
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
    T rhsNorm2 = dot(n, b, b);
    if (rhsNorm2 == 0) {
        for (exint i = 0; i < n; ++i) x[i] = 0;
        *itersOut = 0;
        *errOut = 0;
        return;
    }
    const T considerAsZero = (std::numeric_limits<T>::min)();
    T threshold = std::max(T(tol * tol * rhsNorm2), considerAsZero);
    T residualNorm2 = dot(n, r, r);
    if (residualNorm2 < threshold) {
        *itersOut = 0;
        *errOut = std::sqrt((double)(residualNorm2 / rhsNorm2));
        return;
    }
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < n; ++i) p[i] = invdiag[i] * r[i];
    T absNew = dot(n, r, p);
    int i = 0;
    while (i < maxIters) {
        spmv(n, ptr, col, val, p, tmp);
        T alpha = absNew / dot(n, p, tmp);
#pragma omp parallel for schedule(static)
        for (exint j = 0; j < n; ++j) {
            x[j] += alpha * p[j];
            r[j] -= alpha * tmp[j];
        }
        residualNorm2 = dot(n, r, r);
        if (residualNorm2 < threshold) break;
#pragma omp parallel for schedule(static)
        for (exint j = 0; j < n; ++j) z[j] = invdiag[j] * r[j];
        T absOld = absNew;
        absNew = dot(n, r, z);
        T beta = absNew / absOld;
#pragma omp parallel for schedule(static)
        for (exint j = 0; j < n; ++j) p[j] = z[j] + beta * p[j];
        ++i;
    }
    *errOut = std::sqrt((double)(residualNorm2 / rhsNorm2));
    *itersOut = i;
}

}  // namespace

// =============================================================================
// C interface (ctypes).  Field descriptors mirror include/avs.h but are declared
// independently on purpose: the oracle shares no code with the product.
// =============================================================================
extern "C" {

struct OrcField {
    const float *data;  // null => constant
    int res[3];
    double org[3];      // world position of sample (0,0,0)
    double dx;
    float constant;
};

struct OrcScene {
    int res[3];
    double origin[3];
    double dx;
    OrcField surface, vel[3], faceWeights[3], viscosity, density, collision, collisionVel[3];
};

struct OrcParams {
    double dt, tolerance, extrapolation;
    int maxIterations, numberSuperSamples, octreeLevels, fineBandwidth;
    int useEnhancedGradients, doApplySolidWeights, singlePrecision;
    int stopAfterStage;  // 0 = everything; 3 = octree only; 5 = labels; 9 = system, no solve
};

static Field toField(Oracle *o, const OrcField &f) {
    Field r;
    for (int a = 0; a < 3; ++a) { r.res[a] = f.res[a]; r.org[a] = f.org[a]; }
    r.dx = f.dx;
    r.constant = f.constant;
    if (f.data) {
        size_t n = (size_t)f.res[0] * f.res[1] * f.res[2];
        o->owned.emplace_back(f.data, f.data + n);
        r.data = o->owned.back().data();
    }
    return r;
}

void *orc_create(const OrcScene *s, const OrcParams *p) {
    Oracle *o = new Oracle();
    o->owned.reserve(32);
    for (int a = 0; a < 3; ++a) { o->S.res[a] = s->res[a]; o->S.origin[a] = s->origin[a]; o->N[a] = s->res[a]; }
    o->S.dx = s->dx;
    o->dx0 = (double)(float)s->dx;  // getVoxelSize() is a float32 vector (AV.cpp:242)
    o->S.surface = toField(o, s->surface);
    for (int a = 0; a < 3; ++a) {
        o->S.vel[a] = toField(o, s->vel[a]);
        o->S.faceWeights[a] = toField(o, s->faceWeights[a]);
        o->S.collisionVel[a] = toField(o, s->collisionVel[a]);
    }
    o->S.viscosity = toField(o, s->viscosity);
    o->S.density = toField(o, s->density);
    o->S.collision = toField(o, s->collision);
    o->P.dt = p->dt; o->P.tolerance = p->tolerance; o->P.extrapolation = p->extrapolation;
    o->P.maxIterations = p->maxIterations; o->P.numberSuperSamples = p->numberSuperSamples;
    o->P.octreeLevels = p->octreeLevels; o->P.fineBandwidth = p->fineBandwidth;
    o->P.useEnhancedGradients = p->useEnhancedGradients; o->P.doApplySolidWeights = p->doApplySolidWeights;
    o->P.singlePrecision = p->singlePrecision;
    return o;
}

void orc_destroy(void *h) { delete (Oracle *)h; }
void orc_set_weight_shortcut(void *h, int on) { ((Oracle *)h)->useShortcut = on != 0; }

// Runs the stages of solveGasSubclass in order (AV.cpp:233-653).
int orc_run(void *h, int stopAfterStage) {
    Oracle *o = (Oracle *)h;
    const int stop = stopAfterStage <= 0 ? 100 : stopAfterStage;
    o->buildIntegrationWeights(); o->stage = 1; if (stop <= 1) return 0;
    o->buildMask();               o->stage = 2; if (stop <= 2) return 0;
    o->buildOctree();             o->stage = 3; if (stop <= 3) return 0;
    o->buildRegularVelocityIndices(); o->stage = 4; if (stop <= 4) return 0;
    o->buildOctreeVelocityIndices();
    o->buildEdgeStressIndices();
    o->buildCenterStressIndices(); o->stage = 5; if (stop <= 5) return 0;
    o->buildVelocityMapping();    o->stage = 8; if (stop <= 8) return 0;
    o->buildSystem();             o->stage = 9; if (stop <= 9) return 0;
    o->sol = o->x0;  // solveWithGuess(rhs, viscositySolution) (AV.cpp:627)
    if (o->P.singlePrecision) {
        size_t n = (size_t)o->nFace, nnz = o->val.size();
        std::vector<float> v(nnz), b(n), x(n);
        for (size_t i = 0; i < nnz; ++i) v[i] = (float)o->val[i];
        for (size_t i = 0; i < n; ++i) { b[i] = (float)o->rhs[i]; x[i] = (float)o->sol[i]; }
        eigenCG<float>(o->nFace, o->rowPtr.data(), o->colIdx.data(), v.data(), b.data(), x.data(),
                       o->P.tolerance, o->P.maxIterations, &o->iterations, &o->error);
        for (size_t i = 0; i < n; ++i) o->sol[i] = (double)x[i];
    } else {
        eigenCG<double>(o->nFace, o->rowPtr.data(), o->colIdx.data(), o->val.data(), o->rhs.data(), o->sol.data(),
                        o->P.tolerance, o->P.maxIterations, &o->iterations, &o->error);
    }
    o->stage = 10;
    if (stop <= 10) return 0;
    o->buildInterpolator();
    o->applyToRegularGrid();
    o->stage = 11;
    return 0;
}

// ---- getters ---------------------------------------------------------------
int orc_levels(void *h) { return ((Oracle *)h)->levels; }
int orc_levels_allocated(void *h) { return ((Oracle *)h)->levelsAllocated; }
void orc_padded_res(void *h, int *out) { for (int a = 0; a < 3; ++a) out[a] = ((Oracle *)h)->Pad[a]; }
int64_t orc_count(void *h, int what) {
    Oracle *o = (Oracle *)h;
    switch (what) {
        case 0: return o->nFace;
        case 1: return o->nEdge;
        case 2: return o->nCenter;
        case 3: return o->regularDOFs;
        case 4: return (int64_t)o->val.size();
        case 5: return o->iterations;
    }
    return -1;
}
double orc_error(void *h) { return ((Oracle *)h)->error; }

// kind: 0 centre weights, 1..3 edge weights, 4 mask.  Returns element count; copies when out != null.
int64_t orc_get_float(void *h, int kind, float *out, int *res) {
    Oracle *o = (Oracle *)h;
    const Arr3<float> *a = kind == 0 ? &o->centerW : kind <= 3 ? &o->edgeW[kind - 1] : &o->mask;
    if (res) for (int i = 0; i < 3; ++i) res[i] = a->n[i];
    if (out) std::memcpy(out, a->d.data(), a->d.size() * sizeof(float));
    return (int64_t)a->d.size();
}
int64_t orc_get_labels(void *h, int level, uint8_t *out, int *res) {
    Oracle *o = (Oracle *)h;
    const Arr3<uint8_t> &a = o->labels[level];
    if (res) for (int i = 0; i < 3; ++i) res[i] = a.n[i];
    if (out) std::memcpy(out, a.d.data(), a.d.size());
    return (int64_t)a.d.size();
}
// kind: 0 face, 1 edge, 2 centre, 3 regular face (level ignored)
int64_t orc_get_index_grid(void *h, int kind, int level, int axis, int64_t *out, int *res) {
    Oracle *o = (Oracle *)h;
    const Arr3<exint> *a = kind == 0 ? &o->faceIdx[level][axis] : kind == 1 ? &o->edgeIdx[level][axis]
                         : kind == 2 ? &o->centerIdx[level] : &o->regIdx[axis];
    if (res) for (int i = 0; i < 3; ++i) res[i] = a->n[i];
    if (out) std::memcpy(out, a->d.data(), a->d.size() * sizeof(int64_t));
    return (int64_t)a->d.size();
}
void orc_get_face_keys(void *h, int32_t *out) {
    Oracle *o = (Oracle *)h;
    std::memcpy(out, o->faceKey.data(), o->faceKey.size() * sizeof(int32_t));
}
// what: 0 x0 (restricted u^n), 1 rhs, 2 solution
void orc_get_vector(void *h, int what, double *out) {
    Oracle *o = (Oracle *)h;
    const std::vector<double> &v = what == 0 ? o->x0 : what == 1 ? o->rhs : o->sol;
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}
// regular-grid velocity after stage 11 (float32, shape of the input vel component)
int64_t orc_get_out_velocity(void *h, int axis, float *out, int *res) {
    Oracle *o = (Oracle *)h;
    const Arr3<float> &a = o->outVel[axis];
    if (res) for (int i = 0; i < 3; ++i) res[i] = a.n[i];
    if (out) std::memcpy(out, a.d.data(), a.d.size() * sizeof(float));
    return (int64_t)a.d.size();
}
// node data of the interpolator: kind 0 labels (as float), 1..3 node values of axis kind-1
int64_t orc_get_node_grid(void *h, int kind, int level, float *out, int *res) {
    Oracle *o = (Oracle *)h;
    const I3 n = o->nodeLabel[level].n;
    if (res) for (int i = 0; i < 3; ++i) res[i] = n[i];
    size_t cnt = o->nodeLabel[level].d.size();
    if (out) {
        if (kind == 0) for (size_t i = 0; i < cnt; ++i) out[i] = (float)o->nodeLabel[level].d[i];
        else std::memcpy(out, o->nodeVal[level][kind - 1].d.data(), cnt * sizeof(float));
    }
    return (int64_t)cnt;
}
int64_t orc_interpolated_faces(void *h) { return ((Oracle *)h)->interpolatedFaces; }

// HDK_OctreeGrid::outputOctreeGeometry (OG.cpp:245-308, called from AV.cpp:283-294 under doPrintOctree): one point per
// ACTIVE cell of every built level: P = indexToPos(cell) stored as UT_Vector3 (fp32), pscale = the level's voxel size,
// octreeLevel = level.  The reference appends in tile order; here level-major, x-fastest (the point SET is the contract).
int64_t orc_get_octree_points(void *h, float *pos, float *pscale, int32_t *level) {
    Oracle *o = (Oracle *)h;
    int64_t n = 0;
    for (int l = 0; l < o->levels; ++l) {
        const Arr3<uint8_t> &g = o->labels[l];
        for (int z = 0; z < g.n[2]; ++z)
            for (int y = 0; y < g.n[1]; ++y)
                for (int x = 0; x < g.n[0]; ++x) {
                    if (g.d[g.lin(x, y, z)] != ACTIVE) continue;   // OG.cpp:286
                    if (pos) {
                        double p[3];
                        o->centerPos(mk(x, y, z), l, p);           // OG.cpp:289-292
                        for (int a = 0; a < 3; ++a) pos[3 * n + a] = (float)p[a];
                        pscale[n] = (float)o->levelDx(l);          // OG.cpp:273, 295
                        level[n] = l;                              // OG.cpp:296
                    }
                    ++n;
                }
    }
    return n;
}

void orc_get_csr(void *h, int64_t *rowPtr, int32_t *col, double *val) {
    Oracle *o = (Oracle *)h;
    std::memcpy(rowPtr, o->rowPtr.data(), o->rowPtr.size() * sizeof(int64_t));
    std::memcpy(col, o->colIdx.data(), o->colIdx.size() * sizeof(int32_t));
    std::memcpy(val, o->val.data(), o->val.size() * sizeof(double));
}
// one stencil row of D by geometric key (for debugging / property tests)
// kind 0: edge stress (AV.cpp:1717), kind 1: centre stress (AV.cpp:1910).
int orc_stencil(void *h, int kind, int level, int axis, int i, int j, int k,
                int64_t *idx, double *coef, int *nb, double *bnd, double *weight) {
    Oracle *o = (Oracle *)h;
    Oracle::Stencil st;
    I3 c = mk(i, j, k);
    if (kind == 0) { o->edgeStressFaces(st, c, axis, level); *weight = o->edgeStressWeight(c, axis, level); }
    else { o->centerStressFaces(st, c, axis, level); *weight = o->centerStressWeight(c, level); }
    for (int q = 0; q < st.n; ++q) { idx[q] = st.idx[q]; coef[q] = st.coef[q]; }
    *nb = st.nb;
    for (int q = 0; q < st.nb; ++q) bnd[q] = st.bnd[q];
    return st.n;
}

// ---- stand-alone linear algebra (CPU baseline of the CG hot loop) -----------
void orc_spmv_f64(int64_t n, const int64_t *ptr, const int32_t *col, const double *val, const double *x, double *y) {
    spmv<double>(n, ptr, col, val, x, y);
}
void orc_spmv_f32(int64_t n, const int64_t *ptr, const int32_t *col, const float *val, const float *x, float *y) {
    spmv<float>(n, ptr, col, val, x, y);
}
void orc_cg_f64(int64_t n, const int64_t *ptr, const int32_t *col, const double *val, const double *b, double *x,
                double tol, int maxIters, int *iters, double *err) {
    eigenCG<double>(n, ptr, col, val, b, x, tol, maxIters, iters, err);
}
void orc_cg_f32(int64_t n, const int64_t *ptr, const int32_t *col, const float *val, const float *b, float *x,
                double tol, int maxIters, int *iters, double *err) {
    eigenCG<float>(n, ptr, col, val, b, x, tol, maxIters, iters, err);
}
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
