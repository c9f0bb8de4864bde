// tests/host_labels.cu -- TEST INFRASTRUCTURE (built and used only by tests/test_host_assembly.py; never part of libavs_b200.so).
//
// Stages 1-5 of the CUDA library as far as they live in device FUNCTIONS of csrc/avs_labels.cu: the super-sampling weight sampler
// (sdfWeightSample = computeSDFWeightsSampled, AV.cpp:712-791), the refinement mask / base labels and the octree passes (baseLabel,
// octreePass1 / 2Down / 2Active / 3 = AV.cpp:839-860, OG.cpp:310-840) and the per-sample classification rules (faceHasWeight,
// classifyRegular, classifyFace, classifyEdge, classifyCenter = AV.cpp:1087-1443) -- compiled for the HOST (-DAVS_HOST_TEST) and run on the cell
// labels of a run of the REFERENCE'S OWN CODE; the weights and the face / edge / centre / regular labels they produce are compared
// with the reference's.  Not the product's code: the loops over samples, the marking of the occupied tiles (the kernels
// k_mark_surface_tiles / k_tile_flags, restated below with the library's markTile); the sign-class shortcuts of the weight kernels only exist on the GPU.
#ifndef AVS_HOST_TEST
#error "compile with -DAVS_HOST_TEST"
#endif
#include <vector>

#include "../adaptiveviscositysolver_b200/csrc/avs_labels.cu"
#include "host_scene.h"

extern "C" {

// Stage 1 (avs_stage_weights): the four weight grids by dense sampling, then the solid-weight division (AV.cpp:772-790).
void host_weights(const HostSceneDesc *h, const HostField *surface, const HostField *collision, int n, int solidWeights, float *centerW,
                  float *edgeW0, float *edgeW1, float *edgeW2) {
    DeviceScene S;
    fillScene(*h, S);
    S.surface = toField(*surface);
    S.collision = toField(*collision);
    float *outs[4] = {centerW, edgeW0, edgeW1, edgeW2};
    const double offs[4][3] = {{0.5, 0.5, 0.5}, {0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0.5}};
    for (int gi = 0; gi < 4; ++gi) {
        Grid3<float> g;
        g.d = outs[gi];
        for (int k = 0; k < 3; ++k) g.n[k] = S.N[k] + ((gi == 0 || k == gi - 1) ? 0 : 1);
        for (int z = 0; z < g.n[2]; ++z)
            for (int y = 0; y < g.n[1]; ++y)
                for (int x = 0; x < g.n[0]; ++x)
                    sdfWeightSample(g, S.surface, offs[gi][0], offs[gi][1], offs[gi][2], S.origin[0], S.origin[1], S.origin[2], S.dx0, n, 0.0, nullptr, x,
                                    y, z, g.lin(x, y, z));
        if (solidWeights) {
            std::vector<float> tmp(g.count());
            Grid3<float> s = g;
            s.d = tmp.data();
            for (int z = 0; z < g.n[2]; ++z)
                for (int y = 0; y < g.n[1]; ++y)
                    for (int x = 0; x < g.n[0]; ++x)
                        sdfWeightSample(s, S.collision, offs[gi][0], offs[gi][1], offs[gi][2], S.origin[0], S.origin[1], S.origin[2], S.dx0, n, -S.extrap,
                                        nullptr, x, y, z, s.lin(x, y, z));
            for (size_t i = 0; i < g.count(); ++i)
                if (s.d[i] > 0.f) g.d[i] = g.d[i] / s.d[i];   // k_divide_where_positive
        }
    }
}

// Stages 2 + 3 (avs_stage_octree): refinement mask fused with the base labels (baseLabel = AV.cpp:839-860 + OG.cpp:310-392), then per
// level the passes of HDK_OctreeGrid::init in the order the stage launches them (octreePass1 = setActiveCellsAndParentList,
// octreePass2Down / octreePass2Active = setFaceGrading, octreePass3 = setParentsUp), setTopLevel and the level cap (OG.cpp:198-211).
// labelOut[l] must hold (Pad >> l)^3 bytes for l < allocated levels; returns the number of levels built, *allocated = levels allocated.
// scalarPasses != 0: the one-cell-per-thread pass functions on every level (what the stage uses where a row is not a multiple of 16).
int host_octree(const HostSceneDesc *h, const HostField *surface, const HostField *collision, int octreeLevels, int fineBandwidth,
                uint8_t *const *labelOut, int *allocated, int scalarPasses) {
    DeviceScene S;
    fillScene(*h, S);
    S.surface = toField(*surface);
    S.collision = toField(*collision);
    int L = octreeLevels;  // OG.cpp:32-40
    for (int a = 0; a < 3; ++a) L = std::min(L, ilog2(S.Pad[a]));
    if (L < 1) L = 1;
    if (L > AVS_MAX_LEVELS) L = AVS_MAX_LEVELS;
    *allocated = L;
    for (int l = 0; l < L; ++l) {
        Grid3<uint8_t> &g = S.label[l];
        for (int a = 0; a < 3; ++a) g.n[a] = S.Pad[a] >> l;
        g.d = labelOut[l];
        if (l > 0) memset(g.d, L_INACTIVE, g.count());
    }
    const double fineVoxelWidth = std::max(2.0, (double)fineBandwidth);      // AV.cpp:259
    const double inner = S.dx0 * fineVoxelWidth, outer = 3.0 * S.dx0;         // AV.cpp:261-262
#define FOR_CELLS(g, body)                           \
    for (int z = 0; z < (g).n[2]; ++z)               \
        for (int y = 0; y < (g).n[1]; ++y)           \
            for (int x = 0; x < (g).n[0]; ++x) { const size_t idx = (g).lin(x, y, z); body; }
    FOR_CELLS(S.label[0], S.label[0].d[idx] = baseLabel(S, inner, outer, x, y, z))
    for (int l = 0; l < L - 1; ++l) {
        Grid3<uint8_t> cur = S.label[l], par = S.label[l + 1];
        FOR_CELLS(par, octreePass1(cur, par, x, y, z, idx))
        if (l > 0) FOR_CELLS(cur, octreePass2Down(cur, par, x, y, z, idx))
        if (cur.n[0] % 16 == 0 && !scalarPasses) {   // the stage's choice: 16 cells per thread where the rows allow it
            for (int z = 0; z < cur.n[2]; ++z)
                for (int y = 0; y < cur.n[1]; ++y)
                    for (int x0 = 0; x0 < cur.n[0]; x0 += 16) octreePass2Active16(cur, par, x0, y, z);
            for (int z = 0; z < cur.n[2]; ++z)
                for (int y = 0; y < cur.n[1]; ++y)
                    for (int x0 = 0; x0 < cur.n[0]; x0 += 16) octreePass3_16(cur, par, x0, y, z);
        } else {
            FOR_CELLS(cur, octreePass2Active(cur, par, x, y, z, idx))
            FOR_CELLS(cur, octreePass3(cur, par, x, y, z, idx))
        }
    }
    {
        Grid3<uint8_t> &top = S.label[L - 1];   // k_octree_top, setTopLevel (OG.cpp:843-875)
        for (size_t i = 0; i < top.count(); ++i)
            if (top.d[i] == L_UP) top.d[i] = L_ACTIVE;
    }
#undef FOR_CELLS
    int capped = 0;                              // first level without an ACTIVE cell
    for (; capped < L; ++capped) {
        bool any = false;
        for (size_t i = 0; i < S.label[capped].count() && !any; ++i) any = S.label[capped].d[i] == L_ACTIVE;
        if (!any) break;
    }
    return capped;
}

// Stages 4 and 5: labels of the regular-grid faces and of the octree faces / edges / centres of every level, from the cell labels
// and weights in `h` (the reference's).  face[l][a] receives the LABEL (0 = degree of freedom, else -1 / -2 / -3).
void host_classify(const HostSceneDesc *h, const HostField *surface, const HostField *collision, int8_t *const *regularOut, int8_t *const *faceOut,
                   int8_t *const *edgeOut, int8_t *const *centerOut) {
    DeviceScene S;
    fillScene(*h, S);
    S.surface = toField(*surface);
    S.collision = toField(*collision);
    // ---- occupied face tiles of level 0 / the regular grid (findOccupiedRegularVelocityTiles, AV.cpp:886-943)
    std::vector<uint8_t> tbuf[3];
    Grid3<uint8_t> ft[3];
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 3; ++k) ft[a].n[k] = (S.Pad[k] + 1 + AVS_TILE - 1) / AVS_TILE;
        tbuf[a].assign(ft[a].count(), 0);
        ft[a].d = tbuf[a].data();
    }
    for (int z = 0; z < S.N[2]; ++z)
        for (int y = 0; y < S.N[1]; ++y)
            for (int x = 0; x < S.N[0]; ++x) {
                if (!((double)S.surface.raw(x, y, z) < 2.0 * S.dx0)) continue;   // k_mark_surface_tiles
                const I3 c = mk3(x, y, z);
                for (int a = 0; a < 3; ++a) { markTile(ft[a], c); markTile(ft[a], cellToFace(c, a, 1)); }
            }
    // ---- regular-grid faces (classifyRegular = classifyRegularVelocityFaces AV.cpp:1087-1165)
    for (int axis = 0; axis < 3; ++axis) {
        Grid3<int8_t> g;
        g.d = regularOut[axis];
        for (int k = 0; k < 3; ++k) g.n[k] = S.N[k] + (k == axis);
        for (int z = 0; z < g.n[2]; ++z)
            for (int y = 0; y < g.n[1]; ++y)
                for (int x = 0; x < g.n[0]; ++x) g.d[g.lin(x, y, z)] = classifyRegular(S, axis, ft[axis], x, y, z);
    }
    // ---- octree levels
    for (int l = 0; l < S.levels; ++l) {
        const Grid3<uint8_t> &lab = S.label[l];
        // occupied edge tiles: the 4 a-edges of every ACTIVE cell (findOccupiedEdgeStressTiles AV.cpp:1002-1057; k_tile_flags)
        std::vector<uint8_t> ebuf[3];
        Grid3<uint8_t> et[3];
        for (int a = 0; a < 3; ++a) {
            for (int k = 0; k < 3; ++k) et[a].n[k] = (S.edge[l][a].n[k] + AVS_TILE - 1) / AVS_TILE;
            ebuf[a].assign(et[a].count(), 0);
            et[a].d = ebuf[a].data();
        }
        for (int z = 0; z < lab.n[2]; ++z)
            for (int y = 0; y < lab.n[1]; ++y)
                for (int x = 0; x < lab.n[0]; ++x) {
                    if (lab.d[lab.lin(x, y, z)] != L_ACTIVE) continue;
                    const I3 c = mk3(x, y, z);
                    for (int e = 0; e < 4; ++e)
                        for (int a = 0; a < 3; ++a) markTile(et[a], cellToEdge(c, a, e));
                }
        for (int a = 0; a < 3; ++a) {
            const Grid3<int32_t> &fg = S.face[l][a];
            int8_t *fo = faceOut[l * 3 + a];
            for (int z = 0; z < fg.n[2]; ++z)
                for (int y = 0; y < fg.n[1]; ++y)
                    for (int x = 0; x < fg.n[0]; ++x) fo[fg.lin(x, y, z)] = (int8_t)classifyFace(S, l, a, ft[a], x, y, z);
            const Grid3<int8_t> &eg = S.edge[l][a];
            int8_t *eo = edgeOut[l * 3 + a];
            for (int z = 0; z < eg.n[2]; ++z)
                for (int y = 0; y < eg.n[1]; ++y)
                    for (int x = 0; x < eg.n[0]; ++x) eo[eg.lin(x, y, z)] = classifyEdge(S, l, a, et[a], x, y, z);
        }
        const Grid3<int8_t> &cg = S.center[l];
        int8_t *co = centerOut[l];
        for (int z = 0; z < cg.n[2]; ++z)
            for (int y = 0; y < cg.n[1]; ++y)
                for (int x = 0; x < cg.n[0]; ++x) co[cg.lin(x, y, z)] = classifyCenter(S, l, x, y, z, cg.lin(x, y, z));
    }
}

}  // extern "C"
