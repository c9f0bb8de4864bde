// tests/host_prolong.cu -- TEST INFRASTRUCTURE (built and used only by tests/test_host_assembly.py; never part of libavs_b200.so).
//
// Stage 11 of the CUDA library -- the interpolator's node pyramid (nodeSample, nodeBubble, nodeFinish, nodeNormalize,
// nodeDistribute) and interpSPGrid of csrc/avs_prolong.cu -- compiled for the HOST (-DAVS_HOST_TEST: AVS_DEV = __host__ __device__)
// and run on the label grids AND the solution vector of a run of the REFERENCE'S OWN CODE; the regular-grid velocity it writes is
// compared with the reference's.  The passes run in the order avs_apply_regular launches them (every level sampled, bubbled up,
// finished, normalised, distributed down); within a pass the nodes are visited serially, which a kernel's threads do in parallel.
// Not the product's code: the loops over nodes / faces and the per-face dispatch of k_apply_regular, restated below line for line.
#ifndef AVS_HOST_TEST
#error "compile with -DAVS_HOST_TEST"
#endif
#include <vector>

#include "../adaptiveviscositysolver_b200/csrc/avs_prolong.cu"
#include "host_scene.h"

extern "C" long long host_apply_regular(const HostSceneDesc *h, const double *sol, float *out0, float *out1, float *out2) {
    DeviceScene S;
    fillScene(*h, S);
    const int L = S.levels;
    NodeScene NS;
    memset(&NS, 0, sizeof(NS));
    std::vector<std::vector<float>> fbuf;
    std::vector<std::vector<uint16_t>> hbuf;
    std::vector<std::vector<uint8_t>> bbuf;
    if (L > 1) {
        fbuf.resize((size_t)L * 6);
        hbuf.resize((size_t)L);
        bbuf.resize((size_t)L);
        for (int l = 0; l < L; ++l) {
            NodeLevel &nl = NS.lv[l];
            int n[3];
            for (int a = 0; a < 3; ++a) n[a] = S.label[l].n[a] + 1;
            const size_t cnt = (size_t)n[0] * n[1] * n[2];
            for (int k = 0; k < 6; ++k) fbuf[(size_t)l * 6 + k].assign(cnt, 0.f);
            hbuf[(size_t)l].assign(cnt, 0);
            bbuf[(size_t)l].assign(cnt, N_INACTIVE);
            for (int a = 0; a < 3; ++a) {
                nl.val[a].d = fbuf[(size_t)l * 6 + a].data();
                nl.w[a].d = fbuf[(size_t)l * 6 + 3 + a].data();
                for (int k = 0; k < 3; ++k) { nl.val[a].n[k] = n[k]; nl.w[a].n[k] = n[k]; }
            }
            nl.flag.d = hbuf[(size_t)l].data();
            nl.label.d = bbuf[(size_t)l].data();
            for (int k = 0; k < 3; ++k) { nl.flag.n[k] = n[k]; nl.label.n[k] = n[k]; }
        }
#define FOR_NODES(nl, body)                                     \
    for (int z = 0; z < (nl).label.n[2]; ++z)                   \
        for (int y = 0; y < (nl).label.n[1]; ++y)               \
            for (int x = 0; x < (nl).label.n[0]; ++x) { body; }
        for (int l = 0; l < L; ++l) FOR_NODES(NS.lv[l], nodeSample(S, NS.lv[l], sol, l, x, y, z))
        for (int l = 0; l < L - 1; ++l) FOR_NODES(NS.lv[l + 1], nodeBubble(NS.lv[l], NS.lv[l + 1], x, y, z))
        for (int l = 0; l < L - 1; ++l) FOR_NODES(NS.lv[l], nodeFinish(S, NS.lv[l], sol, l, x, y, z))
        for (int l = 0; l < L; ++l) FOR_NODES(NS.lv[l], nodeNormalize(NS.lv[l], x, y, z))
        for (int l = L - 2; l >= 0; --l) FOR_NODES(NS.lv[l], nodeDistribute(NS.lv[l], NS.lv[l + 1], x, y, z))
#undef FOR_NODES
    }
    float *outs[3] = {out0, out1, out2};
    long long interpolated = 0;
    for (int axis = 0; axis < 3; ++axis) {
        const Grid3<int8_t> g = S.regular[axis];
        float *out = outs[axis];
        for (int z = 0; z < g.n[2]; ++z)
            for (int y = 0; y < g.n[1]; ++y)
                for (int x = 0; x < g.n[0]; ++x) {
                    // ---- k_apply_regular, restated (applyVelocitiesToRegularGrid, AV.cpp:2815-2894)
                    const size_t idx = g.lin(x, y, z);
                    const int8_t lab = g.d[idx];
                    const I3 face = mk3(x, y, z);
                    double p[3];
                    if (lab == F_SOLID) {
                        S.facePos(face, axis, 0, p);
                        out[idx] = (float)S.collisionVel[axis].value(p);
                    } else if (lab >= 0) {
                        int32_t oi = S.face[0][axis].get(face);
                        if (oi >= 0) out[idx] = (float)sol[oi];
                        else if (oi == F_SOLID) {
                            S.facePos(face, axis, 0, p);
                            out[idx] = (float)S.collisionVel[axis].value(p);
                        } else if (oi == F_UNASSIGNED) {
                            S.facePos(face, axis, 0, p);
                            out[idx] = (float)interpSPGrid(S, NS, sol, p, axis);
                            ++interpolated;
                        }
                    }
                }
    }
    return interpolated;
}
