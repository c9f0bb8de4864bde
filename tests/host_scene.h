// tests/host_scene.h -- TEST INFRASTRUCTURE: the scene description the host harnesses (host_assembly.cu, host_prolong.cu) receive from
// Python, and the DeviceScene they fill from it the way the library's stages do -- with HOST pointers.
#pragma once
struct HostField {          // one scalar component, as AvsField of include/avs.h (host pointer or constant)
    const float *data;
    int32_t res[3];
    double org[3];
    double dx;
    float constant;
};
struct HostSceneDesc {
    int32_t N[3];           // liquid surface resolution
    int32_t levels;         // built octree levels
    double origin[3], dx, dt, extrapolation;
    int32_t enhanced;
    HostField viscosity, density, collisionVel[3], faceW[3];
    const float *centerW;
    const float *edgeW[3];
    const uint8_t *label[AVS_MAX_LEVELS];
    const int32_t *face[AVS_MAX_LEVELS][3];   // >= 0: DOF (row) index, else UNASSIGNED / SOLIDBOUNDARY / OUTSIDE
    const int8_t *edge[AVS_MAX_LEVELS][3];    // 0 active, else the negative label
    const int8_t *center[AVS_MAX_LEVELS];
    HostField vel[3];       // u^n, face sampled (read by the restriction only)
    const int8_t *regular[3];   // regular-grid face labels: 0 solved face, else the negative label (read by the write-back only)
};

static DField toField(const HostField &f) {   // csrc/avs_api.cu uploadField: same members, host pointer instead of a device copy
    DField d;
    d.d = f.data;
    for (int a = 0; a < 3; ++a) { d.n[a] = f.res[a]; d.org[a] = f.org[a]; }
    d.dx = f.dx;
    d.constant = f.constant;
    return d;
}

// the part of DeviceScene the row builder reads, filled like avs_stage_upload / avs_stage_octree / avs_stage_octree_labels do
static void fillScene(const HostSceneDesc &h, DeviceScene &S) {
    memset(&S, 0, sizeof(S));
    for (int a = 0; a < 3; ++a) {
        S.N[a] = h.N[a];
        S.origin[a] = h.origin[a];
        int pad = 1;
        while (pad < h.N[a]) pad <<= 1;   // OG.cpp:18-24
        S.Pad[a] = pad;
    }
    S.levels = h.levels;
    S.dx0 = (double)(float)h.dx;          // AV.cpp:242
    S.dt = h.dt;
    S.extrap = S.dx0 * h.extrapolation;   // AV.cpp:243
    S.enhanced = h.enhanced ? 1 : 0;
    S.viscosity = toField(h.viscosity);
    S.density = toField(h.density);
    for (int a = 0; a < 3; ++a) {
        S.collisionVel[a] = toField(h.collisionVel[a]);
        S.faceW[a] = toField(h.faceW[a]);
        S.vel[a] = toField(h.vel[a]);
    }
    S.centerW.d = (float *)h.centerW;
    for (int k = 0; k < 3; ++k) S.centerW.n[k] = h.N[k];
    for (int a = 0; a < 3; ++a) {
        S.edgeW[a].d = (float *)h.edgeW[a];
        for (int k = 0; k < 3; ++k) S.edgeW[a].n[k] = h.N[k] + (k != a);
    }
    for (int l = 0; l < h.levels; ++l) {
        Grid3<uint8_t> &lab = S.label[l];
        lab.d = (uint8_t *)h.label[l];
        for (int k = 0; k < 3; ++k) lab.n[k] = S.Pad[k] >> l;
        for (int a = 0; a < 3; ++a) {
            S.face[l][a].d = (int32_t *)h.face[l][a];
            S.edge[l][a].d = (int8_t *)h.edge[l][a];
            for (int k = 0; k < 3; ++k) {
                S.face[l][a].n[k] = lab.n[k] + (k == a);
                S.edge[l][a].n[k] = lab.n[k] + (k != a);
            }
        }
        S.center[l].d = (int8_t *)h.center[l];
        for (int k = 0; k < 3; ++k) S.center[l].n[k] = lab.n[k];
    }
    for (int a = 0; a < 3; ++a) {
        S.regular[a].d = (int8_t *)h.regular[a];
        for (int k = 0; k < 3; ++k) S.regular[a].n[k] = h.N[k] + (k == a);
    }
}

