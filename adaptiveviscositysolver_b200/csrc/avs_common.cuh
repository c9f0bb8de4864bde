// avs_common.cuh -- shared device/host definitions of the B200 viscosity-solve library.
//
// Data layout in HBM: the reference's tiled SIM_RawField / SIM_RawIndexField
// pyramid (HDK_OctreeGrid.h:325, HDK_AdaptiveViscosity.cpp:337-350) is flattened into dense,
// x-fastest, level-by-level arrays:
//   label[l]   uint8   cell labels            (Pad >> l)^3
//   face[l][a] int32   face DOF index / label (cell res + 1 on axis a)
//   edge[l][a] int8    edge stress label      (cell res + 1 on the two other axes)
//   center[l]  int8    centre stress label    (cell res)
// Stress stencils (rows of D) are never stored: they are pure functions of these grids and are
// re-evaluated inside the assembly kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// AVS_DEV marks the device code that tests/test_host_assembly.py ALSO compiles for the host (nvcc -DAVS_HOST_TEST on a test harness
// that includes avs_system.cu): the row builder of the assembly and everything it calls then exist as __host__ __device__ functions
// and are run on the CPU against the compiled reference.  In the product build AVS_DEV is plain __device__ (SASS unchanged); the
// library itself has no CPU path.
#ifdef AVS_HOST_TEST
#define AVS_DEV __host__ __device__
#else
#define AVS_DEV __device__
#endif

#define AVS_MAX_LEVELS 10
#define AVS_TILE 16  // UT_VoxelArray tile edge: the reference classifies only inside "occupied" tiles

// OG.h:33-39
enum : uint8_t { L_INACTIVE = 0, L_ACTIVE = 1, L_UP = 2, L_DOWN = 3 };
// UTIL.h:18-21
#define F_FLUID 0
#define F_UNASSIGNED (-1)
#define F_SOLID (-2)
#define F_OUTSIDE (-3)

struct I3 {
    int v[3];
    __host__ __device__ __forceinline__ int &operator[](int a) { return v[a]; }
    __host__ __device__ __forceinline__ int operator[](int a) const { return v[a]; }
};
__host__ __device__ __forceinline__ I3 mk3(int x, int y, int z) {
    I3 r;
    r.v[0] = x; r.v[1] = y; r.v[2] = z;
    return r;
}

// ---- index algebra (HDK_Utilities.h:46-217, HDK_OctreeGrid.h:53-142) as integer device functions
__host__ __device__ __forceinline__ I3 cellToFace(I3 c, int axis, int dir) { if (dir == 1) ++c[axis]; return c; }
__host__ __device__ __forceinline__ I3 cellToCell(I3 c, int axis, int dir) { c[axis] += dir ? 1 : -1; return c; }
__host__ __device__ __forceinline__ I3 cellToEdge(I3 c, int ea, int ei) {
    if (ei & 1) ++c[(ea + 1) % 3];
    if (ei & 2) ++c[(ea + 2) % 3];
    return c;
}
__host__ __device__ __forceinline__ I3 faceToCell(I3 f, int axis, int dir) { if (dir == 0) --f[axis]; return f; }
__host__ __device__ __forceinline__ I3 faceToEdge(I3 f, int fa, int ea, int dir) { if (dir == 1) ++f[3 - fa - ea]; return f; }
__host__ __device__ __forceinline__ I3 edgeToFace(I3 e, int ea, int fa, int dir) { if (dir == 0) --e[3 - fa - ea]; return e; }
__host__ __device__ __forceinline__ I3 edgeToCell(I3 e, int ea, int ci) {
    if (!(ci & 1)) --e[(ea + 1) % 3];
    if (!(ci & 2)) --e[(ea + 2) % 3];
    return e;
}
__host__ __device__ __forceinline__ I3 parentOf(I3 c) { return mk3(c[0] >> 1, c[1] >> 1, c[2] >> 1); }  // indices >= 0 only
__host__ __device__ __forceinline__ I3 childFace(I3 f, int axis, int ch) {
    I3 r = mk3(f[0] * 2, f[1] * 2, f[2] * 2);
    if (ch & 1) ++r[(axis + 1) % 3];
    if (ch & 2) ++r[(axis + 2) % 3];
    return r;
}
__host__ __device__ __forceinline__ I3 childEdge(I3 e, int ea, int ch) {
    I3 r = mk3(e[0] * 2, e[1] * 2, e[2] * 2);
    if (ch > 0) ++r[ea];
    return r;
}
__host__ __device__ __forceinline__ I3 childEdgeInFace(I3 f, int fa, int ea, int ch) {
    I3 r = mk3(f[0] * 2, f[1] * 2, f[2] * 2);
    if (ch == 1) ++r[ea];
    ++r[3 - fa - ea];
    return r;
}

// A dense x-fastest grid with clamped reads (HDKgetFieldValue: out-of-range indices clamp).
template <class T>
struct Grid3 {
    T *d;
    int n[3];
    __host__ __device__ __forceinline__ size_t lin(int x, int y, int z) const {
        return (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z);
    }
    __host__ __device__ __forceinline__ size_t count() const { return (size_t)n[0] * n[1] * n[2]; }
    AVS_DEV __forceinline__ T get(const I3 &c) const {
        int x = min(max(c[0], 0), n[0] - 1), y = min(max(c[1], 0), n[1] - 1), z = min(max(c[2], 0), n[2] - 1);
        return d[lin(x, y, z)];
    }
    AVS_DEV __forceinline__ T &at(const I3 &c) const { return d[lin(c[0], c[1], c[2])]; }
};

// One scalar component of an input field, resident on the device (or constant).
struct DField {
    const float *d;  // null => constant
    int n[3];
    double org[3];
    double dx;
    float constant;
    AVS_DEV __forceinline__ float raw(int x, int y, int z) const {
        if (!d) return constant;
        x = min(max(x, 0), n[0] - 1); y = min(max(y, 0), n[1] - 1); z = min(max(z, 0), n[2] - 1);
        return d[(size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z)];
    }
    // SIM_RawField::getValue(pos): trilinear, clamp-to-edge, fp64, a + t*(b-a) without contraction
    // (this translation unit is compiled with -fmad=false so it rounds exactly like the CPU oracle).
    AVS_DEV double value(const double p[3]) const {
        if (!d) return (double)constant;
        int i0[3], i1[3];
        double t[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double g = (p[a] - org[a]) / dx;
            double hi = (double)(n[a] - 1);
            if (g < 0.0) g = 0.0;
            if (g > hi) g = hi;
            double f = floor(g);
            i0[a] = (int)f;
            i1[a] = min(i0[a] + 1, n[a] - 1);
            t[a] = g - f;
        }
        const size_t sy = (size_t)n[0], sz = (size_t)n[0] * n[1];
        const float *b = d;
        double v000 = b[i0[0] + sy * i0[1] + sz * i0[2]], v100 = b[i1[0] + sy * i0[1] + sz * i0[2]];
        double v010 = b[i0[0] + sy * i1[1] + sz * i0[2]], v110 = b[i1[0] + sy * i1[1] + sz * i0[2]];
        double v001 = b[i0[0] + sy * i0[1] + sz * i1[2]], v101 = b[i1[0] + sy * i0[1] + sz * i1[2]];
        double v011 = b[i0[0] + sy * i1[1] + sz * i1[2]], v111 = b[i1[0] + sy * i1[1] + sz * i1[2]];
        double c00 = v000 + t[0] * (v100 - v000);
        double c10 = v010 + t[0] * (v110 - v010);
        double c01 = v001 + t[0] * (v101 - v001);
        double c11 = v011 + t[0] * (v111 - v011);
        double c0 = c00 + t[1] * (c10 - c00);
        double c1 = c01 + t[1] * (c11 - c01);
        return c0 + t[2] * (c1 - c0);
    }
};

// Everything the labelling / assembly kernels need, passed by value (fits the 4 KB parameter space).
struct DeviceScene {
    int N[3];        // liquid surface resolution
    int Pad[3];      // power-of-two padded resolution (OG.cpp:18-24)
    int levels;      // built levels (after capping)
    double origin[3];
    double dx0;      // float32-rounded voxel size (AV.cpp:242)
    double dt;
    double extrap;   // dx0 * extrapolation (AV.cpp:243)
    int enhanced;
    DField surface, vel[3], faceW[3], viscosity, density, collision, collisionVel[3];
    Grid3<float> centerW, edgeW[3];
    Grid3<uint8_t> label[AVS_MAX_LEVELS];
    Grid3<int32_t> face[AVS_MAX_LEVELS][3];
    Grid3<int8_t> edge[AVS_MAX_LEVELS][3];
    Grid3<int8_t> center[AVS_MAX_LEVELS];
    Grid3<int8_t> regular[3];  // regular-grid face labels (only >=0 / SOLID / UNASSIGNED matter, AV.cpp:2843-2890)

    __host__ __device__ __forceinline__ double levelDx(int level) const { return (double)(float)(dx0 * (double)(1 << level)); }
    AVS_DEV __forceinline__ void centerPos(const I3 &c, int level, double p[3]) const {
        double h = levelDx(level);
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = origin[a] + (c[a] + 0.5) * h;
    }
    AVS_DEV __forceinline__ void facePos(const I3 &f, int axis, int level, double p[3]) const {
        double h = levelDx(level);
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = origin[a] + (f[a] + (a == axis ? 0.0 : 0.5)) * h;
    }
    AVS_DEV __forceinline__ void edgePos(const I3 &e, int axis, int level, double p[3]) const {
        double h = levelDx(level);
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = origin[a] + (e[a] + (a == axis ? 0.5 : 0.0)) * h;
    }
};

// Row key: where a velocity DOF lives. 5 x int32 like the oracle's (level, axis, i, j, k).
struct RowKey {
    int32_t level, axis, i, j, k;
};

#define AVS_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            avs_set_last_error(#expr, _e, __FILE__, __LINE__);                                \
            return AVS_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

void avs_set_last_error(const char *what, cudaError_t e, const char *file, int line);

// AVS_TRACE=1 in the environment: one stderr line per pipeline step (debugging aid; a getenv per call site otherwise)
#include <cstdio>
#include <cstdlib>
#define AVS_TRACE(...)                                                          \
    do {                                                                        \
        static int _on = -1;                                                    \
        if (_on < 0) { const char *_e = getenv("AVS_TRACE"); _on = (_e && _e[0] == '1') ? 1 : 0; } \
        if (_on) { fprintf(stderr, "[avs] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); }   \
    } while (0)
