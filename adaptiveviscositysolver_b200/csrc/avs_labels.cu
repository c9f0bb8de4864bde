// avs_labels.cu -- stages 1-5 of the viscosity solve on the GPU:
//   integration weights  (HDK_AdaptiveViscosity.cpp:712-791)
//   refinement mask + octree label pyramid (HDK_AdaptiveViscosity.cpp:793-871, HDK_OctreeGrid.cpp:4-243)
//   regular-grid face labels (HDK_AdaptiveViscosity.cpp:1087-1165, 1445-1512)
//   octree face / edge / centre labels + DOF numbering (HDK_AdaptiveViscosity.cpp:1167-1443, 1514-1715)
//
// All kernels are streaming passes over dense x-fastest grids (HBM-bound, coalesced along x).
// The reference's serial numbering sweeps (AV.cpp:1563-1591) become count -> scan -> assign over
// 8^3-cell bricks in Morton order, so consecutive rows are spatial neighbours (SpMV gather locality).
// Compiled with -fmad=false: sign tests on interpolated SDF values must round like the CPU oracle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "avs_context.h"

#define LAUNCH_1D(ctx, kernel, count, ...)                                                     \
    do {                                                                                       \
        size_t _n = (size_t)(count);                                                           \
        if (_n > 0) {                                                                          \
            unsigned _b = (unsigned)((_n + 255) / 256);                                        \
            kernel<<<_b, 256, 0, (ctx)->stream>>>(__VA_ARGS__);                                \
            ++(ctx)->launches;                                                                 \
        }                                                                                      \
    } while (0)

// grid-stride counting kernels: 16 elements per thread and trip, at most 8 CTAs per SM
#define LAUNCH_COUNT(ctx, kernel, count, ...)                                                  \
    do {                                                                                       \
        size_t _n = (size_t)(count);                                                           \
        if (_n > 0) {                                                                          \
            unsigned _b = (unsigned)std::min<size_t>((_n / 16 + 255) / 256 + 1, (size_t)(ctx)->numSMs * 8); \
            kernel<<<_b, 256, 0, (ctx)->stream>>>(__VA_ARGS__);                                \
            ++(ctx)->launches;                                                                 \
        }                                                                                      \
    } while (0)

static_assert(sizeof(DeviceScene) <= 4000, "DeviceScene must fit the kernel parameter space");

// Grid kernels are launched as (ceil(nx/128), ceil(ny/CELL_YB), nz) x (128, CELL_YB) threads: x and y from the thread index,
// z from the block index -- no 64-bit div/mod per thread (that cost more than the memory traffic of these streaming passes).
// CELL_YB = grid rows per CTA.  Measured at C3 (512^3) with 4 rows per CTA (128 x 4 threads): the trivial streaming passes gain
// ~10 % (weights 5.70 -> 5.15 ms, octree 2.56 -> 2.19), but the gather-heavy classification passes lose more (octree labels
// 9.07 -> 10.79 ms): one row per CTA stays.  Four samples per THREAD with the loads hoisted in front of the decisions (round 2, same
// decisions, bit-identical output) lost as well: octree labels 9.24 -> 11.04 ms, regular labels 3.95 -> 3.79 ms at C3 -- these
// kernels are bound by their integer index arithmetic (~270 instructions per face), not by load latency; not kept.
#define CELL_YB 1
#define LAUNCH_3D(ctx, kernel, n3, ...)                                                        \
    do {                                                                                       \
        if ((n3)[0] > 0 && (n3)[1] > 0 && (n3)[2] > 0) {                                        \
            dim3 _g((unsigned)(((n3)[0] + 127) / 128), (unsigned)(((n3)[1] + CELL_YB - 1) / CELL_YB), (unsigned)(((n3)[2] + ZREP - 1) / ZREP));    \
            kernel<<<_g, dim3(128, CELL_YB), 0, (ctx)->stream>>>(__VA_ARGS__);                  \
            ++(ctx)->launches;                                                                 \
        }                                                                                      \
    } while (0)
#define ZREP 1   // grid cells per thread along z (8 was measured slower: serialises the per-cell loads of a thread)
template <class F>
__device__ __forceinline__ void forCells(const int n[3], F f) {
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)(blockIdx.y * blockDim.y + threadIdx.y);
    const bool inx = x < n[0] && y < n[1];
    for (int r = 0; r < ZREP; ++r) {
        const int z = (int)blockIdx.z * ZREP + r;
        if (z >= n[2]) return;  // uniform over the CTA
        f(x, y, z, (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z), inx);  // inx == false: lane is outside the row
    }
}

// z-window variant: blockIdx.z counts planes from zOff (multi-GPU z-slabs of the regular grid)
template <class F>
__device__ __forceinline__ void forCellsZ(const int n[3], int zOff, F f) {
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)(blockIdx.y * blockDim.y + threadIdx.y), z = (int)blockIdx.z + zOff;
    if (z >= n[2]) return;
    f(x, y, z, (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z), x < n[0] && y < n[1]);
}

// ------------------------------------------------------------------------------------------------
// Stage 1: computeSDFWeightsSampled restated: fraction of the n^3 sub-samples at
// offsets ((k+1/2)/n - 1/2) dx whose interpolated sdf (minus dilate) is negative.
// Early-out: the trilinear interpolant is a convex combination of the voxels it touches, so when
// every voxel under the sample's box has the same sign the count is n^3 or 0 without sampling.
AVS_DEV void sdfWeightSample(Grid3<float> w, const DField &sdf, double off0, double off1, double off2, double o0,
                                double o1, double o2, double dx0, int n, double dilate, const uint8_t *signClass,
                                int x, int y, int z, size_t idx) {
    double c[3] = {o0 + (x + off0) * dx0, o1 + (y + off1) * dx0, o2 + (z + off2) * dx0};
    if (!sdf.d) {
        w.d[idx] = ((double)sdf.constant - dilate < 0.0) ? 1.f : 0.f;
        return;
    }
    // Aligned sdf grid (the liquid surface always is): the voxels under this sample's box are a subset of
    // the 3x3x3 neighbourhood of cell (x,y,z) clamped into the grid, whose sign class was precomputed.
    if (signClass) {
        int cx = min(x, sdf.n[0] - 1), cy = min(y, sdf.n[1] - 1), cz = min(z, sdf.n[2] - 1);
        uint8_t cls = signClass[(size_t)cx + (size_t)sdf.n[0] * ((size_t)cy + (size_t)sdf.n[1] * cz)];
        if (cls == 0) { w.d[idx] = 1.f; return; }
        if (cls == 1) { w.d[idx] = 0.f; return; }
    }
    const double inv = 1.0 / (double)n;
    const double h = (0.5 - 0.5 * inv) * dx0;
    int lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double gl = (c[a] - h - sdf.org[a]) / sdf.dx - 1e-9;
        double gh = (c[a] + h - sdf.org[a]) / sdf.dx + 1e-9;
        double top = (double)(sdf.n[a] - 1);
        gl = fmin(fmax(gl, 0.0), top);
        gh = fmin(fmax(gh, 0.0), top);
        lo[a] = (int)floor(gl);
        hi[a] = min((int)floor(gh) + 1, sdf.n[a] - 1);
    }
    bool allNeg = true, allPos = true;
    const float dl = (float)dilate;  // exact for dilate == 0; conservative test below otherwise
    for (int kz = lo[2]; kz <= hi[2]; ++kz)
        for (int ky = lo[1]; ky <= hi[1]; ++ky)
            for (int kx = lo[0]; kx <= hi[0]; ++kx) {
                double v = (double)sdf.d[(size_t)kx + (size_t)sdf.n[0] * ((size_t)ky + (size_t)sdf.n[1] * kz)] - dilate;
                allNeg = allNeg && (v < 0.0);
                allPos = allPos && (v >= 0.0);
            }
    (void)dl;
    if (allNeg) { w.d[idx] = 1.f; return; }
    if (allPos) { w.d[idx] = 0.f; return; }
    int count = 0;
    if (n <= 4) {
        // The sub-sample lattice is separable, so the interpolation indices/weights of DField::value() take only
        // n distinct values per axis: evaluate those once (same expressions, same rounding) and reuse them.
        int i0[3][4], i1[3][4];
        double tt[3][4];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int s = 0; s < n; ++s) {
                double pa = c[a] + ((s + 0.5) * inv - 0.5) * dx0;
                double g = (pa - sdf.org[a]) / sdf.dx;
                double top = (double)(sdf.n[a] - 1);
                if (g < 0.0) g = 0.0;
                if (g > top) g = top;
                double f = floor(g);
                i0[a][s] = (int)f;
                i1[a][s] = min((int)f + 1, sdf.n[a] - 1);
                tt[a][s] = g - f;
            }
        const size_t sy_ = (size_t)sdf.n[0], sz_ = (size_t)sdf.n[0] * sdf.n[1];
        const float *b = sdf.d;
        for (int sz = 0; sz < n; ++sz)
            for (int sy = 0; sy < n; ++sy)
                for (int sx = 0; sx < n; ++sx) {
                    const int x0 = i0[0][sx], x1 = i1[0][sx], y0 = i0[1][sy], y1 = i1[1][sy], z0 = i0[2][sz], z1 = i1[2][sz];
                    const double tx = tt[0][sx], ty = tt[1][sy], tz = tt[2][sz];
                    double v000 = b[x0 + sy_ * y0 + sz_ * z0], v100 = b[x1 + sy_ * y0 + sz_ * z0];
                    double v010 = b[x0 + sy_ * y1 + sz_ * z0], v110 = b[x1 + sy_ * y1 + sz_ * z0];
                    double v001 = b[x0 + sy_ * y0 + sz_ * z1], v101 = b[x1 + sy_ * y0 + sz_ * z1];
                    double v011 = b[x0 + sy_ * y1 + sz_ * z1], v111 = b[x1 + sy_ * y1 + sz_ * z1];
                    double c00 = v000 + tx * (v100 - v000);
                    double c10 = v010 + tx * (v110 - v010);
                    double c01 = v001 + tx * (v101 - v001);
                    double c11 = v011 + tx * (v111 - v011);
                    double c0 = c00 + ty * (c10 - c00);
                    double c1 = c01 + ty * (c11 - c01);
                    if ((c0 + tz * (c1 - c0)) - dilate < 0.0) ++count;
                }
    } else {
        for (int sz = 0; sz < n; ++sz)
            for (int sy = 0; sy < n; ++sy)
                for (int sx = 0; sx < n; ++sx) {
                    double p[3] = {c[0] + ((sx + 0.5) * inv - 0.5) * dx0, c[1] + ((sy + 0.5) * inv - 0.5) * dx0,
                                   c[2] + ((sz + 0.5) * inv - 0.5) * dx0};
                    if (sdf.value(p) - dilate < 0.0) ++count;
                }
    }
    const double total = (double)n * n * n;
    w.d[idx] = (float)((double)count / total);
}

__global__ void k_sdf_weights(Grid3<float> w, DField sdf, double off0, double off1, double off2, double o0,
                              double o1, double o2, double dx0, int n, double dilate, const uint8_t *signClass,
                              const uint32_t *list, const unsigned long long *listCount) {
    if (list) {  // list-driven: one thread per band sample
        unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= *listCount) return;
        size_t idx = list[t];
        int x = (int)(idx % w.n[0]);
        int y = (int)((idx / w.n[0]) % w.n[1]);
        int z = (int)(idx / ((size_t)w.n[0] * w.n[1]));
        sdfWeightSample(w, sdf, off0, off1, off2, o0, o1, o2, dx0, n, dilate, nullptr, x, y, z, idx);
        return;
    }
    forCells(w.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        sdfWeightSample(w, sdf, off0, off1, off2, o0, o1, o2, dx0, n, dilate, signClass, x, y, z, idx);
    });
}

// ---- sign classes of the 3x3x3 voxel neighbourhood (clamped), separable: x, then y, then z ------------
// class 0: every voxel < 0, 1: every voxel >= 0, 2: mixed.  One byte per voxel of the sdf grid.
__device__ __forceinline__ uint8_t combine3(uint8_t a, uint8_t b, uint8_t c) { return (a == b && b == c) ? a : (uint8_t)2; }
__global__ void k_sign_x(const float *sdf, int nx, int ny, int nz, uint8_t *out) {
    const int n3[3] = {nx, ny, nz};
    forCells(n3, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
    size_t row = idx - x;
    uint8_t a = sdf[row + max(x - 1, 0)] < 0.f ? 0 : 1, b = sdf[idx] < 0.f ? 0 : 1, c = sdf[row + min(x + 1, nx - 1)] < 0.f ? 0 : 1;
    out[idx] = combine3(a, b, c);
    });
}
__global__ void k_sign_axis(const uint8_t *in, int nx, int ny, int nz, int axis, uint8_t *out) {
    const int n3[3] = {nx, ny, nz};
    forCells(n3, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
    size_t stride = axis == 1 ? (size_t)nx : (size_t)nx * ny;
    int pos = axis == 1 ? y : z, top = axis == 1 ? ny - 1 : nz - 1;
    uint8_t a = in[pos > 0 ? idx - stride : idx], b = in[idx], c = in[pos < top ? idx + stride : idx];
    out[idx] = combine3(a, b, c);
    });
}

// ---- the same three passes, four cells per thread (rows whose length is a multiple of 4: every row then starts 4-byte / 16-byte
// aligned).  One cell per thread keeps a single 1- or 4-byte load in flight per thread and the passes end up latency-bound at
// ~1 TB/s; with 4 cells per thread the loads of a thread are independent, 4x wider and issued together.
__device__ __forceinline__ unsigned combine3x4(unsigned a, unsigned b, unsigned c) {   // combine3 on four packed class bytes
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned x = (a >> (8 * k)) & 0xffu, y = (b >> (8 * k)) & 0xffu, z = (c >> (8 * k)) & 0xffu;
        r |= ((x == y && y == z) ? x : 2u) << (8 * k);
    }
    return r;
}
// Also findOccupiedRegularVelocityTiles (AV.cpp:886-943; k_mark_surface_tiles below is the stand-alone form): this pass reads the
// whole surface field anyway, so it marks the face tiles of every cell with sdf < 2 dx on the way (t0.d == nullptr: no marking).
__global__ void k_sign_x4(const float *__restrict__ sdf, int nx, int ny, int nz, uint8_t *__restrict__ out, double twoDx, Grid3<uint8_t> t0,
                          Grid3<uint8_t> t1, Grid3<uint8_t> t2) {
    const int x4 = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z;
    const int x = 4 * x4;
    if (x >= nx) return;
    const size_t row = (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z);
    const float4 v = *(const float4 *)(sdf + row + x);
    if (t0.d) {
        const bool h3 = (double)v.w < twoDx;
        if (h3 || (double)v.x < twoDx || (double)v.y < twoDx || (double)v.z < twoDx) {
            const int tx = x / AVS_TILE, ty = y / AVS_TILE, tz = z / AVS_TILE;   // x .. x+3 lie in one tile, and so do their +1 faces but the last
            t0.d[t0.lin(tx, ty, tz)] = 1;
            t1.d[t1.lin(tx, ty, tz)] = 1; t1.d[t1.lin(tx, (y + 1) / AVS_TILE, tz)] = 1;
            t2.d[t2.lin(tx, ty, tz)] = 1; t2.d[t2.lin(tx, ty, (z + 1) / AVS_TILE)] = 1;
            if (h3) t0.d[t0.lin((x + 4) / AVS_TILE, ty, tz)] = 1;
        }
    }
    const float l = sdf[row + max(x - 1, 0)], r = sdf[row + min(x + 4, nx - 1)];
    const unsigned s[6] = {l < 0.f ? 0u : 1u, v.x < 0.f ? 0u : 1u, v.y < 0.f ? 0u : 1u, v.z < 0.f ? 0u : 1u, v.w < 0.f ? 0u : 1u, r < 0.f ? 0u : 1u};
    unsigned o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) o |= ((s[k] == s[k + 1] && s[k + 1] == s[k + 2]) ? s[k + 1] : 2u) << (8 * k);
    *(unsigned *)(out + row + x) = o;
}
__global__ void k_sign_axis4(const uint8_t *__restrict__ in, int nx, int ny, int nz, int axis, uint8_t *__restrict__ out) {
    const int x4 = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z;
    const int x = 4 * x4;
    if (x >= nx) return;
    const size_t idx = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z);
    const size_t stride = axis == 1 ? (size_t)nx : (size_t)nx * ny;
    const int pos = axis == 1 ? y : z, top = axis == 1 ? ny - 1 : nz - 1;
    const unsigned a = *(const unsigned *)(in + (pos > 0 ? idx - stride : idx)), b = *(const unsigned *)(in + idx),
                   c = *(const unsigned *)(in + (pos < top ? idx + stride : idx));
    *(unsigned *)(out + idx) = combine3x4(a, b, c);
}

// Light pass for ALL FOUR weight grids at once (centre + the three edge grids; aligned sdf grid only): one thread per index
// (x, y, z) of the (N+1)^3 corner lattice, four x positions per thread (x, x+128, x+256, x+384: every load / store instruction of a
// warp stays one contiguous segment).  The sign class of the clamped cell is read ONCE and decides the weight of the centre sample
// and of the three edge samples that carry this index: 1 (all voxels under the sample negative), 0 (all non-negative), or the
// sample goes on its grid's band list for k_sdf_weights.  Replaces four dense passes that each re-read the class array.
struct WeightGrids {
    Grid3<float> g[4];          // centre, edge x, edge y, edge z
    uint32_t *list[4];
    unsigned long long cap[4];  // entries a list can hold; counts beyond it are detected on the host (dense fallback)
};
__global__ void k_weights_classify4(const __grid_constant__ WeightGrids W, int sn0, int sn1, int sn2, const uint8_t *__restrict__ signClass,
                                    unsigned long long *listCount) {
    const int y = (int)blockIdx.y, z = (int)blockIdx.z;
    const int cy = min(y, sn1 - 1), cz = min(z, sn2 - 1);
    const size_t crow = (size_t)sn0 * ((size_t)cy + (size_t)sn1 * (size_t)cz);
    const int xb = (int)blockIdx.x * 512 + (int)threadIdx.x;
    uint8_t cls[4];
    bool band4 = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = xb + 128 * j;
        cls[j] = (x <= sn0) ? signClass[crow + min(x, sn0 - 1)] : (uint8_t)0;
        band4 = band4 || cls[j] == 2;
    }
    // 98 % of the warps see no band sample at all: they only store (no ballots, no list logic)
    const bool warpHasBand = __any_sync(0xffffffffu, band4);
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
        const Grid3<float> &g = W.g[gi];
        if (!(y < g.n[1] && z < g.n[2])) continue;   // uniform over the CTA
        float *rowp = g.d + (size_t)g.n[0] * ((size_t)y + (size_t)g.n[1] * (size_t)z);
        const unsigned rowIdx = (unsigned)((size_t)g.n[0] * ((size_t)y + (size_t)g.n[1] * (size_t)z));   // lattice < 2^32 (checked on the host)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = xb + 128 * j;
            const bool in = x < g.n[0];
            if (in && cls[j] != 2) rowp[x] = cls[j] == 0 ? 1.f : 0.f;
            if (!warpHasBand) continue;
            const bool band = in && cls[j] == 2;
            const unsigned m = __ballot_sync(0xffffffffu, band);   // warp-aggregated append
            if (m) {
                const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(listCount + gi, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
                if (band && slot < W.cap[gi]) W.list[gi][slot] = rowIdx + (unsigned)x;
            }
        }
    }
}

// setScaleDivideThreshold(1, nullptr, &b, 0): a /= b where b > 0 (AV.cpp:781-789)
__global__ void k_divide_where_positive(float *a, const float *b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && b[i] > 0.f) a[i] = a[i] / b[i];
}

// dense supersampling of one weight grid (the collision weights of doApplySolidWeights: any grid, dilated, no sign classes)
static int weightsFor(AvsContext *c, DevBuf &buf, Grid3<float> &g, const DField &sdf, const double off[3],
                      const int res[3], int n, double dilate) {
    g.n[0] = res[0]; g.n[1] = res[1]; g.n[2] = res[2];
    if (buf.reserve(g.count() * sizeof(float))) return AVS_ERR_ALLOC;
    g.d = buf.as<float>();
    LAUNCH_3D(c, k_sdf_weights, g.n, g, sdf, off[0], off[1], off[2], c->S.origin[0], c->S.origin[1],
              c->S.origin[2], c->S.dx0, n, dilate, nullptr, nullptr, nullptr);
    return AVS_OK;
}

static int avs_face_tile_maps(AvsContext *c, Grid3<uint8_t> t[3]);
int avs_stage_weights(AvsContext *c, const AvsParams *p) {
    DeviceScene &S = c->S;
    c->surfaceTilesMarked = false;
    const int n = p->number_super_samples;
    const double offC[3] = {0.5, 0.5, 0.5};
    // sign classes of the surface SDF (only valid as a shortcut when the field sits on the scene grid with dilate = 0,
    // which avs_stage_upload has validated for `surface`)
    const uint8_t *cls = nullptr;
    const size_t cells = (size_t)S.N[0] * S.N[1] * S.N[2];
    if (S.surface.d && n >= 1) {
        if (c->signA.reserve(cells) || c->signB.reserve(cells)) return AVS_ERR_ALLOC;
        if (S.N[0] % 4 == 0) {   // rows start 16-byte aligned (cudaMalloc base, row length a multiple of 4 floats): 4 cells per thread
            const dim3 g4((unsigned)((S.N[0] / 4 + 127) / 128), (unsigned)S.N[1], (unsigned)S.N[2]);
            // the regular-grid face tile maps of stage 4 are filled here, by the pass that reads the surface anyway
            Grid3<uint8_t> ft[3];
            int rct = avs_face_tile_maps(c, ft);
            if (rct) return rct;
            c->surfaceTilesMarked = true;
            k_sign_x4<<<g4, 128, 0, c->stream>>>(S.surface.d, S.N[0], S.N[1], S.N[2], c->signA.as<uint8_t>(), 2.0 * S.dx0, ft[0], ft[1], ft[2]);
            k_sign_axis4<<<g4, 128, 0, c->stream>>>(c->signA.as<uint8_t>(), S.N[0], S.N[1], S.N[2], 1, c->signB.as<uint8_t>());
            k_sign_axis4<<<g4, 128, 0, c->stream>>>(c->signB.as<uint8_t>(), S.N[0], S.N[1], S.N[2], 2, c->signA.as<uint8_t>());
            c->launches += 3;
        } else {
            LAUNCH_3D(c, k_sign_x, S.N, S.surface.d, S.N[0], S.N[1], S.N[2], c->signA.as<uint8_t>());
            LAUNCH_3D(c, k_sign_axis, S.N, c->signA.as<uint8_t>(), S.N[0], S.N[1], S.N[2], 1, c->signB.as<uint8_t>());
            LAUNCH_3D(c, k_sign_axis, S.N, c->signB.as<uint8_t>(), S.N[0], S.N[1], S.N[2], 2, c->signA.as<uint8_t>());
        }
        cls = c->signA.as<uint8_t>();
    }
    // the four grids: centre samples, and per axis a the a-directed edges (centred along a only, HDK_Utilities.h:13-15)
    Grid3<float> *grids[4] = {&S.centerW, &S.edgeW[0], &S.edgeW[1], &S.edgeW[2]};
    DevBuf *bufs[4] = {&c->centerW, &c->edgeW[0], &c->edgeW[1], &c->edgeW[2]};
    double offs[4][3] = {{0.5, 0.5, 0.5}, {0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0.5}};
    for (int gi = 0; gi < 4; ++gi) {
        Grid3<float> &g = *grids[gi];
        for (int k = 0; k < 3; ++k) g.n[k] = S.N[k] + ((gi == 0 || k == gi - 1) ? 0 : 1);
        if (bufs[gi]->reserve(g.count() * sizeof(float))) return AVS_ERR_ALLOC;
        g.d = bufs[gi]->as<float>();
    }
    const size_t lattice = (size_t)(S.N[0] + 1) * (S.N[1] + 1) * (S.N[2] + 1);
    if (cls && lattice < 0xffffffffull) {
        // one fused light pass over the corner lattice writes the 0 / 1 weights of all four grids and lists their band samples
        WeightGrids W;
        const size_t cap = lattice / 4 + 4096;   // a band is a thin shell: a list that overflows falls back to the dense sampler
        if (c->bandList.reserve(4 * cap * sizeof(uint32_t))) return AVS_ERR_ALLOC;
        for (int gi = 0; gi < 4; ++gi) {
            W.g[gi] = *grids[gi];
            W.list[gi] = c->bandList.as<uint32_t>() + (size_t)gi * cap;
            W.cap[gi] = cap;
        }
        unsigned long long *cnt = c->counters.as<unsigned long long>() + 28;   // slots 28..31
        AVS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned long long), c->stream));
        const dim3 grid((unsigned)((S.N[0] + 1 + 511) / 512), (unsigned)(S.N[1] + 1), (unsigned)(S.N[2] + 1));
        k_weights_classify4<<<grid, 128, 0, c->stream>>>(W, S.N[0], S.N[1], S.N[2], cls, cnt);
        ++c->launches;
        unsigned long long h[4] = {0, 0, 0, 0};
        AVS_CUDA_CHECK(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int gi = 0; gi < 4; ++gi) {
            Grid3<float> &g = *grids[gi];
            if (h[gi] > cap) {   // not a thin band (noise field): every sample decides for itself, class shortcut included
                LAUNCH_3D(c, k_sdf_weights, g.n, g, S.surface, offs[gi][0], offs[gi][1], offs[gi][2], S.origin[0], S.origin[1], S.origin[2], S.dx0, n,
                          0.0, cls, nullptr, nullptr);
            } else if (h[gi] > 0) {
                k_sdf_weights<<<(unsigned)((h[gi] + 127) / 128), 128, 0, c->stream>>>(g, S.surface, offs[gi][0], offs[gi][1], offs[gi][2], S.origin[0],
                                                                                  S.origin[1], S.origin[2], S.dx0, n, 0.0, nullptr, W.list[gi], cnt + gi);
                ++c->launches;
            }
        }
    } else {
        for (int gi = 0; gi < 4; ++gi) {
            Grid3<float> &g = *grids[gi];
            LAUNCH_3D(c, k_sdf_weights, g.n, g, S.surface, offs[gi][0], offs[gi][1], offs[gi][2], S.origin[0], S.origin[1], S.origin[2], S.dx0, n, 0.0,
                      cls, nullptr, nullptr);
        }
    }
    int rc = AVS_OK;
    if (p->do_apply_solid_weights) {  // AV.cpp:772-790
        DevBuf &tmp = c->solidW;
        Grid3<float> g;
        rc = weightsFor(c, tmp, g, S.collision, offC, S.N, n, -S.extrap);
        if (rc) return rc;
        LAUNCH_1D(c, k_divide_where_positive, g.count(), S.centerW.d, g.d, g.count());
        for (int a = 0; a < 3; ++a) {
            double off[3] = {0, 0, 0};
            off[a] = 0.5;
            rc = weightsFor(c, tmp, g, S.collision, off, S.edgeW[a].n, n, -S.extrap);
            if (rc) return rc;
            LAUNCH_1D(c, k_divide_where_positive, g.count(), S.edgeW[a].d, g.d, g.count());
        }
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Stage 2+3a: refinement mask (AV.cpp:839-860) fused with setBaseGridLabels (OG.cpp:310-392):
// mask 0 -> ACTIVE, < 0 -> UP, > 0 (or outside the un-padded grid) -> INACTIVE.
AVS_DEV __forceinline__ uint8_t baseLabel(const DeviceScene &S, double inner, double outer, int x, int y, int z) {
    uint8_t out = L_INACTIVE;
    if (x < S.N[0] && y < S.N[1] && z < S.N[2]) {
        double sdf = (double)S.surface.raw(x, y, z);
        if (sdf > 0 && sdf < outer) out = L_ACTIVE;
        else if (sdf <= 0.) {
            if (sdf > -inner) out = L_ACTIVE;
            else {
                double p[3];
                S.centerPos(mk3(x, y, z), 0, p);
                out = (S.collision.value(p) > (-inner - S.extrap)) ? L_ACTIVE : L_UP;
            }
        }
    }
    return out;
}
__global__ void k_base_labels(const __grid_constant__ DeviceScene S, double inner, double outer) {
    const Grid3<uint8_t> lab = S.label[0];
    forCells(lab.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        lab.d[idx] = baseLabel(S, inner, outer, x, y, z);
    });
}

// pass 1, setActiveCellsAndParentList (OG.cpp:394-565): one thread per 2x2x2 sibling block.
AVS_DEV __forceinline__ void octreePass1(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int px, int py, int pz, size_t idx) {
    uint8_t v[8];
    bool any = false;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        v[ch] = cur.d[cur.lin(2 * px + (ch & 1), 2 * py + ((ch >> 1) & 1), 2 * pz + (ch >> 2))];
        any = any || (v[ch] == L_ACTIVE);
    }
    if (!any) return;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        if (v[ch] == L_UP) cur.d[cur.lin(2 * px + (ch & 1), 2 * py + ((ch >> 1) & 1), 2 * pz + (ch >> 2))] = L_ACTIVE;
    par.d[idx] = L_DOWN;
}
__global__ void k_octree_pass1(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    forCells(par.n, [&](int px, int py, int pz, size_t idx, bool ok) {
        if (!ok) return;
        octreePass1(cur, par, px, py, pz, idx);
    });
}

// pass 2, setFaceGrading (OG.cpp:656-754). The reference applies its DOWN list before its ACTIVE
// list (OG.cpp:145, 162); two kernels keep that order.
AVS_DEV __forceinline__ void octreePass2Down(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int x, int y, int z, size_t idx) {
    if (cur.d[idx] != L_DOWN) return;
    par.d[par.lin(x >> 1, y >> 1, z >> 1)] = L_DOWN;
}
AVS_DEV __forceinline__ void octreePass2Active(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int x, int y, int z, size_t idx) {
    if (cur.d[idx] != L_ACTIVE) return;
    I3 c = mk3(x, y, z);
#pragma unroll
    for (int axis = 0; axis < 3; ++axis)
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            I3 a = cellToCell(c, axis, dir);
            if (a[axis] < 0 || a[axis] >= cur.n[axis]) continue;
            if (cur.at(a) == L_UP) par.d[par.lin(a[0] >> 1, a[1] >> 1, a[2] >> 1)] = L_ACTIVE;
        }
}
// pass 3, setParentsUp (OG.cpp:756-840)
AVS_DEV __forceinline__ void octreePass3(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int x, int y, int z, size_t idx) {
    if (cur.d[idx] != L_UP) return;
    size_t pi = par.lin(x >> 1, y >> 1, z >> 1);
    if (par.d[pi] == L_INACTIVE) par.d[pi] = L_UP;
}
__global__ void k_octree_pass2_down(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    forCells(cur.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        octreePass2Down(cur, par, x, y, z, idx);
    });
}
__global__ void k_octree_pass2_active(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    forCells(cur.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        octreePass2Active(cur, par, x, y, z, idx);
    });
}
__global__ void k_octree_pass3(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    forCells(cur.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        octreePass3(cur, par, x, y, z, idx);
    });
}
// ---- the same passes with 16 cells per thread (rows whose length is a multiple of 16: one 128-bit load per thread) ------------
// One byte per thread keeps these sweeps at ~0.25 TB/s.  Only a band of cells is ACTIVE and the UP cells of one 16-cell run share
// 8 parents, so a thread that loads 16 labels at once usually has nothing (pass 2) or eight byte-checks (pass 3) left to do.
AVS_DEV __forceinline__ void octreePass2Active16(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int x0, int y, int z) {
    const size_t row = (size_t)cur.n[0] * ((size_t)y + (size_t)cur.n[1] * (size_t)z);
    __align__(16) uint8_t v[16];
    *(uint4 *)v = *(const uint4 *)(cur.d + row + x0);
    bool any = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) any = any || v[k] == L_ACTIVE;
    if (!any) return;
    for (int k = 0; k < 16; ++k) {
        if (v[k] != L_ACTIVE) continue;
        const I3 c = mk3(x0 + k, y, z);
#pragma unroll
        for (int axis = 0; axis < 3; ++axis)
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
                I3 a = cellToCell(c, axis, dir);
                if (a[axis] < 0 || a[axis] >= cur.n[axis]) continue;
                if (cur.at(a) == L_UP) par.d[par.lin(a[0] >> 1, a[1] >> 1, a[2] >> 1)] = L_ACTIVE;
            }
    }
}
__global__ void k_octree_pass2_active16(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    const int x0 = 16 * (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z;
    if (x0 >= cur.n[0]) return;
    octreePass2Active16(cur, par, x0, y, z);
}
AVS_DEV __forceinline__ void octreePass3_16(const Grid3<uint8_t> &cur, const Grid3<uint8_t> &par, int x0, int y, int z) {
    const size_t row = (size_t)cur.n[0] * ((size_t)y + (size_t)cur.n[1] * (size_t)z);
    __align__(16) uint8_t v[16];
    *(uint4 *)v = *(const uint4 *)(cur.d + row + x0);
    const size_t prow = par.lin(x0 >> 1, y >> 1, z >> 1);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (v[2 * k] == L_UP || v[2 * k + 1] == L_UP) {
            if (par.d[prow + k] == L_INACTIVE) par.d[prow + k] = L_UP;
        }
}
__global__ void k_octree_pass3_16(Grid3<uint8_t> cur, Grid3<uint8_t> par) {
    const int x0 = 16 * (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z;
    if (x0 >= cur.n[0]) return;
    octreePass3_16(cur, par, x0, y, z);
}
// k_base_labels with 4 cells per thread (padded rows are a multiple of 4): one 32-bit store per thread
__global__ void k_base_labels4(const __grid_constant__ DeviceScene S, double inner, double outer) {
    const Grid3<uint8_t> lab = S.label[0];
    const int x0 = 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z;
    if (x0 >= lab.n[0]) return;
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) packed |= (unsigned)baseLabel(S, inner, outer, x0 + k, y, z) << (8 * k);
    *(unsigned *)(lab.d + lab.lin(x0, y, z)) = packed;
}

// setTopLevel (OG.cpp:843-875)
__global__ void k_octree_top(Grid3<uint8_t> g) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < g.count() && g.d[idx] == L_UP) g.d[idx] = L_ACTIVE;
}

// one-byte element types only: 16 elements per thread and trip (128-bit loads), grid-stride, one atomic per CTA
template <class T>
__global__ void k_count_equal(const T *d, size_t n, T value, unsigned long long *counter) {
    static_assert(sizeof(T) == 1, "byte grids");
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    const unsigned char key = (unsigned char)value;
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)d & 15)) & 15));   // elements before the first 16-byte boundary
    int cnt = 0;
    for (size_t i = tid; i < head; i += nthreads) cnt += ((const unsigned char *)d)[i] == key;
    const uint4 *v = (const uint4 *)((const unsigned char *)d + head);
    const size_t nvec = (n - head) / 16;
    for (size_t i = tid; i < nvec; i += nthreads) {
        const uint4 q = v[i];
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) cnt += ((w[j] >> (8 * b)) & 0xffu) == key;
    }
    for (size_t i = head + nvec * 16 + tid; i < n; i += nthreads) cnt += ((const unsigned char *)d)[i] == key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        if (t) atomicAdd(counter, (unsigned long long)t);
    }
}
template <class T>
__global__ void k_count_nonneg(const T *d, size_t n, unsigned long long *counter) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = idx < n && d[idx] >= 0;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        if (t) atomicAdd(counter, (unsigned long long)t);
    }
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}

int avs_stage_octree(AvsContext *c, const AvsParams *p) {
    DeviceScene &S = c->S;
    for (int a = 0; a < 3; ++a) {  // OG.cpp:18-24
        int pad = 1;
        while (pad < S.N[a]) pad <<= 1;
        S.Pad[a] = pad;
    }
    int L = p->octree_levels;  // OG.cpp:32-40
    for (int a = 0; a < 3; ++a) L = std::min(L, ilog2(S.Pad[a]));
    if (L < 1) L = 1;
    if (L > AVS_MAX_LEVELS) L = AVS_MAX_LEVELS;
    c->levelsAllocated = L;
    for (int l = 0; l < L; ++l) {
        Grid3<uint8_t> &g = S.label[l];
        for (int a = 0; a < 3; ++a) g.n[a] = S.Pad[a] >> l;
        if (c->label[l].reserve(g.count())) return AVS_ERR_ALLOC;
        g.d = c->label[l].as<uint8_t>();
        if (l > 0) AVS_CUDA_CHECK(cudaMemsetAsync(g.d, L_INACTIVE, g.count(), c->stream));
    }
    const double fineVoxelWidth = std::max(2.0, (double)p->fine_bandwidth);  // AV.cpp:259
    const double inner = S.dx0 * fineVoxelWidth, outer = 3.0 * S.dx0;         // AV.cpp:261-262
    if (S.label[0].n[0] % 4 == 0) {
        const dim3 g((unsigned)((S.label[0].n[0] / 4 + 127) / 128), (unsigned)S.label[0].n[1], (unsigned)S.label[0].n[2]);
        k_base_labels4<<<g, 128, 0, c->stream>>>(S, inner, outer);
        ++c->launches;
    } else LAUNCH_3D(c, k_base_labels, S.label[0].n, S, inner, outer);
    for (int l = 0; l < L - 1; ++l) {
        Grid3<uint8_t> cur = S.label[l], par = S.label[l + 1];
        LAUNCH_3D(c, k_octree_pass1, par.n, cur, par);
        if (l > 0) LAUNCH_3D(c, k_octree_pass2_down, cur.n, cur, par);
        if (cur.n[0] % 16 == 0) {   // 16 cells per thread
            const dim3 g((unsigned)((cur.n[0] / 16 + 63) / 64), (unsigned)cur.n[1], (unsigned)cur.n[2]);
            k_octree_pass2_active16<<<g, 64, 0, c->stream>>>(cur, par);
            k_octree_pass3_16<<<g, 64, 0, c->stream>>>(cur, par);
            c->launches += 2;
        } else {
            LAUNCH_3D(c, k_octree_pass2_active, cur.n, cur, par);
            LAUNCH_3D(c, k_octree_pass3, cur.n, cur, par);
        }
    }
    LAUNCH_1D(c, k_octree_top, S.label[L - 1].count(), S.label[L - 1]);
    // level capping (OG.cpp:198-211): first level without an ACTIVE cell
    if (c->counters.reserve(64 * sizeof(unsigned long long))) return AVS_ERR_ALLOC;
    unsigned long long *cnt = c->counters.as<unsigned long long>();
    AVS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 64 * sizeof(unsigned long long), c->stream));
    for (int l = 0; l < L; ++l)
        LAUNCH_COUNT(c, k_count_equal<uint8_t>, S.label[l].count(), S.label[l].d, S.label[l].count(), (uint8_t)L_ACTIVE, cnt + l);
    unsigned long long h[AVS_MAX_LEVELS];
    AVS_CUDA_CHECK(cudaMemcpyAsync(h, cnt, L * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    int capped = 0;
    for (; capped < L; ++capped)
        if (h[capped] == 0) break;
    S.levels = capped;
    return AVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Stage 4/5: labelling.  The reference only classifies inside 16^3 tiles it found "occupied"
// (findOccupied*Tiles, AV.cpp:886-1057); everything else keeps UNASSIGNED.  Tile maps are uint8.
// warp-aggregated count of lanes whose predicate holds (every lane of the warp must call it)
__device__ __forceinline__ void countWarp(bool hit, unsigned long long *counter) {
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}
AVS_DEV __forceinline__ void markTile(const Grid3<uint8_t> &t, const I3 &c) {
    t.d[t.lin(c[0] / AVS_TILE, c[1] / AVS_TILE, c[2] / AVS_TILE)] = 1;
}
AVS_DEV __forceinline__ bool tileOccupied(const Grid3<uint8_t> &t, int x, int y, int z) {
    return t.d[t.lin(x / AVS_TILE, y / AVS_TILE, z / AVS_TILE)] != 0;
}

// findOccupiedRegularVelocityTiles (AV.cpp:886-943): both faces of every cell with sdf < 2 dx.  (The six byte stores of a warp's
// lanes hit two or three tile bytes per map and are merged by the hardware: electing one lane per 16-cell segment to do them
// measured no faster -- regular labels 3.99 -> 4.33 ms at C3.)
__global__ void k_mark_surface_tiles(const __grid_constant__ DeviceScene S, Grid3<uint8_t> t0, Grid3<uint8_t> t1,
                                     Grid3<uint8_t> t2) {
    forCells(S.N, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
    if (!((double)S.surface.raw(x, y, z) < 2.0 * S.dx0)) return;
    I3 c = mk3(x, y, z);
    markTile(t0, c); markTile(t0, cellToFace(c, 0, 1));
    markTile(t1, c); markTile(t1, cellToFace(c, 1, 1));
    markTile(t2, c); markTile(t2, cellToFace(c, 2, 1));
    });
}

// face activity test shared by AV.cpp:1127-1150 and AV.cpp:1235-1258
AVS_DEV __forceinline__ bool faceHasWeight(const DeviceScene &S, const I3 &face, int axis) {
    I3 b = faceToCell(face, axis, 0), f = faceToCell(face, axis, 1);
    if (S.centerW.get(b) > 0.f || S.centerW.get(f) > 0.f) return true;
#pragma unroll
    for (int ea = 0; ea < 3; ++ea) {
        if (ea == axis) continue;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir)
            if (S.edgeW[ea].get(faceToEdge(face, axis, ea, dir)) > 0.f) return true;
    }
    return false;
}

// classifyRegularVelocityFaces (AV.cpp:1087-1165): label of one regular-grid face
AVS_DEV __forceinline__ int8_t classifyRegular(const DeviceScene &S, int axis, const Grid3<uint8_t> &tiles, int x, int y, int z) {
    int8_t out = F_UNASSIGNED;
    I3 face = mk3(x, y, z);
    if (tileOccupied(tiles, x, y, z) && face[axis] - 1 >= 0 && face[axis] < S.N[axis]) {
        if (faceHasWeight(S, face, axis)) {
            double p[3];
            S.facePos(face, axis, 0, p);
            out = (S.collision.value(p) > -S.extrap) ? F_SOLID : F_FLUID;
        }
    }
    return out;
}
__global__ void k_classify_regular(const __grid_constant__ DeviceScene S, int axis, Grid3<uint8_t> tiles, int zOff) {
    const Grid3<int8_t> g = S.regular[axis];
    forCellsZ(g.n, zOff, [&](int x, int y, int z, size_t idx, bool ok) {
        if (ok) g.d[idx] = classifyRegular(S, axis, tiles, x, y, z);
    });
}

// classifyOctreeVelocityFaces (AV.cpp:1167-1323): label of one face
AVS_DEV __forceinline__ int32_t classifyFace(const DeviceScene &S, int level, int axis, const Grid3<uint8_t> &tiles, int x, int y, int z) {
    int32_t out = F_UNASSIGNED;
    I3 face = mk3(x, y, z);
    const Grid3<uint8_t> &lab = S.label[level];
    const int vr = (level == 0) ? S.N[axis] : lab.n[axis];  // AV.cpp:1185-1189
    if (level > 0 || tileOccupied(tiles, x, y, z)) {
        I3 b = faceToCell(face, axis, 0), f = face;
        if (b[axis] < 0 || f[axis] >= vr) {
            if (level == 0) out = F_OUTSIDE;  // AV.cpp:1210-1215
        } else {
            const int bl = lab.get(b), fl = lab.get(f);
            if (level == 0) {
                if (bl == L_ACTIVE && fl == L_ACTIVE) {
                    if (faceHasWeight(S, face, axis)) {
                        double p[3];
                        S.facePos(face, axis, 0, p);
                        out = (S.collision.value(p) > -S.extrap) ? F_SOLID : F_FLUID;
                    } else out = F_OUTSIDE;
                } else if (bl == L_INACTIVE || fl == L_INACTIVE) out = F_OUTSIDE;
                else if ((bl == L_UP && fl == L_ACTIVE) || (bl == L_ACTIVE && fl == L_UP)) out = F_FLUID;
            } else if ((bl == L_ACTIVE && fl == L_ACTIVE) || (bl == L_UP && fl == L_ACTIVE) || (bl == L_ACTIVE && fl == L_UP))
                out = F_FLUID;
        }
    }
    return out;
}
__global__ void k_classify_faces(const __grid_constant__ DeviceScene S, int level, int axis, Grid3<uint8_t> tiles) {
    const Grid3<int32_t> g = S.face[level][axis];
    forCells(g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
        g.d[idx] = classifyFace(S, level, axis, tiles, x, y, z);
    });
}

// findOccupiedEdgeStressTiles (AV.cpp:1002-1057): the 4 a-edges of every ACTIVE cell, a = 0,1,2
__global__ void k_mark_edge_tiles(Grid3<uint8_t> lab, Grid3<uint8_t> t0, Grid3<uint8_t> t1, Grid3<uint8_t> t2) {
    forCells(lab.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (!ok) return;
    if (lab.d[idx] != L_ACTIVE) return;
    I3 c = mk3(x, y, z);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        markTile(t0, cellToEdge(c, 0, e));
        markTile(t1, cellToEdge(c, 1, e));
        markTile(t2, cellToEdge(c, 2, e));
    }
    });
}

// classifyEdgeStresses (AV.cpp:1325-1405): label of one edge
AVS_DEV __forceinline__ int8_t classifyEdge(const DeviceScene &S, int level, int axis, const Grid3<uint8_t> &tiles, int x, int y, int z) {
    int8_t out = F_UNASSIGNED;
    if (tileOccupied(tiles, x, y, z)) {
        const Grid3<uint8_t> &lab = S.label[level];
        I3 edge = mk3(x, y, z);
        int vr[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) vr[a] = (level == 0) ? S.N[a] : lab.n[a];  // AV.cpp:1340-1344
        bool active = false;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            I3 cc = edgeToCell(edge, axis, ci);
            if (cc[0] < 0 || cc[1] < 0 || cc[2] < 0 || cc[0] >= vr[0] || cc[1] >= vr[1] || cc[2] >= vr[2]) {
                out = F_OUTSIDE;  // AV.cpp:1365-1370 (isStressActive keeps its value, as in the reference)
                break;
            }
            uint8_t l = lab.get(cc);
            if (l == L_DOWN) { active = false; break; }
            else if (l == L_ACTIVE) active = true;
        }
        if (active) {
            if (level == 0) out = (S.edgeW[axis].get(edge) > 0.f) ? F_FLUID : F_OUTSIDE;
            else out = F_FLUID;
        }
    }
    return out;
}
__global__ void k_classify_edges(const __grid_constant__ DeviceScene S, int level, int axis, Grid3<uint8_t> tiles, unsigned long long *counter) {
    const Grid3<int8_t> g = S.edge[level][axis];
    forCells(g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        int8_t out = F_UNASSIGNED;
        if (ok) {
            out = classifyEdge(S, level, axis, tiles, x, y, z);
            g.d[idx] = out;
        }
        countWarp(ok && out == F_FLUID, counter);
    });
}

// classifyCenterStresses (AV.cpp:1407-1443): label of one cell centre
AVS_DEV __forceinline__ int8_t classifyCenter(const DeviceScene &S, int level, int x, int y, int z, size_t idx) {
    int8_t out = F_UNASSIGNED;
    if (S.label[level].d[idx] == L_ACTIVE) {
        if (level != 0) out = F_FLUID;
        else if (S.centerW.get(mk3(x, y, z)) > 0.f) out = F_FLUID;
    }
    return out;
}
__global__ void k_classify_centers(const __grid_constant__ DeviceScene S, int level, unsigned long long *counter) {
    const Grid3<int8_t> g = S.center[level];
    forCells(g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        int8_t out = F_UNASSIGNED;
        if (ok) {
            out = classifyCenter(S, level, x, y, z, idx);
            g.d[idx] = out;
        }
        countWarp(ok && out == F_FLUID, counter);
    });
}

// ---- level 0 on the tiles that can hold anything but the default label ----------------------------------------------------------
// Level 0 is 87 % of all cells, but only a band around the surface is refined down to it: deep inside the liquid every cell is UP
// and every face / edge / centre label is UNASSIGNED, outside the reference never looks (unoccupied tiles keep UNASSIGNED,
// AV.cpp:886-1085).  So the level-0 label grids are memset to UNASSIGNED and the classification runs only on the 16^3 tiles where
// another label is possible -- decided per tile from two flags of the cell tile (any ACTIVE cell / any cell that is not UP) and the
// reference's own occupancy maps.  Same per-sample functions as the dense kernels above, hence the same labels.
enum : uint8_t { TILE_HAS_ACTIVE = 1, TILE_HAS_NON_UP = 2 };

// One CTA per cell tile: the two flags, and findOccupiedEdgeStressTiles (the 4 a-edges of every ACTIVE cell, a = 0,1,2) on the way.
__global__ void k_tile_flags(Grid3<uint8_t> lab, Grid3<uint8_t> flags, Grid3<uint8_t> t0, Grid3<uint8_t> t1, Grid3<uint8_t> t2) {
    const int tx = (int)blockIdx.x, ty = (int)blockIdx.y, tz = (int)blockIdx.z;
    const int y = ty * AVS_TILE + (int)(threadIdx.x & 15), z = tz * AVS_TILE + (int)(threadIdx.x >> 4);
    int mine = 0;
    if (y < lab.n[1] && z < lab.n[2]) {
        const int x0 = tx * AVS_TILE, x1 = min(x0 + AVS_TILE, lab.n[0]);
        const size_t row = (size_t)lab.n[0] * ((size_t)y + (size_t)lab.n[1] * (size_t)z);
        __align__(16) uint8_t cells[AVS_TILE];
        if (x1 - x0 == AVS_TILE && lab.n[0] % AVS_TILE == 0) *(uint4 *)cells = *(const uint4 *)(lab.d + row + x0);   // rows start 16-byte aligned
        else
            for (int x = x0; x < x1; ++x) cells[x - x0] = lab.d[row + x];
        for (int x = x0; x < x1; ++x) {
            const uint8_t l = cells[x - x0];
            if (l != L_UP) mine |= TILE_HAS_NON_UP;
            if (l == L_ACTIVE) {
                mine |= TILE_HAS_ACTIVE;
                const I3 c = mk3(x, y, z);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    markTile(t0, cellToEdge(c, 0, e));
                    markTile(t1, cellToEdge(c, 1, e));
                    markTile(t2, cellToEdge(c, 2, e));
                }
            }
        }
    }
    // __syncthreads_or returns "any thread's predicate is non-zero", not the bitwise OR: one vote per flag
    const int anyActive = __syncthreads_or(mine & TILE_HAS_ACTIVE), anyNonUp = __syncthreads_or(mine & TILE_HAS_NON_UP);
    if (threadIdx.x == 0) flags.d[flags.lin(tx, ty, tz)] = (uint8_t)((anyActive ? TILE_HAS_ACTIVE : 0) | (anyNonUp ? TILE_HAS_NON_UP : 0));
}

// Compacts the tiles of one grid that need classification into `list` (packed 10-bit coordinates; order irrelevant).
//   kind 0: face tiles of `axis` -- occupied (AV.cpp:886-943) and an adjacent cell may be other than UP, or the tile touches the
//           low / high end of the grid on `axis` (domain-boundary faces are OUTSIDE, AV.cpp:1210-1215)
//   kind 1: edge tiles -- occupied (AV.cpp:1002-1057)          kind 2: cell tiles with an ACTIVE cell
//   kind 3: face tiles of `axis` ABOVE level 0 (occ.d unused, only its dimensions): the reference classifies every face there, but a
//           face is FLUID only next to an ACTIVE cell (AV.cpp:1304-1318) and UNASSIGNED otherwise -- the tile or its lower neighbour
//           on `axis` must hold an ACTIVE cell
__global__ void k_tile_list(Grid3<uint8_t> occ, Grid3<uint8_t> flags, int kind, int axis, uint32_t *list, unsigned int *count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= occ.count()) return;
    const int tx = (int)(i % occ.n[0]), ty = (int)((i / occ.n[0]) % occ.n[1]), tz = (int)(i / ((size_t)occ.n[0] * occ.n[1]));
    bool take;
    if (kind == 2) take = (flags.d[i] & TILE_HAS_ACTIVE) != 0;   // occ == flags grid
    else if (kind == 1) take = occ.d[i] != 0;
    else if (kind == 3) {
        int t[3] = {tx, ty, tz};
        take = false;
        if (t[(axis + 1) % 3] < flags.n[(axis + 1) % 3] && t[(axis + 2) % 3] < flags.n[(axis + 2) % 3]) {
            if (t[axis] < flags.n[axis]) take = (flags.d[flags.lin(t[0], t[1], t[2])] & TILE_HAS_ACTIVE) != 0;
            if (!take && t[axis] > 0) {
                --t[axis];
                take = (flags.d[flags.lin(t[0], t[1], t[2])] & TILE_HAS_ACTIVE) != 0;
            }
        }
    } else {
        take = false;
        if (occ.d[i]) {
            int t[3] = {tx, ty, tz};
            const bool inGrid = tx < flags.n[0] && ty < flags.n[1] && tz < flags.n[2];
            take = !inGrid || t[axis] == 0 || (flags.d[flags.lin(tx, ty, tz)] & TILE_HAS_NON_UP);
            if (!take) {
                --t[axis];
                take = (flags.d[flags.lin(t[0], t[1], t[2])] & TILE_HAS_NON_UP) != 0;
            }
        }
    }
    if (take) list[atomicAdd(count, 1u)] = (uint32_t)tx | ((uint32_t)ty << 10) | ((uint32_t)tz << 20);
}

// one CTA (256 threads = 16 x 16) per listed tile, 16 planes each; launched with an upper bound of CTAs
template <class F>
__device__ __forceinline__ void forTileSamples(const uint32_t *list, const unsigned int *count, const int n[3], F f) {
    if (blockIdx.x >= *count) return;
    const uint32_t t = list[blockIdx.x];
    const int x = (int)(t & 1023u) * AVS_TILE + (int)(threadIdx.x & 15), y = (int)((t >> 10) & 1023u) * AVS_TILE + (int)(threadIdx.x >> 4);
    const int z0 = (int)(t >> 20) * AVS_TILE;
    const bool inxy = x < n[0] && y < n[1];
    for (int dz = 0; dz < AVS_TILE; ++dz) {
        const int z = z0 + dz;
        if (z >= n[2]) return;   // uniform over the CTA
        f(x, y, z, (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z), inxy);
    }
}
__global__ void k_classify_faces_tiles(const __grid_constant__ DeviceScene S, int level, int axis, Grid3<uint8_t> tiles, const uint32_t *list,
                                       const unsigned int *count) {
    const Grid3<int32_t> g = S.face[level][axis];
    forTileSamples(list, count, g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        if (ok) g.d[idx] = classifyFace(S, level, axis, tiles, x, y, z);
    });
}
__global__ void k_classify_edges_tiles(const __grid_constant__ DeviceScene S, int level, int axis, Grid3<uint8_t> tiles, const uint32_t *list,
                                       const unsigned int *count, unsigned long long *counter) {
    const Grid3<int8_t> g = S.edge[level][axis];
    forTileSamples(list, count, g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        int8_t out = F_UNASSIGNED;
        if (ok) {
            out = classifyEdge(S, level, axis, tiles, x, y, z);
            g.d[idx] = out;
        }
        countWarp(ok && out == F_FLUID, counter);
    });
}
__global__ void k_classify_centers_tiles(const __grid_constant__ DeviceScene S, int level, const uint32_t *list, const unsigned int *count,
                                         unsigned long long *counter) {
    const Grid3<int8_t> g = S.center[level];
    forTileSamples(list, count, g.n, [&](int x, int y, int z, size_t idx, bool ok) {
        int8_t out = F_UNASSIGNED;
        if (ok) {
            out = classifyCenter(S, level, x, y, z, idx);
            g.d[idx] = out;
        }
        countWarp(ok && out == F_FLUID, counter);
    });
}

static int tileGrid(AvsContext *c, DevBuf &buf, Grid3<uint8_t> t[3], const int res[3][3]) {
    size_t total = 0, off[3];
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 3; ++k) t[a].n[k] = (res[a][k] + AVS_TILE - 1) / AVS_TILE;
        off[a] = total;
        total += (t[a].count() + 255) / 256 * 256;
    }
    if (buf.reserve(total)) return AVS_ERR_ALLOC;
    for (int a = 0; a < 3; ++a) t[a].d = buf.as<uint8_t>() + off[a];
    AVS_CUDA_CHECK(cudaMemsetAsync(buf.p, 0, total, c->stream));
    return AVS_OK;
}

// the three level-0 face tile maps (c->tiles): dimensions from the PADDED face grids, which avs_stage_octree fixes later -- the
// padding is a pure function of the resolution, so it is computed here as well
static void faceTileRes(const AvsContext *c, int res[3][3]) {
    for (int k = 0; k < 3; ++k) {
        int pad = 1;
        while (pad < c->S.N[k]) pad <<= 1;   // OG.cpp:18-24
        for (int a = 0; a < 3; ++a) res[a][k] = pad + 1;
    }
}
static void avs_face_tile_views(AvsContext *c, Grid3<uint8_t> t[3]) {
    int res[3][3];
    faceTileRes(c, res);
    size_t total = 0;
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 3; ++k) t[a].n[k] = (res[a][k] + AVS_TILE - 1) / AVS_TILE;
        t[a].d = c->tiles.as<uint8_t>() + total;
        total += (t[a].count() + 255) / 256 * 256;
    }
}
static int avs_face_tile_maps(AvsContext *c, Grid3<uint8_t> t[3]) {   // allocates and zeroes
    int res[3][3];
    faceTileRes(c, res);
    return tileGrid(c, c->tiles, t, res);
}

int avs_stage_regular_labels(AvsContext *c) {
    DeviceScene &S = c->S;
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 3; ++k) S.regular[a].n[k] = S.N[k] + (k == a);
        if (c->regular[a].reserve(S.regular[a].count())) return AVS_ERR_ALLOC;
        S.regular[a].d = c->regular[a].as<int8_t>();
    }
    // One tile map per axis, dimensioned for the padded face grids; tile coordinates (idx/16) are the
    // same for the regular grid and for octree level 0, so stage 5 reuses these maps.
    Grid3<uint8_t> t[3];
    int rc = AVS_OK;
    if (c->surfaceTilesMarked) avs_face_tile_views(c, t);   // filled by the weights stage (k_sign_x4)
    else {
        rc = avs_face_tile_maps(c, t);
        if (rc) return rc;
        LAUNCH_3D(c, k_mark_surface_tiles, S.N, S, t[0], t[1], t[2]);
    }
    unsigned long long *cnt = c->counters.as<unsigned long long>();
    AVS_CUDA_CHECK(cudaMemsetAsync(cnt + 16, 0, 3 * sizeof(unsigned long long), c->stream));
    // Multi-GPU: the regular grid is only read by stage 11 (which regular faces are written back, AV.cpp:2843-2890) and by
    // the regular DOF count, both per-face work with no coupling -- so every rank classifies, and later fills, only its own
    // slab of z-planes (avs_slab_cuts); the count is summed over ranks, the output slabs are all-gathered after stage 11.
    int rc2 = avs_slab_cuts(c);
    if (rc2) return rc2;
    for (int a = 0; a < 3; ++a) {
        int z0, z1;
        avs_slab_range(c, a, c->rank, &z0, &z1);
        if (z1 <= z0) continue;
        dim3 g((unsigned)((S.regular[a].n[0] + 127) / 128), (unsigned)((S.regular[a].n[1] + CELL_YB - 1) / CELL_YB), (unsigned)(z1 - z0));
        k_classify_regular<<<g, dim3(128, CELL_YB), 0, c->stream>>>(S, a, t[a], z0);
        ++c->launches;
        // regular DOF count (most of the liquid volume: a fused per-warp atomic would serialise on one address)
        const size_t plane = (size_t)S.regular[a].n[0] * S.regular[a].n[1];
        const size_t cntCells = plane * (size_t)(z1 - z0);
        LAUNCH_COUNT(c, k_count_equal<int8_t>, cntCells, S.regular[a].d + plane * (size_t)z0, cntCells, (int8_t)F_FLUID, cnt + 16);
    }
    if (c->nranks > 1 && (rc2 = avs_dist_allreduce_u64(c, cnt + 16, 1))) return rc2;
    return AVS_OK;
}

// ---- z-slabs of the regular grid (multi-GPU) ---------------------------------------------------------------------
// Liquid cells per z-plane of the surface field: the weight of a plane in stage 11 (faces inside the liquid are the ones
// that get copied or interpolated).
__global__ void k_plane_liquid(DField surface, unsigned long long *planeCount) {
    const int z = (int)blockIdx.x;
    int cnt = 0;
    const size_t base = (size_t)surface.n[0] * surface.n[1] * (size_t)z;
    const size_t plane = (size_t)surface.n[0] * surface.n[1];
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.y * blockDim.x)
        cnt += surface.d[base + i] < 0.f;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(planeCount + z, (unsigned long long)cnt);
}

// slabZ[q] .. slabZ[q+1] = cell planes of rank q.  Identical on every rank (integer counts of identical inputs).
int avs_slab_cuts(AvsContext *c) {
    const int P = c->nranks, nz = c->S.N[2];
    c->slabZ.assign(P + 1, 0);
    c->slabZ[P] = nz;
    if (P == 1) return AVS_OK;
    std::vector<unsigned long long> h(nz, 0);
    if (c->S.surface.d) {
        if (c->scanTmp.reserve((size_t)nz * sizeof(unsigned long long))) return AVS_ERR_ALLOC;
        unsigned long long *d = c->scanTmp.as<unsigned long long>();
        AVS_CUDA_CHECK(cudaMemsetAsync(d, 0, (size_t)nz * sizeof(unsigned long long), c->stream));
        k_plane_liquid<<<dim3((unsigned)nz, 16), 256, 0, c->stream>>>(c->S.surface, d);
        ++c->launches;
        AVS_CUDA_CHECK(cudaMemcpyAsync(h.data(), d, (size_t)nz * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    // every plane also costs one streaming pass over its labels: a floor of 2 % of a full plane keeps empty regions cheap but not free
    const unsigned long long floorW = std::max<unsigned long long>(1, (unsigned long long)c->S.N[0] * c->S.N[1] / 50);
    unsigned long long total = 0;
    for (int z = 0; z < nz; ++z) { h[z] += floorW; total += h[z]; }
    unsigned long long run = 0;
    int q = 1;
    for (int z = 0; z < nz && q < P; ++z) {
        run += h[z];
        while (q < P && run * (unsigned long long)P >= total * (unsigned long long)q) c->slabZ[q++] = z + 1;
    }
    for (; q < P; ++q) c->slabZ[q] = nz;
    return AVS_OK;
}
// planes [z0, z1) of the axis-`axis` face grid that rank q owns (the z-face grid has one more plane: the last rank takes it)
void avs_slab_range(const AvsContext *c, int axis, int q, int *z0, int *z1) {
    const int P = c->nranks;
    if (P == 1 || c->slabZ.size() != (size_t)P + 1) { *z0 = 0; *z1 = c->S.regular[axis].n[2]; return; }
    *z0 = c->slabZ[q];
    *z1 = (q == P - 1) ? c->S.regular[axis].n[2] : c->slabZ[q + 1];
}

// ---- DOF numbering: count -> scan -> assign over 8^3-cell bricks in Morton order -----------------
// Bricks of all levels form one octree: a level-l brick (8^3 cells of level l) covers 2 bricks per axis of
// level l-1 (1 where that level has a single brick on the axis).  Rows are numbered in depth-first order of this
// tree -- a brick's own rows, then its children in Morton order -- so a coarse face sits next to the fine faces
// it couples to (T-junction stencils) and every contiguous row range is one compact region holding ALL levels:
// thin halos for the multi-GPU row partition, and level-0 bricks still appear in Morton order.
struct BrickLayout {
    int nb[AVS_MAX_LEVELS][3];     // bricks per axis
    int lf[AVS_MAX_LEVELS][3];     // log2(children per axis) of a level-l brick (0 for l == 0)
    long long sub[AVS_MAX_LEVELS]; // bricks in the subtree of one level-l brick (itself included)
    long long total;
    int levels;
};
#define BRICK 8

__device__ __forceinline__ void brickCoord(const BrickLayout &bl, int level, long long lin, int b[3]) {
    b[0] = (int)(lin % bl.nb[level][0]);
    b[1] = (int)((lin / bl.nb[level][0]) % bl.nb[level][1]);
    b[2] = (int)(lin / ((long long)bl.nb[level][0] * bl.nb[level][1]));
}

// depth-first position of brick b of `level`
__device__ __forceinline__ long long brickPos(const BrickLayout &bl, int level, const int b[3]) {
    const int T = bl.levels - 1;
    long long pos = 0;
    int sh[3] = {0, 0, 0};  // shift from `level` coordinates to level k-1 coordinates
    for (int k = level + 1; k <= T; ++k) {
        int ci = 0, bit = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (bl.lf[k][a]) {
                ci |= ((b[a] >> sh[a]) & 1) << bit;
                ++bit;
            }
        pos += 1 + (long long)ci * bl.sub[k - 1];
#pragma unroll
        for (int a = 0; a < 3; ++a) sh[a] += bl.lf[k][a];
    }
    long long t = (long long)(b[0] >> sh[0]) + (long long)bl.nb[T][0] * ((long long)(b[1] >> sh[1]) + (long long)bl.nb[T][1] * (b[2] >> sh[2]));
    return pos + t * bl.sub[T];
}

// Row order inside a brick: axis-major (all x-faces of the brick, then y, then z), cells z,y,x inside an axis.
// 32 consecutive rows then share their axis, which keeps the assembly kernel's warps on one control path
// (ncu: 7.3 of 32 lanes active per instruction with the cell-interleaved order).
__device__ __forceinline__ void brickThreadCount(const DeviceScene &S, int level, const int b[3], int t, int32_t vals[6], int n[3]) {
    n[0] = n[1] = n[2] = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int cellLocal = t * 2 + q;
        int x = b[0] * BRICK + (cellLocal & 7), y = b[1] * BRICK + ((cellLocal >> 3) & 7), z = b[2] * BRICK + (cellLocal >> 6);
        const Grid3<uint8_t> &lab = S.label[level];
        bool in = x < lab.n[0] && y < lab.n[1] && z < lab.n[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int32_t v = F_UNASSIGNED;
            if (in) v = S.face[level][a].d[S.face[level][a].lin(x, y, z)];
            vals[q * 3 + a] = v;
            n[a] += (v >= 0);
        }
    }
}

// `tileFlags` (d may be null): flags of the level's 16^3 cell tiles (k_tile_flags).  A face is a DOF only next to an ACTIVE
// cell, so a brick whose tile and whose three lower neighbour tiles hold no ACTIVE cell has no rows: it is not read at all.
__global__ void k_brick_count(const __grid_constant__ DeviceScene S, const __grid_constant__ BrickLayout bl, int level, int32_t *brickCount,
                              int32_t *brickCost, Grid3<uint8_t> tileFlags) {
    int b[3];
    brickCoord(bl, level, blockIdx.x, b);
    if (tileFlags.d) {
        const int t[3] = {b[0] * BRICK / AVS_TILE, b[1] * BRICK / AVS_TILE, b[2] * BRICK / AVS_TILE};
        bool any = (tileFlags.d[tileFlags.lin(t[0], t[1], t[2])] & TILE_HAS_ACTIVE) != 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (any || t[a] == 0 || (b[a] * BRICK) % AVS_TILE != 0) continue;   // only the brick at the low end of a tile sees the tile below
            int q[3] = {t[0], t[1], t[2]};
            --q[a];
            any = (tileFlags.d[tileFlags.lin(q[0], q[1], q[2])] & TILE_HAS_ACTIVE) != 0;
        }
        if (!any) {   // uniform over the CTA
            if (threadIdx.x == 0) {
                const long long pos = brickPos(bl, level, b);
                brickCount[pos] = 0;
                brickCost[pos] = 0;
            }
            return;
        }
    }
    int32_t vals[6];
    int na[3];
    brickThreadCount(S, level, b, threadIdx.x, vals, na);
    int n = na[0] + na[1] + na[2];
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        long long pos = brickPos(bl, level, b);
        brickCount[pos] = t;
        // estimated SpMV cost in units of 1/20 fine row: rows above level 0 carry ~1.65x the non-zeros (T-junction stencils)
        brickCost[pos] = t * (level == 0 ? 20 : 33);
    }
}

__global__ void k_brick_assign(const __grid_constant__ DeviceScene S, const __grid_constant__ BrickLayout bl, int level,
                               const int32_t *brickCount, const long long *brickOffset, RowKey *keys) {
    int b[3];
    brickCoord(bl, level, blockIdx.x, b);
    const long long pos = brickPos(bl, level, b);
    if (brickCount[pos] == 0) return;
    int32_t vals[6];
    int na[3];
    brickThreadCount(S, level, b, threadIdx.x, vals, na);
    // block-wide exclusive scans, one per axis
    __shared__ int warpSum[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        inc[a] = na[a];
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, inc[a], o);
            if (lane >= o) inc[a] += v;
        }
        if (lane == 31) warpSum[a][wid] = inc[a];
    }
    __syncthreads();
    long long axisBase = brickOffset[pos];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int wbase = 0, total = 0;
        for (int i = 0; i < 8; ++i) {
            if (i < wid) wbase += warpSum[a][i];
            total += warpSum[a][i];
        }
        long long idx = axisBase + wbase + inc[a] - na[a];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (vals[q * 3 + a] >= 0) {
                int cellLocal = threadIdx.x * 2 + q;
                int x = b[0] * BRICK + (cellLocal & 7), y = b[1] * BRICK + ((cellLocal >> 3) & 7), z = b[2] * BRICK + (cellLocal >> 6);
                S.face[level][a].d[S.face[level][a].lin(x, y, z)] = (int32_t)idx;
                RowKey k;
                k.level = level; k.axis = a; k.i = x; k.j = y; k.k = z;
                keys[idx] = k;
                ++idx;
            }
        }
        axisBase += total;
    }
}

// ---- exclusive scan int32 -> int64 (three phases, 2048 elements per block) ----------------------
#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
__global__ void k_scan_block_sums(const int32_t *in, long long n, long long *blockSums) {
    long long base = (long long)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
    long long s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        long long j = base + (long long)i * SCAN_BLOCK + threadIdx.x;
        if (j < n) s += in[j];
    }
    __shared__ long long sh[SCAN_BLOCK];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = SCAN_BLOCK / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) blockSums[blockIdx.x] = sh[0];
}
__global__ void k_scan_sums_serial(long long *blockSums, long long nb, long long *total) {
    // nb is a few thousand at most: one thread, exact and deterministic
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long run = 0;
        for (long long i = 0; i < nb; ++i) {
            long long v = blockSums[i];
            blockSums[i] = run;
            run += v;
        }
        *total = run;
    }
}
__global__ void k_scan_apply(const int32_t *in, long long n, const long long *blockSums, long long *out) {
    // each thread owns SCAN_ITEMS consecutive elements
    long long base = (long long)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (long long)threadIdx.x * SCAN_ITEMS;
    long long v[SCAN_ITEMS], s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        long long j = base + i;
        v[i] = (j < n) ? in[j] : 0;
        s += v[i];
    }
    __shared__ long long sh[SCAN_BLOCK];
    sh[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 256 thread sums
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        long long t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    long long run = blockSums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        long long j = base + i;
        if (j < n) out[j] = run;
        run += v[i];
    }
}

int avs_exclusive_scan_i32_to_i64(AvsContext *c, const int32_t *dIn, int64_t *dOut, int64_t n, int64_t *hostTotal) {
    if (n <= 0) { if (hostTotal) *hostTotal = 0; return AVS_OK; }
    long long nb = (n + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
    if (c->scanTmp.reserve((size_t)(nb + 1) * sizeof(long long))) return AVS_ERR_ALLOC;
    long long *sums = c->scanTmp.as<long long>();
    k_scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, c->stream>>>(dIn, n, sums);
    k_scan_sums_serial<<<1, 32, 0, c->stream>>>(sums, nb, sums + nb);
    k_scan_apply<<<(unsigned)nb, SCAN_BLOCK, 0, c->stream>>>(dIn, n, sums, (long long *)dOut);
    c->launches += 3;
    if (hostTotal) {
        long long t = 0;
        AVS_CUDA_CHECK(cudaMemcpyAsync(&t, sums + nb, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        *hostTotal = t;
    }
    return AVS_OK;
}

// rank q's block starts at the first brick whose cost prefix reaches total*q/P (brick-granular, contiguous)
__global__ void k_find_cuts(const long long *costPrefix, const long long *rowOffset, long long nbricks, long long totalCost,
                            long long nRows, int P, long long *rowStarts) {
    int q = threadIdx.x;
    if (q > P) return;
    if (q == 0) { rowStarts[0] = 0; return; }
    if (q == P) { rowStarts[P] = nRows; return; }
    long long target = totalCost * q / P;
    long long lo = 0, hi = nbricks;  // first brick with prefix >= target
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (costPrefix[mid] >= target) hi = mid; else lo = mid + 1;
    }
    rowStarts[q] = lo < nbricks ? rowOffset[lo] : nRows;
}

int avs_stage_octree_labels(AvsContext *c) {
    DeviceScene &S = c->S;
    const int L = S.levels;
    c->tileFlagsGrid.d = nullptr;
    if (L == 0) {  // no ACTIVE cell at all (the reference asserts this away, OG.cpp:206): an empty system
        c->nRows = 0;
        c->nEdge = c->nCenter = 0;
        c->rowStarts.assign(c->nranks + 1, 0);
        unsigned long long h = 0;
        AVS_CUDA_CHECK(cudaMemcpyAsync(&h, c->counters.as<unsigned long long>() + 16, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->nRegular = (int64_t)h;
        if (c->rowKeys.reserve(sizeof(RowKey))) return AVS_ERR_ALLOC;
        return AVS_OK;
    }
    // allocate + classify, level by level
    Grid3<uint8_t> t0[3];
    {   // level-0 face tile maps were produced by avs_stage_regular_labels and still live in c->tiles
        int res[3][3];
        size_t total = 0;
        for (int a = 0; a < 3; ++a) {
            for (int k = 0; k < 3; ++k) { res[a][k] = S.Pad[k] + 1; t0[a].n[k] = (res[a][k] + AVS_TILE - 1) / AVS_TILE; }
            t0[a].d = c->tiles.as<uint8_t>() + total;
            total += (t0[a].count() + 255) / 256 * 256;
        }
    }
    unsigned long long *cnt = c->counters.as<unsigned long long>();
    DevBuf &edgeTiles = c->edgeTiles;
    for (int l = 0; l < L; ++l) {
        const Grid3<uint8_t> &lab = S.label[l];
        for (int a = 0; a < 3; ++a) {
            Grid3<int32_t> &f = S.face[l][a];
            Grid3<int8_t> &e = S.edge[l][a];
            for (int k = 0; k < 3; ++k) { f.n[k] = lab.n[k] + (k == a); e.n[k] = lab.n[k] + (k != a); }
            if (c->face[l][a].reserve(f.count() * sizeof(int32_t))) return AVS_ERR_ALLOC;
            if (c->edge[l][a].reserve(e.count())) return AVS_ERR_ALLOC;
            f.d = c->face[l][a].as<int32_t>();
            e.d = c->edge[l][a].as<int8_t>();
        }
        Grid3<int8_t> &ce = S.center[l];
        for (int k = 0; k < 3; ++k) ce.n[k] = lab.n[k];
        if (c->center[l].reserve(ce.count())) return AVS_ERR_ALLOC;
        ce.d = c->center[l].as<int8_t>();
    }
    // AVS_LABELS=dense: level 0 through the dense sweeps as well (round 1; A/B measurements and the bit-equality test)
    static int dense0 = -1;
    if (dense0 < 0) { const char *e = getenv("AVS_LABELS"); dense0 = (e && strcmp(e, "dense") == 0) ? 1 : 0; }
    // tile flags of every level that is classified tile by tile: levels with at least 4 tiles on every axis (the others are tiny)
    Grid3<uint8_t> flagsL[AVS_MAX_LEVELS];
    size_t flagBytes = 0, flagOff[AVS_MAX_LEVELS];
    for (int l = 0; l < L; ++l) {
        flagsL[l].d = nullptr;
        for (int k = 0; k < 3; ++k) flagsL[l].n[k] = (S.label[l].n[k] + AVS_TILE - 1) / AVS_TILE;
        flagOff[l] = flagBytes;
        const bool tiled = !dense0 && (l == 0 || (flagsL[l].n[0] >= 4 && flagsL[l].n[1] >= 4 && flagsL[l].n[2] >= 4));
        if (tiled) flagBytes += (flagsL[l].count() + 255) / 256 * 256;
        else flagsL[l].n[0] = 0;   // marks "dense"
    }
    if (flagBytes && c->tileFlags.reserve(flagBytes)) return AVS_ERR_ALLOC;
    for (int l = 0; l < L; ++l)
        if (flagsL[l].n[0] > 0) flagsL[l].d = c->tileFlags.as<uint8_t>() + flagOff[l];
    Grid3<uint8_t> tileFlags = flagsL[0];   // level 0: also used by the interpolator (avs_prolong.cu)
    if (!tileFlags.d) tileFlags.n[0] = tileFlags.n[1] = tileFlags.n[2] = 0;
    for (int l = 0; l < L; ++l) {
        Grid3<uint8_t> te[3];
        int res[3][3];
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 3; ++k) res[a][k] = S.edge[l][a].n[k];
        int rc = tileGrid(c, edgeTiles, te, res);
        if (rc) return rc;
        if (flagsL[l].d) {
            // ---- tile-culled (see k_tile_flags) ----
            const Grid3<uint8_t> fl = flagsL[l];
            Grid3<uint8_t> ft[3];   // face tile grids: the occupancy maps at level 0, dimensions only above
            for (int a = 0; a < 3; ++a) {
                ft[a] = t0[a];
                if (l > 0) {
                    ft[a].d = nullptr;
                    for (int k = 0; k < 3; ++k) ft[a].n[k] = (S.label[l].n[k] + 1 + AVS_TILE - 1) / AVS_TILE;
                }
            }
            size_t listCap = fl.count();
            for (int a = 0; a < 3; ++a) listCap = std::max(listCap, std::max(ft[a].count(), te[a].count()));
            if (ft[0].n[0] > 1023 || ft[0].n[1] > 1023 || ft[0].n[2] > 1023) return AVS_ERR_UNSUPPORTED;   // packed tile coordinates
            if (c->tileLists.reserve(7 * listCap * sizeof(uint32_t))) return AVS_ERR_ALLOC;
            unsigned int *lcnt = (unsigned int *)(cnt + 32);   // 7 list lengths in counter slots 32..35
            AVS_CUDA_CHECK(cudaMemsetAsync(lcnt, 0, 8 * sizeof(unsigned int), c->stream));
            k_tile_flags<<<dim3((unsigned)fl.n[0], (unsigned)fl.n[1], (unsigned)fl.n[2]), 256, 0, c->stream>>>(S.label[l], fl, te[0], te[1], te[2]);
            ++c->launches;
            uint32_t *lists = c->tileLists.as<uint32_t>();
            for (int a = 0; a < 3; ++a) {
                AVS_CUDA_CHECK(cudaMemsetAsync(S.face[l][a].d, 0xFF, S.face[l][a].count() * sizeof(int32_t), c->stream));   // F_UNASSIGNED
                AVS_CUDA_CHECK(cudaMemsetAsync(S.edge[l][a].d, 0xFF, S.edge[l][a].count(), c->stream));
            }
            AVS_CUDA_CHECK(cudaMemsetAsync(S.center[l].d, 0xFF, S.center[l].count(), c->stream));
            for (int a = 0; a < 3; ++a) {
                k_tile_list<<<(unsigned)((ft[a].count() + 255) / 256), 256, 0, c->stream>>>(ft[a], fl, l == 0 ? 0 : 3, a, lists + (size_t)a * listCap, lcnt + a);
                k_classify_faces_tiles<<<(unsigned)ft[a].count(), 256, 0, c->stream>>>(S, l, a, t0[a], lists + (size_t)a * listCap, lcnt + a);
                k_tile_list<<<(unsigned)((te[a].count() + 255) / 256), 256, 0, c->stream>>>(te[a], fl, 1, a, lists + (size_t)(3 + a) * listCap, lcnt + 3 + a);
                k_classify_edges_tiles<<<(unsigned)te[a].count(), 256, 0, c->stream>>>(S, l, a, te[a], lists + (size_t)(3 + a) * listCap, lcnt + 3 + a, cnt + 17);
                c->launches += 4;
            }
            k_tile_list<<<(unsigned)((fl.count() + 255) / 256), 256, 0, c->stream>>>(fl, fl, 2, 0, lists + 6 * listCap, lcnt + 6);
            k_classify_centers_tiles<<<(unsigned)fl.count(), 256, 0, c->stream>>>(S, l, lists + 6 * listCap, lcnt + 6, cnt + 18);
            c->launches += 2;
            continue;
        }
        for (int a = 0; a < 3; ++a) LAUNCH_3D(c, k_classify_faces, S.face[l][a].n, S, l, a, t0[a]);
        LAUNCH_3D(c, k_mark_edge_tiles, S.label[l].n, S.label[l], te[0], te[1], te[2]);
        for (int a = 0; a < 3; ++a) LAUNCH_3D(c, k_classify_edges, S.edge[l][a].n, S, l, a, te[a], cnt + 17);
        LAUNCH_3D(c, k_classify_centers, S.center[l].n, S, l, cnt + 18);
    }
    c->tileFlagsGrid = tileFlags;
    // (stress DOF counts -- the reference numbers them, AV.cpp:1632-1715 -- were accumulated by the classify kernels)
    // numbering
    BrickLayout bl;
    memset(&bl, 0, sizeof(bl));
    bl.levels = L;
    for (int l = 0; l < L; ++l)
        for (int a = 0; a < 3; ++a) bl.nb[l][a] = std::max(1, S.label[l].n[a] / BRICK);
    for (int l = 1; l < L; ++l)
        for (int a = 0; a < 3; ++a) bl.lf[l][a] = ilog2(bl.nb[l - 1][a] / bl.nb[l][a]);
    bl.sub[0] = 1;
    for (int l = 1; l < L; ++l) bl.sub[l] = 1 + ((long long)1 << (bl.lf[l][0] + bl.lf[l][1] + bl.lf[l][2])) * bl.sub[l - 1];
    const long long total = (long long)bl.nb[L - 1][0] * bl.nb[L - 1][1] * bl.nb[L - 1][2] * bl.sub[L - 1];
    bl.total = total;
    if (c->brickCount.reserve((size_t)total * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->brickCost.reserve((size_t)total * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->brickOffset.reserve((size_t)total * sizeof(long long))) return AVS_ERR_ALLOC;
    for (int l = 0; l < L; ++l) {
        unsigned nb = (unsigned)((long long)bl.nb[l][0] * bl.nb[l][1] * bl.nb[l][2]);
        k_brick_count<<<nb, 256, 0, c->stream>>>(S, bl, l, c->brickCount.as<int32_t>(), c->brickCost.as<int32_t>(), flagsL[l]);
        ++c->launches;
    }
    int64_t nRows = 0;
    int rc = avs_exclusive_scan_i32_to_i64(c, c->brickCount.as<int32_t>(), c->brickOffset.as<int64_t>(), total, &nRows);
    if (rc) return rc;
    c->nRows = nRows;
    if (nRows >= (int64_t)2147483000) return AVS_ERR_UNSUPPORTED;  // int32 DOF indices
    if (c->rowKeys.reserve((size_t)std::max<int64_t>(nRows, 1) * sizeof(RowKey))) return AVS_ERR_ALLOC;
    for (int l = 0; l < L; ++l) {
        unsigned nb = (unsigned)((long long)bl.nb[l][0] * bl.nb[l][1] * bl.nb[l][2]);
        k_brick_assign<<<nb, 256, 0, c->stream>>>(S, bl, l, c->brickCount.as<int32_t>(), c->brickOffset.as<long long>(),
                                                  c->rowKeys.as<RowKey>());
        ++c->launches;
    }
    // multi-GPU row partition: contiguous blocks of the depth-first order, cut at brick
    // boundaries where the estimated SpMV cost prefix crosses q/P of the total
    c->rowStarts.assign(c->nranks + 1, 0);
    c->rowStarts[c->nranks] = nRows;
    if (c->nranks > 1) {
        if (c->brickCostPrefix.reserve((size_t)(total + 32) * sizeof(long long))) return AVS_ERR_ALLOC;
        int64_t totalCost = 0;
        rc = avs_exclusive_scan_i32_to_i64(c, c->brickCost.as<int32_t>(), c->brickCostPrefix.as<int64_t>(), total, &totalCost);
        if (rc) return rc;
        long long *dStarts = c->brickCostPrefix.as<long long>() + total;
        k_find_cuts<<<1, 32, 0, c->stream>>>(c->brickCostPrefix.as<long long>(), c->brickOffset.as<long long>(), total, totalCost, nRows,
                                            c->nranks, dStarts);
        ++c->launches;
        AVS_CUDA_CHECK(cudaMemcpyAsync(c->rowStarts.data(), dStarts, (size_t)(c->nranks + 1) * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    }
    unsigned long long h[3];
    AVS_CUDA_CHECK(cudaMemcpyAsync(h, cnt + 16, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->nRegular = (int64_t)h[0];
    c->nEdge = (int64_t)h[1];
    c->nCenter = (int64_t)h[2];
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Octree geometry dump (HDK_OctreeGrid::outputOctreeGeometry, OG.cpp:245-308, reached from AV.cpp:283-294 when
// doPrintOctree is set): one point per ACTIVE cell of every built level, carrying the cell centre, pscale = the
// level's voxel size and octreeLevel.  The reference appends points while walking tiles; here the order is
// level-major, x-fastest (deterministic: per-block counts -> exclusive scan -> ordered fill), and the point set
// is what parity is judged on.
#define GEO_BLOCK 256
__global__ void k_geo_count(Grid3<uint8_t> g, int32_t *blockCount) {
    size_t idx = (size_t)blockIdx.x * GEO_BLOCK + threadIdx.x;
    bool hit = idx < g.count() && g.d[idx] == L_ACTIVE;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    __shared__ int s[GEO_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < GEO_BLOCK / 32; ++i) t += s[i];
        blockCount[blockIdx.x] = t;
    }
}
__global__ void k_geo_fill(DeviceScene S, int level, const long long *blockOffset, long long base, float *pos, float *pscale,
                           int32_t *lvl) {
    const Grid3<uint8_t> &g = S.label[level];
    size_t idx = (size_t)blockIdx.x * GEO_BLOCK + threadIdx.x;
    bool hit = idx < g.count() && g.d[idx] == L_ACTIVE;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    __shared__ int s[GEO_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (!hit) return;
    int before = __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s[w];
    long long o = base + blockOffset[blockIdx.x] + before;
    I3 c = mk3((int)(idx % g.n[0]), (int)((idx / g.n[0]) % g.n[1]), (int)(idx / ((size_t)g.n[0] * g.n[1])));
    double p[3];
    S.centerPos(c, level, p);      // SIM_RawField::indexToPos of a centre-sampled grid, stored as UT_Vector3 (fp32)
    pos[3 * o + 0] = (float)p[0];
    pos[3 * o + 1] = (float)p[1];
    pos[3 * o + 2] = (float)p[2];
    pscale[o] = (float)S.levelDx(level);
    lvl[o] = level;
}

// Writes the points to the device buffers of the context (geoPos / geoScale / geoLevel) and returns the count.
int avs_octree_points(AvsContext *c, int64_t *countOut) {
    DeviceScene &S = c->S;
    int64_t total = 0;
    std::vector<int64_t> perLevel(S.levels, 0);
    std::vector<size_t> blocks(S.levels, 0);
    size_t maxBlocks = 1, sumBlocks = 0;
    for (int l = 0; l < S.levels; ++l) {
        blocks[l] = (S.label[l].count() + GEO_BLOCK - 1) / GEO_BLOCK;
        maxBlocks = std::max(maxBlocks, blocks[l]);
        sumBlocks += blocks[l];
    }
    if (c->geoCount.reserve(std::max<size_t>(sumBlocks, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->geoOffset.reserve(std::max<size_t>(sumBlocks, 1) * sizeof(long long))) return AVS_ERR_ALLOC;
    // pass 1: counts and per-level offsets
    size_t bo = 0;
    for (int l = 0; l < S.levels; ++l) {
        k_geo_count<<<(unsigned)blocks[l], GEO_BLOCK, 0, c->stream>>>(S.label[l], c->geoCount.as<int32_t>() + bo);
        ++c->launches;
        int rc = avs_exclusive_scan_i32_to_i64(c, c->geoCount.as<int32_t>() + bo, c->geoOffset.as<int64_t>() + bo, (int64_t)blocks[l],
                                               &perLevel[l]);
        if (rc) return rc;
        total += perLevel[l];
        bo += blocks[l];
    }
    if (c->geoPos.reserve((size_t)std::max<int64_t>(total, 1) * 3 * sizeof(float))) return AVS_ERR_ALLOC;
    if (c->geoScale.reserve((size_t)std::max<int64_t>(total, 1) * sizeof(float))) return AVS_ERR_ALLOC;
    if (c->geoLevel.reserve((size_t)std::max<int64_t>(total, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    // pass 2: ordered fill
    bo = 0;
    long long base = 0;
    for (int l = 0; l < S.levels; ++l) {
        if (perLevel[l] > 0) {
            k_geo_fill<<<(unsigned)blocks[l], GEO_BLOCK, 0, c->stream>>>(S, l, c->geoOffset.as<long long>() + bo, base, c->geoPos.as<float>(),
                                                                          c->geoScale.as<float>(), c->geoLevel.as<int32_t>());
            ++c->launches;
        }
        base += perLevel[l];
        bo += blocks[l];
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    *countOut = total;
    return AVS_OK;
}
