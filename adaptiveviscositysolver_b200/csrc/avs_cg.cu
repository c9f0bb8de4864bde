// avs_cg.cu -- stage 10: the Jacobi-preconditioned conjugate-gradient solve the reference hands to
// Eigen::ConjugateGradient<SparseMatrix<SolveType>, Lower|Upper> (HDK_AdaptiveViscosity.cpp:611-630).
// The iteration follows upstream Eigen's conjugate_gradient() + DiagonalPreconditioner exactly
// (SURVEY.md section 8a row a15), including "break before the iteration counter is bumped".
//
// Matrix format: SJDS-32 ("sliced jagged diagonals, paired").  Rows are grouped in slices of 32
// (one warp); inside a slice rows are sorted by length, and the slice stores, for pair index j, the
// c_j rows that still have entries -- contiguously, no padding.  One lane owns one row and reads
// two consecutive non-zeros per step with one 128-bit value load and one 64-bit column load; the
// loads of a warp are one contiguous segment.  Rows of this system have 2..46 entries (mean ~16),
// so a padded ELL/SELL layout would move ~1.6x the bytes (measured on the oracle's matrices).
//
// Per CG iteration: 3 kernels (SpMV + p.Ap | x,r update + r.r, r.z | p update), all scalars stay on
// the device, dot products are reduced per CTA and finished in fixed order by the consumer kernel
// (deterministic, no atomics), convergence is a device-side flag that turns later launches into
// no-ops, so the host only polls every `check_every` iterations without changing the result.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "avs_context.h"
#include "avs_p2p.cuh"

#define CG_THREADS 256
#define SJDS_HALO_BIT 0x40000000   // meta word: this slice gathers from a halo slot (multi-GPU; set per solve by k_slice_needs_halo)
#define CG_CTAS_PER_SM 8

// ---- SJDS build ---------------------------------------------------------------------------------
// meta[slice*32 + q] = (original lane << 8) | number of pairs, q = position after sorting by length
__global__ void k_row_lengths(long long n, const long long *ptr, int32_t *len) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) len[r] = (int32_t)(ptr[r + 1] - ptr[r]);
}
__global__ void k_sjds_count(long long n, const int32_t *rowLen, long long nslices, int32_t *meta, int32_t *slicePairs) {
    long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    long long r = s * 32 + lane;
    int len = (r < n) ? rowLen[r] : 0;
    int np = (len + 1) >> 1;
    int rank = 0, total = 0;
    for (int o = 0; o < 32; ++o) {
        int other = __shfl_sync(0xffffffffu, np, o);
        rank += (other > np) || (other == np && o < lane);
        total += other;
    }
    meta[s * 32 + rank] = (lane << 8) | np;
    if (lane == 0) slicePairs[s] = total + (total & 1);  // even: slice starts are 16-byte aligned in val2 AND col2 (bulk copies)
}

// fill from the assembly staging area (entry j of row r at j*stride + r)
template <class T, class T2>
__global__ void k_sjds_fill_stage(long long n, const int32_t *rowLen, const int32_t *stageCol, const double *stageVal, long long stride,
                                  long long nslices, const int32_t *meta, const long long *sliceOff, T2 *val2, int2 *col2,
                                  long long rowBegin, long long rowEnd, const long long *haloIndex) {
    long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    int m = meta[s * 32 + lane];
    int np = m & 0xff, orig = (m >> 8) & 31;
    long long r = s * 32 + orig;
    int len = (r < n) ? rowLen[r] : 0;
    long long base = sliceOff[s];
    int maxnp = __shfl_sync(0xffffffffu, np, 0);
    long long off = 0;
    for (int j = 0; j < maxnp; ++j) {
        bool active = j < np;
        int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            long long k0 = (long long)(2 * j) * stride + r, k1 = k0 + stride;
            T2 v;
            int2 cc;
            v.x = (T)stageVal[k0];
            cc.x = stageCol[k0];
            if (2 * j + 1 < len) { v.y = (T)stageVal[k1]; cc.y = stageCol[k1]; }
            else { v.y = (T)0; cc.y = cc.x; }
            // global column -> local index: owned rows first, then the halo slots (multi-GPU row partition)
            cc.x = (cc.x >= rowBegin && cc.x < rowEnd) ? (int)(cc.x - rowBegin) : (int)((rowEnd - rowBegin) + haloIndex[cc.x]);
            cc.y = (cc.y >= rowBegin && cc.y < rowEnd) ? (int)(cc.y - rowBegin) : (int)((rowEnd - rowBegin) + haloIndex[cc.y]);
            val2[base + off + lane] = v;
            col2[base + off + lane] = cc;
        }
        off += cnt;
    }
}
template <class T>
__global__ void k_inv_diag_from_array(long long n, const double *diag, T *invDiag) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) invDiag[r] = (diag[r] != 0.0) ? (T)(1.0 / diag[r]) : (T)1;
}

template <class T, class T2>
__global__ void k_sjds_fill(long long n, const long long *ptr, const int32_t *col, const double *val, long long nslices,
                            const int32_t *meta, const long long *sliceOff, T2 *val2, int2 *col2) {
    long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    int m = meta[s * 32 + lane];
    int np = m & 0xff, orig = (m >> 8) & 31;
    long long r = s * 32 + orig;
    long long p0 = (r < n) ? ptr[r] : 0;
    int len = (r < n) ? (int)(ptr[r + 1] - p0) : 0;
    long long base = sliceOff[s];
    int maxnp = __shfl_sync(0xffffffffu, np, 0);
    long long off = 0;
    for (int j = 0; j < maxnp; ++j) {
        bool active = j < np;
        int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            int k0 = 2 * j, k1 = 2 * j + 1;
            T2 v;
            int2 cc;
            v.x = (T)val[p0 + k0];
            cc.x = col[p0 + k0];
            if (k1 < len) { v.y = (T)val[p0 + k1]; cc.y = col[p0 + k1]; }
            else { v.y = (T)0; cc.y = cc.x; }  // odd row length: a zero paired with a harmless column
            val2[base + off + lane] = v;
            col2[base + off + lane] = cc;
        }
        off += cnt;
    }
}

// DiagonalPreconditioner::factorize: invdiag = 1/A_jj when the diagonal exists and is non-zero, else 1
template <class T>
__global__ void k_inv_diag(long long n, const long long *ptr, const int32_t *col, const double *val, long long colOfRow0, T *invDiag) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0;
    bool found = false;
    for (long long k = ptr[r]; k < ptr[r + 1]; ++k)
        if (col[k] == (int32_t)(r + colOfRow0)) { d = val[k]; found = true; break; }
    invDiag[r] = (found && d != 0.0) ? (T)(1.0 / d) : (T)1;
}

// common part: slice metadata + offsets from the row lengths; returns the number of pairs
static int sjdsLayout(AvsContext *c, SellMatrix &A, int64_t n, const int32_t *dLen, int precision, int64_t *totalPairsOut) {
    A.n = n;
    A.precision = precision;
    A.nslices = (n + 31) / 32;
    const size_t vs = precision == AVS_PRECISION_F32 ? sizeof(float) : sizeof(double);
    if (A.sliceOff.reserve((size_t)(A.nslices + 1) * sizeof(long long))) return AVS_ERR_ALLOC;
    if (c->slicePairs.reserve((size_t)std::max<int64_t>(A.nslices, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (A.meta.reserve((size_t)std::max<int64_t>(A.nslices, 1) * 32 * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (A.invDiag.reserve((size_t)std::max<int64_t>(n, 1) * vs)) return AVS_ERR_ALLOC;
    A.padded = 0;
    *totalPairsOut = 0;
    if (n == 0) return AVS_OK;
    unsigned blocks = (unsigned)((A.nslices * 32 + 255) / 256);
    k_sjds_count<<<blocks, 256, 0, c->stream>>>(n, dLen, A.nslices, A.meta.as<int32_t>(), c->slicePairs.as<int32_t>());
    ++c->launches;
    int64_t totalPairs = 0;
    int rc = avs_exclusive_scan_i32_to_i64(c, c->slicePairs.as<int32_t>(), A.sliceOff.as<int64_t>(), A.nslices, &totalPairs);
    if (rc) return rc;
    AVS_CUDA_CHECK(cudaMemcpyAsync(A.sliceOff.as<long long>() + A.nslices, &totalPairs, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    A.padded = totalPairs * 2;
    if (A.val.reserve((size_t)std::max<int64_t>(totalPairs, 1) * 2 * vs)) return AVS_ERR_ALLOC;
    if (A.col.reserve((size_t)std::max<int64_t>(totalPairs, 1) * sizeof(int2))) return AVS_ERR_ALLOC;
    *totalPairsOut = totalPairs;
    return AVS_OK;
}

int avs_sell_from_csr(AvsContext *c, SellMatrix &A, int64_t n, const int64_t *dPtr, const int32_t *dCol, const double *dVal,
                      int precision) {
    A.nnz = 0;
    if (c->rowCount.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (n > 0) {
        k_row_lengths<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, (const long long *)dPtr, c->rowCount.as<int32_t>());
        ++c->launches;
    }
    int64_t totalPairs = 0;
    int rc = sjdsLayout(c, A, n, c->rowCount.as<int32_t>(), precision, &totalPairs);
    if (rc || n == 0) return rc;
    long long nnz = 0;
    AVS_CUDA_CHECK(cudaMemcpyAsync(&nnz, (const long long *)dPtr + n, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    A.nnz = nnz;
    unsigned blocks = (unsigned)((A.nslices * 32 + 255) / 256);
    if (precision == AVS_PRECISION_F32) {
        k_sjds_fill<float, float2><<<blocks, 256, 0, c->stream>>>(n, (const long long *)dPtr, dCol, dVal, A.nslices, A.meta.as<int32_t>(),
                                                                 A.sliceOff.as<long long>(), A.val.as<float2>(), A.col.as<int2>());
        k_inv_diag<float><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, (const long long *)dPtr, dCol, dVal, 0, A.invDiag.as<float>());
    } else {
        k_sjds_fill<double, double2><<<blocks, 256, 0, c->stream>>>(n, (const long long *)dPtr, dCol, dVal, A.nslices, A.meta.as<int32_t>(),
                                                                   A.sliceOff.as<long long>(), A.val.as<double2>(), A.col.as<int2>());
        k_inv_diag<double><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, (const long long *)dPtr, dCol, dVal, 0, A.invDiag.as<double>());
    }
    c->launches += 2;
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

int avs_sell_from_stage(AvsContext *c, SellMatrix &A, int64_t n, int64_t nnz, const int32_t *dCount, const int32_t *dStageCol,
                        const double *dStageVal, long long stride, const double *dDiag, int precision,
                        long long rowBegin, long long rowEnd, const long long *haloIndex) {
    A.nnz = nnz;
    int64_t totalPairs = 0;
    int rc = sjdsLayout(c, A, n, dCount, precision, &totalPairs);
    if (rc || n == 0) return rc;
    unsigned blocks = (unsigned)((A.nslices * 32 + 255) / 256);
    if (precision == AVS_PRECISION_F32) {
        k_sjds_fill_stage<float, float2><<<blocks, 256, 0, c->stream>>>(n, dCount, dStageCol, dStageVal, stride, A.nslices, A.meta.as<int32_t>(),
                                                                       A.sliceOff.as<long long>(), A.val.as<float2>(), A.col.as<int2>(), rowBegin, rowEnd, haloIndex);
        k_inv_diag_from_array<float><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, dDiag, A.invDiag.as<float>());
    } else {
        k_sjds_fill_stage<double, double2><<<blocks, 256, 0, c->stream>>>(n, dCount, dStageCol, dStageVal, stride, A.nslices, A.meta.as<int32_t>(),
                                                                         A.sliceOff.as<long long>(), A.val.as<double2>(), A.col.as<int2>(), rowBegin, rowEnd, haloIndex);
        k_inv_diag_from_array<double><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, dDiag, A.invDiag.as<double>());
    }
    c->launches += 2;
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// ---- block reduction helpers -------------------------------------------------------------------
__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA; result valid in thread 0
__device__ __forceinline__ double blockSum(double v, double *sh /* [CG_THREADS/32] */) {
    v = warpSum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < CG_THREADS / 32; ++i) t += sh[i];
    return t;
}
// every thread of the CTA gets sum(parts[0..n)) in a fixed order (independent of gridDim of the consumer)
__device__ __forceinline__ double reduceParts(const double *parts, int n, double *sh /* [CG_THREADS/32 + 1] */) {
    double v = 0;
    for (int i = threadIdx.x; i < n; i += CG_THREADS) v += parts[i];
    v = warpSum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < CG_THREADS / 32; ++i) t += sh[i];
        sh[CG_THREADS / 32] = t;
    }
    __syncthreads();
    return sh[CG_THREADS / 32];
}

// device-resident CG state
struct CgScalars {
    double rho[2];        // absNew, double buffered by iteration parity
    double threshold;     // max(tol^2 |b|^2, min)
    double rhsNorm2;
    double residualNorm2;
    int done;             // 1 once |r|^2 < threshold (or breakdown / zero rhs)
    int iters;            // Eigen's `i`
    int breakdown;
    int zeroRhs;
};

// ---- SpMV: y = A x, fused partial dot x.y -------------------------------------------------------
// The slice loop shared by the stand-alone SpMV kernel and the persistent CG kernel.  NC = true gathers x through the
// read-only path (ld.global.nc): only legal when x is not written during the kernel's lifetime, i.e. NOT inside the
// persistent kernel, where p is rewritten every iteration by other CTAs (and by peers over NVLink).
// The matrix as the SpMV kernels see it (passed by value).
template <class T, class T2>
struct SjdsView {
    long long nrows, nslices;
    const long long *sliceOff;   // in pairs
    const int32_t *meta;
    const T2 *val2;
    const int2 *col2;
};
template <class T, class T2>
static SjdsView<T, T2> sjdsView(const SellMatrix &A) {
    SjdsView<T, T2> v;
    v.nrows = A.n; v.nslices = A.nslices;
    v.sliceOff = A.sliceOff.as<long long>(); v.meta = A.meta.as<int32_t>();
    v.val2 = A.val.as<T2>(); v.col2 = A.col.as<int2>();
    return v;
}

// One slice = 32 rows = one warp.  Returns the lane's contribution to x.y (row r of the slice, 0 when !DOT).
// NC = true gathers x through the read-only path (ld.global.nc): only legal when x is not written during the kernel's
// lifetime, i.e. NOT inside the persistent CG kernel, where p is rewritten every iteration by other CTAs (and by peers
// over NVLink).
template <class T, class T2, bool DOT, int SPMV_U, bool NC>
__device__ __forceinline__ double spmvOneSlice(long long s, int lane, const SjdsView<T, T2> &A, const T *x, T *y) {
    double dot = 0;
    {
        const int m = A.meta[s * 32 + lane];
        const int np = m & 0xff;
        const long long r = s * 32 + ((m >> 8) & 31);
        const long long base = A.sliceOff[s] + lane;
        const int maxnp = __shfl_sync(0xffffffffu, np, 0);
        T acc = 0;
        long long off = 0;
        // SPMV_U pair-steps per trip: offsets first (warp ballots, pure ALU), then all matrix loads, then all
        // gathers of x, then the FMAs -- so one lane keeps 2*SPMV_U independent loads in flight.
        for (int j0 = 0; j0 < maxnp; j0 += SPMV_U) {
            long long o[SPMV_U];
            bool act[SPMV_U];
#pragma unroll
            for (int u = 0; u < SPMV_U; ++u) {
                act[u] = (j0 + u) < np;
                o[u] = base + off;
                off += __popc(__ballot_sync(0xffffffffu, act[u]));
            }
            T2 v[SPMV_U];
            int2 cc[SPMV_U];
#pragma unroll
            for (int u = 0; u < SPMV_U; ++u) {
                v[u].x = 0; v[u].y = 0;
                cc[u].x = 0; cc[u].y = 0;
                if (act[u]) { v[u] = A.val2[o[u]]; cc[u] = A.col2[o[u]]; }
            }
            T xa[SPMV_U], xb[SPMV_U];
#pragma unroll
            for (int u = 0; u < SPMV_U; ++u) {
                xa[u] = 0; xb[u] = 0;
                if (act[u]) {
                    if (NC) { xa[u] = __ldg(x + cc[u].x); xb[u] = __ldg(x + cc[u].y); }
                    else { xa[u] = x[cc[u].x]; xb[u] = x[cc[u].y]; }
                }
            }
#pragma unroll
            for (int u = 0; u < SPMV_U; ++u) {
                acc += v[u].x * xa[u];
                acc += v[u].y * xb[u];
            }
        }
        if (r < A.nrows) {
            y[r] = acc;
            if (DOT) dot += (double)(NC ? __ldg(x + r) : x[r]) * (double)acc;
        }
    }
    return dot;
}
// static schedule: warp w of the grid takes slices w, w + W, w + 2W, ...
template <class T, class T2, bool DOT, int SPMV_U, bool NC>
__device__ __forceinline__ double spmvSlices(const SjdsView<T, T2> &A, const T *x, T *y) {
    const int lane = threadIdx.x & 31;
    const long long warpsTotal = ((long long)gridDim.x * CG_THREADS) >> 5;
    double dot = 0;
    for (long long s = ((long long)blockIdx.x * CG_THREADS + threadIdx.x) >> 5; s < A.nslices; s += warpsTotal)
        dot += spmvOneSlice<T, T2, DOT, SPMV_U, NC>(s, lane, A, x, y);
    return dot;
}
// ---- SpMV slice loop, variant "prefetch" -----------------------------------------------------------
// Same arithmetic as spmvOneSlice.  Every trip additionally asks L2 for the lines of the slice's value / column stream
// that the NEXT trip will read (prefetch.global.L2, no destination register), so the dependent chain of a trip is
// L2 latency (columns) + gather latency instead of DRAM latency + gather latency.
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class T, class T2, bool DOT, int SPMV_U, bool NC>
__device__ __forceinline__ double spmvOneSlicePf(long long s, int lane, const SjdsView<T, T2> &A, const T *x, T *y) {
    double dot = 0;
    const int m = A.meta[s * 32 + lane];
    const int np = m & 0xff;
    const long long r = s * 32 + ((m >> 8) & 31);
    const long long base = A.sliceOff[s] + lane;
    const int maxnp = __shfl_sync(0xffffffffu, np, 0);
    T acc = 0;
    long long off = 0;
    for (int j0 = 0; j0 < maxnp; j0 += SPMV_U) {
        long long o[SPMV_U];
        bool act[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            act[u] = (j0 + u) < np;
            o[u] = base + off;
            off += __popc(__ballot_sync(0xffffffffu, act[u]));
        }
        if (j0 + SPMV_U < maxnp) {   // the next trip's entries start at base - lane + off; <= 32 * SPMV_U of them
            const long long nx = base - lane + off;
            if (lane < 4 * SPMV_U) prefetchL2((const char *)(A.val2 + nx) + lane * 128);
            else if (lane < 6 * SPMV_U) prefetchL2((const char *)(A.col2 + nx) + (lane - 4 * SPMV_U) * 128);
        }
        T2 v[SPMV_U];
        int2 cc[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            v[u].x = 0; v[u].y = 0;
            cc[u].x = 0; cc[u].y = 0;
            if (act[u]) { v[u] = A.val2[o[u]]; cc[u] = A.col2[o[u]]; }
        }
        T xa[SPMV_U], xb[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            xa[u] = 0; xb[u] = 0;
            if (act[u]) {
                if (NC) { xa[u] = __ldg(x + cc[u].x); xb[u] = __ldg(x + cc[u].y); }
                else { xa[u] = x[cc[u].x]; xb[u] = x[cc[u].y]; }
            }
        }
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            acc += v[u].x * xa[u];
            acc += v[u].y * xb[u];
        }
    }
    if (r < A.nrows) {
        y[r] = acc;
        if (DOT) dot += (double)(NC ? __ldg(x + r) : x[r]) * (double)acc;
    }
    return dot;
}

// ---- SpMV slice loop, variant "ring" ---------------------------------------------------------------
// The matrix stream (value pairs + column pairs) of a slice is staged through shared memory by the warp that owns the slice:
// every lane copies ITS entry of a pair-step with cp.async (LDGSTS, 16 B values / 8 B columns) into a per-warp ring of
// SPMV_RING pair-steps and later reads back only what it copied itself -- no CTA barrier, no cross-lane hand-over, the
// warps stay independent.  The ring runs SPMV_RING pair-steps ahead of the gathers, so a trip waits for ONE latency (the
// gathers of x) instead of two dependent ones (columns, then gathers), and the in-flight matrix bytes need no registers.
__device__ __forceinline__ uint32_t smemAddr32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpAsync16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <class T2>
__device__ __forceinline__ void cpAsyncPair(T2 *dst, const T2 *src) {
    if (sizeof(T2) == 16) cpAsync16(dst, src); else cpAsync8(dst, src);
}
template <class T2, int SPMV_RING>
__host__ __device__ constexpr size_t spmvRingBytesPerWarp() { return (size_t)SPMV_RING * 32 * (sizeof(T2) + sizeof(int2)); }

// ring: this warp's SPMV_RING * 32 value pairs followed by SPMV_RING * 32 column pairs.  (m, sliceBase) = meta word of this
// lane and sliceOff[s], loaded by the caller one slice ahead.
template <class T, class T2, bool DOT, int SPMV_U, bool NC, int SPMV_RING>
__device__ __forceinline__ double spmvOneSliceRing(long long s, int lane, int m, long long sliceBase, const SjdsView<T, T2> &A, const T *x, T *y,
                                                   T2 *ringVal, int2 *ringCol) {
    static_assert(SPMV_RING % SPMV_U == 0, "ring must hold whole trips");
    constexpr int GROUPS = SPMV_RING / SPMV_U;   // cp.async groups in flight
    const int np = m & 0xff;
    const long long r = s * 32 + ((m >> 8) & 31);
    const int maxnp = __shfl_sync(0xffffffffu, np, 0);
    const T2 *gval = A.val2 + sliceBase + lane;
    const int2 *gcol = A.col2 + sliceBase + lane;
    int offF = 0, jf = 0;   // fetch side: entries of the slice before pair-step jf
    auto fetchGroup = [&]() {
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const int j = jf + u;
            const bool a = j < np;
            if (a) {
                const int slot = (j & (SPMV_RING - 1)) * 32 + lane;
                cpAsyncPair<T2>(ringVal + slot, gval + offF);
                cpAsync8(ringCol + slot, gcol + offF);
            }
            offF += __popc(__ballot_sync(0xffffffffu, a));
        }
        cpAsyncCommit();
        jf += SPMV_U;
    };
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) fetchGroup();
    T acc = 0;
    for (int jc = 0; jc < maxnp; jc += SPMV_U) {
        cpAsyncWait<GROUPS - 1>();
        int2 cc[SPMV_U];
        bool act[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            act[u] = (jc + u) < np;
            cc[u].x = 0; cc[u].y = 0;
            if (act[u]) cc[u] = ringCol[((jc + u) & (SPMV_RING - 1)) * 32 + lane];
        }
        T xa[SPMV_U], xb[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            xa[u] = 0; xb[u] = 0;
            if (act[u]) {
                if (NC) { xa[u] = __ldg(x + cc[u].x); xb[u] = __ldg(x + cc[u].y); }
                else { xa[u] = x[cc[u].x]; xb[u] = x[cc[u].y]; }
            }
        }
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            if (act[u]) {
                const T2 v = ringVal[((jc + u) & (SPMV_RING - 1)) * 32 + lane];
                acc += v.x * xa[u];
                acc += v.y * xb[u];
            }
        }
        fetchGroup();   // refills the slots just read (empty group once the slice is exhausted: keeps the group count uniform)
    }
    double dot = 0;
    if (r < A.nrows) {
        y[r] = acc;
        if (DOT) dot = (double)(NC ? __ldg(x + r) : x[r]) * (double)acc;
    }
    return dot;
}
// static schedule like spmvSlices; meta word and slice offset of the next slice are loaded while the current one runs
template <class T, class T2, bool DOT, int SPMV_U, bool NC, int SPMV_RING>
__device__ __forceinline__ double spmvSlicesRing(const SjdsView<T, T2> &A, const T *x, T *y, unsigned char *smemRing) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T2 *ringVal = (T2 *)(smemRing + (size_t)warp * spmvRingBytesPerWarp<T2, SPMV_RING>());
    int2 *ringCol = (int2 *)(ringVal + SPMV_RING * 32);
    const long long warpsTotal = ((long long)gridDim.x * CG_THREADS) >> 5;
    double dot = 0;
    long long s = ((long long)blockIdx.x * CG_THREADS + threadIdx.x) >> 5;
    if (s >= A.nslices) return 0;
    int m = A.meta[s * 32 + lane];
    long long sb = A.sliceOff[s];
    while (true) {
        const long long sn = s + warpsTotal;
        int mn = 0;
        long long sbn = 0;
        if (sn < A.nslices) { mn = A.meta[sn * 32 + lane]; sbn = A.sliceOff[sn]; }
        dot += spmvOneSliceRing<T, T2, DOT, SPMV_U, NC, SPMV_RING>(s, lane, m, sb, A, x, y, ringVal, ringCol);
        if (sn >= A.nslices) break;
        s = sn; m = mn; sb = sbn;
    }
    return dot;
}
template <class T, class T2, bool DOT, int SPMV_U, bool NC>
__device__ __forceinline__ double spmvSlicesPf(const SjdsView<T, T2> &A, const T *x, T *y) {
    const int lane = threadIdx.x & 31;
    const long long warpsTotal = ((long long)gridDim.x * CG_THREADS) >> 5;
    double dot = 0;
    for (long long s = ((long long)blockIdx.x * CG_THREADS + threadIdx.x) >> 5; s < A.nslices; s += warpsTotal)
        dot += spmvOneSlicePf<T, T2, DOT, SPMV_U, NC>(s, lane, A, x, y);
    return dot;
}

// dynamic schedule: warps grab chunks of PCG_CHUNK consecutive slices from a global counter
#define PCG_CHUNK 4
template <class T, class T2, bool DOT, int SPMV_U, bool NC>
__device__ __forceinline__ double spmvSlicesDynamic(const SjdsView<T, T2> &A, const T *x, T *y, unsigned long long *counter) {
    const int lane = threadIdx.x & 31;
    double dot = 0;
    while (true) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(counter, (unsigned long long)PCG_CHUNK);
        c = __shfl_sync(0xffffffffu, c, 0);
        if ((long long)c >= A.nslices) break;
        const long long e = min((long long)c + PCG_CHUNK, A.nslices);
        for (long long s = (long long)c; s < e; ++s)
            dot += spmvOneSlice<T, T2, DOT, SPMV_U, NC>(s, lane, A, x, y);
    }
    return dot;
}

template <class T, class T2, bool DOT, int SPMV_U>
__global__ void __launch_bounds__(CG_THREADS) k_spmv_sjds(const __grid_constant__ SjdsView<T, T2> A, const T *__restrict__ x, T *__restrict__ y,
                                                          double *__restrict__ parts, const CgScalars *sc) {
    if (sc && sc->done) return;
    __shared__ double sh[CG_THREADS / 32 + 1];
    double dot = spmvSlices<T, T2, DOT, SPMV_U, true>(A, x, y);
    if (DOT) {
        double t = blockSum(dot, sh);
        if (threadIdx.x == 0) parts[blockIdx.x] = t;
    }
}

template <class T, class T2, bool DOT, int SPMV_U>
__global__ void __launch_bounds__(CG_THREADS) k_spmv_sjds_pf(const __grid_constant__ SjdsView<T, T2> A, const T *__restrict__ x, T *__restrict__ y,
                                                             double *__restrict__ parts, const CgScalars *sc) {
    if (sc && sc->done) return;
    __shared__ double sh[CG_THREADS / 32 + 1];
    double dot = spmvSlicesPf<T, T2, DOT, SPMV_U, true>(A, x, y);
    if (DOT) {
        double t = blockSum(dot, sh);
        if (threadIdx.x == 0) parts[blockIdx.x] = t;
    }
}
template <class T, class T2, bool DOT, int SPMV_U, int SPMV_RING>
__global__ void __launch_bounds__(CG_THREADS) k_spmv_sjds_ring(const __grid_constant__ SjdsView<T, T2> A, const T *__restrict__ x, T *__restrict__ y,
                                                               double *__restrict__ parts, const CgScalars *sc) {
    if (sc && sc->done) return;
    extern __shared__ __align__(128) unsigned char dynSmem[];
    __shared__ double sh[CG_THREADS / 32 + 1];
    double dot = spmvSlicesRing<T, T2, DOT, SPMV_U, true, SPMV_RING>(A, x, y, dynSmem);
    if (DOT) {
        double t = blockSum(dot, sh);
        if (threadIdx.x == 0) parts[blockIdx.x] = t;
    }
}

// ---- SpMV, TMA-staged variant --------------------------------------------------------------------
// Same SJDS-32 matrix, same arithmetic per row, different data movement: the matrix is streamed by the TMA unit.
// One persistent CTA per SM (16 warps).  A chunk = 16 consecutive slices = 512 rows; its value pairs and column
// pairs are two contiguous byte ranges, fetched with two 1-D bulk async copies (cp.async.bulk ... mbarrier
// complete_tx, SASS UBLKCP) into one of two shared-memory stages while the warps work on the other stage.
// Warps read their slice out of shared memory (conflict-free LDS.128 / LDS.64), gather x through the read-only
// path, reduce p.Ap with warp shuffles.  Chunks larger than a stage (rows far denser than average) fall back to
// direct global loads.
#define TMA_WARPS 16
#define TMA_THREADS (TMA_WARPS * 32)
#define TMA_CAP_PAIRS 4608

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkG2S(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}

// one slice (32 rows) from `v`/`cc` (shared or global), returns this lane's row sum
template <class T, class T2>
__device__ __forceinline__ T sliceRowSum(const T2 *__restrict__ v, const int2 *__restrict__ cc, int np, int maxnp, int lane,
                                         const T *__restrict__ x) {
    T acc = 0;
    int off = lane;
    for (int j0 = 0; j0 < maxnp; j0 += 4) {
        int o[4];
        bool act[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            act[u] = (j0 + u) < np;
            o[u] = off;
            off += __popc(__ballot_sync(0xffffffffu, act[u]));
        }
        T2 vv[4];
        int2 c4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            vv[u].x = 0; vv[u].y = 0;
            c4[u].x = 0; c4[u].y = 0;
            if (act[u]) { vv[u] = v[o[u]]; c4[u] = cc[o[u]]; }
        }
        T xa[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            xa[u] = 0; xb[u] = 0;
            if (act[u]) { xa[u] = __ldg(x + c4[u].x); xb[u] = __ldg(x + c4[u].y); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc += vv[u].x * xa[u];
            acc += vv[u].y * xb[u];
        }
    }
    return acc;
}

template <class T, class T2, bool DOT>
__global__ void __launch_bounds__(TMA_THREADS, 1) k_spmv_tma(const __grid_constant__ SjdsView<T, T2> A, const T *__restrict__ x, T *__restrict__ y,
                                                             double *__restrict__ parts, const CgScalars *sc) {
    if (sc && sc->done) return;
    const long long nrows = A.nrows, nslices = A.nslices;
    const long long *__restrict__ sliceOff = A.sliceOff;
    const int32_t *__restrict__ meta = A.meta;
    const T2 *__restrict__ val2 = A.val2;
    const int2 *__restrict__ col2 = A.col2;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;
    const size_t stageBytes = (size_t)TMA_CAP_PAIRS * (sizeof(T2) + sizeof(int2));
    T2 *sval[2];
    int2 *scol[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        sval[s] = (T2 *)(smem + 128 + s * stageBytes);
        scol[s] = (int2 *)(smem + 128 + s * stageBytes + (size_t)TMA_CAP_PAIRS * sizeof(T2));
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbarInit(&bar[0], 1);
        mbarInit(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long nChunks = (nslices + TMA_WARPS - 1) / TMA_WARPS;
    auto chunkRange = [&](long long i, long long &s0, long long &s1, long long &p0, long long &np) {
        const long long c = (long long)blockIdx.x + i * gridDim.x;
        s0 = c * TMA_WARPS;
        s1 = min(s0 + (long long)TMA_WARPS, nslices);
        p0 = sliceOff[s0];
        np = sliceOff[s1] - p0;
    };
    auto issue = [&](long long i) {  // thread 0 only
        long long s0, s1, p0, np;
        chunkRange(i, s0, s1, p0, np);
        if (np > 0 && np <= TMA_CAP_PAIRS) {
            const int st = (int)(i & 1);
            mbarExpectTx(&bar[st], (uint32_t)(np * (sizeof(T2) + sizeof(int2))));
            bulkG2S(sval[st], val2 + p0, (uint32_t)(np * sizeof(T2)), &bar[st]);
            bulkG2S(scol[st], col2 + p0, (uint32_t)(np * sizeof(int2)), &bar[st]);
        }
    };
    long long myChunks = 0;
    if ((long long)blockIdx.x < nChunks) myChunks = (nChunks - 1 - blockIdx.x) / gridDim.x + 1;
    uint32_t phase[2] = {0, 0};
    double dot = 0;
    if (myChunks > 0 && tid == 0) issue(0);
    for (long long i = 0; i < myChunks; ++i) {
        if (i + 1 < myChunks && tid == 0) issue(i + 1);
        long long s0, s1, p0, np;
        chunkRange(i, s0, s1, p0, np);
        const int st = (int)(i & 1);
        const bool staged = np > 0 && np <= TMA_CAP_PAIRS;
        if (staged) {
            mbarWait(&bar[st], phase[st]);
            phase[st] ^= 1u;
        }
        const long long s = s0 + warp;
        if (s < s1) {
            const int m = meta[s * 32 + lane];
            const int npl = m & 0xff;
            const long long r = s * 32 + ((m >> 8) & 31);
            const int maxnp = __shfl_sync(0xffffffffu, npl, 0);
            const long long rel = sliceOff[s] - p0;
            T acc;
            if (staged) acc = sliceRowSum<T, T2>(sval[st] + rel, scol[st] + rel, npl, maxnp, lane, x);
            else acc = sliceRowSum<T, T2>(val2 + p0 + rel, col2 + p0 + rel, npl, maxnp, lane, x);
            if (r < nrows) {
                y[r] = acc;
                if (DOT) dot += (double)__ldg(x + r) * (double)acc;
            }
        }
        __syncthreads();  // every warp is done with stage `st` before iteration i+1 refills it (as chunk i+2)
    }
    if (DOT) {
        __shared__ double red[TMA_WARPS];
        double v = warpSum(dot);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (tid == 0) {
            double t = 0;
            for (int w = 0; w < TMA_WARPS; ++w) t += red[w];
            parts[blockIdx.x] = t;
        }
    }
}

// ---- CG vector kernels --------------------------------------------------------------------------
// init: r = b - t (t = A x0), p = invdiag r; partial b.b, r.r, r.p
template <class T>
__global__ void __launch_bounds__(CG_THREADS) k_cg_init(long long n, const T *__restrict__ b, const T *__restrict__ t, const T *__restrict__ invDiag,
                                                        T *__restrict__ r, T *__restrict__ p, double *parts, int nparts) {
    __shared__ double sh[CG_THREADS / 32 + 1];
    double bb = 0, rr = 0, rp = 0;
    for (long long i = (long long)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * CG_THREADS) {
        T bi = b[i];
        T ri = bi - t[i];
        T pi = invDiag[i] * ri;
        r[i] = ri;
        p[i] = pi;
        bb += (double)bi * (double)bi;
        rr += (double)ri * (double)ri;
        rp += (double)ri * (double)pi;
    }
    double s0 = blockSum(bb, sh), s1 = blockSum(rr, sh), s2 = blockSum(rp, sh);
    if (threadIdx.x == 0) {
        parts[blockIdx.x] = s0;
        parts[nparts + blockIdx.x] = s1;
        parts[2 * nparts + blockIdx.x] = s2;
    }
}
__global__ void __launch_bounds__(CG_THREADS) k_cg_init_scalars(const double *parts, int nparts, double tol, double tiny, CgScalars *sc) {
    __shared__ double sh[CG_THREADS / 32 + 1];
    double bb = reduceParts(parts, nparts, sh);
    double rr = reduceParts(parts + nparts, nparts, sh);
    double rp = reduceParts(parts + 2 * nparts, nparts, sh);
    if (threadIdx.x == 0) {
        sc->rhsNorm2 = bb;
        sc->residualNorm2 = rr;
        sc->rho[0] = rp;
        sc->rho[1] = rp;
        sc->iters = 0;
        sc->breakdown = 0;
        sc->zeroRhs = 0;
        sc->done = 0;
        double thr = tol * tol * bb;
        if (thr < tiny) thr = tiny;
        sc->threshold = thr;
        if (bb == 0) { sc->zeroRhs = 1; sc->done = 1; sc->residualNorm2 = 0; }  // Eigen: x = 0, 0 iterations
        else if (rr < thr) sc->done = 1;
    }
}
// x += alpha p; r -= alpha t; partial r.r and r.z with z = invdiag r
template <class T>
__global__ void __launch_bounds__(CG_THREADS) k_cg_update_xr(long long n, const T *__restrict__ p, const T *__restrict__ t,
                                                             const T *__restrict__ invDiag, T *__restrict__ x, T *__restrict__ r,
                                                             const double *ptParts, int nptParts, double *parts, int nparts,
                                                             CgScalars *sc, int parity) {
    if (sc->done) return;
    __shared__ double sh[CG_THREADS / 32 + 1];
    const double pt = reduceParts(ptParts, nptParts, sh);
    const double alphaD = sc->rho[parity] / pt;
    if (!isfinite(alphaD)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->breakdown = 1;  // becomes `done` in k_cg_update_p
        // fall through with a harmless alpha so every CTA takes the same path
    }
    const T alpha = isfinite(alphaD) ? (T)alphaD : (T)0;
    double rr = 0, rz = 0;
    for (long long i = (long long)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * CG_THREADS) {
        T pi = p[i];
        T ri = r[i] - alpha * t[i];
        x[i] += alpha * pi;
        r[i] = ri;
        T zi = invDiag[i] * ri;
        rr += (double)ri * (double)ri;
        rz += (double)ri * (double)zi;
    }
    double s0 = blockSum(rr, sh), s1 = blockSum(rz, sh);
    if (threadIdx.x == 0) {
        parts[blockIdx.x] = s0;
        parts[nparts + blockIdx.x] = s1;
    }
}
// convergence test, then p = z + beta p
template <class T>
__global__ void __launch_bounds__(CG_THREADS) k_cg_update_p(long long n, const T *__restrict__ r, const T *__restrict__ invDiag, T *__restrict__ p,
                                                            const double *parts, int nparts, CgScalars *sc, int parity) {
    if (sc->done) return;
    __shared__ double sh[CG_THREADS / 32 + 1];
    const double rr = reduceParts(parts, nparts, sh);
    const double rz = reduceParts(parts + nparts, nparts, sh);
    const bool stop = (rr < sc->threshold) || sc->breakdown;
    const double beta = rz / sc->rho[parity];
    __syncthreads();
    // CTA 0 publishes the new state.  No other CTA reads residualNorm2 / rho[parity^1] / iters in this
    // kernel; `done` is only ever written when `stop` holds, and then every CTA returns without touching
    // p whether it saw the old or the new value of `done`.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->residualNorm2 = rr;
        sc->rho[parity ^ 1] = rz;
        if (stop) sc->done = 1;       // Eigen: break before ++i
        else sc->iters += 1;
    }
    if (stop) return;
    const T b = (T)beta;
    for (long long i = (long long)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * CG_THREADS)
        p[i] = invDiag[i] * r[i] + b * p[i];
}

// ---- persistent CG kernel --------------------------------------------------------------------------
// One cooperative launch runs the whole iteration loop: SpMV (+ p.Ap) | grid barrier | x,r update (+ r.r, r.z) | grid
// barrier | convergence test, p update | grid barrier.  No kernel boundaries, no host polling, no iterations launched past
// convergence: at ~1 M rows the three launches + gaps of the per-launch path cost as much as the SpMV itself, and at the
// 1/8-of-C3 partitions of the 8-GPU run they are the floor (SURVEY.md section 7 "CG latency floor").  Every CTA finishes
// the per-CTA partial sums itself, in the same fixed order, so all CTAs (and, multi-GPU, all ranks) hold bit-identical
// scalars and take the same branches -- the Eigen loop of the header comment, unchanged.
// Multi-GPU (peer-memory mode): the two scalar all-reduces and the halo exchange happen INSIDE this kernel over NVLink --
// CTA 0 stores this rank's sums into every peer's mailbox and every CTA spins on its own mailbox; after the p update the
// CTAs push the rows their peers need straight into the peers' halo slots, the last CTA to finish raises the peers' flags,
// and the next SpMV waits for the flags of the ranks it receives from.
// Memory visibility between phases comes from the barrier (release fence + atomic, acquire spin + fence: the pattern of
// cooperative_groups' grid sync); p is gathered with ordinary loads, never ld.global.nc.
struct PcgState {                 // device-resident, copied back by the host after every launch
    unsigned long long seqPush, seqReduce;
    unsigned long long phaseNs[3];    // accumulated %globaltimer time of the SpMV / x,r / p phases (CTA 0, barrier to barrier)
    unsigned long long spmvPhases;    // SpMV phases executed
    unsigned int barrier;             // grid barrier arrival counter (zeroed before every launch)
    unsigned int pushArrivals;        // halo push arrival counter (zeroed before every launch)
    unsigned long long work[2];       // dynamic SpMV schedule: next slice, double buffered by iteration parity (zeroed likewise)
    int abort;                        // a spin loop timed out (peer died / co-residency broken): everybody leaves
    int pad;
};
#define PCG_TIMEOUT_NS 8000000000ull

__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ldAcquireGpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ldVolatileU64(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
__device__ __forceinline__ unsigned ldVolatileU32(const unsigned *p) { return *(const volatile unsigned *)p; }

// All CTAs of the (co-resident) grid.  Returns false when the launch is being aborted.
__device__ __forceinline__ bool gridBarrier(PcgState *st, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&st->barrier, 1u);
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldAcquireGpu(&st->barrier) < target) {
            if ((++spins & 0x3ff) == 0) {
                if (*(volatile int *)&st->abort) break;
                unsigned long long now = globalTimerNs();
                if (t0 == 0) t0 = now;
                else if (now - t0 > PCG_TIMEOUT_NS) { *(volatile int *)&st->abort = 1; break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *(volatile int *)&st->abort == 0;
}

// v[0..COUNT) holds this rank's sums (the same bits in every CTA); on return it holds the sum over ranks, in rank order.
template <int COUNT>
__device__ __forceinline__ bool rankSum(double (&v)[COUNT], const PcgDist &D, unsigned long long seq, PcgState *st) {
    const int par = (int)(seq & 1ull);
    P2PHeader *mine = (P2PHeader *)D.peerRegion[D.myRank];
    if (blockIdx.x == 0 && threadIdx.x < D.P) {
        P2PHeader *peer = (P2PHeader *)D.peerRegion[threadIdx.x];
#pragma unroll
        for (int q = 0; q < COUNT; ++q) *(volatile double *)&peer->mail[par][D.myRank][q] = v[q];
        __threadfence_system();
        *(volatile unsigned long long *)&peer->flag[par][D.myRank] = seq;
    }
    if (threadIdx.x < D.P) {
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldVolatileU64(&mine->flag[par][threadIdx.x]) != seq) {
            if ((++spins & 0x3ff) == 0) {
                if (*(volatile int *)&st->abort) break;
                unsigned long long now = globalTimerNs();
                if (t0 == 0) t0 = now;
                else if (now - t0 > PCG_TIMEOUT_NS) { *(volatile int *)&st->abort = 1; break; }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (*(volatile int *)&st->abort) return false;
#pragma unroll
    for (int q = 0; q < COUNT; ++q) {
        double t = 0;
        for (int r = 0; r < D.P; ++r) t += *(const volatile double *)&mine->mail[par][r][q];
        v[q] = t;
    }
    return true;
}

// After the p update (and the barrier behind it): store the rows my peers need into their halo slots; the last CTA to
// finish raises pushFlag[myRank] = seq in every peer I send to.
template <class T>
__device__ __forceinline__ void pushHalo(const T *p, const PcgDist &D, unsigned long long seq, PcgState *st, unsigned &pushTarget) {
    for (long long i = (long long)blockIdx.x * CG_THREADS + threadIdx.x; i < D.nSend; i += (long long)gridDim.x * CG_THREADS) {
        const int2 dst = D.sendDst[i];
        T *peerP = (T *)((char *)D.peerRegion[dst.x] + P2P_HEADER_BYTES);
        *(volatile T *)(peerP + dst.y) = p[D.sendIdx[i] - D.rowBegin];
    }
    __threadfence_system();
    __syncthreads();
    pushTarget += gridDim.x;
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(&st->pushArrivals, 1u);
        if (prev + 1 == pushTarget) {   // every CTA's stores are ordered before this point (fence + atomic chain)
            __threadfence_system();
            for (int q = 0; q < D.P; ++q)
                if ((D.sendMask >> q) & 1u) {
                    P2PHeader *peer = (P2PHeader *)D.peerRegion[q];
                    *(volatile unsigned long long *)&peer->pushFlag[D.myRank] = seq;
                }
        }
    }
}
__device__ __forceinline__ bool waitHalo(const PcgDist &D, unsigned long long seq, PcgState *st) {
    if (threadIdx.x < D.P && ((D.recvMask >> threadIdx.x) & 1u)) {
        const P2PHeader *mine = (const P2PHeader *)D.peerRegion[D.myRank];
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldVolatileU64(&mine->pushFlag[threadIdx.x]) < seq) {
            if ((++spins & 0x3ff) == 0) {
                if (*(volatile int *)&st->abort) break;
                unsigned long long now = globalTimerNs();
                if (t0 == 0) t0 = now;
                else if (now - t0 > PCG_TIMEOUT_NS) { *(volatile int *)&st->abort = 1; break; }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    return *(volatile int *)&st->abort == 0;
}

template <class T, class T2>
struct PcgArgs {
    long long n;
    SjdsView<T, T2> M;
    const T *invDiag;
    T *x, *r, *p, *t;
    double *parts;       // 3 * gridDim.x: p.Ap | r.r | r.z
    CgScalars *sc;
    PcgState *st;
    int itLimit;         // stop once sc->iters reaches this
    int pushFirst;       // 1: the p vector of this launch's first SpMV has not been pushed to the peers yet
    int dynamic;         // 1: warps take slices from a global counter instead of a fixed stride
    PcgDist D;
};

template <class T, class T2, int SPMV_U, int MINB>
__global__ void __launch_bounds__(CG_THREADS, MINB) k_cg_persistent(const __grid_constant__ PcgArgs<T, T2> A) {
    __shared__ double sh[CG_THREADS / 32 + 1];
    __shared__ unsigned long long tm[4];   // thread 0 of CTA 0: phase start stamp + 3 accumulators
    __shared__ double rrLast;
    CgScalars *sc = A.sc;
    PcgState *st = A.st;
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
    if (sc->done) {             // uniform: written before the launch.  Nothing was exchanged: hand the counters back unchanged
        if (timer) { st->seqPush = A.D.seqPush; st->seqReduce = A.D.seqReduce; }
        return;
    }
    const bool dist = A.D.P > 1;
    const int G = gridDim.x;
    double rho = sc->rho[0];
    int iters = sc->iters;
    int k = 0;                  // iterations started in this launch: sequence numbers derive from it
    unsigned barTarget = 0, pushTarget = 0;
    bool stop = false, breakdown = false;
    if (timer) { tm[1] = tm[2] = tm[3] = 0; rrLast = sc->residualNorm2; }
    // sequence numbers: reduce #(2k+1), #(2k+2) in iteration k; push #(pushFirst + k) feeds iteration k's SpMV
    const unsigned long long seqPush0 = A.D.seqPush + (unsigned long long)A.pushFirst;

    if (dist && A.pushFirst) pushHalo<T>(A.p, A.D, seqPush0, st, pushTarget);
    while (iters < A.itLimit) {
        if (timer) tm[0] = globalTimerNs();
        // ---- t = A p, partial p.t
        if (dist && !waitHalo(A.D, seqPush0 + k, st)) break;
        {
            double dot = A.dynamic
                ? spmvSlicesDynamic<T, T2, true, SPMV_U, false>(A.M, A.p, A.t, &st->work[k & 1])
                : spmvSlices<T, T2, true, SPMV_U, false>(A.M, A.p, A.t);
            double s0 = blockSum(dot, sh);
            if (threadIdx.x == 0) A.parts[blockIdx.x] = s0;
        }
        if (!gridBarrier(st, barTarget)) break;
        if (timer) { unsigned long long now = globalTimerNs(); tm[1] += now - tm[0]; tm[0] = now; st->work[(k + 1) & 1] = 0; }
        double pt[1] = {reduceParts(A.parts, G, sh)};
        if (dist && !rankSum<1>(pt, A.D, A.D.seqReduce + 2ull * k + 1, st)) break;
        const double alphaD = rho / pt[0];
        breakdown = !isfinite(alphaD);
        const T alpha = breakdown ? (T)0 : (T)alphaD;
        // ---- x += alpha p, r -= alpha t, partial r.r and r.z
        {
            // two elements per thread and trip (one 128-bit access per vector for fp64): more bytes in flight per thread
            double rr = 0, rz = 0;
            const long long gid = (long long)blockIdx.x * CG_THREADS + threadIdx.x, gstride = (long long)G * CG_THREADS;
            const long long n2 = A.n >> 1;
            for (long long j = gid; j < n2; j += gstride) {
                const T2 pv = ((const T2 *)A.p)[j], tv = ((const T2 *)A.t)[j], dv = ((const T2 *)A.invDiag)[j];
                T2 rv = ((T2 *)A.r)[j], xv = ((T2 *)A.x)[j];
                rv.x -= alpha * tv.x; rv.y -= alpha * tv.y;
                xv.x += alpha * pv.x; xv.y += alpha * pv.y;
                ((T2 *)A.x)[j] = xv;
                ((T2 *)A.r)[j] = rv;
                rr += (double)rv.x * (double)rv.x + (double)rv.y * (double)rv.y;
                rz += (double)rv.x * (double)(T)(dv.x * rv.x) + (double)rv.y * (double)(T)(dv.y * rv.y);
            }
            if ((A.n & 1) && gid == 0) {
                const long long i = A.n - 1;
                T ri = A.r[i] - alpha * A.t[i];
                A.x[i] += alpha * A.p[i];
                A.r[i] = ri;
                rr += (double)ri * (double)ri;
                rz += (double)ri * (double)(T)(A.invDiag[i] * ri);
            }
            double s0 = blockSum(rr, sh), s1 = blockSum(rz, sh);
            if (threadIdx.x == 0) { A.parts[G + blockIdx.x] = s0; A.parts[2 * G + blockIdx.x] = s1; }
        }
        if (!gridBarrier(st, barTarget)) break;
        if (timer) { unsigned long long now = globalTimerNs(); tm[2] += now - tm[0]; tm[0] = now; }
        double rs[2];
        rs[0] = reduceParts(A.parts + G, G, sh);
        rs[1] = reduceParts(A.parts + 2 * G, G, sh);
        if (dist && !rankSum<2>(rs, A.D, A.D.seqReduce + 2ull * k + 2, st)) break;
        ++k;
        if (timer) rrLast = rs[0];
        stop = (rs[0] < sc->threshold) || breakdown;
        if (stop) break;                       // Eigen: break before ++i
        // ---- p = z + beta p
        const T beta = (T)(rs[1] / rho);
        rho = rs[1];
        ++iters;
        {
            const long long gid = (long long)blockIdx.x * CG_THREADS + threadIdx.x, gstride = (long long)G * CG_THREADS;
            const long long n2 = A.n >> 1;
            for (long long j = gid; j < n2; j += gstride) {
                const T2 dv = ((const T2 *)A.invDiag)[j], rv = ((const T2 *)A.r)[j];
                T2 pv = ((T2 *)A.p)[j];
                pv.x = dv.x * rv.x + beta * pv.x;
                pv.y = dv.y * rv.y + beta * pv.y;
                ((T2 *)A.p)[j] = pv;
            }
            if ((A.n & 1) && gid == 0) { const long long i = A.n - 1; A.p[i] = A.invDiag[i] * A.r[i] + beta * A.p[i]; }
        }
        if (!gridBarrier(st, barTarget)) break;
        if (dist) pushHalo<T>(A.p, A.D, seqPush0 + k, st, pushTarget);
        if (timer) tm[3] += globalTimerNs() - tm[0];
    }
    if (timer) {
        sc->rho[0] = rho;
        sc->rho[1] = rho;
        sc->iters = iters;
        sc->residualNorm2 = rrLast;
        if (breakdown) sc->breakdown = 1;
        if (stop) sc->done = 1;
        // k counts completed (SpMV, x/r) phase pairs; a stopped iteration did both reduces but no p update / push
        st->seqReduce = A.D.seqReduce + 2ull * k;
        st->seqPush = seqPush0 + (stop ? k - 1 : k);
        st->phaseNs[0] += tm[1];
        st->phaseNs[1] += tm[2];
        st->phaseNs[2] += tm[3];
        st->spmvPhases += k;
    }
}

// acquire-release fences (the v2 kernel's synchronisation is release/acquire, it never needs sequential consistency)
__device__ __forceinline__ void fenceGpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fenceSys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// ---- persistent CG kernel, version 2 (default) ---------------------------------------------------------------------------------
// Same Eigen loop, same three phases per iteration, but every synchronisation point is ONE wait instead of a chain:
//  * syncs A and B (after the SpMV: p.Ap; after the x,r update: r.r, r.z) merge the grid barrier, the reduction of the per-CTA
//    partial sums and the scalar all-reduce over the ranks.  Every CTA stores its partials and bumps an arrival counter; the LAST CTA
//    to arrive adds the partials in a fixed order and stores the rank's sums straight into the mailbox of every rank (its own
//    included) as flag-in-data words (32 payload bits + 32-bit sequence number per 8-byte store: no fence between data and flag, the
//    idea of NCCL's LL protocol).  All CTAs of all ranks then spin on their own rank's mailbox until the P contributions of this
//    sequence number are there and add them in rank order: bit-identical scalars everywhere, and the arrival of the rank's own
//    contribution doubles as the grid barrier.  v1 paid barrier + 592 redundant reductions + fence.sys + flag + spin in sequence.
//  * p is double buffered (p_{k+1} = z + beta p_k is written to the other buffer), so the rows a peer needs are RECOMPUTED from r,
//    invdiag and p_k by whichever thread gets there first and stored into the peer's halo slots over NVLink before the owner's own
//    sweep -- no barrier between "p updated" and "halo pushed".  Sync C (last arriver raises localReady and the peers' pushFlag)
//    is not waited for at all: the next SpMV waits for localReady, runs the slices that read no halo slot, and only then waits for
//    the peers' pushFlag and runs the boundary slices (about 1 % of the slices), so the NVLink latency hides under the interior SpMV.
// Single GPU: the same kernel with P = 1 (mailbox and flags in a local header, no remote store).
struct Pcg2State {                // device-resident, copied back by the host after every launch
    unsigned long long seqPush, seqReduce;
    unsigned long long phaseNs[3];    // accumulated %globaltimer time of the SpMV / x,r / p phases (thread 0 of CTA 0)
    unsigned long long spmvPhases;
    unsigned int arrive[4];           // arrival counters of syncs A, B, C (zeroed before every launch)
    int abort;
    int pad;
};

// The canonical reduction of the per-CTA partial sums (every CTA of every rank uses this one order, and so does the CTA that posts
// the rank's sum to the peers: bit-identical scalars everywhere).  Loads go straight to L2 (other CTAs wrote the partials).
__device__ __forceinline__ double reducePartsCg(const double *parts, int n, double *sh /* [CG_THREADS/32 + 1] */) {
    double v = 0;
    for (int i = threadIdx.x; i < n; i += CG_THREADS) v += __ldcg(parts + i);
    v = warpSum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < CG_THREADS / 32; ++i) t += sh[i];
        sh[CG_THREADS / 32] = t;
    }
    __syncthreads();
    return sh[CG_THREADS / 32];
}
template <class ST>
__device__ __forceinline__ bool spinTimedOut(unsigned &spins, unsigned long long &t0, ST *st) {
    if ((++spins & 0x3ff) != 0) return false;
    if (*(volatile int *)&st->abort) return true;
    const unsigned long long now = globalTimerNs();
    if (t0 == 0) t0 = now;
    else if (now - t0 > PCG_TIMEOUT_NS) { *(volatile int *)&st->abort = 1; return true; }
    return false;
}

// v[0..COUNT): this CTA's partial sums (valid in thread 0).  On return: the sums over all CTAs of all ranks, the same bits in every
// thread of every CTA of every rank.  Also the grid barrier between two phases.
//   every CTA publishes its partials and bumps the arrival counter, waits for the counter (the local barrier: one release fence, one
//   acquire fence per CTA, like a plain grid barrier) and reduces the partials itself;
//   P > 1: the LAST CTA to arrive reduces first and stores the rank's sums into the mailboxes of the PEERS as flag-in-data words, so
//   the NVLink trip overlaps the other CTAs' own reduction; everybody then waits for the peers' words and adds the P contributions in
//   rank order.  shw: 4 * P2P_MAX_RANKS words.
template <int COUNT>
__device__ __forceinline__ bool syncSum(double (&v)[COUNT], double *parts, unsigned *counter, unsigned &target, const PcgDist &D, P2PHeader *mine,
                                        unsigned long long seq, Pcg2State *st, double *sh, unsigned *shw, int *shLast) {
    const int G = gridDim.x;
    const int par = (int)(seq & 1ull);
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int words = (D.P - 1) * COUNT * 2;
    target += (unsigned)G;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < COUNT; ++q) parts[q * G + blockIdx.x] = v[q];
    }
    __syncthreads();   // every thread's writes of this phase precede thread 0's release below
    if (threadIdx.x == 0) {
        fenceGpu();
        const unsigned prev = atomicAdd(counter, 1u);
        if (D.P > 1) *shLast = (prev + 1u == target) ? 1 : 0;
    }
    double tot[COUNT];
    if (D.P > 1) {
        __syncthreads();
        if (*shLast) {   // CTA-uniform: every CTA of this rank has arrived, its partials are in L2
            if (threadIdx.x == 0) fenceGpu();
#pragma unroll
            for (int q = 0; q < COUNT; ++q) tot[q] = reducePartsCg(parts + q * G, G, sh);
            if (threadIdx.x < words) {   // one thread per (peer, value, half); my slot in the peer's mailbox = my rank
                int peerRank = threadIdx.x / (COUNT * 2);
                if (peerRank >= D.myRank) ++peerRank;
                const int q = (threadIdx.x >> 1) % COUNT, half = threadIdx.x & 1;
                unsigned long long bits = 0;
#pragma unroll
                for (int i = 0; i < COUNT; ++i)
                    if (i == q) bits = (unsigned long long)__double_as_longlong(tot[i]);
                const unsigned long long word = (half ? (bits >> 32) : (bits & 0xffffffffull)) | tag;
                *(volatile unsigned long long *)&((P2PHeader *)D.peerRegion[peerRank])->ll[par][D.myRank][q][half] = word;
            }
        }
    }
    if (threadIdx.x == 0) {   // local barrier
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldVolatileU32(counter) < target)
            if (spinTimedOut(spins, t0, st)) break;
        fenceGpu();
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < COUNT; ++q) tot[q] = reducePartsCg(parts + q * G, G, sh);
    if (D.P > 1) {
        if (threadIdx.x < words) {   // thread = (sender slot, value, half), senders in rank order without me
            int sender = threadIdx.x / (COUNT * 2);
            if (sender >= D.myRank) ++sender;
            const int q = (threadIdx.x >> 1) % COUNT, half = threadIdx.x & 1;
            const unsigned long long *slot = &mine->ll[par][sender][q][half];
            unsigned long long w = ldVolatileU64(slot), t0 = 0;
            unsigned spins = 0;
            while ((w & 0xffffffff00000000ull) != tag) {
                if (spinTimedOut(spins, t0, st)) break;
                w = ldVolatileU64(slot);
            }
            shw[threadIdx.x] = (unsigned)(w & 0xffffffffull);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < COUNT; ++q) {
            double t = 0;
            for (int r = 0; r < D.P; ++r) {
                if (r == D.myRank) { t += tot[q]; continue; }
                const int slotIdx = ((r > D.myRank ? r - 1 : r) * COUNT + q) * 2;
                const unsigned lo = shw[slotIdx], hi = shw[slotIdx + 1];
                t += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
            }
            tot[q] = t;
        }
        __syncthreads();   // shw / shLast are reused by the next sync
    }
    if (*(volatile int *)&st->abort) return false;
#pragma unroll
    for (int q = 0; q < COUNT; ++q) v[q] = tot[q];
    return true;
}

// the one expression for p_{k+1}: the owner's sweep and the halo push must produce the same bits
__device__ __forceinline__ double pNext(double d, double r, double beta, double p) { return __fma_rn(beta, p, __dmul_rn(d, r)); }
__device__ __forceinline__ float pNext(float d, float r, float beta, float p) { return __fmaf_rn(beta, p, __fmul_rn(d, r)); }

template <class T, class T2>
struct Pcg2Args {
    long long n;
    SjdsView<T, T2> M;
    const T *invDiag;
    T *x, *r, *t;
    T *p0;               // p buffer 0 (owned rows, then halo slots); buffer 1 = p0 + pDelta bytes.  Iteration k of a launch reads
    long long pDelta;    // buffer (parity0 + k) & 1.  Two pointers in an ARRAY would be selected with a register-indexed constant-bank
                         // load that ptxas re-issues inside the SpMV loop in front of every gather (measured: +14 % SpMV time).
                         // Nor may the pointer be laundered through inline asm: the gathers then become GENERIC loads (LD.E instead
                         // of LDG.E), the same 14 %.  tests/test_sass.py checks both.
    double *parts;       // 3 * gridDim.x
    CgScalars *sc;
    Pcg2State *st;
    P2PHeader *mine;     // this rank's header (inside the peer-mapped region when P > 1)
    int itLimit;
    int pushFirst;       // 1: first launch of a solve -- p_0 (from k_cg_init, buffer 0) has not been pushed / flagged yet
    int parity0;         // buffer that holds the current p at kernel entry
    PcgDist D;
};

template <class T>
__device__ __forceinline__ T *peerP(const PcgDist &D, int rank, int buf) {
    return (T *)((char *)D.peerRegion[rank] + P2P_HEADER_BYTES + (size_t)buf * D.pStrideBytes);
}

// Sync C: every CTA has written its rows of the new p (and its share of the halo pushes).  Nobody waits here: the next SpMV waits
// for the arrival counter (local part) and for the peers' pushFlag (halo part).
__device__ __forceinline__ void arrivePush(unsigned *counter, unsigned &target, const PcgDist &D, unsigned long long seq, bool pushedRemote) {
    if (pushedRemote) fenceSys();
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        if (pushedRemote) fenceSys(); else fenceGpu();
        const unsigned prev = atomicAdd(counter, 1u);
        if (D.P > 1 && prev + 1u == target) {   // every CTA's pushes are ordered before this point (fence + atomic chain)
            fenceSys();
            for (int q = 0; q < D.P; ++q)
                if ((D.sendMask >> q) & 1u) *(volatile unsigned long long *)&((P2PHeader *)D.peerRegion[q])->pushFlag[D.myRank] = seq;
        }
    }
}
__device__ __forceinline__ bool waitLocalReady(const unsigned *counter, unsigned target, Pcg2State *st) {
    if (threadIdx.x == 0) {
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldVolatileU32(counter) < target)
            if (spinTimedOut(spins, t0, st)) break;
        fenceGpu();
    }
    __syncthreads();
    return *(volatile int *)&st->abort == 0;
}
__device__ __forceinline__ bool waitHaloReady(const PcgDist &D, P2PHeader *mine, unsigned long long seq, Pcg2State *st) {
    if (threadIdx.x < D.P && ((D.recvMask >> threadIdx.x) & 1u)) {
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ldVolatileU64(&mine->pushFlag[threadIdx.x]) < seq)
            if (spinTimedOut(spins, t0, st)) break;
        fenceSys();
    }
    __syncthreads();
    return *(volatile int *)&st->abort == 0;
}

// A/B variant (AVS_SPMV_MODE=call): the slice loop as a function of its own, scheduled and register-allocated independently of the
// synchronisation code around it.  Measured 13 % slower than the inlined loop (0.527 vs 0.464 ms per SpMV phase at C3), so it is not
// the default; the instability it was meant to cure turned out to be the p-buffer pointer (see Pcg2Args::p0).
template <class T, class T2>
__device__ __noinline__ double spmvSliceCall(long long s, int lane, long long nrows, const long long *sliceOff, const int32_t *meta, const T2 *val2,
                                             const int2 *col2, const T *p, T *t) {
    SjdsView<T, T2> M;   // rebuilt from scalars: a by-reference struct would be read through local memory on every use
    M.nrows = nrows; M.nslices = 0; M.sliceOff = sliceOff; M.meta = meta; M.val2 = val2; M.col2 = col2;
    return spmvOneSlice<T, T2, true, 4, false>(s, lane, M, p, t);
}

// SPMV_MODE: 0 register-staged slice loop, 1 + L2 prefetch, 2 per-warp cp.async ring (8 pair-steps, trips of 4)
template <class T, class T2, int SPMV_MODE>
__device__ __forceinline__ double pcgSlice(long long s, int lane, const SjdsView<T, T2> &M, const T *p, T *t, unsigned char *ring) {
    if (SPMV_MODE == 2) {
        const int warp = threadIdx.x >> 5;
        T2 *ringVal = (T2 *)(ring + (size_t)warp * spmvRingBytesPerWarp<T2, 8>());
        int2 *ringCol = (int2 *)(ringVal + 8 * 32);
        return spmvOneSliceRing<T, T2, true, 4, false, 8>(s, lane, M.meta[s * 32 + lane], M.sliceOff[s], M, p, t, ringVal, ringCol);
    }
    if (SPMV_MODE == 1) return spmvOneSlicePf<T, T2, true, 4, false>(s, lane, M, p, t);
    if (SPMV_MODE == 3) return spmvSliceCall<T, T2>(s, lane, M.nrows, M.sliceOff, M.meta, M.val2, M.col2, p, t);   // A/B: measured 13 % slower
    return spmvOneSlice<T, T2, true, 4, false>(s, lane, M, p, t);
}

template <class T, class T2, int SPMV_MODE, int MINB>
__global__ void __launch_bounds__(CG_THREADS, MINB) k_cg_persistent2(const __grid_constant__ Pcg2Args<T, T2> A) {
    extern __shared__ __align__(128) unsigned char dynSmem[];
    __shared__ double sh[CG_THREADS / 32 + 1];
    __shared__ unsigned shw[P2P_MAX_RANKS * 4];
    __shared__ int shLast;
    __shared__ unsigned long long tm[4];   // thread 0 of CTA 0: phase start stamp + 3 accumulators
    __shared__ double rrLast;
    CgScalars *sc = A.sc;
    Pcg2State *st = A.st;
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
    // sequence numbers: reduce #(2k+1), #(2k+2) in iteration k of this launch; push #(seqPush0 + k) feeds iteration k's SpMV
    const unsigned long long seqPush0 = A.D.seqPush + (unsigned long long)A.pushFirst;
    if (sc->done) {             // uniform: written before the launch.  Report the counters unchanged (nothing was exchanged).
        if (timer) { st->seqPush = A.D.seqPush; st->seqReduce = A.D.seqReduce; }
        return;
    }
    const bool dist = A.D.P > 1;
    const int G = gridDim.x;
    const int lane = threadIdx.x & 31;
    const long long gwarp = ((long long)blockIdx.x * CG_THREADS + threadIdx.x) >> 5, warpsTotal = ((long long)G * CG_THREADS) >> 5;
    const long long gid = (long long)blockIdx.x * CG_THREADS + threadIdx.x, gstride = (long long)G * CG_THREADS;
    const long long n2 = A.n >> 1;
    const bool pushes = dist && (long long)blockIdx.x * CG_THREADS < A.D.nSend;   // this CTA stores into peer memory
    double rho = sc->rho[0];
    int iters = sc->iters;
    int k = 0;
    unsigned tgtA = 0, tgtB = 0, tgtC = 0;
    bool stop = false, breakdown = false;
    if (timer) { tm[1] = tm[2] = tm[3] = 0; rrLast = sc->residualNorm2; }

    if (A.pushFirst) {   // p_0 sits in buffer parity0 (written by k_cg_init): copy the rows the peers need, then flag
        const T *pc = (const T *)((const char *)A.p0 + (A.parity0 ? A.pDelta : 0));
        if (dist)
            for (long long i = gid; i < A.D.nSend; i += gstride) {
                const int2 dst = A.D.sendDst[i];
                *(volatile T *)(peerP<T>(A.D, dst.x, A.parity0) + dst.y) = pc[A.D.sendIdx[i] - A.D.rowBegin];
            }
        arrivePush(&st->arrive[2], tgtC, A.D, seqPush0, pushes);
    }
    while (iters < A.itLimit) {
        const int cur = (A.parity0 + k) & 1;
        T *pc = (T *)((char *)A.p0 + (cur ? A.pDelta : 0)), *pn = (T *)((char *)A.p0 + (cur ? 0 : A.pDelta));
        if (timer) tm[0] = globalTimerNs();
        // ---- t = A p, partial p.t: slices that read no halo slot first, the boundary slices after the peers' values have landed
        double pt[1];
        {
            if (!waitLocalReady(&st->arrive[2], tgtC, st)) break;
            double dot = 0;
            if (!dist) {
                if (SPMV_MODE == 2) dot = spmvSlicesRing<T, T2, true, 4, false, 8>(A.M, pc, A.t, dynSmem);
                else
                    for (long long s = gwarp; s < A.M.nslices; s += warpsTotal) dot += pcgSlice<T, T2, SPMV_MODE>(s, lane, A.M, pc, A.t, dynSmem);
            } else {
                for (long long s = gwarp; s < A.M.nslices; s += warpsTotal)
                    if (!(A.M.meta[s * 32] & SJDS_HALO_BIT)) dot += pcgSlice<T, T2, SPMV_MODE>(s, lane, A.M, pc, A.t, dynSmem);
                if (!waitHaloReady(A.D, A.mine, seqPush0 + k, st)) break;
                for (long long i = gwarp; i < A.D.nBoundary; i += warpsTotal)
                    dot += pcgSlice<T, T2, SPMV_MODE>((long long)A.D.boundarySlices[i], lane, A.M, pc, A.t, dynSmem);
            }
            pt[0] = blockSum(dot, sh);
        }
        if (!syncSum<1>(pt, A.parts, &st->arrive[0], tgtA, A.D, A.mine, A.D.seqReduce + 2ull * k + 1, st, sh, shw, &shLast)) break;
        if (timer) { const unsigned long long now = globalTimerNs(); tm[1] += now - tm[0]; tm[0] = now; }
        const double alphaD = rho / pt[0];
        breakdown = !isfinite(alphaD);
        const T alpha = breakdown ? (T)0 : (T)alphaD;
        // ---- x += alpha p, r -= alpha t, partial r.r and r.z (two elements per thread and trip: 128-bit accesses for fp64)
        double rs[2];
        {
            double rr = 0, rz = 0;
            for (long long j = gid; j < n2; j += gstride) {
                const T2 pv = ((const T2 *)pc)[j], tv = ((const T2 *)A.t)[j], dv = ((const T2 *)A.invDiag)[j];
                T2 rv = ((T2 *)A.r)[j], xv = ((T2 *)A.x)[j];
                rv.x -= alpha * tv.x; rv.y -= alpha * tv.y;
                xv.x += alpha * pv.x; xv.y += alpha * pv.y;
                ((T2 *)A.x)[j] = xv;
                ((T2 *)A.r)[j] = rv;
                rr += (double)rv.x * (double)rv.x + (double)rv.y * (double)rv.y;
                rz += (double)rv.x * (double)(T)(dv.x * rv.x) + (double)rv.y * (double)(T)(dv.y * rv.y);
            }
            if ((A.n & 1) && gid == 0) {
                const long long i = A.n - 1;
                T ri = A.r[i] - alpha * A.t[i];
                A.x[i] += alpha * pc[i];
                A.r[i] = ri;
                rr += (double)ri * (double)ri;
                rz += (double)ri * (double)(T)(A.invDiag[i] * ri);
            }
            rs[0] = blockSum(rr, sh);
            rs[1] = blockSum(rz, sh);
        }
        if (!syncSum<2>(rs, A.parts + G, &st->arrive[1], tgtB, A.D, A.mine, A.D.seqReduce + 2ull * k + 2, st, sh, shw, &shLast)) break;
        if (timer) { const unsigned long long now = globalTimerNs(); tm[2] += now - tm[0]; tm[0] = now; }
        ++k;
        if (timer) rrLast = rs[0];
        stop = (rs[0] < sc->threshold) || breakdown;
        if (stop) break;                       // Eigen: break before ++i
        // ---- p_{k+1} = z + beta p_k into the other buffer; the rows the peers need go out first
        const T beta = (T)(rs[1] / rho);
        rho = rs[1];
        ++iters;
        if (dist)
            for (long long i = gid; i < A.D.nSend; i += gstride) {
                const long long row = A.D.sendIdx[i] - A.D.rowBegin;
                const int2 dst = A.D.sendDst[i];
                *(volatile T *)(peerP<T>(A.D, dst.x, cur ^ 1) + dst.y) = pNext(A.invDiag[row], __ldcg(A.r + row), beta, pc[row]);
            }
        for (long long j = gid; j < n2; j += gstride) {
            const T2 dv = ((const T2 *)A.invDiag)[j], rv = ((const T2 *)A.r)[j], pv = ((const T2 *)pc)[j];
            T2 o;
            o.x = pNext(dv.x, rv.x, beta, pv.x);
            o.y = pNext(dv.y, rv.y, beta, pv.y);
            ((T2 *)pn)[j] = o;
        }
        if ((A.n & 1) && gid == 0) { const long long i = A.n - 1; pn[i] = pNext(A.invDiag[i], A.r[i], beta, pc[i]); }
        arrivePush(&st->arrive[2], tgtC, A.D, seqPush0 + k, pushes);
        if (timer) tm[3] += globalTimerNs() - tm[0];
    }
    if (timer) {
        sc->rho[0] = rho;
        sc->rho[1] = rho;
        sc->iters = iters;
        sc->residualNorm2 = rrLast;
        if (breakdown) sc->breakdown = 1;
        if (stop) sc->done = 1;
        // k counts completed (SpMV, x/r) phase pairs; a stopped iteration did both reduces but no p update / push
        st->seqReduce = A.D.seqReduce + 2ull * k;
        st->seqPush = seqPush0 + (stop ? k - 1 : k);
        st->phaseNs[0] += tm[1];
        st->phaseNs[1] += tm[2];
        st->phaseNs[2] += tm[3];
        st->spmvPhases += k;
    }
}

// flag[s] = 1 when slice s gathers from a halo slot (local column >= nLocal)
__global__ void k_slice_needs_halo(long long nslices, const long long *__restrict__ sliceOff, const int2 *__restrict__ col2, int nLocal,
                                   uint8_t *flag, int32_t *flag32, int32_t *meta) {
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    bool hit = false;
    for (long long i = sliceOff[s] + lane; i < sliceOff[s + 1]; i += 32) {
        const int2 c = col2[i];
        hit = hit || c.x >= nLocal || c.y >= nLocal;
    }
    const unsigned any = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) { flag[s] = any ? 1 : 0; flag32[s] = any ? 1 : 0; }
    if (any) meta[s * 32 + lane] |= SJDS_HALO_BIT;   // the SpMV's interior pass skips the slice without touching another array
}
__global__ void k_compact_slices(long long nslices, const uint8_t *flag, const long long *index, int32_t *list) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslices && flag[s]) list[index[s]] = (int32_t)s;
}

template <class T>
__global__ void k_convert_in(long long n, const double *in, T *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (T)in[i];
}
template <class T>
__global__ void k_convert_out(long long n, const T *in, double *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

static int spmvU() {
    static int u = -1;
    if (u < 0) {
        const char *e = getenv("AVS_SPMV_U");
        u = e ? atoi(e) : 4;
        if (u != 2 && u != 4 && u != 8) u = 4;
    }
    return u;
}
// AVS_CG_MODE=launch selects the three-launches-per-iteration loop (kept for A/B measurements and for the NCCL fallback);
// the default is the persistent cooperative kernel.
static bool cgUsePersistent() {
    static int t = -1;
    if (t < 0) {
        const char *e = getenv("AVS_CG_MODE");
        t = (e && strcmp(e, "launch") == 0) ? 0 : 1;
    }
    return t == 1;
}
// AVS_PCG_KERNEL=v1 selects round 1's persistent kernel (three grid barriers + separate all-reduce and halo push per iteration;
// kept for A/B measurements); the default is k_cg_persistent2.  AVS_PCG_MINB=5 (v1 only): register allocation for 5 CTAs per SM.
// Default: k_cg_persistent2 when the solve is row-partitioned over several ranks (its merged synchronisation points are what
// multi-GPU needs), round 1's k_cg_persistent on a single GPU (same speed at 10 M rows, 5-9 % faster on 0.2-1 M rows where the
// per-iteration synchronisation dominates: C2 18.8 vs 20.6 ms, C5 264 vs 278 ms).  AVS_PCG_KERNEL=v1 / v2 forces one.
static int pcgVersion(bool dist, bool fp32) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("AVS_PCG_KERNEL");
        forced = (e && strcmp(e, "v1") == 0) ? 1 : (e && strcmp(e, "v2") == 0) ? 2 : 0;
    }
    return forced ? forced : ((dist || fp32) ? 2 : 1);   // fp32: the prefetching slice loop (spmvMode) exists in k_cg_persistent2 only
}
template <class T, class T2>
static const void *pcgKernel() {
    static int minb = -1;
    if (minb < 0) {
        const char *e = getenv("AVS_PCG_MINB");
        minb = e ? atoi(e) : 4;
    }
    if (minb == 5) return (const void *)k_cg_persistent<T, T2, 4, 5>;
    return (const void *)k_cg_persistent<T, T2, 4, 4>;
}
// AVS_SPMV_MODE selects the slice loop of the stand-alone SpMV kernel and of the persistent CG kernel:
//   base (default) register-staged | pf = + L2 prefetch of the next trip | ring = matrix stream staged through a per-warp cp.async ring
// Default: base for fp64; pf for fp32 -- measured at C3 (N = 10.0 M): fp32 SpMV phase 0.342 ms with the prefetch against 0.366-0.371
// without (solve 127 vs 135 ms), fp64 0.477-0.482 with against 0.462-0.464 without.
static int spmvMode(bool fp32) {
    static int m = -2;
    if (m == -2) {
        const char *e = getenv("AVS_SPMV_MODE");
        m = -1;
        if (e && strcmp(e, "base") == 0) m = 0;
        if (e && strcmp(e, "pf") == 0) m = 1;
        if (e && strcmp(e, "ring") == 0) m = 2;
        if (e && strcmp(e, "ring4") == 0) m = 3;
        if (e && strcmp(e, "call") == 0) m = 4;     // persistent kernel calling the slice loop as a separate function (stand-alone SpMV: base)
    }
    return m >= 0 ? m : (fp32 ? 1 : 0);
}
#define SPMV_MODE_T spmvMode(sizeof(T) == 4)
template <class T, class T2>
static const void *pcg2Kernel(size_t *smemOut) {
    const void *k;
    size_t smem = 0;
    switch (SPMV_MODE_T) {
        case 1: k = (const void *)k_cg_persistent2<T, T2, 1, 4>; break;
        case 4: k = (const void *)k_cg_persistent2<T, T2, 3, 4>; break;
        case 2:
        case 3: k = (const void *)k_cg_persistent2<T, T2, 2, 4>; smem = (CG_THREADS / 32) * spmvRingBytesPerWarp<T2, 8>(); break;
        default: k = (const void *)k_cg_persistent2<T, T2, 0, 4>; break;
    }
    if (smem) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    *smemOut = smem;
    return k;
}
static bool spmvUseTma() {
    static int t = -1;
    if (t < 0) {
        const char *e = getenv("AVS_SPMV_TMA");
        t = (e && e[0] == '1') ? 1 : 0;
    }
    return t == 1;
}
template <class T, class T2>
static size_t tmaSmemBytes() { return 128 + 2 * (size_t)TMA_CAP_PAIRS * (sizeof(T2) + sizeof(int2)); }

template <class T, class T2, bool DOT>
static void launchSpmv(AvsContext *c, SellMatrix &A, const T *x, T *y, double *parts, const CgScalars *sc, int grid) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timeSpmv && c->spmvEventsUsed + 2 <= c->spmvEvents.size()) {
        e0 = c->spmvEvents[c->spmvEventsUsed++];
        e1 = c->spmvEvents[c->spmvEventsUsed++];
        cudaEventRecord(e0, c->stream);
    }
    if (spmvUseTma()) {
        static bool attrSet[2] = {false, false};
        const size_t smem = tmaSmemBytes<T, T2>();
        if (!attrSet[DOT]) {
            cudaFuncSetAttribute(k_spmv_tma<T, T2, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attrSet[DOT] = true;
        }
        k_spmv_tma<T, T2, DOT><<<grid, TMA_THREADS, smem, c->stream>>>(sjdsView<T, T2>(A), x, y, parts, sc);
    } else if (SPMV_MODE_T == 2 || SPMV_MODE_T == 3) {   // per-warp cp.async ring: 8 pair-steps in trips of 4 (mode 2) or 4 pair-steps in trips of 2 (mode 3)
#define RING_LAUNCH(U, R)                                                                                                        \
    {                                                                                                                            \
        static bool attrSet = false;                                                                                             \
        const size_t smem = (CG_THREADS / 32) * spmvRingBytesPerWarp<T2, R>();                                                   \
        if (!attrSet) {                                                                                                          \
            cudaFuncSetAttribute(k_spmv_sjds_ring<T, T2, DOT, U, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            attrSet = true;                                                                                                      \
        }                                                                                                                        \
        k_spmv_sjds_ring<T, T2, DOT, U, R><<<grid, CG_THREADS, smem, c->stream>>>(sjdsView<T, T2>(A), x, y, parts, sc);          \
    }
        if (SPMV_MODE_T == 2) RING_LAUNCH(4, 8) else RING_LAUNCH(2, 4)
#undef RING_LAUNCH
    } else if (SPMV_MODE_T == 1) {   // L2 prefetch of the next trip
        k_spmv_sjds_pf<T, T2, DOT, 4><<<grid, CG_THREADS, 0, c->stream>>>(sjdsView<T, T2>(A), x, y, parts, sc);
    } else {
#define SPMV_LAUNCH(U) k_spmv_sjds<T, T2, DOT, U><<<grid, CG_THREADS, 0, c->stream>>>(sjdsView<T, T2>(A), x, y, parts, sc)
        switch (spmvU()) {
            case 2: SPMV_LAUNCH(2); break;
            case 8: SPMV_LAUNCH(8); break;
            default: SPMV_LAUNCH(4); break;
        }
#undef SPMV_LAUNCH
    }
    if (e1) cudaEventRecord(e1, c->stream);
    ++c->launches;
    ++c->spmvLaunches;
}

// persistent SpMV grid: exactly the number of CTAs that are resident at once (one wave)
template <class T, class T2>
static int spmvGrid(AvsContext *c, long long nslices) {
    if (spmvUseTma()) {
        long long chunks = (nslices + TMA_WARPS - 1) / TMA_WARPS;
        return (int)std::max<long long>(1, std::min<long long>(chunks, c->numSMs));
    }
    int perSM = 0;
    cudaError_t e;
    if (SPMV_MODE_T == 2) {
        const size_t smem = (CG_THREADS / 32) * spmvRingBytesPerWarp<T2, 8>();
        cudaFuncSetAttribute(k_spmv_sjds_ring<T, T2, true, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds_ring<T, T2, true, 4, 8>, CG_THREADS, smem);
    } else if (SPMV_MODE_T == 3) {
        const size_t smem = (CG_THREADS / 32) * spmvRingBytesPerWarp<T2, 4>();
        cudaFuncSetAttribute(k_spmv_sjds_ring<T, T2, true, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds_ring<T, T2, true, 2, 4>, CG_THREADS, smem);
    } else if (SPMV_MODE_T == 1)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds_pf<T, T2, true, 4>, CG_THREADS, 0);
    else
    switch (spmvU()) {
        case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds<T, T2, true, 2>, CG_THREADS, 0); break;
        case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds<T, T2, true, 8>, CG_THREADS, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_sjds<T, T2, true, 4>, CG_THREADS, 0); break;
    }
    if (e != cudaSuccess || perSM < 1) {
        cudaGetLastError();
        perSM = 4;
    }
    const char *env = getenv("AVS_SPMV_WAVES");   // CTAs per SM as a multiple of the resident count (default 1 = one wave)
    int waves = env ? atoi(env) : 2;  // measured at C3: 2 waves 0.464 ms vs 1 wave 0.473 ms
    if (waves < 1) waves = 1;
    long long want = (nslices * 32 + CG_THREADS - 1) / CG_THREADS;
    long long cap = (long long)c->numSMs * perSM * waves;
    return (int)std::max<long long>(1, std::min(want, cap));
}

static int cgGrid(AvsContext *c, long long n) {
    long long want = (n + CG_THREADS - 1) / CG_THREADS;
    long long cap = (long long)c->numSMs * CG_CTAS_PER_SM;
    return (int)std::max<long long>(1, std::min(want, cap));
}

template <class T, class T2>
static int cgRunT(AvsContext *c, SellMatrix &A, const double *dRhs, const double *dX0, double *dXout, const AvsParams *p, AvsResult *res) {
    const long long n = A.n;
    CgWork &w = c->cg;
    const int grid = cgGrid(c, n);
    const int sgrid = spmvGrid<T, T2>(c, A.nslices);
    const bool dist = c->dist != nullptr;
    c->pcgUsed = false;
    c->pcgKernelMs = 0.f;
    c->pcgLaunches = 0;
    int pgrid = 0;   // persistent CG kernel: all CTAs co-resident (cooperative launch)
    {
        int perSM = 0;
        size_t smem2 = 0;
        const void *k2 = pcg2Kernel<T, T2>(&smem2);
        cudaError_t e = pcgVersion(dist, sizeof(T) == 4) == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k2, CG_THREADS, smem2)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, pcgKernel<T, T2>(), CG_THREADS, 0);
        if (e != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
        const char *env = getenv("AVS_PCG_CTAS_PER_SM");
        if (env && atoi(env) >= 1) perSM = std::min(perSM, atoi(env));
        long long want = std::max<long long>((A.nslices * 32 + CG_THREADS - 1) / CG_THREADS, 1);
        pgrid = (int)std::min<long long>(want, (long long)c->numSMs * perSM);
        // ranks that share one device (in-process group, tests) must be co-resident TOGETHER: they spin on each other's flags
        if (c->deviceShare > 1) perSM = std::max(1, perSM / c->deviceShare);
        if (dist) pgrid = (int)((long long)c->numSMs * perSM);   // every rank runs the flag logic even with few rows
        else pgrid = (int)std::min<long long>(pgrid, (long long)c->numSMs * perSM);
    }
    const long long nHalo = dist ? c->nHalo : 0;
    const size_t vb = (size_t)std::max<long long>(n, 1) * sizeof(T);
    if (w.x.reserve(vb) || w.r.reserve(vb) || w.t.reserve(vb)) return AVS_ERR_ALLOC;
    const long long pPad = (n + nHalo + 63) & ~63ll;   // single GPU: the persistent kernel's second p buffer starts here
    if (w.p.reserve((size_t)std::max<long long>(2 * pPad, 1) * sizeof(T))) return AVS_ERR_ALLOC;  // owned rows + halo slots, double buffered
    DevBuf &bbuf = c->cgRhs;
    if (bbuf.reserve(vb)) return AVS_ERR_ALLOC;
    if (w.partials.reserve(((size_t)std::max(grid, pgrid) * 3 + (size_t)sgrid + 16) * sizeof(double))) return AVS_ERR_ALLOC;
    if (w.scalars.reserve(sizeof(CgScalars))) return AVS_ERR_ALLOC;
    T *x = w.x.as<T>(), *r = w.r.as<T>(), *pp = w.p.as<T>(), *t = w.t.as<T>(), *b = bbuf.as<T>();
    if (dist) {  // peer-memory mode: p lives in the region the peers have mapped
        int rcp = AVS_OK;
        void *shared = avs_dist_prepare_p(c, &rcp);
        if (rcp) return rcp;
        if (shared) pp = (T *)shared;
    }
    double *parts = w.partials.as<double>();
    double *ptParts = parts + 3 * (size_t)grid;
    double *red = ptParts + sgrid;  // 16 doubles: globally reduced scalars (multi-GPU)
    CgScalars *sc = w.scalars.as<CgScalars>();
    res->iterations = 0;
    res->error = 0;
    if (n == 0 && !dist) return AVS_OK;
    const unsigned eb = (unsigned)((std::max<long long>(n, 1) + 255) / 256);
    k_convert_in<T><<<eb, 256, 0, c->stream>>>(n, dRhs, b);
    k_convert_in<T><<<eb, 256, 0, c->stream>>>(n, dX0, x);
    c->launches += 2;
    // residual = rhs - A x0 (Eigen: VectorType residual = rhs - mat * x)
    AVS_CUDA_CHECK(cudaMemsetAsync(sc, 0, sizeof(CgScalars), c->stream));
    int rcd;
    if (dist) {  // the SpMV reads its input with halo slots: stage x0 in p's buffer and exchange
        AVS_CUDA_CHECK(cudaMemcpyAsync(pp, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
        if ((rcd = avs_dist_halo_exchange(c, pp, A.precision, nullptr))) return rcd;
        launchSpmv<T, T2, false>(c, A, pp, t, nullptr, nullptr, sgrid);
        // barrier: peers may still be pulling the staged x0 out of this buffer; k_cg_init overwrites it
        if ((rcd = avs_dist_allreduce_parts(c, parts, 0, 1, red + 8, nullptr))) return rcd;
    } else
        launchSpmv<T, T2, false>(c, A, x, t, nullptr, nullptr, sgrid);
    k_cg_init<T><<<grid, CG_THREADS, 0, c->stream>>>(n, b, t, A.invDiag.as<T>(), r, pp, parts, grid);
    const double tiny = (sizeof(T) == 4) ? (double)FLT_MIN : DBL_MIN;
    if (dist) {
        if ((rcd = avs_dist_allreduce_parts(c, parts, grid, 3, red + 4, nullptr))) return rcd;
        k_cg_init_scalars<<<1, CG_THREADS, 0, c->stream>>>(red + 4, 1, p->tolerance, tiny, sc);
    } else
        k_cg_init_scalars<<<1, CG_THREADS, 0, c->stream>>>(parts, grid, p->tolerance, tiny, sc);
    c->launches += 2;

    AVS_TRACE("cg: init done, entering the loop (pgrid %d, grid %d, sgrid %d)", pgrid, grid, sgrid);
    const int maxIters = p->max_iterations;
    CgScalars *hs = (CgScalars *)c->hostScalars;  // pinned
    PcgDist pd;
    const bool persistent = cgUsePersistent() && (!dist || avs_dist_pcg_args(c, &pd));
    if (persistent && pcgVersion(dist, sizeof(T) == 4) == 2 && c->nranks <= 8) {   // the mailbox exchange of syncSum fits the peers' words into one warp
        // ---- k_cg_persistent2: one cooperative launch per `check_every` iterations (default: all of them) ----------------
        if (w.pcgState.reserve(sizeof(Pcg2State) + 64)) return AVS_ERR_ALLOC;
        Pcg2State *st = w.pcgState.as<Pcg2State>();
        Pcg2State *hst = (Pcg2State *)((char *)c->hostScalars + 256);
        AVS_CUDA_CHECK(cudaMemsetAsync(st, 0, sizeof(Pcg2State), c->stream));
        Pcg2Args<T, T2> ka;
        ka.n = n;
        ka.M = sjdsView<T, T2>(A);
        ka.invDiag = A.invDiag.as<T>();
        ka.x = x; ka.r = r; ka.t = t;
        ka.parts = parts; ka.sc = sc; ka.st = st;
        ka.pushFirst = 1;
        ka.parity0 = 0;
        if (dist) {
            ka.D = pd;
            ka.p0 = pp;
            ka.pDelta = (long long)pd.pStrideBytes;
            ka.mine = (P2PHeader *)((char *)pp - P2P_HEADER_BYTES);
            // slices that gather from a halo slot run after the peers' values have landed: flag them, list them in ascending order
            if (w.sliceHalo.reserve((size_t)std::max<long long>(A.nslices, 1)) || w.sliceFlag.reserve((size_t)std::max<long long>(A.nslices, 1) * 4) ||
                w.sliceIndex.reserve((size_t)(std::max<long long>(A.nslices, 1) + 1) * 8))
                return AVS_ERR_ALLOC;
            int64_t nBoundary = 0;
            if (A.nslices > 0) {
                k_slice_needs_halo<<<(unsigned)((A.nslices * 32 + 255) / 256), 256, 0, c->stream>>>(A.nslices, A.sliceOff.as<long long>(), A.col.as<int2>(), (int)n,
                                                                                          w.sliceHalo.as<uint8_t>(), w.sliceFlag.as<int32_t>(), A.meta.as<int32_t>());
                ++c->launches;
                int rcs = avs_exclusive_scan_i32_to_i64(c, w.sliceFlag.as<int32_t>(), w.sliceIndex.as<int64_t>(), A.nslices, &nBoundary);
                if (rcs) return rcs;
            }
            if (w.boundaryList.reserve((size_t)std::max<int64_t>(nBoundary, 1) * 4)) return AVS_ERR_ALLOC;
            if (nBoundary > 0) {
                k_compact_slices<<<(unsigned)((A.nslices + 255) / 256), 256, 0, c->stream>>>(A.nslices, w.sliceHalo.as<uint8_t>(), w.sliceIndex.as<long long>(),
                                                                                   w.boundaryList.as<int32_t>());
                ++c->launches;
            }
            ka.D.sliceHalo = w.sliceHalo.as<uint8_t>();
            ka.D.boundarySlices = w.boundaryList.as<int32_t>();
            ka.D.nBoundary = nBoundary;
            c->pcgBoundarySlices = nBoundary;
        } else {
            // single GPU: mailbox + flags in a local header, "peer table" with one entry
            if (w.pcgLocal.reserve(P2P_HEADER_BYTES + P2P_MAX_RANKS * sizeof(void *))) return AVS_ERR_ALLOC;
            AVS_CUDA_CHECK(cudaMemsetAsync(w.pcgLocal.p, 0, P2P_HEADER_BYTES, c->stream));
            void *self = w.pcgLocal.p;
            AVS_CUDA_CHECK(cudaMemcpyAsync((char *)w.pcgLocal.p + P2P_HEADER_BYTES, &self, sizeof(void *), cudaMemcpyHostToDevice, c->stream));
            ka.D = PcgDist();
            ka.D.peerRegion = (void *const *)((char *)w.pcgLocal.p + P2P_HEADER_BYTES);
            ka.p0 = pp;
            ka.pDelta = (long long)pPad * (long long)sizeof(T);
            ka.mine = (P2PHeader *)w.pcgLocal.p;
        }
        size_t smem2 = 0;
        const void *kern = pcg2Kernel<T, T2>(&smem2);
        const int chunk = p->check_every > 0 ? p->check_every : (p->cancel ? 256 : maxIters);
        int itersKnown = 0;
        while (true) {
            ka.itLimit = std::min(maxIters, itersKnown + std::max(chunk, 1));
            void *kargs[] = {(void *)&ka};
            AVS_CUDA_CHECK(cudaEventRecord(c->evPcg[0], c->stream));
            AVS_TRACE("rank %d cg: cooperative launch, grid %d itLimit %d smem %zu seqPush %llu seqReduce %llu", c->rank, pgrid, ka.itLimit, smem2, ka.D.seqPush, ka.D.seqReduce);
            AVS_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(pgrid), dim3(CG_THREADS), kargs, smem2, c->stream));
            AVS_CUDA_CHECK(cudaEventRecord(c->evPcg[1], c->stream));
            ++c->launches;
            ++c->pcgLaunches;
            AVS_CUDA_CHECK(cudaMemcpyAsync(&hs[0], sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
            AVS_CUDA_CHECK(cudaMemcpyAsync(hst, st, sizeof(Pcg2State), cudaMemcpyDeviceToHost, c->stream));
            AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, c->evPcg[0], c->evPcg[1]) == cudaSuccess) c->pcgKernelMs += ms;
            }
            AVS_TRACE("rank %d cg: kernel returned, iters %d done %d abort %d", c->rank, hs[0].iters, hs[0].done, hst->abort);
            if (hst->abort) {
                c->lastError = "persistent CG kernel timed out waiting for its own CTAs or a peer";
                avs_set_last_error("k_cg_persistent2 (spin-wait timeout)", cudaErrorLaunchTimeout, __FILE__, __LINE__);
                return AVS_ERR_CUDA;
            }
            ka.parity0 = (ka.parity0 + (hs[0].iters - itersKnown)) & 1;   // one buffer flip per completed p update
            itersKnown = hs[0].iters;
            ka.pushFirst = 0;
            ka.D.seqPush = hst->seqPush;
            ka.D.seqReduce = hst->seqReduce;
            if (hs[0].done || itersKnown >= maxIters) break;
            if (p->cancel && *p->cancel) {
                if (dist) avs_dist_pcg_commit(c, hst->seqPush, hst->seqReduce);
                return AVS_ERR_CANCELLED;
            }
            // arrival counters restart with every launch
            AVS_CUDA_CHECK(cudaMemsetAsync((char *)st + offsetof(Pcg2State, arrive), 0, 4 * sizeof(unsigned int), c->stream));
        }
        if (dist) avs_dist_pcg_commit(c, hst->seqPush, hst->seqReduce);
        c->pcgSpmvMs = (float)(hst->phaseNs[0] * 1e-6);
        c->pcgXrMs = (float)(hst->phaseNs[1] * 1e-6);
        c->pcgPMs = (float)(hst->phaseNs[2] * 1e-6);
        c->pcgPhases = (int64_t)hst->spmvPhases;
        c->pcgUsed = true;
    } else if (persistent) {
        // ---- round 1's kernel (AVS_PCG_KERNEL=v1) -----------------------------------------------------------------------
        if (w.pcgState.reserve(sizeof(PcgState))) return AVS_ERR_ALLOC;
        PcgState *st = w.pcgState.as<PcgState>();
        PcgState *hst = (PcgState *)((char *)c->hostScalars + 256);
        AVS_CUDA_CHECK(cudaMemsetAsync(st, 0, sizeof(PcgState), c->stream));
        PcgArgs<T, T2> ka;
        ka.n = n;
        ka.M = sjdsView<T, T2>(A);
        ka.invDiag = A.invDiag.as<T>();
        ka.x = x; ka.r = r; ka.p = pp; ka.t = t;
        ka.parts = parts; ka.sc = sc; ka.st = st;
        ka.D = dist ? pd : PcgDist();
        ka.pushFirst = 1;
        { const char *env = getenv("AVS_PCG_DYN"); ka.dynamic = (env && env[0] == '1') ? 1 : 0; }
        const int chunk = p->check_every > 0 ? p->check_every : (p->cancel ? 256 : maxIters);
        int itersKnown = 0;
        while (true) {
            ka.itLimit = std::min(maxIters, itersKnown + std::max(chunk, 1));
            void *kargs[] = {(void *)&ka};
            AVS_CUDA_CHECK(cudaEventRecord(c->evPcg[0], c->stream));
            AVS_CUDA_CHECK(cudaLaunchCooperativeKernel(pcgKernel<T, T2>(), dim3(pgrid), dim3(CG_THREADS), kargs, 0, c->stream));
            AVS_CUDA_CHECK(cudaEventRecord(c->evPcg[1], c->stream));
            ++c->launches;
            ++c->pcgLaunches;
            AVS_CUDA_CHECK(cudaMemcpyAsync(&hs[0], sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
            AVS_CUDA_CHECK(cudaMemcpyAsync(hst, st, sizeof(PcgState), cudaMemcpyDeviceToHost, c->stream));
            AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, c->evPcg[0], c->evPcg[1]) == cudaSuccess) c->pcgKernelMs += ms;
            }
            if (hst->abort) {
                c->lastError = "persistent CG kernel timed out waiting for a grid barrier or a peer";
                avs_set_last_error("k_cg_persistent (spin-wait timeout)", cudaErrorLaunchTimeout, __FILE__, __LINE__);
                return AVS_ERR_CUDA;
            }
            itersKnown = hs[0].iters;
            ka.pushFirst = 0;
            ka.D.seqPush = hst->seqPush;
            ka.D.seqReduce = hst->seqReduce;
            if (hs[0].done || itersKnown >= maxIters) break;
            if (p->cancel && *p->cancel) {
                if (dist) avs_dist_pcg_commit(c, hst->seqPush, hst->seqReduce);
                return AVS_ERR_CANCELLED;
            }
            // counters restart with every launch
            AVS_CUDA_CHECK(cudaMemsetAsync((char *)st + offsetof(PcgState, barrier), 0, 2 * sizeof(unsigned int) + 2 * sizeof(unsigned long long), c->stream));
        }
        if (dist) avs_dist_pcg_commit(c, hst->seqPush, hst->seqReduce);
        c->pcgSpmvMs = (float)(hst->phaseNs[0] * 1e-6);
        c->pcgXrMs = (float)(hst->phaseNs[1] * 1e-6);
        c->pcgPMs = (float)(hst->phaseNs[2] * 1e-6);
        c->pcgPhases = (int64_t)hst->spmvPhases;
        c->pcgUsed = true;
    } else {
        const int *doneFlag = (const int *)((const char *)sc + offsetof(CgScalars, done));
        // (opt-in, AVS_L2_PERSIST=1: measured -1% on the SpMV but +19% on the x,r update at C3, net loss)
        // Keep the SpMV's gathered vector resident in L2 while the matrix streams through: p is re-read ~17x per launch
        // (once per non-zero), the matrix exactly once.  Persisting window on p, streaming everything else.
        bool l2window = false;
        {
            const char *env = getenv("AVS_L2_PERSIST");
            int maxPersist = 0, maxWindow = 0;
            cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
            cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
            size_t bytes = (size_t)(n + nHalo) * sizeof(T);
            if ((env && env[0] == '1') && maxPersist > 0 && maxWindow > 0 && bytes > 0) {
                size_t setAside = std::min<size_t>((size_t)maxPersist, bytes);
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside);
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = (void *)pp;
                attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)maxWindow);
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)attr.accessPolicyWindow.num_bytes);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                l2window = cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
                if (!l2window) cudaGetLastError();
            }
        }
        int checkEvery = p->check_every > 0 ? p->check_every : 32;
        int launched = 0;
        bool done = false;
        // the flag copy of batch k is awaited only after batch k+1 has been enqueued, so the GPU never idles
        cudaEvent_t evPrev = nullptr;
        int slot = 0;
        while (!done) {
            int batch = std::min(checkEvery, maxIters - launched);
            for (int it = 0; it < batch; ++it) {
                const int parity = (launched + it) & 1;
                if (dist && (rcd = avs_dist_halo_exchange(c, pp, A.precision, doneFlag))) return rcd;
                launchSpmv<T, T2, true>(c, A, pp, t, ptParts, sc, sgrid);
                const double *ptSrc = ptParts, *rrSrc = parts;
                int ptN = sgrid, rrN = grid;
                if (dist) {
                    if ((rcd = avs_dist_allreduce_parts(c, ptParts, sgrid, 1, red, doneFlag))) return rcd;
                    ptSrc = red;
                    ptN = 1;
                }
                cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
                if (c->timeSpmv && c->auxEventsUsed + 3 <= c->auxEvents.size()) {
                    e0 = c->auxEvents[c->auxEventsUsed++]; e1 = c->auxEvents[c->auxEventsUsed++]; e2 = c->auxEvents[c->auxEventsUsed++];
                    cudaEventRecord(e0, c->stream);
                }
                k_cg_update_xr<T><<<grid, CG_THREADS, 0, c->stream>>>(n, pp, t, A.invDiag.as<T>(), x, r, ptSrc, ptN, parts, grid, sc, parity);
                if (e1) cudaEventRecord(e1, c->stream);
                if (dist) {
                    if ((rcd = avs_dist_allreduce_parts(c, parts, grid, 2, red + 1, doneFlag))) return rcd;
                    rrSrc = red + 1;
                    rrN = 1;
                }
                k_cg_update_p<T><<<grid, CG_THREADS, 0, c->stream>>>(n, r, A.invDiag.as<T>(), pp, rrSrc, rrN, sc, parity);
                if (e2) cudaEventRecord(e2, c->stream);
                c->launches += 3;
            }
            launched += batch;
            AVS_CUDA_CHECK(cudaMemcpyAsync(&hs[slot], sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
            AVS_CUDA_CHECK(cudaEventRecord(c->evPoll[slot], c->stream));
            if (evPrev) {
                AVS_CUDA_CHECK(cudaEventSynchronize(evPrev));
                if (hs[slot ^ 1].done) done = true;
            }
            evPrev = c->evPoll[slot];
            slot ^= 1;
            if (launched >= maxIters || batch == 0) {
                AVS_CUDA_CHECK(cudaEventSynchronize(evPrev));
                break;
            }
            if (p->cancel && *p->cancel) {
                AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
                return AVS_ERR_CANCELLED;
            }
        }
        AVS_CUDA_CHECK(cudaMemcpyAsync(&hs[0], sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (l2window) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaCtxResetPersistingL2Cache();
        }
    }
    if (dist && avs_dist_timed_out(c)) {
        c->lastError = "a multi-GPU exchange kernel timed out waiting for a peer";
        avs_set_last_error("halo exchange / all-reduce (spin-wait timeout)", cudaErrorLaunchTimeout, __FILE__, __LINE__);
        return AVS_ERR_CUDA;
    }
    CgScalars fin = hs[0];
    if (fin.zeroRhs) {
        AVS_CUDA_CHECK(cudaMemsetAsync(dXout, 0, (size_t)n * sizeof(double), c->stream));
        res->iterations = 0;
        res->error = 0;
        return AVS_OK;
    }
    k_convert_out<T><<<eb, 256, 0, c->stream>>>(n, x, dXout);
    ++c->launches;
    res->iterations = fin.iters;
    res->error = std::sqrt(fin.residualNorm2 / fin.rhsNorm2);
    AVS_CUDA_CHECK(cudaGetLastError());
    if (fin.breakdown) return AVS_ERR_BREAKDOWN;
    return AVS_OK;
}

int avs_cg_run(AvsContext *c, SellMatrix &A, const double *dRhs, const double *dX0, double *dXout, const AvsParams *p, AvsResult *res) {
    if (A.precision == AVS_PRECISION_F32) return cgRunT<float, float2>(c, A, dRhs, dX0, dXout, p, res);
    return cgRunT<double, double2>(c, A, dRhs, dX0, dXout, p, res);
}

// y = A x once (parity tests)
int avs_spmv_once(AvsContext *c, SellMatrix &A, const double *dX, double *dY) {
    const long long n = A.n;
    if (n == 0) return AVS_OK;
    const int grid = A.precision == AVS_PRECISION_F32 ? spmvGrid<float, float2>(c, A.nslices) : spmvGrid<double, double2>(c, A.nslices);
    const unsigned eb = (unsigned)((n + 255) / 256);
    if (A.precision == AVS_PRECISION_F32) {
        DevBuf xb, yb;
        if (xb.reserve((size_t)n * 4) || yb.reserve((size_t)n * 4)) return AVS_ERR_ALLOC;
        k_convert_in<float><<<eb, 256, 0, c->stream>>>(n, dX, xb.as<float>());
        launchSpmv<float, float2, false>(c, A, xb.as<float>(), yb.as<float>(), nullptr, nullptr, grid);
        k_convert_out<float><<<eb, 256, 0, c->stream>>>(n, yb.as<float>(), dY);
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        xb.release();
        yb.release();
    } else {
        launchSpmv<double, double2, false>(c, A, dX, dY, nullptr, nullptr, grid);
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// Times `repeats` launches of the CG's SpMV (with the fused dot, as in the solve) with CUDA events
// on the library stream.  x is filled with ones; the matrix is far larger than L2 for the bench configs.
template <class T>
__global__ void k_fill(long long n, T v, T *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
int avs_spmv_time(AvsContext *c, SellMatrix &A, int repeats, float *msPerLaunch) {
    const long long n = A.n;
    *msPerLaunch = 0;
    if (n == 0 || repeats <= 0) return AVS_OK;
    const int grid = A.precision == AVS_PRECISION_F32 ? spmvGrid<float, float2>(c, A.nslices) : spmvGrid<double, double2>(c, A.nslices);
    const size_t vs = A.precision == AVS_PRECISION_F32 ? 4 : 8;
    CgWork &w = c->cg;
    if (w.p.reserve((size_t)n * vs) || w.t.reserve((size_t)n * vs) || w.partials.reserve((size_t)grid * 4 * sizeof(double))) return AVS_ERR_ALLOC;
    const unsigned eb = (unsigned)((n + 255) / 256);
    bool saved = c->timeSpmv;
    c->timeSpmv = false;
    cudaEvent_t e0, e1;
    AVS_CUDA_CHECK(cudaEventCreate(&e0));
    AVS_CUDA_CHECK(cudaEventCreate(&e1));
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 = warm-up
        int reps = pass == 0 ? 3 : repeats;
        if (pass == 1) AVS_CUDA_CHECK(cudaEventRecord(e0, c->stream));
        for (int i = 0; i < reps; ++i) {
            if (A.precision == AVS_PRECISION_F32) {
                if (pass == 0 && i == 0) k_fill<float><<<eb, 256, 0, c->stream>>>(n, 1.f, w.p.as<float>());
                launchSpmv<float, float2, true>(c, A, w.p.as<float>(), w.t.as<float>(), w.partials.as<double>(), nullptr, grid);
            } else {
                if (pass == 0 && i == 0) k_fill<double><<<eb, 256, 0, c->stream>>>(n, 1.0, w.p.as<double>());
                launchSpmv<double, double2, true>(c, A, w.p.as<double>(), w.t.as<double>(), w.partials.as<double>(), nullptr, grid);
            }
        }
        if (pass == 1) AVS_CUDA_CHECK(cudaEventRecord(e1, c->stream));
    }
    AVS_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    AVS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *msPerLaunch = ms / (float)repeats;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->timeSpmv = saved;
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}
