// avs_dist.cu -- multi-GPU plumbing of the row-partitioned CG.
//
// One process per GPU.  Rank r owns the contiguous row block [N r/P, N (r+1)/P) of the Morton-brick
// row order; labelling/numbering are replicated, assembly and the CG are per rank.  The matrix couples a
// face only to faces <= 2 cells away (or the adjacent level's parent/children), so the off-rank columns
// of a block are a thin shell ("halo").  Per CG iteration the path has two real exchange steps:
//   (1) halo exchange of p before the SpMV  -- grouped ncclSend/ncclRecv with the few neighbouring ranks;
//   (2) sum-allreduce of p.Ap (1 double) and of {r.r, r.z} (2 doubles).
// NCCL is loaded at run time (dlopen "libnccl.so.2": inside a torch process this resolves to the NCCL torch
// already loaded), so the single-GPU library has no NCCL dependency.  The communicator is created by the
// library from a unique id the host broadcasts (avs_nccl_unique_id on rank 0 -> AvsDeviceConfig).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "avs_context.h"

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi g_nccl;

static bool loadNccl() {
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return false;
#define LOADSYM(field, sym)                                       \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, #sym);        \
    if (!g_nccl.field) return false;
    LOADSYM(GetUniqueId, ncclGetUniqueId)
    LOADSYM(CommInitRank, ncclCommInitRank)
    LOADSYM(CommDestroy, ncclCommDestroy)
    LOADSYM(AllReduce, ncclAllReduce)
    LOADSYM(AllGather, ncclAllGather)
    LOADSYM(Broadcast, ncclBroadcast)
    LOADSYM(Send, ncclSend)
    LOADSYM(Recv, ncclRecv)
    LOADSYM(GroupStart, ncclGroupStart)
    LOADSYM(GroupEnd, ncclGroupEnd)
    LOADSYM(GetErrorString, ncclGetErrorString)
#undef LOADSYM
    g_nccl.handle = h;
    return true;
}

#define NCCL_CHECK(c, expr)                                                                              \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            (c)->lastError = std::string(#expr) + ": " + g_nccl.GetErrorString(_r);                      \
            return AVS_ERR_NCCL;                                                                         \
        }                                                                                                \
    } while (0)

struct DistState {
    ncclComm_t comm = nullptr;
    DevBuf flag, index, haloCols, sendIdx, sendBuf, counts, scal;
    std::vector<int> recvCnt, recvOff, sendCnt, sendOff;
    long long nHalo = 0, nSend = 0;
};

extern "C" int avs_nccl_unique_id(void *out128) {
    if (!out128) return AVS_ERR_INVALID_ARGUMENT;
    if (!loadNccl()) return AVS_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return AVS_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "unique id is 128 bytes");
    memcpy(out128, &id, 128);
    return AVS_OK;
}

int avs_dist_init(AvsContext *c, const void *uniqueId128) {
    if (c->nranks <= 1) return AVS_OK;
    if (!uniqueId128) return AVS_ERR_INVALID_ARGUMENT;
    if (!loadNccl()) { c->lastError = "cannot load libnccl.so.2"; return AVS_ERR_NCCL; }
    DistState *d = new DistState();
    c->dist = d;
    ncclUniqueId id;
    memcpy(&id, uniqueId128, 128);
    NCCL_CHECK(c, g_nccl.CommInitRank(&d->comm, c->nranks, id, c->rank));
    d->recvCnt.assign(c->nranks, 0);
    d->recvOff.assign(c->nranks + 1, 0);
    d->sendCnt.assign(c->nranks, 0);
    d->sendOff.assign(c->nranks + 1, 0);
    if (d->scal.reserve(64 * sizeof(double))) return AVS_ERR_ALLOC;
    return AVS_OK;
}

void avs_dist_destroy(AvsContext *c) {
    DistState *d = (DistState *)c->dist;
    if (!d) return;
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    DevBuf *bufs[] = {&d->flag, &d->index, &d->haloCols, &d->sendIdx, &d->sendBuf, &d->counts, &d->scal};
    for (DevBuf *b : bufs) b->release();
    delete d;
    c->dist = nullptr;
}

// ---- halo discovery --------------------------------------------------------------------------------
__global__ void k_mark_halo(long long nLocal, long long stride, const int32_t *rowCount, const int32_t *stageCol,
                            long long rb, long long re, int32_t *flag) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nLocal) return;
    int n = rowCount[r];
    for (int i = 0; i < n; ++i) {
        int32_t col = stageCol[(long long)i * stride + r];
        if (col < rb || col >= re) flag[col] = 1;
    }
}
__global__ void k_compact_halo(long long n, const int32_t *flag, const long long *index, int32_t *haloCols) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) haloCols[index[i]] = (int32_t)i;
}
template <class T>
__global__ void k_pack_halo(long long n, const int32_t *sendIdx, long long rb, const T *p, T *sendBuf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sendBuf[i] = p[sendIdx[i] - rb];
}
__global__ void k_reduce_parts(const double *parts, int nparts, int count, double *out) {
    // one CTA of 256 threads; fixed order
    __shared__ double sh[256];
    for (int q = 0; q < count; ++q) {
        double v = 0;
        for (int i = threadIdx.x; i < nparts; i += 256) v += parts[(size_t)q * nparts + i];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[q] = sh[0];
        __syncthreads();
    }
}

// Finds the off-rank columns of the local rows, builds the remap table (global column -> halo slot) and the
// send lists.  After this call:  column c of a local row maps to  c - rowBegin            if rowBegin <= c < rowEnd
//                                                              nLocal + index[c]         otherwise.
int avs_dist_build_halo(AvsContext *c) {
    DistState *d = (DistState *)c->dist;
    if (!d) return AVS_OK;
    const int P = c->nranks;
    const long long N = c->nRows, nLocal = c->rowEnd - c->rowBegin;
    if (d->flag.reserve((size_t)std::max<long long>(N, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (d->index.reserve((size_t)(std::max<long long>(N, 1) + 1) * sizeof(long long))) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemsetAsync(d->flag.p, 0, (size_t)N * sizeof(int32_t), c->stream));
    if (nLocal > 0) {
        k_mark_halo<<<(unsigned)((nLocal + 255) / 256), 256, 0, c->stream>>>(nLocal, c->stageStride, c->rowCount.as<int32_t>(),
                                                                           c->stageCol.as<int32_t>(), c->rowBegin, c->rowEnd,
                                                                           d->flag.as<int32_t>());
        ++c->launches;
    }
    int64_t nHalo = 0;
    int rc = avs_exclusive_scan_i32_to_i64(c, d->flag.as<int32_t>(), d->index.as<int64_t>(), N, &nHalo);
    if (rc) return rc;
    d->nHalo = nHalo;
    AVS_CUDA_CHECK(cudaMemcpyAsync(d->index.as<long long>() + N, &d->nHalo, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    if (d->haloCols.reserve((size_t)std::max<long long>(nHalo, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (N > 0) {
        k_compact_halo<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(N, d->flag.as<int32_t>(), d->index.as<long long>(),
                                                                          d->haloCols.as<int32_t>());
        ++c->launches;
    }
    // halo slots are sorted by global column, hence grouped by owner: counts per owner from the scan at the block boundaries
    std::vector<long long> bnd(P + 1);
    for (int q = 0; q <= P; ++q) {
        long long row = N * q / P;
        AVS_CUDA_CHECK(cudaMemcpyAsync(&bnd[q], d->index.as<long long>() + row, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < P; ++q) {
        d->recvOff[q] = (int)bnd[q];
        d->recvCnt[q] = (int)(bnd[q + 1] - bnd[q]);
    }
    d->recvOff[P] = (int)bnd[P];
    // all-gather the P x P count matrix, then exchange the column lists
    if (d->counts.reserve((size_t)P * P * sizeof(int32_t) + (size_t)P * sizeof(int32_t))) return AVS_ERR_ALLOC;
    int32_t *dMine = d->counts.as<int32_t>() + (size_t)P * P;
    AVS_CUDA_CHECK(cudaMemcpyAsync(dMine, d->recvCnt.data(), (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(c, g_nccl.AllGather(dMine, d->counts.p, (size_t)P, ncclInt32, d->comm, c->stream));
    std::vector<int32_t> all((size_t)P * P);
    AVS_CUDA_CHECK(cudaMemcpyAsync(all.data(), d->counts.p, (size_t)P * P * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    long long nSend = 0;
    for (int q = 0; q < P; ++q) {
        d->sendCnt[q] = all[(size_t)q * P + c->rank];  // what rank q needs from me
        d->sendOff[q] = (int)nSend;
        nSend += d->sendCnt[q];
    }
    d->sendOff[P] = (int)nSend;
    d->nSend = nSend;
    if (d->sendIdx.reserve((size_t)std::max<long long>(nSend, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (d->sendBuf.reserve((size_t)std::max<long long>(nSend, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    NCCL_CHECK(c, g_nccl.GroupStart());
    for (int q = 0; q < P; ++q) {
        if (q == c->rank) continue;
        if (d->recvCnt[q] > 0) NCCL_CHECK(c, g_nccl.Send(d->haloCols.as<int32_t>() + d->recvOff[q], (size_t)d->recvCnt[q], ncclInt32, q, d->comm, c->stream));
        if (d->sendCnt[q] > 0) NCCL_CHECK(c, g_nccl.Recv(d->sendIdx.as<int32_t>() + d->sendOff[q], (size_t)d->sendCnt[q], ncclInt32, q, d->comm, c->stream));
    }
    NCCL_CHECK(c, g_nccl.GroupEnd());
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->nHalo = d->nHalo;
    c->haloIndex = d->index.as<long long>();
    return AVS_OK;
}

// p[nLocal + slot] <- owner's p for every halo slot
template <class T>
int avs_dist_halo_exchange_t(AvsContext *c, T *p) {
    DistState *d = (DistState *)c->dist;
    const int P = c->nranks;
    const long long nLocal = c->rowEnd - c->rowBegin;
    if (d->nSend > 0) {
        k_pack_halo<T><<<(unsigned)((d->nSend + 255) / 256), 256, 0, c->stream>>>(d->nSend, d->sendIdx.as<int32_t>(), c->rowBegin, p, d->sendBuf.as<T>());
        ++c->launches;
    }
    const ncclDataType_t dt = sizeof(T) == 4 ? ncclFloat32 : ncclFloat64;
    NCCL_CHECK(c, g_nccl.GroupStart());
    for (int q = 0; q < P; ++q) {
        if (q == c->rank) continue;
        if (d->sendCnt[q] > 0) NCCL_CHECK(c, g_nccl.Send(d->sendBuf.as<T>() + d->sendOff[q], (size_t)d->sendCnt[q], dt, q, d->comm, c->stream));
        if (d->recvCnt[q] > 0) NCCL_CHECK(c, g_nccl.Recv(p + nLocal + d->recvOff[q], (size_t)d->recvCnt[q], dt, q, d->comm, c->stream));
    }
    NCCL_CHECK(c, g_nccl.GroupEnd());
    return AVS_OK;
}
int avs_dist_halo_exchange(AvsContext *c, void *p, int precision) {
    if (!c->dist) return AVS_OK;
    return precision == AVS_PRECISION_F32 ? avs_dist_halo_exchange_t<float>(c, (float *)p) : avs_dist_halo_exchange_t<double>(c, (double *)p);
}

// out[0..count) = sum over ranks of sum over CTAs of parts[q*nparts + i]; identical bits on every rank
int avs_dist_allreduce_parts(AvsContext *c, const double *parts, int nparts, int count, double *out) {
    DistState *d = (DistState *)c->dist;
    k_reduce_parts<<<1, 256, 0, c->stream>>>(parts, nparts, count, out);
    ++c->launches;
    if (d) NCCL_CHECK(c, g_nccl.AllReduce(out, out, (size_t)count, ncclFloat64, ncclSum, d->comm, c->stream));
    return AVS_OK;
}

// every rank ends up with the full solution vector (rank q's block at offset N q/P)
int avs_dist_allgather_solution(AvsContext *c, const double *local, double *full) {
    DistState *d = (DistState *)c->dist;
    const int P = c->nranks;
    const long long N = c->nRows;
    if (!d) return AVS_OK;
    AVS_CUDA_CHECK(cudaMemcpyAsync(full + c->rowBegin, local, (size_t)(c->rowEnd - c->rowBegin) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    NCCL_CHECK(c, g_nccl.GroupStart());
    for (int q = 0; q < P; ++q) {
        long long b = N * q / P, e = N * (q + 1) / P;
        if (e > b) NCCL_CHECK(c, g_nccl.Broadcast(full + b, full + b, (size_t)(e - b), ncclFloat64, q, d->comm, c->stream));
    }
    NCCL_CHECK(c, g_nccl.GroupEnd());
    return AVS_OK;
}
