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
//
// Second way in: ONE process drives all GPUs (avs_create_multi, the shape of a Houdini DOP: solveGasSubclass is called on
// one cook thread, HDK_AdaptiveViscosity.cpp:126-128).  The rank contexts then share a LocalGroup: the set-up collectives
// (counts, halo lists, solution / slab gathers) are a host barrier plus plain device-to-device copies between the ranks'
// buffers -- same address space, so a published pointer is enough -- and the peer regions of the CG are mapped with
// cudaDeviceEnablePeerAccess instead of CUDA IPC.  No NCCL at all; several ranks may even share one device (tests).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "avs_context.h"
#include "avs_p2p.cuh"

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

// Peer-memory ("p2p") mode: every rank exports one cudaMalloc region through CUDA IPC
//   [ header: ready flag, all-reduce mailboxes + sequence flags | the CG's p vector (owned rows, then halo slots) ]
// and maps the regions of all peers.  The hot loop then needs no NCCL call: halo values are LOADED from the owners'
// p vectors over NVLink by k_p2p_halo_pull, and the scalar all-reduce is P remote STORES into the peers' mailboxes
// followed by a local spin (k_p2p_allreduce) -- both inside our own kernels.
// Shared by the rank contexts of one process (avs_create_multi).  Every collective below is called by all rank threads.
struct LocalGroup {
    int P = 0;
    int device[P2P_MAX_RANKS] = {};
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long generation = 0;
    std::atomic<int> failed{0};
    const void *ptr[P2P_MAX_RANKS][4] = {};
    unsigned long long u64[P2P_MAX_RANKS][8] = {};
    const int *hostInts[P2P_MAX_RANKS] = {};
    // returns false when a rank has failed (or nobody showed up for 120 s): callers give up instead of hanging
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long long gen = generation;
        if (++waiting == P) {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return failed.load() == 0;
        }
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
        while (generation == gen) {
            if (cv.wait_for(lk, std::chrono::milliseconds(50)) == std::cv_status::timeout) {
                if (failed.load() || std::chrono::steady_clock::now() > deadline) {
                    failed.store(1);
                    --waiting;
                    return false;
                }
            }
        }
        return failed.load() == 0;
    }
};
void *avs_local_group_create(int P, const int *devices) {
    LocalGroup *g = new LocalGroup();
    g->P = P;
    for (int q = 0; q < P; ++q) g->device[q] = devices[q];
    return g;
}
void avs_local_group_destroy(void *g) { delete (LocalGroup *)g; }
void avs_local_group_fail(void *g) { if (g) { ((LocalGroup *)g)->failed.store(1); ((LocalGroup *)g)->cv.notify_all(); } }
void avs_local_group_reset(void *g) { if (g) ((LocalGroup *)g)->failed.store(0); }

struct DistState {
    ncclComm_t comm = nullptr;
    LocalGroup *local = nullptr;           // single-process mode (avs_create_multi): no NCCL
    // peer-memory mode
    bool p2p = false;
    void *region = nullptr;                // my exported region
    size_t regionBytes = 0;
    void *peerRegion[P2P_MAX_RANKS] = {};  // mapped regions (peerRegion[rank] == region)
    DevBuf peerTable;                      // device copy of peerRegion[]
    DevBuf haloSrc;                        // int2 per halo slot: (owner rank, index in the owner's p)
    unsigned long long seqReady = 0, seqReduce = 0, seqPush = 0;
    DevBuf sendDst;                        // int2 per send entry: (peer, element index in the peer's p) -- push mode
    std::vector<int32_t> allCounts;        // P x P: allCounts[q*P + o] = halo values rank q needs from owner o
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

#define LOCAL_BARRIER(c, d)                                                                      \
    do {                                                                                         \
        if (!(d)->local->barrier()) {                                                            \
            (c)->lastError = "a rank of the in-process group failed or timed out";               \
            return AVS_ERR_NCCL;                                                                 \
        }                                                                                        \
    } while (0)

// ---- the set-up collectives, NCCL or in-process ---------------------------------------------------------------------------
// recv[q * bytes ..] = rank q's send buffer (device pointers)
static int commAllGather(AvsContext *c, DistState *d, const void *dSend, void *dRecv, size_t bytes) {
    if (!d->local) {
        NCCL_CHECK(c, g_nccl.AllGather(dSend, dRecv, bytes, ncclUint8, d->comm, c->stream));
        return AVS_OK;
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    d->local->ptr[c->rank][0] = dSend;
    LOCAL_BARRIER(c, d);
    for (int q = 0; q < c->nranks; ++q)
        AVS_CUDA_CHECK(cudaMemcpyPeerAsync((char *)dRecv + (size_t)q * bytes, c->device, d->local->ptr[q][0], d->local->device[q], bytes, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    LOCAL_BARRIER(c, d);
    return AVS_OK;
}
// in place on the device; op: 0 sum, 1 max, 2 min
static int commAllReduceU64(AvsContext *c, DistState *d, unsigned long long *dv, int count, int op) {
    if (!d->local) {
        const ncclRedOp_t o = op == 0 ? ncclSum : op == 1 ? ncclMax : ncclMin;
        NCCL_CHECK(c, g_nccl.AllReduce(dv, dv, (size_t)count, ncclUint64, o, d->comm, c->stream));
        return AVS_OK;
    }
    if (count > 8) return AVS_ERR_INVALID_ARGUMENT;
    AVS_CUDA_CHECK(cudaMemcpyAsync(d->local->u64[c->rank], dv, (size_t)count * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    LOCAL_BARRIER(c, d);
    unsigned long long r[8];
    for (int i = 0; i < count; ++i) {
        unsigned long long v = d->local->u64[0][i];
        for (int q = 1; q < c->nranks; ++q) {
            const unsigned long long w = d->local->u64[q][i];
            v = op == 0 ? v + w : op == 1 ? std::max(v, w) : std::min(v, w);
        }
        r[i] = v;
    }
    LOCAL_BARRIER(c, d);   // everybody has read before anybody publishes again
    AVS_CUDA_CHECK(cudaMemcpyAsync(dv, r, (size_t)count * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}
// every rank's buffer `base` (same layout on all ranks) receives, for each q, bytes [off[q], off[q] + len[q]) from rank q's buffer
static int commGatherSegments(AvsContext *c, DistState *d, void *base, const size_t *off, const size_t *len, ncclDataType_t dt, size_t elem) {
    const int P = c->nranks;
    if (!d->local) {
        NCCL_CHECK(c, g_nccl.GroupStart());
        for (int q = 0; q < P; ++q)
            if (len[q] > 0) NCCL_CHECK(c, g_nccl.Broadcast((char *)base + off[q], (char *)base + off[q], len[q] / elem, dt, q, d->comm, c->stream));
        NCCL_CHECK(c, g_nccl.GroupEnd());
        return AVS_OK;
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    d->local->ptr[c->rank][1] = base;
    LOCAL_BARRIER(c, d);
    for (int q = 0; q < P; ++q)
        if (q != c->rank && len[q] > 0)
            AVS_CUDA_CHECK(cudaMemcpyPeerAsync((char *)base + off[q], c->device, (const char *)d->local->ptr[q][1] + off[q], d->local->device[q], len[q], c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    LOCAL_BARRIER(c, d);
    return AVS_OK;
}

int avs_dist_init(AvsContext *c, const void *uniqueId128) {
    if (c->nranks <= 1) return AVS_OK;
    if (c->localGroup) {   // single-process mode: the group object replaces the communicator
        if (c->nranks > P2P_MAX_RANKS) return AVS_ERR_UNSUPPORTED;
        DistState *d = new DistState();
        c->dist = d;
        d->local = (LocalGroup *)c->localGroup;
        d->recvCnt.assign(c->nranks, 0);
        d->recvOff.assign(c->nranks + 1, 0);
        d->sendCnt.assign(c->nranks, 0);
        d->sendOff.assign(c->nranks + 1, 0);
        if (d->scal.reserve(64 * sizeof(double))) return AVS_ERR_ALLOC;
        d->p2p = true;
        for (int q = 0; q < c->nranks; ++q) {   // direct loads / stores into the peers' regions
            const int dev = d->local->device[q];
            if (dev == c->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->device, dev);
            if (!can) { c->lastError = "GPUs of the in-process group cannot access each other's memory"; return AVS_ERR_UNSUPPORTED; }
            cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->lastError = "cudaDeviceEnablePeerAccess failed"; cudaGetLastError(); return AVS_ERR_CUDA; }
            cudaGetLastError();
        }
        return AVS_OK;
    }
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
    const char *mode = getenv("AVS_DIST_MODE");   // "p2p" (default) or "nccl"
    d->p2p = !(mode && strcmp(mode, "nccl") == 0) && c->nranks <= P2P_MAX_RANKS;
    return AVS_OK;
}

void avs_dist_destroy(AvsContext *c) {
    DistState *d = (DistState *)c->dist;
    if (!d) return;
    if (!d->local)
        for (int q = 0; q < c->nranks && q < P2P_MAX_RANKS; ++q)
            if (q != c->rank && d->peerRegion[q]) cudaIpcCloseMemHandle(d->peerRegion[q]);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);   // also orders "close" before the peers' "free"
    if (d->region) cudaFree(d->region);
    d->peerTable.release();
    d->haloSrc.release();
    DevBuf *bufs[] = {&d->flag, &d->index, &d->haloCols, &d->sendIdx, &d->sendBuf, &d->counts, &d->scal, &d->sendDst};
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
        long long row = c->rowStarts[q];
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
    {
        int rcg = commAllGather(c, d, dMine, d->counts.p, (size_t)P * sizeof(int32_t));
        if (rcg) return rcg;
    }
    std::vector<int32_t> &all = d->allCounts;
    all.assign((size_t)P * P, 0);
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
    if (d->local) {
        // my send list for rank q = the block of q's (sorted, owner-grouped) halo column list that I own: pull it out of q's buffer
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        d->local->ptr[c->rank][2] = d->haloCols.p;
        LOCAL_BARRIER(c, d);
        for (int q = 0; q < P; ++q) {
            if (q == c->rank || d->sendCnt[q] == 0) continue;
            long long off = 0;
            for (int o = 0; o < c->rank; ++o) off += all[(size_t)q * P + o];
            AVS_CUDA_CHECK(cudaMemcpyPeerAsync(d->sendIdx.as<int32_t>() + d->sendOff[q], c->device, (const int32_t *)d->local->ptr[q][2] + off,
                                               d->local->device[q], (size_t)d->sendCnt[q] * sizeof(int32_t), c->stream));
        }
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        LOCAL_BARRIER(c, d);
    } else {
        NCCL_CHECK(c, g_nccl.GroupStart());
        for (int q = 0; q < P; ++q) {
            if (q == c->rank) continue;
            if (d->recvCnt[q] > 0) NCCL_CHECK(c, g_nccl.Send(d->haloCols.as<int32_t>() + d->recvOff[q], (size_t)d->recvCnt[q], ncclInt32, q, d->comm, c->stream));
            if (d->sendCnt[q] > 0) NCCL_CHECK(c, g_nccl.Recv(d->sendIdx.as<int32_t>() + d->sendOff[q], (size_t)d->sendCnt[q], ncclInt32, q, d->comm, c->stream));
        }
        NCCL_CHECK(c, g_nccl.GroupEnd());
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    c->nHalo = d->nHalo;
    c->haloIndex = d->index.as<long long>();
    return AVS_OK;
}

// ---- peer-memory mode ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ldVolatile(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
// 8 s on %globaltimer since the first call (t0 == 0)
__device__ __forceinline__ bool spinExpired(unsigned long long &t0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > 8000000000ull;
}

struct RowStarts { long long v[P2P_MAX_RANKS + 1]; };
__global__ void k_halo_sources(long long nHalo, const int32_t *haloCols, const __grid_constant__ RowStarts starts, int P, int2 *src) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nHalo) return;
    long long col = haloCols[i];
    int q = 0;
    while (q + 1 < P && starts.v[q + 1] <= col) ++q;   // owner of global row `col`
    src[i] = make_int2(q, (int)(col - starts.v[q]));
}

// halo slot i of my p  <-  p of its owner, read over NVLink.  `seq` identifies the p vector (CG iteration).
template <class T>
__global__ void k_p2p_halo_pull(long long nHalo, const int2 *src, void *const *peerRegion, int myRank, int P,
                                unsigned long long seq, unsigned neighbourMask, long long nLocal, const int *done) {
    if (done && *done) return;  // converged: every rank holds identical scalars, so every rank skips consistently
    P2PHeader *mine = (P2PHeader *)peerRegion[myRank];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // my p (written by the previous kernel on this stream) is complete: publish
        __threadfence_system();
        *(volatile unsigned long long *)&mine->ready = seq;
    }
    if (threadIdx.x < P && ((neighbourMask >> threadIdx.x) & 1u)) {
        const P2PHeader *peer = (const P2PHeader *)peerRegion[threadIdx.x];
        unsigned long long t0 = 0;
        while (ldVolatile(&peer->ready) < seq) {
            __nanosleep(200);
            if (spinExpired(t0)) { *(volatile unsigned long long *)&mine->timedOut = 1; break; }   // a lost peer must not hang the GPU
        }
    }
    __syncthreads();
    __threadfence_system();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nHalo) return;
    int2 s = src[i];
    const T *peerP = (const T *)((const char *)peerRegion[s.x] + P2P_HEADER_BYTES);
    T *myP = (T *)((char *)mine + P2P_HEADER_BYTES);
    myP[nLocal + i] = *(const volatile T *)(peerP + s.y);
}

// out[0..count) = sum over ranks (in rank order) of this rank's sum over CTAs of parts[q*nparts + i].
// One CTA.  Remote stores into every peer's mailbox, then a local spin until all P contributions arrived.
__global__ void k_p2p_allreduce(const double *parts, int nparts, int count, void *const *peerRegion, int myRank, int P,
                                unsigned long long seq, double *out, const int *done) {
    if (done && *done) return;
    __shared__ double sh[256];
    __shared__ double local[4];
    for (int q = 0; q < count; ++q) {
        double v = 0;
        for (int i = threadIdx.x; i < nparts; i += 256) v += parts[(size_t)q * nparts + i];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) local[q] = sh[0];
        __syncthreads();
    }
    const int par = (int)(seq & 1ull);
    if (threadIdx.x < P) {
        P2PHeader *peer = (P2PHeader *)peerRegion[threadIdx.x];
        for (int q = 0; q < count; ++q) *(volatile double *)&peer->mail[par][myRank][q] = local[q];
        __threadfence_system();
        *(volatile unsigned long long *)&peer->flag[par][myRank] = seq;
    }
    P2PHeader *mine = (P2PHeader *)peerRegion[myRank];
    if (threadIdx.x < P) {
        unsigned long long t0 = 0;
        while (ldVolatile(&mine->flag[par][threadIdx.x]) != seq) {
            __nanosleep(100);
            if (spinExpired(t0)) { *(volatile unsigned long long *)&mine->timedOut = 1; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x < count) {
        double t = 0;
        for (int r = 0; r < P; ++r) t += *(const volatile double *)&mine->mail[par][r][threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// (Re)creates the exported region so that it can hold `elems` vector elements of 8 bytes; collective.
static int p2pEnsureRegion(AvsContext *c, size_t elems) {
    DistState *d = (DistState *)c->dist;
    AVS_TRACE("rank %d: ensure region for %zu elements (have %zu bytes)", c->rank, elems, d->regionBytes);
    const int P = c->nranks;
    // agree on the capacity: max over ranks
    unsigned long long need = (unsigned long long)(P2P_HEADER_BYTES + elems * 8 + 256);
    unsigned long long *dNeed = (unsigned long long *)d->scal.as<double>() + 32;
    int rcc;
    AVS_CUDA_CHECK(cudaMemcpyAsync(dNeed, &need, sizeof(need), cudaMemcpyHostToDevice, c->stream));
    if ((rcc = commAllReduceU64(c, d, dNeed, 1, 1))) return rcc;
    AVS_CUDA_CHECK(cudaMemcpyAsync(&need, dNeed, sizeof(need), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (d->region && need <= d->regionBytes) return AVS_OK;
    // tear down the old mappings (everyone closes before anyone frees: the all-reduce above/below orders it)
    for (int q = 0; q < P; ++q)
        if (q != c->rank && d->peerRegion[q]) { if (!d->local) cudaIpcCloseMemHandle(d->peerRegion[q]); d->peerRegion[q] = nullptr; }
    if ((rcc = commAllReduceU64(c, d, dNeed, 1, 1))) return rcc;
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (d->region) cudaFree(d->region);
    d->region = nullptr;
    size_t bytes = (size_t)(need + need / 4);
    if (cudaMalloc(&d->region, bytes) != cudaSuccess) { cudaGetLastError(); return AVS_ERR_ALLOC; }
    d->regionBytes = bytes;
    AVS_CUDA_CHECK(cudaMemsetAsync(d->region, 0, P2P_HEADER_BYTES, c->stream));
    d->seqReady = 0;
    d->seqReduce = 0;
    d->seqPush = 0;
    if (d->local) {   // one address space: the pointer itself is the handle (peer access was enabled in avs_dist_init)
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        d->local->ptr[c->rank][3] = d->region;
        LOCAL_BARRIER(c, d);
        for (int q = 0; q < P; ++q) d->peerRegion[q] = const_cast<void *>(d->local->ptr[q][3]);
        LOCAL_BARRIER(c, d);
        if (d->peerTable.reserve(P2P_MAX_RANKS * sizeof(void *))) return AVS_ERR_ALLOC;
        AVS_CUDA_CHECK(cudaMemcpyAsync(d->peerTable.p, d->peerRegion, P2P_MAX_RANKS * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return AVS_OK;
    }
    // exchange IPC handles
    cudaIpcMemHandle_t mine;
    AVS_CUDA_CHECK(cudaIpcGetMemHandle(&mine, d->region));
    DevBuf hb;
    if (hb.reserve((size_t)(P + 1) * sizeof(cudaIpcMemHandle_t))) return AVS_ERR_ALLOC;
    char *dAll = hb.as<char>(), *dMine = dAll + (size_t)P * sizeof(cudaIpcMemHandle_t);
    AVS_CUDA_CHECK(cudaMemcpyAsync(dMine, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(c, g_nccl.AllGather(dMine, dAll, sizeof(mine), ncclUint8, d->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(P);
    AVS_CUDA_CHECK(cudaMemcpyAsync(all.data(), dAll, (size_t)P * sizeof(mine), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    hb.release();
    int okLocal = 1;
    for (int q = 0; q < P; ++q) {
        if (q == c->rank) { d->peerRegion[q] = d->region; continue; }
        if (cudaIpcOpenMemHandle(&d->peerRegion[q], all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            d->peerRegion[q] = nullptr;
            okLocal = 0;
        }
    }
    // every rank must agree on the mode
    unsigned long long ok = (unsigned long long)okLocal;
    AVS_CUDA_CHECK(cudaMemcpyAsync(dNeed, &ok, sizeof(ok), cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(c, g_nccl.AllReduce(dNeed, dNeed, 1, ncclUint64, ncclMin, d->comm, c->stream));
    AVS_CUDA_CHECK(cudaMemcpyAsync(&ok, dNeed, sizeof(ok), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (!ok) { d->p2p = false; return AVS_OK; }
    if (d->peerTable.reserve(P2P_MAX_RANKS * sizeof(void *))) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(d->peerTable.p, d->peerRegion, P2P_MAX_RANKS * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}

struct SendTable { int off[P2P_MAX_RANKS + 1]; long long base[P2P_MAX_RANKS]; };
__global__ void k_send_dst(long long nSend, const __grid_constant__ SendTable st, int P, int2 *dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSend) return;
    int q = 0;
    while (q + 1 < P && st.off[q + 1] <= i) ++q;
    dst[i] = make_int2(q, (int)(st.base[q] + (i - st.off[q])));
}

// Called once per solve after the halo is known.  Returns the p vector the CG must use (inside the exported region)
// or nullptr when the NCCL path is active.
void *avs_dist_prepare_p(AvsContext *c, int *rcOut) {
    DistState *d = (DistState *)c->dist;
    *rcOut = AVS_OK;
    if (!d || !d->p2p) return nullptr;
    const long long nLocal = c->rowEnd - c->rowBegin;
    int rc = p2pEnsureRegion(c, 2 * (size_t)(nLocal + d->nHalo + 64) + 64);   // two p buffers (the persistent kernel double-buffers p)
    if (rc) { *rcOut = rc; return nullptr; }
    if (!d->p2p) return nullptr;
    if (d->haloSrc.reserve((size_t)std::max<long long>(d->nHalo, 1) * sizeof(int2))) { *rcOut = AVS_ERR_ALLOC; return nullptr; }
    if (d->nHalo > 0) {
        RowStarts rs;
        for (int q = 0; q <= c->nranks; ++q) rs.v[q] = c->rowStarts[q];
        k_halo_sources<<<(unsigned)((d->nHalo + 255) / 256), 256, 0, c->stream>>>(d->nHalo, d->haloCols.as<int32_t>(), rs, c->nranks,
                                                                                 d->haloSrc.as<int2>());
        ++c->launches;
    }
    // push mode (persistent CG kernel): where every value of my send list lands in its consumer's p vector.
    // Rank q keeps its halo slots sorted by global column, i.e. grouped by owner; my block starts at slot
    // sum_{o < me} allCounts[q*P + o] and holds my send list for q in the same order (avs_dist_build_halo).
    if (d->sendDst.reserve((size_t)std::max<long long>(d->nSend, 1) * sizeof(int2))) { *rcOut = AVS_ERR_ALLOC; return nullptr; }
    if (d->nSend > 0) {
        SendTable st;
        for (int q = 0; q < c->nranks; ++q) {
            long long slot0 = 0;
            for (int o = 0; o < c->rank; ++o) slot0 += d->allCounts[(size_t)q * c->nranks + o];
            st.off[q] = d->sendOff[q];
            st.base[q] = (c->rowStarts[q + 1] - c->rowStarts[q]) + slot0;
        }
        st.off[c->nranks] = d->sendOff[c->nranks];
        k_send_dst<<<(unsigned)((d->nSend + 255) / 256), 256, 0, c->stream>>>(d->nSend, st, c->nranks, d->sendDst.as<int2>());
        ++c->launches;
    }
    return (char *)d->region + P2P_HEADER_BYTES;
}

// Everything the persistent CG kernel needs for its in-kernel exchanges; false when the peer-memory mode is off.
bool avs_dist_pcg_args(AvsContext *c, PcgDist *out) {
    DistState *d = (DistState *)c->dist;
    if (!d || !d->p2p || !d->region) return false;
    PcgDist a;
    a.P = c->nranks;
    a.myRank = c->rank;
    a.peerRegion = d->peerTable.as<void *>();
    a.sendIdx = d->sendIdx.as<int32_t>();
    a.sendDst = d->sendDst.as<int2>();
    a.nSend = d->nSend;
    a.rowBegin = c->rowBegin;
    for (int q = 0; q < c->nranks; ++q) {
        if (q == c->rank) continue;
        if (d->recvCnt[q] > 0) a.recvMask |= 1u << q;
        if (d->sendCnt[q] > 0) a.sendMask |= 1u << q;
    }
    a.seqPush = d->seqPush;
    a.seqReduce = d->seqReduce;
    // regionBytes is the same on every rank (p2pEnsureRegion agrees on the maximum), hence so is the buffer distance
    a.pStrideBytes = ((unsigned long long)(d->regionBytes - P2P_HEADER_BYTES) / 2) & ~255ull;
    *out = a;
    return true;
}
void avs_dist_pcg_commit(AvsContext *c, unsigned long long seqPush, unsigned long long seqReduce) {
    DistState *d = (DistState *)c->dist;
    if (!d) return;
    d->seqPush = seqPush;
    d->seqReduce = seqReduce;
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
int avs_dist_halo_exchange(AvsContext *c, void *p, int precision, const int *done) {
    if (!c->dist) return AVS_OK;
    DistState *d = (DistState *)c->dist;
    if (d->p2p) {
        const long long nLocal = c->rowEnd - c->rowBegin;
        unsigned mask = 0;
        for (int q = 0; q < c->nranks; ++q)
            if (q != c->rank && d->recvCnt[q] > 0) mask |= 1u << q;
        ++d->seqReady;
        AVS_TRACE("rank %d: halo pull seq %llu (nHalo %lld, mask %x)", c->rank, d->seqReady, d->nHalo, mask);
        unsigned blocks = (unsigned)std::max<long long>(1, (d->nHalo + 255) / 256);
        if (precision == AVS_PRECISION_F32)
            k_p2p_halo_pull<float><<<blocks, 256, 0, c->stream>>>(d->nHalo, d->haloSrc.as<int2>(), d->peerTable.as<void *>(), c->rank, c->nranks,
                                                                  d->seqReady, mask, nLocal, done);
        else
            k_p2p_halo_pull<double><<<blocks, 256, 0, c->stream>>>(d->nHalo, d->haloSrc.as<int2>(), d->peerTable.as<void *>(), c->rank, c->nranks,
                                                                   d->seqReady, mask, nLocal, done);
        ++c->launches;
        (void)p;
        return AVS_OK;
    }
    if (d->local) { c->lastError = "the in-process group has no NCCL hot loop"; return AVS_ERR_UNSUPPORTED; }
    return precision == AVS_PRECISION_F32 ? avs_dist_halo_exchange_t<float>(c, (float *)p) : avs_dist_halo_exchange_t<double>(c, (double *)p);
}

// out[0..count) = sum over ranks of sum over CTAs of parts[q*nparts + i]; identical bits on every rank
int avs_dist_allreduce_parts(AvsContext *c, const double *parts, int nparts, int count, double *out, const int *done) {
    DistState *d = (DistState *)c->dist;
    if (d && d->p2p) {
        ++d->seqReduce;
        AVS_TRACE("rank %d: p2p allreduce seq %llu (count %d)", c->rank, d->seqReduce, count);
        k_p2p_allreduce<<<1, 256, 0, c->stream>>>(parts, nparts, count, d->peerTable.as<void *>(), c->rank, c->nranks, d->seqReduce, out, done);
        ++c->launches;
        return AVS_OK;
    }
    k_reduce_parts<<<1, 256, 0, c->stream>>>(parts, nparts, count, out);
    ++c->launches;
    if (d && d->local) { c->lastError = "the in-process group has no NCCL hot loop"; return AVS_ERR_UNSUPPORTED; }
    if (d) NCCL_CHECK(c, g_nccl.AllReduce(out, out, (size_t)count, ncclFloat64, ncclSum, d->comm, c->stream));
    return AVS_OK;
}

// every rank ends up with the full solution vector (rank q's block at its row offset)
int avs_dist_allgather_solution(AvsContext *c, const double *local, double *full) {
    DistState *d = (DistState *)c->dist;
    const int P = c->nranks;
    if (!d) return AVS_OK;
    AVS_CUDA_CHECK(cudaMemcpyAsync(full + c->rowBegin, local, (size_t)(c->rowEnd - c->rowBegin) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    size_t off[P2P_MAX_RANKS] = {}, len[P2P_MAX_RANKS] = {};
    for (int q = 0; q < P && q < P2P_MAX_RANKS; ++q) {
        off[q] = (size_t)c->rowStarts[q] * sizeof(double);
        len[q] = (size_t)(c->rowStarts[q + 1] - c->rowStarts[q]) * sizeof(double);
    }
    return commGatherSegments(c, d, full, off, len, ncclFloat64, sizeof(double));
}

int avs_dist_allreduce_u64(AvsContext *c, unsigned long long *d, int count) {
    DistState *ds = (DistState *)c->dist;
    if (!ds) return AVS_OK;
    return commAllReduceU64(c, ds, d, count, 0);
}

// Stage 11 is sharded by z-slabs of the regular grid (avs_slab_cuts): rank q has filled planes [z0_q, z1_q) of every axis.
// In-place gathers from the owners: after this every rank holds the whole velocity field (solveGasSubclass
// updates `vel` in place, AV.cpp:698 -- a caller on any rank sees the complete result).
int avs_dist_allgather_slabs(AvsContext *c, float *dOut[3]) {
    DistState *ds = (DistState *)c->dist;
    if (!ds) return AVS_OK;
    for (int a = 0; a < 3; ++a) {
        const size_t plane = (size_t)c->S.regular[a].n[0] * c->S.regular[a].n[1];
        size_t off[P2P_MAX_RANKS] = {}, len[P2P_MAX_RANKS] = {};
        for (int q = 0; q < c->nranks && q < P2P_MAX_RANKS; ++q) {
            int z0, z1;
            avs_slab_range(c, a, q, &z0, &z1);
            if (z1 <= z0) continue;
            off[q] = plane * (size_t)z0 * sizeof(float);
            len[q] = plane * (size_t)(z1 - z0) * sizeof(float);
        }
        int rc = commGatherSegments(c, ds, dOut[a], off, len, ncclFloat32, sizeof(float));
        if (rc) return rc;
    }
    return AVS_OK;
}

// true when one of the per-launch exchange kernels gave up waiting for a peer (the solve's result is then garbage)
bool avs_dist_timed_out(AvsContext *c) {
    DistState *d = (DistState *)c->dist;
    if (!d || !d->p2p || !d->region) return false;
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, (char *)d->region + offsetof(P2PHeader, timedOut), sizeof(v), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return false;
    cudaStreamSynchronize(c->stream);
    if (v) cudaMemsetAsync((char *)d->region + offsetof(P2PHeader, timedOut), 0, sizeof(v), c->stream);
    return v != 0;
}

// 0 = single GPU, 1 = NCCL hot loop, 2 = peer-memory hot loop
int avs_dist_mode(AvsContext *c) {
    DistState *d = (DistState *)c->dist;
    if (!d) return 0;
    return d->p2p ? 2 : 1;
}
