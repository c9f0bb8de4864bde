// avs_context.h -- host-side state of one solver context (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/avs.h"
#include "avs_common.cuh"

// Growable device buffer: allocations are cached in the context across solves (the reference
// rebuilds everything per call, HDK_AdaptiveViscosity.cpp:233-707; here only the contents are).
// Ranks that SHARE one device (avs_create_multi with a repeated ordinal: the one-GPU test configuration) must never make a
// device-synchronising call while a peer's kernel spins on a flag they are about to raise: cudaFree / cudaMalloc would wait for
// that kernel -> deadlock.  In that mode buffers come from the stream-ordered allocator on the calling thread's context stream.
extern bool g_avsAsyncAlloc;                     // set once by avs_create_multi when two ranks share a device
extern thread_local cudaStream_t g_avsTlsStream; // stream of the context the calling thread is currently working for

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool async = false;
    void freeNow() {
        if (!p) return;
        if (async) {
            cudaStreamSynchronize(g_avsTlsStream);   // kernels of this context that still use the buffer
            cudaFreeAsync(p, g_avsTlsStream);
        } else cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        freeNow();
        size_t want = bytes + bytes / 8 + 256;
        if (g_avsAsyncAlloc) {
            if (cudaMallocAsync(&p, want, g_avsTlsStream) != cudaSuccess) {
                cudaGetLastError();
                want = bytes;
                if (cudaMallocAsync(&p, want, g_avsTlsStream) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
            }
            cudaStreamSynchronize(g_avsTlsStream);   // the context's other streams (copy stream) may use the buffer next
            async = true;
            cap = want;
            return 0;
        }
        async = false;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
            want = bytes;
        }
        cap = want;
        return 0;
    }
    void release() { freeNow(); }
    template <class T> T *as() const { return (T *)p; }
};

// SJDS-32 "paired" sparse matrix (layout described at the top of avs_cg.cu).
struct SellMatrix {
    int64_t n = 0;          // rows (local)
    int64_t nnz = 0;        // true non-zeros (for the roofline figure)
    int64_t nslices = 0;
    int64_t padded = 0;     // stored entries
    DevBuf sliceOff;        // int64[nslices+1], in pairs
    DevBuf val;             // double2 / float2 [padded/2]
    DevBuf col;             // int2 [padded/2]
    DevBuf invDiag;         // Jacobi preconditioner (DiagonalPreconditioner)
    DevBuf meta;            // int32[nslices*32]: (original lane << 8) | pairs, per sorted lane
    int precision = AVS_PRECISION_F64;
};

struct CgWork {
    DevBuf x, r, p, t;      // CG vectors
    DevBuf partials;        // per-CTA partial sums
    DevBuf scalars;         // device-resident CG scalars
    DevBuf pcgState;        // PcgState of the persistent CG kernel
    DevBuf sliceHalo, sliceFlag, sliceIndex, boundaryList;  // multi-GPU: slices that read a halo slot (flag u8 / i32, scan, ascending list)
    DevBuf pcgLocal;        // single GPU: the persistent kernel's mailbox header + one-entry peer table
};

struct AvsContext {
    int device = 0;
    int rank = 0, nranks = 1;
    void *dist = nullptr;            // DistState (avs_dist.cu) when nranks > 1
    void *localGroup = nullptr;      // LocalGroup (avs_dist.cu) when the ranks live in one process (avs_create_multi)
    int deviceShare = 1;             // ranks of the group that run on this rank's device (cooperative grids must co-reside)
    bool slabOutputOnly = false;     // multi-GPU host output: every rank downloads only its z-slab into the caller's arrays
    bool outputIsHost = false;       // set by runApply: the velocity of this call goes to host memory
    long long nHalo = 0;              // off-rank columns referenced by the local rows
    const long long *haloIndex = nullptr;  // device: global column -> halo slot (valid where flagged)
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    cudaStream_t copyStream = nullptr;   // host->device copies that overlap the labelling stages
    cudaEvent_t evCopyGate = nullptr, evUploadDone = nullptr;
    cudaEvent_t evAxis[3] = {}, evDownloadDone = nullptr;   // per-axis device -> host copies of the output under stage 11
    bool timeSpmv = false;
    int numSMs = 148;
    std::string lastError;

    // ---- device copies of the inputs
    DevBuf inSurface, inVel[3], inFaceW[3], inVisc, inDens, inColl, inCollVel[3];
    // ---- stage outputs
    DevBuf centerW, edgeW[3];
    DevBuf label[AVS_MAX_LEVELS], face[AVS_MAX_LEVELS][3], edge[AVS_MAX_LEVELS][3], center[AVS_MAX_LEVELS], regular[3];
    DevBuf tiles;           // tile-occupancy scratch
    DevBuf nodes[AVS_MAX_LEVELS], nodeScene;  // interpolator node pyramid (stage 11)
    DevBuf bandList;        // indices of the samples whose neighbourhood changes sign (supersampled)
    DevBuf signA, signB;    // sign classes of the surface SDF neighbourhoods (weights shortcut)
    DevBuf brickCount, brickCost, brickCostPrefix, brickOffset, scanTmp, counters;
    DevBuf rowKeys;         // RowKey[n]
    DevBuf rowMass;                         // fp64 [local rows]: M_u = rho V_face of every row (rhs += M_u u^n, k_finish_rhs)
    DevBuf faceWRow;                        // fp32 [N]: the face weight of every level-0 row (k_gather_face_weights)
    DevBuf asmRowList;                      // split assembly: local ids of the rows left for the generic kernel
    DevBuf coarseRows;                      // restriction of the rows of level >= 3: row ids | work-item bases | work items | partial sums
    DevBuf x0, rhs, solution, fullSolution;  // fp64: x0 [N], rhs/solution [local rows], fullSolution [N] (multi-GPU apply)
    DevBuf rowCount, rowOffset; // assembly scratch (int32 / int64)
    DevBuf csrPtr, csrCol, csrVal;  // canonical CSR of the local rows (int64 / int32 / fp64); col/val built lazily
    DevBuf stageCol, stageVal, diag; // assembly staging area (column-major, MAX_ROW deep) and the diagonal
    Grid3<uint8_t> tileFlagsGrid = {nullptr, {0, 0, 0}};   // view of tileFlags for the current solve (d == nullptr: dense labelling ran)
    DevBuf nodeTileList;                 // level-0 node tiles of the interpolator's pyramid (avs_prolong.cu)
    DevBuf tileFlags, tileLists;         // level-0 labelling: per-16^3-tile flags (ACTIVE / non-UP present) and the compacted tile lists
    DevBuf slicePairs, edgeTiles, solidW; // per-solve scratch kept across calls (no cudaMalloc/cudaFree in steady state)
    DevBuf geoCount, geoOffset, geoPos, geoScale, geoLevel;  // octree geometry dump (OG.cpp:245-308), built on request
    bool haveOctree = false;
    long long stageStride = 0;
    bool csrValid = false;
    DeviceScene S;          // host copy of the descriptor handed to kernels
    bool haveSystem = false, haveSolution = false;

    int levelsAllocated = 0;
    int64_t nRows = 0, nnz = 0, nEdge = 0, nCenter = 0, nRegular = 0;
    int64_t rowBegin = 0, rowEnd = 0;  // rows owned by this rank
    bool surfaceTilesMarked = false;   // the weights stage has already filled the level-0 face tile maps (c->tiles) for this solve
    long long asmGenericRows = 0;      // rows of the last assembly that went through the generic kernel (split assembly: the non-simple ones)
    int faceWMappedBytes = 0;          // 1: the face weights of this solve are read through mapped host pointers (no bulk upload)
    bool x0AllRows = true;             // x0 holds the restricted velocity of every row (false: only of [rowBegin, rowEnd))
    std::vector<int> slabZ;            // [nranks+1] z-plane cuts of the regular grid (stage 4 labels + stage 11 are slab-sharded)
    std::vector<long long> rowStarts;  // [nranks+1] first row of every rank's block (brick-granular, cost-balanced, identical on all ranks)

    SellMatrix A;
    CgWork cg;
    DevBuf cgRhs;
    void *hostScalars = nullptr;   // pinned, 2 x CgScalars
    cudaEvent_t evPoll[2] = {};
    std::vector<cudaEvent_t> spmvEvents;  // pairs, only when timeSpmv
    size_t spmvEventsUsed = 0;
    std::vector<cudaEvent_t> auxEvents;   // triples around the two CG vector kernels, only when timeSpmv
    size_t auxEventsUsed = 0;

    int64_t launches = 0, spmvLaunches = 0;
    // persistent CG kernel: phase times measured in-kernel (%globaltimer, barrier to barrier)
    bool pcgUsed = false;
    float pcgSpmvMs = 0.f, pcgXrMs = 0.f, pcgPMs = 0.f;
    float pcgKernelMs = 0.f;          // CUDA events around the cooperative launches
    int pcgLaunches = 0;
    cudaEvent_t evPcg[2] = {};
    int64_t pcgPhases = 0;
    int64_t pcgBoundarySlices = 0;
    float spmvMs = 0.f;
    cudaEvent_t ev[AVS_STAGE_COUNT + 2] = {};
};

// stage entry points implemented in the .cu files
int avs_stage_upload(AvsContext *c, const AvsFields *in, const AvsParams *p);
int avs_stage_weights(AvsContext *c, const AvsParams *p);
int avs_stage_octree(AvsContext *c, const AvsParams *p);
int avs_stage_regular_labels(AvsContext *c);
int avs_octree_points(AvsContext *c, int64_t *countOut);
int avs_slab_cuts(AvsContext *c);
void avs_slab_range(const AvsContext *c, int axis, int q, int *z0, int *z1);
int avs_stage_octree_labels(AvsContext *c);
int avs_finish_rhs(AvsContext *c);   // rhs += M_u u^n once the restriction has run (after the assembly)
int avs_stage_restriction(AvsContext *c, bool allRows = false);   // nranks > 1: only the rows this rank owns unless allRows
int avs_stage_system(AvsContext *c, const AvsParams *p);
int avs_stage_solve(AvsContext *c, const AvsParams *p, AvsResult *res);
int avs_apply_regular(AvsContext *c, float *dOut[3], unsigned long long *hostPending, float *const *hostOut);

int avs_sell_from_csr(AvsContext *c, SellMatrix &A, int64_t n, const int64_t *dPtr, const int32_t *dCol,
                      const double *dVal, int precision);
int avs_cg_run(AvsContext *c, SellMatrix &A, const double *dRhs, const double *dX0, double *dXout,
               const AvsParams *p, AvsResult *res);
int avs_sell_from_stage(AvsContext *c, SellMatrix &A, int64_t n, int64_t nnz, const int32_t *dCount, const int32_t *dStageCol,
                        const double *dStageVal, long long stride, const double *dDiag, int precision,
                        long long rowBegin, long long rowEnd, const long long *haloIndex);
int avs_build_csr(AvsContext *c);
int avs_dist_init(AvsContext *c, const void *uniqueId128);
void *avs_local_group_create(int P, const int *devices);
void avs_local_group_destroy(void *g);
void avs_local_group_fail(void *g);
void avs_local_group_reset(void *g);
void avs_dist_destroy(AvsContext *c);
int avs_dist_build_halo(AvsContext *c);
int avs_dist_halo_exchange(AvsContext *c, void *p, int precision, const int *done);
void *avs_dist_prepare_p(AvsContext *c, int *rcOut);
int avs_dist_mode(AvsContext *c);
bool avs_dist_timed_out(AvsContext *c);
struct PcgDist;
bool avs_dist_pcg_args(AvsContext *c, PcgDist *out);
void avs_dist_pcg_commit(AvsContext *c, unsigned long long seqPush, unsigned long long seqReduce);
int avs_dist_allreduce_parts(AvsContext *c, const double *parts, int nparts, int count, double *out, const int *done);
int avs_dist_allgather_solution(AvsContext *c, const double *local, double *full);
int avs_dist_allreduce_u64(AvsContext *c, unsigned long long *d, int count);   // in-place sum over ranks (setup paths only)
int avs_dist_allgather_slabs(AvsContext *c, float *dOut[3]);                  // every rank's z-slab of the regular output -> all ranks
int avs_spmv_time(AvsContext *c, SellMatrix &A, int repeats, float *msPerLaunch);
int avs_spmv_once(AvsContext *c, SellMatrix &A, const double *dX, double *dY);

// generic device helpers (avs_labels.cu)
int avs_exclusive_scan_i32_to_i64(AvsContext *c, const int32_t *dIn, int64_t *dOut, int64_t n, int64_t *hostTotal);
