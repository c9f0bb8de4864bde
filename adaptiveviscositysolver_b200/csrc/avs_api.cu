// avs_api.cu -- the extern "C" surface declared in include/avs.h: context life cycle, the staged
// pipeline that replaces HDK_AdaptiveViscosity.cpp:233-707, and host read-back for parity tests.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler (nsys, ncu --nvtx) is attached

#include "avs_context.h"

// NVTX range per pipeline stage, named after the reference's UT_PerfMonAutoSolveEvent labels (HDK_AdaptiveViscosity.cpp:306, 360,
// 441, 473, 516, 554, 603, 662, 759, 776, 813, 874): a timeline of this library reads like Houdini's performance monitor.
struct NvtxRange {
    explicit NvtxRange(const char *label) { nvtxRangePushA(label); }
    ~NvtxRange() { nvtxRangePop(); }
};

static thread_local char g_lastError[512] = "";
bool g_avsAsyncAlloc = false;
thread_local cudaStream_t g_avsTlsStream = nullptr;
// every entry point: bind the calling thread to the context's device and stream
static inline void enterContext(AvsContext *c) {
    cudaSetDevice(c->device);
    g_avsTlsStream = c->stream;
}

// leaving an entry point: a context-level message (c->lastError, e.g. why a scene is AVS_ERR_UNSUPPORTED) becomes the text
// avs_last_error() returns on this thread; AVS_ERR_CUDA keeps the text of the failed CUDA call (avs_set_last_error)
static inline int leaveContext(AvsContext *c, int rc) {
    if (rc != AVS_OK && rc != AVS_ERR_CUDA && !c->lastError.empty()) snprintf(g_lastError, sizeof(g_lastError), "%s", c->lastError.c_str());
    // a stage that fails returns before the pipeline joins its copy stream: the caller's host arrays must not be in flight any more
    // when the call returns (success paths have waited for evUploadDone / evDownloadDone already)
    if (rc != AVS_OK && c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaGetLastError(); }
    return rc;
}

void avs_set_last_error(const char *what, cudaError_t e, const char *file, int line) {
    snprintf(g_lastError, sizeof(g_lastError), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    cudaGetLastError();
}

extern "C" {

int avs_abi_version(void) { return AVS_ABI_VERSION; }

int avs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *avs_last_error(void) { return g_lastError; }

const char *avs_status_string(int s) {
    switch (s) {
        case AVS_OK: return "ok";
        case AVS_ERR_INVALID_ARGUMENT: return "invalid argument";
        case AVS_ERR_MISSING_FIELD: return "field missing";
        case AVS_ERR_MISALIGNED_FIELD: return "field not aligned";
        case AVS_ERR_ALLOC: return "device allocation failed";
        case AVS_ERR_CUDA: return "CUDA error";
        case AVS_ERR_NCCL: return "NCCL error";
        case AVS_ERR_CANCELLED: return "cancelled";
        case AVS_ERR_BREAKDOWN: return "conjugate gradient breakdown";
        case AVS_ERR_NO_DEVICE: return "no CUDA device";
        case AVS_ERR_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown status";
}

void avs_default_params(AvsParams *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->size = sizeof(AvsParams);
    p->dt = 1.0 / 24.0;
    p->tolerance = 1e-3;
    p->extrapolation = 0.5;
    p->max_iterations = 2500;
    p->number_super_samples = 3;
    p->octree_levels = 4;
    p->fine_bandwidth = 0;
    p->use_enhanced_gradients = 1;
    p->do_apply_solid_weights = 0;
    p->precision = AVS_PRECISION_F64;
    p->check_every = 0;
    p->cancel = nullptr;
}

static int createImpl(const AvsDeviceConfig *cfg, void *localGroup, int deviceShare, AvsContext **out);
int avs_create(const AvsDeviceConfig *cfg, AvsContext **out) { return createImpl(cfg, nullptr, 1, out); }

static int createImpl(const AvsDeviceConfig *cfg, void *localGroup, int deviceShare, AvsContext **out) {
    if (!out) return AVS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        snprintf(g_lastError, sizeof(g_lastError), "no CUDA device visible: this library has no CPU path");
        return AVS_ERR_NO_DEVICE;
    }
    AvsContext *c = new (std::nothrow) AvsContext();
    if (!c) return AVS_ERR_ALLOC;
    if (cfg) {
        if (cfg->size != sizeof(AvsDeviceConfig)) { delete c; return AVS_ERR_INVALID_ARGUMENT; }
        c->device = cfg->device;
        c->rank = cfg->rank;
        c->nranks = std::max(1, cfg->nranks);
        c->timeSpmv = cfg->time_spmv != 0;
        c->slabOutputOnly = cfg->distributed_output != 0 && cfg->nranks > 1;
        c->stream = (cudaStream_t)cfg->stream;
    }
    c->localGroup = localGroup;
    c->deviceShare = std::max(1, deviceShare);
    if (c->device < 0 || c->device >= ndev || c->rank < 0 || c->rank >= c->nranks) { delete c; return AVS_ERR_INVALID_ARGUMENT; }
    if (cudaSetDevice(c->device) != cudaSuccess) { delete c; return AVS_ERR_CUDA; }
    if (!c->stream) {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return AVS_ERR_CUDA; }
        c->ownStream = true;
    }
    g_avsTlsStream = c->stream;
    cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, c->device);
    for (auto &e : c->ev) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->evCopyGate, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evUploadDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evDownloadDone, cudaEventDisableTiming);
    for (auto &e : c->evAxis) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evPoll[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evPoll[1], cudaEventDisableTiming);
    cudaEventCreate(&c->evPcg[0]);
    cudaEventCreate(&c->evPcg[1]);
    if (cudaHostAlloc(&c->hostScalars, 512, cudaHostAllocDefault) != cudaSuccess) { c->hostScalars = nullptr; avs_destroy(c); return AVS_ERR_ALLOC; }
    if (c->timeSpmv) {
        c->spmvEvents.resize(2 * 4096);
        for (auto &e : c->spmvEvents) cudaEventCreate(&e);
        c->auxEvents.resize(3 * 4096);
        for (auto &e : c->auxEvents) cudaEventCreate(&e);
    }
    if (c->counters.reserve(64 * sizeof(unsigned long long))) { avs_destroy(c); return AVS_ERR_ALLOC; }
    memset(&c->S, 0, sizeof(c->S));
    if (c->nranks > 1) {
        int rc = avs_dist_init(c, cfg ? cfg->nccl_unique_id : nullptr);
        if (rc != AVS_OK) {
            snprintf(g_lastError, sizeof(g_lastError), "%s", c->lastError.c_str());
            avs_destroy(c);
            return rc;
        }
    }
    *out = c;
    return AVS_OK;
}

void avs_destroy(AvsContext *c) {
    if (!c) return;
    enterContext(c);
    cudaStreamSynchronize(c->stream);
    avs_dist_destroy(c);
    DevBuf *bufs[] = {&c->fullSolution, &c->inSurface, &c->inVisc, &c->inDens, &c->inColl, &c->centerW, &c->tiles, &c->brickCount, &c->brickOffset,
                      &c->scanTmp, &c->counters, &c->rowKeys, &c->coarseRows, &c->asmRowList, &c->faceWRow, &c->rowMass, &c->x0, &c->rhs, &c->solution, &c->rowCount, &c->rowOffset,
                      &c->csrPtr, &c->csrCol, &c->csrVal, &c->A.sliceOff, &c->A.val, &c->A.col, &c->A.invDiag, &c->A.meta, &c->cg.pcgState, &c->cg.sliceHalo, &c->cg.sliceFlag, &c->cg.sliceIndex, &c->cg.boundaryList, &c->cg.pcgLocal,
                      &c->cg.x, &c->cg.r, &c->cg.p, &c->cg.t, &c->cg.partials, &c->cg.scalars, &c->cgRhs,
                      &c->stageCol, &c->stageVal, &c->diag, &c->slicePairs, &c->edgeTiles, &c->tileFlags, &c->tileLists, &c->nodeTileList, &c->solidW, &c->signA, &c->signB, &c->bandList, &c->brickCost, &c->brickCostPrefix,
                      &c->geoCount, &c->geoOffset, &c->geoPos, &c->geoScale, &c->geoLevel};
    for (DevBuf *b : bufs) b->release();
    for (int a = 0; a < 3; ++a) {
        c->inVel[a].release(); c->inFaceW[a].release(); c->inCollVel[a].release(); c->edgeW[a].release(); c->regular[a].release();
    }
    c->nodeScene.release();
    for (int l = 0; l < AVS_MAX_LEVELS; ++l) {
        c->nodes[l].release();
        c->label[l].release(); c->center[l].release();
        for (int a = 0; a < 3; ++a) { c->face[l][a].release(); c->edge[l][a].release(); }
    }
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->evPoll) if (e) cudaEventDestroy(e);
    for (auto &e : c->evPcg) if (e) cudaEventDestroy(e);
    for (auto &e : c->spmvEvents) if (e) cudaEventDestroy(e);
    for (auto &e : c->auxEvents) if (e) cudaEventDestroy(e);
    if (c->evCopyGate) cudaEventDestroy(c->evCopyGate);
    if (c->evUploadDone) cudaEventDestroy(c->evUploadDone);
    if (c->evDownloadDone) cudaEventDestroy(c->evDownloadDone);
    for (auto &e : c->evAxis) if (e) cudaEventDestroy(e);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->hostScalars) cudaFreeHost(c->hostScalars);
    if (c->ownStream) cudaStreamDestroy(c->stream);
    delete c;
}

}  // extern "C"

// ---- stage 0: host -> device ------------------------------------------------------------------
static int uploadField(AvsContext *c, DevBuf &buf, DField &d, const AvsField &f, cudaStream_t stream = nullptr) {
    if (!stream) stream = c->stream;
    for (int a = 0; a < 3; ++a) { d.n[a] = f.res[a]; d.org[a] = f.org[a]; }
    d.dx = f.dx;
    d.constant = f.constant;
    if (!f.data) {
        d.d = nullptr;
        d.n[0] = d.n[1] = d.n[2] = 1;
        return AVS_OK;
    }
    if (f.res[0] <= 0 || f.res[1] <= 0 || f.res[2] <= 0 || !(f.dx > 0)) return AVS_ERR_INVALID_ARGUMENT;
    if (f.on_device) {  // already resident in HBM: read it in place, the kernels never write input fields
        d.d = f.data;
        return AVS_OK;
    }
    size_t bytes = (size_t)f.res[0] * f.res[1] * f.res[2] * sizeof(float);
    if (buf.reserve(bytes)) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(buf.p, f.data, bytes, cudaMemcpyHostToDevice, stream));
    d.d = buf.as<float>();
    return AVS_OK;
}

static bool aligned(const AvsField &f, const int res[3], const double org[3], double dx) {
    if (!f.data) return true;  // constant fields are aligned with anything
    for (int a = 0; a < 3; ++a) {
        if (f.res[a] != res[a]) return false;
        if (std::fabs(f.org[a] - org[a]) > 1e-6 * dx) return false;
    }
    return std::fabs(f.dx - dx) <= 1e-9 * dx;
}

int avs_stage_upload(AvsContext *c, const AvsFields *in, const AvsParams *p) {
    if (!in || in->size != sizeof(AvsFields) || !p || p->size != sizeof(AvsParams)) return AVS_ERR_INVALID_ARGUMENT;
    if (in->res[0] <= 0 || in->res[1] <= 0 || in->res[2] <= 0 || !(in->dx > 0)) return AVS_ERR_INVALID_ARGUMENT;
    if (p->number_super_samples < 1 || p->octree_levels < 1 || p->max_iterations < 0) return AVS_ERR_INVALID_ARGUMENT;
    // the reference's validation (AV.cpp:152-229): missing fields and alignment
    if (!in->surface.data) return AVS_ERR_MISSING_FIELD;
    for (int a = 0; a < 3; ++a)
        if (!in->vel[a].data || !in->face_weights[a].data) return AVS_ERR_MISSING_FIELD;
    double corg[3];
    for (int a = 0; a < 3; ++a) corg[a] = in->origin[a] + 0.5 * in->dx;
    if (!aligned(in->surface, in->res, corg, in->dx)) return AVS_ERR_MISALIGNED_FIELD;
    if (!aligned(in->viscosity, in->res, corg, in->dx)) return AVS_ERR_MISALIGNED_FIELD;  // AV.cpp:210
    if (!aligned(in->density, in->res, corg, in->dx)) return AVS_ERR_MISALIGNED_FIELD;    // AV.cpp:225
    for (int a = 0; a < 3; ++a) {
        int fres[3] = {in->res[0], in->res[1], in->res[2]};
        fres[a] += 1;
        double forg[3];
        for (int k = 0; k < 3; ++k) forg[k] = in->origin[k] + (k == a ? 0.0 : 0.5 * in->dx);
        if (!aligned(in->vel[a], fres, forg, in->dx)) return AVS_ERR_MISALIGNED_FIELD;           // AV.cpp:157 (face sampled)
        if (!aligned(in->face_weights[a], fres, forg, in->dx)) return AVS_ERR_MISALIGNED_FIELD;  // AV.cpp:169
    }
    DeviceScene &S = c->S;
    for (int a = 0; a < 3; ++a) { S.N[a] = in->res[a]; S.origin[a] = in->origin[a]; }
    S.dx0 = (double)(float)in->dx;  // getVoxelSize() is a float32 vector (AV.cpp:242)
    S.dt = p->dt;
    S.extrap = S.dx0 * p->extrapolation;  // AV.cpp:243
    S.enhanced = p->use_enhanced_gradients ? 1 : 0;
    int rc;
    // Labelling (weights, octree, DOF labels) only reads the surface and the collision SDF: those go first on the
    // compute stream.  Velocity and face weights are first read by the restriction / assembly stages, so their
    // host->device copies run on the copy stream underneath the labelling kernels (uploadDone gates restriction).
    if ((rc = uploadField(c, c->inSurface, S.surface, in->surface))) return rc;
    if ((rc = uploadField(c, c->inColl, S.collision, in->collision))) return rc;
    if ((rc = uploadField(c, c->inVisc, S.viscosity, in->viscosity))) return rc;
    if ((rc = uploadField(c, c->inDens, S.density, in->density))) return rc;
    for (int a = 0; a < 3; ++a)
        if ((rc = uploadField(c, c->inCollVel[a], S.collisionVel[a], in->collision_vel[a]))) return rc;
    // the copy stream must not overwrite buffers a previous solve on the compute stream may still be reading
    AVS_CUDA_CHECK(cudaEventRecord(c->evCopyGate, c->stream));
    AVS_CUDA_CHECK(cudaStreamWaitEvent(c->copyStream, c->evCopyGate, 0));
    for (int a = 0; a < 3; ++a)
        if ((rc = uploadField(c, c->inVel[a], S.vel[a], in->vel[a], c->copyStream))) return rc;
    // Face weights are read once per LEVEL-0 ROW (k_gather_face_weights, avs_system.cu) -- a band around the surface, ~2 % of the
    // three dense arrays.  Pinned (page-locked / registered) host arrays are therefore not copied at all: the gather kernel reads
    // the values it needs through the mapped host pointer.  Pageable arrays cannot be mapped and are uploaded whole as before;
    // AVS_FACEW_UPLOAD=bulk forces that path (A/B measurements).
    static int bulk = -1;
    if (bulk < 0) { const char *e = getenv("AVS_FACEW_UPLOAD"); bulk = (e && strcmp(e, "bulk") == 0) ? 1 : 0; }
    c->faceWMappedBytes = 0;
    for (int a = 0; a < 3; ++a) {
        const AvsField &f = in->face_weights[a];
        const void *mapped = nullptr;
        if (!bulk && f.data && !f.on_device) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, f.data) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
                mapped = attr.devicePointer;
            else cudaGetLastError();
        }
        if (mapped) {
            AvsField g = f;
            g.data = (const float *)mapped;
            g.on_device = 1;   // device-addressable: read in place
            if ((rc = uploadField(c, c->inFaceW[a], S.faceW[a], g, c->copyStream))) return rc;
            c->faceWMappedBytes = 1;
        } else if ((rc = uploadField(c, c->inFaceW[a], S.faceW[a], f, c->copyStream))) return rc;
    }
    AVS_CUDA_CHECK(cudaEventRecord(c->evUploadDone, c->copyStream));
    return AVS_OK;
}

static void fillCounts(AvsContext *c, AvsResult *res) {
    res->levels = c->S.levels;
    res->octree_dofs = c->nRows;
    res->regular_dofs = c->nRegular;
    res->edge_dofs = c->nEdge;
    res->center_dofs = c->nCenter;
    res->nnz = c->nnz;
    res->local_rows = c->rowEnd - c->rowBegin;
    res->kernel_launches = c->launches;
    res->spmv_launches = c->spmvLaunches;
    res->dist_mode = avs_dist_mode(c);
    res->halo_columns = c->nHalo;
}

static void collectStageTimes(AvsContext *c, AvsResult *res, int first, int last) {
    // ev[i] is recorded at the START of stage i; ev[last+1] after the last stage
    cudaEventSynchronize(c->ev[last + 1]);
    float total = 0;
    for (int i = first; i <= last; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]);
        res->stage_ms[i] = ms;
        total += ms;
    }
    res->stage_ms[AVS_STAGE_TOTAL] += total;
}

static int runAssemble(AvsContext *c, const AvsFields *in, const AvsParams *p, AvsResult *res, bool octreeOnly = false) {
    int rc;
    c->haveSystem = c->haveSolution = c->haveOctree = false;
    cudaEventRecord(c->ev[AVS_STAGE_UPLOAD], c->stream);
    { NvtxRange nv("Upload Fields (host -> device)"); AVS_TRACE("avs_stage_upload"); if ((rc = avs_stage_upload(c, in, p))) return rc; }
    cudaEventRecord(c->ev[AVS_STAGE_SURFACE_WEIGHTS], c->stream);
    { NvtxRange nv("Compute Surface Weights"); AVS_TRACE("avs_stage_weights"); if ((rc = avs_stage_weights(c, p))) return rc; }   // + "Compute Collision Weights"
    cudaEventRecord(c->ev[AVS_STAGE_OCTREE], c->stream);
    { NvtxRange nv("Build Mask for Octree + Build Octree"); AVS_TRACE("avs_stage_octree"); if ((rc = avs_stage_octree(c, p))) return rc; }
    c->haveOctree = true;
    cudaEventRecord(c->ev[AVS_STAGE_REGULAR_LABELS], c->stream);
    if (octreeOnly) {   // onlyPrintOctree (AV.cpp:292-293): the reference returns right after the geometry dump
        collectStageTimes(c, res, AVS_STAGE_UPLOAD, AVS_STAGE_OCTREE);
        res->levels = c->S.levels;
        return AVS_OK;
    }
    { NvtxRange nv("Build Regular Grid Velocity Labels"); AVS_TRACE("avs_stage_regular_labels"); if ((rc = avs_stage_regular_labels(c))) return rc; }
    cudaEventRecord(c->ev[AVS_STAGE_OCTREE_LABELS], c->stream);
    { NvtxRange nv("Build Octree Velocity and Stress Labels"); AVS_TRACE("avs_stage_octree_labels"); if ((rc = avs_stage_octree_labels(c))) return rc; }
    // The assembly needs everything but the velocity (u^n only enters the right-hand side, through its mass term), so it runs BEFORE
    // the restriction: the velocity's host -> device copy has labelling + assembly to hide under (avs_stage_upload).
    cudaEventRecord(c->ev[AVS_STAGE_SYSTEM], c->stream);
    {
        NvtxRange nvSystem("Build Edge/Cell Stress Stencils + Build Octree Linear System");
        AVS_TRACE("avs_stage_system"); if ((rc = avs_stage_system(c, p))) return rc;
        // the CG's matrix format is part of "Build Octree Linear System" (the reference's setFromTriplets, AV.cpp:614)
        c->nHalo = 0;
        c->haloIndex = nullptr;
        AVS_TRACE("sjds build");
        if (c->nranks > 1 && (rc = avs_dist_build_halo(c))) return rc;
        if ((rc = avs_sell_from_stage(c, c->A, c->rowEnd - c->rowBegin, c->nnz, c->rowCount.as<int32_t>(), c->stageCol.as<int32_t>(),
                                      c->stageVal.as<double>(), c->stageStride, c->diag.as<double>(), p->precision,
                                      c->rowBegin, c->rowEnd, c->haloIndex))) return rc;
    }
    cudaStreamWaitEvent(c->stream, c->evUploadDone, 0);  // the velocity has landed
    cudaEventRecord(c->ev[AVS_STAGE_RESTRICTION], c->stream);
    {
        NvtxRange nv("Interpolate Regular Grid Velocities at Octree Velocity Faces");
        AVS_TRACE("avs_stage_restriction");
        if ((rc = avs_stage_restriction(c))) return rc;
        if ((rc = avs_finish_rhs(c))) return rc;
    }
    cudaEventRecord(c->ev[AVS_STAGE_SOLVE], c->stream);
    collectStageTimes(c, res, AVS_STAGE_UPLOAD, AVS_STAGE_REGULAR_LABELS);
    {   // octree labels -> system -> restriction -> (solve): events in time order
        const int order[4] = {AVS_STAGE_OCTREE_LABELS, AVS_STAGE_SYSTEM, AVS_STAGE_RESTRICTION, AVS_STAGE_SOLVE};
        cudaEventSynchronize(c->ev[AVS_STAGE_SOLVE]);
        for (int k = 0; k < 3; ++k) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[order[k]], c->ev[order[k + 1]]);
            res->stage_ms[order[k]] = ms;
            res->stage_ms[AVS_STAGE_TOTAL] += ms;
        }
    }
    fillCounts(c, res);
    return AVS_OK;
}

static int runSolve(AvsContext *c, const AvsParams *p, AvsResult *res) {
    if (!c->haveSystem) return AVS_ERR_INVALID_ARGUMENT;
    const long long n = c->rowEnd - c->rowBegin;
    if (c->solution.reserve((size_t)std::max<long long>(n, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    c->spmvEventsUsed = 0;
    c->auxEventsUsed = 0;
    res->cg_kernel_ms = 0.f;
    res->cg_kernel_launches = 0;
    c->spmvMs = 0;
    const int64_t spmv0 = c->spmvLaunches;
    cudaEventRecord(c->ev[AVS_STAGE_SOLVE], c->stream);
    AVS_TRACE("avs_cg_run");
    NvtxRange nvSolve("Solve Linear System");
    int rc = avs_cg_run(c, c->A, c->rhs.as<double>(), c->x0.as<double>() + c->rowBegin, c->solution.as<double>(), p, res);
    cudaEventRecord(c->ev[AVS_STAGE_SOLVE + 1], c->stream);
    collectStageTimes(c, res, AVS_STAGE_SOLVE, AVS_STAGE_SOLVE);
    // Launches after convergence are no-ops (device-side `done` flag): only the launches that did work count.
    // Event pair 0 is the residual SpMV (r = b - A x0), pairs 1..iters+1 are the CG iterations incl. the one that broke out.
    const int64_t realCg = std::min<int64_t>((int64_t)res->iterations + 1, std::max<int64_t>(p->max_iterations, 0));
    if (c->pcgUsed) {   // persistent kernel: phase times come from its own %globaltimer stamps
        res->spmv_ms = c->pcgSpmvMs;
        res->spmv_launches = c->pcgPhases;
        res->cg_update_xr_ms = c->pcgXrMs;
        res->cg_update_p_ms = c->pcgPMs;
        res->cg_kernel_ms = c->pcgKernelMs;
        res->cg_kernel_launches = c->pcgLaunches;
    } else if (c->timeSpmv) {
        const size_t pairs = std::min<size_t>(c->spmvEventsUsed / 2, (size_t)(realCg + 1));
        float total = 0;
        for (size_t i = 0; i < pairs; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c->spmvEvents[2 * i], c->spmvEvents[2 * i + 1]);
            total += ms;
        }
        res->spmv_ms = total;
        res->spmv_launches = (int64_t)pairs;
        const size_t triples = std::min<size_t>(c->auxEventsUsed / 3, (size_t)realCg);
        float xr = 0, pu = 0;
        for (size_t i = 0; i < triples; ++i) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, c->auxEvents[3 * i], c->auxEvents[3 * i + 1]);
            cudaEventElapsedTime(&b, c->auxEvents[3 * i + 1], c->auxEvents[3 * i + 2]);
            xr += a;
            pu += b;
        }
        res->cg_update_xr_ms = xr;
        res->cg_update_p_ms = pu;
    } else {
        res->spmv_launches = realCg + 1;
    }
    (void)spmv0;
    res->kernel_launches = c->launches;
    if (rc == AVS_OK) c->haveSolution = true;
    return rc;
}

static int runApply(AvsContext *c, AvsVelocityOut *out, AvsResult *res) {
    if (!c->haveSolution || !out) return AVS_ERR_INVALID_ARGUMENT;
    AVS_TRACE("apply");
    NvtxRange nvApply("Apply Octree Solution to Regular Grid");
    cudaEventRecord(c->ev[AVS_STAGE_APPLY], c->stream);
    float *dOut[3];
    struct TmpBufs {   // released on every exit path
        DevBuf b[3];
        ~TmpBufs() { for (DevBuf &x : b) x.release(); }
        DevBuf &operator[](int a) { return b[a]; }
    } tmp;
    for (int a = 0; a < 3; ++a) {
        if (!out->vel[a]) return AVS_ERR_INVALID_ARGUMENT;
        size_t bytes = c->S.regular[a].count() * sizeof(float);
        if (out->on_device) dOut[a] = out->vel[a];
        else if (c->S.vel[a].d == c->inVel[a].as<float>() && c->inVel[a].p) {
            // host caller: the library's own device copy of the input velocity is the "in place" target
            // (solveGasSubclass updates `vel` in place, AV.cpp:698); untouched faces keep the input value.
            dOut[a] = c->inVel[a].as<float>();
        } else {
            if (tmp[a].reserve(bytes)) return AVS_ERR_ALLOC;
            AVS_CUDA_CHECK(cudaMemcpyAsync(tmp[a].p, out->vel[a], bytes, cudaMemcpyHostToDevice, c->stream));
            dOut[a] = tmp[a].as<float>();
        }
    }
    if (c->nranks > 1) {
        if (c->fullSolution.reserve((size_t)std::max<int64_t>(c->nRows, 1) * sizeof(double))) return AVS_ERR_ALLOC;
        int rcg = avs_dist_allgather_solution(c, c->solution.as<double>(), c->fullSolution.as<double>());
        if (rcg) return rcg;
    }
    unsigned long long pending = 0;
    c->outputIsHost = !out->on_device;
    const bool streamed = !out->on_device && c->nranks == 1;   // per-axis downloads overlap the remaining apply kernels
    int rc = avs_apply_regular(c, dOut, &pending, streamed ? out->vel : nullptr);
    if (rc) return rc;
    cudaEventRecord(c->ev[AVS_STAGE_DOWNLOAD], c->stream);
    if (streamed) AVS_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->evDownloadDone, 0));
    else if (!out->on_device && c->slabOutputOnly && c->nranks > 1) {
        // in-process group: the ranks share the caller's host arrays, every rank fills the z-slab it computed
        for (int a = 0; a < 3; ++a) {
            int z0, z1;
            avs_slab_range(c, a, c->rank, &z0, &z1);
            if (z1 <= z0) continue;
            const size_t plane = (size_t)c->S.regular[a].n[0] * c->S.regular[a].n[1];
            AVS_CUDA_CHECK(cudaMemcpyAsync(out->vel[a] + plane * (size_t)z0, dOut[a] + plane * (size_t)z0, plane * (size_t)(z1 - z0) * sizeof(float),
                                           cudaMemcpyDeviceToHost, c->stream));
        }
    } else if (!out->on_device)
        for (int a = 0; a < 3; ++a)
            AVS_CUDA_CHECK(cudaMemcpyAsync(out->vel[a], dOut[a], c->S.regular[a].count() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    cudaEventRecord(c->ev[AVS_STAGE_DOWNLOAD + 1], c->stream);
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));   // the caller's host arrays are complete on return
    collectStageTimes(c, res, AVS_STAGE_APPLY, AVS_STAGE_DOWNLOAD);
    res->kernel_launches = c->launches;
    res->interpolated_faces = (int64_t)pending;
    return AVS_OK;
}

extern "C" {

int avs_assemble(AvsContext *c, const AvsFields *in, const AvsParams *p, AvsResult *res) {
    if (!c || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    memset(res->stage_ms, 0, sizeof(res->stage_ms));
    c->launches = 0;
    c->spmvLaunches = 0;
    c->lastError.clear();
    int rc = runAssemble(c, in, p, res);
    res->status = rc;
    return leaveContext(c, rc);
}

int avs_build_octree(AvsContext *c, const AvsFields *in, const AvsParams *p, AvsResult *res) {
    if (!c || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    memset(res->stage_ms, 0, sizeof(res->stage_ms));
    c->launches = 0;
    c->spmvLaunches = 0;
    int rc = runAssemble(c, in, p, res, true);
    res->kernel_launches = c->launches;
    res->status = rc;
    return rc;
}

int avs_get_octree_points(AvsContext *c, int64_t *count, float *pos, float *pscale, int32_t *level) {
    if (!c || !count) return AVS_ERR_INVALID_ARGUMENT;
    if (!c->haveOctree) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    int64_t n = 0;
    int rc = avs_octree_points(c, &n);
    if (rc) return rc;
    *count = n;
    if (n > 0) {
        if (pos) AVS_CUDA_CHECK(cudaMemcpyAsync(pos, c->geoPos.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (pscale) AVS_CUDA_CHECK(cudaMemcpyAsync(pscale, c->geoScale.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (level) AVS_CUDA_CHECK(cudaMemcpyAsync(level, c->geoLevel.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}

int avs_solve_resident(AvsContext *c, const AvsParams *p, AvsResult *res) {
    if (!c || !p || p->size != sizeof(AvsParams) || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    fillCounts(c, res);
    int rc = runSolve(c, p, res);
    res->status = rc;
    return rc;
}

int avs_apply(AvsContext *c, AvsVelocityOut *out, AvsResult *res) {
    if (!c || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    int rc = runApply(c, out, res);
    res->status = rc;
    return rc;
}

int avs_solve(AvsContext *c, const AvsFields *in, const AvsParams *p, AvsVelocityOut *out, AvsResult *res) {
    if (!c || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    memset(res->stage_ms, 0, sizeof(res->stage_ms));
    c->launches = 0;
    c->spmvLaunches = 0;
    cudaEventRecord(c->ev[AVS_STAGE_COUNT], c->stream);
    c->lastError.clear();
    int rc = runAssemble(c, in, p, res);
    if (rc == AVS_OK) rc = runSolve(c, p, res);
    if (rc == AVS_OK && out) rc = runApply(c, out, res);
    // total = first enqueue to last completion on the stream (includes the host gaps between stages)
    cudaEventRecord(c->ev[AVS_STAGE_COUNT + 1], c->stream);
    if (cudaEventSynchronize(c->ev[AVS_STAGE_COUNT + 1]) == cudaSuccess) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[AVS_STAGE_COUNT], c->ev[AVS_STAGE_COUNT + 1]);
        res->stage_ms[AVS_STAGE_TOTAL] = ms;
    }
    res->status = rc;
    return leaveContext(c, rc);
}

// ---- read-back ------------------------------------------------------------------------------
int avs_get_sizes(AvsContext *c, int64_t *nRows, int64_t *nnz, int32_t *levels) {
    if (!c) return AVS_ERR_INVALID_ARGUMENT;
    if (nRows) *nRows = c->nRows;
    if (nnz) *nnz = c->nnz;
    if (levels) *levels = c->S.levels;
    return AVS_OK;
}

int avs_get_local_range(AvsContext *c, int64_t *b, int64_t *e) {
    if (!c) return AVS_ERR_INVALID_ARGUMENT;
    if (b) *b = c->rowBegin;
    if (e) *e = c->rowEnd;
    return AVS_OK;
}

int avs_get_output_slab(AvsContext *c, int axis, int32_t *z0, int32_t *z1) {
    if (!c || axis < 0 || axis > 2 || !z0 || !z1 || !c->haveSystem) return AVS_ERR_INVALID_ARGUMENT;
    int a = 0, b = 0;
    avs_slab_range(c, axis, c->rank, &a, &b);
    *z0 = a;
    *z1 = b;
    return AVS_OK;
}

int avs_get_row_starts(AvsContext *c, int64_t *starts) {
    if (!c || !starts || c->rowStarts.empty()) return AVS_ERR_INVALID_ARGUMENT;
    for (int q = 0; q <= c->nranks; ++q) starts[q] = c->rowStarts[q];
    return AVS_OK;
}

int avs_get_keys(AvsContext *c, int32_t *keys) {
    if (!c || !keys || !c->haveSystem) return AVS_ERR_INVALID_ARGUMENT;   // keys exist only behind a geometric system (not avs_cg_csr)
    enterContext(c);
    AVS_CUDA_CHECK(cudaMemcpyAsync(keys, c->rowKeys.p, (size_t)c->nRows * sizeof(RowKey), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}

int avs_get_system_csr(AvsContext *c, int64_t *rowPtr, int32_t *col, double *val, double *rhs, double *x0) {
    if (!c || !c->haveSystem) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    const long long n = c->rowEnd - c->rowBegin;
    if (col || val) {
        int rc = avs_build_csr(c);
        if (rc) return rc;
    }
    if (rowPtr) AVS_CUDA_CHECK(cudaMemcpyAsync(rowPtr, c->csrPtr.p, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    if (col) AVS_CUDA_CHECK(cudaMemcpyAsync(col, c->csrCol.p, (size_t)c->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (val) AVS_CUDA_CHECK(cudaMemcpyAsync(val, c->csrVal.p, (size_t)c->nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (rhs) AVS_CUDA_CHECK(cudaMemcpyAsync(rhs, c->rhs.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (x0) {
        if (!c->x0AllRows) {   // multi-GPU solves restrict only the rows they own: complete the vector for the read-back
            int rc = avs_stage_restriction(c, true);
            if (rc) return rc;
        }
        AVS_CUDA_CHECK(cudaMemcpyAsync(x0, c->x0.p, (size_t)c->nRows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}

int avs_get_solution(AvsContext *c, double *x) {
    if (!c || !c->haveSolution || !x) return AVS_ERR_INVALID_ARGUMENT;
    enterContext(c);
    AVS_CUDA_CHECK(cudaMemcpyAsync(x, c->solution.p, (size_t)(c->rowEnd - c->rowBegin) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return AVS_OK;
}

int avs_get_grid(AvsContext *c, int kind, int level, int axis, void *out, int32_t *res, int64_t *nbytes) {
    if (!c || axis < 0 || axis > 2) return AVS_ERR_INVALID_ARGUMENT;
    if (kind >= 0 && kind <= 3 && (level < 0 || level >= c->S.levels)) return AVS_ERR_INVALID_ARGUMENT;
    if (kind == 0 ? !c->haveOctree : !c->haveSystem) return AVS_ERR_INVALID_ARGUMENT;   // nothing resident yet (or a caller CSR only)
    enterContext(c);
    const DeviceScene &S = c->S;
    const void *src = nullptr;
    const int *n = nullptr;
    size_t elem = 1;
    switch (kind) {
        case 0: src = S.label[level].d; n = S.label[level].n; elem = 1; break;
        case 1: src = S.face[level][axis].d; n = S.face[level][axis].n; elem = 4; break;
        case 2: src = S.edge[level][axis].d; n = S.edge[level][axis].n; elem = 1; break;
        case 3: src = S.center[level].d; n = S.center[level].n; elem = 1; break;
        case 4: src = S.regular[axis].d; n = S.regular[axis].n; elem = 1; break;
        case 5: src = S.centerW.d; n = S.centerW.n; elem = 4; break;
        case 6: src = S.edgeW[axis].d; n = S.edgeW[axis].n; elem = 4; break;
        default: return AVS_ERR_INVALID_ARGUMENT;
    }
    if (!src) return AVS_ERR_INVALID_ARGUMENT;
    size_t bytes = (size_t)n[0] * n[1] * n[2] * elem;
    if (res) { res[0] = n[0]; res[1] = n[1]; res[2] = n[2]; }
    if (nbytes) *nbytes = (int64_t)bytes;
    if (out) {
        AVS_CUDA_CHECK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return AVS_OK;
}

// ---- stand-alone linear algebra on a caller CSR ---------------------------------------------
static int uploadCsr(AvsContext *c, int64_t n, const int64_t *rowPtr, const int32_t *col, const double *val, int precision) {
    if (n < 0 || !rowPtr) return AVS_ERR_INVALID_ARGUMENT;
    const int64_t nnz = rowPtr[n];
    if (nnz > 0 && (!col || !val)) return AVS_ERR_INVALID_ARGUMENT;
    if (c->csrPtr.reserve((size_t)(n + 1) * sizeof(int64_t)) || c->csrCol.reserve((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)) ||
        c->csrVal.reserve((size_t)std::max<int64_t>(nnz, 1) * sizeof(double)))
        return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(c->csrPtr.p, rowPtr, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) {
        AVS_CUDA_CHECK(cudaMemcpyAsync(c->csrCol.p, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        AVS_CUDA_CHECK(cudaMemcpyAsync(c->csrVal.p, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    c->nRows = n;
    c->nnz = nnz;
    c->rowBegin = 0;
    c->rowEnd = n;
    c->rowStarts.assign(2, 0);
    c->rowStarts[1] = n;
    c->haveSystem = false;  // no geometry behind this matrix
    c->csrValid = true;
    c->haveSolution = false;
    return avs_sell_from_csr(c, c->A, n, c->csrPtr.as<int64_t>(), c->csrCol.as<int32_t>(), c->csrVal.as<double>(), precision);
}

int avs_cg_csr(AvsContext *c, int64_t n, const int64_t *rowPtr, const int32_t *col, const double *val, const double *rhs, double *x,
               const AvsParams *p, AvsResult *res) {
    if (!c || !p || p->size != sizeof(AvsParams) || !res || res->size != sizeof(AvsResult) || !rhs || !x) return AVS_ERR_INVALID_ARGUMENT;
    if (c->nranks > 1) return AVS_ERR_UNSUPPORTED;  // caller matrices are single-GPU only
    enterContext(c);
    memset(res->stage_ms, 0, sizeof(res->stage_ms));
    c->launches = 0;
    c->spmvLaunches = 0;
    int rc = uploadCsr(c, n, rowPtr, col, val, p->precision);
    if (rc) { res->status = rc; return rc; }
    const size_t vb = (size_t)std::max<int64_t>(n, 1) * sizeof(double);
    if (c->rhs.reserve(vb) || c->x0.reserve(vb) || c->solution.reserve(vb)) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(c->rhs.p, rhs, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    AVS_CUDA_CHECK(cudaMemcpyAsync(c->x0.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->spmvEventsUsed = 0;
    cudaEventRecord(c->ev[AVS_STAGE_SOLVE], c->stream);
    rc = avs_cg_run(c, c->A, c->rhs.as<double>(), c->x0.as<double>(), c->solution.as<double>(), p, res);
    cudaEventRecord(c->ev[AVS_STAGE_SOLVE + 1], c->stream);
    collectStageTimes(c, res, AVS_STAGE_SOLVE, AVS_STAGE_SOLVE);
    AVS_CUDA_CHECK(cudaMemcpyAsync(x, c->solution.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    res->octree_dofs = n;
    res->nnz = c->nnz;
    res->local_rows = n;
    res->kernel_launches = c->launches;
    res->spmv_launches = c->spmvLaunches;
    res->status = rc;
    return rc;
}

int avs_spmv_csr(AvsContext *c, int64_t n, const int64_t *rowPtr, const int32_t *col, const double *val, const double *x, double *y,
                 int precision, int repeats, float *msPerLaunch) {
    if (!c || !x || !y) return AVS_ERR_INVALID_ARGUMENT;
    if (c->nranks > 1) return AVS_ERR_UNSUPPORTED;
    enterContext(c);
    int rc = uploadCsr(c, n, rowPtr, col, val, precision);
    if (rc) return rc;
    const size_t vb = (size_t)std::max<int64_t>(n, 1) * sizeof(double);
    if (c->x0.reserve(vb) || c->solution.reserve(vb)) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(c->x0.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    rc = avs_spmv_once(c, c->A, c->x0.as<double>(), c->solution.as<double>());
    if (rc) return rc;
    AVS_CUDA_CHECK(cudaMemcpyAsync(y, c->solution.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (repeats > 0 && msPerLaunch) rc = avs_spmv_time(c, c->A, repeats, msPerLaunch);
    return rc;
}

int avs_time_spmv_resident(AvsContext *c, int precision, int repeats, float *msPerLaunch, double *algorithmicBytes) {
    if (!c || !msPerLaunch || c->A.n == 0) return AVS_ERR_INVALID_ARGUMENT;
    (void)precision;
    enterContext(c);
    const double s = c->A.precision == AVS_PRECISION_F32 ? 4.0 : 8.0;
    if (algorithmicBytes) *algorithmicBytes = (double)c->A.nnz * (s + 4.0) + (double)(c->A.n + 1) * 4.0 + 2.0 * (double)c->A.n * s;
    return avs_spmv_time(c, c->A, repeats, msPerLaunch);
}

// ---- one process, several GPUs ----------------------------------------------------------------------------------------------
// The DOP surface is single-threaded (solveGasSubclass is called on one cook thread, HDK_AdaptiveViscosity.cpp:126-128).
// avs_create_multi builds one rank context per listed device; avs_solve_multi runs the row-partitioned solve on all of them --
// one short-lived host thread per rank (the per-rank pipeline has host-side waits, so the ranks must be driven concurrently;
// the calling thread blocks until all have returned) -- reading the caller's host fields and writing the caller's host
// velocity arrays (every rank downloads the z-slab it computed).  No NCCL, no CUDA IPC: peer access + plain pointers.
struct AvsMulti {
    int n = 0;
    void *group = nullptr;
    std::vector<AvsContext *> ctx;
    std::vector<int> status;
    std::vector<AvsResult> res;
};

int avs_create_multi(const int32_t *devices, int32_t n, int32_t time_spmv, AvsMulti **out) {
    if (!out || !devices || n < 1 || n > 16) return AVS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    AvsMulti *m = new (std::nothrow) AvsMulti();
    if (!m) return AVS_ERR_ALLOC;
    m->n = n;
    std::vector<int> devs(devices, devices + n);
    m->group = n > 1 ? avs_local_group_create(n, devs.data()) : nullptr;
    for (int r = 0; r < n; ++r)
        for (int q = 0; q < r; ++q)
            if (devs[q] == devs[r]) {
                g_avsAsyncAlloc = true;   // ranks share a device: no device-synchronising allocator calls from now on
                // ... and no lazily loaded kernels: the first launch of a kernel may synchronise the context, which cannot complete
                // while the peer rank's kernel on the SAME device spins waiting for this rank (tests/conftest.py sets the variable)
                const char *ml = getenv("CUDA_MODULE_LOADING");
                static bool warned = false;
                if (!warned && !(ml && strcmp(ml, "EAGER") == 0)) {
                    fprintf(stderr, "[avs] avs_create_multi: ranks share a GPU; set CUDA_MODULE_LOADING=EAGER before CUDA initialises "
                                    "(lazy kernel loading can stall a rank until its peer's spin-wait times out)\n");
                    warned = true;
                }
            }
    m->ctx.assign(n, nullptr);
    m->status.assign(n, AVS_OK);
    m->res.resize(n);
    // peer access is enabled per rank inside avs_dist_init; creation itself has no collective, so it can run sequentially
    for (int r = 0; r < n; ++r) {
        AvsDeviceConfig cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.size = sizeof(cfg);
        cfg.device = devs[r];
        cfg.rank = r;
        cfg.nranks = n;
        cfg.time_spmv = time_spmv;
        int share = 0;
        for (int q = 0; q < n; ++q) share += devs[q] == devs[r];
        int rc = createImpl(&cfg, m->group, share, &m->ctx[r]);
        if (rc != AVS_OK) {
            for (int q = 0; q < r; ++q) avs_destroy(m->ctx[q]);
            if (m->group) avs_local_group_destroy(m->group);
            delete m;
            return rc;
        }
        m->ctx[r]->slabOutputOnly = n > 1;
    }
    *out = m;
    return AVS_OK;
}

void avs_destroy_multi(AvsMulti *m) {
    if (!m) return;
    for (AvsContext *c : m->ctx) avs_destroy(c);
    if (m->group) avs_local_group_destroy(m->group);
    delete m;
}

int avs_multi_size(AvsMulti *m) { return m ? m->n : 0; }
AvsContext *avs_multi_context(AvsMulti *m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

int avs_solve_multi(AvsMulti *m, const AvsFields *in, const AvsParams *p, AvsVelocityOut *out, AvsResult *res) {
    if (!m || !in || !p || !res || res->size != sizeof(AvsResult)) return AVS_ERR_INVALID_ARGUMENT;
    if (m->n > 1) {
        // the ranks read the same host arrays: device pointers belong to one GPU and cannot be shared this way
        const AvsField *all[] = {&in->surface, &in->viscosity, &in->density, &in->collision, &in->vel[0], &in->vel[1], &in->vel[2],
                                 &in->face_weights[0], &in->face_weights[1], &in->face_weights[2], &in->collision_vel[0], &in->collision_vel[1],
                                 &in->collision_vel[2]};
        for (const AvsField *f : all)
            if (f->data && f->on_device) return AVS_ERR_UNSUPPORTED;
        if (out && out->on_device) return AVS_ERR_UNSUPPORTED;
        if (p->cancel) return AVS_ERR_UNSUPPORTED;   // a cancelled rank would leave its peers waiting
    }
    avs_local_group_reset(m->group);
    std::vector<std::thread> pool;
    std::vector<std::string> errors(m->n);
    for (int r = 0; r < m->n; ++r) {
        m->res[r] = *res;
        pool.emplace_back([m, r, in, p, out, &errors]() {
            m->status[r] = avs_solve(m->ctx[r], in, p, out, &m->res[r]);
            if (m->status[r] != AVS_OK) {
                errors[r] = std::string(avs_last_error()) + " " + m->ctx[r]->lastError;
                avs_local_group_fail(m->group);   // wake the peers out of their barriers
            }
        });
    }
    for (auto &t : pool) t.join();
    *res = m->res[0];
    // the rank that failed FIRST is the one to report: its peers only see their barriers give up afterwards
    int rc = AVS_OK, primary = -1;
    for (int r = 0; r < m->n; ++r)
        if (m->status[r] != AVS_OK && (primary < 0 || (errors[primary].find("in-process group") != std::string::npos &&
                                                        errors[r].find("in-process group") == std::string::npos)))
            primary = r;
    if (primary >= 0) {
        rc = m->status[primary];
        snprintf(g_lastError, sizeof(g_lastError), "rank %d: %s (%s)", primary, avs_status_string(rc), errors[primary].c_str());
    }
    // stage times: the slowest rank of every stage (the ranks run side by side)
    for (int r = 1; r < m->n; ++r)
        for (int i = 0; i < AVS_STAGE_COUNT; ++i) res->stage_ms[i] = std::max(res->stage_ms[i], m->res[r].stage_ms[i]);
    res->status = rc;
    return rc;
}

}  // extern "C"
