/*
 * avs.h -- C-ABI of the B200-native adaptive-octree viscosity solve.
 *
 * This is the drop-in boundary for ONE hot path of rgoldade/AdaptiveViscositySolver: everything
 * HDK_AdaptiveViscosity::solveGasSubclass does after it has fetched and validated its seven
 * Houdini fields (Source/HDK_AdaptiveViscosity.cpp:233-707), i.e. integration weights, refinement
 * mask, octree labels, DOF labelling, SPD assembly and the conjugate-gradient solve.  The host
 * keeps field lookup / validation / error reporting (HDK_AdaptiveViscosity.cpp:126-231) and calls
 * avs_solve() with flat copies of the fields; INTEGRATION.md shows the ~150-line shim.
 *
 * Conventions
 *   - plain C structs, first member `size` = sizeof(struct) for versioning; no C++/torch types;
 *   - flat arrays are x-fastest: idx = x + nx*(y + ny*z), float32 (Houdini stores fp32 voxels,
 *     HDK_Utilities.h:219-237);
 *   - a field component records the world position of its sample (0,0,0) (`org`) and its spacing,
 *     so centre / face / edge / unaligned collision grids are all described the same way;
 *   - pointers may be host or device pointers (`on_device`); the library never frees caller memory;
 *   - every entry point returns an AvsStatus; negative = error, never aborts, no exceptions cross;
 *   - a context is single-threaded (one solve at a time); separate contexts may run concurrently.
 */
#ifndef AVS_H
#define AVS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVS_ABI_VERSION 3

typedef enum AvsStatus {
    AVS_OK = 0,
    AVS_ERR_INVALID_ARGUMENT = -1,  /* null pointer, bad size field, bad resolution               */
    AVS_ERR_MISSING_FIELD = -2,     /* mirrors the addError paths of AV.cpp:152-229               */
    AVS_ERR_MISALIGNED_FIELD = -3,  /* face weights / viscosity / density not aligned (AV.cpp:169,210,225) */
    AVS_ERR_ALLOC = -4,
    AVS_ERR_CUDA = -5,
    AVS_ERR_NCCL = -6,
    AVS_ERR_CANCELLED = -7,         /* maps UT_Interrupt::opInterrupt (AV.cpp:911, ...)           */
    AVS_ERR_BREAKDOWN = -8,         /* CG produced a non-finite scalar (Eigen: solver.info()!=Success, AV.cpp:621) */
    AVS_ERR_NO_DEVICE = -9,
    AVS_ERR_UNSUPPORTED = -10
} AvsStatus;

typedef enum AvsPrecision {
    AVS_PRECISION_F64 = 0,          /* SolveType = fpreal64 (HDK_Utilities.h:32)                   */
    AVS_PRECISION_F32 = 1           /* USESINGLEPRECISION  (HDK_Utilities.h:25-30)                */
} AvsPrecision;

/* One scalar component of a Houdini field (SIM_RawField). data == NULL => constant field
 * (the reference's field()->isConstant() fast paths, AV.cpp:2090, 2248, 2501). */
typedef struct AvsField {
    const float *data;
    int32_t res[3];
    double org[3];      /* world position of sample (0,0,0) */
    double dx;          /* sample spacing (cubic voxels, AV.cpp:242) */
    float constant;
    int32_t on_device;  /* 1: `data` is a CUDA device pointer on the context's device */
} AvsField;

/* The seven inputs of solveGasSubclass (AV.cpp:138-231), flattened. */
typedef struct AvsFields {
    uint32_t size;
    int32_t res[3];           /* liquid surface resolution                                      */
    double origin[3];         /* world position of the grid corner                              */
    double dx;                /* voxel size                                                     */
    AvsField surface;         /* "surface": liquid SDF, centre sampled, negative inside (AV.cpp:138,195) */
    AvsField vel[3];          /* "vel": face sampled (AV.cpp:139,157)                           */
    AvsField face_weights[3]; /* "surfaceweights": aligned with vel (AV.cpp:144,169)            */
    AvsField viscosity;       /* "viscosity": aligned with surface, or constant (AV.cpp:203-216)*/
    AvsField density;         /* "massdensity": aligned with surface, or constant (AV.cpp:218-231) */
    AvsField collision;       /* "collision": positive inside the solid; any grid (AV.cpp:141)  */
    AvsField collision_vel[3];/* "collisionvel": any grid (AV.cpp:142)                          */
} AvsFields;

/* The DOP parameters read through GET_DATA_FUNC_* (HDK_AdaptiveViscosity.h:28-41); defaults are the
 * reference's *effective* defaults (fineBandwidth and doApplySolidWeights
 * read options no parameter defines). avs_default_params() fills them in. */
typedef struct AvsParams {
    uint32_t size;
    double dt;                       /* timestep (AV.cpp:130)                          */
    double tolerance;                /* 1e-3  (AV.cpp:62-63)                           */
    double extrapolation;            /* 0.5 voxels (AV.cpp:68-69, scaled by dx at :243)*/
    int32_t max_iterations;          /* 2500  (AV.cpp:65-66)                           */
    int32_t number_super_samples;    /* 3     (AV.cpp:104)                             */
    int32_t octree_levels;           /* 4     (AV.cpp:106)                             */
    int32_t fine_bandwidth;          /* 0 => band = max(2, 0) (AV.cpp:259)             */
    int32_t use_enhanced_gradients;  /* 1     (AV.cpp:109)                             */
    int32_t do_apply_solid_weights;  /* 0     (AV.h:37)                                */
    int32_t precision;               /* AvsPrecision                                   */
    int32_t check_every;             /* CG iterations between host convergence polls (0 = default) */
    volatile const int32_t *cancel;  /* optional; polled between CG batches (UT_Interrupt) */
} AvsParams;

/* Caller-allocated output velocity, same layout as AvsFields.vel (AV.cpp:696-706). Only faces the
 * reference would write (regular label >= 0 or SOLIDBOUNDARY, AV.cpp:2843-2890) differ from the INPUT
 * velocity (AvsFields.vel): solveGasSubclass updates `vel` in place (AV.cpp:698), so on return the arrays
 * hold the input velocity with the solved faces overwritten. Device callers (on_device = 1) pre-fill the
 * arrays themselves (typically out == in). */
typedef struct AvsVelocityOut {
    float *vel[3];
    int32_t on_device;
} AvsVelocityOut;

enum { AVS_STAGE_COUNT = 12 };
/* stage_ms[] slots, named after the reference's UT_PerfMonAutoSolveEvent labels (AV.cpp:306-880) */
enum AvsStage {
    AVS_STAGE_UPLOAD = 0,             /* host -> device copies of the fields                     */
    AVS_STAGE_SURFACE_WEIGHTS = 1,    /* "Compute Surface Weights" (+ collision weights)         */
    AVS_STAGE_OCTREE = 2,             /* "Build Mask for Octree" + "Build Octree"                */
    AVS_STAGE_REGULAR_LABELS = 3,     /* "Build Regular Grid Velocity Labels"                    */
    AVS_STAGE_OCTREE_LABELS = 4,      /* "Build Octree Velocity and Stress Labels"               */
    AVS_STAGE_RESTRICTION = 5,        /* "Interpolate Regular Grid Velocities at Octree Velocity Faces" */
    AVS_STAGE_SYSTEM = 6,             /* "Build Edge/Cell Stress Stencils" + "Build Octree Linear System" */
    AVS_STAGE_SOLVE = 7,              /* "Solve Linear System"                                   */
    AVS_STAGE_APPLY = 8,              /* "Apply Octree Solution to Regular Grid"                 */
    AVS_STAGE_DOWNLOAD = 9,           /* device -> host copy of the velocity                     */
    AVS_STAGE_TOTAL = 10
};

/* What the reference reports through setExtraInfo (AV.cpp:645-652) plus per-stage timings. */
typedef struct AvsResult {
    uint32_t size;
    int32_t status;
    int32_t iterations;       /* solver.iterations() (AV.cpp:629)        */
    int32_t levels;           /* octree levels actually built (OG.cpp:198-211) */
    double error;             /* solver.error()      (AV.cpp:630)        */
    int64_t octree_dofs;      /* octreeVelocityDOFCount (AV.cpp:395)     */
    int64_t regular_dofs;     /* regularVelocityDOFcount (AV.cpp:323)    */
    int64_t edge_dofs;        /* edgeStressDOFCount  (AV.cpp:403)        */
    int64_t center_dofs;      /* centerStressDOFCount (AV.cpp:408)       */
    int64_t nnz;              /* non-zeros of the assembled matrix       */
    int64_t local_rows;       /* rows owned by this rank                 */
    int64_t spmv_launches;    /* SpMV launches that did work (residual + CG iterations; launches after convergence are no-ops) */
    int64_t kernel_launches;  /* all kernels this call launched          */
    float stage_ms[AVS_STAGE_COUNT];
    float spmv_ms;            /* accumulated device time of the SpMV phases: persistent CG kernel -- measured inside the kernel
                                 (%globaltimer of CTA 0, grid barrier to grid barrier); per-launch CG (AVS_CG_MODE=launch) --
                                 CUDA events around every SpMV launch, only with AvsDeviceConfig.time_spmv */
    float cg_update_xr_ms;    /* same for the x,r update phase (incl. its scalar all-reduce wait when nranks > 1) */
    float cg_update_p_ms;     /* same for the p update phase (incl. the halo push when nranks > 1)              */
    int32_t dist_mode;        /* 0 single GPU, 1 NCCL hot loop, 2 peer-memory (NVLink loads/stores in our kernels) */
    int32_t reserved0;
    int64_t halo_columns;     /* off-rank columns this rank's rows reference */
    int64_t interpolated_faces; /* regular faces inside coarse cells, filled by the octree interpolator (interpSPGrid,
                                  HDK_OctreeVectorFieldInterpolator.cpp:660-845); 0 when depth == 1 */
    float cg_kernel_ms;       /* device time of the persistent CG kernel, CUDA events around its cooperative launch(es) */
    int32_t cg_kernel_launches; /* cooperative launches of that kernel (1 unless check_every / cancel chunking is on)     */
} AvsResult;

typedef struct AvsDeviceConfig {
    uint32_t size;
    int32_t device;           /* CUDA device ordinal                                        */
    int32_t rank;             /* this process' rank in the row partition (0 when nranks==1) */
    int32_t nranks;           /* number of ranks sharing one solve (1 = single GPU)         */
    const void *nccl_unique_id; /* 128-byte ncclUniqueId from avs_nccl_unique_id() on rank 0, broadcast by the host;
                                 NULL when nranks==1. The library creates its own communicator from it. */
    void *stream;             /* cudaStream_t to run on; NULL = the library creates its own */
    int32_t time_spmv;        /* 1: bracket every SpMV launch with CUDA events (bench only) */
    int32_t distributed_output; /* nranks > 1 only.  0 (default): on return EVERY rank's output arrays hold the whole velocity field
                                 (the z-slabs are all-gathered over NVLink).  1: a rank fills -- and, for host arrays, downloads --
                                 only the z-slab of the regular grid it computed; the other planes keep what the caller put there.
                                 The slabs of all ranks tile the grid (avs_get_output_slab). */
} AvsDeviceConfig;

typedef struct AvsContext AvsContext;

/* life cycle */
int avs_abi_version(void);
int avs_device_count(void);            /* CUDA devices visible to the process (0 = none: every avs_create fails) */
int avs_nccl_unique_id(void *out128);  /* rank 0: fill 128 bytes; the host broadcasts them to every rank */
int avs_create(const AvsDeviceConfig *cfg, AvsContext **out);
void avs_destroy(AvsContext *ctx);
const char *avs_status_string(int status);
const char *avs_last_error(void);   /* detail of the last AVS_ERR_CUDA / AVS_ERR_UNSUPPORTED on this thread */
void avs_default_params(AvsParams *p);

/* The drop-in call: replaces HDK_AdaptiveViscosity.cpp:233-707 in one shot.
 * Scenes the reference itself does not support -- a liquid that reaches the boundary of the grid makes its debug build assert
 * (HDK_AdaptiveViscosity.cpp:411-413, 881) and its release build hand Eigen a triplet with column -3 (:1886-1894) -- are refused
 * with AVS_ERR_UNSUPPORTED (text in avs_last_error()) when a stencil of the assembled system references a face that is not a
 * degree of freedom; nothing is written to `out`, the context stays usable.  Pad the fields so that the liquid stays a few voxels
 * inside the grid. */
int avs_solve(AvsContext *ctx, const AvsFields *in, const AvsParams *p, AvsVelocityOut *out, AvsResult *res);

/* ---- one process, several GPUs ------------------------------------------------------------------
 * The same drop-in call for a host that owns all GPUs from ONE thread -- the shape of the DOP: solveGasSubclass is called
 * on Houdini's cook thread, one object at a time (HDK_AdaptiveViscosity.cpp:126-128, HDK_AdaptiveViscosity.h:57-58).
 * avs_create_multi makes one rank context per entry of `devices` (an ordinal may repeat: ranks then share that GPU, which is
 * how the multi-rank path is tested on a one-GPU box -- that configuration needs CUDA_MODULE_LOADING=EAGER in the environment
 * before CUDA initialises, because the ranks' kernels wait for each other and a lazily loaded kernel cannot start meanwhile); avs_solve_multi row-partitions the solve over them and returns when
 * every rank is done.  Fields and the output velocity must be HOST arrays (every rank reads the same inputs and writes the
 * z-slab of `out` it computed); `cancel` is not supported.  No NCCL and no CUDA IPC are involved: the GPUs must be able to
 * access each other's memory (cudaDeviceCanAccessPeer).  `res` receives rank 0's result with per-stage times maximised over
 * the ranks; avs_multi_context(m, r) exposes rank r's context to the read-back functions below. */
typedef struct AvsMulti AvsMulti;
int avs_create_multi(const int32_t *devices, int32_t n, int32_t time_spmv, AvsMulti **out);
void avs_destroy_multi(AvsMulti *m);
int avs_multi_size(AvsMulti *m);
AvsContext *avs_multi_context(AvsMulti *m, int rank);
int avs_solve_multi(AvsMulti *m, const AvsFields *in, const AvsParams *p, AvsVelocityOut *out, AvsResult *res);

/* ---- staged entry points (tests, benchmarks, multi-step hosts) --------------------------------- */

/* Stages 1-9 only (weights ... linear system); the system stays resident in the context. */
int avs_assemble(AvsContext *ctx, const AvsFields *in, const AvsParams *p, AvsResult *res);
/* CG on the resident system (stage 10), initial guess = restricted u^n (AV.cpp:627). May be called
 * repeatedly; every call restarts from the initial guess. */
int avs_solve_resident(AvsContext *ctx, const AvsParams *p, AvsResult *res);
/* Stage 11 on the resident solution: node pyramid + interpSPGrid + write-back to the regular grid. */
int avs_apply(AvsContext *ctx, AvsVelocityOut *out, AvsResult *res);

/* The octree geometry dump of the DOP's `doPrintOctree` / `onlyPrintOctree` toggles (AV.cpp:283-294;
 * HDK_OctreeGrid::outputOctreeGeometry, HDK_OctreeGrid.cpp:245-308).
 * avs_build_octree runs stages 1-3 only (weights, refinement mask, label pyramid) -- the `onlyPrintOctree` early return.
 * avs_get_octree_points is valid after avs_build_octree / avs_assemble / avs_solve: one point per ACTIVE cell of every
 * built level -- pos[3*i..] = cell centre (fp32, like UT_Vector3), pscale[i] = voxel size of the cell's level,
 * level[i] = "octreeLevel".  HOST buffers; NULL arrays are skipped, so a first call with NULLs returns the count.
 * Point order is level-major then x-fastest; the reference's order follows its tile iterator and is not a contract. */
int avs_build_octree(AvsContext *ctx, const AvsFields *in, const AvsParams *p, AvsResult *res);
int avs_get_octree_points(AvsContext *ctx, int64_t *count, float *pos, float *pscale, int32_t *level);

/* Read-back of the resident state, all to HOST buffers (NULL pointers are skipped).
 * keys: n x 5 int32 (level, axis, i, j, k) -- DOF numbering is NOT part of the parity contract,
 * results are compared by geometric key. */
int avs_get_sizes(AvsContext *ctx, int64_t *n_rows, int64_t *nnz, int32_t *levels);
int avs_get_local_range(AvsContext *ctx, int64_t *row_begin, int64_t *row_end);  /* rows owned by this rank */
int avs_get_row_starts(AvsContext *ctx, int64_t *starts /* [nranks+1] */);  /* first row of every rank's block */
/* z-planes [z0, z1) of velocity component `axis` that this rank computes in stage 11 (the whole grid when nranks == 1) */
int avs_get_output_slab(AvsContext *ctx, int axis, int32_t *z0, int32_t *z1);
int avs_get_keys(AvsContext *ctx, int32_t *keys);
/* row_ptr/col/val/rhs cover the rows this rank owns (global column ids); x0 is the full restricted u^n [n_rows] */
int avs_get_system_csr(AvsContext *ctx, int64_t *row_ptr, int32_t *col, double *val, double *rhs, double *x0);
int avs_get_solution(AvsContext *ctx, double *x);   /* the rows this rank owns: [N*rank/nranks, N*(rank+1)/nranks) */
/* One grid of the resident state; valid after avs_build_octree (kind 0) / avs_assemble / avs_solve, else AVS_ERR_INVALID_ARGUMENT.
 *   kind 0  cell labels of `level`                 uint8   INACTIVE 0 / ACTIVE 1 / UP 2 / DOWN 3
 *   kind 1  octree face labels of (level, axis)    int32   >= 0 DOF row, -1 UNASSIGNED, -2 SOLIDBOUNDARY, -3 OUTSIDE
 *   kind 2  edge-stress labels of (level, axis)    int8    0 active, -1 / -2 / -3 as above
 *   kind 3  centre-stress labels of `level`        int8    0 active, -1 otherwise
 *   kind 4  regular-grid face labels of `axis`     int8    0 solved face, -1 / -2 / -3 as above (level ignored)
 *   kind 5  centre integration weights             float32 (level, axis ignored)
 *   kind 6  edge integration weights of `axis`     float32 (level ignored)
 * Writes the grid resolution to res[3] (x-fastest); copies to the HOST buffer `out` when non-NULL; returns bytes via *nbytes. */
int avs_get_grid(AvsContext *ctx, int kind, int level, int axis, void *out, int32_t *res, int64_t *nbytes);

/* Stand-alone linear algebra on a caller-supplied CSR matrix (host pointers, int64 row_ptr, int32 col):
 * the CG hot loop by itself -- what the reference hands to Eigen::ConjugateGradient (AV.cpp:611-630). */
int avs_cg_csr(AvsContext *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
               const double *rhs, double *x /* in: guess, out: solution */, const AvsParams *p, AvsResult *res);
int avs_spmv_csr(AvsContext *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                 const double *x, double *y, int precision, int repeats, float *ms_per_launch);

/* Benchmark support: time `repeats` launches of the SpMV kernel on the resident system with CUDA events
 * on the library stream; returns average ms per launch and the algorithmic bytes of one launch
 * (nnz*(s+4) + (N+1)*4 + 2*N*s). */
int avs_time_spmv_resident(AvsContext *ctx, int precision, int repeats, float *ms_per_launch, double *algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* AVS_H */
