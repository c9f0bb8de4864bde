// =============================================================================
// ref_harness.cpp -- C entry points around the REFERENCE's own solver class,
// compiled unchanged from /root/reference/Source against oracle/mock_hdk
// (see mock_hdk.h for what is real and what is assumed).  TEST INFRASTRUCTURE:
// built by oracle/Makefile into oracle/_ref/libavs_ref.so; only tests/ and
// bench.py's CPU arm load it.
//
// ref_create() flattens a scene (same structs as oracle/avs_oracle.cpp's OrcScene /
// OrcParams) into the mock's SIM fields, ref_run() calls
// HDK_AdaptiveViscosity::solveGasSubclass (HDK_AdaptiveViscosity.cpp:126) and the
// getters hand back what the reference computed: octree labels, weights, the index
// grids, the assembled matrix / rhs / initial guess (seen through the Eigen stand-in
// when the reference calls solveWithGuess, HDK_AdaptiveViscosity.cpp:627), the
// solution, iteration count and error, the regular-grid velocity it leaves in `vel`,
// and the octree geometry dump.
// =============================================================================
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

// With -DAVS_SHIM_HARNESS the same harness drives integration/hdk/HDK_AdaptiveViscosityB200.cpp instead -- the Houdini-side shim
// of THIS repository (same class name, same DOP surface), which hands the solve to libavs_b200.so: the drop-in, run end to end
// on the stand-in fields (oracle/_ref/libavs_shim.so; needs a GPU at run time).
#ifdef AVS_SHIM_HARNESS
#include "HDK_AdaptiveViscosityB200.h"
#else
#include "HDK_AdaptiveViscosity.h"
#endif

namespace {

struct Grid {
    int n[3] = {0, 0, 0};
    std::vector<float> f;
    std::vector<int64_t> i;
};

struct RefRun {
    // inputs (owned by the mock fields)
    SIM_ScalarField surface, viscosity, density, collision;
    SIM_VectorField vel, faceWeights, collisionVel;
    SIM_Object obj;
    std::map<std::string, double> options;
    double dt = 1.0 / 24.0;
    // outputs
    bool ok = false;
    int levels = 0;
    Grid centerW, edgeW[3];
    std::vector<Grid> labels;                 // per level (float labels as stored by the reference)
    Grid regIdx[3];
    std::vector<Grid> faceIdx[3], edgeIdx[3], centerIdx;
    int64_t nFace = 0, nEdge = 0, nCenter = 0, regularDofs = 0;
    std::vector<int64_t> rowPtr;
    std::vector<int32_t> colIdx;
    std::vector<double> val, rhs, x0, sol;
    int iterations = 0;
    double error = 0;
    std::string extraInfo;
    std::vector<int32_t> faceKey;             // nFace x 5
    Grid outVel[3];
    std::vector<float> geoPos, geoScale;
    std::vector<int32_t> geoLevel;
    unsigned long long seq0 = 0;
    bool snapped = false;
};

template <class F>
void copyGrid(const F &field, Grid &g, bool isIndex) {
    g.n[0] = field.getXRes(); g.n[1] = field.getYRes(); g.n[2] = field.getZRes();
    const size_t cnt = (size_t)g.n[0] * g.n[1] * g.n[2];
    if (isIndex) g.i.resize(cnt); else g.f.resize(cnt);
    size_t k = 0;
    for (int z = 0; z < g.n[2]; ++z)
        for (int y = 0; y < g.n[1]; ++y)
            for (int x = 0; x < g.n[0]; ++x, ++k) {
                if (isIndex) g.i[k] = (int64_t)field.field()->getValue(x, y, z);
                else g.f[k] = (float)field.field()->getValue(x, y, z);
            }
}

struct OrcField {
    const float *data;
    int res[3];
    double org[3];
    double dx;
    float constant;
};
struct OrcScene {
    int res[3];
    double origin[3];
    double dx;
    OrcField surface, vel[3], faceWeights[3], viscosity, density, collision, collisionVel[3];
};
struct OrcParams {
    double dt, tolerance, extrapolation;
    int maxIterations, numberSuperSamples, octreeLevels, fineBandwidth;
    int useEnhancedGradients, doApplySolidWeights, singlePrecision;
    int stopAfterStage;
};

// A flat caller field -> SIM_RawField.  `cellRes`/`orig`/`dx` describe the grid the field lives on; `sample` its sample type.
void loadField(SIM_RawField &dst, const OrcField &f, SIM_FieldSample sample, const int cellRes[3], const double gridOrig[3], double dx) {
    UT_Vector3 orig, size;
    for (int a = 0; a < 3; ++a) {
        orig.setExact(a, gridOrig[a]);
        size.setExact(a, (double)(float)dx * cellRes[a]);
    }
    dst.init(sample, orig, size, cellRes[0], cellRes[1], cellRes[2]);
    if (!f.data) {   // constant field: every tile constant (field()->isConstant() fast paths, AV.cpp:2090, 2248, 2501)
        dst.makeConstant(f.constant);
        return;
    }
    UT_VoxelArray<fpreal32> &v = *dst.fieldNC();
    size_t k = 0;
    for (int z = 0; z < f.res[2]; ++z)
        for (int y = 0; y < f.res[1]; ++y)
            for (int x = 0; x < f.res[0]; ++x, ++k) v.setValue(x, y, z, f.data[k]);
    v.collapseAllTiles();   // Houdini keeps constant regions of a field as constant tiles (AV.cpp:829)
    dst.myFlatSampling = true;
    for (int a = 0; a < 3; ++a) dst.myFlatOrg[a] = f.org[a];
    dst.myFlatDx = f.dx;
}

// a field on its own grid (collision, collisionvel): centre-sampled grid whose sample (0,0,0) sits at f.org
void loadFreeField(SIM_RawField &dst, const OrcField &f) {
    int res[3] = {f.res[0], f.res[1], f.res[2]};
    double orig[3];
    for (int a = 0; a < 3; ++a) orig[a] = f.org[a] - 0.5 * f.dx;
    if (!f.data) { res[0] = res[1] = res[2] = 1; }
    loadField(dst, f, SIM_SAMPLE_CENTER, res, orig, f.data ? f.dx : 1.0);
}

}  // namespace

extern "C" {

void *ref_create(const OrcScene *s, const OrcParams *p) {
    RefRun *r = new RefRun();
    const SIM_FieldSample faceSample[3] = {SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ};
    loadField(r->surface.myField, s->surface, SIM_SAMPLE_CENTER, s->res, s->origin, s->dx);
    loadField(r->viscosity.myField, s->viscosity, SIM_SAMPLE_CENTER, s->res, s->origin, s->dx);
    loadField(r->density.myField, s->density, SIM_SAMPLE_CENTER, s->res, s->origin, s->dx);
    for (int a = 0; a < 3; ++a) {
        loadField(r->vel.myFields[a], s->vel[a], faceSample[a], s->res, s->origin, s->dx);
        loadField(r->faceWeights.myFields[a], s->faceWeights[a], faceSample[a], s->res, s->origin, s->dx);
        loadFreeField(r->collisionVel.myFields[a], s->collisionVel[a]);
    }
    loadFreeField(r->collision.myField, s->collision);
    r->obj.scalars[GAS_NAME_SURFACE] = &r->surface;
    r->obj.scalars["viscosity"] = &r->viscosity;
    r->obj.scalars[GAS_NAME_DENSITY] = &r->density;
    r->obj.scalars[GAS_NAME_COLLISION] = &r->collision;
    r->obj.vectors[GAS_NAME_VELOCITY] = &r->vel;
    r->obj.vectors["faceWeights"] = &r->faceWeights;
    r->obj.vectors[GAS_NAME_COLLISIONVELOCITY] = &r->collisionVel;
    r->options[SIM_NAME_TOLERANCE] = p->tolerance;
    r->options["maxIterations"] = p->maxIterations;
    r->options["numberSuperSamples"] = p->numberSuperSamples;
    r->options["octreeLevels"] = p->octreeLevels;
    r->options["fineBandwidth"] = p->fineBandwidth;
    r->options["useEnhancedGradients"] = p->useEnhancedGradients;
    r->options["doApplySolidWeights"] = p->doApplySolidWeights;
    r->options["extrapolation"] = p->extrapolation;
    r->options["doPrintOctree"] = 1;
    r->options["cudaDevice"] = 0;
    r->options["cudaDeviceCount"] = p->singlePrecision >> 8;        // shim harness: bits 8.. of the precision word = GPU count
    r->options["singlePrecision"] = p->singlePrecision & 1;
    r->options["onlyPrintOctree"] = (p->stopAfterStage > 0 && p->stopAfterStage <= 3) ? 1 : 0;
    r->dt = p->dt;
    return r;
}

// Validation paths of solveGasSubclass (AV.cpp:152-229): break the object between ref_create and ref_run.
//   op 0: remove the field `name` from the object;  op 1: shift the field's grid by one voxel along x (no longer aligned);
//   op 2: give component 0 of the vector field `name` centre sampling (no longer a staggered grid)
int ref_tamper(void *h, int op, const char *name) {
    RefRun *r = (RefRun *)h;
    auto s = r->obj.scalars.find(name);
    auto v = r->obj.vectors.find(name);
    if (s == r->obj.scalars.end() && v == r->obj.vectors.end()) return -1;
    if (op == 0) {
        if (s != r->obj.scalars.end()) r->obj.scalars.erase(s);
        else r->obj.vectors.erase(v);
    } else if (op == 1) {
        if (s != r->obj.scalars.end()) s->second->myField.myOrigD[0] += s->second->myField.myVoxelSizeD[0];
        else for (int a = 0; a < 3; ++a) v->second->myFields[a].myOrigD[0] += v->second->myFields[a].myVoxelSizeD[0];
    } else if (op == 2) {
        if (v == r->obj.vectors.end()) return -1;
        v->second->myFields[0].mySample = SIM_SAMPLE_CENTER;
    } else return -1;
    return 0;
}

void ref_destroy(void *h) { delete (RefRun *)h; }
void ref_set_threads(int n) { mock_hdk::setThreads(n); }
void ref_set_weight_shortcut(int on) { mock_hdk::weightShortcutRef() = on != 0; }

// Snapshot of the reference's locals (all registered fields that were initialised after the solve started).
static void snapshotFields(RefRun *r) {
    if (r->snapped) return;
    std::vector<const SIM_RawField *> floats;
    std::vector<const SIM_RawIndexField *> indices;
    for (const auto &e : mock_hdk::registry().live) {
        if (e.seq < r->seq0) continue;
        if (e.isIndex) indices.push_back((const SIM_RawIndexField *)e.field);
        else floats.push_back((const SIM_RawField *)e.field);
    }
    // init() order inside solveGasSubclass: centre weights, 3 edge weights (AV.cpp:761-765), octree labels per level
    // (OG.cpp:58-69); regular indices x3 (AV.cpp:314-321); per level: 3 face, 3 edge, 1 centre index grids (AV.cpp:370-392)
    if (floats.size() < 5 || indices.size() < 3 || (indices.size() - 3) % 7 != 0) return;
    copyGrid(*floats[0], r->centerW, false);
    for (int a = 0; a < 3; ++a) copyGrid(*floats[1 + a], r->edgeW[a], false);
    r->levels = (int)((indices.size() - 3) / 7);
    r->labels.resize(r->levels);
    for (int l = 0; l < r->levels; ++l) copyGrid(*floats[4 + l], r->labels[l], false);
    for (int a = 0; a < 3; ++a) {
        copyGrid(*indices[a], r->regIdx[a], true);
        r->faceIdx[a].resize(r->levels);
        r->edgeIdx[a].resize(r->levels);
    }
    r->centerIdx.resize(r->levels);
    for (int l = 0; l < r->levels; ++l) {
        for (int a = 0; a < 3; ++a) {
            copyGrid(*indices[3 + 7 * l + a], r->faceIdx[a][l], true);
            copyGrid(*indices[3 + 7 * l + 3 + a], r->edgeIdx[a][l], true);
        }
        copyGrid(*indices[3 + 7 * l + 6], r->centerIdx[l], true);
    }
    auto countDofs = [](const Grid &g) { int64_t n = 0; for (int64_t v : g.i) n += v >= 0; return n; };
    r->nFace = r->nEdge = r->nCenter = r->regularDofs = 0;
    for (int a = 0; a < 3; ++a) r->regularDofs += countDofs(r->regIdx[a]);
    for (int l = 0; l < r->levels; ++l) {
        for (int a = 0; a < 3; ++a) { r->nFace += countDofs(r->faceIdx[a][l]); r->nEdge += countDofs(r->edgeIdx[a][l]); }
        r->nCenter += countDofs(r->centerIdx[l]);
    }
    r->faceKey.assign((size_t)r->nFace * 5, -1);
    for (int l = 0; l < r->levels; ++l)
        for (int a = 0; a < 3; ++a) {
            const Grid &g = r->faceIdx[a][l];
            size_t k = 0;
            for (int z = 0; z < g.n[2]; ++z)
                for (int y = 0; y < g.n[1]; ++y)
                    for (int x = 0; x < g.n[0]; ++x, ++k) {
                        const int64_t idx = g.i[k];
                        if (idx < 0) continue;
                        int32_t *key = &r->faceKey[(size_t)idx * 5];
                        key[0] = l; key[1] = a; key[2] = x; key[3] = y; key[4] = z;
                    }
        }
    r->snapped = true;
}

int ref_run(void *h) {
    RefRun *r = (RefRun *)h;
    r->seq0 = mock_hdk::registry().nextSeq;
    r->snapped = false;
    mock_hdk::hooks().scopeEnd = [r](const char *label) {
        // the label scope closes when all index grids are numbered (AV.cpp:358-421)
        if (std::strcmp(label, "Build Octree Velocity and Stress Labels") == 0) snapshotFields(r);
    };
    mock_hdk::hooks().extraInfo = [r](const char *info) { r->extraInfo = info; };
#ifndef AVS_SHIM_HARNESS
    mock_eigen::hooks().beforeSolve = [r](const Eigen::SparseMatrix<double> &A, const Eigen::VectorXd &b, const Eigen::VectorXd &g) {
        r->rowPtr = A.ptr; r->colIdx = A.col; r->val = A.val;
        r->rhs.assign(b.data(), b.data() + b.size());
        r->x0.assign(g.data(), g.data() + g.size());
    };
    mock_eigen::hooks().afterSolve = [r](const Eigen::VectorXd &x, int it, double err) {
        r->sol.assign(x.data(), x.data() + x.size());
        r->iterations = it;
        r->error = err;
    };
#endif
    HDK_AdaptiveViscosity *solver = HDK_AdaptiveViscosity::mockCreate();
    solver->myOptions = r->options;
    GAS_SubSolver *base = solver;
    SIM_Engine engine;
    r->ok = base->solveGasSubclass(engine, &r->obj, 0.0, r->dt);
    delete base;   // through the (public, virtual) base destructor
    mock_hdk::hooks() = mock_hdk::Hooks();
#ifndef AVS_SHIM_HARNESS
    mock_eigen::hooks() = mock_eigen::Hooks();
#endif
    for (int a = 0; a < 3; ++a) copyGrid(r->vel.myFields[a], r->outVel[a], false);
    auto it = r->obj.geometry.find("octreeGeometry");
    if (it != r->obj.geometry.end()) {
        const GU_Detail &g = it->second->myGdp;
        const size_t n = g.myPos.size();
        r->geoPos.resize(n * 3); r->geoScale.resize(n); r->geoLevel.resize(n);
        const auto &ps = g.myFloat.at("pscale");
        const auto &lv = g.myInt.at("octreeLevel");
        for (size_t i = 0; i < n; ++i) {
            for (int a = 0; a < 3; ++a) r->geoPos[3 * i + a] = (float)(double)g.myPos[i][a];
            r->geoScale[i] = (float)ps[i];
            r->geoLevel[i] = (int32_t)lv[i];
        }
    }
    return r->ok ? 0 : 1;
}

int ref_error_count(void *h) { return (int)((RefRun *)h)->obj.errors.size(); }
const char *ref_error_text(void *h, int i) { return ((RefRun *)h)->obj.errors[(size_t)i].c_str(); }
const char *ref_extra_info(void *h) { return ((RefRun *)h)->extraInfo.c_str(); }
int ref_levels(void *h) { return ((RefRun *)h)->levels; }
int64_t ref_count(void *h, int what) {
    RefRun *r = (RefRun *)h;
    switch (what) {
        case 0: return r->nFace;
        case 1: return r->nEdge;
        case 2: return r->nCenter;
        case 3: return r->regularDofs;
        case 4: return (int64_t)r->val.size();
        case 5: return r->iterations;
    }
    return -1;
}
double ref_error(void *h) { return ((RefRun *)h)->error; }
static int64_t giveFloat(const Grid &g, float *out, int *res) {
    if (res) for (int i = 0; i < 3; ++i) res[i] = g.n[i];
    if (out) std::memcpy(out, g.f.data(), g.f.size() * sizeof(float));
    return (int64_t)g.f.size();
}
static int64_t giveIndex(const Grid &g, int64_t *out, int *res) {
    if (res) for (int i = 0; i < 3; ++i) res[i] = g.n[i];
    if (out) std::memcpy(out, g.i.data(), g.i.size() * sizeof(int64_t));
    return (int64_t)g.i.size();
}
// kind: 0 centre weights, 1..3 edge weights
int64_t ref_get_float(void *h, int kind, float *out, int *res) {
    RefRun *r = (RefRun *)h;
    return giveFloat(kind == 0 ? r->centerW : r->edgeW[kind - 1], out, res);
}
int64_t ref_get_labels(void *h, int level, uint8_t *out, int *res) {
    RefRun *r = (RefRun *)h;
    const Grid &g = r->labels[(size_t)level];
    if (res) for (int i = 0; i < 3; ++i) res[i] = g.n[i];
    if (out) for (size_t i = 0; i < g.f.size(); ++i) out[i] = (uint8_t)g.f[i];
    return (int64_t)g.f.size();
}
// kind: 0 face, 1 edge, 2 centre, 3 regular face (level ignored)
int64_t ref_get_index_grid(void *h, int kind, int level, int axis, int64_t *out, int *res) {
    RefRun *r = (RefRun *)h;
    const Grid &g = kind == 0 ? r->faceIdx[axis][(size_t)level] : kind == 1 ? r->edgeIdx[axis][(size_t)level] : kind == 2 ? r->centerIdx[(size_t)level] : r->regIdx[axis];
    return giveIndex(g, out, res);
}
void ref_get_face_keys(void *h, int32_t *out) {
    RefRun *r = (RefRun *)h;
    std::memcpy(out, r->faceKey.data(), r->faceKey.size() * sizeof(int32_t));
}
// what: 0 x0 (restricted u^n = the CG's initial guess), 1 rhs, 2 solution
void ref_get_vector(void *h, int what, double *out) {
    RefRun *r = (RefRun *)h;
    const std::vector<double> &v = what == 0 ? r->x0 : what == 1 ? r->rhs : r->sol;
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}
int64_t ref_get_out_velocity(void *h, int axis, float *out, int *res) { return giveFloat(((RefRun *)h)->outVel[axis], out, res); }
void ref_get_csr(void *h, int64_t *rowPtr, int32_t *col, double *val) {
    RefRun *r = (RefRun *)h;
    std::memcpy(rowPtr, r->rowPtr.data(), r->rowPtr.size() * sizeof(int64_t));
    std::memcpy(col, r->colIdx.data(), r->colIdx.size() * sizeof(int32_t));
    std::memcpy(val, r->val.data(), r->val.size() * sizeof(double));
}
int64_t ref_get_octree_points(void *h, float *pos, float *pscale, int32_t *level) {
    RefRun *r = (RefRun *)h;
    const size_t n = r->geoScale.size();
    if (pos) {
        std::memcpy(pos, r->geoPos.data(), n * 3 * sizeof(float));
        std::memcpy(pscale, r->geoScale.data(), n * sizeof(float));
        std::memcpy(level, r->geoLevel.data(), n * sizeof(int32_t));
    }
    return (int64_t)n;
}

}  // extern "C"
