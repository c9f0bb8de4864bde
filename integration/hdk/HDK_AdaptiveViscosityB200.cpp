// HDK_AdaptiveViscosityB200.cpp -- host side of the drop-in: DOP registration, parameter surface, field lookup and
// validation stay on the Houdini side (what HDK_AdaptiveViscosity.cpp:20-231 does in the reference); everything the
// reference does after that (HDK_AdaptiveViscosity.cpp:233-707) is one avs_solve() call.
//
// Built only against a Houdini toolkit ($HFS); see CMakeLists.txt in this directory and INTEGRATION.md.
#include "HDK_AdaptiveViscosityB200.h"

#include <GU/GU_Detail.h>
#include <PRM/PRM_Include.h>
#include <SIM/SIM_DopDescription.h>
#include <SIM/SIM_FieldSampler.h>
#include <SIM/SIM_GeometryCopy.h>
#include <SIM/SIM_Object.h>
#include <SIM/SIM_ScalarField.h>
#include <SIM/SIM_VectorField.h>
#include <UT/UT_DSOVersion.h>
#include <UT/UT_Interrupt.h>
#include <UT/UT_PerfMonAutoEvent.h>
#include <UT/UT_WorkBuffer.h>

#include <vector>

#include "avs.h"

// DSO entry point, same symbol as the reference (HDK_AdaptiveViscosity.cpp:20-24)
void initializeSIM(void *) { IMPLEMENT_DATAFACTORY(HDK_AdaptiveViscosity); }

HDK_AdaptiveViscosity::HDK_AdaptiveViscosity(const SIM_DataFactory *factory) : BaseClass(factory) {}
HDK_AdaptiveViscosity::~HDK_AdaptiveViscosity()
{
    if (myMulti) avs_destroy_multi(myMulti);      // owns its rank contexts
    else if (myContext) avs_destroy(myContext);
}

// The parameter surface of the reference (HDK_AdaptiveViscosity.cpp:36-116): names, labels and defaults are kept so existing
// scenes bind; "cudaDevice" and "singlePrecision" are new (the reference selects precision at compile time, USESINGLEPRECISION).
const SIM_DopDescription *HDK_AdaptiveViscosity::getDopDescription()
{
    static PRM_Name surfaceName(GAS_NAME_SURFACE, "Liquid Surface Field");
    static PRM_Default surfaceDefault(0, "surface");
    static PRM_Name faceWeightsName("faceWeights", "Face Weights Field");
    static PRM_Default faceWeightsDefault(0, "surfaceweights");
    static PRM_Name velocityName(GAS_NAME_VELOCITY, "Liquid Velocity Field");
    static PRM_Default velocityDefault(0, "vel");
    static PRM_Name viscosityName("viscosity", "Viscosity Field");
    static PRM_Default viscosityDefault(0, "viscosity");
    static PRM_Name densityName(GAS_NAME_DENSITY, "Density Field");
    static PRM_Default densityDefault(0, "massdensity");
    static PRM_Name solidSurfaceName(GAS_NAME_COLLISION, "Solid Surface Field");
    static PRM_Default solidSurfaceDefault(0, "collision");
    static PRM_Name solidVelocityName(GAS_NAME_COLLISIONVELOCITY, "Solid Velocity Field");
    static PRM_Default solidVelocityDefault(0, "collisionvel");
    static PRM_Name applySolidWeightsName("applySolidWeights", "Apply Solid Weights");
    static PRM_Name toleranceName(SIM_NAME_TOLERANCE, "Relative Solver Tolerance");
    static PRM_Default toleranceDefault(1e-3);
    static PRM_Name maxIterationsName("maxIterations", "Max Solver Iterations");
    static PRM_Default maxIterationsDefault(2500);
    static PRM_Name extrapolationName("extrapolation", "Extrapolation");
    static PRM_Default extrapolationDefault(0.5);
    static PRM_Name superSamplesName("numberSuperSamples", "Samples Per Axis");
    static PRM_Default superSamplesDefault(3);
    static PRM_Name octreeLevelsName("octreeLevels", "Octree Levels");
    static PRM_Default octreeLevelsDefault(4);
    static PRM_Name fineBandwidthName("fineLayerBandwidth", "Fine Layer Bandwidth");
    static PRM_Default fineBandwidthDefault(2);
    static PRM_Name enhancedGradientsName("useEnhancedGradients", "Use Enhanced Gradients");
    static PRM_Name printOctreeName("doPrintOctree", "Output Octree Geometry");
    static PRM_Name onlyPrintOctreeName("onlyPrintOctree", "Only Output Octree");
    static PRM_Name octreeGeometryName("octreeGeometry", "Octree Geometry");
    static PRM_Default octreeGeometryDefault(0, "OctreeGeometry");
    static PRM_Name cudaDeviceName("cudaDevice", "CUDA Device");
    static PRM_Name cudaDeviceCountName("cudaDeviceCount", "CUDA Device Count");
    static PRM_Name singlePrecisionName("singlePrecision", "Single Precision Solve");

    static PRM_Template templates[] = {
        PRM_Template(PRM_STRING, 1, &surfaceName, &surfaceDefault),
        PRM_Template(PRM_STRING, 1, &faceWeightsName, &faceWeightsDefault),
        PRM_Template(PRM_STRING, 1, &velocityName, &velocityDefault),
        PRM_Template(PRM_STRING, 1, &viscosityName, &viscosityDefault),
        PRM_Template(PRM_STRING, 1, &densityName, &densityDefault),
        PRM_Template(PRM_STRING, 1, &solidSurfaceName, &solidSurfaceDefault),
        PRM_Template(PRM_STRING, 1, &solidVelocityName, &solidVelocityDefault),
        PRM_Template(PRM_TOGGLE, 1, &applySolidWeightsName, PRMzeroDefaults),
        PRM_Template(PRM_FLT, 1, &toleranceName, &toleranceDefault),
        PRM_Template(PRM_FLT, 1, &maxIterationsName, &maxIterationsDefault),   // PRM_FLT as in the reference (AV.cpp:100)
        PRM_Template(PRM_FLT, 1, &extrapolationName, &extrapolationDefault),
        PRM_Template(PRM_INT, 1, &superSamplesName, &superSamplesDefault),
        PRM_Template(PRM_INT, 1, &octreeLevelsName, &octreeLevelsDefault),
        PRM_Template(PRM_INT, 1, &fineBandwidthName, &fineBandwidthDefault),
        PRM_Template(PRM_TOGGLE, 1, &enhancedGradientsName, PRMoneDefaults),
        PRM_Template(PRM_TOGGLE, 1, &printOctreeName, PRMzeroDefaults),
        PRM_Template(PRM_TOGGLE, 1, &onlyPrintOctreeName, PRMzeroDefaults),
        PRM_Template(PRM_STRING, 1, &octreeGeometryName, &octreeGeometryDefault),
        PRM_Template(PRM_INT, 1, &cudaDeviceName, PRMzeroDefaults),
        PRM_Template(PRM_INT, 1, &cudaDeviceCountName, PRMoneDefaults),
        PRM_Template(PRM_TOGGLE, 1, &singlePrecisionName, PRMzeroDefaults),
        PRM_Template()};

    static SIM_DopDescription description(true, "HDK_AdaptiveViscosity", "HDK Adaptive Viscosity", "$OS", classname(), templates);
    setGasDescription(description);
    return &description;
}

AvsContext *HDK_AdaptiveViscosity::context(SIM_Object *obj)
{
    const int device = getCudaDevice();
    const int count = getCudaDeviceCount() > 1 ? getCudaDeviceCount() : 1;
    if (myContext && myContextDevice == device && myContextCount == count) return myContext;
    if (myMulti) { avs_destroy_multi(myMulti); myMulti = nullptr; myContext = nullptr; }
    if (myContext) { avs_destroy(myContext); myContext = nullptr; }
    int rc;
    if (count > 1) {
        // the cook thread owns all GPUs: one rank context per device, no NCCL, no second process (include/avs.h)
        std::vector<int32_t> devices(count);
        const int visible = avs_device_count();   // ordinals wrap when fewer GPUs are visible: ranks then share a GPU (tests)
        for (int i = 0; i < count; ++i) devices[i] = visible > 0 ? (device + i) % visible : device + i;
        rc = avs_create_multi(devices.data(), count, 0, &myMulti);
        if (rc == AVS_OK) myContext = avs_multi_context(myMulti, 0);   // read-back (octree geometry) goes through rank 0
    } else {
        AvsDeviceConfig cfg{};
        cfg.size = sizeof cfg;
        cfg.device = device;
        cfg.nranks = 1;
        rc = avs_create(&cfg, &myContext);
    }
    if (rc != AVS_OK) {
        addError(obj, SIM_MESSAGE, avs_last_error(), UT_ERROR_ABORT);   // no GPU: fail loudly, there is no CPU path
        myContext = nullptr;
        myMulti = nullptr;
        return nullptr;
    }
    myContextDevice = device;
    myContextCount = count;
    return myContext;
}

namespace {

// SIM_RawField -> flat x-fastest float32 buffer + sample descriptor (include/avs.h: AvsField)
AvsField flatten(const SIM_RawField &f, std::vector<float> &store, bool allowConstant = true)
{
    AvsField d{};
    fpreal32 c = 0;
    if (allowConstant && f.field()->isConstant(&c)) {   // the reference's isConstant fast paths (AV.cpp:2090, 2248, 2501)
        d.data = nullptr;
        d.constant = c;
        return d;
    }
    const int nx = f.field()->getXRes(), ny = f.field()->getYRes(), nz = f.field()->getZRes();   // samples, not cells
    store.resize(size_t(nx) * ny * nz);
    f.field()->flatten(store.data(), 1, nx, exint(nx) * ny);
    UT_Vector3 p0;
    f.indexToPos(0, 0, 0, p0);                   // world position of sample (0,0,0): covers centre / face / edge / corner sampling
    d.data = store.data();
    d.res[0] = nx; d.res[1] = ny; d.res[2] = nz;
    d.org[0] = p0.x(); d.org[1] = p0.y(); d.org[2] = p0.z();
    d.dx = f.getVoxelSize().maxComponent();
    d.on_device = 0;
    return d;
}

}  // namespace

bool HDK_AdaptiveViscosity::solveGasSubclass(SIM_Engine &engine, SIM_Object *obj, SIM_Time time, SIM_Time timestep)
{
    // ---- field lookup and validation: the contract of HDK_AdaptiveViscosity.cpp:138-231, same messages ----------
    const SIM_ScalarField *surfaceField = getConstScalarField(obj, GAS_NAME_SURFACE);
    SIM_VectorField *velocityField = getVectorField(obj, GAS_NAME_VELOCITY);
    const SIM_ScalarField *solidField = getConstScalarField(obj, GAS_NAME_COLLISION);
    const SIM_VectorField *solidVelocityField = getConstVectorField(obj, GAS_NAME_COLLISIONVELOCITY);
    const SIM_VectorField *faceWeightsField = getConstVectorField(obj, "faceWeights");
    const SIM_ScalarField *viscosityField = getConstScalarField(obj, "viscosity");
    const SIM_ScalarField *densityField = getConstScalarField(obj, GAS_NAME_DENSITY);

    auto fail = [&](const char *msg) { addError(obj, SIM_MESSAGE, msg, UT_ERROR_WARNING); return false; };
    if (!velocityField) return fail("Liquid velocity field missing");
    if (!velocityField->isFaceSampled()) return fail("Liquid velocity field must be a staggered grid");
    if (!faceWeightsField) return fail("Face weights field missing");
    if (!faceWeightsField->isAligned(velocityField)) return fail("Face weights must align with velocity samples");
    if (!solidField) return fail("Solid surface field missing");
    if (!solidVelocityField) return fail("Solid velocity field missing");
    if (!surfaceField) return fail("Liquid surface field is missing");
    if (!viscosityField) return fail("Viscosity field is missing");
    if (!viscosityField->getField()->isAligned(surfaceField->getField())) return fail("Viscosity field must align with the surface volume");
    if (!densityField) return fail("Density field is missing");
    if (!densityField->getField()->isAligned(surfaceField->getField())) return fail("Density field must align with the surface volume");

    AvsContext *ctx = context(obj);
    if (!ctx) return false;

    // ---- flatten the seven fields ---------------------------------------------------------------------------------
    const SIM_RawField &surface = *surfaceField->getField();
    std::vector<float> store[13];
    AvsFields in{};
    in.size = sizeof in;
    surface.getVoxelRes(in.res[0], in.res[1], in.res[2]);
    const UT_Vector3 orig = surface.getOrig();
    in.origin[0] = orig.x(); in.origin[1] = orig.y(); in.origin[2] = orig.z();
    in.dx = velocityField->getVoxelSize().maxComponent();                 // AV.cpp:242
    in.surface = flatten(surface, store[0]);
    in.viscosity = flatten(*viscosityField->getField(), store[1]);
    in.density = flatten(*densityField->getField(), store[2]);
    in.collision = flatten(*solidField->getField(), store[3]);
    for (int a = 0; a < 3; ++a) {
        // velocity and face weights always travel dense: Houdini may hold a component as constant tiles (a w = 0 drop), but the
        // solver reads them per face and writes the velocity back in place
        in.vel[a] = flatten(*velocityField->getField(a), store[4 + a], false);
        in.face_weights[a] = flatten(*faceWeightsField->getField(a), store[7 + a], false);
        in.collision_vel[a] = flatten(*solidVelocityField->getField(a), store[10 + a]);
    }

    AvsParams p;
    avs_default_params(&p);
    p.dt = timestep;                                                        // AV.cpp:130
    p.tolerance = getSolverTolerance();
    p.max_iterations = getMaxIterations();
    p.number_super_samples = getNumberSuperSamples();
    p.octree_levels = getOctreeLevels();
    p.fine_bandwidth = getFineBandwidth();
    p.use_enhanced_gradients = getUseEnhancedGradients();
    p.do_apply_solid_weights = getDoApplySolidWeights();
    p.extrapolation = getExtrapolation();
    p.precision = getSinglePrecision() ? AVS_PRECISION_F32 : AVS_PRECISION_F64;
    // AvsParams.cancel stays NULL: UT_Interrupt has no asynchronous callback this shim could set a flag from, and a non-NULL
    // pointer that nothing sets would only chunk the persistent CG kernel (256 iterations per launch) for no benefit.  A host
    // that owns a watcher thread polling opInterrupt() can pass its flag here; AVS_ERR_CANCELLED then maps to a quiet return.
    UT_Interrupt *boss = UTgetInterrupt();
    p.cancel = nullptr;

    AvsResult r{};
    r.size = sizeof r;

    // ---- octree geometry dump (AV.cpp:283-294) ---------------------------------------------------------------------
    if (getDoPrintOctree() && getOnlyPrintOctree()) {
        int rc = avs_build_octree(ctx, &in, &p, &r);
        if (rc != AVS_OK) return fail(avs_status_string(rc));
    } else {
        // ---- the solve: replaces HDK_AdaptiveViscosity.cpp:233-707 ---------------------------------------------------
        std::vector<float> outv[3];
        AvsVelocityOut out{};
        out.on_device = 0;
        for (int a = 0; a < 3; ++a) {
            outv[a] = store[4 + a];            // in-place semantics: faces the solver does not own keep their value
            out.vel[a] = outv[a].data();
        }
        UT_PerfMonAutoSolveEvent event(this, "Solve Linear System");
        int rc = myMulti ? avs_solve_multi(myMulti, &in, &p, &out, &r) : avs_solve(ctx, &in, &p, &out, &r);
        if (rc == AVS_ERR_CANCELLED) return true;                                                        // user interrupt: quiet
        if (rc != AVS_OK) return fail((rc == AVS_ERR_CUDA || rc == AVS_ERR_UNSUPPORTED || myMulti) ? avs_last_error() : avs_status_string(rc));   // AV.cpp:621-622
        for (int a = 0; a < 3; ++a)              // applyVelocitiesToRegularGrid (AV.cpp:696-706)
            velocityField->getField(a)->fieldNC()->extractFromFlattened(outv[a].data(), in.vel[a].res[0],
                                                                         exint(in.vel[a].res[0]) * in.vel[a].res[1]);
        velocityField->pubHandleModification();
        UT_WorkBuffer extra;                     // AV.cpp:645-652
        extra.sprintf("iterations=%d, error=%.6f, octree DOFS=%d, regular DOFs=%d", r.iterations, r.error,
                      int(r.octree_dofs), int(r.regular_dofs));
        event.setExtraInfo(extra.buffer());
    }
    if (getDoPrintOctree()) {
        // HDK_OctreeGrid::outputOctreeGeometry (HDK_OctreeGrid.cpp:245-308): P, pscale, octreeLevel per ACTIVE cell
        int64_t n = 0;
        if (avs_get_octree_points(ctx, &n, nullptr, nullptr, nullptr) != AVS_OK) return fail("octree geometry unavailable");
        std::vector<float> pos(size_t(n) * 3), pscale(n);
        std::vector<int32_t> level(n);
        if (n && avs_get_octree_points(ctx, &n, pos.data(), pscale.data(), level.data()) != AVS_OK) return fail("octree geometry unavailable");
        SIM_GeometryCopy *geo = getOrCreateGeometry(obj, "octreeGeometry");
        SIM_GeometryAutoWriteLock lock(geo, SIM_DATA_ID_PRESERVE);
        GU_Detail &gdp = lock.getGdp();
        gdp.clear();
        GA_RWHandleF scaleH(gdp.addFloatTuple(GA_ATTRIB_POINT, "pscale", 1, GA_Defaults(0)));
        GA_RWHandleI levelH(gdp.addIntTuple(GA_ATTRIB_POINT, "octreeLevel", 1, GA_Defaults(-1)));
        const GA_Offset first = gdp.appendPointBlock(n);
        for (int64_t i = 0; i < n; ++i) {
            if (!(i & 0xffff) && boss->opInterrupt()) break;
            gdp.setPos3(first + i, UT_Vector3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
            scaleH.set(first + i, pscale[i]);
            levelH.set(first + i, level[i]);
        }
        gdp.getAttributes().bumpAllDataIds(GA_ATTRIB_POINT);
    }
    return true;
}
