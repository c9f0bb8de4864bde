// HDK_AdaptiveViscosityB200.h -- Houdini DOP micro-solver that keeps the reference's plugin surface and hands the solve to
// libavs_b200.so (include/avs.h).
//
// Drop-in for Source/HDK_AdaptiveViscosity.{h,cpp} of rgoldade/AdaptiveViscositySolver: same DOP data type
// ("HDK_AdaptiveViscosity", label "HDK Adaptive Viscosity", HDK_AdaptiveViscosity.h:67-69), same parameter names
// (HDK_AdaptiveViscosity.cpp:36-116) and the same option getters (HDK_AdaptiveViscosity.h:28-41), so scenes that use the
// reference DSO (Scenes/*.hip) cook unchanged.  Only this host shim needs the HDK; it is built when $HFS is set
// (CMakeLists.txt beside this file) and is NOT compiled in the development image (no Houdini there) -- the same boundary
// is exercised from Python in tests/ (adaptiveviscositysolver_b200/solver.py mirrors this class method by method).
#ifndef HDK_ADAPTIVEVISCOSITY_B200_H
#define HDK_ADAPTIVEVISCOSITY_B200_H

#include <GAS/GAS_SubSolver.h>
#include <GAS/GAS_Utils.h>

struct AvsContext;
struct AvsMulti;

class GAS_API HDK_AdaptiveViscosity : public GAS_SubSolver
{
public:
    // the options the reference reads (HDK_AdaptiveViscosity.h:28-41) -- including the two it reads under names no
    // parameter defines ("fineBandwidth", "doApplySolidWeights"), so behaviour matches bit for bit
    GET_DATA_FUNC_F(SIM_NAME_TOLERANCE, SolverTolerance);            // AV.h:28  -> AvsParams.tolerance
    GET_DATA_FUNC_I("maxIterations", MaxIterations);                 // AV.h:29  -> max_iterations
    GET_DATA_FUNC_I("numberSuperSamples", NumberSuperSamples);       // AV.h:30  -> number_super_samples
    GET_DATA_FUNC_I("octreeLevels", OctreeLevels);                   // AV.h:31  -> octree_levels
    GET_DATA_FUNC_I("fineBandwidth", FineBandwidth);                 // AV.h:33  -> fine_bandwidth (no parameter of that name: stays 0)
    GET_DATA_FUNC_B("useEnhancedGradients", UseEnhancedGradients);   // AV.h:35  -> use_enhanced_gradients
    GET_DATA_FUNC_B("doApplySolidWeights", DoApplySolidWeights);     // AV.h:37  -> do_apply_solid_weights (likewise unset)
    GET_DATA_FUNC_B("doPrintOctree", DoPrintOctree);                 // AV.h:39  -> avs_get_octree_points
    GET_DATA_FUNC_B("onlyPrintOctree", OnlyPrintOctree);             // AV.h:40  -> avs_build_octree
    GET_DATA_FUNC_F("extrapolation", Extrapolation);                 // AV.h:41  -> extrapolation
    // additions of this build
    GET_DATA_FUNC_I("cudaDevice", CudaDevice);                       // first CUDA device ordinal
    GET_DATA_FUNC_I("cudaDeviceCount", CudaDeviceCount);             // > 1: row-partition the solve over that many consecutive GPUs
    GET_DATA_FUNC_B("singlePrecision", SinglePrecision);

protected:
    explicit HDK_AdaptiveViscosity(const SIM_DataFactory *factory);
    ~HDK_AdaptiveViscosity() override;

    bool solveGasSubclass(SIM_Engine &engine, SIM_Object *obj, SIM_Time time, SIM_Time timestep) override;

private:
    static const SIM_DopDescription *getDopDescription();
    AvsContext *context(SIM_Object *obj);

    AvsContext *myContext = nullptr;   // device buffers are cached across cooks (the reference reallocates per call)
    AvsMulti *myMulti = nullptr;       // cudaDeviceCount > 1: one rank context per GPU, all driven from this cook thread
    int myContextDevice = -1, myContextCount = 0;

    DECLARE_STANDARD_GETCASTTOTYPE();
    DECLARE_DATAFACTORY(HDK_AdaptiveViscosity, GAS_SubSolver, "HDK Adaptive Viscosity", getDopDescription());
};

#endif
