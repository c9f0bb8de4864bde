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

class GAS_API HDK_AdaptiveViscosity : public GAS_SubSolver
{
public:
    // the options the reference reads (HDK_AdaptiveViscosity.h:28-41) -- including the two it reads under names no
    // parameter defines ("fineBandwidth", "doApplySolidWeights"), so behaviour matches bit for bit
    GET_DATA_FUNC_F(SIM_NAME_TOLERANCE, SolverTolerance);
    GET_DATA_FUNC_I("maxIterations", MaxIterations);
    GET_DATA_FUNC_I("numberSuperSamples", NumberSuperSamples);
    GET_DATA_FUNC_I("octreeLevels", OctreeLevels);
    GET_DATA_FUNC_I("fineBandwidth", FineBandwidth);
    GET_DATA_FUNC_B("useEnhancedGradients", UseEnhancedGradients);
    GET_DATA_FUNC_B("doApplySolidWeights", DoApplySolidWeights);
    GET_DATA_FUNC_B("doPrintOctree", DoPrintOctree);
    GET_DATA_FUNC_B("onlyPrintOctree", OnlyPrintOctree);
    GET_DATA_FUNC_F("extrapolation", Extrapolation);
    // additions of this build
    GET_DATA_FUNC_I("cudaDevice", CudaDevice);
    GET_DATA_FUNC_B("singlePrecision", SinglePrecision);

protected:
    explicit HDK_AdaptiveViscosity(const SIM_DataFactory *factory);
    ~HDK_AdaptiveViscosity() override;

    bool solveGasSubclass(SIM_Engine &engine, SIM_Object *obj, SIM_Time time, SIM_Time timestep) override;

private:
    static const SIM_DopDescription *getDopDescription();
    AvsContext *context(SIM_Object *obj);

    AvsContext *myContext = nullptr;   // device buffers are cached across cooks (the reference reallocates per call)
    int myContextDevice = -1;

    DECLARE_STANDARD_GETCASTTOTYPE();
    DECLARE_DATAFACTORY(HDK_AdaptiveViscosity, GAS_SubSolver, "HDK Adaptive Viscosity", getDopDescription());
};

#endif
