// =============================================================================
// mock_hdk.h -- a dense, single-header stand-in for the slice of the Houdini HDK
// (and of Eigen 3) that rgoldade/AdaptiveViscositySolver touches.
//
// TEST INFRASTRUCTURE ONLY (lives under oracle/).  Purpose: compile the
// reference's OWN sources -- /root/reference/Source/HDK_AdaptiveViscosity.cpp,
// HDK_OctreeGrid.cpp, HDK_OctreeVectorFieldInterpolator.cpp and their headers,
// UNCHANGED, straight from where they lie -- into oracle/_ref/libavs_ref.so, so
// that the restated oracle (oracle/avs_oracle.cpp) and the CUDA path can be
// checked against the reference's real control flow instead of against a
// re-reading of it.  Nothing of the reference is copied into this repository;
// this file only supplies the types and calls the reference expects from
// <SIM/...>, <UT/...>, <GAS/...>, <GU/...>, <PRM/...>, <SYS/...> and "Eigen/Sparse".
//
// What is REAL when the reference runs on this mock: every line of the reference
// (labelling rules, tile-skipping logic, stencils, weights, assembly, restriction,
// interpolation, the order of operations).  What is ASSUMED (Houdini and Eigen are
// closed / absent here; SURVEY.md Appendix D lists the same assumptions, and
// oracle/avs_oracle.cpp fixes the same ones):
//   * UT_VoxelArray: 16^3 tiles, x-fastest, tiles are either constant or dense;
//     writing the tile's own constant into a constant tile keeps it constant, any
//     other write expands it; compress-on-exit / collapseAllTiles re-compress tiles
//     whose voxels are all equal; out-of-range reads clamp (UT_VOXELBORDER_STREAK);
//   * SIM_RawField: sample layout per SIM_FieldSample, voxel size = size / cell
//     resolution rounded to float32 (UT_Vector3 is a float vector), indexToPos =
//     orig + (index + 1/2 on cell-centred axes) * voxel size, posToIndex its inverse,
//     getValue(pos) = trilinear, clamp-to-edge, evaluated in fp64 as a + t (b - a),
//     x then y then z;
//   * computeSDFWeightsSampled(sdf, n, invert=false, minweight, dilate): fraction of
//     the n^3 sub-samples at ((k + 1/2)/n - 1/2) voxels around the sample whose
//     interpolated sdf minus dilate is < 0;
//   * setScaleDivideThreshold(1, nullptr, &b, 0): a <- a / b where b > 0;
//   * UT_ThreadedAlgorithm: THREADED_METHODn runs the Partial method once per job;
//     jobs = mock_hdk::threads() (default 1 -> deterministic order);
//   * Eigen: setFromTriplets sums duplicates; ConjugateGradient<.., Lower|Upper> with
//     the default DiagonalPreconditioner is upstream Eigen 3.3/3.4's
//     conjugate_gradient() loop (restated from the published source, call site
//     HDK_AdaptiveViscosity.cpp:611-630).
// MOCK_HDK_EXACT_POSITIONS (default 1): positions handed out by indexToPos / posToIndex
// keep their fp64 value inside UT_Vector3 (the convention oracle/avs_oracle.cpp and the
// CUDA path use); with 0 they are rounded to float32 like a real UT_Vector3.  Arithmetic
// the REFERENCE performs on UT_Vector3 components is always rounded to float32.
// =============================================================================
#pragma once
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#ifndef MOCK_HDK_EXACT_POSITIONS
#define MOCK_HDK_EXACT_POSITIONS 1
#endif

// ---- SYS -----------------------------------------------------------------------
typedef int64_t exint;
typedef int64_t int64;
typedef int32_t int32;
typedef double fpreal;
typedef float fpreal32;
typedef double fpreal64;
#define SYS_FORCE_INLINE inline
#define GAS_API
#define SIM_API

template <class A, class B>
inline auto SYSmax(A a, B b) -> typename std::common_type<A, B>::type { typedef typename std::common_type<A, B>::type C; return (C)a > (C)b ? (C)a : (C)b; }
template <class A, class B>
inline auto SYSmin(A a, B b) -> typename std::common_type<A, B>::type { typedef typename std::common_type<A, B>::type C; return (C)a < (C)b ? (C)a : (C)b; }

namespace mock_hdk {
inline int &threadsRef() { static int t = 1; return t; }
inline int threads() { return threadsRef(); }
inline void setThreads(int t) { threadsRef() = t < 1 ? 1 : t; }

// A float32 scalar that can carry an fp64 payload.  Assignments and compound assignments made by the reference round to
// float32 (what storing into a UT_Vector3 component does); the mock's own position functions may store unrounded doubles.
struct Real32 {
    double v = 0;
    Real32() {}
    Real32(double x) : v((double)(float)x) {}
    static Real32 exact(double x) { Real32 r; r.v = MOCK_HDK_EXACT_POSITIONS ? x : (double)(float)x; return r; }
    operator double() const { return v; }
    Real32 &operator=(double x) { v = (double)(float)x; return *this; }
    Real32 &operator+=(double x) { v = (double)(float)(v + x); return *this; }
    Real32 &operator-=(double x) { v = (double)(float)(v - x); return *this; }
    Real32 &operator*=(double x) { v = (double)(float)(v * x); return *this; }
    Real32 &operator/=(double x) { v = (double)(float)(v / x); return *this; }
};
// float * float stays float in C++: emulate for component products (HDK_AdaptiveViscosity.cpp:2056)
inline Real32 operator*(const Real32 &a, const Real32 &b) { return Real32(a.v * b.v); }
inline Real32 operator+(const Real32 &a, const Real32 &b) { return Real32(a.v + b.v); }
inline Real32 operator-(const Real32 &a, const Real32 &b) { return Real32(a.v - b.v); }
// float (op) double promotes to double
#define MOCK_REAL32_MIXED(OP)                                                                  \
    template <class A, class = typename std::enable_if<std::is_arithmetic<A>::value>::type>    \
    inline double operator OP(const Real32 &a, A b) { return a.v OP (double)b; }               \
    template <class A, class = typename std::enable_if<std::is_arithmetic<A>::value>::type>    \
    inline double operator OP(A a, const Real32 &b) { return (double)a OP b.v; }
MOCK_REAL32_MIXED(+)
MOCK_REAL32_MIXED(-)
MOCK_REAL32_MIXED(*)
MOCK_REAL32_MIXED(/)
#undef MOCK_REAL32_MIXED
}  // namespace mock_hdk

template <class V, class L, class H>
inline double SYSclamp(V v, L lo, H hi) { double x = (double)v; return x < (double)lo ? (double)lo : (x > (double)hi ? (double)hi : x); }

// ---- UT vectors ------------------------------------------------------------------
template <class T>
struct UT_Vector3T {
    T vec[3];
    UT_Vector3T() { vec[0] = vec[1] = vec[2] = T(); }
    explicit UT_Vector3T(T s) { vec[0] = vec[1] = vec[2] = s; }
    UT_Vector3T(T a, T b, T c) { vec[0] = a; vec[1] = b; vec[2] = c; }
    template <class U>
    UT_Vector3T(const UT_Vector3T<U> &o) { for (int i = 0; i < 3; ++i) vec[i] = (T)o.vec[i]; }
    T &operator[](int i) { return vec[i]; }
    const T &operator[](int i) const { return vec[i]; }
    T &operator()(int i) { return vec[i]; }
    const T &operator()(int i) const { return vec[i]; }
    T &x() { return vec[0]; } T &y() { return vec[1]; } T &z() { return vec[2]; }
    const T &x() const { return vec[0]; } const T &y() const { return vec[1]; } const T &z() const { return vec[2]; }
    UT_Vector3T &operator/=(const UT_Vector3T &o) { for (int i = 0; i < 3; ++i) vec[i] /= o.vec[i]; return *this; }
    UT_Vector3T &operator*=(const UT_Vector3T &o) { for (int i = 0; i < 3; ++i) vec[i] *= o.vec[i]; return *this; }
    UT_Vector3T &operator+=(const UT_Vector3T &o) { for (int i = 0; i < 3; ++i) vec[i] += o.vec[i]; return *this; }
    UT_Vector3T &operator-=(const UT_Vector3T &o) { for (int i = 0; i < 3; ++i) vec[i] -= o.vec[i]; return *this; }
    UT_Vector3T &operator*=(T s) { for (int i = 0; i < 3; ++i) vec[i] *= s; return *this; }
    UT_Vector3T &operator/=(T s) { for (int i = 0; i < 3; ++i) vec[i] /= s; return *this; }
    bool operator==(const UT_Vector3T &o) const { return vec[0] == o.vec[0] && vec[1] == o.vec[1] && vec[2] == o.vec[2]; }
    bool operator!=(const UT_Vector3T &o) const { return !(*this == o); }
    T maxComponent() const { return std::max(vec[0], std::max(vec[1], vec[2])); }
};
template <class T> inline UT_Vector3T<T> operator+(UT_Vector3T<T> a, const UT_Vector3T<T> &b) { a += b; return a; }
template <class T> inline UT_Vector3T<T> operator-(UT_Vector3T<T> a, const UT_Vector3T<T> &b) { a -= b; return a; }
template <class T> inline UT_Vector3T<T> operator*(UT_Vector3T<T> a, const UT_Vector3T<T> &b) { a *= b; return a; }
template <class T> inline UT_Vector3T<T> operator*(UT_Vector3T<T> a, T s) { a *= s; return a; }
typedef UT_Vector3T<int32_t> UT_Vector3i;
typedef UT_Vector3T<int64_t> UT_Vector3I;

// UT_Vector3 (fpreal32 components): see mock_hdk::Real32
struct UT_Vector3 {
    mock_hdk::Real32 vec[3];
    UT_Vector3() {}
    explicit UT_Vector3(double s) { vec[0] = vec[1] = vec[2] = s; }
    UT_Vector3(double a, double b, double c) { vec[0] = a; vec[1] = b; vec[2] = c; }
    template <class U>
    UT_Vector3(const UT_Vector3T<U> &o) { for (int i = 0; i < 3; ++i) vec[i] = (double)o.vec[i]; }
    mock_hdk::Real32 &operator[](int i) { return vec[i]; }
    const mock_hdk::Real32 &operator[](int i) const { return vec[i]; }
    mock_hdk::Real32 &x() { return vec[0]; } mock_hdk::Real32 &y() { return vec[1]; } mock_hdk::Real32 &z() { return vec[2]; }
    const mock_hdk::Real32 &x() const { return vec[0]; } const mock_hdk::Real32 &y() const { return vec[1]; } const mock_hdk::Real32 &z() const { return vec[2]; }
    UT_Vector3 &operator*=(const UT_Vector3 &o) { for (int i = 0; i < 3; ++i) vec[i] *= (double)o.vec[i]; return *this; }
    UT_Vector3 &operator/=(const UT_Vector3 &o) { for (int i = 0; i < 3; ++i) vec[i] /= (double)o.vec[i]; return *this; }
    UT_Vector3 &operator+=(const UT_Vector3 &o) { for (int i = 0; i < 3; ++i) vec[i] += (double)o.vec[i]; return *this; }
    UT_Vector3 &operator-=(const UT_Vector3 &o) { for (int i = 0; i < 3; ++i) vec[i] -= (double)o.vec[i]; return *this; }
    UT_Vector3 &operator*=(double s) { for (int i = 0; i < 3; ++i) vec[i] *= s; return *this; }
    double maxComponent() const { return std::max((double)vec[0], std::max((double)vec[1], (double)vec[2])); }
    void setExact(int i, double v) { vec[i] = mock_hdk::Real32::exact(v); }
};
inline UT_Vector3 operator*(UT_Vector3 a, const UT_Vector3 &b) { a *= b; return a; }
inline UT_Vector3 operator+(UT_Vector3 a, const UT_Vector3 &b) { a += b; return a; }
inline UT_Vector3 operator-(UT_Vector3 a, const UT_Vector3 &b) { a -= b; return a; }
inline UT_Vector3 operator*(UT_Vector3 a, double s) { a *= s; return a; }

struct UT_Vector4i {
    int32_t vec[4];
    UT_Vector4i() { vec[0] = vec[1] = vec[2] = vec[3] = 0; }
    UT_Vector4i(int a, int b, int c, int d) { vec[0] = a; vec[1] = b; vec[2] = c; vec[3] = d; }
    int32_t &operator[](int i) { return vec[i]; }
    const int32_t &operator[](int i) const { return vec[i]; }
    bool operator==(const UT_Vector4i &o) const { return vec[0] == o.vec[0] && vec[1] == o.vec[1] && vec[2] == o.vec[2] && vec[3] == o.vec[3]; }
    bool operator!=(const UT_Vector4i &o) const { return !(*this == o); }
};
template <class T, int N>
struct UT_FixedVector {
    T vec[N];
    UT_FixedVector() { for (int i = 0; i < N; ++i) vec[i] = T(); }
    T &operator[](int i) { return vec[i]; }
    const T &operator[](int i) const { return vec[i]; }
};

// ---- UT_Array ----------------------------------------------------------------------
template <class T>
class UT_Array {
    typedef typename std::conditional<std::is_same<T, bool>::value, unsigned char, T>::type S;
    std::vector<S> d;

public:
    typedef T value_type;
    UT_Array() {}
    exint entries() const { return (exint)d.size(); }
    exint size() const { return (exint)d.size(); }
    bool isEmpty() const { return d.empty(); }
    void setSize(exint n) { d.resize((size_t)n); }
    void setSizeNoInit(exint n) { d.resize((size_t)n); }
    void setCapacity(exint n) { d.reserve((size_t)n); }
    void bumpCapacity(exint n) { if ((size_t)n > d.capacity()) d.reserve((size_t)n); }
    void clear() { d.clear(); }
    void constant(const T &v) { for (auto &e : d) e = (S)v; }
    exint append(const T &v) { d.push_back((S)v); return (exint)d.size() - 1; }
    void concat(const UT_Array<T> &o) { d.insert(d.end(), o.d.begin(), o.d.end()); }
    T &operator[](exint i) { return *reinterpret_cast<T *>(&d[(size_t)i]); }
    const T &operator[](exint i) const { return *reinterpret_cast<const T *>(&d[(size_t)i]); }
    T &operator()(exint i) { return (*this)[i]; }
    const T &operator()(exint i) const { return (*this)[i]; }
    T &last() { return (*this)[entries() - 1]; }
    T *begin() { return reinterpret_cast<T *>(d.data()); }
    T *end() { return reinterpret_cast<T *>(d.data() + d.size()); }
    const T *begin() const { return reinterpret_cast<const T *>(d.data()); }
    const T *end() const { return reinterpret_cast<const T *>(d.data() + d.size()); }
};

// ---- threading / interrupt / perf monitor stand-ins -------------------------------
class UT_JobInfo {
    int myJob, myNumJobs;

public:
    UT_JobInfo(int job, int numJobs) : myJob(job), myNumJobs(numJobs) {}
    int job() const { return myJob; }
    int numJobs() const { return myNumJobs; }
    void divideWork(exint units, exint &start, exint &end) const {
        start = units * myJob / myNumJobs;
        end = units * (myJob + 1) / myNumJobs;
    }
};
struct UT_Thread {
    static int getNumProcessors() { return mock_hdk::threads(); }
};
class UT_Interrupt {
public:
    bool opInterrupt(int = -1) { return false; }
    bool opStart(const char * = nullptr) { return true; }
    void opEnd() {}
};
inline UT_Interrupt *UTgetInterrupt() { static UT_Interrupt boss; return &boss; }

template <class T>
struct UT_BlockedRange {
    T b, e;
    UT_BlockedRange(T b_, T e_) : b(b_), e(e_) {}
    T begin() const { return b; }
    T end() const { return e; }
};
template <class F>
inline void UTparallelForEachNumber(int64 n, const F &f) {
    const int jobs = (int)std::min<int64>(mock_hdk::threads(), std::max<int64>(n, 1));
    if (jobs <= 1) { f(UT_BlockedRange<int64>(0, n)); return; }
    std::vector<std::thread> pool;
    for (int j = 0; j < jobs; ++j) pool.emplace_back([&, j]() { f(UT_BlockedRange<int64>(n * j / jobs, n * (j + 1) / jobs)); });
    for (auto &t : pool) t.join();
}

namespace mock_hdk {
// runs fn(info) once per job; jobs > 1 only when the reference's own "should multithread" predicate says so
template <class F>
inline void runJobs(bool multi, const F &fn) {
    const int jobs = multi ? threads() : 1;
    if (jobs <= 1) { fn(UT_JobInfo(0, 1)); return; }
    std::vector<std::thread> pool;
    for (int j = 0; j < jobs; ++j) pool.emplace_back([&fn, j, jobs]() { fn(UT_JobInfo(j, jobs)); });
    for (auto &t : pool) t.join();
}
}  // namespace mock_hdk

// UT_ThreadedAlgorithm.h: THREADED_METHODn(CLASS, DOMULTI, METHOD, T1, P1, ...) declares METHOD(P1..) that fans out to
// METHODPartial(P1.., const UT_JobInfo &)
#define MOCK_TM_BODY(DOMULTI, CALL) { mock_hdk::runJobs((DOMULTI), [&](const UT_JobInfo &info) { CALL; }); }
#define THREADED_METHOD1(C, M, F, T1, P1) void F(T1 P1) MOCK_TM_BODY(M, F##Partial(P1, info))
#define THREADED_METHOD2(C, M, F, T1, P1, T2, P2) void F(T1 P1, T2 P2) MOCK_TM_BODY(M, F##Partial(P1, P2, info))
#define THREADED_METHOD3(C, M, F, T1, P1, T2, P2, T3, P3) void F(T1 P1, T2 P2, T3 P3) MOCK_TM_BODY(M, F##Partial(P1, P2, P3, info))
#define THREADED_METHOD4(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4) void F(T1 P1, T2 P2, T3 P3, T4 P4) MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, info))
#define THREADED_METHOD5(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5) MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, info))
#define THREADED_METHOD1_CONST(C, M, F, T1, P1) void F(T1 P1) const MOCK_TM_BODY(M, F##Partial(P1, info))
#define THREADED_METHOD2_CONST(C, M, F, T1, P1, T2, P2) void F(T1 P1, T2 P2) const MOCK_TM_BODY(M, F##Partial(P1, P2, info))
#define THREADED_METHOD3_CONST(C, M, F, T1, P1, T2, P2, T3, P3) void F(T1 P1, T2 P2, T3 P3) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, info))
#define THREADED_METHOD4_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, info))
#define THREADED_METHOD5_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, info))
#define THREADED_METHOD6_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5, T6, P6) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5, T6 P6) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, P6, info))
#define THREADED_METHOD7_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5, T6, P6, T7, P7) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5, T6 P6, T7 P7) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, P6, P7, info))
#define THREADED_METHOD8_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5, T6, P6, T7, P7, T8, P8) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5, T6 P6, T7 P7, T8 P8) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, P6, P7, P8, info))
#define THREADED_METHOD9_CONST(C, M, F, T1, P1, T2, P2, T3, P3, T4, P4, T5, P5, T6, P6, T7, P7, T8, P8, T9, P9) \
    void F(T1 P1, T2 P2, T3 P3, T4 P4, T5 P5, T6 P6, T7 P7, T8 P8, T9 P9) const MOCK_TM_BODY(M, F##Partial(P1, P2, P3, P4, P5, P6, P7, P8, P9, info))

class UT_WorkBuffer {
    std::string s;

public:
    void sprintf(const char *fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s = buf;
    }
    const char *buffer() const { return s.c_str(); }
};

namespace mock_hdk {
// Observation hooks: the harness (oracle/ref_harness.cpp) installs callbacks that fire when a PerfMon scope closes or the
// linear solve runs -- the reference keeps everything in locals of solveGasSubclass, so this is how results are read out.
struct Hooks {
    std::function<void(const char *label)> scopeBegin, scopeEnd;
    std::function<void(const char *info)> extraInfo;
};
inline Hooks &hooks() { static Hooks h; return h; }
}  // namespace mock_hdk

class UT_PerfMonAutoSolveEvent {
    std::string label;

public:
    template <class S>
    UT_PerfMonAutoSolveEvent(const S *, const char *l) : label(l) { if (mock_hdk::hooks().scopeBegin) mock_hdk::hooks().scopeBegin(label.c_str()); }
    ~UT_PerfMonAutoSolveEvent() { if (mock_hdk::hooks().scopeEnd) mock_hdk::hooks().scopeEnd(label.c_str()); }
    void setExtraInfo(const char *info) { if (mock_hdk::hooks().extraInfo) mock_hdk::hooks().extraInfo(info); }
};

// ---- UT_VoxelArray ---------------------------------------------------------------------
template <class T>
class UT_VoxelTile {
public:
    int res[3] = {0, 0, 0};
    bool constantTile = true;
    T constantValue = T();
    std::vector<T> data;   // dense storage when !constantTile (x fastest)

    int xres() const { return res[0]; }
    int yres() const { return res[1]; }
    int zres() const { return res[2]; }
    int numVoxels() const { return res[0] * res[1] * res[2]; }
    bool isConstant() const { return constantTile; }
    T operator()(int x, int y, int z) const { return constantTile ? constantValue : data[(size_t)x + (size_t)res[0] * ((size_t)y + (size_t)res[1] * z)]; }
    void makeConstant(T v) {
        constantTile = true;
        constantValue = v;
        std::vector<T>().swap(data);
    }
    void uncompress() {
        if (!constantTile) return;
        data.assign((size_t)numVoxels(), constantValue);
        constantTile = false;
    }
    void setValue(int x, int y, int z, T v) {
        if (constantTile) {
            if (constantValue == v) return;   // UT_VoxelTile::writeThrough: the tile's own constant needs no expansion
            uncompress();
        }
        data[(size_t)x + (size_t)res[0] * ((size_t)y + (size_t)res[1] * z)] = v;
    }
    void tryCompress() {
        if (constantTile || data.empty()) return;
        const T v = data[0];
        for (const T &e : data)
            if (!(e == v)) return;
        makeConstant(v);
    }
};

template <class T>
class UT_VoxelArray {
public:
    enum { TILEBITS = 4, TILESIZE = 16, TILEMASK = 15 };
    int myRes[3] = {0, 0, 0};
    int myTileRes[3] = {0, 0, 0};
    mutable std::vector<UT_VoxelTile<T>> myTiles;   // getTile()/getLinearTile() hand out non-const tiles from a const array, as in the HDK

    void size(int x, int y, int z) {
        myRes[0] = x; myRes[1] = y; myRes[2] = z;
        for (int a = 0; a < 3; ++a) myTileRes[a] = (myRes[a] + TILESIZE - 1) >> TILEBITS;
        myTiles.assign((size_t)myTileRes[0] * myTileRes[1] * myTileRes[2], UT_VoxelTile<T>());
        for (int tz = 0; tz < myTileRes[2]; ++tz)
            for (int ty = 0; ty < myTileRes[1]; ++ty)
                for (int tx = 0; tx < myTileRes[0]; ++tx) {
                    UT_VoxelTile<T> &t = myTiles[(size_t)tx + (size_t)myTileRes[0] * ((size_t)ty + (size_t)myTileRes[1] * tz)];
                    t.res[0] = std::min((int)TILESIZE, myRes[0] - tx * TILESIZE);
                    t.res[1] = std::min((int)TILESIZE, myRes[1] - ty * TILESIZE);
                    t.res[2] = std::min((int)TILESIZE, myRes[2] - tz * TILESIZE);
                }
    }
    int getXRes() const { return myRes[0]; }
    int getYRes() const { return myRes[1]; }
    int getZRes() const { return myRes[2]; }
    int getRes(int a) const { return myRes[a]; }
    int getTileRes(int a) const { return myTileRes[a]; }
    int numTiles() const { return (int)myTiles.size(); }
    exint numVoxels() const { return (exint)myRes[0] * myRes[1] * myRes[2]; }
    bool isValidIndex(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < myRes[0] && y < myRes[1] && z < myRes[2]; }
    int indexToLinearTile(int x, int y, int z) const { return (x >> TILEBITS) + myTileRes[0] * ((y >> TILEBITS) + myTileRes[1] * (z >> TILEBITS)); }
    void linearTileToXYZ(int idx, int &x, int &y, int &z) const {
        x = idx % myTileRes[0];
        idx /= myTileRes[0];
        y = idx % myTileRes[1];
        z = idx / myTileRes[1];
    }
    UT_VoxelTile<T> *getTile(int tx, int ty, int tz) const { return &myTiles[(size_t)tx + (size_t)myTileRes[0] * ((size_t)ty + (size_t)myTileRes[1] * tz)]; }
    UT_VoxelTile<T> *getLinearTile(int idx) const { return &myTiles[(size_t)idx]; }
    T getValue(int x, int y, int z) const { return (*getLinearTile(indexToLinearTile(x, y, z)))(x & TILEMASK, y & TILEMASK, z & TILEMASK); }
    // operator(): out-of-range indices clamp (UT_VOXELBORDER_STREAK)
    T operator()(int x, int y, int z) const {
        x = std::min(std::max(x, 0), myRes[0] - 1);
        y = std::min(std::max(y, 0), myRes[1] - 1);
        z = std::min(std::max(z, 0), myRes[2] - 1);
        return getValue(x, y, z);
    }
    void setValue(int x, int y, int z, T v) { getLinearTile(indexToLinearTile(x, y, z))->setValue(x & TILEMASK, y & TILEMASK, z & TILEMASK, v); }
    void constant(T v) { for (auto &t : myTiles) t.makeConstant(v); }
    bool isConstant(T *v = nullptr) const {
        if (myTiles.empty()) return true;
        const T c = myTiles[0].constantValue;
        for (const auto &t : myTiles)
            if (!t.constantTile || !(t.constantValue == c)) return false;
        if (v) *v = c;
        return true;
    }
    void collapseAllTiles() { for (auto &t : myTiles) t.tryCompress(); }
    // flat copies: element (x, y, z) at x * xstride + y * ystride + z * zstride
    void flatten(T *flat, exint xstride, exint ystride, exint zstride) const {
        for (int z = 0; z < myRes[2]; ++z)
            for (int y = 0; y < myRes[1]; ++y)
                for (int x = 0; x < myRes[0]; ++x) flat[x * xstride + y * ystride + z * zstride] = getValue(x, y, z);
    }
    void extractFromFlattened(const T *flat, exint ystride, exint zstride) {
        for (int z = 0; z < myRes[2]; ++z)
            for (int y = 0; y < myRes[1]; ++y)
                for (int x = 0; x < myRes[0]; ++x) setValue(x, y, z, flat[x + y * ystride + z * zstride]);
        collapseAllTiles();
    }
};

template <class T>
class UT_VoxelArrayIterator {
public:
    UT_VoxelArray<T> *myArray = nullptr;
    int myTileStart = 0, myTileEnd = 0;   // public in the HDK too (HDK_AdaptiveViscosity.cpp:824-825 sets them)
    int myCurTile = 0;
    int myPos[3] = {0, 0, 0};             // voxel position inside the current tile
    bool myCompressOnExit = false;

    UT_VoxelArrayIterator() {}
    explicit UT_VoxelArrayIterator(UT_VoxelArray<T> *a) { setArray(a); }
    void setArray(UT_VoxelArray<T> *a) {
        myArray = a;
        myTileStart = 0;
        myTileEnd = a ? a->numTiles() : 0;
        myCurTile = myTileEnd;
    }
    void setConstArray(const UT_VoxelArray<T> *a) { setArray(const_cast<UT_VoxelArray<T> *>(a)); }
    void setCompressOnExit(bool b) { myCompressOnExit = b; }
    void setPartialRange(int idx, int numranges) {
        const int n = myArray ? myArray->numTiles() : 0;
        myTileStart = (int)((int64)n * idx / numranges);
        myTileEnd = (int)((int64)n * (idx + 1) / numranges);
    }
    void splitByTile(const UT_JobInfo &info) { setPartialRange(info.job(), info.numJobs()); }
    void rewind() {
        myCurTile = myTileStart;
        myPos[0] = myPos[1] = myPos[2] = 0;
        if (myCurTile > myTileEnd) myCurTile = myTileEnd;
    }
    bool atEnd() const { return myCurTile >= myTileEnd; }
    UT_VoxelTile<T> *getTile() const { return myArray->getLinearTile(myCurTile); }
    int getLinearTileNum() const { return myCurTile; }
    bool isTileConstant() const { return getTile()->isConstant(); }
    bool isStartOfTile() const { return myPos[0] == 0 && myPos[1] == 0 && myPos[2] == 0; }
    void leaveTile() { if (myCompressOnExit && myCurTile < myTileEnd) getTile()->tryCompress(); }
    void advanceTile() {
        leaveTile();
        ++myCurTile;
        myPos[0] = myPos[1] = myPos[2] = 0;
    }
    void advance() {
        UT_VoxelTile<T> *t = getTile();
        if (++myPos[0] < t->res[0]) return;
        myPos[0] = 0;
        if (++myPos[1] < t->res[1]) return;
        myPos[1] = 0;
        if (++myPos[2] < t->res[2]) return;
        advanceTile();
    }
    void tileOrigin(int &ox, int &oy, int &oz) const {
        int tx, ty, tz;
        myArray->linearTileToXYZ(myCurTile, tx, ty, tz);
        ox = tx * UT_VoxelArray<T>::TILESIZE; oy = ty * UT_VoxelArray<T>::TILESIZE; oz = tz * UT_VoxelArray<T>::TILESIZE;
    }
    int x() const { int ox, oy, oz; tileOrigin(ox, oy, oz); return ox + myPos[0]; }
    int y() const { int ox, oy, oz; tileOrigin(ox, oy, oz); return oy + myPos[1]; }
    int z() const { int ox, oy, oz; tileOrigin(ox, oy, oz); return oz + myPos[2]; }
    T getValue() const { return (*getTile())(myPos[0], myPos[1], myPos[2]); }
    void setValue(T v) { getTile()->setValue(myPos[0], myPos[1], myPos[2], v); }
    // voxels [start, end) of the current tile
    void getTileVoxels(UT_Vector3I &start, UT_Vector3I &end) const {
        int ox, oy, oz;
        tileOrigin(ox, oy, oz);
        UT_VoxelTile<T> *t = getTile();
        start[0] = ox; start[1] = oy; start[2] = oz;
        end[0] = ox + t->res[0]; end[1] = oy + t->res[1]; end[2] = oz + t->res[2];
    }
};

template <class T>
class UT_VoxelTileIterator {
    UT_VoxelTile<T> *myTile = nullptr;
    int myOrg[3] = {0, 0, 0};
    int myPos[3] = {0, 0, 0};
    bool myEnd = true;

public:
    void setTile(const UT_VoxelArrayIterator<T> &vit) {
        myTile = vit.getTile();
        vit.tileOrigin(myOrg[0], myOrg[1], myOrg[2]);
        rewind();
    }
    void rewind() {
        myPos[0] = myPos[1] = myPos[2] = 0;
        myEnd = !myTile || myTile->numVoxels() == 0;
    }
    bool atEnd() const { return myEnd; }
    void advance() {
        if (++myPos[0] < myTile->res[0]) return;
        myPos[0] = 0;
        if (++myPos[1] < myTile->res[1]) return;
        myPos[1] = 0;
        if (++myPos[2] < myTile->res[2]) return;
        myEnd = true;
    }
    int x() const { return myOrg[0] + myPos[0]; }
    int y() const { return myOrg[1] + myPos[1]; }
    int z() const { return myOrg[2] + myPos[2]; }
    T getValue() const { return (*myTile)(myPos[0], myPos[1], myPos[2]); }
    void setValue(T v) { myTile->setValue(myPos[0], myPos[1], myPos[2], v); }
};
typedef UT_VoxelArray<fpreal32> UT_VoxelArrayF;
typedef UT_VoxelArray<exint> UT_VoxelArrayI;
typedef UT_VoxelArrayIterator<fpreal32> UT_VoxelArrayIteratorF;
typedef UT_VoxelArrayIterator<exint> UT_VoxelArrayIteratorI;
typedef UT_VoxelTileIterator<fpreal32> UT_VoxelTileIteratorF;
typedef UT_VoxelTileIterator<exint> UT_VoxelTileIteratorI;

// ---- SIM fields ----------------------------------------------------------------------------
enum SIM_FieldSample {
    SIM_SAMPLE_CENTER, SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ, SIM_SAMPLE_CORNER,
    SIM_SAMPLE_EDGEXY, SIM_SAMPLE_EDGEXZ, SIM_SAMPLE_EDGEYZ
};

namespace mock_hdk {
// per axis: 1 when the sample sits on the voxel boundary (array one longer, no half-voxel offset)
inline void sampleNodeAxes(SIM_FieldSample s, int on[3]) {
    on[0] = on[1] = on[2] = 0;
    switch (s) {
        case SIM_SAMPLE_CENTER: break;
        case SIM_SAMPLE_FACEX: on[0] = 1; break;
        case SIM_SAMPLE_FACEY: on[1] = 1; break;
        case SIM_SAMPLE_FACEZ: on[2] = 1; break;
        case SIM_SAMPLE_CORNER: on[0] = on[1] = on[2] = 1; break;
        case SIM_SAMPLE_EDGEXY: on[0] = on[1] = 1; break;   // z-directed edge
        case SIM_SAMPLE_EDGEXZ: on[0] = on[2] = 1; break;   // y-directed edge
        case SIM_SAMPLE_EDGEYZ: on[1] = on[2] = 1; break;   // x-directed edge
    }
}

// Every field the reference initialises registers itself here (in init() order), so the harness can find the locals of
// solveGasSubclass (index grids, weights, octree labels) while they are alive.
struct FieldRegistry {
    struct Entry { const void *field; int isIndex; unsigned long long seq; };
    std::vector<Entry> live;
    unsigned long long nextSeq = 1;
    void add(const void *f, int isIndex) { remove(f); live.push_back(Entry{f, isIndex, nextSeq++}); }
    void remove(const void *f) {
        for (size_t i = 0; i < live.size(); ++i)
            if (live[i].field == f) { live.erase(live.begin() + (long)i); return; }
    }
    void moved(const void *from, const void *to) {
        for (auto &e : live)
            if (e.field == from) e.field = to;
    }
};
inline FieldRegistry &registry() { static FieldRegistry r; return r; }
inline bool &weightShortcutRef() { static bool b = true; return b; }
inline bool weightShortcut() { return weightShortcutRef(); }

template <class T>
class RawFieldT {
public:
    SIM_FieldSample mySample = SIM_SAMPLE_CENTER;
    double myOrigD[3] = {0, 0, 0};        // grid corner (fp64 payload of getOrig())
    double mySizeD[3] = {0, 0, 0};
    double myVoxelSizeD[3] = {1, 1, 1};   // float32-rounded
    int myCellRes[3] = {0, 0, 0};
    std::unique_ptr<UT_VoxelArray<T>> myField;

    // fields built from the caller's flat arrays sample with the caller's own origin / spacing (the C-ABI convention)
    bool myFlatSampling = false;
    double myFlatOrg[3] = {0, 0, 0};
    double myFlatDx = 1;

    RawFieldT() : myField(new UT_VoxelArray<T>()) {}
    RawFieldT(const RawFieldT &o) : myField(new UT_VoxelArray<T>()) { *this = o; }
    RawFieldT(RawFieldT &&o) : myField(new UT_VoxelArray<T>()) { *this = std::move(o); }
    ~RawFieldT() { registry().remove(this); }
    void copyHeader(const RawFieldT &o) {
        mySample = o.mySample;
        for (int a = 0; a < 3; ++a) {
            myOrigD[a] = o.myOrigD[a]; mySizeD[a] = o.mySizeD[a]; myVoxelSizeD[a] = o.myVoxelSizeD[a]; myCellRes[a] = o.myCellRes[a];
            myFlatOrg[a] = o.myFlatOrg[a];
        }
        myFlatSampling = o.myFlatSampling;
        myFlatDx = o.myFlatDx;
    }
    RawFieldT &operator=(RawFieldT &&o) {
        if (this == &o) return *this;
        copyHeader(o);
        myField.swap(o.myField);
        registry().remove(this);
        registry().moved(&o, this);
        return *this;
    }
    RawFieldT &operator=(const RawFieldT &o) {
        if (this == &o) return *this;
        copyHeader(o);
        *myField = *o.myField;
        return *this;
    }

    void init(SIM_FieldSample sample, const UT_Vector3 &orig, const UT_Vector3 &size, int xres, int yres, int zres) {
        mySample = sample;
        myCellRes[0] = xres; myCellRes[1] = yres; myCellRes[2] = zres;
        int on[3];
        sampleNodeAxes(sample, on);
        for (int a = 0; a < 3; ++a) {
            myOrigD[a] = (double)orig[a];
            mySizeD[a] = (double)size[a];
            myVoxelSizeD[a] = myCellRes[a] > 0 ? (double)(float)(mySizeD[a] / (double)myCellRes[a]) : 0.0;
        }
        myField->size(xres + on[0], yres + on[1], zres + on[2]);
        myField->constant(T());
        myFlatSampling = false;
        registry().add(this, std::is_same<T, exint>::value ? 1 : 0);
    }
    const UT_VoxelArray<T> *field() const { return myField.get(); }
    UT_VoxelArray<T> *fieldNC() const { return myField.get(); }
    SIM_FieldSample getSample() const { return mySample; }
    void makeConstant(T v) { myField->constant(v); }
    void getVoxelRes(int &x, int &y, int &z) const { x = myCellRes[0]; y = myCellRes[1]; z = myCellRes[2]; }
    int getXRes() const { return myField->getXRes(); }
    int getYRes() const { return myField->getYRes(); }
    int getZRes() const { return myField->getZRes(); }
    UT_Vector3 getOrig() const { UT_Vector3 v; for (int a = 0; a < 3; ++a) v.setExact(a, myOrigD[a]); return v; }
    UT_Vector3 getSize() const { UT_Vector3 v; for (int a = 0; a < 3; ++a) v.setExact(a, mySizeD[a]); return v; }
    UT_Vector3 getVoxelSize() const { return UT_Vector3(myVoxelSizeD[0], myVoxelSizeD[1], myVoxelSizeD[2]); }
    bool shouldMultiThread() const { return myField->numTiles() > 1; }
    double sampleOffset(int a) const { int on[3]; sampleNodeAxes(mySample, on); return on[a] ? 0.0 : 0.5; }
    bool indexToPos(int x, int y, int z, UT_Vector3 &pos) const {
        const int idx[3] = {x, y, z};
        for (int a = 0; a < 3; ++a) pos.setExact(a, myOrigD[a] + ((double)idx[a] + sampleOffset(a)) * myVoxelSizeD[a]);
        return true;
    }
    bool posToIndex(const UT_Vector3 &pos, UT_Vector3 &index) const {
        for (int a = 0; a < 3; ++a) index.setExact(a, ((double)pos[a] - myOrigD[a]) / myVoxelSizeD[a] - sampleOffset(a));
        return true;
    }
    template <class U>
    bool isAligned(const RawFieldT<U> *o) const {
        if (mySample != o->mySample) return false;
        for (int a = 0; a < 3; ++a) {
            if (myCellRes[a] != o->myCellRes[a]) return false;
            if (std::fabs(myOrigD[a] - o->myOrigD[a]) > 1e-6 * myVoxelSizeD[a]) return false;
            if (std::fabs(mySizeD[a] - o->mySizeD[a]) > 1e-6 * myVoxelSizeD[a] * std::max(1, myCellRes[a])) return false;
        }
        return true;
    }
};
}  // namespace mock_hdk

class SIM_RawIndexField : public mock_hdk::RawFieldT<exint> {};

class SIM_RawField : public mock_hdk::RawFieldT<fpreal32> {
public:
    // trilinear, clamp-to-edge, fp64, a + t (b - a), x then y then z
    fpreal getValue(const UT_Vector3 &pos) const {
        const UT_VoxelArray<fpreal32> &f = *myField;
        int i0[3], i1[3];
        double t[3];
        for (int a = 0; a < 3; ++a) {
            // same arithmetic as the flattened fields of the C-ABI: (p - position of sample 0) / dx
            const double org = myFlatSampling ? myFlatOrg[a] : myOrigD[a] + sampleOffset(a) * myVoxelSizeD[a];
            const double h = myFlatSampling ? myFlatDx : myVoxelSizeD[a];
            double g = ((double)pos[a] - org) / h;
            const double hi = (double)(f.getRes(a) - 1);
            if (g < 0.0) g = 0.0;
            if (g > hi) g = hi;
            const double fl = std::floor(g);
            i0[a] = (int)fl;
            i1[a] = std::min(i0[a] + 1, f.getRes(a) - 1);
            t[a] = g - fl;
        }
        auto V = [&](int x, int y, int z) { return (double)f.getValue(x, y, z); };
        auto lerp = [](double a, double b, double tt) { return a + tt * (b - a); };
        const double c00 = lerp(V(i0[0], i0[1], i0[2]), V(i1[0], i0[1], i0[2]), t[0]);
        const double c10 = lerp(V(i0[0], i1[1], i0[2]), V(i1[0], i1[1], i0[2]), t[0]);
        const double c01 = lerp(V(i0[0], i0[1], i1[2]), V(i1[0], i0[1], i1[2]), t[0]);
        const double c11 = lerp(V(i0[0], i1[1], i1[2]), V(i1[0], i1[1], i1[2]), t[0]);
        const double c0 = lerp(c00, c10, t[1]);
        const double c1 = lerp(c01, c11, t[1]);
        return lerp(c0, c1, t[2]);
    }
    // fraction of the n^3 sub-samples of the voxel-sized box around every sample with (sdf - dilate) < 0
    void computeSDFWeightsSampled(const SIM_RawField *sdf, int samplesperaxis, bool invert, fpreal minweight, fpreal dilate = 0) {
        UT_VoxelArray<fpreal32> &f = *myField;
        const int n = samplesperaxis;
        const double inv = 1.0 / (double)n;
        const double total = (double)n * n * n;
        const int rx = f.getXRes(), ry = f.getYRes(), rz = f.getZRes();
        const UT_VoxelArray<fpreal32> &sf = *sdf->field();
        const bool shortcut = mock_hdk::weightShortcut();
        mock_hdk::runJobs(rz > 1, [&](const UT_JobInfo &info) {
            exint z0, z1;
            info.divideWork(rz, z0, z1);
            for (int z = (int)z0; z < (int)z1; ++z)
                for (int y = 0; y < ry; ++y)
                    for (int x = 0; x < rx; ++x) {
                        UT_Vector3 c;
                        indexToPos(x, y, z, c);
                        if (shortcut) {
                            // exact shortcut: the interpolant is a convex combination of the voxels under the sample's box,
                            // so if they all have one sign every sub-sample has it
                            int lo[3], hi[3];
                            for (int a = 0; a < 3; ++a) {
                                const double hbox = (0.5 - 0.5 * inv) * myVoxelSizeD[a];
                                const double org = sdf->myFlatSampling ? sdf->myFlatOrg[a] : sdf->myOrigD[a] + sdf->sampleOffset(a) * sdf->myVoxelSizeD[a];
                                const double h = sdf->myFlatSampling ? sdf->myFlatDx : sdf->myVoxelSizeD[a];
                                double gl = ((double)c[a] - hbox - org) / h - 1e-9, gh = ((double)c[a] + hbox - org) / h + 1e-9;
                                const double top = (double)(sf.getRes(a) - 1);
                                gl = std::min(std::max(gl, 0.0), top);
                                gh = std::min(std::max(gh, 0.0), top);
                                lo[a] = (int)std::floor(gl);
                                hi[a] = std::min((int)std::floor(gh) + 1, sf.getRes(a) - 1);
                            }
                            bool allNeg = true, allPos = true;
                            for (int kz = lo[2]; kz <= hi[2]; ++kz)
                                for (int ky = lo[1]; ky <= hi[1]; ++ky)
                                    for (int kx = lo[0]; kx <= hi[0]; ++kx) {
                                        const double v = (double)sf.getValue(kx, ky, kz) - dilate;
                                        allNeg = allNeg && (v < 0.0);
                                        allPos = allPos && (v >= 0.0);
                                    }
                            if (allNeg || allPos) {
                                double w = (allNeg != invert) ? 1.0 : 0.0;
                                if (w < minweight) w = minweight;
                                if (w != 0.0) f.setValue(x, y, z, (fpreal32)w);
                                continue;
                            }
                        }
                        int count = 0;
                        for (int k = 0; k < n; ++k)
                            for (int j = 0; j < n; ++j)
                                for (int i = 0; i < n; ++i) {
                                    UT_Vector3 p;
                                    p.setExact(0, (double)c[0] + (((double)i + 0.5) * inv - 0.5) * myVoxelSizeD[0]);
                                    p.setExact(1, (double)c[1] + (((double)j + 0.5) * inv - 0.5) * myVoxelSizeD[1]);
                                    p.setExact(2, (double)c[2] + (((double)k + 0.5) * inv - 0.5) * myVoxelSizeD[2]);
                                    const bool inside = (sdf->getValue(p) - dilate) < 0.0;
                                    if (inside != invert) ++count;
                                }
                        double w = (double)count / total;
                        if (w < minweight) w = minweight;
                        f.setValue(x, y, z, (fpreal32)w);
                    }
        });
        f.collapseAllTiles();
    }
    // this <- scale * this / b where b > threshold
    void setScaleDivideThreshold(fpreal scale, const SIM_RawField *a, const SIM_RawField *b, fpreal threshold) {
        UT_VoxelArray<fpreal32> &f = *myField;
        for (int z = 0; z < f.getZRes(); ++z)
            for (int y = 0; y < f.getYRes(); ++y)
                for (int x = 0; x < f.getXRes(); ++x) {
                    const double bv = b ? (double)b->field()->getValue(x, y, z) : 1.0;
                    if (!(bv > threshold)) continue;
                    const double av = a ? (double)a->field()->getValue(x, y, z) : 1.0;
                    f.setValue(x, y, z, (fpreal32)(scale * (double)f.getValue(x, y, z) * av / bv));
                }
    }
};

class SIM_ScalarField {
public:
    SIM_RawField myField;
    const SIM_RawField *getField() const { return &myField; }
    SIM_RawField *getField() { return &myField; }
};
class SIM_VectorField {
public:
    SIM_RawField myFields[3];
    const SIM_RawField *getField(int axis) const { return &myFields[axis]; }
    SIM_RawField *getField(int axis) { return &myFields[axis]; }
    const SIM_RawField *getXField() const { return &myFields[0]; }
    const SIM_RawField *getYField() const { return &myFields[1]; }
    const SIM_RawField *getZField() const { return &myFields[2]; }
    bool isFaceSampled() const {
        return myFields[0].getSample() == SIM_SAMPLE_FACEX && myFields[1].getSample() == SIM_SAMPLE_FACEY && myFields[2].getSample() == SIM_SAMPLE_FACEZ;
    }
    bool isAligned(const SIM_VectorField *o) const {
        for (int a = 0; a < 3; ++a)
            if (!myFields[a].isAligned(&o->myFields[a])) return false;
        return true;
    }
    UT_Vector3 getVoxelSize() const { return myFields[0].getVoxelSize(); }
    int myModifications = 0;
    void pubHandleModification() { ++myModifications; }
};

// ---- GU / GA -------------------------------------------------------------------------------------
typedef exint GA_Offset;
enum GA_AttributeOwner { GA_ATTRIB_VERTEX, GA_ATTRIB_POINT, GA_ATTRIB_PRIMITIVE, GA_ATTRIB_GLOBAL };
struct GA_Defaults { double v; explicit GA_Defaults(double x) : v(x) {} };
class GU_Detail;
struct GA_Attribute { GU_Detail *detail; std::string name; };
struct GA_AttributeSet { void bumpAllDataIds(GA_AttributeOwner) {} };
class GU_Detail {
public:
    std::vector<UT_Vector3> myPos;
    std::map<std::string, std::vector<double>> myFloat;
    std::map<std::string, std::vector<exint>> myInt;
    GA_AttributeSet myAttribs;
    void clear() { myPos.clear(); myFloat.clear(); myInt.clear(); }
    std::vector<std::unique_ptr<GA_Attribute>> myAttribObjects;
    GA_Attribute *addFloatTuple(GA_AttributeOwner, const char *name, int, const GA_Defaults &) {
        myFloat[name].resize(myPos.size(), 0);
        myAttribObjects.emplace_back(new GA_Attribute{this, name});
        return myAttribObjects.back().get();
    }
    GA_Attribute *addIntTuple(GA_AttributeOwner, const char *name, int, const GA_Defaults &) {
        myInt[name].resize(myPos.size(), -1);
        myAttribObjects.emplace_back(new GA_Attribute{this, name});
        return myAttribObjects.back().get();
    }
    GA_Offset appendPoint() { return appendPointBlock(1); }
    GA_Offset appendPointBlock(exint n) {
        const GA_Offset first = (GA_Offset)myPos.size();
        myPos.resize(myPos.size() + (size_t)n);
        for (auto &kv : myFloat) kv.second.resize(myPos.size(), 0);
        for (auto &kv : myInt) kv.second.resize(myPos.size(), -1);
        return first;
    }
    void setPos3(GA_Offset o, const UT_Vector3 &p) { myPos[(size_t)o] = p; }
    GA_AttributeSet &getAttributes() { return myAttribs; }
};
class GA_RWHandleF {
    GU_Detail *g = nullptr;
    std::string name;

public:
    GA_RWHandleF() {}
    GA_RWHandleF(GU_Detail *gd, GA_AttributeOwner, const char *n) : g(gd), name(n) {}
    GA_RWHandleF(GA_Attribute *a) : g(a ? a->detail : nullptr), name(a ? a->name : "") {}
    bool isValid() const { return g && g->myFloat.count(name); }
    void bumpDataId() {}
    void set(GA_Offset o, double v) { g->myFloat[name][(size_t)o] = v; }
};
class GA_RWHandleI {
    GU_Detail *g = nullptr;
    std::string name;

public:
    GA_RWHandleI() {}
    GA_RWHandleI(GU_Detail *gd, GA_AttributeOwner, const char *n) : g(gd), name(n) {}
    GA_RWHandleI(GA_Attribute *a) : g(a ? a->detail : nullptr), name(a ? a->name : "") {}
    bool isValid() const { return g && g->myInt.count(name); }
    void bumpDataId() {}
    void set(GA_Offset o, exint v) { g->myInt[name][(size_t)o] = v; }
};

// ---- SIM / GAS framework stand-ins ------------------------------------------------------------------
typedef double SIM_Time;
class SIM_Engine {};
class SIM_DataFactory {};
class SIM_GeometryCopy { public: GU_Detail myGdp; };
enum { SIM_DATA_ID_PRESERVE = 0 };
class SIM_GeometryAutoWriteLock {
    SIM_GeometryCopy *g;

public:
    SIM_GeometryAutoWriteLock(SIM_GeometryCopy *geo, int) : g(geo) {}
    GU_Detail &getGdp() { return g->myGdp; }
};
// the DOP object: named fields + the options the GET_DATA_FUNC accessors read
class SIM_Object {
public:
    std::map<std::string, SIM_ScalarField *> scalars;
    std::map<std::string, SIM_VectorField *> vectors;
    std::map<std::string, SIM_GeometryCopy *> geometry;
    std::vector<std::string> errors;
};
enum { SIM_MESSAGE = 0 };
enum { UT_ERROR_WARNING = 1, UT_ERROR_ABORT = 2 };
#define SIM_NAME_TOLERANCE "tolerance"
#define GAS_NAME_SURFACE "surface"
#define GAS_NAME_VELOCITY "velocity"
#define GAS_NAME_DENSITY "density"
#define GAS_NAME_COLLISION "collision"
#define GAS_NAME_COLLISIONVELOCITY "collisionvel"

enum PRM_Type { PRM_STRING, PRM_TOGGLE, PRM_FLT, PRM_INT };
struct PRM_Name { PRM_Name(const char * = nullptr, const char * = nullptr) {} };
struct PRM_Default { PRM_Default(double = 0, const char * = nullptr) {} };
static PRM_Default PRMzeroDefaults[1], PRMoneDefaults[1], PRMtwoDefaults[1], PRMthreeDefaults[1], PRMfourDefaults[1];
struct PRM_Template {
    PRM_Template() {}
    PRM_Template(PRM_Type, int, PRM_Name *, PRM_Default * = nullptr) {}
};
struct SIM_DopDescription {
    SIM_DopDescription(bool, const char *, const char *, const char *, const char *, const PRM_Template *) {}
};

class GAS_SubSolver {
public:
    typedef GAS_SubSolver BaseClassRoot;
    explicit GAS_SubSolver(const SIM_DataFactory *) {}
    virtual ~GAS_SubSolver() {}
    // option table behind GET_DATA_FUNC_* (the DOP parameters)
    std::map<std::string, double> myOptions;
    double getOption(const char *name, double dflt) const {
        auto it = myOptions.find(name);
        return it == myOptions.end() ? dflt : it->second;
    }
    // field lookup: the field NAME is itself a string parameter in Houdini; the harness registers fields under the parameter name
    SIM_ScalarField *getScalarField(SIM_Object *obj, const char *name) const { auto it = obj->scalars.find(name); return it == obj->scalars.end() ? nullptr : it->second; }
    const SIM_ScalarField *getConstScalarField(SIM_Object *obj, const char *name) const { return getScalarField(obj, name); }
    SIM_VectorField *getVectorField(SIM_Object *obj, const char *name) const { auto it = obj->vectors.find(name); return it == obj->vectors.end() ? nullptr : it->second; }
    const SIM_VectorField *getConstVectorField(SIM_Object *obj, const char *name) const { return getVectorField(obj, name); }
    SIM_GeometryCopy *getOrCreateGeometry(SIM_Object *obj, const char *name) const {
        auto it = obj->geometry.find(name);
        if (it != obj->geometry.end()) return it->second;
        SIM_GeometryCopy *g = new SIM_GeometryCopy();
        obj->geometry[name] = g;
        return g;
    }
    void addError(SIM_Object *obj, int, const char *text, int) const { if (obj) obj->errors.push_back(text); }
    static void setGasDescription(SIM_DopDescription &) {}
    virtual bool solveGasSubclass(SIM_Engine &, SIM_Object *, SIM_Time, SIM_Time) = 0;
};
#define GET_DATA_FUNC_F(NAME, FN) fpreal get##FN() const { return (fpreal)getOption(NAME, 0.0); }
#define GET_DATA_FUNC_I(NAME, FN) int get##FN() const { return (int)getOption(NAME, 0.0); }
#define GET_DATA_FUNC_B(NAME, FN) bool get##FN() const { return getOption(NAME, 0.0) != 0.0; }
#define DECLARE_STANDARD_GETCASTTOTYPE()
#define DECLARE_DATAFACTORY(CLASS, BASE, DESC, DOPDESC) \
public:                                                  \
    typedef BASE BaseClass;                              \
    static const char *classname() { return #CLASS; }    \
    static CLASS *mockCreate() { return new CLASS(nullptr); } \
    static const SIM_DopDescription *mockDopDescription() { return DOPDESC; } \
private:
#define IMPLEMENT_DATAFACTORY(CLASS)

// ---- UT_SparseMatrix (only named by the non-Eigen branch of HDK_Utilities.h; the library is built with USEEIGEN) ----------
template <class T, bool B> class UT_SparseMatrixT {};
template <class T, bool B> class UT_SparseMatrixELLT {};
template <class T> class UT_SparseMatrixRowT {};
template <class T> class UT_VectorT {};
