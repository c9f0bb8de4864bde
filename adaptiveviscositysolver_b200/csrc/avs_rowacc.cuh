// avs_rowacc.cuh -- per-thread accumulator of one matrix row during assembly (setFromTriplets semantics: duplicates are summed,
// HDK_AdaptiveViscosity.cpp:614).  Plain C++ so that tests/test_rowacc.py can compile it for the host.
#pragma once
#include <stdint.h>

#ifndef AVS_HD
#if defined(__CUDACC__) && defined(AVS_HOST_TEST)
#define AVS_HD __host__ __device__ __forceinline__   // tests/test_host_assembly.py: the row builder compiled for the host as well
#elif defined(__CUDACC__)
#define AVS_HD __device__ __forceinline__
#else
#define AVS_HD inline
#endif
#endif

// Longest row: a coarse face whose whole neighbourhood is one level finer couples to 64 fine faces + itself = 65 entries (32 fine
// faces of its own axis, 16 of each transverse axis).  Sphere drops never produce one (their rows end at 46 entries); a droplet a
// few cells across, or liquid folded over a solid, does: 18 of the first 300 scenes of scripts/fuzz_reference_pin.py hold a 65-entry
// row, none a longer one.  Only the entries a row really has are ever touched in local memory, so the capacity costs nothing there;
// the staging area of the assembly is MAX_ROW deep (avs_stage_system).
#define MAX_ROW 80

// Linear search over the entries collected so far (rows have 2..65 entries, 16.7 on average).
struct RowAcc {
    int n;
    int overflow;
    int32_t col[MAX_ROW];
    double val[MAX_ROW];
    AVS_HD void init() { n = 0; overflow = 0; }
    AVS_HD void add(int32_t c, double v) {
        for (int i = 0; i < n; ++i)
            if (col[i] == c) { val[i] += v; return; }  // setFromTriplets sums duplicates (AV.cpp:614)
        if (n < MAX_ROW) { col[n] = c; val[n] = v; ++n; }
        else overflow = 1;
    }
};

// Default (AVS_ASM_ROW=linear selects the one above): same entries in the same insertion order -- hence bit-identical rows -- but
// the search goes through a 128-slot open-addressing table of entry indices (load factor <= 0.625): ~1.3 probes per add instead of
// n/2 compares.
// ncu on the default (profiles/r1_experiments.md): 39 % of the assembly kernel's stall samples sit on the compares of the linear
// search, 1.0e9 of its 8.0e9 warp instructions are those compares, another 1.1e9 their branches.
struct RowAccHash {
    int n;
    int overflow;
    int32_t col[MAX_ROW];
    double val[MAX_ROW];
    int8_t slot[128];
    AVS_HD void init() {
        n = 0;
        overflow = 0;
        for (int i = 0; i < 128; ++i) slot[i] = -1;
    }
    AVS_HD void add(int32_t c, double v) {
        unsigned h = ((unsigned)c * 2654435761u) >> 25;   // 7 bits
        for (int probe = 0; probe < 128; ++probe) {
            const int s = slot[h];
            if (s < 0) {
                if (n < MAX_ROW) { slot[h] = (int8_t)n; col[n] = c; val[n] = v; ++n; }
                else overflow = 1;
                return;
            }
            if (col[s] == c) { val[s] += v; return; }
            h = (h + 1) & 127u;
        }
        overflow = 1;   // table full: cannot happen with MAX_ROW = 80 < 128
    }
};
