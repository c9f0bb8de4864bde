// tests/host_assembly.cu -- TEST INFRASTRUCTURE (built and used only by tests/test_host_assembly.py; never part of libavs_b200.so).
//
// The row builder of the CUDA library -- buildRow / buildSimpleRow and everything they call in csrc/avs_system.cu and
// csrc/avs_common.cuh: edgeStressFaces, centerStressFaces, the control volumes, the stress weights, applyToMatrix, the row
// accumulators -- compiled for the HOST.  The product marks that code AVS_DEV (= __device__); with -DAVS_HOST_TEST it becomes
// __host__ __device__, and this file, which includes avs_system.cu as a whole, runs it on the CPU over label grids that come from a
// run of the REFERENCE'S OWN CODE (oracle/_ref/libavs_ref.so).  The rows it returns are then compared with the reference's matrix:
// the product's assembly source against the reference without a GPU and without the restated oracle in between.
//
// What is NOT the product's code here: the loop over the rows (the kernels' one-thread-per-row mapping), the fetch of a row's face
// weight (k_gather_face_weights, one line) and the staging / SJDS layout -- those only exist on the GPU (tests -m gpu).
#ifndef AVS_HOST_TEST
#error "compile with -DAVS_HOST_TEST"
#endif
#include "../adaptiveviscositysolver_b200/csrc/avs_system.cu"

#include "host_scene.h"

template <class Row>
static int genericRow(const DeviceScene &S, int32_t vi, const RowKey &k, float fw, int32_t *col, double *val, double &rhsI, double &mass) {
    Row row;
    row.init();
    rhsI = 0;
    mass = 0;
    buildRow(S, vi, k, fw, row, rhsI, mass);
    if (row.overflow) return -1;
    for (int i = 0; i < row.n; ++i) { col[i] = row.col[i]; val[i] = row.val[i]; }
    return row.n;
}

extern "C" {

int host_max_row(void) { return MAX_ROW; }

// mode 0: every row through buildRow with the hashed accumulator (k_assemble<.., RowAccHash>, the default second pass)
// mode 1: every row through buildRow with the linear accumulator (AVS_ASM_ROW=linear)
// mode 2: the split assembly: buildSimpleRow where it accepts the row (k_assemble_simple), buildRow for the rest
// Entries are returned in the builder's own order, MAX_ROW slots per row: col[r * MAX_ROW + i], val[...]; count[r] = entries,
// simple[r] = 1 where buildSimpleRow wrote the row.  Returns 0, or -(row + 1) for a row that overflowed.
long long host_assemble_rows(const HostSceneDesc *h, long long nRows, const int32_t *keys, int mode, int32_t *count, int32_t *col, double *val,
                             double *rhsI, double *mass, int32_t *simple) {
    DeviceScene S;
    fillScene(*h, S);
    for (long long r = 0; r < nRows; ++r) {
        RowKey k;
        k.level = keys[5 * r + 0]; k.axis = keys[5 * r + 1]; k.i = keys[5 * r + 2]; k.j = keys[5 * r + 3]; k.k = keys[5 * r + 4];
        const float fw = (k.level == 0) ? S.faceW[k.axis].raw(k.i, k.j, k.k) : 1.f;   // k_gather_face_weights
        int32_t *c = col + r * MAX_ROW;
        double *v = val + r * MAX_ROW;
        simple[r] = 0;
        int n = -1;
        if (mode == 2 && k.level == 0) {
            // the staging area of k_assemble_simple is column-major with a stride; here stride = 1 makes it this row's slots
            SimpleRowOut out;
            out.col = c;
            out.val = v;
            out.stride = 1;
            out.last = 0;
            out.n = 0;
            double m = 0;
            const I3 face = mk3(k.i, k.j, k.k);
            bool ok;
            if (k.axis == 0) ok = buildSimpleRow<0>(S, (int32_t)r, face, fw, out, m);
            else if (k.axis == 1) ok = buildSimpleRow<1>(S, (int32_t)r, face, fw, out, m);
            else ok = buildSimpleRow<2>(S, (int32_t)r, face, fw, out, m);
            if (ok) {
                n = out.n;
                rhsI[r] = 0;
                mass[r] = m;
                simple[r] = 1;
            }
        }
        if (n < 0) {
            n = (mode == 1) ? genericRow<RowAcc>(S, (int32_t)r, k, fw, c, v, rhsI[r], mass[r])
                            : genericRow<RowAccHash>(S, (int32_t)r, k, fw, c, v, rhsI[r], mass[r]);
            if (n < 0) return -(r + 1);
        }
        count[r] = n;
    }
    return 0;
}

// Stage 8, restriction of u^n to the octree faces (AV.cpp:2291-2402) with the library's leaf / hat-weight functions.  The loops
// are those of k_restrict_fine (levels 0-2, same order of additions) and, for levels >= 3, the plain sum of the terms that
// k_restrict_coarse / k_restrict_coarse_finish add CTA-wise (same terms, another order: <= 1e-13 relative).
void host_restrict_rows(const HostSceneDesc *h, long long nRows, const int32_t *keys, double *x0) {
    DeviceScene S;
    fillScene(*h, S);
    for (long long r = 0; r < nRows; ++r) {
        RowKey k;
        k.level = keys[5 * r + 0]; k.axis = keys[5 * r + 1]; k.i = keys[5 * r + 2]; k.j = keys[5 * r + 3]; k.k = keys[5 * r + 4];
        const I3 face = mk3(k.i, k.j, k.k);
        double acc = 0;
        if (k.level == 0) {
            x0[r] = 1.0 * (double)S.vel[k.axis].raw(face[0], face[1], face[2]);
        } else if (k.level == 1) {
            for (int q = 0; q < 12; ++q) {
                double w;
                double v = restrictLeaf(S, k.axis, 1, face, q, w);
                acc += w * v;
            }
            x0[r] = acc;
        } else if (k.level == 2) {
            for (int q = 0; q < 7 * 16; ++q) acc += restrictHatTerm(S, k.axis, 2, face, q);
            x0[r] = acc * (1.0 / 256.0);
        } else {
            const long long side = 1ll << k.level, terms = (2 * side - 1) * side * side;
            for (long long q = 0; q < terms; ++q) acc += restrictHatTerm(S, k.axis, k.level, face, q);
            x0[r] = acc / (double)(side * side * side * side);
        }
    }
}

}  // extern "C"
