// avs_system.cu -- stages 6-9 and 11 on the GPU:
//   stress stencils + control volumes (HDK_AdaptiveViscosity.cpp:1717-2289), evaluated on the fly
//   restriction of u^n to the octree faces (HDK_AdaptiveViscosity.cpp:2291-2402)
//   SPD assembly  A = M_u + sum_s w_s d_s d_s^T,  rhs  (HDK_AdaptiveViscosity.cpp:2404-2777)
//   scatter of the solution to the regular grid (HDK_AdaptiveViscosity.cpp:2779-2894, level-0 part)
//
// One thread owns one matrix row (one velocity face).  The reference materialises every row of D in
// heap arrays (AV.cpp:429-436) and merges per-thread triplet lists serially (AV.cpp:587-593, 614);
// here a row of D is a pure function of the flattened label/index grids, so each row thread
// re-evaluates the <= ~12 stencils that touch it, merges duplicates in registers/local memory and
// writes its CSR segment directly (count pass -> scan -> fill pass).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include "avs_context.h"
#include "avs_rowacc.cuh"

#define MAX_STENCIL 40

struct Stencil {
    int n, nb;
    int32_t idx[MAX_STENCIL];
    double coef[MAX_STENCIL];
    double bnd[8];
    AVS_DEV __forceinline__ void add(int32_t i, double c) {
        if (n < MAX_STENCIL) { idx[n] = i; coef[n] = c; ++n; }
    }
    AVS_DEV __forceinline__ void addB(double b) {
        if (nb < 8) bnd[nb++] = b;
    }
};

// getEdgeStressFaces (AV.cpp:1717-1908)
AVS_DEV __noinline__ void edgeStressFaces(const DeviceScene &S, Stencil &st, const I3 &edge, int axis, int level) {
    st.n = st.nb = 0;
    const double dx = S.levelDx(level);  // AV.cpp:1733
    bool isAtTransition[3] = {false, false, false};
    bool isFaceOutside[3] = {false, false, false};
    float gradientDx[3] = {0.f, 0.f, 0.f};  // UT_Vector3 (float32), AV.cpp:1738
    for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
        if (faceAxis == axis) continue;
        const Grid3<int32_t> &fg = S.face[level][faceAxis];
        for (int dir = 0; dir < 2; ++dir) {
            I3 face = edgeToFace(edge, axis, faceAxis, dir);
            const int g = 3 - faceAxis - axis;
            if (face[g] < 0 || face[g] >= fg.n[g]) {
                gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                isFaceOutside[g] = true;
            } else {
                int32_t vi = fg.get(face);
                if (vi >= 0) gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                else if (vi == F_OUTSIDE || vi == F_SOLID) {
                    gradientDx[g] = (float)((double)gradientDx[g] + .5 * dx);
                    isFaceOutside[g] = true;
                } else {  // UNASSIGNED: a coarser face sits there (AV.cpp:1777-1782)
                    gradientDx[g] = (float)((double)gradientDx[g] + dx);
                    if (S.enhanced) isAtTransition[g] = true;
                }
            }
        }
    }
    for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
        if (faceAxis == axis) continue;
        const Grid3<int32_t> &fg = S.face[level][faceAxis];
        for (int dir = 0; dir < 2; ++dir) {
            I3 face = edgeToFace(edge, axis, faceAxis, dir);
            const int g = 3 - faceAxis - axis;
            const double sign = (dir == 0) ? -1 : 1;
            const double gdx = (double)gradientDx[g];
            if (face[g] < 0 || face[g] >= fg.n[g]) continue;
            int32_t vi = fg.get(face);
            if (vi >= 0) {
                if (isAtTransition[g] && !isFaceOutside[g]) {  // AV.cpp:1814-1824
                    I3 sib = face;
                    sib[axis] += (edge[axis] % 2 == 0) ? 1 : -1;
                    int32_t si = fg.get(sib);
                    st.add(si, .25 * sign / gdx);
                    st.add(vi, .25 * sign / gdx);
                } else
                    st.add(vi, .5 * sign / gdx);
            } else if (vi == F_UNASSIGNED) {
                if (edge[faceAxis] % 2 != 0) {  // dangling edge, AV.cpp:1835-1884
                    for (int off = -1; off <= 1; off += 2) {
                        I3 of = face;
                        of[faceAxis] += off;
                        I3 pf = parentOf(of);
                        int32_t pi = (level + 1 < S.levels) ? S.face[level + 1][faceAxis].get(pf) : F_OUTSIDE;
                        if (pi >= 0) st.add(pi, .25 * sign / gdx);
                        else if (pi == F_UNASSIGNED) {
                            for (int ch = 0; ch < 4; ++ch) {
                                int32_t ci = fg.get(childFace(pf, faceAxis, ch));
                                if (ci >= 0) st.add(ci, .0625 * sign / gdx);
                            }
                        }
                    }
                } else {  // AV.cpp:1886-1894
                    I3 pf = parentOf(face);
                    int32_t pi = (level + 1 < S.levels) ? S.face[level + 1][faceAxis].get(pf) : F_OUTSIDE;
                    st.add(pi, .5 * sign / gdx);
                }
            } else if (vi == F_SOLID) {  // AV.cpp:1896-1905: component `axis`, as written in the reference
                double p[3];
                S.facePos(face, faceAxis, level, p);
                double lv = S.collisionVel[axis].value(p);
                st.addB(.5 * sign * lv / gdx);
            }
        }
    }
}

// getCenterStressFaces (AV.cpp:1910-1963)
AVS_DEV __noinline__ void centerStressFaces(const DeviceScene &S, Stencil &st, const I3 &cell, int axis, int level) {
    st.n = st.nb = 0;
    const double dx = S.levelDx(level);
    for (int dir = 0; dir < 2; ++dir) {
        I3 face = cellToFace(cell, axis, dir);
        const double sign = (dir == 0) ? -1 : 1;
        int32_t vi = S.face[level][axis].get(face);
        if (vi >= 0) st.add(vi, sign / dx);
        else if (vi == F_UNASSIGNED) {
            if (level > 0)
                for (int ch = 0; ch < 4; ++ch) st.add(S.face[level - 1][axis].get(childFace(face, axis, ch)), .25 * sign / dx);
        } else if (vi == F_SOLID) {
            double p[3];
            S.facePos(face, axis, level, p);
            st.addB(sign * S.collisionVel[axis].value(p) / dx);
        }
    }
}

// faceOctreeVolumes (AV.cpp:1965-2002)
AVS_DEV double faceOctreeVolume(const DeviceScene &S, const I3 &face, int axis, int level) {
    const Grid3<uint8_t> &lab = S.label[level];
    const double dx = (double)(1 << level);
    double g = 0;
    for (int dir = 0; dir < 2; ++dir) {
        I3 c = faceToCell(face, axis, dir);
        if (c[axis] < 0 || c[axis] >= lab.n[axis]) g += .5 * dx;
        else {
            uint8_t l = lab.get(c);
            g += (l == L_ACTIVE || l == L_INACTIVE) ? .5 * dx : dx;
        }
    }
    return dx * dx * g;
}

// edgeOctreeVolumes (AV.cpp:2004-2057)
AVS_DEV double edgeOctreeVolume(const DeviceScene &S, const I3 &edge, int axis, int level) {
    const double dx = (double)(1 << level);
    float v[3] = {0.f, 0.f, 0.f};
    v[axis] = (float)dx;
    for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
        if (faceAxis == axis) continue;
        const Grid3<int32_t> &fg = S.face[level][faceAxis];
        for (int dir = 0; dir < 2; ++dir) {
            I3 face = edgeToFace(edge, axis, faceAxis, dir);
            const int g = 3 - faceAxis - axis;
            if (face[g] < 0 || face[g] >= fg.n[g]) v[g] = (float)((double)v[g] + .5 * dx);
            else {
                int32_t vi = fg.get(face);
                v[g] = (float)((double)v[g] + ((vi == F_UNASSIGNED) ? dx : .5 * dx));
            }
        }
    }
    return (double)(v[0] * v[1] * v[2]);
}

// stress weights: AV.cpp:2124-2155 (edges), AV.cpp:2223-2289 (centres)
AVS_DEV __noinline__ double edgeStressWeight(const DeviceScene &S, const I3 &edge, int axis, int level) {
    double w;
    if (level == 0) {
        w = (double)S.edgeW[axis].get(edge);
        if (w == 1.) w = edgeOctreeVolume(S, edge, axis, level);
    } else w = edgeOctreeVolume(S, edge, axis, level);
    if (!S.viscosity.d) w *= (double)S.viscosity.constant;
    else {
        double p[3];
        S.edgePos(edge, axis, level, p);
        w *= S.viscosity.value(p);
    }
    return 4. * S.dt * w;
}
AVS_DEV double centerStressWeight(const DeviceScene &S, const I3 &cell, int level) {
    double w;
    if (level == 0) w = (double)S.centerW.get(cell);
    else { double dx = (double)(1 << level); w = dx * dx * dx; }
    if (!S.viscosity.d) w *= (double)S.viscosity.constant;
    else {
        double p[3];
        S.centerPos(cell, level, p);
        w *= S.viscosity.value(p);
    }
    return 2. * S.dt * w;
}

// ------------------------------------------------------------------------------------------------
// Stage 8: buildVelocityMapping (AV.cpp:2291-2402).  Level 0 copies the regular face value; level l
// sums 12^l leaves with weights (1/16, 1/8, 1/16) per child and in-axis offset.
AVS_DEV __forceinline__ double restrictLeaf(const DeviceScene &S, int axis, int level, I3 face, long long q, double &wOut) {
    // q enumerates the 12^level leaves in depth-first order (base-12 digits, most significant first)
    double w = 1.0;
    long long div = 1;
    for (int d = 1; d < level; ++d) div *= 12;
    for (int d = 0; d < level; ++d) {
        int digit = (int)(q / div);
        q -= (long long)digit * div;
        div /= 12;
        int ch = digit / 3, o = digit % 3 - 1;
        face = childFace(face, axis, ch);
        face[axis] += o;
        w = (double)(float)(((o == 0) ? 0.125 : 0.0625) * w);  // weight kept as fpreal32 (AV.cpp:2318)
    }
    wOut = w;
    return (double)S.vel[axis].raw(face[0], face[1], face[2]);
}

// Closed form for level >= 2 (same leaves, regrouped): the in-axis offsets of the 12^l leaves are
// D = sum_m 2^(l-m) o_m with o_m in {-1,0,1} weighted (1/4,1/2,1/4), i.e. the difference of two uniform
// l-bit integers, so the weight of in-axis offset d is the hat (2^l - |d|) / 4^l; transverse positions are
// the 2^l x 2^l block, each weighted 4^-l.  (2^(l+1)-1) * 4^l distinct reads instead of 12^l; all weights are
// dyadic, so only the order of the fp64 additions differs from the reference's flat sum (<= 1e-13 relative).
AVS_DEV __forceinline__ double restrictHatTerm(const DeviceScene &S, int axis, int level, const I3 &face, long long q) {
    const int side = 1 << level;
    const int t1 = (int)(q % side);
    const int t2 = (int)((q / side) % side);
    const int d = (int)(q / ((long long)side * side)) - (side - 1);
    I3 f;
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    f[axis] = face[axis] * side + d;
    f[a1] = face[a1] * side + t1;
    f[a2] = face[a2] * side + t2;
    const double w = (double)(side - (d < 0 ? -d : d));
    return w * (double)S.vel[axis].raw(f[0], f[1], f[2]);
}

// levels 0 and 1: leaves added in the reference's order (bit-identical to the oracle); level 2: hat form
__global__ void k_restrict_fine(const __grid_constant__ DeviceScene S, const RowKey *keys, long long base, long long n, double *x0) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i += base;
    RowKey k = keys[i];
    if (k.level > 2) return;
    I3 face = mk3(k.i, k.j, k.k);
    if (k.level == 0) {
        x0[i] = 1.0 * (double)S.vel[k.axis].raw(face[0], face[1], face[2]);
        return;
    }
    double acc = 0;
    if (k.level == 1) {
        for (int q = 0; q < 12; ++q) {
            double w;
            double v = restrictLeaf(S, k.axis, 1, face, q, w);
            acc += w * v;
        }
        x0[i] = acc;
        return;
    }
    for (int q = 0; q < 7 * 16; ++q) acc += restrictHatTerm(S, k.axis, 2, face, q);
    x0[i] = acc * (1.0 / 256.0);  // 16^-l
}
// levels >= 3: a row's (2^(l+1) - 1) 4^l terms are split into restrictSplit(l) work items of <= ~16 K terms, one CTA each -- one
// CTA per ROW left the few top-level rows (4.2 M terms each at level 7) as a serial tail longer than all other rows together.
// k_collect_coarse_rows (two passes: count, then fill) lists the rows and their work items; every CTA reduces its strided share in
// a fixed order into partial[item]; k_restrict_coarse_finish adds a row's parts in ascending order: deterministic, the same on
// every rank count.
#define RESTRICT_MAX_SPLIT 256
__host__ __device__ __forceinline__ int restrictSplit(int level) {
    const long long side = 1ll << level, terms = (2 * side - 1) * side * side;
    const long long s = (terms + 16383) / 16384;
    return (int)(s < 1 ? 1 : (s > RESTRICT_MAX_SPLIT ? RESTRICT_MAX_SPLIT : s));
}
// counters[0] = coarse rows, counters[1] = work items.  fill == 0: count only.
__global__ void k_collect_coarse_rows(const RowKey *keys, long long base, long long n, int fill, int32_t *rows, long long *rowItemBase, int2 *items,
                                      unsigned long long *counters) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i += base;
    const int level = keys[i].level;
    if (level < 3) return;
    const int split = restrictSplit(level);
    const unsigned long long slot = atomicAdd(&counters[0], 1ull);           // order is irrelevant: every row is written exactly once
    const unsigned long long first = atomicAdd(&counters[1], (unsigned long long)split);
    if (!fill) return;
    rows[slot] = (int32_t)i;
    rowItemBase[slot] = (long long)first;
    for (int p = 0; p < split; ++p) items[first + p] = make_int2((int)slot, p);
}
__global__ void k_restrict_coarse(const __grid_constant__ DeviceScene S, const RowKey *keys, const int32_t *rows, const int2 *items, double *partial) {
    const int2 item = items[blockIdx.x];
    const long long i = rows[item.x];
    const RowKey k = keys[i];
    const int split = restrictSplit(k.level), part = item.y;
    const I3 face = mk3(k.i, k.j, k.k);
    const long long side = 1ll << k.level;
    const long long terms = (2 * side - 1) * side * side;
    double acc = 0;
    for (long long q = (long long)part * blockDim.x + threadIdx.x; q < terms; q += (long long)blockDim.x * split) acc += restrictHatTerm(S, k.axis, k.level, face, q);
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_restrict_coarse_finish(const RowKey *keys, const int32_t *rows, const long long *rowItemBase, unsigned long long nCoarse, const double *partial,
                                         double *x0) {
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nCoarse) return;
    const long long i = rows[r];
    const int level = keys[i].level, split = restrictSplit(level);
    const double *part = partial + rowItemBase[r];
    double acc = 0;
    for (int p = 0; p < split; ++p) acc += part[p];
    const long long side = 1ll << level;
    x0[i] = acc / (double)(side * side * side * side);  // 16^-l, exact power of two
}

// Multi-GPU: a rank restricts only the rows it owns -- the assembly reads x0 of its own rows (rhs += M u^n) and the CG stages its
// local x0 in p's buffer and fetches the halo entries from the peers (avs_cg.cu) -- unless `allRows` (read-back of the whole x0).
int avs_stage_restriction(AvsContext *c, bool allRows) {
    const long long nAll = c->nRows;
    if (c->x0.reserve((size_t)std::max<long long>(nAll, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    // (rowBegin / rowEnd are assigned by avs_stage_system, after this stage: take the block from the numbering's row partition)
    const bool all = allRows || c->nranks == 1 || c->rowStarts.size() != (size_t)c->nranks + 1;
    const long long base = all ? 0 : c->rowStarts[c->rank];
    const long long n = all ? nAll : c->rowStarts[c->rank + 1] - c->rowStarts[c->rank];
    c->x0AllRows = all;
    if (n <= 0) return AVS_OK;
    unsigned blocks = (unsigned)((n + 255) / 256);
    k_restrict_fine<<<blocks, 256, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), base, n, c->x0.as<double>());
    ++c->launches;
    if (c->S.levels > 3) {
        // rows of level >= 3 are few (a few percent): count them and their work items, then list them, then one CTA per item
        unsigned long long *cnt = c->counters.as<unsigned long long>() + 40;   // slots 40, 41
        unsigned long long h[2] = {0, 0};
        AVS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), c->stream));
        k_collect_coarse_rows<<<blocks, 256, 0, c->stream>>>(c->rowKeys.as<RowKey>(), base, n, 0, nullptr, nullptr, nullptr, cnt);
        ++c->launches;
        AVS_CUDA_CHECK(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (h[0] > 0) {
            // one buffer: rows int32[h0] | item base int64[h0] | items int2[h1] | partial double[h1]
            const size_t oBase = (h[0] * 4 + 255) / 256 * 256, oItems = oBase + (h[0] * 8 + 255) / 256 * 256, oPart = oItems + (h[1] * 8 + 255) / 256 * 256;
            if (c->coarseRows.reserve(oPart + h[1] * 8)) return AVS_ERR_ALLOC;
            char *buf = c->coarseRows.as<char>();
            int32_t *rows = (int32_t *)buf;
            long long *rowBase = (long long *)(buf + oBase);
            int2 *items = (int2 *)(buf + oItems);
            double *partial = (double *)(buf + oPart);
            AVS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), c->stream));
            k_collect_coarse_rows<<<blocks, 256, 0, c->stream>>>(c->rowKeys.as<RowKey>(), base, n, 1, rows, rowBase, items, cnt);
            k_restrict_coarse<<<(unsigned)h[1], 256, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), rows, items, partial);
            k_restrict_coarse_finish<<<(unsigned)((h[0] + 255) / 256), 256, 0, c->stream>>>(c->rowKeys.as<RowKey>(), rows, rowBase, h[0], partial, c->x0.as<double>());
            c->launches += 3;
        }
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Stage 9: one matrix row (accumulators: avs_rowacc.cuh).

// applyToMatrix (AV.cpp:2404-2457)
template <class Row>
AVS_DEV __noinline__ void applyToMatrix(Row &row, double &rhsI, double &diag, double coefficient, int32_t vi, const Stencil &st) {
    for (int i = 0; i < st.n; ++i)
        if (st.idx[i] == vi) { coefficient *= st.coef[i]; break; }
    for (int i = 0; i < st.n; ++i) {
        double el = coefficient * st.coef[i];
        if (st.idx[i] == vi) diag += el;
        else row.add(st.idx[i], el);
    }
    for (int i = 0; i < st.nb; ++i) rhsI -= coefficient * st.bnd[i];
}

// buildOctreeSystemFromStencilsPartial (AV.cpp:2459-2777) for row vi
template <class Row>
AVS_DEV void buildRow(const DeviceScene &S, int32_t vi, const RowKey &k, float faceWeight, Row &row, double &rhsI, double &massOut) {
    const int level = k.level, axis = k.axis;
    const I3 face = mk3(k.i, k.j, k.k);
    const Grid3<uint8_t> &lab = S.label[level];
    const Grid3<int32_t> &fgrid = S.face[level][axis];
    double diag = 0;
    Stencil st;
    for (int dir = 0; dir < 2; ++dir) {
        I3 cell = faceToCell(face, axis, dir);
        if (cell[axis] < 0 || cell[axis] >= lab.n[axis]) continue;
        I3 sc;
        int sl;
        if (lab.get(cell) == L_ACTIVE) { sc = cell; sl = level; }
        else { sc = parentOf(cell); sl = level + 1; }  // face grading: the parent is ACTIVE (AV.cpp:2562-2572)
        if (sl >= S.levels) continue;
        if (S.center[sl].get(sc) >= 0) {
            centerStressFaces(S, st, sc, axis, sl);
            applyToMatrix(row, rhsI, diag, centerStressWeight(S, sc, sl), vi, st);
        }
        // T-junction ghost stresses (AV.cpp:2614-2649)
        for (int fa = 0; fa < 3; ++fa) {
            if (fa == axis) continue;
            for (int fd = 0; fd < 2; ++fd) {
                I3 af = cellToFace(sc, fa, fd);
                if (sl > 0 && S.face[sl][fa].get(af) == F_UNASSIGNED) {
                    const int ea = 3 - fa - axis;
                    for (int ins = 0; ins < 2; ++ins) {
                        I3 e = childEdgeInFace(af, fa, ea, ins);
                        if (S.edge[sl - 1][ea].get(e) >= 0) {
                            edgeStressFaces(S, st, e, ea, sl - 1);
                            applyToMatrix(row, rhsI, diag, edgeStressWeight(S, e, ea, sl - 1), vi, st);
                        }
                    }
                }
            }
        }
    }
    for (int ea = 0; ea < 3; ++ea) {
        if (ea == axis) continue;
        const Grid3<int8_t> &eg = S.edge[level][ea];
        for (int dir = 0; dir < 2; ++dir) {
            I3 e = faceToEdge(face, axis, ea, dir);
            int8_t ei = eg.get(e);
            if (ei >= 0) {
                if (S.enhanced) {  // AV.cpp:2664-2697
                    const int ta = 3 - ea - axis;
                    I3 af = face;
                    af[ta] += (dir == 0) ? -1 : 1;
                    if (af[ta] >= 0 && af[ta] < fgrid.n[ta] && fgrid.get(af) == F_UNASSIGNED) {
                        I3 se = e;
                        se[ea] += (e[ea] % 2 == 0) ? 1 : -1;
                        if (eg.get(se) >= 0) {
                            edgeStressFaces(S, st, se, ea, level);
                            applyToMatrix(row, rhsI, diag, edgeStressWeight(S, se, ea, level), vi, st);
                        }
                    }
                }
                edgeStressFaces(S, st, e, ea, level);
                applyToMatrix(row, rhsI, diag, edgeStressWeight(S, e, ea, level), vi, st);
            } else if (ei == F_UNASSIGNED && level > 0) {  // AV.cpp:2714-2742
                for (int ch = 0; ch < 2; ++ch) {
                    I3 ce = childEdge(e, ea, ch);
                    if (S.edge[level - 1][ea].get(ce) >= 0) {
                        edgeStressFaces(S, st, ce, ea, level - 1);
                        applyToMatrix(row, rhsI, diag, edgeStressWeight(S, ce, ea, level - 1), vi, st);
                    }
                }
            }
        }
    }
    // velocity control volume (AV.cpp:2748-2772)
    double fw;
    if (level == 0) {
        fw = (double)faceWeight;   // S.faceW[axis] at the face, fetched by k_gather_face_weights
        if (fw == 1.) fw = faceOctreeVolume(S, face, axis, level);
    } else fw = faceOctreeVolume(S, face, axis, level);
    if (!S.density.d) fw *= (double)S.density.constant;
    else {
        double p[3];
        S.facePos(face, axis, level, p);
        fw *= S.density.value(p);
    }
    row.add(vi, fw + diag);
    massOut = fw;   // rhs_i += M_u u^n_i (AV.cpp:2767-2772) is added by k_finish_rhs once the restricted velocity exists
}

// The face-weight field ("surfaceweights", AV.cpp:144) is read at ONE place of the whole path: the control volume of a LEVEL-0
// row (AV.cpp:2748-2760).  So the three dense arrays (1.6 GB at 512^3) need not be resident: this kernel fetches the one value each
// level-0 row of this rank needs into a per-row array.  S.faceW[a].d is either a device copy / the caller's device array, or --
// when the caller's host arrays are pinned -- the mapped HOST pointer, and the reads then go over PCIe: ~N_level0 x 32-byte
// sectors instead of the bulk upload (avs_stage_upload).
__global__ void k_gather_face_weights(const __grid_constant__ DeviceScene S, const RowKey *keys, long long base, long long n, float *fwRow) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i += base;
    const RowKey k = keys[i];
    fwRow[i] = (k.level == 0) ? S.faceW[k.axis].raw(k.i, k.j, k.k) : 1.f;
}

// One pass: every row thread writes its merged entries into a column-major staging area
// (entry j of local row r at j*stride + r, so a warp's writes are coalesced), its entry count, its
// rhs and its diagonal.  The CG's SJDS matrix is filled straight from the staging area; a canonical
// CSR (sorted columns) is only built when a caller asks to read the system back.
template <int MINB, class Row>
__global__ void __launch_bounds__(128, MINB) k_assemble(const __grid_constant__ DeviceScene S, const RowKey *keys, double *rowMass, const float *fwRow,
                                                  long long rowBegin, long long nLocal, long long stride, int32_t *rowCount,
                                                  int32_t *stageCol, double *stageVal, double *rhs, double *diagOut,
                                                  int *overflowFlag, const int32_t *rowList, const unsigned long long *rowListCount) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (rowList) {   // second pass of the split assembly: only the rows k_assemble_simple passed on
        if ((unsigned long long)r >= *rowListCount) return;
        r = rowList[r];
    }
    if (r >= nLocal) return;
    const long long gi = rowBegin + r;
    Row row;
    row.init();
    double rhsI = 0, mass = 0;
    buildRow(S, (int32_t)gi, keys[gi], fwRow[gi], row, rhsI, mass);
    rowMass[r] = mass;
    if (row.overflow) *overflowFlag = 1;
    rowCount[r] = row.n;
    for (int i = 0; i < row.n; ++i) {
        stageCol[(long long)i * stride + r] = row.col[i];
        stageVal[(long long)i * stride + r] = row.val[i];
    }
    rhs[r] = rhsI;
    diagOut[r] = row.val[row.n - 1];  // buildRow adds the diagonal last and it is unique (AV.cpp:2768)
}

// ---- split assembly, first pass (default): the level-0 rows without T-junctions, solids or coarse neighbours -------------------
// buildRow's control flow (which stencils exist, what each one touches) is what makes k_assemble slow: ncu showed 10.8 of 32
// lanes active per instruction and the stencils in local memory.  For a level-0 face whose two cells are ACTIVE (or outside the
// grid) and whose stencil faces are all DOFs or OUTSIDE -- no UNASSIGNED (coarser) face, no SOLID face -- every branch of
// edgeStressFaces / centerStressFaces / applyToMatrix is decided by a handful of labels:
//   * the two centre stresses contribute (self, self -/+ 1 along the axis) with coefficients -/+ 1/dx;
//   * the four edge stresses (two transverse axes x two sides) contribute the own-axis neighbour across the edge and the two
//     transverse faces that meet at the edge, all with +/- 0.5/G, G = float(float(0 + dx/2) + dx/2) (no transition: AV.cpp:1755-1782);
//   * control volumes are 1 (faceOctreeVolume / edgeOctreeVolume of an all-level-0 neighbourhood).
// This kernel fetches all labels, indices and weights of a row up front (one memory round trip after the key), evaluates exactly
// those expressions in buildRow's order -- the entries, their order and every rounding are bit-identical to k_assemble's
// (tests/test_gpu_kernel_variants.py) -- and hands every other row to k_assemble through `rowList`.
struct SimpleRowOut {   // entries go straight to the column-major staging area (the decision "simple" is taken before the first add)
    int n;
    int32_t *col;
    double *val, last;
    long long stride;
    AVS_DEV __forceinline__ void add(int32_t c, double v) {
        col[(long long)n * stride] = c;
        val[(long long)n * stride] = v;
        last = v;
        ++n;
    }
};
template <int AXIS>
AVS_DEV __forceinline__ bool buildSimpleRow(const DeviceScene &S, int32_t vi, const I3 &face, float fwRaw, SimpleRowOut &row, double &massOut) {
    const Grid3<uint8_t> &lab = S.label[0];
    const Grid3<int32_t> &fg = S.face[0][AXIS];
    constexpr int T1 = (AXIS + 1) % 3, T2 = (AXIS + 2) % 3;           // the two transverse axes
    constexpr int EA0 = T1 < T2 ? T1 : T2, EA1 = T1 < T2 ? T2 : T1;   // buildRow visits the edge axes in ascending order
    // ---- everything a decision may need, fetched unconditionally with clamped indices ---------------------------------------
    const I3 cell[2] = {faceToCell(face, AXIS, 0), face};
    bool cellIn[2];
    uint8_t cl[2];
    int8_t cen[2];
    float cw[2];
    int32_t along[2];   // face -/+ 1 along the axis
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        cellIn[d] = cell[d][AXIS] >= 0 && cell[d][AXIS] < lab.n[AXIS];
        cl[d] = lab.get(cell[d]);
        cen[d] = S.center[0].get(cell[d]);
        cw[d] = S.centerW.get(cell[d]);
        I3 f = face;
        f[AXIS] += d ? 1 : -1;
        along[d] = fg.get(f);
    }
    int8_t ei[2][2];
    float ew[2][2];
    int32_t own[2][2], cross[2][2][2];
    bool ownIn[2][2], crossIn[2][2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ea = q ? EA1 : EA0, ta = 3 - AXIS - ea;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const I3 e = faceToEdge(face, AXIS, ea, dir);
            ei[q][dir] = S.edge[0][ea].get(e);
            ew[q][dir] = S.edgeW[ea].get(e);
            // own-axis face across the edge: edgeToFace(e, ea, AXIS, dir) = face -/+ 1 along ta
            I3 of = face;
            of[ta] += dir ? 1 : -1;
            ownIn[q][dir] = of[ta] >= 0 && of[ta] < fg.n[ta];
            own[q][dir] = fg.get(of);
            // the two ta-faces that meet at the edge: edgeToFace(e, ea, ta, d2) = e - (d2 == 0) along AXIS
            const Grid3<int32_t> &tg = S.face[0][ta];
#pragma unroll
            for (int d2 = 0; d2 < 2; ++d2) {
                I3 cf = e;
                if (d2 == 0) --cf[AXIS];
                crossIn[q][dir][d2] = cf[AXIS] >= 0 && cf[AXIS] < tg.n[AXIS];
                cross[q][dir][d2] = tg.get(cf);
            }
        }
    }
    // ---- is it a simple row? ------------------------------------------------------------------------------------------------
    bool simple = true;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (cellIn[d] && cl[d] != L_ACTIVE) simple = false;                        // the stress cell is the parent (AV.cpp:2562-2572)
        if (cellIn[d] && cen[d] >= 0 && along[d] == F_SOLID) simple = false;       // boundary term
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            if (ei[q][dir] < 0) continue;
            if (ownIn[q][dir] && (own[q][dir] == F_UNASSIGNED || own[q][dir] == F_SOLID)) simple = false;
#pragma unroll
            for (int d2 = 0; d2 < 2; ++d2)
                if (crossIn[q][dir][d2] && (cross[q][dir][d2] == F_UNASSIGNED || cross[q][dir][d2] == F_SOLID)) simple = false;
        }
    if (!simple) return false;
    // ---- the row, in buildRow's order ---------------------------------------------------------------------------------------
    const double dx = S.levelDx(0);
    const double G = (double)(float)((double)(float)(0.0 + .5 * dx) + .5 * dx);   // gradientDx of an edge without transition
    double diag = 0;
    row.n = 0;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (!cellIn[d] || cen[d] < 0) continue;
        double w = (double)cw[d];                                                  // centerStressWeight, level 0
        if (!S.viscosity.d) w *= (double)S.viscosity.constant;
        else {
            double p[3];
            S.centerPos(cell[d], 0, p);
            w *= S.viscosity.value(p);
        }
        double coefficient = 2. * S.dt * w;
        const double cLo = -1. / dx, cHi = 1. / dx;                                // sign / dx
        // stencil order: low face, high face.  d == 0: (face - 1, self); d == 1: (self, face + 1)
        coefficient *= d == 0 ? cHi : cLo;
        if (d == 0) {
            if (along[0] >= 0) row.add(along[0], coefficient * cLo);
            diag += coefficient * cHi;
        } else {
            diag += coefficient * cLo;
            if (along[1] >= 0) row.add(along[1], coefficient * cHi);
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ea = q ? EA1 : EA0, ta = 3 - AXIS - ea;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            if (ei[q][dir] < 0) continue;
            double w = (double)ew[q][dir];                                         // edgeStressWeight, level 0; edgeOctreeVolume == 1 here
            if (!S.viscosity.d) w *= (double)S.viscosity.constant;
            else {
                double p[3];
                S.edgePos(faceToEdge(face, AXIS, ea, dir), ea, 0, p);
                w *= S.viscosity.value(p);
            }
            double coefficient = 4. * S.dt * w;
            const double cNeg = .5 * -1. / G, cPos = .5 * 1. / G;
            // own face is edgeToFace(e, ea, AXIS, d2) with d2 = 1 - dir: its coefficient is cPos for dir == 0, cNeg for dir == 1
            coefficient *= dir == 0 ? cPos : cNeg;
            // stencil order: face axes ascending (skipping ea), d2 = 0, 1
            if (AXIS < ta) {
                if (dir == 0) {   // d2 = 0: the neighbour (face - 1 along ta), d2 = 1: self
                    if (ownIn[q][dir] && own[q][dir] >= 0) row.add(own[q][dir], coefficient * cNeg);
                    diag += coefficient * cPos;
                } else {          // d2 = 0: self, d2 = 1: the neighbour (face + 1 along ta)
                    diag += coefficient * cNeg;
                    if (ownIn[q][dir] && own[q][dir] >= 0) row.add(own[q][dir], coefficient * cPos);
                }
            }
#pragma unroll
            for (int d2 = 0; d2 < 2; ++d2)
                if (crossIn[q][dir][d2] && cross[q][dir][d2] >= 0) row.add(cross[q][dir][d2], coefficient * (d2 == 0 ? cNeg : cPos));
            if (!(AXIS < ta)) {
                if (dir == 0) {
                    if (ownIn[q][dir] && own[q][dir] >= 0) row.add(own[q][dir], coefficient * cNeg);
                    diag += coefficient * cPos;
                } else {
                    diag += coefficient * cNeg;
                    if (ownIn[q][dir] && own[q][dir] >= 0) row.add(own[q][dir], coefficient * cPos);
                }
            }
        }
    }
    double fw = (double)fwRaw;                                                     // velocity control volume; faceOctreeVolume == 1 here
    if (!S.density.d) fw *= (double)S.density.constant;
    else {
        double p[3];
        S.facePos(face, AXIS, 0, p);
        fw *= S.density.value(p);
    }
    row.add(vi, fw + diag);
    massOut = fw;   // no boundary terms in a simple row: rhs_i = 0 + M_u u^n_i (k_finish_rhs)
    return true;
}

__global__ void __launch_bounds__(128) k_assemble_simple(const __grid_constant__ DeviceScene S, const RowKey *keys, double *rowMass, const float *fwRow,
                                                         long long rowBegin, long long nLocal, long long stride, int32_t *rowCount,
                                                         int32_t *stageCol, double *stageVal, double *rhs, double *diagOut,
                                                         int32_t *rowList, unsigned long long *rowListCount) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    if (r < nLocal) {
        const long long gi = rowBegin + r;
        const RowKey k = keys[gi];
        pass = true;
        if (k.level == 0) {
            const I3 face = mk3(k.i, k.j, k.k);
            const float fwi = fwRow[gi];
            SimpleRowOut row;
            row.col = stageCol + r;
            row.val = stageVal + r;
            row.stride = stride;
            row.last = 0;
            double mass = 0;
            bool ok;
            if (k.axis == 0) ok = buildSimpleRow<0>(S, (int32_t)gi, face, fwi, row, mass);
            else if (k.axis == 1) ok = buildSimpleRow<1>(S, (int32_t)gi, face, fwi, row, mass);
            else ok = buildSimpleRow<2>(S, (int32_t)gi, face, fwi, row, mass);
            if (ok) {
                pass = false;
                rowCount[r] = row.n;
                rhs[r] = 0;
                rowMass[r] = mass;
                diagOut[r] = row.last;   // the diagonal is added last (AV.cpp:2768)
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);   // warp-aggregated append of the rows left for k_assemble
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(rowListCount, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) rowList[base + __popc(m & ((1u << lane) - 1))] = (int32_t)r;
    }
}

// Column check of the assembled rows.  A scene outside the reference's contract -- liquid that reaches the boundary of the grid: the
// reference's own debug checks (octreeLabels.unitTest, edgeStressUnitTest, centerStresUnitTest, AV.cpp:411-413, 881) fail on such a
// scene -- can leave a LABEL (< 0) where a stencil expects a degree of freedom: getEdgeStressFaces appends the parent of an UNASSIGNED
// face without looking at it (`assert(parentVelocityIndex >= 0)` is compiled out of release builds, AV.cpp:1886-1894), and Eigen's
// setFromTriplets then writes out of bounds.  Here such a row is detected before any kernel uses its columns as addresses (SJDS fill,
// halo discovery, the CG's gathers) and the solve ends with AVS_ERR_UNSUPPORTED.  flag[1] = high half of the 64-bit word whose low
// half is k_assemble's overflow flag.
__global__ void k_check_columns(long long nLocal, long long stride, const int32_t *rowCount, const int32_t *stageCol, long long nRows, int *flag) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nLocal) return;
    const int n = rowCount[r];
    bool bad = false;
    for (int i = 0; i < n; ++i) {
        const int32_t col = stageCol[(long long)i * stride + r];
        bad |= col < 0 || (long long)col >= nRows;
    }
    if (bad) flag[1] = 1;
}

// canonical CSR from the staging area (lazy: only for read-back)
__global__ void k_csr_from_stage(long long nLocal, long long stride, const int32_t *rowCount, const long long *ptr,
                                 const int32_t *stageCol, const double *stageVal, int32_t *col, double *val) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nLocal) return;
    int n = rowCount[r];
    int32_t c[MAX_ROW];
    double v[MAX_ROW];
    for (int i = 0; i < n; ++i) {  // insertion sort by column while loading
        int32_t cc = stageCol[(long long)i * stride + r];
        double vv = stageVal[(long long)i * stride + r];
        int j = i - 1;
        while (j >= 0 && c[j] > cc) { c[j + 1] = c[j]; v[j + 1] = v[j]; --j; }
        c[j + 1] = cc;
        v[j + 1] = vv;
    }
    long long o = ptr[r];
    for (int i = 0; i < n; ++i) { col[o + i] = c[i]; val[o + i] = v[i]; }
}

int avs_stage_system(AvsContext *c, const AvsParams *p) {
    (void)p;
    // row partition: computed with the numbering (avs_stage_octree_labels)
    c->rowBegin = c->rowStarts[c->rank];
    c->rowEnd = c->rowStarts[c->rank + 1];
    const long long nLocal = c->rowEnd - c->rowBegin;
    const long long stride = (nLocal + 31) / 32 * 32;
    c->stageStride = stride;
    c->csrValid = false;
    if (c->rowCount.reserve((size_t)std::max<long long>(nLocal, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->csrPtr.reserve((size_t)(nLocal + 1) * sizeof(long long))) return AVS_ERR_ALLOC;
    if (c->rhs.reserve((size_t)std::max<long long>(nLocal, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    if (c->diag.reserve((size_t)std::max<long long>(nLocal, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    if (c->stageCol.reserve((size_t)std::max<long long>(stride, 32) * MAX_ROW * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->stageVal.reserve((size_t)std::max<long long>(stride, 32) * MAX_ROW * sizeof(double))) return AVS_ERR_ALLOC;
    int *flag = (int *)(c->counters.as<unsigned long long>() + 25);
    AVS_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(unsigned long long), c->stream));
    c->nnz = 0;
    bool rowListCountHost = false;   // the split assembly ran (counter 27 holds the rows it left for k_assemble)
    if (c->faceWRow.reserve((size_t)std::max<long long>(c->nRows, 1) * sizeof(float))) return AVS_ERR_ALLOC;
    if (c->rowMass.reserve((size_t)std::max<long long>(nLocal, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    // face weights that were uploaded whole (pageable host arrays) must have landed; mapped ones are read in place
    if (!c->faceWMappedBytes) cudaStreamWaitEvent(c->stream, c->evUploadDone, 0);
    if (nLocal > 0) {
        unsigned blocks = (unsigned)((nLocal + 127) / 128);
        k_gather_face_weights<<<(unsigned)((nLocal + 255) / 256), 256, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), c->rowBegin, nLocal,
                                                                                     c->faceWRow.as<float>());
        ++c->launches;
        // AVS_ASM_MINB: resident CTAs per SM the register allocation aims at.  Measured at C3 (stage 'system'): 4 (128 registers,
        // no spills) 36.5 ms, 5: 32.9, 6: 32.2, 8 (64 registers, 788 B spills) 30.8 ms -- the kernel is latency-bound (ncu: 19 % warps
        // active, long-scoreboard stalls), so occupancy beats spills.
        static int minb = -1;
        if (minb < 0) { const char *e = getenv("AVS_ASM_MINB"); minb = e ? atoi(e) : 8; }
        // Row accumulator of k_assemble (avs_rowacc.cuh): hashed lookup by default, AVS_ASM_ROW=linear for the linear search.  Same
        // entries in the same order either way.  Measured at C3, stage "system": single-pass assembly 28.9 (hash) vs 30.7 ms (linear);
        // second pass of the split assembly -- the rows it is left with are the long ones (T-junctions, 26-46 entries) -- 24.1 vs 25.7 ms.
        static int hashRow = -1;
        if (hashRow < 0) { const char *e = getenv("AVS_ASM_ROW"); hashRow = (e && e[0] == 'l') ? 0 : 1; }
        // AVS_ASM=generic: every row through k_assemble (the single-pass assembly of round 1; A/B measurements and the bit-equality test)
        static int split = -1;
        if (split < 0) { const char *e = getenv("AVS_ASM"); split = (e && strcmp(e, "generic") == 0) ? 0 : 1; }
        const int32_t *rowList = nullptr;
        const unsigned long long *rowListCount = nullptr;
        if (split) {
            if (c->asmRowList.reserve((size_t)nLocal * sizeof(int32_t))) return AVS_ERR_ALLOC;
            unsigned long long *lc = c->counters.as<unsigned long long>() + 27;
            AVS_CUDA_CHECK(cudaMemsetAsync(lc, 0, sizeof(unsigned long long), c->stream));
            k_assemble_simple<<<blocks, 128, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), c->rowMass.as<double>(), c->faceWRow.as<float>(), c->rowBegin, nLocal, stride,
                                                           c->rowCount.as<int32_t>(), c->stageCol.as<int32_t>(), c->stageVal.as<double>(),
                                                           c->rhs.as<double>(), c->diag.as<double>(), c->asmRowList.as<int32_t>(), lc);
            ++c->launches;
            rowList = c->asmRowList.as<int32_t>();
            rowListCount = lc;
            rowListCountHost = true;
        }
#define ASM_LAUNCH(M)                                                                                                              \
    do {                                                                                                                           \
    if (hashRow)                                                                                                                   \
        k_assemble<M, RowAccHash><<<blocks, 128, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), c->rowMass.as<double>(), c->faceWRow.as<float>(), c->rowBegin, nLocal, stride, \
                                                 c->rowCount.as<int32_t>(), c->stageCol.as<int32_t>(), c->stageVal.as<double>(),      \
                                                 c->rhs.as<double>(), c->diag.as<double>(), flag, rowList, rowListCount);                                    \
    else                                                                                                                           \
    k_assemble<M, RowAcc><<<blocks, 128, 0, c->stream>>>(c->S, c->rowKeys.as<RowKey>(), c->rowMass.as<double>(), c->faceWRow.as<float>(), c->rowBegin, nLocal, stride,       \
                                                 c->rowCount.as<int32_t>(), c->stageCol.as<int32_t>(), c->stageVal.as<double>(),      \
                                                 c->rhs.as<double>(), c->diag.as<double>(), flag, rowList, rowListCount);                                    \
    } while (0)
        if (minb >= 8) ASM_LAUNCH(8);
        else if (minb >= 6) ASM_LAUNCH(6);
        else if (minb == 5) ASM_LAUNCH(5);
        else ASM_LAUNCH(4);
#undef ASM_LAUNCH
        ++c->launches;
        k_check_columns<<<(unsigned)((nLocal + 255) / 256), 256, 0, c->stream>>>(nLocal, stride, c->rowCount.as<int32_t>(), c->stageCol.as<int32_t>(), c->nRows, flag);
        ++c->launches;
        int64_t nnz = 0;
        int rc = avs_exclusive_scan_i32_to_i64(c, c->rowCount.as<int32_t>(), c->csrPtr.as<int64_t>(), nLocal, &nnz);
        if (rc) return rc;
        c->nnz = nnz;
        AVS_CUDA_CHECK(cudaMemcpyAsync(c->csrPtr.as<long long>() + nLocal, &c->nnz, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    } else {
        long long z = 0;
        AVS_CUDA_CHECK(cudaMemcpyAsync(c->csrPtr.p, &z, sizeof(z), cudaMemcpyHostToDevice, c->stream));
    }
    {
        // the two error flags of the assembly (word 25: overflow in the low half, invalid column in the high half): summed over the
        // ranks, so that every rank of a row-partitioned solve leaves with the same status instead of waiting for a peer that left
        if (c->nranks > 1) {
            int rc = avs_dist_allreduce_u64(c, c->counters.as<unsigned long long>() + 25, 1);
            if (rc) return rc;
        }
        unsigned long long hc[3] = {0, 0, 0};   // counters 25 (error flags), 26 (stage 11), 27 (rows left for k_assemble)
        AVS_CUDA_CHECK(cudaMemcpyAsync(hc, c->counters.as<unsigned long long>() + 25, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
        AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if ((int)(hc[0] & 0xffffffffull)) {
            c->lastError = "a matrix row exceeded MAX_ROW entries";
            return AVS_ERR_UNSUPPORTED;
        }
        if (hc[0] >> 32) {
            c->lastError = "a stress stencil references a face that is not a degree of freedom: the liquid reaches the boundary of the grid "
                           "(the reference asserts on such a scene, HDK_AdaptiveViscosity.cpp:411-413, 881, 1891); pad the fields";
            return AVS_ERR_UNSUPPORTED;
        }
        if (nLocal > 0) {
            c->asmGenericRows = rowListCountHost ? (long long)hc[2] : nLocal;
            AVS_TRACE("assembly: %lld of %lld rows through k_assemble", c->asmGenericRows, nLocal);
        }
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    c->haveSystem = true;
    c->haveSolution = false;
    return AVS_OK;
}

// rhs_i += M_u u^n_i (AV.cpp:2767-2772).  The assembly leaves rhs_i = -(boundary terms) and the row's mass M_u = rho V_face;
// the restricted velocity is the LAST input the system needs, so the restriction stage runs after the assembly (the velocity's
// host -> device copy then hides under labelling + assembly) and this kernel completes the right-hand side.  The same two operands
// are added as in buildOctreeSystemFromStencilsPartial's last statement: bit-identical.
__global__ void k_finish_rhs(long long n, long long base, const double *x0, const double *rowMass, double *rhs) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) rhs[r] += rowMass[r] * x0[base + r];
}
int avs_finish_rhs(AvsContext *c) {
    const long long n = c->rowEnd - c->rowBegin;
    if (n > 0) {
        k_finish_rhs<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->rowBegin, c->x0.as<double>(), c->rowMass.as<double>(), c->rhs.as<double>());
        ++c->launches;
    }
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}

// Builds csrCol / csrVal of the resident system on demand (avs_get_system_csr).
int avs_build_csr(AvsContext *c) {
    if (c->csrValid) return AVS_OK;
    const long long nLocal = c->rowEnd - c->rowBegin;
    if (c->csrCol.reserve((size_t)std::max<int64_t>(c->nnz, 1) * sizeof(int32_t))) return AVS_ERR_ALLOC;
    if (c->csrVal.reserve((size_t)std::max<int64_t>(c->nnz, 1) * sizeof(double))) return AVS_ERR_ALLOC;
    if (nLocal > 0) {
        k_csr_from_stage<<<(unsigned)((nLocal + 127) / 128), 128, 0, c->stream>>>(nLocal, c->stageStride, c->rowCount.as<int32_t>(),
                                                                                 c->csrPtr.as<long long>(), c->stageCol.as<int32_t>(),
                                                                                 c->stageVal.as<double>(), c->csrCol.as<int32_t>(),
                                                                                 c->csrVal.as<double>());
        ++c->launches;
    }
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    AVS_CUDA_CHECK(cudaGetLastError());
    c->csrValid = true;
    return AVS_OK;
}

