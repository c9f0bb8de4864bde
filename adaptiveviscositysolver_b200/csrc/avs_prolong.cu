// avs_prolong.cu -- stage 11, "Apply Octree Solution to Regular Grid" (HDK_AdaptiveViscosity.cpp:661-707):
//   HDK_OctreeVectorFieldInterpolator's node pyramid (HDK_OctreeVectorFieldInterpolator.h:30-138,
//   HDK_OctreeVectorFieldInterpolator.cpp:118-658) and interpSPGrid (.cpp:660-845), then
//   applyVelocitiesToRegularGrid (HDK_AdaptiveViscosity.cpp:2815-2894).
//
// The reference stores node values / weights in fp32 SIM_RawFields and the octree velocities in fp32 fields
// (setOctreeVelocity, AV.cpp:2779-2813); every store below rounds to float at the same places.  The octree
// velocity fields are not materialised: vel(level, axis, face) = (float)solution[face index], 0 where there is
// no DOF.  Node grids are dense per level (label u8, 3 x value f32, 3 x weight f32, flag u16); all kernels are
// streaming passes over them.  Compiled with -fmad=false like the other parity-critical translation units.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "avs_context.h"

enum : uint8_t { N_INACTIVE = 0, N_ACTIVE = 1, N_DEPENDENT = 2 };

struct NodeLevel {
    Grid3<uint8_t> label;
    Grid3<float> val[3], w[3];
    Grid3<uint16_t> flag;
};
struct NodeScene {
    NodeLevel lv[AVS_MAX_LEVELS];
};

AVS_DEV __forceinline__ I3 nodeToFace(I3 n, int fa, int fi) {  // UTIL.h:187
    if (!(fi & 1)) --n[(fa + 1) % 3];
    if (!(fi & 2)) --n[(fa + 2) % 3];
    return n;
}
AVS_DEV __forceinline__ I3 faceToNode(I3 f, int fa, int ni) {  // UTIL.h:133
    if (ni & 1) ++f[(fa + 1) % 3];
    if (ni & 2) ++f[(fa + 2) % 3];
    return f;
}
AVS_DEV __forceinline__ I3 cellToNode(I3 c, int ni) {  // UTIL.h:88
    if (ni & 1) ++c[0];
    if (ni & 2) ++c[1];
    if (ni & 4) ++c[2];
    return c;
}
AVS_DEV __forceinline__ float octVel(const DeviceScene &S, const double *sol, int level, int axis, const I3 &f) {
    int32_t vi = S.face[level][axis].get(f);
    return vi >= 0 ? (float)sol[vi] : 0.f;
}

// setActiveNodes (VFI.cpp:118-188) + sampleActiveNodes (VFI.cpp:190-286)
AVS_DEV __forceinline__ void nodeSample(const DeviceScene &S, const NodeLevel &nl, const double *sol, int level, int x, int y, int z) {
    const size_t idx = nl.label.lin(x, y, z);
    const I3 node = mk3(x, y, z);
    // cheap reject: a node with an active face is a corner of an ACTIVE cell
    {
        const Grid3<uint8_t> &lab = S.label[level];
        bool any = false;
        for (int q = 0; q < 8 && !any; ++q) {
            int cx = x - ((q & 1) ? 0 : 1), cy = y - ((q & 2) ? 0 : 1), cz = z - ((q & 4) ? 0 : 1);
            if (cx < 0 || cy < 0 || cz < 0 || cx >= lab.n[0] || cy >= lab.n[1] || cz >= lab.n[2]) continue;
            any = lab.d[lab.lin(cx, cy, cz)] == L_ACTIVE;
        }
        if (!any) { nl.label.d[idx] = N_INACTIVE; return; }
    }
    int32_t vis[12];
    bool act = false, inact = false;
    for (int fa = 0; fa < 3; ++fa) {
        const Grid3<int32_t> &fg = S.face[level][fa];
        const int a1 = (fa + 1) % 3, a2 = (fa + 2) % 3;
        for (int fi = 0; fi < 4; ++fi) {
            I3 f = nodeToFace(node, fa, fi);
            int32_t vi;
            if (f[a1] < 0 || f[a2] < 0 || f[a1] >= fg.n[a1] || f[a2] >= fg.n[a2]) { vi = -100; inact = true; }
            else {
                vi = fg.get(f);
                if (vi >= 0) act = true;
                else if (vi == F_SOLID || vi == F_OUTSIDE) inact = true;
            }
            vis[fa * 4 + fi] = vi;
        }
    }
    // (the reference stops scanning at the first disqualifying face; the outcome -- active iff some face is a
    //  DOF and none is solid / outside / out of range -- does not depend on the scan order)
    if (!(act && !inact)) { nl.label.d[idx] = N_INACTIVE; return; }
    nl.label.d[idx] = N_ACTIVE;
    const double weight = (double)(1 << (S.levels - level - 1));
    unsigned flag = 0;
    for (int fa = 0; fa < 3; ++fa) {
        double av = 0, aw = 0;
        for (int fi = 0; fi < 4; ++fi) {
            int32_t vi = vis[fa * 4 + fi];
            if (vi >= 0) {
                av += weight * (double)(float)sol[vi];
                aw += weight;
                flag += 1u << (fa * 4 + fi);
            }  // UNASSIGNED contributes nothing; other labels cannot occur on an active node
        }
        nl.val[fa].d[idx] = (float)av;
        nl.w[fa].d[idx] = (float)aw;
    }
    nl.flag.d[idx] = (uint16_t)flag;
}

// bubbleActiveNodeValues (VFI.cpp:288-355): one thread per PARENT node (its co-located child is unique)
AVS_DEV __forceinline__ void nodeBubble(const NodeLevel &child, const NodeLevel &par, int x, int y, int z) {
    const int cx = 2 * x, cy = 2 * y, cz = 2 * z;
    if (cx >= child.label.n[0] || cy >= child.label.n[1] || cz >= child.label.n[2]) return;
    const size_t pi = par.label.lin(x, y, z), ci = child.label.lin(cx, cy, cz);
    if (child.label.d[ci] != N_ACTIVE || par.label.d[pi] != N_ACTIVE) return;
    par.flag.d[pi] = (uint16_t)(child.flag.d[ci] + par.flag.d[pi]);
    for (int a = 0; a < 3; ++a) {
        par.w[a].d[pi] = (float)((double)child.w[a].d[ci] + (double)par.w[a].d[pi]);
        par.val[a].d[pi] = (float)((double)child.val[a].d[ci] + (double)par.val[a].d[pi]);
    }
    child.label.d[ci] = N_DEPENDENT;
}
__global__ void k_node_bubble(NodeLevel child, NodeLevel par, int zLo) {
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)(blockIdx.y * blockDim.y + threadIdx.y), z = (int)blockIdx.z + zLo;
    if (x >= par.label.n[0] || y >= par.label.n[1]) return;
    nodeBubble(child, par, x, y, z);
}

// finishIncompleteNodes (VFI.cpp:357-567)
AVS_DEV __forceinline__ void nodeFinish(const DeviceScene &S, const NodeLevel &nl, const double *sol, int level, int x, int y, int z) {
    const size_t idx = nl.label.lin(x, y, z);
    if (nl.label.d[idx] != N_ACTIVE) return;
    unsigned flag = nl.flag.d[idx];
    if (flag == 0xFFF) return;
    const I3 node = mk3(x, y, z);
    const int L = S.levels;
    const double lw = (double)(1 << (L - level - 1));
    unsigned temp = flag;
    for (int bit = 0; flag != 0xFFF && bit < 12; ++bit, temp >>= 1) {
        if (temp & 1u) continue;
        const int fa = bit / 4, fi = bit % 4;
        const I3 f = nodeToFace(node, fa, fi);
        bool found = false;
        double ghost = 0;
        if (node[fa] % 2 == 0) {
            I3 pf = parentOf(f);
            int32_t pvi = S.face[level + 1][fa].get(pf);
            if (pvi >= 0) { ghost = (double)(float)sol[pvi]; found = true; }
        }
        if (!found) {
            I3 cell = f;  // HDKfaceToCell(face, faceAxis, 1)
            int sl = L;   // first level >= `level` whose ancestor of the cell is ACTIVE; all labels fetched at once (no dependent chain)
            {
                uint8_t labs[AVS_MAX_LEVELS];
#pragma unroll
                for (int l = 0; l < AVS_MAX_LEVELS; ++l) {
                    const int sh = max(l - level, 0);
                    labs[l] = (l >= level && l < L) ? S.label[l].get(mk3(f[0] >> sh, f[1] >> sh, f[2] >> sh)) : (uint8_t)L_INACTIVE;
                }
#pragma unroll
                for (int l = AVS_MAX_LEVELS - 1; l >= 0; --l)
                    if (labs[l] == L_ACTIVE) sl = l;
            }
            if (sl >= L) { flag += 1u << bit; continue; }  // asserted impossible in the reference (VFI.cpp:490)
            { const int sh = sl - level; cell = mk3(f[0] >> sh, f[1] >> sh, f[2] >> sh); }
            double fp[3];
            S.facePos(f, fa, level, fp);
            const double idxNode = (fp[fa] - S.origin[fa]) / S.levelDx(sl);
            const double iw = idxNode - floor(idxNode);
            for (int dir = 0; dir < 2; ++dir) {
                I3 of = cellToFace(cell, fa, dir);
                int32_t ovi = S.face[sl][fa].get(of);
                const double liw = dir == 0 ? 1. - iw : iw;
                if (ovi >= 0) ghost += liw * (double)(float)sol[ovi];
                else if (ovi == F_UNASSIGNED && sl > 0)
                    for (int ch = 0; ch < 4; ++ch) {
                        int32_t cvi = S.face[sl - 1][fa].get(childFace(of, fa, ch));
                        if (cvi >= 0) ghost += .25 * liw * (double)(float)sol[cvi];
                    }
            }
        }
        double v = (double)nl.val[fa].d[idx];
        v += lw * ghost;
        nl.val[fa].d[idx] = (float)v;
        double w = (double)nl.w[fa].d[idx];
        w += lw;
        nl.w[fa].d[idx] = (float)w;
        flag += 1u << bit;
    }
    nl.flag.d[idx] = (uint16_t)flag;
}

// normalizeActiveNodes (VFI.cpp:569-613)
AVS_DEV __forceinline__ void nodeNormalize(const NodeLevel &nl, int x, int y, int z) {
    const size_t idx = nl.label.lin(x, y, z);
    if (nl.label.d[idx] != N_ACTIVE) return;
    for (int a = 0; a < 3; ++a) nl.val[a].d[idx] = (float)((double)nl.val[a].d[idx] / (double)nl.w[a].d[idx]);
}

// distributeNodeValuesDown (VFI.cpp:615-658)
AVS_DEV __forceinline__ void nodeDistribute(const NodeLevel &child, const NodeLevel &par, int x, int y, int z) {
    const size_t idx = child.label.lin(x, y, z);
    if (child.label.d[idx] != N_DEPENDENT) return;
    const I3 pn = mk3(x >> 1, y >> 1, z >> 1);
    for (int a = 0; a < 3; ++a) child.val[a].d[idx] = par.val[a].get(pn);
    child.label.d[idx] = N_ACTIVE;
}

// ---- launch shapes of the four per-node passes ------------------------------------------------------------------------------
// dense: one thread per node of planes [zLo, zLo + gridDim.z).  tiles (level 0, where 87 % of the nodes live but only those around
// the refined band are ever ACTIVE): the node labels are memset to INACTIVE and the pass runs on the 16^3 node tiles that touch a
// cell tile with an ACTIVE cell (flags of k_tile_flags, avs_labels.cu) -- a node is ACTIVE only as a corner of an ACTIVE cell.
#define NODE_DENSE_XYZ(nl)                                                                                                              \
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)(blockIdx.y * blockDim.y + threadIdx.y), z = (int)blockIdx.z + zLo; \
    if (x >= (nl).label.n[0] || y >= (nl).label.n[1]) return;
template <class F>
__device__ __forceinline__ void forNodeTiles(const uint32_t *list, const unsigned int *count, const int n[3], int zLo, int zHi, F f) {
    if (blockIdx.x >= *count) return;
    const uint32_t t = list[blockIdx.x];
    const int x = (int)(t & 1023u) * AVS_TILE + (int)(threadIdx.x & 15), y = (int)((t >> 10) & 1023u) * AVS_TILE + (int)(threadIdx.x >> 4);
    if (x >= n[0] || y >= n[1]) return;
    const int z0 = max((int)(t >> 20) * AVS_TILE, zLo), z1 = min(min((int)(t >> 20) * AVS_TILE + AVS_TILE, n[2]), zHi + 1);
    for (int z = z0; z < z1; ++z) f(x, y, z);
}
__global__ void k_node_sample(const __grid_constant__ DeviceScene S, NodeLevel nl, const double *sol, int level, int zLo) {
    NODE_DENSE_XYZ(nl)
    nodeSample(S, nl, sol, level, x, y, z);
}
__global__ void k_node_finish(const __grid_constant__ DeviceScene S, NodeLevel nl, const double *sol, int level, int zLo) {
    NODE_DENSE_XYZ(nl)
    nodeFinish(S, nl, sol, level, x, y, z);
}
__global__ void k_node_normalize(NodeLevel nl, int zLo) {
    NODE_DENSE_XYZ(nl)
    nodeNormalize(nl, x, y, z);
}
__global__ void k_node_distribute(NodeLevel child, NodeLevel par, int zLo) {
    NODE_DENSE_XYZ(child)
    nodeDistribute(child, par, x, y, z);
}
__global__ void k_node_sample_tiles(const __grid_constant__ DeviceScene S, NodeLevel nl, const double *sol, int level, int zLo, int zHi,
                                    const uint32_t *list, const unsigned int *count) {
    forNodeTiles(list, count, nl.label.n, zLo, zHi, [&](int x, int y, int z) { nodeSample(S, nl, sol, level, x, y, z); });
}
__global__ void k_node_finish_tiles(const __grid_constant__ DeviceScene S, NodeLevel nl, const double *sol, int level, int zLo, int zHi,
                                    const uint32_t *list, const unsigned int *count) {
    forNodeTiles(list, count, nl.label.n, zLo, zHi, [&](int x, int y, int z) { nodeFinish(S, nl, sol, level, x, y, z); });
}
__global__ void k_node_normalize_tiles(NodeLevel nl, int zLo, int zHi, const uint32_t *list, const unsigned int *count) {
    forNodeTiles(list, count, nl.label.n, zLo, zHi, [&](int x, int y, int z) { nodeNormalize(nl, x, y, z); });
}
__global__ void k_node_distribute_tiles(NodeLevel child, NodeLevel par, int zLo, int zHi, const uint32_t *list, const unsigned int *count) {
    forNodeTiles(list, count, child.label.n, zLo, zHi, [&](int x, int y, int z) { nodeDistribute(child, par, x, y, z); });
}
// node tiles (of the (n+1)^3 node grid) that touch a cell tile holding an ACTIVE cell; node tile t touches cell tiles t - {0,1}^3
__global__ void k_node_tile_list(int n0, int n1, int n2, Grid3<uint8_t> flags, uint32_t *list, unsigned int *count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n0 * n1 * n2) return;
    const int tx = (int)(i % n0), ty = (int)((i / n0) % n1), tz = (int)(i / ((size_t)n0 * n1));
    bool take = false;
    for (int q = 0; q < 8 && !take; ++q) {
        const int cx = tx - (q & 1), cy = ty - ((q >> 1) & 1), cz = tz - (q >> 2);
        if (cx < 0 || cy < 0 || cz < 0 || cx >= flags.n[0] || cy >= flags.n[1] || cz >= flags.n[2]) continue;
        take = (flags.d[flags.lin(cx, cy, cz)] & 1) != 0;   // TILE_HAS_ACTIVE
    }
    if (take) list[atomicAdd(count, 1u)] = (uint32_t)tx | ((uint32_t)ty << 10) | ((uint32_t)tz << 20);
}

// interpSPGrid (VFI.cpp:660-845)
AVS_DEV double interpSPGrid(const DeviceScene &S, const NodeScene &NS, const double *sol, const double pos[3], int axis) {
    const int L = S.levels;
    I3 cell;
    for (int a = 0; a < 3; ++a) cell[a] = (int)floor((pos[a] - S.origin[a]) / S.levelDx(0));
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    // The reference climbs from level 0 until the cell is ACTIVE (VFI.cpp:676-690): a chain of up to L dependent loads.  The
    // ancestors' indices are known up front, so all L labels are fetched at once and the first ACTIVE one is picked.
    int hit = L;
    {
        uint8_t labs[AVS_MAX_LEVELS];
#pragma unroll
        for (int l = 0; l < AVS_MAX_LEVELS; ++l)
            labs[l] = (l < L) ? S.label[l].get(mk3(cell[0] >> l, cell[1] >> l, cell[2] >> l)) : (uint8_t)L_INACTIVE;
#pragma unroll
        for (int l = AVS_MAX_LEVELS - 1; l >= 0; --l)
            if (labs[l] == L_ACTIVE) hit = l;
    }
    if (hit < L) {
        const int level = hit;
        cell = mk3(cell[0] >> level, cell[1] >> level, cell[2] >> level);
        {
            const double h = S.levelDx(level);
            double ifp[3];
            I3 face;
            for (int a = 0; a < 3; ++a) {
                ifp[a] = (pos[a] - S.origin[a]) / h - (a == axis ? 0.0 : 0.5);
                face[a] = (int)floor(ifp[a]);
            }
            const Grid3<int32_t> &fg = S.face[level][axis];
            int32_t vis[8];
            bool transition = false;
            for (int q = 0; q < 8; ++q) {
                vis[q] = fg.get(cellToNode(face, q));
                if (vis[q] == F_UNASSIGNED) { transition = true; break; }
            }
            if (!transition) {
                float iw[3];  // UT_Vector3 interpolationWeight
                for (int a = 0; a < 3; ++a) {
                    iw[a] = (float)(ifp[a] - (double)face[a]);
                    iw[a] = fminf(fmaxf(iw[a], 0.f), 1.f);
                }
                double value = 0;
                for (int q = 0; q < 8; ++q) {
                    double w = 1.;
                    w *= (q & 1) ? (double)iw[0] : (1. - (double)iw[0]);
                    w *= (q & 2) ? (double)iw[1] : (1. - (double)iw[1]);
                    w *= (q & 4) ? (double)iw[2] : (1. - (double)iw[2]);
                    value += w * (double)(vis[q] >= 0 ? (float)sol[vis[q]] : 0.f);
                }
                return value;
            }
            double ciw = (pos[axis] - S.origin[axis]) / h - (double)cell[axis];
            ciw = fmin(fmax(ciw, 0.), 1.);
            double fiv[2] = {0., 0.};
            for (int dir = 0; dir < 2; ++dir) {
                I3 af = cellToFace(cell, axis, dir);
                int fl = level;
                if (fg.get(af) == F_UNASSIGNED && level > 0) {
                    const double hc = S.levelDx(level - 1);
                    const double c1 = (pos[a1] - S.origin[a1]) / hc, c2 = (pos[a2] - S.origin[a2]) / hc;
                    for (int ch = 0; ch < 4; ++ch) {
                        I3 cf = childFace(af, axis, ch);
                        if ((double)cf[a1] <= c1 && (double)cf[a2] <= c2 && (double)(cf[a1] + 1) >= c1 && (double)(cf[a2] + 1) >= c2) {
                            fl = level - 1;
                            af = cf;
                            break;
                        }
                    }
                }
                const double hf = S.levelDx(fl);
                const double n1 = (pos[a1] - S.origin[a1]) / hf, n2 = (pos[a2] - S.origin[a2]) / hf;
                const double w0 = n1 - floor(n1), w1 = n2 - floor(n2);
                const double faceVelocity = (double)octVel(S, sol, fl, axis, af);
                double avg = 0;
                for (int q = 0; q < 4; ++q) {
                    I3 node = faceToNode(af, axis, q);
                    double w = 1.;
                    w *= (node[a1] - af[a1] == 0) ? 1. - w0 : w0;
                    w *= (node[a2] - af[a2] == 0) ? 1. - w1 : w1;
                    const double nv = (double)NS.lv[fl].val[axis].get(node);
                    avg += nv;
                    fiv[dir] += nv * w;
                }
                fiv[dir] += 2. * (faceVelocity - .25 * avg) * fmin(w0, fmin(w1, fmin(1. - w0, 1. - w1)));
            }
            return (1. - ciw) * fiv[0] + ciw * fiv[1];
        }
    }
    return 0.;
}

// applyVelocitiesToRegularGrid (AV.cpp:2815-2894)
__global__ void k_apply_regular(const __grid_constant__ DeviceScene S, const NodeScene *NS, int axis, const double *sol, float *out,
                                unsigned long long *interpolated, int zOff) {
    const Grid3<int8_t> g = S.regular[axis];
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, z = (int)blockIdx.z + zOff;
    bool interp = false;
    if (x < g.n[0]) {
        const size_t idx = g.lin(x, y, z);
        const int8_t lab = g.d[idx];
        const I3 face = mk3(x, y, z);
        double p[3];
        if (lab == F_SOLID) {  // AV.cpp:2881-2890
            S.facePos(face, axis, 0, p);
            out[idx] = (float)S.collisionVel[axis].value(p);
        } else if (lab >= 0) {
            int32_t oi = S.face[0][axis].get(face);
            if (oi >= 0) out[idx] = (float)sol[oi];  // AV.cpp:2856-2857
            else if (oi == F_SOLID) {                // AV.cpp:2860-2867
                S.facePos(face, axis, 0, p);
                out[idx] = (float)S.collisionVel[axis].value(p);
            } else if (oi == F_UNASSIGNED) {         // AV.cpp:2868-2876
                S.facePos(face, axis, 0, p);
                out[idx] = (float)interpSPGrid(S, *NS, sol, p, axis);
                interp = true;
            }
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, interp);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(interpolated, (unsigned long long)__popc(m));
}

#define NODE_YB 1   // grid rows per CTA; 4 (128 x 4 threads) measured slower for the node pyramid at C3 (apply 22.3 -> 24.4 ms)
// node planes [zLo, zLo + nz) of a level
static dim3 grid3(const int n[3], int nz) { return dim3((unsigned)((n[0] + 127) / 128), (unsigned)((n[1] + NODE_YB - 1) / NODE_YB), (unsigned)nz); }

int avs_apply_regular(AvsContext *c, float *dOut[3], unsigned long long *hostInterpolated, float *const *hostOut) {
    const DeviceScene &S = c->S;
    const int L = S.levels;
    const double *sol = c->nranks > 1 ? c->fullSolution.as<double>() : c->solution.as<double>();
    NodeScene NS;
    memset(&NS, 0, sizeof(NS));
    const bool needNodes = L > 1;  // a uniform grid has no UNASSIGNED face: nothing to interpolate
    if (needNodes) {
        for (int l = 0; l < L; ++l) {
            NodeLevel &nl = NS.lv[l];
            int n[3];
            for (int a = 0; a < 3; ++a) n[a] = S.label[l].n[a] + 1;
            const size_t cnt = (size_t)n[0] * n[1] * n[2];
            // one allocation per level: label | flag | 3 val | 3 w
            const size_t bytes = cnt * (1 + 2 + 24) + 8 * 256;   // take() rounds each of the 8 sub-arrays up to 256 B
            if (c->nodes[l].reserve(bytes)) return AVS_ERR_ALLOC;
            char *base = c->nodes[l].as<char>();
            size_t off = 0;
            auto take = [&](size_t b) { char *p = base + off; off += (b + 255) / 256 * 256; return p; };
            float *vals[6];
            for (int k = 0; k < 6; ++k) vals[k] = (float *)take(cnt * 4);
            uint16_t *fl = (uint16_t *)take(cnt * 2);
            uint8_t *lb = (uint8_t *)take(cnt);
            if (off > c->nodes[l].cap) return AVS_ERR_ALLOC;
            for (int a = 0; a < 3; ++a) {
                nl.val[a].d = vals[a]; nl.w[a].d = vals[3 + a];
                for (int k = 0; k < 3; ++k) { nl.val[a].n[k] = n[k]; nl.w[a].n[k] = n[k]; }
            }
            nl.flag.d = fl;
            nl.label.d = lb;
            for (int k = 0; k < 3; ++k) { nl.flag.n[k] = n[k]; nl.label.n[k] = n[k]; }
        }
        // Multi-GPU: this rank writes back only its z-slab of the regular grid, and interpSPGrid reads node values only on the
        // faces of the ACTIVE cell that contains the sample.  Every node kernel is either local or couples CO-LOCATED nodes of
        // adjacent levels (bubble up / distribute down), so the pyramid is needed -- and built -- only between the fine planes
        // zA and zB: the slab widened to the top-level cells it touches (any ACTIVE cell overlapping the slab nests in those).
        int zA = 0, zB = S.Pad[2];
        if (c->nranks > 1) {
            int s0 = S.Pad[2], s1 = 0;
            for (int a = 0; a < 3; ++a) {
                int z0, z1;
                avs_slab_range(c, a, c->rank, &z0, &z1);
                if (z1 > z0) { s0 = std::min(s0, z0); s1 = std::max(s1, z1); }
            }
            const int top = 1 << (L - 1);
            zA = s1 > s0 ? (s0 / top) * top : 0;
            zB = s1 > s0 ? std::min(S.Pad[2], ((s1 + top - 1) / top) * top) : -1;   // empty slab: no node is read
        }
        int zLo[AVS_MAX_LEVELS], nzW[AVS_MAX_LEVELS];
        for (int l = 0; l < L; ++l) {
            zLo[l] = zA >> l;
            nzW[l] = zB < zA ? 0 : std::min((zB >> l), NS.lv[l].label.n[2] - 1) - zLo[l] + 1;
        }
        // level 0 on node tiles when the labelling stage left its tile flags (default); the other levels are 1/8, 1/64, ... of it
        const uint32_t *tl = nullptr;
        unsigned int *tc = nullptr;
        unsigned tcap = 0;
        if (c->tileFlagsGrid.d && nzW[0] > 0) {
            const int *nn = NS.lv[0].label.n;
            const int t0 = (nn[0] + AVS_TILE - 1) / AVS_TILE, t1 = (nn[1] + AVS_TILE - 1) / AVS_TILE, t2 = (nn[2] + AVS_TILE - 1) / AVS_TILE;
            tcap = (unsigned)((size_t)t0 * t1 * t2);
            if (c->nodeTileList.reserve((size_t)tcap * sizeof(uint32_t))) return AVS_ERR_ALLOC;
            tc = (unsigned int *)(c->counters.as<unsigned long long>() + 36);
            AVS_CUDA_CHECK(cudaMemsetAsync(tc, 0, sizeof(unsigned int), c->stream));
            // nodes of tiles that are not listed keep N_INACTIVE (k_node_sample would have written exactly that)
            AVS_CUDA_CHECK(cudaMemsetAsync(NS.lv[0].label.d + (size_t)nn[0] * nn[1] * (size_t)zLo[0], N_INACTIVE, (size_t)nn[0] * nn[1] * (size_t)nzW[0], c->stream));
            k_node_tile_list<<<(tcap + 255) / 256, 256, 0, c->stream>>>(t0, t1, t2, c->tileFlagsGrid, c->nodeTileList.as<uint32_t>(), tc);
            ++c->launches;
            tl = c->nodeTileList.as<uint32_t>();
        }
        const int zHi0 = zLo[0] + nzW[0] - 1;
        for (int l = 0; l < L; ++l) {
            if (nzW[l] <= 0) continue;
            if (l == 0 && tl) k_node_sample_tiles<<<tcap, 256, 0, c->stream>>>(S, NS.lv[0], sol, 0, zLo[0], zHi0, tl, tc);
            else k_node_sample<<<grid3(NS.lv[l].label.n, nzW[l]), dim3(128, NODE_YB), 0, c->stream>>>(S, NS.lv[l], sol, l, zLo[l]);
            ++c->launches;
        }
        for (int l = 0; l < L - 1; ++l) {
            if (nzW[l + 1] <= 0) continue;
            k_node_bubble<<<grid3(NS.lv[l + 1].label.n, nzW[l + 1]), dim3(128, NODE_YB), 0, c->stream>>>(NS.lv[l], NS.lv[l + 1], zLo[l + 1]);
            ++c->launches;
        }
        for (int l = 0; l < L - 1; ++l) {
            if (nzW[l] <= 0) continue;
            if (l == 0 && tl) k_node_finish_tiles<<<tcap, 256, 0, c->stream>>>(S, NS.lv[0], sol, 0, zLo[0], zHi0, tl, tc);
            else k_node_finish<<<grid3(NS.lv[l].label.n, nzW[l]), dim3(128, NODE_YB), 0, c->stream>>>(S, NS.lv[l], sol, l, zLo[l]);
            ++c->launches;
        }
        for (int l = 0; l < L; ++l) {
            if (nzW[l] <= 0) continue;
            if (l == 0 && tl) k_node_normalize_tiles<<<tcap, 256, 0, c->stream>>>(NS.lv[0], zLo[0], zHi0, tl, tc);
            else k_node_normalize<<<grid3(NS.lv[l].label.n, nzW[l]), dim3(128, NODE_YB), 0, c->stream>>>(NS.lv[l], zLo[l]);
            ++c->launches;
        }
        for (int l = L - 2; l >= 0; --l) {
            if (nzW[l] <= 0) continue;
            if (l == 0 && tl) k_node_distribute_tiles<<<tcap, 256, 0, c->stream>>>(NS.lv[0], NS.lv[1], zLo[0], zHi0, tl, tc);
            else k_node_distribute<<<grid3(NS.lv[l].label.n, nzW[l]), dim3(128, NODE_YB), 0, c->stream>>>(NS.lv[l], NS.lv[l + 1], zLo[l]);
            ++c->launches;
        }
    }
    if (c->nodeScene.reserve(sizeof(NodeScene))) return AVS_ERR_ALLOC;
    AVS_CUDA_CHECK(cudaMemcpyAsync(c->nodeScene.p, &NS, sizeof(NS), cudaMemcpyHostToDevice, c->stream));
    unsigned long long *cnt = c->counters.as<unsigned long long>() + 26;
    AVS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
    for (int a = 0; a < 3; ++a) {   // multi-GPU: only this rank's z-slab (the caller all-gathers the slabs)
        int z0, z1;
        avs_slab_range(c, a, c->rank, &z0, &z1);
        if (z1 <= z0) continue;
        dim3 g((unsigned)((S.regular[a].n[0] + 127) / 128), (unsigned)S.regular[a].n[1], (unsigned)(z1 - z0));
        k_apply_regular<<<g, 128, 0, c->stream>>>(S, c->nodeScene.as<NodeScene>(), a, sol, dOut[a], cnt, z0);
        ++c->launches;
        if (hostOut && c->nranks == 1) {
            // host caller: the device -> host copy of this axis runs on the copy stream underneath the next axis' kernel
            AVS_CUDA_CHECK(cudaEventRecord(c->evAxis[a], c->stream));
            AVS_CUDA_CHECK(cudaStreamWaitEvent(c->copyStream, c->evAxis[a], 0));
            AVS_CUDA_CHECK(cudaMemcpyAsync(hostOut[a], dOut[a], S.regular[a].count() * sizeof(float), cudaMemcpyDeviceToHost, c->copyStream));
        }
    }
    if (hostOut && c->nranks == 1) AVS_CUDA_CHECK(cudaEventRecord(c->evDownloadDone, c->copyStream));
    if (c->nranks > 1) {
        int rcd = avs_dist_allreduce_u64(c, cnt, 1);
        if (rcd) return rcd;
        // distributed output (AvsDeviceConfig.distributed_output, always on for avs_create_multi): every rank keeps / downloads
        // the slab it computed; otherwise the slabs are all-gathered so that every rank holds the whole field
        if (!c->slabOutputOnly && (rcd = avs_dist_allgather_slabs(c, dOut))) return rcd;
    }
    AVS_CUDA_CHECK(cudaMemcpyAsync(hostInterpolated, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    AVS_CUDA_CHECK(cudaStreamSynchronize(c->stream));  // also keeps the stack copy of NS alive until the H2D copy is done
    AVS_CUDA_CHECK(cudaGetLastError());
    return AVS_OK;
}
