// InteractionGraph for the B200 engine - a gather-form, atomics-free re-design of the reference's
// PairlistComputation + InteractionGraph (src/interaction_graph.h:31-557).
//
// Reference: one global edge list per graph (Verlet-cached, SSE left-packed), per-edge value/deriv arrays, a serial
// scatter-add of edge_sensitivity*edge_deriv.  Here, per replica, the pair list is held as ELL neighbour tables
//     nbr1[B][n1][K1] : for each element of group 1 the indices of its partners in group 2 (ascending)
//     nbr2[B][n2][K2] : the transpose (for a symmetric graph a single full table is used for both)
// built by one thread per element scanning shared-memory tiles of the other group.  Every consumer then works in
// gather form - a small lane group per element walks its row, evaluates the pair term, and reduces with warp
// shuffles - so forces need no atomics and are summed in a fixed order (bit-reproducible run to run).  Pair-term
// derivatives are recomputed in the backward pass instead of being stored: at B200's flop:byte ratio ~300 flops is
// cheaper than writing and re-reading 52 bytes per edge from HBM.
//
// The pair predicate is the reference's (pinned build): fp32, no FMA, ((dx*dx+dy*dy)+dz*dz) < cutoff^2 with
// dx = pos1[i1]-pos2[i2], plus the interaction's id exclusion and, for symmetric graphs, i1<i2
// (interaction_graph.h:122-158,223-244).  The reference's emission order is recovered on request by sorting.
#pragma once
#include "common.cuh"
#include "engine.h"

namespace ub {

enum Exclusion { EXCL_NONE = 0, EXCL_ROTAMER = 1, EXCL_SEQ2 = 2, EXCL_SEQ1 = 3 };

__device__ __forceinline__ bool acceptable_id_pair(int excl, int id1, int id2) {
    switch (excl) {
        case EXCL_ROTAMER: return (((unsigned)id1) >> 4) != (((unsigned)id2) >> 4);   // bead_interaction.h:195-197
        case EXCL_SEQ2: return (id1 - id2 > 2) || (id2 - id1 > 2);                     // hbond.cpp:254-259
        case EXCL_SEQ1: return (id1 - id2 > 1) || (id2 - id1 > 1);                     // backbone_steric.cpp:32-35
        default: return true;
    }
}

struct IGraphSide {
    const float* out;   // node output [B][n_node][wp]
    float* sens;
    int n_node, wp;
    const int *loc, *type, *id;   // per element of the group
    int n;
};

struct IGraphDev {
    IGraphSide s1, s2;
    const float* param;   // [n_type1][n_type2][n_param]
    int n_type1, n_type2, n_param;
    float cutoff, cutoff2;
    int symmetric, excl;
    unsigned short* nbr1; int* cnt1; int K1;
    unsigned short* nbr2; int* cnt2; int K2;
    int* error_flag;
};

// Build the ELL rows of group A against group B.  grid (ceil(nA/TILE), G), block TILE.  With rep_list == nullptr
// blockIdx.y is the replica; otherwise the CTAs with the same blockIdx.x stride over the compacted list of replicas that
// need their Verlet cache rebuilt (rep_list[0..*n_list)).
template <int TILE>
__global__ void k_pairlist(IGraphSide A, IGraphSide Bs, unsigned short* __restrict__ nbr, int* __restrict__ cnt, int K,
                           float cutoff2, int excl, int same_group, int a_is_first, int* error_flag,
                           const int* __restrict__ rep_list, const int* __restrict__ n_list, int colmajor) {
    __shared__ float4 tile[TILE];   // x,y,z, id (bit-cast)
    const int n_rep_todo = rep_list ? *n_list : gridDim.y;
    for (int ridx = blockIdx.y; ridx < n_rep_todo; ridx += gridDim.y) {
        const int r = rep_list ? rep_list[ridx] : ridx;
        int i = blockIdx.x * TILE + threadIdx.x;
        bool active = i < A.n;
        float xi = 0.f, yi = 0.f, zi = 0.f;
        int idi = 0;
        if (active) {
            const float* p = A.out + (size_t(r) * A.n_node + A.loc[i]) * A.wp;
            xi = p[0]; yi = p[1]; zi = p[2];
            idi = A.id[i];
        }
        int n = 0;
        // row-major rows [i][k] for the exact tables; the Verlet candidates are stored in slices of eight, [k/8][i][8]: the
        // thread that owns row i (here and in k_refine) moves 16 bytes per access and neighbouring threads touch
        // neighbouring 16-byte words
        unsigned short* row = colmajor ? nbr + size_t(r) * K * A.n : nbr + (size_t(r) * A.n + (active ? i : 0)) * K;
        const int ii = active ? i : 0;
        for (int j0 = 0; j0 < Bs.n; j0 += TILE) {
            int j = j0 + threadIdx.x;
            __syncthreads();
            if (j < Bs.n) {
                const float* p = Bs.out + (size_t(r) * Bs.n_node + Bs.loc[j]) * Bs.wp;
                tile[threadIdx.x] = make_float4(p[0], p[1], p[2], __int_as_float(Bs.id[j]));
            }
            __syncthreads();
            if (!active) continue;
            int jn = min(TILE, Bs.n - j0);
            for (int jj = 0; jj < jn; ++jj) {
                float4 t = tile[jj];
                // pos1 - pos2 with group 1 first, as in the reference refine step; squares make the order immaterial
                float dx = a_is_first ? xi - t.x : t.x - xi;
                float dy = a_is_first ? yi - t.y : t.y - yi;
                float dz = a_is_first ? zi - t.z : t.z - zi;
                float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                bool hit = d2 < cutoff2 && acceptable_id_pair(excl, idi, __float_as_int(t.w)) && !(same_group && (j0 + jj) == i);
                if (hit) {
                    if (n < K) row[colmajor ? (size_t(n >> 3) * A.n + ii) * 8 + (n & 7) : size_t(n)] = (unsigned short)(j0 + jj);
                    ++n;
                }
            }
        }
        if (active) {
            if (n > K) { atomicExch(error_flag, 1); n = K; }
            cnt[size_t(r) * A.n + i] = n;
            // pad the last slice with the sentinel index Bs.n (k_refine keeps a far-away position there)
            if (colmajor) for (int k = n; k & 7; ++k) row[(size_t(k >> 3) * A.n + ii) * 8 + (k & 7)] = (unsigned short)Bs.n;
        }
        __syncthreads();
    }
}

// Verlet cache (reference PairlistComputation::ensure_cache_valid, interaction_graph.h:51-168): candidate rows are built
// with cutoff + skin and reused until some element has moved more than skin/2 from where it was at build time.
// k_cache_check: one CTA per replica; sets flag[r], appends r to the rebuild list and refreshes the cached positions.
static __global__ void k_cache_check(IGraphSide A, IGraphSide Bs, int two_groups, float* __restrict__ cposA, float* __restrict__ cposB,
                              float max_move2, int* __restrict__ flag, int* __restrict__ rep_list, int* __restrict__ n_list) {
    __shared__ int moved;
    const int r = blockIdx.x;
    if (threadIdx.x == 0) moved = flag[r] == 2;   // 2 = never built
    __syncthreads();
    for (int pass = 0; pass < (two_groups ? 2 : 1); ++pass) {
        const IGraphSide& S = pass ? Bs : A;
        const float4* c = reinterpret_cast<const float4*>(pass ? cposB : cposA) + size_t(r) * S.n;
        for (int i = threadIdx.x; i < S.n; i += blockDim.x) {
            const float* p = S.out + (size_t(r) * S.n_node + S.loc[i]) * S.wp;
            float4 q = c[i];
            float dx = p[0] - q.x, dy = p[1] - q.y, dz = p[2] - q.z;
            if (max_move2 < dx * dx + dy * dy + dz * dz) moved = 1;
        }
    }
    __syncthreads();
    if (moved) {
        for (int pass = 0; pass < (two_groups ? 2 : 1); ++pass) {
            const IGraphSide& S = pass ? Bs : A;
            float4* c = reinterpret_cast<float4*>(pass ? cposB : cposA) + size_t(r) * S.n;
            for (int i = threadIdx.x; i < S.n; i += blockDim.x) {
                const float* p = S.out + (size_t(r) * S.n_node + S.loc[i]) * S.wp;
                c[i] = make_float4(p[0], p[1], p[2], 0.f);
            }
        }
    }
    if (threadIdx.x == 0) {
        flag[r] = moved ? 1 : 0;
        if (moved) rep_list[atomicAdd(n_list, 1)] = r;
    }
}

// k_refine: exact rows from candidate rows with the reference's refine predicate (interaction_graph.h:223-244).  One CTA
// per replica stages the positions of both groups in shared memory, then G lanes walk each candidate row (ballot
// compaction keeps the ascending order); for an asymmetric graph the transposed table is refined in the same launch.
struct RefineTable { const unsigned short* cand; const int* ccnt; int Kc; unsigned short* nbr; int* cnt; int K; };

// One thread per candidate row; a slice of eight candidates per 16-byte load, no bounds tests (padding entries index the
// sentinel position, which is never in range).  Hits are packed four to a 64-bit register and leave as 8-byte stores, in
// candidate order (ascending).  The count and the first four slices of a thread's first row are fetched BEFORE the CTA
// stages positions (RefinePrefetch), so that their latency overlaps the staging and its barrier.
struct RefineRole {          // which rows of which table this thread refines
    RefineTable T;
    int nA, a_is_first, t0, n_thread;
    const float4 *posA, *posB;
};
struct RefinePrefetch { int c; uint4 v[4]; };

__device__ __forceinline__ void refine_prefetch(int r, const RefineRole& R, int i, RefinePrefetch& F) {
    F.c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) F.v[q] = make_uint4(0, 0, 0, 0);
    if (i >= R.nA) return;
    F.c = R.T.ccnt[size_t(r) * R.nA + i];
    const uint4* cs = reinterpret_cast<const uint4*>(R.T.cand + size_t(r) * R.T.Kc * R.nA) + i;
    const int n_slice_max = R.T.Kc >> 3;   // slices past the row's count hold stale entries: loaded, never used
#pragma unroll
    for (int q = 0; q < 4; ++q) if (q < n_slice_max) F.v[q] = cs[size_t(q) * R.nA];
}

__device__ __forceinline__ void refine_rows(int r, const RefineRole& R, float cutoff2, int* error_flag, RefinePrefetch& F) {
    const RefineTable& T = R.T;
    const int nA = R.nA;
    for (int i = R.t0; i < nA; i += R.n_thread) {
        if (i != R.t0) refine_prefetch(r, R, i, F);
        const float4 pi = R.posA[i];
        const int c = F.c;
        const uint4* cs = reinterpret_cast<const uint4*>(T.cand + size_t(r) * T.Kc * nA) + i;
        unsigned short* row = T.nbr + (size_t(r) * nA + i) * T.K;   // 16-byte aligned (K is a multiple of 8)
        // the row base stays in one register pair (a store address is then one IMAD.WIDE) and the stores are st.global, which
        // the compiler may move across the shared-memory position loads
        unsigned long long rowp = (unsigned long long)__cvta_generic_to_global(row);
        asm volatile("" : "+l"(rowp));
        unsigned n = 0;
        const int n_slice = (c + 7) >> 3;
        const unsigned last = T.K - 1;
        // slices are fetched four at a time (independent 16-byte loads: one memory round trip per 32 candidates)
        for (int s0 = 0; s0 < n_slice; s0 += 4) {
            uint4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = s0 == 0 ? F.v[q] : (s0 + q < n_slice ? cs[size_t(s0 + q) * nA] : make_uint4(0, 0, 0, 0));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (s0 + q >= n_slice) break;
                const unsigned w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned j = (u & 1) ? (w[u >> 1] >> 16) : (w[u >> 1] & 0xffffu);
                    const float4 pj = R.posB[j];
                    // the reference subtracts group 2 from group 1 (interaction_graph.h:230-232); the squares are the same
                    // either way, bit for bit
                    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    const bool hit = d2 < cutoff2;
                    if (hit) {   // predicated store; an overflowing row is reported below
                        asm volatile("st.global.u16 [%0], %1;" :: "l"(rowp + 2ull * min(n, last)), "h"((unsigned short)j));
                        ++n;
                    }
                }
            }
        }
        if (n > (unsigned)T.K) { atomicExch(error_flag, 1); n = T.K; }
        T.cnt[size_t(r) * nA + i] = (int)n;
    }
}

// Verlet cache handled INSIDE the refine kernel (round 2; k_cache_check and the rebuild launch of k_pairlist are gone from
// the cached path): the CTA has the positions of both groups staged anyway, so it compares them with the positions at build
// time itself; if an element has moved more than skin/2 (or nothing was built yet) it refreshes the cached positions and
// rebuilds ITS replica's candidate rows on the spot - one thread per row against the staged partner positions (a broadcast
// per partner) - before refining.  About one replica in fifteen rebuilds per evaluation; those CTAs run ~3x longer.
struct VerletCache {
    float* cposA; float* cposB;   // positions at build time, [B][n][4]
    int* flag;                    // [B] 2 = never built
    float max_move2, cand_cutoff2;
    int excl, same_group;
};
// candidate rows of the elements of X against Y (cutoff + skin), column-major slices of eight as k_pairlist(colmajor) writes them
__device__ __forceinline__ void rebuild_candidates(int r, const IGraphSide& X, const IGraphSide& Y, const float4* posX, const float4* posY,
                                                   const RefineTable& T, const VerletCache& V, int* error_flag) {
    unsigned short* base = const_cast<unsigned short*>(T.cand) + size_t(r) * T.Kc * X.n;
    for (int i = threadIdx.x; i < X.n; i += blockDim.x) {
        const float4 pi = posX[i];
        const int idi = X.id[i];
        int n = 0;
        for (int j = 0; j < Y.n; ++j) {
            const float4 pj = posY[j];
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < V.cand_cutoff2 && acceptable_id_pair(V.excl, idi, Y.id[j]) && !(V.same_group && j == i)) {
                if (n < T.Kc) base[(size_t(n >> 3) * X.n + i) * 8 + (n & 7)] = (unsigned short)j;
                ++n;
            }
        }
        if (n > T.Kc) { atomicExch(error_flag, 1); n = T.Kc; }
        const_cast<int*>(T.ccnt)[size_t(r) * X.n + i] = n;
        for (int k = n; k & 7; ++k) base[(size_t(k >> 3) * X.n + i) * 8 + (k & 7)] = (unsigned short)Y.n;   // sentinel: the far-away position
    }
}

// `which`: bit 0 = refine table 1, bit 1 = refine the transposed table (two groups only)
#ifndef UB_REFINE_OCC
#define UB_REFINE_OCC 3
#endif
template <int G>
__global__ void __launch_bounds__(256, UB_REFINE_OCC) k_refine(IGraphSide A, IGraphSide Bs, int two_groups, int which, RefineTable T1, RefineTable T2, float cutoff2, int* error_flag, VerletCache V) {
    extern __shared__ float4 sm_pos[];
    const int r = blockIdx.x;
    float4* posA = sm_pos;                                  // [A.n + 1], last = sentinel
    float4* posB = two_groups ? sm_pos + A.n + 1 : sm_pos;  // [Bs.n + 1]
    // one row per thread where the block is large enough; with two groups the first A.n threads (rounded up to whole
    // warps) refine table 1 while the others refine the transposed table
    const int T = blockDim.x;
    const bool two_tables = two_groups && which == 3;
    const int split = two_tables ? max(32, min(T - 32, ((T * A.n / (A.n + Bs.n)) + 16) & ~31)) : 0;   // threads in proportion to rows
    RefineRole R1{T1, A.n, 1, (int)threadIdx.x, T, posA, posB}, R2{T2, Bs.n, 0, (int)threadIdx.x, T, posB, posA};
    const bool both = two_tables && split <= 0;      // block too small to split: every thread does both tables in turn
    if (two_tables && split > 0) {
        if ((int)threadIdx.x < split) R1.n_thread = split;
        else { R1 = R2; R1.t0 = threadIdx.x - split; R1.n_thread = T - split; }
    } else if (two_groups && which == 2) R1 = R2;    // only the transposed table is wanted
    RefinePrefetch F;
    refine_prefetch(r, R1, R1.t0, F);
    const float4 far = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    // stage the positions and, on the way, compare them with the positions the candidate rows were built from
    int moved = V.flag ? V.flag[r] == 2 : 0;
    for (int i = threadIdx.x; i < A.n; i += blockDim.x) {
        const float* p = A.out + (size_t(r) * A.n_node + A.loc[i]) * A.wp;
        const float4 v = make_float4(p[0], p[1], p[2], 0.f);
        posA[i] = v;
        if (V.flag) {
            const float4 q = reinterpret_cast<const float4*>(V.cposA)[size_t(r) * A.n + i];
            const float dx = v.x - q.x, dy = v.y - q.y, dz = v.z - q.z;
            moved |= V.max_move2 < dx * dx + dy * dy + dz * dz;
        }
    }
    if (threadIdx.x == 0) posA[A.n] = far;
    if (two_groups) {
        for (int i = threadIdx.x; i < Bs.n; i += blockDim.x) {
            const float* p = Bs.out + (size_t(r) * Bs.n_node + Bs.loc[i]) * Bs.wp;
            const float4 v = make_float4(p[0], p[1], p[2], 0.f);
            posB[i] = v;
            if (V.flag) {
                const float4 q = reinterpret_cast<const float4*>(V.cposB)[size_t(r) * Bs.n + i];
                const float dx = v.x - q.x, dy = v.y - q.y, dz = v.z - q.z;
                moved |= V.max_move2 < dx * dx + dy * dy + dz * dz;
            }
        }
        if (threadIdx.x == 0) posB[Bs.n] = far;
    }
    moved = __syncthreads_or(moved);   // (also: the staged positions are complete)
    if (moved) {   // uniform: refresh the cache of this replica, then rebuild its candidate rows
        for (int i = threadIdx.x; i < A.n; i += blockDim.x) reinterpret_cast<float4*>(V.cposA)[size_t(r) * A.n + i] = posA[i];
        if (two_groups) for (int i = threadIdx.x; i < Bs.n; i += blockDim.x) reinterpret_cast<float4*>(V.cposB)[size_t(r) * Bs.n + i] = posB[i];
        if (!two_groups || (which & 1)) rebuild_candidates(r, A, two_groups ? Bs : A, posA, posB, T1, V, error_flag);
        if (two_groups && (which & 2)) rebuild_candidates(r, Bs, A, posB, posA, T2, V, error_flag);
        if (threadIdx.x == 0) V.flag[r] = 0;
        __syncthreads();   // this CTA's candidate rows are visible to all its threads
        refine_prefetch(r, R1, R1.t0, F);
    }
    refine_rows(r, R1, cutoff2, error_flag, F);
    if (both) {
        refine_prefetch(r, R2, R2.t0, F);
        refine_rows(r, R2, cutoff2, error_flag, F);
    }
}

// (A warp-cooperative refinement without staging - WarpEdges over the candidate rows, positions gathered from global memory -
// was measured for the sparse asymmetric tables and lost: 803 us against 730 us per evaluation for all pair lists.  A
// candidate costs five global loads there (index, two element rows behind their `loc` indirection) against one shared-memory
// gather here, and candidates outnumber elements ten to one.)

// ---- row scheduling ---------------------------------------------------------------------------------------
// Counting sort of n rows by descending key (row length, clipped to 255) into order[0..n).  Lane groups that take
// consecutive entries of `order` then walk rows of similar length, so a warp's groups finish together.  Results do not
// depend on the processing order (every row is reduced inside its own lane group).  All threads of the CTA must call;
// hist: 256 ints of shared memory.
template <typename KeyF>
__device__ __forceinline__ void sort_rows_desc(int n, KeyF key, int* order, int* hist) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[min(max(key(i), 0), 255)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {   // offsets for descending keys: bin b starts after all larger bins
        int lane = threadIdx.x, sum = 0, v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = hist[255 - (lane * 8 + u)]; sum += v[u]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += t; }
        int run = incl - sum;
#pragma unroll
        for (int u = 0; u < 8; ++u) { hist[255 - (lane * 8 + u)] = run; run += v[u]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&hist[min(max(key(i), 0), 255)], 1)] = i;
    __syncthreads();
}

// ---- element loads ------------------------------------------------------------------------------------
__device__ __forceinline__ const float* elem_ptr(const IGraphSide& s, int r, int e) {
    return s.out + (size_t(r) * s.n_node + s.loc[e]) * s.wp;
}
__device__ __forceinline__ float* elem_sens_ptr(const IGraphSide& s, int r, int e) {
    return s.sens + (size_t(r) * s.n_node + s.loc[e]) * s.wp;
}
__device__ __forceinline__ void load8(const float* p, float* x) {   // rows are 32-byte aligned when wp == 8
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// ---- pair terms ---------------------------------------------------------------------------------------
// quadspline: V = wide(r) + ang1(cos1)*ang2(cos2)*narrow(r); reference bead_interaction.h:30-84.
// x1,x2: (pos, unit direction).  Outputs: value, d/dx1[0..5], d/dx2[0..5].
struct QuadSplineShape { int nka, nk; float inv_dtheta, inv_dx; };

__device__ __forceinline__ float quadspline_edge(const float* __restrict__ p, const QuadSplineShape& q, const float* x1,
                                                 const float* x2, float* d1, float* d2) {
    f3 displace = mk3(x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]);
    f3 rvec1 = mk3(x1[3], x1[4], x1[5]), rvec2 = mk3(x2[3], x2[4], x2[5]);
    float dist2 = mag2(displace);
    float inv_dist = rsqrtf(dist2);
    float dist_coord = dist2 * (inv_dist * q.inv_dx);
    f3 u = inv_dist * displace;
    float cos1 = dot(rvec1, u), cos2 = -dot(rvec2, u);
    float a1v, a1d, a2v, a2d, wv, wd, nv, nd;
    float w[4], d[4];
    {
        float x = (cos1 + 1.f) * q.inv_dtheta + 1.f;
        int b = max(1, min((int)x, q.nka - 3));
        bspline_weights(x - (float)b, w, d);
        bspline_apply(w, d, p[b - 1], p[b], p[b + 1], p[b + 2], a1v, a1d);
        const float* p2 = p + q.nka;
        x = (cos2 + 1.f) * q.inv_dtheta + 1.f;
        b = max(1, min((int)x, q.nka - 3));
        bspline_weights(x - (float)b, w, d);
        bspline_apply(w, d, p2[b - 1], p2[b], p2[b + 1], p2[b + 2], a2v, a2d);
    }
    {   // the wide and narrow radial profiles share the knot interval and hence the weights (clamping: spline.h:275-310)
        const float* wide = p + 2 * q.nka;
        const float* narrow = wide + q.nk;
        const int nk = q.nk;
        if (dist_coord < 1.f) {
            wv = (1.f / 6.f) * wide[0] + (2.f / 3.f) * wide[1] + (1.f / 6.f) * wide[2];
            nv = (1.f / 6.f) * narrow[0] + (2.f / 3.f) * narrow[1] + (1.f / 6.f) * narrow[2];
            wd = nd = 0.f;
        } else if (dist_coord >= (float)(nk - 2)) {
            wv = (1.f / 6.f) * wide[nk - 3] + (2.f / 3.f) * wide[nk - 2] + (1.f / 6.f) * wide[nk - 1];
            nv = (1.f / 6.f) * narrow[nk - 3] + (2.f / 3.f) * narrow[nk - 2] + (1.f / 6.f) * narrow[nk - 1];
            wd = nd = 0.f;
        } else {
            int b = (int)dist_coord;
            bspline_weights(dist_coord - (float)b, w, d);
            bspline_apply(w, d, wide[b - 1], wide[b], wide[b + 1], wide[b + 2], wv, wd);
            bspline_apply(w, d, narrow[b - 1], narrow[b], narrow[b + 1], narrow[b + 2], nv, nd);
        }
    }
    float angular_weight = a1v * a2v;
    float radial_deriv = q.inv_dx * (wd + angular_weight * nd);
    float ang_d1 = q.inv_dtheta * a1d * a2v * nv;
    float ang_d2 = q.inv_dtheta * a1v * a2d * nv;
    f3 rXX = ang_d1 * rvec1 - ang_d2 * rvec2;
    f3 deriv_dir = inv_dist * (rXX - dot(u, rXX) * u);
    f3 d_displace = radial_deriv * u + deriv_dir;
    d1[0] = -d_displace.x; d1[1] = -d_displace.y; d1[2] = -d_displace.z;
    d1[3] = ang_d1 * u.x;  d1[4] = ang_d1 * u.y;  d1[5] = ang_d1 * u.z;
    d2[0] = d_displace.x;  d2[1] = d_displace.y;  d2[2] = d_displace.z;
    d2[3] = -ang_d2 * u.x; d2[4] = -ang_d2 * u.y; d2[5] = -ang_d2 * u.z;
    return wv + angular_weight * nv;
}

// d(quadspline)/d(parameter) (reference quadspline_param_deriv, bead_interaction.h:86-129): the value is linear in the
// coefficients of each of its four splines, so the derivative has 16 non-zero entries - the four basis weights of the
// knot interval of each spline times the other factors.  idx = position inside the type pair's parameter row.
__device__ __forceinline__ void quadspline_param_deriv(const float* __restrict__ p, const QuadSplineShape& q, const float* x1,
                                                       const float* x2, int* idx, float* val) {
    f3 displace = mk3(x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]);
    f3 rvec1 = mk3(x1[3], x1[4], x1[5]), rvec2 = mk3(x2[3], x2[4], x2[5]);
    float dist2 = mag2(displace);
    float inv_dist = rsqrtf(dist2);
    float dist_coord = dist2 * (inv_dist * q.inv_dx);
    f3 u = inv_dist * displace;
    float cos1 = dot(rvec1, u), cos2 = -dot(rvec2, u);
    float wa1[4], wa2[4], wr[4], d[4], a1v, a2v, nv, unused;
    float x = (cos1 + 1.f) * q.inv_dtheta + 1.f;
    const int b1 = max(1, min((int)x, q.nka - 3));
    bspline_weights(x - (float)b1, wa1, d);
    bspline_apply(wa1, d, p[b1 - 1], p[b1], p[b1 + 1], p[b1 + 2], a1v, unused);
    const float* p2 = p + q.nka;
    x = (cos2 + 1.f) * q.inv_dtheta + 1.f;
    const int b2 = max(1, min((int)x, q.nka - 3));
    bspline_weights(x - (float)b2, wa2, d);
    bspline_apply(wa2, d, p2[b2 - 1], p2[b2], p2[b2 + 1], p2[b2 + 2], a2v, unused);
    int rs;   // first coefficient of the radial interval (clamped_deBoor_coeff_deriv, spline.h:375-392)
    if (dist_coord <= 1.f) { rs = 0; wr[0] = 1.f / 6.f; wr[1] = 2.f / 3.f; wr[2] = 1.f / 6.f; wr[3] = 0.f; }
    else if (dist_coord >= (float)(q.nk - 2)) { rs = q.nk - 4; wr[0] = 0.f; wr[1] = 1.f / 6.f; wr[2] = 2.f / 3.f; wr[3] = 1.f / 6.f; }
    else { const int b = (int)dist_coord; rs = b - 1; bspline_weights(dist_coord - (float)b, wr, d); }
    const float* narrow = p + 2 * q.nka + q.nk;
    nv = wr[0] * narrow[rs] + wr[1] * narrow[rs + 1] + wr[2] * narrow[rs + 2] + wr[3] * narrow[rs + 3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        idx[i] = b1 - 1 + i;                       val[i] = a2v * nv * wa1[i];
        idx[4 + i] = q.nka + b2 - 1 + i;           val[4 + i] = a1v * nv * wa2[i];
        idx[8 + i] = 2 * q.nka + rs + i;           val[8 + i] = wr[i];
        idx[12 + i] = 2 * q.nka + q.nk + rs + i;   val[12 + i] = a1v * a2v * wr[i];
    }
}

// HBondCoverageInteraction (hbond.cpp:241-286): quadspline scaled by (1-hb)^2, hb = x1[6]; d1 has 7 components
__device__ __forceinline__ float hbond_coverage_edge(const float* __restrict__ p, const QuadSplineShape& q, const float* x1,
                                                     const float* x2, float* d1, float* d2) {
    float cov = quadspline_edge(p, q, x1, x2, d1, d2);
    float one_m = 1.f - x1[6];
    float pre = one_m * one_m;
#pragma unroll
    for (int k = 0; k < 6; ++k) { d1[k] *= pre; d2[k] *= pre; }
    d1[6] = -cov * one_m * 2.f;
    return pre * cov;
}

// ProteinHBondInteraction (hbond.cpp:152-238).  x1 = (H, rHN), x2 = (O, rOC); returns -log(1-hb) and its derivatives.
// The reference evaluates the angular gate per group of four SIMD lanes (an edge that fails the gate still gets the
// ~1e-6 sigmoid tail when a lane-mate passes); here the gate is per edge, a difference far below tolerance.
__device__ __forceinline__ float protein_hbond_edge(const float* __restrict__ p, const float* x1, const float* x2, float* d1,
                                                    float* d2) {
    f3 H = mk3(x1[0], x1[1], x1[2]), O = mk3(x2[0], x2[1], x2[2]);
    f3 rHN = mk3(x1[3], x1[4], x1[5]), rOC = mk3(x2[3], x2[4], x2[5]);
    f3 HO = H - O;
    float m2 = mag2(HO) + 1e-6f;
    float inv = rsqrtf(m2);
    float mHO = m2 * inv;
    f3 rHO = inv * HO;
    float dotHOC = dot(rHO, rOC), dotOHN = -dot(rHO, rHN);
    f3 dH = mk3(0.f, 0.f, 0.f), drHN = dH, drOC = dH;
    float hb = 0.f;
    if (dotHOC > 0.f && dotOHN > 0.f) {
        float p0 = (*(p)), p1 = (*(p + 1)), p2 = (*(p + 2)), p3 = (*(p + 3)), p4 = (*(p + 4)), p5 = (*(p + 5));
        float ov, od, iv, id_;
        sigmoid_vd((p2 - mHO) * p3, ov, od);   // outer
        sigmoid_vd((mHO - p0) * p1, iv, id_);  // inner
        float rad = ov * iv, rad_d = -p3 * od * iv + p1 * id_ * ov;
        float a1v, a1d, a2v, a2d;
        sigmoid_vd((dotHOC - p4) * p5, a1v, a1d); a1d *= p5;
        sigmoid_vd((dotOHN - p4) * p5, a2v, a2d); a2d *= p5;
        hb = rad * a1v * a2v;
        float c0 = rad_d * a1v * a2v, c1 = rad * a1d * a2v, c2 = -rad * a1v * a2d;
        drOC = c1 * rHO;
        drHN = c2 * rHO;
        dH = c0 * rHO + (c1 * inv) * (rOC - dotHOC * rHO) + (c2 * inv) * (rHN + dotOHN * rHO);
    }
    float hb_log = (hb >= 1.f) ? 100.f : -logf(1.f - hb);
    float pre = fminf(1.f / (1.f - hb), 1e5f);
    d1[0] = dH.x * pre; d1[1] = dH.y * pre; d1[2] = dH.z * pre;
    d1[3] = drHN.x * pre; d1[4] = drHN.y * pre; d1[5] = drHN.z * pre;
    d2[0] = -dH.x * pre; d2[1] = -dH.y * pre; d2[2] = -dH.z * pre;
    d2[3] = drOC.x * pre; d2[4] = drOC.y * pre; d2[5] = drOC.z * pre;
    return hb_log;
}

// EnvironmentCoverageInteraction (environment.cpp:12-68).  x1 = (CB pos, dir), x2 = (bead pos, weight);
// d1 has 6 components, d2 has 4.
__device__ __forceinline__ float environment_edge(const float* __restrict__ p, const float* x1, const float* x2, float* d1,
                                                  float* d2) {
    f3 displace = mk3(x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]);
    f3 rvec1 = mk3(x1[3], x1[4], x1[5]);
    float prob = x2[3];
    float dist2 = mag2(displace);
    float inv_dist = rsqrtf(dist2);
    float dist = dist2 * inv_dist;
    f3 u = inv_dist * displace;
    float r0 = (*(p)), r_sharp = (*(p + 1)), dot0 = (*(p + 2)), dot_sharp = (*(p + 3));
    float dp = dot(u, rvec1);
    float rv, rd, av, ad;
    compact_sigmoid(dist - r0, r_sharp, rv, rd);
    compact_sigmoid(dot0 - dp, dot_sharp, av, ad);
    f3 d_displace = prob * ((rd * av) * u - (rv * ad * inv_dist) * (rvec1 - dp * u));
    float s = -prob * rv * ad;
    d1[0] = -d_displace.x; d1[1] = -d_displace.y; d1[2] = -d_displace.z;
    d1[3] = s * u.x; d1[4] = s * u.y; d1[5] = s * u.z;
    d2[0] = d_displace.x; d2[1] = d_displace.y; d2[2] = d_displace.z;
    float score = rv * av;
    d2[3] = score;
    return prob * score;
}

// ---- host side --------------------------------------------------------------------------------------------
struct IGraphHost {
    bool symmetric;
    int excl;
    CoordNode *node1, *node2;
    int n1, n2, n_type1, n_type2, n_param;
    float cutoff;
    std::vector<int> loc1, loc2, type1, type2, id1, id2;
    std::vector<float> h_param;
    DevBuf<int> d_loc1, d_loc2, d_type1, d_type2, d_id1, d_id2;
    DevBuf<float> d_param;
    DevBuf<unsigned short> nbr1, nbr2;
    DevBuf<int> cnt1, cnt2;
    int K1 = 0, K2 = 0;
    // Verlet cache: candidate rows (cutoff + skin), positions at build time, per-replica rebuild flag + compacted list
    bool use_cache = true;
    float skin = 0.f;
    DevBuf<unsigned short> cand1, cand2;
    DevBuf<int> ccnt1, ccnt2, flag, rep_list, n_list;
    DevBuf<float> cpos1, cpos2;
    int Kc1 = 0, Kc2 = 0;
    // which exact tables the owning node's kernels read (asymmetric graphs; set before allocate()): a table nobody gathers
    // from is neither rebuilt nor refined
    bool need1 = true, need2 = true;
    bool fused_rebuild = true;   // cache check + rebuild inside k_refine (graphs of up to 250 k element pairs)
    bool lists = true;   // false: the owning node builds its own pair structure (rotamer fast build); no tables are allocated
    Engine* engine = nullptr;

    // reads index/type/id(+1/2) and interaction_param (n_type1,n_type2,n_param): interaction_graph.h:305-381
    IGraphHost(const h5l::Node& grp, bool symmetric_, int excl_, int n_dim1, int n_dim2, CoordNode* p1, CoordNode* p2);
    void allocate(Engine* e);          // after cutoff is known
    IGraphDev dev() const;
    void build(cudaStream_t s);        // enqueue the pair-list kernels
    std::vector<float> count_edges_by_type(int replica);
    bool pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2);
    void set_param(const std::vector<float>& p);
};

}  // namespace ub
