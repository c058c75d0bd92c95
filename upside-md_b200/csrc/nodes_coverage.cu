// hbond_coverage (reference hbond.cpp:241-286,371-414) and environment_coverage (environment.cpp:12-109).
//
// One CTA per replica stages both interaction groups (8 floats + type per element) and the parameter table in shared
// memory, then lane groups walk the ELL rows: the forward kernel sums pair values per target element, the backward
// kernel recomputes the pair term and gathers d/d(group 1) and d/d(group 2) in one launch - no atomics, fixed order.
#include <algorithm>
#include <cmath>

#include "edgelist.cuh"
#include "igraph.cuh"

namespace ub {
namespace {

#ifndef UB_CTPB
#define UB_CTPB 256
#endif
constexpr int CTPB = UB_CTPB;
#ifndef UB_COV_RED
#define UB_COV_RED 1   // partner-side derivatives of the backward kernels: 1 = 16-byte global reductions, 0 = shared-memory accumulators
#endif

struct StagedGroup {
    float4* a;   // x,y,z,w0
    float4* b;   // w1..w4
    int* type;
    int* loc;    // element -> row of the parent node's output (replica independent, staged once)
};

// carve `n1`+`n2` staged elements and `n_tab` table floats out of dynamic shared memory; returns the first free byte
__device__ __forceinline__ void* carve_groups(const IGraphDev& g, float4* smem, StagedGroup& S1, StagedGroup& S2, float*& table,
                                              int n_tab) {
    S1.a = smem; S1.b = S1.a + g.s1.n;
    S2.a = S1.b + g.s1.n; S2.b = S2.a + g.s2.n;
    table = reinterpret_cast<float*>(S2.b + g.s2.n);
    S1.type = reinterpret_cast<int*>(table + ((n_tab + 3) & ~3));
    S2.type = S1.type + g.s1.n;
    S1.loc = S2.type + g.s2.n;
    S2.loc = S1.loc + g.s1.n;
    return S1.type + ((2 * (g.s1.n + g.s2.n) + 3) & ~3);
}
// parameter table and element types do not depend on the replica: staged once per (persistent) CTA
__device__ __forceinline__ void stage_table(const IGraphDev& g, const StagedGroup& S1, const StagedGroup& S2, float* table, int n_tab) {
    for (int i = threadIdx.x; i < n_tab; i += blockDim.x) table[i] = g.param[i];
    for (int i = threadIdx.x; i < g.s1.n; i += blockDim.x) { S1.type[i] = g.s1.type[i]; S1.loc[i] = g.s1.loc[i]; }
    for (int i = threadIdx.x; i < g.s2.n; i += blockDim.x) { S2.type[i] = g.s2.type[i]; S2.loc[i] = g.s2.loc[i]; }
    __syncthreads();
}
// both barriers included: the previous replica's readers are done before the overwrite, the data is visible after
// the trailing barrier is the caller's (it usually has more to put into shared memory first)
__device__ __forceinline__ void stage_groups(const IGraphDev& g, int r, const StagedGroup& S1, const StagedGroup& S2) {
    __syncthreads();
    for (int i = threadIdx.x; i < g.s1.n; i += blockDim.x) {
        const float* p = g.s1.out + (size_t(r) * g.s1.n_node + S1.loc[i]) * g.s1.wp;
        S1.a[i] = reinterpret_cast<const float4*>(p)[0];
        S1.b[i] = g.s1.wp >= 8 ? reinterpret_cast<const float4*>(p)[1] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < g.s2.n; i += blockDim.x) {
        const float* p = g.s2.out + (size_t(r) * g.s2.n_node + S2.loc[i]) * g.s2.wp;
        S2.a[i] = reinterpret_cast<const float4*>(p)[0];
        S2.b[i] = g.s2.wp >= 8 ? reinterpret_cast<const float4*>(p)[1] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// Per-replica prologue shared by the four kernels.  The row lengths (and, backward, the rows' sensitivities) are fetched
// into registers BEFORE the barrier that frees the staging area, so that their latency overlaps the staging loads instead
// of following them; the lengths go to E.start for the in-place scan.  RPT = rows per thread (ceil(n_rows / CTPB)).
template <int RPT, bool BWD>
__device__ __forceinline__ void replica_prologue(const IGraphDev& g, int r, const StagedGroup& S1, const StagedGroup& S2, int n_rows,
                                                 const int* __restrict__ cnt, const float* __restrict__ sens, float* sn,
                                                 float* acc, int n_acc, const EdgeScratch& E) {
    int plen[RPT];
    float psn[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int j = threadIdx.x + k * CTPB;
        plen[k] = j < n_rows ? cnt[size_t(r) * n_rows + j] : 0;
        psn[k] = (BWD && j < n_rows) ? sens[size_t(r) * n_rows + j] : 1.f;
    }
    stage_groups(g, r, S1, S2);   // (leading barrier: the previous replica's readers of sn / acc / E.start are done)
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int j = threadIdx.x + k * CTPB;
        if (j < n_rows) {
            if (BWD) sn[j] = psn[k];
            E.start[j] = psn[k] != 0.f ? plen[k] : 0;   // rows without sensitivity contribute nothing
        }
    }
    if (BWD) for (int i = threadIdx.x; i < n_acc; i += CTPB) acc[i] = 0.f;
    __syncthreads();
    scan_row_lengths_inplace(n_rows, E.start, E.wtot);
}
__device__ __forceinline__ void unpack8(const StagedGroup& S, int i, float* x) {
    float4 a = S.a[i], b = S.b[i];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
inline size_t staged_bytes(int n1, int n2, int n_tab) {
    return size_t(n1 + n2) * 2 * sizeof(float4) + sizeof(float) * ((n_tab + 3) & ~3) + sizeof(int) * size_t((2 * (n1 + n2) + 3) & ~3);
}

// shared-memory sizes and persistent grids (as many CTAs as stay resident, striding over the replicas)
struct CoverageLaunch {
    size_t smem_fwd = 0, smem_bwd = 0;
    int grid_fwd = 1, grid_bwd = 1, n_rows_max = 0, rpt = 2;   // rpt: rows per thread of the prologue, 2 / 4 / 8
    static int rows_per_thread(const IGraphHost& ig, const char* what) {
        const int n_rows = ig.need1 ? ig.n1 : ig.n2;
        if (n_rows > 8 * CTPB) throw std::string(what) + ": more than " + std::to_string(8 * CTPB) + " rows per replica";
        return n_rows <= 2 * CTPB ? 2 : (n_rows <= 4 * CTPB ? 4 : 8);
    }
    void init(Engine* e, const IGraphHost& ig, int nv_bwd, int n_sens, const void* kf, const void* kb, const char* what) {
        if ((ig.need1 && ig.K1 > EL_CAP) || (ig.need2 && ig.K2 > EL_CAP)) throw std::string(what) + ": neighbour capacity exceeds the edge chunk size";
        n_rows_max = std::max(ig.n1, ig.n2);
        rpt = rows_per_thread(ig, what);
        size_t staged = staged_bytes(ig.n1, ig.n2, ig.n_type1 * ig.n_type2 * ig.n_param);
        smem_fwd = staged + edge_scratch_bytes(n_rows_max, 1);
        smem_bwd = staged + edge_scratch_bytes(n_rows_max, nv_bwd) + sizeof(float) * n_sens;
        int lim = 0, n_sm = 0, occ_f = 0, occ_b = 0;
        UB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
        UB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device));
        if (smem_bwd > (size_t)lim) throw std::string(what) + ": system too large for the shared-memory kernels";
        UB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fwd));
        UB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd));
        UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, kf, CTPB, smem_fwd));
        UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, kb, CTPB, smem_bwd));
        grid_fwd = std::max(1, std::min(e->n_rep, n_sm * std::max(1, occ_f)));
        grid_bwd = std::max(1, std::min(e->n_rep, n_sm * std::max(1, occ_b)));
    }
};

// ================================================================================================ HBondCoverage
// forward: per bead (group 2) the coverage of every H/O site (group 1) in range
template <int RPT>
__global__ void __launch_bounds__(CTPB) k_hbond_coverage(IGraphDev g, QuadSplineShape q, float* __restrict__ out, int n_rep, int n_rows_max) {
    extern __shared__ float4 smem4[];
    StagedGroup S1, S2;
    float* table;
    const int n_tab = g.n_type1 * g.n_type2 * g.n_param;
    EdgeScratch E = carve_edge_scratch(carve_groups(g, smem4, S1, S2, table, n_tab), n_rows_max, 1);
    stage_table(g, S1, S2, table, n_tab);
    for (int r = blockIdx.x; r < n_rep; r += gridDim.x) {
        replica_prologue<RPT, false>(g, r, S1, S2, g.s2.n, g.cnt2, nullptr, nullptr, nullptr, 0, E);
        for_each_edge<1>(g.s2.n, g.nbr2 + size_t(r) * g.s2.n * g.K2, g.K2, E,
            [&](int j, int i, float* o) {
                float x1[8], x2[8], d1[7], d2[6];
                unpack8(S1, i, x1); unpack8(S2, j, x2);
                o[0] = hbond_coverage_edge(table + (S1.type[i] * g.n_type2 + S2.type[j]) * g.n_param, q, x1, x2, d1, d2);
            },
            [&](int j, int, const float* s) { out[size_t(r) * g.s2.n + j] = s[0]; });
    }
}
// backward: every edge is evaluated ONCE (rows = beads whose coverage has a non-zero sensitivity).  Bead side:
// sens[j] * sum_i dV/d(bead j), summed per row in a fixed order.  Site side: sens[j] * dV/d(site i) (7 components, last =
// d/d hb) goes into per-site accumulators in shared memory with atomics (a site has few partners; forces are summed with
// float atomics elsewhere on the path as well), flushed by one thread per site.
template <int RPT>
__global__ void __launch_bounds__(CTPB) k_hbond_coverage_deriv(IGraphDev g, QuadSplineShape q, const float* __restrict__ sens, int n_rep,
                                                               int n_rows_max) {
    extern __shared__ float4 smem4[];
    StagedGroup S1, S2;
    float* table;
    const int n_tab = g.n_type1 * g.n_type2 * g.n_param;
    EdgeScratch E = carve_edge_scratch(carve_groups(g, smem4, S1, S2, table, n_tab), n_rows_max, 6);
    float* sn = reinterpret_cast<float*>(E.wtot + 33);   // [n2] sens of this replica's beads
    float* acc1 = sn + g.s2.n;                            // [n1][7] site-side sums
    stage_table(g, S1, S2, table, n_tab);
    for (int r = blockIdx.x; r < n_rep; r += gridDim.x) {
        replica_prologue<RPT, true>(g, r, S1, S2, g.s2.n, g.cnt2, sens, sn, acc1, UB_COV_RED ? 0 : g.s1.n * 7, E);
        for_each_edge<6>(g.s2.n, g.nbr2 + size_t(r) * g.s2.n * g.K2, g.K2, E,
            [&](int j, int i, float* o) {
                float x1[8], x2[8], d1[7];
                unpack8(S1, i, x1); unpack8(S2, j, x2);
                hbond_coverage_edge(table + (S1.type[i] * g.n_type2 + S2.type[j]) * g.n_param, q, x1, x2, d1, o);
                const float sj = sn[j];
#if UB_COV_RED
                // two 16-byte reductions straight into the site's sens row (REDG.ADD.F32x4): a shared-memory float atomicAdd
                // would be a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN)
                float4* d = reinterpret_cast<float4*>(g.s1.sens + (size_t(r) * g.s1.n_node + S1.loc[i]) * g.s1.wp);
                atomicAdd(d, make_float4(sj * d1[0], sj * d1[1], sj * d1[2], sj * d1[3]));
                atomicAdd(d + 1, make_float4(sj * d1[4], sj * d1[5], sj * d1[6], 0.f));
#else
#pragma unroll
                for (int c = 0; c < 7; ++c) atomicAdd(&acc1[i * 7 + c], sj * d1[c]);
#endif
            },
            [&](int j, int c, const float* s) {
                if (!c) return;
                // reductions without a return value (RED): the thread does not wait for the row to come back from L2
                float* dst = g.s2.sens + (size_t(r) * g.s2.n_node + S2.loc[j]) * g.s2.wp;
                const float sj = sn[j];
                atomicAdd(reinterpret_cast<float4*>(dst), make_float4(sj * s[0], sj * s[1], sj * s[2], sj * s[3]));
                atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(sj * s[4], sj * s[5], 0.f, 0.f));
            });
#if !UB_COV_RED
        // (for_each_edge ends with a barrier: acc1 is complete)
        for (int i = threadIdx.x; i < g.s1.n; i += blockDim.x) {
            const float* s = acc1 + i * 7;
            if (s[0] == 0.f && s[1] == 0.f && s[2] == 0.f && s[6] == 0.f) continue;   // site without a partner (no table by site is kept)
            float* dst = g.s1.sens + (size_t(r) * g.s1.n_node + S1.loc[i]) * g.s1.wp;
#pragma unroll
            for (int k = 0; k < 7; ++k) atomicAdd(dst + k, s[k]);
        }
#endif
    }
}
// parameter derivative (interaction_graph.h:404-415 with hbond.cpp:278-283): sum over edges of
// sens[bead] * (1-hb)^2 * d(quadspline)/d(param); off the hot path, one thread per bead row, atomics into the table
__global__ void k_hbond_coverage_param_deriv(IGraphDev g, QuadSplineShape q, const float* __restrict__ sens, int r0, int r1,
                                             float* __restrict__ out) {
    const int r = r0 + blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1 || j >= g.s2.n) return;
    const float sj = sens[size_t(r) * g.s2.n + j];
    const int c = g.cnt2[size_t(r) * g.s2.n + j];
    if (!c || sj == 0.f) return;
    float x2[8];
    load8(elem_ptr(g.s2, r, j), x2);
    const unsigned short* row = g.nbr2 + (size_t(r) * g.s2.n + j) * g.K2;
    for (int k = 0; k < c; ++k) {
        const int i = row[k];
        float x1[8], val[16];
        int idx[16];
        load8(elem_ptr(g.s1, r, i), x1);
        const int tp = g.s1.type[i] * g.n_type2 + g.s2.type[j];
        quadspline_param_deriv(g.param + size_t(tp) * g.n_param, q, x1, x2, idx, val);
        const float w = sj * (1.f - x1[6]) * (1.f - x1[6]);
        for (int m = 0; m < 16; ++m) atomicAdd(out + size_t(tp) * g.n_param + idx[m], w * val[m]);
    }
}
struct HBondCoverage : CoordNode {
    IGraphHost ig;
    int nka = 15, nk = 12;
    float knot_spacing = 0.5f;
    CoverageLaunch launch;
    HBondCoverage(Engine&, const h5l::Node& g, CoordNode& hb, CoordNode& sc)
        : CoordNode((int)h5_dims(g, "index2", 1)[0], 1), ig(g, false, EXCL_SEQ2, 7, 6, &hb, &sc) {
        if (hb.wp != 8 || sc.wp != 8) throw std::string("hbond_coverage expects 8-float rows on both arguments");
        // knot counts are compile-time in the reference (bead_interaction.h:12-27); here they follow the table shape:
        // n_param = 2*n_knot_angular + 2*n_knot_radial with (angular, radial, spacing) of the three reference builds
        if (ig.n_param == 2 * 15 + 2 * 12) { nka = 15; nk = 12; knot_spacing = 0.5f; }
        else if (ig.n_param == 2 * 8 + 2 * 12) { nka = 8; nk = 12; knot_spacing = 1.f; }
        else if (ig.n_param == 2 * 8 + 2 * 7) { nka = 8; nk = 7; knot_spacing = 1.f; }
        else throw "unsupported hbond_coverage parameter count " + std::to_string(ig.n_param);
        ig.cutoff = float((nk - 2 - 1e-6) / double(1.f / knot_spacing));   // hbond.cpp:250-252
    }
    void finalize() override {
        ig.need1 = false;   // both kernels walk the rows of the beads (table 2)
        ig.allocate(engine);
        const int rpt = CoverageLaunch::rows_per_thread(ig, "hbond_coverage");
        kf = rpt == 2 ? k_hbond_coverage<2> : (rpt == 4 ? k_hbond_coverage<4> : k_hbond_coverage<8>);
        kb = rpt == 2 ? k_hbond_coverage_deriv<2> : (rpt == 4 ? k_hbond_coverage_deriv<4> : k_hbond_coverage_deriv<8>);
        launch.init(engine, ig, 6, ig.n2 + (UB_COV_RED ? 0 : 7 * ig.n1), (const void*)kf, (const void*)kb, "hbond_coverage");
    }
    void (*kf)(IGraphDev, QuadSplineShape, float*, int, int) = nullptr;
    void (*kb)(IGraphDev, QuadSplineShape, const float*, int, int) = nullptr;
    QuadSplineShape shape() const { QuadSplineShape q; q.nka = nka; q.nk = nk; q.inv_dx = 1.f / knot_spacing; q.inv_dtheta = (nka - 3) / 2.f; return q; }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        kf<<<launch.grid_fwd, CTPB, launch.smem_fwd, s>>>(ig.dev(), shape(), output, engine->n_rep, launch.n_rows_max);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        kb<<<launch.grid_bwd, CTPB, launch.smem_bwd, s>>>(ig.dev(), shape(), sens, engine->n_rep, launch.n_rows_max);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }
    std::vector<float> get_param_deriv(int replica) override {
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        DevBuf<float> acc;
        acc.upload(std::vector<float>(ig.h_param.size(), 0.f));
        const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? engine->n_rep : replica + 1;
        for (int a = r0; n_elem && a < r1; a += 32768) {
            const int b = std::min(r1, a + 32768);
            k_hbond_coverage_param_deriv<<<dim3((ig.n2 + 127) / 128, b - a), 128>>>(ig.dev(), shape(), sens, a, b, acc.p);
        }
        UB_CUDA(cudaDeviceSynchronize());
        return acc.download();
    }
    std::vector<float> get_value_by_name(int replica, const char* nm) override {
        if (std::string(nm) == "count_edges_by_type") return ig.count_edges_by_type(replica);
        throw std::string("Value ") + nm + " not implemented";
    }
};
RegisterNodeType<HBondCoverage, 2> coverage_node("hbond_coverage");

// ================================================================================================ EnvironmentCoverage
// forward: per CB (group 1) the weighted count of side-chain beads (group 2) in its cone
template <int RPT>
__global__ void __launch_bounds__(CTPB) k_env_coverage(IGraphDev g, float* __restrict__ out, int n_rep, int n_rows_max) {
    extern __shared__ float4 smem4[];
    StagedGroup S1, S2;
    float* table;
    const int n_tab = g.n_type1 * g.n_type2 * g.n_param;
    EdgeScratch E = carve_edge_scratch(carve_groups(g, smem4, S1, S2, table, n_tab), n_rows_max, 1);
    stage_table(g, S1, S2, table, n_tab);
    for (int r = blockIdx.x; r < n_rep; r += gridDim.x) {
        replica_prologue<RPT, false>(g, r, S1, S2, g.s1.n, g.cnt1, nullptr, nullptr, nullptr, 0, E);
        for_each_edge<1>(g.s1.n, g.nbr1 + size_t(r) * g.s1.n * g.K1, g.K1, E,
            [&](int i, int j, float* o) {
                float x1[8], d1[6], d2[4];
                unpack8(S1, i, x1);
                float4 v = S2.a[j];
                float x2[4] = {v.x, v.y, v.z, v.w};
                o[0] = environment_edge(table + (S1.type[i] * g.n_type2 + S2.type[j]) * g.n_param, x1, x2, d1, d2);
            },
            [&](int i, int, const float* s) { out[size_t(r) * g.s1.n + i] = s[0]; });
    }
}
// backward: every edge is evaluated once (rows = CBs whose coverage has a non-zero sensitivity); CB side summed per row in a
// fixed order, bead side (position + weight, 4 components) through shared-memory accumulators as in k_hbond_coverage_deriv
template <int RPT>
__global__ void __launch_bounds__(CTPB) k_env_coverage_deriv(IGraphDev g, const float* __restrict__ sens, int n_rep, int n_rows_max) {
    extern __shared__ float4 smem4[];
    StagedGroup S1, S2;
    float* table;
    const int n_tab = g.n_type1 * g.n_type2 * g.n_param;
    EdgeScratch E = carve_edge_scratch(carve_groups(g, smem4, S1, S2, table, n_tab), n_rows_max, 6);
    float* sn = reinterpret_cast<float*>(E.wtot + 33);   // [n1] sens of this replica's CB coverages
    float* acc2 = sn + g.s1.n;                            // [n2][4] bead-side sums
    stage_table(g, S1, S2, table, n_tab);
    for (int r = blockIdx.x; r < n_rep; r += gridDim.x) {
        replica_prologue<RPT, true>(g, r, S1, S2, g.s1.n, g.cnt1, sens, sn, acc2, UB_COV_RED ? 0 : g.s2.n * 4, E);
        for_each_edge<6>(g.s1.n, g.nbr1 + size_t(r) * g.s1.n * g.K1, g.K1, E,
            [&](int i, int j, float* o) {
                float x1[8], d2[4];
                unpack8(S1, i, x1);
                float4 v = S2.a[j];
                float x2[4] = {v.x, v.y, v.z, v.w};
                environment_edge(table + (S1.type[i] * g.n_type2 + S2.type[j]) * g.n_param, x1, x2, o, d2);
                const float si = sn[i];
#if UB_COV_RED
                atomicAdd(reinterpret_cast<float4*>(g.s2.sens + (size_t(r) * g.s2.n_node + S2.loc[j]) * g.s2.wp),
                          make_float4(si * d2[0], si * d2[1], si * d2[2], si * d2[3]));   // one REDG.ADD.F32x4 per edge
#else
#pragma unroll
                for (int c = 0; c < 4; ++c) atomicAdd(&acc2[j * 4 + c], si * d2[c]);
#endif
            },
            [&](int i, int c, const float* s) {
                if (!c) return;
                float* dst = g.s1.sens + (size_t(r) * g.s1.n_node + S1.loc[i]) * g.s1.wp;
                const float si = sn[i];
                atomicAdd(reinterpret_cast<float4*>(dst), make_float4(si * s[0], si * s[1], si * s[2], si * s[3]));
                atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(si * s[4], si * s[5], 0.f, 0.f));
            });
#if !UB_COV_RED
        for (int j = threadIdx.x; j < g.s2.n; j += blockDim.x) {
            if (acc2[4 * j] == 0.f && acc2[4 * j + 1] == 0.f && acc2[4 * j + 2] == 0.f && acc2[4 * j + 3] == 0.f) continue;   // bead outside every cone
            float* dst = g.s2.sens + (size_t(r) * g.s2.n_node + S2.loc[j]) * g.s2.wp;
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(dst + k, acc2[4 * j + k]);
        }
#endif
    }
}
struct EnvironmentCoverage : CoordNode {
    IGraphHost ig;
    CoverageLaunch launch;
    EnvironmentCoverage(Engine&, const h5l::Node& g, CoordNode& cb, CoordNode& wsc)
        : CoordNode((int)h5_dims(g, "index1", 1)[0], 1), ig(g, false, EXCL_SEQ2, 6, 4, &cb, &wsc) {
        if (ig.n_param != 4) throw std::string("environment_coverage expects 4 interaction parameters");
        if (cb.wp != 8 || wsc.wp != 4) throw std::string("environment_coverage expects (8,4)-float rows");
        float c = 0.f;   // environment.cpp:18-20 with compact_sigmoid_cutoff = 1/sharpness
        for (int t = 0; t < ig.n_type1 * ig.n_type2; ++t) c = std::max(c, ig.h_param[t * 4 + 0] + 1.f / ig.h_param[t * 4 + 1]);
        ig.cutoff = c;
    }
    void finalize() override {
        ig.need2 = false;   // both kernels walk the rows of the CBs (table 1)
        ig.allocate(engine);
        const int rpt = CoverageLaunch::rows_per_thread(ig, "environment_coverage");
        kf = rpt == 2 ? k_env_coverage<2> : (rpt == 4 ? k_env_coverage<4> : k_env_coverage<8>);
        kb = rpt == 2 ? k_env_coverage_deriv<2> : (rpt == 4 ? k_env_coverage_deriv<4> : k_env_coverage_deriv<8>);
        launch.init(engine, ig, 6, ig.n1 + (UB_COV_RED ? 0 : 4 * ig.n2), (const void*)kf, (const void*)kb, "environment_coverage");
    }
    void (*kf)(IGraphDev, float*, int, int) = nullptr;
    void (*kb)(IGraphDev, const float*, int, int) = nullptr;
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        kf<<<launch.grid_fwd, CTPB, launch.smem_fwd, s>>>(ig.dev(), output, engine->n_rep, launch.n_rows_max);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        kb<<<launch.grid_bwd, CTPB, launch.smem_bwd, s>>>(ig.dev(), sens, engine->n_rep, launch.n_rows_max);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }   // a cutoff change needs a new engine
    // the reference leaves this derivative unimplemented and returns zeros (environment.cpp:62-65)
    std::vector<float> get_param_deriv(int) override { return std::vector<float>(ig.h_param.size(), 0.f); }
};
RegisterNodeType<EnvironmentCoverage, 2> environment_coverage_node("environment_coverage");

}  // namespace
}  // namespace ub
