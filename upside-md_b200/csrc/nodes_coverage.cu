// hbond_coverage (reference hbond.cpp:241-286,371-414) and environment_coverage (environment.cpp:12-109).
//
// A warp owns 32 ELL rows and deals their edges to its lanes (WarpEdges, edgelist.cuh); operands are gathered from global
// memory.  The forward kernels sum pair values per row; the backward kernels recompute the pair term once per edge, return
// the row's half through segmented shuffles and send the partner's half as 16-byte reductions.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "edgelist.cuh"
#include "igraph.cuh"

namespace ub {
namespace {

// The sparse coverage graphs have a few hundred to a thousand edges per replica.  (A CTA per replica that staged both groups
// and the parameter table in shared memory and ran one thread per edge spent most of its time in per-replica fixed costs -
// staging 600 elements, five barriers, scans: 57-72 us forward and 86-116 us backward per launch against 39-63 / 69-94 us
// here; plain thread-per-row kernels were no faster than the staged ones because two rows in three are empty.)  A warp owns 32 rows,
// its lanes share the rows' edges evenly, operands come straight from global memory (L2-resident), row sums return to the
// row's lane by segmented shuffles in a fixed order, and the backward pass sends the partner's half of every edge to the
// partner's sens row as 16-byte reductions.  No shared memory, no barriers.  Grid (rows/128, replicas).
constexpr int DTPB = 128;
__global__ void __launch_bounds__(DTPB) k_hbond_coverage(IGraphDev g, QuadSplineShape q, float* __restrict__ out) {
    const int r = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = j < g.s2.n;
    const WarpEdges W(in ? g.cnt2[size_t(r) * g.s2.n + j] : 0);
    const int row0 = j - W.lane;
    float sum = 0.f;
    for (int it = 0; it < W.rounds(); ++it) {
        int s, k;
        const bool valid = W.edge(it, s, k);
        float v = 0.f;
        if (valid) {
            const int jr = row0 + s, i = g.nbr2[(size_t(r) * g.s2.n + jr) * g.K2 + k];
            float x1[8], x2[8], d1[7], d2[6];
            load8(elem_ptr(g.s1, r, i), x1);
            load8(elem_ptr(g.s2, r, jr), x2);
            v = hbond_coverage_edge(g.param + (g.s1.type[i] * g.n_type2 + g.s2.type[jr]) * g.n_param, q, x1, x2, d1, d2);
        }
        sum += W.row_sum(it, v, s, valid);
    }
    if (in) out[size_t(r) * g.s2.n + j] = sum;
}
__global__ void __launch_bounds__(DTPB) k_hbond_coverage_deriv(IGraphDev g, QuadSplineShape q, const float* __restrict__ sens) {
    const int r = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = j < g.s2.n;
    const float sj = in ? sens[size_t(r) * g.s2.n + j] : 0.f;
    const WarpEdges W(sj != 0.f ? g.cnt2[size_t(r) * g.s2.n + j] : 0);   // rows without sensitivity contribute nothing
    const int row0 = j - W.lane;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int it = 0; it < W.rounds(); ++it) {
        int s, k;
        const bool valid = W.edge(it, s, k);
        float d2[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const int jr = row0 + s, i = g.nbr2[(size_t(r) * g.s2.n + jr) * g.K2 + k];
            const float sr = sens[size_t(r) * g.s2.n + jr];
            float x1[8], x2[8], d1[7];
            load8(elem_ptr(g.s1, r, i), x1);
            load8(elem_ptr(g.s2, r, jr), x2);
            hbond_coverage_edge(g.param + (g.s1.type[i] * g.n_type2 + g.s2.type[jr]) * g.n_param, q, x1, x2, d1, d2);
            float4* d = reinterpret_cast<float4*>(elem_sens_ptr(g.s1, r, i));
            atomicAdd(d, make_float4(sr * d1[0], sr * d1[1], sr * d1[2], sr * d1[3]));
            atomicAdd(d + 1, make_float4(sr * d1[4], sr * d1[5], sr * d1[6], 0.f));
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) acc[m] += W.row_sum(it, d2[m], s, valid);
    }
    if (W.c) {
        float4* dst = reinterpret_cast<float4*>(elem_sens_ptr(g.s2, r, j));
        atomicAdd(dst, make_float4(sj * acc[0], sj * acc[1], sj * acc[2], sj * acc[3]));
        atomicAdd(dst + 1, make_float4(sj * acc[4], sj * acc[5], 0.f, 0.f));
    }
}
__global__ void __launch_bounds__(DTPB) k_env_coverage(IGraphDev g, float* __restrict__ out) {
    const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < g.s1.n;
    const WarpEdges W(in ? g.cnt1[size_t(r) * g.s1.n + i] : 0);
    const int row0 = i - W.lane;
    float sum = 0.f;
    for (int it = 0; it < W.rounds(); ++it) {
        int s, k;
        const bool valid = W.edge(it, s, k);
        float v = 0.f;
        if (valid) {
            const int ir = row0 + s, j = g.nbr1[(size_t(r) * g.s1.n + ir) * g.K1 + k];
            float x1[8], d1[6], d2[4];
            load8(elem_ptr(g.s1, r, ir), x1);
            const float4 w = *reinterpret_cast<const float4*>(elem_ptr(g.s2, r, j));
            const float x2[4] = {w.x, w.y, w.z, w.w};
            v = environment_edge(g.param + (g.s1.type[ir] * g.n_type2 + g.s2.type[j]) * g.n_param, x1, x2, d1, d2);
        }
        sum += W.row_sum(it, v, s, valid);
    }
    if (in) out[size_t(r) * g.s1.n + i] = sum;
}
__global__ void __launch_bounds__(DTPB) k_env_coverage_deriv(IGraphDev g, const float* __restrict__ sens) {
    const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < g.s1.n;
    const float si = in ? sens[size_t(r) * g.s1.n + i] : 0.f;
    const WarpEdges W(si != 0.f ? g.cnt1[size_t(r) * g.s1.n + i] : 0);
    const int row0 = i - W.lane;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int it = 0; it < W.rounds(); ++it) {
        int s, k;
        const bool valid = W.edge(it, s, k);
        float d1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const int ir = row0 + s, j = g.nbr1[(size_t(r) * g.s1.n + ir) * g.K1 + k];
            const float sr = sens[size_t(r) * g.s1.n + ir];
            float x1[8], d2[4];
            load8(elem_ptr(g.s1, r, ir), x1);
            const float4 w = *reinterpret_cast<const float4*>(elem_ptr(g.s2, r, j));
            const float x2[4] = {w.x, w.y, w.z, w.w};
            environment_edge(g.param + (g.s1.type[ir] * g.n_type2 + g.s2.type[j]) * g.n_param, x1, x2, d1, d2);
            atomicAdd(reinterpret_cast<float4*>(elem_sens_ptr(g.s2, r, j)), make_float4(sr * d2[0], sr * d2[1], sr * d2[2], sr * d2[3]));
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) acc[m] += W.row_sum(it, d1[m], s, valid);
    }
    if (W.c) {
        float4* dst = reinterpret_cast<float4*>(elem_sens_ptr(g.s1, r, i));
        atomicAdd(dst, make_float4(si * acc[0], si * acc[1], si * acc[2], si * acc[3]));
        atomicAdd(dst + 1, make_float4(si * acc[4], si * acc[5], 0.f, 0.f));
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
    }
    QuadSplineShape shape() const { QuadSplineShape q; q.nka = nka; q.nk = nk; q.inv_dx = 1.f / knot_spacing; q.inv_dtheta = (nka - 3) / 2.f; return q; }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        k_hbond_coverage<<<dim3((ig.n2 + DTPB - 1) / DTPB, engine->n_rep), DTPB, 0, s>>>(ig.dev(), shape(), output);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_hbond_coverage_deriv<<<dim3((ig.n2 + DTPB - 1) / DTPB, engine->n_rep), DTPB, 0, s>>>(ig.dev(), shape(), sens);
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
struct EnvironmentCoverage : CoordNode {
    IGraphHost ig;
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
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        k_env_coverage<<<dim3((ig.n1 + DTPB - 1) / DTPB, engine->n_rep), DTPB, 0, s>>>(ig.dev(), output);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_env_coverage_deriv<<<dim3((ig.n1 + DTPB - 1) / DTPB, engine->n_rep), DTPB, 0, s>>>(ig.dev(), sens);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }   // a cutoff change needs a new engine
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // environment.cpp:77-82 (extensive only)
        if (level < 2) return;
        out.push_back({"environment_coverage", {(uint64_t)n_elem}, false, [this](int r) {
            auto o = host_rows(output, r);
            std::vector<float> v(n_elem);
            for (int i = 0; i < n_elem; ++i) v[i] = o[size_t(i) * wp];
            return v;
        }});
    }
    // the reference leaves this derivative unimplemented and returns zeros (environment.cpp:62-65)
    std::vector<float> get_param_deriv(int) override { return std::vector<float>(ig.h_param.size(), 0.f); }
};
RegisterNodeType<EnvironmentCoverage, 2> environment_coverage_node("environment_coverage");

}  // namespace
}  // namespace ub
