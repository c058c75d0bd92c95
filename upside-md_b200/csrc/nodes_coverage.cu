// hbond_coverage (reference hbond.cpp:241-286,371-414) and environment_coverage (environment.cpp:12-109).
//
// One CTA per replica stages both interaction groups (8 floats + type per element) and the parameter table in shared
// memory, then lane groups walk the ELL rows: the forward kernel sums pair values per target element, the backward
// kernel recomputes the pair term and gathers d/d(group 1) and d/d(group 2) in one launch - no atomics, fixed order.
#include <algorithm>
#include <cmath>

#include "igraph.cuh"

namespace ub {
namespace {

constexpr int CTPB = 256;
constexpr int CG = 8;   // lanes per element row

struct StagedGroup {
    float4* a;   // x,y,z,w0
    float4* b;   // w1..w4
    int* type;
};

// carve `n1`+`n2` staged elements and `n_tab` table floats out of dynamic shared memory and fill them
__device__ __forceinline__ void stage_groups(const IGraphDev& g, int r, float4* smem, StagedGroup& S1, StagedGroup& S2,
                                             float*& table, int n_tab) {
    S1.a = smem; S1.b = S1.a + g.s1.n;
    S2.a = S1.b + g.s1.n; S2.b = S2.a + g.s2.n;
    S1.type = reinterpret_cast<int*>(S2.b + g.s2.n);
    S2.type = S1.type + g.s1.n;
    table = reinterpret_cast<float*>(S2.type + g.s2.n);
    for (int i = threadIdx.x; i < g.s1.n; i += blockDim.x) {
        const float* p = elem_ptr(g.s1, r, i);
        S1.a[i] = reinterpret_cast<const float4*>(p)[0];
        S1.b[i] = g.s1.wp >= 8 ? reinterpret_cast<const float4*>(p)[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        S1.type[i] = g.s1.type[i];
    }
    for (int i = threadIdx.x; i < g.s2.n; i += blockDim.x) {
        const float* p = elem_ptr(g.s2, r, i);
        S2.a[i] = reinterpret_cast<const float4*>(p)[0];
        S2.b[i] = g.s2.wp >= 8 ? reinterpret_cast<const float4*>(p)[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        S2.type[i] = g.s2.type[i];
    }
    for (int i = threadIdx.x; i < n_tab; i += blockDim.x) table[i] = g.param[i];
    __syncthreads();
}
__device__ __forceinline__ void unpack8(const StagedGroup& S, int i, float* x) {
    float4 a = S.a[i], b = S.b[i];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
inline size_t staged_bytes(int n1, int n2, int n_tab) { return size_t(n1 + n2) * (2 * sizeof(float4) + sizeof(int)) + sizeof(float) * n_tab; }

// ================================================================================================ HBondCoverage
// forward: per bead (group 2) the coverage of every H/O site (group 1) in range
__global__ void __launch_bounds__(CTPB) k_hbond_coverage(IGraphDev g, QuadSplineShape q, float* __restrict__ out) {
    extern __shared__ float4 smem4[];
    const int r = blockIdx.x;
    StagedGroup S1, S2;
    float* table;
    stage_groups(g, r, smem4, S1, S2, table, g.n_type1 * g.n_type2 * g.n_param);
    const int lane = threadIdx.x % CG, n_grp = CTPB / CG;
    for (int j0 = 0; j0 < g.s2.n; j0 += n_grp) {
        int j = j0 + threadIdx.x / CG;
        float acc = 0.f;
        if (j < g.s2.n) {
            float x2[8];
            unpack8(S2, j, x2);
            int t2 = S2.type[j];
            const unsigned short* row = g.nbr2 + (size_t(r) * g.s2.n + j) * g.K2;
            int cnt = g.cnt2[size_t(r) * g.s2.n + j];
            for (int k = lane; k < cnt; k += CG) {
                int i = row[k];
                float x1[8], d1[7], d2[6];
                unpack8(S1, i, x1);
                acc += hbond_coverage_edge(table + (S1.type[i] * g.n_type2 + t2) * g.n_param, q, x1, x2, d1, d2);
            }
        }
        acc = group_sum<CG>(acc);
        if (j < g.s2.n && lane == 0) out[size_t(r) * g.s2.n + j] = acc;
    }
}
// backward: bead side sens[j] * sum_i dV/d(bead j); site side sum_j sens[j] * dV/d(site i) (7 components, last = d/d hb)
__global__ void __launch_bounds__(CTPB) k_hbond_coverage_deriv(IGraphDev g, QuadSplineShape q, const float* __restrict__ sens) {
    extern __shared__ float4 smem4[];
    const int r = blockIdx.x;
    StagedGroup S1, S2;
    float* table;
    stage_groups(g, r, smem4, S1, S2, table, g.n_type1 * g.n_type2 * g.n_param);
    const int lane = threadIdx.x % CG, n_grp = CTPB / CG;
    const float* sn = sens + size_t(r) * g.s2.n;
    for (int j0 = 0; j0 < g.s2.n; j0 += n_grp) {
        int j = j0 + threadIdx.x / CG;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float sj = 0.f;
        if (j < g.s2.n) {
            sj = sn[j];
            float x2[8];
            unpack8(S2, j, x2);
            int t2 = S2.type[j];
            const unsigned short* row = g.nbr2 + (size_t(r) * g.s2.n + j) * g.K2;
            int cnt = (sj != 0.f) ? g.cnt2[size_t(r) * g.s2.n + j] : 0;
            for (int k = lane; k < cnt; k += CG) {
                int i = row[k];
                float x1[8], d1[7], d2[6];
                unpack8(S1, i, x1);
                hbond_coverage_edge(table + (S1.type[i] * g.n_type2 + t2) * g.n_param, q, x1, x2, d1, d2);
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[c] += d2[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] = group_sum<CG>(acc[c]);
        if (j < g.s2.n && lane == 0 && sj != 0.f) {
            float* dst = elem_sens_ptr(g.s2, r, j);
#pragma unroll
            for (int c = 0; c < 6; ++c) dst[c] += sj * acc[c];
        }
    }
    for (int i0 = 0; i0 < g.s1.n; i0 += n_grp) {
        int i = i0 + threadIdx.x / CG;
        float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (i < g.s1.n) {
            float x1[8];
            unpack8(S1, i, x1);
            int t1 = S1.type[i];
            const unsigned short* row = g.nbr1 + (size_t(r) * g.s1.n + i) * g.K1;
            int cnt = g.cnt1[size_t(r) * g.s1.n + i];
            for (int k = lane; k < cnt; k += CG) {
                int j = row[k];
                float sj = sn[j];
                float x2[8], d1[7], d2[6];
                unpack8(S2, j, x2);
                hbond_coverage_edge(table + (t1 * g.n_type2 + S2.type[j]) * g.n_param, q, x1, x2, d1, d2);
#pragma unroll
                for (int c = 0; c < 7; ++c) acc[c] += sj * d1[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[c] = group_sum<CG>(acc[c]);
        if (i < g.s1.n && lane == 0) {
            float* dst = elem_sens_ptr(g.s1, r, i);
#pragma unroll
            for (int c = 0; c < 7; ++c) dst[c] += acc[c];
        }
    }
}
struct HBondCoverage : CoordNode {
    IGraphHost ig;
    int nka = 15, nk = 12;
    float knot_spacing = 0.5f;
    size_t smem = 0;
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
        ig.allocate(engine);
        smem = staged_bytes(ig.n1, ig.n2, ig.n_type1 * ig.n_type2 * ig.n_param);
        int lim = 0;
        UB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, engine->device));
        if (smem > (size_t)lim) throw std::string("hbond_coverage: system too large for the shared-memory kernels");
        UB_CUDA(cudaFuncSetAttribute(k_hbond_coverage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UB_CUDA(cudaFuncSetAttribute(k_hbond_coverage_deriv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    QuadSplineShape shape() const { QuadSplineShape q; q.nka = nka; q.nk = nk; q.inv_dx = 1.f / knot_spacing; q.inv_dtheta = (nka - 3) / 2.f; return q; }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        k_hbond_coverage<<<engine->n_rep, CTPB, smem, s>>>(ig.dev(), shape(), output);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_hbond_coverage_deriv<<<engine->n_rep, CTPB, smem, s>>>(ig.dev(), shape(), sens);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }
    std::vector<float> get_value_by_name(int replica, const char* nm) override {
        if (std::string(nm) == "count_edges_by_type") return ig.count_edges_by_type(replica);
        throw std::string("Value ") + nm + " not implemented";
    }
};
RegisterNodeType<HBondCoverage, 2> coverage_node("hbond_coverage");

// ================================================================================================ EnvironmentCoverage
__global__ void __launch_bounds__(CTPB) k_env_coverage(IGraphDev g, float* __restrict__ out) {
    extern __shared__ float4 smem4[];
    const int r = blockIdx.x;
    StagedGroup S1, S2;
    float* table;
    stage_groups(g, r, smem4, S1, S2, table, g.n_type1 * g.n_type2 * g.n_param);
    const int lane = threadIdx.x % CG, n_grp = CTPB / CG;
    for (int i0 = 0; i0 < g.s1.n; i0 += n_grp) {
        int i = i0 + threadIdx.x / CG;
        float acc = 0.f;
        if (i < g.s1.n) {
            float x1[8];
            unpack8(S1, i, x1);
            const float* p = table + S1.type[i] * g.n_type2 * g.n_param;
            const unsigned short* row = g.nbr1 + (size_t(r) * g.s1.n + i) * g.K1;
            int cnt = g.cnt1[size_t(r) * g.s1.n + i];
            for (int k = lane; k < cnt; k += CG) {
                int j = row[k];
                float4 v = S2.a[j];
                float x2[4] = {v.x, v.y, v.z, v.w}, d1[6], d2[4];
                acc += environment_edge(p + S2.type[j] * g.n_param, x1, x2, d1, d2);
            }
        }
        acc = group_sum<CG>(acc);
        if (i < g.s1.n && lane == 0) out[size_t(r) * g.s1.n + i] = acc;
    }
}
__global__ void __launch_bounds__(CTPB) k_env_coverage_deriv(IGraphDev g, const float* __restrict__ sens) {
    extern __shared__ float4 smem4[];
    const int r = blockIdx.x;
    StagedGroup S1, S2;
    float* table;
    stage_groups(g, r, smem4, S1, S2, table, g.n_type1 * g.n_type2 * g.n_param);
    const int lane = threadIdx.x % CG, n_grp = CTPB / CG;
    const float* sn = sens + size_t(r) * g.s1.n;
    for (int i0 = 0; i0 < g.s1.n; i0 += n_grp) {
        int i = i0 + threadIdx.x / CG;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float si = 0.f;
        if (i < g.s1.n) {
            si = sn[i];
            float x1[8];
            unpack8(S1, i, x1);
            const float* p = table + S1.type[i] * g.n_type2 * g.n_param;
            const unsigned short* row = g.nbr1 + (size_t(r) * g.s1.n + i) * g.K1;
            int cnt = (si != 0.f) ? g.cnt1[size_t(r) * g.s1.n + i] : 0;
            for (int k = lane; k < cnt; k += CG) {
                int j = row[k];
                float4 v = S2.a[j];
                float x2[4] = {v.x, v.y, v.z, v.w}, d1[6], d2[4];
                environment_edge(p + S2.type[j] * g.n_param, x1, x2, d1, d2);
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[c] += d1[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] = group_sum<CG>(acc[c]);
        if (i < g.s1.n && lane == 0 && si != 0.f) {
            float* dst = elem_sens_ptr(g.s1, r, i);
#pragma unroll
            for (int c = 0; c < 6; ++c) dst[c] += si * acc[c];
        }
    }
    for (int j0 = 0; j0 < g.s2.n; j0 += n_grp) {
        int j = j0 + threadIdx.x / CG;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < g.s2.n) {
            float4 v = S2.a[j];
            float x2[4] = {v.x, v.y, v.z, v.w};
            int t2 = S2.type[j];
            const unsigned short* row = g.nbr2 + (size_t(r) * g.s2.n + j) * g.K2;
            int cnt = g.cnt2[size_t(r) * g.s2.n + j];
            for (int k = lane; k < cnt; k += CG) {
                int i = row[k];
                float si = sn[i];
                float x1[8], d1[6], d2[4];
                unpack8(S1, i, x1);
                environment_edge(table + (S1.type[i] * g.n_type2 + t2) * g.n_param, x1, x2, d1, d2);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] += si * d2[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = group_sum<CG>(acc[c]);
        if (j < g.s2.n && lane == 0) {
            float4* dst = reinterpret_cast<float4*>(elem_sens_ptr(g.s2, r, j));
            float4 o = *dst;
            o.x += acc[0]; o.y += acc[1]; o.z += acc[2]; o.w += acc[3];
            *dst = o;
        }
    }
}
struct EnvironmentCoverage : CoordNode {
    IGraphHost ig;
    size_t smem = 0;
    EnvironmentCoverage(Engine&, const h5l::Node& g, CoordNode& cb, CoordNode& wsc)
        : CoordNode((int)h5_dims(g, "index1", 1)[0], 1), ig(g, false, EXCL_SEQ2, 6, 4, &cb, &wsc) {
        if (ig.n_param != 4) throw std::string("environment_coverage expects 4 interaction parameters");
        if (cb.wp != 8 || wsc.wp != 4) throw std::string("environment_coverage expects (8,4)-float rows");
        float c = 0.f;   // environment.cpp:18-20 with compact_sigmoid_cutoff = 1/sharpness
        for (int t = 0; t < ig.n_type1 * ig.n_type2; ++t) c = std::max(c, ig.h_param[t * 4 + 0] + 1.f / ig.h_param[t * 4 + 1]);
        ig.cutoff = c;
    }
    void finalize() override {
        ig.allocate(engine);
        smem = staged_bytes(ig.n1, ig.n2, ig.n_type1 * ig.n_type2 * ig.n_param);
        int lim = 0;
        UB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, engine->device));
        if (smem > (size_t)lim) throw std::string("environment_coverage: system too large for the shared-memory kernels");
        UB_CUDA(cudaFuncSetAttribute(k_env_coverage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UB_CUDA(cudaFuncSetAttribute(k_env_coverage_deriv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        ig.build(s);
        k_env_coverage<<<engine->n_rep, CTPB, smem, s>>>(ig.dev(), output);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_env_coverage_deriv<<<engine->n_rep, CTPB, smem, s>>>(ig.dev(), sens);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }   // a cutoff change needs a new engine
};
RegisterNodeType<EnvironmentCoverage, 2> environment_coverage_node("environment_coverage");

}  // namespace
}  // namespace ub
