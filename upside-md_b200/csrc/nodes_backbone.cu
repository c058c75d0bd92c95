// backbone_pairs (reference backbone_steric.cpp:38-147) and membrane_potential (membrane_potential.cpp:105-152).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "igraph.cuh"
#include "spline_fit.h"

namespace ub {
namespace {

constexpr int TPB = 128;
constexpr int G = 8;

// ================================================================================================ BackbonePairs
// Residue pairs within dist_cutoff of each other's frame origin and |i-j| >= 2, then all N/CA/C/CB atom pairs with
// the compact-sigmoid wall 4*sigma_c(r^2-9, 1/0.3).  Gather form: every residue sums the force and torque it receives
// from all its partners, so the affine sens needs no atomics; the energy counts each pair once (j > i).
__global__ void k_bb_atoms(const float* __restrict__ affine, float* __restrict__ atoms, const int* __restrict__ residue,
                           const float* __restrict__ ref_pos, int n, int n_aff) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    const float* aff = affine + (size_t(r) * n_aff + residue[i]) * 8;
    float4 a0 = reinterpret_cast<const float4*>(aff)[0], a1 = reinterpret_cast<const float4*>(aff)[1];
    f3 t = mk3(a0.x, a0.y, a0.z);
    float q[4] = {a0.w, a1.x, a1.y, a1.z}, U[9];
    quat_to_rot(U, q);
    float4* o = reinterpret_cast<float4*>(atoms) + (size_t(r) * n + i) * 4;
    for (int a = 0; a < 4; ++a) {
        f3 x = rot_apply(U, ld3(ref_pos + (i * 4 + a) * 3)) + t;
        o[a] = make_float4(x.x, x.y, x.z, 0.f);
    }
}
__global__ void k_backbone_pairs(IGraphSide S, const unsigned short* __restrict__ nbr, const int* __restrict__ cnt, int K,
                                 const float* __restrict__ atoms, const int* __restrict__ n_atom,
                                 float* __restrict__ pot, int want_pot) {
    __shared__ float sc[32];
    int r = blockIdx.y;
    int i = (blockIdx.x * blockDim.x + threadIdx.x) / G, lane = threadIdx.x % G;
    bool active = i < S.n;
    const float cutoff2 = 3.f * 3.f + 0.1f * 3.f;
    const float sharp = 1.f / (3.0f * 0.10f);
    f3 d = mk3(0.f, 0.f, 0.f), tq = d;
    float en = 0.f;
    if (active) {
        const float4* ai = reinterpret_cast<const float4*>(atoms) + (size_t(r) * S.n + i) * 4;
        const float* aff = S.out + (size_t(r) * S.n_node + S.loc[i]) * S.wp;
        f3 ti = mk3(aff[0], aff[1], aff[2]);
        int na_i = n_atom[i];
        f3 xi[4];
        for (int a = 0; a < 4; ++a) { float4 v = ai[a]; xi[a] = mk3(v.x, v.y, v.z); }
        const unsigned short* row = nbr + (size_t(r) * S.n + i) * K;
        int c = cnt[size_t(r) * S.n + i];
        for (int k = lane; k < c; k += G) {
            int j = row[k];
            const float4* aj = reinterpret_cast<const float4*>(atoms) + (size_t(r) * S.n + j) * 4;
            int na_j = n_atom[j];
            for (int b = 0; b < na_j; ++b) {
                float4 v = aj[b];
                f3 xj = mk3(v.x, v.y, v.z);
                for (int a = 0; a < na_i; ++a) {
                    f3 rv = xi[a] - xj;
                    float r2 = mag2(rv);
                    if (r2 > cutoff2) continue;
                    float val, der;
                    compact_sigmoid(r2 - 9.f, sharp, val, der);
                    f3 g = (2.f * 4.f * der) * rv;
                    d += g;
                    tq += cross(xi[a] - ti, g);
                    if (j > i) en += 4.f * val;
                }
            }
        }
    }
    d.x = group_sum<G>(d.x); d.y = group_sum<G>(d.y); d.z = group_sum<G>(d.z);
    tq.x = group_sum<G>(tq.x); tq.y = group_sum<G>(tq.y); tq.z = group_sum<G>(tq.z);
    if (active && lane == 0) {
        float* s = S.sens + (size_t(r) * S.n_node + S.loc[i]) * S.wp;
        add3(s, d);
        add3(s + 3, tq);
    }
    if (want_pot) {
        en = block_sum(en, sc);
        if (threadIdx.x == 0) atomicAdd(pot + r, en);
    }
}
// Fused hot-path form: one CTA per replica places the backbone atoms of every residue in shared memory, tests all residue
// pairs (i < j, |id_i - id_j| >= 2) against the frame-origin cutoff and compacts the survivors into a shared-memory list
// (integer atomics), then runs one thread per listed pair: the 4x4 atom pairs are evaluated ONCE, force and torque go to
// both residues' affine sens rows as 16-byte reductions.  Pairs beyond the origin cutoff cannot contain an atom pair in
// range, so the result does not depend on the list; the reference-ordered list for the accessor is built on demand.
constexpr int BB_TPB = 128;
__global__ void __launch_bounds__(BB_TPB) k_backbone_fused(IGraphSide S, const float* __restrict__ ref_pos, const int* __restrict__ n_atom,
                                                            float origin_cutoff2, int cap, float* __restrict__ pot, int want_pot,
                                                            int* __restrict__ error_flag) {
    extern __shared__ float4 bb_smem[];
    __shared__ long long sc[BB_TPB / 32];
    __shared__ int n_list;
    const int r = blockIdx.x, n = S.n, tid = threadIdx.x;
    float4* atoms = bb_smem;            // [n][4]
    float4* org = atoms + 4 * n;        // [n] frame origin, w = id
    unsigned* list = reinterpret_cast<unsigned*>(org + n);   // [cap] i << 16 | j
    if (tid == 0) n_list = 0;
    for (int i = tid; i < n; i += BB_TPB) {
        const float* aff = S.out + (size_t(r) * S.n_node + S.loc[i]) * S.wp;
        const float4 a0 = reinterpret_cast<const float4*>(aff)[0], a1 = reinterpret_cast<const float4*>(aff)[1];
        const f3 t = mk3(a0.x, a0.y, a0.z);
        float q[4] = {a0.w, a1.x, a1.y, a1.z}, U[9];
        quat_to_rot(U, q);
        for (int a = 0; a < 4; ++a) {
            const f3 x = rot_apply(U, ld3(ref_pos + (i * 4 + a) * 3)) + t;
            atoms[4 * i + a] = make_float4(x.x, x.y, x.z, 0.f);
        }
        org[i] = make_float4(t.x, t.y, t.z, __int_as_float(S.id[i]));
    }
    __syncthreads();
    {   // phase 1: a warp per row i, its lanes over the partners j > i
        const int lane = tid & 31, w = tid >> 5;
        for (int i = w; i < n; i += BB_TPB / 32) {
            const float4 oi = org[i];
            for (int j = i + 1 + lane; j < n; j += 32) {
                const float4 oj = org[j];
                const float dx = oi.x - oj.x, dy = oi.y - oj.y, dz = oi.z - oj.z;
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d2 < origin_cutoff2 && acceptable_id_pair(EXCL_SEQ1, __float_as_int(oi.w), __float_as_int(oj.w))) {
                    const int slot = atomicAdd(&n_list, 1);
                    if (slot < cap) list[slot] = unsigned(i) << 16 | unsigned(j);
                }
            }
        }
    }
    __syncthreads();
    if (n_list > cap && tid == 0) atomicExch(error_flag, 1);
    const int m = min(n_list, cap);
    const float cutoff2 = 3.f * 3.f + 0.1f * 3.f;
    const float sharp = 1.f / (3.0f * 0.10f);
    // the list order varies from run to run (slots come from an atomic counter): the energy is summed in 2^-32 fixed point,
    // where addition is associative, so identical replicas still give bit-identical energies
    long long en_fx = 0;
    for (int e = tid; e < m; e += BB_TPB) {   // phase 2: one thread per residue pair
        const int i = list[e] >> 16, j = list[e] & 0xffffu;
        const int na_i = n_atom[i], na_j = n_atom[j];
        const float4 oi = org[i], oj = org[j];
        f3 fi = mk3(0.f, 0.f, 0.f), ti = fi, tj = fi;
        float en = 0.f;
        bool any = false;
        for (int b = 0; b < na_j; ++b) {
            const float4 vj = atoms[4 * j + b];
            const f3 xj = mk3(vj.x, vj.y, vj.z);
            for (int a = 0; a < na_i; ++a) {
                const float4 vi = atoms[4 * i + a];
                const f3 xi = mk3(vi.x, vi.y, vi.z);
                const f3 rv = xi - xj;
                const float r2 = mag2(rv);
                if (r2 > cutoff2) continue;
                float val, der;
                compact_sigmoid(r2 - 9.f, sharp, val, der);
                const f3 g = (2.f * 4.f * der) * rv;   // force on atom a of residue i; -g on atom b of residue j
                fi += g;
                ti += cross(xi - mk3(oi.x, oi.y, oi.z), g);
                tj -= cross(xj - mk3(oj.x, oj.y, oj.z), g);
                en += 4.f * val;
                any = true;
            }
        }
        if (any) {
            en_fx += __float2ll_rn(en * 4294967296.f);
            float4* si = reinterpret_cast<float4*>(S.sens + (size_t(r) * S.n_node + S.loc[i]) * S.wp);
            float4* sj = reinterpret_cast<float4*>(S.sens + (size_t(r) * S.n_node + S.loc[j]) * S.wp);
            atomicAdd(si, make_float4(fi.x, fi.y, fi.z, ti.x));
            atomicAdd(si + 1, make_float4(ti.y, ti.z, 0.f, 0.f));
            atomicAdd(sj, make_float4(-fi.x, -fi.y, -fi.z, tj.x));
            atomicAdd(sj + 1, make_float4(tj.y, tj.z, 0.f, 0.f));
        }
    }
    if (want_pot) {
#pragma unroll
        for (int o = 16; o; o >>= 1) en_fx += __shfl_down_sync(UB_FULL_MASK, en_fx, o);
        if ((tid & 31) == 0) sc[tid >> 5] = en_fx;
        __syncthreads();
        if (tid == 0) {
            long long tot = 0;
            for (int w = 0; w < BB_TPB / 32; ++w) tot += sc[w];
            atomicAdd(pot + r, float(double(tot) * (1.0 / 4294967296.0)));
        }
    }
}

struct BackbonePairs : PotentialNode {
    CoordNode& alignment;
    int n_residue, K = 0;
    float dist_cutoff;
    std::vector<int> residue, id, n_atom;
    DevBuf<int> d_residue, d_id, d_n_atom, cnt;
    DevBuf<float> d_ref, atoms;
    DevBuf<unsigned short> nbr;
    BackbonePairs(Engine&, const h5l::Node& g, CoordNode& alignment_) : alignment(alignment_) {
        check_elem_width(alignment, 7);
        n_residue = (int)h5_dims(g, "id", 1)[0];
        h5_check_size(g, "n_atom", {(uint64_t)n_residue});
        h5_check_size(g, "ref_pos", {(uint64_t)n_residue, 4, 3});
        id = h5_read<int>(g, "id");
        residue = id;   // the reference uses `id` both as the affine index and as the exclusion id (:66)
        n_atom = h5_read<int>(g, "n_atom");
        auto ref = h5_read<float>(g, "ref_pos");
        float max_dev = 0.f;
        for (int nr = 0; nr < n_residue; ++nr) {
            if (residue[nr] < 0 || residue[nr] >= alignment.n_elem) throw std::string("residue index out of range");
            if (n_atom[nr] < 0 || n_atom[nr] > 4) throw std::string("n_atom must be in [0,4]");
            for (int a = 0; a < n_atom[nr]; ++a) {
                const float* p = &ref[(nr * 4 + a) * 3];
                max_dev = std::max(max_dev, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
            }
            for (int a = n_atom[nr]; a < 4; ++a) for (int d = 0; d < 3; ++d) ref[(nr * 4 + a) * 3 + d] = 0.f;   // NaN rows (GLY CB)
        }
        dist_cutoff = 2 * max_dev + sqrtf(3.f * 3.f + 0.1f * 3.f);
        d_residue.upload(residue); d_id.upload(id); d_n_atom.upload(n_atom); d_ref.upload(ref);
    }
    int list_cap = 0;
    size_t smem_fused = 0;
    bool fused = false;
    void finalize() override {
        double k = 8. + 0.55 * double(dist_cutoff) * dist_cutoff * dist_cutoff / 4.;   // residues, not beads
        K = std::max(1, (int)std::min<double>(n_residue, std::ceil(k)));
        nbr.alloc(size_t(engine->n_rep) * n_residue * K);
        cnt.alloc(size_t(engine->n_rep) * n_residue);
        atoms.alloc(size_t(engine->n_rep) * n_residue * 16);
        // fused kernel: atoms + origins + the pair list (every pair once) of one replica in shared memory
        list_cap = std::max(64, (n_residue * K + 1) / 2);
        smem_fused = sizeof(float4) * 5 * size_t(n_residue) + sizeof(unsigned) * size_t(list_cap);
        int lim = 0;
        UB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, engine->device));
        fused = n_residue < 65536 && smem_fused <= (size_t)lim && !getenv("UPSIDE_B200_BACKBONE_GATHER");
        if (fused) UB_CUDA(cudaFuncSetAttribute(k_backbone_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused));
    }
    IGraphSide side() const {
        return IGraphSide{alignment.output, alignment.sens, alignment.n_elem, alignment.wp, d_residue.p, nullptr, d_id.p, n_residue};
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_residue) return;
        IGraphSide S = side();
        if (fused) {
            k_backbone_fused<<<engine->n_rep, BB_TPB, smem_fused, s>>>(S, d_ref.p, d_n_atom.p, dist_cutoff * dist_cutoff, list_cap, potential,
                                                                       mode == PotentialAndDerivMode, engine->error_flag.p);
            return;
        }
        constexpr int TILE = 128;
        k_pairlist<TILE><<<dim3((n_residue + TILE - 1) / TILE, engine->n_rep), TILE, 0, s>>>(
            S, S, nbr.p, cnt.p, K, dist_cutoff * dist_cutoff, EXCL_SEQ1, 1, 1, engine->error_flag.p, nullptr, nullptr, 0);
        k_bb_atoms<<<dim3((n_residue + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(alignment.output, atoms.p, d_residue.p,
                                                                                    d_ref.p, n_residue, alignment.n_elem);
        k_backbone_pairs<<<dim3((n_residue * G + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(
            S, nbr.p, cnt.p, K, atoms.p, d_n_atom.p, potential, mode == PotentialAndDerivMode);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override {
        engine->sync_and_check();
        if (fused) {   // the hot path keeps no list in global memory: build the reference's list for the accessor
            constexpr int TILE = 128;
            k_pairlist<TILE><<<dim3((n_residue + TILE - 1) / TILE, engine->n_rep), TILE>>>(
                side(), side(), nbr.p, cnt.p, K, dist_cutoff * dist_cutoff, EXCL_SEQ1, 1, 1, engine->error_flag.p, nullptr, nullptr, 0);
            engine->sync_and_check();
        }
        std::vector<unsigned short> rows(size_t(n_residue) * K);
        std::vector<int> c(n_residue);
        UB_CUDA(cudaMemcpy(rows.data(), nbr.p + size_t(replica) * n_residue * K, rows.size() * 2, cudaMemcpyDeviceToHost));
        UB_CUDA(cudaMemcpy(c.data(), cnt.p + size_t(replica) * n_residue, c.size() * 4, cudaMemcpyDeviceToHost));
        i1.clear(); i2.clear();
        for (int i = 0; i < n_residue; ++i)
            for (int k = 0; k < c[i]; ++k) if (i < rows[size_t(i) * K + k]) { i1.push_back(i); i2.push_back(rows[size_t(i) * K + k]); }
        sort_reference_order(i1, i2);
        return true;
    }
};
RegisterNodeType<BackbonePairs, 1> backbone_pairs_node("backbone_pairs");

// ================================================================================================ MembranePotential
__device__ __forceinline__ void clamped1d(const float* __restrict__ coeff, const float* __restrict__ left,
                                          const float* __restrict__ right, int nx, int layer, float x, float& val, float& der) {
    // LayeredClampedSpline1D<1>::evaluate_value_and_deriv, spline.h:495-515
    if (x >= nx - 1) { der = 0.f; val = right[layer]; }
    else if (x <= 0) { der = 0.f; val = left[layer]; }
    else {
        int b = (int)x;
        float f = x - b;
        const float* c = coeff + (size_t(layer) * (nx - 1) + b) * 4;
        der = c[1] + 2.f * f * c[2] + 3.f * f * f * c[3];
        val = c[0] + f * c[1] + f * f * c[2] + f * f * f * c[3];
    }
}
struct MembraneDev {
    const float *cb, *env, *hb;
    float *cb_sens, *env_sens, *hb_sens;
    int n_cb, wp_cb, n_env, wp_env, n_hb;
    const int *cb_index, *env_index, *restype;
    const float *cov_mid, *cov_sharp;
    const float *cb_coeff, *cb_left, *cb_right, *uhb_coeff, *uhb_left, *uhb_right;
    int cb_nx, uhb_nx, n_elem, n_donor, n_virtual;
    float cb_shift, cb_scale, uhb_shift, uhb_scale;
};
__global__ void k_membrane(MembraneDev M, float* __restrict__ pot, int want_pot) {
    __shared__ float sc[32];
    int t = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float en = 0.f;
    if (t < M.n_elem) {
        int ci = M.cb_index[t], ei = M.env_index[t], rt = M.restype[t];
        float z = M.cb[(size_t(r) * M.n_cb + ci) * M.wp_cb + 2];
        float v, d;
        clamped1d(M.cb_coeff, M.cb_left, M.cb_right, M.cb_nx, rt, (z + M.cb_shift) * M.cb_scale, v, d);
        d *= M.cb_scale;
        float sv, sd;
        compact_sigmoid(M.env[(size_t(r) * M.n_env + ei) * M.wp_env] - M.cov_mid[rt], M.cov_sharp[rt], sv, sd);
        en = v * sv;
        atomicAdd(&M.cb_sens[(size_t(r) * M.n_cb + ci) * M.wp_cb + 2], d * sv);
        atomicAdd(&M.env_sens[(size_t(r) * M.n_env + ei) * M.wp_env], v * sd);
    } else if (t < M.n_elem + M.n_virtual) {
        int nv = t - M.n_elem;
        const float* h = M.hb + (size_t(r) * M.n_hb + nv) * 8;
        float v, d;
        clamped1d(M.uhb_coeff, M.uhb_left, M.uhb_right, M.uhb_nx, nv >= M.n_donor ? 1 : 0, (h[2] + M.uhb_shift) * M.uhb_scale, v, d);
        d *= M.uhb_scale;
        float u = 1.f - h[6];
        en = v * u * u;
        float* s = M.hb_sens + (size_t(r) * M.n_hb + nv) * 8;
        s[2] += d * u * u;
        s[6] += -2.f * v * u;
    }
    if (want_pot) {
        en = block_sum(en, sc);
        if (threadIdx.x == 0) atomicAdd(pot + r, en);
    }
}
struct MembranePotential : PotentialNode {
    CoordNode &res_pos, &env, &hb;
    int n_elem, n_restype, n_donor, n_acceptor;
    DevBuf<int> cb_index, env_index, restype;
    DevBuf<float> cov_mid, cov_sharp, cb_coeff, cb_left, cb_right, uhb_coeff, uhb_left, uhb_right;
    int cb_nx, uhb_nx;
    float cb_shift, cb_scale, uhb_shift, uhb_scale;
    MembranePotential(Engine&, const h5l::Node& g, CoordNode& res_pos_, CoordNode& env_, CoordNode& hb_)
        : res_pos(res_pos_), env(env_), hb(hb_) {
        check_elem_width_lower_bound(res_pos, 3);
        check_elem_width_lower_bound(env, 1);
        check_elem_width(hb, 7);
        n_elem = (int)h5_dims(g, "cb_index", 1)[0];
        auto cd = h5_dims(g, "cb_energy", 2), ud = h5_dims(g, "uhb_energy", 2);
        n_restype = (int)cd[0];
        cb_nx = (int)cd[1];
        uhb_nx = (int)ud[1];
        n_donor = (int)h5_dims(g, "donor_residue_ids", 1)[0];
        n_acceptor = (int)h5_dims(g, "acceptor_residue_ids", 1)[0];
        if (n_donor + n_acceptor != hb.n_elem) throw std::string("membrane_potential: donor/acceptor counts do not match protein_hbond");
        h5_check_size(g, "env_index", {(uint64_t)n_elem});
        h5_check_size(g, "residue_type", {(uint64_t)n_elem});
        h5_check_size(g, "cov_midpoint", {(uint64_t)n_restype});
        h5_check_size(g, "cov_sharpness", {(uint64_t)n_restype});
        h5_check_size(g, "uhb_energy", {2, (uint64_t)uhb_nx});
        auto ci = h5_read<int>(g, "cb_index"), ei = h5_read<int>(g, "env_index"), rt = h5_read<int>(g, "residue_type");
        for (int v : ci) if (v < 0 || v >= res_pos.n_elem) throw std::string("cb_index out of range");
        for (int v : ei) if (v < 0 || v >= env.n_elem) throw std::string("env_index out of range");
        for (int v : rt) if (v < 0 || v >= n_restype) throw std::string("residue_type out of range");
        cb_index.upload(ci); env_index.upload(ei); restype.upload(rt);
        cov_mid.upload(h5_read<float>(g, "cov_midpoint"));
        cov_sharp.upload(h5_read<float>(g, "cov_sharpness"));
        auto cbd = h5_read<double>(g, "cb_energy"), uhd = h5_read<double>(g, "uhb_energy");
        auto s1 = fit_clamped_spline_1d(n_restype, cb_nx, 1, cbd.data());
        auto s2 = fit_clamped_spline_1d(2, uhb_nx, 1, uhd.data());
        cb_coeff.upload(s1.coeff); cb_left.upload(s1.left); cb_right.upload(s1.right);
        uhb_coeff.upload(s2.coeff); uhb_left.upload(s2.left); uhb_right.upload(s2.right);
        cb_shift = -h5_attr<float>(g, "cb_energy", "z_min");
        cb_scale = (cb_nx - 1) / (h5_attr<float>(g, "cb_energy", "z_max") + cb_shift);
        uhb_shift = -h5_attr<float>(g, "uhb_energy", "z_min");
        uhb_scale = (uhb_nx - 1) / (h5_attr<float>(g, "uhb_energy", "z_max") + uhb_shift);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        MembraneDev M;
        M.cb = res_pos.output; M.env = env.output; M.hb = hb.output;
        M.cb_sens = res_pos.sens; M.env_sens = env.sens; M.hb_sens = hb.sens;
        M.n_cb = res_pos.n_elem; M.wp_cb = res_pos.wp; M.n_env = env.n_elem; M.wp_env = env.wp; M.n_hb = hb.n_elem;
        M.cb_index = cb_index.p; M.env_index = env_index.p; M.restype = restype.p;
        M.cov_mid = cov_mid.p; M.cov_sharp = cov_sharp.p;
        M.cb_coeff = cb_coeff.p; M.cb_left = cb_left.p; M.cb_right = cb_right.p;
        M.uhb_coeff = uhb_coeff.p; M.uhb_left = uhb_left.p; M.uhb_right = uhb_right.p;
        M.cb_nx = cb_nx; M.uhb_nx = uhb_nx; M.n_elem = n_elem; M.n_donor = n_donor; M.n_virtual = n_donor + n_acceptor;
        M.cb_shift = cb_shift; M.cb_scale = cb_scale; M.uhb_shift = uhb_shift; M.uhb_scale = uhb_scale;
        int n = n_elem + M.n_virtual;
        if (!n) return;
        k_membrane<<<dim3((n + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(M, potential, mode == PotentialAndDerivMode);
    }
};
RegisterNodeType<MembranePotential, 3> membrane_potential_node("membrane_potential");

}  // namespace
}  // namespace ub
