// InteractionGraph host side + the three pair-term coordinate nodes that are plain sums over edges:
// protein_hbond (hbond.cpp:290-368), hbond_coverage (hbond.cpp:371-414), environment_coverage
// (environment.cpp:71-109).  All kernels are gather-form over the ELL neighbour tables of igraph.cuh.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "igraph.cuh"

namespace ub {

// ================================================================================================ IGraphHost
static int neighbor_capacity(float cutoff, int n_other) {
    double scale = 1.0;
    if (const char* s = getenv("UPSIDE_B200_NEIGHBOR_SCALE")) scale = std::max(0.05, atof(s));
    // 0.75 elements per cubic Angstrom of the cutoff sphere: side-chain beads come up to six per residue (one per rotamer
    // state), so clashing starting structures pack far more of them around a site than a relaxed chain does
    double k = scale * (16. + 0.75 * double(cutoff) * cutoff * cutoff);
    int K = (int)std::min<double>(n_other, std::ceil(k));
    return (std::max(K, 1) + 7) & ~7;   // multiple of eight: every ELL row starts on a 16-byte boundary
}

IGraphHost::IGraphHost(const h5l::Node& grp, bool symmetric_, int excl_, int n_dim1, int n_dim2, CoordNode* p1, CoordNode* p2)
    : symmetric(symmetric_), excl(excl_), node1(p1), node2(symmetric_ ? p1 : p2) {
    if (!(symmetric ^ bool(p2))) throw std::string("second node must be null iff symmetric interaction");
    auto sfx = [&](const char* b) { return std::string(b) + (symmetric ? "" : "1"); };
    n1 = (int)h5_dims(grp, sfx("index"), 1)[0];
    n2 = symmetric ? n1 : (int)h5_dims(grp, "index2", 1)[0];
    auto pd = h5_dims(grp, "interaction_param", 3);
    n_type1 = (int)pd[0]; n_type2 = (int)pd[1]; n_param = (int)pd[2];
    check_elem_width_lower_bound(*node1, n_dim1);
    if (!symmetric) check_elem_width_lower_bound(*node2, n_dim2);
    if (n1 >= 65536 || n2 >= 65536) throw std::string("interaction graph groups are limited to 65535 elements");
    h_param = h5_read<float>(grp, "interaction_param");
    h5_check_size(grp, sfx("type"), {(uint64_t)n1});
    h5_check_size(grp, sfx("id"), {(uint64_t)n1});
    loc1 = h5_read<int>(grp, sfx("index"));
    type1 = h5_read<int>(grp, sfx("type"));
    id1 = h5_read<int>(grp, sfx("id"));
    if (!symmetric) {
        h5_check_size(grp, "type2", {(uint64_t)n2});
        h5_check_size(grp, "id2", {(uint64_t)n2});
        loc2 = h5_read<int>(grp, "index2");
        type2 = h5_read<int>(grp, "type2");
        id2 = h5_read<int>(grp, "id2");
    } else { loc2 = loc1; type2 = type1; id2 = id1; }
    for (int v : loc1) if (v < 0 || v >= node1->n_elem) throw std::string("index out of range for first interaction group");
    for (int v : loc2) if (v < 0 || v >= node2->n_elem) throw std::string("index out of range for second interaction group");
    for (int v : type1) if (v < 0 || v >= n_type1) throw std::string("type out of range for first interaction group");
    for (int v : type2) if (v < 0 || v >= n_type2) throw std::string("type out of range for second interaction group");
    d_loc1.upload(loc1); d_type1.upload(type1); d_id1.upload(id1);
    d_loc2.upload(loc2); d_type2.upload(type2); d_id2.upload(id2);
    d_param.upload(h_param);
}

void IGraphHost::allocate(Engine* e) {
    engine = e;
    K1 = neighbor_capacity(cutoff, n2);
    K2 = symmetric ? K1 : neighbor_capacity(cutoff, n1);
    if (!lists) { use_cache = false; return; }
    if (symmetric) need1 = need2 = true;
    if (!need1 && !need2) throw std::string("interaction graph without a neighbour table");
    if (need1) {
        nbr1.alloc(size_t(e->n_rep) * n1 * K1);
        cnt1.alloc(size_t(e->n_rep) * n1);
    }
    if (!symmetric && need2) {
        nbr2.alloc(size_t(e->n_rep) * n2 * K2);
        cnt2.alloc(size_t(e->n_rep) * n2);
    }
    if (const char* s = getenv("UPSIDE_B200_NO_VERLET_CACHE")) use_cache = atoi(s) == 0;
    if (use_cache) {
        // The skin only trades rebuild frequency against candidates per refine; the exact list does not depend on it.  The
        // reference uses 1 + 0.2*cutoff (cache_buffer, interaction_graph.h:395-396); here an all-pairs rebuild costs more
        // relative to a refine than on the CPU, and 1.5x that skin measured 2% faster end to end (0.35x: 9% slower).
        skin = 1.5f * (1.0f + 0.2f * cutoff);
        if (const char* sk = getenv("UPSIDE_B200_SKIN_SCALE")) skin = std::max(0.1f, (float)atof(sk)) * (1.0f + 0.2f * cutoff);
        Kc1 = neighbor_capacity(cutoff + skin, n2);   // multiple of eight: slices (see k_pairlist)
        Kc2 = symmetric ? Kc1 : neighbor_capacity(cutoff + skin, n1);
        if (need1) {
            cand1.alloc(size_t(e->n_rep) * n1 * Kc1);
            ccnt1.alloc(size_t(e->n_rep) * n1);
        }
        cpos1.alloc(size_t(e->n_rep) * n1 * 4);
        if (!symmetric) {
            if (need2) {
                cand2.alloc(size_t(e->n_rep) * n2 * Kc2);
                ccnt2.alloc(size_t(e->n_rep) * n2);
            }
            cpos2.alloc(size_t(e->n_rep) * n2 * 4);
        }
        {   // k_refine stages the positions of both groups in shared memory: opt in above the 48 KB default, and give a
            // system that does not fit the uncached all-pairs path instead of a launch failure
            int device_smem = 0;
            UB_CUDA(cudaDeviceGetAttribute(&device_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
            const size_t smem = sizeof(float4) * size_t(symmetric ? n1 + 1 : n1 + n2 + 2);
            if (smem > (size_t)device_smem) {
                use_cache = false;
                cand1.alloc(0); cand2.alloc(0); ccnt1.alloc(0); ccnt2.alloc(0); cpos1.alloc(0); cpos2.alloc(0);
                return;
            }
            UB_CUDA(cudaFuncSetAttribute(k_refine<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        }
        flag.upload(std::vector<int>(e->n_rep, 2));   // 2 = never built
        fused_rebuild = double(n1) * n2 <= 2.5e5;
        if (const char* fr = getenv("UPSIDE_B200_FUSED_REBUILD")) fused_rebuild = atoi(fr) != 0;
        rep_list.alloc(e->n_rep);
        n_list.alloc(1);
    }
}

IGraphDev IGraphHost::dev() const {
    IGraphDev d;
    d.s1 = IGraphSide{node1->output, node1->sens, node1->n_elem, node1->wp, d_loc1.p, d_type1.p, d_id1.p, n1};
    d.s2 = IGraphSide{node2->output, node2->sens, node2->n_elem, node2->wp, d_loc2.p, d_type2.p, d_id2.p, n2};
    d.param = d_param.p;
    d.n_type1 = n_type1; d.n_type2 = n_type2; d.n_param = n_param;
    d.cutoff = cutoff; d.cutoff2 = cutoff * cutoff;
    d.symmetric = symmetric; d.excl = excl;
    d.nbr1 = nbr1.p; d.cnt1 = cnt1.p; d.K1 = K1;
    d.nbr2 = symmetric ? nbr1.p : nbr2.p; d.cnt2 = symmetric ? cnt1.p : cnt2.p; d.K2 = symmetric ? K1 : K2;
    d.error_flag = engine->error_flag.p;
    return d;
}

void IGraphHost::build(cudaStream_t s) {
    IGraphDev d = dev();
    constexpr int TILE = 128;
    constexpr int RGL = 8;
    if (!n1 || !n2) return;
    const int B = engine->n_rep;
    if (!use_cache) {
        if (need1)
            k_pairlist<TILE><<<dim3((n1 + TILE - 1) / TILE, B), TILE, 0, s>>>(d.s1, d.s2, d.nbr1, d.cnt1, d.K1, d.cutoff2, d.excl,
                                                                              symmetric, 1, d.error_flag, nullptr, nullptr, 0);
        if (!symmetric && need2)
            k_pairlist<TILE><<<dim3((n2 + TILE - 1) / TILE, B), TILE, 0, s>>>(d.s2, d.s1, d.nbr2, d.cnt2, d.K2, d.cutoff2, d.excl,
                                                                              0, 0, d.error_flag, nullptr, nullptr, 0);
        return;
    }
    const float cc = cutoff + skin;
    RefineTable T1{cand1.p, ccnt1.p, Kc1, d.nbr1, d.cnt1, d.K1};
    RefineTable T2{cand2.p, ccnt2.p, Kc2, symmetric ? nullptr : d.nbr2, symmetric ? nullptr : d.cnt2, d.K2};
    size_t smem = sizeof(float4) * size_t(symmetric ? n1 + 1 : n1 + n2 + 2);
    const int which = symmetric ? 1 : (need1 ? 1 : 0) | (need2 ? 2 : 0);
    const int rows = which == 3 ? ((n1 + 31) & ~31) + n2 : (which == 2 ? n2 : n1);
    const int tpb = std::min(256, std::max(64, (rows + 31) & ~31));   // smaller blocks keep more of them resident
    if (fused_rebuild) {
        // cache check, rebuild of the replicas that need it and refine in ONE launch (see VerletCache)
        VerletCache V{cpos1.p, cpos2.p, flag.p, (0.5f * skin) * (0.5f * skin), cc * cc, d.excl, symmetric ? 1 : 0};
        k_refine<RGL><<<B, tpb, smem, s>>>(d.s1, d.s2, symmetric ? 0 : 1, which, T1, T2, d.cutoff2, d.error_flag, V);
    } else {
        // large graphs: an all-pairs rebuild inside one CTA would hold up its whole launch (300 residues: 1.2 M pair tests per
        // replica); the flagged replicas are rebuilt by a tiled kernel over many CTAs instead
        const float max_move2 = (0.5f * skin) * (0.5f * skin);
        constexpr int TILE = 128;
        UB_CUDA(cudaMemsetAsync(n_list.p, 0, sizeof(int), s));
        k_cache_check<<<B, 128, 0, s>>>(d.s1, d.s2, symmetric ? 0 : 1, cpos1.p, cpos2.p, max_move2, flag.p, rep_list.p, n_list.p);
        const int gy = std::min(B, 592);
        if (need1)
            k_pairlist<TILE><<<dim3((n1 + TILE - 1) / TILE, gy), TILE, 0, s>>>(d.s1, d.s2, cand1.p, ccnt1.p, Kc1, cc * cc, d.excl, symmetric,
                                                                               1, d.error_flag, rep_list.p, n_list.p, 1);
        if (!symmetric && need2)
            k_pairlist<TILE><<<dim3((n2 + TILE - 1) / TILE, gy), TILE, 0, s>>>(d.s2, d.s1, cand2.p, ccnt2.p, Kc2, cc * cc, d.excl, 0, 0,
                                                                               d.error_flag, rep_list.p, n_list.p, 1);
        VerletCache V{nullptr, nullptr, nullptr, 0.f, 0.f, 0, 0};
        k_refine<RGL><<<B, tpb, smem, s>>>(d.s1, d.s2, symmetric ? 0 : 1, which, T1, T2, d.cutoff2, d.error_flag, V);
    }
    engine->mark(s, "(pairlist)");
}

bool IGraphHost::pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) {
    if (replica < 0 || replica >= engine->n_rep) throw std::string("replica out of range");
    engine->sync_and_check();
    const bool t1 = need1;   // read whichever table the node keeps; the transposed one lists the same pairs by second index
    const int n = t1 ? n1 : n2, K = t1 ? K1 : K2;
    std::vector<unsigned short> rows(size_t(n) * K);
    std::vector<int> cnt(n);
    UB_CUDA(cudaMemcpy(rows.data(), (t1 ? nbr1.p : nbr2.p) + size_t(replica) * n * K, rows.size() * sizeof(unsigned short), cudaMemcpyDeviceToHost));
    UB_CUDA(cudaMemcpy(cnt.data(), (t1 ? cnt1.p : cnt2.p) + size_t(replica) * n, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
    i1.clear();
    i2.clear();
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < cnt[i]; ++k) {
            int j = rows[size_t(i) * K + k];
            if (symmetric && !(i < j)) continue;   // the reference keeps i1<i2 only (interaction_graph.h:142-144)
            i1.push_back(t1 ? i : j);
            i2.push_back(t1 ? j : i);
        }
    sort_reference_order(i1, i2);
    return true;
}

std::vector<float> IGraphHost::count_edges_by_type(int replica) {
    std::vector<int> i1, i2;
    pairlist(replica, i1, i2);
    std::vector<float> ret(size_t(n_type1) * n_type2, 0.f);
    for (size_t e = 0; e < i1.size(); ++e) ret[type1[i1[e]] * n_type2 + type2[i2[e]]] += 1.f;
    return ret;
}

void IGraphHost::set_param(const std::vector<float>& p) {
    if (p.size() != h_param.size())
        throw "Bad param size, got " + std::to_string(p.size()) + " params, but expected " + std::to_string(h_param.size()) +
            " params of shape (" + std::to_string(n_type1) + ", " + std::to_string(n_type2) + ", " + std::to_string(n_param) + ")";
    h_param = p;
    d_param.upload(h_param);
}

namespace {

constexpr int TPB = 128;

// ================================================================================================ ProteinHBond
// forward: one thread per virtual site (a site has at most a few partners inside 3.5 A) sums its edge values in row order
__global__ void k_protein_hbond(IGraphDev g, const float* __restrict__ infer, float* __restrict__ out, int n_donor, int n_virtual) {
    const int r = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_virtual) return;
    const bool donor = e < n_donor;
    const int me = donor ? e : e - n_donor;
    const IGraphSide& mine = donor ? g.s1 : g.s2;
    const IGraphSide& other = donor ? g.s2 : g.s1;
    const int cnt = donor ? g.cnt1[size_t(r) * g.s1.n + me] : g.cnt2[size_t(r) * g.s2.n + me];
    const float* src = infer + (size_t(r) * n_virtual + e) * 8;
    const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
    float acc = 0.f;
    if (cnt) {
        float xs[8];
        load8(elem_ptr(mine, r, me), xs);
        const unsigned short* row = donor ? g.nbr1 + (size_t(r) * g.s1.n + me) * g.K1 : g.nbr2 + (size_t(r) * g.s2.n + me) * g.K2;
        for (int k = 0; k < cnt; ++k) {
            float xo[8], d1[6], d2[6];
            load8(elem_ptr(other, r, row[k]), xo);
            acc += donor ? protein_hbond_edge(g.param, xs, xo, d1, d2) : protein_hbond_edge(g.param, xo, xs, d1, d2);
        }
    }
    float* o = out + (size_t(r) * n_virtual + e) * 8;
    reinterpret_cast<float4*>(o)[0] = a;
    reinterpret_cast<float4*>(o)[1] = make_float4(b.x, b.y, 1.f - expf(-acc), 0.f);
}
// backward: one thread per virtual site.  Every site passes the six copied components of its sens row through; a donor
// also walks its row and evaluates each H-bond edge ONCE: its own half of the derivative stays in registers, the
// acceptor's half goes straight to the acceptor's sens row as two 16-byte reductions (REDG.ADD.F32x4).
__global__ void k_protein_hbond_deriv(IGraphDev g, const float* __restrict__ out, const float* __restrict__ sens, int n_donor,
                                      int n_virtual) {
    const int r = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_virtual) return;
    const bool donor = e < n_donor;
    const int me = donor ? e : e - n_donor;
    const float4* s_me = reinterpret_cast<const float4*>(sens + (size_t(r) * n_virtual + e) * 8);
    const float4 sa = s_me[0], sb = s_me[1];
    float acc[6] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y};   // pass-through of the six copied components
    if (donor) {
        const int cnt = g.cnt1[size_t(r) * g.s1.n + me];
        if (cnt) {
            float xs[8];
            load8(elem_ptr(g.s1, r, me), xs);
            const float ss_me = sb.z * (1.f - out[(size_t(r) * n_virtual + e) * 8 + 6]);
            const unsigned short* row = g.nbr1 + (size_t(r) * g.s1.n + me) * g.K1;
            for (int k = 0; k < cnt; ++k) {
                const int o = row[k], eo = n_donor + o;
                const float ss_o = sens[(size_t(r) * n_virtual + eo) * 8 + 6] * (1.f - out[(size_t(r) * n_virtual + eo) * 8 + 6]);
                const float es = ss_me + ss_o;
                float xo[8], d1[6], d2[6];
                load8(elem_ptr(g.s2, r, o), xo);
                protein_hbond_edge(g.param, xs, xo, d1, d2);
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[c] += es * d1[c];
                float4* d = reinterpret_cast<float4*>(elem_sens_ptr(g.s2, r, o));
                atomicAdd(d, make_float4(es * d2[0], es * d2[1], es * d2[2], es * d2[3]));
                atomicAdd(d + 1, make_float4(es * d2[4], es * d2[5], 0.f, 0.f));
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(elem_sens_ptr(donor ? g.s1 : g.s2, r, me));
    atomicAdd(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
    atomicAdd(dst + 1, make_float4(acc[4], acc[5], 0.f, 0.f));
}
// ================================================================================================ radial / hbond_sc_radial
// RadialHelper (sidechain_radial.cpp:16-79): clamped cubic B-spline of the distance, 16 knots behind p[0] = 1/dx per type
// pair; exclusion |id1-id2| <= 2; SidechainRadialPairs (:83-105, one group, symmetric) and HBondSidechainRadialPairs
// (:108-136, two groups).  Gather form: one thread per element walks its ELL row, keeps its own half of every pair's force
// in registers (a symmetric graph lists every pair in both rows; an asymmetric one is walked from both tables) and adds it
// to the element's sens row once; the energy counts every pair once.
constexpr int RADIAL_KNOT = 16;
__device__ __forceinline__ float radial_edge(const float* __restrict__ p, f3 x1, f3 x2, f3& d1) {
    const float inv_dx = p[0];
    const f3 disp = x1 - x2;
    const float dist2 = mag2(disp);
    const float inv_dist = rsqrtf(dist2 + 1e-7f);   // 1e-7 is divergence protection (sidechain_radial.cpp:54)
    const float dist_coord = dist2 * (inv_dist * inv_dx);
    float v, dv;
    clamped_deboor_vd(p + 1, RADIAL_KNOT, dist_coord, v, dv);
    d1 = (inv_dist * inv_dx * dv) * disp;
    return v;
}
// second = 0: elements of group 1 over nbr1; second = 1: elements of group 2 over nbr2 (asymmetric graphs only)
__global__ void k_radial(IGraphDev g, int second, int want_pot, float* __restrict__ potential) {
    const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const IGraphSide& mine = second ? g.s2 : g.s1;
    const IGraphSide& other = second ? g.s1 : g.s2;
    float e = 0.f;
    if (i < mine.n) {
        const int cnt = second ? g.cnt2[size_t(r) * mine.n + i] : g.cnt1[size_t(r) * mine.n + i];
        const unsigned short* row = second ? g.nbr2 + (size_t(r) * mine.n + i) * g.K2 : g.nbr1 + (size_t(r) * mine.n + i) * g.K1;
        const f3 xi = ld3(elem_ptr(mine, r, i));
        const int ti = mine.type[i];
        f3 acc = mk3(0.f, 0.f, 0.f);
        for (int k = 0; k < cnt; ++k) {
            const int j = row[k];
            const f3 xj = ld3(elem_ptr(other, r, j));
            const int tj = other.type[j];
            const int t1 = second ? tj : ti, t2 = second ? ti : tj;
            f3 d1;
            // operands in (group 1, group 2) order - for a symmetric graph (lower index, higher index), the reference's edge
            const bool first = second ? false : (g.symmetric ? i < j : true);
            const float v = first ? radial_edge(g.param + size_t(t1 * g.n_type2 + t2) * g.n_param, xi, xj, d1)
                                  : radial_edge(g.param + size_t(t1 * g.n_type2 + t2) * g.n_param, xj, xi, d1);
            acc += first ? d1 : -d1;
            if (first) e += v;   // every pair counted once
        }
        if (cnt) atomic_add3(elem_sens_ptr(mine, r, i), acc);
    }
    if (want_pot && !second) {
        e = warp_sum(e);
        if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(potential + r, e);
    }
}
template <bool SYM> struct RadialPairs : PotentialNode {
    IGraphHost ig;
    RadialPairs(Engine&, const h5l::Node& g, CoordNode& a) : ig(g, true, EXCL_SEQ2, 3, 3, &a, nullptr) { init(); }
    RadialPairs(Engine&, const h5l::Node& g, CoordNode& a, CoordNode& b) : ig(g, false, EXCL_SEQ2, 3, 3, &a, &b) { init(); }
    void init() {
        if (ig.n_param != 1 + RADIAL_KNOT) throw "radial interaction expects " + std::to_string(1 + RADIAL_KNOT) + " parameters per type pair";
        update_cutoff();
    }
    void update_cutoff() {   // update_cutoffs (interaction_graph.h:383-398) with RadialHelper::cutoff (sidechain_radial.cpp:33-36)
        float c = 0.f;
        for (int t1 = 0; t1 < ig.n_type1; ++t1)
            for (int t2 = 0; t2 < ig.n_type2; ++t2) {
                const float* p = &ig.h_param[size_t(t1 * ig.n_type2 + t2) * ig.n_param];
                c = std::max(c, float((RADIAL_KNOT - 2 - 1e-6) / p[0]));
                if (SYM)
                    for (int k = 0; k < ig.n_param; ++k)
                        if (p[k] != ig.h_param[size_t(t2 * ig.n_type2 + t1) * ig.n_param + k]) throw std::string("incompatible parameters");
            }
        ig.cutoff = c;
    }
    void finalize() override { ig.allocate(engine); }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!ig.n1 || !ig.n2) return;
        ig.build(s);
        const int want = mode == PotentialAndDerivMode;
        k_radial<<<dim3((ig.n1 + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(ig.dev(), 0, want, potential);
        if (!SYM) k_radial<<<dim3((ig.n2 + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(ig.dev(), 1, want, potential);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override {   // the cutoff follows the parameters; the tables keep their capacity
        const float old = ig.cutoff;
        ig.set_param(p);
        update_cutoff();
        if (ig.cutoff > old) { ig.cutoff = old; throw std::string("radial: new parameters enlarge the cutoff beyond the allocated pair-list capacity"); }
    }
};
typedef RadialPairs<true> SidechainRadialPairs;
typedef RadialPairs<false> HBondSidechainRadialPairs;

struct ProteinHBond : CoordNode {
    CoordNode& infer;
    IGraphHost ig;
    int n_donor, n_acceptor, n_virtual;
    ProteinHBond(Engine&, const h5l::Node& g, CoordNode& infer_)
        : CoordNode((int)(h5_dims(g, "index1", 1)[0] + h5_dims(g, "index2", 1)[0]), 7), infer(infer_),
          ig(g, false, EXCL_NONE, 6, 6, &infer_, &infer_) {
        if (ig.n_param != 8) throw std::string("protein_hbond expects 8 interaction parameters");
        n_donor = ig.n1; n_acceptor = ig.n2; n_virtual = n_donor + n_acceptor;
        if (n_virtual != infer.n_elem) throw std::string("protein_hbond expects one element per virtual site");
        ig.cutoff = sqrtf(3.5f * 3.5f);   // hbond.cpp:124,158-160
    }
    void finalize() override { ig.allocate(engine); }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_virtual) return;
        ig.build(s);
        k_protein_hbond<<<dim3((n_virtual + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(ig.dev(), infer.output, output, n_donor, n_virtual);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_virtual) return;
        k_protein_hbond_deriv<<<dim3((n_virtual + TPB - 1) / TPB, engine->n_rep), TPB, 0, s>>>(ig.dev(), output, sens, n_donor, n_virtual);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // hbond.cpp:306-311: H-bond fraction per site
        if (level < 1) return;
        out.push_back({"hbond", {(uint64_t)n_virtual}, false, [this](int r) {
            auto o = host_rows(output, r);
            std::vector<float> v(n_virtual);
            for (int i = 0; i < n_virtual; ++i) v[i] = o[size_t(i) * wp + 6];
            return v;
        }});
    }
};
RegisterNodeType<ProteinHBond, 1> protein_hbond_node("protein_hbond");
RegisterNodeType<SidechainRadialPairs, 1> radial_node("radial");                         // sidechain_radial.cpp:209
RegisterNodeType<HBondSidechainRadialPairs, 2> hbond_sc_radial_node("hbond_sc_radial");   // sidechain_radial.cpp:210

}  // namespace
}  // namespace ub
