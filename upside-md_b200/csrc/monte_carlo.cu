// Monte-Carlo pivot and jump moves for every replica of a batch at once (reference src/monte_carlo_sampler.cpp).
//
// Reference: per system, per sampler - copy the positions, evaluate the potential, propose a move on the CPU, evaluate
// again, Metropolis test with the sampler's own counter-based random stream, restore the copy on rejection
// (monte_carlo_step :255-284).  Here one MC step of a sampler is five enqueued operations for ALL replicas: a device copy
// of the positions, one batched evaluation, one proposal kernel (a CTA per replica: thread 0 draws and builds the rigid
// transform, the CTA applies it to the moved atoms), a second batched evaluation, and the accept/restore kernel.  The
// random numbers are the reference's: RandomGenerator(seed_r, stream, 0, round), draws in the same order (rng.cuh).
#include <cmath>

#include "engine.h"
#include "rng.cuh"

namespace ub {

namespace {

constexpr int MC_TPB = 128;
constexpr uint32_t PIVOT_MOVE_RANDOM_STREAM = 2u, JUMP_MOVE_RANDOM_STREAM = 3u;   // random.h:12-17
constexpr float M_PI_F = 3.141592653589793f;

struct PivotLoc { int rama_atom[5]; int range0, range1, restype; };
struct JumpChain { int first_atom, next_first; float sigma_trans, sigma_rot; };

__device__ __forceinline__ f3 rot3(const float* U, f3 v) {
    return mk3(U[0] * v.x + U[1] * v.y + U[2] * v.z, U[3] * v.x + U[4] * v.y + U[5] * v.z, U[6] * v.x + U[7] * v.y + U[8] * v.z);
}
__device__ __forceinline__ f3 normalized3(f3 v) { return (1.f / sqrtf(mag2(v))) * v; }

// PivotSampler::propose_random_move (:80-155)
__global__ void __launch_bounds__(MC_TPB) k_mc_pivot(float* __restrict__ pos, int n_atom, const uint32_t* __restrict__ seed,
                                                     unsigned long long round, const PivotLoc* __restrict__ locs, int n_loc,
                                                     const float* __restrict__ proposal_pot, const float* __restrict__ cdf, int n_bin,
                                                     float* __restrict__ delta_lprob) {
    __shared__ float sU[18];
    __shared__ float sO[6];
    __shared__ int sR[2];
    const int r = blockIdx.x;
    float4* x = reinterpret_cast<float4*>(pos) + size_t(r) * n_atom;
    if (threadIdx.x == 0) {
        DeviceRandom random(seed[r], PIVOT_MOVE_RANDOM_STREAM, 0u, round);
        float u[4];
        random.uniform_open_closed(u);
        int loc = int(n_loc * u[2]);
        if (loc == n_loc) loc--;   // this may occur due to rounding
        const PivotLoc p = locs[loc];
        const int nb2 = n_bin * n_bin;
        const float* c = cdf + size_t(p.restype) * nb2;
        int lo = 0, hi = nb2;      // std::lower_bound: first bin whose cdf is not less than the variate
        while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] < u[3]) lo = mid + 1; else hi = mid; }
        const int pivot_bin = min(lo, nb2 - 1);
        const float new_lprob = proposal_pot[size_t(p.restype) * nb2 + pivot_bin];
        const int phi_bin = pivot_bin / n_bin, psi_bin = pivot_bin % n_bin;
        // a random location in that bin; the half-bin shift puts the centre of the left-most bin at -pi
        const float new_phi = (2.f * M_PI_F / n_bin) * (phi_bin + u[0] - 0.5f) - M_PI_F;
        const float new_psi = (2.f * M_PI_F / n_bin) * (psi_bin + u[1] - 0.5f) - M_PI_F;
        f3 d1, d2, d3, d4;
        const float4 a0 = x[p.rama_atom[0]], a1 = x[p.rama_atom[1]], a2 = x[p.rama_atom[2]], a3 = x[p.rama_atom[3]], a4 = x[p.rama_atom[4]];
        const f3 prevC = mk3(a0.x, a0.y, a0.z), N = mk3(a1.x, a1.y, a1.z), CA = mk3(a2.x, a2.y, a2.z), C = mk3(a3.x, a3.y, a3.z),
                 nextN = mk3(a4.x, a4.y, a4.z);
        const float old_phi = dihedral_germ(prevC, N, CA, C, d1, d2, d3, d4);
        const float old_psi = dihedral_germ(N, CA, C, nextN, d1, d2, d3, d4);
        int old_phi_bin = int((old_phi + M_PI_F) * (0.5f / M_PI_F) * n_bin + 0.5f);   // reverse the half-bin shift
        int old_psi_bin = int((old_psi + M_PI_F) * (0.5f / M_PI_F) * n_bin + 0.5f);
        old_phi_bin = old_phi_bin >= n_bin ? 0 : old_phi_bin;                          // periodicity
        old_psi_bin = old_psi_bin >= n_bin ? 0 : old_psi_bin;
        const float old_lprob = proposal_pot[(size_t(p.restype) * n_bin + old_phi_bin) * n_bin + old_psi_bin];
        axis_angle_to_rot(sU, new_phi - old_phi, normalized3(CA - N));
        axis_angle_to_rot(sU + 9, new_psi - old_psi, normalized3(C - CA));
        sO[0] = CA.x; sO[1] = CA.y; sO[2] = CA.z; sO[3] = C.x; sO[4] = C.y; sO[5] = C.z;
        sR[0] = p.range0; sR[1] = p.range1;
        delta_lprob[r] = new_lprob - old_lprob;
        // C and nextN move with the rotated part (pivot_range cannot contain them)
        const f3 phi_o = CA, psi_o = C;
        f3 y = phi_o + rot3(sU, (psi_o + rot3(sU + 9, C - psi_o)) - phi_o);
        x[p.rama_atom[3]] = make_float4(y.x, y.y, y.z, a3.w);
        y = phi_o + rot3(sU, (psi_o + rot3(sU + 9, nextN - psi_o)) - phi_o);
        x[p.rama_atom[4]] = make_float4(y.x, y.y, y.z, a4.w);
    }
    __syncthreads();
    const f3 phi_o = mk3(sO[0], sO[1], sO[2]), psi_o = mk3(sO[3], sO[4], sO[5]);
    for (int na = sR[0] + threadIdx.x; na < sR[1]; na += MC_TPB) {
        const float4 v = x[na];
        const f3 after_psi = psi_o + rot3(sU + 9, mk3(v.x, v.y, v.z) - psi_o);
        const f3 after_phi = phi_o + rot3(sU, after_psi - phi_o);
        x[na] = make_float4(after_phi.x, after_phi.y, after_phi.z, v.w);
    }
}

// JumpSampler::propose_random_move (:203-251): rigid translation or rotation about the centre of mass of one chain
__global__ void __launch_bounds__(MC_TPB) k_mc_jump(float* __restrict__ pos, int n_atom, const uint32_t* __restrict__ seed,
                                                    unsigned long long round, const JumpChain* __restrict__ chains, int n_chain,
                                                    float* __restrict__ delta_lprob) {
    __shared__ float sU[9];
    __shared__ float sV[3];
    __shared__ float red[3][32];
    __shared__ int sI[3];
    const int r = blockIdx.x;
    float4* x = reinterpret_cast<float4*>(pos) + size_t(r) * n_atom;
    if (threadIdx.x == 0) {
        DeviceRandom random(seed[r], JUMP_MOVE_RANDOM_STREAM, 0u, round);
        float u[4], n[4];
        random.uniform_open_closed(u);
        const int type = int(2 * u[0]);
        int chain = int(n_chain * u[3]);
        if (chain == n_chain) chain--;
        const JumpChain j = chains[chain];
        random.normal(n);
        if (type == 0) {
            const float s = j.sigma_trans / sqrtf(3.f);
            sV[0] = s * n[0]; sV[1] = s * n[1]; sV[2] = s * n[2];
        } else {
            const float angle = j.sigma_rot * n[0];
            f3 axis = mk3(n[1], n[2], n[3]);
            axis = (1.f / (sqrtf(mag2(axis)) + 1e-16f)) * axis;   // 1e-16 is the reference's paranoia against division by zero
            axis_angle_to_rot(sU, angle, axis);
        }
        sI[0] = type; sI[1] = j.first_atom; sI[2] = j.next_first;
        delta_lprob[r] = 0.f;
    }
    __syncthreads();
    const int type = sI[0], a0 = sI[1], a1 = sI[2];
    if (type == 0) {
        for (int na = a0 + threadIdx.x; na < a1; na += MC_TPB) {
            float4 v = x[na];
            v.x += sV[0]; v.y += sV[1]; v.z += sV[2];
            x[na] = v;
        }
        return;
    }
    f3 com = mk3(0.f, 0.f, 0.f);
    for (int na = a0 + threadIdx.x; na < a1; na += MC_TPB) { const float4 v = x[na]; com += mk3(v.x, v.y, v.z); }
    com.x = warp_sum(com.x); com.y = warp_sum(com.y); com.z = warp_sum(com.z);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = com.x; red[1][threadIdx.x >> 5] = com.y; red[2][threadIdx.x >> 5] = com.z; }
    __syncthreads();
    com = mk3(0.f, 0.f, 0.f);
    for (int w = 0; w < MC_TPB / 32; ++w) com += mk3(red[0][w], red[1][w], red[2][w]);
    com = (1.f / float(a1 - a0)) * com;
    for (int na = a0 + threadIdx.x; na < a1; na += MC_TPB) {
        const float4 v = x[na];
        const f3 y = com + rot3(sU, mk3(v.x, v.y, v.z) - com);
        x[na] = make_float4(y.x, y.y, y.z, v.w);
    }
}

// Metropolis test of monte_carlo_step (:271-283); the uniform variate is the sampler's NEXT draw (draw index n_draw)
__global__ void __launch_bounds__(MC_TPB) k_mc_accept(float* __restrict__ pos, const float* __restrict__ pos_copy, int n_atom,
                                                      const uint32_t* __restrict__ seed, unsigned long long round, uint32_t stream,
                                                      uint32_t n_draw, const float* __restrict__ temperature,
                                                      const float* __restrict__ old_pot, const float* __restrict__ new_pot,
                                                      const float* __restrict__ delta_lprob, unsigned long long* __restrict__ stats) {
    __shared__ int accept;
    const int r = blockIdx.x;
    if (threadIdx.x == 0) {
        const float lboltz_diff = delta_lprob[r] - (1.f / temperature[r]) * (new_pot[r] - old_pot[r]);
        DeviceRandom random(seed[r], stream, 0u, round);
        random.ctr[3] = n_draw;
        float u[4];
        random.uniform_open_closed(u);
        accept = lboltz_diff >= 0.f || expf(lboltz_diff) >= u[0];
        stats[2 * r + 1] += 1ull;
        if (accept) stats[2 * r] += 1ull;
    }
    __syncthreads();
    if (accept) return;
    float4* x = reinterpret_cast<float4*>(pos) + size_t(r) * n_atom;
    const float4* c = reinterpret_cast<const float4*>(pos_copy) + size_t(r) * n_atom;
    for (int na = threadIdx.x; na < n_atom; na += MC_TPB) x[na] = c[na];
}

}  // namespace

struct MonteCarlo {
    struct Sampler {
        std::string name;
        uint32_t stream;
        int n_bin = 0, n_loc = 0, n_chain = 0;
        DevBuf<PivotLoc> locs;
        DevBuf<float> pot, cdf;
        DevBuf<JumpChain> chains;
        DevBuf<unsigned long long> stats;   // [B][2] n_success, n_attempt
    };
    std::vector<std::unique_ptr<Sampler>> samplers;
    DevBuf<float> pos_copy, old_pot, delta_lprob;
};

void mc_destroy(MonteCarlo* m) { delete m; }

// MultipleMonteCarloSampler (:292-308) + the samplers' constructors (:28-78, :173-201)
void Engine::mc_init(const h5l::Node& input) {
    UB_CUDA(cudaSetDevice(device));
    auto m = std::unique_ptr<MonteCarlo>(new MonteCarlo);
    if (h5_has(input, "pivot_moves")) {
        const h5l::Node& g = h5_child(input, "pivot_moves");
        auto s = std::unique_ptr<MonteCarlo::Sampler>(new MonteCarlo::Sampler);
        s->name = "pivot";
        s->stream = PIVOT_MOVE_RANDOM_STREAM;
        auto dims = h5_dims(g, "proposal_pot", 3);
        const int n_layer = (int)dims[0];
        s->n_bin = (int)dims[1];
        s->n_loc = (int)h5_dims(g, "pivot_atom", 2)[0];
        h5_check_size(g, "proposal_pot", {(uint64_t)n_layer, (uint64_t)s->n_bin, (uint64_t)s->n_bin});
        h5_check_size(g, "pivot_atom", {(uint64_t)s->n_loc, 5});
        h5_check_size(g, "pivot_range", {(uint64_t)s->n_loc, 2});
        h5_check_size(g, "pivot_restype", {(uint64_t)s->n_loc});
        auto atom = h5_read<int>(g, "pivot_atom");
        auto range = h5_read<int>(g, "pivot_range");
        auto restype = h5_read<int>(g, "pivot_restype");
        std::vector<PivotLoc> locs(s->n_loc);
        for (int i = 0; i < s->n_loc; ++i) {
            PivotLoc& p = locs[i];
            for (int k = 0; k < 5; ++k) p.rama_atom[k] = atom[5 * i + k];
            p.range0 = range[2 * i]; p.range1 = range[2 * i + 1]; p.restype = restype[i];
            if (p.restype < 0 || p.restype >= n_layer) throw std::string("invalid pivot restype");
            for (int k = 0; k < 5; ++k) {
                if (p.rama_atom[k] < 0 || p.rama_atom[k] >= n_atom) throw std::string("pivot_atom out of range");
                if (p.range0 <= p.rama_atom[k] && p.rama_atom[k] < p.range1)
                    throw std::string("pivot_range cannot contain any atoms in pivot_atom ") + std::to_string(p.range0) + " <= " +
                        std::to_string(p.rama_atom[k]) + " < " + std::to_string(p.range1);
            }
            if (p.range0 < 0 || p.range1 > n_atom) throw std::string("pivot_range out of range");
        }
        std::vector<float> pot = h5_read<float>(g, "proposal_pot"), cdf(pot.size());
        const int nb2 = s->n_bin * s->n_bin;
        for (int nl = 0; nl < n_layer; ++nl) {   // normalise the negative log probability and its cdf, in double (:58-76)
            double sum_prob = 0.;
            for (int i = 0; i < nb2; ++i) {
                sum_prob += std::exp(-pot[size_t(nl) * nb2 + i]);
                cdf[size_t(nl) * nb2 + i] = (float)sum_prob;
            }
            const double inv = 1. / sum_prob, lsum = std::log(sum_prob);
            for (int i = 0; i < nb2; ++i) {
                cdf[size_t(nl) * nb2 + i] = (float)(double(cdf[size_t(nl) * nb2 + i]) * inv);   // float *= double, as the reference
                pot[size_t(nl) * nb2 + i] = (float)(double(pot[size_t(nl) * nb2 + i]) + lsum);
            }
            cdf[size_t(nl + 1) * nb2 - 1] = 1.f;   // ensure no rounding error here
        }
        s->locs.upload(locs);
        s->pot.upload(pot);
        s->cdf.upload(cdf);
        s->stats.upload(std::vector<unsigned long long>(size_t(n_rep) * 2, 0ull));
        if (s->n_loc) m->samplers.push_back(std::move(s));
    }
    if (h5_has(input, "jump_moves")) {
        const h5l::Node& g = h5_child(input, "jump_moves");
        auto s = std::unique_ptr<MonteCarlo::Sampler>(new MonteCarlo::Sampler);
        s->name = "jump";
        s->stream = JUMP_MOVE_RANDOM_STREAM;
        s->n_chain = (int)h5_dims(g, "atom_range", 2)[0];
        h5_check_size(g, "atom_range", {(uint64_t)s->n_chain, 2});
        h5_check_size(g, "sigma_trans", {(uint64_t)s->n_chain});
        h5_check_size(g, "sigma_rot", {(uint64_t)s->n_chain});
        auto range = h5_read<int>(g, "atom_range");
        auto st = h5_read<float>(g, "sigma_trans");
        auto sr = h5_read<float>(g, "sigma_rot");
        std::vector<JumpChain> chains(s->n_chain);
        for (int i = 0; i < s->n_chain; ++i) {
            chains[i] = JumpChain{range[2 * i], range[2 * i + 1], st[i], sr[i]};
            if (chains[i].first_atom < 0 || chains[i].next_first > n_atom || chains[i].first_atom >= chains[i].next_first)
                throw std::string("jump_moves atom_range out of range");
        }
        s->chains.upload(chains);
        s->stats.upload(std::vector<unsigned long long>(size_t(n_rep) * 2, 0ull));
        if (s->n_chain) m->samplers.push_back(std::move(s));
    }
    m->pos_copy.alloc(size_t(n_rep) * n_atom * 4);
    m->old_pot.alloc(n_rep);
    m->delta_lprob.alloc(n_rep);
    if (mc) mc_destroy(mc);
    mc = m.release();
}

void Engine::mc_execute(uint64_t round) {
    if (!mc) throw std::string("Monte-Carlo samplers are not initialised");
    if (seed.n != size_t(n_rep)) throw std::string("md_init must precede Monte-Carlo moves (seeds and temperatures)");
    UB_CUDA(cudaSetDevice(device));
    const size_t bytes = sizeof(float) * size_t(n_rep) * n_atom * 4;
    for (auto& sp : mc->samplers) {
        MonteCarlo::Sampler& s = *sp;
        UB_CUDA(cudaMemcpyAsync(mc->pos_copy.p, pos->output, bytes, cudaMemcpyDeviceToDevice, stream));
        compute(PotentialAndDerivMode);
        UB_CUDA(cudaMemcpyAsync(mc->old_pot.p, potential.p, sizeof(float) * n_rep, cudaMemcpyDeviceToDevice, stream));
        uint32_t n_draw;
        if (s.stream == PIVOT_MOVE_RANDOM_STREAM) {
            k_mc_pivot<<<n_rep, MC_TPB, 0, stream>>>(pos->output, n_atom, seed.p, round, s.locs.p, s.n_loc, s.pot.p, s.cdf.p, s.n_bin,
                                                     mc->delta_lprob.p);
            n_draw = 1u;
        } else {
            k_mc_jump<<<n_rep, MC_TPB, 0, stream>>>(pos->output, n_atom, seed.p, round, s.chains.p, s.n_chain, mc->delta_lprob.p);
            n_draw = 2u;
        }
        compute(PotentialAndDerivMode);
        k_mc_accept<<<n_rep, MC_TPB, 0, stream>>>(pos->output, mc->pos_copy.p, n_atom, seed.p, round, s.stream, n_draw, temperature.p,
                                                  mc->old_pot.p, potential.p, mc->delta_lprob.p, s.stats.p);
    }
}

int Engine::mc_n_samplers() const { return mc ? (int)mc->samplers.size() : 0; }
std::string Engine::mc_sampler_name(int i) const {
    if (!mc || i < 0 || i >= (int)mc->samplers.size()) throw std::string("no such Monte-Carlo sampler");
    return mc->samplers[i]->name;
}
void Engine::mc_stats(int i, uint64_t* n_success, uint64_t* n_attempt, bool reset) {
    if (!mc || i < 0 || i >= (int)mc->samplers.size()) throw std::string("no such Monte-Carlo sampler");
    UB_CUDA(cudaSetDevice(device));
    sync_and_check();
    auto v = mc->samplers[i]->stats.download();
    for (int r = 0; r < n_rep; ++r) { n_success[r] = v[2 * r]; n_attempt[r] = v[2 * r + 1]; }
    if (reset) UB_CUDA(cudaMemset(mc->samplers[i]->stats.p, 0, sizeof(unsigned long long) * 2 * n_rep));
}

}  // namespace ub
