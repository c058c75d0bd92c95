// RotamerSidechain on the B200: one CTA per replica runs the whole node - 1-body energies, bead-pair energies,
// residue-pair graph construction, damped loopy belief propagation, Bethe free energy and the backward pass -
// with the BP state (beliefs, residue adjacency bitmap) resident in shared memory.
//
// Reference: src/rotamer.cpp: fill_holders :793-852, solve_for_marginals :1005-1061, EdgeHolder::update_beliefs
// :453-522, NodeHolder::standardize_belief_update :258-273, calculate_marginals :275-281,403-429, free energies
// :292-302,431-451, propagate_derivatives :956-985.  Bead ids encode (residue k << 8 | n_rot << 4 | rot)
// (upside_config.py:976-983).
//
// Differences of formulation (not of result): the reference keys residue pairs through an open-addressed table
// (EdgeLocator :134-206) in bead-pair emission order; here a residue-adjacency bitmap in shared memory gives every
// residue pair a slot by prefix popcount, so the pair -> slot map needs no hashing and is deterministic.  Messages are
// L1-normalised with an exact reciprocal instead of the 12-bit rcpps (:513); both only rescale messages.
#include <algorithm>
#include <cmath>

#include "igraph.cuh"

namespace ub {
namespace {

constexpr int MAXR = 6;           // most rotamer states per residue (UPPER_ROT-1 in the reference)
constexpr int RTPB = 256;         // threads per replica CTA
constexpr int RG = 8;             // lanes per bead row
constexpr int MAX_PROB_NODES = 4;

struct RotamerDev {
    IGraphDev g;
    QuadSplineShape q;
    int n_bead, n_res, n_words;
    const int *bead_res, *bead_rot, *res_nrot;
    int n_prob;
    const float* prob_out[MAX_PROB_NODES];
    float* prob_sens[MAX_PROB_NODES];
    int prob_wp[MAX_PROB_NODES], prob_n[MAX_PROB_NODES];
    float damping, tol;
    int max_iter, chunk, max_pairs;
    // per-replica scratch in global memory
    float* pmat;              // [B][max_pairs][36]  pair energy -> probability -> marginal
    float* msg;               // [B][2][max_pairs][12]
    unsigned short* pair_ab;  // [B][max_pairs][2]
    int* inc;                 // [B][2*max_pairs]
    float* node_marg;         // [B][n_res][6]
    int* stats;               // [B][4]: n_iter, n_pair, converged, -
    float* potential;
    int* error_flag;
};

// number of set bits of `row` (nW words) strictly between positions lo and hi
__device__ __forceinline__ int rank_between(const unsigned* row, int nW, int lo, int hi) {
    int cnt = 0;
    const int w0 = (lo + 1) >> 5, wh = hi >> 5;
    const int w1 = min(wh, nW - 1);
    for (int w = w0; w <= w1; ++w) {
        unsigned bits = row[w];
        if (w == w0) bits &= ~0u << ((lo + 1) & 31);
        if (w == wh) bits &= (hi & 31) ? (~0u >> (32 - (hi & 31))) : 0u;
        cnt += __popc(bits);
    }
    return cnt;
}

// value only (forward); ordering (lo,hi) as the reference's i1<i2 edge
__device__ __forceinline__ float bead_pair_value(const RotamerDev& P, int r, int lo, int hi) {
    float x1[8], x2[8], d1[6], d2[6];
    load8(elem_ptr(P.g.s1, r, lo), x1);
    load8(elem_ptr(P.g.s1, r, hi), x2);
    const float* prm = P.g.param + (size_t(P.g.s1.type[lo]) * P.g.n_type2 + P.g.s1.type[hi]) * P.g.n_param;
    return quadspline_edge(prm, P.q, x1, x2, d1, d2);
}

__device__ __forceinline__ float block_max_bcast(float v, float* red) {
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float m = red[0];
    for (int k = 1; k < RTPB / 32; ++k) m = fmaxf(m, red[k]);
    return m;
}

// one message sweep over all residue-pair edges: old beliefs/messages -> new messages (rotamer.cpp:468-520)
__device__ __forceinline__ void bp_messages(const RotamerDev& P, int n_pair, const unsigned short* pair_ab, const float* pmat,
                                            const float* bel_old, const float* msg_old, float* msg_new) {
    for (int e = threadIdx.x; e < n_pair; e += RTPB) {
        int A = pair_ab[2 * e], B = pair_ab[2 * e + 1];
        int nA = P.res_nrot[A], nB = P.res_nrot[B];
        float v1[MAXR], v2[MAXR];
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            v1[a] = a < nA ? bel_old[A * MAXR + a] / (1e-10f + msg_old[e * 12 + a]) : 0.f;
            v2[a] = a < nB ? bel_old[B * MAXR + a] / (1e-10f + msg_old[e * 12 + 6 + a]) : 0.f;
        }
        const float* M = pmat + size_t(e) * 36;
        float m1[MAXR], m2[MAXR];
#pragma unroll
        for (int a = 0; a < MAXR; ++a) { m1[a] = 0.f; m2[a] = 0.f; }
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            if (a < nA) {
#pragma unroll
                for (int b = 0; b < MAXR; ++b) {
                    if (b < nB) {
                        float p = M[a * 6 + b];
                        m1[a] = fmaf(p, v2[b], m1[a]);   // apply_left : message to A
                        m2[b] = fmaf(v1[a], p, m2[b]);   // apply_right: message to B
                    }
                }
            }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int a = 0; a < MAXR; ++a) { s1 += m1[a]; s2 += m2[a]; }
        float i1 = 1.f / s1, i2 = 1.f / s2;
#pragma unroll
        for (int a = 0; a < MAXR; ++a) { msg_new[e * 12 + a] = m1[a] * i1; msg_new[e * 12 + 6 + a] = m2[a] * i2; }
    }
}

// node update: belief = prob * prod(incoming messages), max-normalised and damped (rotamer.cpp:488-499,258-273);
// returns this thread's largest signed deviation cur-old
__device__ __forceinline__ float bp_nodes(const RotamerDev& P, const int* istart, const int* inc, const float* prob,
                                          const float* msg_new, const float* bel_old, float* bel_new, float damping) {
    float dev = 0.f;
    for (int A = threadIdx.x; A < P.n_res; A += RTPB) {
        int nA = P.res_nrot[A];
        if (nA < 2) continue;
        float b[MAXR];
#pragma unroll
        for (int a = 0; a < MAXR; ++a) b[a] = prob[A * MAXR + a];
        for (int t = istart[A]; t < istart[A + 1]; ++t) {
            int code = inc[t];
            const float* m = msg_new + (code >> 1) * 12 + (code & 1) * 6;
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < MAXR; ++a) { b[a] *= m[a]; s += b[a]; }
            float is = 1.f / s;
#pragma unroll
            for (int a = 0; a < MAXR; ++a) b[a] *= is;
        }
        float mx = b[0];
#pragma unroll
        for (int a = 1; a < MAXR; ++a) mx = fmaxf(mx, b[a]);
        float imx = 1.f / mx;
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            float o = bel_old[A * MAXR + a];
            float n = (damping != 0.f) ? (1.f - damping) * imx * b[a] + damping * o : imx * b[a];
            if (a < nA) dev = fmaxf(dev, n - o);
            bel_new[A * MAXR + a] = n;
        }
    }
    return dev;
}

__global__ void __launch_bounds__(RTPB) k_rotamer(RotamerDev P, int want_pot) {
    extern __shared__ float smem[];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int nR = P.n_res, nW = P.n_words;
    float* Enode = smem;                       // [nR][6] energy, then unused
    float* prob = Enode + nR * MAXR;           // [nR][6]
    float* bel0 = prob + nR * MAXR;            // [nR][6]
    float* bel1 = bel0 + nR * MAXR;            // [nR][6]
    float* offs = bel1 + nR * MAXR;            // [nR]
    unsigned* bitmap = reinterpret_cast<unsigned*>(offs + nR);   // [nR][nW] symmetric residue adjacency
    int* estart = reinterpret_cast<int*>(bitmap + nR * nW);      // [nR+1] first slot of pairs (A,B>A)
    int* istart = estart + nR + 1;                                // [nR+1] incidence CSR
    float* red = reinterpret_cast<float*>(istart + nR + 1);      // [32]

    float* pmat = P.pmat + size_t(r) * P.max_pairs * 36;
    float* msg0 = P.msg + size_t(r) * 2 * P.max_pairs * 12;
    float* msg1 = msg0 + size_t(P.max_pairs) * 12;
    unsigned short* pair_ab = P.pair_ab + size_t(r) * P.max_pairs * 2;
    int* inc = P.inc + size_t(r) * 2 * P.max_pairs;
    float* node_marg = P.node_marg + size_t(r) * nR * MAXR;
    const unsigned short* nbr = P.g.nbr1 + size_t(r) * P.n_bead * P.g.K1;
    const int* cnt = P.g.cnt1 + size_t(r) * P.n_bead;

    // ---- 1. one-body energies and residue adjacency -------------------------------------------------------
    for (int i = tid; i < nR * MAXR; i += RTPB) Enode[i] = 0.f;
    for (int i = tid; i < nR * nW; i += RTPB) bitmap[i] = 0u;
    __syncthreads();
    for (int i = tid; i < P.n_bead; i += RTPB) {
        float e = 0.f;
        int loc = P.g.s1.loc[i];
        for (int p = 0; p < P.n_prob; ++p) e += P.prob_out[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]];
        atomicAdd(&Enode[P.bead_res[i] * MAXR + P.bead_rot[i]], e);
        int A = P.bead_res[i];
        if (P.res_nrot[A] > 1) {
            const unsigned short* row = nbr + size_t(i) * P.g.K1;
            for (int k = 0; k < cnt[i]; ++k) {
                int Bq = P.bead_res[row[k]];
                if (P.res_nrot[Bq] > 1) atomicOr(&bitmap[A * nW + (Bq >> 5)], 1u << (Bq & 31));
            }
        }
    }
    __syncthreads();
    for (int A = tid; A < nR; A += RTPB) {   // energy offset = smallest 1-body energy (convert_energy_to_prob :239-256)
        float m = Enode[A * MAXR];
        for (int a = 1; a < P.res_nrot[A]; ++a) m = fminf(m, Enode[A * MAXR + a]);
        offs[A] = m;
    }
    if (tid == 0) {
        int eu = 0, ei = 0;
        for (int A = 0; A < nR; ++A) {
            estart[A] = eu;
            istart[A] = ei;
            const unsigned* row = bitmap + A * nW;
            int deg = 0;
            for (int w = 0; w < nW; ++w) deg += __popc(row[w]);
            eu += rank_between(row, nW, A, nR);
            ei += deg;
        }
        estart[nR] = eu;
        istart[nR] = ei;
    }
    __syncthreads();
    const int n_pair = estart[nR];
    if (n_pair > P.max_pairs) {   // uniform across the block
        if (tid == 0) atomicExch(P.error_flag, 2);
        return;
    }
    // ---- 2. pair slots, incidence lists, zeroed pair energies -----------------------------------------------
    for (int A = tid; A < nR; A += RTPB) {
        const unsigned* row = bitmap + A * nW;
        int t = istart[A], up = 0;
        for (int w = 0; w < nW; ++w) {
            unsigned bits = row[w];
            while (bits) {
                int C = (w << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                if (C > A) {
                    int e = estart[A] + up++;
                    pair_ab[2 * e] = (unsigned short)A;
                    pair_ab[2 * e + 1] = (unsigned short)C;
                    inc[t++] = 2 * e;
                } else {
                    int e = estart[C] + rank_between(bitmap + C * nW, nW, C, A);
                    inc[t++] = 2 * e + 1;
                }
            }
        }
    }
    for (int i = tid; i < n_pair * 36; i += RTPB) pmat[i] = 0.f;
    __syncthreads();
    // ---- 3. bead-pair energies: residue-pair matrices, single-state partners folded into node energies -------
    float e11 = 0.f;   // energy of (1-state, 1-state) pairs, needed only for the potential
    {
        const int grp = tid / RG, lane = tid % RG;
        for (int i0 = 0; i0 < P.n_bead; i0 += RTPB / RG) {
            int i = i0 + grp;
            float fold = 0.f;
            int A = 0, ra = 0, nA = 0;
            if (i < P.n_bead) {
                A = P.bead_res[i]; ra = P.bead_rot[i]; nA = P.res_nrot[A];
                const unsigned short* row = nbr + size_t(i) * P.g.K1;
                int c = cnt[i];
                for (int k = lane; k < c; k += RG) {
                    int j = row[k];
                    int Bq = P.bead_res[j], rb = P.bead_rot[j], nB = P.res_nrot[Bq];
                    if (nA > 1 && nB > 1) {
                        if (j > i) {
                            float V = bead_pair_value(P, r, i, j);
                            int e, idx;
                            if (A < Bq) { e = estart[A] + rank_between(bitmap + A * nW, nW, A, Bq); idx = ra * 6 + rb; }
                            else { e = estart[Bq] + rank_between(bitmap + Bq * nW, nW, Bq, A); idx = rb * 6 + ra; }
                            atomicAdd(&pmat[size_t(e) * 36 + idx], V);
                        }
                    } else if (nA > 1) {
                        fold += bead_pair_value(P, r, min(i, j), max(i, j));
                    } else if (nB == 1 && j > i) {
                        if (want_pot) e11 += bead_pair_value(P, r, i, j);
                    }
                }
            }
            fold = group_sum<RG>(fold);
            if (i < P.n_bead && lane == 0 && nA > 1) atomicAdd(&Enode[A * MAXR + ra], fold);
        }
    }
    __syncthreads();
    // ---- 4. energies -> probabilities --------------------------------------------------------------------------
    for (int i = tid; i < nR * MAXR; i += RTPB) {
        int A = i / MAXR, a = i % MAXR;
        float p = a < P.res_nrot[A] ? __expf(offs[A] - Enode[i]) : 0.f;
        prob[i] = p;
        bel0[i] = p;
    }
    for (int i = tid; i < n_pair * 36; i += RTPB) pmat[i] = __expf(-pmat[i]);
    for (int e = tid; e < n_pair; e += RTPB) {
        int nA = P.res_nrot[pair_ab[2 * e]], nB = P.res_nrot[pair_ab[2 * e + 1]];
        for (int a = 0; a < 6; ++a) { msg0[e * 12 + a] = a < nA ? 1.f : 0.f; msg0[e * 12 + 6 + a] = a < nB ? 1.f : 0.f; }
    }
    __syncthreads();
    // ---- 5. belief propagation ----------------------------------------------------------------------------------
    // initial sweep: first messages from (prob, unit messages); node beliefs restart from prob/max (rotamer.cpp:1034)
    bp_messages(P, n_pair, pair_ab, pmat, bel0, msg0, msg1);
    for (int A = tid; A < nR; A += RTPB) {
        float mx = prob[A * MAXR];
        for (int a = 1; a < MAXR; ++a) mx = fmaxf(mx, prob[A * MAXR + a]);
        float imx = 1.f / mx;
        for (int a = 0; a < MAXR; ++a) bel1[A * MAXR + a] = prob[A * MAXR + a] * imx;
    }
    __syncthreads();
    float* bel_cur = bel1; float* bel_old = bel0;
    float* msg_cur = msg1; float* msg_old = msg0;
    float max_dev = 1e10f;
    int iter = 0;
    for (; max_dev > P.tol && iter < P.max_iter; iter += P.chunk) {
        float dev = 0.f;
        for (int j = 0; j < P.chunk; ++j) {
            float* t = bel_cur; bel_cur = bel_old; bel_old = t;
            t = msg_cur; msg_cur = msg_old; msg_old = t;
            bp_messages(P, n_pair, pair_ab, pmat, bel_old, msg_old, msg_cur);
            __syncthreads();
            dev = bp_nodes(P, istart, inc, prob, msg_cur, bel_old, bel_cur, P.damping);
            __syncthreads();
        }
        max_dev = block_max_bcast(dev, red);
    }
    if (tid == 0) {
        int* st = P.stats + size_t(r) * 4;
        st[0] = iter; st[1] = n_pair; st[2] = max_dev <= P.tol;
    }
    // ---- 6. marginals (and Bethe free energy) ---------------------------------------------------------------------
    float en = e11;
    for (int A = tid; A < nR; A += RTPB) {
        int nA = P.res_nrot[A];
        float b[MAXR], s = 0.f;
        for (int a = 0; a < MAXR; ++a) { b[a] = nA > 1 ? bel_cur[A * MAXR + a] : (a == 0 ? 1.f : 0.f); s += b[a]; }
        float is = 1.f / s;
        for (int a = 0; a < MAXR; ++a) { b[a] *= is; bel_cur[A * MAXR + a] = b[a]; node_marg[A * MAXR + a] = b[a]; }
        if (want_pot) {
            float e = offs[A];
            for (int a = 0; a < nA; ++a) e += b[a] * __logf((1e-10f + b[a]) / (1e-10f + prob[A * MAXR + a]));
            en += e;
        }
    }
    __syncthreads();
    for (int e = tid; e < n_pair; e += RTPB) {
        int A = pair_ab[2 * e], Bq = pair_ab[2 * e + 1];
        int nA = P.res_nrot[A], nB = P.res_nrot[Bq];
        float bc1[MAXR], bc2[MAXR];
        for (int a = 0; a < MAXR; ++a) {
            bc1[a] = a < nA ? bel_cur[A * MAXR + a] / (1e-10f + msg_cur[e * 12 + a]) : 0.f;
            bc2[a] = a < nB ? bel_cur[Bq * MAXR + a] / (1e-10f + msg_cur[e * 12 + 6 + a]) : 0.f;
        }
        float* M = pmat + size_t(e) * 36;
        float s = 0.f;
        for (int a = 0; a < nA; ++a) for (int b = 0; b < nB; ++b) s += M[a * 6 + b] * bc1[a] * bc2[b];
        float is = 1.f / s;
        for (int a = 0; a < MAXR; ++a)
            for (int b = 0; b < MAXR; ++b) {
                float pr = M[a * 6 + b];
                float mg = (a < nA && b < nB) ? pr * bc1[a] * bc2[b] * is : 0.f;
                if (want_pot && a < nA && b < nB)
                    en += mg * __logf((1e-10f + mg) / (1e-10f + pr * bel_cur[A * MAXR + a] * bel_cur[Bq * MAXR + b]));
                M[a * 6 + b] = mg;
            }
    }
    if (want_pot) {
        float tot = block_sum(en, red);
        if (tid == 0) P.potential[r] = tot;
    }
    __syncthreads();
    // ---- 7. backward: d/d(bead) = sum over partners of marginal * dV/d(bead); 1-body sens += node marginal ---------
    {
        const int grp = tid / RG, lane = tid % RG;
        for (int i0 = 0; i0 < P.n_bead; i0 += RTPB / RG) {
            int i = i0 + grp;
            float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (i < P.n_bead) {
                int A = P.bead_res[i], ra = P.bead_rot[i], nA = P.res_nrot[A];
                float xi[8];
                load8(elem_ptr(P.g.s1, r, i), xi);
                int ti = P.g.s1.type[i];
                const unsigned short* row = nbr + size_t(i) * P.g.K1;
                int c = cnt[i];
                for (int k = lane; k < c; k += RG) {
                    int j = row[k];
                    int Bq = P.bead_res[j], rb = P.bead_rot[j], nB = P.res_nrot[Bq];
                    float s;
                    if (nA > 1 && nB > 1) {
                        if (A < Bq) s = pmat[size_t(estart[A] + rank_between(bitmap + A * nW, nW, A, Bq)) * 36 + ra * 6 + rb];
                        else s = pmat[size_t(estart[Bq] + rank_between(bitmap + Bq * nW, nW, Bq, A)) * 36 + rb * 6 + ra];
                    } else if (nA > 1) s = bel_cur[A * MAXR + ra];
                    else if (nB > 1) s = bel_cur[Bq * MAXR + rb];
                    else s = 1.f;
                    float xj[8], d1[6], d2[6];
                    load8(elem_ptr(P.g.s1, r, j), xj);
                    int tj = P.g.s1.type[j];
                    if (i < j) {
                        quadspline_edge(P.g.param + (size_t(ti) * P.g.n_type2 + tj) * P.g.n_param, P.q, xi, xj, d1, d2);
#pragma unroll
                        for (int q = 0; q < 6; ++q) acc[q] += s * d1[q];
                    } else {
                        quadspline_edge(P.g.param + (size_t(tj) * P.g.n_type2 + ti) * P.g.n_param, P.q, xj, xi, d1, d2);
#pragma unroll
                        for (int q = 0; q < 6; ++q) acc[q] += s * d2[q];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[q] = group_sum<RG>(acc[q]);
            if (i < P.n_bead && lane == 0) {
                float* dst = elem_sens_ptr(P.g.s1, r, i);
#pragma unroll
                for (int q = 0; q < 6; ++q) dst[q] += acc[q];
                float m = bel_cur[P.bead_res[i] * MAXR + P.bead_rot[i]];
                int loc = P.g.s1.loc[i];
                for (int p = 0; p < P.n_prob; ++p) P.prob_sens[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]] += m;
            }
        }
    }
}

struct RotamerSidechain : PotentialNode {
    std::vector<CoordNode*> prob_nodes;
    IGraphHost ig;
    int nka = 15, nk = 16;
    float knot_spacing = 0.5f;
    int n_res = 0, n_words = 0, max_pairs = 0;
    std::vector<int> bead_res, bead_rot, res_nrot, res_key;
    DevBuf<int> d_bead_res, d_bead_rot, d_res_nrot, inc, stats;
    DevBuf<float> pmat, msg, node_marg;
    DevBuf<unsigned short> pair_ab;
    float damping, tol;
    int max_iter, chunk;
    size_t smem_bytes = 0;

    RotamerSidechain(Engine&, const h5l::Node& g, const ArgList& args)
        : prob_nodes(args.begin() + 1, args.end()), ig(h5_child(g, "pair_interaction"), true, EXCL_ROTAMER, 6, 6, args[0], nullptr) {
        if (args[0]->wp != 8) throw std::string("rotamer expects 8-float rows for bead positions");
        if ((int)prob_nodes.size() > MAX_PROB_NODES) throw std::string("too many 1-body probability nodes for rotamer");
        for (size_t i = 0; i < prob_nodes.size(); ++i)
            if (ig.node1->n_elem != prob_nodes[i]->n_elem)
                throw "rotamer positions have " + std::to_string(ig.node1->n_elem) + " elements but the " + std::to_string(i) +
                    "-th (0-indexed) probability node has only " + std::to_string(prob_nodes[i]->n_elem) + " elements.";
        // knot counts follow the table shape (compile-time macros in the reference, bead_interaction.h:12-27)
        if (ig.n_param == 2 * 15 + 2 * 16) { nka = 15; nk = 16; knot_spacing = 0.5f; }
        else if (ig.n_param == 2 * 8 + 2 * 12) { nka = 8; nk = 12; knot_spacing = 1.f; }
        else if (ig.n_param == 2 * 8 + 2 * 9) { nka = 8; nk = 9; knot_spacing = 1.f; }
        else throw "unsupported rotamer pair_interaction parameter count " + std::to_string(ig.n_param);
        ig.cutoff = float((nk - 2 - 1e-6) / double(1.f / knot_spacing));   // bead_interaction.h:191-193
        check_compatible();
        damping = h5_attr<float>(g, ".", "damping");
        max_iter = h5_attr<int>(g, ".", "max_iter");
        tol = h5_attr<float>(g, ".", "tol");
        chunk = std::max(1, h5_attr<int>(g, ".", "iteration_chunk_size"));
        // residues = distinct (n_rot, k) in order of first appearance
        std::map<int, int> key_to_res;
        for (int b = 0; b < ig.n1; ++b) {
            unsigned id = (unsigned)ig.id1[b];
            int rot = id & 15, n_rot = (id >> 4) & 15, key = id >> 4;
            if (rot >= n_rot) throw std::string("invalid rotamer number");
            if (n_rot > MAXR) throw "invalid rotamer count " + std::to_string(n_rot);
            auto it = key_to_res.find(key);
            if (it == key_to_res.end()) {
                it = key_to_res.emplace(key, (int)res_nrot.size()).first;
                res_nrot.push_back(n_rot);
                res_key.push_back(key);
            }
            bead_res.push_back(it->second);
            bead_rot.push_back(rot);
        }
        n_res = (int)res_nrot.size();
        n_words = (n_res + 31) / 32;
        if (n_res >= 65536) throw std::string("too many residues for rotamer node");
        int n_multi = 0;
        for (int n : res_nrot) n_multi += n > 1;
        long full = long(n_multi) * (n_multi - 1) / 2;
        double scale = 1.0;
        if (const char* s = getenv("UPSIDE_B200_NEIGHBOR_SCALE")) scale = std::max(0.05, atof(s));
        max_pairs = (int)std::max<long>(1, std::min<long>(full, (long)std::ceil(16 * scale * n_res)));
        d_bead_res.upload(bead_res);
        d_bead_rot.upload(bead_rot);
        d_res_nrot.upload(res_nrot);
    }
    void check_compatible() {
        // symmetric tables must satisfy p(t1,t2).ang1 == p(t2,t1).ang2 and equal radial parts (bead_interaction.h:209-218)
        int n = ig.n_type1, np = ig.n_param;
        if (ig.n_type1 != ig.n_type2) throw std::string("incompatible parameters");
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                const float* p1 = &ig.h_param[(size_t(a) * n + b) * np];
                const float* p2 = &ig.h_param[(size_t(b) * n + a) * np];
                for (int k = 0; k < nka; ++k)
                    if (p1[k] != p2[k + nka] || p1[k + nka] != p2[k]) throw std::string("bad angular match");
                for (int k = 0; k < 2 * nk; ++k)
                    if (p1[2 * nka + k] != p2[2 * nka + k]) throw std::string("incompatible parameters");
            }
    }
    void finalize() override {
        ig.allocate(engine);
        size_t B = engine->n_rep;
        pmat.alloc(B * max_pairs * 36);
        msg.alloc(B * 2 * max_pairs * 12);
        pair_ab.alloc(B * max_pairs * 2);
        inc.alloc(B * 2 * max_pairs);
        node_marg.alloc(B * n_res * MAXR);
        stats.alloc(B * 4);
        smem_bytes = sizeof(float) * (size_t(n_res) * MAXR * 4 + n_res + 32) + sizeof(unsigned) * size_t(n_res) * n_words +
                     sizeof(int) * 2 * (n_res + 1);
        if (smem_bytes > 200 * 1024) throw std::string("rotamer node: system too large for the shared-memory BP kernel");
        UB_CUDA(cudaFuncSetAttribute(k_rotamer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    }
    RotamerDev dev() {
        RotamerDev P;
        P.g = ig.dev();
        P.q.nka = nka; P.q.nk = nk; P.q.inv_dx = 1.f / knot_spacing; P.q.inv_dtheta = (nka - 3) / 2.f;
        P.n_bead = ig.n1; P.n_res = n_res; P.n_words = n_words;
        P.bead_res = d_bead_res.p; P.bead_rot = d_bead_rot.p; P.res_nrot = d_res_nrot.p;
        P.n_prob = (int)prob_nodes.size();
        for (int i = 0; i < P.n_prob; ++i) {
            P.prob_out[i] = prob_nodes[i]->output; P.prob_sens[i] = prob_nodes[i]->sens;
            P.prob_wp[i] = prob_nodes[i]->wp; P.prob_n[i] = prob_nodes[i]->n_elem;
        }
        P.damping = damping; P.tol = tol; P.max_iter = max_iter; P.chunk = chunk; P.max_pairs = max_pairs;
        P.pmat = pmat.p; P.msg = msg.p; P.pair_ab = pair_ab.p; P.inc = inc.p; P.node_marg = node_marg.p; P.stats = stats.p;
        P.potential = potential; P.error_flag = engine->error_flag.p;
        return P;
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!ig.n1) return;
        ig.build(s);
        k_rotamer<<<engine->n_rep, RTPB, smem_bytes, s>>>(dev(), mode == PotentialAndDerivMode);
    }
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override { return ig.pairlist(replica, i1, i2); }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); check_compatible(); }

    std::vector<float> get_value_by_name(int replica, const char* log_name) override {
        std::string nm(log_name);
        engine->sync_and_check();
        if (nm == "count_edges_by_type") return ig.count_edges_by_type(replica);
        if (nm == "n_node") return {float(n_res)};
        if (nm == "bead_marginal") {   // B200 extension: node marginal of each bead's (residue, rotamer)
            std::vector<float> nmg(size_t(n_res) * MAXR);
            UB_CUDA(cudaMemcpy(nmg.data(), node_marg.p + size_t(replica) * n_res * MAXR, nmg.size() * sizeof(float), cudaMemcpyDeviceToHost));
            std::vector<float> out(ig.n1);
            for (int b = 0; b < ig.n1; ++b) out[b] = nmg[bead_res[b] * MAXR + bead_rot[b]];
            return out;
        }
        if (nm == "solve_stats") {     // B200 extension: (n_iter, n residue pairs, converged)
            std::vector<int> st(4);
            UB_CUDA(cudaMemcpy(st.data(), stats.p + size_t(replica) * 4, 4 * sizeof(int), cudaMemcpyDeviceToHost));
            return {float(st[0]), float(st[1]), float(st[2])};
        }
        throw std::string("Value ") + log_name + " not implemented";
    }
};
RegisterNodeType<RotamerSidechain, -1> rotamer_node("rotamer");

}  // namespace
}  // namespace ub
