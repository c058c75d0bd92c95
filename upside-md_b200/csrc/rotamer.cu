// RotamerSidechain on the B200 (reference src/rotamer.cpp).  Four kernels per evaluation, all batched over replicas:
//
//   k_rot_prep    one CTA per replica: residue adjacency bitmap (shared memory) from the bead ELL rows -> a slot for
//                 every residue pair by prefix popcount, incidence lists, and the bead pairs as CSR rows (partner + a
//                 "code" saying where the pair energy goes / which marginal weights its derivative), rows sorted by
//                 length for the thread-per-row consumers, 1-body node energies.
//   k_rot_energy  throughput-shaped, persistent: the B-spline table staged once per CTA and shared by the two replicas a CTA
//                 serves at a time, beads staged per replica; ONE THREAD PER CSR ROW walks the partners above the bead,
//                 pair energies -> residue-pair matrices; partners with a single rotamer state are folded into the bead's
//                 node energy (fill_holders, rotamer.cpp:793-852).
//   k_rot_bp2     one CTA per replica, state resident in shared memory (pair matrices, messages, beliefs): damped loopy
//                 belief propagation (solve_for_marginals :1005-1061, update_beliefs :453-522, standardize_belief_update
//                 :258-273), marginals (:275-281,403-429), Bethe free energy (:292-302,431-451), and the backward weight of
//                 every CSR entry; k_rot_bp is the general path for replicas that do not fit.
//   k_rot_deriv   same shape as k_rot_energy: per bead, sum over ALL partners of weight * dV/d(bead) with the pair term
//                 recomputed (propagate_derivatives :956-985) - gather form, fixed order, no atomics; node marginals into
//                 the 1-body sens.
//
// Bead ids encode (residue k << 8 | n_rot << 4 | rot) (upside_config.py:976-983).  Differences of formulation, not of
// result: residue pairs get slots from an adjacency bitmap instead of the reference's open-addressed EdgeLocator
// (:134-206); messages are L1-normalised with an exact reciprocal instead of the 12-bit rcpps (:513).
#include <algorithm>
#include <cmath>

#include "igraph.cuh"
#include "edgelist.cuh"

namespace ub {
namespace {

constexpr int MAXR = 6;            // most rotamer states per residue (UPPER_ROT-1 in the reference)
constexpr int PREP_TPB = 256;
#ifndef UB_EDGE_MAXT
#define UB_EDGE_MAXT 448   // two replicas per CTA (2 x 224 threads at 100 residues) ...
#endif
#ifndef UB_EDGE_OCC
#define UB_EDGE_OCC 2     // ... and two CTAs per SM: 28 warps per SM instead of 21 with one table copy per replica (measured -11 %)
#endif
constexpr int EDGE_TPB = 256;            // most threads on the rows of ONE replica
constexpr int EDGE_MAXT = UB_EDGE_MAXT;  // most threads of an edge-kernel CTA (several replicas share one staged table)
constexpr int EDGE_OCC = UB_EDGE_OCC;
constexpr int BP_TPB = 384;
constexpr int MAX_PROB_NODES = 4;
constexpr int CODE_SS = INT_MIN;   // (single, single)
// Entry codes.  code >= 0: index into the replica's pair-matrix array (slot*36 + a*6 + b).  Otherwise, unless CODE_SS,
// t = -2-code names a node state (t>>1 = res*6+rot) whose marginal weights the entry in the backward pass: bit 0 set
// ("fold") = (multi-state bead, single-state partner), the node is the bead's own and the pair energy folds into its
// 1-body energy; bit 0 clear = (single-state bead, multi-state partner), the node is the partner's.
__host__ __device__ constexpr int code_node(int node, bool fold) { return -2 - (2 * node + (fold ? 1 : 0)); }
__device__ __forceinline__ bool code_is_fold(int cd) { return cd < 0 && cd != CODE_SS && ((-2 - cd) & 1); }

struct BeadRec {   // 32 bytes, staged in shared memory by the edge kernels
    float x, y, z, dx, dy, dz;
    int type;
    int res_rot;   // res << 3 | rot
};

struct RotamerDev {
    IGraphDev g;
    QuadSplineShape q;
    int n_bead, n_res, n_words, n_type;
    const int *bead_res, *bead_rot, *res_nrot;
    int loc_identity;       // bead b is element b of the position and probability nodes (ff_1): no index indirection
    const int* res_first;   // first bead of a residue (fast build: the beads of a residue are contiguous, one per state)
    const float* table;   // symmetric-compressed B-spline table: rows (t1<=t2), n_param floats each
    int n_prob;
    const float* prob_out[MAX_PROB_NODES];
    float* prob_sens[MAX_PROB_NODES];
    int prob_wp[MAX_PROB_NODES], prob_n[MAX_PROB_NODES];
    float damping, tol;
    int max_iter, chunk, max_pairs, smem_pairs, multi_bead_states, cap_e;
    // per-replica scratch in global memory.  The bead pair list of a replica is kept as CSR rows (one row per bead, both
    // directions of every pair, partners ascending): entry e of row i = rowstart[i] + k.
    int* rowstart;            // [B][n_bead+1]
    unsigned short* dj;       // [B][cap_e]  partner bead
    int* code;                // [B][cap_e]  where the pair energy goes / which marginal weights its derivative
    float* ss;                // [B][cap_e]  that marginal (written by the BP kernels)
    unsigned short* lower;    // [B][n_bead] first row entry the energy kernel evaluates
    unsigned short* order_e;  // [B][n_bead] rows by falling number of entries the energy kernel evaluates
    unsigned short* order_d;  // [B][n_bead] rows by falling length
    float* enode;             // [B][n_res][6]  1-body energy per (residue, state)
    float* fold;              // [B][n_bead]    energy of single-state partners
    float* e11;               // [B]
    float* pmat;              // [B][max_pairs][36]  pair energy (-> marginal on the general BP path)
    unsigned short* pair_ab;  // [B][max_pairs][2]
    int* inc;                 // [B][2*max_pairs]
    int* istart;              // [B][n_res+1]
    float* node_marg;         // [B][n_res][6]
    int* stats;               // [B][4]: n_iter, n_pair, converged, -
    int* n_bad;               // [B] solves that ran to (within a chunk of) max_iter (n_bad_solve, rotamer.cpp:604,784-785)
    int* slow_list;           // [B] replicas the fast BP kernel declined
    int* n_slow;              // [1] (reset by k_rot_prep)
    // per-residue free energy (residue_free_energies, rotamer.cpp:868-902): node term + half of every incident pair term;
    // only filled while *fe_flag != 0 (logging runs and the rotamer_free_energy accessor)
    float* res_fe;            // [B][n_res]
    const int* fe_flag;       // [1]
    int n_rep;
    float* potential;
    int* error_flag;
};

// rows by falling key (clipped to 255) into order[0..n): counting sort, see sort_rows_desc (igraph.cuh)
template <typename KeyF>
__device__ __forceinline__ void sort_rows_desc_u16(int n, KeyF key, unsigned short* order, int* hist) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[min(max(key(i), 0), 255)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {
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
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&hist[min(max(key(i), 0), 255)], 1)] = (unsigned short)i;
    __syncthreads();
}

// ================================================================================================ prep
__global__ void __launch_bounds__(PREP_TPB) k_rot_prep(RotamerDev P) {
    extern __shared__ unsigned smem_u[];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int nR = P.n_res, nW = P.n_words, K = P.g.K1, nb = P.n_bead;
    unsigned* bitmap = smem_u;                                   // [nR][nW] symmetric residue adjacency (multi-state only)
    int* estart = reinterpret_cast<int*>(bitmap + nR * nW);      // [nR+1] first slot of pairs (A,B>A)
    int* istart = estart + nR + 1;                               // [nR+1]
    int* deg = istart + nR + 1;                                  // [nR]
    int* ebase = deg + nR;                                       // [nR] slot base: estart - (neighbours below)
    float* en = reinterpret_cast<float*>(ebase + nR);            // [nR*6]
    int* rr = reinterpret_cast<int*>(en + nR * MAXR);            // [n_bead] res << 4 | rot << 1 | multi-state
    int* rs = rr + nb;                                           // [n_bead+1] CSR row starts
    int* wtot = rs + nb + 1;                                     // [33]
    int* hist = wtot + 33;                                       // [256]
    unsigned short* wpre = reinterpret_cast<unsigned short*>(hist + 256);   // [nR][nW] set bits in the words before w
    unsigned short* lo_s = wpre + nR * nW;                       // [n_bead]
    unsigned short* ce_s = lo_s + nb;                            // [n_bead] entries the energy kernel evaluates
    const unsigned short* nbr = P.g.nbr1 + size_t(r) * nb * K;
    const int* cnt = P.g.cnt1 + size_t(r) * nb;

    for (int i = tid; i < nR * nW; i += PREP_TPB) bitmap[i] = 0u;
    for (int i = tid; i < nR * MAXR; i += PREP_TPB) en[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < nb; i += PREP_TPB) {
        float e = 0.f;
        int loc = P.g.s1.loc[i];
        for (int p = 0; p < P.n_prob; ++p) e += P.prob_out[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]];
        int A = P.bead_res[i], ra = P.bead_rot[i];
        // one contributor per state unless several beads share a state (a shared-memory float atomicAdd is a CAS loop)
        if (P.multi_bead_states) atomicAdd(&en[A * MAXR + ra], e); else en[A * MAXR + ra] = e;
        rr[i] = (A << 4) | (ra << 1) | (P.res_nrot[A] > 1 ? 1 : 0);
    }
    __syncthreads();
    // Pass 1, one thread per bead row (16-byte loads: eight partners each): residue adjacency bits; the number of partners
    // below the bead, how many of those fold into it, and how many entries the energy kernel will evaluate.
    for (int i = tid; i < nb; i += PREP_TPB) {
        const uint4* row4 = reinterpret_cast<const uint4*>(nbr + size_t(i) * K);   // rows are 16-byte aligned (K % 8 == 0)
        const int c = cnt[i], me = rr[i];
        const bool mA = me & 1;
        int lo = 0, nfl = 0, nev = 0;
        int last_res = -1;   // partners ascend and the beads of a residue are neighbours: one adjacency update per run
        for (int k0 = 0; k0 < c; k0 += 8) {
            const uint4 v = row4[k0 >> 3];
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (k0 + u >= c) break;
                const int j = (u & 1) ? int(w[u >> 1] >> 16) : int(w[u >> 1] & 0xffffu);
                const int other = rr[j];
                const bool mB = other & 1, below = j < i;
                if (mA && mB && (other >> 4) != last_res) {
                    // both directions, so that the adjacency stays symmetric even if a row was truncated by a capacity
                    // overflow (reported through error_flag): every index derived below relies on that symmetry
                    last_res = other >> 4;
                    atomicOr(&bitmap[(me >> 4) * nW + (other >> 9)], 1u << ((other >> 4) & 31));
                    atomicOr(&bitmap[(other >> 4) * nW + (me >> 9)], 1u << ((me >> 4) & 31));
                }
                lo += below;
                nfl += below && mA && !mB;
                nev += below ? (mA && !mB) : (mA || !mB);   // entries above a single-state bead with a multi-state partner are skipped
            }
        }
        lo_s[i] = (unsigned short)(lo - nfl);
        ce_s[i] = (unsigned short)nev;
    }
    scan_row_lengths(nb, [&](int i) { return cnt[i]; }, rs, wtot);   // (ends with a barrier: the bitmap is complete too)
    // Per residue: running popcounts of its adjacency row (wpre), degree, and the number of neighbours below itself.  With
    // them the slot of pair (A,B>A) is ebase[A] + R_A(B), R_A(B) = wpre[A][B>>5] + popc(row_A[B>>5] below bit B): two
    // independent shared-memory loads instead of a loop over the row.
    for (int A = tid; A < nR; A += PREP_TPB) {
        const unsigned* row = bitmap + A * nW;
        int run = 0, below = 0;
        for (int w = 0; w < nW; ++w) {
            unsigned bits = row[w];
            wpre[A * nW + w] = (unsigned short)run;
            if (w == (A >> 5)) below = run + __popc(bits & ((1u << (A & 31)) - 1u));
            run += __popc(bits);
        }
        deg[A] = run;
        ebase[A] = -below;            // completed after the scan
        estart[A] = run - below;      // upper degree, scanned in place below
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scans of upper degree (pair slots) and full degree (incidence lists), one warp
        int per = (nR + 31) / 32, a0 = min(nR, tid * per), a1 = min(nR, a0 + per);
        int su = 0, sf = 0;
        for (int A = a0; A < a1; ++A) { su += estart[A]; sf += deg[A]; }
        int pu = su, pf = sf;
        for (int o = 1; o < 32; o <<= 1) {
            int tu = __shfl_up_sync(UB_FULL_MASK, pu, o), tf = __shfl_up_sync(UB_FULL_MASK, pf, o);
            if (tid >= o) { pu += tu; pf += tf; }
        }
        int eu = pu - su, ef = pf - sf;
        for (int A = a0; A < a1; ++A) {
            int up = estart[A];
            estart[A] = eu; istart[A] = ef; ebase[A] += eu;
            eu += up; ef += deg[A];
        }
        if (tid == 31) { estart[nR] = pu; istart[nR] = pf; }
    }
    __syncthreads();
    auto slot_of = [&](int A, int B) {   // A < B, adjacent
        return ebase[A] + (int)wpre[A * nW + (B >> 5)] + __popc(bitmap[A * nW + (B >> 5)] & ((1u << (B & 31)) - 1u));
    };
    const int n_pair = estart[nR];
    if (tid == 0) { P.stats[size_t(r) * 4 + 1] = n_pair; P.e11[r] = 0.f; }
    if (tid == 0 && r == 0) *P.n_slow = 0;
    if (*P.fe_flag) for (int A = tid; A < nR; A += PREP_TPB) P.res_fe[size_t(r) * nR + A] = 0.f;
    int* rowstart = P.rowstart + size_t(r) * (nb + 1);
    if (n_pair > P.max_pairs || rs[nb] > P.cap_e) {   // uniform across the block: report, and leave an empty graph behind
        if (tid == 0) { atomicExch(P.error_flag, n_pair > P.max_pairs ? 2 : 3); P.stats[size_t(r) * 4 + 1] = 0; }
        for (int A = tid; A <= nR; A += PREP_TPB) P.istart[size_t(r) * (nR + 1) + A] = 0;
        for (int i = tid; i <= nb; i += PREP_TPB) rowstart[i] = 0;
        for (int i = tid; i < nb; i += PREP_TPB) {
            P.lower[size_t(r) * nb + i] = 0; P.order_e[size_t(r) * nb + i] = P.order_d[size_t(r) * nb + i] = (unsigned short)i;
        }
        return;
    }
    unsigned short* pair_ab = P.pair_ab + size_t(r) * P.max_pairs * 2;
    int* inc = P.inc + size_t(r) * 2 * P.max_pairs;
    for (int A = tid; A <= nR; A += PREP_TPB) P.istart[size_t(r) * (nR + 1) + A] = istart[A];
    for (int i = tid; i <= nb; i += PREP_TPB) rowstart[i] = rs[i];
    for (int A = tid; A < nR; A += PREP_TPB) {
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
                    inc[t++] = 2 * slot_of(C, A) + 1;
                }
            }
        }
    }
    for (int i = tid; i < nR * MAXR; i += PREP_TPB) P.enode[size_t(r) * nR * MAXR + i] = en[i];
    {   // pair energies start from zero: 16-byte stores
        float4* pm4 = reinterpret_cast<float4*>(P.pmat + size_t(r) * P.max_pairs * 36);
        for (int i = tid; i < n_pair * 9; i += PREP_TPB) pm4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // CSR rows: partner + code per entry.  Inside a row the partners below the bead come first, those of them whose energy
    // folds into this bead last among them, then the partners above in ascending order: the energy kernel evaluates the
    // contiguous tail that starts at `lower` = (partners below) - (folding partners below).
    unsigned short* dj = P.dj + size_t(r) * P.cap_e;
    int* code = P.code + size_t(r) * P.cap_e;
    for (int i = tid; i < nb; i += PREP_TPB) {   // pass 2, one thread per row: place the entries
        const uint4* row4 = reinterpret_cast<const uint4*>(nbr + size_t(i) * K);
        const int c = cnt[i], base = rs[i], me = rr[i], A = me >> 4, ra = (me >> 1) & 7;
        const bool mA = me & 1;
        int run_f = lo_s[i], run_o = 0;   // next slot of a folding / other partner below the bead
        int last_res = -1, last_slot36 = 0;   // the pair slot is looked up once per run of beads of the same partner residue
        for (int k0 = 0; k0 < c; k0 += 8) {
            const uint4 v = row4[k0 >> 3];
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k >= c) break;
                const int j = (u & 1) ? int(w[u >> 1] >> 16) : int(w[u >> 1] & 0xffffu);
                const int other = rr[j], Bq = other >> 4, rb = (other >> 1) & 7;
                const bool mB = other & 1, below = j < i;
                int cd;
                if (mA && mB) {
                    if (Bq != last_res) { last_res = Bq; last_slot36 = 36 * (A < Bq ? slot_of(A, Bq) : slot_of(Bq, A)); }
                    cd = last_slot36 + (A < Bq ? ra * 6 + rb : rb * 6 + ra);
                }
                else if (mA) cd = code_node(A * MAXR + ra, true);
                else if (mB) cd = code_node(Bq * MAXR + rb, false);
                else cd = CODE_SS;
                const int at = !below ? k : ((mA && !mB) ? run_f++ : run_o++);
                dj[base + at] = (unsigned short)j;
                code[base + at] = cd;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nb; i += PREP_TPB) P.lower[size_t(r) * nb + i] = lo_s[i];
    // thread-per-row consumers take the rows in order of falling length, so that the threads of a warp finish together
    sort_rows_desc_u16(nb, [&](int i) { return (int)ce_s[i]; }, P.order_e + size_t(r) * nb, hist);
    sort_rows_desc_u16(nb, [&](int i) { return rs[i + 1] - rs[i]; }, P.order_d + size_t(r) * nb, hist);
}

// ================================================================================================ build (fast path)
// k_rot_build replaces the Verlet cache, k_refine and k_rot_prep for the rotamer graph when every (residue, state) owns
// exactly one bead and the beads of a residue are contiguous (ff_1).  The beads of a residue are its rotamer states - up
// to six points a few Angstrom apart - so the bead pair list is a RESIDUE pair list with a 36-bit mask per pair:
//   1. one bounding sphere per residue; all residue pairs (A<B) are tested sphere against sphere (n_res^2/2 cheap tests,
//      no cached candidate list to keep valid), survivors compacted in (A,B) order;
//   2. one thread per surviving residue pair applies the reference's exact bead predicate (interaction_graph.h:223-244,
//      __fadd_rn/__fmul_rn, no FMA) to its nA x nB bead pairs -> mask; non-empty masks set the residue adjacency bits;
//   3. the same slot arithmetic as k_rot_prep (prefix popcounts over adjacency rows) names every active pair, and one
//      thread per bead walks its residue's adjacency row and reads its partners off the masks: the CSR rows, codes, pair
//      slots and incidence lists come out exactly as k_rot_prep writes them, so the consumers are unchanged.
// Nothing is read from or written to global memory between the bead positions and the CSR rows.
constexpr int BUILD_TPB_SMALL = 256, BUILD_TPB_LARGE = 1024;
// capacities of the shared-memory arrays (sphere-test survivors, active residue pairs); what does not fit spills to a
// per-replica global scratch area (clashing start structures have several times the pairs of a relaxed chain) up to
// capc_tot / capa_tot, beyond which the error flag is raised
struct BuildLay { int capc, capa, capc_tot, capa_tot; unsigned long long* spill; int* bstats; };
template <typename T, bool SPILL> struct SpillArr {
    T* s; T* g; int cap;
    // SPILL: one generic-address access behind a pointer select; otherwise a plain shared-memory access
    __device__ __forceinline__ T& operator[](int k) const {
        if (!SPILL) return s[k];
        T* p = k < cap ? s + k : g + (k - cap);
        return *p;
    }
};

__device__ __forceinline__ int block_excl_scan(int v, int* wtot, int& total) {   // every thread calls; contains barriers
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += t; }
    __syncthreads();   // previous readers of wtot are done
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < nw; ++k) { const int t = wtot[k]; base += k < w ? t : 0; tot += t; }
    total = tot;
    return base + incl - v;
}

// one neighbour of a residue in 64 bits: the pair's mask as seen from that residue (bits 0-35: bit a*6+b, a = own state,
// b = the neighbour's), whether the neighbour lies above (bit 36), the neighbour (bits 37-47) and the pair's matrix slot
// (bits 48-63; both multi-state, else unused)
__device__ __forceinline__ unsigned long long nbr_pack(unsigned long long mask, bool up, int C, int slot) {
    return mask | ((unsigned long long)(up ? 1 : 0) << 36) | ((unsigned long long)C << 37) | ((unsigned long long)slot << 48);
}
constexpr int BUILD_MAX_RES = 2047, BUILD_MAX_SLOT = 65535;

__host__ __device__ inline size_t build_spill_words(const BuildLay& L) {   // 64-bit words of global scratch per replica
    const size_t n_row_g = 2 * size_t(L.capa_tot - L.capa), n_c_g = size_t(L.capc_tot - L.capc);
    return n_row_g + n_c_g + (n_c_g + 1) / 2 + (6 * n_row_g + 3) / 4 + (n_row_g + 3) / 4;
}

// returns false (uniformly, nothing written) if SPILL is off and the shared-memory capacities do not hold this replica
template <bool SPILL, int BUILD_TPB>
__device__ __forceinline__ bool rot_build_body(const RotamerDev& P, const BuildLay& L, unsigned long long* smem_ull) {
    const int r = blockIdx.x, tid = threadIdx.x;
    const int nR = P.n_res, nW = P.n_words, nb = P.n_bead;
    unsigned long long* rown_s = smem_ull;                                  // [2*capa] neighbour rows of the residues (CSR by astart)
    unsigned long long* cmask_s = smem_ull + 2 * size_t(L.capa);            // [capc] bead-pair mask of candidate c (bit a*6+b)
    unsigned* cand_s = reinterpret_cast<unsigned*>(cmask_s + L.capc);       // [capc] A | B << 16   (capc is a multiple of 4)
    // entry offsets per (neighbour-row entry, own state), written by pass 1 of step 5 when masks and candidates are dead
    unsigned short* off_s = reinterpret_cast<unsigned short*>(cmask_s);     // [2*capa][6]  (12*capc >= 24*capa bytes)
    float4* bpos = reinterpret_cast<float4*>(cand_s + L.capc);              // [nb]
    float4* rc = bpos + nb;                                                 // [nR] bounding sphere (centre, radius)
    unsigned* bitmap = reinterpret_cast<unsigned*>(rc + nR);                // [nR][nW] adjacency, both residues multi-state
    unsigned* adj = bitmap + nR * nW;                                       // [nR][nW] adjacency, every active pair
    int* estart = reinterpret_cast<int*>(adj + nR * nW);                    // [nR+1] first matrix slot of pairs (A, B>A)
    int* istart = estart + nR + 1;                                          // [nR+1] incidence lists of the residue graph
    int* astart = istart + nR + 1;                                          // [nR+1] neighbour rows
    int* ebase = astart + nR + 1;                                           // [nR]
    int* rfirst = ebase + nR;                                               // [nR] first bead of the residue
    int* nrot = rfirst + nR;                                                // [nR]
    float* en = reinterpret_cast<float*>(nrot + nR);                        // [nR*6]
    int* rs = reinterpret_cast<int*>(en + nR * MAXR);                       // [nb+1]
    int* wtot = rs + nb + 1;                                                // [33]
    int* hist = wtot + 33;                                                  // [256]
    unsigned short* wpre = reinterpret_cast<unsigned short*>(hist + 256);   // [nR][nW] set bits of `bitmap` in the words before w
    unsigned short* awpre = wpre + nR * nW;                                 // [nR][nW] the same for `adj`
    unsigned short* lo_s = awpre + nR * nW;                                 // [nb]
    unsigned short* ce_s = lo_s + nb;                                       // [nb]
    unsigned short* rowner_s = ce_s + nb;                                   // [2*capa] residue that owns neighbour-row entry t

    // spill areas of this replica: [2*(capa_tot-capa)] rows | [capc_tot-capc] masks | [capc_tot-capc] candidates | offsets
    const size_t n_row_g = 2 * size_t(L.capa_tot - L.capa), n_c_g = size_t(L.capc_tot - L.capc);
    unsigned long long* sp = L.spill + size_t(r) * build_spill_words(L);
    const SpillArr<unsigned long long, SPILL> rown{rown_s, sp, 2 * L.capa};
    const SpillArr<unsigned long long, SPILL> cmask{cmask_s, sp + n_row_g, L.capc};
    const SpillArr<unsigned, SPILL> cand{cand_s, reinterpret_cast<unsigned*>(sp + n_row_g + n_c_g), L.capc};
    const SpillArr<unsigned short, SPILL> off{off_s, reinterpret_cast<unsigned short*>(sp + n_row_g + n_c_g + (n_c_g + 1) / 2), 12 * L.capa};
    const SpillArr<unsigned short, SPILL> rowner{rowner_s, reinterpret_cast<unsigned short*>(sp + n_row_g + n_c_g + (n_c_g + 1) / 2 + (6 * n_row_g + 3) / 4), 2 * L.capa};
    const int capc_use = SPILL ? L.capc_tot : L.capc, capa_use = SPILL ? L.capa_tot : L.capa;

    for (int i = tid; i < nR * nW; i += BUILD_TPB) { bitmap[i] = 0u; adj[i] = 0u; }
    for (int i = tid; i < nR * MAXR; i += BUILD_TPB) en[i] = 0.f;
    for (int A = tid; A < nR; A += BUILD_TPB) { nrot[A] = P.res_nrot[A]; rfirst[A] = P.res_first[A]; }
    __syncthreads();
    for (int i = tid; i < nb; i += BUILD_TPB) {
        const int loc = P.loc_identity ? i : P.g.s1.loc[i];   // (an index load in front of every gather is a second trip to memory)
        bpos[i] = *reinterpret_cast<const float4*>(P.g.s1.out + (size_t(r) * P.g.s1.n_node + loc) * P.g.s1.wp);
        float e = 0.f;
        for (int p = 0; p < P.n_prob; ++p) e += P.prob_out[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]];
        en[P.bead_res[i] * MAXR + P.bead_rot[i]] = e;   // one bead per state
    }
    __syncthreads();
    for (int A = tid; A < nR; A += BUILD_TPB) {   // bounding sphere: centre = mean of the beads
        const int f = rfirst[A], n = nrot[A];
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int a = 0; a < n; ++a) { const float4 p = bpos[f + a]; cx += p.x; cy += p.y; cz += p.z; }
        const float inv = 1.f / float(n);
        cx *= inv; cy *= inv; cz *= inv;
        float r2 = 0.f;
        for (int a = 0; a < n; ++a) {
            const float4 p = bpos[f + a];
            const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
            r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
        }
        rc[A] = make_float4(cx, cy, cz, sqrtf(r2));
    }
    __syncthreads();

    // ---- 1. sphere tests over all residue pairs, survivors in (A,B) order ------------------------------------------------
    const int nP = nR * (nR - 1) / 2;
    const int CH = min(32, max(1, (nP + BUILD_TPB - 1) / BUILD_TPB));
    const float reach = P.g.cutoff + 1e-3f;   // slack >> the rounding of centre, radius and distance
    int nc = 0;
    for (int q00 = 0; q00 < nP; q00 += BUILD_TPB * CH) {
        const int q0 = q00 + tid * CH;
        unsigned hits = 0u;
        int A0 = 0, B0 = 0;
        if (q0 < nP) {
            const float t = float(2 * nR - 1);
            int A = (int)((t - sqrtf(fmaxf(0.f, t * t - 8.f * float(q0)))) * 0.5f);
            A = max(0, min(A, nR - 2));
            while (A + 1 <= nR - 2 && ((A + 1) * (2 * nR - A - 2)) / 2 <= q0) ++A;
            while (A > 0 && (A * (2 * nR - A - 1)) / 2 > q0) --A;
            int B = q0 - (A * (2 * nR - A - 1)) / 2 + A + 1;
            A0 = A; B0 = B;
            float4 ca = rc[A];
            for (int k = 0; k < CH && q0 + k < nP; ++k) {
                const float4 cb = rc[B];
                const float dx = ca.x - cb.x, dy = ca.y - cb.y, dz = ca.z - cb.z;
                const float lim = reach + ca.w + cb.w;
                if (dx * dx + dy * dy + dz * dz < lim * lim) hits |= 1u << k;
                if (++B == nR) { ++A; B = A + 1; if (A < nR - 1) ca = rc[A]; }
            }
        }
        int tot;
        int at = nc + block_excl_scan(__popc(hits), wtot, tot);
        while (hits) {   // pair number q0 + k of the upper triangle, walked from (A0,B0)
            const int k = __ffs(hits) - 1;
            hits &= hits - 1;
            int A = A0, B = B0 + k;
            while (B >= nR) { B = B - nR + A + 2; ++A; }
            if (at < capc_use) cand[at] = (unsigned)A | ((unsigned)B << 16);
            ++at;
        }
        nc += tot;
    }
    __syncthreads();
    const bool cand_overflow = nc > capc_use;
    if (cand_overflow && !SPILL) return false;
    if (cand_overflow) nc = 0;

    // ---- 2. exact bead predicate per surviving residue pair ----------------------------------------------------------------
    const float cutoff2 = P.g.cutoff2;
    for (int c = tid; c < nc; c += BUILD_TPB) {
        const unsigned ab = cand[c];
        const int A = ab & 0xffffu, B = ab >> 16;
        const int nA = nrot[A], nB = nrot[B], fA = rfirst[A], fB = rfirst[B];
        // the beads of B in registers (absent states sit far away), then one unrolled row of six tests per bead of A
        float4 pj[MAXR];
#pragma unroll
        for (int b = 0; b < MAXR; ++b) pj[b] = b < nB ? bpos[fB + b] : make_float4(1e18f, 1e18f, 1e18f, 0.f);
        unsigned rows_lo = 0u, rows_hi = 0u;   // states 0-2 and 3-5 of A, six bits each
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            if (a < nA) {
                const float4 pi = bpos[fA + a];
                unsigned row = 0u;
#pragma unroll
                for (int b = 0; b < MAXR; ++b) {
                    const float dx = pi.x - pj[b].x, dy = pi.y - pj[b].y, dz = pi.z - pj[b].z;
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    row |= (d2 < cutoff2 ? 1u : 0u) << b;
                }
                if (a < 3) rows_lo |= row << (6 * a); else rows_hi |= row << (6 * (a - 3));
            }
        }
        const unsigned long long m = (unsigned long long)rows_lo | ((unsigned long long)rows_hi << 18);
        cmask[c] = m;
        if (m) {
            atomicOr(&adj[A * nW + (B >> 5)], 1u << (B & 31));
            atomicOr(&adj[B * nW + (A >> 5)], 1u << (A & 31));
            if (nA > 1 && nB > 1) {
                atomicOr(&bitmap[A * nW + (B >> 5)], 1u << (B & 31));
                atomicOr(&bitmap[B * nW + (A >> 5)], 1u << (A & 31));
            }
        }
    }
    __syncthreads();

    // ---- 3. per residue: running popcounts of both adjacency rows, degrees; scans ------------------------------------------
    // With them the position of B among the neighbours of A is wpre[A][B>>5] + popc(row_A[B>>5] below bit B) - two
    // independent loads - and the matrix slot of pair (A,B>A) is ebase[A] + that position.
    for (int A = tid; A < nR; A += BUILD_TPB) {
        int run = 0, below = 0, arun = 0;
        for (int w = 0; w < nW; ++w) {
            const unsigned bits = bitmap[A * nW + w], abits = adj[A * nW + w];
            wpre[A * nW + w] = (unsigned short)run;
            awpre[A * nW + w] = (unsigned short)arun;
            if (w == (A >> 5)) below = run + __popc(bits & ((1u << (A & 31)) - 1u));
            run += __popc(bits);
            arun += __popc(abits);
        }
        istart[A] = run;              // degree in the residue graph, scanned below
        ebase[A] = -below;
        estart[A] = run - below;      // upper degree
        astart[A] = arun;             // active degree
    }
    __syncthreads();
    if (tid < 32) {
        int per = (nR + 31) / 32, a0 = min(nR, tid * per), a1 = min(nR, a0 + per);
        int su = 0, sf = 0, sa = 0;
        for (int A = a0; A < a1; ++A) { su += estart[A]; sf += istart[A]; sa += astart[A]; }
        int pu = su, pf = sf, pa = sa;
        for (int o = 1; o < 32; o <<= 1) {
            int tu = __shfl_up_sync(UB_FULL_MASK, pu, o), tf = __shfl_up_sync(UB_FULL_MASK, pf, o), ta = __shfl_up_sync(UB_FULL_MASK, pa, o);
            if (tid >= o) { pu += tu; pf += tf; pa += ta; }
        }
        int eu = pu - su, ef = pf - sf, ea = pa - sa;
        for (int A = a0; A < a1; ++A) {
            const int up = estart[A], dg = istart[A], ad = astart[A];
            estart[A] = eu; istart[A] = ef; astart[A] = ea; ebase[A] += eu;
            eu += up; ef += dg; ea += ad;
        }
        if (tid == 31) { estart[nR] = pu; istart[nR] = pf; astart[nR] = pa; }
    }
    __syncthreads();
    const int n_pair = estart[nR], n_act2 = astart[nR];   // n_act2 = 2 x active pairs
    if (!SPILL && n_act2 > 2 * capa_use) return false;
    if (tid == 0) { P.stats[size_t(r) * 4 + 1] = n_pair; P.e11[r] = 0.f; }
    if (tid == 0 && r == 0) *P.n_slow = 0;
    if (*P.fe_flag) for (int A = tid; A < nR; A += BUILD_TPB) P.res_fe[size_t(r) * nR + A] = 0.f;
    int* rowstart = P.rowstart + size_t(r) * (nb + 1);
    auto fail = [&](int code) {   // uniform across the block: report, and leave an empty graph behind
        if (tid == 0) { atomicExch(P.error_flag, code); P.stats[size_t(r) * 4 + 1] = 0; }
        for (int A = tid; A <= nR; A += BUILD_TPB) P.istart[size_t(r) * (nR + 1) + A] = 0;
        for (int i = tid; i <= nb; i += BUILD_TPB) rowstart[i] = 0;
        for (int i = tid; i < nb; i += BUILD_TPB) {
            P.lower[size_t(r) * nb + i] = 0; P.order_e[size_t(r) * nb + i] = P.order_d[size_t(r) * nb + i] = (unsigned short)i;
        }
    };
    if (tid == 0 && L.bstats) { L.bstats[2 * r] = nc; L.bstats[2 * r + 1] = n_act2 / 2; }
    if (cand_overflow || n_act2 > 2 * capa_use) { fail(4); return true; }
    if (n_pair > P.max_pairs) { fail(2); return true; }

    // ---- 4. one thread per active pair: both neighbour-row entries, the matrix slot, both incidence-list entries -----------
    unsigned short* pair_ab = P.pair_ab + size_t(r) * P.max_pairs * 2;
    int* inc = P.inc + size_t(r) * 2 * P.max_pairs;
    for (int c = tid; c < nc; c += BUILD_TPB) {
        const unsigned long long m = cmask[c];
        if (!m) continue;
        const unsigned ab = cand[c];
        const int A = ab & 0xffffu, B = ab >> 16;
        unsigned long long mt = 0ull;   // the transpose: bit b*6+a
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const unsigned row = (unsigned)(m >> (a * 6)) & 63u;
#pragma unroll
            for (int b = 0; b < 6; ++b) mt |= (unsigned long long)((row >> b) & 1u) << (b * 6 + a);
        }
        const unsigned lowB = (1u << (B & 31)) - 1u, lowA = (1u << (A & 31)) - 1u;
        const int pa = (int)awpre[A * nW + (B >> 5)] + __popc(adj[A * nW + (B >> 5)] & lowB);   // B among A's neighbours
        const int pb = (int)awpre[B * nW + (A >> 5)] + __popc(adj[B * nW + (A >> 5)] & lowA);   // A among B's neighbours
        int slot = 0;
        if (nrot[A] > 1 && nrot[B] > 1) {
            const int ia = (int)wpre[A * nW + (B >> 5)] + __popc(bitmap[A * nW + (B >> 5)] & lowB);
            const int ib = (int)wpre[B * nW + (A >> 5)] + __popc(bitmap[B * nW + (A >> 5)] & lowA);
            const int e = ebase[A] + ia;
            slot = e;
            pair_ab[2 * e] = (unsigned short)A;
            pair_ab[2 * e + 1] = (unsigned short)B;
            inc[istart[A] + ia] = 2 * e;        // incidence lists list the neighbours in ascending order, as k_rot_prep does
            inc[istart[B] + ib] = 2 * e + 1;
        }
        rown[astart[A] + pa] = nbr_pack(m, true, B, slot);
        rown[astart[B] + pb] = nbr_pack(mt, false, A, slot);
        rowner[astart[A] + pa] = (unsigned short)A;
        rowner[astart[B] + pb] = (unsigned short)B;
    }
    for (int A = tid; A <= nR; A += BUILD_TPB) P.istart[size_t(r) * (nR + 1) + A] = istart[A];
    for (int i = tid; i < nR * MAXR; i += BUILD_TPB) P.enode[size_t(r) * nR * MAXR + i] = en[i];
    {
        float4* pm4 = reinterpret_cast<float4*>(P.pmat + size_t(r) * P.max_pairs * 36);
        for (int i = tid; i < n_pair * 9; i += BUILD_TPB) pm4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // ---- 5. CSR rows ------------------------------------------------------------------------------------------------------
    // pass 1, one thread per bead: walk the residue's neighbour row; row length, partners below, folding partners below,
    // evaluated entries - and for every neighbour the offset its entries get inside the bead's row (among the partners
    // above: position in the whole row; among those below: position inside the folding / non-folding group)
    for (int i = tid; i < nb; i += BUILD_TPB) {
        const int A = P.bead_res[i], a = P.bead_rot[i], a6 = 6 * a;
        const bool mA = nrot[A] > 1;
        int cnt = 0, lo = 0, nfl = 0, nev = 0;
        for (int t = astart[A]; t < astart[A + 1]; ++t) {
            const unsigned long long nr = rown[t];
            const int n = __popc((unsigned)(nr >> a6) & 63u);
            const bool mB = nrot[(int)(nr >> 37) & 0x7ff] > 1, below = !((nr >> 36) & 1ull), fold = mA && !mB;
            off[6 * t + a] = (unsigned short)(below ? (fold ? nfl : lo - nfl) : cnt);
            cnt += n;
            if (below) { lo += n; if (fold) { nfl += n; nev += n; } }
            else if (mA || !mB) nev += n;
        }
        rs[i] = cnt;   // scanned in place below
        lo_s[i] = (unsigned short)(lo - nfl);
        ce_s[i] = (unsigned short)nev;
    }
    __syncthreads();
    {   // exclusive scan of the row lengths in place (rs[nb] = total)
        const int per = (nb + BUILD_TPB - 1) / BUILD_TPB;
        const int r0 = min(nb, tid * per), r1 = min(nb, r0 + per);
        int s = 0;
        for (int i = r0; i < r1; ++i) s += rs[i];
        int tot;
        int run = block_excl_scan(s, wtot, tot);
        for (int i = r0; i < r1; ++i) { const int c = rs[i]; rs[i] = run; run += c; }
        if (tid == 0) rs[nb] = tot;
    }
    __syncthreads();
    if (rs[nb] > P.cap_e) { fail(3); return true; }
    for (int i = tid; i <= nb; i += BUILD_TPB) rowstart[i] = rs[i];
    unsigned short* dj = P.dj + size_t(r) * P.cap_e;
    int* code = P.code + size_t(r) * P.cap_e;
    // pass 2, one thread per (neighbour-row entry, own state): the up to six entries of one bead against one neighbouring
    // residue, at positions known in advance (offset recorded by pass 1 + rank of the partner state among the set bits), so
    // the stores are predicated, not branched over, and the lanes of a warp stay together.
    // Order inside a row as k_rot_prep: partners below that do not fold, those that fold, partners above ascending.
    for (int id = tid; id < 6 * n_act2; id += BUILD_TPB) {
        const int t = id / 6, a = id - 6 * t;
        const unsigned long long nr = rown[t];
        const unsigned pb = (unsigned)(nr >> (6 * a)) & 63u;
        if (!pb) continue;   // (also: states the owner does not have)
        const int A = rowner[t], C = (int)(nr >> 37) & 0x7ff, slot36 = 36 * (int)(nr >> 48);
        const bool up = (nr >> 36) & 1ull;
        const int fC = rfirst[C], i = rfirst[A] + a;
        const bool mA = nrot[A] > 1, mB = nrot[C] > 1, fold = mA && !mB;
        const int at = rs[i] + (int)off[6 * t + a] + ((!up && fold) ? (int)lo_s[i] : 0);
        // code of partner state b = cbase + b * cstep (matrix element / node whose marginal weights the entry)
        int cbase, cstep;
        if (mA && mB) { cbase = slot36 + (up ? a * 6 : a); cstep = up ? 1 : 6; }
        else if (mA) { cbase = code_node(A * MAXR + a, true); cstep = 0; }
        else if (mB) { cbase = code_node(C * MAXR, false); cstep = -2; }
        else { cbase = CODE_SS; cstep = 0; }
#pragma unroll
        for (int b = 0; b < MAXR; ++b) {
            const int where = at + __popc(pb & ((1u << b) - 1u));
            if (pb & (1u << b)) {
                dj[where] = (unsigned short)(fC + b);
                code[where] = cbase + b * cstep;
            }
        }
    }
    for (int i = tid; i < nb; i += BUILD_TPB) P.lower[size_t(r) * nb + i] = lo_s[i];
    // both row orders in one pass: two histograms side by side (keys clipped to 127), one set of barriers
    {
        int* hist_e = hist;          // [128]
        int* hist_d = hist + 128;    // [128]
        auto key_e = [&](int i) { return min((int)ce_s[i], 127); };
        auto key_d = [&](int i) { return min(rs[i + 1] - rs[i], 127); };
        for (int b = tid; b < 256; b += BUILD_TPB) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < nb; i += BUILD_TPB) { atomicAdd(&hist_e[key_e(i)], 1); atomicAdd(&hist_d[key_d(i)], 1); }
        __syncthreads();
        if (tid < 64) {   // warp 0: order_e, warp 1: order_d; offsets for falling keys (bin b starts after all larger bins)
            int* h = tid < 32 ? hist_e : hist_d;
            const int lane = tid & 31;
            int sum = 0, v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { v[u] = h[127 - (lane * 4 + u)]; sum += v[u]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += t; }
            int run = incl - sum;
#pragma unroll
            for (int u = 0; u < 4; ++u) { h[127 - (lane * 4 + u)] = run; run += v[u]; }
        }
        __syncthreads();
        unsigned short* oe = P.order_e + size_t(r) * nb;
        unsigned short* od = P.order_d + size_t(r) * nb;
        for (int i = tid; i < nb; i += BUILD_TPB) {
            oe[atomicAdd(&hist_e[key_e(i)], 1)] = (unsigned short)i;
            od[atomicAdd(&hist_d[key_d(i)], 1)] = (unsigned short)i;
        }
    }
    return true;
}

// the shared-memory-only body first; a replica that overflows its capacities (clashing start structures) is redone by the
// body whose arrays continue in global memory
// BUILD_TPB x OCC: 256 threads and four CTAs per SM for the shared-memory plan of ~100 residues; a large system, whose plan
// leaves room for one or two CTAs per SM, gets 1024 threads so that its 45 k sphere tests and 10 k entries do not queue
// behind eight warps
template <int BUILD_TPB, int OCC>
__global__ void __launch_bounds__(BUILD_TPB, OCC) k_rot_build(RotamerDev P, BuildLay L) {
    extern __shared__ unsigned long long smem_ull[];
    if (rot_build_body<false, BUILD_TPB>(P, L, smem_ull)) return;
    __syncthreads();
    rot_build_body<true, BUILD_TPB>(P, L, smem_ull);
}

// ================================================================================================ edge kernels
// row (t1,t2) of the symmetric-compressed table; swap = angular blocks exchanged (bead_interaction.h:209-218)
__device__ __forceinline__ int sym_row(int t1, int t2, int nT, bool& swap) {
    swap = t1 > t2;
    int a = swap ? t2 : t1, b = swap ? t1 : t2;
    return a * nT - (a * (a - 1)) / 2 + (b - a);
}

// the B-spline table is staged once per CTA (CTAs are persistent over replicas), beads once per replica.  `rowoff` maps an
// ordered type pair to (float offset of its row) | swap bit, so the per-pair table lookup is one shared-memory load.
__device__ __forceinline__ void stage_table(const RotamerDev& P, float* table, int* rowoff) {
    const int n4 = ((P.n_type * (P.n_type + 1) / 2) * P.g.n_param) / 4;   // n_param is even and rows come in pairs: multiple of 4
    const float4* src = reinterpret_cast<const float4*>(P.table);
    float4* dst = reinterpret_cast<float4*>(table);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    for (int i = n4 * 4 + threadIdx.x; i < (P.n_type * (P.n_type + 1) / 2) * P.g.n_param; i += blockDim.x) table[i] = P.table[i];
    for (int i = threadIdx.x; i < P.n_type * P.n_type; i += blockDim.x) {
        bool swap;
        int row = sym_row(i / P.n_type, i % P.n_type, P.n_type, swap);
        rowoff[i] = (row * P.g.n_param) << 1 | (swap ? 1 : 0);
    }
}
__device__ __forceinline__ void stage_beads(const RotamerDev& P, int r, BeadRec* beads, int t_in, int tpr) {
    for (int i = t_in; i < P.n_bead; i += tpr) {
        const float* p = elem_ptr(P.g.s1, r, i);
        float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        BeadRec br;
        br.x = a.x; br.y = a.y; br.z = a.z; br.dx = a.w; br.dy = b.x; br.dz = b.y;
        const int t = P.g.s1.type[i];
        br.type = (t * P.n_type) << 8 | t;     // rowoff index of (b1,b2) = (b1.type >> 8) + (b2.type & 0xff)
        br.res_rot = (P.bead_res[i] << 3) | P.bead_rot[i];
        beads[i] = br;
    }
}

// quadspline on shared-memory operands, (lo,hi) ordering as the reference's i1<i2 edge.  NKA/NK > 0: knot counts known at
// compile time (ff_1: 15 angular, 16 radial), all sub-table offsets become immediates; 0: read from P.q.
struct PairTab { const float* table; const int* rowoff; };
template <bool DERIV, int NKA, int NK>
__device__ __forceinline__ float pair_term(const RotamerDev& P, const PairTab& T, const BeadRec& b1, const BeadRec& b2, float* d1,
                                           float* d2) {
    const int nka = NKA ? NKA : P.q.nka, nk = NK ? NK : P.q.nk;
    const int ro = T.rowoff[(b1.type >> 8) + (b2.type & 0xff)];
    const bool swap = ro & 1;
    const float* row = T.table + (ro >> 1);
    const float* ang1 = row + (swap ? nka : 0);
    const float* ang2 = row + (swap ? 0 : nka);
    const float* wide = row + 2 * nka;
    const float* narrow = wide + nk;
    f3 displace = mk3(b2.x - b1.x, b2.y - b1.y, b2.z - b1.z);
    f3 rvec1 = mk3(b1.dx, b1.dy, b1.dz), rvec2 = mk3(b2.dx, b2.dy, b2.dz);
    float dist2 = mag2(displace);
    float inv_dist = rsqrtf(dist2);
    float dist_coord = dist2 * (inv_dist * P.q.inv_dx);
    f3 u = inv_dist * displace;
    float cos1 = dot(rvec1, u), cos2 = -dot(rvec2, u);
    float a1v, a1d, a2v, a2d, wv, wd, nv, nd;
    float w[4], d[4];
    {
        float x = (cos1 + 1.f) * P.q.inv_dtheta + 1.f;
        int b = max(1, min((int)x, nka - 3));
        bspline_weights(x - (float)b, w, d);
        bspline_apply(w, d, ang1[b - 1], ang1[b], ang1[b + 1], ang1[b + 2], a1v, a1d);
        x = (cos2 + 1.f) * P.q.inv_dtheta + 1.f;
        b = max(1, min((int)x, nka - 3));
        bspline_weights(x - (float)b, w, d);
        bspline_apply(w, d, ang2[b - 1], ang2[b], ang2[b + 1], ang2[b + 2], a2v, a2d);
    }
    {   // both radial splines share the knot interval and hence the weights (clamped rule of spline.h:275-310)
        float x = dist_coord;
        if (x < 1.f) {
            wv = (1.f / 6.f) * wide[0] + (2.f / 3.f) * wide[1] + (1.f / 6.f) * wide[2];
            nv = (1.f / 6.f) * narrow[0] + (2.f / 3.f) * narrow[1] + (1.f / 6.f) * narrow[2];
            wd = nd = 0.f;
        } else if (x >= (float)(nk - 2)) {
            wv = (1.f / 6.f) * wide[nk - 3] + (2.f / 3.f) * wide[nk - 2] + (1.f / 6.f) * wide[nk - 1];
            nv = (1.f / 6.f) * narrow[nk - 3] + (2.f / 3.f) * narrow[nk - 2] + (1.f / 6.f) * narrow[nk - 1];
            wd = nd = 0.f;
        } else {
            int b = (int)x;
            bspline_weights(x - (float)b, w, d);
            bspline_apply(w, d, wide[b - 1], wide[b], wide[b + 1], wide[b + 2], wv, wd);
            bspline_apply(w, d, narrow[b - 1], narrow[b], narrow[b + 1], narrow[b + 2], nv, nd);
        }
    }
    float angular_weight = a1v * a2v;
    if (DERIV) {
        float radial_deriv = P.q.inv_dx * (wd + angular_weight * nd);
        float ang_d1 = P.q.inv_dtheta * a1d * a2v * nv;
        float ang_d2 = P.q.inv_dtheta * a1v * a2d * nv;
        f3 rXX = ang_d1 * rvec1 - ang_d2 * rvec2;
        f3 deriv_dir = inv_dist * (rXX - dot(u, rXX) * u);
        f3 dd = radial_deriv * u + deriv_dir;
        d1[0] = -dd.x; d1[1] = -dd.y; d1[2] = -dd.z; d1[3] = ang_d1 * u.x; d1[4] = ang_d1 * u.y; d1[5] = ang_d1 * u.z;
        if (d2) { d2[0] = dd.x; d2[1] = dd.y; d2[2] = dd.z; d2[3] = -ang_d2 * u.x; d2[4] = -ang_d2 * u.y; d2[5] = -ang_d2 * u.z; }
    }
    return wv + angular_weight * nv;
}

constexpr int PF = 4;   // row entries fetched per step: the index/code/weight loads of a step are independent

// Edge kernels: ONE THREAD PER CSR ROW, rows taken in order of falling length (order_e / order_d from k_rot_prep) so that
// the threads of a warp walk rows of about the same length; a thread keeps its row's sums in registers, so nothing is
// reduced across lanes and every sum has a fixed order.  CTAs are persistent over replicas (blockIdx.y strides); with a
// small batch the rows of a replica are split over gridDim.x CTAs.
template <int NKA, int NK>
__global__ void __launch_bounds__(EDGE_MAXT, EDGE_OCC) k_rot_energy(RotamerDev P, int want_pot, int n_rep, int tpr) {
    extern __shared__ float4 smem4[];
    // a CTA serves G = blockDim.x / tpr replicas at a time: one staged table, G bead arrays, `tpr` threads (whole warps) each
    const int G = blockDim.x / tpr, g = threadIdx.x / tpr, t_in = threadIdx.x - g * tpr;
    BeadRec* beads = reinterpret_cast<BeadRec*>(smem4) + size_t(g) * P.n_bead;
    float* table = reinterpret_cast<float*>(reinterpret_cast<BeadRec*>(smem4) + size_t(G) * P.n_bead);
    int* rowoff = reinterpret_cast<int*>(table + (((P.n_type * (P.n_type + 1) / 2) * P.g.n_param + 3) & ~3));
    const PairTab T{table, rowoff};
    const int nb = P.n_bead;
    stage_table(P, table, rowoff);
    const bool fe_on = want_pot && *P.fe_flag;
    for (int rb = blockIdx.y * G; rb < n_rep; rb += gridDim.y * G) {
        const int r = rb + g;
        __syncthreads();   // previous replica's readers are done with `beads`
        if (r < n_rep) stage_beads(P, r, beads, t_in, tpr);
        __syncthreads();
        if (r >= n_rep) continue;
        const int* rowstart = P.rowstart + size_t(r) * (nb + 1);
        const unsigned short* dj = P.dj + size_t(r) * P.cap_e;
        const int* code = P.code + size_t(r) * P.cap_e;
        const unsigned short* order = P.order_e + size_t(r) * nb;
        const unsigned short* lower = P.lower + size_t(r) * nb;
        float* pmat = P.pmat + size_t(r) * P.max_pairs * 36;
        float e11 = 0.f;
        // boustrophedon over the length-sorted rows: a thread's long row of one pass is followed by a short one in the next
        for (int t0 = 0, pass = 0; t0 < nb; t0 += gridDim.x * tpr, ++pass) {
            const int tl = blockIdx.x * tpr + t_in;
            const int t = t0 + ((pass & 1) ? gridDim.x * tpr - 1 - tl : tl);
            if (t >= nb) continue;
            const int i = order[t];
            const BeadRec bi = beads[i];
            const int base = rowstart[i], c = rowstart[i + 1] - base;
            const int lo = lower[i];   // first entry to evaluate: folding partners below the bead, then the partners above
            float fold = 0.f;
            // software pipeline: the partner / code loads of the NEXT four entries are in flight while these four are evaluated
            // (the kernel sat on the long scoreboard: 6 stalled warps per issue, 38 % issue-slot utilisation)
            int js_n[PF], cds_n[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int k = lo + u;
                js_n[u] = k < c ? (int)dj[base + k] : -1;
                cds_n[u] = k < c ? code[base + k] : CODE_SS;
            }
            for (int k0 = lo; k0 < c; k0 += PF) {
                int js[PF], cds[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) { js[u] = js_n[u]; cds[u] = cds_n[u]; }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int k = k0 + PF + u;
                    js_n[u] = k < c ? (int)dj[base + k] : -1;
                    cds_n[u] = k < c ? code[base + k] : CODE_SS;
                }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int cd = cds[u];
                    if (js[u] < 0) continue;
                    if (cd == CODE_SS) { if (!want_pot) continue; }
                    else if (cd < 0 && !code_is_fold(cd)) continue;   // (single, multi): handled from the partner's row
                    const float V = pair_term<false, NKA, NK>(P, T, bi, beads[js[u]], nullptr, nullptr);
                    if (cd >= 0) { if (P.multi_bead_states) atomicAdd(&pmat[cd], V); else pmat[cd] = V; }
                    else if (cd == CODE_SS) {
                        e11 += V;
                        if (fe_on) {   // edges11: half to each residue (rotamer.cpp:877-881)
                            atomicAdd(&P.res_fe[size_t(r) * P.n_res + (bi.res_rot >> 3)], 0.5f * V);
                            atomicAdd(&P.res_fe[size_t(r) * P.n_res + (beads[js[u]].res_rot >> 3)], 0.5f * V);
                        }
                    }
                    else fold += V;
                }
            }
            P.fold[size_t(r) * nb + i] = fold;
        }
        if (want_pot) {
            e11 = warp_sum(e11);
            if ((threadIdx.x & 31) == 0 && e11 != 0.f) atomicAdd(&P.e11[r], e11);
        }
    }
}

// (Evaluating every pair once and adding the partner's half into shared-memory accumulators was measured at 663-679 us per
// launch against 358 us for this gather form: sm_100a has no native shared-memory float add, atomicAdd compiles to an
// ATOMS.CAST.SPIN compare-and-swap loop.)
template <int NKA, int NK>
__global__ void __launch_bounds__(EDGE_MAXT, EDGE_OCC) k_rot_deriv(RotamerDev P, int n_rep, int tpr) {
    extern __shared__ float4 smem4[];
    const int G = blockDim.x / tpr, g = threadIdx.x / tpr, t_in = threadIdx.x - g * tpr;
    BeadRec* beads = reinterpret_cast<BeadRec*>(smem4) + size_t(g) * P.n_bead;
    float* table = reinterpret_cast<float*>(reinterpret_cast<BeadRec*>(smem4) + size_t(G) * P.n_bead);
    int* rowoff = reinterpret_cast<int*>(table + (((P.n_type * (P.n_type + 1) / 2) * P.g.n_param + 3) & ~3));
    const PairTab T{table, rowoff};
    const int nb = P.n_bead;
    stage_table(P, table, rowoff);
    for (int rb = blockIdx.y * G; rb < n_rep; rb += gridDim.y * G) {
        const int r = rb + g;
        __syncthreads();
        if (r < n_rep) stage_beads(P, r, beads, t_in, tpr);
        __syncthreads();
        if (r >= n_rep) continue;
        const int* rowstart = P.rowstart + size_t(r) * (nb + 1);
        const unsigned short* dj = P.dj + size_t(r) * P.cap_e;
        const float* ssr = P.ss + size_t(r) * P.cap_e;
        const unsigned short* order = P.order_d + size_t(r) * nb;
        const float* nm = P.node_marg + size_t(r) * P.n_res * MAXR;
        // boustrophedon over the length-sorted rows: a thread's long row of one pass is followed by a short one in the next
        for (int t0 = 0, pass = 0; t0 < nb; t0 += gridDim.x * tpr, ++pass) {
            const int tl = blockIdx.x * tpr + t_in;
            const int t = t0 + ((pass & 1) ? gridDim.x * tpr - 1 - tl : tl);
            if (t >= nb) continue;
            const int i = order[t];
            const BeadRec bi = beads[i];
            const int base = rowstart[i], c = rowstart[i + 1] - base;
            // the read half of the read-modify-writes is issued before the row walk hides its latency
            float* dst = elem_sens_ptr(P.g.s1, r, i);
            float4 old_a = reinterpret_cast<const float4*>(dst)[0], old_b = reinterpret_cast<const float4*>(dst)[1];
            const int loc = P.g.s1.loc[i];
            float old_p[MAX_PROB_NODES] = {0.f, 0.f, 0.f, 0.f};
            for (int p = 0; p < P.n_prob; ++p) old_p[p] = P.prob_sens[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]];
            const float my_marg = nm[(bi.res_rot >> 3) * MAXR + (bi.res_rot & 7)];
            float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int k0 = 0; k0 < c; k0 += PF) {
                int js[PF];
                float ss[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int k = k0 + u;
                    js[u] = k < c ? (int)dj[base + k] : -1;
                    ss[u] = k < c ? ssr[base + k] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    if (js[u] < 0) continue;
                    // The pair term is symmetric under exchange of its operands down to the last bit - the displacement and
                    // every quantity odd in it change sign exactly, the angular tables trade places through the row-offset
                    // table - so the row's bead always goes first and only ITS derivative is formed (no operand ordering by
                    // index, no selects, no partner half: -30 instructions per entry)
                    float d1[6];
                    pair_term<true, NKA, NK>(P, T, bi, beads[js[u]], d1, nullptr);
#pragma unroll
                    for (int q = 0; q < 6; ++q) acc[q] += ss[u] * d1[q];
                }
            }
            old_a.x += acc[0]; old_a.y += acc[1]; old_a.z += acc[2]; old_a.w += acc[3]; old_b.x += acc[4]; old_b.y += acc[5];
            reinterpret_cast<float4*>(dst)[0] = old_a;
            reinterpret_cast<float4*>(dst)[1] = old_b;
            for (int p = 0; p < P.n_prob; ++p) P.prob_sens[p][(size_t(r) * P.prob_n[p] + loc) * P.prob_wp[p]] = old_p[p] + my_marg;
        }
    }
}

// weight of every CSR entry in the backward pass (propagate_derivatives, rotamer.cpp:956-985): the pair marginal for two
// multi-state residues, a node marginal when one of them has a single state, 1 for two single-state residues.
// pair_marg(code) and node_marg(node) read the marginals wherever the calling BP kernel holds them.
template <typename PairF, typename NodeF>
__device__ __forceinline__ void emit_entry_weights(const RotamerDev& P, int r, int n_thread, PairF pair_marg, NodeF node_marg) {
    const int n_ent = P.rowstart[size_t(r) * (P.n_bead + 1) + P.n_bead];
    const int* code = P.code + size_t(r) * P.cap_e;
    float* ss = P.ss + size_t(r) * P.cap_e;
    // four entries per step: the code loads of a step are independent, so their global-memory latency overlaps (one load per
    // step left this loop waiting on the long scoreboard for 10 % of the BP kernel's time)
    for (int e0 = threadIdx.x; e0 < n_ent; e0 += 4 * n_thread) {
        int cd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cd[u] = e0 + u * n_thread < n_ent ? code[e0 + u * n_thread] : CODE_SS;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (e0 + u * n_thread < n_ent)
                ss[e0 + u * n_thread] = cd[u] >= 0 ? pair_marg(cd[u]) : (cd[u] == CODE_SS ? 1.f : node_marg((-2 - cd[u]) >> 1));
    }
}

// ================================================================================================ belief propagation
__device__ __forceinline__ float block_max_bcast(float v, float* red) {
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float m = red[0];
    for (int k = 1; k < BP_TPB / 32; ++k) m = fmaxf(m, red[k]);
    return m;
}

// Pair matrices and messages are addressed as (pair e, component k) -> e*se + k*sk: component-major in shared memory
// (se=1: consecutive threads = consecutive pairs hit consecutive banks), pair-major in the global spill area (sk=1).
struct Lay { int se, sk; };
__device__ __forceinline__ int at(const Lay& l, int e, int k) { return e * l.se + k * l.sk; }

// messages of every residue pair from old beliefs and old messages, in place (rotamer.cpp:468-520): the new message to A
// needs only the old message to B of the same pair, so one thread updates both directions of its pair without a copy
__device__ __forceinline__ void bp_messages(const int* res_nrot, int n_pair, const unsigned short* pair_ab, const float* Pm,
                                            Lay lp, const float* bel, float* msg, Lay lm) {
    for (int e = threadIdx.x; e < n_pair; e += BP_TPB) {
        int A = pair_ab[2 * e], B = pair_ab[2 * e + 1];
        int nA = res_nrot[A], nB = res_nrot[B];
        float v1[MAXR], v2[MAXR], m1[MAXR], m2[MAXR];
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            v1[a] = a < nA ? __fdividef(bel[A * MAXR + a], 1e-10f + msg[at(lm, e, a)]) : 0.f;
            v2[a] = a < nB ? __fdividef(bel[B * MAXR + a], 1e-10f + msg[at(lm, e, 6 + a)]) : 0.f;
            m1[a] = 0.f; m2[a] = 0.f;
        }
#pragma unroll
        for (int a = 0; a < MAXR; ++a)
#pragma unroll
            for (int b = 0; b < MAXR; ++b) {
                float p = Pm[at(lp, e, a * 6 + b)];   // entries outside (nA,nB) meet a zero cavity factor
                m1[a] = fmaf(p, v2[b], m1[a]);         // apply_left : message to A
                m2[b] = fmaf(v1[a], p, m2[b]);         // apply_right: message to B
            }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int a = 0; a < MAXR; ++a) { m1[a] = a < nA ? m1[a] : 0.f; m2[a] = a < nB ? m2[a] : 0.f; s1 += m1[a]; s2 += m2[a]; }
        float i1 = __fdividef(1.f, s1), i2 = __fdividef(1.f, s2);
#pragma unroll
        for (int a = 0; a < MAXR; ++a) { msg[at(lm, e, a)] = m1[a] * i1; msg[at(lm, e, 6 + a)] = m2[a] * i2; }
    }
}

// node update in place: belief = prob * prod(incoming messages), max-normalised and damped (rotamer.cpp:488-499,258-273)
__device__ __forceinline__ float bp_nodes(const int* res_nrot, int n_res, const int* istart, const int* inc, const float* prob,
                                          const float* msg, Lay lm, float* bel, float damping) {
    float dev = 0.f;
    for (int A = threadIdx.x; A < n_res; A += BP_TPB) {
        int nA = res_nrot[A];
        if (nA < 2) continue;
        float b[MAXR];
#pragma unroll
        for (int a = 0; a < MAXR; ++a) b[a] = prob[A * MAXR + a];
        int t0 = istart[A], t1 = istart[A + 1];
        // four incident pairs per step: the index loads and the 24 message loads of a step are independent, so the
        // serial chain per node is one shared-memory round trip per four pairs instead of two per pair
        for (int t = t0; t < t1; t += 4) {
            int cd[4];
            float m[4][MAXR];
#pragma unroll
            for (int u = 0; u < 4; ++u) cd[u] = t + u < t1 ? inc[t + u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int a = 0; a < MAXR; ++a) m[u][a] = cd[u] >= 0 ? msg[at(lm, cd[u] >> 1, (cd[u] & 1) * 6 + a)] : 1.f;
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < MAXR; ++a) { b[a] *= (m[0][a] * m[1][a]) * (m[2][a] * m[3][a]); s += b[a]; }
            // renormalise (pure rescaling, rotamer.cpp:492-493): a product of four L1-normalised messages cannot underflow
            float is = __fdividef(1.f, s);
#pragma unroll
            for (int a = 0; a < MAXR; ++a) b[a] *= is;
        }
        float mx = b[0];
#pragma unroll
        for (int a = 1; a < MAXR; ++a) mx = fmaxf(mx, b[a]);
        float imx = __fdividef(1.f, mx);
#pragma unroll
        for (int a = 0; a < MAXR; ++a) {
            float o = bel[A * MAXR + a];
            float n = (damping != 0.f) ? (1.f - damping) * imx * b[a] + damping * o : imx * b[a];
            if (a < nA) dev = fmaxf(dev, n - o);
            bel[A * MAXR + a] = n;
        }
    }
    return dev;
}

// General path: any state count <= 6, any pair count (spills to global memory).  With `use_list` the CTAs stride over the
// replicas the fast kernel declined (slow_list); otherwise CTA r solves replica r.
__device__ void bp_solve_generic(const RotamerDev& P, int want_pot, int r, float* smem);
__global__ void __launch_bounds__(BP_TPB) k_rot_bp(RotamerDev P, int want_pot, int use_list) {
    extern __shared__ float smem[];
    if (!use_list) { bp_solve_generic(P, want_pot, blockIdx.x, smem); return; }
    const int n = *P.n_slow;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        __syncthreads();
        bp_solve_generic(P, want_pot, P.slow_list[i], smem);
    }
}
__device__ void bp_solve_generic(const RotamerDev& P, int want_pot, int r, float* smem) {
    const int tid = threadIdx.x;
    const int nR = P.n_res, SP = P.smem_pairs;
    const bool fe_on = want_pot && *P.fe_flag;
    const int n_pair = P.stats[size_t(r) * 4 + 1];
    float* prob = smem;                         // [nR][6]
    float* bel = prob + nR * MAXR;              // [nR][6]
    float* offs = bel + nR * MAXR;              // [nR]
    float* red = offs + nR;                     // [32]
    int* istart = reinterpret_cast<int*>(red + 32);           // [nR+1]
    int* nrot = istart + nR + 1;                               // [nR]
    float* sm_msg = reinterpret_cast<float*>(nrot + nR);      // [12][SP]
    float* sm_P = sm_msg + size_t(SP) * 12;                   // [36][SP]
    int* sm_inc = reinterpret_cast<int*>(sm_P + size_t(SP) * 36);           // [2*SP]
    unsigned short* sm_ab = reinterpret_cast<unsigned short*>(sm_inc + 2 * SP);   // [2*SP]

    float* g_pmat = P.pmat + size_t(r) * P.max_pairs * 36;
    const bool in_smem = n_pair <= SP;
    // replicas whose pair count exceeds the shared-memory budget run the same code on their global scratch; their
    // messages spill to the area allocated behind the node marginals
    float* Pm = in_smem ? sm_P : g_pmat;
    float* msg = in_smem ? sm_msg : P.node_marg + size_t(P.n_rep) * nR * MAXR + size_t(r) * P.max_pairs * 12;
    const Lay lp = in_smem ? Lay{1, SP} : Lay{36, 1};
    const Lay lm = in_smem ? Lay{1, SP} : Lay{12, 1};
    const unsigned short* pair_ab = P.pair_ab + size_t(r) * P.max_pairs * 2;
    const int* inc = P.inc + size_t(r) * 2 * P.max_pairs;
    float* node_marg = P.node_marg + size_t(r) * nR * MAXR;

    // ---- node energies -> probabilities (convert_energy_to_prob :239-256; single-state partners already folded) ----------
    for (int i = tid; i < nR; i += BP_TPB) nrot[i] = P.res_nrot[i];
    for (int i = tid; i <= nR; i += BP_TPB) istart[i] = P.istart[size_t(r) * (nR + 1) + i];
    for (int i = tid; i < nR * MAXR; i += BP_TPB) bel[i] = P.enode[size_t(r) * nR * MAXR + i];   // 1-body energy
    __syncthreads();
    for (int A = tid; A < nR; A += BP_TPB) {   // energy offset = smallest 1-body energy
        float m = bel[A * MAXR];
        for (int a = 1; a < nrot[A]; ++a) m = fminf(m, bel[A * MAXR + a]);
        offs[A] = m;
    }
    __syncthreads();
    for (int i = tid; i < P.n_bead; i += BP_TPB) {
        float f = P.fold[size_t(r) * P.n_bead + i];
        if (f != 0.f) { float* b = &bel[P.bead_res[i] * MAXR + P.bead_rot[i]]; if (P.multi_bead_states) atomicAdd(b, f); else *b += f; }
    }
    __syncthreads();
    for (int i = tid; i < nR * MAXR; i += BP_TPB) {
        int A = i / MAXR, a = i % MAXR;
        float p = a < nrot[A] ? __expf(offs[A] - bel[i]) : 0.f;
        prob[i] = p;
    }
    {   // pair energies -> probabilities, 16-byte loads (36 floats per pair = 9 float4, rows are 16-byte aligned)
        const float4* src = reinterpret_cast<const float4*>(g_pmat);
        for (int i = tid; i < n_pair * 9; i += BP_TPB) {
            float4 v = src[i];
            int e = i / 9, k = (i % 9) * 4;
            Pm[at(lp, e, k)] = __expf(-v.x); Pm[at(lp, e, k + 1)] = __expf(-v.y);
            Pm[at(lp, e, k + 2)] = __expf(-v.z); Pm[at(lp, e, k + 3)] = __expf(-v.w);
        }
    }
    if (in_smem) {
        for (int i = tid; i < 2 * n_pair; i += BP_TPB) { sm_inc[i] = inc[i]; sm_ab[i] = pair_ab[i]; }
        inc = sm_inc;
        pair_ab = sm_ab;
    }
    __syncthreads();
    for (int i = tid; i < nR * MAXR; i += BP_TPB) bel[i] = prob[i];
    for (int e = tid; e < n_pair; e += BP_TPB) {
        int nA = nrot[pair_ab[2 * e]], nB = nrot[pair_ab[2 * e + 1]];
        for (int a = 0; a < 6; ++a) { msg[at(lm, e, a)] = a < nA ? 1.f : 0.f; msg[at(lm, e, 6 + a)] = a < nB ? 1.f : 0.f; }
    }
    __syncthreads();
    // ---- belief propagation -----------------------------------------------------------------------------------------------
    // initial sweep: first messages from (prob, unit messages); node beliefs restart from prob/max (rotamer.cpp:1034)
    bp_messages(nrot, n_pair, pair_ab, Pm, lp, bel, msg, lm);
    __syncthreads();
    for (int A = tid; A < nR; A += BP_TPB) {
        float mx = prob[A * MAXR];
        for (int a = 1; a < MAXR; ++a) mx = fmaxf(mx, prob[A * MAXR + a]);
        float imx = 1.f / mx;
        for (int a = 0; a < MAXR; ++a) bel[A * MAXR + a] = prob[A * MAXR + a] * imx;
    }
    __syncthreads();
    float max_dev = 1e10f;
    int iter = 0;
    for (; max_dev > P.tol && iter < P.max_iter; iter += P.chunk) {
        float dev = 0.f;
        for (int j = 0; j < P.chunk; ++j) {
            bp_messages(nrot, n_pair, pair_ab, Pm, lp, bel, msg, lm);
            __syncthreads();
            dev = bp_nodes(nrot, nR, istart, inc, prob, msg, lm, bel, P.damping);
            __syncthreads();
        }
        max_dev = block_max_bcast(dev, red);
    }
    if (tid == 0) {
        int* st = P.stats + size_t(r) * 4;
        st[0] = iter; st[2] = max_dev <= P.tol;
        if (iter >= P.max_iter - P.chunk - 1) P.n_bad[r] += 1;   // the reference's criterion (rotamer.cpp:784-785)
    }
    // ---- marginals (and Bethe free energy) ----------------------------------------------------------------------------------
    float en = 0.f;
    for (int A = tid; A < nR; A += BP_TPB) {
        int nA = nrot[A];
        float b[MAXR], s = 0.f;
        for (int a = 0; a < MAXR; ++a) { b[a] = nA > 1 ? bel[A * MAXR + a] : (a == 0 ? 1.f : 0.f); s += b[a]; }
        float is = 1.f / s;
        for (int a = 0; a < MAXR; ++a) { b[a] *= is; bel[A * MAXR + a] = b[a]; node_marg[A * MAXR + a] = b[a]; }
        if (want_pot) {
            float e = offs[A];
            for (int a = 0; a < nA; ++a) e += b[a] * __logf((1e-10f + b[a]) / (1e-10f + prob[A * MAXR + a]));
            en += e;
            if (fe_on) atomicAdd(&P.res_fe[size_t(r) * nR + A], e);
        }
    }
    __syncthreads();
    for (int e = tid; e < n_pair; e += BP_TPB) {
        int A = pair_ab[2 * e], Bq = pair_ab[2 * e + 1];
        int nA = nrot[A], nB = nrot[Bq];
        float bc1[MAXR], bc2[MAXR];
        for (int a = 0; a < MAXR; ++a) {
            bc1[a] = a < nA ? bel[A * MAXR + a] / (1e-10f + msg[at(lm, e, a)]) : 0.f;
            bc2[a] = a < nB ? bel[Bq * MAXR + a] / (1e-10f + msg[at(lm, e, 6 + a)]) : 0.f;
        }
        float* out = g_pmat + size_t(e) * 36;
        float s = 0.f;
        for (int a = 0; a < nA; ++a) for (int b = 0; b < nB; ++b) s += Pm[at(lp, e, a * 6 + b)] * bc1[a] * bc2[b];
        float is = 1.f / s, en_pair = 0.f;
        for (int a = 0; a < MAXR; ++a)
            for (int b = 0; b < MAXR; ++b) {
                float pr = Pm[at(lp, e, a * 6 + b)];
                float mg = (a < nA && b < nB) ? pr * bc1[a] * bc2[b] * is : 0.f;
                if (want_pot && a < nA && b < nB)
                    en_pair += mg * __logf((1e-10f + mg) / (1e-10f + pr * bel[A * MAXR + a] * bel[Bq * MAXR + b]));
                out[a * 6 + b] = mg;
            }
        en += en_pair;
        if (fe_on) { atomicAdd(&P.res_fe[size_t(r) * nR + A], 0.5f * en_pair); atomicAdd(&P.res_fe[size_t(r) * nR + Bq], 0.5f * en_pair); }
    }
    __syncthreads();   // this CTA's global writes of the pair marginals are visible to all its threads
    emit_entry_weights(P, r, BP_TPB, [&](int cd) { return g_pmat[cd]; }, [&](int node) { return bel[node]; });
    if (want_pot) {
        float tot = block_sum(en, red);
        if (tid == 0) P.potential[r] = tot + P.e11[r];
    }
}

// ================================================================================================ belief propagation, fast path
// k_rot_bp2: the solver of k_rot_bp re-laid for throughput.  Residue pairs are oriented (fewer states first) and sorted
// by class (6x6 | 3x6 | 3x3; the reference keeps separate 33/36/66 edge holders, rotamer.cpp:527-539), so that
//   * pair matrices occupy 36/18/9 floats instead of a padded 36 (three resident CTAs per SM instead of two),
//   * a warp of the message phase runs ONE unrolled nA x nB body without padding work,
// and the node update runs 8 lanes per residue (one per state) instead of one thread per residue, which cuts the serial
// chain per sweep to one shared-memory round trip per four incident pairs.  The sweep-end barrier doubles as the
// convergence vote (__syncthreads_or).  Replicas that do not fit the shared-memory budget (or configurations with state
// counts other than 1/3/6) are appended to slow_list and solved by k_rot_bp.
#ifndef UB_BP2_TPB
#define UB_BP2_TPB 384
#endif
#ifndef UB_BP2_WANT10
#define UB_BP2_WANT10 38   // pair capacity the fast kernel must offer, per residue, x10
#endif
#ifndef UB_BP2_OCC
#define UB_BP2_OCC 3
#endif
constexpr int BP2_TPB = UB_BP2_TPB;
constexpr int BP2_OCC = UB_BP2_OCC;

struct Bp2Lay {
    int nRp;     // residue count of the node-major probability / belief arrays
    int SP;      // pair capacity (== 4 mod 32)
    int Pcap;    // floats available for pair matrices
};

// Pair matrices and messages are PAIR-major with odd strides (NA*NB + 1 for the 6x6 and 3x6 blocks, 9 for 3x3, MSG_STRIDE =
// 13 for the twelve message components): a thread's own block sits at one base address, so every element is a load with
// an immediate offset (component-major blocks cost one address instruction per element), and consecutive threads still hit
// different banks because the stride is coprime with 32.
constexpr int MSG_STRIDE = 13;
constexpr int NODE_STRIDE = 7;   // probabilities and beliefs node-major: six states + one word of padding (odd stride)
template <int NA, int NB>
__device__ __forceinline__ void bp2_pair(const float* __restrict__ bel, int nRp, int F, int S, float* __restrict__ msg_p,
                                         const float* __restrict__ Pq) {
    float v1[NA], v2[NB], m1[NA], m2[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) { v1[a] = bel[(F) * NODE_STRIDE + a] * rcp_fast(1e-10f + msg_p[a]); m1[a] = 0.f; }
#pragma unroll
    for (int b = 0; b < NB; ++b) { v2[b] = bel[(S) * NODE_STRIDE + b] * rcp_fast(1e-10f + msg_p[6 + b]); m2[b] = 0.f; }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float pr = Pq[a * NB + b];
            m1[a] = fmaf(pr, v2[b], m1[a]);   // apply_left : message to the first residue
            m2[b] = fmaf(v1[a], pr, m2[b]);   // apply_right: message to the second residue
        }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) s1 += m1[a];
#pragma unroll
    for (int b = 0; b < NB; ++b) s2 += m2[b];
    float i1 = rcp_fast(s1), i2 = rcp_fast(s2);
#pragma unroll
    for (int a = 0; a < NA; ++a) msg_p[a] = m1[a] * i1;
#pragma unroll
    for (int b = 0; b < NB; ++b) msg_p[6 + b] = m2[b] * i2;
}

// pair marginal (rotamer.cpp:403-429), in place of the pair's probability matrix in shared memory, plus its Bethe term
// (:431-451)
template <int NA, int NB>
__device__ __forceinline__ float bp2_pair_marginal(const float* __restrict__ bel, int nRp, int F, int S, const float* __restrict__ msg_p,
                                                   float* __restrict__ Pq, int want_pot) {
    float bc1[NA], bc2[NB], b1[NA], b2[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) { b1[a] = bel[(F) * NODE_STRIDE + a]; bc1[a] = b1[a] / (1e-10f + msg_p[a]); }
#pragma unroll
    for (int b = 0; b < NB; ++b) { b2[b] = bel[(S) * NODE_STRIDE + b]; bc2[b] = b2[b] / (1e-10f + msg_p[6 + b]); }
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) s += Pq[a * NB + b] * bc1[a] * bc2[b];
    float is = 1.f / s, en = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float pr = Pq[a * NB + b];
            float mg = pr * bc1[a] * bc2[b] * is;
            if (want_pot) en += mg * __logf((1e-10f + mg) / (1e-10f + pr * b1[a] * b2[b]));
            Pq[a * NB + b] = mg;
        }
    return en;
}

// node update: belief = prob * prod(incoming messages), max-normalised and damped (rotamer.cpp:488-499,258-273).  One thread
// per residue; four incident pairs per step so that the index loads and the 4*NA message loads of a step are independent
// (one shared-memory round trip per four pairs).  inc2 holds word offsets of the messages' component-0 entries; entries
// past the end of the list read the column of ones.
template <int NA>
__device__ __forceinline__ float bp2_node(int A, const float* __restrict__ prob, float* __restrict__ bel, int nRp, int t0, int t1,
                                          const int* __restrict__ inc2, const float* __restrict__ msg, int dummy, float damping) {
    float b[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) b[a] = prob[(A) * NODE_STRIDE + a];
    for (int t = t0; t < t1; t += 4) {
        int off[4];
        float m[4][NA];
#pragma unroll
        for (int u = 0; u < 4; ++u) off[u] = t + u < t1 ? inc2[t + u] : dummy;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int a = 0; a < NA; ++a) m[u][a] = msg[off[u] + a];
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) { b[a] *= (m[0][a] * m[1][a]) * (m[2][a] * m[3][a]); s += b[a]; }
        // renormalise (pure rescaling, rotamer.cpp:492-493): a product of four L1-normalised messages cannot underflow
        const float is = __fdividef(1.f, s);
#pragma unroll
        for (int a = 0; a < NA; ++a) b[a] *= is;
    }
    float mx = b[0];
#pragma unroll
    for (int a = 1; a < NA; ++a) mx = fmaxf(mx, b[a]);
    const float imx = __fdividef(1.f, mx);
    float dev = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        const float o = bel[(A) * NODE_STRIDE + a];
        const float n = (damping != 0.f) ? (1.f - damping) * imx * b[a] + damping * o : imx * b[a];
        dev = fmaxf(dev, n - o);
        bel[(A) * NODE_STRIDE + a] = n;
    }
    return dev;
}

// TPB / OCC: 384 threads x 3 CTAs per SM where the pair capacity of a replica fits a third of the shared memory (100 residues);
// larger systems get 512 x 2 or 1024 x 1, so that a CTA that has the SM to itself still fills it with warps
template <int TPB, int OCC>
__global__ void __launch_bounds__(TPB, OCC) k_rot_bp2(RotamerDev P, Bp2Lay L, int want_pot) {
    constexpr int BP2_TPB = TPB;   // (shadows the default block size)
    extern __shared__ float smem[];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int nR = P.n_res, nRp = L.nRp, SP = L.SP;
    const bool fe_on = want_pot && *P.fe_flag;
    const int n_pair = P.stats[size_t(r) * 4 + 1];
    float* prob = smem;                          // [nRp][7] node-major
    float* bel = prob + NODE_STRIDE * nRp;       // [nRp][7]
    float* offs = bel + NODE_STRIDE * nRp;       // [nR]
    float* red = offs + nR;                      // [32]
    int* istart = reinterpret_cast<int*>(red + 32);     // [nR+1]
    int* nrot = istart + nR + 1;                         // [nR]
    unsigned long long* wscan = reinterpret_cast<unsigned long long*>(nrot + nR + ((nR + 1) & 1));   // [34], 8-byte aligned
    float* msg = reinterpret_cast<float*>(wscan + 34);   // [SP][13] pair-major: side 0 in 0..5, side 1 in 6..11
    float* Pm = msg + MSG_STRIDE * SP;                   // [Pcap]: 6x6 block [n66][37] | 3x6 block [n36][19] | 3x3 block [n33][9]
    int* inc2 = reinterpret_cast<int*>(Pm + L.Pcap);     // [2*SP] word offset of component 0 of each incident message
    unsigned short* pos = reinterpret_cast<unsigned short*>(inc2 + 2 * SP);   // [SP] slot e -> sorted position p
    unsigned short* perm = pos + SP;                     // [SP] p -> e
    unsigned short* fs = perm + SP;                      // [2*SP] first/second residue of p, bit 15 of fs[2p] = swapped
    unsigned short* nlist = fs + 2 * SP;                 // [nR] multi-state residues: 6-state first, then by falling degree
    __shared__ int n_multi_s;

    const unsigned short* pair_ab = P.pair_ab + size_t(r) * P.max_pairs * 2;
    const int* inc = P.inc + size_t(r) * 2 * P.max_pairs;
    float* g_pmat = P.pmat + size_t(r) * P.max_pairs * 36;
    float* node_marg = P.node_marg + size_t(r) * nR * MAXR;
    int* st = P.stats + size_t(r) * 4;

    if (tid == 0) n_multi_s = 0;
    for (int i = tid; i < nR; i += BP2_TPB) nrot[i] = P.res_nrot[i];
    for (int i = tid; i <= nR; i += BP2_TPB) istart[i] = P.istart[size_t(r) * (nR + 1) + i];
    for (int i = tid; i < NODE_STRIDE * nRp; i += BP2_TPB) { bel[i] = 0.f; prob[i] = 0.f; }
    __syncthreads();

    // ---- orient and sort the pairs by class: one packed scan (three 21-bit counters) -------------------------------------
    int n66, n36, n33;
    {
        const int lane = tid & 31, w = tid >> 5;
        const int per = (n_pair + BP2_TPB - 1) / BP2_TPB;
        const int e0 = min(n_pair, tid * per), e1 = min(n_pair, e0 + per);
        auto cls = [&](int e) {
            int nA = nrot[pair_ab[2 * e]], nB = nrot[pair_ab[2 * e + 1]];
            return (nA == 6 && nB == 6) ? 0 : ((nA == 3 && nB == 3) ? 2 : 1);
        };
        unsigned long long s = 0ull;
        for (int e = e0; e < e1; ++e) s += 1ull << (21 * cls(e));
        unsigned long long incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long v = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) wscan[w] = incl;
        __syncthreads();
        if (w == 0) {
            unsigned long long v = lane < BP2_TPB / 32 ? wscan[lane] : 0ull, iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long u = __shfl_up_sync(UB_FULL_MASK, iv, o); if (lane >= o) iv += u; }
            wscan[lane] = iv - v;
            if (lane == 31) wscan[32] = iv;
        }
        __syncthreads();
        const unsigned long long tot = wscan[32];
        n66 = int(tot & 0x1fffff); n36 = int((tot >> 21) & 0x1fffff); n33 = int((tot >> 42) & 0x1fffff);
        const bool fits = n_pair <= SP - 1 && 37 * n66 + 19 * n36 + 9 * n33 <= L.Pcap;   // row SP-1 is the neutral message
        if (!fits) {   // uniform: k_rot_bp solves this replica
            if (tid == 0) P.slow_list[atomicAdd(P.n_slow, 1)] = r;
            return;
        }
        unsigned long long run = wscan[w] + incl - s;
        const int cbase[3] = {0, n66, n66 + n36};
        for (int e = e0; e < e1; ++e) {
            int c = cls(e);
            int p = cbase[c] + int((run >> (21 * c)) & 0x1fffff);
            run += 1ull << (21 * c);
            int A = pair_ab[2 * e], B = pair_ab[2 * e + 1];
            bool sw = nrot[A] > nrot[B];
            pos[e] = (unsigned short)p;
            perm[p] = (unsigned short)e;
            fs[2 * p] = (unsigned short)((sw ? B : A) | (sw ? 0x8000 : 0));
            fs[2 * p + 1] = (unsigned short)(sw ? A : B);
        }
    }
    float* P66 = Pm;
    float* P36 = Pm + 37 * n66;
    float* P33 = P36 + 19 * n36;

    // ---- node energies -> probabilities (convert_energy_to_prob :239-256; single-state partners already folded) ----------
    for (int i = tid; i < nR * MAXR; i += BP2_TPB) { int A = i / MAXR, a = i % MAXR; bel[(A) * NODE_STRIDE + a] = P.enode[size_t(r) * nR * MAXR + i]; }
    __syncthreads();
    for (int A = tid; A < nR; A += BP2_TPB) {   // energy offset = smallest 1-body energy
        float m = bel[A * NODE_STRIDE];
        for (int a = 1; a < nrot[A]; ++a) m = fminf(m, bel[(A) * NODE_STRIDE + a]);
        offs[A] = m;
    }
    __syncthreads();
    for (int i = tid; i < P.n_bead; i += BP2_TPB) {
        float f = P.fold[size_t(r) * P.n_bead + i];
        if (f != 0.f) { float* b = &bel[P.bead_res[i] * NODE_STRIDE + P.bead_rot[i]]; if (P.multi_bead_states) atomicAdd(b, f); else *b += f; }
    }
    __syncthreads();
    for (int i = tid; i < 6 * nR; i += BP2_TPB) {
        int a = i / nR, A = i - a * nR;
        float pr = a < nrot[A] ? __expf(offs[A] - bel[(A) * NODE_STRIDE + a]) : 0.f;
        prob[(A) * NODE_STRIDE + a] = pr;
    }
    {   // pair energies -> probabilities: coalesced 16-byte reads of the replica's pair-major energy array
        const float4* src = reinterpret_cast<const float4*>(g_pmat);
        for (int i = tid; i < n_pair * 9; i += BP2_TPB) {
            float4 v = src[i];
            const int e = i / 9, k0 = (i - e * 9) * 4;
            const int A = pair_ab[2 * e], B = pair_ab[2 * e + 1], nA = nrot[A], nB = nrot[B];
            const int p = pos[e];
            const bool sw = nA > nB;
            const int nF = sw ? nB : nA, nS = sw ? nA : nB;
            float* blk;   // the pair's own block
            if (nF == 6) blk = P66 + 37 * p;
            else if (nS == 6) blk = P36 + 19 * (p - n66);
            else blk = P33 + 9 * (p - n66 - n36);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int k = k0 + u, a = k / 6, b = k - a * 6;
                if (a < nA && b < nB) {
                    int f = sw ? b : a, s2 = sw ? a : b;
                    blk[f * nS + s2] = __expf(-vv[u]);
                }
            }
        }
    }
    for (int t = tid; t < 2 * n_pair; t += BP2_TPB) {
        int cd = inc[t], p = pos[cd >> 1];
        inc2[t] = p * MSG_STRIDE + ((cd & 1) ^ (fs[2 * p] >> 15)) * 6;
    }
    for (int k = tid; k < 12; k += BP2_TPB) msg[(SP - 1) * MSG_STRIDE + k] = 1.f;
    // residues in the order the node update takes them: rank by (state count, degree) falling, ties by index, so that the
    // threads of a warp run the same unrolled body for about the same number of steps
    for (int A = tid; A < nR; A += BP2_TPB) {
        const int nA = nrot[A];
        if (nA < 2) continue;
        const int key = (nA << 16) | (istart[A + 1] - istart[A]);
        int rank = 0;
        for (int C = 0; C < nR; ++C) {
            const int nC = nrot[C];
            const int kc = (nC << 16) | (istart[C + 1] - istart[C]);
            rank += (nC >= 2) && (kc > key || (kc == key && C < A));
        }
        nlist[rank] = (unsigned short)A;
        atomicAdd(&n_multi_s, 1);
    }
    for (int p = tid; p < n_pair; p += BP2_TPB) {
        int nF = nrot[fs[2 * p] & 0x7fff], nS = nrot[fs[2 * p + 1]];
        for (int a = 0; a < 6; ++a) { msg[p * MSG_STRIDE + a] = a < nF ? 1.f : 0.f; msg[p * MSG_STRIDE + 6 + a] = a < nS ? 1.f : 0.f; }
    }
    __syncthreads();
    for (int i = tid; i < NODE_STRIDE * nRp; i += BP2_TPB) bel[i] = prob[i];
    __syncthreads();

    // ---- sweeps -------------------------------------------------------------------------------------------------------------
    auto messages = [&]() {
        for (int p = tid; p < n_pair; p += BP2_TPB) {
            const int F = fs[2 * p] & 0x7fff, S = fs[2 * p + 1];
            if (p < n66) bp2_pair<6, 6>(bel, nRp, F, S, msg + p * MSG_STRIDE, P66 + 37 * p);
            else if (p < n66 + n36) bp2_pair<3, 6>(bel, nRp, F, S, msg + p * MSG_STRIDE, P36 + 19 * (p - n66));
            else bp2_pair<3, 3>(bel, nRp, F, S, msg + p * MSG_STRIDE, P33 + 9 * (p - n66 - n36));
        }
    };
    const int n_multi = n_multi_s;
    const float damping = P.damping;
    const int dummy = (SP - 1) * MSG_STRIDE;   // row SP-1 holds ones: the neutral message
    (void)dummy;
#ifndef UB_BP2_NODE_THREAD
    // node update: belief = prob * prod(incoming messages), max-normalised and damped (rotamer.cpp:488-499,258-273).  FOUR lanes
    // per residue, each owning one or two states (lane s: states s and s+4): a lane multiplies its states' components of the
    // incident messages - every factor scaled by 6, so that a uniform message is neutral and neither 40 factors overflow nor
    // the product vanishes for every state at once (pure rescaling, as rotamer.cpp:492-493) - and the four lanes agree on the
    // maximum by two shuffles.  Against one thread per residue (three busy warps with 24 dependent loads per step) this
    // spreads the phase over ten warps and cuts its serial chain to one load and one multiply per incident pair.
    auto nodes = [&]() -> float {
        float dev = 0.f;
        const int sub = tid & 3;
        for (int i0 = (tid & ~31) >> 2; i0 < n_multi; i0 += BP2_TPB >> 2) {   // i0: first residue of this warp's eight
            const int i = i0 + ((tid & 31) >> 2);
            const bool act = i < n_multi;
            const int A = act ? nlist[i] : 0;
            const int nA = act ? nrot[A] : 0;
            const bool has0 = sub < nA, has1 = sub + 4 < nA;
            float b0 = has0 ? prob[A * NODE_STRIDE + sub] : 0.f, b1 = has1 ? prob[A * NODE_STRIDE + sub + 4] : 0.f;
            const int t1 = act ? istart[A + 1] : 0;
            for (int t = act ? istart[A] : 0; t < t1; ++t) {
                const int off = inc2[t];
                if (has0) b0 *= 6.f * msg[off + sub];
                if (has1) b1 *= 6.f * msg[off + sub + 4];
            }
            float mx = fmaxf(b0, b1);
            mx = fmaxf(mx, __shfl_xor_sync(UB_FULL_MASK, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(UB_FULL_MASK, mx, 2));
            const float sc = (damping != 0.f ? 1.f - damping : 1.f) * rcp_fast(mx);
            if (has0) {
                const float o = bel[A * NODE_STRIDE + sub], n = fmaf(sc, b0, damping * o);
                dev = fmaxf(dev, n - o);
                bel[A * NODE_STRIDE + sub] = n;
            }
            if (has1) {
                const float o = bel[A * NODE_STRIDE + sub + 4], n = fmaf(sc, b1, damping * o);
                dev = fmaxf(dev, n - o);
                bel[A * NODE_STRIDE + sub + 4] = n;
            }
        }
        return dev;
    };
#else
    auto nodes = [&]() -> float {
        float dev = 0.f;
        for (int i = tid; i < n_multi; i += BP2_TPB) {
            const int A = nlist[i];
            if (nrot[A] == 6) dev = fmaxf(dev, bp2_node<6>(A, prob, bel, nRp, istart[A], istart[A + 1], inc2, msg, dummy, damping));
            else dev = fmaxf(dev, bp2_node<3>(A, prob, bel, nRp, istart[A], istart[A + 1], inc2, msg, dummy, damping));
        }
        return dev;
    };
#endif
    // initial sweep: first messages from (prob, unit messages); node beliefs restart from prob/max (rotamer.cpp:1034)
    messages();
    __syncthreads();
    for (int A = tid; A < nR; A += BP2_TPB) {
        float mx = prob[A * NODE_STRIDE];
        for (int k = 1; k < MAXR; ++k) mx = fmaxf(mx, prob[(A) * NODE_STRIDE + k]);
        float imx = 1.f / mx;
        for (int k = 0; k < MAXR; ++k) bel[(A) * NODE_STRIDE + k] = prob[(A) * NODE_STRIDE + k] * imx;
    }
    __syncthreads();
    int iter = 0, unconverged = 1;
    for (; unconverged && iter < P.max_iter; iter += P.chunk) {
        float dev = 0.f;
        for (int j = 0; j < P.chunk; ++j) {
            messages();
            __syncthreads();
            dev = nodes();
            if (j + 1 < P.chunk) __syncthreads();
        }
        unconverged = __syncthreads_or(dev > P.tol);
    }
    if (tid == 0) { st[0] = iter; st[2] = !unconverged; st[3] = n66 | (n36 << 16); if (iter >= P.max_iter - P.chunk - 1) P.n_bad[r] += 1; }   // st[3]: class counts for the flop audit

    // ---- marginals (and Bethe free energy) ----------------------------------------------------------------------------------
    float en = 0.f;
    for (int A = tid; A < nR; A += BP2_TPB) {
        int nA = nrot[A];
        float b[MAXR], s = 0.f;
        for (int k = 0; k < MAXR; ++k) { b[k] = nA > 1 ? bel[(A) * NODE_STRIDE + k] : (k == 0 ? 1.f : 0.f); s += b[k]; }
        float is = 1.f / s;
        for (int k = 0; k < MAXR; ++k) { b[k] *= is; bel[(A) * NODE_STRIDE + k] = b[k]; node_marg[A * MAXR + k] = b[k]; }
        if (want_pot) {
            float e = offs[A];
            for (int k = 0; k < nA; ++k) e += b[k] * __logf((1e-10f + b[k]) / (1e-10f + prob[(A) * NODE_STRIDE + k]));
            en += e;
            if (fe_on) atomicAdd(&P.res_fe[size_t(r) * nR + A], e);
        }
    }
    __syncthreads();
    for (int p = tid; p < n_pair; p += BP2_TPB) {
        const int F = fs[2 * p] & 0x7fff, S = fs[2 * p + 1];
        float en_pair;
        if (p < n66) en_pair = bp2_pair_marginal<6, 6>(bel, nRp, F, S, msg + p * MSG_STRIDE, P66 + 37 * p, want_pot);
        else if (p < n66 + n36) en_pair = bp2_pair_marginal<3, 6>(bel, nRp, F, S, msg + p * MSG_STRIDE, P36 + 19 * (p - n66), want_pot);
        else en_pair = bp2_pair_marginal<3, 3>(bel, nRp, F, S, msg + p * MSG_STRIDE, P33 + 9 * (p - n66 - n36), want_pot);
        en += en_pair;
        if (fe_on) { atomicAdd(&P.res_fe[size_t(r) * nR + F], 0.5f * en_pair); atomicAdd(&P.res_fe[size_t(r) * nR + S], 0.5f * en_pair); }
    }
    __syncthreads();
    // backward-pass weight of every bead-pair entry, read from the marginals while they are in shared memory
    emit_entry_weights(P, r, BP2_TPB,
        [&](int cd) {
            const int slot = cd / 36, ab = cd - slot * 36, a = ab / 6, b = ab - a * 6;
            const int p = pos[slot];
            const bool sw = fs[2 * p] >> 15;
            const int f = sw ? b : a, s2 = sw ? a : b;
            if (p < n66) return P66[37 * p + f * 6 + s2];
            if (p < n66 + n36) return P36[19 * (p - n66) + f * 6 + s2];
            return P33[9 * (p - n66 - n36) + f * 3 + s2];
        },
        [&](int node) { const int A = node / MAXR; return bel[(A) * NODE_STRIDE + (node - A * MAXR)]; });
    if (want_pot) {
        float tot = block_sum(en, red);
        if (tid == 0) P.potential[r] = tot + P.e11[r];
    }
}

// (Measured and rejected in round 2: a kernel that gives every DIRECTED message its own lane - cavity, matrix-vector product,
// normalisation, then the product of a residue's incoming messages inside an aligned power-of-two lane group, beliefs and
// messages double-buffered, ONE barrier per sweep.  It removes k_rot_bp2's second barrier and its idle warps, but every lane
// repeats the cavity division and loads its own copy of the pair matrix (rows for one direction, columns for the other,
// so the two kinds of lane diverge): 175 k warp instructions per solve against 112 k, 1377 us against 650 us per batched
// evaluation at B = 4096, and 56 bytes of shared memory per lane left room for two CTAs per SM only.)

// parameter derivative (interaction_graph.h:404-415 with bead_interaction.h:204-207): sum over bead pairs (i<j, the
// reference's edge orientation) of the backward weight ss[e] (pair marginal / node marginal / 1) times
// d(quadspline)/d(param) of the ordered type pair; off the hot path, one thread per CSR row, atomics into the table
__global__ void k_rot_param_deriv(RotamerDev P, int r0, int r1, float* __restrict__ out) {
    const int r = r0 + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1 || i >= P.n_bead) return;
    const int* rowstart = P.rowstart + size_t(r) * (P.n_bead + 1);
    const unsigned short* dj = P.dj + size_t(r) * P.cap_e;
    const float* ss = P.ss + size_t(r) * P.cap_e;
    float x1[8];
    load8(elem_ptr(P.g.s1, r, i), x1);
    for (int e = rowstart[i]; e < rowstart[i + 1]; ++e) {
        const int j = dj[e];
        if (j <= i) continue;
        float x2[8], val[16];
        int idx[16];
        load8(elem_ptr(P.g.s1, r, j), x2);
        const int tp = P.g.s1.type[i] * P.g.n_type2 + P.g.s1.type[j];
        quadspline_param_deriv(P.g.param + size_t(tp) * P.g.n_param, P.q, x1, x2, idx, val);
        const float w = ss[e];
        for (int m = 0; m < 16; ++m) atomicAdd(out + size_t(tp) * P.g.n_param + idx[m], w * val[m]);
    }
}

struct RotamerSidechain : PotentialNode {
    std::vector<CoordNode*> prob_nodes;
    IGraphHost ig;
    int nka = 15, nk = 16;
    float knot_spacing = 0.5f;
    int n_res = 0, n_words = 0, max_pairs = 0, smem_pairs = 0, multi_bead_states = 0;
    std::vector<int> bead_res, bead_rot, res_nrot, res_key;
    DevBuf<int> d_bead_res, d_bead_rot, d_res_nrot, d_res_first, inc, istart, stats, code, rowstart;
    DevBuf<float> pmat, node_marg, enode, fold, e11, table, ss;
    DevBuf<unsigned short> pair_ab, dj, lower, order_e, order_d;
    int cap_e = 0, edge_tpb = EDGE_TPB, edge_split = 1, edge_G = 1;
    float damping, tol;
    int max_iter, chunk;
    size_t smem_prep = 0, smem_edge = 0, smem_bp = 0, smem_bp2 = 0, smem_build = 0;
    int bp2_occ = BP2_OCC;   // CTAs per SM the fast BP kernel was planned for (picks its block size)
    bool fast_build = false;   // k_rot_build instead of Verlet cache + k_refine + k_rot_prep
    bool build_large = false;  // ... with 1024-thread CTAs (large systems)
    BuildLay blay{0, 0, 0, 0, nullptr, nullptr};
    DevBuf<unsigned long long> build_spill;
    DevBuf<int> build_stats;
    Bp2Lay lay2{0, 0, 0};
    bool fast_bp = false;
    DevBuf<int> slow_list, n_slow, n_bad, fe_flag;
    DevBuf<float> res_fe;
    bool fe_always = false;   // a logger wants the per-residue free energies at every evaluation

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
        upload_table();
        damping = h5_attr<float>(g, ".", "damping");
        max_iter = h5_attr<int>(g, ".", "max_iter");
        tol = h5_attr<float>(g, ".", "tol");
        chunk = std::max(1, h5_attr<int>(g, ".", "iteration_chunk_size"));
        // residues = distinct (n_rot, k) in order of first appearance
        std::map<int, int> key_to_res;
        std::map<int, int> state_count;
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
            if (++state_count[it->second * 8 + rot] > 1) multi_bead_states = 1;
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
        // fast build: one bead per (residue, state), the beads of a residue contiguous with ascending state (ff_1)
        std::vector<int> res_first(n_res, -1);
        fast_build = !multi_bead_states && !getenv("UPSIDE_B200_NO_FAST_BUILD");
        for (int b = 0; b < ig.n1 && fast_build; ++b) {
            if (bead_rot[b] == 0) res_first[bead_res[b]] = b;
            if (res_first[bead_res[b]] < 0 || b - res_first[bead_res[b]] != bead_rot[b]) fast_build = false;
        }
        for (int A = 0; A < n_res && fast_build; ++A)
            if (res_first[A] < 0 || (A + 1 < n_res ? res_first[A + 1] : ig.n1) != res_first[A] + res_nrot[A]) fast_build = false;
        d_res_first.upload(res_first);
    }
    // symmetric tables must satisfy p(t1,t2).ang1 == p(t2,t1).ang2 and equal radial parts (bead_interaction.h:209-218);
    // that lets the device keep only the rows t1<=t2
    void upload_table() {
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
        std::vector<float> t;
        for (int a = 0; a < n; ++a)
            for (int b = a; b < n; ++b) t.insert(t.end(), &ig.h_param[(size_t(a) * n + b) * np], &ig.h_param[(size_t(a) * n + b) * np] + np);
        table.upload(t);
    }
    void finalize() override {
        size_t B = engine->n_rep;
        int device_smem = 0;
        UB_CUDA(cudaDeviceGetAttribute(&device_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, engine->device));
        if (fast_build) {   // capacities per residue: sphere-test survivors (A<B) and active pairs; an overflow is reported
            double scale = 1.0;
            if (const char* s = getenv("UPSIDE_B200_NEIGHBOR_SCALE")) scale = std::max(0.05, atof(s));
            const long all_pairs = long(n_res) * (n_res - 1) / 2;
            // shared-memory capacities cover a relaxed chain (measured at 100 residues, T = 0.8: see DESIGN.md); the totals
            // cover clashing start structures
            double cs = 12, as = 6;
            if (const char* e = getenv("UPSIDE_B200_BUILD_CAPS")) sscanf(e, "%lf,%lf", &cs, &as);
            auto cap = [&](double per_res) { return (int)((std::max<long>(4, std::min<long>(all_pairs, (long)std::ceil(per_res * scale * n_res))) + 3) & ~3L); };
            blay.capa = cap(as); blay.capc = std::max(cap(cs), 2 * blay.capa);   // (the offsets of pass 2 reuse the mask + candidate arrays)
            blay.capc_tot = std::max(blay.capc, cap(64)); blay.capa_tot = std::max(blay.capa, cap(32));
            smem_build = build_bytes();
            // large systems keep the Verlet path
            if (n_res > BUILD_MAX_RES || max_pairs > BUILD_MAX_SLOT || smem_build > (size_t)device_smem) fast_build = false;
        }
        ig.lists = !fast_build;
        if (fast_build) {
            build_spill.alloc(B * build_spill_words(blay) + 1);
            build_stats.alloc(B * 2);
            blay.spill = build_spill.p; blay.bstats = build_stats.p;
        }
        ig.allocate(engine);
        // BP kernel: fixed part + 204 bytes per pair; aim for two resident CTAs per SM, never more pairs than can occur
        size_t fixed_bp = sizeof(float) * (size_t(n_res) * MAXR * 2 + n_res + 32) + sizeof(int) * (2 * n_res + 1);
        size_t per_pair = 12 * 4 + 36 * 4 + 2 * 4 + 2 * 2;
        size_t budget = std::min<size_t>(device_smem, 110 * 1024);
        smem_pairs = budget > fixed_bp ? (int)std::min<size_t>(max_pairs, (budget - fixed_bp) / per_pair) : 0;
        smem_bp = fixed_bp + size_t(smem_pairs) * per_pair + 16;
        smem_prep = sizeof(unsigned) * size_t(n_res) * n_words + sizeof(int) * (4 * n_res + 2) + sizeof(float) * n_res * MAXR +
                    sizeof(int) * (2 * size_t(ig.n1) + 1 + 33 + 256) + sizeof(unsigned short) * (size_t(n_res) * n_words + 2 * size_t(ig.n1)) + 16;
        // CSR capacity: 64 partners per bead on average (a collapsed 100-residue globule has ~50 within 7 A), never more than
        // the ELL rows can hold; an overflow raises the engine's error flag
        {
            double scale = 1.0;
            if (const char* sc = getenv("UPSIDE_B200_NEIGHBOR_SCALE")) scale = std::max(0.05, atof(sc));
            long want = (long)std::ceil(64 * scale * ig.n1);
            cap_e = (int)std::max<long>(16, std::min<long>(long(ig.n1) * ig.K1, want));
            cap_e = (cap_e + 7) & ~7;   // keeps the per-replica arrays 16-byte aligned
        }
        // one thread per bead row: the smallest block (whole warps, <= 256 threads) that covers the rows in equal passes;
        // with a small batch the rows are split over several CTAs instead
        if (engine->n_rep >= 148) {
            int passes = (ig.n1 + EDGE_TPB - 1) / EDGE_TPB;
            edge_tpb = std::min(EDGE_TPB, (((ig.n1 + passes - 1) / passes) + 31) & ~31);
            edge_split = 1;
        } else {
            edge_tpb = 64;
            edge_split = std::max(1, std::min((ig.n1 + 63) / 64, 600 / std::max(1, engine->n_rep)));
        }
        // replicas per CTA (they share one staged copy of the 52 KB table): as many as the thread bound of the build allows
        edge_G = edge_split == 1 ? std::max(1, EDGE_MAXT / edge_tpb) : 1;
        if (const char* e = getenv("UPSIDE_B200_EDGE_G")) edge_G = std::max(1, std::min(atoi(e), EDGE_MAXT / edge_tpb));
        auto edge_bytes = [&](int G) {
            return sizeof(BeadRec) * size_t(ig.n1) * G + sizeof(float) * (table.n + 4) + sizeof(int) * size_t(ig.n_type1) * ig.n_type1 + 64;
        };
        // sharing pays while two CTAs still fit an SM; a large system falls back to one replica per CTA
        while (edge_G > 1 && edge_bytes(edge_G) > std::min<size_t>(device_smem, 112 * 1024)) --edge_G;
        smem_edge = edge_bytes(edge_G);
        if (ig.n_type1 > 255) throw std::string("rotamer node: more than 255 bead types");
        if (smem_prep > (size_t)device_smem || smem_edge > (size_t)device_smem || fixed_bp > (size_t)device_smem)
            throw std::string("rotamer node: system too large for the shared-memory kernels");
        UB_CUDA(cudaFuncSetAttribute(k_rot_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
        // block size of the build kernel: many warps per CTA once fewer than three CTAs fit an SM
        build_large = fast_build && smem_build * 3 > (size_t)device_smem && !getenv("UPSIDE_B200_BUILD_SMALL_BLOCKS");
        if (fast_build && build_large) UB_CUDA(cudaFuncSetAttribute(k_rot_build<BUILD_TPB_LARGE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_build));
        else if (fast_build) UB_CUDA(cudaFuncSetAttribute(k_rot_build<BUILD_TPB_SMALL, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_build));
        UB_CUDA(cudaFuncSetAttribute(k_rot_energy<15, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_edge));
        UB_CUDA(cudaFuncSetAttribute(k_rot_deriv<15, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_edge));
        UB_CUDA(cudaFuncSetAttribute(k_rot_energy<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_edge));
        UB_CUDA(cudaFuncSetAttribute(k_rot_deriv<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_edge));
        UB_CUDA(cudaFuncSetAttribute(k_rot_bp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bp));
        plan_fast_bp(device_smem);
        slow_list.alloc(B);
        n_bad.upload(std::vector<int>(B, 0));
        n_slow.alloc(1);
        fe_flag.upload(std::vector<int>(1, 0));
        res_fe.alloc(B * n_res);
        pmat.alloc(B * max_pairs * 36);
        pair_ab.alloc(B * max_pairs * 2);
        inc.alloc(B * 2 * max_pairs);
        istart.alloc(B * (n_res + 1));
        node_marg.alloc(B * n_res * MAXR + B * size_t(max_pairs) * 12);   // + message spill area for oversized replicas
        enode.alloc(B * n_res * MAXR);
        fold.alloc(B * ig.n1);
        e11.alloc(B);
        stats.alloc(B * 4);
        code.alloc(B * size_t(cap_e));
        dj.alloc(B * size_t(cap_e));
        ss.alloc(B * size_t(cap_e));
        rowstart.alloc(B * (ig.n1 + 1));
        lower.alloc(B * ig.n1);
        order_e.alloc(B * ig.n1);
        order_d.alloc(B * ig.n1);
    }
    // Shared-memory plan of k_rot_bp2: the most resident CTAs per SM (at most 3) whose pair capacity still covers a typical
    // replica (3.8 residue pairs per residue; measured 2.6-3.3 on coil-like chains) with 27 floats of
    // pair matrix per pair (class mix of a uniform sequence: 21.5).  Larger replicas take the general kernel.
    static size_t bp2_bytes(int nR, const Bp2Lay& L) {
        size_t words = size_t(2 * NODE_STRIDE) * L.nRp + nR + 32 + (2 * nR + 1) + ((nR + 1) & 1) + 2 * 34 + size_t(MSG_STRIDE) * L.SP + L.Pcap;
        return (words + 2 * size_t(L.SP)) * 4 + (size_t(4) * L.SP + nR) * 2 + 16;
    }
    void plan_fast_bp(int device_smem) {
        fast_bp = false;
        for (int n : res_nrot) if (n != 1 && n != 3 && n != 6) return;   // the class-specialised kernel knows 3 and 6 states
        if (getenv("UPSIDE_B200_NO_FAST_BP")) return;
        int sm_total = 0;
        UB_CUDA(cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, engine->device));
        auto pad4 = [](int v) { return ((v + 27) / 32) * 32 + 4; };   // smallest value >= v that is 4 mod 32
        const int want_pairs = std::min<long>(max_pairs, std::max<long>(32, (long)std::ceil(0.1 * UB_BP2_WANT10 * n_res)));
        for (int occ = BP2_OCC; occ >= 1; --occ) {
            size_t budget = std::min<size_t>(device_smem, size_t(sm_total) / occ - 1024);
            Bp2Lay L;
            L.nRp = n_res;
            L.SP = pad4(4); L.Pcap = 0;
            size_t fixed = bp2_bytes(n_res, L);
            if (fixed >= budget) continue;
            int sp = (int)std::min<size_t>(max_pairs + 1, (budget - fixed) / (MSG_STRIDE * 4 + 27 * 4 + 16));
            sp = std::min(sp, 32767);
            if (sp < want_pairs + 1 && occ > 1) continue;
            if (sp < 8) continue;
            L.SP = sp;   // (pair-major rows with odd strides: no alignment constraint on the capacity)
            if (L.SP < 4) continue;
            L.Pcap = 27 * L.SP;
            lay2 = L;
            smem_bp2 = bp2_bytes(n_res, L);
            bp2_occ = occ;
            if (const char* e = getenv("UPSIDE_B200_BP2_SMALL_BLOCKS")) if (atoi(e)) bp2_occ = BP2_OCC;   // 384 threads whatever the occupancy
            if (bp2_occ >= BP2_OCC) UB_CUDA(cudaFuncSetAttribute(k_rot_bp2<BP2_TPB, BP2_OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bp2));
            else if (bp2_occ == 2) UB_CUDA(cudaFuncSetAttribute(k_rot_bp2<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bp2));
            else UB_CUDA(cudaFuncSetAttribute(k_rot_bp2<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bp2));
            fast_bp = true;
            return;
        }
    }
    size_t build_bytes() const {
        const size_t nR = n_res, nW = n_words, nb = ig.n1;
        return 16 * size_t(blay.capa) + 8 * size_t(blay.capc) + 16 * (nb + nR) + 4 * (2 * nR * nW + blay.capc) + 4 * (6 * nR + 3) +
               4 * nR * MAXR + 4 * (nb + 1 + 33 + 256) + 2 * (2 * nR * nW + 2 * nb + 2 * size_t(blay.capa)) + 16;
    }
    RotamerDev dev() {
        RotamerDev P;
        P.g = ig.dev();
        P.q.nka = nka; P.q.nk = nk; P.q.inv_dx = 1.f / knot_spacing; P.q.inv_dtheta = (nka - 3) / 2.f;
        P.n_bead = ig.n1; P.n_res = n_res; P.n_words = n_words; P.n_type = ig.n_type1;
        P.bead_res = d_bead_res.p; P.bead_rot = d_bead_rot.p; P.res_nrot = d_res_nrot.p; P.res_first = d_res_first.p;
        P.loc_identity = 1;
        for (int b = 0; b < ig.n1; ++b) if (ig.loc1[b] != b) P.loc_identity = 0;
        P.table = table.p;
        P.n_prob = (int)prob_nodes.size();
        for (int i = 0; i < P.n_prob; ++i) {
            P.prob_out[i] = prob_nodes[i]->output; P.prob_sens[i] = prob_nodes[i]->sens;
            P.prob_wp[i] = prob_nodes[i]->wp; P.prob_n[i] = prob_nodes[i]->n_elem;
        }
        P.damping = damping; P.tol = tol; P.max_iter = max_iter; P.chunk = chunk; P.max_pairs = max_pairs;
        P.smem_pairs = smem_pairs; P.multi_bead_states = multi_bead_states;
        P.cap_e = cap_e;
        P.code = code.p; P.dj = dj.p; P.ss = ss.p; P.rowstart = rowstart.p; P.lower = lower.p; P.order_e = order_e.p; P.order_d = order_d.p;
        P.enode = enode.p; P.fold = fold.p; P.e11 = e11.p;
        P.pmat = pmat.p; P.pair_ab = pair_ab.p; P.inc = inc.p; P.istart = istart.p; P.node_marg = node_marg.p; P.stats = stats.p;
        P.potential = potential; P.error_flag = engine->error_flag.p;
        P.res_fe = res_fe.p; P.fe_flag = fe_flag.p;
        P.slow_list = slow_list.p; P.n_slow = n_slow.p; P.n_bad = n_bad.p; P.n_rep = engine->n_rep;
        return P;
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!ig.n1) return;
        RotamerDev P = dev();
        int want = mode == PotentialAndDerivMode;
        // edge kernels: persistent CTAs (as many as stay resident) striding over the replicas, edge_G replicas at a time
        int persist = std::min((engine->n_rep + edge_G - 1) / edge_G, 148 * EDGE_OCC);
        if (fast_build) {
            if (build_large) k_rot_build<BUILD_TPB_LARGE, 1><<<engine->n_rep, BUILD_TPB_LARGE, smem_build, s>>>(P, blay);
            else k_rot_build<BUILD_TPB_SMALL, 4><<<engine->n_rep, BUILD_TPB_SMALL, smem_build, s>>>(P, blay);
            engine->mark(s, "rotamer/build");
        } else {
            ig.build(s);
            engine->mark(s, "rotamer/pairlist");
            k_rot_prep<<<engine->n_rep, PREP_TPB, smem_prep, s>>>(P);
            engine->mark(s, "rotamer/prep");
        }
        const bool ff1_knots = nka == 15 && nk == 16;   // the PARAM_7A_CUTOFF build of the reference (bead_interaction.h:12-27)
        if (ff1_knots) k_rot_energy<15, 16><<<dim3(edge_split, persist), edge_tpb * edge_G, smem_edge, s>>>(P, want, engine->n_rep, edge_tpb);
        else k_rot_energy<0, 0><<<dim3(edge_split, persist), edge_tpb * edge_G, smem_edge, s>>>(P, want, engine->n_rep, edge_tpb);
        engine->mark(s, "rotamer/energy");
        if (fast_bp) {
            if (bp2_occ >= BP2_OCC) k_rot_bp2<BP2_TPB, BP2_OCC><<<engine->n_rep, BP2_TPB, smem_bp2, s>>>(P, lay2, want);
            else if (bp2_occ == 2) k_rot_bp2<512, 2><<<engine->n_rep, 512, smem_bp2, s>>>(P, lay2, want);
            else k_rot_bp2<1024, 1><<<engine->n_rep, 1024, smem_bp2, s>>>(P, lay2, want);
            k_rot_bp<<<std::min(engine->n_rep, 148), BP_TPB, smem_bp, s>>>(P, want, 1);   // replicas the fast path declined
        } else {
            k_rot_bp<<<engine->n_rep, BP_TPB, smem_bp, s>>>(P, want, 0);
        }
        engine->mark(s, "rotamer/bp");
        if (ff1_knots) k_rot_deriv<15, 16><<<dim3(edge_split, persist), edge_tpb * edge_G, smem_edge, s>>>(P, engine->n_rep, edge_tpb);
        else k_rot_deriv<0, 0><<<dim3(edge_split, persist), edge_tpb * edge_G, smem_edge, s>>>(P, engine->n_rep, edge_tpb);
        engine->mark(s, "rotamer/deriv");
    }
    // the bead pair list in the reference's emission order; the fast build keeps no ELL table, the CSR rows hold every pair
    // in both directions
    bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) override {
        if (!fast_build) return ig.pairlist(replica, i1, i2);
        if (replica < 0 || replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        std::vector<int> rs(ig.n1 + 1);
        UB_CUDA(cudaMemcpy(rs.data(), rowstart.p + size_t(replica) * (ig.n1 + 1), rs.size() * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<unsigned short> partner(rs[ig.n1]);
        if (!partner.empty())
            UB_CUDA(cudaMemcpy(partner.data(), dj.p + size_t(replica) * cap_e, partner.size() * sizeof(unsigned short), cudaMemcpyDeviceToHost));
        i1.clear();
        i2.clear();
        for (int i = 0; i < ig.n1; ++i)
            for (int e = rs[i]; e < rs[i + 1]; ++e)
                if (i < (int)partner[e]) { i1.push_back(i); i2.push_back(partner[e]); }
        sort_reference_order(i1, i2);
        return true;
    }
    std::vector<float> count_edges_by_type(int replica) {
        std::vector<int> i1, i2;
        get_pairlist(replica, i1, i2);
        std::vector<float> ret(size_t(ig.n_type1) * ig.n_type2, 0.f);
        for (size_t e = 0; e < i1.size(); ++e) ret[ig.type1[i1[e]] * ig.n_type2 + ig.type2[i2[e]]] += 1.f;
        return ret;
    }
    std::vector<float> get_param() const override { return ig.h_param; }
    void set_param(const std::vector<float>& p) override { ig.set_param(p); upload_table(); }
    std::vector<float> get_param_deriv(int replica) override {
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        DevBuf<float> acc;
        acc.upload(std::vector<float>(ig.h_param.size(), 0.f));
        const int first = replica < 0 ? 0 : replica, last = replica < 0 ? engine->n_rep : replica + 1;
        for (int r0 = first; ig.n1 && r0 < last; r0 += 32768) {
            const int r1 = std::min(last, r0 + 32768);
            k_rot_param_deriv<<<dim3((ig.n1 + 127) / 128, r1 - r0), 128>>>(dev(), r0, r1, acc.p);
        }
        UB_CUDA(cudaDeviceSynchronize());
        return acc.download();
    }

    void set_fe_flag(int v) { UB_CUDA(cudaMemcpy(fe_flag.p, &v, sizeof(int), cudaMemcpyHostToDevice)); }
    // rotamer.cpp:657-672: cumulative bad solves, free energy per residue, marginal-weighted 1-body energy per residue
    // and probability node.  The free energies are accumulated by the kernels themselves from now on (fe_flag).
    void add_loggers(int level, std::vector<NodeLogger>& out) override {
        if (level < 1) return;
        fe_always = true;
        set_fe_flag(1);
        out.push_back({"rotamer_bad_solves_cumulative", {1}, true, [this](int r) { return get_value_by_name(r, "read n_bad_solve"); }});
        out.push_back({"rotamer_free_energy", {(uint64_t)n_res}, false, [this](int r) { return get_value_by_name(r, "rotamer_free_energy"); }});
        const int n_prob = (int)prob_nodes.size();
        for (int p = 0; p < n_prob; ++p)
            out.push_back({"rotamer_1body_energy" + std::to_string(p), {(uint64_t)n_res}, false, [this, p, n_prob](int r) {
                auto all = get_value_by_name(r, "rotamer_1body_energy");
                std::vector<float> v(n_res);
                for (int i = 0; i < n_res; ++i) v[i] = all[size_t(i) * n_prob + p];
                return v;
            }});
    }
    std::vector<float> get_value_by_name(int replica, const char* log_name) override {
        std::string nm(log_name);
        engine->sync_and_check();
        if (nm == "count_edges_by_type") return count_edges_by_type(replica);
        if (nm == "n_node") return {float(n_res)};
        if (nm == "bead_marginal") {   // B200 extension: node marginal of each bead's (residue, rotamer)
            std::vector<float> nmg(size_t(n_res) * MAXR);
            UB_CUDA(cudaMemcpy(nmg.data(), node_marg.p + size_t(replica) * n_res * MAXR, nmg.size() * sizeof(float), cudaMemcpyDeviceToHost));
            std::vector<float> out(ig.n1);
            for (int b = 0; b < ig.n1; ++b) out[b] = nmg[bead_res[b] * MAXR + bead_rot[b]];
            return out;
        }
        if (nm == "build_stats") {     // B200 extension: (sphere-test survivors, active residue pairs) of the fast build
            std::vector<int> st(2, 0);
            if (fast_build) UB_CUDA(cudaMemcpy(st.data(), build_stats.p + size_t(replica) * 2, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            return {float(st[0]), float(st[1])};
        }
        if (nm == "solve_stats") {     // B200 extension: (n_iter, n residue pairs, converged, n 6x6 pairs, n 3x6 pairs)
            std::vector<int> st(4);
            UB_CUDA(cudaMemcpy(st.data(), stats.p + size_t(replica) * 4, 4 * sizeof(int), cudaMemcpyDeviceToHost));
            return {float(st[0]), float(st[1]), float(st[2]), float(st[3] & 0xffff), float((st[3] >> 16) & 0xffff)};
        }
        if (nm == "read n_bad_solve" || nm == "read n_bad_solve and reset") {   // rotamer.cpp:764-770
            int v = 0;
            UB_CUDA(cudaMemcpy(&v, n_bad.p + replica, sizeof(int), cudaMemcpyDeviceToHost));
            if (nm != "read n_bad_solve") UB_CUDA(cudaMemset(n_bad.p + replica, 0, sizeof(int)));
            return {float(v)};
        }
        // residues in the reference's reporting order: grouped by state count (1, 3, 6), inside a group by the index the
        // bead ids carry (nodes1/nodes3/nodes6, rotamer.cpp:695-707); arrange_energies (:929-955) lists them in order of
        // first appearance instead
        auto download = [&](const float* p, size_t n) {
            std::vector<float> v(n);
            UB_CUDA(cudaMemcpy(v.data(), p, n * sizeof(float), cudaMemcpyDeviceToHost));
            return v;
        };
        if (nm == "rotamer_free_energy") {   // residue_free_energies (:868-902): filled by the kernels while fe_flag is set
            if (!fe_always) {
                set_fe_flag(1);
                engine->compute(PotentialAndDerivMode);
                engine->sync_and_check();
                set_fe_flag(0);
            }
            return download(res_fe.p + size_t(replica) * n_res, n_res);
        }
        if (nm == "rotamer_1body_energy") {   // marginal-weighted 1-body energy per residue and probability node (:680-691,904-927)
            std::vector<float> nmg = download(node_marg.p + size_t(replica) * n_res * MAXR, size_t(n_res) * MAXR);
            const int n_prob = (int)prob_nodes.size();
            std::vector<float> out(size_t(n_res) * n_prob, 0.f);
            for (int p = 0; p < n_prob; ++p) {
                CoordNode* pn = prob_nodes[p];
                std::vector<float> o = download(pn->output + size_t(replica) * pn->stride(), pn->stride());
                for (int b = 0; b < ig.n1; ++b)
                    out[size_t(bead_res[b]) * n_prob + p] += nmg[bead_res[b] * MAXR + bead_rot[b]] * o[size_t(ig.loc1[b]) * pn->wp];
            }
            return out;
        }
        if (nm == "node_energy") {   // -log(prob) per (node, state), 1e5 for absent states (:698-711)
            std::vector<float> en = download(enode.p + size_t(replica) * n_res * MAXR, size_t(n_res) * MAXR);
            std::vector<float> fo = download(fold.p + size_t(replica) * ig.n1, ig.n1);
            std::vector<float> tot(en);
            std::vector<float> offs(n_res, 0.f);
            for (int A = 0; A < n_res; ++A) {
                float m = en[A * MAXR];
                for (int a = 1; a < res_nrot[A]; ++a) m = std::min(m, en[A * MAXR + a]);
                offs[A] = m;
            }
            for (int b = 0; b < ig.n1; ++b) tot[bead_res[b] * MAXR + bead_rot[b]] += fo[b];
            std::vector<int> order;
            for (int n_rot : {1, 3, 6}) {
                std::vector<std::pair<int, int>> grp;
                for (int A = 0; A < n_res; ++A) if (res_nrot[A] == n_rot) grp.push_back({res_key[A] >> 4, A});
                std::sort(grp.begin(), grp.end());
                for (auto& g : grp) order.push_back(g.second);
            }
            std::vector<float> out(order.size() * 6, 1e5f);
            for (size_t nn = 0; nn < order.size(); ++nn) {
                int A = order[nn];
                for (int a = 0; a < res_nrot[A]; ++a) out[nn * 6 + a] = tot[A * MAXR + a] - offs[A];
            }
            return out;
        }
        throw std::string("Value ") + log_name + " not implemented";
    }
};
RegisterNodeType<RotamerSidechain, -1> rotamer_node("rotamer");

}  // namespace
}  // namespace ub
