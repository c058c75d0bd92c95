// Edge-centric evaluation over ELL neighbour rows, inside one CTA that owns one replica.
//
// The ELL rows of a sparse interaction graph are short and uneven (hbond_coverage: 1.4 partners per bead on average), so
// lane groups walking rows leave most lanes idle.  Here the CTA numbers the edges of a run of rows by an exclusive scan of
// the row lengths (rows ascending, partners ascending inside a row), then runs ONE THREAD PER EDGE over the pair term - the
// thread finds its row by bisection of the scan - parks the per-edge results in shared memory (component-major, conflict
// free) and finally lets one thread per row sum its own contiguous segment.  Every sum has a fixed order => results are
// bit-reproducible; nothing is accumulated with atomics.
#pragma once
#include "common.cuh"

namespace ub {

#ifndef UB_EL_CAP
#define UB_EL_CAP 1024
#endif
constexpr int EL_CAP = UB_EL_CAP;   // edges evaluated per chunk (a chunk is a run of whole rows)

struct EdgeScratch {
    int* start;        // [n_rows_max + 1] exclusive scan of the row lengths
    float* vals;       // [NV_MAX][EL_CAP]
    int* wtot;         // [33]
};
inline size_t edge_scratch_bytes(int n_rows_max, int nv_max) {
    return sizeof(int) * (size_t(n_rows_max) + 1 + 33) + sizeof(float) * size_t(nv_max) * EL_CAP;
}
__device__ __forceinline__ EdgeScratch carve_edge_scratch(void* p, int n_rows_max, int nv_max) {
    EdgeScratch s;
    s.vals = reinterpret_cast<float*>(p);
    s.start = reinterpret_cast<int*>(s.vals + size_t(nv_max) * EL_CAP);
    s.wtot = s.start + n_rows_max + 1;
    return s;
}

// start[0..n] = exclusive scan of len(row); every thread of the CTA must call (contains barriers)
template <typename LenF>
__device__ __forceinline__ void scan_row_lengths(int n, LenF len, int* start, int* wtot) {
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int per = (n + T - 1) / T;
    const int r0 = min(n, t * per), r1 = min(n, r0 + per);
    int s = 0;
    for (int r = r0; r < r1; ++r) s += len(r);
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int nw = (T + 31) >> 5;
        int v = lane < nw ? wtot[lane] : 0, iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(UB_FULL_MASK, iv, o); if (lane >= o) iv += u; }
        wtot[lane] = iv - v;
        if (lane == 31) wtot[32] = iv;
    }
    __syncthreads();
    int run = wtot[w] + incl - s;
    for (int r = r0; r < r1; ++r) { start[r] = run; run += len(r); }
    if (t == 0) start[n] = wtot[32];
    __syncthreads();
}

// the same scan, in place: start[0..n) holds the row lengths on entry (written by any threads, visible after a barrier
// the caller has already passed); no global memory is touched
__device__ __forceinline__ void scan_row_lengths_inplace(int n, int* start, int* wtot) {
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int per = (n + T - 1) / T;
    const int r0 = min(n, t * per), r1 = min(n, r0 + per);
    int s = 0;
    for (int r = r0; r < r1; ++r) s += start[r];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int nw = (T + 31) >> 5;
        int v = lane < nw ? wtot[lane] : 0, iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(UB_FULL_MASK, iv, o); if (lane >= o) iv += u; }
        wtot[lane] = iv - v;
        if (lane == 31) wtot[32] = iv;
    }
    __syncthreads();
    int run = wtot[w] + incl - s;
    for (int r = r0; r < r1; ++r) { const int len = start[r]; start[r] = run; run += len; }
    if (t == 0) start[n] = wtot[32];
    __syncthreads();
}

// Evaluate `edge(row, partner, out[NV])` once per edge of rows [0,n) and hand every row the sum over its edges:
// `row_done(row, n_edges_of_row, sum[NV])` is called by exactly one thread per row (also for empty rows).
// nbr/K: the replica's ELL table for these rows; start: scan_row_lengths of the lengths to use (a caller may zero rows).
template <int NV, typename EdgeF, typename RowF>
__device__ __forceinline__ void for_each_edge(int n, const unsigned short* __restrict__ nbr, int K, const EdgeScratch& S,
                                              EdgeF edge, RowF row_done) {
    const int T = blockDim.x, t = threadIdx.x;
    int ra = 0;
    while (ra < n) {
        // rb = largest row index with start[rb] - start[ra] <= EL_CAP (at least ra+1: a row never exceeds K <= EL_CAP)
        const int base = S.start[ra];
        int lo = ra + 1, hi = n;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (S.start[mid] - base <= EL_CAP) lo = mid; else hi = mid - 1;
        }
        const int rb = lo, ne = S.start[rb] - base;
        // one thread per edge: find the row by bisection of the scan (shared memory), fetch the partner (consecutive threads
        // read consecutive entries of a row), evaluate
        for (int e = t; e < ne; e += T) {
            const int ge = base + e;
            int a = ra, b = rb - 1;
            while (a < b) {
                int mid = (a + b + 1) >> 1;
                if (S.start[mid] <= ge) a = mid; else b = mid - 1;
            }
            const int partner = nbr[size_t(a) * K + (ge - S.start[a])];
            float out[NV];
            edge(a, partner, out);
#pragma unroll
            for (int c = 0; c < NV; ++c) S.vals[c * EL_CAP + e] = out[c];
        }
        __syncthreads();
        for (int row = ra + t; row < rb; row += T) {
            const int s = S.start[row] - base, c = S.start[row + 1] - S.start[row];
            float sum[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) sum[q] = 0.f;
#pragma unroll 4
            for (int k = 0; k < c; ++k)
#pragma unroll
                for (int q = 0; q < NV; ++q) sum[q] += S.vals[q * EL_CAP + s + k];
            row_done(row, c, sum);
        }
        __syncthreads();
        ra = rb;
    }
}

}  // namespace ub
