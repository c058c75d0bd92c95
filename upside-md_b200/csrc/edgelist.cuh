// Helpers for kernels that walk ELL neighbour rows: a CTA-wide exclusive scan of row lengths (rotamer prep) and the
// warp-cooperative edge walk of the sparse pair-term kernels.
#pragma once
#include "common.cuh"

namespace ub {

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

// ---- warp-cooperative edge walk ------------------------------------------------------------------------------------------
// Thread-per-row kernels over sparse rows leave most lanes idle (hbond_coverage: two rows in three are empty, the others
// hold 1-4 partners).  WarpEdges deals the edges of a warp's 32 rows round-robin to its lanes instead: lane l of round `it`
// evaluates edge number 32*it + l in (row, position) order, and the per-row sums come back to the row's owner through a
// segmented shuffle reduction in a fixed lane order (bit-reproducible, no atomics).  No shared memory, no barriers; all 32
// lanes of the warp must take part in every call.
struct WarpEdges {
    int lane, c, excl, total;
    __device__ explicit WarpEdges(int count) : lane(threadIdx.x & 31), c(count) {
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(UB_FULL_MASK, incl, o); if (lane >= o) incl += v; }
        excl = incl - c;
        total = __shfl_sync(UB_FULL_MASK, incl, 31);
    }
    __device__ int rounds() const { return (total + 31) >> 5; }
    // the edge this lane evaluates in round `it`: owner lane `s` (the last lane whose rows start at or before the edge: empty
    // rows share their start with the next non-empty one) and position `k` inside the owner's row; false = padding
    __device__ bool edge(int it, int& s, int& k) const {
        const int e = 32 * it + lane;
        s = 0;
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int cand = s + step;
            const int ex = __shfl_sync(UB_FULL_MASK, excl, cand & 31);
            if (cand < 32 && ex <= e) s = cand;
        }
        k = e - __shfl_sync(UB_FULL_MASK, excl, s);
        return e < total;
    }
    // sum of `v` over the lanes of round `it` that carry edges of the CALLER's row (0 if none); s/valid as returned by edge()
    __device__ float row_sum(int it, float v, int s, bool valid) const {
        const int key = valid ? s : -1 - lane;
        if (!valid) v = 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {   // segmented suffix sums: equal keys are contiguous
            const float w = __shfl_down_sync(UB_FULL_MASK, v, o);
            const int ko = __shfl_down_sync(UB_FULL_MASK, key, o);
            if (lane + o < 32 && ko == key) v += w;
        }
        const int lo = max(excl, 32 * it), hi = min(excl + c, 32 * it + 32);
        const float head = __shfl_sync(UB_FULL_MASK, v, (lo - 32 * it) & 31);
        return lo < hi ? head : 0.f;
    }
};

}  // namespace ub
