// Measured denominators for the roofline bench.py reports ("measure with an FMA microbenchmark on
// the box - MEASURED_PEAKS.json only has HBM and bf16").  The path is FP32-issue bound, so the number that matters is
// what the FP32 FMA pipe sustains on THIS GPU at the clocks it actually runs: a persistent grid (one CTA of 1024 threads
// per SM slot) of independent FFMA chains, timed with CUDA events.
#include <cuda_runtime.h>

#include <string>

#include "../../include/upside_b200.h"
#include "engine.h"

namespace {

constexpr int CHAINS = 8;       // independent accumulators per thread: hides the 4-cycle FFMA latency twice over
constexpr int INNER = 4096;     // FFMAs per chain per launch

__global__ void __launch_bounds__(1024) k_fma_peak(float* out, float a, float b) {
    float acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = float(threadIdx.x + c);
    for (int i = 0; i < INNER; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fmaf(acc[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 12345.678f) out[0] = s;   // never true for the operands used; keeps the chains alive
}

}  // namespace

extern "C" int ub_measure_fp32_peak(int device, float* tflops, float* sm_mhz_nominal) {
    try {
        UB_CUDA(cudaSetDevice(device));
        cudaDeviceProp p;
        UB_CUDA(cudaGetDeviceProperties(&p, device));
        float* d = nullptr;
        UB_CUDA(cudaMalloc(&d, 4));
        const int grid = p.multiProcessorCount * 2;   // 2 x 1024 threads = the SM's 64 resident warps
        cudaEvent_t e0, e1;
        UB_CUDA(cudaEventCreate(&e0));
        UB_CUDA(cudaEventCreate(&e1));
        float best = 0.f;
        for (int rep = 0; rep < 6; ++rep) {   // first repetitions are warm-up (clock ramp); best of the rest
            UB_CUDA(cudaEventRecord(e0));
            for (int k = 0; k < 8; ++k) k_fma_peak<<<grid, 1024>>>(d, 0.999f, 0.001f);
            UB_CUDA(cudaEventRecord(e1));
            UB_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            UB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double flop = 2.0 * CHAINS * double(INNER) * 1024.0 * grid * 8;
            const float tf = float(flop / (ms * 1e-3) / 1e12);
            if (rep >= 2 && tf > best) best = tf;
        }
        UB_CUDA(cudaEventDestroy(e0));
        UB_CUDA(cudaEventDestroy(e1));
        UB_CUDA(cudaFree(d));
        if (tflops) *tflops = best;
        if (sm_mhz_nominal) *sm_mhz_nominal = p.clockRate * 1e-3f;
        return 0;
    } catch (const std::string& e) {
        fprintf(stderr, "ERROR: %s\n", e.c_str());
        return 1;
    }
}
