// Device RNG shared by the thermostat and the Monte-Carlo samplers: the reference's RandomGenerator (src/random.h:19-66)
// = Random123 Threefry-4x32-20 keyed (seed, stream, 0, 0) with counter (t_lo, t_hi, atom, draw), u01/uneg11
// (uniform.hpp:145-154,171-180) and Box-Muller (boxmuller.hpp:109-117).
#pragma once
#include "common.cuh"

namespace ub {

// Threefry-4x32 with 20 rounds (Random123 threefry.h: rotation constants :110-117, key-schedule parity 0x1BD11BDA)
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
__device__ __forceinline__ void threefry4x32_20(uint32_t out[4], const uint32_t ctr[4], const uint32_t key[4]) {
    const int R[8][2] = {{10, 26}, {11, 21}, {13, 27}, {23, 5}, {6, 20}, {17, 11}, {25, 10}, {18, 20}};
    uint32_t ks[5];
    ks[4] = 0x1BD11BDAu;
    uint32_t X[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ks[i] = key[i]; X[i] = ctr[i]; ks[4] ^= key[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) X[i] += ks[i];
#pragma unroll
    for (int r = 0; r < 20; ++r) {
        int ra = R[r & 7][0], rb = R[r & 7][1];
        if ((r & 1) == 0) {
            X[0] += X[1]; X[1] = rotl32(X[1], ra); X[1] ^= X[0];
            X[2] += X[3]; X[3] = rotl32(X[3], rb); X[3] ^= X[2];
        } else {
            X[0] += X[3]; X[3] = rotl32(X[3], ra); X[3] ^= X[0];
            X[2] += X[1]; X[1] = rotl32(X[1], rb); X[1] ^= X[2];
        }
        if ((r & 3) == 3) {
            int s = r / 4 + 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) X[i] += ks[(s + i) % 5];
            X[3] += (uint32_t)s;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = X[i];
}
__device__ __forceinline__ float u01_f(uint32_t w) { return w * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }   // uniform.hpp:145-154
__device__ __forceinline__ float uneg11_f(uint32_t w) { return (float)(int32_t)w * 4.6566128730773926e-10f + 2.3283064365386963e-10f; }  // :171-180
__device__ __forceinline__ void boxmuller_f(uint32_t u0, uint32_t u1, float& x, float& y) {   // boxmuller.hpp:109-117
    const float PI = 3.1415926535897932f;
    float a = PI * uneg11_f(u0);
    float s = sinf(a), c = cosf(a);
    float r = sqrtf(-2.f * logf(u01_f(u1)));
    x = s * r;
    y = c * r;
}


// RandomGenerator with its draw counter (c.v[3]++ per draw, random.h:26-30)
struct DeviceRandom {
    uint32_t key[4], ctr[4];
    __device__ DeviceRandom(uint32_t seed, uint32_t stream, uint32_t atom, unsigned long long timestep) {
        key[0] = seed; key[1] = stream; key[2] = 0u; key[3] = 0u;
        ctr[0] = (uint32_t)(timestep & 0xffffffffull); ctr[1] = (uint32_t)(timestep >> 32); ctr[2] = atom; ctr[3] = 0u;
    }
    __device__ void bits(uint32_t* out) { threefry4x32_20(out, ctr, key); ctr[3]++; }
    __device__ void uniform_open_closed(float* u) {
        uint32_t b[4];
        bits(b);
        for (int i = 0; i < 4; ++i) u[i] = u01_f(b[i]);
    }
    __device__ void normal(float* n) {
        uint32_t b[4];
        bits(b);
        boxmuller_f(b[0], b[1], n[0], n[1]);
        boxmuller_f(b[2], b[3], n[2], n[3]);
    }
};

}  // namespace ub
