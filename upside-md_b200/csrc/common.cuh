// Shared device helpers for the Upside B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define UB_FULL_MASK 0xffffffffu

// padded row width of a node output: 1->1, 2->2, 3,4->4, 5..8->8 (float4-aligned rows for vector loads;
// the reference pads to multiples of 4 as well, vector_math.h:23-25)
__host__ __device__ inline int ub_padded_width(int w) { return w <= 2 ? w : ((w + 3) & ~3); }

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return f3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3& operator+=(f3& a, f3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ f3& operator-=(f3& a, f3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float mag2(f3 a) { return dot(a, a); }
__device__ __forceinline__ f3 ld3(const float* p) { return f3{p[0], p[1], p[2]}; }
__device__ __forceinline__ f3 ld3v(const float* p) {  // 16-byte aligned row
    float4 v = *reinterpret_cast<const float4*>(p);
    return f3{v.x, v.y, v.z};
}
__device__ __forceinline__ void st3(float* p, f3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ void add3(float* p, f3 a) { p[0] += a.x; p[1] += a.y; p[2] += a.z; }
__device__ __forceinline__ void atomic_add3(float* p, f3 a) {
    atomicAdd(p + 0, a.x); atomicAdd(p + 1, a.y); atomicAdd(p + 2, a.z);
}
// the same into a 16-byte aligned row whose fourth float is padding (or tolerates +0): ONE 16-byte reduction
// (REDG.ADD.F32x4, sm_90+) instead of three 4-byte ones
__device__ __forceinline__ void atomic_add3v(float* p16, f3 a) {
    atomicAdd(reinterpret_cast<float4*>(p16), make_float4(a.x, a.y, a.z, 0.f));
}

// One-instruction reciprocal (MUFU.RCP).  `__fdividef(x, y)` and `1.f / y` guard against denormal denominators with a compare and
// two predicated multiplies per call (five instructions); where the denominator is known to be a normal number - 1e-10 plus a
// probability, a sum of L1-normalised messages, a maximum of beliefs - the guarded and the plain form return the same bits.
__device__ __forceinline__ float rcp_fast(float y) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
}

// quaternion (a,b,c,d) -> row-major rotation matrix; reference affine.h:99-108
__device__ __forceinline__ void quat_to_rot(float* U, const float* q) {
    float a = q[0], b = q[1], c = q[2], d = q[3];
    U[0] = a * a + b * b - c * c - d * d; U[1] = 2.f * b * c - 2.f * a * d;     U[2] = 2.f * b * d + 2.f * a * c;
    U[3] = 2.f * b * c + 2.f * a * d;     U[4] = a * a - b * b + c * c - d * d; U[5] = 2.f * c * d - 2.f * a * b;
    U[6] = 2.f * b * d - 2.f * a * c;     U[7] = 2.f * c * d + 2.f * a * b;     U[8] = a * a - b * b - c * c + d * d;
}
__device__ __forceinline__ f3 rot_apply(const float* U, f3 r) {
    return f3{U[0] * r.x + U[1] * r.y + U[2] * r.z, U[3] * r.x + U[4] * r.y + U[5] * r.z,
              U[6] * r.x + U[7] * r.y + U[8] * r.z};
}
__device__ __forceinline__ f3 rot_apply_inv(const float* U, f3 r) {
    return f3{U[0] * r.x + U[3] * r.y + U[6] * r.z, U[1] * r.x + U[4] * r.y + U[7] * r.z,
              U[2] * r.x + U[5] * r.y + U[8] * r.z};
}

// compact_sigmoid(x, sharpness): 1 for y<-1, 0 for y>1, else (y+2)(y-1)^2/4; reference vector_math.h:639-658
__device__ __forceinline__ void compact_sigmoid(float x, float sharpness, float& val, float& deriv) {
    float y = x * sharpness;
    if (y < -1.f) { val = 1.f; deriv = 0.f; }
    else if (y > 1.f) { val = 0.f; deriv = 0.f; }
    else { val = 0.25f * (y + 2.f) * (y - 1.f) * (y - 1.f); deriv = (sharpness * 0.75f) * (y * y - 1.f); }
}
// sigmoid(x) -> (1/(1+e^-x), e^-x/(1+e^-x)^2); reference vector_math.h:626-631
__device__ __forceinline__ void sigmoid_vd(float x, float& val, float& deriv) {
    float z = __expf(-x);
    float w = 1.f / (1.f + z);
    val = w;
    deriv = z * w * w;
}

// Uniform cubic B-spline, de Boor form; x >= 1, coefficients c[b-1..b+2] with b=int(x).
// Same recurrence as the reference's uniform_deBoor_algorithm (spline.h:136-174).
__device__ __forceinline__ void deboor_core(float c0, float c1, float c2, float c3, float y, float& val, float& der) {
    const float third = 1.f / 3.f;
    float a11 = third * (y + 2.f), a12 = third * (y + 1.f), a13 = third * y;
    float c11 = fmaf(1.f - a11, c0, a11 * c1), d11 = c1 - c0;
    float c12 = fmaf(1.f - a12, c1, a12 * c2), d12 = c2 - c1;
    float c13 = fmaf(1.f - a13, c2, a13 * c3), d13 = c3 - c2;
    float a22 = 0.5f * (y + 1.f), a23 = 0.5f * y;
    float c22 = fmaf(1.f - a22, c11, a22 * c12), d22 = fmaf(1.f - a22, d11, a22 * d12);
    float c23 = fmaf(1.f - a23, c12, a23 * c13), d23 = fmaf(1.f - a23, d12, a23 * d13);
    val = fmaf(1.f - y, c22, y * c23);
    der = fmaf(1.f - y, d22, y * d23);
}
// The same cubic B-spline in basis-weight form: value = sum_k w[k] c[b-1+k], d/dy = sum_k d[k] c[b-1+k].  Two splines that
// share y (the wide and narrow radial profiles of a quadspline) share the weights, and the evaluation is 4 + 4 FMAs per
// spline instead of the 3-level recurrence.  Algebraically identical to deboor_core (agreement ~1 ulp of the value).
__device__ __forceinline__ void bspline_weights(float y, float* __restrict__ w, float* __restrict__ d) {
    const float z = 1.f - y, y2 = y * y, z2 = z * z;
    w[0] = (1.f / 6.f) * z2 * z;
    w[3] = (1.f / 6.f) * y2 * y;
    w[1] = fmaf(y2, fmaf(0.5f, y, -1.f), 2.f / 3.f);     // (3y^3 - 6y^2 + 4)/6
    w[2] = fmaf(z2, fmaf(0.5f, z, -1.f), 2.f / 3.f);     // the mirror image, (3z^3 - 6z^2 + 4)/6
    d[0] = -0.5f * z2;
    d[3] = 0.5f * y2;
    d[1] = y * fmaf(1.5f, y, -2.f);                      // (3y^2 - 4y)/2
    d[2] = -z * fmaf(1.5f, z, -2.f);
}
__device__ __forceinline__ void bspline_apply(const float* __restrict__ w, const float* __restrict__ d, float c0, float c1, float c2,
                                              float c3, float& val, float& der) {
    val = fmaf(w[0], c0, fmaf(w[1], c1, fmaf(w[2], c2, w[3] * c3)));
    der = fmaf(d[0], c0, fmaf(d[1], c1, fmaf(d[2], c2, d[3] * c3)));
}

// unclamped evaluation; `n` = number of coefficients available, used only to keep the 4-wide window in range
// (the reference reads one float past the block when x lands exactly on the last knot, with weight 0)
__device__ __forceinline__ void deboor_vd(const float* __restrict__ c, int n, float x, float& val, float& der) {
    int b = (int)x;
    b = max(1, min(b, n - 3));
    float y = x - (float)b;
    deboor_core((*(c + b - 1)), (*(c + b)), (*(c + b + 1)), (*(c + b + 2)), y, val, der);
}
// clamped spline, Float4 flavour of the reference (spline.h:275-310): x<1 -> left value, x>=n-2 -> right value
__device__ __forceinline__ void clamped_deboor_vd(const float* __restrict__ c, int n, float x, float& val, float& der) {
    if (x < 1.f) {
        val = (1.f / 6.f) * (*(c)) + (2.f / 3.f) * (*(c + 1)) + (1.f / 6.f) * (*(c + 2));
        der = 0.f;
    } else if (x >= (float)(n - 2)) {
        val = (1.f / 6.f) * (*(c + n - 3)) + (2.f / 3.f) * (*(c + n - 2)) + (1.f / 6.f) * (*(c + n - 1));
        der = 0.f;
    } else {
        int b = (int)x;
        float y = x - (float)b;
        deboor_core((*(c + b - 1)), (*(c + b)), (*(c + b + 1)), (*(c + b + 2)), y, val, der);
    }
}

// sum over the lanes of an aligned power-of-two sub-group of width G (G <= 32)
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(UB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(UB_FULL_MASK, v, o));
    return v;
}
// block-wide sum; every thread must call; result valid in thread 0.  scratch: >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? scratch[lane] : 0.f;
        v = warp_sum(v);
    }
    return v;
}

// Blondel & Karplus dihedral and its gradient; reference vector_math.h:704-735
__device__ __forceinline__ float dihedral_germ(f3 r1, f3 r2, f3 r3, f3 r4, f3& d1, f3& d2, f3& d3, f3& d4) {
    f3 F = r1 - r2, G = r2 - r3, H = r4 - r3;
    f3 A = cross(F, G), B = cross(H, G), C = cross(B, A);
    float iA = 1.f / mag2(A), iB = 1.f / mag2(B);
    float G2 = mag2(G), iG = rsqrtf(G2), Gm = G2 * iG;
    d1 = (-Gm * iA) * A;
    d4 = (Gm * iB) * B;
    f3 fm = (dot(F, G) * iA * iG) * A - (dot(H, G) * iB * iG) * B;
    d2 = fm - d1;
    d3 = -d4 - fm;
    return atan2f(dot(C, G), dot(A, B) * Gm);
}

// rotation matrix (row-major) of `angle` about the unit vector `axis`; reference affine.h:49-64
__device__ __forceinline__ void axis_angle_to_rot(float* U, float angle, f3 axis) {
    const float x = axis.x, y = axis.y, z = axis.z;
    const float c = cosf(angle), s = sinf(angle), C = 1.f - c;
    U[0] = x * x * C + c;     U[1] = x * y * C - z * s; U[2] = x * z * C + y * s;
    U[3] = y * x * C + z * s; U[4] = y * y * C + c;     U[5] = y * z * C - x * s;
    U[6] = z * x * C - y * s; U[7] = z * y * C + x * s; U[8] = z * z * C + c;
}
