// Waveform-domain augmentation on the staged batch, the step right before the extractors
// in the reference's training loop (src/models/model_module.py:53-58):
//
//   foa_rotate_kernel   FOA Rotation (src/augment/rotate.py:47-99): per clip, channels 1..3 become a signed
//                       permutation of channels 1..3 (channel 0 is never touched); the reference does this
//                       clip by clip with torch.stack, here one launch rewrites the rotated clips in place.
//   wavmix_kernel       WavMix line src/augment/wavmix.py:50:
//                           x[dst_k] = lam_k * x[dst_k] + (1 - lam_k) * x[src_k]        (all right-hand sides
//                       taken before any assignment), in place: the host orders the pairs along the chains
//                       dst_k -> src_k (seld_wavmix_order) so a thread that walks the list with the previous
//                       source still in registers reads every clip once before it is overwritten.
//
// Both are streaming kernels; a thread owns a fixed sample quad, so there is no cross-thread hazard.
// Rounding is the reference's: a sign flip is exact; the mix is mul, mul, add with (1 - lam) rounded to fp32
// first (torch eager evaluates the expression op by op, no fused multiply-add).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seldfeat.h"
#include "seld_plan.h"

namespace seld {
namespace aug {

constexpr int kThreads = 256;
constexpr int kQuadsPerThread = 4;
constexpr int kChunk = kThreads * kQuadsPerThread * 4;      // samples of one channel handled by one block

template <typename V> struct Vec;
template <> struct Vec<float4> {
    static constexpr int n = 4;
    static __device__ __forceinline__ float4 neg(float4 v, bool s) { return s ? make_float4(-v.x, -v.y, -v.z, -v.w) : v; }
    static __device__ __forceinline__ float4 mix(float4 a, float4 b, float l, float o) {
        return make_float4(__fadd_rn(__fmul_rn(l, a.x), __fmul_rn(o, b.x)), __fadd_rn(__fmul_rn(l, a.y), __fmul_rn(o, b.y)),
                           __fadd_rn(__fmul_rn(l, a.z), __fmul_rn(o, b.z)), __fadd_rn(__fmul_rn(l, a.w), __fmul_rn(o, b.w)));
    }
};
template <> struct Vec<float> {
    static constexpr int n = 1;
    static __device__ __forceinline__ float neg(float v, bool s) { return s ? -v : v; }
    static __device__ __forceinline__ float mix(float a, float b, float l, float o) {
        return __fadd_rn(__fmul_rn(l, a), __fmul_rn(o, b));
    }
};

// code: bits 0-1 / 2-3 / 4-5 = source channel (1..3) of output channel 1 / 2 / 3, bits 8 / 9 / 10 = negate it;
// SELD_ROT_IDENTITY = clip left alone (no traffic)
template <typename V>
__global__ void __launch_bounds__(kThreads)
foa_rotate_kernel(float* __restrict__ x, int64_t L, int64_t stride_b, int64_t stride_c, const int32_t* __restrict__ codes,
                  int chunks_per_clip) {
    const int b = blockIdx.x / chunks_per_clip;
    const int chunk = blockIdx.x - b * chunks_per_clip;
    const int32_t code = __ldg(codes + b);
    if (code == SELD_ROT_IDENTITY) return;
    const int s1 = code & 3, s2 = (code >> 2) & 3, s3 = (code >> 4) & 3;
    const bool n1 = code & 0x100, n2 = code & 0x200, n3 = code & 0x400;
    float* xb = x + (int64_t)b * stride_b;
    const int64_t n = L / Vec<V>::n;                                      // whole vectors (the launcher picks V so that n * V::n == L)
    const int64_t i0 = (int64_t)chunk * (kChunk / Vec<V>::n);
    V* c1 = reinterpret_cast<V*>(xb + stride_c);
    V* c2 = reinterpret_cast<V*>(xb + 2 * stride_c);
    V* c3 = reinterpret_cast<V*>(xb + 3 * stride_c);
    const V* q1 = reinterpret_cast<const V*>(xb + s1 * stride_c);
    const V* q2 = reinterpret_cast<const V*>(xb + s2 * stride_c);
    const V* q3 = reinterpret_cast<const V*>(xb + s3 * stride_c);
#pragma unroll
    for (int j = 0; j < kQuadsPerThread * (4 / Vec<V>::n); ++j) {
        const int64_t i = i0 + (int64_t)j * kThreads + threadIdx.x;
        if (i < n) {
            const V a = q1[i], bq = q2[i], c = q3[i];                     // all three sources before any store
            c1[i] = Vec<V>::neg(a, n1);
            c2[i] = Vec<V>::neg(bq, n2);
            c3[i] = Vec<V>::neg(c, n3);
        }
    }
}

// ops: ordered by seld_wavmix_order.  Thread = one sample quad of one channel, walks the whole list.
template <typename V>
__global__ void __launch_bounds__(kThreads)
wavmix_kernel(float* __restrict__ x, int64_t L, int64_t stride_b, int64_t stride_c, int C,
              const seld_mix_op* __restrict__ ops, int n_ops) {
    const int64_t n = L / Vec<V>::n;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n || c >= C) return;
    V* xc = reinterpret_cast<V*>(x + (int64_t)c * stride_c) + i;
    const int64_t sb = stride_b / Vec<V>::n;
    V cur = V(), head = V();
    for (int k = 0; k < n_ops; ++k) {
        const int4 raw = __ldg(reinterpret_cast<const int4*>(ops + k));   // {dst, src, lam bits, flags}
        const float lam = __int_as_float(raw.z);
        if (raw.w & SELD_MIX_BEGIN) { cur = xc[(int64_t)raw.x * sb]; head = cur; }
        const V nxt = (raw.w & SELD_MIX_USE_HEAD) ? head : xc[(int64_t)raw.y * sb];
        xc[(int64_t)raw.x * sb] = Vec<V>::mix(cur, nxt, lam, __fsub_rn(1.0f, lam));
        cur = nxt;                                                        // if the chain goes on, the next dst is this src
    }
}

}  // namespace aug

static bool vec4_ok(const void* x, int64_t L, int64_t stride_b, int64_t stride_c) {
    return (((uintptr_t)x & 15) == 0) && (L % 4 == 0) && (stride_b % 4 == 0) && (stride_c % 4 == 0);
}

cudaError_t foa_rotate_launch(float* x, int64_t B, int64_t L, int64_t stride_b, int64_t stride_c, const int32_t* codes,
                              cudaStream_t st) {
    using namespace aug;
    const int64_t chunks = (L + kChunk - 1) / kChunk;
    if (B * chunks > INT32_MAX) return cudaErrorInvalidConfiguration;
    if (vec4_ok(x, L, stride_b, stride_c))
        foa_rotate_kernel<float4><<<(unsigned)(B * chunks), kThreads, 0, st>>>(x, L, stride_b, stride_c, codes, (int)chunks);
    else
        foa_rotate_kernel<float><<<(unsigned)(B * chunks), kThreads, 0, st>>>(x, L, stride_b, stride_c, codes, (int)chunks);
    return cudaGetLastError();
}

cudaError_t wavmix_launch(float* x, int C, int64_t L, int64_t stride_b, int64_t stride_c, const seld_mix_op* ops, int n_ops,
                          cudaStream_t st) {
    using namespace aug;
    if (vec4_ok(x, L, stride_b, stride_c)) {
        const int64_t blocks = (L / 4 + kThreads - 1) / kThreads;
        if (blocks > INT32_MAX) return cudaErrorInvalidConfiguration;
        wavmix_kernel<float4><<<dim3((unsigned)blocks, C), kThreads, 0, st>>>(x, L, stride_b, stride_c, C, ops, n_ops);
    } else {
        const int64_t blocks = (L + kThreads - 1) / kThreads;
        if (blocks > INT32_MAX) return cudaErrorInvalidConfiguration;
        wavmix_kernel<float><<<dim3((unsigned)blocks, C), kThreads, 0, st>>>(x, L, stride_b, stride_c, C, ops, n_ops);
    }
    return cudaGetLastError();
}

}  // namespace seld
