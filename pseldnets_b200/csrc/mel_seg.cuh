// Segment-walk mel projection shared by the iv3 and MIC kernels (the iv2 kernel carries its own
// copy of the same scheme).
//
// A triangular bank whose bands overlap only their neighbours has a "segment" structure: the bins
// between two consecutive band centres (segment s) feed exactly bands s-1 and s, with weights
// (a_k, b_k).  So  out[m] = V[m] + U[m+1],  U[s] = sum_{k in s} a_k q_k,  V[s] = sum_{k in s} b_k q_k.
//   step 1 (mel_walk):    lane c owns bins [16c, 16c+16) (+ bin 512 for c = 31) of each row; it
//                         reads them with conflict-free 128-bit loads (rows are XOR-swizzled per
//                         chunk), accumulates (U, V) of the current run as ONE FFMA2 per bin and row,
//                         and stores each finished run's pair at float2 index g (its run number)
//                         in the first words of the same row.
//   step 2 (mel_combine): lane m sums the V of segment m's runs and the U of segment m+1's runs.
// Row layout: bin k = 16c + j lives at word 16c + 4*((j>>2) ^ ((c>>1)&3)) + (j&3); a writer holding
// bin lane + 32*kb therefore stores at 32*kb + wofs[kb&3] (see the kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft32.cuh"

namespace seld {
namespace melseg {

constexpr int kRowWords = 528;            // 33 chunks of 16 bins (bin 512 opens chunk 32)
constexpr int kWabStride = 36;            // floats per lane in the (a, b) table: 17 float2 + pad; 36*l mod 32 = 4l

__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void lane_offsets(int lane, int (&wofs)[4], int (&rofs)[4]) {
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        wofs[x] = 16 * (lane >> 4) + 4 * (((lane >> 2) & 3) ^ x) + (lane & 3);   // writer: bin lane+32kb -> 32kb + wofs[kb&3]
        rofs[x] = 16 * lane + 4 * (x ^ ((lane >> 1) & 3));                       // reader: quad x of chunk `lane`
    }
}

__device__ __forceinline__ void load_weights(const float* wab_s, int lane, float2 (&wv)[17]) {
    const float4* wp = reinterpret_cast<const float4*>(wab_s + lane * kWabStride);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = wp[i];
        wv[2 * i] = make_float2(v.x, v.y);
        wv[2 * i + 1] = make_float2(v.z, v.w);
    }
    wv[16] = *reinterpret_cast<const float2*>(wab_s + lane * kWabStride + 32);
}

// Walk NF rows (compile-time row ids) of the frame region: per-run (U, V) partial sums are left in
// the first words of each row.  wv: the lane's 17 (a, b) weight pairs.
template <int NF, int R0, int R1, int R2, int R3 = 0>
__device__ __forceinline__ void mel_walk(float* R, const float2 (&wv)[17], const int (&rofs)[4],
                                         uint32_t runmask, int g0, int lane) {
    constexpr int rows[4] = {R0, R1, R2, R3};
    float q[NF][17];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float* row = R + rows[f] * kRowWords;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(row + rofs[i]);
            q[f][4 * i] = v.x; q[f][4 * i + 1] = v.y; q[f][4 * i + 2] = v.z; q[f][4 * i + 3] = v.w;
        }
        q[f][16] = lane == 31 ? row[512] : 0.0f;
    }
    __syncwarp();                                                   // every lane holds its bins: rows may be overwritten
    float2 acc[NF];
    int po = g0;                                                    // float2 index of the lane's current run
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = vmuls(wv[0], q[f][0]);
    // branch-free: where a new run starts the finished pair is stored and the accumulator restarts
    // (acc * keep, keep = 0): one FMUL2 + one FFMA2 + one predicated store per bin and row
    static_for<1, 17>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        const bool start = (runmask >> j) & 1u;
        const float keep = start ? 0.0f : 1.0f;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if (start) reinterpret_cast<float2*>(R + rows[f] * kRowWords)[po] = acc[f];
            acc[f] = __ffma2_rn(wv[j], make_float2(q[f][j], q[f][j]), vmuls(acc[f], keep));
        }
        po += start ? 1 : 0;
    });
#pragma unroll
    for (int f = 0; f < NF; ++f) reinterpret_cast<float2*>(R + rows[f] * kRowWords)[po] = acc[f];
}

// Band-per-lane combine of NF rows: out[m] = sum V(runs of segment m) + sum U(runs of segment m+1).
template <int NF, int R0, int R1, int R2, int R3, bool kDb>
__device__ __forceinline__ void mel_combine(const float* R, const int* gseg_s, int M, int lane, float amin,
                                            float* const (&o)[4]) {
    constexpr int rows[4] = {R0, R1, R2, R3};
    for (int m = lane; m < M; m += 32) {
        const int ga = gseg_s[m], gb = gseg_s[m + 1], gc = gseg_s[m + 2];
        float v[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) v[f] = 0.0f;
        for (int g = ga; g < gb; ++g) {
#pragma unroll
            for (int f = 0; f < NF; ++f) v[f] += reinterpret_cast<const float2*>(R + rows[f] * kRowWords)[g].y;
        }
        for (int g = gb; g < gc; ++g) {
#pragma unroll
            for (int f = 0; f < NF; ++f) v[f] += reinterpret_cast<const float2*>(R + rows[f] * kRowWords)[g].x;
        }
#pragma unroll
        for (int f = 0; f < NF; ++f)
            o[f][m] = kDb ? 3.01029995663981195f * lg2_ftz(fmaxf(v[f], amin)) : v[f];   // 10*log10(max(v, amin))
    }
}

}  // namespace melseg
}  // namespace seld
