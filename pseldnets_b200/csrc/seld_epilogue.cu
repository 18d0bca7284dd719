// Backbone-input stage of the feature map, consumer side of the extractors:
//
//   scalar_kernel          the eval-mode per-channel BatchNorm2d "scalar" of the backbones
//                          (src/models/accdoa.py:222-227, 318-321; einv2.py:106-109, 292-295), in place;
//   scalar_wav2img_kernel  the same affine map fused with HTS-AT's reshape_wav2img
//                          (src/models/components/htsat.py:493-511): zero-pad (or crop) the time axis to
//                          r * S frames and fold it r-fold onto the mel axis,
//                          img[b][c][k*M + m][t] = y[b][c][k*S + t][m],  r = S / M.
//
// Both are pure streaming kernels (HBM-bound: every feature value is read once and written once);
// the fold is a 64 x 64 tiled transpose through shared memory with 128-bit accesses on all four
// sides (global load, shared store, shared load, global store).
//
// Rounding follows torch's CPU kernel exactly (tests/golden/epilogue.npz is bit-identical):
//   a = weight * (1 / sqrt(var + eps)),  b = fma(-mean, a, bias),  y = fma(x, a, b).
#include <cuda_runtime.h>
#include <stdint.h>

#include "seld_plan.h"

namespace seld {
namespace epi {

constexpr int kThreads = 256;
constexpr int kTile = 64;                  // tile edge (frames and mel bins)
constexpr int kBlocksPerSM = 6;            // register budget of the fold kernel: 6 x 256 threads x 40 registers

struct Affine { float4 a, b; };

// multiplier / offset of mel bins [m, m + 4) of channel c; identity when the scalar is absent
__device__ __forceinline__ Affine load_affine(const ScalarArgs& s, int c, int m, int M) {
    Affine r;
    if (!s.mean) {
        r.a = make_float4(1.f, 1.f, 1.f, 1.f);
        r.b = make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    const int64_t o = (int64_t)c * M + m;
    const float4 mean = __ldg(reinterpret_cast<const float4*>(s.mean + o));
    const float4 var = __ldg(reinterpret_cast<const float4*>(s.var + o));
    const float4 w = s.weight ? __ldg(reinterpret_cast<const float4*>(s.weight + o)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bias = s.bias ? __ldg(reinterpret_cast<const float4*>(s.bias + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.a.x = __fmul_rn(w.x, __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var.x, s.eps))));
    r.a.y = __fmul_rn(w.y, __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var.y, s.eps))));
    r.a.z = __fmul_rn(w.z, __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var.z, s.eps))));
    r.a.w = __fmul_rn(w.w, __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var.w, s.eps))));
    r.b.x = __fmaf_rn(-mean.x, r.a.x, bias.x);
    r.b.y = __fmaf_rn(-mean.y, r.a.y, bias.y);
    r.b.z = __fmaf_rn(-mean.z, r.a.z, bias.z);
    r.b.w = __fmaf_rn(-mean.w, r.a.w, bias.w);
    return r;
}

__device__ __forceinline__ float4 apply(const float4 v, const Affine& f) {
    return make_float4(__fmaf_rn(v.x, f.a.x, f.b.x), __fmaf_rn(v.y, f.a.y, f.b.y),
                       __fmaf_rn(v.z, f.a.z, f.b.z), __fmaf_rn(v.w, f.a.w, f.b.w));
}

// ------------------------------------------------------------------------------------------------
// In place: one block walks a slab of frames of one (clip, channel) plane.  A thread keeps its
// column quad (4 mel bins) for the whole slab, so the affine terms live in registers.
// Requires kThreads % (M / 4) == 0 (the launcher falls back to kGeneral otherwise).
template <bool kGeneral>
__global__ void __launch_bounds__(kThreads)
scalar_kernel(float* __restrict__ x, const ScalarArgs s, int C, int T, int M, int slabs_per_plane, int frames_per_slab) {
    const int M4 = M >> 2;
    const int plane = blockIdx.x / slabs_per_plane;                  // b * C + c
    const int slab = blockIdx.x - plane * slabs_per_plane;
    const int c = plane % C;
    const int t0 = slab * frames_per_slab;
    const int t1 = min(T, t0 + frames_per_slab);
    float4* p = reinterpret_cast<float4*>(x + (int64_t)plane * T * M);
    const int64_t i0 = (int64_t)t0 * M4, i1 = (int64_t)t1 * M4;
    if constexpr (!kGeneral) {
        const Affine f = load_affine(s, c, (threadIdx.x % M4) * 4, M);
        for (int64_t i = i0 + threadIdx.x; i < i1; i += kThreads) p[i] = apply(p[i], f);
    } else {
        for (int64_t i = i0 + threadIdx.x; i < i1; i += kThreads) {
            const Affine f = load_affine(s, c, (int)(i % M4) * 4, M);
            p[i] = apply(p[i], f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fold: tile = 64 output columns (frames t of piece k) x 64 mel bins of one (clip, channel).
// Thread (tq, m4) owns the 4 x 4 block frames 4tq..4tq+3 x bins 4m4..4m4+3: it loads the four frame
// quads as 128-bit words (16 lanes side by side cover one whole 256-byte frame row), applies the
// affine map and stores the block's four columns as 128-bit words into the transposed tile
// [mel bin][frame]; the tile then leaves as 256-byte row segments.  The tile is unpadded; the
// 16-byte chunk index of row `r` is XOR-swizzled with r / 4, which makes both the column-quad
// stores (8 lanes = 8 different m4, same tq) and the row reads (8 consecutive chunks of one row)
// hit 8 different bank groups.
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
scalar_wav2img_kernel(const float* __restrict__ x, float* __restrict__ img, const ScalarArgs s,
                      int B, int C, int T, int M, int S, int R, int tiles_t, int tiles_m, uint32_t n_tiles) {
    __shared__ __align__(16) float tile[kTile * kTile];
    const int tid = threadIdx.x;
    const int m4 = tid & 15, tq = tid >> 4;
    const int T_in = min(T, R * S);                                   // frames that survive the pad / crop
    int cur_c = -1, cur_mt = -1;
    Affine f;
    // strided walk: all resident blocks sweep the map together (measured: a contiguous run of tiles per block, or a
    // channel-slowest order that would save recomputing the affine terms, both cost 5-8 % in DRAM locality)
    for (uint32_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {  // 32-bit tile arithmetic: the decode is on every tile's path
        uint32_t q = tl;
        const int mt = (int)(q % (uint32_t)tiles_m); q /= (uint32_t)tiles_m;
        const int tt = (int)(q % (uint32_t)tiles_t); q /= (uint32_t)tiles_t;
        const int k = (int)(q % (uint32_t)R); q /= (uint32_t)R;
        const uint32_t plane = q;                                     // b * C + c
        const int c = (int)(plane % (uint32_t)C);
        const int m = mt * kTile + 4 * m4;                            // first of this thread's 4 mel bins
        const bool m_ok = m < M;
        if (c != cur_c || mt != cur_mt) {
            if (m_ok) f = load_affine(s, c, m, M);
            cur_c = c; cur_mt = mt;
        }
        const int tcol = tt * kTile + 4 * tq;                         // first of this thread's 4 output columns
        const int t_in = k * S + tcol;                                // ... = input frames t_in .. t_in + 3
        const float* xp = x + ((int64_t)plane * T + t_in) * M + m;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = m_ok && tcol < S && t_in + i < T_in;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);                    // padding frames stay exactly zero (padded after the scalar)
            if (ok) v[i] = __ldcs(reinterpret_cast<const float4*>(xp + (int64_t)i * M));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (m_ok && tcol < S && t_in + i < T_in) v[i] = apply(v[i], f);
        float* tp = tile + (4 * m4) * kTile + 4 * (tq ^ m4);          // rows 4m4 .. 4m4+3 share the swizzle key m4
        *reinterpret_cast<float4*>(tp + 0 * kTile) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
        *reinterpret_cast<float4*>(tp + 1 * kTile) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
        *reinterpret_cast<float4*>(tp + 2 * kTile) = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
        *reinterpret_cast<float4*>(tp + 3 * kTile) = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
        __syncthreads();
        float* ob = img + ((int64_t)plane * R * M + (int64_t)k * M + mt * kTile) * S + tt * kTile;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = i * kThreads + tid;
            const int row = idx >> 4, c4 = idx & 15;                  // mel bin row of the tile, column quad
            if (mt * kTile + row < M && tt * kTile + 4 * c4 < S)
                __stcs(reinterpret_cast<float4*>(ob + (int64_t)row * S + 4 * c4),
                       *reinterpret_cast<const float4*>(tile + row * kTile + 4 * (c4 ^ (row >> 2))));
        }
        __syncthreads();
    }
}

}  // namespace epi

cudaError_t scalar_launch(float* x, const ScalarArgs& s, int64_t B, int C, int T, int M, int sm_count, cudaStream_t st) {
    using namespace epi;
    const int M4 = M / 4;
    // slabs of whole frames, about 32 KB each, and at least ~4 blocks per SM when the batch is small
    int frames_per_slab = (8192 + M - 1) / M;
    const int64_t planes = B * C;
    while (frames_per_slab > 16 && planes * ((T + frames_per_slab - 1) / frames_per_slab) < 4LL * sm_count) frames_per_slab /= 2;
    const int slabs = (T + frames_per_slab - 1) / frames_per_slab;
    const int64_t blocks = planes * slabs;
    if (blocks > INT32_MAX) return cudaErrorInvalidConfiguration;
    if (kThreads % M4 == 0)
        scalar_kernel<false><<<(unsigned)blocks, kThreads, 0, st>>>(x, s, C, T, M, slabs, frames_per_slab);
    else
        scalar_kernel<true><<<(unsigned)blocks, kThreads, 0, st>>>(x, s, C, T, M, slabs, frames_per_slab);
    return cudaGetLastError();
}

cudaError_t scalar_wav2img_launch(const float* x, float* img, const ScalarArgs& s, int64_t B, int C, int T, int M, int S,
                                  int sm_count, cudaStream_t st) {
    using namespace epi;
    const int R = S / M;
    const int tiles_t = (S + kTile - 1) / kTile, tiles_m = (M + kTile - 1) / kTile;
    const int64_t n_tiles = B * C * R * tiles_t * tiles_m;
    if (n_tiles > INT32_MAX) return cudaErrorInvalidConfiguration;
    static int per_sm = 0;                                            // resident blocks per SM (persistent grid)
    if (per_sm == 0) {
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scalar_wav2img_kernel, kThreads, 0);
        if (e != cudaSuccess) return e;
        per_sm = n > 0 ? n : 1;
    }
    const int64_t resident = (int64_t)per_sm * sm_count;
    const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
    scalar_wav2img_kernel<<<grid, kThreads, 0, st>>>(x, img, s, (int)B, C, T, M, S, R, tiles_t, tiles_m, (uint32_t)n_tiles);
    return cudaGetLastError();
}

}  // namespace seld
