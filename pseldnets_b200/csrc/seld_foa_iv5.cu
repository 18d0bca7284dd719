// FOA log-mel + intensity-vector kernel, tensor-core generation ("iv5"): the headline path.
//
//   LogmelIV_Extractor.forward (feature.py:39-56) + intensityvector (feature.py:93-117), 4 channels.
//
// Same warp-per-frame packed fp32 transform as iv2 (two 1024-point FFTs in FADD2/FMUL2/FFMA2 registers),
// but the mel projection -- `mel_scale(|X|^2)` (feature.py:50) and the three `(I_j / normal) @ melW`
// contractions (feature.py:112-114) -- runs on the 5th-generation tensor cores:
//
//   * the eight transform warps of a block write their frame's seven per-bin rows (4 powers, 3 normalised
//     intensities) as bf16 hi/lo pairs straight into the shared-memory A operand of a tcgen05.mma:
//     MN-major, no swizzle, 16 rows per frame ([P0 P2 P1 P3]hi [..]lo | [n1 n3 n2 0]hi [..]lo), one
//     16-byte store per 8-row block and bin; 8 frames = one M = 128 tile;
//   * the mel bank is the B operand: 33 chunks of 16 bins, each a K-major tile holding only the bands the
//     chunk touches (a window of 8 or 16 bands), hi and lo parts interleaved along N, so ONE tcgen05.mma per
//     chunk accumulates  A_hi*B_hi, A_hi*B_lo, A_lo*B_hi and A_lo*B_lo  into a column window of the
//     fp32 accumulator in tensor memory (the first chunk is issued full-width and overwrites: that is the
//     zero-initialisation);
//   * whichever warp hands its rows over last issues the 33 MMAs of the tile (one elected lane, operands
//     from the constant bank; a ninth warp would cost the register allocation of four); every warp picks
//     the finished tile up two frames later (tcgen05.ld of its 32-lane quarter), adds the hi/lo halves
//     (in-thread for B, one shuffle for A), takes 10*log10 and stores.
//
// Warps stay decoupled: a transform warp only waits for the MMA of the PREVIOUS tile before it reuses its
// row area as FFT exchange buffer, roughly a third of a frame after it handed the rows over.
//
// Error budget of the split (measured with tools/umma_probe.py on the B200): 2 x bf16 carries 16-17
// mantissa bits: <= 9e-6 of the row maximum, 6e-5 dB on the log-mel rows -- against 1e-4 of the block maximum.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "fft32.cuh"
#include "seld_plan.h"

namespace seld {
namespace iv5 {

constexpr int kW = 8;                     // transform warps (= frames per tile)
constexpr int kThreads = kW * 32;
constexpr int kBins = 528;                // 33 chunks of 16 bins (513 used)
constexpr int kPlane = kBins * 16;        // bytes of one 8-row block of one frame: 16 B per bin
constexpr int kSlot = 2 * kPlane;         // a frame: power block, then intensity block
constexpr int kABytes = kW * kSlot;       // 135168
constexpr int kTwStride = 68;             // floats per lane in the twiddle table: 32 float2 + pad
constexpr int kWinStride = 36;            // floats per lane in the window table: 32 floats + pad
constexpr int kXStride = 34;              // exchange buffer row stride in float2 (even: 128-bit reads)
constexpr int kAccCols = 128;             // accumulator columns per buffer: (hi, lo) interleaved x 64 bands
constexpr int kTmemCols = 2 * kAccCols;   // double-buffered

static_assert(32 * kXStride * 8 <= kSlot, "exchange buffer must fit the warp's own frame slot");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// two fp32 -> packed bf16 pair (lo half = a, hi half = b), round to nearest even
__device__ __forceinline__ uint32_t bf16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// hi/lo split of a pair: h = bf16(v), l = bf16(v - h)
__device__ __forceinline__ void split_pair(float2 v, uint32_t& h, uint32_t& l) {
    h = bf16x2(v.x, v.y);
    const float2 hf = make_float2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
    const float2 r = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), v);
    l = bf16x2(r.x, r.y);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

}  // namespace iv5

template <typename TIn>
__global__ void __launch_bounds__(iv5::kThreads, 1)
foa_iv5_kernel(const FoaArgs a, const PlanDev pd, const MelTiles mt) {
    using namespace iv5;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* A_s = smem_raw;                                           // kW frame slots
    unsigned char* B_s = A_s + kABytes;                                      // mel tiles (mt.b_bytes)
    float* tw_s = reinterpret_cast<float*>(B_s + mt.b_bytes);                // [lane][kTwStride]
    float* win_s = tw_s + 32 * kTwStride;                                    // [lane][kWinStride]
    uint64_t* bars = reinterpret_cast<uint64_t*>(win_s + 32 * kWinStride);   // [0] rows full, [1], [2] MMA done (per accumulator)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    unsigned int* arrivals = tmem_slot + 1;                                  // counts row hand-overs: the 8th of a tile issues its MMAs

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float in_scale = a.in_scale;
    for (int i = tid; i < 1024; i += kThreads) {
        const int l = i & 31, r = i >> 5;                                    // pd.tw is [ka][lane], pd.win is [32*m + lane]
        const float2 w = pd.tw[brev5(r) * 32 + l];
        tw_s[l * kTwStride + 2 * r] = w.x; tw_s[l * kTwStride + 2 * r + 1] = w.y;
        win_s[l * kWinStride + r] = pd.win[i] * in_scale;
    }
    for (int i = tid; i < mt.b_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(B_s)[i] = mt.b_img[i];
    for (int i = tid; i < kABytes / 16; i += kThreads) reinterpret_cast<uint4*>(A_s)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        mbar_init(bars + 0, kW);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        *arrivals = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // tiles of this block: blockIdx.x, + gridDim.x, ... (a tile = 8 consecutive frames of one clip)
    const int n_my = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // ---------------------------------------------------------------------- MMA issue (one lane of the last warp to arrive)
    // A: MN-major, no swizzle: 8-row blocks `kPlane` apart (SBO), 8-bin groups 128 B apart (LBO)
    const uint64_t a_desc = ((uint64_t)1 << 46) | ((uint64_t)(kPlane >> 4) << 32) | ((uint64_t)(128 >> 4) << 16) |
                            (uint64_t)((smem_u32(A_s) >> 4) & 0x3fff);
    // B: K-major, no swizzle: 8-column groups 256 B apart (SBO), the two 8-bin halves 128 B apart (LBO)
    const uint64_t b_desc = ((uint64_t)1 << 46) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)(128 >> 4) << 16) |
                            (uint64_t)((smem_u32(B_s) >> 4) & 0x3fff);
    auto issue_tile = [&](int i) {
        // kind::f16, D = f32, A = B = bf16, A MN-major, B K-major, M = 128; N goes in per chunk
        constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((128u >> 4) << 24);
        mbar_wait(bars + 0, i & 1);                                          // completes at once (this lane arrived last): acquire
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d0 = tmem + (uint32_t)((i & 1) * kAccCols);
        static_for<0, 33>([&](auto ci) {
            constexpr int c = decltype(ci)::value;
            const uint32_t e = mt.chunk[c];                                  // b_off16 | dcol << 16 | (N >> 3) << 24
            const uint64_t ad = a_desc + (uint64_t)(c * 16);                 // 16 bins x 16 B = 256 B per chunk
            const uint64_t bd = b_desc + (uint64_t)(e & 0xffffu);
            const uint32_t dcol = d0 + ((e >> 16) & 0xffu);
            const uint32_t idesc = kIdesc | ((e >> 24) << 17);
            const uint32_t acc = c > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol),
                "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                : "memory");
        });
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + 1 + (i & 1)))
                     : "memory");
    };

    {
        // ------------------------------------------------------------------ transform warps
        unsigned char* slot = A_s + warp * kSlot;
        float2* scratch = reinterpret_cast<float2*>(slot);                   // 32 x 34 float2 exchange buffer, aliases the rows
        const int hop = pd.hop, M = pd.n_mels;
        const float eps = pd.eps, amin = pd.amin;
        const int64_t ch_stride = (int64_t)a.T * M;

        // epilogue of a finished tile: this warp owns TMEM lanes [32q, 32q+32) (q = warp & 3) = frames 2q, 2q+1 of
        // the tile, 16 rows each, and the column half `warp >> 2` (32 bands)
        auto epilogue = [&](int tile, int buf) {
            const int b = tile / a.tiles_per_clip, tr = tile - b * a.tiles_per_clip;
            uint32_t v[32], u[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * kAccCols + 64 * (warp >> 2));
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // columns come as (hi, lo) pairs of one band; rows r and r ^ 4 are the hi and lo halves of the data
            float s[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                s[j] = __uint_as_float(v[2 * j]) + __uint_as_float(v[2 * j + 1]);
                s[16 + j] = __uint_as_float(u[2 * j]) + __uint_as_float(u[2 * j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
            const int r16 = lane & 15, f = 2 * (warp & 3) + (lane >> 4);
            const int t = tr * kW + f;
            const int q = r16 & 3;                                           // row within its 4-row group
            const bool is_p = r16 < 8;
            // rows: P0 P2 P1 P3 | n1 n3 n2 -   ->   output channels 0 2 1 3 | C+0 C+2 C+1
            const int ch = is_p ? ((q & 1) << 1 | (q >> 1)) : a.C + (q == 0 ? 0 : (q == 1 ? 2 : 1));
            const bool live = t < a.T && q + (is_p ? 0 : 1) < 4;             // intensity block has 3 rows
            if (is_p) {
#pragma unroll
                for (int j = 0; j < 32; ++j) s[j] = 3.01029995663981195f * lg2_ftz(fmaxf(s[j], amin));   // 10*log10(max(v, amin))
            }
            if (live) {
                // the hi-row lane stores bands [0, 16) of this warp's 32, the lo-row lane bands [16, 32)
                const int half = (r16 >> 2) & 1;
                float* o = a.out + ((int64_t)b * a.Cout + ch) * ch_stride + (int64_t)t * M + 32 * (warp >> 2) + 16 * half;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w4 = half ? make_float4(s[16 + 4 * j], s[17 + 4 * j], s[18 + 4 * j], s[19 + 4 * j])
                                           : make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
                    *reinterpret_cast<float4*>(o + 4 * j) = w4;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        };

        int tile = blockIdx.x;
        for (int i = 0; i < n_my; ++i, tile += gridDim.x) {
            if (i >= 2) {                                                    // tile i-2 is long finished: drain it
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                epilogue(tile - 2 * (int)gridDim.x, i & 1);
            }
            const int b = tile / a.tiles_per_clip;
            const int t = (tile - b * a.tiles_per_clip) * kW + warp;
            const bool has = t < a.T;
            float2 re[32], im[32];
            if (has) {
                const TIn* xb = reinterpret_cast<const TIn*>(a.x) + (int64_t)b * a.stride_b;
                // ---------------- load + window: re = (ch0, ch2), im = (ch1, ch3)
                const int64_t s0 = (int64_t)t * hop - 512;
                if (s0 >= 0 && s0 + 1024 <= a.L) {
                    const TIn* p0 = xb + s0 + lane;
                    const TIn* p1 = p0 + a.stride_c;
                    const TIn* p2 = p1 + a.stride_c;
                    const TIn* p3 = p2 + a.stride_c;
                    static_for<0, 32>([&](auto mi) {
                        constexpr int m = decltype(mi)::value;
                        re[m] = make_float2((float)__ldg(p0 + 32 * m), (float)__ldg(p2 + 32 * m));
                        im[m] = make_float2((float)__ldg(p1 + 32 * m), (float)__ldg(p3 + 32 * m));
                    });
                } else {                                                     // reflect padding at the clip edges
                    static_for<0, 32>([&](auto mi) {
                        constexpr int m = decltype(mi)::value;
                        int64_t sidx = s0 + 32 * m + lane;
                        if (sidx < 0) sidx = -sidx;
                        if (sidx >= a.L) sidx = 2 * (a.L - 1) - sidx;
                        const TIn* p = xb + sidx;
                        re[m] = make_float2((float)__ldg(p), (float)__ldg(p + 2 * a.stride_c));
                        im[m] = make_float2((float)__ldg(p + a.stride_c), (float)__ldg(p + 3 * a.stride_c));
                    });
                }
                static_for<0, 8>([&](auto mi) {
                    constexpr int m4 = decltype(mi)::value;
                    const float4 w4 = *reinterpret_cast<const float4*>(win_s + lane * kWinStride + 4 * m4);
                    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        re[4 * m4 + e] = vmuls(re[4 * m4 + e], w[e]);
                        im[4 * m4 + e] = vmuls(im[4 * m4 + e], w[e]);
                    }
                });
                // ---------------- two 1024-point FFTs at once: 32-pt, twiddle, exchange, 32-pt
                fft32(re, im);
                static_for<0, 16>([&](auto pi) {
                    constexpr int p2 = decltype(pi)::value;                  // positions 2*p2, 2*p2+1
                    const float4 w4 = *reinterpret_cast<const float4*>(tw_s + lane * kTwStride + 4 * p2);
                    if constexpr (p2 > 0) {                                  // position 0 is ka = 0: twiddle 1
                        const float2 r = re[2 * p2], i2 = im[2 * p2];
                        re[2 * p2] = vfmas(i2, -w4.y, vmuls(r, w4.x));
                        im[2 * p2] = vfmas(i2, w4.x, vmuls(r, w4.y));
                    }
                    const float2 r = re[2 * p2 + 1], i2 = im[2 * p2 + 1];
                    re[2 * p2 + 1] = vfmas(i2, -w4.w, vmuls(r, w4.z));
                    im[2 * p2 + 1] = vfmas(i2, w4.z, vmuls(r, w4.w));
                });
            }
            // the rows of the previous tile must have been consumed before the slot becomes exchange buffer again
            if (i >= 1) mbar_wait(bars + 1 + ((i - 1) & 1), ((i - 1) >> 1) & 1);
            if (has) {
                static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; scratch[brev5(p) * kXStride + lane] = re[p]; });
                __syncwarp();
                static_for<0, 16>([&](auto ji) {
                    constexpr int j = decltype(ji)::value;
                    const float4 v = *reinterpret_cast<const float4*>(scratch + lane * kXStride + 2 * j);
                    re[2 * j] = make_float2(v.x, v.y); re[2 * j + 1] = make_float2(v.z, v.w);
                });
                __syncwarp();
                static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; scratch[brev5(p) * kXStride + lane] = im[p]; });
                __syncwarp();
                static_for<0, 16>([&](auto ji) {
                    constexpr int j = decltype(ji)::value;
                    const float4 v = *reinterpret_cast<const float4*>(scratch + lane * kXStride + 2 * j);
                    im[2 * j] = make_float2(v.x, v.y); im[2 * j + 1] = make_float2(v.z, v.w);
                });
                __syncwarp();
                fft32(re, im);                                               // position p: Z[lane + 32*brev5(p)]

                // ---------------- per-bin quantities -> bf16 hi/lo rows of the A operand
                const int src = (32 - lane) & 31;
                const bool lane0 = lane == 0;
                uint4* rowP = reinterpret_cast<uint4*>(slot) + lane;          // bin k = lane + 32*kb -> 16 B at k*16
                uint4* rowI = reinterpret_cast<uint4*>(slot + kPlane) + lane;
                static_for<0, 17>([&](auto kbi) {
                    constexpr int kb = decltype(kbi)::value;
                    constexpr int p = brev5(kb & 31);
                    const float2 zr = re[p], zi = im[p];
                    float2 pr, pi;
                    if constexpr (kb == 16) {
                        pr = zr; pi = zi;
                    } else {
                        constexpr int pp = brev5(31 - kb), p0 = brev5((32 - kb) & 31);
                        const float sx = __shfl_sync(0xffffffffu, re[pp].x, src), sy = __shfl_sync(0xffffffffu, re[pp].y, src);
                        const float tx = __shfl_sync(0xffffffffu, im[pp].x, src), ty = __shfl_sync(0xffffffffu, im[pp].y, src);
                        pr = make_float2(lane0 ? re[p0].x : sx, lane0 ? re[p0].y : sy);
                        pi = make_float2(lane0 ? im[p0].x : tx, lane0 ? im[p0].y : ty);
                    }
                    // window was pre-scaled by 0.5: A = Z[k] + conj(Z[N-k]), B = (Z[k] - conj(Z[N-k])) / i
                    const float2 ar = vadd(zr, pr), ai = vsub(zi, pi);        // (X0, X2)
                    const float2 br = vadd(zi, pi), bi = vsub(pr, zr);        // (X1, X3)
                    const float2 p02 = __ffma2_rn(ai, ai, __fmul2_rn(ar, ar));
                    const float2 p13 = __ffma2_rn(bi, bi, __fmul2_rn(br, br));
                    const float2 i13 = vfmas(bi, ai.x, vmuls(br, ar.x));      // Re(conj(X0) X1), Re(conj(X0) X3)
                    const float i2 = fmaf(ai.x, ai.y, ar.x * ar.y);           // Re(conj(X0) X2)
                    const float sq = fmaf(i13.y, i13.y, fmaf(i2, i2, i13.x * i13.x));
                    const float inv = rcp_ftz(sqrt_ftz(sq) + eps);
                    uint4 wp, wi;
                    split_pair(p02, wp.x, wp.z);
                    split_pair(p13, wp.y, wp.w);
                    split_pair(vmuls(i13, inv), wi.x, wi.z);
                    split_pair(make_float2(i2 * inv, 0.0f), wi.y, wi.w);
                    if constexpr (kb < 16) {
                        rowP[32 * kb] = wp;
                        rowI[32 * kb] = wi;
                    } else {                                                 // bin 512 (lane 0); bins 513..527 are padding: zeros
                        if (lane < 16) {
                            rowP[32 * kb] = lane0 ? wp : make_uint4(0u, 0u, 0u, 0u);
                            rowI[32 * kb] = lane0 ? wi : make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                });
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bars + 0);
                if ((atomicAdd(arrivals, 1u) & (kW - 1)) == kW - 1) issue_tile(i);
            }
            __syncwarp();
        }
        // drain the last two tiles
        if (n_my >= 2) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue(tile - 2 * (int)gridDim.x, n_my & 1);
        }
        if (n_my >= 1) {
            mbar_wait(bars + 1 + ((n_my - 1) & 1), ((n_my - 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue(tile - (int)gridDim.x, (n_my - 1) & 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// ---------------------------------------------------------------------------------------------
static size_t iv5_smem_bytes(const MelTiles& mt) {
    return (size_t)iv5::kABytes + mt.b_bytes + (size_t)(32 * iv5::kTwStride + 32 * iv5::kWinStride) * sizeof(float) + 64;
}

bool foa_iv5_supported(const PlanDev& pd, const MelTiles& mt, size_t smem_optin) {
    return mt.ok && pd.n_mels == 64 && iv5_smem_bytes(mt) <= smem_optin;
}

int foa_iv5_frames_per_tile() { return iv5::kW; }

template <typename TIn>
static cudaError_t iv5_launch_t(const FoaArgs& a, const PlanDev& pd, const MelTiles& mt, int sm_count, cudaStream_t st) {
    const size_t smem = iv5_smem_bytes(mt);
    static std::atomic<uint64_t> attr_done{0};                               // per device, once per process and instantiation
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_relaxed) & bit)) {
        e = cudaFuncSetAttribute(foa_iv5_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.fetch_or(bit, std::memory_order_relaxed);
    }
    const int gx = sm_count < a.n_tiles ? sm_count : a.n_tiles;
    foa_iv5_kernel<TIn><<<gx, iv5::kThreads, smem, st>>>(a, pd, mt);
    return cudaGetLastError();
}

cudaError_t foa_iv5_launch(const FoaArgs& a, const PlanDev& pd, const MelTiles& mt, int sm_count, cudaStream_t st) {
    return a.in_i16 ? iv5_launch_t<int16_t>(a, pd, mt, sm_count, st) : iv5_launch_t<float>(a, pd, mt, sm_count, st);
}

}  // namespace seld
