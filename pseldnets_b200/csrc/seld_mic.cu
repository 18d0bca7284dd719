// MIC feature kernel: log-mel (top_db-limited) + GCC-PHAT for a 4-microphone array.
//
// Reference behaviour restated (not ported): /root/reference/src/utils/feature.py
//   Features_Extractor_MIC._spectrogram :146-153 (librosa.stft: zero 'constant' centre padding,
//   frames cropped to int(L/hop)), _get_logmel_spectrogram :155-162 (|X|^2 @ mel_bank,
//   power_to_db with top_db=80 per channel plane), _get_gcc :164-175 (R = conj(X_m) X_n,
//   irfft(exp(j*angle(R))), lags [-M/2, M/2)), assembled channel-first as preprocess.py:549-556.
//
// One warp per frame, warps independent.  Per frame:
//   1. both packed complex FFTs ((mic0,mic1), (mic2,mic3)) together in float2 halves (as iv2);
//   2. untangle -> the four spectra go to shared memory (4 x 513 complex), powers to 4 swizzled rows;
//   3. segment-walk mel of the 4 power rows -> dB, stored unclamped; the running per-(clip, mic)
//      maximum is kept in registers and flushed with one atomicMax per clip change;
//   4. GCC-PHAT: three complex inverse transforms for the six pairs.  A transform's input is
//      Z = ph_a + i*ph_b for two mic pairs (its real/imag outputs are the two correlations); lane l
//      builds Z[l + 32m] for all m from the shared unit phasors X_c/|X_c| (bins above 512 are the
//      conjugates of 1024-k), runs the 32-point inverse stage in registers, twiddles, exchanges, and -- because
//      only lags [-32, 32) are kept -- evaluates just the two needed outputs of the second stage.
//      Pass 0 does two transforms in the two float2 halves (pairs 01 + i 12 | 03 + i 23: with the phasors kept as
//      the pairs (u0, u2), (u1, u3) of the untangle step those four products are packed operations) and leaves
//      the products of pairs 02 and 13 where the phasors were; pass 1 transforms those as one internally packed
//      transform (fft32_dit) with scalar exchange planes.
// Order in the kernel: 1, 2, 4, 3 -- the GCC passes exchange through phasor rows 2 and 3 (free once pass 0 has its input), so the
// power rows of step 3 stay in the exchange area of step 1 until the mel step reads them (measured -1.2 % against mel first).
// A second, element-wise kernel applies the top_db floor once every frame's maximum is known.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "fft32.cuh"
#include "mel_seg.cuh"
#include "seld_plan.h"
#include "tmem_tables.cuh"

// the 64 twiddle floats of the lane, (cos, cos', -sin, -sin') per pair of positions: from tensor memory (two loads, then
// compile-time indexing) or from the shared-memory table
#ifdef SELD_TMEM_TABLES
#define TW_FETCH() uint32_t tvA[32], tvB[32]; tmem_ld32(tmem_w + 32, tvA); tmem_ld32(tmem_w + 64, tvB);
#define TW4(p2) make_float4(__uint_as_float(((p2) < 8 ? tvA : tvB)[4 * ((p2) & 7)]), __uint_as_float(((p2) < 8 ? tvA : tvB)[4 * ((p2) & 7) + 1]), \
                            __uint_as_float(((p2) < 8 ? tvA : tvB)[4 * ((p2) & 7) + 2]), __uint_as_float(((p2) < 8 ? tvA : tvB)[4 * ((p2) & 7) + 3]))
#else
#define TW_FETCH() do { } while (0)
#define TW4(p2) (*reinterpret_cast<const float4*>(tw_s + lane * kTwStride + 4 * (p2)))
#endif

namespace seld {
namespace mic {
using namespace melseg;

constexpr int kW = 8;                      // warps (frames) per block
constexpr int kSpecStride = 544;           // float2 per phasor row: 513 + pad to a whole number of 128-byte lines (with 516 the 256-byte warp
                                           // accesses of rows 1-3 straddled three lines instead of two); two rows hold a 32 x 34 exchange buffer
constexpr int kSpec = 4 * kSpecStride * 2; // floats: four spectra
constexpr int kXStride = 34;                // exchange buffer row stride in float2 (even: 128-bit reads)
constexpr int kRowsArea = 32 * kXStride * 2; // floats: 32x34 float2 exchange buffer; the 4 power rows (2112) alias it
constexpr int kTwStride = 68;               // lane-major twiddle table row: 32 float2 + pad
constexpr int kWinStride = 36;              // lane-major window table row: 32 floats + pad
constexpr int kItemRow = 544;               // float2 per power pair-row in natural bin order: (P0, P2) and (P1, P3) fill the exchange area exactly
constexpr int kItemSlots = 136, kItemZero = 128;   // piece sums: 4 arrays of kItemSlots float2 (4 classes x 32 lanes, the slot kept at zero, pad)
static_assert(2 * 2 * kItemRow <= kRowsArea && 4 * 2 * kItemSlots <= kRowsArea, "rows and piece sums live in the exchange area");
static_assert(2 * kSpecStride * 2 >= kRowsArea, "the GCC passes exchange through phasor rows 2 and 3");
constexpr int kRegion = kSpec + kRowsArea;
static_assert(kRegion % 32 == 0 && kSpec % 32 == 0 && (kSpecStride * 2) % 32 == 0, "per-warp regions and their parts start on 128-byte lines");
// imbalance rule as in seld_foa_iv2.cu: a loose ratio that three of eight neighbouring bands must cross, a strict one
// that a single band may cross; every microphone's phases enter three GCC planes, hence the loose ratio of 40 dB
constexpr float kTauLoose = 1e-4f, kTauStrict = 1e-7f;
constexpr uint32_t kRedoMark = 0x7fc5e1d0u;  // quiet NaN with a payload, in element 0 of the frame's first log-mel row

// order-preserving float <-> int key for atomicMax
__device__ __forceinline__ int f2key(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }

// X / |X| given |X|^2 (0 for a vanishing bin, so that products with it vanish too)
__device__ __forceinline__ float inv_mag(float p) { return p > 1e-37f ? rsqrt_ftz(p) : 0.0f; }

// conj(ua) * ub for unit (or zero) phasors; a vanishing product means angle(0) = 0 -> phasor 1
// kCheck = false: the caller knows that no channel of this frame has a vanishing bin, so no product vanishes
template <bool kCheck>
__device__ __forceinline__ float2 cross_phasor(float2 ua, float2 ub) {
    float re = fmaf(ua.y, ub.y, ua.x * ub.x);
    const float im = fmaf(-ua.y, ub.x, ua.x * ub.y);
    if constexpr (kCheck) {
        if (re == 0.0f && im == 0.0f) re = 1.0f;
    }
    return make_float2(re, im);
}
}  // namespace mic

// kMode 0: waveform -> features (the fused path).
// kMode 1: waveform -> complex spectrogram only (Features_Extractor_MIC._spectrogram, feature.py:146-153): a.spec receives
//          (B, T, 513, 4) complex64, i.e. the reference's (T, F, C) layout per clip.
// kMode 2: features from a given spectrogram in that layout (_get_logmel_spectrogram / _get_gcc, feature.py:155-175):
//          the forward transform is skipped, everything after it is the same code.
// kRedo (kMode 0 only): second launch that recomputes the frames the main launch marked as too unbalanced for the packed
//          transform -- a microphone whose mel bands lie far under those of the microphone it shares a transform with sees
//          that one's fp32 rounding noise, and PHAT turns noise into phase.  Those frames run two passes with each
//          microphone alone in its transform (partner slot zero): exact down to a digitally silent microphone, whose
//          cross-spectra then vanish exactly (angle(0) = 0: phasor 1, a unit pulse at lag 0).
template <int kMode, bool kRedo = false>
__global__ void __launch_bounds__(mic::kW * 32, 1)
mic_features_kernel(const FoaArgs a, const PlanDev pd, int* __restrict__ maxkey) {
    using namespace mic;
    constexpr int W = kW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // lane-major tables read with 128-bit loads (row stride = 4 mod 32 words: conflict-free)
    float* tw_s = reinterpret_cast<float*>(smem_raw);                      // [lane][kTwStride]: W1024^(lane*brev5(p)) as (cos_p, cos_p+1, -sin_p, -sin_p+1), p even
    float* win_s = tw_s + 32 * kTwStride;                                  // [lane][kWinStride]: window[32*m + lane] * 0.5
    float* wab_s = win_s + 32 * kWinStride;                                // item form of the mel bank: iw, iP * 32 float2 (a, b)
#ifdef SELD_TMEM_TABLES
    float* R_all = reinterpret_cast<float*>(smem_raw);                     // tables in tensor memory: shared memory holds the per-warp regions only
    constexpr bool kSmemTables = false;
#else
    float* R_all = wab_s + 64 * pd.iP;                                     // W * kRegion
    constexpr bool kSmemTables = true;
#endif
    int* marked_s = reinterpret_cast<int*>(R_all + W * kRegion);           // main form: some warp of this block marked a frame

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if constexpr (kSmemTables)
    for (int i = tid; i < 1024; i += W * 32) {
        const int l = i & 31, r = i >> 5;                                  // pd.tw is [ka][lane], pd.win is [32*m + lane]
        const float2 w = pd.tw[brev5(r) * 32 + l];                         // positions (2j, 2j+1) share one float4: (cos, cos', -sin, -sin')
        tw_s[l * kTwStride + 4 * (r >> 1) + (r & 1)] = w.x; tw_s[l * kTwStride + 4 * (r >> 1) + 2 + (r & 1)] = w.y;
        win_s[l * kWinStride + r] = pd.win[i];
    }
    if constexpr (kSmemTables)
        for (int i = tid; i < 64 * pd.iP; i += W * 32) wab_s[i] = reinterpret_cast<const float*>(pd.iw)[i];
    if (tid == 0) *marked_s = 0;
    // the lane-private tables also go to tensor memory (tmem_tables.cuh): window [0, 32), twiddles in this kernel's layout
    // [32, 96), mel weights [96, 96 + 2 iP); the loop reads them from there, off the shared-memory pipe
    uint32_t tmem_w = 0;
#ifdef SELD_TMEM_TABLES
    tmem_w = tmem_tables_alloc(reinterpret_cast<uint32_t*>(marked_s + 1), warp);
    if (warp < 4) {
        uint32_t v[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) v[m] = __float_as_uint(pd.win[32 * m + lane]);
        tmem_st32(tmem_w, v);
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                       // positions 16 h .. 16 h + 15 as (cos, cos', -sin, -sin') per pair
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float2 w = pd.tw[brev5(16 * h + r) * 32 + lane];
                v[4 * (r >> 1) + (r & 1)] = __float_as_uint(w.x); v[4 * (r >> 1) + 2 + (r & 1)] = __float_as_uint(w.y);
            }
            tmem_st32(tmem_w + 32 + 32 * h, v);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int pos = 16 * h + r;
                const float2 w = pos < pd.iP ? pd.iw[pos * 32 + lane] : make_float2(0.0f, 0.0f);
                v[2 * r] = __float_as_uint(w.x); v[2 * r + 1] = __float_as_uint(w.y);
            }
            tmem_st32(tmem_w + 96 + 32 * h, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
    __syncthreads();
#ifdef SELD_TMEM_TABLES
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif

    float2* spec = reinterpret_cast<float2*>(R_all + warp * kRegion);      // [4][kSpecStride]
    float* R = R_all + warp * kRegion + kSpec;                             // 4 rows / exchange buffer
    float2* scratch = reinterpret_cast<float2*>(R);
    const int hop = pd.hop, M = pd.n_mels;
    const float amin = pd.amin;
    const int64_t ch_stride = (int64_t)a.T * M;
    const int src = (32 - lane) & 31;
    // mel step in its item form (see seld_foa_iv2.cu): lane l walks one piece of a segment per class, from bin ist[c] on;
    // bands lane and lane + 32 then add up at most four piece sums per list (segment m: V sums, segment m + 1: U sums)
    int ist[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ist[c] = pd.istart[c * 32 + lane];
    const uint32_t slotV0 = pd.islot[2 * lane], slotU0 = pd.islot[2 * lane + 1];
    const uint32_t slotV1 = pd.islot[2 * (lane + 32)], slotU1 = pd.islot[2 * (lane + 32) + 1];
    const bool four = __any_sync(0xffffffffu, (slotV0 >> 24) != kItemZero || (slotU0 >> 24) != kItemZero ||
                                              (slotV1 >> 24) != kItemZero || (slotU1 >> 24) != kItemZero);

    int cur_b = -1;
    float rmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float rmin[4] = {INFINITY, INFINITY, INFINITY, INFINITY};               // ... and minima: a plane whose minimum is above its floor is left alone
    auto flush_max = [&]() {
        if (kMode == 1 || cur_b < 0) return;                               // spectrogram mode: no dB planes, no maxima
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = rmax[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) atomicMax(maxkey + cur_b * 4 + c, f2key(v));
            rmax[c] = -INFINITY;
            float u = rmin[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) u = fminf(u, __shfl_xor_sync(0xffffffffu, u, o));
            if (lane == 0) atomicMin(maxkey + (a.B + cur_b) * 4 + c, f2key(u));
            rmin[c] = INFINITY;
        }
    };

    // frames of this warp: the main form walks its tiles; the redo form looks at 32 frames at a time (one per lane)
    // and then takes the marked ones in turn
    // a block owns a contiguous range of tiles (neighbouring tiles share samples in L1; -3 % against grid-stride tiles in
    // the FOA kernel, -2 % here); block i of the redo grid looks through the tiles of main block i
    const int nblk = kRedo ? a.redo_grid : (int)gridDim.x;
    const int tile_lo = (int)(((int64_t)blockIdx.x * a.n_tiles) / nblk), tile_hi = (int)(((int64_t)(blockIdx.x + 1) * a.n_tiles) / nblk);
    int tile = tile_lo - 1;
    int tb = tile_lo / a.tiles_per_clip, tr = tile_lo - tb * a.tiles_per_clip - 1;   // (clip, tile within the clip), advanced without a division per frame
    // redo form: item j = (tile number j / W of that block, warp slot j % W)
    const int redo_tiles = kRedo ? tile_hi - tile_lo : 0;
    const int redo_items = redo_tiles * W;
    auto redo_item = [&](int j, int& b, int& t) -> bool {
        const int tl = tile_lo + j / W;
        b = tl / a.tiles_per_clip;
        t = (tl - b * a.tiles_per_clip) * W + (j % W);
        return j < redo_items && t < a.T;
    };
    int scan_j0 = warp * 32 - W * 32;
    uint32_t todo = 0u;
    for (;;) {
        int b, t;
        if constexpr (!kRedo) {
            ++tile;
            if (tile >= tile_hi) break;
            if (++tr >= a.tiles_per_clip) { tr = 0; ++tb; }
            b = tb;
            t = tr * W + warp;
            if (t >= a.T) continue;
        } else {
            while (todo == 0u) {
                scan_j0 += W * 32;
                if (scan_j0 >= redo_items) break;
                int gb, gt;
                bool marked = false;
                if (redo_item(scan_j0 + lane, gb, gt))
                    marked = __float_as_uint(a.out[(((int64_t)gb * a.Cout) * a.T + gt) * M]) == kRedoMark;
                todo = __ballot_sync(0xffffffffu, marked);
            }
            if (todo == 0u) break;
            redo_item(scan_j0 + (__ffs(todo) - 1), b, t);
            todo &= todo - 1;
        }
        if (b != cur_b) { flush_max(); cur_b = b; }
        const float* xb = static_cast<const float*>(a.x) + (int64_t)b * a.stride_b;
        const int64_t s0 = (int64_t)t * hop - 512;

        float2 re[32], im[32];
        float* const spec_g = static_cast<float*>(a.spec) + (((int64_t)b * a.T + t) * 513) * 8;   // kMode 1 / 2: this frame's (513, 4) complex block
        float2 keep02[17];                                                  // kRedo: (|X0|^2, |X2|^2) of pass 0, written as rows after pass 1's exchange
        float min_n = 1.0f;                                                 // becomes 0 if any channel has a vanishing bin
#pragma unroll 1
        for (int pass = 0; pass < (kRedo ? 2 : 1); ++pass) {
        if constexpr (kMode != 2) {
        // ---------------- load + window: re = (mic0, mic2), im = (mic1, mic3); zeros outside the clip
        if constexpr (kRedo) {                                              // pass p: re = (mic p, mic p+2), im = 0
            const float* pa = xb + (int64_t)pass * a.stride_c;
            static_for<0, 32>([&](auto mi) {
                constexpr int m = decltype(mi)::value;
                const int64_t s = s0 + 32 * m + lane;
                const bool in = s >= 0 && s < a.L;
                const float* p = pa + (in ? s : 0);
                re[m] = in ? make_float2(__ldg(p), __ldg(p + 2 * a.stride_c)) : make_float2(0.f, 0.f);
                im[m] = make_float2(0.f, 0.f);
            });
        } else
        if (s0 >= 0 && s0 + 1024 <= a.L) {
            const float* p0 = xb + s0 + lane;
            const float* p1 = p0 + a.stride_c;
            const float* p2 = p1 + a.stride_c;
            const float* p3 = p2 + a.stride_c;
            static_for<0, 32>([&](auto mi) {
                constexpr int m = decltype(mi)::value;
                re[m] = make_float2(__ldg(p0 + 32 * m), __ldg(p2 + 32 * m));
                im[m] = make_float2(__ldg(p1 + 32 * m), __ldg(p3 + 32 * m));
            });
        } else {
            static_for<0, 32>([&](auto mi) {
                constexpr int m = decltype(mi)::value;
                const int64_t s = s0 + 32 * m + lane;
                const bool in = s >= 0 && s < a.L;
                const float* p = xb + (in ? s : 0);
                re[m] = in ? make_float2(__ldg(p), __ldg(p + 2 * a.stride_c)) : make_float2(0.f, 0.f);
                im[m] = in ? make_float2(__ldg(p + a.stride_c), __ldg(p + 3 * a.stride_c)) : make_float2(0.f, 0.f);
            });
        }
#ifdef SELD_TMEM_TABLES
        {
            uint32_t wv[32];
            tmem_ld32(tmem_w, wv);
            static_for<0, 32>([&](auto mi) {
                constexpr int m = decltype(mi)::value;
                re[m] = vmuls(re[m], __uint_as_float(wv[m]));
                im[m] = vmuls(im[m], __uint_as_float(wv[m]));
            });
        }
#else
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
#endif

        // ---------------- forward: two 1024-point FFTs at once
        fft32(re, im);
        TW_FETCH();
        static_for<0, 16>([&](auto pi) {
            constexpr int p2 = decltype(pi)::value;                         // positions 2*p2, 2*p2+1
            const float4 w4 = TW4(p2);
            if constexpr (p2 > 0) {
                const float2 r = re[2 * p2], i = im[2 * p2];
                re[2 * p2] = vfmas(i, -w4.z, vmuls(r, w4.x));
                im[2 * p2] = vfmas(i, w4.x, vmuls(r, w4.z));
            }
            const float2 r = re[2 * p2 + 1], i = im[2 * p2 + 1];
            re[2 * p2 + 1] = vfmas(i, -w4.w, vmuls(r, w4.y));
            im[2 * p2 + 1] = vfmas(i, w4.y, vmuls(r, w4.w));
        });
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
        fft32(re, im);
        }  // kMode != 2

        // ---------------- untangle: spectra -> spec[c][k], powers -> rows
        static_for<0, 17>([&](auto kbi) {
            constexpr int kb = decltype(kbi)::value;
            constexpr int p = brev5(kb & 31);
            float2 ar, ai, br, bi;                                          // (X0, X2) and (X1, X3), real and imaginary parts
            if constexpr (kMode == 2) {
                float4 c01 = make_float4(0.f, 0.f, 0.f, 0.f), c23 = c01;
                if (kb < 16 || lane == 0) {
                    const float4* g4 = reinterpret_cast<const float4*>(spec_g + (int64_t)(lane + 32 * kb) * 8);
                    c01 = __ldg(g4); c23 = __ldg(g4 + 1);
                }
                ar = make_float2(c01.x, c23.x); ai = make_float2(c01.y, c23.y);
                br = make_float2(c01.z, c23.z); bi = make_float2(c01.w, c23.w);
            } else {
                const float2 zr = re[p], zi = im[p];
                float2 pr, pi;
                if constexpr (kb == 16) {
                    pr = zr; pi = zi;
                } else {
                    constexpr int pp = brev5(31 - kb), p0 = brev5((32 - kb) & 31);
#ifdef MABL_NOSHFL
                    pr = re[pp]; pi = im[pp];
#else
                    pr.x = __shfl_sync(0xffffffffu, re[pp].x, src);
                    pr.y = __shfl_sync(0xffffffffu, re[pp].y, src);
                    pi.x = __shfl_sync(0xffffffffu, im[pp].x, src);
                    pi.y = __shfl_sync(0xffffffffu, im[pp].y, src);
                    if (lane == 0) { pr = re[p0]; pi = im[p0]; }
#endif
                }
                ar = vadd(zr, pr); ai = vsub(zi, pi);
                br = vadd(zi, pi); bi = vsub(pr, zr);
            }
            if constexpr (kMode == 1) {
                if (kb < 16 || lane == 0) {
                    float4* g4 = reinterpret_cast<float4*>(spec_g + (int64_t)(lane + 32 * kb) * 8);
                    g4[0] = make_float4(ar.x, ai.x, br.x, bi.x);
                    g4[1] = make_float4(ar.y, ai.y, br.y, bi.y);
                }
                return;
            }
            const float2 p02 = __ffma2_rn(ai, ai, __fmul2_rn(ar, ar));
            const float2 p13 = __ffma2_rn(bi, bi, __fmul2_rn(br, br));
            if constexpr (kRedo) {
                // the live slots are (mic pass, mic pass + 2) = (ar, ai); the other two are zero
                if (kb < 16 || lane == 0) {
                    const int k = lane + 32 * kb;
                    const float2 n02 = make_float2(inv_mag(p02.x), inv_mag(p02.y));
                    min_n = fminf(min_n, fminf(n02.x, n02.y));
                    const float2 ur = __fmul2_rn(ar, n02), ui = __fmul2_rn(ai, n02);
                    spec[(2 * pass) * kSpecStride + k] = ur;               // pass 0: (u0, u2), pass 1: (u1, u3), real and imaginary parts
                    spec[(2 * pass + 1) * kSpecStride + k] = ui;
                    if (pass == 1) {                                        // pair rows (P0, P2), (P1, P3) in natural bin order
                        reinterpret_cast<float2*>(R)[k] = keep02[kb];
                        reinterpret_cast<float2*>(R)[kItemRow + k] = p02;
                    }
                }
                if (pass == 0) keep02[kb] = p02;
                return;
            }
            if (kb < 16 || lane == 0) {
                const int k = lane + 32 * kb;
                // PHAT only needs phases: keep X_c / |X_c| (unit phasors) for the GCC passes
#ifdef MABL_NOINVMAG
                const float2 n02 = p02, n13 = p13;
#else
                const float2 n02 = make_float2(inv_mag(p02.x), inv_mag(p02.y)), n13 = make_float2(inv_mag(p13.x), inv_mag(p13.y));
                min_n = fminf(min_n, fminf(fminf(n02.x, n02.y), fminf(n13.x, n13.y)));
#endif
                const float2 ur02 = __fmul2_rn(ar, n02), ui02 = __fmul2_rn(ai, n02);
                const float2 ur13 = __fmul2_rn(br, n13), ui13 = __fmul2_rn(bi, n13);
#ifdef MABL_NOSPECST
                if (__float_as_uint(ur02.x + ui02.x + ur13.x + ui13.x + ur02.y + ui02.y + ur13.y + ui13.y) == 0x12345678u)
#endif
                {
                spec[0 * kSpecStride + k] = ur02;                           // pairs as they come out of the untangle step:
                spec[1 * kSpecStride + k] = ui02;                           // (Re u0, Re u2), (Im u0, Im u2), (Re u1, Re u3), (Im u1, Im u3)
                spec[2 * kSpecStride + k] = ur13;
                spec[3 * kSpecStride + k] = ui13;
                }
                reinterpret_cast<float2*>(R)[k] = p02;                      // pair rows (P0, P2), (P1, P3) in natural bin order
                reinterpret_cast<float2*>(R)[kItemRow + k] = p13;
            }
        });
        __syncwarp();
        }  // pass
        if constexpr (kMode == 1) continue;

        float* ob = a.out + (((int64_t)b * a.Cout) * a.T + t) * M;
        // (the GCC passes come before the mel step: they exchange through phasor rows 2 and 3, the power rows stay where they are)
        // ---------------- GCC-PHAT.  Two real correlations per complex inverse transform (Z = ph_a + i ph_b):
        // pass 0 = pairs (01, 02 | 03, 12) as two transforms packed in float2 halves, pass 1 = pairs (13, 23) as
        // one transform packed internally (fft32_dit).
        constexpr float kInvN = 1.0f / 1024.0f;
        float2* const xg = spec + 2 * kSpecStride;                          // exchange buffer of both passes: phasor rows 2 and 3, free once pass 0 has its input
        // angle(0) = 0: a vanishing cross-spectrum bin must contribute the phasor 1.  That needs two compares and a
        // select per product; frames without any vanishing bin (all but digital silence / dead channels) skip them.
        const bool any_zero = __any_sync(0xffffffffu, min_n == 0.0f);
#ifndef MABL_NOGCC0
        {
            // phasors kept as the pairs (u0, u2) and (u1, u3): conj(u0) (u1, u3) and conj(u2) (u1, u3) are four packed operations
            // each and give the pairs 01, 03, 21, 23; transforms: x halves Z = P01 + i P12 (P12 = conj(P21)), y halves Z = P03 + i P23.
            // The two products left, P02 and P13, are taken here as well, from the same four loads, for the bins a lane owns
            // (k <= 512), and written, already combined into pass 1's input, over the phasors of those bins (rows 0 and 1: the
            // mirrored reads of all lanes come first), so pass 1 loads one value per input and computes nothing.
            auto build = [&](auto check_c) {
                constexpr bool kCheck = decltype(check_c)::value;
                auto one = [&](auto mi) {
                    constexpr int m = decltype(mi)::value;
                    const int k = lane + 32 * m;
                    const bool up = k > 512;
                    const int kk = up ? 1024 - k : k;
                    const float sg = up ? -1.0f : 1.0f;
                    // bins above 512 are the conjugates of bins 1024 - k: with every phasor conjugated every product is, too
                    // (m = 16 is the one row where that depends on the lane)
                    const float2 r02 = spec[0 * kSpecStride + kk], r13 = spec[2 * kSpecStride + kk];
                    float2 i02 = spec[1 * kSpecStride + kk], i13 = spec[3 * kSpecStride + kk];
                    if constexpr (m == 16) { i02 = vmuls(i02, sg); i13 = vmuls(i13, sg); }
                    float2 Ar = vfmas(i13, i02.x, vmuls(r13, r02.x));         // Re (P01, P03)
                    float2 Ai = vfmas(r13, -i02.x, vmuls(i13, r02.x));        // Im (P01, P03)
                    float2 Br = vfmas(i13, i02.y, vmuls(r13, r02.y));         // Re (P21, P23)
                    float2 Bi = vfmas(r13, -i02.y, vmuls(i13, r02.y));        // Im (P21, P23)
                    if constexpr (kCheck) {
                        if (Ar.x == 0.0f && Ai.x == 0.0f) Ar.x = 1.0f;
                        if (Ar.y == 0.0f && Ai.y == 0.0f) Ar.y = 1.0f;
                        if (Br.x == 0.0f && Bi.x == 0.0f) Br.x = 1.0f;
                        if (Br.y == 0.0f && Bi.y == 0.0f) Br.y = 1.0f;
                    }
                    if constexpr (m > 16) {
                        re[m] = __ffma2_rn(Bi, make_float2(-1.0f, 1.0f), Ar);
                        im[m] = vsub(Br, Ai);
                    } else {
                        re[m] = __ffma2_rn(Bi, make_float2(1.0f, -1.0f), Ar); // Re P01 - Im P12 | Re P03 - Im P23
                        im[m] = vadd(Ai, Br);                                 // Im P01 + Re P12 | Im P03 + Re P23
                    }
                    if constexpr (m <= 16) {
                        if (m < 16 || lane == 0) {
                            const float2 p02 = cross_phasor<kCheck>(make_float2(r02.x, i02.x), make_float2(r02.y, i02.y));
                            const float2 p13 = cross_phasor<kCheck>(make_float2(r13.x, i13.x), make_float2(r13.y, i13.y));
                            // pass 1 transforms Z = P02 + i P13: row 0 takes Z[k] itself, row 1 what the reader of the mirrored
                            // bin needs, conj(P02) + i conj(P13) = Z[1024 - k]: one load per input there
                            spec[0 * kSpecStride + k] = make_float2(p02.x - p13.y, p02.y + p13.x);
                            spec[1 * kSpecStride + k] = make_float2(p02.x + p13.y, p13.x - p02.y);
                        }
                    }
                };
                static_for<16, 32>(one);                                      // mirrored reads (m = 16: lane 0 owns bin 512, which nobody else reads)
                __syncwarp();                                                 // every mirrored read is done: a lane's own bins may be overwritten
                static_for<0, 16>(one);
                __syncwarp();
            };
            if (any_zero) build(std::true_type{}); else build(std::false_type{});
            // inverse 32-point stage over m: swap(FFT(swap(z)))
            fft32(im, re);                                                  // position p: A[lane][n2 = brev5(p)]
            TW_FETCH();
            static_for<0, 16>([&](auto pi) {                                // table holds (cos, -sin): multiply by (cos + i sin)
                constexpr int p2 = decltype(pi)::value;
                const float4 w4 = TW4(p2);
                if constexpr (p2 > 0) {
                    const float2 r = re[2 * p2], i = im[2 * p2];
                    re[2 * p2] = vfmas(i, w4.z, vmuls(r, w4.x));
                    im[2 * p2] = vfmas(i, w4.x, vmuls(r, -w4.z));
                }
                const float2 r = re[2 * p2 + 1], i = im[2 * p2 + 1];
                re[2 * p2 + 1] = vfmas(i, w4.w, vmuls(r, w4.y));
                im[2 * p2 + 1] = vfmas(i, w4.y, vmuls(r, -w4.w));
            });
            static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; xg[brev5(p) * kXStride + lane] = re[p]; });
            __syncwarp();
            static_for<0, 16>([&](auto ji) {
                constexpr int j = decltype(ji)::value;
                const float4 v = *reinterpret_cast<const float4*>(xg + lane * kXStride + 2 * j);
                re[2 * j] = make_float2(v.x, v.y); re[2 * j + 1] = make_float2(v.z, v.w);
            });
            __syncwarp();
            static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; xg[brev5(p) * kXStride + lane] = im[p]; });
            __syncwarp();
            static_for<0, 16>([&](auto ji) {
                constexpr int j = decltype(ji)::value;
                const float4 v = *reinterpret_cast<const float4*>(xg + lane * kXStride + 2 * j);
                im[2 * j] = make_float2(v.x, v.y); im[2 * j + 1] = make_float2(v.z, v.w);
            });
            __syncwarp();
            // second stage, only n1 = 0 (lag n2 = lane) and n1 = 31 (lag lane - 32):
            //   c0 = sum_k1 A'[k1],  c31 = sum_k1 A'[k1] * (cos(2 pi k1/32) - i sin(2 pi k1/32))
            float2 c0r = re[0], c0i = im[0], c31r = re[0], c31i = im[0];
            // terms k1 and 32 - k1 share the cosine and have opposite sines: sums and differences first
            static_for<1, 16>([&](auto ki) {
                constexpr int k1 = decltype(ki)::value;
                constexpr float c = (float)cos32(k1), sn = (float)sin32(k1);
                const float2 sr = vadd(re[k1], re[32 - k1]), si = vadd(im[k1], im[32 - k1]);
                const float2 dr = vsub(re[k1], re[32 - k1]), di = vsub(im[k1], im[32 - k1]);
                c0r = vadd(c0r, sr); c0i = vadd(c0i, si);
                c31r = vfmas(di, sn, vfmas(sr, c, c31r));
                c31i = vfmas(dr, -sn, vfmas(si, c, c31i));
            });
            c0r = vadd(c0r, re[16]); c0i = vadd(c0i, im[16]);
            c31r = vsub(c31r, re[16]); c31i = vsub(c31i, im[16]);
            // real part = first pair of a transform, imaginary part = second pair
            float* g = ob + (int64_t)4 * ch_stride;
            g[0 * ch_stride + lane] = c31r.x * kInvN;  g[0 * ch_stride + 32 + lane] = c0r.x * kInvN;   // pair 01
            g[3 * ch_stride + lane] = c31i.x * kInvN;  g[3 * ch_stride + 32 + lane] = c0i.x * kInvN;   // pair 12
            g[2 * ch_stride + lane] = c31r.y * kInvN;  g[2 * ch_stride + 32 + lane] = c0r.y * kInvN;   // pair 03
            g[5 * ch_stride + lane] = c31i.y * kInvN;  g[5 * ch_stride + 32 + lane] = c0i.y * kInvN;   // pair 23
        }
#endif
#ifndef MABL_NOGCC1
        {
            // one transform: position p of (zr, zi) holds the input pair (z[m = 2p], z[m = 2p + 1])
            float2 zr[16], zi[16];
            {
                static_for<0, 16>([&](auto pi) {
                    constexpr int p = decltype(pi)::value;
                    float rr[2], ii[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = lane + 32 * (2 * p + e);
                        const bool up = k > 512;
                        const int kk = up ? 1024 - k : k;
                        const float2 z = spec[(up ? 1 : 0) * kSpecStride + kk];   // Z = P02 + i P13 as pass 0 left it (row 1: for the mirrored bin)
                        rr[e] = z.x; ii[e] = z.y;
                    }
                    zr[p] = make_float2(rr[0], rr[1]); zi[p] = make_float2(ii[0], ii[1]);
                });
            }
            fft32_dit(zi, zr);                                              // position q': (A[lane][q], A[lane][q + 16]), q = brev4(q')
            float* fr = reinterpret_cast<float*>(xg);                       // two planes of 32 x 34 floats in the exchange area
            float* fi = fr + 32 * kXStride;
            TW_FETCH();
            static_for<0, 16>([&](auto qi) {
                constexpr int qp = decltype(qi)::value;
                constexpr int q = brev4(qp);
                // table positions brev5(q) (even) and brev5(q + 16) = brev5(q) + 1 share one float4: (cos, cos', -sin, -sin')
                const float4 w4 = TW4(brev5(q) / 2);
                float2 r = zr[qp], i = zi[qp];
                if constexpr (q == 0) {                                     // W^0 = 1 for n2 = 0; n2 = 16 still needs its factor
                    const float r1 = r.y, i1 = i.y;
                    r.y = fmaf(i1, w4.w, r1 * w4.y);
                    i.y = fmaf(i1, w4.y, r1 * -w4.w);
                } else {
                    const float2 C2 = make_float2(w4.x, w4.y), S2 = make_float2(w4.z, w4.w);
                    const float2 t = __fmul2_rn(make_float2(-r.x, -r.y), S2);
                    r = __ffma2_rn(i, S2, __fmul2_rn(r, C2));
                    i = __ffma2_rn(i, C2, t);
                }
                fr[q * kXStride + lane] = r.x; fr[(q + 16) * kXStride + lane] = r.y;
                fi[q * kXStride + lane] = i.x; fi[(q + 16) * kXStride + lane] = i.y;
            });
            __syncwarp();
            // second stage for n1 = 0 and 31, two k1 terms per packed operation
            float2 a0r = make_float2(0.f, 0.f), a0i = a0r, a31r = a0r, a31i = a0r;
            static_for<0, 16>([&](auto ji) {
                constexpr int j = decltype(ji)::value;
                const float2 pr = *reinterpret_cast<const float2*>(fr + lane * kXStride + 2 * j);   // (A'[2j], A'[2j + 1])
                const float2 pi = *reinterpret_cast<const float2*>(fi + lane * kXStride + 2 * j);
                constexpr int k0 = 2 * j, k1 = 2 * j + 1;
                constexpr float c0 = (float)(k0 <= 16 ? cos32(k0) : cos32(32 - k0)), c1 = (float)(k1 <= 16 ? cos32(k1) : cos32(32 - k1));
                constexpr float s0 = (float)(k0 <= 16 ? sin32(k0) : -sin32(32 - k0)), s1 = (float)(k1 <= 16 ? sin32(k1) : -sin32(32 - k1));
                a0r = __fadd2_rn(a0r, pr); a0i = __fadd2_rn(a0i, pi);
                a31r = __ffma2_rn(pi, make_float2(s0, s1), __ffma2_rn(pr, make_float2(c0, c1), a31r));
                a31i = __ffma2_rn(pr, make_float2(-s0, -s1), __ffma2_rn(pi, make_float2(c0, c1), a31i));
            });
            __syncwarp();                                                   // the next frame reuses the exchange area
            float* g = ob + (int64_t)5 * ch_stride;                         // pairs 02 and 13: planes 5 and 8
            g[0 * ch_stride + lane] = (a31r.x + a31r.y) * kInvN;  g[0 * ch_stride + 32 + lane] = (a0r.x + a0r.y) * kInvN;
            g[3 * ch_stride + lane] = (a31i.x + a31i.y) * kInvN;  g[3 * ch_stride + 32 + lane] = (a0i.x + a0i.y) * kInvN;
        }
#endif

        // ---------------- log-mel of the four power rows (unclamped dB + running maximum)
#ifndef MABL_NOMEL
        {
            float2 aU[4][2], aV[4][2];                                      // piece sums per class, both pair rows
            {
                const float2* const iw_s = reinterpret_cast<const float2*>(wab_s);
                const float2* const Q = reinterpret_cast<const float2*>(R);
                static_for<0, 4>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    aU[c][0] = aU[c][1] = aV[c][0] = aV[c][1] = make_float2(0.f, 0.f);
                    const int Lc = pd.iL[c];
                    const float2* qp = Q + ist[c];
                    const float2* wp = iw_s + 32 * pd.ioff[c] + lane;
#pragma unroll 1
                    for (int j0 = 0; j0 < Lc; j0 += 4) {                    // class lengths are multiples of four
#ifdef SELD_TMEM_TABLES
                    uint32_t w8[8];
                    tmem_ld8(tmem_w + 96 + 2 * (pd.ioff[c] + j0), w8);
#endif
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj, ++qp, wp += 32) {
                        const float2 q0 = qp[0], q1 = qp[kItemRow];
#ifdef SELD_TMEM_TABLES
                        const float2 w = make_float2(__uint_as_float(w8[2 * jj]), __uint_as_float(w8[2 * jj + 1]));
#else
                        const float2 w = *wp;
#endif
                        const float2 aa = make_float2(w.x, w.x), bb = make_float2(w.y, w.y);
                        aU[c][0] = __ffma2_rn(aa, q0, aU[c][0]); aV[c][0] = __ffma2_rn(bb, q0, aV[c][0]);
                        aU[c][1] = __ffma2_rn(aa, q1, aU[c][1]); aV[c][1] = __ffma2_rn(bb, q1, aV[c][1]);
                    }
                    }
                });
            }
            __syncwarp();                                                   // every lane is through with the rows: the sums may overwrite them
            {
                float2* const S = reinterpret_cast<float2*>(R);             // [U row 0 | V row 0 | U row 1 | V row 1], kItemSlots float2 each
                static_for<0, 4>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    if (c < pd.iK) {
                        S[0 * kItemSlots + 32 * c + lane] = aU[c][0]; S[1 * kItemSlots + 32 * c + lane] = aV[c][0];
                        S[2 * kItemSlots + 32 * c + lane] = aU[c][1]; S[3 * kItemSlots + 32 * c + lane] = aV[c][1];
                    }
                });
                if (lane < 4) S[lane * kItemSlots + kItemZero] = make_float2(0.f, 0.f);
            }
            __syncwarp();
            // band per lane; run numbers of segment m (V) and m+1 (U) in fixed slots (absent -> the zero run), all
            // loads issued up front (n_mels == 64 on this path: bands lane and lane + 32)
            bool bad = false;                                               // some microphone too far under its transform partner
            float rmax0[4], rmin0[4];                                       // the extrema before this frame (a marked frame must not count)
#pragma unroll
            for (int f = 0; f < 4; ++f) { rmax0[f] = rmax[f]; rmin0[f] = rmin[f]; }
            auto unbalanced = [](const float (&v)[4], float t) {
                return v[0] < t * v[1] || v[1] < t * v[0] || v[2] < t * v[3] || v[3] < t * v[2];
            };
            auto clustered = [](uint32_t m) {                               // an aligned group of eight bands with >= 3 bits set
                uint32_t c = m - ((m >> 1) & 0x55555555u);
                c = (c & 0x33333333u) + ((c >> 2) & 0x33333333u);
                c = (c + (c >> 4)) & 0x0f0f0f0fu;
                return ((c + 0x05050505u) & 0x08080808u) != 0u;
            };
            auto combine = [&](auto four_c) {
                constexpr bool kFour = decltype(four_c)::value;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int m = lane + 32 * r;
                    const uint32_t pv = r ? slotV1 : slotV0, pu = r ? slotU1 : slotU0;
                    float v[4];
                    {
                        const float2* S = reinterpret_cast<const float2*>(R);
                        const int iv0 = pv & 0xff, iv1 = (pv >> 8) & 0xff, iv2 = (pv >> 16) & 0xff, iv3 = pv >> 24;
                        const int iu0 = pu & 0xff, iu1 = (pu >> 8) & 0xff, iu2 = (pu >> 16) & 0xff, iu3 = pu >> 24;
                        float2 vv[2];
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const float2* su = S + (2 * f) * kItemSlots;
                            const float2* sv = su + kItemSlots;
                            const float2 v0 = sv[iv0], v1 = sv[iv1], v2 = sv[iv2], u0 = su[iu0], u1 = su[iu1], u2 = su[iu2];
                            if constexpr (kFour) vv[f] = vadd(vadd(vadd(v0, v1), vadd(v2, sv[iv3])), vadd(vadd(u0, u1), vadd(u2, su[iu3])));
                            else vv[f] = vadd(vadd(vadd(v0, v1), v2), vadd(vadd(u0, u1), u2));
                        }
                        v[0] = vv[0].x; v[2] = vv[0].y; v[1] = vv[1].x; v[3] = vv[1].y;
                    }
                    if constexpr (kMode == 0 && !kRedo) {                   // every lane votes (no short-circuit in front of the ballot)
                        const uint32_t lm = __ballot_sync(0xffffffffu, unbalanced(v, kTauLoose));
                        bad = clustered(lm) || unbalanced(v, kTauStrict) || bad;
                    }
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const float db = 3.01029995663981195f * lg2_ftz(fmaxf(v[f], amin));
                        rmax[f] = fmaxf(rmax[f], db);
                        rmin[f] = fminf(rmin[f], db);
                        ob[f * ch_stride + m] = db;
                    }
                }
            };
            if (four) combine(std::true_type{}); else combine(std::false_type{});
            __syncwarp();                                                   // rows become the exchange buffer again
            if constexpr (kMode == 0 && !kRedo) {
                if (__any_sync(0xffffffffu, bad)) {
                    // leave the frame to the redo launch: mark it, take its values back out of the running maxima
                    if (lane == 0) { ob[0] = __uint_as_float(kRedoMark); *marked_s = 1; }   // lane 0 wrote that element itself: program order
#pragma unroll
                    for (int f = 0; f < 4; ++f) { rmax[f] = rmax0[f]; rmin[f] = rmin0[f]; }
                    continue;
                }
            }
        }
#endif

    }
    flush_max();
#ifdef SELD_TMEM_TABLES
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    tmem_tables_free(tmem_w, warp);
#endif
    if constexpr (kMode == 0 && !kRedo) {
        // a block that marked frames launches the redo form for them into the tail of this grid: it runs when the whole
        // grid is done and before the stream's next kernel (the top_db floor)
        __syncthreads();
        if (tid == 0) {                                                     // the last block to finish launches one redo grid (as in seld_foa_iv2.cu)
            if (*marked_s != 0) atomicExch(&a.redo_flags[1], 1);
            __threadfence();
            if (atomicAdd(&a.redo_flags[0], 1) == (int)gridDim.x - 1) {
                __threadfence();
                const int any = atomicExch(&a.redo_flags[1], 0);
                atomicExch(&a.redo_flags[0], 0);
#ifndef SELD_NO_DEVICE_LAUNCH
                if (any) {
                    FoaArgs ar = a;
                    ar.redo_grid = (int)gridDim.x;
                    mic_features_kernel<0, true><<<gridDim.x, W * 32, a.smem_bytes, cudaStreamTailLaunch>>>(ar, pd, maxkey);
                }
#else
                (void)any;
#endif
            }
        }
    }
}

// top_db floor of the log-mel planes: v = max(v, max_over_plane - top_db)
__global__ void mic_topdb_kernel(float* __restrict__ out, const int* __restrict__ maxkey, int B, int Cout,
                                 int64_t plane, float top_db) {
    const int bc = blockIdx.y;                                             // b * 4 + c
    const int b = bc >> 2, c = bc & 3;
    const int key = maxkey[bc], kmin = maxkey[4 * B + bc];
    const float mx = __int_as_float(key >= 0 ? key : key ^ 0x7fffffff);
    const float floor_db = mx - top_db;
    if (__int_as_float(kmin >= 0 ? kmin : kmin ^ 0x7fffffff) >= floor_db) return;   // nothing in this plane lies under the floor
    float4* p = reinterpret_cast<float4*>(out + ((int64_t)b * Cout + c) * plane);
    const int64_t n4 = plane >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = p[i];
        v.x = fmaxf(v.x, floor_db); v.y = fmaxf(v.y, floor_db); v.z = fmaxf(v.z, floor_db); v.w = fmaxf(v.w, floor_db);
        p[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
static size_t mic_smem_bytes(const PlanDev& pd) {
#ifdef SELD_TMEM_TABLES
    (void)pd;
    return (size_t)(mic::kW * mic::kRegion + 4) * sizeof(float);           // the tables live in tensor memory
#else
    return (size_t)(32 * mic::kTwStride + 32 * mic::kWinStride + 64 * pd.iP + mic::kW * mic::kRegion + 4) * sizeof(float);
#endif
}

bool mic_supported(const PlanDev& pd, size_t smem_optin) {
    return pd.item_ok && pd.n_mels == 64 && mic_smem_bytes(pd) <= smem_optin;
}

int mic_frames_per_tile() { return mic::kW; }

template <int kMode>
static cudaError_t mic_set_attr() {
    static std::atomic<uint64_t> done{0};                                   // per device, once per process and instantiation
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_relaxed) & bit)) {
        e = cudaFuncSetAttribute(mic_features_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        done.fetch_or(bit, std::memory_order_relaxed);
    }
    return cudaSuccess;
}

static cudaError_t mic_set_attr_redo() {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_relaxed) & bit)) {
        e = cudaFuncSetAttribute(mic_features_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        done.fetch_or(bit, std::memory_order_relaxed);
    }
    return cudaSuccess;
}

// waveform -> (B, T, 513, 4) complex64 spectrogram in a.spec
cudaError_t mic_spectrogram_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    cudaError_t e = mic_set_attr<1>();
    if (e != cudaSuccess) return e;
    const int gx = sm_count < a.n_tiles ? sm_count : a.n_tiles;
    mic_features_kernel<1><<<gx, mic::kW * 32, mic_smem_bytes(pd), st>>>(a, pd, nullptr);
    return cudaGetLastError();
}

// from_spectra: a.spec holds the (B, T, 513, 4) complex64 spectrogram, a.x is unused
cudaError_t mic_launch(const FoaArgs& a, const PlanDev& pd, int* maxkey, float top_db, bool use_top_db,
                       int sm_count, cudaStream_t st, bool from_spectra) {
    const size_t smem = mic_smem_bytes(pd);
    cudaError_t e = from_spectra ? mic_set_attr<2>() : mic_set_attr<0>();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(maxkey, 0x80, (size_t)a.B * 4 * sizeof(int), st);  // maxima: key 0x80808080, below any dB value
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(maxkey + (size_t)a.B * 4, 0x7f, (size_t)a.B * 4 * sizeof(int), st);   // minima: key 0x7f7f7f7f, above any
    if (e != cudaSuccess) return e;
    int gx = sm_count < a.n_tiles ? sm_count : a.n_tiles;
    if (from_spectra) {
        mic_features_kernel<2><<<gx, mic::kW * 32, smem, st>>>(a, pd, maxkey);
        e = cudaGetLastError();
    } else {
        e = mic_set_attr_redo();                                          // the device-side launch of the redo form needs its opt-in too
        if (e != cudaSuccess) return e;
        FoaArgs aa = a;
        aa.smem_bytes = (int)smem;
        mic_features_kernel<0><<<gx, mic::kW * 32, smem, st>>>(aa, pd, maxkey);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess || !use_top_db) return e;
    const int64_t plane = (int64_t)a.T * pd.n_mels;                         // multiple of 4 (n_mels = 64)
    dim3 grid((unsigned)((plane / 4 + 255) / 256 > 64 ? 64 : (plane / 4 + 255) / 256), (unsigned)(a.B * 4));
    mic_topdb_kernel<<<grid, 256, 0, st>>>(a.out, maxkey, a.B, a.Cout, plane, top_db);
    return cudaGetLastError();
}

}  // namespace seld
