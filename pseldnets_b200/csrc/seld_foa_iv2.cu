// FOA log-mel + intensity-vector kernel, second generation ("iv2"): the headline path.
//
//   LogmelIV_Extractor.forward (feature.py:39-56) + intensityvector (feature.py:93-117) for the
//   4-channel case, one warp per frame, warps fully independent (no block barrier in the loop).
//
// What changed against the first kernel (seld_foa.cu, kept as the general fallback):
//   * both packed complex FFTs of a frame ((ch0,ch1) and (ch2,ch3)) run TOGETHER in the two
//     halves of float2 registers: every butterfly is one FADD2/FMUL2/FFMA2 (sm_100a packed
//     fp32), and window / twiddle loads are shared -> about half the issue slots;
//   * samples are read straight from global memory with coalesced 128-byte warp loads (the
//     1024-hop overlap of neighbouring frames is served by L1/L2) -> no staging buffer, no
//     __syncthreads, no exposed staging latency;
//   * the mel step.  For a triangular bank every bin feeds exactly two bands: the bins between two consecutive band
//     centres (a "segment" s) feed bands s-1 and s, so out[m] = V[m] + U[m+1] with U / V the a- / b-weighted sums of a
//     segment.  Two forms of it live in this file:
//       - ITEM form (kItem = true; main form of both modes, round 2): the packed spectra go to shared memory as they leave
//         the transform (natural bin order, bin 1024 - k beside bin k), and every lane walks one PIECE of a segment per
//         class, doing the untangle / power / IV arithmetic of a bin and the seven weighted accumulates in one go -- one
//         instruction stream for all lanes, no run boundaries, no shuffles (see the comment at the kernel);
//       - RUN form (kItem = false; round 1; the redo form and banks the item plan does not fit): the seven per-bin
//         quantities are written to swizzled rows, lane c owns bins [16c, 16c+16) and accumulates per run of one segment,
//         with predicated partial-sum stores at the run boundaries.
//     Both end in a band-per-lane combine step that adds up at most four partial sums per list.
//   * the lane-private tables (window, twiddles, mel weights) of the item form live in tensor memory (tmem_tables.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "fft32.cuh"
#include "seld_plan.h"
#include "tmem_tables.cuh"

#ifndef SELD_STRIDED_TILES
#define SELD_CONTIG 1      // a block owns a contiguous range of tiles (measured -3 % against grid-stride tiles: neighbouring tiles share samples in L1)
#endif

namespace seld {

#ifdef SELD_PHASE_TIMING
// Developer build: per-warp clock64 time of each phase, summed over the warp's frames.
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(i) do { const long long _t = clock64(); if (lane == 0) phase_acc[i] += _t - phase_t0; phase_t0 = _t; } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

// Rows are kept in PAIRS, interleaved as float2 per bin, the way the untangle step produces them:
//   pair 0 = (P0, P2), pair 1 = (P1, P3), pair 2 = (n1, n3), pair 3 = (n2, -)       (log-mel only: pairs 0 and 1)
// so the mel walk and the combine step handle two rows per packed instruction.
constexpr int kPairWords = 1056;          // 33 chunks of 16 bins x float2 (bin 512 opens chunk 32)
constexpr int kPairs = 4;
constexpr int kRows = 7;                  // P0 P1 P2 P3 n1 n2 n3 (log-mel only: the first 4)
constexpr int kRegion = kPairs * kPairWords; // floats per warp; the 32x34 float2 exchange buffer (2176) aliases it
constexpr int kZeroRun = 127;             // float2 slot of every row kept at (0, 0) during the combine step
constexpr int kTwStride = 68;             // 32 float2 + pad
constexpr int kWinStride = 36;            // 32 floats + pad
constexpr int kXStride = 34;              // exchange buffer row stride in float2 (even: 128-bit reads)
constexpr int kItemRow = 544;             // item form: float2 per pair-row in natural bin order (513 bins + room to read a class length past the end)
constexpr int kItemRegion = 4 * 2 * kItemRow;   // item form: floats per warp, a whole number of 128-byte lines (a stride of 187.5 lines was measured 12 % slower:
                                                // every other warp's 256-byte accesses then straddle three lines)
static_assert(kItemRegion % 32 == 0, "warp regions start on 128-byte lines");
constexpr int kWabStride = 36;            // floats per lane in the (a, b) weight table: 17 float2 + pad, 36*l mod 32 = 4l
// Two real channels share each complex transform; splitting the packed spectrum leaves a channel with its partner's
// rounding noise, about -120 dB relative to the PARTNER (measured: tests/test_imbalance_gpu.py).  A frame in which some
// mel band of a channel lies that far under its partner's that the noise would show -- 50 dB for a log-mel row, 30 dB
// for W (channel 0), whose relative error goes straight into the normalised intensity vector -- is noted in a
// per-block list and transformed again at the end of the kernel with every channel alone in its transform (the
// partner slot zero), which is exact down to digital silence.  Ordinary material never takes that path.
// Mechanics: the main kernel marks such a frame by a sentinel (a NaN pattern no arithmetic produces) in element 0 of
// one of its output rows.  A block that marked anything launches, from the device and into the tail of its own grid
// (cudaStreamTailLaunch: it runs once the whole main grid has finished, and the stream's next kernel waits for it),
// one block of the same kernel template in its kRedo form, which looks through that block's frames for sentinels and
// recomputes them.  Ordinary material launches nothing.  (Measured alternatives: the slow path inside the main kernel
// wrecks the register allocation of the main loop, 0.73 ms instead of 0.41 ms; an unconditional second launch from
// the host that scans all frames costs 10 us per call.)  No state outside the output tensor, so calls on different
// streams do not interact.
// Band powers of stationary noise fluctuate (a narrow band is a chi-square with few degrees of freedom), so a single
// band far under its partner's is no evidence of a level difference between channels: the loose thresholds only
// count when three of eight neighbouring bands agree; one band alone must cross the strict threshold.
constexpr float kTauRow = 1e-5f;          // loose: band power ratio under which a log-mel row would lose accuracy (50 dB)
constexpr float kTauW = 2.5e-3f;          // loose, channel 0 of the IV kernel (26 dB: its relative error goes into every IV value)
constexpr float kTauRowStrict = 1e-7f;    // strict (70 dB)
constexpr float kTauWStrict = 1e-5f;      // strict, channel 0 of the IV kernel (50 dB)
constexpr uint32_t kRedoMark = 0x7fc5e1d0u;   // quiet NaN with a payload

__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// TIn = float (the reference's input) or int16_t (PCM as decoded from wav/flac: soundfile's float32
// conversion is s / 32768, folded exactly into the window: a.in_scale = 2^-15)
// kIV = true : one warp = one frame of a 4-channel clip -> 4 log-mel + 3 IV rows (LogmelIV_Extractor).
// kIV = false: log-mel only (Logmel_Extractor, or channels >= 4 of a wider IV call).  The four transform
//              slots of a warp then take four consecutive (frame, channel) jobs, j = t * Cj + c with
//              Cj = C - c_lo channels, so any channel count keeps all four slots busy (C = 1: four frames).
// kItem = true: item form of the mel projection (main form of the 4-channel IV kernel).  The packed spectra Z go to shared
//              memory as they leave the transform, already in the order the mel step reads them: bin k <= 512 into a "P" plane,
//              bin 1024 - k into an "M" plane at the same place, each at (position, lane) of the piece of its segment that
//              the plan gave to `lane` (PlanDev::iw / idst).  Lane l then walks its pieces: per position four conflict-free
//              64-bit loads, the untangle / power / IV arithmetic of that bin (no shuffles: both halves of the conjugate pair
//              are in the lane) and seven FFMA2 into the sums of the piece.  Every lane runs the same instruction stream, a
//              piece belongs to one segment, so there are no run boundaries, no predicated partial-sum stores and no
//              restart multiplies; the band-per-lane combine step reads at most four piece sums per list as before.
template <int W, typename TIn, bool kIV, bool kRedo = false, bool kItem = false>
__global__ void __launch_bounds__(W * 32, 1)
foa_iv2_kernel(const FoaArgs a, const PlanDev pd) {
    static_assert(!kItem || !kRedo, "item form: main form only (the redo form keeps the run form of the mel step)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // lane-major tables read with 128-bit loads: row stride = 4 (mod 32) words -> conflict-free
    float* tw_s = reinterpret_cast<float*>(smem_raw);                      // [lane][kTwStride]: (cos, -sin) of W1024^(lane*brev5(p)), p = 0..31
    float* win_s = tw_s + 32 * kTwStride;                                  // [lane][kWinStride]: window[32*m + lane] * 0.5, m = 0..31
    float* wab_s = win_s + 32 * kWinStride;                                // 32 * kWabStride          (item form: iw, iP * 32 float2)
    int* gseg_s = reinterpret_cast<int*>(wab_s + (kItem ? 64 * pd.iP : 32 * kWabStride));   // gseg_pad (item form: none)
#ifdef SELD_TMEM_TABLES
    // item form with the tables in tensor memory: shared memory holds the per-warp regions only (what is left of the 228 KB is L1)
    float* R_all = kItem ? reinterpret_cast<float*>(smem_raw) : reinterpret_cast<float*>(gseg_s + pd.gseg_pad);   // W * region
#else
    float* R_all = reinterpret_cast<float*>(gseg_s + (kItem ? 0 : pd.gseg_pad));            // W * region
#endif
    constexpr int region = kItem ? kItemRegion : kRegion;                  // floats per warp (the exchange buffer aliases it)
    int* marked_s = reinterpret_cast<int*>(R_all + W * region);            // main form: some warp of this block marked a frame

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float in_scale = a.in_scale;
    // redo form: the work items of main block `(int)blockIdx.x`: its tiles redo_block, + redo_grid, ..., W items each
    // (kIV: frames; log-mel only: groups of four (frame, channel) jobs), item j = (tile number j / W, warp slot j % W)
    const int redo_Cj = a.C - a.c_lo;
    const int64_t redo_J = (int64_t)a.T * redo_Cj;                          // log-mel only: jobs per clip
#ifdef SELD_CONTIG
    const int redo_t0 = kRedo ? (int)(((int64_t)(int)blockIdx.x * a.n_tiles) / a.redo_grid) : 0;
    const int redo_tiles = kRedo ? (int)(((int64_t)((int)blockIdx.x + 1) * a.n_tiles) / a.redo_grid) - redo_t0 : 0;
#else
    const int redo_tiles = (kRedo && a.n_tiles > (int)blockIdx.x) ? (a.n_tiles - 1 - (int)blockIdx.x) / a.redo_grid + 1 : 0;
#endif
    const int redo_items = redo_tiles * W;
    auto redo_item = [&](int j, int& b, int& grp) -> bool {                 // false: no such frame / group
#ifdef SELD_CONTIG
        const int tile = redo_t0 + j / W;
#else
        const int tile = (int)blockIdx.x + (j / W) * a.redo_grid;
#endif
        b = tile / a.tiles_per_clip;
        grp = (tile - b * a.tiles_per_clip) * W + (j % W);
        return kIV ? grp < a.T : (int64_t)4 * grp < redo_J;
    };
    auto is_marked = [&](int j) -> bool {
        int b, grp;
        if (j >= redo_items || !redo_item(j, b, grp)) return false;
        const float* const o0 = a.out + ((int64_t)b * a.Cout) * ((int64_t)a.T * pd.n_mels);
        if constexpr (kIV) {
            return __float_as_uint(o0[((int64_t)a.C * a.T + grp) * pd.n_mels]) == kRedoMark;
        } else {
            bool marked = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t jj = (int64_t)4 * grp + q;
                if (jj < redo_J)
                    marked = marked || __float_as_uint(o0[((int64_t)(a.c_lo + (int)(jj % redo_Cj)) * a.T + (jj / redo_Cj)) * pd.n_mels]) == kRedoMark;
            }
            return marked;
        }
    };
#ifdef SELD_TMEM_TABLES
    constexpr bool kSmemTables = !kItem;
#else
    constexpr bool kSmemTables = true;
#endif
    if constexpr (kSmemTables)
    for (int i = tid; i < 1024; i += W * 32) {
        const int l = i & 31, r = i >> 5;                                  // pd.tw is [ka][lane], pd.win is [32*m + lane]
        const float2 w = pd.tw[brev5(r) * 32 + l];
        tw_s[l * kTwStride + 2 * r] = w.x; tw_s[l * kTwStride + 2 * r + 1] = w.y;
        win_s[l * kWinStride + r] = pd.win[i] * in_scale;                   // int16 PCM: the 2^-15 of soundfile's conversion, exact
    }
    if constexpr (kItem) {
        if constexpr (kSmemTables)
            for (int i = tid; i < 64 * pd.iP; i += W * 32) wab_s[i] = reinterpret_cast<const float*>(pd.iw)[i];
        for (int i = tid; i < W * region; i += W * 32) R_all[i] = 0.0f;     // the entries behind bin 512 are read with zero weights: keep them finite
    } else {
        for (int i = tid; i < 32 * kWabStride; i += W * 32) wab_s[i] = pd.wab[i];
        for (int i = tid; i < pd.n_mels + 2; i += W * 32) gseg_s[i] = pd.gseg[i];
    }
    if (tid == 0) *marked_s = 0;
    uint32_t tmem_w = 0;                                                    // item form: tensor-memory address of this warp's lane quarter, column 0
#ifdef SELD_TMEM_TABLES
    if constexpr (kItem) {
        tmem_w = tmem_tables_alloc(reinterpret_cast<uint32_t*>(marked_s + 1), warp);
        if (warp < 4) {                                                     // one copy of the tables per lane quarter (warps w and w + 4 share it)
            uint32_t v[32];
#pragma unroll
            for (int m = 0; m < 32; ++m) v[m] = __float_as_uint(pd.win[32 * m + lane] * in_scale);
            tmem_st32(tmem_w, v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {                                   // positions 16 h .. 16 h + 15: (cos, -sin) pairs
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 w = pd.tw[brev5(16 * h + r) * 32 + lane];
                    v[2 * r] = __float_as_uint(w.x); v[2 * r + 1] = __float_as_uint(w.y);
                }
                tmem_st32(tmem_w + 32 + 32 * h, v);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {                                   // mel weights (a, b) of positions 16 h .. 16 h + 15
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
    }
#endif
    __syncthreads();
#ifdef SELD_TMEM_TABLES
    if constexpr (kItem) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif

    float* R = R_all + warp * region;
    float2* scratch = reinterpret_cast<float2*>(R);
    const uint32_t runmask = pd.runmask[lane];
    const int g0 = pd.g0[lane];
    const int hop = pd.hop, M = pd.n_mels;
    const float eps = pd.eps, amin = pd.amin;
    // A pair-row is cut in chunks of 16 bins = 128 bytes; the 16-byte sub-chunk s (two bins) of chunk c is stored at
    // position s ^ (c & 7), which makes both sides conflict-free:
    // writer: bin k = lane + 32*kb (chunk 2*kb + lane/16) lands at word 64*kb + wofs[kb & 3] of its pair-row
    int wofs[4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
        wofs[x] = 32 * (lane >> 4) + 4 * ((((lane & 15) >> 1) ^ ((2 * x + (lane >> 4)) & 7))) + 2 * (lane & 1);
    // reader: lane owns chunk c = lane; its logical sub-chunk i sits at word 32c + 4*(i ^ (c & 7)) (computed per frame)

    const int64_t ch_stride = (int64_t)a.T * M;
    // combine step, bands lane and lane+32: run numbers of segment m (V) and m+1 (U), four slots each
    auto pack_runs = [&](int lo, int hi) {
        uint32_t p = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) p |= (uint32_t)(lo + i < hi ? lo + i : kZeroRun) << (8 * i);
        return p;
    };
    constexpr uint32_t kNoRuns = kZeroRun * 0x01010101u;
    uint32_t slotV0 = kNoRuns, slotU0 = kNoRuns, slotV1 = kNoRuns, slotU1 = kNoRuns;
    int ist[4] = {0, 0, 0, 0};                                             // item form: first bin this lane reads in class c
    if constexpr (kItem) {
#pragma unroll
        for (int c = 0; c < 4; ++c) ist[c] = pd.istart[c * 32 + lane];
    }
    constexpr int kItemZero = 128, kItemSlots = 136;                       // item form: slot kept at zero, slots per array (4 classes x 32 lanes + zero + pad)
    if constexpr (kItem) {
        slotV0 = slotU0 = slotV1 = slotU1 = kItemZero * 0x01010101u;
        if (lane < M) { slotV0 = pd.islot[2 * lane]; slotU0 = pd.islot[2 * lane + 1]; }
        if (lane + 32 < M) { slotV1 = pd.islot[2 * (lane + 32)]; slotU1 = pd.islot[2 * (lane + 32) + 1]; }
    } else {
        if (lane < M) { slotV0 = pack_runs(gseg_s[lane], gseg_s[lane + 1]); slotU0 = pack_runs(gseg_s[lane + 1], gseg_s[lane + 2]); }
        if (lane + 32 < M) { slotV1 = pack_runs(gseg_s[lane + 32], gseg_s[lane + 33]); slotU1 = pack_runs(gseg_s[lane + 33], gseg_s[lane + 34]); }
    }
    // item form: longest V / U list of the bands of round 0 (bands 0-31) and round 1, four bits each (warp-uniform)
    int itemCnt = 0;
    if constexpr (kItem) {
        auto cnt = [&](uint32_t p) {
            int n = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) n += ((p >> (8 * i)) & 0xff) != (uint32_t)kItemZero;
            return __reduce_max_sync(0xffffffffu, n > 0 ? n : 1);
        };
        itemCnt = cnt(slotV0) | (cnt(slotU0) << 4) | (cnt(slotV1) << 8) | (cnt(slotU1) << 12);
    }
    // fourth slot in use by any band?  (warp-uniform; most banks never split a segment over four chunks)
    constexpr uint32_t kAbsent = kItem ? kItemZero : kZeroRun;
    const bool four = __any_sync(0xffffffffu, (slotV0 >> 24) != kAbsent || (slotU0 >> 24) != kAbsent ||
                                              (slotV1 >> 24) != kAbsent || (slotU1 >> 24) != kAbsent);

#ifdef SELD_PHASE_TIMING
    long long phase_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long phase_t0 = clock64();
#endif
    // ---- mel projection of the rows a warp has just written (walk + combine + dB + store).  kCheck: also report whether
    // some band of a slot lies too far under its partner's for the packed transform (the frame is then redone, see below)
    auto mel_rows = [&](auto check_c, int b, int t, const int (&tk)[4], const int (&ck)[4], const bool (&vk)[4]) -> bool {
        constexpr bool kCheck = decltype(check_c)::value;
        bool bad = false;
        PHASE_MARK(6);   // pointwise
        // ---------------- mel step 1: chunk walk of all rows, per-run partial sums (U, V) left in the rows
        constexpr int NR = kIV ? kRows : 4;
        constexpr int NP = kIV ? kPairs : 2;                                // pair-rows in use
#ifndef ABL_NOWALK
        // the walk handles NPP pair-rows at a time: all of them in the 255-register builds, two at a time where registers
        // are scarce (W > 8: 168 registers per thread)
        auto walk_pairs = [&](auto f0_c, auto npp_c) {
            constexpr int F0 = decltype(f0_c)::value, NPP = decltype(npp_c)::value;
            float2 wv[17];
            const float4* wp = reinterpret_cast<const float4*>(wab_s + lane * kWabStride);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = wp[i];
                wv[2 * i] = make_float2(v.x, v.y);
                wv[2 * i + 1] = make_float2(v.z, v.w);
            }
            wv[16] = *reinterpret_cast<const float2*>(wab_s + lane * kWabStride + 32);
            float2 q[NPP][17];                                              // (row, row') values of the lane's 16 (+1) bins
#pragma unroll
            for (int f = 0; f < NPP; ++f) {
                const float* row = R + (F0 + f) * kPairWords;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(row + 32 * lane + 4 * (i ^ (lane & 7)));
                    q[f][2 * i] = make_float2(v.x, v.y); q[f][2 * i + 1] = make_float2(v.z, v.w);
                }
                q[f][16] = lane == 31 ? *reinterpret_cast<const float2*>(row + 1024) : make_float2(0.0f, 0.0f);
            }
            __syncwarp();                                                   // everyone holds its bins: rows may be overwritten
            PHASE_MARK(9);   // walk: weights + rows in registers
            if (lane < NPP) reinterpret_cast<float4*>(R + (F0 + lane) * kPairWords)[kZeroRun] = make_float4(0.f, 0.f, 0.f, 0.f);
            // per pair-row: U = sum a_k q_k and V = sum b_k q_k of the current run, each for both rows of the pair
            float2 U[NPP], V[NPP];
            float2* po = reinterpret_cast<float2*>(R + F0 * kPairWords) + 2 * g0;   // slot of run r: words 4r..4r+3 = (U, U', V, V'), two 64-bit stores
                                                                            // (one 128-bit store would need the four values moved into an aligned register quad)
            {
                const float2 aa = make_float2(wv[0].x, wv[0].x), bb = make_float2(wv[0].y, wv[0].y);
#pragma unroll
                for (int f = 0; f < NPP; ++f) { U[f] = __fmul2_rn(aa, q[f][0]); V[f] = __fmul2_rn(bb, q[f][0]); }
            }
            // branch-free: where a new run starts the finished sums are stored and the accumulators restart
            // (acc * keep with keep = 0); the weights are broadcast once per bin for all rows
            static_for<1, 17>([&](auto ji) {
                constexpr int j = decltype(ji)::value;
                const bool start = (runmask >> j) & 1u;
                const float keep = start ? 0.0f : 1.0f;
                const float2 kk = make_float2(keep, keep);
                const float2 aa = make_float2(wv[j].x, wv[j].x), bb = make_float2(wv[j].y, wv[j].y);
#pragma unroll
                for (int f = 0; f < NPP; ++f) {
                    if (start) { po[f * (kPairWords / 2)] = U[f]; po[f * (kPairWords / 2) + 1] = V[f]; }
                    U[f] = __ffma2_rn(aa, q[f][j], __fmul2_rn(U[f], kk));
                    V[f] = __ffma2_rn(bb, q[f][j], __fmul2_rn(V[f], kk));
                }
                po += start ? 2 : 0;
            });
#pragma unroll
            for (int f = 0; f < NPP; ++f) { po[f * (kPairWords / 2)] = U[f]; po[f * (kPairWords / 2) + 1] = V[f]; }
        };
        if constexpr (kItem) {
            // the piece sums are in place already (main loop)
        } else if constexpr (W > 8 && NP == 4) {
            walk_pairs(std::integral_constant<int, 0>{}, std::integral_constant<int, 2>{});
            walk_pairs(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{});
        } else {
            walk_pairs(std::integral_constant<int, 0>{}, std::integral_constant<int, NP>{});
        }
#endif
        __syncwarp();

        PHASE_MARK(7);   // mel walk
#ifndef ABL_NOCOMB
        // ---------------- mel step 2: band per lane, out[m] = sum V(runs of segment m) + sum U(runs of segment m+1)
        {
            // destination of row f (one 64-float line of the output per row): kIV: channels 0-3 and the three
            // IV channels of frame t; log-mel only: (channel, frame) of slot f, if that slot holds a job
            float* const ob = a.out + ((int64_t)b * a.Cout) * ch_stride + (kIV ? (int64_t)t * M : 0);
            auto emit = [&](int f, int m, float v) {
                if (f < 4) v = 3.01029995663981195f * lg2_ftz(fmaxf(v, amin));   // 10*log10(max(v, amin))
                if constexpr (kIV) {
                    ob[(f < 4 ? f : a.C + f - 4) * ch_stride + m] = v;
                } else {
                    if (vk[f & 3]) ob[ck[f & 3] * ch_stride + (int64_t)tk[f & 3] * M + m] = v;
                }
            };
            // imbalance check (see kTauRow / kTauW): rows 0..3 are the powers of the four transform slots, partners (0,1), (2,3)
            // (log-mel only: a slot without a job is exactly zero and leaks nothing into its partner)
            const bool pair01 = kIV || (vk[0] && vk[1]), pair23 = kIV || (vk[2] && vk[3]);
            auto unbalanced = [&](const float (&v)[NR], float t0, float t) {
                return (pair01 && (v[0] < t0 * v[1] || v[1] < t * v[0])) || (pair23 && (v[2] < t * v[3] || v[3] < t * v[2]));
            };
            // some aligned group of eight bands (a byte of the ballot) with three or more bits set
            auto clustered = [](uint32_t m) {
                uint32_t c = m - ((m >> 1) & 0x55555555u);
                c = (c & 0x33333333u) + ((c >> 2) & 0x33333333u);
                c = (c + (c >> 4)) & 0x0f0f0f0fu;
                return ((c + 0x05050505u) & 0x08080808u) != 0u;
            };
            if (M <= 64) {
                // fixed slots: <= 4 runs per segment, run numbers held packed in registers (absent -> the zero run);
                // all loads are issued up front (a data-dependent slot count was measured 2.4 % slower)
                auto combine = [&](auto four_c) {
                    constexpr bool kFour = decltype(four_c)::value;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int m = lane + 32 * r;
                        bool loose = false;
                        if (m < M) {
                            const uint32_t pv = r ? slotV1 : slotV0, pu = r ? slotU1 : slotU0;
                            float2 vv[NP];                                  // both rows of a pair per packed add
                            if constexpr (kItem) {
                                // piece sums: arrays [U pair 0 | V pair 0 | U 1 | V 1 | U 2 | V 2 | (U, V) of n2], kItemSlots float2 each
                                const float2* S = reinterpret_cast<const float2*>(R);
                                const int iv0 = pv & 0xff, iv1 = (pv >> 8) & 0xff, iv2 = (pv >> 16) & 0xff, iv3 = pv >> 24;
                                const int iu0 = pu & 0xff, iu1 = (pu >> 8) & 0xff, iu2 = (pu >> 16) & 0xff, iu3 = pu >> 24;
                                // only as many list entries as some band of this round has (warp-uniform counts: the low bands
                                // are one or two pieces each, the high ones up to four)
                                const int nv = (r ? itemCnt >> 8 : itemCnt) & 15, nu = (r ? itemCnt >> 12 : itemCnt >> 4) & 15;
#pragma unroll
                                for (int f = 0; f < (kIV ? 3 : 2); ++f) {
                                    const float2* su = S + (2 * f) * kItemSlots;
                                    const float2* sv = su + kItemSlots;
                                    float2 acc = vadd(sv[iv0], su[iu0]);
                                    if (nv > 1) acc = vadd(acc, sv[iv1]);
                                    if (nu > 1) acc = vadd(acc, su[iu1]);
                                    if (nv > 2) acc = vadd(acc, sv[iv2]);
                                    if (nu > 2) acc = vadd(acc, su[iu2]);
                                    if (nv > 3) acc = vadd(acc, sv[iv3]);
                                    if (nu > 3) acc = vadd(acc, su[iu3]);
                                    vv[f] = acc;
                                }
                                if constexpr (kIV) {
                                    const float2* s3 = S + 6 * kItemSlots;
                                    float t3 = s3[iv0].y + s3[iu0].x;
                                    if (nv > 1) t3 += s3[iv1].y;
                                    if (nu > 1) t3 += s3[iu1].x;
                                    if (nv > 2) t3 += s3[iv2].y;
                                    if (nu > 2) t3 += s3[iu2].x;
                                    if (nv > 3) t3 += s3[iv3].y;
                                    if (nu > 3) t3 += s3[iu3].x;
                                    vv[NP - 1] = make_float2(t3, 0.0f);
                                }
                            } else {
#pragma unroll
                                for (int f = 0; f < NP; ++f) {
                                    const float2* rowp = reinterpret_cast<const float2*>(R + f * kPairWords);   // slot r: [2r] = U pair, [2r+1] = V pair
                                    const float2 v0 = rowp[2 * (pv & 0xff) + 1], v1 = rowp[2 * ((pv >> 8) & 0xff) + 1];
                                    const float2 v2 = rowp[2 * ((pv >> 16) & 0xff) + 1];
                                    const float2 u0 = rowp[2 * (pu & 0xff)], u1 = rowp[2 * ((pu >> 8) & 0xff)];
                                    const float2 u2 = rowp[2 * ((pu >> 16) & 0xff)];
                                    if constexpr (kFour) {
                                        const float2 v3 = rowp[2 * (pv >> 24) + 1], u3 = rowp[2 * (pu >> 24)];
                                        vv[f] = vadd(vadd(vadd(v0, v1), vadd(v2, v3)), vadd(vadd(u0, u1), vadd(u2, u3)));
                                    } else {
                                        vv[f] = vadd(vadd(vadd(v0, v1), v2), vadd(vadd(u0, u1), u2));
                                    }
                                }
                            }
                            float v[NR];
                            v[0] = vv[0].x; v[2] = vv[0].y; v[1] = vv[1].x; v[3] = vv[1].y;
                            if constexpr (kIV) { v[4] = vv[2].x; v[6] = vv[2].y; v[5] = vv[3].x; }
                            if constexpr (kCheck) {
                                bad = bad || unbalanced(v, kIV ? kTauWStrict : kTauRowStrict, kTauRowStrict);
                                loose = unbalanced(v, kIV ? kTauW : kTauRow, kTauRow);
                            }
#pragma unroll
                            for (int f = 0; f < NR; ++f) emit(f, m, v[f]);
                        }
                        if constexpr (kCheck) {
                            const uint32_t lm = __ballot_sync(0xffffffffu, loose);   // every lane votes (no short-circuit in front of it)
                            bad = clustered(lm) || bad;
                        }
                    }
                };
                if (four) combine(std::true_type{}); else combine(std::false_type{});
            } else {
                const float4* P = reinterpret_cast<const float4*>(R);        // slot g of pair-row f: (U, U', V, V')
                for (int m = lane; m < M; m += 32) {
                    const int ga = gseg_s[m], gb = gseg_s[m + 1], gc = gseg_s[m + 2];
                    float2 vv[NP];
#pragma unroll
                    for (int f = 0; f < NP; ++f) vv[f] = make_float2(0.0f, 0.0f);
                    for (int g = ga; g < gb; ++g) {
#pragma unroll
                        for (int f = 0; f < NP; ++f) { const float4 t = P[f * (kPairWords / 4) + g]; vv[f] = vadd(vv[f], make_float2(t.z, t.w)); }
                    }
                    for (int g = gb; g < gc; ++g) {
#pragma unroll
                        for (int f = 0; f < NP; ++f) { const float4 t = P[f * (kPairWords / 4) + g]; vv[f] = vadd(vv[f], make_float2(t.x, t.y)); }
                    }
                    float v[NR];
                    v[0] = vv[0].x; v[2] = vv[0].y; v[1] = vv[1].x; v[3] = vv[1].y;
                    if constexpr (kIV) { v[4] = vv[2].x; v[6] = vv[2].y; v[5] = vv[3].x; }
                    if constexpr (kCheck) bad = bad || unbalanced(v, kIV ? kTauW : kTauRow, kTauRow);   // wide banks: any band
#pragma unroll
                    for (int f = 0; f < NR; ++f) emit(f, m, v[f]);
                }
            }
        }
#endif
        return bad;
    };

    // tile = (clip tb, tile tr within the clip), advanced by the grid size without a division per frame (the
    // division was a 500-cycle dependent chain at the top of every frame)
#ifdef SELD_CONTIG
    // a block owns a contiguous range of tiles: consecutive tiles of a clip share 1024 - hop samples per frame, which the
    // block then finds in its L1 instead of fetching every tile cold from L2
    const int tile_lo = (int)(((int64_t)blockIdx.x * a.n_tiles) / gridDim.x);
    int tiles_left = (int)(((int64_t)(blockIdx.x + 1) * a.n_tiles) / gridDim.x) - tile_lo;
    int tb = tile_lo / a.tiles_per_clip, tr = tile_lo - tb * a.tiles_per_clip;
    auto next_tile = [&]() {
        --tiles_left; ++tr;
        if (tr >= a.tiles_per_clip) { tr = 0; ++tb; }
    };
    if constexpr (!kRedo) {
    for (; tiles_left > 0; next_tile()) {
#else
    int tb = blockIdx.x / a.tiles_per_clip, tr = blockIdx.x - tb * a.tiles_per_clip;
    auto next_tile = [&]() {
        tb += a.step_clip; tr += a.step_tile;
        if (tr >= a.tiles_per_clip) { tr -= a.tiles_per_clip; ++tb; }
    };
    if constexpr (!kRedo) {
    for (; tb < a.B; next_tile()) {
#endif
        const int b = tb;
        const int grp = tr * W + warp;                                      // kIV: the frame; else: group of 4 jobs
        int tk[4] = {0, 0, 0, 0}, ck[4] = {0, 0, 0, 0};                     // log-mel only: frame / channel of each transform slot
        bool vk[4] = {true, true, true, true};
        if constexpr (kIV) {
            if (grp >= a.T) continue;
        } else {
            const int Cj = a.C - a.c_lo;
            const int64_t J = (int64_t)a.T * Cj;
            if ((int64_t)4 * grp >= J) continue;
            // job j = 4 grp + k -> (frame j / Cj, channel j % Cj): ONE division per group, the other three jobs by counting on
            // (four 64-bit divisions and remainders per group were 11 % of this mode's time)
            const int64_t j0 = (int64_t)4 * grp;
            int tq, cq;
            if (J <= 0x7fffffff) { tq = (int)((uint32_t)j0 / (uint32_t)Cj); cq = (int)((uint32_t)j0 - (uint32_t)tq * (uint32_t)Cj); }
            else { tq = (int)(j0 / Cj); cq = (int)(j0 - (int64_t)tq * Cj); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                vk[k] = j0 + k < J;
                tk[k] = vk[k] ? tq : 0;
                ck[k] = a.c_lo + (vk[k] ? cq : 0);
                if (++cq == Cj) { cq = 0; ++tq; }
            }
        }
        const int t = grp;                                                  // kIV: the frame
        PHASE_MARK(0);
        const TIn* xb = reinterpret_cast<const TIn*>(a.x) + (int64_t)b * a.stride_b;

        float2 re[32], im[32];
        // ---------------- load + window: re = (slot 0, slot 2), im = (slot 1, slot 3)
        if constexpr (kIV) {                                                // slots = channels 0-3 of frame t
            const int64_t s0 = (int64_t)t * hop - 512;
#ifdef ABL_NOLOAD
            if (true) {
                static_for<0, 32>([&](auto mi) {
                    constexpr int m = decltype(mi)::value;
                    const float v = (float)(lane + t) * 1e-3f + (float)m;
                    re[m] = make_float2(v, v + 1.0f);
                    im[m] = make_float2(v + 2.0f, v + 3.0f);
                });
            } else
#endif
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
            } else {                                                        // reflect padding at the clip edges
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
        } else {
            bool interior = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t sk = (int64_t)tk[k] * hop - 512;
                interior = interior && vk[k] && sk >= 0 && sk + 1024 <= a.L;
            }
            if (interior) {
                const TIn* p0 = xb + ck[0] * a.stride_c + ((int64_t)tk[0] * hop - 512) + lane;
                const TIn* p1 = xb + ck[1] * a.stride_c + ((int64_t)tk[1] * hop - 512) + lane;
                const TIn* p2 = xb + ck[2] * a.stride_c + ((int64_t)tk[2] * hop - 512) + lane;
                const TIn* p3 = xb + ck[3] * a.stride_c + ((int64_t)tk[3] * hop - 512) + lane;
                static_for<0, 32>([&](auto mi) {
                    constexpr int m = decltype(mi)::value;
                    re[m] = make_float2((float)__ldg(p0 + 32 * m), (float)__ldg(p2 + 32 * m));
                    im[m] = make_float2((float)__ldg(p1 + 32 * m), (float)__ldg(p3 + 32 * m));
                });
            } else {                                                        // reflect padding at the clip edges / empty slots
                auto edge = [&](int k, int m) -> float {
                    if (!vk[k]) return 0.0f;
                    int64_t sidx = (int64_t)tk[k] * hop - 512 + 32 * m + lane;
                    if (sidx < 0) sidx = -sidx;
                    if (sidx >= a.L) sidx = 2 * (a.L - 1) - sidx;
                    return (float)__ldg(xb + ck[k] * a.stride_c + sidx);
                };
                static_for<0, 32>([&](auto mi) {
                    constexpr int m = decltype(mi)::value;
                    re[m] = make_float2(edge(0, m), edge(2, m));
                    im[m] = make_float2(edge(1, m), edge(3, m));
                });
            }
        }
        PHASE_MARK(1);   // loads issued
#ifdef SELD_PREFETCH_L1
        if constexpr (kIV) {
            // the hop of new samples this warp's next frame (t + W) adds to what the block has touched: one line per lane
            const int64_t pf = (int64_t)(t + W) * hop + 512 - hop + (lane & 7) * (128 / (int)sizeof(TIn));
            if (t + W < a.T && pf < a.L) {
                const TIn* pp = xb + (lane >> 3) * a.stride_c + pf;
                asm volatile("prefetch.global.L1 [%0];" :: "l"(pp));
            }
        }
#endif
#ifdef SELD_TMEM_TABLES
        if constexpr (kItem) {
            uint32_t wv[32];
            tmem_ld32(tmem_w, wv);
            static_for<0, 32>([&](auto mi) {
                constexpr int m = decltype(mi)::value;
                re[m] = vmuls(re[m], __uint_as_float(wv[m]));
                im[m] = vmuls(im[m], __uint_as_float(wv[m]));
            });
        } else
#endif
#ifndef ABL_NOWIN
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

        PHASE_MARK(2);   // loads landed + window
        // ---------------- two 1024-point FFTs at once: 32-pt, twiddle, exchange, 32-pt
#ifndef ABL_NOFFT1
        fft32(re, im);
#endif
        PHASE_MARK(3);   // first 32-pt
#ifdef SELD_TMEM_TABLES
        if constexpr (kItem) {
            static_for<0, 2>([&](auto hi) {
                constexpr int h = decltype(hi)::value;
                uint32_t tv[32];
                tmem_ld32(tmem_w + 32 + 32 * h, tv);
                static_for<0, 8>([&](auto qi) {
                    constexpr int p2 = 8 * h + decltype(qi)::value;         // positions 2*p2, 2*p2+1
                    constexpr int o = 4 * decltype(qi)::value;
                    const float wx = __uint_as_float(tv[o]), wy = __uint_as_float(tv[o + 1]), wz = __uint_as_float(tv[o + 2]), ww = __uint_as_float(tv[o + 3]);
                    if constexpr (p2 > 0) {                                 // position 0 is ka = 0: twiddle 1
                        const float2 r = re[2 * p2], i = im[2 * p2];
                        re[2 * p2] = vfmas(i, -wy, vmuls(r, wx));
                        im[2 * p2] = vfmas(i, wx, vmuls(r, wy));
                    }
                    const float2 r = re[2 * p2 + 1], i = im[2 * p2 + 1];
                    re[2 * p2 + 1] = vfmas(i, -ww, vmuls(r, wz));
                    im[2 * p2 + 1] = vfmas(i, wz, vmuls(r, ww));
                });
            });
        } else
#endif
#ifndef ABL_NOTW
        static_for<0, 16>([&](auto pi) {
            constexpr int p2 = decltype(pi)::value;                         // positions 2*p2, 2*p2+1
            const float4 w4 = *reinterpret_cast<const float4*>(tw_s + lane * kTwStride + 4 * p2);
            if constexpr (p2 > 0) {                                         // position 0 is ka = 0: twiddle 1
                const float2 r = re[2 * p2], i = im[2 * p2];
                re[2 * p2] = vfmas(i, -w4.y, vmuls(r, w4.x));
                im[2 * p2] = vfmas(i, w4.x, vmuls(r, w4.y));
            }
            const float2 r = re[2 * p2 + 1], i = im[2 * p2 + 1];
            re[2 * p2 + 1] = vfmas(i, -w4.w, vmuls(r, w4.z));
            im[2 * p2 + 1] = vfmas(i, w4.z, vmuls(r, w4.w));
        });
#endif
#ifndef ABL_NOEXCH
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
#endif
        PHASE_MARK(4);   // twiddle + exchange
#ifndef ABL_NOFFT2
        fft32(re, im);                                                      // position p: Z[lane + 32*brev5(p)]
#endif
        PHASE_MARK(5);   // second 32-pt

        if constexpr (kItem) {
            // ---------------- item form, untangle in the walk: the packed spectra go to shared memory as they are, bin k <= 512
            // to the P planes and bin 1024 - k to the M planes at index k (natural order: consecutive lanes, consecutive words)
            float2* const Pre = reinterpret_cast<float2*>(R);
            float2* const Pim = Pre + kItemRow;
            float2* const Mre = Pre + 2 * kItemRow;
            float2* const Mim = Pre + 3 * kItemRow;
            static_for<0, 16>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                Pre[32 * kb + lane] = re[brev5(kb)]; Pim[32 * kb + lane] = im[brev5(kb)];
            });
            static_for<16, 32>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;                    // bin 32 kb + lane = 1024 - m
                Mre[1024 - 32 * kb - lane] = re[brev5(kb)]; Mim[1024 - 32 * kb - lane] = im[brev5(kb)];
            });
            if (lane == 0) {                                                // bins 0 and 512 are their own mirror
                Mre[0] = re[brev5(0)]; Mim[0] = im[brev5(0)];
                Pre[512] = re[brev5(16)]; Pim[512] = im[brev5(16)];
            }
        } else
        // ---------------- per-bin quantities -> 7 rows
        {
            const int src = (32 - lane) & 31;
            const bool lane0 = lane == 0;
            static_for<0, 17>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                constexpr int p = brev5(kb & 31);
                const float2 zr = re[p], zi = im[p];
                float2 pr, pi;
#ifdef ABL_NOSHFL
                if constexpr (true) {
                    constexpr int pp = brev5(31 - (kb & 15));
                    pr = re[pp]; pi = im[pp];
                } else
#endif
                if constexpr (kb == 16) {
                    pr = zr; pi = zi;
                } else {
                    constexpr int pp = brev5(31 - kb), p0 = brev5((32 - kb) & 31);
                    // lane 0 holds its own partners (bins 32*kb <-> 32*(32-kb)); written as selects so that each costs one FSEL
                    const float sx = __shfl_sync(0xffffffffu, re[pp].x, src), sy = __shfl_sync(0xffffffffu, re[pp].y, src);
                    const float tx = __shfl_sync(0xffffffffu, im[pp].x, src), ty = __shfl_sync(0xffffffffu, im[pp].y, src);
                    pr = make_float2(lane0 ? re[p0].x : sx, lane0 ? re[p0].y : sy);
                    pi = make_float2(lane0 ? im[p0].x : tx, lane0 ? im[p0].y : ty);
                }
                // window was pre-scaled by 0.5: A = Z[k] + conj(Z[N-k]), B = (Z[k] - conj(Z[N-k])) / i
                const float2 ar = vadd(zr, pr), ai = vsub(zi, pi);          // (X0, X2)
                const float2 br = vadd(zi, pi), bi = vsub(pr, zr);          // (X1, X3)
                const float2 p02 = __ffma2_rn(ai, ai, __fmul2_rn(ar, ar));
                const float2 p13 = __ffma2_rn(bi, bi, __fmul2_rn(br, br));
                if constexpr (kIV) {
                    const float2 i13 = vfmas(bi, ai.x, vmuls(br, ar.x));    // Re(conj(X0) X1), Re(conj(X0) X3)
                    const float i2 = fmaf(ai.x, ai.y, ar.x * ar.y);         // Re(conj(X0) X2)
                    const float s = fmaf(i13.y, i13.y, fmaf(i2, i2, i13.x * i13.x));
#ifdef ABL_NOMUFU
                    const float inv = s + eps;
#else
                    const float nrm = sqrt_ftz(s) + eps;                     // one MUFU; sqrt(0) = 0, subnormal sums flush to 0 (far below eps)
                    const float inv = rcp_ftz(nrm);
#endif
#ifdef ABL_NOROWST
                    if (__float_as_uint(p02.x + p13.x + i13.x * inv + i2 * inv + p02.y + p13.y + i13.y) == 0x12345678u) {
#else
                    if (kb < 16 || lane == 0) {
#endif
                        float2* q = reinterpret_cast<float2*>(R + 64 * kb + wofs[kb & 3]);
                        q[0 * (kPairWords / 2)] = p02;
                        q[1 * (kPairWords / 2)] = p13;
                        q[2 * (kPairWords / 2)] = vmuls(i13, inv);
                        q[3 * (kPairWords / 2)] = make_float2(i2 * inv, 0.0f);
                    }
                } else {
                    if (kb < 16 || lane == 0) {
                        float2* q = reinterpret_cast<float2*>(R + 64 * kb + wofs[kb & 3]);
                        q[0 * (kPairWords / 2)] = p02;
                        q[1 * (kPairWords / 2)] = p13;
                    }
                }
            });
        }
        __syncwarp();
        if constexpr (kItem) {
            // ---------------- item form of the mel step: every lane walks its pieces (one per class) through the rows
            PHASE_MARK(6);   // pointwise
            float2 aU[4][3], aV[4][3], a3[4];                               // piece sums per class: three row pairs (U, V), and (U, V) of n2
            const float2* const iw_s = reinterpret_cast<const float2*>(wab_s);
            const float2* const Q = reinterpret_cast<const float2*>(R);
            static_for<0, 4>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
#pragma unroll
                for (int f = 0; f < 3; ++f) { aU[c][f] = make_float2(0.0f, 0.0f); aV[c][f] = make_float2(0.0f, 0.0f); }
                a3[c] = make_float2(0.0f, 0.0f);
                const int Lc = pd.iL[c];
                const float2* zp = Q + ist[c];
                const float2* wp = iw_s + 32 * pd.ioff[c] + lane;
#pragma unroll 1
                for (int j0 = 0; j0 < Lc; j0 += 4) {                        // class lengths are multiples of four: no remainder code
#ifdef SELD_TMEM_TABLES
                // the four weight pairs of this trip, from tensor memory.  (Asking for them one trip ahead, or for the window
                // before the global loads, was measured 17 % SLOWER: registers a tcgen05.ld has been told to write must not be
                // touched until the wait, and tying them to the wait costs copies and spills.)
                uint32_t w8[8];
                tmem_ld8(tmem_w + 96 + 2 * (pd.ioff[c] + j0), w8);
#endif
#pragma unroll
                for (int jj = 0; jj < 4; ++jj, ++zp, wp += 32) {
                    const float2 zr = zp[0], zi = zp[kItemRow], pr = zp[2 * kItemRow], pi = zp[3 * kItemRow];
#ifdef SELD_TMEM_TABLES
                    const float2 w = make_float2(__uint_as_float(w8[2 * jj]), __uint_as_float(w8[2 * jj + 1]));
#else
                    const float2 w = *wp;
#endif
                    // window was pre-scaled by 0.5: A = Z[k] + conj(Z[N-k]), B = (Z[k] - conj(Z[N-k])) / i
                    const float2 ar = vadd(zr, pr), ai = vsub(zi, pi);      // (X0, X2)
                    const float2 br = vadd(zi, pi), bi = vsub(pr, zr);      // (X1, X3)
                    const float2 p02 = __ffma2_rn(ai, ai, __fmul2_rn(ar, ar));
                    const float2 p13 = __ffma2_rn(bi, bi, __fmul2_rn(br, br));
                    const float2 aa = make_float2(w.x, w.x), bb = make_float2(w.y, w.y);
                    aU[c][0] = __ffma2_rn(aa, p02, aU[c][0]); aV[c][0] = __ffma2_rn(bb, p02, aV[c][0]);
                    aU[c][1] = __ffma2_rn(aa, p13, aU[c][1]); aV[c][1] = __ffma2_rn(bb, p13, aV[c][1]);
                    if constexpr (kIV) {
                        const float2 i13 = vfmas(bi, ai.x, vmuls(br, ar.x));    // Re(conj(X0) X1), Re(conj(X0) X3)
                        const float i2 = fmaf(ai.x, ai.y, ar.x * ar.y);         // Re(conj(X0) X2)
                        const float sq = fmaf(i13.y, i13.y, fmaf(i2, i2, i13.x * i13.x));
                        const float inv = rcp_ftz(sqrt_ftz(sq) + eps);          // one MUFU each; sqrt(0) = 0, subnormal sums flush to 0 (far below eps)
                        const float2 n13 = vmuls(i13, inv);
                        const float n2 = i2 * inv;
                        aU[c][2] = __ffma2_rn(aa, n13, aU[c][2]); aV[c][2] = __ffma2_rn(bb, n13, aV[c][2]);
                        a3[c] = __ffma2_rn(w, make_float2(n2, n2), a3[c]);
                    }
                }
                }
            });
            __syncwarp();                                                   // every lane is through with the rows: the sums may overwrite them
            PHASE_MARK(9);   // piece walk
            float2* const S = reinterpret_cast<float2*>(R);
            static_for<0, 4>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
                if (c < pd.iK) {
#pragma unroll
                    for (int f = 0; f < (kIV ? 3 : 2); ++f) {
                        S[(2 * f) * kItemSlots + 32 * c + lane] = aU[c][f];
                        S[(2 * f + 1) * kItemSlots + 32 * c + lane] = aV[c][f];
                    }
                    if constexpr (kIV) S[6 * kItemSlots + 32 * c + lane] = a3[c];
                }
            });
            if (lane < (kIV ? 7 : 4)) S[lane * kItemSlots + kItemZero] = make_float2(0.0f, 0.0f);
            __syncwarp();
        }

        const bool bad = mel_rows(std::true_type{}, b, t, tk, ck, vk);
        if (__any_sync(0xffffffffu, bad) && lane == 0) {                    // lane 0 also wrote element 0 of every row: program order
            *marked_s = 1;
            float* const o0 = a.out + ((int64_t)b * a.Cout) * ch_stride;
            if constexpr (kIV) {
                o0[(int64_t)a.C * ch_stride + (int64_t)t * M] = __uint_as_float(kRedoMark);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (vk[q]) o0[(int64_t)ck[q] * ch_stride + (int64_t)tk[q] * M] = __uint_as_float(kRedoMark);
            }
        }
        __syncwarp();                                                       // rows are reused by the next frame's exchange
        PHASE_MARK(8);   // mel combine + store
    }
#ifdef SELD_TMEM_TABLES
    if constexpr (kItem) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
    __syncthreads();
#ifdef SELD_TMEM_TABLES
    if constexpr (kItem) {
        tmem_tables_free(tmem_w, warp);
    }
#endif
    if (tid == 0) {
        // the last block to finish launches the redo grid if any block marked a frame (one launch, all SMs; tail launches of
        // single blocks would run one after the other)
        if (*marked_s != 0) atomicExch(&a.redo_flags[1], 1);
        __threadfence();
        if (atomicAdd(&a.redo_flags[0], 1) == (int)gridDim.x - 1) {
            __threadfence();
            const int any = atomicExch(&a.redo_flags[1], 0);
            atomicExch(&a.redo_flags[0], 0);                               // ready for the next launch that is handed this pair
#ifndef SELD_NO_DEVICE_LAUNCH                                               // (sanitizer builds: racecheck / synccheck do not support device-side launches)
            if (any) {
                FoaArgs ar = a;
                ar.redo_grid = (int)gridDim.x;
                foa_iv2_kernel<W, TIn, kIV, true><<<gridDim.x, W * 32, a.smem_bytes, cudaStreamTailLaunch>>>(ar, pd);
            }
#else
            (void)any;
#endif
        }
    }

    } else {
    // ---------------- frames whose channels were too unbalanced for the packed transform: once more, every slot alone
    // in its transform (pass 0: slots 0 and 2, pass 1: slots 1 and 3; the partner slot is zero, so nothing leaks and a
    // silent channel comes out as exact zeros, like the reference's one-FFT-per-channel, feature.py:49)
    {
        const int Cj = redo_Cj;
        const int64_t J = redo_J;
        auto redo = [&](int b, int grp) {
            int tk[4] = {grp, grp, grp, grp}, ck[4] = {0, 1, 2, 3};
            bool vk[4] = {true, true, true, true};
            if constexpr (!kIV) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t j = (int64_t)4 * grp + q;
                    vk[q] = j < J;
                    tk[q] = vk[q] ? (int)(j / Cj) : 0;
                    ck[q] = a.c_lo + (vk[q] ? (int)(j % Cj) : 0);
                }
            }
            const TIn* xb = reinterpret_cast<const TIn*>(a.x) + (int64_t)b * a.stride_b;
            float x0r[17], x0i[17], i2s[17];                                  // kIV: W spectrum and Re(conj(X0) X2) of pass 0
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                float2 re[32], im[32];
                const int ta = tk[pass], tb2 = tk[pass + 2];
                const bool va = vk[pass], vb = vk[pass + 2];
                const TIn* pa = xb + (int64_t)ck[pass] * a.stride_c;
                const TIn* pb = xb + (int64_t)ck[pass + 2] * a.stride_c;
                auto sample = [&](const TIn* p, bool valid, int tt, int m) -> float {
                    if (!valid) return 0.0f;
                    int64_t sidx = (int64_t)tt * hop - 512 + 32 * m + lane;
                    if (sidx < 0) sidx = -sidx;
                    if (sidx >= a.L) sidx = 2 * (a.L - 1) - sidx;
                    return (float)__ldg(p + sidx);
                };
                static_for<0, 32>([&](auto mi) {
                    constexpr int m = decltype(mi)::value;
                    re[m] = make_float2(sample(pa, va, ta, m), sample(pb, vb, tb2, m));
                    im[m] = make_float2(0.0f, 0.0f);
                });
                static_for<0, 8>([&](auto mi) {
                    constexpr int m4 = decltype(mi)::value;
                    const float4 w4 = *reinterpret_cast<const float4*>(win_s + lane * kWinStride + 4 * m4);
                    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) re[4 * m4 + e] = vmuls(re[4 * m4 + e], w[e]);
                });
                fft32(re, im);
                static_for<0, 16>([&](auto pi) {
                    constexpr int p2 = decltype(pi)::value;
                    const float4 w4 = *reinterpret_cast<const float4*>(tw_s + lane * kTwStride + 4 * p2);
                    if constexpr (p2 > 0) {
                        const float2 r = re[2 * p2], i = im[2 * p2];
                        re[2 * p2] = vfmas(i, -w4.y, vmuls(r, w4.x));
                        im[2 * p2] = vfmas(i, w4.x, vmuls(r, w4.y));
                    }
                    const float2 r = re[2 * p2 + 1], i = im[2 * p2 + 1];
                    re[2 * p2 + 1] = vfmas(i, -w4.w, vmuls(r, w4.z));
                    im[2 * p2 + 1] = vfmas(i, w4.z, vmuls(r, w4.w));
                });
                __syncwarp();
                // exchange behind pair-row 0 (which holds pass 0's powers during pass 1): floats [1056, 3232) of the region
                float2* xs = scratch + kPairWords / 2;
                static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; xs[brev5(p) * kXStride + lane] = re[p]; });
                __syncwarp();
                static_for<0, 16>([&](auto ji) {
                    constexpr int j = decltype(ji)::value;
                    const float4 v = *reinterpret_cast<const float4*>(xs + lane * kXStride + 2 * j);
                    re[2 * j] = make_float2(v.x, v.y); re[2 * j + 1] = make_float2(v.z, v.w);
                });
                __syncwarp();
                static_for<0, 32>([&](auto pi) { constexpr int p = decltype(pi)::value; xs[brev5(p) * kXStride + lane] = im[p]; });
                __syncwarp();
                static_for<0, 16>([&](auto ji) {
                    constexpr int j = decltype(ji)::value;
                    const float4 v = *reinterpret_cast<const float4*>(xs + lane * kXStride + 2 * j);
                    im[2 * j] = make_float2(v.x, v.y); im[2 * j + 1] = make_float2(v.z, v.w);
                });
                __syncwarp();
                fft32(re, im);
                const int src = (32 - lane) & 31;
                const bool lane0 = lane == 0;
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
                    const float2 ar = vadd(zr, pr), ai = vsub(zi, pi);      // spectra of the two live slots (the other two are zero)
                    const float2 pw = __ffma2_rn(ai, ai, __fmul2_rn(ar, ar));
                    const bool wr = kb < 16 || lane == 0;
                    float2* q = reinterpret_cast<float2*>(R + 64 * kb + wofs[kb & 3]);
                    if (wr) q[pass * (kPairWords / 2)] = pw;               // pass 0: (P0, P2), pass 1: (P1, P3)
                    if constexpr (kIV) {
                        if (pass == 0) {
                            x0r[kb] = ar.x; x0i[kb] = ai.x;
                            i2s[kb] = fmaf(ai.x, ai.y, ar.x * ar.y);       // Re(conj(X0) X2)
                        } else {
                            const float2 i13 = vfmas(ai, x0i[kb], vmuls(ar, x0r[kb]));   // Re(conj(X0) X1), Re(conj(X0) X3)
                            const float i2 = i2s[kb];
                            const float sq = fmaf(i13.y, i13.y, fmaf(i2, i2, i13.x * i13.x));
                            const float inv = rcp_ftz(sqrt_ftz(sq) + eps);
                            if (wr) {
                                q[2 * (kPairWords / 2)] = vmuls(i13, inv);
                                q[3 * (kPairWords / 2)] = make_float2(i2 * inv, 0.0f);
                            }
                        }
                    }
                });
                __syncwarp();
            }
            mel_rows(std::false_type{}, b, tk[0], tk, ck, vk);
            __syncwarp();
        };
        // scan: a warp looks at 32 items at a time (one per lane), then redoes the marked ones
        for (int j0 = warp * 32; j0 < redo_items; j0 += W * 32) {
            uint32_t todo = __ballot_sync(0xffffffffu, is_marked(j0 + lane));
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                int b, grp;
                redo_item(j0 + l, b, grp);
                redo(b, grp);
            }
        }
    }
    }
#ifdef SELD_PHASE_TIMING
    if (lane == 0) for (int i = 0; i < 10; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)phase_acc[i]);
#endif
}

// ---------------------------------------------------------------------------------------------
template <int W>
static size_t iv2_smem_bytes(const PlanDev& pd) {
#ifndef SELD_SMEM_PAD
#define SELD_SMEM_PAD 0
#endif
    return (size_t)(32 * kTwStride + 32 * kWinStride + 32 * kWabStride + pd.gseg_pad + W * kRegion + 4) * sizeof(float) + SELD_SMEM_PAD;
}

// Warps (= frames) per block.  One block per SM; more warps hide latency, fewer leave more
// registers per thread (8 -> 255, 12 -> 168; warps are allocated in fours).  Measured at cfg2: 8 -> 0.46 ms,
// 12 (spills) -> 0.51 ms; the 12-warp build only exists under -DSELD_EXPERIMENTS (SELD_IV2_WARPS=12).
static int iv2_warps() {
#ifdef SELD_IV2_W
    return SELD_IV2_W;
#elif defined(SELD_EXPERIMENTS)
    static int w = [] {
        const char* e = getenv("SELD_IV2_WARPS");
        const int v = e ? atoi(e) : 8;
        return (v == 8 || v == 12) ? v : 8;
    }();
    return w;
#else
    return 8;
#endif
}

bool foa_iv2_supported(const PlanDev& pd, size_t smem_optin) {
    return pd.fast_ok && iv2_smem_bytes<12>(pd) - SELD_SMEM_PAD <= smem_optin;
}

int foa_iv2_frames_per_tile() { return iv2_warps(); }

static size_t iv2_item_smem_bytes(const PlanDev& pd, int W) {
#ifdef SELD_TMEM_TABLES
    (void)pd;
    return (size_t)(W * kItemRegion + 4) * sizeof(float);                   // the tables live in tensor memory
#else
    return (size_t)(32 * kTwStride + 32 * kWinStride + 64 * pd.iP + W * kItemRegion + 4) * sizeof(float);
#endif
}

template <int W, typename TIn, bool kIV, bool kItem = false>
static cudaError_t iv2_launch_t(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    const size_t smem_redo = iv2_smem_bytes<W>(pd);                         // the redo form always runs the run form of the mel step
    const size_t smem = kItem ? iv2_item_smem_bytes(pd, W) : smem_redo;
    static std::atomic<uint64_t> attr_done{0};                              // per device, once per process and instantiation
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_relaxed) & bit)) {
        e = cudaFuncSetAttribute(foa_iv2_kernel<W, TIn, kIV, false, kItem>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(foa_iv2_kernel<W, TIn, kIV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.fetch_or(bit, std::memory_order_relaxed);
    }
    int gx = sm_count < a.n_tiles ? sm_count : a.n_tiles;
    FoaArgs aa = a;
    aa.step_clip = gx / a.tiles_per_clip; aa.step_tile = gx - aa.step_clip * a.tiles_per_clip;
    aa.smem_bytes = (int)smem_redo;                                         // blocks that mark frames launch the redo form themselves
    foa_iv2_kernel<W, TIn, kIV, false, kItem><<<gx, W * 32, smem, st>>>(aa, pd);
    return cudaGetLastError();
}

#ifdef SELD_PHASE_TIMING
extern "C" void seld_dev_phase_cycles(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_cycles, 16 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
}
#endif

cudaError_t foa_iv2_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
#ifdef SELD_IV2_W
    if (pd.item_ok && iv2_item_smem_bytes(pd, SELD_IV2_W) <= 227 * 1024) {
        if (a.in_i16) return iv2_launch_t<SELD_IV2_W, int16_t, true, true>(a, pd, sm_count, st);
        return iv2_launch_t<SELD_IV2_W, float, true, true>(a, pd, sm_count, st);
    }
    if (a.in_i16) return iv2_launch_t<SELD_IV2_W, int16_t, true>(a, pd, sm_count, st);
    return iv2_launch_t<SELD_IV2_W, float, true>(a, pd, sm_count, st);
#endif
#ifndef SELD_NO_ITEM
    if (pd.item_ok && iv2_item_smem_bytes(pd, 8) <= 227 * 1024) {           // item form of the mel step (the default for every bank it fits)
        if (a.in_i16) return iv2_launch_t<8, int16_t, true, true>(a, pd, sm_count, st);
        return iv2_launch_t<8, float, true, true>(a, pd, sm_count, st);
    }
#endif
    if (a.in_i16) return iv2_launch_t<8, int16_t, true>(a, pd, sm_count, st);   // PCM input: always the 8-warp build (frames_per_tile says 8 then, too)
#ifdef SELD_EXPERIMENTS
    if (iv2_warps() == 12) return iv2_launch_t<12, float, true>(a, pd, sm_count, st);
#endif
    return iv2_launch_t<8, float, true>(a, pd, sm_count, st);
}

// log-mel only, any channel count: channels [a.c_lo, a.C) of every clip; a.tiles_per_clip counts tiles of
// 8 warps x 4 (frame, channel) jobs
int foa_lm4_jobs_per_tile() { return 8 * 4; }
cudaError_t foa_lm4_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
#ifndef SELD_NO_ITEM
    if (pd.item_ok && iv2_item_smem_bytes(pd, 8) <= 227 * 1024) return iv2_launch_t<8, float, false, true>(a, pd, sm_count, st);
#endif
    return iv2_launch_t<8, float, false>(a, pd, sm_count, st);
}

}  // namespace seld
