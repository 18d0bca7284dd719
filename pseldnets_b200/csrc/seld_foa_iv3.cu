// FOA log-mel + intensity-vector kernel, third generation ("iv3"): the headline path.
//
//   LogmelIV_Extractor.forward (feature.py:39-56) + intensityvector (feature.py:93-117), C = 4.
//
// Two warps per frame.  Warp A transforms the packed pair (ch0 + i*ch1), warp B (ch2 + i*ch3);
// each 1024-point FFT is 32 x 32 over the warp's lanes with the two 32-point stages done in the
// internally packed "split" form of fft32.cuh (FADD2/FMUL2/FFMA2, 64 data registers per thread
// instead of the 128 the dual-transform kernel needed), so twice as many warps fit on an SM
// and latencies (global loads, the shared-memory exchange, shuffles, MUFU) hide behind each
// other.  The pair meets on a 64-thread named barrier:
//     both : load + window, FFT32, twiddle, exchange, FFT32                       -> bar (scratch free)
//     A    : untangle -> rows P0 P1 X0re X0im I1            -> bar ; mel walk of P0 P1
//     B    : bar.sync ; untangle, I2 I3, normalise -> rows P2 P3 n1 n2 n3 (over X0/I1)
//     both : bar ; mel walk of the remaining rows (A: P2 P3, B: n1 n2 n3) ; bar ;
//            band-per-lane combine + dB + coalesced store (A: 4 log-mel maps, B: 3 IV maps)
// The mel projection is the segment walk of the iv2 kernel (see seld_foa_iv2.cu).
// Samples come straight from global memory (coalesced 128-byte warp loads; the 1024-hop overlap
// of neighbouring frames is served by L1/L2); the next frame's lines are prefetched into L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fft32.cuh"
#include "mel_seg.cuh"
#include "seld_plan.h"

namespace seld {

namespace iv3 {
using melseg::kRowWords;
using melseg::kWabStride;
using melseg::mel_walk;
using melseg::mel_combine;
constexpr int kPlane = 32 * 34;           // exchange plane: [ka][lane], row stride 34 (even: 64-bit reads)
constexpr int kArea = 2 * kPlane;         // re + im plane of one warp
constexpr int kRegion = 2 * kArea;        // floats per frame (warp pair); the 7 rows (3696) alias it

using melseg::rsqrt_ftz;
using melseg::rcp_ftz;
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace iv3

template <int NP>
__global__ void __launch_bounds__(NP * 64, 1)
foa_iv3_kernel(const FoaArgs a, const PlanDev pd) {
    using namespace iv3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* tw4_s = reinterpret_cast<float4*>(smem_raw);                   // [16][32] (c_q, c_q+16, s_q, s_q+16)
    float2* win2_s = reinterpret_cast<float2*>(tw4_s + 512);               // [16][32] 0.5*(w[32*2p+l], w[32*(2p+1)+l])
    float* wab_s = reinterpret_cast<float*>(win2_s + 512);                 // 32 * kWabStride
    int* gseg_s = reinterpret_cast<int*>(wab_s + 32 * kWabStride);         // gseg_pad
    float* R_all = reinterpret_cast<float*>(gseg_s + pd.gseg_pad);         // NP * kRegion

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = warp & 1, pair = warp >> 1;
    for (int i = tid; i < 512; i += NP * 64) { tw4_s[i] = pd.tw4[i]; win2_s[i] = pd.win2[i]; }
    for (int i = tid; i < 32 * kWabStride; i += NP * 64) wab_s[i] = pd.wab[i];
    for (int i = tid; i < pd.n_mels + 2; i += NP * 64) gseg_s[i] = pd.gseg[i];
    __syncthreads();

    float* R = R_all + pair * kRegion;
    float* mre = R + role * kArea;                                         // this warp's exchange planes
    float* mim = mre + kPlane;
    const int bar = 1 + pair;
    const uint32_t runmask = pd.runmask[lane];
    const int g0 = pd.g0[lane];
    const int hop = pd.hop, M = pd.n_mels;
    const float eps = pd.eps, amin = pd.amin;
    int wofs[4], rofs[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        wofs[x] = 16 * (lane >> 4) + 4 * (((lane >> 2) & 3) ^ x) + (lane & 3);   // writer: bin lane+32kb -> 32kb + wofs[kb&3]
        rofs[x] = 16 * lane + 4 * (x ^ ((lane >> 1) & 3));                       // reader: quad x of chunk `lane`
    }
    const int64_t ch_stride = (int64_t)a.T * M;
    const int src = (32 - lane) & 31;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_clip;
        const int t = (tile - b * a.tiles_per_clip) * NP + pair;
        if (t >= a.T) continue;                                            // both warps of the pair skip together
        const float* xa = static_cast<const float*>(a.x) + (int64_t)b * a.stride_b + (int64_t)(2 * role) * a.stride_c;
        const int64_t s0 = (int64_t)t * hop - 512;

        float2 pr[16], pi[16];
        // ---------------- load + window: position p = samples (32*2p + lane, 32*(2p+1) + lane)
        if (s0 >= 0 && s0 + 1024 <= a.L) {
            const float* pa = xa + s0 + lane;
            const float* pb = pa + a.stride_c;
            static_for<0, 16>([&](auto pi_) {
                constexpr int p = decltype(pi_)::value;
                pr[p] = make_float2(__ldg(pa + 64 * p), __ldg(pa + 64 * p + 32));
                pi[p] = make_float2(__ldg(pb + 64 * p), __ldg(pb + 64 * p + 32));
            });
        } else {                                                            // reflect padding at the clip edges
            static_for<0, 16>([&](auto pi_) {
                constexpr int p = decltype(pi_)::value;
                int64_t s1 = s0 + 64 * p + lane, s2 = s1 + 32;
                if (s1 < 0) s1 = -s1;
                if (s1 >= a.L) s1 = 2 * (a.L - 1) - s1;
                if (s2 < 0) s2 = -s2;
                if (s2 >= a.L) s2 = 2 * (a.L - 1) - s2;
                pr[p] = make_float2(__ldg(xa + s1), __ldg(xa + s2));
                pi[p] = make_float2(__ldg(xa + a.stride_c + s1), __ldg(xa + a.stride_c + s2));
            });
        }
        {   // next frame of this pair: pull its 2 x 32 lines towards L2 while this one is computed
            const int tile_n = tile + gridDim.x;
            if (tile_n < a.n_tiles) {
                const int bn = tile_n / a.tiles_per_clip;
                const int tn = (tile_n - bn * a.tiles_per_clip) * NP + pair;
                const int64_t sn = (int64_t)tn * hop - 512 + 32 * lane;
                if (tn < a.T && sn >= 0 && sn + 32 <= a.L) {
                    const float* pn = static_cast<const float*>(a.x) + (int64_t)bn * a.stride_b + (int64_t)(2 * role) * a.stride_c + sn;
                    prefetch_l2(pn);
                    prefetch_l2(pn + a.stride_c);
                }
            }
        }
        static_for<0, 16>([&](auto pi_) {
            constexpr int p = decltype(pi_)::value;
            const float2 w2 = win2_s[p * 32 + lane];
            pr[p] = __fmul2_rn(pr[p], w2);
            pi[p] = __fmul2_rn(pi[p], w2);
        });

        // ---------------- 1024-point FFT of (cha + i*chb): 32-pt, twiddle, exchange, 32-pt
        fft32_dit(pr, pi);                                                  // position q': (Y[q], Y[q+16])
        pair_sync(bar);                                                     // partner is done with the previous frame's rows
        static_for<0, 16>([&](auto qi) {
            constexpr int qp = decltype(qi)::value;
            constexpr int q = brev4(qp);
            const float4 tw = tw4_s[q * 32 + lane];
            const float2 c2 = make_float2(tw.x, tw.y), s2 = make_float2(tw.z, tw.w);
            const float2 r = pr[qp], i = pi[qp];
            const float2 nr = __ffma2_rn(i, s2, __fmul2_rn(r, c2));
            const float2 ni = __ffma2_rn(r, make_float2(-s2.x, -s2.y), __fmul2_rn(i, c2));
            mre[q * 34 + lane] = nr.x; mre[(q + 16) * 34 + lane] = nr.y;
            mim[q * 34 + lane] = ni.x; mim[(q + 16) * 34 + lane] = ni.y;
        });
        __syncwarp();
        static_for<0, 16>([&](auto pi_) {
            constexpr int p = decltype(pi_)::value;
            pr[p] = *reinterpret_cast<const float2*>(mre + lane * 34 + 2 * p);
            pi[p] = *reinterpret_cast<const float2*>(mim + lane * 34 + 2 * p);
        });
        fft32_dit(pr, pi);                                                  // position q': Z[lane + 32q] (.x), Z[lane + 32(q+16)] (.y)
        pair_sync(bar);                                                     // both exchanges done: the region now holds rows

        float* ob = a.out + (((int64_t)b * a.Cout) * a.T + t) * M;

        // ---------------- untangle the two real channels, per-bin quantities -> rows
        // rows: 0 P0, 1 P1, 2 X0re -> n1, 3 X0im -> n2, 4 I1 -> n3, 5 P2, 6 P3
        if (role == 1) pair_sync(bar);                                      // B needs A's X0 / I1
        static_for<0, 17>([&](auto kbi) {
            constexpr int kb = decltype(kbi)::value;
            float zr, zi, qr, qi;
            if constexpr (kb == 16) {                                       // bin 512 (lane 0): its own partner
                zr = pr[brev4(0)].y; zi = pi[brev4(0)].y; qr = zr; qi = zi;
            } else {
                zr = pr[brev4(kb)].x; zi = pi[brev4(kb)].x;
                qr = __shfl_sync(0xffffffffu, pr[brev4(15 - kb)].y, src);
                qi = __shfl_sync(0xffffffffu, pi[brev4(15 - kb)].y, src);
                if (lane == 0) {                                            // bins 32*kb: partner 1024-32kb is in lane 0 too
                    if constexpr (kb == 0) { qr = zr; qi = zi; }
                    else { qr = pr[brev4(16 - kb)].y; qi = pi[brev4(16 - kb)].y; }
                }
            }
            // window pre-scaled by 0.5: Xa = Z[k] + conj(Z[N-k]), Xb = (Z[k] - conj(Z[N-k])) / i
            const float ar = zr + qr, ai = zi - qi, br = zi + qi, bi = qr - zr;
            const float pa = fmaf(ai, ai, ar * ar), pb = fmaf(bi, bi, br * br);
            if (kb < 16 || lane == 0) {
                float* q = R + 32 * kb + wofs[kb & 3];
                if (role == 0) {
                    q[0 * kRowWords] = pa;
                    q[1 * kRowWords] = pb;
                    q[2 * kRowWords] = ar;
                    q[3 * kRowWords] = ai;
                    q[4 * kRowWords] = fmaf(ai, bi, ar * br);               // I1 = Re(conj(X0) X1)
                } else {
                    const float x0r = q[2 * kRowWords], x0i = q[3 * kRowWords], i1 = q[4 * kRowWords];
                    const float i2 = fmaf(x0i, ai, x0r * ar);               // Re(conj(X0) X2)
                    const float i3 = fmaf(x0i, bi, x0r * br);               // Re(conj(X0) X3)
                    const float s = fmaf(i3, i3, fmaf(i2, i2, i1 * i1));
                    const float nrm = (s > 1e-37f ? s * rsqrt_ftz(s) : 0.0f) + eps;
                    const float inv = rcp_ftz(nrm);
                    q[5 * kRowWords] = pa;
                    q[6 * kRowWords] = pb;
                    q[2 * kRowWords] = i1 * inv;
                    q[3 * kRowWords] = i2 * inv;
                    q[4 * kRowWords] = i3 * inv;
                }
            }
        });
        float2 wv[17];                                                      // the lane's (a, b) mel weights
        {
            const float4* wp = reinterpret_cast<const float4*>(wab_s + lane * kWabStride);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = wp[i];
                wv[2 * i] = make_float2(v.x, v.y);
                wv[2 * i + 1] = make_float2(v.z, v.w);
            }
            wv[16] = *reinterpret_cast<const float2*>(wab_s + lane * kWabStride + 32);
        }
        if (role == 0) {
            pair_sync(bar);                                                 // X0 / I1 / P0 / P1 are in place (B is already waiting)
            mel_walk<2, 0, 1, 0>(R, wv, rofs, runmask, g0, lane);
        }
        pair_sync(bar);                                                     // all seven rows final
        if (role == 0) mel_walk<2, 5, 6, 0>(R, wv, rofs, runmask, g0, lane);
        else           mel_walk<3, 2, 3, 4>(R, wv, rofs, runmask, g0, lane);
        pair_sync(bar);                                                     // all partial sums in place
        if (role == 0) {
            float* const o[4] = {ob, ob + ch_stride, ob + 2 * ch_stride, ob + 3 * ch_stride};
            mel_combine<4, 0, 1, 5, 6, true>(R, gseg_s, M, lane, amin, o);
        } else {
            float* iv = ob + (int64_t)a.C * ch_stride;
            float* const o[4] = {iv, iv + ch_stride, iv + 2 * ch_stride, nullptr};
            mel_combine<3, 2, 3, 4, 0, false>(R, gseg_s, M, lane, amin, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int NP>
static size_t iv3_smem_bytes(const PlanDev& pd) {
    return 512 * 16 + 512 * 8 + (size_t)(32 * iv3::kWabStride + pd.gseg_pad + NP * iv3::kRegion) * sizeof(float);
}

// Frame slots (warp pairs) per block; one block per SM.  SELD_IV3_PAIRS overrides for experiments.
static int iv3_pairs() {
    static int v = [] {
        const char* e = getenv("SELD_IV3_PAIRS");
        const int n = e ? atoi(e) : 8;
        return (n == 6 || n == 8 || n == 9 || n == 10) ? n : 8;
    }();
    return v;
}

bool foa_iv3_supported(const PlanDev& pd, size_t smem_optin) {
    return pd.fast_ok && pd.tw4 != nullptr && iv3_smem_bytes<10>(pd) <= smem_optin;
}

int foa_iv3_frames_per_tile() { return iv3_pairs(); }

template <int NP>
static cudaError_t iv3_launch_t(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    const size_t smem = iv3_smem_bytes<NP>(pd);
    cudaError_t e = cudaFuncSetAttribute(foa_iv3_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int gx = sm_count < a.n_tiles ? sm_count : a.n_tiles;
    foa_iv3_kernel<NP><<<gx, NP * 64, smem, st>>>(a, pd);
    return cudaGetLastError();
}

cudaError_t foa_iv3_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    switch (iv3_pairs()) {
        case 6: return iv3_launch_t<6>(a, pd, sm_count, st);
        case 9: return iv3_launch_t<9>(a, pd, sm_count, st);
        case 10: return iv3_launch_t<10>(a, pd, sm_count, st);
        default: return iv3_launch_t<8>(a, pd, sm_count, st);
    }
}

}  // namespace seld
