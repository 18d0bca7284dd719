// Fused FOA feature front-end for sm_100a: framing + window + 1024-point STFT + power + mel + dB
// + intensity vector, one pass from the waveform in HBM to the (B, C+3, T, M) feature map in HBM.
//
// Reference behaviour restated (not ported): /root/reference/src/utils/feature.py
//   LogmelIV_Extractor.forward :39-56, Logmel_Extractor.forward :76-91, intensityvector :93-117,
// whose arithmetic lives in torchaudio Spectrogram/MelScale/AmplitudeToDB (torch.stft with
// center=True, pad_mode='reflect', onesided; |X|^2 @ fb; 10*log10(clamp(.,1e-10))).
//
// Work decomposition
//   tile   = kWarps consecutive frames of one clip (one channel group of <= 4 channels)
//   block  = kWarps warps, persistent over tiles; 2 blocks resident per SM
//   warp   = one frame: two packed complex 1024-point FFTs (channels (0,1) and (2,3)), each done
//            as 32 x 32: a 32-point FFT in registers per lane, twiddle, 32x32 exchange through
//            shared memory, second 32-point FFT in registers.  Lane l ends up holding bins
//            l + 32*kb; the conjugate-symmetric partner bin 1024-k sits in lane (32-l)&31, so the
//            two-real-channels-per-complex-FFT untangle is one warp shuffle per value.
//   mel    = band-sparse: each lane owns two mel bands and walks their (lo, count) bin ranges
//            over the per-bin quantities the warp just wrote to its shared-memory rows.
// Each input sample is read from HBM/L2 once per tile (halo re-read 1024-hop per tile), each
// output element is written once; nothing else touches global memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft32.cuh"
#include "seld_plan.h"

namespace seld {

constexpr int kWarps = 6;                 // frames per tile
constexpr int kThreads = kWarps * 32;
constexpr int kScratch = 32 * 33 * 2;     // floats: 32x33 float2 exchange buffer == 4 rows of kRow + 48
constexpr int kRow = 516;                 // 513 bins padded to a multiple of 4


// ---------------------------------------------------------------------------------------------
// 1024-point complex FFT across one warp.  In: lane l holds z[32*m + l] at position m.
// Out: position p holds Z[l + 32*brev5(p)].
__device__ __forceinline__ void fft1024_warp(float (&re)[32], float (&im)[32], float2* scratch,
                                             const float2* tw_s, int lane) {
    fft32(re, im);                                   // position p: Y_l[ka], ka = brev5(p)
    static_for<0, 32>([&](auto pi) {
        constexpr int p = decltype(pi)::value;
        constexpr int ka = brev5(p);
        float r = re[p], i = im[p];
        if constexpr (ka != 0) {                     // times W1024^(l*ka) = (c, -s)
            const float2 w = tw_s[ka * 32 + lane];
            const float tr = r * w.x - i * w.y;
            const float ti = r * w.y + i * w.x;
            r = tr; i = ti;
        }
        scratch[ka * 33 + lane] = make_float2(r, i);
    });
    __syncwarp();
    static_for<0, 32>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        const float2 v = scratch[lane * 33 + j];
        re[j] = v.x; im[j] = v.y;
    });
    __syncwarp();                                    // scratch is free again
    fft32(re, im);
}

// Window and load one channel pair of one frame.  sA/sB point at the frame's first staged sample.
template <bool kHasB>
__device__ __forceinline__ void load_pair(float (&re)[32], float (&im)[32], const float* sA,
                                          const float* sB, const float* win_s, int lane) {
    static_for<0, 32>([&](auto mi) {
        constexpr int m = decltype(mi)::value;
        const float w = win_s[32 * m + lane];
        re[m] = sA[32 * m + lane] * w;
        im[m] = kHasB ? sB[32 * m + lane] * w : 0.0f;
    });
}

// Split the packed transform Z = A + iB of two real channels at slot kb (bin k = lane + 32*kb).
// The analysis window was pre-scaled by 0.5, so A = Z[k] + conj(Z[N-k]) exactly.
template <int KB>
__device__ __forceinline__ void untangle(const float (&re)[32], const float (&im)[32], int lane,
                                         float& ar, float& ai, float& br, float& bi) {
    constexpr int p = brev5(KB & 31);
    const float zr = re[p], zi = im[p];
    float pr, pi;
    if constexpr (KB == 16) {                        // only lane 0 (bin 512) is meaningful: self-partner
        pr = zr; pi = zi;
    } else {
        constexpr int pp = brev5(31 - KB);           // partner slot in lane (32-l)&31, l != 0
        constexpr int p0 = brev5((32 - KB) & 31);    // lane 0: partner is in lane 0 itself
        const int src = (32 - lane) & 31;
        pr = __shfl_sync(0xffffffffu, re[pp], src);
        pi = __shfl_sync(0xffffffffu, im[pp], src);
        if (lane == 0) { pr = re[p0]; pi = im[p0]; }
    }
    ar = zr + pr; ai = zi - pi;
    br = zi + pi; bi = pr - zr;
}

// ---------------------------------------------------------------------------------------------
// Band-sparse mel projections.  Lane owns band m in round r: m = 32*r + (r odd ? 31-lane : lane),
// which pairs a narrow low band with a wide high band.
struct MelTab {
    const float* wt; const int* blo; const int* bcnt; const int* boff;
    int n_mels; float amin;
};

__device__ __forceinline__ int band_of(int r, int lane, int n_mels) {
    const int m = 32 * r + ((r & 1) ? 31 - lane : lane);
    return m < n_mels ? m : -1;
}

__device__ __forceinline__ float to_db(float v, float amin) {
    return 10.0f * log10f(fmaxf(v, amin));
}

// two power rows (interleaved float2 per bin) -> dB; o0/o1 may be null
__device__ __forceinline__ void mel_pow2(const float2* q, const MelTab& mt, int lane, float* o0, float* o1) {
    for (int r = 0; r * 32 < mt.n_mels; ++r) {
        const int m = band_of(r, lane, mt.n_mels);
        float a0 = 0.f, a1 = 0.f;
        if (m >= 0) {
            const int lo = mt.blo[m], cnt = mt.bcnt[m];
            const float* w = mt.wt + mt.boff[m];
            for (int i = 0; i < cnt; ++i) {
                const float wv = w[i];
                const float2 v = q[lo + i];
                a0 = fmaf(v.x, wv, a0); a1 = fmaf(v.y, wv, a1);
            }
            if (o0) o0[m] = to_db(a0, mt.amin);
            if (o1) o1[m] = to_db(a1, mt.amin);
        }
    }
}

// two power rows -> dB (o0, o1) and two linear rows (o2, o3) in one walk
__device__ __forceinline__ void mel_pow2_lin2(const float2* qa, const float2* qb, const MelTab& mt, int lane,
                                              float* o0, float* o1, float* o2, float* o3) {
    for (int r = 0; r * 32 < mt.n_mels; ++r) {
        const int m = band_of(r, lane, mt.n_mels);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (m >= 0) {
            const int lo = mt.blo[m], cnt = mt.bcnt[m];
            const float* w = mt.wt + mt.boff[m];
            for (int i = 0; i < cnt; ++i) {
                const float wv = w[i];
                const float2 u = qa[lo + i];
                const float2 v = qb[lo + i];
                a0 = fmaf(u.x, wv, a0); a1 = fmaf(u.y, wv, a1);
                a2 = fmaf(v.x, wv, a2); a3 = fmaf(v.y, wv, a3);
            }
            o0[m] = to_db(a0, mt.amin);
            o1[m] = to_db(a1, mt.amin);
            o2[m] = a2; o3[m] = a3;
        }
    }
}

// one linear row
__device__ __forceinline__ void mel_lin1(const float* q, const MelTab& mt, int lane, float* o) {
    for (int r = 0; r * 32 < mt.n_mels; ++r) {
        const int m = band_of(r, lane, mt.n_mels);
        float a = 0.f;
        if (m >= 0) {
            const int lo = mt.blo[m], cnt = mt.bcnt[m];
            const float* w = mt.wt + mt.boff[m];
            for (int i = 0; i < cnt; ++i) a = fmaf(q[lo + i], w[i], a);
            o[m] = a;
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <bool kIV>
__global__ void __launch_bounds__(kThreads, 2)
foa_features_kernel(const FoaArgs a, const PlanDev pd) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);                     // 1024 float2
    float* win_s = reinterpret_cast<float*>(tw_s + 1024);                   // 1024
    float* wt_s = win_s + 1024;                                             // nnz_pad
    int* blo_s = reinterpret_cast<int*>(wt_s + pd.nnz_pad);                 // n_mels_pad x3
    int* bcnt_s = blo_s + pd.n_mels_pad;
    int* boff_s = bcnt_s + pd.n_mels_pad;
    float* R_all = reinterpret_cast<float*>(boff_s + pd.n_mels_pad);        // kWarps * kScratch
    float* stage = R_all + kWarps * kScratch;                               // 4 * span

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-block constant tables (once: the block is persistent)
    for (int i = tid; i < 1024; i += kThreads) { tw_s[i] = pd.tw[i]; win_s[i] = pd.win[i]; }
    for (int i = tid; i < pd.nnz_pad; i += kThreads) wt_s[i] = pd.wt[i];
    for (int i = tid; i < pd.n_mels; i += kThreads) {
        blo_s[i] = pd.blo[i]; bcnt_s[i] = pd.bcnt[i]; boff_s[i] = pd.boff[i];
    }
    MelTab mt{wt_s, blo_s, bcnt_s, boff_s, pd.n_mels, pd.amin};

    float* R = R_all + warp * kScratch;
    float2* scratch = reinterpret_cast<float2*>(R);
    const int hop = pd.hop, span = a.span, M = pd.n_mels;

    // channel group handled by this block row
    const int c_base = a.c_lo + blockIdx.y * 4;
    const int nc = min(4, a.C - c_base);
    const bool do_iv = kIV && c_base == 0;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_clip;
        const int t0 = (tile - b * a.tiles_per_clip) * kWarps;
        const int nf = min(kWarps, a.T - t0);
        const int64_t s0 = (int64_t)t0 * hop - 512;                         // first staged sample
        const int valid = (nf - 1) * hop + 1024;                            // samples actually needed
        const float* xb = static_cast<const float*>(a.x) + (int64_t)b * a.stride_b + (int64_t)c_base * a.stride_c;

        __syncthreads();                                                    // previous tile's readers done
        if (a.vec_ok && s0 >= 0 && s0 + span <= a.L) {                      // interior: 16-byte loads
            const int nv = span >> 2;
            for (int c = 0; c < nc; ++c) {
                const float4* src = reinterpret_cast<const float4*>(xb + c * a.stride_c + s0);
                float4* dst = reinterpret_cast<float4*>(stage + c * span);
                for (int i = tid; i < nv; i += kThreads) dst[i] = __ldg(src + i);
            }
        } else {                                                            // edges: reflect padding
            for (int c = 0; c < nc; ++c) {
                const float* src = xb + c * a.stride_c;
                float* dst = stage + c * span;
                for (int i = tid; i < valid; i += kThreads) {
                    int64_t s = s0 + i;
                    if (s < 0) s = -s;
                    if (s >= a.L) s = 2 * (a.L - 1) - s;
                    dst[i] = __ldg(src + s);
                }
            }
        }
        __syncthreads();

        if (warp >= nf) continue;                                           // no frame for this warp
        const int t = t0 + warp;
        const float* sf = stage + warp * hop;                               // this frame's samples, ch 0
        float* ob = a.out + (((int64_t)b * a.Cout) * a.T + t) * M;          // + ch*T*M
        const int64_t ch_stride = (int64_t)a.T * M;

        float re[32], im[32];
        float Wr[17], Wi[17], Iy[17];

        // ---------------- channel pair (0, 1)
        if (nc >= 2) load_pair<true>(re, im, sf, sf + span, win_s, lane);
        else         load_pair<false>(re, im, sf, sf, win_s, lane);
        fft1024_warp(re, im, scratch, tw_s, lane);
        {
            float2* q = reinterpret_cast<float2*>(R);
            static_for<0, 17>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                float ar, ai, br, bi;
                untangle<kb>(re, im, lane, ar, ai, br, bi);
                const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
                if (kb < 16 || lane == 0) q[lane + 32 * kb] = make_float2(pa, pb);
                if (kIV) { Wr[kb] = ar; Wi[kb] = ai; Iy[kb] = ar * br + ai * bi; }
            });
            __syncwarp();
            mel_pow2(q, mt, lane, ob + (c_base + 0) * ch_stride, nc >= 2 ? ob + (c_base + 1) * ch_stride : nullptr);
            __syncwarp();
        }
        if (nc <= 2) continue;

        // ---------------- channel pair (2, 3)
        if (nc >= 4) load_pair<true>(re, im, sf + 2 * span, sf + 3 * span, win_s, lane);
        else         load_pair<false>(re, im, sf + 2 * span, sf, win_s, lane);
        fft1024_warp(re, im, scratch, tw_s, lane);
        float* o2 = ob + (c_base + 2) * ch_stride;
        float* o3 = nc >= 4 ? ob + (c_base + 3) * ch_stride : nullptr;
        if (!do_iv) {
            float2* q = reinterpret_cast<float2*>(R);
            static_for<0, 17>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                float ar, ai, br, bi;
                untangle<kb>(re, im, lane, ar, ai, br, bi);
                const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
                if (kb < 16 || lane == 0) q[lane + 32 * kb] = make_float2(pa, pb);
            });
            __syncwarp();
            mel_pow2(q, mt, lane, o2, o3);
            __syncwarp();
        } else {
            // intensity vector: I_j = Re(conj(W) P_j), j = ch1, ch2, ch3; n_j = I_j / (|I| + eps)
            float2* qa = reinterpret_cast<float2*>(R);                      // (P_2, P_3)
            float2* qb = qa + kRow;                                         // (n_1, n_2)
            const float eps = pd.eps;
            static_for<0, 17>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                float ar, ai, br, bi;
                untangle<kb>(re, im, lane, ar, ai, br, bi);
                const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
                const float i1 = Iy[kb];
                const float i2 = Wr[kb] * ar + Wi[kb] * ai;
                const float i3 = Wr[kb] * br + Wi[kb] * bi;
                const float nrm = sqrtf(i1 * i1 + i2 * i2 + i3 * i3) + eps;
                const float inv = 1.0f / nrm;
                if (kb < 16 || lane == 0) {
                    qa[lane + 32 * kb] = make_float2(pa, pb);
                    qb[lane + 32 * kb] = make_float2(i1 * inv, i2 * inv);
                }
                Iy[kb] = i3 * inv;
            });
            __syncwarp();
            float* iv0 = ob + (int64_t)a.C * ch_stride;
            mel_pow2_lin2(qa, qb, mt, lane, o2, o3, iv0, iv0 + ch_stride);
            __syncwarp();
            float* q1 = R;
            static_for<0, 17>([&](auto kbi) {
                constexpr int kb = decltype(kbi)::value;
                if (kb < 16 || lane == 0) q1[lane + 32 * kb] = Iy[kb];
            });
            __syncwarp();
            mel_lin1(q1, mt, lane, iv0 + 2 * ch_stride);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
size_t foa_smem_bytes(const PlanDev& pd, int span) {
    size_t f = 2 * 1024 + 1024 + pd.nnz_pad + 3 * pd.n_mels_pad + kWarps * kScratch + 4 * (size_t)span;
    return f * sizeof(float);
}

int foa_frames_per_tile() { return kWarps; }

template <bool kIV>
static cudaError_t launch_t(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    const size_t smem = foa_smem_bytes(pd, a.span);
    cudaError_t e = cudaFuncSetAttribute(foa_features_kernel<kIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int groups = (a.C - a.c_lo + 3) / 4;
    int gx = (2 * sm_count) / groups;
    if (gx < 1) gx = 1;
    if (gx > a.n_tiles) gx = a.n_tiles;
    dim3 grid(gx, groups);
    foa_features_kernel<kIV><<<grid, kThreads, smem, st>>>(a, pd);
    return cudaGetLastError();
}

cudaError_t foa_launch(bool iv, const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st) {
    return iv ? launch_t<true>(a, pd, sm_count, st) : launch_t<false>(a, pd, sm_count, st);
}

}  // namespace seld
