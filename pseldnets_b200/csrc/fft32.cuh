// 32-point complex FFT held entirely in one thread's registers, and the small compile-time
// helpers around it.  Two of these per thread (with a 32x32 exchange through shared memory in
// between) make the warp-wide 1024-point transform of the feature kernels.
//
// Radix-2 decimation in frequency, fully unrolled with compile-time twiddles, so every
// twiddle is an immediate operand and the trivial ones (1, -i, (1-i)/sqrt2 ...) cost adds only.
// Output is left in bit-reversed register order: after fft32(), position p holds X[brev5(p)];
// since every register index is a compile-time constant the permutation is free.
//
// The element type V is either `float` (one transform) or `float2` (two independent transforms
// in the two halves of a packed register pair: on sm_100a add/mul/fma of a float2 is ONE
// FADD2/FMUL2/FFMA2 instruction, which halves the issue slots the butterflies need).
#pragma once
#include <utility>

#ifndef __CUDACC__
// Host build (tests/host/fft_model.cpp compiles this header with g++ to replay the kernels' index
// algebra on the CPU): stand-ins for the CUDA vector type and the sm_100a packed-fp32 intrinsics.
#include <cmath>
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif

namespace seld {

#ifdef __CUDACC__
// log2 of a normal positive number in one MUFU (__log2f adds a compare, a scale and a correction for subnormal
// arguments; the kernels only take it of max(v, amin) with amin >= FLT_MIN)
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int brev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos/sin(2*pi*k/32), k = 0..8, correctly rounded from double.
__host__ __device__ constexpr double cos32_q(int k) {
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.98078528040323044913;
        case 2: return 0.92387953251128675613;
        case 3: return 0.83146961230254523708;
        case 4: return 0.70710678118654752440;
        case 5: return 0.55557023301960222474;
        case 6: return 0.38268343236508977173;
        case 7: return 0.19509032201612826785;
        default: return 0.0;
    }
}
__host__ __device__ constexpr double cos32(int k) {   // k in [0, 16]
    return k <= 8 ? cos32_q(k) : -cos32_q(16 - k);
}
__host__ __device__ constexpr double sin32(int k) {   // k in [0, 16]
    return k <= 8 ? cos32_q(8 - k) : cos32_q(k - 8);
}

// ---- element ops: scalar
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vmuls(float a, float s) { return a * s; }                   // a*s
__device__ __forceinline__ float vfmas(float a, float s, float c) { return a * s + c; }      // a*s + c

// ---- element ops: two transforms packed in a float2 (FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 vmuls(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 vfmas(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

// (r + i*im) *= exp(-2*pi*i*K/32), K in [0, 16)
template <int K, class V>
__device__ __forceinline__ void mul_w32(V& r, V& i) {
    if constexpr (K == 0) {
    } else if constexpr (K == 8) {
        const V t = r; r = i; i = vneg(t);
    } else if constexpr (K == 4) {
        constexpr float h = (float)cos32(4);
        const V a = vadd(r, i), b = vsub(i, r);
        r = vmuls(a, h); i = vmuls(b, h);
    } else if constexpr (K == 12) {
        constexpr float h = (float)cos32(4);
        const V a = vsub(i, r), b = vadd(r, i);
        r = vmuls(a, h); i = vmuls(b, -h);
    } else {
        // r' = r*c + i*s, i' = i*c - r*s, written so each result can land in its own register
        constexpr float c = (float)cos32(K), s = (float)sin32(K);
        const V a = vmuls(i, s), b = vmuls(r, -s);
        r = vfmas(r, c, a);
        i = vfmas(i, c, b);
    }
}

// One DIF stage: butterflies of span HALF inside groups of 2*HALF.
template <int HALF, class V>
__device__ __forceinline__ void dif_stage(V (&re)[32], V (&im)[32]) {
    constexpr int STEP = 16 / HALF;            // twiddle exponent step in units of W32
    static_for<0, 16>([&](auto bi) {
        constexpr int b = decltype(bi)::value;
        constexpr int g = b / HALF, k = b % HALF;
        constexpr int i0 = g * 2 * HALF + k, i1 = i0 + HALF;
        const V ur = re[i0], ui = im[i0], vr = re[i1], vi = im[i1];
        re[i0] = vadd(ur, vr); im[i0] = vadd(ui, vi);
        V dr = vsub(ur, vr), di = vsub(ui, vi);
        mul_w32<k * STEP>(dr, di);
        re[i1] = dr; im[i1] = di;
    });
}

// ---- decimation in time with natural-order input and bit-reversed output: the twiddle of a butterfly depends on its
// group only, and multiplies the second input BEFORE the add/subtract.  That form fuses: with w = c (1 - i t), t = tan,
//   b w = c (br + t bi, bi - t br),   a +- b w = a +- c * tmp
// is 2 + 4 fused multiply-adds per butterfly instead of 4 adds + 4 multiplies (cotangent form where |c| < |s|); the
// (1 -+ i) / sqrt2 twiddles take 2 adds + 4 FMAs instead of 6 adds + 2 multiplies.  68 fewer operations per 32-point
// transform than the decimation-in-frequency form above (388 instead of 456), one rounding less per output.
__host__ __device__ constexpr int brev_n(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int x) { int r = 0; while (x > 1) { x >>= 1; ++r; } return r; }

// (ar, ai), (br, bi) -> a + b W32^K, a - b W32^K        K in [0, 16)
template <int K, class V>
__device__ __forceinline__ void dit_bfly(V& ar, V& ai, V& br, V& bi) {
    if constexpr (K == 0) {
        const V r0 = vadd(ar, br), i0 = vadd(ai, bi), r1 = vsub(ar, br), i1 = vsub(ai, bi);
        ar = r0; ai = i0; br = r1; bi = i1;
    } else if constexpr (K == 8) {                 // b * (-i) = (bi, -br)
        const V r0 = vadd(ar, bi), i0 = vsub(ai, br), r1 = vsub(ar, bi), i1 = vadd(ai, br);
        ar = r0; ai = i0; br = r1; bi = i1;
    } else if constexpr (K == 4) {                 // b * h (1 - i) = h (br + bi, bi - br)
        constexpr float h = (float)cos32(4);
        const V tr = vadd(br, bi), ti = vsub(bi, br);
        const V r0 = vfmas(tr, h, ar), i0 = vfmas(ti, h, ai), r1 = vfmas(tr, -h, ar), i1 = vfmas(ti, -h, ai);
        ar = r0; ai = i0; br = r1; bi = i1;
    } else if constexpr (K == 12) {                // b * (-h) (1 + i) = -h (br - bi, br + bi)
        constexpr float h = (float)cos32(4);
        const V tr = vsub(br, bi), ti = vadd(br, bi);
        const V r0 = vfmas(tr, -h, ar), i0 = vfmas(ti, -h, ai), r1 = vfmas(tr, h, ar), i1 = vfmas(ti, h, ai);
        ar = r0; ai = i0; br = r1; bi = i1;
    } else {
        // W32^K = c - i s:  b w = (br c + bi s, bi c - br s)
        constexpr double cd = cos32(K), sd = sin32(K);
        if constexpr ((cd < 0 ? -cd : cd) >= sd) {  // = c (br + t bi, bi - t br)
            constexpr float t = (float)(sd / cd), c = (float)cd;
            const V tr = vfmas(bi, t, br), ti = vfmas(br, -t, bi);
            const V r0 = vfmas(tr, c, ar), i0 = vfmas(ti, c, ai), r1 = vfmas(tr, -c, ar), i1 = vfmas(ti, -c, ai);
            ar = r0; ai = i0; br = r1; bi = i1;
        } else {                                   // = s (u br + bi, u bi - br),  u = c / s
            constexpr float u = (float)(cd / sd), sn = (float)sd;
            const V tr = vfmas(br, u, bi), ti = vfmas(bi, u, vneg(br));
            const V r0 = vfmas(tr, sn, ar), i0 = vfmas(ti, sn, ai), r1 = vfmas(tr, -sn, ar), i1 = vfmas(ti, -sn, ai);
            ar = r0; ai = i0; br = r1; bi = i1;
        }
    }
}

// One DIT stage: butterflies of span HALF inside groups of 2*HALF; group g uses W32^(HALF * brev(g)).
template <int HALF, class V>
__device__ __forceinline__ void dit_stage(V (&re)[32], V (&im)[32]) {
    static_for<0, 16>([&](auto bi_) {
        constexpr int b = decltype(bi_)::value;
        constexpr int g = b / HALF, k = b % HALF;
        constexpr int i0 = g * 2 * HALF + k, i1 = i0 + HALF;
        constexpr int K = HALF * brev_n(g, ilog2(16 / HALF));
        dit_bfly<K>(re[i0], im[i0], re[i1], im[i1]);
    });
}

// In-place forward 32-point DFT; result X[brev5(p)] at position p.
template <class V>
__device__ __forceinline__ void fft32(V (&re)[32], V (&im)[32]) {
#ifdef SELD_FFT_DIF
    dif_stage<16>(re, im);
    dif_stage<8>(re, im);
    dif_stage<4>(re, im);
    dif_stage<2>(re, im);
    dif_stage<1>(re, im);
#else
    dit_stage<16>(re, im);
    dit_stage<8>(re, im);
    dit_stage<4>(re, im);
    dit_stage<2>(re, im);
    dit_stage<1>(re, im);
#endif
}


// ---------------------------------------------------------------------------------------------
// The same 32-point DFT for ONE transform, packed internally ("split" form, decimation in time):
// the even- and odd-indexed inputs are two independent 16-point transforms E = FFT16(x[2p]) and
// O = FFT16(x[2p+1]); they run in the two halves of float2 registers (same twiddles in both
// halves).  The last radix-2 stage X[q] = E[q] + W32^q O[q], X[q+16] = E[q] - W32^q O[q] is one
// scalar complex multiply plus ONE FFMA2 per component: (e + t, e - t) = (1, -1) * t + e with t
// and e as broadcast scalar operands.
//   in : position p  holds the input pair  (x[2p], x[2p+1])
//   out: position q' holds the output pair (X[q], X[q+16]),  q = brev4(q')
__host__ __device__ constexpr int brev4(int x) {
    return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3);
}

template <int HALF>
__device__ __forceinline__ void dif16_stage(float2 (&re)[16], float2 (&im)[16]) {
    constexpr int STEP = 16 / HALF;            // twiddle exponent step in units of W32 (W16^k = W32^2k)
    static_for<0, 8>([&](auto bi) {
        constexpr int b = decltype(bi)::value;
        constexpr int g = b / HALF, k = b % HALF;
        constexpr int i0 = g * 2 * HALF + k, i1 = i0 + HALF;
        const float2 ur = re[i0], ui = im[i0], vr = re[i1], vi = im[i1];
        re[i0] = vadd(ur, vr); im[i0] = vadd(ui, vi);
        float2 dr = vsub(ur, vr), di = vsub(ui, vi);
        mul_w32<k * STEP>(dr, di);
        re[i1] = dr; im[i1] = di;
    });
}

__device__ __forceinline__ void fft32_dit(float2 (&pr)[16], float2 (&pi)[16]) {
    dif16_stage<8>(pr, pi);
    dif16_stage<4>(pr, pi);
    dif16_stage<2>(pr, pi);
    dif16_stage<1>(pr, pi);                    // position q': (E[q], O[q]), q = brev4(q')
    const float2 pm = make_float2(1.0f, -1.0f);
    static_for<0, 16>([&](auto qi) {
        constexpr int qp = decltype(qi)::value;
        constexpr int q = brev4(qp);
        float tr = pr[qp].y, ti = pi[qp].y;
        mul_w32<q>(tr, ti);                    // t = W32^q * O[q]
        const float er = pr[qp].x, ei = pi[qp].x;
        pr[qp] = __ffma2_rn(pm, make_float2(tr, tr), make_float2(er, er));
        pi[qp] = __ffma2_rn(pm, make_float2(ti, ti), make_float2(ei, ei));
    });
}

}  // namespace seld
