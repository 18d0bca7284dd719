// 32-point complex FFT held entirely in one thread's registers, and the small compile-time
// helpers around it.  Two of these per thread (with a 32x32 exchange through shared memory in
// between) make the warp-wide 1024-point transform of seld_foa.cu.
//
// Radix-2 decimation in frequency, fully unrolled with compile-time twiddles, so every
// twiddle is an immediate operand and the trivial ones (1, -i, (1-i)/sqrt2 ...) cost adds only.
// Output is left in bit-reversed register order: after fft32(), position p holds X[brev5(p)];
// since every register index is a compile-time constant the permutation is free.
#pragma once
#include <utility>

namespace seld {

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

// (r + i*im) *= exp(-2*pi*i*K/32), K in [0, 16)
template <int K>
__device__ __forceinline__ void mul_w32(float& r, float& i) {
    if constexpr (K == 0) {
    } else if constexpr (K == 8) {
        const float t = r; r = i; i = -t;
    } else if constexpr (K == 4) {
        constexpr float h = (float)cos32(4);
        const float a = r + i, b = i - r;
        r = a * h; i = b * h;
    } else if constexpr (K == 12) {
        constexpr float h = (float)cos32(4);
        const float a = i - r, b = r + i;
        r = a * h; i = -(b * h);
    } else {
        constexpr float c = (float)cos32(K), s = (float)sin32(K);
        const float tr = r * c + i * s;
        const float ti = i * c - r * s;
        r = tr; i = ti;
    }
}

// One DIF stage: butterflies of span HALF inside groups of 2*HALF.
template <int HALF>
__device__ __forceinline__ void dif_stage(float (&re)[32], float (&im)[32]) {
    constexpr int STEP = 16 / HALF;            // twiddle exponent step in units of W32
    static_for<0, 16>([&](auto bi) {
        constexpr int b = decltype(bi)::value;
        constexpr int g = b / HALF, k = b % HALF;
        constexpr int i0 = g * 2 * HALF + k, i1 = i0 + HALF;
        const float ur = re[i0], ui = im[i0], vr = re[i1], vi = im[i1];
        re[i0] = ur + vr; im[i0] = ui + vi;
        float dr = ur - vr, di = ui - vi;
        mul_w32<k * STEP>(dr, di);
        re[i1] = dr; im[i1] = di;
    });
}

// In-place forward 32-point DFT; result X[brev5(p)] at position p.
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
    dif_stage<16>(re, im);
    dif_stage<8>(re, im);
    dif_stage<4>(re, im);
    dif_stage<2>(re, im);
    dif_stage<1>(re, im);
}

}  // namespace seld
