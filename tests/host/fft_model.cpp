// Host-side model of the warp-wide 1024-point transform of pseldnets_b200/csrc/seld_foa.cu.
// It compiles the SAME fft32.cuh (device qualifiers defined away) and replays, lane by lane, the
// index algebra the kernel uses: stride-32 load, 32-point FFT, W1024^(l*ka) twiddle, 32x33
// exchange, second 32-point FFT, bit-reversed slots, and the (32-l)&31 partner-lane untangle of
// two real channels packed into one complex transform.  Exposed through a tiny C ABI so the CPU
// test-suite can compare it with numpy's rfft -- no GPU needed to catch an indexing mistake.
#include <cmath>
#include <cstring>
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../pseldnets_b200/csrc/fft32.cuh"

using namespace seld;

// a, b: two real frames of 1024 samples (already windowed*0.5 upstream is NOT applied here; the
// caller passes what the kernel would have in registers).  Outputs A, B: 513 complex bins each,
// interleaved (re, im), equal to rfft(2a) and rfft(2b) -- i.e. the caller halves the input.
extern "C" void model_fft1024_pair(const float* a, const float* b, float* A, float* B) {
    static float re[32][32], im[32][32];
    static float sc_r[32 * 33], sc_i[32 * 33];
    for (int lane = 0; lane < 32; ++lane) {
        for (int m = 0; m < 32; ++m) { re[lane][m] = a[32 * m + lane]; im[lane][m] = b[32 * m + lane]; }
        fft32(re[lane], im[lane]);
        for (int p = 0; p < 32; ++p) {
            const int ka = brev5(p);
            float r = re[lane][p], i = im[lane][p];
            if (ka != 0) {
                const double ang = 2.0 * M_PI * (double)((ka * lane) % 1024) / 1024.0;
                const float c = (float)cos(ang), s = (float)(-sin(ang));
                const float tr = r * c - i * s, ti = r * s + i * c;
                r = tr; i = ti;
            }
            sc_r[ka * 33 + lane] = r; sc_i[ka * 33 + lane] = i;
        }
    }
    for (int lane = 0; lane < 32; ++lane) {
        for (int j = 0; j < 32; ++j) { re[lane][j] = sc_r[lane * 33 + j]; im[lane][j] = sc_i[lane * 33 + j]; }
        fft32(re[lane], im[lane]);
    }
    for (int lane = 0; lane < 32; ++lane) {
        for (int kb = 0; kb <= 16; ++kb) {
            if (kb == 16 && lane != 0) continue;
            const int p = brev5(kb & 31);
            const float zr = re[lane][p], zi = im[lane][p];
            float pr, pi;
            if (kb == 16) { pr = zr; pi = zi; }
            else if (lane == 0) { const int p0 = brev5((32 - kb) & 31); pr = re[0][p0]; pi = im[0][p0]; }
            else { const int pp = brev5(31 - kb); const int src = (32 - lane) & 31; pr = re[src][pp]; pi = im[src][pp]; }
            const int k = lane + 32 * kb;
            A[2 * k] = zr + pr; A[2 * k + 1] = zi - pi;
            B[2 * k] = zi + pi; B[2 * k + 1] = pr - zr;
        }
    }
}

extern "C" void model_fft32(float* re, float* im) {   // in place, natural order out
    float r[32], i[32], orr[32], oi[32];
    memcpy(r, re, sizeof r); memcpy(i, im, sizeof i);
    fft32(r, i);
    for (int p = 0; p < 32; ++p) { orr[brev5(p)] = r[p]; oi[brev5(p)] = i[p]; }
    memcpy(re, orr, sizeof r); memcpy(im, oi, sizeof i);
}

// ---- second/third generation kernels: packed transforms -------------------------------------

// fft32_dit: in position p = (x[2p], x[2p+1]); out natural order re/im[32]
extern "C" void model_fft32_dit(float* re, float* im) {
    float2 pr[16], pi[16];
    for (int p = 0; p < 16; ++p) { pr[p] = make_float2(re[2 * p], re[2 * p + 1]); pi[p] = make_float2(im[2 * p], im[2 * p + 1]); }
    fft32_dit(pr, pi);
    for (int qp = 0; qp < 16; ++qp) {
        const int q = brev4(qp);
        re[q] = pr[qp].x; re[q + 16] = pr[qp].y; im[q] = pi[qp].x; im[q + 16] = pi[qp].y;
    }
}

// iv3 warp algebra for one channel pair (a, b): DIT-split 32-pt, packed W1024 twiddle pairs
// (ka, ka+16), 32x34 plane exchange read back as (Y_2p, Y_2p+1) pairs, DIT-split 32-pt,
// partner = .y half of position brev4(15-kb) in lane (32-l)&31.  Same outputs as above.
extern "C" void model_iv3_pair(const float* a, const float* b, float* A, float* B) {
    static float2 pr[32][16], pi[32][16];
    static float mre[32 * 34], mim[32 * 34];
    for (int lane = 0; lane < 32; ++lane) {
        for (int p = 0; p < 16; ++p) {
            pr[lane][p] = make_float2(a[32 * (2 * p) + lane], a[32 * (2 * p + 1) + lane]);
            pi[lane][p] = make_float2(b[32 * (2 * p) + lane], b[32 * (2 * p + 1) + lane]);
        }
        fft32_dit(pr[lane], pi[lane]);
        for (int qp = 0; qp < 16; ++qp) {
            const int q = brev4(qp);
            float c[2], s[2];
            for (int h = 0; h < 2; ++h) {
                const double ang = 2.0 * M_PI * (double)(((q + 16 * h) * lane) % 1024) / 1024.0;
                c[h] = (float)cos(ang); s[h] = (float)sin(ang);
            }
            const float2 c2 = make_float2(c[0], c[1]), s2 = make_float2(s[0], s[1]);
            const float2 r = pr[lane][qp], i = pi[lane][qp];
            const float2 nr = __ffma2_rn(i, s2, __fmul2_rn(r, c2));
            const float2 ni = __ffma2_rn(r, make_float2(-s2.x, -s2.y), __fmul2_rn(i, c2));
            mre[q * 34 + lane] = nr.x; mre[(q + 16) * 34 + lane] = nr.y;
            mim[q * 34 + lane] = ni.x; mim[(q + 16) * 34 + lane] = ni.y;
        }
    }
    for (int lane = 0; lane < 32; ++lane) {
        for (int p = 0; p < 16; ++p) {
            pr[lane][p] = make_float2(mre[lane * 34 + 2 * p], mre[lane * 34 + 2 * p + 1]);
            pi[lane][p] = make_float2(mim[lane * 34 + 2 * p], mim[lane * 34 + 2 * p + 1]);
        }
        fft32_dit(pr[lane], pi[lane]);
    }
    for (int lane = 0; lane < 32; ++lane) {
        for (int kb = 0; kb <= 16; ++kb) {
            if (kb == 16 && lane != 0) continue;
            float zr, zi, qr, qi;
            if (kb == 16) { zr = pr[0][brev4(0)].y; zi = pi[0][brev4(0)].y; qr = zr; qi = zi; }
            else {
                zr = pr[lane][brev4(kb)].x; zi = pi[lane][brev4(kb)].x;
                if (lane != 0) { const int src = (32 - lane) & 31; qr = pr[src][brev4(15 - kb)].y; qi = pi[src][brev4(15 - kb)].y; }
                else if (kb == 0) { qr = zr; qi = zi; }
                else { qr = pr[0][brev4(16 - kb)].y; qi = pi[0][brev4(16 - kb)].y; }
            }
            const int k = lane + 32 * kb;
            A[2 * k] = zr + qr; A[2 * k + 1] = zi - qi;
            B[2 * k] = zi + qi; B[2 * k + 1] = qr - zr;
        }
    }
}
