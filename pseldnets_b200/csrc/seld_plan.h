// Internal (not exported) declarations shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace seld {

// Device-resident constant tables of one extractor configuration ("plan").
struct PlanDev {
    const float2* tw;     // [32][32] (cos, -sin)(2*pi*ka*j/1024) at [ka*32 + j]
    const float* win;     // [n_fft] analysis window * 0.5
    const float* wt;      // band-sparse mel weights, band after band
    const int* blo;       // [n_mels] first bin of the band's support
    const int* bcnt;      // [n_mels] bins in the support
    const int* boff;      // [n_mels] offset of the band's weights in wt
    int nnz_pad;          // floats in wt (multiple of 4)
    int n_mels, n_mels_pad;
    int hop;
    float amin, eps;
};

struct FoaArgs {
    const float* x;          // (B, C, L) fp32, strides in elements
    int64_t stride_b, stride_c;
    float* out;              // (B, Cout, T, M) contiguous
    int64_t L;
    int B, C, Cout, T;
    int tiles_per_clip, n_tiles;
    int span;                // staged samples per channel per tile (multiple of 4)
    int vec_ok;              // x base/strides allow 16-byte loads
};

size_t foa_smem_bytes(const PlanDev& pd, int span);
int foa_frames_per_tile();
cudaError_t foa_launch(bool iv, const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);

}  // namespace seld
