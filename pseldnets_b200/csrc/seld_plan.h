// Internal (not exported) declarations shared by the kernels and the C-ABI layer.
#pragma once
#include "../../include/seldfeat.h"
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace seld {

// Device-resident constant tables of one extractor configuration ("plan").
struct PlanDev {
    const float2* tw;     // [32][32] (cos, -sin)(2*pi*ka*j/1024) at [ka*32 + j]
    const float* win;     // [n_fft] analysis window * 0.5
    const float4* tw4;    // iv3: [16][32] (cos q, cos q+16, sin q, sin q+16)(2*pi*ka*lane/1024)
    const float2* win2;   // iv3: [16][32] 0.5 * (w[32*2p + lane], w[32*(2p+1) + lane])
    const float* wt;      // band-sparse mel weights, band after band
    const int* blo;       // [n_mels] first bin of the band's support
    const int* bcnt;      // [n_mels] bins in the support
    const int* boff;      // [n_mels] offset of the band's weights in wt
    // segment form of the bank for the iv2 kernel (valid when fast_ok): every bin k feeds only
    // bands s_k-1 and s_k with weights (a_k, b_k); lane c owns bins [16c, 16c+16) (+512 for c=31)
    const float* wab;     // [32][36] (a, b) pairs per lane, 17 used
    const uint32_t* runmask;  // [32] bit j: a new run (segment change) starts at the lane's bin j
    const int* g0;        // [32] index of the lane's first run
    const int* gseg;      // [n_mels + 2] first run of segment s; runs of s are [gseg[s], gseg[s+1])
    // item form of the bank for the iv2 kernel's main path (valid when item_ok): every segment is cut into pieces of at most
    // iL[class] bins; lane l works through one piece per class (positions ioff[c] .. ioff[c] + iL[c], read from bins
    // istart[c][l] + j of the rows), so all lanes share one instruction stream with no run boundaries inside it (outside its
    // piece a lane's weights are zero)
    const float2* iw;     // [iP][32] (a, b) of position p of lane l
    const int* istart;    // [4 classes][32 lanes] first bin lane l reads in class c
    const uint32_t* islot;// [n_mels][2] slot lists, 4 x 8 bits each: pieces of segment m (their V sums) and of segment m + 1 (U sums)
    int item_ok, iP, iK;
    int iL[4], ioff[4];
    int gseg_pad;         // ints reserved for gseg in shared memory (multiple of 4)
    int fast_ok;          // bank has the segment structure and the run table fits
    int nnz_pad;          // floats in wt (multiple of 4)
    int n_mels, n_mels_pad;
    int hop;
    float amin, eps;
};

// Mel bank as tensor-core B operand (iv5 kernel): 33 chunks of 16 bins.  Chunk c is a K-major, unswizzled bf16
// tile [N_c columns x 16 bins] holding only the bands the chunk touches, hi and lo halves of each weight interleaved
// along N (column 2j = hi of band col0_c + j, column 2j + 1 = lo), so one MMA per chunk lands in accumulator
// columns [2 col0_c, 2 col0_c + N_c).  Chunk 0 is stored full width (N = 2 * 64): it overwrites the accumulator.
struct MelTiles {
    const uint4* b_img;      // device image of all tiles
    int b_bytes;             // multiple of 16
    int ok;                  // bank fits this form (n_mels == 64, tiles fit shared memory)
    uint32_t chunk[33];      // tile offset in 16-byte units | first accumulator column << 16 | (N_c >> 3) << 24
};

struct FoaArgs {
    const void* x;           // (B, C, L) fp32 (or int16 PCM when in_i16), strides in elements
    float in_scale;          // 1 for fp32 input, 2^-15 for int16 PCM
    int in_i16;
    int64_t stride_b, stride_c;
    float* out;              // (B, Cout, T, M) contiguous
    int64_t L;
    int B, C, Cout, T;       // C = channels of x, Cout = channels of out
    int c_lo;                // first input channel this launch covers (log-mel of channels c_lo..C-1)
    int tiles_per_clip, n_tiles;
    int step_clip, step_tile;  // iv2: grid size split as step_clip * tiles_per_clip + step_tile (set by the launcher)
    int redo_grid;             // redo form: blocks of the main grid (redo block i looks through the frames main block i processed)
    int* redo_flags;           // [0] blocks of the main grid that have finished, [1] some block marked a frame; both zero between launches
    int smem_bytes;            // dynamic shared memory of the launch (the device-side launch of the redo form needs it)
    void* spec;              // MIC only: (B, T, 513, 4) complex64 spectrogram, written (spectrogram mode) or read (from-spectra mode)
    int span;                // staged samples per channel per tile (multiple of 4)
    int vec_ok;              // x base/strides allow 16-byte loads
};

size_t foa_smem_bytes(const PlanDev& pd, int span);
int foa_frames_per_tile();
cudaError_t foa_launch(bool iv, const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);

// second-generation 4-channel log-mel + IV kernel (seld_foa_iv2.cu)
bool foa_iv2_supported(const PlanDev& pd, size_t smem_optin);
int foa_iv2_frames_per_tile();
cudaError_t foa_iv2_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);
int foa_lm4_jobs_per_tile();
cudaError_t foa_lm4_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);

// tensor-core generation: fp32 transform + tcgen05 mel projection (seld_foa_iv5.cu)
bool foa_iv5_supported(const PlanDev& pd, const MelTiles& mt, size_t smem_optin);
int foa_iv5_frames_per_tile();
cudaError_t foa_iv5_launch(const FoaArgs& a, const PlanDev& pd, const MelTiles& mt, int sm_count, cudaStream_t st);

// third-generation kernel: two warps per frame (seld_foa_iv3.cu)
bool foa_iv3_supported(const PlanDev& pd, size_t smem_optin);
int foa_iv3_frames_per_tile();
cudaError_t foa_iv3_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);


// MIC kernel: log-mel + GCC-PHAT of 4 microphones (seld_mic.cu)
bool mic_supported(const PlanDev& pd, size_t smem_optin);
int mic_frames_per_tile();
cudaError_t mic_launch(const FoaArgs& a, const PlanDev& pd, int* maxkey, float top_db, bool use_top_db,
                       int sm_count, cudaStream_t st, bool from_spectra = false);
cudaError_t mic_spectrogram_launch(const FoaArgs& a, const PlanDev& pd, int sm_count, cudaStream_t st);

// backbone-input stage (seld_epilogue.cu): eval-mode BatchNorm "scalar" terms, each (C, M) on the device;
// mean == nullptr means "no scalar" (identity)
struct ScalarArgs {
    const float* mean;
    const float* var;
    const float* weight;   // nullptr: 1
    const float* bias;     // nullptr: 0
    float eps;
};
cudaError_t scalar_launch(float* x, const ScalarArgs& s, int64_t B, int C, int T, int M, int sm_count, cudaStream_t st);
cudaError_t scalar_wav2img_launch(const float* x, float* img, const ScalarArgs& s, int64_t B, int C, int T, int M, int S,
                                  int sm_count, cudaStream_t st);

// waveform-domain augmentation (seld_augment.cu)
cudaError_t foa_rotate_launch(float* x, int64_t B, int64_t L, int64_t stride_b, int64_t stride_c, const int32_t* codes,
                              cudaStream_t st);
cudaError_t wavmix_launch(float* x, int C, int64_t L, int64_t stride_b, int64_t stride_c, const struct seld_mix_op* ops,
                          int n_ops, cudaStream_t st);

}  // namespace seld
