/* seldfeat.h -- C ABI of the B200-native SELD feature front-end (libseldfeat.so).
 *
 * Drop-in boundary for the hot path of Jinbo-Hu/PSELDNets: the waveform -> feature-map
 * extractors of /root/reference/src/utils/feature.py.  The reference has no FFI (it is pure
 * Python on torchaudio); these are the entry points a binding for that path would call.  Every
 * function cites the reference interface it stands in for.
 *
 * Conventions: plain pointers and sizes only (no torch types); device pointers are owned by the
 * caller (e.g. the PyTorch allocator); compute entry points never allocate, never synchronise
 * and never throw -- they enqueue on `stream` (a cudaStream_t passed as void*) and return 0 or a
 * negative SELD_E* code.  All tensors are fp32.
 */
#ifndef SELDFEAT_H_
#define SELDFEAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELD_OK 0
#define SELD_EINVAL (-1)      /* bad argument (null pointer, non-positive size, C < 4 for IV ...) */
#define SELD_EUNSUPPORTED (-2) /* configuration outside what the kernels implement (n_fft != 1024, hop too large) */
#define SELD_ESHORT (-3)      /* clip too short for reflect padding: torch.stft needs n_fft/2 < L */
#define SELD_ECUDA (-4)       /* a CUDA runtime call failed; see seld_last_cuda_error() */
#define SELD_ENOMEM (-5)

typedef struct seld_plan seld_plan;

/* Build the constant tables of one extractor configuration on `device`.
 * Stands in for LogmelIV_Extractor.__init__ / Logmel_Extractor.__init__ (feature.py:21-37,
 * 60-75): `window_host` is the (n_fft,) analysis window the reference keeps as
 * `stft_extractor.window`, `fb_host` the (n_fft/2+1, n_mels) row-major mel bank it keeps as
 * `mel_scale.fb`; amin is AmplitudeToDB's clamp (1e-10; values below FLT_MIN are raised to it), eps the intensity-vector epsilon
 * (feature.py:8).  Allocates device memory; not on the per-step path. */
int seld_plan_create(seld_plan** plan, int device, const float* window_host, const float* fb_host,
                     int n_fft, int hop, int n_mels, float amin, float eps);
void seld_plan_destroy(seld_plan* plan);

/* Frames produced for a clip of L samples: 1 + L / hop (torch.stft, center=True). */
int64_t seld_num_frames(const seld_plan* plan, int64_t L);

/* LogmelIV_Extractor.forward (feature.py:39-56): x (B, C>=4, L) device fp32 with element
 * strides (stride_b, stride_c, 1) -> out (B, C+3, T, n_mels) contiguous device fp32:
 * C log-mel maps followed by the 3 mel-projected normalised intensity-vector maps of channels
 * 1..3 against channel 0 (intensityvector, feature.py:93-117). */
int seld_logmel_iv_f32(const seld_plan* plan, const float* x, int64_t B, int C, int64_t L,
                       int64_t stride_b, int64_t stride_c, float* out, void* stream);

/* Same with 16-bit PCM input: x holds the int16 samples a wav/flac decoder yields;
 * the kernel converts exactly as soundfile's float32 read does (s / 32768), so the result is
 * bit-identical to seld_logmel_iv_f32 on the converted waveform while reading half the bytes.
 * C must be 4. */
int seld_logmel_iv_i16(const seld_plan* plan, const int16_t* x, int64_t B, int C, int64_t L,
                       int64_t stride_b, int64_t stride_c, float* out, void* stream);

/* Same computation with HOST buffers: x_host (B, C, L) contiguous and out_host (B, C+3, T, n_mels)
 * contiguous, ideally page-locked.  The batch is cut into chunks of `chunk_clips` clips (<= 0:
 * library default) that flow through three plan-owned device slots on three internal streams, so
 * the host->device copy of chunk i+1, the kernel of chunk i and the device->host copy of chunk i-1
 * overlap (PCIe is full duplex).  Ordered after prior work on `stream` and before later work on it;
 * returns once everything is enqueued.  Device slots are allocated on first use and kept in the
 * plan (this is the one compute entry point that may allocate, on its first call per size). */
int seld_logmel_iv_f32_host(seld_plan* plan, const float* x_host, int64_t B, int C, int64_t L,
                            float* out_host, int chunk_clips, void* stream);
/* ... and with int16 PCM host input (half the host->device bytes). */
int seld_logmel_iv_i16_host(seld_plan* plan, const int16_t* x_host, int64_t B, int C, int64_t L,
                            float* out_host, int chunk_clips, void* stream);

/* Logmel_Extractor.forward (feature.py:76-91): x (B, C>=1, L) -> out (B, C, T, n_mels). */
int seld_logmel_f32(const seld_plan* plan, const float* x, int64_t B, int C, int64_t L,
                    int64_t stride_b, int64_t stride_c, float* out, void* stream);

/* MIC features: Features_Extractor_MIC._spectrogram / _get_logmel_spectrogram / _get_gcc
 * (feature.py:146-175) assembled as preprocess.py:546-556.  x (B, C=4, L) -> out
 * (B, 4 + 6, T, n_mels), T = L / hop: 4 log-mel planes (librosa.power_to_db: amin clamp, then a
 * floor at the plane's maximum - top_db; pass top_db < 0 for "None") followed by the 6
 * GCC-PHAT planes of mic pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), lags [-n_mels/2, n_mels/2).
 * Frames use librosa.stft's zero ('constant') centre padding.  The plan's fb is the mel bank
 * (librosa.filters.mel(sr, n_fft, n_mels).T).  `workspace` is device scratch of at least
 * seld_workspace_bytes(plan, B, C) bytes (per-plane running maxima and minima). */
int64_t seld_num_frames_mic(const seld_plan* plan, int64_t L);
size_t seld_workspace_bytes(const seld_plan* plan, int64_t B, int C);
int seld_logmel_gcc_f32(const seld_plan* plan, const float* x, int64_t B, int C, int64_t L,
                        int64_t stride_b, int64_t stride_c, float top_db, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* The same path in the reference's three stages, for callers that drive them separately as
 * Preprocess.extract_mic_features does (preprocess.py:546-556):
 * seld_mic_spectrogram_f32           Features_Extractor_MIC._spectrogram (feature.py:146-153): x (B, C=4, L) ->
 *                                    spec (B, T, n_fft/2+1, C) complex64 (interleaved re, im; 16-byte aligned),
 *                                    i.e. the reference's (T, F, C) array per clip, T = L / hop;
 * seld_logmel_gcc_from_spectra_f32   _get_logmel_spectrogram + _get_gcc (feature.py:155-175) of such an array ->
 *                                    out (B, 4 + 6, T, n_mels) as seld_logmel_gcc_f32 lays it out. */
int seld_mic_spectrogram_f32(const seld_plan* plan, const float* x, int64_t B, int C, int64_t L,
                             int64_t stride_b, int64_t stride_c, float* spec, void* stream);
int seld_logmel_gcc_from_spectra_f32(const seld_plan* plan, const float* spec, int64_t B, int C, int64_t T,
                                     float top_db, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream);

/* ---- Backbone-input stage: what every backbone does first with the feature map.
 *
 * seld_scalar_f32: the per-channel eval-mode BatchNorm2d "scalar" loop of the backbones
 * (src/models/accdoa.py:222-227, 318-321; src/models/einv2.py:106-109, 292-295), in place on
 * x (B, C, T, M) contiguous device fp32.  mean / var / weight / bias are (C, M) device fp32: row c
 * holds running_mean / running_var / weight / bias of scalar[c]; weight and bias may be null
 * (affine=False); mean == var == null means "no scalar" (no-op).  Rounds exactly like torch's CPU
 * kernel: a = weight / sqrt(var + eps) [as w * (1 / sqrt)], b = fma(-mean, a, bias), y = fma(x, a, b).
 * Training-mode batch statistics are not part of this path.  M % 4 == 0, 16-byte aligned pointers.
 *
 * seld_scalar_wav2img_f32: the same map fused with HTSAT_Swin_Transformer.reshape_wav2img
 * (src/models/components/htsat.py:493-511): x (B, C, T, M) -> img (B, C, spec_size, spec_size),
 * img[b][c][k*M + m][t] = y[b][c][k*spec_size + t][m] for k < spec_size / M; frames past T are zero
 * (F.pad after the scalar), frames past (spec_size / M) * spec_size are dropped (negative pad).
 * x is not modified.  spec_size % M == 0 (else SELD_EINVAL), spec_size % 4 == 0. */
int seld_scalar_f32(float* x, int64_t B, int C, int64_t T, int M, const float* mean, const float* var,
                    const float* weight, const float* bias, float eps, void* stream);
int seld_scalar_wav2img_f32(const float* x, int64_t B, int C, int64_t T, int M, int spec_size,
                            const float* mean, const float* var, const float* weight, const float* bias,
                            float eps, float* img, void* stream);

/* ---- Waveform-domain augmentation of the staged batch, the step before the extractors in
 * the reference's training loop (src/models/model_module.py:53-58).  The random draws and the label
 * bookkeeping stay with the caller; these entry points do the waveform arithmetic of a whole batch in
 * one launch, in place, bit-identical to the reference's torch expressions.
 *
 * seld_foa_rotate_f32: Rotation.transform48 / transform16 (src/augment/rotate.py:47-99),
 *   x[b] <- stack(x[b][0], sy * x[b][s_x], sz * x[b][s_y], sx * x[b][s_z])   for every rotated clip b.
 * x (B, C >= 4, L) device fp32 with element strides (stride_b, stride_c, 1); codes (B,) device int32:
 * SELD_ROT_CODE(source channel of output channels 1, 2, 3; their negate flags), or SELD_ROT_IDENTITY for
 * a clip the draw left alone (no memory traffic).  Channel 0 and channels >= 4 are never touched. */
#define SELD_ROT_CODE(s1, s2, s3, n1, n2, n3) \
    ((int32_t)((s1) | ((s2) << 2) | ((s3) << 4) | ((n1) ? 0x100 : 0) | ((n2) ? 0x200 : 0) | ((n3) ? 0x400 : 0)))
#define SELD_ROT_IDENTITY ((int32_t)-1)
int seld_foa_rotate_f32(float* x, int64_t B, int C, int64_t L, int64_t stride_b, int64_t stride_c,
                        const int32_t* codes, void* stream);

/* seld_wavmix_f32: WavMix's waveform line (src/augment/wavmix.py:50),
 *   x[dst_k] <- lam_k * x[dst_k] + (1 - lam_k) * x[src_k],  k < n_ops,
 * every right-hand side taken before any assignment (the reference gathers first).  `ops` is a device
 * array in the order seld_wavmix_order produced: the kernel walks it keeping the previous source in
 * registers, which is what makes the in-place update safe when sources are also destinations.
 *
 * seld_wavmix_order: host-only helper (no CUDA call).  dst / src: n clip indices each in [0, B), no
 * index repeated within dst nor within src (true for the reference: an index list and a permutation);
 * lam: n mixing weights.  Writes n ops to ops_host, ordered along the chains dst_k -> src_k. */
typedef struct seld_mix_op {
    int32_t dst, src;
    float lam;
    int32_t flags;     /* SELD_MIX_* */
} seld_mix_op;
#define SELD_MIX_BEGIN 1      /* first op of a chain: load x[dst] (and remember it as the chain head) */
#define SELD_MIX_USE_HEAD 2   /* the source is the chain head, already overwritten: use the remembered copy */
int seld_wavmix_order(const int64_t* dst, const int64_t* src, const float* lam, int n, int64_t B,
                      seld_mix_op* ops_host);
int seld_wavmix_f32(float* x, int64_t B, int C, int64_t L, int64_t stride_b, int64_t stride_c,
                    const seld_mix_op* ops, int n_ops, void* stream);

/* Kernels enqueued by this library since load (all entry points, all plans). */
uint64_t seld_launch_count(void);
/* cudaError_t of the most recent failing runtime call on this thread (0 if none). */
int seld_last_cuda_error(void);
const char* seld_strerror(int code);
/* "seldfeat <version> sm_100a" */
const char* seld_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SELDFEAT_H_ */
