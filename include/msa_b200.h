/* msa_b200 — C ABI of the B200 (sm_100a) implementation of the audio-features -> fusion hot path.
 *
 * The reference (Joaonic/multimodal-sentiment-analyzer) has no FFI of its own: its boundary for
 * this path is two Python classes.  Every entry point below names the reference interface it
 * replaces (paths relative to the reference root); INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says host; buffers are caller-owned;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 *     on that stream and never synchronise the device;
 *   - return value: MSA_OK (0) or a negative MSA_ERR_* / positive cudaError_t; nothing throws
 *     (the reference's convention is "never raise, return a default": the Python shim maps a
 *     non-zero code to the reference's documented default AND logs it);
 *   - one caller thread per stream; no internal threads; constant tables are built once per
 *     device on first use (thread-safe).
 */
#ifndef MSA_B200_H
#define MSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSA_OK 0
#define MSA_ERR_BAD_ARGUMENT (-1)
#define MSA_ERR_UNSUPPORTED_LENGTH (-2) /* segment too long for one cluster's shared memory */
#define MSA_ERR_NOT_PACKED (-3)
#define MSA_ERR_WORKSPACE (-4)

/* flags of msa_features_* */
#define MSA_FEAT_STRICT_NAN 1 /* mono intensity = NaN exactly like audio_analyzer.py:194-196 (default) */
                              /* bits 2 and 4 were tuning / experiment switches of round 1 (lockstep barrier, folded wave statistics):
                               * removed with the fp32 STFT-512 path they belonged to; set bits are ignored */
/* parts mask: which feature groups to compute (the rest take the reference's exception defaults) */
#define MSA_PART_WAVE 1  /* rhythm, speech_rate, snr, consistency */
#define MSA_PART_MFCC 2  /* timbre, clarity */
#define MSA_PART_PITCH 4 /* STFT-512 -> ISTFT residual (fp16 round trip on the tensor cores: the feature is the mean of the
                          * z-scored residual, |v| <= 1e-6 whatever the residual's precision; detail[65:68] = its mean, std, max) */
#define MSA_PART_ALL 7

#define MSA_DETAIL_STRIDE 96

int msa_version(void);
const char* msa_strerror(int code);

/* Number of CTAs per cluster used for a segment of T samples: the smallest that lets two CTAs share an
 * SM, else the smallest that fits at all (0 if T is unsupported; a 5 s segment needs 1; the limit is
 * about 3 minutes), and the dynamic shared memory per CTA.  With cluster_size = 0 the launch
 * uses this value for large batches and up to 8 CTAs per segment when B is small (streaming). */
int msa_features_cluster_size(int T);
int msa_features_smem_bytes(int T, int cluster_size);

/* Batched AudioAnalyzer feature body: B independent mono segments of T samples each.
 * Replaces, per segment, audio_analyzer.py:89-131 (_analyze_pitch :175-188, _analyze_intensity
 * :190-201, _analyze_timbre :203-217, _analyze_speech_rate :219-233, _analyze_rhythm :235-263,
 * _calculate_audio_quality/_signal_noise_ratio/_clarity/_consistency :265-329), the
 * AudioFeatureNormalizer (src/utils/normalization.py:26-44) and the audio-row assembly +
 * nan_to_num of streaming_processor.py:250-268, 295-298.
 *
 *   wav     [B, T] fp32 in [-1, 1] (torchaudio.load layout, one channel)        (device)
 *   emo8    [B, 8] output of _analyze_emotion (out of scope: wav2vec2), or NULL = uniform 1/8
 *   feat31  [B, 31] out: LayerNorm31(raw27)[:27] ++ [audio_quality, snr, clarity, consistency],
 *           NaN -> 0 (the row AdvancedFusionModel.forward receives as audio_probs)
 *   detail  [B, 96] out or NULL: [0:27] raw features before LayerNorm in analyze()'s concat order
 *           (emotion8, pitch, intensity, timbre13, speech_rate, rhythm3), [27:31] the four quality
 *           floats, [32:63] the full LayerNorm(31) row (NaN where the reference is NaN),
 *           [64:80] diagnostics (top_db max, residual mean/std/max, energies, counts, clamped-pass flag, min dB,
 *           candidate level, clamp flag)
 *   dbg_mfcc [B, T/200+1, 13] out or NULL: the MFCC matrix (frames x coefficients)
 *   cluster_size 0 = auto, else 1/2/4/8/16 (16 is a non-portable cluster size: MSA_ERR_BAD_ARGUMENT where the device does
 *   not co-schedule 16 CTAs of the kernel; auto picks it for a handful of segments only)
 */
int msa_features_f32(const float* wav, int B, int T, const float* emo8, float* feat31, float* detail,
                     float* dbg_mfcc, int flags, int parts, int cluster_size, void* stream);

/* Optional scratch table for msa_features_ws_*: when a segment holds long stretches far below its loudest mel
 * value (pauses, digital silence), the top_db = 80 clamp touches more values than the kernel's on-chip candidate
 * lists hold; with a workspace of msa_features_workspace_bytes(B, T) bytes (device memory, contents irrelevant)
 * those quads save their mel dB values there and the clamp is applied without a second FFT.  Without it (or with
 * the plain msa_features_* calls) the affected quads are recomputed: the results are bit-identical either way. */
size_t msa_features_workspace_bytes(int B, int T);
int msa_features_ws_f32(const float* wav, int B, int T, const float* emo8, float* feat31, float* detail, float* dbg_mfcc,
                        int flags, int parts, int cluster_size, void* workspace, size_t workspace_bytes, void* stream);
int msa_features_ws_s16(const int16_t* pcm, int B, int T, const float* emo8, float* feat31, float* detail, float* dbg_mfcc,
                        int flags, int parts, int cluster_size, void* workspace, size_t workspace_bytes, void* stream);

/* Same, from int16 PCM (pcm_s16le as written by offline_processor.py:87-91 and
 * streaming_processor.py:185-196); samples are scaled by 1/32768 like torchaudio.load. */
int msa_features_s16(const int16_t* pcm, int B, int T, const float* emo8, float* feat31, float* detail,
                     float* dbg_mfcc, int flags, int parts, int cluster_size, void* stream);

/* ---- fusion model (src/models/fusion_model.py) ------------------------------------------- */

/* Size in bytes of the packed-weight blob and of the activation workspace for batch B. */
size_t msa_fusion_packed_bytes(void);
size_t msa_fusion_workspace_bytes(int B);

/* Repack an AdvancedFusionModel state_dict (fusion_model.py:44-103) into the
 * device layout the kernels read.  `tensors` is a HOST array of 42 HOST fp32 pointers in the
 * order msa_fusion_tensor_name(i) gives; `packed` is a DEVICE buffer of msa_fusion_packed_bytes().
 * Synchronous (done once per checkpoint load, fusion_model.py:259-294). */
int msa_fusion_num_tensors(void);
const char* msa_fusion_tensor_name(int i);
size_t msa_fusion_tensor_numel(int i);
int msa_fusion_pack(const float* const* tensors_host, void* packed_dev, void* stream);

/* AdvancedFusionModel.forward in eval mode (fusion_model.py:131-190):
 *   face [B,27], audio [B,31], text [B,783] or NULL.  Supported: all three (_fuse_all, :386-408)
 *   and face+audio (_fuse_face_audio, :296-321).  Every other combination never produces a
 *   "fused" tensor in the reference (pass-through / always-raising pairs); the Python shim
 *   reproduces those dict results without calling the device.
 *   logits7 [B,7] out (the reference's "fused": raw logits, no softmax), argmax [B] int32 out or NULL.
 */
int msa_fusion_forward(const float* face, const float* audio, const float* text, int B, const void* packed,
                       void* workspace, size_t workspace_bytes, float* logits7, int32_t* argmax, void* stream);

/* Batch-size dispatch of msa_fusion_forward: 0 = default (tcgen05 tensor-core kernels; batches of up to 8 rows, i.e. the
 * streaming path's one row per chunk, run as fp32 matrix-vector kernels instead), 2 = tcgen05 kernels for every batch size
 * (also MSA_FUSION_IMPL=tc in the environment).  Same ABI, same results to fp32 rounding.  (Value 1, round 1's fp32
 * CUDA-core cross-check, is no longer part of the library: it lives in tests/xcheck as test infrastructure.) */
int msa_fusion_set_impl(int impl);

/* ---- speaker / timeline aggregation (src/processors/offline_processor.py:259-298) --------- */

/* label [S] int32 = argmax of the fused logits per segment (in segment order), speaker [S] int32 in
 * [0, n_speakers).  hist [n_speakers,7] out: label counts; dominant [n_speakers] out: mode of the
 * labels (:287-290, ties -> smallest label, -1 if the speaker has no segment); run3 [S] out: 1 where
 * a segment starts three equal consecutive labels within its speaker's own sequence (:293-298). */
int msa_aggregate_speakers(const int32_t* label, const int32_t* speaker, int S, int n_speakers, int32_t* hist,
                           int32_t* dominant, int32_t* run3, void* stream);

/* ---- PCM ingest: resample to the analyzer's rate (src/analyzers/audio_analyzer.py:74-77) ------- */

/* torchaudio.transforms.Resample(orig_freq, new_freq) with its defaults (sinc_interp_hann, lowpass_filter_width 6,
 * rolloff 0.99): x [B, length] fp32 (or int16 PCM, scaled by 1/32768 like torchaudio.load) -> y [B, out_length]
 * fp32 with out_length = msa_resample_out_len(length, orig_freq, new_freq) = ceil(new * length / orig).
 * Rate pairs whose reduced filter (2 * width + orig/gcd taps) does not fit one CTA's staging buffer return
 * MSA_ERR_UNSUPPORTED_LENGTH (every common audio rate to 16 kHz fits). */
int msa_resample_out_len(int length, int orig_freq, int new_freq);
int msa_resample_f32(const float* x, int B, int length, int orig_freq, int new_freq, float* y, int out_length, void* stream);
int msa_resample_s16(const int16_t* pcm, int B, int length, int orig_freq, int new_freq, float* y, int out_length, void* stream);
/* The polyphase filter bank the two calls above use, built on the HOST ([phases][taps] like torchaudio's
 * `Resample.kernel`); kernel_out may be NULL to query the sizes.  No device needed (tests, capacity planning). */
int msa_resample_kernel_host(int orig_freq, int new_freq, float* kernel_out, int capacity, int* width, int* taps, int* phases);

/* ---- feature-row normalisation (src/utils/normalization.py:19-98) ------------------------------ */

/* {Face,Text,Audio}FeatureNormalizer.normalize for B rows: x [B, d_in] (row stride ld_in) is zero-padded or
 * truncated to target_dim (27 / 783 / 31) and LayerNorm'ed (biased variance, eps 1e-5 in the reference; gamma /
 * beta [target_dim] or NULL = the untrained 1 / 0) into y [B, target_dim] (row stride ld_out).
 * flags & 1: torch.nan_to_num(nan=0) on the result (streaming_processor.py:293-300). */
#define MSA_ROWS_NAN_TO_NUM 1
int msa_rows_layernorm(const float* x, int B, int d_in, int ld_in, int target_dim, const float* gamma, const float* beta,
                       float eps, float* y, int ld_out, int flags, void* stream);

/* ---- additive descriptors (north-star vocabulary; NOT computed by the reference) ---------- */

/* Real f0 track and frame-level voicing of B mono segments of T samples (16 kHz):
 *   lags   [B, msa_pitch_frames(T)]  int32 out: best NCCF lag per 10 ms frame (torchaudio _find_max_per_frame)
 *   f0     [B, msa_pitch_outputs(T)] fp32 out: 16000 / lower-median-of-30(lags) = torchaudio.functional.
 *          detect_pitch_frequency(waveform, 16000) with its defaults
 *   voiced [B, msa_voiced_frames(T)] int32 out or NULL: 400/160 frame energy > 0.1 * mean frame energy, the
 *          frame-level analogue of audio_analyzer.py:223-228
 * The oracle for these is the restated torchaudio algorithm (oracle/descriptors_np.py), not the reference. */
int msa_pitch_frames(int T);
int msa_pitch_outputs(int T);
int msa_voiced_frames(int T);
int msa_pitch_track_f32(const float* wav, int B, int T, int32_t* lags, float* f0, int32_t* voiced, void* stream);
int msa_pitch_track_s16(const int16_t* pcm, int B, int T, int32_t* lags, float* f0, int32_t* voiced, void* stream);
/* Spectral timbre descriptors and the onset envelope per frame of the MFCC's own STFT grid (n_fft 400, hop 200, periodic
 * Hann, centre + reflect: the transform inside torchaudio.transforms.MFCC, audio_analyzer.py:207-210):
 *   out4 [B, msa_spectral_frames(T), 4] fp32 = spectral centroid [Hz] (= torchaudio.functional.spectral_centroid; NaN for
 *   a silent frame like torchaudio), roll-off [Hz] (first bin whose cumulative magnitude reaches 85 %), spectral flux
 *   (L2 norm of the magnitude difference to the previous frame) and onset strength (mean over the 128 HTK mel bands of
 *   the positive log-power difference).  T must exceed 200 (torch.stft's reflect padding).  Oracle: oracle/descriptors_np.py. */
int msa_spectral_frames(int T);
int msa_spectral_f32(const float* wav, int B, int T, float* out4, void* stream);
int msa_spectral_s16(const int16_t* pcm, int B, int T, float* out4, void* stream);
/* probs [B,7] = softmax(logits [B,7]) (fusion_model.py:94 returns raw logits; consumers argmax them). */
int msa_softmax7(const float* logits, int B, float* probs, void* stream);

/* torch.nan_to_num(x, nan=0.0) in place over n contiguous floats (row assembly, streaming_processor.py:293-300). */
int msa_nan_to_num(float* x, long long n, void* stream);

/* Per-segment result table (the only data that crosses GPUs): rows40 [n, 40] 32-bit words =
 * audio row 31 | logits 7 | argmax (int32 bits) | segment id first_id + r (int32 bits). */
int msa_pack_rows(const float* audio31, const float* logits7, const int32_t* argmax, int first_id, int n, float* rows40,
                  void* stream);

/* Number of kernel launches the last call of each kind issued on this thread (bench accounting). */
int msa_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MSA_B200_H */
