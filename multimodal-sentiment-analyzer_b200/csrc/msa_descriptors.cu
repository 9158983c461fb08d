// Additive descriptors the north star names and the reference does not compute (SURVEY.md sections 2.3, 8(f) rank 3):
// a real f0 track, frame-level voicing flags and class probabilities.  PARITY IS UNPINNED BY THE REFERENCE; the
// oracle (oracle/descriptors_np.py) restates the same third-party dependency the reference uses:
//
//   f0      torchaudio.functional.detect_pitch_frequency(waveform, 16000) with its defaults (functional.py:
//           _compute_nccf, _find_max_per_frame, _median_smoothing): frames of 160 samples, 189 lags (85 Hz),
//           nccf[f][lag] = <s1, s2> / (1e-9 + |s1|)^2 / (1e-9 + |s2|)^2, best lag >= 6 (3400 Hz) with the lower
//           half of the lag range preferred when within 1 %, lower median over 30 frames, f0 = 16000 / lag.
//   voiced  per 400/160 frame: energy > 0.1 * mean frame energy, the frame-level analogue of the reference's
//           segment-level `energy > 0.1 * energy.mean()` (/root/reference/src/analyzers/audio_analyzer.py:223-228).
//   probs   softmax over the 7 fused logits (fusion_model.py:94 leaves them raw).
//
// One warp per 10 ms frame (CTAs of 64 frames): the frame and its look-ahead sit in a shared-memory row; lane l takes
// lags l + 1, l + 33, ...  Every sum of 160 terms (<s1, s2>, |s1|^2, |s2|^2) is formed in fp32 in EXACTLY the order of
// the oracle's (and numpy's) pairwise summation - products rounded before they are added, two blocks of 80, eight
// running sums per block combined as ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)) - so that the NCCF values, and
// with them every arg-max decision including the near-ties, are bit-identical to oracle/descriptors_np.py.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "msa_api_internal.h"

namespace msa {

constexpr int kPFrame = 160, kPLags = 189, kPLagMin = 5, kPMedWin = 30, kPWarps = 8;
constexpr int kPChunk = 64;                    // frames per CTA of the lag kernel (8 per warp)
constexpr int kPRow = 352;                     // frame + 189 samples of look-ahead, rounded to 32
constexpr int kVWin = 400, kVHop = 160;

__device__ __forceinline__ float pt_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float pt_load(const int16_t* p) {
  return __fmaf_rn(__int_as_float(0x4B400000 + (int)__ldg(p)), 1.0f / 32768.0f, -384.0f);
}

// sum_i a[i] * b[i] over 160 terms in numpy's pairwise order (see the header): no fused multiply-adds
__device__ __forceinline__ float np_dot160(const float* __restrict__ a, const float* __restrict__ b) {
  float total = 0.0f;
#pragma unroll 1
  for (int blk = 0; blk < 2; ++blk) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fmul_rn(a[80 * blk + j], b[80 * blk + j]);
#pragma unroll 3
    for (int i = 8; i < 80; i += 8)
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __fmul_rn(a[80 * blk + i + j], b[80 * blk + i + j]));
    const float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    total = (blk == 0) ? res : __fadd_rn(total, res);
  }
  return total;
}

// grid (frame chunks, B): a CTA takes kPChunk consecutive frames of one segment (fine-grained CTAs keep the last
// wave of a batch short: a whole 5 s segment per CTA left a third of the GPU idle at 1024 segments)
template <class InT>
__global__ void __launch_bounds__(kPWarps * 32) pitch_lags_kernel(const InT* __restrict__ wav, int T, int nf,
                                                                  int32_t* __restrict__ lags_out) {
  __shared__ float rows[kPWarps][kPRow];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const InT* x = wav + (size_t)blockIdx.y * T;
  int32_t* lags = lags_out + (size_t)blockIdx.y * nf;
  float* s = rows[warp];
  const int f_end = min(nf, (int)(blockIdx.x + 1) * kPChunk);

  for (int f = blockIdx.x * kPChunk + warp; f < f_end; f += kPWarps) {
    const int base = f * kPFrame;
#pragma unroll
    for (int k = 0; k < kPRow / 32; ++k) {
      const int t = base + lane + 32 * k;
      s[lane + 32 * k] = (t < T) ? pt_load(x + t) : 0.0f;          // torch pads the waveform with zeros
    }
    __syncwarp();
    const float r1 = __fadd_rn(1e-9f, __fsqrt_rn(np_dot160(s, s)));
    const float n1 = __fmul_rn(r1, r1);                            // (EPS + |s1|) ** 2
    float bv = -3.4e38f, hv = -3.4e38f;
    int bl = 0x7fffffff, hl = 0x7fffffff;
#pragma unroll 1
    for (int lag = 1 + lane; lag <= kPLags; lag += 32) {           // ascending lags per lane: the first maximum wins
      const float r2 = __fadd_rn(1e-9f, __fsqrt_rn(np_dot160(s + lag, s + lag)));
      const float v = __fdiv_rn(__fdiv_rn(np_dot160(s, s + lag), n1), __fmul_rn(r2, r2));
      if (lag > kPLagMin) {
        if (v > bv) { bv = v; bl = lag; }
        if (lag <= kPLags / 2 && v > hv) { hv = v; hl = lag; }     // nccf[..., lag_min : lags // 2]
      }
    }
    __syncwarp();                                                  // the row is free for the next frame
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
      const float pv = __shfl_xor_sync(0xffffffffu, hv, o);
      const int pl = __shfl_xor_sync(0xffffffffu, hl, o);
      if (pv > hv || (pv == hv && pl < hl)) { hv = pv; hl = pl; }
    }
    if (lane == 0) lags[f] = (hv > __fmul_rn(0.99f, bv)) ? hl : bl;   // _combine_max(half, best, thresh = 0.99)
  }
}

// grid B: smoothing of the lag track and the voicing flags of one segment
template <class InT>
__global__ void __launch_bounds__(kPWarps * 32) pitch_finish_kernel(const InT* __restrict__ wav, int T, int nf, int n_out,
                                                                    const int32_t* __restrict__ lags_in, float* __restrict__ f0_out,
                                                                    int32_t* __restrict__ voiced_out, int nv) {
  __shared__ double red[kPWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const InT* x = wav + (size_t)blockIdx.x * T;
  const int32_t* lags = lags_in + (size_t)blockIdx.x * nf;

  // lower median over windows of 30 frames, 14 copies of the first value in front (_median_smoothing)
  float* f0 = f0_out + (size_t)blockIdx.x * n_out;
  for (int t = threadIdx.x; t < n_out; t += blockDim.x) {
    int w[kPMedWin];
#pragma unroll
    for (int j = 0; j < kPMedWin; ++j) {
      const int p = t + j - (kPMedWin - 1) / 2;
      w[j] = lags[p < 0 ? 0 : p];
    }
    int med = w[0];
#pragma unroll 1
    for (int j = 0; j < kPMedWin; ++j) {
      int less = 0, le = 0;
#pragma unroll
      for (int k = 0; k < kPMedWin; ++k) { less += (w[k] < w[j]); le += (w[k] <= w[j]); }
      if (less <= (kPMedWin - 1) / 2 && (kPMedWin - 1) / 2 < le) med = w[j];
    }
    f0[t] = (1.0f / (1e-9f + (float)med)) * 16000.0f;              // torch: reciprocal(tensor) * sample_rate
  }

  // frame-level voicing: energy of the 400/160 frames against a tenth of their mean
  if (voiced_out != nullptr && nv > 0) {
    float* eg = reinterpret_cast<float*>(voiced_out + (size_t)blockIdx.x * nv);      // energies first, flags in place
    for (int g = warp; g < nv; g += kPWarps) {
      float e = 0.0f;
      for (int i = lane; i < kVWin; i += 32) { const float v = pt_load(x + g * kVHop + i); e = fmaf(v, v, e); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if (lane == 0) eg[g] = e;
    }
    __syncthreads();
    double acc = 0.0;
    for (int g = threadIdx.x; g < nv; g += blockDim.x) acc += (double)eg[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int w2 = 0; w2 < kPWarps; ++w2) tot += red[w2];
    const double thr = 0.1 * (tot / (double)nv);
    __syncthreads();
    for (int g = threadIdx.x; g < nv; g += blockDim.x) {
      const int flag = ((double)eg[g] > thr) ? 1 : 0;
      voiced_out[(size_t)blockIdx.x * nv + g] = flag;
    }
  }
}

__global__ void softmax7_kernel(const float* __restrict__ logits, int B, float* __restrict__ probs) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  float v[7], m = -3.4e38f;
#pragma unroll
  for (int j = 0; j < 7; ++j) { v[j] = logits[(size_t)r * 7 + j]; m = fmaxf(m, v[j]); }
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 7; ++j) { v[j] = expf(v[j] - m); sum += v[j]; }
#pragma unroll
  for (int j = 0; j < 7; ++j) probs[(size_t)r * 7 + j] = v[j] / sum;
}

template <class InT>
static int launch_pitch(const InT* wav, int B, int T, int32_t* lags, float* f0, int32_t* voiced, cudaStream_t st) {
  if (!wav || !lags || B < 0 || T < 1) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  const int nf = (T + kPFrame - 1) / kPFrame;
  const int n_out = nf - (kPMedWin - 1 - (kPMedWin - 1) / 2) > 0 ? nf - (kPMedWin - 1 - (kPMedWin - 1) / 2) : 0;
  if (n_out > 0 && !f0) return MSA_ERR_BAD_ARGUMENT;           // fewer than 16 frames: no smoothed output (torchaudio raises)
  const int nv = (T >= kVWin) ? (T - kVWin) / kVHop + 1 : 0;
  pitch_lags_kernel<InT><<<dim3((nf + kPChunk - 1) / kPChunk, B), kPWarps * 32, 0, st>>>(wav, T, nf, lags);
  pitch_finish_kernel<InT><<<B, kPWarps * 32, 0, st>>>(wav, T, nf, n_out, lags, f0, voiced, nv);
  note_launches(2);
  return (int)cudaGetLastError();
}

}  // namespace msa

extern "C" int msa_pitch_frames(int T) { return T < 1 ? 0 : (T + msa::kPFrame - 1) / msa::kPFrame; }
extern "C" int msa_pitch_outputs(int T) {
  const int nf = msa_pitch_frames(T);
  return nf - 15 > 0 ? nf - 15 : 0;                                // nf + 14 front copies - 30 + 1 windows
}
extern "C" int msa_voiced_frames(int T) { return T >= msa::kVWin ? (T - msa::kVWin) / msa::kVHop + 1 : 0; }

extern "C" int msa_pitch_track_f32(const float* wav, int B, int T, int32_t* lags, float* f0, int32_t* voiced, void* stream) {
  msa::reset_launches();
  return msa::launch_pitch<float>(wav, B, T, lags, f0, voiced, (cudaStream_t)stream);
}
extern "C" int msa_pitch_track_s16(const int16_t* pcm, int B, int T, int32_t* lags, float* f0, int32_t* voiced, void* stream) {
  msa::reset_launches();
  return msa::launch_pitch<int16_t>(pcm, B, T, lags, f0, voiced, (cudaStream_t)stream);
}

extern "C" int msa_softmax7(const float* logits, int B, float* probs, void* stream) {
  msa::reset_launches();
  if (!logits || !probs || B < 0) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  msa::softmax7_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits, B, probs);
  msa::note_launches(1);
  return (int)cudaGetLastError();
}
