// The rows either side of the hot path.
//
//  1. PCM ingest + resample to 16 kHz: restates torchaudio.transforms.Resample(sr, 16000) as the reference
//     calls it (/root/reference/src/analyzers/audio_analyzer.py:74-77; torchaudio functional.py
//     _get_sinc_resample_kernel / _apply_sinc_resample_kernel, defaults sinc_interp_hann, lowpass_filter_width 6,
//     rolloff 0.99).  After reducing by the gcd the transform is a polyphase FIR: output sample
//     f * nw + j = sum_k kern[j][k] * xpad[f * orig + k], xpad = x shifted by `width` with zeros outside,
//     K = 2 width + orig taps.  One CTA stages the input span of a tile of frames in shared memory; every
//     thread owns one phase j and four frames, so a tap costs one coalesced table load, four shared-memory
//     loads and four FMAs.  The windowed sinc of phase j is non-zero only within lowpass_filter_width zero
//     crossings of its centre (torchaudio clamps t to +-6, where the Hann window vanishes): of the 475 taps
//     of the 441 -> 160 filter (44.1 kHz) at most 35 per phase are non-zero.  The table is therefore stored
//     compactly as [tap - first[j]][j] and a thread runs over its phase's run only; skipped taps are exact
//     zeros, so the sums are bit-identical to the dense filter.  Both cases are bandwidth bound.
//
//  2. Feature-row normalisation: Face/Text/AudioFeatureNormalizer.normalize of src/utils/normalization.py:19-98
//     (zero-pad or truncate to the target width, LayerNorm with eps 1e-5 and biased variance) and the
//     nan_to_num of the row assembly (src/processors/streaming_processor.py:293-300).  One warp per row.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <mutex>
#include <vector>

#include "msa_api_internal.h"

namespace msa {

// ------------------------------------------------------------------------------ resample tables
constexpr int kLowpassWidth = 6;
constexpr double kRolloff = 0.99;
constexpr int kResampleSmemFloats = 11000;     // 44 KB: no opt-in needed
constexpr int kFrameTile = 4;                  // frames per thread

struct ResampleTable {
  int dev, orig, nw, width, K;
  int run;                                     // longest run of non-zero taps of a phase
  float* kt;                                   // [run][nw] on the device: kt[i][j] = kernel[j][first[j] + i]
  int* first;                                  // [nw] first tap of phase j's run (first[j] + run <= K)
};
static std::mutex g_rs_mutex;
static std::vector<ResampleTable> g_rs_tables;

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// torchaudio builds the kernel in float64 (the phase offsets -j / nw go through one float32 division:
// torch.arange(0, -nw, -1) / nw is a float32 tensor) and casts it to float32.
static void build_resample_kernel(int orig, int nw, int* width_out, std::vector<float>& kt) {
  const double PI = 3.14159265358979323846;
  const double base = (double)(orig < nw ? orig : nw) * kRolloff;
  const int width = (int)std::ceil((double)kLowpassWidth * orig / base);
  const int K = 2 * width + orig;
  kt.assign((size_t)K * nw, 0.0f);
  for (int j = 0; j < nw; ++j) {
    const double phase = (double)((float)(-j) / (float)nw);
    for (int k = 0; k < K; ++k) {
      double t = (phase + (double)(k - width) / (double)orig) * base;
      if (t < -(double)kLowpassWidth) t = -(double)kLowpassWidth;
      if (t > (double)kLowpassWidth) t = (double)kLowpassWidth;
      const double c = std::cos(t * PI / kLowpassWidth / 2.0);
      const double window = c * c;
      t *= PI;
      const double sinc = (t == 0.0) ? 1.0 : std::sin(t) / t;
      kt[(size_t)k * nw + j] = (float)(sinc * (window * (base / (double)orig)));
    }
  }
  *width_out = width;
}

static int get_resample_table(int orig, int nw, ResampleTable* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lock(g_rs_mutex);
  for (const ResampleTable& t : g_rs_tables)
    if (t.dev == dev && t.orig == orig && t.nw == nw) { *out = t; return MSA_OK; }
  ResampleTable t{dev, orig, nw, 0, 0, 0, nullptr, nullptr};
  std::vector<float> host;
  build_resample_kernel(orig, nw, &t.width, host);
  t.K = 2 * t.width + orig;
  // per phase: the run [lo, hi] of taps that are not exactly zero in fp32
  std::vector<int> lo(nw, 0), hi(nw, -1);
  for (int j = 0; j < nw; ++j) {
    int a = t.K, b = -1;
    for (int k = 0; k < t.K; ++k)
      if (host[(size_t)k * nw + j] != 0.0f) { if (k < a) a = k; b = k; }
    if (b < 0) { a = 0; b = 0; }
    lo[j] = a; hi[j] = b;
    if (b - a + 1 > t.run) t.run = b - a + 1;
  }
  std::vector<float> compact((size_t)t.run * nw, 0.0f);
  for (int j = 0; j < nw; ++j) {
    if (lo[j] + t.run > t.K) lo[j] = t.K - t.run;              // keep the run inside the staged span (extra taps are zeros)
    for (int i = 0; i < t.run; ++i) compact[(size_t)i * nw + j] = host[(size_t)(lo[j] + i) * nw + j];
  }
  e = cudaMalloc(&t.kt, compact.size() * sizeof(float));
  if (e != cudaSuccess) return (int)e;
  e = cudaMalloc(&t.first, (size_t)nw * sizeof(int));
  if (e != cudaSuccess) { cudaFree(t.kt); return (int)e; }
  e = cudaMemcpy(t.kt, compact.data(), compact.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t.first, lo.data(), (size_t)nw * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t.kt); cudaFree(t.first); return (int)e; }
  g_rs_tables.push_back(t);
  *out = t;
  return MSA_OK;
}

__device__ __forceinline__ float load_sample(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_sample(const int16_t* p) {
  return __fmaf_rn(__int_as_float(0x4B400000 + (int)__ldg(p)), 1.0f / 32768.0f, -384.0f);   // exact v / 32768 without I2F
}

// grid (frame tiles, B); FR frames per CTA (multiple of 4)
template <class InT>
__global__ void __launch_bounds__(256) resample_kernel(const InT* __restrict__ x, int L, int orig, int nw, int width, int K,
                                                       int run, const float* __restrict__ kt, const int* __restrict__ first_tap,
                                                       float* __restrict__ y, int L_out, int FR) {
  extern __shared__ float xs[];
  const int f0 = blockIdx.x * FR;
  const InT* xb = x + (size_t)blockIdx.y * L;
  const int span = FR * orig + K;
  const int first = f0 * orig - width;                       // input index of xs[0]
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int t = first + i;
    xs[i] = (t >= 0 && t < L) ? load_sample(xb + t) : 0.0f;
  }
  __syncthreads();
  const int FQ = FR / kFrameTile;
  float* yb = y + (size_t)blockIdx.y * L_out;
  for (int w = threadIdx.x; w < FQ * nw; w += blockDim.x) {
    const int j = w % nw, fi = w / nw;
    float acc[kFrameTile] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float* xq = xs + fi * orig + __ldg(first_tap + j);
    const float* kj = kt + j;
    const int qs = FQ * orig;                                // frames fi, fi + FQ, fi + 2 FQ, fi + 3 FQ
    for (int i = 0; i < run; ++i) {                          // taps first[j] .. first[j] + run - 1, ascending like the dense sum
      const float wt = __ldg(kj + (size_t)i * nw);
#pragma unroll
      for (int q = 0; q < kFrameTile; ++q) acc[q] = fmaf(wt, xq[q * qs + i], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kFrameTile; ++q) {
      const long long o = (long long)(f0 + fi + q * FQ) * nw + j;
      if (o < L_out) yb[o] = acc[q];
    }
  }
}

template <class InT>
static int launch_resample(const InT* x, int B, int L, int orig_freq, int new_freq, float* y, int L_out, cudaStream_t st) {
  if (!x || !y || B < 0 || L < 1 || orig_freq < 1 || new_freq < 1) return MSA_ERR_BAD_ARGUMENT;
  const int g = gcd_int(orig_freq, new_freq);
  const int orig = orig_freq / g, nw = new_freq / g;
  const long long target = ((long long)nw * L + orig - 1) / orig;          // ceil(new * length / orig)
  if (L_out != (int)target) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  ResampleTable t;
  int rc = get_resample_table(orig, nw, &t);
  if (rc != MSA_OK) return rc;
  if (t.K + kFrameTile * orig > kResampleSmemFloats) return MSA_ERR_UNSUPPORTED_LENGTH;   // rates with a tiny gcd
  int FR = 4096 / nw;
  const int cap = (kResampleSmemFloats - t.K) / orig;
  if (FR > cap) FR = cap;
  FR = (FR / kFrameTile) * kFrameTile;
  if (FR < kFrameTile) FR = kFrameTile;
  const int n_frames = (int)((target + nw - 1) / nw);
  dim3 grid((n_frames + FR - 1) / FR, B);
  const size_t smem = (size_t)(FR * orig + t.K) * sizeof(float);
  resample_kernel<InT><<<grid, 256, smem, st>>>(x, L, orig, nw, t.width, t.K, t.run, t.kt, t.first, y, L_out, FR);
  note_launches(1);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------ row normalisation
// y[r, 0:D] = LayerNorm(pad_or_truncate(x[r, 0:d_in], D)) * gamma + beta  (gamma / beta may be null = 1 / 0),
// optional nan_to_num (NaN -> 0, +-inf -> +-FLT_MAX).  One warp per row.
__global__ void __launch_bounds__(256) rows_layernorm_kernel(const float* __restrict__ x, int B, int d_in, int ld_in, int D,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float eps, float* __restrict__ y, int ld_out, int flags) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (size_t)row * ld_in;
  const int n = d_in < D ? d_in : D;                          // columns that come from the input, the rest are zeros
  float s = 0.0f;
  for (int c = lane; c < n; c += 32) s += __ldg(xr + c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.0f;
  for (int c = lane; c < D; c += 32) {
    const float d = ((c < n) ? __ldg(xr + c) : 0.0f) - mean;
    q = fmaf(d, d, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)D + eps);
  float* yr = y + (size_t)row * ld_out;
  for (int c = lane; c < D; c += 32) {
    float v = (((c < n) ? __ldg(xr + c) : 0.0f) - mean) * rstd;
    if (gamma) v *= __ldg(gamma + c);
    if (beta) v += __ldg(beta + c);
    if (flags & 1) {
      if (v != v) v = 0.0f;
      else if (v > 3.4028234663852886e38f) v = 3.4028234663852886e38f;
      else if (v < -3.4028234663852886e38f) v = -3.4028234663852886e38f;
    }
    yr[c] = v;
  }
}

// torch.nan_to_num(x, nan=0.0) in place: NaN -> 0, +-inf -> +-FLT_MAX (the torch defaults the reference relies on)
__global__ void nan_to_num_kernel(float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (v != v) v = 0.0f;
  else if (v > 3.4028234663852886e38f) v = 3.4028234663852886e38f;
  else if (v < -3.4028234663852886e38f) v = -3.4028234663852886e38f;
  x[i] = v;
}

}  // namespace msa

extern "C" int msa_nan_to_num(float* x, long long n, void* stream) {
  msa::reset_launches();
  if (!x || n < 0) return MSA_ERR_BAD_ARGUMENT;
  if (n == 0) return MSA_OK;
  msa::nan_to_num_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n);
  msa::note_launches(1);
  return (int)cudaGetLastError();
}

extern "C" int msa_resample_out_len(int length, int orig_freq, int new_freq) {
  if (length < 0 || orig_freq < 1 || new_freq < 1) return -1;
  const int g = msa::gcd_int(orig_freq, new_freq);
  const long long orig = orig_freq / g, nw = new_freq / g;
  return (int)((nw * length + orig - 1) / orig);
}

extern "C" int msa_resample_kernel_host(int orig_freq, int new_freq, float* kernel_out, int capacity, int* width, int* taps,
                                        int* phases) {
  if (orig_freq < 1 || new_freq < 1 || !width || !taps || !phases) return MSA_ERR_BAD_ARGUMENT;
  const int g = msa::gcd_int(orig_freq, new_freq);
  const int orig = orig_freq / g, nw = new_freq / g;
  std::vector<float> kt;
  msa::build_resample_kernel(orig, nw, width, kt);
  *taps = 2 * *width + orig;
  *phases = nw;
  if (kernel_out) {
    if (capacity < (int)kt.size()) return MSA_ERR_WORKSPACE;
    for (int j = 0; j < nw; ++j)                                 // returned as torchaudio lays it out: [phase][tap]
      for (int k = 0; k < *taps; ++k) kernel_out[(size_t)j * *taps + k] = kt[(size_t)k * nw + j];
  }
  return MSA_OK;
}

extern "C" int msa_resample_f32(const float* x, int B, int length, int orig_freq, int new_freq, float* y, int out_length,
                                void* stream) {
  msa::reset_launches();
  return msa::launch_resample<float>(x, B, length, orig_freq, new_freq, y, out_length, (cudaStream_t)stream);
}

extern "C" int msa_resample_s16(const int16_t* pcm, int B, int length, int orig_freq, int new_freq, float* y, int out_length,
                                void* stream) {
  msa::reset_launches();
  return msa::launch_resample<int16_t>(pcm, B, length, orig_freq, new_freq, y, out_length, (cudaStream_t)stream);
}

extern "C" int msa_rows_layernorm(const float* x, int B, int d_in, int ld_in, int target_dim, const float* gamma,
                                  const float* beta, float eps, float* y, int ld_out, int flags, void* stream) {
  msa::reset_launches();
  if (!x || !y || B < 0 || d_in < 1 || target_dim < 1 || ld_in < d_in || ld_out < target_dim) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  msa::rows_layernorm_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, B, d_in, ld_in, target_dim, gamma, beta, eps, y,
                                                                           ld_out, flags);
  msa::note_launches(1);
  return (int)cudaGetLastError();
}
