// Spectral timbre descriptors and the onset envelope (north-star vocabulary; NOT computed by the reference, SURVEY.md
// sections 2.3, 8(f) rank 3: PARITY UNPINNED BY THE REFERENCE; oracle = oracle/descriptors_np.py, whose centroid is pinned
// against torchaudio.functional.spectral_centroid).  Per frame of the MFCC's own STFT grid (n_fft 400, hop 200, periodic
// Hann, centre + reflect padding: the transform inside torchaudio.transforms.MFCC, audio_analyzer.py:207-210):
//
//   centroid [Hz]  sum_k f_k S_k / sum_k S_k        S = magnitude spectrum, f_k = 40 k Hz   (0 / 0 = NaN like torchaudio)
//   rolloff  [Hz]  f_k of the first bin whose cumulative magnitude reaches 85 % of the frame's total
//   flux           || S_t - S_(t-1) ||_2, 0 for the first frame
//   onset          mean over the 128 HTK mel bands (the MFCC's bank) of max(0, L_t - L_(t-1)),
//                  L = 10 log10(max(mel power, 1e-10)), 0 for the first frame
//
// One warp per frame t: frames t - 1 and t are transformed TOGETHER as the real and imaginary part of one complex
// 400-point FFT (the feature kernel's own two-pass register FFT, msa_fft.cuh: 16 x 25), so the two differences need no
// state carried between warps; every frame is transformed twice, which is fine for a descriptor off the hot path.
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "msa_api_internal.h"
#include "msa_fft.cuh"
#include "msa_tables.hpp"

namespace msa {

constexpr int kSpWarps = 8;
constexpr int kSpRow = 25;                               // row stride (complex words) of the pass-A -> pass-B tile
constexpr int kSpBins = 208;                             // 201 bins + zero pad for the mel trips

struct SpectralTables {                                  // the slice of SmemTables this kernel needs
  float tw400[2 * 15 * 25 + 2];
  float win400[kNfftM];
  float mel_w[kMelTrips * 32];
  uint16_t mel_lo[4 * 32];
};

__device__ __forceinline__ float sp_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float sp_load(const int16_t* p) {
  return __fmaf_rn(__int_as_float(0x4B400000 + (int)__ldg(p)), 1.0f / 32768.0f, -384.0f);
}

template <class InT>
__global__ void __launch_bounds__(kSpWarps * 32) spectral_kernel(const InT* __restrict__ wav, int T, int nF,
                                                                 const SpectralTables* __restrict__ gtab, float* __restrict__ out) {
  __shared__ SpectralTables tab;
  __shared__ __align__(16) c32 tiles[kSpWarps][16 * kSpRow];   // 400 complex; reused for three 208-float rows afterwards
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int4* s = reinterpret_cast<const int4*>(gtab);
    int4* d = reinterpret_cast<int4*>(&tab);
    for (int i = threadIdx.x; i < (int)(sizeof(SpectralTables) / 16); i += blockDim.x) d[i] = __ldg(s + i);
  }
  __syncthreads();
  const InT* x = wav + (size_t)blockIdx.y * T;
  const int t = blockIdx.x * kSpWarps + warp;
  if (t >= nF) return;
  auto xr = [&](int i) -> float {                        // reflect-101 padding of torch.stft(center=True)
    if (i < 0) i = -i;
    else if (i >= T) i = 2 * (T - 1) - i;
    return (i >= 0 && i < T) ? sp_load(x + i) : 0.0f;
  };
  c32* tile = tiles[warp];
  const c32* tw = reinterpret_cast<const c32*>(tab.tw400);
  // pass A: lane n2 < 25 takes the 16 stride-25 samples of frame t - 1 (real part) and frame t (imaginary part)
  bool nz_a = false, nz_b = false;                       // does the frame hold any non-zero sample?
  if (lane < 25) {
    const int sb = kHopM * t - kNfftM / 2;               // first sample of frame t
    c32 z[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float w = tab.win400[25 * n1 + lane];
      const float a = (t > 0) ? xr(sb - kHopM + 25 * n1 + lane) : 0.0f;
      const float b = xr(sb + 25 * n1 + lane);
      nz_a = nz_a || (a != 0.0f);
      nz_b = nz_b || (b != 0.0f);
      z[n1] = c32{w * a, w * b};
    }
    dft16<false>(z);
    tile[lane] = z[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) tile[k1 * kSpRow + lane] = cmul(z[k1], tw[(k1 - 1) * 25 + lane]);
  }
  // a frame of digital silence has an exactly zero spectrum (centroid 0 / 0 = NaN like torchaudio); separated from its
  // packed partner it would come out as that partner's rounding residue instead
  nz_a = __any_sync(0xffffffffu, nz_a);
  nz_b = __any_sync(0xffffffffu, nz_b);
  __syncwarp();
  if (lane < 16) {                                       // pass B: row k1 -> Z[k1 + 16 k2]
    c32* row = tile + lane * kSpRow;
    c32 v[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) v[i] = row[i];
    dft25<false>(v);
#pragma unroll
    for (int i = 0; i < 25; ++i) row[i] = v[i];
  }
  __syncwarp();
  // spectra of both frames: A_k = (Z_k + conj Z_(N-k)) / 2, B_k = (Z_k - conj Z_(N-k)) / (2 i)
  float sum_s = 0.0f, sum_fs = 0.0f, flux = 0.0f;
  float pa[7], pb[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int k = lane + 32 * i;
    pa[i] = pb[i] = 0.0f;
    if (k < kBinsM) {
      const int kk = (k == 0) ? 0 : kNfftM - k;
      const c32 zk = tile[(k & 15) * kSpRow + (k >> 4)], zn = tile[(kk & 15) * kSpRow + (kk >> 4)];
      const c32 sm = add_conj(zk, zn), df = sub_conj(zk, zn);
      pa[i] = nz_a ? 0.25f * fmaf(sm.x, sm.x, sm.y * sm.y) : 0.0f;
      pb[i] = nz_b ? 0.25f * fmaf(df.x, df.x, df.y * df.y) : 0.0f;
    }
  }
  __syncwarp();                                          // every lane has read the tile: its memory now holds the rows
  float* S1 = reinterpret_cast<float*>(tile);            // magnitude of frame t, power of frames t - 1 and t
  float* P0 = S1 + kSpBins;
  float* P1 = P0 + kSpBins;
  static_assert(3 * kSpBins * 4 <= 16 * kSpRow * 8, "rows fit the tile");
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int k = lane + 32 * i;
    if (k < kSpBins) {
      const float sa = sqrtf(pa[i]), sb2 = sqrtf(pb[i]);
      S1[k] = sb2; P0[k] = pa[i]; P1[k] = pb[i];
      sum_s += sb2;
      sum_fs = fmaf(40.0f * (float)k, sb2, sum_fs);
      flux = fmaf(sb2 - sa, sb2 - sa, flux);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum_s += __shfl_xor_sync(0xffffffffu, sum_s, o);
    sum_fs += __shfl_xor_sync(0xffffffffu, sum_fs, o);
    flux += __shfl_xor_sync(0xffffffffu, flux, o);
  }
  __syncwarp();
  // roll-off: lane l scans bins 7 l .. 7 l + 6 behind the exclusive prefix of the lanes before it
  float loc[7], run = 0.0f;
#pragma unroll
  for (int j = 0; j < 7; ++j) { const int k = 7 * lane + j; run += (k < kBinsM) ? S1[k] : 0.0f; loc[j] = run; }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  const float before = incl - run, thr = 0.85f * total;
  int first = 0x7fffffff;
#pragma unroll
  for (int j = 6; j >= 0; --j) { const int k = 7 * lane + j; if (k < kBinsM && before + loc[j] >= thr) first = k; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  // onset strength: the lane's four mel filters (32 s + lane) of both frames
  float onset = 0.0f;
  {
    const int trips[4] = {kMelTrip0, kMelTrip1, kMelTrip2, kMelTrip3};
    int toff = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int lo = tab.mel_lo[32 * s + lane];
      float ea = 0.0f, eb = 0.0f;
      for (int p = 0; p < trips[s]; ++p) {
        const float w = tab.mel_w[(toff + p) * 32 + lane];
        ea = fmaf(w, P0[lo + p], ea);
        eb = fmaf(w, P1[lo + p], eb);
      }
      toff += trips[s];
      const float la = 3.0102999566398120f * __log2f(fmaxf(ea, 1e-10f)), lb = 3.0102999566398120f * __log2f(fmaxf(eb, 1e-10f));
      onset += fmaxf(lb - la, 0.0f);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) onset += __shfl_xor_sync(0xffffffffu, onset, o);
  if (lane == 0) {
    float* o4 = out + ((size_t)blockIdx.y * nF + t) * 4;
    o4[0] = sum_fs / sum_s;
    o4[1] = (first == 0x7fffffff) ? 0.0f : 40.0f * (float)first;
    o4[2] = (t > 0) ? sqrtf(flux) : 0.0f;
    o4[3] = (t > 0) ? onset * (1.0f / (float)kMels) : 0.0f;
  }
}

static std::mutex g_sp_mutex;
static SpectralTables* g_sp_dev[64] = {nullptr};

static int get_spectral_tables(const SpectralTables** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return MSA_ERR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(g_sp_mutex);
  if (!g_sp_dev[dev]) {
    static FeatureTables ft;
    static SpectralTables host;
    build_feature_tables(ft);
    std::memcpy(host.tw400, ft.s.tw400, sizeof(ft.s.tw400));
    std::memcpy(host.win400, ft.s.win400, sizeof(host.win400));
    std::memcpy(host.mel_w, ft.s.mel_w, sizeof(host.mel_w));
    std::memcpy(host.mel_lo, ft.s.mel_lo, sizeof(host.mel_lo));
    SpectralTables* d = nullptr;
    e = cudaMalloc(&d, sizeof(SpectralTables));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d, &host, sizeof(SpectralTables), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return (int)e; }
    g_sp_dev[dev] = d;
  }
  *out = g_sp_dev[dev];
  return MSA_OK;
}

template <class InT>
static int launch_spectral(const InT* wav, int B, int T, float* out, cudaStream_t st) {
  static_assert(sizeof(SpectralTables) % 16 == 0, "copied as int4");
  if (!wav || !out || B < 0 || T <= kNfftM / 2) return MSA_ERR_BAD_ARGUMENT;   // reflect padding needs T > n_fft / 2 (torch.stft raises)
  if (B == 0) return MSA_OK;
  const SpectralTables* tab = nullptr;
  const int rc = get_spectral_tables(&tab);
  if (rc != MSA_OK) return rc;
  const int nF = T / kHopM + 1;
  spectral_kernel<InT><<<dim3((nF + kSpWarps - 1) / kSpWarps, B), kSpWarps * 32, 0, st>>>(wav, T, nF, tab, out);
  note_launches(1);
  return (int)cudaGetLastError();
}

}  // namespace msa

extern "C" int msa_spectral_frames(int T) { return T < 1 ? 0 : T / msa::kHopM + 1; }
extern "C" int msa_spectral_f32(const float* wav, int B, int T, float* out4, void* stream) {
  msa::reset_launches();
  return msa::launch_spectral<float>(wav, B, T, out4, (cudaStream_t)stream);
}
extern "C" int msa_spectral_s16(const int16_t* pcm, int B, int T, float* out4, void* stream) {
  msa::reset_launches();
  return msa::launch_spectral<int16_t>(pcm, B, T, out4, (cudaStream_t)stream);
}
