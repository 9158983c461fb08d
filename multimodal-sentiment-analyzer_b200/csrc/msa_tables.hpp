// Constant tables of the audio-feature kernel, built on the host (fp64 -> fp32).
//
// They restate what torchaudio 2.x builds inside the reference's transforms (not vendored in
// /root/reference; pinned 2.5.1 in requirements.txt:350):
//   - periodic Hann windows for n_fft 400 (MFCC, audio_analyzer.py:207-210) and 512
//     (PitchShift, audio_analyzer.py:43-47), and the overlap-added squared window of torch.istft,
//   - the 128-filter HTK mel bank over 201 bins, f in [0, 8000] (MelScale defaults): every
//     filter is a run of consecutive bins (394 non-zeros out of 25,728; filters 0, 3, 6, 13 empty),
//   - the ortho DCT-II 128 -> 13 (create_dct),
//   - the inter-pass twiddles of the 16 x 32 and 16 x 25 FFT factorisations (msa_fft.cuh).
//
// Lane-major layouts: in the kernel lane l of a warp owns mel filters m = 32 s + l (s = 0..3), so
// every per-filter constant is stored [slot][lane] and a warp reads it conflict-free.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace msa {

constexpr int kSampleRate = 16000;
constexpr int kNfftM = 400, kHopM = 200, kBinsM = 201;   // MFCC STFT
constexpr int kMels = 128, kMfcc = 13;
constexpr int kNfftP = 512, kHopP = 128;                 // "pitch" STFT/ISTFT
constexpr int kAtom = 80;                                // common unit of the 400/160 rhythm frames and 1600 blocks
constexpr int kRhythmWin = 400, kRhythmHop = 160, kBlock = 1600;
// trip counts of the mel accumulation per filter slot s (= max non-zeros of filters 32 s .. 32 s + 31)
constexpr int kMelTrip0 = 2, kMelTrip1 = 3, kMelTrip2 = 4, kMelTrip3 = 8;
constexpr int kMelTrips = kMelTrip0 + kMelTrip1 + kMelTrip2 + kMelTrip3;   // 17
constexpr int kDctQuads = 13;                            // 4 slots x 13 coefficients = 52 floats = 13 float4 per lane
constexpr int kPowStride = 208;                          // power spectrum row: 201 bins + zero pad for the mel trips

// The part every CTA stages in shared memory (copied as 16-byte words: keep the size a multiple of 16).
struct alignas(16) SmemTables {
  float tw512[2 * 15 * 32];       // [(k1-1)*32 + n2] = W_512^(n2 k1) as (cos, -sin)
  float tw400[2 * 15 * 32];       // [(k1-1)*32 + n2] = W_400^(n2 k1), n2 < 25
  float win400[kNfftM];
  float win512[kNfftP];
  float ienv[kHopP];              // 1 / (512 sum_j win512[j*128 + o]^2): interior window envelope of torch.istft
                                  // times the 1/512 of the unnormalised inverse FFT (exact: power of two)
  float mel_w[kMelTrips * 32];    // [(trip offset of slot s + p)*32 + lane], zero padded
  float dctq[kDctQuads * 32 * 4]; // float4 [i*32 + lane]: flattened (slot, k) = divmod(4 i + c, 13); 0 for empty filters
  float dct_dead[16];             // sum over the empty filters of dct[m][k]
  uint16_t mel_lo[4 * 32];        // first bin of filter 32 s + lane (0 for the empty filters)
  uint16_t mel_dead[4 * 32];      // 1 where the filter has no non-zero weight
};
static_assert(sizeof(SmemTables) % 16 == 0, "SmemTables is copied as int4");

// Everything the device needs, laid out exactly as it is copied to global memory.
struct FeatureTables {
  SmemTables s;
  float dct[kMels * 16];          // plain [m][k] table (tests / reference restatement)
  int mel_nnz;
  int pad[3];
};

inline int build_feature_tables(FeatureTables& ft) {
  std::memset(&ft, 0, sizeof(ft));
  SmemTables& t = ft.s;
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < kNfftM; ++n) t.win400[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftM));
  for (int n = 0; n < kNfftP; ++n) t.win512[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftP));
  for (int o = 0; o < kHopP; ++o) {
    float e = 0.0f;
    for (int j = 0; j < 4; ++j) e += t.win512[j * kHopP + o] * t.win512[j * kHopP + o];
    t.ienv[o] = (1.0f / e) * (1.0f / (float)kNfftP);
  }
  for (int k1 = 1; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = 2.0 * PI * (double)(n2 * k1) / 512.0;
      t.tw512[2 * ((k1 - 1) * 32 + n2)] = (float)std::cos(a);
      t.tw512[2 * ((k1 - 1) * 32 + n2) + 1] = (float)(-std::sin(a));
      const double b = 2.0 * PI * (double)((n2 < 25 ? n2 : 0) * k1) / 400.0;
      t.tw400[2 * ((k1 - 1) * 32 + n2)] = (float)std::cos(b);
      t.tw400[2 * ((k1 - 1) * 32 + n2) + 1] = (float)(-std::sin(b));
    }
  // DCT-II ortho (torchaudio.functional.create_dct(13, 128, "ortho"))
  for (int m = 0; m < kMels; ++m)
    for (int k = 0; k < kMfcc; ++k) {
      double v = std::cos(PI / kMels * (m + 0.5) * k) * std::sqrt(2.0 / kMels);
      if (k == 0) v *= 1.0 / std::sqrt(2.0);
      ft.dct[m * 16 + k] = (float)v;
    }
  // HTK mel bank (torchaudio.functional.melscale_fbanks, norm=None, mel_scale="htk")
  const int trips[4] = {kMelTrip0, kMelTrip1, kMelTrip2, kMelTrip3};
  int trip_off[4], off = 0;
  for (int s = 0; s < 4; ++s) { trip_off[s] = off; off += trips[s]; }
  const int n_freqs = kBinsM;
  const double f_max = kSampleRate / 2.0;
  const double m_max = 2595.0 * std::log10(1.0 + f_max / 700.0);
  std::vector<double> f_pts(kMels + 2);
  for (int i = 0; i < kMels + 2; ++i) f_pts[i] = 700.0 * (std::pow(10.0, (m_max * i / (kMels + 1)) / 2595.0) - 1.0);
  int nnz = 0, bad = 0;
  for (int m = 0; m < kMels; ++m) {
    const int s = m / 32, lane = m % 32;
    int first = -1, cnt = 0;
    for (int k = 0; k < n_freqs; ++k) {
      const double f = f_max * k / (n_freqs - 1);
      const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      const double w = std::fmin(down, up);
      if (w > 1e-9) {   // fp64 leaves a 1e-14 crumb at the Nyquist corner; fp32 torch has 0 there
        if (first < 0) first = k;
        if (k != first + cnt || cnt >= trips[s]) { ++bad; continue; }   // filters must be runs that fit their slot
        t.mel_w[(trip_off[s] + cnt) * 32 + lane] = (float)w;
        ++cnt;
        ++nnz;
      }
    }
    t.mel_lo[s * 32 + lane] = (uint16_t)(first < 0 ? 0 : first);
    t.mel_dead[s * 32 + lane] = (uint16_t)(cnt == 0);
    for (int k = 0; k < kMfcc; ++k) {
      const int e = s * kMfcc + k;                       // flattened (slot, k)
      if (cnt == 0) t.dct_dead[k] += ft.dct[m * 16 + k];
      else t.dctq[((e / 4) * 32 + lane) * 4 + (e % 4)] = ft.dct[m * 16 + k];
    }
  }
  ft.mel_nnz = nnz;
  return bad;   // 0 when the compile-time trip counts match the filter bank
}

}  // namespace msa
