// Constant tables of the audio feature path (host side, double precision, rounded once to fp32).
//
// These restate the constants the reference obtains from torchaudio (not vendored in
// /root/reference; pinned 2.5.1 in requirements.txt:350):
//   - periodic Hann windows for n_fft 400 (MFCC, audio_analyzer.py:207-210) and 512
//     (PitchShift, audio_analyzer.py:43-47),
//   - the 128-filter HTK mel bank over 201 bins, f in [0, 8000] (MelScale defaults),
//     stored sparse (CSR by filter: 394 non-zeros out of 25,728),
//   - the ortho DCT-II 128 -> 13 (create_dct),
//   - twiddles and digit-reversal maps of the in-place mixed-radix FFTs.
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
constexpr int kAtom = 80;                                // gcd-ish unit of the 400/160 frames and 1600 blocks
constexpr int kRhythmWin = 400, kRhythmHop = 160, kBlock = 1600;
constexpr int kMelNnzMax = 448;                          // >= 394 actual non-zeros
constexpr int kDctStride = 16;                           // 13 coefficients padded to 16 floats

// Everything the device needs, laid out exactly as it is copied to global memory.
struct FeatureTables {
  float win400[kNfftM];
  float win512[kNfftP];
  // per-stage twiddles W_NS^(j*k) stored [k-1][j] (consecutive lanes -> consecutive words), (cos, -sin)
  float tw512_s1[2 * 7 * 64];     // NS = 512, radix 8, m = 64
  float tw512_s2[2 * 7 * 8];      // NS = 64,  radix 8, m = 8
  float tw400_s1[2 * 15 * 25];    // NS = 400, radix 16, m = 25
  float tw400_s2[2 * 4 * 5];      // NS = 25,  radix 5, m = 5
  uint16_t perm400[kNfftM];       // position in the DIF output that holds bin k (radices 16,5,5)
  uint16_t mel_ptr[kMels + 1];    // CSR row pointers per mel filter
  uint16_t mel_bin[kMelNnzMax];
  uint16_t mel_pos[2 * kMelNnzMax];   // (position of bin k, position of bin N-k) in the DIF output, per non-zero
  float mel_w[kMelNnzMax];
  float dct[kMels * kDctStride];  // dct[m*16 + k], k < 13
};

inline void build_feature_tables(FeatureTables& t) {
  std::memset(&t, 0, sizeof(t));
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < kNfftM; ++n) t.win400[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftM));
  for (int n = 0; n < kNfftP; ++n) t.win512[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftP));
  auto fill = [&](float* dst, int NS, int R) {
    const int m = NS / R;
    for (int k = 1; k < R; ++k)
      for (int j = 0; j < m; ++j) {
        const double a = 2.0 * PI * (double)(j * k) / NS;
        dst[2 * ((k - 1) * m + j)] = (float)std::cos(a);
        dst[2 * ((k - 1) * m + j) + 1] = (float)(-std::sin(a));
      }
  };
  fill(t.tw512_s1, 512, 8);
  fill(t.tw512_s2, 64, 8);
  fill(t.tw400_s1, 400, 16);
  fill(t.tw400_s2, 25, 5);
  // digit reversal of the in-place DIF with radices (16, 5, 5): position p = d1*25 + d2*5 + d3
  // holds bin k = d1 + 16*d2 + 80*d3.
  for (int p = 0; p < kNfftM; ++p) {
    int d1 = p / 25, d2 = (p % 25) / 5, d3 = p % 5;
    t.perm400[d1 + 16 * d2 + 80 * d3] = (uint16_t)p;
  }
  // HTK mel bank (torchaudio.functional.melscale_fbanks, norm=None, mel_scale="htk")
  {
    const int n_freqs = kBinsM;
    const double f_max = kSampleRate / 2.0;
    const double m_min = 2595.0 * std::log10(1.0 + 0.0 / 700.0);
    const double m_max = 2595.0 * std::log10(1.0 + f_max / 700.0);
    std::vector<double> f_pts(kMels + 2);
    for (int i = 0; i < kMels + 2; ++i) {
      double m = m_min + (m_max - m_min) * i / (kMels + 1);
      f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    }
    int nnz = 0;
    for (int m = 0; m < kMels; ++m) {
      t.mel_ptr[m] = (uint16_t)nnz;
      for (int k = 0; k < n_freqs; ++k) {
        double f = f_max * k / (n_freqs - 1);
        double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
        double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
        double w = std::fmin(down, up);
        if (w > 1e-9 && nnz < kMelNnzMax) {   // fp64 leaves a 1e-14 crumb at the Nyquist corner; fp32 torch has 0 there
          t.mel_bin[nnz] = (uint16_t)k;
          t.mel_w[nnz] = (float)w;
          ++nnz;
        }
      }
    }
    t.mel_ptr[kMels] = (uint16_t)nnz;
    for (int p = 0; p < nnz; ++p) {
      const int k = t.mel_bin[p];
      t.mel_pos[2 * p] = t.perm400[k];
      t.mel_pos[2 * p + 1] = t.perm400[(kNfftM - k) % kNfftM];
    }
  }
  // DCT-II ortho (torchaudio.functional.create_dct(13, 128, "ortho")), stored [mel][16]
  for (int m = 0; m < kMels; ++m)
    for (int k = 0; k < kMfcc; ++k) {
      double v = std::cos(PI / kMels * (m + 0.5) * k) * std::sqrt(2.0 / kMels);
      if (k == 0) v *= 1.0 / std::sqrt(2.0);
      t.dct[m * kDctStride + k] = (float)v;
    }
}

}  // namespace msa
