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

// fp32 -> fp16 bits, round to nearest even (host-side table construction; the values are all in [-1, 1])
inline uint16_t f32_to_f16_bits(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const int32_t e = (int32_t)((x >> 23) & 0xffu) - 127 + 15;
  uint32_t m = x & 0x7fffffu;
  if (e >= 31) return (uint16_t)(sign | 0x7bffu);                       // saturate (never reached by the tables)
  if (e <= 0) {                                                          // subnormal half or zero
    if (e < -10) return (uint16_t)sign;
    m |= 0x800000u;
    const int shift = 14 - e;                                            // 14 .. 24
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;               // may carry into the exponent: still correct
  return (uint16_t)(sign | h);
}
inline uint32_t f16x2_bits(double lo, double hi) {
  return (uint32_t)f32_to_f16_bits((float)lo) | ((uint32_t)f32_to_f16_bits((float)hi) << 16);
}

// Constant FRAGMENTS of the tensor-core STFT-512 round trip (msa_pitch_tc.cuh): every entry is the fp16 pair one lane
// (g = lane / 4, t = lane % 4) of a warp holds in one register of an mma.sync m16n8k16 operand, so the kernel reads
// them with one conflict-free vector load per lane.  512 = 16 x 32, n = 32 n1 + n2, k = k1 + 16 k2.
struct alignas(16) PitchSmemTables {      // read inside the frame loop
  uint32_t cs[16][32][2];                 // [8 sin + 4 ks + m][lane] B fragment (rows 16 ks + .., columns 8 m + g) of cos|sin(2 pi r c / 32) / 16
  uint32_t ws[4][32][4];                  // [i][lane][j] synthesis window pair at n1 = (g + 4 i) % 16, n2 = 8 j + 2 t
  uint32_t a4[4][32][4];                  // [Vre^0, -Vim^0, Vre^1, Vim^1][lane] A fragments of exp(+2 pi i n1 k1 / 16) / 2, rows rho = (n1 + 4 a) % 16
};
struct alignas(16) PitchRegTables {       // read once per work item into registers
  uint32_t a1[3][32][4];                  // [cos, -sin, +sin][lane] A fragments of W_16^(k1 n1)
  uint32_t tw[8][32][2];                  // [2 j + h][lane] {cos, sin}(2 pi k1 n2 / 512) pairs, k1 = g + 8 h, n2 = 8 j + 2 t
  uint32_t wa[4][32][2];                  // [j][lane] analysis window pairs in B-fragment order (rows n1 = 2 t.., column n2 = 8 j + g)
};

// The part every CTA stages in shared memory (copied as 16-byte words: keep the size a multiple of 16).
struct alignas(16) SmemTables {
  float tw400[2 * 15 * 25];       // [(k1-1)*25 + n2] = W_400^(n2 k1) as (cos, -sin), n2 < 25 (8-byte entries)
  float tw400_pad[2];             // keeps the next member 16-byte aligned
  float win400[kNfftM];
  float mel_w[kMelTrips * 32];    // [(trip offset of slot s + p)*32 + lane], zero padded
  // the 13 x 128 DCT as mma.sync m16n8k16 A fragments (rows = coefficient k, padded to 16; K = mel filter), split into
  // fp16 hi + lo parts: [hi | lo][k-step kappa][lane][4 registers]; rows of EMPTY mel filters are zero (they enter through
  // dct_dead).  lane (g, t): {D[g][16 kappa + 2 t], D[g][.. + 1]}, {D[g + 8][..]}, {D[g][16 kappa + 2 t + 8], ..}, {D[g + 8][..]}
  uint32_t dct_frag[2][8][32][4];
  float dct_dead[16];             // sum over the empty filters of dct[m][k]
  uint16_t mel_lo[4 * 32];        // first bin of filter 32 s + lane (0 for the empty filters)
  uint16_t mel_dead[4 * 32];      // 1 where the filter has no non-zero weight
  PitchSmemTables pt;
};
static_assert(sizeof(SmemTables) % 16 == 0, "SmemTables is copied as int4");

// Everything the device needs, laid out exactly as it is copied to global memory.
struct FeatureTables {
  SmemTables s;
  float dct[kMels * 16];          // plain [m][k] table (tests / reference restatement)
  float dctq[kDctQuads * 32 * 4]; // float4 [i*32 + lane]: flattened (slot, k) = divmod(4 i + c, 13); 0 for empty filters (the
                                  // clamp-delta paths read it from global memory: they are rare)
  PitchRegTables pr;
  int mel_nnz;
  int pad[3];
};

inline int build_feature_tables(FeatureTables& ft) {
  std::memset(&ft, 0, sizeof(ft));
  SmemTables& t = ft.s;
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < kNfftM; ++n) t.win400[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftM));
  for (int k1 = 1; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      if (n2 < 25) {
        const double b = 2.0 * PI * (double)(n2 * k1) / 400.0;
        t.tw400[2 * ((k1 - 1) * 25 + n2)] = (float)std::cos(b);
        t.tw400[2 * ((k1 - 1) * 25 + n2) + 1] = (float)(-std::sin(b));
      }
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
      else ft.dctq[((e / 4) * 32 + lane) * 4 + (e % 4)] = ft.dct[m * 16 + k];
    }
  }
  ft.mel_nnz = nnz;
  // ---- DCT fragments (fp16 hi + lo of every entry: hi = fp16(v), lo = fp16(v - hi))
  for (int kap = 0; kap < 8; ++kap)
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, tq = lane & 3;
      auto val = [&](int k, int m) -> float {
        if (k >= kMfcc) return 0.0f;
        if (t.mel_dead[(m / 32) * 32 + (m % 32)]) return 0.0f;
        return ft.dct[m * 16 + k];
      };
      auto split = [&](float v, int part) -> uint16_t {
        const uint16_t hb = f32_to_f16_bits(v);
        if (part == 0) return hb;
        // value of hb as float
        const uint32_t sign = (hb & 0x8000u) << 16, e = (hb >> 10) & 0x1fu, mant = hb & 0x3ffu;
        float hv;
        if (e == 0) hv = std::ldexp((float)mant, -24);
        else hv = std::ldexp((float)(mant | 0x400u), (int)e - 25);
        if (sign) hv = -hv;
        return f32_to_f16_bits(v - hv);
      };
      for (int part = 0; part < 2; ++part) {
        uint32_t* o = t.dct_frag[part][kap][lane];
        const int m0 = 16 * kap + 2 * tq;
        o[0] = (uint32_t)split(val(g, m0), part) | ((uint32_t)split(val(g, m0 + 1), part) << 16);
        o[1] = (uint32_t)split(val(g + 8, m0), part) | ((uint32_t)split(val(g + 8, m0 + 1), part) << 16);
        o[2] = (uint32_t)split(val(g, m0 + 8), part) | ((uint32_t)split(val(g, m0 + 9), part) << 16);
        o[3] = (uint32_t)split(val(g + 8, m0 + 8), part) | ((uint32_t)split(val(g + 8, m0 + 9), part) << 16);
      }
    }
  // ---- fragments of the tensor-core STFT-512 round trip
  auto hann = [&](int n) { return 0.5 - 0.5 * std::cos(2.0 * PI * n / kNfftP); };
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, tq = lane & 3;
    for (int sn = 0; sn < 2; ++sn)
      for (int ks = 0; ks < 2; ++ks)
        for (int m = 0; m < 4; ++m) {
          auto v = [&](int row) {                                      // matrix [n2 or k2][k2 or n2], symmetric
            const double a = 2.0 * PI * (double)((row * (8 * m + g)) % 32) / 32.0;
            return (sn ? std::sin(a) : std::cos(a)) / 16.0;
          };
          uint32_t* o = t.pt.cs[sn * 8 + ks * 4 + m][lane];
          o[0] = f16x2_bits(v(16 * ks + 2 * tq), v(16 * ks + 2 * tq + 1));
          o[1] = f16x2_bits(v(16 * ks + 2 * tq + 8), v(16 * ks + 2 * tq + 9));
        }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        const int n = 32 * ((g + 4 * i) % 16) + 8 * j + 2 * tq;
        t.pt.ws[i][lane][j] = f16x2_bits(hann(n), hann(n + 1));
      }
    for (int f = 0; f < 4; ++f) {
      const int a = f >> 1;                                            // rotation of the output rows: rho = (n1 + 4 a) % 16
      auto v = [&](int rho, int k1) {
        const int n1 = (rho - 4 * a + 16) % 16;
        const double ang = 2.0 * PI * (double)((n1 * k1) % 16) / 16.0;
        const double c = std::cos(ang) * 0.5, s = std::sin(ang) * 0.5;
        return f == 0 ? c : (f == 1 ? -s : (f == 2 ? c : s));
      };
      uint32_t* o = t.pt.a4[f][lane];
      o[0] = f16x2_bits(v(g, 2 * tq), v(g, 2 * tq + 1));
      o[1] = f16x2_bits(v(g + 8, 2 * tq), v(g + 8, 2 * tq + 1));
      o[2] = f16x2_bits(v(g, 2 * tq + 8), v(g, 2 * tq + 9));
      o[3] = f16x2_bits(v(g + 8, 2 * tq + 8), v(g + 8, 2 * tq + 9));
    }
    for (int f = 0; f < 3; ++f) {
      auto v = [&](int k1, int n1) {
        const double ang = 2.0 * PI * (double)((n1 * k1) % 16) / 16.0;
        return f == 0 ? std::cos(ang) : (f == 1 ? -std::sin(ang) : std::sin(ang));
      };
      uint32_t* o = ft.pr.a1[f][lane];
      o[0] = f16x2_bits(v(g, 2 * tq), v(g, 2 * tq + 1));
      o[1] = f16x2_bits(v(g + 8, 2 * tq), v(g + 8, 2 * tq + 1));
      o[2] = f16x2_bits(v(g, 2 * tq + 8), v(g, 2 * tq + 9));
      o[3] = f16x2_bits(v(g + 8, 2 * tq + 8), v(g + 8, 2 * tq + 9));
    }
    for (int j = 0; j < 4; ++j) {
      for (int h = 0; h < 2; ++h) {
        const int k1 = g + 8 * h, n2 = 8 * j + 2 * tq;
        const double a0 = 2.0 * PI * (double)(k1 * n2) / 512.0, a1 = 2.0 * PI * (double)(k1 * (n2 + 1)) / 512.0;
        ft.pr.tw[2 * j + h][lane][0] = f16x2_bits(std::cos(a0), std::cos(a1));
        ft.pr.tw[2 * j + h][lane][1] = f16x2_bits(std::sin(a0), std::sin(a1));
      }
      const int n2 = 8 * j + g;
      ft.pr.wa[j][lane][0] = f16x2_bits(hann(32 * (2 * tq) + n2), hann(32 * (2 * tq + 1) + n2));
      ft.pr.wa[j][lane][1] = f16x2_bits(hann(32 * (2 * tq + 8) + n2), hann(32 * (2 * tq + 9) + n2));
    }
  }
  return bad;   // 0 when the compile-time trip counts match the filter bank
}

}  // namespace msa
