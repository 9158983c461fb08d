// Fused audio-feature kernel body: one thread-block CLUSTER per 5 s segment.
//
// Restates, for one mono segment [T], what AudioAnalyzer computes per call
// (/root/reference/src/analyzers/audio_analyzer.py):
//   _analyze_pitch        :175-188  STFT 512/128 -> identity vocoder -> ISTFT, |x - x^|, z-score, mean
//   _analyze_intensity    :190-201  NaN for mono (std of one element)
//   _analyze_timbre       :203-217  MFCC(13) -> global z-score -> time mean
//   _analyze_speech_rate  :219-233
//   _analyze_rhythm       :235-263  400/160 frame energies -> mean, unbiased std, L/16000
//   _calculate_signal_noise_ratio :278-293, _calculate_clarity :295-311,
//   _calculate_consistency :313-329, _calculate_audio_quality :265-276
//   AudioFeatureNormalizer.normalize (src/utils/normalization.py:26-44) -> LayerNorm(31)
//   audio row for fusion (src/processors/streaming_processor.py:250-268, 295-298)
//
// Work split: the segment's samples [0,T) are cut into `nranks` contiguous slices; CTA `rank`
// stages its slice (+512-sample halos) in shared memory ONCE (the only HBM read of the
// waveform) and computes, from shared memory, every MFCC frame whose centre, every output
// sample of the STFT->ISTFT round trip and every 80-sample energy atom that falls in its slice.
// The whole-segment dependencies (top_db max, z-score moments, frame-energy statistics) are
// exchanged through distributed shared memory; rank 0 assembles the 31-float row.
//
// FFTs: one warp per transform, in place in shared memory, two real frames packed into one
// complex transform (msa_fft.cuh).  Warps run independently; only the overlap-add of the ISTFT
// needs block barriers (once per batch of 2*nwarps frames).
//
// This file is compiled by nvcc (GpuEnv, msa_features.cu) and by g++ (CpuEnv, tests/emu)
// so the index logic can be exercised without a GPU.  It must only use the Env primitives.
#pragma once
#include <cstdint>
#include "msa_fft.cuh"
#include "msa_hd.h"
#include "msa_tables.hpp"

namespace msa {

constexpr int kHalo = 512;
constexpr int kDbStride = kMels + 1;       // 129: conflict-free row-per-thread reads in the DCT step
constexpr int kDetailStride = 96;

enum : int { kPartWave = 1, kPartMfcc = 2, kPartPitch = 4, kPartAll = 7 };
enum : int { kFlagStrictNan = 1, kFlagBulkCopy = 2 };

struct FeatParams {
  const void* wav;       // [B, T] fp32 or int16
  int is_s16;
  int B, T;
  int slice_len;         // samples per cluster rank, multiple of 80
  int noise_n;           // int(0.05 * T) computed by the host exactly like the reference
  const float* emo8;     // [B, 8] or null -> 1/8
  float* feat31;         // [B, 31]  LN31[:27] ++ quality4, nan_to_num'd (fusion input row)
  float* detail;         // [B, 96]  or null: raw27, quality4, ln31, diagnostics
  float* dbg_mfcc;       // [B, nFm, 13] or null
  const FeatureTables* tab;
  int flags;
  int parts;
};

// tables staged in shared memory (the DCT matrix stays in global: warp-uniform reads through L1)
struct SmemTables {
  c32 tw512_s1[7 * 64];
  c32 tw512_s2[7 * 8];
  c32 tw400_s1[15 * 25];
  c32 tw400_s2[4 * 5];
  float win400[kNfftM];
  float win512[kNfftP];
  float mel_w[kMelNnzMax];
  uint16_t mel_pos[2 * kMelNnzMax];
  uint16_t mel_ptr[kMels + 2];
};

struct Partials {
  double mf_sum[kMfcc];
  double mf_sumsq, mf_abs_lo, mf_abs_hi;
  double p_n, p_sum, p_sumsq;
  double e_total, e_noise;
  float db_max, p_max;
  int mf_frames, n_atoms;
};

struct FeatLayout {
  int wave_off, zb_off, dbs_off, mfcc_off, carry_off, atoms_off, tab_off, red_off, part_off, bar_off;
  int wave_cap, dbs_rows, total;
};

MSA_FN int ceil_div(int a, int b) { return (a + b - 1) / b; }

// identical on host (launch configuration) and device (carve-up)
inline
#ifdef __CUDACC__
__host__ __device__
#endif
FeatLayout feat_layout(int slice_len, int nwarps) {
  FeatLayout l;
  int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
  l.wave_cap = slice_len + 2 * kHalo;
  l.dbs_rows = slice_len / kHopM + 2;
  l.wave_off = take(l.wave_cap * 4);
  l.zb_off = take(nwarps * kPad512 * 8);
  l.dbs_off = take(l.dbs_rows * kDbStride * 4);
  l.mfcc_off = take(l.dbs_rows * kMfcc * 4);
  l.carry_off = take(2 * 3 * kHopP * 4);
  l.atoms_off = take((slice_len / kAtom + 1) * 4);
  l.tab_off = take((int)sizeof(SmemTables));
  l.red_off = take(64 * 8);
  l.part_off = take((int)sizeof(Partials));
  l.bar_off = take(16);
  l.total = off;
  return l;
}

template <class InT> MSA_FN float to_f32(InT v);
template <> MSA_FN float to_f32<float>(float v) { return v; }
template <> MSA_FN float to_f32<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }

// python semantics of min(max(v, 0), 1): max(v,0) returns v unless 0 > v; min(w,1) returns w unless 1 < w
MSA_FN double py_clip01(double v) {
  double w = (0.0 > v) ? 0.0 : v;
  return (1.0 < w) ? 1.0 : w;
}

template <class Env, class InT>
MSA_KFN void features_cta(Env& env, const FeatParams& P, unsigned char* smem) {
  constexpr int LANES = Env::kLanes;
  const int T = P.T, L = P.slice_len;
  const int seg = env.cluster_id, r = env.rank;
  const int t0 = (r * L < T) ? r * L : T;
  const int t1 = (t0 + L < T) ? t0 + L : T;
  const bool has = t1 > t0;
  const int lo = (t0 - kHalo > 0) ? t0 - kHalo : 0;
  const int hi = (t1 + kHalo < T) ? t1 + kHalo : T;
  const FeatLayout lay = feat_layout(L, env.nwarps);

  float* wave = reinterpret_cast<float*>(smem + lay.wave_off);
  c32* zb_all = reinterpret_cast<c32*>(smem + lay.zb_off);
  float* dbs = reinterpret_cast<float*>(smem + lay.dbs_off);
  float* mfcc = reinterpret_cast<float*>(smem + lay.mfcc_off);
  float* carry = reinterpret_cast<float*>(smem + lay.carry_off);
  float* atoms = reinterpret_cast<float*>(smem + lay.atoms_off);
  SmemTables* tb = reinterpret_cast<SmemTables*>(smem + lay.tab_off);
  double* red = reinterpret_cast<double*>(smem + lay.red_off);
  Partials* part = reinterpret_cast<Partials*>(smem + lay.part_off);

  // reflect-101 indexing of torch.stft(center=True, pad_mode="reflect") into the staged slice
  auto W = [&](int t) -> float {
    if (t < 0) t = -t;
    else if (t >= T) t = 2 * (T - 1) - t;
    return wave[t - lo];
  };

  // ---------------------------------------------------------------- stage tables + slice
  {
    const FeatureTables* g = P.tab;
    const c32* s1 = reinterpret_cast<const c32*>(g->tw512_s1);
    const c32* s2 = reinterpret_cast<const c32*>(g->tw512_s2);
    const c32* m1 = reinterpret_cast<const c32*>(g->tw400_s1);
    const c32* m2 = reinterpret_cast<const c32*>(g->tw400_s2);
    for (int i = env.tid; i < 7 * 64; i += env.nthreads) tb->tw512_s1[i] = s1[i];
    for (int i = env.tid; i < 7 * 8; i += env.nthreads) tb->tw512_s2[i] = s2[i];
    for (int i = env.tid; i < 15 * 25; i += env.nthreads) tb->tw400_s1[i] = m1[i];
    for (int i = env.tid; i < 4 * 5; i += env.nthreads) tb->tw400_s2[i] = m2[i];
    for (int i = env.tid; i < kNfftM; i += env.nthreads) tb->win400[i] = g->win400[i];
    for (int i = env.tid; i < kNfftP; i += env.nthreads) tb->win512[i] = g->win512[i];
    for (int i = env.tid; i < kMelNnzMax; i += env.nthreads) {
      tb->mel_w[i] = g->mel_w[i];
      tb->mel_pos[2 * i] = g->mel_pos[2 * i];
      tb->mel_pos[2 * i + 1] = g->mel_pos[2 * i + 1];
    }
    for (int i = env.tid; i <= kMels; i += env.nthreads) tb->mel_ptr[i] = g->mel_ptr[i];
    for (int i = env.tid; i < 2 * 3 * kHopP; i += env.nthreads) carry[i] = 0.0f;
  }
  if (has) {
    const InT* src = reinterpret_cast<const InT*>(P.wav) + (size_t)seg * T + lo;
    env.template load_slice<InT>(wave, src, hi - lo, smem + lay.bar_off, (P.flags & kFlagBulkCopy) != 0);
  }
  env.sync();

  // ---------------------------------------------------------------- K1: energy atoms, totals
  double e_total = 0.0, e_noise = 0.0;
  int n_atoms_local = 0;
  if (has && (P.parts & kPartWave)) {
    const int full_end = T - T % kAtom;
    const int a0 = t0 / kAtom;
    const int a1 = ((t1 < full_end) ? t1 : full_end) / kAtom;
    n_atoms_local = (a1 > a0) ? a1 - a0 : 0;
    for (int a = a0 + env.warp; a < a1; a += env.nwarps) {
      float s = 0.0f;
      for (int i = env.lane; i < kAtom; i += LANES) { float x = wave[a * kAtom + i - lo]; s = fmaf(x, x, s); }
      double sd = env.wsum((double)s);
      if (env.lane == 0) atoms[a - a0] = (float)sd;
    }
    const int nn = P.noise_n;
    for (int t = t0 + env.tid; t < t1; t += env.nthreads) {
      double x = wave[t - lo];
      double x2 = x * x;
      e_total += x2;
      if (t < nn || t >= T - nn) e_noise += x2;
    }
  }
  e_total = env.bsum(e_total, red);
  e_noise = env.bsum(e_noise, red);

  // ---------------------------------------------------------------- K2 phase A: STFT-400 -> power -> mel -> dB
  const int nFm = T / kHopM + 1;
  int fm_begin = 0, fm_end = 0;
  if (has) {
    fm_begin = ceil_div(t0, kHopM);
    fm_end = (t1 == T) ? nFm : ceil_div(t1, kHopM);
  }
  const int nfr = (P.parts & kPartMfcc) ? (fm_end - fm_begin) : 0;
  float dbmax = -3.0e38f;
  {
    c32* zb = zb_all + env.warp * kPad512;
    const int npairs = (nfr + 1) / 2;
    for (int pp = env.warp; pp < npairs; pp += env.nwarps) {
      const int fa = fm_begin + 2 * pp;
      const bool hasb = (2 * pp + 1) < nfr;
      const int ca = fa * kHopM - kNfftM / 2, cb = ca + kHopM;
      if (ca >= 0 && cb + kNfftM <= T) {                 // interior pair: no reflection (warp-uniform branch)
        const float* wa = wave + (ca - lo);
        for (int n = env.lane; n < kNfftM; n += LANES) {
          const float w = tb->win400[n];
          zb[n] = c32{w * wa[n], hasb ? w * wa[n + kHopM] : 0.0f};
        }
      } else {
        for (int n = env.lane; n < kNfftM; n += LANES) {
          const float w = tb->win400[n];
          zb[n] = c32{w * W(ca + n), hasb ? w * W(cb + n) : 0.0f};
        }
      }
      env.wsync();
      fft_stage<kNfftM, 16, 400, false, PadNone, LANES>(zb, tb->tw400_s1, env.lane);
      env.wsync();
      fft_stage<kNfftM, 5, 25, false, PadNone, LANES>(zb, tb->tw400_s2, env.lane);
      env.wsync();
      fft_stage<kNfftM, 5, 5, false, PadNone, LANES>(zb, nullptr, env.lane);
      env.wsync();
      // mel energies straight from the packed spectrum: with Z = FFT(a + i b),
      //   |A_k|^2 = |Z_k + conj Z_{N-k}|^2 / 4,  |B_k|^2 = |Z_k - conj Z_{N-k}|^2 / 4
      for (int m = env.lane; m < kMels; m += LANES) {
        const int p0 = tb->mel_ptr[m], p1 = tb->mel_ptr[m + 1];
        float ea = 0.0f, eb = 0.0f;
        for (int p = p0; p < p1; ++p) {
          const c32 zk = zb[tb->mel_pos[2 * p]];
          const c32 zn = zb[tb->mel_pos[2 * p + 1]];
          const float w = 0.25f * tb->mel_w[p];
          const float sx = zk.x + zn.x, sy = zk.y - zn.y;
          const float dx = zk.x - zn.x, dy = zk.y + zn.y;
          ea = fmaf(w, fmaf(sx, sx, sy * sy), ea);
          eb = fmaf(w, fmaf(dx, dx, dy * dy), eb);
        }
        const float da = 10.0f * log10f(fmaxf(ea, 1e-10f));
        dbs[(2 * pp) * kDbStride + m] = da;
        dbmax = fmaxf(dbmax, da);
        if (hasb) {
          const float db = 10.0f * log10f(fmaxf(eb, 1e-10f));
          dbs[(2 * pp + 1) * kDbStride + m] = db;
          dbmax = fmaxf(dbmax, db);
        }
      }
      env.wsync();
    }
  }

  // ---------------------------------------------------------------- K3: STFT-512 -> ISTFT residual
  double p_n = 0.0, p_sum = 0.0, p_sumsq = 0.0;
  float p_max = 0.0f;
  if (P.parts & kPartPitch) {
    const int nFp = T / kHopP + 1;
    // samples [t0,t1) sit at padded positions [t0+256, t1+256); position tp is covered by frames tp/128-3 .. tp/128
    const bool any = has;
    const int pf_lo = ((t0 + kNfftP / 2) >> 7) - 3;
    const int pf_hi = (t1 + kNfftP / 2 - 1) >> 7;
    const int pf_begin = (pf_lo > 0) ? pf_lo : 0;
    const int pf_end = (pf_hi < nFp - 1) ? pf_hi : nFp - 1;             // inclusive
    const int FB = 2 * env.nwarps;
    const float inv_n = 1.0f / (float)kNfftP;
    float* cin = carry;
    float* cout = carry + 3 * kHopP;
    c32* zb = zb_all + env.warp * kPad512;
    for (int fb0 = pf_begin; any && fb0 <= pf_end; fb0 += FB) {
      const int fa = fb0 + 2 * env.warp;
      if (fa <= pf_end) {
        const bool hasb = fa + 1 <= pf_end;
        const int sa = fa * kHopP - kNfftP / 2, sb = sa + kHopP;
        if (sa >= 0 && sb + kNfftP <= T) {
          const float* wa = wave + (sa - lo);
          for (int n = env.lane; n < kNfftP; n += LANES) {
            const float w = tb->win512[n];
            zb[Pad8::at(n)] = c32{w * wa[n], hasb ? w * wa[n + kHopP] : 0.0f};
          }
        } else {
          for (int n = env.lane; n < kNfftP; n += LANES) {
            const float w = tb->win512[n];
            zb[Pad8::at(n)] = c32{w * W(sa + n), hasb ? w * W(sb + n) : 0.0f};
          }
        }
        env.wsync();
        fft_stage<kNfftP, 8, 512, false, Pad8, LANES>(zb, tb->tw512_s1, env.lane);
        env.wsync();
        fft_stage<kNfftP, 8, 64, false, Pad8, LANES>(zb, tb->tw512_s2, env.lane);
        env.wsync();
        fft_stage<kNfftP, 8, 8, false, Pad8, LANES>(zb, nullptr, env.lane);
        env.wsync();
        // phase_vocoder(rate = 1.0) returns its input: the spectrum goes straight back
        fft_stage<kNfftP, 8, 8, true, Pad8, LANES>(zb, nullptr, env.lane);
        env.wsync();
        fft_stage<kNfftP, 8, 64, true, Pad8, LANES>(zb, tb->tw512_s2, env.lane);
        env.wsync();
        fft_stage<kNfftP, 8, 512, true, Pad8, LANES>(zb, tb->tw512_s1, env.lane);
      }
      env.sync();
      // overlap-add by gathering: every padded position sums the <= 4 frames of this batch that cover it
      const int f_hi = (fb0 + FB - 1 < pf_end) ? fb0 + FB - 1 : pf_end;
      for (int i = env.tid; i < (FB + 3) * kHopP; i += env.nthreads) {
        const int tp = fb0 * kHopP + i;                 // position in the reflect-padded signal
        const int fq = tp >> 7, nq = tp & (kHopP - 1);
        float s = (i < 3 * kHopP) ? cin[i] : 0.0f;
#pragma unroll
        for (int j = 3; j >= 0; --j) {
          const int f = fq - j;
          if (f >= fb0 && f <= f_hi) {
            const int n = nq + j * kHopP;
            const c32 z = zb_all[((f - fb0) >> 1) * kPad512 + Pad8::at(n)];
            s = fmaf(tb->win512[n] * inv_n, ((f - fb0) & 1) ? z.y : z.x, s);
          }
        }
        if (i < FB * kHopP) {
          const int t = tp - kNfftP / 2;
          if (t >= t0 && t < t1) {
            float env_w = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int f = fq - j;
              if (f >= 0 && f < nFp) { const float w = tb->win512[nq + j * kHopP]; env_w = fmaf(w, w, env_w); }
            }
            const float xh = s / env_w;
            const float pv = fabsf(wave[t - lo] - xh);
            p_n += 1.0; p_sum += (double)pv; p_sumsq += (double)pv * (double)pv;
            p_max = fmaxf(p_max, pv);
          }
        } else {
          cout[i - FB * kHopP] = s;
        }
      }
      env.sync();
      float* tmp = cin; cin = cout; cout = tmp;
    }
  }
  p_n = env.bsum(p_n, red);
  p_sum = env.bsum(p_sum, red);
  p_sumsq = env.bsum(p_sumsq, red);
  p_max = env.bmax(p_max, red);
  dbmax = env.bmax(dbmax, red);

  if (env.tid == 0) {
    part->db_max = dbmax; part->p_max = p_max;
    part->p_n = p_n; part->p_sum = p_sum; part->p_sumsq = p_sumsq;
    part->e_total = e_total; part->e_noise = e_noise;
    part->mf_frames = nfr; part->n_atoms = n_atoms_local;
  }
  env.csync();                                            // #1: every rank's dB maximum is visible

  float gmax = -3.0e38f;
  for (int rr = 0; rr < env.nranks; ++rr) gmax = fmaxf(gmax, env.remote(part, rr)->db_max);

  // ---------------------------------------------------------------- K2 phase B: top_db clamp -> DCT -> moments
  {
    const float thr = gmax - 80.0f;
    const float* dct = P.tab->dct;
    // one thread per (frame, group of 4 coefficients): 13 = 4 + 4 + 4 + 1
    for (int it = env.tid; it < nfr * 4; it += env.nthreads) {
      const int fl = it >> 2, kg = it & 3;
      const float* row = dbs + fl * kDbStride;
      const float* dk = dct + kg * 4;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
      for (int m = 0; m < kMels; ++m) {
        const float v = fmaxf(row[m], thr);
        const float* d = dk + m * kDctStride;
        a0 = fmaf(v, d[0], a0);
        a1 = fmaf(v, d[1], a1);                         // columns 13..15 of the padded table are zero
        a2 = fmaf(v, d[2], a2);
        a3 = fmaf(v, d[3], a3);
      }
      float* o = mfcc + fl * kMfcc + kg * 4;
      o[0] = a0;
      if (kg < 3) { o[1] = a1; o[2] = a2; o[3] = a3; }
    }
    env.sync();
    if (P.dbg_mfcc) {
      float* o = P.dbg_mfcc + ((size_t)seg * nFm + fm_begin) * kMfcc;
      for (int i = env.tid; i < nfr * kMfcc; i += env.nthreads) o[i] = mfcc[i];
    }
    double ss = 0.0, alo = 0.0, ahi = 0.0;
    for (int i = env.tid; i < nfr * kMfcc; i += env.nthreads) {
      const double v = mfcc[i];
      ss += v * v;
      if (i % kMfcc < 6) alo += fabs(v); else ahi += fabs(v);
    }
    ss = env.bsum(ss, red); alo = env.bsum(alo, red); ahi = env.bsum(ahi, red);
    // per-coefficient time sums: warp w reduces coefficient k = w, w + nwarps, ...
    for (int k = env.warp; k < kMfcc; k += env.nwarps) {
      double sk = 0.0;
      for (int fl = env.lane; fl < nfr; fl += LANES) sk += (double)mfcc[fl * kMfcc + k];
      sk = env.wsum(sk);
      if (env.lane == 0) part->mf_sum[k] = sk;
    }
    if (env.tid == 0) { part->mf_sumsq = ss; part->mf_abs_lo = alo; part->mf_abs_hi = ahi; }
  }
  env.csync();                                            // #2: all partial moments and atoms are visible

  // ---------------------------------------------------------------- rank 0: merge and assemble the row
  if (r == 0) {
    // gather the energy atoms of all ranks (re-using the FFT buffers, idle by now)
    float* all_atoms = reinterpret_cast<float*>(zb_all);
    const int atoms_cap = env.nwarps * kPad512 * 2;
    int nA = 0;
    for (int rr = 0; rr < env.nranks; ++rr) {
      const Partials* rp = env.remote(part, rr);
      const float* ra = env.remote(atoms, rr);
      const int n = rp->n_atoms;
      for (int i = env.tid; i < n && nA + i < atoms_cap; i += env.nthreads) all_atoms[nA + i] = ra[i];
      nA += n;
    }
    env.sync();
    // rhythm: frame energies e_g = sum of 5 atoms at stride 2 (400 = 5*80, 160 = 2*80)
    const int nG = (T >= kRhythmWin) ? (T - kRhythmWin) / kRhythmHop + 1 : 0;
    double gs = 0.0;
    for (int g = env.tid; g < nG; g += env.nthreads) {
      const float* a = all_atoms + 2 * g;
      gs += (double)(a[0] + a[1] + a[2] + a[3] + a[4]);
    }
    gs = env.bsum(gs, red);
    const double gmean = (nG > 0) ? gs / nG : 0.0;
    double gq = 0.0;
    for (int g = env.tid; g < nG; g += env.nthreads) {
      const float* a = all_atoms + 2 * g;
      const double d = (double)(a[0] + a[1] + a[2] + a[3] + a[4]) - gmean;
      gq += d * d;
    }
    gq = env.bsum(gq, red);
    // consistency: 1600-sample block mean-squares = 20 atoms / 1600
    const int nBk = T / kBlock;
    double bs = 0.0;
    for (int b = env.tid; b < nBk; b += env.nthreads) {
      float acc = 0.0f;
      for (int j = 0; j < 20; ++j) acc += all_atoms[20 * b + j];
      bs += (double)(acc * (1.0f / (float)kBlock));
    }
    bs = env.bsum(bs, red);
    const double bmean = (nBk > 0) ? bs / nBk : 0.0;
    double bq = 0.0;
    for (int b = env.tid; b < nBk; b += env.nthreads) {
      float acc = 0.0f;
      for (int j = 0; j < 20; ++j) acc += all_atoms[20 * b + j];
      const double d = (double)(acc * (1.0f / (float)kBlock)) - bmean;
      bq += d * d;
    }
    bq = env.bsum(bq, red);

    if (env.tid == 0) {
      const double NaN = nan("");
      Partials tot = *part;
      for (int rr = 1; rr < env.nranks; ++rr) {
        const Partials* rp = env.remote(part, rr);
        for (int k = 0; k < kMfcc; ++k) tot.mf_sum[k] += rp->mf_sum[k];
        tot.mf_sumsq += rp->mf_sumsq; tot.mf_abs_lo += rp->mf_abs_lo; tot.mf_abs_hi += rp->mf_abs_hi;
        tot.p_n += rp->p_n; tot.p_sum += rp->p_sum; tot.p_sumsq += rp->p_sumsq;
        tot.e_total += rp->e_total; tot.e_noise += rp->e_noise;
        tot.p_max = fmaxf(tot.p_max, rp->p_max);
        tot.mf_frames += rp->mf_frames;
      }
      float raw[27];
      const float* emo = P.emo8 ? P.emo8 + (size_t)seg * 8 : nullptr;
      for (int k = 0; k < 8; ++k) raw[k] = emo ? emo[k] : 0.125f;
      // pitch: mean of the z-scored residual. mu and sigma are fp32 tensors in the reference, so the
      // value is the rounding residue of mu; the residual itself is fp32 FFT noise (~1e-8).
      double p_mean = 0.0, p_std = 0.0;
      {
        float v = 0.0f;
        if (tot.p_n > 1.0) {
          p_mean = tot.p_sum / tot.p_n;
          double var = (tot.p_sumsq - tot.p_sum * p_mean) / (tot.p_n - 1.0);
          p_std = sqrt(var > 0.0 ? var : 0.0);
          const float mu32 = (float)p_mean, sd32 = (float)p_std;
          v = (float)((tot.p_sum - tot.p_n * (double)mu32) / (tot.p_n * ((double)sd32 + 1e-6)));
        }
        raw[8] = v;
      }
      // intensity: (e - mean(e)) / (std(e) + 1e-6) over ONE channel: std of one element is NaN
      raw[9] = (P.flags & kFlagStrictNan) ? (float)NaN : 0.0f;
      // timbre
      double clarity = 0.0;
      {
        const double n = (double)tot.mf_frames * kMfcc;
        if (tot.mf_frames > 0) {
          double ssum = 0.0;
          for (int k = 0; k < kMfcc; ++k) ssum += tot.mf_sum[k];
          const double mu = ssum / n;
          const double var = (n > 1.0) ? (tot.mf_sumsq - ssum * mu) / (n - 1.0) : NaN;
          const double sd = sqrt(var > 0.0 ? var : (var == var ? 0.0 : NaN));
          for (int k = 0; k < kMfcc; ++k) raw[10 + k] = (float)((tot.mf_sum[k] / tot.mf_frames - mu) / (sd + 1e-6));
          const double hi_m = tot.mf_abs_hi / (7.0 * tot.mf_frames), lo_m = tot.mf_abs_lo / (6.0 * tot.mf_frames);
          clarity = py_clip01((double)((float)hi_m / ((float)lo_m + 1e-6f)));
        } else {
          for (int k = 0; k < kMfcc; ++k) raw[10 + k] = 0.0f;
        }
      }
      // speech rate (mono): energy > 0.1 * energy in fp32
      {
        const float e = (float)tot.e_total;
        raw[23] = (e > e * 0.1f) ? 1.0f : 0.0f;
      }
      // rhythm
      if (nG > 0) {
        raw[24] = (float)gmean;
        raw[25] = (nG > 1) ? (float)sqrt(gq / (nG - 1)) : (float)NaN;
        raw[26] = (float)((double)nG / (double)kSampleRate);
      } else {
        raw[24] = raw[25] = raw[26] = 0.0f;
      }
      // quality scalars (python floats in the reference: double arithmetic on fp32 .item() values)
      double snr = 0.0, consistency = 0.0;
      if (P.noise_n > 0) {
        const float noise_p = (float)(tot.e_noise / (2.0 * P.noise_n));
        const float sig_p = (float)(tot.e_total / (double)T);
        const float snr_db = 10.0f * log10f(sig_p / (noise_p + 1e-6f));
        snr = py_clip01((double)snr_db / 30.0);
      }
      if (nBk > 0) {
        const float sd = (nBk > 1) ? (float)sqrt(bq / (nBk - 1)) : (float)NaN;
        const double cv = (double)(sd / ((float)bmean + 1e-6f));
        consistency = 1.0 - ((1.0 < cv) ? 1.0 : cv);   // python min(cv, 1.0): NaN stays NaN
      }
      if (!(P.parts & kPartMfcc)) clarity = 0.0;
      const double quality = 0.4 * snr + 0.3 * clarity + 0.3 * consistency;
      const float q4[4] = {(float)quality, (float)snr, (float)clarity, (float)consistency};

      // AudioFeatureNormalizer: pad 27 -> 31 with zeros, LayerNorm(31) (gamma 1, beta 0, eps 1e-5, biased var)
      float ln[31];
      {
        double m = 0.0;
        for (int k = 0; k < 27; ++k) m += (double)raw[k];
        m /= 31.0;
        double v = 0.0;
        for (int k = 0; k < 31; ++k) { const double d = ((k < 27) ? (double)raw[k] : 0.0) - m; v += d * d; }
        v /= 31.0;
        const double rs = 1.0 / sqrt(v + 1e-5);
        for (int k = 0; k < 31; ++k) ln[k] = (float)((((k < 27) ? (double)raw[k] : 0.0) - m) * rs);
      }
      // fusion input row: LN slices ++ quality, torch.nan_to_num(nan=0.0) (+-inf -> +-FLT_MAX)
      float* out = P.feat31 + (size_t)seg * 31;
      for (int k = 0; k < 31; ++k) {
        float v = (k < 27) ? ln[k] : q4[k - 27];
        if (v != v) v = 0.0f;
        else if (v > 3.4028234663852886e38f) v = 3.4028234663852886e38f;
        else if (v < -3.4028234663852886e38f) v = -3.4028234663852886e38f;
        out[k] = v;
      }
      if (P.detail) {
        float* d = P.detail + (size_t)seg * kDetailStride;
        for (int k = 0; k < 27; ++k) d[k] = raw[k];
        for (int k = 0; k < 4; ++k) d[27 + k] = q4[k];
        d[31] = 0.0f;
        for (int k = 0; k < 31; ++k) d[32 + k] = ln[k];
        d[63] = 0.0f;
        d[64] = gmax; d[65] = (float)p_mean; d[66] = (float)p_std; d[67] = tot.p_max;
        d[68] = (float)tot.e_total; d[69] = (float)tot.e_noise; d[70] = (float)tot.mf_frames; d[71] = (float)nG;
        d[72] = (float)tot.p_n; d[73] = (float)nBk; d[74] = (float)nA;
        for (int k = 75; k < kDetailStride; ++k) d[k] = 0.0f;
      }
    }
  }
  env.csync();                                            // #3: nobody exits while rank 0 still reads its smem
}

}  // namespace msa
