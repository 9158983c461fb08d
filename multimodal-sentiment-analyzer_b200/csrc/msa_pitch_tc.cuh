// "Pitch" feature on the tensor cores: STFT 512/128 -> identity vocoder -> ISTFT of one segment, |x - x^| statistics.
// Restates torchaudio.transforms.PitchShift(n_steps = 0) as called by audio_analyzer.py:175-188.
//
// The reference value is the MEAN OF A Z-SCORE of the round-trip residual, i.e. zero up to fp32 rounding (~1e-9) for
// any residual; its tolerance is 1e-6 absolute.  The round trip therefore runs in fp16 on the tensor cores (mma.sync
// m16n8k16, fp16 accumulate): the residual becomes fp16 rounding noise (~1e-3 of the signal) instead of fp32 rounding
// noise, the feature stays ~1e-8, and the FMA pipe is left to the MFCC pass.
//
// Work unit = one WARP x one QUAD of four consecutive frames = two complex 512-point transforms with two real frames
// packed in each (re, im).  512 = 16 x 32 (n = 32 n1 + n2, k = k1 + 16 k2); every stage is a small matrix product whose
// DATA operand is a register fragment produced by the stage before it (the accumulator layout of m16n8k16 is the
// A-operand layout of the next product), so a transform never leaves the registers:
//   stage 1   Y[k1][n2]  = sum_n1 W16[k1][n1] z[n1][n2]        A = W16 (registers), B = windowed samples (ldmatrix.trans
//                                                              from the warp's fp16 ring of the padded signal)
//   twiddle   Y'         = Y * W512^(k1 n2)                     fp16 pairs, lane-local
//   stage 2   Z[k1][k2]  = sum_n2 Y'[k1][n2] W32[n2][k2] / 16   A = Y' (registers), B = cos|sin fragments (shared memory)
//   (phase_vocoder(rate = 1) returns its input)
//   stage 3   O[k1][n2]  = sum_k2 Z[k1][k2] conj W32[k2][n2] / 16
//   twiddle   O'         = O * conj W512^(k1 n2)
//   transpose O' -> B operand (movmatrix.trans per 8 x 8 tile)
//   stage 4   y[rho][n2] = sum_k1 V16[rho][k1] O'[k1][n2] / 2   rows ROTATED by the frame's phase: rho = (n1 + 4 (f % 4)) % 16
// The rotation makes the 16 x 32 output tile of every frame land on the same ring of 16 rows (= 512 samples) of
// overlap-add accumulators: a frame adds window * y to all 16 rows, completes the 4 oldest (one hop-block of 128
// samples, which is compared with the input and zeroed) and starts 4 new ones.  Every warp owns a contiguous run of
// quads and warms the ring up with the quad before its run.
#pragma once
#include "msa_fft.cuh"
#include "msa_half.h"
#include "msa_tables.hpp"

namespace msa {

constexpr int kRing = 2048;     // fp16 samples of the padded signal a warp keeps staged (positions modulo kRing)

// Where padded position p lives in the ring (in samples).  The ring is 64 rows of 32 samples (64 bytes); ldmatrix reads
// one 16-byte chunk from each of 8 CONSECUTIVE rows, which would hit only two of the eight 16-byte bank groups, so the
// four chunks of row R are stored XOR-swizzled by (R / 2) % 4: eight consecutive rows then cover all eight groups
// (ldmatrix and the read-back of the input in the residual are conflict-free).
MSA_FN int ring_off(int p) {
  const int R = p >> 5, c = (p >> 3) & 3;
  return ((R & 63) << 5) | ((c ^ ((R >> 1) & 3)) << 3) | (p & 7);
}

#define MSA_R(expr) [&](int li_) { return (expr); }

// x: the segment; T samples; [q_begin, q_end): this CTA's quads; ring: this warp's kRing fp16 samples (16-byte aligned)
template <class Env, class InT>
MSA_KFN void pitch_tc(Env& env, const InT* x, int T, int q_begin, int q_end, uint16_t* ring, const PitchSmemTables* pt,
                      const PitchRegTables* gt, double (&ps)[Env::kStates], double (&pq)[Env::kStates],
                      double (&pn)[Env::kStates], float (&pmax)[Env::kStates]) {
  constexpr int S = Env::kStates;
  const int NW = env.nwarps;
  const int nFp = T / kHopP + 1;
  const int wper = (q_end - q_begin + NW - 1) / NW;
  const int wq_lo = q_begin + env.warp * wper;
  const int wq_begin = (wq_lo < q_end) ? wq_lo : q_end;
  const int wq_end = (wq_begin + wper < q_end) ? wq_begin + wper : q_end;
  if (wq_begin >= wq_end) return;
  const int wq_first = (wq_begin > 0) ? wq_begin - 1 : wq_begin;       // warm-up quad: only its overlap-add tail is used

  // reflect-101 padded signal of torch.stft(center=True, pad_mode="reflect"); 0 outside the padding
  auto xr = [&](int t) -> float {
    if (t < 0) t = -t;
    else if (t >= T) t = 2 * (T - 1) - t;
    return (t >= 0 && t < T) ? env.ld(x + t) : 0.0f;
  };

  // ---- per-lane constant fragments (registers)
  u32 a1[S][3][4], tc[S][8], ts[S][8], wa[S][4][2];
  env.lanes([&](int lane, int li) {
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int i = 0; i < 4; ++i) a1[li][f][i] = env.ldu(&gt->a1[f][lane][i]);
#pragma unroll
    for (int r = 0; r < 8; ++r) { tc[li][r] = env.ldu(&gt->tw[r][lane][0]); ts[li][r] = env.ldu(&gt->tw[r][lane][1]); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { wa[li][j][0] = env.ldu(&gt->wa[j][lane][0]); wa[li][j][1] = env.ldu(&gt->wa[j][lane][1]); }
  });
  u32 acc[S][4][2];                                                     // overlap-add ring: rows rho = g (0) / g + 8 (1), columns 8 j + 2 t
  // the residual's sums stay in fp32 for the warp's whole run (a few hundred terms of similar size per lane; the feature
  // built from them is rounding residue by construction) and reach the fp64 block reduction once: the fp64 pipe of this
  // chip is narrow, and three DADDs per pair of frames showed up as 4 % of the kernel's stall samples
  float rs1[S], rs2[S], rmx[S];
  int rcnt[S];
  for (int i = 0; i < S; ++i) {
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0u;
    rs1[i] = rs2[i] = rmx[i] = 0.0f;
    rcnt[i] = 0;
  }

  // ---- staging: 512 padded positions per step, 16 per lane, as fp16 into the ring, one quad ahead of their use.
  // (Splitting a step into loads at the top of a quad and conversion + stores at its end, to hide the L2 latency behind
  // the quad's matrix products, costs 16 live registers: measured slower, 0.448 vs 0.420 ms for this part alone.)
  int staged = kHopP * 4 * wq_first;                                    // first padded position not yet in the ring
  auto stage_chunk = [&]() {
    env.lanes([&](int lane, int li) {
      (void)li;
      const int p = staged + 16 * lane, t0 = p - kNfftP / 2;
      float v[16];
      if (t0 >= 0 && t0 + 16 <= T) {
        env.ld16(x, t0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = xr(t0 + i);
      }
      u32 w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = h2_pack(v[2 * i], v[2 * i + 1]);
      env.st8(ring + ring_off(p), w);                                   // two chunks of one row
      env.st8(ring + ring_off(p + 8), w + 4);
    });
    staged += 512;
  };

  for (int quad = wq_first; quad < wq_end; ++quad) {
    const int f0 = 4 * quad;
    // this quad reads positions [128 f0, 128 f0 + 896); the chunk behind it is staged one quad ahead
    while (staged < kHopP * f0 + 896 + 512 && staged < kHopP * (4 * wq_end) + 896) stage_chunk();
    env.wsync();
    const bool owned = quad >= wq_begin;

    static_for<0, 2>([&](auto uc) {
      constexpr int U = decltype(uc)::value;
      const int fa = f0 + 2 * U;                                        // frames fa (real part) and fa + 1 (imaginary part)
      const bool oka = fa < nFp, okb = fa + 1 < nFp;

      // ---- B fragments of both frames: z[n1][n2] = window * padded signal, tile j = columns 8 j .. 8 j + 7
      u32 bre[S][4][2], bim[S][4][2];
      {
        const uint16_t* rp[S];
#pragma unroll
        for (int fr = 0; fr < 2; ++fr)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            env.lanes([&](int lane, int li) {
              const int n1 = (lane & 7) + 8 * ((lane >> 3) & 1), tile = 2 * jj + (lane >> 4);
              rp[li] = ring + ring_off(kHopP * (fa + fr) + 32 * n1 + 8 * tile);
            });
            if (fr == 0) env.ldsm4t(MSA_R(&bre[li_][2 * jj][0]), MSA_R(rp[li_]));
            else env.ldsm4t(MSA_R(&bim[li_][2 * jj][0]), MSA_R(rp[li_]));
          }
        env.lanes([&](int lane, int li) {
          (void)lane;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              bre[li][j][i] = oka ? h2_mul(bre[li][j][i], wa[li][j][i]) : 0u;
              bim[li][j][i] = okb ? h2_mul(bim[li][j][i], wa[li][j][i]) : 0u;
            }
        });
      }

      // ---- stage 1: Y = W16 z (W16 = cos - i sin)
      u32 yre[S][4][2], yim[S][4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        env.mma(MSA_R(yre[li_][j]), MSA_R(a1[li_][0]), MSA_R(bre[li_][j]), false);
        env.mma(MSA_R(yim[li_][j]), MSA_R(a1[li_][0]), MSA_R(bim[li_][j]), false);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        env.mma(MSA_R(yre[li_][j]), MSA_R(a1[li_][2]), MSA_R(bim[li_][j]), true);
        env.mma(MSA_R(yim[li_][j]), MSA_R(a1[li_][1]), MSA_R(bre[li_][j]), true);
      }
      // ---- twiddle W512^(k1 n2) = tc - i ts; the negated real part feeds the "- Y're sin" products of stage 2
      u32 nyre[S][4][2];
      env.lanes([&](int lane, int li) {
        (void)lane;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const u32 yr = yre[li][j][h], yi = yim[li][j][h], c = tc[li][2 * j + h], s = ts[li][2 * j + h];
            const u32 re = h2_fma(yr, c, h2_mul(yi, s));
            yim[li][j][h] = h2_fms(yi, c, h2_mul(yr, s));
            yre[li][j][h] = re;
            nyre[li][j][h] = h2_neg(re);
          }
      });

      // ---- stage 2: Z = Y' W32 / 16 (W32 = cos - i sin): K = (re | im) x n2 = 4 steps of 16, N = k2 tiles m
      u32 zre[S][4][2], zim[S][4][2], nzim[S][4][2];
      // four rounds (cos | sin) x (n2 < 16 | n2 >= 16) of eight INDEPENDENT products: a warp never waits for its own accumulator
      static_for<0, 4>([&](auto rc) {
        constexpr int R = decltype(rc)::value;
        u32 b[S][4][2];
        env.lanes([&](int lane, int li) {
#pragma unroll
          for (int m = 0; m < 4; ++m) env.lds2(b[li][m], pt->cs[4 * R + m][lane]);
        });
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (R == 0) {
            env.mma(MSA_R(zre[li_][m]), MSA_R(&yre[li_][0][0]), MSA_R(b[li_][m]), false);
            env.mma(MSA_R(zim[li_][m]), MSA_R(&yim[li_][0][0]), MSA_R(b[li_][m]), false);
          } else if (R == 1) {
            env.mma(MSA_R(zre[li_][m]), MSA_R(&yre[li_][2][0]), MSA_R(b[li_][m]), true);
            env.mma(MSA_R(zim[li_][m]), MSA_R(&yim[li_][2][0]), MSA_R(b[li_][m]), true);
          } else if (R == 2) {
            env.mma(MSA_R(zre[li_][m]), MSA_R(&yim[li_][0][0]), MSA_R(b[li_][m]), true);
            env.mma(MSA_R(zim[li_][m]), MSA_R(&nyre[li_][0][0]), MSA_R(b[li_][m]), true);
          } else {
            env.mma(MSA_R(zre[li_][m]), MSA_R(&yim[li_][2][0]), MSA_R(b[li_][m]), true);
            env.mma(MSA_R(zim[li_][m]), MSA_R(&nyre[li_][2][0]), MSA_R(b[li_][m]), true);
          }
        }
      });
      env.lanes([&](int lane, int li) {
        (void)lane;
#pragma unroll
        for (int m = 0; m < 4; ++m) { nzim[li][m][0] = h2_neg(zim[li][m][0]); nzim[li][m][1] = h2_neg(zim[li][m][1]); }
      });

      // ---- stage 3: O = Z conj(W32) / 16: K = (re | im) x k2, N = n2 tiles j
      u32 ore[S][4][2], oim[S][4][2];
      static_for<0, 4>([&](auto rc) {
        constexpr int R = decltype(rc)::value;
        u32 b[S][4][2];
        env.lanes([&](int lane, int li) {
#pragma unroll
          for (int j = 0; j < 4; ++j) env.lds2(b[li][j], pt->cs[4 * R + j][lane]);
        });
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (R == 0) {
            env.mma(MSA_R(ore[li_][j]), MSA_R(&zre[li_][0][0]), MSA_R(b[li_][j]), false);
            env.mma(MSA_R(oim[li_][j]), MSA_R(&zim[li_][0][0]), MSA_R(b[li_][j]), false);
          } else if (R == 1) {
            env.mma(MSA_R(ore[li_][j]), MSA_R(&zre[li_][2][0]), MSA_R(b[li_][j]), true);
            env.mma(MSA_R(oim[li_][j]), MSA_R(&zim[li_][2][0]), MSA_R(b[li_][j]), true);
          } else if (R == 2) {
            env.mma(MSA_R(ore[li_][j]), MSA_R(&nzim[li_][0][0]), MSA_R(b[li_][j]), true);
            env.mma(MSA_R(oim[li_][j]), MSA_R(&zre[li_][0][0]), MSA_R(b[li_][j]), true);
          } else {
            env.mma(MSA_R(ore[li_][j]), MSA_R(&nzim[li_][2][0]), MSA_R(b[li_][j]), true);
            env.mma(MSA_R(oim[li_][j]), MSA_R(&zre[li_][2][0]), MSA_R(b[li_][j]), true);
          }
        }
      });
      // ---- conjugate twiddle, then every 8 x 8 tile is transposed into the B-operand layout of stage 4
      env.lanes([&](int lane, int li) {
        (void)lane;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const u32 orr = ore[li][j][h], oi = oim[li][j][h], c = tc[li][2 * j + h], s = ts[li][2 * j + h];
            ore[li][j][h] = h2_fms(orr, c, h2_mul(oi, s));
            oim[li][j][h] = h2_fma(oi, c, h2_mul(orr, s));
          }
      });
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          env.movmt(MSA_R(&ore[li_][j][h]));
          env.movmt(MSA_R(&oim[li_][j][h]));
        }

      // ---- stage 4: y = V16 O' / 2 (V16 = cos + i sin), output rows rotated by the frame's phase (2 U and 2 U + 1);
      // the fragments of phases 2, 3 are those of phases 0, 1 with the row halves exchanged
      u32 vre_e[S][4], nvim_e[S][4], vre_o[S][4], vim_o[S][4];
      env.lanes([&](int lane, int li) {
        u32 r[4][4];
#pragma unroll
        for (int f = 0; f < 4; ++f) env.lds4(r[f], pt->a4[f][lane]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int s = (U == 0) ? i : (i ^ 1);
          vre_e[li][i] = r[0][s]; nvim_e[li][i] = r[1][s]; vre_o[li][i] = r[2][s]; vim_o[li][i] = r[3][s];
        }
      });
      u32 ye[S][4][2], yo[S][4][2];                                     // frame fa (phase 2 U) and frame fa + 1 (phase 2 U + 1)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        env.mma(MSA_R(ye[li_][j]), MSA_R(vre_e[li_]), MSA_R(ore[li_][j]), false);
        env.mma(MSA_R(yo[li_][j]), MSA_R(vre_o[li_]), MSA_R(oim[li_][j]), false);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        env.mma(MSA_R(ye[li_][j]), MSA_R(nvim_e[li_]), MSA_R(oim[li_][j]), true);
        env.mma(MSA_R(yo[li_][j]), MSA_R(vim_o[li_]), MSA_R(ore[li_][j]), true);
      }

      // ---- synthesis window and overlap-add; the phase-a frame completes ring rows 4 a .. 4 a + 3
      u32 done[S][4];
      env.lanes([&](int lane, int li) {
        const int g = lane >> 2;
        constexpr int ae = 2 * U, ao = 2 * U + 1;
        u32 w0[4], w1[4];
        env.lds4(w0, pt->ws[(0 - ae) & 3][lane]);
        env.lds4(w1, pt->ws[(2 - ae) & 3][lane]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[li][j][0] = h2_fma(w0[j], ye[li][j][0], acc[li][j][0]);
          acc[li][j][1] = h2_fma(w1[j], ye[li][j][1], acc[li][j][1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          done[li][j] = acc[li][j][U];
          if (g < 4) acc[li][j][U] = 0u;
        }
        env.lds4(w0, pt->ws[(0 - ao) & 3][lane]);
        env.lds4(w1, pt->ws[(2 - ao) & 3][lane]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[li][j][0] = h2_fma(w0[j], yo[li][j][0], acc[li][j][0]);
          acc[li][j][1] = h2_fma(w1[j], yo[li][j][1], acc[li][j][1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (g >= 4) { done[li][j] = acc[li][j][U]; acc[li][j][U] = 0u; }
        }
      });

      // ---- finished hop-blocks fa (lanes g < 4) and fa + 1 (lanes g >= 4): compare with the input
      // (torch.istft divides by the overlap-added squared window and trims the n_fft / 2 padding: t = position - 256)
      if (owned) {
        const bool fast = (fa >= 3) && (fa + 1 <= nFp - 1) && (kHopP * (fa + 2) - kNfftP / 2 <= T);
        env.lanes([&](int lane, int li) {
          const int g = lane >> 2, tq = lane & 3;
          const int blk = fa + (g >> 2);
          float s1 = 0.0f, s2 = 0.0f, mx = 0.0f;
          int cnt = 0;
          if (fast) {
            // four frames cover the block: sum of the squared periodic Hann windows at hop N / 4 is exactly 3 / 2
            const u32 k23 = h2_pack(2.0f / 3.0f, 2.0f / 3.0f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int p = kHopP * blk + 32 * (g & 3) + 8 * j + 2 * tq;
              const u32 xv = env.lds1(reinterpret_cast<const u32*>(ring + ring_off(p)));
              const u32 r = h2_abs(h2_fnma(done[li][j], k23, xv));
              const float r0 = h2_lo(r), r1 = h2_hi(r);
              s1 += r0 + r1;
              s2 = fmaf(r0, r0, fmaf(r1, r1, s2));
              mx = fmaxf(mx, fmaxf(r0, r1));
            }
            cnt = 8;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int o = 32 * (g & 3) + 8 * j + 2 * tq + e;
                const int t = kHopP * blk + o - kNfftP / 2;
                if (t >= 0 && t < T) {
                  float en = 0.0f;
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const int f = blk - jj;
                    if (f >= 0 && f < nFp) {
                      const float w = 0.5f - 0.5f * env.cospi((float)(jj * kHopP + o) * (1.0f / 256.0f));
                      en = fmaf(w, w, en);
                    }
                  }
                  const float y = (e == 0 ? h2_lo(done[li][j]) : h2_hi(done[li][j])) / en;
                  const float pv = fabsf(xr(t) - y);
                  s1 += pv; s2 = fmaf(pv, pv, s2); mx = fmaxf(mx, pv); ++cnt;
                }
              }
          }
          rs1[li] += s1; rs2[li] += s2; rcnt[li] += cnt;
          rmx[li] = fmaxf(rmx[li], mx);
        });
      }
    });
    env.wsync();                                                        // the ring positions behind this quad may be overwritten
  }
  for (int i = 0; i < S; ++i) {
    ps[i] += (double)rs1[i]; pq[i] += (double)rs2[i]; pn[i] += (double)rcnt[i];
    pmax[i] = fmaxf(pmax[i], rmx[i]);
  }
}

}  // namespace msa
