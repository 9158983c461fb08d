// Fused audio-feature kernel body: one thread-block CLUSTER (1, 2, 4 or 8 CTAs) per 5 s segment.
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
// Work unit = one WARP x one QUAD of four consecutive frames.
//
//   "pitch"  (n_fft 512, hop 128; msa_pitch_tc.cuh)  STFT -> vocoder at rate 1.0 (identity) -> ISTFT as four chained
//            mma.sync stages per pair of frames: 512 = 16 x 32, the DFT-16 / DFT-32 factors and their conjugates are
//            constant fp16 A operands, the accumulator layout of one stage is the B layout of the next, the
//            inter-stage twiddles are half2 FMAs, and the last stage's rows are rotated so that the overlap-add of
//            consecutive frames stays in registers.  Every warp owns a contiguous run of quads, stages the waveform
//            as fp16 into a swizzled 2048-sample ring of its own and reads the original samples back from it for
//            |x - x^|.  The feature is the mean of a z-scored sequence, 0 up to 1e-6 whatever the residual, so fp16
//            tensor-core arithmetic is free; the residual's moments still accumulate in fp32 / fp64.
//   MFCC     (n_fft 400, hop 160; msa_fft.cuh)  two real frames packed per complex FFT, every pass in registers:
//            pass A lane n2: 16 strided samples x window -> radix-16 -> twiddle -> smem [k1][n2];  pass B lane k1: one
//            row -> radix-25.  The waveform is read straight from global memory (each 128-byte line reaches the SM
//            from HBM once and is re-read from L1 by the overlapping frames).
//            Power of both packed frames from Z_k and Z_(N-k), sparse mel (each lane owns 4 filters), dB, then the
//            DCT 128 -> 13 of the quad's four frames as mma.sync with the dB values split into fp16 hi + lo words
//            (three products, fp32 accumulate: error below 1e-6 of the coefficients).
//            top_db needs the segment maximum, which is only known after the last frame: the DCT is
//            linear, so the four EMPTY mel filters (max(-100, max-80) dB in every frame) enter as one
//            constant vector afterwards, and the few live values that can fall below max-80 dB (known
//            in advance from an upper bound of the maximum) go on per-warp candidate lists and are
//            patched in as the DCT of their clamp deltas; a list that overflows redoes its quads
//            (or re-reads their dB values from an optional scratch table) with the deltas inline.
//
// Whole-segment dependencies (top_db maximum, z-score moments, frame-energy statistics) are
// exchanged through distributed shared memory; rank 0 assembles the 31-float row.
//
// This file is compiled by nvcc (GpuEnv, msa_features.cu) and by g++ (CpuEnv, tests/emu: the 32
// lanes of a warp run sequentially between warp barriers, one OS thread per warp) so the index
// logic can be exercised without a GPU.  It must only use the Env primitives.
#pragma once
#include <cstdint>
#include <cstring>
#include "msa_fft.cuh"
#include "msa_hd.h"
#include "msa_tables.hpp"
#include "msa_pitch_tc.cuh"

namespace msa {

constexpr int kDetailStride = 96;
constexpr int kRow512 = 33, kRow400 = 25;        // row strides (complex words) of the pass-A -> pass-B tiles
constexpr int kFftHalf = 16 * kRow512;           // spacing (complex words) of the two power-spectrum row pairs in a warp's buffer
constexpr int kGroup = 640;                      // wave-statistics unit: 8 energy atoms, 5 float4 per lane
constexpr int kRedSlots = 16;
#ifndef MSA_K1_GROUPS
#define MSA_K1_GROUPS 2
#endif
constexpr int kK1Groups = MSA_K1_GROUPS;                     // load groups in flight per warp in the wave-statistics phase (160 floats of scratch each; 4 in flight measured no faster)
constexpr int kTileM = 16 * kRow400;             // one MFCC FFT tile: 400 complex = 3200 bytes
constexpr int kShareBytes = 52 * 33 * 4;         // DCT-share transpose tile (the largest MFCC use of a warp's buffer)
// top_db candidates of one warp live behind the transpose tile: 198 entries of (frame << 7 | filter, dB)
constexpr int kClampCap = 166;
constexpr int kDbRow = 136;                      // row stride (32-bit words) of the quad's [frame][filter] dB table: 128 + 8 keeps the
                                                 // four rows on different banks for the 8-byte fragment loads
// a warp's private buffer: MFCC FFT tiles (2 x 3200 B), power rows, the transpose tile + candidate list; the wave-statistics
// scratch and the fp16 ring of the STFT-512 round trip (4096 B) reuse it in their own phases.  Sized so that the whole
// layout of a 5 s segment stays below half an SM's shared memory (two CTAs per SM).
constexpr int kWarpBufBytes = kShareBytes + kClampCap * 8;
static_assert(kWarpBufBytes % 16 == 0 && kWarpBufBytes >= 2 * kFftHalf * 4 + 2 * kPowStride * 4 && kWarpBufBytes >= kRing * 2, "warp buffer");

enum : int { kPartWave = 1, kPartMfcc = 2, kPartPitch = 4, kPartAll = 7 };
enum : int { kFlagStrictNan = 1 };

struct FeatParams {
  const void* wav;       // [B, T] fp32 or int16
  int is_s16;
  int B, T;
  int noise_n;           // int(0.05 * T) computed by the host exactly like the reference
  const float* emo8;     // [B, 8] or null -> 1/8
  float* feat31;         // [B, 31]  LN31[:27] ++ quality4, nan_to_num'd (fusion input row)
  float* detail;         // [B, 96]  or null: raw27, quality4, ln31, diagnostics
  float* dbg_mfcc;       // [B, nFm, 13] or null
  float* dbscratch;      // [B, ceil(nFm / 4), 16, 32] or null: mel dB values of the quads whose top_db candidates did not fit
                         // their warp's list, so that applying the clamp later needs no second FFT
  const FeatureTables* tab;
  int flags;
  int parts;
};

struct Partials {
  double mf_sum[kMfcc];
  double mf_sumsq, mf_abs_lo, mf_abs_hi;
  double p_n, p_sum, p_sumsq;
  double e_total, e_noise, e_left;
  float db_max, db_min, p_max, a_max;
  int mf_frames, n_atoms, slow_pass;
};

struct FeatLayout {
  int buf_off, mfl_off, atoms_off, tab_off, wred_off, out_off, part_off, ctr_off;
  int mfl_frames, atoms_cap, total;
};

MSA_FN int ceil_div(int a, int b) { return (a + b - 1) / b; }
MSA_FN int float_as_int(float f) {
#ifdef __CUDACC__
  return __float_as_int(f);
#else
  int v;
  std::memcpy(&v, &f, 4);
  return v;
#endif
}
MSA_FN float int_as_float(int v) {
#ifdef __CUDACC__
  return __int_as_float(v);
#else
  float f;
  std::memcpy(&f, &v, 4);
  return f;
#endif
}

// STFT-512 quads (4 frames each) of a segment, and how many of them a rank of the cluster takes
inline
#ifdef __CUDACC__
__host__ __device__
#endif
int pitch_quads(int T) {
  const int nFp = T / kHopP + 1;
  const int nBl = (T - 1 + kNfftP / 2) / kHopP + 1;
  return ((nFp > nBl ? nFp : nBl) + 3) / 4;
}
// identical on host (launch configuration) and device (carve-up)
inline
#ifdef __CUDACC__
__host__ __device__
#endif
FeatLayout feat_layout(int T, int nranks, int nwarps) {
  FeatLayout l;
  int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
  const int nFm = T / kHopM + 1;
  l.mfl_frames = 4 * ((((nFm + 3) / 4) + nranks - 1) / nranks);
  l.atoms_cap = 8 * ((((T + kGroup - 1) / kGroup) + nranks - 1) / nranks);
  int buf = nwarps * kWarpBufBytes;
  // rank 0 gathers all energy atoms and every rank's Partials there at the end
  const int gather = (((T / kAtom + 8) * 4 + 15) & ~15) + (nranks > 8 ? nranks : 8) * (int)sizeof(Partials);
  if (buf < gather) buf = gather;
  l.buf_off = take(buf);
  // the MFCC rows share their region with the "pitch" phase's per-warp copy of the quad's input samples
  int mx = l.mfl_frames * kMfcc * 4;
  if (mx < nwarps * 4 * kHopP * 4) mx = nwarps * 4 * kHopP * 4;
  l.mfl_off = take(mx);
  l.atoms_off = take(l.atoms_cap * 4);
  l.tab_off = take((int)sizeof(SmemTables));
  l.wred_off = take(kRedSlots * nwarps * 8);
  l.out_off = take(kRedSlots * 8);
  l.part_off = take((int)sizeof(Partials));
  l.ctr_off = take(16);
  l.total = off;
  return l;
}

// python semantics of min(max(v, 0), 1): max(v,0) returns v unless 0 > v; min(w,1) returns w unless 1 < w
MSA_FN double py_clip01(double v) {
  double w = (0.0 > v) ? 0.0 : v;
  return (1.0 < w) ? 1.0 : w;
}

// ---- FFT passes (see msa_fft.cuh) ------------------------------------------------------------
// pass A forward: 16 packed samples -> radix-16 -> twiddle W_N^(n2 k1) -> tile[k1][n2]  (tw: [k1 - 1][n2], row stride ROW)
template <int ROW> MSA_FN void pass_a_fwd(c32* z, const c32* tw, int n2, c32* tile) {
  dft16<false>(z);
  tile[n2] = z[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) tile[k1 * ROW + n2] = cmul(z[k1], tw[(k1 - 1) * ROW + n2]);
}
enum : int { kOpSum = 0, kOpMax = 1, kOpMin = 2 };
MSA_FN double red_comb(int op, double a, double b) {
  return op == kOpSum ? a + b : (op == kOpMax ? (a > b ? a : b) : (a < b ? a : b));
}

// Deterministic block reduction of K per-lane doubles: out[k] (shared memory) = op_k over all threads.
// Level 1 (reduce_warp_stage) is the Env's warp reduction (shuffles on the GPU) into the warp's scratch slots
// slot0 .. slot0 + K - 1, level 2 (reduce_block_stage) combines the per-warp values of slots 0 .. K - 1 behind a
// barrier.  The two levels can be apart: a phase deposits its per-warp values without a barrier and a later
// phase's reduction picks them up.
template <int K, class Env, class Get>
MSA_KFN void reduce_warp_stage(Env& env, double* wred, int slot0, const int* ops, Get get) {
  const int NW = env.nwarps;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = env.warp_reduce(ops[k], [&](int li) { return get(li, k); });
    env.lanes([&](int lane, int li) { (void)li; if (lane == 0) wred[(slot0 + k) * NW + env.warp] = r; });
  }
}
template <int K, class Env>
MSA_KFN void reduce_block_stage(Env& env, double* wred, double* out, const int* ops) {
  static_assert(K <= kRedSlots, "reduction scratch");
  const int NW = env.nwarps;
  env.sync();
  if (env.warp == 0) {
    env.lanes([&](int lane, int li) {
      (void)li;
      if (lane < K) {
        double a = wred[lane * NW];
        for (int w = 1; w < NW; ++w) a = red_comb(ops[lane], a, wred[lane * NW + w]);
        out[lane] = a;
      }
    });
  }
  env.sync();
}
template <int K, class Env, class Get>
MSA_KFN void block_reduce(Env& env, double* wred, double* out, const int* ops, Get get) {
  reduce_warp_stage<K>(env, wred, 0, ops, get);
  reduce_block_stage<K>(env, wred, out, ops);
}

template <class Env, class InT>
MSA_KFN void features_cta(Env& env, const FeatParams& P, unsigned char* smem) {
  constexpr int S = Env::kStates;           // per-lane state copies: 1 on the GPU (registers), 32 in the CPU emulation
  const int T = P.T;
  const int seg = env.cluster_id, r = env.rank, NR = env.nranks, NW = env.nwarps;
  const FeatLayout lay = feat_layout(T, NR, NW);
  const InT* x = reinterpret_cast<const InT*>(P.wav) + (size_t)seg * T;

  c32* wbuf = reinterpret_cast<c32*>(smem + lay.buf_off + env.warp * kWarpBufBytes);   // this warp's private buffer
  double* wred = reinterpret_cast<double*>(smem + lay.wred_off);
  float* mfl = reinterpret_cast<float*>(smem + lay.mfl_off);
  float* atoms = reinterpret_cast<float*>(smem + lay.atoms_off);
  const SmemTables* tb = reinterpret_cast<const SmemTables*>(smem + lay.tab_off);
  double* rout = reinterpret_cast<double*>(smem + lay.out_off);
  Partials* part = reinterpret_cast<Partials*>(smem + lay.part_off);
  int* ctr = reinterpret_cast<int*>(smem + lay.ctr_off);
  const c32* tw400 = reinterpret_cast<const c32*>(tb->tw400);

  // reflect-101 padded signal of torch.stft(center=True, pad_mode="reflect"); 0 outside the padding
  auto xr = [&](int t) -> float {
    if (t < 0) t = -t;
    else if (t >= T) t = 2 * (T - 1) - t;
    return (t >= 0 && t < T) ? env.ld(x + t) : 0.0f;
  };

  // ---------------------------------------------------------------- stage the constant tables
  env.copy16(smem + lay.tab_off, &P.tab->s, (int)sizeof(SmemTables));
  if (env.warp == 0) env.lanes([&](int lane, int li) { (void)li; if (lane < 4) ctr[lane] = 0; });   // ctr[2]: candidate-list overflow flag
  env.sync();

  // ---------------------------------------------------------------- K1: energy atoms, totals
  // groups of 640 samples (8 atoms of 80): lane l loads float4 j at sample 128 j + 4 l, a 160-entry
  // shared-memory tile per group turns the per-lane partial sums into per-atom sums (20 entries each)
  int n_atoms_local = 0;
  double e_tot[S], e_noi[S], e_left[S];
  float a_max[S];
  for (int i = 0; i < S; ++i) { e_tot[i] = e_noi[i] = e_left[i] = 0.0; a_max[i] = 0.0f; }
  if (P.parts & kPartWave) {
    const int nG = ceil_div(T, kGroup);
    const int gper = ceil_div(nG, NR);
    const int g0 = (r * gper < nG) ? r * gper : nG;
    const int g1 = (g0 + gper < nG) ? g0 + gper : nG;
    const int full_atoms = T / kAtom;
    const int a_lo = g0 * 8;
    const int a_hi = (g1 * 8 < full_atoms) ? g1 * 8 : full_atoms;
    n_atoms_local = (a_hi > a_lo) ? a_hi - a_lo : 0;
    float* scr = reinterpret_cast<float*>(wbuf);
    // kK1Groups groups of 640 samples are in flight per warp and step: this phase is pure load latency
    for (int gp = g0 + kK1Groups * env.warp; gp < g1; gp += kK1Groups * NW) {
      const int ng = (g1 - gp < kK1Groups) ? g1 - gp : kK1Groups;
      // every 4 consecutive samples form one partial sum (entry = sample offset / 4 of the kK1Groups * 640 samples of
      // the step); an atom is 20 consecutive entries.  fp32 input: one 16-byte load per entry; int16 input: one
      // 16-byte load per TWO entries, so the bytes in flight (what this phase lives on) stay the same.  Entries are
      // formed and added identically for both input types (bit-identical results).
      env.lanes([&](int lane, int li) {
        constexpr int kPer = (sizeof(InT) == 2) ? 8 : 4;                       // samples per load
        constexpr int kLoads = kK1Groups * kGroup / (32 * kPer);               // loads per lane and step
        float v[kLoads][kPer];
        const int first = gp * kGroup, limit = (gp + ng) * kGroup;             // samples of this step
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
          const int off = 32 * kPer * k + kPer * lane;
          if (first + off < limit) env.ldv(x, first + off, T, v[k]);
          else {
#pragma unroll
            for (int i = 0; i < kPer; ++i) v[k][i] = 0.0f;
          }
        }
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < kLoads; ++k)
#pragma unroll
          for (int h4 = 0; h4 < kPer / 4; ++h4) {
            const float* q = v[k] + 4 * h4;
            const float sq = fmaf(q[3], q[3], fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0])));
            scr[(32 * kPer * k + kPer * lane) / 4 + h4] = sq;
            tot += (double)sq;
          }
        e_tot[li] += tot;
      });
      env.wsync();
      env.lanes([&](int lane, int li) {
        if (lane < 8 * ng) {
          const int a = gp * 8 + lane;
          float s = 0.0f;
#pragma unroll
          for (int j = 0; j < 20; ++j) s += scr[20 * lane + j];
          if (a < full_atoms) {
            atoms[a - a_lo] = s;
            a_max[li] = fmaxf(a_max[li], s);
          } else {
            e_left[li] += (double)s;                     // the T mod 80 samples behind the last full atom
          }
        }
      });
      env.wsync();
    }
    // noise power: first and last int(0.05 T) samples (audio_analyzer.py:282-285), split over the ranks
    const int nn = P.noise_n;
    const int per = ceil_div(2 * nn, NR);
    const int i0 = (r * per < 2 * nn) ? r * per : 2 * nn;
    const int i1 = (i0 + per < 2 * nn) ? i0 + per : 2 * nn;
    // exact squares summed in fp64: which lane visits which sample depends on the cluster size and the CTA shape,
    // and the result must not (a float running sum per lane would differ in its last bits)
    env.lanes([&](int lane, int li) {
      double acc = 0.0;
      for (int i = i0 + env.warp * 32 + lane; i < i1; i += env.nthreads) {
        const double v = (double)env.ld(x + ((i < nn) ? i : T - 2 * nn + i));
        acc += v * v;
      }
      e_noi[li] = acc;
    });
  }
  // per-warp totals go to scratch slots 4..7 WITHOUT a barrier: every warp moves on to its STFT-512 quads as
  // soon as its own loads are in (this phase is pure load latency), and the block-level sums are formed together
  // with the residual's, behind that phase's barrier
  const int ops[4] = {kOpSum, kOpSum, kOpSum, kOpMax};
  reduce_warp_stage<4>(env, wred, 4, ops, [&](int li, int k) {
    return k == 0 ? e_tot[li] : (k == 1 ? e_noi[li] : (k == 2 ? e_left[li] : (double)a_max[li]));
  });

  // ---------------------------------------------------------------- K3: STFT-512 -> ISTFT residual (tensor cores)
  {
    double ps[S], pq[S], pn[S];
    float pmax[S];
    for (int i = 0; i < S; ++i) { ps[i] = pq[i] = pn[i] = 0.0; pmax[i] = 0.0f; }
    if (P.parts & kPartPitch) {
      const int nQ = pitch_quads(T);
      const int qper = ceil_div(nQ, NR);
      const int q_begin = (r * qper < nQ) ? r * qper : nQ;
      const int q_end = (q_begin + qper < nQ) ? q_begin + qper : nQ;
      // msa_pitch_tc.cuh: every warp owns a contiguous run of this rank's quads; its fp16 ring of the padded signal is
      // the warp's own buffer (the wave-statistics scratch of the same warp is dead by now: program order)
      env.wsync();
      pitch_tc<Env, InT>(env, x, T, q_begin, q_end, reinterpret_cast<uint16_t*>(wbuf), &tb->pt, &P.tab->pr, ps, pq, pn, pmax);
    }
    const int ops[8] = {kOpSum, kOpSum, kOpSum, kOpMax, kOpSum, kOpSum, kOpSum, kOpMax};
    reduce_warp_stage<4>(env, wred, 0, ops, [&](int li, int k) {
      return k == 0 ? ps[li] : (k == 1 ? pq[li] : (k == 2 ? pn[li] : (double)pmax[li]));
    });
    reduce_block_stage<8>(env, wred, rout, ops);          // slots 4..7: the wave-statistics totals deposited above
    if (env.tid == 0) {
      part->p_sum = rout[0]; part->p_sumsq = rout[1]; part->p_n = rout[2]; part->p_max = (float)rout[3];
      part->e_total = rout[4]; part->e_noise = rout[5]; part->e_left = rout[6]; part->a_max = (float)rout[7];
      part->n_atoms = n_atoms_local;
    }
  }

  // ---------------------------------------------------------------- K2: STFT-400 -> power -> mel -> dB -> DCT shares
  const int nFm = T / kHopM + 1;
  const int nQm = ceil_div(nFm, 4);
  const int mper = ceil_div(nQm, NR);
  const int mq_begin = (r * mper < nQm) ? r * mper : nQm;
  const int mq_end = (mq_begin + mper < nQm) ? mq_begin + mper : nQm;
  const int mf_begin = 4 * mq_begin;
  const int mf_end = (4 * mq_end < nFm) ? 4 * mq_end : nFm;
  const int nfr = ((P.parts & kPartMfcc) && mf_end > mf_begin) ? mf_end - mf_begin : 0;

  // MFCC passes over this rank's quads.  Pass 0: no clamp on live filters (thr = -inf) and every live
  // (frame, filter) below `cand` dB goes on the warp's candidate list.  Pass 1 (only when a candidate
  // list overflowed) redoes the quads clamped at the now known threshold.
  int2* clist = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(wbuf) + kShareBytes);
  // An upper bound of the segment's largest mel energy, known before the pass: a frame holds at most 400
  // samples of the reflect-padded signal, i.e. at most 8 energy atoms' worth (twice 4: a padded frame repeats
  // up to 201 samples) plus the ragged end; Parseval with window <= 1 and mel weights <= 1 gives
  // mel <= 400 * E_frame.  Anything 80 dB below the bound or higher can never be clamped by top_db.
  env.csync();                                            // #0: every rank's atom maximum is visible
  float cand = -3.0e38f;
  {
    double amax = 0.0, eleft = 0.0;
    for (int rr = 0; rr < NR; ++rr) {
      const Partials* rp = env.remote(part, rr);
      amax = (amax > (double)rp->a_max) ? amax : (double)rp->a_max;
      eleft += rp->e_left;
    }
    const float bound = (float)(400.0 * (8.0 * amax + 2.0 * eleft)) * 1.001f;
    if (P.parts & kPartWave) cand = 3.0102999566398120f * env.log2(fmaxf(bound, 1e-30f)) - 80.0f + 0.01f;
    else cand = 3.0e38f;                                  // no atoms: every live value is a candidate (overflow -> clamped pass)
  }
  float thr = -3.0e38f, gmax = -3.0e38f, gmin = 3.0e38f;
  int ncand = 0;                                           // warp-uniform length of this warp's candidate list
  int ovf_task = 0x7fffffff;                               // warp-uniform: first task (quad) whose candidates did not fit
  bool fix = false, slow = false, slow_w = false;          // slow: some warp of this CTA overflowed; slow_w: this warp did
  for (int pass = 0; pass < 2; ++pass) {
    const float cand_p = (pass == 0) ? cand : -3.0e38f;
      float dmax[S], dmin[S];
      for (int i = 0; i < S; ++i) { dmax[i] = -3.0e38f; dmin[i] = 3.0e38f; }
      float pw[S][28];
      float share[S][52];
      float* fbuf = reinterpret_cast<float*>(wbuf);
      float cdelta[S][4][4];                                       // [frame][slot] clamp deltas of the current quad
      float dbs[S][4][4];                                          // [frame][slot] mel dB values of the current quad
      float casum[S][2];
      // DCT of the clamp deltas in cdelta, summed over the lanes in lane order -> casum (v = lane + 32 half): the one
      // arithmetic every way of applying the clamp uses (patch list, saved dB values, recomputed quad)
      auto delta_dct = [&]() {
        env.lanes([&](int lane, int li) {
          const float (&cl)[4][4] = cdelta[li];
  #pragma unroll
          for (int v = 0; v < 52; ++v) share[li][v] = 0.0f;
          const float* dq = P.tab->dctq;                           // global memory: this path is rare
          static_for<0, kDctQuads>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const float qc[4] = {env.ldf(dq + (i * 32 + lane) * 4), env.ldf(dq + (i * 32 + lane) * 4 + 1),
                                 env.ldf(dq + (i * 32 + lane) * 4 + 2), env.ldf(dq + (i * 32 + lane) * 4 + 3)};
            static_for<0, 4>([&](auto cc) {
              constexpr int c = decltype(cc)::value;
              constexpr int s = (4 * i + c) / kMfcc, k = (4 * i + c) % kMfcc;
  #pragma unroll
              for (int j = 0; j < 4; ++j) share[li][j * kMfcc + k] = fmaf(cl[j][s], qc[c], share[li][j * kMfcc + k]);
            });
          });
        });
        env.lanes([&](int lane, int li) {
  #pragma unroll
          for (int v = 0; v < 52; ++v) fbuf[v * 33 + lane] = share[li][v];
        });
        env.wsync();
        env.lanes([&](int lane, int li) {
  #pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int v = lane + 32 * half;
            float ca = 0.0f;
            if (v < 52) {
  #pragma unroll
              for (int j = 0; j < 32; ++j) ca += fbuf[v * 33 + j];
            }
            casum[li][half] = ca;
          }
        });
        env.wsync();
      };
      // a contiguous run of quads per warp (adjacent quads share a hop of samples in L1)
      const int ntasks = (nfr > 0) ? mq_end - mq_begin : 0;
      const int tper = ceil_div(ntasks, NW);
      const int t_lo = env.warp * tper;
      const int t_begin = (t_lo < ntasks) ? t_lo : ntasks;
      const int t_end = (t_begin + tper < ntasks) ? t_begin + tper : ntasks;
      // pass 1: only the warps whose list overflowed, from the quad at which it did (earlier quads are patched from the list)
      for (int task = (pass == 0) ? t_begin : (slow_w ? ovf_task : t_end); task < t_end; ++task) {
        const int m0 = 4 * (mq_begin + task);
        if (pass == 1 && P.dbscratch != nullptr) {
          // the quad's dB values were saved by pass 0: clamp deltas -> delta DCT -> add to the rows, no FFT
          const float* sv = P.dbscratch + (((size_t)seg * nQm + (mq_begin + task)) * 16) * 32;
          env.lanes([&](int lane, int li) {
  #pragma unroll
            for (int j = 0; j < 4; ++j)
  #pragma unroll
              for (int s = 0; s < 4; ++s) {
                const float d = sv[(4 * j + s) * 32 + lane];
                const bool live = tb->mel_dead[32 * s + lane] == 0 && (m0 + j) < nFm;
                cdelta[li][j][s] = (live && d < thr) ? thr - d : 0.0f;
              }
          });
          delta_dct();
          env.lanes([&](int lane, int li) {
  #pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int v = lane + 32 * half;
              const int fr = m0 + v / kMfcc;
              if (v < 52 && fr < nFm) mfl[(fr - mf_begin) * kMfcc + v % kMfcc] += casum[li][half];
            }
          });
          env.wsync();
          continue;
        }
        const int s0 = kHopM * m0 - kNfftM / 2;
        const bool interior = (s0 >= 0) && (s0 + 5 * kHopM <= T);
        // the next quad of this warp reads samples s0 + 800 .. s0 + 1799 (the first 200 of them are in L1 already):
        // requested into L1 now, one 128-byte line per lane, so that its first loads do not wait for L2 / HBM
        if (task + 1 < t_end) {
          env.lanes([&](int lane, int li) {
            (void)li;
            constexpr int kPerLine = 128 / (int)sizeof(InT);
            const int tn = s0 + 5 * kHopM + lane * kPerLine;
            if (lane * kPerLine < 4 * kHopM + kPerLine && tn >= 0 && tn < T) env.prefetch_l1(x + tn);
          });
        }
        for (int h = 0; h < 2; ++h) {
          env.lanes([&](int lane, int li) {
            (void)li;
            if (lane < 25) {
              const int sb = s0 + 2 * h * kHopM;
              const bool oka = (m0 + 2 * h) < nFm, okb = (m0 + 2 * h + 1) < nFm;
              float raw[24];                                       // (sharing the quad's 40 samples between the two FFTs costs more in registers than the 8 loads it saves)
              if (interior) {
  #pragma unroll
                for (int i = 0; i < 24; ++i) raw[i] = env.ld_last(x + sb + 25 * i + lane);
              } else {
  #pragma unroll
                for (int i = 0; i < 24; ++i) raw[i] = xr(sb + 25 * i + lane);
              }
              c32 z[16];
              if (interior) {
  #pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                  const float w = tb->win400[25 * n1 + lane];
                  z[n1] = c32{w * raw[n1], w * raw[n1 + 8]};
                }
              } else {
  #pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                  const float w = tb->win400[25 * n1 + lane];
                  z[n1] = c32{oka ? w * raw[n1] : 0.0f, okb ? w * raw[n1 + 8] : 0.0f};
                }
              }
              pass_a_fwd<kRow400>(z, tw400, lane, wbuf + h * kTileM);
            }
          });
        }
        env.wsync();
        env.lanes([&](int lane, int li) {
          (void)li;
          c32* row = wbuf + (lane >> 4) * kTileM + (lane & 15) * kRow400;
          c32 v[25];
  #pragma unroll
          for (int i = 0; i < 25; ++i) v[i] = row[i];
          dft25<false>(v);
  #pragma unroll
          for (int i = 0; i < 25; ++i) row[i] = v[i];                // row[k2] = Z[k1 + 16 k2]
        });
        env.wsync();
        // power of both packed frames: with Z = FFT(a + i b),
        //   |A_k|^2 = |Z_k + conj Z_{N-k}|^2 / 4,  |B_k|^2 = |Z_k - conj Z_{N-k}|^2 / 4
        env.lanes([&](int lane, int li) {
  #pragma unroll
          for (int h = 0; h < 2; ++h) {
            const c32* zt = wbuf + h * kTileM;
  #pragma unroll
            for (int i = 0; i < 7; ++i) {
              const int k = lane + 32 * i;
              float pa = 0.0f, pb = 0.0f;
              if (k < kBinsM) {
                const int kk = (k == 0) ? 0 : kNfftM - k;
                const c32 zk = zt[(k & 15) * kRow400 + (k >> 4)];
                const c32 zn = zt[(kk & 15) * kRow400 + (kk >> 4)];
                const c32 sm = add_conj(zk, zn), df = sub_conj(zk, zn);
                pa = 0.25f * fmaf(sm.x, sm.x, sm.y * sm.y);
                pb = 0.25f * fmaf(df.x, df.x, df.y * df.y);
              }
              pw[li][14 * h + 2 * i] = pa;
              pw[li][14 * h + 2 * i + 1] = pb;
            }
          }
        });
        env.wsync();
        env.lanes([&](int lane, int li) {
  #pragma unroll
          for (int h = 0; h < 2; ++h) {
            float* pr = fbuf + h * (2 * kFftHalf);                    // same bytes as tile h, as floats
  #pragma unroll
            for (int i = 0; i < 7; ++i) {
              const int k = lane + 32 * i;
              if (k < kPowStride) {                                  // bins 201..207 are written as zeros
                pr[k] = pw[li][14 * h + 2 * i];
                pr[kPowStride + k] = pw[li][14 * h + 2 * i + 1];
              }
            }
          }
        });
        env.wsync();
        // mel energies of the lane's 4 filters for the 4 frames, dB, DCT shares
        env.lanes([&](int lane, int li) {
          float (&db)[4][4] = dbs[li];                               // [frame][slot], never clamped: the DCT of the clamp
          float (&cl)[4][4] = cdelta[li];                            // deltas (thr - dB where dB < thr) is added separately
          static_for<0, 4>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int trips = (s == 0) ? kMelTrip0 : (s == 1) ? kMelTrip1 : (s == 2) ? kMelTrip2 : kMelTrip3;
            constexpr int toff = (s == 0) ? 0 : (s == 1) ? kMelTrip0 : (s == 2) ? kMelTrip0 + kMelTrip1 : kMelTrip0 + kMelTrip1 + kMelTrip2;
            const int lo = tb->mel_lo[32 * s + lane];
            const bool dead = tb->mel_dead[32 * s + lane] != 0;
            float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f, e3 = 0.0f;
  #pragma unroll
            for (int p = 0; p < trips; ++p) {
              const float w = tb->mel_w[(toff + p) * 32 + lane];
              e0 = fmaf(w, fbuf[lo + p], e0);
              e1 = fmaf(w, fbuf[kPowStride + lo + p], e1);
              e2 = fmaf(w, fbuf[2 * kFftHalf + lo + p], e2);
              e3 = fmaf(w, fbuf[2 * kFftHalf + kPowStride + lo + p], e3);
            }
            const float e[4] = {e0, e1, e2, e3};
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              float d = 3.0102999566398120f * env.log2(fmaxf(e[j], 1e-10f));   // 10 log10(max(x, amin))
              const bool live = !dead && (m0 + j) < nFm;
              env.push(live && d < cand_p, clist, ncand, kClampCap, ((m0 + j - mf_begin) << 7) | (32 * s + lane), d);
              if (live) {
                dmax[li] = fmaxf(dmax[li], d);
                dmin[li] = fminf(dmin[li], d);
              }
              cl[j][s] = (live && d < thr) ? thr - d : 0.0f;         // thr = -inf in pass 0
              db[j][s] = d;
            }
          });
        });
        // the DCT runs on the tensor cores (below): every dB value goes to shared memory as an fp16 pair (hi, lo) with
        // hi + lo = d to ~2^-22, word [frame][filter] (rows padded to 136 words: conflict-free fragment loads).  The
        // power rows this warp just read are dead; a warp barrier separates the last read from the first write.
        env.wsync();
        env.lanes([&](int lane, int li) {
          u32* dbw = reinterpret_cast<u32*>(fbuf);
  #pragma unroll
          for (int j = 0; j < 4; ++j)
  #pragma unroll
            for (int s = 0; s < 4; ++s) {
              const float d = dbs[li][j][s];
              const float hi = int_as_float(float_as_int(d) & (int)0xFFFFE000);     // 11 significant bits: exact in fp16
              dbw[j * kDbRow + 32 * s + lane] = h2_pack(hi, d - hi);
            }
        });
        env.wsync();
        if (pass == 0) {
          // the list overflowed at this quad (or earlier): from here on the quads' dB values go to the scratch
          // table instead, so pass 1 can apply the clamp without recomputing them
          if (ncand > kClampCap && ovf_task > task) ovf_task = task;
          if (ovf_task <= task && P.dbscratch != nullptr) {
            float* sv = P.dbscratch + (((size_t)seg * nQm + (mq_begin + task)) * 16) * 32;
            env.lanes([&](int lane, int li) {
  #pragma unroll
              for (int j = 0; j < 4; ++j)
  #pragma unroll
                for (int s = 0; s < 4; ++s) sv[(4 * j + s) * 32 + lane] = dbs[li][j][s];
            });
          }
        }
        env.wsync();
        // MFCC[k][frame] = sum_m dct[m][k] dB[frame][m] as 8 k-steps of mma.sync m16n8k16 (fp16 operands, fp32 accumulate):
        // A = DCT fragments (rows = 13 coefficients padded to 16, constant, shared memory), B = the quad's dB values
        // (rows = 16 mel filters of the k-step, columns = frames: lane (g, t) supplies frame g % 4).  Both operands are
        // split hi + lo and the three significant products are issued (hi hi + lo hi + hi lo): ~2^-21 relative.
        {
          float acc[S][4], acc1[S][4], acc2[S][4];                       // three independent accumulation chains
          u32 bh[S][2], bl[S][2], ah[S][4], al[S][4];
          env.lanes([&](int lane, int li) {
            (void)lane;
  #pragma unroll
            for (int i = 0; i < 4; ++i) acc[li][i] = acc1[li][i] = acc2[li][i] = 0.0f;
          });
          const u32* dbw = reinterpret_cast<const u32*>(fbuf);
  #pragma unroll
          for (int kap = 0; kap < 8; ++kap) {
            env.lanes([&](int lane, int li) {
              const int g = lane >> 2, tq = lane & 3;
              u32 w0[2], w1[2];
              env.lds2(w0, dbw + (g & 3) * kDbRow + 16 * kap + 2 * tq);
              env.lds2(w1, dbw + (g & 3) * kDbRow + 16 * kap + 2 * tq + 8);
              bh[li][0] = h2_lows(w0[0], w0[1]); bl[li][0] = h2_highs(w0[0], w0[1]);
              bh[li][1] = h2_lows(w1[0], w1[1]); bl[li][1] = h2_highs(w1[0], w1[1]);
              env.lds4(ah[li], tb->dct_frag[0][kap][lane]);
              env.lds4(al[li], tb->dct_frag[1][kap][lane]);
            });
            env.mma_f32(MSA_R(acc[li_]), MSA_R(ah[li_]), MSA_R(bh[li_]));
            env.mma_f32(MSA_R(acc1[li_]), MSA_R(ah[li_]), MSA_R(bl[li_]));
            env.mma_f32(MSA_R(acc2[li_]), MSA_R(al[li_]), MSA_R(bh[li_]));
          }
          env.lanes([&](int lane, int li) {
            (void)lane;
  #pragma unroll
            for (int i = 0; i < 4; ++i) acc[li][i] += acc1[li][i] + acc2[li][i];     // the two correction terms first
          });
          env.wsync();
          // lane (g, t < 2) holds coefficients g and g + 8 of frames 2 t and 2 t + 1
          env.lanes([&](int lane, int li) {
            const int g = lane >> 2, tq = lane & 3;
            if (tq < 2) {
  #pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int fr = m0 + 2 * tq + e;
                if (fr < nFm) {
                  mfl[(fr - mf_begin) * kMfcc + g] = acc[li][e];
                  if (g + 8 < kMfcc) mfl[(fr - mf_begin) * kMfcc + g + 8] = acc[li][2 + e];
                }
              }
            }
          });
          env.wsync();
        }
        if (pass == 1) {                                             // recomputed quad (no scratch table): add the delta DCT
          delta_dct();
          env.lanes([&](int lane, int li) {
  #pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int v = lane + 32 * half;
              const int fr = m0 + v / kMfcc;
              if (v < 52 && fr < nFm) mfl[(fr - mf_begin) * kMfcc + v % kMfcc] += casum[li][half];
            }
          });
          env.wsync();
        }
      }
    const int ops[2] = {kOpMax, kOpMin};
    block_reduce<2>(env, wred, rout, ops, [&](int li, int k) { return k == 0 ? (double)dmax[li] : (double)dmin[li]; });
    if (pass == 1) break;
    const float dbmax = (float)rout[0], dbmin = (float)rout[1];
    if (ncand > kClampCap) env.lanes([&](int lane, int li) { (void)li; if (lane == 0) ctr[2] = 1; });
    if (env.tid == 0) { part->db_max = dbmax; part->db_min = dbmin; part->mf_frames = nfr; }
    env.csync();                                          // #1: every rank's dB extrema (and this CTA's overflow flag) are visible
    for (int rr = 0; rr < NR; ++rr) {
      const Partials* rp = env.remote(part, rr);
      gmax = fmaxf(gmax, rp->db_max);
      gmin = fminf(gmin, rp->db_min);
    }
    // amplitude_to_DB(top_db = 80): clamp to (segment max - 80).  The DCT is linear, so a clamped value only
    // adds dct[m][k] * (thr - dB) to its frame: every warp patches the frames it produced from its own
    // candidate list (a frame belongs to exactly one warp; list order is program order: deterministic).
    thr = gmax - 80.0f;
    fix = (P.parts & kPartMfcc) && (dbmin < thr);
    // the two ways of applying the clamp give the same bits, so each WARP chooses: patch from its list, or, if the
    // list overflowed, redo its own quads with the deltas computed inline (pass 1 runs if any warp needs it)
    slow = fix && ctr[2] != 0;
    slow_w = fix && ncand > kClampCap;
    if (!slow) break;
  }
  if (fix) {                                               // quads before the overflow (all of them when the list sufficed)
    const int ovf_quad = ovf_task;                         // task index = this rank's local quad index = list key >> 9
    // clamp_fix: per quad with clamped candidates, scatter delta * dct into the [frame x coefficient][lane] tile
    // (entries of one (frame, lane) arrive in slot order), sum the tile over the lanes in lane order and add the
    // result to the quad's rows: bit for bit what the clamped pass computes
    float* fb = reinterpret_cast<float*>(wbuf);
    int e0 = 0;
    const int n_list = (ncand < kClampCap) ? ncand : kClampCap;
    while (e0 < n_list) {                                  // the list and its length are warp-uniform
      const int quad = (clist[e0].x >> 7) >> 2;
      int e1 = e0;
      bool any = false;
      while (e1 < n_list && ((clist[e1].x >> 7) >> 2) == quad) { any = any || (int_as_float(clist[e1].y) < thr); ++e1; }
      if (quad >= ovf_quad) break;                           // that quad and the later ones are handled by pass 1
      if (any) {
        env.lanes([&](int lane, int li) {
          (void)li;
          for (int v = 0; v < 52; ++v) fb[v * 33 + lane] = 0.0f;
        });
        env.wsync();
        env.lanes([&](int lane, int li) {
          (void)li;
          if (lane < kMfcc) {
            for (int e = e0; e < e1; ++e) {
              const int2 ce = clist[e];
              const float d = int_as_float(ce.y);
              if (d < thr) {
                const int m = ce.x & 127, j = (ce.x >> 7) & 3, q = (m >> 5) * kMfcc + lane;
                const float w = env.ldf(P.tab->dctq + ((q >> 2) * 32 + (m & 31)) * 4 + (q & 3));
                float* o = fb + (j * kMfcc + lane) * 33 + (m & 31);
                *o = fmaf(thr - d, w, *o);
              }
            }
          }
        });
        env.wsync();
        env.lanes([&](int lane, int li) {
          (void)li;
          for (int half = 0; half < 2; ++half) {
            const int v = lane + 32 * half;
            const int fl = 4 * quad + v / kMfcc;
            if (v < 52 && fl < nfr) {
              float ca = 0.0f;
              for (int j = 0; j < 32; ++j) ca += fb[v * 33 + j];
              mfl[fl * kMfcc + v % kMfcc] += ca;
            }
          }
        });
        env.wsync();
      }
      e0 = e1;
    }
  }
  env.sync();

  // ---------------------------------------------------------------- MFCC moments (timbre z-score, clarity)
  {
    const float cdead = fmaxf(-100.0f, thr);              // every frame of an empty mel filter sits at this level
    double acc[S][16];
    for (int i = 0; i < S; ++i)
      for (int k = 0; k < 16; ++k) acc[i][k] = 0.0;
    env.lanes([&](int lane, int li) {
      for (int fl = env.warp * 32 + lane; fl < nfr; fl += env.nthreads) {
        float* o = P.dbg_mfcc ? P.dbg_mfcc + ((size_t)seg * nFm + mf_begin + fl) * kMfcc : nullptr;
#pragma unroll
        for (int k = 0; k < kMfcc; ++k) {
          const float vf = fmaf(cdead, tb->dct_dead[k], mfl[fl * kMfcc + k]);
          if (o) o[k] = vf;
          const double v = vf;
          acc[li][k] += v;
          acc[li][13] += v * v;
          if (k < 6) acc[li][14] += fabs(v); else acc[li][15] += fabs(v);
        }
      }
    });
    const int ops[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    block_reduce<16>(env, wred, rout, ops, [&](int li, int k) { return acc[li][k]; });
    if (env.tid == 0) {
      for (int k = 0; k < kMfcc; ++k) part->mf_sum[k] = rout[k];
      part->mf_sumsq = rout[13]; part->mf_abs_lo = rout[14]; part->mf_abs_hi = rout[15];
      part->slow_pass = slow ? 1 : 0;
    }
  }
  env.csync();                                            // #2: all partial moments and atoms are visible

  // ---------------------------------------------------------------- rank 0: merge and assemble the row
  if (r == 0) {
    // gather the energy atoms and the partial moments of all ranks (re-using the FFT tiles, idle by now) in ONE
    // round of remote reads: which rank holds which atoms follows from (T, NR) alone, and the Partials are copied
    // word by word, so no read waits for another (a serial walk over 8 ranks costs ~150 DSMEM round trips)
    float* all_atoms = reinterpret_cast<float*>(smem + lay.buf_off);
    Partials* gathered = reinterpret_cast<Partials*>(smem + lay.buf_off + (((T / kAtom + 8) * 4 + 15) & ~15));
    int nA = 0;
    if (P.parts & kPartWave) {
      const int nGr = ceil_div(T, kGroup), gper_a = ceil_div(nGr, NR), full_atoms = T / kAtom;
      nA = (nGr * 8 < full_atoms) ? nGr * 8 : full_atoms;
      env.lanes([&](int lane, int li) {
        (void)li;
        for (int a = env.warp * 32 + lane; a < nA; a += env.nthreads) {
          const int rr = a / (8 * gper_a);
          all_atoms[a] = env.remote(atoms, rr)[a - 8 * gper_a * rr];
        }
      });
    }
    {
      constexpr int kWords = (int)(sizeof(Partials) / 8);
      static_assert(sizeof(Partials) % 8 == 0, "Partials is copied as 8-byte words");
      env.lanes([&](int lane, int li) {
        (void)li;
        for (int i = env.warp * 32 + lane; i < NR * kWords; i += env.nthreads) {
          const int rr = i / kWords, w = i - rr * kWords;
          reinterpret_cast<double*>(gathered + rr)[w] = reinterpret_cast<const double*>(env.remote(part, rr))[w];
        }
      });
    }
    env.sync();
    // rhythm: frame energies e_g = sum of 5 atoms at stride 2 (400 = 5*80, 160 = 2*80)
    const int nG = (T >= kRhythmWin && (P.parts & kPartWave)) ? (T - kRhythmWin) / kRhythmHop + 1 : 0;
    const int nBk = (P.parts & kPartWave) ? T / kBlock : 0;
    auto frame_e = [&](int g) { const float* a = all_atoms + 2 * g; return (double)(a[0] + a[1] + a[2] + a[3] + a[4]); };
    auto block_ms = [&](int b) {
      float s = 0.0f;
      for (int j = 0; j < 20; ++j) s += all_atoms[20 * b + j];
      return (double)(s * (1.0f / (float)kBlock));
    };
    // the means and centred sums of squares are tiny (498 frames, 50 blocks): every thread of warp 0 .. NW-1
    // accumulates a strided share, two deterministic block reductions finish them
    double gs[S], bs[S];
    env.lanes([&](int lane, int li) {
      double a = 0.0, b = 0.0;
      for (int g = env.warp * 32 + lane; g < nG; g += env.nthreads) a += frame_e(g);
      for (int k = env.warp * 32 + lane; k < nBk; k += env.nthreads) b += block_ms(k);
      gs[li] = a; bs[li] = b;
    });
    const int ops2[2] = {kOpSum, kOpSum};
    block_reduce<2>(env, wred, rout, ops2, [&](int li, int k) { return k == 0 ? gs[li] : bs[li]; });
    const double gmean = (nG > 0) ? rout[0] / nG : 0.0;
    const double bmean = (nBk > 0) ? rout[1] / nBk : 0.0;
    env.sync();
    env.lanes([&](int lane, int li) {
      double a = 0.0, b = 0.0;
      for (int g = env.warp * 32 + lane; g < nG; g += env.nthreads) { const double d = frame_e(g) - gmean; a += d * d; }
      for (int k = env.warp * 32 + lane; k < nBk; k += env.nthreads) { const double d = block_ms(k) - bmean; b += d * d; }
      gs[li] = a; bs[li] = b;
    });
    block_reduce<2>(env, wred, rout, ops2, [&](int li, int k) { return k == 0 ? gs[li] : bs[li]; });
    const double gq = rout[0], bq = rout[1];

    // The row.  The segment-level scalars fall into three independent groups, each a serial chain of fp64 divisions and
    // square roots, so THREE warps form them side by side (warp 0: emotion, "pitch", intensity, speech rate, rhythm;
    // warp 1: timbre and clarity; warp 2: SNR and consistency), one raw feature per lane where a feature has 13 entries;
    // warp 0 then runs the LayerNorm, nan_to_num and the stores, one entry per lane.  Every scalar is computed by the
    // same operations in the same order as a single thread would (the bits do not depend on this split).  (As one
    // warp's chain this block took ~20 us per segment, 8 % of a CTA's life, with the CTA's other warps parked at the
    // last barrier: profiles/r2_v202_features_ncu_full_B1024.txt.)
    float* raw = reinterpret_cast<float*>(wred);            // [0:27] raw features
    double* qd = rout;                                       // [0] snr, [1] clarity, [2] consistency (python floats = doubles),
                                                             // [4:11] diagnostics of the detail record
    const float* emo = P.emo8 ? P.emo8 + (size_t)seg * 8 : nullptr;
    env.sync();                                              // wred / rout are free (the last block reduction has been read)
    if (env.warp == 0) {
      env.lanes([&](int lane, int li) {
        (void)li;
        const double NaN = nan("");
        double p_n = part->p_n, p_sum = part->p_sum, p_sumsq = part->p_sumsq, e_total = part->e_total;
        float p_max = part->p_max;
#pragma unroll 1
        for (int rr = 1; rr < NR; ++rr) {
          const Partials* rp = gathered + rr;
          p_n += rp->p_n; p_sum += rp->p_sum; p_sumsq += rp->p_sumsq; e_total += rp->e_total;
          p_max = fmaxf(p_max, rp->p_max);
        }
        float mine = 0.0f;                                    // raw[lane]
        if (lane < 8) mine = emo ? emo[lane] : 0.125f;
        // pitch: mean of the z-scored residual. mu and sigma are fp32 tensors in the reference, so the
        // value is the rounding residue of mu; the residual itself is fp16 rounding noise of the tensor-core round trip.
        double p_mean = 0.0, p_std = 0.0;
        if (p_n > 1.0) {
          p_mean = p_sum / p_n;
          double var = (p_sumsq - p_sum * p_mean) / (p_n - 1.0);
          p_std = sqrt(var > 0.0 ? var : 0.0);
          const float mu32 = (float)p_mean, sd32 = (float)p_std;
          if (lane == 8) mine = (float)((p_sum - p_n * (double)mu32) / (p_n * ((double)sd32 + 1e-6)));
        }
        // intensity: (e - mean(e)) / (std(e) + 1e-6) over ONE channel: std of one element is NaN
        if (lane == 9) mine = (P.flags & kFlagStrictNan) ? (float)NaN : 0.0f;
        // speech rate (mono): energy > 0.1 * energy in fp32
        if (lane == 23) {
          const float e = (float)e_total;
          mine = (e > e * 0.1f) ? 1.0f : 0.0f;
        }
        // rhythm
        if (nG > 0) {
          if (lane == 24) mine = (float)gmean;
          if (lane == 25) mine = (nG > 1) ? (float)sqrt(gq / (nG - 1)) : (float)NaN;
          if (lane == 26) mine = (float)((double)nG / (double)kSampleRate);
        }
        if (lane < 10 || (lane >= 23 && lane < 27)) raw[lane] = mine;
        if (lane == 0) { qd[4] = p_mean; qd[5] = p_std; qd[6] = (double)p_max; qd[7] = e_total; qd[10] = p_n; }
      });
    }
    // groups 1 and 2 run on warps 1 and 2 when the CTA has them, else on warp 0 after its own group
    const int w1 = (NW > 1) ? 1 : 0, w2 = (NW > 2) ? 2 : 0;
    if (env.warp == w1) {
      env.lanes([&](int lane, int li) {
        (void)li;
        const double NaN = nan("");
        double mf_sumsq = part->mf_sumsq, mf_abs_lo = part->mf_abs_lo, mf_abs_hi = part->mf_abs_hi;
        int mf_frames = part->mf_frames;
#pragma unroll 1
        for (int rr = 1; rr < NR; ++rr) {
          const Partials* rp = gathered + rr;
          mf_sumsq += rp->mf_sumsq; mf_abs_lo += rp->mf_abs_lo; mf_abs_hi += rp->mf_abs_hi;
          mf_frames += rp->mf_frames;
        }
        double ssum = 0.0, mf_mine = 0.0;                     // sum of all coefficient sums; the lane's own coefficient
#pragma unroll 1
        for (int k = 0; k < kMfcc; ++k) {
          double t = part->mf_sum[k];
#pragma unroll 1
          for (int rr = 1; rr < NR; ++rr) t += gathered[rr].mf_sum[k];
          ssum += t;
          if (k == lane - 10) mf_mine = t;
        }
        float mine = 0.0f;
        double clarity = 0.0;
        if (mf_frames > 0) {
          const double n = (double)mf_frames * kMfcc;
          const double mu = ssum / n;
          const double var = (n > 1.0) ? (mf_sumsq - ssum * mu) / (n - 1.0) : NaN;
          const double sd = sqrt(var > 0.0 ? var : (var == var ? 0.0 : NaN));
          if (lane >= 10 && lane < 10 + kMfcc) mine = (float)((mf_mine / mf_frames - mu) / (sd + 1e-6));
          const double hi_m = mf_abs_hi / (7.0 * mf_frames), lo_m = mf_abs_lo / (6.0 * mf_frames);
          clarity = py_clip01((double)((float)hi_m / ((float)lo_m + 1e-6f)));
        }
        if (!(P.parts & kPartMfcc)) clarity = 0.0;
        if (lane >= 10 && lane < 10 + kMfcc) raw[lane] = mine;
        if (lane == 0) { qd[1] = clarity; qd[9] = (double)mf_frames; }
      });
    }
    if (env.warp == w2) {
      env.lanes([&](int lane, int li) {
        (void)li;
        const double NaN = nan("");
        double e_total = part->e_total, e_noise = part->e_noise;
#pragma unroll 1
        for (int rr = 1; rr < NR; ++rr) { e_total += gathered[rr].e_total; e_noise += gathered[rr].e_noise; }
        // quality scalars (python floats in the reference: double arithmetic on fp32 .item() values)
        double snr = 0.0, consistency = 0.0;
        if (P.noise_n > 0 && (P.parts & kPartWave)) {
          const float noise_p = (float)(e_noise / (2.0 * P.noise_n));
          const float sig_p = (float)(e_total / (double)T);
          const float snr_db = 10.0f * log10f(sig_p / (noise_p + 1e-6f));
          snr = py_clip01((double)snr_db / 30.0);
        }
        if (nBk > 0) {
          const float sd = (nBk > 1) ? (float)sqrt(bq / (nBk - 1)) : (float)NaN;
          const double cv = (double)(sd / ((float)bmean + 1e-6f));
          consistency = 1.0 - ((1.0 < cv) ? 1.0 : cv);   // python min(cv, 1.0): NaN stays NaN
        }
        if (lane == 0) { qd[0] = snr; qd[2] = consistency; qd[8] = e_noise; }
      });
    }
    env.sync();
    if (env.warp == 0) {
      env.lanes([&](int lane, int li) {
        (void)li;
        const double quality = 0.4 * qd[0] + 0.3 * qd[1] + 0.3 * qd[2];
        const float q4[4] = {(float)quality, (float)qd[0], (float)qd[1], (float)qd[2]};
        // AudioFeatureNormalizer: pad 27 -> 31 with zeros, LayerNorm(31) (gamma 1, beta 0, eps 1e-5, biased var)
        double m = 0.0;
#pragma unroll 1
        for (int k = 0; k < 27; ++k) m += (double)raw[k];
        m /= 31.0;
        double v = 0.0;
#pragma unroll 1
        for (int k = 0; k < 31; ++k) { const double d = ((k < 27) ? (double)raw[k] : 0.0) - m; v += d * d; }
        v /= 31.0;
        const double rs = 1.0 / sqrt(v + 1e-5);
        const float mine = (lane < 27) ? raw[lane] : 0.0f;
        const float lnv = (float)(((double)mine - m) * rs);
        // fusion input row: LN slices ++ quality, torch.nan_to_num(nan=0.0) (+-inf -> +-FLT_MAX)
        if (lane < 31) {
          float o = (lane < 27) ? lnv : q4[lane - 27];
          if (o != o) o = 0.0f;
          else if (o > 3.4028234663852886e38f) o = 3.4028234663852886e38f;
          else if (o < -3.4028234663852886e38f) o = -3.4028234663852886e38f;
          P.feat31[(size_t)seg * 31 + lane] = o;
        }
        if (P.detail) {
          float* d = P.detail + (size_t)seg * kDetailStride;
          if (lane < 27) d[lane] = mine;
          if (lane < 4) d[27 + lane] = q4[lane];
          if (lane < 31) d[32 + lane] = lnv;
          if (lane == 31) { d[31] = 0.0f; d[63] = 0.0f; }
          if (lane == 0) {
            d[64] = gmax; d[65] = (float)qd[4]; d[66] = (float)qd[5]; d[67] = (float)qd[6];
            d[68] = (float)qd[7]; d[69] = (float)qd[8]; d[70] = (float)qd[9]; d[71] = (float)nG;
            d[72] = (float)qd[10]; d[73] = (float)nBk; d[74] = (float)nA;
            d[75] = slow ? 1.0f : 0.0f; d[76] = gmin; d[77] = cand; d[78] = fix ? 1.0f : 0.0f;
            d[79] = 0.0f;
          }
          if (lane < kDetailStride - 80) d[80 + lane] = 0.0f;
        }
      });
    }
  }
  env.csync();                                            // #3: nobody exits while rank 0 still reads its smem
}

}  // namespace msa
