// Register-resident DFT butterflies (radix 4, 5, 8, 16, 25, 32) on complex fp32 values.
//
// The two transform sizes of the audio path are factored so that every pass runs entirely in
// registers and a transform crosses shared memory exactly once per direction:
//     512 = 16 x 32   (STFT of torchaudio.transforms.PitchShift, audio_analyzer.py:43-47)
//     400 = 16 x 25   (STFT inside torchaudio.transforms.MFCC,   audio_analyzer.py:207-210)
// With n = N2 n1 + n2 and k = k1 + 16 k2 (N2 = 32 or 25):
//     X[k1 + 16 k2] = sum_n2 W_N2^(n2 k2) * W_N^(n2 k1) * ( sum_n1 x[N2 n1 + n2] W_16^(n1 k1) )
// Pass A: lane n2 runs one radix-16 butterfly over the stride-N2 samples it loaded itself and
// applies the inter-pass twiddle; pass B: lane k1 runs one radix-N2 butterfly on one row.
// All butterflies take and return NATURAL order (pure register renaming), so there are no
// digit-reversal tables.  Twiddle factors inside a butterfly are compile-time constants.
//
// Compiled by nvcc (device) and by g++ (tests/emu, CPU emulation of a warp).
#pragma once
#include "msa_hd.h"

namespace msa {

struct alignas(8) c32 { float x, y; };   // 8-byte aligned: one LDS.64 / STS.64 per complex value, one register PAIR

// ---- pair primitives ---------------------------------------------------------------------------
// A complex value is a register pair, and sm_100a has packed fp32 arithmetic on pairs (FADD2 / FMUL2 /
// FFMA2: two independent IEEE operations per instruction, one issue slot instead of two) whose pair
// operand takes a half swap (LO_HI) and a sign per half as free modifiers, and whose other multiplicand
// may be a scalar register or immediate broadcast to both halves.  That is exactly complex arithmetic:
// a +- b, a +- i b, (a, i a) * s + c with a real s.  Every function below is two independent fp32
// operations, written out per half for the CPU build (tests/emu) and as one packed instruction for the
// device (same roundings: results do not depend on which form runs).  The kernel is bound by instruction
// issue, so halving the butterfly instructions is worth ~1/4 of its run time (scripts/ubench/f32x2.cu).
#if defined(__CUDA_ARCH__)
MSA_FN float2 pk_f2(c32 a) { return make_float2(a.x, a.y); }
MSA_FN c32 pk_c(float2 a) { return c32{a.x, a.y}; }
MSA_FN c32 operator+(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), pk_f2(b))); }
MSA_FN c32 operator-(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), make_float2(-b.x, -b.y))); }
MSA_FN c32 add_mi(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), make_float2(b.y, -b.x))); }       // a - i b
MSA_FN c32 add_pi(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), make_float2(-b.y, b.x))); }       // a + i b
MSA_FN c32 add_conj(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), make_float2(b.x, -b.y))); }     // a + conj(b)
MSA_FN c32 sub_conj(c32 a, c32 b) { return pk_c(__fadd2_rn(pk_f2(a), make_float2(-b.x, b.y))); }     // a - conj(b)
MSA_FN c32 mul_s(c32 a, float s) { return pk_c(__fmul2_rn(pk_f2(a), make_float2(s, s))); }            // a s
MSA_FN c32 fma_s(c32 a, float s, c32 c) { return pk_c(__ffma2_rn(pk_f2(a), make_float2(s, s), pk_f2(c))); }               // a s + c
MSA_FN c32 fma_mi_s(c32 a, float s, c32 c) { return pk_c(__ffma2_rn(make_float2(a.y, -a.x), make_float2(s, s), pk_f2(c))); }  // (-i a) s + c
MSA_FN c32 fma_pi_s(c32 a, float s, c32 c) { return pk_c(__ffma2_rn(make_float2(-a.y, a.x), make_float2(s, s), pk_f2(c))); }  // (i a) s + c
MSA_FN c32 fms2(c32 e, c32 lo) { return pk_c(__ffma2_rn(pk_f2(e), make_float2(2.0f, 2.0f), make_float2(-lo.x, -lo.y))); }   // 2 e - lo
#else
MSA_FN c32 operator+(c32 a, c32 b) { return {a.x + b.x, a.y + b.y}; }
MSA_FN c32 operator-(c32 a, c32 b) { return {a.x - b.x, a.y - b.y}; }
MSA_FN c32 add_mi(c32 a, c32 b) { return {a.x + b.y, a.y - b.x}; }
MSA_FN c32 add_pi(c32 a, c32 b) { return {a.x - b.y, a.y + b.x}; }
MSA_FN c32 add_conj(c32 a, c32 b) { return {a.x + b.x, a.y - b.y}; }
MSA_FN c32 sub_conj(c32 a, c32 b) { return {a.x - b.x, a.y + b.y}; }
MSA_FN c32 mul_s(c32 a, float s) { return {a.x * s, a.y * s}; }
MSA_FN c32 fma_s(c32 a, float s, c32 c) { return {fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)}; }
MSA_FN c32 fma_mi_s(c32 a, float s, c32 c) { return {fmaf(a.y, s, c.x), fmaf(-a.x, s, c.y)}; }
MSA_FN c32 fma_pi_s(c32 a, float s, c32 c) { return {fmaf(-a.y, s, c.x), fmaf(a.x, s, c.y)}; }
MSA_FN c32 fms2(c32 e, c32 lo) { return {fmaf(e.x, 2.0f, -lo.x), fmaf(e.y, 2.0f, -lo.y)}; }
#endif
// a b = a b.x + (i a) b.y   and   a conj(b) = a b.x + (-i a) b.y: the twiddle is the broadcast operand
MSA_FN c32 cmul(c32 a, c32 b) { return fma_pi_s(a, b.y, mul_s(a, b.x)); }
MSA_FN c32 cmulc(c32 a, c32 b) { return fma_mi_s(a, b.y, mul_s(a, b.x)); }

// cos(pi * j / 16), j = 0..16
template <int J> MSA_FN constexpr float cospi16() {
  constexpr float t[17] = {1.0f,
                           0.98078528040323043058f,
                           0.92387953251128673848f,
                           0.83146961230254523567f,
                           0.70710678118654757274f,
                           0.55557023301960228867f,
                           0.38268343236508983729f,
                           0.19509032201612833135f,
                           0.0f,
                           -0.19509032201612833135f,
                           -0.38268343236508983729f,
                           -0.55557023301960228867f,
                           -0.70710678118654757274f,
                           -0.83146961230254523567f,
                           -0.92387953251128673848f,
                           -0.98078528040323043058f,
                           -1.0f};
  return t[J];
}
template <int J> MSA_FN constexpr float sinpi16() { return J <= 8 ? cospi16<8 - (J <= 8 ? J : 8)>() : cospi16<(J > 8 ? J : 8) - 8>(); }

// cos / sin (2 pi j / 25), j = 0..24
template <int J> MSA_FN constexpr float cos2pi25() {
  constexpr float t[13] = {1.0f,
                           0.96858316112863107605f,
                           0.87630668004386358394f,
                           0.72896862742141155245f,
                           0.53582679497899654564f,
                           0.30901699437494745126f,
                           0.06279051952931352654f,
                           -0.18738131458572460097f,
                           -0.42577929156507271502f,
                           -0.63742398974868974548f,
                           -0.80901699437494734024f,
                           -0.92977648588825134723f,
                           -0.99211470131447776488f};
  return t[J <= 12 ? J : 25 - J];
}
template <int J> MSA_FN constexpr float sin2pi25() {
  constexpr float t[13] = {0.0f,
                           0.24868988716485479484f,
                           0.48175367410171532345f,
                           0.68454710592868861507f,
                           0.84432792550201507531f,
                           0.95105651629515353118f,
                           0.99802672842827155897f,
                           0.98228725072868872115f,
                           0.90482705246601946580f,
                           0.77051324277578925326f,
                           0.58778525229247324813f,
                           0.36812455268467814129f,
                           0.12533323356430453588f};
  return J <= 12 ? t[J <= 12 ? J : 0] : -t[J > 12 ? 25 - J : 0];
}

// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1
template <int I> struct IC { static constexpr int value = I; };
template <int I, int N, class F> MSA_FN void static_for(F&& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for<I + 1, N>(f);
  }
}

template <bool INV> MSA_FN void dft4(c32& a, c32& b, c32& c, c32& d) {
  const c32 t0 = a + c, t1 = a - c, t2 = b + d, u = b - d;      // forward: * W4 = -i
  a = t0 + t2; c = t0 - t2;
  b = INV ? add_pi(t1, u) : add_mi(t1, u);
  d = INV ? add_mi(t1, u) : add_pi(t1, u);
}

// Radix-2 butterfly with the twiddle folded into the multiply-adds:
//   lo = e + o W_32^K,  hi = e - o W_32^K
// Generic twiddle: lo costs four FMAs, and hi = 2 e - lo two more (6 instead of 8 instructions);
// W_32^4 / W_32^12 (45 degrees): two adds and four FMAs; K = 0 and K = 8 are four adds.
template <int K, bool INV> MSA_FN void bfly(c32 e, c32 o, c32& lo, c32& hi) {
  if constexpr (K == 0) {
    lo = e + o; hi = e - o;
  } else if constexpr (K == 8) {                                  // forward: o W = -i o
    lo = INV ? add_pi(e, o) : add_mi(e, o);
    hi = INV ? add_mi(e, o) : add_pi(e, o);
  } else if constexpr (K == 4 || K == 12) {
    constexpr float h = 0.70710678118654752440f;
    const c32 u = add_mi(o, o);                                    // (o.x + o.y, o.y - o.x) = (1 - i) o
    // o W / h:  K = 4 forward u, inverse i u;  K = 12 forward -i u, inverse -u
    if constexpr (K == 4 && !INV) { lo = fma_s(u, h, e); hi = fma_s(u, -h, e); }
    else if constexpr (K == 4 && INV) { lo = fma_mi_s(u, -h, e); hi = fma_mi_s(u, h, e); }
    else if constexpr (K == 12 && !INV) { lo = fma_mi_s(u, h, e); hi = fma_mi_s(u, -h, e); }
    else { lo = fma_s(u, -h, e); hi = fma_s(u, h, e); }
  } else {
    constexpr float c = cospi16<K>(), sn = sinpi16<K>();       // angle = pi K / 16
    constexpr float s = INV ? -sn : sn;                        // forward w = (c, -sn): o w = o c + (-i o) sn
    lo = fma_mi_s(o, s, fma_s(o, c, e));
    hi = fms2(e, lo);
  }
}

// v[k] <- sum_q v[q] W_N^{qk}, radix-2 decimation in time built on dft4 (N = 4, 8, 16, 32)
template <int N, bool INV> MSA_FN void dft_pow2(c32* v) {
  if constexpr (N == 4) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
  } else {
    c32 e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft_pow2<N / 2, INV>(e);
    dft_pow2<N / 2, INV>(o);
    static_for<0, N / 2>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      bfly<k*(32 / N), INV>(e[k], o[k], v[k], v[k + N / 2]);
    });
  }
}
template <bool INV> MSA_FN void dft16(c32* v) { dft_pow2<16, INV>(v); }
template <bool INV> MSA_FN void dft32(c32* v) { dft_pow2<32, INV>(v); }

template <bool INV> MSA_FN void dft5(c32& v0, c32& v1, c32& v2, c32& v3, c32& v4) {
  constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
  constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
  const c32 a1 = v1 + v4, a2 = v2 + v3, d1 = v1 - v4, d2 = v2 - v3;
  const c32 y0 = (v0 + a1) + a2;
  const c32 a = fma_s(a2, c2, fma_s(a1, c1, v0));
  const c32 b = fma_s(a2, c1, fma_s(a1, c2, v0));
  const c32 e = fma_s(d2, s2, mul_s(d1, s1));
  const c32 g = fma_s(d2, -s1, mul_s(d1, s2));
  // forward: y1 = a - i e, y4 = a + i e, y2 = b - i g, y3 = b + i g   (inverse: conjugate signs)
  v0 = y0;
  v1 = INV ? add_pi(a, e) : add_mi(a, e); v4 = INV ? add_mi(a, e) : add_pi(a, e);
  v2 = INV ? add_pi(b, g) : add_mi(b, g); v3 = INV ? add_mi(b, g) : add_pi(b, g);
}

// 25-point DFT, natural order in and out: n = 5a + b, k = c + 5d
//   X[c + 5d] = sum_b W5^{bd} * ( W25^{bc} * sum_a x[5a + b] W5^{ac} )
template <bool INV> MSA_FN void dft25(c32* v) {
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5<INV>(v[b], v[5 + b], v[10 + b], v[15 + b], v[20 + b]);   // over a -> index c at v[5c + b]
  static_for<1, 5>([&](auto bc) {
    static_for<1, 5>([&](auto cc) {
      constexpr int b = decltype(bc)::value, c = decltype(cc)::value;
      constexpr float wc = cos2pi25<b * c>(), ws = sin2pi25<b * c>();
      const c32 a = v[5 * c + b];                                 // forward: a (wc, -ws) = a wc + (-i a) ws
      v[5 * c + b] = fma_mi_s(a, INV ? -ws : ws, mul_s(a, wc));
    });
  });
#pragma unroll
  for (int c = 0; c < 5; ++c) dft5<INV>(v[5 * c], v[5 * c + 1], v[5 * c + 2], v[5 * c + 3], v[5 * c + 4]);  // over b -> d at v[5c + d]
  // v[5c + d] holds X[c + 5d]: transpose the 5x5 register tile to natural order
#pragma unroll
  for (int c = 0; c < 5; ++c)
#pragma unroll
    for (int d = c + 1; d < 5; ++d) { const c32 t = v[5 * c + d]; v[5 * c + d] = v[5 * d + c]; v[5 * d + c] = t; }
}

}  // namespace msa
