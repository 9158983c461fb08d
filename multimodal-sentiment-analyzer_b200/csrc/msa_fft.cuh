// In-place mixed-radix FFTs executed by ONE WARP per transform, data in shared memory.
//
// Forward = decimation in frequency (natural order in, digit-reversed out); inverse =
// the exact stage-by-stage inverse (digit-reversed in, natural out, unnormalised).  The
// STFT -> ISTFT round trip of the "pitch" feature therefore needs no permutation at all,
// and the MFCC power spectrum reads its bins through a 400-entry position table.
//
// Compiled both by nvcc (device) and by g++ (tests/emu, CPU emulation of a warp:
// the lane loop runs sequentially, which is valid because within one stage every
// butterfly reads and writes only its own R positions).
#pragma once
#include "msa_hd.h"

namespace msa {

struct alignas(8) c32 { float x, y; };   // 8-byte aligned: one LDS.64 / STS.64 per complex value
MSA_FN c32 operator+(c32 a, c32 b) { return {a.x + b.x, a.y + b.y}; }
MSA_FN c32 operator-(c32 a, c32 b) { return {a.x - b.x, a.y - b.y}; }
MSA_FN c32 cmul(c32 a, c32 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
MSA_FN c32 cmulc(c32 a, c32 b) { return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a * conj(b)
template <bool INV> MSA_FN c32 rot90(c32 a) { return INV ? c32{-a.y, a.x} : c32{a.y, -a.x}; }  // * W4 (fwd: -i)

template <bool INV> MSA_FN void dft2(c32& a, c32& b) { c32 t = a - b; a = a + b; b = t; }

template <bool INV> MSA_FN void dft4(c32& a, c32& b, c32& c, c32& d) {
  c32 t0 = a + c, t1 = a - c, t2 = b + d, t3 = rot90<INV>(b - d);
  a = t0 + t2; b = t1 + t3; c = t0 - t2; d = t1 - t3;
}

// v[k] <- sum_q v[q] W8^{qk}
template <bool INV> MSA_FN void dft8(c32* v) {
  const float h = 0.70710678118654752440f;
  dft4<INV>(v[0], v[2], v[4], v[6]);   // E0..E3 in v0,v2,v4,v6
  dft4<INV>(v[1], v[3], v[5], v[7]);   // O0..O3 in v1,v3,v5,v7
  c32 o1 = v[3], o3 = v[7];
  c32 w1, w3;
  if (INV) { w1 = {(o1.x - o1.y) * h, (o1.x + o1.y) * h}; w3 = {(-o3.x - o3.y) * h, (o3.x - o3.y) * h}; }
  else     { w1 = {(o1.x + o1.y) * h, (o1.y - o1.x) * h}; w3 = {(o3.y - o3.x) * h, -(o3.x + o3.y) * h}; }
  c32 w0 = v[1], w2 = rot90<INV>(v[5]);
  c32 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  v[0] = e0 + w0; v[4] = e0 - w0;
  v[1] = e1 + w1; v[5] = e1 - w1;
  v[2] = e2 + w2; v[6] = e2 - w2;
  v[3] = e3 + w3; v[7] = e3 - w3;
}

template <bool INV> MSA_FN void dft16(c32* v) {
  c32 e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
  dft8<INV>(e);
  dft8<INV>(o);
  // W16^k = (cos(pi k/8), -/+ sin(pi k/8))
  const float c[8] = {1.0f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f,
                      0.0f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f};
  const float s[8] = {0.0f, 0.38268343236508977173f, 0.70710678118654752440f, 0.92387953251128675613f,
                      1.0f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    c32 w = {c[k], INV ? s[k] : -s[k]};
    c32 t = cmul(o[k], w);
    v[k] = e[k] + t;
    v[k + 8] = e[k] - t;
  }
}

template <bool INV> MSA_FN void dft5(c32* v) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
  c32 a1 = v[1] + v[4], a2 = v[2] + v[3], d1 = v[1] - v[4], d2 = v[2] - v[3];
  c32 y0 = {v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y};
  c32 a = {v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y};
  c32 b = {v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y};
  c32 e = {s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y};
  c32 g = {s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y};
  // forward: y1 = a - i e, y4 = a + i e, y2 = b - i g, y3 = b + i g   (inverse: conjugate signs)
  c32 ie = INV ? c32{-e.y, e.x} : c32{e.y, -e.x};     // (-/+ i) * e  -> fwd: -i e
  c32 ig = INV ? c32{-g.y, g.x} : c32{g.y, -g.x};
  v[0] = y0;
  v[1] = a + ie; v[4] = a - ie;
  v[2] = b + ig; v[3] = b - ig;
}

template <int R, bool INV> MSA_FN void dftR(c32* v) {
  if (R == 2) dft2<INV>(v[0], v[1]);
  else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
  else if (R == 5) dft5<INV>(v);
  else if (R == 8) dft8<INV>(v);
  else dft16<INV>(v);
}

struct PadNone { MSA_FN static int at(int i) { return i; } };
struct Pad8 { MSA_FN static int at(int i) { return i + (i >> 3); } };   // 512-pt: kills the stride-8 bank conflicts
constexpr int kPad512 = 512 + 64;

// One radix-R stage over an N-point transform whose current sub-transform size is NS.
// tw = per-stage twiddle table W_NS^(j*k) stored [k-1][j] as (cos, -sin); unused when m == 1.
// LANES is the number of lanes sharing the N/R butterflies (32 on the GPU -> fully unrolled, 1 in the
// CPU emulation); every lane's butterflies are independent, so the compiler can overlap their loads.
template <int N, int R, int NS, bool INV, class P, int LANES>
MSA_FN void fft_stage(c32* zb, const c32* tw, int lane) {
  constexpr int m = NS / R;
  constexpr int items = N / R;
  constexpr int per_lane = (items + LANES - 1) / LANES;
#pragma unroll
  for (int u = 0; u < per_lane; ++u) {
    const int it = lane + u * LANES;
    if (items % LANES == 0 || it < items) {
      const int b = it / m, j = it - b * m;
      const int base = b * NS + j;
      c32 v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) v[q] = zb[P::at(base + q * m)];
      if (INV) {
        if (m > 1) {
#pragma unroll
          for (int k = 1; k < R; ++k) {
            v[k] = cmulc(v[k], tw[(k - 1) * m + j]);
          }
        }
        dftR<R, true>(v);
      } else {
        dftR<R, false>(v);
        if (m > 1) {
#pragma unroll
          for (int k = 1; k < R; ++k) {
            v[k] = cmul(v[k], tw[(k - 1) * m + j]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) zb[P::at(base + q * m)] = v[q];
    }
  }
}

}  // namespace msa
