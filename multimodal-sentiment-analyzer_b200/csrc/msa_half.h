// Packed fp16 pairs in 32-bit registers (the operand format of the tensor-core STFT-512 round trip,
// msa_pitch_tc.cuh): lane-local arithmetic on them, for the device (HFMA2 / HMUL2 / F2FP) and for the
// CPU emulation build (tests/emu, _Float16).  A pair is {lo = first element, hi = second element},
// exactly how mma.sync / ldmatrix / movmatrix lay two consecutive matrix elements out in a register.
#pragma once
#include <cstdint>
#include <cstring>
#include "msa_hd.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace msa {

using u32 = uint32_t;

#if defined(__CUDA_ARCH__)
MSA_FN __half2 h2_as(u32 v) { return *reinterpret_cast<__half2*>(&v); }
MSA_FN u32 h2_bits(__half2 v) { return *reinterpret_cast<u32*>(&v); }
MSA_FN u32 h2_mul(u32 a, u32 b) { return h2_bits(__hmul2(h2_as(a), h2_as(b))); }
MSA_FN u32 h2_add(u32 a, u32 b) { return h2_bits(__hadd2(h2_as(a), h2_as(b))); }
MSA_FN u32 h2_fma(u32 a, u32 b, u32 c) { return h2_bits(__hfma2(h2_as(a), h2_as(b), h2_as(c))); }   // a b + c
MSA_FN u32 h2_fms(u32 a, u32 b, u32 c) { return h2_bits(__hfma2(h2_as(a), h2_as(b), __hneg2(h2_as(c)))); }    // a b - c
MSA_FN u32 h2_fnma(u32 a, u32 b, u32 c) { return h2_bits(__hfma2(__hneg2(h2_as(a)), h2_as(b), h2_as(c))); }   // c - a b
MSA_FN u32 h2_max(u32 a, u32 b) { return h2_bits(__hmax2(h2_as(a), h2_as(b))); }
// two fp32 -> one pair, round to nearest, +-inf saturated to +-65504 (a sample can never poison a frame)
MSA_FN u32 h2_pack(float lo, float hi) {
  u32 r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
MSA_FN float h2_lo(u32 v) { return __low2float(h2_as(v)); }
MSA_FN float h2_hi(u32 v) { return __high2float(h2_as(v)); }
#else
using f16 = _Float16;
MSA_FN f16 h_of(uint16_t b) { f16 h; std::memcpy(&h, &b, 2); return h; }
MSA_FN uint16_t h_bits(f16 h) { uint16_t b; std::memcpy(&b, &h, 2); return b; }
MSA_FN uint16_t h_from_f32(float f) {
  if (f > 65504.0f) f = 65504.0f;
  if (f < -65504.0f) f = -65504.0f;
  return h_bits((f16)f);
}
template <class F> MSA_FN u32 h2_map(u32 a, u32 b, u32 c, F f) {
  const uint16_t lo = h_bits(f(h_of((uint16_t)(a & 0xffffu)), h_of((uint16_t)(b & 0xffffu)), h_of((uint16_t)(c & 0xffffu))));
  const uint16_t hi = h_bits(f(h_of((uint16_t)(a >> 16)), h_of((uint16_t)(b >> 16)), h_of((uint16_t)(c >> 16))));
  return (u32)lo | ((u32)hi << 16);
}
MSA_FN u32 h2_mul(u32 a, u32 b) { return h2_map(a, b, 0, [](f16 x, f16 y, f16) { return (f16)(x * y); }); }
MSA_FN u32 h2_add(u32 a, u32 b) { return h2_map(a, b, 0, [](f16 x, f16 y, f16) { return (f16)(x + y); }); }
MSA_FN u32 h2_fma(u32 a, u32 b, u32 c) {
  return h2_map(a, b, c, [](f16 x, f16 y, f16 z) { return (f16)((float)x * (float)y + (float)z); });
}
MSA_FN u32 h2_fms(u32 a, u32 b, u32 c) {
  return h2_map(a, b, c, [](f16 x, f16 y, f16 z) { return (f16)((float)x * (float)y - (float)z); });
}
MSA_FN u32 h2_fnma(u32 a, u32 b, u32 c) {
  return h2_map(a, b, c, [](f16 x, f16 y, f16 z) { return (f16)((float)z - (float)x * (float)y); });
}
MSA_FN u32 h2_max(u32 a, u32 b) { return h2_map(a, b, 0, [](f16 x, f16 y, f16) { return x > y ? x : y; }); }
MSA_FN u32 h2_pack(float lo, float hi) { return (u32)h_from_f32(lo) | ((u32)h_from_f32(hi) << 16); }
MSA_FN float h2_lo(u32 v) { return (float)h_of((uint16_t)(v & 0xffffu)); }
MSA_FN float h2_hi(u32 v) { return (float)h_of((uint16_t)(v >> 16)); }
#endif
// {lo(a), lo(b)} and {hi(a), hi(b)}: regroup two words that each hold one value's (hi, lo) fp16 parts
#if defined(__CUDA_ARCH__)
MSA_FN u32 h2_lows(u32 a, u32 b) { return __byte_perm(a, b, 0x5410); }
MSA_FN u32 h2_highs(u32 a, u32 b) { return __byte_perm(a, b, 0x7632); }
#else
MSA_FN u32 h2_lows(u32 a, u32 b) { return (a & 0xffffu) | (b << 16); }
MSA_FN u32 h2_highs(u32 a, u32 b) { return (a >> 16) | (b & 0xffff0000u); }
#endif
MSA_FN u32 h2_neg(u32 a) { return a ^ 0x80008000u; }
MSA_FN u32 h2_abs(u32 a) { return a & 0x7fff7fffu; }

}  // namespace msa
