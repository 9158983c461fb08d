// Device execution environment for msa_features_body.cuh: 32-lane warps (per-lane state lives in
// registers, so there is ONE state copy and lanes() runs its body once), block barriers,
// thread-block cluster barriers and distributed shared memory (DSMEM), read-only global loads.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include "msa_half.h"
#include "msa_features_body.cuh"

namespace msa {
namespace cg = cooperative_groups;

struct GpuEnv {
  static constexpr int kStates = 1;
  int tid, nthreads, lane, warp, nwarps, rank, nranks, cluster_id;

  template <class F> __device__ __forceinline__ void lanes(F&& f) { f(lane, 0); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void wsync() { __syncwarp(); }
  // cluster barrier; a one-CTA "cluster" only needs the block barrier (the hardware cluster barrier costs ~400 cycles
  // and flushes L1 even then)
  __device__ __forceinline__ void csync() {
    if (nranks > 1) cg::this_cluster().sync();
    else __syncthreads();
  }
  template <class T> __device__ __forceinline__ T* remote(T* p, int r) {
    return cg::this_cluster().map_shared_rank(p, r);
  }
  __device__ __forceinline__ float log2(float v) { return __log2f(v); }

  // one sample as fp32 (int16 PCM is scaled by 1/32768 like torchaudio.load)
  __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  // int16 -> fp32 on the integer/FMA pipes instead of the quarter-rate I2F: 1.5 * 2^23 + v is exact in fp32
  static __device__ __forceinline__ float s16_to_f32(int v) {
    // (12582912 + v) * 2^-15 - 384 in one FMA: every intermediate is exactly representable
    return __fmaf_rn(__int_as_float(0x4B400000 + v), 1.0f / 32768.0f, -384.0f);
  }
  __device__ __forceinline__ float ld(const int16_t* p) { return s16_to_f32((int)__ldg(p)); }
  // the same for the LAST pass over a segment (the MFCC frames): evict-first, so that lines nobody will read again
  // leave L2 before the live segments of the other 295 CTAs do
  __device__ __forceinline__ float ld_last(const float* p) { return __ldcs(p); }
  __device__ __forceinline__ float ld_last(const int16_t* p) { return s16_to_f32((int)__ldcs(p)); }

  // request the 128-byte line at p into L1 (no register, no wait)
  __device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
  // request the 128-byte line at p into L2 (no register, no wait)
  __device__ __forceinline__ void prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

  // samples idx .. idx+3 of a segment of T samples (0 beyond the end), one vector load when aligned
  __device__ __forceinline__ void ld4(const float* x, int idx, int T, float* v) {
    const float* p = x + idx;
    if (idx + 3 < T && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (idx + i < T) ? __ldg(p + i) : 0.0f;
    }
  }
  __device__ __forceinline__ void ld4(const int16_t* x, int idx, int T, float* v) {
    const int16_t* p = x + idx;
    if (idx + 3 < T && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
      const int2 q = __ldg(reinterpret_cast<const int2*>(p));
      v[0] = s16_to_f32((int)(short)(q.x & 0xffff)); v[1] = s16_to_f32(q.x >> 16);
      v[2] = s16_to_f32((int)(short)(q.y & 0xffff)); v[3] = s16_to_f32(q.y >> 16);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (idx + i < T) ? s16_to_f32((int)__ldg(p + i)) : 0.0f;
    }
  }

  // one 16-byte load: 4 fp32 samples or 8 int16 samples starting at idx (0 beyond the end of the segment)
  __device__ __forceinline__ void ldv(const float* x, int idx, int T, float* v) { ld4(x, idx, T, v); }
  __device__ __forceinline__ void ldv(const int16_t* x, int idx, int T, float* v) {
    const int16_t* p = x + idx;
    if (idx + 7 < T && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(p));
      const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = s16_to_f32((int)(short)(w[i] & 0xffff));
        v[2 * i + 1] = s16_to_f32(w[i] >> 16);
      }
    } else {
      ld4(x, idx, T, v);
      ld4(x, idx + 4, T, v + 4);
    }
  }

  // ---- tensor-core STFT-512 round trip (msa_pitch_tc.cuh): fp16 pairs in 32-bit registers
  __device__ __forceinline__ u32 ldu(const u32* p) { return __ldg(p); }
  __device__ __forceinline__ float cospi(float v) { return cospif(v); }
  // 16 consecutive samples idx .. idx + 15, all inside the segment
  __device__ __forceinline__ void ld16(const float* x, int idx, float* v) {
    const float* p = x + idx;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldg(p + i);
    }
  }
  __device__ __forceinline__ void ld16(const int16_t* x, int idx, float* v) {
    const int16_t* p = x + idx;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(p) + i);
        const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[8 * i + 2 * k] = s16_to_f32((int)(short)(w[k] & 0xffff));
          v[8 * i + 2 * k + 1] = s16_to_f32(w[k] >> 16);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = s16_to_f32((int)__ldg(p + i));
    }
  }
  // 8 fp16 values (4 pairs) to 16-byte aligned shared memory
  __device__ __forceinline__ void st8(uint16_t* dst, const u32* w) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ __forceinline__ u32 lds1(const u32* p) { return *p; }
  __device__ __forceinline__ void lds2(u32* d, const u32* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    d[0] = v.x; d[1] = v.y;
  }
  __device__ __forceinline__ void lds4(u32* d, const u32* p) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  // D (16 x 8, fp16) = A (16 x 16) B (16 x 8) [+ D]: the getters return the lane's registers
  template <class D, class A, class B> __device__ __forceinline__ void mma(D dg, A ag, B bg, bool accumulate) {
    u32* d = dg(0);
    const u32* a = ag(0);
    const u32* b = bg(0);
    const u32 c0 = accumulate ? d[0] : 0u, c1 = accumulate ? d[1] : 0u;
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%9};"
                 : "=r"(d[0]), "=r"(d[1])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(c0), "r"(c1));
  }
  // D (16 x 8, fp32) += A (16 x 16, fp16) B (16 x 8, fp16)
  template <class D, class A, class B> __device__ __forceinline__ void mma_f32(D dg, A ag, B bg) {
    float* d = dg(0);
    const u32* a = ag(0);
    const u32* b = bg(0);
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  __device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
  // four transposed 8 x 8 fp16 tiles: lane l supplies the address of row l % 8 of tile l / 8 (8 contiguous values)
  template <class R, class P> __device__ __forceinline__ void ldsm4t(R rg, P pg) {
    u32* r = rg(0);
    const u32 addr = (u32)__cvta_generic_to_shared(pg(0));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
  }
  // transpose one 8 x 8 fp16 tile held across the warp (lane g, t: row g, columns 2 t, 2 t + 1), in place
  template <class D> __device__ __forceinline__ void movmt(D dg) {
    u32* d = dg(0);
    u32 o;
    asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(o) : "r"(d[0]));
    d[0] = o;
  }

  // block-wide copy of a 16-byte aligned table into shared memory
  __device__ __forceinline__ void copy16(void* dst, const void* src, int bytes) {
    const int4* s = reinterpret_cast<const int4*>(src);
    int4* d = reinterpret_cast<int4*>(dst);
    for (int i = tid; i < bytes / 16; i += nthreads) d[i] = __ldg(s + i);
  }

  // op over the 32 lanes' values (butterfly of shuffles: every lane gets the same, deterministic result)
  template <class G> __device__ __forceinline__ double warp_reduce(int op, G&& get) {
    double v = get(0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = red_comb(op, v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
  }

  // append (key, value) of the lanes with `pred` to a per-warp list in lane order; `cnt` is warp-uniform
  // and keeps counting past `cap` (overflow is the caller's signal)
  __device__ __forceinline__ void push(bool pred, int2* list, int& cnt, int cap, int key, float val) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m) {
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (pred && pos < cap) list[pos] = make_int2(key, __float_as_int(val));
      cnt += __popc(m);
    }
  }
};

}  // namespace msa
