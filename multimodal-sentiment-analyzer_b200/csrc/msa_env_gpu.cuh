// Device execution environment for msa_features_body.cuh: 32-lane warps, block barriers,
// thread-block cluster barriers and distributed shared memory (DSMEM), and the slice loader
// (TMA bulk copy global -> shared when the fp32 source is 16-byte aligned).
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include "msa_features_body.cuh"

namespace msa {
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct GpuEnv {
  static constexpr int kLanes = 32;
  int tid, nthreads, lane, nlanes, warp, nwarps, rank, nranks, cluster_id;

  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void wsync() { __syncwarp(); }
  __device__ __forceinline__ void csync() { cg::this_cluster().sync(); }
  template <class T> __device__ __forceinline__ T* remote(T* p, int r) {
    return cg::this_cluster().map_shared_rank(p, r);
  }
  __device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ double bsum(double v, double* red) {
    v = wsum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += red[w];
    __syncthreads();
    return s;
  }
  __device__ __forceinline__ float bmax(float v, double* red) {
    float* r = reinterpret_cast<float*>(red);
    v = wmax(v);
    if (lane == 0) r[warp] = v;
    __syncthreads();
    float s = r[0];
    for (int w = 1; w < nwarps; ++w) s = fmaxf(s, r[w]);
    __syncthreads();
    return s;
  }

  // Stage n samples of the segment slice into shared memory as fp32.
  template <class InT>
  __device__ __forceinline__ void load_slice(float* dst, const InT* src, int n, void* bar_mem, bool bulk);
};

template <>
__device__ __forceinline__ void GpuEnv::load_slice<float>(float* dst, const float* src, int n, void* bar_mem, bool bulk) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  int done = 0;
  if (bulk && aligned && n >= 4) {
    // TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP), completion counted in bytes on an mbarrier
    const uint32_t bar = smem_u32(bar_mem);
    const int nb = (n & ~3) * 4;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb) : "memory");
      const int chunk = 16384;
      for (int off = 0; off < nb; off += chunk) {
        const int sz = (nb - off < chunk) ? nb - off : chunk;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(reinterpret_cast<unsigned char*>(dst) + off)),
            "l"(reinterpret_cast<const unsigned char*>(src) + off), "r"(sz), "r"(bar)
            : "memory");
      }
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(0)
          : "memory");
    }
    done = n & ~3;
  } else if (aligned) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < n / 4; i += nthreads) d4[i] = __ldg(s4 + i);
    done = n & ~3;
  }
  for (int i = done + tid; i < n; i += nthreads) dst[i] = __ldg(src + i);
}

template <>
__device__ __forceinline__ void GpuEnv::load_slice<int16_t>(float* dst, const int16_t* src, int n, void*, bool) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  int done = 0;
  if (aligned) {
    const int4* s8 = reinterpret_cast<const int4*>(src);
    const float k = 1.0f / 32768.0f;
    for (int i = tid; i < n / 8; i += nthreads) {
      const int4 v = __ldg(s8 + i);
      float4 a, b;
      a.x = (float)(short)(v.x & 0xffff) * k; a.y = (float)(short)(v.x >> 16) * k;
      a.z = (float)(short)(v.y & 0xffff) * k; a.w = (float)(short)(v.y >> 16) * k;
      b.x = (float)(short)(v.z & 0xffff) * k; b.y = (float)(short)(v.z >> 16) * k;
      b.z = (float)(short)(v.w & 0xffff) * k; b.w = (float)(short)(v.w >> 16) * k;
      reinterpret_cast<float4*>(dst)[2 * i] = a;
      reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
    }
    done = n & ~7;
  }
  for (int i = done + tid; i < n; i += nthreads) dst[i] = (float)__ldg(src + i) * (1.0f / 32768.0f);
}

}  // namespace msa
