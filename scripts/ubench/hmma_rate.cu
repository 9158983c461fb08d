// Issue rate of the legacy tensor path (mma.sync.m16n8k16, SASS HMMA.16816) on sm_100a: the denominator of the "tensor
// floor" of the STFT-512 round trip in DESIGN.md 4.1.  Every warp runs 8 INDEPENDENT accumulator chains so that the pipe,
// not the latency, bounds the loop; 148 x 2 CTAs x 8 warps (the feature kernel's occupancy) and 148 x 1 x 32 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/hmma_rate scripts/ubench/hmma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <bool F32ACC>
__global__ void hmma_loop(uint32_t* out, int iters) {
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x38003800u, 0x3c003800u}, b[2] = {0x3c003c00u, 0x34003400u};
  if (F32ACC) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(s);
  } else {
    uint32_t c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0u;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                     : "+r"(c[i][0]), "+r"(c[i][1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= c[i][0] ^ c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  uint32_t* out;
  cudaMalloc(&out, (size_t)sms * 2048 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  printf("{\"what\": \"mma.sync m16n8k16 issue rate\", \"sms\": %d, \"max_clock_mhz\": %d, \"runs\": [", sms, khz / 1000);
  bool first = true;
  for (int acc32 = 0; acc32 < 2; ++acc32)
    for (int shape = 0; shape < 2; ++shape) {
      const int ctas = shape == 0 ? 2 * sms : sms, threads = shape == 0 ? 256 : 1024;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (acc32) hmma_loop<true><<<ctas, threads>>>(out, iters); else hmma_loop<false><<<ctas, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0) continue;
        const double mmas = (double)ctas * (threads / 32) * 8.0 * iters;
        printf("%s{\"accumulate\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"hmma_per_us_per_sm\": %.1f, \"dense_tflops\": %.1f}", first ? "" : ", ",
               acc32 ? "f32" : "f16", ctas * (threads / 32) / sms, ms, mmas / (ms * 1e3) / sms, mmas * 4096.0 / (ms * 1e-3) / 1e12);
        first = false;
      }
    }
  printf("]}\n");
  return cudaGetLastError() != cudaSuccess;
}
