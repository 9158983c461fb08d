// Microbenchmark: issue cost of packed fp32 (fma.rn.f32x2 / add.rn.f32x2) against scalar FFMA / FADD on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint64_t pk(float a, float b) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float c0, float c1) {
  const float s = (float)threadIdx.x * 1e-3f;
  if (MODE == 0) {            // 16 independent scalar FFMA chains
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = s + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], c0, c1);
    }
    float r = 0; 
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 1) {     // 8 independent packed FFMA2 chains (same flops as mode 0)
    uint64_t a[8]; const uint64_t C0 = pk(c0, c0), C1 = pk(c1, c1);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(s + i, s - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], C0, C1);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk(a[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 2) {     // 16 scalar FADD chains
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = s + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = a[i] + a[(i + 1) & 15];
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 3) {     // 8 packed FADD2 chains
    uint64_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(s + i, s - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = add2(a[i], a[(i + 1) & 7]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk(a[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 4) {     // packed FFMA2 whose multiplier is a compile-time constant pair
    uint64_t a[8]; const uint64_t C1 = pk(c1, c1);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(s + i, s - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], pk(0.92387953251128673848f, 0.92387953251128673848f), C1);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk(a[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  }
}

template <int MODE> void run(const char* name, int flops_per_thread_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {              // warps per SM scheduler = blocks*8/4
    const int blocks_per_sm = warps / 2;                       // 256 threads = 8 warps = 2 per scheduler
    k<MODE><<<148 * blocks_per_sm, 256>>>(out, 10, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * blocks_per_sm, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)148 * blocks_per_sm * 256 * iters * flops_per_thread_iter;
    printf("%-28s warps/sched %2d  %.3f ms  %.1f TFLOP/s (or Tadd/s)\n", name, warps, ms, fl / ms / 1e9);
  }
  cudaFree(out);
}

int main() {
  run<0>("scalar FFMA x16", 32);
  run<1>("packed FFMA2 x8", 32);
  run<4>("packed FFMA2 x8 const mult", 32);
  run<2>("scalar FADD x16", 16);
  run<3>("packed FADD2 x8", 16);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
