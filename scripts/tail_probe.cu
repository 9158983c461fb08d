// How the feature kernel's time steps with the number of segments (waves of 2 CTAs x SMs), and what a remainder costs
// as 1-, 2- or 4-CTA clusters.  dlopen of libmsa_b200.so, CUDA events, inputs rotated through a 2048-segment buffer
// (655 MB, larger than L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tail_probe scripts/tail_probe.cu -ldl
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
typedef int (*feat_f32_t)(const float*, int, int, const float*, float*, float*, float*, int, int, int, void*);
__device__ unsigned hash_u32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void synth(float* wav, float* emo, int B, int T) {
  const size_t n = (size_t)B * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int seg = (int)(i / T), t = (int)(i % T);
    const float f0 = 80.0f + 220.0f * (hash_u32(seg * 2 + 1) * (1.0f / 4294967296.0f));
    const float ts = t / 16000.0f;
    float v = 0.0f;
    for (int k = 1; k <= 5; ++k) v += (0.3f / k) * __sinf(6.2831853f * f0 * k * ts);
    v *= 0.5f + 0.5f * __sinf(6.2831853f * 3.0f * ts);
    const unsigned h = hash_u32((unsigned)i * 2654435761u + 12345u);
    v += 0.02f * ((float)(h & 0xffff) + (float)(h >> 16) - 65535.0f) * (1.0f / 26754.0f);
    wav[i] = fminf(fmaxf(v, -1.0f), 1.0f);
    if (t < 8) emo[seg * 8 + t] = 0.125f;
  }
}
int main(int argc, char** argv) {
  void* h = dlopen(argc > 1 ? argv[1] : "multimodal-sentiment-analyzer_b200/libmsa_b200.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("{\"error\": \"%s\"}\n", dlerror()); return 2; }
  feat_f32_t feat = (feat_f32_t)dlsym(h, "msa_features_f32");
  const int T = 80000, NB = 2048;
  float *wav, *emo, *out;
  CK(cudaMalloc(&wav, (size_t)NB * T * 4)); CK(cudaMalloc(&emo, NB * 8 * 4)); CK(cudaMalloc(&out, NB * 31 * 4));
  synth<<<1184, 256>>>(wav, emo, NB, T);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct Case { int B, c; };
  const Case cases[] = {{148, 1}, {296, 1}, {444, 1}, {592, 1}, {740, 1}, {888, 1}, {960, 1}, {1024, 1}, {1184, 1}, {1480, 1},
                        {136, 1}, {136, 2}, {136, 4}, {68, 2}, {68, 4}, {74, 4}, {148, 2},
                        {1, 4}, {1, 8}, {1, 16}, {4, 8}, {4, 16}, {8, 8}, {8, 16}, {16, 8}, {16, 16}, {1, 0}, {8, 0}, {16, 0}};
  printf("{\"cases\": [");
  bool first = true;
  for (const Case& cs : cases) {
    std::vector<float> ms;
    int start = 0;
    for (int it = 0; it < 13; ++it) {
      start = (start + cs.B) % (NB - cs.B + 1);
      CK(cudaEventRecord(e0));
      int rc = feat(wav + (size_t)start * T, cs.B, T, emo, out, nullptr, nullptr, 0, 7, cs.c, nullptr);
      if (rc) { printf("{\"error\": \"rc %d\"}\n", rc); return 2; }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float t; CK(cudaEventElapsedTime(&t, e0, e1));
      if (it >= 3) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    printf("%s{\"B\": %d, \"cluster\": %d, \"ms_median\": %.4f, \"ms_min\": %.4f}", first ? "" : ", ", cs.B, cs.c, ms[ms.size() / 2], ms[0]);
    first = false;
  }
  printf("]}\n");
  return 0;
}
