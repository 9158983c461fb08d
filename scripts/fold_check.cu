// Stand-alone check of the feature kernel's FOLD variant (kFlagFoldWave = flag bit 4) against the default build of the
// same kernel, straight through the C ABI (dlopen of libmsa_b200.so; no Python, no torch: starts in a second on a
// fresh box).  Compares the 31-float rows and the raw detail columns bit for bit over several segment lengths and
// cluster sizes, then times both variants on BASELINE configs[1] (1024 x 5 s, fp32 and int16) with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/fold_check scripts/fold_check.cu -ldl
//   ./scripts/fold_check multimodal-sentiment-analyzer_b200/libmsa_b200.so > gpurun_out/fold_check.json
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

typedef int (*feat_f32_t)(const float*, int, int, const float*, float*, float*, float*, int, int, int, void*);
typedef int (*feat_s16_t)(const int16_t*, int, int, const float*, float*, float*, float*, int, int, int, void*);

__device__ unsigned hash_u32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// voiced-speech-like segments: 5 harmonics under a syllabic envelope plus noise, quantised to int16;
// every 7th segment has a stretch of digital silence
__global__ void synth_kernel(int16_t* pcm, float* wav, int B, int T) {
  const size_t n = (size_t)B * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int seg = (int)(i / T), t = (int)(i % T);
    const float f0 = 80.0f + 220.0f * (hash_u32(seg * 2 + 1) * (1.0f / 4294967296.0f));
    const float r = 2.0f + 4.0f * (hash_u32(seg * 2 + 2) * (1.0f / 4294967296.0f));
    const float ts = t / 16000.0f;
    float v = 0.0f;
    for (int k = 1; k <= 5; ++k) v += (0.3f / k) * __sinf(6.2831853f * f0 * k * ts);
    v *= 0.5f + 0.5f * __sinf(6.2831853f * r * ts);
    const unsigned h = hash_u32((unsigned)i * 2654435761u + 12345u);
    v += 0.02f * ((float)(h & 0xffff) + (float)(h >> 16) - 65535.0f) * (1.0f / 26754.0f);
    if (seg % 7 == 3 && t > T / 3 && t < T / 2) v = 0.0f;
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    const int q = __float2int_rn(v * 32767.0f);
    pcm[i] = (int16_t)q;
    wav[i] = (float)q * (1.0f / 32768.0f);
  }
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "multimodal-sentiment-analyzer_b200/libmsa_b200.so";
  void* h = dlopen(path, RTLD_NOW);
  if (!h) { printf("{\"error\": \"dlopen: %s\"}\n", dlerror()); return 2; }
  feat_f32_t f32 = (feat_f32_t)dlsym(h, "msa_features_f32");
  feat_s16_t s16 = (feat_s16_t)dlsym(h, "msa_features_s16");
  if (!f32 || !s16) { printf("{\"error\": \"missing symbols\"}\n"); return 2; }

  const int Bmax = 1024, Tmax = 80640;
  int16_t* pcm; float *wav, *feat[2], *det[2];
  CK(cudaMalloc(&pcm, (size_t)Bmax * Tmax * 2));
  CK(cudaMalloc(&wav, (size_t)Bmax * Tmax * 4));
  for (int v = 0; v < 2; ++v) { CK(cudaMalloc(&feat[v], Bmax * 31 * 4)); CK(cudaMalloc(&det[v], Bmax * 96 * 4)); }
  std::vector<float> hf[2], hd[2];
  for (int v = 0; v < 2; ++v) { hf[v].resize(Bmax * 31); hd[v].resize(Bmax * 96); }

  printf("{\"what\": \"features kernel: FOLD (flag 4) vs default, bitwise\", \"cases\": [");
  struct Case { int B, T, c, is16; };
  const Case cases[] = {{1024, 80000, 0, 0}, {1024, 80000, 0, 1}, {64, 80000, 2, 0}, {16, 80000, 4, 0}, {4, 80000, 8, 0}, {1, 80000, 0, 0},
                        {64, 12345, 1, 0}, {64, 80129, 1, 0}, {64, 80127, 2, 1}, {64, 30001, 1, 0}, {64, 513, 1, 0}, {64, 1700, 1, 1}, {64, 257, 1, 0}};
  int bad_total = 0;
  bool first = true;
  for (const Case& cs : cases) {
    synth_kernel<<<592, 256>>>(pcm, wav, cs.B, cs.T);
    CK(cudaGetLastError());
    int rc[2];
    for (int v = 0; v < 2; ++v) {
      CK(cudaMemset(feat[v], 0xFF, Bmax * 31 * 4));
      CK(cudaMemset(det[v], 0xFF, Bmax * 96 * 4));
      const int flags = 1 | (v ? 4 : 0);
      rc[v] = cs.is16 ? s16(pcm, cs.B, cs.T, nullptr, feat[v], det[v], nullptr, flags, 7, cs.c, nullptr)
                      : f32(wav, cs.B, cs.T, nullptr, feat[v], det[v], nullptr, flags, 7, cs.c, nullptr);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hf[v].data(), feat[v], cs.B * 31 * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hd[v].data(), det[v], cs.B * 96 * 4, cudaMemcpyDeviceToHost));
    }
    int bad_rows = 0, bad_col = -1, folded = 0;
    for (int b = 0; b < cs.B; ++b) {
      bool bad = memcmp(&hf[0][b * 31], &hf[1][b * 31], 31 * 4) != 0;
      for (int c = 0; c < 79 && !bad; ++c)
        if (memcmp(&hd[0][b * 96 + c], &hd[1][b * 96 + c], 4) != 0) { bad = true; bad_col = c; }
      bad_rows += bad;
      folded += hd[1][b * 96 + 79] == 1.0f;
    }
    bad_total += bad_rows + (rc[0] != 0) + (rc[1] != 0);
    printf("%s{\"B\": %d, \"T\": %d, \"cluster\": %d, \"s16\": %d, \"rc\": [%d, %d], \"rows_folded\": %d, \"rows_differing\": %d, \"first_bad_col\": %d, \"quality_row0\": %.7g}",
           first ? "" : ", ", cs.B, cs.T, cs.c, cs.is16, rc[0], rc[1], folded, bad_rows, bad_col, hf[1][27]);
    first = false;
  }
  printf("], \"rows_differing_total\": %d, \"timing_ms_per_1024_segments\": {", bad_total);

  // timing: 1024 x 80000, 3 warm-up + 20 timed launches per variant, CUDA events on the launching (default) stream
  synth_kernel<<<592, 256>>>(pcm, wav, 1024, 80000);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  first = true;
  for (int is16 = 0; is16 < 2; ++is16)
    for (int rep = 0; rep < 2; ++rep)          // default, fold, default, fold: drift shows up as a difference between the repeats
      for (int v = 0; v < 2; ++v) {
        const int flags = 1 | (v ? 4 : 0);
        for (int i = 0; i < 3; ++i) is16 ? s16(pcm, 1024, 80000, nullptr, feat[v], nullptr, nullptr, flags, 7, 0, nullptr)
                                         : f32(wav, 1024, 80000, nullptr, feat[v], nullptr, nullptr, flags, 7, 0, nullptr);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; ++i) is16 ? s16(pcm, 1024, 80000, nullptr, feat[v], nullptr, nullptr, flags, 7, 0, nullptr)
                                          : f32(wav, 1024, 80000, nullptr, feat[v], nullptr, nullptr, flags, 7, 0, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s\"%s_%s_%d\": %.4f", first ? "" : ", ", is16 ? "s16" : "f32", v ? "fold" : "default", rep, ms / 20.0f);
        first = false;
      }
  printf("}}\n");
  return bad_total ? 1 : 0;
}
