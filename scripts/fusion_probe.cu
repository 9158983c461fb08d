// One process, one batch size: the 3-modal fusion forward of libmsa_b200.so (random state_dict and rows), warmed up and
// timed with CUDA events.  Small enough to sit under ncu:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/fusion_probe scripts/fusion_probe.cu -ldl
//   ./scripts/fusion_probe [lib.so] [B] [reps]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
typedef int (*fus_num_t)(void);
typedef size_t (*fus_numel_t)(int);
typedef size_t (*fus_bytes_t)(void);
typedef size_t (*fus_ws_t)(int);
typedef int (*fus_pack_t)(const float* const*, void*, void*);
typedef int (*fus_fwd_t)(const float*, const float*, const float*, int, const void*, void*, size_t, float*, int32_t*, void*);
typedef const char* (*fus_name_t)(int);
__global__ void fill(float* p, size_t n, unsigned seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = scale * ((x & 0xffff) * (1.0f / 65536.0f) - 0.5f);
  }
}
int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "multimodal-sentiment-analyzer_b200/libmsa_b200.so";
  const int B = argc > 2 ? atoi(argv[2]) : 65536, reps = argc > 3 ? atoi(argv[3]) : 10;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("{\"error\": \"%s\"}\n", dlerror()); return 2; }
  fus_num_t num = (fus_num_t)dlsym(h, "msa_fusion_num_tensors");
  fus_numel_t numel = (fus_numel_t)dlsym(h, "msa_fusion_tensor_numel");
  fus_name_t tname = (fus_name_t)dlsym(h, "msa_fusion_tensor_name");
  fus_bytes_t pbytes = (fus_bytes_t)dlsym(h, "msa_fusion_packed_bytes");
  fus_ws_t wbytes = (fus_ws_t)dlsym(h, "msa_fusion_workspace_bytes");
  fus_pack_t pack = (fus_pack_t)dlsym(h, "msa_fusion_pack");
  fus_fwd_t fwd = (fus_fwd_t)dlsym(h, "msa_fusion_forward");
  if (!num || !numel || !tname || !pbytes || !wbytes || !pack || !fwd) { printf("{\"error\": \"symbols\"}\n"); return 2; }
  float *face, *audio, *text, *logits; int32_t* amax; void *packed, *ws;
  CK(cudaMalloc(&face, (size_t)B * 27 * 4)); CK(cudaMalloc(&audio, (size_t)B * 31 * 4)); CK(cudaMalloc(&text, (size_t)B * 783 * 4));
  CK(cudaMalloc(&logits, (size_t)B * 7 * 4)); CK(cudaMalloc(&amax, (size_t)B * 4));
  fill<<<592, 256>>>(face, (size_t)B * 27, 1u, 2.0f); fill<<<592, 256>>>(audio, (size_t)B * 31, 2u, 2.0f); fill<<<592, 256>>>(text, (size_t)B * 783, 3u, 2.0f);
  const int nt = num();
  std::vector<std::vector<float>> tens(nt);
  std::vector<const float*> ptr(nt);
  unsigned st = 777u;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xffff) * (1.0f / 65536.0f) - 0.5f; };
  for (int i = 0; i < nt; ++i) {
    const size_t n = numel(i);
    const char* nm = tname(i);
    const bool is_w = strstr(nm, "weight") != nullptr, is_norm = strstr(nm, "norm") != nullptr;
    tens[i].resize(n);
    for (size_t k = 0; k < n; ++k) tens[i][k] = (is_w && !is_norm && n > 2048) ? 0.1f * rnd() : ((is_w && n > 3) ? 1.0f + 0.2f * rnd() : (n > 3 ? 0.2f * rnd() : rnd()));
    ptr[i] = tens[i].data();
  }
  CK(cudaMalloc(&packed, pbytes()));
  if (pack(ptr.data(), packed, nullptr)) { printf("{\"error\": \"pack\"}\n"); return 2; }
  const size_t wsb = wbytes(B);
  CK(cudaMalloc(&ws, wsb));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 3; ++i) if (fwd(face, audio, text, B, packed, ws, wsb, logits, amax, nullptr)) { printf("{\"error\": \"forward\"}\n"); return 2; }
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) fwd(face, audio, text, B, packed, ws, wsb, logits, amax, nullptr);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  float l0[7]; CK(cudaMemcpy(l0, logits, 28, cudaMemcpyDeviceToHost));
  printf("{\"lib\": \"%s\", \"B\": %d, \"us_per_forward\": %.2f, \"logit0\": %.6g", path, B, 1000.0f * ms / reps, l0[0]);
  // MSA_TC_TRACE builds: mean clock64 distance between the phase stamps of a CTA, per launch of the forward
  typedef int (*trace_t)(long long*, size_t);
  trace_t trace = (trace_t)dlsym(h, "msa_debug_tc_trace");
  if (trace) {
    const int S = 4, E = 8, C = 4096;
    std::vector<long long> t((size_t)S * E * C);
    if (trace(t.data(), t.size() * 8) == 0) {
      const char* names[4] = {"proj", "proc", "fusion0", "fusion4"};
      printf(", \"trace_cycles\": {");
      for (int sl = 0; sl < S; ++sl) {
        auto at = [&](int e, int c) { return t[((size_t)sl * E + e) * C + c]; };
        if (getenv("MSA_PAIR_MODE") == nullptr || getenv("MSA_PAIR_MODE")[0] != '0') {
          // persistent kernel: per-CTA SUMS over its tiles; event 7 carries the tile count in its upper bits
          double a[8] = {0}; long long nc = 0, nlead = 0, tiles = 0;
          for (int c = 0; c < C; ++c) {
            const long long e7 = at(7, c);
            if (e7 == 0) continue;
            ++nc; tiles += e7 >> 40;
            for (int e = 3; e < 7; ++e) a[e] += (double)at(e, c);
            a[7] += (double)(e7 & ((1ll << 40) - 1));
            if (at(2, c) != 0) { ++nlead; for (int e = 0; e < 3; ++e) a[e] += (double)at(e, c); }
          }
          if (!nc) continue;
          const double tp = (double)tiles / nc;
          printf("%s\"%s\": {\"ctas\": %lld, \"tiles_per_cta\": %.2f, \"per_tile\": {\"mma_wait_tmem\": %.0f, \"mma_wait_full\": %.0f, \"mma_issue\": %.0f, \"epi_wait_accum\": %.0f, \"epi_pass1\": %.0f, \"epi_tmem_ld\": %.0f, \"epi_pass2\": %.0f, \"epi_release\": %.0f}}",
                 sl ? ", " : "", names[sl], nc, tp, nlead ? a[0] / nlead / tp : 0.0, nlead ? a[1] / nlead / tp : 0.0, nlead ? a[2] / nlead / tp : 0.0,
                 a[3] / nc / tp, a[4] / nc / tp, a[5] / nc / tp, a[6] / nc / tp, a[7] / nc / tp);
          continue;
        }
        double d[8] = {0}; long long n = 0, nl = 0; double lead[2] = {0, 0};
        for (int c = 0; c < C; ++c) {
          if (at(0, c) == 0 || at(7, c) == 0) continue;
          ++n;
          d[0] += (double)(at(1, c) - at(0, c));      // prologue
          d[1] += (double)(at(4, c) - at(1, c));      // mainloop as the epilogue sees it
          d[2] += (double)(at(5, c) - at(4, c));      // pass 1
          d[3] += (double)(at(6, c) - at(5, c));      // sync + pass 2
          d[4] += (double)(at(7, c) - at(6, c));      // final cluster sync
          d[5] += (double)(at(7, c) - at(0, c));      // whole CTA
          if (at(2, c) != 0 && at(3, c) != 0 && at(3, c) > at(2, c)) { ++nl; lead[0] += (double)(at(2, c) - at(1, c)); lead[1] += (double)(at(3, c) - at(2, c)); }
        }
        if (!n) continue;
        printf("%s\"%s\": {\"ctas\": %lld, \"prologue\": %.0f, \"mainloop\": %.0f, \"pass1\": %.0f, \"pass2\": %.0f, \"exit_sync\": %.0f, \"total\": %.0f, \"first_tile_wait\": %.0f, \"mma_issue_span\": %.0f}",
               sl ? ", " : "", names[sl], n, d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[4] / n, d[5] / n, nl ? lead[0] / nl : 0.0, nl ? lead[1] / nl : 0.0);
      }
      printf("}");
    }
  }
  printf("}\n");
  return 0;
}
