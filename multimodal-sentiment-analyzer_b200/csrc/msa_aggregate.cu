// Speaker / timeline aggregation over gathered per-segment results.
//
// Restates /root/reference/src/processors/offline_processor.py:259-298 on device:
//   - group segments by speaker (in segment order),
//   - dominant emotion per speaker = mode of the argmax labels (:287-290; ties -> smallest label,
//     which is what max(set(ints), key=count) yields for small ints),
//   - "patterns": every position i of a speaker's own sequence with e[i] == e[i+1] == e[i+2] (:293-298).
// Integer work only: results are bit-exact and order-independent (integer atomics).
#include <cuda_runtime.h>

#include <cstdint>

#include "msa_api_internal.h"

namespace msa {

constexpr int kEmo = 7;

__global__ void aggregate_segments_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ speaker, int S,
                                          int n_speakers, int32_t* __restrict__ hist, int32_t* __restrict__ run3) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int p = speaker[s], e = label[s];
  if (p < 0 || p >= n_speakers || e < 0 || e >= kEmo) { run3[s] = 0; return; }
  atomicAdd(&hist[p * kEmo + e], 1);
  // next two segments of the same speaker
  int found = 0, ok = 1;
  for (int t = s + 1; t < S && found < 2; ++t) {
    if (speaker[t] == p) { ok &= (label[t] == e); ++found; }
  }
  run3[s] = (found == 2 && ok) ? 1 : 0;
}

__global__ void dominant_kernel(const int32_t* __restrict__ hist, int n_speakers, int32_t* __restrict__ dominant) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_speakers) return;
  int best = -1, bc = 0;
  for (int e = 0; e < kEmo; ++e) {
    const int c = hist[p * kEmo + e];
    if (c > bc) { bc = c; best = e; }
  }
  dominant[p] = best;          // -1: speaker without segments
}

// result table row = audio row 31 | logits 7 | argmax | segment id (40 x 32-bit words)
__global__ void pack_rows_kernel(const float* __restrict__ audio, const float* __restrict__ logits,
                                 const int32_t* __restrict__ argmax, int first_id, int n, float* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 40) return;
  const int r = i / 40, c = i - r * 40;
  float v;
  if (c < 31) v = audio[r * 31 + c];
  else if (c < 38) v = logits[r * 7 + (c - 31)];
  else v = __int_as_float(c == 38 ? argmax[r] : first_id + r);
  rows[i] = v;
}

}  // namespace msa

extern "C" int msa_pack_rows(const float* audio31, const float* logits7, const int32_t* argmax, int first_id, int n, float* rows40,
                             void* stream) {
  msa::reset_launches();
  if (!audio31 || !logits7 || !argmax || !rows40 || n < 0) return MSA_ERR_BAD_ARGUMENT;
  if (n == 0) return MSA_OK;
  msa::pack_rows_kernel<<<(n * 40 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(audio31, logits7, argmax, first_id, n, rows40);
  msa::note_launches(1);
  return (int)cudaGetLastError();
}

// label [S] int32 (argmax of the fused logits), speaker [S] int32 in [0, n_speakers)
// hist [n_speakers, 7] int32 out, dominant [n_speakers] int32 out, run3 [S] int32 out
extern "C" int msa_aggregate_speakers(const int32_t* label, const int32_t* speaker, int S, int n_speakers, int32_t* hist,
                                      int32_t* dominant, int32_t* run3, void* stream) {
  using namespace msa;
  reset_launches();
  if (!label || !speaker || !hist || !dominant || !run3 || S < 0 || n_speakers < 1) return MSA_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(int32_t) * n_speakers * kEmo, st);
  if (e != cudaSuccess) return (int)e;
  if (S > 0) {
    aggregate_segments_kernel<<<(S + 255) / 256, 256, 0, st>>>(label, speaker, S, n_speakers, hist, run3);
    note_launches(1);
  }
  dominant_kernel<<<(n_speakers + 127) / 128, 128, 0, st>>>(hist, n_speakers, dominant);
  note_launches(1);
  return (int)cudaGetLastError();
}
