// AdvancedFusionModel.forward for a HANDFUL of rows (batch <= 8): the streaming path and the per-segment
// reference call (src/processors/streaming_processor.py:295-304 runs the model on ONE row per chunk).
//
// Reference: /root/reference/src/models/fusion_model.py:44-98 (layers), :296-321 (_fuse_face_audio),
// :386-408 (_fuse_all), eval mode.
//
// With one row a layer is a matrix-vector product: there is nothing for a 128-row MMA tile to do, and the
// tensor-core chain's cost (TMEM allocation, a TMA/MMA pipeline over K, cluster exchange of the LayerNorm
// statistics) is pure latency, 5 launches x ~23 us.  Here a layer is one launch of small CTAs that stream the
// fp32 weights (18 MB in all, L2-resident between chunks) straight from nn.Linear's [N, K] layout:
//   prologue  every CTA stages the previous layer's RAW output rows [B, K] in shared memory and applies that
//             layer's LayerNorm (+ ReLU) itself (per 512-column segment for the concatenated branches), so a
//             layer needs no grid-wide reduction and no second launch for its normalisation;
//   body      one warp per output column: coalesced 16-byte weight loads, the rows from shared memory, fp32
//             FMAs, a shuffle reduction, bias;  the last layer (512 -> 7) also writes the argmax.
// Five launches (projections, processors, fusion.0 / fusion2, fusion.4, fusion.8), fp32 throughout: results agree
// with the tensor-core path to fp32 rounding (tests/test_gpu_fusion.py).
#include <cuda_runtime.h>

#include <cstdint>

#include "msa_api_internal.h"
#include "msa_fusion_common.cuh"

namespace msa {

constexpr int kRowsThreads = 128, kRowsWarps = kRowsThreads / 32;
constexpr int kRowsMaxQ = 1536 / 128, kRowsMaxS = (kTextDim + 31) / 32;     // registers of a preloaded weight row per lane

struct RowJob {
  const float* x;            // raw input rows [B, K], row stride ldx
  int ldx, K, nseg;          // LayerNorm over each of the nseg segments of K / nseg columns
  const float* gamma[3];
  const float* beta[3];
  int relu;
  const float* W;            // nn.Linear weight [N, K]
  const float* bias;
  int N;
  float* y;                  // raw outputs y[b * ldy + n]
  int ldy;
};
struct RowLayer {
  RowJob job[3];
  int B;
  int per_cta;               // output columns per CTA (a multiple of the 4 warps)
  int32_t* argmax;           // not null: last layer (N = 7, one CTA), also writes argmax[b]
};

template <int NB>
__global__ void __launch_bounds__(kRowsThreads) rows_linear_kernel(const RowLayer L) {
  extern __shared__ __align__(16) float xs[];                     // [NB][K]
  const RowJob& J = L.job[blockIdx.y];
  const int K = J.K, B = L.B, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * L.per_cta;
  // Programmatic dependent launch: this grid is launched while the previous kernel (the feature kernel, or the layer
  // before) still runs.  Everything that does not depend on it happens first: the warp pulls the weight row of its
  // first output column (constant, L2-resident between chunks) into registers; then the next layer may be launched, and
  // only then the warp waits for the previous kernel's results.  A layer's launch latency and its weight fetch are
  // thereby off the chain of five dependent launches that a streaming hop is.
  const int nfirst = n0 + warp;
  const bool vec = (K & 3) == 0;
  float4 wq[kRowsMaxQ];                                            // vec: K / 128 float4 per lane (K <= 1536)
  float ws1[kRowsMaxS];                                            // scalar rows (K = 27 / 31 / 783): K / 32 floats per lane
  const bool pre = n0 < J.N && nfirst < n0 + L.per_cta && nfirst < J.N;
  if (pre) {
    const float* w = J.W + (size_t)nfirst * K;
    if (vec) {
#pragma unroll
      for (int i = 0; i < kRowsMaxQ; ++i) {
        const int k = 4 * lane + 128 * i;
        wq[i] = (k < K) ? __ldg(reinterpret_cast<const float4*>(w + k)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kRowsMaxS; ++i) {
        const int k = lane + 32 * i;
        ws1[i] = (k < K) ? __ldg(w + k) : 0.0f;
      }
    }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (n0 >= J.N) return;                                          // a job with fewer columns than the widest of the launch
  for (int i = threadIdx.x; i < NB * K; i += kRowsThreads) {
    const int b = i / K, k = i - b * K;
    xs[i] = (b < B) ? __ldg(J.x + (size_t)b * J.ldx + k) : 0.0f;
  }
  __syncthreads();
  // LayerNorm (biased variance, eps 1e-5) of every (row, segment), one warp each
  const int segw = K / J.nseg;
  for (int p = warp; p < B * J.nseg; p += kRowsWarps) {
    const int b = p / J.nseg, s = p - b * J.nseg;
    float* v = xs + b * K + s * segw;
    const float* g = J.gamma[s];
    const float* be = J.beta[s];
    float sum = 0.0f;
    for (int k = lane; k < segw; k += 32) sum += v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)segw;
    float sq = 0.0f;
    for (int k = lane; k < segw; k += 32) { const float d = v[k] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / (float)segw + 1e-5f);
    for (int k = lane; k < segw; k += 32) {
      float t = (v[k] - mean) * rstd * __ldg(g + k) + __ldg(be + k);
      if (J.relu) t = fmaxf(t, 0.0f);
      v[k] = t;
    }
  }
  __syncthreads();
  for (int n = n0 + warp; n < n0 + L.per_cta && n < J.N; n += kRowsWarps) {
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.0f;
    const float* w = J.W + (size_t)n * K;
    if (n == nfirst && pre) {                                     // the preloaded row: same products in the same order as below
      if (vec) {
#pragma unroll
        for (int i = 0; i < kRowsMaxQ; ++i) {
          const int k = 4 * lane + 128 * i;
          if (k < K) {
            const float4 q = wq[i];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const float4 xv = *reinterpret_cast<const float4*>(xs + b * K + k);
              acc[b] = fmaf(q.w, xv.w, fmaf(q.z, xv.z, fmaf(q.y, xv.y, fmaf(q.x, xv.x, acc[b]))));
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kRowsMaxS; ++i) {
          const int k = lane + 32 * i;
          if (k < K) {
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b] = fmaf(ws1[i], xs[b * K + k], acc[b]);
          }
        }
      }
    } else if ((K & 3) == 0) {                                    // rows of W are 16-byte aligned (tensors start on 256 bytes)
#pragma unroll 4
      for (int k = 4 * lane; k < K; k += 128) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + b * K + k);
          acc[b] = fmaf(q.w, xv.w, fmaf(q.z, xv.z, fmaf(q.y, xv.y, fmaf(q.x, xv.x, acc[b]))));
        }
      }
    } else {
#pragma unroll 4
      for (int k = lane; k < K; k += 32) {
        const float q = __ldg(w + k);
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = fmaf(q, xs[b * K + k], acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    if (lane == 0) {
      const float bv = __ldg(J.bias + n);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < B) J.y[(size_t)b * J.ldy + n] = acc[b] + bv;
    }
  }
  if (L.argmax != nullptr) {                                      // last layer: one CTA holds all 7 logits of every row
    __syncthreads();
    if (threadIdx.x < B) {
      const float* l = J.y + (size_t)threadIdx.x * J.ldy;
      int best = 0;
      float bv = l[0];
      for (int k = 1; k < J.N; ++k)
        if (l[k] > bv) { bv = l[k]; best = k; }
      L.argmax[threadIdx.x] = best;
    }
  }
}

static void launch_rows(const RowLayer& L, int njobs, int maxN, int maxK, cudaStream_t s) {
  const dim3 grid((maxN + L.per_cta - 1) / L.per_cta, njobs);
  const int nb = L.B <= 1 ? 1 : (L.B <= 2 ? 2 : (L.B <= 4 ? 4 : 8));
  const size_t smem = (size_t)nb * maxK * sizeof(float);          // <= 8 * 1536 * 4 = 48 KB
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kRowsThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  switch (nb) {
    case 1: cudaLaunchKernelEx(&cfg, rows_linear_kernel<1>, L); break;
    case 2: cudaLaunchKernelEx(&cfg, rows_linear_kernel<2>, L); break;
    case 4: cudaLaunchKernelEx(&cfg, rows_linear_kernel<4>, L); break;
    default: cudaLaunchKernelEx(&cfg, rows_linear_kernel<8>, L); break;
  }
  note_launches(1);
}

int fusion_forward_rows(const float* face, const float* audio, const float* text, int B, const unsigned char* packed,
                        const PackedHeader& h, unsigned char* ws, const Workspace& wl, float* logits7, int32_t* argmax,
                        cudaStream_t s) {
  if (B < 1 || B > kFusionRowsMaxBatch) return MSA_ERR_BAD_ARGUMENT;
  auto T = [&](int t) { return reinterpret_cast<const float*>(packed + h.f32_off[t]); };
  float* proj = reinterpret_cast<float*>(ws + wl.f32_a);          // [B, 3 * 1024] raw projections, later [B, 1024] raw fusion.0
  float* cat = reinterpret_cast<float*>(ws + wl.f32_b);           // [B, 1536] raw processor outputs, later [B, 512] raw fusion.4
  const bool three = text != nullptr;
  const int nm = three ? 3 : 2, cat_w = nm * kHalf;
  const float* xin[3] = {face, audio, text};
  const int din[3] = {kFaceDim, kAudioDim, kTextDim};
  const int t_norm[3] = {T_FACE_NORM_W, T_AUDIO_NORM_W, T_TEXT_NORM_W}, t_proj[3] = {T_FACE_PROJ_W, T_AUDIO_PROJ_W, T_TEXT_PROJ_W};
  const int t_p0[3] = {T_FACE_P0_W, T_AUDIO_P0_W, T_TEXT_P0_W}, t_p3[3] = {T_FACE_P3_W, T_AUDIO_P3_W, T_TEXT_P3_W};
  const int t_p4[3] = {T_FACE_P4_W, T_AUDIO_P4_W, T_TEXT_P4_W};
  RowLayer L{};
  L.B = B;
  L.per_cta = kRowsWarps;
  L.argmax = nullptr;
  // 1: LN(d) -> Linear(d, 1024), all modalities in one launch
  for (int m = 0; m < nm; ++m)
    L.job[m] = RowJob{xin[m], din[m], din[m], 1, {T(t_norm[m]), nullptr, nullptr}, {T(t_norm[m] + 1), nullptr, nullptr}, 0,
                      T(t_proj[m]), T(t_proj[m] + 1), kHidden, proj + m * kHidden, 3 * kHidden};
  launch_rows(L, nm, kHidden, three ? kTextDim : kAudioDim, s);
  // 2: LN(1024) -> ReLU -> Linear(1024, 512), written side by side = the concatenation
  for (int m = 0; m < nm; ++m)
    L.job[m] = RowJob{proj + m * kHidden, 3 * kHidden, kHidden, 1, {T(t_p0[m]), nullptr, nullptr}, {T(t_p0[m] + 1), nullptr, nullptr}, 1,
                      T(t_p3[m]), T(t_p3[m] + 1), kHalf, cat + m * kHalf, cat_w};
  launch_rows(L, nm, kHalf, kHidden, s);
  // 3: per-branch LN(512) -> ReLU, concat -> fusion.0 (1536 -> 1024) or fusion2 (1024 -> 1024)
  L.job[0] = RowJob{cat, cat_w, cat_w, nm, {T(t_p4[0]), T(t_p4[1]), T(t_p4[2])}, {T(t_p4[0] + 1), T(t_p4[1] + 1), T(t_p4[2] + 1)}, 1,
                    T(three ? T_FUS0_W : T_FUS2_W), T(three ? T_FUS0_B : T_FUS2_B), kHidden, proj, kHidden};
  launch_rows(L, 1, kHidden, cat_w, s);
  // 4: LN(1024) -> ReLU -> fusion.4 (1024 -> 512)
  L.job[0] = RowJob{proj, kHidden, kHidden, 1, {T(T_FUS1_W), nullptr, nullptr}, {T(T_FUS1_B), nullptr, nullptr}, 1,
                    T(T_FUS4_W), T(T_FUS4_B), kHalf, cat, kHalf};
  launch_rows(L, 1, kHalf, kHidden, s);
  // 5: LN(512) -> ReLU -> fusion.8 (512 -> 7) + argmax, one CTA
  L.job[0] = RowJob{cat, kHalf, kHalf, 1, {T(T_FUS5_W), nullptr, nullptr}, {T(T_FUS5_B), nullptr, nullptr}, 1,
                    T(T_FUS8_W), T(T_FUS8_B), kOut, logits7, kOut};
  L.per_cta = 2 * kRowsWarps;
  L.argmax = argmax;
  launch_rows(L, 1, kOut, kHalf, s);
  return (int)cudaGetLastError();
}

}  // namespace msa
