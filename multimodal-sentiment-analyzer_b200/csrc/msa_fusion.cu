// AdvancedFusionModel.forward (eval) on device: weight packing, the fp32 SIMT bring-up /
// cross-check path, and dispatch to the tcgen05 path (msa_fusion_tc.cu).
//
// Reference: /root/reference/src/models/fusion_model.py:296-321 (_fuse_face_audio),
// :386-408 (_fuse_all), :44-98 (layers), :114-120 (init).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "msa_api_internal.h"
#include "msa_fusion_common.cuh"

namespace msa {

// ------------------------------------------------------------------------------ packing
// W (fp32 [N, K]) -> hi = bf16(W), lo = bf16(W - hi), both [N, Kpad] with zero padding.
__global__ void split_weight_kernel(const float* __restrict__ w, int N, int K, int Kpad,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const size_t total = (size_t)N * Kpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kpad), k = (int)(i % Kpad);
    float v = (k < K) ? w[(size_t)n * K + k] : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

int fusion_pack(const float* const* tensors_host, void* packed_dev, cudaStream_t stream) {
  if (!tensors_host || !packed_dev) return MSA_ERR_BAD_ARGUMENT;
  PackedHeader h;
  packed_layout(h);
  unsigned char* base = static_cast<unsigned char*>(packed_dev);
  cudaError_t e = cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < kNumTensors; ++i) {
    if (!tensors_host[i]) return MSA_ERR_BAD_ARGUMENT;
    e = cudaMemcpyAsync(base + h.f32_off[i], tensors_host[i], tensor_numel(i) * 4, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return (int)e;
  }
  for (int g = 0; g < kNumGemmWeights; ++g) {
    const GemmWeight& gw = kGemmWeights[g];
    split_weight_kernel<<<296, 256, 0, stream>>>(reinterpret_cast<const float*>(base + h.f32_off[gw.tensor]), gw.N, gw.K,
                                                 gw.Kpad, reinterpret_cast<__nv_bfloat16*>(base + h.hi_off[g]),
                                                 reinterpret_cast<__nv_bfloat16*>(base + h.lo_off[g]));
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(stream);   // host tensors may be freed by the caller after return
  return (int)e;
}

// ------------------------------------------------------------------------------ SIMT bring-up path
// C[M, N] = A[M, K] * W[N, K]^T + bias, fp32 FFMA, 64x64x16 tiles, 256 threads x (4x4).
__global__ void __launch_bounds__(256) simt_gemm_bias_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                             int ldw, const float* __restrict__ bias, float* __restrict__ C,
                                                             int ldc, int M, int N, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      const int m = m0 + r, n = n0 + r, k = k0 + c;
      As[c][r] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.0f;
      Ws[c][r] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j] + (bias ? bias[n] : 0.0f);
    }
}

// dst[row, 0:N] = act(LayerNorm(src[row, 0:N]) * gamma + beta); one warp per row, N <= 1024.
template <bool RELU>
__global__ void __launch_bounds__(256) row_ln_kernel(const float* src, int ld_src, int N,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* dst, int ld_dst, int M) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* s = src + (size_t)row * ld_src;
  float v[32];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < N) ? s[c] : 0.0f;
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)N;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < N) ? v[i] - mean : 0.0f;
    sq = fmaf(d, d, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)N + 1e-5f);
  float* d = dst + (size_t)row * ld_dst;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < N) {
      float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (RELU) y = fmaxf(y, 0.0f);
      d[c] = y;
    }
  }
}

__global__ void argmax7_kernel(const float* __restrict__ logits, int32_t* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* l = logits + (size_t)i * kOut;
  int best = 0;
  float bv = l[0];
  for (int k = 1; k < kOut; ++k)
    if (l[k] > bv) { bv = l[k]; best = k; }
  out[i] = best;
}

struct Blob {
  const unsigned char* base;
  PackedHeader h;
  const float* f32(int t) const { return reinterpret_cast<const float*>(base + h.f32_off[t]); }
};

static void simt_gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                      int K, cudaStream_t s) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  simt_gemm_bias_kernel<<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  note_launches(1);
}
static void row_ln(bool relu, const float* src, int ld_src, int N, const float* g, const float* b, float* dst, int ld_dst, int M,
                   cudaStream_t s) {
  if (relu) row_ln_kernel<true><<<(M + 7) / 8, 256, 0, s>>>(src, ld_src, N, g, b, dst, ld_dst, M);
  else row_ln_kernel<false><<<(M + 7) / 8, 256, 0, s>>>(src, ld_src, N, g, b, dst, ld_dst, M);
  note_launches(1);
}

int fusion_forward_simt(const float* face, const float* audio, const float* text, int B, const Blob& w, unsigned char* ws,
                        const Workspace& wl, float* logits7, int32_t* argmax, cudaStream_t s) {
  float* bufA = reinterpret_cast<float*>(ws + wl.f32_a);   // [B, 1536]
  float* bufB = reinterpret_cast<float*>(ws + wl.f32_b);   // [B, 1536]
  float* cat = reinterpret_cast<float*>(ws + wl.h_hi[0]);  // bring-up path re-uses the bf16 activation area as fp32 [B,1536]
  const bool three = text != nullptr;
  const int cat_w = three ? 1536 : 1024;
  struct Mod { const float* x; int d; int nw, nb, pw, pb, l0w, l0b, p3w, p3b, l4w, l4b; };
  const Mod mods[3] = {
      {face, kFaceDim, T_FACE_NORM_W, T_FACE_NORM_B, T_FACE_PROJ_W, T_FACE_PROJ_B, T_FACE_P0_W, T_FACE_P0_B, T_FACE_P3_W, T_FACE_P3_B, T_FACE_P4_W, T_FACE_P4_B},
      {audio, kAudioDim, T_AUDIO_NORM_W, T_AUDIO_NORM_B, T_AUDIO_PROJ_W, T_AUDIO_PROJ_B, T_AUDIO_P0_W, T_AUDIO_P0_B, T_AUDIO_P3_W, T_AUDIO_P3_B, T_AUDIO_P4_W, T_AUDIO_P4_B},
      {text, kTextDim, T_TEXT_NORM_W, T_TEXT_NORM_B, T_TEXT_PROJ_W, T_TEXT_PROJ_B, T_TEXT_P0_W, T_TEXT_P0_B, T_TEXT_P3_W, T_TEXT_P3_B, T_TEXT_P4_W, T_TEXT_P4_B},
  };
  for (int m = 0; m < (three ? 3 : 2); ++m) {
    const Mod& md = mods[m];
    row_ln(false, md.x, md.d, md.d, w.f32(md.nw), w.f32(md.nb), bufB, md.d, B, s);
    simt_gemm(bufB, md.d, w.f32(md.pw), md.d, w.f32(md.pb), bufA, kHidden, B, kHidden, md.d, s);
    row_ln(true, bufA, kHidden, kHidden, w.f32(md.l0w), w.f32(md.l0b), bufA, kHidden, B, s);
    simt_gemm(bufA, kHidden, w.f32(md.p3w), kHidden, w.f32(md.p3b), bufB, kHalf, B, kHalf, kHidden, s);
    row_ln(true, bufB, kHalf, kHalf, w.f32(md.l4w), w.f32(md.l4b), cat + m * kHalf, cat_w, B, s);
  }
  if (three) simt_gemm(cat, 1536, w.f32(T_FUS0_W), 1536, w.f32(T_FUS0_B), bufA, kHidden, B, kHidden, 1536, s);
  else simt_gemm(cat, 1024, w.f32(T_FUS2_W), 1024, w.f32(T_FUS2_B), bufA, kHidden, B, kHidden, 1024, s);
  row_ln(true, bufA, kHidden, kHidden, w.f32(T_FUS1_W), w.f32(T_FUS1_B), bufA, kHidden, B, s);
  simt_gemm(bufA, kHidden, w.f32(T_FUS4_W), kHidden, w.f32(T_FUS4_B), bufB, kHalf, B, kHalf, kHidden, s);
  row_ln(true, bufB, kHalf, kHalf, w.f32(T_FUS5_W), w.f32(T_FUS5_B), bufB, kHalf, B, s);
  simt_gemm(bufB, kHalf, w.f32(T_FUS8_W), kHalf, w.f32(T_FUS8_B), logits7, kOut, B, kOut, kHalf, s);
  if (argmax) { argmax7_kernel<<<(B + 255) / 256, 256, 0, s>>>(logits7, argmax, B); note_launches(1); }
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

int fusion_forward_tc(const float* face, const float* audio, const float* text, int B, const unsigned char* packed,
                      const PackedHeader& h, unsigned char* ws, const Workspace& wl, float* logits7, int32_t* argmax,
                      cudaStream_t s);   // msa_fusion_tc.cu
int fusion_forward_rows(const float* face, const float* audio, const float* text, int B, const unsigned char* packed,
                        const PackedHeader& h, unsigned char* ws, const Workspace& wl, float* logits7, int32_t* argmax,
                        cudaStream_t s);   // msa_fusion_rows.cu

static int fusion_impl();

// the header lives in device memory; keep a host copy keyed by blob pointer (layout is static anyway)
static const PackedHeader& host_header() {
  static PackedHeader h;
  static bool init = false;
  if (!init) { packed_layout(h); init = true; }
  return h;
}

}  // namespace msa

namespace msa {
static int g_impl = -1;
static int fusion_impl() {
  if (g_impl < 0) {
    const char* e = std::getenv("MSA_FUSION_IMPL");
    g_impl = (e && std::strcmp(e, "simt") == 0) ? 1 : 0;
  }
  return g_impl;
}
}  // namespace msa

extern "C" int msa_fusion_set_impl(int impl) {
  if (impl < 0 || impl > 2) return MSA_ERR_BAD_ARGUMENT;
  msa::g_impl = impl;
  return MSA_OK;
}

extern "C" size_t msa_fusion_packed_bytes(void) { return msa::host_header().total_bytes; }
extern "C" size_t msa_fusion_workspace_bytes(int B) {
  msa::Workspace w;
  msa::workspace_layout(B < 1 ? 1 : B, w);
  return w.total;
}
extern "C" int msa_fusion_num_tensors(void) { return msa::kNumTensors; }
extern "C" const char* msa_fusion_tensor_name(int i) { return (i >= 0 && i < msa::kNumTensors) ? msa::kTensors[i].name : nullptr; }
extern "C" size_t msa_fusion_tensor_numel(int i) { return (i >= 0 && i < msa::kNumTensors) ? msa::tensor_numel(i) : 0; }
extern "C" int msa_fusion_pack(const float* const* tensors_host, void* packed_dev, void* stream) {
  return msa::fusion_pack(tensors_host, packed_dev, (cudaStream_t)stream);
}

extern "C" int msa_fusion_forward(const float* face, const float* audio, const float* text, int B, const void* packed,
                                  void* workspace, size_t workspace_bytes, float* logits7, int32_t* argmax, void* stream) {
  using namespace msa;
  reset_launches();
  if (!face || !audio || !packed || !workspace || !logits7 || B < 0) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  Workspace wl;
  workspace_layout(B, wl);
  if (workspace_bytes < wl.total) return MSA_ERR_WORKSPACE;
  const PackedHeader& h = host_header();
  // MSA_FUSION_IMPL=simt selects the fp32 CUDA-core bring-up kernels (on-device cross-check of the
  // tensor-core path; same C ABI, same results to fp32 rounding).  Default: tcgen05, except that a handful
  // of rows (the streaming path: one row per chunk) takes the matrix-vector kernels of msa_fusion_rows.cu.
  if (fusion_impl() == 0 && B <= kFusionRowsMaxBatch)
    return fusion_forward_rows(face, audio, text, B, static_cast<const unsigned char*>(packed), h,
                               static_cast<unsigned char*>(workspace), wl, logits7, argmax, (cudaStream_t)stream);
  if (fusion_impl() == 1) {
    Blob blob{static_cast<const unsigned char*>(packed), h};
    return fusion_forward_simt(face, audio, text, B, blob, static_cast<unsigned char*>(workspace), wl, logits7, argmax,
                               (cudaStream_t)stream);
  }
  return fusion_forward_tc(face, audio, text, B, static_cast<const unsigned char*>(packed), h,
                           static_cast<unsigned char*>(workspace), wl, logits7, argmax, (cudaStream_t)stream);
}
