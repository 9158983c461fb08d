// fp32 CUDA-core restatement of AdvancedFusionModel.forward (eval) ON THE DEVICE: TEST INFRASTRUCTURE.
//
// An independent implementation of /root/reference/src/models/fusion_model.py:296-321 (_fuse_face_audio) and :386-408
// (_fuse_all) - plain fp32 FFMA GEMM tiles, two-pass LayerNorm, no tensor cores, no bf16 split - that reads the SAME packed
// weight blob as the product kernels (csrc/msa_fusion_common.cuh).  It is built into tests/xcheck/libmsa_xcheck.so by
// __graft_entry__.build() and loaded by tests/test_gpu_fusion.py only (large-batch agreement checks where the fp64 numpy
// oracle would take minutes).  It is NOT part of libmsa_b200.so: the product has one fusion implementation per batch range.
#include <cuda_runtime.h>

#include <cstdint>

#include "msa_fusion_common.cuh"

#define MSA_OK 0
#define MSA_ERR_BAD_ARGUMENT (-1)
#define MSA_ERR_WORKSPACE (-4)

namespace msa {
static inline void note_launches(int) {}

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

static int fusion_forward_simt(const float* face, const float* audio, const float* text, int B, const Blob& w, unsigned char* ws,
                               float* logits7, int32_t* argmax, cudaStream_t s) {
  const size_t plane = (size_t)B * 1536 * sizeof(float);
  float* bufA = reinterpret_cast<float*>(ws);              // three fp32 [B, 1536] planes at the start of the workspace
  float* bufB = reinterpret_cast<float*>(ws + plane);
  float* cat = reinterpret_cast<float*>(ws + 2 * plane);
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

}  // namespace msa

// same signature as msa_fusion_forward (include/msa_b200.h); needs 3 * B * 1536 * 4 bytes of workspace, which the product's
// own workspace of msa_fusion_workspace_bytes(B) always covers
extern "C" int msa_xcheck_fusion_forward(const float* face, const float* audio, const float* text, int B, const void* packed,
                                         void* workspace, size_t workspace_bytes, float* logits7, int32_t* argmax, void* stream) {
  using namespace msa;
  if (!face || !audio || !packed || !workspace || !logits7 || B < 0) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;
  if (workspace_bytes < (size_t)3 * B * 1536 * sizeof(float)) return MSA_ERR_WORKSPACE;
  PackedHeader h;
  packed_layout(h);
  Blob blob{static_cast<const unsigned char*>(packed), h};
  return fusion_forward_simt(face, audio, text, B, blob, static_cast<unsigned char*>(workspace), logits7, argmax, (cudaStream_t)stream);
}
