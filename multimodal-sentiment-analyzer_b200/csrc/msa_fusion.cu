// AdvancedFusionModel.forward (eval) on device: weight packing and the dispatch between the tcgen05 kernels
// (msa_fusion_tc.cu) and the matrix-vector kernels for a handful of rows (msa_fusion_rows.cu).  (The fp32 CUDA-core
// cross-check of round 1 lives in tests/xcheck/ now: test infrastructure, not part of this library.)
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
    g_impl = (e && std::strcmp(e, "tc") == 0) ? 2 : 0;
  }
  return g_impl;
}
}  // namespace msa

extern "C" int msa_fusion_set_impl(int impl) {
  if (impl != 0 && impl != 2) return MSA_ERR_BAD_ARGUMENT;
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
  // tcgen05 kernels, except that a handful of rows (the streaming path: one row per chunk) takes the matrix-vector
  // kernels of msa_fusion_rows.cu (msa_fusion_set_impl(2) forces the tensor-core kernels for every batch size)
  if (fusion_impl() == 0 && B <= kFusionRowsMaxBatch)
    return fusion_forward_rows(face, audio, text, B, static_cast<const unsigned char*>(packed), h,
                               static_cast<unsigned char*>(workspace), wl, logits7, argmax, (cudaStream_t)stream);
  return fusion_forward_tc(face, audio, text, B, static_cast<const unsigned char*>(packed), h,
                           static_cast<unsigned char*>(workspace), wl, logits7, argmax, (cudaStream_t)stream);
}
