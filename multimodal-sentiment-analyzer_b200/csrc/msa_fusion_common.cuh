// Shared definitions of the fusion-model kernels: state_dict tensor table, packed-weight blob
// layout and activation workspace layout.
//
// Network (src/models/fusion_model.py:44-98, eval mode):
//   per modality m in {face 27, audio 31, text 783}:
//     LN(d_m) -> Linear(d_m, 1024) -> LN(1024) -> ReLU -> Linear(1024, 512) -> LN(512) -> ReLU
//   3-modal: concat [face|audio|text] 1536 -> Linear(1536,1024) -> LN -> ReLU -> Linear(1024,512)
//            -> LN -> ReLU -> Linear(512,7)
//   face+audio: concat [face|audio] 1024 -> fusion2 Linear(1024,1024) -> same tail (fusion[1:]).
#pragma once
#include <cstddef>
#include <cstdint>

namespace msa {

constexpr int kFaceDim = 27, kAudioDim = 31, kTextDim = 783, kHidden = 1024, kHalf = 512, kOut = 7;
constexpr int kFusionRowsMaxBatch = 8;                     // batches up to this size run as matrix-vector products (msa_fusion_rows.cu)
constexpr int kFaceK = 64, kAudioK = 64, kTextK = 832;     // K padded to a multiple of 64 (one 128-byte swizzle row of bf16)

struct TensorInfo { const char* name; int rows; int cols; };   // cols == 0: vector of `rows`

// order of the `tensors` array of msa_fusion_pack
static const TensorInfo kTensors[] = {
    {"face_norm.weight", 27, 0},  {"face_norm.bias", 27, 0},
    {"audio_norm.weight", 31, 0}, {"audio_norm.bias", 31, 0},
    {"text_norm.weight", 783, 0}, {"text_norm.bias", 783, 0},
    {"face_proj.weight", 1024, 27},  {"face_proj.bias", 1024, 0},
    {"audio_proj.weight", 1024, 31}, {"audio_proj.bias", 1024, 0},
    {"text_proj.weight", 1024, 783}, {"text_proj.bias", 1024, 0},
    {"face_processor.0.weight", 1024, 0}, {"face_processor.0.bias", 1024, 0},
    {"face_processor.3.weight", 512, 1024}, {"face_processor.3.bias", 512, 0},
    {"face_processor.4.weight", 512, 0}, {"face_processor.4.bias", 512, 0},
    {"audio_processor.0.weight", 1024, 0}, {"audio_processor.0.bias", 1024, 0},
    {"audio_processor.3.weight", 512, 1024}, {"audio_processor.3.bias", 512, 0},
    {"audio_processor.4.weight", 512, 0}, {"audio_processor.4.bias", 512, 0},
    {"text_processor.0.weight", 1024, 0}, {"text_processor.0.bias", 1024, 0},
    {"text_processor.3.weight", 512, 1024}, {"text_processor.3.bias", 512, 0},
    {"text_processor.4.weight", 512, 0}, {"text_processor.4.bias", 512, 0},
    {"fusion.0.weight", 1024, 1536}, {"fusion.0.bias", 1024, 0},
    {"fusion.1.weight", 1024, 0}, {"fusion.1.bias", 1024, 0},
    {"fusion.4.weight", 512, 1024}, {"fusion.4.bias", 512, 0},
    {"fusion.5.weight", 512, 0}, {"fusion.5.bias", 512, 0},
    {"fusion.8.weight", 7, 512}, {"fusion.8.bias", 7, 0},
    {"fusion2.weight", 1024, 1024}, {"fusion2.bias", 1024, 0},
};
constexpr int kNumTensors = sizeof(kTensors) / sizeof(kTensors[0]);
enum TensorId {
  T_FACE_NORM_W, T_FACE_NORM_B, T_AUDIO_NORM_W, T_AUDIO_NORM_B, T_TEXT_NORM_W, T_TEXT_NORM_B,
  T_FACE_PROJ_W, T_FACE_PROJ_B, T_AUDIO_PROJ_W, T_AUDIO_PROJ_B, T_TEXT_PROJ_W, T_TEXT_PROJ_B,
  T_FACE_P0_W, T_FACE_P0_B, T_FACE_P3_W, T_FACE_P3_B, T_FACE_P4_W, T_FACE_P4_B,
  T_AUDIO_P0_W, T_AUDIO_P0_B, T_AUDIO_P3_W, T_AUDIO_P3_B, T_AUDIO_P4_W, T_AUDIO_P4_B,
  T_TEXT_P0_W, T_TEXT_P0_B, T_TEXT_P3_W, T_TEXT_P3_B, T_TEXT_P4_W, T_TEXT_P4_B,
  T_FUS0_W, T_FUS0_B, T_FUS1_W, T_FUS1_B, T_FUS4_W, T_FUS4_B, T_FUS5_W, T_FUS5_B, T_FUS8_W, T_FUS8_B,
  T_FUS2_W, T_FUS2_B,
};

inline size_t tensor_numel(int i) { return (size_t)kTensors[i].rows * (kTensors[i].cols ? kTensors[i].cols : 1); }

// The seven big Linear layers run on tensor cores from split-bf16 copies: W = hi + lo, both bf16,
// K-major [N, Kpad] (nn.Linear's own [out, in] layout, K zero-padded to a multiple of 64).
struct GemmWeight { int tensor; int N; int K; int Kpad; };
static const GemmWeight kGemmWeights[] = {
    {T_FACE_PROJ_W, 1024, 27, kFaceK},   {T_AUDIO_PROJ_W, 1024, 31, kAudioK}, {T_TEXT_PROJ_W, 1024, 783, kTextK},
    {T_FACE_P3_W, 512, 1024, 1024},      {T_AUDIO_P3_W, 512, 1024, 1024},     {T_TEXT_P3_W, 512, 1024, 1024},
    {T_FUS0_W, 1024, 1536, 1536},        {T_FUS4_W, 512, 1024, 1024},         {T_FUS2_W, 1024, 1024, 1024},
};
constexpr int kNumGemmWeights = sizeof(kGemmWeights) / sizeof(kGemmWeights[0]);
enum GemmId { G_FACE_PROJ, G_AUDIO_PROJ, G_TEXT_PROJ, G_FACE_P3, G_AUDIO_P3, G_TEXT_P3, G_FUS0, G_FUS4, G_FUS2 };

// Packed blob: [header][fp32 copy of all 42 tensors][bf16 hi | bf16 lo per GEMM weight]
struct PackedHeader {
  uint32_t magic;                       // 'MSAF'
  uint32_t version;
  uint64_t f32_off[kNumTensors];        // byte offsets from blob start
  uint64_t hi_off[kNumGemmWeights];
  uint64_t lo_off[kNumGemmWeights];
  uint64_t total_bytes;
};
constexpr uint32_t kPackedMagic = 0x4D534146u;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline void packed_layout(PackedHeader& h) {
  h.magic = kPackedMagic;
  h.version = 1;
  size_t off = align_up(sizeof(PackedHeader), 256);
  for (int i = 0; i < kNumTensors; ++i) { h.f32_off[i] = off; off = align_up(off + tensor_numel(i) * 4, 256); }
  for (int g = 0; g < kNumGemmWeights; ++g) {
    const size_t bytes = (size_t)kGemmWeights[g].N * kGemmWeights[g].Kpad * 2;
    h.hi_off[g] = off; off = align_up(off + bytes, 256);
    h.lo_off[g] = off; off = align_up(off + bytes, 256);
  }
  h.total_bytes = off;
}

constexpr int kSmallBatchRows = 3072;                      // at or below: one CTA per 128 rows x 128 columns; above: CTA pairs per 256 rows x 512 columns (measured crossover: 197 vs 208 us at 3072 rows, 251 vs 216 us at 4096; profiles/r2_v204_fusion_threshold.txt)

// Activation workspace for batch B (rows padded to one MMA tile of rows: 128, or 256 where CTA pairs run).
struct Workspace {
  size_t x_face_hi, x_face_lo, x_audio_hi, x_audio_lo, x_text_hi, x_text_lo;   // bf16 [Bp, Kpad] LayerNorm'd inputs
  size_t h_hi[3], h_lo[3];                                                     // bf16 [Bp, 1024] per modality
  size_t cat_hi, cat_lo;                                                       // bf16 [Bp, 1536]
  size_t f1_hi, f1_lo;                                                         // bf16 [Bp, 1024]
  size_t f32_a, f32_b;                                                         // fp32 [128, 3072] scratch of the matrix-vector path (batches <= 8)
  size_t total;
  int Bp;
};

inline void workspace_layout(int B, Workspace& w) {
  const size_t tile = B > kSmallBatchRows ? 256 : 128;
  const size_t Bp = ((size_t)B + tile - 1) / tile * tile;
  w.Bp = (int)Bp;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  w.x_face_hi = take(Bp * kFaceK * 2);  w.x_face_lo = take(Bp * kFaceK * 2);
  w.x_audio_hi = take(Bp * kAudioK * 2); w.x_audio_lo = take(Bp * kAudioK * 2);
  w.x_text_hi = take(Bp * kTextK * 2);  w.x_text_lo = take(Bp * kTextK * 2);
  for (int m = 0; m < 3; ++m) { w.h_hi[m] = take(Bp * kHidden * 2); w.h_lo[m] = take(Bp * kHidden * 2); }
  w.cat_hi = take(Bp * 1536 * 2); w.cat_lo = take(Bp * 1536 * 2);
  w.f1_hi = take(Bp * kHidden * 2); w.f1_lo = take(Bp * kHidden * 2);
  w.f32_a = take((size_t)128 * 3072 * 4); w.f32_b = take((size_t)128 * 3072 * 4);
  w.total = off;
}

}  // namespace msa
