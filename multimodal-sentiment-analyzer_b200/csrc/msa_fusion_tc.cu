// Tensor-core path of AdvancedFusionModel.forward: every Linear(+LayerNorm+ReLU) of
// fusion_model.py:44-98 as ONE warp-specialised tcgen05 kernel per layer.
//
//   Y = ReLU(LayerNorm(X W^T + b))          X: [B, K] split-bf16 (hi + lo), W: [N, K] split-bf16
//
// Precision: plain bf16 operands fail the parity bar (99.58 % argmax agreement),
// so each product is issued as three bf16 MMAs with fp32 accumulation in tensor memory:
//   X W^T ~= Xhi Whi^T + Xlo Whi^T + Xhi Wlo^T      (relative error ~2^-16)
//
// Up to kSmallBatchRows rows (streaming, the 1024-segment step), tc_linear_ln_kernel: one CTA owns 128 rows x 128
// output columns in tensor memory (thread <-> TMEM lane <-> row in the epilogue), and the N / 128 CTAs that share a row
// block form a cluster (4 or 8) which exchanges the per-row LayerNorm statistics - and, in the last layer, the partial
// 7 logits - through distributed shared memory, always combined in rank order so every CTA sees the same bits.  A
// layer's latency there is one CTA's MMA chain, hence many short CTAs.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
// Operands are staged by TMA into 128-byte-swizzled K-major tiles.
// Above kSmallBatchRows rows: the persistent CTA-pair kernel of msa_fusion_pair.cuh (tcgen05.mma.cta_group::2).
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "msa_api_internal.h"
#include "msa_fusion_common.cuh"

namespace msa {
namespace cg = cooperative_groups;

constexpr int BLOCK_M = 128, BLOCK_K = 64;
constexpr int kTileA = BLOCK_M * BLOCK_K * 2;            // 16 KB
constexpr int kTcThreads = 256;
constexpr uint32_t kSpinLimit = 400u * 1000u * 1000u;    // a lost barrier traps instead of hanging the GPU

// per-variant shapes: NC columns per CTA, issued in MMAs of NSUB columns
template <int NC> struct TcShape {
  static constexpr int kNSub = NC >= 256 ? 256 : NC;
  static constexpr int kTileW = kNSub * BLOCK_K * 2;                 // 32 KB or 16 KB
  static constexpr int kStageBytes = 2 * kTileA + 2 * kTileW;        // A_hi, A_lo, W_hi, W_lo: 96 KB or 64 KB
  static constexpr int kStages = NC >= 256 ? 2 : 3;
};
constexpr int kMaxStages = 3;

struct TcTail {            // smem after the stages
  uint64_t full[kMaxStages], empty[kMaxStages], accum_full;
  uint32_t tmem_base, pad;
  float2 stats[BLOCK_M];
  float part7[BLOCK_M][8]; // last layer, cluster > 1: this CTA's share of the 7 logits per row
  float bias[512], gamma[512], beta[512];
  float w8[kOut * 512];    // final layer only
  float b8[8];
};
template <int NC> constexpr int tc_smem_bytes() { return TcShape<NC>::kStages * TcShape<NC>::kStageBytes + (int)sizeof(TcTail) + 1024; }

struct TcEpilogue {
  const float* bias;       // [n_total]
  const float* gamma;      // LayerNorm weight [n_total]
  const float* beta;
  __nv_bfloat16* out_hi;   // [rows, ld_out] at col_off
  __nv_bfloat16* out_lo;
  int ld_out, col_off;
  int n_total;             // 512 or 1024
  int m_valid;             // B
  int k_blocks;            // Kpad / 64
  // final layer (kFinal): Linear(512 -> 7) fused after the ReLU
  const float* w8;
  const float* b8;
  float* logits;
  int32_t* argmax;
};

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s2u(bar);
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && ++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s2u(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(s2u(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128-byte swizzle: 8-row groups of 1024 bytes (SBO), one swizzle atom along K (LBO unused)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_ptr) {
  uint64_t d = 0;
  d |= (uint64_t)((s2u(smem_ptr) >> 4) & 0x3FFF);        // start address, bits [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load in two steps, so that the next chunk's load is in flight while the current chunk is stored: the wait
// names the registers as in/out operands, which keeps every use of them behind it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, float* v) {       // 64 consecutive columns: two loads, one wait
  uint32_t r[64];
#pragma unroll
  for (int h = 0; h < 2; ++h)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[32 * h + 0]), "=r"(r[32 * h + 1]), "=r"(r[32 * h + 2]), "=r"(r[32 * h + 3]), "=r"(r[32 * h + 4]), "=r"(r[32 * h + 5]),
          "=r"(r[32 * h + 6]), "=r"(r[32 * h + 7]), "=r"(r[32 * h + 8]), "=r"(r[32 * h + 9]), "=r"(r[32 * h + 10]), "=r"(r[32 * h + 11]),
          "=r"(r[32 * h + 12]), "=r"(r[32 * h + 13]), "=r"(r[32 * h + 14]), "=r"(r[32 * h + 15]), "=r"(r[32 * h + 16]), "=r"(r[32 * h + 17]),
          "=r"(r[32 * h + 18]), "=r"(r[32 * h + 19]), "=r"(r[32 * h + 20]), "=r"(r[32 * h + 21]), "=r"(r[32 * h + 22]), "=r"(r[32 * h + 23]),
          "=r"(r[32 * h + 24]), "=r"(r[32 * h + 25]), "=r"(r[32 * h + 26]), "=r"(r[32 * h + 27]), "=r"(r[32 * h + 28]), "=r"(r[32 * h + 29]),
          "=r"(r[32 * h + 30]), "=r"(r[32 * h + 31])
        : "r"(taddr + 32 * h));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// Up to three independent layers of the same shape class (the per-modality branches of fusion_model.py:44-86)
// share one launch: blockIdx.z selects the layer.
struct TcBatch {
  CUtensorMap maps[3][4];   // A_hi, A_lo, W_hi, W_lo
  TcEpilogue ep[3];
  int trace_slot;           // MSA_TC_TRACE builds: which launch of the forward this is
};

// -DMSA_TC_TRACE (scripts/build_variant.sh): per-CTA clock64 stamps of the pair kernel's phases, read back by
// msa_debug_tc_trace (scripts/fusion_probe.cu).  Not in the default build.
#ifdef MSA_TC_TRACE
constexpr int kTraceSlots = 4, kTraceEvents = 8, kTraceCtas = 4096;
__device__ long long g_tc_trace[kTraceSlots][kTraceEvents][kTraceCtas];
#endif

// NC: output columns per CTA (128 in the product; the shape traits also cover 256 / 512); CL: CTAs per cluster along grid.y = n_total / NC;
// kFinal: fuse Linear(512 -> 7) + argmax behind the ReLU.
template <int NC, int CL, bool kFinal>
__global__ void __launch_bounds__(kTcThreads, 1) tc_linear_ln_kernel(const __grid_constant__ TcBatch batch) {
  using Sh = TcShape<NC>;
  constexpr int N_SUB = Sh::kNSub, kTileW = Sh::kTileW, kStageBytes = Sh::kStageBytes, kStages = Sh::kStages;
  const CUtensorMap& tmA_hi = batch.maps[blockIdx.z][0];
  const CUtensorMap& tmA_lo = batch.maps[blockIdx.z][1];
  const CUtensorMap& tmW_hi = batch.maps[blockIdx.z][2];
  const CUtensorMap& tmW_lo = batch.maps[blockIdx.z][3];
  const TcEpilogue& ep = batch.ep[blockIdx.z];
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcTail* tail = reinterpret_cast<TcTail*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * NC;
  const int k_blocks = ep.k_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    mbar_init(&tail->accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&tail->tmem_base)), "r"(NC));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tail->tmem_base;
  // Programmatic dependent launch: everything above (barriers, tensor-memory allocation) touches no global memory and
  // overlaps the tail of the previous layer; the next layer may be scheduled from here on; nothing below runs before the
  // previous layer's grid has completed and flushed.  (Measured, scripts/ab_check: 102.5 -> 95.3 us per 1024-row
  // forward, profiles/r2_v201_ab_pdl.json.)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ns = 0; ns < NC / N_SUB; ++ns) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          unsigned char* st = smem + stage * kStageBytes;
          mbar_expect_tx(&tail->full[stage], kStageBytes);
          tma_load_2d(&tmA_hi, &tail->full[stage], st, kb * BLOCK_K, m0);
          tma_load_2d(&tmA_lo, &tail->full[stage], st + kTileA, kb * BLOCK_K, m0);
          tma_load_2d(&tmW_hi, &tail->full[stage], st + 2 * kTileA, kb * BLOCK_K, n0 + ns * N_SUB);
          tma_load_2d(&tmW_lo, &tail->full[stage], st + 2 * kTileA + kTileW, kb * BLOCK_K, n0 + ns * N_SUB);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      // kind::f16: D fp32, A/B bf16, both K-major, M = 128, N = N_SUB
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_SUB >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int ns = 0; ns < NC / N_SUB; ++ns) {
        const uint32_t tmem_d = tmem_base + ns * N_SUB;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          unsigned char* st = smem + stage * kStageBytes;
          const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kTileA);
          const uint64_t w_hi = umma_desc_sw128(st + 2 * kTileA), w_lo = umma_desc_sw128(st + 2 * kTileA + kTileW);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);          // 32 bytes per K step inside the swizzle atom
            umma_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
            umma_bf16(tmem_d, a_lo + adv, w_hi + adv, idesc, 1);
            umma_bf16(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
          }
          umma_commit(&tail->empty[stage]);                              // frees the smem stage when the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(&tail->accum_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue pass 1: row statistics of this CTA's columns
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 128;
    for (int i = et; i < NC; i += 128) {
      tail->bias[i] = ep.bias[n0 + i];
      tail->gamma[i] = ep.gamma[n0 + i];
      tail->beta[i] = ep.beta[n0 + i];
    }
    if (kFinal) {
      for (int i = et; i < kOut * NC; i += 128) tail->w8[i] = ep.w8[(i / NC) * ep.n_total + n0 + (i % NC)];
      if (et < kOut) tail->b8[et] = ep.b8[et];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");                       // epilogue warps only
    mbar_wait(&tail->accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // Row statistics of this CTA's NC columns as (mean, centred sum of squares), formed from sums SHIFTED by the row's
    // first value: sum (x - s) and sum (x - s)^2 with s within a few sigma of the mean lose nothing when |mean| >> sigma
    // (a trained checkpoint's biases), unlike sum x^2 / n - mean^2.  One sweep over tensor memory, like before.
    float sum = 0.0f, sumsq = 0.0f, shift = 0.0f;
    for (int c = 0; c < NC; c += 32) {
      float v[32];
      tmem_ld32(trow + c, v);
      if (c == 0) shift = v[0] + tail->bias[0];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = v[i] + tail->bias[c + i] - shift;
        sum += x;
        sumsq = fmaf(x, x, sumsq);
      }
    }
    const float dm = sum * (1.0f / (float)NC);
    tail->stats[row_in_tile] = make_float2(shift + dm, fmaxf(sumsq - sum * dm, 0.0f));     // (mean_i, M2_i)
  }

  cg::cluster_group cluster = cg::this_cluster();
  if (CL > 1) cluster.sync();                                            // every CTA's row statistics are published
  else __syncthreads();
  if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue pass 2: normalise, ReLU, store / project
    const int row_in_tile = (warp & 3) * 32 + lane;
    // combine the CL partial (mean_i, M2_i) of equal size NC (Chan et al.): mean = avg mean_i, M2 = sum M2_i + NC sum (mean_i - mean)^2;
    // rank order, so every CTA of the cluster forms the same bits
    float mi[CL], msum = 0.0f, m2 = 0.0f;
#pragma unroll
    for (int r = 0; r < CL; ++r) {
      const float2* peer = (CL > 1) ? cluster.map_shared_rank(&tail->stats[0], r) : &tail->stats[0];
      const float2 o = peer[row_in_tile];
      mi[r] = o.x;
      msum += o.x;
      m2 += o.y;
    }
    const float mean = msum * (1.0f / (float)CL);
    float spread = 0.0f;
#pragma unroll
    for (int r = 0; r < CL; ++r) spread = fmaf(mi[r] - mean, mi[r] - mean, spread);
    const float var = (m2 + (float)NC * spread) / (float)ep.n_total;
    const float rstd = rsqrtf(var + 1e-5f);
    const uint32_t trow = tail->tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int row = m0 + row_in_tile;
    float acc[kOut];
#pragma unroll
    for (int j = 0; j < kOut; ++j) acc[j] = 0.0f;
    for (int c = 0; c < NC; c += 32) {
      float v[32];
      tmem_ld32(trow + c, v);
      if (kFinal) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y = (v[i] + tail->bias[c + i] - mean) * rstd * tail->gamma[c + i] + tail->beta[c + i];
          y = fmaxf(y, 0.0f);
#pragma unroll
          for (int j = 0; j < kOut; ++j) acc[j] = fmaf(y, tail->w8[j * NC + c + i], acc[j]);
        }
      } else {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float y0 = (v[i] + tail->bias[c + i] - mean) * rstd * tail->gamma[c + i] + tail->beta[c + i];
          float y1 = (v[i + 1] + tail->bias[c + i + 1] - mean) * rstd * tail->gamma[c + i + 1] + tail->beta[c + i + 1];
          y0 = fmaxf(y0, 0.0f);
          y1 = fmaxf(y1, 0.0f);
          const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(y0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(y1 - __bfloat162float(h1));
          hi[i / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[i / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        if (row < ep.m_valid) {
          const size_t o = (size_t)row * ep.ld_out + ep.col_off + n0 + c;
          uint4* ph = reinterpret_cast<uint4*>(ep.out_hi + o);
          uint4* pl = reinterpret_cast<uint4*>(ep.out_lo + o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ph[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            pl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
      }
    }
    if (kFinal) {
      if (CL > 1) {
#pragma unroll
        for (int j = 0; j < kOut; ++j) tail->part7[row_in_tile][j] = acc[j];
      } else if (row < ep.m_valid) {
        int best = 0;
        float bv = 0.0f;
#pragma unroll
        for (int j = 0; j < kOut; ++j) {
          const float l = acc[j] + tail->b8[j];
          ep.logits[(size_t)row * kOut + j] = l;
          if (j == 0 || l > bv) { bv = l; best = j; }
        }
        if (ep.argmax) ep.argmax[row] = best;
      }
    }
  }
  if (kFinal && CL > 1) {
    cluster.sync();                                                      // every CTA's share of the logits is published
    if (warp >= 4 && cluster.block_rank() == 0) {
      const int row_in_tile = (warp & 3) * 32 + lane;
      const int row = m0 + row_in_tile;
      if (row < ep.m_valid) {
        float l[kOut];
#pragma unroll
        for (int j = 0; j < kOut; ++j) l[j] = 0.0f;
#pragma unroll
        for (int r = 0; r < CL; ++r) {
          const float* peer = cluster.map_shared_rank(&tail->part7[0][0], r) + row_in_tile * 8;
#pragma unroll
          for (int j = 0; j < kOut; ++j) l[j] += peer[j];
        }
        int best = 0;
        float bv = 0.0f;
#pragma unroll
        for (int j = 0; j < kOut; ++j) {
          l[j] += tail->b8[j];
          ep.logits[(size_t)row * kOut + j] = l[j];
          if (j == 0 || l[j] > bv) { bv = l[j]; best = j; }
        }
        if (ep.argmax) ep.argmax[row] = best;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CL > 1) cluster.sync();                                            // nobody exits while a peer still reads its smem
  else __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NC));
  }
}

// ------------------------------------------------------------------------------ large batches: CTA pairs
// Above kSmallBatchRows rows a layer runs as tcgen05.mma.cta_group::2 in a persistent kernel (msa_fusion_pair.cuh): two
// CTAs on the two SMs of a TPC own 256 rows x 512 columns; each stages ITS 128 rows of the activations and HALF of the
// 256 weight rows of an MMA (hi + lo: 64 KB per stage, three stages) and the leader issues M = 256, N = 256 MMAs that
// read both CTAs' shared memory and write both CTAs' tensor memory.  Against one CTA per 128 x 512 tile this halves the
// weight bytes every SM pulls from L2 and cuts the shared-memory traffic per MMA k-step (operand reads + TMA writes)
// from 60 KB to 40 KB per 384 tensor-pipe cycles - the single-CTA form asks for 156 B/clk of a 128 B/clk shared memory.
constexpr int kPairTileB = 128 * BLOCK_K * 2;                            // this CTA's half of a 256-row weight tile: 16 KB
constexpr int kPairStageBytes = 2 * kTileA + 2 * kPairTileB;             // 64 KB
constexpr int kPairStages = 3;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA load into THIS CTA's shared memory whose bytes complete on a barrier that may live in the pair's other CTA
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s2u(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s2u(bar)),
               "h"(cta_mask)
               : "memory");
}

}  // namespace msa
#include "msa_fusion_pair.cuh"
namespace msa {

// ------------------------------------------------------------------------------ input LayerNorm + split
// One warp per row and modality: y = LN(x) * gamma + beta -> (hi, lo) bf16, K zero-padded; rows >= B zeroed.
struct PrepArgs {
  const float* x[3];
  const float* gamma[3];
  const float* beta[3];
  __nv_bfloat16* hi[3];
  __nv_bfloat16* lo[3];
  int d[3], kpad[3];
  int B, Bp, nmod;
  int lane_rows;            // short rows (<= 32 columns): one lane per row (large batches) instead of one warp per row
};

constexpr int kPrepMaxK = 832;                                           // 783 padded to a multiple of 64

// One row of one modality, NI = ceil(d / 32) column groups in registers (all loads of the row in flight at once)
template <int NI>
__device__ __forceinline__ void prep_row(const float* x, const float* gamma, const float* beta, int d, int kpad, int lane,
                                         __nv_bfloat16* sh, __nv_bfloat16* sl) {
  float v[NI];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < d) ? x[c] : 0.0f;
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)d;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane + 32 * i;
    const float dd = (c < d) ? v[i] - mean : 0.0f;
    sq = fmaf(dd, dd, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < NI + 1; ++i) {                                     // kpad <= 32 (NI + 1)
    const int c = lane + 32 * i;
    if (c < kpad) {
      float y = 0.0f;
      if (c < d) y = (v[i < NI ? i : NI - 1] - mean) * rstd * gamma[c] + beta[c];
      const __nv_bfloat16 h = __float2bfloat16_rn(y);
      sh[c] = h;
      sl[c] = __float2bfloat16_rn(y - __bfloat162float(h));
    }
  }
}

// Rows of at most 32 columns (face 27, audio 31): one LANE per row, the whole row in registers, no shuffles; every lane
// writes its row's 64 (hi) + 64 (lo) zero-padded bf16 as whole 128-byte lines.  (One warp per such row made 16,384
// eight-warp CTAs of a few instructions each: half of this kernel's 189 us at 65,536 rows.)
__device__ __forceinline__ void prep_short_row(const float* x, const float* gamma, const float* beta, int d, uint4* hi, uint4* lo) {
  float v[32];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = (i < d) ? x[i] : 0.0f;
    sum += v[i];
  }
  const float mean = sum / (float)d;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float dd = (i < d) ? v[i] - mean : 0.0f;
    sq = fmaf(dd, dd, sq);
  }
  const float rstd = rsqrtf(sq / (float)d + 1e-5f);
  uint32_t h[16], l[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float y0 = (i < d) ? (v[i] - mean) * rstd * __ldg(gamma + i) + __ldg(beta + i) : 0.0f;
    const float y1 = (i + 1 < d) ? (v[i + 1] - mean) * rstd * __ldg(gamma + i + 1) + __ldg(beta + i + 1) : 0.0f;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(y0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(y1 - __bfloat162float(h1));
    h[i / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    lo[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
    hi[4 + j] = make_uint4(0u, 0u, 0u, 0u);                               // columns 32..63: K padding
    lo[4 + j] = make_uint4(0u, 0u, 0u, 0u);
  }
}

static_assert(kFaceK == 64 && kAudioK == 64, "prep_short_row writes 64 padded columns");

__global__ void __launch_bounds__(256, 4) tc_input_prep_kernel(const PrepArgs a) {
  // a long row (text, 783 columns) is normalised by one warp with lane-strided columns (coalesced loads), staged as bf16
  // pairs in shared memory and written out 16 bytes per lane: the 2-byte lane-strided stores of the first version ran at
  // a quarter of the HBM rate.
  __shared__ __align__(16) __nv_bfloat16 stage[8][2][kPrepMaxK];
  // the projection layer behind this kernel is a programmatic dependent launch: for small batches its barrier set-up
  // and tensor-memory allocation may start now (it waits for this grid before it loads a row): 95.4 -> 94.5 us per
  // 1024-row forward.  Not for large batches: the persistent projection CTAs take a whole SM's shared memory and would
  // keep this kernel's later CTAs off the SMs (4096 rows: 216 -> 224 us when tried).
  if (!a.lane_rows) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int m = blockIdx.y;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = a.d[m], kpad = a.kpad[m];
  if (d <= 32 && a.lane_rows) {                                           // kpad == 64
    const int row = (blockIdx.x * 8 + w) * 32 + lane;
    if (row >= a.Bp) return;
    uint4* hi = reinterpret_cast<uint4*>(a.hi[m] + (size_t)row * kpad);
    uint4* lo = reinterpret_cast<uint4*>(a.lo[m] + (size_t)row * kpad);
    if (row >= a.B) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { hi[j] = make_uint4(0u, 0u, 0u, 0u); lo[j] = make_uint4(0u, 0u, 0u, 0u); }
      return;
    }
    prep_short_row(a.x[m] + (size_t)row * d, a.gamma[m], a.beta[m], d, hi, lo);
    return;
  }
  const int row = blockIdx.x * 8 + w;
  if (row >= a.Bp) return;
  uint4* hi = reinterpret_cast<uint4*>(a.hi[m] + (size_t)row * kpad);
  uint4* lo = reinterpret_cast<uint4*>(a.lo[m] + (size_t)row * kpad);
  if (row >= a.B) {
    for (int c = lane; c < kpad / 8; c += 32) { hi[c] = make_uint4(0u, 0u, 0u, 0u); lo[c] = make_uint4(0u, 0u, 0u, 0u); }
    return;
  }
  if (d <= 32) prep_row<1>(a.x[m] + (size_t)row * d, a.gamma[m], a.beta[m], d, kpad, lane, stage[w][0], stage[w][1]);   // small batch: latency, not throughput
  else prep_row<25>(a.x[m] + (size_t)row * d, a.gamma[m], a.beta[m], d, kpad, lane, stage[w][0], stage[w][1]);         // 783 <= 25 * 32
  __syncwarp();
  const uint4* sh = reinterpret_cast<const uint4*>(stage[w][0]);
  const uint4* sl = reinterpret_cast<const uint4*>(stage[w][1]);
  for (int c = lane; c < kpad / 8; c += 32) { hi[c] = sh[c]; lo[c] = sl[c]; }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] -> TMA map with box [box_rows, 64 cols], 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MSA_ERR_BAD_ARGUMENT;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSA_OK : 1000 + (int)r;
}

struct LayerLaunch {
  const void *a_hi, *a_lo; int a_cols;           // activations [Bp, a_cols]
  int gemm;                                      // weight id
  int bias_t, gamma_t, beta_t;                   // tensor ids
  void *out_hi, *out_lo; int ld_out, col_off;
  bool final_layer;
};

template <int NC, int CL, bool kFinal>
static cudaError_t launch_variant(const TcBatch& batch, int Bp, int N, int count, cudaStream_t s) {
  constexpr int smem = tc_smem_bytes<NC>();
  cudaError_t e = cudaFuncSetAttribute(tc_linear_ln_kernel<NC, CL, kFinal>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(Bp / BLOCK_M, N / NC, count);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CL;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchAttribute attr2[2];
  attr2[0] = attr[0];
  attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr2[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr2;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, tc_linear_ln_kernel<NC, CL, kFinal>, batch);
}

template <int CLN, bool kFinal>
static cudaError_t launch_pair_persistent(const TcBatch& batch, int Bp, int N, int count, cudaStream_t s) {
  auto kern = tc_pair_persistent_kernel<CLN, kFinal>;
  constexpr int smem = pp_smem_bytes<kFinal>();
  if (Bp % (2 * BLOCK_M) != 0 || N != 512 * CLN) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  PpSched sched{Bp / (2 * BLOCK_M), count};
  const int total = sched.count * sched.tiles_per_branch;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kPpThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = CLN;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // clusters that fit the device at once (asked once per device and variant): the launch is exactly that many
  static int fit[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (fit[dev] == 0) {
    cfg.gridDim = dim3(2 * 148, CLN, 1);
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    fit[dev] = n;
  }
  const int G = total < fit[dev] ? total : fit[dev];
  cfg.gridDim = dim3(2 * G, CLN, 1);
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, batch, sched);
}

// One launch for `count` (1..3) layers of the same output width (all 1024-wide or all 512-wide).
static int launch_layers(const LayerLaunch* Ls, int count, int B, int Bp, const unsigned char* packed, const PackedHeader& h,
                         float* logits, int32_t* argmax, cudaStream_t s) {
  TcBatch batch;
  std::memset(&batch, 0, sizeof(batch));
  const int N = kGemmWeights[Ls[0].gemm].N;
  const bool small = B <= kSmallBatchRows;                       // 128 columns per CTA: more, shorter CTAs
  const int nsub = 128;                                          // weight rows per TMA box: one CTA's tile, or its half of a pair's
  for (int i = 0; i < count; ++i) {
    const LayerLaunch& L = Ls[i];
    const GemmWeight& gw = kGemmWeights[L.gemm];
    if (gw.N != N || L.final_layer != Ls[0].final_layer) return MSA_ERR_BAD_ARGUMENT;
    int rc;
    if ((rc = make_map(&batch.maps[i][0], L.a_hi, Bp, L.a_cols, BLOCK_M))) return rc;
    if ((rc = make_map(&batch.maps[i][1], L.a_lo, Bp, L.a_cols, BLOCK_M))) return rc;
    if ((rc = make_map(&batch.maps[i][2], packed + h.hi_off[L.gemm], gw.N, gw.Kpad, nsub))) return rc;
    if ((rc = make_map(&batch.maps[i][3], packed + h.lo_off[L.gemm], gw.N, gw.Kpad, nsub))) return rc;
    TcEpilogue& ep = batch.ep[i];
    ep.bias = reinterpret_cast<const float*>(packed + h.f32_off[L.bias_t]);
    ep.gamma = reinterpret_cast<const float*>(packed + h.f32_off[L.gamma_t]);
    ep.beta = reinterpret_cast<const float*>(packed + h.f32_off[L.beta_t]);
    ep.out_hi = static_cast<__nv_bfloat16*>(L.out_hi);
    ep.out_lo = static_cast<__nv_bfloat16*>(L.out_lo);
    ep.ld_out = L.ld_out;
    ep.col_off = L.col_off;
    ep.n_total = gw.N;
    ep.m_valid = B;
    ep.k_blocks = gw.Kpad / BLOCK_K;
    ep.w8 = reinterpret_cast<const float*>(packed + h.f32_off[T_FUS8_W]);
    ep.b8 = reinterpret_cast<const float*>(packed + h.f32_off[T_FUS8_B]);
    ep.logits = logits;
    ep.argmax = argmax;
  }
  const bool fin = Ls[0].final_layer;
  batch.trace_slot = fin ? 3 : (N == 1024 ? (count > 1 ? 0 : 2) : 1);    // projections, processors, fusion.0, fusion.4
  cudaError_t e;
  if (N == 1024) {
    if (fin) return MSA_ERR_BAD_ARGUMENT;
    e = small ? launch_variant<128, 8, false>(batch, Bp, N, count, s) : launch_pair_persistent<2, false>(batch, Bp, N, count, s);
  } else if (N == 512) {
    if (fin) e = small ? launch_variant<128, 4, true>(batch, Bp, N, count, s) : launch_pair_persistent<1, true>(batch, Bp, N, count, s);
    else e = small ? launch_variant<128, 4, false>(batch, Bp, N, count, s) : launch_pair_persistent<1, false>(batch, Bp, N, count, s);
  } else {
    return MSA_ERR_BAD_ARGUMENT;
  }
  if (e != cudaSuccess) return (int)e;
  note_launches(1);
  return MSA_OK;
}

int fusion_forward_tc(const float* face, const float* audio, const float* text, int B, const unsigned char* packed,
                      const PackedHeader& h, unsigned char* ws, const Workspace& wl, float* logits7, int32_t* argmax,
                      cudaStream_t s) {
  const bool three = text != nullptr;
  const int Bp = wl.Bp;
  auto f32 = [&](int t) { return reinterpret_cast<const float*>(packed + h.f32_off[t]); };
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };

  PrepArgs pa{};
  pa.x[0] = face; pa.x[1] = audio; pa.x[2] = text;
  pa.gamma[0] = f32(T_FACE_NORM_W); pa.beta[0] = f32(T_FACE_NORM_B);
  pa.gamma[1] = f32(T_AUDIO_NORM_W); pa.beta[1] = f32(T_AUDIO_NORM_B);
  pa.gamma[2] = f32(T_TEXT_NORM_W); pa.beta[2] = f32(T_TEXT_NORM_B);
  pa.hi[0] = bf(wl.x_face_hi); pa.lo[0] = bf(wl.x_face_lo);
  pa.hi[1] = bf(wl.x_audio_hi); pa.lo[1] = bf(wl.x_audio_lo);
  pa.hi[2] = bf(wl.x_text_hi); pa.lo[2] = bf(wl.x_text_lo);
  pa.d[0] = kFaceDim; pa.d[1] = kAudioDim; pa.d[2] = kTextDim;
  pa.kpad[0] = kFaceK; pa.kpad[1] = kAudioK; pa.kpad[2] = kTextK;
  pa.B = B; pa.Bp = Bp; pa.nmod = three ? 3 : 2;
  pa.lane_rows = B > kSmallBatchRows;
  tc_input_prep_kernel<<<dim3((Bp + 7) / 8, pa.nmod), 256, 0, s>>>(pa);
  note_launches(1);

  const int cat_w = three ? 1536 : 1024;
  const size_t xin_hi[3] = {wl.x_face_hi, wl.x_audio_hi, wl.x_text_hi}, xin_lo[3] = {wl.x_face_lo, wl.x_audio_lo, wl.x_text_lo};
  const int xk[3] = {kFaceK, kAudioK, kTextK};
  const int proj_g[3] = {G_FACE_PROJ, G_AUDIO_PROJ, G_TEXT_PROJ}, proj_b[3] = {T_FACE_PROJ_B, T_AUDIO_PROJ_B, T_TEXT_PROJ_B};
  const int l0w[3] = {T_FACE_P0_W, T_AUDIO_P0_W, T_TEXT_P0_W}, l0b[3] = {T_FACE_P0_B, T_AUDIO_P0_B, T_TEXT_P0_B};
  const int p3_g[3] = {G_FACE_P3, G_AUDIO_P3, G_TEXT_P3}, p3_b[3] = {T_FACE_P3_B, T_AUDIO_P3_B, T_TEXT_P3_B};
  const int l4w[3] = {T_FACE_P4_W, T_AUDIO_P4_W, T_TEXT_P4_W}, l4b[3] = {T_FACE_P4_B, T_AUDIO_P4_B, T_TEXT_P4_B};
  int rc;
  // the modality branches are independent: all projections in one launch, all processors in the next
  LayerLaunch proj[3], proc[3];
  for (int m = 0; m < pa.nmod; ++m) {
    proj[m] = LayerLaunch{bf(xin_hi[m]), bf(xin_lo[m]), xk[m], proj_g[m], proj_b[m], l0w[m], l0b[m], bf(wl.h_hi[m]), bf(wl.h_lo[m]), kHidden, 0, false};
    proc[m] = LayerLaunch{bf(wl.h_hi[m]), bf(wl.h_lo[m]), kHidden, p3_g[m], p3_b[m], l4w[m], l4b[m], bf(wl.cat_hi), bf(wl.cat_lo), cat_w, m * kHalf, false};
  }
  if ((rc = launch_layers(proj, pa.nmod, B, Bp, packed, h, nullptr, nullptr, s))) return rc;
  if ((rc = launch_layers(proc, pa.nmod, B, Bp, packed, h, nullptr, nullptr, s))) return rc;
  LayerLaunch f0{bf(wl.cat_hi), bf(wl.cat_lo), cat_w, three ? G_FUS0 : G_FUS2, three ? T_FUS0_B : T_FUS2_B, T_FUS1_W, T_FUS1_B,
                 bf(wl.f1_hi), bf(wl.f1_lo), kHidden, 0, false};
  if ((rc = launch_layers(&f0, 1, B, Bp, packed, h, nullptr, nullptr, s))) return rc;
  LayerLaunch f4{bf(wl.f1_hi), bf(wl.f1_lo), kHidden, G_FUS4, T_FUS4_B, T_FUS5_W, T_FUS5_B, nullptr, nullptr, 0, 0, true};
  if ((rc = launch_layers(&f4, 1, B, Bp, packed, h, logits7, argmax, s))) return rc;
  return (int)cudaGetLastError();
}

}  // namespace msa

#ifdef MSA_TC_TRACE
// out[slot][event][cta] clock64 stamps of the last forward (MSA_TC_TRACE builds only)
extern "C" int msa_debug_tc_trace(long long* out, size_t bytes) {
  if (bytes < sizeof(msa::g_tc_trace)) return -1;
  return (int)cudaMemcpyFromSymbol(out, msa::g_tc_trace, sizeof(msa::g_tc_trace));
}
#endif
