// Persistent CTA-pair layer kernel of the fusion chain (batches above kSmallBatchRows rows).  Included by
// msa_fusion_tc.cu, which holds the barrier / TMA / tcgen05 helpers and the host side.
//
//   Y = ReLU(LayerNorm(X W^T + b)),  X and W split-bf16 (three MMAs per product), as in tc_linear_ln_kernel.
//
// One cluster = 2 x CLN CTAs (CLN = N / 512): the pair along x shares every MMA (tcgen05.mma.cta_group::2, M = 256 rows
// over the two CTAs, N = 256 columns per instruction), the CLN column blocks along y only exchange LayerNorm partials.
// Each CTA keeps 128 rows x 512 columns of fp32 accumulators = all of its tensor memory, as two halves of 256 columns.
// A cluster walks over row blocks (and, where a launch carries several branches, over the branches: heaviest K first):
// the launch is sized to the clusters that fit the device at once.
//
// Roles per CTA (384 threads): warp 0 = TMA producer (one lane, both CTAs: each loads ITS 128 activation rows and ITS
// half of the 256 weight rows of an MMA: 64 KB per stage, three stages, bytes complete on the LEADER's full barrier),
// warp 1 = MMA issuer (one lane of the leader; tcgen05.commit multicasts to both CTAs), warp 2 = tensor-memory
// allocator, warps 4-11 = epilogue (two warps per 32-lane quadrant: each thread owns one row and 128 of the 256
// columns of a half).
//
// Overlap: the MMAs of a tile run half by half (columns 0-255, then 256-511).  The epilogue sweeps half 0 for its row
// statistics while the MMAs of half 1 run; when half 1 is complete it sweeps that, combines the partials (Chan), and
// normalises / stores half 0, hands half 0 back to the MMA warp (tmem_empty), which starts the NEXT tile's half 0
// while the epilogue normalises half 1.  The producer runs ahead across tiles, so a tile never waits for its first
// operands.  Per tile only (statistics of half 1) + (normalise half 0) is not hidden behind MMAs.
//
// Stores: a thread owns a ROW, so direct stores would touch 32 different 128-byte lines per instruction (measured:
// 31,000 cycles per 128 x 512 tile against 51,000 of MMA).  Each warp passes 32 rows x 32 columns of hi and of lo through
// a swizzled staging tile of its own and stores them with four lanes per 64-byte row segment.  (Tensor stores from a
// single staging tile per warp were measured slower here: every chunk waited ~1,900 cycles for the previous store's
// read-out behind the operand loads on the same TMA unit, and there is no shared memory left for deeper staging.)
#pragma once

namespace msa {

constexpr int kPpThreads = 384;
constexpr int kPpStoreTile = 32 * 32 * 2;                                // 32 rows x 32 bf16 columns: 2 KB

struct PpBars {
  uint64_t full[kPairStages], empty[kPairStages];
  uint64_t accum_full[2];      // half h of the accumulators is complete (both CTAs, by multicast commit)
  uint64_t tmem_empty[2];      // leader only: both CTAs' epilogues have read half h out
  uint64_t stats_full;         // CLN > 1: every column block of this row block has published its partials
  uint64_t stats_read;         // CLN > 1: the other column blocks have read this CTA's partials (8 warps each)
  uint32_t tmem_base, pad;
};
template <bool kFinal> struct alignas(1024) PpTail {
  unsigned char store[kFinal ? 16 : 8 * kPpStoreTile];                  // non-final: one staging tile per epilogue warp (first: 1024-byte aligned)
  alignas(16) float cst[3][512];                                         // bias, LayerNorm weight and bias of this CTA's 512 columns (current branch)
  PpBars b;
  float2 stats[2][BLOCK_M];                                              // [column part of the half][row]: (mean, M2) over 256 columns
  float part7[kFinal ? 2 : 1][kFinal ? BLOCK_M : 1][8];                  // final: the two column parts of a row meet here
  alignas(16) float w8[kFinal ? kOut * 512 : 4];
  float b8[8];
};
template <bool kFinal> constexpr int pp_smem_bytes() { return kPairStages * kPairStageBytes + (int)sizeof(PpTail<kFinal>); }

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope
  const uint32_t addr = s2u(bar);
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && ++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

struct PpSched {
  int tiles_per_branch;      // Bp / 256
  int count;                 // branches in this launch (1..3)
};

#ifdef MSA_TC_TRACE
#define PP_T0() long long pp_t_ = clock64()
#define PP_ACC(var) do { const long long n_ = clock64(); var += n_ - pp_t_; pp_t_ = n_; } while (0)
#define PP_DUMP(ev, val)                                                                                           \
  do {                                                                                                             \
    const unsigned cta_ = blockIdx.x + gridDim.x * blockIdx.y;                                                     \
    if (cta_ < (unsigned)kTraceCtas && batch.trace_slot < kTraceSlots) g_tc_trace[batch.trace_slot][ev][cta_] = (val); \
  } while (0)
#else
#define PP_T0() do { } while (0)
#define PP_ACC(var) do { } while (0)
#define PP_DUMP(ev, val) do { } while (0)
#endif

template <int CLN, bool kFinal>
__global__ void __launch_bounds__(kPpThreads, 1) tc_pair_persistent_kernel(const __grid_constant__ TcBatch batch, const PpSched sched) {
  constexpr int NC = 512, N_SUB = 256;
  extern __shared__ __align__(1024) unsigned char smem[];
  PpTail<kFinal>* tail = reinterpret_cast<PpTail<kFinal>*>(smem + kPairStages * kPairStageBytes);
  PpBars* B = &tail->b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();                              // x + 2 y
  const uint32_t px = crank & 1u, py = crank >> 1;
  const bool leader = px == 0;
  const int G = gridDim.x >> 1, g = blockIdx.x >> 1;                     // clusters in flight, this cluster
  const int total = sched.count * sched.tiles_per_branch;
  const int n0 = (int)py * NC;
  cg::cluster_group cluster = cg::this_cluster();
  if ((s2u(smem) & 1023u) != 0) __trap();                                // the swizzled tiles rely on it

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) { mbar_init(&B->full[s], 1); mbar_init(&B->empty[s], 1); }
    for (int h = 0; h < 2; ++h) { mbar_init(&B->accum_full[h], 1); mbar_init(&B->tmem_empty[h], 2); }
    mbar_init(&B->stats_full, CLN);
    mbar_init(&B->stats_read, 8 * (CLN > 1 ? CLN - 1 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&B->tmem_base)), "r"(NC));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = B->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // tile t of this launch -> (branch, row block); the text branch (K = 832 against 64) goes first
  auto branch_of = [&](int t) { const int zi = t / sched.tiles_per_branch; return sched.count == 3 ? 2 - zi : zi; };
  auto rows_of = [&](int t) { return (t % sched.tiles_per_branch) * (2 * BLOCK_M) + (int)px * BLOCK_M; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane of EACH CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = g; t < total; t += G) {
        const int z = branch_of(t), m0 = rows_of(t), k_blocks = batch.ep[z].k_blocks;
        const CUtensorMap* mp = batch.maps[z];
        for (int ns = 0; ns < NC / N_SUB; ++ns) {
          const int wrow = n0 + ns * N_SUB + (int)px * (N_SUB / 2);
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&B->empty[stage], phase ^ 1);
            unsigned char* st = smem + stage * kPairStageBytes;
            if (leader) mbar_expect_tx(&B->full[stage], 2 * kPairStageBytes);
            const uint32_t bar = mapa_shared(s2u(&B->full[stage]), crank & ~1u);
            tma_load_2d_pair(&mp[0], bar, st, kb * BLOCK_K, m0);
            tma_load_2d_pair(&mp[1], bar, st + kTileA, kb * BLOCK_K, m0);
            tma_load_2d_pair(&mp[2], bar, st + 2 * kTileA, kb * BLOCK_K, wrow);
            tma_load_2d_pair(&mp[3], bar, st + 2 * kTileA + kPairTileB, kb * BLOCK_K, wrow);
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one lane of the leader)
    if (lane == 0 && leader) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_SUB >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
      const uint16_t pair_mask = (uint16_t)(3u << (2 * py));
      int stage = 0;
      uint32_t phase = 0, it = 0;
      [[maybe_unused]] long long w_tmem = 0, w_full = 0, w_issue = 0;
      PP_T0();
      for (int t = g; t < total; t += G, ++it) {
        const int k_blocks = batch.ep[branch_of(t)].k_blocks;
        for (int ns = 0; ns < NC / N_SUB; ++ns) {
          mbar_wait_cluster(&B->tmem_empty[ns], (it & 1u) ^ 1u);         // the previous tile's half has been read out by both CTAs
          PP_ACC(w_tmem);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + ns * N_SUB;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&B->full[stage], phase);
            PP_ACC(w_full);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned char* st = smem + stage * kPairStageBytes;
            const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kTileA);
            const uint64_t w_hi = umma_desc_sw128(st + 2 * kTileA), w_lo = umma_desc_sw128(st + 2 * kTileA + kPairTileB);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              umma_bf16_pair(tmem_d, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
              umma_bf16_pair(tmem_d, a_lo + adv, w_hi + adv, idesc, 1);
              umma_bf16_pair(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
            }
            umma_commit_pair(&B->empty[stage], pair_mask);
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
            PP_ACC(w_issue);
          }
          umma_commit_pair(&B->accum_full[ns], pair_mask);
        }
      }
      PP_DUMP(0, w_tmem); PP_DUMP(1, w_full); PP_DUMP(2, w_issue);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (256 threads)
    const int et = threadIdx.x - 128;
    const int q = warp & 3, sub = (warp - 4) >> 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* stg = tail->store + (kFinal ? 0 : (warp - 4) * kPpStoreTile);
    const uint32_t tmem_empty_leader[2] = {mapa_shared(s2u(&B->tmem_empty[0]), crank & ~1u), mapa_shared(s2u(&B->tmem_empty[1]), crank & ~1u)};
    if (kFinal) {
      const TcEpilogue& e0 = batch.ep[0];
      for (int i = et; i < kOut * NC; i += 256) tail->w8[i] = e0.w8[(i / NC) * e0.n_total + n0 + (i % NC)];
      if (et < kOut) tail->b8[et] = e0.b8[et];
      epi_bar(1);
    }
    uint32_t it = 0;
    int z_staged = -1;
    [[maybe_unused]] long long e_wait = 0, e_p1 = 0, e_sync = 0, e_p2 = 0, e_rel = 0, e_ld = 0;
    PP_T0();
    for (int t = g; t < total; t += G, ++it) {
      const int z = branch_of(t), m0 = rows_of(t);
      const TcEpilogue& ep = batch.ep[z];
      if (z != z_staged) {
        // every thread has left the previous tile's last read of the constants (it passed that tile's last barrier)
        for (int i = et; i < NC; i += 256) {
          tail->cst[0][i] = __ldg(ep.bias + n0 + i);
          tail->cst[1][i] = __ldg(ep.gamma + n0 + i);
          tail->cst[2][i] = __ldg(ep.beta + n0 + i);
        }
        z_staged = z;
        epi_bar(3);
      }
      const float4* bias4 = reinterpret_cast<const float4*>(tail->cst[0]);
      const float4* gamma4 = reinterpret_cast<const float4*>(tail->cst[1]);
      const float4* beta4 = reinterpret_cast<const float4*>(tail->cst[2]);
      // ---- statistics: (mean, M2) of this thread's 128 columns of each half, shifted by the first value
      float hm[2], hM2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&B->accum_full[h], it & 1u);
        PP_ACC(e_wait);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int c0 = h * N_SUB + sub * 128;
        float2 sum2 = make_float2(0.0f, 0.0f), sq2 = make_float2(0.0f, 0.0f);
        float shift = 0.0f;
        for (int c = c0; c < c0 + 128; c += 64) {
          float v[64];
          PP_ACC(e_p1);
          tmem_ld32x2(trow + c, v);                                       // two loads in flight, one wait
          PP_ACC(e_ld);
          if (c == c0) shift = v[0] + tail->cst[0][c0];
          const float2 ns2 = make_float2(-shift, -shift);
#pragma unroll
          for (int i = 0; i < 64; i += 4) {                               // packed fp32: two columns per instruction
            const float4 b4 = bias4[(c + i) / 4];
            const float2 xa = __fadd2_rn(__fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y)), ns2);
            const float2 xb = __fadd2_rn(__fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w)), ns2);
            sum2 = __fadd2_rn(sum2, xa);
            sq2 = __ffma2_rn(xa, xa, sq2);
            sum2 = __fadd2_rn(sum2, xb);
            sq2 = __ffma2_rn(xb, xb, sq2);
          }
        }
        const float sum = sum2.x + sum2.y, sumsq = sq2.x + sq2.y;
        const float dm = sum * (1.0f / 128.0f);
        hm[h] = shift + dm;
        hM2[h] = fmaxf(sumsq - sum * dm, 0.0f);
        PP_ACC(e_p1);
      }
      {
        // the two halves' partials of this thread -> one partial over its 256 columns (Chan, equal sizes)
        const float m = 0.5f * (hm[0] + hm[1]);
        const float d0 = hm[0] - m, d1 = hm[1] - m;
        if (CLN > 1) mbar_wait_cluster(&B->stats_read, (it & 1u) ^ 1u);  // the other column blocks are done with the previous tile's partials
        tail->stats[sub][row_in_tile] = make_float2(m, hM2[0] + hM2[1] + 128.0f * (d0 * d0 + d1 * d1));
      }
      epi_bar(1);
      if (CLN > 1) {
        if (et == 0) {
#pragma unroll
          for (int r = 0; r < CLN; ++r) mbar_arrive_remote(mapa_shared(s2u(&B->stats_full), px + 2u * (uint32_t)r));
        }
        mbar_wait_cluster(&B->stats_full, it & 1u);
      }
      constexpr int NP = 2 * CLN;                                        // partials per row, 256 columns each, fixed order
      float mi[NP], msum = 0.0f, m2 = 0.0f;
#pragma unroll
      for (int r = 0; r < CLN; ++r) {
        const float2* peer = (CLN > 1) ? cluster.map_shared_rank(&tail->stats[0][0], (int)px + 2 * r) : &tail->stats[0][0];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float2 o = peer[s * BLOCK_M + row_in_tile];
          mi[2 * r + s] = o.x;
          msum += o.x;
          m2 += o.y;
        }
      }
      if (CLN > 1) {
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < CLN; ++r)
            if (r != (int)py) mbar_arrive_remote(mapa_shared(s2u(&B->stats_read), px + 2u * (uint32_t)r));
        }
      }
      PP_ACC(e_sync);
      const float mean = msum * (1.0f / (float)NP);
      float spread = 0.0f;
#pragma unroll
      for (int r = 0; r < NP; ++r) spread = fmaf(mi[r] - mean, mi[r] - mean, spread);
      const float var = (m2 + 256.0f * spread) / (float)ep.n_total;
      const float rstd = rsqrtf(var + 1e-5f);
      const int row = m0 + row_in_tile;
      float2 acc2[kOut];                                                 // final layer: the 7 logits, even and odd columns apart
#pragma unroll
      for (int j = 0; j < kOut; ++j) acc2[j] = make_float2(0.0f, 0.0f);
      // ---- normalise, ReLU, store (or project): half 0 first, so that the next tile's MMAs can start
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int c0 = h * N_SUB + sub * 128;
        uint32_t rv[32];
        tmem_ld32_issue(trow + c0, rv);
#pragma unroll 1
        for (int c = c0; c < c0 + 128; c += 32) {
          float v[32];
          PP_ACC(e_p2);
          tmem_ld32_wait(rv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rv[i]);
          PP_ACC(e_ld);
          const float2 nm2 = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);
          // ((v + bias - mean) * rstd) * gamma + beta, two columns per packed instruction
          auto norm2 = [&](float va, float vb, float ba, float bb, float ga, float gb, float ta, float tb) {
            float2 y = __fadd2_rn(__fadd2_rn(make_float2(va, vb), make_float2(ba, bb)), nm2);
            y = __ffma2_rn(__fmul2_rn(y, rs2), make_float2(ga, gb), make_float2(ta, tb));
            return make_float2(fmaxf(y.x, 0.0f), fmaxf(y.y, 0.0f));
          };
          if (kFinal) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = bias4[(c + i) / 4], g4 = gamma4[(c + i) / 4], t4 = beta4[(c + i) / 4];
              const float2 ya = norm2(v[i], v[i + 1], b4.x, b4.y, g4.x, g4.y, t4.x, t4.y);
              const float2 yb = norm2(v[i + 2], v[i + 3], b4.z, b4.w, g4.z, g4.w, t4.z, t4.w);
#pragma unroll
              for (int j = 0; j < kOut; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(&tail->w8[j * NC + c + i]);
                acc2[j] = __ffma2_rn(ya, make_float2(w4.x, w4.y), acc2[j]);
                acc2[j] = __ffma2_rn(yb, make_float2(w4.z, w4.w), acc2[j]);
              }
            }
            if (c + 32 < c0 + 128) tmem_ld32_issue(trow + c + 32, rv);
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = bias4[(c + i) / 4], g4 = gamma4[(c + i) / 4], t4 = beta4[(c + i) / 4];
              const float2 ya = norm2(v[i], v[i + 1], b4.x, b4.y, g4.x, g4.y, t4.x, t4.y);
              const float2 yb = norm2(v[i + 2], v[i + 3], b4.z, b4.w, g4.z, g4.w, t4.z, t4.w);
              const __nv_bfloat162 ha = __float22bfloat162_rn(ya), hb = __float22bfloat162_rn(yb);
              const float2 fa = __bfloat1622float2(ha), fb = __bfloat1622float2(hb);
              const __nv_bfloat162 la = __float22bfloat162_rn(__fadd2_rn(ya, make_float2(-fa.x, -fa.y)));
              const __nv_bfloat162 lb = __float22bfloat162_rn(__fadd2_rn(yb, make_float2(-fb.x, -fb.y)));
              hi[i / 2] = *reinterpret_cast<const uint32_t*>(&ha);
              hi[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&hb);
              lo[i / 2] = *reinterpret_cast<const uint32_t*>(&la);
              lo[i / 2 + 1] = *reinterpret_cast<const uint32_t*>(&lb);
            }
            if (c + 32 < c0 + 128) tmem_ld32_issue(trow + c + 32, rv);   // in flight while this chunk is stored
            // through the warp's staging tile (64-byte rows, chunk ^ ((row >> 1) & 3): conflict-free both ways) to global
            // memory, hi then lo: four lanes cover one row's 64 bytes, eight rows per instruction
            const int ch = lane & 3;
            const size_t col = (size_t)(ep.col_off + n0 + c + ch * 8);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              const uint32_t* src = part ? lo : hi;
              __nv_bfloat16* dst = part ? ep.out_lo : ep.out_hi;
              __syncwarp();                                               // the tile's previous content has been read out
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int r = 8 * k + (lane >> 2);
                *reinterpret_cast<uint4*>(dst + (size_t)(m0 + q * 32 + r) * ep.ld_out + col) =
                    *reinterpret_cast<const uint4*>(stg + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
              }
            }
          }
        }
        // this half of tensor memory is free for the next tile's MMAs once both CTAs' epilogues have left it
        PP_ACC(e_p2);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        epi_bar(2);
        if (et == 0) mbar_arrive_remote(tmem_empty_leader[h]);
        PP_ACC(e_rel);
      }
      if (kFinal) {                                                      // CLN == 1: the two column parts of a row meet in shared memory
#pragma unroll
        for (int j = 0; j < kOut; ++j) tail->part7[sub][row_in_tile][j] = acc2[j].x + acc2[j].y;
        epi_bar(1);
        if (sub == 0 && row < ep.m_valid) {
          int best = 0;
          float bv = 0.0f;
#pragma unroll
          for (int j = 0; j < kOut; ++j) {
            const float l = tail->part7[0][row_in_tile][j] + tail->part7[1][row_in_tile][j] + tail->b8[j];
            ep.logits[(size_t)row * kOut + j] = l;
            if (j == 0 || l > bv) { bv = l; best = j; }
          }
          if (ep.argmax) ep.argmax[row] = best;
        }
        epi_bar(2);                                                      // part7 is free again
      }
    }
    if (et == 0) { PP_DUMP(3, e_wait); PP_DUMP(4, e_p1); PP_DUMP(5, e_ld); PP_DUMP(6, e_p2); PP_DUMP(7, e_rel + ((long long)it << 40)); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();                                                        // nobody frees tensor memory or exits while the cluster still works
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NC));
  }
}

}  // namespace msa
