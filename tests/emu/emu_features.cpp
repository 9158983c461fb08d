// CPU emulation of the fused feature kernel (TEST INFRASTRUCTURE).
//
// Compiles the SAME kernel body as the CUDA build (csrc/msa_features_body.cuh) with g++ and
// runs one thread-block cluster per segment on OS threads: one std::thread per WARP, whose 32
// lanes run sequentially inside every lanes() call (per-lane state is kept in 32 copies, warp
// barriers are therefore no-ops), std::barrier for the block and cluster barriers, and plain
// pointers for distributed shared memory.  It exists so that the quad / tail / reflect-padding /
// ownership logic and the register FFTs can be checked against the oracle in a container without
// a GPU.  It is not a product path and is never loaded by msa_b200.
#include <atomic>
#include <barrier>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct float4 { float x, y, z, w; };
struct int2 { int x, y; };

#include "msa_features_body.cuh"

namespace msa {

struct CpuCluster {
  std::barrier<> bar;
  std::vector<unsigned char*> smem;
  explicit CpuCluster(int nthreads_total) : bar(nthreads_total) {}
};

struct CpuEnv {
  static constexpr int kStates = 32;
  int tid, nthreads, lane, warp, nwarps, rank, nranks, cluster_id;
  std::barrier<>* block;
  CpuCluster* cl;

  template <class F> void lanes(F&& f) {
    for (int l = 0; l < 32; ++l) f(l, l);
  }
  void sync() { block->arrive_and_wait(); }
  void wsync() {}
  void csync() { cl->bar.arrive_and_wait(); }
  template <class T> T* remote(T* p, int r) {
    return reinterpret_cast<T*>(cl->smem[r] + (reinterpret_cast<unsigned char*>(p) - cl->smem[rank]));
  }
  float log2(float v) { return std::log2(v); }
  float ld(const float* p) { return *p; }
  float ld(const int16_t* p) { return (float)*p * (1.0f / 32768.0f); }
  void prefetch(const void*) {}
  void prefetch_l1(const void*) {}
  float ld_last(const float* p) { return ld(p); }
  float ld_last(const int16_t* p) { return ld(p); }
  template <class InT> void ld4(const InT* x, int idx, int T, float* v) {
    for (int i = 0; i < 4; ++i) v[i] = (idx + i < T) ? ld(x + idx + i) : 0.0f;
  }
  void ldv(const float* x, int idx, int T, float* v) { ld4(x, idx, T, v); }
  void ldv(const int16_t* x, int idx, int T, float* v) { ld4(x, idx, T, v); ld4(x, idx + 4, T, v + 4); }
  // ---- tensor-core primitives of msa_pitch_tc.cuh, emulated over the 32 per-lane register copies
  uint32_t ldu(const uint32_t* p) { return *p; }
  float cospi(float v) { return (float)std::cos(3.14159265358979323846 * (double)v); }
  template <class InT> void ld16(const InT* x, int idx, float* v) { for (int i = 0; i < 16; ++i) v[i] = ld(x + idx + i); }
  void st8(uint16_t* dst, const uint32_t* w) { std::memcpy(dst, w, 16); }
  uint32_t lds1(const uint32_t* p) { return *p; }
  void lds2(uint32_t* d, const uint32_t* p) { d[0] = p[0]; d[1] = p[1]; }
  void lds4(uint32_t* d, const uint32_t* p) { for (int i = 0; i < 4; ++i) d[i] = p[i]; }
  static float h16(uint32_t v, int hi) { return hi ? msa::h2_hi(v) : msa::h2_lo(v); }
  template <class D, class A, class B> void mma(D dg, A ag, B bg, bool accumulate) {
    float Am[16][16], Bm[16][8], Cm[16][8];
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      const uint32_t* a = ag(l);
      const uint32_t* b = bg(l);
      const uint32_t* c = dg(l);
      for (int e = 0; e < 2; ++e) {
        Am[g][2 * t + e] = h16(a[0], e); Am[g + 8][2 * t + e] = h16(a[1], e);
        Am[g][2 * t + 8 + e] = h16(a[2], e); Am[g + 8][2 * t + 8 + e] = h16(a[3], e);
        Bm[2 * t + e][g] = h16(b[0], e); Bm[2 * t + 8 + e][g] = h16(b[1], e);
        Cm[g][2 * t + e] = accumulate ? h16(c[0], e) : 0.0f; Cm[g + 8][2 * t + e] = accumulate ? h16(c[1], e) : 0.0f;
      }
    }
    float Dm[16][8];
    for (int i = 0; i < 16; ++i)
      for (int j = 0; j < 8; ++j) {
        float s = Cm[i][j];
        for (int k = 0; k < 16; ++k) s += Am[i][k] * Bm[k][j];
        Dm[i][j] = s;
      }
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      uint32_t* d = dg(l);
      d[0] = msa::h2_pack(Dm[g][2 * t], Dm[g][2 * t + 1]);
      d[1] = msa::h2_pack(Dm[g + 8][2 * t], Dm[g + 8][2 * t + 1]);
    }
  }
  template <class D, class A, class B> void mma_f32(D dg, A ag, B bg) {
    float Am[16][16], Bm[16][8], Dm[16][8];
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      const uint32_t* a = ag(l);
      const uint32_t* b = bg(l);
      const float* c = dg(l);
      for (int e = 0; e < 2; ++e) {
        Am[g][2 * t + e] = h16(a[0], e); Am[g + 8][2 * t + e] = h16(a[1], e);
        Am[g][2 * t + 8 + e] = h16(a[2], e); Am[g + 8][2 * t + 8 + e] = h16(a[3], e);
        Bm[2 * t + e][g] = h16(b[0], e); Bm[2 * t + 8 + e][g] = h16(b[1], e);
        Dm[g][2 * t + e] = c[e]; Dm[g + 8][2 * t + e] = c[2 + e];
      }
    }
    for (int i = 0; i < 16; ++i)
      for (int j = 0; j < 8; ++j) {
        double s = Dm[i][j];                       // the tensor core adds the 16 exact products and the accumulator with
        for (int k = 0; k < 16; ++k) s += (double)Am[i][k] * (double)Bm[k][j];   // extra internal bits and rounds once
        Dm[i][j] = (float)s;
      }
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      float* d = dg(l);
      for (int e = 0; e < 2; ++e) { d[e] = Dm[g][2 * t + e]; d[2 + e] = Dm[g + 8][2 * t + e]; }
    }
  }
  float ldf(const float* p) { return *p; }
  template <class R, class P> void ldsm4t(R rg, P pg) {
    for (int m = 0; m < 4; ++m) {
      uint16_t tile[8][8];
      for (int r = 0; r < 8; ++r) std::memcpy(tile[r], pg(8 * m + r), 16);
      for (int l = 0; l < 32; ++l) {
        const int g = l >> 2, t = l & 3;
        rg(l)[m] = (uint32_t)tile[2 * t][g] | ((uint32_t)tile[2 * t + 1][g] << 16);     // transposed distribution
      }
    }
  }
  template <class D> void movmt(D dg) {
    uint16_t tile[8][8];
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      const uint32_t v = dg(l)[0];
      tile[g][2 * t] = (uint16_t)(v & 0xffffu); tile[g][2 * t + 1] = (uint16_t)(v >> 16);
    }
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      dg(l)[0] = (uint32_t)tile[2 * t][g] | ((uint32_t)tile[2 * t + 1][g] << 16);
    }
  }
  void copy16(void* dst, const void* src, int bytes) {
    if (warp == 0) std::memcpy(dst, src, bytes);
  }
  template <class G> double warp_reduce(int op, G&& get) {
    double a = get(0);
    for (int l = 1; l < 32; ++l) a = red_comb(op, a, get(l));
    return a;
  }
  void push(bool pred, int2* list, int& cnt, int cap, int key, float val) {   // lanes run in order: same list order as the GPU
    if (!pred) return;
    if (cnt < cap) { int bits; std::memcpy(&bits, &val, 4); list[cnt] = int2{key, bits}; }
    ++cnt;
  }
};

}  // namespace msa

extern "C" int emu_features_ws(const void* wav, int is_s16, int B, int T, int nranks, int nwarps, const float* emo8,
                               float* feat31, float* detail, float* dbg_mfcc, int flags, int parts, float* dbscratch);

extern "C" int emu_features(const void* wav, int is_s16, int B, int T, int nranks, int nwarps, const float* emo8,
                            float* feat31, float* detail, float* dbg_mfcc, int flags, int parts) {
  return emu_features_ws(wav, is_s16, B, T, nranks, nwarps, emo8, feat31, detail, dbg_mfcc, flags, parts, nullptr);
}

// dbscratch: [B, ceil((T / 200 + 1) / 4), 16, 32] floats or null (see FeatParams)
extern "C" int emu_features_ws(const void* wav, int is_s16, int B, int T, int nranks, int nwarps, const float* emo8,
                               float* feat31, float* detail, float* dbg_mfcc, int flags, int parts, float* dbscratch) {
  using namespace msa;
  static FeatureTables tab;
  static bool built = false;
  if (!built) {
    if (build_feature_tables(tab) != 0) return -100;
    built = true;
  }
  if (T <= kNfftP / 2) parts &= ~kPartPitch;
  if (T <= kNfftM / 2) parts &= ~kPartMfcc;
  FeatParams P{};
  P.wav = wav; P.is_s16 = is_s16; P.B = B; P.T = T; P.noise_n = (int)(0.05 * (double)T);
  P.emo8 = emo8; P.feat31 = feat31; P.detail = detail; P.dbg_mfcc = dbg_mfcc; P.dbscratch = dbscratch; P.tab = &tab; P.flags = flags; P.parts = parts;
  const FeatLayout lay = feat_layout(T, nranks, nwarps);
  for (int seg = 0; seg < B; ++seg) {
    CpuCluster cl(nranks * nwarps);
    std::vector<std::unique_ptr<unsigned char[]>> mem;
    std::vector<std::unique_ptr<std::barrier<>>> bars;
    for (int r = 0; r < nranks; ++r) {
      mem.emplace_back(new unsigned char[lay.total + 64]);
      std::memset(mem.back().get(), 0xCD, lay.total + 64);      // poison: catches reads of unwritten smem
      cl.smem.push_back(mem.back().get());
      bars.emplace_back(new std::barrier<>(nwarps));
    }
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r)
      for (int w = 0; w < nwarps; ++w)
        th.emplace_back([&, r, w]() {
          CpuEnv env{w * 32, nwarps * 32, 0, w, nwarps, r, nranks, seg, bars[r].get(), &cl};
          if (is_s16) features_cta<CpuEnv, int16_t>(env, P, cl.smem[r]);
          else features_cta<CpuEnv, float>(env, P, cl.smem[r]);
        });
    for (auto& t : th) t.join();
  }
  return 0;
}

extern "C" int emu_layout_bytes(int T, int nranks, int nwarps) { return msa::feat_layout(T, nranks, nwarps).total; }

// ---- the register butterflies on their own (checked against numpy.fft in tests/test_emu_features.py)
extern "C" void emu_dft(int n, int inverse, const float* in, float* out) {
  using namespace msa;
  c32 v[32];
  for (int i = 0; i < n; ++i) v[i] = c32{in[2 * i], in[2 * i + 1]};
  if (n == 16) { if (inverse) dft16<true>(v); else dft16<false>(v); }
  else if (n == 32) { if (inverse) dft32<true>(v); else dft32<false>(v); }
  else if (n == 25) { if (inverse) dft25<true>(v); else dft25<false>(v); }
  else if (n == 5) { if (inverse) dft5<true>(v[0], v[1], v[2], v[3], v[4]); else dft5<false>(v[0], v[1], v[2], v[3], v[4]); }
  for (int i = 0; i < n; ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; }
}

// full two-pass transform of one complex vector of 400 points through the tile layout
extern "C" int emu_fft(int n, const float* in, float* out) {
  using namespace msa;
  static FeatureTables tab;
  static bool built = false;
  if (!built) { if (build_feature_tables(tab) != 0) return -100; built = true; }
  if (n != 400) return -1;
  std::vector<c32> tile(kFftHalf);
  const int n2 = n / 16;
  const c32* tw = reinterpret_cast<const c32*>(tab.s.tw400);
  for (int lane = 0; lane < n2; ++lane) {
    c32 z[16];
    for (int n1 = 0; n1 < 16; ++n1) z[n1] = c32{in[2 * (n2 * n1 + lane)], in[2 * (n2 * n1 + lane) + 1]};
    pass_a_fwd<kRow400>(z, tw, lane, tile.data());
  }
  for (int k1 = 0; k1 < 16; ++k1) {
    c32 v[32];
    c32* row = tile.data() + k1 * kRow400;
    for (int i = 0; i < n2; ++i) v[i] = row[i];
    dft25<false>(v);
    for (int k2 = 0; k2 < n2; ++k2) { out[2 * (k1 + 16 * k2)] = v[k2].x; out[2 * (k1 + 16 * k2) + 1] = v[k2].y; }
  }
  return 0;
}
