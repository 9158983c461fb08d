// CPU emulation of the fused feature kernel (TEST INFRASTRUCTURE).
//
// Compiles the SAME kernel body as the CUDA build (csrc/msa_features_body.cuh) with g++ and
// runs one thread-block cluster per segment on OS threads: one std::thread per WARP (lanes are
// executed sequentially inside the per-lane loops, nlanes = 1), std::barrier for the block and
// cluster barriers, and plain pointers for distributed shared memory.  It exists so that the
// slicing / ownership / overlap-add / digit-reversal logic can be checked against the oracle in
// a container without a GPU.  It is not a product path and is never loaded by msa_b200.
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "msa_features_body.cuh"

namespace msa {

struct CpuCluster {
  std::barrier<> bar;
  std::vector<unsigned char*> smem;
  explicit CpuCluster(int nthreads_total) : bar(nthreads_total) {}
};

struct CpuEnv {
  static constexpr int kLanes = 1;
  int tid, nthreads, lane, nlanes, warp, nwarps, rank, nranks, cluster_id;
  std::barrier<>* block;
  CpuCluster* cl;

  void sync() { block->arrive_and_wait(); }
  void wsync() {}
  void csync() { cl->bar.arrive_and_wait(); }
  template <class T> T* remote(T* p, int r) {
    return reinterpret_cast<T*>(cl->smem[r] + (reinterpret_cast<unsigned char*>(p) - cl->smem[rank]));
  }
  double wsum(double v) { return v; }
  float wmax(float v) { return v; }
  double bsum(double v, double* red) {
    red[tid] = v;
    sync();
    double s = 0.0;
    for (int i = 0; i < nthreads; ++i) s += red[i];
    sync();
    return s;
  }
  float bmax(float v, double* red) {
    float* r = reinterpret_cast<float*>(red);
    r[tid] = v;
    sync();
    float s = r[0];
    for (int i = 1; i < nthreads; ++i) s = std::fmax(s, r[i]);
    sync();
    return s;
  }
  template <class InT> void load_slice(float* dst, const InT* src, int n, void*, bool) {
    for (int i = tid; i < n; i += nthreads) dst[i] = to_f32<InT>(src[i]);
  }
};

}  // namespace msa

extern "C" int emu_features(const void* wav, int is_s16, int B, int T, int nranks, int nwarps, const float* emo8,
                            float* feat31, float* detail, float* dbg_mfcc, int flags, int parts) {
  using namespace msa;
  static FeatureTables tab;
  static bool built = false;
  if (!built) { build_feature_tables(tab); built = true; }
  int L = (T + nranks - 1) / nranks;
  L = ((L + kAtom - 1) / kAtom) * kAtom;
  FeatParams P{};
  P.wav = wav; P.is_s16 = is_s16; P.B = B; P.T = T; P.slice_len = L; P.noise_n = (int)(0.05 * (double)T);
  P.emo8 = emo8; P.feat31 = feat31; P.detail = detail; P.dbg_mfcc = dbg_mfcc; P.tab = &tab; P.flags = flags; P.parts = parts;
  const FeatLayout lay = feat_layout(L, nwarps);
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
          CpuEnv env{w, nwarps, 0, 1, w, nwarps, r, nranks, seg, bars[r].get(), &cl};
          if (is_s16) features_cta<CpuEnv, int16_t>(env, P, cl.smem[r]);
          else features_cta<CpuEnv, float>(env, P, cl.smem[r]);
        });
    for (auto& t : th) t.join();
  }
  return 0;
}

extern "C" int emu_layout_bytes(int T, int nranks, int nwarps) {
  int L = (T + nranks - 1) / nranks;
  L = ((L + msa::kAtom - 1) / msa::kAtom) * msa::kAtom;
  return msa::feat_layout(L, nwarps).total;
}
