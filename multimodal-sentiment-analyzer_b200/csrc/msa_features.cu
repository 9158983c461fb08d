// Launch side of the fused audio-feature kernel (see msa_features_body.cuh for the algorithm).
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "msa_api_internal.h"
#include "msa_env_gpu.cuh"

namespace msa {

template <class InT>
__global__ void __launch_bounds__(kFeatThreads, 2) features_kernel(const FeatParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  GpuEnv env;
  env.tid = threadIdx.x;
  env.nthreads = blockDim.x;
  env.lane = threadIdx.x & 31;
  env.warp = threadIdx.x >> 5;
  env.nwarps = blockDim.x >> 5;
  env.rank = (int)cluster.block_rank();
  env.nranks = (int)cluster.num_blocks();
  env.cluster_id = blockIdx.x / env.nranks;
  // a kernel launched behind this one with programmatic stream serialization (the first layer of the streaming fusion
  // path, msa_fusion_rows.cu) may be scheduled now: it fetches its weights and then waits for this grid to complete
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  features_cta<GpuEnv, InT>(env, P, smem);
}

static std::mutex g_tab_mutex;
static FeatureTables* g_tab_dev[64] = {nullptr};

static int get_tables(const FeatureTables** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return MSA_ERR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (!g_tab_dev[dev]) {
    static FeatureTables host;
    build_feature_tables(host);
    FeatureTables* d = nullptr;
    e = cudaMalloc(&d, sizeof(FeatureTables));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d, &host, sizeof(FeatureTables), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return (int)e; }
    g_tab_dev[dev] = d;
  }
  *out = g_tab_dev[dev];
  return MSA_OK;
}

int feat_threads() { return kFeatThreads; }

// SMs of the current device (read once per device; 148 on B200)
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Smallest cluster whose per-CTA shared-memory layout lets two CTAs share an SM, else the smallest that
// fits at all (the per-segment MFCC tile and the energy atoms are split over the ranks; everything else
// is per warp).
size_t features_workspace_bytes(int B, int T);

int features_cluster_size(int T) {
  if (T < 1) return 0;
  const int nwarps = feat_threads() / 32;
  for (int c = 1; c <= 8; c *= 2)                       // two CTAs per SM hide each other's serial phases
    if (feat_layout(T, c, nwarps).total <= kHalfSmem) return c;
  for (int c = 1; c <= 8; c *= 2)
    if (feat_layout(T, c, nwarps).total <= kMaxSmem) return c;
  return 0;
}

// Largest cluster the device schedules for this kernel: 16 CTAs (a non-portable size, asked for explicitly) where a GPC
// takes them, else 8.  Asked once per device and input type.
template <class InT>
static int max_cluster_size(int T) {
  static int cached[64] = {0}, cached_smem[64] = {0};                   // the answer depends on the CTA's shared memory
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 8;
  const FeatLayout lay = feat_layout(T, 16, feat_threads() / 32);
  if (cached[dev] == 0 || cached_smem[dev] != lay.total) {
    cached[dev] = 8;
    cached_smem[dev] = lay.total;
    auto kern = features_kernel<InT>;
    if (lay.total <= kMaxSmem && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total) == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(16, 1, 1);
      cfg.blockDim = dim3(feat_threads(), 1, 1);
      cfg.dynamicSmemBytes = lay.total;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 16;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) cached[dev] = 16;
    }
    (void)cudaGetLastError();
  }
  return cached[dev];
}

// Few segments (streaming: B = 1) spread over more CTAs to cut latency; large batches use the smallest
// cluster so that the quad lists per CTA stay long.
template <class InT>
static int auto_cluster_size(int B, int T) {
  int c = features_cluster_size(T);
  if (c == 0) return 0;
  // 16 CTAs per segment pay for a handful of segments only (measured per launch: 1 segment 52.6 -> 47.4 us, 4 segments
  // 53.4 -> 49.0, 8 segments 53.6 = 53.5, 16 segments 67 -> 97: profiles/r2_v204_tail_probe.json)
  const int cmax = ((long long)B * 16 * 2 <= sm_count()) ? max_cluster_size<InT>(T) : 8;
  while (c < cmax && (long long)B * c * 2 <= 2 * sm_count()) c *= 2;   // up to two CTAs per SM
  return c;
}

size_t features_workspace_bytes(int B, int T) {
  if (B < 1 || T < 1) return 0;
  const size_t quads = (size_t)((T / kHopM + 1) + 3) / 4;
  return (size_t)B * quads * 16 * 32 * sizeof(float);
}

int features_smem_bytes(int T, int c) {
  if (T < 1 || c < 1) return 0;
  return feat_layout(T, c, feat_threads() / 32).total;
}

template <class InT>
static int launch_features(const InT* wav, int B, int T, const float* emo8, float* feat31, float* detail,
                           float* dbg_mfcc, int flags, int parts, int cluster_size, void* workspace, size_t ws_bytes,
                           cudaStream_t stream) {
  if (B < 0 || T < 1) return MSA_ERR_BAD_ARGUMENT;
  if (B == 0) return MSA_OK;                                   // an empty batch has no buffers to check
  if (!wav || !feat31) return MSA_ERR_BAD_ARGUMENT;
  int c = cluster_size ? cluster_size : auto_cluster_size<InT>(B, T);
  if (c != 1 && c != 2 && c != 4 && c != 8 && c != 16) return c == 0 ? MSA_ERR_UNSUPPORTED_LENGTH : MSA_ERR_BAD_ARGUMENT;
  if (c == 16 && max_cluster_size<InT>(T) < 16) return MSA_ERR_BAD_ARGUMENT;   // this device does not co-schedule 16 CTAs of this kernel
  // torch.stft's reflect padding needs T > n_fft/2: below that the reference's method raises and
  // returns its default, which is what a cleared part bit produces.
  if (T <= kNfftP / 2) parts &= ~kPartPitch;
  if (T <= kNfftM / 2) parts &= ~kPartMfcc;
  const FeatureTables* tab = nullptr;
  int rc = get_tables(&tab);
  if (rc != MSA_OK) return rc;

  FeatParams P{};
  P.wav = wav;
  P.is_s16 = sizeof(InT) == 2;
  P.B = B;
  P.T = T;
  P.noise_n = (int)(0.05 * (double)T);   // int(0.05 * waveform.shape[1]), audio_analyzer.py:282
  P.emo8 = emo8;
  P.feat31 = feat31;
  P.detail = detail;
  P.dbg_mfcc = dbg_mfcc;
  // optional scratch table of mel dB values (see FeatParams::dbscratch); too small or absent: quads are recomputed
  P.dbscratch = (workspace != nullptr && ws_bytes >= features_workspace_bytes(B, T)) ? static_cast<float*>(workspace) : nullptr;
  P.tab = tab;
  P.flags = flags;
  P.parts = parts;
  const int threads = feat_threads();
  const FeatLayout lay = feat_layout(T, c, threads / 32);
  if (lay.total > kMaxSmem) return MSA_ERR_UNSUPPORTED_LENGTH;

  auto kern = features_kernel<InT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)B * c, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = lay.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) return (int)e;
  note_launches(1);
  return MSA_OK;
}

}  // namespace msa

extern "C" int msa_features_cluster_size(int T) { return msa::features_cluster_size(T); }
extern "C" int msa_features_smem_bytes(int T, int c) { return msa::features_smem_bytes(T, c); }

extern "C" size_t msa_features_workspace_bytes(int B, int T) { return msa::features_workspace_bytes(B, T); }

extern "C" int msa_features_f32(const float* wav, int B, int T, const float* emo8, float* feat31, float* detail,
                                float* dbg_mfcc, int flags, int parts, int cluster_size, void* stream) {
  msa::reset_launches();
  return msa::launch_features<float>(wav, B, T, emo8, feat31, detail, dbg_mfcc, flags, parts, cluster_size, nullptr, 0,
                                     (cudaStream_t)stream);
}

extern "C" int msa_features_ws_f32(const float* wav, int B, int T, const float* emo8, float* feat31, float* detail,
                                   float* dbg_mfcc, int flags, int parts, int cluster_size, void* workspace, size_t ws_bytes,
                                   void* stream) {
  msa::reset_launches();
  return msa::launch_features<float>(wav, B, T, emo8, feat31, detail, dbg_mfcc, flags, parts, cluster_size, workspace, ws_bytes,
                                     (cudaStream_t)stream);
}

extern "C" int msa_features_ws_s16(const int16_t* pcm, int B, int T, const float* emo8, float* feat31, float* detail,
                                   float* dbg_mfcc, int flags, int parts, int cluster_size, void* workspace, size_t ws_bytes,
                                   void* stream) {
  msa::reset_launches();
  return msa::launch_features<int16_t>(pcm, B, T, emo8, feat31, detail, dbg_mfcc, flags, parts, cluster_size, workspace, ws_bytes,
                                       (cudaStream_t)stream);
}

extern "C" int msa_features_s16(const int16_t* pcm, int B, int T, const float* emo8, float* feat31, float* detail,
                                float* dbg_mfcc, int flags, int parts, int cluster_size, void* stream) {
  msa::reset_launches();
  return msa::launch_features<int16_t>(pcm, B, T, emo8, feat31, detail, dbg_mfcc, flags, parts, cluster_size, nullptr, 0,
                                       (cudaStream_t)stream);
}
