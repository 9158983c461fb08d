// Small C-ABI utilities: version, error strings, launch accounting.
#include <cuda_runtime.h>

#include "msa_api_internal.h"

namespace msa {
static thread_local int g_launches = 0;
void reset_launches() { g_launches = 0; }
void note_launches(int n) { g_launches += n; }
}  // namespace msa

extern "C" int msa_version(void) { return 207; }   // 2xx: round 2 (tensor-core STFT-512 round trip)
extern "C" int msa_last_launch_count(void) { return msa::g_launches; }

extern "C" const char* msa_strerror(int code) {
  switch (code) {
    case MSA_OK: return "ok";
    case MSA_ERR_BAD_ARGUMENT: return "bad argument";
    case MSA_ERR_UNSUPPORTED_LENGTH: return "segment too long for one cluster (see msa_features_cluster_size)";
    case MSA_ERR_NOT_PACKED: return "fusion weights not packed";
    case MSA_ERR_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown msa error";
}
