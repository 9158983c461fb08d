// Internal helpers shared by the translation units of libmsa_b200.so.
#pragma once
#include "../../include/msa_b200.h"

namespace msa {
constexpr int kFeatThreads = 256;      // default threads per feature CTA: 8 warps x 128 registers, two CTAs per SM
int sm_count();                       // SMs of the current device (msa_features.cu)
constexpr int kMaxSmem = 232448;   // 227 KB opt-in shared memory per CTA on sm_100
constexpr int kHalfSmem = 115712;  // (228 KB - 2 x 1 KB reserved) / 2: two CTAs per SM
void reset_launches();
void note_launches(int n);
}  // namespace msa
