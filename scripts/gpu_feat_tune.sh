#!/bin/bash
# v2 feature kernel bring-up: feature parity tests, then kernel-only timings (CUDA events).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-v2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q -x > gpurun_out/${TAG}_feat.log 2>&1; echo "feat rc=$?"; tail -15 gpurun_out/${TAG}_feat.log
for th in 512 256; do
  MSA_FEAT_THREADS=$th timeout 120 python scripts/time_features.py 1024 f32 2>&1 | tail -1
  MSA_FEAT_THREADS=$th timeout 120 python scripts/time_features.py 1024 s16 2>&1 | tail -1
done | tee gpurun_out/${TAG}_tune.log
for B in 1 8 148 296 2048; do
  timeout 120 python scripts/time_features.py $B f32 2>&1 | tail -1
done | tee -a gpurun_out/${TAG}_tune.log
