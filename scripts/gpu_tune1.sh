#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/t_feat.log 2>&1; echo "feat rc=$?"; tail -4 gpurun_out/t_feat.log
for th in 256 512; do for sl in 20000 10000 5000; do
  MSA_FEAT_THREADS=$th MSA_FEAT_SLICE=$sl timeout 120 python scripts/time_features.py 1024 f32 2>&1 | tail -1
done; done | tee gpurun_out/tune1.log
MSA_FEAT_THREADS=512 MSA_FEAT_SLICE=10000 timeout 120 python scripts/time_features.py 1024 s16 2>&1 | tail -1 | tee -a gpurun_out/tune1.log
