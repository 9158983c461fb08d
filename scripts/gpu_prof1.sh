#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/t_feat.log 2>&1; echo "feat rc=$?"; tail -4 gpurun_out/t_feat.log
timeout 300 python scripts/prof_features.py 296 features > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features_kernel -s 2 -c 1 -o gpurun_out/feat_r1 python scripts/prof_features.py 296 features > gpurun_out/ncu_feat.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_feat.log
