#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export MSA_FEAT_THREADS=256 MSA_FEAT_SLICE=20000
timeout 300 python scripts/prof_features.py 296 features > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features_kernel -s 2 -c 1 -o gpurun_out/feat_r2 python scripts/prof_features.py 296 features > gpurun_out/ncu_feat.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_feat.log
