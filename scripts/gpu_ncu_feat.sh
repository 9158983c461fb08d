#!/bin/bash
# One full ncu capture of the feature kernel (after a plain run of the same command) + kernel-only timings.
# usage: scripts/gpu_ncu_feat.sh <tag> [f32|s16]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-prof}; DT=${2:-f32}
O=gpurun_out; mkdir -p $O
timeout 120 python scripts/time_features.py 1024 f32 2>&1 | tail -1 | tee $O/${TAG}_time.log
timeout 120 python scripts/time_features.py 1024 s16 2>&1 | tail -1 | tee -a $O/${TAG}_time.log
timeout 300 python scripts/prof_features.py 296 features > $O/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features -s 2 -c 1 -o $O/${TAG}_feat \
    python scripts/prof_features.py 296 features > $O/${TAG}_ncu_feat.log 2>&1
echo "ncu rc=$?"; tail -2 $O/${TAG}_ncu_feat.log
