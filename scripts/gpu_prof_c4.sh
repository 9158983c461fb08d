#!/bin/bash
# feature kernel ncu at the bench batch (traffic per launch) + fusion-only config (timing, ncu of the GEMM kernels)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-c4}; O=gpurun_out; mkdir -p $O
for B in 1024 8192 65536; do timeout 300 python scripts/time_fusion.py $B 2>&1 | tail -1; done | tee $O/${TAG}_fusion_time.log
timeout 300 python scripts/prof_features.py 1024 features > $O/${TAG}_plainf.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:features -s 2 -c 1 -o $O/${TAG}_feat1024 python scripts/prof_features.py 1024 features > $O/${TAG}_ncu_f.log 2>&1
echo "ncu feat rc=$?"
timeout 300 python scripts/time_fusion.py 65536 2 > $O/${TAG}_plainc4.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -k regex:tc_ -s 16 -c 5 -o $O/${TAG}_fusion65536 python scripts/time_fusion.py 65536 2 > $O/${TAG}_ncu_c4.log 2>&1
echo "ncu fusion rc=$?"; tail -2 $O/${TAG}_ncu_c4.log
