#!/bin/bash
# Round-end evidence: full round (tests, smoke, bench both arms, launch list, ncu of the feature kernel at B=296)
# plus one ncu --set full capture at the bench batch (traffic per launch).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-final}
bash scripts/gpu_round.sh $TAG
O=gpurun_out
timeout 300 python scripts/prof_features.py 1024 features > $O/${TAG}_plainf.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:features -s 2 -c 1 -o $O/${TAG}_feat1024 python scripts/prof_features.py 1024 features > $O/${TAG}_ncu_f.log 2>&1
echo "ncu feat1024 rc=$?"
