#!/bin/bash
# Round 2 GPU call: A/B of the tensor-core pitch build against v11, the GPU tests, one full ncu capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt
MSA_AB_PITCH_TOL=1 MSA_AB_FUSION=${AB_FUSION:-0} timeout 300 ./scripts/ab_check scripts/ab/libmsa_v11.so multimodal-sentiment-analyzer_b200/libmsa_b200.so ${AB_EXTRA} > $O/${TAG}_ab.json 2> $O/${TAG}_ab.err
echo "ab rc=$?"; tail -c 1500 $O/${TAG}_ab.json; echo
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -15 $O/${TAG}_tests.log
timeout 300 python scripts/prof_features.py 1024 features > $O/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features -s 2 -c 1 -o $O/${TAG}_feat1024 \
    python scripts/prof_features.py 1024 features > $O/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/${TAG}_ncu.log
