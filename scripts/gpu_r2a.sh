#!/bin/bash
# Round 2, first GPU call: A/B of the tensor-core pitch build against v11, the GPU tests, one full ncu capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_gpu.txt
MSA_AB_PITCH_TOL=1 MSA_AB_FUSION=0 timeout 200 ./scripts/ab_check scripts/ab/libmsa_v11.so multimodal-sentiment-analyzer_b200/libmsa_b200.so > $O/r2a_ab.json 2> $O/r2a_ab.err
echo "ab rc=$?"; tail -c 1500 $O/r2a_ab.json; echo
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -5 $O/r2a_tests.log
timeout 300 python scripts/prof_features.py 1024 features > $O/r2a_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features -s 2 -c 1 -o $O/r2a_feat1024 \
    python scripts/prof_features.py 1024 features > $O/r2a_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r2a_ncu.log
