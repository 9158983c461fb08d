#!/bin/bash
# First GPU bring-up: each stage in its own process with its own timeout; logs under gpurun_out/.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== features"; timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q -x > gpurun_out/t_feat.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_feat.log
echo "== fusion simt"; timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q -k "simt" > gpurun_out/t_simt.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_simt.log
echo "== fusion tcgen05 golden"; timeout 300 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x -k "tcgen05 and golden" > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_tc.log
echo "== fusion rest"; timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q -k "not simt and not golden" > gpurun_out/t_rest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_rest.log
echo "== bench simt"; MSA_FUSION_IMPL=simt timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_simt.log
echo "== bench tc"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_tc.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
