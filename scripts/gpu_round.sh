#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (native + reference arm), ncu launch list of the
# bench command and one full ncu capture of the feature kernel.  Logs land in gpurun_out/<tag>_*.
# usage: scripts/gpu_round.sh <tag> [skip-tests]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
if [ "$2" != "skip-tests" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -4 $O/${TAG}_tests.log
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${TAG}_smoke.log
fi
echo "== bench"; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "rc=$?"; tail -c 3000 $O/${TAG}_bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "rc=$?"; tail -c 1500 $O/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1
echo "rc=$?"
echo "== ncu full (feature kernel)"
timeout 300 python scripts/prof_features.py 296 features > $O/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features -s 2 -c 1 -o $O/${TAG}_feat \
    python scripts/prof_features.py 296 features > $O/${TAG}_ncu_feat.log 2>&1
echo "rc=$?"; tail -2 $O/${TAG}_ncu_feat.log
