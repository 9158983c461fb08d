#!/bin/bash
# N-GPU run: sharded-hour check (configs[2]) and the bench contract line at N GPUs (both arms).
# usage: scripts/gpu_multi.sh <N> <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; TAG=${2:-multi}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/${TAG}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/check_sharded.py > $O/${TAG}_sharded.json 2> $O/${TAG}_sharded.err; echo "sharded rc=$?"; tail -1 $O/${TAG}_sharded.json
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 2500 $O/${TAG}_bench.json
tail -3 $O/${TAG}_bench.err
