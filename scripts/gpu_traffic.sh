#!/bin/bash
# DRAM bytes of one feature-kernel launch at the bench batch (ncu, two metrics only) + its plain timing.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-traffic}; O=gpurun_out; mkdir -p $O
timeout 120 python scripts/time_features.py 1024 f32 | tail -1
timeout 120 python scripts/time_features.py 1024 s16 | tail -1
timeout 300 python scripts/prof_features.py 1024 features > $O/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:features -s 2 -c 1 --csv \
    --log-file $O/${TAG}_dram.csv python scripts/prof_features.py 1024 features > $O/${TAG}_ncu.log 2>&1
tail -5 $O/${TAG}_dram.csv | cut -d, -f5,13-
