#!/bin/bash
# full GPU parity suite + smoke, then one full ncu capture of the feature kernel (after a plain run of the same command)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-v2b}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
timeout 120 python scripts/time_features.py 1024 f32 2>&1 | tail -1
timeout 120 python scripts/time_features.py 1024 s16 2>&1 | tail -1
timeout 300 python scripts/prof_features.py 296 features > $O/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features -s 2 -c 1 -o $O/${TAG}_feat \
    python scripts/prof_features.py 296 features > $O/${TAG}_ncu_feat.log 2>&1
echo "ncu rc=$?"; tail -2 $O/${TAG}_ncu_feat.log
