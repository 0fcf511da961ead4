#!/bin/bash
# Exploration call: stream counts / pipeline depths, VQ group-min screen, one-wave skinning launch, LBS launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "vq or lbs" > $O/r2c_pytest_new.log 2>&1
el "targeted pytest rc=$?"; tail -3 $O/r2c_pytest_new.log
timeout 120 python scripts/dev_vq_lbs.py > $O/r2c_vq_lbs.log 2>&1; el "vq/lbs rc=$?"; cat $O/r2c_vq_lbs.log
timeout 300 python scripts/dev_streams.py 24 > $O/r2c_streams.log 2>&1; el "streams rc=$?"; cat $O/r2c_streams.log | tail -16
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2c_lbs_launches.csv \
    python scripts/dev_vq_lbs.py lbs > $O/r2c_ncu_lbs_list.log 2>&1
el "ncu lbs list rc=$?"
