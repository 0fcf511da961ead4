#!/bin/bash
# Last call of the round: whole GPU suite on the final tree, ncu --set full of the packed-key pass-1 kernel + launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 300 python -m pytest tests -m gpu -q > $O/r2h_pytest_gpu.log 2>&1
el "full pytest rc=$?"; tail -4 $O/r2h_pytest_gpu.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tn_kernel -s 6 -c 3 -f -o $O/r2h_prof_vq \
    python scripts/dev_vq_lbs.py vq > $O/r2h_ncu_vq.log 2>&1
el "ncu vq full rc=$?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2h_vq_launches.csv \
    python scripts/dev_vq_lbs.py vq > $O/r2h_ncu_vq_list.log 2>&1
el "ncu vq list rc=$?"
