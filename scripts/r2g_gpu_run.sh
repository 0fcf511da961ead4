#!/bin/bash
# Final-state check: whole GPU suite, VQ (packed-key screen) / LBS timings, bench line, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
export THMR_BENCH_WATCHDOG=300
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 400 python -m pytest tests -m gpu -q > $O/r2g_pytest_gpu.log 2>&1
el "full pytest rc=$?"; tail -6 $O/r2g_pytest_gpu.log
timeout 120 python scripts/dev_vq_lbs.py > $O/r2g_vq_lbs.log 2>&1; el "vq/lbs rc=$?"; cat $O/r2g_vq_lbs.log
timeout 420 python bench.py --steps 20 --warmup 5 > $O/r2g_bench_b200_n1.json 2> $O/r2g_bench_b200_n1.err
el "bench rc=$?"; grep early $O/r2g_bench_b200_n1.err | head -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1
el "smoke rc=$?"; tail -2 $O/r2g_smoke.log
