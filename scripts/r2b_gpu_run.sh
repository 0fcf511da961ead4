#!/bin/bash
# Round-2, second session: ONE short gpurun call (the round's GPU budget was nearly spent).  In order of value:
# targeted tests of the new paths, the full bench line (one-stream vs two-stream replay, VQ screened vs exact, skinning launch
# shapes), smoke(), the whole GPU suite, then ncu launch list + full captures of the screened VQ kernels and the skinning kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
export THMR_BENCH_WATCHDOG=300
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/r2b_gpu.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -k "vq or lbs or pipeline or gemm" > $O/r2b_pytest_new.log 2>&1
el "targeted pytest rc=$?"; tail -3 $O/r2b_pytest_new.log
timeout 420 python bench.py --steps 20 --warmup 5 > $O/r2b_bench_b200_n1.json 2> $O/r2b_bench_b200_n1.err
el "bench rc=$?"; grep early $O/r2b_bench_b200_n1.err; tail -c 600 $O/r2b_bench_b200_n1.err | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1
el "smoke rc=$?"; tail -2 $O/r2b_smoke.log
timeout 400 python -m pytest tests -m gpu -q > $O/r2b_pytest_gpu.log 2>&1
el "full pytest rc=$?"; tail -4 $O/r2b_pytest_gpu.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file $O/r2b_vq_launches.csv \
    python scripts/dev_vq_lbs.py vq > $O/r2b_ncu_vq_list.log 2>&1
el "ncu vq list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tn_kernel -s 6 -c 3 -f -o $O/r2b_prof_vq \
    python scripts/dev_vq_lbs.py vq > $O/r2b_ncu_vq.log 2>&1
el "ncu vq full rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:smpl_skin_kernel -s 16 -c 1 -f -o $O/r2b_prof_lbs \
    python scripts/dev_vq_lbs.py lbs > $O/r2b_ncu_lbs.log 2>&1
el "ncu lbs full rc=$?"
ls -la $O | tail -20
