#!/bin/bash
# 2-GPU check of the sharded path with this session's bench / pipeline changes: NCCL tests + the N=2 bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
export THMR_BENCH_WATCHDOG=240
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q > $O/r2f_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; tail -2 $O/r2f_pytest_nccl.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/r2f_nccl_%h_%p.log timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2f_bench_b200_n2.json 2> $O/r2f_bench_b200_n2.err; echo "bench n2 rc=$?"
tail -c 700 $O/r2f_bench_b200_n2.json | head -c 700; echo; grep -h "nranks" $O/r2f_nccl_*.log | head -2
