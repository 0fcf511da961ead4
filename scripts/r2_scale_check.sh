#!/bin/bash
# One multi-GPU bench run launched the way the driver launches it (torchrun, NCCL_DEBUG=INFO), bounded by a watchdog.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=${1:-2}
NCCL_DEBUG=INFO THMR_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_scale_n$N.out 2> $O/r2_scale_n$N.err
echo "== N=$N rc=$?"
grep -c "NCCL INFO" $O/r2_scale_n$N.out; grep "Init COMPLETE" $O/r2_scale_n$N.out | head -2
tail -n 1 $O/r2_scale_n$N.out > $O/r2_bench_b200_n$N.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_b200_n$N.json").read()); print("N", d["n_gpus"], "value %.1f"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "d2h", d["e2e"]["d2h_bytes_per_step"], d["clocks"], d.get("exchange"))
except Exception as e: print("ERR", e)
PY
grep -E "File|Error|error" $O/r2_scale_n$N.err | head -10
