#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=${1:-2}
NCCL_DEBUG=INFO THMR_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) bench.py --gpus $N --steps 20 --warmup 5 > $O/c6_n$N.out 2> $O/c6_n$N.err
echo "== N=$N rc=$?"
grep -c "NCCL INFO" $O/c6_n$N.out; grep "Init COMPLETE" $O/c6_n$N.out | head -4
tail -n 1 $O/c6_n$N.out | head -c 1200; echo
grep -E "File|Error|error" $O/c6_n$N.err | head -20
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/c6_n1.out 2>$O/c6_n1.err
python - <<PY
import json
for f in ("gpurun_out/c6_n1.out","gpurun_out/c6_n$N.out"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "value %.1f"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], d["clocks"], d.get("exchange"))
    except Exception as e: print(f, "ERR", e)
PY
