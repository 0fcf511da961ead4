#!/bin/bash
# Round-2 GPU check 2: stream-K + L2 prefetch of the reduce-add target: parity, isolated GEMM rates, in-step A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_strict.py -m gpu -q -x -s > $O/c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/c2_pytest.log
tail -3 $O/c2_pytest.log
for v in "X=1" "THMR_GEMM_STREAMK=0" "THMR_GEMM_DBG=512" ; do
  n=$(echo $v | tr '=' '_')
  env $v timeout 300 python scripts/dev_gemm_perf.py > $O/c2_gemm_$n.log 2>&1
  echo "== $v"; cat $O/c2_gemm_$n.log
done
for v in "X=1" "THMR_GEMM_STREAMK=0" "THMR_GEMM_DBG=512"; do
  n=$(echo $v | tr '=' '_')
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/c2_bench_$n.json 2> $O/c2_bench_$n.err; echo "$v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c2_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        fam=d['kernel_families']
        print(f, 'ms/step %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], 'clk',d['clocks']['sm_mhz'],
              ' '.join('%s=%.2f'%(k.split('.')[1][:6],v['ms_per_step']) for k,v in fam.items() if k.startswith('vit')))
    except Exception as e:
        print(f, 'ERR', e)
PY
