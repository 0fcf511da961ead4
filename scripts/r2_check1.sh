#!/bin/bash
# Round-2 GPU check 1: smplx probe, full GPU test suite, bench, and A/B runs of the new epilogue / sub-batching.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $O/c1_gpu.txt 2>&1
python -c "import smplx; print('smplx', smplx.__version__, smplx.__file__)" > $O/c1_smplx.log 2>&1; echo "rc=$?" >> $O/c1_smplx.log
python -m pip list 2>/dev/null | grep -i -E "smpl|chumpy|trimesh" >> $O/c1_smplx.log
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 -s > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
tail -5 $O/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc=$?"
for v in "THMR_GEMM_DBG=256" "THMR_VIT_SUB=32" "THMR_VIT_SUB=16"; do
  n=$(echo $v | tr '=' '_')
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/c1_bench_$n.json 2> $O/c1_bench_$n.err; echo "$v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c1_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        fam=d['kernel_families']
        print(f, 'ms/step %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], 'clk',d['clocks']['sm_mhz'],
              ' '.join('%s=%.2f'%(k.split('.')[1][:6],v['ms_per_step']) for k,v in fam.items() if k.startswith('vit')))
        if d.get('strict'): print('  strict', d['strict']['value'], d['strict']['ms_per_step'])
        if d.get('standalone'): print('  standalone', json.dumps(d['standalone'])[:600])
    except Exception as e:
        print(f, 'ERR', e)
PY
