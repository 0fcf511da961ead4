#!/bin/bash
# A/B of env-selected variants with alternation (the box's power / clock state drifts between runs): usage
#   scripts/r2_ab.sh <rounds> <steps> VAR1=.. VAR2=.. ...   (use X=1 for the default)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
R=$1; S=$2; shift 2
for r in $(seq 1 $R); do
  for v in "$@"; do
    n=$(echo $v | tr '=' '_' | tr ' ' '_')
    env $v timeout 300 python bench.py --steps $S --warmup 5 --no-cpu-baseline --no-extras > $O/ab_${n}_$r.json 2> $O/ab_${n}_$r.err || echo "$v run $r failed"
  done
done
python - <<'PY'
import json,glob,collections
acc=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    name=f.split('/')[-1][3:].rsplit('_',1)[0]
    fam=d['kernel_families']
    acc[name].append((d['ms_per_step'], d['clocks']['sm_mhz'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in fam.items()}))
for name,rows in acc.items():
    ms=[r[0] for r in rows]
    print(f"{name:28s} ms/step {' '.join('%.3f'%m for m in ms)}  min {min(ms):.3f} | clk {' '.join('%d'%r[1] for r in rows)} | e2e {' '.join('%.2f'%r[2] for r in rows)}")
    fam=rows[ms.index(min(ms))][3]
    print('      ', ' '.join('%s=%.2f'%(k.split('.')[1][:7],v) for k,v in fam.items() if k.startswith('vit')))
PY
