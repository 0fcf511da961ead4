#!/bin/bash
# same-box A/B of engine variants: prints ms/step and the family table for each env setting
export PYTHONPATH=.
for cfg in ${AB_CFGS:-base THMR_ATTN_GEN=2 THMR_ATTN_GEN=1 THMR_GEMM_2CTA=0}; do
  if [ "$cfg" = "base" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
f=d['kernel_families']
print('%-22s %.2f ms/step  %.0f img/s | attn %.2f  qkv %.2f proj %.2f fc1 %.2f fc2 %.2f ln %.2f dec %.2f | clk %s' % ('$cfg', d['ms_per_step'], d['value'], f['vit.attention']['ms_per_step'], f['vit.qkv_gemm']['ms_per_step'], f['vit.proj_gemm']['ms_per_step'], f['vit.fc1_gelu_gemm']['ms_per_step'], f['vit.fc2_gemm']['ms_per_step'], f['vit.layernorm']['ms_per_step'], f['dec.token_ops']['ms_per_step'], d['clocks']['sm_mhz']))
"
done
