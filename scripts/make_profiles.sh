#!/bin/bash
# One gpurun call that produces every artefact under profiles/ (run on the GPU box from the repo root):
#   launch list (per-kernel gpu__time_duration) of one eager bs=64 forward, and ncu --set full captures of the
#   CTA-pair GEMM (the four launches of ViT block 1: qkv, proj, fc1+GELU, fc2), the attention kernel and LayerNorm.
export PYTHONPATH=.
mkdir -p gpurun_out
R=${ROUND:-r1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bs64.csv \
    python scripts/profile_forward.py 64 2 > gpurun_out/launches.log 2>&1
# eager forward #1: ViT block 0's four big GEMMs are CTA-pair launches 0..3, block 1's are 4..7
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_f16_tn_2cta -s 4 -c 4 -f \
    -o gpurun_out/${R}_prof_gemm python scripts/profile_forward.py 64 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:vit_attention3 -s 1 -c 1 -f \
    -o gpurun_out/${R}_prof_attn python scripts/profile_forward.py 64 1 > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:layernorm_reg -s 2 -c 1 -f \
    -o gpurun_out/${R}_prof_ln python scripts/profile_forward.py 64 1 > gpurun_out/ncu_ln.log 2>&1
ls -la gpurun_out | tail -8
