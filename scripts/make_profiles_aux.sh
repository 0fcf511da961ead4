#!/bin/bash
# ncu --set full captures of the stand-alone configurations (BASELINE.json configs 4 / 5) and of the pre-processing kernel:
#   VQ nearest-code arg-min GEMM (1 M queries), SMPL skinning kernel (4096 poses), 8-bit crop kernel (64 persons, 1080p).
export PYTHONPATH=.
mkdir -p gpurun_out
R=${ROUND:-r1}
timeout 300 ncu --set full --clock-control none -k regex:gemm_f16_tn_kernel -s 2 -c 1 -f -o gpurun_out/${R}_prof_vq \
    python scripts/dev_vq_lbs.py vq > gpurun_out/ncu_vq.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:smpl_skin_kernel -s 16 -c 1 -f -o gpurun_out/${R}_prof_lbs \
    python scripts/dev_vq_lbs.py lbs > gpurun_out/ncu_lbs.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:preproc_warp_u8 -s 3 -c 1 -f -o gpurun_out/${R}_prof_pre \
    python scripts/dev_pre_enc_perf.py > gpurun_out/ncu_pre.log 2>&1
ls -la gpurun_out/*.ncu-rep
