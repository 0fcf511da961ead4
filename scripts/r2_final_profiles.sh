#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, LBS / VQ stand-alone timings, ncu launch list and
# --set full captures (GEMMs, attention, LayerNorm, VQ, LBS), sustained GEMM table vs cuBLAS, and the bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2_pytest_gpu.log
python scripts/dev_vq_lbs.py > $O/r2_vq_lbs.log 2>&1; cat $O/r2_vq_lbs.log
ROUND=r2 bash scripts/make_profiles.sh > $O/r2_make_profiles.log 2>&1
ROUND=r2 bash scripts/make_profiles_aux.sh > $O/r2_make_profiles_aux.log 2>&1
python scripts/dev_sustained.py 2.0 > $O/r2_sustained_gemm.log 2>&1; cat $O/r2_sustained_gemm.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_b200_n1.json 2> $O/r2_bench_b200_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_cpu.json 2> $O/r2_bench_reference_cpu.err; echo "ref rc=$?"
ls -la $O/*.ncu-rep | tail -8
