#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH=.
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/dev_streams.py 24 4,6,8 4x4,8x4,12x4,6x3,12x6,8x8 > $O/r2e_streams.log 2>&1; echo "streams rc=$?"; tail -14 $O/r2e_streams.log | head -13
