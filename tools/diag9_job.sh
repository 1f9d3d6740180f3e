#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism7.log; : > $O
run() { env "$@" timeout 240 python tools/determinism.py C4r 150 2>&1 | grep -v "^$" | tail -8 >> $O; tail -1 $O | cut -c1-330; }
run A=1; run CB200_TMA_TILE=128
timeout 400 python tools/fine_breakdown.py C5 C4 > gpurun_out/r02_fine_lagged.log 2>&1; grep -E "^==|schur_gemm|panel_update|piv_" gpurun_out/r02_fine_lagged.log | cut -c1-170
timeout 300 python -m pytest tests/test_gpu_kkt.py -x -q -m gpu > gpurun_out/r02_pytest_lagged.log 2>&1; tail -2 gpurun_out/r02_pytest_lagged.log
