#!/bin/bash
# Diagnostic GPU job: read-before-write hunt (poisoned allocations, initcheck) + merged-rows experiment.
cd "$(dirname "$0")/.."
O=gpurun_out
CB200_POISON=1 timeout 300 python tools/solve_check.py C4t 1 > $O/r02_poison_c4t.log 2>&1; tail -3 $O/r02_poison_c4t.log | cut -c1-200
CB200_POISON=1 timeout 300 python tools/solve_check.py C4r 1 > $O/r02_poison_c4r.log 2>&1; tail -3 $O/r02_poison_c4r.log | cut -c1-200
CB200_POISON=1 timeout 300 python tools/solve_check.py C4 1 > $O/r02_poison_c4.log 2>&1; tail -3 $O/r02_poison_c4.log | cut -c1-200
CB200_POISON=2 CB200_GRAPH=0 CB200_MAX_ITER=2 timeout 700 compute-sanitizer --tool initcheck --print-limit 60 python tools/solve_check.py C4t 1 > $O/r02_initcheck_c4t.log 2>&1; grep -c "Uninitialized" $O/r02_initcheck_c4t.log; tail -3 $O/r02_initcheck_c4t.log | cut -c1-200
CB200_POISON=1 timeout 900 python -m pytest tests/test_gpu_kkt.py tests/test_gpu_nonsym.py tests/test_gpu_updates.py -x -q -m gpu > $O/r02_pytest_poison.log 2>&1; tail -3 $O/r02_pytest_poison.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k c4 > $O/r02_pytest_c4_alone.log 2>&1; tail -3 $O/r02_pytest_c4_alone.log
for v in 0 1; do CB200_MERGED_ROWS=$v timeout 400 python tools/fine_breakdown.py C5 C3 > $O/r02_fine_merged$v.log 2>&1; grep -E "^==|fwd_warp|bwd_warp|fwd_cta|bwd_cta" $O/r02_fine_merged$v.log | cut -c1-160; done
CB200_MERGED_ROWS=1 timeout 600 python -m pytest tests/test_gpu_kkt.py -x -q -m gpu > $O/r02_pytest_merged.log 2>&1; tail -2 $O/r02_pytest_merged.log
