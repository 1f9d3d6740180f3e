#!/bin/bash
cd "$(dirname "$0")/.."
timeout 500 python tools/c4_repeat.py > gpurun_out/r02_c4_repeat.log 2>&1; cat gpurun_out/r02_c4_repeat.log | cut -c1-200
C4_IMPORT_ORACLE=0 C4_WARM=0 C4_REPEAT=1 timeout 200 python tools/c4_repeat.py > gpurun_out/r02_c4_repeat_b.log 2>&1; cat gpurun_out/r02_c4_repeat_b.log | cut -c1-200
