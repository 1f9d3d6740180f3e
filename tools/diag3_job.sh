#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism.log; : > $O
run() { env "$@" timeout 300 python tools/determinism.py $W 6 2>&1 | tail -1 >> $O; tail -1 $O | cut -c1-250; }
W=C4r; run A=1; run CB200_NO_TMA=1; run CB200_MULTISTREAM=0; run CB200_GRAPH=0; run CB200_GRAPH=0 CB200_MULTISTREAM=0; run CB200_GRAPH=0 CB200_MULTISTREAM=0 CB200_NO_TMA=1
W=C1; run A=1
W=C3; run A=1
W=C5; run A=1
W=C4t; CB200_GRAPH=0 CB200_MAX_ITER=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 40 python tools/determinism.py C4t 2 > gpurun_out/r02_racecheck_c4t.log 2>&1; grep -c "Race reported\|hazard" gpurun_out/r02_racecheck_c4t.log; tail -2 gpurun_out/r02_racecheck_c4t.log | cut -c1-200
