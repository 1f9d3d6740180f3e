#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism2.log; : > $O
run() { env "$@" timeout 300 python tools/determinism.py $W 40 2>&1 | grep -v "^$" | tail -6 >> $O; tail -1 $O | cut -c1-250; }
W=C4r; run A=1; run CB200_NO_TMA=1; run CB200_MULTISTREAM=0; run CB200_GRAPH=0; run CB200_GRAPH=1; run CB200_GRAPH=2; run CB200_TMA_TILE=128; run CB200_SORT_BATCHES=0
W=C4; env A=1 timeout 400 python tools/determinism.py C4 12 2>&1 | tail -6 >> $O; tail -1 $O | cut -c1-250
