#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism4.log; : > $O
run() { env DET_SKIP_L=1 DET_SLEEP=0.35 "$@" timeout 200 python tools/determinism.py C4r 70 2>&1 | grep -v "^$" | tail -12 >> $O; tail -1 $O | cut -c1-300; }
run A=1; run CB200_NO_TMA=1; run CB200_TMA_TILE=128; run CB200_TMA_FENCE=1; run CB200_TMA_PAD_KB=100; run CB200_GRAPH=0 CB200_MULTISTREAM=0; run CB200_NO_TMA=1 CB200_GRAPH=0 CB200_MULTISTREAM=0
