#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism6.log; : > $O
V=$PWD/clarabel.jl_b200/_variants/libvariant_nofence.so
run() { env "$@" timeout 240 python tools/determinism.py C4r 120 2>&1 | grep -v "^$" | tail -8 >> $O; tail -1 $O | cut -c1-330; }
run CB200_LIB_PATH=$V CB200_TMA_ATTR_KB=64
run CB200_TMA_ATTR_KB=64
run CB200_LIB_PATH=$V
