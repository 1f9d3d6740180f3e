#!/bin/bash
# Round-end ncu evidence (1 GPU): launch list of one C5 step with DRAM bytes (-> profiles/r02_traffic.json)
# and a --set full capture of the shipped TMA GEMM (Schur launches of C4: big K).
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1250 -c 1400 --csv \
    --log-file gpurun_out/r02_launches_c5_step_final.csv python tools/profile_step.py C5 2 > gpurun_out/r02_ncu_list_final.log 2>&1
tail -1 gpurun_out/r02_ncu_list_final.log | cut -c1-80
ncu --set full --clock-control none --import-source on -k regex:k_ldl_update_tma -s 300 -c 4 -o gpurun_out/r02_tma_gemm_final \
    python tools/profile_step.py C4 1 > gpurun_out/r02_ncu_tma_final.log 2>&1
tail -1 gpurun_out/r02_ncu_tma_final.log | cut -c1-80
