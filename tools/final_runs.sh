#!/bin/bash
# Round-end measurement sequence (1 GPU): full -m gpu test suite, bench lines of every BASELINE
# config (C5 with the full-size CPU leg), the reference arm.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_final.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_C5_final.json 2> gpurun_out/r02_bench_C5_final.err; tail -c 500 gpurun_out/r02_bench_C5_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_C5_reference.json 2> gpurun_out/r02_bench_C5_reference.err; tail -c 300 gpurun_out/r02_bench_C5_reference.json; echo
for w in C1 C2 C3; do python bench.py --workload $w --steps 8 > gpurun_out/r02_bench_${w}_final.json 2> gpurun_out/r02_bench_${w}_final.err; tail -c 250 gpurun_out/r02_bench_${w}_final.json; echo; done
python bench.py --workload C4 --steps 6 --no-cpu-baseline > gpurun_out/r02_bench_C4_final.json 2> gpurun_out/r02_bench_C4_final.err; tail -c 250 gpurun_out/r02_bench_C4_final.json; echo
