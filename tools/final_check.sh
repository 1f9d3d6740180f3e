#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_kkt.py -x -q -m gpu -k "resident or inner_boundary" > gpurun_out/r02_pytest_resident.log 2>&1; tail -3 gpurun_out/r02_pytest_resident.log
timeout 200 python bench.py --workload C3 --steps 6 > gpurun_out/r02_bench_C3_check.json 2> gpurun_out/r02_bench_C3_check.err; python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_C3_check.json") if l.startswith("{")][-1])
print("C3", d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["parity"].get("resident_vs_host_rel_diff"))
P
tail -2 gpurun_out/r02_bench_C3_check.err
