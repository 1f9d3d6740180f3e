#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02_bench_C5_2gpu_final.json 2> gpurun_out/r02_bench_C5_2gpu_final.err; tail -c 400 gpurun_out/r02_bench_C5_2gpu_final.json; echo; tail -3 gpurun_out/r02_bench_C5_2gpu_final.err
