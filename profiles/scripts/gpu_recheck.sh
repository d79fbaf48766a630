#!/bin/bash
# gpurun command file: short re-validation of HEAD after a container restart (parity suite in 8 worker
# processes, default bench line without the CPU leg, smoke)
cd /root/repo
mkdir -p gpurun_out/r1d
( time timeout 330 python -m pytest tests -q -m gpu -x -n 8 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/r1d/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r1d/bench_battle512.json 2> gpurun_out/r1d/bench_battle512.err
tail -c 1500 gpurun_out/r1d/bench_battle512.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r1d/smoke.log
