#!/bin/bash
# gpurun command file: full GPU parity suite, smoke, default bench line, launch list (round 1, after the general rule binder)
cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 2500 gpurun_out/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_rules.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -n 3 gpurun_out/ncu_launch.log | cut -c1-300
