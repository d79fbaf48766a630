#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r1c
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r1c/bench_battle512.json 2> gpurun_out/r1c/bench_battle512.err; tail -c 3000 gpurun_out/r1c/bench_battle512.json
timeout 600 python bench.py --impl reference > gpurun_out/r1c/bench_reference.json 2> gpurun_out/r1c/bench_reference.err; tail -c 800 gpurun_out/r1c/bench_reference.json
timeout 600 python bench.py --workload battle1 --steps 300 --warmup 20 > gpurun_out/r1c/bench_battle1.json 2> gpurun_out/r1c/bench_battle1.err; tail -c 300 gpurun_out/r1c/bench_battle1.json
timeout 600 python bench.py --workload gather64 --steps 100 --warmup 10 > gpurun_out/r1c/bench_gather64.json 2> gpurun_out/r1c/bench_gather64.err; tail -c 300 gpurun_out/r1c/bench_gather64.json
timeout 900 python bench.py --workload battle1m --steps 10 --warmup 3 --no-cpu > gpurun_out/r1c/bench_battle1m.json 2> gpurun_out/r1c/bench_battle1m.err; tail -c 300 gpurun_out/r1c/bench_battle1m.json
timeout 900 python bench.py --obs-dtype f16 --steps 30 --warmup 5 --no-cpu > gpurun_out/r1c/bench_battle512_f16.json 2> gpurun_out/r1c/bench_battle512_f16.err; tail -c 300 gpurun_out/r1c/bench_battle512_f16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c/launches_battle512.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/r1c/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 1 -f -o gpurun_out/r1c/obs_render_v13 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r1c/ncu_obs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel_cta -s 1 -c 1 -f -o gpurun_out/r1c/step_cta_v13 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r1c/ncu_step.log 2>&1
ls -la gpurun_out/r1c
