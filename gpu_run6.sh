set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_battle512.json 2> gpurun_out/bench_battle512.err; tail -c 3500 gpurun_out/bench_battle512.json; tail -5 gpurun_out/bench_battle512.err
timeout 600 python bench.py --workload battle1 --steps 200 --warmup 10 --no-cpu --no-e2e > gpurun_out/bench_battle1.json 2> gpurun_out/bench_battle1.err; tail -c 2500 gpurun_out/bench_battle1.json; tail -5 gpurun_out/bench_battle1.err
timeout 600 python bench.py --workload gather64 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_gather64.json 2> gpurun_out/bench_gather64.err; tail -c 2000 gpurun_out/bench_gather64.json; tail -5 gpurun_out/bench_gather64.err
timeout 900 python bench.py --workload battle1m --steps 5 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_battle1m.json 2> gpurun_out/bench_battle1m.err; tail -c 2000 gpurun_out/bench_battle1m.json; tail -5 gpurun_out/bench_battle1m.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_battle512.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 1 -o gpurun_out/obs_render_r1e python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel_cta -s 1 -c 1 -o gpurun_out/step_cta_r1b python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
