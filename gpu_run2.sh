set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|^CPU\(s\)" 
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_battle512.json 2> gpurun_out/bench_battle512.err; tail -c 3000 gpurun_out/bench_battle512.json; tail -5 gpurun_out/bench_battle512.err
timeout 600 python bench.py --workload battle1 --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_battle1.json 2> gpurun_out/bench_battle1.err; tail -c 2500 gpurun_out/bench_battle1.json; tail -5 gpurun_out/bench_battle1.err
# launch list (shares only) for the default workload, short run
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_battle512.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --arenas 128 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
# full capture of the top kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 2 -o gpurun_out/obs_render_r1 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --arenas 128 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
