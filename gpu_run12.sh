set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_battle512.json 2> gpurun_out/bench_battle512.err; python -c "
import json; j=json.load(open('gpurun_out/bench_battle512.json')); print('battle512 value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"; tail -3 gpurun_out/bench_battle512.err
timeout 900 python bench.py --workload gather64 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_gather64.json 2> gpurun_out/bench_gather64.err; python -c "
import json; j=json.load(open('gpurun_out/bench_gather64.json')); print('gather64 value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"
timeout 900 python bench.py --workload battle1m --steps 5 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_battle1m.json 2> gpurun_out/bench_battle1m.err; python -c "
import json; j=json.load(open('gpurun_out/bench_battle1m.json')); print('battle1m value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 1 -o gpurun_out/obs_render_r1j python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
