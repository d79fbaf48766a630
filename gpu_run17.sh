set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-cpu --no-e2e > gpurun_out/b512.json 2> gpurun_out/b512.err; python -c "
import json; j=json.load(open('gpurun_out/b512.json')); print('battle512 value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"; tail -3 gpurun_out/b512.err
timeout 600 python bench.py --workload battle1 --steps 300 --warmup 20 --no-e2e > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json; j=json.load(open('gpurun_out/b1.json')); print('battle1 value %.3e ms/step %.3f obs_ms %.4f cpu %s'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['cpu_baseline'] and (j['cpu_baseline']['value'], j['cpu_baseline']['sample'][-120:])))"
timeout 600 python bench.py --workload gather64 --steps 100 --warmup 10 --no-cpu --no-e2e > gpurun_out/g64.json 2> gpurun_out/g64.err; python -c "
import json; j=json.load(open('gpurun_out/g64.json')); print('gather64 value %.3e ms/step %.3f obs_ms %.4f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_battle1.csv python bench.py --workload battle1 --steps 5 --warmup 2 --no-e2e --no-cpu > /dev/null 2>&1
