set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for m in 0 1; do
MAGENT_B200_STEP_SMEM=$m timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_smem$m.json 2> gpurun_out/bench_smem$m.err; python -c "
import json; j=json.load(open('gpurun_out/bench_smem$m.json')); print('SMEM=$m value %.3e ms/step %.3f obs_ms %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms']))"
done
MAGENT_B200_STEP_SMEM=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -c 4 --csv --log-file gpurun_out/step_smem0.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
MAGENT_B200_STEP_SMEM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -c 4 --csv --log-file gpurun_out/step_smem1.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
grep step_kernel gpurun_out/step_smem0.csv | cut -d, -f5,12- ; grep step_kernel gpurun_out/step_smem1.csv | cut -d, -f5,12-
timeout 600 python bench.py --workload battle1 --steps 200 --warmup 10 --no-cpu --no-e2e > gpurun_out/bench_battle1.json 2> gpurun_out/bench_battle1.err; python -c "
import json; j=json.load(open('gpurun_out/bench_battle1.json')); print('battle1 value %.3e ms/step %.3f'%(j['value'], j['ms_per_step']))"
