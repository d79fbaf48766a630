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
# per-phase timeline of the step kernel (profiling variant) and one full ncu capture of it
for w in battle1 gather64 battle512; do
  MAGENT_B200_LIB=$PWD/magent_b200/lib/variants/libmagent_timing.so timeout 200 python profiles/scripts/phase_timeline.py $w 2>&1 | tail -20
done | tee gpurun_out/phase_timeline.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel_cta -s 3 -c 1 -f -o gpurun_out/step_kernel_r1 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_step.log 2>&1
tail -n 2 gpurun_out/ncu_step.log | cut -c1-200
for w in battle1 gather64 battle1m; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_$w.json')); print('$w value %.3e ms/step %.4f obs_ms %.4f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))" || tail -3 gpurun_out/bench_$w.err
done
