#!/bin/bash
# gpurun command file: measurements of HEAD after the step-kernel tuning commits (launch list, phase timeline,
# one full ncu capture of step_kernel_cta, the other workloads)
cd /root/repo
o=gpurun_out/r1e; mkdir -p $o
for w in battle512 battle1; do
  MAGENT_B200_LIB=$PWD/magent_b200/lib/variants/libmagent_timing.so timeout 100 python profiles/scripts/phase_timeline.py $w 2>&1 | tail -18
done | tee $o/phase_timeline.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_battle512_head.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > $o/ncu_launch.log 2>&1
tail -n 2 $o/ncu_launch.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel_cta -s 2 -c 1 -f -o $o/step_kernel_head python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > $o/ncu_step.log 2>&1
tail -n 2 $o/ncu_step.log | cut -c1-200
for w in battle1 gather64 battle1m; do
  timeout 120 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > $o/bench_$w.json 2> $o/bench_$w.err
  python -c "
import json; j=json.load(open('$o/bench_$w.json')); print('$w value %.3e ms/step %.4f obs_ms %.4f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))" || tail -3 $o/bench_$w.err
done
ls -la $o
