#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
for v in base k3 k5 k6 c9 c10; do
  lib=$PWD/magent_b200/lib/variants/libmagent_$v.so; [ $v = base ] && lib=$PWD/magent_b200/lib/libmagent.so
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err
  python -c "
import json; j=json.load(open('gpurun_out/var_${v}.json')); print('VAR $v rep$rep value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/var_${v}.err
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 1 -f -o gpurun_out/obs_render_v13b python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_v13b.log 2>&1
