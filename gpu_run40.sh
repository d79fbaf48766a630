#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in b d; do
  lib=$PWD/magent_b200/lib/variants/libmagent_$v.so
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err
  python -c "
import json; j=json.load(open('gpurun_out/var_${v}.json')); print('VAR $v rep$rep value %.3e ms/step %.3f obs_ms %.3f frac %.3f clocks %s'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['clocks']))" || tail -3 gpurun_out/var_${v}.err
done
done
nvidia-smi --query-gpu=name,clocks.mem,clocks.max.mem,clocks.sm,power.limit --format=csv
