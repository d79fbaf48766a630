#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
for rep in 1 2; do
for v in base k1 k2 k8 k16; do
  lib=$PWD/magent_b200/lib/variants/libmagent_$v.so; [ $v = base ] && lib=$PWD/magent_b200/lib/libmagent.so
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err
  python -c "
import json; j=json.load(open('gpurun_out/var_${v}.json')); print('VAR $v rep$rep value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/var_${v}.err
done
done
for v in base k8; do
lib=$PWD/magent_b200/lib/variants/libmagent_$v.so; [ $v = base ] && lib=$PWD/magent_b200/lib/libmagent.so
for w in gather64 battle1m battle1; do
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/v13_$w.json 2> gpurun_out/v13_$w.err
  python -c "
import json; j=json.load(open('gpurun_out/v13_$w.json')); print('WL $v $w value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/v13_$w.err
done
MAGENT_B200_LIB=$lib timeout 600 python bench.py --obs-dtype f16 --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/v13_f16.json 2> gpurun_out/v13_f16.err
python -c "
import json; j=json.load(open('gpurun_out/v13_f16.json')); print('WL $v f16 value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"
done
