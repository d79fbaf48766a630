#!/bin/bash
# A/B of engine library variants: gpu_ab.sh <outdir> <variant> [<variant> ...]   ("base" = magent_b200/lib/libmagent.so)
cd /root/repo
o=gpurun_out/$1; shift; mkdir -p $o
for rep in 1 2; do
for v in "$@"; do
  lib=$PWD/magent_b200/lib/variants/libmagent_$v.so; [ $v = base ] && lib=$PWD/magent_b200/lib/libmagent.so
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $o/var_${v}.json 2> $o/var_${v}.err
  python -c "
import json; j=json.load(open('$o/var_${v}.json')); print('VAR $v rep$rep value %.3e ms/step %.4f obs_ms %.4f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 $o/var_${v}.err
done
done
