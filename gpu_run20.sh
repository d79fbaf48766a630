set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in base n4c8 n4c9 n4c10 n4c11 n4p0c10 n4p0c11; do
  lib=$PWD/magent_b200/lib/variants/libmagent_$v.so; [ $v = base ] && lib=$PWD/magent_b200/lib/libmagent.so
  MAGENT_B200_LIB=$lib timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err
  python -c "
import json; j=json.load(open('gpurun_out/var_${v}.json')); print('VAR $v rep$rep value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/var_${v}.err
done
done
