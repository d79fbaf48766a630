set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in c10 c11 t8c5 t8c4; do
  for w in battle512 battle1m; do
    st=20; [ $w = battle1m ] && st=5
    MAGENT_B200_LIB=$PWD/magent_b200/lib/variants/libmagent_$v.so timeout 600 python bench.py --workload $w --steps $st --warmup 3 --no-cpu --no-e2e > gpurun_out/var_${v}_$w.json 2> gpurun_out/var_${v}_$w.err
    python -c "
import json; j=json.load(open('gpurun_out/var_${v}_$w.json')); print('VAR $v $w value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/var_${v}_$w.err
  done
done
