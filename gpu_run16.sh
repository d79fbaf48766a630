set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for pf in 0 4 8 16 -1; do
  if [ $pf = -1 ]; then unset MAGENT_B200_OBS_PF; else export MAGENT_B200_OBS_PF=$pf; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/pf_$pf.json 2> gpurun_out/pf_$pf.err
  python -c "
import json; j=json.load(open('gpurun_out/pf_$pf.json')); print('PF=$pf value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))" || tail -3 gpurun_out/pf_$pf.err
done
