set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --no-cpu > gpurun_out/b512.json 2> gpurun_out/b512.err; python -c "
import json; j=json.load(open('gpurun_out/b512.json')); print('battle512 value %.3e ms/step %.3f obs_ms %.3f frac %.3f e2e %.3e traffic %s'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value'], j['roofline']['traffic']))"; tail -3 gpurun_out/b512.err
