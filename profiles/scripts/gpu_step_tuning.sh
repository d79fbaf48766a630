cd /root/repo; mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu.log
for w in battle512 battle1; do MAGENT_B200_LIB=$PWD/magent_b200/lib/variants/libmagent_timing.so timeout 200 python profiles/scripts/phase_timeline.py $w 2>&1 | tail -17; done | tee gpurun_out/phase_timeline2.txt
for rep in 1 2; do python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/b512_$rep.json 2>gpurun_out/b512.err; python -c "
import json; j=json.load(open('gpurun_out/b512_$rep.json')); print('b512 value %.4e ms/step %.4f obs_ms %.4f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"; done
python bench.py --workload battle1 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/b1.json 2>gpurun_out/b1.err; python -c "
import json; j=json.load(open('gpurun_out/b1.json')); print('battle1 value %.4e ms/step %.4f'%(j['value'], j['ms_per_step']))"
