#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
for rep in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/b512_r45.json 2> gpurun_out/b512_r45.err; python -c "
import json; j=json.load(open('gpurun_out/b512_r45.json')); print('F32 value %.3e ms/step %.3f obs_ms %.3f frac %.3f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac']))"
done
python bench.py --obs-dtype f16 --steps 30 --warmup 5 --no-cpu > gpurun_out/f16_r45.json 2> gpurun_out/f16_r45.err; python -c "
import json; j=json.load(open('gpurun_out/f16_r45.json')); print('F16 value %.3e ms/step %.3f obs_ms %.3f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r45.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_r45.log 2>&1
