#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
python bench.py --obs-dtype f16 --steps 30 --warmup 5 --no-cpu > gpurun_out/f16_r44.json 2> gpurun_out/f16_r44.err; python -c "
import json; j=json.load(open('gpurun_out/f16_r44.json')); print('F16 value %.3e ms/step %.3f obs_ms %.3f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))"
