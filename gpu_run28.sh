#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 2 -c 1 -f -o gpurun_out/obs_render_v12c python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_v12c.log 2>&1
tail -3 gpurun_out/ncu_v12c.log
ls -la gpurun_out/obs_render_v12c.ncu-rep
