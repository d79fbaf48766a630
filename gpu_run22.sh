#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 > gpurun_out/b512_r22.json 2> gpurun_out/b512_r22.err
cat gpurun_out/b512_r22.json
