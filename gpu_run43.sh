#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/b512_r43.json 2> gpurun_out/b512_r43.err; cut -c1-400 gpurun_out/b512_r43.json
