#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python tests/debug_diff.py > gpurun_out/debug_diff.txt 2>&1
cat gpurun_out/debug_diff.txt
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15
