#!/bin/bash
cd /root/repo
o=gpurun_out/r2c; mkdir -p $o
timeout 1500 python -m pytest tests -v -m gpu -x --durations=8 > $o/pytest_gpu_full.log 2>&1
grep -n "PASSED\|FAILED\|ERROR" $o/pytest_gpu_full.log | tail -5
grep -n "FATAL\|Fatal Python\|CUDA error\|Segmentation\|Abort" -A12 $o/pytest_gpu_full.log | head -60
tail -5 $o/pytest_gpu_full.log
