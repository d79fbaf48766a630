#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 ./profiles/probes/bin/probe_tma_store > gpurun_out/probe_tma_store.txt 2>&1
cat gpurun_out/probe_tma_store.txt
timeout 900 python -m pytest tests -q -m gpu -k "f16" 2>&1 | tail -8
