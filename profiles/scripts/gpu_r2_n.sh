#!/bin/bash
# NUMA-split receive buffers: e2e with buffers and threads spread over both sockets vs bound to the GPU's node
cd /root/repo
o=gpurun_out/${1:-r2n}; mkdir -p $o
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "wire or graph or deferred" > $o/pytest_wire.log 2>&1; tail -2 $o/pytest_wire.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --no-cpu --steps 10 --warmup 3 $EXTRA > $o/$name.json 2> $o/$name.err; python -c "
import json; j=json.load(open('$o/$name.json')); e=j['e2e']; print('$name e2e %.3e steps %d threads %d numa: %s' % (e['value'], e['steps'], e['host_threads'], e['numa']))" || tail -3 $o/$name.err; }
EXTRA="" run split_default A=1
EXTRA="--numa-bind" run bound_to_gpu_node A=1
EXTRA="" run split_off MAGENT_B200_NUMA=off
EXTRA="" run split_t8 MAGENT_B200_HOST_THREADS=8
EXTRA="" run split_t24 MAGENT_B200_HOST_THREADS=24
EXTRA="" run split_t32 MAGENT_B200_HOST_THREADS=32
