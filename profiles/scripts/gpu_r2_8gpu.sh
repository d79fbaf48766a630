#!/bin/bash
# eight GPUs: one bench.py line at N=8 (weak scaling, the driver's SCALE configuration)
cd /root/repo
o=gpurun_out/${1:-r2n8}; mkdir -p $o
export MAGENT_B200_BENCH_RANK_REPORT=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 20 --warmup 5 > $o/bench_8gpu.json 2> $o/bench_8gpu.err; grep "^rank" $o/bench_8gpu.err | sort
python -c "
import json; j=json.loads(open('$o/bench_8gpu.json').read().strip().splitlines()[-1]); print('N=8 value %.3e ms/step %.4f e2e %.3e threads/rank %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['host_threads']), j['e2e']['ms_per_step_by_phase'])" || tail -20 $o/bench_8gpu.err
