#!/bin/bash
cd /root/repo
o=gpurun_out/${1:-r2t}; mkdir -p $o
export MAGENT_B200_BENCH_RANK_REPORT=1
nvidia-smi topo -m | head -5 | cut -c1-60,100-150
for rep in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2964$rep bench.py --gpus 4 --steps 20 --warmup 5 > $o/bench_4gpu_$rep.json 2> $o/bench_4gpu_$rep.err; grep "^rank" $o/bench_4gpu_$rep.err
python -c "
import json; j=json.loads(open('$o/bench_4gpu_$rep.json').read().strip().splitlines()[-1]); print('rep $rep value %.3e ms/step %.4f e2e %.3e' % (j['value'], j['ms_per_step'], j['e2e']['value']), j['e2e']['ms_per_step_by_phase'])"
done
