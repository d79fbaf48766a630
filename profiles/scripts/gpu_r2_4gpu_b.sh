#!/bin/bash
cd /root/repo
o=gpurun_out/${1:-r2s}; mkdir -p $o
export MAGENT_B200_BENCH_RANK_REPORT=1
for steps in 20 100; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2963$((steps/20)) bench.py --gpus 4 --steps $steps --warmup 5 --no-e2e > $o/bench_4gpu_$steps.json 2> $o/bench_4gpu_$steps.err; grep "^rank" $o/bench_4gpu_$steps.err
python -c "
import json; j=json.loads(open('$o/bench_4gpu_$steps.json').read().strip().splitlines()[-1]); print('steps $steps value %.3e ms/step %.4f' % (j['value'], j['ms_per_step']))"
done
# the same four GPUs one at a time (is one of them slower on its own?)
for d in 0 1 2 3; do CUDA_VISIBLE_DEVICES=$d timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu 2> $o/single_$d.err | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gpu $d alone ms/step %.4f render %.4f' % (j['ms_per_step'], j['roofline']['mean_launch_ms']))"; done
