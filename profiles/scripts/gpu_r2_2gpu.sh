#!/bin/bash
# two GPUs: engines on two devices in one process; bench at N=2 (weak, then strong scaling of 1024 arenas)
cd /root/repo
o=gpurun_out/${1:-r2m}; mkdir -p $o
nvidia-smi topo -m | head -6
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x > $o/pytest_multigpu.log 2>&1; tail -3 $o/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > $o/bench_2gpu.json 2> $o/bench_2gpu.err; tail -2 $o/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --scaling strong --arenas 1024 > $o/bench_2gpu_strong1024.json 2> $o/bench_2gpu_strong1024.err; tail -2 $o/bench_2gpu_strong1024.err
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --scaling strong --arenas 1024 --no-cpu > $o/bench_1gpu_strong1024.json 2> $o/bench_1gpu_strong1024.err
python - <<PY
import json
for n in ("bench_2gpu", "bench_2gpu_strong1024", "bench_1gpu_strong1024"):
    try:
        j = json.loads(open("$o/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "n_gpus", j["n_gpus"], "value %.3e ms/step %.4f e2e %.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["scaling"], j["e2e"].get("numa"), "threads", j["e2e"].get("host_threads"))
    except Exception as e: print(n, "failed", e)
PY
