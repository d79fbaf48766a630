#!/bin/bash
# four GPUs: bench at N=4, weak and strong (4096 arenas in total); checks the multi-rank path of the host threads / NUMA code
cd /root/repo
o=gpurun_out/${1:-r2r}; mkdir -p $o
nvidia-smi topo -m | head -6 | cut -c1-150
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 > $o/bench_4gpu.json 2> $o/bench_4gpu.err; tail -2 $o/bench_4gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 20 --warmup 5 --scaling strong > $o/bench_4gpu_strong4096.json 2> $o/bench_4gpu_strong4096.err; tail -2 $o/bench_4gpu_strong4096.err
python - <<PY
import json
for n in ("bench_4gpu", "bench_4gpu_strong4096"):
    try:
        j = json.loads(open("$o/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "n_gpus", j["n_gpus"], "value %.3e ms/step %.4f e2e %.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["scaling"], "threads/rank", j["e2e"].get("host_threads"), j["e2e"]["ms_per_step_by_phase"])
    except Exception as e: print(n, "failed", e)
PY
