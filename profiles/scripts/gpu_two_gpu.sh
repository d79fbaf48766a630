set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_mg$N.json 2> gpurun_out/bench_mg$N.err
tail -c 3000 gpurun_out/bench_mg$N.json; tail -5 gpurun_out/bench_mg$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ref_mg$N.json 2> gpurun_out/bench_ref_mg$N.err
tail -c 1500 gpurun_out/bench_ref_mg$N.json; tail -3 gpurun_out/bench_ref_mg$N.err
