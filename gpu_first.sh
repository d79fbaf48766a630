set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
cd $GRAFT_REPO_ROOT
ls oracle/_ref/ magent_b200/lib/*.so
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -30
