#!/bin/bash
# round 2: first run of the wire host path on the GPU -- parity suite, then the bench lines
cd /root/repo
o=gpurun_out/r2b; mkdir -p $o
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -25 ) 2>&1 | tee $o/pytest_gpu.log
timeout 400 python bench.py > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -c 1500 $o/bench_battle512.json; tail -5 $o/bench_battle512.err
timeout 300 python bench.py --host-path dense --no-cpu > $o/bench_battle512_dense.json 2> $o/bench_battle512_dense.err; python - <<'PY'
import json
for n in ("bench_battle512", "bench_battle512_dense"):
    try:
        j = json.load(open("gpurun_out/r2b/%s.json" % n)); print(n, "value %.3e ms/step %.4f frac %.3f e2e %.3e" % (j["value"], j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"]), j["e2e"])
    except Exception as e: print(n, "failed", e)
PY
for t in 4 8 12; do MAGENT_B200_HOST_THREADS=$t timeout 200 python bench.py --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $t e2e %.3e' % j['e2e']['value'])"; done
