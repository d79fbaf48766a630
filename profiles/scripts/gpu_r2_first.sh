#!/bin/bash
# round 2, first GPU call: host-side bandwidth probe (DMA vs host-thread expansion), the full GPU suite on the
# round-1 HEAD (tests/test_zz_fullsize_gpu.py never completed a GPU run), every bench workload, launch list.
cd /root/repo
o=gpurun_out/r2a; mkdir -p $o
timeout 600 profiles/probes/bin/probe_host_expand > $o/probe_host_expand.txt 2>&1; tail -40 $o/probe_host_expand.txt
( time timeout 1200 python -m pytest tests -q -m gpu -x --durations=15 2>&1 | tail -30 ) 2>&1 | tee $o/pytest_gpu.log
timeout 300 python bench.py > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -c 600 $o/bench_battle512.json
for w in battle1 gather64 battle1m battle1m_sparse; do
  timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > $o/bench_$w.json 2> $o/bench_$w.err
  python -c "
import json; j=json.load(open('$o/bench_$w.json')); print('$w value %.3e ms/step %.4f obs_ms %.4f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))" || tail -3 $o/bench_$w.err
done
ls -la $o
