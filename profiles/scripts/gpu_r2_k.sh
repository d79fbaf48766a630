#!/bin/bash
cd /root/repo
o=gpurun_out/${1:-r2k}; mkdir -p $o
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x > $o/pytest_parity.log 2>&1; tail -3 $o/pytest_parity.log
for w in battle1 gather64; do
  for g in off on; do
    timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu --graph $g > $o/bench_${w}_graph_$g.json 2> $o/bench_${w}_graph_$g.err
    python -c "
import json; j=json.load(open('$o/bench_${w}_graph_$g.json')); print('$w graph=$g value %.3e ms/step %.4f render ms %.4f e2e %.3e launches %d'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['e2e']['value'], j['gpu_launches']))" || tail -5 $o/bench_${w}_graph_$g.err
  done
done
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 50 python profiles/scripts/sanitize_micro.py > $o/sanitizer_$tool.log 2>&1; grep "scenario\|SUMMARY" $o/sanitizer_$tool.log
done
