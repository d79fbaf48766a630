#!/bin/bash
# gpurun command file for the first GPU call of the next round: everything that was queued at the end of round 1
# (profiles/README.md "Open measurements").  ~6-8 minutes of box time on one GPU.
#   gpurun --timeout 900 -- 'bash profiles/scripts/gpu_next_round_first_call.sh'
cd /root/repo
o=gpurun_out/r2a; mkdir -p $o
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=15 2>&1 | tail -30 ) 2>&1 | tee $o/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/smoke.log
timeout 300 python bench.py > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -c 600 $o/bench_battle512.json
timeout 300 python bench.py --impl reference > $o/bench_reference.json 2> $o/bench_reference.err; tail -c 400 $o/bench_reference.json
for w in battle512_blocks battle1 gather64 battle1m battle1m_sparse; do
  timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > $o/bench_$w.json 2> $o/bench_$w.err
  python -c "
import json; j=json.load(open('$o/bench_$w.json')); print('$w value %.3e ms/step %.4f obs_ms %.4f frac %.3f e2e %.3e'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms'], j['roofline']['frac'], j['e2e']['value']))" || tail -3 $o/bench_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_battle512.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > $o/ncu_launch.log 2>&1
tail -n 2 $o/ncu_launch.log | cut -c1-200
ls -la $o
