#!/bin/bash
# gpurun command file: A/B of the step kernel's CTA width on the default workload (512 arenas of 2x1000 agents; the
# launch picks 512 threads = 2 CTAs per SM today) -- MAGENT_B200_STEP_THREADS is a measurement knob of launch_step.
# Each setting first replays a few parity tests (the phase functions must not care about the team width).
cd /root/repo
o=gpurun_out/step_threads; mkdir -p $o
for t in 256 384 512 640 768 1024; do
  export MAGENT_B200_STEP_THREADS=$t
  timeout 200 python -m pytest tests/test_parity_gpu.py -q -x -k "battle_small or kills or bands or pursuit or double_attack" 2>&1 | tail -1
  timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu > $o/b512_$t.json 2> $o/b512_$t.err
  python -c "
import json; j=json.load(open('$o/b512_$t.json')); print('threads $t value %.4e ms/step %.4f obs_ms %.4f'%(j['value'], j['ms_per_step'], j['roofline']['mean_launch_ms']))" || tail -3 $o/b512_$t.err
done 2>&1 | tee $o/summary.txt
