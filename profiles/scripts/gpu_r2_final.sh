#!/bin/bash
# round 2, final pass on one GPU: whole GPU suite, smoke, bench (both arms, every workload), launch list, ncu --set full
# captures of the render / step / wire kernels (raw reports kept under profiles/r2/)
cd /root/repo
o=gpurun_out/${1:-r2z}; mkdir -p $o
( time timeout 1800 python -m pytest tests -q -m gpu -x --durations=10 ) > $o/pytest_gpu.log 2>&1; tail -16 $o/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $o/smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/bench_reference.json 2> $o/bench_reference.err
timeout 400 python bench.py --steps 20 --warmup 5 > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -2 $o/bench_battle512.err
timeout 900 python bench.py --workload all --steps 40 --warmup 5 --no-cpu > $o/bench_all.jsonl 2> $o/bench_all.err
python - <<PY
import json
r = json.loads(open("$o/bench_reference.json").read().strip().splitlines()[-1])
j = json.loads(open("$o/bench_battle512.json").read().strip().splitlines()[-1])
print("reference %.3e (%s cores)  b200 value %.3e ms/step %.4f frac %.3f  e2e %.3e  ratio e2e/ref %.2f" % (r["value"], r["cpu_baseline"]["cores"], j["value"], j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"], j["e2e"]["value"] / r["value"]))
print(j["e2e"]["ms_per_step_by_phase"], j["e2e"]["numa"])
for line in open("$o/bench_all.jsonl"):
    try:
        j = json.loads(line); print(j["config"]["workload"][:44], "| value %.3e ms/step %.4f render ms %.4f frac %.3f e2e %.3e | %s" % (j["value"], j["ms_per_step"], j["roofline"]["mean_launch_ms"], j["roofline"]["frac"], j["e2e"]["value"], j["config"]["launch"][:30]))
    except Exception as e: print("line failed", e, line[:200])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_battle512.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > $o/ncu_launch.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$o/launches_battle512.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0]; v = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-40s n=%4d avg %10.1f ns share %.3f" % (k[-40:], v[0], v[1] / v[0], v[1] / tot))
PY
for k in obs_render step_kernel_cta; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $o/${k}_r2 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > $o/ncu_$k.log 2>&1
  ls -la $o/${k}_r2.ncu-rep
done
