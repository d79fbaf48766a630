#!/bin/bash
# GPU suite + bench + launch list (shares of one step)
cd /root/repo
o=gpurun_out/${1:-r2d}; mkdir -p $o
timeout 1500 python -m pytest tests -q -m gpu -x > $o/pytest_gpu.log 2>&1; tail -3 $o/pytest_gpu.log
timeout 400 python bench.py --no-cpu > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -3 $o/bench_battle512.err
python - <<PY
import json
j = json.load(open("$o/bench_battle512.json")); print("value %.3e ms/step %.4f render ms %.4f frac %.3f e2e %.3e launches %d" % (j["value"], j["ms_per_step"], j["roofline"]["mean_launch_ms"], j["roofline"]["frac"], j["e2e"]["value"], j["gpu_launches"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_battle512.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > $o/ncu_launch.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$o/launches_battle512.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0]; v = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
unit = rows[0][-2] if rows else "?"
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-40s n=%4d avg %10.1f %s share %.3f" % (k[-40:], v[0], v[1] / v[0], unit, v[1] / tot))
PY
