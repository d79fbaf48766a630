#!/bin/bash
# e2e phase breakdown + launch list of the host-buffer loop + 1 M sparse geometry parity
cd /root/repo
o=gpurun_out/${1:-r2o}; mkdir -p $o
timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $o/bench.json 2> $o/bench.err; python -c "
import json; j=json.load(open('$o/bench.json')); e=j['e2e']; print('e2e %.3e' % e['value'], e['ms_per_step_by_phase'])"
timeout 600 python -m pytest tests/test_zz_fullsize_gpu.py -q -m gpu -x -k "reference_1m_geometry or 2x400k" > $o/pytest_1m.log 2>&1; tail -3 $o/pytest_1m.log
cat > /tmp/e2e_loop.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, magent_b200 as magent
env = magent.GridWorld("battle", map_size=200, _num_arenas=512); env.reset()
hs = env.get_handles()
for h in hs: env.add_agents(h, method="random", n=1000)
acts = [np.random.randint(0, 21, size=env.get_num(h)).astype(np.int32) for h in hs]
for t in range(4):
    for h in hs: env.get_observation(h)
    for h, a in zip(hs, acts): env.set_action(h, a[:env.get_num(h)])
    env.step()
    for h in hs: env.get_reward(h)
    env.clear_dead()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_e2e.csv python /tmp/e2e_loop.py > $o/ncu_e2e.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$o/launches_e2e.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0]; v = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-40s n=%4d avg %10.1f ns" % (k[-40:], v[0], v[1] / v[0]))
PY
