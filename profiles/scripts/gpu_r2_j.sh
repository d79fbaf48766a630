#!/bin/bash
# full GPU suite on the tuned kernels, bench (all workloads), traffic capture, compute-sanitizer logs
cd /root/repo
o=gpurun_out/${1:-r2j}; mkdir -p $o
timeout 1500 python -m pytest tests -q -m gpu -x > $o/pytest_gpu.log 2>&1; tail -3 $o/pytest_gpu.log
timeout 400 python bench.py > $o/bench_battle512.json 2> $o/bench_battle512.err; tail -3 $o/bench_battle512.err
timeout 300 python bench.py --impl reference > $o/bench_reference.json 2> $o/bench_reference.err
timeout 900 python bench.py --workload all --steps 30 --warmup 5 --no-cpu > $o/bench_all.jsonl 2> $o/bench_all.err
python - <<PY
import json
for f in ("$o/bench_battle512.json", "$o/bench_reference.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.3e ms/step %.4f" % (j["value"], j["ms_per_step"]), "e2e %.3e" % j["e2e"]["value"], "frac", (j.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "failed", e)
for line in open("$o/bench_all.jsonl"):
    try:
        j = json.loads(line); print(j["config"]["workload"][:40], "value %.3e ms/step %.4f render ms %.4f frac %.3f e2e %.3e" % (j["value"], j["ms_per_step"], j["roofline"]["mean_launch_ms"], j["roofline"]["frac"], j["e2e"]["value"]))
    except Exception as e: print("line failed", e, line[:200])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:obs_render -s 4 -c 1 -f -o $o/obs_render_r2 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > $o/ncu_obs_render.log 2>&1; ls -la $o/obs_render_r2.ncu-rep
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/scripts/sanitize_micro.py > $o/sanitizer_$tool.log 2>&1; tail -4 $o/sanitizer_$tool.log
done
