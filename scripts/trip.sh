#!/bin/bash
# developer helper for one gpurun trip: parity tests, then bench lines of the main workloads and a launch list; tag = $1
tag=${1:-trip}
mkdir -p gpurun_out
(time python -m pytest tests -q -x -m gpu) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -4 gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
python bench.py --workload c5_many_light --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c3_dragon --steps 200 --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c4_tree_sv --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c1_teapot --steps 300 --no-cpu-baseline > gpurun_out/${tag}_bench_c1.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload dragon_pcss --steps 200 --no-cpu-baseline > gpurun_out/${tag}_bench_dragon_pcss.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/${tag}_launches_c5.csv python bench.py --workload c5_many_light --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch_c5.log 2>&1
tail -5 gpurun_out/${tag}_bench.err
for f in c2 c5 c3 c4 c1 dragon_pcss; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$f.json").read())
    print("$f", "fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["pass_ms"].items()}, "frac", d["roofline"] and round(d["roofline"]["frac"],3))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
python scripts/ncu_summary.py gpurun_out/${tag}_launches_c2.csv gpurun_out/${tag}_launches_c5.csv | cut -c1-180
