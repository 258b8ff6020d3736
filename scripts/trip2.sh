#!/bin/bash
tag=${1:-trip}
mkdir -p gpurun_out
(time python -m pytest tests -q -x -m gpu) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -4 gpurun_out/${tag}_gpu_tests.log
python scripts/ab_probe.py 2>&1 | tee gpurun_out/${tag}_ab.txt
python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
python bench.py --workload c5_many_light --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2>> gpurun_out/${tag}_bench.err
for f in c2 c5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$f.json").read())
    print("$f", "fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["pass_ms"].items()}, "frac", d["roofline"] and round(d["roofline"]["frac"],3))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/${tag}_launches_c5.csv python bench.py --workload c5_many_light --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch_c5.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_launches_c2.csv gpurun_out/${tag}_launches_c5.csv | cut -c1-180
ncu --set full --clock-control none --import-source on -k regex:"k_setup_bin|k_order" -s 8 -c 4 -o gpurun_out/${tag}_prof_bin -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full_bin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_order" -s 30 -c 3 -o gpurun_out/${tag}_prof_c5 -f python bench.py --workload c5_many_light --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/${tag}_bench.err
