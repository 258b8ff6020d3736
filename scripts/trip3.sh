#!/bin/bash
tag=${1:-trip}
mkdir -p gpurun_out
(time python -m pytest tests -q -x -m gpu) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -6 gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_c2.json").read())
print("c2 fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["pass_ms"].items()})
print("sharded", json.dumps(d.get("sharded"), indent=0)[:1500])
PY
python bench.py --workload c5_many_light --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2>> gpurun_out/${tag}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_c5.json').read())
print('c5 unfused fps %.1f' % d['value'], {k: round(v,4) for k,v in d['pass_ms'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_ncu_launch.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_launches_c2.csv | cut -c1-180
