#!/bin/bash
tag=${1:-trip}
mkdir -p gpurun_out
(time python -m pytest tests -q -x -m gpu) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -6 gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_c2.json").read())
print("c2 fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["pass_ms"].items()})
print("sharded ms", (d.get("sharded") or {}).get("ms_per_frame"), (d.get("sharded") or {}).get("pass_ms_rank0"))
for k, v in (d.get("secondary") or {}).items(): print("secondary", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("value", "ms_per_step", "prism_fragments_per_frame", "tile_sv_ms", "prism_fragments_per_s")})
print("cpu_baseline", d.get("cpu_baseline"), "lit", d["run"]["lit_fraction"], "| shadow_pass", {k: v for k, v in (d.get("shadow_pass") or {}).items() if k in ("launch_ms", "hbm_frac", "l2_taps_frac")})
PY
python bench.py --workload c4_tree_sv --steps 40 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_c4.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c4_tree_sv_pertri --steps 40 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_c4p.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c4_tree_sv_1080p --steps 40 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_c4hd.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload c4_tree_sv_zfail --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_c4z.json 2>> gpurun_out/${tag}_bench.err
python -c "
import json
import glob
[print(f, 'fps %.1f' % json.loads(open(f).read())['value'], {k: round(v,4) for k,v in json.loads(open(f).read())['pass_ms'].items()}) for f in sorted(glob.glob('gpurun_out/'+'${tag}'+'_bench_c4*.json'))]"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_ncu_launch.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_launches_c2.csv | cut -c1-180
