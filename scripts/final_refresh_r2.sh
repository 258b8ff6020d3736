#!/bin/bash
# refresh of the round-2 evidence after the last change (k_order without spills): GPU tests, the default bench line, launch lists
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu) > gpurun_out/r2_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_gpu_tests.log | head -1
python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/r2_launches_c5_sharded.csv python bench.py --sharded-only > gpurun_out/r2_ncu_launch_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --workload c4_tree_sv --steps 4 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_launch_c4.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_launches_c2.csv | cut -c1-150
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c2.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['e2e']['value'], d['sharded']['ms_per_frame'], {k: v['value'] for k, v in d['secondary'].items()})"
