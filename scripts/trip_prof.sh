#!/bin/bash
# profile one workload: launch list + full capture of the top kernels; $1 = workload, $2 = tag, $3 = kernel regex, $4 = skip count
w=${1:-c5_sandiego}; tag=${2:-prof}; kre=${3:-"k_tile|k_setup_bin"}; skip=${4:-20}
mkdir -p gpurun_out
python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_$w.json').read())
print('$w fps %.1f' % d['value'], {k: round(v,4) for k,v in d['pass_ms'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 150 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_ncu_launch.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_launches_$w.csv | cut -c1-180
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 4 -o gpurun_out/${tag}_prof_$w -f python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_prof_$w.ncu-rep
