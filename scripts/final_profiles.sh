#!/bin/bash
# round-end evidence: bench JSON lines, ncu launch list of the bench command, one --set full capture (c2 workload)
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_c2.json 2> gpurun_out/r1_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2>> gpurun_out/r1_bench_c2.err
python bench.py --workload c5_many_light --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_c5.json 2>> gpurun_out/r1_bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r1_launches_c2_final.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_visibility" -s 12 -c 6 -o gpurun_out/r1_prof_c2_final -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
tail -2 gpurun_out/r1_ncu_full.log
