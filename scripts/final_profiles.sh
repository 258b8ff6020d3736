#!/bin/bash
# round-end evidence: GPU test log, bench JSON lines of every workload, ncu launch lists of the bench command, --set full captures
set -x
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu) > gpurun_out/r1_gpu_tests.log 2>&1
python bench.py > gpurun_out/r1_bench_c2.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2>> gpurun_out/r1_bench.err
python bench.py --workload c1_teapot --steps 1000 > gpurun_out/r1_bench_c1_teapot.json 2>> gpurun_out/r1_bench.err
python bench.py --workload c3_dragon --steps 500 > gpurun_out/r1_bench_c3_dragon.json 2>> gpurun_out/r1_bench.err
python bench.py --workload c4_tree_sv --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_c4_tree_sv.json 2>> gpurun_out/r1_bench.err
python bench.py --workload c4_tree_sv_1080p --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_c4_tree_sv_1080p.json 2>> gpurun_out/r1_bench.err
python bench.py --workload c5_many_light --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_c5_1gpu.json 2>> gpurun_out/r1_bench.err
for w in c2_sponza_pcf c2_sponza_pcss_k7 dragon_pcss; do python bench.py --workload $w --steps 500 > gpurun_out/r1_bench_$w.json 2>> gpurun_out/r1_bench.err; done
for t in vsm esm evsm msm; do python bench.py --workload c2_sponza_$t --steps 1000 > gpurun_out/r1_bench_c2_$t.json 2>> gpurun_out/r1_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r1_launches_c2_final.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r1_launches_c2_vsm.csv python bench.py --workload c2_sponza_vsm --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launch_vsm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_visibility" -s 12 -c 3 -o gpurun_out/r1_prof_c2_final -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_mom" -s 12 -c 4 -o gpurun_out/r1_prof_c2_vsm -f python bench.py --workload c2_sponza_vsm --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full_vsm.log 2>&1
tail -3 gpurun_out/r1_gpu_tests.log; tail -2 gpurun_out/r1_ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
