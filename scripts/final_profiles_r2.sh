#!/bin/bash
# round-2 evidence (one GPU): GPU test log, bench JSON lines of every workload, ncu launch lists, --set full captures
set -x
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu) > gpurun_out/r2_gpu_tests.log 2>&1
python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
for w in c5_sandiego c5_many_light; do python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_bench_$w.json 2>> gpurun_out/r2_bench.err; done
for w in c1_teapot c3_dragon dragon_pcss c2_sponza_pcf c2_sponza_vsm; do python bench.py --workload $w --steps 300 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_bench_$w.json 2>> gpurun_out/r2_bench.err; done
for w in c4_tree_sv c4_tree_sv_1080p c4_tree_sv_pertri c4_tree_sv_zfail; do python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_bench_$w.json 2>> gpurun_out/r2_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/r2_launches_c5_sharded.csv python bench.py --sharded-only > gpurun_out/r2_ncu_launch_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --workload c4_tree_sv --steps 4 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_launch_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_visibility|k_setup_bin|k_order" -s 16 -c 6 -o gpurun_out/r2_prof_c2 -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_visibility_multi_fused" -s 70 -c 3 -o gpurun_out/r2_prof_c5 -f python bench.py --sharded-only > gpurun_out/r2_ncu_full_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_setup_bin" -s 12 -c 4 -o gpurun_out/r2_prof_c4 -f python bench.py --workload c4_tree_sv --steps 2 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/r2_ncu_full_c4.log 2>&1
tail -3 gpurun_out/r2_gpu_tests.log; tail -3 gpurun_out/r2_bench.err; ls -la gpurun_out/*.ncu-rep
