#!/bin/bash
# c4 A/B with the big-record kernel forced on + launch list
tag=${1:-trip}
mkdir -p gpurun_out
for w in c4_tree_sv c4_tree_sv_1080p c2_sponza c1_teapot; do
for cfg in "" "SGI_TILE_BIN_BIG=1 SGI_TILE_BIN_BIG_WORK=0"; do
env $cfg python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --no-sharded --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', '[$cfg]', 'fps %.1f' % d['value'], {k: round(v,4) for k,v in d['pass_ms'].items()})"
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 150 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --workload c4_tree_sv --steps 6 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary > gpurun_out/${tag}_ncu_launch.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_launches_c4.csv | cut -c1-180
