#!/bin/bash
# quick c2 / dragon_pcss / pcf numbers
for w in c2_sponza dragon_pcss c2_sponza_pcf; do
python bench.py --workload $w --steps 300 --warmup 10 --no-cpu-baseline --no-sharded --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', 'fps %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], {k: round(v,4) for k,v in d['pass_ms'].items()})"
done
