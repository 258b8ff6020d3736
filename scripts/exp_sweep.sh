#!/bin/bash
# developer helper: the standard set of workloads (fps / per-pass ms) with the current defaults (env overrides apply)
python bench.py --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 fps %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: round(v,3) for k,v in d['pass_ms'].items()})"
python scripts/perf_probe.py teapot 1280 720 1024 hard 30
python scripts/perf_probe.py dragon 3840 2160 4096 rbsm_noncons 20
python scripts/perf_probe.py teapot 7680 4320 8192 hard 10
python scripts/perf_probe.py sv tree 640 480
python scripts/perf_probe.py sv dragon 1920 1080
python scripts/perf_probe.py app c5_many_light 5
python scripts/perf_probe.py teapot 1920 1080 2048 rbssm 5
python scripts/perf_probe.py dragon 1920 1080 2048 rbssm 5
