#!/bin/bash
# developer helper: CTA size x sub-tile split threshold of the tile kernel over several workloads (fps / per-pass ms)
for nt in ${NTS:-256 512 1024}; do
 for sp in ${SPLITS:-0 64 256}; do
  echo "== SGI_TILE_THREADS=$nt SGI_TILE_SPLIT=$sp"
  export SGI_TILE_THREADS=$nt SGI_TILE_SPLIT=$sp
  python bench.py --steps ${STEPS:-150} --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 fps %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: round(v,3) for k,v in d['pass_ms'].items()})"
  python scripts/perf_probe.py teapot 1280 720 1024 hard 30
  python scripts/perf_probe.py dragon 3840 2160 4096 rbsm_noncons 20
  python scripts/perf_probe.py teapot 7680 4320 8192 hard 10
  python scripts/perf_probe.py sv tree 640 480
 done
done
