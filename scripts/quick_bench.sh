#!/bin/bash
# developer helper: parity tests + per-pass timings for a few CTA sizes of the tile kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for nt in 256 512 1024; do
  echo "== SGI_TILE_THREADS=$nt"
  SGI_TILE_THREADS=$nt python bench.py --steps 100 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: round(v,3) for k,v in d['pass_ms'].items()})"
  SGI_TILE_THREADS=$nt python scripts/perf_probe.py teapot 1280 720 1024 hard 20
  SGI_TILE_THREADS=$nt python scripts/perf_probe.py dragon 3840 2160 4096 rbsm_noncons 20
  SGI_TILE_THREADS=$nt python scripts/perf_probe.py teapot 1920 1080 2048 pcss 20
done
