"""Developer A/B helper: per-pass timings of bench workloads under different SGI_* environment switches (one context each);
passes run one after the other (overlap off), so tile_depth / tile_gbuffer are the kernels' own durations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.perf_probe import probe_app

CONFIGS = [{}, {"SGI_TILE_BULK": "0"}, {"SGI_TILE_THREADS": "128"}, {"SGI_TILE_THREADS": "256"}, {"SGI_TILE_THREADS": "512"}]
if __name__ == "__main__":
    workloads = sys.argv[1:] or ["c2_sponza", "c5_many_light"]
    for w in workloads:
        for cfg in CONFIGS:
            for k in ("SGI_TILE_BULK", "SGI_TILE_THREADS"):
                os.environ.pop(k, None)
            os.environ.update(cfg)
            print(cfg, end=" ", flush=True)
            probe_app(w, 20, overlap=False)
