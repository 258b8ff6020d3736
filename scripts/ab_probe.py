"""Developer A/B helper: frame rate of bench workloads under different SGI_* environment switches (one bench.py run each)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = [{}, {"SGI_SV_SPLIT_LISTS": "0"}, {"SGI_TILE_BIN_BIG": "0"}]
if __name__ == "__main__":
    for w, steps in (("c4_tree_sv", 40), ("c4_tree_sv_1080p", 40), ("c4_tree_sv_pertri", 20), ("c3_dragon", 200), ("c5_sandiego", 30)):
        for cfg in CONFIGS:
            env = dict(os.environ); env.update(cfg)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline",
                                "--no-sharded", "--no-secondary"], capture_output=True, text=True, env=env)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                print(w, cfg, "fps %.1f" % d["value"], {k: round(v, 4) for k, v in d["pass_ms"].items() if k.startswith("tile")}, flush=True)
            except Exception as e:
                print(w, cfg, "FAILED", r.stderr[-300:], flush=True)
