"""Copy the evidence scripts/final_profiles.sh left in gpurun_out/ into profiles/ (tracked) with the read-here summaries:
bench JSON lines, launch lists + per-kernel shares, key metrics and DRAM bytes of the --set full captures.
Run on the CPU box after the gpurun call:  python scripts/collect_profiles.py"""
import glob, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_summary  # noqa: E402


def main():
    rows = []
    for f in sorted(glob.glob(os.path.join(G, "r1_bench_*.json"))):
        lines = [l for l in open(f).read().strip().splitlines() if l.startswith("{")]
        if not lines:
            continue
        d = json.loads(lines[-1])
        with open(os.path.join(P, os.path.basename(f)), "w") as o:
            o.write(json.dumps(d) + "\n")
        r, sp, cb = d.get("roofline") or {}, d.get("shadow_pass") or {}, d.get("cpu_baseline") or {}
        rows.append(f"{os.path.basename(f)[9:-5]:22s} n_gpus {d.get('n_gpus', 1)} value {d['value']:9.1f} e2e {d.get('e2e', {}).get('value', 0):9.1f} "
                    f"cpu {cb.get('value', 0):6.2f} | top kernel {r.get('kernel', '-')[:34]:34s} frac {r.get('frac', 0):.3f} | shadow pass hbm "
                    f"{sp.get('hbm_frac', 0):.3f} l2 taps {sp.get('l2_taps_frac', '-')}")
    with open(os.path.join(P, "r1_bench_summary.txt"), "w") as o:
        o.write("# one line per profiles/r1_bench_*.json (frames/s; roofline.frac of the kernel with the largest share; shadow pass vs HBM / L2 peaks)\n")
        o.write("\n".join(rows) + "\n")
    print("\n".join(rows))
    for f in glob.glob(os.path.join(G, "r1_launches_*.csv")):
        dst = os.path.join(P, os.path.basename(f))
        shutil.copy(f, dst)
        with open(dst[:-4] + "_summary.txt", "w") as o:
            o.write(ncu_summary.summarise(os.path.relpath(dst, ROOT)) + "\n")
    heads = {
        "r1_prof_c2_final": ("r1_ncu_full_c2_final_keymetrics.txt", "dram_traffic_c2.json",
                             "# round 1 (final build of the round) — ncu --set full --clock-control none --import-source on, bench.py c2 (Sponza-like 1920x1080, S=2048, PCSS)\n"
                             "# kernels: k_tile<3,512> = camera G-buffer tile rasteriser + resolve (with the albedo target), k_tile<0,512> = light-view depth tile rasteriser, k_visibility<2,7,15> = PCSS\n"
                             "# note: under ncu every kernel is serialised and starts with cold caches; compare shares, not absolutes (the .ncu-rep is not committed); extracted with scripts/ncu_keymetrics.py\n"),
        "r1_prof_c2_vsm": ("r1_ncu_full_c2_vsm_keymetrics.txt", "dram_traffic_c2_vsm.json",
                           "# round 1 (final build) — ncu --set full, bench.py --workload c2_sponza_vsm: k_tile<4,512> = light-view moment pass (raster + moment resolve), k_mom_filter<1,0,7> / <0,0,7> = blur X / Y (order 7), k_tile<3,512> as above\n"),
    }
    for rep, (txt, js, head) in heads.items():
        path = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_keymetrics.py"), os.path.relpath(path, ROOT), os.path.join(P, js)],
                             capture_output=True, text=True, cwd=ROOT).stdout
        with open(os.path.join(P, txt), "w") as o:
            o.write(head + out)
    log = os.path.join(G, "r1_gpu_tests.log")
    if os.path.exists(log):
        shutil.copy(log, os.path.join(P, "r1_gpu_tests.log"))


if __name__ == "__main__":
    main()
