"""Copies the outputs of scripts/final_profiles_r2.sh from gpurun_out/ into profiles/ and derives the text summaries that are committed
(the .ncu-rep files themselves are not): launch-list summaries, key metrics + DRAM traffic of the --set full captures, source hot spots."""
import glob, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_summary  # noqa: E402


def run(*args):
    return subprocess.run([sys.executable, *args], capture_output=True, text=True, cwd=ROOT).stdout


def main():
    for f in sorted(glob.glob(os.path.join(G, "r2_bench_*.json"))):
        if os.path.getsize(f) > 0:
            shutil.copy(f, os.path.join(P, os.path.basename(f)))
    for f in glob.glob(os.path.join(G, "r2_launches_*.csv")):
        dst = os.path.join(P, os.path.basename(f))
        shutil.copy(f, dst)
        with open(dst[:-4] + "_summary.txt", "w") as o:
            o.write(ncu_summary.summarise(os.path.relpath(dst, ROOT)) + "\n")
    note = ("# note: under ncu every kernel is serialised and starts with cold caches; compare shares, not absolutes (the .ncu-rep is not committed); "
            "extracted with scripts/ncu_keymetrics.py\n")
    heads = {
        "r2_prof_c2": ("r2_ncu_full_c2_keymetrics.txt", "dram_traffic_c2.json",
                       "# round 2 (final build) - ncu --set full --clock-control none --import-source on, bench.py c2 (Sponza-like 1920x1080, S=2048, PCSS)\n"
                       "# k_tile<3,512> = camera G-buffer tile rasteriser + resolve, k_tile<0,512> = light-view depth tile rasteriser, k_visibility<2,7,15> = PCSS, k_setup_bin / k_order = binning chain\n"),
        "r2_prof_c5": ("r2_ncu_full_c5_sharded_keymetrics.txt", None,
                       "# round 2 (final build) - ncu --set full, bench.py --sharded-only (c5 on the SanDiego scene, N = 1): k_tile<0,256> = 8192^2 depth pass of one light, "
                       "k_tile<5,256> = primitive-id camera pass, k_visibility_multi_fused = 16-light accumulation\n"),
        "r2_prof_c4": ("r2_ncu_full_c4_keymetrics.txt", None,
                       "# round 2 (final build) - ncu --set full, bench.py --workload c4_tree_sv: k_tile<2,1024> = stencil counting of the silhouette prisms (hot tiles shared by list segment), "
                       "k_setup_bin = set-up + binning of the 229 200 prism slots (tile-major walk)\n"),
    }
    hot = ["# ncu --set full --import-source on, scripts/ncu_hotspots.py: share of warp-stall samples / executed warp instructions per CUDA source line (final build of round 2)\n"]
    for rep, (txt, js, head) in heads.items():
        path = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(path):
            print("missing", path)
            continue
        args = [os.path.join(ROOT, "scripts", "ncu_keymetrics.py"), os.path.relpath(path, ROOT)]
        if js:
            args.append(os.path.join(P, js))
        with open(os.path.join(P, txt), "w") as o:
            o.write(head + note + run(*args))
        hot.append("## " + rep + "\n" + run(os.path.join(ROOT, "scripts", "ncu_hotspots.py"), os.path.relpath(path, ROOT), "", "16"))
    with open(os.path.join(P, "r2_source_hotspots.txt"), "w") as o:
        o.write("\n".join(hot))
    log = os.path.join(G, "r2_gpu_tests.log")
    if os.path.exists(log):
        shutil.copy(log, os.path.join(P, "r2_gpu_tests.log"))
    print("done")


if __name__ == "__main__":
    main()
