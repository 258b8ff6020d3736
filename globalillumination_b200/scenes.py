"""The workloads of BASELINE.json as data: each is written out in the reference's Configs/*.txt grammar
 and loaded through the C++ SceneLoader, so the bench exercises the same host path a
reference user would.  Values are those of the reference's config files (cited per entry); geometry that is
absent from the reference mount (.MISSING_LARGE_BLOBS) is replaced by the seeded procedural stand-ins of
host/procedural.cpp via `o procedural:<spec>` lines.  /root/reference is never read at run time.
"""
import os
import tempfile

WORKLOADS = {
    # Configs/Sponza.txt:1-12 — asset OBJ/Sponza/sponza.obj is missing from the mount -> procedural atrium
    "c2_sponza": dict(
        lines=["o procedural:sponza_like?seed=1", "s 2 2 2", "r 0 90 0", "t 0.0 0.0 -30.0", "c 0.5 0.0 0.0", "+",
               "ve 0.0 19.0 -52.0", "va 0.0 -2.0 -22.0", "le 0.0 43.0 -50.0", "la 0.0 -17.0 -17.0", "d 0"],
        W=1920, H=1080, S=2048, program="soft_shadow_mapping", technique="pcss",
        params=dict(blockerSearchSize=7, kernelSize=15, lightSourceRadius=8),
        scene="procedural sponza_like(seed=1), 69424 triangles (reference asset absent from the mount)"),
    # Configs/SanDiego.txt:1-26 — building.obj is not shipped to the GPU box, spheres are missing upstream ->
    # procedural city block of spheres over the plane; 16 lights = 4x4 UniformSampledLightSource, size 16
    "c5_many_light": dict(
        lines=["o procedural:sponza_like?seed=5", "s 1.2 1.2 1.2", "t -10.0 -8.0 0.0", "r 0.0 90.0 0.0", "c 1.0 1.0 0.5", "+",
               "o procedural:sphere?seed=2", "s 1.5 1.5 1.5", "t 1.5 -8.0 6.0", "+",
               "o procedural:sphere?seed=2", "s 1.5 1.5 1.5", "t -2.5 -8.0 6.0", "+",
               "o procedural:plane", "s 60.0 1.0 40.0", "t 0.0 -8.0 0.0", "+",
               "ve 0.0 41.0 -50.0", "va 0.0 16.0 -10.0", "le 10.0 130.0 100.0", "la 0.0 0.0 0.0", "d 0.000025"],
        W=7680, H=4320, S=8192, program="soft_shadow_mapping", technique="montecarlo",
        params=dict(numberOfSamples=16, lightSourceSize=16),
        scene="procedural stand-in for Configs/SanDiego.txt (building/sphere assets not available on the GPU box)"),
    # The other configurations of BASELINE.json (parity-test cases; measurable with `bench.py --workload ...`).  Their
    # assets exist in the reference but /root/reference is not on the GPU box: the geometry is the reference loader's own
    # output for the config file, committed as globalillumination_b200/data/scene_<name>.npz (made by tests/golden/make_golden.py).
    # Configs/Teapot.txt — c1: hard shadow mapping (`naive`)
    "c1_teapot": dict(golden="teapot", lines=[], W=1280, H=720, S=1024, program="shadow_mapping", technique="naive", params={},
                      scene="Configs/Teapot.txt through the reference's SceneLoader (golden scene_teapot.npz), 15706 triangles"),
    # Configs/Dragon.txt — c3: revectorization-based shadow mapping at 4K
    "c3_dragon": dict(golden="dragon", lines=[], W=3840, H=2160, S=4096, program="shadow_mapping", technique="smsr", params={},
                      scene="Configs/Dragon.txt through the reference's SceneLoader (golden scene_dragon.npz), 100004 triangles"),
    # Configs/TreeWithLeaves.txt (without the missing TreeSub1.obj) — c4: shadow volumes at the reference's window size
    # north_star (4): silhouette extrusion - the side quads of interior edges cancel in pairs and are not drawn (svSilhouette);
    # the `_pertri` variants draw the reference's per-triangle prisms (ShadowVolume::update as is: the parity mode)
    "c4_tree_sv": dict(golden="tree", lines=[], W=640, H=480, S=64, program="shadow_volumes", technique="naive", params=dict(svSilhouette=1),
                       scene="Configs/TreeWithLeaves.txt minus the missing TreeSub1.obj (scene_tree.npz), 38200 triangles; silhouette quads: 74644 of 229200 prism triangles"),
    "c4_tree_sv_1080p": dict(golden="tree", lines=[], W=1920, H=1080, S=64, program="shadow_volumes", technique="naive", params=dict(svSilhouette=1),
                             scene="as c4_tree_sv at 1920x1080"),
    "c4_tree_sv_pertri": dict(golden="tree", lines=[], W=640, H=480, S=64, program="shadow_volumes", technique="naive", params={},
                              scene="Configs/TreeWithLeaves.txt minus the missing TreeSub1.obj (scene_tree.npz), 38200 triangles -> 229200 prism triangles (per-triangle prisms)"),
    "c4_tree_sv_zfail": dict(golden="tree", lines=[], W=640, H=480, S=64, program="shadow_volumes", technique="naive", params=dict(svSilhouette=1, svZfail=1),
                             scene="as c4_tree_sv, depth-fail counting over capped volumes"),
}


# Config c5 on the reference's own geometry: Configs/SanDiego.txt through the reference's SceneLoader (building.obj + plane; the two
# spheres are missing upstream, .MISSING_LARGE_BLOBS), 16 lights = 4x4 UniformSampledLightSource of size 16, 8192^2 maps, 7680x4320.
# This is the workload of bench.py's `sharded` record (light shards over 1/2/4/8 GPUs).
WORKLOADS["c5_sandiego"] = dict(golden="sandiego", lines=[], W=7680, H=4320, S=8192, program="soft_shadow_mapping", technique="montecarlo",
                                params=dict(numberOfSamples=16, lightSourceSize=16),
                                scene="Configs/SanDiego.txt through the reference's SceneLoader (scene_sandiego.npz), 39500 triangles; 16 lights")


# Moment shadow maps on the headline scene: the ShadowMapping program with VSM / ESM / EVSM / MSM set
# (ShadowMapping/src/main.cpp:459-472), Gaussian order 7 (:859).  Measurable with `bench.py --workload c2_sponza_vsm` etc.
for _t in ("vsm", "esm", "evsm", "msm"):
    WORKLOADS["c2_sponza_" + _t] = dict(WORKLOADS["c2_sponza"], program="shadow_mapping", technique=_t, params={},
                                        scene=WORKLOADS["c2_sponza"]["scene"] + f"; moment shadow map ({_t}), blur order 7")


# The other two parameter sets SURVEY 8(d) lists for config c2: (i) PCF order 7, penumbra 1 (ShadowMapping program, bilinearPCF) and
# (ii) PCSS with kernelSize 7 (the state after the reference's "reset"); the headline c2_sponza is (iii), kernelSize 15.
WORKLOADS["c2_sponza_pcf"] = dict(WORKLOADS["c2_sponza"], program="shadow_mapping", technique="pcf", params=dict(kernelOrder=7, penumbraSize=1),
                                  scene=WORKLOADS["c2_sponza"]["scene"] + "; PCF 7x7")
WORKLOADS["c2_sponza_pcss_k7"] = dict(WORKLOADS["c2_sponza"], params=dict(blockerSearchSize=7, kernelSize=7, lightSourceRadius=8),
                                      scene=WORKLOADS["c2_sponza"]["scene"] + "; PCSS kernelSize 7")
# PCSS where its filter loop really runs: on the Sponza-like light every blocker average is below the shader's 0.99 cut-off
# (PlausibleSoftShadow.frag:368) and the pass ends after the blocker search; under the Dragon / Teapot light
# (10,130,100) the light-space depths are 0.991-0.996 and penumbra pixels take all kernelSize^2 filter taps as well.
WORKLOADS["dragon_pcss"] = dict(golden="dragon", lines=[], W=1920, H=1080, S=2048, program="soft_shadow_mapping", technique="pcss",
                                params=dict(blockerSearchSize=7, kernelSize=15, lightSourceRadius=8),
                                scene="Configs/Dragon.txt through the reference's SceneLoader (golden scene_dragon.npz), 100004 triangles; PCSS at the c2 sizes")


def golden_scene(name):
    """The reference loader's output for a config (input fixture shipped with the package: globalillumination_b200/data/,
    written by tests/golden/make_golden.py from the reference's own SceneLoader)."""
    import numpy as np
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"scene_{name}.npz")))


def write_config(name, directory=None):
    """Write WORKLOADS[name] as a Configs-style text file and return its path."""
    w = WORKLOADS[name]
    directory = directory or tempfile.mkdtemp(prefix="shadowgi_cfg_")
    path = os.path.join(directory, name + ".txt")
    with open(path, "w") as f:
        f.write("\n".join(w["lines"]))          # no trailing newline, like the reference's files
    return path
