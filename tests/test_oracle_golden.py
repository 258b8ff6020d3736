"""The CPU oracle against the golden vectors made by the reference's own sources (tests/golden/make_golden.py):
every technique of the reference's GLSL fragment shaders, GLM matrices, ShadowVolume prisms, light samples.
Runs without /root/reference; when oracle/_ref is present the same comparisons are repeated live on fresh inputs."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util

TECHS = ["hard", "pcf", "pcss", "rbsm_noncons", "rbsm_cons", "rpcf_noncons", "rpcf_cons", "rsmss", "rbssm"]
ALT = dict(kernel_order=9, kernel_size=7, shadow_intensity=0.5, max_search=8)


def _frame(g):
    fm = {k[3:]: g[k] for k in g if k.startswith("fm_")}
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    return fm, cam


@pytest.mark.parametrize("variant", ["default", "alt"])
@pytest.mark.parametrize("tech", TECHS)
def test_visibility_matches_reference_shader_golden(tech, variant):
    g = util.golden("golden_shaders.npz")
    fm, cam = _frame(g)
    S = int(g["S"])
    p = O.default_params(tech, S, depth_threshold=float(g["depth_threshold"]), **(ALT if variant == "alt" else {}))
    vis = O.visibility(p, cam, fm["light_mvp_b"], g["pos"], g["nrm"], g["sm"])
    ref = g[f"vis/{tech}/{variant}"]
    assert util.bits_equal(vis, ref), util.describe_diff(vis, ref)
    fg = g["pos"][..., 0] != 0
    assert (ref[~fg] == 0).all() and 0.02 < (ref[fg] < 1).mean() < 0.98


@pytest.mark.parametrize("tech", ["edtsm_noncons", "edtsm_cons"])
def test_edt_shadow_mapping_stages_match_reference_shaders_golden(tech):
    """EDTSM: the hard-shadow target (RBSM shader with EDTSM == 1) and both MeanFilter.frag passes bit-exact against the
    reference's shaders; the filter inputs stored with them are what the oracle's own site / Voronoi / normalise stages
    produce from that target."""
    g = util.golden("golden_shaders.npz")
    fm, cam = _frame(g)
    S = int(g["S"])
    p = O.default_params(tech, S, depth_threshold=float(g["depth_threshold"]), penumbra_size=5)
    img = O.edt_hard_image(p, cam, fm["cam_mvp"], fm["light_mvp_b"], g["pos"], g["nrm"], g["sm"])
    assert util.bits_equal(img, g[f"edt/{tech}/hard"]), util.describe_diff(img, g[f"edt/{tech}/hard"])
    near = O.edt_nearest(O.edt_sites(img))
    a2 = O.edt_normalize(img, g["pos"], near, np.float32(p.penumbra_size / 5.0), p.shadow_intensity)
    assert util.bits_equal(a2, g[f"edt/{tech}/filter_x_in"])
    fg = g["pos"][..., 0] != 0
    bx = O.mean_filter(a2, g["pos"], fm["cam_mv"], p.kernel_order, True)
    assert util.bits_equal(bx[fg], g[f"edt/{tech}/filter_x"][fg])
    by = O.mean_filter(bx, g["pos"], fm["cam_mv"], p.kernel_order, False, linear=True)
    assert util.bits_equal(by[fg], g[f"edt/{tech}/filter_y"][fg])
    vis, near2 = O.edtsm(p, cam, fm["cam_mvp"], fm["light_mvp_b"], g["pos"], g["nrm"], g["sm"])
    assert util.bits_equal(vis[fg], by[..., 0][fg]) and np.array_equal(near2, near)
    ramp = (vis > p.shadow_intensity) & (vis < 1.0)
    assert ramp.sum() > 200                                   # a penumbra exists


@pytest.mark.parametrize("W,H,n_sites,seed", [(37, 23, 5, 0), (64, 64, 40, 1), (101, 33, 1, 2), (50, 70, 600, 3), (16, 16, 0, 4)])
def test_edt_nearest_site_is_the_brute_force_answer(W, H, n_sites, seed):
    """Exact Euclidean nearest site, ties to the smallest (y, x); no sites -> MARKER everywhere."""
    rng = np.random.default_rng(seed)
    sites = np.full((H, W, 2), O.EDT_MARKER, np.int16)
    flat = rng.choice(W * H, size=n_sites, replace=False) if n_sites else np.zeros(0, np.int64)
    ys, xs = flat // W, flat % W
    sites[ys, xs, 0] = xs; sites[ys, xs, 1] = ys
    near = O.edt_nearest(sites)
    if n_sites == 0:
        assert (near == O.EDT_MARKER).all()
        return
    yy, xx = np.mgrid[0:H, 0:W]
    d = (xx[..., None] - xs) ** 2 + (yy[..., None] - ys) ** 2                  # [H, W, n]
    key = d.astype(np.int64) * (1 << 32) + ys.astype(np.int64) * (1 << 16) + xs.astype(np.int64)
    k = key.argmin(-1)
    assert np.array_equal(near[..., 0], xs[k]) and np.array_equal(near[..., 1], ys[k])


def test_many_light_matches_reference_shader_golden():
    g = util.golden("golden_shaders.npz")
    S = int(g["S"])
    p = O.default_params("multi_hard", S)
    mvpb = g["multi/mvpb"]
    vis = O.visibility_multi(p, mvpb[-1], mvpb[:, 12:16], g["pos"], g["multi/maps"])
    assert util.bits_equal(vis, g["multi/vis"]), util.describe_diff(vis, g["multi/vis"])


def test_golden_inputs_are_what_the_oracle_rasterises():
    """The stored frame (depth map, G-buffer) is the oracle rasteriser's output for the teapot scene: guards the
    rasteriser definition against silent changes."""
    g = util.golden("golden_shaders.npz")
    sc = util.scene("teapot")
    fm, _ = _frame(g)
    W, H, S = int(g["W"]), int(g["H"]), int(g["S"])
    sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    assert util.bits_equal(sm, g["sm"]) and util.bits_equal(pos, g["pos"]) and util.bits_equal(nrm, g["nrm"])


@pytest.mark.parametrize("name", ["teapot", "door", "dragon", "raptor"])
def test_matrices_match_glm_golden(name):
    g = util.golden("golden_host.npz")
    sc = util.scene(name)
    for (W, H, S) in ((1280, 720, 1024), (1920, 1080, 2048), (640, 480, 512)):
        fm = util.frame(sc, W, H, S)
        for k, v in fm.items():
            assert util.bits_equal(v, g[f"fm/{name}/{W}x{H}x{S}/{k}"]), (name, W, H, S, k)


def test_rotate_matches_glm_golden():
    g = util.golden("golden_host.npz")
    assert util.bits_equal(O.rotate(33.5, [0, 1, 0]), g["rotate/33.5/y"])


def test_shadow_volume_prisms_match_reference_golden():
    g = util.golden("golden_host.npz")
    d = util.scene("door")
    pxyz, pidx = O.sv_build_prisms(d["xyz"], d["nrm"], d["idx"], d["light_eye"], 100)
    assert util.bits_equal(pxyz, g["sv/door/xyz"]) and np.array_equal(pidx, g["sv/door/idx"])
    assert len({tuple(r) for r in pidx.reshape(-1, 18) - 6 * np.arange(len(d["idx"]))[:, None]}) == 2   # both windings occur


@pytest.mark.parametrize("n,size", [(16, 16), (289, 16), (4, 8)])
def test_uniform_light_samples_match_reference_golden(n, size):
    g = util.golden("golden_host.npz")
    got = np.stack([O.uniform_light_sample(np.array([10, 130, 100], np.float32), size, n, i) for i in range(n)])
    assert util.bits_equal(got, g[f"uls/{n}/{size}"])


def test_pcf_float_loop_tap_counts():
    """the float loops give 7 / 8 taps per axis at order 7, 10 at order 9, 16 at 15, 11 at 11 (`<`)."""
    assert len(O.pcf_offsets(7, 1, False)) == 7 and len(O.pcf_offsets(7, 1, True)) == 8
    assert len(O.pcf_offsets(9, 1, False)) == 10 and len(O.pcf_offsets(15, 1, False)) == 16 and len(O.pcf_offsets(11, 1, False)) == 11
    off = O.pcf_offsets(7, 1, False)
    assert off[0] == -1.0 and not np.allclose(off, -off[::-1])        # asymmetric, as in the reference


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("tech", TECHS)
def test_live_reference_shader_on_fresh_inputs(tech):
    """Same comparison on a different scene / size / parameters, executing the reference's shader source now."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(util.GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    sc = util.scene("raptor")
    W, H, S = 200, 112, 160
    fm = util.frame(sc, W, H, S)
    sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    p = O.default_params(tech, S, depth_threshold=2.5e-5, kernel_order=5, kernel_size=9, blocker_search_size=5, max_search=12)
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis = O.visibility(p, cam, fm["light_mvp_b"], pos, nrm, sm)
    ref = mg.ref_visibility(tech, fm, pos, nrm, sm, S, p, W, H)
    assert util.bits_equal(vis, ref), util.describe_diff(vis, ref)


def test_phong_shading_matches_reference_shader_golden():
    """Deferred shading (PhongShading.frag, next-row §8f): oracle == the reference's shader on the golden frame."""
    g = util.golden("golden_shaders.npz")
    fm, cam = _frame(g)
    img = O.shade_phong(cam, 0.25, g["pos"], g["nrm"], g["phong/albedo"], g["vis/pcf/default"])
    fg = g["pos"][..., 0] != 0
    assert util.bits_equal(img[fg], g["phong/image"][fg])
    assert np.allclose(img[~fg], O.CLEAR_COLOR)                       # discarded pixels keep glClearColor (main.cpp:453)
    sc = util.scene("teapot")
    _, _, alb, _ = O.raster_gbuffer_rgb(sc["xyz"], sc["nrm"], g["phong/rgb"], sc["idx"], fm["cam_mvp"], int(g["W"]), int(g["H"]))
    assert util.bits_equal(alb, g["phong/albedo"])
