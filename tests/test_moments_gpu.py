"""Moment shadow maps (VSM / ESM / EVSM / MSM) through the C ABI against the CPU oracle.

Everything without exp / log is required bit-identical (moment target, Gaussian blur, VSM and MSM reconstruction); the
exponential paths (ESM's log-space blur, ESM / EVSM reconstruction) go through expf / logf, whose CUDA and glibc versions
differ by a few ulp, so they carry an explicit tolerance far inside north_star's 1e-3 for soft visibility."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util

SOFT_TOL = 1e-3        # north_star: soft visibility within 1e-3 absolute
EXP_TOL = 2e-5         # what expf / logf implementations may differ by after the c = 80 exponent (measured: < 1e-5)


def test_host_quantization_equals_oracle_and_reference_glm():
    """sgi_moment_quantization is host arithmetic (no device): same bits as the oracle and as the reference's GLM golden."""
    from globalillumination_b200 import capi
    g = util.golden("golden_moments.npz")
    m, mi, t = capi.moment_quantization()
    assert util.bits_equal(m, g["quant/m"]) and util.bits_equal(mi, g["quant/minv"]) and util.bits_equal(t, g["quant/t"])


@pytest.fixture(scope="module")
def ctx():
    from globalillumination_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _setup(ctx, sc, W, H, S, pg):
    fm = util.frame(sc, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    return fm


def _close(a, b, tol):
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(((np.abs(a - b) <= tol) | both_nan).all())


@pytest.mark.gpu
@pytest.mark.parametrize("tech", O.MOMENT_TECHS)
@pytest.mark.parametrize("name,W,H,S,kw", [("teapot", 320, 180, 256, {}), ("teapot", 1280, 720, 1024, {}),
                                           ("raptor", 333, 217, 300, dict(kernel_order=5, shadow_intensity=0.5)),
                                           ("dragon", 640, 360, 512, dict(kernel_order=11))])
def test_moment_chain_matches_oracle(ctx, tech, name, W, H, S, kw):
    sc = util.scene(name)
    po, pg = util.params_pair(tech, S, **kw)
    fm = _setup(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    mom, fx, fy, vis = ctx.read("moments"), ctx.read("moments_x"), ctx.read("moments_filtered"), ctx.read("visibility")
    pos, nrm = ctx.read("gbuf_pos"), ctx.read("gbuf_nrm")
    # 1. light-view moment target: bit-exact (no transcendental functions on this stage)
    mom_o = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, tech)
    assert util.bits_equal(mom, mom_o), "moment target: " + util.describe_diff(mom, mom_o)
    assert (mom_o[..., 0] != 0).mean() > 0.05
    # 2. filterShadowMap: Gaussian bit-exact; ESM's log-space blur within EXP_TOL
    order = po.kernel_order
    fx_o = O.filter_moments(mom, W, H, order, True, tech == "esm")
    fy_o = O.filter_moments(fx, W, H, order, False, tech == "esm")          # from the GPU's own X pass: stage-wise comparison
    if tech == "esm":
        assert _close(fx, fx_o, EXP_TOL) and _close(fy, fy_o, EXP_TOL), (np.nanmax(np.abs(fx - fx_o)), np.nanmax(np.abs(fy - fy_o)))
    else:
        assert util.bits_equal(fx, fx_o), "X pass: " + util.describe_diff(fx, fx_o)
        assert util.bits_equal(fy, fy_o), "Y pass: " + util.describe_diff(fy, fy_o)
    # 3. Shadow.frag reconstruction on the GPU's own filtered map
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility_moments(po, cam, fm["light_mvp_b"], pos, nrm, fy)
    assert np.array_equal(np.isnan(vis), np.isnan(vis_o))
    ok = ~np.isnan(vis_o)
    assert (~ok).mean() < 0.01                                               # MSM: sqrt of a negative discriminant, as in the shader
    assert np.abs(vis[ok] - vis_o[ok]).max() <= SOFT_TOL
    if tech in ("vsm", "msm"):
        assert util.bits_equal(vis[ok], vis_o[ok]), util.describe_diff(vis[ok], vis_o[ok])
    else:
        assert np.abs(vis[ok] - vis_o[ok]).max() <= EXP_TOL, float(np.abs(vis[ok] - vis_o[ok]).max())
    fg = pos[..., 0] != 0
    assert (vis_o[~fg] == 0).all() and 0.02 < (vis_o[fg & ok] < 0.999).mean() < 0.98
    # 4. end to end against the all-oracle chain: inside north_star's soft-shadow tolerance
    vis_full = O.visibility_moments(po, cam, fm["light_mvp_b"], pos, nrm, O.filter_shadow_map(mom_o, W, H, order, tech))
    okf = ok & ~np.isnan(vis_full)
    assert np.abs(vis[okf] - vis_full[okf]).max() <= SOFT_TOL


@pytest.mark.gpu
def test_moment_passes_need_their_inputs_and_switch_back_to_depth_maps(ctx):
    from globalillumination_b200 import capi
    sc = util.scene("teapot")
    W, H, S = 160, 90, 128
    po, pg = util.params_pair("vsm", S)
    fm = _setup(ctx, sc, W, H, S, pg)
    ctx.render_gbuffer()
    with pytest.raises(capi.SgiError):
        ctx.compute_visibility()                    # no moment map yet
    ctx.render_shadow_map()
    with pytest.raises(capi.SgiError):
        ctx.compute_visibility()                    # not filtered yet
    ctx.filter_shadow_map(); ctx.compute_visibility()
    v_vsm = ctx.read("visibility")
    # another moment technique needs its own map
    po2, pg2 = util.params_pair("esm", S)
    ctx.set_params(pg2)
    with pytest.raises(capi.SgiError):
        ctx.filter_shadow_map()
    with pytest.raises(capi.SgiError):
        ctx.compute_visibility()
    # blur order outside the shader's kernel[] array is refused
    for bad in (35, 8, 1):                          # past `kernel[33]`, even, too small
        with pytest.raises(capi.SgiError):
            ctx.set_params(capi.default_params("vsm", kernel_order=bad))
    # back to a depth-map technique on the same context: unchanged results
    po3, pg3 = util.params_pair("pcf", S)
    ctx.set_params(pg3)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    v_o = O.visibility(po3, cam, fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    assert util.bits_equal(ctx.read("visibility"), v_o)
    assert not np.array_equal(v_vsm, v_o)


@pytest.mark.gpu
def test_moment_visibility_respects_the_screen_rectangle(ctx):
    """Multi-GPU screen tiles: a rank evaluates only its rectangle; the union of two strips equals the whole frame."""
    sc = util.scene("teapot")
    W, H, S = 320, 180, 256
    po, pg = util.params_pair("evsm", S)
    _setup(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    whole = ctx.read("visibility").copy()
    parts = np.zeros_like(whole)
    for y0, y1 in ((0, 77), (77, H)):
        po_r, pg_r = util.params_pair("evsm", S, rect_x0=0, rect_y0=y0, rect_x1=W, rect_y1=y1)
        ctx.set_params(pg_r)
        ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        parts[y0:y1] = ctx.read("visibility")[y0:y1]
    assert util.bits_equal(parts, whole)


@pytest.mark.gpu
@pytest.mark.parametrize("tech", ["vsm", "msm"])
def test_host_display_runs_the_reference_frame_order(tech):
    """ShadowApp::display with shadowParams.VSM / MSM set: renderShadowMap, filterShadowMap, renderGBuffer,
    computeHardShadows (ShadowMapping/src/main.cpp:459-472) through the C++ host; the result equals the oracle chain."""
    from globalillumination_b200 import hostapi, scenes
    w = scenes.WORKLOADS["c1_teapot"]
    W, H, S = w["W"] // 2, w["H"] // 2, w["S"] // 2
    sc = scenes.golden_scene(w["golden"])              # Configs/Teapot.txt through the reference's loader
    app = hostapi.App(0)
    try:
        app.set_scene(sc); app.configure(W, H, S); app.set_technique(tech)
        app.display("shadow_mapping")
        c = app.context()
        vis, mom, fy = c.read("visibility"), c.read("moments"), c.read("moments_filtered")
        fm = util.frame(sc, W, H, S)
        assert (mom[..., 0] != 0).mean() > 0.2
        mom_o = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, tech)
        assert util.bits_equal(mom, mom_o), util.describe_diff(mom, mom_o)
        fy_o = O.filter_shadow_map(mom_o, W, H, 7, tech)
        assert util.bits_equal(fy, fy_o), util.describe_diff(fy, fy_o)
        pos_o, nrm_o, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        vis_o = O.visibility_moments(O.default_params(tech, S), cam, fm["light_mvp_b"], pos_o, nrm_o, fy_o)
        ok = ~np.isnan(vis_o)
        assert np.array_equal(np.isnan(vis), ~ok) and util.bits_equal(vis[ok], vis_o[ok]), util.describe_diff(vis[ok], vis_o[ok])
        # frames queued back to back without a host sync (the shadow pass of frame k overlaps the raster passes of frame k+1):
        # every frame still gets its own filtered map
        app.set(animationOn=1, animation=-1800.0)
        frames = []
        for k in range(4):
            app.display("shadow_mapping"); frames.append(c.read("visibility").copy()); app.step_animation(120.0)
        app.set(animation=-1800.0)
        for k in range(4):
            app.display("shadow_mapping"); app.step_animation(120.0)
        last = c.read("visibility")
        assert util.bits_equal(np.nan_to_num(last), np.nan_to_num(frames[3])) and not util.bits_equal(np.nan_to_num(frames[0]), np.nan_to_num(frames[3]))
    finally:
        app.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tech", ["vsm", "esm"])
def test_moment_chain_at_the_headline_size(tech):
    """BASELINE config c2's scene and sizes (Sponza-like, 1920x1080, 2048^2 map) with a pre-filtered technique, through the
    host's display(): moment target and Gaussian blur bit-exact, visibility within the stated tolerances; plus the size-
    independent properties of the chain (blur weights sum to one: a constant region stays constant; visibility in
    [shadowIntensity, 1] on the foreground, 0 on the background)."""
    from globalillumination_b200 import hostapi, scenes
    cfg = scenes.write_config("c2_sponza_" + tech)
    w = scenes.WORKLOADS["c2_sponza_" + tech]
    W, H, S = w["W"], w["H"], w["S"]
    app = hostapi.App(0)
    try:
        app.load_scene(cfg); app.configure(W, H, S); app.set_technique(tech)
        app.display("shadow_mapping")
        c = app.context()
        vis, mom, fx, fy = c.read("visibility"), c.read("moments"), c.read("moments_x"), c.read("moments_filtered")
        pos, nrm = c.read("gbuf_pos"), c.read("gbuf_nrm")
        sc = hostapi.load_scene(cfg)
        fm = util.frame(sc, W, H, S)
        mom_o = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, tech)
        assert util.bits_equal(mom, mom_o), util.describe_diff(mom, mom_o)
        fx_o = O.filter_moments(mom, W, H, 7, True, tech == "esm")
        fy_o = O.filter_moments(fx, W, H, 7, False, tech == "esm")
        if tech == "esm":
            assert _close(fx, fx_o, EXP_TOL) and _close(fy, fy_o, EXP_TOL)
        else:
            assert util.bits_equal(fx, fx_o) and util.bits_equal(fy, fy_o)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        vis_o = O.visibility_moments(O.default_params(tech, S), cam, fm["light_mvp_b"], pos, nrm, fy)
        assert np.abs(vis - vis_o).max() <= (EXP_TOL if tech == "esm" else 0.0)
        fg = pos[..., 0] != 0
        # (ESM clamps to [shadowIntensity, 1]; Chebyshev's variance / (variance + d^2) does not: where the blurred moments give a
        #  slightly negative variance the reference's formula leaves [0, 1], and so do the oracle and the kernel)
        assert (vis[~fg] == 0).all() and np.isfinite(vis).all() and 0.02 < (vis[fg] < 0.999).mean() < 0.98
        if tech == "esm":
            assert (vis[fg] >= 0.25).all() and (vis[fg] <= 1.0).all()
        else:
            assert ((vis[fg] < 0.25) | (vis[fg] > 1.0)).mean() < 0.01
        if tech == "vsm":
            # first moment of the blurred map stays inside the range of the un-blurred one (convex combination + border zeros)
            assert fy[..., 0].max() <= mom[..., 0].max() * (1 + 1e-6) and fy[..., 0].min() >= 0.0
    finally:
        app.close()


@pytest.mark.gpu
def test_host_selects_tricubic_pcf_like_the_filtering_menu():
    """shadowFilteringMenu case 1 (ShadowMapping/src/main.cpp:672-675): tricubicPCF alone -> every PCF tap is textureBicubic();
    with bilinearPCF set as well the bilinear assignment wins (Shadow.frag:101-104)."""
    from globalillumination_b200 import hostapi, scenes
    w = scenes.WORKLOADS["c1_teapot"]
    W, H, S = w["W"] // 2, w["H"] // 2, w["S"] // 2
    sc = scenes.golden_scene(w["golden"])
    app = hostapi.App(0)
    try:
        app.set_scene(sc); app.configure(W, H, S); app.set_technique("tricubic")
        app.display("shadow_mapping")
        c = app.context()
        vis, sm, pos, nrm = c.read("visibility"), c.read("shadow_map")[0], c.read("gbuf_pos"), c.read("gbuf_nrm")
        fm = util.frame(sc, W, H, S)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        vis_o = O.visibility(O.default_params("pcf_tricubic", S), cam, fm["light_mvp_b"], pos, nrm, sm)
        assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
        app.set_technique("pcf"); app.display("shadow_mapping")
        vis_pcf = c.read("visibility")
        assert util.bits_equal(vis_pcf, O.visibility(O.default_params("pcf", S), cam, fm["light_mvp_b"], pos, nrm, sm))
        assert not util.bits_equal(vis, vis_pcf)
    finally:
        app.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tech", ["vsm", "msm"])
def test_moment_chain_ragged_sizes_and_empty_scene(ctx, tech):
    """Non-square shadow map, window sizes that are no multiple of the CTA footprint, a window LARGER than the map (the blur then
    magnifies), degenerate / NaN triangles, and a scene without triangles: cleared moment target (0,0,0,1), its blur, and what
    Shadow.frag makes of it."""
    from globalillumination_b200 import capi
    sc = util.scene("teapot")
    W, H, SW, SH = 333, 217, 200, 120
    fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, SW, SH)
    po, pg = util.params_pair(tech, SW, shadow_map_height=SH, kernel_order=9)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], SW, SH)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    mom, fy, vis = ctx.read("moments"), ctx.read("moments_filtered"), ctx.read("visibility")
    assert mom.shape == (SH, SW, 4) and fy.shape == (H, W, 4)
    mom_o = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], SW, SH, tech)
    assert util.bits_equal(mom, mom_o), util.describe_diff(mom, mom_o)
    fy_o = O.filter_shadow_map(mom_o, W, H, 9, tech)
    assert util.bits_equal(fy, fy_o), util.describe_diff(fy, fy_o)
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility_moments(po, cam, fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), fy_o)
    ok = ~np.isnan(vis_o)
    assert np.array_equal(np.isnan(vis), ~ok) and util.bits_equal(vis[ok], vis_o[ok]), util.describe_diff(vis[ok], vis_o[ok])
    # degenerate triangles, a NaN vertex, a triangle behind the light
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [np.nan, 0, 0], [0, 500, 900], [1, 500, 900], [0, 501, 900]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (7, 1))
    idx = np.array([[0, 0, 1], [0, 1, 2], [3, 1, 2], [4, 5, 6]], np.int32)
    ctx.set_mesh(xyz, nrm, idx)
    ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    mom_d = O.raster_moments(xyz, idx, fm["light_mvp"], SW, SH, tech)
    assert util.bits_equal(ctx.read("moments"), mom_d)
    assert util.bits_equal(ctx.read("moments_filtered"), O.filter_shadow_map(mom_d, W, H, 9, tech))
    # no triangles at all
    ctx.set_mesh(xyz, nrm, np.zeros((0, 3), np.int32))
    ctx.render_shadow_map(); ctx.filter_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    mom_e = ctx.read("moments")
    assert (mom_e == np.array([0, 0, 0, 1], np.float32)).all()
    assert util.bits_equal(ctx.read("moments_filtered"), O.filter_shadow_map(mom_e, W, H, 9, tech))
    assert (ctx.read("visibility") == 0.0).all()                       # every pixel is background
    # restore a valid mesh for the tests that follow on this context
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
