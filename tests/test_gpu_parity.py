"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs — bit-exact.

Bar (BASELINE.json north_star): hard masks and shadow-volume counts bit-exact, soft visibility within 1e-3.
The oracle and the kernels evaluate the same fp32 expressions in the same order with FMA contraction off
on both sides, so here EVERYTHING (depth maps, G-buffer, all visibilities, counts) is required to be
bit-identical; the 1e-3 tolerance is asserted separately so a future relaxation stays visible.
"""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util

pytestmark = pytest.mark.gpu

SOFT_TOL = 1e-3
TECHS = ["hard", "pcf", "pcss", "rbsm_noncons", "rbsm_cons", "rpcf_noncons", "rpcf_cons", "rsmss", "rbssm", "pcf_tricubic"]


@pytest.fixture(scope="module")
def ctx():
    from globalillumination_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def setup_frame(ctx, sc, W, H, S, pg):
    fm = util.frame(sc, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    return fm


@pytest.mark.parametrize("name,W,H,S", [("teapot", 320, 180, 256), ("door", 200, 150, 128), ("teapot", 1280, 720, 1024),
                                        ("raptor", 333, 217, 300)])
def test_depth_and_gbuffer_bit_exact(ctx, name, W, H, S):
    sc = util.scene(name)
    po, pg = util.params_pair("hard", S)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map()
    ctx.render_gbuffer()
    sm = ctx.read("shadow_map")[0]
    pos, nrm, dep = ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("cam_depth")
    sm_o = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    pos_o, nrm_o, dep_o = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    assert util.bits_equal(sm, sm_o), "shadow map: " + util.describe_diff(sm, sm_o)
    assert util.bits_equal(dep, dep_o), "camera depth: " + util.describe_diff(dep, dep_o)
    assert util.bits_equal(pos, pos_o), "G-buffer position: " + util.describe_diff(pos, pos_o)
    assert util.bits_equal(nrm, nrm_o), "G-buffer normal: " + util.describe_diff(nrm, nrm_o)
    assert (sm_o < 1).mean() > 0.05 and (dep_o < 1).mean() > 0.05     # the frame is not empty


@pytest.mark.parametrize("tech", TECHS)
@pytest.mark.parametrize("name,W,H,S", [("teapot", 320, 180, 256), ("teapot", 1280, 720, 1024)])
def test_visibility_bit_exact(ctx, name, W, H, S, tech):
    if tech in ("rpcf_noncons", "rpcf_cons", "rsmss", "rbssm") and W > 640:
        W, H = 640, 360                                   # the oracle needs seconds per frame for 64x RBSM taps
    sc = util.scene(name)
    po, pg = util.params_pair(tech, S, depth_threshold=float(sc["depth_threshold"]))
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map()
    ctx.render_gbuffer()
    ctx.compute_visibility()
    vis = ctx.read("visibility")
    sm, pos, nrm = ctx.read("shadow_map")[0], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm")
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility(po, cam, fm["light_mvp_b"], pos, nrm, sm)
    assert np.abs(vis - vis_o).max() <= SOFT_TOL
    assert np.array_equal(vis == 1.0, vis_o == 1.0), "hard mask: " + util.describe_diff(vis == 1.0, vis_o == 1.0)
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
    fg = pos[..., 0] != 0
    assert 0.05 < (vis_o[fg] == 1.0).mean() < 0.999        # both lit and shadowed pixels exist


@pytest.mark.parametrize("tech,name,W,H,S,kw", [("edtsm_noncons", "teapot", 320, 180, 256, dict(penumbra_size=5)),
                                                ("edtsm_cons", "teapot", 640, 360, 512, dict(penumbra_size=3, kernel_order=9)),
                                                ("edtsm_noncons", "dragon", 1280, 720, 1024, dict(penumbra_size=10)),
                                                ("edtsm_noncons", "door", 333, 217, 300, dict(penumbra_size=1, shadow_intensity=0.5))])
def test_edt_shadow_mapping_bit_exact(ctx, tech, name, W, H, S, kw):
    """EDTSM (next row f3): nearest-site map (exact Voronoi diagram) identical to the oracle's, final filtered visibility
    bit-exact; the penumbra ramp exists."""
    sc = util.scene(name)
    po, pg = util.params_pair(tech, S, depth_threshold=float(sc["depth_threshold"]), **kw)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    vis, near = ctx.read("visibility"), ctx.read("edt_nearest")
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o, near_o = O.edtsm(po, cam, fm["cam_mvp"], fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    assert np.array_equal(near, near_o), f"nearest site differs at {int((near != near_o).any(-1).sum())} pixels"
    assert np.abs(vis - vis_o).max() <= SOFT_TOL
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
    assert (near_o[..., 0] != O.EDT_MARKER).all() and ((vis_o > po.shadow_intensity) & (vis_o < 1.0)).sum() > 100


def test_rbssm_work_list_form_equals_thread_per_pixel_form(ctx):
    """RBSSM runs as a work list of penumbra pixels with one warp per pixel; the plain one-thread-per-pixel kernel must give
    the same bits (tap contributions are summed in the shader's loop order in both)."""
    sc = util.scene("dragon")
    W, H, S = 640, 360, 1024
    for kw in (dict(), dict(kernel_size=8, light_source_radius=16), dict(kernel_size=19, max_search=6)):
        po, pg = util.params_pair("rbssm", S, depth_threshold=float(sc["depth_threshold"]), **kw)
        setup_frame(ctx, sc, W, H, S, pg)
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        a = ctx.read("visibility")
        ctx.set_option("rbssm_compact", 0)
        try:
            ctx.compute_visibility()
            b = ctx.read("visibility")
        finally:
            ctx.set_option("rbssm_compact", 1)
        assert util.bits_equal(a, b), (kw, util.describe_diff(a, b))
        assert ((a > po.shadow_intensity) & (a < 1.0)).sum() > 1000


def test_edt_shadow_mapping_without_any_shadow_boundary(ctx):
    """A frame with no site (light inside nothing: everything lit or everything failing the site test) keeps the hard
    shadows and reports MARKER everywhere."""
    xyz = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], np.float32)
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    sc = dict(xyz=xyz, nrm=nrm, idx=idx, cam_eye=np.array([0, 41, -50], np.float32), cam_at=np.array([0, 0, 0], np.float32),
              light_eye=np.array([10, 130, 100], np.float32), light_at=np.zeros(3, np.float32))
    W, H, S = 160, 90, 128
    po, pg = util.params_pair("edtsm_noncons", S)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    near = ctx.read("edt_nearest")
    assert (near == O.EDT_MARKER).all()
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o, near_o = O.edtsm(po, cam, fm["cam_mvp"], fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    assert np.array_equal(near, near_o) and util.bits_equal(ctx.read("visibility"), vis_o)


@pytest.mark.parametrize("kw", [dict(kernel_order=9, shadow_intensity=0.5), dict(kernel_order=15, penumbra_size=2),
                                dict(kernel_size=7, blocker_search_size=5, light_source_radius=4), dict(max_search=4, kernel_order=3)])
@pytest.mark.parametrize("tech", ["pcf", "pcss", "rbsm_noncons", "rpcf_cons", "rsmss", "rbssm"])
def test_visibility_parameter_sweep(ctx, tech, kw):
    sc = util.scene("teapot")
    W, H, S = 256, 144, 200
    po, pg = util.params_pair(tech, S, depth_threshold=float(sc["depth_threshold"]), **kw)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    vis = ctx.read("visibility")
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility(po, cam, fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)


def test_visibility_rect_only_touches_rect(ctx):
    sc = util.scene("teapot")
    W, H, S = 320, 180, 256
    po, pg = util.params_pair("pcf", S)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    full = ctx.read("visibility")
    po, pg = util.params_pair("hard", S, rect_x0=37, rect_y0=11, rect_x1=200, rect_y1=97)
    ctx.set_params(pg)
    ctx.compute_visibility()
    part = ctx.read("visibility")
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    ref = full.copy()
    sub = O.visibility(po, cam, fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    ref[11:97, 37:200] = sub[11:97, 37:200]
    assert util.bits_equal(part, ref), util.describe_diff(part, ref)


@pytest.mark.parametrize("n_lights,S", [(4, 128), (16, 256)])
def test_many_light_bit_exact(ctx, n_lights, S):
    sc = util.scene("teapot")
    W, H = 320, 180
    po, pg = util.params_pair("multi_hard", S)
    fm = util.frame(sc, W, H, S)
    mvp, mvpb = util.multi_lights(sc, n_lights, 16, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(mvp, mvpb, fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    maps = ctx.read("shadow_map")
    for i in range(n_lights):
        m_o = O.raster_depth(sc["xyz"], sc["idx"], mvp[i], S, S)
        assert util.bits_equal(maps[i], m_o), f"light {i}: " + util.describe_diff(maps[i], m_o)
    vis = ctx.read("visibility")
    vis_o = O.visibility_multi(po, mvpb[-1], mvpb[:, 12:16], ctx.read("gbuf_pos"), maps)
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
    assert len(np.unique(vis_o)) > 3                        # penumbra levels exist


@pytest.mark.parametrize("name,W,H,func", [("door", 320, 240, O.DEPTH_LEQUAL), ("door", 320, 240, O.DEPTH_LESS),
                                           ("teapot", 640, 480, O.DEPTH_LEQUAL), ("raptor", 640, 480, O.DEPTH_LEQUAL)])
def test_shadow_volume_counts_bit_exact(ctx, name, W, H, func):
    sc = util.scene(name)
    S = 64
    po, pg = util.params_pair("hard", S, sv_depth_func=func)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_gbuffer()
    ctx.compute_shadow_volume(sc["light_eye"])
    cnt, st = ctx.read("sv_count"), ctx.read("sv_stencil")
    pxyz, pidx = ctx.read("sv_prism_xyz"), ctx.read("sv_prism_idx")
    pxyz_o, pidx_o = O.sv_build_prisms(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], 100)
    assert util.bits_equal(pxyz, pxyz_o), "prism vertices: " + util.describe_diff(pxyz, pxyz_o)
    assert np.array_equal(pidx, pidx_o)
    dep = ctx.read("cam_depth")
    cnt_o, st_o = O.sv_count(pxyz_o, pidx_o, fm["cam_mvp"], W, H, dep, func)
    assert np.array_equal(cnt, cnt_o), "counts: " + util.describe_diff(cnt, cnt_o)
    assert np.array_equal(st, st_o)
    assert (cnt_o != 0).mean() > 0.01


@pytest.mark.parametrize("tech", ["hard", "pcf", "pcss", "rbsm_noncons", "rbssm", "edtsm_noncons"])
def test_non_square_shadow_map(ctx, tech):
    """Shadow map width != height (the light frustum's aspect follows, displaySceneFromLightPOV main.cpp:255): depth map and
    visibility bit-exact."""
    sc = util.scene("teapot")
    W, H, SW, SH = 320, 180, 384, 200
    fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, SW, SH)
    po, pg = util.params_pair(tech, SW, depth_threshold=float(sc["depth_threshold"]), shadow_map_height=SH, penumbra_size=5 if tech.startswith("edt") else 1)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], SW, SH)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    sm = ctx.read("shadow_map")[0]
    assert sm.shape == (SH, SW)
    sm_o = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], SW, SH)
    assert util.bits_equal(sm, sm_o), util.describe_diff(sm, sm_o)
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    pos, nrm = ctx.read("gbuf_pos"), ctx.read("gbuf_nrm")
    if tech.startswith("edt"):
        vis_o, _ = O.edtsm(po, cam, fm["cam_mvp"], fm["light_mvp_b"], pos, nrm, sm)
    else:
        vis_o = O.visibility(po, cam, fm["light_mvp_b"], pos, nrm, sm)
    vis = ctx.read("visibility")
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
    fg = pos[..., 0] != 0
    assert 0.02 < (vis_o[fg] < 1.0).mean() < 0.98


def test_tile_list_overflow_is_reported_and_recovered():
    """When a frame needs more tile-list space than the context sized from earlier frames, the call that next touches the
    host reports SGI_ERR_OVERFLOW, the lists are re-sized, and running the frame again gives the right answer."""
    from globalillumination_b200 import capi
    c = capi.Context(0)
    try:
        W, H, S = 640, 480, 64
        small, big = util.scene("door"), util.scene("tree")
        po, pg = util.params_pair("hard", S)
        fm = setup_frame(c, small, W, H, S, pg)
        c.render_gbuffer(); c.compute_shadow_volume(small["light_eye"]); c.synchronize()      # lists sized for a few prisms
        fm = setup_frame(c, big, W, H, S, pg)                                                  # 229 200 prism triangles now
        errors = 0
        for attempt in range(6):
            try:
                c.render_gbuffer(); c.compute_shadow_volume(big["light_eye"])
                cnt = c.read("sv_count")
                break
            except capi.SgiError as e:
                assert e.code == -4 and "overflow" in str(e)
                errors += 1
        else:
            raise AssertionError("the lists never became large enough")
        assert errors >= 1, "expected at least one overflow report on the way"
        depth = c.read("cam_depth")
        pxyz, pidx = O.sv_build_prisms(big["xyz"], big["nrm"], big["idx"], big["light_eye"])
        cnt_o, _ = O.sv_count(pxyz, pidx, fm["cam_mvp"], W, H, depth)
        assert np.array_equal(cnt, cnt_o)
    finally:
        c.close()


def test_tile_list_overflow_on_the_pipelined_path():
    """with frames pipelined (sgi_read_async, the next frame queued before the previous ticket is waited for) an overflow
    cannot be charged to one frame; every ticket issued before the overflow was seen must report it - no truncated frame is
    delivered as good - and re-issued frames come out right."""
    import ctypes as C
    from globalillumination_b200 import capi
    c = capi.Context(0)
    try:
        lib = c.lib
        W, H, S = 640, 480, 64
        small, big = util.scene("door"), util.scene("tree")
        po, pg = util.params_pair("hard", S)
        fm = setup_frame(c, small, W, H, S, pg)
        c.render_gbuffer(); c.compute_shadow_volume(small["light_eye"]); c.synchronize()      # lists sized for a few prisms
        fm = setup_frame(c, big, W, H, S, pg)
        ptrs, bufs = [], []
        for _ in range(2):
            p = C.c_void_p()
            assert lib.sgi_alloc_host(C.byref(p), C.c_size_t(W * H * 4)) == 0
            ptrs.append(p); bufs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), (H, W)))
        depth = None
        good, reports = [], 0
        for rnd in range(8):
            tickets = []
            for k in range(2):                                                               # two frames in flight
                c.render_gbuffer(); c.compute_shadow_volume(big["light_eye"])
                tickets.append(c.read_async("sv_count", ptrs[k].value, W * H * 4))
            ok = []
            for k, t in enumerate(tickets):
                try:
                    c.read_wait(t); ok.append(k)
                except capi.SgiError as e:
                    assert e.code == -4
                    reports += 1
            if len(ok) == 2:
                good = [bufs[0].copy(), bufs[1].copy()]
                break
        assert reports >= 1 and good, (reports, len(good))
        depth = c.read("cam_depth")
        pxyz, pidx = O.sv_build_prisms(big["xyz"], big["nrm"], big["idx"], big["light_eye"])
        cnt_o, _ = O.sv_count(pxyz, pidx, fm["cam_mvp"], W, H, depth)
        assert np.array_equal(good[0], cnt_o) and np.array_equal(good[1], cnt_o)
        c.synchronize()
        for p in ptrs:
            lib.sgi_free_host(p)
    finally:
        c.close()


def test_empty_and_degenerate_inputs(ctx):
    from globalillumination_b200 import capi
    sc = util.scene("door")
    W, H, S = 64, 48, 32
    po, pg = util.params_pair("hard", S)
    fm = util.frame(sc, W, H, S)
    # degenerate triangles (repeated vertex), a triangle behind the camera, NaN vertex
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [np.nan, 0, 0], [0, 500, -900], [1, 500, -900], [0, 501, -900]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (7, 1))
    idx = np.array([[0, 0, 1], [0, 1, 2], [3, 1, 2], [4, 5, 6]], np.int32)
    ctx.set_mesh(xyz, nrm, idx)
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    sm_o = O.raster_depth(xyz, idx, fm["light_mvp"], S, S)
    pos_o, nrm_o, dep_o = O.raster_gbuffer(xyz, nrm, idx, fm["cam_mvp"], W, H)
    assert util.bits_equal(ctx.read("shadow_map")[0], sm_o)
    assert util.bits_equal(ctx.read("gbuf_pos"), pos_o)
    assert util.bits_equal(ctx.read("cam_depth"), dep_o)
    # zero triangles: cleared outputs
    ctx.set_mesh(xyz, nrm, np.zeros((0, 3), np.int32))
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    assert (ctx.read("shadow_map") == 1.0).all() and (ctx.read("visibility") == 0.0).all()
    assert (ctx.read("gbuf_pos")[..., 3] == 1.0).all() and (ctx.read("gbuf_pos")[..., :3] == 0.0).all()
    # error paths: bad index, pass before inputs
    with pytest.raises(capi.SgiError):
        ctx.set_mesh(xyz, nrm, np.array([[0, 1, 99]], np.int32))
    c2 = capi.Context(0)
    with pytest.raises(capi.SgiError):
        c2.render_shadow_map()
    c2.close()


def test_many_light_shards_sum_to_the_single_context_result(ctx):
    """Light sharding: partial sums of two shards (common matrix = last light of the WHOLE set) add up
    bit-exactly to the un-sharded result, and each shard matches the oracle."""
    sc = util.scene("teapot")
    W, H, S, n_l = 256, 144, 128, 6
    fm = util.frame(sc, W, H, S)
    mvp, mvpb = util.multi_lights(sc, n_l, 16, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    po, pg = util.params_pair("multi_hard", S)
    ctx.set_multi_light_common(None)
    ctx.set_lights(mvp, mvpb, fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    full = ctx.read("visibility")
    pos = ctx.read("gbuf_pos")
    total = np.zeros_like(full)
    po, pg = util.params_pair("multi_hard", S, multi_partial=1)
    for r in range(2):
        own = list(range(r, n_l, 2))
        ctx.set_lights(mvp[own], mvpb[own], fm["light_pos_shading"], S, S)
        ctx.set_multi_light_common(mvpb[-1])
        ctx.set_params(pg)
        ctx.render_shadow_map(); ctx.compute_visibility()
        part = ctx.read("visibility")
        maps = ctx.read("shadow_map")
        part_o = O.visibility_multi(po, mvpb[-1], mvpb[own][:, 12:16], pos, maps)
        assert util.bits_equal(part, part_o), util.describe_diff(part, part_o)
        total += part
    ctx.set_multi_light_common(None)
    assert util.bits_equal(total / np.float32(n_l), full), util.describe_diff(total / np.float32(n_l), full)


# ---- BASELINE.json configs at (or near) their full sizes -------------------------------------------------------------
def _full_frame(ctx, sc, W, H, S, tech, **kw):
    po, pg = util.params_pair(tech, S, depth_threshold=float(sc["depth_threshold"]), **kw)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    return po, fm


def test_config_c3_dragon_rbsm_4k(ctx):
    """c3: Dragon, 3840x2160, 4096^2 map, RBSM conservative + non-conservative: every buffer bit-exact vs the oracle."""
    sc = util.scene("dragon")
    W, H, S = 3840, 2160, 4096
    po, fm = _full_frame(ctx, sc, W, H, S, "rbsm_noncons")
    sm, pos, nrm = ctx.read("shadow_map")[0], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm")
    sm_o = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    assert util.bits_equal(sm, sm_o), util.describe_diff(sm, sm_o)
    pos_o, nrm_o, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    assert util.bits_equal(pos, pos_o) and util.bits_equal(nrm, nrm_o)
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis = ctx.read("visibility")
    vis_o = O.visibility(po, cam, fm["light_mvp_b"], pos, nrm, sm)
    assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
    po2, pg2 = util.params_pair("rbsm_cons", S, depth_threshold=float(sc["depth_threshold"]))
    ctx.set_params(pg2); ctx.compute_visibility()
    vis2, vis2_o = ctx.read("visibility"), O.visibility(po2, cam, fm["light_mvp_b"], pos, nrm, sm)
    assert util.bits_equal(vis2, vis2_o), util.describe_diff(vis2, vis2_o)
    # revectorisation only moves shadow boundaries: both variants agree with the plain hard test away from edges
    po3, pg3 = util.params_pair("hard", S)
    ctx.set_params(pg3); ctx.compute_visibility()
    hard = ctx.read("visibility")
    assert ((vis != hard).mean() < 0.02) and ((vis2 != hard).mean() < 0.02) and (vis2 <= hard + 1e-6).all()   # conservative only adds shadow


def test_config_c2_sponza_full_size_through_the_host(ctx):
    """c2: the bench workload at full size through the C++ host (SceneLoader -> ShadowApp), PCSS and PCF, vs the oracle."""
    from globalillumination_b200 import hostapi, scenes
    cfg = scenes.write_config("c2_sponza")
    w = scenes.WORKLOADS["c2_sponza"]
    W, H, S = w["W"], w["H"], w["S"]
    app = hostapi.App(0)
    try:
        app.load_scene(cfg); app.configure(W, H, S); app.set_technique("pcss"); app.set(**w["params"])
        app.display("soft_shadow_mapping")
        c = app.context()
        vis, sm, pos, nrm = c.read("visibility"), c.read("shadow_map")[0], c.read("gbuf_pos"), c.read("gbuf_nrm")
        sc = hostapi.load_scene(cfg)
        fm = util.frame(sc, W, H, S)
        sm_o = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
        pos_o, nrm_o, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        assert util.bits_equal(sm, sm_o), util.describe_diff(sm, sm_o)
        assert util.bits_equal(pos, pos_o), util.describe_diff(pos, pos_o)
        assert util.bits_equal(nrm, nrm_o), util.describe_diff(nrm, nrm_o)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        vis_o = O.visibility(O.default_params("pcss", S), cam, fm["light_mvp_b"], pos_o, nrm_o, sm_o)
        assert util.bits_equal(vis, vis_o), util.describe_diff(vis, vis_o)
        app.set_technique("pcf"); app.display("shadow_mapping")
        vis_pcf = c.read("visibility")
        vis_pcf_o = O.visibility(O.default_params("pcf", S), cam, fm["light_mvp_b"], pos_o, nrm_o, sm_o)
        assert util.bits_equal(vis_pcf, vis_pcf_o), util.describe_diff(vis_pcf, vis_pcf_o)
        assert 0.3 < (vis_pcf_o == 1).mean() < 0.9
    finally:
        app.close()


def test_end_to_end_pipeline_frames_are_their_own(ctx):
    """The pipelined host-buffer path of bench.py's e2e (Mesh arrays page-locked in place and uploaded by DMA on the upload
    stream, three frames in flight, spare visibility buffer, asynchronous copy-out): every frame of an animated light must
    equal, bit for bit, the same frame rendered through the blocking call."""
    import ctypes as C
    from globalillumination_b200 import hostapi, scenes
    cfg = scenes.write_config("c2_sponza")
    w = scenes.WORKLOADS["c2_sponza"]
    W, H, S = w["W"], w["H"], w["S"]
    n_frames, depth, nbytes = 10, 3, W * H * 4
    lib = ctx.lib
    ptrs = []
    for _ in range(depth + 1):
        p = C.c_void_p()
        assert lib.sgi_alloc_host(C.byref(p), C.c_size_t(nbytes)) == 0
        ptrs.append(p)
    view = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (H, W))
    app = hostapi.App(0)
    try:
        app.load_scene(cfg); app.configure(W, H, S); app.set_technique("pcss"); app.set(**w["params"])
        app.set(animationOn=1, animation=-1800.0)
        blocking = []
        for k in range(n_frames):
            app.display_e2e("soft_shadow_mapping", "visibility", ptrs[depth].value, nbytes)
            blocking.append(view(ptrs[depth]).copy())
            app.step_animation(60.0)
        assert not util.bits_equal(blocking[0], blocking[-1])            # the light really moves
        app.set(animation=-1800.0)
        pending, got = [], []
        for k in range(n_frames):
            pending.append((app.display_e2e_async("soft_shadow_mapping", "visibility", ptrs[k % depth].value, nbytes), k))
            app.step_animation(60.0)
            if len(pending) >= depth:
                t, j = pending.pop(0)
                app.e2e_wait(t); got.append((j, view(ptrs[j % depth]).copy()))
        for t, j in pending:
            app.e2e_wait(t); got.append((j, view(ptrs[j % depth]).copy()))
        assert [j for j, _ in got] == list(range(n_frames))
        for j, img in got:
            assert util.bits_equal(img, blocking[j]), (j, util.describe_diff(img, blocking[j]))
    finally:
        app.close()
        for p in ptrs:
            lib.sgi_free_host(p)


def test_host_saves_the_shaded_frame_as_png(ctx, tmp_path):
    """Next row f1 end: shadeScene() + PNG output through the host; the decoded file is the SHADED buffer flipped to
    top-down and converted like a framebuffer (clamp, *255, round)."""
    from PIL import Image
    from globalillumination_b200 import hostapi, scenes
    app = hostapi.App(0)
    try:
        app.load_scene(scenes.write_config("c2_sponza")); app.configure(480, 270, 512); app.set_technique("pcf")
        app.display("shadow_mapping")
        path = tmp_path / "frame.png"
        app.save_image(path)
        shaded = app.context().read("shaded")
        want = np.rint(np.clip(shaded[::-1], 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
        got = np.asarray(Image.open(path).convert("RGBA"))
        assert got.shape == (270, 480, 4) and np.array_equal(got, want)
        assert len(np.unique(got[..., :3].reshape(-1, 3), axis=0)) > 100          # a real image, not a flat colour
    finally:
        app.close()


def test_lent_device_pointer_sees_each_frame_on_the_callers_stream(ctx):
    """A caller that holds the visibility buffer's device pointer and queues its own work on the context's stream (what
    bench.py's light-shard mode does with NCCL) must see every frame's result without any explicit synchronisation, although
    the shadow pass runs on an internal stream."""
    import torch

    class DevView:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    sc = util.scene("dragon")
    W, H, S = 1280, 720, 2048
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    try:
        po, pg = util.params_pair("pcss", S)
        util_frames = []
        fm0 = setup_frame(ctx, sc, W, H, S, pg)
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        ptr, nbytes = ctx.device_ptr("visibility")
        vis_t = torch.as_tensor(DevView(ptr, nbytes // 4), device="cuda:0")
        snaps, expect = [], []
        for k in range(6):
            light = sc["light_eye"] + np.array([6.0 * k, 0, -4.0 * k], np.float32)
            fm = util.frame(sc, W, H, S, light_eye=light)
            ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
            ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
            with torch.cuda.stream(stream):
                snaps.append(vis_t.clone())                     # queued on the context's stream right behind the frame
            util_frames.append(fm)
        for k in range(6):                                      # the same frames again, read through the blocking call
            fm = util_frames[k]
            ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
            ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
            expect.append(ctx.read("visibility"))
        torch.cuda.synchronize()
        for k in range(6):
            assert util.bits_equal(snaps[k].cpu().numpy().reshape(H, W), expect[k]), k
        assert not util.bits_equal(expect[0], expect[5])
    finally:
        ctx.synchronize()
        ctx.set_stream(None)


def test_config_c4_tree_shadow_volumes(ctx):
    """c4: TreeWithLeaves (the present half of it), 640x480 as in the reference: signed z-pass counts and 8-bit stencil."""
    sc = util.scene("tree")
    W, H, S = 640, 480, 64
    po, pg = util.params_pair("hard", S)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_gbuffer(); ctx.compute_shadow_volume(sc["light_eye"])
    cnt, st, dep = ctx.read("sv_count"), ctx.read("sv_stencil"), ctx.read("cam_depth")
    pxyz_o, pidx_o = O.sv_build_prisms(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], 100)
    assert util.bits_equal(ctx.read("sv_prism_xyz"), pxyz_o) and np.array_equal(ctx.read("sv_prism_idx"), pidx_o)
    cnt_o, st_o = O.sv_count(pxyz_o, pidx_o, fm["cam_mvp"], W, H, dep, O.DEPTH_LEQUAL)
    assert np.array_equal(cnt, cnt_o), util.describe_diff(cnt, cnt_o)
    assert np.array_equal(st, st_o) and (st_o != 0).mean() > 0.01


def test_config_c5_many_light_properties_at_full_size(ctx):
    """c5 at full size (7680x4320, 16 x 8192^2 maps) is beyond what the oracle finishes in seconds: size-independent
    properties instead.  (i) every light's map equals the same light rendered alone; (ii) the 16-light result equals the
    mean of 16 single-light hard results computed by the same kernels; (iii) value set and background."""
    sc = util.scene("sandiego")
    W, H, S, n_l = 7680, 4320, 8192, 16
    fm = util.frame(sc, W, H, S)
    mvp, mvpb = util.multi_lights(sc, n_l, 16, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    po, pg = util.params_pair("multi_hard", S)
    ctx.set_multi_light_common(None)
    ctx.set_lights(mvp, mvpb, fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    vis = ctx.read("visibility")
    pos = ctx.read("gbuf_pos")
    fg = pos[..., 0] != 0
    assert (vis[~fg] == 0).all() and fg.mean() > 0.3
    levels = np.unique(vis[fg])
    assert set(np.round((levels - 0.25) / 0.75 * 16).astype(int)) <= set(range(17)) and len(levels) > 4   # k/16 lit lights
    maps_crc = [int(np.bitwise_xor.reduce(m.view(np.uint32).ravel())) for m in ctx.read("shadow_map")]
    # (i)+(ii) on a strip (keeps the single-light re-runs cheap): partial sums over shards of one light each
    po1, pg1 = util.params_pair("multi_hard", S, multi_partial=1, rect_x0=0, rect_y0=2048, rect_x1=W, rect_y1=2048 + 256)
    acc = np.zeros((256, W), np.float32)
    for i in range(n_l):
        ctx.set_lights(mvp[i:i + 1], mvpb[i:i + 1], fm["light_pos_shading"], S, S)
        ctx.set_multi_light_common(mvpb[-1]); ctx.set_params(pg1)
        ctx.render_shadow_map(); ctx.compute_visibility()
        assert int(np.bitwise_xor.reduce(ctx.read("shadow_map")[0].view(np.uint32).ravel())) == maps_crc[i]
        acc += ctx.read("visibility")[2048:2048 + 256]
    ctx.set_multi_light_common(None)
    assert util.bits_equal(acc / np.float32(n_l), vis[2048:2048 + 256])
    # spot-check one light's map against the oracle at full size
    m0 = O.raster_depth(sc["xyz"], sc["idx"], mvp[3], S, S)
    assert int(np.bitwise_xor.reduce(m0.view(np.uint32).ravel())) == maps_crc[3]


def test_pipelined_readback_and_double_buffered_mesh(ctx):
    """sgi_read_async / sgi_read_wait with the next frame (a different mesh!) issued before the previous result is
    awaited: every result must be the one of its own frame."""
    import ctypes as C
    lib = ctx.lib
    W, H, S = 320, 180, 256
    names = ["teapot", "door", "raptor", "teapot"]
    bufs, ptrs = [], []
    for _ in names:
        p = C.c_void_p()
        assert lib.sgi_alloc_host(C.byref(p), C.c_size_t(W * H * 4)) == 0
        ptrs.append(p)
        bufs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (H, W)))
    expect, tickets = [], []
    for k, name in enumerate(names):
        sc = util.scene(name)
        po, pg = util.params_pair("pcf", S)
        fm = setup_frame(ctx, sc, W, H, S, pg)
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        tickets.append(ctx.read_async("visibility", ptrs[k].value, W * H * 4))
        sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
        pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        expect.append(O.visibility(po, cam, fm["light_mvp_b"], pos, nrm, sm))
    for k in range(len(names)):
        ctx.read_wait(tickets[k])
        assert util.bits_equal(bufs[k].copy(), expect[k]), (k, util.describe_diff(bufs[k], expect[k]))
    ctx.synchronize()
    for p in ptrs:
        lib.sgi_free_host(p)


@pytest.mark.parametrize("option,value", [("vis_staged", 1), ("overlap_passes", 0), ("tile_threads", 256), ("tile_threads", 512), ("tile_threads", 1024),
                                          ("tile_order", 0), ("tile_split", 0), ("tile_split", 16), ("tile_direct", 0), ("tile_direct", 128),
                                          ("tile_bin_big", 0), ("tile_bin_big_work", 0), ("sv_split_lists", 0), ("tile_static_items", 0), ("tile_static_items", 1),
                                          ("tile_refresh_full", 0), ("tile_refresh_full", 1)])
@pytest.mark.parametrize("tech,name,W,H,S", [("pcss", "teapot", 640, 360, 512), ("pcf", "raptor", 333, 217, 300), ("pcss", "dragon", 1920, 1080, 4096)])
def test_implementation_switches_do_not_change_results(ctx, option, value, tech, name, W, H, S):
    """Shared-memory staged taps, single-stream execution, every tile-CTA size, raster-order tile launch, no / aggressive
    hot-tile subdivision, the register path of short depth lists off / up to 128 entries, big records binned / tested per tile,
    both work-item scheduling forms: all give the same bits as the defaults."""
    sc = util.scene(name)
    po, pg = util.params_pair(tech, S, kernel_size=15 if name != "raptor" else 9)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    base = (ctx.read("visibility"), ctx.read("shadow_map"), ctx.read("gbuf_pos"))
    ctx.set_option(option, value)
    try:
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        alt = (ctx.read("visibility"), ctx.read("shadow_map"), ctx.read("gbuf_pos"))
    finally:
        ctx.set_option(option, {"vis_staged": 0, "overlap_passes": 1, "tile_threads": 0, "tile_order": 1, "tile_split": 256, "tile_direct": 32,
                                "tile_bin_big": 4096, "tile_bin_big_work": 1 << 20, "sv_split_lists": 1, "tile_static_items": 2, "tile_refresh_full": 2}[option])
    for a, b in zip(base, alt):
        assert util.bits_equal(a, b), (option, util.describe_diff(a, b))
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility(po, cam, fm["light_mvp_b"], base[2], ctx.read("gbuf_nrm"), base[1][0])
    assert util.bits_equal(alt[0], vis_o)


def test_albedo_target_and_phong_shading(ctx):
    """Next-row §8f: third G-buffer target (vertex colours) bit-exact; deferred Phong image within 2e-6 relative (powf is
    the only operation whose CPU and GPU implementations are not both correctly rounded)."""
    sc = util.scene("teapot")
    W, H, S = 640, 360, 512
    rgb = np.random.default_rng(5).uniform(0.05, 1.0, (len(sc["xyz"]), 3)).astype(np.float32)
    po, pg = util.params_pair("pcf", S)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    ctx.set_mesh_colors(rgb)
    try:
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility(); ctx.shade_phong()
        alb, img = ctx.read("gbuf_albedo"), ctx.read("shaded")
        pos, nrm, vis = ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("visibility")
        pos_o, nrm_o, alb_o, _ = O.raster_gbuffer_rgb(sc["xyz"], sc["nrm"], rgb, sc["idx"], fm["cam_mvp"], W, H)
        assert util.bits_equal(pos, pos_o) and util.bits_equal(nrm, nrm_o)
        assert util.bits_equal(alb, alb_o), util.describe_diff(alb, alb_o)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        img_o = O.shade_phong(cam, 0.25, pos, nrm, alb, vis)
        assert np.allclose(img, img_o, rtol=2e-6, atol=1e-7), float(np.abs(img - img_o).max())
        assert (img[pos[..., 0] == 0] == O.CLEAR_COLOR).all()
    finally:
        ctx.set_mesh_colors(None)
    ctx.render_gbuffer(); ctx.compute_visibility(); ctx.shade_phong()           # without colours: white albedo
    img2 = ctx.read("shaded")
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    img2_o = O.shade_phong(cam, 0.25, ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), None, ctx.read("visibility"))
    assert np.allclose(img2, img2_o, rtol=2e-6, atol=1e-7)


def test_pcss_early_out_option_is_exact(ctx):
    """Option "pcss_early_out": pixels with light-space depth in (0, 0.989) return 1.0 without taps.  The result must equal the
    oracle's (which runs every tap) bit for bit - on the Sponza-like light, where the shortcut covers every pixel, under the Teapot
    light, where it covers none, and with a light moved close enough that the depth range straddles the threshold."""
    from globalillumination_b200 import hostapi, scenes
    cfg = scenes.write_config("c2_sponza")
    sponza = hostapi.load_scene(cfg)
    teapot = util.scene("teapot")
    near = dict(teapot)
    near["light_eye"] = (np.asarray(teapot["light_eye"], np.float32) * np.float32(0.55)).astype(np.float32)      # depths around 0.989
    hit = []
    for sc, (W, H, S) in ((sponza, (480, 270, 512)), (teapot, (320, 180, 256)), (near, (320, 180, 256))):
        for kw in (dict(), dict(kernel_size=7, blocker_search_size=5, light_source_radius=16)):
            po, pg = util.params_pair("pcss", S, **kw)
            fm = setup_frame(ctx, sc, W, H, S, pg)
            ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
            plain = ctx.read("visibility").copy()
            ctx.set_option("pcss_early_out", 1)
            try:
                ctx.compute_visibility()
                fast = ctx.read("visibility").copy()
            finally:
                ctx.set_option("pcss_early_out", 0)
            cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
            pos, nrm, sm = ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0]
            vis_o = O.visibility(po, cam, fm["light_mvp_b"], pos, nrm, sm)
            assert util.bits_equal(plain, vis_o), util.describe_diff(plain, vis_o)
            assert util.bits_equal(fast, vis_o), util.describe_diff(fast, vis_o)
            # which share of the foreground the shortcut applies to (light-space depth of the pixel)
            p4 = pos.reshape(-1, 4).astype(np.float64) @ fm["light_mvp_b"].reshape(4, 4).astype(np.float64)
            z = p4[:, 2] / p4[:, 3]
            fg = pos.reshape(-1, 4)[:, 0] != 0
            hit.append(float(((z > 0) & (z < 0.989))[fg].mean()))
    assert hit[0] > 0.9 and hit[2] < 0.01 and 0.05 < hit[4] < 0.95, hit


# ---- min-max cull of the PCF / PCSS tap windows (block extrema of the depth map, dilated over the window's reach) ----------
@pytest.mark.parametrize("name,tech,W,H,S,kw,near_light", [
    ("teapot", "pcf", 640, 360, 512, {}, False), ("teapot", "pcss", 640, 360, 512, {}, False),
    ("dragon", "pcss", 480, 270, 1024, {}, False), ("dragon", "pcss", 480, 270, 300, dict(kernel_size=7), False),
    ("teapot", "pcf", 320, 180, 200, dict(kernel_order=9, penumbra_size=3), False),
    ("teapot", "pcss", 400, 300, 256, {}, True), ("teapot", "pcf", 400, 300, 256, {}, True)])
def test_minmax_cull_is_exact(ctx, name, tech, W, H, S, kw, near_light):
    """With the cull on and off (the default) the visibility is the same to the bit, and equal to the oracle, which runs every tap.
    near_light: a light frustum that does not contain the scene, so tap windows cross the map border (CLAMP_TO_BORDER depth 0)."""
    sc = dict(util.scene(name))
    if near_light:
        sc["light_eye"] = (np.asarray(sc["light_eye"], np.float32) * np.float32(0.25)).astype(np.float32)
    po, pg = util.params_pair(tech, S, depth_threshold=float(sc["depth_threshold"]), **kw)
    fm = setup_frame(ctx, sc, W, H, S, pg)
    out = []
    for on in (1, 0, 1):
        ctx.set_option("vis_minmax_cull", on)
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        out.append(ctx.read("visibility"))
    ctx.set_option("vis_minmax_cull", 0)
    assert util.bits_equal(out[0], out[1]) and util.bits_equal(out[0], out[2]), util.describe_diff(out[0], out[1])
    cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
    vis_o = O.visibility(po, cam, fm["light_mvp_b"], ctx.read("gbuf_pos"), ctx.read("gbuf_nrm"), ctx.read("shadow_map")[0])
    assert util.bits_equal(out[0], vis_o), util.describe_diff(out[0], vis_o)
    fg = vis_o > 0
    assert 0.02 < (vis_o[fg] == 1.0).mean() < 0.999                      # lit and shadowed regions both present


@pytest.mark.parametrize("direct,bin_big", [(0, 0), (32, 0), (0, 1), (128, 1), (32, 4096)])
def test_sparse_map_paths_match_the_oracle(ctx, direct, bin_big):
    """The city under a 4096^2 map (4096 tiles, a few hundred records that span more than 256 of them, most tiles holding only
    those): the register path of short lists and the binning of the big records, alone and together, against the oracle."""
    sc = util.scene("sandiego")
    S = 4096
    po, pg = util.params_pair("hard", S)
    fm = setup_frame(ctx, sc, 320, 180, S, pg)
    ctx.set_option("tile_direct", direct); ctx.set_option("tile_bin_big", bin_big); ctx.set_option("tile_bin_big_work", 0 if bin_big == 1 else 1 << 20)
    try:
        ctx.render_shadow_map(); ctx.render_shadow_map()                    # (the second pass knows the first one's big-record count)
        got = ctx.read("shadow_map")[0]
    finally:
        ctx.set_option("tile_direct", 32); ctx.set_option("tile_bin_big", 4096); ctx.set_option("tile_bin_big_work", 1 << 20)
    want = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    assert util.bits_equal(got, want), util.describe_diff(got, want)


def test_shared_reciprocal_divide_has_the_bits_of_the_plain_division(ctx):
    """The projective divides of the shadow pass share one reciprocal (sgi_internal.cuh sgi_div3): 3 x 2^30 quotients over random
    operand bits and over exponents where light-space coordinates live - every one identical to `a / b`."""
    for seed in (1, 2, 3, 4):
        assert ctx.divide_selftest(1 << 28, seed) == 0
