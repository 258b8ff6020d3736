"""Shadow volumes the way north_star (4) words them: silhouette extraction (the side quads of interior edges cancel in pairs and
are not drawn) and depth-fail counting over capped volumes.  The parity target stays the reference's per-triangle, depth-pass
stencil count (ShadowVolumes/src/ShadowVolume.cpp:116-195, main.cpp:160-172); the two modes are checked against it:
  * depth-fail == depth-pass, count for count, on every foreground pixel (eye outside the volumes, nothing near-clipped);
  * silhouette == per-triangle except on depth-equality pixels: the two quads of a cancelling pair are triangulated along different
    diagonals (the reference's index pattern), so where the scene depth sits within rounding of the quad's depth one of them can
    pass and the other fail.  Measured: <= 0.01 % of the foreground pixels change their shadow mask (budget: 0.05 %).
CPU tests pin the oracle's two modes against its per-triangle form; GPU tests (-m gpu) compare CUDA with the oracle bit for bit."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util


def _per_triangle(sc, W, H):
    fm = util.frame(sc, W, H, 64)
    _, _, dep = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    pxyz, pidx = O.sv_build_prisms(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"])
    cnt, _ = O.sv_count(pxyz, pidx, fm["cam_mvp"], W, H, dep)
    return fm, dep, cnt


@pytest.mark.parametrize("name", ["teapot", "raptor", "door"])
def test_oracle_silhouette_keeps_what_does_not_cancel(name):
    sc = util.scene(name)
    idx, nrm, L = sc["idx"], sc["nrm"], sc["light_eye"]
    keep = O.sv_silhouette_keep(nrm, idx, L)
    # independent restatement with numpy: per (undirected edge, class) the surplus of one direction survives
    n = (nrm[idx[:, 0]] + nrm[idx[:, 1]] + nrm[idx[:, 2]]) / np.float32(3)
    cls = (n[:, 0] * L[0] + n[:, 1] * L[1] + n[:, 2] * L[2]) >= 0
    a, b = idx, np.roll(idx, -1, axis=1)
    key = np.minimum(a, b).astype(np.int64) << 32 | np.maximum(a, b)
    fwd = a < b
    from collections import defaultdict
    tally = defaultdict(lambda: [0, 0])
    for t in range(idx.shape[0]):
        for e in range(3):
            tally[(int(key[t, e]), bool(cls[t]))][int(fwd[t, e])] += 1
    expect = sum(abs(v[0] - v[1]) for v in tally.values())
    assert int(keep.sum()) == expect
    assert 0 < keep.sum() < keep.size // 2                      # a closed-ish mesh: most interior quads are gone


@pytest.mark.parametrize("name,W,H", [("teapot", 320, 240), ("raptor", 320, 240), ("door", 200, 150)])
def test_oracle_silhouette_and_zfail_against_the_per_triangle_count(name, W, H):
    sc = util.scene(name)
    fm, dep, cnt = _per_triangle(sc, W, H)
    fg = dep < 1.0
    keep = O.sv_silhouette_keep(sc["nrm"], sc["idx"], sc["light_eye"])
    pxyz, vidx = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], keep=keep)
    cnt_s, _ = O.sv_count_ex(pxyz, vidx, fm["cam_mvp"], W, H, dep)
    mask_changed = ((cnt != 0) != (cnt_s != 0)) & fg
    assert mask_changed.sum() <= 0.0005 * fg.sum(), f"{mask_changed.sum()} of {fg.sum()} foreground pixels"       # the 0.05 % budget
    assert (cnt != cnt_s).sum() <= 0.002 * cnt.size
    assert not ((cnt != cnt_s) & ~fg).any()
    # depth-fail over capped volumes: the same counts on the foreground, per-triangle and silhouette alike
    pxyz8, v8 = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], caps=True)
    cnt_f, _ = O.sv_count_ex(pxyz8, v8, fm["cam_mvp"], W, H, dep, zfail=True, per=8)
    assert np.array_equal(cnt_f[fg], cnt[fg])
    pxyz8s, v8s = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], keep=keep, caps=True)
    cnt_fs, _ = O.sv_count_ex(pxyz8s, v8s, fm["cam_mvp"], W, H, dep, zfail=True, per=8)
    assert np.array_equal(cnt_fs[fg], cnt_s[fg])


def test_oracle_volume_layout():
    sc = util.scene("door")
    T = sc["idx"].shape[0]
    pxyz, pidx = O.sv_build_prisms(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"])
    px6, v6 = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"])
    assert np.array_equal(pxyz, px6) and np.array_equal(pidx, v6)                      # no keep, no caps: ShadowVolume::update as is
    px8, v8 = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], caps=True)
    v8 = v8.reshape(T, 8, 3)
    assert np.array_equal(v8[:, :6].reshape(-1, 3), pidx)
    assert np.array_equal(v8[:, 6, 0], np.arange(T) * 6)                              # the near cap starts at the triangle's first vertex
    assert set(np.unique(v8[:, 6] - (np.arange(T) * 6)[:, None])) == {0, 1, 2}
    assert set(np.unique(v8[:, 7] - (np.arange(T) * 6)[:, None])) == {3, 4, 5}


# ---------------------------------------------------------------- CUDA vs oracle
@pytest.fixture(scope="module")
def ctx():
    from globalillumination_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _gpu_counts(ctx, sc, W, H, **kw):
    po, pg = util.params_pair("hard", 64, **kw)
    fm = util.frame(sc, W, H, 64)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], 64, 64)
    ctx.set_params(pg)
    ctx.render_gbuffer()
    ctx.compute_shadow_volume(sc["light_eye"])
    return fm, ctx.read("cam_depth"), ctx.read("sv_count"), ctx.read("sv_stencil"), ctx.read("sv_prism_xyz"), ctx.read("sv_prism_idx")


@pytest.mark.gpu
@pytest.mark.parametrize("name,W,H,silhouette,zfail", [("teapot", 640, 480, 1, 0), ("raptor", 640, 480, 1, 0), ("door", 320, 240, 1, 0),
                                                       ("teapot", 640, 480, 0, 1), ("raptor", 333, 217, 1, 1), ("tree", 640, 480, 1, 0),
                                                       ("tree", 320, 240, 1, 1)])
def test_shadow_volume_modes_bit_exact_vs_oracle(ctx, name, W, H, silhouette, zfail):
    sc = util.scene(name)
    fm, dep, cnt, st, pxyz, pidx = _gpu_counts(ctx, sc, W, H, sv_silhouette=silhouette, sv_zfail=zfail)
    keep = O.sv_silhouette_keep(sc["nrm"], sc["idx"], sc["light_eye"]) if silhouette else None
    pxyz_o, vidx_o = O.sv_build_volumes(sc["xyz"], sc["nrm"], sc["idx"], sc["light_eye"], keep=keep, caps=bool(zfail))
    assert util.bits_equal(pxyz, pxyz_o)
    assert np.array_equal(pidx, vidx_o), "volume triangles (dropped quads, caps)"
    cnt_o, st_o = O.sv_count_ex(pxyz_o, vidx_o, fm["cam_mvp"], W, H, dep, zfail=bool(zfail), per=8 if zfail else 6)
    assert np.array_equal(cnt, cnt_o), util.describe_diff(cnt, cnt_o)
    assert np.array_equal(st, st_o)
    assert (cnt_o != 0).mean() > 0.005


@pytest.mark.gpu
def test_silhouette_and_zfail_agree_with_the_reference_form_on_the_gpu(ctx):
    """The equalities the CPU tests establish for the oracle, on the CUDA path at the reference's window size."""
    sc = util.scene("tree")
    W, H = 640, 480
    _, dep, cnt, *_ = _gpu_counts(ctx, sc, W, H)
    fg = dep < 1.0
    _, _, cnt_s, *_ = _gpu_counts(ctx, sc, W, H, sv_silhouette=1)
    assert (((cnt != 0) != (cnt_s != 0)) & fg).sum() <= 0.0005 * fg.sum()
    _, _, cnt_f, *_ = _gpu_counts(ctx, sc, W, H, sv_zfail=1)
    assert (cnt_f[fg] != cnt[fg]).sum() <= 2                              # (one pixel differs in the oracle too: a volume grazing the near plane)
    ctx.set_option("sv_count_fragments", 1)
    _gpu_counts(ctx, sc, W, H)
    n_all = ctx.sv_fragments()
    _gpu_counts(ctx, sc, W, H, sv_silhouette=1)
    n_sil = ctx.sv_fragments()
    ctx.set_option("sv_count_fragments", 0)
    assert 0 < n_sil < n_all


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(sv_silhouette=1), dict()], ids=["silhouette", "per_triangle"])
def test_c4_tree_shadow_volumes_at_1080p_bit_exact(ctx, kw):
    """the 1080p shadow-volume workload was benchmarked but never compared with the oracle.  (The per-triangle form
    costs the CPU oracle about a minute at this size: it runs with SGI_SLOW_TESTS=1; the bench workload is the silhouette form.)"""
    import os
    if not kw and os.environ.get("SGI_SLOW_TESTS") != "1":
        pytest.skip("per-triangle prisms at 1080p: set SGI_SLOW_TESTS=1 (about a minute of CPU oracle)")
    sc = util.scene("tree")
    W, H = 1920, 1080
    fm, dep, cnt, st, pxyz, pidx = _gpu_counts(ctx, sc, W, H, **kw)
    cnt_o, st_o = O.sv_count_ex(pxyz, pidx, fm["cam_mvp"], W, H, dep)
    assert np.array_equal(cnt, cnt_o), util.describe_diff(cnt, cnt_o)
    assert np.array_equal(st, st_o)


@pytest.mark.gpu
@pytest.mark.parametrize("name,W,H,kw", [("tree", 640, 480, dict(sv_silhouette=1)), ("tree", 320, 240, dict(sv_silhouette=1, sv_zfail=1)), ("raptor", 333, 217, dict())])
def test_hot_tiles_shared_by_list_segment_or_by_region_count_the_same(ctx, name, W, H, kw):
    """Option "sv_split_lists": hot tiles shared between CTAs by list segment (counts added atomically, the default) or by
    sub-region; option "tile_few_walk": the binner's tile-major walk for passes of few tiles or the pair walk: identical counts
    and stencil values."""
    sc = util.scene(name)
    _, _, cnt1, st1, *_ = _gpu_counts(ctx, sc, W, H, **kw)
    for option in ("sv_split_lists", "tile_few_walk"):
        ctx.set_option(option, 0)
        try:
            _, _, cnt0, st0, *_ = _gpu_counts(ctx, sc, W, H, **kw)
        finally:
            ctx.set_option(option, 1)
        assert np.array_equal(cnt0, cnt1) and np.array_equal(st0, st1), option
    assert (cnt1 != 0).mean() > 0.005
