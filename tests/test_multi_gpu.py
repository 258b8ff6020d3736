"""Many-light fast path (primitive-id camera pass + fused position resolve) and the multi-GPU exchange of the C ABI
(sgi_comm_init / sgi_gather / sgi_reduce_lights).  Single-GPU tests run everywhere a GPU is; the 2-rank test needs two GPUs
(it launches scripts/multi_gpu_check.py under torchrun, NCCL) and is skipped on a one-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from globalillumination_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _many_light_frame(ctx, sc, W, H, S, n_lights, **kw):
    po, pg = util.params_pair("multi_hard", S, **kw)
    fm = util.frame(sc, W, H, S)
    mvp, mvpb = util.multi_lights(sc, n_lights, 16, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_multi_light_common(None)
    ctx.set_lights(mvp, mvpb, fm["light_pos_shading"], S, S)
    ctx.set_params(pg)
    return po, pg, fm, mvp, mvpb


@pytest.mark.parametrize("name,n_lights,S,W,H,eye", [("teapot", 4, 128, 320, 180, None), ("teapot", 16, 256, 333, 187, None),
                                                     ("teapot", 5, 200, 640, 360, ((0.0, 6.0, 6.0), (0.0, -19.0, 46.0))),   # camera close to the floor and the teapot: clipped triangles
                                                     ("sandiego", 16, 512, 960, 540, None)])
def test_many_light_fused_equals_the_gbuffer_path_and_the_oracle(ctx, name, n_lights, S, W, H, eye):
    sc = dict(util.scene(name))
    if eye is not None:
        sc["cam_eye"], sc["cam_at"] = np.asarray(eye[0], np.float32), np.asarray(eye[1], np.float32)
    po, pg, fm, mvp, mvpb = _many_light_frame(ctx, sc, W, H, S, n_lights)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    vis_a, pos, dep = ctx.read("visibility"), ctx.read("gbuf_pos"), ctx.read("cam_depth")
    maps = ctx.read("shadow_map")
    vis_o = O.visibility_multi(po, mvpb[-1], mvpb[:, 12:16], pos, maps)
    assert util.bits_equal(vis_a, vis_o), util.describe_diff(vis_a, vis_o)
    # the same frame from primitive ids: positions resolved inside the accumulation kernel
    _, pg_f = util.params_pair("multi_hard", S, multi_fused=1)
    ctx.set_params(pg_f)
    ctx.render_prim_ids(); ctx.compute_visibility()
    vis_b, ids = ctx.read("visibility"), ctx.read("prim_id")
    assert util.bits_equal(vis_a, vis_b), util.describe_diff(vis_a, vis_b)
    assert np.array_equal(ids != 0xFFFFFFFF, dep < 1.0)                  # a primitive wherever the depth pre-pass has a surface
    assert (ids != 0xFFFFFFFF).mean() > 0.2 and (ids[ids != 0xFFFFFFFF] >> 3).max() < sc["idx"].shape[0]
    if eye is not None:
        assert ((ids[ids != 0xFFFFFFFF] & 7) != 0).any()                # fan triangles of clipped polygons are visible
    # id strips rendered separately (what the ranks of a light shard do) assemble the same buffer
    for (y0, y1) in ((0, H // 3), (H // 3, H)):
        _, pg_s = util.params_pair("multi_hard", S, multi_fused=1, rect_x0=0, rect_y0=y0, rect_x1=W, rect_y1=y1)
        ctx.set_params(pg_s); ctx.render_prim_ids()
    assert np.array_equal(ctx.read("prim_id"), ids)
    ctx.set_params(pg_f); ctx.compute_visibility()
    assert util.bits_equal(ctx.read("visibility"), vis_a)


def test_fused_needs_the_id_pass(ctx):
    from globalillumination_b200 import capi
    sc = util.scene("door")
    po, pg, fm, mvp, mvpb = _many_light_frame(ctx, sc, 160, 120, 64, 4, multi_fused=1)
    ctx.render_shadow_map(); ctx.render_gbuffer()
    with pytest.raises(capi.SgiError):
        ctx.compute_visibility()                                         # G-buffer rendered, but no primitive ids


def test_comm_single_rank_strip_gather_and_reduce():
    """One rank: the strip is the screen, sgi_gather is a no-op, sgi_reduce_lights is the division by the light count."""
    from globalillumination_b200 import capi
    c = capi.Context(0)
    try:
        c.comm_init(capi.comm_unique_id(), 0, 1)
        sc = util.scene("teapot")
        W, H, S, n_l = 256, 145, 128, 6
        po, pg, fm, mvp, mvpb = _many_light_frame(c, sc, W, H, S, n_l)
        assert c.comm_strip(0) == (0, H)
        c.render_shadow_map(); c.render_gbuffer(); c.compute_visibility()
        full = c.read("visibility")
        _, pg_p = util.params_pair("multi_hard", S, multi_partial=1, multi_fused=1)
        c.set_params(pg_p)
        c.render_prim_ids(); c.gather("prim_id"); c.compute_visibility(); c.reduce_lights(n_l)
        assert util.bits_equal(c.read("visibility"), full)
        c.comm_destroy()
    finally:
        c.close()


@pytest.mark.parametrize("si", [0.25, 0.3])
def test_lit_masks_single_rank_and_two_half_sets(si):
    """params.multi_partial = 2: the many-light pass writes which lights reach each pixel (one bit per light of the whole set), the
    exchange sums the ranks' masks and the strip's visibility is accumulated in the reference's light order.  Checked here without
    a second GPU: (a) one rank, all lights: sgi_reduce_lights gives the un-sharded frame; (b) the lights dealt to two contexts
    (odd / even): the union of their masks, replayed on the host in light order, gives the un-sharded frame as well - also for a
    shadow intensity that is not a dyadic fraction, where partial float sums would depend on how the lights were dealt."""
    from globalillumination_b200 import capi
    sc = util.scene("teapot")
    W, H, S, n_l = 256, 145, 128, 11
    c = capi.Context(0)
    try:
        c.comm_init(capi.comm_unique_id(), 0, 1)
        po, pg, fm, mvp, mvpb = _many_light_frame(c, sc, W, H, S, n_l, shadow_intensity=si)
        c.render_shadow_map(); c.render_gbuffer(); c.compute_visibility()
        full = c.read("visibility")
        _, pg_m = util.params_pair("multi_hard", S, multi_partial=2, multi_fused=1, shadow_intensity=si)
        c.set_params(pg_m)
        c.set_light_ids(np.arange(n_l), n_l)
        c.render_prim_ids(); c.gather("prim_id"); c.compute_visibility(); c.reduce_lights(n_l)
        assert util.bits_equal(c.read("visibility"), full)
        mask_all = c.read_light_mask()
        fg = c.read("prim_id") != 0xFFFFFFFF
        assert (mask_all[~fg] == 0).all() and (mask_all >> n_l == 0).all()
        # (b) two half sets on two contexts of this device
        union = np.zeros((H, W), np.uint32)
        for part in (np.arange(0, n_l, 2), np.arange(1, n_l, 2)):
            d = capi.Context(0)
            try:
                d.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
                d.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
                d.set_multi_light_common(mvpb[-1])                       # the common term comes from the last light of the WHOLE set
                d.set_lights(mvp[part], mvpb[part], fm["light_pos_shading"], S, S)
                d.set_params(pg_m)
                d.set_light_ids(part, n_l)
                d.render_shadow_map(); d.render_prim_ids(); d.compute_visibility()
                m = d.read_light_mask()
                assert (m & union == 0).all()                            # disjoint bits: the byte sum of the exchange is their union
                union |= m
            finally:
                d.close()
        assert np.array_equal(union, mask_all)
        acc = np.zeros((H, W), np.float32); cnt = np.float32(0)
        for l in range(n_l):
            acc = (acc + np.where((union >> np.uint32(l)) & 1, np.float32(1.0), np.float32(si)).astype(np.float32)).astype(np.float32)
            cnt = np.float32(cnt + np.float32(1.0))
        replay = np.where(full != 0, acc / cnt, np.float32(0)).astype(np.float32)
        assert util.bits_equal(np.where(full != 0, full, np.float32(0)), replay)
        # a mask pass without light indices for the current lights is refused
        c.set_lights(mvp[:3], mvpb[:3], fm["light_pos_shading"], S, S)
        c.render_shadow_map()
        with pytest.raises(capi.SgiError):
            c.compute_visibility()
        c.comm_destroy()
    finally:
        c.close()


def test_two_contexts_on_one_device_and_kernel_attributes_per_context():
    """function attributes (dynamic shared memory opt-in) are per device; every context configures its own.  With one
    GPU this checks two contexts in one process side by side; with two GPUs the second context lives on device 1."""
    import torch
    from globalillumination_b200 import capi
    dev2 = 1 if torch.cuda.device_count() > 1 else 0
    a, b = capi.Context(0), capi.Context(dev2)
    try:
        sc = util.scene("teapot")
        W, H, S = 640, 360, 512
        out = []
        for c in (a, b):
            po, pg = util.params_pair("pcf", S)
            fm = util.frame(sc, W, H, S)
            c.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
            c.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
            c.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
            c.set_params(pg)
            c.render_shadow_map(); c.render_gbuffer(); c.compute_visibility()
            out.append((c.read("shadow_map"), c.read("visibility")))
        assert util.bits_equal(out[0][0], out[1][0]) and util.bits_equal(out[0][1], out[1][1])
        sm_o = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
        assert util.bits_equal(out[1][0][0], sm_o)
    finally:
        a.close(); b.close()


def test_two_ranks_over_nccl_match_the_unsharded_frame():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29677", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_config_c5_sandiego_full_size_against_the_oracle(ctx):
    """c5 at full size (SanDiego geometry, 7680x4320, 16 lights x 8192^2) against the oracle itself, not only
    through size-independent properties: every depth map, the vertex map and the 16-light visibility of the whole frame, bit for
    bit, through the G-buffer path and through the fused primitive-id path (the one bench.py's `sharded` record runs)."""
    sc = util.scene("sandiego")
    W, H, S, n_l = 7680, 4320, 8192, 16
    po, pg, fm, mvp, mvpb = _many_light_frame(ctx, sc, W, H, S, n_l)
    ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    vis_a = ctx.read("visibility")
    maps = ctx.read("shadow_map")
    for i in range(n_l):
        m_o = O.raster_depth(sc["xyz"], sc["idx"], mvp[i], S, S)
        assert util.bits_equal(maps[i], m_o), f"light {i}: " + util.describe_diff(maps[i], m_o)
    pos_o, nrm_o, dep_o = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    assert util.bits_equal(ctx.read("gbuf_pos"), pos_o)
    vis_o = O.visibility_multi(po, mvpb[-1], mvpb[:, 12:16], pos_o, maps)
    assert util.bits_equal(vis_a, vis_o), util.describe_diff(vis_a, vis_o)
    del maps
    _, pg_f = util.params_pair("multi_hard", S, multi_fused=1)
    ctx.set_params(pg_f)
    ctx.render_prim_ids(); ctx.compute_visibility()
    vis_b = ctx.read("visibility")
    assert util.bits_equal(vis_b, vis_o), util.describe_diff(vis_b, vis_o)
    fg = pos_o[..., 0] != 0
    assert fg.mean() > 0.3 and len(np.unique(vis_o[fg])) > 4
