"""Texture albedo: the texture select of GBuffer.frag:11-30 (useTextureForColoring) on the (u, v, texture id) varying, with the scene
textures of loadRGBTexture (RGB8, GL_LINEAR, GL_REPEAT).  CPU: the oracle's fragment colour against the unmodified GBuffer.frag
(golden fixture tests/golden/golden_gbuffer.npz, made by make_golden_gbuffer below with the reference build; live when oracle/_ref is
present).  GPU: the textured G-buffer against the oracle bit for bit, the shaded frame within the powf tolerance (2e-6 relative)."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_gbuffer.npz")


def _inputs(seed=5, n=64 * 48):
    rng = np.random.default_rng(seed)
    tex = [rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), rng.integers(0, 256, (16, 16, 3), dtype=np.uint8), None]
    uvw = np.zeros((n, 3), np.float32)
    uvw[:, 0] = rng.uniform(-2.5, 3.5, n); uvw[:, 1] = rng.uniform(-2.5, 3.5, n)
    uvw[:, 2] = rng.choice(np.float32([0, 1, 2, 3, 0.995, 1.0005, 1.9995, 2.5, 1.002, 0.98, 0.99, 1.001]), n)
    rgb = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    return tex, uvw, rgb


def _run_reference(tex, uvw, rgb, W=64, H=48):
    def img(v3):
        o = np.zeros((H, W, 4), np.float32); o[..., :3] = v3.reshape(H, W, 3); return o

    def texf(t):
        if t is None:
            return np.zeros((1, 1, 4), np.float32)
        o = np.ones(t.shape[:2] + (4,), np.float32); o[..., :3] = t.astype(np.float32) / np.float32(255.0); return o
    u = {"useTextureForColoring": np.int32(1), "useMeshColor": np.int32(1), "varying:GBufferTextureCoordinates": img(uvw),
         "varying:GBufferColor": img(rgb), "texture0": ("tex", texf(tex[0]), "linear_repeat"),
         "texture1": ("tex", texf(tex[1]), "linear_repeat"), "texture2": ("tex", texf(tex[2]), "linear_repeat")}
    return O.ref_run_shader("gbuffer", u, W, H).reshape(-1, 4)


def make_golden_gbuffer():
    """python -c 'from tests.test_texture_albedo import make_golden_gbuffer as m; m()'  (needs /root/reference + oracle/_ref)"""
    tex, uvw, rgb = _inputs()
    np.savez_compressed(GOLDEN, out=_run_reference(tex, uvw, rgb))


def test_fragment_color_matches_the_reference_shader_golden():
    tex, uvw, rgb = _inputs()
    out = O.fragment_color(uvw, rgb, tex)
    gold = np.load(GOLDEN)["out"]
    assert util.bits_equal(out, gold), util.describe_diff(out, gold)
    # all three kinds of fragments are present: texture 0, texture 1, vertex colour (and the unbound texture 2)
    assert (out[:, 3] == 1.0).any() and (out[:, 3] == 0.0).any()


@pytest.mark.skipif(not O.ref_available(), reason="needs the reference build (oracle/_ref)")
def test_fragment_color_matches_the_reference_shader_live():
    tex, uvw, rgb = _inputs(seed=11)
    assert util.bits_equal(O.fragment_color(uvw, rgb, tex), _run_reference(tex, uvw, rgb))


def _textured_scene():
    sc = dict(util.scene("teapot"))
    xyz = sc["xyz"]
    uv = np.zeros_like(xyz)
    uv[:, 0] = xyz[:, 0] * np.float32(0.07); uv[:, 1] = xyz[:, 2] * np.float32(0.05)
    uv[:, 2] = np.where(xyz[:, 1] > -7.9, 1.0, 2.0).astype(np.float32)        # teapot: texture 1, floor: texture 2 (`m` numbering)
    uv[::97, 2] = 0.0                                                          # a few vertices of untextured objects
    rng = np.random.default_rng(3)
    tex = [rng.integers(0, 256, (64, 48, 3), dtype=np.uint8), rng.integers(0, 256, (33, 17, 3), dtype=np.uint8), None]
    rgb = rng.uniform(0, 1, xyz.shape).astype(np.float32)
    return sc, uv, tex, rgb


def test_oracle_textured_gbuffer_reduces_to_the_colour_form_without_textures():
    sc, uv, tex, rgb = _textured_scene()
    W, H = 160, 90
    fm = util.frame(sc, W, H, 64)
    a = O.raster_gbuffer_rgb(sc["xyz"], sc["nrm"], rgb, sc["idx"], fm["cam_mvp"], W, H)
    uv0 = uv.copy(); uv0[:, 2] = 0.0
    b = O.raster_gbuffer_tex(sc["xyz"], sc["nrm"], rgb, uv0, sc["idx"], fm["cam_mvp"], W, H, tex)
    for x, y in zip(a, b):
        assert util.bits_equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("with_rgb", [True, False])
def test_textured_gbuffer_bit_exact_and_shading(with_rgb):
    from globalillumination_b200 import capi
    sc, uv, tex, rgb = _textured_scene()
    W, H, S = 640, 360, 256
    ctx = capi.Context(0)
    try:
        po, pg = util.params_pair("hard", S)
        fm = util.frame(sc, W, H, S)
        ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
        ctx.set_mesh_colors(rgb if with_rgb else None)
        ctx.set_mesh_uv(uv)
        for k in range(3):
            ctx.set_texture(k, tex[k])
        ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
        ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
        ctx.set_params(pg)
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
        pos_o, nrm_o, alb_o, dep_o = O.raster_gbuffer_tex(sc["xyz"], sc["nrm"], rgb if with_rgb else None, uv, sc["idx"], fm["cam_mvp"], W, H, tex)
        assert util.bits_equal(ctx.read("gbuf_pos"), pos_o) and util.bits_equal(ctx.read("gbuf_nrm"), nrm_o)
        alb = ctx.read("gbuf_albedo")
        assert util.bits_equal(alb, alb_o), util.describe_diff(alb, alb_o)
        ctx.shade_phong()
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        shaded_o = O.shade_phong(cam, po.shadow_intensity, pos_o, nrm_o, alb_o, ctx.read("visibility"))
        # powf is the one operation whose CPU and GPU implementations are not both correctly rounded (as in test_gpu_parity.py)
        assert np.allclose(ctx.read("shaded"), shaded_o, rtol=2e-6, atol=1e-7)
        fg = dep_o < 1
        assert len(np.unique(alb_o[fg][:, 0])) > 500                          # filtered texels, not a handful of flat colours
        # unbinding the textures returns to the vertex-colour form
        for k in range(3):
            ctx.set_texture(k, None)
        ctx.render_gbuffer()
        if with_rgb:
            _, _, alb_c, _ = O.raster_gbuffer_rgb(sc["xyz"], sc["nrm"], rgb, sc["idx"], fm["cam_mvp"], W, H)
            assert util.bits_equal(ctx.read("gbuf_albedo"), alb_c)
    finally:
        ctx.close()
