"""Host side (C++: SceneLoader / Mesh / OBJ reader / matrices / procedural stand-ins / technique selection) —
against the reference loader's golden output, GLM goldens and hand-checked small cases.  No GPU needed."""
import hashlib
import os

import numpy as np
import pytest

from globalillumination_b200 import hostapi
from oracle import oracle_py as O
from tests import util

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
@pytest.mark.parametrize("name,cfg", [("teapot", "Configs/Teapot.txt"), ("door", "Configs/Door.txt"), ("dragon", "Configs/Dragon.txt"),
                                      ("raptor", "Configs/Raptor.txt")])
def test_scene_loader_bit_identical_to_reference_loader(name, cfg):
    """Our SceneLoader+Mesh+OBJ reader on the reference's own config/asset files == golden arrays produced by the
    reference's SceneLoader.cpp/Mesh.cpp/OBJLoader.cpp (tests/golden/make_golden.py)."""
    sc, g = hostapi.load_scene(os.path.join(REF, cfg), REF), util.scene(name)
    for k in ("xyz", "nrm", "idx", "cam_eye", "cam_at", "light_eye", "light_at", "depth_threshold"):
        assert util.bits_equal(sc[k], g[k]), (name, k)
    assert sc["substitutions"] == []


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
def test_missing_assets_are_substituted_and_reported():
    sc = hostapi.load_scene(os.path.join(REF, "Configs/Sponza.txt"), REF)
    assert sc["substitutions"] == ["OBJ/Sponza/sponza.obj"] and sc["idx"].shape[0] > 60000
    assert np.array_equal(sc["cam_eye"], [0, 19, -52]) and np.array_equal(sc["light_at"], [0, -17, -17])
    with pytest.raises(hostapi.HostError):
        hostapi.load_scene(os.path.join(REF, "Configs/Armadillo.txt"), REF)      # no stand-in registered: error, not exit()


@pytest.mark.parametrize("name", ["teapot", "dragon"])
def test_frame_matrices_bit_identical_to_glm(name):
    g, sc = util.golden("golden_host.npz"), util.scene(name)
    for (W, H, S) in ((1280, 720, 1024), (1920, 1080, 2048), (640, 480, 512)):
        fm = hostapi.frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, S, S)
        for k, v in fm.items():
            assert util.bits_equal(v, g[f"fm/{name}/{W}x{H}x{S}/{k}"]), (W, H, S, k)
        fo = util.frame(sc, W, H, S)
        assert all(util.bits_equal(fm[k], fo[k]) for k in fm)              # host == oracle as well


def test_obj_reader_tokeniser_rules():
    cwd = os.getcwd()
    os.chdir(os.path.dirname(DATA))
    try:
        cfg = os.path.join(DATA, "_one.txt")
        with open(cfg, "w") as f:
            f.write("o data/quadfan.obj\n+\nve 0 0 0\nva 0 0 1\nle 0 1 0\nla 0 0 0")
        sc = hostapi.load_scene(cfg, os.path.dirname(DATA))
    finally:
        os.chdir(cwd)
    assert sc["xyz"].shape == (6, 3)
    assert np.array_equal(sc["xyz"][0], [0, 0, 0]) and np.array_equal(sc["xyz"][4], [0.5, 1.5, 0.25]) and np.array_equal(sc["xyz"][5], [2, 0, 0])
    # quad -> fan (0,1,2),(0,2,3); v//n; negative indices (-5,-4,-3 of 5 vertices = 1,2,3); v/t; plain
    assert sc["idx"].tolist() == [[0, 1, 2], [0, 2, 3], [3, 2, 4], [0, 1, 2], [0, 2, 4], [1, 5, 2]]
    # computeNormals always overrides file normals: running mean of face normals, all faces here have +-z normals
    assert np.allclose(sc["nrm"][5], [0, 0, 1], atol=1e-6)        # vertex 5 only belongs to the in-plane face (1,5,2)


def test_scene_grammar_transform_order_and_keys():
    sc = hostapi.load_scene(os.path.join(DATA, "two_objects.txt"), os.path.dirname(DATA))
    assert sc["xyz"].shape == (10, 3) and sc["idx"].shape == (8, 3)
    assert sc["idx"][6:].min() == 6                                        # second object's indices are re-based
    # first object: scale 2, rotate y 90 (Rx*Ry*Rz transposed, row-vector product), translate (1,-1,0.5)
    v = np.array([1.0, 0.0, 0.0], np.float32) * 2
    R = O.rotate(90.0, [0, 1, 0]).reshape(4, 4).T[:3, :3]                   # column-major -> row-major 3x3
    expect = (R @ v).astype(np.float32) + np.array([1, -1, 0.5], np.float32)   # transposing twice: a plain R*p
    assert np.allclose(sc["xyz"][1], expect, atol=1e-6)
    assert np.allclose(sc["xyz"][6:, 1], -1.0)                              # plane: y=1 * 1 + (-2)
    assert sc["depth_threshold"] == np.float32(0.0000025)
    assert np.array_equal(sc["light_eye"], [10, 130, 100])


def test_procedural_scenes_are_deterministic():
    a, ai = hostapi.procedural("sponza_like?seed=1")
    b, bi = hostapi.procedural("sponza_like?seed=1")
    c, _ = hostapi.procedural("sponza_like?seed=2")
    assert util.bits_equal(a, b) and np.array_equal(ai, bi) and not np.array_equal(a, c)
    assert ai.shape[0] == 69424 and ai.max() == a.shape[0] - 1 and ai.min() == 0
    assert hostapi.procedural("sphere?seed=2")[1].shape[0] == 16128
    with pytest.raises(hostapi.HostError):
        hostapi.procedural("no_such_scene")


def test_workload_configs_load_through_the_scene_loader():
    from globalillumination_b200 import scenes
    sc = hostapi.load_scene(scenes.write_config("c2_sponza"))
    assert sc["idx"].shape[0] == 69424 and np.array_equal(sc["light_eye"], [0, 43, -50])
    # scale 2, rotate y 90, translate z -30 of a 35 x 15 x 15 model
    assert np.allclose(sc["xyz"].min(0), [-15, 0, -65], atol=1e-3) and np.allclose(sc["xyz"].max(0), [15, 30, 5], atol=1e-3)
    sc5 = hostapi.load_scene(scenes.write_config("c5_many_light"))
    assert sc5["idx"].shape[0] > 100000


def test_png_writer_round_trip(tmp_path):
    """The host library's PNG encoder (stored deflate blocks): decoded by an independent reader the pixels are the input;
    images larger than one 64 KiB deflate block included."""
    from PIL import Image
    rng = np.random.default_rng(3)
    for (H, W) in ((1, 1), (7, 5), (90, 160), (300, 257)):
        img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        path = tmp_path / f"t_{W}x{H}.png"
        hostapi.write_png(path, img)
        back = np.asarray(Image.open(path).convert("RGBA"))
        assert back.shape == img.shape and np.array_equal(back, img)


@pytest.mark.gpu
def test_set_scene_rejects_out_of_range_indices():
    """ShadowApp::setScene validates the indices before Mesh::computeNormals walks the vertex arrays with them."""
    from globalillumination_b200 import hostapi
    app = hostapi.App(0)
    try:
        sc = dict(util.scene("door"))
        bad = sc["idx"].copy(); bad[3, 1] = sc["xyz"].shape[0] + 7
        sc["idx"] = bad
        with pytest.raises(hostapi.HostError):
            app.set_scene(sc)
        sc["idx"][3, 1] = -1
        with pytest.raises(hostapi.HostError):
            app.set_scene(sc)
    finally:
        app.close()
