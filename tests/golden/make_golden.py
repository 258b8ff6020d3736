#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own sources (runs only where /root/reference exists).

Everything here comes out of oracle/_ref — the reference's unmodified GLSL fragment shaders and its
Mesh/SceneLoader/OBJLoader/ShadowVolume/UniformSampledLightSource/GLM sources compiled on the CPU by
oracle/ref_build/build_ref.py.  The .npz files are the committed golden vectors; this script is how they
were made:   python tests/golden/make_golden.py

  scene_<name>.npz        the reference loader's output for Configs/<Name>.txt: xyz, nrm, idx + views
  golden_host.npz         GLM frame matrices per config, ShadowVolume prisms, uniform light samples
  golden_shaders.npz      one small frame (inputs + uniforms) and gl_FragData[0].r of every technique
  golden_moments.npz      VSM / ESM / EVSM / MSM: the light-view moment programs on seeded fragments, both passes of
                          filterShadowMap and Shadow.frag's reconstruction on a small frame, GLM's quantisation matrices
                          (python tests/golden/make_golden.py --only-moments regenerates this file alone)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCENE_DIR = os.path.join(os.path.dirname(os.path.dirname(HERE)), "globalillumination_b200", "data")   # input scenes ship with the package
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))
from oracle import oracle_py as O  # noqa: E402

SCENES = {"teapot": "Configs/Teapot.txt", "door": "Configs/Door.txt", "dragon": "Configs/Dragon.txt",
          "raptor": "Configs/Raptor.txt"}
# configs whose asset list is only partly present in the mount: the present objects of the config, same directives
PARTIAL = {
    # Configs/TreeWithLeaves.txt without the missing TreeSub1.obj (config c4)
    "tree": ["o OBJ/TreeWithLeaves/TreeSub0.obj", "s 2.0 2.0 2.0", "t 0.0 -7.0 9.0", "+",
             "o OBJ/TreeWithLeaves/plane.obj", "m OBJ/TreeWithLeaves/white.png", "s 60.0 1.0 40.0", "t 0.0 -8.0 0.0", "+",
             "ve 0.0 41.0 -57.0", "va 0.0 16.0 -17.0", "le 10.0 130.0 100.0", "la 0.0 0.0 0.0", "d 0.000025"],
    # Configs/SanDiego.txt without the two missing sphere.obj objects (config c5)
    "sandiego": ["o OBJ/SanDiego/building.obj", "s 30.0 30.0 30.0", "t -10.0 -8.0 0.0", "r 0.0 90.0 0.0", "c 1.0 1.0 0.5", "+",
                 "o OBJ/SanDiego/plane.obj", "m OBJ/SanDiego/white.png", "s 60.0 1.0 40.0", "t 0.0 -8.0 0.0", "+",
                 "ve 0.0 41.0 -50.0", "va 0.0 16.0 -10.0", "le 10.0 130.0 100.0", "la 0.0 0.0 0.0", "d 0.000025"],
}

SHADER_OF = {
    "hard": ("shadow", dict(naive=1, bilinearPCF=1)),
    "pcf": ("shadow", dict(naive=0, bilinearPCF=1)),
    "pcss": ("plausible", dict(PCSS=1)),
    "rbsm_noncons": ("nonconservative", dict(SMSR=1)),
    "rbsm_cons": ("conservative", dict(SMSR=1)),
    "rpcf_noncons": ("nonconservative", dict(SMSR=0, RPCFPlusSMSR=1)),
    "rpcf_cons": ("conservative", dict(SMSR=0, RPCFPlusSMSR=1)),
    "rsmss": ("filtered", dict(SMSR=0, RPCFPlusSMSR=1)),
    "rbssm": ("rbssm", dict()),
}


def shader_uniforms(fm, pos, nrm, sm, S, p):
    return dict(
        shadowMap=("tex", sm), vertexMap=("tex", pos), normalMap=("tex", nrm), MV=fm["cam_mv"], lightMVP=fm["light_mvp_b"],
        normalMatrix=fm["normal_matrix"], lightPosition=fm["light_pos_shading"], shadowIntensity=np.float32(p.shadow_intensity),
        shadowMapWidth=np.int32(S), shadowMapHeight=np.int32(S), zNear=np.int32(p.z_near), zFar=np.int32(p.z_far),
        kernelOrder=np.int32(p.kernel_order), penumbraSize=np.int32(p.penumbra_size),
        shadowMapStep=np.array([np.float32(1.0 / S), np.float32(1.0 / S)], np.float32),
        depthThreshold=np.float32(p.depth_threshold), maxSearch=np.int32(p.max_search),
        blockerSearchSize=np.int32(p.blocker_search_size), kernelSize=np.int32(p.kernel_size),
        lightSourceRadius=np.int32(p.light_source_radius))


def ref_visibility(tech, fm, pos, nrm, sm, S, p, W, H):
    shader, extra = SHADER_OF[tech]
    u = shader_uniforms(fm, pos, nrm, sm, S, p)
    if shader == "shadow":      # uniforms keep their last value between runs of one program: clear the branches other runs may have set
        u.update({k: np.int32(0) for k in ("VSM", "ESM", "EVSM", "MSM", "tricubicPCF")})
    u.update({k: np.int32(v) for k, v in extra.items()})
    return O.ref_run_shader(shader, u, W, H)[..., 0].copy()


def multi_light_setup(sc, n_lights, size, W, H, S):
    """renderMonteCarlo's light set: SoftShadowMapping/src/main.cpp:756-811 (+UniformSampledLightSource)."""
    mvps, mvpbs = [], []
    for i in range(n_lights):
        e = O.ref_uniform_sample(sc["light_eye"], size, n_lights, i)
        a = O.ref_uniform_sample(sc["light_at"], size, n_lights, i)
        fm = O.ref_frame_matrices(sc["cam_eye"], sc["cam_at"], e, a, W, H, S, S)
        mvps.append(fm["light_mvp"]); mvpbs.append(fm["light_mvp_b"])
    return np.stack(mvps), np.stack(mvpbs)


# ---- moment shadow maps (VSM / ESM / EVSM / MSM): golden_moments.npz -------------------------------------------------
MOMENT_TYPED = [-2.07224649, 32.2370378, -68.5710746, 39.3703274, 13.7948857, -59.4683976, 82.035975, -35.3649032,
                0.105877704, -1.90774663, 9.34965551, -6.65434907, 9.79240621, -33.76521106, 47.9456097, -23.9728048]
MOMENT_SHADER = {"vsm": ("moments", dict(VSM=1, MSM=0)), "msm": ("moments", dict(VSM=0, MSM=1)),
                 "esm": ("exponential", {}), "evsm": ("expmoments", {})}


def lin32(d, n=1, f=1000):
    n, f = np.float32(n), np.float32(f)
    return (np.float32(2.0) * n) / (f + n - d * (f - n))


def moment_texel_inputs(seed=5, W=32, H=24):
    """Random per-fragment inputs of the light-view moment programs: window depth of the fragment and of its two quad
    partners, the `position` varying that yields it and the dFdx / dFdy the fine quad differences give."""
    rng = np.random.default_rng(seed)
    ndc = rng.uniform(0.2, 0.9999, (H, W)).astype(np.float32)
    ndc[::5, ::3] = rng.uniform(-1.0, 0.2, ndc[::5, ::3].shape).astype(np.float32)
    zwin = ndc * np.float32(0.5) + np.float32(0.5)
    zpx = (zwin + rng.uniform(-2e-4, 2e-4, (H, W)).astype(np.float32)).astype(np.float32)
    zpy = (zwin + rng.uniform(-2e-4, 2e-4, (H, W)).astype(np.float32)).astype(np.float32)
    xo = (np.arange(W)[None, :] & 1).astype(bool) & np.ones((H, W), bool)
    yo = (np.arange(H)[:, None] & 1).astype(bool) & np.ones((H, W), bool)
    d, dpx, dpy = lin32(zwin), lin32(zpx), lin32(zpy)
    ddx = np.where(xo, d - dpx, dpx - d).astype(np.float32)
    ddy = np.where(yo, d - dpy, dpy - d).astype(np.float32)
    pos = np.zeros((H, W, 4), np.float32)
    pos[..., 2] = ndc; pos[..., 3] = 1.0
    return dict(zwin=zwin, zpx=zpx, zpy=zpy, ddx=ddx, ddy=ddy, position=pos)


def ref_moment_texels(tech, ti, m_q, t_q):
    shader, flags = MOMENT_SHADER[tech]
    H, W = ti["zwin"].shape
    u = {"varying:position": ti["position"], "dFdx": ti["ddx"], "dFdy": ti["ddy"], "zNear": np.int32(1), "zFar": np.int32(1000),
         "mQuantization": m_q, "tQuantization": t_q}
    u.update({k: np.int32(v) for k, v in flags.items()})
    return O.ref_run_shader(shader, u, W, H)


def ref_filter_pass(src4, W, H, order, horizontal, tech):
    k = np.zeros(33, np.float32)
    k[:order] = O.gaussian_kernel(order)
    u = dict(image=("tex", src4, "linear"), width=np.int32(W), height=np.int32(H), order=np.int32(order),
             horizontal=np.int32(horizontal), vertical=np.int32(1 - horizontal), kernel=k)
    return O.ref_run_shader("loggaussian" if tech == "esm" else "gaussian", u, W, H)


def ref_visibility_moments(tech, fm, pos, nrm, fmap, p, W, H, m_qi, t_q):
    u = shader_uniforms(fm, pos, nrm, np.zeros((2, 2), np.float32), p.shadow_map_width, p)
    u["shadowMap"] = ("tex", fmap, "linear")
    u.update(dict(naive=np.int32(0), bilinearPCF=np.int32(0), VSM=np.int32(tech == "vsm"), ESM=np.int32(tech == "esm"),
                  EVSM=np.int32(tech == "evsm"), MSM=np.int32(tech == "msm"), mQuantizationInverse=m_qi, tQuantization=t_q))
    return O.ref_run_shader("shadow", u, W, H)[..., 0].copy()


def make_moments(scenes):
    out = {}
    m_q, m_qi = O.ref_moment_quantization(MOMENT_TYPED)       # the reference's GLM: transpose + inverse
    t_q = np.array([0.0359558848, 0, 0, 0], np.float32)
    out["quant/m"], out["quant/minv"], out["quant/t"] = m_q, m_qi, t_q
    ti = moment_texel_inputs()
    for k, v in ti.items():
        out[f"texel/{k}"] = v
    W, H, S, order = 128, 72, 96, 7
    sc = scenes["teapot"]
    fm = O.ref_frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, S, S)
    pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    out.update(W=W, H=H, S=S, order=order, **{f"fm_{k}": v for k, v in fm.items()})
    for tech in O.MOMENT_TECHS:
        out[f"texel/{tech}"] = ref_moment_texels(tech, ti, m_q, t_q)
        # the chain's input is the moment target the oracle's own rasteriser produces (the GL rasteriser has no source)
        mom = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, tech)
        fx = ref_filter_pass(mom, W, H, order, 1, tech)
        fy = ref_filter_pass(fx, W, H, order, 0, tech)
        out[f"chain/{tech}/filter_x"], out[f"chain/{tech}/filter_y"] = fx, fy
        for variant, si in (("default", 0.25), ("alt", 0.5)):
            p = O.default_params(tech, S, shadow_intensity=si)
            out[f"chain/{tech}/vis/{variant}"] = ref_visibility_moments(tech, fm, pos, nrm, fy, p, W, H, m_qi, t_q)
    # tricubic PCF (Shadow.frag:41-84,101 with tricubicPCF == 1): same small frame with the plain depth map
    sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    for variant, kw in (("default", {}), ("alt", dict(kernel_order=5, penumbra_size=2, shadow_intensity=0.5))):
        p = O.default_params("pcf_tricubic", S, **kw)
        u = shader_uniforms(fm, pos, nrm, sm, S, p)
        u.update(dict(naive=np.int32(0), bilinearPCF=np.int32(0), tricubicPCF=np.int32(1), VSM=np.int32(0), ESM=np.int32(0),
                      EVSM=np.int32(0), MSM=np.int32(0)))          # uniforms keep their last value between runs of one program
        out[f"tricubic/vis/{variant}"] = O.ref_run_shader("shadow", u, W, H)[..., 0].copy()
    np.savez_compressed(os.path.join(HERE, "golden_moments.npz"), **out)
    print("golden_moments.npz", os.path.getsize(os.path.join(HERE, "golden_moments.npz")) // 1024, "KiB")


def main():
    if "--only-moments" in sys.argv:
        assert O.build_ref(), "oracle/_ref could not be built (is /root/reference mounted?)"
        return make_moments({"teapot": O.ref_load_scene(SCENES["teapot"])})
    assert O.build_ref(), "oracle/_ref could not be built (is /root/reference mounted?)"
    scenes = {}
    for name, cfg in SCENES.items():
        sc = O.ref_load_scene(cfg)
        scenes[name] = sc
        np.savez_compressed(os.path.join(SCENE_DIR, f"scene_{name}.npz"), **sc)
        print(name, sc["xyz"].shape, sc["idx"].shape)
    import tempfile
    for name, lines in PARTIAL.items():
        cfg = os.path.join(tempfile.mkdtemp(), name + ".txt")
        with open(cfg, "w") as f:
            f.write("\n".join(lines))
        sc = O.ref_load_scene(cfg)
        np.savez_compressed(os.path.join(SCENE_DIR, f"scene_{name}.npz"), **sc)
        print(name, sc["xyz"].shape, sc["idx"].shape)

    host = {}
    for name, sc in scenes.items():
        for (W, H, S) in ((1280, 720, 1024), (1920, 1080, 2048), (640, 480, 512)):
            fm = O.ref_frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, S, S)
            for k, v in fm.items():
                host[f"fm/{name}/{W}x{H}x{S}/{k}"] = v
    d = scenes["door"]
    pxyz, pidx = O.ref_sv_prisms(d["xyz"], d["nrm"], d["idx"], d["light_eye"], 100)
    host["sv/door/xyz"], host["sv/door/idx"] = pxyz, pidx
    for n, size in ((16, 16), (289, 16), (4, 8)):
        host[f"uls/{n}/{size}"] = np.stack([O.ref_uniform_sample(np.array([10, 130, 100], np.float32), size, n, i) for i in range(n)])
    host["rotate/33.5/y"] = np.zeros(16, np.float32)
    O.ref_host().ref_rotate(O.C.c_float(33.5), O._fp(np.array([0, 1, 0], np.float32)), O._fp(host["rotate/33.5/y"]))
    np.savez_compressed(os.path.join(HERE, "golden_host.npz"), **host)

    # one small frame through every technique of the reference's shaders
    W, H, S = 160, 90, 128
    sc = scenes["teapot"]
    fm = O.ref_frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"], sc["light_at"], W, H, S, S)
    sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
    pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
    out = dict(W=W, H=H, S=S, pos=pos, nrm=nrm, sm=sm, depth_threshold=sc["depth_threshold"], **{f"fm_{k}": v for k, v in fm.items()})
    for tech in SHADER_OF:
        for variant, kw in (("default", {}), ("alt", dict(kernel_order=9, kernel_size=7, shadow_intensity=0.5, max_search=8))):
            p = O.default_params(tech, S, depth_threshold=float(sc["depth_threshold"]), **kw)
            out[f"vis/{tech}/{variant}"] = ref_visibility(tech, fm, pos, nrm, sm, S, p, W, H)
    n_l = 4
    mvps, mvpbs = multi_light_setup(sc, n_l, 8, W, H, S)
    maps = np.stack([O.raster_depth(sc["xyz"], sc["idx"], mvps[i], S, S) for i in range(n_l)])
    u = dict(shadowMapArray=("tex", maps, "array"), vertexMap=("tex", pos), normalMap=("tex", nrm), lightMVP=mvpbs[-1],
             lightMVPTrans=np.ascontiguousarray(mvpbs[:, 12:16]), shadowIntensity=np.float32(0.25), numberOfSamples=np.int32(n_l),
             shadowMapWidth=np.int32(S), shadowMapHeight=np.int32(S), monteCarlo=np.int32(1), adaptiveSampling=np.int32(0),
             adaptiveSamplingLowerAccuracy=np.int32(0))
    out["multi/maps"], out["multi/mvp"], out["multi/mvpb"] = maps, mvps, mvpbs
    out["multi/vis"] = O.ref_run_shader("accurate", u, W, H)[..., 0].copy()
    # deferred shading (PhongShading.frag) of the PCF frame with seeded per-vertex colours
    rgb = np.random.default_rng(11).uniform(0.1, 1.0, (len(sc["xyz"]), 3)).astype(np.float32)
    _, _, alb, _ = O.raster_gbuffer_rgb(sc["xyz"], sc["nrm"], rgb, sc["idx"], fm["cam_mvp"], W, H)
    vis_pcf = out["vis/pcf/default"]
    hs = np.ascontiguousarray(np.stack([vis_pcf, np.zeros_like(vis_pcf), np.zeros_like(vis_pcf), np.ones_like(vis_pcf)], -1))
    u = dict(hardShadowMap=("tex", hs), colorMap=("tex", alb), vertexMap=("tex", pos), normalMap=("tex", nrm), MV=fm["cam_mv"],
             normalMatrix=fm["normal_matrix"], lightPosition=fm["light_pos_shading"], shadowIntensity=np.float32(0.25))
    out["phong/rgb"], out["phong/albedo"] = rgb, alb
    out["phong/image"] = O.ref_run_shader("phong", u, W, H)
    # EDT shadow mapping: the RBSM shaders' EDTSM target and both MeanFilter.frag passes (the CUDA stages between them -
    # site detection, Voronoi diagram, normalisation - have no compilable reference: oracle restatement, see its header)
    for tech, shader in (("edtsm_noncons", "nonconservative"), ("edtsm_cons", "conservative")):
        p = O.default_params(tech, S, depth_threshold=float(sc["depth_threshold"]), penumbra_size=5)
        u = shader_uniforms(fm, pos, nrm, sm, S, p)
        u.update(dict(SMSR=np.int32(0), RPCFPlusSMSR=np.int32(0), EDTSM=np.int32(1), MVP=fm["cam_mvp"]))
        img = O.ref_run_shader(shader, u, W, H)
        out[f"edt/{tech}/hard"] = img
        near = O.edt_nearest(O.edt_sites(img))
        a2 = O.edt_normalize(img, pos, near, np.float32(p.penumbra_size / 5.0), p.shadow_intensity)
        stage = a2
        for axis, horizontal, mode in (("x", 1, ()), ("y", 0, ("linear",))):
            rgba = np.zeros((H, W, 4), np.float32)
            rgba[..., :2] = stage
            fu = dict(image=("tex", rgba) + mode, vertexMap=("tex", pos), MV=fm["cam_mv"], shadowIntensity=np.float32(p.shadow_intensity),
                      fov=np.float32(45.0), width=np.int32(W), height=np.int32(H), order=np.int32(p.kernel_order),
                      horizontal=np.int32(horizontal), vertical=np.int32(1 - horizontal), zNear=np.int32(1), zFar=np.int32(1000))
            stage = np.ascontiguousarray(O.ref_run_shader("meanfilter", fu, W, H)[..., :2])
            out[f"edt/{tech}/filter_{axis}_in"] = rgba[..., :2].copy()
            out[f"edt/{tech}/filter_{axis}"] = stage
    np.savez_compressed(os.path.join(HERE, "golden_shaders.npz"), **out)
    make_moments(scenes)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
