"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built, for the reference's own
sources compiled on the CPU (oracle/_ref/libref_shaders.so, libref_host.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from globalillumination_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

TECH = {
    "hard": 0, "pcf": 1, "pcss": 2, "rbsm_noncons": 3, "rbsm_cons": 4,
    "rpcf_noncons": 5, "rpcf_cons": 6, "rsmss": 7, "multi_hard": 8, "rbssm": 9, "edtsm_noncons": 10, "edtsm_cons": 11,
    "vsm": 12, "esm": 13, "evsm": 14, "msm": 15, "pcf_tricubic": 16,
}
MOMENT_TECHS = ("vsm", "esm", "evsm", "msm")
DEPTH_LESS, DEPTH_LEQUAL = 0, 1


class Params(C.Structure):
    """orc_params (same layout as sgi_params)."""
    _fields_ = [
        ("technique", C.c_int32),
        ("shadow_map_width", C.c_int32), ("shadow_map_height", C.c_int32),
        ("shadow_intensity", C.c_float),
        ("kernel_order", C.c_int32), ("penumbra_size", C.c_int32),
        ("blocker_search_size", C.c_int32), ("kernel_size", C.c_int32), ("light_source_radius", C.c_int32),
        ("max_search", C.c_int32), ("depth_threshold", C.c_float),
        ("z_near", C.c_int32), ("z_far", C.c_int32),
        ("polygon_offset_factor", C.c_float), ("polygon_offset_units", C.c_float),
        ("sv_depth_func", C.c_int32), ("sv_infinity", C.c_int32),
        ("rect_x0", C.c_int32), ("rect_y0", C.c_int32), ("rect_x1", C.c_int32), ("rect_y1", C.c_int32),
        ("multi_partial", C.c_int32), ("multi_fused", C.c_int32), ("sv_silhouette", C.c_int32), ("sv_zfail", C.c_int32),
    ]


def default_params(technique="hard", S=1024, **kw):
    """Reference defaults: ShadowMapping/src/main.cpp:859-877, SoftShadowMapping/src/main.cpp:1598-1614."""
    p = Params(
        technique=TECH[technique] if isinstance(technique, str) else technique,
        shadow_map_width=S, shadow_map_height=S, shadow_intensity=0.25,
        kernel_order=7, penumbra_size=1, blocker_search_size=7, kernel_size=15, light_source_radius=8,
        max_search=16, depth_threshold=0.0, z_near=1, z_far=1000,
        polygon_offset_factor=4.0, polygon_offset_units=20.0,
        sv_depth_func=DEPTH_LEQUAL, sv_infinity=100, rect_x0=0, rect_y0=0, rect_x1=0, rect_y1=0, multi_partial=0,
    )
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Camera(C.Structure):
    _fields_ = [("mv", C.c_float * 16), ("normal_matrix", C.c_float * 9), ("light_pos", C.c_float * 3)]


def make_camera(mv, normal_matrix, light_pos):
    c = Camera()
    c.mv[:] = [float(x) for x in np.asarray(mv, np.float32).ravel()]
    c.normal_matrix[:] = [float(x) for x in np.asarray(normal_matrix, np.float32).ravel()]
    c.light_pos[:] = [float(x) for x in np.asarray(light_pos, np.float32).ravel()]
    return c


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle_raster.c", "oracle_shadow.c", "oracle.h", "Makefile", "oracle_rbssm_impl.h", "oracle_edt_impl.h", "oracle_moments_impl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_num_threads.restype = C.c_int
    return _lib


# ---- matrices ---------------------------------------------------------------------------------------------
def perspective(fovy, aspect, zn, zf):
    o = np.zeros(16, np.float32)
    lib().orc_perspective(C.c_float(fovy), C.c_float(aspect), C.c_float(zn), C.c_float(zf), _fp(o))
    return o


def look_at(eye, at, up):
    o = np.zeros(16, np.float32)
    lib().orc_look_at(_fp(_f32(eye)), _fp(_f32(at)), _fp(_f32(up)), _fp(o))
    return o


def mat4_mul(a, b):
    o = np.zeros(16, np.float32)
    lib().orc_mat4_mul(_fp(_f32(a)), _fp(_f32(b)), _fp(o))
    return o


def rotate(angle_deg, axis, m=None):
    o = np.zeros(16, np.float32)
    m = _f32(np.eye(4).ravel() if m is None else m)
    lib().orc_rotate(_fp(m), C.c_float(angle_deg), _fp(_f32(axis)), _fp(o))
    return o


def frame_matrices(cam_eye, cam_at, light_eye, light_at, W, H, SW, SH):
    """All per-frame uniforms of one light, composed as the reference's display() does."""
    up = np.array([0, 0, 1], np.float32)
    L = lib()
    lp = perspective(45.0, np.float32(SW) / np.float32(SH), 1.0, 1000.0)
    lv = look_at(light_eye, light_at, up)
    light_mvp = mat4_mul(mat4_mul(lp, lv), _model_identity())
    cp = perspective(45.0, np.float32(W) / np.float32(H), 1.0, 1000.0)
    cv = look_at(cam_eye, cam_at, up)
    model = _model_identity()
    cam_mvp = mat4_mul(mat4_mul(cp, cv), model)
    cam_mv = mat4_mul(cv, model)
    nm = np.zeros(9, np.float32)
    L.orc_normal_matrix(_fp(cam_mv), _fp(nm))
    lb = np.zeros(16, np.float32)
    L.orc_bias_mul(_fp(light_mvp), _fp(lb))
    ls = np.zeros(3, np.float32)
    L.orc_rotate_light_180(_fp(_f32(light_eye)), _fp(ls))
    return dict(cam_mvp=cam_mvp, cam_mv=cam_mv, normal_matrix=nm, light_mvp=light_mvp, light_mvp_b=lb,
                light_pos_shading=ls)


def _model_identity():
    """model = I * translate(0) * rotate(0,x) * rotate(0,y) * rotate(0,z)  (main.cpp:190-195)"""
    L = lib()
    m = _f32(np.eye(4).ravel())
    t = np.zeros(16, np.float32)
    L.orc_translate(_fp(m), _fp(np.zeros(3, np.float32)), _fp(t))
    m = mat4_mul(m, t)
    for ax in ([1, 0, 0], [0, 1, 0], [0, 0, 1]):
        m = mat4_mul(m, rotate(0.0, ax))
    return m


def uniform_light_sample(p, size, n_lights, index):
    o = np.zeros(3, np.float32)
    lib().orc_uniform_light_sample(_fp(_f32(p)), int(size), int(n_lights), int(index), _fp(o))
    return o


def pcf_offsets(kernel_order, penumbra_size, inclusive):
    o = np.zeros(256, np.float32)
    n = lib().orc_pcf_offsets(int(kernel_order), int(penumbra_size), int(bool(inclusive)), _fp(o), 256)
    if n < 0:
        raise ValueError("bad PCF parameters")
    return o[:n].copy()


# ---- passes -----------------------------------------------------------------------------------------------
def raster_depth(xyz, idx, mvp, W, H, factor=4.0, units=20.0):
    xyz, idx, mvp = _f32(xyz), _i32(idx), _f32(mvp)
    d = np.empty((H, W), np.float32)
    rc = lib().orc_raster_depth(_fp(xyz), xyz.size // 3, _ip(idx), idx.size // 3, _fp(mvp), W, H,
                                C.c_float(factor), C.c_float(units), _fp(d))
    assert rc == 0
    return d


def raster_gbuffer(xyz, nrm, idx, mvp, W, H):
    xyz, nrm, idx, mvp = _f32(xyz), _f32(nrm), _i32(idx), _f32(mvp)
    pos = np.empty((H, W, 4), np.float32)
    nr = np.empty((H, W, 4), np.float32)
    d = np.empty((H, W), np.float32)
    rc = lib().orc_raster_gbuffer(_fp(xyz), _fp(nrm), xyz.size // 3, _ip(idx), idx.size // 3, _fp(mvp), W, H,
                                  _fp(pos), _fp(nr), _fp(d))
    assert rc == 0
    return pos, nr, d


def raster_gbuffer_rgb(xyz, nrm, rgb, idx, mvp, W, H):
    xyz, nrm, rgb, idx, mvp = _f32(xyz), _f32(nrm), _f32(rgb), _i32(idx), _f32(mvp)
    pos, nr, alb = (np.empty((H, W, 4), np.float32) for _ in range(3))
    d = np.empty((H, W), np.float32)
    rc = lib().orc_raster_gbuffer_ex(_fp(xyz), _fp(nrm), _fp(rgb), xyz.size // 3, _ip(idx), idx.size // 3, _fp(mvp), W, H,
                                     _fp(pos), _fp(nr), _fp(alb), _fp(d))
    assert rc == 0
    return pos, nr, alb, d


class Texture(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32)]


def _textures(textures):
    """up to three uint8 [h, w, 3] arrays (None = unbound) -> (orc_texture[3], keep-alive list)"""
    arr = (Texture * 3)()
    keep = []
    for k in range(3):
        t = textures[k] if textures is not None and k < len(textures) else None
        if t is None:
            arr[k] = Texture(None, 0, 0)
        else:
            a = np.ascontiguousarray(t, np.uint8)
            assert a.ndim == 3 and a.shape[2] == 3
            keep.append(a)
            arr[k] = Texture(a.ctypes.data, a.shape[1], a.shape[0])
    return arr, keep


def fragment_color(uvw, rgb, textures):
    """GBuffer.frag:11-30 with useTextureForColoring == 1 on n fragments: uvw [n,3], rgb [n,3] -> [n,4]."""
    uvw, rgb = _f32(uvw).reshape(-1, 3), _f32(rgb).reshape(-1, 3)
    tex, keep = _textures(textures)
    out = np.empty((uvw.shape[0], 4), np.float32)
    f = lib().orc_fragment_color
    for i in range(uvw.shape[0]):
        f(_fp(uvw[i]), _fp(rgb[i]), tex, _fp(out[i]))
    return out


def raster_gbuffer_tex(xyz, nrm, rgb, uv, idx, mvp, W, H, textures):
    """G-buffer with the texture select of GBuffer.frag:11-30: rgb may be None, uv [V,3] = (u, v, texture id)."""
    xyz, nrm, idx, mvp, uv = _f32(xyz), _f32(nrm), _i32(idx), _f32(mvp), _f32(uv)
    rgb = None if rgb is None else _f32(rgb)
    tex, keep = _textures(textures)
    pos, nr, alb = (np.empty((H, W, 4), np.float32) for _ in range(3))
    d = np.empty((H, W), np.float32)
    rc = lib().orc_raster_gbuffer_tex(_fp(xyz), _fp(nrm), _fp(rgb) if rgb is not None else None, _fp(uv), xyz.size // 3, _ip(idx),
                                      idx.size // 3, _fp(mvp), W, H, tex, _fp(pos), _fp(nr), _fp(alb), _fp(d))
    assert rc == 0
    return pos, nr, alb, d


CLEAR_COLOR = np.array([0.63, 0.82, 0.96, 1.0], np.float32)      # shadeScene, ShadowMapping/src/main.cpp:453


def shade_phong(cam, shadow_intensity, pos4, nrm4, albedo4, vis):
    H, W = pos4.shape[:2]
    out = np.empty((H, W, 4), np.float32)
    lib().orc_shade_phong(C.byref(cam), C.c_float(shadow_intensity), _fp(_f32(pos4)), _fp(_f32(nrm4)),
                          _fp(_f32(albedo4)) if albedo4 is not None else None, _fp(_f32(vis)), W, H, _fp(CLEAR_COLOR), _fp(out))
    return out


def visibility(params, cam, light_mvp_b, pos4, nrm4, shadow_map):
    H, W = pos4.shape[:2]
    vis = np.zeros((H, W), np.float32)
    lib().orc_visibility(C.byref(params), C.byref(cam), _fp(_f32(light_mvp_b)), _fp(_f32(pos4)), _fp(_f32(nrm4)),
                         W, H, _fp(_f32(shadow_map)), _fp(vis))
    return vis


EDT_MARKER = -32768


def pcss_tap_count(params, cam, light_mvp_b, pos4, nrm4, shadow_map):
    """(pixels running the blocker search, pixels running the filter loop, foreground pixels) of a PCSS frame."""
    pos4, nrm4, sm, lm = _f32(pos4), _f32(nrm4), _f32(shadow_map), _f32(light_mvp_b)
    H, W = pos4.shape[:2]
    out = np.zeros(3, np.int64)
    lib().orc_pcss_tap_count(C.byref(params), C.byref(cam), _fp(lm), _fp(pos4), _fp(nrm4), W, H, _fp(sm),
                             out.ctypes.data_as(C.POINTER(C.c_int64)))
    return int(out[0]), int(out[1]), int(out[2])


def edt_hard_image(params, cam, cam_mvp, light_mvp_b, pos4, nrm4, shadow_map):
    """(shadow, camera depth, pre-evaluated shadow, 1) target of the RBSM shaders with EDTSM == 1."""
    H, W = pos4.shape[:2]
    img = np.zeros((H, W, 4), np.float32)
    lib().orc_edt_hard_image(C.byref(params), C.byref(cam), _fp(_f32(cam_mvp)), _fp(_f32(light_mvp_b)), _fp(_f32(pos4)),
                             _fp(_f32(nrm4)), W, H, _fp(_f32(shadow_map)), _fp(img))
    return img


def _i16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def edt_sites(img4):
    H, W = img4.shape[:2]
    out = np.empty((H, W, 2), np.int16)
    lib().orc_edt_sites(_fp(_f32(img4)), W, H, _i16p(out))
    return out


def edt_nearest(site2):
    site2 = np.ascontiguousarray(site2, np.int16)
    H, W = site2.shape[:2]
    out = np.empty((H, W, 2), np.int16)
    lib().orc_edt_nearest(_i16p(site2), W, H, _i16p(out))
    return out


def edt_normalize(img4, pos4, near2, penumbra_size, shadow_intensity):
    H, W = img4.shape[:2]
    near2 = np.ascontiguousarray(near2, np.int16)
    out = np.empty((H, W, 2), np.float32)
    lib().orc_edt_normalize(_fp(_f32(img4)), _fp(_f32(pos4)), _i16p(near2), W, H, C.c_float(penumbra_size),
                            C.c_float(shadow_intensity), _fp(out))
    return out


def mean_filter(in2, pos4, cam_mv, order, horizontal, z_near=1, z_far=1000, linear=False):
    H, W = pos4.shape[:2]
    out = np.empty((H, W, 2), np.float32)
    lib().orc_mean_filter(_fp(_f32(in2)), _fp(_f32(pos4)), _fp(_f32(cam_mv)), W, H, int(order), int(bool(horizontal)),
                          int(z_near), int(z_far), int(bool(linear)), _fp(out))
    return out


def edtsm(params, cam, cam_mvp, light_mvp_b, pos4, nrm4, shadow_map):
    """Whole EDTSM pass sequence; returns (visibility[H,W], nearest site[H,W,2] int16)."""
    H, W = pos4.shape[:2]
    vis = np.zeros((H, W), np.float32)
    near = np.empty((H, W, 2), np.int16)
    lib().orc_edtsm(C.byref(params), C.byref(cam), _fp(_f32(cam_mvp)), _fp(_f32(light_mvp_b)), _fp(_f32(pos4)), _fp(_f32(nrm4)),
                    W, H, _fp(_f32(shadow_map)), _fp(vis), _i16p(near))
    return vis, near


def msm_quantization():
    """(mQuantization, mQuantizationInverse, tQuantization) as the uniforms hold them (column-major)."""
    m, mi, t = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(4, np.float32)
    lib().orc_msm_quantization(_fp(m), _fp(mi), _fp(t))
    return m, mi, t


def moment_texel(technique, zwin, zwin_px, zwin_py, x_odd, y_odd, z_near=1, z_far=1000):
    out = np.zeros(4, np.float32)
    lib().orc_moment_texel(TECH[technique], C.c_float(zwin), C.c_float(zwin_px), C.c_float(zwin_py), int(x_odd), int(y_odd),
                           z_near, z_far, _fp(out))
    return out


def raster_moments(xyz, idx, mvp, W, H, technique, factor=4.0, units=20.0, z_near=1, z_far=1000):
    xyz, idx, mvp = _f32(xyz), _i32(idx), _f32(mvp)
    out = np.empty((H, W, 4), np.float32)
    rc = lib().orc_raster_moments(_fp(xyz), xyz.size // 3, _ip(idx), idx.size // 3, _fp(mvp), W, H, C.c_float(factor),
                                  C.c_float(units), TECH[technique], z_near, z_far, _fp(out))
    assert rc == 0
    return out


def gaussian_kernel(order):
    k = np.zeros(34, np.float32)               # `uniform float kernel[33]`: entries past `order` stay 0
    lib().orc_gaussian_kernel(order, _fp(k))
    return k[:order]


def _kernel33(order):
    k = np.zeros(34, np.float32)
    lib().orc_gaussian_kernel(order, _fp(k))
    return k


def filter_moments(src4, W, H, order, horizontal, log_space=False):
    """One pass of filterShadowMap: src4[sh, sw, 4] -> float32[H, W, 4]."""
    src4 = _f32(src4)
    k = _kernel33(order)
    out = np.empty((H, W, 4), np.float32)
    lib().orc_filter_moments(_fp(src4), src4.shape[1], src4.shape[0], W, H, order, _fp(k), int(bool(horizontal)),
                             int(bool(log_space)), _fp(out))
    return out


def filter_shadow_map(mom4, W, H, order, technique):
    """filterShadowMap(), ShadowMapping/src/main.cpp:374-398: X pass then Y pass into W x H targets."""
    logs = technique == "esm"
    return filter_moments(filter_moments(mom4, W, H, order, True, logs), W, H, order, False, logs)


def visibility_moments(params, cam, light_mvp_b, pos4, nrm4, fmap4):
    pos4, nrm4, fmap4, lm = _f32(pos4), _f32(nrm4), _f32(fmap4), _f32(light_mvp_b)
    H, W = pos4.shape[:2]
    vis = np.zeros((H, W), np.float32)
    lib().orc_visibility_moments(C.byref(params), C.byref(cam), _fp(lm), _fp(pos4), _fp(nrm4), W, H, _fp(fmap4),
                                 fmap4.shape[1], fmap4.shape[0], _fp(vis))
    return vis


def visibility_multi(params, light_mvp_b_common, trans4, pos4, shadow_maps):
    H, W = pos4.shape[:2]
    trans4 = _f32(trans4)
    vis = np.zeros((H, W), np.float32)
    lib().orc_visibility_multi(C.byref(params), _fp(_f32(light_mvp_b_common)), trans4.shape[0], _fp(trans4),
                               _fp(_f32(pos4)), W, H, _fp(_f32(shadow_maps)), _fp(vis))
    return vis


def sv_build_prisms(xyz, nrm, idx, light, infinity=100):
    xyz, nrm, idx = _f32(xyz), _f32(nrm), _i32(idx)
    T = idx.size // 3
    pxyz = np.empty((T * 6, 3), np.float32)
    pidx = np.empty((T * 6, 3), np.int32)
    lib().orc_sv_build_prisms(_fp(xyz), _fp(nrm), xyz.size // 3, _ip(idx), T, _fp(_f32(light)), int(infinity),
                              _fp(pxyz), _ip(pidx))
    return pxyz, pidx


def sv_count(prism_xyz, prism_idx, mvp, W, H, scene_depth, depth_func=DEPTH_LEQUAL):
    pxyz, pidx = _f32(prism_xyz), _i32(prism_idx)
    cnt = np.zeros((H, W), np.int32)
    st = np.zeros((H, W), np.uint8)
    rc = lib().orc_sv_count(_fp(pxyz), pxyz.size // 3, _ip(pidx), pidx.size // 3, _fp(_f32(mvp)), W, H,
                            _fp(_f32(scene_depth)), int(depth_func), _ip(cnt), st.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0
    return cnt, st


def sv_silhouette_keep(nrm, idx, light):
    """keep[T, 3]: which side quads survive the pairwise cancellation of interior edges (silhouette form)."""
    nrm, idx = _f32(nrm), _i32(idx)
    T = idx.size // 3
    keep = np.zeros((T, 3), np.uint8)
    lib().orc_sv_silhouette_keep(_fp(nrm), _ip(idx), T, _fp(_f32(light)), keep.ctypes.data_as(C.POINTER(C.c_uint8)))
    return keep


def sv_build_volumes(xyz, nrm, idx, light, infinity=100, keep=None, caps=False):
    """(prism vertices [6T,3], volume triangles [(8 if caps else 6) T, 3]); dropped quads are (0,0,0)."""
    xyz, nrm, idx = _f32(xyz), _f32(nrm), _i32(idx)
    T = idx.size // 3
    per = 8 if caps else 6
    pxyz = np.empty((T * 6, 3), np.float32)
    vidx = np.empty((T * per, 3), np.int32)
    kp = None if keep is None else np.ascontiguousarray(keep, np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
    lib().orc_sv_build_volumes(_fp(xyz), _fp(nrm), xyz.size // 3, _ip(idx), T, _fp(_f32(light)), int(infinity), kp, int(bool(caps)),
                               _fp(pxyz), _ip(vidx))
    return pxyz, vidx


def sv_count_ex(prism_xyz, vol_idx, mvp, W, H, scene_depth, depth_func=DEPTH_LEQUAL, zfail=False, per=6):
    pxyz, pidx = _f32(prism_xyz), _i32(vol_idx)
    cnt = np.zeros((H, W), np.int32)
    st = np.zeros((H, W), np.uint8)
    rc = lib().orc_sv_count_ex(_fp(pxyz), pxyz.size // 3, _ip(pidx), pidx.size // 3, _fp(_f32(mvp)), W, H,
                               _fp(_f32(scene_depth)), int(depth_func), int(bool(zfail)), int(per), _ip(cnt),
                               st.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0
    return cnt, st


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


# ---- the reference's own sources on the CPU (oracle/_ref) -------------------------------------------------
def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "libref_shaders.so")) and os.path.exists(os.path.join(REF_DIR, "libref_host.so"))


def build_ref(reference="/root/reference"):
    if os.path.isdir(reference):
        subprocess.check_call(["python3", os.path.join(HERE, "ref_build", "build_ref.py"), "--reference", reference])
    return ref_available()


class _Binding(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("size", C.c_uint64)]


class _Sampler(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("channels", C.c_int32), ("layers", C.c_int32)]


_ref_sh = None
_ref_host = None


def ref_shaders():
    global _ref_sh
    if _ref_sh is None:
        _ref_sh = C.CDLL(os.path.join(REF_DIR, "libref_shaders.so"))
    return _ref_sh


def ref_host():
    global _ref_host
    if _ref_host is None:
        _ref_host = C.CDLL(os.path.join(REF_DIR, "libref_host.so"))
        _ref_host.ref_scene_load.restype = C.c_void_p
    return _ref_host


def ref_run_shader(name, uniforms, W, H, rect=None):
    """Run one of the reference's fragment shaders (name in build_ref.SHADERS) over the screen.

    uniforms: dict name -> np.ndarray (float32/int32 scalars, vectors, column-major matrices) or
              ('tex', array[H,W] | array[H,W,4] | array[L,H,W]) for samplers.
    Returns gl_FragData[0] as float32[H,W,4]; discarded pixels keep the clear colour (0,0,0,1).
    """
    keep, binds = [], []
    for k, v in uniforms.items():
        if isinstance(v, tuple) and v[0] == "tex":
            a = _f32(v[1])
            if len(v) >= 3 and v[2] == "linear":       # RGBA32F with GL_LINEAR (level 0)
                s = _Sampler(a.ctypes.data, a.shape[1], a.shape[0], 4 | 0x100, 1)
            elif len(v) >= 3 and v[2] == "linear_repeat":   # scene texture: GL_LINEAR, GL_REPEAT
                s = _Sampler(a.ctypes.data, a.shape[1], a.shape[0], 4 | 0x100 | 0x200, 1)
            elif a.ndim == 2:
                s = _Sampler(a.ctypes.data, a.shape[1], a.shape[0], 1, 1)
            elif a.ndim == 3 and a.shape[2] == 4 and len(v) < 3:
                s = _Sampler(a.ctypes.data, a.shape[1], a.shape[0], 4, 1)
            else:       # depth array [L,H,W]
                s = _Sampler(a.ctypes.data, a.shape[2], a.shape[1], 1, a.shape[0])
            keep += [a, s]
            binds.append(_Binding(k.encode(), C.cast(C.pointer(s), C.c_void_p), C.sizeof(s)))
        else:
            a = np.ascontiguousarray(v)
            if a.dtype not in (np.float32, np.int32):
                a = a.astype(np.float32 if a.dtype.kind == "f" else np.int32)
            keep.append(a)
            binds.append(_Binding(k.encode(), a.ctypes.data, a.nbytes))
    arr = (_Binding * len(binds))(*binds)
    out = np.zeros((H, W, 4), np.float32)
    out[..., 3] = 1.0
    x0, y0, x1, y1 = rect if rect else (0, 0, W, H)
    fn = getattr(ref_shaders(), f"ref_{name}_run")
    rc = fn(arr, len(binds), W, H, x0, y0, x1, y1, _fp(out))
    assert rc == 0
    del keep
    return out


def ref_load_scene(config, root="/root/reference"):
    """Load a Configs/*.txt through the reference's own SceneLoader/Mesh/OBJLoader."""
    h = ref_host().ref_scene_load(root.encode(), config.encode())
    if not h:
        raise FileNotFoundError(config)
    h = C.c_void_p(h)
    nv, nt = C.c_int(), C.c_int()
    ref_host().ref_scene_counts(h, C.byref(nv), C.byref(nt))
    xyz = np.empty((nv.value, 3), np.float32)
    nrm = np.empty((nv.value, 3), np.float32)
    idx = np.empty((nt.value, 3), np.int32)
    ref_host().ref_scene_copy(h, _fp(xyz), _fp(nrm), _ip(idx))
    v = [np.zeros(3, np.float32) for _ in range(4)]
    dt = C.c_float()
    ref_host().ref_scene_views(h, _fp(v[0]), _fp(v[1]), _fp(v[2]), _fp(v[3]), C.byref(dt))
    return dict(xyz=xyz, nrm=nrm, idx=idx, cam_eye=v[0], cam_at=v[1], light_eye=v[2], light_at=v[3],
                depth_threshold=np.float32(dt.value))


def ref_frame_matrices(cam_eye, cam_at, light_eye, light_at, W, H, SW, SH):
    o = dict(cam_mvp=np.zeros(16, np.float32), cam_mv=np.zeros(16, np.float32), normal_matrix=np.zeros(9, np.float32),
             light_mvp=np.zeros(16, np.float32), light_mvp_b=np.zeros(16, np.float32),
             light_pos_shading=np.zeros(3, np.float32))
    ref_host().ref_frame_matrices(_fp(_f32(cam_eye)), _fp(_f32(cam_at)), _fp(_f32(light_eye)), _fp(_f32(light_at)),
                                  W, H, SW, SH, _fp(o["cam_mvp"]), _fp(o["cam_mv"]), _fp(o["normal_matrix"]),
                                  _fp(o["light_mvp"]), _fp(o["light_mvp_b"]), _fp(o["light_pos_shading"]))
    return o


def ref_sv_prisms(xyz, nrm, idx, light, infinity=100):
    xyz, nrm, idx = _f32(xyz), _f32(nrm), _i32(idx)
    T = idx.size // 3
    pxyz = np.empty((T * 6, 3), np.float32)
    pidx = np.empty((T * 6, 3), np.int32)
    ref_host().ref_sv_prisms(_fp(xyz), _fp(nrm), xyz.size // 3, _ip(idx), T, _fp(_f32(light)), int(infinity),
                             _fp(pxyz), _ip(pidx))
    return pxyz, pidx


def ref_moment_quantization(typed16):
    t = _f32(typed16)
    m, mi = np.zeros(16, np.float32), np.zeros(16, np.float32)
    ref_host().ref_moment_quantization(_fp(t), _fp(m), _fp(mi))
    return m, mi


def ref_uniform_sample(p, size, n_lights, index):
    o = np.zeros(3, np.float32)
    ref_host().ref_uniform_sample(_fp(_f32(p)), int(size), int(n_lights), int(index), _fp(o))
    return o
