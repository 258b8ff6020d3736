"""ctypes view of include/shadowgi_host.h (libshadowgi_host.so): the C++ host side that mirrors the reference's
SceneLoader / Mesh / OBJ loader / matrix set-up / per-technique render-pass interface."""
import ctypes as C
import os

import numpy as np

from . import capi

PKG = os.path.dirname(os.path.abspath(__file__))
PROGRAM = {"shadow_mapping": 0, "soft_shadow_mapping": 1, "shadow_volumes": 2}

EXPORTS = [
    "sgh_last_error", "sgh_scene_load", "sgh_scene_free", "sgh_scene_counts", "sgh_scene_copy", "sgh_scene_views",
    "sgh_scene_substitutions", "sgh_frame_matrices", "sgh_app_create", "sgh_app_destroy", "sgh_app_error", "sgh_app_context",
    "sgh_app_load_scene", "sgh_app_set_scene", "sgh_app_scene_counts", "sgh_app_scene_copy", "sgh_app_configure", "sgh_app_set_rect", "sgh_app_set_light_shard", "sgh_app_comm_init", "sgh_app_light_costs", "sgh_app_set_light_owners",
    "sgh_app_set_technique", "sgh_app_set_int", "sgh_app_set_float", "sgh_app_upload_scene", "sgh_app_set_texture", "sgh_app_render_shadow_map",
    "sgh_app_render_gbuffer", "sgh_app_filter_shadow_map", "sgh_app_compute_hard_shadows", "sgh_app_render_soft_shadows", "sgh_app_render_monte_carlo",
    "sgh_app_render_shadow_volumes", "sgh_app_shade_scene", "sgh_app_save_image", "sgh_write_png", "sgh_app_display", "sgh_app_display_e2e", "sgh_app_display_e2e_async", "sgh_app_e2e_wait", "sgh_app_step_animation", "sgh_procedural", "sgh_free",
]

_lib = None


def load():
    global _lib
    if _lib is None:
        from . import _build
        capi.load()                       # libshadowgi.so first (the host library links against it)
        _lib = C.CDLL(_build.build_host())
        _lib.sgh_last_error.restype = C.c_char_p
        _lib.sgh_scene_load.restype = C.c_void_p
        _lib.sgh_scene_substitutions.restype = C.c_char_p
        _lib.sgh_app_create.restype = C.c_void_p
        _lib.sgh_app_error.restype = C.c_char_p
        _lib.sgh_app_context.restype = C.c_void_p
        _lib.sgh_app_display_e2e.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
        _lib.sgh_app_display_e2e_async.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)]
        _lib.sgh_app_e2e_wait.argtypes = [C.c_void_p, C.c_int32]
        _lib.sgh_app_save_image.argtypes = [C.c_void_p, C.c_char_p]
        _lib.sgh_app_set_texture.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]
        _lib.sgh_app_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32]
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class HostError(RuntimeError):
    pass


def load_scene(config, base_dir=""):
    """SceneLoader::load -> dict(xyz, nrm, idx, cam_eye, cam_at, light_eye, light_at, depth_threshold, substitutions)."""
    L = load()
    h = L.sgh_scene_load(str(config).encode(), str(base_dir).encode())
    if not h:
        raise HostError(L.sgh_last_error().decode())
    h = C.c_void_p(h)
    try:
        nv, nt = C.c_int32(), C.c_int32()
        L.sgh_scene_counts(h, C.byref(nv), C.byref(nt))
        xyz, nrm = np.empty((nv.value, 3), np.float32), np.empty((nv.value, 3), np.float32)
        idx = np.empty((nt.value, 3), np.int32)
        L.sgh_scene_copy(h, _fp(xyz), _fp(nrm), _ip(idx))
        v = [np.zeros(3, np.float32) for _ in range(4)]
        dt = C.c_float()
        L.sgh_scene_views(h, _fp(v[0]), _fp(v[1]), _fp(v[2]), _fp(v[3]), C.byref(dt))
        subs = [s for s in L.sgh_scene_substitutions(h).decode().split(";") if s]
    finally:
        L.sgh_scene_free(h)
    return dict(xyz=xyz, nrm=nrm, idx=idx, cam_eye=v[0], cam_at=v[1], light_eye=v[2], light_at=v[3],
                depth_threshold=np.float32(dt.value), substitutions=subs)


def frame_matrices(cam_eye, cam_at, light_eye, light_at, W, H, SW, SH):
    o = dict(cam_mvp=np.zeros(16, np.float32), cam_mv=np.zeros(16, np.float32), normal_matrix=np.zeros(9, np.float32),
             light_mvp=np.zeros(16, np.float32), light_mvp_b=np.zeros(16, np.float32), light_pos_shading=np.zeros(3, np.float32))
    rc = load().sgh_frame_matrices(_fp(_f32(cam_eye)), _fp(_f32(cam_at)), _fp(_f32(light_eye)), _fp(_f32(light_at)), W, H, SW, SH,
                                   _fp(o["cam_mvp"]), _fp(o["cam_mv"]), _fp(o["normal_matrix"]), _fp(o["light_mvp"]),
                                   _fp(o["light_mvp_b"]), _fp(o["light_pos_shading"]))
    if rc:
        raise HostError("sgh_frame_matrices failed")
    return o


def write_png(path, rgba8):
    """8-bit RGBA image [H, W, 4], row 0 = top, through the host library's encoder (no GPU needed)."""
    a = np.ascontiguousarray(rgba8, np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4
    L = load()
    L.sgh_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
    rc = L.sgh_write_png(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0])
    if rc:
        raise OSError(f"sgh_write_png({path}) failed")


def procedural(spec):
    L = load()
    xyz, idx = C.POINTER(C.c_float)(), C.POINTER(C.c_int32)()
    nv, nt = C.c_int32(), C.c_int32()
    if L.sgh_procedural(spec.encode(), C.byref(xyz), C.byref(nv), C.byref(idx), C.byref(nt)):
        raise HostError(L.sgh_last_error().decode())
    a = np.ctypeslib.as_array(xyz, (nv.value, 3)).copy()
    b = np.ctypeslib.as_array(idx, (nt.value, 3)).copy()
    L.sgh_free(xyz); L.sgh_free(idx)
    return a, b


class App:
    """sgh::ShadowApp — the reference's display() loop over the C ABI (one per GPU)."""

    def __init__(self, device=0):
        self.L = load()
        h = self.L.sgh_app_create(int(device))
        if not h:
            raise HostError(self.L.sgh_last_error().decode())
        self.h = C.c_void_p(h)
        self.W = self.H = self.SW = self.SH = 0

    def _ck(self, rc):
        if rc:
            raise HostError(f"rc={rc}: " + self.L.sgh_app_error(self.h).decode() + " / " + self.L.sgh_last_error().decode())

    def close(self):
        if self.h:
            self.L.sgh_app_destroy(self.h)
            self.h = None

    def load_scene(self, config, base_dir=""):
        self._ck(self.L.sgh_app_load_scene(self.h, str(config).encode(), str(base_dir).encode()))

    def set_scene(self, sc):
        xyz, nrm, idx = _f32(sc["xyz"]), _f32(sc["nrm"]), np.ascontiguousarray(sc["idx"], np.int32)
        self._ck(self.L.sgh_app_set_scene(self.h, _fp(xyz), _fp(nrm), xyz.size // 3, _ip(idx), idx.size // 3, _fp(_f32(sc["cam_eye"])),
                                          _fp(_f32(sc["cam_at"])), _fp(_f32(sc["light_eye"])), _fp(_f32(sc["light_at"])),
                                          C.c_float(float(sc["depth_threshold"]))))

    def scene_arrays(self):
        nv, nt = C.c_int32(), C.c_int32()
        self._ck(self.L.sgh_app_scene_counts(self.h, C.byref(nv), C.byref(nt)))
        xyz, nrm = np.empty((nv.value, 3), np.float32), np.empty((nv.value, 3), np.float32)
        idx = np.empty((nt.value, 3), np.int32)
        self._ck(self.L.sgh_app_scene_copy(self.h, _fp(xyz), _fp(nrm), _ip(idx)))
        return xyz, nrm, idx

    def configure(self, W, H, SW, SH=None):
        SH = SW if SH is None else SH
        self.W, self.H, self.SW, self.SH = W, H, SW, SH
        self._ck(self.L.sgh_app_configure(self.h, W, H, SW, SH))

    def set_rect(self, x0, y0, x1, y1):
        self._ck(self.L.sgh_app_set_rect(self.h, x0, y0, x1, y1))

    def set_light_shard(self, rank, world):
        self._ck(self.L.sgh_app_set_light_shard(self.h, int(rank), int(world)))

    def comm_init(self, unique_id, rank, world):
        """Join the NCCL communicator inside the library (id from capi.comm_unique_id() on rank 0): the many-light frame is then
        sharded by lights with the exchanges done by ShadowApp::renderMonteCarlo."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.sgh_app_comm_init(self.h, buf, 128, int(rank), int(world)))

    def light_costs(self, n):
        """Depth-pass time (ms) of each of the n lights of the many-light frame, measured one at a time."""
        ms = np.zeros(n, np.float32)
        self._ck(self.L.sgh_app_light_costs(self.h, _fp(ms), int(n)))
        return ms

    def set_light_owners(self, owners):
        o = np.ascontiguousarray(owners, np.int32)
        self._ck(self.L.sgh_app_set_light_owners(self.h, _ip(o), int(o.size)))

    def set_technique(self, name):
        self._ck(self.L.sgh_app_set_technique(self.h, name.encode()))

    def set(self, **kw):
        for k, v in kw.items():
            if isinstance(v, float):
                self._ck(self.L.sgh_app_set_float(self.h, k.encode(), C.c_float(v)))
            else:
                self._ck(self.L.sgh_app_set_int(self.h, k.encode(), int(v)))

    def set_texture(self, index, rgb):
        a = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        self._ck(self.L.sgh_app_set_texture(self.h, int(index), None if a is None else a.ctypes.data_as(C.c_void_p), 0 if a is None else a.shape[1], 0 if a is None else a.shape[0]))

    def upload_scene(self):
        self._ck(self.L.sgh_app_upload_scene(self.h))

    def display(self, program="shadow_mapping"):
        self._ck(self.L.sgh_app_display(self.h, PROGRAM[program]))

    def display_e2e(self, program, which, host_ptr, nbytes):
        self._ck(self.L.sgh_app_display_e2e(self.h, PROGRAM[program], capi.BUF[which], C.c_void_p(host_ptr), nbytes))

    def display_e2e_async(self, program, which, host_ptr, nbytes):
        t = C.c_int32()
        self._ck(self.L.sgh_app_display_e2e_async(self.h, PROGRAM[program], capi.BUF[which], C.c_void_p(host_ptr), C.c_size_t(nbytes), C.byref(t)))
        return t.value

    def e2e_wait(self, ticket):
        self._ck(self.L.sgh_app_e2e_wait(self.h, int(ticket)))

    def step_animation(self, delta=6.0):
        self._ck(self.L.sgh_app_step_animation(self.h, C.c_float(delta)))

    def render_shadow_map(self): self._ck(self.L.sgh_app_render_shadow_map(self.h))
    def render_gbuffer(self): self._ck(self.L.sgh_app_render_gbuffer(self.h))
    def filter_shadow_map(self): self._ck(self.L.sgh_app_filter_shadow_map(self.h))
    def compute_hard_shadows(self): self._ck(self.L.sgh_app_compute_hard_shadows(self.h))
    def render_soft_shadows(self): self._ck(self.L.sgh_app_render_soft_shadows(self.h))
    def render_monte_carlo(self): self._ck(self.L.sgh_app_render_monte_carlo(self.h))
    def render_shadow_volumes(self): self._ck(self.L.sgh_app_render_shadow_volumes(self.h))
    def shade_scene(self): self._ck(self.L.sgh_app_shade_scene(self.h))

    def save_image(self, path):
        """shadeScene() + the frame written as an 8-bit RGBA PNG."""
        self._ck(self.L.sgh_app_save_image(self.h, str(path).encode()))

    def context(self):
        """A capi.Context view over the app's sgi_ctx (borrowed: do not close)."""
        c = capi.Context.__new__(capi.Context)
        c.lib = capi.load()
        c.h = C.c_void_p(self.L.sgh_app_context(self.h))
        nv, nt = C.c_int32(), C.c_int32()
        self.L.sgh_app_scene_counts(self.h, C.byref(nv), C.byref(nt))
        c.W, c.H, c.SW, c.SH, c.N, c.T = self.W, self.H, self.SW, self.SH, 1, nt.value
        c.close = lambda: None
        return c
