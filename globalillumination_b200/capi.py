"""ctypes view of include/shadowgi.h (libshadowgi.so).  Thin: every method is one C-ABI call.

There is no CPU fallback: if the library is missing it must be built (globalillumination_b200._build), and
sgi_create fails loudly on a box without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))

TECH = {"hard": 0, "pcf": 1, "pcss": 2, "rbsm_noncons": 3, "rbsm_cons": 4, "rpcf_noncons": 5, "rpcf_cons": 6,
        "rsmss": 7, "multi_hard": 8, "rbssm": 9, "edtsm_noncons": 10, "edtsm_cons": 11,
        "vsm": 12, "esm": 13, "evsm": 14, "msm": 15, "pcf_tricubic": 16}
MOMENT_TECHS = ("vsm", "esm", "evsm", "msm")
BUF = {"shadow_map": 0, "gbuf_pos": 1, "gbuf_nrm": 2, "cam_depth": 3, "visibility": 4, "sv_count": 5,
       "sv_stencil": 6, "sv_prism_xyz": 7, "sv_prism_idx": 8, "gbuf_albedo": 9, "shaded": 10, "edt_nearest": 11,
       "moments": 12, "moments_x": 13, "moments_filtered": 14, "prim_id": 15, "light_mask": 16}
PASS = {"shadow_map": 0, "gbuffer": 1, "visibility": 2, "shadow_volume": 3, "vis_kernel": 4, "tile_depth": 5,
        "tile_gbuffer": 6, "tile_sv": 7, "moment_filter": 8}
DEPTH_LESS, DEPTH_LEQUAL = 0, 1
SGI_ERR_OVERFLOW = -4


class SgiParams(C.Structure):
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


EXPORTS = [
    "sgi_create", "sgi_destroy", "sgi_set_stream", "sgi_set_mesh", "sgi_set_mesh_colors", "sgi_shade_phong", "sgi_set_camera", "sgi_set_lights", "sgi_set_params", "sgi_set_multi_light_common", "sgi_set_option",
    "sgi_default_params", "sgi_render_shadow_map", "sgi_render_gbuffer", "sgi_compute_visibility", "sgi_filter_shadow_map", "sgi_moment_quantization",
    "sgi_compute_shadow_volume", "sgi_read", "sgi_read_async", "sgi_read_wait", "sgi_device_ptr", "sgi_synchronize", "sgi_join", "sgi_enable_timing",
    "sgi_render_prim_ids", "sgi_sv_fragments", "sgi_divide_selftest", "sgi_set_mesh_uv", "sgi_set_texture", "sgi_comm_unique_id", "sgi_comm_init", "sgi_comm_destroy", "sgi_comm_strip", "sgi_gather", "sgi_reduce_lights", "sgi_set_light_ids",
    "sgi_pass_time_ms", "sgi_reset_timing", "sgi_alloc_host", "sgi_free_host", "sgi_register_host", "sgi_unregister_host", "sgi_kernel_launches", "sgi_last_error", "sgi_version",
]

_lib = None


def library_path():
    return os.path.join(PKG, "libshadowgi.so")


def load():
    """dlopen libshadowgi.so (building it in-tree first if it is stale or missing)."""
    global _lib
    if _lib is None:
        from . import _build
        _lib = C.CDLL(_build.build_cuda())
        _lib.sgi_last_error.restype = C.c_char_p
        _lib.sgi_version.restype = C.c_char_p
        _lib.sgi_read.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        _lib.sgi_device_ptr.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        _lib.sgi_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        _lib.sgi_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        _lib.sgi_set_texture.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]
        _lib.sgi_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32]
    return _lib


class SgiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"shadowgi error {code}: {msg}")
        self.code = code


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def moment_quantization():
    """sgi_moment_quantization: (mQuantization, mQuantizationInverse, tQuantization), host arithmetic only."""
    m, mi, t = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(4, np.float32)
    load().sgi_moment_quantization(_fp(m), _fp(mi), _fp(t))
    return m, mi, t


def comm_unique_id():
    """sgi_comm_unique_id: the 128-byte NCCL id (made on rank 0; ship it to the other ranks and pass it to comm_init)."""
    buf = (C.c_char * 128)()
    rc = load().sgi_comm_unique_id(buf, 128)
    if rc != 0:
        raise SgiError(rc, "sgi_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf)


def default_params(technique="hard", **kw):
    p = SgiParams()
    load().sgi_default_params(C.byref(p))
    p.technique = TECH[technique] if isinstance(technique, str) else int(technique)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Context:
    """One sgi_ctx (one GPU)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.sgi_create(C.byref(self.h), int(device))
        if rc != 0:
            raise SgiError(rc, "sgi_create failed (no CUDA device? this library has no CPU path)")
        self.W = self.H = self.N = self.SW = self.SH = self.T = 0

    def _ck(self, rc):
        if rc != 0:
            raise SgiError(rc, self.lib.sgi_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.sgi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.sgi_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_mesh(self, xyz, nrm, idx):
        xyz, nrm = _f32(xyz), _f32(nrm)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.T = idx.size // 3
        self._ck(self.lib.sgi_set_mesh(self.h, _fp(xyz), _fp(nrm), xyz.size // 3,
                                       idx.ctypes.data_as(C.POINTER(C.c_int32)), self.T))

    def set_mesh_raw(self, xyz_ptr, nrm_ptr, V, idx_ptr, T):
        """Borrowed host pointers (e.g. pinned memory) — no numpy conversion on the hot e2e path."""
        self.T = T
        self._ck(self.lib.sgi_set_mesh(self.h, C.c_void_p(xyz_ptr), C.c_void_p(nrm_ptr), int(V), C.c_void_p(idx_ptr), int(T)))

    def set_mesh_colors(self, rgb):
        if rgb is None:
            self._ck(self.lib.sgi_set_mesh_colors(self.h, None))
        else:
            self._ck(self.lib.sgi_set_mesh_colors(self.h, _fp(_f32(rgb))))

    def set_mesh_uv(self, uv):
        self._ck(self.lib.sgi_set_mesh_uv(self.h, None if uv is None else _fp(_f32(uv))))

    def set_texture(self, index, rgb):
        """rgb: uint8 [h, w, 3] (row 0 = t 0) or None to unbind texture<index>."""
        if rgb is None:
            self._ck(self.lib.sgi_set_texture(self.h, int(index), None, 0, 0))
        else:
            a = np.ascontiguousarray(rgb, np.uint8)
            self._ck(self.lib.sgi_set_texture(self.h, int(index), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0]))

    def shade_phong(self, clear=(0.63, 0.82, 0.96, 1.0)):
        self._ck(self.lib.sgi_shade_phong(self.h, _fp(_f32(clear))))

    def set_camera(self, mvp, mv, normal_matrix, W, H):
        self.W, self.H = int(W), int(H)
        self._ck(self.lib.sgi_set_camera(self.h, _fp(_f32(mvp)), _fp(_f32(mv)), _fp(_f32(normal_matrix)), self.W, self.H))

    def set_lights(self, light_mvp, light_mvp_b, light_pos_shading, SW, SH):
        a, b = _f32(light_mvp).reshape(-1, 16), _f32(light_mvp_b).reshape(-1, 16)
        assert a.shape == b.shape
        self.N, self.SW, self.SH = a.shape[0], int(SW), int(SH)
        self._ck(self.lib.sgi_set_lights(self.h, self.N, _fp(a), _fp(b), _fp(_f32(light_pos_shading)), self.SW, self.SH))

    def set_multi_light_common(self, light_mvp_b):
        if light_mvp_b is None:
            self._ck(self.lib.sgi_set_multi_light_common(self.h, None))
        else:
            self._ck(self.lib.sgi_set_multi_light_common(self.h, _fp(_f32(light_mvp_b))))

    def set_option(self, name, value):
        self._ck(self.lib.sgi_set_option(self.h, name.encode(), int(value)))

    def set_params(self, params):
        self.params = params
        self._ck(self.lib.sgi_set_params(self.h, C.byref(params)))

    def render_shadow_map(self):
        self._ck(self.lib.sgi_render_shadow_map(self.h))

    def render_gbuffer(self):
        self._ck(self.lib.sgi_render_gbuffer(self.h))

    def render_prim_ids(self):
        self._ck(self.lib.sgi_render_prim_ids(self.h))

    # ---- multi-GPU (NCCL inside the library) ----
    def comm_init(self, unique_id, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.sgi_comm_init(self.h, buf, 128, int(rank), int(nranks)))

    def comm_destroy(self):
        self._ck(self.lib.sgi_comm_destroy(self.h))

    def comm_strip(self, rank):
        r0, r1 = C.c_int32(), C.c_int32()
        self._ck(self.lib.sgi_comm_strip(self.h, int(rank), C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def gather(self, which):
        self._ck(self.lib.sgi_gather(self.h, BUF[which]))

    def reduce_lights(self, total_lights):
        self._ck(self.lib.sgi_reduce_lights(self.h, int(total_lights)))

    def set_light_ids(self, ids, total_lights):
        """Light shards with lit masks (params.multi_partial = 2): index in the whole set of each light of set_lights."""
        a = np.ascontiguousarray(ids, np.int32)
        self._ck(self.lib.sgi_set_light_ids(self.h, int(a.size), a.ctypes.data_as(C.POINTER(C.c_int32)), int(total_lights)))
        self.mask_total = int(total_lights)

    def read_light_mask(self):
        """SGI_BUF_LIGHT_MASK on one rank (no communicator): uint32 lit mask per pixel, assembled from the byte planes."""
        planes = (self.mask_total + 7) // 8
        raw = np.empty((planes, self.H, self.W), np.uint8)
        self._ck(self.lib.sgi_read(self.h, BUF["light_mask"], raw.ctypes.data, raw.nbytes))
        out = np.zeros((self.H, self.W), np.uint32)
        for p in range(planes):
            out |= raw[p].astype(np.uint32) << np.uint32(8 * p)
        return out

    def filter_shadow_map(self):
        self._ck(self.lib.sgi_filter_shadow_map(self.h))

    def compute_visibility(self):
        self._ck(self.lib.sgi_compute_visibility(self.h))

    def compute_shadow_volume(self, light_pos):
        self._ck(self.lib.sgi_compute_shadow_volume(self.h, _fp(_f32(light_pos))))

    def sv_fragments(self):
        n = C.c_int64()
        self._ck(self.lib.sgi_sv_fragments(self.h, C.byref(n)))
        return n.value

    def divide_selftest(self, n, seed=1):
        m = C.c_uint64()
        self._ck(self.lib.sgi_divide_selftest(self.h, C.c_uint64(n), C.c_uint32(seed), C.byref(m)))
        return m.value

    def synchronize(self):
        self._ck(self.lib.sgi_synchronize(self.h))

    def _shape(self, which):
        W, H, N, SW, SH, T = self.W, self.H, self.N, self.SW, self.SH, self.T
        return {
            "shadow_map": ((N, SH, SW), np.float32), "gbuf_pos": ((H, W, 4), np.float32), "gbuf_nrm": ((H, W, 4), np.float32),
            "cam_depth": ((H, W), np.float32), "visibility": ((H, W), np.float32), "sv_count": ((H, W), np.int32),
            "sv_stencil": ((H, W), np.uint8), "gbuf_albedo": ((H, W, 4), np.float32), "shaded": ((H, W, 4), np.float32), "edt_nearest": ((H, W, 2), np.int16), "sv_prism_xyz": ((T * 6, 3), np.float32), "sv_prism_idx": ((T * (8 if getattr(getattr(self, "params", None), "sv_zfail", 0) else 6), 3), np.int32),
            "moments": ((SH, SW, 4), np.float32), "moments_x": ((H, W, 4), np.float32), "moments_filtered": ((H, W, 4), np.float32),
            "prim_id": ((H, W), np.uint32),
        }[which]

    def read(self, which, out=None):
        shape, dt = self._shape(which)
        if out is None:
            out = np.empty(shape, dt)
        self._ck(self.lib.sgi_read(self.h, BUF[which], out.ctypes.data, out.nbytes))
        return out

    def read_raw(self, which, host_ptr, nbytes):
        self._ck(self.lib.sgi_read(self.h, BUF[which], C.c_void_p(host_ptr), nbytes))

    def read_async(self, which, host_ptr, nbytes):
        t = C.c_int32()
        self._ck(self.lib.sgi_read_async(self.h, BUF[which], C.c_void_p(host_ptr), C.c_size_t(nbytes), C.byref(t)))
        return t.value

    def read_wait(self, ticket):
        self._ck(self.lib.sgi_read_wait(self.h, int(ticket)))

    def device_ptr(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.sgi_device_ptr(self.h, BUF[which], C.byref(p), C.byref(n)))
        return p.value, n.value

    def join(self):
        self._ck(self.lib.sgi_join(self.h))

    def enable_timing(self, on=True):
        self._ck(self.lib.sgi_enable_timing(self.h, int(bool(on))))

    def reset_timing(self):
        self._ck(self.lib.sgi_reset_timing(self.h))

    def pass_time_ms(self, which):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.sgi_pass_time_ms(self.h, PASS[which], C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_launches(self):
        n = C.c_int64()
        self._ck(self.lib.sgi_kernel_launches(self.h, C.byref(n)))
        return n.value
