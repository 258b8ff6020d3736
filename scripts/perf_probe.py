"""Ad-hoc per-pass timing probe (developer tool, not the bench): golden scenes at bench-like sizes."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from globalillumination_b200 import capi
from tests import util

def probe(name, W, H, S, tech, iters=20, **kw):
    sc = util.scene(name)
    ctx = capi.Context(0)
    fm = util.frame(sc, W, H, S)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_lights(fm["light_mvp"], fm["light_mvp_b"], fm["light_pos_shading"], S, S)
    ctx.set_params(capi.default_params(tech, depth_threshold=float(sc["depth_threshold"]), **kw))
    for _ in range(3):
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    ctx.synchronize()
    ctx.enable_timing(True); ctx.reset_timing()
    t0 = time.time()
    for _ in range(iters):
        ctx.render_shadow_map(); ctx.render_gbuffer(); ctx.compute_visibility()
    ctx.synchronize()
    wall = (time.time() - t0) / iters * 1e3
    out = {p: ctx.pass_time_ms(p)[0] / max(1, ctx.pass_time_ms(p)[1]) for p in ("shadow_map", "gbuffer", "visibility", "vis_kernel")}
    print(f"{name} {W}x{H} S={S} {tech} {kw}: wall {wall:.3f} ms/frame | " + " ".join(f"{k}={v:.3f}ms" for k, v in out.items()), flush=True)
    ctx.close()

def probe_app(workload, iters=10, overlap=True):
    """A bench.py workload through the C++ host (ShadowApp).  overlap=False: passes one after the other on one stream, so the
    per-kernel event times are the kernels' own durations."""
    from globalillumination_b200 import hostapi, scenes
    w = scenes.WORKLOADS[workload]
    app = hostapi.App(0)
    app.load_scene(scenes.write_config(workload))
    app.configure(w["W"], w["H"], w["S"]); app.set_technique(w["technique"]); app.set(**w["params"])
    ctx = app.context()
    if not overlap:
        ctx.set_option("overlap_passes", 0)
    for _ in range(3):
        app.display(w["program"])
    ctx.synchronize()
    ctx.enable_timing(True); ctx.reset_timing()
    t0 = time.time()
    for _ in range(iters):
        app.display(w["program"])
    ctx.synchronize()
    wall = (time.time() - t0) / iters * 1e3
    out = {p: ctx.pass_time_ms(p)[0] / iters for p in ("shadow_map", "gbuffer", "visibility", "vis_kernel", "tile_depth", "tile_gbuffer")}
    print(f"{workload}: wall {wall:.3f} ms/frame | " + " ".join(f"{k}={v:.3f}ms" for k, v in out.items()), flush=True)
    app.close()


def probe_sv(name, W, H, iters=10):
    sc = util.scene(name)
    ctx = capi.Context(0)
    fm = util.frame(sc, W, H, 64)
    ctx.set_mesh(sc["xyz"], sc["nrm"], sc["idx"])
    ctx.set_camera(fm["cam_mvp"], fm["cam_mv"], fm["normal_matrix"], W, H)
    ctx.set_params(capi.default_params("hard"))
    for _ in range(2):
        ctx.render_gbuffer(); ctx.compute_shadow_volume(sc["light_eye"])
    ctx.synchronize()
    ctx.enable_timing(True); ctx.reset_timing()
    t0 = time.time()
    for _ in range(iters):
        ctx.render_gbuffer(); ctx.compute_shadow_volume(sc["light_eye"])
    ctx.synchronize()
    wall = (time.time() - t0) / iters * 1e3
    out = {p: ctx.pass_time_ms(p)[0] / iters for p in ("gbuffer", "shadow_volume", "tile_sv")}
    cnt = ctx.read("sv_count")
    print(f"SV {name} {W}x{H} T={sc['idx'].shape[0]}: wall {wall:.3f} ms/frame | " + " ".join(f"{k}={v:.3f}ms" for k, v in out.items()) +
          f" | shadowed {float((cnt != 0).mean()):.3f} max|count| {int(abs(cnt).max())}", flush=True)
    ctx.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "app":
    probe_app(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 10)
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sv":
    probe_sv(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
elif __name__ == "__main__" and len(sys.argv) > 1:
    name, W, H, S, tech = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    probe(name, W, H, S, tech, iters=int(sys.argv[6]) if len(sys.argv) > 6 else 5)
elif __name__ == "__main__":
    probe("teapot", 1280, 720, 1024, "hard")
    probe("teapot", 1920, 1080, 2048, "pcf")
    probe("teapot", 1920, 1080, 2048, "pcss")
    probe("teapot", 1920, 1080, 2048, "pcss", kernel_size=7)
    probe("dragon", 1920, 1080, 2048, "pcss")
    probe("dragon", 3840, 2160, 4096, "rbsm_noncons")
    probe("dragon", 3840, 2160, 4096, "rbsm_cons")
    probe("dragon", 3840, 2160, 4096, "hard")
    probe("teapot", 7680, 4320, 8192, "hard", iters=5)
