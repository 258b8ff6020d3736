"""Developer probe: host time to queue a frame (no waiting) vs the device time of the frame, c2 workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from globalillumination_b200 import hostapi, scenes
w = scenes.WORKLOADS["c2_sponza"]
app = hostapi.App(0)
app.load_scene(scenes.write_config("c2_sponza")); app.configure(w["W"], w["H"], w["S"]); app.set_technique(w["technique"]); app.set(**w["params"])
app.set(animationOn=1)
ctx = app.context()
app.upload_scene()
for _ in range(20):
    app.display(w["program"]); app.step_animation(6.0)
ctx.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    app.display(w["program"]); app.step_animation(6.0)
t1 = time.perf_counter()
ctx.synchronize()
t2 = time.perf_counter()
print(f"display: issue {1e3*(t1-t0)/N:.3f} ms/frame, issue+drain {1e3*(t2-t0)/N:.3f} ms/frame")
t0 = time.perf_counter()
for _ in range(N):
    app.upload_scene(); app.step_animation(6.0)
t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
print(f"upload_scene: issue {1e3*(t1-t0)/N:.3f} ms/call, with drain {1e3*(t2-t0)/N:.3f}")
host = torch.empty(w["W"] * w["H"], dtype=torch.float32).pin_memory()
t0 = time.perf_counter()
for _ in range(N):
    app.display_e2e(w["program"], "visibility", host.data_ptr(), host.numel() * 4); app.step_animation(6.0)
t2 = time.perf_counter()
print(f"display_e2e blocking: {1e3*(t2-t0)/N:.3f} ms/frame")
hosts = [torch.empty(w["W"] * w["H"], dtype=torch.float32).pin_memory() for _ in range(3)]
nb = hosts[0].numel() * 4
def loop(n, upload=True, read=True, depth=3):
    pend = []
    for k in range(n):
        if upload and read:
            pend.append(app.display_e2e_async(w["program"], "visibility", hosts[k % 3].data_ptr(), nb))
        else:
            if upload: app.upload_scene()
            app.display(w["program"])
            if read: pend.append(ctx.read_async("visibility", hosts[k % 3].data_ptr(), nb))
        app.step_animation(6.0)
        if len(pend) >= depth: ctx.read_wait(pend.pop(0))
    for t in pend: ctx.read_wait(t)
    ctx.synchronize()
for name, kw in (("full e2e depth3", {}), ("full e2e depth2", dict(depth=2)), ("no upload", dict(upload=False)), ("no readback", dict(read=False)), ("neither", dict(upload=False, read=False))):
    loop(10, **kw)
    t0 = time.perf_counter(); loop(N, **kw); dt = (time.perf_counter() - t0) / N
    print(f"{name}: {1e3*dt:.3f} ms/frame -> {1/dt:.0f} fps")
def loop_t(n, depth=3):
    pend = []; ti = tw = 0.0
    for k in range(n):
        a = time.perf_counter()
        pend.append(app.display_e2e_async(w["program"], "visibility", hosts[k % 3].data_ptr(), nb))
        app.step_animation(6.0)
        b = time.perf_counter()
        if len(pend) >= depth: ctx.read_wait(pend.pop(0))
        c = time.perf_counter()
        ti += b - a; tw += c - b
    for t in pend: ctx.read_wait(t)
    ctx.synchronize()
    return ti / n, tw / n
loop_t(10)
t0 = time.perf_counter(); ti, tw = loop_t(N); dt = (time.perf_counter() - t0) / N
print(f"full e2e: {1e3*dt:.3f} ms/frame; host inside display_e2e_async {1e3*ti:.3f} ms, inside read_wait {1e3*tw:.3f} ms")
