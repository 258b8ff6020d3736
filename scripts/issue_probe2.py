"""Developer probe: what the asynchronous read-back costs on the device timeline (c2 workload)."""
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
hosts = [torch.empty(w["W"] * w["H"], dtype=torch.float32).pin_memory() for _ in range(3)]
N = 300
def loop(n, nbytes, depth=3):
    pend = []
    for k in range(n):
        app.display(w["program"])
        if nbytes: pend.append(ctx.read_async("visibility", hosts[k % 3].data_ptr(), nbytes))
        app.step_animation(6.0)
        if len(pend) >= depth: ctx.read_wait(pend.pop(0))
    for t in pend: ctx.read_wait(t)
    ctx.synchronize()
for nbytes in (0, 4, 1 << 20, 4 << 20, w["W"] * w["H"] * 4):
    loop(10, nbytes)
    t0 = time.perf_counter(); loop(N, nbytes); dt = (time.perf_counter() - t0) / N
    print(f"read {nbytes:>9d} B/frame: {1e3*dt:.3f} ms/frame")
for name in ("overlap_passes",):
    ctx.set_option(name, 0)
    loop(10, w["W"] * w["H"] * 4)
    t0 = time.perf_counter(); loop(N, w["W"] * w["H"] * 4); dt = (time.perf_counter() - t0) / N
    print(f"full read, {name}=0: {1e3*dt:.3f} ms/frame")
    loop(10, 0)
    t0 = time.perf_counter(); loop(N, 0); dt = (time.perf_counter() - t0) / N
    print(f"no read, {name}=0: {1e3*dt:.3f} ms/frame")
