import time, sys
sys.path.insert(0, "/root/repo")
import torch
from globalillumination_b200 import hostapi, scenes
w = scenes.WORKLOADS["c2_sponza"]
cfg = scenes.write_config("c2_sponza")
app = hostapi.App(0)
app.load_scene(cfg); app.configure(w["W"], w["H"], w["S"]); app.set_technique(w["technique"]); app.set(**w["params"])
app.set(animationOn=1, animation=-1800.0)
ctx = app.context()
for _ in range(50):
    app.display(w["program"]); app.step_animation(6.0)
ctx.synchronize()
for n in (200, 2000):
    t0 = time.perf_counter()
    for _ in range(n):
        app.display(w["program"]); app.step_animation(6.0)
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print(f"n={n}: host enqueue {1e6*(t1-t0)/n:.1f} us/frame, total {1e6*(t2-t0)/n:.1f} us/frame, drain {1e3*(t2-t1):.2f} ms")
