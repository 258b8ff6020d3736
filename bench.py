#!/usr/bin/env python3
"""bench.py — frames/s of the shadow hot path on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2_sponza|c5_many_light]

A "step" is one frame of the hot path = light-view depth pass (K1) + camera G-buffer pass (K2) + per-pixel
shadow pass (K3) on the workload BASELINE.json's metric is quoted on (c2: Sponza, 1920x1080, 2048^2 PCSS),
driven through the C++ host side (SceneLoader -> ShadowApp::display) and the C ABI.  The light moves by the
reference's animation step every frame (ShadowMapping/src/main.cpp:217,481).

  value   frames/s, geometry resident in HBM, no readback; per-step CUDA-event times on the context's stream,
          L2 flushed (256 MiB memset) before every timed step, max over ranks
  e2e     frames/s through ShadowApp's host-buffer entry (geometry re-uploaded from pinned host memory every
          frame as the reference's loadVBOs does + visibility read back to pinned host memory)
  roofline  the kernel with the largest share of the step, timed with CUDA events around its launches
  cpu_baseline  the CPU oracle (oracle/, a port of the reference's passes) on the host cores, a few frames
N > 1: frame-parallel — every rank renders its own frames of the same animation (no data-path collective,
"scaling": "weak"); `value` is the sum over ranks / max time.  With --workload c5_many_light --mode lights the
ranks instead share ONE frame: each builds and samples the depth maps of its own lights and the partial sums are
all-reduced over NCCL ("scaling": "strong").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec at 1920x1080 (Sponza, 2048^2 PCSS)"
ANIMATION_STEP = 6.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def metric_name(workload, w):
    """BASELINE.json's metric for the headline workload; the other configurations are named after themselves."""
    return METRIC if workload == "c2_sponza" else f"frames/sec at {w['W']}x{w['H']} ({workload}, {w['program']}/{w['technique']})"


def bench_config(workload, w, V, T):
    """`config` of the JSON line: the workload only, identical in both arms (what differs per run lives under `run`)."""
    n_l = w["params"].get("numberOfSamples", 1) if w["technique"] == "montecarlo" else 1
    return {"workload": workload, "W": w["W"], "H": w["H"], "shadow_map": w["S"], "technique": w["technique"], "lights": n_l,
            "params": w["params"], "triangles": int(T), "vertices": int(V), "scene": w["scene"]}


def algorithmic_bytes(w, V, T, L):
    """SURVEY.md §8(d) / BASELINE.md §3 per-frame algorithmic bytes of each pass."""
    px, S = w["W"] * w["H"], w["S"]
    # G-buffer: 32 B/px (position + normal), 48 B/px when the scene has vertex colours and the albedo target is written too
    gb = (48 if any(l.startswith("c ") for l in w["lines"]) else 32) * px + 12 * (V + T)
    ab = {"shadow_map": L * (4 * S * S + 12 * (V + T)), "gbuffer": gb, "visibility": 36 * px + L * 4 * S * S}
    ab["shadow_volume"] = 8 * px + 36 * T + 144 * T          # depth read + count write + geometry read + prisms written and read back
    if w["technique"] in ("vsm", "esm", "evsm", "msm"):
        # moment shadow maps: float4 moment target written once; blur X reads it and writes a window-sized float4 target, blur Y
        # reads that and writes another; the shadow pass reads G-buffer + filtered map and writes the visibility
        ab["shadow_map"] = 16 * S * S + 12 * (V + T)
        ab["moment_filter"] = 16 * S * S + 48 * px
        ab["visibility"] = 36 * px + 16 * px
    return ab


def run_reference(args, w, cfg_path):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference's GL
    passes cannot run here: no OpenGL; its shaders compiled as C++ are used as the oracle's pin)."""
    from oracle import oracle_py as O
    from globalillumination_b200 import hostapi, scenes
    # all the host threads the process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the
    # CPU arm on one core at N > 1
    try:
        n_host = len(os.sched_getaffinity(0))
    except AttributeError:
        n_host = os.cpu_count() or 1
    O.set_num_threads(n_host)
    sc = scenes.golden_scene(w["golden"]) if w.get("golden") else hostapi.load_scene(cfg_path)
    W, H, S = w["W"], w["H"], w["S"]
    n_l = w["params"].get("numberOfSamples", 1) if w["technique"] == "montecarlo" else 1
    tech = {"pcss": "pcss", "montecarlo": "multi_hard", "naive": "hard", "smsr": "rbsm_noncons", "vsm": "vsm", "esm": "esm", "evsm": "evsm",
            "msm": "msm", "pcf": "pcf"}[w["technique"]]
    p = O.default_params(tech, S, depth_threshold=float(sc["depth_threshold"]),
                         **{k2: w["params"][k1] for k1, k2 in (("blockerSearchSize", "blocker_search_size"), ("kernelSize", "kernel_size"),
                                                                ("lightSourceRadius", "light_source_radius"), ("kernelOrder", "kernel_order"),
                                                                ("penumbraSize", "penumbra_size")) if k1 in w["params"]})

    def frame(anim):
        le = sc["light_eye"]
        if anim is not None:
            r = O.rotate(anim / 10.0, [0, 1, 0]).reshape(4, 4).T[:3, :3]
            le = (r @ le).astype(np.float32)
        if w["program"] == "shadow_volumes":                    # ShadowVolumes/src/main.cpp:126-172
            fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], le, sc["light_at"], W, H, S, S)
            _, _, depth = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
            pxyz, pidx = O.sv_build_prisms(sc["xyz"], sc["nrm"], sc["idx"], le)
            return O.sv_count(pxyz, pidx, fm["cam_mvp"], W, H, depth)[0]
        if n_l == 1 and tech not in O.MOMENT_TECHS:
            fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], le, sc["light_at"], W, H, S, S)
            sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
            pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
            cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
            return O.visibility(p, cam, fm["light_mvp_b"], pos, nrm, sm)
        if tech in O.MOMENT_TECHS:                               # ShadowMapping/src/main.cpp:459-472 with VSM / ESM / EVSM / MSM
            fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], le, sc["light_at"], W, H, S, S)
            mom = O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, tech)
            fmap = O.filter_shadow_map(mom, W, H, p.kernel_order, tech)
            pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
            cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
            return O.visibility_moments(p, cam, fm["light_mvp_b"], pos, nrm, fmap)
        fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], le, sc["light_at"], W, H, S, S)
        pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        maps, mvpb = [], []
        for i in range(n_l):
            e = O.uniform_light_sample(le, w["params"]["lightSourceSize"], n_l, i)
            a = O.uniform_light_sample(sc["light_at"], w["params"]["lightSourceSize"], n_l, i)
            f2 = O.frame_matrices(sc["cam_eye"], sc["cam_at"], e, a, W, H, S, S)
            maps.append(O.raster_depth(sc["xyz"], sc["idx"], f2["light_mvp"], S, S)); mvpb.append(f2["light_mvp_b"])
        mvpb = np.stack(mvpb)
        return O.visibility_multi(p, mvpb[-1], mvpb[:, 12:16], pos, np.stack(maps))

    static_light = w["program"] == "shadow_volumes"          # as in our arm (see main)
    anim = -1800.0
    for _ in range(args.warmup):
        frame(None if static_light else anim); anim += ANIMATION_STEP
    # the whole run has to end within a few minutes whatever --steps says: frames beyond a 150 s budget are not run
    # (the frame time of this arm is stable to a few per cent, so fewer frames give the same figure); "steps" reports what ran
    requested = args.steps
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        frame(None if static_light else anim); anim += ANIMATION_STEP
        done += 1
        if done >= 3 and time.perf_counter() - t0 > 150.0:
            break
    dt = time.perf_counter() - t0
    args = argparse.Namespace(**{**vars(args), "steps": done})
    fps = args.steps / dt
    cores = O.num_threads()
    taps = None
    if tech == "pcss":
        # exact tap count of one frame (the first of the animation): which pixels run the blocker search, which the filter loop
        r = O.rotate(-1800.0 / 10.0, [0, 1, 0]).reshape(4, 4).T[:3, :3]
        fm = O.frame_matrices(sc["cam_eye"], sc["cam_at"], (r @ sc["light_eye"]).astype(np.float32), sc["light_at"], W, H, S, S)
        sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
        pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        n_search, n_filter, n_fg = O.pcss_tap_count(p, cam, fm["light_mvp_b"], pos, nrm, sm)
        bsz = 2 * int((p.blocker_search_size - 1) * 0.5) + 1
        ksz = 2 * int((p.kernel_size - 1) * 0.5) + 1
        taps = {"foreground_pixels": n_fg, "blocker_search_pixels": n_search, "filter_pixels": n_filter,
                "taps": bsz * bsz * n_search + ksz * ksz * n_filter, "frame": "first frame of the animation (counted with the CPU port)"}
    return {
        "pcss_taps": taps,
        "impl": "reference", "metric": metric_name(args.workload, w), "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, w, sc["xyz"].shape[0], sc["idx"].shape[0]), "run": {"steps_requested": requested},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full frames (depth + G-buffer + shadow pass) of the same workload, OpenMP over {cores} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_secondary(local_rank, names=("dragon_pcss", "c3_dragon", "c1_teapot", "c4_tree_sv", "c4_tree_sv_pertri"), steps=60, warmup=6):
    """Short runs of the other configurations through the same loop as the headline (inputs resident, L2 flushed before every
    timed step, CUDA events per step), so that the driver-run line carries them: PCSS where the filter loop really runs
    (dragon_pcss), c3 (Dragon 4K RBSM), c1 (Teapot hard), c4 (shadow volumes: silhouette form and the reference's per-triangle
    prisms, with prism fragments per second - the unit SURVEY 8(d) names for that pass)."""
    import torch
    from globalillumination_b200 import capi, hostapi, scenes
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    for name in names:
        w = scenes.WORKLOADS[name]
        app = hostapi.App(local_rank)
        try:
            if w.get("golden"):
                app.set_scene(scenes.golden_scene(w["golden"]))
            else:
                app.load_scene(scenes.write_config(name))
            app.configure(w["W"], w["H"], w["S"]); app.set_technique(w["technique"]); app.set(**w["params"])
            sv = w["program"] == "shadow_volumes"
            app.set(animationOn=0) if sv else app.set(animationOn=1, animation=-1800.0)
            ctx = app.context()
            stream = torch.cuda.Stream(device=local_rank)
            ctx.set_stream(stream.cuda_stream)
            with torch.cuda.stream(stream):
                for attempt in range(6):
                    try:
                        for _ in range(warmup):
                            app.display(w["program"]); app.step_animation(ANIMATION_STEP)
                        ctx.synchronize()
                        break
                    except (capi.SgiError, hostapi.HostError) as e:
                        if "overflow" not in str(e) or attempt == 5:
                            raise
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                for k in range(steps):
                    flush.zero_()
                    ev[k][0].record(stream)
                    app.display(w["program"]); ctx.join()
                    ev[k][1].record(stream)
                    app.step_animation(ANIMATION_STEP)
                ctx.synchronize()
                ms = sum(a.elapsed_time(b) for a, b in ev) / steps
                rec = {"value": 1e3 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": steps, "W": w["W"], "H": w["H"], "shadow_map": w["S"],
                       "technique": w["technique"], "params": w["params"], "scene": w["scene"]}
                if sv:
                    ctx.set_option("sv_count_fragments", 1)
                    app.display(w["program"]); ctx.synchronize()
                    frags = ctx.sv_fragments()
                    ctx.set_option("sv_count_fragments", 0)
                    ctx.set_option("overlap_passes", 0); ctx.enable_timing(True); ctx.reset_timing()
                    for _ in range(5):
                        app.display(w["program"])
                    ctx.synchronize()
                    t_sv = ctx.pass_time_ms("tile_sv")[0] / 5
                    ctx.enable_timing(False)
                    rec.update({"prism_fragments_per_frame": int(frags), "tile_sv_ms": t_sv,
                                "prism_fragments_per_s": frags / (t_sv * 1e-3) if t_sv > 0 else None,
                                "note": "fragments = pixel centres covered by a volume triangle (before the depth test), counted by the kernel in a separate frame"})
            out[name] = rec
        finally:
            app.close()
    return out


SHARDED_N1_CACHE = os.path.join(ROOT, "gpurun_out", ".sharded_n1.json")


def run_sharded(rank, world, local_rank, steps=24, warmup=4):
    """The north-star multi-GPU partitioning, measured at every N (same keys at N = 1, 2, 4, 8): config c5 on the reference's
    SanDiego geometry (16 lights x 8192^2 depth maps, 7680x4320), ONE frame shared by all ranks.  Rank r owns lights
    l = r (mod N): it renders only those depth maps; the camera pass is reduced to primitive ids, each rank rasterising its own
    screen strip, strips all-gathered over NCCL (4 B/pixel); the accumulation kernel resolves positions from the ids and sums the
    rank's lights; partial sums are reduce-scattered and divided (sgi_reduce_lights), so rank r ends with the final visibility
    of its strip.  Both collectives are issued by the library (C ABI) on its communication stream and overlap the next frame's
    depth passes.  Time = one CUDA-event pair around `steps` frames on every rank + the tail of the last exchange, max over ranks."""
    import torch
    import torch.distributed as dist
    from globalillumination_b200 import capi, hostapi, scenes
    name = "c5_sandiego"
    w = scenes.WORKLOADS[name]
    app = hostapi.App(local_rank)
    app.set_scene(scenes.golden_scene(w["golden"]))
    app.configure(w["W"], w["H"], w["S"])
    app.set_technique(w["technique"])
    app.set(**w["params"])
    app.set(fusedMonteCarlo=1, animationOn=1, animation=-1800.0)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        app.comm_init(uid[0], rank, world)
    ctx = app.context()
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    n_l = w["params"]["numberOfSamples"]
    owners, costs = None, None
    if world > 1:
        # the depth pass of a light costs what the light sees (here up to 1.4x apart): shards are balanced with measured costs
        # (longest first, to the least loaded rank) instead of dealt round-robin; rank 0's measurement decides for everybody
        from globalillumination_b200 import sharding
        box = [[float(c) for c in app.light_costs(n_l)] if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        costs = box[0]
        owners = sharding.balance_lights(costs, world)
        app.set_light_owners(owners)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(n):
                app.display(w["program"]); app.step_animation(ANIMATION_STEP)
            ctx.join()
            e1.record(stream)
            ctx.synchronize()                    # includes the communication stream: the last frame's exchange has landed
        t_host = time.perf_counter()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / n

    with torch.cuda.stream(stream):
        for attempt in range(6):
            try:
                for _ in range(warmup):
                    app.display(w["program"]); app.step_animation(ANIMATION_STEP)
                ctx.synchronize()
                break
            except (capi.SgiError, hostapi.HostError) as e:      # tile lists grown from the first measured frame: warm up again
                if "overflow" not in str(e) or attempt == 5:
                    raise
    ms_frame = timed(steps)
    # what the exchanges cost on the critical path: the same frames with the collectives skipped (camera static: the gathered ids stay valid)
    ms_nocomm = None
    if world > 1:
        app.set(commSkip=1)
        timed(2)
        ms_nocomm = timed(steps)
        app.set(commSkip=0)
        timed(1)
    # rank 0's own pass times (passes one after the other, CUDA events around each)
    ctx.set_option("overlap_passes", 0)
    ctx.enable_timing(True); ctx.reset_timing()
    n_p = 6
    with torch.cuda.stream(stream):
        for _ in range(n_p):
            app.display(w["program"]); app.step_animation(ANIMATION_STEP)
        ctx.synchronize()
    passes = {}
    for pname in capi.PASS:
        ms, n = ctx.pass_time_ms(pname)
        if n:
            passes[pname] = ms / n_p
    ctx.enable_timing(False)
    ctx.set_option("overlap_passes", 1)
    all_passes = [passes]
    if world > 1:
        all_passes = [None] * world
        dist.all_gather_object(all_passes, passes)
    r0, r1 = ctx.comm_strip(rank) if world > 1 else (0, w["H"])
    vis = ctx.read("visibility")[r0:r1]
    lit = float((vis == 1.0).mean())
    xyz, nrm, idx = app.scene_arrays()
    app.close()
    if rank != 0:
        return None
    px = w["W"] * w["H"]
    rec = {
        "workload": name, "mode": "lights" if world > 1 else "single GPU (same code path, no exchange)", "n_gpus": world,
        "W": w["W"], "H": w["H"], "shadow_map": w["S"], "lights": n_l, "lights_per_rank": max(1, n_l // world), "triangles": int(idx.shape[0]),
        "scene": w["scene"], "frames_per_s": 1e3 / ms_frame, "ms_per_frame": ms_frame, "steps": steps, "scaling": "strong",
        "collective": ("ncclAllGather(primitive-id strips, in place, %d MB) + ncclReduceScatter(fp32 partial visibility, in place, %d MB), "
                       "issued by the C ABI (sgi_gather / sgi_reduce_lights); host flag commMasks exchanges lit masks (%d MB) instead"
                       % (px * 4 // 1000000, px * 4 // 1000000, px * ((n_l + 7) // 8) // 1000000)) if world > 1 else None,
        "ms_per_frame_without_exchanges": ms_nocomm, "exposed_comm_ms": (ms_frame - ms_nocomm) if ms_nocomm is not None else 0.0,
        "pass_ms_rank0": passes, "strip_rows_rank0": [r0, r1], "lit_fraction_rank0_strip": lit,
        "tile_depth_ms_per_rank": [round(p.get("tile_depth", 0.0), 4) for p in all_passes],
        "light_owner": owners, "light_cost_ms": [round(c, 4) for c in costs] if costs else None,
    }
    try:
        if world == 1:
            os.makedirs(os.path.dirname(SHARDED_N1_CACHE), exist_ok=True)
            with open(SHARDED_N1_CACHE, "w") as f:
                json.dump({"ms_per_frame": ms_frame}, f)
            rec["speedup_vs_n1"] = 1.0
        elif os.path.exists(SHARDED_N1_CACHE):
            with open(SHARDED_N1_CACHE) as f:
                rec["speedup_vs_n1"] = json.load(f)["ms_per_frame"] / ms_frame
            rec["speedup_vs_n1_source"] = "N=1 record of an earlier run on this box (gpurun_out/.sharded_n1.json)"
        else:
            rec["speedup_vs_n1"] = None
    except Exception:
        rec["speedup_vs_n1"] = None
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_sponza")
    ap.add_argument("--mode", default="frames", choices=["frames", "lights"], help="N>1: frame-parallel, or light shards of one frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the `sharded` record (config c5 light shards, every N)")
    ap.add_argument("--sharded-only", action="store_true", help="only the `sharded` record (its own JSON line): profiling / scaling runs")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` records (short runs of the other configurations, N=1 only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from globalillumination_b200 import scenes
    w = scenes.WORKLOADS[args.workload]
    cfg_path = scenes.write_config(args.workload)

    if args.impl == "reference":
        if rank != 0:
            return 0
        args.steps = args.steps or 3
        args.warmup = 1 if args.warmup is None else args.warmup
        print(json.dumps(run_reference(args, w, cfg_path)), flush=True)
        return 0

    args.steps = args.steps or 2000        # ~0.7 s of timed frames: several 100 ms nvidia-smi clock samples fall inside
    args.warmup = 20 if args.warmup is None else max(3, args.warmup)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the shadow path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from globalillumination_b200 import capi, hostapi
    if args.sharded_only:
        rec = run_sharded(rank, world, local_rank)
        if rank == 0:
            print(json.dumps({"sharded": rec}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0
    app = hostapi.App(local_rank)
    if w.get("golden"):
        app.set_scene(scenes.golden_scene(w["golden"]))
    else:
        app.load_scene(cfg_path)
    app.configure(w["W"], w["H"], w["S"])
    app.set_technique(w["technique"])
    app.set(**w["params"])
    lights_mode = args.mode == "lights" and world > 1 and w["technique"] == "montecarlo"
    if lights_mode:
        app.set(animationOn=1, animation=-1800.0)                          # every rank works on the same frame
        app.set_light_shard(rank, world)
    elif w["program"] == "shadow_volumes":
        # static light: the prism lists of a moving light change size by large factors from frame to frame and the context
        # re-sizes them with an error return (SGI_ERR_OVERFLOW, "run the frame again"), which a timed loop cannot absorb
        app.set(animationOn=0)
    else:
        app.set(animationOn=1, animation=-1800.0 + ANIMATION_STEP * rank)     # rank r renders frames r, r+N, ...
    anim_stride = ANIMATION_STEP * (1 if lights_mode else world)
    program = w["program"]
    result_buf = "sv_count" if program == "shadow_volumes" else "visibility"
    ctx = app.context()
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    xyz, nrm, idx = app.scene_arrays()
    V, T = xyz.shape[0], idx.shape[0]
    n_l = w["params"].get("numberOfSamples", 1) if w["technique"] == "montecarlo" else 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class _DevView:      # torch tensor over the context's visibility buffer (the NCCL send buffer: no staging copy)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    vis_t = None
    staging, vis_strip, ev_ready, ev_reduced = [None, None], [None, None], None, None
    comm_stream = torch.cuda.Stream(device=local_rank) if lights_mode else None
    frame_no = 0

    def frame():
        """One frame.  Light shards: the partial sums leave the context's buffer through a device copy into one of two
        staging buffers and are reduce-scattered on a side stream, so the collective of frame k runs under the depth and
        G-buffer passes of frame k+1 (a strip of the final image per rank is what tile-local shading consumes).
        One NCCL collective per frame; the timed region ends only when the last collective has completed."""
        nonlocal vis_t, ev_ready, ev_reduced, frame_no
        app.display(program)
        ctx.join()            # the shadow pass runs on an internal stream: order `stream` (the timing events) after it
        if lights_mode:
            if vis_t is None:
                ptr, nbytes = ctx.device_ptr("visibility")
                vis_t = torch.as_tensor(_DevView(ptr, nbytes // 4), device=f"cuda:{local_rank}")
                assert vis_t.numel() % world == 0
                for b in range(2):
                    staging[b] = torch.empty_like(vis_t)
                    vis_strip[b] = torch.empty(vis_t.numel() // world, dtype=torch.float32, device=f"cuda:{local_rank}")
                ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
                ev_reduced = [torch.cuda.Event(), torch.cuda.Event()]
            b = frame_no & 1
            if frame_no >= 2:
                stream.wait_event(ev_reduced[b])                     # the collective that last read this staging buffer
            staging[b].copy_(vis_t)
            ev_ready[b].record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev_ready[b])
                dist.reduce_scatter_tensor(vis_strip[b], staging[b])
                vis_strip[b].mul_(1.0 / n_l)                         # AccurateSoftShadow.frag:127
                ev_reduced[b].record(comm_stream)
            frame_no += 1

    def join_comm():
        if lights_mode and ev_reduced is not None:
            for e in ev_reduced:
                stream.wait_event(e)

    app.upload_scene()
    with torch.cuda.stream(stream):
        for attempt in range(6):
            try:
                for _ in range(args.warmup):
                    frame(); app.step_animation(anim_stride)
                join_comm()
                ctx.synchronize()
                break
            except capi.SgiError as e:            # tile lists grown from the first measured frame: run the warm-up again
                if "overflow" not in str(e) or attempt == 5:
                    raise

        # ---- timed region: K steps, per-step events, L2 flushed before each ----
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        launches0 = ctx.kernel_launches()
        barrier()
        ev_all = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev_all[0].record(stream)
        for k in range(args.steps):
            if not lights_mode:
                flush.zero_()
            ev[k][0].record(stream)
            frame()
            ev[k][1].record(stream)
            app.step_animation(anim_stride)
        join_comm()
        ev_all[1].record(stream)
        ctx.synchronize()
        barrier()
        launches = ctx.kernel_launches() - launches0
        # nvidia-smi samples every 100 ms: a short timed region (small --steps) may hold fewer than three samples.  Keep the
        # same load running (untimed, same frames) until three are in, so the clocks line still describes the GPU under
        # this workload; "extended_s" says how long that took (0 when the timed region alone was long enough).
        extended_s = 0.0
        if rank == 0 and clocks.proc is not None and not lights_mode:
            t_ext = time.perf_counter()
            while len(clocks.rows) < 3 and time.perf_counter() - t_ext < 1.0:
                for _ in range(20):
                    frame(); app.step_animation(anim_stride)
                ctx.synchronize()
            extended_s = time.perf_counter() - t_ext if len(clocks.rows) and time.perf_counter() - t_ext > 0.01 else 0.0
        clock_info = clocks.stop() if rank == 0 else None
        if clock_info is not None:
            clock_info["extended_s"] = round(extended_s, 3)
        step_ms = [a.elapsed_time(b) for a, b in ev]
        # light shards: frames overlap their predecessors' collective, so the whole loop is timed with one event pair
        # (no L2 flush inside it: the 16 depth maps + G-buffer of this workload are far larger than L2 anyway)
        total_ms = float(ev_all[0].elapsed_time(ev_all[1])) if lights_mode else float(sum(step_ms))

        # ---- per-kernel shares (same steps again, CUDA events around the kernels of each pass) ----
        # (pass overlap off here: with the depth and G-buffer passes on concurrent streams an event pair around one kernel
        #  also counts the time it shares the SMs with the other stream's kernels; alone on the stream it is the kernel's own
        #  duration, which is what the roofline and the ncu launch list describe)
        ctx.set_option("overlap_passes", 0)
        ctx.enable_timing(True); ctx.reset_timing()
        for k in range(min(args.steps, 100)):
            flush.zero_()
            frame(); app.step_animation(anim_stride)
        join_comm()
        ctx.synchronize()
        ctx.set_option("overlap_passes", 1)
        passes = {}
        for name in capi.PASS:
            ms, n = ctx.pass_time_ms(name)
            if n:
                passes[name] = ms / n * (n / min(args.steps, 100))       # ms per frame
        ctx.enable_timing(False)
        sv_frags = None
        if program == "shadow_volumes":
            ctx.set_option("sv_count_fragments", 1)
            frame(); ctx.synchronize()
            sv_frags = int(ctx.sv_fragments())
            ctx.set_option("sv_count_fragments", 0)

        # ---- secondary figure, PCSS only: the same loop with the exact early-out of the blocker search switched on
        #      (sgi_set_option "pcss_early_out": bit-identical results, off by default so that `value` counts every tap) ----
        early = None
        if w["technique"] == "pcss" and not lights_mode:
            ctx.set_option("pcss_early_out", 1)
            n_e = min(args.steps, 500)
            for k in range(5):
                frame(); app.step_animation(anim_stride)
            ctx.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_e)]
            for k in range(n_e):                   # timed exactly like the headline loop: L2 flush outside each event pair
                flush.zero_()
                evs[k][0].record(stream); frame(); evs[k][1].record(stream)
                app.step_animation(anim_stride)
            ctx.synchronize()
            ms_e = sum(e0.elapsed_time(e1) for e0, e1 in evs)
            ctx.set_option("pcss_early_out", 0)
            early = {"value": world * n_e / (ms_e / 1e3), "unit": "frames/s", "steps": n_e,
                     "note": "option pcss_early_out = 1: pixels with light-space depth in (0, 0.989) return 1.0 without taps (provably the program's result, bit-identical); not the headline"}

        # ---- e2e: host buffers in, host buffer out, every frame ----
        vis_bytes = w["W"] * w["H"] * 4
        E2E_DEPTH = 3            # frames in flight: the host queues frame k while k-1 renders and k-2 is copied out
        host_vis2 = [torch.empty(w["W"] * w["H"], dtype=torch.float32).pin_memory() for _ in range(E2E_DEPTH)]
        host_vis = host_vis2[0]
        e2e_steps = max(10, min(args.steps, 200))

        def e2e_loop(n):
            """Every frame: geometry host->device (the Mesh arrays), all passes, visibility device->host (pinned).
            Frames are pipelined: frame k's copy-out overlaps the passes of the frames queued behind it; all of it inside
            the timed region, which ends when the last frame's pixels are in host memory."""
            pending = []
            for k in range(n):
                pending.append(app.display_e2e_async(program, result_buf, host_vis2[k % E2E_DEPTH].data_ptr(), vis_bytes))
                app.step_animation(anim_stride)
                if len(pending) >= E2E_DEPTH:
                    app.e2e_wait(pending.pop(0))
            for t in pending:
                app.e2e_wait(t)

        e2e_loop(4)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        # the blocking form (upload, passes, read back, return) for reference
        t0 = time.perf_counter()
        for _ in range(20):
            app.display_e2e(program, result_buf, host_vis.data_ptr(), vis_bytes); app.step_animation(anim_stride)
        e2e_blocking_s = (time.perf_counter() - t0) / 20
        lit = None
        if result_buf == "visibility" and not lights_mode:        # a fixed frame (the first of the animation), not whichever the loops ended on
            app.set(animation=-1800.0)
            app.display_e2e(program, result_buf, host_vis.data_ptr(), vis_bytes)
            lit = float((host_vis == 1.0).float().mean())

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    sharded = None
    if not args.no_sharded and args.workload == "c2_sponza" and not lights_mode:
        app.close()                                   # free the headline context's targets first
        sharded = run_sharded(rank, world, local_rank)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fps = (1 if lights_mode else world) * args.steps / (total_ms / 1e3)
    peak, peak_src = load_peaks()
    # light shards: a rank's kernels process only its own lights (pass_ms are rank 0's), so the per-kernel bytes count those
    n_l_rank = max(1, n_l // world) if lights_mode else n_l
    ab = algorithmic_bytes(w, V, T, n_l_rank)
    kernel_of = {"tile_sv": ("k_tile<SVCOUNT> (shadow-volume prism counting)", ab["shadow_volume"]),
                 "vis_kernel": ("k_visibility (per-pixel shadow test/filter)", ab["visibility"]),
                 "tile_depth": ("k_tile<DEPTH> (light-view tile rasteriser)", ab["shadow_map"] // max(1, n_l_rank)),
                 "tile_gbuffer": ("k_tile<GBUFFER> (camera-view tile rasteriser + resolve)", ab["gbuffer"])}
    if "moment_filter" in ab:
        kernel_of["moment_filter"] = ("k_mom_filter (separable blur of the moment map, X + Y launches)", ab["moment_filter"] // 2)
        kernel_of["tile_depth"] = ("k_tile<MOMENTS> (light-view tile rasteriser + moment resolve)", ab["shadow_map"])
        kernel_of["vis_kernel"] = ("k_mom_visibility (moment reconstruction)", ab["visibility"])
    cand = {k: v for k, v in passes.items() if k in kernel_of}
    roof = None
    if cand:
        top = max(cand, key=cand.get)
        calls = n_l_rank if top == "tile_depth" else (2 if top == "moment_filter" else 1)
        per_launch_ms = cand[top] / calls
        achieved = kernel_of[top][1] / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "dram_traffic_c2.json")
        if args.workload == "c2_sponza" and os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            key = {"vis_kernel": "k_visibility", "tile_depth": "k_tile<DEPTH>", "tile_gbuffer": "k_tile<GBUFFER>"}[top]
            if key in tj["kernels"]:
                traffic = tj["kernels"][key]["dram_bytes_read"] + tj["kernels"][key]["dram_bytes_write"]
                traffic_src = tj["source"]
        roof = {"bound": "hbm", "kernel": kernel_of[top][0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": kernel_of[top][1], "launch_ms": per_launch_ms, "peak_source": peak_src,
                "share_of_step": cand[top] / (total_ms / args.steps)}
    if roof and program == "shadow_volumes":
        # the unit of work of the stencil pass is the prism fragment, not the byte: the HBM fraction of this kernel says
        # nothing (it moves 8 B/pixel); fragments are tallied by the kernel itself in one extra frame (option sv_count_fragments)
        roof["note"] = "fill-bound pass: see prism_fragments_per_s; the HBM fraction is not a measure of this kernel"
        roof["prism_fragments_per_frame"] = sv_frags
        roof["prism_fragments_per_s"] = sv_frags / (per_launch_ms * 1e-3) if sv_frags else None
    # the shadow pass against both of its rooflines - HBM (G-buffer in, visibility out, every map texel once)
    # and L2 (4 bytes per shadow-map tap; measured L2 read peak: profiles/r1_l2_bandwidth.txt)
    shadow_pass = None
    if "vis_kernel" in passes and passes["vis_kernel"] > 0:
        px = w["W"] * w["H"]
        k = w["params"].get("kernelSize", 15)
        bs = w["params"].get("blockerSearchSize", 7)
        ko = w["params"].get("kernelOrder", 7)
        taps = {"pcss": bs * bs + k * k, "montecarlo": n_l, "pcf": ko * ko}.get(w["technique"], 1)
        t_s = passes["vis_kernel"] * 1e-3
        shadow_pass = {"launch_ms": passes["vis_kernel"], "hbm_GBs": ab["visibility"] / t_s / 1e9, "hbm_frac": ab["visibility"] / t_s / 1e9 / peak,
                       "taps_per_lit_pixel": taps, "l2_taps_GBs_upper": 4.0 * taps * px / t_s / 1e9, "l2_read_peak_GBs": 9555.0,
                       "note": "l2_taps is an upper bound (every foreground pixel taking every tap); PCSS pixels without blockers stop after the blocker search"}
    med = float(np.median(step_ms)) if step_ms else None
    out = {
        "metric": metric_name(args.workload, w), "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "ms_per_step_median": med, "higher_is_better": True, "scaling": "strong" if lights_mode else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(args.workload, w, V, T),
        "run": {"l2": ("not flushed: per-frame inputs (depth maps + G-buffer) exceed L2; whole loop timed with one event pair" if lights_mode else
                       "flushed before every timed step (256 MiB memset on the same stream, outside the event pair)"),
                "parallelism": (f"lights x{world} + reduce-scatter (pipelined one frame deep on a side stream)" if lights_mode else f"frames x{world}") if world > 1 else "single GPU",
                "lit_fraction": lit, "lit_fraction_frame": "animation = -1800 (the first frame of the sequence), rendered after the timed loops"},
        "clocks": clock_info, "gpu_launches": int(launches),
        "e2e": {"value": (1 if lights_mode else world) * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(V * (36 if any(l.startswith("c ") for l in w["lines"]) else 24) + T * 12),   # xyz + normals (+ colours) + indices
                "d2h_bytes_per_step": int(vis_bytes), "steps": e2e_steps, "pipelined_frames_in_flight": 3,
                "blocking_call_ms": 1e3 * e2e_blocking_s},
        "pass_ms": passes, "roofline": roof, "shadow_pass": shadow_pass,
    }
    if early:
        out["with_pcss_early_out"] = early
    if sharded:
        out["sharded"] = sharded
    if world == 1 and args.workload == "c2_sponza" and not args.no_secondary:
        out["secondary"] = run_secondary(local_rank)
    if not args.no_cpu_baseline and world == 1:
        a2 = argparse.Namespace(**vars(args)); a2.steps, a2.warmup = 3, 1
        ref = run_reference(a2, w, cfg_path)
        out["cpu_baseline"] = ref["cpu_baseline"]
        if ref.get("pcss_taps") and out.get("shadow_pass"):
            # L2_taps / t_K3 with the taps the program really executes (4 bytes each), against the measured L2 read peak
            sp, tp = out["shadow_pass"], ref["pcss_taps"]
            sp["taps_executed"] = tp
            sp["l2_taps_GBs"] = 4.0 * tp["taps"] / (sp["launch_ms"] * 1e-3) / 1e9
            sp["l2_taps_frac"] = sp["l2_taps_GBs"] / sp["l2_read_peak_GBs"]
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
