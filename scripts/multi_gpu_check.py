"""Multi-GPU parity check on real GPUs (run under torchrun, NCCL):
  tiles  — every rank renders the full shadow map, its own 64-row-aligned strip of the G-buffer and of the visibility
           (sgi_params.rect_*); strips are all-gathered; result must equal rank 0's un-sharded frame bit for bit
  lights — every rank owns lights l = rank (mod N) of a 16-light frame; partial sums are all-reduced; result must equal
           the un-sharded 16-light frame bit for bit
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py
"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from globalillumination_b200 import capi, hostapi, scenes, sharding


class DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # ---- tiles, PCF on the c2 scene at 1920x1080
    w = scenes.WORKLOADS["c2_sponza"]
    W, H, S = w["W"], w["H"], w["S"]
    app = hostapi.App(local)
    app.load_scene(scenes.write_config("c2_sponza")); app.configure(W, H, S); app.set_technique("pcf")
    app.display("shadow_mapping")
    full = torch.from_numpy(app.context().read("visibility")).cuda()
    rects = sharding.strip_rects(W, H, world)
    app.set_rect(*rects[rank])
    app.display("shadow_mapping")
    ctx = app.context()
    ptr, nbytes = ctx.device_ptr("visibility")
    ctx.synchronize()
    vis = torch.as_tensor(DevView(ptr, nbytes // 4), device=f"cuda:{local}").view(H, W)
    out = sharding.gather_strips(vis, rects)
    same = bool(torch.equal(out, full))
    ok &= same
    if rank == 0:
        print(f"tiles x{world}: gathered image == un-sharded image: {same} (strips {rects})", flush=True)
    # ---- tiles with a pre-filtered technique (VSM): the moment target and its blur are replicated, each rank reconstructs its strip
    app.set_rect(0, 0, 0, 0); app.set_technique("vsm")
    app.display("shadow_mapping")
    full = torch.from_numpy(ctx.read("visibility")).cuda()
    app.set_rect(*rects[rank])
    app.display("shadow_mapping")
    ptr, nbytes = ctx.device_ptr("visibility")
    ctx.synchronize()
    vis = torch.as_tensor(DevView(ptr, nbytes // 4), device=f"cuda:{local}").view(H, W)
    out = sharding.gather_strips(vis, rects)
    same = bool(torch.equal(out, full))
    ok &= same
    if rank == 0:
        print(f"tiles x{world}, VSM: gathered image == un-sharded image: {same}", flush=True)
    app.close()
    # ---- lights, 16 lights, 2048x1152, 1024^2 maps
    app = hostapi.App(local)
    app.load_scene(scenes.write_config("c5_many_light")); app.configure(2048, 1152, 1024); app.set_technique("montecarlo")
    app.set(numberOfSamples=16, lightSourceSize=16)
    app.display("soft_shadow_mapping")
    full = torch.from_numpy(app.context().read("visibility")).cuda()
    app.set_light_shard(rank, world)
    app.display("soft_shadow_mapping")
    ctx = app.context()
    ptr, nbytes = ctx.device_ptr("visibility")
    ctx.synchronize()
    part = torch.as_tensor(DevView(ptr, nbytes // 4), device=f"cuda:{local}").clone()
    dist.all_reduce(part)
    total = (part / 16.0).view(1152, 2048)
    same = bool(torch.equal(total, full))
    ok &= same
    if rank == 0:
        print(f"lights x{world}: all-reduced partial sums / 16 == un-sharded 16-light frame: {same}", flush=True)
    app.close()

    # ---- the same two partitionings with the exchange inside the C ABI (sgi_comm_init / sgi_gather / sgi_reduce_lights)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    # tiles: PCF on the c2 scene, strips = sgi_comm_strip, one in-place ncclAllGather
    app = hostapi.App(local)
    app.load_scene(scenes.write_config("c2_sponza")); app.configure(W, H, S); app.set_technique("pcf")
    app.display("shadow_mapping")
    full = app.context().read("visibility")
    app.comm_init(uid[0], rank, world)
    app.display("shadow_mapping")                       # (sizes the padded targets for the rank count)
    ctx = app.context()
    r0, r1 = ctx.comm_strip(rank)
    app.set_rect(0, r0, W, r1)
    app.display("shadow_mapping")
    ctx.gather("visibility")
    same = bool(np.array_equal(ctx.read("visibility").view(np.uint32), full.view(np.uint32)))
    ok &= same
    if rank == 0:
        print(f"C ABI tiles x{world}: sgi_gather(visibility) == un-sharded image: {same} (rows {r0}..{r1} on rank 0)", flush=True)
    app.close()
    # lights: 16 lights, fused many-light path (primitive-id strips all-gathered, partial sums reduce-scattered + divided)
    # (exchange of lit masks - host flag commMasks, up to 32 lights - also with a shadow intensity that is not a dyadic fraction, where only
    #  the masks reproduce the un-sharded accumulation order; and of float partial sums)
    for (Wl, Hl, masks, si) in ((2048, 1152, 1, None), (1000, 563, 1, 0.3), (1000, 563, 0, None)):   # 563 rows do not divide by the rank count: padded strips
        app = hostapi.App(local)
        app.load_scene(scenes.write_config("c5_many_light")); app.configure(Wl, Hl, 1024); app.set_technique("montecarlo")
        app.set(numberOfSamples=16, lightSourceSize=16, commMasks=masks)
        if si is not None:
            app.set(shadowIntensity=si)
        app.display("soft_shadow_mapping")
        full = app.context().read("visibility")
        app.set(fusedMonteCarlo=1)
        uid2 = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        app.comm_init(uid2[0], rank, world)
        if Hl == 563:                                    # a light-to-rank table other than round robin (shards balanced by cost)
            app.set_light_owners([min(world - 1, s * world // 12) for s in range(16)])      # (unequal shares)
        for _ in range(3):                               # several frames: the exchanges of consecutive frames overlap
            app.display("soft_shadow_mapping")
        ctx = app.context()
        r0, r1 = ctx.comm_strip(rank)
        mine = ctx.read("visibility")[r0:r1]
        same = bool(np.array_equal(mine.view(np.uint32), full[r0:r1].view(np.uint32)))
        ok &= same
        print(f"C ABI lights x{world} {Wl}x{Hl} {'masks' if masks else 'float sums'}{'' if si is None else ' si=%g' % si}, rank {rank}: rows {r0}..{r1} of sgi_reduce_lights == un-sharded frame: {same}", flush=True)
        app.close()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
