"""Multi-process host logic of the multi-GPU modes on CPU: world_size 2, gloo.  The per-rank
compute is stood in for by the oracle so that the partition + collective + reassembly logic is what is tested."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from globalillumination_b200 import sharding


def test_strip_rects_cover_the_screen_exactly():
    for W, H, n in ((1920, 1080, 8), (7680, 4320, 8), (640, 480, 3), (100, 37, 5), (64, 64, 1)):
        rects = sharding.strip_rects(W, H, n)
        assert len(rects) == n and rects[0][1] == 0 and rects[-1][3] == H
        for a, b in zip(rects, rects[1:]):
            assert a[3] == b[1] and a[0] == 0 and a[2] == W
        if H >= 64 * n:
            assert all(r[1] % 64 == 0 for r in rects)


def test_light_and_frame_shards_partition():
    for n, w in ((16, 8), (289, 8), (5, 2), (3, 4)):
        got = sorted(i for r in range(w) for i in sharding.light_shard(n, r, w))
        assert got == list(range(n))
    assert sharding.frame_indices(10, 7, 1, 3) == [11, 14]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as O
        from tests import util
        O.set_num_threads(2)
        sc = util.scene("door")
        W, H, S = 96, 130, 64
        fm = util.frame(sc, W, H, S)
        sm = O.raster_depth(sc["xyz"], sc["idx"], fm["light_mvp"], S, S)
        pos, nrm, _ = O.raster_gbuffer(sc["xyz"], sc["nrm"], sc["idx"], fm["cam_mvp"], W, H)
        cam = O.make_camera(fm["cam_mv"], fm["normal_matrix"], fm["light_pos_shading"])
        full = O.visibility(O.default_params("pcf", S), cam, fm["light_mvp_b"], pos, nrm, sm)
        # tiles: each rank evaluates only its strip, strips are all-gathered
        rects = sharding.strip_rects(W, H, world)
        x0, y0, x1, y1 = rects[rank]
        p = O.default_params("pcf", S, rect_x0=x0, rect_y0=y0, rect_x1=x1, rect_y1=y1)
        mine = O.visibility(p, cam, fm["light_mvp_b"], pos, nrm, sm)
        assert (mine[:y0] == 0).all() and (mine[y1:] == 0).all()
        out = sharding.gather_strips(torch.from_numpy(mine), rects).numpy()
        ok_tiles = np.array_equal(out, full)
        # tiles with a moment shadow map (VSM): the moment target and its blur are replicated like the depth map, every rank
        # reconstructs only its strip
        fmap = O.filter_shadow_map(O.raster_moments(sc["xyz"], sc["idx"], fm["light_mvp"], S, S, "vsm"), W, H, 7, "vsm")
        full_m = O.visibility_moments(O.default_params("vsm", S), cam, fm["light_mvp_b"], pos, nrm, fmap)
        mine_m = O.visibility_moments(O.default_params("vsm", S, rect_x0=x0, rect_y0=y0, rect_x1=x1, rect_y1=y1), cam, fm["light_mvp_b"], pos, nrm, fmap)
        out_m = sharding.gather_strips(torch.from_numpy(mine_m), rects).numpy()
        ok_tiles = ok_tiles and np.array_equal(out_m, full_m) and (mine_m[:y0] == 0).all() and (mine_m[y1:] == 0).all()
        # lights: each rank accumulates its own lights, partial sums are reduced
        n_l = 4
        mvp, mvpb = util.multi_lights(sc, n_l, 16, W, H, S)
        maps = np.stack([O.raster_depth(sc["xyz"], sc["idx"], mvp[i], S, S) for i in range(n_l)])
        pm = O.default_params("multi_hard", S)
        ref = O.visibility_multi(pm, mvpb[-1], mvpb[:, 12:16], pos, maps)
        own = sharding.light_shard(n_l, rank, world)
        part = O.visibility_multi(pm, mvpb[-1], mvpb[own, 12:16], pos, maps[own]) * len(own)     # sum over own lights
        fg = torch.from_numpy((pos[..., 0] != 0).astype(np.float32))
        tot = sharding.reduce_light_partials(torch.from_numpy(part), fg * len(own)).numpy()
        ok_lights = np.allclose(tot, ref, atol=1e-6)
        q.put((rank, ok_tiles, ok_lights))
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_gather_and_light_reduce_gloo():
    world, port = 2, 29500 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "tile gather mismatch"
    assert all(r[2] for r in res), "light reduce mismatch"


def test_balance_lights_is_deterministic_and_beats_round_robin():
    from globalillumination_b200 import sharding
    costs = [0.57, 0.41, 0.40, 0.52, 0.55, 0.42, 0.39, 0.50, 0.58, 0.43, 0.41, 0.51, 0.56, 0.40, 0.38, 0.49]
    own = sharding.balance_lights(costs, 8)
    assert own == sharding.balance_lights(list(costs), 8) and sorted(set(own)) == list(range(8))
    load = lambda o: max(sum(c for c, r in zip(costs, o) if r == k) for k in range(8))
    assert load(own) < load([s % 8 for s in range(16)])
    assert sharding.balance_lights([1.0], 4) == [0]
