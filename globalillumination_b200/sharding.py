"""Multi-GPU partitioning of one frame.  One process per GPU; geometry is replicated.

  frames   every rank renders whole frames of its own (rank r: frames r, r+N, ...) — no data-path collective;
           this is what bench.py measures at N > 1 ("scaling": "weak").
  tiles    single/few lights: every rank builds the (cheap) shadow map, rasterises and shades only its
           horizontal strip of the screen (sgi_params.rect_*); strips are gathered to every rank.
  lights   many lights (config c5): rank r owns lights l = r (mod N): it builds only those depth maps and
           accumulates their hard visibility over the whole screen; partial sums are reduced.
Both collectives are a single call per frame on the visibility image (W*H floats).
"""
import numpy as np


def strip_rects(W, H, n):
    """n horizontal strips covering [0,H): rows are split as evenly as possible, 64-row aligned where the
    screen is tall enough (the rasteriser works on 64x64 tiles, so aligned strips do no redundant tile work)."""
    if n <= 0:
        raise ValueError("n must be positive")
    align = 64 if H >= 64 * n else 1
    units = (H + align - 1) // align
    rects, y = [], 0
    for r in range(n):
        rows = (units // n + (1 if r < units % n else 0)) * align
        y1 = min(H, y + rows)
        rects.append((0, y, W, y1))
        y = y1
    assert y == H or rects[-1][3] == H
    return rects


def light_shard(num_lights, rank, world):
    """Indices of the lights rank `rank` owns (round robin)."""
    return list(range(rank, num_lights, world))


def balance_lights(costs, world):
    """Owner rank of every light, longest-processing-time first: lights sorted by measured cost, each given to the rank with the least
    load so far (ties: lowest rank).  Deterministic for a given cost vector; every rank must use the same table."""
    order = sorted(range(len(costs)), key=lambda s: (-float(costs[s]), s))
    load = [0.0] * world
    count = [0] * world
    owner = [0] * len(costs)
    for s in order:
        r = min(range(world), key=lambda k: (load[k], count[k], k))
        owner[s] = r
        load[r] += float(costs[s]); count[r] += 1
    return owner


def frame_indices(first, count, rank, world):
    """Frames rank `rank` renders in frame-parallel mode."""
    return list(range(first + rank, first + count, world))


def gather_strips(local_vis, rects, group=None):
    """All-gather the per-rank strips into the full image (works on CPU tensors with gloo and CUDA tensors with NCCL).
    local_vis: full-size [H,W] tensor of which only this rank's strip is valid."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    H, W = local_vis.shape
    rows = max(r[3] - r[1] for r in rects)
    send = torch.zeros((rows, W), dtype=local_vis.dtype, device=local_vis.device)
    x0, y0, x1, y1 = rects[rank]
    send[: y1 - y0] = local_vis[y0:y1]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out = torch.empty_like(local_vis)
    for r, (x0, y0, x1, y1) in enumerate(rects):
        out[y0:y1] = recv[r][: y1 - y0]
    return out


def reduce_light_partials(partial_sum, weight_sum, group=None):
    """Sum the per-rank (sum_l w_l*vis_l, sum_l w_l) images and normalise (AccurateSoftShadow.frag:127)."""
    import torch.distributed as dist
    dist.all_reduce(partial_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(weight_sum, op=dist.ReduceOp.SUM, group=group)
    return partial_sum / weight_sum.clamp_min(1e-30)
