"""Host I/O probe behind the end-to-end numbers: pinned D2H of a frame's visibility (8.3 MB) and H2D of its geometry (2.2 MB),
alone and with every rank of the job doing the same at once (run under torchrun: ranks start together after a barrier).
Prints per-rank and aggregate GB/s; rank 0 prints the summary line."""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = torch.empty(8294400 // 4, dtype=torch.float32, device="cuda"); h = torch.empty(8294400 // 4, dtype=torch.float32).pin_memory()
h2 = torch.empty(2162496 // 4, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(h2, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(n, both):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


for both in (False, True):
    run(20, both)
    dt = run(300, both)
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        b = 8.2944e6 + (2.162496e6 if both else 0)
        print(f"ranks {world}: {'D2H 8.3 MB + H2D 2.2 MB' if both else 'D2H 8.3 MB'} per frame, all ranks at once: {float(t[0]) * 1e3:.3f} ms per frame on the slowest rank"
              f" -> {b / float(t[0]) / 1e9:.1f} GB/s per rank, {world * b / float(t[0]) / 1e9:.1f} GB/s aggregate, {world / float(t[0]):.0f} frames/s bound", flush=True)
if world > 1:
    dist.destroy_process_group()
