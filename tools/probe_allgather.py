"""Under torchrun: what one frame's geometry costs to distribute — 1/N of 28 MB from pinned memory per rank, then an NCCL
all-gather of the 28 MB — each alone, CUDA events, max over ranks."""
import json, os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = 28048096
per = ((total + world - 1) // world + 15) // 16 * 16
full = torch.zeros(per * world, dtype=torch.uint8, device="cuda")
host = torch.zeros(per, dtype=torch.uint8).pin_memory()
mine = full[rank * per:(rank + 1) * per]
band = torch.zeros(33177600 // world, dtype=torch.uint8, device="cuda")
band_host = torch.zeros(33177600 // world, dtype=torch.uint8).pin_memory()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


out = {"world": world, "shard_bytes": per}
out["h2d_ms"] = timed(lambda: mine.copy_(host, non_blocking=True))
out["allgather_ms"] = timed(lambda: dist.all_gather_into_tensor(full, mine))
out["h2d_then_allgather_ms"] = timed(lambda: (mine.copy_(host, non_blocking=True), dist.all_gather_into_tensor(full, mine)))
out["d2h_band_ms"] = timed(lambda: band_host.copy_(band, non_blocking=True))
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
