"""Does torch symmetric memory (peer-mapped buffers over NVLink) work on this box? torchrun --nproc-per-node 2."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
h = symm_mem.rendezvous(t, dist.group.WORLD)
ptrs = [int(p) for p in h.buffer_ptrs]
t.fill_(rank + 1)
torch.cuda.synchronize(); h.barrier(); 
peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
print(rank, "ptrs", [hex(p) for p in ptrs], "local", hex(t.data_ptr()), "peer first byte", int(peer[0]), flush=True)
h.barrier()
dist.destroy_process_group()
