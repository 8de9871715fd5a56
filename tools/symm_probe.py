"""Probe: does torch's symmetric memory give this rank store access to its peers' buffers on this box?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
t.fill_(-1.0)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, flush=True)
hdl.barrier()
for p in range(world):
    hdl.get_buffer(p, (1024,), torch.float32)[rank * 8:(rank + 1) * 8] = float(rank + 1)
hdl.barrier()
torch.cuda.synchronize()
print(rank, "after peer writes:", t[:8 * world:8].tolist(), flush=True)
dist.destroy_process_group()
