"""Does torch symmetric memory work on this box?  (developer probe, run under torchrun with 2+ GPUs)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((4, 8), dtype=torch.float32, device=dev)
t.fill_(float(rank))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (4, 8), torch.float32)
peer[rank].fill_(100.0 + rank)          # remote store into the next rank's table
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "local table after peers wrote:", t[:, 0].tolist(), flush=True)
dist.destroy_process_group()
