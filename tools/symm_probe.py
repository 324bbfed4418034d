"""Probe: does torch symmetric memory (peer-mapped buffers + stream-ordered signals) work on this box?
   torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
shape = (2, 1, 4, 25, 72, 128)
t = symm.empty(shape, dtype=torch.float16, device=dev)
t.zero_()
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, [n for n in dir(hdl) if not n.startswith("_")][:30], flush=True)
dist.barrier()
nxt, prv = (rank + 1) % world, (rank - 1) % world
peer = hdl.get_buffer(nxt, shape, torch.float16)
src = torch.full(shape[1:], float(rank + 1), device=dev, dtype=torch.float16)
for it in range(4):                                # warm-up
    slot = it % 2
    peer[slot].copy_(src)
    hdl.put_signal(nxt, channel=slot)
    hdl.wait_signal(prv, channel=slot)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for it in range(40):
    slot = it % 2
    peer[slot].copy_(src)                      # P2P store into the next rank's slot
    hdl.put_signal(nxt, channel=slot)          # stream-ordered: after the copy
    hdl.wait_signal(prv, channel=slot)         # stream-ordered: my slot was filled by the previous rank
e1.record()
host = (time.perf_counter() - t0) / 40
torch.cuda.synchronize()
got = t[1].float().mean()
print(rank, "value from prev rank:", float(got), "expected", float(prv + 1),
      f"device {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per hop, host enqueue {host * 1e6:.1f} us", flush=True)
# NCCL for comparison
buf = torch.empty(shape[1:], device=dev, dtype=torch.float16)
for _ in range(4):
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, src, nxt), dist.P2POp(dist.irecv, buf, prv)]):
        w.wait()
torch.cuda.synchronize()
e0.record()
for _ in range(40):
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, src, nxt), dist.P2POp(dist.irecv, buf, prv)]):
        w.wait()
e1.record()
torch.cuda.synchronize()
print(rank, f"NCCL batch_isend_irecv: device {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per hop", flush=True)
dist.barrier()
dist.destroy_process_group()
