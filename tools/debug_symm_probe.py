"""Debug tool: does torch symmetric memory (peer-mapped buffers) work on this box?  torchrun --nproc-per-node 2."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
print(rank, 'ptrs', [hex(p) for p in hdl.buffer_ptrs], flush=True)
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1 << 20,), torch.float32)
print(rank, 'peer value', float(pt[12345]), 'expected', float(peer + 1), flush=True)
# bandwidth of a plain peer read
dst = torch.empty(54 << 18, dtype=torch.float32, device=dev)   # 54 MB
big = symm_mem.empty(54 << 18, dtype=torch.float32, device=dev)
hb = symm_mem.rendezvous(big, dist.group.WORLD)
pb = hb.get_buffer(peer, (54 << 18,), torch.float32)
for _ in range(3):
    dst.copy_(pb)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    dst.copy_(pb)
e.record(); torch.cuda.synchronize()
print(rank, 'peer copy 54 MB: %.3f ms -> %.0f GB/s' % (s.elapsed_time(e) / 10, 54 * 1.048576 / (s.elapsed_time(e) / 10)), flush=True)
dist.barrier()
dist.destroy_process_group()
