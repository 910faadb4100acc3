import os, time, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
for mb in (1, 16, 54, 256):
    x = torch.empty(mb * 1024 * 1024 // 4, device=dev)
    out = torch.empty(world * x.numel(), device=dev)
    for name, fn in (('allgather', lambda: dist.all_gather_into_tensor(out, x)), ('allreduce', lambda: dist.all_reduce(x))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        if rank == 0: print('%s %4d MB: %.3f ms  (%.1f GB/s per rank payload)' % (name, mb, dt * 1e3, mb / 1024 / dt), flush=True)
dist.destroy_process_group()
