"""Probe torch symmetric memory on this box (development aid): rendezvous, peer pointers, barrier cost."""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty((512 * 32,), dtype=torch.float64, device=torch.device('cuda', local))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, 'rendezvous ok; buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs], 'dev array', hex(hdl.buffer_ptrs_dev),
          'signal pads', hex(hdl.signal_pad_ptrs_dev), 'multicast', hex(hdl.multicast_ptr) if hdl.has_multicast_support else None, flush=True)
    t.fill_(rank + 1.0)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (512 * 32,), torch.float64)
    print(rank, 'peer value', float(peer[0]), flush=True)
    torch.cuda.synchronize()
    for name, fn in (('symm barrier', lambda: hdl.barrier(channel=0)),
                     ('nccl all_reduce 131KB', lambda: dist.all_reduce(t)),
                     ('nccl all_gather 8.5KB', None)):
        if fn is None:
            src = torch.zeros(1057, dtype=torch.float64, device='cuda'); dst = torch.zeros(world * 1057, dtype=torch.float64, device='cuda')
            fn = lambda: dist.all_gather_into_tensor(dst, src)
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): fn()
        e1.record(); torch.cuda.synchronize()
        if rank == 0: print('%s: %.2f us per call (back to back)' % (name, e0.elapsed_time(e1) * 1000 / 200), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
