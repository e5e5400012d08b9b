"""The one collective of the path, alone: all-reduce(sum) of the FP32 P-vector (P = 2 837 314, 11.3 MB) over the ranks of
one box, through the direct NCCL communicator the solver uses.  Run under torchrun with different NCCL_ALGO / NCCL_PROTO
settings to see what the library picks and what it costs."""
import os, sys, torch, torch.distributed as dist
sys.path[:0] = ['.']
from pytorchhessianfree_b200.dist import all_reduce_sum
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev); g = dist.group.WORLD; n = dist.get_world_size()
for P in (2837314, 669706):
    buf = torch.randn(P, device=dev)
    for _ in range(10): all_reduce_sum(buf, g)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): all_reduce_sum(buf, g)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    if rank == 0:
        print(f"NCCL_ALGO={os.environ.get('NCCL_ALGO','auto')} NCCL_PROTO={os.environ.get('NCCL_PROTO','auto')} ranks={n} P={P}: {us:.1f} us, bus {2*(n-1)/n*4*P/us/1e3:.0f} GB/s", flush=True)
dist.destroy_process_group()
