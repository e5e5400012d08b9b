import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(4 * 1024 * 1024, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        print("symm ok: world", hdl.world_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "multicast_ptr", hex(hdl.multicast_ptr),
              "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2], "buffer_ptrs_dev", hex(hdl.buffer_ptrs_dev), "signal_pad_ptrs_dev", hex(hdl.signal_pad_ptrs_dev), flush=True)
        print([a for a in dir(hdl) if not a.startswith('_')], flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    # try the built-in multimem all-reduce if present
    try:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        if rank == 0: print("multimem_all_reduce_ ok:", t[:3].tolist(), flush=True)
    except Exception as e:
        if rank == 0: print("multimem_all_reduce_ failed:", repr(e)[:300], flush=True)
except Exception as e:
    print(rank, "symm failed:", repr(e)[:500], flush=True)
dist.destroy_process_group()
