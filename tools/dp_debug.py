"""Two-rank acc_step with a stack dump if it stalls (debugging aid): torchrun --nproc-per-node 2 tools/dp_debug.py"""
import faulthandler
import os
import sys
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
faulthandler.dump_traceback_later(45, exit=True)
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from test_gpu_dist import _train  # noqa: E402

try:
    flat, iters, losses = _train(rank, dist.get_world_size(), dist.group.WORLD)
    torch.cuda.synchronize()
    print(rank, "ok", iters, losses, float(flat.double().sum()), flush=True)
except Exception as e:  # noqa: BLE001
    import traceback
    traceback.print_exc()
    print(rank, "FAILED", repr(e)[:300], flush=True)
dist.destroy_process_group()
