"""Back-to-back timing of distributed operators: torchrun --nproc-per-node N tools/prof_dist.py n reps ops..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pyranda_b200.distributed import DistributedParcop

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]); reps = int(sys.argv[2]); ops = sys.argv[3:] or ["ddz", "sfilter", "gfilter"]
gz = int(os.environ.get("PB_GLOBAL_NZ", n * world))
L = 2 * np.pi
per = os.environ.get("PB_BOUNDED", "0") != "1"
eng = DistributedParcop(n, n, gz, 0, L, 0, L, 0, L, periodic=(per,) * 3, device=local)
f = eng.empty(); f.copy_(torch.rand(tuple(reversed(eng.plan.shape)), dtype=torch.float64, device="cuda").permute(2, 1, 0))
out = eng.empty()
for name in ops:
    for _ in range(5):
        eng.apply_into(name, f, out)
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            eng.apply_into(name, f, out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    t = torch.tensor([sorted(ts)[1]], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-9s %.4f ms  (env: %s)" % (name, t.item(), " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("PB_"))), flush=True)
dist.barrier(); dist.destroy_process_group()
