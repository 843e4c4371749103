"""Time the pieces of a distributed z sweep (halo exchange, local pass, interface exchange, finish)
with CUDA events: torchrun --nproc-per-node N tools/prof_zsplit.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pyranda_b200._lib import OP, check
from pyranda_b200.distributed import DistributedParcop, _IMPLICIT, _ZOPS

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = 2 * np.pi
eng = DistributedParcop(n, n, n * world, 0, L, 0, L, 0, L * world, periodic=(True,) * 3, device=local)
f = eng.empty(); f.copy_(torch.rand((n, n, n), dtype=torch.float64, device="cuda"))
out = eng.empty()
P, Lb = eng.plan, eng.plan.L


def ev():
    return torch.cuda.Event(enable_timing=True)


for name in ("ddz", "sfilterz", "gfilterz", "dd8z"):
    opname, h = _ZOPS[name]
    code = OP[opname]
    st = torch.cuda.current_stream().cuda_stream
    acc = np.zeros(4)
    reps = 12
    for it in range(reps):
        e = [ev() for _ in range(5)]
        if eng._pb is not None:
            eng._pb.k += 1
            rlo, rhi = eng._pb.view("lo"), eng._pb.view("hi")
            iall, iloc = eng._pb.iface_all(), eng._pb.view("iface", rank)
        else:
            rlo, rhi, iall, iloc = eng.recv_lo, eng.recv_hi, eng.iface_all, eng.iface_local
        e[0].record()
        eng._halo_exchange(f, h)
        e[1].record()
        check(Lb, Lb.pb_z_local(P._h, code, f.data_ptr(), rlo.data_ptr(), rhi.data_ptr(), out.data_ptr(), iloc.data_ptr(), st))
        e[2].record()
        if _IMPLICIT[name]:
            eng._iface_exchange(code, iall, iloc)
        e[3].record()
        if _IMPLICIT[name]:
            check(Lb, Lb.pb_z_finish(P._h, code, f.data_ptr(), iall.data_ptr(), out.data_ptr(), st))
        e[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            acc += np.array([e[i].elapsed_time(e[i + 1]) for i in range(4)])
    acc /= reps - 2
    if rank == 0:
        print("%-9s halo %.3f  local %.3f  iface-exchange %.3f  finish %.3f  total %.3f ms  (%s)" % (
            name, acc[0], acc[1], acc[2], acc[3], acc.sum(), eng._xmask.get(code, "-")), flush=True)
dist.barrier()
dist.destroy_process_group()
