"""z-slab operators over NCCL on 2+ GPUs of one box against the one-rank oracle (same worker as the
gloo test, real library, CUDA tensors). Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import domain, rel_linf, synthetic_field
from oracle import oracle
from pyranda_b200.distributed import DistributedParcop
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
worst = 0.0
for n in ((32, 48, 32 * world), (64, 64, 64 * world)):
    for periodic in (True, False):
        (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
        o = oracle.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
        f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
        eng = DistributedParcop(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, device=local)
        az = n[2] // world
        sl = slice(rank * az, (rank + 1) * az)
        loc = eng.empty(); loc.copy_(torch.from_numpy(f[:, :, sl].copy()))
        for name, ref in (("ddx", o.ddx), ("ddy", o.ddy), ("ddz", o.ddz), ("dd8z", o.dd8z), ("d2z", o.d2z),
                          ("sfilter", o.sfilter), ("gfilter", o.gfilter), ("laplacian", o.plaplacian), ("ring", o.pring)):
            got = eng.apply(name, loc).cpu().numpy()
            err = rel_linf(got, ref(f)[:, :, sl])
            worst = max(worst, err)
            assert err < 1e-12, (name, n, periodic, rank, err)
        got = eng.divergence(loc, loc * 2, loc * loc).cpu().numpy()
        assert rel_linf(got, o.divergence(f, 2 * f, f * f)[:, :, sl]) < 1e-12
        assert abs(eng.sum3D(loc) - f.sum()) < 1e-9 * np.abs(f).sum()
        assert eng.max3D(loc) == f.max() and eng.min3D(loc) == f.min()
        # host-array entry point (pinned staging inside the engine)
        hin = np.asfortranarray(f[:, :, sl]); hout = np.empty_like(hin, order="F")
        eng.apply_host_into("ddz", hin, hout)
        assert rel_linf(hout, o.ddz(f)[:, :, sl]) < 1e-12
        # neighbour planes for viz dumps (pyrandaMPI.ghost): P2P over NCCL on the device fields
        from pyranda_b200.distributed import _make_backend
        gh = _make_backend()(eng).ghost_host(loc)
        g_lo = rank * az - (1 if rank > 0 else 0)
        g_n = az + (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
        assert gh.shape == (n[0], n[1], g_n) and np.array_equal(gh, f[:, :, g_lo:g_lo + g_n]), (rank, gh.shape)
        for name, ref in (("ddx", o.ddx), ("ddy", o.ddy)):  # local directions: the rank's own slab pipeline
            eng.apply_host_into(name, hin, hout)
            assert rel_linf(hout, ref(f)[:, :, sl]) < 1e-12, (name, n, periodic, rank)
# long slabs: neighbour-only interface exchange and the pipelined kernels with halo planes
from pyranda_b200._lib import OP
for n in ((32, 32, 128 * world), (48, 32, 256 * world)):
    for periodic in (True, False):
        (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
        o = oracle.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
        f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
        eng = DistributedParcop(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, device=local)
        az = n[2] // world
        sl = slice(rank * az, (rank + 1) * az)
        loc = eng.empty(); loc.copy_(torch.from_numpy(f[:, :, sl].copy()))
        for name, ref in (("ddz", o.ddz), ("d2z", o.d2z), ("dd8z", o.dd8z), ("sfilter", o.sfilter), ("gfilter", o.gfilter)):
            err = rel_linf(eng.apply(name, loc).cpu().numpy(), ref(f)[:, :, sl])
            worst = max(worst, err)
            assert err < 1e-12, (name, n, periodic, rank, err)
        assert "ddz" in eng._ring or eng._xmask[OP["ddz"]] == "neighbours", (eng._ring, eng._xmask)
        if os.environ.get("PB_NO_PEER_MEMORY", "0") != "1" and os.environ.get("PB_NO_RING", "0") != "1" and az >= 256:
            assert "ddz" in eng._ring and "sfilterz" in eng._ring, eng._ring   # the fused sweep ran
from pyranda_b200 import _lib
print("rank", rank, "worst", worst, "ring launches", _lib.load().pb_ring_launch_count())
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "partitioned-peer", "partitioned-nccl"])
def test_zslab_nccl(tmp_path, mode):
    import torch
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if world < 4 else (4 if world < 8 else 8)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    # fused: ring kernel with chunk states over peer memory; partitioned-peer: local solve + correction with the
    # peer-memory exchange kernel; partitioned-nccl: the same over NCCL send/recv (north_star's plumbing)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    if mode == "partitioned-peer":
        env["PB_NO_RING"] = "1"
    if mode == "partitioned-nccl":
        env["PB_NO_PEER_MEMORY"] = "1"
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    at = r.stderr.find("Traceback")
    assert r.returncode == 0, r.stdout[-1500:] + (r.stderr[at:at + 3000] if at >= 0 else r.stderr[-3000:])
    assert r.stdout.count("worst") == world
