"""z-slab multi-GPU operators: one process per GPU, torch.distributed for the plumbing.

Reference counterpart: the np > 1 branches of eval_compact_op1z_{d1,r3,r4}
(pyranda/parcop/compact_d1.f90:719-746,858-928, compact_r4.f90:640-656,820-...), i.e. the halo
MPI_Sendrecv pair and the mpi_allgather of the four interface unknowns per line followed by the
redundant reduced solve.  Here only z is split (x and y sweeps stay local), the halo is a grouped
send/recv of 3-4 xy-planes straight out of the field's own memory (z-planes are contiguous), the
interface exchange is one all_gather_into_tensor, and the reduced solve + spike correction +
scale / add-back are one kernel (pb_z_finish).

Works with NCCL on GPUs and, for host-logic tests, with gloo on CPU tensors when the plan is bound
to the emulated library (tests/emul).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import OP, check
from .plan import ParcopPlan

# z operator behind each distributed call and its halo width (nor of the stencil)
_ZOPS = {"ddz": ("ddz", 3), "dd8z": ("dd8z", 4), "d2z": ("d2z", 3), "sfilterz": ("sfilterz", 4), "gfilterz": ("gfilterz", 4)}
_IMPLICIT = {"ddz": True, "dd8z": True, "d2z": True, "sfilterz": True, "gfilterz": False}


class DistributedParcop:
    def __init__(self, nx, ny, nz, x1=0.0, xn=1.0, y1=0.0, yn=1.0, z1=0.0, zn=1.0, periodic=(False, False, False),
                 coordsys=0, device=-1, group=None, lib=None, tensor_device=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.plan = ParcopPlan(nx, ny, nz, x1, xn, y1, yn, z1, zn, periodic=periodic, px=1, py=1, pz=self.world,
                               coords=(0, 0, self.rank), coordsys=coordsys, device=device, lib=lib)
        self.periodic_z = bool(periodic[2])
        ax, ay, az = self.plan.shape
        self.plane = ax * ay
        if tensor_device is None:
            tensor_device = torch.device("cuda", torch.cuda.current_device())
        self.dev = torch.device(tensor_device)
        mk = lambda n: torch.zeros(n, dtype=torch.float64, device=self.dev)
        self.recv_lo, self.recv_hi = mk(4 * self.plane), mk(4 * self.plane)
        self.iface_local, self.iface_all = mk(4 * self.plane), mk(self.world * 4 * self.plane)
        self._tmp = None
        # MPI_CART_SHIFT (comm.f90:186)
        self.lo = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.periodic_z else None)
        self.hi = self.rank + 1 if self.rank < self.world - 1 else (0 if self.periodic_z else None)

    # ---------------------------------------------------------------- helpers
    def empty(self):
        ax, ay, az = self.plan.shape
        return torch.empty((az, ay, ax), dtype=torch.float64, device=self.dev).permute(2, 1, 0)

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream if self.dev.type == "cuda" else 0

    def _planes(self, f):
        """(az, ay, ax) contiguous view of a Fortran-strided field: z-planes are contiguous."""
        ax, ay, az = self.plan.shape
        assert f.stride() == (1, ax, ax * ay), "fields must have Fortran strides (use .empty())"
        return f.permute(2, 1, 0)

    def _global_rank(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _halo_exchange(self, f, h):
        """compact_d1.f90:719-735: last h planes -> hi neighbour's lower halo, first h planes -> lo
        neighbour's upper halo.  Sent straight from the field (no pack)."""
        pl = self._planes(f)
        n = h * self.plane
        ops = []
        if self.hi is not None:
            ops.append(dist.P2POp(dist.isend, pl[pl.shape[0] - h:].reshape(-1), self._global_rank(self.hi), self.group, tag=0))
        if self.lo is not None:
            ops.append(dist.P2POp(dist.isend, pl[:h].reshape(-1), self._global_rank(self.lo), self.group, tag=1))
        if self.lo is not None:
            ops.append(dist.P2POp(dist.irecv, self.recv_lo[:n], self._global_rank(self.lo), self.group, tag=0))
        if self.hi is not None:
            ops.append(dist.P2POp(dist.irecv, self.recv_hi[:n], self._global_rank(self.hi), self.group, tag=1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    # ---------------------------------------------------------------- operators
    def zop_into(self, name, f, out):
        """One distributed z sweep: halo exchange -> local solve -> all-gather -> finish."""
        opname, h = _ZOPS[name]
        code = OP[opname]
        P, L = self.plan, self.plan.L
        if self.world > 1:
            self._halo_exchange(f, h)
        st = self._stream()
        check(L, L.pb_z_local(P._h, code, f.data_ptr(), self.recv_lo.data_ptr(), self.recv_hi.data_ptr(), out.data_ptr(),
                              self.iface_local.data_ptr(), st))
        if self.world > 1 and _IMPLICIT[name]:
            dist.all_gather_into_tensor(self.iface_all, self.iface_local, group=self.group)  # compact_d1.f90:890
            check(L, L.pb_z_finish(P._h, code, f.data_ptr(), self.iface_all.data_ptr(), out.data_ptr(), st))
        return out

    def _local_into(self, opname, f, out):
        check(self.plan.L, self.plan.L.pb_apply(self.plan._h, OP[opname], f.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def apply_into(self, name, f, out):
        """ddx ddy ddz dd8x dd8y dd8z d2x d2y d2z sfilter gfilter gfilterx/y/z laplacian ring."""
        if name in ("ddx", "ddy", "dd8x", "dd8y", "d2x", "d2y", "gfilterx", "gfiltery", "sfilterx", "sfiltery"):
            return self._local_into(name, f, out)
        if name in _ZOPS:
            return self.zop_into(name, f, out)
        if self._tmp is None:
            self._tmp = self.empty()
        tmp = self._tmp
        if name in ("sfilter", "gfilter"):  # operators.f90:849-851,873-875
            k = name[0]
            self._local_into(k + "filterx", f, out)
            self._local_into(k + "filtery", out, tmp)
            return self.zop_into(k + "filterz", tmp, out)
        if name == "laplacian":  # operators.f90:523-526
            self._local_into("d2x", f, out)
            self._local_into("d2y", f, tmp)
            out.add_(tmp)
            self.zop_into("d2z", f, tmp)
            return out.add_(tmp)
        if name == "ring":  # operators.f90:638 with the Cartesian d1 = dx, d2 = dy, d3 = dz
            self._local_into("dd8x", f, out)
            out.abs_().mul_(self.plan.dx ** 2)
            for nm, d in (("dd8y", self.plan.dy),):
                self._local_into(nm, f, tmp)
                torch.maximum(out, tmp.abs_().mul_(d ** 2), out=out)
            self.zop_into("dd8z", f, tmp)
            return torch.maximum(out, tmp.abs_().mul_(self.plan.dz ** 2), out=out)
        raise KeyError(name)

    def apply(self, name, f):
        return self.apply_into(name, f, self.empty())

    def divergence(self, fx, fy, fz):  # operators.f90:48-52
        out = self.empty()
        if self._tmp is None:
            self._tmp = self.empty()
        self._local_into("ddx", fx, out)
        self._local_into("ddy", fy, self._tmp)
        out.add_(self._tmp)
        self.zop_into("ddz", fz, self._tmp)
        return out.add_(self._tmp)

    def grads(self, f):  # operators.f90:191-193
        return self.apply("ddx", f), self.apply("ddy", f), self.apply("ddz", f)

    def apply_host_into(self, name, a_in, a_out):
        """Host arrays in / out (the f2py call shape) around the distributed device operator."""
        if not hasattr(self, "_hin"):
            self._hin, self._hout = self.empty(), self.empty()
        self._planes(self._hin).copy_(torch.from_numpy(a_in.T), non_blocking=True)
        self.apply_into(name, self._hin, self._hout)
        torch.from_numpy(a_out.T).copy_(self._planes(self._hout), non_blocking=True)
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()

    # pyrandaMPI.py:307-326
    def sum3D(self, f):
        t = torch.tensor([self._reduce("sum", f)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.item()

    def max3D(self, f):
        t = torch.tensor([self._reduce("max", f)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t.item()

    def min3D(self, f):
        t = torch.tensor([self._reduce("min", f)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t.item()

    def _reduce(self, kind, f):
        import ctypes
        out = ctypes.c_double()
        from ._lib import REDUCE
        check(self.plan.L, self.plan.L.pb_reduce(self.plan._h, REDUCE[kind], f.numel(), f.data_ptr(), ctypes.byref(out), self._stream()))
        return out.value
