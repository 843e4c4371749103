"""ctypes binding of libparcop_b200.so (C ABI: include/parcop_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``python -m pyranda_b200.build``
and must be present: there is no CPU fallback, a missing library is an ImportError-style failure.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libparcop_b200.so")

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

# opcodes (include/parcop_b200.h)
OP = dict(ddx=0, ddy=1, ddz=2, dd8x=3, dd8y=4, dd8z=5, d2x=6, d2y=7, d2z=8, laplacian=9, ring=10,
          sfilter=11, gfilter=12, gfilterx=13, gfiltery=14, gfilterz=15, sfilterx=16, sfiltery=17,
          sfilterz=18, ddx_odd=19, ddy_odd=20, ddz_odd=21, dd4x=22, dd4y=23, dd4z=24, dd8x_odd=25, dd8y_odd=26, dd8z_odd=27)
REDUCE = dict(sum=0, max=1, min=2)

EXPORTS = [
    "pb_last_error", "pb_version", "pb_plan_create", "pb_plan_destroy", "pb_plan_extents",
    "pb_plan_spacing", "pb_plan_set_mesh", "pb_getvar", "pb_getvar_device", "pb_apply",
    "pb_divergence", "pb_grads", "pb_divergence_tensor", "pb_ring_vector", "pb_host_divergence_tensor", "pb_host_ring_vector", "pb_rk4_stage", "pb_reduce", "pb_reduce_device", "pb_z_pack_halo", "pb_z_local",
    "pb_z_finish", "pb_z_exchange_ranks", "pb_peer_exchange", "pb_host_apply", "pb_host_divergence", "pb_host_grads", "pb_launch_count",
    "pb_pipe_launch_count", "pb_set_tuning", "pb_ring_launch_count", "pb_set_ring",
    "pb_z_ring_info", "pb_z_ring", "pb_z_ring_mode", "pb_apply_epi",
]


class XRingC(ctypes.Structure):
    """pb_xring of include/parcop_b200.h."""
    _fields_ = [("epoch", ctypes.c_uint), ("en_in", _vp), ("st_in", _vp), ("en_out", _vp * 3), ("st_out", _vp * 3),
                ("push", ctypes.c_int), ("npeers", ctypes.c_int), ("halo_dst", _vp * 2), ("flag_remote", _vp * 2),
                ("flag_local", _vp * 2), ("halo_epoch", ctypes.c_ulonglong), ("counter", _vp)]


class ParcopError(RuntimeError):
    pass


def declare(L):
    """Attach argtypes/restypes of every entry point declared in include/parcop_b200.h."""
    i, d, cp = ctypes.c_int, ctypes.c_double, ctypes.c_char_p
    L.pb_last_error.restype = cp
    L.pb_version.restype = cp
    L.pb_plan_create.argtypes = [ctypes.POINTER(_vp)] + [i] * 10 + [d] * 6 + [cp] * 6 + [i]
    L.pb_plan_destroy.argtypes = [_vp]
    L.pb_plan_extents.argtypes = [_vp] + [ctypes.POINTER(i)] * 3
    L.pb_plan_spacing.argtypes = [_vp] + [_dp] * 3
    L.pb_plan_set_mesh.argtypes = [_vp, _vp, _vp, _vp, i]
    L.pb_getvar.argtypes = [_vp, cp, _vp]
    L.pb_getvar_device.argtypes = [_vp, cp, ctypes.POINTER(_vp)]
    L.pb_apply.argtypes = [_vp, i, _vp, _vp, _vp]
    L.pb_divergence.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
    L.pb_grads.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
    L.pb_rk4_stage.argtypes = [_vp, ctypes.c_long, d, d, d, _vp, _vp, _vp, _vp]
    L.pb_reduce.argtypes = [_vp, i, ctypes.c_long, _vp, _dp, _vp]
    L.pb_z_pack_halo.argtypes = [_vp, i, _vp, _vp, _vp, _vp]
    L.pb_z_local.argtypes = [_vp, i, _vp, _vp, _vp, _vp, _vp, _vp]
    L.pb_z_finish.argtypes = [_vp, i, _vp, _vp, _vp, _vp]
    L.pb_peer_exchange.argtypes = [i, _vp, _vp, _vp, i, _vp, _vp, ctypes.c_ulonglong, _vp, _vp]
    L.pb_z_exchange_ranks.argtypes = [_vp, i, ctypes.POINTER(ctypes.c_ulonglong)]
    L.pb_host_apply.argtypes = [_vp, i, _vp, _vp]
    L.pb_host_divergence.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.pb_host_grads.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.pb_divergence_tensor.argtypes = [_vp] * 14
    L.pb_ring_vector.argtypes = [_vp] * 6
    L.pb_host_divergence_tensor.argtypes = [_vp, _vp, _vp]
    L.pb_host_ring_vector.argtypes = [_vp] * 5
    L.pb_reduce_device.argtypes = [_vp, i, ctypes.c_long, _vp, _vp, _vp]
    L.pb_launch_count.restype = ctypes.c_long
    L.pb_pipe_launch_count.restype = ctypes.c_long
    L.pb_set_tuning.argtypes = [i, i, i]
    L.pb_ring_launch_count.restype = ctypes.c_long
    L.pb_set_ring.argtypes = [i, i]
    L.pb_z_ring_info.argtypes = [_vp, i] + [ctypes.POINTER(i)] * 4
    L.pb_z_ring.argtypes = [_vp, i, _vp, _vp, _vp, _vp, ctypes.POINTER(XRingC), i, d, _vp]
    L.pb_z_ring_mode.argtypes = [_vp, i]
    L.pb_apply_epi.argtypes = [_vp, i, _vp, _vp, i, d, _vp]
    return L


_LIB = None


def load(path=None):
    """Load (once) and return the CUDA library.  Raises if it has not been built."""
    global _LIB
    if path is not None:
        return declare(ctypes.CDLL(path))
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ParcopError(
                "libparcop_b200.so is missing (%s): build it with `python -m pyranda_b200.build`; "
                "there is no CPU fallback" % LIB_PATH)
        _LIB = declare(ctypes.CDLL(LIB_PATH))
    return _LIB


def check(L, rc):
    if rc != 0:
        raise ParcopError("parcop_b200 error %d: %s" % (rc, L.pb_last_error().decode()))
