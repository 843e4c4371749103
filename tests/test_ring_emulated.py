"""The ring kernel (pyranda_b200/csrc/ring.cuh) in the host-emulated build: thread-block clusters for
long lines, 64-line tiles for short slabs, and the fused z-slab sweep whose chunk states cross the
slab faces as self-validating records -- here between host threads that stand in for the ranks (one
plan per "rank", record / halo buffers in process memory, the kernels of all ranks running
concurrently exactly as they do on neighbouring GPUs)."""
import ctypes
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", EMUL, "-s", "-j8"])
    from pyranda_b200 import _lib
    L = _lib.load(os.path.join(EMUL, "libparcop_emul.so"))
    L.pb_set_tuning(0, 0, 32)
    return L


def _pair(n, periodic, oracle_mod, lib):
    from pyranda_b200 import ParcopPlan
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, lib=lib, tensor_device="cpu")
    p.set_mesh()
    return o, p, synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("n,lines,names", [
    ((16, 1024, 3), 0, ("ddy", "sfilter")),          # 4-CTA clusters, 32-line tiles, partial tile
    ((40, 3, 1024), 16, ("ddz", "d2z", "dd8z")),     # 2-CTA clusters, 16-line tiles
    ((70, 128, 3), 0, ("ddy", "sfilter", "pring", "plaplacian")),  # 64-line tiles, TMA reduce epilogues
])
def test_ring_kernel_long_and_short_lines(n, lines, names, periodic, oracle_mod, lib):
    lib.pb_set_ring(1, lines)
    try:
        o, p, f = _pair(n, periodic, oracle_mod, lib)
        for nm in names:
            r0 = lib.pb_ring_launch_count()
            got = getattr(p, nm)(f)
            assert lib.pb_ring_launch_count() > r0, "the ring kernel did not run"
            assert rel_linf(got, getattr(o, nm)(f)) < 1e-13, (nm, n, periodic)
    finally:
        lib.pb_set_ring(1, 0)


@pytest.mark.parametrize("periodic", [True, False])
def test_ring_kernel_wherever_it_fits(periodic, oracle_mod, lib):
    """mode 2: the ring kernel also where the one-CTA pipelined kernel fits (256- and 512-point lines;
    512 points with 32-line tiles is a 2-CTA cluster)."""
    lib.pb_set_ring(2, 32)
    try:
        for n, names in (((20, 3, 512), ("ddz", "sfilter")), ((20, 256, 2), ("ddy", "dd8y"))):
            o, p, f = _pair(n, periodic, oracle_mod, lib)
            for nm in names:
                r0 = lib.pb_ring_launch_count()
                got = getattr(p, nm)(f)
                assert lib.pb_ring_launch_count() > r0
                assert rel_linf(got, getattr(o, nm)(f)) < 1e-13, (nm, n, periodic)
    finally:
        lib.pb_set_ring(1, 0)


def _fused_zslab(lib, oracle_mod, n, world, periodic, ops, push=False):
    """z-slab of `world` ranks as threads: halo planes copied by hand, chunk states exchanged by the
    kernels themselves through the record buffers."""
    from pyranda_b200 import ParcopPlan
    from pyranda_b200._lib import OP, XRingC, check
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    ax, ay, az = n[0], n[1], n[2] // world
    plane = ax * ay
    plans = [ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, pz=world, coords=(0, 0, r), lib=lib, tensor_device="cpu")
             for r in range(world)]
    slabs = [np.asfortranarray(f[:, :, r * az:(r + 1) * az]) for r in range(world)]
    cap = 8
    en = [np.zeros(cap * plane * 4, dtype=np.uint64) for _ in range(world)]
    st = [np.zeros(cap * plane * 4, dtype=np.uint64) for _ in range(world)]
    epoch = 0
    worst = {}
    modes = set()
    flags = [np.zeros(world, dtype=np.uint64) for _ in range(world)]   # flags[r][p]: set by rank p in rank r's memory
    counters = [np.zeros(2, dtype=np.uint32) for _ in range(world)]
    for name, h, epi, s2, ref in ops:
        code = OP[name]
        info = [ctypes.c_int() for _ in range(4)]
        check(lib, lib.pb_z_ring_info(plans[0]._h, code, *[ctypes.byref(c) for c in info]))
        if info[0].value + info[1].value == 0:  # states would wrap onto the rank itself: no fused form (SPIKE path)
            worst[name + str(epi)] = None
            continue
        modes.add(lib.pb_z_ring_mode(plans[0]._h, code))
        epoch += 1
        outs = [np.asfortranarray(np.full((ax, ay, az), 0.5)) for _ in range(world)]
        halos = []
        for r in range(world):  # compact_d1.f90:719-735
            lo = np.zeros((ax, ay, 4), order="F"); hi = np.zeros((ax, ay, 4), order="F")
            below, above = r - 1, r + 1
            if periodic:
                below %= world; above %= world
            if below >= 0 and not push:
                lo[:, :, :h] = slabs[below][:, :, az - h:]
            if above < world and not push:
                hi[:, :, :h] = slabs[above][:, :, :h]
            halos.append((lo, hi))
        errs = []

        def run(r):
            try:
                x = XRingC()
                x.epoch = epoch
                x.en_in, x.st_in = en[r].ctypes.data, st[r].ctypes.data
                for k in range(3):
                    x.en_out[k] = en[(r + 1 + k) % world].ctypes.data
                    x.st_out[k] = st[(r - 1 - k) % world].ctypes.data
                if push:  # the kernel moves the halo planes itself and waits for the neighbours' flags
                    below, above = r - 1, r + 1
                    if periodic:
                        below %= world; above %= world
                    below = below if below >= 0 else None
                    above = above if above < world else None
                    peers = sorted({q for q in (below, above) if q is not None})
                    x.push, x.npeers, x.halo_epoch = 1, len(peers), epoch
                    x.halo_dst[0] = None if below is None else halos[below][1].ctypes.data
                    x.halo_dst[1] = None if above is None else halos[above][0].ctypes.data
                    for q, pr in enumerate(peers):
                        x.flag_remote[q] = flags[pr].ctypes.data + 8 * r
                        x.flag_local[q] = flags[r].ctypes.data + 8 * pr
                    x.counter = counters[r].ctypes.data
                check(lib, lib.pb_z_ring(plans[r]._h, code, slabs[r].ctypes.data, halos[r][0].ctypes.data, halos[r][1].ctypes.data,
                                         outs[r].ctypes.data, ctypes.byref(x), epi, s2, None))
            except Exception as exc:  # noqa: BLE001
                errs.append(exc)
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        want = ref(o, f)
        got = np.concatenate(outs, axis=2)
        worst[name + str(epi)] = rel_linf(got, want)
    worst["modes"] = modes
    return worst


OPS = [
    ("ddz", 3, 0, 0.0, lambda o, f: o.ddz(f)),
    ("sfilterz", 4, 0, 0.0, lambda o, f: o.dir_op("sf", 2, f)),
    ("d2z", 3, 1, 0.0, lambda o, f: o.d2z(f) + 0.5),                       # out += val on the 0.5 the test preloads
    ("dd8z", 4, 3, 2.0, lambda o, f: np.maximum(0.5, np.abs(o.dd8z(f)) * 2.0)),
]


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("push", [False, True])
@pytest.mark.parametrize("n,world,fused", [((20, 3, 256), 2, 3), ((34, 2, 384), 3, 4), ((16, 2, 512), 4, 4), ((16, 2, 1024), 2, 4),
                                           ((64, 64, 1024), 8, 4)])   # the bench line's parity slab at 8 ranks: 64-line tiles, 128-plane slabs
def test_fused_zslab_ranks_as_threads(n, world, fused, push, periodic, oracle_mod, lib):
    worst = _fused_zslab(lib, oracle_mod, n, world, periodic, OPS, push)
    modes = worst.pop("modes")
    assert 2 in modes, modes   # the non-waiting form ran for at least one operator
    if n[2] // world == 128 and world == 3 and periodic:
        assert 1 in modes, modes   # the compact filter's states come from two ranks away: the waiting form
    done = {k: v for k, v in worst.items() if v is not None}
    assert len(done) >= (fused if periodic else 4), worst
    assert max(done.values()) < 2e-14, worst   # a dropped far-chunk state of the compact filter is 4e-12
