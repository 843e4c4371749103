#!/usr/bin/env python
"""bench.py -- operator microbenchmark of the parcop compact-operator hot path on B200.

One "step" applies ddx, ddy, ddz, filter (compact 8th-order, 3 sweeps) and gfilter (Gaussian,
3 sweeps) once each to a synthetic fp64 field of 512^3 points per GPU (BASELINE.json metric
"fp64 Gpoints/s for compact ddx/ddy/ddz+filter at 512^3").  value = operator applications x grid
points / second over the whole job.  With N GPUs the domain is 512 x 512 x (512 N), z-slab
partitioned (weak scaling): x / y sweeps are local, z sweeps exchange halos and interface unknowns
over NCCL.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the CPU path (oracle port, all host threads)
    ... bench.py --gpus N --global-n 1024   # BASELINE configs[2] as the headline workload (strong scaling)

Every line also carries, measured in the same run:
  strong_1024     BASELINE configs[2]: the same five operators on a FIXED 1024^3 periodic field split into N z-slabs
                  (strong scaling: efficiency = T(1) / (N T(N)) across the lines of a 1/2/4/8 run);
  per_op_bounded  the operators (+ ring, laplacian) with one-sided closures (BC NONE) at 512^3 per GPU;
  parity          a 64 x 64 x (128 N) field through the same engine against the CPU oracle, max relative L-inf.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OPS = ("ddx", "ddy", "ddz", "sfilter", "gfilter")
SWEEPS = {"ddx": 1, "ddy": 1, "ddz": 1, "sfilter": 3, "gfilter": 3}
BYTES_PER_POINT = {"ddx": 16, "ddy": 16, "ddz": 16, "sfilter": 48, "gfilter": 48}
METRIC = "fp64 Gpoints/s (operator applications x points / s), compact ddx+ddy+ddz+filter+gfilter"
NPER = 512  # points per side per GPU
# ncu --set full, 512^3 launch of the periodic y/z sweep kernel (the kernel of ddy and ddz): 1.076544 GB read + 1.034750 GB
# written (profiles/r2_pipe_and_ring_kernels_ncu_full.txt; round 1: 1.073867 + 1.035293, profiles/r1_v7_tma_kernels_ncu_full.txt)
NCU_TRAFFIC_DDZ_512 = 2111294000


def bench_config(world, n, global_n=0):
    """The workload both arms are quoted on (BASELINE.json configs[1]: 512^3 fp64 per GPU); with
    --global-n N the fixed N^3 grid of configs[2] split into z-slabs (strong scaling)."""
    grid = [global_n] * 3 if global_n else [n, n, n * world]
    per = [grid[0], grid[1], grid[2] // world]
    return {"workload": "operator microbench: ddx, ddy, ddz, filter, gfilter once each per step on a periodic fp64 field",
            "global_grid": grid, "per_gpu_grid": per, "partition": "z-slab x%d" % world,
            "l2": "input %.2f GB per field per GPU, larger than the 126 MB L2; no explicit flush" % (per[0] * per[1] * per[2] * 8 / 1e9),
            "chunk_len": 32}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use the
    whole host (it runs on rank 0 alone), so the variable is reset before the OpenMP runtime loads."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


class CpuArm:
    """The CPU oracle (port of the reference's algorithm, OpenMP over line bundles, all host cores) on
    the five operators of one step, n^3 periodic field built once."""

    def __init__(self, n):
        import numpy as np
        self.cores = all_host_threads()
        from oracle import oracle
        oracle.build()
        self.n = n
        L = 2 * np.pi * (n - 1) / n
        self.o = oracle.Oracle(n, n, n, 0, L, 0, L, 0, L, periodic=(True,) * 3)
        ax = L / (n - 1) * np.arange(n) if n > 1 else np.zeros(1)
        rng = np.random.default_rng(1234)
        f = np.sin(3 * ax)[:, None, None] * np.cos(2 * ax)[None, :, None] * np.cos(ax)[None, None, :]
        f += 0.1 * rng.uniform(-1, 1, size=f.shape)
        self.f = np.asfortranarray(f)
        self.threads = oracle.num_threads()

    def step(self):
        t0 = time.perf_counter()
        for name in OPS:
            getattr(self.o, name)(self.f)
        return time.perf_counter() - t0

    def describe(self, value, reps):
        n = self.n
        return {"value": value, "unit": "Gpoints/s", "cores": self.threads, "kind": "port",
                "sample": ("the same 5 operators on a %d^3 periodic fp64 field (%s one GPU's share of the workload), %d step(s), host arrays, "
                           "OpenMP threads = host cores" % (n, "all of" if n >= NPER else "1/%d of" % ((NPER // n) ** 3), reps))}


def cpu_sample(n=256, reps=1):
    """A bounded sample of the workload on the CPU arm (the in-line cpu_baseline of the N = 1 line)."""
    arm = CpuArm(n)
    arm.step()  # warm-up
    dt = sum(arm.step() for _ in range(reps)) / reps
    return arm.describe(len(OPS) * n ** 3 / dt / 1e9, reps)


def tgv_step_time(n, warm=2, steps=3):
    """BASELINE.json configs[1]: Taylor-Green vortex (examples/TaylorGreen.py deck, verbatim) at n^3,
    device-resident RK4 steps through pyranda_b200.sim; returns seconds per RK4 step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    from pyranda_b200.sim import pyrandaSim
    ss = pyrandaSim("TGvortex", tgv_mesh(n))
    ss.EOM(TGV_EOM)
    ss.setIC(TGV_IC)
    time_, dt = 0.0, ss.variables["dt"] * 0.5
    l0 = ss.B.plan.launch_count()
    for _ in range(warm):
        time_ = ss.rk4(time_, dt)
        dt = ss.variables["dt"] * 0.5
    torch.cuda.synchronize()
    l0 = ss.B.plan.launch_count()
    f0 = ss.fuser.launches if ss.fuser is not None else 0
    t0 = time.perf_counter()
    for _ in range(steps):
        time_ = ss.rk4(time_, dt)
        dt = ss.variables["dt"] * 0.5
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    return {"config": "Taylor-Green vortex %d^3 periodic fp64, CFL 0.5, deck of examples/TaylorGreen.py" % n,
            "ms_per_rk4_step": sec * 1e3, "steps": steps, "sweeps_per_step": 255,
            "library_launches_per_step": (ss.B.plan.launch_count() - l0) / steps,
            "fused_pointwise_launches_per_step": ((ss.fuser.launches - f0) / steps) if ss.fuser is not None else 0,
            "gpoints_per_s": n ** 3 / sec / 1e9}


def tgv_cpu_sample(n=64, steps=1):
    """The same Taylor-Green deck on the CPU oracle's operators (numpy pointwise algebra, all host cores)
    at a bounded size: one RK4 step after one warm-up step."""
    all_host_threads()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    from oracle import oracle
    from oracle_backend import make_sim
    oracle.build()
    ss = make_sim(oracle, "TGvortex", tgv_mesh(n))
    ss.EOM(TGV_EOM)
    ss.setIC(TGV_IC)
    time_, dt = 0.0, ss.variables["dt"] * 0.5
    time_ = ss.rk4(time_, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        time_ = ss.rk4(time_, ss.variables["dt"] * 0.5)
    sec = (time.perf_counter() - t0) / steps
    return {"kind": "port", "cores": oracle.num_threads(), "sample": "%d^3 (the GPU figure is %s)" % (n, "256^3"),
            "ms_per_rk4_step": sec * 1e3, "gpoints_per_s": n ** 3 / sec / 1e9}


_JSON_FD = None


def quiet_stdout():
    """Libraries (NCCL's version banner, NVRTC, torch) may write to stdout; the contract is ONE JSON
    line there.  Everything written to fd 1 from here on goes to stderr, the line goes to the real one."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n  # the field one GPU holds in the other arm (512^3): the real size, not a reduced sample
    t0 = time.perf_counter()
    arm = CpuArm(n)
    for _ in range(min(max(args.warmup, 1), 3)):
        arm.step()
    K = max(1, args.steps)
    secs = [arm.step() for _ in range(K)]
    wall = time.perf_counter() - t0
    sec = sum(secs) / len(secs)
    value = len(OPS) * n ** 3 / sec / 1e9
    cb = arm.describe(value, K)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args.gpus, args.n),
            "note": "the Fortran/MPI reference cannot be built on this box (probed: no gfortran / mpif90 / mpirun / mpi4py, "
                    "profiles/r2_gpu_box_toolchain_probe.log); this is the oracle port of its algorithm on all host cores. "
                    "Each step applies the five operators to a %d^3 field (one GPU's share of the workload; Gpoints/s "
                    "does not depend on the number of slabs)" % n,
            "cpu_baseline": cb, "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    emit(line)


def build_engine(world, local, nx, ny, nz, periodic):
    """One plan (N = 1) or the z-slab engine (N > 1) on an nx x ny x nz grid; periodic -> the decks'
    [0, 2 pi (n-1)/n] extents, bounded -> [0, 1]."""
    import numpy as np
    from pyranda_b200 import ParcopPlan
    ext = [(2 * np.pi * (k - 1) / k) if periodic else 1.0 for k in (nx, ny, nz)]
    if world > 1:
        from pyranda_b200.distributed import DistributedParcop
        eng = DistributedParcop(nx, ny, nz, 0, ext[0], 0, ext[1], 0, ext[2], periodic=(periodic,) * 3, device=local)
        plan = eng.plan
    else:
        plan = ParcopPlan(nx, ny, nz, 0, ext[0], 0, ext[1], 0, ext[2], periodic=(periodic,) * 3, device=local)
        eng = None
    plan.set_mesh()
    return plan, eng


def synthetic_device_field(plan, rank, dev):
    """SURVEY 8d's field, generated on the device: this rank's slab of sin(3x) cos(2y) cos(z) + 0.1 U(-1, 1)."""
    import torch
    ax, ay, az = plan.shape
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    xi = torch.arange(ax, dtype=torch.float64, device=dev) * plan.dx
    yi = torch.arange(ay, dtype=torch.float64, device=dev) * plan.dy
    zi = (torch.arange(az, dtype=torch.float64, device=dev) + rank * az) * plan.dz
    f = plan.empty_device()
    f.copy_(torch.sin(3 * xi).view(ax, 1, 1) * torch.cos(2 * yi).view(1, ay, 1) * torch.cos(zi).view(1, 1, az))
    noise = torch.rand((az, ay, ax), dtype=torch.float64, device=dev, generator=g).permute(2, 1, 0)
    f.add_(0.2 * noise - 0.1)
    del noise
    return f


def timed_block(plan, eng, f, outs, ops, K, warm, world, dev, sampler=None):
    """W untimed + K timed steps of `ops` (one application each per step), CUDA events on the launching
    stream, barrier + synchronize on both sides, max over ranks.  Returns (ms per step, {op: ms}, launches)."""
    import torch
    import torch.distributed as dist
    stream = torch.cuda.current_stream().cuda_stream

    def one_op(name):
        if eng is not None:
            eng.apply_into(name, f, outs[name])
        else:
            plan.apply_ptr(name, f.data_ptr(), outs[name].data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        for name in ops:
            one_op(name)
    barrier()
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = plan.launch_count()
    barrier()
    e0.record()
    for k in range(K):
        for j, name in enumerate(ops):
            ev[k][j][0].record()
            one_op(name)
            ev[k][j][1].record()
    e1.record()
    barrier()
    launches = plan.launch_count() - launches0
    per = [sum(ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(K)) / K for j in range(len(ops))]
    t = torch.tensor([e0.elapsed_time(e1)] + per, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t[0].item() / K, {name: t[1 + j].item() for j, name in enumerate(ops)}, launches


def per_op_table(per_op_ms, npts, world, peak):
    out = {}
    for name, ms in per_op_ms.items():
        gbs = BYTES_PER_POINT[name] * npts / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "gpoints_per_s": npts * world / (ms * 1e-3) / 1e9, "algorithmic_GBps_per_gpu": gbs,
                     "frac_of_hbm_peak": gbs / peak}
    return out


def strong_1024_block(world, rank, local, dev, peak, steps=5, warm=3, n=1024):
    """BASELINE configs[2]: the five operators on a fixed n^3 periodic field, z-slab split over the N ranks."""
    import torch
    if n % world or n // world < 16:
        return None
    plan, eng = build_engine(world, local, n, n, n, True)
    f = synthetic_device_field(plan, rank, dev)
    out = plan.empty_device()
    outs = {name: out for name in OPS}  # one output field: 8.6 GB per field at N = 1
    ms, per_op_ms, _ = timed_block(plan, eng, f, outs, OPS, steps, warm, world, dev)
    npts = plan.npts
    res = {"config": "operator microbench on a fixed %d^3 periodic fp64 field, z-slab x%d (BASELINE configs[2])" % (n, world),
           "global_grid": [n, n, n], "per_gpu_grid": list(plan.shape), "scaling": "strong", "steps": steps, "warmup": warm,
           "ms_per_step": ms, "value": len(OPS) * npts * world / (ms * 1e-3) / 1e9, "unit": "Gpoints/s",
           "per_op": per_op_table(per_op_ms, npts, world, peak),
           "z_path": ("one rank" if eng is None else ("fused ring kernel: " + ",".join(sorted(eng._ring)) if eng._ring else "partitioned")),
           "efficiency": "T(1) / (N T(N)) over the lines of a 1/2/4/8 run (each line reports its own T(N))"}
    del f, out, outs, plan, eng
    torch.cuda.empty_cache()
    return res


def bounded_block(world, rank, local, dev, peak, n, steps=5, warm=3):
    """The operators with one-sided closures on every axis (BC NONE), n^3 per GPU: the table-coefficient chunks."""
    import torch
    ops = OPS + ("ring", "laplacian")
    plan, eng = build_engine(world, local, n, n, n * world, False)
    f = synthetic_device_field(plan, rank, dev)
    out = plan.empty_device()
    outs = {name: out for name in ops}
    ms, per_op_ms, _ = timed_block(plan, eng, f, outs, ops, steps, warm, world, dev)
    res = {"config": "BC NONE on every axis, %d^3 per GPU" % n, "ms_per_step": ms, "per_op": {}}
    for name, t in per_op_ms.items():
        sweeps = 3 if name in ("sfilter", "gfilter", "ring", "laplacian") else 1
        gbs = sweeps * 16 * plan.npts / (t * 1e-3) / 1e9
        res["per_op"][name] = {"ms": t, "sweeps": sweeps, "algorithmic_GBps_per_gpu": gbs, "frac_of_hbm_peak": gbs / peak}
        if name in ("ring", "laplacian"):
            # the second and third sweep accumulate into the output through TMA reduce stores (max / add): the memory
            # system reads the old value, so the three sweeps move 16 + 24 + 24 bytes per point, not 48
            moved = 64 * plan.npts / (t * 1e-3) / 1e9
            res["per_op"][name].update({"moved_bytes_per_point": 64, "moved_GBps_per_gpu": moved, "moved_frac_of_hbm_peak": moved / peak})
    del f, out, outs, plan, eng
    torch.cuda.empty_cache()
    return res


def parity_block(world, rank, local, dev):
    """The same engine on a 64 x 64 x (128 N) field, periodic and bounded, against the CPU oracle (the
    checker; never on the measured path): max relative L-infinity over the five operators."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import oracle
    oracle.build()
    worst, per_op, paths = 0.0, {}, {}
    for periodic in (True, False):
        n = (64, 64, 128 * world)
        ext = [(2 * np.pi * (k - 1) / k) if periodic else 1.0 for k in n]
        o = oracle.Oracle(*n, 0, ext[0], 0, ext[1], 0, ext[2], periodic=(periodic,) * 3)
        x, y, z = o.getvar("x"), o.getvar("y"), o.getvar("z")
        rng = np.random.default_rng(1234)
        fh = np.asfortranarray(np.sin(3 * x) * np.cos(2 * y) * np.cos(z) + 0.1 * rng.uniform(-1, 1, size=x.shape))
        plan, eng = build_engine(world, local, *n, periodic)
        az = n[2] // world
        sl = slice(rank * az, (rank + 1) * az)
        f = plan.empty_device()
        f.copy_(torch.from_numpy(np.ascontiguousarray(fh[:, :, sl])).to(dev))
        for name in OPS:
            got = (eng.apply(name, f) if eng is not None else plan.apply(name, f)).cpu().numpy()
            ref = getattr(o, name)(fh)[:, :, sl]
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            key = name + ("" if periodic else "_bounded")
            per_op[key] = err
            worst = max(worst, err)
        paths["periodic" if periodic else "bounded"] = ("one rank" if eng is None else
                                                        ("fused: " + ",".join(sorted(eng._ring)) if eng._ring else "partitioned"))
        del plan, eng, f
    t = torch.tensor([worst] + [per_op[k] for k in sorted(per_op)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"grid": [64, 64, 128 * world], "max_rel": t[0].item(), "per_op": {k: t[1 + j].item() for j, k in enumerate(sorted(per_op))},
            "tolerance": 1e-12, "against": "oracle/parcop_oracle.c (CPU restatement of the reference), same inputs", "z_path": paths}


def other_configs(world, rank):
    """BASELINE configs[3] (RT3D, bounded x, z-slab over the ranks; 1.5 N x N x N with N = 512 on 8 GPUs, the same
    points per GPU -- N = 256 -- on one) and configs[4] (curvilinear cylinder 1024 x 512 x 64, one GPU), a few
    RK4 steps each through the deck interpreter (tools/run_configs.py).  Failures are reported, not fatal."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    out = {}
    try:
        import run_configs
        npts = {1: 256, 8: 512}.get(world)
        if npts is not None:
            out["rt3d"] = run_configs.run_rt3d(npts, 2)
        if world == 1:
            torch.cuda.empty_cache()
            out["cylinder_curv"] = run_configs.run_cylinder(1024, 512, 64, 2)
    except Exception as exc:  # noqa: BLE001
        out["error"] = "%s: %s" % (type(exc).__name__, exc)
    torch.cuda.empty_cache()
    return out or None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=NPER, help="points per side per GPU (default 512)")
    ap.add_argument("--global-n", type=int, default=0,
                    help="fixed global N^3 grid split into z-slabs as the headline workload (strong scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tgv", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong_1024 block")
    ap.add_argument("--no-bounded", action="store_true", help="skip the per_op_bounded block")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[3] / configs[4] (RT3D, curvilinear cylinder)")
    ap.add_argument("--tgv-n", type=int, default=256)
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    warm = max(args.warmup, 3)
    n = args.n
    nx, ny, nz = n, n, n * world
    if args.global_n:
        if args.global_n % world or args.global_n // world < 16:
            raise SystemExit("--global-n %d cannot be split into %d z-slabs of at least 16 planes" % (args.global_n, world))
        nx = ny = nz = args.global_n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peaks()
    plan, eng = build_engine(world, local, nx, ny, nz, True)
    ax, ay, az = plan.shape
    npts = ax * ay * az
    f = synthetic_device_field(plan, rank, dev)
    outs = {name: plan.empty_device() for name in OPS}
    sampler = ClockSampler(local) if rank == 0 else None
    K = args.steps
    ms_per_step, per_op_ms, launches = timed_block(plan, eng, f, outs, OPS, K, warm, world, dev, sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = len(OPS) * npts * world / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end through the host-array API (the f2py call shape), pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        Ke = max(1, min(K, 4))
        hin = torch.empty((az, ay, ax), dtype=torch.float64).pin_memory()
        hin.copy_(f.permute(2, 1, 0))
        hout = torch.empty((az, ay, ax), dtype=torch.float64).pin_memory()
        a_in = hin.numpy().T  # Fortran-ordered (ax, ay, az) views of the pinned buffers
        a_out = hout.numpy().T

        def e2e_step():
            for name in OPS:
                if eng is not None:
                    eng.apply_host_into(name, a_in, a_out)
                else:
                    plan.apply_host_into(name, a_in, a_out)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / Ke], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": len(OPS) * npts * world / dt.item() / 1e9, "unit": "Gpoints/s",
               "h2d_bytes_per_step": len(OPS) * npts * 8 * world, "d2h_bytes_per_step": len(OPS) * npts * 8 * world,
               "steps": Ke, "ms_per_step": dt.item() * 1e3,
               "note": "pb_host_apply per operator: pinned host array -> H2D -> kernels -> D2H -> pinned host array"}
        del hin, hout
    z_path = "one rank" if eng is None else ("fused ring kernel: " + ",".join(sorted(eng._ring)) if eng._ring else "partitioned")
    del f, outs, plan, eng
    torch.cuda.empty_cache()

    # ---- the other driver-visible blocks: every rank takes part, rank 0 reports ----
    bounded = None if args.no_bounded else bounded_block(world, rank, local, dev, peak, n)
    strong = None if (args.no_strong or args.global_n) else strong_1024_block(world, rank, local, dev, peak)
    parity = None if args.no_parity else parity_block(world, rank, local, dev)
    configs = None
    if not args.no_configs:
        configs = other_configs(world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    per_op = per_op_table(per_op_ms, npts, world, peak)
    # dominant kernel: the fused y/z sweep (6 of the 9 sweeps of a step); one launch == ddz at N = 1, ddy on a z-slab
    dom = "ddz" if world == 1 else "ddy"
    ach = 16.0 * npts / (per_op_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "sweep_yz_pipe_kernel<D1,16> (%s, one launch per application)" % dom, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_DDZ_512 if (n == 512 and world == 1 and not args.global_n) else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the sweep_yz_pipe_kernel<D1,16> launch in profiles/r2_pipe_and_ring_kernels_ncu_full.txt",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 16 * npts}
    tgv = None
    if world == 1 and not args.no_tgv:
        tgv = tgv_step_time(args.tgv_n)
    cpu = None if (args.no_cpu or world > 1) else cpu_sample(256, 1)  # N = 1 only (contract)
    if cpu:
        cpu.pop("seconds_per_step", None)
        if tgv is not None:
            tgv["cpu_baseline"] = tgv_cpu_sample()
    cfg = bench_config(world, n, args.global_n)
    cfg["z_path"] = z_path
    line = {"metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": K, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "per_op": per_op, "per_op_bounded": bounded, "strong_1024": strong, "parity": parity, "configs": configs, "tgv": tgv,
            "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
