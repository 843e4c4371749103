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
    ... bench.py --gpus N --global-n 1024   # BASELINE configs[2]: fixed 1024^3 grid, strong scaling
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
# ncu --set full, 512^3 ddz launch: 1.073867 GB read + 1.035293 GB written (profiles/r1_v7_tma_kernels_ncu_full.txt)
NCU_TRAFFIC_DDZ_512 = 2109160000


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


def cpu_sample(n=256, reps=1):
    """Time the CPU oracle (port of the reference's algorithm, OpenMP over line bundles) on a bounded
    sample of the same workload: the five operators on an n^3 periodic field."""
    import numpy as np
    from oracle import oracle
    oracle.build()
    L = 2 * np.pi * (n - 1) / n
    o = oracle.Oracle(n, n, n, 0, L, 0, L, 0, L, periodic=(True,) * 3)
    x, y, z = o.getvar("x"), o.getvar("y"), o.getvar("z")
    rng = np.random.default_rng(1234)
    f = np.asfortranarray(np.sin(3 * x) * np.cos(2 * y) * np.cos(z) + 0.1 * rng.uniform(-1, 1, size=x.shape))
    for name in OPS:  # warm-up
        getattr(o, name)(f)
    t0 = time.perf_counter()
    for _ in range(reps):
        for name in OPS:
            getattr(o, name)(f)
    dt = (time.perf_counter() - t0) / reps
    return {"value": len(OPS) * n ** 3 / dt / 1e9, "unit": "Gpoints/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "the same 5 operators on a %d^3 periodic fp64 field (1/%d of the GPU workload), %d rep(s), host arrays" % (n, (NPER // n) ** 3, reps),
            "seconds_per_step": dt}


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
    n = 256
    for _ in range(max(args.warmup - 1, 0)):
        cpu_sample(n, 1)
    t0 = time.perf_counter()
    res = [cpu_sample(n, 1) for _ in range(max(1, args.steps))]
    wall = time.perf_counter() - t0
    sec = sum(r["seconds_per_step"] for r in res) / len(res)
    value = len(OPS) * n ** 3 / sec / 1e9
    cb = dict(res[0]); cb["value"] = value; cb.pop("seconds_per_step", None)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args.gpus, args.n),
            "note": "the Fortran/MPI reference cannot be built in this image (no Fortran compiler, no MPI); this is the oracle "
                    "port of its algorithm on all host threads, each step a %d^3 sample of the workload" % n,
            "cpu_baseline": cb, "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=NPER, help="points per side per GPU (default 512)")
    ap.add_argument("--global-n", type=int, default=0,
                    help="fixed global N^3 grid split into z-slabs (strong scaling; BASELINE configs[2] is N = 1024)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tgv", action="store_true")
    ap.add_argument("--tgv-n", type=int, default=256)
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pyranda_b200 import ParcopPlan
    from pyranda_b200._lib import OP

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
    Lx, Ly, Lz = (2 * np.pi * (k - 1) / k for k in (nx, ny, nz))

    if world > 1:
        from pyranda_b200.distributed import DistributedParcop
        eng = DistributedParcop(nx, ny, nz, 0, Lx, 0, Ly, 0, Lz, periodic=(True,) * 3, device=local)
        plan = eng.plan
    else:
        plan = ParcopPlan(nx, ny, nz, 0, Lx, 0, Ly, 0, Lz, periodic=(True,) * 3, device=local)
        eng = None
    plan.set_mesh()
    ax, ay, az = plan.shape
    npts = ax * ay * az

    # synthetic field of SURVEY 8d, generated on the device (its own slab of the global field)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    xi = torch.arange(ax, dtype=torch.float64, device=dev) * plan.dx
    yi = torch.arange(ay, dtype=torch.float64, device=dev) * plan.dy
    zi = (torch.arange(az, dtype=torch.float64, device=dev) + rank * az) * plan.dz
    f = plan.empty_device()
    f.copy_(torch.sin(3 * xi).view(ax, 1, 1) * torch.cos(2 * yi).view(1, ay, 1) * torch.cos(zi).view(1, 1, az))
    noise = torch.rand((az, ay, ax), dtype=torch.float64, device=dev, generator=g).permute(2, 1, 0)
    f.add_(0.2 * noise - 0.1)
    del noise
    outs = {name: plan.empty_device() for name in OPS}
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
        for name in OPS:
            one_op(name)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    K = args.steps
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in OPS] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = plan.launch_count()
    barrier()
    e0.record()
    for k in range(K):
        for j, name in enumerate(OPS):
            ev[k][j][0].record()
            one_op(name)
            ev[k][j][1].record()
    e1.record()
    barrier()
    launches = plan.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    per_op_ms = {name: sum(ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(K)) / K for j, name in enumerate(OPS)}
    t = torch.tensor([total_ms] + [per_op_ms[nme] for nme in OPS], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t[0].item()
    per_op_ms = {name: t[1 + j].item() for j, name in enumerate(OPS)}
    ms_per_step = total_ms / K
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    per_op = {}
    for name in OPS:
        gbs = BYTES_PER_POINT[name] * npts / (per_op_ms[name] * 1e-3) / 1e9
        per_op[name] = {"ms": per_op_ms[name], "gpoints_per_s": npts * world / (per_op_ms[name] * 1e-3) / 1e9,
                        "algorithmic_GBps_per_gpu": gbs, "frac_of_hbm_peak": gbs / peak}
    # dominant kernel: the fused y/z sweep (6 of the 9 sweeps of a step); one launch == ddz
    dom = "ddz" if world == 1 else "ddy"
    ach = 16.0 * npts / (per_op_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "sweep_yz_pipe_kernel<D1,16> (%s, one launch per application)" % dom, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_DDZ_512 if (n == 512 and world == 1 and not args.global_n) else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the ddz launch in profiles/r1_v7_tma_kernels_ncu_full.txt",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 16 * npts}
    tgv = None
    if world == 1 and not args.no_tgv:
        tgv = tgv_step_time(args.tgv_n)
    cpu = None if (args.no_cpu or world > 1) else cpu_sample(256, 1)  # N = 1 only (contract)
    if cpu:
        cpu.pop("seconds_per_step", None)
    line = {"metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": K, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(world, n, args.global_n),
            "per_op": per_op, "tgv": tgv, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
