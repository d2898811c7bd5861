#!/usr/bin/env python
"""bench.py — headline benchmark of the Physis b200 backend.

Metric (BASELINE.json): 7-pt diffusion GLUP/s (+ Himeno GLUP/s as an extra key)
and the achieved fraction of the measured HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" is one full pass of the reference benchmark's hot path over one
synthetic input: PSStencilRun of `--count` (default 1000, BASELINE config 2)
7-pt sweeps on a 512^3 fp32 grid per GPU.
  value : device-resident throughput (inputs already in HBM), all ranks
  e2e   : the same through run_kernel_physis(): PSGridCopyin from pinned host
          memory + the sweeps + PSGridCopyout, copies inside the timed region
Timing is on the device (CUDA events on the runtime's stream, through the
C ABI), max over ranks; the 1 GiB working set is far larger than the 126 MB L2.
Extra keys on the same line: `himeno` (BASELINE config 3: XL, sweep-only and with
the per-sweep residual + PSReduce), `periodic_staggered_fp64` (config 5) and
`strong_scaling_1024` (config 4: a fixed 1024^3 grid cut over the N ranks).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(kernel):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version
    banner at communicator creation, ...) is sent to stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


_ORIG_AFFINITY = None


def _bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask), so that its pinned host
    buffers are allocated in that socket's memory: with 8 ranks copying at once the host side of
    the PCIe transfers is the bottleneck of `e2e`.  Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < ncpu:
            global _ORIG_AFFINITY
            _ORIG_AFFINITY = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} of {ncpu} CPUs (GPU {local_rank}'s NUMA node)"
        return "no narrower affinity reported"
    except Exception as e:  # noqa: BLE001 -- NVML absent or not permitted: run unbound
        return f"unbound ({type(e).__name__})"


def _dist_setup(ngpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    return rank, world, dist


def _max_over_ranks(dist, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


# ------------------------------------------------------------------ reference arm

def _ref_target_glups(lib, n, sweeps, reps=1, warm=0):
    """The Physis REFERENCE target (libphysis_rt_ref + translator-shaped sweep): sequential code."""
    import helpers as H
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.initialize_physis(0, None, n, n, n)
    lib.initialize_benchmark_physis(n, n, n)
    lib.copyin_physis.argtypes = [C.c_void_p]
    lib.copyin_physis(f0.ctypes.data)
    lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
    co = [float(c) for c in p[:7]]
    for _ in range(warm):
        lib.run_sweeps_only_physis(sweeps, n, n, n, *co)
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.run_sweeps_only_physis(sweeps, n, n, n, *co)
    dt = time.perf_counter() - t0
    lib.finalize_benchmark_physis()
    return n ** 3 * sweeps * reps / dt / 1e9, dt


def _ref_openmp_glups(lib, n, sweeps, reps=1, warm=0):
    """The reference's own multi-threaded CPU form of the same sweep
    (examples/diffusion-benchmark/diffusion3d_openmp.cc, unmodified, all host threads)."""
    import helpers as H
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    lib.ref_openmp_load.argtypes = [C.c_int] * 3 + [C.c_void_p]
    lib.ref_openmp_store.argtypes = [C.c_void_p]
    try:  # launchers (torchrun) pin OMP_NUM_THREADS=1; this arm is meant to use every host core
        C.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count() or 1)
    except OSError:
        pass
    lib.ref_openmp_load(n, n, n, f0.ctypes.data)
    for _ in range(warm):
        lib.ref_openmp_sweeps(sweeps)
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.ref_openmp_sweeps(sweeps)
    dt = time.perf_counter() - t0
    lib.ref_openmp_store(f0.ctypes.data)
    return n ** 3 * sweeps * reps / dt / 1e9, dt, int(lib.ref_openmp_threads())


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores: its OpenMP
    form of the sweep on all host threads (the headline of this arm), with the Physis
    REFERENCE target (sequential by construction, translator/reference_runtime_builder.cc:605-664)
    beside it.  Falls back to the oracle port (1 thread) where oracle/_ref was never built."""
    if rank != 0:
        return
    import helpers as H
    n = args.size
    lib = H.oracle_ref()
    extra = {}
    if lib is not None:
        sweeps = 20   # bounded sample: ~0.1 s per 512^3 sweep on 16 cores
        glups, dt, threads = _ref_openmp_glups(lib, n, sweeps, reps=max(args.steps, 1), warm=args.warmup)
        kind, cores = "reference", threads
        sample = (f"{sweeps} sweeps of {n}^3 per step, the reference's diffusion3d_openmp.cc "
                  f"on {threads} host threads")
        ms_per_step = dt / max(args.steps, 1) * 1e3
        g1, _ = _ref_target_glups(lib, min(n, 256), 4)
        extra["ref_target_1thread"] = {"value": g1, "unit": "GLUP/s", "cores": 1,
                                       "sample": f"4 sweeps of {min(n, 256)}^3, Physis REFERENCE target"}
    else:
        lib = H.oracle_port()
        sweeps = 2
        glups, dt = _ref_target_glups(lib, n, sweeps, reps=max(args.steps, 1), warm=args.warmup)
        kind, cores = "port", 1
        sample = f"{sweeps} sweeps of {n}^3 per step, CPU restatement of the REFERENCE target, 1 thread"
        ms_per_step = dt / max(args.steps, 1) * 1e3
    cpu = {"value": glups, "unit": "GLUP/s", "cores": cores, "kind": kind, "sample": sample,
           "host_cores": os.cpu_count()}
    cpu.update(extra)
    line = {
        "impl": "reference", "metric": "7-pt diffusion GLUP/s", "value": glups, "unit": "GLUP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"7-pt 3D diffusion fp32 {n}^3 (BASELINE config 2 shape), "
                               f"bounded sample of {sweeps} sweeps/step on host CPU"},
        "cpu_baseline": cpu,
        "e2e": {"value": glups, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------ b200 arm

def cpu_baseline_sample():
    """Config 1 (256^3, 100 sweeps) on the host cores: the reference's OpenMP form on all
    threads and the Physis REFERENCE target on one (its codegen is sequential)."""
    import helpers as H
    if _ORIG_AFFINITY is not None:   # the CPU baseline gets every host core back
        os.sched_setaffinity(0, _ORIG_AFFINITY)
    n, sweeps = 256, 100
    lib = H.oracle_ref()
    if lib is None:
        glups, dt = _ref_target_glups(H.oracle_port(), n, sweeps)
        return {"value": glups, "unit": "GLUP/s", "cores": 1, "kind": "port",
                "host_cores": os.cpu_count(), "seconds": dt,
                "sample": f"BASELINE config 1: {n}^3 fp32, {sweeps} sweeps, CPU restatement of the "
                          "REFERENCE target, 1 thread"}
    g1, dt1 = _ref_target_glups(lib, n, sweeps)
    gomp, dto, threads = _ref_openmp_glups(lib, n, sweeps, reps=3, warm=1)
    return {"value": gomp, "unit": "GLUP/s", "cores": threads, "kind": "reference",
            "host_cores": os.cpu_count(), "seconds": dt1 + dto,
            "sample": f"BASELINE config 1: {n}^3 fp32, {sweeps} sweeps x3, the reference's "
                      f"diffusion3d_openmp.cc on {threads} host threads",
            "ref_target_1thread": {"value": g1, "unit": "GLUP/s", "cores": 1, "seconds": dt1,
                                   "sample": f"{n}^3 fp32, {sweeps} sweeps, Physis REFERENCE target "
                                             "(libphysis_rt_ref + translator-shaped sweep; REF codegen "
                                             "is sequential)"}}


def himeno_line(args, api, lib, world, dist):
    """Extra: Himeno XL (1024x512x512 per GPU, z-slabs) sweep-only and with the residual emitted
    every sweep + PSReduce.  Weak scaling: the global grid is 1024 x 512 x (512*world)."""
    mi, mj, mk = (1024, 512, 512) if args.himeno == "XL" else (512, 256, 256)
    gmk = mk * world
    lib.himeno_init_local.argtypes = [C.c_int] * 3
    lib.himeno_init_local(mi, mj, gmk)
    nn = args.himeno_nn
    lib.himeno_sweeps_only.argtypes = [C.c_int, C.c_int]
    lib.himeno_reduce_gosa.restype = C.c_float
    out = {}
    pts = (mi - 2) * (mj - 2) * (gmk - 2)
    peak, _ = _peaks()
    lib.himeno_jacobi_gosa_each.argtypes = [C.c_int]
    lib.himeno_jacobi_gosa_each.restype = C.c_float
    for with_gosa in (0, 1):
        # with_residual: the original benchmark's structure -- the residual of every iteration:
        # each PSStencilRun of the ping-pong pair (two sweeps, ss*ss emitted) is followed by
        # PSReduce(&gosa, PS_SUM, gosa_g), which folds the per-CTA partial sums the sweep left
        run = (lambda: lib.himeno_jacobi_gosa_each(nn)) if with_gosa else (lambda: lib.himeno_sweeps_only(nn, 0))
        for _ in range(2):
            run()
        api.rt().__PSB200Synchronize()
        _barrier(dist)
        api.rt().__PSB200ResetStats()
        api.rt().__PSB200TimerStart()
        gosa = run()
        ms = api.rt().__PSB200TimerStopMs()
        _barrier(dist)
        ms = _max_over_ranks(dist, ms)
        # 12 coefficient/source reads + p read + p write (himenobmtxpa_physis.c:418-432); with the
        # residual, + the 4-byte ss*ss emit of the second sweep of every pair (the first one's would be
        # overwritten before anything reads it, so that sweep runs in its plain form): 58 on
        # average; the reduction reads one fp64 per CTA, not the grid
        bpl = 58 if with_gosa else 56
        key = "with_residual" if with_gosa else "sweep_only"
        gbs = pts * nn * bpl / ms / 1e6
        out[key] = {"glups": pts * nn / ms / 1e6, "ms_per_sweep": ms / nn,
                    "alg_bytes_per_lup": bpl, "gbs": gbs, "roofline_frac_per_gpu": gbs / world / peak}
        if with_gosa:
            st = api.stats()
            # the same sum by a full pass over the emitted grid (GPU, fp32 tree)
            api.set_option("reduce_fuse=0")
            full = float(lib.himeno_reduce_gosa())
            api.set_option("reduce_fuse=1")
            out[key].update({"gosa": float(gosa), "gosa_full_pass": full,
                             "gosa_ok": bool(abs(float(gosa) - full) <= 2e-5 * abs(full)),
                             "reduces": nn // 2, "reduces_from_partials": int(st.reduces_from_partials),
                             "schedule": "PSStencilRun(pair, 1) + PSReduce per iteration"})
    # the original benchmark's OUTPUT is the residual of the last sweep only: one PSStencilRun of all
    # the sweeps (ss*ss emitted by every sweep of the DSL program; fused passes skip the emits later
    # sweeps overwrite) and one PSReduce at the end
    lib.himeno_jacobi_gosa.argtypes = [C.c_int]
    lib.himeno_jacobi_gosa.restype = C.c_float
    for _ in range(2):
        lib.himeno_jacobi_gosa(nn)
    api.rt().__PSB200Synchronize()
    _barrier(dist)
    api.rt().__PSB200ResetStats()
    api.rt().__PSB200TimerStart()
    gosa_end = lib.himeno_jacobi_gosa(nn)
    ms = api.rt().__PSB200TimerStopMs()
    _barrier(dist)
    ms = _max_over_ranks(dist, ms)
    st = api.stats()
    out["residual_at_end"] = {"glups": pts * nn / ms / 1e6, "ms_per_sweep": ms / nn, "gosa": float(gosa_end),
                              "fused_passes": int(st.fused_pairs),
                              "reduces_from_partials": int(st.reduces_from_partials),
                              "schedule": "one PSStencilRun of all sweeps + one PSReduce"}
    lib.himeno_finalize()
    out["size"] = f"{mi}x{mj}x{gmk} over {world} GPU(s)"
    out["sweeps"] = nn
    return out


def strong_line(args, api, lib, world, dist):
    """Extra: BASELINE config 4 -- 7-pt diffusion fp32 on a FIXED 1024^3 grid cut into z-slabs over
    the ranks (strong scaling; at N=1 the whole grid on one GPU).  Rows of 1024 floats are wider
    than the fused tile, so this runs sweep by sweep: 8 B/LUP."""
    n = args.strong_size
    sweeps = args.strong_count
    lib.initialize_physis(0, None, n, n, n)
    lib.initialize_benchmark_physis(n, n, n)
    zo, zl = C.c_int(), C.c_int()
    lib.local_size_physis(C.byref(zo), C.byref(zl))
    i = np.arange(n, dtype=np.float64)
    ax = (1.0 - np.cos(2 * np.pi * (i + 0.5) / n)).astype(np.float32)
    az = ax[zo.value:zo.value + zl.value]
    plane = (0.125 * ax[:, None] * ax[None, :]).astype(np.float32)
    host, host_ptr = api.pinned_empty(zl.value * n * n * 4, np.float32)
    hv = host.reshape(zl.value, n, n)
    for k in range(zl.value):
        np.multiply(plane, az[k], out=hv[k])
    lib.copyin_local_physis(host.ctypes.data)
    co = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.4]
    r = api.rt()
    lib.run_sweeps_only_physis(20, n, n, n, *co)
    r.__PSB200Synchronize()
    _barrier(dist)
    r.__PSB200ResetStats()
    r.__PSB200TimerStart()
    lib.run_sweeps_only_physis(sweeps, n, n, n, *co)
    ms = r.__PSB200TimerStopMs()
    _barrier(dist)
    ms = _max_over_ranks(dist, ms)
    st = api.stats()
    lib.finalize_benchmark_physis()
    r.__PSB200HostFree(C.c_void_p(host_ptr))
    peak, _ = _peaks()
    gbs = n ** 3 * sweeps * 8 / ms / 1e6
    return {"glups": n ** 3 * sweeps / ms / 1e6, "ms_per_sweep": ms / sweeps, "alg_bytes_per_lup": 8,
            "gbs": gbs, "roofline_frac_per_gpu": gbs / world / peak,
            "size": f"{n}^3 over {world} GPU(s) ({n}x{n}x{zl.value} z-slab per GPU)", "sweeps": sweeps,
            "scaling": "strong", "fused_passes": int(st.fused_pairs)}


def pstag_line(args, api, lib, world, dist):
    """Extra: BASELINE config 5 -- fp64 periodic 7-pt on a user type {p,q} with a staggered
    coefficient grid, 512^3 cells per GPU (weak scaling), device SoA: 24 B/LUP."""
    n = args.pstag_size
    gnz = n * world
    lib.pstag_init.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.pstag_init(0, None, n, n, gnz)
    uo, ul, ko, kl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.pstag_local_size(C.byref(uo), C.byref(ul), C.byref(ko), C.byref(kl))
    rng = np.random.default_rng(7)
    u = np.zeros((ul.value * n * n, 2), np.float64)
    u[:, 0] = rng.random(ul.value * n * n)
    kap = np.full(kl.value * (n + 1) * (n + 1), 0.05, np.float64)
    lib.pstag_copyin_local.argtypes = [C.c_void_p, C.c_void_p]
    lib.pstag_copyin_local(u.ctypes.data, kap.ctypes.data)
    lib.pstag_sweeps_only.argtypes = [C.c_int] * 4
    count = args.pstag_count
    for _ in range(2):
        lib.pstag_sweeps_only(count, n, n, gnz)
    api.rt().__PSB200Synchronize()
    _barrier(dist)
    api.rt().__PSB200TimerStart()
    lib.pstag_sweeps_only(count, n, n, gnz)
    ms = api.rt().__PSB200TimerStopMs()
    _barrier(dist)
    ms = _max_over_ranks(dist, ms)
    lib.pstag_finalize()
    pts = n * n * gnz
    peak, _ = _peaks()
    gbs = pts * count * 24 / ms / 1e6
    return {"glups": pts * count / ms / 1e6, "ms_per_sweep": ms / count, "alg_bytes_per_lup": 24,
            "gbs": gbs, "roofline_frac_per_gpu": gbs / world / peak,
            "size": f"{n}x{n}x{gnz} cells over {world} GPU(s)", "sweeps": count, "dtype": "f64"}


def expected_samples(n, gnz, z_off, idx, sweeps, co):
    """Exact solution of the DISCRETE problem at flat slab indices `idx` after `sweeps` sweeps:
    the benchmark's initial field 0.125 * prod_d (1 - cos(theta_d)) is a sum of cell-centred cosine
    modes, each an eigenvector of the clamped 7-point update, so the field after T sweeps is
    0.125 * sum over subsets S of axes of (-1)^|S| * lambda_S^T * prod_{d in S} cos(theta_d),
    lambda_S = cc + sum_d (c_d- + c_d+) * (cos(2 pi / n_d) if d in S else 1).  fp64; independent of
    the oracle."""
    ce, cw, cn, cs, ct, cb, cc = [float(np.float32(c)) for c in co]
    i = idx % n
    j = (idx // n) % n
    k = idx // (n * n) + z_off
    th = [2 * np.pi * (i + 0.5) / n, 2 * np.pi * (j + 0.5) / n, 2 * np.pi * (k + 0.5) / gnz]
    pair = [ce + cw, cn + cs, ct + cb]
    step = [np.cos(2 * np.pi / n), np.cos(2 * np.pi / n), np.cos(2 * np.pi / gnz)]
    out = np.zeros(idx.size, np.float64)
    for mask in range(8):
        lam, term, sign = cc, np.ones(idx.size, np.float64), 1.0
        for d in range(3):
            if mask >> d & 1:
                lam += pair[d] * step[d]
                term = term * np.cos(th[d])
                sign = -sign
            else:
                lam += pair[d]
        out += sign * (lam ** sweeps) * term
    return 0.125 * out


def parity_check(world, rank, dist):
    """After the timed regions: one small fixed case per kernel family, run across the ranks
    through the same C ABI and compared bit for bit with the CPU oracle (the checker, never the
    thing measured).  Returns (ok on every rank, sha256 of rank 0's GPU results, case names)."""
    import hashlib
    import helpers as H
    port, prog = H.oracle_port(), H.b200_programs()
    h = hashlib.sha256()
    ok, cases = True, []

    def same(name, want, got):
        nonlocal ok
        w, g = np.ascontiguousarray(want), np.ascontiguousarray(got)
        good = w.tobytes() == g.tobytes()
        ok = ok and good
        cases.append(name if good else name + " MISMATCH")
        h.update(g.tobytes())

    rng = np.random.default_rng(1234)
    iso = np.array([0.1] * 6 + [0.4], np.float32)
    gen = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32)
    for (nx, ny, nz), count, co in [((256, 48, 8 * world), 6, iso), ((512, 20, 8 * world), 8, gen),
                                    ((1024, 24, 8 * world), 6, iso)]:
        f0 = rng.random(nx * ny * nz, dtype=np.float32)
        same(f"diffusion7 {nx}x{ny}x{nz} x{count}", H.run_diffusion(port, f0, nx, ny, nz, count, co),
             H.run_diffusion(prog, f0, nx, ny, nz, count, co))
    dims = (128, 24, 4 * world + 2)
    a = H.run_himeno(port, dims, 4, gosa=True, seed=5, each=True)
    b = H.run_himeno(prog, dims, 4, gosa=True, seed=5, each=True)
    same(f"himeno19 {dims} p0", a[0], b[0])
    same(f"himeno19 {dims} p1", a[1], b[1])
    same(f"himeno19 {dims} ss^2", a[3], b[3])
    exact = float(np.sum(a[3].astype(np.float64)))
    good = abs(b[2] - exact) <= 8e-6 * abs(exact)
    ok = ok and good
    cases.append("himeno19 gosa vs fp64 sum" + ("" if good else " MISMATCH"))
    nx, ny, nz = 128, 32, 4 * world
    u, kap = H.pstag_inputs(nx, ny, nz)
    same(f"periodic7_staggered {nx}x{ny}x{nz} x3", H.run_pstag(port, u, kap, nx, ny, nz, 3),
         H.run_pstag(prog, u, kap, nx, ny, nz, 3))
    if dist is not None:
        import torch
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() > 0.5)
    return ok, h.hexdigest(), cases


def small_runs_line(args, api, lib, world, dist):
    """Extra: BASELINE config 1's grid (256^3 per GPU, 100 sweeps) on the GPU -- through one
    PSStencilRun(50 iterations) and through the common idiom of 50 PSStencilRun(1 iteration)
    calls (no fused passes: one run call covers two sweeps; prepared plans are reused)."""
    n, sweeps = 256, 100
    gnz = n * world
    co = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.4]
    lib.initialize_physis(0, None, n, n, gnz)
    lib.initialize_benchmark_physis(n, n, gnz)
    zo, zl = C.c_int(), C.c_int()
    lib.local_size_physis(C.byref(zo), C.byref(zl))
    i = np.arange(n, dtype=np.float64)
    ax = (1.0 - np.cos(2 * np.pi * (i + 0.5) / n)).astype(np.float32)
    k = np.arange(zo.value, zo.value + zl.value, dtype=np.float64)
    az = (1.0 - np.cos(2 * np.pi * (k + 0.5) / gnz)).astype(np.float32)
    f0 = (0.125 * az[:, None, None] * ax[None, :, None] * ax[None, None, :]).astype(np.float32).ravel()
    lib.copyin_local_physis(f0.ctypes.data)
    lib.run_sweeps_iter1_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
    r = api.rt()
    out = {}
    for key, fn in (("one_run_call", lib.run_sweeps_only_physis), ("iter1_loop", lib.run_sweeps_iter1_physis)):
        for _ in range(3):
            fn(sweeps, n, n, gnz, *co)
        r.__PSB200Synchronize()
        _barrier(dist)
        r.__PSB200ResetStats()
        t0 = time.perf_counter()
        r.__PSB200TimerStart()
        reps = 5
        for _ in range(reps):
            fn(sweeps, n, n, gnz, *co)
        host_ms = (time.perf_counter() - t0) * 1e3   # time to ENQUEUE (the host never synchronises)
        ms = r.__PSB200TimerStopMs()
        _barrier(dist)
        ms = _max_over_ranks(dist, ms)
        st = api.stats()
        out[key] = {"glups": n * n * gnz * sweeps * reps / ms / 1e6, "ms_per_sweep": ms / (sweeps * reps),
                    "host_enqueue_ms_per_sweep": host_ms / (sweeps * reps),
                    "launches_per_100_sweeps": int(st.kernel_launches) // reps,
                    "plan_cache_hits": int(st.plan_cache_hits)}
    lib.finalize_benchmark_physis()
    out["size"] = f"{n}x{n}x{gnz} over {world} GPU(s), {sweeps} sweeps (BASELINE config 1's grid per GPU)"
    out["l2"] = "working set 134 MB per GPU ~ L2 (126 MB): partly L2-resident, not an HBM roofline case"
    return out


def run_b200(args, rank, world, dist):
    import physis_b200
    from physis_b200 import api
    import helpers as H

    lib = physis_b200.load_programs()
    n, count = args.size, args.count
    # weak scaling: n^3 points per GPU, the global grid is n x n x (n*world) cut into z-slabs
    # (strong scaling: --strong keeps the global grid at n^3)
    gnz = n if args.strong else n * world
    co = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.4]

    lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.initialize_physis(0, None, n, n, gnz)
    for kv in args.opt:
        api.set_option(kv)
    lib.initialize_benchmark_physis(n, n, gnz)
    lib.copyin_local_physis.argtypes = [C.c_void_p]
    lib.copyout_local_physis.argtypes = [C.c_void_p]
    lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
    lib.run_kernel_local_physis.argtypes = [C.c_int, C.c_void_p] + [C.c_int] * 3 + [C.c_float] * 7
    r = api.rt()
    zo, zl = C.c_int(), C.c_int()
    lib.local_size_physis(C.byref(zo), C.byref(zl))
    z_off, nz_loc = zo.value, zl.value
    npts_loc = n * n * nz_loc
    npts_glob = n * n * gnz

    # synthetic initial field of the benchmark's shape (smooth cosine product), this rank's slab
    def axis(m, lo=0, cnt=None):
        i = np.arange(lo, lo + (m if cnt is None else cnt), dtype=np.float64)
        return (1.0 - np.cos(2 * np.pi * (i + 0.5) / m)).astype(np.float32)
    ax, az = axis(n), axis(gnz, z_off, nz_loc)
    f0 = (0.125 * az[:, None, None] * ax[None, :, None] * ax[None, None, :]).astype(np.float32).ravel()

    host, host_ptr = api.pinned_empty(npts_loc * 4, np.float32)
    host[:] = f0
    lib.copyin_local_physis(host.ctypes.data)

    # ---- value: device-resident sweeps --------------------------------------
    for _ in range(args.warmup):
        lib.run_sweeps_only_physis(count, n, n, gnz, *co)
    r.__PSB200Synchronize()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    _barrier(dist)
    r.__PSB200ResetStats()
    # time_kernels: the run function brackets its fused passes with CUDA events on the
    # runtime's stream (one event synchronisation per PSStencilRun of `count` sweeps)
    api.set_option("time_kernels=1")
    r.__PSB200TimerStart()
    for _ in range(args.steps):
        lib.run_sweeps_only_physis(count, n, n, gnz, *co)
    ms = r.__PSB200TimerStopMs()
    api.set_option("time_kernels=0")
    _barrier(dist)
    st = api.stats()
    launches = int(st.kernel_launches)
    pairs = int(st.fused_pairs)
    pair_ms = float(st.fused_pair_ms) / max(int(st.fused_pairs_timed), 1)
    ms = _max_over_ranks(dist, ms)
    clocks = sampler.stop()
    value = npts_glob * count * args.steps / ms / 1e6  # GLUP/s, all ranks

    # ---- e2e: copyin (pinned host) + sweeps + copyout per step --------------
    for _ in range(min(args.warmup, 2)):
        lib.run_kernel_local_physis(count, host.ctypes.data, n, n, gnz, *co)
    host[:] = f0
    _barrier(dist)
    r.__PSB200TimerStart()
    for _ in range(args.steps):
        lib.run_kernel_local_physis(count, host.ctypes.data, n, n, gnz, *co)
    ms_e2e = r.__PSB200TimerStopMs()
    _barrier(dist)
    ms_e2e = _max_over_ranks(dist, ms_e2e)
    e2e = npts_glob * count * args.steps / ms_e2e / 1e6
    checksum = float(np.sum(host[::4097], dtype=np.float64))
    # ... against the exact solution of the discrete problem after count*steps sweeps (fp64)
    idx = np.arange(0, npts_loc, 4097, dtype=np.int64)
    want_sum = float(np.sum(expected_samples(n, gnz, z_off, idx, count * args.steps, co)))
    checksum_ok = abs(checksum - want_sum) <= 2e-4 * abs(want_sum)
    if dist is not None:
        checksum_ok = _max_over_ranks(dist, 0.0 if checksum_ok else 1.0) == 0.0

    # ---- roofline of the dominant kernel (per GPU) ------------------------------
    peak, peak_src = _peaks()
    if pairs:
        # dominant kernel: the fused two-sweep pass.  One launch advances every point by two
        # sweeps, so by the benchmark's own definition (8 B per point per sweep,
        # examples/diffusion-benchmark/diffusion3d.h:97-100) it accounts for 16 B per point;
        # temporal blocking moves about half of that through DRAM (`traffic`), which is how
        # `frac` exceeds 1: the kernel is no longer HBM-bound but fp32-issue-bound.
        kernel = "Star7PairKernel<float>"
        alg_bytes = 16 * npts_loc
        launch_ms = pair_ms
    else:
        kernel = "Star7Kernel<float>"
        alg_bytes = 8 * npts_loc              # 1 fp32 read + 1 fp32 write per point per launch
        launch_ms = ms / max(count * args.steps, 1)
    achieved = alg_bytes / launch_ms / 1e6     # GB/s
    traffic = _traffic(kernel)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel, "alg_bytes_per_launch": alg_bytes,
                "launch_ms": launch_ms, "peak_source": peak_src, "per": "GPU",
                "sweeps_per_launch": 2 if pairs else 1,
                "launches_timed": int(st.fused_pairs_timed) if pairs else count * args.steps}
    if traffic:
        roofline["dram_frac"] = traffic / launch_ms / 1e6 / peak
    if pairs:
        roofline["note"] = ("temporal blocking: one launch = two sweeps; algorithmic bytes follow the "
                            "benchmark's 8 B/point/sweep, DRAM traffic per launch is about half of them")

    lib.finalize_benchmark_physis()
    r.__PSB200HostFree(C.c_void_p(host_ptr))

    line = {
        "metric": "7-pt diffusion GLUP/s", "value": value, "unit": "GLUP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"7-pt 3D diffusion fp32 {n}x{n}x{gnz} over {world} GPU(s) "
                               f"({n}x{n}x{nz_loc} z-slab per GPU), {count} sweeps per step "
                               "(BASELINE config 2 per GPU)",
                   "l2": f"working set {2 * npts_loc * 4 / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                   "parallelism": f"z-slabs x{world}, halo planes stored peer-to-peer by the sweep",
                   "schedule": (f"{pairs // max(args.steps, 1)} fused two-sweep passes + "
                                f"{(launches - pairs) // max(args.steps, 1)} single sweeps per step"),
                   "options": args.opt, "cpu_binding": args.cpu_binding},
        "e2e": {"value": e2e, "unit": "GLUP/s", "h2d_bytes_per_step": npts_loc * 4 * world,
                "d2h_bytes_per_step": npts_loc * 4 * world, "ms_per_step": ms_e2e / args.steps,
                "checksum": checksum, "checksum_expected": want_sum, "checksum_ok": bool(checksum_ok)},
        "gpu_launches": launches,
        "roofline": roofline,
        "clocks": clocks,
    }
    if world > 1:
        line["halo_bytes_per_step"] = int(2 * n * n * 4 * count)
    if not args.no_himeno:
        h = himeno_line(args, api, lib, world, dist)
        ps = pstag_line(args, api, lib, world, dist)
        line["himeno"] = h
        line["periodic_staggered_fp64"] = ps
    if not args.no_strong and not args.strong and 1024 % world == 0:
        line["strong_scaling_1024"] = strong_line(args, api, lib, world, dist)
    if not args.no_small:
        line["config1_256"] = small_runs_line(args, api, lib, world, dist)
    if not args.no_parity:
        ok, sha, cases = parity_check(world, rank, dist)
        line["parity_ok"] = bool(ok and checksum_ok and
                                 line.get("himeno", {}).get("with_residual", {}).get("gosa_ok", True))
        line["parity"] = {"oracle": "oracle/liboracle.so (CPU restatement of the REFERENCE target)",
                          "bit_exact_cases": cases, "sha256_of_gpu_results": sha,
                          "checksum_vs_exact_discrete_solution": bool(checksum_ok), "ranks": world}
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_sample()
    if rank == 0:
        _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--count", type=int, default=1000)
    ap.add_argument("--opt", action="append", default=[], help="runtime option key=value")
    ap.add_argument("--no-himeno", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity cases against the oracle")
    ap.add_argument("--no-small", action="store_true", help="skip the 256^3 / iter=1-loop entry")
    ap.add_argument("--himeno", default="XL")
    ap.add_argument("--himeno-nn", type=int, default=20)
    ap.add_argument("--pstag-size", type=int, default=512)
    ap.add_argument("--pstag-count", type=int, default=100)
    ap.add_argument("--strong", action="store_true", help="keep the global grid at size^3 (strong scaling)")
    ap.add_argument("--no-strong", action="store_true", help="skip the extra config-4 (1024^3 strong scaling) entry")
    ap.add_argument("--strong-size", type=int, default=1024)
    ap.add_argument("--strong-count", type=int, default=100)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.cpu_binding = _bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world, dist = _dist_setup(args.gpus)
    run_b200(args, rank, world, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
