#!/usr/bin/env python
"""bench.py — spin-updates/s of the fused fp64 exchange + Zeeman LLG-Heun step with Langevin white noise.

    python bench.py --gpus N --steps K --warmup W                      (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference's CPU path on the host cores)

Workload (BASELINE.json configs[2], SURVEY.md 8d "C3"): simple cubic 256^3 (16 777 216 spins), NN Heisenberg
exchange J = 3.5e-21 J + Zeeman B = (0,0,1) T, alpha = 0.1, dt = 1e-4 ps, Langevin thermostat at T = 100 K
(Philox noise drawn inside both stage kernels), seeded uniform-on-sphere initial spins.  At N > 1 every rank owns a
256 x 256 x 256 x-slab of a (256 N) x 256 x 256 lattice (weak scaling; N = 8 is the 134 M spins of configs[4]) and the
boundary planes cross NVLink as peer stores from inside the stage kernels, ordered by epoch flags the same kernels poll
and publish (two launches per step at any N).

One "step" = one Heun step (predictor + corrector = two kernel launches) of every spin.
  value : device-resident throughput, K steps timed with CUDA events on the launching stream, max over ranks.
          Each spin array (403 MB per component set) is larger than L2 (126 MB), so no flush is needed.
  e2e   : the same K steps driven like the JAMS main loop (core/jams++.cc:333-341) through the plugin surface with
          HOST buffers: every `monitor_interval` steps the adapter imports globals::s from pinned host memory (H2D,
          the physics module may have rewritten it), runs the interval, exports the spins back into the pinned host
          array for the monitors (D2H) and reduces the magnetisation.  The reference's default interval is 100
          (helpers/defaults.h:28); here it is min(100, K / 4) so that a short run still crosses PCIe at least four
          times, and one untimed interval runs first.  `e2e_every_step` is the same with a monitor interval of 1 (the
          full state crosses PCIe in both directions around every step).
  roofline : the dominant launch (CUDA-event average over the timed region) against the measured HBM copy bandwidth;
          `step_frac` = 144 B x spins / ms_per_step / peak (SURVEY.md 8d's per-update figure).
  cpu_baseline / --impl reference : the reference's CPU path (oracle/_ref: the reference's own SparseMatrix::multiply +
          restated HeunLLGSolver::run) on a bounded sample, timed at 1 OpenMP thread and at all host threads; the faster
          of the two is the value (the reference's small OpenMP loops get slower with many threads on these boxes).
  reference_cuda : the reference's own CUDA path (cuSPARSE field + its Heun kernels, oracle/_ref/libjams_ref_cuda.so) on this
          GPU, on a bounded sample, with the product timed on the same lattice beside it (N = 1 only; a baseline like cpu_baseline)
The oracle (CPU restatement / reference-header builds) is executed only in the `cpu_baseline`, `reference_cuda` and `--impl reference` legs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spin-updates/sec (fp64 LLG-Heun+exchange)"
UNIT = "spin-updates/s"
BYTES_PER_UPDATE = 144.0          # SURVEY.md 8d / BASELINE.md 3: the two-stage data flow s, s*, u
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md
TEMPERATURE = 100.0
OUTPUT_STEPS = 100                # the reference's default monitor interval (helpers/defaults.h:28)
N_CELLS = 256
KERNEL_TIMING_EVERY = 4   # the stage launches of every 4th step of the timed region are bracketed by CUDA events (roofline.stage_ms)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        for key in ("hbm_gbs", "hbm_GBs", "hbm_copy_gbs"):
            if key in d:
                return float(d[key]), "measured (MEASURED_PEAKS.json %s)" % key
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.mem_samples, self.power_samples = [], []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
             0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:   # memory clock and board power: an HBM-bound kernel can slow down under sustained load at the full SM clock
                    self.mem_samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM)))
                    self.power_samples.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:  # noqa: BLE001
                    pass
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        out = {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples), "sm_mhz_min": int(min(self.samples))}
        if self.mem_samples:
            out["mem_mhz"] = int(np.median(self.mem_samples)); out["mem_mhz_min"] = int(min(self.mem_samples))
        if self.power_samples:
            out["power_w_max"] = float(max(self.power_samples))
        return out


def workload(n_ranks, dims=None):
    from jams_b200 import workloads as W
    dims = dims or (N_CELLS * n_ranks, N_CELLS, N_CELLS)
    return W.c3_sc(dims=dims, temperature=TEMPERATURE)


# ---------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref when it was built, else the restatement) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_worker(sample_n, steps, warmup):
    """runs in a subprocess whose OMP_NUM_THREADS the parent chose; prints one JSON line"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import build_cpu_sim
    which, kind = ("reference", "reference") if oracle.have_ref() else ("restatement", "port")
    w = workload(1, dims=(sample_n, sample_n, sample_n))
    lat = w["lattice"]
    sim = build_cpu_sim(w, which=which, dt_ps=1e-4, seed=1)
    sim.set_spins(lat.initial_spins(seed=1))
    threads = int(sim.L.omp_threads())
    sim.run(warmup)
    t0 = time.perf_counter()
    sim.run(steps)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(rate=lat.num_spins * steps / dt, s_per_step=dt / steps, threads=threads, kind=kind, spins=lat.num_spins)), flush=True)


def cpu_run(sample_n, steps, warmup, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-worker", "--cpu-sample", str(sample_n),
                        "--steps", str(steps), "--warmup", str(warmup)], env=env, capture_output=True, text=True, timeout=1500)
    for line in reversed((r.stdout or "").splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu worker failed: " + (r.stderr or "")[-400:])


# ---------------------------------------------------------------------------------------------------
# the reference's own CUDA path (llg-heun-gpu: curand normals, cuSPARSE SpMV field, cuda_heun_llg_kernelA/B) on this GPU, from
# oracle/_ref/libjams_ref_cuda.so (the reference's kernels compiled where they lie, oracle/ref_cuda_wrap.cu).  A second baseline
# next to cpu_baseline, on the headline lattice itself (sc 256^3: 302 M CSR non-zeros, 3.6 GB on the device; the reference's
# Builder needs ~75 s of host time to sort them, outside the timed region), falling back to sc 128^3 if that fails.  The product
# is timed on the same lattice in the same process, after it.
# ---------------------------------------------------------------------------------------------------
def refcuda_worker(sample_n, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import oracle
    from helpers import build_cpu_sim
    from jams_b200 import workloads as W
    torch.cuda.set_device(0)
    w = workload(1, dims=(sample_n, sample_n, sample_n))
    lat = w["lattice"]
    s0 = lat.initial_spins(seed=1)
    t0 = time.perf_counter()
    ref = build_cpu_sim(w, which="reference_cuda", dt_ps=1e-4, seed=1)
    build_s = time.perf_counter() - t0
    ref.set_spins(s0)
    ref_ms = ref.time_heun(steps, warmup)
    ref_rk4_ms = ref.time_heun(max(steps // 2, 2), 2, rk4=True)
    nnz = ref.exchange_nnz(ref.terms["exchange"])
    ref.close()

    def own(module):
        from jams_b200.solver import create_hamiltonian, create_solver
        solver = create_solver(dict(module=module, t_step=W.T_STEP, t_max=1e-9, seed=1), lat)
        for h in w["hamiltonians"]:
            solver.register_hamiltonian(create_hamiltonian(h, lat))
        solver.set_temperature(TEMPERATURE)
        solver.set_spins(s0)
        solver.run(warmup)
        ctx = solver.ctx
        stream = torch.cuda.ExternalStream(ctx.stream(), device=0)
        ctx.synchronize(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.run(steps)
        e1.record(stream)
        e1.synchronize(); ctx.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctx.close()
        return ms

    print(json.dumps(dict(spins=lat.num_spins, ref_ms=ref_ms, own_ms=own("llg-heun-b200-gpu"), ref_rk4_ms=ref_rk4_ms, own_rk4_ms=own("llg-rk4-b200-gpu"),
                          nnz=nnz, build_s=build_s)), flush=True)


def reference_cuda_rate(steps, warmup, sample_n=N_CELLS):
    sys.path.insert(0, ROOT)
    import oracle
    if not oracle.have_ref_cuda():
        return {"unavailable": "oracle/_ref/libjams_ref_cuda.so was not built (needs the reference tree at build time)"}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "refcuda-worker", "--cpu-sample", str(sample_n),
                            "--steps", str(steps), "--warmup", str(warmup)], capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired:
        r = None
    if (r is None or r.returncode != 0) and sample_n > 128:
        return reference_cuda_rate(steps, warmup, 128)
    for line in reversed((r.stdout or "").splitlines()):
        if line.startswith("{"):
            res = json.loads(line)
            n = res["spins"]
            return {"value": n / (res["ref_ms"] * 1e-3), "unit": UNIT, "ms_per_step": res["ref_ms"], "kind": "reference",
                    "path": "CUDAHeunLLGSolver::run: curandGenerateNormalDouble + scaling kernel, 2 x (cuSPARSE SpMV on the 3N x 3N CSR "
                            "+ Zeeman field + daxpy), cuda_heun_llg_kernelA / B; the reference's sources compiled for sm_100a",
                    "sample": f"sc {sample_n}^3 ({n} spins, {res['nnz']} CSR non-zeros) NN exchange + Zeeman, T={TEMPERATURE} K, {steps} Heun steps "
                              f"after {warmup} warm-up, CUDA events; device-resident; matrix assembly {res['build_s']:.0f} s on the host, untimed",
                    "product_same_lattice": {"value": n / (res["own_ms"] * 1e-3), "unit": UNIT, "ms_per_step": res["own_ms"]},
                    "speedup_same_lattice": res["ref_ms"] / res["own_ms"],
                    "rk4": {"reference_ms_per_step": res["ref_rk4_ms"], "product_ms_per_step": res["own_rk4_ms"],
                            "reference_value": n / (res["ref_rk4_ms"] * 1e-3), "product_value": n / (res["own_rk4_ms"] * 1e-3), "unit": UNIT,
                            "speedup_same_lattice": res["ref_rk4_ms"] / res["own_rk4_ms"],
                            "path": "CudaRK4BaseSolver::run (llg-rk4-gpu): 4 x (SpMV + Zeeman + daxpy + cuda_llg_rk4_kernel), cublas mid-points, "
                                    "combination + normalisation kernels"}}
    return {"error": "reference CUDA worker failed: " + ((r.stderr or "")[-300:] or "no output")}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def cpu_reference_rate(steps, warmup, budget_s=100.0):
    """the reference CPU path for `steps` timed steps after `warmup`, on the largest sample lattice (sc n^3, n <= 96) that keeps
    the run inside `budget_s`; 1 thread and all host threads are probed, the faster runs the requested steps"""
    nthr = host_threads()
    probes = {}
    for thr in sorted({1, min(4, nthr), nthr}):
        probes[thr] = cpu_run(48, 3, 1, thr)                # 110 592 spins: a fraction of a second per step
    best_thr = max(probes, key=lambda k: probes[k]["rate"])
    per_spin_step = 1.0 / probes[best_thr]["rate"]
    n = 96
    for cand in (96, 80, 64, 48, 32):
        n = cand
        if (steps + warmup) * cand ** 3 * per_spin_step <= budget_s:
            break
    res = cpu_run(n, steps, warmup, best_thr)
    tried = {str(k): v["rate"] for k, v in probes.items()}
    what = ("reference headers (SparseMatrix::multiply, MultiArray, Vec3) + restated HeunLLGSolver::run, oracle/_ref"
            if res["kind"] == "reference" else "oracle restatement (oracle/jams_oracle.cpp)")
    base = dict(value=res["rate"], unit=UNIT, cores=res["threads"], kind=res["kind"], threads_tried=tried,
                sample=f"sc {n}^3 ({res['spins']} spins) NN exchange + Zeeman, T={TEMPERATURE} K, {steps} Heun steps after {warmup} warm-up; {what}; "
                       f"OMP threads={res['threads']} (probe on sc 48^3, spin-updates/s per thread count: {tried}), host cpus={nthr}")
    return base, res["s_per_step"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base, s_per_step = cpu_reference_rate(steps, warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 sc 256^3 NN Heisenberg + Zeeman, Langevin T=100 K, dt=1e-16 s (bounded sample: see cpu_baseline.sample)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from jams_b200 import workloads as W
    from jams_b200.distributed import TorchComm
    from jams_b200.solver import MagnetisationMonitor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm(periodic_x=True, device=f"cuda:{local_rank}")

    K, Wm = args.steps, max(3, args.warmup)
    options = {}
    if args.kernel is not None:
        options["kernel"] = args.kernel

    def build(dims, seed=20261017, random_init=True):
        """solver + pinned host image of globals::s for this rank's slab (random_init = False: the materials' initial spin
        direction -- stage times do not depend on the spin values, and a 134 M-spin random stream per rank is slow to make)"""
        w = workload(world, dims=dims)
        lat = w["lattice"]
        solver = W.make_solver(w, comm=comm, seed=seed, options=dict(options), device=local_rank)
        n_local = solver.nx * lat.dims[1] * lat.dims[2] * lat.M
        host = torch.empty((n_local, 3), dtype=torch.float64).pin_memory()
        host.numpy()[:] = lat.initial_spins(solver.x0, solver.nx, seed=1 if random_init else None)
        solver._spins0 = host.numpy()
        solver._build()          # materials, template, halos, first import
        return solver, lat, host, n_local

    def barrier(ctx):
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(ctx, fn):
        """fn() enqueues / runs the region; returns max-over-ranks device milliseconds"""
        stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
        barrier(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ctx.synchronize()
        ms = e0.elapsed_time(e1)
        barrier(ctx)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def device_resident(solver, steps, warm):
        ctx = solver.ctx
        solver.run(warm)
        barrier(ctx)
        # per-launch CUDA events inside the timed region, on every 4th step: an event record between two launches costs the
        # stream about 2 us (measured: 0.3840 against 0.3765 ms per step with a record around every launch, profiles r02av), so
        # bracketing every launch would take 2 % off the number being measured
        ctx.set_option("time_kernels", KERNEL_TIMING_EVERY)
        ctx.last_step_kernel_ms()
        l0 = ctx.kernel_launches()
        ms = timed(ctx, lambda: solver.run(steps))
        launches = ctx.kernel_launches() - l0
        sampled = max(ctx.timed_steps(), 1)
        stage_ms = ctx.last_step_kernel_ms() / sampled     # average launch duration of stage A and stage B over the sampled steps of the timed region
        ctx.set_option("time_kernels", 0)
        return ms, launches, stage_ms

    # ---- the headline: weak scaling, every rank owns 256^3 of a (256 N) x 256 x 256 lattice ---------------------------
    solver, lat, host, n_local = build(None)
    ctx = solver.ctx
    n_total = lat.num_spins
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, stage_ms = device_resident(solver, K, Wm)
    clocks = sampler.result()
    value = n_total * K / (ms * 1e-3)

    peak, peak_src = hbm_peak()
    # data flow of the default kernel (option recover_u): the predictor moves 24 B in + 24 B out per spin, the
    # corrector 48 B in + 24 B out: 120 B of HBM traffic per update.  SURVEY.md 8d's per-update figure stays 144 B: step_frac.
    recover = args.kernel in (None, 2)
    stage_bytes = [48.0, 72.0] if recover else [72.0, 72.0]
    dom = int(np.argmax(stage_ms))
    ach = stage_bytes[dom] * n_local / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": ["stage_pair_kernel<0> (predictor)", "stage_pair_kernel<1> (corrector)"][dom], "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                "algorithmic_bytes_per_launch": stage_bytes[dom] * n_local,
                "stage_ms": [float(stage_ms[0]), float(stage_ms[1])],
                "stage_ms_sampling": "CUDA events around the stage launches of every %d-th step of the timed region (a record between two "
                                     "launches costs the stream ~2 us: bracketing every launch would slow the timed steps by 2 %%)" % KERNEL_TIMING_EVERY,
                "stage_frac": [stage_bytes[k] * n_local / (stage_ms[k] * 1e-3) / 1e9 / peak for k in range(2)],
                "data_flow": ("recover_u: stage A moves 48 B and stage B 72 B per spin (120 B per update)" if recover
                              else "store u: 72 B per spin and stage (144 B per update)"),
                "step_frac": BYTES_PER_UPDATE * n_local / (ms / K * 1e-3) / 1e9 / peak,
                "step_model": "144 B per spin-update (SURVEY.md 8d) x spins per GPU / ms_per_step"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture (256^3 launches)
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(["stage_A", "stage_B"][dom] + ("_recover_u" if recover else ""))
        except Exception:  # noqa: BLE001
            pass

    # ---- end to end through the plugin surface with host buffers --------------------------------------
    mon = MagnetisationMonitor(dict(output_steps=OUTPUT_STEPS), lat)
    host_ptr = host.data_ptr()

    def e2e_loop(interval, steps):
        done = 0
        while done < steps:
            n = min(interval, steps - done)
            if world > 1:
                comm.barrier(ctx)                      # neighbours must be done reading the ghosts the import overwrites
            ctx.import_spins_ptr(host_ptr, 0)          # H2D from pinned globals::s
            solver.run(n)
            ctx.export_spins_ptr(host_ptr, 0)          # D2H into pinned globals::s (synchronises)
            mon.update(solver)                         # magnetisation reduce + 32 B D2H (+ all-reduce at N > 1)
            done += n

    def e2e_measure(interval, steps):
        e2e_loop(interval, min(interval, steps))       # one untimed interval: first-touch allocations, pinned-page mapping
        t0 = time.perf_counter()
        ms_dev = timed(ctx, lambda: e2e_loop(interval, steps))
        wall = time.perf_counter() - t0
        n_int = (steps + interval - 1) // interval
        return {"value": n_total * steps / (ms_dev * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_local * 24.0 * n_int / steps,
                "d2h_bytes_per_step": (n_local * 24.0 + 32.0) * n_int / steps, "steps": steps, "monitor_interval": interval,
                "ms_per_step": ms_dev / steps, "wall_ms_per_step": wall * 1e3 / steps,
                # bytes that crossed PCIe (both directions, one after the other) over the time not spent in the stage kernels
                "pcie_GBs_effective": 2.0 * n_local * 24.0 * n_int / max(1e-9, (ms_dev - (ms / K) * steps) * 1e-3) / 1e9}

    def pcie_raw():
        """plain pinned <-> device copies of 256 MiB on the copy engine: what this box's PCIe link gives, for context"""
        n = 32 * 1024 * 1024
        h = torch.empty(n, dtype=torch.float64).pin_memory()
        d = torch.empty(n, dtype=torch.float64, device=f"cuda:{local_rank}")
        out = {}
        for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); e1.synchronize()
            out[name] = n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        return out

    if world > 1:
        comm.barrier(ctx)
    interval = max(1, min(OUTPUT_STEPS, K // 4))
    e2e = e2e_measure(interval, K)
    e2e["note"] = (f"JAMS main-loop form: import globals::s, {interval} steps, export + magnetisation, repeated; the reference's default monitor "
                   f"interval is {OUTPUT_STEPS} steps, shortened to K/4 here so that the state crosses PCIe at least four times")
    e2e_every = e2e_measure(1, min(K, 10))
    try:
        e2e["pcie_raw_GBs"] = pcie_raw()
    except Exception as e:  # noqa: BLE001
        e2e["pcie_raw_GBs"] = {"error": str(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 sc {lat.dims[0]}x{lat.dims[1]}x{lat.dims[2]} NN Heisenberg + Zeeman, Langevin T={TEMPERATURE} K, dt=1e-16 s, alpha=0.1",
                       "spins": n_total, "spins_per_gpu": n_local, "partition": f"x-slabs x{world}" if world > 1 else "single slab",
                       "halo": "P2P stores + epoch flags, both inside the two stage kernels" if world > 1 else "none",
                       "l2": "working set per stage 0.8-1.2 GB >> 126 MB L2, no flush needed"},
            "clocks": clocks, "e2e": e2e, "e2e_every_step": e2e_every, "gpu_launches": int(launches), "roofline": roofline}
    if world > 1:
        comm.barrier(ctx)
    solver.ctx.close()
    del solver, host

    # ---- N > 1: parity of the slab decomposition, and strong scaling of BASELINE config 5 (sc 512^3) ------------------
    if world > 1 and not args.no_extra:
        line["parity"] = slab_parity(world, rank, local_rank, comm, build)
        s5, lat5, host5, n5 = build((512, 512, 512), random_init=False)
        ms5, _, st5 = device_resident(s5, min(K, 50), 3)
        k5 = min(K, 50)
        line["strong"] = {"workload": "C5 sc 512x512x512 (134 217 728 spins) cut into %d x-slabs" % world, "value": lat5.num_spins * k5 / (ms5 * 1e-3),
                          "unit": UNIT, "steps": k5, "ms_per_step": ms5 / k5, "stage_ms": [float(st5[0]), float(st5[1])],
                          "step_frac_per_gpu": BYTES_PER_UPDATE * n5 / (ms5 / k5 * 1e-3) / 1e9 / peak, "scaling": "strong"}
        comm.barrier(s5.ctx)
        s5.ctx.close()
        del s5, host5

    # ---- N = 1: the other BASELINE configurations as secondary entries, and the CPU arm ------------------------------------
    if rank == 0 and world == 1 and not args.no_extra:
        line["other_configs"] = other_configs(min(K, 20), peak)
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"], _ = cpu_reference_rate(args.cpu_steps, 2, budget_s=20.0)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
        try:
            line["reference_cuda"] = reference_cuda_rate(min(K, 20), 3)
        except Exception as e:  # noqa: BLE001
            line["reference_cuda"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def slab_parity(world, rank, local_rank, comm, build):
    """a small lattice stepped by `world` slabs against the same lattice in ONE slab on rank 0's GPU: the Philox noise is keyed by
    the global site, so the two runs must agree bit for bit"""
    import torch
    import torch.distributed as dist
    from jams_b200 import workloads as W
    dims = (8 * world, 12, 20)
    steps = 25
    s, lat, host, n_local = build(dims, seed=4242)
    s.run(steps)
    mine = torch.from_numpy(s.spins().copy()).to(f"cuda:{local_rank}")
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    comm.barrier(s.ctx)
    s.ctx.close()
    out = {"lattice": "sc %dx%dx%d, T=%g K, %d steps" % (dims + (TEMPERATURE, steps))}
    if rank == 0:
        got = torch.cat(parts).cpu().numpy()
        w = workload(1, dims=dims)
        single = W.make_solver(w, seed=4242, device=local_rank)
        single.set_spins(w["lattice"].initial_spins(seed=1))
        single.run(steps)
        out["max_abs_diff_vs_single_slab"] = float(np.abs(single.spins() - got).max())
        single.ctx.close()
    return out


def other_configs(steps, peak_gbs=None):
    """BASELINE configs 2 and 4 and the RK4 solver on the bench workload (parity-test cases, not the headline): stage-kernel rates on
    this GPU.  C4 also reports the gather rate SURVEY.md 8d asks for (neighbour-spin operands delivered per second)."""
    from jams_b200 import workloads as W
    from jams_b200.solver import create_hamiltonian, create_solver
    names = {0: "direct gathers", 2: "TMA pair kernel", 4: "TMA rows kernel", 5: "general neighbour list"}
    out = []
    for name, w, z in (("C2 bcc Fe 64^3 NN+NNN (z = 14), T = 300 K (BASELINE config 2)", W.c2_bcc_fe(64, temperature=300.0), 14),
                       ("C2 bcc Fe 128^3 NN+NNN (z = 14), T = 300 K", W.c2_bcc_fe(128, temperature=300.0), 14),
                       ("C2 bcc Fe 256^3 NN+NNN (z = 14), T = 300 K (the same kernel on a lattice that fills the GPU)", W.c2_bcc_fe(256, temperature=300.0), 14),
                       ("C4 bcc 128^3, 8 shells (z = 112), T = 0", W.c4_bcc_long_range(128, temperature=0.0), 112)):
        try:
            s = W.make_solver(w, options=dict(time_kernels=1), random_spins_seed=1)
            s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
            s.run(steps); s.ctx.synchronize()
            st = s.ctx.last_step_kernel_ms() / steps
            n = w["lattice"].num_spins
            out.append({"config": name, "spins": n, "stage_ms": [float(st[0]), float(st[1])], "value": n / (float(st.sum()) * 1e-3), "unit": UNIT,
                        "kernel": names.get(s.ctx.stage_kernel(), "?"), "gather_TBs": [n * z * 24.0 / (float(t) * 1e-3) / 1e12 for t in st],
                        "frac_of_144B_model": (144.0 * n / (float(st.sum()) * 1e-3) / 1e9 / peak_gbs) if peak_gbs else None,
                        "timing": "sum of the two stage launches, CUDA events"})
            s.ctx.close()
        except Exception as e:  # noqa: BLE001
            out.append({"config": name, "error": str(e)})
    try:   # llg-rk4-b200-gpu on the bench workload: four launches per step, 336 B of HBM traffic per spin-update
        w = W.c3_sc(256, temperature=TEMPERATURE)
        lat = w["lattice"]
        s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=W.T_STEP, t_max=1e-9, seed=3, options=dict(time_kernels=1)), lat)
        for h in w["hamiltonians"]:
            s.register_hamiltonian(create_hamiltonian(h, lat))
        s.set_temperature(TEMPERATURE)
        s.set_spins(lat.initial_spins(seed=1))
        s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
        s.run(steps); s.ctx.synchronize()
        st = s.ctx.last_step_kernel_ms() / steps
        ms = float(st.sum())
        rec = {"config": "RK4-LLG on C3 sc 256^3, T = %g K" % TEMPERATURE, "spins": lat.num_spins, "stage_ms": [float(st[0]), float(st[1])],
               "stage_ms_note": "stages 1 + 2, stages 3 + 4", "value": lat.num_spins / (ms * 1e-3), "unit": UNIT,
               "kernel": names.get(s.ctx.stage_kernel(), "?"), "bytes_per_update": 336,
               "bytes_note": "48 + 72 + 120 + 96 B: no stored k-sum (DESIGN.md 3.2b); 408 B with one"}
        if peak_gbs:
            rec["hbm_frac"] = 336.0 * lat.num_spins / (ms * 1e-3) / 1e9 / peak_gbs
            rec["frac_of_408B_model"] = 408.0 * lat.num_spins / (ms * 1e-3) / 1e9 / peak_gbs
        out.append(rec)
        s.ctx.close()
    except Exception as e:  # noqa: BLE001
        out.append({"config": "RK4-LLG on C3", "error": str(e)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu-worker", "refcuda-worker"])
    ap.add_argument("--kernel", type=int, default=None, help="0 = direct gathers through L1/L2, 2 = the TMA stage kernel (the library default)")
    ap.add_argument("--temperature", type=float, default=None, help="thermostat temperature of the workload in K (default 100; 0 = the deterministic T = 0 variant, a profile artefact and not the headline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary records (other BASELINE configs at N = 1; slab parity and strong scaling at N > 1)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=96, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.temperature is not None:
        global TEMPERATURE
        TEMPERATURE = float(args.temperature)
    if args.impl == "cpu-worker":
        cpu_worker(args.cpu_sample, args.steps, args.warmup)
    elif args.impl == "refcuda-worker":
        refcuda_worker(args.cpu_sample, args.steps, args.warmup)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
