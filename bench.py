#!/usr/bin/env python
"""bench.py — spin-updates/s of the fused fp64 exchange + Zeeman LLG-Heun step with Langevin white noise.

    python bench.py --gpus N --steps K --warmup W                      (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference's CPU path on the host cores)

Workload (BASELINE.json configs[2], SURVEY.md 8d "C3"): simple cubic 256^3 (16 777 216 spins), NN Heisenberg
exchange J = 3.5e-21 J + Zeeman B = (0,0,1) T, alpha = 0.1, dt = 1e-4 ps, Langevin thermostat at T = 100 K
(Philox noise drawn inside both stage kernels), seeded uniform-on-sphere initial spins.  At N > 1 every rank owns a
256 x 256 x 256 x-slab of a (256 N) x 256 x 256 lattice (weak scaling; N = 8 is the 134 M spins of configs[4]) and the
boundary planes cross NVLink as peer stores from inside the stage kernels.

One "step" = one Heun step (predictor + corrector = two kernel launches) of every spin.
  value : device-resident throughput, K steps timed with CUDA events on the launching stream, max over ranks.
          Each spin array (403 MB per component set) is larger than L2 (126 MB), so no flush is needed.
  e2e   : the same K steps driven like the JAMS main loop (core/jams++.cc:333-341) through the plugin surface with
          HOST buffers: every `output_steps` (100, helpers/defaults.h:28) steps the adapter imports globals::s from
          pinned host memory (H2D, the physics module may have rewritten it), runs the interval, exports the spins
          back into the pinned host array for the monitors (D2H) and reduces the magnetisation.
          `e2e_every_step` is the same with a monitor interval of 1 (full state H2D + D2H around every step).
The oracle (CPU restatement / reference-header build) is executed only in the `cpu_baseline` and `--impl reference` legs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spin-updates/sec (fp64 LLG-Heun+exchange)"
UNIT = "spin-updates/s"
BYTES_PER_UPDATE = 144.0          # SURVEY.md 8d / DESIGN.md: 72 B per stage launch per spin
BYTES_PER_STAGE = 72.0
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md
TEMPERATURE = 100.0
OUTPUT_STEPS = 100                # monitor interval of the e2e loop
N_CELLS = 256


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
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload(n_ranks, dims=None):
    from jams_b200 import workloads as W
    dims = dims or (N_CELLS * n_ranks, N_CELLS, N_CELLS)
    w = W.c3_sc(dims=dims, temperature=TEMPERATURE)
    return w


# ---------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref when it was built, else the restatement) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, sample_n=96):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import build_cpu_sim
    which, kind = ("reference", "reference") if oracle.have_ref() else ("restatement", "port")
    w = workload(1, dims=(sample_n, sample_n, sample_n))
    lat = w["lattice"]
    sim = build_cpu_sim(w, which=which, dt_ps=1e-4, seed=1)
    sim.set_spins(lat.initial_spins(seed=1))
    cores = int(sim.L.omp_threads())
    sim.run(warmup)
    t0 = time.perf_counter()
    sim.run(steps)
    dt = time.perf_counter() - t0
    rate = lat.num_spins * steps / dt
    what = ("reference headers (SparseMatrix::multiply, MultiArray, Vec3) + restated HeunLLGSolver::run, oracle/_ref"
            if kind == "reference" else "oracle restatement (oracle/jams_oracle.cpp)")
    return dict(value=rate, unit=UNIT, cores=cores, kind=kind,
                sample=f"sc {sample_n}^3 ({lat.num_spins} spins) NN exchange + Zeeman, T={TEMPERATURE} K, {steps} Heun steps after {warmup} warm-up; {what}; "
                       f"OMP threads={cores}, host cpus={os.cpu_count()}"), dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))
    warmup = max(1, min(args.warmup, 2))
    base, s_per_step = cpu_reference_rate(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 sc 256^3 NN Heisenberg + Zeeman, Langevin T=100 K, dt=1e-16 s (bounded sample: see cpu_baseline.sample)",
                       "requested_steps": args.steps},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from jams_b200 import capi, workloads as W
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
    # default: weak scaling, every rank owns 256^3 of a (256 N) x 256 x 256 lattice (N = 8: the 134 M spins of BASELINE config 5);
    # --strong: config 5 itself, sc 512^3 cut into N x-slabs
    w = workload(world, dims=(512, 512, 512) if args.strong else None)
    lat = w["lattice"]
    options = dict(time_kernels=0)
    if args.kernel is not None:
        options["kernel"] = args.kernel
    solver = W.make_solver(w, comm=comm, seed=20261017, options=options, device=local_rank)
    n_local = solver.nx * lat.dims[1] * lat.dims[2] * lat.M
    n_total = lat.num_spins

    # pinned host image of globals::s for this slab
    host = torch.empty((n_local, 3), dtype=torch.float64).pin_memory()
    host.numpy()[:] = lat.initial_spins(solver.x0, solver.nx, seed=1)
    host_ptr = host.data_ptr()
    solver._spins0 = host.numpy()
    solver._build()          # materials, template, halos, first import
    ctx = solver.ctx
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        """fn() enqueues / runs the region; returns max-over-ranks device milliseconds"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ctx.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput -------------------------------------------------------------
    solver.run(Wm)
    barrier()
    ctx.set_option("time_kernels", 1)
    ctx.last_step_kernel_ms()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.kernel_launches()
    ms = timed(lambda: solver.run(K))
    launches = ctx.kernel_launches() - l0
    clocks = sampler.result()
    stage_ms = ctx.last_step_kernel_ms() / K     # average launch duration of stage A and stage B over the timed region
    ctx.set_option("time_kernels", 0)
    value = n_total * K / (ms * 1e-3)

    peak, peak_src = hbm_peak()
    fused = bool(stage_ms[1] == 0.0)     # fused step kernel: one launch per Heun step (jb_step_fused.cu)
    recover = False
    if fused:
        # the roofline figure is 144 B per spin-update (the two-stage data flow: s, s*, u).  The fused kernel
        # keeps s* and u on the SM and moves 48 B per update (24 B read + 24 B written), so its fraction of the 144 B
        # model may exceed 1; both readings are reported and labelled.
        ach = BYTES_PER_UPDATE * n_local / (stage_ms[0] * 1e-3) / 1e9
        ach48 = 48.0 * n_local / (stage_ms[0] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "step_fused_kernel (predictor + corrector in one launch)", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                    "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * n_local,
                    "model": "144 B per spin-update (SURVEY.md 8d, two-stage data flow); the fused kernel's own minimum is 48 B per update",
                    "achieved_48B_model": ach48, "frac_48B_model": ach48 / peak,
                    "stage_ms": [float(stage_ms[0]), 0.0], "step_frac": ach / peak}
        dom = 0
    else:
        dom = int(np.argmax(stage_ms))
        # the pair kernel's T = 0 default stores no Heun intermediate (option recover_u): its predictor moves 48 B per spin
        recover = TEMPERATURE == 0.0 and (args.kernel is None or args.kernel == 2)
        stage_bytes = 48.0 if (recover and dom == 0) else BYTES_PER_STAGE
        ach = stage_bytes * n_local / (stage_ms[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": ["stage A (predictor)", "stage B (corrector)"][dom], "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                    "algorithmic_bytes_per_launch": stage_bytes * n_local,
                    "stage_ms": [float(stage_ms[0]), float(stage_ms[1])],
                    "data_flow": ("recover_u: stage A moves 48 B and stage B 72 B per spin (120 B per update); step_frac is quoted against SURVEY 8d's 144 B per update model"
                                  if recover else "store u: 72 B per spin and stage (144 B per update)"),
                    "step_frac": BYTES_PER_UPDATE * n_local / ((stage_ms[0] + stage_ms[1]) * 1e-3) / 1e9 / peak}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(traffic_file) and not args.strong:   # the captures are of 256^3 launches
        try:
            key = "step_fused" if fused else ["stage_A", "stage_B"][dom] + ("_recover_u" if recover else "")
            roofline["traffic"] = json.load(open(traffic_file)).get(key)
        except Exception:  # noqa: BLE001
            pass

    # ---- end to end through the plugin surface with host buffers --------------------------------------
    mon = MagnetisationMonitor(dict(output_steps=OUTPUT_STEPS), lat)

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
        t0 = time.perf_counter()
        ms_dev = timed(lambda: e2e_loop(interval, steps))
        wall = time.perf_counter() - t0
        n_int = (steps + interval - 1) // interval
        return {"value": n_total * steps / (ms_dev * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_local * 24.0 * n_int / steps,
                "d2h_bytes_per_step": (n_local * 24.0 + 32.0) * n_int / steps, "steps": steps, "monitor_interval": interval,
                "ms_per_step": ms_dev / steps, "wall_ms_per_step": wall * 1e3 / steps}

    if world > 1:
        comm.barrier(ctx)
    e2e = e2e_measure(OUTPUT_STEPS, K)
    e2e_every = e2e_measure(1, min(K, 20))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{'C5' if args.strong else 'C3'} sc {lat.dims[0]}x{lat.dims[1]}x{lat.dims[2]} NN Heisenberg + Zeeman, Langevin T={TEMPERATURE} K, dt=1e-16 s, alpha=0.1",
                       "spins": n_total, "spins_per_gpu": n_local, "partition": f"x-slabs x{world}" if world > 1 else "single slab",
                       "halo": "P2P stores from the stage kernels + epoch flags" if world > 1 else "none",
                       "l2": "working set per stage 1.2 GB >> 126 MB L2, no flush needed"},
            "clocks": clocks, "e2e": e2e, "e2e_every_step": e2e_every, "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"], _ = cpu_reference_rate(args.cpu_steps, 2)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        comm.barrier(ctx)
        solver.ctx.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel", type=int, default=None, help="0 = direct gathers, 1 = TMA plane ring (one site per thread), 2 = pair kernel, two launches per step (the library default), 3 = fused step kernel")
    ap.add_argument("--temperature", type=float, default=None, help="thermostat temperature of the workload in K (default 100; 0 = the deterministic T = 0 variant, a profile artefact and not the headline)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: BASELINE config 5 (sc 512^3, 134 M spins) cut into --gpus x-slabs instead of 256^3 per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=10)
    args = ap.parse_args()
    if args.temperature is not None:
        global TEMPERATURE
        TEMPERATURE = float(args.temperature)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
