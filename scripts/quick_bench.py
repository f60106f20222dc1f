"""Development helper: time the stage kernels for a few tilings (not the contract bench; see bench.py).
    python scripts/quick_bench.py [n] [T]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np
from jams_b200 import workloads as W

PEAK = 6538.9e9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass


def run(dims, options, T=0.0, steps=20, label="", make=W.c3_sc):
    w = make(dims=dims, temperature=T) if make is W.c3_sc else make(dims, T)
    try:
        s = W.make_solver(w, options=dict(options, time_kernels=1), random_spins_seed=1)
        s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
        t0 = time.perf_counter(); s.run(steps); s.ctx.synchronize(); wall = time.perf_counter() - t0
        ms = s.ctx.last_step_kernel_ms()
    except Exception as e:  # noqa: BLE001
        print(f"{label:44s} FAILED: {e}", flush=True)
        return
    N = w["lattice"].num_spins
    rate = N * steps / (ms.sum() * 1e-3)
    print(f"{label:44s} T={T:5.0f}: A {ms[0]/steps:.3f} ms  B {ms[1]/steps:.3f} ms  wall/step {wall/steps*1e3:.3f} ms "
          f"-> {rate/1e9:6.2f} G upd/s = {rate*144/PEAK*100:5.1f}% of HBM roofline", flush=True)
    s.ctx.close()


# (TY, TZ, SPT, R, RU, u_tma)
CONFIGS = [(7, 64, 1, 4, 2, 1), (7, 64, 1, 5, 2, 1), (7, 64, 1, 5, 2, 0), (7, 64, 1, 6, 2, 0), (6, 64, 1, 5, 2, 1), (6, 64, 1, 6, 2, 1), (3, 128, 1, 4, 2, 1),
           (3, 128, 1, 5, 2, 1), (15, 32, 1, 4, 2, 1), (15, 32, 1, 5, 2, 1), (4, 64, 1, 6, 3, 1), (4, 64, 1, 8, 3, 1)]
EXTRA = [dict()]
# pair kernel: (TY, TZ, SPT, R, RU, ctas_per_sm); 0 = heuristic
PAIR_EXTRA = []   # (label, options) appended by the caller through JB_QB_EXTRA (json list)
if os.environ.get("JB_QB_EXTRA"):
    PAIR_EXTRA = [(str(o), o) for o in json.loads(os.environ["JB_QB_EXTRA"])]
PAIR_CONFIGS = [] if os.environ.get("JB_QB_EXTRA") else [(8, 64, 1, 4, 2, 0), (8, 64, 1, 5, 2, 0), (8, 64, 1, 0, 3, 0), (16, 64, 1, 0, 2, 1), (16, 64, 1, 6, 3, 1), (4, 128, 1, 0, 2, 0),
                (8, 128, 1, 0, 2, 1), (16, 64, 2, 0, 2, 1), (8, 64, 2, 0, 2, 0), (16, 32, 1, 0, 2, 0)]

if __name__ == "__main__":
    # every configuration runs in its own process: a CUDA error is sticky for the process that hit it
    import subprocess
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        n, T = int(sys.argv[2]), float(sys.argv[3])
        opts = json.loads(sys.argv[4])
        dims = tuple(int(v) for v in os.environ["JB_QB_DIMS"].split("x")) if os.environ.get("JB_QB_DIMS") else (n, n, n)
        run(dims, opts, T=T, label=sys.argv[5])
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    temps = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.0, 100.0]
    for T in temps:
        jobs = [(dict(kernel=2, verbose=1), "pair default")] if os.environ.get("JB_QB_EXTRA") else [(dict(kernel=1, verbose=1), "tile (one site per thread) default"), (dict(kernel=2, verbose=1), "pair default")]
        for TY, TZ, SPT, R, RU, cps in PAIR_CONFIGS:
            jobs.append((dict(kernel=2, tile_y=TY, tile_z=TZ, spt=SPT, ring=R, ring_u=RU, ctas_per_sm=cps, verbose=1),
                         f"pair TY={TY} TZ={TZ} SPT={SPT} R={R} RU={RU} ctas={cps}"))
        for label, o in PAIR_EXTRA:
            jobs.append((dict(dict(kernel=2, verbose=1), **o), "pair " + label))
        for TY, TZ, SPT, R, RU, ut in (CONFIGS if os.environ.get("JB_QB_TILE") else []):
            for ex in EXTRA:
                jobs.append((dict(kernel=1, tile_y=TY, tile_z=TZ, spt=SPT, ring=R, ring_u=RU, **dict(dict(u_tma=ut), **ex)),
                             f"tile TY={TY} TZ={TZ} SPT={SPT} R={R} RU={RU} {ex}"))
        for opts, label in jobs:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(n), str(T), json.dumps(opts), label],
                               capture_output=True, text=True, timeout=300)
            out = (r.stdout or "").strip()
            for line in (r.stderr or "").splitlines():
                if line.startswith("jams_b200:"):
                    print("    " + line, flush=True)
            print(out if out else f"{label:44s} CRASHED rc={r.returncode}: {(r.stderr or '').strip()[-300:]}", flush=True)
