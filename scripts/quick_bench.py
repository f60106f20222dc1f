"""Development helper: time the stage kernels for a few tilings (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jams_b200 import workloads as W

def run(dims, options, T=0.0, steps=20, label=""):
    w = W.c3_sc(dims=dims, temperature=T)
    s = W.make_solver(w, options=dict(options, time_kernels=1), random_spins_seed=1)
    s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
    t0 = time.perf_counter(); s.run(steps); s.ctx.synchronize(); wall = time.perf_counter() - t0
    ms = s.ctx.last_step_kernel_ms()
    N = w["lattice"].num_spins
    rate = N * steps / (ms.sum() * 1e-3)
    print(f"{label:40s} dims={dims} T={T}: A {ms[0]/steps:.3f} ms  B {ms[1]/steps:.3f} ms  wall/step {wall/steps*1e3:.3f} ms "
          f"-> {rate/1e9:.2f} G upd/s = {rate*144/6532.2e9*100:.1f}% of HBM roofline", flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dims = (n, n, n)
    run(dims, dict(kernel=0), label="direct")
    for TY, TZ, XC, R, thr in [(8, 64, 0, 4, 256), (8, 64, 16, 8, 256), (4, 64, 0, 4, 256), (16, 64, 0, 4, 512), (8, 128, 0, 4, 256),
                               (16, 32, 0, 4, 256), (8, 64, 32, 4, 256), (8, 64, 8, 4, 256), (4, 128, 0, 8, 256)]:
        try:
            run(dims, dict(kernel=1, tile_y=TY, tile_z=TZ, chunk_x=XC, ring=R, threads=thr), label=f"tma TY={TY} TZ={TZ} XC={XC} R={R} thr={thr}")
        except Exception as e:
            print("failed", TY, TZ, XC, R, thr, e)
    run(dims, dict(kernel=0), T=300.0, label="direct thermal")
    run(dims, dict(kernel=1), T=300.0, label="tma thermal (default tiling)")
