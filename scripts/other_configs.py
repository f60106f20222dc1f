"""Development helper: stage-kernel times of the other BASELINE.json configurations (parity-test cases, not bench lines).
    python scripts/other_configs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jams_b200 import workloads as W


def run(name, w, steps, options=None):
    s = W.make_solver(w, options=dict(options or {}, time_kernels=1, verbose=0), random_spins_seed=None if w.get("spins") is not None else 1)
    s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
    t0 = time.perf_counter(); s.run(steps); s.ctx.synchronize(); wall = (time.perf_counter() - t0) / steps
    ms = s.ctx.last_step_kernel_ms() / steps
    N = w["lattice"].num_spins
    print(f"{name:58s} N={N:9d}  A {ms[0]:.4f} ms  B {ms[1]:.4f} ms  wall/step {wall*1e3:.4f} ms -> {N/ (ms.sum()*1e-3)/1e9:6.2f} G upd/s "
          f"({144*N/(ms.sum()*1e-3)/1e12:.2f} TB/s of the 144 B model)", flush=True)
    s.ctx.close()


if __name__ == "__main__":
    run("C1 bloch wall sc 256x16x16, T=0 (open x)", W.c1_bloch_wall((256, 16, 16)), 200)
    run("C2 bcc Fe 64^3 NN+NNN (z=14), T=300 K", W.c2_bcc_fe(64, temperature=300.0), 100)
    run("C2 bcc Fe 128^3 NN+NNN (z=14), T=300 K", W.c2_bcc_fe(128, temperature=300.0), 50)
    run("C4 bcc 128^3, 8 shells (z=112), T=0", W.c4_bcc_long_range(128), 10)
    run("C4 bcc 128^3, 8 shells (z=112), T=0, direct kernel", W.c4_bcc_long_range(128), 10, dict(kernel=0))
    run("C3 sc 256^3, T=100 K (bench workload)", W.c3_sc(256, temperature=100.0), 50)
