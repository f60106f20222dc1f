"""Development helper: the reference's own CUDA path (oracle/_ref/libjams_ref_cuda.so) timed on BASELINE config 2 (bcc Fe n^3, NN + NNN).
    python scripts/ref_cuda_time.py [n]"""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle
from helpers import build_cpu_sim
from jams_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = W.c2_bcc_fe(n, temperature=300.0)
lat = w["lattice"]
t0 = time.time()
ref = build_cpu_sim(w, which="reference_cuda")
print("build s", time.time() - t0, "nnz", ref.exchange_nnz(0), flush=True)
ref.set_spins(lat.initial_spins(seed=1))
ms = ref.time_heun(10, 2)
print(f"reference CUDA Heun step, C2 bcc Fe {n}^3 ({lat.num_spins} spins): {ms:.4f} ms -> {lat.num_spins / ms / 1e6:.3f} G upd/s", flush=True)
ms4 = ref.time_heun(5, 1, rk4=True)
print(f"reference CUDA RK4 step: {ms4:.4f} ms", flush=True)
