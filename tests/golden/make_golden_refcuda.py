"""Generates tests/golden/refcuda_*.npz ON THE GPU BOX from the reference's own CUDA kernels (oracle/_ref/libjams_ref_cuda.so =
oracle/ref_cuda_wrap.cu: the reference's sources compiled for sm_100a where they lie).  Run:

    gpurun -- 'python tests/golden/make_golden_refcuda.py gpurun_out/refcuda_golden'   # then copy the .npz files into tests/golden/

The vectors pin the three restatements the reference ships no CPU code or test for -- CudaRK4BaseSolver::run, the biquadratic
exchange field kernel, PinnedBoundariesPhysics::update's CUDA branch -- and the CUDA Heun kernels, in the CPU suite as well."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
import refcuda_cases as RC  # noqa: E402
from helpers import ENERGY_UNITS, build_cpu_sim, oracle_exchange_pairs, random_unit_spins  # noqa: E402
from jams_b200 import workloads as W  # noqa: E402
from jams_b200.solver import create_hamiltonian, create_solver  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    # RK4 and Heun at T = 0 on two lattices
    for name, w in (("sc", RC.sc_three_terms((12, 9, 20))), ("bcc", RC.bcc_fe())):
        lat = w["lattice"]
        s0 = random_unit_spins(lat.num_spins, 21)
        res = {"s0": s0}
        for solver, steps in (("rk4", RC.RK4_STEPS), ("heun", RC.HEUN_STEPS)):
            ref = build_cpu_sim(w, which="reference_cuda")
            ref.set_spins(s0)
            (ref.run_rk4 if solver == "rk4" else ref.run)(steps)
            res[solver] = ref.get_spins()
            ref.close()
        ref = build_cpu_sim(w, which="reference_cuda")
        ref.set_spins(s0)
        res["h"] = ref.get_h()
        ref.close()
        np.savez_compressed(os.path.join(out, f"refcuda_T0_{name}.npz"), **res)
    # same-noise T > 0: the product's Philox normals go through the reference's thermostat scaling kernel
    w = RC.sc_three_terms((8, 7, 10), temperature=RC.THERMAL_T)
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 9)
    s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=W.T_STEP, t_max=1e-9, seed=RC.THERMAL_SEED), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_temperature(RC.THERMAL_T)
    s.set_spins(s0)
    normals = np.stack([s.ctx.noise(s.step_size, RC.THERMAL_T, RC.THERMAL_SEED, n, normals_only=True) for n in range(RC.THERMAL_STEPS)])
    res = {"s0": s0, "normals": normals}
    for solver in ("rk4", "heun"):
        ref = build_cpu_sim(w, which="reference_cuda")
        ref.set_spins(s0)
        (ref.run_rk4 if solver == "rk4" else ref.run)(RC.THERMAL_STEPS, normals)
        res[solver] = ref.get_spins()
        ref.close()
    np.savez_compressed(os.path.join(out, "refcuda_thermal_sc.npz"), **res)
    # biquadratic field
    w = RC.biquadratic()
    lat = w["lattice"]
    bq = w["hamiltonians"][0]
    s0 = random_unit_spins(lat.num_spins, 31)
    i, j, J9, _ = oracle_exchange_pairs(lat, dict(bq, interaction_prefactor=1.0))
    keep = J9[:, 0] > 0.0 * ENERGY_UNITS["joules"]
    np.savez_compressed(os.path.join(out, "refcuda_biquadratic.npz"), s0=s0,
                        h=oracle.ref_cuda_biquadratic_field(lat.num_spins, i[keep], j[keep], J9[keep, 0], s0))
    # pinned boundaries: reduce, rotation matrix, rotate -- all the reference's code
    w, _ = RC.pinned_wall()
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 31) * 0.2 + w["spins"]
    s0 /= np.linalg.norm(s0, axis=1, keepdims=True)
    left, right = RC.pinned_regions(lat)
    a, mag_left = oracle.ref_cuda_pin_region(s0, lat.mus(), left, [0.0, 0.0, -1.0])
    a, mag_right = oracle.ref_cuda_pin_region(a, lat.mus(), right, [0.0, 0.0, 1.0])
    np.savez_compressed(os.path.join(out, "refcuda_pinned.npz"), s0=s0, rotated=a, mag_left=mag_left, mag_right=mag_right)
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_refcuda_out"))
