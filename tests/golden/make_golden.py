"""Generate the golden vectors under tests/golden/ from the REFERENCE'S OWN code.

Run in a container that has /root/reference:  `python tests/golden/make_golden.py`
It builds oracle/_ref/libjams_ref.so (the reference's header-only SparseMatrix / InteractionList /
Vec3 / Mat3 compiled in place, see oracle/ref_wrap.cpp) and records what that code produces for small
seeded inputs.  The tests then hold BOTH the self-contained restatement (oracle/jams_oracle.cpp) and the
CUDA path against these files, also where the reference tree is absent (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import build_cpu_sim, oracle_exchange_pairs, random_unit_spins  # noqa: E402
from golden_cases import CASES, tensor_case_pairs  # noqa: E402


def main():
    oracle.build(ref=True)
    assert oracle.have_ref(), "reference tree not available: cannot regenerate golden vectors"
    ref = oracle.reference()

    # 1. CSR construction + SpMV of the reference Builder on a pair list with full 3x3 tensors
    i, j, J9, N = tensor_case_pairs()
    mus = np.full(N, 0.17); gyro = np.full(N, 0.176); alpha = np.full(N, 0.1)
    sim = oracle.CpuSim(mus, gyro, alpha, "reference")
    sim.add_exchange(i, j, J9, check_symmetric=True)
    row, col, val = sim.exchange_csr(0)
    s = random_unit_spins(N, 11)
    sim.set_spins(s)
    np.savez_compressed(os.path.join(HERE, "csr_tensor.npz"), i=i, j=j, J9=J9, row=row, col=col, val=val, s=s,
                        field=sim.term_fields(0), energy=sim.term_total_energy(0))

    # 2. InteractionList storage order (pairs inserted in a shuffled generation order, 3 distinct values)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(i))
    oi, oj, ov, vals = oracle.reference_interaction_list(i[perm], j[perm], J9[perm])
    np.savez_compressed(os.path.join(HERE, "interaction_list.npz"), in_i=i[perm], in_j=j[perm], in_J9=J9[perm],
                        out_i=oi, out_j=oj, out_v=ov, values=vals)

    # 3. small Vec3 / Mat3 helpers
    rng = np.random.default_rng(7)
    a = rng.standard_normal((64, 3)); b = rng.standard_normal((64, 3))
    a[0] = 0.0; a[1] = [0, 0, 1]; b[1] = [0, 0, 1]; a[2] = [0, 0, 1]; b[2] = [0, 0, -1]; a[3] = [1, 0, 0]; b[3] = [-1, 0, 0]
    unit = np.zeros_like(a); rhs = np.zeros_like(a); rot = np.zeros((64, 9))
    for n in range(64):
        ref.lib.jref_unit_vector(np.ascontiguousarray(a[n]), unit[n])
        ref.lib.jref_llg_rhs(np.ascontiguousarray(a[n]), np.ascontiguousarray(b[n]), 0.176, 0.1, rhs[n])
        if n > 0:
            ref.rotation_matrix_between_vectors(np.ascontiguousarray(a[n]), np.ascontiguousarray(b[n]), rot[n])
    np.savez_compressed(os.path.join(HERE, "vec_ops.npz"), a=a, b=b, unit=unit, rhs=rhs, rot=rot)

    # 4. Heun trajectories, fields and energies of the named cases
    for name, case in CASES.items():
        w = case["workload"]()
        lat = w["lattice"]
        sim = build_cpu_sim(w, "reference", dt_ps=case["dt_ps"])
        s0 = case["spins"](w)
        sim.set_spins(s0)
        out = dict(s0=s0, sigma=sim.sigma())
        for tname, term in sim.terms.items():
            out["field_" + tname] = sim.term_fields(term, 0.0)
            out["energy_" + tname] = sim.term_total_energy(term, 0.0)
        normals = None
        if w.get("temperature", 0.0) > 0.0:
            normals = np.random.default_rng(case["noise_seed"]).standard_normal((case["steps"], lat.num_spins, 3))
            out["normals"] = normals
        sim.run(case["steps"], normals)
        out["s_final"] = sim.get_spins()
        out["h_final"] = sim.get_h()
        out["time_final"] = sim.time()
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        print(name, "N =", lat.num_spins, "steps =", case["steps"])


if __name__ == "__main__":
    main()
