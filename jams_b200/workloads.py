"""The BASELINE.json configurations as synthetic inputs.  Everything is generated from
integers and fixed seeds; energies are given in joules like a JAMS config file and converted by the
Hamiltonian classes."""
from __future__ import annotations

import numpy as np

from .lattice import Lattice, Material, bloch_domain_wall
from .solver import (B200HeunLLGSolver, create_hamiltonian)

T_STEP = 1e-16  # s, recommended Heun step (docs/source/solvers/llg-heun-gpu.rst; bloch_domain_wall.cfg:65-69)


def c1_bloch_wall(size=(256, 16, 16)):
    """examples/bloch_domain_wall/bloch_domain_wall.cfg:12-58 patched to T = 0 / physics 'empty'"""
    lat = Lattice([Material("A", 3.0, alpha=0.1, spin=(0, 0, 1))], np.eye(3), [("A", (0, 0, 0))], size,
                  periodic=(False, True, True))
    hams = [dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 1e-23)]),
            dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)])]
    spins = bloch_domain_wall(lat.positions(), lat.initial_spins(), width=41.56, center=size[0] / 2.0)
    return dict(name="C1 bloch wall sc %dx%dx%d" % tuple(size), lattice=lat, hamiltonians=hams, spins=spins, temperature=0.0)


def c2_bcc_fe(n=64, temperature=300.0):
    """bcc Fe, conventional cell with a 2-site motif, NN (8) + NNN (6) exchange, Langevin thermostat"""
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (n, n, n))
    hams = [dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21)])]
    return dict(name=f"C2 bcc Fe {n}^3 NN+NNN", lattice=lat, hamiltonians=hams, spins=None, temperature=temperature)


def c3_sc(n=256, temperature=0.0, field=(0.0, 0.0, 1.0), dims=None):
    """simple cubic NN Heisenberg + Zeeman"""
    dims = dims or (n, n, n)
    lat = Lattice([Material("A", 2.0, alpha=0.1)], np.eye(3), [("A", (0, 0, 0))], dims)
    hams = [dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]),
            dict(module="zeeman", dc_local_field=[list(field)])]
    return dict(name="C3 sc %dx%dx%d NN + Zeeman" % tuple(dims), lattice=lat, hamiltonians=hams, spins=None, temperature=temperature)


# bcc shells (conventional-cell units), representative vector of each of the first 8 shells: 8,6,12,24,8,6,24,24 = 112
_BCC_SHELLS = [(0.5, 0.5, 0.5), (1, 0, 0), (1, 1, 0), (1.5, 0.5, 0.5), (1, 1, 1), (2, 0, 0), (1.5, 1.5, 0.5), (2, 1, 0)]


def c4_bcc_long_range(n=128, temperature=0.0):
    """bcc with 8 shells (112 neighbours/spin) and a synthetic decaying J(r) (no ab-initio table ships with JAMS)"""
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (n, n, n))
    inter = []
    for k, r in enumerate(_BCC_SHELLS):
        rr = float(np.sqrt(np.dot(r, r)))
        J = 3.2e-21 * np.exp(-(rr - 0.8660254037844386) / 0.6) * (1.0 if k % 3 != 2 else -0.35)
        inter.append(("Fe", "Fe", list(r), float(J)))
    hams = [dict(module="exchange", interactions=inter, radius_cutoff=3.0)]
    return dict(name=f"C4 bcc {n}^3 8 shells z=112", lattice=lat, hamiltonians=hams, spins=None, temperature=temperature)


def c5_sc(n=512, dims=None):
    return dict(c3_sc(dims=dims or (n, n, n)), name="C5 sc %s NN + Zeeman" % (str(dims or (n, n, n)),))


def make_solver(workload, comm=None, t_max=1e-9, seed=0, options=None, random_spins_seed=None, device=-1):
    """Build the solver + Hamiltonians of a workload through the plugin surface"""
    lat = workload["lattice"]
    settings = dict(module="llg-heun-b200-gpu", t_step=T_STEP, t_max=t_max, seed=seed, options=options or {}, device=device)
    solver = B200HeunLLGSolver(settings, lat, comm)
    for h in workload["hamiltonians"]:
        solver.register_hamiltonian(create_hamiltonian(h, lat))
    solver.set_temperature(workload.get("temperature", 0.0))
    if random_spins_seed is not None:
        solver.set_spins(lat.initial_spins(solver.x0, solver.nx, seed=random_spins_seed))
    elif workload.get("spins") is not None:
        per_plane = lat.dims[1] * lat.dims[2] * lat.M
        solver.set_spins(workload["spins"][solver.x0 * per_plane:(solver.x0 + solver.nx) * per_plane])
    return solver
