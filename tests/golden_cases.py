"""Definitions of the small golden cases shared by tests/golden/make_golden.py and the tests."""
import numpy as np

from jams_b200 import workloads as W
from jams_b200.lattice import Lattice, Material, bloch_domain_wall
from helpers import random_unit_spins


def tensor_case_pairs():
    """sc 4x3x5, periodic (T,F,T): NN pairs with a symmetric-exchange + DMI-like tensor whose transpose sits on
    the reversed bond (so the 3N x 3N matrix is symmetric), plus a zero component to exercise the `!= 0` skip."""
    dims, periodic = (4, 3, 5), (True, False, True)
    N = dims[0] * dims[1] * dims[2]

    def site(x, y, z):
        return (x * dims[1] + y) * dims[2] + z

    base = {0: np.array([[1.5, 0.3, 0.0], [-0.3, 1.5, 0.2], [0.0, -0.2, 1.1]]),
            1: np.array([[0.7, 0.0, -0.4], [0.0, 0.9, 0.0], [0.4, 0.0, 0.7]]),
            2: np.array([[2.0, 0.1, 0.0], [-0.1, 2.0, 0.0], [0.0, 0.0, 2.5]])}
    pi, pj, pJ = [], [], []
    for x in range(dims[0]):
        for y in range(dims[1]):
            for z in range(dims[2]):
                for axis in range(3):
                    for sign in (+1, -1):
                        d = [x, y, z]
                        d[axis] += sign
                        if not periodic[axis] and (d[axis] < 0 or d[axis] >= dims[axis]):
                            continue
                        d[axis] %= dims[axis]
                        pi.append(site(x, y, z)); pj.append(site(*d))
                        pJ.append((base[axis] if sign > 0 else base[axis].T).reshape(9))
    order = np.lexsort((pj, pi))
    return (np.array(pi, np.int32)[order], np.array(pj, np.int32)[order], np.array(pJ)[order], N)


def _sc_small():
    lat = Lattice([Material("A", 3.0, alpha=0.1)], np.eye(3), [("A", (0, 0, 0))], (6, 5, 8), periodic=(False, True, True))
    hams = [dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 1e-23)]),
            dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)])]
    return dict(name="sc 6x5x8 open-x", lattice=lat, hamiltonians=hams, temperature=0.0)


def _bloch_small():
    w = W.c1_bloch_wall((24, 4, 4))
    return w


def _bcc_small_T():
    w = W.c2_bcc_fe(5, temperature=300.0)
    w["hamiltonians"].append(dict(module="uniaxial", order="K2", anisotropies=[(1, [1.0, 1.0, 0.0], 2e-23), (2, [0.0, 0.0, 1.0], 1e-23)]))
    w["hamiltonians"].append(dict(module="zeeman", dc_local_field=[[0.0, 0.5, 1.0]], ac_local_field=[[0.2, 0.0, 0.0]],
                                  ac_local_frequency=[0.5]))
    return w


def _two_material():
    lat = Lattice([Material("A", 2.0, alpha=0.05, gyro=1.0), Material("B", 1.2, alpha=0.2, gyro=0.9)], np.eye(3),
                  [("A", (0, 0, 0)), ("B", (0.5, 0.5, 0.5))], (4, 6, 5), gilbert_prefactor=True)
    hams = [dict(module="exchange", energy_units="meV",
                 interactions=[("A", "B", [0.5, 0.5, 0.5], -8.0), ("B", "A", [0.5, 0.5, 0.5], -8.0),
                               ("A", "A", [1.0, 0.0, 0.0], 3.0), ("B", "B", [1.0, 0.0, 0.0], 1.0)]),
            dict(module="uniaxial", order="K3", energy_units="meV", anisotropies=[("A", [0.0, 0.0, 1.0], 0.05), ("B", [1.0, 0.0, 0.0], 0.02)]),
            dict(module="zeeman", dc_local_field=[[0.0, 0.0, 2.0], [0.3, 0.0, -1.0]])]
    return dict(name="two-material bcc-like 4x6x5", lattice=lat, hamiltonians=hams, temperature=0.0)


CASES = {
    "sc_open_T0": dict(workload=_sc_small, spins=lambda w: random_unit_spins(w["lattice"].num_spins, 1), dt_ps=1e-4, steps=200, noise_seed=0),
    "bloch_T0": dict(workload=_bloch_small, spins=lambda w: w["spins"], dt_ps=1e-4, steps=100, noise_seed=0),
    "bcc_T300": dict(workload=_bcc_small_T, spins=lambda w: random_unit_spins(w["lattice"].num_spins, 2), dt_ps=1e-4, steps=20, noise_seed=3),
    "two_material_T0": dict(workload=_two_material, spins=lambda w: random_unit_spins(w["lattice"].num_spins, 4), dt_ps=1e-4, steps=100, noise_seed=0),
}
