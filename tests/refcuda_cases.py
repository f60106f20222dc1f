"""Small cases whose answers were produced by the reference's OWN CUDA kernels on a B200 (tests/golden/make_golden_refcuda.py, run on
the GPU box through oracle/_ref/libjams_ref_cuda.so) and committed as tests/golden/refcuda_*.npz.  Shared by the generator, the CPU
tests (oracle restatements against the vectors) and the GPU tests (product against the vectors)."""
import numpy as np

from jams_b200 import workloads as W
from jams_b200.lattice import Lattice, Material

RK4_STEPS, HEUN_STEPS, THERMAL_STEPS, THERMAL_T, THERMAL_SEED = 30, 30, 20, 40.0, 4321


def sc_three_terms(dims, temperature=0.0):
    """sc NN exchange + K2 uniaxial on a tilted axis + Zeeman with an AC part (evaluated at t0, t0 + dt/2, t0 + dt by RK4)"""
    w = W.c3_sc(dims=dims, temperature=temperature)
    w["hamiltonians"] = [h for h in w["hamiltonians"] if h["module"] != "zeeman"]
    w["hamiltonians"].append(dict(module="uniaxial", order="K2", anisotropies=[("A", [0.0, 0.6, 0.8], 2e-23)]))
    w["hamiltonians"].append(dict(module="zeeman", dc_local_field=[[0.0, 0.0, 0.5]], ac_local_field=[[2.0, 0.0, 0.0]], ac_local_frequency=[0.5]))
    return w


def bcc_fe():
    return W.c2_bcc_fe(6, temperature=0.0)


def biquadratic():
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 7), periodic=(True, True, False))
    bq = dict(module="biquadratic-exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 0.8e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 0.3e-21),
                                                          ("Fe", "Fe", [1.0, 1.0, 0.0], -0.2e-21)])
    return dict(name="bq", lattice=lat, hamiltonians=[bq], spins=None, temperature=0.0)


def pinned_wall():
    w = W.c1_bloch_wall((32, 6, 5))
    phys = dict(module="pinned_boundaries", left_pinned_magnetisation=[0.0, 0.0, -1.0], right_pinned_magnetisation=[0.0, 0.0, 1.0],
                left_pinned_cells=3, right_pinned_cells=2)
    return w, phys


def pinned_regions(lat, left_cells=3, right_cells=2):
    """site indices of the x-edge regions as PinnedBoundariesPhysics selects them (physics/pinned_boundaries.cc:20-27)"""
    nx, ny, nz = lat.dims
    per_plane = ny * nz * lat.M
    left = np.arange(0, left_cells * per_plane, dtype=np.int32)
    right = np.arange((nx - right_cells) * per_plane, nx * per_plane, dtype=np.int32)
    return left, right
