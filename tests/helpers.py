"""Test glue: build the CPU oracle for the same workload definition the GPU solver gets.

The oracle side goes from the raw config-level settings through the ORACLE's own template expansion and
neighbour-list construction (oracle/jams_oracle.cpp); the product side goes through jams_b200.lattice.
Nothing here is imported by the product."""
from __future__ import annotations

import numpy as np

import oracle
from jams_b200.consts import ENERGY_UNITS
from jams_b200.solver import create_hamiltonian


def oracle_exchange_pairs(lat, settings):
    """raw `interactions` setting -> (i, j, J9 per pair in meV) via the oracle (exchange.cc:125-169)"""
    unit = ENERGY_UNITS[settings.get("energy_units", "joules")]
    kkr = isinstance(settings["interactions"][0][0], (int, np.integer))
    inter = []
    for ti, tj, r, J in settings["interactions"]:
        if kkr:
            inter.append((ti - 1, tj - 1, r, J))
        else:
            inter.append((lat.material_index[ti], lat.material_index[tj], r, J))
    tmpl = oracle.expand_template(lat.cell, lat.motif_frac, lat.motif_material, inter, fmt="kkr" if kkr else "jams",
                                  frac_coords=settings.get("coordinate_format", "cartesian").lower() == "fractional",
                                  use_symops=settings.get("symops", True), symops=(lat.symops[0].reshape(-1, 9), lat.symops[1]),
                                  energy_cutoff=settings.get("energy_cutoff", 0.0), radius_cutoff=settings.get("radius_cutoff", 100.0),
                                  distance_tolerance=settings.get("distance_tolerance", 1e-4))
    i, j, v, vals = oracle.neighbour_list(lat.dims, lat.periodic, lat.M, lat.site_material(), tmpl, motif_type=lat.motif_material)
    prefactor = settings.get("interaction_prefactor", 1.0)
    J9 = (prefactor * unit * vals)[v]                      # Jij = prefactor * unit * J   (exchange.cc:165)
    keep = np.max(np.abs(J9), axis=1) > settings.get("energy_cutoff", 0.0) * unit  # (exchange.cc:166)
    return i[keep], j[keep], J9[keep], tmpl


def brute_force_functional_pairs(lat, functionals, tol=1e-4):
    """what ExchangeFunctionalHamiltonian's near-tree walk produces (hamiltonian/exchange_functional.cc:206-243), by brute force:
    every ordered pair of sites (i, j != i) whose minimum-image distance is within the cutoff of its material pair.
    ``functionals``: {(name_i, name_j): (r_cutoff, J(r_ij) in meV)}; lengths in lattice parameters.  Returns i, j, J9."""
    pos = lat.positions()
    names = [m.name for m in lat.materials]
    mat = lat.site_material()
    A = [lat.cell[:, k] * lat.dims[k] for k in range(3)]
    shifts = [sx * A[0] * lat.periodic[0] + sy * A[1] * lat.periodic[1] + sz * A[2] * lat.periodic[2]
              for sx in (-1, 0, 1) for sy in (-1, 0, 1) for sz in (-1, 0, 1)]
    shifts = np.unique(np.array(shifts), axis=0)
    I, J, V = [], [], []
    for i in range(lat.num_spins):
        for j in range(lat.num_spins):
            if i == j:
                continue
            key = (names[mat[i]], names[mat[j]])
            if key not in functionals:
                continue
            rc, fn = functionals[key]
            for sh in shifts:
                rij = pos[j] + sh - pos[i]
                r = np.linalg.norm(rij)
                if (r - rc) < max(abs(r), abs(rc)) * tol:
                    I.append(i); J.append(j); V.append(float(fn(rij)) * np.eye(3).reshape(9))
    return np.array(I, np.int32), np.array(J, np.int32), np.array(V).reshape(-1, 9)


def build_cpu_sim(workload, which="restatement", dt_ps=1e-4, seed=1):
    lat = workload["lattice"]
    sim = oracle.CpuSim(lat.mus(), lat.gyro(), lat.alpha(), which)
    terms = {}
    for hs in workload["hamiltonians"]:
        module = hs["module"].lower()
        if module == "exchange":
            i, j, J9, _ = oracle_exchange_pairs(lat, hs)
            terms[module] = sim.add_exchange(i, j, J9)
        elif module == "exchange-functional":
            i, j, J9 = workload["functional_pairs"]   # brute-force list built by the test (brute_force_functional_pairs)
            terms[module] = sim.add_exchange(i, j, J9)
        elif module == "uniaxial":
            h = create_hamiltonian(hs, lat)   # parameter parsing only (no numerics)
            K, axis = h.site_arrays()
            terms[module] = sim.add_uniaxial(h.power, K, axis)
        elif module == "zeeman":
            h = create_hamiltonian(hs, lat)
            dc, ac, om = h.site_arrays()
            terms[module] = sim.add_zeeman(dc, ac, om)
        elif module == "applied-field":
            B = np.asarray(hs["field"], float)
            kind = str(hs.get("type", "static")).lower()
            if kind == "static":
                terms[module] = sim.add_zeeman(lat.mus()[:, None] * B[None, :])
            else:   # applied_field.cc:37-38,66-68: seconds -> ps, Hz -> THz
                terms[module] = sim.add_applied_field(B, kind, float(hs["time_center"]) / 1e-12, float(hs["freq_bandwidth"]) / 1e12,
                                                      float(hs.get("freq_center", 0.0)) / 1e12)
        else:
            raise RuntimeError(module)
    sim.init_solver(dt_ps, lat.gilbert_prefactor, seed)
    sim.set_temperature(workload.get("temperature", 0.0))
    sim.terms = terms
    return sim


def random_unit_spins(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)
