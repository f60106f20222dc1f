"""Test glue: build the CPU oracle for the same workload definition the GPU solver gets.

The oracle side goes from the RAW config-level settings (material table, motif, `anisotropies`, `dc_local_field`, ...)
through this file's own restatement of the reference's parameter handling and the ORACLE's own template expansion and
neighbour-list construction (oracle/jams_oracle.cpp); the product side goes through jams_b200.lattice / jams_b200.solver.
Nothing of the product's parser is used on the checker's side: a unit-conversion or material-indexing
bug in jams_b200 shows up as a parity failure.  Nothing here is imported by the product."""
from __future__ import annotations

import numpy as np

import oracle

# reference helpers/consts.h:29-36 and core/units.h:15-26, typed in again here on purpose
REF_HBAR_IU = 0.6582119569              # meV ps
REF_BOHR_MAGNETON_IU = 0.0578838181     # meV / T
REF_ELECTRON_G = 2.0023193043625
REF_GYRO_IU = REF_ELECTRON_G * REF_BOHR_MAGNETON_IU / REF_HBAR_IU   # rad / (ps T)
REF_BOLTZMANN_IU = 0.0861733326         # meV / K
REF_JOULE_TO_MEV = 6.24150907e21
REF_MRYD_TO_MEV = 13.605693123
ENERGY_UNITS = {"joules": REF_JOULE_TO_MEV, "J": REF_JOULE_TO_MEV, "milli_electron_volts": 1.0, "meV": 1.0,
                "milli_rydbergs": REF_MRYD_TO_MEV, "mRyd": REF_MRYD_TO_MEV, "rydbergs": REF_MRYD_TO_MEV * 1e3, "Ryd": REF_MRYD_TO_MEV * 1e3,
                "Kelvin": REF_BOLTZMANN_IU, "K": REF_BOLTZMANN_IU}


def ref_site_tables(lat):
    """per-site (material id, motif position) in the reference's site order ((i Ny + j) Nz + k) M + m (core/lattice.cc:622-657)"""
    cells = int(lat.dims[0]) * int(lat.dims[1]) * int(lat.dims[2])
    M = len(lat.motif_material)
    mat = np.tile(np.asarray(lat.motif_material, dtype=np.int64), cells)
    if getattr(lat, "has_impurities", False):
        # lattice.impurities: which sites were substituted is a random draw (pcg32 stream, unpinned); the checker takes the realised
        # site materials as an input, like the normals of a thermal run, and derives everything else from the raw settings
        mat = np.asarray(lat.site_material(), dtype=np.int64)
    return mat, np.tile(np.arange(M, dtype=np.int64), cells)


def ref_material_arrays(lat):
    """globals::mus / gyro / alpha from the raw material table (containers/material.h:30-34, core/lattice.cc:91-97,703-713)"""
    mat, _ = ref_site_tables(lat)
    moment = np.array([m.moment * REF_BOHR_MAGNETON_IU for m in lat.materials])
    alpha = np.array([float(m.alpha) for m in lat.materials])
    gyro = np.array([m.gyro * REF_GYRO_IU for m in lat.materials])
    if lat.gilbert_prefactor:
        gyro = gyro / (1.0 + alpha * alpha)
    return moment[mat], gyro[mat], alpha[mat]


def ref_uniaxial_arrays(lat, settings):
    """power_, magnitude_ (meV), axis_ of UniaxialAnisotropyHamiltonian (hamiltonian/uniaxial_anisotropy.cc:40-114)"""
    power = {"K1": 2, "K2": 4, "K3": 6}[settings["order"]]
    unit = ENERGY_UNITS[settings.get("energy_units", "joules")]
    mat, motif = ref_site_tables(lat)
    names = [m.name for m in lat.materials]
    K = np.zeros(mat.size)
    axis = np.zeros((mat.size, 3))
    for who, ax, energy in settings["anisotropies"]:
        a = np.asarray(ax, dtype=np.float64)
        a = a / np.sqrt(a @ a)
        sel = (motif == int(who) - 1) if isinstance(who, (int, np.integer)) else (mat == names.index(who))
        K[sel] = float(energy) * unit
        axis[sel] = a
    return power, K, axis


def ref_zeeman_arrays(lat, settings):
    """dc_local_field_, ac_local_field_ (meV), ac_local_frequency_ (rad / ps) of ZeemanHamiltonian (hamiltonian/zeeman.cc:26-71)"""
    mat, _ = ref_site_tables(lat)
    mus = ref_material_arrays(lat)[0]
    dc = np.zeros((mat.size, 3))
    if "dc_local_field" in settings:
        dc = np.asarray(settings["dc_local_field"], dtype=np.float64).reshape(-1, 3)[mat] * mus[:, None]
    if "ac_local_field" in settings:
        ac = np.asarray(settings["ac_local_field"], dtype=np.float64).reshape(-1, 3)[mat] * mus[:, None]
        om = 2.0 * np.pi * np.asarray(settings["ac_local_frequency"], dtype=np.float64)[mat]
        return dc, ac, om
    return dc, None, None


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
    i, j, v, vals = oracle.neighbour_list(lat.dims, lat.periodic, lat.M, ref_site_tables(lat)[0].astype(np.int32), tmpl, motif_type=lat.motif_material)
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
    mat = ref_site_tables(lat)[0].astype(np.int32)
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
    mus, gyro, alpha = ref_material_arrays(lat)
    # "reference_cuda": the reference's own CUDA kernels + cuSPARSE field on the GPU (oracle/ref_cuda_wrap.cu), same assembly
    sim = oracle.RefCudaSim(mus, gyro, alpha) if which == "reference_cuda" else oracle.CpuSim(mus, gyro, alpha, which)
    terms = {}
    for hs in workload["hamiltonians"]:
        module = hs["module"].lower()
        if module == "exchange":
            i, j, J9, _ = oracle_exchange_pairs(lat, hs)
            terms[module] = sim.add_exchange(i, j, J9)
        elif module == "biquadratic-exchange":
            # cuda_biquadratic_exchange.cu:127-134: value = unit * J[0][0] (no interaction_prefactor), kept only if value > energy_cutoff * unit
            unit = ENERGY_UNITS[hs.get("energy_units", "joules")]
            i, j, J9, _ = oracle_exchange_pairs(lat, dict(hs, interaction_prefactor=1.0))
            keep = J9[:, 0] > hs.get("energy_cutoff", 0.0) * unit
            terms[module] = sim.add_biquadratic(i[keep], j[keep], J9[keep, 0], hs.get("check_sparse_matrix_symmetry", True))
        elif module == "exchange-functional":
            i, j, J9 = workload["functional_pairs"]   # brute-force list built by the test (brute_force_functional_pairs)
            terms[module] = sim.add_exchange(i, j, J9)
        elif module == "uniaxial":
            power, K, axis = ref_uniaxial_arrays(lat, hs)
            terms[module] = sim.add_uniaxial(power, K, axis)
        elif module == "zeeman":
            dc, ac, om = ref_zeeman_arrays(lat, hs)
            terms[module] = sim.add_zeeman(dc, ac, om)
        elif module == "applied-field":
            B = np.asarray(hs["field"], float)
            kind = str(hs.get("type", "static")).lower()
            if kind == "static":
                terms[module] = sim.add_zeeman(mus[:, None] * B[None, :])
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
