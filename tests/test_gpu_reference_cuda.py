"""The reference's OWN CUDA path, run on the GPU box beside the product and the CPU oracle.

`oracle/_ref/libjams_ref_cuda.so` is nvcc's build (sm_100a) of the reference's kernels and cuSPARSE field path from the sources where
they lie under /root/reference/src (oracle/ref_cuda_wrap.cu lists them; `make -C oracle refcuda`, prebuilt in the container that
has the reference tree, travels like the other built libraries).  It is what `llg-heun-gpu` / `llg-rk4-gpu` execute:
  CUDAHeunLLGSolver::run (solvers/cuda_llg_heun.cu:66-127) with cuda_heun_llg_kernelA/B, SparseMatrix::multiply_gpu (cuSPARSE SpMV),
  CudaRK4BaseSolver::run (solvers/cuda_rk4_base.cu:50-108) with cuda_llg_rk4_kernel / cuda_rk4_combination_kernel / normalise_spins_cuda,
  cuda_biquadratic_exchange_field_kernel, cuda_uniaxial_field_kernel, cuda_zeeman_ac_field_kernel,
  vector_field_*_reduce_cuda and rotate_spins_cuda (PinnedBoundariesPhysics::update's CUDA branch).

These tests pin the three restatements the oracle header lists as "unpinned by reference vectors" -- rk4_run, the biquadratic term
and pin_region -- to the reference's kernels themselves, and compare the product with the very code it replaces.  Tolerances: the
fp64 1e-10 trajectory bar of BASELINE.json; fields 1e-13 relative (cuSPARSE sums a row in its own order)."""
import os

import numpy as np
import pytest

import oracle
import refcuda_cases as RC
from helpers import ENERGY_UNITS, build_cpu_sim, oracle_exchange_pairs, random_unit_spins
from jams_b200 import workloads as W
from jams_b200.lattice import Lattice, Material
from jams_b200.solver import create_hamiltonian, create_physics, create_solver

TRAJ_TOL = 1e-10


def test_reference_cuda_library_exports_every_entry_point():
    """CPU check: where the library was built (a container with /root/reference) it loads and exports the wrapper's C entry points"""
    if not oracle.have_ref_cuda():
        pytest.skip("oracle/_ref/libjams_ref_cuda.so was not built here (no reference tree)")
    L = oracle.reference_cuda()
    for name in oracle.REF_CUDA_SYMBOLS:
        assert getattr(L, name, None) is not None, name


needs_lib = pytest.mark.skipif(not oracle.have_ref_cuda(), reason="oracle/_ref/libjams_ref_cuda.so was not built (no reference tree)")


def _solver(w, module, seed=0, options=None):
    lat = w["lattice"]
    s = create_solver(dict(module=module, t_step=W.T_STEP, t_max=1e-9, seed=seed, options=options or {}), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_temperature(w.get("temperature", 0.0))
    return s


def _two_term_sc(dims, temperature=0.0):
    w = W.c3_sc(dims=dims, temperature=temperature)
    w["hamiltonians"] = [h for h in w["hamiltonians"] if h["module"] != "zeeman"]
    w["hamiltonians"].append(dict(module="uniaxial", order="K2", anisotropies=[("A", [0.0, 0.6, 0.8], 2e-23)]))
    w["hamiltonians"].append(dict(module="zeeman", dc_local_field=[[0.0, 0.0, 0.5]], ac_local_field=[[2.0, 0.0, 0.0]], ac_local_frequency=[0.5]))
    return w


@pytest.mark.gpu
@needs_lib
@pytest.mark.parametrize("make_w,steps", [(lambda: _two_term_sc((12, 10, 36)), 30), (lambda: W.c2_bcc_fe(8, temperature=0.0), 30),
                                          (lambda: W.c1_bloch_wall((32, 6, 6)), 30), (lambda: W.c4_bcc_long_range(8), 6)])
def test_heun_T0_reference_cuda_kernels_cpu_reference_and_product_agree(make_w, steps):
    """llg-heun-gpu as the reference runs it (cuSPARSE field, kernels A / B) == llg-heun-cpu (the oracle) == the product"""
    w = make_w()
    lat = w["lattice"]
    s0 = w["spins"] if w.get("spins") is not None else random_unit_spins(lat.num_spins, 17)
    ref = build_cpu_sim(w, which="reference_cuda")
    cpu = build_cpu_sim(w)
    assert ref.exchange_nnz(ref.terms["exchange"]) == cpu.L.sim_exchange_nnz(cpu.h, cpu.terms["exchange"])      # the same CSR out of the reference's Builder
    ref.set_spins(s0); cpu.set_spins(s0)
    h_ref, s = ref.get_h(), _solver(w, "llg-heun-b200-gpu")
    s.set_spins(s0)
    h_gpu = s.compute_fields()
    assert np.abs(h_gpu - h_ref).max() <= 1e-13 * np.abs(h_ref).max()
    ref.run(steps); cpu.run(steps); s.run(steps)
    a, b, c = ref.get_spins(), cpu.get_spins(), s.spins()
    assert np.abs(a - b).max() <= TRAJ_TOL      # the reference's two implementations agree (so the CPU oracle stands for both)
    assert np.abs(c - a).max() <= TRAJ_TOL      # the product against the CUDA code it replaces
    assert np.abs(c - b).max() <= TRAJ_TOL


@pytest.mark.gpu
@needs_lib
def test_heun_thermal_same_noise_reference_cuda_kernels_and_product_agree():
    """T > 0 with the product's Philox normals handed to the reference's thermostat scaling kernel (sigma sqrt(T) n,
    cuda_thermostat_classical.cc:55) and to the CPU oracle"""
    T, seed, steps = 150.0, 77, 20
    w = _two_term_sc((10, 9, 20), temperature=T)
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 4)
    s = _solver(w, "llg-heun-b200-gpu", seed=seed)
    s.set_spins(s0)
    normals = np.stack([s.ctx.noise(s.step_size, T, seed, n, normals_only=True) for n in range(steps)])
    ref = build_cpu_sim(w, which="reference_cuda"); cpu = build_cpu_sim(w)
    ref.set_spins(s0); cpu.set_spins(s0)
    ref.run(steps, normals); cpu.run(steps, normals); s.run(steps)
    a, b, c = ref.get_spins(), cpu.get_spins(), s.spins()
    assert np.abs(a - b).max() <= TRAJ_TOL and np.abs(c - a).max() <= TRAJ_TOL
    # the reference's own curand stream also runs (a different stream: only the invariants hold)
    ref.set_spins(s0)
    ref.run(5)
    d = ref.get_spins()
    assert np.abs(np.linalg.norm(d, axis=1) - 1.0).max() < 1e-12 and np.abs(d - s0).max() > 1e-6


@pytest.mark.gpu
@needs_lib
def test_zero_safe_reference_kernels_and_product_agree_on_vacancies():
    """zero-length spins switch the reference to cuda_zero_safe_heun_llg_kernelA/B (solvers/cuda_llg_heun.cu:42-54).  The zero-safe
    kernel B multiplies the field by gyro a second time (cuda_llg_heun_kernel.cuh: `h[p0] = (...) * gyro_dev[idx]`), which the CPU
    solver does not; with gyro = 1 (below) both of the reference's implementations and the product agree."""
    w = W.c3_sc(dims=(10, 8, 12), temperature=0.0)
    w["hamiltonians"].append(dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 1e-22)]))
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 5)
    holes = np.random.default_rng(3).choice(lat.num_spins, lat.num_spins // 10, replace=False)
    s0[holes] = 0.0
    cpu = build_cpu_sim(w)
    cpu.set_spins(s0); cpu.run(20)
    s = _solver(w, "llg-heun-b200-gpu")
    s.set_spins(s0); s.run(20)
    assert np.abs(s.spins() - cpu.get_spins()).max() <= TRAJ_TOL and np.array_equal(s.spins()[holes], np.zeros((len(holes), 3)))
    # the reference's zero-safe CUDA kernels with unit gyro, against the CPU oracle assembled with the same arrays
    from helpers import ref_material_arrays
    mus, gyro, alpha = ref_material_arrays(lat)
    one = np.ones_like(gyro)
    ref = oracle.RefCudaSim(mus, one, alpha); cpu1 = oracle.CpuSim(mus, one, alpha)
    i, j, J9, _ = oracle_exchange_pairs(lat, w["hamiltonians"][0])
    for sim in (ref, cpu1):
        sim.add_exchange(i, j, J9)
        sim.init_solver(1e-4, lat.gilbert_prefactor, 1)
        sim.set_spins(s0)
        sim.run(20)
    a, b = ref.get_spins(), cpu1.get_spins()
    assert np.array_equal(a[holes], np.zeros((len(holes), 3)))
    assert np.abs(a - b).max() <= TRAJ_TOL


@pytest.mark.gpu
@needs_lib
@pytest.mark.parametrize("make_w,steps", [(lambda: _two_term_sc((12, 9, 20)), 30), (lambda: W.c2_bcc_fe(6, temperature=0.0), 25),
                                          (lambda: W.c1_bloch_wall((32, 6, 6)), 30)])
def test_rk4_T0_oracle_restatement_is_pinned_to_the_reference_kernels(make_w, steps):
    """the oracle's rk4_run (a restatement of CudaRK4BaseSolver::run, the reference has no CPU RK4) against the
    reference's kernels executing that very sequence (cublasDcopy / cublasDaxpy mid-points, cuda_llg_rk4_kernel, the combination
    kernel, normalise_spins_cuda), and the product's ring / direct RK4 stages against both"""
    w = make_w()
    lat = w["lattice"]
    s0 = w["spins"] if w.get("spins") is not None else random_unit_spins(lat.num_spins, 21)
    ref = build_cpu_sim(w, which="reference_cuda"); cpu = build_cpu_sim(w)
    ref.set_spins(s0); cpu.set_spins(s0)
    ref.run_rk4(steps); cpu.run_rk4(steps)
    a, b = ref.get_spins(), cpu.get_spins()
    assert np.abs(a - b).max() <= 1e-12          # restatement == reference kernels (the pin: two orders below the trajectory bar)
    for options in (None, dict(kernel=0)):
        s = _solver(w, "llg-rk4-b200-gpu", options=options)
        s.set_spins(s0); s.run(steps)
        assert np.abs(s.spins() - a).max() <= TRAJ_TOL


@pytest.mark.gpu
@needs_lib
def test_rk4_thermal_same_noise_oracle_restatement_is_pinned_to_the_reference_kernels():
    """one noise draw per step for all four stages (cuda_rk4_base.cu:65), AC field at t0, t0 + dt/2, t0 + dt"""
    T, seed, steps = 40.0, 4321, 20
    w = _two_term_sc((8, 7, 10), temperature=T)
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 9)
    s = _solver(w, "llg-rk4-b200-gpu", seed=seed)
    s.set_spins(s0)
    normals = np.stack([s.ctx.noise(s.step_size, T, seed, n, normals_only=True) for n in range(steps)])
    ref = build_cpu_sim(w, which="reference_cuda"); cpu = build_cpu_sim(w)
    ref.set_spins(s0); cpu.set_spins(s0)
    ref.run_rk4(steps, normals); cpu.run_rk4(steps, normals); s.run(steps)
    a, b, c = ref.get_spins(), cpu.get_spins(), s.spins()
    assert np.abs(a - b).max() <= 1e-12 and np.abs(c - a).max() <= TRAJ_TOL


@pytest.mark.gpu
@needs_lib
def test_biquadratic_restatement_and_product_are_pinned_to_the_reference_kernel():
    """cuda_biquadratic_exchange_field_kernel on the scalar N x N CSR out of the reference's Builder, against the
    oracle's restated term and the product's template kernel"""
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 7), periodic=(True, True, False))
    bq = dict(module="biquadratic-exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 0.8e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 0.3e-21),
                                                          ("Fe", "Fe", [1.0, 1.0, 0.0], -0.2e-21)])
    w = dict(name="bq", lattice=lat, hamiltonians=[bq], spins=None, temperature=0.0)
    s0 = random_unit_spins(lat.num_spins, 31)
    # the pair list as cuda_biquadratic_exchange.cu:127-134 keeps it (no prefactor, value > energy_cutoff)
    unit = ENERGY_UNITS["joules"]
    i, j, J9, _ = oracle_exchange_pairs(lat, dict(bq, interaction_prefactor=1.0))
    keep = J9[:, 0] > 0.0 * unit
    h_ref = oracle.ref_cuda_biquadratic_field(lat.num_spins, i[keep], j[keep], J9[keep, 0], s0)
    cpu = build_cpu_sim(w)
    cpu.set_spins(s0)
    h_cpu = cpu.term_fields(cpu.terms["biquadratic-exchange"], 0.0)
    assert np.abs(h_cpu - h_ref).max() <= 1e-15 * np.abs(h_ref).max() * 8       # same CSR order, same arithmetic: rounding of FMA contraction only
    s = _solver(w, "llg-heun-b200-gpu")
    s.set_spins(s0)
    h_gpu = s.hamiltonians[0].calculate_fields(0.0)
    assert np.abs(h_gpu - h_ref).max() <= 1e-13 * np.abs(h_ref).max()


@pytest.mark.gpu
@needs_lib
def test_pinned_boundaries_restatement_and_product_are_pinned_to_the_reference_cuda_branch():
    """physics/pinned_boundaries.cc:36-40: vector_field_indexed_scale_and_reduce_cuda -> rotation_matrix_between_vectors ->
    rotate_spins_cuda, all the reference's code, against oracle.pin_region and the product's jb_rotate_region path"""
    w = W.c1_bloch_wall((32, 6, 5))
    lat = w["lattice"]
    phys = dict(module="pinned_boundaries", left_pinned_magnetisation=[0.0, 0.0, -1.0], right_pinned_magnetisation=[0.0, 0.0, 1.0],
                left_pinned_cells=3, right_pinned_cells=2)
    s = _solver(w, "llg-heun-b200-gpu")
    p = create_physics(phys, lat)
    s.register_physics_module(p)
    s0 = random_unit_spins(lat.num_spins, 31) * 0.2 + w["spins"]
    s0 /= np.linalg.norm(s0, axis=1, keepdims=True)
    left = p.region_sites(0, False, 3, 0, lat.dims[0]); right = p.region_sites(0, True, 2, 0, lat.dims[0])
    mus = lat.mus()
    a, mag_left = oracle.ref_cuda_pin_region(s0, mus, left, [0.0, 0.0, -1.0])
    a, mag_right = oracle.ref_cuda_pin_region(a, mus, right, [0.0, 0.0, 1.0])
    b = oracle.pin_region(oracle.pin_region(s0, mus, left, [0.0, 0.0, -1.0]), mus, right, [0.0, 0.0, 1.0])
    assert np.abs(a - b).max() <= 1e-13          # restatement == the reference's CUDA branch (block-tree sum vs host loop)
    want = (mus[left, None] * s0[left]).sum(axis=0)
    assert np.abs(mag_left - want).max() <= 1e-12 * np.abs(want).max()
    s.set_spins(s0)
    s.update_physics_module()
    assert np.abs(s.spins() - a).max() <= 1e-13
    m4 = s.ctx.region_moment(0)
    # the product's region moment after the rotation against the reference's reduction of the rotated spins
    after = oracle.ref_cuda_reduce(3, a, mus, left)
    assert np.abs(m4[:3] - after).max() <= 1e-12 * np.abs(after).max()


@pytest.mark.gpu
@needs_lib
def test_magnetisation_monitor_against_the_reference_cuda_reductions():
    """row a19: vector_field_reduce_cuda / vector_field_scale_and_reduce_cuda / the indexed forms (cuda/cuda_array_reduction.cu,
    block-tree + shuffle sums) against the product's magnetisation monitor kernel, per material as MagnetisationMonitor groups"""
    from golden_cases import CASES
    from jams_b200.solver import MagnetisationMonitor
    w = CASES["two_material_T0"]["workload"]()
    lat = w["lattice"]
    x = random_unit_spins(lat.num_spins, 21)
    s = _solver(w, "llg-heun-b200-gpu")
    s.set_spins(x)
    row = MagnetisationMonitor(dict(grouping="materials"), lat).update(s)
    mus, mat = lat.mus(), lat.site_material()
    assert np.abs(oracle.ref_cuda_reduce(0, x) - x.sum(axis=0)).max() <= 1e-11
    assert np.abs(oracle.ref_cuda_reduce(1, x, mus) - (mus[:, None] * x).sum(axis=0)).max() <= 1e-11 * mus.max()
    for gidx in range(2):
        idx = np.nonzero(mat == gidx)[0]
        m = oracle.ref_cuda_reduce(3, x, mus, idx)          # monitors/magnetisation.cc:86-99 normalises by the summed moments
        want = list(m / mus[idx].sum()) + [np.linalg.norm(m) / mus[idx].sum()]
        assert np.allclose(row[2 + 4 * gidx: 6 + 4 * gidx], want, rtol=0, atol=1e-13)
        assert np.abs(oracle.ref_cuda_reduce(2, x, None, idx) - x[idx].sum(axis=0)).max() <= 1e-11


@pytest.mark.gpu
@needs_lib
def test_adapter_data_path_through_the_reference_multiarray():
    """row a18: the JAMS adapter hands globals::s.device_data() to jb_import_spins / jb_export_spins with on_device = 1
    (integration/jams/solvers/b200_llg_heun.cc).  Here globals::s is the reference's real jams::MultiArray (SyncedMemory in CUDA
    mode): the const device pointer leaves the host copy valid, the non-const one marks it stale, and the next host read -- what a
    monitor does -- downloads the exported spins."""
    w = _two_term_sc((10, 8, 12))
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 3)
    s = _solver(w, "llg-heun-b200-gpu")
    s.set_spins(random_unit_spins(lat.num_spins, 99))        # something else, so that the import matters
    s.run(1)
    steps = 7

    def between(_handle):
        s.ctx.step(steps, s.step_size, 0.0, 0.0, 0, 0)

    seen, before = oracle.ref_cuda_multiarray_contract(s.ctx.lib, s.ctx.h, s0, between)
    assert np.array_equal(before, s0)                          # import through the const pointer: host copy untouched
    want = s.ctx.export_spins()                                # the product's own host export of the same state
    assert np.array_equal(seen, want)
    cpu = build_cpu_sim(w)
    cpu.set_spins(s0)
    cpu.run(steps)
    assert np.abs(seen - cpu.get_spins()).max() <= TRAJ_TOL


@pytest.mark.gpu
@needs_lib
@pytest.mark.parametrize("rk4", [False, True])
def test_three_uniaxial_modules_reference_cuda_kernels_and_product_agree(rk4):
    """K1 + K2 + K3 as three "uniaxial" modules: the reference launches cuda_uniaxial_field_kernel once per module and sums the
    fields with daxpy (cuda/cuda_solver.cc:11-26); the product keeps them in three slots (jb_set_uniaxial_term)"""
    from helpers import ref_material_arrays, ref_uniaxial_arrays
    lat = Lattice([Material("A", 2.0, alpha=0.05), Material("B", 1.2, alpha=0.2)], np.eye(3), [("A", (0, 0, 0)), ("B", (0.5, 0.5, 0.5))], (6, 5, 8))
    hams = [dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 4e-23), ("B", [1.0, 0.0, 0.0], 2e-23)]),
            dict(module="exchange", interactions=[("A", "B", [0.5, 0.5, 0.5], 3.0e-21), ("B", "A", [0.5, 0.5, 0.5], 3.0e-21)]),
            dict(module="uniaxial", order="K2", anisotropies=[("B", [0.0, 0.6, 0.8], 3e-23)]),
            dict(module="uniaxial", order="K3", anisotropies=[(1, [1.0, 1.0, 1.0], 1e-23), (2, [0.0, 1.0, 0.0], -2e-23)])]
    w = dict(name="three uniaxial", lattice=lat, hamiltonians=hams, spins=None, temperature=0.0)
    s0 = random_unit_spins(lat.num_spins, 77)
    ref = oracle.RefCudaSim(*ref_material_arrays(lat))
    for hs in hams:
        if hs["module"] == "exchange":
            i, j, J9, _ = oracle_exchange_pairs(lat, hs)
            ref.add_exchange(i, j, J9)
        else:
            ref.add_uniaxial(*ref_uniaxial_arrays(lat, hs))
    ref.init_solver(1e-4, lat.gilbert_prefactor, 1)
    ref.set_spins(s0)
    s = _solver(w, "llg-rk4-b200-gpu" if rk4 else "llg-heun-b200-gpu")
    s.set_spins(s0)
    h_ref = ref.get_h()
    assert np.abs(s.compute_fields() - h_ref).max() <= 1e-13 * np.abs(h_ref).max()
    (ref.run_rk4 if rk4 else ref.run)(25)
    s.run(25)
    assert np.abs(s.spins() - ref.get_spins()).max() <= TRAJ_TOL


# ---- the product against the committed vectors the reference's CUDA kernels produced (tests/golden/refcuda_*.npz, generated by
# ---- tests/golden/make_golden_refcuda.py on the GPU box): these need no reference library at run time -----------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sc", "bcc"])
@pytest.mark.parametrize("solver", ["heun", "rk4", "rk4_direct"])
def test_product_reproduces_the_reference_cuda_golden_trajectories(name, solver):
    g = np.load(os.path.join(GOLD, f"refcuda_T0_{name}.npz"))
    w = RC.sc_three_terms((12, 9, 20)) if name == "sc" else RC.bcc_fe()
    s = _solver(w, "llg-heun-b200-gpu" if solver == "heun" else "llg-rk4-b200-gpu", options=dict(kernel=0) if solver == "rk4_direct" else None)
    s.set_spins(g["s0"])
    h = s.compute_fields()
    assert np.abs(h - g["h"]).max() <= 1e-13 * np.abs(g["h"]).max()
    s.run(RC.HEUN_STEPS if solver == "heun" else RC.RK4_STEPS)
    assert np.abs(s.spins() - g["heun" if solver == "heun" else "rk4"]).max() <= TRAJ_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["heun", "rk4"])
def test_product_reproduces_the_reference_cuda_golden_thermal_trajectories(solver):
    """the fixture's normals are the product's own Philox draws for (seed, step): regenerated here bit for bit, then the trajectory
    the reference's kernels integrated with them"""
    g = np.load(os.path.join(GOLD, "refcuda_thermal_sc.npz"))
    w = RC.sc_three_terms((8, 7, 10), temperature=RC.THERMAL_T)
    s = _solver(w, "llg-heun-b200-gpu" if solver == "heun" else "llg-rk4-b200-gpu", seed=RC.THERMAL_SEED)
    s.set_spins(g["s0"])
    normals = np.stack([s.ctx.noise(s.step_size, RC.THERMAL_T, RC.THERMAL_SEED, n, normals_only=True) for n in range(RC.THERMAL_STEPS)])
    assert np.array_equal(normals, g["normals"])
    s.run(RC.THERMAL_STEPS)
    assert np.abs(s.spins() - g[solver]).max() <= TRAJ_TOL


@pytest.mark.gpu
def test_product_reproduces_the_reference_cuda_golden_biquadratic_field_and_pinned_rotation():
    g = np.load(os.path.join(GOLD, "refcuda_biquadratic.npz"))
    s = _solver(RC.biquadratic(), "llg-heun-b200-gpu")
    s.set_spins(g["s0"])
    h = s.hamiltonians[0].calculate_fields(0.0)
    assert np.abs(h - g["h"]).max() <= 1e-13 * np.abs(g["h"]).max()
    g = np.load(os.path.join(GOLD, "refcuda_pinned.npz"))
    w, phys = RC.pinned_wall()
    s = _solver(w, "llg-heun-b200-gpu")
    s.register_physics_module(create_physics(phys, w["lattice"]))
    s.set_spins(g["s0"])
    s.update_physics_module()
    assert np.abs(s.spins() - g["rotated"]).max() <= 1e-13
