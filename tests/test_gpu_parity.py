"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the C ABI
(jams_b200.capi -> libjams_b200.so), against the CPU oracle and the committed golden vectors.

Tolerances: neighbour structure / import-export are bit-exact; fp64 fields agree to 1e-13 relative
(different summation instruction mix: FMA on the GPU, separate mul+add in the reference build);
T = 0 (and fixed-noise T > 0) trajectories agree to 1e-10 per spin component after N steps, the bar
BASELINE.json states."""
import gc
import os

import numpy as np
import pytest

import oracle
from helpers import build_cpu_sim, random_unit_spins
from golden_cases import CASES
from jams_b200 import capi, workloads as W
from jams_b200.lattice import Lattice, Material
from jams_b200.solver import create_hamiltonian

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAJ_TOL = 1e-10
KERNELS = {"direct": dict(kernel=0), "pair": dict(kernel=2),
           # recover_u = 1 (default): the TMA kernel rebuilds the Heun intermediate from s_n and s* (120 B per update); 0: stores it (144 B)
           "pair_store_u": dict(kernel=2, recover_u=0), "pair_recover_u": dict(kernel=2, recover_u=1),
           # other tilings / work-item plans of the same kernel: small tiles (many columns), equal chunks, short chunks only
           "pair_small_tile": dict(kernel=2, tile_y=2, tile_z=16, chunks=3), "pair_small_tile_store_u": dict(kernel=2, tile_y=2, tile_z=16, chunks=3, recover_u=0),
           "pair_short_chunks": dict(kernel=2, tile_y=3, tile_z=32, chunk_long=4, chunk_short=2, tail_pct=50),
           # the rows kernel (four y rows per thread, deep isotropic templates; forced here on every isotropic template; tensor
           # couplings fall through to the direct kernel), default tile and the smallest tile with equal chunks
           "rows": dict(kernel=4), "rows_small_tile": dict(kernel=4, tile_y=4, chunks=3)}
ALL_TMA = ["pair", "pair_store_u", "pair_recover_u", "pair_small_tile", "pair_small_tile_store_u", "pair_short_chunks", "rows", "rows_small_tile"]


def gold(name):
    return np.load(os.path.join(GOLD, name))


def make(workload, options=None, pairs=False, **kw):
    """pairs=True: hand the neighbour list over and force the general ELL kernel; pairs="auto": hand the neighbour list
    over and let the library recognise the template (what the JAMS adapter does)"""
    w = dict(workload)
    if pairs:
        w["hamiltonians"] = [dict(h, use_neighbour_list=True) if h["module"] in ("exchange", "exchange-functional") else h for h in w["hamiltonians"]]
        if pairs != "auto":
            options = dict(options or {}, detect_template=0)
    return W.make_solver(w, options=options, **kw)


# ------------------------------------------------------------------------------------------------
def test_extension_is_loaded_and_launches_kernels():
    s = make(W.c3_sc(dims=(8, 8, 8)))
    s.run(2)
    s.ctx.synchronize()
    assert s.ctx.kernel_launches() >= 3   # import + 2 fused steps (or 2 stages x 2 steps)
    assert any("libjams_b200.so" in line for line in open("/proc/self/maps"))


@pytest.mark.parametrize("dims,M,periodic", [((5, 4, 7), 1, (True, True, True)), ((3, 6, 9), 2, (False, True, False)),
                                              ((8, 3, 33), 1, (True, False, True))])
def test_import_export_round_trip_is_exact(dims, M, periodic):
    motif = [("A", (0, 0, 0)), ("A", (0.5, 0.5, 0.5))][:M]
    lat = Lattice([Material("A", 1.0)], np.eye(3), motif, dims, periodic=periodic)
    w = dict(lattice=lat, hamiltonians=[dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 1e-21)])], temperature=0.0)
    s = make(w)
    x = random_unit_spins(lat.num_spins, 5)
    s.set_spins(x)
    assert np.array_equal(s.spins(), x)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("pairs", [False, True])
def test_fields_and_energies_match_reference_golden(name, pairs):
    case, g = CASES[name], gold(f"case_{name}.npz")
    w = case["workload"]()
    s = make(w, pairs=pairs)
    s.set_spins(g["s0"])
    total = np.zeros_like(g["s0"])
    for h in s.hamiltonians:
        key = h.settings["module"].lower()
        ref_f = g["field_" + key]
        f = h.calculate_fields(0.0)
        scale = max(np.abs(ref_f).max(), 1e-300)
        assert np.abs(f - ref_f).max() <= 1e-13 * scale, key
        total += ref_f
        e = h.calculate_total_energy(0.0)
        assert abs(e - float(g["energy_" + key])) <= 1e-12 * max(abs(float(g["energy_" + key])), 1.0), key
    assert np.abs(s.compute_fields() - total).max() <= 1e-13 * np.abs(total).max()


@pytest.mark.parametrize("name", [n for n in CASES if "T0" in n])
@pytest.mark.parametrize("variant", ["direct"] + ALL_TMA + ["pairs", "pairs_auto"])
def test_T0_trajectories_match_reference_golden(name, variant):
    case, g = CASES[name], gold(f"case_{name}.npz")
    w = case["workload"]()
    s = make(w, options=KERNELS.get(variant), pairs={"pairs": True, "pairs_auto": "auto"}.get(variant, False))
    assert abs(s.step_size - case["dt_ps"]) < 1e-18
    s.set_spins(g["s0"])
    s.run(case["steps"])
    out = s.spins()
    assert np.abs(out - g["s_final"]).max() <= TRAJ_TOL
    assert np.abs(np.linalg.norm(out, axis=1) - 1.0).max() < 1e-14
    assert abs(s.time - float(g["time_final"])) < 1e-15


@pytest.mark.parametrize("variant", ["direct"] + ALL_TMA + ["pairs"])
def test_thermal_trajectory_matches_oracle_given_the_same_noise(variant):
    """T > 0: the reference's CPU (pcg) and GPU (XORWOW) noise streams already differ, so parity is defined on the
    integrator given identical noise: dump the Philox normals the kernels use and feed them to the oracle."""
    case = CASES["bcc_T300"]
    w = case["workload"]()
    lat = w["lattice"]
    steps, seed = 25, 1234
    s = make(w, options=KERNELS.get(variant), pairs=(variant == "pairs"), seed=seed)
    s0 = random_unit_spins(lat.num_spins, 8)
    s.set_spins(s0)
    normals = np.stack([s.ctx.noise(s.step_size, 300.0, seed, n, normals_only=True) for n in range(steps)])
    sim = build_cpu_sim(w, dt_ps=case["dt_ps"])
    sim.set_spins(s0)
    # Thermostat::device_data: xi = sigma sqrt(T) n
    xi = s.thermostat.noise(step=3)
    assert np.abs(xi - normals[3] * (sim.sigma() * np.sqrt(300.0))[:, None]).max() <= 1e-12 * np.abs(xi).max()
    sim.run(steps, normals)
    s.run(steps)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL


def test_golden_thermal_case_with_reference_normals_is_reproduced_by_oracle_and_statistics_of_gpu_noise():
    """the Philox/Box-Muller draws are N(0,1): moments, tails, and independence across sites, components, steps"""
    lat = Lattice([Material("A", 1.0)], np.eye(3), [("A", (0, 0, 0))], (32, 32, 32))
    w = dict(lattice=lat, hamiltonians=[dict(module="zeeman", dc_local_field=[[0, 0, 1.0]])], temperature=10.0)
    s = make(w, seed=7)
    s.spins()
    a = s.ctx.noise(1e-4, 10.0, 7, 0, normals_only=True)
    b = s.ctx.noise(1e-4, 10.0, 7, 1, normals_only=True)
    c = s.ctx.noise(1e-4, 10.0, 8, 0, normals_only=True)
    n = a.size
    for x in (a, b, c):
        assert abs(x.mean()) < 5 / np.sqrt(n)
        assert abs(x.var() - 1.0) < 5 * np.sqrt(2.0 / n)
        assert abs((x ** 4).mean() - 3.0) < 5 * np.sqrt(96.0 / n)
        assert abs((x ** 3).mean()) < 5 * np.sqrt(15.0 / n)
        assert 3.5 < np.abs(x).max() < 6.7
    for x, y in ((a[:, 0], a[:, 1]), (a[:, 0], a[:, 2]), (a[:, 1], a[:, 2]), (a.ravel(), b.ravel()), (a.ravel(), c.ravel()),
                 (a[:-1, 0], a[1:, 0])):
        assert abs(np.mean(x * y)) < 5 / np.sqrt(x.size)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a, s.ctx.noise(1e-4, 10.0, 7, 0, normals_only=True))   # counter-based: reproducible


def test_magnetisation_and_energy_monitors():
    case = CASES["two_material_T0"]
    w = case["workload"]()
    lat = w["lattice"]
    s = make(w)
    x = random_unit_spins(lat.num_spins, 21)
    s.set_spins(x)
    from jams_b200.solver import EnergyMonitor, MagnetisationMonitor
    mon = MagnetisationMonitor(dict(grouping="materials"), lat)
    row = mon.update(s)
    ref = oracle.magnetisation(x, lat.mus(), lat.site_material(), 2)
    for gidx in range(2):
        want = list(ref[gidx, :3] / ref[gidx, 3]) + [np.linalg.norm(ref[gidx, :3]) / ref[gidx, 3]]
        assert np.allclose(row[2 + 4 * gidx: 6 + 4 * gidx], want, rtol=0, atol=1e-13)
    sim = build_cpu_sim(w)
    sim.set_spins(x)
    erow = EnergyMonitor().update(s)
    for k, h in enumerate(s.hamiltonians):
        want = sim.term_total_energy(sim.terms[h.settings["module"].lower()], 0.0)
        assert abs(erow[1 + k] - want) <= 1e-12 * max(abs(want), 1.0)
        e = h.calculate_energies(0.0)
        assert np.abs(e - sim.term_energies(sim.terms[h.settings["module"].lower()])).max() <= 1e-12 * max(np.abs(e).max(), 1.0)


# ---- sizes beyond what the fixtures hold: oracle on the same seeded input, and size-independent properties ----
@pytest.mark.parametrize("make_w,steps", [(lambda: W.c3_sc(dims=(20, 18, 70)), 30), (lambda: W.c2_bcc_fe(12, temperature=0.0), 30),
                                          (lambda: W.c4_bcc_long_range(8), 10), (lambda: W.c1_bloch_wall((64, 16, 16)), 50)])
def test_midsize_trajectories_match_oracle(make_w, steps):
    w = make_w()
    lat = w["lattice"]
    s0 = w["spins"] if w.get("spins") is not None else random_unit_spins(lat.num_spins, 3)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    sim.run(steps)
    want = sim.get_spins()
    for variant in ["direct"] + ALL_TMA:
        s = make(w, options=KERNELS[variant])
        s.set_spins(s0)
        s.run(steps)
        assert np.abs(s.spins() - want).max() <= TRAJ_TOL, variant


@pytest.mark.parametrize("dims", [(7, 9, 37), (33, 5, 130), (4, 70, 3), (5, 3, 64), (6, 19, 66), (9, 4, 4), (40, 10, 6)])
def test_partial_tiles_and_odd_sizes(dims):
    w = W.c3_sc(dims=dims)
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 13)
    res = {}
    for variant in ["direct"] + ALL_TMA:
        s = make(w, options=KERNELS[variant])
        s.set_spins(s0)
        s.run(5)
        res[variant] = s.spins()
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    sim.run(5)
    assert np.abs(res["direct"] - sim.get_spins()).max() <= TRAJ_TOL
    for variant in ALL_TMA:
        assert np.abs(res[variant] - res["direct"]).max() <= 1e-14, variant


def test_full_size_properties_sc_128():
    """size-independent properties at a size the oracle cannot run in seconds: norm conservation, the
    ferromagnetic fixed point, energy dissipation at T = 0, and kernel-variant agreement"""
    w = W.c3_sc(dims=(128, 128, 128))
    lat = w["lattice"]
    s = make(w, options=KERNELS["pair"])
    s.set_spins(np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1)))
    s.run(10)
    assert np.array_equal(s.spins(), np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1)))
    s0 = lat.initial_spins(seed=2)
    s.set_spins(s0)
    e0 = sum(h.calculate_total_energy(0.0) for h in s.hamiltonians)
    s.run(40)
    out = s.spins()
    e1 = sum(h.calculate_total_energy(0.0) for h in s.hamiltonians)
    assert np.abs(np.linalg.norm(out, axis=1) - 1.0).max() < 1e-14
    assert e1 < e0
    for variant in ("direct", "pair_store_u", "pair_short_chunks"):
        d = make(w, options=KERNELS[variant])
        d.set_spins(s0)
        d.run(40)
        assert np.abs(d.spins() - out).max() <= 1e-13, variant


@pytest.mark.parametrize("n_slabs", [2, 4])
@pytest.mark.parametrize("periodic_x", [True, False])
@pytest.mark.parametrize("kernel", ["2r", "2u", "2r_fold", "2u_fold", "rows", "rows_fold", "rk4", "rk4_fold", "rk4_direct", "pairs", "direct", "direct_uni3"])
def test_slab_decomposition_in_one_process_matches_single_slab(n_slabs, periodic_x, kernel):
    """several contexts (one per x-slab) on this GPU, halos exchanged by peer stores + epoch flags; thermal noise is
    keyed by the global site so the result must equal the undecomposed run bit for bit"""
    dims = (16, 6, 10)
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], dims,
                  periodic=(periodic_x, True, True))
    hs = dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21)])
    h = create_hamiltonian(hs, lat)
    t = h.template
    nbr = h.neighbour_list() if kernel == "pairs" else None
    s0 = lat.initial_spins(seed=17)
    dt, T, seed, steps = 1e-4, 50.0, 99, 12

    def new_ctx(rank, n):
        nx = dims[0] // n
        c = capi.Context(dims, lat.M, lat.periodic, x_begin=rank * nx, nx_local=nx, rank=rank, n_ranks=n)
        if kernel.startswith("rk4"):
            c.set_option("kernel", 0 if kernel == "rk4_direct" else 2)
            c.set_option("fold_halo", 2 if kernel.endswith("fold") else 0)
        elif kernel.startswith("direct"):   # the direct-gather Heun stages (what biquadratic exchange and uniaxial slots 1, 2 run on)
            c.set_option("kernel", 0)
        else:
            c.set_option("kernel", 4 if kernel.startswith("rows") else 2)
            c.set_option("recover_u", 0 if kernel.startswith("2u") else 1)   # two launches per step, two halo exchanges
            # fold: the epoch handshake inside the stage kernel (what multi-GPU runs use); the slabs share this GPU here, which
            # works because these lattices leave most of the SMs free for the neighbour's kernel
            c.set_option("fold_halo", 2 if kernel.endswith("fold") else 0)
        c.set_materials(lat.mus(rank * nx, nx), lat.gyro(rank * nx, nx), lat.alpha(rank * nx, nx))
        if kernel == "direct_uni3":   # three uniaxial Hamiltonians: slots 1 and 2 force the direct kernels whatever the option says
            n_loc = nx * dims[1] * dims[2] * lat.M
            for slot, (power, K, axis) in enumerate([(2, 0.02, (0.0, 0.0, 1.0)), (4, 0.03, (0.0, 0.6, 0.8)), (6, -0.01, (1.0, 0.0, 0.0))]):
                c.set_uniaxial(power, np.full(n_loc, K), np.tile(np.asarray(axis), (n_loc, 1)), slot=slot)
        if kernel == "pairs":   # the general neighbour list, GLOBAL site ids on every rank: neighbours across a slab face sit in the x ghost planes
            c.set_option("detect_template", 0)
            c.set_exchange_pairs(*nbr)
        else:
            c.set_exchange_template(t["mi"], t["mj"], t["T"], t["J9"])
        return c

    step = (lambda c, *a: c.step_rk4(*a)) if kernel.startswith("rk4") else (lambda c, *a: c.step(*a))   # RK4: four exchanges per step
    single = new_ctx(0, 1)
    single.import_spins(s0)
    step(single, steps, dt, 0.0, T, seed, 0)
    want = single.export_spins()

    # The slabs share ONE GPU and one host thread here, so a slab's stage may sit in its halo wait until the host has launched the
    # neighbour's stage.  Anything that makes the host wait for the device in between deadlocks until the 10 s peer timeout --
    # in particular the cudaFree of an earlier test's context when Python's cyclic collector happens to run inside the loop.
    # (One process per GPU, the product's multi-GPU form, has no such coupling.)
    gc.collect()
    gc.disable()
    try:
        ctxs = [new_ctx(r, n_slabs) for r in range(n_slabs)]
        blobs = [c.halo_export_handle() for c in ctxs]
        for r, c in enumerate(ctxs):
            lo = r - 1 if r > 0 else (n_slabs - 1 if periodic_x else None)
            hi = r + 1 if r < n_slabs - 1 else (0 if periodic_x else None)
            c.halo_connect(blobs[lo] if lo is not None else None, blobs[hi] if hi is not None else None)
        per = lat.num_spins // n_slabs
        for r, c in enumerate(ctxs):
            c.import_spins(s0[r * per:(r + 1) * per])
        for n in range(steps):
            for c in ctxs:
                step(c, 1, dt, n * dt, T, seed, n)
        got = np.concatenate([c.export_spins() for c in ctxs])
        for c in ctxs:
            c.synchronize()
    finally:
        gc.enable()
    assert np.array_equal(got, want)


def test_kernel_choice_deep_templates_go_to_the_rows_kernel():
    for w, opts, want in ((W.c3_sc(dims=(12, 10, 40)), None, 2), (W.c3_sc(dims=(12, 10, 40)), dict(kernel=4), 4), (W.c3_sc(dims=(12, 10, 40)), dict(kernel=0), 0),
                          (W.c4_bcc_long_range(8), None, 4), (W.c4_bcc_long_range(8), dict(kernel=0), 0)):
        s = make(w, options=opts, random_spins_seed=1)
        assert s.ctx.stage_kernel() == -1
        s.run(1)
        assert s.ctx.stage_kernel() == want, (w["name"], opts)


@pytest.mark.parametrize("dims,periodic", [((8, 9, 40), (True, True, True)), ((7, 13, 33), (True, True, True)), ((9, 8, 70), (False, True, False)),
                                           ((8, 21, 8), (True, False, True)), ((16, 36, 70), (True, True, True))])   # the last one: 3 x 3 column tiles, several chunks
@pytest.mark.parametrize("T", [0.0, 200.0])
def test_deep_template_rows_kernel_matches_oracle(dims, periodic, T):
    """BASELINE config 4's template (bcc, eight shells, 112 neighbours per spin, ghost depth 3) on small lattices with ragged
    tiles, short z rows and open faces: the rows kernel against the oracle, T = 0 and T > 0 with the kernels' own noise"""
    w = W.c4_bcc_long_range(8, temperature=T)
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], dims, periodic=periodic)
    w["lattice"] = lat
    steps, seed = 8, 77
    s0 = random_unit_spins(lat.num_spins, 21)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    normals = None
    res = {}
    for variant in ("rows", "rows_small_tile", "direct"):
        s = make(w, options=KERNELS[variant], seed=seed)
        s.set_spins(s0)
        if T > 0 and normals is None:
            normals = np.stack([s.ctx.noise(s.step_size, T, seed, n, normals_only=True) for n in range(steps)])
        s.run(steps)
        assert s.ctx.stage_kernel() == (0 if variant == "direct" else 4)
        res[variant] = s.spins()
    sim.run(steps, normals)
    for variant, got in res.items():
        assert np.abs(got - sim.get_spins()).max() <= TRAJ_TOL, variant


def test_error_behaviour_mirrors_the_reference():
    # periodic dimension too short for the template: the reference throws "Multiple interactions"
    lat = Lattice([Material("A", 1.0)], np.eye(3), [("A", (0, 0, 0))], (16, 2, 8))
    w = dict(lattice=lat, hamiltonians=[dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 1e-21)])], temperature=0.0)
    with pytest.raises(capi.JamsB200Error, match="Multiple interactions"):
        make(w).run(1)
    c = capi.Context((4, 4, 4))
    with pytest.raises(capi.JamsB200Error, match="no spins"):
        c.step(1, 1e-4)
    with pytest.raises(RuntimeError, match="Unsupported anisotropy"):
        create_hamiltonian(dict(module="uniaxial", order="K7", anisotropies=[]), lat)
    with pytest.raises(RuntimeError, match="unknown hamiltonian"):
        create_hamiltonian(dict(module="dipole-fft"), lat)


def test_a_second_hamiltonian_of_one_kind_is_refused_not_dropped():
    """the reference sums any number of Hamiltonians (core/solver.cc:43-57); the fused kernels hold one exchange list, one Zeeman field
    ... so a second one must fail loudly instead of silently replacing the first.  Uniaxial terms have three slots."""
    from jams_b200.solver import create_solver
    w = W.c3_sc(dims=(4, 4, 4))
    lat = w["lattice"]
    s = create_solver(dict(module="llg-heun-b200-gpu", t_step=1e-16, t_max=1e-12), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    with pytest.raises(RuntimeError, match="same kind of term"):
        s.register_hamiltonian(create_hamiltonian(dict(module="zeeman", dc_local_field=[[0.0, 0.0, 0.5]]), lat))
    with pytest.raises(RuntimeError, match="same kind of term"):
        s.register_hamiltonian(create_hamiltonian(dict(module="exchange", interactions=[("A", "A", [0.0, 1.0, 0.0], 1e-21)]), lat))
    for order in ("K1", "K2", "K3"):
        s.register_hamiltonian(create_hamiltonian(dict(module="uniaxial", order=order, anisotropies=[("A", [0.0, 0.0, 1.0], 1e-23)]), lat))
    with pytest.raises(RuntimeError, match="more than 3 uniaxial"):
        s.register_hamiltonian(create_hamiltonian(dict(module="uniaxial", order="K1", anisotropies=[("A", [1.0, 0.0, 0.0], 1e-23)]), lat))


@pytest.mark.parametrize("solver_module", ["llg-heun-b200-gpu", "llg-rk4-b200-gpu"])
def test_several_uniaxial_hamiltonians_are_summed_like_the_reference(solver_module):
    """K1 + K2 + K3 written as three "uniaxial" modules (one power each, uniaxial_anisotropy.cc:89-114) on a two-material bcc lattice,
    different axes, the K2 module on one material only: per-term fields and energies, the total field, T = 0 and same-noise T > 0
    trajectories against the oracle, which sums its terms like Solver::compute_fields.  Slots 1 and 2 live in a table only the
    direct-gather kernels read, so the step must not run on a TMA kernel."""
    from jams_b200.solver import create_solver
    lat = Lattice([Material("A", 2.0, alpha=0.05), Material("B", 1.2, alpha=0.2)], np.eye(3), [("A", (0, 0, 0)), ("B", (0.5, 0.5, 0.5))], (6, 5, 8))
    hams = [dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 4e-23), ("B", [1.0, 0.0, 0.0], 2e-23)]),
            dict(module="exchange", interactions=[("A", "B", [0.5, 0.5, 0.5], 3.0e-21), ("B", "A", [0.5, 0.5, 0.5], 3.0e-21)]),
            dict(module="uniaxial", order="K2", anisotropies=[("B", [0.0, 0.6, 0.8], 3e-23)]),
            dict(module="uniaxial", order="K3", anisotropies=[(1, [1.0, 1.0, 1.0], 1e-23), (2, [0.0, 1.0, 0.0], -2e-23)])]
    w = dict(name="three uniaxial", lattice=lat, hamiltonians=hams, spins=None, temperature=0.0)
    rk4 = "rk4" in solver_module
    s0 = random_unit_spins(lat.num_spins, 77)
    s = create_solver(dict(module=solver_module, t_step=1e-16, t_max=1e-9, seed=5), lat)
    for h in hams:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    assert [h.term for h in s.hamiltonians] == [capi.TERM_UNIAXIAL, capi.TERM_EXCHANGE, capi.TERM_UNIAXIAL_2, capi.TERM_UNIAXIAL_3]
    s.set_spins(s0)
    # the oracle takes the terms one by one (helpers.build_cpu_sim keys them by module name: assemble it here)
    from helpers import ref_material_arrays, ref_uniaxial_arrays, oracle_exchange_pairs
    mus, gyro, alpha = ref_material_arrays(lat)

    def cpu_sim(T=0.0):
        sim = oracle.CpuSim(mus, gyro, alpha)
        ids = []
        for hs in hams:
            if hs["module"] == "exchange":
                i, j, J9, _ = oracle_exchange_pairs(lat, hs)
                ids.append(sim.add_exchange(i, j, J9))
            else:
                ids.append(sim.add_uniaxial(*ref_uniaxial_arrays(lat, hs)))
        sim.init_solver(1e-4, lat.gilbert_prefactor, 1)
        sim.set_temperature(T)
        return sim, ids

    sim, ids = cpu_sim()
    sim.set_spins(s0)
    total = np.zeros_like(s0)
    for h, tid in zip(s.hamiltonians, ids):
        ref_f = sim.term_fields(tid, 0.0)
        f = h.calculate_fields(0.0)
        assert np.abs(f - ref_f).max() <= 1e-13 * np.abs(ref_f).max(), h.settings
        total += ref_f
        e, tot = s.ctx.energies(h.term, 0.0)
        assert np.abs(e - sim.term_energies(tid)).max() <= 1e-12 * np.abs(e).max(), h.settings
        assert abs(tot - sim.term_total_energy(tid, 0.0)) <= 1e-12 * abs(tot), h.settings
    assert np.abs(s.compute_fields() - total).max() <= 1e-13 * np.abs(total).max()
    steps = 30
    (sim.run_rk4 if rk4 else sim.run)(steps)
    s.run(steps)
    assert s.ctx.stage_kernel() in (-1, 0)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL
    if not rk4:
        T, seed = 90.0, 5
        s.set_temperature(T)
        s.set_spins(s0)
        it0 = s.iteration
        normals = np.stack([s.ctx.noise(s.step_size, T, seed, it0 + n, normals_only=True) for n in range(10)])
        sim2, _ = cpu_sim(T)
        sim2.set_spins(s0)
        sim2.run(10, normals)
        s.run(10)
        assert np.abs(s.spins() - sim2.get_spins()).max() <= TRAJ_TOL
    # clearing the extra slots gives the TMA kernel back
    s.ctx.set_uniaxial(0, None, None, slot=1); s.ctx.set_uniaxial(0, None, None, slot=2)
    s.ctx.step(1, s.step_size, s.time, 0.0, 0, s.iteration)
    if not rk4:
        assert s.ctx.stage_kernel() == 2


# ---- RK4-LLG: jb_step_rk4 against the restatement of CudaRK4BaseSolver::run ----
def _rk4_solver(w, **kw):
    from jams_b200.solver import create_solver
    lat = w["lattice"]
    s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=W.T_STEP, t_max=1e-9, seed=kw.get("seed", 0), options=kw.get("options", {})), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_temperature(w.get("temperature", 0.0))
    return s


RK4_KERNELS = {"ring": None, "ring_small_tile": dict(tile_y=2, tile_z=16, chunks=3), "direct": dict(kernel=0)}   # the four stages on the TMA pair kernel / direct gathers


@pytest.mark.parametrize("variant", list(RK4_KERNELS))
@pytest.mark.parametrize("make_w,steps", [(lambda: W.c3_sc(dims=(12, 9, 20)), 40), (lambda: W.c2_bcc_fe(6, temperature=0.0), 30),
                                          (lambda: W.c1_bloch_wall((32, 6, 6)), 40), (lambda: W.c4_bcc_long_range(8), 8),
                                          (lambda: W.c3_sc(dims=(7, 13, 70)), 12), (lambda: W.c3_sc(dims=(40, 32, 128)), 6)])
def test_rk4_T0_trajectories_match_oracle(make_w, steps, variant):
    w = make_w()
    if "sc 12" in w["name"] or w["name"].startswith("C3"):
        w["hamiltonians"].append(dict(module="uniaxial", order="K2", anisotropies=[("A", [0.0, 0.6, 0.8], 2e-23)]))
    lat = w["lattice"]
    s0 = w["spins"] if w.get("spins") is not None else random_unit_spins(lat.num_spins, 21)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    sim.run_rk4(steps)
    s = _rk4_solver(w, options=RK4_KERNELS[variant] or {})
    s.set_spins(s0)
    s.run(steps)
    got = s.spins()
    if variant == "direct" or (variant == "ring" and "C4" in w["name"]):   # C4's deep template is beyond the pair kernel's default tiles: direct gathers
        assert s.ctx.stage_kernel() == 0
    elif variant == "ring":
        assert s.ctx.stage_kernel() == 2
    assert np.abs(got - sim.get_spins()).max() <= TRAJ_TOL
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() < 1e-14
    # a Heun step afterwards continues from the RK4 state (shared device state)
    s.ctx.step(3, s.step_size, s.time, 0.0, 0, s.iteration)
    sim.run(3)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL


@pytest.mark.parametrize("variant", list(RK4_KERNELS))
def test_rk4_thermal_trajectory_matches_oracle_given_the_same_noise_and_ac_field(variant):
    """one noise draw per step for all four stages (cuda_rk4_base.cu:65); AC Zeeman field evaluated at t0, t0 + dt/2, t0 + dt"""
    lat = Lattice([Material("A", 2.0, alpha=0.05)], np.eye(3), [("A", (0, 0, 0))], (8, 7, 10))
    w = dict(name="sc ac", lattice=lat, temperature=40.0, spins=None,
             hamiltonians=[dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]),
                           dict(module="zeeman", dc_local_field=[[0.0, 0.0, 0.5]], ac_local_field=[[2.0, 0.0, 0.0]], ac_local_frequency=[0.5])])
    steps, seed = 20, 4321
    s = _rk4_solver(w, seed=seed, options=RK4_KERNELS[variant] or {})
    s0 = random_unit_spins(lat.num_spins, 9)
    s.set_spins(s0)
    normals = np.stack([s.ctx.noise(s.step_size, 40.0, seed, n, normals_only=True) for n in range(steps)])
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    sim.run_rk4(steps, normals)
    s.run(steps)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL


def test_rk4_full_size_properties_and_fixed_point():
    w = W.c3_sc(dims=(96, 96, 96))
    lat = w["lattice"]
    s = _rk4_solver(w)
    up = np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1))
    s.set_spins(up)
    s.run(5)
    assert np.array_equal(s.spins(), up)
    s0 = lat.initial_spins(seed=2)
    s.set_spins(s0)
    e0 = sum(h.calculate_total_energy(0.0) for h in s.hamiltonians)
    s.run(20)
    out = s.spins()
    e1 = sum(h.calculate_total_energy(0.0) for h in s.hamiltonians)
    assert np.abs(np.linalg.norm(out, axis=1) - 1.0).max() < 1e-14 and e1 < e0
    d = _rk4_solver(w, options=dict(kernel=0))   # the ring and the direct kernels agree to rounding at a size the oracle does not run
    d.set_spins(s0)
    d.run(20)
    assert s.ctx.stage_kernel() == 2 and d.ctx.stage_kernel() == 0
    assert np.abs(d.spins() - out).max() <= 1e-13
    # Heun from the same start stays within its own (second-order) error of the RK4 trajectory
    h = make(w)
    h.set_spins(s0)
    h.run(20)
    assert np.abs(h.spins() - out).max() < 1e-4


# ---- physics hook on the device: pinned_boundaries, i.e. the shipped example's own set-up ----
@pytest.mark.parametrize("module", ["llg-rk4-b200-gpu", "llg-heun-b200-gpu"])
def test_pinned_boundaries_with_the_bloch_wall_example_matches_oracle(module):
    """examples/bloch_domain_wall as shipped (bloch_domain_wall.cfg:75-81,132-139): RK4 + pinned_boundaries, here at T = 0 on a
    smaller box; every iteration: update_physics_module (rotate the two edge regions), then run (core/jams++.cc:334-341)"""
    from jams_b200.solver import create_solver, create_physics
    w = W.c1_bloch_wall((32, 6, 5))
    lat = w["lattice"]
    phys = dict(module="pinned_boundaries", left_pinned_magnetisation=[0.0, 0.0, -1.0], right_pinned_magnetisation=[0.0, 0.0, 1.0],
                left_pinned_cells=3, right_pinned_cells=2)
    s = create_solver(dict(module=module, t_step=W.T_STEP, t_max=1e-9), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    p = create_physics(phys, lat)
    s.register_physics_module(p)
    s0 = random_unit_spins(lat.num_spins, 31) * 0.2 + w["spins"]      # a disturbed wall, so the rotations are not the identity
    s0 /= np.linalg.norm(s0, axis=1, keepdims=True)
    s.set_spins(s0)
    sim = build_cpu_sim(w)
    cur = s0.copy()
    left = p.region_sites(0, False, 3, 0, lat.dims[0]); right = p.region_sites(0, True, 2, 0, lat.dims[0])
    assert len(left) == 3 * 6 * 5 and len(right) == 2 * 6 * 5
    steps = 15
    for n in range(steps):
        s.update_physics_module()
        s.run(1)
        cur = oracle.pin_region(cur, lat.mus(), left, [0.0, 0.0, -1.0])
        cur = oracle.pin_region(cur, lat.mus(), right, [0.0, 0.0, 1.0])
        sim.set_spins(cur)
        (sim.run_rk4 if "rk4" in module else sim.run)(1)
        cur = sim.get_spins()
    got = s.spins()
    assert np.abs(got - cur).max() <= TRAJ_TOL
    # the device reduction equals the host loop
    m4 = s.ctx.region_moment(0)
    want = (lat.mus()[left, None] * got[left]).sum(axis=0)
    assert np.abs(m4[:3] - want).max() <= 1e-12 * np.abs(want).max() and abs(m4[3] - lat.mus()[left].sum()) <= 1e-12 * m4[3]


# ---- T > 0: statistics (BASELINE.json north_star, third correctness check) ----
@pytest.mark.parametrize("T", [150.0, 300.0])
def test_thermal_equilibrium_statistics_match_the_reference_arithmetic_and_the_thermostat(T):
    """The reference's CPU (pcg + std::normal_distribution) and GPU (XORWOW) noise streams differ, so T > 0 parity is
    statistical: <m_z>, <E>/N and the spin temperature sum|s x H|^2 / (2 kB sum s.H) (monitors/spin_temperature.cc:22-37) of
    a Philox-driven GPU run against an independent run of the oracle with its own generator, and against the thermostat
    temperature.  gilbert_prefactor = true makes the reference's sigma satisfy the fluctuation-dissipation relation exactly
    (with false the LL form runs at T (1 + alpha^2), in the reference and here alike).  Error bars: the run-to-run
    spread of the oracle at these sizes is 0.002 in m_z, 0.06 meV in E/N and 0.3 K in T_s (two seeds, measured)."""
    lat = Lattice([Material("A", 2.0, alpha=0.5)], np.eye(3), [("A", (0, 0, 0))], (16, 16, 16), gilbert_prefactor=True)
    w = dict(name="stat", lattice=lat, spins=None, temperature=T,
             hamiltonians=[dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]),
                           dict(module="zeeman", dc_local_field=[[0.0, 0.0, 1.0]])])
    dt_s, equil, meas, every = 5e-16, 1000, 2000, 10
    up = np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1))

    def stats(spins_fn, step_fn, energy_fn, field_fn):
        step_fn(equil)
        mz, en, ts = [], [], []
        for _ in range(meas // every):
            step_fn(every)
            s = spins_fn()
            mz.append(s[:, 2].mean()); en.append(energy_fn() / lat.num_spins); ts.append(oracle.spin_temperature(s, field_fn()))
        return np.mean(mz), np.mean(en), np.mean(ts)

    from jams_b200.solver import create_solver
    g = create_solver(dict(module="llg-heun-b200-gpu", t_step=dt_s, t_max=1e-9, seed=2024, gilbert_prefactor=True), lat)
    for h in w["hamiltonians"]:
        g.register_hamiltonian(create_hamiltonian(h, lat))
    g.set_temperature(T)
    g.set_spins(up)
    gm, ge, gt = stats(g.spins, g.run, lambda: sum(h.calculate_total_energy(g.time) for h in g.hamiltonians), g.compute_fields)

    sim = build_cpu_sim(w, dt_ps=dt_s / 1e-12, seed=7)
    sim.set_spins(up)
    cm, ce, ct = stats(sim.get_spins, sim.run, lambda: sum(sim.term_total_energy(t, 0.0) for t in sim.terms.values()),
                       lambda: sum(sim.term_fields(t, 0.0) for t in sim.terms.values()))
    assert abs(gm - cm) <= 0.01, (gm, cm)
    assert abs(ge - ce) <= 0.006 * abs(ce), (ge, ce)
    assert abs(gt - ct) <= 0.02 * T and abs(gt - T) <= 0.03 * T, (gt, ct, T)
    assert 0.0 < gm < 1.0


def test_MT_sweep_with_chained_temperatures_matches_the_reference_arithmetic():
    """BASELINE config 3 is an M(T) sweep: one chained run over temperatures (the state of T_k starts T_k+1), GPU with its Philox
    noise against the oracle with its own generator, through the ordered phase, near and above the Curie temperature
    (k_B T_c = 1.44 J -> 366 K for J = 3.5e-21 J).  Error bars: block standard errors of both runs + the oracle's seed-to-seed
    spread (0.002 ordered, ~0.01 near T_c at this size)."""
    sys_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mt_sweep", os.path.join(sys_path, "mt_sweep.py"))
    mt = importlib.util.module_from_spec(spec); spec.loader.exec_module(mt)
    temps, equil, meas, every, dims = [100.0, 200.0, 300.0, 450.0], 1500, 3000, 10, (16, 16, 16)
    lat, got = mt.sweep(0, temps, equil, meas, every, dims=dims, log=lambda *a: None)
    w = dict(name="mt", lattice=lat, spins=None, temperature=temps[0],
             hamiltonians=[dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]), dict(module="zeeman", dc_local_field=[[0.0, 0.0, 1.0]])])
    sim = build_cpu_sim(w, dt_ps=5e-16 / 1e-12, seed=5)
    sim.set_spins(np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1)))
    prev = 1.0
    for T, g in zip(temps, got):
        sim.set_temperature(T)
        sim.run(equil)
        mz = []
        for _ in range(meas // every):
            sim.run(every)
            mz.append(sim.get_spins()[:, 2].mean())
        cm = float(np.mean(mz))
        tol = 0.01 if T < 250 else 0.03
        assert abs(g["mz"] - cm) <= tol + 3 * g["mz_err"], (T, g, cm)
        assert g["mz"] < prev   # the magnetisation falls monotonically along the sweep
        prev = g["mz"]
    assert got[0]["mz"] > 0.85 and got[-1]["mz"] < 0.25


# ---- edge cases: vacancies, no exchange at all, empty step counts, sizes beyond one slab ----
@pytest.mark.parametrize("variant", ["direct", "pair", "pair_store_u", "pair_small_tile"])
def test_vacancies_stay_zero_and_do_not_act_on_their_neighbours(variant):
    """zero-length spins (vacancies) are left unchanged by unit_vector (containers/vec3.h:276-283) and add nothing to J.s"""
    w = W.c3_sc(dims=(10, 8, 12), temperature=0.0)
    w["hamiltonians"].append(dict(module="uniaxial", order="K1", anisotropies=[("A", [0.0, 0.0, 1.0], 1e-22)]))
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 5)
    holes = np.random.default_rng(3).choice(lat.num_spins, lat.num_spins // 10, replace=False)
    s0[holes] = 0.0
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    sim.run(25)
    s = make(w, options=KERNELS[variant])
    s.set_spins(s0)
    s.run(25)
    got = s.spins()
    assert np.array_equal(got[holes], np.zeros((len(holes), 3)))
    assert np.abs(got - sim.get_spins()).max() <= TRAJ_TOL


def test_zeeman_only_single_spin_precession_and_empty_step_count():
    """no exchange Hamiltonian at all (no template): one spin in a field precesses at gamma B; nsteps = 0 is a no-op"""
    from jams_b200.consts import kGyromagneticRatioIU
    lat = Lattice([Material("A", 1.0, alpha=0.0)], np.eye(3), [("A", (0, 0, 0))], (1, 1, 1), periodic=(False, False, False))
    w = dict(name="one spin", lattice=lat, hamiltonians=[dict(module="zeeman", dc_local_field=[[0.0, 0.0, 10.0]])], temperature=0.0, spins=None)
    s = make(w)
    s.set_spins(np.array([[1.0, 0.0, 0.0]]))
    s.run(0)
    assert np.array_equal(s.spins(), [[1.0, 0.0, 0.0]])
    s.run(1000)
    phi = kGyromagneticRatioIU * 10.0 * 1000 * 1e-4
    out = s.spins()[0]
    assert abs(out[2]) < 1e-12 and abs(out[0] - np.cos(phi)) < 1e-6 and abs(out[1] - np.sin(phi)) < 1e-6


def test_a_lattice_too_large_for_one_slab_is_refused_not_truncated():
    """the reference throws when nnz overflows int32 (containers/sparse_matrix_builder.h:265-269, sc 512^3 already does);
    here the limit is 2^31 box elements per slab and the message says what to do"""
    with pytest.raises(capi.JamsB200Error, match="more ranks|too large|2\\^31"):
        c = capi.Context((1400, 1300, 1300))
        c.set_materials(np.ones(1), np.ones(1), np.ones(1))


# ---- exchange-functional: another producer of the same scalar CSR matrix (hamiltonian/exchange_functional.cc) ----
@pytest.mark.parametrize("variant", ["direct", "pair", "pair_store_u", "pair_small_tile", "pairs", "pairs_auto"])
def test_exchange_functional_fields_energy_and_trajectory_match_oracle(variant):
    """two-material bcc lattice (periodic x and z, open y), J(r) from gaussian / exponential / rkky forms inside cutoffs: the
    oracle integrates with the pair list of a brute-force minimum-image search, the GPU path with the template built by
    the host layer (long range: 32 neighbours per Co site, ghost depth 1)"""
    from test_host_logic import _functional_case
    from helpers import brute_force_functional_pairs
    lat, settings, fns = _functional_case()
    w = dict(name="functional", lattice=lat, hamiltonians=[settings, dict(module="zeeman", dc_local_field=[[0.0, 0.0, 0.5], [0.0, 0.3, 0.0]])],
             temperature=0.0, spins=None, functional_pairs=brute_force_functional_pairs(lat, fns))
    s0 = random_unit_spins(lat.num_spins, 21)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    s = make(w, options=KERNELS.get(variant), pairs={"pairs": True, "pairs_auto": "auto"}.get(variant, False))
    s.set_spins(s0)
    h = s.hamiltonians[0]
    f = h.calculate_fields(0.0)
    want_f = sim.term_fields(sim.terms["exchange-functional"], 0.0)
    assert np.abs(f - want_f).max() <= 1e-13 * np.abs(want_f).max()
    e = h.calculate_total_energy(0.0)
    want_e = sim.term_total_energy(sim.terms["exchange-functional"], 0.0)
    assert abs(e - want_e) <= 1e-12 * abs(want_e)
    sim.run(30)
    s.run(30)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL


# ---- time-dependent applied field (hamiltonian/applied_field.cc:10-82: static, sinc, sinc-cos) ----
@pytest.mark.parametrize("kind", ["sinc", "sinc-cos"])
@pytest.mark.parametrize("variant", ["direct", "pair", "pair_store_u", "pair_small_tile", "pairs"])
def test_applied_field_pulse_fields_energy_and_trajectory_match_oracle(kind, variant):
    """B(t) = B g(t): both Heun stages see the field at their own time (predictor t, corrector t + dt,
    cpu_llg_heun.cc:46,103-104); the pulse is centred inside the run so the amplitude changes sign and size over the steps"""
    w = W.c3_sc(dims=(9, 8, 14), temperature=0.0)
    hs = dict(module="applied-field", type=kind, field=[0.3, -0.2, 1.5], time_center=2.0e-15, freq_bandwidth=4.0e14 if kind == "sinc-cos" else 1.2e15)
    if kind == "sinc-cos":
        hs["freq_center"] = 9.0e14
    w["hamiltonians"].append(hs)
    lat = w["lattice"]
    s0 = random_unit_spins(lat.num_spins, 31)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    s = make(w, options=KERNELS.get(variant), pairs=(variant == "pairs"))
    s.set_spins(s0)
    h = s.hamiltonians[-1]
    assert h.name == "applied-field-" + kind
    for t in (0.0, 1.3e-3, 2.0e-3, 3.7e-3):     # ps
        f, want = h.calculate_fields(t), sim.term_fields(sim.terms["applied-field"], t)
        assert np.abs(f - want).max() <= 1e-13 * max(np.abs(want).max(), 1e-30)
        e, want_e = h.calculate_total_energy(t), sim.term_total_energy(sim.terms["applied-field"], t)
        assert abs(e - want_e) <= 1e-12 * max(abs(want_e), 1.0)
    amp = [sim.term_fields(sim.terms["applied-field"], n * 1e-4)[0, 2] for n in range(40)]
    assert max(amp) > 0 and min(amp) < 0                              # the pulse really varies over the run
    sim.run(40)
    s.run(25)
    s.run(15)                                                         # time is carried across calls
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL


# ---- round 2 hardening (VERDICT r01 "what's weak" 1-6) --------------------------------------------------------------------
_ORACLE_CACHE = {}


def _oracle_run(dims, T, steps, seed, normals):
    """the oracle on the bench workload's Hamiltonian at a size it does in a few seconds per case"""
    w = W.c3_sc(dims=dims, temperature=T)
    sim = build_cpu_sim(w)
    s0 = w["lattice"].initial_spins(seed=3)
    sim.set_spins(s0)
    sim.run(steps, normals)
    return s0, sim.get_spins()


@pytest.mark.parametrize("dims", [(96, 96, 96), (64, 32, 256), (24, 256, 256)])
@pytest.mark.parametrize("T", [0.0, 100.0])
def test_bench_code_path_matches_oracle(dims, T):
    """The configuration the headline runs -- 4 x 128 tiles (128 columns at 256 x 256), more work items than resident CTAs, several
    x-chunks plus the taper, the TMA ring wrapping across item boundaries -- compared with the oracle directly, T = 0 and T > 0
    with the kernels' own noise handed to the oracle; both data flows of the stage kernel and the direct kernel."""
    steps, seed = 5, 97
    w = W.c3_sc(dims=dims, temperature=T)
    lat = w["lattice"]
    normals = None
    if T > 0:
        probe = make(w, seed=seed)
        probe.spins()
        normals = np.stack([probe.ctx.noise(probe.step_size, T, seed, n, normals_only=True) for n in range(steps)])
        probe.ctx.close()
    s0, want = _oracle_run(dims, T, steps, seed, normals)
    for variant in ("pair", "pair_store_u", "direct"):
        s = make(w, options=dict(KERNELS[variant], verbose=0), seed=seed)
        s.set_spins(s0)
        s.run(steps)
        got = s.spins()
        assert np.abs(got - want).max() <= TRAJ_TOL, (variant, dims, T)
        s.ctx.close()


@pytest.mark.parametrize("solver_module", ["llg-heun-b200-gpu", "llg-rk4-b200-gpu"])
def test_biquadratic_exchange_fields_energies_and_trajectory_match_oracle(solver_module):
    """module = "biquadratic-exchange" (hamiltonian/cuda_biquadratic_exchange.{cu,kernel.cuh}): h_i = sum_j 2 B_ij s_j (s_i . s_j) next
    to bilinear exchange and a uniaxial term on bcc (NN + NNN biquadratic shells; a negative coupling is dropped by the reference's
    value > energy_cutoff filter and must be dropped here).  Fields and per-spin energies per term, the reference's total (half the
    sum of the per-spin energies), T = 0 Heun and RK4 trajectories, a same-noise T > 0 Heun trajectory.  The reference has no test or
    CPU field implementation of this term: the oracle restates the CUDA kernel; tests/test_gpu_reference_cuda.py pins both to that kernel itself."""
    from jams_b200.solver import create_solver
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 7), periodic=(True, True, False))
    hams = [dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21)]),
            dict(module="biquadratic-exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 0.8e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 0.3e-21),
                                                             ("Fe", "Fe", [1.0, 1.0, 0.0], -0.2e-21)]),
            dict(module="uniaxial", order="K1", anisotropies=[("Fe", [0.0, 0.0, 1.0], 1e-23)])]
    w = dict(name="bq", lattice=lat, hamiltonians=hams, spins=None, temperature=0.0)
    s0 = random_unit_spins(lat.num_spins, 31)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    rk4 = "rk4" in solver_module
    s = create_solver(dict(module=solver_module, t_step=1e-16, t_max=1e-9, seed=3), lat)
    for h in hams:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    bq = s.hamiltonians[1]
    assert len(bq.template["B"]) == 2 * (8 + 6) and (bq.template["B"] > 0).all()     # the negative shell is gone
    s.set_spins(s0)
    total = np.zeros_like(s0)
    for h in s.hamiltonians:
        key = h.settings["module"].lower()
        ref_f = sim.term_fields(sim.terms[key], 0.0)
        f = h.calculate_fields(0.0)
        assert np.abs(f - ref_f).max() <= 1e-13 * np.abs(ref_f).max(), key
        total += ref_f
        e, tot = s.ctx.energies(h.term, 0.0)
        assert np.abs(e - sim.term_energies(sim.terms[key])).max() <= 1e-12 * np.abs(e).max(), key
        assert abs(tot - sim.term_total_energy(sim.terms[key], 0.0)) <= 1e-12 * abs(tot), key
    assert np.abs(s.compute_fields() - total).max() <= 1e-13 * np.abs(total).max()
    steps = 30
    (sim.run_rk4 if rk4 else sim.run)(steps)
    s.run(steps)
    assert s.ctx.stage_kernel() in (-1, 0)    # direct gathers (-1: the RK4 solver does not go through jb_step)
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL
    if not rk4:   # same-noise T > 0
        T, seed = 120.0, 3
        s.set_temperature(T)
        s.set_spins(s0)
        it0 = s.iteration
        normals = np.stack([s.ctx.noise(s.step_size, T, seed, it0 + n, normals_only=True) for n in range(10)])
        sim2 = build_cpu_sim(dict(w, temperature=T))
        sim2.set_spins(s0)
        sim2.run(10, normals)
        s.run(10)
        assert np.abs(s.spins() - sim2.get_spins()).max() <= TRAJ_TOL


def test_lattice_impurities_random_alloy_matches_oracle():
    """lattice.impurities (core/lattice.cc:614-640): 30 % of the Fe sites of a bcc lattice become Co.  The neighbour list is no longer
    translation invariant (pairs with a substituted end are dropped, core/interactions.cc:381-385): general neighbour-list kernel with
    per-site classes (moment, damping, anisotropy and Zeeman field by the site's own material).  Fields and energies per term, a T = 0
    trajectory and a same-noise T > 0 trajectory against the oracle, which gets the realised site materials as an input."""
    from jams_b200.solver import create_solver
    mats = [Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05)]
    lat = Lattice(mats, np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (8, 7, 9), periodic=(True, True, False),
                  impurities=[("Fe", "Co", 0.3)], impurities_seed=5)
    hams = [dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21)]),
            dict(module="uniaxial", order="K1", anisotropies=[("Fe", [0.0, 0.0, 1.0], 1e-23), ("Co", [1.0, 0.0, 0.0], 4e-23)]),
            dict(module="zeeman", dc_local_field=[[0.0, 0.0, 1.0], [0.0, 0.5, 0.0]])]
    w = dict(name="alloy", lattice=lat, hamiltonians=hams, spins=None, temperature=0.0)
    assert 0.2 < (lat.site_material() == 1).mean() < 0.4
    s0 = random_unit_spins(lat.num_spins, 41)
    sim = build_cpu_sim(w)
    sim.set_spins(s0)
    s = create_solver(dict(module="llg-heun-b200-gpu", t_step=1e-16, t_max=1e-9, seed=9), lat)
    for h in hams:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_spins(s0)
    for h in s.hamiltonians:
        key = h.settings["module"].lower()
        ref_f = sim.term_fields(sim.terms[key], 0.0)
        assert np.abs(h.calculate_fields(0.0) - ref_f).max() <= 1e-13 * np.abs(ref_f).max(), key
        assert abs(h.calculate_total_energy(0.0) - sim.term_total_energy(sim.terms[key], 0.0)) <= 1e-12 * abs(sim.term_total_energy(sim.terms[key], 0.0)), key
    sim.run(25)
    s.run(25)
    assert s.ctx.stage_kernel() == 5   # general neighbour list
    assert np.abs(s.spins() - sim.get_spins()).max() <= TRAJ_TOL
    T, seed = 80.0, 9
    s.set_temperature(T)
    s.set_spins(s0)
    it0 = s.iteration
    normals = np.stack([s.ctx.noise(s.step_size, T, seed, it0 + n, normals_only=True) for n in range(10)])
    sim2 = build_cpu_sim(dict(w, temperature=T))
    sim2.set_spins(s0)
    sim2.run(10, normals)
    s.run(10)
    assert np.abs(s.spins() - sim2.get_spins()).max() <= TRAJ_TOL


def test_exchange_symmetry_guard_mirrors_the_reference():
    """SparseInteractionHamiltonian::finalize throws "sparse matrix for exchange is not symmetric" (hamiltonian/sparse_interaction.cc:114-118,
    containers/sparse_matrix_builder.h:320-362) unless check_sparse_matrix_symmetry = false (hamiltonian/exchange.cc:104-110)"""
    c = capi.Context((6, 6, 6))
    c.set_materials(np.full(216, 0.1), np.full(216, 0.17), np.full(216, 0.1))
    eye = np.eye(3).reshape(9)
    mi, mj = np.zeros(2, np.int32), np.zeros(2, np.int32)
    T = np.array([[1, 0, 0], [-1, 0, 0]], np.int32)
    c.set_exchange_template(mi, mj, T, np.stack([eye, eye]))                        # symmetric: accepted
    with pytest.raises(capi.JamsB200Error, match="sparse matrix for exchange is not symmetric"):
        c.set_exchange_template(mi, mj, T, np.stack([eye, 2 * eye]))               # J_ij != J_ji
    with pytest.raises(capi.JamsB200Error, match="not symmetric"):
        c.set_exchange_template(mi[:1], mj[:1], T[:1], eye[None, :])                # (j, i) missing
    dm = np.array([0, 1.0, 0, -1.0, 0, 0, 0, 0, 0])                                  # antisymmetric (DM-like) tensor: J_ji = J_ij^T is symmetric overall
    c.set_exchange_template(mi, mj, T, np.stack([dm, -dm]))
    with pytest.raises(capi.JamsB200Error, match="not symmetric"):
        c.set_exchange_template(mi, mj, T, np.stack([dm, dm]))
    c.set_option("check_symmetry", 0)
    c.set_exchange_template(mi, mj, T, np.stack([eye, 2 * eye]))                    # the reference's check_sparse_matrix_symmetry = false
    c.set_option("check_symmetry", 1)
    # the general list (no translation invariance -> ELL path): a one-way bond is refused too
    c.set_option("detect_template", 0)
    i = np.array([0, 1, 5], np.int32); j = np.array([1, 0, 7], np.int32)
    with pytest.raises(capi.JamsB200Error, match="not symmetric"):
        c.set_exchange_pairs(i, j, np.zeros(3, np.int32), eye[None, :])
    c.set_exchange_pairs(i[:2], j[:2], np.zeros(2, np.int32), eye[None, :])
    # through the plugin surface: the setting reaches the library
    lat = Lattice([Material("A", 1.0)], np.eye(3), [("A", (0, 0, 0))], (6, 6, 6))
    hs = dict(module="exchange", symops=False, interactions=[("A", "A", [1.0, 0.0, 0.0], 1e-21), ("A", "A", [-1.0, 0.0, 0.0], 2e-21)])
    with pytest.raises(capi.JamsB200Error, match="not symmetric"):
        make(dict(lattice=lat, hamiltonians=[hs], temperature=0.0)).run(1)
    make(dict(lattice=lat, hamiltonians=[dict(hs, check_sparse_matrix_symmetry=False)], temperature=0.0)).run(1)


def test_device_pointer_abi_path():
    """`on_device = 1`: what the JAMS adapter passes (MultiArray::device_data(), containers/multiarray.h:239-253) -- import, export,
    fields and noise with device pointers of another owner (a torch tensor) against the host-pointer path"""
    import torch
    w = W.c2_bcc_fe(6, temperature=0.0)
    w["hamiltonians"].append(dict(module="uniaxial", order="K1", anisotropies=[("Fe", [0.0, 0.6, 0.8], 3e-23)]))
    lat = w["lattice"]
    s = make(w, seed=11)
    s0 = random_unit_spins(lat.num_spins, 4)
    d_in = torch.from_numpy(s0).cuda()
    torch.cuda.synchronize()
    s._build()
    s.ctx.import_spins_ptr(d_in.data_ptr(), 1)
    s.ctx.synchronize()
    assert np.array_equal(s.ctx.export_spins(), s0)
    d_out = torch.zeros_like(d_in)
    s.ctx.export_spins_ptr(d_out.data_ptr(), 1)
    s.ctx.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), s0)
    # jb_fields into a device array (globals::h on the device)
    import ctypes as C
    d_h = torch.zeros_like(d_in)
    s.ctx._ck(s.ctx.lib.jb_fields(s.ctx.h, capi.TERM_TOTAL, 0.0, C.c_void_p(d_h.data_ptr()), 1))
    s.ctx.synchronize()
    assert np.array_equal(d_h.cpu().numpy(), s.ctx.fields(capi.TERM_TOTAL, 0.0))
    # Thermostat::device_data
    d_n = torch.zeros_like(d_in)
    s.ctx._ck(s.ctx.lib.jb_noise(s.ctx.h, 1e-4, 50.0, 11, 3, 0, 0, C.c_void_p(d_n.data_ptr()), 1))
    s.ctx.synchronize()
    assert np.array_equal(d_n.cpu().numpy(), s.ctx.noise(1e-4, 50.0, 11, 3))
    # a trajectory started from the device import equals one started from the host import
    s.run(10)
    a = s.spins()
    s2 = make(w, seed=11)
    s2.set_spins(s0)
    s2.run(10)
    assert np.array_equal(a, s2.spins())


def test_noise_distribution_in_depth():
    """6.3 M draws of the pair-keyed Philox / Box-Muller stream (jb_device.cuh): Kolmogorov-Smirnov and chi-square against N(0,1),
    moments to order 8, the hard cap of the 23-bit radius, and the independence structure inside a pair of sites -- the two
    members of a Box-Muller pair share their radius, so besides their correlation the correlation of their SQUARES must vanish"""
    from scipy import stats
    lat = Lattice([Material("A", 1.0)], np.eye(3), [("A", (0, 0, 0))], (128, 128, 128))
    w = dict(lattice=lat, hamiltonians=[dict(module="zeeman", dc_local_field=[[0, 0, 1.0]])], temperature=10.0)
    s = make(w, seed=2026)
    s.spins()
    x = s.ctx.noise(1e-4, 10.0, 2026, 17, normals_only=True)      # (N, 3)
    flat = x.ravel()
    n = flat.size
    assert stats.kstest(flat[::3], "norm").pvalue > 1e-3 and stats.kstest(flat[1::5], "norm").pvalue > 1e-3
    edges = np.linspace(-4.0, 4.0, 81)
    obs, _ = np.histogram(flat, bins=np.concatenate([[-np.inf], edges, [np.inf]]))
    exp = np.diff(stats.norm.cdf(np.concatenate([[-np.inf], edges, [np.inf]]))) * n
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    assert chi2 < stats.chi2.ppf(1 - 1e-4, df=len(obs) - 1), chi2
    for k, mk, var in ((1, 0.0, 1.0), (2, 1.0, 2.0), (3, 0.0, 15.0), (4, 3.0, 96.0), (6, 15.0, 10170.0), (8, 105.0, 2016000.0)):
        assert abs((flat ** k).mean() - mk) < 5 * np.sqrt(var / n), k
    assert 4.9 < np.abs(flat).max() <= np.sqrt(2 * 23 * np.log(2)) + 1e-6      # radius uniform has 23 bits
    pairs = x.reshape(-1, 2, 3)                                      # (z even, z odd) of every row: one Philox call each
    e, o = pairs[:, 0, :], pairs[:, 1, :]
    m = e.shape[0]
    bm = ((e[:, 0], e[:, 1]), (e[:, 2], o[:, 0]), (o[:, 1], o[:, 2]))   # the three Box-Muller pairs
    other = ((e[:, 0], e[:, 2]), (e[:, 1], o[:, 1]), (e[:, 0], o[:, 2]), (o[:, 0], o[:, 1]))
    for a, b in bm + other:
        assert abs(np.mean(a * b)) < 5 / np.sqrt(m)
        assert abs(np.mean((a * a - 1) * (b * b - 1))) < 5 * 2 / np.sqrt(m)
    # neighbouring pairs / rows / steps are independent calls
    assert abs(np.mean(e[:-1, 0] * e[1:, 0])) < 5 / np.sqrt(m)
    y = s.ctx.noise(1e-4, 10.0, 2026, 18, normals_only=True).ravel()
    assert abs(np.mean(flat * y)) < 5 / np.sqrt(n)
