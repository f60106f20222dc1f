"""CPU tests of the host side: lattice / template construction bit-exact with the oracle, the C-ABI
library loads and exports every declared symbol, the product refuses to run without a GPU, and the
multi-rank plumbing works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from helpers import oracle_exchange_pairs
from jams_b200 import capi, workloads as W
from jams_b200.distributed import ring_neighbours, slab_range
from jams_b200.lattice import Lattice, Material
from jams_b200.solver import create_hamiltonian

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sorted_template(t):
    key = np.lexsort((t["T"][:, 2], t["T"][:, 1], t["T"][:, 0], t["mj"], t["mi"]))
    return t["mi"][key], t["mj"][key], t["T"][key], t["J9"][key]


@pytest.mark.parametrize("make", [lambda: W.c1_bloch_wall((16, 4, 4)), lambda: W.c2_bcc_fe(4), lambda: W.c3_sc(dims=(5, 4, 6)),
                                  lambda: W.c4_bcc_long_range(6)])
def test_template_and_neighbour_list_bit_exact_with_oracle(make):
    w = make()
    lat = w["lattice"]
    hs = next(h for h in w["hamiltonians"] if h["module"] == "exchange")
    h = create_hamiltonian(hs, lat)
    oi, oj, oJ9, otmpl = oracle_exchange_pairs(lat, hs)
    # processed template: same entries (order of symmetric copies is irrelevant, the list is kept sorted)
    unit = h.input_energy_unit_conversion
    a = _sorted_template(h.template)
    b = _sorted_template(dict(mi=otmpl["mi"], mj=otmpl["mj"], T=otmpl["T"], J9=otmpl["J9"] * unit))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # neighbour list: identical pairs, order, and tensors
    i, j, v, vals = h.neighbour_list()
    assert np.array_equal(i, oi) and np.array_equal(j, oj) and np.array_equal(vals[v], oJ9)


@pytest.mark.parametrize("make", [lambda: W.c1_bloch_wall((16, 4, 4)), lambda: W.c2_bcc_fe(4), lambda: W.c3_sc(dims=(5, 4, 6)),
                                  lambda: W.c4_bcc_long_range(6)])
def test_translation_invariant_neighbour_list_is_recognised(make):
    """jb_detect_exchange_template (host only): ExchangeHamiltonian::neighbour_list() -> the template it came from"""
    from jams_b200 import capi
    w = make()
    lat = w["lattice"]
    h = create_hamiltonian(next(hs for hs in w["hamiltonians"] if hs["module"] == "exchange"), lat)
    i, j, v, vals = h.neighbour_list()
    t = capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i, j, v, vals)
    assert t is not None
    for x, y in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    # a slab sees the same template from the global list
    nx = lat.dims[0] // 2
    t2 = capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i, j, v, vals, x_begin=nx, nx_local=lat.dims[0] - nx)
    for x, y in zip(_sorted_template(t2), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    # a vacancy (one pair missing), a changed coupling, or too small a capacity: not a template
    keep = np.ones(i.size, bool); keep[i.size // 3] = False
    assert capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i[keep], j[keep], v[keep], vals) is None
    vals2 = np.concatenate([vals, vals[:1] * 1.5]); v2 = v.copy(); v2[i.size // 2] = len(vals)
    assert capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i, j, v2, vals2) is None
    assert capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i, j, v, vals, capacity=2) is None
    with pytest.raises(capi.JamsB200Error):
        capi.detect_exchange_template(lat.dims, lat.M, lat.periodic, i, j + lat.num_spins, v, vals)


def test_c4_template_has_112_neighbours_per_site():
    w = W.c4_bcc_long_range(6)
    h = create_hamiltonian(w["hamiltonians"][0], w["lattice"])
    assert np.array_equal(np.bincount(h.template["mi"]), [112, 112])


def test_site_arrays_follow_reference_numbering():
    lat = Lattice([Material("A", 2.0, alpha=0.05), Material("B", 1.0, alpha=0.2)], np.eye(3),
                  [("A", (0, 0, 0)), ("B", (0.5, 0.5, 0.5))], (3, 4, 5))
    assert lat.site_index(2, 3, 4, 1) == lat.num_spins - 1
    assert np.array_equal(lat.site_material()[:4], [0, 1, 0, 1])
    pos = lat.positions()
    assert np.allclose(pos[lat.site_index(1, 2, 3, 1)], [1.5, 2.5, 3.5])
    # slab slices are contiguous ranges of the global order
    assert np.array_equal(lat.mus(1, 2), lat.mus()[1 * 40:3 * 40])
    assert np.array_equal(lat.initial_spins(1, 2, seed=3), lat.initial_spins(seed=3)[40:120])


def test_slab_and_ring_helpers():
    assert slab_range(512, 3, 8) == (192, 64)
    with pytest.raises(RuntimeError):
        slab_range(10, 0, 4)
    assert ring_neighbours(0, 4, True) == (3, 1) and ring_neighbours(0, 4, False) == (None, 1)
    assert ring_neighbours(3, 4, False) == (2, None) and ring_neighbours(0, 1, True) == (None, None)
    assert ring_neighbours(1, 2, True) == (0, 0)


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jams_b200.h")).read()
    declared = set(re.findall(r"JB_API [^;]*?\b(jb_\w+)\(", header))
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    lib = capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.jb_abi_version() == 1
    # the library is the in-tree build, not something on the system path
    assert os.path.dirname(capi.LIB_PATH) == os.path.join(ROOT, "jams_b200")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.JamsB200Error, match="no CUDA device"):
        capi.Context((4, 4, 4))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jams_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libjams_oracle" not in text and "libjams_ref" not in text, f


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from jams_b200.distributed import TorchComm, ring_neighbours, slab_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
comm = TorchComm(periodic_x=True)
blobs = comm.all_gather_bytes(bytes([comm.rank]) * 256)
assert [b[0] for b in blobs] == [0, 1] and all(len(b) == 256 for b in blobs)
tot = comm.allreduce_sum(np.array([[1.0 + comm.rank, 2.0, 3.0, 4.0]]))
assert tot.shape == (1, 4) and tot[0, 0] == 3.0 and tot[0, 3] == 8.0
assert comm.allreduce_max(float(comm.rank)) == 1.0
x0, nx = slab_range(8, comm.rank, comm.world_size)
assert (x0, nx) == (4 * comm.rank, 4)
assert ring_neighbours(comm.rank, 2, True) == (1 - comm.rank, 1 - comm.rank)
comm.barrier()
dist.destroy_process_group()
print("ok", comm.rank)
"""


def test_gloo_world_size_2_plumbing(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_pinned_boundary_regions_follow_the_reference_predicates():
    """physics/pinned_boundaries.h:86-107: lower boundary cell[dim] < n, upper boundary cell[dim] >= size[dim] - n;
    indices in the reference site order, per slab when the lattice is split along x"""
    from jams_b200.lattice import Lattice, Material
    from jams_b200.solver import PinnedBoundariesPhysics
    lat = Lattice([Material("A", 1.0)], np.eye(3), [("A", (0, 0, 0)), ("A", (0.5, 0.5, 0.5))], (8, 3, 4), periodic=(False, True, True))
    p = PinnedBoundariesPhysics(dict(module="pinned_boundaries", left_pinned_magnetisation=[0, 0, -1], top_pinned_magnetisation=[1, 0, 0],
                                     left_pinned_cells=2), lat)
    assert [b[0] for b in p.boundaries] == ["left", "top"] and p.boundaries[0][3] == 2 and p.boundaries[1][3] == 1
    left = p.region_sites(0, False, 2, 0, 8)
    assert len(left) == 2 * 3 * 4 * 2 and left.max() == 2 * 3 * 4 * 2 - 1          # the first two x planes are the first sites
    top = p.region_sites(2, True, 1, 0, 8)
    z = (top // lat.M) % 4
    assert len(top) == 8 * 3 * 2 and np.all(z == 3)
    # slab 1 of 2 (x in [4, 8)): the left region is empty, the right one is the last plane of the slab in local numbering
    assert len(p.region_sites(0, False, 2, 4, 4)) == 0
    right = p.region_sites(0, True, 1, 4, 4)
    assert len(right) == 3 * 4 * 2 and right.min() == 3 * 3 * 4 * 2


def test_spin_snapshot_text_format_round_trips_exactly(tmp_path):
    from jams_b200.lattice import load_spins_tsv, write_spins_tsv
    s = np.random.default_rng(1).standard_normal((37, 3))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    write_spins_tsv(tmp_path / "a.tsv", s, iteration=12, time_ps=1.2e-3)
    assert np.array_equal(load_spins_tsv(tmp_path / "a.tsv", 37), s)
    with pytest.raises(RuntimeError, match="expected size"):
        load_spins_tsv(tmp_path / "a.tsv", 36)
    with pytest.raises(RuntimeError, match="failed to open file"):
        load_spins_tsv(tmp_path / "nope.tsv", 37)


# ---- physics modules that ramp the temperature (core/solver.cc:94-97 re-reads T every step) ----
class _Clock:
    """what Physics::update gets from the solver: iteration, time (ps) and the step (core/solver.cc:85-87)"""
    def __init__(self, dt):
        self.iteration, self.time, self.step_size = 0, 0.0, dt

    def tick(self):
        self.iteration += 1
        self.time = self.iteration * self.step_size


def test_field_cool_physics_ramps_like_the_reference():
    """physics/field_cool.cc:56-78: continuous mode adds (final - init) * sim.t_step / CoolTime per iteration while
    t_eq < time < CoolTime; TSteps mode holds plateaus of CoolTime / TSteps after sim.t_eq"""
    from jams_b200.solver import create_physics
    clk = _Clock(1e-4)
    p = create_physics(dict(module="field-cool", InitialTemperature=300.0, FinalTemperature=100.0, InitialField=[0.0, 0.0, 1.0],
                            FinalField=[0.0, 0.0, 0.0], CoolTime=0.01, applied_field=[0.5, 0.0, 0.0]), None, dict(t_step=1e-4))
    assert p.temperature == 300.0 and np.array_equal(p.applied_field, [0.5, 0.0, 1.0])
    temps = []
    for _ in range(150):
        p.update(clk)          # main loop order: update_physics_module, then run (core/jams++.cc:334-341)
        temps.append(p.temperature)
        clk.tick()
    want, T = [], 300.0
    for n in range(150):
        t = n * 1e-4
        if t > 0.0 and t < 0.01:
            T += (100.0 - 300.0) * 1e-4 / 0.01
        want.append(T)
    assert temps == want and abs(temps[-1] - 102.0) < 1e-9      # 99 increments: time = 0 is not > t_eq, time = CoolTime is not < CoolTime
    assert abs(p.applied_field[2] - 0.01) < 1e-12 and p.applied_field[0] == 0.5
    clk = _Clock(1e-4)
    q = create_physics(dict(module="field-cool", InitialTemperature=300.0, FinalTemperature=100.0, InitialField=[0, 0, 0],
                            FinalField=[0, 0, 0], CoolTime=0.01, TSteps=4), None, dict(t_step=1e-4, t_eq=0.002))
    seen = []
    for _ in range(200):
        q.update(clk)
        seen.append(q.temperature)
        clk.tick()
    assert seen[:21] == [300.0] * 21                           # time <= t_eq
    assert seen[46] == 300.0 - 1 * 50.0 and seen[71] == 300.0 - 2 * 50.0 and seen[199] == 100.0
    assert sorted(set(seen), reverse=True) == [300.0, 250.0, 200.0, 150.0, 100.0]


def test_two_temperature_model_physics_follows_the_reference_recursion():
    """physics/two_temperature_model.cc:66-95 (forward Euler with the solver's step; the thermostat follows T_electron)"""
    from jams_b200.solver import create_physics
    cfg = dict(module="two-temperature-model", InitialTemperature=300.0, ReversingField=[0.0, 0.0, -0.2], Ce=700.0, Cl=3.0e6,
               Gep=1.7e6, Gps=1.7e3, output_steps=10, laserPulses=[dict(width=0.05, fluence=4.0e-11, t_start=0.01), dict(width=0.02, fluence=1.0e-11, t_start=0.3)])
    p = create_physics(cfg, None)
    clk = _Clock(1e-3)
    Te = Tp = Ts = 300.0
    for _ in range(800):
        p.update(clk)
        pump = 0.0
        for w, f, t0 in ((0.05, 1.152e20 * 4.0e-11, 0.01), (0.02, 1.152e20 * 1.0e-11, 0.3)):
            rel = clk.time - t0
            if 0.0 < rel <= 10 * w:
                pump += f * np.exp(-((rel - 3 * w) / w) ** 2)
        Te = Te + ((-1.7e6 * (Te - Tp) + pump) * 1e-3) / (700.0 * Te)
        Tp = Tp + ((1.7e6 * (Te - Tp) - 1.7e3 * (Tp - Ts)) * 1e-3) / 3.0e6
        assert p.temperature == Te and p.phonon_temp == Tp
        clk.tick()
    assert max(r[1] for r in p.records) > 320.0 and len(p.records) == 80 and np.array_equal(p.applied_field, [0.0, 0.0, -0.2])


# ---- exchange-functional (another producer of the same scalar CSR matrix) ----
def _functional_case():
    from jams_b200.lattice import Lattice, Material
    lat = Lattice([Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05)], np.eye(3), [("Fe", (0, 0, 0)), ("Co", (0.5, 0.5, 0.5))], (5, 4, 6),
                  periodic=(True, False, True))
    settings = dict(module="exchange-functional", energy_units="meV",
                    interactions=[("Fe", "Fe", "gaussian", 1.01, 12.0, 1.0, 0.3), ("Fe", "Co", "exponential", 0.9, 20.0, 0.8, 0.25),
                                  ("Co", "Fe", "exponential", 0.9, 20.0, 0.8, 0.25), ("Co", "Co", "rkky", 1.45, 3.0, 0.2, [1.3])])
    fns = {("Fe", "Fe"): (1.01, lambda r: 12.0 * np.exp(-(np.linalg.norm(r) - 1.0) ** 2 / (2 * 0.3 ** 2))),
           ("Fe", "Co"): (0.9, lambda r: 20.0 * np.exp(-(np.linalg.norm(r) - 0.8) / 0.25)),
           ("Co", "Fe"): (0.9, lambda r: 20.0 * np.exp(-(np.linalg.norm(r) - 0.8) / 0.25)),
           ("Co", "Co"): (1.45, lambda r: -3.0 * ((2 * 1.3 * (np.linalg.norm(r) - 0.2)) * np.cos(2 * 1.3 * (np.linalg.norm(r) - 0.2)) - np.sin(2 * 1.3 * (np.linalg.norm(r) - 0.2)))
                                 / (2 * 1.3 * (np.linalg.norm(r) - 0.2)) ** 4)}
    return lat, settings, fns


def test_exchange_functional_template_equals_a_brute_force_pair_search():
    """hamiltonian/exchange_functional.cc:206-243: every ordered pair within its material pair's cutoff gets J(r_ij); the
    template route (what the kernels consume) must give the same list as a brute-force minimum-image search over all pairs"""
    from helpers import brute_force_functional_pairs
    from jams_b200.solver import create_hamiltonian
    lat, settings, fns = _functional_case()
    h = create_hamiltonian(settings, lat)
    i, j, v, vals = h.neighbour_list()
    bi, bj, bJ = brute_force_functional_pairs(lat, fns)
    got = {(int(a), int(b)): vals[c][0] for a, b, c in zip(i, j, v)}
    want = {(int(a), int(b)): J9[0] for a, b, J9 in zip(bi, bj, bJ)}
    assert len(i) == len(bi) == len(got) and set(got) == set(want)
    assert max(abs(got[k] - want[k]) for k in got) <= 1e-13 * max(abs(x) for x in want.values())
    # Fe-Fe: 6 neighbours at 1.0 (4 across the open y faces for boundary cells), Fe-Co: 8 at 0.866, Co-Co: 6 + 12 (1.0, 1.414)
    counts = np.bincount(i, minlength=lat.num_spins)
    inner = lat.site_index(2, 1, 3, 0), lat.site_index(2, 1, 3, 1)
    assert counts[inner[0]] == 6 + 8 and counts[inner[1]] == 8 + 6 + 12
    # reference error behaviour (exchange_functional.cc:13-88,118-160)
    for bad, msg in ((dict(settings, interactions=[("Fe", "Fe", "gaussian", 1.0, 1.0, 1.0)]), "expects 3 parameters"),
                     (dict(settings, interactions=[("Fe", "Fe", "sinc", 1.0, 1.0)]), "unknown exchange functional"),
                     (dict(settings, interactions=[("Fe", "Fe", "gaussian", 1.0, 1.0, 1.0, 0.0)]), "non-zero parameter 'sigma'"),
                     (dict(settings, interactions=[("Fe", "Ni", "step", 1.0, 1.0, 1.0)]), "does not exist"),
                     (dict(settings, interactions=[("Fe", "Fe", "step", 1.0, 1.0, 1.0), ("Fe", "Fe", "step", 1.0, 1.0, 1.0)]), "defined more than once"),
                     (dict(settings, interactions=[("Fe", "Fe", "step", 9.0, 1.0, 9.0)]), "larger than the maximum cutoff radius"),
                     (dict(settings, distance_units="furlongs"), "distance units")):
        with pytest.raises(RuntimeError, match=msg):
            create_hamiltonian(bad, lat)


def test_biquadratic_template_keeps_only_couplings_above_the_cutoff_and_matches_the_oracle_pairs():
    """cuda_biquadratic_exchange.cu:127-134: B_ij = unit * J[0][0], inserted only if it exceeds energy_cutoff * unit"""
    from helpers import oracle_exchange_pairs
    lat = Lattice([Material("Fe", 2.2)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (4, 4, 4))
    hs = dict(module="biquadratic-exchange", energy_units="meV", interaction_prefactor=7.0,   # the prefactor is not a setting of this module
              interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 0.5), ("Fe", "Fe", [1.0, 0.0, 0.0], 0.2), ("Fe", "Fe", [1.0, 1.0, 0.0], -0.1)])
    h = create_hamiltonian(hs, lat)
    t = h.template
    assert len(t["B"]) == 2 * (8 + 6) and set(np.round(t["B"], 12)) == {0.5, 0.2}
    i, j, J9, _ = oracle_exchange_pairs(lat, dict(hs, interaction_prefactor=1.0))
    keep = J9[:, 0] > 0.0
    # expand the template over the lattice and compare with the oracle's neighbour list as (i, j, B) sets
    nl = lat.neighbour_list(dict(mi=t["mi"], mj=t["mj"], T=t["T"], J9=np.repeat(t["B"][:, None], 9, axis=1)))
    got = sorted(zip(nl[0].tolist(), nl[1].tolist(), np.round(nl[3][nl[2]][:, 0], 12).tolist()))
    want = sorted(zip(i[keep].tolist(), j[keep].tolist(), np.round(J9[keep, 0], 12).tolist()))
    assert got == want and len(got) == lat.num_spins * 14
    hs_cut = dict(hs, energy_cutoff=0.3)
    assert len(create_hamiltonian(hs_cut, lat).template["B"]) == 2 * 8


def test_pcg32_known_answer_and_impurity_substitution():
    """lattice.impurities (core/lattice.cc:424-427,614-640,1077-1109): pcg32 restated from its published definition (the demo
    vector of pcg32(42, stream 54) is the known answer), one uniform draw per candidate site in site order"""
    from jams_b200.lattice import Pcg32

    class Demo(Pcg32):
        INC = (54 << 1) | 1
    r = Demo(42)
    assert [r() for _ in range(6)] == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    mats = [Material("A", 2.0), Material("B", 1.0, alpha=0.2), Material("C", 3.0)]
    motif = [("A", (0, 0, 0)), ("C", (0.5, 0.5, 0.5))]
    lat = Lattice(mats, np.eye(3), motif, (12, 10, 8), impurities=[("A", "B", 0.25)], impurities_seed=11)
    mat = lat.site_material()
    assert set(mat[1::2]) == {2}                                  # C sites are not candidates
    frac = (mat[0::2] == 1).mean()
    assert abs(frac - 0.25) < 0.04 and set(mat[0::2]) == {0, 1}
    assert np.array_equal(mat, Lattice(mats, np.eye(3), motif, (12, 10, 8), impurities=[("A", "B", 0.25)], impurities_seed=11).site_material())
    assert not np.array_equal(mat, Lattice(mats, np.eye(3), motif, (12, 10, 8), impurities=[("A", "B", 0.25)], impurities_seed=12).site_material())
    # the draws in site order: site q of the A sublattice is substituted iff the q-th uniform number is below the fraction
    rng = Pcg32(11)
    want = np.array([1 if rng.uniform_real() < 0.25 else 0 for _ in range(mat.size // 2)])
    assert np.array_equal(mat[0::2], want)
    from jams_b200.consts import kBohrMagnetonIU
    assert np.array_equal(lat.mus(), np.array([m.moment * kBohrMagnetonIU for m in mats])[mat])
    x0, nx = 3, 4
    per_plane = 10 * 8 * 2
    assert np.array_equal(lat.site_material(x0, nx), mat[x0 * per_plane:(x0 + nx) * per_plane])
    for bad, msg in (([("Z", "B", 0.1)], "materialA"), ([("A", "Z", 0.1)], "materialB"), ([("A", "B", 1.0)], "fraction"),
                     ([("A", "B", 0.1), ("A", "C", 0.1)], "redefines")):
        with pytest.raises(RuntimeError, match=msg):
            Lattice(mats, np.eye(3), motif, (4, 4, 4), impurities=bad)


def test_neighbour_list_with_impurities_bit_exact_with_oracle():
    """neighbour_list_from_interactions skips pairs whose site materials differ from the entry's types
    (core/interactions.cc:381-385): the product's list against the oracle's, given the same site materials"""
    mats = [Material("Fe", 2.2), Material("Co", 1.7)]
    lat = Lattice(mats, np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 4), periodic=(True, True, False),
                  impurities=[("Fe", "Co", 0.3)], impurities_seed=3)
    hs = dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21)])
    h = create_hamiltonian(hs, lat)
    assert h.use_pairs
    i, j, v, vals = h.neighbour_list()
    oi, oj, oJ9, _ = oracle_exchange_pairs(lat, hs)
    assert np.array_equal(i, oi) and np.array_equal(j, oj) and np.array_equal(vals[v], oJ9)
    mat = lat.site_material()
    assert (mat[i] == 0).all() and (mat[j] == 0).all() and 0 < len(i) < lat.num_spins * 14


def test_work_item_plan_covers_the_slab_exactly_once():
    """jb_plan_work_items (host only): for any slab the x-chunks tile [0, nx) without gaps or overlaps, no chunk is shorter than the
    ghost depth, the face chunks lead the queue, long chunks come before the taper; small lattices (BASELINE configs 1 and 2) get
    chunks short enough to give every resident CTA an item"""
    rng = np.random.default_rng(0)
    cases = [(256, 1, 128, 296), (256, 1, 256, 592), (128, 1, 64, 296), (128, 2, 32, 148), (64, 1, 16, 296), (64, 1, 512, 592), (256, 1, 4, 296),
             (16, 1, 4, 296), (8, 1, 2, 296), (5, 2, 3, 148), (1, 0, 7, 296), (3, 3, 1, 148), (512, 1, 512, 296), (2048, 1, 128, 296)]
    cases += [(int(rng.integers(1, 600)), int(rng.integers(0, 4)), int(rng.integers(1, 300)), int(rng.choice([148, 296, 592]))) for _ in range(200)]
    for nx, gx, n_cols, G in cases:
        if nx < max(gx, 1):
            continue
        plan = capi.plan_work_items(nx, gx, n_cols, G)
        assert 1 <= len(plan) <= 160, (nx, gx, n_cols, G)
        covered = np.zeros(nx, int)
        for x0, xc in plan:
            assert xc >= 1 and 0 <= x0 and x0 + xc <= nx, (nx, gx, n_cols, G, plan)
            covered[x0:x0 + xc] += 1
        assert (covered == 1).all(), (nx, gx, n_cols, G, plan)
        if len(plan) > 1:
            assert min(xc for _, xc in plan) >= max(gx, 1), (nx, gx, n_cols, G, plan)
        face = [(x0 < gx) or (x0 + xc > nx - gx) for x0, xc in plan]
        n_face = sum(face)
        assert all(face[:n_face]) and not any(face[n_face:]), (nx, gx, n_cols, G, plan)     # face chunks first
    # small lattices: every resident CTA gets an item when there is enough work to go round
    for nx, gx, n_cols, G in ((64, 1, 16, 296), (256, 1, 4, 296), (32, 1, 16, 296)):
        plan = capi.plan_work_items(nx, gx, n_cols, G)
        assert len(plan) * n_cols >= min(G, nx * n_cols) * 0.8, (nx, n_cols, G, len(plan))
    # the bench workload: long chunks first, then a taper that ends in short chunks
    plan = capi.plan_work_items(256, 1, 256, 592)
    inner = [xc for x0, xc in plan if not ((x0 < 1) or (x0 + xc > 255))]
    long = [xc for xc in inner if xc >= 16]
    assert inner[:len(long)] == long and max(long) - min(long) <= 1           # equal long chunks lead ...
    assert inner[len(long):] == sorted(inner[len(long):], reverse=True) and inner[-1] <= 4   # ... the taper ends in short ones

