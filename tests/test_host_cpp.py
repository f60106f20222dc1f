"""The C++ host layer (jams_b200/host/: libconfig-grammar front end, Lattice, Hamiltonians, Solver, Monitors, `jams-b200`)
against the Python mirror that the oracle parity tests validated: same config file -> same lattice arrays, the same
exchange template bit for bit, and on the GPU the same trajectory and JAMS-format monitor files."""
import os
import subprocess

import numpy as np
import pytest

from jams_b200 import host, workloads as W
from jams_b200.lattice import Lattice, Material
from jams_b200.solver import create_hamiltonian

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "bloch_wall_small.cfg")
PATCH_B200 = 'solver : { module = "llg-heun-b200-gpu"; };'


def _sorted_template(t):
    key = np.lexsort((t["T"][:, 2], t["T"][:, 1], t["T"][:, 0], t["mj"], t["mi"]))
    return t["mi"][key], t["mj"][key], t["T"][key], t["J9"][key]


def test_config_grammar_and_types():
    d = host.config_to_dict("""
        # hash comment
        a = 1; b : 2.5, c = "str" "ing";   /* block
        comment */ d = true; e = FALSE; f = 0x1F; g = 12345678901L; h = 1e-23; i = -3;
        grp : { x = [1, 2, 3]; y = ( "A", [0.5, 1], { k = 1; } ); empty = (); };
        list = ( (1, 2.0, "three"), [1.0, 2] )
    """)
    assert d["a"] == 1 and d["b"] == 2.5 and d["c"] == "string" and d["d"] is True and d["e"] is False
    assert d["f"] == 31 and d["g"] == 12345678901 and d["h"] == 1e-23 and d["i"] == -3
    assert d["grp"] == {"x": [1, 2, 3], "y": ["A", [0.5, 1], {"k": 1}], "empty": []}
    assert d["list"] == [[1, 2.0, "three"], [1.0, 2]]
    for bad in ("a = ;", "a = 1; a = 2;", "x = [1, \"s\"];", "g = { a = 1;", 's = "unterminated'):
        with pytest.raises(host.HostError):
            host.config_to_dict(bad)


def test_config_patches_merge_left_to_right_like_jams():
    """core/jams++.cc:48-84 + interface/config.cc:12-145: scalars are replaced, missing settings added, lists patched by position"""
    d = host.config_to_dict(FIXTURE, 'solver : { module = "llg-heun-b200-gpu"; t_max = 2e-15; }; physics : { temperature = 30.0; };',
                            'lattice : { size = [8, 8, 8]; }; hamiltonians = ( { order = "K2"; } ); sim : { seed = 7; };')
    assert d["solver"] == {"t_step": 1e-16, "module": "llg-heun-b200-gpu", "t_max": 2e-15}
    assert d["physics"]["temperature"] == 30.0 and d["sim"]["seed"] == 7
    assert d["lattice"] == {"size": [8, 8, 8], "periodic": [False, True, True]}
    assert d["hamiltonians"][0]["order"] == "K2" and d["hamiltonians"][0]["module"] == "uniaxial" and d["hamiltonians"][1]["module"] == "exchange"
    assert d["materials"][0]["name"] == "A"


def test_cpp_lattice_and_template_equal_the_python_mirror_bit_for_bit():
    w = W.c1_bloch_wall((32, 4, 4))
    lat = w["lattice"]
    la = host.lattice_arrays(FIXTURE)
    assert la["num_spins"] == lat.num_spins and la["M"] == lat.M and la["dims"] == lat.dims and la["periodic"] == lat.periodic
    assert np.array_equal(la["mus"], lat.mus()) and np.array_equal(la["gyro"], lat.gyro()) and np.array_equal(la["alpha"], lat.alpha())
    assert np.array_equal(la["spins"], lat.initial_spins()) and np.array_equal(la["positions"], lat.positions())
    h = create_hamiltonian(w["hamiltonians"][1], lat)
    t = host.exchange_template(FIXTURE, ham_index=1)
    for x, y in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    assert t["n_pairs"] == len(h.neighbour_list()[0])


BCC_CFG = """
materials = ( { name = "Fe"; moment = 2.2; alpha = 0.1; }, { name = "Co"; moment = 1.7; alpha = 0.05; gyro = 1.1; spin = [90.0, 0.0]; } );
unitcell : { parameter = 2.87e-10; basis = ([1.0,0,0],[0,1.0,0],[0,0,1.0]); positions = ( ("Fe", [0,0,0]), ("Co", [0.5,0.5,0.5]) ); };
lattice : { size = [6, 5, 4]; periodic = [true, true, false]; };
hamiltonians = ( { module = "exchange"; energy_units = "meV"; energy_cutoff = 0.5;
                   interactions = ( ("Fe","Co",[0.5,0.5,0.5], [20.0,0,0, 0,20.0,0, 0,0,20.0]), ("Co","Fe",[0.5,0.5,0.5], [20.0,0,0, 0,20.0,0, 0,0,20.0]),
                                    ("Fe","Fe",[1.0,0,0], [10.0,0,0, 0,10.0,0, 0,0,10.0]), ("Co","Co",[1.0,0,0], [8.0,0.1,0, -0.1,8.0,0, 0,0,7.5]),
                                    ("Fe","Fe",[1.0,1.0,0], [0.4,0,0, 0,0.4,0, 0,0,0.4]) ); } );
solver : { module = "llg-heun-b200-gpu"; t_step = 1e-16; t_max = 1e-15; gilbert_prefactor = true; };
"""


def test_two_material_bcc_template_with_cutoffs_and_tensors():
    from jams_b200.lattice import Lattice, Material
    lat = Lattice([Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, gyro=1.1, alpha=0.05, spin=(1.0, 0.0, 0.0))], np.eye(3),
                  [("Fe", (0, 0, 0)), ("Co", (0.5, 0.5, 0.5))], (6, 5, 4), periodic=(True, True, False), gilbert_prefactor=True)
    inter = [("Fe", "Co", [0.5, 0.5, 0.5], 20.0), ("Co", "Fe", [0.5, 0.5, 0.5], 20.0), ("Fe", "Fe", [1.0, 0, 0], 10.0),
             ("Co", "Co", [1.0, 0, 0], [8.0, 0.1, 0, -0.1, 8.0, 0, 0, 0, 7.5]), ("Fe", "Fe", [1.0, 1.0, 0], 0.4)]
    h = create_hamiltonian(dict(module="exchange", energy_units="meV", energy_cutoff=0.5, interactions=inter), lat)
    t = host.exchange_template(BCC_CFG, ham_index=0)
    for x, y in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    assert t["n_pairs"] == len(h.neighbour_list()[0])
    la = host.lattice_arrays(BCC_CFG)
    assert np.array_equal(la["mus"], lat.mus()) and np.array_equal(la["gyro"], lat.gyro()) and np.array_equal(la["alpha"], lat.alpha())
    assert np.allclose(la["spins"], lat.initial_spins(), rtol=0, atol=1e-16)   # spherical-angle spin setting
    assert np.array_equal(la["positions"], lat.positions())


def test_lattice_impurities_cpp_equals_python(tmp_path):
    """lattice.impurities (core/lattice.cc:424-427,614-640): both host layers restate pcg32 + generate_canonical, so the same seed
    substitutes the same sites; the neighbour list drops the pairs that touch a substituted site (core/interactions.cc:381-385)"""
    cfg = BCC_CFG.replace('("Co", [0.5,0.5,0.5])', '("Fe", [0.5,0.5,0.5])').replace(
        'lattice : { size = [6, 5, 4];', 'lattice : { impurities = ( ("Fe", "Co", 0.35) ); impurities_seed = 21; size = [6, 5, 4];')
    path = tmp_path / "alloy.cfg"
    open(path, "w").write(cfg)
    la = host.lattice_arrays(str(path))
    mats = [Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05, gyro=1.1, spin=(1.0, 0.0, 0.0))]
    lat = Lattice(mats, np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 4), periodic=(True, True, False), gilbert_prefactor=True,
                  impurities=[("Fe", "Co", 0.35)], impurities_seed=21)
    assert 0.25 < (lat.site_material() == 1).mean() < 0.45
    assert np.array_equal(la["mus"], lat.mus()) and np.array_equal(la["alpha"], lat.alpha()) and np.array_equal(la["gyro"], lat.gyro())
    assert np.abs(la["spins"] - lat.initial_spins()).max() <= 1e-15
    hs = dict(module="exchange", energy_units="meV", energy_cutoff=0.5,
              interactions=[("Fe", "Co", [0.5, 0.5, 0.5], [20.0, 0, 0, 0, 20.0, 0, 0, 0, 20.0]), ("Co", "Fe", [0.5, 0.5, 0.5], [20.0, 0, 0, 0, 20.0, 0, 0, 0, 20.0]),
                            ("Fe", "Fe", [1.0, 0, 0], [10.0, 0, 0, 0, 10.0, 0, 0, 0, 10.0]), ("Co", "Co", [1.0, 0, 0], [8.0, 0.1, 0, -0.1, 8.0, 0, 0, 0, 7.5]),
                            ("Fe", "Fe", [1.0, 1.0, 0], [0.4, 0, 0, 0, 0.4, 0, 0, 0, 0.4])])
    h = create_hamiltonian(hs, lat)
    t = host.exchange_template(str(path), ham_index=0)
    for x, y in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    n_py = len(h.neighbour_list()[0])
    assert t["n_pairs"] == n_py and 0 < n_py < lat.num_spins * 6


def test_lattice_rotations_cpp_equals_python(tmp_path):
    """lattice.orientation_axis / orientation_lattice_vector and lattice.global_rotation (core/lattice.cc:434-454,515-575): the unit
    cell vectors are rotated (orientation first), positions follow; interaction vectors are Cartesian in the ROTATED frame, so the bcc
    template of a lattice whose [111] was turned onto z and then rotated by 90 degrees about z is found with the rotated vectors"""
    from jams_b200.lattice import rotation_matrix_between_vectors
    Rz = [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
    R1 = rotation_matrix_between_vectors(np.array([1.0, 1.0, 1.0]) / np.sqrt(3.0), np.array([0.0, 0.0, 1.0]))
    R = np.array(Rz) @ R1
    nn = R @ np.array([0.5, 0.5, 0.5])
    cfg = BCC_CFG.replace('("Co", [0.5,0.5,0.5])', '("Fe", [0.5,0.5,0.5])').replace(
        'lattice : { size = [6, 5, 4];',
        'lattice : { orientation_axis = [0.0, 0.0, 1.0]; orientation_lattice_vector = [1.0, 1.0, 1.0]; '
        'global_rotation = ([0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]); size = [6, 5, 4];')
    cfg = cfg[:cfg.index("hamiltonians")] + ('hamiltonians = ( { module = "exchange"; interactions = ( ("Fe", "Fe", [%.17g, %.17g, %.17g], 3.2e-21) ); } );\n' % tuple(nn)) + cfg[cfg.index("solver :"):]
    path = tmp_path / "rot.cfg"
    open(path, "w").write(cfg)
    mats = [Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05, gyro=1.1, spin=(1.0, 0.0, 0.0))]
    lat = Lattice(mats, np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 4), periodic=(True, True, False), gilbert_prefactor=True,
                  orientation_axis=(0.0, 0.0, 1.0), orientation_lattice_vector=(1.0, 1.0, 1.0), global_rotation=Rz)
    assert np.abs(lat.cell - R).max() <= 1e-15
    la = host.lattice_arrays(str(path))
    assert np.abs(la["positions"] - lat.positions()).max() <= 1e-14
    h = create_hamiltonian(dict(module="exchange", interactions=[("Fe", "Fe", list(nn), 3.2e-21)]), lat)
    assert len(h.template["mi"]) == 16          # 8 nearest neighbours per motif site, as on the unrotated lattice
    t = host.exchange_template(str(path), ham_index=0)
    for x, y in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(x, y)
    plain = Lattice(mats, np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 4), periodic=(True, True, False))
    h0 = create_hamiltonian(dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21)]), plain)
    for x, y in zip(_sorted_template(h.template), _sorted_template(h0.template)):
        assert np.array_equal(x, y)


def test_known_neighbour_count_sc_8_cubed():
    """sc 8^3 NN with symmetry operations -> 8*8*8*6 interactions (reference src/jams/test/interactions.h:241-252)"""
    t = host.exchange_template(FIXTURE, "lattice : { size = [8, 8, 8]; periodic = [true, true, true]; };", ham_index=1)
    assert len(t["mi"]) == 6 and t["n_pairs"] == 8 * 8 * 8 * 6
    assert np.allclose(t["J9"][:, [0, 4, 8]], 3.5e-21 * 6.24150907e21) and np.all(t["J9"][:, [1, 2, 3, 5, 6, 7]] == 0)


def test_error_behaviour_mirrors_the_reference():
    with pytest.raises(host.HostError, match="unknown solver"):
        host.run(FIXTURE, 'solver : { module = "monte-carlo-metropolis-cpu"; };', num_spins=512)
    with pytest.raises(host.HostError, match="energy units"):
        host.exchange_template(FIXTURE, 'hamiltonians = ( {}, { energy_units = "furlongs"; } );', ham_index=1)
    with pytest.raises(host.HostError, match="Multiple interactions"):   # core/interactions.cc:373-381
        host.exchange_template(FIXTURE, "lattice : { size = [32, 1, 4]; };", ham_index=1)
    with pytest.raises(host.HostError, match="required setting 'moment'"):
        host.lattice_arrays('materials = ( { name = "A"; } ); unitcell : { parameter = 1e-10; basis = ([1.0,0,0],[0,1.0,0],[0,0,1.0]); positions = (("A",[0,0,0])); }; lattice : { size = [1,1,1]; };')


@pytest.mark.gpu
def test_cpp_several_uniaxial_hamiltonians_and_the_one_of_a_kind_rule(tmp_path):
    """the reference sums any number of Hamiltonians (core/solver.cc:43-57): two "uniaxial" modules (K1 + K2) run in two slots and
    give the Python mirror's trajectory bit for bit; a second exchange module is refused instead of replacing the first"""
    text = open(FIXTURE).read()
    assert text.count('{\n    module = "exchange";') == 1
    k2 = '{ module = "uniaxial"; order = "K2"; anisotropies = ( ( "A", [ 0.0, 0.6, 0.8 ], 3e-23 ) ); },\n  {\n    module = "exchange";'
    cfg = tmp_path / "two_uniaxial.cfg"
    cfg.write_text(text.replace('{\n    module = "exchange";', k2))
    got, done = host.run(str(cfg), PATCH_B200, name="k1k2", output_dir=str(tmp_path))
    w = W.c1_bloch_wall((32, 4, 4))
    w["hamiltonians"].insert(1, dict(module="uniaxial", order="K2", anisotropies=[("A", [0.0, 0.6, 0.8], 3e-23)]))
    assert [h["module"] for h in w["hamiltonians"]] == ["uniaxial", "uniaxial", "exchange"]
    init, _ = host.run(str(cfg), PATCH_B200, name="init", output_dir=str(tmp_path), max_steps=0)
    s = W.make_solver(w)
    s.set_spins(init)
    s.run(done)
    assert s.ctx.stage_kernel() == 0 and np.array_equal(got, s.spins())
    eng = open(tmp_path / "k1k2_eng.tsv").read().splitlines()
    assert eng[0].split() == ["time", "uniaxial_E_meV", "uniaxial_E_meV", "exchange_E_meV"]
    two_exchange = tmp_path / "two_exchange.cfg"
    two_exchange.write_text(text.replace('{\n    module = "exchange";', '{ module = "exchange"; interactions = (("A", "A", [0.0, 1.0, 0.0], 1e-21)); },\n  {\n    module = "exchange";'))
    with pytest.raises(host.HostError, match="same kind of term"):
        host.run(str(two_exchange), PATCH_B200, name="dup", output_dir=str(tmp_path))


@pytest.mark.gpu
def test_cpp_simulation_equals_python_mirror_and_writes_jams_monitor_files(tmp_path):
    """the same config through the C++ Simulation (C++ lattice/template/initializer/main loop) and through the Python mirror"""
    steps = 40
    got, done = host.run(FIXTURE, PATCH_B200, name="wall", output_dir=str(tmp_path))
    assert done == steps
    w = W.c1_bloch_wall((32, 4, 4))
    from jams_b200.lattice import bloch_domain_wall
    lat = w["lattice"]
    w["spins"] = bloch_domain_wall(lat.positions(), lat.initial_spins(), width=8.0, center=16.0)
    # the initializer's rotation matrices are built by different code (C++ here, numpy there): agreement to rounding
    init, zero = host.run(FIXTURE, PATCH_B200, name="init", output_dir=str(tmp_path), max_steps=0)
    assert zero == 0 and np.abs(init - w["spins"]).max() <= 1e-15
    s = W.make_solver(w)
    s.set_spins(init)
    s.run(steps)
    assert np.array_equal(got, s.spins())
    mag = open(tmp_path / "wall_mag.tsv").read().splitlines()
    assert mag[0].split() == ["time", "T", "hx", "hy", "hz", "A_mx", "A_my", "A_mz", "A_m"]
    assert len(mag) == 1 + 4 and all(len(line) == 16 * 9 for line in mag[1:])     # jams::fmt::sci columns, steps 0,10,20,30
    assert float(mag[2].split()[0]) == pytest.approx(10 * 1e-4)
    eng = open(tmp_path / "wall_eng.tsv").read().splitlines()
    assert eng[0].split() == ["time", "uniaxial_E_meV", "exchange_E_meV"] and len(eng) == 1 + 2
    # the command-line driver runs the same thing
    r = subprocess.run([host.EXE_PATH, "--name", "cli", "--output", str(tmp_path), FIXTURE, PATCH_B200], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "cli_mag.tsv").read() == open(tmp_path / "wall_mag.tsv").read()


@pytest.mark.gpu
def test_cpp_simulation_with_biquadratic_exchange_equals_python_mirror(tmp_path):
    """module = "biquadratic-exchange" (hamiltonian/cuda_biquadratic_exchange.cu) through the C++ host: the patch replaces the
    Hamiltonian list; a negative coupling must be dropped by the value > energy_cutoff filter; the energy monitor writes the
    reference's total (half the sum of the per-spin energies)"""
    from jams_b200.solver import create_hamiltonian, create_solver
    patch = ('hamiltonians = ( { module = "exchange"; interactions = (("A", "A", [1.0, 0.0, 0.0], 3.5e-21)); }, '
             '{ module = "biquadratic-exchange"; interactions = (("A", "A", [1.0, 0.0, 0.0], 0.7e-21), ("A", "A", [1.0, 1.0, 0.0], -0.2e-21)); } ); ')
    got, done = host.run(FIXTURE, PATCH_B200, patch, name="bq", output_dir=str(tmp_path))
    assert done == 40
    init, _ = host.run(FIXTURE, PATCH_B200, patch, name="bq0", output_dir=str(tmp_path), max_steps=0)
    lat = W.c1_bloch_wall((32, 4, 4))["lattice"]
    s = create_solver(dict(module="llg-heun-b200-gpu", t_step=1e-16, t_max=1e-9, seed=0), lat)
    hs = [dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]),
          dict(module="biquadratic-exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 0.7e-21), ("A", "A", [1.0, 1.0, 0.0], -0.2e-21)])]
    for h in hs:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    assert len(s.hamiltonians[1].template["B"]) == 6
    s.set_spins(init)
    e0 = s.hamiltonians[1].calculate_total_energy(0.0)
    s.run(40)
    assert np.array_equal(got, s.spins())
    eng = open(tmp_path / "bq_eng.tsv").read().splitlines()
    assert eng[0].split() == ["time", "exchange_E_meV", "biquadratic-exchange_E_meV"]
    assert float(eng[1].split()[2]) == pytest.approx(e0, rel=1e-12)


@pytest.mark.gpu
def test_cpp_simulation_runs_the_shipped_example_setup_rk4_with_pinned_boundaries(tmp_path):
    """examples/bloch_domain_wall as shipped: solver llg-rk4-gpu + physics pinned_boundaries (bloch_domain_wall.cfg:75-81,132-139),
    through the C++ Simulation and through the Python mirror (main loop order of core/jams++.cc:333-341)"""
    from jams_b200.solver import create_solver, create_hamiltonian, create_physics
    from jams_b200.lattice import bloch_domain_wall
    patch = ('solver : { module = "llg-rk4-b200-gpu"; t_step = 5e-16; t_max = 1.5e-14; }; '
             'physics : { module = "pinned_boundaries"; temperature = 0.0; left_pinned_magnetisation = [0.0, 0.0, -1.0]; '
             'right_pinned_magnetisation = [0.0, 0.0, 1.0]; left_pinned_cells = 2; right_pinned_cells = 3; };')
    got, done = host.run(FIXTURE, patch, name="pinned", output_dir=str(tmp_path))
    assert done == 30
    w = W.c1_bloch_wall((32, 4, 4))
    lat = w["lattice"]
    init = bloch_domain_wall(lat.positions(), lat.initial_spins(), width=8.0, center=16.0)
    s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=5e-16, t_max=1.5e-14), lat)
    for h in w["hamiltonians"]:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.register_physics_module(create_physics(dict(module="pinned_boundaries", left_pinned_magnetisation=[0.0, 0.0, -1.0],
                                                  right_pinned_magnetisation=[0.0, 0.0, 1.0], left_pinned_cells=2, right_pinned_cells=3), lat))
    s.set_spins(init)
    while s.is_running():
        s.update_physics_module()
        s.run(1)
    assert s.iteration == 30
    assert np.abs(got - s.spins()).max() <= 1e-12
    # the pinned regions point where they were told to
    m = got.reshape(32, 4, 4, 3)
    assert m[:2, ..., 2].mean() < -0.9 and m[-3:, ..., 2].mean() > 0.9


def test_lattice_spins_file_is_read_like_the_reference_loader(tmp_path):
    """lattice.spins = "file" (core/lattice.cc:738-748) through the text route of load_array_from_file (helpers/load.h:21-61):
    whitespace separated, empty lines and '#' / '//' comment lines skipped, the element count must match"""
    base = host.lattice_arrays(FIXTURE)
    N = base["num_spins"]
    rng = np.random.default_rng(5)
    s = rng.standard_normal((N, 3)); s /= np.linalg.norm(s, axis=1, keepdims=True)
    f = tmp_path / "state.tsv"
    with open(f, "w") as fh:
        fh.write("# spins\n\n// another comment\n")
        for k, row in enumerate(s):
            fh.write(("%.17g\t%.17g %.17g\n" % tuple(row)) if k % 2 else ("  %.17g %.17g %.17g  \n" % tuple(row)))
    got = host.lattice_arrays(FIXTURE, 'lattice : { spins = "%s"; };' % f)
    assert np.array_equal(got["spins"], s)
    with open(f, "a") as fh:
        fh.write("1.0 0.0\n")
    with pytest.raises(host.HostError, match="expected size"):
        host.lattice_arrays(FIXTURE, 'lattice : { spins = "%s"; };' % f)
    with pytest.raises(host.HostError, match="failed to open file"):
        host.lattice_arrays(FIXTURE, 'lattice : { spins = "%s"; };' % (tmp_path / "missing.tsv"))
    with pytest.raises(host.HostError, match="HDF5 is not available"):
        host.lattice_arrays(FIXTURE, 'lattice : { spins = "state.h5"; };')


@pytest.mark.gpu
def test_checkpoint_and_restart_through_the_spin_snapshot_monitor(tmp_path):
    """the reference's checkpoint / resume idiom (monitors/hdf5.cc + lattice.spins): a run that stops after 20 steps,
    writes <name>_final, and is resumed from that file reproduces the uninterrupted 40-step run bit for bit (T = 0)"""
    import re
    cfg = tmp_path / "noinit.cfg"   # the fixture without its initializer group (a resumed run must not re-initialise) and random initial spins
    text = re.sub(r"initializer\s*:\s*\{.*?\};", "", open(FIXTURE).read(), flags=re.S)
    open(cfg, "w").write(text.replace("spin      = [0.0, 0.0, 1.0];", 'spin      = "random";') + "\nsim : { seed = 7; };\n")
    cfg = str(cfg)
    mon = 'monitors = ( { module = "hdf5"; output_steps = 10; } ); '
    full, n = host.run(cfg, PATCH_B200, mon, name="full", output_dir=str(tmp_path))
    assert n == 40 and np.abs(full - host.lattice_arrays(cfg)["spins"]).max() > 1e-6     # something happened
    half, n = host.run(cfg, PATCH_B200, mon, 'solver : { t_max = 2e-15; };', name="half", output_dir=str(tmp_path))
    assert n == 20
    for k in (0, 10):
        assert (tmp_path / ("half_%07d.tsv" % k)).exists()
    final = np.loadtxt(tmp_path / "half_final.tsv")
    assert np.array_equal(final, half)
    rest, n = host.run(cfg, PATCH_B200, 'solver : { t_max = 2e-15; }; lattice : { spins = "%s"; };' % (tmp_path / "half_final.tsv"),
                       name="rest", output_dir=str(tmp_path))
    assert n == 20
    assert np.array_equal(rest, full)
    # T > 0: the snapshot header carries the iteration, the resumed run continues the Philox noise stream (step index = counter) instead
    # of replaying the first segment's draws -- so it, too, equals the uninterrupted run bit for bit
    hot = 'physics : { temperature = 50.0; }; '
    full, _ = host.run(cfg, PATCH_B200, mon, hot, name="fullT", output_dir=str(tmp_path))
    half, _ = host.run(cfg, PATCH_B200, mon, hot, 'solver : { t_max = 2e-15; };', name="halfT", output_dir=str(tmp_path))
    rest, n = host.run(cfg, PATCH_B200, hot, 'solver : { t_max = 2e-15; }; lattice : { spins = "%s"; };' % (tmp_path / "halfT_final.tsv"),
                       name="restT", output_dir=str(tmp_path))
    assert n == 20 and np.abs(half - full).max() > 1e-9
    assert np.array_equal(rest, full)


@pytest.mark.gpu
def test_cpp_simulation_with_temperature_ramp_physics_equals_python_mirror(tmp_path):
    """physics.module = "field-cool" and "two-temperature-model" (physics/field_cool.cc, two_temperature_model.cc): the C++
    Simulation hands physics()->temperature() to jb_step every iteration (core/solver.cc:94-97); the magnetisation monitor's T
    column (monitors/magnetisation.cc:83) shows the ramp, and the trajectory equals the Python mirror's main loop
    (same Philox noise: seed and step index are the same)"""
    from jams_b200.solver import create_solver, create_hamiltonian, create_physics
    from jams_b200.lattice import bloch_domain_wall
    cases = {
        "cool": (dict(module="field-cool", InitialTemperature=40.0, FinalTemperature=10.0, InitialField=[0.0, 0.0, 0.5], FinalField=[0.0, 0.0, 0.0], CoolTime=0.003),
                 'physics : { module = "field-cool"; InitialTemperature = 40.0; FinalTemperature = 10.0; InitialField = [0.0, 0.0, 0.5]; '
                 'FinalField = [0.0, 0.0, 0.0]; CoolTime = 0.003; };'),
        "ttm": (dict(module="two-temperature-model", InitialTemperature=20.0, ReversingField=[0.0, 0.0, -0.1], Ce=700.0, Cl=3.0e6, Gep=1.7e6, Gps=1.7e3,
                     laserPulses=[dict(width=0.0005, fluence=4.0e-9, t_start=0.0002)]),
                'physics : { module = "two-temperature-model"; InitialTemperature = 20.0; ReversingField = [0.0, 0.0, -0.1]; Ce = 700.0; Cl = 3.0e6; '
                'Gep = 1.7e6; Gps = 1.7e3; laserPulses = ( { width = 0.0005; fluence = 4.0e-9; t_start = 0.0002; } ); };'),
    }
    w = W.c1_bloch_wall((32, 4, 4))
    lat = w["lattice"]
    # the C++ initializer's spins (they agree with bloch_domain_wall() to rounding only, see the test above)
    init, _ = host.run(FIXTURE, PATCH_B200, name="init", output_dir=str(tmp_path), max_steps=0)
    assert np.abs(init - bloch_domain_wall(lat.positions(), lat.initial_spins(), width=8.0, center=16.0)).max() <= 1e-15
    for name, (py_cfg, patch) in cases.items():
        got, done = host.run(FIXTURE, PATCH_B200, patch, 'sim : { seed = 11; t_step = 1e-4; }; monitors = ( { module = "magnetisation"; output_steps = 5; } );',
                             name=name, output_dir=str(tmp_path))
        assert done == 40
        s = create_solver(dict(module="llg-heun-b200-gpu", t_step=1e-16, t_max=4e-15, seed=11), lat)
        for h in w["hamiltonians"]:
            s.register_hamiltonian(create_hamiltonian(h, lat))
        s.register_physics_module(create_physics(py_cfg, lat, dict(t_step=1e-4)))   # sim.t_step as written: field_cool.cc uses it next to times in ps
        s.set_spins(init)
        temps = []
        while s.is_running():
            s.update_physics_module()
            temps.append(s.temperature)
            s.run(1)
        assert len(set(temps)) > 5 and max(temps) - min(temps) > 1.0   # the temperature really moves
        assert np.abs(got - s.spins()).max() <= 1e-12, name            # std::exp / np.exp may differ in the last bit of T
        rows = [line.split() for line in open(tmp_path / (name + "_mag.tsv")).read().splitlines()[1:]]
        assert len(rows) == 8
        for k, r in enumerate(rows):                                  # monitor steps 0, 5, ..., 35: T and the reported applied field
            assert float(r[1]) == pytest.approx(temps[5 * k], rel=1e-7)
    assert (tmp_path / "ttm_ttm.tsv").exists() and len(open(tmp_path / "ttm_ttm.tsv").read().splitlines()) >= 2


def test_exc_file_gives_the_same_template_as_inline_interactions(tmp_path):
    """exc_file (hamiltonian/exchange.cc:118-133; core/interactions.cc:126-171,205-252) in the C++ host and the Python mirror:
    JAMS format (material names) with scalar J, KKR format (1-based motif indices) with a 3x3 tensor, comments and blank
    lines, and the reference's error behaviour"""
    from jams_b200.lattice import read_interaction_file
    f = tmp_path / "wall.exc"
    f.write_text("# J between nearest neighbours\n\n// (joules)\nA A 1.0 0.0 0.0 3.5e-21\n")
    inline = host.exchange_template(FIXTURE, ham_index=1)
    from_file = host.exchange_template(FIXTURE, 'hamiltonians = ( {}, { exc_file = "%s"; } );' % f, ham_index=1)
    for x, y in zip(_sorted_template(from_file), _sorted_template(inline)):
        assert np.array_equal(x, y)
    assert read_interaction_file(str(f)) == [("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]
    w = W.c1_bloch_wall((32, 4, 4))
    hs = dict(w["hamiltonians"][1]); hs.pop("interactions"); hs["exc_file"] = str(f)
    py = create_hamiltonian(hs, w["lattice"]).template
    for x, y in zip(_sorted_template(py), _sorted_template(inline)):
        assert np.array_equal(x, y)
    # KKR + tensor: 14 columns
    t9 = [1e-21, 2e-22, 0.0, -2e-22, 1e-21, 0.0, 0.0, 0.0, 1.5e-21]
    g = tmp_path / "kkr.exc"
    g.write_text("1 1  1.0 0.0 0.0  " + " ".join(repr(v) for v in t9) + "\n")
    # (a config of its own: a patch cannot turn the fixture's scalar entry into an array in place)
    text = open(FIXTURE).read()
    line = 'interactions = (("A", "A", [ 1.0, 0.0, 0.0], 3.5e-21));'
    assert line in text
    extra = " symops = false; check_sparse_matrix_symmetry = false;"
    cfg_file, cfg_inline = tmp_path / "tens_file.cfg", tmp_path / "tens_inline.cfg"
    cfg_file.write_text(text.replace(line, 'exc_file = "%s";' % g + extra))
    cfg_inline.write_text(text.replace(line, "interactions = ( (1, 1, [1.0, 0.0, 0.0], [%s]) );" % ", ".join(repr(v) for v in t9) + extra))
    tens_file = host.exchange_template(str(cfg_file), ham_index=1)
    tens_inline = host.exchange_template(str(cfg_inline), ham_index=1)
    assert len(tens_file["mi"]) == 1 and np.array_equal(tens_file["J9"][0], np.array(t9) * 6.24150907e21)
    for x, y in zip(_sorted_template(tens_file), _sorted_template(tens_inline)):
        assert np.array_equal(x, y)
    assert read_interaction_file(str(g)) == [(1, 1, [1.0, 0.0, 0.0], t9)]
    # errors (core/interactions.cc:143,168,240; exchange.cc:122-124)
    bad = tmp_path / "bad.exc"
    bad.write_text("A A 1.0 0.0 0.0\n")
    for reader in (lambda p: host.exchange_template(FIXTURE, 'hamiltonians = ( {}, { exc_file = "%s"; } );' % p, ham_index=1), lambda p: read_interaction_file(str(p))):
        with pytest.raises((host.HostError, RuntimeError), match="incorrect number of columns"):
            reader(bad)
        with pytest.raises((host.HostError, RuntimeError), match="failed to open file"):
            reader(tmp_path / "missing.exc")
    bad.write_text("# only comments\n\n")
    with pytest.raises(host.HostError, match="failed to discover interaction file format"):
        host.exchange_template(FIXTURE, 'hamiltonians = ( {}, { exc_file = "%s"; } );' % bad, ham_index=1)
    bad.write_text("A A 1.0 0.0 0.0 3.5e-21\nA A 0.0 1.0 zero 3.5e-21\n")
    with pytest.raises(host.HostError, match="failed to read line"):
        host.exchange_template(FIXTURE, 'hamiltonians = ( {}, { exc_file = "%s"; } );' % bad, ham_index=1)
    with pytest.raises(RuntimeError, match="failed to read line"):
        read_interaction_file(str(bad))


def test_exchange_functional_template_cpp_equals_python(tmp_path):
    """module = "exchange-functional" (hamiltonian/exchange_functional.cc) in the C++ host against the Python mirror (which
    tests/test_host_logic.py checks against a brute-force pair search): same template entries, couplings to rounding"""
    cfg = tmp_path / "functional.cfg"
    cfg.write_text("""
materials = ( { name = "Fe"; moment = 2.2; alpha = 0.1; spin = [0.0, 0.0, 1.0]; }, { name = "Co"; moment = 1.7; alpha = 0.05; spin = [0.0, 0.0, 1.0]; } );
unitcell : { parameter = 0.2866e-9; basis = ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]); positions = (("Fe", [0.0, 0.0, 0.0]), ("Co", [0.5, 0.5, 0.5])); };
lattice : { size = [5, 4, 6]; periodic = [true, false, true]; };
hamiltonians = ( { module = "exchange-functional"; energy_units = "meV"; distance_units = "angstroms";
    interactions = ( ("Fe", "Fe", "gaussian", 2.894, 12.0, 2.866, 0.86), ("Fe", "Co", "exponential", 2.58, 20.0, 2.29, 0.72),
                     ("Co", "Fe", "exponential", 2.58, 20.0, 2.29, 0.72), ("Co", "Co", "rkky", 4.16, 3.0, 0.57, [1.3]) ); } );
solver : { module = "llg-heun-b200-gpu"; t_step = 1e-16; t_max = 1e-15; };
""")
    from jams_b200.lattice import Lattice, Material
    lat = Lattice([Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05)], np.eye(3), [("Fe", (0, 0, 0)), ("Co", (0.5, 0.5, 0.5))], (5, 4, 6),
                  periodic=(True, False, True))
    a = 0.2866e-9
    py = create_hamiltonian(dict(module="exchange-functional", energy_units="meV", distance_units="angstroms", lattice_parameter=a,
                                 interactions=[("Fe", "Fe", "gaussian", 2.894, 12.0, 2.866, 0.86), ("Fe", "Co", "exponential", 2.58, 20.0, 2.29, 0.72),
                                               ("Co", "Fe", "exponential", 2.58, 20.0, 2.29, 0.72), ("Co", "Co", "rkky", 4.16, 3.0, 0.57, [1.3])]), lat).template
    cpp = host.exchange_template(str(cfg), ham_index=0)
    a_, b_ = _sorted_template(cpp), _sorted_template(py)
    assert len(a_[0]) == len(b_[0]) == 6 + 8 + 8 + 18
    for k in range(3):
        assert np.array_equal(a_[k], b_[k])
    assert np.abs(a_[3] - b_[3]).max() <= 1e-13 * np.abs(b_[3]).max()
    assert cpp["n_pairs"] == len(create_hamiltonian(dict(module="exchange-functional", energy_units="meV", distance_units="angstroms", lattice_parameter=a,
                                                         interactions=[("Fe", "Fe", "gaussian", 2.894, 12.0, 2.866, 0.86), ("Fe", "Co", "exponential", 2.58, 20.0, 2.29, 0.72),
                                                                       ("Co", "Fe", "exponential", 2.58, 20.0, 2.29, 0.72), ("Co", "Co", "rkky", 4.16, 3.0, 0.57, [1.3])]), lat).neighbour_list()[0])
    with pytest.raises(host.HostError, match="unknown exchange functional"):
        host.exchange_template(str(cfg), 'hamiltonians = ( { interactions = ( ("Fe", "Fe", "sinc", 2.894, 12.0, 2.866, 0.86) ); } );', ham_index=0)
    with pytest.raises(host.HostError, match="larger than the maximum cutoff radius"):
        host.exchange_template(str(cfg), 'hamiltonians = ( { interactions = ( ("Fe", "Fe", "gaussian", 28.0, 12.0, 2.866, 0.86) ); } );', ham_index=0)


def test_exchange_neartree_equals_exchange_with_explicit_vectors(tmp_path):
    """module = "exchange-neartree" (hamiltonian/exchange_neartree.cc): shells given by (material, material, radius, J) produce the
    same template as the `exchange` module given the shell's vectors with symmetry expansion (bcc: 8 NN at sqrt(3)/2, 6 NNN at 1)
    -- in the C++ host and in the Python mirror -- plus the reference's checks"""
    from jams_b200.lattice import Lattice, Material
    base = """
materials = ( { name = "Fe"; moment = 2.2; alpha = 0.1; spin = [0.0, 0.0, 1.0]; } );
unitcell : { parameter = 0.2866e-9; basis = ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]); positions = (("Fe", [0.0, 0.0, 0.0]), ("Fe", [0.5, 0.5, 0.5])); };
lattice : { size = [6, 5, 7]; periodic = [true, true, false]; };
solver : { module = "llg-heun-b200-gpu"; t_step = 1e-16; t_max = 1e-15; };
hamiltonians = ( %s );
"""
    shells = tmp_path / "shells.cfg"
    shells.write_text(base % '{ module = "exchange-neartree"; interactions = ( ("Fe", "Fe", 0.8660254037844386, 3.2e-21), ("Fe", "Fe", 1.0, 1.6e-21) ); }')
    vectors = tmp_path / "vectors.cfg"
    vectors.write_text(base % '{ module = "exchange"; interactions = ( ("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21) ); }')
    a_, b_ = _sorted_template(host.exchange_template(str(shells), ham_index=0)), _sorted_template(host.exchange_template(str(vectors), ham_index=0))
    assert len(a_[0]) == 28
    for x, y in zip(a_, b_):
        assert np.array_equal(x, y)
    lat = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (6, 5, 7), periodic=(True, True, False))
    py = create_hamiltonian(dict(module="exchange-neartree", interactions=[("Fe", "Fe", 0.8660254037844386, 3.2e-21), ("Fe", "Fe", 1.0, 1.6e-21)]), lat)
    for x, y in zip(_sorted_template(py.template), b_):
        assert np.array_equal(x, y)
    assert len(py.neighbour_list()[0]) == host.exchange_template(str(shells), ham_index=0)["n_pairs"]
    # a shell so wide that it holds both distances reaches the NN twice when listed next to the NN shell; tiny J is dropped; unknown material
    for patch, msg in (('hamiltonians = ( { shell_width = 0.3; } );', "multiple interactions"),
                       ('hamiltonians = ( { interactions = ( ("Fe", "Ni", 1.0, 1e-21) ); } );', "does not exist in the config")):
        with pytest.raises(host.HostError, match=msg):
            host.exchange_template(str(shells), patch, ham_index=0)
    with pytest.raises(RuntimeError, match="multiple interactions"):
        create_hamiltonian(dict(module="exchange-neartree", shell_width=0.3, interactions=[("Fe", "Fe", 0.8660254037844386, 3.2e-21), ("Fe", "Fe", 1.0, 1.6e-21)]), lat)
    dropped = host.exchange_template(str(shells), 'hamiltonians = ( { energy_cutoff = 2e-21; } );', ham_index=0)
    assert len(dropped["mi"]) == 16      # only the 8 + 8 nearest-neighbour entries survive


# ---- the reference's own known answers and shipped configuration files (copies under tests/golden/reference_cfg/, made by
# ---- tests/golden/make_reference_cfg_fixtures.py) ----------------------------------------------------------------------
REF_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cfg")


def _python_exchange_from_cfg(cfg, ham, base_dir):
    """the Python mirror fed from the parsed configuration (so both host layers see the same text)"""
    from jams_b200.lattice import Lattice, Material, read_interaction_file
    mats = [Material(m["name"], m["moment"], m.get("gyro", 1.0), m.get("alpha", 0.01)) for m in cfg["materials"]]
    lat = Lattice(mats, np.array(cfg["unitcell"]["basis"], float), [(p[0], tuple(p[1])) for p in cfg["unitcell"]["positions"]],
                  tuple(cfg["lattice"]["size"]), periodic=tuple(cfg["lattice"]["periodic"]))
    hs = dict(ham)
    if "exc_file" in hs:
        hs["interactions"] = read_interaction_file(os.path.join(base_dir, hs.pop("exc_file")))
    else:
        hs["interactions"] = [tuple(x) for x in hs["interactions"]]
    return create_hamiltonian(hs, lat)


def test_known_answer_crps4_local_point_groups_give_1536_interactions():
    """/root/reference/test/test_exchange_symops.py:14 ("computed interactions: 1536") on test/test_exchange_symops.cfg, unchanged:
    CrPS4 4^3, a low-symmetry cell where the crystal point group (with inversion) would invent interactions.  Needs the symmetry
    operations of the crystal itself (jams_host.cc find_point_operations / lattice.py find_space_group_operations)."""
    path = os.path.join(REF_CFG, "test_exchange_symops.cfg")
    t = host.exchange_template(path, ham_index=1)
    assert t["n_pairs"] == 1536 and len(t["mi"]) == 24
    cfg = host.config_to_dict(path)
    h = _python_exchange_from_cfg(cfg, cfg["hamiltonians"][1], REF_CFG)
    assert len(h.neighbour_list()[0]) == 1536
    for a, b in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(a, b)


def test_known_answer_yig_8_cubed_gives_217088_interactions():
    """src/jams/test/interactions.h:255-757 (generate_interactions_yig): 424 interactions per primitive cell x 8^3.  That (stale)
    gtest hands the complete list to generate_neighbour_list without the symmetry expansion, so `symops = false` here -- with
    symops the current reference would copy entry (i, j, r) to the images of r while keeping j (core/interactions.cc:24-37) and
    report 'Multiple interactions', which this host layer reproduces."""
    patch = 'hamiltonians = ( { module = "exchange"; exc_file = "yig_princep_exc.in"; symops = false; } );'
    cwd = os.getcwd()
    os.chdir(REF_CFG)   # exc_file is relative to the working directory, as in JAMS
    try:
        t = host.exchange_template("yig_8x8x8.cfg", patch, ham_index=0)
        with pytest.raises(host.HostError, match="Multiple interactions"):
            host.exchange_template("yig_8x8x8.cfg", ham_index=0)
    finally:
        os.chdir(cwd)
    assert t["n_pairs"] == 424 * 8 * 8 * 8 == 217088 and len(t["mi"]) == 424
    cfg = host.config_to_dict(os.path.join(REF_CFG, "yig_8x8x8.cfg"), patch)
    h = _python_exchange_from_cfg(cfg, cfg["hamiltonians"][0], REF_CFG)
    assert len(h.template["mi"]) == 424
    for a, b in zip(_sorted_template(t), _sorted_template(h.template)):
        assert np.array_equal(a, b)


def test_shipped_bloch_domain_wall_example_parses_unchanged():
    """examples/bloch_domain_wall/bloch_domain_wall.cfg as shipped: 256 x 16 x 16 sc, open along x -> 6 N - 2 * 16 * 16 pairs"""
    path = os.path.join(REF_CFG, "bloch_domain_wall.cfg")
    la = host.lattice_arrays(path)
    assert la["num_spins"] == 65536 and la["dims"] == (256, 16, 16) and la["periodic"] == (False, True, True)
    t = host.exchange_template(path, ham_index=1)
    assert len(t["mi"]) == 6 and t["n_pairs"] == 6 * 65536 - 2 * 16 * 16
    cfg = host.config_to_dict(path)
    assert cfg["solver"]["module"] == "llg-rk4-gpu" and cfg["physics"]["module"] == "pinned_boundaries" and cfg["monitors"][0]["module"] == "magnetisation-layers"
    assert cfg["initializer"] == {"module": "bloch_domain_wall", "width": 41.56, "center": 128.0}


def test_space_group_search_known_answers():
    """find_space_group_operations against textbook counts: sc O_h (48), conventional bcc (48 x 2 centring translations), the
    bcc-primitive YIG cell (Ia-3d: 48 operations, 6 of them without translation: the -3 axis through the origin)"""
    from jams_b200.lattice import cubic_point_group, find_space_group_operations
    R, T = find_space_group_operations(np.eye(3), [[0, 0, 0]], [0])
    assert sorted(tuple(r.ravel()) for r in R) == sorted(tuple(r.ravel()) for r in cubic_point_group()[0]) and np.all(T == 0)
    R, T = find_space_group_operations(np.eye(3), [[0, 0, 0], [0.5, 0.5, 0.5]], [0, 0])
    assert len(R) == 96 and sum(bool(np.allclose(t, 0)) for t in T) == 48
    R, T = find_space_group_operations(np.eye(3), [[0, 0, 0], [0.5, 0.5, 0.5]], [0, 1])   # CsCl: the centring is gone
    assert len(R) == 48
    cfg = host.config_to_dict(os.path.join(REF_CFG, "yig_8x8x8.cfg"))
    pos = [p[1] for p in cfg["unitcell"]["positions"]]
    types = [0 if p[0] == "FeA" else 1 for p in cfg["unitcell"]["positions"]]
    R, T = find_space_group_operations(np.array(cfg["unitcell"]["basis"], float), pos, types)
    assert len(R) == 48 and sum(bool(np.allclose(t, 0)) for t in T) == 6


def _read_mag_layers(path):
    layers, rows = {}, []
    for line in open(path):
        t = line.split()
        if line.startswith("# group"):
            layers[int(t[4])] = float(t[6])
        elif not line.startswith("#") and t:
            rows.append((int(t[0]), float(t[1]), t[2], int(t[3]), float(t[4]), float(t[5]), float(t[6])))
    return layers, rows


@pytest.mark.gpu
def test_shipped_bloch_domain_wall_example_runs_unchanged(tmp_path):
    """examples/bloch_domain_wall/bloch_domain_wall.cfg exactly as shipped (llg-rk4-gpu + pinned_boundaries at T = 1 K +
    bloch_domain_wall initializer + magnetisation-layers monitor, N = 65 536); only t_max is shortened by a command-line
    patch, the way a JAMS user would (core/jams++.cc:48-84)"""
    cfg = os.path.join(REF_CFG, "bloch_domain_wall.cfg")
    got, done = host.run(cfg, "solver : { t_max = 2.0e-12; };", name="wall", output_dir=str(tmp_path))
    assert done == 2000 and got.shape == (65536, 3)
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() < 1e-12
    layers, rows = _read_mag_layers(tmp_path / "wall_mag_layers.tsv")
    assert len(layers) == 256 and abs(layers[1] - layers[0] - 0.3) < 1e-9        # layer spacing = lattice parameter 0.3 nm
    assert sorted({r[0] for r in rows}) == [0, 1000] or sorted({r[0] for r in rows}) == [0, 1000, 2000]
    last = [r for r in rows if r[0] == max(q[0] for q in rows)]
    mz = np.array([r[6] for r in sorted(last, key=lambda r: r[3])]) / (3.0 * 256)   # moment 3 mu_B x 256 spins per layer
    x = np.arange(256)
    assert np.abs(mz - np.tanh(np.pi * (x - 128.0) / 41.56)).max() < 0.05          # 1 K: the profile stays where the initializer put it
    assert mz[0] < -0.99 and mz[-1] > 0.99                                         # pinned ends


@pytest.mark.gpu
def test_bloch_wall_relaxes_to_the_analytic_width(tmp_path):
    """BASELINE.md 1 / bloch_domain_wall.cfg:84-106: w/a = pi sqrt(J / 2k) = 41.56 at T = 0.  Start the shipped example's wall 28 %
    too narrow (width 30), relax it with the llg-heun path at T = 0 (alpha = 1 to get there in 60 ps; the fixed point does not
    depend on alpha) and fit m_z(x) = tanh(pi (x - c) / w) to the final state."""
    from scipy.optimize import curve_fit
    cfg = os.path.join(REF_CFG, "bloch_domain_wall.cfg")
    patch = ('lattice : { size = [256, 4, 4]; }; initializer : { width = 30.0; }; physics : { module = "empty"; temperature = 0.0; }; '
             'materials = ( { name = "A"; moment = 3.0; alpha = 1.0; spin = [0.0, 0.0, 1.0]; } ); '
             'solver : { module = "llg-heun-gpu"; t_step = 5.0e-16; t_max = 6.0e-11; }; '
             'monitors = ( { module = "magnetisation-layers"; output_steps = 40000; layer_normal = [1, 0, 0]; } );')
    got, done = host.run(cfg, patch, name="relax", output_dir=str(tmp_path))
    assert done == 120000
    mz = got.reshape(256, 4, 4, 3)[..., 2].mean(axis=(1, 2))
    x = np.arange(256, dtype=float)
    (c, wfit), _ = curve_fit(lambda xx, c, w: np.tanh(np.pi * (xx - c) / w), x, mz, p0=(128.0, 30.0))
    assert abs(wfit - 41.56) <= 0.01 * 41.56, wfit
    assert abs(c - 127.5) < 1.0 or abs(c - 128.0) < 1.0
    # the first output of the layers monitor holds the narrow initial wall, the last one the relaxed wall
    layers, rows = _read_mag_layers(tmp_path / "relax_mag_layers.tsv")
    first = np.array([r[6] for r in sorted([r for r in rows if r[0] == 0], key=lambda r: r[3])]) / (3.0 * 16)
    (c0, w0), _ = curve_fit(lambda xx, c, w: np.tanh(np.pi * (xx - c) / w), x, first, p0=(128.0, 30.0))
    assert abs(w0 - 30.0) < 0.05
