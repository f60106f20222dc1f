"""Host-side mirror of the JAMS plugin surface for the llg-heun hot path.

Names, argument meaning and error behaviour follow the reference's classes so that code (and tests)
written against JAMS read the same here:

=========================  ==========================================================================
reference                   here
=========================  ==========================================================================
``Solver``                  :class:`Solver` (core/solver.h:15-90): ``initialize``, ``run``, ``time``,
                            ``iteration``, ``is_running``, ``register_hamiltonian``,
                            ``register_monitor``, ``notify_monitors``, ``compute_fields``
``CUDAHeunLLGSolver``       :class:`B200HeunLLGSolver`, module name ``"llg-heun-b200-gpu"``
                            (solvers/cuda_llg_heun.cu:21-122; maths of solvers/cpu_llg_heun.cc:45-148)
``Hamiltonian``             :class:`Hamiltonian` (core/hamiltonian.h:15-78): ``calculate_fields``,
                            ``calculate_energies``, ``calculate_total_energy``
``ExchangeHamiltonian``     :class:`ExchangeHamiltonian`      (hamiltonian/exchange.cc:12-172)
``UniaxialAnisotropy...``   :class:`UniaxialAnisotropyHamiltonian` (hamiltonian/uniaxial_anisotropy.cc:79-115)
``ZeemanHamiltonian``       :class:`ZeemanHamiltonian`        (hamiltonian/zeeman.cc:12-72)
``AppliedFieldHamiltonian`` :class:`AppliedFieldHamiltonian`  (hamiltonian/applied_field.cc:84-148, static type)
``Thermostat``              :class:`LangevinWhiteThermostat`  (thermostats/cuda_thermostat_classical.cc:22-56)
``MagnetisationMonitor``    :class:`MagnetisationMonitor`     (monitors/magnetisation.cc:21-104)
``EnergyMonitor``           :class:`EnergyMonitor`            (monitors/energy.cc:16-33)
=========================  ==========================================================================

All numerical work happens in ``libjams_b200.so`` through :mod:`jams_b200.capi`; this module only holds
parameters and sequences calls (as the C++ adapter in INTEGRATION.md does inside JAMS itself).
"""
from __future__ import annotations

import numpy as np

from . import capi
from .consts import ENERGY_UNITS, kBohrMagnetonIU
from .lattice import Lattice


class Hamiltonian:
    """core/hamiltonian.h:15-78.  ``energy_units`` defaults to joules (helpers/defaults.h:25)."""
    term = None
    name = "hamiltonian"

    def __init__(self, settings: dict, lattice: Lattice):
        self.settings = dict(settings)
        self.lattice = lattice
        unit_name = self.settings.get("unit_name", self.settings.get("energy_units", "joules"))
        if unit_name not in ENERGY_UNITS:
            raise RuntimeError(f"energy units: {unit_name} is not known")
        self.input_energy_unit_conversion = ENERGY_UNITS[unit_name]
        self.input_energy_unit_name = unit_name
        self.solver = None

    # push parameters for the slab [x0, x0+nx) into a context
    def attach(self, ctx: capi.Context, x0: int, nx: int):
        raise NotImplementedError

    def calculate_fields(self, time: float):
        """field_ of this term, N x 3, meV (not divided by mu)"""
        self.solver._build()
        return self.solver.ctx.fields(self.term, time)

    def calculate_energies(self, time: float):
        self.solver._build()
        return self.solver.ctx.energies(self.term, time, per_spin=True)[0]

    def calculate_total_energy(self, time: float):
        self.solver._build()
        return self.solver.ctx.energies(self.term, time, per_spin=False)[1]


class ExchangeHamiltonian(Hamiltonian):
    term = capi.TERM_EXCHANGE
    name = "exchange"

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings, lattice)
        s = self.settings
        if "exc_file" in s:          # kept by the reference for backwards compatibility; wins over 'interactions' (exchange.cc:118-145)
            from .lattice import read_interaction_file
            interactions = read_interaction_file(s["exc_file"])
        elif "interactions" in s:
            interactions = s["interactions"]
        else:
            raise RuntimeError("'exc_file' or 'interactions' settings are required")
        self.template = lattice.expand_interactions(
            interactions, energy_units=self.input_energy_unit_name,
            coordinate_format=s.get("coordinate_format", "cartesian"), use_symops=s.get("symops", True),
            energy_cutoff=s.get("energy_cutoff", 0.0), radius_cutoff=s.get("radius_cutoff", 100.0),
            distance_tolerance=s.get("distance_tolerance", 1e-4), interaction_prefactor=s.get("interaction_prefactor", 1.0))
        self.use_pairs = bool(s.get("use_neighbour_list", False)) or lattice.has_impurities   # impurities break translation invariance
        self._nbr = None

    def neighbour_list(self):
        if self._nbr is None:
            self._nbr = self.lattice.neighbour_list(self.template)
        return self._nbr

    def attach(self, ctx, x0, nx):
        # check_sparse_matrix_symmetry = false switches the symmetry check off (hamiltonian/exchange.cc:104-110); default: checked
        ctx.set_option("check_symmetry", 0 if self.settings.get("check_sparse_matrix_symmetry", True) is False else 1)
        if self.use_pairs:
            i, j, v, vals = self.neighbour_list()
            ctx.set_exchange_pairs(i, j, v, vals)
        else:
            t = self.template
            ctx.set_exchange_template(t["mi"], t["mj"], t["T"], t["J9"])


class BiquadraticExchangeHamiltonian(Hamiltonian):
    """``module = "biquadratic-exchange"`` (hamiltonian/cuda_biquadratic_exchange.{h,cu}): E = -1/2 sum_ij B_ij (s_i . s_j)^2 with the
    field h_i = sum_j 2 B_ij s_j (s_i . s_j).  Same ``interactions`` / ``exc_file`` grammar and cutoffs as ``exchange``; the scalar
    B_ij is element [0][0] of the interaction tensor times the unit conversion (no ``interaction_prefactor``), and only values
    above ``energy_cutoff`` are kept (cuda_biquadratic_exchange.cu:127-134: negative couplings are dropped at the default cutoff
    of 0, as in the reference)."""
    term = capi.TERM_BIQUADRATIC
    name = "biquadratic-exchange"

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings, lattice)
        if lattice.has_impurities:   # the library takes this term as a translation-invariant template only
            raise RuntimeError(self.name + " is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer")
        s = self.settings
        if "exc_file" in s:
            from .lattice import read_interaction_file
            interactions = read_interaction_file(s["exc_file"])
        elif "interactions" in s:
            interactions = s["interactions"]
        else:
            raise RuntimeError("'exc_file' or 'interactions' settings are required for exchange hamiltonian")
        t = lattice.expand_interactions(
            interactions, energy_units=self.input_energy_unit_name,
            coordinate_format=s.get("coordinate_format", "cartesian"), use_symops=s.get("symops", True),
            energy_cutoff=s.get("energy_cutoff", 0.0), radius_cutoff=s.get("radius_cutoff", 100.0),
            distance_tolerance=s.get("distance_tolerance", 1e-4), interaction_prefactor=1.0)
        B = np.asarray(t["J9"], dtype=np.float64).reshape(-1, 9)[:, 0]
        keep = B > s.get("energy_cutoff", 0.0) * self.input_energy_unit_conversion
        self.template = dict(mi=np.asarray(t["mi"])[keep], mj=np.asarray(t["mj"])[keep], T=np.asarray(t["T"]).reshape(-1, 3)[keep], B=B[keep])

    def attach(self, ctx, x0, nx):
        ctx.set_option("check_symmetry", 0 if self.settings.get("check_sparse_matrix_symmetry", True) is False else 1)
        t = self.template
        ctx.set_biquadratic_template(t["mi"], t["mj"], t["T"], t["B"])


class ExchangeFunctionalHamiltonian(ExchangeHamiltonian):
    """``module = "exchange-functional"`` (hamiltonian/exchange_functional.{h,cc}): isotropic exchange J(r_ij) from a closed
    form inside a cutoff radius, one entry per ordered material pair:
    ``interactions = ( (type_i, type_j, functional, r_cutoff, params...), ... )`` with functionals ``step`` (J0, r_cut),
    ``exponential`` / ``gaussian`` / ``kaneyoshi`` (J0, r0, sigma), ``gaussian_multi`` (three such triples), ``rkky``
    (J0, r0, k_F) and ``c3z`` (14 parameters); energies in ``energy_units``, lengths in ``distance_units``
    (lattice_constants | nanometers | angstroms, core/hamiltonian.cc:140-155).  The reference fills the same scalar CSR
    matrix as ``exchange`` (SparseInteractionHamiltonian); here the list becomes an exchange template for the same kernels."""
    name = "exchange-functional"
    _NPAR = {"rkky": 3, "exponential": 3, "gaussian": 3, "gaussian_multi": 9, "kaneyoshi": 3, "c3z": 14, "step": 2}

    def __init__(self, settings: dict, lattice: Lattice, lattice_parameter: float | None = None):
        Hamiltonian.__init__(self, settings, lattice)
        if lattice.has_impurities:   # the template form below assumes a translation-invariant lattice
            raise RuntimeError(self.name + " is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer")
        s = self.settings
        dunit = s.get("distance_units", "lattice_constants")
        a = lattice_parameter if lattice_parameter is not None else s.get("lattice_parameter", getattr(lattice, "parameter", None))
        conv = {"lattice_constants": 1.0}
        if a:
            conv.update(nanometers=1e-9 / float(a), angstroms=1e-10 / float(a))
        if dunit not in conv:
            raise RuntimeError(f"distance units: {dunit} is not known")
        self.input_distance_unit_conversion = conv[dunit]
        if "interactions" not in s:
            raise RuntimeError("no 'interactions' setting in ExchangeFunctional hamiltonian")
        functionals, rmax = {}, 0.0
        for entry in s["interactions"]:
            if len(entry) < 4:
                raise RuntimeError("interaction requires at least 4 elements")
            ti, tj, fname = str(entry[0]), str(entry[1]), str(entry[2])
            rc = self.input_distance_unit_conversion * float(entry[3])
            for t in (ti, tj):
                if t not in lattice.material_index:
                    raise RuntimeError(f"material {t} does not exist in config")
            if rc < 0.0 and (0.0 - rc) > abs(rc) * 1e-4:
                raise RuntimeError("cutoff radius cannot be negative")
            if (ti, tj) in functionals:
                raise RuntimeError(f'Interaction between types "{ti}" and "{tj}" is defined more than once.')
            if rc > lattice.max_interaction_radius():
                raise RuntimeError(f"cutoff radius {rc:f} is larger than the maximum cutoff radius {lattice.max_interaction_radius():f}")
            params = []
            for v in entry[4:]:
                params.extend([float(x) for x in v] if isinstance(v, (list, tuple, np.ndarray)) else [float(v)])
            functionals[(ti, tj)] = (rc, self._functional(fname, params))
        self.template = lattice.functional_template(functionals)
        self.use_pairs = bool(s.get("use_neighbour_list", False))
        self._nbr = None

    def _functional(self, name, p):
        """validate_functional_params + functional_from_params (exchange_functional.cc:13-88,358-416)"""
        if name not in self._NPAR:
            raise RuntimeError("unknown exchange functional: " + name)
        if len(p) != self._NPAR[name]:
            raise RuntimeError(f"exchange functional '{name}' expects {self._NPAR[name]} parameters, got {len(p)}")
        E, D, tol = self.input_energy_unit_conversion, self.input_distance_unit_conversion, 1e-4
        nonzero = {"rkky": [(2, "k_F")], "exponential": [(2, "sigma")], "gaussian": [(2, "sigma")], "kaneyoshi": [(2, "sigma")],
                   "gaussian_multi": [(2, "sigma0"), (5, "sigma1"), (8, "sigma2")]}.get(name, [])
        for idx, pname in nonzero:
            if abs(p[idx]) <= tol:
                raise RuntimeError(f"exchange functional '{name}' requires non-zero parameter '{pname}'")
        if name == "c3z":
            for idx, pname in ((10, "l0"), (11, "l1s"), (12, "l1c")):
                if not (p[idx] - 0.0) > abs(p[idx]) * tol:
                    raise RuntimeError(f"exchange functional 'c3z' requires positive parameter '{pname}'")
        norm = np.linalg.norm
        gauss = lambda r, J0, r0, sg: J0 * np.exp(-(r - r0) ** 2 / (2 * sg ** 2))   # noqa: E731
        if name == "step":
            J0, rc = E * p[0], D * p[1]
            return lambda rij: J0 if (norm(rij) - rc) < max(abs(norm(rij)), abs(rc)) * tol else 0.0
        if name == "exponential":
            J0, r0, sg = E * p[0], D * p[1], D * p[2]
            return lambda rij: J0 * np.exp(-(norm(rij) - r0) / sg)
        if name == "gaussian":
            J0, r0, sg = E * p[0], D * p[1], D * p[2]
            return lambda rij: gauss(norm(rij), J0, r0, sg)
        if name == "gaussian_multi":
            q = [(E * p[3 * k], D * p[3 * k + 1], D * p[3 * k + 2]) for k in range(3)]
            return lambda rij: sum(gauss(norm(rij), *t) for t in q)
        if name == "kaneyoshi":
            J0, r0, sg = E * p[0], D * p[1], D * p[2]
            return lambda rij: J0 * (norm(rij) - r0) ** 2 * np.exp(-(norm(rij) - r0) ** 2 / (2 * sg ** 2))
        if name == "rkky":
            J0, r0, kF = E * p[0], D * p[1], p[2]

            def rkky(rij):
                kr = 2 * kF * (norm(rij) - r0)
                if abs(kr) <= tol:
                    raise RuntimeError("exchange functional rkky is singular for k_F*(r-r0) = 0")
                return -J0 * (kr * np.cos(kr) - np.sin(kr)) / kr ** 4
            return rkky
        # c3z (exchange_functional.cc:310-356)
        qs1, qc1 = np.array(p[0:3]) / D, np.array(p[3:6]) / D
        J0, J1s, J1c = E * p[6], E * p[7], E * p[8]
        d0, l0, l1s, l1c, rstar = (D * v for v in p[9:14])

        def rotz(t):
            return np.array([[np.cos(t), -np.sin(t), 0.0], [np.sin(t), np.cos(t), 0.0], [0.0, 0.0, 1.0]])

        def c3z(rij):
            r = norm(rij)
            rpar = np.array([rij[0], rij[1], 0.0])
            ssum = sum(np.sin(np.dot(rotz(t) @ qs1, rpar)) for t in (0.0, 2 * np.pi / 3, 4 * np.pi / 3))
            csum = sum(np.cos(np.dot(rotz(t) @ qc1, rpar)) for t in (0.0, 2 * np.pi / 3, 4 * np.pi / 3))
            return J0 * np.exp(-abs(r - d0) / l0) + J1s * np.exp(-abs(r - rstar) / l1s) * ssum + J1c * np.exp(-abs(r - rstar) / l1c) * csum
        return c3z


class ExchangeNeartreeHamiltonian(ExchangeHamiltonian):
    """``module = "exchange-neartree"`` (hamiltonian/exchange_neartree.{h,cc}): isotropic exchange by distance shells,
    ``interactions = ( (material A, material B, radius, J), ... )``: every A site couples with J to the B sites at distance
    radius -+ ``shell_width`` / 2 (default 1e-3), and the mirrored entry is added for A != B (:78-88); couplings with
    |J| <= ``energy_cutoff`` (default 1e-26 in the input units) are dropped (:123-126).  Same scalar CSR matrix as ``exchange``
    in the reference; here a template for the same kernels."""
    name = "exchange-neartree"

    def __init__(self, settings: dict, lattice: Lattice, lattice_parameter: float | None = None):
        Hamiltonian.__init__(self, settings, lattice)
        if lattice.has_impurities:   # the template form below assumes a translation-invariant lattice
            raise RuntimeError(self.name + " is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer")
        s = self.settings
        dunit = s.get("distance_units", "lattice_constants")
        a = lattice_parameter if lattice_parameter is not None else s.get("lattice_parameter", getattr(lattice, "parameter", None))
        conv = {"lattice_constants": 1.0}
        if a:
            conv.update(nanometers=1e-9 / float(a), angstroms=1e-10 / float(a))
        if dunit not in conv:
            raise RuntimeError(f"distance units: {dunit} is not known")
        D, E = conv[dunit], self.input_energy_unit_conversion
        energy_cutoff = float(s.get("energy_cutoff", 1e-26)) * E
        shell_width = float(s.get("shell_width", 1e-3)) * D
        for i in range(lattice.M):             # exchange_neartree.cc:43-56
            for j in range(i + 1, lattice.M):
                d = float(np.linalg.norm(lattice.motif_frac[i] - lattice.motif_frac[j]))
                if d < shell_width:
                    raise RuntimeError(f"Atoms {i} and {j} in the unit cell are close together ({d}) than the shell_width ({shell_width}).")
        if "interactions" not in s:
            raise RuntimeError("no 'interactions' setting in ExchangeNeartree hamiltonian")
        shells = []
        for n, (ta, tb, radius, J) in enumerate(s["interactions"]):
            for t in (ta, tb):
                if t not in lattice.material_index:
                    raise RuntimeError(f"exchange neartree interaction {n}: material {t} does not exist in the config")
            A, B = lattice.material_index[ta], lattice.material_index[tb]
            shells.append((A, B, float(radius) * D, float(J) * E))
            if A != B:
                shells.append((B, A, float(radius) * D, float(J) * E))
        self.template = lattice.shell_template(shells, shell_width, energy_cutoff)
        self.use_pairs = bool(s.get("use_neighbour_list", False))
        self._nbr = None


class UniaxialAnisotropyHamiltonian(Hamiltonian):
    term = capi.TERM_UNIAXIAL
    name = "uniaxial"
    slot = 0          # the n-th uniaxial Hamiltonian of a configuration (register_hamiltonian): term = capi.UNIAXIAL_TERMS[slot]
    _POWER = {"K1": 2, "K2": 4, "K3": 6}

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings, lattice)
        s = self.settings
        for old in ("d2z", "d4z", "d6z", "K1", "K2", "K3"):
            if old in s:
                raise RuntimeError("UniaxialHamiltonian: anisotropy should only be specified for a single K1, K2 or K3.")
        if s["order"] not in self._POWER:
            raise RuntimeError("Unsupported anisotropy: " + str(s["order"]))
        self.power = self._POWER[s["order"]]
        M = lattice.M
        self.motif_K = np.zeros(M)
        self.motif_axis = np.zeros((M, 3))
        self.by_material = {}   # material index -> (K, axis): what a site with that material gets when the lattice has impurities
        for who, axis, energy in s["anisotropies"]:
            axis = np.asarray(axis, dtype=np.float64)
            axis = axis / np.sqrt(axis @ axis)  # normalize(), uniaxial_anisotropy.cc:65
            for m in range(M):
                if isinstance(who, (int, np.integer)):
                    if who - 1 < 0 or who - 1 >= M:
                        raise RuntimeError("uniaxial anisotropy motif position is invalid")
                    hit = (m == who - 1)
                else:
                    if who not in lattice.material_index:
                        raise RuntimeError("uniaxial anisotropy material is invalid")
                    hit = lattice.motif_material[m] == lattice.material_index[who]
                if hit:
                    self.motif_K[m] = energy * self.input_energy_unit_conversion
                    self.motif_axis[m] = axis
            if not isinstance(who, (int, np.integer)):
                self.by_material[lattice.material_index[who]] = (energy * self.input_energy_unit_conversion, axis)

    def site_arrays(self, x0=0, nx=None):
        lat = self.lattice
        K, axis = lat._tile(self.motif_K, x0, nx), lat._tile(self.motif_axis, x0, nx)
        if lat.has_impurities:   # by material name: the site's own material decides (uniaxial_anisotropy.cc:89-114)
            K, axis = np.array(K), np.array(axis)
            mat = lat.site_material(x0, nx)
            motif_mat = lat._tile(lat.motif_material, x0, nx)
            changed = mat != motif_mat
            K[changed] = 0.0; axis[changed] = 0.0
            for m_idx, (k, a) in self.by_material.items():
                sel = changed & (mat == m_idx)
                K[sel] = k; axis[sel] = a
        return K, axis

    def attach(self, ctx, x0, nx):
        K, axis = self.site_arrays(x0, nx)
        ctx.set_uniaxial(self.power, K, axis, slot=self.slot)


class ZeemanHamiltonian(Hamiltonian):
    term = capi.TERM_ZEEMAN
    name = "zeeman"

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings, lattice)
        s = self.settings
        nmat = len(lattice.materials)
        self.dc = np.zeros((nmat, 3))
        if "dc_local_field" in s:
            if len(s["dc_local_field"]) != nmat:
                raise RuntimeError("dc_local_field: field must be specified for every material")
            self.dc = np.asarray(s["dc_local_field"], dtype=np.float64).reshape(nmat, 3)
        self.has_ac = ("ac_local_field" in s) or ("ac_local_frequency" in s)
        if self.has_ac:
            if not ("ac_local_field" in s and "ac_local_frequency" in s):
                raise RuntimeError("ac_local_field: must have a field and a frequency")
            if len(s["ac_local_field"]) != nmat or len(s["ac_local_frequency"]) != nmat:
                raise RuntimeError("ac_local_frequency: must be specified for every material")
            self.ac = np.asarray(s["ac_local_field"], dtype=np.float64).reshape(nmat, 3)
            self.freq = 2.0 * np.pi * np.asarray(s["ac_local_frequency"], dtype=np.float64)

    def site_arrays(self, x0=0, nx=None):
        lat = self.lattice
        mat = lat.site_material(x0, nx)
        mus = lat.mus(x0, nx)
        dc = self.dc[mat] * mus[:, None]                   # zeeman.cc:32-37
        if self.has_ac:
            return dc, self.ac[mat] * mus[:, None], self.freq[mat]
        return dc, None, None

    def attach(self, ctx, x0, nx):
        ctx.set_zeeman(*self.site_arrays(x0, nx))


class AppliedFieldHamiltonian(Hamiltonian):
    """``module = "applied-field"`` (hamiltonian/applied_field.{h,cc}): homogeneous B(t) = ``field`` x g(t), the field on site i is
    mu_i B(t).  ``type``: ``static`` (default, g = 1), ``sinc`` (g = sinc(pi f_bw (t - t0))), ``sinc-cos`` (... x cos(2 pi f_c (t - t0)));
    ``time_center`` in seconds, ``freq_bandwidth`` / ``freq_center`` in Hz (converted to ps / THz like applied_field.cc:37-38,66-68)."""
    term = capi.TERM_APPLIED
    name = "applied-field-static"

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings, lattice)
        kind = str(self.settings.get("type", "static")).lower()
        if kind not in ("static", "sinc", "sinc-cos"):
            raise RuntimeError("Unknown field pulse type " + kind)
        self.kind = kind
        self.name = "applied-field-" + kind
        self.field = np.asarray(self.settings["field"], dtype=np.float64)
        self.time_center = float(self.settings["time_center"]) / 1e-12 if kind != "static" else 0.0
        self.freq_bandwidth = float(self.settings["freq_bandwidth"]) / 1e12 if kind != "static" else 0.0
        self.freq_center = float(self.settings["freq_center"]) / 1e12 if kind == "sinc-cos" else 0.0

    def attach(self, ctx, x0, nx):
        if self.kind == "static":
            ctx.set_applied_field(self.field, True)
        else:
            ctx.set_applied_field_pulse(self.field, self.kind, self.time_center, self.freq_bandwidth, self.freq_center)


_HAMILTONIANS = {"exchange": ExchangeHamiltonian, "exchange-functional": ExchangeFunctionalHamiltonian,
                 "exchange-neartree": ExchangeNeartreeHamiltonian, "uniaxial": UniaxialAnisotropyHamiltonian,
                 "zeeman": ZeemanHamiltonian, "applied-field": AppliedFieldHamiltonian, "biquadratic-exchange": BiquadraticExchangeHamiltonian}


def create_hamiltonian(settings: dict, lattice: Lattice) -> Hamiltonian:
    """Hamiltonian::create (core/hamiltonian.cc:80-115): dispatch on lower-cased ``module``"""
    module = str(settings["module"]).lower()
    if module not in _HAMILTONIANS:
        raise RuntimeError("unknown hamiltonian " + module)
    return _HAMILTONIANS[module](settings, lattice)


class LangevinWhiteThermostat:
    """thermostats/cuda_thermostat_classical.cc:22-56: xi_ij = sigma_i sqrt(T) n_ij (Tesla).  The draws come
    from Philox4x32-10 keyed by (seed; global site, step) inside the stage kernels; this object only exposes
    them (Thermostat::device_data) and carries the temperature."""

    def __init__(self, solver, seed=0):
        self.solver, self.seed, self.temperature = solver, int(seed), 0.0

    def set_temperature(self, T):
        self.temperature = float(T)

    def noise(self, step=None):
        s = self.solver
        s._build()
        return s.ctx.noise(s.step_size, self.temperature, self.seed, s.iteration if step is None else step,
                           s.lattice.gilbert_prefactor)


class Monitor:
    def __init__(self, settings: dict | None = None):
        settings = settings or {}
        self.output_step_freq = int(settings.get("output_steps", 100))  # helpers/defaults.h:28
        self.records = []

    def is_updating(self, iteration):  # core/monitor.cc
        return iteration % self.output_step_freq == 0

    def update(self, solver):
        raise NotImplementedError


class MagnetisationMonitor(Monitor):
    """monitors/magnetisation.cc:21-104, grouping = materials (default) | positions | none"""

    def __init__(self, settings=None, lattice: Lattice | None = None):
        super().__init__(settings)
        settings = settings or {}
        self.grouping = str(settings.get("grouping", "materials")).lower()
        if self.grouping not in ("none", "materials", "positions"):
            raise RuntimeError("unknown magnetisation grouping: " + self.grouping)
        self.normalize = bool(settings.get("normalize", True))
        self._registered = None   # the context whose device copy of the group array is current

    def groups(self, solver):
        lat = solver.lattice
        if self.grouping == "none":
            return None, 1
        if self.grouping == "materials":
            return lat.site_material(solver.x0, solver.nx), len(lat.materials)
        return lat.site_motif(solver.x0, solver.nx), lat.M

    def update(self, solver):
        solver._build()
        if self._registered is not solver.ctx:   # the groups are fixed at construction (monitors/magnetisation.cc:21-60): upload once
            g, ng = self.groups(solver)
            solver.ctx.set_magnetisation_groups(g, ng)
            self._registered, self._ng = solver.ctx, ng
        ng = self._ng
        M4 = solver.reduce_sum(solver.ctx.magnetisation(None, ng))
        row = [solver.time, solver.temperature]
        for n in range(ng):
            mag = M4[n, :3]
            factor = 1.0 / M4[n, 3] if self.normalize else 1.0 / kBohrMagnetonIU
            row += [mag[0] * factor, mag[1] * factor, mag[2] * factor, float(np.sqrt(mag @ mag)) * factor]
        self.records.append(row)
        return row


class EnergyMonitor(Monitor):
    """monitors/energy.cc:22-33: one total per registered Hamiltonian, meV"""

    def update(self, solver):
        row = [solver.time] + [float(solver.reduce_sum(np.array([h.calculate_total_energy(solver.time)]))[0])
                               for h in solver.hamiltonians]
        self.records.append(row)
        return row


class Physics:
    """core/physics.h:14-40 with physics/empty.h: constant ``temperature`` and ``applied_field`` from the ``physics`` group"""

    def __init__(self, settings: dict | None = None):
        settings = settings or {}
        self.temperature = float(settings.get("temperature", 0.0))
        self.applied_field = np.asarray(settings.get("applied_field", [0.0, 0.0, 0.0]), dtype=np.float64)

    def update(self, solver):   # Physics::update(iterations, time, dt), called once per iteration (core/jams++.cc:334)
        pass


class PinnedBoundariesPhysics(Physics):
    """``physics.module = "pinned_boundaries"`` (physics/pinned_boundaries.{h,cc}): every iteration the spins of each edge
    region are rotated so that the region's moment sum mu_i s_i points along the pinned direction.  Regions: ``left/right``
    (a), ``front/back`` (b), ``bottom/top`` (c), ``<name>_pinned_cells`` unit cells deep (default 1)
    (pinned_boundaries.h:86-107).  The reduction and the rotation run on the device (``jb_region_moment``,
    ``jb_rotate_region``); the 3x3 rotation is built here from the (all-reduced) moment like the reference does
    (``rotation_matrix_between_vectors``, containers/mat3.h:334-366)."""
    NAMES = {"left": (0, False), "right": (0, True), "front": (1, False), "back": (1, True), "bottom": (2, False), "top": (2, True)}

    def __init__(self, settings: dict, lattice: Lattice):
        super().__init__(settings)
        self.lattice = lattice
        self.boundaries = []   # (name, dim, upper, n_cells, pinned magnetisation)
        for name, (dim, upper) in self.NAMES.items():
            if name + "_pinned_magnetisation" not in settings:
                continue
            m = np.asarray(settings[name + "_pinned_magnetisation"], dtype=np.float64)
            self.boundaries.append((name, dim, upper, int(settings.get(name + "_pinned_cells", 1)), m))
        self._attached = None

    def region_sites(self, dim, upper, n, x0, nx):
        """local site ids (reference order) of the spins whose cell offset lies in the region (pinned_boundaries.cc:21-27)"""
        lat = self.lattice
        Nx, Ny, Nz = lat.dims
        x, y, z, m = np.meshgrid(np.arange(x0, x0 + nx), np.arange(Ny), np.arange(Nz), np.arange(lat.M), indexing="ij")
        cell = (x, y, z)[dim]
        size = lat.dims[dim]
        mask = (cell >= size - n) if upper else (cell < n)
        local = (((x - x0) * Ny + y) * Nz + z) * lat.M + m
        return local[mask].astype(np.int32)

    def attach(self, solver):
        if self._attached is solver:
            return
        for r, (name, dim, upper, n, m) in enumerate(self.boundaries):
            solver.ctx.set_region(r, self.region_sites(dim, upper, n, solver.x0, solver.nx))
        self._attached = solver

    def update(self, solver):
        from .lattice import rotation_matrix_between_vectors
        solver._build()
        self.attach(solver)
        for r, (name, dim, upper, n, target) in enumerate(self.boundaries):
            mag = solver.reduce_sum(solver.ctx.region_moment(r))[:3]
            solver.ctx.rotate_region(r, rotation_matrix_between_vectors(mag, target))


class FieldCoolPhysics(Physics):
    """``physics.module = "field-cool"`` (physics/field_cool.{h,cc}): temperature (and the reported applied field) ramp from
    ``InitialTemperature`` / ``InitialField`` to ``FinalTemperature`` / ``FinalField`` over ``CoolTime`` -- continuously
    (per-iteration increments, field_cool.cc:64-76) or, with ``TSteps``, in that many plateaus after ``sim.t_eq``
    (:58-63).  The solver re-reads the temperature every step (core/solver.cc:94-97).  Restated literally, including the
    reference's mixed units: ``time`` arrives in ps (core/solver.cc:85-87) while ``CoolTime``, ``sim.t_eq`` and
    ``sim.t_step`` are used as written in the config file."""

    def __init__(self, settings: dict, sim: dict | None = None):
        super().__init__(settings)
        sim = sim or {}
        self.init_temp = float(settings["InitialTemperature"])
        self.final_temp = float(settings["FinalTemperature"])
        self.integration_time_step = float(sim.get("t_step", settings.get("t_step", 0.0)))   # globals::config->lookup("sim.t_step")
        self.init_field = np.asarray(settings["InitialField"], dtype=np.float64)
        self.final_field = np.asarray(settings["FinalField"], dtype=np.float64)
        self.cool_time = float(settings["CoolTime"])
        self.applied_field = self.applied_field + self.init_field        # field_cool.cc:37-39
        self.t_eq = 0.0
        self.step_toggle = "TSteps" in settings
        if self.step_toggle:                                             # :41-49
            self.n_steps = int(settings["TSteps"])
            self.delta_t = (self.init_temp - self.final_temp) / self.n_steps
            self.t_plateau = self.cool_time / self.n_steps
            self.t_eq = float(sim.get("t_eq", 0.0))
        self.temperature = self.init_temp

    def update(self, solver):
        time = solver.time
        if not time > self.t_eq:
            return
        if self.step_toggle:
            count = int((time - self.t_eq) / self.t_plateau)
            if count < self.n_steps + 1:
                self.temperature = self.init_temp - count * self.delta_t
        elif time < self.cool_time:
            self.applied_field = self.applied_field + (self.final_field - self.init_field) * self.integration_time_step / self.cool_time
            self.temperature += (self.final_temp - self.init_temp) * self.integration_time_step / self.cool_time


class TTMPhysics(Physics):
    """``physics.module = "two-temperature-model"`` (physics/two_temperature_model.{h,cc}): electron and phonon temperatures
    driven by Gaussian laser pulses, forward Euler with the solver's step (two_temperature_model.cc:66-95); the thermostat
    follows the electron temperature.  ``records`` holds what the reference writes to ``ttm.tsv`` (:91-94)."""

    def __init__(self, settings: dict, output_steps: int | None = None):
        super().__init__(settings)
        self.phonon_temp = float(settings["InitialTemperature"])
        self.electron_temp = self.phonon_temp
        self.sink_temp = self.phonon_temp
        pulses = settings.get("laserPulses", [])
        self.pulse_width = np.array([float(q["width"]) for q in pulses])
        self.pulse_fluence = np.array([1.152e20 * float(q["fluence"]) for q in pulses])    # pumpPower, two_temperature_model.h:23
        self.pulse_start = np.array([float(q["t_start"]) for q in pulses])
        self.Ce = float(settings.get("Ce", 7.0e2))
        self.Cl = float(settings.get("Cl", 3.0e6))
        self.G = float(settings.get("Gep", 17.0e17))
        self.Gsink = float(settings.get("Gps", 17.0e14))
        self.reversing_field = np.asarray(settings["ReversingField"], dtype=np.float64)
        self.output_steps = int(output_steps if output_steps is not None else settings.get("output_steps", 100))
        self.pump_temp = 0.0
        self.temperature = self.electron_temp if "temperature" not in settings else self.temperature
        self.records = []

    def update(self, solver):
        time, dt = solver.time, solver.step_size
        self.applied_field = self.reversing_field.copy()
        pump = 0.0
        for w, f, t0 in zip(self.pulse_width, self.pulse_fluence, self.pulse_start):
            rel = time - t0
            if 0.0 < rel <= 10 * w:
                pump += f * np.exp(-((rel - 3 * w) / w) ** 2)
        self.pump_temp = pump
        self.electron_temp = self.electron_temp + ((-self.G * (self.electron_temp - self.phonon_temp) + pump) * dt) / (self.Ce * self.electron_temp)
        self.phonon_temp = self.phonon_temp + ((self.G * (self.electron_temp - self.phonon_temp) - self.Gsink * (self.phonon_temp - self.sink_temp)) * dt) / self.Cl
        self.temperature = self.electron_temp
        if solver.iteration % self.output_steps == 0:
            self.records.append((time, self.electron_temp, self.phonon_temp, pump))


def create_physics(settings: dict | None, lattice: Lattice, sim: dict | None = None) -> Physics:
    """Physics::create (core/physics.cc:79-126), the modules on this path"""
    module = str((settings or {}).get("module", "empty")).lower()
    if module == "empty":
        return Physics(settings)
    if module == "pinned_boundaries":
        return PinnedBoundariesPhysics(settings, lattice)
    if module == "field-cool":
        return FieldCoolPhysics(settings, sim)
    if module == "two-temperature-model":
        return TTMPhysics(settings)
    raise RuntimeError("unknown physics module " + module)


class Solver:
    """core/solver.h:15-90"""
    name = "solver"

    def __init__(self):
        self.iteration = 0
        self.time = 0.0
        self.step_size = 1.0
        self.max_steps = 0
        self.min_steps = 0
        self.temperature = 0.0
        self.hamiltonians = []
        self.monitors = []
        self.physics = Physics()

    def register_physics_module(self, p: Physics):   # core/solver.h:47
        self.physics = p
        self.set_temperature(p.temperature)

    def set_temperature(self, T):
        self.temperature = float(T)

    def update_physics_module(self):   # core/solver.cc:85-87, called before notify_monitors and run (core/jams++.cc:334)
        self.physics.update(self)
        if self.physics.temperature != self.temperature:   # update_thermostat re-reads it every step (core/solver.cc:94-97)
            self.set_temperature(self.physics.temperature)

    def is_cuda_solver(self):
        return False

    def is_running(self):
        return self.iteration < self.max_steps

    def register_hamiltonian(self, h: Hamiltonian):
        # the reference sums any number of Hamiltonians (core/solver.cc:43-57).  The fused kernels hold ONE bilinear exchange list, one
        # biquadratic list, one Zeeman field and one applied field -- a second one of those is refused, not dropped -- and up to
        # three uniaxial terms (K1 + K2 + K3 as separate modules: jb_set_uniaxial_term slots 0, 1, 2)
        if isinstance(h, UniaxialAnisotropyHamiltonian):
            n = sum(isinstance(o, UniaxialAnisotropyHamiltonian) for o in self.hamiltonians)
            if n >= len(capi.UNIAXIAL_TERMS):
                raise RuntimeError(f"{self.name}: more than {len(capi.UNIAXIAL_TERMS)} uniaxial hamiltonians (use the reference's solver)")
            h.slot, h.term = n, capi.UNIAXIAL_TERMS[n]
        else:
            for other in self.hamiltonians:
                if other.term == h.term:
                    raise RuntimeError(f"{self.name}: hamiltonians '{other.settings.get('module')}' and '{h.settings.get('module')}' are the "
                                       "same kind of term; the fused solver holds one of each kind (merge them, or use the reference's solver)")
        h.solver = self
        self.hamiltonians.append(h)

    def register_monitor(self, m: Monitor):
        self.monitors.append(m)

    def notify_monitors(self):  # core/solver.cc:110-116
        for m in self.monitors:
            if m.is_updating(self.iteration):
                m.update(self)


class B200HeunLLGSolver(Solver):
    """``module = "llg-heun-b200-gpu"``: drop-in alternative to ``llg-heun-gpu``.

    settings keys (solvers/cuda_llg_heun.cu:23-37, solvers/cpu_llg_heun.cc:17-33): ``t_step``, ``t_max``,
    ``t_min`` in seconds, ``gilbert_prefactor``; extras: ``seed`` (sim.seed), ``device``.
    One instance drives one x-slab; ``comm`` (see :mod:`jams_b200.distributed`) supplies rank/world and the
    halo-handle exchange when the lattice is split over several GPUs."""
    name = "llg-heun-b200-gpu"

    def __init__(self, settings: dict, lattice: Lattice, comm=None):
        super().__init__()
        self.lattice = lattice
        self.comm = comm
        self.rank = comm.rank if comm else 0
        self.n_ranks = comm.world_size if comm else 1
        from .distributed import slab_range
        self.x0, self.nx = slab_range(lattice.dims[0], self.rank, self.n_ranks)
        self.ctx = None
        self._built = False
        self._physics_dirty = False
        self.initialize(settings)

    def is_cuda_solver(self):
        return True

    def initialize(self, settings: dict):
        self.step_size = float(settings["t_step"]) / 1e-12          # ps
        t_max = float(settings["t_max"]) / 1e-12
        t_min = float(settings.get("t_min", 0.0)) / 1e-12
        self.max_steps = int(t_max / self.step_size)
        self.min_steps = int(t_min / self.step_size)
        self.lattice.gilbert_prefactor = bool(settings.get("gilbert_prefactor", self.lattice.gilbert_prefactor))
        self.seed = int(settings.get("seed", 0))
        self.thermostat = LangevinWhiteThermostat(self, self.seed)
        lat = self.lattice
        self.ctx = capi.Context(lat.dims, lat.M, lat.periodic, x_begin=self.x0, nx_local=self.nx,
                                rank=self.rank, n_ranks=self.n_ranks, device=int(settings.get("device", -1)))
        for key, val in settings.get("options", {}).items():
            self.ctx.set_option(key, val)
        self._spins0 = lat.initial_spins(self.x0, self.nx)

    # Hamiltonians are registered after the solver exists (core/jams++.cc:274-288): build lazily
    def _build(self):
        if self._built:
            return
        lat, ctx = self.lattice, self.ctx
        ctx.set_materials(lat.mus(self.x0, self.nx), lat.gyro(self.x0, self.nx), lat.alpha(self.x0, self.nx))
        for h in self.hamiltonians:
            h.attach(ctx, self.x0, self.nx)
        if self.n_ranks > 1:
            self.comm.connect_halos(ctx)
        ctx.import_spins(self._spins0)
        self._built = True

    def set_spins(self, s_aos):
        """globals::s = ... (local slab, N x 3)"""
        self._spins0 = np.ascontiguousarray(s_aos, dtype=np.float64).reshape(-1, 3)
        if self._built:
            if self.n_ranks > 1:
                self.comm.barrier(self.ctx)
            self.ctx.import_spins(self._spins0)
        elif self.hamiltonians:
            self._build()     # Hamiltonians are registered: the device structures can be built now

    def spins(self):
        self._build()
        return self.ctx.export_spins()

    def set_temperature(self, T):
        """physics_module_->temperature() is re-read every step (core/solver.cc:94-97)"""
        self.temperature = float(T)
        self.thermostat.set_temperature(T)

    def run(self, nsteps: int = 1):
        """``nsteps`` Heun steps (one in the reference's main loop, core/jams++.cc:341)"""
        self._build()
        self.ctx.step(nsteps, self.step_size, self.time, self.temperature, self.seed, self.iteration,
                      self.lattice.gilbert_prefactor)
        self.iteration += nsteps
        self.time = self.iteration * self.step_size   # cpu_llg_heun.cc:146-147

    def compute_fields(self):
        """globals::h = sum_k field_k (core/solver.cc:43-57), N x 3 meV"""
        self._build()
        return self.ctx.fields(capi.TERM_TOTAL, self.time)

    def notify_monitors(self):
        self._build()
        super().notify_monitors()

    def reduce_sum(self, arr):
        """all-reduce of monitor partial sums over slabs (no-op on one GPU)"""
        return self.comm.allreduce_sum(arr) if self.comm and self.n_ranks > 1 else arr


class B200RK4LLGSolver(B200HeunLLGSolver):
    """``module = "llg-rk4-b200-gpu"``: drop-in alternative to ``llg-rk4-gpu`` (CudaRK4BaseSolver + CUDALLGRK4Solver,
    solvers/cuda_rk4_base.cu:10-108, solvers/cuda_llg_rk4.cu:17-34).  Same settings keys as the Heun solver
    (cuda_rk4_base.cu:12-29); the shipped example runs this integrator with a ten times larger ``t_step``
    (examples/bloch_domain_wall/bloch_domain_wall.cfg:75-81)."""
    name = "llg-rk4-b200-gpu"

    def run(self, nsteps: int = 1):
        """``nsteps`` RK4 steps (cuda_rk4_base.cu:50-108)"""
        self._build()
        self.ctx.step_rk4(nsteps, self.step_size, self.time, self.temperature, self.seed, self.iteration,
                          self.lattice.gilbert_prefactor)
        self.iteration += nsteps
        self.time = self.iteration * self.step_size   # cuda_rk4_base.cu:105-106


def create_solver(settings: dict, lattice: Lattice, comm=None) -> Solver:
    """Solver::create (core/solver.cc:60-77)"""
    module = str(settings["module"]).lower()
    if module == "llg-heun-b200-gpu":
        return B200HeunLLGSolver(settings, lattice, comm)
    if module == "llg-rk4-b200-gpu":
        return B200RK4LLGSolver(settings, lattice, comm)
    raise RuntimeError("unknown solver " + str(settings["module"]))
