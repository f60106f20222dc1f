"""oracle — CPU checkers for the llg-heun + exchange hot path.  TEST INFRASTRUCTURE ONLY.

Two libraries, one Python face:

* ``libjams_oracle.so``  — self-contained restatement (``oracle/jams_oracle.cpp``), prefix ``jo_``.
* ``_ref/libjams_ref.so`` — the reference's own header-only code compiled in place from
  ``/root/reference/src`` (``oracle/ref_wrap.cpp``), prefix ``jref_``.  Exists only if it was built in a
  container that has the reference tree; it then travels to the GPU box as a prebuilt file.
* ``_ref/libjams_ref_cuda.so`` — the reference's own CUDA kernels and cuSPARSE field path for the same step
  (``oracle/ref_cuda_wrap.cu``, nvcc for sm_100a from the same tree), prefix ``jrc_``: ``RefCudaSim``.  Needs a GPU to run.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``jams_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libjams_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libjams_ref.so")
REF_CUDA_SO = os.path.join(HERE, "_ref", "libjams_ref_cuda.so")

_c_double_p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_c_int_p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (``make -C oracle oracle [ref]``)."""
    targets = ["oracle"] + (["ref", "refcuda"] if ref else [])
    subprocess.run(["make", "-C", HERE, "--no-print-directory"] + targets, check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_ref_cuda() -> bool:
    return os.path.exists(REF_CUDA_SO)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _Lib:
    """ctypes face shared by both libraries (function names differ only by prefix)."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing; run `make -C oracle` (python -c 'import oracle; oracle.build()')")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        p = prefix
        L = self.lib

        def sig(name, restype, *argtypes, optional=False):
            f = getattr(L, p + name, None)
            if f is None:
                if optional:
                    return None      # the reference-header build (oracle/_ref) has no such entry
                raise AttributeError(p + name)
            f.restype, f.argtypes = restype, list(argtypes)
            return f

        self.last_error = sig("last_error", C.c_char_p)
        self.omp_threads = sig("omp_threads", C.c_int)
        self.sim_create = sig("sim_create", C.c_void_p, C.c_int, _c_double_p, _c_double_p, _c_double_p)
        self.sim_destroy = sig("sim_destroy", None, C.c_void_p)
        self.sim_add_exchange = sig("sim_add_exchange", C.c_int, C.c_void_p, C.c_int64, _c_int_p, _c_int_p, _c_double_p, C.c_int)
        self.sim_add_uniaxial = sig("sim_add_uniaxial", C.c_int, C.c_void_p, C.c_int, _c_double_p, _c_double_p)
        self.sim_add_zeeman = sig("sim_add_zeeman", C.c_int, C.c_void_p, _c_double_p, C.c_void_p, C.c_void_p)
        self.sim_add_applied_field = sig("sim_add_applied_field", C.c_int, C.c_void_p, _c_double_p, C.c_int, C.c_double, C.c_double, C.c_double, optional=True)
        self.sim_add_biquadratic = sig("sim_add_biquadratic", C.c_int, C.c_void_p, C.c_int64, _c_int_p, _c_int_p, _c_double_p, C.c_int, optional=True)
        self.sim_exchange_nnz = sig("sim_exchange_nnz", C.c_int64, C.c_void_p, C.c_int)
        self.sim_exchange_csr = sig("sim_exchange_csr", None, C.c_void_p, C.c_int, _c_int_p, _c_int_p, _c_double_p)
        self.sim_set_spins = sig("sim_set_spins", None, C.c_void_p, _c_double_p)
        self.sim_get_spins = sig("sim_get_spins", None, C.c_void_p, _c_double_p)
        self.sim_get_h = sig("sim_get_h", None, C.c_void_p, _c_double_p)
        self.sim_init_solver = sig("sim_init_solver", None, C.c_void_p, C.c_double, C.c_int, C.c_uint64)
        self.sim_get_sigma = sig("sim_get_sigma", None, C.c_void_p, _c_double_p)
        self.sim_set_temperature = sig("sim_set_temperature", None, C.c_void_p, C.c_double)
        self.sim_time = sig("sim_time", C.c_double, C.c_void_p)
        self.sim_run = sig("sim_run", None, C.c_void_p, C.c_int, C.c_void_p)
        self.sim_term_fields = sig("sim_term_fields", None, C.c_void_p, C.c_int, C.c_double, _c_double_p)
        self.sim_term_total_energy = sig("sim_term_total_energy", C.c_double, C.c_void_p, C.c_int, C.c_double)
        self.rotation_matrix_between_vectors = sig("rotation_matrix_between_vectors", None, _c_double_p, _c_double_p, _c_double_p)

    def error(self) -> str:
        return (self.last_error() or b"").decode()


_libs: dict[str, _Lib] = {}


def restatement() -> _Lib:
    if "jo" not in _libs:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = _Lib(ORACLE_SO, "jo_")
        L = lib.lib
        L.jo_expand_template.restype = C.c_int64
        L.jo_expand_template.argtypes = [
            _c_double_p, C.c_int, _c_double_p, _c_int_p, C.c_int, C.c_int, C.c_int64, _c_int_p, _c_int_p,
            _c_double_p, _c_double_p, C.c_int, C.c_int, _c_double_p, _c_double_p, C.c_double, C.c_double, C.c_double,
            C.c_int64, _c_int_p, _c_int_p, _c_int_p, _c_double_p, _c_double_p]
        L.jo_neighbour_list.restype = C.c_int64
        L.jo_neighbour_list.argtypes = [
            _c_int_p, _c_int_p, C.c_int, _c_int_p, C.c_int64, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_double_p,
            C.c_int64, _c_int_p, _c_int_p, _c_int_p, C.c_int, C.POINTER(C.c_int), _c_double_p]
        L.jo_init_bloch_domain_wall.restype = None
        L.jo_init_bloch_domain_wall.argtypes = [C.c_int64, _c_double_p, C.c_double, C.c_double, _c_double_p, _c_double_p, _c_double_p]
        L.jo_magnetisation.restype = None
        L.jo_magnetisation.argtypes = [C.c_int64, _c_double_p, _c_double_p, C.c_int, _c_int_p, _c_double_p]
        L.jo_spin_temperature.restype = C.c_double
        L.jo_spin_temperature.argtypes = [C.c_int64, _c_double_p, _c_double_p]
        L.jo_sim_run_rk4.restype = None
        L.jo_sim_run_rk4.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.jo_sim_term_energies.restype = None
        L.jo_sim_term_energies.argtypes = [C.c_void_p, C.c_int, C.c_double, _c_double_p]
        _libs["jo"] = lib
    return _libs["jo"]


def reference() -> _Lib:
    """The reference-header build (raises FileNotFoundError where it was never built)."""
    if "jref" not in _libs:
        lib = _Lib(REF_SO, "jref_")
        L = lib.lib
        L.jref_interaction_list.restype = C.c_int
        L.jref_interaction_list.argtypes = [C.c_int64, _c_int_p, _c_int_p, _c_double_p, _c_int_p, _c_int_p, _c_int_p,
                                            C.POINTER(C.c_int), _c_double_p]
        L.jref_unit_vector.restype = None
        L.jref_unit_vector.argtypes = [_c_double_p, _c_double_p]
        L.jref_llg_rhs.restype = None
        L.jref_llg_rhs.argtypes = [_c_double_p, _c_double_p, C.c_double, C.c_double, _c_double_p]
        L.jref_approximately_equal.restype = C.c_int
        L.jref_approximately_equal.argtypes = [C.c_double, C.c_double, C.c_double]
        L.jref_have_pcg.restype = C.c_int
        _libs["jref"] = lib
    return _libs["jref"]


REF_CUDA_SYMBOLS = ("jrc_last_error", "jrc_device_count", "jrc_sim_create", "jrc_sim_destroy", "jrc_sim_add_exchange",
                    "jrc_sim_add_uniaxial", "jrc_sim_add_zeeman", "jrc_sim_exchange_nnz", "jrc_sim_set_spins", "jrc_sim_get_spins",
                    "jrc_sim_get_h", "jrc_sim_init_solver", "jrc_sim_run_heun", "jrc_sim_run_rk4", "jrc_sim_time_heun", "jrc_sim_time_rk4",
                    "jrc_biquadratic_field", "jrc_pin_region", "jrc_reduce", "jrc_multiarray_contract")


def reference_cuda():
    """ctypes handle of the reference's CUDA path (raises FileNotFoundError where it was never built)."""
    if "jrc" not in _libs:
        if not os.path.exists(REF_CUDA_SO):
            raise FileNotFoundError(f"{REF_CUDA_SO} missing; `make -C oracle refcuda` in a container that has /root/reference")
        L = C.CDLL(REF_CUDA_SO)
        L.jrc_last_error.restype = C.c_char_p
        L.jrc_device_count.restype = C.c_int
        L.jrc_sim_create.restype = C.c_void_p
        L.jrc_sim_create.argtypes = [C.c_int, _c_double_p, _c_double_p, _c_double_p]
        L.jrc_sim_destroy.restype = None
        L.jrc_sim_destroy.argtypes = [C.c_void_p]
        L.jrc_sim_add_exchange.restype = C.c_int
        L.jrc_sim_add_exchange.argtypes = [C.c_void_p, C.c_long, _c_int_p, _c_int_p, _c_double_p, C.c_int]
        L.jrc_sim_add_uniaxial.restype = C.c_int
        L.jrc_sim_add_uniaxial.argtypes = [C.c_void_p, C.c_int, _c_double_p, _c_double_p]
        L.jrc_sim_add_zeeman.restype = C.c_int
        L.jrc_sim_add_zeeman.argtypes = [C.c_void_p, _c_double_p, C.c_void_p, C.c_void_p]
        L.jrc_sim_exchange_nnz.restype = C.c_long
        L.jrc_sim_exchange_nnz.argtypes = [C.c_void_p, C.c_int]
        for name in ("jrc_sim_set_spins", "jrc_sim_get_spins", "jrc_sim_get_h"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p, _c_double_p]
        L.jrc_sim_init_solver.restype = C.c_int
        L.jrc_sim_init_solver.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_ulonglong]
        for name in ("jrc_sim_run_heun", "jrc_sim_run_rk4"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for name in ("jrc_sim_time_heun", "jrc_sim_time_rk4"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.jrc_biquadratic_field.restype = C.c_int
        L.jrc_biquadratic_field.argtypes = [C.c_int, C.c_long, _c_int_p, _c_int_p, _c_double_p, _c_double_p, _c_double_p]
        L.jrc_pin_region.restype = C.c_int
        L.jrc_pin_region.argtypes = [C.c_int, _c_double_p, _c_double_p, C.c_int, _c_int_p, _c_double_p, _c_double_p]
        L.jrc_reduce.restype = C.c_int
        L.jrc_reduce.argtypes = [C.c_int, C.c_int, _c_double_p, C.c_void_p, C.c_int, C.c_void_p, _c_double_p]
        _libs["jrc"] = L
    return _libs["jrc"]


class RefCudaSim:
    """The reference's own CUDA step (llg-heun-gpu / llg-rk4-gpu: cuSPARSE field + its kernels) on the current device, with the
    method names of ``CpuSim`` so that tests/helpers.build_cpu_sim can assemble either."""

    def __init__(self, mus, gyro, alpha):
        self.L = reference_cuda()
        if self.L.jrc_device_count() < 1:
            raise RuntimeError("the reference's CUDA path needs a GPU")
        self.N = len(mus)
        self._arrays = (_f64(mus), _f64(gyro), _f64(alpha))
        self.h = self.L.jrc_sim_create(self.N, *self._arrays)
        if not self.h:
            raise RuntimeError(self.error())
        self.n_terms = 0
        self.T = 0.0
        self._solver = None

    def error(self):
        return (self.L.jrc_last_error() or b"").decode()

    def close(self):
        if self.h:
            self.L.jrc_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.error())

    def add_exchange(self, i, j, J9_per_pair, check_symmetric=True):
        self._check(self.L.jrc_sim_add_exchange(self.h, len(i), _i32(i), _i32(j), _f64(J9_per_pair, (-1,)), int(check_symmetric)))
        self.n_terms += 1
        return self.n_terms - 1

    def add_uniaxial(self, power, magnitude, axis):
        self._check(self.L.jrc_sim_add_uniaxial(self.h, int(power), _f64(magnitude), _f64(axis, (-1,))))
        self.n_terms += 1
        return self.n_terms - 1

    def add_zeeman(self, dc, ac=None, omega=None):
        dc = _f64(dc, (-1,))
        if ac is not None:
            ac = _f64(ac, (-1,)); omega = _f64(omega)
            rc = self.L.jrc_sim_add_zeeman(self.h, dc, ac.ctypes.data, omega.ctypes.data)
        else:
            rc = self.L.jrc_sim_add_zeeman(self.h, dc, None, None)
        self._check(rc)
        self.n_terms += 1
        return self.n_terms - 1

    def exchange_nnz(self, term=0):
        return int(self.L.jrc_sim_exchange_nnz(self.h, term))

    def set_spins(self, s):
        self._check(self.L.jrc_sim_set_spins(self.h, _f64(s, (-1,))))

    def get_spins(self):
        s = np.zeros(3 * self.N)
        self._check(self.L.jrc_sim_get_spins(self.h, s))
        return s.reshape(-1, 3)

    def get_h(self):
        h = np.zeros(3 * self.N)
        self._check(self.L.jrc_sim_get_h(self.h, h))
        return h.reshape(-1, 3)

    def init_solver(self, dt_ps, gilbert_prefactor=False, seed=1):
        # sigma_(i, j) as CudaThermostatClassical's constructor fills it (cuda_thermostat_classical.cc:34-44 = the CPU solver's
        # formula, cpu_llg_heun.cc:33-42): taken from the restatement, which test_oracle_cpu.py pins to the reference bit for bit
        cpu = CpuSim(*self._arrays)
        cpu.init_solver(dt_ps, gilbert_prefactor, seed)
        self._sigma = np.ascontiguousarray(np.repeat(cpu.sigma()[:, None], 3, axis=1))
        cpu.close()
        self._solver = (float(dt_ps), int(seed))
        self._check(self.L.jrc_sim_init_solver(self.h, float(dt_ps), self._sigma.ctypes.data, float(self.T), int(seed)))

    def set_temperature(self, T):
        self.T = float(T)
        if self._solver is not None:
            self._check(self.L.jrc_sim_init_solver(self.h, self._solver[0], self._sigma.ctypes.data, self.T, self._solver[1]))

    def _run(self, fn, nsteps, normals):
        if normals is not None:
            normals = _f64(normals, (-1,))
            assert normals.size == nsteps * 3 * self.N
            self._check(fn(self.h, int(nsteps), normals.ctypes.data))
        else:
            self._check(fn(self.h, int(nsteps), None))

    def run(self, nsteps=1, normals=None):
        """CUDAHeunLLGSolver::run; ``normals`` = None draws them with curand like the reference"""
        self._run(self.L.jrc_sim_run_heun, nsteps, normals)

    def run_rk4(self, nsteps=1, normals=None):
        """CudaRK4BaseSolver::run with CUDALLGRK4Solver's function kernel"""
        self._run(self.L.jrc_sim_run_rk4, nsteps, normals)

    def time_heun(self, steps, warmup, rk4=False):
        """milliseconds per step of the reference's CUDA Heun (RK4) step (CUDA events around ``steps`` steps)"""
        ms = (self.L.jrc_sim_time_rk4 if rk4 else self.L.jrc_sim_time_heun)(self.h, int(steps), int(warmup))
        if ms < 0:
            raise RuntimeError(self.error())
        return ms


def ref_cuda_multiarray_contract(product_lib, product_ctx, s_aos, between):
    """the JAMS adapter's import -> steps -> export through the reference's own MultiArray (SyncedMemory in CUDA mode).
    ``product_lib`` / ``product_ctx``: the ctypes library and context handle of the product (jb_import_spins / jb_export_spins are handed
    over as function pointers); ``between(ctx_handle)`` runs the steps.  Returns (spins a monitor reads after the export, the host
    copy as it stood before the export)."""
    L = reference_cuda()
    IMPORT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
    EXPORT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
    BETWEEN = C.CFUNCTYPE(None, C.c_void_p)
    imp = C.cast(product_lib.jb_import_spins, IMPORT)
    exp = C.cast(product_lib.jb_export_spins, EXPORT)
    cb = BETWEEN(lambda h: between(h))
    s = _f64(s_aos, (-1,))
    out, before = np.zeros_like(s), np.zeros_like(s)
    L.jrc_multiarray_contract.restype = C.c_int
    L.jrc_multiarray_contract.argtypes = [C.c_int, _c_double_p, _c_double_p, _c_double_p, IMPORT, EXPORT, C.c_void_p, BETWEEN]
    if L.jrc_multiarray_contract(s.size // 3, s, out, before, imp, exp, product_ctx, cb) != 0:
        raise RuntimeError((L.jrc_last_error() or b"").decode())
    return out.reshape(-1, 3), before.reshape(-1, 3)


def ref_cuda_biquadratic_field(n, i, j, B, s_aos):
    h = np.zeros(3 * n)
    L = reference_cuda()
    if L.jrc_biquadratic_field(int(n), len(i), _i32(i), _i32(j), _f64(B), _f64(s_aos, (-1,)), h) != 0:
        raise RuntimeError((L.jrc_last_error() or b"").decode())
    return h.reshape(-1, 3)


def ref_cuda_pin_region(s_aos, mus, indices, target):
    """PinnedBoundariesPhysics::update's CUDA branch on one region: returns (rotated spins, region moment before the rotation)"""
    L = reference_cuda()
    s = _f64(s_aos, (-1,)).copy()
    mag = np.zeros(3)
    idx = _i32(indices)
    if L.jrc_pin_region(len(mus), s, _f64(mus), len(idx), idx, _f64(target, (3,)), mag) != 0:
        raise RuntimeError((L.jrc_last_error() or b"").decode())
    return s.reshape(-1, 3), mag


def ref_cuda_reduce(kind, s_aos, mus=None, indices=None):
    L = reference_cuda()
    out = np.zeros(3)
    s = _f64(s_aos, (-1,))
    m = _f64(mus) if mus is not None else None
    idx = _i32(indices) if indices is not None else None
    rc = L.jrc_reduce(int(kind), s.size // 3, s, m.ctypes.data if m is not None else None, len(idx) if idx is not None else 0,
                      idx.ctypes.data if idx is not None else None, out)
    if rc != 0:
        raise RuntimeError((L.jrc_last_error() or b"").decode())
    return out


# ------------------------------------------------------------------------------------------------
# lattice / interaction construction (restatement only; the reference's versions need spglib)
# ------------------------------------------------------------------------------------------------
def cubic_point_group():
    """The 48 signed permutation matrices of O_h with zero translation.

    Stands in for the spglib operation list the reference uses (core/lattice.cc:942-967) for the
    cubic cells of the BASELINE configs."""
    rots = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            R = np.zeros((3, 3))
            for r in range(3):
                R[r, perm[r]] = signs[r]
            rots.append(R.reshape(9))
    rots = np.array(rots)
    return rots, np.zeros((len(rots), 3))


def site_index(dims, M, i, j, k, m):
    return ((i * dims[1] + j) * dims[2] + k) * M + m


def expand_template(cell, motif_frac, motif_type, interactions, *, fmt="jams", frac_coords=False, use_symops=True,
                    symops=None, energy_cutoff=0.0, radius_cutoff=100.0, distance_tolerance=1e-4):
    """``post_process_interactions`` (core/interactions.cc:292-347).

    interactions: list of (type_i, type_j, r[3], J) with J a scalar or 9 numbers, energies in input units."""
    lib = restatement()
    motif_frac = _f64(motif_frac, (-1, 3))
    M = motif_frac.shape[0]
    n_in = len(interactions)
    ti = _i32([it[0] for it in interactions])
    tj = _i32([it[1] for it in interactions])
    r = _f64([it[2] for it in interactions], (n_in, 3))
    J9 = np.zeros((n_in, 9))
    for n, it in enumerate(interactions):
        J = np.asarray(it[3], dtype=np.float64)
        J9[n] = (J * np.eye(3)).reshape(9) if J.ndim == 0 else J.reshape(9)
    if symops is None:
        symops = cubic_point_group()
    rot, trans = _f64(symops[0], (-1, 9)), _f64(symops[1], (-1, 3))
    cap = max(64, n_in * M * (len(rot) + 1))
    mi = np.zeros(cap, np.int32); mj = np.zeros(cap, np.int32); T = np.zeros((cap, 3), np.int32)
    oJ = np.zeros((cap, 9)); orr = np.zeros((cap, 3))
    n = lib.lib.jo_expand_template(_f64(cell, (9,)), M, motif_frac, _i32(motif_type), 0 if fmt == "jams" else 1,
                                   int(frac_coords), n_in, ti, tj, r, J9, int(use_symops), len(rot), rot, trans,
                                   float(energy_cutoff), float(radius_cutoff), float(distance_tolerance),
                                   cap, mi, mj, T, oJ, orr)
    if n < 0:
        raise RuntimeError(lib.error())
    return {"mi": mi[:n].copy(), "mj": mj[:n].copy(), "T": T[:n].copy(), "J9": oJ[:n].copy(), "r": orr[:n].copy()}


def neighbour_list(dims, periodic, M, site_type, template, entry_type_i=None, entry_type_j=None, motif_type=None):
    """``neighbour_list_from_interactions`` (core/interactions.cc:349-395) in InteractionList storage order.

    Returns (i, j, value_id, values9)."""
    lib = restatement()
    dims = _i32(dims); periodic = _i32([int(bool(p)) for p in periodic])
    n_t = len(template["mi"])
    if entry_type_i is None:
        mt = _i32(motif_type if motif_type is not None else np.zeros(M, np.int32))
        entry_type_i = mt[template["mi"]]
        entry_type_j = mt[template["mj"]]
    N = int(np.prod(dims)) * M
    cap = N * max(n_t, 1)
    oi = np.zeros(cap, np.int32); oj = np.zeros(cap, np.int32); ov = np.zeros(cap, np.int32)
    vals = np.zeros((max(n_t, 1), 9)); nv = C.c_int(0)
    n = lib.lib.jo_neighbour_list(dims, periodic, M, _i32(site_type), n_t, _i32(template["mi"]), _i32(template["mj"]),
                                  _i32(template["T"]).reshape(-1), _i32(entry_type_i), _i32(entry_type_j),
                                  _f64(template["J9"], (-1,)), cap, oi, oj, ov, vals.shape[0], C.byref(nv), vals)
    if n < 0:
        raise RuntimeError(lib.error())
    return oi[:n].copy(), oj[:n].copy(), ov[:n].copy(), vals[:nv.value].copy()


def reference_interaction_list(i, j, J9_per_pair):
    """Feed pairs (in generation order) through the reference's jams::InteractionList."""
    lib = reference()
    n = len(i)
    oi = np.zeros(n, np.int32); oj = np.zeros(n, np.int32); ov = np.zeros(n, np.int32)
    vals = np.zeros((max(n, 1), 9)); nv = C.c_int(0)
    rc = lib.lib.jref_interaction_list(n, _i32(i), _i32(j), _f64(J9_per_pair, (-1,)), oi, oj, ov, C.byref(nv), vals)
    if rc != 0:
        raise RuntimeError(lib.error())
    return oi, oj, ov, vals[:nv.value].copy()


def init_bloch_domain_wall(positions, s_aos, width, center, normal=(1, 0, 0), domain=(0, 0, 1)):
    lib = restatement()
    s = _f64(s_aos, (-1, 3)).copy()
    lib.lib.jo_init_bloch_domain_wall(len(s), _f64(positions, (-1, 3)), float(width), float(center),
                                      _f64(normal, (3,)), _f64(domain, (3,)), s)
    return s


def magnetisation(s_aos, mus, group_of_spin=None, n_groups=1):
    lib = restatement()
    s = _f64(s_aos, (-1, 3))
    g = _i32(group_of_spin if group_of_spin is not None else np.zeros(len(s), np.int32))
    out = np.zeros((n_groups, 4))
    lib.lib.jo_magnetisation(len(s), s, _f64(mus), n_groups, g, out)
    return out


def pin_region(s_aos, mus, indices, target):
    """PinnedBoundariesPhysics::update for one boundary on the host path (physics/pinned_boundaries.cc:41-44):
    mag = jams::sum_spins_moments(s, mus, indices) (helpers/spinops.cc:55-67: plain loop over the indices),
    R = rotation_matrix_between_vectors(mag, target) (containers/mat3.h:334-366, the oracle's restatement),
    jams::rotate_spins(s, R, indices) (helpers/spinops.cc: s_i <- R s_i).  Returns the new N x 3 array."""
    s = np.array(_f64(s_aos, (-1, 3)), copy=True)
    mus = _f64(mus)
    mag = np.zeros(3)
    for i in indices:            # same accumulation order as the reference loop
        mag += mus[i] * s[i]
    R = np.zeros(9)
    restatement().rotation_matrix_between_vectors(np.ascontiguousarray(mag), _f64(target), R)
    R = R.reshape(3, 3)
    for i in indices:
        v = s[i]
        s[i] = [R[0, 0] * v[0] + R[0, 1] * v[1] + R[0, 2] * v[2], R[1, 0] * v[0] + R[1, 1] * v[1] + R[1, 2] * v[2],
                R[2, 0] * v[0] + R[2, 1] * v[1] + R[2, 2] * v[2]]
    return s


def spin_temperature(s_aos, h_aos):
    lib = restatement()
    s = _f64(s_aos, (-1, 3))
    return lib.lib.jo_spin_temperature(len(s), s, _f64(h_aos, (-1, 3)))


class CpuSim:
    """llg-heun-cpu on one of the two CPU libraries (``which`` = "restatement" | "reference")."""

    def __init__(self, mus, gyro, alpha, which="restatement"):
        self.L = restatement() if which == "restatement" else reference()
        self.which = which
        self.N = len(mus)
        self.h = self.L.sim_create(self.N, _f64(mus), _f64(gyro), _f64(alpha))
        self.n_terms = 0

    def close(self):
        if self.h:
            self.L.sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.error())

    def add_exchange(self, i, j, J9_per_pair, check_symmetric=True):
        self._check(self.L.sim_add_exchange(self.h, len(i), _i32(i), _i32(j), _f64(J9_per_pair, (-1,)), int(check_symmetric)))
        self.n_terms += 1
        return self.n_terms - 1

    def add_biquadratic(self, i, j, B_per_pair, check_symmetric=True):
        """CudaBiquadraticExchangeHamiltonian (hamiltonian/cuda_biquadratic_exchange.cu): scalar N x N matrix; restatement only"""
        if getattr(self.L, "sim_add_biquadratic", None) is None:
            raise RuntimeError("this oracle build has no biquadratic-exchange term")
        self._check(self.L.sim_add_biquadratic(self.h, len(i), _i32(i), _i32(j), _f64(B_per_pair, (-1,)), int(check_symmetric)))
        self.n_terms += 1
        return self.n_terms - 1

    def add_uniaxial(self, power, magnitude, axis):
        self._check(self.L.sim_add_uniaxial(self.h, int(power), _f64(magnitude), _f64(axis, (-1,))))
        self.n_terms += 1
        return self.n_terms - 1

    def add_zeeman(self, dc, ac=None, omega=None):
        dc = _f64(dc, (-1,))
        if ac is not None:
            ac = _f64(ac, (-1,)); omega = _f64(omega)
            rc = self.L.sim_add_zeeman(self.h, dc, ac.ctypes.data, omega.ctypes.data)
        else:
            rc = self.L.sim_add_zeeman(self.h, dc, None, None)
        self._check(rc)
        self.n_terms += 1
        return self.n_terms - 1

    def add_applied_field(self, B, kind="static", time_center_ps=0.0, freq_bandwidth_THz=0.0, freq_center_THz=0.0):
        """AppliedFieldHamiltonian (hamiltonian/applied_field.cc): mu_i B g(t); restatement only"""
        if getattr(self.L, "sim_add_applied_field", None) is None:
            raise RuntimeError("this oracle build has no applied-field term")
        rc = self.L.sim_add_applied_field(self.h, _f64(B, (3,)), {"static": 0, "sinc": 1, "sinc-cos": 2}[kind], float(time_center_ps),
                                          float(freq_bandwidth_THz), float(freq_center_THz))
        self._check(rc)
        self.n_terms += 1
        return self.n_terms - 1

    def exchange_csr(self, term=0):
        nnz = self.L.sim_exchange_nnz(self.h, term)
        row = np.zeros(3 * self.N + 1, np.int32); col = np.zeros(nnz, np.int32); val = np.zeros(nnz)
        self.L.sim_exchange_csr(self.h, term, row, col, val)
        return row, col, val

    def set_spins(self, s):
        self.L.sim_set_spins(self.h, _f64(s, (-1,)))

    def get_spins(self):
        s = np.zeros(3 * self.N)
        self.L.sim_get_spins(self.h, s)
        return s.reshape(-1, 3)

    def get_h(self):
        s = np.zeros(3 * self.N)
        self.L.sim_get_h(self.h, s)
        return s.reshape(-1, 3)

    def init_solver(self, dt_ps, gilbert_prefactor=False, seed=1):
        self.L.sim_init_solver(self.h, float(dt_ps), int(gilbert_prefactor), int(seed))

    def sigma(self):
        s = np.zeros(self.N)
        self.L.sim_get_sigma(self.h, s)
        return s

    def set_temperature(self, T):
        self.L.sim_set_temperature(self.h, float(T))

    def time(self):
        return self.L.sim_time(self.h)

    def run(self, nsteps=1, normals=None):
        if normals is not None:
            normals = _f64(normals, (-1,))
            assert normals.size == nsteps * 3 * self.N
            self.L.sim_run(self.h, int(nsteps), normals.ctypes.data)
        else:
            self.L.sim_run(self.h, int(nsteps), None)

    def run_rk4(self, nsteps=1, normals=None):
        """CudaRK4BaseSolver::run restated on the host (restatement library only: the reference has no CPU RK4)"""
        fn = getattr(self.L.lib, self.L.prefix + "sim_run_rk4", None)
        if fn is None:
            raise RuntimeError("run_rk4 exists in the restatement library only")
        if normals is not None:
            normals = _f64(normals, (-1,))
            assert normals.size == nsteps * 3 * self.N
            fn(C.c_void_p(self.h), int(nsteps), C.c_void_p(normals.ctypes.data))
        else:
            fn(C.c_void_p(self.h), int(nsteps), None)

    def term_fields(self, term, time=0.0):
        f = np.zeros(3 * self.N)
        self.L.sim_term_fields(self.h, int(term), float(time), f)
        return f.reshape(-1, 3)

    def term_total_energy(self, term, time=0.0):
        return self.L.sim_term_total_energy(self.h, int(term), float(time))

    def term_energies(self, term, time=0.0):
        assert self.which == "restatement"
        e = np.zeros(self.N)
        self.L.lib.jo_sim_term_energies(self.h, int(term), float(time), e)
        return e
