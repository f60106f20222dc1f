"""ctypes binding of the C ABI in include/jams_b200.h (libjams_b200.so, built in-tree by
jams_b200/csrc/Makefile).  This is the only way Python reaches the kernels; there is no fallback:
a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjams_b200.so")

JB_OK, JB_ERR_INVALID, JB_ERR_CUDA, JB_ERR_UNSUPPORTED, JB_ERR_PEER = range(5)
TERM_EXCHANGE, TERM_UNIAXIAL, TERM_ZEEMAN, TERM_APPLIED, TERM_TOTAL, TERM_BIQUADRATIC, TERM_UNIAXIAL_2, TERM_UNIAXIAL_3 = range(8)
UNIAXIAL_TERMS = (TERM_UNIAXIAL, TERM_UNIAXIAL_2, TERM_UNIAXIAL_3)   # jb_set_uniaxial_term slots 0, 1, 2
HALO_HANDLE_BYTES = 256


class JamsB200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"jams_b200 status {status}: {message}")
        self.status = status


class LatticeDesc(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("num_motif", C.c_int32), ("periodic", C.c_int32 * 3),
                ("x_begin", C.c_int32), ("nx_local", C.c_int32), ("rank", C.c_int32), ("n_ranks", C.c_int32),
                ("device", C.c_int32)]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

# every symbol include/jams_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "jb_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(LatticeDesc)]),
    "jb_destroy": (None, [C.c_void_p]),
    "jb_last_error": (C.c_char_p, [C.c_void_p]),
    "jb_abi_version": (C.c_int, []),
    "jb_set_materials": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "jb_set_exchange_template": (C.c_int, [C.c_void_p, C.c_int32, _ip, _ip, _ip, _dp]),
    "jb_set_biquadratic_template": (C.c_int, [C.c_void_p, C.c_int32, _ip, _ip, _ip, _dp]),
    "jb_set_exchange_pairs": (C.c_int, [C.c_void_p, C.c_int64, _ip, _ip, _ip, C.c_int32, _dp]),
    "jb_detect_exchange_template": (C.c_int, [C.POINTER(LatticeDesc), C.c_int64, _ip, _ip, _ip, C.c_int32, _dp, C.c_int32,
                                               C.POINTER(C.c_int32), _ip, _ip, _ip, _dp]),
    "jb_set_uniaxial": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "jb_set_uniaxial_term": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "jb_timed_steps": (C.c_int64, [C.c_void_p]),
    "jb_set_zeeman": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jb_set_applied_field": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "jb_set_applied_field_pulse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double]),
    "jb_import_spins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "jb_export_spins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "jb_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int32]),
    "jb_set_region": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "jb_region_moment": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "jb_rotate_region": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "jb_step_rk4": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int32]),
    "jb_noise": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "jb_fields": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32]),
    "jb_energies": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "jb_magnetisation": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, _dp]),
    "jb_set_magnetisation_groups": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "jb_halo_export_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jb_halo_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "jb_kernel_launches": (C.c_int64, [C.c_void_p]),
    "jb_stage_kernel": (C.c_int, [C.c_void_p]),
    "jb_last_step_kernel_ms": (C.c_int, [C.c_void_p, _dp]),
    "jb_synchronize": (C.c_int, [C.c_void_p]),
    "jb_stream": (C.c_void_p, [C.c_void_p]),
    "jb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "jb_plan_work_items": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _ip, _ip]),
    "jb_last_stage_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
}

_lib = None


def load():
    """Load libjams_b200.so and bind every declared symbol.  Raises if the extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def detect_exchange_template(dims, num_motif, periodic, i, j, value_id, values9, x_begin=0, nx_local=None, capacity=1024):
    """jb_detect_exchange_template: (mi, mj, T, J9) of a translation-invariant neighbour list, or None.  Host only."""
    lib = load()
    d = LatticeDesc()
    d.dims[:] = [int(v) for v in dims]
    d.num_motif = int(num_motif)
    d.periodic[:] = [int(bool(v)) for v in periodic]
    d.x_begin = int(x_begin)
    d.nx_local = int(dims[0] if nx_local is None else nx_local)
    d.rank, d.n_ranks, d.device = 0, 1, -1
    i = np.ascontiguousarray(i, np.int32); j = np.ascontiguousarray(j, np.int32)
    v = np.ascontiguousarray(value_id, np.int32); J = _f64(values9).reshape(-1)
    mi = np.zeros(capacity, np.int32); mj = np.zeros(capacity, np.int32); T = np.zeros(3 * capacity, np.int32)
    J9 = np.zeros(9 * capacity)
    n = C.c_int32(-1)
    rc = lib.jb_detect_exchange_template(C.byref(d), i.size, i, j, v, J.size // 9, J, capacity, C.byref(n), mi, mj, T, J9)
    if rc != JB_OK:
        raise JamsB200Error(rc, "jb_detect_exchange_template: invalid arguments")
    if n.value < 0:
        return None
    k = n.value
    return dict(mi=mi[:k].copy(), mj=mj[:k].copy(), T=T[:3 * k].reshape(k, 3).copy(), J9=J9[:9 * k].reshape(k, 9).copy())


def plan_work_items(nx_local, ghost_x, n_columns, n_ctas):
    """jb_plan_work_items: the stage kernel's x-chunk plan in queue order, as a list of (x0, xc).  Host only."""
    lib = load()
    x0 = np.zeros(160, np.int32); xc = np.zeros(160, np.int32)
    n = C.c_int32(0)
    rc = lib.jb_plan_work_items(int(nx_local), int(ghost_x), int(n_columns), int(n_ctas), 160, C.byref(n), x0, xc)
    if rc != JB_OK:
        raise JamsB200Error(rc, "jb_plan_work_items: invalid arguments")
    return [(int(x0[k]), int(xc[k])) for k in range(n.value)]


class Context:
    """One jb_ctx: the spins of one x-slab on one GPU."""

    def __init__(self, dims, num_motif=1, periodic=(True, True, True), x_begin=0, nx_local=None, rank=0, n_ranks=1, device=-1):
        self.lib = load()
        d = LatticeDesc()
        d.dims[:] = [int(v) for v in dims]
        d.num_motif = int(num_motif)
        d.periodic[:] = [int(bool(v)) for v in periodic]
        d.x_begin = int(x_begin)
        d.nx_local = int(dims[0] if nx_local is None else nx_local)
        d.rank, d.n_ranks, d.device = int(rank), int(n_ranks), int(device)
        self.desc = d
        self.N = d.nx_local * d.dims[1] * d.dims[2] * d.num_motif
        h = C.c_void_p()
        rc = self.lib.jb_create(C.byref(h), C.byref(d))
        if rc != JB_OK:
            raise JamsB200Error(rc, (self.lib.jb_last_error(None) or b"").decode())
        self.h = h

    # ---- plumbing
    def _ck(self, rc):
        if rc != JB_OK:
            raise JamsB200Error(rc, (self.lib.jb_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.jb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- parameters
    def set_materials(self, mus, gyro, alpha):
        mus, gyro, alpha = _f64(mus), _f64(gyro), _f64(alpha)
        assert mus.size == gyro.size == alpha.size == self.N
        self._ck(self.lib.jb_set_materials(self.h, mus, gyro, alpha))

    def set_exchange_template(self, mi, mj, T, J9):
        mi = np.ascontiguousarray(mi, np.int32); mj = np.ascontiguousarray(mj, np.int32)
        T = np.ascontiguousarray(T, np.int32).reshape(-1); J9 = _f64(J9).reshape(-1)
        assert T.size == 3 * mi.size and J9.size == 9 * mi.size
        self._ck(self.lib.jb_set_exchange_template(self.h, mi.size, mi, mj, T, J9))

    def set_biquadratic_template(self, mi, mj, T, B):
        mi = np.ascontiguousarray(mi, np.int32); mj = np.ascontiguousarray(mj, np.int32)
        T = np.ascontiguousarray(T, np.int32).reshape(-1); B = _f64(B).reshape(-1)
        assert T.size == 3 * mi.size and B.size == mi.size
        self._ck(self.lib.jb_set_biquadratic_template(self.h, mi.size, mi, mj, T, B))

    def set_exchange_pairs(self, i, j, value_id, values9):
        i = np.ascontiguousarray(i, np.int32); j = np.ascontiguousarray(j, np.int32)
        v = np.ascontiguousarray(value_id, np.int32); J = _f64(values9).reshape(-1)
        self._ck(self.lib.jb_set_exchange_pairs(self.h, i.size, i, j, v, J.size // 9, J))

    def set_uniaxial(self, power, magnitude, axis, slot=0):
        """slot 0 = jb_set_uniaxial; slots 1, 2: a second / third uniaxial Hamiltonian (jb_set_uniaxial_term)"""
        if not power:
            self._ck(self.lib.jb_set_uniaxial_term(self.h, int(slot), 0, None, None)); return
        K, a = _f64(magnitude), _f64(axis).reshape(-1)
        assert K.size == self.N and a.size == 3 * self.N
        self._ck(self.lib.jb_set_uniaxial_term(self.h, int(slot), int(power), _ptr(K), _ptr(a)))

    def set_zeeman(self, dc, ac=None, omega=None):
        dc = None if dc is None else _f64(dc).reshape(-1)
        ac = None if ac is None else _f64(ac).reshape(-1)
        omega = None if omega is None else _f64(omega)
        self._ck(self.lib.jb_set_zeeman(self.h, _ptr(dc), _ptr(ac), _ptr(omega)))

    def set_applied_field(self, B, enable=True):
        B = _f64(B if B is not None else [0, 0, 0])
        self._ck(self.lib.jb_set_applied_field(self.h, _ptr(B), int(enable)))

    FIELD_TYPES = {"static": 0, "sinc": 1, "sinc-cos": 2}

    def set_applied_field_pulse(self, B, kind, time_center_ps=0.0, freq_bandwidth_THz=0.0, freq_center_THz=0.0):
        """B(t) = B g(t): static, sinc or sinc-cos amplitude (hamiltonian/applied_field.cc:10-82)"""
        B = _f64(B)
        self._ck(self.lib.jb_set_applied_field_pulse(self.h, _ptr(B), self.FIELD_TYPES[kind], float(time_center_ps),
                                                     float(freq_bandwidth_THz), float(freq_center_THz)))

    def set_option(self, key, value):
        self._ck(self.lib.jb_set_option(self.h, key.encode(), int(value)))

    # ---- state
    def import_spins(self, s_aos):
        s = _f64(s_aos).reshape(-1)
        assert s.size == 3 * self.N
        self._ck(self.lib.jb_import_spins(self.h, _ptr(s), 0))

    def import_spins_ptr(self, ptr, on_device):
        self._ck(self.lib.jb_import_spins(self.h, C.c_void_p(ptr), int(on_device)))

    def export_spins(self, out=None):
        out = np.empty((self.N, 3)) if out is None else out
        self._ck(self.lib.jb_export_spins(self.h, _ptr(out), 0))
        return out

    def export_spins_ptr(self, ptr, on_device):
        self._ck(self.lib.jb_export_spins(self.h, C.c_void_p(ptr), int(on_device)))

    # ---- hot path
    def step(self, nsteps, dt_ps, time_ps=0.0, temperature=0.0, seed=0, first_step=0, gilbert_prefactor=False):
        self._ck(self.lib.jb_step(self.h, int(nsteps), float(dt_ps), float(time_ps), float(temperature),
                                  int(seed), int(first_step), int(gilbert_prefactor)))

    # ---- physics hooks (PinnedBoundariesPhysics)
    def set_region(self, region, site_index):
        idx = np.ascontiguousarray(site_index, dtype=np.int32)
        self._ck(self.lib.jb_set_region(self.h, int(region), int(idx.size), _ptr(idx)))

    def region_moment(self, region):
        out = np.zeros(4)
        self._ck(self.lib.jb_region_moment(self.h, int(region), _ptr(out)))
        return out

    def rotate_region(self, region, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        self._ck(self.lib.jb_rotate_region(self.h, int(region), _ptr(R)))

    def step_rk4(self, nsteps, dt_ps, time_ps=0.0, temperature=0.0, seed=0, first_step=0, gilbert_prefactor=False):
        self._ck(self.lib.jb_step_rk4(self.h, int(nsteps), float(dt_ps), float(time_ps), float(temperature),
                                      int(seed), int(first_step), int(gilbert_prefactor)))

    def noise(self, dt_ps, temperature, seed, step, gilbert_prefactor=False, normals_only=False):
        out = np.empty((self.N, 3))
        self._ck(self.lib.jb_noise(self.h, float(dt_ps), float(temperature), int(seed), int(step),
                                   int(gilbert_prefactor), int(normals_only), _ptr(out), 0))
        return out

    # ---- Hamiltonian / Monitor surface
    def fields(self, term=TERM_TOTAL, time_ps=0.0):
        out = np.empty((self.N, 3))
        self._ck(self.lib.jb_fields(self.h, int(term), float(time_ps), _ptr(out), 0))
        return out

    def energies(self, term, time_ps=0.0, per_spin=True):
        e = np.empty(self.N) if per_spin else None
        tot = C.c_double(0.0)
        self._ck(self.lib.jb_energies(self.h, int(term), float(time_ps), _ptr(e), 0, C.byref(tot)))
        return e, tot.value

    def magnetisation(self, group_of_spin=None, n_groups=1):
        g = None if group_of_spin is None else np.ascontiguousarray(group_of_spin, np.int32)
        out = np.zeros((n_groups, 4))
        self._ck(self.lib.jb_magnetisation(self.h, int(n_groups), _ptr(g), out))
        return out

    def set_magnetisation_groups(self, group_of_spin, n_groups):
        """register the monitor's groups once; afterwards ``magnetisation(None, n_groups)`` uses them"""
        g = None if group_of_spin is None else np.ascontiguousarray(group_of_spin, np.int32)
        self._ck(self.lib.jb_set_magnetisation_groups(self.h, int(n_groups), _ptr(g)))

    # ---- halo plumbing
    def halo_export_handle(self) -> bytes:
        buf = C.create_string_buffer(HALO_HANDLE_BYTES)
        self._ck(self.lib.jb_halo_export_handle(self.h, buf))
        return buf.raw

    def halo_connect(self, blob_lo, blob_hi):
        lo = C.create_string_buffer(blob_lo, HALO_HANDLE_BYTES) if blob_lo is not None else None
        hi = C.create_string_buffer(blob_hi, HALO_HANDLE_BYTES) if blob_hi is not None else None
        self._ck(self.lib.jb_halo_connect(self.h, lo, hi))

    # ---- introspection
    def kernel_launches(self):
        return int(self.lib.jb_kernel_launches(self.h))

    def stage_kernel(self):
        """jb_stage_kernel: 0 direct, 2 pair, 4 rows, 5 general neighbour list (-1 before the first step)"""
        return int(self.lib.jb_stage_kernel(self.h))

    def last_step_kernel_ms(self):
        """totals since the last call (option time_kernels > 0); divide by ``timed_steps()`` taken BEFORE this call"""
        out = np.zeros(2)
        self._ck(self.lib.jb_last_step_kernel_ms(self.h, out))
        return out

    def timed_steps(self):
        return int(self.lib.jb_timed_steps(self.h))

    def synchronize(self):
        self._ck(self.lib.jb_synchronize(self.h))

    def last_stage_trace(self, capacity=4096):
        """option trace = 1: per resident CTA of the last stage launch (SM id, first clock ns, last clock ns, items taken, then per
        item: item id << 40 | start ns after the CTA's first clock)"""
        out = np.zeros((capacity, 32), dtype=np.uint64)
        n = C.c_int32(0)
        self._ck(self.lib.jb_last_stage_trace(self.h, _ptr(out), int(capacity), C.byref(n)))
        return out[:n.value].copy()

    def stream(self):
        return self.lib.jb_stream(self.h)
