"""ctypes face of the C++ host layer (jams_b200/host/, libjams_b200_host.so): JAMS config files in, the llg-heun
B200 path out.  The Python mirror in solver.py drives the same C ABI; this module exists so tests can check that the
C++ layer parses the same files into the same lattice / template / trajectory."""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjams_b200_host.so")
EXE_PATH = os.path.join(HERE, "jams-b200")
_lib = None


class HostError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make -C jams_b200/host` (__graft_entry__.build())")
        capi.load()   # libjams_b200.so first (rpath $ORIGIN finds it too)
        lib = C.CDLL(LIB_PATH)
        lib.jbh_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def _args(args):
    arr = (C.c_char_p * len(args))(*[a.encode() for a in args])
    return arr, len(args)


def _ck(rc):
    if rc != 0:
        raise HostError(load().jbh_last_error().decode())


def config_to_dict(*args: str) -> dict:
    """merge config files / strings left to right (core/jams++.cc:48-84) and return the result"""
    lib = load()
    a, n = _args(args)
    buf = C.create_string_buffer(1 << 20)
    _ck(lib.jbh_config_to_json(a, n, buf, C.c_longlong(len(buf))))
    return json.loads(buf.value.decode())


def lattice_arrays(*args: str):
    lib = load()
    a, n = _args(args)
    ns, M = C.c_int(), C.c_int()
    dims, per = (C.c_int * 3)(), (C.c_int * 3)()
    _ck(lib.jbh_lattice_info(a, n, C.byref(ns), C.byref(M), dims, per))
    N = ns.value
    out = dict(num_spins=N, M=M.value, dims=tuple(dims), periodic=tuple(bool(p) for p in per), mus=np.zeros(N), gyro=np.zeros(N), alpha=np.zeros(N),
               spins=np.zeros((N, 3)), positions=np.zeros((N, 3)))
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    _ck(lib.jbh_lattice_arrays(a, n, p(out["mus"]), p(out["gyro"]), p(out["alpha"]), p(out["spins"]), p(out["positions"])))
    return out


def exchange_template(*args: str, ham_index: int, capacity: int = 4096):
    lib = load()
    a, n = _args(args)
    mi = np.zeros(capacity, np.int32); mj = np.zeros(capacity, np.int32); T = np.zeros(3 * capacity, np.int32); J9 = np.zeros(9 * capacity)
    k, pairs = C.c_int(), C.c_longlong()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    _ck(lib.jbh_exchange_template(a, n, int(ham_index), int(capacity), C.byref(k), p(mi), p(mj), p(T), p(J9), C.byref(pairs)))
    k = k.value
    return dict(mi=mi[:k].copy(), mj=mj[:k].copy(), T=T[:3 * k].reshape(k, 3).copy(), J9=J9[:9 * k].reshape(k, 9).copy(), n_pairs=pairs.value)


def run(*args: str, name="jams", output_dir=".", max_steps=-1, num_spins=None):
    """run a configuration on the GPU through the C++ Simulation; returns (final spins N x 3, steps done)"""
    lib = load()
    a, n = _args(args)
    if num_spins is None:
        num_spins = lattice_arrays(*args)["num_spins"]
    spins = np.zeros((num_spins, 3))
    steps = C.c_int()
    _ck(lib.jbh_run(a, n, name.encode(), output_dir.encode(), int(max_steps), spins.ctypes.data_as(C.c_void_p), C.byref(steps)))
    return spins, steps.value
