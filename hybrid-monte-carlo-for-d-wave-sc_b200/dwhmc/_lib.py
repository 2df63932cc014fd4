"""ctypes binding of libdwhmc.so (include/dwhmc.h).  There is no CPU fallback: if the CUDA
library has not been built, importing this module fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("DWHMC_LIB", os.path.join(_PKG_DIR, "libdwhmc.so"))

OK, E_BADARG, E_CUDA, E_NOCONV, E_NODEVICE, E_STATE = range(6)
NOBS = 9


class DwhmcError(RuntimeError):
    """Nonzero return code of a libdwhmc call (the reference raises Julia exceptions)."""

    def __init__(self, code: int, text: str):
        super().__init__(f"libdwhmc error {code}: {text}")
        self.code = code


class EigenConvergenceError(DwhmcError):
    """Twin of LAPACKException from eigen! (src/Hamiltonian.jl:106)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C {os.path.join(_PKG_DIR, 'csrc')}` "
        "(or __graft_entry__.build()); dwhmc has no CPU fallback")

lib = C.CDLL(LIB_PATH)

_vp, _i, _dp, _ip = C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/dwhmc.h
PROTOTYPES = {
    "dwhmc_create": [C.POINTER(_vp), _i, _i, _i, _i, _i64p, _i64p],
    "dwhmc_destroy": [_vp],
    "dwhmc_last_error": [_vp],
    "dwhmc_version": [],
    "dwhmc_dims": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "dwhmc_set_params": [_vp, _dp, _dp, _dp, _dp, _dp, _dp],
    "dwhmc_set_disorder": [_vp, _dp],
    "dwhmc_set_field": [_vp, _dp],
    "dwhmc_get_field": [_vp, _dp],
    "dwhmc_set_momentum": [_vp, _dp],
    "dwhmc_get_momentum": [_vp, _dp],
    "dwhmc_seed": [_vp, C.c_uint64],
    "dwhmc_init_static_H": [_vp],
    "dwhmc_update_H": [_vp],
    "dwhmc_diagonalize": [_vp],
    "dwhmc_compute_forces": [_vp],
    "dwhmc_total_energy": [_vp, _dp],
    "dwhmc_measure_observables": [_vp, _dp],
    "dwhmc_get_H": [_vp, _dp],
    "dwhmc_get_eigenvalues": [_vp, _dp],
    "dwhmc_get_eigenvectors": [_vp, _dp],
    "dwhmc_get_forces": [_vp, _dp],
    "dwhmc_get_fermi": [_vp, _dp],
    "dwhmc_trajectory": [_vp, _ip, _dp, _dp, _dp, _dp, _dp],
    "dwhmc_commit": [_vp, _ip],
    "dwhmc_hmc_sweep": [_vp, _ip, _dp, _dp, _dp, _ip, _dp],
    "dwhmc_run_sweeps": [_vp, _i, _ip, _dp, _ip, _dp, _dp],
    "dwhmc_eigensolver_route": [_vp, C.POINTER(C.c_int)],
    "dwhmc_init_state": [_vp, _dp, _dp],
    "dwhmc_get_disorder": [_vp, _dp],
    "dwhmc_measure_transport": [_vp, C.c_double, _dp, _i, _dp, _i, _dp, _dp, _dp, _dp, _dp],
    "dwhmc_get_timers": [_vp, _dp],
    "dwhmc_reset_timers": [_vp],
    "dwhmc_last_elapsed_ms": [_vp, _dp],
    "dwhmc_set_profiling": [_vp, _i],
    "dwhmc_debug_tridiagonalize": [_vp, _dp, _dp],
    "dwhmc_debug_stedc": [_vp, _dp, _dp, _dp, _dp],
    "dwhmc_debug_heev": [_vp, _dp, _dp, _dp],
}
_RESTYPES = {"dwhmc_last_error": C.c_char_p, "dwhmc_version": C.c_char_p}

for _name, _args in PROTOTYPES.items():
    _f = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _f.argtypes = _args
    _f.restype = _RESTYPES.get(_name, C.c_int)


def dptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, np.complex128), (a.dtype, a.flags)
    return a.ctypes.data_as(_dp)


def iptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.int32
    return a.ctypes.data_as(_ip)


def check(rc: int, handle=None) -> None:
    if rc == OK:
        return
    text = lib.dwhmc_last_error(handle)
    text = text.decode() if text else ""
    raise (EigenConvergenceError if rc == E_NOCONV else DwhmcError)(rc, text)


def version() -> str:
    return lib.dwhmc_version().decode()
