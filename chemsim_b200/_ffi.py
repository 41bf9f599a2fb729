"""ctypes binding of include/chemsim_lbm.h (the C-ABI library libchemsim_lbm.so).

This is the Python-side equivalent of the `extern "C"` block a Rust shim would
declare.  There is no fallback: if the library is missing or a
call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CHEMSIM_LBM_LIB selects an experimental build of the same library (tools/variants.sh)
LIB_PATH = os.environ.get("CHEMSIM_LBM_LIB") or os.path.join(HERE, "libchemsim_lbm.so")

F32, F64 = 0, 1
EDGE_ZEROFILL, EDGE_PERIODIC = 0, 1
OK, ERR_INVALID_ARGUMENT, ERR_INVALID_SLICE_SIZE, ERR_CUDA, ERR_NCCL, ERR_NOT_READY, ERR_UNSUPPORTED = range(7)
NCCL_ID_BYTES = 128


class LbmError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"chemsim_lbm status {status}: {message}")
        self.status = status
        self.message = message


class InvalidSliceSize(LbmError, ValueError):
    """matrix::Error::InvalidSliceSize (/root/reference/src/matrix.rs:15-19)."""


_H = C.c_void_p
_SZ = C.c_size_t
_P = C.c_void_p
_I = C.c_int
_D = C.c_double

# name -> (restype, argtypes); every symbol include/chemsim_lbm.h declares
PROTOTYPES = {
    "chemsim_lbm_abi_version": (_I, []),
    "chemsim_lbm_last_error": (C.c_char_p, [_H]),
    "chemsim_lbm_create": (_I, [_I, _I, _I, _I, _I, C.POINTER(_H)]),
    "chemsim_lbm_create_slab": (_I, [_I, _I, _I, _I, _I, _I, _I, _P, C.POINTER(_H)]),
    "chemsim_lbm_nccl_unique_id": (_I, [_P]),
    "chemsim_lbm_enable_p2p_halo": (_I, [_H]),
    "chemsim_lbm_halo_mode": (_I, [_H, C.POINTER(_I)]),
    "chemsim_lbm_destroy": (_I, [_H]),
    "chemsim_lbm_shape": (_I, [_H, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "chemsim_lbm_set_discretization": (_I, [_H, _D, _D]),
    "chemsim_lbm_set_bgk": (_I, [_H, _D]),
    "chemsim_lbm_set_trt": (_I, [_H, _D, _D]),
    "chemsim_lbm_set_regularized": (_I, [_H, _D]),
    "chemsim_lbm_set_kbc": (_I, [_H, _D]),
    "chemsim_lbm_kinematic_shear_viscosity": (_I, [_H, C.POINTER(_D)]),
    "chemsim_lbm_kinematic_bulk_viscosity": (_I, [_H, C.POINTER(_D)]),
    "chemsim_lbm_init_equilibrium": (_I, [_H, _P, _P, _P, _SZ]),
    "chemsim_lbm_init_equilibrium_rows": (_I, [_H, _I, _I, _P, _P, _P, _SZ]),
    "chemsim_lbm_set_population": (_I, [_H, _I, _P, _SZ]),
    "chemsim_lbm_set_geometry_rows": (_I, [_H, _I, _I, _P, _SZ]),
    "chemsim_lbm_set_geometry_async": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_get_density_async": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_set_geometry": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_step": (_I, [_H, _I]),
    "chemsim_lbm_synchronize": (_I, [_H]),
    "chemsim_lbm_time": (_I, [_H, C.POINTER(_D)]),
    "chemsim_lbm_get_density": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_get_pressure": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_get_speed": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_get_velocity": (_I, [_H, _P, _P, _SZ]),
    "chemsim_lbm_get_momentum_density": (_I, [_H, _P, _P, _SZ]),
    "chemsim_lbm_get_population": (_I, [_H, _I, _P, _SZ]),
    "chemsim_lbm_get_equilibrium": (_I, [_H, _I, _P, _SZ]),
    "chemsim_lbm_get_non_equilibrium": (_I, [_H, _I, _P, _SZ]),
    "chemsim_lbm_get_geometry": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_total_mass": (_I, [_H, C.POINTER(_D)]),
    "chemsim_lbm_total_mass_global": (_I, [_H, C.POINTER(_D)]),
    "chemsim_lbm_render": (_I, [_H, _I, _I, _P, _SZ]),
    "chemsim_lbm_is_unstable": (_I, [_H, C.POINTER(_I)]),
    "chemsim_lbm_cuda_stream": (_I, [_H, C.POINTER(_P)]),
    "chemsim_lbm_kernel_launches": (_I, [_H, C.POINTER(C.c_uint64)]),
    "chemsim_lbm_step_kernel_name": (C.c_char_p, [_H]),
    "chemsim_lbm_barrier": (_I, [_H]),
    "chemsim_lbm_set_p2p_timeout": (_I, [_H, _D]),
    "chemsim_lbm_set_stream_convention": (_I, [_H, _I]),
    "chemsim_lbm_fill_geometry": (_I, [_H, _I]),
    "chemsim_lbm_paint_rect": (_I, [_H, _I, _I, _I, _I, _I]),
    "chemsim_lbm_get_async": (_I, [_H, _I, _I, _P, _P, _SZ]),
    "chemsim_lbm_checkpoint_bytes": (_I, [_H, C.POINTER(_SZ)]),
    "chemsim_lbm_checkpoint": (_I, [_H, _P, _SZ]),
    "chemsim_lbm_restore": (_I, [_H, _P, _SZ]),
}
ABI_VERSION = 2
(FIELD_DENSITY, FIELD_PRESSURE, FIELD_SPEED, FIELD_VELOCITY, FIELD_MOMENTUM_DENSITY, FIELD_POPULATION,
 FIELD_EQUILIBRIUM, FIELD_NON_EQUILIBRIUM) = range(8)



class HaloMsg(C.Structure):
    _fields_ = [("is_send", C.c_int), ("peer", C.c_int), ("q", C.c_int), ("row", C.c_int)]


HALO_PLAN_MAX = 48
(ROW_FIRST, ROW_LAST, ROW_GHOST_ABOVE, ROW_GHOST_BELOW, ROW_SECOND, ROW_SECOND_LAST, ROW_GHOST_ABOVE2,
 ROW_GHOST_BELOW2) = range(8)
PROTOTYPES["chemsim_lbm_slab_rows"] = (_I, [_I, _I, _I, C.POINTER(_I), C.POINTER(_I)])
PROTOTYPES["chemsim_lbm_halo_plan"] = (_I, [_I, _I, _I, _I, C.POINTER(HaloMsg), C.POINTER(_I)])

_lib = None


def load():
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m chemsim_b200.build` "
                "(there is no CPU or PyTorch fallback for the D2Q9 path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)      # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, handle=None):
    if status == OK:
        return
    msg = load().chemsim_lbm_last_error(handle)
    msg = msg.decode() if msg else ""
    if status == ERR_INVALID_SLICE_SIZE:
        raise InvalidSliceSize(status, msg)
    raise LbmError(status, msg)
