"""ctypes binding of libebm_b200.so (the C ABI declared in include/ebm_b200.h).

There is no fallback: if the library is missing or fails to load, every product entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
# EBM_B200_LIB: load another build of the same ABI (A/B timing of kernel variants); still no fallback of any kind
LIB_PATH = os.environ.get("EBM_B200_LIB") or os.path.join(HERE, "lib", "libebm_b200.so")

EBM_ABI_VERSION = 9

ENERGY_DOUBLE_WELL, ENERGY_HARMONIC, ENERGY_RASTRIGIN, ENERGY_GAUSSIAN, ENERGY_MOG, ENERGY_MLP = range(6)
ACT_SILU, ACT_TANH, ACT_RELU, ACT_SOFTPLUS = range(4)
RNG_INJECTED, RNG_TORCH, RNG_NATIVE = range(3)
MASS_NONE, MASS_SCALAR, MASS_VECTOR = range(3)
MLP_FP32, MLP_BF16X3, MLP_BF16 = range(3)
MLP_PRECISIONS = {"fp32": MLP_FP32, "bf16x3": MLP_BF16X3, "bf16": MLP_BF16}
ERR_INVALID, ERR_UNSUPPORTED = -1, -2

RNG_MODES = {"injected": RNG_INJECTED, "torch": RNG_TORCH, "native": RNG_NATIVE}


class EbmEnergyDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dim", C.c_int32),
        ("n_components", C.c_int32),
        ("hidden1", C.c_int32),
        ("hidden2", C.c_int32),
        ("activation", C.c_int32),
        ("precision", C.c_int32),
        ("sm_margin", C.c_int32),
        ("hidden3", C.c_int32),
        ("reserved", C.c_int32),
        ("p", C.c_float * 4),
        ("buf", C.c_void_p * 10),
    ]


class EbmLibraryError(RuntimeError):
    pass


class EbmUnsupported(EbmLibraryError):
    """The request is valid but this build has no kernel for it (EBM_ERR_UNSUPPORTED)."""


_P = C.c_void_p
_I32, _I64, _U64, _F64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double
_DESC = C.POINTER(EbmEnergyDesc)
_PD = C.POINTER(C.c_double)
_PF = C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every symbol of include/ebm_b200.h (tests/test_cabi.py checks)
PROTOTYPES = {
    "ebm_abi_version": (C.c_int, []),
    "ebm_last_error": (C.c_char_p, []),
    "ebm_device_sm_count": (C.c_int, [C.c_int]),
    "ebm_torch_rng_threads": (_I64, [C.c_int, _I64]),
    "ebm_torch_rng_offset_increment": (_I64, [C.c_int, _I64]),
    "ebm_workspace_bytes": (_I64, [_DESC]),
    "ebm_energy_f32": (C.c_int, [_DESC, _P, _I64, _P, _P]),
    "ebm_gradient_f32": (C.c_int, [_DESC, _P, _I64, _P, _P]),
    "ebm_euler_maruyama_step_f32": (C.c_int, [_P, _P, _P, _P, _I64, _F64, _F64, _P]),
    "ebm_langevin_burst_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64, _U64, _P, _P, _I32, _P]),
    "ebm_langevin_heun_burst_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64, _U64, _P, _P, _I32, _P]),
    "ebm_peer_push_f32": (C.c_int, [_P, _I64, C.POINTER(C.c_void_p), _I32, _I64, _I32, _P]),
    "ebm_descent_burst_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _PD, _I32, _F64, _P, _P, _I32, _P]),
    "ebm_langevin_burst_gather_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64, _U64,
                                                C.POINTER(C.c_void_p), _I32, _I64, _P]),
    "ebm_langevin_burst_host_f32": (C.c_int, [_DESC, _P, _P, _P, _I64, _I32, _F64, _F64, _I32, _U64, _U64, _P]),
    "ebm_leapfrog_f32": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I32, _F64, _I32, _F64, _P, _I32, _P]),
    "ebm_hmc_burst_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _I32, _PD, _I32, _I32, _F64, _P, _I32, _U64, _U64, _P, _P, _P, _I32, _P, _P, _P]),
    "ebm_pcd_gather_f32": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _P, _P, _I64, _P]),
    "ebm_pcd_scatter_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, C.POINTER(C.c_int64), _P]),
    "ebm_pcd_langevin_fused": (C.c_int, [_DESC]),
    "ebm_pcd_langevin_burst_f32": (C.c_int, [_DESC, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64, _U64,
                                             _P, _P, _I64, _P, C.POINTER(C.c_int64), _P]),
    "ebm_pcd_langevin_burst_gather_f32": (C.c_int, [_DESC, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64,
                                                    _U64, _P, _P, _I64, _P, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), _I32,
                                                    _I64, _P]),
    "ebm_langevin_burst_diag_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _PD, _PD, _I32, _PF, _I32, _U64, _U64, _P, _P, _I32, _I32,
                                              _P, _P, _P, _P, _P, _P]),
    "ebm_hmc_burst_diag_f32": (C.c_int, [_DESC, _P, _P, _I64, _I32, _I32, _PD, _I32, _I32, _F64, _P, _I32, _U64, _U64, _P, _P, _P,
                                         _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ebm_ess_f32": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "ebm_rng_fill_f32": (C.c_int, [_P, _I64, _I32, _I32, _U64, _U64, _P]),
}

_lock = threading.Lock()
_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the CUDA library or raise.  Never falls back to another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EbmLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m torchebm_b200.build` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.ebm_abi_version() != EBM_ABI_VERSION:
            raise EbmLibraryError(f"ABI mismatch: library {lib.ebm_abi_version()} vs binding {EBM_ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = load().ebm_last_error().decode(errors="replace")
    if rc == ERR_UNSUPPORTED:
        raise EbmUnsupported(f"{what}: {msg}")
    if rc == ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise EbmLibraryError(f"{what}: CUDA error {rc}: {msg}")


def doubles(values) -> "C.Array":
    arr = (C.c_double * len(values))(*values)
    return arr
