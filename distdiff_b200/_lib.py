"""ctypes binding of libdistdiff_sm100.so (include/distdiff_sm100.h).

There is NO fallback: if the shared library is missing, or a call returns non-zero, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DD_LIB_PATH: development only -- load an experimental build of the same library (tools/build_variant.sh)
LIB_PATH = os.environ.get("DD_LIB_PATH") or os.path.join(HERE, "libdistdiff_sm100.so")

DD_F32, DD_F16, DD_BF16 = 0, 1, 2
ABI_VERSION = 1

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_z = C.c_size_t

# name -> (restype, argtypes) ; mirrors include/distdiff_sm100.h one to one
SIGNATURES = {
    "dd_abi_version": (_i, []),
    "dd_last_error": (C.c_char_p, []),
    "dd_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dd_cfg_ddim_fwd": (_i, [_p, _p, _p, _l, _i, _f, _f, _f, _p, _f, _p, _p, _p]),
    "dd_cfg_ddim_bwd": (_i, [_p, _p, _l, _i, _f, _f, _f, _i, _p, _p, _p, _p]),
    "dd_affine_project_fwd": (_i, [_p, _p, _p, _p, _l, _l, _i, _f, _p, _p]),
    "dd_affine_bwd": (_i, [_p, _p, _p, _l, _l, _i, _p, _p, _p, _p]),
    "dd_add_noise": (_i, [_p, _p, _l, _i, _f, _p, _p]),
    "dd_bicubic_resize_fwd": (_i, [_p, _l, _i, _i, _i, _i, _i, _p, _p]),
    "dd_bicubic_resize_bwd": (_i, [_p, _l, _i, _i, _i, _i, _i, _p, _p]),
    "dd_image_to_uint8": (_i, [_p, _l, _i, _i, _i, _i, _i, _p, _p]),
    "dd_energy_workspace_bytes": (_z, [_i, _i]),
    "dd_energy_fwd_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p, _z, _i, _p]),
    "dd_proto_workspace_bytes": (_z, [_i, _i, _i]),
    "dd_rownorm_classsum": (_i, [_p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _z, _p]),
    "dd_class_mean": (_i, [_p, _p, _l, _i, _p, _p, _p]),
    "dd_normalize_rows": (_i, [_p, _l, _i, _p, _p]),
    "dd_kmeans_seed": (_i, [_p, _p, _l, _i, _p, _p, _p]),
    "dd_kmeans_assign_accum": (_i, [_p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "dd_kmeans_update": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "dd_agglo_workspace_bytes": (_z, [_l, _i]),
    "dd_agglo_average": (_i, [_p, _p, _i, _i, _i, _l, _p, _p, _p, _p, _p, _z, _p]),
    "dd_comm_unique_id": (_i, [_p]),
    "dd_comm_init": (_i, [_i, _i, _p, C.POINTER(_p)]),
    "dd_comm_allreduce": (_i, [_p, _p, _z, _p, _z, _p]),
    "dd_comm_destroy": (_i, [_p]),
    "dd_peer_create": (_i, [_i, _i, _z, C.POINTER(_p), _p]),
    "dd_peer_connect": (_i, [_p, _p]),
    "dd_peer_local": (_p, [_p]),
    "dd_peer_header_bytes": (_z, []),
    "dd_peer_arena_bytes": (_z, [_i, _i, _i]),
    "dd_peer_kmeans_exchange": (_i, [_p, _z, _z, _z, _z, _z, _i, _i, _p]),
    "dd_peer_status": (_i, [_p, _p, C.POINTER(_i)]),
    "dd_peer_timing": (_i, [_p, _p, C.POINTER(C.c_double)]),
    "dd_kmeans_lloyd": (_i, [_p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _z, _p, _p, _i, _p]),
    "dd_peer_destroy": (_i, [_p]),
}

_lib = None


class DistDiffError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the CDLL with argtypes set.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DistDiffError(
            f"{LIB_PATH} is missing: build it with `python -m distdiff_b200.build` "
            "(there is no CPU / PyTorch fallback for the guidance hot path)")
    # NCCL is resolved through torch's bundled copy (rpath); importing torch first guarantees one instance.
    import torch  # noqa: F401
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    got = handle.dd_abi_version()
    if got != ABI_VERSION:
        raise DistDiffError(f"libdistdiff_sm100.so ABI {got} != expected {ABI_VERSION}; rebuild")
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dd_last_error().decode(errors="replace")
        raise DistDiffError(f"{what} failed (code {rc}): {msg}")
