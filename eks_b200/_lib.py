"""ctypes binding of libeks_b200.so (the C ABI declared in include/eks_b200.h).

There is NO CPU fallback: importing this module on a machine where the library has not been built,
or calling into it without a CUDA device, raises.  The library is built in-tree by
``python -m eks_b200.build`` (or ``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_longlong, c_size_t, c_void_p

import numpy as np
import torch

F32, F64 = 0, 1
ABI_VERSION = 202      # eks_version() of the library this package was written against (include/eks_b200.h)
MAX_CHAN, MAX_STATE, CAM_STRIDE = 16, 6, 29

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EKS_B200_LIB') or os.path.join(_HERE, 'lib', 'libeks_b200.so')
_lib = None


class EksB200Error(RuntimeError):
    pass


_SIGS = {
    'eks_last_error': (ctypes.c_char_p, []),
    'eks_version': (c_int, []),
    'eks_ensemble_tile_frames': (c_int, [c_int, c_int, c_int, c_int]),
    'eks_ensemble_stats': (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_double, c_void_p, c_int, c_longlong, c_longlong, c_longlong, c_void_p,
                                   c_void_p, c_void_p]),
    'eks_center_moments': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'eks_initial_guess': (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p]),
    'eks_const_R_median_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'eks_const_R_median': (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_double, c_void_p, c_void_p, c_size_t, c_void_p]),
    'eks_nll_grad': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'eks_optimize_s_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    'eks_optimize_s': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double,
                               c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_size_t, c_void_p]),
    'eks_filter_smooth_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'eks_filter_smooth': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                  c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    # (round 1 loaded the entry points below only if present; a stale library now fails at load time)
    'eks_triangulate_mean': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'eks_geometric_init_workspace_bytes': (c_size_t, [c_int, c_int]),
    'eks_geometric_init': (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'eks_last_launch_count': (c_int, []),
    'eks_last_unverified_count': (c_int, []),
    'eks_mc_valid_moments': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                     c_longlong, c_void_p, c_void_p, c_longlong, c_void_p, c_double, c_double,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'eks_mc_inflate_step': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p,
                                    c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double,
                                    c_void_p, c_void_p, c_void_p]),
    'eks_mc_prestage_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'eks_mc_center': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_longlong,
                              c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'eks_mc_pca_moments': (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    'eks_mc_latent_init': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'eks_pupil_optimize_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'eks_pupil_optimize': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                   c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int, c_void_p, c_void_p,
                                   c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p, c_size_t, c_void_p]),
    'eks_reproject': (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p, c_longlong, c_void_p, c_int, c_void_p, c_longlong, c_longlong,
                              c_void_p, c_void_p]),
    'eks_diag_smooth_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'eks_diag_smooth': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p,
                                c_void_p, c_void_p, c_longlong, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
}


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it is missing -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EksB200Error(
                f'{LIB_PATH} not found: build it with `python -m eks_b200.build` '
                '(eks_b200 has no CPU fallback)')
        L = ctypes.CDLL(LIB_PATH)
        missing = [name for name in _SIGS if not hasattr(L, name)]
        if missing:   # every symbol of include/eks_b200.h is mandatory: never select another path on a stale build
            raise EksB200Error(f'{LIB_PATH} is stale: missing {missing}; rebuild with `python -m eks_b200.build --force`')
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.eks_version() != ABI_VERSION:
            raise EksB200Error(f'{LIB_PATH} has ABI version {L.eks_version()}, this package needs {ABI_VERSION}; '
                               'rebuild with `python -m eks_b200.build --force`')
        _lib = L
    return _lib


def exported_symbols() -> list[str]:
    return list(_SIGS)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().eks_last_error().decode(errors='replace')
        raise EksB200Error(f'{what} failed (rc={rc}): {msg}')


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise EksB200Error('eks_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError(f'unsupported dtype {dtype}')


def ptr(t) -> int | None:
    """Device (or host) address of a tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    return t.ctypes.data


def i64_host(vals) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(vals, dtype=np.int64))


def i32_host(vals) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(vals, dtype=np.int32))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
