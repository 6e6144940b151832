"""Drop-in mirror of eks/core.py's public functions, running on the CUDA library.

  ensemble()              <- eks/core.py:25-101
  run_kalman_smoother()   <- eks/core.py:159-302
  optimize_smooth_param() <- eks/core.py:306-401 (+ fast path :562-699, block path :403-559)
  compute_initial_guesses / constant_R_from_timevarying <- eks/core.py:104-133, :702-709

Same names, argument meaning, defaults, return types and error behaviour.  Arguments may be NumPy
arrays, torch tensors or anything np.asarray-able; results come back as NumPy (as the reference
returns).  All arithmetic runs in the kernels: host code here only validates arguments, lays the data
out as frame-major planes on the device and reads results back.  No CPU fallback exists.

Precision: float32 is the production precision of the reference; call
``eks_b200.set_precision('float64')`` for the parity mode (the analogue of jax_enable_x64).
"""

from __future__ import annotations

import logging
import time
from collections.abc import Callable
from typing import Literal

import numpy as np
import torch

from eks_b200 import ops
from eks_b200._lib import require_cuda
from eks_b200.marker_array import MarkerArray
from eks_b200.ops import Model, PlaneView
from eks_b200.utils import normalize_spans

logger = logging.getLogger(__name__)

_PRECISION = torch.float32


def set_precision(precision: str) -> None:
    """'float32' (production, default) or 'float64' (parity mode)."""
    global _PRECISION
    if precision not in ('float32', 'float64'):
        raise ValueError("precision must be 'float32' or 'float64'")
    _PRECISION = torch.float32 if precision == 'float32' else torch.float64


def get_precision() -> torch.dtype:
    return _PRECISION


class PinholeProjection:
    """Calibrated multi-view emission h: R^3 -> R^{2V} (eks/multicam_smoother.py:806-885).

    The reference passes an arbitrary JAX callable as ``h_fn``; the CUDA path needs the camera
    parameters themselves, so ``make_projection_from_camgroup`` returns this object (callable on the
    host for convenience, via the library)."""

    def __init__(self, cams: np.ndarray):
        self.cams = np.ascontiguousarray(np.asarray(cams, dtype=np.float64).reshape(-1, 29))

    @property
    def n_cameras(self) -> int:
        return self.cams.shape[0]

    def __call__(self, object_points) -> np.ndarray:
        """h(X): (..., 3) world points -> (..., 2V) pixels [u_0, v_0, u_1, v_1, ...] (one camera: (..., 2)), evaluated
        by the device projection of the library (eks_reproject) in float64 -- the same code the EKF linearises."""
        dev = require_cuda()
        X = np.asarray(object_points, dtype=np.float64)
        lead, pts = X.shape[:-1], X.reshape(-1, 3)
        N, V = pts.shape[0], self.n_cameras
        ms = torch.as_tensor(np.ascontiguousarray(pts)).to(dev).reshape(1, N, 3)
        Vs = torch.zeros((1, N, 3, 3), dtype=torch.float64, device=dev)
        out = torch.empty((V, 4, N), dtype=torch.float64, device=dev)
        ops.reproject(ms, Vs, V, out, V * 4 * N, 4 * N, [0, N, 2 * N, 3 * N],
                      cams=torch.as_tensor(self.cams).to(dev))
        uv = out[:, :2, :].permute(2, 0, 1).reshape(N, 2 * V).cpu().numpy()
        return uv.reshape(lead + (2 * V,))


def _to_device(x, dtype, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), device='cpu').to(device=device, dtype=dtype)


# ----------------------------------------------------------------------------- ensemble
def ensemble(
    marker_array: MarkerArray,
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
    nan_replacement: float = 1000.0,
) -> MarkerArray:
    """Ensemble mean/median and variance across models -> MarkerArray (1, V, T, K, 5) with fields
    ['x', 'y', 'var_x', 'var_y', 'likelihood']."""
    dev = require_cuda()
    dtype = get_precision()
    M, V, T, K, _ = marker_array.shape
    arr = marker_array.slice_fields('x', 'y', 'likelihood').array if list(marker_array.data_fields) != [
        'x', 'y', 'likelihood'] else marker_array.array
    raw = torch.as_tensor(np.ascontiguousarray(arr)).to(dev)
    if raw.dtype not in (torch.float32, torch.float64):
        raw = raw.to(torch.float64)
    if raw.dtype == torch.float32 and dtype == torch.float64:
        raw = raw.to(torch.float64)
    planes = torch.empty((V, K, 5, T), dtype=dtype, device=dev)
    ops.ensemble_stats(raw.reshape(1, M, V, T, K, 3), planes, 0, K * 5 * T, 5 * T, [f * T for f in range(5)],
                       avg_mode=avg_mode, var_mode=var_mode, nan_replacement=nan_replacement)
    out = planes.permute(0, 3, 1, 2).contiguous().cpu().numpy()  # (V,T,K,5)
    return MarkerArray(out[None, ...], data_fields=['x', 'y', 'var_x', 'var_y', 'likelihood'])


# ----------------------------------------------------------------------------- small mirrors (API)
def compute_initial_guesses(ensemble_vars) -> float:
    """Host mirror of eks/core.py:104-133 (the device path computes the same quantity in
    eks_initial_guess)."""
    ev = np.asarray(ensemble_vars)[:2000]
    if ev.shape[0] < 2:
        raise ValueError('Not enough frames to compute temporal differences.')
    d = ev[1:] - ev[:-1]
    return float(round(np.nanstd(d), 5))


def constant_R_from_timevarying(R_t_np: np.ndarray, min_var: float = 1e-4) -> np.ndarray:
    """Host mirror of eks/core.py:702-709 (the device path uses eks_const_R_median)."""
    diag_ts = np.diagonal(R_t_np, axis1=-2, axis2=-1)
    med = np.clip(np.nanmedian(diag_ts, axis=0), min_var, np.inf)
    return np.diag(med).astype(R_t_np.dtype)


# ----------------------------------------------------------------------------- device staging
def _is_diag_model(m0s, S0s, As, Cs, Qs, h_fn) -> bool:
    """True for the decoupled single-camera structure: D == O == 2, diagonal A, C, Q, S0 (checked on
    the host copies of the tiny parameter arrays before upload)."""
    if h_fn is not None or isinstance(S0s, torch.Tensor) and S0s.is_cuda:
        return False
    try:
        S0, A, C, Q = (np.asarray(x, dtype=np.float64) for x in (S0s, As, Cs, Qs))
    except Exception:
        return False
    if A.shape[-2:] != (2, 2) or C.shape[-2:] != (2, 2):
        return False
    off = lambda x: np.all(x[..., 0, 1] == 0) and np.all(x[..., 1, 0] == 0)
    return bool(off(S0) and off(A) and off(C) and off(Q))


def _stage(ys, m0s, S0s, As, Cs, Qs, h_fn, dev, dtype):
    """(K,T,O) observations -> frame-major planes [K][O][T] + per-sequence model on the device."""
    y = _to_device(ys, dtype, dev)
    K, T, O = y.shape
    y_planes = y.permute(0, 2, 1).contiguous()
    if h_fn is not None:
        if not isinstance(h_fn, PinholeProjection):
            raise TypeError(
                'h_fn must be the PinholeProjection returned by eks_b200.multicam_smoother.'
                'make_projection_from_camgroup: the CUDA path needs camera parameters, arbitrary callables '
                'cannot be traced (there is no CPU fallback)')
        cams = _to_device(h_fn.cams, dtype, dev).contiguous()
        C = None
    else:
        cams = None
        C = _to_device(Cs, dtype, dev).contiguous()
    model = Model(_to_device(m0s, dtype, dev).contiguous(), _to_device(S0s, dtype, dev).contiguous(),
                  _to_device(As, dtype, dev).contiguous(), _to_device(Qs, dtype, dev).contiguous(), C, cams)
    yv = PlaneView(y_planes, O * T, [o * T for o in range(O)])
    return model, yv, K, T, O


def _finalize_s(opt, K, s_bounds_log):
    s_log = opt['s_log'].double().cpu().numpy()
    loss = opt['loss'].double().cpu().numpy()
    iters = opt['iters'].cpu().numpy()
    s_out = np.empty(K, dtype=float)
    for j, blk in enumerate(opt['blocks']):
        s_star = float(np.exp(np.clip(s_log[j], s_bounds_log[0], s_bounds_log[1])))
        for k in blk:
            s_out[k] = s_star
        logger.debug(f'[opt s | block {list(blk)}] s={s_star:.6g}, iters={int(iters[j])}, NLL={float(loss[j]):.6f}')
    return s_out, iters, loss


# ----------------------------------------------------------------------------- public API
def run_kalman_smoother(
    ys,                              # (K, T, obs)
    m0s,                             # (K, D)
    S0s,                             # (K, D, D)
    As,                              # (K, D, D)
    Cs,                              # (K, obs, D)
    Qs,                              # (K, D, D)
    ensemble_vars,                   # (T, K, obs)
    s_frames: list | None = None,
    smooth_param: float | list | None = None,
    blocks: list | None = None,
    lr: float = 0.25,
    s_bounds_log: tuple = (-8.0, 8.0),
    tol: float = 1e-2,
    safety_cap: int = 300,
    h_fn: Callable | None = None,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Optimise the process-noise scale s per block of keypoints on the EKF filter NLL, then run the
    EKF smoother.  Returns (s_finals (K,), ms (K,T,D), Vs (K,T,D,D))."""
    dev = require_cuda()
    dtype = get_precision()
    model, yv, K, T, O = _stage(ys, m0s, S0s, As, Cs, Qs, h_fn, dev, dtype)
    structure = ops.STRUCT_DIAG if _is_diag_model(m0s, S0s, As, Cs, Qs, h_fn) else ops.STRUCT_GENERAL
    if not blocks:
        blocks = [[k] for k in range(K)]
    logger.debug(f'correlated keypoint blocks: {blocks}')

    t0 = time.perf_counter()
    ev = _to_device(ensemble_vars, dtype, dev)                 # (T,K,O)
    var_planes = ev.permute(1, 2, 0).contiguous()               # [K][O][T]
    vv = PlaneView(var_planes, O * T, [o * T for o in range(O)])
    logger.debug(f'[profile]   build_R: {time.perf_counter() - t0:.3f}s')
    if T < 2:
        raise ValueError('Not enough frames to compute temporal differences.')

    s_finals = np.empty(K, dtype=float)
    if smooth_param is not None:
        if isinstance(smooth_param, (int, float)):
            s_finals[:] = float(smooth_param)
        else:
            s_finals[:] = np.asarray(smooth_param, dtype=float)
    else:
        t0 = time.perf_counter()
        spans = normalize_spans(T, s_frames)
        guess, s_log0 = ops.initial_guess(vv, K, T)
        _optimize_on_device(model, yv, vv, T, spans, blocks, guess, s_log0, lr, s_bounds_log, tol, safety_cap,
                            1e-4, s_finals, structure=structure)
        logger.debug(f'[profile]   optimize_smooth_param: {time.perf_counter() - t0:.3f}s')

    t0 = time.perf_counter()
    s_dev = torch.as_tensor(s_finals, device=dev).to(dtype)
    if structure == ops.STRUCT_DIAG:
        # decoupled model: time-parallel filter + RTS kernels, latent moments as planes [K][4][T]
        lat = torch.empty((K, 4, T), dtype=dtype, device=dev)
        ops.diag_smooth(model, yv, vv, T, s_dev, None, lat, 4 * T, [0, T, 2 * T, 3 * T], latent_out=True)
        lat_np = lat.cpu().numpy()
        ms_np = np.ascontiguousarray(np.transpose(lat_np[:, 0:2, :], (0, 2, 1)))
        Vs_np = np.zeros((K, T, 2, 2), dtype=lat_np.dtype)
        Vs_np[:, :, 0, 0] = lat_np[:, 2, :]
        Vs_np[:, :, 1, 1] = lat_np[:, 3, :]
    else:
        ms, Vs = ops.filter_smooth(model, yv, vv, T, s_dev)
        ms_np, Vs_np = ms.cpu().numpy(), Vs.cpu().numpy()
    logger.debug(f'[profile]   final smoother pass ({K} keypoints): {time.perf_counter() - t0:.3f}s')
    return s_finals, ms_np, Vs_np


def _optimize_on_device(model, yv, vv, T, spans, blocks, guess, s_log0, lr, s_bounds_log, tol, safety_cap,
                        min_R_var, s_finals, trace_cap=0, structure=0):
    K = model.B
    Rconst = ops.const_R_median(vv, K, T, spans=spans, min_var=min_R_var)
    if all(len(b) == 1 for b in blocks) and [b[0] for b in blocks] == list(range(K)):
        s0 = s_log0
    else:  # block seed: mean of member guesses, clipped, float32 log (core.py:439-441)
        g = guess.cpu().numpy()
        s0_host = np.array([np.log(np.clip(np.mean([g[k] for k in b]), 1e-6, 1e3)) for b in blocks])
        s0 = torch.as_tensor(s0_host.astype(np.float32), device=guess.device).to(model.m0.dtype)
    opt = ops.optimize_s(model, yv, T, Rconst, s0, blocks=blocks, spans=spans, lr=lr, s_bounds_log=s_bounds_log,
                         tol=tol, safety_cap=safety_cap, trace_cap=trace_cap,
                         structure=structure if (spans is None or len(spans) == 1) else ops.STRUCT_GENERAL)
    s_out, iters, loss = _finalize_s(opt, K, s_bounds_log)
    covered = sorted(k for b in blocks for k in b)
    for k in covered:
        s_finals[k] = s_out[k]
    return opt, iters, loss


def optimize_smooth_param(
    ys, m0s, S0s, As, Cs, Qs,
    Rs,                              # (K, T, obs, obs) time-varying R_t (only its diagonal is used)
    blocks: list | None,
    s_finals: np.ndarray,            # (K,), filled in place
    s_frames: list | None,
    s_guess_per_k: np.ndarray,       # (K,)
    lr: float = 0.25,
    s_bounds_log: tuple = (-8.0, 8.0),
    tol: float = 1e-3,
    safety_cap: int = 300,
    min_R_var: float = 1e-4,
    h_fn_combined: Callable | None = None,
) -> None:
    """Optimise one scalar s per block by minimising the summed EKF filter NLL (constant median R on
    the cropped frames).  Writes into ``s_finals`` in place, as the reference does."""
    dev = require_cuda()
    dtype = get_precision()
    model, yv, K, T, O = _stage(ys, m0s, S0s, As, Cs, Qs, h_fn_combined, dev, dtype)
    structure = ops.STRUCT_DIAG if _is_diag_model(m0s, S0s, As, Cs, Qs, h_fn_combined) else ops.STRUCT_GENERAL
    if not blocks:
        blocks = [[k] for k in range(K)]
    R = _to_device(Rs, dtype, dev)
    var_planes = torch.diagonal(R, dim1=-2, dim2=-1).permute(0, 2, 1).contiguous()   # [K][O][T]
    vv = PlaneView(var_planes, O * T, [o * T for o in range(O)])
    spans = normalize_spans(T, s_frames)
    guess = torch.as_tensor(np.asarray(s_guess_per_k, dtype=np.float64), device=dev)
    s0 = np.log(np.clip(np.asarray(s_guess_per_k, dtype=np.float64), 1e-6, 1e3)).astype(np.float32)
    s_log0 = torch.as_tensor(s0, device=dev).to(dtype)
    _optimize_on_device(model, yv, vv, T, spans, blocks, guess, s_log0, lr, s_bounds_log, tol, safety_cap,
                        min_R_var, s_finals, structure=structure)
