"""Device-level operators: thin, typed wrappers over the C ABI (include/eks_b200.h).

Every function takes CUDA tensors (allocated by torch: plumbing only), enqueues kernels on the current
stream and returns device tensors.  Nothing here computes on the CPU.
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from eks_b200 import _lib
from eks_b200._lib import check, dt_code, i32_host, i64_host, lib, ptr, stream_ptr

# plane order of the singlecam / per-camera output block: the reference's 9 output columns
# (eks/singlecam_smoother.py:231-234, eks/multicam_smoother.py:515-520)
OUT_COLS = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var',
            'x_posterior_var', 'y_posterior_var']
# where the ensemble kernel's (x_avg, y_avg, var_x, var_y, likelihood) land inside that block
ENS_TO_OUT = [3, 4, 5, 6, 2]
STRUCT_GENERAL, STRUCT_DIAG, STRUCT_DIAG_STREAM = 0, 1, 2
# number of kernels of libeks_b200.so launched through this module (bench.py reports it)
LAUNCH_COUNT = 0


def _count(n: int) -> None:
    global LAUNCH_COUNT
    LAUNCH_COUNT += n


@dataclass
class PlaneView:
    """(base tensor, seq_stride, chan_off): element (b, o, t) = base.flatten()[b*seq_stride + chan_off[o] + t]."""
    base: torch.Tensor
    seq_stride: int
    chan_off: list

    @property
    def n_chan(self) -> int:
        return len(self.chan_off)


@dataclass
class Model:
    """Per-sequence state-space model (params_nlgssm_for_keypoint, eks/core.py:136-155)."""
    m0: torch.Tensor            # (B, D)
    S0: torch.Tensor            # (B, D, D)
    A: torch.Tensor             # (B, D, D)
    Q: torch.Tensor             # (B, D, D)
    C: torch.Tensor | None      # (B, O, D)  linear emission
    cams: torch.Tensor | None = None  # (ncam, 29) pinhole emission

    @property
    def B(self) -> int:
        return self.m0.shape[0]

    @property
    def D(self) -> int:
        return self.m0.shape[1]

    @property
    def ncam(self) -> int:
        return 0 if self.cams is None else self.cams.shape[0]

    def O(self) -> int:
        return 2 * self.ncam if self.cams is not None else self.C.shape[1]

    def cast(self, dtype, device) -> 'Model':
        f = lambda t: None if t is None else t.to(device=device, dtype=dtype).contiguous()
        return Model(f(self.m0), f(self.S0), f(self.A), f(self.Q), f(self.C), f(self.cams))


def _spans(spans, T):
    if not spans:
        return 0, None, None
    s0 = i32_host([a for a, _ in spans])
    s1 = i32_host([b for _, b in spans])
    return len(spans), s0, s1


def ensemble_stats(raw: torch.Tensor, out: torch.Tensor, sess_stride: int, cam_stride: int, kp_stride: int,
                   plane_off, avg_mode: str = 'median', var_mode: str = 'confidence_weighted_var',
                   nan_replacement: float = 1000.0, moments: bool = False):
    """raw (S,M,V,T,K,3) device tensor -> 5 planes per (s,v,k) written into `out` (see eks_b200.h).

    Returns the per-tile moment partials tensor (or None)."""
    assert raw.is_cuda and raw.is_contiguous() and raw.dim() == 6 and raw.shape[-1] == 3
    S, M, V, T, K, _ = raw.shape
    TT = lib().eks_ensemble_tile_frames(M, K, dt_code(raw.dtype), dt_code(out.dtype))
    ntiles = (T + TT - 1) // TT
    partials = torch.empty((S * V * K, ntiles, 4), dtype=torch.float64, device=raw.device) if moments else None
    po = i64_host(plane_off)
    vm = 1 if var_mode in ('conf_weighted_var', 'confidence_weighted_var') else 0
    check(lib().eks_ensemble_stats(ptr(raw), dt_code(raw.dtype), M * V * T * K * 3, S, M, V, T, K,
                                   int(avg_mode == 'median'), vm, float(nan_replacement), ptr(out),
                                   dt_code(out.dtype), sess_stride, cam_stride, kp_stride, ptr(po), ptr(partials),
                                   stream_ptr()), 'eks_ensemble_stats')
    _count(1)
    return partials


def center_moments(partials: torch.Tensor, T: int, dtype):
    n_seq = partials.shape[0]
    mean = torch.empty((n_seq, 2), dtype=dtype, device=partials.device)
    var = torch.empty((n_seq, 2), dtype=dtype, device=partials.device)
    check(lib().eks_center_moments(ptr(partials), n_seq, partials.shape[1], T, ptr(mean), ptr(var), dt_code(dtype),
                                   stream_ptr()),
          'eks_center_moments')
    _count(1)
    return mean, var


def initial_guess(var: PlaneView, B: int, T: int):
    """-> (guess (B,) float64, s_log0 (B,) real) on device."""
    dev, dtype = var.base.device, var.base.dtype
    guess = torch.empty(B, dtype=torch.float64, device=dev)
    s_log0 = torch.empty(B, dtype=dtype, device=dev)
    off = i64_host(var.chan_off)
    check(lib().eks_initial_guess(ptr(var.base), var.seq_stride, ptr(off), dt_code(dtype), B, var.n_chan, T,
                                  ptr(guess), ptr(s_log0), stream_ptr()), 'eks_initial_guess')
    _count(1)
    return guess, s_log0


def const_R_median(var: PlaneView, B: int, T: int, spans=None, min_var: float = 1e-4) -> torch.Tensor:
    dev, dtype = var.base.device, var.base.dtype
    O = var.n_chan
    out = torch.empty((B, O), dtype=dtype, device=dev)
    nbytes = lib().eks_const_R_median_workspace_bytes(dt_code(dtype), B, O, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    off = i64_host(var.chan_off)
    n, s0, s1 = _spans(spans, T)
    check(lib().eks_const_R_median(ptr(var.base), var.seq_stride, ptr(off), dt_code(dtype), B, O, T, n, ptr(s0),
                                   ptr(s1), float(min_var), ptr(out), ptr(ws), nbytes, stream_ptr()),
          'eks_const_R_median')
    _count(int(lib().eks_last_launch_count()))
    return out


def nll_grad(model: Model, y: PlaneView, T: int, Rconst: torch.Tensor, s: torch.Tensor, ymean=None, spans=None):
    dtype, dev = model.m0.dtype, model.m0.device
    B, D, O = model.B, model.D, model.O()
    nll = torch.empty(B, dtype=dtype, device=dev)
    dn = torch.empty(B, dtype=dtype, device=dev)
    off = i64_host(y.chan_off)
    n, s0, s1 = _spans(spans, T)
    check(lib().eks_nll_grad(dt_code(dtype), B, D, O, T, ptr(model.m0), ptr(model.S0), ptr(model.A), ptr(model.Q),
                             ptr(model.C), model.ncam, ptr(model.cams), ptr(y.base), y.seq_stride, ptr(off),
                             ptr(ymean), ptr(Rconst), n, ptr(s0), ptr(s1), ptr(s), ptr(nll), ptr(dn),
                             stream_ptr()), 'eks_nll_grad')
    _count(1)
    return nll, dn


def optimize_s(model: Model, y: PlaneView, T: int, Rconst: torch.Tensor, s_log0: torch.Tensor, blocks=None,
               ymean=None, spans=None, lr=0.25, s_bounds_log=(-8.0, 8.0), tol=1e-2, safety_cap=300,
               trace_cap: int = 0, structure: int = 0):
    """Device-resident Adam loop.  Returns dict of device tensors: s_log, loss, iters (+trace)."""
    dtype, dev = model.m0.dtype, model.m0.device
    B, D, O = model.B, model.D, model.O()
    if not blocks:
        # singleton blocks: built on the device (a pageable host->device copy would synchronise the stream with the
        # host and serialise the session groups of pipeline.singlecam_smooth_sessions)
        blocks = None
        nb = B
        d_boff = torch.arange(nb + 1, dtype=torch.int32, device=dev)
        d_mem = torch.arange(nb, dtype=torch.int32, device=dev)
    else:
        nb = len(blocks)
        boff = np.zeros(nb + 1, dtype=np.int32)
        boff[1:] = np.cumsum([len(b) for b in blocks])
        members = np.asarray([k for b in blocks for k in b], dtype=np.int32)
        d_boff = torch.from_numpy(boff).to(dev)
        d_mem = torch.from_numpy(members).to(dev)
    s_log = torch.empty(nb, dtype=dtype, device=dev)
    loss = torch.empty(nb, dtype=dtype, device=dev)
    iters = torch.empty(nb, dtype=torch.int32, device=dev)
    trace = torch.full((nb, trace_cap, 3), float('nan'), dtype=dtype, device=dev) if trace_cap else None
    nbytes = lib().eks_optimize_s_workspace_bytes(dt_code(dtype), nb, B, D, O, T)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    off = i64_host(y.chan_off)
    n, s0, s1 = _spans(spans, T)
    check(lib().eks_optimize_s(dt_code(dtype), B, D, O, T, ptr(model.m0), ptr(model.S0), ptr(model.A),
                               ptr(model.Q), ptr(model.C), model.ncam, ptr(model.cams), ptr(y.base), y.seq_stride,
                               ptr(off), ptr(ymean), ptr(Rconst), n, ptr(s0), ptr(s1), nb, ptr(d_boff), ptr(d_mem),
                               ptr(s_log0), float(lr), float(s_bounds_log[0]), float(s_bounds_log[1]), float(tol),
                               int(safety_cap), ptr(s_log), ptr(loss), ptr(iters), ptr(trace), int(trace_cap),
                               int(structure), ptr(ws), nbytes, stream_ptr()), 'eks_optimize_s')
    launches = int(lib().eks_last_launch_count())   # the library reports what this call enqueued
    _count(launches)
    unverified = int(lib().eks_last_unverified_count())
    if unverified:
        import logging
        logging.getLogger(__name__).warning(
            f'optimize_s: {unverified} evaluation(s) were accepted with a run boundary that still disagreed at the '
            f'float32 warm-up cap (rounding-level disagreement of ill-conditioned covariances)')
    return dict(s_log=s_log, loss=loss, iters=iters, trace=trace, launches=launches, unverified=unverified,
                blocks=blocks if blocks is not None else [[k] for k in range(B)], _keep=(d_boff, d_mem, ws))


def filter_smooth(model: Model, y: PlaneView, var: PlaneView, T: int, s: torch.Tensor, ymean=None):
    """-> ms (B,T,D), Vs (B,T,D,D) device tensors (reference return layout, eks/core.py:296-297)."""
    dtype, dev = model.m0.dtype, model.m0.device
    B, D, O = model.B, model.D, model.O()
    ms = torch.empty((B, T, D), dtype=dtype, device=dev)
    Vs = torch.empty((B, T, D, D), dtype=dtype, device=dev)
    nbytes = lib().eks_filter_smooth_workspace_bytes(dt_code(dtype), B, D, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    yo, vo = i64_host(y.chan_off), i64_host(var.chan_off)
    check(lib().eks_filter_smooth(dt_code(dtype), B, D, O, T, ptr(model.m0), ptr(model.S0), ptr(model.A),
                                  ptr(model.Q), ptr(model.C), model.ncam, ptr(model.cams), ptr(y.base),
                                  y.seq_stride, ptr(yo), ptr(ymean), ptr(var.base), var.seq_stride, ptr(vo),
                                  ptr(s), ptr(ms), ptr(Vs), ptr(ws), nbytes, stream_ptr()), 'eks_filter_smooth')
    _count(4 if T >= 512 else 1)
    return ms, Vs


def diag_smooth(model: Model, y: PlaneView, var: PlaneView, T: int, s: torch.Tensor, ymean, out: torch.Tensor,
                out_seq_stride: int, out_off, latent_out: bool = False, exact_scan: bool = False):
    """Decoupled (singlecam) final pass: filter + RTS + reprojection straight into the output planes.

    out_off: element offsets (within a sequence's block) of [x plane, y plane, x post-var plane,
    y post-var plane]."""
    dtype, dev = model.m0.dtype, model.m0.device
    B = model.B
    nbytes = lib().eks_diag_smooth_workspace_bytes(dt_code(dtype), B, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    yo, vo, oo = i64_host(y.chan_off), i64_host(var.chan_off), i64_host(out_off)
    check(lib().eks_diag_smooth(dt_code(dtype), B, T, ptr(model.m0), ptr(model.S0), ptr(model.A), ptr(model.Q),
                                ptr(model.C), ptr(y.base), y.seq_stride, ptr(yo), ptr(ymean), ptr(var.base),
                                var.seq_stride, ptr(vo), ptr(s), ptr(out), out_seq_stride, ptr(oo),
                                int(latent_out) | (2 if exact_scan else 0), ptr(ws), nbytes, stream_ptr()),
          'eks_diag_smooth')
    _count(int(lib().eks_last_launch_count()))
    return ws


def reproject(ms: torch.Tensor, Vs: torch.Tensor, V: int, out: torch.Tensor, out_seq_stride: int,
              out_cam_stride: int, plane_off, C=None, ymean=None, cams=None, var: PlaneView | None = None,
              pinhole_var_quirk: bool = False):
    """Smoothed latent moments -> per-camera planes (x, y, posterior var x, posterior var y)."""
    B, T, D = ms.shape
    po = i64_host(plane_off)
    vo = i64_host(var.chan_off) if var is not None else None
    check(lib().eks_reproject(dt_code(ms.dtype), B, T, D, V, ptr(ms), ptr(Vs), ptr(C), ptr(ymean),
                              0 if cams is None else cams.shape[0], ptr(cams),
                              ptr(var.base) if var is not None else None, var.seq_stride if var is not None else 0,
                              ptr(vo), int(pinhole_var_quirk), ptr(out), out_seq_stride, out_cam_stride, ptr(po),
                              stream_ptr()), 'eks_reproject')
    _count(1)


def pupil_optimize(m0, S0, C, var3, y: PlaneView, var: PlaneView, T: int, ymean=None, spans=None, lr=5e-3,
                   tol=1e-6, safety_cap=5000, trace_cap: int = 0):
    """IBL pupil AR(1) parameter optimisation for B sessions.  Returns dict(u, s, loss, iters, trace)."""
    dtype, dev = m0.dtype, m0.device
    B = m0.shape[0]
    u = torch.empty((B, 2), dtype=dtype, device=dev)
    s = torch.empty((B, 2), dtype=dtype, device=dev)
    loss = torch.empty(B, dtype=dtype, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    trace = torch.full((B, trace_cap, 3), float('nan'), dtype=dtype, device=dev) if trace_cap else None
    nbytes = lib().eks_pupil_optimize_workspace_bytes(dt_code(dtype), B, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    yo, vo = i64_host(y.chan_off), i64_host(var.chan_off)
    n, s0, s1 = _spans(spans, T)
    check(lib().eks_pupil_optimize(dt_code(dtype), B, T, ptr(m0), ptr(S0), ptr(C), ptr(var3), ptr(y.base),
                                   y.seq_stride, ptr(yo), ptr(ymean), ptr(var.base), var.seq_stride, ptr(vo), n,
                                   ptr(s0), ptr(s1), float(lr), float(tol), int(safety_cap), ptr(u), ptr(s),
                                   ptr(loss), ptr(iters), ptr(trace), int(trace_cap), ptr(ws), nbytes,
                                   stream_ptr()), 'eks_pupil_optimize')
    _count(1 + 2 * int(iters.max().item()))
    return dict(u=u, s=s, loss=loss, iters=iters, trace=trace)


# ----------------------------------------------------------------------------- multi-camera pre-stage
def mc_center(y: PlaneView, var: PlaneView, S: int, K: int, T: int, quantile_keep: float):
    """center_predictions on the device -> (ymean (B,O), n_good (B,2) int32 [good frames, frames used], workspace)."""
    dtype, dev = y.base.dtype, y.base.device
    B, O = S * K, y.n_chan
    ymean = torch.empty((B, O), dtype=dtype, device=dev)
    n_good = torch.empty((B, 2), dtype=torch.int32, device=dev)
    nbytes = lib().eks_mc_prestage_workspace_bytes(dt_code(dtype), B, O, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    yo, vo = i64_host(y.chan_off), i64_host(var.chan_off)
    check(lib().eks_mc_center(dt_code(dtype), S, K, O, T, ptr(y.base), y.seq_stride, ptr(yo), ptr(var.base),
                              var.seq_stride, ptr(vo), float(quantile_keep), ptr(ymean), ptr(n_good), ptr(ws), nbytes,
                              stream_ptr()), 'eks_mc_center')
    _count(2 + (6 if dtype == torch.float32 else 12) + 4)
    return ymean, n_good, ws


def mc_pca_moments(y: PlaneView, ymean: torch.Tensor, T: int, ws: torch.Tensor) -> torch.Tensor:
    """-> (B, 1 + O + O*O) float64: n, sum x, sum x x^T of the centred variance-filtered frames."""
    B, O = ymean.shape
    mom = torch.empty((B, 1 + O + O * O), dtype=torch.float64, device=ymean.device)
    yo = i64_host(y.chan_off)
    check(lib().eks_mc_pca_moments(dt_code(ymean.dtype), B, O, T, ptr(y.base), y.seq_stride, ptr(yo), ptr(ymean),
                                   ptr(mom), ptr(ws), ws.numel(), stream_ptr()), 'eks_mc_pca_moments')
    _count(2)
    return mom


def mc_latent_init(y: PlaneView, ymean: torch.Tensor, pca_mean: torch.Tensor, comps: torch.Tensor, T: int,
                   ws: torch.Tensor):
    """comps (B,O,L) = pca.components_.T -> (S0 (B,L,L), Q (B,L,L))."""
    B, O, L = comps.shape
    dtype, dev = ymean.dtype, ymean.device
    S0 = torch.empty((B, L, L), dtype=dtype, device=dev)
    Q = torch.empty((B, L, L), dtype=dtype, device=dev)
    yo = i64_host(y.chan_off)
    check(lib().eks_mc_latent_init(dt_code(dtype), B, O, L, T, ptr(y.base), y.seq_stride, ptr(yo), ptr(ymean),
                                   ptr(pca_mean), ptr(comps), ptr(S0), ptr(Q), ptr(ws), ws.numel(), stream_ptr()),
          'eks_mc_latent_init')
    _count(2)
    return S0, Q


def mc_valid_moments(y: PlaneView, ymean: torch.Tensor, var: PlaneView, T: int, ws: torch.Tensor, v_quantile=50.0,
                     lik: PlaneView | None = None, lik_threshold: float = 0.9, active: torch.Tensor | None = None):
    """Moments (n, sum x, sum x x^T) of the rows that feed the FactorAnalysis fit -> (B, 1 + O + O*O) float64."""
    B, O = ymean.shape
    mom = torch.zeros((B, 1 + O + O * O), dtype=torch.float64, device=ymean.device)
    yo, vo = i64_host(y.chan_off), i64_host(var.chan_off)
    lo = i64_host(lik.chan_off) if lik is not None else None
    check(lib().eks_mc_valid_moments(dt_code(ymean.dtype), B, O // 2, T, ptr(y.base), y.seq_stride, ptr(yo), ptr(ymean),
                                     ptr(var.base), var.seq_stride, ptr(vo), ptr(lik.base) if lik is not None else None,
                                     lik.seq_stride if lik is not None else 0, ptr(lo), float(lik_threshold),
                                     -1.0 if v_quantile is None else float(v_quantile), ptr(active), ptr(mom),
                                     ptr(ws), ws.numel(), stream_ptr()), 'eks_mc_valid_moments')
    _count(2 + (1 + (6 if ymean.dtype == torch.float32 else 12) if v_quantile is not None else 0))
    return mom


def mc_inflate_step(y: PlaneView, ymean: torch.Tensor, var: PlaneView, T: int, loading: torch.Tensor,
                    mean: torch.Tensor, epsilon=1e-6, threshold=5.0, scalar=10.0, active: torch.Tensor | None = None):
    """One Mahalanobis inflation pass, variances updated in place -> flags (B,) int32 (1 = something inflated)."""
    B, O, L = loading.shape
    flags = torch.empty(B, dtype=torch.int32, device=ymean.device)
    yo, vo = i64_host(y.chan_off), i64_host(var.chan_off)
    check(lib().eks_mc_inflate_step(dt_code(ymean.dtype), B, O // 2, L, T, ptr(y.base), y.seq_stride, ptr(yo),
                                    ptr(ymean), ptr(var.base), var.seq_stride, ptr(vo), ptr(loading), ptr(mean),
                                    float(epsilon), float(threshold), float(scalar), ptr(active), ptr(flags),
                                    stream_ptr()), 'eks_mc_inflate_step')
    _count(1)
    return flags


def geometric_init(tri: torch.Tensor):
    """initialize_kalman_filter_geometric (eks/multicam_smoother.py:600-650) on the device.
    tri: (B, T, 3) float64 CUDA tensor of triangulated ensemble means -> (m0 (B,3), S0_diag (B,3), Q_diag (B,3)) float64."""
    assert tri.is_cuda and tri.dtype == torch.float64 and tri.dim() == 3 and tri.shape[-1] == 3
    tri = tri.contiguous()
    B, T, _ = tri.shape
    dev = tri.device
    m0 = torch.empty((B, 3), dtype=torch.float64, device=dev)
    S0d = torch.empty((B, 3), dtype=torch.float64, device=dev)
    Qd = torch.empty((B, 3), dtype=torch.float64, device=dev)
    nbytes = lib().eks_geometric_init_workspace_bytes(B, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib().eks_geometric_init(B, T, ptr(tri), ptr(m0), ptr(S0d), ptr(Qd), ptr(ws), nbytes, stream_ptr()),
          'eks_geometric_init')
    _count(int(lib().eks_last_launch_count()))
    return m0, S0d, Qd


def triangulate_mean(raw: torch.Tensor, cams: torch.Tensor) -> torch.Tensor:
    """raw (M,V,T,K,3) device tensor (pixels), cams (V,29) float64 -> (K,T,3) float64: mean over the ensemble of
    the triangulated points (triangulate_3d_models(...).mean(axis=0))."""
    assert raw.is_cuda and raw.is_contiguous() and raw.dim() == 5 and raw.shape[-1] == 3
    M, V, T, K, _ = raw.shape
    cams = cams.to(device=raw.device, dtype=torch.float64).contiguous()
    out = torch.empty((K, T, 3), dtype=torch.float64, device=raw.device)
    check(lib().eks_triangulate_mean(ptr(raw), dt_code(raw.dtype), M, V, T, K, ptr(cams), ptr(out), stream_ptr()),
          'eks_triangulate_mean')
    _count(1)
    return out
