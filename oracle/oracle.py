"""CPU ORACLE for the EKS smoothing hot path -- TEST INFRASTRUCTURE ONLY.

Thin ctypes wrapper over ``oracle/liboracle.so`` (built from ``oracle/eks_oracle.cpp``) plus
NumPy restatements of the small host-side steps of the reference path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product package ``eks_b200`` never does.

PARITY STATUS: **parity unpinned** at the dynamax boundary (dynamax<=1.0.1 / jax<=0.4.36 / optax
are not installable here, see the header of eks_oracle.cpp).  Every function cites the reference
file:line it restates (paths relative to /root/reference).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'liboracle.so')
_lib = None

CAM_STRIDE = 29


def build(force: bool = False) -> str:
    """Compile liboracle.so with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, 'eks_oracle.cpp')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, 'liboracle.so'], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return 'f32'
    if dtype == np.float64:
        return 'f64'
    raise TypeError(f'oracle supports float32/float64, got {dtype}')


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _real(dtype, x):
    return (ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double)(float(x))


# ----------------------------------------------------------------------------- cameras
def rodrigues(rvec) -> np.ndarray:
    """rvec (3,) -> R (3,3); restates eks/multicam_smoother.py:771-793 (OpenCV convention)."""
    rvec = np.asarray(rvec, dtype=np.float64).ravel()
    theta = np.linalg.norm(rvec)
    if theta < 1e-12:
        rx, ry, rz = rvec
        Km = np.array([[0.0, -rz, ry], [rz, 0.0, -rx], [-ry, rx, 0.0]])
        return np.eye(3) + Km
    rx, ry, rz = rvec / theta
    Km = np.array([[0.0, -rz, ry], [rz, 0.0, -rx], [-ry, rx, 0.0]])
    return np.eye(3) + np.sin(theta) * Km + (1.0 - np.cos(theta)) * (Km @ Km)


def pack_camera(rvec_or_R, tvec, Kmat, dist) -> np.ndarray:
    """Pack one camera into the 29-real layout used by the oracle.

    dist follows OpenCV ordering [k1,k2,p1,p2,k3,k4,k5,k6,s1,s2,s3,s4,(tx,ty ignored)], padded with
    zeros (parse_dist, eks/multicam_smoother.py:796-803).
    """
    r = np.asarray(rvec_or_R, dtype=np.float64)
    Rm = r if r.shape == (3, 3) else rodrigues(r)
    Kmat = np.asarray(Kmat, dtype=np.float64)
    d = np.zeros(14)
    dist = np.asarray(dist, dtype=np.float64).ravel()
    d[:min(14, dist.size)] = dist[:14]
    out = np.zeros(CAM_STRIDE)
    out[0:9] = Rm.ravel()
    out[9:12] = np.asarray(tvec, dtype=np.float64).ravel()
    out[12:17] = [Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2], Kmat[0, 1]]
    out[17:29] = d[:12]
    return out


def project(cams, X, dtype=np.float64, jac=False):
    """h(x): (N,3) -> (N,2V) and optionally its Jacobian (N,2V,3)."""
    cams = _c(np.asarray(cams).reshape(-1, CAM_STRIDE), dtype)
    X = _c(np.asarray(X).reshape(-1, 3), dtype)
    V, N = cams.shape[0], X.shape[0]
    uv = np.empty((N, 2 * V), dtype=dtype)
    J = np.empty((N, 2 * V, 3), dtype=dtype) if jac else None
    getattr(lib(), f'eks_oracle_project_{_sfx(dtype)}')(V, _p(cams), N, _p(X), _p(uv), _p(J))
    return (uv, J) if jac else uv


# ----------------------------------------------------------------------------- ensemble
def ensemble(raw, avg_mode='median', var_mode='confidence_weighted_var', nan_replacement=1000.0,
             dtype=np.float32) -> np.ndarray:
    """raw (M,V,T,K,3) -> (V,T,K,5) [x,y,var_x,var_y,likelihood]; eks/core.py:25-101."""
    raw = _c(raw, dtype)
    M, V, T, K, F = raw.shape
    assert F == 3 and M <= 64
    out = np.empty((V, T, K, 5), dtype=dtype)
    vm = 1 if var_mode in ('conf_weighted_var', 'confidence_weighted_var') else 0
    getattr(lib(), f'eks_oracle_ensemble_{_sfx(dtype)}')(
        _p(raw), M, V, T, K, int(avg_mode == 'median'), vm, _real(dtype, nan_replacement), _p(out))
    return out


# ----------------------------------------------------------------------------- small host steps
def crop_frames(y, s_frames):
    """eks/utils.py:235-290 restated (validation + concatenation of [start,end) spans)."""
    n = len(y)
    if s_frames is None or len(s_frames) == 0 or (len(s_frames) == 1 and s_frames[0] == (None, None)):
        return y
    if not isinstance(s_frames, list):
        raise TypeError('s_frames must be a list of (start, end) tuples or None.')
    spans = []
    for i, fr in enumerate(s_frames):
        if not (isinstance(fr, tuple) and len(fr) == 2):
            raise ValueError(f's_frames[{i}] must be a (start, end) tuple, got {fr!r}')
        a, b = fr
        if a is not None and not isinstance(a, int):
            raise ValueError(f's_frames[{i}].start must be int or None, got {a!r}')
        if b is not None and not isinstance(b, int):
            raise ValueError(f's_frames[{i}].end must be int or None, got {b!r}')
        a = 0 if a is None else a
        b = n if b is None else b
        if a < 0 or b > n:
            raise ValueError(f'Range ({a}, {b}) out of bounds for length {n}.')
        if a >= b:
            raise ValueError(f'Invalid range ({a}, {b}).')
        spans.append((a, b))
    spans.sort(key=lambda s: s[0])
    for i in range(1, len(spans)):
        if spans[i][0] < spans[i - 1][1]:
            raise ValueError(f'Overlapping or out-of-order intervals: {spans[i - 1]} and {spans[i]}')
    return np.concatenate([y[a:b] for a, b in spans], axis=0)


def compute_initial_guess(ensemble_vars_k) -> float:
    """eks/core.py:104-133 + caller fallback core.py:235-236.  ensemble_vars_k: (T, obs)."""
    ev = np.asarray(ensemble_vars_k)[:2000]
    if ev.shape[0] < 2:
        raise ValueError('Not enough frames to compute temporal differences.')
    d = ev[1:] - ev[:-1]
    g = float(round(np.nanstd(d), 5)) or 2.0
    return g if (np.isfinite(g) and g > 0.0) else 2.0


def constant_R(var_cropped, min_var=1e-4) -> np.ndarray:
    """(T',obs) variances (already clipped at 1e-12) -> (obs,) ; eks/core.py:702-709."""
    med = np.nanmedian(var_cropped, axis=0)
    return np.clip(med, min_var, np.inf).astype(var_cropped.dtype)


# ----------------------------------------------------------------------------- kernels of the path
def _stack_model(m0s, S0s, As, Qs, Cs, cams, dtype):
    m0s = _c(m0s, dtype)
    K, D = m0s.shape
    S0s = _c(S0s, dtype).reshape(K, D, D)
    As = _c(As, dtype).reshape(K, D, D)
    Qs = _c(Qs, dtype).reshape(K, D, D)
    if cams is not None:
        cams = _c(np.asarray(cams).reshape(-1, CAM_STRIDE), dtype)
        ncam = cams.shape[0]
        O = 2 * ncam
        Cs_ = None
        assert D == 3
    else:
        Cs_ = _c(Cs, dtype)
        O = Cs_.shape[1]
        ncam = 0
    assert D <= lib().eks_oracle_dmax() and O <= lib().eks_oracle_omax()
    return K, D, O, m0s, S0s, As, Qs, Cs_, ncam, cams


def nll_grad(ys, m0s, S0s, As, Cs, Qs, Rdiag, s, cams=None, dtype=np.float64):
    """Filter NLL and d NLL / d s per sequence.  ys (K,T,O); Rdiag (K,O) constant or (K,T,O)."""
    K, D, O, m0s, S0s, As, Qs, Cs_, ncam, cams = _stack_model(m0s, S0s, As, Qs, Cs, cams, dtype)
    ys = _c(ys, dtype)
    T = ys.shape[1]
    Rdiag = _c(Rdiag, dtype)
    tv = int(Rdiag.ndim == 3)
    s = _c(np.broadcast_to(np.asarray(s, dtype=dtype), (K,)), dtype)
    nll = np.empty(K, dtype=dtype)
    dn = np.empty(K, dtype=dtype)
    getattr(lib(), f'eks_oracle_nll_grad_{_sfx(dtype)}')(
        K, D, O, _p(m0s), _p(S0s), _p(As), _p(Qs), _p(Cs_), ncam, _p(cams), _p(ys), _p(Rdiag), tv, T,
        _p(s), _p(nll), _p(dn))
    return nll, dn


def optimize(ys, m0s, S0s, As, Cs, Qs, Rconst, blocks, s_guess_per_k, lr=0.25, s_bounds_log=(-8.0, 8.0),
             tol=1e-3, safety_cap=300, cams=None, dtype=np.float32, trace_cap=0):
    """Adam on log s per block; eks/core.py:306-401, 403-559, 562-699.

    ys (K,T',O) cropped; Rconst (K,O).  Returns dict(s_finals(K,), s_log(B,), loss(B,), iters(B,),
    trace (B,trace_cap,3) [s_log, loss, lr*grad])."""
    K, D, O, m0s, S0s, As, Qs, Cs_, ncam, cams = _stack_model(m0s, S0s, As, Qs, Cs, cams, dtype)
    ys = _c(ys, dtype)
    T = ys.shape[1]
    Rconst = _c(Rconst, dtype)
    if not blocks:
        blocks = [[k] for k in range(K)]
    order = [k for b in blocks for k in b]
    off = np.zeros(len(blocks) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(b) for b in blocks])
    idx = np.asarray(order, dtype=int)
    s0 = np.array([np.clip(np.mean([s_guess_per_k[k] for k in b]), 1e-6, 1e3) for b in blocks])
    s_log0 = np.log(s0).astype(np.float32).astype(dtype)  # core.py:441, 622 -- float32 seed
    B = len(blocks)
    s_log = np.empty(B, dtype=dtype)
    loss = np.empty(B, dtype=dtype)
    iters = np.empty(B, dtype=np.int32)
    trace = np.full((B, max(trace_cap, 1), 3), np.nan, dtype=dtype) if trace_cap else None
    getattr(lib(), f'eks_oracle_optimize_{_sfx(dtype)}')(
        B, _p(off), D, O, _p(_c(m0s[idx], dtype)), _p(_c(S0s[idx], dtype)), _p(_c(As[idx], dtype)),
        _p(_c(Qs[idx], dtype)), _p(None if Cs_ is None else _c(Cs_[idx], dtype)), ncam, _p(cams),
        _p(_c(ys[idx], dtype)), _p(_c(Rconst[idx], dtype)), T, _p(s_log0), _real(dtype, lr),
        _real(dtype, s_bounds_log[0]), _real(dtype, s_bounds_log[1]), _real(dtype, tol), int(safety_cap),
        _p(s_log), _p(loss), _p(iters), _p(trace), int(trace_cap))
    s_finals = np.empty(K, dtype=float)
    for b, blk in enumerate(blocks):
        s_star = float(np.exp(np.clip(s_log[b], s_bounds_log[0], s_bounds_log[1])))  # core.py:550, 694
        for k in blk:
            s_finals[k] = s_star
    return dict(s_finals=s_finals, s_log=s_log, loss=loss, iters=iters, trace=trace, s_log0=s_log0)


def smooth(ys, m0s, S0s, As, Cs, Qs, Rdiag, s, cams=None, dtype=np.float32, return_filtered=False):
    """EKF filter + RTS smoother per sequence; eks/core.py:274-295.  Rdiag (K,T,O) or (K,O)."""
    K, D, O, m0s, S0s, As, Qs, Cs_, ncam, cams = _stack_model(m0s, S0s, As, Qs, Cs, cams, dtype)
    ys = _c(ys, dtype)
    T = ys.shape[1]
    Rdiag = _c(Rdiag, dtype)
    tv = int(Rdiag.ndim == 3)
    s = _c(np.broadcast_to(np.asarray(s, dtype=dtype), (K,)), dtype)
    ms = np.empty((K, T, D), dtype=dtype)
    Vs = np.empty((K, T, D, D), dtype=dtype)
    mfs = np.empty((K, T, D), dtype=dtype) if return_filtered else None
    Pfs = np.empty((K, T, D, D), dtype=dtype) if return_filtered else None
    ll = np.empty(K, dtype=dtype)
    bad = np.zeros(K, dtype=np.int32)
    getattr(lib(), f'eks_oracle_smooth_{_sfx(dtype)}')(
        K, D, O, _p(m0s), _p(S0s), _p(As), _p(Qs), _p(Cs_), ncam, _p(cams), _p(ys), _p(Rdiag), tv, T, _p(s),
        _p(ms), _p(Vs), _p(mfs), _p(Pfs), _p(ll), _p(bad))
    if return_filtered:
        return ms, Vs, mfs, Pfs, ll
    return ms, Vs


# ----------------------------------------------------------------------------- the path, end to end
def run_kalman_smoother(ys, m0s, S0s, As, Cs, Qs, ensemble_vars, s_frames=None, smooth_param=None,
                        blocks=None, lr=0.25, s_bounds_log=(-8.0, 8.0), tol=1e-2, safety_cap=300,
                        cams=None, dtype=np.float32, min_R_var=1e-4, trace_cap=0):
    """eks/core.py:159-302 restated.  ys (K,T,O); ensemble_vars (T,K,O).

    Returns (s_finals (K,) float64, ms (K,T,D), Vs (K,T,D,D), info dict)."""
    ys = np.asarray(ys, dtype=dtype)
    K, T, O = ys.shape
    ev = np.asarray(ensemble_vars, dtype=dtype)
    var_kto = np.clip(np.swapaxes(ev, 0, 1), 1e-12, None)  # build_R_from_vars, utils.py:373
    guesses = np.array([compute_initial_guess(ev[:, k, :]) for k in range(K)])
    info = dict(guesses=guesses)
    if smooth_param is not None:
        s_finals = np.empty(K, dtype=float)
        if isinstance(smooth_param, (int, float)):
            s_finals[:] = float(smooth_param)
        else:
            s_finals[:] = np.asarray(smooth_param, dtype=float)
    else:
        y_c = np.stack([crop_frames(ys[k], s_frames) for k in range(K)])
        R_c = np.stack([constant_R(crop_frames(var_kto[k], s_frames), min_R_var) for k in range(K)])
        opt = optimize(y_c, m0s, S0s, As, Cs, Qs, R_c, blocks, guesses, lr=lr, s_bounds_log=s_bounds_log,
                       tol=tol, safety_cap=safety_cap, cams=cams, dtype=dtype, trace_cap=trace_cap)
        s_finals = opt['s_finals']
        info.update(opt)
        info['Rconst'] = R_c
    ms, Vs = smooth(ys, m0s, S0s, As, Cs, Qs, var_kto, s_finals, cams=cams, dtype=dtype)
    return s_finals, ms, Vs, info


def singlecam(raw, smooth_param=None, s_frames=None, blocks=None, avg_mode='median',
              var_mode='confidence_weighted_var', dtype=np.float32, trace_cap=0):
    """ensemble_kalman_smoother_singlecam restated (eks/singlecam_smoother.py:105-284).

    raw: (M,1,T,K,3).  Returns dict(out (T,K,9) float64 in the reference column order
    [x,y,likelihood,x_ens_median,y_ens_median,x_ens_var,y_ens_var,x_posterior_var,y_posterior_var],
    s_finals, info)."""
    raw = np.asarray(raw)
    M, V, T, K, _ = raw.shape
    assert V == 1
    ens = ensemble(raw, avg_mode, var_mode, dtype=dtype)[0]  # (T,K,5)
    preds = ens[..., 0:2].astype(np.float64) if dtype == np.float64 else ens[..., 0:2]
    # center_predictions(quantile 100): every frame is kept (utils.py:318-343) -> mean over all frames
    means = np.mean(preds, axis=0, keepdims=True)  # (1,K,2)
    centered = preds - means
    ys = np.transpose(centered, (1, 0, 2))  # (K,T,2)
    m0s = np.zeros((K, 2))
    S0s = np.zeros((K, 2, 2))
    for k in range(K):
        S0s[k, 0, 0] = np.nanvar(centered[:, k, 0])
        S0s[k, 1, 1] = np.nanvar(centered[:, k, 1])
    eye = np.tile(np.eye(2), (K, 1, 1))
    s_finals, ms, Vs, info = run_kalman_smoother(
        ys, m0s, S0s, eye, eye, eye, ens[..., 2:4], s_frames=s_frames, smooth_param=smooth_param,
        blocks=blocks, dtype=dtype, trace_cap=trace_cap)
    out = np.empty((T, K, 9), dtype=np.float64)
    out[..., 0] = ms[:, :, 0].T.astype(np.float64) + means[0, :, 0][None, :]
    out[..., 1] = ms[:, :, 1].T.astype(np.float64) + means[0, :, 1][None, :]
    out[..., 2] = ens[..., 4]
    out[..., 3] = ens[..., 0]
    out[..., 4] = ens[..., 1]
    out[..., 5] = ens[..., 2]
    out[..., 6] = ens[..., 3]
    out[..., 7] = Vs[:, :, 0, 0].T
    out[..., 8] = Vs[:, :, 1, 1].T
    info['means'] = means
    info['S0s'] = S0s
    return dict(out=out, s_finals=s_finals, info=info, ms=ms, Vs=Vs)


def multicam(raw, quantile_keep_pca=50.0, n_latent=3, camgroup=None, smooth_param=None, s_frames=None,
             avg_mode='median', var_mode='confidence_weighted_var', dtype=np.float32, trace_cap=0,
             inflate_vars=False, inflate_vars_kwargs=None):
    """ensemble_kalman_smoother_multicam restated (eks/multicam_smoother.py:279-551), inflate_vars=False.

    raw: (M,V,T,K,3).  Linear model: centring and PCA initialisation are the oracle's own restatement
    (mc_center_predictions / mc_pca_init, NumPy + scikit-learn's PCA).  Nonlinear model (camgroup = path of an
    Anipose calibration TOML, or the list load_calibration returns): triangulation and the geometric initialisation
    are the oracle's own NumPy restatement (load_calibration / triangulate_3d_models / geometric_init below).
    Returns dict(cam_out (V,T,K,9), out3d (T,K,2*D), s_finals, info)."""
    raw = np.asarray(raw)
    M, V, T, K, _ = raw.shape
    ens = ensemble(raw, avg_mode, var_mode, dtype=dtype)                    # (V,T,K,5)
    stacked = lambda lo, hi: np.transpose(ens[..., lo:hi], (2, 1, 0, 3)).reshape(K, T, 2 * V)   # channel o = 2 v + xy
    cams = None
    if camgroup is not None:
        # triangulation + geometric initialisation: the oracle's OWN restatement (NumPy only; see the
        # "calibration" section below) -- nothing here comes from the product package
        calib = load_calibration(camgroup) if isinstance(camgroup, (str, os.PathLike)) else camgroup
        tri = triangulate_3d_models(raw, calib)                                # (M,K,T,3)
        m0s, S0s, As, Qs, Cs = geometric_init(tri.mean(axis=0))
        cams = pack_calibration(calib)
        ys = stacked(0, 2)
        D = 3
    else:
        mask_o, cen_o, good_o, means_o, _ = mc_center_predictions(ens, quantile_keep_pca)
        m0s, S0s, As, Qs, Cs = mc_pca_init(mask_o, cen_o, good_o, n_latent)
        ys = cen_o
        D = n_latent
    ev = stacked(2, 4)                                                        # (K,T,2V)
    ev_out = ev                                                               # variances of the output columns
    if inflate_vars:   # eks/multicam_smoother.py:353-361 (on the CENTRED predictions, also for the nonlinear model)
        kw = dict(inflate_vars_kwargs or {})
        cen_i = mc_center_predictions(ens, quantile_keep_pca)[1]
        likes = np.transpose(ens[..., 4], (2, 1, 0)) if kw.pop('likelihoods', None) is not None else None   # (K,T,V)
        ev = np.stack([mc_inflate_variance(cen_i[k], ev[k], n_latent=n_latent,
                                           likes=None if likes is None else likes[k], **kw)[0] for k in range(K)])
        if cams is None:
            ev_out = ev
    s_finals, ms, Vs, info = run_kalman_smoother(ys, m0s, S0s, As, Cs, Qs, np.swapaxes(ev, 0, 1),
                                                 s_frames=s_frames, smooth_param=smooth_param, cams=cams,
                                                 dtype=dtype, trace_cap=trace_cap)
    ms64, Vs64 = ms.astype(np.float64), Vs.astype(np.float64)
    cam_out = np.empty((V, T, K, 9))
    for k in range(K):
        if cams is not None:
            uv, J = project(cams, ms64[k], jac=True)                          # (T,2V), (T,2V,3)
            cov = np.einsum('tij,tjk,tlk->til', J, Vs64[k], J)
        else:
            Ck = np.asarray(Cs[k], dtype=np.float64)
            uv = ms64[k] @ Ck.T + means_o[k].astype(np.float64)[None, :]
            cov = np.einsum('ij,tjk,lk->til', Ck, Vs64[k], Ck)
        for c in range(V):
            cam_out[c, :, k, 0] = uv[:, 2 * c]
            cam_out[c, :, k, 1] = uv[:, 2 * c + 1]
            cam_out[c, :, k, 2] = ens[c, :, k, 4]
            cam_out[c, :, k, 3:5] = ens[c, :, k, 0:2]
            cam_out[c, :, k, 5:7] = ev_out[k][:, 2 * c:2 * c + 2]       # linear: the inflated variances (:505-508)
            if cams is not None:  # :943-944 -- always variance columns 0 / 1
                cam_out[c, :, k, 7] = cov[:, 2 * c, 2 * c] + ev[k][:, 0]
                cam_out[c, :, k, 8] = cov[:, 2 * c + 1, 2 * c + 1] + ev[k][:, 1]
            else:
                cam_out[c, :, k, 7] = cov[:, 2 * c, 2 * c] + ev[k][:, 2 * c]
                cam_out[c, :, k, 8] = cov[:, 2 * c + 1, 2 * c + 1] + ev[k][:, 2 * c + 1]
    out3d = np.empty((T, K, 6))
    for k in range(K):
        out3d[:, k, 0:3] = ms64[k][:, 0:3]
        out3d[:, k, 3] = Vs64[k][:, 0, 0]
        out3d[:, k, 4] = Vs64[k][:, 1, 1]
        out3d[:, k, 5] = Vs64[k][:, 2, 2]
    return dict(cam_out=cam_out, out3d=out3d, s_finals=s_finals, info=info, ms=ms, Vs=Vs)




# ----------------------------------------------------------------------------- calibration, triangulation, 3-D init
# Independent NumPy restatement (no OpenCV, no product code) of the calibrated pre-stage:
#   * aniposelib.cameras.CameraGroup.load / Camera.undistort_points / CameraGroup.triangulate(fast=True)
#     (third-party, aniposelib>=0.8.0, pyproject.toml:35; call sites eks/multicam_smoother.py:233, 902): per camera
#     cv2.undistortPoints with its default criteria (five fixed-point iterations of the inverse distortion), then for
#     every camera PAIR the homogeneous DLT of cv2.triangulatePoints (null vector of the 4x4 system by SVD), and the
#     nan-median over the pairs;
#   * triangulate_3d_models (eks/multicam_smoother.py:888-911) and initialize_kalman_filter_geometric (:600-650).
# Pinned against OpenCV itself in tests/test_oracle.py (cv2.undistortPoints, cv2.triangulatePoints).
def load_calibration(path) -> list:
    """Anipose calibration TOML -> list of dicts(name, matrix (3,3), dist (n,), rvec (3,), tvec (3,)), sorted by
    section name (cam_0, cam_1, ...), which is the order CameraGroup.load keeps."""
    import tomllib
    with open(path, 'rb') as f:
        cfg = tomllib.load(f)
    cams = []
    for key in sorted(k for k in cfg if k.startswith('cam_')):
        c = cfg[key]
        cams.append(dict(name=c['name'], matrix=np.asarray(c['matrix'], dtype=np.float64),
                         dist=np.asarray(c['distortions'], dtype=np.float64).ravel(),
                         rvec=np.asarray(c['rotation'], dtype=np.float64).ravel(),
                         tvec=np.asarray(c['translation'], dtype=np.float64).ravel()))
    return cams


def pack_calibration(calib) -> np.ndarray:
    """(V, 29) packed cameras (make_projection_from_camgroup, eks/multicam_smoother.py:862-885)."""
    return np.stack([pack_camera(c['rvec'], c['tvec'], c['matrix'], c['dist']) for c in calib])


def undistort_points(pts, Kmat, dist, iters: int = 5) -> np.ndarray:
    """cv2.undistortPoints(pts, K, dist) with the default criteria restated: normalise with the camera matrix, then
    `iters` fixed-point iterations x <- (x0 - tangential(x) - prism(x)) / radial(x).  pts (N,2) pixels -> (N,2)."""
    k = np.zeros(14)
    d = np.asarray(dist, dtype=np.float64).ravel()
    k[:min(14, d.size)] = d[:14]
    Kmat = np.asarray(Kmat, dtype=np.float64)
    fx, fy, cx, cy = Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2]
    pts = np.asarray(pts, dtype=np.float64)
    x0 = (pts[:, 0] - cx) / fx
    y0 = (pts[:, 1] - cy) / fy
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icd = (1.0 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1.0 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        icd = np.where(icd < 0, 1.0, icd)            # OpenCV: a negative ratio leaves the point at x0 (never hit here)
        dx = 2.0 * k[2] * x * y + k[3] * (r2 + 2.0 * x * x) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2.0 * y * y) + 2.0 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
        x = (x0 - dx) * icd
        y = (y0 - dy) * icd
    return np.stack([x, y], axis=1)


def triangulate_pair(P1, P2, x1, x2) -> np.ndarray:
    """cv2.triangulatePoints restated: rows x*P[2]-P[0], y*P[2]-P[1] of both views, right singular vector of the
    smallest singular value, de-homogenised.  P1, P2 (3,4); x1, x2 (N,2) normalised coordinates -> (N,3)."""
    N = x1.shape[0]
    A = np.empty((N, 4, 4))
    A[:, 0] = x1[:, 0:1] * P1[2] - P1[0]
    A[:, 1] = x1[:, 1:2] * P1[2] - P1[1]
    A[:, 2] = x2[:, 0:1] * P2[2] - P2[0]
    A[:, 3] = x2[:, 1:2] * P2[2] - P2[1]
    X = np.full((N, 3), np.nan)
    ok = np.isfinite(A).all(axis=(1, 2))
    if ok.any():
        vh = np.linalg.svd(A[ok])[2][:, 3, :]
        with np.errstate(all='ignore'):
            X[ok] = vh[:, :3] / vh[:, 3:4]
    return X


def triangulate_fast(points, calib) -> np.ndarray:
    """CameraGroup.triangulate(points, fast=True): points (V,N,2) pixels -> (N,3)."""
    import itertools
    V = len(calib)
    und = [undistort_points(points[c], calib[c]['matrix'], calib[c]['dist']) for c in range(V)]
    Rt = []
    for c in calib:
        E = np.empty((3, 4))
        E[:, :3] = rodrigues(c['rvec'])
        E[:, 3] = c['tvec']
        Rt.append(E)
    tris = [triangulate_pair(Rt[a], Rt[b], und[a], und[b]) for a, b in itertools.combinations(range(V), 2)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=RuntimeWarning)
        return np.nanmedian(np.stack(tris), axis=0)


def triangulate_3d_models(raw, calib) -> np.ndarray:
    """eks/multicam_smoother.py:888-911: raw (M,V,T,K,>=2) -> (M,K,T,3), one triangulate(fast=True) per (m, k)."""
    raw = np.asarray(raw, dtype=np.float64)
    M, V, T, K, _ = raw.shape
    tri = np.zeros((M, K, T, 3))
    for m in range(M):
        for k in range(K):
            tri[m, k] = triangulate_fast(raw[m, :, :, k, :2], calib)
    return tri


def geometric_init(ys3d):
    """initialize_kalman_filter_geometric (eks/multicam_smoother.py:600-650): ys3d (K,T,3) ->
    m0 = mean of the first 10 frames, S0 = diag(nanvar + 1e-4), A = C = I, Q = diag(max((1.4826 MAD(diff))^2, 1e-8))."""
    ys3d = np.asarray(ys3d, dtype=np.float64)
    K, T, D = ys3d.shape
    m0s = ys3d[:, :10].mean(axis=1)
    S0s = np.stack([np.diag(np.nanvar(ys3d[k], axis=0) + 1e-4) for k in range(K)])
    eye = np.tile(np.eye(D), (K, 1, 1))
    Qs = np.empty((K, D, D))
    for k in range(K):
        dx = np.diff(ys3d[k], axis=0)
        med = np.median(dx, axis=0)
        mad = np.median(np.abs(dx - med), axis=0) + 1e-12
        Qs[k] = np.diag(np.maximum((1.4826 * mad) ** 2, 1e-8))
    return m0s, S0s, eye, Qs, eye.copy()


# ----------------------------------------------------------------------------- multi-camera pre-stage (linear model)
def mc_center_predictions(ens, quantile_keep_pca):
    """center_predictions restated (eks/utils.py:293-365).  ens (V,T,K,5) [x,y,var_x,var_y,lik].
    Returns (mask (T,K) bool, centered (K,T,2V), good_centered list of (n,2V), means (K,2V), n_used)."""
    V, T, K, _ = ens.shape
    max_vars = ens[..., 2:4].max(axis=(0, 3))                                  # (T,K)
    thr = np.percentile(max_vars, quantile_keep_pca, axis=0)
    mask = max_vars <= thr
    good = [np.where(mask[:, k])[0] for k in range(K)]
    n_used = min(len(g) for g in good)
    preds = np.transpose(ens[..., 0:2], (2, 1, 0, 3)).reshape(K, T, 2 * V)      # channel o = 2 v + xy
    means = np.empty((K, 2 * V), dtype=ens.dtype)
    centered = np.empty_like(preds)
    good_centered = []
    for k in range(K):
        gp = preds[k][good[k][:n_used]]
        means[k] = gp.mean(axis=0)
        centered[k] = preds[k] - means[k]
        good_centered.append(gp - means[k])
    return mask, centered, good_centered, means, n_used


def mc_pca_init(mask, centered, good_centered, n_latent=3):
    """compute_pca + initialize_kalman_filter_pca restated (eks/stats.py:9-64, eks/multicam_smoother.py:554-597)
    with scikit-learn's PCA itself (the reference's dependency).  Returns (m0s, S0s, As, Qs, Cs)."""
    from sklearn.decomposition import PCA
    K = len(good_centered)
    m0s = np.zeros((K, n_latent))
    As = np.tile(np.eye(n_latent), (K, 1, 1))
    S0s, Qs, Cs = [], [], []
    for k in range(K):
        pca = PCA(n_components=n_latent).fit(good_centered[k])
        pcs = pca.transform(centered[k])[np.where(mask[:, k])[0]]
        S0s.append(np.diag(pcs.var(axis=0)))
        cov = np.atleast_2d(np.cov(np.diff(pcs, axis=0).T))
        mx = np.abs(cov).max()
        Qs.append(cov / mx if mx > 0 else cov)
        Cs.append(pca.components_.T)
    return m0s, np.stack(S0s), As, np.stack(Qs), np.stack(Cs)


def mc_inflate_variance(centered, evars, n_latent=3, likes=None, likelihood_threshold=0.9, v_quantile_threshold=50.0,
                        epsilon=1e-6, loading_matrix=None, mean=None, threshold=5.0, scalar=10.0):
    """mA_compute_maha + compute_mahalanobis + inflate_variance restated for ONE keypoint
    (eks/multicam_smoother.py:653-764, eks/stats.py:67-157), with scikit-learn's FactorAnalysis itself and the
    per-frame loops written as batched float64 NumPy.  centered, evars: (T, 2V).  Returns the inflated variances."""
    from sklearn.decomposition import FactorAnalysis
    x = np.asarray(centered)
    v = np.array(evars, copy=True)
    V = x.shape[1] // 2
    assert V >= 2, 'must have >=2 views to inflate variance'
    rounds = 0
    while True:
        rounds += 1
        if loading_matrix is None or mean is None:
            valid = np.ones(x.shape[0], dtype=bool)
            if likes is not None and likelihood_threshold is not None:
                valid = np.min(likes, axis=1) >= likelihood_threshold
            if v_quantile_threshold is not None:
                ev_max = v.max(axis=1)
                valid = valid & (ev_max < np.percentile(ev_max, v_quantile_threshold))
            fa = FactorAnalysis(n_components=n_latent).fit(x[valid])
            W, mu = fa.components_.T.astype(np.float64), fa.mean_.astype(np.float64)
        else:
            W, mu = np.asarray(loading_matrix, dtype=np.float64), np.asarray(mean, dtype=np.float64)
        v64, x64 = v.astype(np.float64), x.astype(np.float64)
        iv = 1.0 / (v64 + epsilon)
        Bm = np.linalg.inv(np.einsum('oi,to,oj->tij', W, iv, W))
        z = np.einsum('tij,oj,to->ti', Bm, W, iv * (x64 - mu))
        diff = x64 - (z @ W.T + mu)
        flag = np.zeros((x.shape[0], V), dtype=bool)
        for c in range(V):
            Wc = W[2 * c:2 * c + 2]
            Q = np.einsum('ai,tij,bj->tab', Wc, Bm, Wc)
            Q[:, 0, 0] += v64[:, 2 * c]
            Q[:, 1, 1] += v64[:, 2 * c + 1]
            d = diff[:, 2 * c:2 * c + 2]
            with np.errstate(all='ignore'):
                flag[:, c] = np.einsum('ta,tab,tb->t', d, np.linalg.inv(Q), d) > threshold
        full = np.repeat(flag, 2, axis=1)
        if V == 2:
            full |= full.any(axis=1, keepdims=True)
        v[full] *= scalar
        if not full.any():
            return v, rounds


# ----------------------------------------------------------------------------- IBL pupil model
PUPIL_POINTS = ['pupil_top_r', 'pupil_bottom_r', 'pupil_right_r', 'pupil_left_r']
PUPIL_C = np.array([[0, 1, 0], [-.5, 0, 1], [0, 1, 0], [.5, 0, 1], [.5, 1, 0], [0, 0, 1], [-.5, 1, 0], [0, 0, 1]],
                   dtype=np.float64)


def pupil_to_s(u, eps=1e-3):
    """_to_stable_s, eks/ibl_pupil_smoother.py:518-520."""
    return 1.0 / (1.0 + np.exp(-np.asarray(u, dtype=np.float64))) * (1.0 - 2 * eps) + eps


def pupil_nll_grad(ys, m0, S0, C, var3, Rdiag, u, dtype=np.float64):
    """-marginal_loglik of the pupil AR(1) model at u and its gradient (eks/ibl_pupil_smoother.py:551-563).
    ys (T,8); Rdiag (T,8) (already clipped); u (2,)."""
    ys, Rdiag = _c(ys, dtype), _c(Rdiag, dtype)
    m0, S0, C, var3, u = _c(m0, dtype), _c(S0, dtype), _c(C, dtype), _c(var3, dtype), _c(u, dtype)
    nll = np.empty(1, dtype=dtype)
    grad = np.empty(2, dtype=dtype)
    getattr(lib(), f'eks_oracle_pupil_nll_grad_{_sfx(dtype)}')(
        _p(m0), _p(S0), _p(C), _p(var3), _p(ys), _p(Rdiag), ys.shape[0], _p(u), _p(nll), _p(grad))
    return float(nll[0]), grad.astype(np.float64)


def pupil_optimize(ys, m0, S0, C, var3, Rdiag, lr=5e-3, tol=1e-6, safety_cap=5000, dtype=np.float32, trace_cap=0):
    """Adam on u from s0 = [0.99, 0.98] with the stop rule of eks/ibl_pupil_smoother.py:570-607."""
    ys, Rdiag = _c(ys, dtype), _c(Rdiag, dtype)
    m0, S0, C, var3 = _c(m0, dtype), _c(S0, dtype), _c(C, dtype), _c(var3, dtype)
    u = np.empty(2, dtype=dtype)
    loss = np.empty(1, dtype=dtype)
    iters = np.zeros(1, dtype=np.int32)
    trace = np.full((trace_cap, 3), np.nan, dtype=dtype) if trace_cap else None
    getattr(lib(), f'eks_oracle_pupil_optimize_{_sfx(dtype)}')(
        _p(m0), _p(S0), _p(C), _p(var3), _p(ys), _p(Rdiag), ys.shape[0], _real(dtype, lr), _real(dtype, tol),
        int(safety_cap), _p(u), _p(loss), _p(iters), _p(trace), int(trace_cap))
    return dict(u=u.astype(np.float64), s=pupil_to_s(u), loss=float(loss[0]), iters=int(iters[0]), trace=trace)


def _pupil_points(preds8):
    """(T,8) columns top/bottom/right/left x,y -> dict of (T,2) arrays."""
    return {n: preds8[:, 2 * i:2 * i + 2] for i, n in enumerate(['top', 'bottom', 'right', 'left'])}


def pupil_diameter(preds8):
    """get_pupil_diameter, eks/ibl_pupil_smoother.py:70-100."""
    p = _pupil_points(preds8)
    dist = lambda a, b: np.sqrt(((p[a] - p[b]) ** 2).sum(axis=1))
    ds = [dist('top', 'bottom'), dist('left', 'right')]
    ds += [dist(a, b) * 2 ** 0.5 for a, b in [('top', 'left'), ('top', 'right'), ('bottom', 'left'), ('bottom', 'right')]]
    with np.errstate(all='ignore'):
        return np.nanmedian(np.stack(ds), axis=0)


def pupil_location(preds8):
    """get_pupil_location, eks/ibl_pupil_smoother.py:33-67."""
    import warnings
    p = _pupil_points(preds8)
    out = np.zeros((preds8.shape[0], 2))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=RuntimeWarning)
        xa = np.nanmedian(np.stack([p['top'][:, 0], p['bottom'][:, 0]], 1), axis=1)
        xb = np.median(np.stack([p['right'][:, 0], p['left'][:, 0]], 1), axis=1)
        out[:, 0] = np.nanmedian(np.stack([xa, xb], 1), axis=1)
        ya = np.median(np.stack([p['top'][:, 1], p['bottom'][:, 1]], 1), axis=1)
        yb = np.nanmedian(np.stack([p['right'][:, 1], p['left'][:, 1]], 1), axis=1)
        out[:, 1] = np.nanmedian(np.stack([ya, yb], 1), axis=1)
    return out


def ibl_pupil(raw, smooth_params=None, s_frames=None, avg_mode='median', var_mode='confidence_weighted_var',
              dtype=np.float32, trace_cap=0):
    """ensemble_kalman_smoother_ibl_pupil restated (eks/ibl_pupil_smoother.py:197-359).
    raw (M, T, 4, 3) in the fixed point order top, bottom, right, left.
    Returns dict(out (4, 9, T) float64 in the reference's key-pair order top,right,bottom,left, s_finals, ms, Vs, info)."""
    raw = np.asarray(raw, dtype=np.float64)
    M, T, K, _ = raw.shape
    ens = ensemble(raw[:, None], avg_mode=avg_mode, var_mode=var_mode, dtype=np.float64)[0]   # (T,K,5)
    preds = ens[..., 0:2].reshape(T, -1).astype(np.float64)
    evars = ens[..., 2:4].reshape(T, -1).astype(np.float64)
    like = ens[..., 4].astype(np.float64)
    diam, loc = pupil_diameter(preds), pupil_location(preds)
    mx, my = loc[:, 0].mean(), loc[:, 1].mean()
    xt, yt = loc[:, 0] - mx, loc[:, 1] - my
    m0 = np.array([diam.mean(), 0.0, 0.0])
    S0 = np.diag([np.nanvar(diam), np.nanvar(xt), np.nanvar(yt)])
    var3 = np.array([diam.var(), xt.var(), yt.var()])
    y = preds.copy()
    y[:, 0::2] -= mx
    y[:, 1::2] -= my
    Rdiag = np.clip(evars, 1e-12, None)
    info = {}
    if smooth_params is not None and all(v is not None for v in smooth_params):
        s = np.clip(np.asarray(smooth_params, dtype=np.float32), 1e-3, 1 - 1e-3).astype(np.float64)
    else:
        opt = pupil_optimize(crop_frames(y, s_frames), m0, S0, PUPIL_C, var3, crop_frames(Rdiag, s_frames),
                             dtype=dtype, trace_cap=trace_cap)
        s = opt['s']
        info.update(opt)
    sd = np.array([s[0], s[1], s[1]])
    A, Q = np.diag(sd), np.diag(var3 * (1 - sd ** 2))
    ms, Vs = smooth(y[None], m0[None], S0[None], A[None], PUPIL_C[None], Q[None], Rdiag[None], 1.0, dtype=dtype)
    ms, Vs = ms[0].astype(np.float64), Vs[0].astype(np.float64)
    ym = ms @ PUPIL_C.T
    ym[:, 0::2] += mx
    ym[:, 1::2] += my
    yv = np.einsum('ij,tjk,lk->til', PUPIL_C, Vs, PUPIL_C)
    order = [0, 2, 1, 3]                               # key pairs: top, right, bottom, left
    ens_idx = [(0, 1), (4, 5), (2, 3), (6, 7)]
    out = np.empty((4, 9, T))
    for i, k in enumerate(order):
        out[i] = [ym[:, 2 * k], ym[:, 2 * k + 1], like[:, i], preds[:, ens_idx[i][0]], preds[:, ens_idx[i][1]],
                  evars[:, ens_idx[i][0]], evars[:, ens_idx[i][1]], yv[:, i, i], yv[:, i + 1, i + 1]]
    return dict(out=out, s_finals=[float(s[0]), float(s[1])], ms=ms, Vs=Vs, info=info)
