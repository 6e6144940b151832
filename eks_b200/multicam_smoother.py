"""Multi-camera EKS entry points (drop-in for eks/multicam_smoother.py).

  ensemble_kalman_smoother_multicam   <- eks/multicam_smoother.py:279-551
  fit_eks_multicam / fit_eks_mirrored_multicam  <- :156-276 / :28-153
  initialize_kalman_filter_pca / _geometric     <- :554-597 / :600-650
  mA_compute_maha / inflate_variance            <- :653-764
  rodrigues / parse_dist / make_projection_from_camgroup / triangulate_3d_models /
  project_3d_covariance_to_2d                   <- :771-946

The hot path (ensemble -> run_kalman_smoother -> reprojection) runs on the CUDA library; PCA / factor
analysis / triangulation are one-off host pre-stages.  aniposelib is not available here, so a
minimal Anipose-TOML camera group (`CameraGroup`) with the `triangulate(fast=True)` recipe is included.
"""

from __future__ import annotations

import itertools
import logging
import os
import time
from typing import Literal

import numpy as np
import pandas as pd
import torch

from eks_b200 import _xfer, core, ops
from eks_b200._lib import require_cuda
from eks_b200.core import PinholeProjection, ensemble, run_kalman_smoother
from eks_b200.io import write_dlc_csv
from eks_b200.marker_array import MarkerArray, input_dfs_to_markerArray, mA_to_stacked_array, stacked_array_to_mA
from eks_b200.stats import compute_mahalanobis, compute_pca
from eks_b200.utils import center_predictions, make_dlc_pandas_index

logger = logging.getLogger(__name__)

LABELS = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
          'y_posterior_var']


# ----------------------------------------------------------------------------- calibration
def rodrigues(rvec) -> np.ndarray:
    """OpenCV-style Rodrigues rvec (3,) -> R (3,3) (eks/multicam_smoother.py:771-793)."""
    rvec = np.asarray(rvec, dtype=np.float64).ravel()
    theta = np.linalg.norm(rvec)
    if theta < 1e-12:
        rx, ry, rz = rvec
        return np.eye(3) + np.array([[0.0, -rz, ry], [rz, 0.0, -rx], [-ry, rx, 0.0]])
    rx, ry, rz = rvec / theta
    Km = np.array([[0.0, -rz, ry], [rz, 0.0, -rx], [-ry, rx, 0.0]])
    return np.eye(3) + np.sin(theta) * Km + (1.0 - np.cos(theta)) * (Km @ Km)


def parse_dist(dist_coeffs) -> dict:
    """OpenCV ordering [k1,k2,p1,p2,k3,k4,k5,k6,s1,s2,s3,s4,tx,ty], zero padded (:796-803)."""
    dc = np.zeros(14)
    d = np.asarray(dist_coeffs, dtype=np.float64).ravel()
    dc[:min(14, d.size)] = d[:14]
    names = ['k1', 'k2', 'p1', 'p2', 'k3', 'k4', 'k5', 'k6', 's1', 's2', 's3', 's4']
    return {n: dc[i] for i, n in enumerate(names)}


def pack_camera(rvec_or_R, tvec, Kmat, dist) -> np.ndarray:
    """One camera -> the 29-value layout of include/eks_b200.h (EKS_CAM_STRIDE)."""
    r = np.asarray(rvec_or_R, dtype=np.float64)
    Rm = r if r.shape == (3, 3) else rodrigues(r)
    Kmat = np.asarray(Kmat, dtype=np.float64)
    d = parse_dist(dist)
    out = np.zeros(29)
    out[0:9] = Rm.ravel()
    out[9:12] = np.asarray(tvec, dtype=np.float64).ravel()
    out[12:17] = [Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2], Kmat[0, 1]]
    out[17:29] = [d[n] for n in ['k1', 'k2', 'p1', 'p2', 'k3', 'k4', 'k5', 'k6', 's1', 's2', 's3', 's4']]
    return out


def make_jax_projection_fn(rvec, tvec, K, dist_coeffs) -> PinholeProjection:
    """One calibrated camera as an emission object (eks/multicam_smoother.py:806-859): callable (..., 3) -> (..., 2)
    like the reference's JAX closure (the name is kept for drop-in compatibility; there is no JAX here)."""
    return PinholeProjection(pack_camera(rvec, tvec, K, dist_coeffs)[None])


class Camera:
    """Minimal stand-in for aniposelib.cameras.Camera (only what the EKS path uses)."""

    def __init__(self, name, matrix, dist, rvec, tvec, size=None):
        self.name, self.size = name, size
        self.matrix = np.asarray(matrix, dtype=np.float64)
        self.dist = np.asarray(dist, dtype=np.float64).ravel()
        self.rvec = np.asarray(rvec, dtype=np.float64).ravel()
        self.tvec = np.asarray(tvec, dtype=np.float64).ravel()

    def get_rotation(self):
        return self.rvec

    def get_translation(self):
        return self.tvec

    def get_camera_matrix(self):
        return self.matrix

    def get_distortions(self):
        return self.dist

    def get_extrinsics_mat(self):
        E = np.eye(4)
        E[:3, :3] = rodrigues(self.rvec)
        E[:3, 3] = self.tvec
        return E

    def undistort_points(self, points):
        import cv2
        pts = np.asarray(points, dtype=np.float64).reshape(-1, 1, 2)
        out = cv2.undistortPoints(pts, self.matrix, self.dist)
        return out.reshape(np.asarray(points).shape)


class CameraGroup:
    """Anipose calibration TOML -> cameras; `triangulate(fast=True)` = undistort, pairwise DLT
    (cv2.triangulatePoints) and nan-median over camera pairs (aniposelib's fast path, restated)."""

    def __init__(self, cameras):
        self.cameras = list(cameras)

    @staticmethod
    def load(path: str) -> 'CameraGroup':
        import tomllib
        with open(path, 'rb') as f:
            cfg = tomllib.load(f)
        cams = []
        for key in sorted(k for k in cfg if k.startswith('cam_')):
            c = cfg[key]
            cams.append(Camera(c['name'], c['matrix'], c['distortions'], c['rotation'], c['translation'],
                               c.get('size')))
        return CameraGroup(cams)

    def triangulate(self, points, undistort=True, fast=True, disable_64bit=True, **_):
        import cv2
        points = np.asarray(points, dtype=np.float64)           # (C, N, 2)
        C, N, _ = points.shape
        pts = np.stack([cam.undistort_points(points[c]) if undistort else points[c]
                        for c, cam in enumerate(self.cameras)])
        Rt = [cam.get_extrinsics_mat()[:3] for cam in self.cameras]
        tris = []
        for j1, j2 in itertools.combinations(range(C), 2):
            t4 = cv2.triangulatePoints(Rt[j1], Rt[j2], pts[j1].T, pts[j2].T)
            tris.append((t4[:3] / t4[3]).T)
        with np.errstate(all='ignore'):
            return np.nanmedian(np.stack(tris), axis=0)         # (N, 3)


def make_projection_from_camgroup(camgroup) -> tuple[PinholeProjection, list[PinholeProjection]]:
    """(combined h: R^3 -> R^{2V}, per-camera heads) as parameter objects for the CUDA path (:862-885)."""
    import cv2
    packed = []
    for cam in camgroup.cameras:
        rot = np.array(cam.get_rotation())
        rvec = cv2.Rodrigues(rot)[0].ravel() if rot.shape == (3, 3) else rot.ravel()
        packed.append(pack_camera(rvec, np.array(cam.get_translation()).ravel(), np.array(cam.get_camera_matrix()),
                                  np.array(cam.get_distortions()).ravel()))
    packed = np.stack(packed)
    return PinholeProjection(packed), [PinholeProjection(packed[c:c + 1]) for c in range(packed.shape[0])]


def triangulate_3d_models(marker_array: MarkerArray, camgroup) -> np.ndarray:
    """(M,K,T,3) triangulated points per model and keypoint (:888-911)."""
    M, C, T, K, _ = marker_array.shape
    raw = marker_array.get_array()
    tri = np.zeros((M, K, T, 3), dtype=float)
    for m in range(M):
        for k in range(K):
            tri[m, k] = camgroup.triangulate(raw[m, :, :, k, :2], fast=True, disable_64bit=True)
    return tri


# ----------------------------------------------------------------------------- initialisation
def initialize_kalman_filter_pca(good_pcs_list, ensemble_pca, n_latent: int) -> tuple:
    K = len(good_pcs_list)
    m0s = np.zeros((K, n_latent))
    S0s = np.array([np.diag([np.var(good_pcs_list[k][:, i]) for i in range(n_latent)]) for k in range(K)])
    As = np.tile(np.eye(n_latent), (K, 1, 1))
    Cs = np.stack([pca.components_.T for pca in ensemble_pca])
    covs = []
    for k in range(K):
        d_t = np.diff(good_pcs_list[k], axis=0)
        cov = np.atleast_2d(np.cov(d_t.T))
        mx = np.max(np.abs(cov))
        covs.append(cov / mx if mx > 0 else cov)
    return m0s, S0s, As, np.stack(covs), Cs


def initialize_kalman_filter_geometric(ys: np.ndarray) -> tuple:
    K, T, D = ys.shape
    m0s = np.array([ys[k, :10].mean(axis=0) for k in range(K)])
    S0s = np.array([np.diag([np.nanvar(ys[k, :, d]) + 1e-4 for d in range(D)]) for k in range(K)])
    As = np.tile(np.eye(D), (K, 1, 1))
    Cs = np.tile(np.eye(D), (K, 1, 1))
    Qs = []
    for k in range(K):
        dx = np.diff(ys[k], axis=0)
        med = np.median(dx, axis=0)
        mad = np.median(np.abs(dx - med), axis=0) + 1e-12
        Qs.append(np.diag(np.maximum((1.4826 * mad) ** 2, 1e-8)))
    return m0s, S0s, As, np.array(Qs), Cs


# ----------------------------------------------------------------------------- variance inflation
def inflate_variance(v, maha_dict, threshold: float = 5.0, scalar: float = 10.0) -> tuple:
    assert len(maha_dict) >= 2, 'must have >=2 views to inflate variance'
    updated = v.copy()
    N, _ = v.shape
    C = len(maha_dict)
    mask = np.zeros((N, C), dtype=bool)
    for c, d in maha_dict.items():
        mask[:, c] = d[:, 0] > threshold
    full = np.repeat(mask, 2, axis=1)
    if C == 2:
        full |= full.any(axis=1, keepdims=True)
    updated[full] *= scalar
    return updated, full.any()


def mA_compute_maha(centered_emA_preds, emA_vars, emA_likes, n_latent, inflate_vars_kwargs={}, threshold=5.0,
                    scalar=10.0) -> MarkerArray:
    _, V, _, K, _ = centered_emA_preds.shape
    out = []
    for k in range(K):
        preds = mA_to_stacked_array(centered_emA_preds, k)
        tmp = mA_to_stacked_array(emA_vars, k)
        likes = mA_to_stacked_array(emA_likes, k)
        inflate_vars_kwargs.setdefault('likelihood_threshold', 0.9)
        inflate_vars_kwargs.setdefault('v_quantile_threshold', 50.0)
        inflated = True
        while inflated:
            kw = dict(inflate_vars_kwargs)
            if kw.get('likelihoods', None) is not None:
                kw['likelihoods'] = likes
            res = compute_mahalanobis(preds, tmp, n_latent=n_latent, **kw)
            tmp, inflated = inflate_variance(tmp, res['mahalanobis'], threshold, scalar)
        out.append(stacked_array_to_mA(tmp, V, data_fields=['var_x', 'var_y']))
    return MarkerArray.stack(out, 'keypoints')


def project_3d_covariance_to_2d(ms_k, Vs_k, h_cam: PinholeProjection, inflated_vars_k) -> tuple:
    """diag(J V J^T) + ensemble variance columns 0/1 for one camera head (:914-946), on the device."""
    dev = require_cuda()
    dtype = core.get_precision()
    T = ms_k.shape[0]
    ms = torch.as_tensor(np.asarray(ms_k)[None], device=dev).to(dtype).contiguous()
    Vs = torch.as_tensor(np.asarray(Vs_k)[None], device=dev).to(dtype).contiguous()
    var = torch.as_tensor(np.asarray(inflated_vars_k).T.copy(), device=dev).to(dtype).contiguous()  # [O][T]
    out = torch.empty((4, T), dtype=dtype, device=dev)
    cams = torch.as_tensor(h_cam.cams, device=dev).to(dtype).contiguous()
    O = var.shape[0]
    ops.reproject(ms, Vs, 1, out, 4 * T, 0, [0, T, 2 * T, 3 * T], cams=cams,
                  var=ops.PlaneView(var, O * T, [o * T for o in range(O)]), pinhole_var_quirk=True)
    o = out.double().cpu().numpy()
    return o[2], o[3]



def _multicam_linear_device(marker_array, keypoint_names, smooth_param, quantile_keep_pca, s_frames, avg_mode,
                            var_mode, n_latent, inflate_vars=False, inflate_vars_kwargs=None, cams=None,
                            n_cams_out: int | None = None, pca=None) -> tuple:
    """Linear PCA-latent model without variance inflation: every per-frame stage runs on the device
    (eks_b200.pipeline.multicam_smooth_sessions); the host only packs the DataFrames."""
    from eks_b200.pipeline import multicam_smooth_sessions
    from eks_b200.utils import normalize_spans
    dev = require_cuda()
    dtype = core.get_precision()
    M, V, T, K, _ = marker_array.shape
    t0 = time.perf_counter()
    arr = marker_array.array
    if list(marker_array.data_fields or ['x', 'y', 'likelihood']) != ['x', 'y', 'likelihood']:
        arr = marker_array.slice_fields('x', 'y', 'likelihood').array      # reorder / drop extra fields (as core.ensemble)
    raw = _xfer.to_device(arr, dev).to(dtype)
    sp = None
    if smooth_param is not None:   # scalar (also NumPy scalars / 0-d arrays) or one value per keypoint, as np.asarray would
        sp = [float(x) for x in np.broadcast_to(np.asarray(smooth_param, dtype=float), (K,))]
    res = multicam_smooth_sessions(raw[None], smooth_param=sp, spans=normalize_spans(T, s_frames),
                                   quantile_keep_pca=quantile_keep_pca, n_latent=n_latent, avg_mode=avg_mode,
                                   var_mode=var_mode, dtype=dtype, inflate_vars=inflate_vars,
                                   inflate_vars_kwargs=inflate_vars_kwargs, cams=cams, pca=pca)
    out_dev = torch.empty((V, T, K, 9), dtype=torch.float64, device=dev)
    out_dev.copy_(res.out[0].permute(1, 3, 0, 2))                                   # planes -> (V,T,K,9) float64
    out = _xfer.to_host(out_dev)
    del out_dev
    # 3-D DataFrame block assembled on the device: per keypoint [x, y, z, var_x, var_y, var_z] (:530-543)
    lat = torch.cat([res.ms[:, :, :3], torch.diagonal(res.Vs, dim1=2, dim2=3)[:, :, :3]], dim=2)   # (K,T,6)
    arr3d_dev = torch.empty((T, K, 6), dtype=torch.float64, device=dev)
    arr3d_dev.copy_(lat.permute(1, 0, 2))
    arr3d = _xfer.to_host(arr3d_dev.view(T, K * 6))
    del arr3d_dev, lat
    s_finals = res.s_finals[0].cpu().numpy()
    logger.debug(f'[profile] device pipeline (upload, smooth, download): {time.perf_counter() - t0:.3f}s')
    t0 = time.perf_counter()
    pdindex = make_dlc_pandas_index(keypoint_names, labels=LABELS)
    # one DataFrame per NAMED camera (the reference loops over camera_names, eks/multicam_smoother.py:452, :490)
    camera_dfs = [pd.DataFrame(out[c].reshape(T, K * 9), columns=pdindex, copy=False)
                  for c in range(min(V, n_cams_out or V))]                              # fresh arrays: no copy
    labels_3d = ['x', 'y', 'z', 'x_posterior_var', 'y_posterior_var', 'z_posterior_var']
    df_3d = pd.DataFrame(arr3d, columns=make_dlc_pandas_index(keypoint_names, labels=labels_3d), copy=False)
    logger.debug(f'[profile] packaging: {time.perf_counter() - t0:.3f}s')
    return camera_dfs, s_finals, df_3d


# ----------------------------------------------------------------------------- main entry point
def ensemble_kalman_smoother_multicam(
    marker_array: MarkerArray,
    keypoint_names: list,
    camera_names: list,
    smooth_param: float | list | None = None,
    quantile_keep_pca: float = 50.0,
    s_frames: list | None = None,
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
    inflate_vars: bool = False,
    inflate_vars_kwargs: dict = {},
    pca_object=None,
    n_latent: int = 3,
    camgroup=None,
) -> tuple:
    """Multi-view EKS: linear PCA latent (default) or calibrated pinhole EKF (camgroup given).
    Returns (camera_dfs, s_finals, df_3d)."""
    if camera_names is None or len(camera_names) == 0:
        raise ValueError('camera_names must be provided')
    dev = require_cuda()
    dtype = core.get_precision()
    M, V, T, K, _ = marker_array.shape
    t_total = time.perf_counter()

    if camgroup is None and os.environ.get('EKS_B200_HOST_PRESTAGE') != '1':
        if inflate_vars and inflate_vars_kwargs.get('mean', None) is not None:      # :355-357
            inflate_vars_kwargs['mean'] = np.zeros_like(inflate_vars_kwargs['mean'])
        pca = None
        if pca_object is not None:   # the caller's fitted PCA replaces the per-keypoint fit (eks/stats.py:52-56)
            pca = (np.asarray(pca_object.mean_, dtype=np.float64),
                   np.asarray(pca_object.components_, dtype=np.float64)[:n_latent])
        return _multicam_linear_device(marker_array, keypoint_names, smooth_param, quantile_keep_pca, s_frames,
                                       avg_mode, var_mode, n_latent, inflate_vars, inflate_vars_kwargs,
                                       n_cams_out=len(camera_names), pca=pca)
    if camgroup is not None and os.environ.get('EKS_B200_HOST_PRESTAGE') != '1':
        h_all, _ = make_projection_from_camgroup(camgroup)          # calibrated model, device-resident pipeline
        if inflate_vars and inflate_vars_kwargs.get('mean', None) is not None:      # :355-357
            inflate_vars_kwargs['mean'] = np.zeros_like(inflate_vars_kwargs['mean'])
        return _multicam_linear_device(marker_array, keypoint_names, smooth_param, quantile_keep_pca, s_frames,
                                       avg_mode, var_mode, n_latent if inflate_vars else 3, inflate_vars,
                                       inflate_vars_kwargs, cams=h_all.cams, n_cams_out=len(camera_names))

    t0 = time.perf_counter()
    ema = ensemble(marker_array, avg_mode=avg_mode, var_mode=var_mode)
    emA_unsm = ema.slice_fields('x', 'y')
    emA_vars = ema.slice_fields('var_x', 'var_y')
    emA_likes = ema.slice_fields('likelihood')
    valid_mask, emA_centered, emA_good_centered, emA_means = center_predictions(ema, quantile_keep_pca)
    logger.debug(f'[profile] ensemble + centering: {time.perf_counter() - t0:.3f}s')

    t0 = time.perf_counter()
    if inflate_vars:
        if inflate_vars_kwargs.get('mean', None) is not None:
            inflate_vars_kwargs['mean'] = np.zeros_like(inflate_vars_kwargs['mean'])
        emA_inflated = mA_compute_maha(emA_centered, emA_vars, emA_likes, n_latent,
                                       inflate_vars_kwargs=inflate_vars_kwargs)
    else:
        emA_inflated = emA_vars
    logger.debug(f'[profile] variance inflation: {time.perf_counter() - t0:.3f}s')

    nonlinear = camgroup is not None
    if nonlinear:
        h_fn, _ = make_projection_from_camgroup(camgroup)
        # triangulate_3d_models(...).mean(axis=0) on the device (eks_triangulate_mean): one thread per
        # (keypoint, frame) instead of M*K host calls into OpenCV
        raw_dev = torch.as_tensor(np.ascontiguousarray(marker_array.slice_fields('x', 'y', 'likelihood').array
                                                       if list(marker_array.data_fields) != ['x', 'y', 'likelihood']
                                                       else marker_array.array)).to(dev)
        ys_3d = ops.triangulate_mean(raw_dev, torch.as_tensor(np.asarray(h_fn.cams, dtype=np.float64))).cpu().numpy()
        del raw_dev
        m0s, S0s, As, Qs, Cs = initialize_kalman_filter_geometric(ys_3d)
        ys = np.stack([mA_to_stacked_array(emA_unsm, k) for k in range(K)])          # un-centred (:392-405)
        ens_vars = np.stack([mA_to_stacked_array(emA_inflated, k) for k in range(K)])
        D = 3
    else:
        pcas, good_pcs = compute_pca(valid_mask, emA_centered, emA_good_centered, n_components=n_latent,
                                     pca_object=pca_object)
        m0s, S0s, As, Qs, Cs = initialize_kalman_filter_pca(good_pcs, pcas, n_latent)
        ys = np.stack([mA_to_stacked_array(emA_centered, k) for k in range(K)])
        ens_vars = np.stack([mA_to_stacked_array(emA_inflated, k) for k in range(K)])
        h_fn = None
        D = n_latent

    t0 = time.perf_counter()
    s_finals, ms, Vs = run_kalman_smoother(ys=ys, m0s=m0s, S0s=S0s, As=As, Cs=Cs, Qs=Qs,
                                           ensemble_vars=np.swapaxes(ens_vars, 0, 1), s_frames=s_frames,
                                           smooth_param=smooth_param, h_fn=h_fn)
    logger.debug(f'[profile] run_kalman_smoother (total): {time.perf_counter() - t0:.3f}s')

    # reprojection epilogue on the device: planes [K][V][4][T] = x, y, posterior var x, posterior var y
    t0 = time.perf_counter()
    d_ms = torch.as_tensor(ms, device=dev).to(dtype).contiguous()
    d_Vs = torch.as_tensor(Vs, device=dev).to(dtype).contiguous()
    d_var = torch.as_tensor(np.ascontiguousarray(np.transpose(ens_vars, (0, 2, 1))), device=dev).to(dtype)
    vview = ops.PlaneView(d_var, 2 * V * T, [o * T for o in range(2 * V)])
    proj = torch.empty((K, V, 4, T), dtype=dtype, device=dev)
    if nonlinear:
        ops.reproject(d_ms, d_Vs, V, proj, V * 4 * T, 4 * T, [0, T, 2 * T, 3 * T],
                      cams=torch.as_tensor(h_fn.cams, device=dev).to(dtype).contiguous(), var=vview,
                      pinhole_var_quirk=True)
    else:
        means = np.stack([emA_means.array[0, :, 0, k, :].reshape(-1) for k in range(K)])   # (K, 2V)
        ops.reproject(d_ms, d_Vs, V, proj, V * 4 * T, 4 * T, [0, T, 2 * T, 3 * T],
                      C=torch.as_tensor(np.asarray(Cs), device=dev).to(dtype).contiguous(),
                      ymean=torch.as_tensor(means, device=dev).to(dtype).contiguous(), var=vview)
    proj = proj.double().cpu().numpy()
    out_vars = emA_vars if nonlinear else emA_inflated    # :474-477 vs :505-508
    pdindex = make_dlc_pandas_index(keypoint_names, labels=LABELS)
    camera_dfs = []
    for c in range(min(V, len(camera_names))):     # one DataFrame per named camera (:452, :490)
        cols = []
        for k in range(K):
            cols.extend([proj[k, c, 0], proj[k, c, 1], emA_likes.array[0, c, :, k, 0], emA_unsm.array[0, c, :, k, 0],
                         emA_unsm.array[0, c, :, k, 1], out_vars.array[0, c, :, k, 0], out_vars.array[0, c, :, k, 1],
                         proj[k, c, 2], proj[k, c, 3]])
        camera_dfs.append(pd.DataFrame(np.asarray(cols, dtype=np.float64).T, columns=pdindex, copy=False))
    labels_3d = ['x', 'y', 'z', 'x_posterior_var', 'y_posterior_var', 'z_posterior_var']
    arr_3d = []
    for k in range(K):
        arr_3d.extend([ms[k][:, 0], ms[k][:, 1], ms[k][:, 2], Vs[k][:, 0, 0], Vs[k][:, 1, 1], Vs[k][:, 2, 2]])
    df_3d = pd.DataFrame(np.asarray(arr_3d).T, columns=make_dlc_pandas_index(keypoint_names, labels=labels_3d), copy=False)
    logger.debug(f'[profile] reprojection + packaging: {time.perf_counter() - t0:.3f}s')
    logger.debug(f'[profile] ensemble_kalman_smoother_multicam total: {time.perf_counter() - t_total:.3f}s')
    return camera_dfs, s_finals, df_3d


def fit_eks_multicam(input_source, save_dir: str, bodypart_list: list | None = None,
                     smooth_param: float | list | None = None, s_frames: list | None = None,
                     camera_names: list | None = None, quantile_keep_pca: float = 50.0,
                     avg_mode: str = 'median', var_mode: str = 'confidence_weighted_var', inflate_vars: bool = False,
                     n_latent: int = 3, calibration: str | None = None, save_3d_outputs: bool = True) -> tuple:
    """Load per-camera seed CSVs, run the multi-camera EKS, save one CSV per camera (:156-276)."""
    from eks_b200.io import format_data
    if calibration is not None:
        camgroup = CameraGroup.load(calibration)
        if camera_names is not None:
            logger.warning('camera_names argument is ignored when calibration is provided; '
                           'camera names will be read from the calibration file')
        camera_names = [cam.name for cam in camgroup.cameras]
    else:
        camgroup = None
        if camera_names is None:
            raise ValueError('camera_names must be provided when no calibration file is given')
    input_dfs_list, keypoint_names = format_data(input_source, camera_names=camera_names)
    if bodypart_list is None:
        bodypart_list = keypoint_names
    marker_array = input_dfs_to_markerArray(input_dfs_list, bodypart_list, camera_names)
    camera_dfs, s_finals, df_3d = ensemble_kalman_smoother_multicam(
        marker_array=marker_array, keypoint_names=bodypart_list, smooth_param=smooth_param,
        quantile_keep_pca=quantile_keep_pca, camera_names=camera_names, s_frames=s_frames, avg_mode=avg_mode,
        var_mode=var_mode, inflate_vars=inflate_vars, n_latent=n_latent, camgroup=camgroup)
    os.makedirs(save_dir, exist_ok=True)
    for c, name in enumerate(camera_names):
        write_dlc_csv(camera_dfs[c], os.path.join(save_dir, f'multicam_{name}_results.csv'))
    if save_3d_outputs and calibration is not None:
        write_dlc_csv(df_3d, os.path.join(save_dir, 'multicam_3d_results.csv'))
    return camera_dfs, s_finals, input_dfs_list, bodypart_list, df_3d


def fit_eks_mirrored_multicam(input_source, save_file: str, bodypart_list: list | None = None,
                              smooth_param: float | list | None = None, s_frames: list | None = None,
                              camera_names: list = [], quantile_keep_pca: float = 50.0, avg_mode: str = 'median',
                              var_mode: str = 'confidence_weighted_var', inflate_vars: bool = False,
                              n_latent: int = 3) -> tuple:
    """Mirrored multi-camera data: every CSV holds all views, keypoints named '{bodypart}_{camera}'
    (eks/multicam_smoother.py:37-153).  Splits the columns per camera, runs the multi-camera smoother and writes ONE
    CSV whose keypoints carry the camera suffix again.  Returns (final_df, s_finals, input_dfs, bodypart_list)."""
    from eks_b200.io import format_data
    input_dfs_list, keypoint_names = format_data(input_source)
    if bodypart_list is None:
        bodypart_list = list(dict.fromkeys(name.split('_')[0] for name in keypoint_names))
    per_camera = []
    for cam in camera_names:
        dfs = []
        for df in input_dfs_list:
            rename = {c: c.replace(f'_{cam}', '') for c in df.columns if f'_{cam}_' in c}
            dfs.append(df[list(rename)].rename(columns=rename))
        per_camera.append(dfs)
    marker_array = input_dfs_to_markerArray(per_camera, bodypart_list, camera_names)
    camera_dfs, s_finals, _ = ensemble_kalman_smoother_multicam(
        marker_array=marker_array, keypoint_names=bodypart_list, smooth_param=smooth_param,
        quantile_keep_pca=quantile_keep_pca, camera_names=camera_names, s_frames=s_frames, avg_mode=avg_mode,
        var_mode=var_mode, inflate_vars=inflate_vars, n_latent=n_latent)
    parts = []
    for cam, cdf in zip(camera_names, camera_dfs):
        cdf = cdf.copy()
        cdf.columns = pd.MultiIndex.from_tuples([(sc, f'{kp}_{cam}', attr) for sc, kp, attr in cdf.columns],
                                                names=cdf.columns.names)
        parts.append(cdf)
    final_df = pd.concat(parts, axis=1)
    os.makedirs(os.path.dirname(save_file), exist_ok=True)
    write_dlc_csv(final_df, save_file)
    return final_df, s_finals, input_dfs_list, bodypart_list
