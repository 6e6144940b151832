"""IBL pupil smoother (drop-in for eks/ibl_pupil_smoother.py).

  fit_eks_pupil                        <- eks/ibl_pupil_smoother.py:120-194
  ensemble_kalman_smoother_ibl_pupil   <- :197-359
  run_pupil_kalman_smoother            <- :363-448   (pupil_optimize_smooth :452-607 runs in eks_pupil_optimize)
  get_pupil_location / get_pupil_diameter / add_mean_to_array  <- :33-117

Geometry helpers are O(T) host pre-stages in NumPy; the ensemble statistics, the Adam optimisation of the two
AR(1) parameters over the time-varying-R filter NLL and the final filter + RTS smoother run on the CUDA library.
"""

from __future__ import annotations

import logging
import os
import warnings
from typing import Literal

import numpy as np
import pandas as pd
import torch

from eks_b200 import core, ops
from eks_b200._lib import require_cuda
from eks_b200.core import ensemble
from eks_b200.marker_array import MarkerArray, input_dfs_to_markerArray
from eks_b200.ops import Model, PlaneView
from eks_b200.utils import make_dlc_pandas_index, normalize_spans

logger = logging.getLogger(__name__)

__all__ = ['fit_eks_pupil', 'ensemble_kalman_smoother_ibl_pupil', 'get_pupil_location', 'get_pupil_diameter']

# observation matrix of the pupil model: rows = (top, bottom, right, left) x (x, y); columns = (diameter, com_x, com_y)
PUPIL_C = np.asarray([[0, 1, 0], [-.5, 0, 1], [0, 1, 0], [.5, 0, 1], [.5, 1, 0], [0, 0, 1], [-.5, 1, 0], [0, 0, 1]],
                     dtype=np.float64)


def get_pupil_location(dlc: dict) -> np.ndarray:
    """Pupil centre of mass per frame from the four pupil points (robust to one missing point per axis)."""
    t = np.vstack((dlc['pupil_top_r_x'], dlc['pupil_top_r_y'])).T
    b = np.vstack((dlc['pupil_bottom_r_x'], dlc['pupil_bottom_r_y'])).T
    le = np.vstack((dlc['pupil_left_r_x'], dlc['pupil_left_r_y'])).T
    r = np.vstack((dlc['pupil_right_r_x'], dlc['pupil_right_r_y'])).T
    center = np.zeros(t.shape)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=RuntimeWarning)
        x1 = np.nanmedian(np.stack([t[:, 0], b[:, 0]], axis=1), axis=1)      # either may be NaN
        x2 = np.median(np.stack([r[:, 0], le[:, 0]], axis=1), axis=1)        # both must be present
        center[:, 0] = np.nanmedian(np.stack([x1, x2], axis=1), axis=1)
        y1 = np.median(np.stack([t[:, 1], b[:, 1]], axis=1), axis=1)
        y2 = np.nanmedian(np.stack([r[:, 1], le[:, 1]], axis=1), axis=1)
        center[:, 1] = np.nanmedian(np.stack([y1, y2], axis=1), axis=1)
    return center


def get_pupil_diameter(dlc: dict) -> np.ndarray:
    """Median over six diameter estimates (two direct, four via the circle assumption)."""
    top, bottom, left, right = [np.vstack((dlc[f'pupil_{p}_r_x'], dlc[f'pupil_{p}_r_y']))
                                for p in ['top', 'bottom', 'left', 'right']]
    ds = [np.linalg.norm(top - bottom, axis=0), np.linalg.norm(left - right, axis=0)]
    for a, b in [(top, left), (top, right), (bottom, left), (bottom, right)]:
        ds.append(np.linalg.norm(a - b, axis=0) * 2 ** 0.5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=RuntimeWarning)
        return np.nanmedian(ds, axis=0)


def add_mean_to_array(pred_arr: np.ndarray, keys: list, mean_x, mean_y) -> dict:
    out = {}
    for i, key in enumerate(keys):
        out[key] = pred_arr[:, i] + (mean_x if 'x' in key else mean_y)
    return out


def pupil_model_arrays(ensemble_preds: np.ndarray):
    """(T, 8) ensemble medians -> (y_obs centred (T,8), m0 (3,), S0 (3,3), var3 (3,), mean_x, mean_y)."""
    keys = [f'{kp}_{c}' for kp in ['pupil_top_r', 'pupil_bottom_r', 'pupil_right_r', 'pupil_left_r'] for c in 'xy']
    d = {k: ensemble_preds[:, i] for i, k in enumerate(keys)}
    diam = get_pupil_diameter(d)
    loc = get_pupil_location(d)
    mean_x, mean_y = np.mean(loc[:, 0]), np.mean(loc[:, 1])
    x_t, y_t = loc[:, 0] - mean_x, loc[:, 1] - mean_y
    m0 = np.asarray([np.mean(diam), 0.0, 0.0])
    S0 = np.diag([np.nanvar(diam), np.nanvar(x_t), np.nanvar(y_t)])
    var3 = np.asarray([np.var(diam), np.var(x_t), np.var(y_t)])
    y_obs = ensemble_preds.copy()
    y_obs[:, 0::2] -= mean_x
    y_obs[:, 1::2] -= mean_y
    return y_obs, m0, S0, var3, mean_x, mean_y


def run_pupil_kalman_smoother(ys, m0, S0, C, ensemble_vars, diameters_var, x_var, y_var, s_frames=None,
                              smooth_params=None, lr: float = 5e-3, tol: float = 1e-6, safety_cap: int = 5000):
    """Optimise [s_diam, s_com] on the EKF-filter NLL (time-varying R_t), then run the EKF smoother.
    Returns ([s_diam, s_com], ms (T,3), Vs (T,3,3))."""
    dev = require_cuda()
    dtype = core.get_precision()
    ys = np.asarray(ys, dtype=np.float64)
    T = ys.shape[0]
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev).to(dtype).contiguous()
    y_planes = f(ys.T[None])                                    # [1][8][T]
    v_planes = f(np.asarray(ensemble_vars, dtype=np.float64).T[None])
    yv = PlaneView(y_planes, 8 * T, [o * T for o in range(8)])
    vv = PlaneView(v_planes, 8 * T, [o * T for o in range(8)])
    m0_d, S0_d, C_d = f(np.asarray(m0)[None]), f(np.asarray(S0)[None]), f(np.asarray(C)[None])
    var3 = np.asarray([diameters_var, x_var, y_var], dtype=np.float64)
    if smooth_params is not None and all(v is not None for v in smooth_params):
        s = np.clip(np.asarray(smooth_params, dtype=np.float32), 1e-3, 1 - 1e-3).astype(np.float64)
    else:
        spans = normalize_spans(T, s_frames)
        opt = ops.pupil_optimize(m0_d, S0_d, C_d, f(var3[None]), yv, vv, T, spans=spans, lr=lr, tol=tol,
                                 safety_cap=safety_cap)
        s = opt['s'][0].double().cpu().numpy()
        logger.debug(f'[pupil] iters={int(opt["iters"][0])}  s_diam={s[0]:.6f}  s_com={s[1]:.6f}  '
                     f'NLL={float(opt["loss"][0]):.6f}')
    s_d, s_c = float(s[0]), float(s[1])
    A = np.diag([s_d, s_c, s_c])
    Q = np.diag([var3[0] * (1 - s_d ** 2), var3[1] * (1 - s_c ** 2), var3[2] * (1 - s_c ** 2)])
    model = Model(m0_d, S0_d, f(A[None]), f(Q[None]), C_d)
    ms, Vs = ops.filter_smooth(model, yv, vv, T, torch.ones(1, dtype=dtype, device=dev))
    return [s_d, s_c], ms[0].cpu().numpy(), Vs[0].cpu().numpy()


def ensemble_kalman_smoother_ibl_pupil(
    marker_array: MarkerArray,
    keypoint_names: list,
    smooth_params: list | None = None,
    s_frames: list | None = None,
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
) -> tuple:
    """Ensemble Kalman smoothing of IBL pupil data.  Returns (DataFrame (T x 36), [s_diam, s_com])."""
    M, V, T, K, _ = marker_array.shape
    keys = [f'{kp}_{coord}' for kp in keypoint_names for coord in ['x', 'y']]
    ema = ensemble(marker_array, avg_mode=avg_mode, var_mode=var_mode)
    ensemble_preds = ema.slice_fields('x', 'y').get_array(squeeze=True).reshape(T, -1)
    ensemble_vars = ema.slice_fields('var_x', 'var_y').get_array(squeeze=True).reshape(T, -1)
    ensemble_likes = ema.slice_fields('likelihood').get_array(squeeze=True)
    y_obs, m0, S0, var3, mean_x, mean_y = pupil_model_arrays(ensemble_preds.astype(np.float64))
    s_finals, ms, Vs = run_pupil_kalman_smoother(y_obs, m0, S0, PUPIL_C, ensemble_vars, var3[0], var3[1], var3[2],
                                                 s_frames=s_frames, smooth_params=smooth_params)
    y_m = (PUPIL_C @ ms.astype(np.float64).T).T
    y_v = np.einsum('ij,tjk,lk->til', PUPIL_C, Vs.astype(np.float64), PUPIL_C)
    processed = add_mean_to_array(y_m, keys, mean_x, mean_y)
    key_pairs = [['pupil_top_r_x', 'pupil_top_r_y'], ['pupil_right_r_x', 'pupil_right_r_y'],
                 ['pupil_bottom_r_x', 'pupil_bottom_r_y'], ['pupil_left_r_x', 'pupil_left_r_y']]
    ens_idx = [(0, 1), (4, 5), (2, 3), (6, 7)]
    labels = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
              'y_posterior_var']
    data = []
    for i, (kx, ky) in enumerate(key_pairs):   # the reference's (quirky) index pairing is kept: :324-351
        data.extend([processed[kx], processed[ky], ensemble_likes[:, i], ensemble_preds[:, ens_idx[i][0]],
                     ensemble_preds[:, ens_idx[i][1]], ensemble_vars[:, ens_idx[i][0]], ensemble_vars[:, ens_idx[i][1]],
                     y_v[:, i, i], y_v[:, i + 1, i + 1]])
    df = pd.DataFrame(np.asarray(data, dtype=np.float64).T, columns=make_dlc_pandas_index(keypoint_names, labels=labels),
                      copy=False)
    return df, s_finals


def fit_eks_pupil(input_source, save_file: str, smooth_params: list | None = None, s_frames: list | None = None,
                  avg_mode: str = 'median', var_mode: str = 'confidence_weighted_var') -> tuple:
    """Load seed CSVs, run the pupil EKS, save the smoothed CSV (order of the four points is fixed)."""
    from eks_b200.io import format_data
    bodypart_list = ['pupil_top_r', 'pupil_bottom_r', 'pupil_right_r', 'pupil_left_r']
    input_dfs_list, _ = format_data(input_source)
    marker_array = input_dfs_to_markerArray([input_dfs_list], bodypart_list, [''])
    df, s = ensemble_kalman_smoother_ibl_pupil(marker_array, bodypart_list, smooth_params=smooth_params,
                                               s_frames=s_frames, avg_mode=avg_mode, var_mode=var_mode)
    os.makedirs(os.path.dirname(save_file), exist_ok=True)
    from eks_b200.io import write_dlc_csv
    write_dlc_csv(df, save_file)
    return df, s, input_dfs_list, bodypart_list
