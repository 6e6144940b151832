"""Single-camera EKS entry points (drop-in for eks/singlecam_smoother.py).

  fit_eks_singlecam                    <- eks/singlecam_smoother.py:23-102
  ensemble_kalman_smoother_singlecam   <- eks/singlecam_smoother.py:105-243
  initialize_kalman_filter             <- eks/singlecam_smoother.py:246-284
"""

from __future__ import annotations

import logging
import os
from typing import Literal

import numpy as np
import pandas as pd
import torch

from eks_b200 import _xfer, core, ops
from eks_b200._lib import require_cuda
from eks_b200.marker_array import MarkerArray, input_dfs_to_markerArray
from eks_b200.pipeline import singlecam_smooth_sessions
from eks_b200.utils import make_dlc_pandas_index, normalize_spans

logger = logging.getLogger(__name__)

__all__ = ['fit_eks_singlecam', 'ensemble_kalman_smoother_singlecam', 'initialize_kalman_filter']


def fit_eks_singlecam(
    input_source: str | list,
    save_file: str,
    bodypart_list: list | None = None,
    smooth_param: float | list | None = None,
    s_frames: list | None = None,
    blocks: list = [],
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
) -> tuple:
    """Load seed CSVs, run the single-camera EKS, save the smoothed CSV.  Returns
    (df_smoothed, s_finals, input_dfs, bodypart_list)."""
    from eks_b200.io import format_data
    input_dfs_list, keypoint_names = format_data(input_source)
    if bodypart_list is None:
        bodypart_list = keypoint_names
        logger.info(f'input data loaded for keypoints:\n{bodypart_list}')
    marker_array = input_dfs_to_markerArray([input_dfs_list], bodypart_list, [''])
    df_smoothed, smooth_params_final = ensemble_kalman_smoother_singlecam(
        marker_array=marker_array, keypoint_names=bodypart_list, smooth_param=smooth_param, s_frames=s_frames,
        blocks=blocks, avg_mode=avg_mode, var_mode=var_mode,
    )
    os.makedirs(os.path.dirname(save_file), exist_ok=True)
    from eks_b200.io import write_dlc_csv
    write_dlc_csv(df_smoothed, save_file)
    logger.info('dataframes successfully converted to CSV')
    return df_smoothed, smooth_params_final, input_dfs_list, bodypart_list


def ensemble_kalman_smoother_singlecam(
    marker_array: MarkerArray,
    keypoint_names: list,
    smooth_param: float | list | None = None,
    s_frames: list | None = None,
    blocks: list = [],
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
) -> tuple:
    """Ensemble Kalman smoothing of single-camera data.  Returns (DataFrame (T x 9K), s_finals (K,))."""
    dev = require_cuda()
    dtype = core.get_precision()
    M, V, T, K, F = marker_array.shape
    assert V == 1, 'single-camera smoother expects n_cameras == 1'
    if T < 2:
        raise ValueError('Not enough frames to compute temporal differences.')
    arr = marker_array.array if list(marker_array.data_fields or ['x', 'y', 'likelihood']) == [
        'x', 'y', 'likelihood'] else marker_array.slice_fields('x', 'y', 'likelihood').array
    raw = _xfer.to_device(arr, dev)                  # chunked, double-buffered through pinned staging
    if raw.dtype not in (torch.float32, torch.float64) or (raw.dtype == torch.float32 and dtype == torch.float64):
        raw = raw.to(torch.float64)
    spans = normalize_spans(T, s_frames)
    res = singlecam_smooth_sessions(raw.reshape(1, M, 1, T, K, 3), smooth_param=smooth_param, spans=spans,
                                    blocks=blocks or None, avg_mode=avg_mode, var_mode=var_mode, dtype=dtype)
    final_dev = torch.empty((T, K, 9), dtype=torch.float64, device=dev)
    final_dev.copy_(res.out[0].permute(2, 0, 1))     # one transpose + cast kernel: planes -> (T, K, 9) float64
    final = _xfer.to_host(final_dev.view(T, K * 9))  # (T, K*9), keypoint-major
    del final_dev
    labels = ops.OUT_COLS
    # copy=False: `final` is a fresh array this function owns (pandas 3 would otherwise copy 1.4 GB per 10^6 frames: 0.8 s)
    markers_df = pd.DataFrame(final, columns=make_dlc_pandas_index(keypoint_names, labels=labels), copy=False)
    s_finals = res.s_finals[0].cpu().numpy().astype(float)
    return markers_df, s_finals


def initialize_kalman_filter(emA_centered_preds: MarkerArray) -> tuple:
    """(m0s, S0s, As, Qs, Cs) for the singlecam model: m0 = 0, S0 = diag(nanvar), A = Q = C = I."""
    K = emA_centered_preds.shape[3]
    cp = emA_centered_preds.slice_fields('x', 'y').get_array(squeeze=True)  # (T,K,2)
    m0s = np.zeros((K, 2))
    S0s = np.zeros((K, 2, 2))
    for k in range(K):
        S0s[k, 0, 0] = np.nanvar(cp[:, k, 0])
        S0s[k, 1, 1] = np.nanvar(cp[:, k, 1])
    eye = np.tile(np.eye(2), (K, 1, 1))
    return m0s, S0s, eye.copy(), eye.copy(), eye.copy()
