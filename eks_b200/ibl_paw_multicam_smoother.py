"""IBL two-camera paw driver around the multi-camera smoother (mirror of eks/ibl_paw_multicam_smoother.py).

Host I/O only: the right camera's markers are resampled onto the left camera's timestamps (linear interpolation,
frames outside the right camera's time range dropped), its x coordinates are flipped into the left camera's frame,
paws are swapped, a zero likelihood field is attached, and `ensemble_kalman_smoother_multicam` (the device pipeline)
does the rest.  The per-timestamp Python loop of the reference (:191-218) is one vectorised `np.interp` per column.
"""

from __future__ import annotations

import os
from collections.abc import Sequence
from typing import Literal

import numpy as np
import pandas as pd

from eks_b200.io import convert_lp_dlc, write_dlc_csv
from eks_b200.marker_array import MarkerArray, input_dfs_to_markerArray
from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam

__all__ = ['fit_eks_multicam_ibl_paw', 'add_camera_means', 'remove_camera_means', 'pca']


def _shift_camera_columns(ensemble_stacks: list, camera_means: Sequence, sign: float) -> list:
    """Column `camera_id` of every (T, n_cameras * 2) stack shifted by sign * camera_means[camera_id]; the stacks are
    modified in place and returned in a new list, as the reference's list.copy() does (:34-39, :57-62)."""
    out = list(ensemble_stacks)
    for stack in out:
        for camera_id, mean in enumerate(camera_means):
            stack[:, camera_id] = stack[:, camera_id] + sign * mean
    return out


def remove_camera_means(ensemble_stacks: list, camera_means: Sequence) -> list:
    """Subtract the per-camera mean coordinates (eks/ibl_paw_multicam_smoother.py:21-40)."""
    return _shift_camera_columns(ensemble_stacks, camera_means, -1.0)


def add_camera_means(ensemble_stacks: list, camera_means: Sequence) -> list:
    """Add the per-camera mean coordinates back (eks/ibl_paw_multicam_smoother.py:43-62)."""
    return _shift_camera_columns(ensemble_stacks, camera_means, 1.0)


def pca(S: np.ndarray, n_comps: int) -> tuple:
    """Fitted sklearn PCA and its explained-variance ratios (eks/ibl_paw_multicam_smoother.py:65-79)."""
    from sklearn.decomposition import PCA
    model = PCA(n_components=n_comps).fit(S)
    return model, model.explained_variance_ratio_


def fit_eks_multicam_ibl_paw(
    input_source: str,
    save_dir: str,
    smooth_param: float | list | None = None,
    s_frames: list | None = None,
    quantile_keep_pca: float = 50.0,
    avg_mode: Literal['mean', 'median'] = 'median',
    var_mode: Literal['var', 'confidence_weighted_var'] = 'confidence_weighted_var',
    img_width: int = 128,
    inflate_vars: bool = False,
    n_latent: int = 3,
) -> tuple:
    """Load the left / right seed CSVs and timestamp files of `input_source`, align the cameras in time, run the
    multi-camera EKS and save one CSV per camera.  Returns (camera_dfs, s_finals, input_dfs_list, bodypart_list)
    (eks/ibl_paw_multicam_smoother.py:82-256)."""
    bodypart_list = ['paw_l', 'paw_r']           # the IBL paw smoother works on this fixed set of points
    camera_names = ['left', 'right']
    swap = {'paw_l_x': 'paw_r_x', 'paw_l_y': 'paw_r_y', 'paw_l_likelihood': 'paw_r_likelihood',
            'paw_r_x': 'paw_l_x', 'paw_r_y': 'paw_l_y', 'paw_r_likelihood': 'paw_l_likelihood'}
    dfs = {'left': [], 'right': []}
    stamps = {'left': None, 'right': None}
    for filename in os.listdir(input_source):
        side = 'left' if 'left' in filename else 'right'
        path = os.path.join(input_source, filename)
        if 'timestamps' in filename:
            stamps[side] = np.load(path)
            continue
        df = convert_lp_dlc(pd.read_csv(path, header=[0, 1, 2], index_col=0), bodypart_list)
        if side == 'right':                      # the right camera sees the paws mirrored: swap them (:167-176)
            df = df.rename(columns=swap).loc[:, list(swap.keys())]
        dfs[side].append(df)
    if stamps['left'] is None or stamps['right'] is None:
        raise ValueError('Need timestamps for both cameras')
    if len(dfs['right']) != len(dfs['left']) or len(dfs['left']) == 0:
        raise ValueError('Need same number of left and right camera models and >=1 model for each.')

    t_left, t_right = np.asarray(stamps['left'], dtype=float), np.asarray(stamps['right'], dtype=float)
    keep = (t_left >= t_right[0]) & (t_left <= t_right[-1])     # left frames inside the right camera's time range
    keys = ['paw_l_x', 'paw_l_y', 'paw_r_x', 'paw_r_y']
    xy_cols = [0, 1, 3, 4]                                      # x, y of both paws in the flat layout
    input_dfs_list = [[], []]
    for df_l, df_r in zip(dfs['left'], dfs['right']):
        left = df_l.to_numpy(dtype=float)[keep][:, xy_cols]
        right_all = df_r.to_numpy(dtype=float)
        right = np.stack([np.interp(t_left[keep], t_right, right_all[:, j]) for j in xy_cols], axis=1)
        right[:, 0] = img_width - right[:, 0]                   # flip x into the left camera's frame
        right[:, 2] = img_width - right[:, 2]
        input_dfs_list[0].append(pd.DataFrame(left, columns=keys))
        input_dfs_list[1].append(pd.DataFrame(right, columns=keys))

    marker_array = input_dfs_to_markerArray(input_dfs_list, bodypart_list, camera_names, data_fields=['x', 'y'])
    lik_shape = list(marker_array.shape)
    lik_shape[-1] = 1                                           # zero likelihoods: use var_mode='var' (cli/cmd_ibl_paw.py:56)
    marker_array = MarkerArray.stack_fields(
        marker_array, MarkerArray(shape=tuple(lik_shape), data_fields=['likelihood'], dtype=marker_array.array.dtype))

    camera_dfs, s_finals, _ = ensemble_kalman_smoother_multicam(
        marker_array=marker_array, keypoint_names=bodypart_list, smooth_param=smooth_param,
        quantile_keep_pca=quantile_keep_pca, camera_names=camera_names, s_frames=s_frames, avg_mode=avg_mode,
        var_mode=var_mode, inflate_vars=inflate_vars, n_latent=n_latent, inflate_vars_kwargs={'likelihoods': None})
    os.makedirs(save_dir, exist_ok=True)
    for c, camera in enumerate(camera_names):
        write_dlc_csv(camera_dfs[c], os.path.join(save_dir, f'multicam_{camera}_results.csv'))
    return camera_dfs, s_finals, input_dfs_list, bodypart_list
