"""Host-side helpers of the path: frame cropping, R construction semantics, centring, DLC index.

Mirrors eks/utils.py (crop_frames :235-290, center_predictions :293-365, build_R_from_vars :368-377,
crop_R :380-398, make_dlc_pandas_index :15-32).  These are O(T) one-off host steps or pure argument
validation; the per-frame arithmetic lives in the CUDA library.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from eks_b200.io import convert_lp_dlc, format_data, get_keypoint_names  # noqa: F401  (eks/utils.py:35-232 live in io.py)
from eks_b200.marker_array import MarkerArray


def make_dlc_pandas_index(keypoint_names, labels=('x', 'y', 'likelihood')) -> pd.MultiIndex:
    return pd.MultiIndex.from_product([['ensemble-kalman_tracker'], list(keypoint_names), list(labels)],
                                      names=['scorer', 'bodyparts', 'coords'])


def normalize_spans(n: int, s_frames):
    """Validate an s_frames spec and return sorted [(start, end)] or None for 'all frames'.

    Same rules and error messages as crop_frames (eks/utils.py:245-284)."""
    if s_frames is None or (len(s_frames) == 1 and s_frames[0] == (None, None)) or len(s_frames) == 0:
        return None
    if not isinstance(s_frames, list):
        raise TypeError('s_frames must be a list of (start, end) tuples or None.')
    spans = []
    for i, frame in enumerate(s_frames):
        if not (isinstance(frame, tuple) and len(frame) == 2):
            raise ValueError(f's_frames[{i}] must be a (start, end) tuple, got {frame!r}')
        start, end = frame
        if start is not None and not isinstance(start, int):
            raise ValueError(f's_frames[{i}].start must be int or None, got {start!r}')
        if end is not None and not isinstance(end, int):
            raise ValueError(f's_frames[{i}].end must be int or None, got {end!r}')
        a = 0 if start is None else start
        b = n if end is None else end
        if a < 0 or b > n:
            raise ValueError(f'Range ({a}, {b}) out of bounds for length {n}.')
        if a >= b:
            raise ValueError(f'Invalid range ({a}, {b}).')
        spans.append((a, b))
    spans.sort(key=lambda s: s[0])
    for i in range(1, len(spans)):
        if spans[i][0] < spans[i - 1][1]:
            raise ValueError(f'Overlapping or out-of-order intervals: {spans[i - 1]} and {spans[i]}')
    return spans


def crop_frames(y, s_frames):
    """Crop the leading (time) axis of y to the union of [start, end) spans."""
    spans = normalize_spans(len(y), s_frames)
    if spans is None:
        return y
    if len(spans) == 1:
        return y[spans[0][0]:spans[0][1]]
    return np.concatenate([y[a:b] for a, b in spans], axis=0)


def build_R_from_vars(ev) -> np.ndarray:
    """(..., T, O) variances -> dense diagonal (..., T, O, O), clipped at 1e-12 (API parity only; the
    device path never materialises the dense form)."""
    ev = np.clip(np.asarray(ev), 1e-12, None)
    return ev[..., :, None] * np.eye(ev.shape[-1], dtype=ev.dtype)


def crop_R(R, s_frames):
    if not s_frames:
        return np.asarray(R)
    R = np.asarray(R)
    lead = R.shape[:-3]
    T, O, O2 = R.shape[-3:]
    assert O == O2, 'R_tv must be square in its last two dims'
    flat = R.reshape((-1, T, O, O))
    out = np.stack([crop_frames(b, s_frames) for b in flat], axis=0)
    return out.reshape((*lead, -1, O, O))


def center_predictions(ensemble_marker_array: MarkerArray, quantile_keep_pca: float):
    """Variance-quantile frame mask, per-keypoint mean over good frames, centred predictions.

    Same outputs as eks/utils.py:293-365: (valid_frames_mask (T,K), emA_centered_preds,
    emA_good_centered_preds, emA_means)."""
    n_models, V, T, K, _ = ensemble_marker_array.shape
    assert n_models == 1, 'MarkerArray should have n_models = 1 after ensembling.'
    preds = ensemble_marker_array.slice_fields('x', 'y').array       # (1,V,T,K,2)
    evars = ensemble_marker_array.slice_fields('var_x', 'var_y').array
    max_vars = np.max(evars, axis=(0, 1, 4))                           # (T,K)
    thresholds = np.percentile(max_vars, quantile_keep_pca, axis=0)
    mask = max_vars <= thresholds
    good = [np.where(mask[:, k])[0] for k in range(K)]
    min_frames = min(len(g) for g in good)
    centered = np.empty_like(preds)
    good_centered = np.empty((1, V, min_frames, K, 2), dtype=preds.dtype)
    means = np.empty((1, V, 1, K, 2), dtype=preds.dtype)
    for k in range(K):
        idx = good[k][:min_frames]
        gp = preds[:, :, idx, k, :]                                    # (1,V,n,2)
        mu = np.mean(gp, axis=2)                                       # (1,V,2)
        means[:, :, 0, k, :] = mu
        centered[:, :, :, k, :] = preds[:, :, :, k, :] - mu[:, :, None, :]
        good_centered[:, :, :, k, :] = gp - mu[:, :, None, :]
    fields = ['x', 'y']
    return (mask, MarkerArray(centered, data_fields=fields), MarkerArray(good_centered, data_fields=fields),
            MarkerArray(means, data_fields=fields))
