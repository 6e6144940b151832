"""MarkerArray: the 5-D (models, cameras, frames, keypoints, fields) container of the EKS API.

Host-side mirror of eks/marker_array.py:15-355 (same constructor, attributes and method names so code
written against the reference keeps working).  The container itself stays a NumPy array on the host;
the device layout used by the kernels is frame-major planes and is produced by
eks_b200.pipeline / eks_b200.core at the boundary.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = ['MarkerArray', 'input_dfs_to_markerArray', 'mA_to_stacked_array', 'stacked_array_to_mA']

_AXES = ('models', 'cameras', 'frames', 'keypoints', 'fields')


class MarkerArray:
    """5-D array container with named-axis slicing / stacking."""

    def __init__(self, array=None, shape=None, data_fields=None, marker_array=None, dtype=np.float32):
        if marker_array is not None:
            assert isinstance(marker_array, MarkerArray), 'marker_array must be a MarkerArray.'
            src = marker_array.array if array is None else array
            self.array = np.array(src, dtype=dtype)
            self.data_fields = marker_array.data_fields if data_fields is None else data_fields
        elif array is not None:
            if not isinstance(array, np.ndarray):
                try:  # torch tensors / anything array-like that is not a list
                    import torch
                    if isinstance(array, torch.Tensor):
                        array = array.detach().cpu().numpy()
                except ImportError:  # pragma: no cover
                    pass
            assert isinstance(array, np.ndarray), 'Input must be a NumPy array.'
            assert array.ndim == 5, 'Expected shape (n_models, n_cameras, n_frames, n_keypoints, n_fields).'
            self.array = array
            self.data_fields = data_fields
        elif shape is not None:
            assert len(shape) == 5, 'Shape must be (n_models, n_cameras, n_frames, n_keypoints, n_fields).'
            self.array = np.zeros(shape, dtype=dtype)
            self.data_fields = data_fields
        else:
            raise AssertionError('Provide either `array`, `shape`, or `marker_array`.')
        self.n_models, self.n_cameras, self.n_frames, self.n_keypoints, self.n_fields = self.array.shape
        self.axis_map = {name: i for i, name in enumerate(_AXES)}

    @property
    def shape(self):
        return self.array.shape

    def get_array(self, squeeze: bool = False) -> np.ndarray:
        return np.squeeze(self.array) if squeeze else self.array

    def slice(self, axis: str, indices) -> 'MarkerArray':
        assert axis in self.axis_map, f'Invalid slice axis: {axis}. Must be one of {list(self.axis_map)}.'
        if isinstance(indices, (int, np.integer)):
            indices = [int(indices)]
        return MarkerArray(np.take(self.array, indices, axis=self.axis_map[axis]), data_fields=self.data_fields)

    def slice_fields(self, *fields: str) -> 'MarkerArray':
        for f in fields:
            assert f in self.data_fields, f"Field '{f}' not found in data_fields: {self.data_fields}"
        idx = [self.data_fields.index(f) for f in fields]
        return MarkerArray(np.take(self.array, idx, axis=4), data_fields=list(fields))

    @staticmethod
    def stack(others, axis: str) -> 'MarkerArray':
        assert len(others) > 0, 'At least one MarkerArray must be provided for stacking.'
        ref = others[0]
        assert axis in ref.axis_map, f'Invalid stack axis: {axis}. Must be one of {list(ref.axis_map)}.'
        ax = ref.axis_map[axis]
        rest = lambda a: a.array.shape[:ax] + a.array.shape[ax + 1:]
        for o in others[1:]:
            assert isinstance(o, MarkerArray), "All elements in 'others' must be MarkerArray instances."
            assert rest(ref) == rest(o), f"Shape mismatch: Cannot stack along '{axis}' due to differing dimensions."
        return MarkerArray(np.concatenate([o.array for o in others], axis=ax), data_fields=ref.data_fields)

    def stack_fields(*marker_arrays: 'MarkerArray') -> 'MarkerArray':
        assert len(marker_arrays) > 0, 'At least one MarkerArray must be provided for stacking.'
        ref = marker_arrays[0]
        fields = []
        for o in marker_arrays:
            assert isinstance(o, MarkerArray), 'All inputs must be MarkerArray instances.'
            assert ref.array.shape[:4] == o.array.shape[:4], \
                "Shape mismatch: Cannot stack along 'fields' due to differing dimensions."
            assert o.data_fields is not None, 'All MarkerArrays must have data_fields defined.'
            fields.extend(o.data_fields)
        return MarkerArray(np.concatenate([o.array for o in marker_arrays], axis=4), data_fields=fields)

    def reorder_data_fields(self, new_order) -> 'MarkerArray':
        assert set(new_order) == set(self.data_fields), \
            f'Mismatch in data fields: Expected {self.data_fields}, but got {new_order}'
        idx = [self.data_fields.index(f) for f in new_order]
        return MarkerArray(marker_array=self, data_fields=list(new_order), array=np.take(self.array, idx, axis=4),
                           dtype=self.array.dtype)

    def __repr__(self) -> str:
        dims = ', '.join(f'{n}={s}' for n, s in zip(_AXES, self.array.shape))
        return f'MarkerArray({dims}, data_fields={self.data_fields}, type=NumPy)'


def input_dfs_to_markerArray(input_dfs_list, bodypart_list, camera_names, data_fields=('x', 'y', 'likelihood')):
    """list (per camera) of lists (per model) of flat '{kp}_{field}' DataFrames -> (M,V,T,K,F) float64.

    Mirrors eks/marker_array.py:269-299; the 4-deep Python loop is replaced by one column gather per
    (camera, model)."""
    data_fields = list(data_fields)
    K, V, M = len(bodypart_list), len(camera_names), len(input_dfs_list[0])
    T, F = input_dfs_list[0][0].shape[0], len(data_fields)
    cols = [f'{kp}_{f}' for kp in bodypart_list for f in data_fields]
    arr = np.zeros((M, V, T, K, F))
    for c in range(V):
        for m in range(M):
            arr[m, c] = input_dfs_list[c][m][cols].to_numpy(dtype=float).reshape(T, K, F)
    return MarkerArray(arr, data_fields=data_fields)


def mA_to_stacked_array(marker_array: MarkerArray, keypoint_idx: int) -> np.ndarray:
    """(1,V,T,K,F) -> (T, V*F) for one keypoint (eks/marker_array.py:302-324)."""
    _, V, T, K, F = marker_array.shape
    assert 0 <= keypoint_idx < K, f'keypoint_idx {keypoint_idx} is out of range (0-{K - 1})'
    sel = marker_array.array[0, :, :, keypoint_idx, :]  # (V,T,F)
    return np.transpose(sel, (1, 0, 2)).reshape(T, V * F)


def stacked_array_to_mA(reshaped_x: np.ndarray, n_cameras: int, data_fields) -> MarkerArray:
    """(T, V*F) -> (1,V,T,1,F) (eks/marker_array.py:327-355)."""
    T, total = reshaped_x.shape
    assert total % n_cameras == 0, 'Input shape mismatch: total fields must be divisible by n_cameras.'
    F = total // n_cameras
    x = reshaped_x.reshape(T, n_cameras, F).transpose(1, 0, 2)[None, :, :, None, :]
    return MarkerArray(x, data_fields=data_fields)
