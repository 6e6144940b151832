"""CSV ingest in the DLC / Lightning Pose 3-row-header format (mirror of eks/utils.py:35-232).

Host I/O only; `.slp` inputs need sleap_io, which is imported lazily."""

from __future__ import annotations

import logging
import os

import pandas as pd

logger = logging.getLogger(__name__)


def get_keypoint_names(df: pd.DataFrame) -> list:
    kps = df.columns[df.columns.get_level_values('coords') == 'x'].get_level_values('bodyparts')
    return kps.tolist()


def convert_lp_dlc(df_lp: pd.DataFrame, keypoint_names: list, model_name: str | None = None) -> pd.DataFrame:
    """3-level (scorer, bodypart, coord) columns -> flat '{keypoint}_{coord}' columns."""
    if model_name is None:
        model_name = str(df_lp.columns[0][0])
    cols = {}
    for kp in keypoint_names:
        for coord in ('x', 'y', 'likelihood'):
            key = (model_name, kp, coord)
            if any(isinstance(level, str) and level.startswith('Unnamed') for level in key):
                continue
            if key in df_lp.columns:
                cols[f'{kp}_{coord}'] = df_lp.loc[:, key]
    return pd.DataFrame(cols, index=df_lp.index)


def _read_one(file_path: str):
    if file_path.endswith('.slp'):
        raise NotImplementedError('.slp ingest needs sleap_io, which is not available in this environment')
    df = pd.read_csv(file_path, header=[0, 1, 2], index_col=0)
    kps = get_keypoint_names(df)
    return convert_lp_dlc(df, kps), kps


def format_data(input_source, camera_names: list | None = None) -> tuple[list, list]:
    """Directory / list of files / {camera: files} -> (list of flat DataFrames [per camera], keypoints)."""
    if isinstance(input_source, str) and os.path.isdir(input_source):
        file_paths = sorted(os.path.join(input_source, f) for f in os.listdir(input_source))
    elif isinstance(input_source, list):
        file_paths = sorted(input_source)
    elif isinstance(input_source, dict):
        file_paths = input_source
    else:
        raise ValueError('input_source must be a directory path, a list of file paths, or a map from camera '
                         'names to list of file paths')
    dfs, keypoint_names = [], None
    if camera_names is None:
        for fp in file_paths:
            if not (fp.endswith('.csv') or fp.endswith('.slp')):
                continue
            df, keypoint_names = _read_one(fp)
            dfs.append(df)
    else:
        for camera in camera_names:
            files = file_paths if isinstance(file_paths, list) else file_paths.get(camera, [])
            valid = [fp for fp in files if camera in os.path.basename(fp)
                     and (fp.endswith('.csv') or fp.endswith('.slp'))]
            if len(valid) == 0:
                raise FileNotFoundError(
                    f"no files matching camera '{camera}' found in {input_source}. "
                    f'ensure the camera name appears as a substring of each filename.')
            per_cam = []
            for fp in valid:
                df, keypoint_names = _read_one(fp)
                per_cam.append(df)
            dfs.append(per_cam)
        counts = [len(d) for d in dfs]
        if len(set(counts)) > 1:
            logger.warning('unequal number of seed files per camera (' + ', '.join(
                f'{c}: {n}' for c, n in zip(camera_names, counts)) + ')')
    if len(dfs) == 0:
        raise FileNotFoundError(f'no valid marker input files found in {input_source}')
    assert keypoint_names is not None
    return dfs, keypoint_names
