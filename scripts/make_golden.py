"""Generate tests/golden/*.npz from the reference's bundled data (run in the build container only:
/root/reference does not exist on the GPU box).

Inputs are the raw seed predictions packed exactly as input_dfs_to_markerArray packs them
(eks/marker_array.py:269-299); outputs are the ORACLE's results (oracle/oracle.py) in fp64 and fp32.
The reference itself cannot be executed here (jax/dynamax/optax absent), so these vectors pin the
product against the restated oracle -- "parity unpinned" w.r.t. the real JAX path.
"""
import glob
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = '/root/reference/data'
OUT = os.path.join(ROOT, 'tests', 'golden')


def load_csvs(files, keypoints=None):
    dfs = [pd.read_csv(f, header=[0, 1, 2], index_col=0) for f in files]
    sc = dfs[0].columns[0][0]
    kps = keypoints or dfs[0].columns[dfs[0].columns.get_level_values(2) == 'x'].get_level_values(1).tolist()
    T = len(dfs[0])
    raw = np.zeros((len(dfs), T, len(kps), 3))
    for m, df in enumerate(dfs):
        for k, kp in enumerate(kps):
            for d, f in enumerate(['x', 'y', 'likelihood']):
                raw[m, :, k, d] = df[(sc, kp, f)].to_numpy()
    return raw, kps


def singlecam_case(name, files, keypoints=None, raw_from=None, **kw):
    raw, kps = load_csvs(files, keypoints)
    raw = raw[:, None].astype(np.float32).astype(np.float64)  # (M,1,T,K,3); CSV values are float32
    res = {}
    for tag, dt in (('f64', np.float64), ('f32', np.float32)):
        r = oracle.singlecam(raw, dtype=dt, trace_cap=300, **kw)
        res[f'out_{tag}'] = r['out'].astype(np.float64 if tag == 'f64' else np.float32)
        res[f's_{tag}'] = r['s_finals']
        if 'iters' in r['info']:
            res[f'iters_{tag}'] = r['info']['iters']
            res[f'loss_{tag}'] = r['info']['loss']
            res[f'guess_{tag}'] = r['info']['guesses']
            res[f'Rconst_{tag}'] = r['info']['Rconst']
    if raw_from is None:
        res['raw'] = raw.astype(np.float32)  # the CSV values are float32-exact
    else:
        res['raw_from'] = np.array(raw_from)
    np.savez_compressed(os.path.join(OUT, f'{name}.npz'), keypoints=np.array(kps), **res)
    print(name, raw.shape, {k: (v.shape if hasattr(v, 'shape') else v) for k, v in res.items() if k.startswith('s_') or k.startswith('iters')},
          res.get('iters_f64'), res.get('iters_f32'))


def multicam_case(name, cam_files, camera_names, keypoints=None, calibration=None, **kw):
    """cam_files: {camera: [csv per seed]}.  Stores raw (M,V,T,K,3) float32 + oracle outputs."""
    raws = []
    kps = keypoints
    for cam in camera_names:
        r, kps = load_csvs(cam_files[cam], kps)
        raws.append(r)
    raw = np.stack(raws, axis=1).astype(np.float32).astype(np.float64)   # (M,V,T,K,3)
    camgroup = calibration   # path of the Anipose TOML: the oracle has its own loader / triangulation
    res = {'raw': raw.astype(np.float32), 'keypoints': np.array(kps), 'cameras': np.array(camera_names)}
    for tag, dt in (('f64', np.float64), ('f32', np.float32)):
        r = oracle.multicam(raw, camgroup=camgroup, dtype=dt, **kw)
        res[f'cam_out_{tag}'] = r['cam_out'].astype(np.float64 if tag == 'f64' else np.float32)
        res[f'out3d_{tag}'] = r['out3d'].astype(np.float64 if tag == 'f64' else np.float32)
        res[f's_{tag}'] = r['s_finals']
        res[f'iters_{tag}'] = r['info']['iters']
    np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **res)
    print(name, raw.shape, res['s_f64'], res['iters_f64'], res['iters_f32'])


def pupil_case(name, files, **kw):
    """IBL pupil smoother (eks/ibl_pupil_smoother.py) on the fixed point order; raw is the singlecam_ibl_pupil
    golden's raw with keypoints re-ordered, so only the oracle outputs are stored."""
    from oracle.oracle import PUPIL_POINTS
    raw, kps = load_csvs(files, PUPIL_POINTS)
    raw = raw.astype(np.float32).astype(np.float64)     # (M,T,4,3)
    res = {'raw_from': np.array('singlecam_ibl_pupil'), 'keypoints': np.array(kps)}
    for tag, dt in (('f64', np.float64), ('f32', np.float32)):
        r = oracle.ibl_pupil(raw, dtype=dt, trace_cap=64, **kw)
        res[f'out_{tag}'] = r['out'].astype(np.float64 if tag == 'f64' else np.float32)
        res[f's_{tag}'] = np.asarray(r['s_finals'])
        if 'iters' in r['info']:
            res[f'iters_{tag}'] = np.asarray(r['info']['iters'])
            res[f'loss_{tag}'] = np.asarray(r['info']['loss'])
            res[f'trace_{tag}'] = r['info']['trace']
    np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **res)
    print(name, raw.shape, res['s_f64'], res['s_f32'], res.get('iters_f64'), res.get('iters_f32'))


def pupil_goldens():
    files = sorted(glob.glob(f'{REF}/ibl-pupil/*.csv'))
    pupil_case('ibl_pupil', files)
    pupil_case('ibl_pupil_sframes', files, s_frames=[(100, 700), (1200, None)])
    pupil_case('ibl_pupil_fixed_s', files, smooth_params=[0.9, 0.95])


def multicam_goldens():
    # BASELINE config 3 family: multicam linear on data/mirror-mouse-separate (2 cams x 10? seeds, 501 frames)
    d = f'{REF}/mirror-mouse-separate'
    cams = ['top', 'bot']
    files = {c: sorted(glob.glob(f'{d}/*.{c}.csv')) for c in cams}
    multicam_case('multicam_mirror_mouse_separate', files, cams, quantile_keep_pca=95.0)
    # BASELINE config 4 family: calibrated nonlinear EKF on data/fly, bodyparts L1A, L1B
    # (reference tests/integration/test_multicam.py:31-41)
    d = f'{REF}/fly'
    cams = ['Cam-A', 'Cam-B', 'Cam-C']
    files = {c: sorted(glob.glob(f'{d}/*{c}*.csv')) for c in cams}
    multicam_case('multicam_fly_nonlinear', files, cams, keypoints=['L1A', 'L1B'],
                  calibration=f'{d}/calibration.toml', quantile_keep_pca=95.0)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'pupil':
        pupil_goldens()
        sys.exit(0)
    # BASELINE config 1: `eks singlecam` on data/ibl-pupil (tests/integration/test_singlecam.py:4-10)
    singlecam_case('singlecam_ibl_pupil', sorted(glob.glob(f'{REF}/ibl-pupil/*.csv')))
    # fixed smoothing parameter and s_frames variants
    singlecam_case('singlecam_ibl_pupil_fixed_s', sorted(glob.glob(f'{REF}/ibl-pupil/*.csv')), smooth_param=[0.5],
                   raw_from='singlecam_ibl_pupil')
    singlecam_case('singlecam_ibl_pupil_sframes', sorted(glob.glob(f'{REF}/ibl-pupil/*.csv')),
                   s_frames=[(100, 700), (1200, None)], raw_from='singlecam_ibl_pupil')
    # mirror-mouse (singlecam on 5 seeds, 501 frames, many keypoints): first 6 keypoints
    singlecam_case('singlecam_mirror_mouse', sorted(glob.glob(f'{REF}/mirror-mouse/*.csv')),
                   keypoints=None)
    multicam_goldens()
    pupil_goldens()
