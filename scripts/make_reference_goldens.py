"""Pin the oracle to the REAL reference: run paninski-lab/eks itself (needs jax, dynamax, optax;
aniposelib for the calibrated case) on the raw arrays stored in tests/golden/*.npz and write its results to
tests/golden/reference/<name>.npz in the schema tests/test_reference_goldens.py consumes:

    out_ref  the nine output columns, same array layout as the oracle's golden (`out_f64` / `cam_out_f64`)
    s_ref    the smoothing parameters the reference selected

None of these packages can be installed in the build container or on the GPU boxes (no wheels, no network), so this
script exits with a clear message there; it is meant for any machine where `pip install ensemble-kalman-smoother`
works:   python scripts/make_reference_goldens.py [--eks-path /path/to/eks/checkout]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
COLS = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
        'y_posterior_var']


def _df_to_cols(df, kps):
    sc = df.columns[0][0]
    return np.stack([np.stack([df[(sc, kp, c)].to_numpy() for c in COLS], axis=-1) for kp in kps], axis=1)  # (T,K,9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--eks-path', default=None, help='checkout of paninski-lab/eks to import (default: installed eks)')
    a = ap.parse_args()
    if a.eks_path:
        sys.path.insert(0, a.eks_path)
    try:
        import jax  # noqa: F401
        from eks.marker_array import MarkerArray
        from eks.multicam_smoother import ensemble_kalman_smoother_multicam
        from eks.singlecam_smoother import ensemble_kalman_smoother_singlecam
    except Exception as e:   # pragma: no cover
        sys.exit(f'the reference is not importable here ({type(e).__name__}: {e}); run this where jax/dynamax/optax '
                 f'are installed')
    out_dir = os.path.join(GOLD, 'reference')
    os.makedirs(out_dir, exist_ok=True)
    fields = ['x', 'y', 'likelihood']

    def load(name):
        g = dict(np.load(os.path.join(GOLD, f'{name}.npz'), allow_pickle=True))
        if 'raw' not in g:
            g['raw'] = np.load(os.path.join(GOLD, f"{str(g['raw_from'])}.npz"))['raw']
        return g

    for name, kw in (('singlecam_ibl_pupil', {}), ('singlecam_ibl_pupil_fixed_s', dict(smooth_param=[0.5])),
                     ('singlecam_ibl_pupil_sframes', dict(s_frames=[(100, 700), (1200, None)])),
                     ('singlecam_mirror_mouse', {})):
        g = load(name)
        kps = [str(k) for k in g['keypoints']]
        df, s = ensemble_kalman_smoother_singlecam(MarkerArray(g['raw'].astype(np.float64), data_fields=fields), kps, **kw)
        np.savez_compressed(os.path.join(out_dir, f'{name}.npz'), out_ref=_df_to_cols(df, kps), s_ref=np.asarray(s))
        print(name, np.asarray(s))
    g = load('multicam_mirror_mouse_separate')
    kps, cams = [str(k) for k in g['keypoints']], [str(c) for c in g['cameras']]
    dfs, s, _ = ensemble_kalman_smoother_multicam(MarkerArray(g['raw'].astype(np.float64), data_fields=fields), kps, cams,
                                                  quantile_keep_pca=95.0)
    np.savez_compressed(os.path.join(out_dir, 'multicam_mirror_mouse_separate.npz'),
                        out_ref=np.stack([_df_to_cols(d, kps) for d in dfs]), s_ref=np.asarray(s))
    print('multicam_mirror_mouse_separate', np.asarray(s))
    try:
        from aniposelib.cameras import CameraGroup
        g = load('multicam_fly_nonlinear')
        kps, cams = [str(k) for k in g['keypoints']], [str(c) for c in g['cameras']]
        cg = CameraGroup.load(os.path.join(GOLD, 'fly_calibration.toml'))
        dfs, s, _ = ensemble_kalman_smoother_multicam(MarkerArray(g['raw'].astype(np.float64), data_fields=fields), kps,
                                                      cams, quantile_keep_pca=95.0, camgroup=cg)
        np.savez_compressed(os.path.join(out_dir, 'multicam_fly_nonlinear.npz'),
                            out_ref=np.stack([_df_to_cols(d, kps) for d in dfs]), s_ref=np.asarray(s))
        print('multicam_fly_nonlinear', np.asarray(s))
    except ImportError as e:   # pragma: no cover
        print('skipping the calibrated case:', e)


if __name__ == '__main__':
    main()
