"""Pins to the REAL reference, consumed when they exist (VERDICT r1 "missing" 2 / item 6).

Two sources, both absent in the build container and on the GPU boxes (no network; jax / dynamax / optax cannot be
installed), so both tests SKIP LOUDLY today -- a skip is not a pass, the oracle stays "parity unpinned":

1. tests/golden/reference/*.npz written by scripts/make_reference_goldens.py on a machine where the reference runs:
   the ORACLE (fp32 mode = the reference's production precision) is compared with them at the reference's own golden
   tolerance (atol 1e-4, tests/conftest.py:95-101 of the reference).  CPU test.
2. eks_golden.zip (the reference's downloaded fixture, tests/conftest.py:12,44-49): the calls of the reference's
   integration tests (tests/integration/test_singlecam.py:4-10, test_multicam.py:4-14,31-41) are made through THIS
   package and every output CSV is compared exactly as the reference's `compare_to_golden` does.  GPU test; needs
   EKS_GOLDEN_ZIP (or tests/golden/eks_golden.zip) and the bundled data directory (EKS_DATA_DIR or
   /root/reference/data)."""
import glob
import os
import zipfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
ATOL = 1e-4


def test_oracle_matches_reference_generated_goldens():
    files = sorted(glob.glob(os.path.join(GOLD, 'reference', '*.npz')))
    if not files:
        pytest.skip('tests/golden/reference/*.npz absent (scripts/make_reference_goldens.py needs jax/dynamax/optax): '
                    'oracle NOT pinned to the real reference -- SKIPPED, not passed')
    from oracle import oracle
    for f in files:
        name = os.path.splitext(os.path.basename(f))[0]
        ref = np.load(f)
        g = dict(np.load(os.path.join(GOLD, f'{name}.npz'), allow_pickle=True))
        raw = g['raw'] if 'raw' in g else np.load(os.path.join(GOLD, f"{str(g['raw_from'])}.npz"))['raw']
        raw = raw.astype(np.float64)
        if name.startswith('singlecam'):
            kw = {'singlecam_ibl_pupil_fixed_s': dict(smooth_param=[0.5]),
                  'singlecam_ibl_pupil_sframes': dict(s_frames=[(100, 700), (1200, None)])}.get(name, {})
            r = oracle.singlecam(raw, dtype=np.float32, **kw)
            out = r['out']
        else:
            cal = os.path.join(GOLD, 'fly_calibration.toml') if 'fly' in name else None
            r = oracle.multicam(raw, dtype=np.float32, quantile_keep_pca=95.0, camgroup=cal)
            out = r['cam_out']
        np.testing.assert_allclose(r['s_finals'], ref['s_ref'], rtol=1e-3, err_msg=name)
        np.testing.assert_allclose(out, ref['out_ref'], rtol=0, atol=ATOL, err_msg=name)


def _compare_dir(out_dir, golden_dir, test_name):
    """The reference's compare_to_golden (tests/conftest.py:60-101), restated."""
    import pandas as pd
    csvs = sorted(glob.glob(os.path.join(out_dir, '*.csv')))
    assert csvs, f'no CSV written to {out_dir}'
    for c in csvs:
        gpath = os.path.join(golden_dir, test_name, os.path.basename(c))
        assert os.path.exists(gpath), f'golden file not found: {gpath}'
        actual, expected = pd.read_csv(c, index_col=0), pd.read_csv(gpath, index_col=0)
        assert actual.shape == expected.shape and list(actual.columns) == list(expected.columns), os.path.basename(c)
        np.testing.assert_allclose(actual.select_dtypes('number').values, expected.select_dtypes('number').values,
                                   rtol=0, atol=ATOL, err_msg=f'{test_name}/{os.path.basename(c)}')


@pytest.mark.gpu
def test_public_entry_points_match_eks_golden_zip(tmp_path):
    zpath = os.environ.get('EKS_GOLDEN_ZIP', os.path.join(GOLD, 'eks_golden.zip'))
    data = os.environ.get('EKS_DATA_DIR', '/root/reference/data')
    if not (os.path.exists(zpath) and os.path.isdir(data)):
        pytest.skip(f'eks_golden.zip ({zpath}) or the bundled data ({data}) absent: no pin to the real reference -- '
                    f'SKIPPED, not passed')
    from eks_b200 import fit_eks_multicam, fit_eks_singlecam
    gdir = tmp_path / 'golden'
    with zipfile.ZipFile(zpath) as zf:
        zf.extractall(gdir)
    roots = [str(gdir)] + [str(p) for p in gdir.iterdir() if p.is_dir()]
    gdir = next(r for r in roots if os.path.isdir(os.path.join(r, 'test_singlecam_defaults')))
    cases = [
        ('test_singlecam_defaults', lambda d: fit_eks_singlecam(
            input_source=f'{data}/ibl-pupil', save_file=f'{d}/eks_singlecam.csv')),
        ('test_singlecam_fixed_smooth_param', lambda d: fit_eks_singlecam(
            input_source=f'{data}/ibl-pupil', save_file=f'{d}/eks_singlecam.csv', smooth_param=[10.0])),
        ('test_multicam_defaults', lambda d: fit_eks_multicam(
            input_source=f'{data}/mirror-mouse-separate', save_dir=d, bodypart_list=['paw1LH', 'paw2LF'],
            camera_names=['top', 'bot'], quantile_keep_pca=95, inflate_vars=True)),
        ('test_multicam_fixed_smooth_param', lambda d: fit_eks_multicam(
            input_source=f'{data}/mirror-mouse-separate', save_dir=d, bodypart_list=['paw1LH', 'paw2LF'],
            camera_names=['top', 'bot'], smooth_param=[10.0], quantile_keep_pca=95, inflate_vars=True)),
        ('test_multicam_defaults_nonlinear', lambda d: fit_eks_multicam(
            input_source=f'{data}/fly', save_dir=d, bodypart_list=['L1A', 'L1B'],
            calibration=f'{data}/fly/calibration.toml', quantile_keep_pca=95, inflate_vars=True)),
        ('test_multicam_fixed_smooth_param_nonlinear', lambda d: fit_eks_multicam(
            input_source=f'{data}/fly', save_dir=d, bodypart_list=['L1A', 'L1B'],
            calibration=f'{data}/fly/calibration.toml', smooth_param=[10.0], quantile_keep_pca=95,
            inflate_vars=True)),
    ]
    for name, call in cases:
        d = tmp_path / name
        d.mkdir()
        call(str(d))
        _compare_dir(str(d), gdir, name)
