"""CPU tests: host logic of the drop-in layer, the C ABI surface, and the kernel arithmetic compiled for
the host (tests/hostcheck) against the oracle.  No GPU needed; no compute call goes to the CUDA library."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, synth_singlecam
from oracle import oracle

HC_DIR = os.path.join(ROOT, 'tests', 'hostcheck')


@pytest.fixture(scope='module')
def hc():
    subprocess.run(['make', '-C', HC_DIR, 'libhostcheck.so'], check=True, capture_output=True)
    return ctypes.CDLL(os.path.join(HC_DIR, 'libhostcheck.so'))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def hc_nll(hc, D, O, y_TO, m0, S0, A, Q, C, cams, Rc, s, dt):
    sf = 'f32' if dt == np.float32 else 'f64'
    cr = ctypes.c_float if dt == np.float32 else ctypes.c_double
    T = y_TO.shape[0]
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=dt) for x in (m0, S0, A, Q, C, cams)]
    yp = np.ascontiguousarray(y_TO.T, dtype=dt)
    nll, dn = np.zeros(1, dt), np.zeros(1, dt)
    ncam = 0 if cams is None else np.asarray(cams).reshape(-1, 29).shape[0]
    getattr(hc, 'hostcheck_nll_grad_' + sf)(D, O, T, *[_p(a) for a in arrs[:5]], ncam, _p(arrs[5]), _p(yp),
                                           _p(np.ascontiguousarray(Rc, dtype=dt)), cr(s), _p(nll), _p(dn))
    return nll[0], dn[0]


def hc_smooth(hc, D, O, y_TO, var_TO, m0, S0, A, Q, C, cams, s, dt):
    sf = 'f32' if dt == np.float32 else 'f64'
    cr = ctypes.c_float if dt == np.float32 else ctypes.c_double
    T = y_TO.shape[0]
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=dt) for x in (m0, S0, A, Q, C, cams)]
    yp, vp = np.ascontiguousarray(y_TO.T, dtype=dt), np.ascontiguousarray(var_TO.T, dtype=dt)
    mf, Pf = np.zeros((T, D), dt), np.zeros((T, D, D), dt)
    ms, Vs = np.zeros((T, D), dt), np.zeros((T, D, D), dt)
    ncam = 0 if cams is None else np.asarray(cams).reshape(-1, 29).shape[0]
    getattr(hc, 'hostcheck_smooth_' + sf)(D, O, T, *[_p(a) for a in arrs[:5]], ncam, _p(arrs[5]), _p(yp), _p(vp),
                                         cr(s), _p(mf), _p(Pf), _p(ms), _p(Vs))
    return ms, Vs


# ------------------------------------------------------------------ kernel arithmetic on the host
@pytest.mark.parametrize('D,O', [(2, 2), (3, 4), (4, 8), (5, 6)])
def test_kernel_math_linear_vs_oracle(hc, D, O):
    from test_oracle import random_linear
    y, m0, S0, A, C, Q, Rt = random_linear(D, O, 300, seed=D + O)
    for dt, rtol in ((np.float64, 1e-7), (np.float32, 2e-3)):
        n_o, g_o = oracle.nll_grad(y[None], m0[None], S0[None], A[None], C[None], Q[None], Rt[0][None], 0.4,
                                   dtype=dt)
        n_h, g_h = hc_nll(hc, D, O, y, m0, S0, A, Q, C, None, Rt[0], 0.4, dt)
        np.testing.assert_allclose(n_h, n_o[0], rtol=rtol)
        np.testing.assert_allclose(g_h, g_o[0], rtol=rtol * 10)
    ms_o, Vs_o = oracle.smooth(y[None], m0[None], S0[None], A[None], C[None], Q[None], Rt[None], 0.4,
                               dtype=np.float64)
    ms_h, Vs_h = hc_smooth(hc, D, O, y, Rt, m0, S0, A, Q, C, None, 0.4, np.float64)
    np.testing.assert_allclose(ms_h, ms_o[0], rtol=1e-6, atol=1e-7 * np.abs(ms_o).max())
    np.testing.assert_allclose(Vs_h, Vs_o[0], rtol=1e-6, atol=1e-7 * np.abs(Vs_o).max())


def test_kernel_math_pinhole_vs_oracle(hc):
    from test_oracle import fly_cams
    cams = fly_cams()
    rng = np.random.default_rng(1)
    T = 200
    X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((T, 3)) * 1e-3, axis=0)
    var = rng.uniform(0.1, 1.0, (T, 6))
    y = oracle.project(cams, X) + rng.standard_normal((T, 6)) * np.sqrt(var)
    m0, S0, A, Q = X[0] + 0.01, np.eye(3) * 1e-2, np.eye(3), np.diag([1e-6, 2e-6, 1.5e-6])
    n_o, g_o = oracle.nll_grad(y[None], m0[None], S0[None], A[None], None, Q[None], var[0][None], 1.3, cams=cams)
    n_h, g_h = hc_nll(hc, 3, 6, y, m0, S0, A, Q, None, cams, var[0], 1.3, np.float64)
    np.testing.assert_allclose(n_h, n_o[0], rtol=1e-8)
    np.testing.assert_allclose(g_h, g_o[0], rtol=1e-6)
    ms_o, Vs_o = oracle.smooth(y[None], m0[None], S0[None], A[None], None, Q[None], var[None], 1.3, cams=cams,
                               dtype=np.float64)
    ms_h, Vs_h = hc_smooth(hc, 3, 6, y, var, m0, S0, A, Q, None, cams, 1.3, np.float64)
    np.testing.assert_allclose(ms_h, ms_o[0], rtol=1e-7)
    np.testing.assert_allclose(Vs_h, Vs_o[0], rtol=1e-5, atol=1e-6 * np.abs(Vs_o).max())
    # analytic Jacobian of the kernel vs the oracle's nested-dual Jacobian
    uv_o, J_o = oracle.project(cams, X[:20], jac=True)
    uv, J = np.zeros((20, 3, 2)), np.zeros((20, 3, 6))
    hc.hostcheck_project_f64(3, _p(np.ascontiguousarray(cams)), 20, _p(np.ascontiguousarray(X[:20])), _p(uv), _p(J))
    np.testing.assert_allclose(uv.reshape(20, 6), uv_o, rtol=1e-12)
    np.testing.assert_allclose(J.reshape(20, 6, 3), J_o, rtol=1e-9)


def test_kernel_adam_matches_oracle_trace(hc):
    raw = synth_singlecam(M=5, K=1, T=500, seed=2)
    r = oracle.singlecam(raw, dtype=np.float64, trace_cap=300)
    tr = r['info']['trace'][0]
    n = int(r['info']['iters'][0])
    loss = np.ascontiguousarray(tr[:n, 1])
    g = np.ascontiguousarray(tr[:n, 2] / 0.25)       # trace stores lr * grad
    s_log, it = np.zeros(1), np.zeros(1, np.int32)
    hc.hostcheck_adam_f64(n, _p(loss), _p(g), ctypes.c_double(tr[0, 0]), ctypes.c_double(0.25),
                          ctypes.c_double(1e-2), 300, _p(s_log), _p(it))
    assert it[0] == n
    np.testing.assert_allclose(np.exp(s_log[0]), r['s_finals'][0], rtol=1e-12)


# ------------------------------------------------------------------ C ABI surface
def test_library_exports_every_declared_symbol():
    from eks_b200 import _lib, build
    if not os.path.exists(build.LIB_PATH):
        build.build()
    L = _lib.lib()
    header = open(os.path.join(ROOT, 'include', 'eks_b200.h')).read()
    declared = set(re.findall(r'\b(eks_[a-z_A-Z0-9]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/eks_b200.h but not exported'
    assert L.eks_version() >= 100
    assert L.eks_ensemble_tile_frames(10, 20, 0, 0) > 0


def test_no_cpu_fallback_without_gpu():
    import torch
    import eks_b200
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from eks_b200._lib import EksB200Error
    ma = eks_b200.MarkerArray(np.random.rand(2, 1, 10, 2, 3), data_fields=['x', 'y', 'likelihood'])
    with pytest.raises(EksB200Error):
        eks_b200.ensemble(ma)
    with pytest.raises(EksB200Error):
        eks_b200.ensemble_kalman_smoother_singlecam(ma, ['a', 'b'])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'eks_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('# oracle', ''), f'{f} references oracle/'


def test_oracle_never_imports_product():
    """The checker must be independent of what it checks: nothing under oracle/ imports or loads eks_b200."""
    import re
    for f in os.listdir(os.path.join(ROOT, 'oracle')):
        if f.endswith(('.py', '.cpp', '.h', 'Makefile')):
            src = open(os.path.join(ROOT, 'oracle', f)).read()
            assert not re.search(r'(import|from)\s+eks_b200|libeks_b200', src), f'oracle/{f} depends on the product'


# ------------------------------------------------------------------ host logic (reference tests mirrored)
def test_crop_frames_semantics():
    """reference tests/test_utils.py:28-98."""
    from eks_b200.utils import crop_frames
    y = np.arange(20).reshape(10, 2)
    assert crop_frames(y, None) is y
    assert crop_frames(y, [(None, None)]) is y
    np.testing.assert_array_equal(crop_frames(y, [(2, 5)]), y[2:5])
    np.testing.assert_array_equal(crop_frames(y, [(None, 3), (7, None)]), np.concatenate([y[:3], y[7:]]))
    np.testing.assert_array_equal(crop_frames(y, [(6, 8), (1, 3)]), np.concatenate([y[1:3], y[6:8]]))
    for bad in ([(3, 3)], [(5, 2)], [(0, 11)], [(-1, 4)], [(0, 5), (4, 8)]):
        with pytest.raises(ValueError):
            crop_frames(y, bad)
    with pytest.raises(ValueError):
        crop_frames(y, [(1.5, 3)])
    with pytest.raises(ValueError):
        crop_frames(y, [[1, 3]])
    with pytest.raises(TypeError):
        crop_frames(y, ((1, 3),))
    for spec in ([(2, 5)], [(None, 3), (7, None)]):
        np.testing.assert_array_equal(crop_frames(y, spec), oracle.crop_frames(y, spec))


def test_marker_array_api():
    """reference tests/test_marker_array.py."""
    from eks_b200 import MarkerArray
    arr = np.random.rand(3, 2, 6, 4, 3)
    ma = MarkerArray(arr, data_fields=['x', 'y', 'likelihood'])
    assert ma.shape == (3, 2, 6, 4, 3) and (ma.n_models, ma.n_cameras, ma.n_frames, ma.n_keypoints) == (3, 2, 6, 4)
    assert ma.slice('keypoints', 1).shape == (3, 2, 6, 1, 3)
    assert ma.slice('cameras', [0, 1]).shape == (3, 2, 6, 4, 3)
    xy = ma.slice_fields('y', 'x')
    assert xy.data_fields == ['y', 'x'] and np.array_equal(xy.array[..., 0], arr[..., 1])
    st = MarkerArray.stack([ma, ma], 'models')
    assert st.shape == (6, 2, 6, 4, 3)
    sf = MarkerArray.stack_fields(ma.slice_fields('x'), ma.slice_fields('likelihood'))
    assert sf.data_fields == ['x', 'likelihood'] and sf.shape[-1] == 2
    ro = ma.reorder_data_fields(['likelihood', 'x', 'y'])
    assert np.array_equal(ro.array[..., 0], arr[..., 2])
    assert MarkerArray(shape=(1, 2, 5, 4, 2), data_fields=['x', 'y']).array.dtype == np.float32
    with pytest.raises(AssertionError):
        MarkerArray()
    with pytest.raises(AssertionError):
        ma.slice_fields('z')
    with pytest.raises(AssertionError):
        MarkerArray.stack([ma, ma.slice('frames', [0, 1])], 'models')


def test_input_dfs_to_marker_array_and_dlc_index():
    import pandas as pd
    from eks_b200.marker_array import input_dfs_to_markerArray
    from eks_b200.utils import make_dlc_pandas_index
    kps = ['a', 'b']
    rng = np.random.default_rng(0)
    dfs = [[pd.DataFrame({f'{k}_{f}': rng.random(7) for k in kps for f in ('x', 'y', 'likelihood')})
            for _ in range(3)] for _ in range(2)]
    ma = input_dfs_to_markerArray(dfs, kps, ['c0', 'c1'])
    assert ma.shape == (3, 2, 7, 2, 3) and ma.array.dtype == np.float64
    assert np.array_equal(ma.array[1, 1, :, 1, 2], dfs[1][1]['b_likelihood'].to_numpy())
    idx = make_dlc_pandas_index(kps, labels=['x', 'y'])
    assert list(idx.names) == ['scorer', 'bodyparts', 'coords'] and idx[0] == ('ensemble-kalman_tracker', 'a', 'x')


def test_center_predictions_matches_plain_numpy():
    from eks_b200 import MarkerArray
    from eks_b200.utils import center_predictions
    rng = np.random.default_rng(0)
    arr = rng.random((1, 2, 50, 3, 5))
    ema = MarkerArray(arr, data_fields=['x', 'y', 'var_x', 'var_y', 'likelihood'])
    mask, cen, good, means = center_predictions(ema, 50.0)
    mv = np.max(arr[..., 2:4], axis=(0, 1, 4))
    thr = np.percentile(mv, 50.0, axis=0)
    assert np.array_equal(mask, mv <= thr)
    nmin = min(int(m.sum()) for m in mask.T)
    assert good.shape == (1, 2, nmin, 3, 2)
    for k in range(3):
        idx = np.where(mask[:, k])[0][:nmin]
        mu = arr[0, :, idx, k, 0:2].mean(axis=0) if False else arr[0][:, idx][:, :, k, 0:2].mean(axis=1)
        np.testing.assert_allclose(means.array[0, :, 0, k, :], mu)
        np.testing.assert_allclose(cen.array[0, :, :, k, :], arr[0, :, :, k, 0:2] - mu[:, None, :])


def test_build_R_and_const_R_mirrors():
    from eks_b200.core import compute_initial_guesses, constant_R_from_timevarying
    from eks_b200.utils import build_R_from_vars, crop_R
    ev = np.abs(np.random.default_rng(0).standard_normal((2, 30, 3))) + 1e-3
    ev[0, 0, 0] = 0.0
    R = build_R_from_vars(ev)
    assert R.shape == (2, 30, 3, 3) and R[0, 0, 0, 0] == 1e-12 and R[1, 4, 0, 1] == 0
    Rc = crop_R(R, [(5, 10), (20, None)])
    assert Rc.shape == (2, 15, 3, 3)
    cR = constant_R_from_timevarying(R[0], min_var=1e-4)
    np.testing.assert_allclose(np.diag(cR), oracle.constant_R(np.clip(ev[0], 1e-12, None), 1e-4))
    np.testing.assert_allclose(compute_initial_guesses(ev[0]), oracle.compute_initial_guess(ev[0]))
    with pytest.raises(ValueError):
        compute_initial_guesses(ev[0][:1])


def test_fast_csv_reader_equals_pandas_path(tmp_path, monkeypatch):
    """pyarrow ingest returns the same flat DataFrames as pd.read_csv + convert_lp_dlc."""
    import pandas as pd
    from eks_b200 import io
    rng = np.random.default_rng(0)
    cols = pd.MultiIndex.from_product([['heatmap_tracker'], ['paw', 'nose', 'tail base'], ['x', 'y', 'likelihood']],
                                      names=['scorer', 'bodyparts', 'coords'])
    for i in range(3):
        df = pd.DataFrame(rng.random((257, 9)).astype(np.float32) * 300, columns=cols)
        df.iloc[5, 0] = np.nan
        df.to_csv(tmp_path / f'pred_rng={i}.csv')
    fast, kp_fast = io.format_data(str(tmp_path))
    monkeypatch.setenv('EKS_B200_PANDAS_CSV', '1')
    slow, kp_slow = io.format_data(str(tmp_path))
    assert kp_fast == kp_slow == ['paw', 'nose', 'tail base'] and len(fast) == len(slow) == 3
    for a, b in zip(fast, slow):
        assert list(a.columns) == list(b.columns) and (a.index == b.index).all()
        np.testing.assert_array_equal(a.to_numpy(), b.to_numpy())


def test_triangulation_templates_match_opencv():
    """The device triangulation (triangulate.cuh, compiled for the host) against the OpenCV primitives of the host
    mirror: cv2.undistortPoints, and CameraGroup.triangulate = pairwise cv2.triangulatePoints + nan-median."""
    import ctypes
    import cv2
    from eks_b200.multicam_smoother import CameraGroup, make_projection_from_camgroup
    from oracle import oracle
    hc = ctypes.CDLL(os.path.join(ROOT, 'tests', 'hostcheck', 'libhostcheck.so'))
    cg = CameraGroup.load(os.path.join(ROOT, 'tests', 'golden', 'fly_calibration.toml'))
    cams = np.ascontiguousarray(make_projection_from_camgroup(cg)[0].cams, dtype=np.float64)     # (V,29)
    V = cams.shape[0]
    rng = np.random.default_rng(0)
    n = 400
    X = np.array([-1.75, -0.30, 3.5]) + rng.normal(0, 0.05, (n, 3))
    uv = oracle.project(cams, X) + rng.normal(0, 0.7, (n, 2 * V))                                 # noisy pixels
    uv[5, 2:4] = np.nan                                                                           # one view missing
    # undistortion, camera by camera
    for c, cam in enumerate(cg.cameras):
        pts = np.ascontiguousarray(uv[:, 2 * c:2 * c + 2])
        ok = np.isfinite(pts).all(axis=1)
        got = np.empty_like(pts)
        hc.hc_undistort(cams[c].ctypes.data_as(ctypes.c_void_p), n, pts.ctypes.data_as(ctypes.c_void_p),
                        got.ctypes.data_as(ctypes.c_void_p))
        ref = cv2.undistortPoints(pts[ok].reshape(-1, 1, 2), cam.matrix, cam.dist).reshape(-1, 2)
        np.testing.assert_allclose(got[ok], ref, rtol=1e-12, atol=1e-14)
    # full triangulation
    got = np.empty((n, 3))
    uvc = np.ascontiguousarray(uv)
    hc.hc_triangulate_points(cams.ctypes.data_as(ctypes.c_void_p), V, n, uvc.ctypes.data_as(ctypes.c_void_p),
                             got.ctypes.data_as(ctypes.c_void_p))
    ref = cg.triangulate(np.stack([uv[:, 2 * c:2 * c + 2] for c in range(V)]), fast=True)
    assert np.isfinite(got).all() and np.isfinite(ref).all()
    np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9)
    assert np.abs(got - X).max() < 0.05                                                            # and it is the point


def test_fast_csv_writer_equals_pandas_to_csv(tmp_path):
    """io.write_dlc_csv (pyarrow) against DataFrame.to_csv on the smoothed-output layout: identical header rows, cells
    that parse to identical doubles (NaN as an empty cell), identical frames after the reference's own read call."""
    import pandas as pd
    from eks_b200.io import write_dlc_csv
    from eks_b200.utils import make_dlc_pandas_index
    labels = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
              'y_posterior_var']
    rng = np.random.default_rng(3)
    arr = rng.normal(100, 30, size=(300, 27))
    arr[5, 3] = np.nan
    arr[7, 0] = 1e-7
    arr[9, 1] = 123456789.125
    arr[11, 2] = -0.0
    df = pd.DataFrame(arr, columns=make_dlc_pandas_index(['a', 'b b', 'c'], labels=labels))
    pa_path, pd_path = str(tmp_path / 'fast.csv'), str(tmp_path / 'pandas.csv')
    write_dlc_csv(df, pa_path)
    df.to_csv(pd_path)
    la, lb = open(pa_path).readlines(), open(pd_path).readlines()
    assert la[:3] == lb[:3] and len(la) == len(lb)
    a = pd.read_csv(pa_path, header=[0, 1, 2], index_col=0, float_precision='round_trip')
    b = pd.read_csv(pd_path, header=[0, 1, 2], index_col=0, float_precision='round_trip')
    assert list(a.columns) == list(b.columns) and a.index.equals(b.index)
    np.testing.assert_array_equal(a.to_numpy(), b.to_numpy())
    np.testing.assert_array_equal(a.to_numpy(), arr)
    # a layout the fast writer does not handle falls back to pandas
    odd = pd.DataFrame({'s': ['x', 'y'], 'v': [1.0, 2.0]})
    write_dlc_csv(odd, str(tmp_path / 'odd.csv'))
    assert open(tmp_path / 'odd.csv').read() == odd.to_csv()


def test_ibl_paw_alignment_equals_the_reference_loop(monkeypatch, tmp_path):
    """fit_eks_multicam_ibl_paw's host stage (vectorised resampling of the right camera onto the left camera's
    timestamps, x flip, paw swap, zero likelihoods) against a loop restatement of eks/ibl_paw_multicam_smoother.py:
    178-236 on the bundled data/ibl-paw.  The smoother itself is replaced by a stub: no device needed."""
    import pandas as pd
    data = os.environ.get('EKS_DATA_DIR', '/root/reference/data') + '/ibl-paw'
    if not os.path.isdir(data):
        pytest.skip(f'{data} not present (build container only)')
    from scipy.interpolate import interp1d
    from eks_b200 import ibl_paw_multicam_smoother as paw
    from eks_b200.io import convert_lp_dlc
    captured = {}

    def stub(marker_array, keypoint_names, camera_names, **kw):
        captured['ma'], captured['kw'] = marker_array, kw
        T = marker_array.shape[2]
        cols = pd.MultiIndex.from_product([['s'], keypoint_names, ['x']])
        return [pd.DataFrame(np.zeros((T, len(keypoint_names))), columns=cols) for _ in camera_names], [1.0, 1.0], None
    monkeypatch.setattr(paw, 'ensemble_kalman_smoother_multicam', stub)
    dfs, s, input_dfs_list, bodyparts = paw.fit_eks_multicam_ibl_paw(data, str(tmp_path), var_mode='var')
    assert bodyparts == ['paw_l', 'paw_r'] and captured['kw']['inflate_vars_kwargs'] == {'likelihoods': None}
    ma = captured['ma']
    assert list(ma.data_fields) == ['x', 'y', 'likelihood'] and np.all(ma.array[..., 2] == 0)
    # loop restatement for the first model (files in os.listdir order, as the product and the reference read them)
    names = os.listdir(data)
    left = [n for n in names if 'timestamps' not in n and 'left' in n][0]
    right = [n for n in names if 'timestamps' not in n and 'left' not in n][0]
    tl = np.load(os.path.join(data, [n for n in names if 'timestamps' in n and 'left' in n][0]))
    tr = np.load(os.path.join(data, [n for n in names if 'timestamps' in n and 'left' not in n][0]))
    L = convert_lp_dlc(pd.read_csv(os.path.join(data, left), header=[0, 1, 2], index_col=0), bodyparts).to_numpy()
    R = convert_lp_dlc(pd.read_csv(os.path.join(data, right), header=[0, 1, 2], index_col=0), bodyparts).to_numpy()
    R = R[:, [3, 4, 5, 0, 1, 2]]                                   # right camera: paws swapped
    f = [interp1d(tr, R[:, j]) for j in range(6)]
    rows_l, rows_r = [], []
    for i, ts in enumerate(tl):
        if ts > tr[-1] or ts < tr[0]:
            continue
        rows_l.append(L[i, [0, 1, 3, 4]])
        r = np.array([f[j](ts) for j in [0, 1, 3, 4]])
        r[0], r[2] = 128 - r[0], 128 - r[2]
        rows_r.append(r)
    ref_l, ref_r = np.asarray(rows_l), np.asarray(rows_r)
    assert ma.shape[2] == len(ref_l)
    got_l = ma.array[0, 0, :, :, :2].reshape(len(ref_l), 4)       # (T, [paw_l x y, paw_r x y])
    got_r = ma.array[0, 1, :, :, :2].reshape(len(ref_r), 4)
    np.testing.assert_allclose(got_l, ref_l, rtol=0, atol=0)
    np.testing.assert_allclose(got_r, ref_r, rtol=1e-12, atol=1e-10)
