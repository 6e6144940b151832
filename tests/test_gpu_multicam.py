"""GPU parity tests for the multi-camera paths (linear PCA latent, calibrated pinhole EKF) through the
drop-in API and the generic kernels, against oracle-generated golden vectors (scripts/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
RTOL64, RTOL32 = 1e-5, 1e-3


def _ma(raw):
    from eks_b200 import MarkerArray
    return MarkerArray(np.ascontiguousarray(raw), data_fields=['x', 'y', 'likelihood'])


def _cam_array(camera_dfs, K):
    return np.stack([df.to_numpy().reshape(len(df), K, 9) for df in camera_dfs])  # (V,T,K,9)


def _check(a, b, rtol, label):
    for c in range(a.shape[-1]):
        x, y = a[..., c], b[..., c]
        scale = np.maximum(np.abs(y), 1e-6)
        err = np.max(np.abs(x - y) / scale)
        assert err <= rtol, f'{label}: column {c} rel err {err:.3e} > {rtol}'


@pytest.fixture(autouse=True)
def _precision():
    import eks_b200
    yield
    eks_b200.set_precision('float32')


def test_multicam_linear_fp64_matches_oracle():
    import eks_b200
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    g = load_golden('multicam_mirror_mouse_separate')
    eks_b200.set_precision('float64')
    kps = [str(k) for k in g['keypoints']]
    cams = [str(c) for c in g['cameras']]
    dfs, s, df3 = ensemble_kalman_smoother_multicam(_ma(g['raw'].astype(np.float64)), kps, cams,
                                                    quantile_keep_pca=95.0)
    np.testing.assert_allclose(s, g['s_f64'], rtol=RTOL64)
    _check(_cam_array(dfs, len(kps)), g['cam_out_f64'], RTOL64, 'multicam linear')
    _check(df3.to_numpy().reshape(-1, len(kps), 6), g['out3d_f64'], RTOL64, 'multicam linear 3d')
    assert list(dfs[0].columns.get_level_values('coords')[:9]) == [
        'x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
        'y_posterior_var']


def test_multicam_linear_fp32():
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    g = load_golden('multicam_mirror_mouse_separate')
    kps = [str(k) for k in g['keypoints']]
    cams = [str(c) for c in g['cameras']]
    # at the oracle's s: isolates the smoother/reprojection from the fp32 knife-edge stop rule
    dfs, s, df3 = ensemble_kalman_smoother_multicam(_ma(g['raw']), kps, cams, quantile_keep_pca=95.0,
                                                    smooth_param=list(g['s_f64']))
    assert np.allclose(s, g['s_f64'])
    out = _cam_array(dfs, len(kps))
    _check(out, g['cam_out_f64'], RTOL32, 'multicam linear fp32 at the oracle s')
    # optimised s in float32: the public call must return exactly what the device pipeline computes, and that
    # optimisation is held to the float32 stop protocol (tests/parity.py) against the float64 oracle trace
    import torch
    from eks_b200.pipeline import multicam_smooth_sessions
    from oracle import oracle
    from parity import fp32_stop_protocol
    dfs2, s2, _ = ensemble_kalman_smoother_multicam(_ma(g['raw']), kps, cams, quantile_keep_pca=95.0)
    res = multicam_smooth_sessions(torch.as_tensor(g['raw']).cuda()[None], quantile_keep_pca=95.0, trace_cap=300)
    np.testing.assert_array_equal(s2, res.s_finals[0].cpu().numpy())
    ref = oracle.multicam(g['raw'].astype(np.float64), quantile_keep_pca=95.0, dtype=np.float64, trace_cap=300)
    ref32 = oracle.multicam(g['raw'].astype(np.float64), quantile_keep_pca=95.0, dtype=np.float32, trace_cap=300)
    trace = multicam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    it = res.iters[0].cpu().numpy()
    for k in range(len(kps)):
        # kappa: the float32 model inputs (PCA components, centring offsets: 6e-8 relative) move each innovation of a
        # +-100 px coordinate by ~6e-6 px, i.e. the 2000-term loss by ~25 float32 ulps at this tiny T (501 frames); the
        # float32 oracle itself is 25-135 ulps from the float64 one here, and the product must not be further than that
        fp32_stop_protocol(f'mirror-mouse-separate kp{k}', trace[k], it[k], ref['info']['trace'][k],
                           ref['info']['iters'][k], kappa=64.0, ref32_trace=ref32['info']['trace'][k],
                           n_ref32=ref32['info']['iters'][k])
    print('[parity fp32] mirror-mouse-separate |ds|/s vs fp64 oracle', np.abs(s2 - ref['s_finals']) / ref['s_finals'])


def test_multicam_nonlinear_fp64_matches_oracle():
    import eks_b200
    from eks_b200.multicam_smoother import CameraGroup, ensemble_kalman_smoother_multicam
    g = load_golden('multicam_fly_nonlinear')
    eks_b200.set_precision('float64')
    kps = [str(k) for k in g['keypoints']]
    cams = [str(c) for c in g['cameras']]
    camgroup = CameraGroup.load(os.path.join(GOLDEN, 'fly_calibration.toml'))
    dfs, s, df3 = ensemble_kalman_smoother_multicam(_ma(g['raw'].astype(np.float64)), kps, cams,
                                                    quantile_keep_pca=95.0, camgroup=camgroup)
    np.testing.assert_allclose(s, g['s_f64'], rtol=1e-4)
    _check(_cam_array(dfs, len(kps)), g['cam_out_f64'], 1e-4, 'multicam nonlinear')
    _check(df3.to_numpy().reshape(-1, len(kps), 6), g['out3d_f64'], 1e-4, 'multicam nonlinear 3d')


def test_multicam_nonlinear_with_inflation_fp64_matches_oracle():
    """Calibrated model with inflate_vars=True (the CLI default; reference integration call
    tests/integration/test_multicam.py:31-41): Mahalanobis inflation on the centred predictions runs on the device for
    this branch too, the smoother uses the inflated variances, the output reports the raw ensemble variances."""
    import eks_b200
    from eks_b200.multicam_smoother import CameraGroup, ensemble_kalman_smoother_multicam
    from oracle import oracle
    g = load_golden('multicam_fly_nonlinear')
    cal = os.path.join(GOLDEN, 'fly_calibration.toml')
    raw = g['raw'].astype(np.float64).copy()
    raw[:, 1, 100:140, 0, 0:2] += 25.0            # one camera disagrees for a while: these frames must be inflated
    ref = oracle.multicam(raw, camgroup=cal, quantile_keep_pca=95.0, dtype=np.float64, inflate_vars=True)
    ref_plain = oracle.multicam(raw, camgroup=cal, quantile_keep_pca=95.0, dtype=np.float64, inflate_vars=False)
    assert np.abs(ref['cam_out'] - ref_plain['cam_out']).max() > 1e-3, 'the inflation did not change anything'
    eks_b200.set_precision('float64')
    kps = [str(k) for k in g['keypoints']]
    cams = [str(c) for c in g['cameras']]
    try:
        dfs, s, df3 = ensemble_kalman_smoother_multicam(_ma(raw), kps, cams, quantile_keep_pca=95.0,
                                                        camgroup=CameraGroup.load(cal), inflate_vars=True)
    finally:
        eks_b200.set_precision('float32')
    np.testing.assert_allclose(s, ref['s_finals'], rtol=1e-4)
    _check(_cam_array(dfs, len(kps)), ref['cam_out'], 1e-4, 'multicam nonlinear + inflation')


def test_fixed_smooth_param_and_latent_dims():
    """reference tests/test_multicam_smoother.py:196-228: n_latent in {3,4,5} with 4 cameras, s echoed."""
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    rng = np.random.default_rng(0)
    M, V, T, K = 3, 4, 120, 2
    lat = np.cumsum(rng.standard_normal((T, K, 3)), axis=0)
    W = rng.standard_normal((2 * V, 3))
    obs = np.einsum('tkd,od->tko', lat, W).reshape(T, K, V, 2).transpose(2, 0, 1, 3)      # (V,T,K,2)
    raw = np.zeros((M, V, T, K, 3))
    raw[..., 0:2] = obs[None] + rng.standard_normal((M, V, T, K, 2)) * 0.3 + 100.0
    raw[..., 2] = rng.uniform(0.8, 1.0, (M, V, T, K))
    for n_latent in (3, 4, 5):
        dfs, s, df3 = ensemble_kalman_smoother_multicam(_ma(raw), ['a', 'b'], ['c0', 'c1', 'c2', 'c3'],
                                                        smooth_param=10.0, n_latent=n_latent)
        assert len(dfs) == V and dfs[0].shape == (T, K * 9) and np.all(s == 10.0)
        assert np.all(np.isfinite(dfs[0].to_numpy()))
    with pytest.raises(ValueError):
        ensemble_kalman_smoother_multicam(_ma(raw), ['a', 'b'], [])


def test_run_kalman_smoother_rejects_python_callable():
    import eks_b200
    K, T = 1, 20
    with pytest.raises(TypeError):
        eks_b200.run_kalman_smoother(np.zeros((K, T, 6)), np.zeros((K, 3)), np.tile(np.eye(3), (K, 1, 1)),
                                     np.tile(np.eye(3), (K, 1, 1)), None, np.tile(np.eye(3), (K, 1, 1)),
                                     np.ones((T, K, 6)), h_fn=lambda x: x)


def test_core_api_shapes_and_blocks():
    """reference tests/test_core.py:155-232 (slow path fills s_finals; members share s; s_frames)."""
    import eks_b200
    rng = np.random.default_rng(0)
    K, T = 3, 30
    ys = rng.standard_normal((K, T, 2))
    eye = np.tile(np.eye(2), (K, 1, 1))
    Rs = np.tile(np.eye(2), (K, T, 1, 1))
    for s_frames in (None, [(0, 15)]):
        s_finals = np.empty(K)
        eks_b200.optimize_smooth_param(ys=ys, m0s=np.zeros((K, 2)), S0s=eye, As=eye, Cs=eye, Qs=eye, Rs=Rs,
                                       blocks=[[0, 1], [2]], s_finals=s_finals, s_frames=s_frames,
                                       s_guess_per_k=np.ones(K), safety_cap=5)
        assert np.all(np.isfinite(s_finals)) and np.all(s_finals > 0) and s_finals[0] == s_finals[1]
    s, ms, Vs = eks_b200.run_kalman_smoother(ys, np.zeros((K, 2)), eye, eye, eye, eye, np.ones((T, K, 2)),
                                             smooth_param=[1.0, 2.0, 3.0])
    assert ms.shape == (K, T, 2) and Vs.shape == (K, T, 2, 2) and list(s) == [1.0, 2.0, 3.0]


def _linear_case(K, T, seed, r_scale=1.0, q_scale=1.0):
    rng = np.random.default_rng(seed)
    V = 2
    W = np.linalg.qr(rng.standard_normal((2 * V, 3)))[0]
    lat = np.cumsum(rng.normal(0, 0.3 * np.sqrt(q_scale), (K, T, 3)), axis=1)
    ev = rng.uniform(0.1, 0.6, (T, K, 2 * V)) * r_scale
    ys = lat @ W.T + rng.standard_normal((K, T, 2 * V)) * np.sqrt(np.swapaxes(ev, 0, 1))
    ys -= ys.mean(axis=1, keepdims=True)
    Q = np.array([[1.0, 0.2, 0.05], [0.2, 0.7, 0.1], [0.05, 0.1, 0.5]])
    return (ys, np.zeros((K, 3)), np.tile(np.diag([5.0, 4.0, 3.0]), (K, 1, 1)), np.tile(np.eye(3), (K, 1, 1)),
            np.tile(W, (K, 1, 1)), np.tile(Q, (K, 1, 1)), ev)


@pytest.mark.parametrize('T', [512, 777, 4096, 13001])
def test_run_parallel_generic_linear_matches_oracle(T):
    """T >= 512 switches the generic path to verified run-parallel execution (generic_runs.cu): results must
    equal the sequential oracle -- identical Adam iteration counts in fp64."""
    import eks_b200
    from oracle import oracle
    args = _linear_case(3, T, seed=T)
    eks_b200.set_precision('float64')
    s, ms, Vs = eks_b200.run_kalman_smoother(*args)
    s_o, ms_o, Vs_o, info = oracle.run_kalman_smoother(*args, dtype=np.float64)
    np.testing.assert_allclose(s, s_o, rtol=RTOL64)
    np.testing.assert_allclose(ms, ms_o, rtol=RTOL64, atol=1e-7 * np.abs(ms_o).max())
    np.testing.assert_allclose(Vs, Vs_o, rtol=RTOL64, atol=1e-9 * np.abs(Vs_o).max())


def test_run_parallel_generic_slow_forgetting_escalates():
    """huge observation noise + small process noise: the filter forgets slowly, the 64-frame warm-up fails the
    boundary verification and must be escalated; the result still equals the sequential oracle."""
    import eks_b200
    from oracle import oracle
    args = _linear_case(2, 6000, seed=3, r_scale=2000.0, q_scale=1e-3)
    eks_b200.set_precision('float64')
    s, ms, Vs = eks_b200.run_kalman_smoother(*args, smooth_param=1e-3)
    s_o, ms_o, Vs_o, _ = oracle.run_kalman_smoother(*args, smooth_param=1e-3, dtype=np.float64)
    np.testing.assert_allclose(ms, ms_o, rtol=RTOL64, atol=1e-7 * np.abs(ms_o).max())
    np.testing.assert_allclose(Vs, Vs_o, rtol=RTOL64, atol=1e-9 * np.abs(Vs_o).max())
    s, ms, Vs = eks_b200.run_kalman_smoother(*args)
    s_o, ms_o, Vs_o, _ = oracle.run_kalman_smoother(*args, dtype=np.float64)
    np.testing.assert_allclose(s, s_o, rtol=RTOL64)
    np.testing.assert_allclose(ms, ms_o, rtol=RTOL64, atol=1e-7 * np.abs(ms_o).max())


def test_run_parallel_generic_pinhole_matches_oracle():
    import eks_b200
    from eks_b200.core import PinholeProjection
    from oracle import oracle
    from test_oracle import fly_cams
    cams = fly_cams()
    rng = np.random.default_rng(2)
    K, T = 2, 5000
    X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((K, T, 3)) * 1e-3, axis=1)
    ys = np.stack([oracle.project(cams, X[k]) for k in range(K)]) + rng.standard_normal((K, T, 6)) * 0.5
    ev = rng.uniform(0.1, 0.5, (T, K, 6))
    args = (ys, X[:, 0, :] + 0.01, np.tile(np.eye(3) * 1e-2, (K, 1, 1)), np.tile(np.eye(3), (K, 1, 1)), None,
            np.tile(np.eye(3) * 1e-6, (K, 1, 1)), ev)
    eks_b200.set_precision('float64')
    s, ms, Vs = eks_b200.run_kalman_smoother(*args, h_fn=PinholeProjection(cams))
    s_o, ms_o, Vs_o, _ = oracle.run_kalman_smoother(*args, cams=cams, dtype=np.float64)
    np.testing.assert_allclose(s, s_o, rtol=1e-4)
    np.testing.assert_allclose(ms, ms_o, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(Vs, Vs_o, rtol=1e-3, atol=1e-6 * np.abs(Vs_o).max())


def test_device_triangulation_matches_oracle():
    """eks_triangulate_mean (one thread per keypoint-frame, fp64) against the ORACLE's independent NumPy restatement of
    triangulate_3d_models(...).mean(axis=0) (oracle.triangulate_3d_models: undistortion iterations + pairwise SVD DLT +
    nan-median; itself pinned to OpenCV in tests/test_oracle.py) on the fly fixture, with a missing view injected."""
    import os
    import torch
    from conftest import GOLDEN
    from eks_b200 import ops
    from oracle import oracle
    g = load_golden('multicam_fly_nonlinear')
    raw = g['raw'].astype(np.float64)[:, :, :200].copy()          # (M,V,T,K,3)
    raw[1, 2, 17, 0, :2] = np.nan
    calib = oracle.load_calibration(os.path.join(GOLDEN, 'fly_calibration.toml'))
    ref = oracle.triangulate_3d_models(raw, calib).mean(axis=0)
    cams = torch.as_tensor(oracle.pack_calibration(calib))
    for dt in (torch.float64, torch.float32):
        got = ops.triangulate_mean(torch.as_tensor(raw).to('cuda', dt).contiguous(), cams).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-6 if dt == torch.float64 else 1e-4, atol=1e-8)


def test_shared_s_block_of_linear_sequences_through_the_lag_path():
    """blocks=[[0, 1], [2]] on a D = 3 / O = 4 model with A = I and 3000 frames: the lag-statistics optimiser sums the
    member losses of a shared-s block (eks/core.py:474-476) and must reproduce the oracle's iteration counts and s."""
    import eks_b200
    from oracle import oracle
    args = _linear_case(3, 3000, seed=21)
    eks_b200.set_precision('float64')
    try:
        s, ms, Vs = eks_b200.run_kalman_smoother(*args, blocks=[[0, 1], [2]])
    finally:
        eks_b200.set_precision('float32')
    s_o, ms_o, Vs_o, info = oracle.run_kalman_smoother(*args, blocks=[[0, 1], [2]], dtype=np.float64)
    assert s[0] == s[1]
    np.testing.assert_allclose(s, s_o, rtol=1e-5)
    np.testing.assert_allclose(ms, ms_o, rtol=1e-5, atol=1e-6)


def test_pca_object_runs_on_the_device_and_equals_the_host_prestage(monkeypatch):
    """pca_object (a fitted sklearn PCA shared by all keypoints, eks/stats.py:52-56): the device pipeline takes its
    mean_ / components_ instead of fitting per keypoint; same results as the host pre-stage (NumPy centring +
    sklearn transform + run_kalman_smoother)."""
    import eks_b200
    from sklearn.decomposition import PCA
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    from test_gpu_multicam_pipeline import synth_multicam
    raw = synth_multicam(M=4, V=2, K=2, T=900, seed=17)
    rng = np.random.default_rng(1)
    lat = np.cumsum(rng.normal(0, 0.3, (4000, 3)), axis=0)
    pca = PCA(n_components=3).fit(lat @ rng.standard_normal((3, 4)))
    eks_b200.set_precision('float64')
    try:
        dfs_d, s_d, _ = ensemble_kalman_smoother_multicam(_ma(raw), ['a', 'b'], ['c0', 'c1'], quantile_keep_pca=80.0,
                                                          pca_object=pca)
        monkeypatch.setenv('EKS_B200_HOST_PRESTAGE', '1')
        dfs_h, s_h, _ = ensemble_kalman_smoother_multicam(_ma(raw), ['a', 'b'], ['c0', 'c1'], quantile_keep_pca=80.0,
                                                          pca_object=pca)
    finally:
        eks_b200.set_precision('float32')
    np.testing.assert_allclose(s_d, s_h, rtol=1e-6)
    for a, b in zip(dfs_d, dfs_h):
        np.testing.assert_allclose(a.to_numpy(), b.to_numpy(), rtol=1e-6, atol=1e-7)
