"""GPU parity tests of the device-resident multi-camera pipeline: centring, PCA moments, latent
initialisation and the whole linear multicam smoother against the CPU oracle (NumPy + scikit-learn PCA)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL64, RTOL32 = 1e-5, 1e-3


def synth_multicam(M=5, V=2, K=3, T=3000, seed=0, outlier_frac=0.0):
    rng = np.random.default_rng(seed)
    lat = np.cumsum(rng.normal(0, 0.3, (T, K, 3)), axis=0)
    W = rng.standard_normal((K, 2 * V, 3))
    truth = np.einsum('tkl,kol->tko', lat, W) + rng.uniform(50, 300, (1, K, 2 * V))
    occ = rng.random((T, K)) < 0.05
    sigma = np.where(occ, 4.0, 0.5)[:, :, None]
    if outlier_frac > 0:   # confident but geometrically inconsistent predictions in one view (all seeds agree)
        out = rng.random((T, K)) < outlier_frac
        view = rng.integers(0, V, (T, K))
        for v in range(V):
            truth[:, :, 2 * v] += np.where(out & (view == v), rng.uniform(8, 25, (T, K)), 0.0)
    raw = np.empty((M, V, T, K, 3))
    for m in range(M):
        noisy = truth + rng.standard_normal((T, K, 2 * V)) * sigma
        raw[m, :, :, :, :2] = noisy.reshape(T, K, V, 2).transpose(2, 0, 1, 3)
        raw[m, :, :, :, 2] = np.where(occ[None], rng.uniform(0.1, 0.5, (V, T, K)), rng.uniform(0.9, 1.0, (V, T, K)))
    return raw.astype(np.float32).astype(np.float64)


def _run(raw, dtype, **kw):
    from eks_b200.pipeline import multicam_smooth_sessions
    t = torch.as_tensor(raw).cuda().to(dtype)
    res = multicam_smooth_sessions(t[None], dtype=dtype, **kw)
    torch.cuda.synchronize()
    return res


def _cam_out(res, s=0):
    return res.out[s].permute(1, 3, 0, 2).double().cpu().numpy()     # (V,T,K,9)


def _check(out, ref, rtol, label):
    for c in range(9):
        a, b = out[..., c], ref[..., c]
        scale = np.maximum(np.abs(b), 1e-6 if c >= 5 else 1.0)
        err = np.max(np.abs(a - b) / scale)
        assert err <= rtol, f'{label}: column {c} rel err {err:.3e} > {rtol}'


@pytest.mark.parametrize('V,T,q', [(2, 3000, 50.0), (3, 2500, 95.0), (2, 1024, 100.0), (4, 777, 25.0)])
def test_prestage_matches_oracle_fp64(V, T, q):
    from eks_b200 import ops
    from eks_b200.pipeline import pca_from_moments
    from oracle import oracle
    raw = synth_multicam(V=V, T=T, seed=T)
    M, _, _, K, _ = raw.shape
    ens = oracle.ensemble(raw, dtype=np.float64)                         # (V,T,K,5)
    mask, cen, good, means, n_used = oracle.mc_center_predictions(ens, q)
    m0s, S0s, As, Qs, Cs = oracle.mc_pca_init(mask, cen, good, 3)
    dev = torch.device('cuda')
    planes = lambda lo: torch.as_tensor(np.ascontiguousarray(
        np.transpose(ens[..., lo:lo + 2], (2, 0, 3, 1)).reshape(K, 2 * V, T))).to(dev)    # [K][2V][T]
    yv = ops.PlaneView(planes(0), 2 * V * T, [o * T for o in range(2 * V)])
    vv = ops.PlaneView(planes(2), 2 * V * T, [o * T for o in range(2 * V)])
    ymean, n_good, ws = ops.mc_center(yv, vv, 1, K, T, q)
    np.testing.assert_array_equal(n_good[:, 0].cpu().numpy(), mask.sum(axis=0))
    np.testing.assert_array_equal(n_good[:, 1].cpu().numpy(), np.full(K, n_used))
    np.testing.assert_allclose(ymean.cpu().numpy(), means, rtol=1e-12)
    mom = ops.mc_pca_moments(yv, ymean, T, ws).cpu().numpy()
    pm, comps = pca_from_moments(mom, 2 * V, 3)
    np.testing.assert_allclose(np.swapaxes(comps, 1, 2), Cs, rtol=1e-7, atol=1e-9)
    C = torch.as_tensor(np.ascontiguousarray(np.swapaxes(comps, 1, 2))).to(dev)
    S0, Q = ops.mc_latent_init(yv, ymean, torch.as_tensor(pm).to(dev), C, T, ws)
    np.testing.assert_allclose(S0.cpu().numpy(), S0s, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(Q.cpu().numpy(), Qs, rtol=1e-7, atol=1e-10)


def test_mirror_mouse_separate_fp64_matches_golden():
    g = load_golden('multicam_mirror_mouse_separate')
    res = _run(g['raw'].astype(np.float64), torch.float64, quantile_keep_pca=95.0)
    assert list(res.iters[0].cpu().numpy()) == list(g['iters_f64'])
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), g['s_f64'], rtol=RTOL64)
    _check(_cam_out(res), g['cam_out_f64'], RTOL64, 'mirror-mouse-separate fp64')
    K = res.ms.shape[0]
    out3d = g['out3d_f64']                                                 # (T,K,6)
    ms = res.ms.cpu().numpy()
    Vs = res.Vs.cpu().numpy()
    for k in range(K):
        np.testing.assert_allclose(ms[k], out3d[:, k, :3], rtol=RTOL64, atol=1e-6 * np.abs(out3d[:, k, :3]).max())
        for d in range(3):
            np.testing.assert_allclose(Vs[k][:, d, d], out3d[:, k, 3 + d], rtol=RTOL64)


def test_mirror_mouse_separate_fp32():
    g = load_golden('multicam_mirror_mouse_separate')
    from eks_b200.pipeline import multicam_smooth_sessions
    from oracle import oracle
    from parity import fp32_stop_protocol
    res = _run(g['raw'], torch.float32, quantile_keep_pca=95.0, trace_cap=300)
    ref_t = oracle.multicam(g['raw'].astype(np.float64), quantile_keep_pca=95.0, dtype=np.float64, trace_cap=300)
    ref32 = oracle.multicam(g['raw'].astype(np.float64), quantile_keep_pca=95.0, dtype=np.float32, trace_cap=300)
    trace = multicam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    it = res.iters[0].cpu().numpy()
    for k in range(trace.shape[0]):   # float32 stop protocol (tests/parity.py) instead of a loose tolerance on s
        # kappa: the float32 model inputs (PCA components, centring offsets: 6e-8 relative) move each innovation of a
        # +-100 px coordinate by ~6e-6 px, i.e. the 2000-term loss by ~25 float32 ulps at this tiny T (501 frames); the
        # float32 oracle itself is 25-135 ulps from the float64 one here, and the product must not be further than that
        fp32_stop_protocol(f'mirror-mouse-separate kp{k}', trace[k], it[k], ref_t['info']['trace'][k],
                           ref_t['info']['iters'][k], kappa=64.0, ref32_trace=ref32['info']['trace'][k],
                           n_ref32=ref32['info']['iters'][k])
    ref = oracle.multicam(g['raw'].astype(np.float64), quantile_keep_pca=95.0, dtype=np.float64,
                          smooth_param=res.s_finals[0].cpu().numpy())
    _check(_cam_out(res), ref['cam_out'], RTOL32, 'mirror-mouse-separate fp32')


@pytest.mark.parametrize('V', [2, 3])
def test_synthetic_vs_oracle_fp64(V):
    from oracle import oracle
    raw = synth_multicam(V=V, T=2200, seed=11 + V)
    res = _run(raw, torch.float64, quantile_keep_pca=50.0)
    ref = oracle.multicam(raw, quantile_keep_pca=50.0, dtype=np.float64)
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters'])
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    _check(_cam_out(res), ref['cam_out'], RTOL64, f'synthetic V={V}')


def test_sessions_batch_equals_single():
    from eks_b200.pipeline import multicam_smooth_sessions
    raws = [synth_multicam(T=1500, seed=s) for s in (1, 2)]
    both = multicam_smooth_sessions(torch.as_tensor(np.stack(raws)).cuda().float())
    for i, r in enumerate(raws):
        one = _run(r, torch.float32)
        torch.testing.assert_close(both.out[i], one.out[0], rtol=0, atol=0)
        torch.testing.assert_close(both.s_finals[i], one.s_finals[0], rtol=0, atol=0)


def test_fixed_smooth_param_and_spans():
    from oracle import oracle
    raw = synth_multicam(T=1800, seed=5)
    res = _run(raw, torch.float64, smooth_param=3.0)
    ref = oracle.multicam(raw, dtype=np.float64, smooth_param=3.0)
    _check(_cam_out(res), ref['cam_out'], RTOL64, 'fixed s')
    res = _run(raw, torch.float64, spans=[(100, 900), (1200, 1800)])
    ref = oracle.multicam(raw, dtype=np.float64, s_frames=[(100, 900), (1200, None)])
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters'])
    _check(_cam_out(res), ref['cam_out'], RTOL64, 's_frames')


@pytest.mark.parametrize('V,kw', [(2, {}), (3, {}), (2, {'likelihoods': True, 'likelihood_threshold': 0.6}),
                                  (3, {'v_quantile_threshold': 80.0, 'epsilon': 1e-4})])
def test_variance_inflation_matches_oracle_fp64(V, kw):
    """Mahalanobis variance inflation on the device (FactorAnalysis moments + per-frame kernel) vs the oracle, which
    runs scikit-learn's FactorAnalysis and the reference's per-frame algebra in NumPy."""
    from oracle import oracle
    raw = synth_multicam(V=V, T=2600, seed=21 + V, outlier_frac=0.03)
    res = _run(raw, torch.float64, inflate_vars=True, inflate_vars_kwargs=dict(kw))
    ref = oracle.multicam(raw, dtype=np.float64, inflate_vars=True, inflate_vars_kwargs=dict(kw))
    out = _cam_out(res)
    infl = ref['cam_out'][..., 5:7]
    base = oracle.multicam(raw, dtype=np.float64, smooth_param=1.0)['cam_out'][..., 5:7]
    assert (infl != base).sum() > 50, 'test data must trigger inflation'
    np.testing.assert_allclose(out[..., 5:7], infl, rtol=1e-12)           # identical inflation decisions
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters'])
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    _check(out, ref['cam_out'], RTOL64, f'inflation V={V}')


def test_variance_inflation_fixed_loading_matrix():
    from oracle import oracle
    raw = synth_multicam(V=2, T=1500, seed=8, outlier_frac=0.04)
    rng = np.random.default_rng(0)
    Wm = np.linalg.qr(rng.standard_normal((4, 3)))[0] * 5.0
    kw = dict(loading_matrix=Wm, mean=np.zeros(4))
    res = _run(raw, torch.float64, inflate_vars=True, inflate_vars_kwargs=dict(kw), smooth_param=2.0)
    ref = oracle.multicam(raw, dtype=np.float64, inflate_vars=True, inflate_vars_kwargs=dict(kw), smooth_param=2.0)
    _check(_cam_out(res), ref['cam_out'], RTOL64, 'inflation with a given loading matrix')


def test_public_entry_point_with_inflation():
    """ensemble_kalman_smoother_multicam(inflate_vars=True) (the CLI default) runs on the device pipeline."""
    import eks_b200
    from eks_b200.marker_array import MarkerArray
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    from oracle import oracle
    raw = synth_multicam(V=2, T=1200, seed=4, outlier_frac=0.03)
    eks_b200.set_precision('float64')
    try:
        dfs, s, df3d = ensemble_kalman_smoother_multicam(
            MarkerArray(raw, data_fields=['x', 'y', 'likelihood'], dtype=np.float64), ['a', 'b', 'c'], ['top', 'bot'],
            inflate_vars=True)
    finally:
        eks_b200.set_precision('float32')
    ref = oracle.multicam(raw, dtype=np.float64, inflate_vars=True)
    np.testing.assert_allclose(s, ref['s_finals'], rtol=RTOL64)
    for c in range(2):
        _check(dfs[c].to_numpy().reshape(1200, 3, 9), ref['cam_out'][c], RTOL64, f'camera {c}')


def test_fit_eks_mirrored_multicam(tmp_path):
    """fit_eks_mirrored_multicam (reference eks/multicam_smoother.py:37-153): CSVs with '{bodypart}_{camera}' keypoints
    are split per camera, smoothed, and written back as one CSV with the camera suffix restored."""
    import pandas as pd
    from eks_b200.marker_array import MarkerArray
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam, fit_eks_mirrored_multicam
    from eks_b200.utils import make_dlc_pandas_index
    raw = synth_multicam(M=3, V=2, K=2, T=900, seed=31)                    # (M,V,T,K,3)
    cams, parts = ['top', 'bot'], ['paw1', 'paw2']
    kps = [f'{p}_{c}' for c in cams for p in parts]
    for m in range(3):
        cols = np.concatenate([raw[m, v, :, k, :] for v in range(2) for k in range(2)], axis=1)
        pd.DataFrame(cols, columns=make_dlc_pandas_index(kps)).to_csv(tmp_path / f'pred_rng={m}.csv')
    out = tmp_path / 'out' / 'mirrored.csv'
    final_df, s, dfs, bps = fit_eks_mirrored_multicam(str(tmp_path), str(out), camera_names=cams, smooth_param=3.0)
    assert bps == parts and out.exists() and final_df.shape == (900, 2 * 2 * 9)
    assert [c[1] for c in final_df.columns[::9]] == ['paw1_top', 'paw2_top', 'paw1_bot', 'paw2_bot']
    ref_dfs, _, _ = ensemble_kalman_smoother_multicam(MarkerArray(raw.astype(np.float32), data_fields=['x', 'y', 'likelihood']),
                                                      parts, cams, smooth_param=3.0)
    np.testing.assert_allclose(final_df.to_numpy(), np.concatenate([d.to_numpy() for d in ref_dfs], axis=1), rtol=1e-6)


@pytest.mark.parametrize('T', [2, 7, 1000, 150_001])
def test_geometric_init_on_device_matches_oracle(T):
    """eks_geometric_init (means, nan-variances, median / MAD of the lag-1 differences by the exact radix select, short
    sequences through the multi-pass select, long ones through the one-pass bracketed median) against the oracle's
    NumPy restatement of initialize_kalman_filter_geometric (eks/multicam_smoother.py:600-650)."""
    from eks_b200 import ops
    from oracle import oracle
    rng = np.random.default_rng(T)
    B = 4
    tri = np.cumsum(rng.normal(0, [1e-3, 2e-3, 5e-4], size=(B, T, 3)), axis=1) + rng.uniform(-2, 4, size=(B, 1, 3))
    tri[0, :, 1] += (rng.random(T) < 0.05) * rng.normal(0, 0.3, size=T)       # heavy-tailed jumps: MAD != std
    if T > 20:
        tri[2, T // 2, 2] = np.nan                                            # np.median propagates, nanvar does not
    m0, S0d, Qd = ops.geometric_init(torch.as_tensor(tri).cuda())
    with np.errstate(all='ignore'):
        m0s, S0s, _, Qs, _ = oracle.geometric_init(tri)
    np.testing.assert_allclose(m0.cpu().numpy(), m0s, rtol=1e-13)
    np.testing.assert_allclose(S0d.cpu().numpy(), np.diagonal(S0s, axis1=1, axis2=2), rtol=1e-11)
    np.testing.assert_allclose(Qd.cpu().numpy(), np.diagonal(Qs, axis1=1, axis2=2), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize('V,dtype,T', [(2, torch.float64, 3000), (3, torch.float64, 2500), (4, torch.float64, 2400),
                                       (2, torch.float32, 30_000)])
def test_lag_statistics_optimiser_equals_run_parallel_path(V, dtype, T, monkeypatch):
    """lin_lag.cu (one pass over the observations: O x O x W lag statistics of the stationary signal + closed-form NLL
    in a persistent Adam kernel) against the run-parallel evaluation of generic_runs.cu on the same data: same
    iteration counts and, iterate by iterate, the same loss / gradient (float64: to 1e-9 relative)."""
    from eks_b200.pipeline import multicam_smooth_sessions
    raw = synth_multicam(M=4, V=V, K=3, T=T, seed=40 + V)
    x = torch.as_tensor(raw).cuda()[None].to(dtype)
    monkeypatch.setenv('EKS_NO_LINLAG', '1')
    # float32 mode is checked against the float64 run-parallel evaluation of the same (float32-exact) data
    ref = multicam_smooth_sessions(x.double(), dtype=torch.float64, trace_cap=300)
    tr_ref = multicam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    monkeypatch.delenv('EKS_NO_LINLAG')
    monkeypatch.setenv('EKS_DEBUG_RUNS', '1')
    res = multicam_smooth_sessions(x, dtype=dtype, trace_cap=300)
    tr = multicam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    assert multicam_smooth_sessions.last_opt['launches'] == 5, \
        'the lag-statistics path did not run (fell back to the run-parallel path)'
    it, it_ref = res.iters[0].cpu().numpy(), ref.iters[0].cpu().numpy()
    if dtype == torch.float64:
        np.testing.assert_array_equal(it, it_ref)
        for k in range(raw.shape[3]):
            n = it[k]
            np.testing.assert_allclose(tr[k, :n, 1], tr_ref[k, :n, 1], rtol=1e-9)
            np.testing.assert_allclose(tr[k, :n, 2], tr_ref[k, :n, 2], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(res.s_finals.cpu().numpy(), ref.s_finals.cpu().numpy(), rtol=1e-8)
    else:   # float32 mode: the lag path evaluates the float32 data and model in float64 arithmetic
        from parity import fp32_stop_protocol
        for k in range(raw.shape[3]):
            fp32_stop_protocol(f'lag (float32) vs runs (float64) kp{k}', tr[k], it[k], tr_ref[k], it_ref[k])


def test_lag_statistics_path_with_one_offset_span_fp64():
    """s_frames with ONE span that does not start at frame 0: the lag-statistics optimiser runs on the cropped frames
    (t_begin > 0, n < T) and must still reproduce the oracle's iteration counts; the final pass covers every frame."""
    from eks_b200.pipeline import multicam_smooth_sessions
    from oracle import oracle
    raw = synth_multicam(T=3000, seed=9)
    res = _run(raw, torch.float64, spans=[(250, 2900)])
    assert multicam_smooth_sessions.last_opt['launches'] == 5, 'the lag-statistics path did not run'
    ref = oracle.multicam(raw, dtype=np.float64, s_frames=[(250, 2900)])
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters'])
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    _check(_cam_out(res), ref['cam_out'], RTOL64, 'one offset span')


def test_fit_eks_multicam_ibl_paw(tmp_path):
    """fit_eks_multicam_ibl_paw (reference eks/ibl_paw_multicam_smoother.py:82-256) end to end on synthetic IBL-style
    files: left / right seed CSVs + timestamp files, right camera on its own clock; the call must equal the multi-camera
    smoother run on the aligned MarkerArray the driver is documented to build."""
    import pandas as pd
    from eks_b200.ibl_paw_multicam_smoother import fit_eks_multicam_ibl_paw
    from eks_b200.marker_array import MarkerArray
    from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
    from eks_b200.utils import make_dlc_pandas_index
    rng = np.random.default_rng(5)
    T, M, width = 700, 2, 128
    t_left = np.arange(T) / 60.0
    t_right = np.arange(T + 40) / 60.0 * 0.97 - 0.05                        # different rate and offset
    lat = np.cumsum(rng.normal(0, 0.4, (T + 80, 2, 2)), axis=0) + 60.0       # (frames, paw, xy) in the left camera's frame
    at = lambda ts: np.stack([np.stack([np.interp(ts, np.arange(T + 80) / 60.0 - 0.1, lat[:, p, c]) for c in range(2)], 1)
                              for p in range(2)], 1)                         # (len(ts), paw, xy)
    idx = make_dlc_pandas_index(['paw_l', 'paw_r'])
    for m in range(M):
        xl = at(t_left) + rng.normal(0, 0.3, (T, 2, 2))
        left = np.concatenate([np.concatenate([xl[:, p], np.full((T, 1), 0.95)], 1) for p in range(2)], 1)
        xr = at(t_right) + rng.normal(0, 0.3, (T + 40, 2, 2))
        xr[..., 0] = width - xr[..., 0]                                      # the right camera is mirrored ...
        right = np.concatenate([np.concatenate([xr[:, p], np.full((T + 40, 1), 0.95)], 1) for p in (1, 0)], 1)   # ... and swaps paws
        pd.DataFrame(left, columns=idx).to_csv(tmp_path / f'sess.left.rng={m}.csv')
        pd.DataFrame(right, columns=idx).to_csv(tmp_path / f'sess.right.rng={m}.csv')
    np.save(tmp_path / 'sess.timestamps.left.npy', t_left)
    np.save(tmp_path / 'sess.timestamps.right.npy', t_right)
    out_dir = tmp_path / 'out'
    dfs, s, input_dfs_list, bps = fit_eks_multicam_ibl_paw(str(tmp_path), str(out_dir), smooth_param=5.0, var_mode='var')
    assert bps == ['paw_l', 'paw_r'] and (out_dir / 'multicam_left_results.csv').exists() and list(s) == [5.0, 5.0]
    n = len(input_dfs_list[0][0])
    assert 0 < n <= T and dfs[0].shape == (n, 2 * 9)
    # after alignment both cameras see the same trajectory: the smoothed x of the two cameras agree to the noise level
    assert np.abs(dfs[0].to_numpy()[:, 0] - dfs[1].to_numpy()[:, 0]).mean() < 1.0
    arr = np.zeros((M, 2, n, 2, 3))
    for c in range(2):
        for m in range(M):
            arr[m, c, :, :, :2] = input_dfs_list[c][m].to_numpy().reshape(n, 2, 2)
    ref, _, _ = ensemble_kalman_smoother_multicam(MarkerArray(arr, data_fields=['x', 'y', 'likelihood']),
                                                  ['paw_l', 'paw_r'], ['left', 'right'], smooth_param=5.0, var_mode='var',
                                                  inflate_vars_kwargs={'likelihoods': None})
    for a, b in zip(dfs, ref):
        np.testing.assert_allclose(a.to_numpy(), b.to_numpy(), rtol=1e-6, atol=1e-6)
