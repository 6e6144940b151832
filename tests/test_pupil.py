"""CPU tests of the IBL pupil path: the oracle's AR(1) model against an independent NumPy Kalman filter, its
forward-mode gradient against finite differences, the committed golden vectors, and the host geometry helpers
(mirrors reference tests/test_ibl_pupil_smoother.py:40-170)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle


def pupil_raw_from_golden():
    """(M,T,4,3) in the fixed order top, bottom, right, left from the singlecam ibl-pupil golden's raw."""
    g = load_golden('singlecam_ibl_pupil')
    kps = [str(k) for k in g['keypoints']]
    idx = [kps.index(k) for k in oracle.PUPIL_POINTS]
    return g['raw'][:, 0][:, :, idx, :].astype(np.float64)


def synth_pupil(T=400, M=4, seed=0):
    rng = np.random.default_rng(seed)
    diam = 20 + 3 * np.sin(np.arange(T) / 40.0) + np.cumsum(rng.normal(0, 0.05, T))
    com = np.cumsum(rng.normal(0, 0.2, (T, 2)), axis=0) + [60.0, 45.0]
    pts = np.stack([com + np.stack([0 * diam, -diam / 2], 1), com + np.stack([0 * diam, diam / 2], 1),
                    com + np.stack([diam / 2, 0 * diam], 1), com + np.stack([-diam / 2, 0 * diam], 1)], axis=1)
    pred = pts[None] + rng.normal(0, 0.4, (M, T, 4, 2))
    lik = rng.uniform(0.8, 1.0, (M, T, 4, 1))
    return np.concatenate([pred, lik], axis=-1).astype(np.float32).astype(np.float64)


def numpy_pupil_nll(ys, m0, S0, C, var3, Rdiag, u):
    """Independent restatement in the textbook (joint-update) form, float64 NumPy."""
    s = oracle.pupil_to_s(u)
    sd = np.array([s[0], s[1], s[1]])
    A, Q = np.diag(sd), np.diag(var3 * (1 - sd ** 2))
    m, P, ll = m0.copy(), S0.copy(), 0.0
    for t in range(ys.shape[0]):
        R = np.diag(Rdiag[t])
        S = C @ P @ C.T + R
        r = ys[t] - C @ m
        ll += -0.5 * (r @ np.linalg.solve(S, r) + np.linalg.slogdet(S)[1] + len(r) * np.log(2 * np.pi))
        K = np.linalg.solve(S, C @ P).T
        m = m + K @ r
        P = P - K @ S @ K.T
        m, P = A @ m, A @ P @ A.T + Q
    return -ll


def _model(T=300, seed=1):
    rng = np.random.default_rng(seed)
    raw = synth_pupil(T=T, seed=seed)
    ens = oracle.ensemble(raw[:, None], dtype=np.float64)[0]
    preds = ens[..., :2].reshape(T, 8)
    Rdiag = np.clip(ens[..., 2:4].reshape(T, 8), 1e-12, None)
    diam, loc = oracle.pupil_diameter(preds), oracle.pupil_location(preds)
    y = preds.copy()
    y[:, 0::2] -= loc[:, 0].mean()
    y[:, 1::2] -= loc[:, 1].mean()
    m0 = np.array([diam.mean(), 0, 0])
    var3 = np.array([diam.var(), loc[:, 0].var(), loc[:, 1].var()])
    return y, m0, np.diag(var3), var3, Rdiag, rng


def test_pupil_nll_matches_numpy_filter():
    y, m0, S0, var3, Rdiag, _ = _model()
    for u in ([4.6, 3.9], [0.3, -1.0], [2.0, 6.0]):
        nll, _ = oracle.pupil_nll_grad(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, u)
        ref = numpy_pupil_nll(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, np.array(u))
        assert abs(nll - ref) <= 1e-7 * abs(ref), (u, nll, ref)


def test_pupil_gradient_matches_finite_differences():
    y, m0, S0, var3, Rdiag, _ = _model()
    u = np.array([3.0, 2.0])
    _, g = oracle.pupil_nll_grad(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, u)
    for k in range(2):
        h = 1e-5
        up, um = u.copy(), u.copy()
        up[k] += h
        um[k] -= h
        fd = (oracle.pupil_nll_grad(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, up)[0] -
              oracle.pupil_nll_grad(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, um)[0]) / (2 * h)
        assert abs(g[k] - fd) <= 1e-5 * max(1.0, abs(fd)), (k, g[k], fd)


def test_pupil_optimizer_descends_and_stops():
    y, m0, S0, var3, Rdiag, _ = _model()
    r = oracle.pupil_optimize(y, m0, S0, oracle.PUPIL_C, var3, Rdiag, dtype=np.float64, trace_cap=5000)
    tr = r['trace'][:r['iters']]
    assert 2 <= r['iters'] <= 5000
    assert tr[-1, 2] < tr[0, 2]                     # loss went down
    np.testing.assert_allclose(tr[0, :2], np.log(np.float32([0.99, 0.98]) / (1 - np.float32([0.99, 0.98]))), rtol=1e-6)
    assert np.all((r['s'] > 1e-3) & (r['s'] < 1 - 1e-3))
    if r['iters'] < 5000:                           # stop rule, eks/ibl_pupil_smoother.py:588-596
        assert abs(tr[-1, 2] - tr[-2, 2]) < 1e-6 * abs(np.log(tr[-2, 2])) + 1e-6


@pytest.mark.parametrize('name,kw', [('ibl_pupil_fixed_s', dict(smooth_params=[0.9, 0.95])),
                                     ('ibl_pupil_sframes', dict(s_frames=[(100, 700), (1200, None)]))])
def test_pupil_golden_reproducible(name, kw):
    g = load_golden(name)
    r = oracle.ibl_pupil(pupil_raw_from_golden(), dtype=np.float64, **kw)
    np.testing.assert_allclose(r['out'], g['out_f64'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r['s_finals'], g['s_f64'], rtol=1e-12)
    if 'iters_f64' in g:
        assert r['info']['iters'] == int(g['iters_f64'])


def test_fixed_params_are_clipped_like_the_reference():
    r = oracle.ibl_pupil(synth_pupil(T=50), smooth_params=[0.0, 2.0], dtype=np.float64)
    np.testing.assert_allclose(r['s_finals'], [np.float32(1e-3), np.float32(1 - 1e-3)], rtol=1e-7)


# ---- host geometry helpers of the product (NumPy pre-stages; no GPU needed) ------------------------------
def _mock_dlc(n=10, seed=0):
    rng = np.random.default_rng(seed)
    d = {f'pupil_{p}_r_{c}': rng.random(n) for p in ['top', 'bottom', 'left', 'right'] for c in 'xy'}
    d['pupil_top_r_x'][2] = np.nan
    d['pupil_left_r_y'][5] = np.nan
    return d


def test_get_pupil_location_and_diameter():
    from eks_b200.ibl_pupil_smoother import get_pupil_diameter, get_pupil_location
    d = _mock_dlc()
    c = get_pupil_location(d)
    assert c.shape == (10, 2) and np.isfinite(c).all()
    diam = get_pupil_diameter(d)
    assert diam.shape == (10,) and np.isfinite(diam).all()
    assert np.isnan(get_pupil_diameter({k: np.full(10, np.nan) for k in d})).all()
    # same numbers as the oracle's independent restatement
    preds = np.stack([d[f'{p}_{c}'] for p in oracle.PUPIL_POINTS for c in 'xy'], axis=1)
    np.testing.assert_allclose(c, oracle.pupil_location(preds), rtol=1e-14)
    np.testing.assert_allclose(diam, oracle.pupil_diameter(preds), rtol=1e-14)


def test_add_mean_to_array():
    from eks_b200.ibl_pupil_smoother import add_mean_to_array
    arr = np.array([[1.0, 2.0, 3.0, 4.0]])
    out = add_mean_to_array(arr, ['key1_x', 'key2_y', 'key3_x', 'key4_y'], 2.0, 3.0)
    assert {k: float(v[0]) for k, v in out.items()} == {'key1_x': 3.0, 'key2_y': 5.0, 'key3_x': 5.0, 'key4_y': 7.0}
    assert add_mean_to_array(np.zeros((0, 0)), [], 2.0, 3.0) == {}
