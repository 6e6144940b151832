"""CPU tests of the multi-camera pre-stage restatements: the oracle's centring / PCA init / variance inflation against
the product's host mirrors of the reference functions, and the host halves of the device path (PCA and
FactorAnalysis on sufficient statistics) against scikit-learn itself."""
import numpy as np
import pytest

from oracle import oracle


def synth(V=2, K=3, T=1500, M=5, seed=0, outlier_frac=0.0):
    rng = np.random.default_rng(seed)
    lat = np.cumsum(rng.normal(0, 0.3, (T, K, 3)), axis=0)
    W = rng.standard_normal((K, 2 * V, 3))
    truth = np.einsum('tkl,kol->tko', lat, W) + rng.uniform(50, 300, (1, K, 2 * V))
    if outlier_frac > 0:
        out = rng.random((T, K)) < outlier_frac
        truth[:, :, 0] += np.where(out, rng.uniform(8, 25, (T, K)), 0.0)
    occ = rng.random((T, K)) < 0.05
    sigma = np.where(occ, 4.0, 0.5)[:, :, None]
    raw = np.empty((M, V, T, K, 3))
    for m in range(M):
        noisy = truth + rng.standard_normal((T, K, 2 * V)) * sigma
        raw[m, :, :, :, :2] = noisy.reshape(T, K, V, 2).transpose(2, 0, 1, 3)
        raw[m, :, :, :, 2] = rng.uniform(0.8, 1.0, (V, T, K))
    return raw


@pytest.mark.parametrize('q', [50.0, 95.0, 100.0])
def test_oracle_centering_equals_host_mirror(q):
    from eks_b200.marker_array import MarkerArray, mA_to_stacked_array
    from eks_b200.utils import center_predictions
    raw = synth(seed=int(q))
    ens = oracle.ensemble(raw, dtype=np.float64)
    ema = MarkerArray(ens[None], data_fields=['x', 'y', 'var_x', 'var_y', 'likelihood'], dtype=np.float64)
    mask, cen, good, means = center_predictions(ema, q)
    mask_o, cen_o, good_o, means_o, n_used = oracle.mc_center_predictions(ens, q)
    np.testing.assert_array_equal(mask, mask_o)
    for k in range(raw.shape[3]):
        np.testing.assert_allclose(mA_to_stacked_array(cen, k), cen_o[k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(mA_to_stacked_array(good, k), good_o[k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(means.array[0, :, 0, k, :].reshape(-1), means_o[k], rtol=1e-14)


def test_oracle_pca_init_equals_host_mirror():
    from eks_b200.marker_array import MarkerArray
    from eks_b200.multicam_smoother import initialize_kalman_filter_pca
    from eks_b200.stats import compute_pca
    from eks_b200.utils import center_predictions
    raw = synth(V=3, seed=3)
    ens = oracle.ensemble(raw, dtype=np.float64)
    ema = MarkerArray(ens[None], data_fields=['x', 'y', 'var_x', 'var_y', 'likelihood'], dtype=np.float64)
    mask, cen, good, _ = center_predictions(ema, 50.0)
    pcas, good_pcs = compute_pca(mask, cen, good, n_components=3)
    ref = initialize_kalman_filter_pca(good_pcs, pcas, 3)
    mask_o, cen_o, good_o, _, _ = oracle.mc_center_predictions(ens, 50.0)
    got = oracle.mc_pca_init(mask_o, cen_o, good_o, 3)
    for a, b in zip(ref, got):
        np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('O', [4, 6, 8])
def test_pca_from_moments_equals_sklearn(O):
    from sklearn.decomposition import PCA
    from eks_b200.pipeline import pca_from_moments
    rng = np.random.default_rng(O)
    X = rng.standard_normal((5000, 3)) @ rng.standard_normal((3, O)) * 3 + rng.standard_normal((5000, O)) * 0.3 + 1e-3
    pca = PCA(n_components=3).fit(X)
    mom = np.concatenate([[X.shape[0]], X.sum(0), (X.T @ X).ravel()])[None]
    mean, comps = pca_from_moments(mom, O, 3)
    np.testing.assert_allclose(mean[0], pca.mean_, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(comps[0], pca.components_, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize('O', [4, 6, 8])
def test_fa_from_moments_equals_sklearn(O):
    from sklearn.decomposition import FactorAnalysis
    from eks_b200.pipeline import fa_from_moments
    rng = np.random.default_rng(10 + O)
    X = (rng.standard_normal((20000, 3)) @ (rng.standard_normal((3, O)) * 2) +
         rng.standard_normal((20000, O)) * rng.uniform(0.2, 1.0, O) + rng.uniform(-3, 3, O))
    fa = FactorAnalysis(n_components=3).fit(X)
    mom = np.concatenate([[X.shape[0]], X.sum(0), (X.T @ X).ravel()])
    mean, W = fa_from_moments(mom, O, 3)
    ref = fa.components_.T
    sgn = np.sign((W * ref).sum(axis=0))
    np.testing.assert_allclose(mean, fa.mean_, rtol=1e-12)
    np.testing.assert_allclose(W * sgn, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())


def test_oracle_inflation_equals_host_mirror():
    from eks_b200.multicam_smoother import inflate_variance
    from eks_b200.stats import compute_mahalanobis
    raw = synth(T=1800, seed=5, outlier_frac=0.04)
    ens = oracle.ensemble(raw, dtype=np.float64)
    _, cen, _, _, _ = oracle.mc_center_predictions(ens, 50.0)
    ev = np.transpose(ens[..., 2:4], (2, 1, 0, 3)).reshape(3, 1800, 4)
    for k in range(3):
        v_o, rounds = oracle.mc_inflate_variance(cen[k], ev[k])
        tmp, infl, r2 = ev[k].copy(), True, 0
        while infl:
            res = compute_mahalanobis(cen[k], tmp, n_latent=3, likelihood_threshold=0.9, v_quantile_threshold=50.0)
            tmp, infl = inflate_variance(tmp, res['mahalanobis'], 5.0, 10.0)
            r2 += 1
        assert rounds == r2 and rounds >= 2
        np.testing.assert_allclose(v_o, tmp, rtol=1e-13)
        assert (v_o != ev[k]).sum() > 20
