"""PCA initialisation and Mahalanobis-distance variance inflation (host side; mirror of eks/stats.py).

Out of scope as kernels (one-off tiny SVDs / pre-stage); kept on the host with
scikit-learn, with the reference's per-frame Python loops replaced by batched NumPy linear algebra.
"""

from __future__ import annotations

import numpy as np
from sklearn.decomposition import PCA, FactorAnalysis

from eks_b200.marker_array import MarkerArray, mA_to_stacked_array


def compute_pca(valid_frames_mask, emA_centered_preds: MarkerArray, emA_good_centered_preds: MarkerArray,
                n_components: int = 3, pca_object: PCA | None = None) -> tuple[list, list]:
    """Per-keypoint PCA on the variance-filtered centred predictions (eks/stats.py:9-64)."""
    n_models, V, T, K, _ = emA_centered_preds.shape
    assert n_models == 1, 'MarkerArray should have n_models = 1 after ensembling.'
    ensemble_pca, good_pcs_list = [], []
    for k in range(K):
        good_idx = np.where(valid_frames_mask[:, k])[0]
        gsp = mA_to_stacked_array(emA_good_centered_preds, k)
        sp = mA_to_stacked_array(emA_centered_preds, k)
        pca_k = PCA(n_components=n_components).fit(gsp) if pca_object is None else pca_object
        pcs = pca_k.transform(sp)
        ensemble_pca.append(pca_k)
        good_pcs_list.append(pcs[good_idx])
    return ensemble_pca, good_pcs_list


def compute_mahalanobis(x, v, n_latent: int = 3, v_quantile_threshold: float | None = 50.0, likelihoods=None,
                        likelihood_threshold: float | None = 0.9, epsilon: float | None = 1e-6,
                        loading_matrix=None, mean=None) -> dict:
    """Factor-analysis posterior predictive Mahalanobis distance per view (eks/stats.py:67-157)."""
    x, v = np.asarray(x, dtype=float), np.asarray(v, dtype=float)
    if loading_matrix is None or mean is None:
        if likelihoods is not None and likelihood_threshold is not None:
            valid = np.min(likelihoods, axis=1) >= likelihood_threshold
        else:
            valid = np.ones(x.shape[0], dtype=bool)
        if v_quantile_threshold is not None:
            ev_max = v.max(axis=1)
            valid = valid & (ev_max < np.percentile(ev_max, v_quantile_threshold))
        fa = FactorAnalysis(n_components=n_latent).fit(x[valid])
        W, mu = fa.components_.T, fa.mean_
    else:
        W, mu = np.asarray(loading_matrix), np.asarray(mean)
    iv = 1.0 / (v + epsilon)                                    # (N, 2C)
    Binv = np.einsum('di,nd,dj->nij', W, iv, W)                 # W^T diag(1/v) W
    B = np.linalg.inv(Binv)                                     # (N, L, L)
    z = np.einsum('nij,dj,nd,nd->ni', B, W, iv, x - mu)
    xhat = z @ W.T + mu
    diff = x - xhat
    n_views = x.shape[1] // 2
    Q, M = {}, {}
    for c in range(n_views):
        Wc = W[2 * c:2 * c + 2]
        Qc = np.einsum('ai,nij,bj->nab', Wc, B, Wc)
        Qc[:, 0, 0] += v[:, 2 * c]
        Qc[:, 1, 1] += v[:, 2 * c + 1]
        d = diff[:, 2 * c:2 * c + 2]
        Q[c] = Qc
        M[c] = np.einsum('na,nab,nb->n', d, np.linalg.inv(Qc), d)[:, None]
    return {'mahalanobis': M, 'posterior_variance': Q, 'reconstructed': xhat}
