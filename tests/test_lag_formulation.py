"""The closed form that eks_b200/csrc/lin_lag.cu evaluates, restated in NumPy and checked against the sequential
Kalman filter NLL (CPU test; the device kernel is checked against the run-parallel path and the oracle in
tests/test_gpu_multicam_pipeline.py).

Model: x_{t+1} = x_t + w (A = I), y_t = C x_t + v, R diagonal and constant.  Once the covariance recursion has
converged, with C = Co Rc, U = [Co | V] orthogonal, K the steady gain, N = S^-1:
    e1_{t+1} = Phi e1_t + Gam g_{t+1},   g_i = [Co^T (y_i - y_{i-1}); V^T y_{i-1}],   Phi = I - Rc K1,   Gam = [I, -Rc K2]
and sum_t e_t^T N e_t is a quadratic form in the lag statistics Rg_m = sum_i g_i g_{i+m}^T (terms A, B, C below)."""
import numpy as np
import pytest


def _sequential_nll(y, Cm, r, Q, S0, m0, s, n_frames):
    O = y.shape[1]
    P, mu, nll = S0.copy(), m0.copy(), 0.0
    for t in range(n_frames):
        for g in range(O):
            h = Cm[g]
            Ph = P @ h
            si = r[g] + h @ Ph
            k = Ph / si
            e = y[t, g] - h @ mu
            nll += 0.5 * (np.log(2 * np.pi * si) + e * e / si)
            mu = mu + k * e
            P = P - np.outer(k, Ph)
        P = 0.5 * (P + P.T) + s * Q
    return nll, mu, P


def _closed_form_nll(y, Cm, r, Q, S0, m0, s, T0, W):
    n, O = y.shape
    D = Cm.shape[1]
    head, mu, P = _sequential_nll(y, Cm, r, Q, S0, m0, s, T0)
    # steady state in the information form (the device code's cov_step)
    J = Cm.T @ np.diag(1 / r) @ Cm
    Pp = np.linalg.inv(np.linalg.inv(P) + J)
    logdetS = np.sum(np.log(r)) + np.log(np.linalg.det(P) * np.linalg.det(np.linalg.inv(P) + J))
    K = Pp @ Cm.T @ np.diag(1 / r)
    N = np.diag(1 / r) - np.diag(1 / r) @ Cm @ Pp @ Cm.T @ np.diag(1 / r)
    Co, Rc = np.linalg.qr(Cm)
    U = np.linalg.qr(np.concatenate([Co, np.eye(O)], axis=1))[0][:, :O]
    U[:, :D] = Co
    V = U[:, D:]
    Kh = K @ U
    Phi = np.eye(D) - Rc @ Kh[:, :D]
    B = Rc @ Kh[:, D:]
    Gam = np.concatenate([np.eye(D), -B], axis=1)
    Nh = U.T @ N @ U
    N11, N12, N22 = Nh[:D, :D], Nh[:D, D:], Nh[D:, D:]
    S2 = np.concatenate([np.zeros((O - D, D)), np.eye(O - D)], axis=1)
    g = np.zeros((n + 1, O))
    g[1:n, :D] = (y[1:] - y[:-1]) @ Co
    g[1:n + 1, D:] = y[0:n] @ V
    G = np.zeros((D, D))
    Pj = np.eye(D)
    for _ in range(4 * W):
        G += Pj.T @ N11 @ Pj
        Pj = Pj @ Phi
    e0 = (U.T @ (y[T0] - Cm @ mu))[:D]
    A, Bt, Ct = e0 @ G @ e0, 0.0, 0.0
    tl = np.zeros(D)
    Pm, Pm1 = np.eye(D), None
    for m in range(W):
        i = np.arange(T0 + 1, n - m + 1)
        Rg = g[i].T @ g[i + m]
        RA = Rg - np.outer(g[n - m], g[n])
        A += (1 if m == 0 else 2) * np.sum((Gam.T @ Pm.T @ G @ Gam) * RA)
        if m >= 1:
            A += 2 * e0 @ (Pm.T @ G @ Gam @ g[T0 + m])
            Bt += 2 * np.sum((Gam.T @ Pm1.T @ N12 @ S2) * Rg)
        Bt += 2 * e0 @ (Pm.T @ N12 @ (S2 @ g[T0 + m + 1]))
        tl += Pm @ Gam @ g[n - 1 - m]
        if m == 0:
            Ct = np.sum(N22 * Rg[D:, D:])
        Pm1, Pm = Pm, Pm @ Phi
    pt = Phi @ tl
    A -= pt @ G @ pt
    rho = np.abs(np.linalg.eigvals(Phi)).max()
    return head + (n - T0) * (O * 0.5 * np.log(2 * np.pi) + 0.5 * logdetS) + 0.5 * (A + Bt + Ct), rho


@pytest.mark.parametrize('O,s', [(4, 0.13), (4, 5.0), (6, 0.02), (8, 0.4)])
def test_lag_statistics_closed_form_equals_sequential_filter(O, s):
    rng = np.random.default_rng(O * 100 + int(s * 1000))
    D, n = 3, 5000
    Cm = np.linalg.qr(rng.standard_normal((O, D)))[0] @ np.array([[1.2, .1, 0], [0, .9, .2], [0, 0, 1.1]])
    Q = np.array([[1, .3, .1], [.3, .8, .05], [.1, .05, .6]])
    r = rng.uniform(0.03, 0.08, O)
    lat = np.cumsum(rng.normal(0, .3, (n, D)), axis=0)
    y = lat @ Cm.T * 1.5 + rng.normal(0, .2, (n, O))
    S0, m0 = np.diag([2., 1., .5]), np.zeros(D)
    full, _, _ = _sequential_nll(y, Cm, r, Q, S0, m0, s, n)
    closed, rho = _closed_form_nll(y, Cm, r, Q, S0, m0, s, T0=256, W=256)
    assert rho < 0.95                       # Phi is a contraction (I - C K is not: it has O - D unit eigenvalues)
    assert abs(full - closed) <= 1e-9 * abs(full), (full, closed)


@pytest.mark.parametrize('s,r', [(0.06, 0.03), (0.005, 0.05), (3.0, 0.02)])
def test_scalar_lag_closed_form_equals_sequential_filter(s, r):
    """The single-camera form of eks_b200/csrc/diag_lag.cu (A = C = 1): e_{t+1} = alpha e_t + d_{t+1} and
    sum e^2 = [e0^2 + 2 e0 H + F - alpha^2 Tl^2] / (1 - alpha^2) with F = R_0 + 2 sum_m alpha^m R_m."""
    rng = np.random.default_rng(int(s * 1e4))
    n, T0, W = 6000, 256, 256
    y = np.cumsum(rng.normal(0, 0.3, n)) + rng.normal(0, 0.4, n)
    P, m, nll = 1.7, 0.0, 0.0
    e_steady = None
    for t in range(n):
        if t == T0:
            P_T0, m_T0, head = P, m, nll
        S = P + r
        e = y[t] - m
        nll += 0.5 * (np.log(2 * np.pi * S) + e * e / S)
        K = P / S
        m = m + K * e
        P = P - K * S * K + s
    Pinf = P_T0                                    # converged long before T0 for these parameters
    S = Pinf + r
    alpha = r / S
    assert alpha ** W < 1e-13
    d = np.diff(y, prepend=0.0)
    e0 = y[T0] - m_T0
    i = np.arange(T0 + 1, n)
    F = H = Tl = 0.0
    for mlag in range(W):
        Rm = np.sum(d[T0 + 1:n - mlag] * d[T0 + 1 + mlag:n])
        F += (1 if mlag == 0 else 2) * alpha ** mlag * Rm
        if mlag >= 1:
            H += alpha ** mlag * d[T0 + mlag]
        Tl += alpha ** mlag * d[n - 1 - mlag]
    E2 = (e0 * e0 + 2 * e0 * H + F - alpha ** 2 * Tl ** 2) / (1 - alpha ** 2)
    closed = head + (n - T0) * 0.5 * np.log(2 * np.pi * S) + 0.5 * E2 / S
    assert abs(nll - closed) <= 1e-9 * abs(nll), (nll, closed)
