// =====================================================================================
// eks_oracle.cpp -- CPU ORACLE for the EKS smoothing hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a from-scratch CPU restatement of the reference algorithm.  It is NOT part of
// the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it.  The product path (eks_b200/) never imports or links it.
//
// PARITY STATUS: "parity unpinned" at the dynamax boundary.  The Kalman arithmetic of the
// reference lives in the third-party package dynamax (pinned <=1.0.1 in
// /root/reference/pyproject.toml:39; with jax<=0.4.36, optax unpinned) which is neither vendored
// in /root/reference nor installable here (no network, no wheel).  The recursion below restates
// dynamax.nonlinear_gaussian_ssm.inference_ekf (extended_kalman_filter /
// extended_kalman_smoother, psd_solve with diagonal_boost=1e-9, symmetrize) and optax.adam's
// published algorithm, anchored on the reference's call sites:
//   eks/core.py:25-101   ensemble()                       -> oracle_ensemble
//   eks/core.py:136-155  params_nlgssm_for_keypoint       -> Model (Q scaled by s)
//   eks/core.py:640-650  loss = -marginal_loglik, non-finite -> 1e12   -> nll_and_grad
//   eks/core.py:654-681  adam(1.0) on lr-scaled grads, tol stop, safety cap -> oracle_optimize
//   eks/core.py:445-476  block loss = sum over members    -> oracle_optimize (nmem > 1)
//   eks/core.py:274-295  final extended_kalman_smoother   -> oracle_smooth
//   eks/multicam_smoother.py:806-859  pinhole projection  -> project_cam
// What IS pinned offline: the projection against cv2.projectPoints / cv2.Rodrigues, the gradient
// against torch reverse-mode autograd and finite differences, and the dense recursion against an
// independent NumPy restatement (tests/test_oracle.py).
//
// Design: every routine is a template over a scalar type S.  S = float/double evaluates values;
// S = Dual<real> carries d/ds by forward-mode AD (exactly what jax.value_and_grad computes, by a
// different but mathematically identical mode), and the projection Jacobian H(x) = dh/dx is
// obtained by a second, nested forward pass (mirrors jacfwd inside grad).
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <vector>

namespace {

constexpr int DMAX = 6;
constexpr int OMAX = 16;
constexpr int CAM_STRIDE = 29;  // R(9) t(3) fx fy cx cy skew k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4

// ---------------------------------------------------------------- forward-mode dual number
template <class T>
struct Dual {
    T v, d;
    Dual() : v(0), d(0) {}
    Dual(T v_) : v(v_), d(0) {}
    Dual(T v_, T d_) : v(v_), d(d_) {}
};
template <class T> inline Dual<T> operator+(Dual<T> a, Dual<T> b) { return {a.v + b.v, a.d + b.d}; }
template <class T> inline Dual<T> operator-(Dual<T> a, Dual<T> b) { return {a.v - b.v, a.d - b.d}; }
template <class T> inline Dual<T> operator-(Dual<T> a) { return {-a.v, -a.d}; }
template <class T> inline Dual<T> operator*(Dual<T> a, Dual<T> b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <class T> inline Dual<T> operator/(Dual<T> a, Dual<T> b) {
    T q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
template <class T> inline Dual<T>& operator+=(Dual<T>& a, Dual<T> b) { a = a + b; return a; }
template <class T> inline Dual<T>& operator-=(Dual<T>& a, Dual<T> b) { a = a - b; return a; }

inline float sqrt_(float x) { return std::sqrt(x); }
inline double sqrt_(double x) { return std::sqrt(x); }
inline float log_(float x) { return std::log(x); }
inline double log_(double x) { return std::log(x); }
template <class T> inline Dual<T> sqrt_(Dual<T> a) { auto r = sqrt_(a.v); return {r, a.d / (r + r)}; }
template <class T> inline Dual<T> log_(Dual<T> a) { return {log_(a.v), a.d / a.v}; }

template <class S> struct ScalarTraits { using real = S; static real val(S x) { return x; } };
template <class T> struct ScalarTraits<Dual<T>> {
    using real = typename ScalarTraits<T>::real;
    static real val(Dual<T> x) { return ScalarTraits<T>::val(x.v); }
};
template <class S> inline typename ScalarTraits<S>::real val(S x) { return ScalarTraits<S>::val(x); }

// ---------------------------------------------------------------- pinhole projection
// Restates make_jax_projection_fn.project (eks/multicam_smoother.py:824-857): polynomial radial
// term with k1..k6 (NOT OpenCV's rational model), tangential p1,p2, thin-prism s1..s4, skew.
template <class S, class R>
inline void project_cam(const R* cam, const S X[3], S uv[2]) {
    S Xc[3];
    for (int i = 0; i < 3; ++i)
        Xc[i] = S(cam[3 * i]) * X[0] + S(cam[3 * i + 1]) * X[1] + S(cam[3 * i + 2]) * X[2] + S(cam[9 + i]);
    const R fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15], skew = cam[16];
    const R k1 = cam[17], k2 = cam[18], p1 = cam[19], p2 = cam[20], k3 = cam[21], k4 = cam[22],
            k5 = cam[23], k6 = cam[24], s1 = cam[25], s2 = cam[26], s3 = cam[27], s4 = cam[28];
    S x = Xc[0] / Xc[2];
    S y = Xc[1] / Xc[2];
    S r2 = x * x + y * y;
    S r4 = r2 * r2;
    S r6 = r4 * r2;
    S r8 = r4 * r4;
    S r10 = r8 * r2;
    S r12 = r6 * r6;
    S radial = S(R(1)) + S(k1) * r2 + S(k2) * r4 + S(k3) * r6 + S(k4) * r8 + S(k5) * r10 + S(k6) * r12;
    S x_tan = S(R(2) * p1) * x * y + S(p2) * (r2 + S(R(2)) * x * x);
    S y_tan = S(p1) * (r2 + S(R(2)) * y * y) + S(R(2) * p2) * x * y;
    S x_tp = S(s1) * r2 + S(s2) * r4;
    S y_tp = S(s3) * r2 + S(s4) * r4;
    S xd = x * radial + x_tan + x_tp;
    S yd = y * radial + y_tan + y_tp;
    uv[0] = S(fx) * xd + S(skew) * yd + S(cx);
    uv[1] = S(fy) * yd + S(cy);
}

// ---------------------------------------------------------------- model
template <class R>
struct Model {
    int D, O, ncam;
    const R *m0, *S0, *A, *Q, *C, *cams;
};

// emission h(x) and its Jacobian H(x) (O x D, row-major).  Linear: C x.  Pinhole: concatenation
// over cameras (eks/multicam_smoother.py:877-882); Jacobian by forward-mode duals over the 3
// state directions (mirrors jax.jacfwd used by dynamax).
template <class S, class R>
inline void emission(const Model<R>& mdl, const S* x, S* yhat, S* H) {
    const int D = mdl.D, O = mdl.O;
    if (mdl.ncam == 0) {
        for (int i = 0; i < O; ++i) {
            S acc = S(R(0));
            for (int j = 0; j < D; ++j) {
                H[i * D + j] = S(mdl.C[i * D + j]);
                acc += H[i * D + j] * x[j];
            }
            yhat[i] = acc;
        }
        return;
    }
    for (int c = 0; c < mdl.ncam; ++c) {
        const R* cam = mdl.cams + c * CAM_STRIDE;
        for (int j = 0; j < 3; ++j) {
            Dual<S> Xd[3], uv[2];
            for (int i = 0; i < 3; ++i) Xd[i] = Dual<S>(x[i], i == j ? S(R(1)) : S(R(0)));
            project_cam<Dual<S>, R>(cam, Xd, uv);
            H[(2 * c) * D + j] = uv[0].d;
            H[(2 * c + 1) * D + j] = uv[1].d;
            if (j == 0) { yhat[2 * c] = uv[0].v; yhat[2 * c + 1] = uv[1].v; }
        }
    }
}

// ---------------------------------------------------------------- small dense helpers
// in-place lower Cholesky of n x n (row-major, leading dim n); returns false if not PD / NaN.
template <class S>
inline bool cholesky(S* a, int n) {
    for (int j = 0; j < n; ++j) {
        S d = a[j * n + j];
        for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
        if (!(val(d) > 0)) return false;
        d = sqrt_(d);
        a[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            S s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = s / d;
        }
    }
    return true;
}
// solve L L^T x = b for one right-hand side (L lower from cholesky)
template <class S>
inline void chol_solve(const S* L, int n, const S* b, S* x) {
    S z[OMAX];
    for (int i = 0; i < n; ++i) {
        S s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * z[k];
        z[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        S s = z[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
}

template <class S>
struct FilterState {
    S m[DMAX];
    S P[DMAX * DMAX];
};

// One EKF step at frame t (dynamax inference_ekf._step restated):
//   ll += log N(y; h(m), H P H^T + R);  K = psd_solve(S, H P)^T (boost 1e-9);
//   P_f = sym(P - K S K^T);  m_f = m + K (y - h(m));  then predict with A, s*Q.
// Returns false when the innovation covariance is not positive definite / not finite.
template <class S, class R>
inline bool ekf_step(const Model<R>& mdl, const R* y, const R* Rdiag, S s, FilterState<S>& st, S& ll,
                     S* mf_out, S* Pf_out, const S* A_s = nullptr, const S* Q_s = nullptr) {
    const int D = mdl.D, O = mdl.O;
    S yhat[OMAX], H[OMAX * DMAX], HP[OMAX * DMAX], Sm[OMAX * OMAX], L[OMAX * OMAX];
    emission<S, R>(mdl, st.m, yhat, H);
    for (int i = 0; i < O; ++i)
        for (int j = 0; j < D; ++j) {
            S acc = S(R(0));
            for (int k = 0; k < D; ++k) acc += H[i * D + k] * st.P[k * D + j];
            HP[i * D + j] = acc;
        }
    for (int i = 0; i < O; ++i)
        for (int j = 0; j < O; ++j) {
            S acc = S(R(0));
            for (int k = 0; k < D; ++k) acc += HP[i * D + k] * H[j * D + k];
            if (i == j) acc += S(Rdiag[i]);
            Sm[i * O + j] = acc;
        }
    // log-likelihood via Cholesky of the un-boosted S (tfp MVN full covariance)
    for (int i = 0; i < O * O; ++i) L[i] = Sm[i];
    if (!cholesky(L, O)) return false;
    S r[OMAX], z[OMAX];
    for (int i = 0; i < O; ++i) r[i] = S(y[i]) - yhat[i];
    S logdet = S(R(0)), quad = S(R(0));
    for (int i = 0; i < O; ++i) {
        S sacc = r[i];
        for (int k = 0; k < i; ++k) sacc -= L[i * O + k] * z[k];
        z[i] = sacc / L[i * O + i];
        quad += z[i] * z[i];
        logdet += log_(L[i * O + i]);
    }
    const R LOG2PI = R(1.8378770664093454835606594728112);
    ll += S(R(-0.5)) * (S(R(O) * LOG2PI) + S(R(2)) * logdet + quad);
    // gain: psd_solve(S, HP) = chol(sym(S) + 1e-9 I) \ HP
    S Lb[OMAX * OMAX];
    for (int i = 0; i < O; ++i)
        for (int j = 0; j < O; ++j) {
            Lb[i * O + j] = S(R(0.5)) * (Sm[i * O + j] + Sm[j * O + i]);
            if (i == j) Lb[i * O + j] += S(R(1e-9));
        }
    if (!cholesky(Lb, O)) return false;
    S K[DMAX * OMAX];  // D x O
    for (int j = 0; j < D; ++j) {
        S b[OMAX], x[OMAX];
        for (int i = 0; i < O; ++i) b[i] = HP[i * D + j];
        chol_solve(Lb, O, b, x);
        for (int i = 0; i < O; ++i) K[j * O + i] = x[i];
    }
    // P_f = P - K S K^T, symmetrised ; m_f = m + K r
    S KS[DMAX * OMAX];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < O; ++j) {
            S acc = S(R(0));
            for (int k = 0; k < O; ++k) acc += K[i * O + k] * Sm[k * O + j];
            KS[i * O + j] = acc;
        }
    S Pf[DMAX * DMAX], mf[DMAX];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            S acc = S(R(0));
            for (int k = 0; k < O; ++k) acc += KS[i * O + k] * K[j * O + k];
            Pf[i * D + j] = st.P[i * D + j] - acc;
        }
    for (int i = 0; i < D; ++i)
        for (int j = i + 1; j < D; ++j) {
            S a = S(R(0.5)) * (Pf[i * D + j] + Pf[j * D + i]);
            Pf[i * D + j] = a;
            Pf[j * D + i] = a;
        }
    for (int i = 0; i < D; ++i) {
        S acc = st.m[i];
        for (int k = 0; k < O; ++k) acc += K[i * O + k] * r[k];
        mf[i] = acc;
    }
    if (mf_out) for (int i = 0; i < D; ++i) mf_out[i] = mf[i];
    if (Pf_out) for (int i = 0; i < D * D; ++i) Pf_out[i] = Pf[i];
    // predict: m = A m_f ; P = A P_f A^T + s Q
    S AP[DMAX * DMAX];
    for (int i = 0; i < D; ++i) {
        S acc = S(R(0));
        for (int k = 0; k < D; ++k) acc += (A_s ? A_s[i * D + k] : S(mdl.A[i * D + k])) * mf[k];
        st.m[i] = acc;
        for (int j = 0; j < D; ++j) {
            S a2 = S(R(0));
            for (int k = 0; k < D; ++k) a2 += (A_s ? A_s[i * D + k] : S(mdl.A[i * D + k])) * Pf[k * D + j];
            AP[i * D + j] = a2;
        }
    }
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            S acc = S(R(0));
            for (int k = 0; k < D; ++k) acc += AP[i * D + k] * (A_s ? A_s[j * D + k] : S(mdl.A[j * D + k]));
            st.P[i * D + j] = acc + (Q_s ? Q_s[i * D + j] : s * S(mdl.Q[i * D + j]));
        }
    return true;
}

// Filter negative log-likelihood and d/ds via forward sensitivities (Dual numbers).
// R_tv != 0: Rdiag is (T, O) time-varying; else (O,) constant.
template <class R>
inline void nll_and_grad(const Model<R>& mdl, const R* y, const R* Rdiag, int R_tv, int T, R s, R* nll,
                         R* dnll_ds) {
    using S = Dual<R>;
    FilterState<S> st;
    for (int i = 0; i < mdl.D; ++i) st.m[i] = S(mdl.m0[i]);
    for (int i = 0; i < mdl.D * mdl.D; ++i) st.P[i] = S(mdl.S0[i]);
    S sd(s, R(1));
    S ll = S(R(0));
    bool ok = true;
    for (int t = 0; t < T && ok; ++t)
        ok = ekf_step<S, R>(mdl, y + (size_t)t * mdl.O, Rdiag + (R_tv ? (size_t)t * mdl.O : 0), sd, st, ll,
                            nullptr, nullptr);
    R v = -ll.v, g = -ll.d;
    if (!ok || !std::isfinite(v)) {  // core.py:650 -- non-finite NLL -> 1e12 (constant => zero grad)
        v = R(1e12);
        g = R(0);
    }
    *nll = v;
    *dnll_ds = g;
}

// Adam on log s restated from core.py:654-681 with optax.adam(1.0) defaults (b1=.9, b2=.999,
// eps=1e-8, eps_root=0, bias-corrected).  All optimiser scalars live in the working precision R
// (float32 in production: core.py:441,622).
template <class R>
inline void optimize(int nmem, const Model<R>* mdls, const R* const* ys, const R* const* Rcs, int T, R s_log0,
                     R lr, R lo, R hi, R tol, int cap, R* s_log_out, R* last_loss_out, int* iters_out,
                     R* trace, int trace_cap) {
    R s_log = s_log0, mu = 0, nu = 0, prev = std::numeric_limits<R>::infinity();
    const R b1 = R(0.9), b2 = R(0.999), eps = R(1e-8);
    int iters = 0;
    bool done = false;
    while (!done && iters < cap) {
        R sc = std::min(std::max(s_log, lo), hi);
        R s = std::exp(sc);
        R inside = (s_log >= lo && s_log <= hi) ? R(1) : R(0);
        R loss = 0, g = 0;
        for (int i = 0; i < nmem; ++i) {
            R v, dv;
            nll_and_grad<R>(mdls[i], ys[i], Rcs[i], 0, T, s, &v, &dv);
            loss += v;
            g += dv * s * inside;  // d/d log s through exp(clip(.))
        }
        g *= lr;
        int count = iters + 1;
        mu = b1 * mu + (R(1) - b1) * g;
        nu = b2 * nu + (R(1) - b2) * g * g;
        R mu_hat = mu / (R(1) - std::pow(b1, R(count)));
        R nu_hat = nu / (R(1) - std::pow(b2, R(count)));
        R upd = -mu_hat / (std::sqrt(nu_hat) + eps);
        if (trace && iters < trace_cap) {
            trace[3 * iters + 0] = s_log;
            trace[3 * iters + 1] = loss;
            trace[3 * iters + 2] = g;
        }
        s_log = s_log + upd;
        R rel_tol = tol * std::fabs(std::log(std::max(prev, R(1e-12))));
        bool stop = std::isfinite(prev) ? (std::fabs(loss - prev) < rel_tol + R(1e-6)) : false;
        prev = loss;
        iters += 1;
        done = stop;
    }
    *s_log_out = s_log;
    *last_loss_out = prev;
    *iters_out = iters;
}

// EKF filter + RTS smoother with time-varying diagonal R (extended_kalman_smoother restated).
template <class R>
inline int smooth(const Model<R>& mdl, const R* y, const R* Rdiag, int R_tv, int T, R s, R* ms, R* Vs, R* mfs,
                  R* Pfs, R* ll_out) {
    const int D = mdl.D;
    std::vector<R> mf((size_t)T * D), Pf((size_t)T * D * D);
    FilterState<R> st;
    for (int i = 0; i < D; ++i) st.m[i] = mdl.m0[i];
    for (int i = 0; i < D * D; ++i) st.P[i] = mdl.S0[i];
    R ll = 0;
    int bad = 0;
    for (int t = 0; t < T; ++t) {
        bool ok = ekf_step<R, R>(mdl, y + (size_t)t * mdl.O, Rdiag + (R_tv ? (size_t)t * mdl.O : 0), s, st, ll,
                                 &mf[(size_t)t * D], &Pf[(size_t)t * D * D]);
        if (!ok) {
            bad = 1;
            R nan = std::numeric_limits<R>::quiet_NaN();
            for (size_t i = (size_t)t * D; i < (size_t)T * D; ++i) mf[i] = nan;
            for (size_t i = (size_t)t * D * D; i < (size_t)T * D * D; ++i) Pf[i] = nan;
            break;
        }
    }
    if (ll_out) *ll_out = ll;
    if (mfs) std::memcpy(mfs, mf.data(), sizeof(R) * mf.size());
    if (Pfs) std::memcpy(Pfs, Pf.data(), sizeof(R) * Pf.size());
    // backward pass
    std::memcpy(ms + (size_t)(T - 1) * D, &mf[(size_t)(T - 1) * D], sizeof(R) * D);
    std::memcpy(Vs + (size_t)(T - 1) * D * D, &Pf[(size_t)(T - 1) * D * D], sizeof(R) * D * D);
    for (int t = T - 2; t >= 0; --t) {
        const R* mft = &mf[(size_t)t * D];
        const R* Pft = &Pf[(size_t)t * D * D];
        R mp[DMAX], AP[DMAX * DMAX], Sp[DMAX * DMAX], Lb[DMAX * DMAX], G[DMAX * DMAX];
        for (int i = 0; i < D; ++i) {
            R acc = 0;
            for (int k = 0; k < D; ++k) acc += mdl.A[i * D + k] * mft[k];
            mp[i] = acc;
            for (int j = 0; j < D; ++j) {
                R a2 = 0;
                for (int k = 0; k < D; ++k) a2 += mdl.A[i * D + k] * Pft[k * D + j];
                AP[i * D + j] = a2;  // F P_f
            }
        }
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                R acc = 0;
                for (int k = 0; k < D; ++k) acc += AP[i * D + k] * mdl.A[j * D + k];
                Sp[i * D + j] = acc + s * mdl.Q[i * D + j];
            }
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                Lb[i * D + j] = R(0.5) * (Sp[i * D + j] + Sp[j * D + i]);
                if (i == j) Lb[i * D + j] += R(1e-9);
            }
        bool ok = cholesky(Lb, D);
        for (int j = 0; j < D; ++j) {  // G = psd_solve(S_p, F P_f)^T : columns of F P_f
            R b[OMAX], x[OMAX];
            for (int i = 0; i < D; ++i) b[i] = AP[i * D + j];
            if (ok) chol_solve(Lb, D, b, x);
            else for (int i = 0; i < D; ++i) x[i] = std::numeric_limits<R>::quiet_NaN();
            for (int i = 0; i < D; ++i) G[j * D + i] = x[i];
        }
        const R* msn = ms + (size_t)(t + 1) * D;
        const R* Vsn = Vs + (size_t)(t + 1) * D * D;
        R* mst = ms + (size_t)t * D;
        R* Vst = Vs + (size_t)t * D * D;
        R dV[DMAX * DMAX], GdV[DMAX * DMAX];
        for (int i = 0; i < D * D; ++i) dV[i] = Vsn[i] - Sp[i];
        for (int i = 0; i < D; ++i) {
            R acc = mft[i];
            for (int k = 0; k < D; ++k) acc += G[i * D + k] * (msn[k] - mp[k]);
            mst[i] = acc;
            for (int j = 0; j < D; ++j) {
                R a2 = 0;
                for (int k = 0; k < D; ++k) a2 += G[i * D + k] * dV[k * D + j];
                GdV[i * D + j] = a2;
            }
        }
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                R acc = 0;
                for (int k = 0; k < D; ++k) acc += GdV[i * D + k] * G[j * D + k];
                Vst[i * D + j] = Pft[i * D + j] + acc;
            }
    }
    return bad;
}

// ---------------------------------------------------------------- ensemble statistics
// Restates compute_stats (eks/core.py:58-85) per (camera, frame, keypoint).
template <class R>
inline void ensemble(const R* raw, int M, int V, int T, int K, int avg_median, int var_mode, R nan_repl,
                     R* out) {
    const R NaN = std::numeric_limits<R>::quiet_NaN();
    const R FMAX = std::numeric_limits<R>::max();
#pragma omp parallel for collapse(2) schedule(static)
    for (int v = 0; v < V; ++v)
        for (int t = 0; t < T; ++t) {
            R buf[64];
            for (int k = 0; k < K; ++k) {
                R res[5];
                R conf = 0;
                for (int m = 0; m < M; ++m)
                    conf += raw[((((size_t)m * V + v) * T + t) * K + k) * 3 + 2];
                R mean_conf = conf / R(M);
                for (int c = 0; c < 2; ++c) {
                    int n = 0;
                    R sum = 0;
                    for (int m = 0; m < M; ++m) {
                        R x = raw[((((size_t)m * V + v) * T + t) * K + k) * 3 + c];
                        if (!std::isnan(x)) { buf[n++] = x; sum += x; }
                    }
                    R avg, var;
                    if (n == 0) { avg = NaN; var = NaN; }
                    else {
                        R mean = sum / R(n);
                        R ss = 0;  // nanvar, ddof 0, accumulated in seed order
                        for (int i = 0; i < n; ++i) ss += (buf[i] - mean) * (buf[i] - mean);
                        var = ss / R(n);
                        if (avg_median) {
                            std::sort(buf, buf + n);
                            avg = (n & 1) ? buf[n / 2] : (buf[n / 2 - 1] * R(0.5) + buf[n / 2] * R(0.5));
                        } else avg = mean;
                    }
                    if (M == 1) var = R(1) / std::max(mean_conf, R(1e-5));
                    else if (var_mode == 1) var = var / mean_conf;
                    // jnp.nan_to_num(nan=nan_replacement): nan -> repl, +inf -> max, -inf -> -max
                    if (std::isnan(var)) var = nan_repl;
                    else if (std::isinf(var)) var = var > 0 ? FMAX : -FMAX;
                    res[c] = avg;
                    res[2 + c] = var;
                }
                res[4] = mean_conf;
                R* o = out + (((size_t)v * T + t) * K + k) * 5;
                for (int i = 0; i < 5; ++i) o[i] = res[i];
            }
        }
}


// ---------------------------------------------------------------- IBL pupil model
// Restates pupil_optimize_smooth / run_pupil_kalman_smoother (eks/ibl_pupil_smoother.py:363-607):
// 3 states [diameter, com_x, com_y], AR(1) dynamics A = diag(s_d, s_c, s_c),
// Q = diag(var_d (1 - s_d^2), var_x (1 - s_c^2), var_y (1 - s_c^2)), 8 observations through a fixed C,
// time-varying diagonal R_t in the loss, two parameters u -> s = sigmoid(u) (1 - 2e-3) + 1e-3,
// optax.adam(lr) on u with the relative-tolerance stop rule.
inline float exp_(float x) { return std::exp(x); }
inline double exp_(double x) { return std::exp(x); }
template <class T> inline Dual<T> exp_(Dual<T> a) { auto e = exp_(a.v); return {e, a.d * e}; }

template <class S, class R>
inline void pupil_AQ(const S u[2], const R var3[3], S* A_s, S* Q_s) {
    S sv[2];
    for (int i = 0; i < 2; ++i) {
        S sg = S(R(1)) / (S(R(1)) + exp_(-u[i]));                    // jax.nn.sigmoid
        sv[i] = sg * S(R(1) - R(2) * R(1e-3)) + S(R(1e-3));          // _to_stable_s
    }
    const S sd[3] = {sv[0], sv[1], sv[1]};
    for (int i = 0; i < 9; ++i) { A_s[i] = S(R(0)); Q_s[i] = S(R(0)); }
    for (int i = 0; i < 3; ++i) {
        A_s[i * 3 + i] = sd[i];
        Q_s[i * 3 + i] = S(var3[i]) * (S(R(1)) - sd[i] * sd[i]);
    }
}

template <class R>
inline void pupil_nll_and_grad(const Model<R>& mdl, const R var3[3], const R* y, const R* Rdiag, int T,
                               const R u[2], R* nll, R grad[2]) {
    using S = Dual<R>;
    for (int k = 0; k < 2; ++k) {   // one forward-mode pass per parameter
        S ud[2] = {S(u[0], k == 0 ? R(1) : R(0)), S(u[1], k == 1 ? R(1) : R(0))};
        S A_s[9], Q_s[9];
        pupil_AQ<S, R>(ud, var3, A_s, Q_s);
        FilterState<S> st;
        for (int i = 0; i < 3; ++i) st.m[i] = S(mdl.m0[i]);
        for (int i = 0; i < 9; ++i) st.P[i] = S(mdl.S0[i]);
        S ll = S(R(0));
        bool ok = true;
        for (int t = 0; t < T && ok; ++t)
            ok = ekf_step<S, R>(mdl, y + (size_t)t * mdl.O, Rdiag + (size_t)t * mdl.O, S(R(1)), st, ll, nullptr,
                                nullptr, A_s, Q_s);
        *nll = ok ? -ll.v : std::numeric_limits<R>::quiet_NaN();
        grad[k] = ok ? -ll.d : std::numeric_limits<R>::quiet_NaN();
    }
}

template <class R>
inline void pupil_optimize(const Model<R>& mdl, const R var3[3], const R* y, const R* Rdiag, int T, R lr, R tol,
                           int cap, R u_out[2], R* last_loss, int* iters_out, R* trace, int trace_cap) {
    const R b1 = R(0.9), b2 = R(0.999), eps = R(1e-8);
    const float s0[2] = {0.99f, 0.98f};
    R u[2], mu[2] = {0, 0}, nu[2] = {0, 0};
    for (int i = 0; i < 2; ++i) u[i] = R(std::log(s0[i] / (1.0f - s0[i])));  // float32 seed
    R prev = std::numeric_limits<R>::infinity();
    int iters = 0;
    bool done = false;
    while (!done && iters < cap) {
        R loss, g[2];
        pupil_nll_and_grad<R>(mdl, var3, y, Rdiag, T, u, &loss, g);
        if (trace && iters < trace_cap) { trace[3 * iters] = u[0]; trace[3 * iters + 1] = u[1]; trace[3 * iters + 2] = loss; }
        const int count = iters + 1;
        for (int i = 0; i < 2; ++i) {
            mu[i] = b1 * mu[i] + (R(1) - b1) * g[i];
            nu[i] = b2 * nu[i] + (R(1) - b2) * g[i] * g[i];
            const R mh = mu[i] / (R(1) - std::pow(b1, R(count)));
            const R nh = nu[i] / (R(1) - std::pow(b2, R(count)));
            u[i] = u[i] - lr * mh / (std::sqrt(nh) + eps);
        }
        const R rel_tol = tol * std::fabs(std::log(std::max(prev, R(1e-12))));
        const bool stop = std::isfinite(prev) ? (std::fabs(loss - prev) < rel_tol + R(1e-6)) : false;
        prev = loss;
        iters += 1;
        done = stop;
    }
    u_out[0] = u[0]; u_out[1] = u[1];
    *last_loss = prev;
    *iters_out = iters;
}

template <class R>
Model<R> make_model(int D, int O, const R* m0, const R* S0, const R* A, const R* Q, const R* C, int ncam,
                    const R* cams) {
    Model<R> m;
    m.D = D; m.O = O; m.ncam = ncam; m.m0 = m0; m.S0 = S0; m.A = A; m.Q = Q; m.C = C; m.cams = cams;
    return m;
}

}  // namespace

// =====================================================================================
// C ABI (loaded with ctypes by oracle/oracle.py)
// =====================================================================================
#define ORACLE_API(P, R)                                                                                      \
    extern "C" void eks_oracle_ensemble_##P(const R* raw, int M, int V, int T, int K, int avg_median,         \
                                            int var_mode, R nan_repl, R* out) {                               \
        ensemble<R>(raw, M, V, T, K, avg_median, var_mode, nan_repl, out);                                    \
    }                                                                                                         \
    extern "C" void eks_oracle_project_##P(int ncam, const R* cams, int N, const R* X, R* uv, R* J) {         \
        Model<R> mdl = make_model<R>(3, 2 * ncam, nullptr, nullptr, nullptr, nullptr, nullptr, ncam, cams);   \
        for (int n = 0; n < N; ++n) {                                                                         \
            R yh[OMAX], H[OMAX * DMAX];                                                                       \
            emission<R, R>(mdl, X + 3 * n, yh, H);                                                            \
            for (int i = 0; i < 2 * ncam; ++i) uv[(size_t)n * 2 * ncam + i] = yh[i];                          \
            if (J) for (int i = 0; i < 6 * ncam; ++i) J[(size_t)n * 6 * ncam + i] = H[i];                     \
        }                                                                                                     \
    }                                                                                                         \
    /* batch over K independent sequences: arrays stacked on the leading axis */                              \
    extern "C" void eks_oracle_nll_grad_##P(int K, int D, int O, const R* m0, const R* S0, const R* A,        \
                                            const R* Q, const R* C, int ncam, const R* cams, const R* y,      \
                                            const R* Rdiag, int R_tv, int T, const R* s, R* nll, R* dnll) {   \
        _Pragma("omp parallel for schedule(dynamic)") for (int k = 0; k < K; ++k) {                           \
            Model<R> mdl = make_model<R>(D, O, m0 + (size_t)k * D, S0 + (size_t)k * D * D,                    \
                                         A + (size_t)k * D * D, Q + (size_t)k * D * D,                        \
                                         C ? C + (size_t)k * O * D : nullptr, ncam, cams);                    \
            nll_and_grad<R>(mdl, y + (size_t)k * T * O, Rdiag + (size_t)k * (R_tv ? (size_t)T * O : O), R_tv, \
                            T, s[k], nll + k, dnll + k);                                                      \
        }                                                                                                     \
    }                                                                                                         \
    /* nblocks blocks; block b owns members [off[b], off[b+1]) of the member-stacked arrays */                \
    extern "C" void eks_oracle_optimize_##P(int nblocks, const int* off, int D, int O, const R* m0,           \
                                            const R* S0, const R* A, const R* Q, const R* C, int ncam,        \
                                            const R* cams, const R* y, const R* Rconst, int T,                \
                                            const R* s_log0, R lr, R lo, R hi, R tol, int cap, R* s_log,      \
                                            R* last_loss, int* iters, R* trace, int trace_cap) {              \
        _Pragma("omp parallel for schedule(dynamic)") for (int b = 0; b < nblocks; ++b) {                     \
            int n = off[b + 1] - off[b];                                                                      \
            std::vector<Model<R>> mdls(n);                                                                    \
            std::vector<const R*> ys(n), rs(n);                                                               \
            for (int i = 0; i < n; ++i) {                                                                     \
                size_t k = (size_t)off[b] + i;                                                                \
                mdls[i] = make_model<R>(D, O, m0 + k * D, S0 + k * D * D, A + k * D * D, Q + k * D * D,       \
                                        C ? C + k * O * D : nullptr, ncam, cams);                             \
                ys[i] = y + k * (size_t)T * O;                                                                \
                rs[i] = Rconst + k * O;                                                                       \
            }                                                                                                 \
            optimize<R>(n, mdls.data(), ys.data(), rs.data(), T, s_log0[b], lr, lo, hi, tol, cap, s_log + b,  \
                        last_loss + b, iters + b, trace ? trace + (size_t)b * 3 * trace_cap : nullptr,        \
                        trace_cap);                                                                           \
        }                                                                                                     \
    }                                                                                                         \
    extern "C" void eks_oracle_pupil_nll_grad_##P(const R* m0, const R* S0, const R* C, const R* var3, const R* y,  \
                                                  const R* Rdiag, int T, const R* u, R* nll, R* grad) {           \
        Model<R> mdl = make_model<R>(3, 8, m0, S0, nullptr, nullptr, C, 0, nullptr);                              \
        pupil_nll_and_grad<R>(mdl, var3, y, Rdiag, T, u, nll, grad);                                              \
    }                                                                                                             \
    extern "C" void eks_oracle_pupil_optimize_##P(const R* m0, const R* S0, const R* C, const R* var3, const R* y,  \
                                                  const R* Rdiag, int T, R lr, R tol, int cap, R* u_out,          \
                                                  R* last_loss, int* iters, R* trace, int trace_cap) {            \
        Model<R> mdl = make_model<R>(3, 8, m0, S0, nullptr, nullptr, C, 0, nullptr);                              \
        pupil_optimize<R>(mdl, var3, y, Rdiag, T, lr, tol, cap, u_out, last_loss, iters, trace, trace_cap);       \
    }                                                                                                             \
    extern "C" void eks_oracle_smooth_##P(int K, int D, int O, const R* m0, const R* S0, const R* A,          \
                                          const R* Q, const R* C, int ncam, const R* cams, const R* y,        \
                                          const R* Rdiag, int R_tv, int T, const R* s, R* ms, R* Vs, R* mfs,  \
                                          R* Pfs, R* ll, int* bad) {                                          \
        _Pragma("omp parallel for schedule(dynamic)") for (int k = 0; k < K; ++k) {                           \
            Model<R> mdl = make_model<R>(D, O, m0 + (size_t)k * D, S0 + (size_t)k * D * D,                    \
                                         A + (size_t)k * D * D, Q + (size_t)k * D * D,                        \
                                         C ? C + (size_t)k * O * D : nullptr, ncam, cams);                    \
            int b = smooth<R>(mdl, y + (size_t)k * T * O, Rdiag + (size_t)k * (R_tv ? (size_t)T * O : O),     \
                              R_tv, T, s[k], ms + (size_t)k * T * D, Vs + (size_t)k * T * D * D,              \
                              mfs ? mfs + (size_t)k * T * D : nullptr,                                        \
                              Pfs ? Pfs + (size_t)k * T * D * D : nullptr, ll ? ll + k : nullptr);            \
            if (bad) bad[k] = b;                                                                              \
        }                                                                                                     \
    }

ORACLE_API(f32, float)
ORACLE_API(f64, double)

extern "C" int eks_oracle_dmax() { return DMAX; }
extern "C" int eks_oracle_omax() { return OMAX; }
