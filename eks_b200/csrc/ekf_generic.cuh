// ekf_generic.cuh -- generic-dimension EKF filter / NLL(+d/ds) / RTS smoother for ONE sequence,
// written as __host__ __device__ templates so the identical arithmetic can be exercised on the CPU
// (tests/hostcheck) and runs one sequence per thread on the GPU (generic.cu).
//
// Reference behaviour restated: dynamax extended_kalman_filter / extended_kalman_smoother as called
// from eks/core.py:290, 469, 500, 648.  Because the observation covariance R is
// always diagonal on this path (eks/utils.py:368-377, eks/core.py:702-709), the measurement update
// is carried out as O sequential scalar updates on the linearised model -- algebraically identical
// to the batch update K = P H^T S^-1, P_f = P - K S K^T, ll += log N(y; h(m), S), but needing only
// D-sized temporaries that live in registers.  (dynamax's 1e-9 diagonal boost inside the gain solve
// is not representable in this form; its effect is O(1e-9 / S) relative, see DESIGN.md.)
#pragma once
#include "common.cuh"

#ifndef EKS_HD
#define EKS_HD __host__ __device__ inline
#endif

namespace eks {

// ------------------------------------------------------------------ pinhole camera (Anipose model)
// reference: make_jax_projection_fn.project, eks/multicam_smoother.py:824-857.  Returns the pixel
// coordinates and the analytic 2x3 Jacobian d(u,v)/dX.  S may be a Dual so that d/ds of both the
// projection and its Jacobian (second derivatives of h) are carried automatically.
template <class S, class P>
EKS_HD void project_cam_jac(const P* __restrict__ cam, const S* X, S* uv, S* J /*[2][3]*/) {
    S Xc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        Xc[i] = S(cam[3 * i]) * X[0] + S(cam[3 * i + 1]) * X[1] + S(cam[3 * i + 2]) * X[2] + S(cam[9 + i]);
    const P fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15], skew = cam[16];
    const P k1 = cam[17], k2 = cam[18], p1 = cam[19], p2 = cam[20], k3 = cam[21], k4 = cam[22], k5 = cam[23],
            k6 = cam[24], s1 = cam[25], s2 = cam[26], s3 = cam[27], s4 = cam[28];
    const S iz = S(P(1)) / Xc[2];
    const S x = Xc[0] * iz, y = Xc[1] * iz;
    const S r2 = x * x + y * y;
    const S two_x = x + x, two_y = y + y;
    S dxd_dx, dxd_dy, dyd_dx, dyd_dy;
    // Anipose calibrates k1 only by default: every other coefficient is exactly zero.  Their terms then add exact
    // zeros (0 * finite = 0, a + 0 = a), so skipping them is bit-identical for finite inputs and removes ~45 % of the
    // projection's arithmetic (the test is uniform across the warp: the cameras are shared by all sequences).
    const bool k1_only = k2 == P(0) && k3 == P(0) && k4 == P(0) && k5 == P(0) && k6 == P(0) && p1 == P(0) &&
                         p2 == P(0) && s1 == P(0) && s2 == P(0) && s3 == P(0) && s4 == P(0);
    if (k1_only) {
        const S radial = S(P(1)) + S(k1) * r2;
        const S drad = S(k1);
        const S xd = x * radial, yd = y * radial;
        uv[0] = S(fx) * xd + S(skew) * yd + S(cx);
        uv[1] = S(fy) * yd + S(cy);
        dxd_dx = radial + x * drad * two_x;
        dxd_dy = x * drad * two_y;
        dyd_dx = y * drad * two_x;
        dyd_dy = radial + y * drad * two_y;
    } else {
        const S r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4, r10 = r8 * r2, r12 = r6 * r6;
        const S radial = S(P(1)) + S(k1) * r2 + S(k2) * r4 + S(k3) * r6 + S(k4) * r8 + S(k5) * r10 + S(k6) * r12;
        // d radial / d r2
        const S drad = S(k1) + S(P(2) * k2) * r2 + S(P(3) * k3) * r4 + S(P(4) * k4) * r6 + S(P(5) * k5) * r8 +
                       S(P(6) * k6) * r10;
        const S xd = x * radial + S(P(2) * p1) * x * y + S(p2) * (r2 + S(P(2)) * x * x) + S(s1) * r2 + S(s2) * r4;
        const S yd = y * radial + S(p1) * (r2 + S(P(2)) * y * y) + S(P(2) * p2) * x * y + S(s3) * r2 + S(s4) * r4;
        uv[0] = S(fx) * xd + S(skew) * yd + S(cx);
        uv[1] = S(fy) * yd + S(cy);
        const S tpx = S(s1) + S(P(2) * s2) * r2, tpy = S(s3) + S(P(2) * s4) * r2;  // d thin-prism / d r2
        // d(xd,yd)/d(x,y)
        dxd_dx = radial + x * drad * two_x + S(P(2) * p1) * y + S(P(6) * p2) * x + tpx * two_x;
        dxd_dy = x * drad * two_y + S(P(2) * p1) * x + S(P(2) * p2) * y + tpx * two_y;
        dyd_dx = y * drad * two_x + S(P(2) * p1) * x + S(P(2) * p2) * y + tpy * two_x;
        dyd_dy = radial + y * drad * two_y + S(P(6) * p1) * y + S(P(2) * p2) * x + tpy * two_y;
    }
    // d(u,v)/d(x,y)
    const S du_dx = S(fx) * dxd_dx + S(skew) * dyd_dx, du_dy = S(fx) * dxd_dy + S(skew) * dyd_dy;
    const S dv_dx = S(fy) * dyd_dx, dv_dy = S(fy) * dyd_dy;
    // d(x,y)/dXc = [[iz, 0, -x iz], [0, iz, -y iz]] ; dXc/dX = R
    const S gux[3] = {du_dx * iz, du_dy * iz, -(du_dx * x + du_dy * y) * iz};
    const S gvx[3] = {dv_dx * iz, dv_dy * iz, -(dv_dx * x + dv_dy * y) * iz};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        J[j] = gux[0] * S(cam[j]) + gux[1] * S(cam[3 + j]) + gux[2] * S(cam[6 + j]);
        J[3 + j] = gvx[0] * S(cam[j]) + gvx[1] * S(cam[3 + j]) + gvx[2] * S(cam[6 + j]);
    }
}

// ------------------------------------------------------------------ per-sequence model
template <class P>
struct SeqModel {
    int D, O, ncam;
    const P *m0, *S0, *A, *Q, *C, *cams;  // this sequence's arrays (C: O x D row-major; cams shared)
};

// observation access: y[o] = plane(o)[t] - mean[o], r[o] = max(var plane(o)[t], floor) or constant
template <class P>
struct SeqObs {
    const P* y_base;      // sequence base pointer
    const P* var_base;    // may be null when R is constant
    const long long* y_off;
    const long long* var_off;
    const P* ymean;       // [O] or null
    const P* Rconst;      // [O] or null (=> time-varying from var planes)
    P var_floor;
};

template <int DC, int OC, bool FIXED>
struct Dims {
    int d, o;
    EKS_HD int D() const { return FIXED ? DC : d; }
    EKS_HD int O() const { return FIXED ? OC : o; }
};

// One EKF step on the predicted state (m, Pm): sequential scalar updates over the O channels,
// accumulating nll += 0.5 * (log 2pi + log s_i + e_i^2 / s_i); then (optionally) the filtered
// moments are exported and the state is predicted forward with A, s*Q.
// Returns false if an innovation variance is not positive/finite.
//
// BOOST (final smoother pass): dynamax computes the gain with psd_solve's 1e-9 diagonal boost,
// K = P H^T (S + eps I)^-1, and the filtered covariance as P - K S K^T with the UN-boosted S.  That is
// represented exactly here: the sequential scalar updates run with R' = R + eps, which gives exactly that K (hence
// the filtered mean) and P' = P - K (S + eps I) K^T; then P_f = P' + eps K K^T with K = P' H^T R'^-1, i.e.
// P_f = P' + eps P' (H^T R'^-2 H) P'.  It matters on frames where an ensemble variance is tiny (eps / R up to 1e-6
// relative); the loss path (constant R >= 1e-4) keeps BOOST = false: its effect there is O(eps / S) on the NLL.
template <class S, class P, int DC, int OC, bool FIXED, bool NL, bool BOOST = false>
EKS_HD bool ekf_step(const Dims<DC, OC, FIXED>& dm, const SeqModel<P>& mdl, const P* yv, const P* rv, S s, S* m,
                     S* Pm, S& nll, S* mf_out, S* Pf_out, const S* Adiag = nullptr, const S* Qdiag = nullptr,
                     bool a_identity = false) {
    const int D = dm.D(), O = dm.O();
    S Mb[BOOST ? DC * DC : 1];   // H^T R'^-2 H
    if (BOOST) {
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) Mb[BOOST ? i : 0] = S(P(0));
    }
    const P HALF_LOG2PI = P(0.91893853320467274178032973640562);
    S delta[DC];  // m_cur - m_pred
#pragma unroll
    for (int i = 0; i < DC; ++i) delta[i] = S(P(0));
    S m_pred[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) m_pred[i] = m[i];
    bool ok = true;
    const int npair = NL ? mdl.ncam : O;
#pragma unroll
    for (int g = 0; g < (NL ? OC / 2 : OC); ++g) {
        if (g >= npair) break;
        S yhat[2], Hrow[2 * DC];
        const int nrow = NL ? 2 : 1;
        if (NL) {
            project_cam_jac<S, P>(mdl.cams + g * CAM_STRIDE, m_pred, yhat, Hrow);
        } else {
            S acc = S(P(0));
#pragma unroll
            for (int j = 0; j < DC; ++j)
                if (j < D) { Hrow[j] = S(mdl.C[g * D + j]); acc += Hrow[j] * m_pred[j]; }
            yhat[0] = acc;
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            if (rr >= nrow) break;
            const int ch = NL ? 2 * g + rr : g;
            const S* h = Hrow + rr * (NL ? 3 : 0);
            // innovation on the linearised model
            S e = S(yv[ch]) - yhat[rr];
            S Ph[DC];
#pragma unroll
            for (int i = 0; i < DC; ++i) {
                if (i < D) {
                    S acc = S(P(0));
#pragma unroll
                    for (int j = 0; j < DC; ++j)
                        if (j < D) acc += Pm[i * D + j] * h[j];
                    Ph[i] = acc;
                }
            }
            S si = S(BOOST ? rv[ch] + P(1e-9) : rv[ch]);
            if (BOOST) {
                const S ir2 = S(P(1)) / (si * si);   // si == R' here
#pragma unroll
                for (int i = 0; i < DC; ++i)
                    if (i < D) {
#pragma unroll
                        for (int j = 0; j < DC; ++j)
                            if (j < D) Mb[BOOST ? i * D + j : 0] += h[i] * h[j] * ir2;
                    }
            }
#pragma unroll
            for (int j = 0; j < DC; ++j)
                if (j < D) { si += h[j] * Ph[j]; e -= h[j] * delta[j]; }
            if (!(val(si) > 0) || !isfinite((double)val(si))) ok = false;
            const S isi = S(P(1)) / si;
            nll += S(HALF_LOG2PI) + S(P(0.5)) * (log_(si) + e * e * isi);
#pragma unroll
            for (int i = 0; i < DC; ++i) {
                if (i < D) {
                    const S k = Ph[i] * isi;
                    delta[i] += k * e;
#pragma unroll
                    for (int j = 0; j < DC; ++j)
                        if (j < D) Pm[i * D + j] -= k * Ph[j];
                }
            }
        }
    }
    if (BOOST) {   // P_f = P' + eps P' (H^T R'^-2 H) P'
        S PM[DC * DC];
#pragma unroll
        for (int i = 0; i < DC; ++i)
            if (i < D) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) {
                        S acc = S(P(0));
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) acc += Pm[i * D + k] * Mb[BOOST ? k * D + j : 0];
                        PM[i * D + j] = acc;
                    }
            }
        S add[DC * DC];
#pragma unroll
        for (int i = 0; i < DC; ++i)
            if (i < D) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) {
                        S acc = S(P(0));
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) acc += PM[i * D + k] * Pm[k * D + j];
                        add[i * D + j] = acc;
                    }
            }
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] += S(P(1e-9)) * add[i];
    }
    // symmetrise the filtered covariance (dynamax symmetrize) and form the filtered mean
    S mf[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) {
        if (i < D) {
            mf[i] = m_pred[i] + delta[i];
#pragma unroll
            for (int j = 0; j < DC; ++j)
                if (j < D && j > i) {
                    const S a = S(P(0.5)) * (Pm[i * D + j] + Pm[j * D + i]);
                    Pm[i * D + j] = a;
                    Pm[j * D + i] = a;
                }
        }
    }
    if (mf_out) {
#pragma unroll
        for (int i = 0; i < DC; ++i) if (i < D) mf_out[i] = mf[i];
    }
    if (Pf_out) {
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pf_out[i] = Pm[i];
    }
    // predict: m = A m_f ; P = A P_f A^T + s Q
    if (Adiag != nullptr) {  // parameter-dependent diagonal dynamics (IBL pupil AR(1) model): A = diag, Q~ = diag
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            if (i < D) {
                m[i] = Adiag[i] * mf[i];
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) Pm[i * D + j] = Adiag[i] * Pm[i * D + j] * Adiag[j] + (i == j ? Qdiag[i] : S(P(0)));
            }
        }
        return ok;
    }
    if (a_identity) {   // A = I (every multi-camera model of the reference): A m_f = m_f and A P A^T = P, bit for bit
#pragma unroll
        for (int i = 0; i < DC; ++i) if (i < D) m[i] = mf[i];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = Pm[i] + s * S(mdl.Q[i]);
        return ok;
    }
    S AP[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) {
        if (i < D) {
            S acc = S(P(0));
#pragma unroll
            for (int k = 0; k < DC; ++k) if (k < D) acc += S(mdl.A[i * D + k]) * mf[k];
            m[i] = acc;
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                if (j < D) {
                    S a2 = S(P(0));
#pragma unroll
                    for (int k = 0; k < DC; ++k) if (k < D) a2 += S(mdl.A[i * D + k]) * Pm[k * D + j];
                    AP[i * D + j] = a2;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < DC; ++i) {
        if (i < D) {
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                if (j < D) {
                    S acc = S(P(0));
#pragma unroll
                    for (int k = 0; k < DC; ++k) if (k < D) acc += AP[i * D + k] * S(mdl.A[j * D + k]);
                    Pm[i * D + j] = acc + s * S(mdl.Q[i * D + j]);
                }
            }
        }
    }
    return ok;
}

template <class P, int OC>
EKS_HD void load_obs(const SeqObs<P>& ob, int O, long long t, P* yv, P* rv) {
#pragma unroll
    for (int o = 0; o < OC; ++o) {
        if (o < O) {
            P y = ob.y_base[ob.y_off[o] + t];
            if (ob.ymean) y -= ob.ymean[o];
            yv[o] = y;
            if (ob.Rconst) rv[o] = ob.Rconst[o];
            else {
                const P v = ob.var_base[ob.var_off[o] + t];
                rv[o] = v > ob.var_floor ? v : ob.var_floor;  // np.clip(ev, 1e-12, None) keeps NaN out only via >
                if (v != v) rv[o] = v;
            }
        }
    }
}

// Filter NLL and d NLL / d s over frames given by frame(i), i in [0, n).
template <class P, int DC, int OC, bool FIXED, bool NL, class FrameFn>
EKS_HD void seq_nll_grad(const Dims<DC, OC, FIXED>& dm, const SeqModel<P>& mdl, const SeqObs<P>& ob, int n,
                         FrameFn frame, P s, P* nll_out, P* dnll_ds_out) {
    using S = Dual<P>;
    const int D = dm.D(), O = dm.O();
    S m[DC], Pm[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) m[i] = S(mdl.m0[i]);
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = S(mdl.S0[i]);
    S sd(s, P(1));
    // per-frame terms in the working precision, their sum in float64: a float32 running sum of ~2000 terms loses
    // ~25 ulps of the loss, which is what the reference's stop rule looks at
    double nll_sum = 0, dnll_sum = 0;
    bool ok = true;
    for (int i = 0; i < n; ++i) {
        P yv[OC], rv[OC];
        load_obs<P, OC>(ob, O, frame(i), yv, rv);
        S nll = S(P(0));
        ok = ekf_step<S, P, DC, OC, FIXED, NL>(dm, mdl, yv, rv, sd, m, Pm, nll, (S*)nullptr, (S*)nullptr) && ok;
        nll_sum += (double)nll.v;
        dnll_sum += (double)nll.d;
    }
    P v = (P)nll_sum, g = (P)dnll_sum;
    if (!ok || !isfinite((double)v)) { v = P(1e12); g = P(0); }  // core.py:650
    *nll_out = v;
    *dnll_ds_out = g;
}

// optax.adam(1.0) state + the reference's tol loop (eks/core.py:654-681); R-precision scalars.
template <class P>
struct AdamState {
    P s_log, mu, nu, prev;
    int iters;
    bool done;
};
template <class P>
EKS_HD void adam_init(AdamState<P>& a, P s_log0) {
    a.s_log = s_log0; a.mu = P(0); a.nu = P(0);
    a.prev = P(INFINITY);
    a.iters = 0; a.done = false;
}
template <class P> EKS_HD P pow_(P a, P b);
template <> EKS_HD float pow_<float>(float a, float b) { return powf(a, b); }
template <> EKS_HD double pow_<double>(double a, double b) { return pow(a, b); }
template <class P> EKS_HD P exp_(P a);
template <> EKS_HD float exp_<float>(float a) { return expf(a); }
template <> EKS_HD double exp_<double>(double a) { return exp(a); }

// current s and the chain factor d s / d s_log (zero outside the clip bounds)
template <class P>
EKS_HD P adam_current_s(const AdamState<P>& a, P lo, P hi, P* dsdlog) {
    const P sc = a.s_log < lo ? lo : (a.s_log > hi ? hi : a.s_log);
    const P s = exp_(sc);
    *dsdlog = (a.s_log >= lo && a.s_log <= hi) ? s : P(0);
    return s;
}
// consume (loss, d loss / d s_log) -> Adam update + stop test
template <class P>
EKS_HD void adam_step(AdamState<P>& a, P loss, P g_log, P lr, P tol, int cap) {
    const P b1 = P(0.9), b2 = P(0.999), eps = P(1e-8);
    const P g = g_log * lr;
    const int count = a.iters + 1;
    a.mu = b1 * a.mu + (P(1) - b1) * g;
    a.nu = b2 * a.nu + (P(1) - b2) * g * g;
    const P mu_hat = a.mu / (P(1) - pow_(b1, P(count)));
    const P nu_hat = a.nu / (P(1) - pow_(b2, P(count)));
    a.s_log = a.s_log + (-mu_hat / (sqrt_(nu_hat) + eps));
    const P pm = a.prev > P(1e-12) ? a.prev : P(1e-12);
    const P rel_tol = tol * fabs(log_(pm));
    const bool stop = isfinite((double)a.prev) ? (fabs(loss - a.prev) < rel_tol + P(1e-6)) : false;
    a.prev = loss;
    a.iters = count;
    a.done = stop || (count >= cap);
}

// In-place Cholesky factorisation / solve for the RTS gain (D x D, runtime or fixed D)
template <class P, int DC>
EKS_HD bool chol_inplace(P* a, int D) {
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (j < D) {
            P d = a[j * D + j];
#pragma unroll
            for (int k = 0; k < DC; ++k) if (k < j) d -= a[j * D + k] * a[j * D + k];
            if (!(d > P(0))) return false;
            d = sqrt_(d);
            a[j * D + j] = d;
#pragma unroll
            for (int i = 0; i < DC; ++i) {
                if (i > j && i < D) {
                    P sacc = a[i * D + j];
#pragma unroll
                    for (int k = 0; k < DC; ++k) if (k < j) sacc -= a[i * D + k] * a[j * D + k];
                    a[i * D + j] = sacc / d;
                }
            }
        }
    }
    return true;
}
template <class P, int DC>
EKS_HD void chol_solve_vec(const P* L, int D, const P* b, P* x) {
    P z[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) {
        if (i < D) {
            P sacc = b[i];
#pragma unroll
            for (int k = 0; k < DC; ++k) if (k < i) sacc -= L[i * D + k] * z[k];
            z[i] = sacc / L[i * D + i];
        }
    }
#pragma unroll
    for (int ii = 0; ii < DC; ++ii) {
        const int i = DC - 1 - ii;
        if (i < D) {
            P sacc = z[i];
#pragma unroll
            for (int k = 0; k < DC; ++k) if (k > i && k < D) sacc -= L[k * D + i] * x[k];
            x[i] = sacc / L[i * D + i];
        }
    }
}

// One RTS step: given the filtered moments (mft, Pft) of frame t and the smoothed moments
// (msn, Vsn) of frame t+1, overwrite (msn, Vsn) with the smoothed moments of frame t.
// G = psd_solve(A P_f A^T + sQ, A P_f)^T with the 1e-9 boost.
template <class P, int DC, int OC, bool FIXED>
EKS_HD void rts_step(const Dims<DC, OC, FIXED>& dm, const SeqModel<P>& mdl, P s, const P* mft, const P* Pft, P* msn,
                     P* Vsn) {
    const int D = dm.D();
        P mp[DC], AP[DC * DC], Sp[DC * DC], L[DC * DC], G[DC * DC];
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            if (i < D) {
                P acc = P(0);
#pragma unroll
                for (int k = 0; k < DC; ++k) if (k < D) acc += mdl.A[i * D + k] * mft[k];
                mp[i] = acc;
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    if (j < D) {
                        P a2 = P(0);
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) a2 += mdl.A[i * D + k] * Pft[k * D + j];
                        AP[i * D + j] = a2;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < DC; ++i)
            if (i < D) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) {
                        P acc = P(0);
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) acc += AP[i * D + k] * mdl.A[j * D + k];
                        Sp[i * D + j] = acc + s * mdl.Q[i * D + j];
                    }
            }
#pragma unroll
        for (int i = 0; i < DC; ++i)
            if (i < D) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) {
                        L[i * D + j] = P(0.5) * (Sp[i * D + j] + Sp[j * D + i]);
                        if (i == j) L[i * D + j] += P(1e-9);
                    }
            }
        const bool ok = chol_inplace<P, DC>(L, D);
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            if (j < D) {
                P bcol[DC], x[DC];
#pragma unroll
                for (int i = 0; i < DC; ++i) if (i < D) bcol[i] = AP[i * D + j];
                if (ok) chol_solve_vec<P, DC>(L, D, bcol, x);
#pragma unroll
                for (int i = 0; i < DC; ++i) if (i < D) G[j * D + i] = ok ? x[i] : P(NAN);
            }
        }
        P dV[DC * DC], GdV[DC * DC], mst[DC], Vst[DC * DC];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) dV[i] = Vsn[i] - Sp[i];
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            if (i < D) {
                P acc = mft[i];
#pragma unroll
                for (int k = 0; k < DC; ++k) if (k < D) acc += G[i * D + k] * (msn[k] - mp[k]);
                mst[i] = acc;
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    if (j < D) {
                        P a2 = P(0);
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) a2 += G[i * D + k] * dV[k * D + j];
                        GdV[i * D + j] = a2;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < DC; ++i)
            if (i < D) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < D) {
                        P acc = P(0);
#pragma unroll
                        for (int k = 0; k < DC; ++k) if (k < D) acc += GdV[i * D + k] * G[j * D + k];
                        Vst[i * D + j] = Pft[i * D + j] + acc;
                    }
            }
#pragma unroll
        for (int i = 0; i < DC; ++i) if (i < D) msn[i] = mst[i];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Vsn[i] = Vst[i];
}

// EKF filter + RTS smoother for one sequence; mf/Pf are scratch of size T*D, T*D*D (row-major per
// frame); ms/Vs are the outputs in the reference layout (T,D), (T,D,D).
template <class P, int DC, int OC, bool FIXED, bool NL>
EKS_HD void seq_smooth(const Dims<DC, OC, FIXED>& dm, const SeqModel<P>& mdl, const SeqObs<P>& ob, int T, P s,
                       P* mf, P* Pf, P* ms, P* Vs) {
    const int D = dm.D(), O = dm.O();
    P m[DC], Pm[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) m[i] = mdl.m0[i];
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = mdl.S0[i];
    P nll = P(0);
    for (int t = 0; t < T; ++t) {
        P yv[OC], rv[OC];
        load_obs<P, OC>(ob, O, t, yv, rv);
        ekf_step<P, P, DC, OC, FIXED, NL, true>(dm, mdl, yv, rv, s, m, Pm, nll, mf + (long long)t * D,
                                                Pf + (long long)t * D * D);
    }
    // backward pass: G = psd_solve(A P_f A^T + sQ, A P_f)^T, boost 1e-9
    P msn[DC], Vsn[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) { msn[i] = mf[(long long)(T - 1) * D + i]; ms[(long long)(T - 1) * D + i] = msn[i]; }
#pragma unroll
    for (int i = 0; i < DC * DC; ++i)
        if (i < D * D) { Vsn[i] = Pf[(long long)(T - 1) * D * D + i]; Vs[(long long)(T - 1) * D * D + i] = Vsn[i]; }
    for (int t = T - 2; t >= 0; --t) {
        P mft[DC], Pft[DC * DC];
#pragma unroll
        for (int i = 0; i < DC; ++i) if (i < D) mft[i] = mf[(long long)t * D + i];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pft[i] = Pf[(long long)t * D * D + i];
        rts_step<P, DC, OC, FIXED>(dm, mdl, s, mft, Pft, msn, Vsn);
#pragma unroll
        for (int i = 0; i < DC; ++i) if (i < D) ms[(long long)t * D + i] = msn[i];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Vs[(long long)t * D * D + i] = Vsn[i];
    }
}

}  // namespace eks
