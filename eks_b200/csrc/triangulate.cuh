// triangulate.cuh -- per-point multi-view triangulation in __host__ __device__ templates (fp64), SURVEY 8 row f3.
//
// Restates what the reference obtains from aniposelib's CameraGroup.triangulate(points, fast=True)
// (eks/multicam_smoother.py:888-911 -> triangulate_3d_models): undistort every view to normalised image
// coordinates, linear (DLT) triangulation of every camera pair, nan-median over the pairs.  aniposelib is not in
// this image; the restatement follows the OpenCV primitives that the host mirror (eks_b200.multicam_smoother.
// CameraGroup, cv2.undistortPoints + cv2.triangulatePoints) calls, and is checked against them on the CPU
// (tests/hostcheck).  Camera parameters use the packed layout of the pinhole emission (CAM_STRIDE = 29).
#pragma once
#include "common.cuh"

namespace eks {

#ifndef EKS_HD
#define EKS_HD __host__ __device__ inline
#endif

// cv::undistortPoints with no R / P and the default criteria (5 fixed-point iterations): pixel -> normalised
// coordinates.  cam: R(9) t(3) fx fy cx cy skew k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (OpenCV ignores the skew here).
EKS_HD void undistort_point(const double* cam, double u, double v, double& xo, double& yo) {
    const double fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15];
    const double* k = cam + 17;   // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4
    double x = (u - cx) / fx, y = (v - cy) / fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    xo = x; yo = y;
}

// Eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix G (cyclic Jacobi, fp64).  G is destroyed.
EKS_HD void smallest_eigvec4(double* G, double* vec) {
    double V[16];
    for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += G[p * 4 + q] * G[p * 4 + q];
        if (off < 1e-300) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = G[p * 4 + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (G[q * 4 + q] - G[p * 4 + p]) / (2 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(tt * tt + 1), s = tt * c;
                for (int r = 0; r < 4; ++r) {   // G <- G J
                    const double grp = G[r * 4 + p], grq = G[r * 4 + q];
                    G[r * 4 + p] = c * grp - s * grq;
                    G[r * 4 + q] = s * grp + c * grq;
                }
                for (int r = 0; r < 4; ++r) {   // G <- J^T G
                    const double gpr = G[p * 4 + r], gqr = G[q * 4 + r];
                    G[p * 4 + r] = c * gpr - s * gqr;
                    G[q * 4 + r] = s * gpr + c * gqr;
                }
                for (int r = 0; r < 4; ++r) {
                    const double vrp = V[r * 4 + p], vrq = V[r * 4 + q];
                    V[r * 4 + p] = c * vrp - s * vrq;
                    V[r * 4 + q] = s * vrp + c * vrq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i) if (G[i * 4 + i] < G[best * 4 + best]) best = i;
    for (int r = 0; r < 4; ++r) vec[r] = V[r * 4 + best];
}

// The same eigenvector by inverse iteration on G + mu I (LDL^T factorisation, mu = 1e-15 trace G): for a triangulation
// system the smallest eigenvalue (the squared residual) is far below the other three, so two or three solves converge
// to rounding -- ~400 flops against ~3600 for the Jacobi sweeps.  Returns false (caller falls back to Jacobi) if the
// factorisation breaks down or the iteration has not converged after 6 solves (degenerate geometry).  G is kept.
EKS_HD bool smallest_eigvec4_invit(const double* G, double* vec) {
    const double tr = G[0] + G[5] + G[10] + G[15];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    const double mu = 1e-15 * tr;
    // LDL^T of the symmetric 4x4 matrix (unit lower L, diagonal D), fully unrolled
    const double a00 = G[0] + mu, a10 = G[4], a11 = G[5] + mu, a20 = G[8], a21 = G[9], a22 = G[10] + mu, a30 = G[12],
                 a31 = G[13], a32 = G[14], a33 = G[15] + mu;
    const double d0 = a00;
    if (!(d0 > 0.0)) return false;
    const double i0 = 1.0 / d0, l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
    const double d1 = a11 - l10 * a10;
    if (!(d1 > 0.0)) return false;
    const double i1 = 1.0 / d1, l21 = (a21 - l20 * a10) * i1, l31 = (a31 - l30 * a10) * i1;
    const double d2 = a22 - l20 * a20 - l21 * l21 * d1;
    if (!(d2 > 0.0)) return false;
    const double i2 = 1.0 / d2, l32 = (a32 - l30 * a20 - l31 * l21 * d1) * i2;
    const double d3 = a33 - l30 * a30 - l31 * l31 * d1 - l32 * l32 * d2;
    if (!(d3 > 0.0)) return false;
    const double i3 = 1.0 / d3;
    double x0 = 0.5, x1 = 0.5, x2 = 0.5, x3 = 0.5;
    for (int it = 0; it < 6; ++it) {
        // L y = x ; D z = y ; L^T w = z
        const double y0 = x0, y1 = x1 - l10 * y0, y2 = x2 - l20 * y0 - l21 * y1, y3 = x3 - l30 * y0 - l31 * y1 - l32 * y2;
        const double w3 = y3 * i3, w2 = y2 * i2 - l32 * w3, w1 = y1 * i1 - l21 * w2 - l31 * w3,
                     w0 = y0 * i0 - l10 * w1 - l20 * w2 - l30 * w3;
        const double nrm = sqrt(w0 * w0 + w1 * w1 + w2 * w2 + w3 * w3);
        if (!(nrm > 0.0) || !isfinite(nrm)) return false;
        const double inv = 1.0 / nrm, n0 = w0 * inv, n1 = w1 * inv, n2 = w2 * inv, n3 = w3 * inv;
        // converged when the direction no longer changes (up to sign)
        const double dot = n0 * x0 + n1 * x1 + n2 * x2 + n3 * x3;
        const double sg = dot < 0 ? -1.0 : 1.0;
        const double e0 = n0 - sg * x0, e1 = n1 - sg * x1, e2 = n2 - sg * x2, e3 = n3 - sg * x3;
        x0 = n0; x1 = n1; x2 = n2; x3 = n3;
        if (it > 0 && e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3 < 1e-30) {
            vec[0] = x0; vec[1] = x1; vec[2] = x2; vec[3] = x3;
            return true;
        }
    }
    return false;
}

// cv::triangulatePoints for one point seen by two cameras with projection matrices [R|t] and normalised
// coordinates: null vector of the 4x4 system (x P_3 - P_1; y P_3 - P_2 for both views), dehomogenised.
EKS_HD void triangulate_pair(const double* cam1, double x1, double y1, const double* cam2, double x2, double y2,
                             double* X) {
    double A[16];
    const double* cams[2] = {cam1, cam2};
    const double xs[2] = {x1, x2}, ys[2] = {y1, y2};
    for (int c = 0; c < 2; ++c) {
        const double* R = cams[c];
        const double* t = cams[c] + 9;
        for (int j = 0; j < 4; ++j) {
            const double p0 = j < 3 ? R[0 * 3 + j] : t[0];
            const double p1 = j < 3 ? R[1 * 3 + j] : t[1];
            const double p2 = j < 3 ? R[2 * 3 + j] : t[2];
            A[(2 * c) * 4 + j] = xs[c] * p2 - p0;
            A[(2 * c + 1) * 4 + j] = ys[c] * p2 - p1;
        }
    }
    double G[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int r = 0; r < 4; ++r) s += A[r * 4 + i] * A[r * 4 + j];
            G[i * 4 + j] = s;
        }
    double h[4];
    if (!smallest_eigvec4_invit(G, h)) smallest_eigvec4(G, h);
    for (int i = 0; i < 3; ++i) X[i] = h[i] / h[3];
}

// np.nanmedian of n <= 28 values (camera pairs of up to 8 views)
EKS_HD double nanmedian_small(double* v, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i) if (!isnan(v[i])) v[m++] = v[i];
    if (m == 0) return nan("");
    for (int i = 1; i < m; ++i) {   // insertion sort
        const double x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    return (m & 1) ? v[m / 2] : 0.5 * (v[m / 2 - 1] + v[m / 2]);
}

// One point seen by V cameras (pixel coordinates uv[2 V]): nan-median over all camera pairs of the pairwise DLT.
EKS_HD void triangulate_point(const double* cams, int V, const double* uv, double* X) {
    double xn[8], yn[8];
    for (int c = 0; c < V; ++c) undistort_point(cams + c * CAM_STRIDE, uv[2 * c], uv[2 * c + 1], xn[c], yn[c]);
    double px[28], py[28], pz[28];
    int np_ = 0;
    for (int a = 0; a < V; ++a)
        for (int b = a + 1; b < V; ++b) {
            double P3[3];
            triangulate_pair(cams + a * CAM_STRIDE, xn[a], yn[a], cams + b * CAM_STRIDE, xn[b], yn[b], P3);
            px[np_] = P3[0]; py[np_] = P3[1]; pz[np_] = P3[2];
            ++np_;
        }
    X[0] = nanmedian_small(px, np_);
    X[1] = nanmedian_small(py, np_);
    X[2] = nanmedian_small(pz, np_);
}

}  // namespace eks
