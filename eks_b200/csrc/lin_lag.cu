// lin_lag.cu -- the smoothing-parameter optimiser of LINEAR models with A = I and a constant observation noise (the
// loss path of the multi-camera PCA-latent model: eks/core.py:562-699 on the model of eks/multicam_smoother.py:409-443)
// with ONE pass over the observations instead of one pass per Adam evaluation -- the matrix form of diag_lag.cu.
//
// With A = I the predicted mean obeys m_{t+1} = m_t + K e_t, e_t = y_t - C m_t.  Let C = Co Rc (Co: O x D orthonormal,
// Rc: D x D) and U = [Co | V] an orthogonal completion (V spans the part of observation space the state never reaches).
// In the basis U the innovation splits into e1 = Co^T e (D values) and e2 = V^T e = V^T y =: w (data, not filtered), and
// once the covariance recursion has reached its fixed point (constant K, N = S^-1)
//       e1_{t+1} = Phi e1_t + Gam g_{t+1},   g_i = [Co^T (y_i - y_{i-1}) ; V^T y_{i-1}],   Gam = [I, -Rc K2],
//       Phi = I - Rc K1,   [K1 K2] = K U            (Phi is a contraction; g is stationary: increments and residuals)
// so the quadratic part of the NLL, sum_t [e1;w]^T (U^T N U) [e1;w], is a quadratic form in the lag statistics
//       Rg_m[a][b] = sum_i g_i[a] g_{i+m}[b],  m = 0 .. W-1            (O x O x W numbers per sequence)
// plus couplings with the state at frame T0 and a correction for the last frame (the terms A, B, C of linlag_eval;
// checked against the sequential filter to 1e-10 in a NumPy prototype and by the parity tests).  mlag_stats_kernel
// computes Rg in one streaming pass; lin_lag_opt_kernel then runs the WHOLE Adam loop of a block in one warp: per
// evaluation the exact filter (sequential scalar updates, covariance recursion with its s-sensitivity in forward duals)
// over the first T0 frames, the steady-state algebra in Dual<double>, the Adam step and the reference's stop rule.
// The head filter and the steady-state gain use the information form of the update (R diagonal and constant:
// P+ = (P^-1 + C^T R^-1 C)^-1, K = P+ C^T R^-1, N = R^-1 - R^-1 C P+ C^T R^-1): the same numbers as the sequential scalar
// updates of the run-parallel path up to float64 rounding.
//
// The closed form is used only when it is exact to rounding: A = I, one contiguous span, n >= T0 + 4 W, the covariance
// recursion converged within T0 = 256 frames, and |Phi^W| below 1e-7 (float32 mode, W = 128) / 1e-13 (float64 mode,
// W = 256).  Otherwise the block is flagged and the caller (generic_runs_optimize) runs the run-parallel path.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "ekf_generic.cuh"
#include "generic.cuh"
#include "lin_lag.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int ML_T0 = 256;      // statistics are kept for the start frames T0 = 32, 64, ..., 256 (8 levels): an
constexpr int ML_NT0 = 8;       // evaluation walks the head only to the first level past its variance transient
constexpr int ML_LSTEP = ML_T0 / ML_NT0;
__host__ __device__ constexpr int ml_level_T0(int lvl) { return ML_T0 - lvl * ML_LSTEP; }   // lvl 0 = 256 ... 7 = 32
constexpr int ML_CH = 4096;     // increments per shared-memory tile
constexpr int ML_RM = 16;       // lags per thread
constexpr int ML_RP = 16;       // frames per thread and step
constexpr int ML_NT = 256;
constexpr int ML_CPB = 8;       // tiles per CTA

constexpr int ML_MAXO = 8;      // observation channels the closed form is compiled for (D = 3 latent dimensions)

// per-sequence observation basis (ml_basis_kernel): U = [Co | V] row major [channel][basis vector], C = Co Rc
struct MLBasis {
    double U[ML_MAXO * ML_MAXO];
    double Rc[9];
    double ym[ML_MAXO];
};

template <class P>
struct MLagStatArgs {
    PlaneView y;
    int B, O, t_begin, n, nsig, nchunk, nx;   // nsig = n + 1: the signal g_i is defined for i = 1 .. n
    const MLBasis* basis;
    P* gsig;           // [B][O][nsig]: the stationary signal g_i, i = 0 .. n (ml_signal_kernel), working precision
    double* partial;   // [B O O][nx][W]
    double* R;         // [ML_NT0][B O O][W]
};

template <class P> struct MLVec;
template <> struct MLVec<float> { using type = float4; static constexpr int VW = 4; };
template <> struct MLVec<double> { using type = double2; static constexpr int VW = 2; };

template <class P>
__device__ __forceinline__ int ml_phys(int x) { return x + (x >> 4) * (16 / (int)sizeof(P)); }

// C = Co Rc by Gram-Schmidt with re-orthogonalisation, then an orthonormal completion V from the unit vectors with the
// largest residuals.  One thread per sequence; D = 3.
template <class P>
__global__ void ml_basis_kernel(int B, int O, const P* __restrict__ C, const P* __restrict__ ymean,
                                MLBasis* __restrict__ basis) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr int D = 3;
    MLBasis bs;
    for (int i = 0; i < ML_MAXO * ML_MAXO; ++i) bs.U[i] = 0.0;
    for (int i = 0; i < 9; ++i) bs.Rc[i] = 0.0;
    for (int k = 0; k < ML_MAXO; ++k) bs.ym[k] = (k < O && ymean) ? (double)ymean[(long long)b * O + k] : 0.0;
    for (int j = 0; j < O; ++j) {
        double v[ML_MAXO];
        int pick = -1;
        if (j < D) {
            for (int k = 0; k < O; ++k) v[k] = (double)C[((long long)b * O + k) * D + j];
        } else {   // unit vector with the largest component outside the span so far
            double best = -1.0;
            for (int k = 0; k < O; ++k) {
                double r = 1.0;
                for (int q = 0; q < j; ++q) r -= bs.U[k * O + q] * bs.U[k * O + q];
                if (r > best) { best = r; pick = k; }
            }
            for (int k = 0; k < O; ++k) v[k] = (k == pick) ? 1.0 : 0.0;
        }
        for (int pass = 0; pass < 2; ++pass)
            for (int q = 0; q < j; ++q) {
                double dot = 0.0;
                for (int k = 0; k < O; ++k) dot += bs.U[k * O + q] * v[k];
                for (int k = 0; k < O; ++k) v[k] -= dot * bs.U[k * O + q];
                if (j < D) bs.Rc[q * D + j] += dot;
            }
        double nrm = 0.0;
        for (int k = 0; k < O; ++k) nrm += v[k] * v[k];
        nrm = sqrt(nrm);
        if (j < D) bs.Rc[j * D + j] = nrm;
        for (int k = 0; k < O; ++k) bs.U[k * O + j] = v[k] / nrm;
    }
    basis[b] = bs;
}

// g_i[o], i = 1 .. n (0 outside): o < 3: Co^T (y_i - y_{i-1}) (i <= n - 1); o >= 3: V^T (y_{i-1} - ymean)
template <class P>
__device__ __forceinline__ double ml_g(const P* __restrict__ yb, const long long* __restrict__ off, const MLBasis& bs,
                                       int O, int n, int i, int o) {
    if (i < 1 || i > n) return 0.0;
    double acc = 0.0;
    if (o < 3) {
        if (i > n - 1) return 0.0;
        for (int k = 0; k < O; ++k) acc += bs.U[k * O + o] * ((double)__ldg(yb + off[k] + i) - (double)__ldg(yb + off[k] + i - 1));
    } else {
        for (int k = 0; k < O; ++k) acc += bs.U[k * O + o] * ((double)__ldg(yb + off[k] + i - 1) - bs.ym[k]);
    }
    return acc;
}

// The signal planes: g[b][o][i] for i = 0 .. n, formed in float64 from the O observation planes and rounded ONCE to the
// working precision (the statistics kernel then stages plain loads; forming the signal while staging, once per channel
// PAIR, made it 63 % non-FMA instructions).  grid = (ceil(nsig / 256), B O).
template <class P>
__global__ void __launch_bounds__(256) ml_signal_kernel(const __grid_constant__ MLagStatArgs<P> a) {
    __shared__ MLBasis bs;
    const int bo = blockIdx.y, b = bo / a.O, o = bo - b * a.O;
    for (int i = threadIdx.x; i < (int)(sizeof(MLBasis) / sizeof(double)); i += 256)
        reinterpret_cast<double*>(&bs)[i] = reinterpret_cast<const double*>(a.basis + b)[i];
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.nsig) return;
    const P* yb = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin;
    a.gsig[(long long)bo * a.nsig + i] = (P)ml_g<P>(yb, a.y.chan_off, bs, a.O, a.n, i, o);
}

// Rg_m[a][c] partial sums.  grid = (nx, B O O); CTA (x, (b, a, c)) handles the tiles x, x + nx, ... : the signal
// components a (frames of the tile) and c (tile + W halo) are formed in float64 from the observations and staged in
// shared memory in the working precision; warp w owns lag groups w, w + 8, ... and lane l the frames
// (step * 32 + l) * 16 ... + 15: a 16 x 16 register tile of products per step.
template <class P, int W>
__global__ void __launch_bounds__(ML_NT) mlag_stats_kernel(const __grid_constant__ MLagStatArgs<P> a) {
    constexpr int PADE = 16 / (int)sizeof(P);
    constexpr int NLOG = ML_CH + W;
    constexpr int NPHYS = NLOG + (NLOG / 16) * PADE + PADE;
    constexpr int NG = W / ML_RM;
    constexpr int GPW = (NG + 7) / 8;
    constexpr int VW = MLVec<P>::VW;
    using V = typename MLVec<P>::type;
    extern __shared__ __align__(16) unsigned char ml_smem[];
    P* smA = reinterpret_cast<P*>(ml_smem);
    P* smC = smA + NPHYS;
    const int O = a.O;
    const int bac = blockIdx.y, b = bac / (O * O), ac = bac - b * O * O, ca = ac / O, cc = ac - ca * O;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const P* ga = a.gsig + ((long long)b * O + ca) * a.nsig;
    const P* gc = a.gsig + ((long long)b * O + cc) * a.nsig;
    double accd[GPW][ML_RM];
#pragma unroll
    for (int q = 0; q < GPW; ++q)
#pragma unroll
        for (int j = 0; j < ML_RM; ++j) accd[q][j] = 0.0;
    for (int chunk = blockIdx.x; chunk < a.nchunk; chunk += a.nx) {
        const int i0 = ML_T0 + 1 + chunk * ML_CH;
        __syncthreads();
        {   // every load of the tile is issued before the first use
            constexpr int U = (NLOG + ML_NT - 1) / ML_NT;
            P vc[U], va[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = threadIdx.x + u * ML_NT, i = i0 + x;
                vc[u] = (x < NLOG && i < a.nsig) ? __ldg(gc + i) : P(0);
                va[u] = (x < ML_CH && i < a.nsig) ? __ldg(ga + i) : P(0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = threadIdx.x + u * ML_NT;
                if (x < NLOG) smC[ml_phys<P>(x)] = vc[u];
                if (x < ML_CH) smA[ml_phys<P>(x)] = va[u];
            }
        }
        __syncthreads();
        const int nvalid = min(ML_CH, a.nsig - i0);
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int g = warp + 8 * q;
            if (g >= NG) break;
            const int m0 = g * ML_RM;
            P acc[ML_RM];
#pragma unroll
            for (int j = 0; j < ML_RM; ++j) acc[j] = P(0);
            for (int step = 0; step < ML_CH / (32 * ML_RP); ++step) {
                const int p = (step * 32 + lane) * ML_RP;
                if (step * 32 * ML_RP >= nvalid) break;
                P av[ML_RP], bv[ML_RP + ML_RM];
                const P* pa = smA + ml_phys<P>(p);
                const P* pb0 = smC + ml_phys<P>(p + m0);
                const P* pb1 = smC + ml_phys<P>(p + m0 + 16);
#pragma unroll
                for (int i = 0; i < ML_RP / VW; ++i) {
                    const V v = *reinterpret_cast<const V*>(pa + i * VW);
                    const V w0 = *reinterpret_cast<const V*>(pb0 + i * VW);
                    const V w1 = *reinterpret_cast<const V*>(pb1 + i * VW);
                    const P* ev = reinterpret_cast<const P*>(&v);
                    const P* e0 = reinterpret_cast<const P*>(&w0);
                    const P* e1 = reinterpret_cast<const P*>(&w1);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                        av[i * VW + k] = ev[k];
                        bv[i * VW + k] = e0[k];
                        bv[16 + i * VW + k] = e1[k];
                    }
                }
#pragma unroll
                for (int i = 0; i < ML_RP; ++i)
#pragma unroll
                    for (int j = 0; j < ML_RM; ++j) acc[j] = fma(av[i], bv[i + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < ML_RM; ++j) accd[q][j] += (double)acc[j];
        }
    }
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        const int g = warp + 8 * q;
        if (g >= NG) break;
#pragma unroll
        for (int j = 0; j < ML_RM; ++j) {
            const double v = warp_sum(accd[q][j]);
            if (lane == 0) a.partial[((long long)bac * a.nx + blockIdx.x) * W + g * ML_RM + j] = v;
        }
    }
}

// fixed-order sum of the per-CTA partials, one thread per (sequence, a, c, lag); the lower levels add the few extra
// products of the signal between their start frame and frame 256.
template <class P, int W>
__global__ void mlag_reduce_kernel(const __grid_constant__ MLagStatArgs<P> a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int O = a.O;
    const long long per = (long long)a.B * O * O * W;
    if (idx >= per) return;
    const long long bac = idx / W;
    const int m = (int)(idx - bac * W);
    double s = 0;
    for (int x = 0; x < a.nx; ++x) s += a.partial[(bac * a.nx + x) * W + m];
    a.R[idx] = s;
    const int b = (int)(bac / (O * O)), ac = (int)(bac - (long long)b * O * O), ca = ac / O, cc = ac - ca * O;
    const MLBasis& bs = a.basis[b];
    const P* yb = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin;
    int hi = ML_T0;
    for (int lvl = 1; lvl < ML_NT0; ++lvl) {
        const int lo = ml_level_T0(lvl);
        for (int i = hi; i > lo; --i)
            s += ml_g<P>(yb, a.y.chan_off, bs, O, a.n, i, ca) * ml_g<P>(yb, a.y.chan_off, bs, O, a.n, i + m, cc);
        a.R[lvl * per + idx] = s;
        hi = lo;
    }
}

// ---------------------------------------------------------------------------------------------------- the optimiser
using SD = Dual<double>;

template <int N>
__device__ __forceinline__ void ml_matmul(const SD* __restrict__ A, const SD* __restrict__ B, SD* __restrict__ Cm) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            SD acc(0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) acc += A[i * N + k] * B[k * N + j];
            Cm[i * N + j] = acc;
        }
}

template <int N>
__device__ __forceinline__ void ml_matmul_tn(const SD* __restrict__ A, const SD* __restrict__ B, SD* __restrict__ Cm) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            SD acc(0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) acc += A[k * N + i] * B[k * N + j];
            Cm[i * N + j] = acc;
        }
}

__device__ __forceinline__ SD ml_warp_sum(SD v) { return SD(warp_sum(v.v), warp_sum(v.d)); }
__device__ __forceinline__ SD ml_scale(double c, SD v) { return SD(c * v.v, c * v.d); }

struct LinLagArgs {
    const double* R;        // [ML_NT0][B O O][W]
    const MLBasis* basis;   // [B]
    int W, nlog;            // lags; log2(W)
    double tolF;
    int* flag;              // [n_blocks]: 1 = the closed form did not apply, run the run-parallel path
};

// One evaluation of one sequence by a whole warp (sequential parts redundantly on every lane, the lag series split over
// the lanes).  Returns false if the closed form does not apply.
template <class P, int OC>
__device__ bool linlag_eval(const GArgs<P>& a, const LinLagArgs& la, int b, double s_val, int lane, double& nll_out,
                            double& dnll_out) {
    constexpr int D = 3, O = OC, E = OC - 3;
    const int n = a.sp.total, t_begin = a.sp.start[0], W = la.W;
    const double HALF_LOG2PI = 0.91893853320467274178;
    const P* yb = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + t_begin;
    const MLBasis& bs = la.basis[b];
    double Cm[O * D], Qm[D * D], rc[O];
    const P* Ap = a.A + (long long)b * D * D;
    bool a_id = true;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) a_id = a_id && (Ap[i * D + j] == (i == j ? P(1) : P(0)));
    if (!a_id) return false;
#pragma unroll
    for (int i = 0; i < O * D; ++i) Cm[i] = (double)a.C[(long long)b * O * D + i];
#pragma unroll
    for (int i = 0; i < D * D; ++i) Qm[i] = (double)a.Q[(long long)b * D * D + i];
#pragma unroll
    for (int g = 0; g < O; ++g) rc[g] = (double)a.Rconst[(long long)b * O + g];
    const SD s(s_val, 1.0);
    SD Pm[D * D], mu[D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) Pm[i] = SD((double)a.S0[(long long)b * D * D + i]);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = SD((double)a.m0[(long long)b * D + i]);
    SD lsum(0.0), nll(0.0);
    bool bad = false, singular = false;
    // INFORMATION FORM of the measurement update (R is diagonal and constant): with J = C^T R^-1 C,
    //     P+ = (P^-1 + J)^-1,   K = P+ C^T R^-1,   log det S = sum log r_g + log(det P det(P^-1 + J)),
    //     e^T S^-1 e = e^T R^-1 e - z^T P+ z,  z = C^T R^-1 e                       (Woodbury / determinant lemma)
    // -- two symmetric 3 x 3 inverses per frame instead of O rank-one updates (O divisions, O logarithms): the same
    // numbers as the sequential scalar updates of ekf_step up to float64 rounding, at half the dependent instructions
    // (this loop is the critical path of the optimiser: one warp per sequence, ~45 frames per evaluation).
    double Jm[D * D], CtRi[D * O], ri[O], sumlogr = 0.0;
#pragma unroll
    for (int g = 0; g < O; ++g) { ri[g] = 1.0 / rc[g]; sumlogr += log(rc[g]); if (!(rc[g] > 0.0)) bad = true; }
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int g = 0; g < O; ++g) CtRi[i * O + g] = Cm[g * D + i] * ri[g];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int g = 0; g < O; ++g) acc += Cm[g * D + i] * ri[g] * Cm[g * D + j];
            Jm[i * D + j] = acc;
        }
    }
    SD Pp[D * D];          // filtered covariance P+ of the current frame
    // symmetric 3 x 3 inverse by cofactors; returns the determinant
    auto inv3 = [](const SD* A, SD* Inv) {
        const SD c00 = A[4] * A[8] - A[5] * A[5], c01 = A[2] * A[5] - A[1] * A[8], c02 = A[1] * A[5] - A[2] * A[4];
        const SD c11 = A[0] * A[8] - A[2] * A[2], c12 = A[1] * A[2] - A[0] * A[5], c22 = A[0] * A[4] - A[1] * A[1];
        const SD det = A[0] * c00 + A[1] * c01 + A[2] * c02;
        const SD id = SD(1.0) / det;
        Inv[0] = c00 * id; Inv[1] = c01 * id; Inv[2] = c02 * id;
        Inv[3] = Inv[1];   Inv[4] = c11 * id; Inv[5] = c12 * id;
        Inv[6] = Inv[2];   Inv[7] = Inv[5];   Inv[8] = c22 * id;
        return det;
    };
    // one covariance step: Pm (predicted) -> Pp (filtered), lsum = log det S; then Pm <- Pp + s Q
    auto cov_step = [&]() {
        SD Lam[D * D];
        const SD detP = inv3(Pm, Lam);
#pragma unroll
        for (int i = 0; i < D * D; ++i) Lam[i] = Lam[i] + SD(Jm[i]);
        const SD detL = inv3(Lam, Pp);
        const SD dd = detP * detL;                         // det(I + P J) = det S / det R
        if (!(detP.v > 0) || !(dd.v > 0) || !isfinite(dd.v)) singular = true;   // (near-)singular P: not for this form
        lsum = SD(sumlogr) + log_(dd);
#pragma unroll
        for (int i = 0; i < D * D; ++i) Pm[i] = Pp[i] + ml_scale(Qm[i], s);
    };
    // the observations of frame t through the current filtered covariance: NLL terms and the mean update
    auto obs_step = [&](int t) {
        SD z[D], q(0.0);
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = SD(0.0);
#pragma unroll
        for (int g = 0; g < O; ++g) {
            SD e((double)__ldg(yb + a.y.chan_off[g] + t) - bs.ym[g]);
#pragma unroll
            for (int j = 0; j < D; ++j) e -= ml_scale(Cm[g * D + j], mu[j]);
            q += ml_scale(ri[g], e * e);
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] += ml_scale(CtRi[i * O + g], e);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            SD pz(0.0);
#pragma unroll
            for (int j = 0; j < D; ++j) pz += Pp[i * D + j] * z[j];
            q -= z[i] * pz;
            mu[i] += pz;
        }
        nll += SD((double)O * HALF_LOG2PI) + ml_scale(0.5, lsum) + ml_scale(0.5, q);
    };

    // ---- head: exact filter; the covariance recursion is followed until its extrapolated distance to the fixed point
    // (geometric convergence: step * ratio / (1 - ratio)) is below `tol` (relative), or the steps have reached the
    // rounding floor of float64.  float32 mode: 1e-10 -- the loss of 10^6 frames moves by < 1e-3, a hundredth of its
    // float32 resolution; float64 mode: 1e-13.
    const double tol = sizeof(P) == 4 ? 1e-10 : 1e-13, floor_rel = 64.0 * 2.220446049250313e-16;
    double prev_cv = INFINITY, prev_cd = INFINITY;
    int stall = 0, t = 0;
    bool conv = false;
    for (; t < ML_T0 && !conv; ++t) {
        double old_v[D * D], old_d[D * D];
#pragma unroll
        for (int i = 0; i < D * D; ++i) { old_v[i] = Pm[i].v; old_d[i] = Pm[i].d; }
        cov_step();
        obs_step(t);
        double cv = 0, cd = 0, sv = 0, sdv = 0;
#pragma unroll
        for (int i = 0; i < D * D; ++i) {
            cv = fmax(cv, fabs(Pm[i].v - old_v[i])); cd = fmax(cd, fabs(Pm[i].d - old_d[i]));
            sv = fmax(sv, fabs(Pm[i].v)); sdv = fmax(sdv, fabs(Pm[i].d));
        }
        // cv * r / (1 - r) <= tol * sv with r = cv / prev_cv (capped at 0.999), without the divisions
        const bool shrink_v = cv < 0.999 * prev_cv, shrink_d = cd < 0.999 * prev_cd;
        const bool cvok = cv <= floor_rel * sv || (shrink_v && isfinite(prev_cv) && cv * cv <= tol * sv * (prev_cv - cv)) ||
                          (!shrink_v && isfinite(prev_cv) && cv * 999.0 <= tol * sv);
        const bool cdok = cd <= floor_rel * sdv || (shrink_d && isfinite(prev_cd) && cd * cd <= tol * sdv * (prev_cd - cd)) ||
                          (!shrink_d && isfinite(prev_cd) && cd * 999.0 <= tol * sdv);
        if (cv >= prev_cv && cd >= prev_cd) ++stall;
        conv = (cvok && cdok) || stall >= 24;
        prev_cv = cv; prev_cd = cd;
    }
    if (!conv || singular) return false;   // the run-parallel path (sequential scalar updates) handles these
    cov_step();                         // steady gains from the converged predicted covariance
    int lvl = ML_NT0 - 1;               // smallest statistics start frame >= the transient length
    while (lvl > 0 && ml_level_T0(lvl) < t) --lvl;
    const int T0 = ml_level_T0(lvl);
    for (; t < T0; ++t) obs_step(t);
    if (bad) { nll_out = nan(""); dnll_out = 0.0; return true; }

    // ---- steady-state algebra in the basis U
    SD Phi[D * D], Bm[D * (E > 0 ? E : 1)], N11[D * D], N12[D * (E > 0 ? E : 1)], N22[(E > 0 ? E : 1) * (E > 0 ? E : 1)];
    SD e0h[D];
    {
        SD Nn[O * O], Kj[D * O];      // K = P+ C^T R^-1,  N = S^-1 = R^-1 - R^-1 C P+ C^T R^-1
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int c = 0; c < O; ++c) {
                SD acc(0.0);
#pragma unroll
                for (int j = 0; j < D; ++j) acc += ml_scale(CtRi[j * O + c], Pp[i * D + j]);
                Kj[i * O + c] = acc;
            }
#pragma unroll
        for (int x = 0; x < O; ++x)
#pragma unroll
            for (int c = 0; c < O; ++c) {
                SD acc(x == c ? ri[x] : 0.0);
#pragma unroll
                for (int i = 0; i < D; ++i) acc -= ml_scale(CtRi[i * O + x], Kj[i * O + c]);
                Nn[x * O + c] = acc;
            }
        SD Kh[D * O];          // K U
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int o = 0; o < O; ++o) {
                SD acc(0.0);
#pragma unroll
                for (int k = 0; k < O; ++k) acc += ml_scale(bs.U[k * O + o], Kj[i * O + k]);
                Kh[i * O + o] = acc;
            }
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                SD acc(i == j ? 1.0 : 0.0);
#pragma unroll
                for (int l = 0; l < D; ++l) acc -= ml_scale(bs.Rc[i * D + l], Kh[l * O + j]);
                Phi[i * D + j] = acc;
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                SD acc(0.0);
#pragma unroll
                for (int l = 0; l < D; ++l) acc += ml_scale(bs.Rc[i * D + l], Kh[l * O + D + e]);
                Bm[i * E + e] = acc;
            }
        }
        SD NU[O * O];          // N U, then U^T (N U)
#pragma unroll
        for (int k = 0; k < O; ++k)
#pragma unroll
            for (int o = 0; o < O; ++o) {
                SD acc(0.0);
#pragma unroll
                for (int c = 0; c < O; ++c) acc += ml_scale(bs.U[c * O + o], Nn[k * O + c]);
                NU[k * O + o] = acc;
            }
#pragma unroll
        for (int p = 0; p < O; ++p)
#pragma unroll
            for (int o = 0; o < O; ++o) {
                SD acc(0.0);
#pragma unroll
                for (int k = 0; k < O; ++k) acc += ml_scale(bs.U[k * O + p], NU[k * O + o]);
                if (p < D && o < D) N11[p * D + o] = acc;
                else if (p < D) N12[p * E + (o - D)] = acc;
                else if (o >= D) N22[(p - D) * E + (o - D)] = acc;
            }
        // joint innovation at T0, first block in the basis U
#pragma unroll
        for (int j = 0; j < D; ++j) e0h[j] = SD(0.0);
#pragma unroll
        for (int g = 0; g < O; ++g) {
            SD e((double)__ldg(yb + a.y.chan_off[g] + T0) - bs.ym[g]);
#pragma unroll
            for (int j = 0; j < D; ++j) e -= ml_scale(Cm[g * D + j], mu[j]);
#pragma unroll
            for (int j = 0; j < D; ++j) e0h[j] += ml_scale(bs.U[g * O + j], e);
        }
    }
    // G = sum_j (Phi^j)^T N11 Phi^j by doubling; Pm1 = Phi^lane; P32 = Phi^32; finally Pw = Phi^W
    SD G[D * D], Pw[D * D], Pm1[D * D], P32[D * D], tmp[D * D], tmp2[D * D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) { G[i] = N11[i]; Pw[i] = Phi[i]; Pm1[i] = SD((i / D == i % D) ? 1.0 : 0.0); P32[i] = SD(0.0); }
    for (int it = 0; it < la.nlog; ++it) {
        if (it == 5) {
#pragma unroll
            for (int i = 0; i < D * D; ++i) P32[i] = Pw[i];
        }
        if (it < 5) {
            ml_matmul<D>(Pm1, Pw, tmp);
            const bool take = (lane >> it) & 1;
#pragma unroll
            for (int i = 0; i < D * D; ++i) if (take) Pm1[i] = tmp[i];
        }
        ml_matmul<D>(G, Pw, tmp);
        ml_matmul_tn<D>(Pw, tmp, tmp2);
#pragma unroll
        for (int i = 0; i < D * D; ++i) G[i] += tmp2[i];
        ml_matmul<D>(Pw, Pw, tmp);
#pragma unroll
        for (int i = 0; i < D * D; ++i) Pw[i] = tmp[i];
    }
    double pmax = 0;
#pragma unroll
    for (int i = 0; i < D * D; ++i) pmax = fmax(pmax, fabs(Pw[i].v));
    if (!(pmax <= la.tolF)) return false;       // slow forgetting (or NaN): the truncated lag series is not exact

    const double* Rb = la.R + ((long long)lvl * a.B + b) * O * O * W;
    auto gvec = [&](int i, double* out) {
#pragma unroll
        for (int o = 0; o < O; ++o) out[o] = ml_g<P>(yb, a.y.chan_off, bs, O, n, i, o);
    };
    // Gam v = v[0:D] - Bm v[D:]
    auto gam = [&](const double* v, SD* out) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            SD acc(v[i]);
#pragma unroll
            for (int e = 0; e < E; ++e) acc -= ml_scale(v[D + e], Bm[i * E + e]);
            out[i] = acc;
        }
    };
    // <T, Gam R Gam^T> for a real O x O matrix R and a dual D x D matrix T
    auto quadA = [&](const SD* T, const double* Rm) {
        SD acc(0.0);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            SD Y1[O];
#pragma unroll
            for (int c = 0; c < O; ++c) {
                SD y(Rm[i * O + c]);
#pragma unroll
                for (int e = 0; e < E; ++e) y -= ml_scale(Rm[(D + e) * O + c], Bm[i * E + e]);
                Y1[c] = y;
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {
                SD x = Y1[j];
#pragma unroll
                for (int e = 0; e < E; ++e) x -= Y1[D + e] * Bm[j * E + e];
                acc += T[i * D + j] * x;
            }
        }
        return acc;
    };
    double gn[O];
    gvec(n, gn);
    SD FA(0.0), FB(0.0), HA[D], HB[D], tl[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { HA[i] = SD(0.0); HB[i] = SD(0.0); tl[i] = SD(0.0); }
    SD Pmm[D * D];                                  // Phi^m = Phi^(m-1) Phi
    for (int m = lane + 1; m < W; m += 32) {
        ml_matmul<D>(Pm1, Phi, Pmm);
        double Rm[O * O], RA[O * O], gh[O], gh1[O], gt[O], gnm[O];
#pragma unroll
        for (int x = 0; x < O * O; ++x) Rm[x] = Rb[(long long)x * W + m];
        gvec(T0 + m, gh); gvec(T0 + m + 1, gh1); gvec(n - 1 - m, gt); gvec(n - m, gnm);
#pragma unroll
        for (int x = 0; x < O; ++x)
#pragma unroll
            for (int c = 0; c < O; ++c) RA[x * O + c] = Rm[x * O + c] - gnm[x] * gn[c];
        ml_matmul_tn<D>(Pmm, G, tmp);                 // TA = (Phi^m)^T G
        FA += ml_scale(2.0, quadA(tmp, RA));
        SD v[D];
        gam(gh, v);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) HA[i] += tmp[i * D + j] * v[j];
        gam(gt, v);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) tl[i] += Pmm[i * D + j] * v[j];
        if (E > 0) {
            // term B: TB = (Phi^(m-1))^T N12; <TB, Gam Rg[:, D:]> and the coupling with the state at T0
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    SD tb(0.0);
#pragma unroll
                    for (int l = 0; l < D; ++l) tb += Pm1[l * D + i] * N12[l * E + e];
                    SD yb2(Rm[i * O + D + e]);
#pragma unroll
                    for (int e2 = 0; e2 < E; ++e2) yb2 -= ml_scale(Rm[(D + e2) * O + D + e], Bm[i * E + e2]);
                    FB += ml_scale(2.0, tb * yb2);
                }
            SD nw[D];                                // N12 w_{T0+m}
#pragma unroll
            for (int l = 0; l < D; ++l) {
                SD acc(0.0);
#pragma unroll
                for (int e = 0; e < E; ++e) acc += ml_scale(gh1[D + e], N12[l * E + e]);
                nw[l] = acc;
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int l = 0; l < D; ++l) HB[i] += Pmm[l * D + i] * nw[l];
        }
        if (m + 32 < W) {
            ml_matmul<D>(Pm1, P32, tmp2);
#pragma unroll
            for (int i = 0; i < D * D; ++i) Pm1[i] = tmp2[i];
        }
    }
    FA = ml_warp_sum(FA); FB = ml_warp_sum(FB);
#pragma unroll
    for (int i = 0; i < D; ++i) { HA[i] = ml_warp_sum(HA[i]); HB[i] = ml_warp_sum(HB[i]); tl[i] = ml_warp_sum(tl[i]); }
    // lag 0 (every lane, redundantly)
    SD Ct(0.0);
    {
        double R0[O * O], RA[O * O], gh1[O], gt[O];
#pragma unroll
        for (int x = 0; x < O * O; ++x) R0[x] = Rb[(long long)x * W];
        gvec(T0 + 1, gh1); gvec(n - 1, gt);
#pragma unroll
        for (int x = 0; x < O; ++x)
#pragma unroll
            for (int c = 0; c < O; ++c) RA[x * O + c] = R0[x * O + c] - gn[x] * gn[c];
        FA += quadA(G, RA);
        SD v[D];
        gam(gt, v);
#pragma unroll
        for (int i = 0; i < D; ++i) tl[i] += v[i];
#pragma unroll
        for (int l = 0; l < D; ++l)
#pragma unroll
            for (int e = 0; e < E; ++e) HB[l] += ml_scale(gh1[D + e], N12[l * E + e]);
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int e2 = 0; e2 < E; ++e2) Ct += ml_scale(R0[(D + e) * O + D + e2], N22[e * E + e2]);
    }
    SD E2 = FA + FB + Ct;
    {
        SD pt[D];      // Phi tl
#pragma unroll
        for (int x = 0; x < D; ++x) {
            SD acc(0.0);
#pragma unroll
            for (int c = 0; c < D; ++c) acc += Phi[x * D + c] * tl[c];
            pt[x] = acc;
        }
#pragma unroll
        for (int x = 0; x < D; ++x) {
            SD ge(0.0), gp(0.0);
#pragma unroll
            for (int c = 0; c < D; ++c) { ge += G[x * D + c] * e0h[c]; gp += G[x * D + c] * pt[c]; }
            E2 += e0h[x] * ge - pt[x] * gp + ml_scale(2.0, e0h[x] * (HA[x] + HB[x]));
        }
    }
    const double nB = (double)(n - T0);
    nll += ml_scale(nB, SD((double)O * HALF_LOG2PI) + ml_scale(0.5, lsum)) + ml_scale(0.5, E2);
    nll_out = nll.v;
    dnll_out = nll.d;
    return true;
}

template <class P, int OC>
__global__ void __launch_bounds__(32) lin_lag_opt_kernel(const __grid_constant__ GArgs<P> a,
                                                         const __grid_constant__ LinLagArgs la) {
    const int j = blockIdx.x, lane = threadIdx.x;
    AdamState<P> adam;                       // every lane carries the same state
    adam_init(adam, a.s_log0[j]);
    if (a.cap <= 0) {
        if (lane == 0) { a.s_log_out[j] = adam.s_log; a.last_loss_out[j] = adam.prev; a.iters_out[j] = 0; la.flag[j] = 0; }
        return;
    }
    const int m_lo = a.block_off[j], m_hi = a.block_off[j + 1];
    while (true) {
        P dsdlog;
        const P s = adam_current_s(adam, a.lo, a.hi, &dsdlog);
        P loss = P(0), grad = P(0);
        for (int mi = m_lo; mi < m_hi; ++mi) {      // members in order (eks/core.py:474-476)
            double nll, dnll;
            const bool ok = linlag_eval<P, OC>(a, la, a.members[mi], (double)s, lane, nll, dnll);
            if (!__all_sync(0xffffffffu, ok)) {
                if (lane == 0) la.flag[j] = 1;
                return;
            }
            P v = (P)nll, g = (P)dnll;
            if (!isfinite(nll) || !isfinite((double)v)) { v = P(1e12); g = P(0); }   // core.py:650
            loss += v;
            grad += g * dsdlog;
        }
        if (lane == 0 && a.trace && adam.iters < a.trace_cap) {
            P* tr = a.trace + ((long long)j * a.trace_cap + adam.iters) * 3;
            tr[0] = adam.s_log; tr[1] = loss; tr[2] = grad * a.lr;
        }
        adam_step(adam, loss, grad, a.lr, a.tol, a.cap);
        if (adam.done) {
            if (lane == 0) {
                a.s_log_out[j] = adam.s_log; a.last_loss_out[j] = adam.prev; a.iters_out[j] = adam.iters;
                la.flag[j] = 0;
            }
            return;
        }
    }
}

static int ml_W(int dtype) { return dtype == EKS_F32 ? 128 : 256; }
static size_t ml_align(size_t x) { return (x + 255) & ~(size_t)255; }

bool lin_lag_applicable(int dtype, int D, int O, int n_spans, int n) {
    if (getenv("EKS_NO_LINLAG")) return false;
    if (n_spans != 1 || n < ML_T0 + 4 * ml_W(dtype)) return false;
    return D == 3 && (O == 4 || O == 6 || O == 8);
}

size_t lin_lag_workspace_bytes(int dtype, int n_blocks, int B, int O, int T) {
    const int W = ml_W(dtype);
    const int nchunk = (T + ML_CH - 1) / ML_CH + 1;
    const int nx = (nchunk + ML_CPB - 1) / ML_CPB;
    return ml_align((size_t)ML_NT0 * B * O * O * W * sizeof(double)) + ml_align((size_t)B * O * O * nx * W * sizeof(double)) +
           ml_align((size_t)B * sizeof(MLBasis)) + ml_align((size_t)n_blocks * sizeof(int)) +
           ml_align((size_t)B * O * (T + 1) * (dtype == EKS_F32 ? 4 : 8)) + 256;
}

template <class P, int W>
static int lin_lag_run(const GArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st, int* used) {
    const int dtype = sizeof(P) == 4 ? EKS_F32 : EKS_F64;
    *used = 0;
    if (!workspace || workspace_bytes < lin_lag_workspace_bytes(dtype, a.n_blocks, a.B, a.O, a.T)) return 0;
    const int n = a.sp.total, O = a.O;
    unsigned char* w = (unsigned char*)workspace;
    double* R = (double*)w; w += ml_align((size_t)ML_NT0 * a.B * O * O * W * sizeof(double));
    MLagStatArgs<P> sa;
    sa.y = a.y; sa.B = a.B; sa.O = O; sa.t_begin = a.sp.start[0]; sa.n = n; sa.nsig = n + 1;
    sa.nchunk = (sa.nsig - (ML_T0 + 1) + ML_CH - 1) / ML_CH;
    sa.nx = (sa.nchunk + ML_CPB - 1) / ML_CPB;
    sa.partial = (double*)w; w += ml_align((size_t)a.B * O * O * sa.nx * W * sizeof(double));
    MLBasis* basis = (MLBasis*)w; w += ml_align((size_t)a.B * sizeof(MLBasis));
    sa.basis = basis;
    sa.R = R;
    int* flag = (int*)w; w += ml_align((size_t)a.n_blocks * sizeof(int));
    sa.gsig = (P*)w;
    constexpr int PADE = 16 / (int)sizeof(P);
    constexpr int NLOG = ML_CH + W;
    constexpr int NPHYS = NLOG + (NLOG / 16) * PADE + PADE;
    const int smem = 2 * NPHYS * (int)sizeof(P);
    cudaError_t e = cudaFuncSetAttribute(mlag_stats_kernel<P, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        set_error("mlag_stats_kernel: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    cudaMemsetAsync(flag, 0xFF, (size_t)a.n_blocks * sizeof(int), st);     // -1: not finished
    ml_basis_kernel<P><<<(a.B + 63) / 64, 64, 0, st>>>(a.B, O, a.C, a.ymean, basis);
    ml_signal_kernel<P><<<dim3((sa.nsig + 255) / 256, a.B * O), 256, 0, st>>>(sa);
    mlag_stats_kernel<P, W><<<dim3(sa.nx, a.B * O * O), ML_NT, smem, st>>>(sa);
    int rc = check_launch("mlag_stats_kernel");
    if (rc) return rc;
    const long long nred = (long long)a.B * O * O * W;
    mlag_reduce_kernel<P, W><<<(unsigned)((nred + 255) / 256), 256, 0, st>>>(sa);
    rc = check_launch("mlag_reduce_kernel");
    if (rc) return rc;
    LinLagArgs la;
    la.R = R; la.basis = basis; la.W = W; la.nlog = W == 128 ? 7 : 8; la.tolF = dtype == EKS_F32 ? 1e-7 : 1e-13;
    la.flag = flag;
    if (O == 4) lin_lag_opt_kernel<P, 4><<<a.n_blocks, 32, 0, st>>>(a, la);
    else if (O == 6) lin_lag_opt_kernel<P, 6><<<a.n_blocks, 32, 0, st>>>(a, la);
    else lin_lag_opt_kernel<P, 8><<<a.n_blocks, 32, 0, st>>>(a, la);
    rc = check_launch("lin_lag_opt_kernel");
    if (rc) return rc;
    // the run-parallel path is the fallback of flagged blocks: this needs the flags on the host (this entry point
    // synchronises the stream, as the run-parallel optimiser does between its chunks)
    std::vector<int> h(a.n_blocks);
    e = cudaMemcpyAsync(h.data(), flag, (size_t)a.n_blocks * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("lin_lag_optimize: %s", cudaGetErrorString(e));
        return (int)e;
    }
    bool all_ok = true;
    for (int v : h) all_ok = all_ok && (v == 0);
    if (getenv("EKS_DEBUG_RUNS")) {
        int nf = 0;
        for (int v : h) nf += (v != 0);
        fprintf(stderr, "[eks lin_lag] %d blocks, %d flagged for the run-parallel path\n", a.n_blocks, nf);
    }
    if (all_ok) {
        *used = 1;
        note_launches(5);
    }
    return 0;
}

template <class P>
int lin_lag_optimize(const GArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st, int* used) {
    if (sizeof(P) == 4) return lin_lag_run<P, 128>(a, workspace, workspace_bytes, st, used);
    return lin_lag_run<P, 256>(a, workspace, workspace_bytes, st, used);
}
template int lin_lag_optimize<float>(const GArgs<float>&, void*, size_t, cudaStream_t, int*);
template int lin_lag_optimize<double>(const GArgs<double>&, void*, size_t, cudaStream_t, int*);

}  // namespace eks
