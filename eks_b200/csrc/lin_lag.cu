// lin_lag.cu -- the smoothing-parameter optimiser of LINEAR models with A = I and a constant observation noise (the
// loss path of the multi-camera PCA-latent model: eks/core.py:562-699 on the model of eks/multicam_smoother.py:409-443)
// with ONE pass over the observations instead of one pass per Adam evaluation -- the matrix form of diag_lag.cu.
//
// With A = I the predicted mean obeys m_{t+1} = m_t + K e_t, so the joint innovation e_t = y_t - C m_t obeys
//       e_{t+1} = Psi e_t + d_{t+1},      d_t = y_t - y_{t-1},      Psi = I - C K      (O x O)
// once the covariance recursion has reached its fixed point (constant K, S).  The quadratic part of the NLL is then
//       sum_{t >= T0} e_t^T N e_t,   N = S^-1,
//   =   e0^T G e0 + 2 e0^T H + F - (Psi tl)^T G (Psi tl)
//       G  = sum_j (Psi^j)^T N Psi^j                        (discrete Lyapunov equation, by doubling)
//       F  = <G, R_0> + 2 sum_{m>=1} <(Psi^m)^T G, R_m>,    R_m[a][b] = sum_i d_i[a] d_{i+m}[b]   (lagged cross products)
//       H  = sum_{m>=1} (Psi^m)^T G d_{T0+m}                (coupling of the state at frame T0 with the first increments)
//       tl = sum_{m>=0} Psi^m d_{n-1-m}                     (innovation at the last frame: removes the tail of the sum)
// which depends on the data only through the O x O x W lag statistics R_m.  mlag_stats_kernel computes them in one
// streaming pass; lin_lag_opt_kernel then runs the WHOLE Adam loop of a block in one warp: per evaluation the exact
// filter (sequential scalar updates, covariance recursion with its s-sensitivity in forward duals) over the first T0
// frames, the steady-state algebra above in Dual<double>, the Adam step and the reference's stop rule.  The quantities
// K, N, Psi are assembled from the SAME sequential-scalar-update gains the run-parallel path uses (lin_cov_step in
// generic_runs.cu): with M[g][j] = h_g . k_j (j < g) and Wm = (I + M)^-1, K = [k_0 .. k_{O-1}] Wm and
// N = Wm^T diag(1/s_g) Wm.
//
// The closed form is used only when it is exact to rounding: A = I, one contiguous span, n >= T0 + 4 W, the covariance
// recursion converged within T0 = 256 frames, and |Psi^W| below 1e-7 (float32 mode, W = 128) / 1e-13 (float64 mode,
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

constexpr int ML_T0 = 256;      // statistics start after frame T0 (levels 256, 128, 64 as in diag_lag.cu)
constexpr int ML_NT0 = 3;
constexpr int ML_CH = 4096;     // increments per shared-memory tile
constexpr int ML_RM = 16;       // lags per thread
constexpr int ML_RP = 16;       // frames per thread and step
constexpr int ML_NT = 256;
constexpr int ML_CPB = 8;       // tiles per CTA

template <class P>
struct MLagStatArgs {
    PlaneView y;
    int B, O, t_begin, n, nchunk, nx;
    double* partial;   // [B O O][nx][W]
    double* R;         // [ML_NT0][B O O][W]
};

template <class P> struct MLVec;
template <> struct MLVec<float> { using type = float4; static constexpr int VW = 4; };
template <> struct MLVec<double> { using type = double2; static constexpr int VW = 2; };

template <class P>
__device__ __forceinline__ int ml_phys(int x) { return x + (x >> 4) * (16 / (int)sizeof(P)); }

// R_m[a][c] partial sums.  grid = (nx, B O O); CTA (x, (b, a, c)) handles the tiles x, x + nx, ... : increments of channel
// a (frames of the tile) and of channel c (tile + W halo) staged in shared memory; warp w owns lag groups w, w + 8, ...
// and lane l the frames (step * 32 + l) * 16 ... + 15: a 16 x 16 register tile of products per step.
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
    const P* ya = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[ca];
    const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[cc];
    double accd[GPW][ML_RM];
#pragma unroll
    for (int q = 0; q < GPW; ++q)
#pragma unroll
        for (int j = 0; j < ML_RM; ++j) accd[q][j] = 0.0;
    for (int chunk = blockIdx.x; chunk < a.nchunk; chunk += a.nx) {
        const int i0 = ML_T0 + 1 + chunk * ML_CH;
        __syncthreads();
        for (int x = threadIdx.x; x < NLOG; x += ML_NT) {
            const int i = i0 + x;
            P dc = P(0), da = P(0);
            if (i < a.n) {
                dc = __ldg(yc + i) - __ldg(yc + i - 1);
                if (x < ML_CH) da = __ldg(ya + i) - __ldg(ya + i - 1);
            }
            smC[ml_phys<P>(x)] = dc;
            if (x < ML_CH) smA[ml_phys<P>(x)] = da;
        }
        __syncthreads();
        const int nvalid = min(ML_CH, a.n - i0);
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

// fixed-order sum of the per-CTA partials, one thread per (sequence, a, c, lag); levels 128 / 64 add the few extra
// products of the increments between their start frame and frame 256.
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
    const P* ya = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[ca];
    const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[cc];
    int hi = ML_T0;
#pragma unroll
    for (int lvl = 1; lvl < ML_NT0; ++lvl) {
        const int lo = ML_T0 >> lvl;
        for (int i = hi; i > lo; --i)
            s += ((double)ya[i] - (double)ya[i - 1]) * ((double)yc[i + m] - (double)yc[i + m - 1]);
        a.R[lvl * per + idx] = s;
        hi = lo;
    }
}

// ---------------------------------------------------------------------------------------------------- the optimiser
using SD = Dual<double>;

template <int OC>
__device__ __forceinline__ void ml_matmul(const SD* __restrict__ A, const SD* __restrict__ B, SD* __restrict__ Cm) {
    // Cm = A B (OC x OC, row major)
#pragma unroll
    for (int i = 0; i < OC; ++i)
#pragma unroll
        for (int j = 0; j < OC; ++j) {
            SD acc(0.0);
#pragma unroll
            for (int k = 0; k < OC; ++k) acc += A[i * OC + k] * B[k * OC + j];
            Cm[i * OC + j] = acc;
        }
}

template <int OC>
__device__ __forceinline__ void ml_matmul_tn(const SD* __restrict__ A, const SD* __restrict__ B, SD* __restrict__ Cm) {
    // Cm = A^T B
#pragma unroll
    for (int i = 0; i < OC; ++i)
#pragma unroll
        for (int j = 0; j < OC; ++j) {
            SD acc(0.0);
#pragma unroll
            for (int k = 0; k < OC; ++k) acc += A[k * OC + i] * B[k * OC + j];
            Cm[i * OC + j] = acc;
        }
}

__device__ __forceinline__ SD ml_warp_sum(SD v) { return SD(warp_sum(v.v), warp_sum(v.d)); }

struct LinLagArgs {
    const double* R;       // [ML_NT0][B O O][W]
    int W, nlog;           // lags; log2(W)
    double tolF;
    int* flag;             // [n_blocks]: 1 = the closed form did not apply, run the run-parallel path
};

// One evaluation of one sequence by a whole warp (sequential parts redundantly on every lane, the lag series split over
// the lanes).  Returns false if the closed form does not apply.
template <class P, int DC, int OC>
__device__ bool linlag_eval(const GArgs<P>& a, const LinLagArgs& la, int b, double s_val, int lane, double& nll_out,
                            double& dnll_out) {
    constexpr int D = DC, O = OC;
    const int n = a.sp.total, t_begin = a.sp.start[0], W = la.W;
    const double HALF_LOG2PI = 0.91893853320467274178;
    const P* yb = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + t_begin;
    double Cm[O * D], Qm[D * D], rc[O], ym[O];
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
    for (int g = 0; g < O; ++g) {
        rc[g] = (double)a.Rconst[(long long)b * O + g];
        ym[g] = a.ymean ? (double)a.ymean[(long long)b * O + g] : 0.0;
    }
    const SD s(s_val, 1.0);
    SD Pm[D * D], mu[D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) Pm[i] = SD((double)a.S0[(long long)b * D * D + i]);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = SD((double)a.m0[(long long)b * D + i]);
    SD kg[O * D], isi[O], lsum(0.0), nll(0.0);
    bool bad = false;

    // one covariance step (sequential scalar updates on the predicted covariance, symmetrise, + s Q): gains -> kg, isi
    auto cov_step = [&]() {
        lsum = SD(0.0);
#pragma unroll
        for (int g = 0; g < O; ++g) {
            SD Ph[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                SD acc(0.0);
#pragma unroll
                for (int j = 0; j < D; ++j) acc += Pm[i * D + j] * SD(Cm[g * D + j]);
                Ph[i] = acc;
            }
            SD si(rc[g]);
#pragma unroll
            for (int j = 0; j < D; ++j) si += SD(Cm[g * D + j]) * Ph[j];
            if (!(si.v > 0) || !isfinite(si.v)) bad = true;
            const SD inv = SD(1.0) / si;
            lsum += log_(si);
            isi[g] = inv;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const SD k = Ph[i] * inv;
                kg[g * D + i] = k;
#pragma unroll
                for (int j = 0; j < D; ++j) Pm[i * D + j] -= k * Ph[j];
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i + 1; j < D; ++j) {
                const SD v = SD(0.5) * (Pm[i * D + j] + Pm[j * D + i]);
                Pm[i * D + j] = v; Pm[j * D + i] = v;
            }
#pragma unroll
        for (int i = 0; i < D * D; ++i) Pm[i] = Pm[i] + s * SD(Qm[i]);
    };
    // the observations of frame t through the current gains: NLL terms and the mean update
    auto obs_step = [&](int t) {
        SD q(0.0);
#pragma unroll
        for (int g = 0; g < O; ++g) {
            SD e((double)__ldg(yb + a.y.chan_off[g] + t) - ym[g]);
#pragma unroll
            for (int j = 0; j < D; ++j) e -= SD(Cm[g * D + j]) * mu[j];
            q += e * e * isi[g];
#pragma unroll
            for (int i = 0; i < D; ++i) mu[i] += kg[g * D + i] * e;
        }
        nll += SD((double)O * HALF_LOG2PI) + SD(0.5) * lsum + SD(0.5) * q;
    };

    // ---- head: exact filter; the covariance recursion is followed until its extrapolated distance to the fixed
    // point is below 1e-13 (relative) or it has reached its rounding floor
    const double tol = 1e-13;
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
        const double rv = fmin(cv / prev_cv, 0.999), rd = fmin(cd / prev_cd, 0.999);
        const bool cvok = (cv == 0.0) || (isfinite(prev_cv) && cv * rv / (1.0 - rv) <= tol * sv);
        const bool cdok = (cd == 0.0) || (isfinite(prev_cd) && cd * rd / (1.0 - rd) <= tol * sdv);
        if (cv >= prev_cv && cd >= prev_cd) ++stall;
        conv = (cvok && cdok) || stall >= 24;
        prev_cv = cv; prev_cd = cd;
    }
    if (!conv) return false;
    cov_step();                         // steady gains from the converged predicted covariance
    int lvl = ML_NT0 - 1;               // smallest statistics start frame >= the transient length
    while (lvl > 0 && (ML_T0 >> lvl) < t) --lvl;
    const int T0 = ML_T0 >> lvl;
    for (; t < T0; ++t) obs_step(t);
    if (bad) { nll_out = nan(""); dnll_out = 0.0; return true; }

    // ---- steady-state algebra
    SD Wm[O * O], Nn[O * O], Psi[O * O];
#pragma unroll
    for (int g = 0; g < O; ++g)
#pragma unroll
        for (int c = 0; c < O; ++c) {
            SD acc(g == c ? 1.0 : 0.0);
#pragma unroll
            for (int j = 0; j < O; ++j) {
                if (j < g) {
                    SD mgj(0.0);
#pragma unroll
                    for (int i = 0; i < D; ++i) mgj += SD(Cm[g * D + i]) * kg[j * D + i];
                    acc -= mgj * Wm[j * O + c];
                }
            }
            Wm[g * O + c] = acc;
        }
#pragma unroll
    for (int x = 0; x < O; ++x)
#pragma unroll
        for (int c = 0; c < O; ++c) {
            SD acc(0.0);
#pragma unroll
            for (int g = 0; g < O; ++g) acc += Wm[g * O + x] * isi[g] * Wm[g * O + c];
            Nn[x * O + c] = acc;
        }
    {
        SD Kj[D * O];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int c = 0; c < O; ++c) {
                SD acc(0.0);
#pragma unroll
                for (int g = 0; g < O; ++g) acc += kg[g * D + i] * Wm[g * O + c];
                Kj[i * O + c] = acc;
            }
#pragma unroll
        for (int x = 0; x < O; ++x)
#pragma unroll
            for (int c = 0; c < O; ++c) {
                SD acc(x == c ? 1.0 : 0.0);
#pragma unroll
                for (int i = 0; i < D; ++i) acc -= SD(Cm[x * D + i]) * Kj[i * O + c];
                Psi[x * O + c] = acc;
            }
    }
    // joint innovation at T0
    SD e0[O];
#pragma unroll
    for (int g = 0; g < O; ++g) {
        SD e((double)__ldg(yb + a.y.chan_off[g] + T0) - ym[g]);
#pragma unroll
        for (int j = 0; j < D; ++j) e -= SD(Cm[g * D + j]) * mu[j];
        e0[g] = e;
    }
    // G = sum_j (Psi^j)^T N Psi^j by doubling; Pl = Psi^lane; P32 = Psi^32; finally Pw = Psi^W
    SD G[O * O], Pw[O * O], Pl[O * O], P32[O * O], tmp[O * O], tmp2[O * O];
#pragma unroll
    for (int i = 0; i < O * O; ++i) { G[i] = Nn[i]; Pw[i] = Psi[i]; Pl[i] = SD((i / O == i % O) ? 1.0 : 0.0); }
    for (int it = 0; it < la.nlog; ++it) {
        if (it == 5) {
#pragma unroll
            for (int i = 0; i < O * O; ++i) P32[i] = Pw[i];
        }
        if (it < 5) {
            ml_matmul<O>(Pl, Pw, tmp);
            const bool take = (lane >> it) & 1;
#pragma unroll
            for (int i = 0; i < O * O; ++i) if (take) Pl[i] = tmp[i];
        }
        ml_matmul<O>(G, Pw, tmp);
        ml_matmul_tn<O>(Pw, tmp, tmp2);
#pragma unroll
        for (int i = 0; i < O * O; ++i) G[i] += tmp2[i];
        ml_matmul<O>(Pw, Pw, tmp);
#pragma unroll
        for (int i = 0; i < O * O; ++i) Pw[i] = tmp[i];
    }
    double pmax = 0;
#pragma unroll
    for (int i = 0; i < O * O; ++i) pmax = fmax(pmax, fabs(Pw[i].v));
    if (!(pmax <= la.tolF)) return false;       // slow forgetting (or NaN): the truncated lag series is not exact

    // ---- lag series, lane l takes the lags l, l + 32, ...
    const double* Rb = la.R + ((long long)lvl * a.B + b) * O * O * W;
    SD F(0.0), H[O], tl[O];
#pragma unroll
    for (int g = 0; g < O; ++g) { H[g] = SD(0.0); tl[g] = SD(0.0); }
    for (int m = lane; m < W; m += 32) {
        ml_matmul_tn<O>(Pl, G, tmp);                            // (Psi^m)^T G
        const double cm = m == 0 ? 1.0 : 2.0;
        SD f(0.0);
#pragma unroll
        for (int x = 0; x < O; ++x)
#pragma unroll
            for (int c = 0; c < O; ++c) f += tmp[x * O + c] * SD(Rb[((long long)x * O + c) * W + m]);
        F += SD(cm) * f;
        double dh[O], dt[O];
#pragma unroll
        for (int g = 0; g < O; ++g) {
            const P* yg = yb + a.y.chan_off[g];
            dh[g] = m >= 1 ? (double)__ldg(yg + T0 + m) - (double)__ldg(yg + T0 + m - 1) : 0.0;
            dt[g] = (double)__ldg(yg + n - 1 - m) - (double)__ldg(yg + n - 2 - m);
        }
#pragma unroll
        for (int x = 0; x < O; ++x) {
            SD hx(0.0), tx(0.0);
#pragma unroll
            for (int c = 0; c < O; ++c) { hx += tmp[x * O + c] * SD(dh[c]); tx += Pl[x * O + c] * SD(dt[c]); }
            H[x] += hx; tl[x] += tx;
        }
        if (m + 32 < W) {
            ml_matmul<O>(Pl, P32, tmp2);
#pragma unroll
            for (int i = 0; i < O * O; ++i) Pl[i] = tmp2[i];
        }
    }
    F = ml_warp_sum(F);
#pragma unroll
    for (int g = 0; g < O; ++g) { H[g] = ml_warp_sum(H[g]); tl[g] = ml_warp_sum(tl[g]); }
    SD E2(0.0);
    {
        SD pt[O];      // Psi tl
#pragma unroll
        for (int x = 0; x < O; ++x) {
            SD acc(0.0);
#pragma unroll
            for (int c = 0; c < O; ++c) acc += Psi[x * O + c] * tl[c];
            pt[x] = acc;
        }
#pragma unroll
        for (int x = 0; x < O; ++x) {
            SD ge(0.0), gp(0.0);
#pragma unroll
            for (int c = 0; c < O; ++c) { ge += G[x * O + c] * e0[c]; gp += G[x * O + c] * pt[c]; }
            E2 += e0[x] * ge - pt[x] * gp + SD(2.0) * e0[x] * H[x];
        }
        E2 += F;
    }
    const double nB = (double)(n - T0);
    nll += SD(nB) * (SD((double)O * HALF_LOG2PI) + SD(0.5) * lsum) + SD(0.5) * E2;
    nll_out = nll.v;
    dnll_out = nll.d;
    return true;
}

template <class P, int DC, int OC>
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
            const bool ok = linlag_eval<P, DC, OC>(a, la, a.members[mi], (double)s, lane, nll, dnll);
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
           ml_align((size_t)n_blocks * sizeof(int)) + 256;
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
    sa.y = a.y; sa.B = a.B; sa.O = O; sa.t_begin = a.sp.start[0]; sa.n = n;
    sa.nchunk = (n - (ML_T0 + 1) + ML_CH - 1) / ML_CH;
    sa.nx = (sa.nchunk + ML_CPB - 1) / ML_CPB;
    sa.partial = (double*)w; w += ml_align((size_t)a.B * O * O * sa.nx * W * sizeof(double));
    sa.R = R;
    int* flag = (int*)w;
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
    mlag_stats_kernel<P, W><<<dim3(sa.nx, a.B * O * O), ML_NT, smem, st>>>(sa);
    int rc = check_launch("mlag_stats_kernel");
    if (rc) return rc;
    const long long nred = (long long)a.B * O * O * W;
    mlag_reduce_kernel<P, W><<<(unsigned)((nred + 255) / 256), 256, 0, st>>>(sa);
    rc = check_launch("mlag_reduce_kernel");
    if (rc) return rc;
    LinLagArgs la;
    la.R = R; la.W = W; la.nlog = W == 128 ? 7 : 8; la.tolF = dtype == EKS_F32 ? 1e-7 : 1e-13; la.flag = flag;
    if (O == 4) lin_lag_opt_kernel<P, 3, 4><<<a.n_blocks, 32, 0, st>>>(a, la);
    else if (O == 6) lin_lag_opt_kernel<P, 3, 6><<<a.n_blocks, 32, 0, st>>>(a, la);
    else lin_lag_opt_kernel<P, 3, 8><<<a.n_blocks, 32, 0, st>>>(a, la);
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
    if (all_ok) {
        *used = 1;
        note_launches(3);
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
