// diag.cu -- time-parallel kernels for DECOUPLED models: D == O == 2 with diagonal A, C, Q, S0 (the
// single-camera EKS model, eks/singlecam_smoother.py:246-284).  With a diagonal R the 2-D filter is
// two independent scalar filters that share the smoothing parameter s.
//
// diag_optimize_kernel: the whole Adam loop of _vmap_optimize_singletons / the block path
// (eks/core.py:562-699, 403-559) for one block per CTA, persistent over all iterations:
//   * loss path has CONSTANT R (core.py:702-709) => the variance recursion is data independent.
//     Lane c of warp 0 runs it sequentially (with its s-sensitivity) until it reaches its floating
//     point fixed point ("transient", a few tens of frames), accumulating the NLL there exactly as
//     the sequential filter does;
//   * for the remaining frames the gain is constant and the predicted mean m_t and its sensitivity
//     dm_t/ds obey a constant-coefficient linear recurrence
//         z_{t+1} = Phi z_t + b y_t,  z = (m, dm),  Phi = [[alpha,0],[dalpha,alpha]], b = (beta,dbeta)
//     which is evaluated EXACTLY in parallel: each thread owns L consecutive frames in registers,
//     computes its zero-state response (2 FMAs / frame), a block-wide scan with the closed-form
//     powers Phi^(L 2^k) hands every thread its carry-in, and a second register pass accumulates
//     sum e^2 and sum e dm (7 FMAs / frame).  y is read from HBM/L2 once per Adam iteration.
//   * NLL = n/2 log 2pi + 1/2 sum log S_t + 1/2 sum e_t^2 / S_t, its derivative likewise; partial
//     sums are kept in fp64 so that the reference's relative-tolerance stop rule stays meaningful
//     at 10^6 frames even in float32 mode.
#include "common.cuh"
#include "ekf_generic.cuh"
#include "diag.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int DIAG_NT = 256;
constexpr int DIAG_NW = DIAG_NT / 32;

template <class P> struct DiagTraits;
template <> struct DiagTraits<float> {
    static constexpr int L = 16;
    using vec_t = float4;
    static constexpr int VW = 4;
    __device__ static float eps() { return 1.1920929e-7f; }
};
template <> struct DiagTraits<double> {
    static constexpr int L = 8;
    using vec_t = double2;
    static constexpr int VW = 2;
    __device__ static double eps() { return 2.220446049250313e-16; }
};

template <class P>
struct ChanConst {
    P alpha, beta, a, cc, dalpha, dbeta, iS, diS, logS, dlogS;
    P aL[5], bL[5];  // Phi^(L 2^k) = [[aL,0],[bL,aL]]
    P aW, bW;        // Phi^(32 L)
};

template <class P>
struct DiagShared {
    ChanConst<P> ch[2];
    P z_tile[2][2][2];        // [buf][chan][m,dm]
    P agg[2][DIAG_NW][2][2];  // [buf][warp][chan][m,dm]
    double tsum[2][5];        // transient sums per channel: logS, dlogS, e2 iS, e2 diS, cc e dm iS
    double red[DIAG_NW][4];
    int t_c;
    int done;
    P s, dsdlog;
    double loss_acc, grad_acc;
    AdamState<P> adam;
};

template <class P>
struct DiagOptArgs {
    int B, t_begin, n;
    const P *m0, *S0, *A, *Q, *C;
    PlaneView y;
    const P *ymean, *Rconst;
    int n_blocks;
    const int *block_off, *members;
    const P* s_log0;
    P lr, lo, hi, tol;
    int cap;
    P *s_log_out, *last_loss_out;
    int* iters_out;
    P* trace;
    int trace_cap;
};

// ---- transient: sequential scalar filter with s-sensitivities until the variance recursion has
// converged (or the sequence ends).  Executed by lane c (< 2) of warp 0; all 32 lanes vote.
template <class P>
__device__ void diag_transient(const DiagOptArgs<P>& a, int b, P s, DiagShared<P>& sh) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < 2;
    const int c = act ? lane : 0;
    const P av = a.A[(long long)b * 4 + c * 3], cc = a.C[(long long)b * 4 + c * 3], Qc = a.Q[(long long)b * 4 + c * 3];
    const P r = a.Rconst[(long long)b * 2 + c];
    const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
    const P* yp = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.y.chan_off[c] + a.t_begin;
    P Pv = a.S0[(long long)b * 4 + c * 3], dP = P(0), m = a.m0[(long long)b * 2 + c], dm = P(0);
    double sl = 0, sdl = 0, se = 0, sde = 0, sg = 0;
    const P tol = P(8) * DiagTraits<P>::eps();
    P prevdP_chg = P(INFINITY), prevP_chg = P(INFINITY);
    int stall = 0;
    int t = 0;
    const int n = a.n;
    P S, iS, dS, diS, K, dK;
    while (true) {
        S = cc * cc * Pv + r;
        iS = P(1) / S;
        dS = cc * cc * dP;
        diS = -dS * iS * iS;
        K = Pv * cc / (S + P(1e-9));
        dK = cc * (dP * iS + Pv * diS);
        const P Pf = Pv - K * K * S;
        const P dPf = r * (dP * iS + Pv * diS);
        const P Pn = av * av * Pf + s * Qc;
        const P dPn = av * av * dPf + Qc;
        // convergence of (P, dP): relative step below tol * (1 - rho), rho = alpha^2 the contraction
        // factor, or the iteration has hit its rounding floor (steps no longer shrinking)
        const P alpha = av * (P(1) - K * cc);
        const P gap = P(1) - alpha * alpha;
        const P chgP = fabs(Pn - Pv), chgd = fabs(dPn - dP);
        bool conv = (chgP <= tol * gap * fabs(Pn)) && (chgd <= tol * gap * fabs(dPn));
        if (chgP >= prevP_chg && chgd >= prevdP_chg) ++stall;
        if (stall >= 24) conv = true;
        prevP_chg = chgP;
        prevdP_chg = chgd;
        const unsigned all_conv = __all_sync(0xffffffffu, conv || !act);
        if ((all_conv && (t & 3) == 0) || t >= n) break;
        if (act) {
            const P y = yp[t] - mean;
            const P e = y - cc * m;
            sl += (double)log_(S);
            sdl += (double)(dS * iS);
            se += (double)(e * e * iS);
            sde += (double)(e * e * diS);
            sg += (double)(cc * e * dm * iS);
            const P mf = m + K * e;
            const P dmf = dm + dK * e - K * cc * dm;
            m = av * mf;
            dm = av * dmf;
            Pv = Pn;
            dP = dPn;
        }
        ++t;
    }
    if (act) {
        ChanConst<P>& k = sh.ch[c];
        k.a = av; k.cc = cc;
        k.alpha = av * (P(1) - K * cc);
        k.beta = av * K;
        k.dalpha = -av * cc * dK;
        k.dbeta = av * dK;
        k.iS = iS; k.diS = diS;
        k.logS = log_(S);
        k.dlogS = dS * iS;
        constexpr int L = DiagTraits<P>::L;
        P aL = pow_(k.alpha, P(L));
        P bL = P(L) * pow_(k.alpha, P(L - 1)) * k.dalpha;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            k.aL[i] = aL; k.bL[i] = bL;
            bL = P(2) * aL * bL;
            aL = aL * aL;
        }
        k.aW = aL; k.bW = bL;
        sh.z_tile[0][c][0] = m;
        sh.z_tile[0][c][1] = dm;
        sh.tsum[c][0] = sl; sh.tsum[c][1] = sdl; sh.tsum[c][2] = se; sh.tsum[c][3] = sde; sh.tsum[c][4] = sg;
        if (c == 0) sh.t_c = t;
    }
}

template <class P, int L>
__device__ inline void load_chunk(const P* __restrict__ p, bool vec, int nvalid, P mean, P (&out)[L]) {
    using V = typename DiagTraits<P>::vec_t;
    constexpr int VW = DiagTraits<P>::VW;
    if (vec && nvalid == L) {
        const V* pv = reinterpret_cast<const V*>(p);
#pragma unroll
        for (int i = 0; i < L / VW; ++i) {
            const V v = __ldg(pv + i);
            const P* e = reinterpret_cast<const P*>(&v);
#pragma unroll
            for (int q = 0; q < VW; ++q) out[i * VW + q] = e[q] - mean;
        }
    } else {
#pragma unroll
        for (int i = 0; i < L; ++i) out[i] = (i < nvalid) ? (__ldg(p + i) - mean) : P(0);
    }
}

// One tile of DIAG_NT * L frames for both channels.  E2/G are this thread's fp64 accumulators.
template <class P, bool FULL>
__device__ inline void diag_tile(const P* __restrict__ y0, const P* __restrict__ y1, P mean0, P mean1, bool vec,
                                 int t0, int n, int buf, DiagShared<P>& sh, const P (&a_lane)[2],
                                 const P (&b_lane)[2], double (&E2)[2], double (&G)[2]) {
    constexpr int L = DiagTraits<P>::L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int start = t0 + threadIdx.x * L;
    const int nvalid = FULL ? L : max(0, min(L, n - start));
    P y[2][L];
    load_chunk<P, L>(y0 + start, vec, nvalid, mean0, y[0]);
    load_chunk<P, L>(y1 + start, vec, nvalid, mean1, y[1]);
    P zm[2], zd[2];
    // phase 1: zero-state response of the chunk.  U = sum alpha^(L-1-i) y_i, W = dU/dalpha
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const P alpha = sh.ch[c].alpha;
        P U = P(0), W = P(0);
#pragma unroll
        for (int i = 0; i < L; ++i) {
            W = fma(alpha, W, U);
            U = fma(alpha, U, y[c][i]);
        }
        zm[c] = sh.ch[c].beta * U;
        zd[c] = sh.ch[c].dbeta * U + sh.ch[c].beta * sh.ch[c].dalpha * W;
    }
    // warp inclusive scan with the closed-form powers of Phi
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int d = 1 << k;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const P pm = __shfl_up_sync(0xffffffffu, zm[c], d);
            const P pd = __shfl_up_sync(0xffffffffu, zd[c], d);
            if (lane >= d) {
                zd[c] = fma(sh.ch[c].aL[k], pd, fma(sh.ch[c].bL[k], pm, zd[c]));
                zm[c] = fma(sh.ch[c].aL[k], pm, zm[c]);
            }
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < 2; ++c) { sh.agg[buf][warp][c][0] = zm[c]; sh.agg[buf][warp][c][1] = zd[c]; }
    }
    P em[2], ed[2];  // exclusive prefix inside the warp
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        em[c] = __shfl_up_sync(0xffffffffu, zm[c], 1);
        ed[c] = __shfl_up_sync(0xffffffffu, zd[c], 1);
        if (lane == 0) { em[c] = P(0); ed[c] = P(0); }
    }
    __syncthreads();
    // carry at the start of this warp: tile carry pushed through the preceding warps' aggregates
    P m_in[2], d_in[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        P cm = sh.z_tile[buf][c][0], cd = sh.z_tile[buf][c][1];
        const P aW = sh.ch[c].aW, bW = sh.ch[c].bW;
        for (int w = 0; w < warp; ++w) {
            const P nm = fma(aW, cm, sh.agg[buf][w][c][0]);
            cd = fma(aW, cd, fma(bW, cm, sh.agg[buf][w][c][1]));
            cm = nm;
        }
        m_in[c] = fma(a_lane[c], cm, em[c]);
        d_in[c] = fma(a_lane[c], cd, fma(b_lane[c], cm, ed[c]));
        if (warp == DIAG_NW - 1 && lane == 0) {  // carry for the next tile (other buffer)
            const P nm = fma(aW, cm, sh.agg[buf][warp][c][0]);
            const P nd = fma(aW, cd, fma(bW, cm, sh.agg[buf][warp][c][1]));
            sh.z_tile[buf ^ 1][c][0] = nm;
            sh.z_tile[buf ^ 1][c][1] = nd;
        }
    }
    // phase 3: innovations with the true carry-in; e = y - cc m, m' = a m + beta e, dm' = alpha dm + dbeta e
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const P alpha = sh.ch[c].alpha, beta = sh.ch[c].beta, dbeta = sh.ch[c].dbeta, av = sh.ch[c].a,
                cc = sh.ch[c].cc;
        P m = m_in[c], dm = d_in[c], e2 = P(0), g = P(0);
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const P e = fma(-cc, m, y[c][i]);
            if (FULL || i < nvalid) {
                e2 = fma(e, e, e2);
                g = fma(e, dm, g);
            }
            dm = fma(alpha, dm, dbeta * e);
            m = fma(beta, e, av * m);
        }
        E2[c] += (double)e2;
        G[c] += (double)g;
    }
}

template <class P>
__global__ void __launch_bounds__(DIAG_NT, 2) diag_optimize_kernel(const __grid_constant__ DiagOptArgs<P> a) {
    __shared__ DiagShared<P> sh;
    constexpr int L = DiagTraits<P>::L;
    constexpr int TILE = DIAG_NT * L;
    const int j = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        adam_init(sh.adam, a.s_log0[j]);
        sh.done = (a.cap <= 0);
    }
    __syncthreads();
    const double HALF_LOG2PI = 0.91893853320467274178;
    while (true) {
        if (threadIdx.x == 0 && !sh.done) {
            P dsdlog;
            sh.s = adam_current_s(sh.adam, a.lo, a.hi, &dsdlog);
            sh.dsdlog = dsdlog;
            sh.loss_acc = 0.0;
            sh.grad_acc = 0.0;
        }
        __syncthreads();
        if (sh.done) break;
        const P s = sh.s;
        for (int mi = a.block_off[j]; mi < a.block_off[j + 1]; ++mi) {
            const int b = a.members[mi];
            if (warp == 0) diag_transient<P>(a, b, s, sh);
            __syncthreads();
            const int t_c = sh.t_c;
            const P* ybase = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin;
            const P* y0 = ybase + a.y.chan_off[0];
            const P* y1 = ybase + a.y.chan_off[1];
            const P mean0 = a.ymean ? a.ymean[(long long)b * 2] : P(0);
            const P mean1 = a.ymean ? a.ymean[(long long)b * 2 + 1] : P(0);
            const bool vec = ((reinterpret_cast<uintptr_t>(y0 + t_c) | reinterpret_cast<uintptr_t>(y1 + t_c)) & 15) == 0;
            // per-thread powers Phi^(L lane) for folding the warp carry into the exclusive prefix
            P a_lane[2], b_lane[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const P alpha = sh.ch[c].alpha;
                const P nl = P(L * lane);
                a_lane[c] = lane == 0 ? P(1) : pow_(alpha, nl);
                b_lane[c] = lane == 0 ? P(0) : nl * pow_(alpha, nl - P(1)) * sh.ch[c].dalpha;
            }
            double E2[2] = {0, 0}, G[2] = {0, 0};
            int buf = 0;
            for (int t0 = t_c; t0 < a.n; t0 += TILE, buf ^= 1) {
                if (t0 + TILE <= a.n) diag_tile<P, true>(y0, y1, mean0, mean1, vec, t0, a.n, buf, sh, a_lane, b_lane, E2, G);
                else diag_tile<P, false>(y0, y1, mean0, mean1, vec, t0, a.n, buf, sh, a_lane, b_lane, E2, G);
            }
            // deterministic block reduction of the four fp64 sums
            double v4[4] = {E2[0], E2[1], G[0], G[1]};
#pragma unroll
            for (int q = 0; q < 4; ++q) v4[q] = warp_sum(v4[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) sh.red[warp][q] = v4[q];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot[4] = {0, 0, 0, 0};
                for (int w = 0; w < DIAG_NW; ++w)
                    for (int q = 0; q < 4; ++q) tot[q] += sh.red[w][q];
                const double nB = (double)(a.n - t_c);
                double nll = 0, dnll = 0;
                for (int c = 0; c < 2; ++c) {
                    const ChanConst<P>& k = sh.ch[c];
                    nll += (double)a.n * HALF_LOG2PI + 0.5 * sh.tsum[c][0] + 0.5 * sh.tsum[c][2] +
                           0.5 * nB * (double)k.logS + 0.5 * (double)k.iS * tot[c];
                    dnll += 0.5 * sh.tsum[c][1] + 0.5 * sh.tsum[c][3] - sh.tsum[c][4] + 0.5 * nB * (double)k.dlogS +
                            0.5 * (double)k.diS * tot[c] - (double)k.cc * (double)k.iS * tot[2 + c];
                }
                P v = (P)nll, g = (P)dnll;
                if (!isfinite(nll) || !isfinite((double)v)) { v = P(1e12); g = P(0); }  // core.py:650
                sh.loss_acc += (double)v;
                sh.grad_acc += (double)(g * sh.dsdlog);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const P loss = (P)sh.loss_acc, g = (P)sh.grad_acc;
            if (a.trace && sh.adam.iters < a.trace_cap) {
                P* tr = a.trace + ((long long)j * a.trace_cap + sh.adam.iters) * 3;
                tr[0] = sh.adam.s_log; tr[1] = loss; tr[2] = g * a.lr;
            }
            adam_step(sh.adam, loss, g, a.lr, a.tol, a.cap);
            sh.done = sh.adam.done;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.s_log_out[j] = sh.adam.s_log;
        a.last_loss_out[j] = sh.adam.prev;
        a.iters_out[j] = sh.adam.iters;
    }
}

size_t diag_optimize_workspace_bytes(int dtype, int n_blocks) {
    (void)dtype; (void)n_blocks;
    return 16;
}

template <class P>
static int diag_optimize_launch(const DiagOptArgs<P>& a, cudaStream_t st) {
    diag_optimize_kernel<P><<<a.n_blocks, DIAG_NT, 0, st>>>(a);
    return check_launch("diag_optimize_kernel");
}

int diag_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q, const void* C,
                  const void* y_base, long long y_seq_stride, const long long* y_off, const void* ymean,
                  const void* Rconst, int t_begin, int n, int n_blocks, const int* block_off, const int* members,
                  const void* s_log0, double lr, double lo, double hi, double tol, int cap, void* s_log_out,
                  void* last_loss_out, int* iters_out, void* trace, int trace_cap, cudaStream_t st) {
    (void)T;
#define EKS_FILL(PT)                                                                                        \
    DiagOptArgs<PT> a;                                                                                      \
    a.B = B; a.t_begin = t_begin; a.n = n;                                                                  \
    a.m0 = (const PT*)m0; a.S0 = (const PT*)S0; a.A = (const PT*)A; a.Q = (const PT*)Q; a.C = (const PT*)C; \
    a.y.base = y_base; a.y.seq_stride = y_seq_stride;                                                       \
    for (int i = 0; i < MAX_CHAN; ++i) a.y.chan_off[i] = i < 2 ? y_off[i] : 0;                              \
    a.ymean = (const PT*)ymean; a.Rconst = (const PT*)Rconst;                                               \
    a.n_blocks = n_blocks; a.block_off = block_off; a.members = members; a.s_log0 = (const PT*)s_log0;      \
    a.lr = (PT)lr; a.lo = (PT)lo; a.hi = (PT)hi; a.tol = (PT)tol; a.cap = cap;                              \
    a.s_log_out = (PT*)s_log_out; a.last_loss_out = (PT*)last_loss_out; a.iters_out = iters_out;            \
    a.trace = (PT*)trace; a.trace_cap = trace_cap;                                                          \
    return diag_optimize_launch<PT>(a, st);
    if (dtype == EKS_F32) { EKS_FILL(float) }
    EKS_FILL(double)
#undef EKS_FILL
}

// =====================================================================================================
// Final pass for decoupled models: forward filter + RTS smoother with TIME-VARYING diagonal R_t
// (replaces vmap(_smooth_one) / extended_kalman_smoother, eks/core.py:274-295, for the singlecam model)
// fused with the reprojection epilogue of eks/singlecam_smoother.py:189-217 (x = C m + mean,
// posterior variance = C V C^T): results land directly in the output planes.
//
// One CTA per (sequence, channel) scalar problem, tiles of DIAG_NT*L frames, every thread owns L
// consecutive frames in registers.
//   forward : the predicted-variance recursion P' = a^2 P r/(c^2 P + r) + q is a Moebius map of P, so
//             chunk products of 2x2 matrices are scanned across the block (Sarkka & Garcia-Fernandez
//             style temporal parallelisation, scalar case) to give every thread its exact P at chunk
//             start; the thread then runs the ordinary per-frame filter arithmetic (gain with the 1e-9
//             boost, P_f = P - K S K) while composing the affine map of the mean, a second scan delivers
//             the carry-in mean, and a register pass writes the filtered moments.
//   backward: m_s[t] = G_t m_s[t+1] + (1 - G_t a) m_f[t],  P_s[t] = G_t^2 P_s[t+1] + (P_f[t] - G_t^2 S_p)
//             are affine recurrences with known coefficients -> one scan in reversed thread order.
// =====================================================================================================
template <class P>
struct DiagSmoothArgs {
    int B, T;
    const P *m0, *S0, *A, *Q, *C;
    PlaneView y, var;
    const P* ymean;
    const P* s;
    P* mf;  // workspace planes [B][2][T]
    P* Pf;
    P* out;
    long long out_seq_stride;
    long long out_off[4];  // x plane ch0, ch1 ; posterior-variance plane ch0, ch1
};

template <class P> __device__ inline P pow2_scale(P sum);
template <> __device__ inline float pow2_scale<float>(float sum) {
    // 2^-e with e the unbiased exponent of sum (sum > 0, finite): keeps products in range
    const int E = (__float_as_int(sum) >> 23) & 0xff;
    return __int_as_float((254 - E) << 23);
}
template <> __device__ inline double pow2_scale<double>(double sum) {
    const int E = (__double2hiint(sum) >> 20) & 0x7ff;
    return __hiloint2double((2046 - E) << 20, 0);
}

template <class P>
struct Mob { P a, b, c, d; };  // [[a,b],[c,d]] acting on P: (a P + b) / (c P + d)

template <class P>
__device__ inline Mob<P> mob_mul(const Mob<P>& l, const Mob<P>& r) {  // l applied after r
    Mob<P> o;
    o.a = fma(l.a, r.a, l.b * r.c);
    o.b = fma(l.a, r.b, l.b * r.d);
    o.c = fma(l.c, r.a, l.d * r.c);
    o.d = fma(l.c, r.b, l.d * r.d);
    return o;
}
template <class P>
__device__ inline void mob_norm(Mob<P>& m) {
    const P sc = pow2_scale<P>(m.a + m.b + m.c + m.d);
    m.a *= sc; m.b *= sc; m.c *= sc; m.d *= sc;
}
template <class P>
__device__ inline Mob<P> mob_shfl_up(const Mob<P>& m, int d) {
    Mob<P> o;
    o.a = __shfl_up_sync(0xffffffffu, m.a, d);
    o.b = __shfl_up_sync(0xffffffffu, m.b, d);
    o.c = __shfl_up_sync(0xffffffffu, m.c, d);
    o.d = __shfl_up_sync(0xffffffffu, m.d, d);
    return o;
}

template <class P, int L>
__device__ inline void store_chunk(P* __restrict__ p, bool vec, int nvalid, const P (&v)[L]) {
    using V = typename DiagTraits<P>::vec_t;
    constexpr int VW = DiagTraits<P>::VW;
    if (vec && nvalid == L) {
        V* pv = reinterpret_cast<V*>(p);
#pragma unroll
        for (int i = 0; i < L / VW; ++i) {
            V t;
            P* e = reinterpret_cast<P*>(&t);
#pragma unroll
            for (int q = 0; q < VW; ++q) e[q] = v[i * VW + q];
            pv[i] = t;
        }
    } else {
#pragma unroll
        for (int i = 0; i < L; ++i)
            if (i < nvalid) p[i] = v[i];
    }
}

template <class P>
struct FwdShared {
    Mob<P> magg[2][DIAG_NW];
    P aagg[2][DIAG_NW][2];
    P carry[2][2];  // [buf][P, m] predicted state at tile start
};

template <class P>
__global__ void __launch_bounds__(DIAG_NT, 2) diag_filter_kernel(const __grid_constant__ DiagSmoothArgs<P> a) {
    __shared__ FwdShared<P> sh;
    constexpr int L = DiagTraits<P>::L;
    constexpr int TILE = DIAG_NT * L;
    const int b = blockIdx.x >> 1, c = blockIdx.x & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const P av = a.A[(long long)b * 4 + c * 3], cc = a.C[(long long)b * 4 + c * 3];
    const P q = a.s[b] * a.Q[(long long)b * 4 + c * 3];
    const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
    const P* yp = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.y.chan_off[c];
    const P* vp = reinterpret_cast<const P*>(a.var.base) + (long long)b * a.var.seq_stride + a.var.chan_off[c];
    P* mfp = a.mf + ((long long)b * 2 + c) * a.T;
    P* Pfp = a.Pf + ((long long)b * 2 + c) * a.T;
    const bool vec = (((reinterpret_cast<uintptr_t>(yp) | reinterpret_cast<uintptr_t>(vp) |
                        reinterpret_cast<uintptr_t>(mfp) | reinterpret_cast<uintptr_t>(Pfp)) & 15) == 0);
    if (threadIdx.x == 0) {
        sh.carry[0][0] = a.S0[(long long)b * 4 + c * 3];
        sh.carry[0][1] = a.m0[(long long)b * 2 + c];
    }
    const P a2 = av * av, c2 = cc * cc, qc2 = q * c2;
    int buf = 0;
    for (int t0 = 0; t0 < a.T; t0 += TILE, buf ^= 1) {
        const int start = t0 + threadIdx.x * L;
        const int nvalid = max(0, min(L, a.T - start));
        P y[L], r[L], Pf[L];
        load_chunk<P, L>(yp + start, vec, nvalid, mean, y);
        load_chunk<P, L>(vp + start, vec, nvalid, P(0), r);
#pragma unroll
        for (int i = 0; i < L; ++i) {
            if (i >= nvalid) r[i] = P(1);
            else if (r[i] < P(1e-12)) r[i] = P(1e-12);  // np.clip(ev, 1e-12, None); NaN passes through
        }
        // ---- phase 1a: chunk Moebius product, each factor pre-divided by r_i
        Mob<P> M{P(1), P(0), P(0), P(1)};
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const P ir = P(1) / r[i];
            const Mob<P> Mi{fma(qc2, ir, a2), q, c2 * ir, P(1)};
            M = mob_mul(Mi, M);
            if (i & 1) mob_norm(M);
        }
        // inclusive scan over the warp (later chunk multiplies from the left)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int d = 1 << k;
            const Mob<P> prev = mob_shfl_up(M, d);
            if (lane >= d) { M = mob_mul(M, prev); mob_norm(M); }
        }
        if (lane == 31) sh.magg[buf][warp] = M;
        Mob<P> ex = mob_shfl_up(M, 1);
        if (lane == 0) ex = Mob<P>{P(1), P(0), P(0), P(1)};
        __syncthreads();
        Mob<P> cm{P(1), P(0), P(0), P(1)};
        for (int w = 0; w < warp; ++w) { cm = mob_mul(sh.magg[buf][w], cm); mob_norm(cm); }
        const P P_tile = sh.carry[buf][0], m_tile = sh.carry[buf][1];
        const Mob<P> tot = mob_mul(ex, cm);
        P Pv = (tot.a * P_tile + tot.b) / (tot.c * P_tile + tot.d);
        if (warp == DIAG_NW - 1 && lane == 31) {  // predicted variance at the start of the next tile
            const Mob<P> all = mob_mul(M, cm);
            sh.carry[buf ^ 1][0] = (all.a * P_tile + all.b) / (all.c * P_tile + all.d);
        }
        // ---- phase 1b: exact per-frame filter arithmetic from the chunk's true P; affine map of m
        P Aacc = P(1), bacc = P(0);
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const P S = fma(c2, Pv, r[i]);
            const P K = Pv * cc / (S + P(1e-9));
            const P Pfi = Pv - K * K * S;
            const P alpha = av * (P(1) - K * cc), beta = av * K;
            bacc = fma(alpha, bacc, beta * y[i]);
            Aacc *= alpha;
            r[i] = K;      // r is dead from here on: reuse its registers for the gain
            Pf[i] = Pfi;
            Pv = fma(a2, Pfi, q);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int d = 1 << k;
            const P pA = __shfl_up_sync(0xffffffffu, Aacc, d);
            const P pb = __shfl_up_sync(0xffffffffu, bacc, d);
            if (lane >= d) { bacc = fma(Aacc, pb, bacc); Aacc *= pA; }
        }
        if (lane == 31) { sh.aagg[buf][warp][0] = Aacc; sh.aagg[buf][warp][1] = bacc; }
        P eA = __shfl_up_sync(0xffffffffu, Aacc, 1), eb = __shfl_up_sync(0xffffffffu, bacc, 1);
        if (lane == 0) { eA = P(1); eb = P(0); }
        __syncthreads();
        P mw = m_tile;
        for (int w = 0; w < warp; ++w) mw = fma(sh.aagg[buf][w][0], mw, sh.aagg[buf][w][1]);
        P m = fma(eA, mw, eb);
        if (warp == DIAG_NW - 1 && lane == 31) sh.carry[buf ^ 1][1] = fma(Aacc, mw, bacc);
        // ---- phase 3: filtered means
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const P e = fma(-cc, m, y[i]);
            const P mfi = fma(r[i], e, m);
            y[i] = mfi;
            m = av * mfi;
        }
        store_chunk<P, L>(mfp + start, vec, nvalid, y);
        store_chunk<P, L>(Pfp + start, vec, nvalid, Pf);
    }
}

template <class P>
struct BwdShared {
    P agg[2][DIAG_NW][3];  // G product, mean offset, variance offset
    P carry[2][2];         // [buf][m_s, P_s] at the first frame AFTER the tile
};

template <class P>
__global__ void __launch_bounds__(DIAG_NT, 2) diag_rts_kernel(const __grid_constant__ DiagSmoothArgs<P> a) {
    __shared__ BwdShared<P> sh;
    constexpr int L = DiagTraits<P>::L;
    constexpr int TILE = DIAG_NT * L;
    const int b = blockIdx.x >> 1, c = blockIdx.x & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const P av = a.A[(long long)b * 4 + c * 3], cc = a.C[(long long)b * 4 + c * 3];
    const P q = a.s[b] * a.Q[(long long)b * 4 + c * 3];
    const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
    const P* mfp = a.mf + ((long long)b * 2 + c) * a.T;
    const P* Pfp = a.Pf + ((long long)b * 2 + c) * a.T;
    P* xo = a.out + (long long)b * a.out_seq_stride + a.out_off[c];
    P* vo = a.out + (long long)b * a.out_seq_stride + a.out_off[2 + c];
    const bool vec = (((reinterpret_cast<uintptr_t>(mfp) | reinterpret_cast<uintptr_t>(Pfp) |
                        reinterpret_cast<uintptr_t>(xo) | reinterpret_cast<uintptr_t>(vo)) & 15) == 0);
    if (threadIdx.x == 0) { sh.carry[0][0] = P(0); sh.carry[0][1] = P(0); }
    const P a2 = av * av, c2 = cc * cc;
    const int ntiles = (a.T + TILE - 1) / TILE;
    int buf = 0;
    for (int tile = ntiles - 1; tile >= 0; --tile, buf ^= 1) {
        // thread index increases BACKWARD in time so that an ordinary inclusive scan runs in reverse time
        const int start = tile * TILE + (DIAG_NT - 1 - threadIdx.x) * L;
        const int nvalid = max(0, min(L, a.T - start));
        P mf[L], Pf[L], G[L];
        load_chunk<P, L>(mfp + start, vec, nvalid, P(0), mf);
        load_chunk<P, L>(Pfp + start, vec, nvalid, P(0), Pf);
        // ---- phase 1: compose the chunk's affine maps, last frame first
        P Ag = P(1), bm = P(0), bP = P(0);
#pragma unroll
        for (int ii = 0; ii < L; ++ii) {
            const int i = L - 1 - ii;
            const P Sp = fma(a2, Pf[i], q);
            P g = av * Pf[i] / (Sp + P(1e-9));
            if (start + i >= a.T - 1) g = P(0);  // last frame: smoothed = filtered; padding frames: inert
            G[i] = g;
            const P g2 = g * g;
            bm = fma(g, bm, fma(-g * av, mf[i], mf[i]));
            bP = fma(g2, bP, fma(-g2, Sp, Pf[i]));
            Ag *= g;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int d = 1 << k;
            const P pA = __shfl_up_sync(0xffffffffu, Ag, d);
            const P pm = __shfl_up_sync(0xffffffffu, bm, d);
            const P pP = __shfl_up_sync(0xffffffffu, bP, d);
            if (lane >= d) {
                bm = fma(Ag, pm, bm);
                bP = fma(Ag * Ag, pP, bP);
                Ag *= pA;
            }
        }
        if (lane == 31) { sh.agg[buf][warp][0] = Ag; sh.agg[buf][warp][1] = bm; sh.agg[buf][warp][2] = bP; }
        P eA = __shfl_up_sync(0xffffffffu, Ag, 1), em = __shfl_up_sync(0xffffffffu, bm, 1),
          eP = __shfl_up_sync(0xffffffffu, bP, 1);
        if (lane == 0) { eA = P(1); em = P(0); eP = P(0); }
        __syncthreads();
        P ms = sh.carry[buf][0], Ps = sh.carry[buf][1];
        for (int w = 0; w < warp; ++w) {
            const P g = sh.agg[buf][w][0];
            ms = fma(g, ms, sh.agg[buf][w][1]);
            Ps = fma(g * g, Ps, sh.agg[buf][w][2]);
        }
        if (warp == DIAG_NW - 1 && lane == 31) {  // smoothed state at the first frame of this tile
            sh.carry[buf ^ 1][0] = fma(Ag, ms, bm);
            sh.carry[buf ^ 1][1] = fma(Ag * Ag, Ps, bP);
        }
        ms = fma(eA, ms, em);
        Ps = fma(eA * eA, Ps, eP);
        // ---- phase 3: smoothed moments, reprojected into the output planes
#pragma unroll
        for (int ii = 0; ii < L; ++ii) {
            const int i = L - 1 - ii;
            const P g = G[i], g2 = g * g;
            const P Sp = fma(a2, Pf[i], q);
            ms = fma(g, ms, fma(-g * av, mf[i], mf[i]));
            Ps = fma(g2, Ps, fma(-g2, Sp, Pf[i]));
            mf[i] = fma(cc, ms, mean);  // x = C m + mean   (singlecam_smoother.py:190-197)
            Pf[i] = c2 * Ps;            // diag(C V C^T)    (singlecam_smoother.py:191, 210-211)
        }
        store_chunk<P, L>(xo + start, vec, nvalid, mf);
        store_chunk<P, L>(vo + start, vec, nvalid, Pf);
    }
}

size_t diag_smooth_workspace_bytes(int dtype, int B, int T) {
    return (size_t)B * 2 * T * 2 * (dtype == EKS_F32 ? 4 : 8) + 64;
}

template <class P>
static int diag_smooth_launch(int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                              const void* C, const PlaneView& y, const PlaneView& var, const void* ymean,
                              const void* s, void* out, long long out_seq_stride, const long long* out_off,
                              void* workspace, cudaStream_t st) {
    DiagSmoothArgs<P> a;
    a.B = B; a.T = T;
    a.m0 = (const P*)m0; a.S0 = (const P*)S0; a.A = (const P*)A; a.Q = (const P*)Q; a.C = (const P*)C;
    a.y = y; a.var = var; a.ymean = (const P*)ymean; a.s = (const P*)s;
    // keep the workspace planes 16-byte aligned when T allows
    a.mf = (P*)workspace;
    a.Pf = a.mf + (size_t)B * 2 * T;
    a.out = (P*)out; a.out_seq_stride = out_seq_stride;
    for (int i = 0; i < 4; ++i) a.out_off[i] = out_off[i];
    diag_filter_kernel<P><<<B * 2, DIAG_NT, 0, st>>>(a);
    int rc = check_launch("diag_filter_kernel");
    if (rc) return rc;
    diag_rts_kernel<P><<<B * 2, DIAG_NT, 0, st>>>(a);
    return check_launch("diag_rts_kernel");
}

}  // namespace eks

using namespace eks;

extern "C" size_t eks_diag_smooth_workspace_bytes(int dtype, int B, int T) { return diag_smooth_workspace_bytes(dtype, B, T); }

extern "C" int eks_diag_smooth(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                               const void* C, const void* y_base, long long y_seq_stride, const long long* y_off,
                               const void* ymean, const void* var_base, long long var_seq_stride,
                               const long long* var_off, const void* s, void* out, long long out_seq_stride,
                               const long long* out_off, void* workspace, size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(m0 && S0 && A && Q && C && y_base && y_off && var_base && var_off && s && out && out_off,
                "diag_smooth: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1, "diag_smooth: bad dims");
    EKS_REQUIRE(workspace && workspace_bytes >= diag_smooth_workspace_bytes(dtype, B, T),
                "diag_smooth: workspace too small");
    PlaneView y, var;
    y.base = y_base; y.seq_stride = y_seq_stride;
    var.base = var_base; var.seq_stride = var_seq_stride;
    for (int i = 0; i < MAX_CHAN; ++i) { y.chan_off[i] = i < 2 ? y_off[i] : 0; var.chan_off[i] = i < 2 ? var_off[i] : 0; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        return diag_smooth_launch<float>(B, T, m0, S0, A, Q, C, y, var, ymean, s, out, out_seq_stride, out_off,
                                         workspace, st);
    return diag_smooth_launch<double>(B, T, m0, S0, A, Q, C, y, var, ymean, s, out, out_seq_stride, out_off,
                                      workspace, st);
}
