// diag_lag.cu -- the smoothing-parameter optimiser of the single-camera model (eks/core.py:562-699 on the decoupled
// model of eks/singlecam_smoother.py:246-284) with ONE pass over the observations instead of one pass per Adam
// iteration.
//
// The loss the reference minimises is the constant-R filter NLL.  For A = C = 1 the steady-state innovation obeys
//       e_{t+1} = alpha e_t + d_{t+1},      d_t = y_t - y_{t-1},      alpha = (r + eps) / (P + r + eps)
// (alpha + K = 1 exactly, so the filter has unit DC gain and only the INCREMENTS of y enter), hence
//       sum_t e_t^2 = [ e0^2 + 2 e0 H(alpha) + F(alpha) - alpha^2 Tl(alpha)^2 ] / (1 - alpha^2)
//       F  = R_0 + 2 sum_{m>=1} alpha^m R_m,     R_m = sum_i d_i d_{i+m}          (lagged products of the increments)
//       H  = sum_{m>=1} alpha^m d_{T0+m}   (coupling of the state e0 at frame T0 with the first increments)
//       Tl = sum_{m>=0} alpha^m d_{n-1-m}  (the innovation at the last frame: removes the tail of the infinite sum)
// which depends on the data only through R_0..R_{W-1} once alpha^W is below rounding.  lag_stats_kernel computes the
// R_m in one streaming pass (HBM: 8 bytes per keypoint-frame, once); diag_lag_opt_kernel then runs the WHOLE Adam
// loop of a block in one persistent CTA: per evaluation the exact sequential filter over the first T0 = 256 frames
// (transient of the variance recursion with its s-sensitivity, then the constant-gain recursion), the closed form
// above with its derivative d/ds (through alpha and e0), the Adam step and the reference's stop rule.  All of this is
// float64 arithmetic on the device in both precision modes (the data, the lag products' partial sums and the Adam
// state keep the working precision), so the float32 mode's loss is the float64 loss of the float32 data.
//
// The closed form is used only when it is exact to rounding:  A = C = 1,  n >= T0 + 4 W,  the variance recursion has
// converged within T0 frames, and the truncation bound 2 alpha^W / (1 - alpha) is below 1e-7 (float32 mode, W = 128) /
// 1e-13 (float64 mode, W = 256).  Otherwise THAT evaluation streams the observations inside the same kernel
// (diag_stream_cta: the exact time-parallel evaluation of diag_stream.cuh), so slow-forgetting data stay correct --
// only slower.  No host synchronisation, one launch for the whole optimisation.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ekf_generic.cuh"
#include "diag.cuh"
#include "diag_stream.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int LAG_T0 = 256;     // frames [0, T0) are filtered sequentially in an evaluation; statistics start after
constexpr int LAG_NT0 = 5;      // statistics are kept for T0 = 256, 128, 64, 32, 16 (lag_reduce_kernel); the evaluation picks
constexpr int LAG_CH = 4096;    // increments per shared-memory tile of lag_stats_kernel
constexpr int LAG_RM = 16;      // lags per thread (register tile), float64; float32 uses LAG_RM32
constexpr int LAG_RM32 = 32;    // float32: 16 frames x 32 lags per step = 16 16-byte shared loads per 512 FMAs (16 x 16: 12 per
                                // 256, measured shared-memory bound at 54 % of the FMA pipe)
constexpr int LAG_NPART_MAX = 2;   // warps sharing one lag group split the tile's steps between them
constexpr int LAG_RP = 16;      // frames per thread and step (register tile)
constexpr int LAG_NT = 256;
constexpr int LAG_CPB = 16;     // tiles per CTA (accumulated in registers before the partial sums are written; the final warp reduction
                                // of a CTA was 5 % of the instructions at 8 tiles)

template <class P>
struct LagStatArgs {
    PlaneView y;
    int B, t_begin, n, nchunk, nx;   // nx = gridDim.x
    double* partial;                 // [2B][nx][W]
    double* R;                       // [LAG_NT0][2B][W]
};

template <class P> struct LagVec;
template <> struct LagVec<float> { using type = float4; static constexpr int VW = 4; };
template <> struct LagVec<double> { using type = double2; static constexpr int VW = 2; };

// physical position of logical element x: 16 bytes of padding after every 16 elements, so that the 16-byte shared
// loads of lanes whose chunks are 16 elements apart fall into different bank groups (stride 80 B / 144 B)
template <class P>
__device__ __forceinline__ int lag_phys(int x) { return x + (x >> 4) * (16 / (int)sizeof(P)); }

// R_m partial sums.  grid = (nx, 2B); CTA (x, 2b+c) handles the tiles x, x + nx, ... of channel c of sequence b.
// A tile = LAG_CH increments d_i (+ W halo) staged in shared memory; warp w owns the lag groups w, w + 8, ... (16 lags
// each) and lane l the frames (step * 32 + l) * 16 ... + 15 of the tile: a 16 x 16 register tile of products per step
// from 12 16-byte shared loads.  float32 mode: products and the <= 128-term partial sums per tile in float32, promoted
// to float64 per tile (error of a partial ~1e-6 relative, of the 10^6-frame sum ~1e-8; the data are float32 anyway).
template <class P, int W>
__global__ void __launch_bounds__(LAG_NT, sizeof(P) == 4 ? 2 : 1) lag_stats_kernel(const __grid_constant__ LagStatArgs<P> a) {
    constexpr int PADE = 16 / (int)sizeof(P);
    constexpr int NLOG = LAG_CH + W;
    constexpr int NPHYS = NLOG + (NLOG / 16) * PADE + PADE;
    constexpr int RM = sizeof(P) == 4 ? LAG_RM32 : LAG_RM;   // lags per thread
    constexpr int NG = W / RM;              // lag groups
    constexpr int NW = LAG_NT / 32;
    constexpr int GPW = (NG + NW - 1) / NW; // lag groups per warp (NG >= 8)
    constexpr int NPART = NG < NW ? NW / NG : 1;   // warps per lag group (NG < 8): they interleave the steps of a tile
    static_assert(NPART <= LAG_NPART_MAX && (NG >= NW || NW % NG == 0), "lag tile geometry");
    constexpr int VW = LagVec<P>::VW;
    using V = typename LagVec<P>::type;
    __shared__ __align__(16) P sm[NPHYS];
    const int bc = blockIdx.y, b = bc >> 1, c = bc & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = NPART > 1 ? warp / NG : 0;
    const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[c];
    double accd[GPW][RM];
#pragma unroll
    for (int q = 0; q < GPW; ++q)
#pragma unroll
        for (int j = 0; j < RM; ++j) accd[q][j] = 0.0;
    for (int chunk = blockIdx.x; chunk < a.nchunk; chunk += a.nx) {
        const int i0 = LAG_T0 + 1 + chunk * LAG_CH;
        __syncthreads();   // the previous tile has been consumed
        {   // staging: every load of the tile is issued before the first use (a rolled loop exposed one DRAM latency per
            // element: ncu long_scoreboard 27 % of the stalls); the left neighbour comes from the previous lane
            constexpr int U = (NLOG + LAG_NT - 1) / LAG_NT;
            P cur[U], left[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + threadIdx.x + u * LAG_NT;
                cur[u] = (i < a.n) ? __ldg(yc + i) : P(0);
                left[u] = (lane == 0 && i < a.n) ? __ldg(yc + i - 1) : P(0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = threadIdx.x + u * LAG_NT;
                const int i = i0 + x;
                const P up = __shfl_up_sync(0xffffffffu, cur[u], 1);
                const P prev = lane == 0 ? left[u] : up;
                if (x < NLOG) sm[lag_phys<P>(x)] = (i < a.n) ? cur[u] - prev : P(0);
            }
        }
        __syncthreads();
        const int nvalid = min(LAG_CH, a.n - i0);   // increments of this tile that exist
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int g = NPART > 1 ? warp % NG : warp + NW * q;
            if (g >= NG) break;
            const int m0 = g * RM;
            P acc[RM];
#pragma unroll
            for (int j = 0; j < RM; ++j) acc[j] = P(0);
            for (int step = part; step < LAG_CH / (32 * LAG_RP); step += NPART) {
                const int p = (step * 32 + lane) * LAG_RP;
                if (step * 32 * LAG_RP >= nvalid) break;      // warp-uniform: nothing left in this tile
                P av[LAG_RP], bv[LAG_RP + RM];
                const P* pa = sm + lag_phys<P>(p);
#pragma unroll
                for (int i = 0; i < LAG_RP / VW; ++i) {
                    const V v = *reinterpret_cast<const V*>(pa + i * VW);
                    const P* ev = reinterpret_cast<const P*>(&v);
#pragma unroll
                    for (int k = 0; k < VW; ++k) av[i * VW + k] = ev[k];
                }
#pragma unroll
                for (int blk = 0; blk < (LAG_RP + RM) / 16; ++blk) {     // the lag window, 16 elements (one padded run) at a time
                    const P* pb = sm + lag_phys<P>(p + m0 + 16 * blk);
#pragma unroll
                    for (int i = 0; i < 16 / VW; ++i) {
                        const V w0 = *reinterpret_cast<const V*>(pb + i * VW);
                        const P* e0 = reinterpret_cast<const P*>(&w0);
#pragma unroll
                        for (int k = 0; k < VW; ++k) bv[16 * blk + i * VW + k] = e0[k];
                    }
                }
                // (a packed-FFMA2 version of the 16 x 16 tile -- even / odd frames as the two lanes, shifted copy of the
                // window for the odd lags -- was measured SLOWER: 688 vs 542 us per 80 channel-planes; shared-memory bound)
#pragma unroll
                for (int i = 0; i < LAG_RP; ++i)
#pragma unroll
                    for (int j = 0; j < RM; ++j) acc[j] = fma(av[i], bv[i + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < RM; ++j) accd[q][j] += (double)acc[j];
        }
    }
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        const int g = NPART > 1 ? warp % NG : warp + NW * q;
        if (g >= NG) break;
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const double v = warp_sum(accd[q][j]);
            if (lane == 0)
                a.partial[(((long long)bc * a.nx + blockIdx.x) * LAG_NPART_MAX + part) * W + g * RM + j] = v;
        }
    }
}

// fixed-order sum of the per-CTA partials: one thread per (sequence, channel, lag).  R[0] = statistics of the increments
// after frame T0 = 256 (what lag_stats_kernel summed); R[1], R[2] = the same from frame 128 / 64 on (the few extra
// products are added here), so that an evaluation whose variance transient ends early need not walk to frame 256.
template <class P, int W>
__global__ void lag_reduce_kernel(const __grid_constant__ LagStatArgs<P> a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)a.B * 2 * W;
    if (idx >= per) return;
    const long long bc = idx / W;
    const int m = (int)(idx - bc * W);
    double s = 0;
    for (int x = 0; x < a.nx * LAG_NPART_MAX; ++x) s += a.partial[(bc * a.nx * LAG_NPART_MAX + x) * W + m];   // unused slots are zero
    a.R[idx] = s;
    const int b = (int)(bc >> 1), c = (int)(bc & 1);
    const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[c];
    int hi = LAG_T0;
#pragma unroll
    for (int lvl = 1; lvl < LAG_NT0; ++lvl) {
        const int lo = LAG_T0 >> lvl;              // increments lo + 1 .. hi join the statistics
        for (int i = hi; i > lo; --i)
            s += ((double)yc[i] - (double)yc[i - 1]) * ((double)yc[i + m] - (double)yc[i + m - 1]);
        a.R[lvl * per + idx] = s;
        hi = lo;
    }
}

// ---- one (sequence, channel) evaluation, float64: exact sequential filter over the first T0 frames
struct LagPair {
    double alpha, dalpha, iS, diS, logS, dlogS, e0, de0;
    double sl, sdl, se, sde, sg;   // transient sums: logS, dlogS, e2 iS, e2 diS, e dm iS
    double E2b, Gb;                // steady-state frames [t_c, T0): sum e^2, sum e dm
    int t_c;
    int T0, lvl;                   // first frame of the statistics used by this evaluation and its index in R
};

// yh: this pair's first T0 + 1 observations (centred, float64).  Returns false if the closed form does not apply.
template <class P>
__device__ bool lag_prepare(const DiagOptArgs<P>& a, int b, int c, double s, const double* __restrict__ yh, int W,
                            double tolF, LagPair& o) {
    const double av = (double)a.A[(long long)b * 4 + c * 3], cc = (double)a.C[(long long)b * 4 + c * 3];
    const double Qc = (double)a.Q[(long long)b * 4 + c * 3], r = (double)a.Rconst[(long long)b * 2 + c];
    if (av != 1.0 || cc != 1.0) return false;
    double Pv = (double)a.S0[(long long)b * 4 + c * 3], dP = 0, m = (double)a.m0[(long long)b * 2 + c], dm = 0;
    double sl = 0, sdl = 0, se = 0, sde = 0, sg = 0;
    // The variance recursion is followed frame by frame only until it is within 1e-8 (relative) of its fixed point;
    // the fixed point itself is then polished by Newton steps in float64, so the steady-state constants are exact and
    // the frames after t_c differ from the exact filter by < 1e-8 relative in P for a few frames (loss error ~1e-8
    // absolute: five orders below the reference's stop threshold).  Following it to rounding took 2.5x more frames.
    const double tol = 1e-8, BOOST = 1e-9;
    double prevP = INFINITY, prevd = INFINITY;
    int stall = 0, t = 0;
    double iS, diS, K, dK, alpha, S, dS, prodS = 1.0;
    while (true) {
        S = Pv + r;
        iS = 1.0 / S;
        const double x = BOOST * iS;
        const double iSb = iS * (1.0 - x + x * x);      // 1 / (S + eps) to rounding (x ~ 1e-8: third order < 1e-24)
        dS = dP;
        diS = -dS * iS * iS;
        K = Pv * iSb;
        dK = dP * iSb - Pv * dS * iSb * iSb;
        const double Pf = Pv * iSb * (r + BOOST * (1.0 + K));   // P - K S K without cancellation (diag_transient)
        const double dPf = r * (dP * iS + Pv * diS);
        const double Pn = Pf + s * Qc, dPn = dPf + Qc;
        alpha = iSb * (r + BOOST);
        const double gap = 1.0 - alpha * alpha;
        const double chgP = fabs(Pn - Pv), chgd = fabs(dPn - dP);
        bool conv = (chgP <= tol * gap * fabs(Pn)) && (chgd <= tol * gap * fabs(dPn));
        if (chgP >= prevP && chgd >= prevd) ++stall;
        if (stall >= 24) conv = true;
        prevP = chgP; prevd = chgd;
        if (conv) break;
        if (t >= LAG_T0) return false;                          // slow convergence: stream this evaluation
        const double e = yh[t] - m;
        prodS *= S;                                             // sum log S_t through products of 8 (S stays in range)
        if ((t & 7) == 7) { sl += log(prodS); prodS = 1.0; }
        sdl += dS * iS;
        se += e * e * iS;
        sde += e * e * diS;
        sg += e * dm * iS;
        dm = dm + dK * e - K * dm;
        m = m + K * e;
        Pv = Pn; dP = dPn;
        ++t;
    }
    {   // Newton on F(P) - P = 0, F(P) = P iSb (r + eps (1 + K)) + s Q  (the exact update of the loop above)
        const double sQ = s * Qc;
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
            const double iSb = 1.0 / (Pv + r + BOOST), Kk = Pv * iSb, cst = r + BOOST * (1.0 + Kk);
            const double al = (r + BOOST) * iSb;
            const double F = Pv * iSb * cst + sQ;
            const double dF = al * iSb * cst + Pv * iSb * iSb * BOOST * al;
            Pv -= (F - Pv) / (dF - 1.0);
        }
        S = Pv + r;
        iS = 1.0 / S;
        const double iSb = 1.0 / (S + BOOST);
        const double rr = r * iS;
        dP = Qc / (1.0 - rr * rr);                 // fixed point of dP' = r (dP iS + P diS) + Q, the loop's recursion
        dS = dP;
        diS = -dS * iS * iS;
        K = Pv * iSb;
        dK = dP * iSb - Pv * dS * iSb * iSb;
        alpha = iSb * (r + BOOST);
    }
    if (!(alpha > 0.0) || !(alpha < 1.0) || !isfinite(alpha)) return false;
    // truncation of the lag series: |sum_{m >= W} 2 alpha^m R_m| <= 2 alpha^W / (1 - alpha) R_0
    double aW = alpha;                                           // alpha^W, W a power of two: log2(W) squarings
    for (int w = W; w > 1; w >>= 1) aW *= aW;
    if (2.0 * aW > tolF * (1.0 - alpha)) return false;
    sl += log(prodS);
    o.t_c = t;
    int lvl = LAG_NT0 - 1;                                      // smallest T0 of the table that is >= t_c
    while (lvl > 0 && (LAG_T0 >> lvl) < t) --lvl;
    const int T0 = LAG_T0 >> lvl;
    o.T0 = T0; o.lvl = lvl;
    double E2b = 0, Gb = 0;
    for (; t < T0; ++t) {               // constant-gain frames before the statistics start
        const double e = yh[t] - m;
        E2b = fma(e, e, E2b);
        Gb = fma(e, dm, Gb);
        dm = fma(alpha, dm, dK * e);
        m = fma(K, e, m);
    }
    o.e0 = yh[T0] - m;
    o.de0 = -dm;
    o.alpha = alpha; o.dalpha = -dK; o.iS = iS; o.diS = diS; o.logS = log(S); o.dlogS = dS * iS;
    o.sl = sl; o.sdl = sdl; o.se = se; o.sde = sde; o.sg = sg; o.E2b = E2b; o.Gb = Gb;
    return true;
}

template <class P>
struct LagOptArgs {
    DiagOptArgs<P> d;
    const double* R;   // [LAG_NT0][2B][W]
    int W;
    int fast;          // 0: the closed form is never applicable (short sequences): stream every evaluation
    double tolF;
};

// loss / gradient of a block from the streamed partial sums (the consumer half of diag_adam_body, nseg == 1)
template <class P>
__device__ void lag_stream_loss(const DiagOptArgs<P>& a, int m_lo, int m_hi, P dsdlog, P& loss, P& grad) {
    const double HALF_LOG2PI = 0.91893853320467274178;
    loss = P(0); grad = P(0);
    for (int mi = m_lo; mi < m_hi; ++mi) {
        const int b = a.members[mi];
        double nll = 0, dnll = 0;
        for (int c = 0; c < 2; ++c) {
            const ChanState<P>& cs = a.cstate[(long long)b * 2 + c];
            const double* part = a.partials + ((long long)b * 2 + c) * 2;
            const double te = part[0], tg = part[1];
            const double nB = (double)(a.n - cs.t_c);
            const ChanConst<P>& k = cs.k;
            nll += (double)a.n * HALF_LOG2PI + 0.5 * cs.tsum[0] + 0.5 * cs.tsum[2] + 0.5 * nB * (double)k.logS +
                   0.5 * (double)k.iS * te;
            dnll += 0.5 * cs.tsum[1] + 0.5 * cs.tsum[3] - cs.tsum[4] + 0.5 * nB * (double)k.dlogS +
                    0.5 * (double)k.diS * te - (double)k.cc * (double)k.iS * (double)k.dbeta * tg;
        }
        P v = (P)nll, g = (P)dnll;
        if (!isfinite(nll) || !isfinite((double)v)) { v = P(1e12); g = P(0); }  // core.py:650
        loss += v;
        grad += g * dsdlog;
    }
}

// ---- the persistent optimiser: one CTA per block (group of sequences sharing one s), whole Adam loop
template <class P>
__global__ void __launch_bounds__(OPT_NT, 2) diag_lag_opt_kernel(const __grid_constant__ LagOptArgs<P> la) {
    const DiagOptArgs<P>& a = la.d;
    __shared__ ChanConst<P> shk;
    __shared__ double red[OPT_NW][2];
    __shared__ int sh_mode, sh_done, sh_head_ok;
    __shared__ double sh_s;
    extern __shared__ __align__(16) unsigned char ring[];
    double* yh_all = reinterpret_cast<double*>(ring);     // head cache: [32 pairs][T0 + 1], valid in FAST mode only
    constexpr int YH = LAG_T0 + 1;
    const int j = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BlockState<P>& bs = a.bstate[j];
    const int m_lo = a.block_off[j], m_hi = a.block_off[j + 1];
    const int npair = (m_hi - m_lo) * 2;
    const int W = la.W;
    const double HALF_LOG2PI = 0.91893853320467274178;
    if (threadIdx.x == 0) {
        adam_init(bs.adam, a.s_log0[j]);
        bs.done = (a.cap <= 0);
        if (bs.done) { a.s_log_out[j] = bs.adam.s_log; a.last_loss_out[j] = bs.adam.prev; a.iters_out[j] = 0; }
        sh_done = bs.done;
        sh_head_ok = 0;
    }
    __syncthreads();
    if (sh_done) return;
    while (true) {
        P loss = P(0), grad = P(0);      // meaningful on thread 0
        if (warp == 0) {
            if (lane == 0) {
                P dsdlog;
                bs.s = adam_current_s(bs.adam, a.lo, a.hi, &dsdlog);
                bs.dsdlog = dsdlog;
                sh_s = (double)bs.s;
            }
            __syncwarp();
            const double s = sh_s;
            bool fast = la.fast != 0;
            for (int p0 = 0; fast && p0 < npair; p0 += 32) {
                const int np = min(32, npair - p0);
                // head cache: the first T0 + 1 observations of each pair of this group (centred, float64)
                if (!(sh_head_ok && npair <= 32)) {
                    for (int q = 0; q < np; ++q) {
                        const int b = a.members[m_lo + ((p0 + q) >> 1)], c = (p0 + q) & 1;
                        const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin +
                                      a.y.chan_off[c];
                        const double mean = a.ymean ? (double)a.ymean[(long long)b * 2 + c] : 0.0;
                        for (int t = lane; t < YH; t += 32) yh_all[q * YH + t] = (double)yc[t] - mean;
                    }
                    __syncwarp();
                    if (lane == 0) sh_head_ok = (npair <= 32);
                }
                LagPair lp;
                bool ok = true;
                int b = 0, c = 0;
                if (lane < np) {
                    b = a.members[m_lo + ((p0 + lane) >> 1)]; c = (p0 + lane) & 1;
                    ok = lag_prepare<P>(a, b, c, s, yh_all + lane * YH, W, la.tolF, lp);
                }
                if (!__all_sync(0xffffffffu, ok)) { fast = false; break; }
                // lag series of each pair, all lanes cooperating: lane l takes the lags l, l + 32, ...
                double F = 0, dF = 0, H = 0, dH = 0, Tl = 0, dTl = 0;
                for (int q = 0; q < np; ++q) {
                    const double alq = __shfl_sync(0xffffffffu, lp.alpha, q);
                    const int bq = __shfl_sync(0xffffffffu, b, q), cq = __shfl_sync(0xffffffffu, c, q);
                    const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)bq * a.y.seq_stride + a.t_begin +
                                  a.y.chan_off[cq];
                    const int T0q = __shfl_sync(0xffffffffu, lp.T0, q), lvq = __shfl_sync(0xffffffffu, lp.lvl, q);
                    const double* Rq = la.R + ((long long)lvq * a.B * 2 + (long long)bq * 2 + cq) * W;
                    // alpha^lane and alpha^32 by binary powering (10 multiplications; log + two exp calls were ~300
                    // instructions on the critical path of every evaluation, ncu r2); relative error a few ulps
                    double pw = 1.0, sq = alq;
#pragma unroll
                    for (int bit = 0; bit < 5; ++bit) {
                        if ((lane >> bit) & 1) pw *= sq;
                        sq *= sq;
                    }
                    const double pw32 = sq, ial = 1.0 / alq;
                    double f = 0, df = 0, h = 0, dh = 0, tl = 0, dtl = 0;
                    for (int m = lane; m < W; m += 32) {
                        const double pm1 = (double)m * pw * ial;            // m alpha^(m-1)
                        const double Rm = Rq[m];
                        const double cm = m == 0 ? 1.0 : 2.0;
                        f = fma(cm * pw, Rm, f);
                        df = fma(cm * pm1, Rm, df);
                        const double lm = (double)yc[a.n - 1 - m] - (double)yc[a.n - 2 - m];
                        tl = fma(pw, lm, tl);
                        dtl = fma(pm1, lm, dtl);
                        if (m >= 1) {
                            const double hm = (double)yc[T0q + m] - (double)yc[T0q + m - 1];
                            h = fma(pw, hm, h);
                            dh = fma(pm1, hm, dh);
                        }
                        pw *= pw32;
                    }
                    f = warp_sum(f); df = warp_sum(df); h = warp_sum(h); dh = warp_sum(dh);
                    tl = warp_sum(tl); dtl = warp_sum(dtl);
                    if (lane == q) { F = f; dF = df; H = h; dH = dh; Tl = tl; dTl = dtl; }
                }
                double nll = 0, dnll = 0;
                if (lane < np) {
                    const double al = lp.alpha, den = 1.0 - al * al, e0 = lp.e0;
                    const double Nn = e0 * e0 + 2.0 * e0 * H + F - al * al * Tl * Tl;
                    const double E2s = Nn / den;
                    const double Nn_a = 2.0 * e0 * dH + dF - 2.0 * al * Tl * Tl - 2.0 * al * al * Tl * dTl;
                    const double E2s_a = (Nn_a + 2.0 * al * E2s) / den;
                    const double E2s_e = (2.0 * e0 + 2.0 * H) / den;
                    const double dE2s = E2s_a * lp.dalpha + E2s_e * lp.de0;
                    const double nB = (double)(a.n - lp.t_c), E2 = lp.E2b + E2s;
                    nll = (double)a.n * HALF_LOG2PI + 0.5 * lp.sl + 0.5 * lp.se + 0.5 * nB * lp.logS + 0.5 * lp.iS * E2;
                    dnll = 0.5 * lp.sdl + 0.5 * lp.sde - lp.sg + 0.5 * nB * lp.dlogS + 0.5 * lp.diS * E2 -
                           lp.iS * lp.Gb + 0.5 * lp.iS * dE2s;
                }
                // members in order (the reference sums the member losses sequentially, eks/core.py:474-476)
                nll += __shfl_xor_sync(0xffffffffu, nll, 1);
                dnll += __shfl_xor_sync(0xffffffffu, dnll, 1);
                for (int q = 0; q < np; q += 2) {
                    const double nq = __shfl_sync(0xffffffffu, nll, q), dq = __shfl_sync(0xffffffffu, dnll, q);
                    P v = (P)nq, g = (P)dq;
                    if (!isfinite(nq) || !isfinite((double)v)) { v = P(1e12); g = P(0); }  // core.py:650
                    loss += v;
                    grad += g * bs.dsdlog;
                }
            }
            if (lane == 0) sh_mode = fast ? 1 : 0;
        }
        __syncthreads();
        if (sh_mode == 0) {
            // stream this evaluation: transient in the working precision, then one pass over every member's channels
            if (warp == 0) {
                const P s = bs.s;
                for (int p0 = 0; p0 < npair; p0 += 32) {
                    const int p = p0 + lane;
                    if (p < npair) {
                        const int b = a.members[m_lo + (p >> 1)], c = p & 1;
                        diag_transient<P>(a, b, c, s, a.cstate[(long long)b * 2 + c]);
                    }
                }
                if (lane == 0) sh_head_ok = 0;      // the ring is about to overwrite the head cache
            }
            __threadfence_block();
            __syncthreads();
            for (int p = 0; p < npair; ++p) {
                const int b = a.members[m_lo + (p >> 1)], c = p & 1;
                double te, tg;
                diag_stream_cta<P>(a, b, c, 0, 1, ring, shk, red, te, tg);
                if (threadIdx.x == 0) {
                    double* part = a.partials + ((long long)b * 2 + c) * 2;
                    part[0] = te; part[1] = tg;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) lag_stream_loss<P>(a, m_lo, m_hi, bs.dsdlog, loss, grad);
        }
        if (threadIdx.x == 0) {
            if (a.trace && bs.adam.iters < a.trace_cap) {
                P* tr = a.trace + ((long long)j * a.trace_cap + bs.adam.iters) * 3;
                tr[0] = bs.adam.s_log; tr[1] = loss; tr[2] = grad * a.lr;
            }
            adam_step(bs.adam, loss, grad, a.lr, a.tol, a.cap);
            if (bs.adam.done) {
                bs.done = 1;
                a.s_log_out[j] = bs.adam.s_log;
                a.last_loss_out[j] = bs.adam.prev;
                a.iters_out[j] = bs.adam.iters;
            }
            sh_done = bs.adam.done;
        }
        __syncthreads();
        if (sh_done) break;
    }
}

static int lag_W(int dtype) { return dtype == EKS_F32 ? 128 : 256; }

size_t diag_lag_workspace_bytes(int dtype, int n_blocks, int B, int T) {
    const int W = lag_W(dtype);
    const int nchunk = (T + LAG_CH - 1) / LAG_CH + 1;
    const int nx = (nchunk + LAG_CPB - 1) / LAG_CPB;
    size_t bytes = diag_optimize_workspace_bytes(dtype, n_blocks, B, T);
    bytes += (size_t)LAG_NT0 * B * 2 * W * sizeof(double) + 256;
    bytes += (size_t)B * 2 * nx * LAG_NPART_MAX * W * sizeof(double) + 256;
    return bytes;
}

template <class P, int W>
static int diag_lag_run(DiagOptArgs<P>& a, void* workspace, size_t workspace_bytes, int dtype, int T, cudaStream_t st) {
    static_assert(sizeof(BlockState<P>) <= 128 && sizeof(ChanState<P>) <= 1024, "workspace bound");
    EKS_REQUIRE(workspace && workspace_bytes >= diag_lag_workspace_bytes(dtype, a.n_blocks, a.B, T),
                "optimize_s: workspace too small");
    a.nseg = 1;
    unsigned char* w = (unsigned char*)workspace;
    a.n_active = (int*)w; w += 256;
    a.bstate = (BlockState<P>*)w; w += (size_t)a.n_blocks * 128;
    a.cstate = (ChanState<P>*)w; w += (size_t)a.B * 2 * 1024;
    a.partials = (double*)w; w += (size_t)a.B * 2 * 2 * sizeof(double);
    w = (unsigned char*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    a.seq_block = nullptr; a.block_counter = nullptr;
    a.blk_lo = 0; a.blk_hi = a.n_blocks;
    LagOptArgs<P> la;
    la.d = a;
    la.W = W;
    la.fast = (a.n >= LAG_T0 + 4 * W) ? 1 : 0;
    la.tolF = dtype == EKS_F32 ? 1e-7 : 1e-13;
    double* R = (double*)w; w += (size_t)LAG_NT0 * a.B * 2 * W * sizeof(double) + 256;
    la.R = R;
    int launches = 1;
    if (la.fast) {
        LagStatArgs<P> sa;
        sa.y = a.y; sa.B = a.B; sa.t_begin = a.t_begin; sa.n = a.n;
        sa.nchunk = (a.n - (LAG_T0 + 1) + LAG_CH - 1) / LAG_CH;
        sa.nx = (sa.nchunk + LAG_CPB - 1) / LAG_CPB;
        sa.partial = (double*)w;
        sa.R = R;
        cudaMemsetAsync(sa.partial, 0, (size_t)a.B * 2 * sa.nx * LAG_NPART_MAX * W * sizeof(double), st);
        lag_stats_kernel<P, W><<<dim3(sa.nx, 2 * a.B), LAG_NT, 0, st>>>(sa);
        int rc = check_launch("lag_stats_kernel");
        if (rc) return rc;
        const long long nred = (long long)a.B * 2 * W;
        lag_reduce_kernel<P, W><<<(unsigned)((nred + 255) / 256), 256, 0, st>>>(sa);
        rc = check_launch("lag_reduce_kernel");
        if (rc) return rc;
        launches += 2;
    }
    const int smem = OPT_NW * OPT_STAGES * WRP_STAGE_BYTES;
    static_assert(OPT_NW * OPT_STAGES * WRP_STAGE_BYTES >= 32 * (LAG_T0 + 1) * (int)sizeof(double),
                  "the ring must hold the head cache");
    cudaError_t e = cudaFuncSetAttribute(diag_lag_opt_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        set_error("diag_lag_opt_kernel: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    diag_lag_opt_kernel<P><<<a.n_blocks, OPT_NT, smem, st>>>(la);
    note_launches(launches);
    return check_launch("diag_lag_opt_kernel");
}

int diag_lag_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                      const void* C, const void* y_base, long long y_seq_stride, const long long* y_off,
                      const void* ymean, const void* Rconst, int t_begin, int n, int n_blocks, const int* block_off,
                      const int* members, const void* s_log0, double lr, double lo, double hi, double tol, int cap,
                      void* s_log_out, void* last_loss_out, int* iters_out, void* trace, int trace_cap, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
#define EKS_FILL(PT, WW)                                                                                    \
    DiagOptArgs<PT> a;                                                                                      \
    memset(&a, 0, sizeof(a));                                                                               \
    a.B = B; a.t_begin = t_begin; a.n = n;                                                                  \
    a.m0 = (const PT*)m0; a.S0 = (const PT*)S0; a.A = (const PT*)A; a.Q = (const PT*)Q; a.C = (const PT*)C; \
    a.y.base = y_base; a.y.seq_stride = y_seq_stride;                                                       \
    for (int i = 0; i < MAX_CHAN; ++i) a.y.chan_off[i] = i < 2 ? y_off[i] : 0;                              \
    a.ymean = (const PT*)ymean; a.Rconst = (const PT*)Rconst;                                               \
    a.n_blocks = n_blocks; a.block_off = block_off; a.members = members; a.s_log0 = (const PT*)s_log0;      \
    a.lr = (PT)lr; a.lo = (PT)lo; a.hi = (PT)hi; a.tol = (PT)tol; a.cap = cap;                              \
    a.s_log_out = (PT*)s_log_out; a.last_loss_out = (PT*)last_loss_out; a.iters_out = iters_out;            \
    a.trace = (PT*)trace; a.trace_cap = trace_cap;                                                          \
    return diag_lag_run<PT, WW>(a, workspace, workspace_bytes, dtype, T, st);
    if (dtype == EKS_F32) { EKS_FILL(float, 128) }
    EKS_FILL(double, 256)
#undef EKS_FILL
}

}  // namespace eks
