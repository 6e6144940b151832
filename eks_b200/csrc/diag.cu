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
#include <cstdlib>
#include "common.cuh"
#include "ekf_generic.cuh"
#include "diag.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int DIAG_NT = 256;
constexpr int DIAG_NW = DIAG_NT / 32;
#ifndef EKS_OPT_NW
#define EKS_OPT_NW 8     // warps per CTA of diag_nll_kernel (each warp = one independent run of warp-tiles)
#endif
constexpr int OPT_NW = EKS_OPT_NW, OPT_NT = 32 * OPT_NW;
constexpr int OPT_NSEG_MAX = 16 * 8 / OPT_NW;   // at most 128 runs per (sequence, channel)

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
    P gamma;         // -c beta: coupling of the scaled recursion (see diag_warp_tile)
    P aL[5], bL[5];  // Phi^(L 2^k) = [[aL,0],[bL,aL]]
    P aW, bW;        // Phi^(32 L)
    P aH, bH;        // Phi^(L/2)
    P aQ, bQ;        // Phi^(L/4)
};

// shared-memory tile ring: every thread's chunk is CHUNK_BYTES of frames, padded to PAD_BYTES so that the
// per-thread 16-byte reads are bank-conflict free (stride 144 B = 9 x 16 B)
constexpr int OPT_CHUNK_BYTES = 128;
constexpr int OPT_PAD_BYTES = 144;
#ifndef EKS_OPT_STAGES
#define EKS_OPT_STAGES 2
#endif
#ifndef EKS_OPT_MINBLOCKS
#define EKS_OPT_MINBLOCKS 3
#endif
#ifndef EKS_OPT_RELOAD
#define EKS_OPT_RELOAD 0
#endif
constexpr int OPT_STAGES = EKS_OPT_STAGES;

// ---- device-resident optimiser state -----------------------------------------------------------------
template <class P>
struct BlockState {          // one per block (group of sequences sharing one s)
    AdamState<P> adam;
    P s, dsdlog;
    int done;
    int pad;
};
template <class P>
struct ChanState {           // one per (sequence, channel): produced by diag_adam_kernel for the current s
    ChanConst<P> k;
    P z0[2];                 // (m, dm) at frame t_c
    P a_lane[32], b_lane[32];// Phi^(L lane) = [[a_lane,0],[b_lane,a_lane]] for folding a warp carry
    double tsum[5];          // transient sums: logS, dlogS, e2 iS, e2 diS, cc e dm iS
    int t_c;                 // first steady-state frame (multiple of 4)
    int warm;                // frames after which a zero carry-in is forgotten below rounding
};

template <class P>
struct DiagOptArgs {
    int B, t_begin, n, nseg;
    const P *m0, *S0, *A, *Q, *C;
    PlaneView y;
    const P *ymean, *Rconst;
    int n_blocks;
    const int *block_off, *members;
    const int* seq_block;        // [B] block index of every sequence
    const P* s_log0;
    P lr, lo, hi, tol;
    int cap;
    P *s_log_out, *last_loss_out;
    int* iters_out;
    P* trace;
    int trace_cap;
    BlockState<P>* bstate;       // [n_blocks]
    ChanState<P>* cstate;        // [B][2]
    double* partials;            // [B][2][nseg][2]  (sum e^2, sum e dm) per segment
    int* n_active;
    int* block_counter;          // [n_blocks] CTAs of the current evaluation that have finished
    int blk_lo, blk_hi;          // this launch evaluates the blocks in [blk_lo, blk_hi) only
};

__device__ inline void cp_async_16(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ inline void cp_async_8(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ inline void cp_async_4(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }


// ---- transient: sequential scalar filter with s-sensitivities until the variance recursion has reached
// its floating-point fixed point (or the sequence ends).  One thread per (sequence, channel).
template <class P>
__device__ void diag_transient(const DiagOptArgs<P>& a, int b, int c, P s, ChanState<P>& out) {
    const P av = a.A[(long long)b * 4 + c * 3], cc = a.C[(long long)b * 4 + c * 3], Qc = a.Q[(long long)b * 4 + c * 3];
    const P r = a.Rconst[(long long)b * 2 + c];
    const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
    const P* yp = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.y.chan_off[c] + a.t_begin;
    P Pv = a.S0[(long long)b * 4 + c * 3], dP = P(0), m = a.m0[(long long)b * 2 + c], dm = P(0);
    double sl = 0, sdl = 0, se = 0, sde = 0, sg = 0;
    const P tol = P(8) * DiagTraits<P>::eps();
    const P BOOST = P(1e-9);
    P prevdP_chg = P(INFINITY), prevP_chg = P(INFINITY);
    int stall = 0;
    int t = 0;
    const int n = a.n;
    P S, iS, dS, diS, K, dK, alpha;
    while (true) {
        S = cc * cc * Pv + r;
        iS = P(1) / S;
        const P iSb = P(1) / (S + BOOST);       // psd_solve boosts the gain solve only
        dS = cc * cc * dP;
        diS = -dS * iS * iS;
        K = Pv * cc * iSb;
        dK = cc * (dP * iSb - Pv * dS * iSb * iSb);
        // P_f = P - K S K and alpha = a (1 - K c), written without cancellation (identical algebra)
        const P Pf = Pv * iSb * (r + BOOST * (P(1) + cc * K));
        const P dPf = r * (dP * iS + Pv * diS);
        const P Pn = av * av * Pf + s * Qc;
        const P dPn = av * av * dPf + Qc;
        alpha = av * iSb * (r + BOOST);
        // convergence of (P, dP): relative step below tol * (1 - rho), rho = alpha^2 the contraction
        // factor, or the iteration has hit its rounding floor (steps no longer shrinking)
        const P gap = P(1) - alpha * alpha;
        const P chgP = fabs(Pn - Pv), chgd = fabs(dPn - dP);
        bool conv = (chgP <= tol * gap * fabs(Pn)) && (chgd <= tol * gap * fabs(dPn));
        if (chgP >= prevP_chg && chgd >= prevdP_chg) ++stall;
        if (stall >= 24) conv = true;
        prevP_chg = chgP;
        prevdP_chg = chgd;
        if ((conv && (t & 3) == 0) || t >= n) break;
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
        ++t;
    }
    ChanConst<P>& k = out.k;
    k.a = av; k.cc = cc;
    k.alpha = alpha;
    k.beta = av * K;
    k.dalpha = -av * cc * dK;
    k.dbeta = av * dK;
    k.iS = iS; k.diS = diS;
    k.logS = log_(S);
    k.dlogS = dS * iS;
    constexpr int L = OPT_CHUNK_BYTES / (int)sizeof(P);
    k.gamma = -cc * k.beta;
    // Phi^(L/2) of the scaled recursion by repeated squaring of Phi = [[alpha,0],[gamma,alpha]]
    {
        P pa = k.alpha, pb = k.gamma;
        for (int h = 1; h < L / 4; h <<= 1) { pb = P(2) * pa * pb; pa = pa * pa; }
        k.aQ = pa; k.bQ = pb;
        pb = P(2) * pa * pb; pa = pa * pa;   // Phi^(L/2) = (Phi^(L/4))^2: keeps quarter/half/full powers consistent
        k.aH = pa; k.bH = pb;
    }
    P aL = k.aH * k.aH;                 // Phi^L = (Phi^(L/2))^2: keeps the half/full powers consistent
    P bL = P(2) * k.aH * k.bH;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        k.aL[i] = aL; k.bL[i] = bL;
        bL = P(2) * aL * bL;
        aL = aL * aL;
    }
    k.aW = aL; k.bW = bL;
    out.a_lane[0] = P(1);
    out.b_lane[0] = P(0);
    for (int l = 1; l < 32; ++l) {  // Phi^(L l) = Phi^L Phi^(L (l-1))
        out.a_lane[l] = k.aL[0] * out.a_lane[l - 1];
        out.b_lane[l] = k.aL[0] * out.b_lane[l - 1] + k.bL[0] * out.a_lane[l - 1];
    }
    // scaled state: mt = m / beta, dt = dm / dbeta  (dm' = alpha dm + dbeta e  =>  dt' = alpha dt + e)
    out.z0[0] = m / k.beta;
    out.z0[1] = (k.dbeta != P(0)) ? dm / k.dbeta : P(0);
    out.tsum[0] = sl; out.tsum[1] = sdl; out.tsum[2] = se; out.tsum[3] = sde; out.tsum[4] = sg;
    out.t_c = t;
    // frames after which the response to the state at their start has decayed below rounding:
    // alpha^W (and W alpha^(W-1)) negligible against 1 -> W = ln(eps_w) / ln(alpha), eps_w far below eps
    const double la = log((double)fmin(fmax(alpha, P(1e-30)), P(1)));
    const double lw = (sizeof(P) == 4) ? -32.0 : -64.0;  // ln(1.3e-14) / ln(1.6e-28)
    double w = (la < -1e-12) ? lw / la : 2.0e9;
    w = fmin(w + 64.0, 2.0e9);
    out.warm = (int)w;
}

template <class P> __device__ void diag_adam_body(const DiagOptArgs<P>& a, int j, bool first);

// One WARP-tile: 32 lanes x L frames of one channel, y[] = the lane's register-resident chunk (already
// centred).  Every warp owns a contiguous run of warp-tiles, so the whole evaluation needs no block-wide
// barrier: the only cross-lane traffic is the 5-step shuffle scan below.
//
// The recursion runs in SCALED variables mt = m / beta, dt = (dm/ds) / dbeta, which removes every
// multiply that is not fused:   mt' = alpha mt + y,   e = y + gamma mt  (gamma = -c beta),   dt' = alpha dt + e,
// i.e. z' = Phi z + (y, y) with Phi = [[alpha, 0], [gamma, alpha]]  (5 FMA per frame in phase 3).
// sum e^2 is unchanged and sum e dm = dbeta * sum e dt (applied once, in diag_adam_kernel).
// The chunk is processed as two independent half-chunks (two dependency chains in flight per lane):
// the zero-state responses of the halves are combined with Phi^(L/2), and the second half of phase 3
// starts from the exact mid-chunk state Phi^(L/2) z_in + z_a.
// (cm, cd) is the warp's carry: state at the first frame of this warp-tile on entry, of the next on exit.
template <class P, int L, bool FULL, bool ACC>
__device__ inline void diag_warp_tile(const P (&y)[L], int nvalid, const ChanConst<P>& k, P a_lane, P b_lane,
                                      P& cm, P& cd, double& E2, double& G) {
    constexpr int H = L / 2;
    const int lane = threadIdx.x & 31;
    const P alpha = k.alpha, gamma = k.gamma;
    // phase 1: zero-state responses.  U = sum alpha^(H-1-i) y_i (= mt),  W = sum alpha^(H-1-i) U_i
    P Ua = P(0), Wa = P(0), Ub = P(0), Wb = P(0);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        Wa = fma(alpha, Wa, Ua);
        Wb = fma(alpha, Wb, Ub);
        Ua = fma(alpha, Ua, y[i]);
        Ub = fma(alpha, Ub, y[H + i]);
    }
    const P zam = Ua, zad = fma(gamma, Wa, Ua);   // first half from a zero state: (mt, dt)
    const P zbm = Ub, zbd = fma(gamma, Wb, Ub);   // second half from a zero state
    const P aH = k.aH, bH = k.bH;                 // Phi^(L/2)
    P zm = fma(aH, zam, zbm);
    P zd = fma(aH, zad, fma(bH, zam, zbd));
    // warp inclusive scan with the closed-form powers of Phi
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int d = 1 << q;
        const P pm = __shfl_up_sync(0xffffffffu, zm, d);
        const P pd = __shfl_up_sync(0xffffffffu, zd, d);
        if (lane >= d) {
            zd = fma(k.aL[q], pd, fma(k.bL[q], pm, zd));
            zm = fma(k.aL[q], pm, zm);
        }
    }
    P em = __shfl_up_sync(0xffffffffu, zm, 1), ed = __shfl_up_sync(0xffffffffu, zd, 1);
    if (lane == 0) { em = P(0); ed = P(0); }
    const P tm = __shfl_sync(0xffffffffu, zm, 31), td = __shfl_sync(0xffffffffu, zd, 31);  // warp aggregate
    const P cm0 = cm, cd0 = cd;
    cm = fma(k.aW, cm0, tm);                       // carry for the next warp-tile: Phi^(32 L) c + aggregate
    cd = fma(k.aW, cd0, fma(k.bW, cm0, td));
    if (!ACC) return;
    // exact states at the start of the two half-chunks
    P m0 = fma(a_lane, cm0, em);
    P d0 = fma(a_lane, cd0, fma(b_lane, cm0, ed));
    P m1 = fma(aH, m0, zam);
    P d1 = fma(aH, d0, fma(bH, m0, zad));
    // phase 3 (5 FMA per frame)
    P e2a = P(0), ga = P(0), e2b = P(0), gb = P(0);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const P ea = fma(gamma, m0, y[i]);
        const P eb = fma(gamma, m1, y[H + i]);
        m0 = fma(alpha, m0, y[i]);
        m1 = fma(alpha, m1, y[H + i]);
        if (FULL || i < nvalid) { e2a = fma(ea, ea, e2a); ga = fma(ea, d0, ga); }
        if (FULL || H + i < nvalid) { e2b = fma(eb, eb, e2b); gb = fma(eb, d1, gb); }
        d0 = fma(alpha, d0, ea);
        d1 = fma(alpha, d1, eb);
    }
    E2 += (double)(e2a + e2b);
    G += (double)(ga + gb);
}

#ifndef EKS_EARLY_ISSUE
#define EKS_EARLY_ISSUE 1
#endif
#ifndef EKS_L2_PREFETCH
#define EKS_L2_PREFETCH 0   // warp-tiles of look-ahead for an L2 prefetch of the observation stream (0 = off).
                            // Measured on the c5 bench: 2 -> 35.4 ms, 4 -> 37.4 ms, 8 -> 45.0 ms against 32.4 ms without:
                            // the memory system is already saturated, extra requests only add contention
#endif
#ifndef EKS_FFMA2
#define EKS_FFMA2 2   // fp32: packed FFMA2 (sm_100) for the two half-chunk chains of diag_warp_tile.
                      // 0 = scalar FFMA (33.0 ms optimiser stage on the c5 bench), 1 = 4-byte cp.async into an interleaved
                      // SMEM layout so that pairs load directly (34.9 ms: the 4x LDGSTS count costs more than the MOVs it
                      // saves), 2 = 16-byte ring, pairs formed in registers (32.4 ms), 3 = same with four quarter-chunk
                      // chains (32.4 ms: the dependent-FMA depth is not the limiter either)
#endif

// fp32 variant of diag_warp_tile on PACKED pairs: y2[i] = (y[i], y[H + i]) holds one frame of each half-chunk, so
// the two independent dependency chains of the scalar version become the two lanes of one FFMA2 (Blackwell's
// packed fp32 FMA: two IEEE fused multiply-adds per issue slot, bit-identical to the scalar code).  Halves the
// floating-point instruction count of a kernel that is issue-bound (ncu: 67 % issue utilisation at 73 % of HBM peak).
template <int L, bool FULL, bool ACC>
__device__ inline void diag_warp_tile_f2(const float2 (&y2)[L / 2], int nvalid, const ChanConst<float>& k, float a_lane,
                                         float b_lane, float& cm, float& cd, double& E2, double& G) {
    constexpr int H = L / 2;
    const int lane = threadIdx.x & 31;
    const float alpha = k.alpha, gamma = k.gamma;
    const float2 al2 = make_float2(alpha, alpha), ga2 = make_float2(gamma, gamma);
    float2 U = make_float2(0.f, 0.f), W = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        W = __ffma2_rn(al2, W, U);
        U = __ffma2_rn(al2, U, y2[i]);
    }
    const float2 zd2 = __ffma2_rn(ga2, W, U);      // (zad, zbd)
    const float zam = U.x, zbm = U.y, zad = zd2.x, zbd = zd2.y;
    const float aH = k.aH, bH = k.bH;
    float zm = fmaf(aH, zam, zbm);
    float zd = fmaf(aH, zad, fmaf(bH, zam, zbd));
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int d = 1 << q;
        const float pm = __shfl_up_sync(0xffffffffu, zm, d);
        const float pd = __shfl_up_sync(0xffffffffu, zd, d);
        if (lane >= d) {
            zd = fmaf(k.aL[q], pd, fmaf(k.bL[q], pm, zd));
            zm = fmaf(k.aL[q], pm, zm);
        }
    }
    float em = __shfl_up_sync(0xffffffffu, zm, 1), ed = __shfl_up_sync(0xffffffffu, zd, 1);
    if (lane == 0) { em = 0.f; ed = 0.f; }
    const float tm = __shfl_sync(0xffffffffu, zm, 31), td = __shfl_sync(0xffffffffu, zd, 31);
    const float cm0 = cm, cd0 = cd;
    cm = fmaf(k.aW, cm0, tm);
    cd = fmaf(k.aW, cd0, fmaf(k.bW, cm0, td));
    if (!ACC) return;
    const float m0 = fmaf(a_lane, cm0, em);
    const float d0 = fmaf(a_lane, cd0, fmaf(b_lane, cm0, ed));
    float2 m = make_float2(m0, fmaf(aH, m0, zam));
    float2 dd = make_float2(d0, fmaf(aH, d0, fmaf(bH, m0, zad)));
    float2 e2 = make_float2(0.f, 0.f), gg = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float2 e = __ffma2_rn(ga2, m, y2[i]);
        m = __ffma2_rn(al2, m, y2[i]);
        if (FULL) {
            e2 = __ffma2_rn(e, e, e2);
            gg = __ffma2_rn(e, dd, gg);
        } else {
            if (i < nvalid) { e2.x = fmaf(e.x, e.x, e2.x); gg.x = fmaf(e.x, dd.x, gg.x); }
            if (H + i < nvalid) { e2.y = fmaf(e.y, e.y, e2.y); gg.y = fmaf(e.y, dd.y, gg.y); }
        }
        dd = __ffma2_rn(al2, dd, e);
    }
    E2 += (double)(e2.x + e2.y);
    G += (double)(gg.x + gg.y);
}

// Quarter-chunk version of diag_warp_tile_f2: FOUR independent chains of L/4 frames (two FFMA2 streams), which
// halves the dependent-FMA depth of phases 1 and 3 (the kernel is latency bound: ~50 % issue utilisation with six
// warps per scheduler).  ya[i] = (y[i], y[Q+i]), yb[i] = (y[2Q+i], y[3Q+i]).
template <int L, bool FULL, bool ACC>
__device__ inline void diag_warp_tile_f4(const float2 (&ya)[L / 4], const float2 (&yb)[L / 4], int nvalid,
                                         const ChanConst<float>& k, float a_lane, float b_lane, float& cm, float& cd,
                                         double& E2, double& G) {
    constexpr int Q = L / 4;
    const int lane = threadIdx.x & 31;
    const float alpha = k.alpha, gamma = k.gamma;
    const float2 al2 = make_float2(alpha, alpha), ga2 = make_float2(gamma, gamma);
    float2 Ua = make_float2(0.f, 0.f), Wa = Ua, Ub = Ua, Wb = Ua;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        Wa = __ffma2_rn(al2, Wa, Ua);
        Wb = __ffma2_rn(al2, Wb, Ub);
        Ua = __ffma2_rn(al2, Ua, ya[i]);
        Ub = __ffma2_rn(al2, Ub, yb[i]);
    }
    const float2 Da = __ffma2_rn(ga2, Wa, Ua), Db = __ffma2_rn(ga2, Wb, Ub);   // zero-state (mt, dt) of the quarters
    const float z0m = Ua.x, z1m = Ua.y, z2m = Ub.x, z3m = Ub.y;
    const float z0d = Da.x, z1d = Da.y, z2d = Db.x, z3d = Db.y;
    const float aQ = k.aQ, bQ = k.bQ, aH = k.aH, bH = k.bH;
    // halves: Phi^Q z0 + z1 and Phi^Q z2 + z3; chunk: Phi^(2Q) (first half) + second half
    const float h0m = fmaf(aQ, z0m, z1m), h0d = fmaf(aQ, z0d, fmaf(bQ, z0m, z1d));
    const float h1m = fmaf(aQ, z2m, z3m), h1d = fmaf(aQ, z2d, fmaf(bQ, z2m, z3d));
    float zm = fmaf(aH, h0m, h1m);
    float zd = fmaf(aH, h0d, fmaf(bH, h0m, h1d));
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int d = 1 << q;
        const float pm = __shfl_up_sync(0xffffffffu, zm, d);
        const float pd = __shfl_up_sync(0xffffffffu, zd, d);
        if (lane >= d) {
            zd = fmaf(k.aL[q], pd, fmaf(k.bL[q], pm, zd));
            zm = fmaf(k.aL[q], pm, zm);
        }
    }
    float em = __shfl_up_sync(0xffffffffu, zm, 1), ed = __shfl_up_sync(0xffffffffu, zd, 1);
    if (lane == 0) { em = 0.f; ed = 0.f; }
    const float tm = __shfl_sync(0xffffffffu, zm, 31), td = __shfl_sync(0xffffffffu, zd, 31);
    const float cm0 = cm, cd0 = cd;
    cm = fmaf(k.aW, cm0, tm);
    cd = fmaf(k.aW, cd0, fmaf(k.bW, cm0, td));
    if (!ACC) return;
    // exact states at the start of the four quarters
    const float m0 = fmaf(a_lane, cm0, em);
    const float d0 = fmaf(a_lane, cd0, fmaf(b_lane, cm0, ed));
    const float m1 = fmaf(aQ, m0, z0m), d1 = fmaf(aQ, d0, fmaf(bQ, m0, z0d));
    const float m2 = fmaf(aH, m0, h0m), d2 = fmaf(aH, d0, fmaf(bH, m0, h0d));
    const float m3 = fmaf(aQ, m2, z2m), d3 = fmaf(aQ, d2, fmaf(bQ, m2, z2d));
    float2 ma = make_float2(m0, m1), da = make_float2(d0, d1), mb = make_float2(m2, m3), db = make_float2(d2, d3);
    float2 e2a = make_float2(0.f, 0.f), ga = e2a, e2b = e2a, gb = e2a;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float2 ea = __ffma2_rn(ga2, ma, ya[i]);
        const float2 eb = __ffma2_rn(ga2, mb, yb[i]);
        ma = __ffma2_rn(al2, ma, ya[i]);
        mb = __ffma2_rn(al2, mb, yb[i]);
        if (FULL) {
            e2a = __ffma2_rn(ea, ea, e2a); ga = __ffma2_rn(ea, da, ga);
            e2b = __ffma2_rn(eb, eb, e2b); gb = __ffma2_rn(eb, db, gb);
        } else {
            if (i < nvalid) { e2a.x = fmaf(ea.x, ea.x, e2a.x); ga.x = fmaf(ea.x, da.x, ga.x); }
            if (Q + i < nvalid) { e2a.y = fmaf(ea.y, ea.y, e2a.y); ga.y = fmaf(ea.y, da.y, ga.y); }
            if (2 * Q + i < nvalid) { e2b.x = fmaf(eb.x, eb.x, e2b.x); gb.x = fmaf(eb.x, db.x, gb.x); }
            if (3 * Q + i < nvalid) { e2b.y = fmaf(eb.y, eb.y, e2b.y); gb.y = fmaf(eb.y, db.y, gb.y); }
        }
        da = __ffma2_rn(al2, da, ea);
        db = __ffma2_rn(al2, db, eb);
    }
    E2 += (double)((e2a.x + e2a.y) + (e2b.x + e2b.y));
    G += (double)((ga.x + ga.y) + (gb.x + gb.y));
}

// fp32 ring fill for the packed variant: 4-byte cp.async, element q of a lane's chunk lands in the interleaved slot
// (q mod H) * 2 + q / H, so that one 16-byte shared load delivers two (y[i], y[H + i]) register pairs.
// Instruction i moves chunk i: 32 consecutive frames = one coalesced 128-byte request, conflict-free in SMEM.
__device__ inline void warp_issue_tile_f2(unsigned char* stage, const float* __restrict__ plane, int t0, int e_min,
                                          int n, bool inner) {
    constexpr int L = OPT_CHUNK_BYTES / 4, H = L / 2;
    const int lane = threadIdx.x & 31;
    const int pos = ((lane & (H - 1)) << 1) | (lane / H);
    unsigned char* dst = stage + pos * 4;
    const float* src = plane + t0 + lane;
    if (inner) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i * OPT_PAD_BYTES);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src + i * L) : "memory");
        }
    } else {
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int e = t0 + i * L + lane;
            const int valid = (e < n && e >= e_min) ? 4 : 0;
            cp_async_4(dst + i * OPT_PAD_BYTES, plane + (valid > 0 ? e : 0), valid);
        }
    }
}

// per-warp ring stage: 32 padded chunks
constexpr int WRP_STAGE_BYTES = 32 * OPT_PAD_BYTES;

// Asynchronous copy of one warp-tile (32 x L frames starting at frame t0) into a warp's ring stage; frames
// outside [e_min, n) are zero-filled without touching memory.  Lane l fetches granules l, l+32, ...:
// every instruction is one fully coalesced 512-byte request.
template <class P>
__device__ inline void warp_issue_tile(unsigned char* stage, const P* __restrict__ plane, int t0, int e_min, int n,
                                       bool vec, bool inner) {
    constexpr int L = OPT_CHUNK_BYTES / (int)sizeof(P);
    constexpr int EPG = 16 / (int)sizeof(P);
    constexpr int GPC = OPT_CHUNK_BYTES / 16;
    const int lane = threadIdx.x & 31;
    if (vec && inner) {
        const int j0 = lane / GPC, q = lane % GPC;
        const P* src = plane + t0 + j0 * L + q * EPG;
        unsigned char* dst = stage + j0 * OPT_PAD_BYTES + q * 16;
        constexpr int CPI = 32 / GPC;  // chunks covered per instruction
#pragma unroll
        for (int i = 0; i < GPC; ++i) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i * CPI * OPT_PAD_BYTES);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i * CPI * L) : "memory");
        }
    } else if (vec) {
#pragma unroll
        for (int i = 0; i < GPC; ++i) {
            const int v = i * 32 + lane;
            const int j = v / GPC, q = v - j * GPC;
            const int e = t0 + j * L + q * EPG;
            int valid = max(0, min(EPG, n - e)) * (int)sizeof(P);
            if (e < e_min) valid = 0;  // e_min is chunk aligned: whole granules
            cp_async_16(stage + j * OPT_PAD_BYTES + q * 16, plane + (valid > 0 ? e : 0), valid);
        }
    } else {
#pragma unroll 4
        for (int i = 0; i < L; ++i) {
            const int v = i * 32 + lane;
            const int j = v / L, q = v - j * L;
            const int e = t0 + j * L + q;
            const int valid = (e < n && e >= e_min) ? (int)sizeof(P) : 0;
            if (sizeof(P) == 4) cp_async_4(stage + j * OPT_PAD_BYTES + q * 4, plane + (valid > 0 ? e : 0), valid);
            else cp_async_8(stage + j * OPT_PAD_BYTES + q * 8, plane + (valid > 0 ? e : 0), valid);
        }
    }
}

// ---- kernel A: one NLL(+d/ds) evaluation.  grid = (nseg, 2 * B): CTA (k, 2b+c) handles segment k of
// channel c of sequence b, and each of its 8 warps an independent contiguous run of warp-tiles (32 lanes x
// L frames) inside it.  A run starts from the exact state at t_c (first run, or slow forgetting) or from a
// zero state `warm` frames earlier, which is exact to rounding because the steady-state filter forgets its
// initial state geometrically (alpha^warm < 1e-14 / 1e-28).  Warps never synchronise with each other until
// the final reduction; each streams its own 2-stage cp.async ring.
template <class P>
__global__ void __launch_bounds__(OPT_NT, EKS_OPT_MINBLOCKS) diag_nll_kernel(const __grid_constant__ DiagOptArgs<P> a) {
    __shared__ ChanConst<P> shk;
    __shared__ double red[OPT_NW][2];
    extern __shared__ __align__(16) unsigned char ring[];
    constexpr int L = OPT_CHUNK_BYTES / (int)sizeof(P);
    constexpr int WT = 32 * L;  // frames per warp-tile
    using V = typename DiagTraits<P>::vec_t;
    constexpr int VW = DiagTraits<P>::VW;
    const int seg = blockIdx.x, b = blockIdx.y >> 1, c = blockIdx.y & 1;
    const int blk = a.seq_block[b];
    if (blk < a.blk_lo || blk >= a.blk_hi || a.bstate[blk].done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ChanState<P>& cs = a.cstate[(long long)b * 2 + c];
    double* part = a.partials + (((long long)b * 2 + c) * a.nseg + seg) * 2;
    const int t_c = cs.t_c;
    if (threadIdx.x == 0) shk = cs.k;
    __syncthreads();
    // this warp's run of warp-tiles
    const int nwt = (a.n - t_c + WT - 1) / WT;
    const int nrun = a.nseg * OPT_NW;
    const int wpr = (nwt + nrun - 1) / nrun;
    const int run = seg * OPT_NW + warp;
    const int wt_lo = min(nwt, run * wpr), wt_hi = min(nwt, wt_lo + wpr);
    double E2 = 0, G = 0;
    if (wt_lo < wt_hi) {
        // warm-up: whole chunks, starting `warm` frames before the run
        const int warm_chunks = (cs.warm + L - 1) / L;
        const long long e_min_ll = (long long)t_c + (long long)wt_lo * WT - (long long)warm_chunks * L;
        int first_wt, e_min;
        P cm, cd;
        if (wt_lo == 0 || e_min_ll <= (long long)t_c) {
            first_wt = 0; e_min = 0; cm = cs.z0[0]; cd = cs.z0[1];
        } else {
            e_min = (int)e_min_ll;
            first_wt = (e_min - t_c) / WT;
            cm = P(0); cd = P(0);
        }
        const P* yc = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.t_begin + a.y.chan_off[c];
        const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
        const bool vec = (reinterpret_cast<uintptr_t>(yc + t_c) & 15) == 0;
        const P a_lane = cs.a_lane[lane], b_lane = cs.b_lane[lane];
        unsigned char* wring = ring + warp * (OPT_STAGES * WRP_STAGE_BYTES);
        const int nt = wt_hi - first_wt;
        auto issue = [&](int stage, int wt) {
            const int t0i = t_c + wt * WT;
            if constexpr (sizeof(P) == 4 && EKS_FFMA2 == 1)
                warp_issue_tile_f2(wring + stage * WRP_STAGE_BYTES, reinterpret_cast<const float*>(yc), t0i, e_min,
                                   a.n, t0i >= e_min && t0i + WT <= a.n);
            else
                warp_issue_tile<P>(wring + stage * WRP_STAGE_BYTES, yc, t0i, e_min, a.n, vec,
                                   t0i >= e_min && t0i + WT <= a.n);
#if EKS_L2_PREFETCH > 0
            // pull a later warp-tile of this run into L2 so that the ring is filled at L2 latency: the SMEM ring
            // (all of the SM's shared memory at 3 CTAs x 8 warps x 2 stages) cannot hold a DRAM latency of bytes
            const int wtp = wt + EKS_L2_PREFETCH;
            if (lane == 0 && vec && wtp < wt_hi) {
                const P* pa = yc + t_c + (long long)wtp * WT;
                const int bytes = (int)min((long long)WT, (long long)a.n - (t_c + (long long)wtp * WT)) * (int)sizeof(P) & ~15;
                if (bytes > 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(pa), "r"(bytes) : "memory");
            }
#endif
        };
        // EKS_EARLY_ISSUE: a stage is free as soon as its tile sits in registers, i.e. BEFORE the arithmetic on that
        // tile.  Refilling it there (tile it+2) keeps two tiles outstanding per warp during the arithmetic instead of
        // one -- twice the bytes in flight from the same shared memory (the ring already fills the SM).
        constexpr bool EARLY = (EKS_EARLY_ISSUE != 0) && (OPT_STAGES == 2) && !(sizeof(P) == 4 && EKS_FFMA2 == 1);
#pragma unroll
        for (int st = 0; st < (EARLY ? OPT_STAGES : OPT_STAGES - 1); ++st) {
            if (st < nt) issue(st, first_wt + st);
            cp_async_commit();
        }
        for (int it = 0; it < nt; ++it) {
            if (!EARLY) {
                const int nx = it + OPT_STAGES - 1;
                // the stage about to be refilled was read in iteration it-1; make sure every lane is done with it
                __syncwarp();
                if (nx < nt) issue(nx % OPT_STAGES, first_wt + nx);
                cp_async_commit();
            }
            cp_async_wait<OPT_STAGES - 1>();
            __syncwarp();  // tile `it` has landed for every lane of this warp
            const unsigned char* mine = wring + (it % OPT_STAGES) * WRP_STAGE_BYTES + lane * OPT_PAD_BYTES;
            const int t0 = t_c + (first_wt + it) * WT;
            const int cstart = t0 + lane * L;
            const bool inner = (t0 >= e_min) && (t0 + WT <= a.n);  // warp-uniform: no masked frames
            const bool acc = (first_wt + it) >= wt_lo;
            if constexpr (sizeof(P) == 4 && EKS_FFMA2 == 1) {
                constexpr int H = L / 2;
                float2 y2[H];
                const float2 nm2 = make_float2(-(float)mean, -(float)mean);
                if (inner) {
#pragma unroll
                    for (int i = 0; i < L / 4; ++i) {
                        const float4 v = *reinterpret_cast<const float4*>(mine + i * 16);
                        y2[2 * i] = __fadd2_rn(make_float2(v.x, v.y), nm2);
                        y2[2 * i + 1] = __fadd2_rn(make_float2(v.z, v.w), nm2);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < L / 4; ++i) {
                        const float4 v = *reinterpret_cast<const float4*>(mine + i * 16);
                        const float e[4] = {v.x, v.y, v.z, v.w};
                        float o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int fr = cstart + (q & 1) * H + 2 * i + (q >> 1);
                            o[q] = (fr >= e_min && fr < a.n) ? e[q] - (float)mean : 0.f;
                        }
                        y2[2 * i] = make_float2(o[0], o[1]);
                        y2[2 * i + 1] = make_float2(o[2], o[3]);
                    }
                }
                const ChanConst<float>& kk = reinterpret_cast<const ChanConst<float>&>(shk);
                float fcm = (float)cm, fcd = (float)cd;
                if (!acc) diag_warp_tile_f2<L, true, false>(y2, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else if (inner) diag_warp_tile_f2<L, true, true>(y2, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else diag_warp_tile_f2<L, false, true>(y2, max(0, min(L, a.n - cstart)), kk, (float)a_lane, (float)b_lane,
                                                       fcm, fcd, E2, G);
                cm = (P)fcm; cd = (P)fcd;
                continue;
            }
            P y[L];
            if (inner) {
#pragma unroll
                for (int i = 0; i < L / VW; ++i) {
                    const V v = *reinterpret_cast<const V*>(mine + i * 16);
                    const P* e = reinterpret_cast<const P*>(&v);
#pragma unroll
                    for (int q = 0; q < VW; ++q) y[i * VW + q] = e[q] - mean;
                }
            } else {
#pragma unroll
                for (int i = 0; i < L / VW; ++i) {
                    const V v = *reinterpret_cast<const V*>(mine + i * 16);
                    const P* e = reinterpret_cast<const P*>(&v);
#pragma unroll
                    for (int q = 0; q < VW; ++q) {
                        const int fr = cstart + i * VW + q;
                        y[i * VW + q] = (fr >= e_min && fr < a.n) ? e[q] - mean : P(0);
                    }
                }
            }
            if (EARLY) {   // tile `it` is in registers: refill its stage with tile it+2 before the arithmetic
                __syncwarp();
                if (it + 2 < nt) issue(it % OPT_STAGES, first_wt + it + 2);
                cp_async_commit();
            }
            if constexpr (sizeof(P) == 4 && EKS_FFMA2 == 3) {   // 16-byte ring, four quarter-chunk chains
                constexpr int Q = L / 4;
                float2 ya[Q], yb[Q];
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    ya[i] = make_float2((float)y[i], (float)y[Q + i]);
                    yb[i] = make_float2((float)y[2 * Q + i], (float)y[3 * Q + i]);
                }
                const ChanConst<float>& kk = reinterpret_cast<const ChanConst<float>&>(shk);
                float fcm = (float)cm, fcd = (float)cd;
                if (!acc) diag_warp_tile_f4<L, true, false>(ya, yb, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else if (inner) diag_warp_tile_f4<L, true, true>(ya, yb, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else diag_warp_tile_f4<L, false, true>(ya, yb, max(0, min(L, a.n - cstart)), kk, (float)a_lane,
                                                       (float)b_lane, fcm, fcd, E2, G);
                cm = (P)fcm; cd = (P)fcd;
                continue;
            }
            if constexpr (sizeof(P) == 4 && EKS_FFMA2 == 2) {   // 16-byte ring, pairs formed in registers
                constexpr int H = L / 2;
                float2 y2[H];
#pragma unroll
                for (int i = 0; i < H; ++i) y2[i] = make_float2((float)y[i], (float)y[H + i]);
                const ChanConst<float>& kk = reinterpret_cast<const ChanConst<float>&>(shk);
                float fcm = (float)cm, fcd = (float)cd;
                if (!acc) diag_warp_tile_f2<L, true, false>(y2, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else if (inner) diag_warp_tile_f2<L, true, true>(y2, L, kk, (float)a_lane, (float)b_lane, fcm, fcd, E2, G);
                else diag_warp_tile_f2<L, false, true>(y2, max(0, min(L, a.n - cstart)), kk, (float)a_lane, (float)b_lane,
                                                       fcm, fcd, E2, G);
                cm = (P)fcm; cd = (P)fcd;
                continue;
            }
            if (!acc) diag_warp_tile<P, L, true, false>(y, L, shk, a_lane, b_lane, cm, cd, E2, G);
            else if (inner) diag_warp_tile<P, L, true, true>(y, L, shk, a_lane, b_lane, cm, cd, E2, G);
            else diag_warp_tile<P, L, false, true>(y, max(0, min(L, a.n - cstart)), shk, a_lane, b_lane, cm, cd, E2, G);
        }
        cp_async_wait<0>();
    }
    E2 = warp_sum(E2);
    G = warp_sum(G);
    if (lane == 0) { red[warp][0] = E2; red[warp][1] = G; }
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        double te = 0, tg = 0;
        for (int w = 0; w < OPT_NW; ++w) { te += red[w][0]; tg += red[w][1]; }
        part[0] = te;
        part[1] = tg;
        // the CTA that completes its block's evaluation takes the Adam step and prepares the next evaluation
        // while other blocks are still streaming (no separate launch, no idle gap between evaluations)
        __threadfence();
        const int expected = a.nseg * 2 * (a.block_off[blk + 1] - a.block_off[blk]);
        const int prev = atomicAdd(&a.block_counter[blk], 1);
        is_last = (prev + 1 == expected);
        if (is_last) a.block_counter[blk] = 0;
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        diag_adam_body<P>(a, blk, false);
    }
}

// ---- Adam step for one block, executed by ONE WARP.  Consumes the partial sums of the evaluation that just
// finished (stop rule of eks/core.py:654-681), then prepares the next one (new s; per-channel transient and
// steady-state constants).  first = true: initialise instead of consuming.
template <class P>
__device__ void diag_adam_body(const DiagOptArgs<P>& a, int j, bool first) {
    const int lane = threadIdx.x & 31;
    BlockState<P>& bs = a.bstate[j];
    const int m_lo = a.block_off[j], m_hi = a.block_off[j + 1];
    if (first) {
        if (lane == 0) {
            adam_init(bs.adam, a.s_log0[j]);
            bs.done = (a.cap <= 0);
            if (bs.done) {
                a.s_log_out[j] = bs.adam.s_log; a.last_loss_out[j] = bs.adam.prev; a.iters_out[j] = 0;
                atomicSub(a.n_active, 1);
            }
        }
    } else {
        if (bs.done) return;
        if (lane == 0) {
            const double HALF_LOG2PI = 0.91893853320467274178;
            P loss = P(0), grad = P(0);
            for (int mi = m_lo; mi < m_hi; ++mi) {
                const int b = a.members[mi];
                double nll = 0, dnll = 0;
                for (int c = 0; c < 2; ++c) {
                    const ChanState<P>& cs = a.cstate[(long long)b * 2 + c];
                    const double* part = a.partials + ((long long)b * 2 + c) * a.nseg * 2;
                    double te = 0, tg = 0;
                    // written by other CTAs of this launch: read through L2 (bypass this SM's L1)
                    for (int q = 0; q < a.nseg; ++q) { te += __ldcg(part + 2 * q); tg += __ldcg(part + 2 * q + 1); }
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
                grad += g * bs.dsdlog;
            }
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
                atomicSub(a.n_active, 1);
            }
        }
    }
    __syncwarp();
    if (bs.done) return;
    if (lane == 0) {
        P dsdlog;
        bs.s = adam_current_s(bs.adam, a.lo, a.hi, &dsdlog);
        bs.dsdlog = dsdlog;
    }
    __syncwarp();
    const P s = bs.s;
    const int npair = (m_hi - m_lo) * 2;  // (member, channel) pairs, one per lane
    for (int p0 = 0; p0 < npair; p0 += 32) {
        const int p = p0 + lane;
        if (p < npair) {
            const int b = a.members[m_lo + (p >> 1)], c = p & 1;
            diag_transient<P>(a, b, c, s, a.cstate[(long long)b * 2 + c]);
        }
    }
}

// ---- kernel B: initialisation (one warp per block): Adam state + the first transient
template <class P>
__global__ void __launch_bounds__(32) diag_adam_kernel(const __grid_constant__ DiagOptArgs<P> a) {
    diag_adam_body<P>(a, blockIdx.x, true);
}

// sequence -> block index table + active-block counter
__global__ void diag_seq_block_kernel(int n_blocks, const int* __restrict__ block_off, const int* __restrict__ members,
                                      int* __restrict__ seq_block, int* __restrict__ n_active) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) *n_active = n_blocks;
    if (j >= n_blocks) return;
    for (int mi = block_off[j]; mi < block_off[j + 1]; ++mi) seq_block[members[mi]] = j;
}

static int diag_nseg(int dtype, int n, int B) {
    // Work unit = one warp's run of warp-tiles (32 lanes x L frames); a CTA holds 8 runs.  Each run pays a
    // warm-up of ~1 warp-tile, and the grid (nseg x 2B CTAs) is executed in waves of (SMs x 3) resident CTAs.
    // Pick the segment count that maximises  wave efficiency x useful fraction of a run.
    static int slots = 0;
    if (slots == 0) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = sms * EKS_OPT_MINBLOCKS;
    }
    const int L = OPT_CHUNK_BYTES / (dtype == EKS_F32 ? 4 : 8);
    const int nwt = (n + 32 * L - 1) / (32 * L);
    int best = 1;
    double best_eff = -1.0;
    for (int nseg = 1; nseg <= OPT_NSEG_MAX; ++nseg) {
        const int run = (nwt + nseg * OPT_NW - 1) / (nseg * OPT_NW);  // warp-tiles per run
        if (run < 1 || (nseg > 1 && run < 2)) break;
        const double waves = (double)nseg * 2.0 * B / slots;
        const double wave_eff = waves / ceil(waves);
        const double eff = wave_eff * run / (run + 1.0);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = nseg; }
    }
    return best;
}

size_t diag_optimize_workspace_bytes(int dtype, int n_blocks, int B, int T) {
    const size_t real = dtype == EKS_F32 ? 4 : 8;
    (void)real;
    const int nseg = OPT_NSEG_MAX;  // upper bound of diag_nseg
    size_t bytes = 256;
    bytes += (size_t)n_blocks * 128;                   // BlockState
    bytes += (size_t)B * 2 * 1024;                     // ChanState (generous bound)
    bytes += (size_t)B * 2 * nseg * 2 * sizeof(double);
    bytes += (size_t)B * sizeof(int) + 256;
    bytes += (size_t)n_blocks * sizeof(int) + 256;
    return bytes;
}

template <class P>
static int diag_optimize_run(DiagOptArgs<P>& a, void* workspace, size_t workspace_bytes, int dtype, int T,
                             cudaStream_t st) {
    static_assert(sizeof(BlockState<P>) <= 128 && sizeof(ChanState<P>) <= 1024, "workspace bound");
    a.nseg = diag_nseg(dtype, a.n, a.B);
    EKS_REQUIRE(workspace && workspace_bytes >= diag_optimize_workspace_bytes(dtype, a.n_blocks, a.B, T),
                "optimize_s: workspace too small");
    unsigned char* w = (unsigned char*)workspace;
    a.n_active = (int*)w; w += 256;
    a.bstate = (BlockState<P>*)w; w += (size_t)a.n_blocks * 128;
    a.cstate = (ChanState<P>*)w; w += (size_t)a.B * 2 * 1024;
    a.partials = (double*)w; w += (size_t)a.B * 2 * a.nseg * 2 * sizeof(double);
    int* seq_block = (int*)w; w += ((size_t)a.B * sizeof(int) + 255) / 256 * 256;
    a.seq_block = seq_block;
    a.block_counter = (int*)w;
    cudaMemsetAsync(a.block_counter, 0, (size_t)a.n_blocks * sizeof(int), st);
    const int smem = OPT_NW * OPT_STAGES * WRP_STAGE_BYTES;
    cudaError_t e = cudaFuncSetAttribute(diag_nll_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        set_error("diag_nll_kernel: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    cudaMemsetAsync(seq_block, 0xFF, (size_t)a.B * sizeof(int), st);  // -1: sequence belongs to no block
    diag_seq_block_kernel<<<(a.n_blocks + 127) / 128, 128, 0, st>>>(a.n_blocks, a.block_off, a.members, seq_block,
                                                                    a.n_active);
    // One initialisation launch, then ONE streaming launch per evaluation; the Adam step between evaluations
    // runs inside the streaming kernel (last CTA of each block).  The loop is unrolled on the stream without
    // host synchronisation: finished blocks make their CTAs exit immediately.
    const dim3 grid(a.nseg, 2 * a.B);
    a.blk_lo = 0; a.blk_hi = a.n_blocks;
    diag_adam_kernel<P><<<a.n_blocks, 32, 0, st>>>(a);
    // The blocks are independent optimisation problems, but launches on one stream serialise them: every evaluation
    // would end with a drain of the whole GPU (measured: a full launch streams 5.5 TB/s on its own, the loop
    // averaged 4.8 TB/s).  The blocks are therefore split over EKS_OPT_STREAMS internal streams whose launches
    // overlap: while one group's evaluation drains, the other group's fills the SMs.
    static int n_streams_dev[64];
    static cudaStream_t hs_dev[64][4];
    static bool hs_init[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!hs_init[dev]) {   // internal streams belong to the device that is current at first use (one table per device)
        const char* e2 = getenv("EKS_OPT_STREAMS");
        int nsd = e2 ? atoi(e2) : 2;
        if (nsd < 1) nsd = 1;
        if (nsd > 4) nsd = 4;
        for (int i = 0; i < nsd; ++i)
            if (cudaStreamCreateWithFlags(&hs_dev[dev][i], cudaStreamNonBlocking) != cudaSuccess) { nsd = 1; break; }
        n_streams_dev[dev] = nsd;
        hs_init[dev] = true;
    }
    const int n_streams = n_streams_dev[dev];
    cudaStream_t* hs = hs_dev[dev];
    // small problems (less than two waves of CTAs per evaluation) are launch bound: a second stream only doubles
    // the number of no-op launches after convergence
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool big = (long long)a.nseg * 2 * a.B >= 2LL * sms * EKS_OPT_MINBLOCKS;
    const int ns = (a.n_blocks >= 2 * n_streams && big) ? n_streams : 1;
    if (ns == 1) {
        for (int it = 0; it < a.cap; ++it) diag_nll_kernel<P><<<grid, OPT_NT, smem, st>>>(a);
        note_launches(2 + a.cap);
        return check_launch("diag optimise kernels");
    }
    cudaEvent_t fork, join[4];
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    cudaEventRecord(fork, st);
    DiagOptArgs<P> ai[4];
    for (int i = 0; i < ns; ++i) {
        cudaStreamWaitEvent(hs[i], fork, 0);
        ai[i] = a;
        ai[i].blk_lo = (int)((long long)a.n_blocks * i / ns);
        ai[i].blk_hi = (int)((long long)a.n_blocks * (i + 1) / ns);
    }
    for (int it = 0; it < a.cap; ++it)
        for (int i = 0; i < ns; ++i) diag_nll_kernel<P><<<grid, OPT_NT, smem, hs[i]>>>(ai[i]);
    for (int i = 0; i < ns; ++i) {
        cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming);
        cudaEventRecord(join[i], hs[i]);
    }
    for (int i = 0; i < ns; ++i) { cudaStreamWaitEvent(st, join[i], 0); cudaEventDestroy(join[i]); }
    cudaEventDestroy(fork);
    note_launches(2 + ns * a.cap);
    return check_launch("diag optimise kernels");
}

int diag_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q, const void* C,
                  const void* y_base, long long y_seq_stride, const long long* y_off, const void* ymean,
                  const void* Rconst, int t_begin, int n, int n_blocks, const int* block_off, const int* members,
                  const void* s_log0, double lr, double lo, double hi, double tol, int cap, void* s_log_out,
                  void* last_loss_out, int* iters_out, void* trace, int trace_cap, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
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
    return diag_optimize_run<PT>(a, workspace, workspace_bytes, dtype, T, st);
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
#ifndef EKS_SMOOTH_MINBLOCKS
// 3 CTAs per SM (<= 85 registers): 148 x 3 = 444 resident CTAs, so that the 2 x (sessions x keypoints)
// persistent CTAs of a typical batch fit in ONE wave (a second, nearly empty wave would double the time)
#define EKS_SMOOTH_MINBLOCKS 3
#endif

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
    int latent_out;        // 1: write the latent smoothed moments (m_s, P_s) instead of C m + mean, C V C^T
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
__global__ void __launch_bounds__(DIAG_NT, EKS_SMOOTH_MINBLOCKS) diag_filter_kernel(const __grid_constant__ DiagSmoothArgs<P> a) {
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
            // gain with the 1e-9 boost of psd_solve; P_f = P - K S K and 1 - K c in their cancellation-free
            // (algebraically identical) forms
            const P S = fma(c2, Pv, r[i]);
            const P iSb = P(1) / (S + P(1e-9));
            const P K = Pv * cc * iSb;
            const P Pfi = Pv * iSb * (r[i] + P(1e-9) * (P(1) + cc * K));
            const P alpha = av * iSb * (r[i] + P(1e-9)), beta = av * K;
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
__global__ void __launch_bounds__(DIAG_NT, EKS_SMOOTH_MINBLOCKS) diag_rts_kernel(const __grid_constant__ DiagSmoothArgs<P> a) {
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
            // G = a P_f / (S_p + 1e-9);  offsets (1 - G a) m_f and P_f - G^2 S_p in cancellation-free form
            const P Sp = fma(a2, Pf[i], q);
            const P iSpb = P(1) / (Sp + P(1e-9));
            P g = av * Pf[i] * iSpb;
            P om = mf[i] * iSpb * (q + P(1e-9));
            P oP = Pf[i] * iSpb * (q + P(1e-9) * (P(1) + av * g));
            if (start + i >= a.T - 1) { g = P(0); om = mf[i]; oP = Pf[i]; }  // last frame: smoothed = filtered
            G[i] = g;
            mf[i] = om;   // keep the offsets: phase 3 reuses them
            Pf[i] = oP;
            bm = fma(g, bm, om);
            bP = fma(g * g, bP, oP);
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
            const P g = G[i];
            ms = fma(g, ms, mf[i]);
            Ps = fma(g * g, Ps, Pf[i]);
            mf[i] = a.latent_out ? ms : fma(cc, ms, mean);  // x = C m + mean   (singlecam_smoother.py:190-197)
            Pf[i] = a.latent_out ? Ps : c2 * Ps;            // diag(C V C^T)    (singlecam_smoother.py:191, 210-211)
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
                              int latent_out, void* workspace, cudaStream_t st) {
    DiagSmoothArgs<P> a;
    a.B = B; a.T = T;
    a.m0 = (const P*)m0; a.S0 = (const P*)S0; a.A = (const P*)A; a.Q = (const P*)Q; a.C = (const P*)C;
    a.y = y; a.var = var; a.ymean = (const P*)ymean; a.s = (const P*)s;
    // keep the workspace planes 16-byte aligned when T allows
    a.mf = (P*)workspace;
    a.Pf = a.mf + (size_t)B * 2 * T;
    a.out = (P*)out; a.out_seq_stride = out_seq_stride;
    for (int i = 0; i < 4; ++i) a.out_off[i] = out_off[i];
    a.latent_out = latent_out;
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
                               const long long* out_off, int latent_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
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
                                         latent_out, workspace, st);
    return diag_smooth_launch<double>(B, T, m0, S0, A, Q, C, y, var, ymean, s, out, out_seq_stride, out_off,
                                      latent_out, workspace, st);
}
