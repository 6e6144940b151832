// diag_stream.cuh -- streaming ("one pass over the observations") evaluation of the constant-R loss of DECOUPLED
// models (D == O == 2, diagonal A, C, Q, S0: the single-camera EKS model, eks/singlecam_smoother.py:246-284), shared
// by the per-evaluation kernel of diag.cu and by the persistent optimiser of diag_lag.cu (its fallback path).
//
//   * loss path has CONSTANT R (core.py:702-709) => the variance recursion is data independent.  diag_transient runs
//     it sequentially (with its s-sensitivity) until it reaches its floating point fixed point ("transient", a few
//     tens of frames), accumulating the NLL there exactly as the sequential filter does;
//   * for the remaining frames the gain is constant and the predicted mean m_t and its sensitivity dm_t/ds obey a
//     constant-coefficient linear recurrence
//         z_{t+1} = Phi z_t + b y_t,  z = (m, dm),  Phi = [[alpha,0],[dalpha,alpha]], b = (beta,dbeta)
//     which diag_stream_cta evaluates EXACTLY in parallel: each lane owns L consecutive frames in registers, computes
//     its zero-state response (2 FMAs / frame), a warp scan with the closed-form powers Phi^(L 2^k) hands every lane
//     its carry-in, and a second register pass accumulates sum e^2 and sum e dm.
//   * NLL = n/2 log 2pi + 1/2 sum log S_t + 1/2 sum e_t^2 / S_t, its derivative likewise; partial sums are kept in
//     fp64 so that the reference's relative-tolerance stop rule stays meaningful at 10^6 frames in float32 mode.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "ekf_generic.cuh"

namespace eks {

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
#ifndef EKS_EARLY_ISSUE
#define EKS_EARLY_ISSUE 1
#endif

// ---- one streaming evaluation of (sum e^2, sum e dt) for segment `seg` (of `nseg`) of channel c of sequence b, by
// ONE CTA of OPT_NT threads.  Each of its OPT_NW warps owns an independent contiguous run of warp-tiles (32 lanes x
// L frames).  A run starts from the exact state at t_c (first run, or slow forgetting) or from a zero state `warm`
// frames earlier, which is exact to rounding because the steady-state filter forgets its initial state geometrically
// (alpha^warm < 1e-14 / 1e-28).  Warps never synchronise with each other until the final reduction; each streams its
// own 2-stage cp.async ring (ring: OPT_NW * OPT_STAGES * WRP_STAGE_BYTES bytes of dynamic shared memory).
// Measured history of this loop (fp32, c5 bench, optimiser stage): scalar FFMA 33.0 ms; packed FFMA2 with pairs
// formed in registers 32.4 ms (kept); 4-byte cp.async into an interleaved layout 34.9 ms; four quarter-chunk chains
// 32.4 ms; cp.async.bulk L2 prefetch 2/4/8 tiles ahead 35.4 / 37.4 / 45.0 ms (all rejected, removed).
// On return thread 0 holds the CTA's sums in (te, tg); contains __syncthreads (call from uniform control flow).
template <class P>
__device__ inline void diag_stream_cta(const DiagOptArgs<P>& a, int b, int c, int seg, int nseg, unsigned char* ring,
                                       ChanConst<P>& shk, double (*red)[2], double& te, double& tg) {
    constexpr int L = OPT_CHUNK_BYTES / (int)sizeof(P);
    constexpr int WT = 32 * L;  // frames per warp-tile
    using V = typename DiagTraits<P>::vec_t;
    constexpr int VW = DiagTraits<P>::VW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ChanState<P>& cs = a.cstate[(long long)b * 2 + c];
    const int t_c = cs.t_c;
    __syncthreads();   // previous users of shk / red are done
    if (threadIdx.x == 0) shk = cs.k;
    __syncthreads();
    // this warp's run of warp-tiles
    const int nwt = (a.n - t_c + WT - 1) / WT;
    const int nrun = nseg * OPT_NW;
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
            warp_issue_tile<P>(wring + stage * WRP_STAGE_BYTES, yc, t0i, e_min, a.n, vec,
                               t0i >= e_min && t0i + WT <= a.n);
        };
        // EKS_EARLY_ISSUE: a stage is free as soon as its tile sits in registers, i.e. BEFORE the arithmetic on that
        // tile.  Refilling it there (tile it+2) keeps two tiles outstanding per warp during the arithmetic instead of
        // one -- twice the bytes in flight from the same shared memory (the ring already fills the SM).
        constexpr bool EARLY = (EKS_EARLY_ISSUE != 0) && (OPT_STAGES == 2);
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
            if constexpr (sizeof(P) == 4) {   // packed FFMA2: the two half-chunk chains are the lanes of one instruction
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
            } else {
                if (!acc) diag_warp_tile<P, L, true, false>(y, L, shk, a_lane, b_lane, cm, cd, E2, G);
                else if (inner) diag_warp_tile<P, L, true, true>(y, L, shk, a_lane, b_lane, cm, cd, E2, G);
                else diag_warp_tile<P, L, false, true>(y, max(0, min(L, a.n - cstart)), shk, a_lane, b_lane, cm, cd, E2, G);
            }
        }
        cp_async_wait<0>();
    }
    E2 = warp_sum(E2);
    G = warp_sum(G);
    if (lane == 0) { red[warp][0] = E2; red[warp][1] = G; }
    __syncthreads();
    te = 0; tg = 0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < OPT_NW; ++w) { te += red[w][0]; tg += red[w][1]; }
    }
}

}  // namespace eks
