// diag_smooth.cu -- the final filter + RTS smoother pass of DECOUPLED models (D == O == 2, diagonal A, C, Q, S0: the
// single-camera EKS model) with time-varying R_t, fused with the reprojection epilogue.
#include <cstdlib>
#include "common.cuh"
#include "ekf_generic.cuh"
#include "diag.cuh"
#include "diag_stream.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int DIAG_NT = 256;
constexpr int DIAG_NW = DIAG_NT / 32;

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
    const int* only_flagged;   // [B] or null: the exact kernels process only the sequences flagged by the fused kernel
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
    if (a.only_flagged && !a.only_flagged[b]) return;
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
    if (a.only_flagged && !a.only_flagged[b]) return;
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

// =====================================================================================================
// FUSED, time-segmented final pass: one kernel, y and R_t read once, outputs written once, no filtered
// moments in HBM.  grid = (segments, 2B); CTA (g, 2b+c) owns the frames [g SEG, (g+1) SEG) of channel c of
// sequence b and loads ONE register tile (DIAG_NT x L frames) = its frames + HALO frames before + HALO after:
//   forward : the tile body of diag_filter_kernel from the prior (m0, S0) at the window start -- exact for
//             the first segment; for the others the filter forgets its start state geometrically and the
//             HALO frames before the segment absorb it;
//   backward: the tile body of diag_rts_kernel on the filtered moments still in registers, started at the
//             window end with "smoothed = filtered" -- exact for the last segment; for the others the RTS
//             recursion forgets its start state over the HALO frames after the segment.
// This is VERIFIED, not assumed: with P_t >= q the forward contraction per frame is bounded by
// a r_t / (c^2 q + r_t) and the backward one (G_t) by a r_t / (a^2 r_t + c^2 q), both computable from the
// data; the CTA sums their logarithms over its two halos and flags the sequence (fail[b]) unless both
// products are below the rounding level of the working precision (also flagged: any non-finite
// observation or variance, which the exact kernels must propagate over the whole sequence).  Flagged
// sequences are redone by the exact scan kernels above (diag_filter_kernel + diag_rts_kernel, which
// return immediately for the others).  HBM traffic: 16 bytes per channel-frame x (1 + 2 HALO / SEG).
// =====================================================================================================
// reciprocal for the fused pass: float32 uses the hardware approximation (MUFU.RCP, <= 1 ulp) -- the IEEE division
// sequence is ~8 instructions, three per frame -- float64 the exact division
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }

template <class P>
struct FusedShared {
    Mob<P> magg[DIAG_NW];
    P aagg[DIAG_NW][2];
    P bagg[DIAG_NW][3];
    float chk[DIAG_NW][2];
};

template <class P>
__global__ void __launch_bounds__(DIAG_NT, EKS_SMOOTH_MINBLOCKS)
diag_smooth_fused_kernel(const __grid_constant__ DiagSmoothArgs<P> a, int seg, int halo, float logtol, int* __restrict__ fail) {
    __shared__ FusedShared<P> sh;
    constexpr int L = DiagTraits<P>::L;
    constexpr int TILE = DIAG_NT * L;
    const int b = blockIdx.y >> 1, c = blockIdx.y & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u0 = blockIdx.x * seg, u1 = min(a.T, u0 + seg);      // frames this CTA writes
    const int ws = max(0, u0 - halo), we = min(a.T, ws + TILE);    // frames it reads
    const P av = a.A[(long long)b * 4 + c * 3], cc = a.C[(long long)b * 4 + c * 3];
    const P q = a.s[b] * a.Q[(long long)b * 4 + c * 3];
    const P mean = a.ymean ? a.ymean[(long long)b * 2 + c] : P(0);
    const P* yp = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride + a.y.chan_off[c];
    const P* vp = reinterpret_cast<const P*>(a.var.base) + (long long)b * a.var.seq_stride + a.var.chan_off[c];
    P* xo = a.out + (long long)b * a.out_seq_stride + a.out_off[c];
    P* vo = a.out + (long long)b * a.out_seq_stride + a.out_off[2 + c];
    const bool vec_in = (((reinterpret_cast<uintptr_t>(yp) | reinterpret_cast<uintptr_t>(vp)) & 15) == 0);
    const bool vec_out = (((reinterpret_cast<uintptr_t>(xo) | reinterpret_cast<uintptr_t>(vo)) & 15) == 0);
    const P a2 = av * av, c2 = cc * cc, qc2 = q * c2;
    const int start = ws + threadIdx.x * L;
    const int nvalid = max(0, min(L, a.T - start));
    P y[L], r[L], Pf[L];
    load_chunk<P, L>(yp + start, vec_in, nvalid, mean, y);
    load_chunk<P, L>(vp + start, vec_in, nvalid, P(0), r);
    // non-finite observation or variance anywhere in the window -> exact kernels.  x * 0 is 0 for finite x and NaN for
    // +-inf / NaN, so ONE fused multiply-add per value accumulates the test (isfinite on the value converted to double
    // was 21 % of this kernel's instructions, ncu r2).
    P nonfinite = P(0);
#pragma unroll
    for (int i = 0; i < L; ++i) {
        if (i < nvalid) { nonfinite = fma(y[i], P(0), nonfinite); nonfinite = fma(r[i], P(0), nonfinite); }
        if (i >= nvalid) r[i] = P(1);
        else if (r[i] < P(1e-12)) r[i] = P(1e-12);  // np.clip(ev, 1e-12, None)
    }
    const bool bad = (nonfinite != P(0));           // NaN != 0
    // contraction bounds over this chunk, if it lies in one of the halos (chunk aligned: seg, halo are multiples of L)
    float lf = 0.f, lb = 0.f;
    const bool in_fh = (start < u0), in_bh = (start >= u1 && start < we);
    if (in_fh || in_bh) {
        float pf = 1.f, pb = 1.f;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            if (i < nvalid) {
                pf *= (float)(av * r[i] / (qc2 + r[i]));
                pb *= (float)(av * r[i] / (a2 * r[i] + qc2));
            }
            if ((i & 3) == 3 || i == L - 1) {   // keep the running products away from underflow
                if (in_fh) lf += __logf(fmaxf(pf, 1e-37f));
                if (in_bh) lb += __logf(fmaxf(pb, 1e-37f));
                pf = 1.f; pb = 1.f;
            }
        }
    }
    // ---- forward, phase 1a: chunk Moebius product, each factor pre-divided by r_i
    Mob<P> M{P(1), P(0), P(0), P(1)};
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const P ir = rcp_fast(r[i]);
        const Mob<P> Mi{fma(qc2, ir, a2), q, c2 * ir, P(1)};
        M = mob_mul(Mi, M);
        if (i & 1) mob_norm(M);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int d = 1 << k;
        const Mob<P> prev = mob_shfl_up(M, d);
        if (lane >= d) { M = mob_mul(M, prev); mob_norm(M); }
    }
    if (lane == 31) sh.magg[warp] = M;
    Mob<P> ex = mob_shfl_up(M, 1);
    if (lane == 0) ex = Mob<P>{P(1), P(0), P(0), P(1)};
    lf = warp_sum(lf);
    lb = warp_sum(lb);
    if (lane == 0) { sh.chk[warp][0] = lf; sh.chk[warp][1] = lb; }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    Mob<P> cm{P(1), P(0), P(0), P(1)};
    for (int w = 0; w < warp; ++w) { cm = mob_mul(sh.magg[w], cm); mob_norm(cm); }
    const P P_tile = a.S0[(long long)b * 4 + c * 3], m_tile = a.m0[(long long)b * 2 + c];
    const Mob<P> tot = mob_mul(ex, cm);
    P Pv = (tot.a * P_tile + tot.b) / (tot.c * P_tile + tot.d);
    if (threadIdx.x == 0) {   // verification of the two halos
        float cf = 0.f, cb = 0.f;
        for (int w = 0; w < DIAG_NW; ++w) { cf += sh.chk[w][0]; cb += sh.chk[w][1]; }
        const bool need_f = ws > 0, need_b = we < a.T;
        if (any_bad || (need_f && !(cf <= logtol)) || (need_b && !(cb <= logtol))) atomicExch(&fail[b], 1);
    }
    // ---- forward, phase 1b: exact per-frame filter arithmetic from the chunk's P; affine map of m
    P Aacc = P(1), bacc = P(0);
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const P S = fma(c2, Pv, r[i]);
        const P iSb = rcp_fast(S + P(1e-9));
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
    if (lane == 31) { sh.aagg[warp][0] = Aacc; sh.aagg[warp][1] = bacc; }
    P eA = __shfl_up_sync(0xffffffffu, Aacc, 1), eb = __shfl_up_sync(0xffffffffu, bacc, 1);
    if (lane == 0) { eA = P(1); eb = P(0); }
    __syncthreads();
    P mw = m_tile;
    for (int w = 0; w < warp; ++w) mw = fma(sh.aagg[w][0], mw, sh.aagg[w][1]);
    P m = fma(eA, mw, eb);
    // ---- forward, phase 3: filtered means (into y)
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const P e = fma(-cc, m, y[i]);
        const P mfi = fma(r[i], e, m);
        y[i] = mfi;
        m = av * mfi;
    }
    // ---- backward, phase 1: compose the chunk's affine maps, last frame first; r[] := G
    const int last = we - 1;      // at the last frame of the window: smoothed = filtered
    P Ag = P(1), bm = P(0), bP = P(0);
#pragma unroll
    for (int ii = 0; ii < L; ++ii) {
        const int i = L - 1 - ii;
        const P Sp = fma(a2, Pf[i], q);
        const P iSpb = rcp_fast(Sp + P(1e-9));
        P g = av * Pf[i] * iSpb;
        P om = y[i] * iSpb * (q + P(1e-9));
        P oP = Pf[i] * iSpb * (q + P(1e-9) * (P(1) + av * g));
        if (start + i >= last) { g = P(0); om = y[i]; oP = Pf[i]; }
        r[i] = g;
        y[i] = om;
        Pf[i] = oP;
        bm = fma(g, bm, om);
        bP = fma(g * g, bP, oP);
        Ag *= g;
    }
    // suffix scan: later chunks (higher lanes / warps) are applied first
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int d = 1 << k;
        const P nA = __shfl_down_sync(0xffffffffu, Ag, d);
        const P nm = __shfl_down_sync(0xffffffffu, bm, d);
        const P nP = __shfl_down_sync(0xffffffffu, bP, d);
        if (lane + d < 32) {
            bm = fma(Ag, nm, bm);
            bP = fma(Ag * Ag, nP, bP);
            Ag *= nA;
        }
    }
    if (lane == 0) { sh.bagg[warp][0] = Ag; sh.bagg[warp][1] = bm; sh.bagg[warp][2] = bP; }
    P xA = __shfl_down_sync(0xffffffffu, Ag, 1), xm = __shfl_down_sync(0xffffffffu, bm, 1),
      xP = __shfl_down_sync(0xffffffffu, bP, 1);
    if (lane == 31) { xA = P(1); xm = P(0); xP = P(0); }
    __syncthreads();
    P ms = P(0), Ps = P(0);     // state after the window: irrelevant (g = 0 at the window's last frame)
    for (int w = DIAG_NW - 1; w > warp; --w) {
        const P g = sh.bagg[w][0];
        ms = fma(g, ms, sh.bagg[w][1]);
        Ps = fma(g * g, Ps, sh.bagg[w][2]);
    }
    ms = fma(xA, ms, xm);
    Ps = fma(xA * xA, Ps, xP);
    // ---- backward, phase 3: smoothed moments, reprojected into the output planes
#pragma unroll
    for (int ii = 0; ii < L; ++ii) {
        const int i = L - 1 - ii;
        const P g = r[i];
        ms = fma(g, ms, y[i]);
        Ps = fma(g * g, Ps, Pf[i]);
        y[i] = a.latent_out ? ms : fma(cc, ms, mean);  // x = C m + mean   (singlecam_smoother.py:190-197)
        Pf[i] = a.latent_out ? Ps : c2 * Ps;           // diag(C V C^T)    (singlecam_smoother.py:191, 210-211)
    }
    if (start >= u0 && start < u1) {
        const int nst = min(L, u1 - start);
        store_chunk<P, L>(xo + start, vec_out, nst, y);
        store_chunk<P, L>(vo + start, vec_out, nst, Pf);
    }
}

size_t diag_smooth_workspace_bytes(int dtype, int B, int T) {
    return (size_t)B * 2 * T * 2 * (dtype == EKS_F32 ? 4 : 8) + 64 + (size_t)B * sizeof(int) + 256;
}

template <class P>
static int diag_smooth_launch(int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                              const void* C, const PlaneView& y, const PlaneView& var, const void* ymean,
                              const void* s, void* out, long long out_seq_stride, const long long* out_off,
                              int flags, void* workspace, cudaStream_t st) {
    DiagSmoothArgs<P> a;
    a.B = B; a.T = T;
    a.m0 = (const P*)m0; a.S0 = (const P*)S0; a.A = (const P*)A; a.Q = (const P*)Q; a.C = (const P*)C;
    a.y = y; a.var = var; a.ymean = (const P*)ymean; a.s = (const P*)s;
    // keep the workspace planes 16-byte aligned when T allows
    a.mf = (P*)workspace;
    a.Pf = a.mf + (size_t)B * 2 * T;
    a.out = (P*)out; a.out_seq_stride = out_seq_stride;
    for (int i = 0; i < 4; ++i) a.out_off[i] = out_off[i];
    a.latent_out = flags & 1;
    a.only_flagged = nullptr;
    int launches = 2;
    if (!(flags & 2)) {   // fused, time-segmented pass first; the exact kernels then redo only what it flagged
        constexpr int L = DiagTraits<P>::L;
        constexpr int TILE = DIAG_NT * L;
        static const int halo_env = [] { const char* e = getenv("EKS_SMOOTH_HALO"); return e ? atoi(e) : 256; }();
        int halo = (halo_env / L) * L;
        if (halo < L) halo = L;
        if (halo > TILE / 4) halo = TILE / 4;
        const int seg = (T <= TILE) ? TILE : TILE - 2 * halo;
        const int nseg = (T + seg - 1) / seg;
        int* fail = (int*)((unsigned char*)workspace + (((size_t)B * 2 * T * 2 * sizeof(P) + 64 + 255) & ~(size_t)255));
        cudaMemsetAsync(fail, 0, (size_t)B * sizeof(int), st);
        // log of the residual influence of a halo's start state: rounding level of the working precision relative to
        // coordinates of O(100) px after a start error of a few px
        const float logtol = sizeof(P) == 4 ? -18.f : -36.f;
        diag_smooth_fused_kernel<P><<<dim3(nseg, 2 * B), DIAG_NT, 0, st>>>(a, seg, halo, logtol, fail);
        const int rc = check_launch("diag_smooth_fused_kernel");
        if (rc) return rc;
        a.only_flagged = fail;
        launches = 3;
    }
    diag_filter_kernel<P><<<B * 2, DIAG_NT, 0, st>>>(a);
    int rc = check_launch("diag_filter_kernel");
    if (rc) return rc;
    diag_rts_kernel<P><<<B * 2, DIAG_NT, 0, st>>>(a);
    note_launches(launches);
    return check_launch("diag_rts_kernel");
}

}  // namespace eks

using namespace eks;

extern "C" size_t eks_diag_smooth_workspace_bytes(int dtype, int B, int T) { return diag_smooth_workspace_bytes(dtype, B, T); }

extern "C" int eks_diag_smooth(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                               const void* C, const void* y_base, long long y_seq_stride, const long long* y_off,
                               const void* ymean, const void* var_base, long long var_seq_stride,
                               const long long* var_off, const void* s, void* out, long long out_seq_stride,
                               const long long* out_off, int flags, void* workspace, size_t workspace_bytes,
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
                                         flags, workspace, st);
    return diag_smooth_launch<double>(B, T, m0, S0, A, Q, C, y, var, ymean, s, out, out_seq_stride, out_off,
                                      flags, workspace, st);
}
