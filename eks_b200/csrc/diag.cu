// diag.cu -- time-parallel kernels for DECOUPLED models: D == O == 2 with diagonal A, C, Q, S0 (the
// single-camera EKS model, eks/singlecam_smoother.py:246-284).  With a diagonal R the 2-D filter is
// two independent scalar filters that share the smoothing parameter s.
//
// (1) diag_optimize, "stream" mode: the Adam loop of _vmap_optimize_singletons / the block path
//     (eks/core.py:562-699, 403-559) as ONE streaming launch per evaluation (diag_nll_kernel; the evaluation itself is
//     diag_stream_cta in diag_stream.cuh) with the Adam step taken by the last CTA of each block.  The default
//     optimiser is the lag-statistics one of diag_lag.cu, which reads the observations once; this mode is kept for
//     EKS_OPT_MODE=stream and as its cross-check.
// (2) the final filter + RTS smoother pass with time-varying R_t (below).
#include <cstdlib>
#include "common.cuh"
#include "ekf_generic.cuh"
#include "diag.cuh"
#include "diag_stream.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int DIAG_NT = 256;
constexpr int DIAG_NW = DIAG_NT / 32;

template <class P> __device__ void diag_adam_body(const DiagOptArgs<P>& a, int j, bool first);

// ---- kernel A: one NLL(+d/ds) evaluation.  grid = (nseg, 2 * B): CTA (k, 2b+c) handles segment k of
// channel c of sequence b (diag_stream_cta).
template <class P>
__global__ void __launch_bounds__(OPT_NT, EKS_OPT_MINBLOCKS) diag_nll_kernel(const __grid_constant__ DiagOptArgs<P> a) {
    __shared__ ChanConst<P> shk;
    __shared__ double red[OPT_NW][2];
    extern __shared__ __align__(16) unsigned char ring[];
    const int seg = blockIdx.x, b = blockIdx.y >> 1, c = blockIdx.y & 1;
    const int blk = a.seq_block[b];
    if (blk < a.blk_lo || blk >= a.blk_hi || a.bstate[blk].done) return;
    const int warp = threadIdx.x >> 5;
    double* part = a.partials + (((long long)b * 2 + c) * a.nseg + seg) * 2;
    double te, tg;
    diag_stream_cta<P>(a, b, c, seg, a.nseg, ring, shk, red, te, tg);
    __shared__ int is_last;
    if (threadIdx.x == 0) {
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
