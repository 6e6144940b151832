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
#include <mutex>
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
    static std::mutex hs_mutex;   // the table is the library's only cross-call state: host threads may race to create it
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::lock_guard<std::mutex> hs_lock(hs_mutex);
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

}  // namespace eks
