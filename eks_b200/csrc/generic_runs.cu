// generic_runs.cu -- time-parallel execution of the GENERIC models (any linear model with D <= 6, obs <= 16,
// and the calibrated pinhole EKF) for long sequences.
//
// The sequential kernels of generic.cu give one thread a whole sequence: at 10^6 frames that is ~10^9 dependent
// cycles per pass.  Here a sequence is cut into runs of `run_len` frames, one thread per (sequence, run).  A run
// that does not start at frame 0 starts `W` frames early from the sequence's prior (m0, S0): a stable Kalman
// filter forgets its initial condition geometrically, so after the warm-up its state equals the sequential
// filter's to rounding.  This is not assumed but VERIFIED: every run records its state at its first frame
// (after warm-up) and at its end; adjacent records must agree to a tight relative tolerance.  By induction
// from run 0 (exact start) agreement at every boundary proves the whole trajectory -- NLL, d NLL/ds, filtered
// and smoothed moments -- equal to the sequential recursion up to that tolerance.  On disagreement the
// sequence's warm-up length is quadrupled and the evaluation repeated (no Adam step is taken on an
// unverified loss); with W >= n every run starts at frame 0, i.e. the scheme degrades to the exact one.
//
// Replaces the same reference code as generic.cu: eks/core.py:274-295 (final smoother), :403-559, :562-699
// (s-optimisation) for non-decoupled models.
#include "common.cuh"
#include "ekf_generic.cuh"
#include "generic.cuh"
#include "lin_lag.cuh"
#include "../../include/eks_b200.h"
#include <cstdio>
#include <vector>
#include <cstdlib>

namespace eks {

static int runs_w0() {            // initial warm-up length (frames); escalated x4 per failed boundary verification
    static int w0 = 0;
    if (w0 == 0) { const char* e = getenv("EKS_RUNS_W0"); w0 = e ? atoi(e) : 64; if (w0 < 4) w0 = 4; }
    return w0;
}
#define RUNS_W0 (runs_w0())
constexpr int RUNS_WMAX32 = 16384;  // fp32: beyond this warm-up a remaining boundary mismatch is rounding noise, not memory
constexpr int RUNS_EXTRA = 12;
constexpr int RUNS_RED_NT = 256;   // runs per CTA of the verification / reduction kernel    // extra evaluation slots for warm-up escalations (64 * 4^10 > 10^7 frames)

template <class P>
struct RunBlockState {
    AdamState<P> adam;
    P s, dsdlog;
    int done;
    int redo;        // diagnostic: number of repeated evaluations
    int unverified;  // diagnostic: evaluations accepted at the fp32 warm-up cap with a boundary mismatch left
};

template <class P>
struct RunArgs {
    int run_len, nruns, ns;   // ns = 2 (D + D^2): values per boundary record (m, P, dm, dP)
    int final_slot;           // evaluation slot index of this launch (for the forced finish)
    int total_slots;
    RunBlockState<P>* bstate; // [n_blocks]
    int* warm;                // [B] per-sequence warm-up length
    const int* seq_block;     // [B]
    double* part;             // [B][nruns][3]: nll, d nll/ds, bad
    P* bnd_start;             // [B][nruns][ns]
    P* bnd_end;               // [B][nruns][ns]
    int* flag;                // smoother: boundary mismatch flag
    P tol;                    // boundary agreement tolerance
    const P* lin_pinf;        // linear tabulated-gain path: [B][ns] covariance template; records hold (mu, dmu) only
    int probe;                // 1: evaluate once at the given a.s[b] and write nll / dnll (eks_nll_grad), no Adam
    double* part2;            // [B][nred][4]: per-256-run sums of part + boundary mismatch flag
    int nred;                 // ceil(nruns / RUNS_RED_NT)
    int* ndone;               // number of finished blocks (host reads it between chunks of evaluation slots)
};

template <class P> __host__ __device__ inline P runs_tol() { return sizeof(P) == 4 ? P(1e-4) : P(1e-10); }
// host: default tolerance, overridable for experiments with EKS_RUNS_TOL32 / EKS_RUNS_TOL64
template <class P> static P runs_tol_host() {
    const char* e = getenv(sizeof(P) == 4 ? "EKS_RUNS_TOL32" : "EKS_RUNS_TOL64");
    return e ? (P)atof(e) : runs_tol<P>();
}
// rounding floor: two different computation histories of the same quantity x differ by a few ulp of |x|
template <class P> __host__ __device__ inline P runs_ulp() { return sizeof(P) == 4 ? P(64 * 1.2e-7) : P(64 * 2.3e-16); }

// ---- one run of the filter NLL with s-sensitivities --------------------------------------------------------
template <class P, int DC, int OC, bool FIXED, bool NL>
__global__ void __launch_bounds__(32) gen_nll_runs_kernel(const __grid_constant__ GArgs<P> a,
                                                          const __grid_constant__ RunArgs<P> g) {
    using S = Dual<P>;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * g.nruns) return;
    const int b = idx / g.nruns, r = idx - b * g.nruns;
    const int blk = g.seq_block[b];
    if (blk < 0 || g.bstate[blk].done) return;
    const int n = a.sp.total;
    const int t0 = r * g.run_len, t1 = min(n, t0 + g.run_len);
    double* part = g.part + ((long long)b * g.nruns + r) * 3;
    if (t0 >= n) { part[0] = 0; part[1] = 0; part[2] = 0; return; }
    const int start = max(0, t0 - g.warm[b]);
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    const int D = dm.D(), O = dm.O();
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, false);
    FrameMap fm{a.sp};
    S m[DC], Pm[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) m[i] = S(mdl.m0[i]);
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = S(mdl.S0[i]);
    const S sd(g.bstate[blk].s, P(1));
    S nll = S(P(0));
    bool ok = true;
    bool a_id = true;       // A = I: the prediction step is skipped bit-exactly (63 of ~350 dual multiplications per frame)
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) a_id = a_id && (mdl.A[i * D + j] == (i == j ? P(1) : P(0)));
    P* bs = g.bnd_start + ((long long)b * g.nruns + r) * g.ns;
    P* be = g.bnd_end + ((long long)b * g.nruns + r) * g.ns;
    // (sector-wide grouped loads as in lin_runs_kernel were measured SLOWER here -- 0.85 -> 1.07 ms per evaluation for the
    // pinhole model: the extra registers cost occupancy and this kernel is bound by its ~3000 dependent
    // instructions per frame, not by L2 traffic)
    for (int i = start; i < t1; ++i) {
        if (i == t0) {  // state after warm-up = predicted state of the run's first frame
            for (int q = 0; q < D; ++q) { bs[q] = m[q].v; bs[D + D * D + q] = m[q].d; }
            for (int q = 0; q < D * D; ++q) { bs[D + q] = Pm[q].v; bs[2 * D + D * D + q] = Pm[q].d; }
            nll = S(P(0));
            ok = true;
        }
        P yv[OC], rv[OC];
        load_obs<P, OC>(ob, O, fm(i), yv, rv);
        ok = ekf_step<S, P, DC, OC, FIXED, NL>(dm, mdl, yv, rv, sd, m, Pm, nll, (S*)nullptr, (S*)nullptr, (const S*)nullptr,
                                               (const S*)nullptr, a_id) && ok;
    }
    for (int q = 0; q < D; ++q) { be[q] = m[q].v; be[D + D * D + q] = m[q].d; }
    for (int q = 0; q < D * D; ++q) { be[D + q] = Pm[q].v; be[2 * D + D * D + q] = Pm[q].d; }
    part[0] = (double)nll.v;
    part[1] = (double)nll.d;
    part[2] = ok ? 0.0 : 1.0;
}

// boundary agreement: |a - b| <= tol * scale, scale from the covariance (means), the variances (covariance)
// and the magnitude of the sensitivities themselves
template <class P>
__device__ inline bool runs_boundary_ok(const P* e, const P* s, int D, P sval, P tol) {
    bool ok = true;
    for (int i = 0; i < D; ++i) {
        const P sd = sqrt_(fabs(e[D + i * D + i]) + P(1e-30));
        ok = ok && (fabs(e[i] - s[i]) <= tol * sd + runs_ulp<P>() * fabs(e[i]));
        const P dsc = fabs(e[D + D * D + i]) + sd / sval;
        // the sensitivity inherits the rounding floor of the mean it is driven by (innovations y - C m cancel)
        ok = ok && (fabs(e[D + D * D + i] - s[D + D * D + i]) <= P(10) * tol * dsc + runs_ulp<P>() * fabs(e[i]) / sval);
        for (int j = 0; j < D; ++j) {
            const P pv = sqrt_(fabs(e[D + i * D + i] * e[D + j * D + j]) + P(1e-30));
            ok = ok && (fabs(e[D + i * D + j] - s[D + i * D + j]) <= tol * pv);
            const P dpv = fabs(e[2 * D + D * D + i * D + j]) + pv / sval;
            ok = ok && (fabs(e[2 * D + D * D + i * D + j] - s[2 * D + D * D + i * D + j]) <= P(10) * tol * dpv);
        }
    }
    return ok;
}

// compact records of the linear tabulated-gain path: (mu, dmu) only, the covariance is the same fixed point on both
// sides (pinf = [0, P_inf, 0, dP_inf] supplies the scales)
template <class P>
__device__ inline bool runs_boundary_ok_lin(const P* e, const P* s, const P* pinf, int D, P sval, P tol) {
    bool ok = true;
    for (int i = 0; i < D; ++i) {
        const P sd = sqrt_(fabs(pinf[D + i * D + i]) + P(1e-30));
        ok = ok && (fabs(e[i] - s[i]) <= tol * sd + runs_ulp<P>() * fabs(e[i]));
        const P dsc = fabs(e[D + i]) + sd / sval;
        ok = ok && (fabs(e[D + i] - s[D + i]) <= P(10) * tol * dsc + runs_ulp<P>() * fabs(e[i]) / sval);
    }
    return ok;
}

// ---- verification + reduction: one thread per run compares its start record with the previous run's end record;
// each CTA reduces 256 runs' partial sums in a fixed order.
template <class P>
__global__ void __launch_bounds__(RUNS_RED_NT) gen_runs_reduce_kernel(const __grid_constant__ GArgs<P> a,
                                                                      const __grid_constant__ RunArgs<P> g) {
    __shared__ double scratch[32];
    const int b = blockIdx.x, r = blockIdx.y * RUNS_RED_NT + threadIdx.x;   // problems on grid x: no 65535 limit
    const int blk = g.seq_block[b];
    if (blk < 0 || g.bstate[blk].done) return;
    const int n = a.sp.total;
    bool ok = true;
    double v = 0, dv = 0, bad = 0;
    if (r < g.nruns) {
        const int t0 = r * g.run_len;
        if (r >= 1 && t0 < n && t0 - g.warm[b] > 0) {   // runs that started at frame 0 are exact
            if (g.lin_pinf)
                ok = runs_boundary_ok_lin<P>(g.bnd_end + ((long long)b * g.nruns + r - 1) * 2 * a.D,
                                             g.bnd_start + ((long long)b * g.nruns + r) * 2 * a.D,
                                             g.lin_pinf + (long long)b * g.ns, a.D, g.bstate[blk].s, g.tol);
            else
                ok = runs_boundary_ok<P>(g.bnd_end + ((long long)b * g.nruns + r - 1) * g.ns,
                                         g.bnd_start + ((long long)b * g.nruns + r) * g.ns, a.D, g.bstate[blk].s, g.tol);
        }
        const double* p = g.part + ((long long)b * g.nruns + r) * 3;
        v = p[0]; dv = p[1]; bad = p[2];
    }
    const int all_ok = __syncthreads_and(ok);
    v = block_sum(v, scratch);
    dv = block_sum(dv, scratch);
    bad = block_sum(bad, scratch);
    if (threadIdx.x == 0) {
        double* o = g.part2 + ((long long)b * g.nred + blockIdx.y) * 4;
        o[0] = v; o[1] = dv; o[2] = bad; o[3] = all_ok ? 0.0 : 1.0;
    }
}

// ---- per block: Adam step on the verified loss (or escalate the warm-up and repeat the evaluation) ----------
constexpr int ADAM_RUNS_NT = 128;

template <class P>
__global__ void __launch_bounds__(ADAM_RUNS_NT) gen_adam_runs_kernel(const __grid_constant__ GArgs<P> a,
                                                                     const __grid_constant__ RunArgs<P> g, int first) {
    __shared__ double scratch[32];
    const int j = blockIdx.x, tid = threadIdx.x;
    RunBlockState<P>& bs = g.bstate[j];
    const int n = a.sp.total;
    if (first) {
        if (tid != 0) return;
        bs.redo = 0;
        bs.unverified = 0;
        if (g.probe) {   // singleton blocks: block j = sequence j, d s / d log s = 1 so that grad = d nll / d s
            bs.done = 0;
            bs.s = a.s[j];
            bs.dsdlog = P(1);
            return;
        }
        adam_init(bs.adam, a.s_log0[j]);
        bs.done = (a.cap <= 0);
        if (bs.done) {
            a.s_log_out[j] = bs.adam.s_log; a.last_loss_out[j] = bs.adam.prev; a.iters_out[j] = 0;
            if (g.ndone) atomicAdd(g.ndone, 1);
            return;
        }
        P dsdlog;
        bs.s = adam_current_s(bs.adam, a.lo, a.hi, &dsdlog);
        bs.dsdlog = dsdlog;
        return;
    }
    if (bs.done) return;      // uniform across the CTA (written by thread 0 of an earlier launch)
    bool verified = true;
    P loss = P(0), grad = P(0);
    for (int mi = a.block_off[j]; mi < a.block_off[j + 1]; ++mi) {
        const int b = a.members[mi];
        const int warm = g.warm[b];
        double v = 0, dv = 0, bad = 0, mism = 0;
        for (int c = tid; c < g.nred; c += ADAM_RUNS_NT) {
            const double* p = g.part2 + ((long long)b * g.nred + c) * 4;
            v += p[0]; dv += p[1]; bad += p[2]; mism += p[3];
        }
        v = block_sum(v, scratch);
        dv = block_sum(dv, scratch);
        bad = block_sum(bad, scratch);
        mism = block_sum(mism, scratch);
        if (mism > 0) {
            if (sizeof(P) == 4 && warm >= RUNS_WMAX32) {
                if (tid == 0) { bs.unverified += 1; if (g.ndone) atomicAdd(g.ndone + 1, 1); }
            } else {
                verified = false;
                if (tid == 0) g.warm[b] = (warm >= n / 4) ? n : warm * 4;
            }
        }
        P vv = (P)v, gg = (P)dv;
        if (bad > 0 || !isfinite(v) || !isfinite((double)vv)) { vv = P(1e12); gg = P(0); }  // core.py:650
        loss += vv;
        grad += gg * bs.dsdlog;
    }
    const bool last_slot = (g.final_slot == g.total_slots - 1);
    if (tid != 0) return;
    if (!verified && !last_slot) {
        bs.redo += 1;   // same s again with longer warm-ups; no Adam step on an unverified loss
        return;
    }
    if (g.probe) {
        a.nll_out[j] = loss;
        a.dnll_out[j] = grad;
        bs.done = 1;
        if (g.ndone) atomicAdd(g.ndone, 1);
        return;
    }
    if (a.trace && bs.adam.iters < a.trace_cap) {
        P* tr = a.trace + ((long long)j * a.trace_cap + bs.adam.iters) * 3;
        tr[0] = bs.adam.s_log; tr[1] = loss; tr[2] = grad * a.lr;
    }
    adam_step(bs.adam, loss, grad, a.lr, a.tol, a.cap);
    if (last_slot) bs.adam.done = true;
    if (bs.adam.done) {
        bs.done = 1;
        a.s_log_out[j] = bs.adam.s_log;
        a.last_loss_out[j] = bs.adam.prev;
        a.iters_out[j] = bs.adam.iters;
        if (g.ndone) atomicAdd(g.ndone, 1);
        return;
    }
    P dsdlog;
    bs.s = adam_current_s(bs.adam, a.lo, a.hi, &dsdlog);
    bs.dsdlog = dsdlog;
}

__global__ void gen_runs_init_kernel(int B, int n_blocks, const int* __restrict__ block_off,
                                     const int* __restrict__ members, int* __restrict__ seq_block,
                                     int* __restrict__ warm, int w0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) warm[i] = w0;
    if (i < n_blocks)
        for (int mi = block_off[i]; mi < block_off[i + 1]; ++mi) seq_block[members[mi]] = i;
}

static int runs_geometry(int n, int B, int& run_len) {
    // The per-run work is a latency-bound sequential recursion, so the runs are made as short as the 64-frame
    // warm-up makes sensible (run_len >= 64: at most 2x the frames) until ~96k threads are in flight; beyond
    // that the runs grow and the warm-up overhead shrinks.
    // (measured on B200, pinhole 6 x 5e5 frames: minimum run length 32 / 64 / 128 -> 1.20 / 0.85 / 1.01 ms per evaluation)
    static int min_len = 0;
    if (min_len == 0) { const char* e = getenv("EKS_RUNS_MINLEN"); min_len = e ? atoi(e) : 64; if (min_len < 32) min_len = 32; }
    int nruns = (98304 + B - 1) / B;
    run_len = (n + nruns - 1) / nruns;
    if (run_len < min_len) run_len = min_len;
    run_len = (run_len + 31) / 32 * 32;
    return (n + run_len - 1) / run_len;
}

size_t generic_runs_optimize_workspace_bytes(int dtype, int n_blocks, int B, int D, int T) {
    const size_t w = dtype == EKS_F32 ? 4 : 8;
    int run_len;
    const int nruns = runs_geometry(T, B, run_len);
    const size_t ns = 2 * (size_t)(D + D * D);
    size_t bytes = 1024 + 256;
    bytes += (size_t)n_blocks * 128;
    bytes += 2 * ((size_t)B * sizeof(int) + 256);
    bytes += (size_t)B * nruns * 3 * sizeof(double) + 256;
    bytes += 2 * ((size_t)B * nruns * ns * w + 256);
    bytes += (size_t)B * ((nruns + RUNS_RED_NT - 1) / RUNS_RED_NT) * 4 * sizeof(double) + 256;
    return bytes;
}
size_t linear_steady_workspace_bytes(int dtype, int B, int D, int O, int T);

// The evaluation slots are enqueued in chunks; between chunks the host reads the number of finished blocks and stops
// as soon as all are done (the reference's cap is 300 evaluations, typical counts are 50-140: enqueueing every slot
// cost ~700 no-op launches per call).  This entry point therefore synchronises the stream once per chunk.
constexpr int RUNS_SLOT_CHUNK = 32;
static bool runs_all_done(const int* ndone, int n_blocks, cudaStream_t st) {
    int h[2] = {0, 0};      // [0] finished blocks, [1] evaluations accepted with an unverified boundary (float32 cap)
    if (cudaMemcpyAsync(h, ndone, 2 * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    note_unverified(h[1]);
    return h[0] >= n_blocks;
}

template <class P, int DC, int OC, bool FIXED, bool NL>
static int runs_optimize_launch(const GArgs<P>& a, RunArgs<P> g, cudaStream_t st) {
    const int nthreads = a.B * g.nruns;
    // extra slots for warm-up escalations (a repeated evaluation consumes a slot): every member of a shared-s block can
    // escalate at a different iteration, so the allowance grows with the largest possible block (B - n_blocks + 1
    // members); unused slots cost nothing, the loop stops when every block is done
    const int max_members = a.B - a.n_blocks + 1 < 1 ? 1 : (a.B - a.n_blocks + 1 > 8 ? 8 : a.B - a.n_blocks + 1);
    const int slots = a.cap + RUNS_EXTRA * max_members;
    g.total_slots = slots;
    g.final_slot = -1;
    // (splitting the blocks over internal streams as diag_optimize_run does was measured and does not pay here:
    // linear 13.0 -> 13.6 ms, pinhole 102.6 -> 104.1 ms with two streams; the chain is latency bound per sequence)
    gen_adam_runs_kernel<P><<<a.n_blocks, ADAM_RUNS_NT, 0, st>>>(a, g, 1);
    int launched = 1;
    for (int it = 0; it < slots; ++it) {
        g.final_slot = it;
        gen_nll_runs_kernel<P, DC, OC, FIXED, NL><<<(nthreads + 31) / 32, 32, 0, st>>>(a, g);
        gen_runs_reduce_kernel<P><<<dim3(a.B, g.nred), RUNS_RED_NT, 0, st>>>(a, g);
        gen_adam_runs_kernel<P><<<a.n_blocks, ADAM_RUNS_NT, 0, st>>>(a, g, 0);
        launched += 3;
        if ((it + 1) % RUNS_SLOT_CHUNK == 0 && it + 1 < slots && runs_all_done(g.ndone, a.n_blocks, st)) break;
    }
    if (getenv("EKS_DEBUG_RUNS")) {
        cudaStreamSynchronize(st);
        std::vector<int> w(a.B);
        cudaMemcpy(w.data(), g.warm, a.B * sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[eks runs] optimiser: run_len %d nruns %d tol %g warm", g.run_len, g.nruns, (double)g.tol);
        for (int b = 0; b < a.B && b < 16; ++b) fprintf(stderr, " %d", w[b]);
        fprintf(stderr, "\n");
    }
    note_launches(launched);
    runs_all_done(g.ndone, a.n_blocks, st);     // final counters (this entry point synchronises the stream)
    return check_launch("generic run-parallel optimise kernels");
}


// =====================================================================================================
// final pass: forward filter runs, then RTS runs (right-to-left, warm-up on the right), both verified
// =====================================================================================================
template <class P>
struct SmoothRunArgs {
    int run_len, nruns, warm, nm;  // nm = D + D^2 values per boundary record
    P* bnd_f_start;  // [B][nruns][nm] filter state after warm-up at the run's first frame
    P* bnd_f_end;    // [B][nruns][nm] filter state after the run's last frame
    P* bnd_b;        // [B][nruns][nm] this run's estimate of the smoothed moments at frame t1 (next run's first)
    int* flag;       // [2]: forward / backward mismatch
    P tol;
};

template <class P, int DC, int OC, bool FIXED, bool NL>
__global__ void __launch_bounds__(32) gen_filter_runs_kernel(const __grid_constant__ GArgs<P> a,
                                                             const __grid_constant__ SmoothRunArgs<P> g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * g.nruns) return;
    const int b = idx / g.nruns, r = idx - b * g.nruns;
    const int T = a.T;
    const int t0 = r * g.run_len, t1 = min(T, t0 + g.run_len);
    if (t0 >= T) return;
    const int start = max(0, t0 - g.warm);
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    const int D = dm.D(), O = dm.O();
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, true);
    P m[DC], Pm[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) m[i] = mdl.m0[i];
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = mdl.S0[i];
    const P s = a.s[b];
    P nll = P(0);
    P* bs = g.bnd_f_start + ((long long)b * g.nruns + r) * g.nm;
    P* be = g.bnd_f_end + ((long long)b * g.nruns + r) * g.nm;
    P* mfb = a.mf + (long long)b * T * D;
    P* Pfb = a.Pf + (long long)b * T * D * D;
    for (int t = start; t < t1; ++t) {
        if (t == t0) {
            for (int q = 0; q < D; ++q) bs[q] = m[q];
            for (int q = 0; q < D * D; ++q) bs[D + q] = Pm[q];
        }
        P yv[OC], rv[OC];
        load_obs<P, OC>(ob, O, t, yv, rv);
        P mf[DC], Pf[DC * DC];
        ekf_step<P, P, DC, OC, FIXED, NL, true>(dm, mdl, yv, rv, s, m, Pm, nll, mf, Pf);
        if (t >= t0) {
            for (int q = 0; q < D; ++q) mfb[(long long)t * D + q] = mf[q];
            for (int q = 0; q < D * D; ++q) Pfb[(long long)t * D * D + q] = Pf[q];
        }
    }
    for (int q = 0; q < D; ++q) be[q] = m[q];
    for (int q = 0; q < D * D; ++q) be[D + q] = Pm[q];
}

template <class P, int DC, int OC, bool FIXED>
__global__ void __launch_bounds__(32) gen_rts_runs_kernel(const __grid_constant__ GArgs<P> a,
                                                          const __grid_constant__ SmoothRunArgs<P> g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * g.nruns) return;
    const int b = idx / g.nruns, r = idx - b * g.nruns;
    const int T = a.T;
    const int t0 = r * g.run_len, t1 = min(T, t0 + g.run_len);
    if (t0 >= T) return;
    const int stop = min(T - 1, t1 - 1 + g.warm);  // first frame of the backward recursion (exact if T-1)
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    const int D = dm.D();
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, true);
    const P s = a.s[b];
    const P* mfb = a.mf + (long long)b * T * D;
    const P* Pfb = a.Pf + (long long)b * T * D * D;
    P* msb = a.ms + (long long)b * T * D;
    P* Vsb = a.Vs + (long long)b * T * D * D;
    P* bb = g.bnd_b + ((long long)b * g.nruns + r) * g.nm;
    P msn[DC], Vsn[DC * DC];
    for (int q = 0; q < D; ++q) msn[q] = mfb[(long long)stop * D + q];
    for (int q = 0; q < D * D; ++q) Vsn[q] = Pfb[(long long)stop * D * D + q];
    for (int t = stop; t >= t0; --t) {
        if (t < stop) {
            P mft[DC], Pft[DC * DC];
            for (int q = 0; q < D; ++q) mft[q] = mfb[(long long)t * D + q];
            for (int q = 0; q < D * D; ++q) Pft[q] = Pfb[(long long)t * D * D + q];
            rts_step<P, DC, OC, FIXED>(dm, mdl, s, mft, Pft, msn, Vsn);
        }
        if (t == t1) {  // own estimate of the next run's first frame
            for (int q = 0; q < D; ++q) bb[q] = msn[q];
            for (int q = 0; q < D * D; ++q) bb[D + q] = Vsn[q];
        }
        if (t < t1) {
            for (int q = 0; q < D; ++q) msb[(long long)t * D + q] = msn[q];
            for (int q = 0; q < D * D; ++q) Vsb[(long long)t * D * D + q] = Vsn[q];
        }
    }
}

template <class P>
__device__ inline bool moments_agree(const P* x, const P* y, int D, P tol) {
    bool ok = true;
    for (int i = 0; i < D; ++i) {
        const P sd = sqrt_(fabs(x[D + i * D + i]) + P(1e-30));
        ok = ok && (fabs(x[i] - y[i]) <= tol * sd + runs_ulp<P>() * fabs(x[i]));
        for (int j = 0; j < D; ++j) {
            const P pv = sqrt_(fabs(x[D + i * D + i] * x[D + j * D + j]) + P(1e-30));
            ok = ok && (fabs(x[D + i * D + j] - y[D + i * D + j]) <= tol * pv);
        }
    }
    return ok;
}

// which = 0: filter boundaries (bnd_f_end[r-1] vs bnd_f_start[r]); which = 1: smoother boundaries
// (bnd_b[r] vs the smoothed moments written by run r+1 at its first frame)
template <class P>
__global__ void gen_smooth_check_kernel(const __grid_constant__ GArgs<P> a, const __grid_constant__ SmoothRunArgs<P> g,
                                        int which) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * g.nruns) return;
    const int b = idx / g.nruns, r = idx - b * g.nruns;
    const int T = a.T, D = a.D;
    const int t0 = r * g.run_len, t1 = min(T, t0 + g.run_len);
    if (t0 >= T) return;
    bool ok = true;
    if (which == 0) {
        if (r > 0 && t0 - g.warm > 0)
            ok = moments_agree<P>(g.bnd_f_end + ((long long)b * g.nruns + r - 1) * g.nm,
                                  g.bnd_f_start + ((long long)b * g.nruns + r) * g.nm, D, g.tol);
    } else {
        if (t1 < T && t1 - 1 + g.warm < T - 1) {
            P y[EKS_MAX_STATE + EKS_MAX_STATE * EKS_MAX_STATE];
            for (int q = 0; q < D; ++q) y[q] = a.ms[((long long)b * T + t1) * D + q];
            for (int q = 0; q < D * D; ++q) y[D + q] = a.Vs[((long long)b * T + t1) * D * D + q];
            ok = moments_agree<P>(y, g.bnd_b + ((long long)b * g.nruns + r) * g.nm, D, g.tol);
        }
    }
    if (!ok) atomicExch(g.flag + which, 1);
}

template <class P, int DC, int OC, bool FIXED, bool NL>
static int runs_smooth_launch(const GArgs<P>& a, SmoothRunArgs<P> g, cudaStream_t st) {
    const int nthreads = a.B * g.nruns;
    const int blocks = (nthreads + 31) / 32;
    int h_flag[2];
    // NOTE: unlike the other entry points this driver synchronises the stream: the boundary verification
    // decides on the host whether the pass has to be repeated with a longer warm-up.
    for (int attempt = 0; attempt < 16; ++attempt) {
        cudaMemsetAsync(g.flag, 0, 2 * sizeof(int), st);
        gen_filter_runs_kernel<P, DC, OC, FIXED, NL><<<blocks, 32, 0, st>>>(a, g);
        gen_smooth_check_kernel<P><<<(nthreads + 127) / 128, 128, 0, st>>>(a, g, 0);
        gen_rts_runs_kernel<P, DC, OC, FIXED><<<blocks, 32, 0, st>>>(a, g);
        gen_smooth_check_kernel<P><<<(nthreads + 127) / 128, 128, 0, st>>>(a, g, 1);
        cudaError_t e = cudaMemcpyAsync(h_flag, g.flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("generic run-parallel smoother: %s", cudaGetErrorString(e));
            return (int)e;
        }
        if (getenv("EKS_DEBUG_RUNS"))
            fprintf(stderr, "[eks runs] smoother attempt %d: warm %d run_len %d nruns %d flags %d %d\n", attempt, g.warm,
                    g.run_len, g.nruns, h_flag[0], h_flag[1]);
        if (!h_flag[0] && !h_flag[1]) return 0;
        if (g.warm >= a.T || (sizeof(P) == 4 && g.warm >= RUNS_WMAX32)) break;  // exact / fp32 rounding floor  // already exact: nothing more to escalate
        g.warm = (g.warm >= a.T / 4) ? a.T : g.warm * 4;
    }
    return check_launch("generic run-parallel smoother kernels");
}

template <class P>
int generic_runs_smooth(const GArgs<P>& a, cudaStream_t st) {
    SmoothRunArgs<P> g;
    g.nruns = runs_geometry(a.T, a.B, g.run_len);
    g.warm = RUNS_W0;
    g.tol = runs_tol_host<P>();
    g.nm = a.D + a.D * a.D;
    // boundary records live behind the filtered moments in the caller's workspace (sized by
    // eks_filter_smooth_workspace_bytes)
    P* w = a.Pf + (size_t)a.B * a.T * a.D * a.D;
    const size_t rec = (size_t)a.B * g.nruns * g.nm;
    g.bnd_f_start = w;
    g.bnd_f_end = w + rec;
    g.bnd_b = w + 2 * rec;
    g.flag = (int*)(w + 3 * rec);
    const int D = a.D, O = a.O;
    if (a.ncam > 0) {
        if (O == 4) return runs_smooth_launch<P, 3, 4, true, true>(a, g, st);
        if (O == 6) return runs_smooth_launch<P, 3, 6, true, true>(a, g, st);
        if (O == 8) return runs_smooth_launch<P, 3, 8, true, true>(a, g, st);
        return runs_smooth_launch<P, 3, EKS_MAX_CHAN, false, true>(a, g, st);
    }
    if (D == 2 && O == 2) return runs_smooth_launch<P, 2, 2, true, false>(a, g, st);
    if (D == 3 && O == 4) return runs_smooth_launch<P, 3, 4, true, false>(a, g, st);
    if (D == 3 && O == 6) return runs_smooth_launch<P, 3, 6, true, false>(a, g, st);
    if (D == 3 && O == 8) return runs_smooth_launch<P, 3, 8, true, false>(a, g, st);
    return runs_smooth_launch<P, EKS_MAX_STATE, EKS_MAX_CHAN, false, false>(a, g, st);
}
template int generic_runs_smooth<float>(const GArgs<float>&, cudaStream_t);
template int generic_runs_smooth<double>(const GArgs<double>&, cudaStream_t);

// =====================================================================================================
// Linear models with a CONSTANT observation noise (the loss path of every linear model, eks/core.py:702-709):
// the covariance recursion is data independent, so ONE thread per sequence runs it (with its s-sensitivity) from
// S0 until it reaches its fixed point, tabulating the per-channel gains k_g, 1/s_g (and d/ds) of the sequential
// scalar updates for the transient frames.  The run threads then only carry the mean and its sensitivity:
//     e_g = y_g - h_g . mu,  mu += k_g e_g  (g = 0..O-1),  mu <- A mu          (same order as ekf_step)
// with the tabulated gains (t < n_tr) or the steady ones -- ~100 FMAs per frame instead of a full EKF step.
// The data independent part of the NLL (sum of log s_g) is added analytically.  Warm-up, boundary verification,
// escalation and the Adam step are those of the generic run-parallel path above.
// =====================================================================================================
template <class P>
struct LinArgs {
    int ent, ncap;          // values per table entry = O (2 D + 2); table capacity in frames
    P* table;               // [B][ncap + 1][ent]: entries of frames 0..ncap-1, then the steady entry
    P* pinf;                // [B][2 (D + D^2)]: boundary-record template (0, P_inf, 0, dP_inf)
    int* ntr;               // [B] transient length
    double* cnll;           // [B][3]: data independent NLL, its derivative, bad flag
};

// one sequential-scalar-update covariance step on the predicted covariance Pm (dual), recording the gains
template <class P, int DC, int OC, bool FIXED>
__device__ inline bool lin_cov_step(const Dims<DC, OC, FIXED>& dm, const SeqModel<P>& mdl, const P* rconst, Dual<P> s,
                                    Dual<P>* Pm, P* ent, Dual<P>& lsum, bool a_identity) {
    using S = Dual<P>;
    const int D = dm.D(), O = dm.O();
    bool ok = true;
    lsum = S(P(0));
#pragma unroll
    for (int g = 0; g < OC; ++g) {
        if (g >= O) break;
        S h[DC], Ph[DC];
#pragma unroll
        for (int j = 0; j < DC; ++j) if (j < D) h[j] = S(mdl.C[g * D + j]);
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            if (i < D) {
                S acc = S(P(0));
#pragma unroll
                for (int j = 0; j < DC; ++j) if (j < D) acc += Pm[i * D + j] * h[j];
                Ph[i] = acc;
            }
        }
        S si = S(rconst[g]);
#pragma unroll
        for (int j = 0; j < DC; ++j) if (j < D) si += h[j] * Ph[j];
        if (!(si.v > 0) || !isfinite((double)si.v)) ok = false;
        const S isi = S(P(1)) / si;
        lsum += log_(si);
        P* e = ent + g * (2 * D + 2);
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            if (i < D) {
                const S k = Ph[i] * isi;
                e[i] = k.v; e[D + i] = k.d;
#pragma unroll
                for (int j = 0; j < DC; ++j) if (j < D) Pm[i * D + j] -= k * Ph[j];
            }
        }
        e[2 * D] = isi.v; e[2 * D + 1] = isi.d;
    }
#pragma unroll
    for (int i = 0; i < DC; ++i)
#pragma unroll
        for (int j = 0; j < DC; ++j)
            if (i < D && j < D && j > i) {
                const S a = S(P(0.5)) * (Pm[i * D + j] + Pm[j * D + i]);
                Pm[i * D + j] = a; Pm[j * D + i] = a;
            }
    if (a_identity) {   // A = I (every multi-camera model of the reference): A P A^T is P itself, bit for bit
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = Pm[i] + s * S(mdl.Q[i]);
        return ok;
    }
    S AP[DC * DC];
#pragma unroll
    for (int i = 0; i < DC; ++i)
#pragma unroll
        for (int j = 0; j < DC; ++j)
            if (i < D && j < D) {
                S a2 = S(P(0));
#pragma unroll
                for (int k = 0; k < DC; ++k) if (k < D) a2 += S(mdl.A[i * D + k]) * Pm[k * D + j];
                AP[i * D + j] = a2;
            }
#pragma unroll
    for (int i = 0; i < DC; ++i)
#pragma unroll
        for (int j = 0; j < DC; ++j)
            if (i < D && j < D) {
                S acc = S(P(0));
#pragma unroll
                for (int k = 0; k < DC; ++k) if (k < D) acc += AP[i * D + k] * S(mdl.A[j * D + k]);
                Pm[i * D + j] = acc + s * S(mdl.Q[i * D + j]);
            }
    return ok;
}

template <class P>
__device__ inline bool is_identity(const P* A, int D) {
    bool id = true;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) id = id && (A[i * D + j] == (i == j ? P(1) : P(0)));
    return id;
}

template <class P, int DC, int OC, bool FIXED>
__global__ void __launch_bounds__(32) lin_prep_kernel(const __grid_constant__ GArgs<P> a,
                                                      const __grid_constant__ RunArgs<P> g,
                                                      const __grid_constant__ LinArgs<P> l) {
    using S = Dual<P>;
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int blk = g.seq_block[b];
    if (blk < 0 || g.bstate[blk].done) return;
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    const int D = dm.D(), O = dm.O(), n = a.sp.total;
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, false);
    const S s(g.bstate[blk].s, P(1));
    const bool a_id = is_identity<P>(mdl.A, D);
    S Pm[DC * DC];
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Pm[i] = S(mdl.S0[i]);
    const P HALF_LOG2PI = P(0.91893853320467274178032973640562);
    const P tol = P(8) * (sizeof(P) == 4 ? P(1.1920929e-7) : P(2.220446049250313e-16));
    P* tab = l.table + (long long)b * (l.ncap + 1) * l.ent;
    double c0 = 0, c1 = 0;
    bool ok = true;
    P prev_cv = P(INFINITY), prev_cd = P(INFINITY);
    int stall = 0, t = 0;
    const int cap = n < l.ncap ? n : l.ncap;
    bool conv = false;
    for (; t < cap && !conv; ++t) {
        P old_v[DC * DC], old_d[DC * DC];
#pragma unroll
        for (int i = 0; i < DC * DC; ++i) if (i < D * D) { old_v[i] = Pm[i].v; old_d[i] = Pm[i].d; }
        S lsum;
        ok = lin_cov_step<P, DC, OC, FIXED>(dm, mdl, ob.Rconst, s, Pm, tab + (long long)t * l.ent, lsum, a_id) && ok;
        c0 += (double)(P(O) * HALF_LOG2PI) + 0.5 * (double)lsum.v;
        c1 += 0.5 * (double)lsum.d;
        // distance to the fixed point from the last two steps (geometric convergence), or the rounding floor
        P cv = 0, cd = 0, sv = 0, sdv = 0;
#pragma unroll
        for (int i = 0; i < DC * DC; ++i)
            if (i < D * D) {
                cv = fmax(cv, fabs(Pm[i].v - old_v[i])); cd = fmax(cd, fabs(Pm[i].d - old_d[i]));
                sv = fmax(sv, fabs(Pm[i].v)); sdv = fmax(sdv, fabs(Pm[i].d));
            }
        const P rv = fmin(cv / prev_cv, P(0.999)), rd = fmin(cd / prev_cd, P(0.999));
        const bool cvok = (cv == P(0)) || (isfinite((double)prev_cv) && cv * rv / (P(1) - rv) <= tol * sv);
        const bool cdok = (cd == P(0)) || (isfinite((double)prev_cd) && cd * rd / (P(1) - rd) <= tol * sdv);
        if (cv >= prev_cv && cd >= prev_cd) ++stall;
        conv = (cvok && cdok) || stall >= 24;
        prev_cv = cv; prev_cd = cd;
    }
    const int ntr = t;
    // steady entry from the converged predicted covariance; boundary-record template
    P* pinf = l.pinf + (long long)b * g.ns;
    for (int q = 0; q < D; ++q) { pinf[q] = P(0); pinf[D + D * D + q] = P(0); }
    for (int q = 0; q < D * D; ++q) { pinf[D + q] = Pm[q].v; pinf[2 * D + D * D + q] = Pm[q].d; }
    S lsum;
    ok = lin_cov_step<P, DC, OC, FIXED>(dm, mdl, ob.Rconst, s, Pm, tab + (long long)l.ncap * l.ent, lsum, a_id) && ok;
    const double rest = (double)(n - ntr);
    c0 += rest * ((double)(P(O) * HALF_LOG2PI) + 0.5 * (double)lsum.v);
    c1 += rest * 0.5 * (double)lsum.d;
    l.ntr[b] = ntr;
    l.cnll[3 * b] = c0; l.cnll[3 * b + 1] = c1; l.cnll[3 * b + 2] = ok ? 0.0 : 1.0;
}

template <class P, int DC, int OC, bool FIXED>
__global__ void __launch_bounds__(64) lin_runs_kernel(const __grid_constant__ GArgs<P> a,
                                                      const __grid_constant__ RunArgs<P> g,
                                                      const __grid_constant__ LinArgs<P> l) {
    constexpr int ENTC = OC * (2 * DC + 2);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * g.nruns) return;
    const int b = idx / g.nruns, r = idx - b * g.nruns;
    const int blk = g.seq_block[b];
    if (blk < 0 || g.bstate[blk].done) return;
    const int n = a.sp.total;
    const int t0 = r * g.run_len, t1 = min(n, t0 + g.run_len);
    double* part = g.part + ((long long)b * g.nruns + r) * 3;
    if (t0 >= n) { part[0] = 0; part[1] = 0; part[2] = 0; return; }
    const int start = max(0, t0 - g.warm[b]);
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    const int D = dm.D(), O = dm.O(), ES = 2 * D + 2;
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, false);
    FrameMap fm{a.sp};
    const int ntr = l.ntr[b];
    const P* tab = l.table + (long long)b * (l.ncap + 1) * l.ent;
    P Cm[OC * DC], Am[DC * DC];
#pragma unroll
    for (int i = 0; i < OC * DC; ++i) if (i < O * D) Cm[i] = mdl.C[i];
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) if (i < D * D) Am[i] = mdl.A[i];
    const bool a_id = is_identity<P>(mdl.A, D);
    P ent[ENTC];
    bool steady_loaded = false;
    P mu[DC], dmu[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) if (i < D) { mu[i] = start == 0 ? mdl.m0[i] : P(0); dmu[i] = P(0); }
    P q = P(0), dq = P(0);
    P* bs = g.bnd_start + ((long long)b * g.nruns + r) * 2 * D;   // compact records: (mu, dmu)
    P* be = g.bnd_end + ((long long)b * g.nruns + r) * 2 * D;
    // one frame of the mean recursion with the observations yv (not yet centred)
    auto frame = [&](int i, const P* yv) {
        if (i < ntr) {
            const P* e = tab + (long long)i * l.ent;
#pragma unroll
            for (int k = 0; k < ENTC; ++k) if (k < O * ES) ent[k] = e[k];
        } else if (!steady_loaded) {
            const P* e = tab + (long long)l.ncap * l.ent;
#pragma unroll
            for (int k = 0; k < ENTC; ++k) if (k < O * ES) ent[k] = e[k];
            steady_loaded = true;
        }
#pragma unroll
        for (int c = 0; c < OC; ++c) {
            if (c < O) {
                P y = yv[c];
                if (ob.ymean) y -= ob.ymean[c];
                P e = y, de = P(0);
#pragma unroll
                for (int j = 0; j < DC; ++j) if (j < D) { e -= Cm[c * D + j] * mu[j]; de -= Cm[c * D + j] * dmu[j]; }
                const P* en = ent + c * ES;
                const P isi = en[2 * D], disi = en[2 * D + 1];
                q += e * e * isi;
                dq += P(2) * e * de * isi + e * e * disi;
#pragma unroll
                for (int j = 0; j < DC; ++j) if (j < D) { mu[j] += en[j] * e; dmu[j] += en[D + j] * e + en[j] * de; }
            }
        }
        if (a_id) return;        // A = I: mu, dmu unchanged by the prediction
        P nm[DC], ndm[DC];
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            if (j < D) {
                P s1 = P(0), s2 = P(0);
#pragma unroll
                for (int k = 0; k < DC; ++k) if (k < D) { s1 += Am[j * D + k] * mu[k]; s2 += Am[j * D + k] * dmu[k]; }
                nm[j] = s1; ndm[j] = s2;
            }
        }
#pragma unroll
        for (int j = 0; j < DC; ++j) if (j < D) { mu[j] = nm[j]; dmu[j] = ndm[j]; }
    };
    // Each lane walks its own run, so a per-frame scalar load touches one 32-byte sector per lane and uses 4 bytes
    // of it: 8x the useful L2 traffic (measured: 170 us per evaluation at 4 x 1e6 frames, L2-bound).  Fixed-size
    // models therefore fetch G = 8 consecutive frames per channel with 16-byte loads (whole sectors, once).
    constexpr int G = FIXED ? 32 / (int)sizeof(P) : 1;   // one 32-byte sector per channel
    constexpr int VE = 16 / (int)sizeof(P);            // elements per 16-byte load
    const bool contiguous = (a.sp.n == 1);
    int i = start;
    while (i < t1) {
        if (i == t0) {
            for (int k = 0; k < D; ++k) { bs[k] = mu[k]; bs[D + k] = dmu[k]; }
            q = P(0); dq = P(0);
        }
        const long long f = fm(i);
        bool grouped = false;
        if (G > 1 && contiguous && i + G <= t1 && (i >= t0 || i + G <= t0)) {
            bool aligned = true;
#pragma unroll
            for (int c = 0; c < OC; ++c)
                if (c < O) aligned = aligned && ((reinterpret_cast<uintptr_t>(ob.y_base + ob.y_off[c] + f) & 15) == 0);
            if (aligned) {
                P yb[OC][G];
#pragma unroll
                for (int c = 0; c < OC; ++c) {
                    if (c < O) {
                        const P* p = ob.y_base + ob.y_off[c] + f;
#pragma unroll
                        for (int v4 = 0; v4 < G / VE; ++v4) {
                            if (sizeof(P) == 4) {
                                const float4 w4 = *reinterpret_cast<const float4*>(p + v4 * VE);
                                yb[c][v4 * VE + 0] = (P)w4.x; yb[c][v4 * VE + 1] = (P)w4.y;
                                yb[c][v4 * VE + 2] = (P)w4.z; yb[c][v4 * VE + 3] = (P)w4.w;
                            } else {
                                const double2 w2 = *reinterpret_cast<const double2*>(p + v4 * VE);
                                yb[c][v4 * VE + 0] = (P)w2.x; yb[c][v4 * VE + 1] = (P)w2.y;
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    P yv[OC];
#pragma unroll
                    for (int c = 0; c < OC; ++c) if (c < O) yv[c] = yb[c][j];
                    frame(i + j, yv);
                }
                i += G;
                grouped = true;
            }
        }
        if (!grouped) {
            P yv[OC];
#pragma unroll
            for (int c = 0; c < OC; ++c) if (c < O) yv[c] = ob.y_base[ob.y_off[c] + f];
            frame(i, yv);
            i += 1;
        }
    }
    for (int k = 0; k < D; ++k) { be[k] = mu[k]; be[D + k] = dmu[k]; }
    double v = 0.5 * (double)q, dv = 0.5 * (double)dq, bad = 0;
    if (r == 0) { v += l.cnll[3 * b]; dv += l.cnll[3 * b + 1]; bad = l.cnll[3 * b + 2]; }
    part[0] = v; part[1] = dv; part[2] = bad;
}

static int lin_table_cap(int dtype, int B, int D, int O, int n) {
    // transient table: at most ~256 MB, between 1024 and 16384 frames per sequence
    const size_t ent_bytes = (size_t)O * (2 * D + 2) * (dtype == EKS_F32 ? 4 : 8);
    long long cap = (long long)((256ull << 20) / ((size_t)B * ent_bytes));
    if (cap > 16384) cap = 16384;
    if (cap < 1024) cap = 1024;
    if (cap > n) cap = n;
    return (int)cap;
}

size_t linear_steady_workspace_bytes(int dtype, int B, int D, int O, int T) {
    const size_t w = dtype == EKS_F32 ? 4 : 8;
    const int cap = lin_table_cap(dtype, B, D, O, T);
    return (size_t)B * (cap + 1) * O * (2 * D + 2) * w + 256 + (size_t)B * 2 * (D + D * D) * w + 256 +
           (size_t)B * 4 + 256 + (size_t)B * 3 * 8 + 256;
}

template <class P, int DC, int OC, bool FIXED>
static int lin_optimize_launch(const GArgs<P>& a, RunArgs<P> g, const LinArgs<P>& l, cudaStream_t st) {
    const int nthreads = a.B * g.nruns;
    // extra slots for warm-up escalations (a repeated evaluation consumes a slot): every member of a shared-s block can
    // escalate at a different iteration, so the allowance grows with the largest possible block (B - n_blocks + 1
    // members); unused slots cost nothing, the loop stops when every block is done
    const int max_members = a.B - a.n_blocks + 1 < 1 ? 1 : (a.B - a.n_blocks + 1 > 8 ? 8 : a.B - a.n_blocks + 1);
    const int slots = a.cap + RUNS_EXTRA * max_members;
    g.total_slots = slots;
    g.final_slot = -1;
    gen_adam_runs_kernel<P><<<a.n_blocks, ADAM_RUNS_NT, 0, st>>>(a, g, 1);
    int launched = 1;
    for (int it = 0; it < slots; ++it) {
        g.final_slot = it;
        lin_prep_kernel<P, DC, OC, FIXED><<<a.B, 32, 0, st>>>(a, g, l);
        lin_runs_kernel<P, DC, OC, FIXED><<<(nthreads + 63) / 64, 64, 0, st>>>(a, g, l);
        gen_runs_reduce_kernel<P><<<dim3(a.B, g.nred), RUNS_RED_NT, 0, st>>>(a, g);
        gen_adam_runs_kernel<P><<<a.n_blocks, ADAM_RUNS_NT, 0, st>>>(a, g, 0);
        launched += 4;
        if ((it + 1) % RUNS_SLOT_CHUNK == 0 && it + 1 < slots && runs_all_done(g.ndone, a.n_blocks, st)) break;
    }
    if (getenv("EKS_DEBUG_RUNS")) {
        cudaStreamSynchronize(st);
        std::vector<int> w(a.B), nt(a.B);
        cudaMemcpy(w.data(), g.warm, a.B * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(nt.data(), l.ntr, a.B * sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[eks runs] linear steady optimiser: run_len %d nruns %d cap %d (warm, n_tr)", g.run_len, g.nruns,
                l.ncap);
        for (int b = 0; b < a.B && b < 16; ++b) fprintf(stderr, " (%d, %d)", w[b], nt[b]);
        fprintf(stderr, "\n");
    }
    note_launches(launched);
    runs_all_done(g.ndone, a.n_blocks, st);
    return check_launch("linear steady-state optimise kernels");
}

template <class P>
int generic_runs_optimize(const GArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    static_assert(sizeof(RunBlockState<P>) <= 128, "workspace bound");
    RunArgs<P> g;
    g.tol = runs_tol_host<P>();
    g.probe = a.nll_out != nullptr;
    g.lin_pinf = nullptr;
    const int n = a.sp.total;
    g.nruns = runs_geometry(n, a.B, g.run_len);
    g.ns = 2 * (a.D + a.D * a.D);
    const int dtype = sizeof(P) == 4 ? EKS_F32 : EKS_F64;
    EKS_REQUIRE(workspace && workspace_bytes >= generic_runs_optimize_workspace_bytes(dtype, a.n_blocks, a.B, a.D, a.T) +
                                                    (a.ncam == 0 ? linear_steady_workspace_bytes(dtype, a.B, a.D, a.O, a.T) : 0),
                "optimize_s: workspace too small for the run-parallel generic path");
    unsigned char* w = (unsigned char*)workspace;
    // Linear model with A = I, one span, long sequence: lag-statistics optimiser (lin_lag.cu) -- one pass over the
    // observations and one persistent launch; its scratch lies behind the run-parallel path's, which stays the
    // fallback for blocks where the closed form does not apply.
    if (a.ncam == 0 && !g.probe && lin_lag_applicable(dtype, a.D, a.O, a.sp.n, n)) {
        const size_t front = ((generic_runs_optimize_workspace_bytes(dtype, a.n_blocks, a.B, a.D, a.T) +
                               linear_steady_workspace_bytes(dtype, a.B, a.D, a.O, a.T)) + 255) / 256 * 256;
        if (workspace_bytes > front) {
            int used = 0;
            if (int rc = lin_lag_optimize<P>(a, w + front, workspace_bytes - front, st, &used)) return rc;
            if (used) return 0;
        }
    }
    auto take = [&](size_t bytes) { unsigned char* p = w; w += (bytes + 255) / 256 * 256; return p; };
    g.bstate = (RunBlockState<P>*)take((size_t)a.n_blocks * 128);
    g.warm = (int*)take((size_t)a.B * sizeof(int));
    int* seq_block = (int*)take((size_t)a.B * sizeof(int));
    g.seq_block = seq_block;
    g.part = (double*)take((size_t)a.B * g.nruns * 3 * sizeof(double));
    g.bnd_start = (P*)take((size_t)a.B * g.nruns * g.ns * sizeof(P));
    g.bnd_end = (P*)take((size_t)a.B * g.nruns * g.ns * sizeof(P));
    g.nred = (g.nruns + RUNS_RED_NT - 1) / RUNS_RED_NT;
    g.part2 = (double*)take((size_t)a.B * g.nred * 4 * sizeof(double));
    g.flag = nullptr;
    g.ndone = (int*)take(2 * sizeof(int));
    cudaMemsetAsync(g.ndone, 0, 2 * sizeof(int), st);
    note_unverified(0);
    cudaMemsetAsync(seq_block, 0xFF, (size_t)a.B * sizeof(int), st);
    const int nmax = a.B > a.n_blocks ? a.B : a.n_blocks;
    gen_runs_init_kernel<<<(nmax + 127) / 128, 128, 0, st>>>(a.B, a.n_blocks, a.block_off, a.members, seq_block,
                                                             g.warm, RUNS_W0);
    const int D = a.D, O = a.O;
    if (a.ncam == 0 && !getenv("EKS_NO_STEADY")) {   // linear model, constant R: steady-state mean recursion
        LinArgs<P> l;
        l.ent = O * (2 * D + 2);
        l.ncap = lin_table_cap(dtype, a.B, D, O, n);
        l.table = (P*)take((size_t)a.B * (l.ncap + 1) * l.ent * sizeof(P));
        l.pinf = (P*)take((size_t)a.B * g.ns * sizeof(P));
        g.lin_pinf = l.pinf;
        l.ntr = (int*)take((size_t)a.B * sizeof(int));
        l.cnll = (double*)take((size_t)a.B * 3 * sizeof(double));
        if (D == 2 && O == 2) return lin_optimize_launch<P, 2, 2, true>(a, g, l, st);
        if (D == 3 && O == 4) return lin_optimize_launch<P, 3, 4, true>(a, g, l, st);
        if (D == 3 && O == 6) return lin_optimize_launch<P, 3, 6, true>(a, g, l, st);
        if (D == 3 && O == 8) return lin_optimize_launch<P, 3, 8, true>(a, g, l, st);
        return lin_optimize_launch<P, EKS_MAX_STATE, EKS_MAX_CHAN, false>(a, g, l, st);
    }
    if (a.ncam > 0) {
        if (O == 4) return runs_optimize_launch<P, 3, 4, true, true>(a, g, st);
        if (O == 6) return runs_optimize_launch<P, 3, 6, true, true>(a, g, st);
        if (O == 8) return runs_optimize_launch<P, 3, 8, true, true>(a, g, st);
        return runs_optimize_launch<P, 3, EKS_MAX_CHAN, false, true>(a, g, st);
    }
    if (D == 2 && O == 2) return runs_optimize_launch<P, 2, 2, true, false>(a, g, st);
    if (D == 3 && O == 4) return runs_optimize_launch<P, 3, 4, true, false>(a, g, st);
    if (D == 3 && O == 6) return runs_optimize_launch<P, 3, 6, true, false>(a, g, st);
    if (D == 3 && O == 8) return runs_optimize_launch<P, 3, 8, true, false>(a, g, st);
    return runs_optimize_launch<P, EKS_MAX_STATE, EKS_MAX_CHAN, false, false>(a, g, st);
}
template int generic_runs_optimize<float>(const GArgs<float>&, void*, size_t, cudaStream_t);
template int generic_runs_optimize<double>(const GArgs<double>&, void*, size_t, cudaStream_t);


// eks_nll_grad for long sequences: one verified run-parallel evaluation per sequence at the given s (the machinery of
// generic_runs_optimize with singleton blocks and no Adam step).  The scratch is stream-ordered (cudaMallocAsync).
__global__ void runs_iota_kernel(int B, int* __restrict__ block_off, int* __restrict__ members) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= B) block_off[i] = i;
    if (i < B) members[i] = i;
}

template <class P>
int generic_runs_nll_grad(const GArgs<P>& a_in, cudaStream_t st) {
    GArgs<P> a = a_in;
    const int dtype = sizeof(P) == 4 ? EKS_F32 : EKS_F64;
    const size_t ws_bytes = generic_runs_optimize_workspace_bytes(dtype, a.B, a.B, a.D, a.T) +
                            (a.ncam == 0 ? linear_steady_workspace_bytes(dtype, a.B, a.D, a.O, a.T) : 0);
    const size_t idx_bytes = ((size_t)(2 * a.B + 1) * sizeof(int) + 255) / 256 * 256;
    unsigned char* buf = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&buf, ws_bytes + idx_bytes, st);
    if (e != cudaSuccess) { set_error("nll_grad: scratch allocation failed: %s", cudaGetErrorString(e)); return (int)e; }
    int* block_off = (int*)buf;
    int* members = block_off + a.B + 1;
    runs_iota_kernel<<<(a.B + 128) / 128, 128, 0, st>>>(a.B, block_off, members);
    a.n_blocks = a.B; a.block_off = block_off; a.members = members;
    a.s_log0 = nullptr; a.cap = 1; a.trace = nullptr; a.trace_cap = 0;
    a.lr = P(1); a.lo = P(-8); a.hi = P(8); a.tol = P(0);
    const int rc = generic_runs_optimize<P>(a, buf + idx_bytes, ws_bytes, st);
    cudaFreeAsync(buf, st);
    return rc;
}
template int generic_runs_nll_grad<float>(const GArgs<float>&, cudaStream_t);
template int generic_runs_nll_grad<double>(const GArgs<double>&, cudaStream_t);

// =====================================================================================================
// IBL pupil model (eks/ibl_pupil_smoother.py:363-607): 3 states [diameter, com_x, com_y], AR(1) dynamics
// A = diag(s_d, s_c, s_c), Q = diag(var_i (1 - s_i^2)), 8 observations through a fixed C, time-varying diagonal
// R_t in the loss, two parameters u -> s = sigmoid(u)(1 - 2e-3) + 1e-3, optax.adam(lr) on u, relative-tolerance
// stop rule, cap 5000.  One thread per (session, run, parameter direction): forward-mode dual on u_k.
// Same verified run-parallel scheme as above (nruns = 1 for short sequences = the exact sequential filter).
// =====================================================================================================
template <class P>
struct PupilState {
    P u[2], mu[2], nu[2], prev, s[2];
    int iters, done, redo;
};

template <class P>
struct PupilArgs {
    int B, T, n;                    // n = cropped frames of the loss
    GSpans sp;
    const P *m0, *S0, *C, *var3;    // [B][3], [B][3][3], [B][8][3], [B][3]
    PlaneView y, var;
    const P* ymean;                 // [B][8] or null
    P lr, tol, btol;
    int cap;
    int run_len, nruns, slot, total_slots;
    PupilState<P>* st;              // [B]
    int* warm;                      // [B]
    double* part;                   // [B][nruns][2 dirs][3]
    P* bnd_start;                   // [B][nruns][2][24]
    P* bnd_end;
    P *u_out, *s_out, *last_loss_out;
    int* iters_out;
    P* trace;
    int trace_cap;
};

template <class S, class P>
__device__ inline void pupil_AQ(const S u[2], const P* var3, S* Ad, S* Qd) {
    S sv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const S ex(exp_(-u[i].v), -u[i].d * exp_(-u[i].v));
        const S sg = S(P(1)) / (S(P(1)) + ex);                          // jax.nn.sigmoid
        sv[i] = sg * S(P(1) - P(2) * P(1e-3)) + S(P(1e-3));             // _to_stable_s
    }
    const S sd[3] = {sv[0], sv[1], sv[1]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Ad[i] = sd[i];
        Qd[i] = S(var3[i]) * (S(P(1)) - sd[i] * sd[i]);
    }
}

template <class P>
__global__ void __launch_bounds__(32) pupil_nll_runs_kernel(const __grid_constant__ PupilArgs<P> a) {
    using S = Dual<P>;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.nruns * 2) return;
    const int k = idx & 1, br = idx >> 1;
    const int b = br / a.nruns, r = br - b * a.nruns;
    if (a.st[b].done) return;
    const int n = a.n;
    const int t0 = r * a.run_len, t1 = min(n, t0 + a.run_len);
    double* part = a.part + (((long long)b * a.nruns + r) * 2 + k) * 3;
    if (t0 >= n) { part[0] = 0; part[1] = 0; part[2] = 0; return; }
    const int start = max(0, t0 - a.warm[b]);
    Dims<3, 8, true> dm{3, 8};
    SeqModel<P> mdl{3, 8, 0, a.m0 + b * 3, a.S0 + b * 9, nullptr, nullptr, a.C + b * 24, nullptr};
    SeqObs<P> ob;
    ob.y_base = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride;
    ob.y_off = a.y.chan_off;
    ob.ymean = a.ymean ? a.ymean + b * 8 : nullptr;
    ob.var_base = reinterpret_cast<const P*>(a.var.base) + (long long)b * a.var.seq_stride;
    ob.var_off = a.var.chan_off;
    ob.Rconst = nullptr;
    ob.var_floor = P(1e-12);
    FrameMap fm{a.sp};
    const S u[2] = {S(a.st[b].u[0], k == 0 ? P(1) : P(0)), S(a.st[b].u[1], k == 1 ? P(1) : P(0))};
    S Ad[3], Qd[3];
    pupil_AQ<S, P>(u, a.var3 + b * 3, Ad, Qd);
    S m[3], Pm[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) m[i] = S(mdl.m0[i]);
#pragma unroll
    for (int i = 0; i < 9; ++i) Pm[i] = S(mdl.S0[i]);
    S nll = S(P(0));
    bool ok = true;
    P* bs = a.bnd_start + (((long long)b * a.nruns + r) * 2 + k) * 24;
    P* be = a.bnd_end + (((long long)b * a.nruns + r) * 2 + k) * 24;
    for (int i = start; i < t1; ++i) {
        if (i == t0) {
            for (int q = 0; q < 3; ++q) { bs[q] = m[q].v; bs[12 + q] = m[q].d; }
            for (int q = 0; q < 9; ++q) { bs[3 + q] = Pm[q].v; bs[15 + q] = Pm[q].d; }
            nll = S(P(0));
            ok = true;
        }
        P yv[8], rv[8];
        load_obs<P, 8>(ob, 8, fm(i), yv, rv);
        ok = ekf_step<S, P, 3, 8, true, false>(dm, mdl, yv, rv, S(P(1)), m, Pm, nll, (S*)nullptr, (S*)nullptr, Ad, Qd) && ok;
    }
    for (int q = 0; q < 3; ++q) { be[q] = m[q].v; be[12 + q] = m[q].d; }
    for (int q = 0; q < 9; ++q) { be[3 + q] = Pm[q].v; be[15 + q] = Pm[q].d; }
    part[0] = (double)nll.v;
    part[1] = (double)nll.d;
    part[2] = ok ? 0.0 : 1.0;
}

// one warp per session: lanes stride over the runs (boundary check, fixed-order partial sums), lane 0 steps Adam
template <class P>
__global__ void __launch_bounds__(32) pupil_adam_kernel(const __grid_constant__ PupilArgs<P> a, int first) {
    const int b = blockIdx.x, lane = threadIdx.x;
    PupilState<P>& st = a.st[b];
    if (first) {
        if (lane == 0) {
            const float s0[2] = {0.99f, 0.98f};
            for (int i = 0; i < 2; ++i) { st.u[i] = P(logf(s0[i] / (1.0f - s0[i]))); st.mu[i] = P(0); st.nu[i] = P(0); }
            st.prev = P(INFINITY);
            st.iters = 0; st.redo = 0;
            st.done = (a.cap <= 0);
            a.warm[b] = 64;
        }
        return;
    }
    if (st.done) return;
    const int n = a.n, warm = a.warm[b];
    bool verified = true;
    for (int r = 1 + lane; r < a.nruns; r += 32) {
        const int t0 = r * a.run_len;
        if (t0 >= n || t0 - warm <= 0) continue;
        for (int k = 0; k < 2; ++k)
            verified = runs_boundary_ok<P>(a.bnd_end + (((long long)b * a.nruns + r - 1) * 2 + k) * 24,
                                           a.bnd_start + (((long long)b * a.nruns + r) * 2 + k) * 24, 3, P(1), a.btol) && verified;
    }
    verified = __all_sync(0xffffffffu, verified);
    if (sizeof(P) == 4 && warm >= RUNS_WMAX32) verified = true;   // fp32 rounding floor
    const bool last_slot = (a.slot == a.total_slots - 1);
    if (!verified && !last_slot) {
        if (lane == 0) {
            a.warm[b] = (warm >= n / 4) ? n : warm * 4;
            st.redo += 1;
        }
        return;
    }
    double v = 0, g0 = 0, g1 = 0, bad = 0;
    for (int r = lane; r < a.nruns; r += 32) {
        const double* p = a.part + ((long long)b * a.nruns + r) * 6;
        v += p[0]; g0 += p[1]; g1 += p[4]; bad += p[2] + p[5];
    }
    v = warp_sum(v); g0 = warp_sum(g0); g1 = warp_sum(g1); bad = warp_sum(bad);
    if (lane != 0) return;
    P loss = (P)v, gr[2] = {(P)g0, (P)g1};
    if (bad > 0) { loss = P(NAN); gr[0] = gr[1] = P(NAN); }   // the reference has no finite-guard here
    if (a.trace && st.iters < a.trace_cap) {
        P* tr = a.trace + ((long long)b * a.trace_cap + st.iters) * 3;
        tr[0] = st.u[0]; tr[1] = st.u[1]; tr[2] = loss;
    }
    const P b1 = P(0.9), b2 = P(0.999), eps = P(1e-8);
    const int count = st.iters + 1;
    for (int i = 0; i < 2; ++i) {
        st.mu[i] = b1 * st.mu[i] + (P(1) - b1) * gr[i];
        st.nu[i] = b2 * st.nu[i] + (P(1) - b2) * gr[i] * gr[i];
        const P mh = st.mu[i] / (P(1) - pow_(b1, P(count)));
        const P nh = st.nu[i] / (P(1) - pow_(b2, P(count)));
        st.u[i] = st.u[i] - a.lr * mh / (sqrt_(nh) + eps);
    }
    const P pm = st.prev > P(1e-12) ? st.prev : P(1e-12);
    const P rel_tol = a.tol * fabs(log_(pm));
    const bool stop = isfinite((double)st.prev) ? (fabs(loss - st.prev) < rel_tol + P(1e-6)) : false;
    st.prev = loss;
    st.iters = count;
    if (stop || count >= a.cap || last_slot) {
        st.done = 1;
        for (int i = 0; i < 2; ++i) {
            const P sg = P(1) / (P(1) + exp_(-st.u[i]));
            a.u_out[b * 2 + i] = st.u[i];
            a.s_out[b * 2 + i] = sg * (P(1) - P(2) * P(1e-3)) + P(1e-3);
        }
        a.last_loss_out[b] = st.prev;
        a.iters_out[b] = st.iters;
    }
}

template <class P>
__global__ void pupil_done_kernel(const PupilState<P>* st, int B, int* flag) {
    int all = 1;
    for (int b = threadIdx.x; b < B; b += 32) all &= st[b].done;
    all = __all_sync(0xffffffffu, all);
    if (threadIdx.x == 0) *flag = all;
}

static int pupil_geometry(int n, int B, int& run_len) {
    // the evaluation is latency-bound (one thread walks run_len + warm-up frames, ~5000 evaluations possible), so
    // runs are as short as the 64-frame warm-up allows; at most ~128k threads in flight
    if (n < 256) { run_len = n; return 1; }
    const long long cap_threads = 131072;
    long long rl = ((long long)n * 2 * B + cap_threads - 1) / cap_threads;
    if (rl < 64) rl = 64;
    run_len = (int)((rl + 31) / 32 * 32);
    return (n + run_len - 1) / run_len;
}

size_t pupil_optimize_workspace_bytes(int dtype, int B, int T) {
    const size_t w = dtype == EKS_F32 ? 4 : 8;
    int run_len;
    const int nruns = pupil_geometry(T, B, run_len);
    return 2048 + (size_t)B * 256 + (size_t)B * 4 + 256 + (size_t)B * nruns * 6 * sizeof(double) + 256 +
           2 * ((size_t)B * nruns * 2 * 24 * w + 256);
}

template <class P>
int pupil_optimize_run(PupilArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    static_assert(sizeof(PupilState<P>) <= 256, "workspace bound");
    a.nruns = pupil_geometry(a.n, a.B, a.run_len);
    a.btol = runs_tol_host<P>();
    const int dtype = sizeof(P) == 4 ? EKS_F32 : EKS_F64;
    EKS_REQUIRE(workspace && workspace_bytes >= pupil_optimize_workspace_bytes(dtype, a.B, a.T),
                "pupil_optimize: workspace too small");
    unsigned char* w = (unsigned char*)workspace;
    auto take = [&](size_t bytes) { unsigned char* p = w; w += (bytes + 255) / 256 * 256; return p; };
    a.st = (PupilState<P>*)take((size_t)a.B * 256);
    a.warm = (int*)take((size_t)a.B * 4);
    a.part = (double*)take((size_t)a.B * a.nruns * 6 * sizeof(double));
    a.bnd_start = (P*)take((size_t)a.B * a.nruns * 2 * 24 * sizeof(P));
    a.bnd_end = (P*)take((size_t)a.B * a.nruns * 2 * 24 * sizeof(P));
    const int nthreads = a.B * a.nruns * 2;
    a.total_slots = a.cap + (a.nruns > 1 ? RUNS_EXTRA : 0);
    a.slot = -1;
    pupil_adam_kernel<P><<<a.B, 32, 0, st>>>(a, 1);
    // NOTE: the reference's cap is 5000 evaluations; the loop is unrolled on the stream in chunks and the host
    // checks a completion flag between chunks so that converged problems do not pay for thousands of no-op
    // launches (this entry point therefore synchronises the stream).
    int h_done = 0;
    int* d_flag = (int*)take(256);
    const int chunk = 64;
    for (int it = 0; it < a.total_slots; ++it) {
        a.slot = it;
        pupil_nll_runs_kernel<P><<<(nthreads + 31) / 32, 32, 0, st>>>(a);
        pupil_adam_kernel<P><<<a.B, 32, 0, st>>>(a, 0);
        if ((it + 1) % chunk == 0 || it + 1 == a.total_slots) {
            pupil_done_kernel<P><<<1, 32, 0, st>>>(a.st, a.B, d_flag);
            cudaError_t e = cudaMemcpyAsync(&h_done, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error("pupil_optimize: %s", cudaGetErrorString(e)); return (int)e; }
            if (h_done) break;
        }
    }
    return check_launch("pupil optimise kernels");
}

size_t generic_runs_smooth_extra_bytes(int dtype, int B, int D, int T) {
    int run_len;
    const int nruns = runs_geometry(T, B, run_len);
    return 3 * (size_t)B * nruns * (D + D * D) * (dtype == EKS_F32 ? 4 : 8) + 256;
}

}  // namespace eks

using namespace eks;

extern "C" size_t eks_pupil_optimize_workspace_bytes(int dtype, int B, int T) {
    return pupil_optimize_workspace_bytes(dtype, B, T);
}

extern "C" int eks_pupil_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* C,
                                  const void* var3, const void* y_base, long long y_seq_stride,
                                  const long long* y_off, const void* ymean, const void* var_base,
                                  long long var_seq_stride, const long long* var_off, int n_spans,
                                  const int* span_start, const int* span_end, double lr, double tol, int cap,
                                  void* u_out, void* s_out, void* last_loss_out, int* iters_out, void* trace,
                                  int trace_cap, void* workspace, size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(m0 && S0 && C && var3 && y_base && y_off && var_base && var_off && u_out && s_out && last_loss_out &&
                    iters_out, "pupil_optimize: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && cap >= 0, "pupil_optimize: bad dims");
    GSpans sp;
    if (n_spans <= 0) {
        sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = T; sp.total = T;
    } else {
        EKS_REQUIRE(n_spans <= G_MAX_SPANS, "at most %d frame spans supported on device", G_MAX_SPANS);
        sp.n = n_spans; sp.cum[0] = 0;
        for (int i = 0; i < n_spans; ++i) {
            EKS_REQUIRE(span_start[i] >= 0 && span_end[i] <= T && span_start[i] < span_end[i], "bad span %d", i);
            sp.start[i] = span_start[i];
            sp.cum[i + 1] = sp.cum[i] + (span_end[i] - span_start[i]);
        }
        sp.total = sp.cum[n_spans];
    }
#define EKS_FILL(PT)                                                                                           \
    PupilArgs<PT> a;                                                                                           \
    memset(&a, 0, sizeof(a));                                                                                  \
    a.B = B; a.T = T; a.n = sp.total; a.sp = sp;                                                               \
    a.m0 = (const PT*)m0; a.S0 = (const PT*)S0; a.C = (const PT*)C; a.var3 = (const PT*)var3;                  \
    a.y.base = y_base; a.y.seq_stride = y_seq_stride; a.var.base = var_base; a.var.seq_stride = var_seq_stride; \
    for (int i = 0; i < MAX_CHAN; ++i) { a.y.chan_off[i] = i < 8 ? y_off[i] : 0; a.var.chan_off[i] = i < 8 ? var_off[i] : 0; } \
    a.ymean = (const PT*)ymean; a.lr = (PT)lr; a.tol = (PT)tol; a.cap = cap;                                   \
    a.u_out = (PT*)u_out; a.s_out = (PT*)s_out; a.last_loss_out = (PT*)last_loss_out; a.iters_out = iters_out; \
    a.trace = (PT*)trace; a.trace_cap = trace_cap;                                                             \
    return pupil_optimize_run<PT>(a, workspace, workspace_bytes, (cudaStream_t)stream);
    if (dtype == EKS_F32) { EKS_FILL(float) }
    EKS_FILL(double)
#undef EKS_FILL
}
