// generic.cu -- generic-dimension kernels: one sequence (or one shared-s block) per thread, sequential
// in time.  Covers every model family of the path (singlecam 2-D, multicam linear D=3..6, calibrated
// pinhole EKF) and is the correctness anchor for the specialised time-parallel kernels (diag.cu).
#include <cstdarg>
#include "common.cuh"
#include "ekf_generic.cuh"
#include <cstring>
#include <cstdlib>
#include "generic.cuh"
#include "lin_lag.cuh"
#include "diag.cuh"
#include "../../include/eks_b200.h"

namespace eks {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local int g_unverified = 0;
void note_launches(int n) { g_launches = n; }
void note_unverified(int n) { g_unverified = n; }
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

template <class P, int DC, int OC, bool FIXED, bool NL>
__global__ void __launch_bounds__(32) nll_grad_generic_kernel(const __grid_constant__ GArgs<P> a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, false);
    FrameMap fm{a.sp};
    P v, g;
    seq_nll_grad<P, DC, OC, FIXED, NL>(dm, mdl, ob, a.sp.total, fm, a.s[b], &v, &g);
    a.nll_out[b] = v;
    a.dnll_out[b] = g;
}

template <class P, int DC, int OC, bool FIXED, bool NL>
__global__ void __launch_bounds__(32) optimize_generic_kernel(const __grid_constant__ GArgs<P> a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_blocks) return;
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    FrameMap fm{a.sp};
    AdamState<P> ad;
    adam_init(ad, a.s_log0[j]);
    while (!ad.done) {
        P dsdlog;
        const P s = adam_current_s(ad, a.lo, a.hi, &dsdlog);
        P loss = P(0), g = P(0);
        for (int mi = a.block_off[j]; mi < a.block_off[j + 1]; ++mi) {
            SeqModel<P> mdl;
            SeqObs<P> ob;
            make_seq(a, a.members[mi], mdl, ob, false);
            P v, dv;
            seq_nll_grad<P, DC, OC, FIXED, NL>(dm, mdl, ob, a.sp.total, fm, s, &v, &dv);
            loss += v;
            g += dv * dsdlog;
        }
        if (a.trace && ad.iters < a.trace_cap) {
            P* tr = a.trace + ((long long)j * a.trace_cap + ad.iters) * 3;
            tr[0] = ad.s_log; tr[1] = loss; tr[2] = g * a.lr;
        }
        adam_step(ad, loss, g, a.lr, a.tol, a.cap);
    }
    a.s_log_out[j] = ad.s_log;
    a.last_loss_out[j] = ad.prev;
    a.iters_out[j] = ad.iters;
}

template <class P, int DC, int OC, bool FIXED, bool NL>
__global__ void __launch_bounds__(32) smooth_generic_kernel(const __grid_constant__ GArgs<P> a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    Dims<DC, OC, FIXED> dm{a.D, a.O};
    SeqModel<P> mdl;
    SeqObs<P> ob;
    make_seq(a, b, mdl, ob, true);
    const long long TD = (long long)a.T * a.D, TDD = TD * a.D;
    seq_smooth<P, DC, OC, FIXED, NL>(dm, mdl, ob, a.T, a.s[b], a.mf + b * TD, a.Pf + b * TDD, a.ms + b * TD,
                                     a.Vs + b * TDD);
}

enum GenericOp { OP_NLL, OP_OPT, OP_SMOOTH };

template <class P, int DC, int OC, bool FIXED, bool NL>
int launch_op(GenericOp op, const GArgs<P>& a, cudaStream_t st) {
    const int n = (op == OP_OPT) ? a.n_blocks : a.B;
    const int threads = 32, blocks = (n + threads - 1) / threads;
    if (op == OP_NLL) nll_grad_generic_kernel<P, DC, OC, FIXED, NL><<<blocks, threads, 0, st>>>(a);
    else if (op == OP_OPT) optimize_generic_kernel<P, DC, OC, FIXED, NL><<<blocks, threads, 0, st>>>(a);
    else smooth_generic_kernel<P, DC, OC, FIXED, NL><<<blocks, threads, 0, st>>>(a);
    return check_launch("generic kernel");
}

template <class P>
int dispatch(GenericOp op, const GArgs<P>& a, cudaStream_t st) {
    const int D = a.D, O = a.O;
    if (a.ncam > 0) {
        EKS_REQUIRE(D == 3 && O == 2 * a.ncam, "pinhole emission needs D == 3 and O == 2*ncam");
        if (O == 4) return launch_op<P, 3, 4, true, true>(op, a, st);
        if (O == 6) return launch_op<P, 3, 6, true, true>(op, a, st);
        if (O == 8) return launch_op<P, 3, 8, true, true>(op, a, st);
        return launch_op<P, 3, EKS_MAX_CHAN, false, true>(op, a, st);
    }
    if (D == 2 && O == 2) return launch_op<P, 2, 2, true, false>(op, a, st);
    if (D == 3 && O == 4) return launch_op<P, 3, 4, true, false>(op, a, st);
    if (D == 3 && O == 6) return launch_op<P, 3, 6, true, false>(op, a, st);
    if (D == 3 && O == 8) return launch_op<P, 3, 8, true, false>(op, a, st);
    return launch_op<P, EKS_MAX_STATE, EKS_MAX_CHAN, false, false>(op, a, st);
}

static int fill_spans(int T, int n_spans, const int* s0, const int* s1, GSpans& sp) {
    if (n_spans <= 0) {
        sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = T; sp.total = T;
        return 0;
    }
    EKS_REQUIRE(n_spans <= G_MAX_SPANS, "at most %d frame spans supported on device", G_MAX_SPANS);
    sp.n = n_spans; sp.cum[0] = 0;
    for (int i = 0; i < n_spans; ++i) {
        EKS_REQUIRE(s0[i] >= 0 && s1[i] <= T && s0[i] < s1[i], "bad span %d", i);
        EKS_REQUIRE(i == 0 || s0[i] >= s1[i - 1], "spans must be sorted and non-overlapping");
        sp.start[i] = s0[i];
        sp.cum[i + 1] = sp.cum[i] + (s1[i] - s0[i]);
    }
    sp.total = sp.cum[n_spans];
    return 0;
}

static PlaneView view_of(const void* base, long long seq_stride, const long long* off, int O) {
    PlaneView v;
    v.base = base; v.seq_stride = seq_stride;
    for (int i = 0; i < MAX_CHAN; ++i) v.chan_off[i] = (off && i < O) ? off[i] : 0;
    return v;
}

template <class P>
static int common_args(GArgs<P>& a, int B, int D, int O, int T, const void* m0, const void* S0, const void* A,
                       const void* Q, const void* C, int ncam, const void* cams) {
    EKS_REQUIRE(B >= 1 && T >= 1, "bad batch/frame count");
    EKS_REQUIRE(D >= 1 && D <= EKS_MAX_STATE, "latent dimension %d outside [1,%d]", D, EKS_MAX_STATE);
    EKS_REQUIRE(O >= 1 && O <= EKS_MAX_CHAN, "observation dimension %d outside [1,%d]", O, EKS_MAX_CHAN);
    EKS_REQUIRE(m0 && S0 && A && Q, "null model pointer");
    EKS_REQUIRE(ncam > 0 ? cams != nullptr : C != nullptr, "need C (linear) or cams (pinhole)");
    memset(&a, 0, sizeof(a));
    a.B = B; a.D = D; a.O = O; a.T = T; a.ncam = ncam;
    a.m0 = (const P*)m0; a.S0 = (const P*)S0; a.A = (const P*)A; a.Q = (const P*)Q; a.C = (const P*)C;
    a.cams = (const P*)cams;
    return 0;
}

template <class P>
int nll_grad_impl(int B, int D, int O, int T, const void* m0, const void* S0, const void* A, const void* Q,
                  const void* C, int ncam, const void* cams, const void* y_base, long long y_seq_stride,
                  const long long* y_off, const void* ymean, const void* Rconst, int n_spans, const int* s0,
                  const int* s1, const void* s, void* nll_out, void* dnll_out, cudaStream_t st) {
    GArgs<P> a;
    if (common_args<P>(a, B, D, O, T, m0, S0, A, Q, C, ncam, cams)) return -1;
    EKS_REQUIRE(y_base && y_off && Rconst && s && nll_out && dnll_out, "nll_grad: null pointer");
    a.y = view_of(y_base, y_seq_stride, y_off, O);
    a.ymean = (const P*)ymean; a.Rconst = (const P*)Rconst;
    if (fill_spans(T, n_spans, s0, s1, a.sp)) return -1;
    a.s = (const P*)s; a.nll_out = (P*)nll_out; a.dnll_out = (P*)dnll_out;
    if (a.sp.total >= GEN_RUNS_MIN_FRAMES) return generic_runs_nll_grad<P>(a, st);
    return dispatch<P>(OP_NLL, a, st);
}

template <class P>
int optimize_impl(int B, int D, int O, int T, const void* m0, const void* S0, const void* A, const void* Q,
                  const void* C, int ncam, const void* cams, const void* y_base, long long y_seq_stride,
                  const long long* y_off, const void* ymean, const void* Rconst, int n_spans, const int* s0,
                  const int* s1, int n_blocks, const int* block_off, const int* members, const void* s_log0,
                  double lr, double lo, double hi, double tol, int cap, void* s_log_out, void* last_loss_out,
                  int* iters_out, void* trace, int trace_cap, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
    GArgs<P> a;
    if (common_args<P>(a, B, D, O, T, m0, S0, A, Q, C, ncam, cams)) return -1;
    EKS_REQUIRE(y_base && y_off && Rconst && block_off && members && s_log0 && s_log_out && last_loss_out &&
                    iters_out, "optimize_s: null pointer");
    EKS_REQUIRE(n_blocks >= 1 && cap >= 0, "optimize_s: bad block count / cap");
    a.y = view_of(y_base, y_seq_stride, y_off, O);
    a.ymean = (const P*)ymean; a.Rconst = (const P*)Rconst;
    if (fill_spans(T, n_spans, s0, s1, a.sp)) return -1;
    a.n_blocks = n_blocks; a.block_off = block_off; a.members = members; a.s_log0 = (const P*)s_log0;
    a.lr = (P)lr; a.lo = (P)lo; a.hi = (P)hi; a.tol = (P)tol; a.cap = cap;
    a.s_log_out = (P*)s_log_out; a.last_loss_out = (P*)last_loss_out; a.iters_out = iters_out;
    a.trace = (P*)trace; a.trace_cap = trace_cap;
    // long sequences: verified run-parallel execution (generic_runs.cu); short ones: one thread per block
    if (a.sp.total >= GEN_RUNS_MIN_FRAMES) return generic_runs_optimize<P>(a, workspace, workspace_bytes, st);
    note_launches(1);
    return dispatch<P>(OP_OPT, a, st);
}

template <class P>
int smooth_impl(int B, int D, int O, int T, const void* m0, const void* S0, const void* A, const void* Q,
                const void* C, int ncam, const void* cams, const void* y_base, long long y_seq_stride,
                const long long* y_off, const void* ymean, const void* var_base, long long var_seq_stride,
                const long long* var_off, const void* s, void* ms_out, void* Vs_out, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
    GArgs<P> a;
    if (common_args<P>(a, B, D, O, T, m0, S0, A, Q, C, ncam, cams)) return -1;
    EKS_REQUIRE(y_base && y_off && var_base && var_off && s && ms_out && Vs_out, "filter_smooth: null pointer");
    const size_t need = (size_t)B * T * (D + D * D) * sizeof(P) +
                        (T >= GEN_RUNS_MIN_FRAMES ? generic_runs_smooth_extra_bytes(sizeof(P) == 4 ? EKS_F32 : EKS_F64, B, D, T) : 0);
    EKS_REQUIRE(workspace && workspace_bytes >= need, "filter_smooth: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    a.y = view_of(y_base, y_seq_stride, y_off, O);
    a.var = view_of(var_base, var_seq_stride, var_off, O);
    a.ymean = (const P*)ymean;
    if (fill_spans(T, 0, nullptr, nullptr, a.sp)) return -1;
    a.s = (const P*)s;
    a.mf = (P*)workspace;
    a.Pf = a.mf + (size_t)B * T * D;
    a.ms = (P*)ms_out; a.Vs = (P*)Vs_out;
    if (T >= GEN_RUNS_MIN_FRAMES) return generic_runs_smooth<P>(a, st);
    return dispatch<P>(OP_SMOOTH, a, st);
}

}  // namespace eks

using namespace eks;

extern "C" const char* eks_last_error(void) { return g_err; }
extern "C" int eks_last_launch_count(void) { return g_launches; }
extern "C" int eks_last_unverified_count(void) { return g_unverified; }
extern "C" int eks_version(void) { return 202; }

extern "C" int eks_nll_grad(int dtype, int B, int D, int O, int T, const void* m0, const void* S0, const void* A,
                            const void* Q, const void* C, int ncam, const void* cams, const void* y_base,
                            long long y_seq_stride, const long long* y_off, const void* ymean, const void* Rconst,
                            int n_spans, const int* s0, const int* s1, const void* s, void* nll_out, void* dnll_out,
                            void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        return nll_grad_impl<float>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean,
                                    Rconst, n_spans, s0, s1, s, nll_out, dnll_out, st);
    return nll_grad_impl<double>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean, Rconst,
                                 n_spans, s0, s1, s, nll_out, dnll_out, st);
}

extern "C" size_t eks_optimize_s_workspace_bytes(int dtype, int n_blocks, int B, int D, int O, int T) {
    const size_t a1 = diag_lag_workspace_bytes(dtype, n_blocks, B, T);   // >= diag_optimize_workspace_bytes
    const size_t a2 = T >= GEN_RUNS_MIN_FRAMES ? generic_runs_optimize_workspace_bytes(dtype, n_blocks, B, D, T) +
                                                     linear_steady_workspace_bytes(dtype, B, D, O, T) + 256 +
                                                     (lin_lag_applicable(dtype, D, O, 1, T)
                                                          ? lin_lag_workspace_bytes(dtype, n_blocks, B, O, T) : 0) : 0;
    return a1 > a2 ? a1 : a2;
}

extern "C" int eks_optimize_s(int dtype, int B, int D, int O, int T, const void* m0, const void* S0, const void* A,
                              const void* Q, const void* C, int ncam, const void* cams, const void* y_base,
                              long long y_seq_stride, const long long* y_off, const void* ymean, const void* Rconst,
                              int n_spans, const int* s0, const int* s1, int n_blocks, const int* block_off,
                              const int* members, const void* s_log0, double lr, double lo, double hi, double tol,
                              int cap, void* s_log_out, void* last_loss_out, int* iters_out, void* trace,
                              int trace_cap, int model_structure, void* workspace, size_t workspace_bytes,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    note_unverified(0);
    // model_structure == EKS_STRUCT_DIAG: the caller asserts D == O == 2 with diagonal A, C, Q, S0
    // (single-camera model) -> time-parallel persistent kernel (diag.cu); one contiguous span only
    if ((model_structure == EKS_STRUCT_DIAG || model_structure == EKS_STRUCT_DIAG_STREAM) && D == 2 && O == 2 &&
        ncam == 0 && n_spans <= 1) {
        EKS_REQUIRE(y_base && y_off && Rconst && block_off && members && s_log0 && s_log_out && last_loss_out &&
                        iters_out && m0 && S0 && A && Q && C, "optimize_s: null pointer");
        int t_begin = 0, n = T;
        if (n_spans == 1) {
            EKS_REQUIRE(s0[0] >= 0 && s1[0] <= T && s0[0] < s1[0], "bad span 0");
            t_begin = s0[0]; n = s1[0] - s0[0];
        }
        // default: lag-statistics optimiser (one pass over the observations, one persistent launch; diag_lag.cu).
        // EKS_OPT_MODE=stream selects the one-streaming-launch-per-evaluation loop of diag.cu (cross-check).
        static const bool stream_mode = [] { const char* e = getenv("EKS_OPT_MODE"); return e && !strcmp(e, "stream"); }();
        if (stream_mode || model_structure == EKS_STRUCT_DIAG_STREAM)
            return diag_optimize(dtype, B, T, m0, S0, A, Q, C, y_base, y_seq_stride, y_off, ymean, Rconst, t_begin, n,
                                 n_blocks, block_off, members, s_log0, lr, lo, hi, tol, cap, s_log_out, last_loss_out,
                                 iters_out, trace, trace_cap, workspace, workspace_bytes, st);
        return diag_lag_optimize(dtype, B, T, m0, S0, A, Q, C, y_base, y_seq_stride, y_off, ymean, Rconst, t_begin, n,
                                 n_blocks, block_off, members, s_log0, lr, lo, hi, tol, cap, s_log_out, last_loss_out,
                                 iters_out, trace, trace_cap, workspace, workspace_bytes, st);
    }
    if (dtype == EKS_F32)
        return optimize_impl<float>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean,
                                    Rconst, n_spans, s0, s1, n_blocks, block_off, members, s_log0, lr, lo, hi, tol,
                                    cap, s_log_out, last_loss_out, iters_out, trace, trace_cap, workspace,
                                    workspace_bytes, st);
    return optimize_impl<double>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean, Rconst,
                                 n_spans, s0, s1, n_blocks, block_off, members, s_log0, lr, lo, hi, tol, cap,
                                 s_log_out, last_loss_out, iters_out, trace, trace_cap, workspace, workspace_bytes, st);
}

extern "C" size_t eks_filter_smooth_workspace_bytes(int dtype, int B, int D, int T) {
    return (size_t)B * T * (D + D * D) * (dtype == EKS_F32 ? 4 : 8) +
           (T >= GEN_RUNS_MIN_FRAMES ? generic_runs_smooth_extra_bytes(dtype, B, D, T) : 0);
}

extern "C" int eks_filter_smooth(int dtype, int B, int D, int O, int T, const void* m0, const void* S0,
                                 const void* A, const void* Q, const void* C, int ncam, const void* cams,
                                 const void* y_base, long long y_seq_stride, const long long* y_off,
                                 const void* ymean, const void* var_base, long long var_seq_stride,
                                 const long long* var_off, const void* s, void* ms_out, void* Vs_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        return smooth_impl<float>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean,
                                  var_base, var_seq_stride, var_off, s, ms_out, Vs_out, workspace, workspace_bytes,
                                  st);
    return smooth_impl<double>(B, D, O, T, m0, S0, A, Q, C, ncam, cams, y_base, y_seq_stride, y_off, ymean, var_base,
                               var_seq_stride, var_off, s, ms_out, Vs_out, workspace, workspace_bytes, st);
}
