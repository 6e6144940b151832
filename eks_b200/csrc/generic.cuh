// generic.cuh -- argument block and helpers shared by the generic-model kernels (generic.cu: one sequence
// per thread; generic_runs.cu: one run of frames per thread).
#pragma once
#include "common.cuh"
#include "ekf_generic.cuh"

namespace eks {

constexpr int G_MAX_SPANS = 16;
struct GSpans {
    int n, total;
    int start[G_MAX_SPANS];
    int cum[G_MAX_SPANS + 1];
};
struct FrameMap {
    GSpans sp;
    __device__ long long operator()(int i) const {
        if (sp.n == 1) return sp.start[0] + i;
        int j = 0;
        while (j + 1 < sp.n && i >= sp.cum[j + 1]) ++j;
        return sp.start[j] + (i - sp.cum[j]);
    }
};

template <class P>
struct GArgs {
    int B, D, O, T, ncam;
    const P *m0, *S0, *A, *Q, *C, *cams;
    PlaneView y, var;
    const P *ymean, *Rconst;
    GSpans sp;
    // optimise
    int n_blocks;
    const int *block_off, *members;
    const P* s_log0;
    P lr, lo, hi, tol;
    int cap;
    P *s_log_out, *last_loss_out;
    int* iters_out;
    P* trace;
    int trace_cap;
    // nll_grad / smooth
    const P* s;
    P *nll_out, *dnll_out;
    P *mf, *Pf, *ms, *Vs;
};

template <class P>
__device__ inline void make_seq(const GArgs<P>& a, int b, SeqModel<P>& mdl, SeqObs<P>& ob, bool use_var) {
    const int D = a.D, O = a.O;
    mdl.D = D; mdl.O = O; mdl.ncam = a.ncam;
    mdl.m0 = a.m0 + (long long)b * D;
    mdl.S0 = a.S0 + (long long)b * D * D;
    mdl.A = a.A + (long long)b * D * D;
    mdl.Q = a.Q + (long long)b * D * D;
    mdl.C = a.C ? a.C + (long long)b * O * D : nullptr;
    mdl.cams = a.cams;
    ob.y_base = reinterpret_cast<const P*>(a.y.base) + (long long)b * a.y.seq_stride;
    ob.y_off = a.y.chan_off;
    ob.ymean = a.ymean ? a.ymean + (long long)b * O : nullptr;
    if (use_var) {
        ob.var_base = reinterpret_cast<const P*>(a.var.base) + (long long)b * a.var.seq_stride;
        ob.var_off = a.var.chan_off;
        ob.Rconst = nullptr;
    } else {
        ob.var_base = nullptr;
        ob.var_off = nullptr;
        ob.Rconst = a.Rconst + (long long)b * O;
    }
    ob.var_floor = P(1e-12);
}


// run-parallel variants (generic_runs.cu)
template <class P>
int generic_runs_optimize(const GArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t generic_runs_optimize_workspace_bytes(int dtype, int n_blocks, int B, int D, int T);
template <class P>
int generic_runs_nll_grad(const GArgs<P>& a, cudaStream_t st);
size_t linear_steady_workspace_bytes(int dtype, int B, int D, int O, int T);
template <class P>
int generic_runs_smooth(const GArgs<P>& a, cudaStream_t st);
size_t generic_runs_smooth_extra_bytes(int dtype, int B, int D, int T);
constexpr int GEN_RUNS_MIN_FRAMES = 512;   // below this the sequential per-sequence kernels are used

}  // namespace eks
