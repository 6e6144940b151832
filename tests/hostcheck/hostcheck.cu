// hostcheck.cu -- runs the library's __host__ __device__ per-sequence templates (ekf_generic.cuh) on
// the CPU so that the kernel arithmetic can be compared with the oracle WITHOUT a GPU.  Test-only.
#include "../../eks_b200/csrc/common.cuh"
#include "../../eks_b200/csrc/ekf_generic.cuh"

using namespace eks;
namespace eks { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

struct HostFrame { __host__ __device__ long long operator()(int i) const { return i; } };

template <class P, int DC, int OC, bool FIXED, bool NL>
static void run_nll(int D, int O, int T, const P* m0, const P* S0, const P* A, const P* Q, const P* C, int ncam,
                    const P* cams, const P* y /*O planes of T*/, const P* Rconst, P s, P* nll, P* dnll) {
    Dims<DC, OC, FIXED> dm{D, O};
    SeqModel<P> mdl{D, O, ncam, m0, S0, A, Q, C, cams};
    long long off[MAX_CHAN];
    for (int o = 0; o < O; ++o) off[o] = (long long)o * T;
    SeqObs<P> ob{y, nullptr, off, nullptr, nullptr, Rconst, P(1e-12)};
    seq_nll_grad<P, DC, OC, FIXED, NL>(dm, mdl, ob, T, HostFrame(), s, nll, dnll);
}
template <class P, int DC, int OC, bool FIXED, bool NL>
static void run_smooth(int D, int O, int T, const P* m0, const P* S0, const P* A, const P* Q, const P* C, int ncam,
                       const P* cams, const P* y, const P* var, P s, P* mf, P* Pf, P* ms, P* Vs) {
    Dims<DC, OC, FIXED> dm{D, O};
    SeqModel<P> mdl{D, O, ncam, m0, S0, A, Q, C, cams};
    long long off[MAX_CHAN];
    for (int o = 0; o < O; ++o) off[o] = (long long)o * T;
    SeqObs<P> ob{y, var, off, off, nullptr, nullptr, P(1e-12)};
    seq_smooth<P, DC, OC, FIXED, NL>(dm, mdl, ob, T, s, mf, Pf, ms, Vs);
}


template <class P>
static void nll_grad_any(int D, int O, int T, const P* m0, const P* S0, const P* A, const P* Q, const P* C, int ncam,
                         const P* cams, const P* y, const P* Rconst, P s, P* nll, P* dnll) {
#define ARGS D, O, T, m0, S0, A, Q, C, ncam, cams, y, Rconst, s, nll, dnll
    if (ncam > 0) {
        if (O == 4) run_nll<P, 3, 4, true, true>(ARGS);
        else if (O == 6) run_nll<P, 3, 6, true, true>(ARGS);
        else run_nll<P, 3, 16, false, true>(ARGS);
    } else if (D == 2 && O == 2) run_nll<P, 2, 2, true, false>(ARGS);
    else if (D == 3 && O == 4) run_nll<P, 3, 4, true, false>(ARGS);
    else run_nll<P, 6, 16, false, false>(ARGS);
#undef ARGS
}
template <class P>
static void smooth_any(int D, int O, int T, const P* m0, const P* S0, const P* A, const P* Q, const P* C, int ncam,
                       const P* cams, const P* y, const P* var, P s, P* mf, P* Pf, P* ms, P* Vs) {
#define ARGS D, O, T, m0, S0, A, Q, C, ncam, cams, y, var, s, mf, Pf, ms, Vs
    if (ncam > 0) {
        if (O == 4) run_smooth<P, 3, 4, true, true>(ARGS);
        else if (O == 6) run_smooth<P, 3, 6, true, true>(ARGS);
        else run_smooth<P, 3, 16, false, true>(ARGS);
    } else if (D == 2 && O == 2) run_smooth<P, 2, 2, true, false>(ARGS);
    else if (D == 3 && O == 4) run_smooth<P, 3, 4, true, false>(ARGS);
    else run_smooth<P, 6, 16, false, false>(ARGS);
#undef ARGS
}
template <class P>
static void project_any(int ncam, const P* cams, int N, const P* X, P* uv, P* J) {
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < ncam; ++c)
            project_cam_jac<P, P>(cams + c * CAM_STRIDE, X + 3 * n, uv + ((size_t)n * ncam + c) * 2,
                                  J + ((size_t)n * ncam + c) * 6);
}
template <class P>
static void adam_any(int n, const P* loss, const P* g, P s_log0, P lr, P tol, int cap, P* s_log_out, int* iters_out) {
    AdamState<P> a;
    adam_init(a, s_log0);
    for (int i = 0; i < n && !a.done; ++i) adam_step(a, loss[i], g[i], lr, tol, cap);
    *s_log_out = a.s_log;
    *iters_out = a.iters;
}

#define HOSTCHECK_API(SFX, PT)                                                                                       \
    extern "C" void hostcheck_nll_grad_##SFX(int D, int O, int T, const PT* m0, const PT* S0, const PT* A,           \
                                             const PT* Q, const PT* C, int ncam, const PT* cams, const PT* y,        \
                                             const PT* Rconst, PT s, PT* nll, PT* dnll) {                            \
        nll_grad_any<PT>(D, O, T, m0, S0, A, Q, C, ncam, cams, y, Rconst, s, nll, dnll);                             \
    }                                                                                                                \
    extern "C" void hostcheck_smooth_##SFX(int D, int O, int T, const PT* m0, const PT* S0, const PT* A,             \
                                           const PT* Q, const PT* C, int ncam, const PT* cams, const PT* y,          \
                                           const PT* var, PT s, PT* mf, PT* Pf, PT* ms, PT* Vs) {                    \
        smooth_any<PT>(D, O, T, m0, S0, A, Q, C, ncam, cams, y, var, s, mf, Pf, ms, Vs);                             \
    }                                                                                                                \
    extern "C" void hostcheck_project_##SFX(int ncam, const PT* cams, int N, const PT* X, PT* uv, PT* J) {           \
        project_any<PT>(ncam, cams, N, X, uv, J);                                                                    \
    }                                                                                                                \
    extern "C" void hostcheck_adam_##SFX(int n, const PT* loss, const PT* g, PT s_log0, PT lr, PT tol, int cap,      \
                                         PT* s_log_out, int* iters_out) {                                            \
        adam_any<PT>(n, loss, g, s_log0, lr, tol, cap, s_log_out, iters_out);                                        \
    }

HOSTCHECK_API(f32, float)
HOSTCHECK_API(f64, double)

// ---- triangulation templates (triangulate.cuh) on the host: compared with cv2 in tests/test_host.py
#include "../../eks_b200/csrc/triangulate.cuh"
extern "C" void hc_undistort(const double* cam, int n, const double* uv, double* out) {
    for (int i = 0; i < n; ++i) undistort_point(cam, uv[2 * i], uv[2 * i + 1], out[2 * i], out[2 * i + 1]);
}
extern "C" void hc_triangulate_points(const double* cams, int V, int n, const double* uv /*[n][2V]*/, double* X /*[n][3]*/) {
    for (int i = 0; i < n; ++i) triangulate_point(cams, V, uv + (size_t)i * 2 * V, X + (size_t)i * 3);
}
