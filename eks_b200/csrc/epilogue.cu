// epilogue.cu -- reprojection of smoothed latent moments into camera planes.
// Replaces the per-keypoint Python loops of eks/multicam_smoother.py:450-511 (linear: C m + mean,
// diag(C V C^T) + ensemble variance) and :914-946 project_3d_covariance_to_2d (pinhole: h_c(m),
// diag(J V J^T) + ensemble variance), and eks/singlecam_smoother.py:189-217.
#include "common.cuh"
#include "ekf_generic.cuh"
#include "../../include/eks_b200.h"

namespace eks {

template <class P>
struct ReprojArgs {
    int B, T, D, V, ncam, var_quirk;
    const P *ms, *Vs, *C, *ymean, *cams;
    PlaneView var;
    int has_var;
    P* out;
    long long out_seq_stride, out_cam_stride;
    long long plane_off[4];
};

template <class P>
__global__ void __launch_bounds__(256) reproject_kernel(const __grid_constant__ ReprojArgs<P> a) {
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;      // problems on grid x: no 65535 limit
    if (t >= a.T) return;
    const int D = a.D;
    P m[EKS_MAX_STATE], Vm[EKS_MAX_STATE * EKS_MAX_STATE];
    const P* mp = a.ms + ((long long)b * a.T + t) * D;
    const P* vp = a.Vs + ((long long)b * a.T + t) * D * D;
    for (int i = 0; i < D; ++i) m[i] = mp[i];
    for (int i = 0; i < D * D; ++i) Vm[i] = vp[i];
    const P* varb = a.has_var ? reinterpret_cast<const P*>(a.var.base) + (long long)b * a.var.seq_stride : nullptr;
    for (int c = 0; c < a.V; ++c) {
        P xy[2], J[2 * EKS_MAX_STATE];
        if (a.ncam > 0) {
            project_cam_jac<P, P>(a.cams + c * CAM_STRIDE, m, xy, J);  // J is 2 x 3
        } else {
            for (int r = 0; r < 2; ++r) {
                const P* Cr = a.C + ((long long)b * 2 * a.V + 2 * c + r) * D;
                P acc = P(0);
                for (int j = 0; j < D; ++j) { J[r * D + j] = Cr[j]; acc += Cr[j] * m[j]; }
                xy[r] = acc + (a.ymean ? a.ymean[(long long)b * 2 * a.V + 2 * c + r] : P(0));
            }
        }
        P* ob = a.out + (long long)b * a.out_seq_stride + (long long)c * a.out_cam_stride;
        for (int r = 0; r < 2; ++r) {
            P pv = P(0);
            for (int i = 0; i < D; ++i) {
                P acc = P(0);
                for (int j = 0; j < D; ++j) acc += Vm[i * D + j] * J[r * D + j];
                pv += J[r * D + i] * acc;
            }
            if (a.has_var) {
                // the nonlinear path of the reference adds columns 0 / 1 of the (T, 2V) variance array for
                // EVERY camera (multicam_smoother.py:459-460, :943-944); the linear path adds the camera's own
                const int ch = a.var_quirk ? r : 2 * c + r;
                pv += varb[a.var.chan_off[ch] + t];
            }
            ob[a.plane_off[r] + t] = xy[r];
            ob[a.plane_off[2 + r] + t] = pv;
        }
    }
}

template <class P>
int reproject_launch(const ReprojArgs<P>& a, cudaStream_t st) {
    dim3 grid(a.B, (a.T + 255) / 256);
    reproject_kernel<P><<<grid, 256, 0, st>>>(a);
    return check_launch("reproject_kernel");
}

}  // namespace eks

using namespace eks;

extern "C" int eks_reproject(int dtype, int B, int T, int D, int V, const void* ms, const void* Vs, const void* C,
                             const void* ymean, int ncam, const void* cams, const void* var_base,
                             long long var_seq_stride, const long long* var_chan_off, int pinhole_var_quirk,
                             void* out, long long out_seq_stride, long long out_cam_stride,
                             const long long* plane_off, void* stream) {
    EKS_REQUIRE(ms && Vs && out && plane_off, "reproject: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && D >= 1 && D <= EKS_MAX_STATE && V >= 1 && 2 * V <= EKS_MAX_CHAN,
                "reproject: bad dims");
    EKS_REQUIRE(ncam > 0 ? (cams != nullptr && D == 3 && ncam == V) : C != nullptr,
                "reproject: need C (linear) or cams with D == 3 (pinhole)");
#define EKS_FILL(PT)                                                                                     \
    ReprojArgs<PT> a;                                                                                    \
    a.B = B; a.T = T; a.D = D; a.V = V; a.ncam = ncam; a.var_quirk = pinhole_var_quirk;                  \
    a.ms = (const PT*)ms; a.Vs = (const PT*)Vs; a.C = (const PT*)C; a.ymean = (const PT*)ymean;          \
    a.cams = (const PT*)cams;                                                                            \
    a.has_var = (var_base != nullptr && var_chan_off != nullptr);                                        \
    a.var.base = var_base; a.var.seq_stride = var_seq_stride;                                            \
    for (int i = 0; i < MAX_CHAN; ++i) a.var.chan_off[i] = (a.has_var && i < 2 * V) ? var_chan_off[i] : 0; \
    a.out = (PT*)out; a.out_seq_stride = out_seq_stride; a.out_cam_stride = out_cam_stride;              \
    for (int i = 0; i < 4; ++i) a.plane_off[i] = plane_off[i];                                           \
    return reproject_launch<PT>(a, (cudaStream_t)stream);
    if (dtype == EKS_F32) { EKS_FILL(float) }
    EKS_FILL(double)
#undef EKS_FILL
}
