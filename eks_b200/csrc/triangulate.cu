// triangulate.cu -- triangulate_3d_models(...).mean(axis=0) of the calibrated multi-camera path on the device
// (reference eks/multicam_smoother.py:888-911 and :385-386): one thread per (keypoint, frame), looping over the
// ensemble members; fp64 throughout like the host code it replaces.
#include "triangulate.cuh"
#include "../../include/eks_b200.h"

namespace eks {

template <class Tin>
__global__ void __launch_bounds__(128) triangulate_mean_kernel(const Tin* __restrict__ raw, int M, int V, int T, int K,
                                                               const double* __restrict__ cams,
                                                               double* __restrict__ out /*[K][T][3]*/) {
    extern __shared__ double scam[];
    for (int i = threadIdx.x; i < V * CAM_STRIDE; i += blockDim.x) scam[i] = cams[i];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)K * T) return;
    const int k = (int)(idx / T), t = (int)(idx - (long long)k * T);
    double acc[3] = {0, 0, 0};
    for (int m = 0; m < M; ++m) {
        double uv[16];
        for (int c = 0; c < V; ++c) {
            const Tin* p = raw + ((((long long)m * V + c) * T + t) * K + k) * 3;
            uv[2 * c] = (double)p[0];
            uv[2 * c + 1] = (double)p[1];
        }
        double X[3];
        triangulate_point(scam, V, uv, X);
        acc[0] += X[0]; acc[1] += X[1]; acc[2] += X[2];
    }
    double* o = out + ((long long)k * T + t) * 3;
    o[0] = acc[0] / M; o[1] = acc[1] / M; o[2] = acc[2] / M;
}

}  // namespace eks

using namespace eks;

extern "C" int eks_triangulate_mean(const void* raw, int raw_dtype, int M, int V, int T, int K, const double* cams,
                                    double* out, void* stream) {
    EKS_REQUIRE(raw && cams && out, "triangulate_mean: null pointer");
    EKS_REQUIRE(M >= 1 && T >= 1 && K >= 1 && V >= 2 && V <= 8, "triangulate_mean: needs 2..8 cameras");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)K * T;
    const int blocks = (int)((n + 127) / 128);
    const size_t smem = (size_t)V * CAM_STRIDE * sizeof(double);
    if (raw_dtype == EKS_F32)
        triangulate_mean_kernel<float><<<blocks, 128, smem, st>>>((const float*)raw, M, V, T, K, cams, out);
    else
        triangulate_mean_kernel<double><<<blocks, 128, smem, st>>>((const double*)raw, M, V, T, K, cams, out);
    return check_launch("triangulate_mean_kernel");
}
