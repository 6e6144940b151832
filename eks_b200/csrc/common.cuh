// common.cuh -- shared device helpers for the eks_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#ifndef __CUDA_ARCH__
using std::isfinite;
using std::isinf;
using std::isnan;
#endif

namespace eks {

constexpr int MAX_CHAN = 16;   // max observation channels (2 * cameras)
constexpr int CAM_STRIDE = 29; // R(9) t(3) fx fy cx cy skew k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4

// Channel-plane view of per-frame data: element (sequence b, channel o, frame t) lives at
//   base[b * seq_stride + chan_off[o] + t]      (frame-major planes, t contiguous)
struct PlaneView {
    const void* base;
    long long seq_stride;
    long long chan_off[MAX_CHAN];
};

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void note_launches(int n);   // kernels enqueued by the current entry point (eks_last_launch_count)
void note_unverified(int n); // evaluations accepted with an unverified run boundary (eks_last_unverified_count)

#define EKS_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            eks::set_error(__VA_ARGS__);  \
            return -1;                    \
        }                                 \
    } while (0)

// ------------------------------------------------------------------ forward-mode dual number
template <class T>
struct Dual {
    T v, d;
    __host__ __device__ Dual() : v(0), d(0) {}
    __host__ __device__ Dual(T v_) : v(v_), d(0) {}
    __host__ __device__ Dual(T v_, T d_) : v(v_), d(d_) {}
};
template <class T> __host__ __device__ inline Dual<T> operator+(Dual<T> a, Dual<T> b) { return {a.v + b.v, a.d + b.d}; }
template <class T> __host__ __device__ inline Dual<T> operator-(Dual<T> a, Dual<T> b) { return {a.v - b.v, a.d - b.d}; }
template <class T> __host__ __device__ inline Dual<T> operator-(Dual<T> a) { return {-a.v, -a.d}; }
template <class T> __host__ __device__ inline Dual<T> operator*(Dual<T> a, Dual<T> b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <class T> __host__ __device__ inline Dual<T> operator/(Dual<T> a, Dual<T> b) {
    T q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
template <class T> __host__ __device__ inline Dual<T>& operator+=(Dual<T>& a, Dual<T> b) { a = a + b; return a; }
template <class T> __host__ __device__ inline Dual<T>& operator-=(Dual<T>& a, Dual<T> b) { a = a - b; return a; }

__host__ __device__ inline float sqrt_(float x) { return sqrtf(x); }
__host__ __device__ inline double sqrt_(double x) { return sqrt(x); }
__host__ __device__ inline float log_(float x) { return logf(x); }
__host__ __device__ inline double log_(double x) { return log(x); }
template <class T> __host__ __device__ inline Dual<T> sqrt_(Dual<T> a) { T r = sqrt_(a.v); return {r, a.d / (r + r)}; }
template <class T> __host__ __device__ inline Dual<T> log_(Dual<T> a) { return {log_(a.v), a.d / a.v}; }

template <class S> struct Scalar { using real = S; __host__ __device__ static real val(S x) { return x; } };
template <class T> struct Scalar<Dual<T>> {
    using real = typename Scalar<T>::real;
    __host__ __device__ static real val(Dual<T> x) { return Scalar<T>::val(x.v); }
};
template <class S> __host__ __device__ inline typename Scalar<S>::real val(S x) { return Scalar<S>::val(x); }

// ------------------------------------------------------------------ warp / block reductions
template <class T>
__device__ inline T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree order).  `scratch` needs >= 32 elements of T.
template <class T>
__device__ inline T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T r = (threadIdx.x < nwarp) ? scratch[threadIdx.x] : T(0);
    if (warp == 0) r = warp_sum(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    r = scratch[0];
    return r;
}

template <class T> __device__ inline T ldg_as(const float* p) { return T(__ldg(p)); }
template <class T> __device__ inline T ldg_as(const double* p) { return T(__ldg(p)); }

}  // namespace eks
