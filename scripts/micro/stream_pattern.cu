// Micro-benchmark: does the per-warp-run streaming pattern of diag_nll_kernel limit DRAM bandwidth?
// Every warp streams 4 KB warp-tiles through a 2-stage cp.async ring (same ring geometry as the kernel) and only
// sums one word per lane.  mode 0: each warp owns a contiguous run of tiles (the kernel's pattern);
// mode 1: the 8 warps of a CTA interleave tiles (adjacent addresses in flight at the same time).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NW = 8, PAD = 144, STAGE = 32 * PAD, STAGES = 2, WT = 1024;
__device__ inline void issue(unsigned char* stage, const float* src, int lane) {
    const int j0 = lane / 8, q = lane % 8;
    const float* s = src + j0 * 32 + q * 4;
    unsigned char* d = stage + j0 * PAD + q * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(d + i * 4 * PAD);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(s + i * 4 * 32) : "memory");
    }
}
template <int MODE>
__global__ void __launch_bounds__(256, 3) stream_kernel(const float* __restrict__ base, long long plane, int tiles_per_cta,
                                                        float* out) {
    extern __shared__ __align__(16) unsigned char ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* p = base + (long long)blockIdx.y * plane + (long long)blockIdx.x * tiles_per_cta * WT;
    const int per_warp = tiles_per_cta / NW;
    unsigned char* wr = ring + warp * STAGES * STAGE;
    auto tile_ptr = [&](int k) { return p + (long long)(MODE == 0 ? warp * per_warp + k : k * NW + warp) * WT; };
    float acc = 0;
    issue(wr, tile_ptr(0), lane);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int it = 0; it < per_warp; ++it) {
        __syncwarp();
        if (it + 1 < per_warp) issue(wr + ((it + 1) & 1) * STAGE, tile_ptr(it + 1), lane);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        const float4* mine = reinterpret_cast<const float4*>(wr + (it & 1) * STAGE + lane * PAD);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float4 v = mine[i]; acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    const int planes = 2560;              // 1280 sequences x 2 channels
    const long long T = 1000448;          // multiple of 1024 * 8
    float *d, *o;
    cudaMalloc(&d, planes * T * sizeof(float));
    cudaMalloc(&o, 4);
    cudaMemset(d, 0, planes * T * sizeof(float));
    for (int nseg : {1, 2, 4}) {
        const int tiles = (int)(T / WT);
        const int tiles_per_cta = tiles / nseg / NW * NW;
        const size_t smem = NW * STAGES * STAGE;
        for (int mode = 0; mode < 2; ++mode) {
            auto k = mode == 0 ? stream_kernel<0> : stream_kernel<1>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                k<<<dim3(nseg, planes), 256, smem>>>(d, T, tiles_per_cta, o);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)planes * nseg * tiles_per_cta * WT * 4;
            printf("nseg %d mode %d: %.3f ms  %.1f GB/s  (%s)\n", nseg, mode, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
