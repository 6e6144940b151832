// ensemble.cu -- per-frame ensemble statistics across seed predictions (reference:
// eks/core.py:25-101, inner compute_stats :58-85).
//
// Input : raw[S][M][V][T][K][3]  (reference MarkerArray layout, fields innermost; f32 or f64)
// Output: five frame-major planes per (session, camera, keypoint):
//           out[s*sess_stride + v*cam_stride + k*kp_stride + plane_off[f] + t],  f in
//           {0:x_avg, 1:y_avg, 2:var_x, 3:var_y, 4:mean likelihood}
//         plus optional per-tile partial moments (sum x, sum y, sum x^2, sum y^2 of the averaged
//         coordinates, fp64) so that centring / S0 need no extra pass over HBM.
//
// One CTA = one (session, camera, tile of TT frames) for all K keypoints.  Loads walk the AoS
// input with keypoint fastest (fully coalesced), results are transposed through shared memory and
// written with frame fastest (fully coalesced).  HBM-bound: 12*M bytes in, 20 bytes out per cell.
#include <cstdlib>
#include "common.cuh"
#include "sort_networks.cuh"
#include "../../include/eks_b200.h"

namespace eks {

struct EnsOut {
    long long sess_stride, cam_stride, kp_stride;
    long long plane_off[5];
};

template <class P> __device__ inline P pos_inf();
template <> __device__ inline float pos_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ inline double pos_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <class P> __device__ inline P real_max();
template <> __device__ inline float real_max<float>() { return 3.402823466e+38f; }
template <> __device__ inline double real_max<double>() { return 1.7976931348623157e+308; }

// bitonic sorting network on N (power of two) register values, ascending
template <class P, int N>
__device__ inline void sort_network(P (&v)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = ((i & k) == 0);
                    const P a = v[i], b = v[l];
                    const bool sw = up ? (a > b) : (a < b);
                    v[i] = sw ? b : a;
                    v[l] = sw ? a : b;
                }
            }
        }
    }
}

template <class P, int N>
__device__ __forceinline__ void sort_values(P (&v)[N]) {
    if constexpr (N <= 16) SortNet<N>::run(v);
    else sort_network<P, N>(v);
}

// statistics of one coordinate over the M seeds (NaN-aware), reference core.py:64-79.
// EXACT: the array length MAXM is the seed count (all `m < M` predicates fold away).
template <class P, int MAXM, bool EXACT = false>
__device__ inline void coord_stats(const P (&x)[MAXM], int M_rt, bool avg_median, P& avg, P& var) {
    const int M = EXACT ? MAXM : M_rt;
    int n = 0;
    P sum = P(0);
#pragma unroll
    for (int m = 0; m < MAXM; ++m)
        if (m < M && !isnan(x[m])) { sum += x[m]; ++n; }
    if (n == 0) {
        avg = nan("");
        var = nan("");
        return;
    }
    const P mean = sum / P(n);
    P ss = P(0);
#pragma unroll
    for (int m = 0; m < MAXM; ++m)
        if (m < M && !isnan(x[m])) { const P d = x[m] - mean; ss += d * d; }
    var = ss / P(n);
    if (!avg_median) {
        avg = mean;
        return;
    }
    P s[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) s[m] = (m < M && !isnan(x[m])) ? x[m] : pos_inf<P>();
    sort_values<P, MAXM>(s);
    // nanmedian: middle element, or mean of the two middle elements, of the n valid values
    P a, b;
    if (EXACT && n == MAXM) {  // no NaN among the seeds (the common case): fixed positions
        a = s[(MAXM - 1) / 2];
        b = s[MAXM / 2];
    } else {
        const int hi = n >> 1, lo = (n & 1) ? hi : hi - 1;
        a = P(0); b = P(0);
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m == lo) a = s[m];
            if (m == hi) b = s[m];
        }
    }
    avg = (a == b) ? a : a * P(0.5) + b * P(0.5);
}

template <class Tin, class P, int MAXM>
__global__ void __launch_bounds__(256) ensemble_kernel(const Tin* __restrict__ raw, long long raw_sess_stride, int M,
                                                       int V, int T, int K, int avg_median, int var_mode,
                                                       P nan_repl, P* __restrict__ out, EnsOut eo,
                                                       double* __restrict__ partials, int TT) {
    extern __shared__ unsigned char smem_raw[];
    P* tile = reinterpret_cast<P*>(smem_raw);  // [5][K][TT+1]
    const int tile_idx = blockIdx.x, v = blockIdx.y, sess = blockIdx.z;
    const int t0 = tile_idx * TT;
    const int nt = min(TT, T - t0);
    const int ld = TT + 1;
    const Tin* base = raw + (long long)sess * raw_sess_stride;
    const long long m_stride = (long long)V * T * K * 3;
    const long long off0 = ((long long)v * T + t0) * K * 3;

    for (int e = threadIdx.x; e < nt * K; e += blockDim.x) {
        const int tl = e / K, k = e - tl * K;
        P xs[MAXM], ys[MAXM];
        P conf = P(0);
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < M) {
                const Tin* p = base + (long long)m * m_stride + off0 + (long long)e * 3;
                xs[m] = P(__ldg(p));       // cast to the compute precision first (core.py:90-92)
                ys[m] = P(__ldg(p + 1));
                conf += P(__ldg(p + 2));   // likelihood sum is NOT NaN-aware (core.py:67-68)
            } else {
                xs[m] = P(0);
                ys[m] = P(0);
            }
        }
        const P mean_conf = conf / P(M);
        P ax, vx, ay, vy;
        coord_stats<P, MAXM>(xs, M, avg_median != 0, ax, vx);
        coord_stats<P, MAXM>(ys, M, avg_median != 0, ay, vy);
        if (M == 1) {
            vx = vy = (mean_conf != mean_conf) ? mean_conf : P(1) / fmax(mean_conf, P(1e-5));   // NaN propagates (jnp.maximum)
        } else if (var_mode == 1) {
            vx = vx / mean_conf;
            vy = vy / mean_conf;
        }
        // jnp.nan_to_num(nan=nan_replacement): nan -> repl, +-inf -> +-max
        if (isnan(vx)) vx = nan_repl; else if (isinf(vx)) vx = vx > 0 ? real_max<P>() : -real_max<P>();
        if (isnan(vy)) vy = nan_repl; else if (isinf(vy)) vy = vy > 0 ? real_max<P>() : -real_max<P>();
        tile[(0 * K + k) * ld + tl] = ax;
        tile[(1 * K + k) * ld + tl] = ay;
        tile[(2 * K + k) * ld + tl] = vx;
        tile[(3 * K + k) * ld + tl] = vy;
        tile[(4 * K + k) * ld + tl] = mean_conf;
    }
    __syncthreads();
    // coalesced plane writes: frame fastest
    P* obase = out + (long long)sess * eo.sess_stride + (long long)v * eo.cam_stride;
    for (int idx = threadIdx.x; idx < 5 * K * nt; idx += blockDim.x) {
        const int tl = idx % nt, fk = idx / nt;
        const int f = fk / K, k = fk - f * K;
        obase[(long long)k * eo.kp_stride + eo.plane_off[f] + t0 + tl] = tile[(f * K + k) * ld + tl];
    }
    // per-tile partial moments of the averaged coordinates (deterministic: fixed lane tree)
    if (partials != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
        const int ntiles = gridDim.x;
        for (int k = warp; k < K; k += nwarp) {
            double sx = 0, sy = 0, sxx = 0, syy = 0;
            for (int tl = lane; tl < nt; tl += 32) {
                const double x = (double)tile[(0 * K + k) * ld + tl];
                const double y = (double)tile[(1 * K + k) * ld + tl];
                sx += x; sy += y; sxx += x * x; syy += y * y;
            }
            sx = warp_sum(sx); sy = warp_sum(sy); sxx = warp_sum(sxx); syy = warp_sum(syy);
            if (lane == 0) {
                double* pp = partials + ((((long long)sess * V + v) * K + k) * ntiles + tile_idx) * 4;
                pp[0] = sx; pp[1] = sy; pp[2] = sxx; pp[3] = syy;
            }
        }
    }
}

// ---- staged variant (default): the tile's raw values are brought into shared memory with fully
// coalesced 16-byte cp.async requests (one contiguous chunk per seed), then every (cell, coordinate) work
// item reads its M values from shared memory.  Compared with direct strided loads this cuts the L1
// wavefronts per request from 12 sectors to 4 and halves the register footprint.
__device__ inline void ens_cp_async_16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
template <int BYTES>
__device__ inline void ens_cp_async_small(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(sa), "l"(gmem), "n"(BYTES) : "memory");
}

// NaN-free statistics of one coordinate (all M seeds valid): mean, ddof-0 variance in seed order, median
// from the fixed middle positions of the exact-size sorting network.
template <class P, int M>
__device__ __forceinline__ void coord_stats_clean(const P (&x)[M], bool avg_median, P& avg, P& var) {
    P sum = P(0);
#pragma unroll
    for (int m = 0; m < M; ++m) sum += x[m];
    const P mean = sum / P(M);
    P ss = P(0);
#pragma unroll
    for (int m = 0; m < M; ++m) { const P d = x[m] - mean; ss += d * d; }
    var = ss / P(M);
    if (!avg_median) { avg = mean; return; }
    P s[M];
#pragma unroll
    for (int m = 0; m < M; ++m) s[m] = x[m];
    sort_values<P, M>(s);
    const P a = s[(M - 1) / 2], b = s[M / 2];
    avg = (M & 1) ? b : a * P(0.5) + b * P(0.5);
}

template <class Tin, class P, int MAXM, bool EXACT>
__global__ void __launch_bounds__(256) ensemble_staged_kernel(const Tin* __restrict__ raw, long long raw_sess_stride,
                                                              int M_rt, int V, int T, int K, int avg_median,
                                                              int var_mode, P nan_repl, P* __restrict__ out, EnsOut eo,
                                                              double* __restrict__ partials, int TT, int log2TT,
                                                              unsigned invK) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = EXACT ? MAXM : M_rt;
    const int tile_idx = blockIdx.x, v = blockIdx.y, sess = blockIdx.z;
    const int t0 = tile_idx * TT;
    const int nt = min(TT, T - t0);
    const int ld = TT + 1;
    const int chunk = TT * K * 3;                     // elements per seed in a full tile
    const int chunk_pad = (chunk * (int)sizeof(Tin) + 15) / 16 * 16;  // bytes, keeps every seed 16-B aligned
    P* tile = reinterpret_cast<P*>(smem_raw + (size_t)M * chunk_pad);  // [5][K][TT+1]
    const Tin* base = raw + (long long)sess * raw_sess_stride;
    const long long m_stride = (long long)V * T * K * 3;
    const long long off0 = ((long long)v * T + t0) * K * 3;
    const int nel = nt * K * 3;
    constexpr int EPG = 16 / (int)sizeof(Tin);
    const bool vec = ((reinterpret_cast<uintptr_t>(base + off0) & 15) == 0) && ((m_stride * (int)sizeof(Tin)) % 16 == 0);
    const int ngran = vec ? nel / EPG : 0;
    if (EXACT && vec && ngran * EPG == nel) {
        // whole 16-byte granules (every full tile): each seed's chunk is ONE contiguous run in global memory, so the M
        // chunks are brought in by M bulk asynchronous copies (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) issued by a
        // single thread and completed on an mbarrier.  The per-granule cp.async loop this replaces was 13 % of the
        // kernel's instructions (ncu r2: issue slots 83 % busy, the kernel was issue bound, not HBM bound).
        __shared__ __align__(8) unsigned long long mbar;
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mb) : "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)nel * (unsigned)sizeof(Tin);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(bytes * (unsigned)MAXM)
                         : "memory");
            const Tin* src = base + off0;
            unsigned dst = (unsigned)__cvta_generic_to_shared(smem_raw);
#pragma unroll 1
            for (int m = 0; m < MAXM; ++m, src += m_stride, dst += chunk_pad)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                             "l"(src), "r"(bytes), "r"(mb)
                             : "memory");
        }
        {   // every thread waits for the phase (parity 0) to complete: the copied bytes are then visible to it
            unsigned done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done)
                             : "r"(mb)
                             : "memory");
        }
    } else {   // stage the M contiguous seed chunks: coalesced 16-byte cp.async, pointers advanced by constant strides
        const Tin* src = base + off0;
        unsigned char* dst = smem_raw;
        for (int m = 0; m < M; ++m, src += m_stride, dst += chunk_pad) {
            for (int g = threadIdx.x; g < ngran; g += blockDim.x) ens_cp_async_16(dst + g * 16, src + g * EPG);
            for (int e = ngran * EPG + threadIdx.x; e < nel; e += blockDim.x)
                ens_cp_async_small<(int)sizeof(Tin)>(dst + e * sizeof(Tin), src + e);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
    }
    // one thread per cell (frame, keypoint): 3*M shared-memory reads at a constant stride
    const int ncell = nt * K;
    for (int e = threadIdx.x; e < ncell; e += blockDim.x) {
        const int tl = (K == 1) ? e : (int)__umulhi((unsigned)e, invK);  // e / K without a divide
        const int k = e - tl * K;
        P xs[MAXM], ys[MAXM];
        P conf = P(0), sumx = P(0), sumy = P(0);
        const unsigned char* sp = smem_raw + (size_t)e * 3 * sizeof(Tin);
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < M) {
                const Tin* q = reinterpret_cast<const Tin*>(sp + (size_t)m * chunk_pad);
                xs[m] = P(q[0]);                     // cast to the compute precision first (core.py:90-92)
                ys[m] = P(q[1]);
                conf += P(q[2]);                     // likelihood sum is NOT NaN-aware (core.py:67-68)
                sumx += xs[m];
                sumy += ys[m];
            } else {
                xs[m] = P(0);
                ys[m] = P(0);
            }
        }
        const P mean_conf = conf / P(M);
        P ax, vx, ay, vy;
        // a NaN among the seeds makes the plain sum NaN: one test per coordinate instead of one per value.  (inf - inf
        // also lands on the NaN-aware path, which treats +-inf exactly like the plain path does.)
        if (EXACT && !(sumx != sumx) && !(sumy != sumy)) {
            coord_stats_clean<P, MAXM>(xs, avg_median != 0, ax, vx);
            coord_stats_clean<P, MAXM>(ys, avg_median != 0, ay, vy);
        } else {
            coord_stats<P, MAXM, EXACT>(xs, M, avg_median != 0, ax, vx);
            coord_stats<P, MAXM, EXACT>(ys, M, avg_median != 0, ay, vy);
        }
        if (M == 1) {
            // jnp.maximum propagates NaN (a NaN likelihood -> NaN variance -> nan_replacement); fmax would drop it
            vx = vy = (mean_conf != mean_conf) ? mean_conf : P(1) / fmax(mean_conf, P(1e-5));
        } else if (var_mode == 1) {
            vx = vx / mean_conf;
            vy = vy / mean_conf;
        }
        // jnp.nan_to_num(nan=nan_replacement): nan -> repl, +-inf -> +-max  (branch free: clamp, then select)
        {
            const P cx = fmin(fmax(vx, -real_max<P>()), real_max<P>()), cy = fmin(fmax(vy, -real_max<P>()), real_max<P>());
            vx = (vx != vx) ? nan_repl : cx;
            vy = (vy != vy) ? nan_repl : cy;
        }
        P* tp = tile + k * ld + tl;
        const int fs = K * ld;
        tp[0] = ax;
        tp[fs] = ay;
        tp[2 * fs] = vx;
        tp[3 * fs] = vy;
        tp[4 * fs] = mean_conf;
    }
    __syncthreads();
    // coalesced plane writes: frame fastest (TT is a power of two: no integer division).  A thread owns fixed
    // (keypoint, frame) slots and computes their addresses ONCE for the five planes (the per-plane loops with a 64-bit
    // multiply per element were 21 % of the kernel's instructions).
    P* obase = out + (long long)sess * eo.sess_stride + (long long)v * eo.cam_stride + t0;
    const int fstride = K * ld;
    for (int idx = threadIdx.x; idx < (K << log2TT); idx += blockDim.x) {
        const int k = idx >> log2TT, tl = idx & (TT - 1);
        if (tl < nt) {
            P* o = obase + (long long)k * eo.kp_stride + tl;
            const P* tsrc = tile + k * ld + tl;
#pragma unroll
            for (int f = 0; f < 5; ++f) o[eo.plane_off[f]] = tsrc[f * fstride];
        }
    }
    if (partials != nullptr) {
        // per-tile moments of the averaged coordinates: row = (coordinate, keypoint), LPR lanes per row with
        // 4 frames each, fixed xor tree inside the lane group (deterministic)
        const int log2LPR = log2TT - 2;                // TT in {8,16,32,64} -> 2..16 lanes per row (powers of two:
        const int LPR = 1 << log2LPR;                  // shifts instead of the integer divisions, 12 % of the instructions)
        const int log2RPP = 8 - log2LPR;               // blockDim.x == 256
        const int rows_per_pass = 1 << log2RPP;
        const int q = threadIdx.x & (LPR - 1);
        const int ntiles = gridDim.x;
        const int npass = (2 * K + rows_per_pass - 1) >> log2RPP;
        for (int ps = 0; ps < npass; ++ps) {
            const int row = (ps << log2RPP) + (threadIdx.x >> log2LPR);
            const bool rv = row < 2 * K;
            const int c = (rv && row >= K) ? 1 : 0, k = rv ? row - c * K : 0;
            double sm = 0, sq = 0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int tl = q * 4 + jj;
                if (rv && tl < nt) {
                    const double x = (double)tile[(c * K + k) * ld + tl];
                    sm += x;
                    sq += x * x;
                }
            }
            for (int off = LPR >> 1; off > 0; off >>= 1) {
                sm += __shfl_xor_sync(0xffffffffu, sm, off);
                sq += __shfl_xor_sync(0xffffffffu, sq, off);
            }
            if (rv && q == 0) {
                double* pp = partials + ((((long long)sess * V + v) * K + k) * ntiles + tile_idx) * 4;
                pp[c] = sm;
                pp[2 + c] = sq;
            }
        }
    }
}

// reduce the per-tile partials in a fixed order: one CTA per (session, camera, keypoint)
template <class P>
__global__ void __launch_bounds__(1024) moments_finalize_kernel(const double* __restrict__ partials, int nseq,
                                                               int ntiles, long long T, P* __restrict__ mean_out,
                                                               P* __restrict__ var_out) {
    __shared__ double scratch[32];
    const int seq = blockIdx.x;
    const double* pp = partials + (long long)seq * ntiles * 4;
    double a[4] = {0, 0, 0, 0};
    {   // four independent loads in flight per thread (a rolled loop with one outstanding load per thread made this
        // 2 MB reduction per sequence latency bound: 0.15 ms); fixed order: slot u of every batch, then u = 0..3
        constexpr int UN = 4;
        double acc[UN][4];
#pragma unroll
        for (int u = 0; u < UN; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0;
        for (int i0 = threadIdx.x; i0 < ntiles; i0 += UN * blockDim.x) {
            double2 lo[UN], hi[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * blockDim.x;
                lo[u] = hi[u] = make_double2(0.0, 0.0);
                if (i < ntiles) {
                    lo[u] = *reinterpret_cast<const double2*>(pp + (long long)i * 4);
                    hi[u] = *reinterpret_cast<const double2*>(pp + (long long)i * 4 + 2);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) { acc[u][0] += lo[u].x; acc[u][1] += lo[u].y; acc[u][2] += hi[u].x; acc[u][3] += hi[u].y; }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] += acc[u][q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = block_sum(a[q], scratch);
    if (threadIdx.x == 0) {
        const double n = (double)T;
        const double mx = a[0] / n, my = a[1] / n;
        // the centred data are formed in precision P (x - P(mean)); its variance about its own mean
        // equals the variance of x, so S0 = E[x^2] - mean^2 in fp64 (utils.py:350-351,
        // singlecam_smoother.py:262-266)
        mean_out[seq * 2 + 0] = P(mx);
        mean_out[seq * 2 + 1] = P(my);
        // np.nanvar of an all-NaN column (a NaN mean makes every centred value NaN) is NaN: keep it, do not clamp to 0
        const double vx = a[2] / n - mx * mx, vy = a[3] / n - my * my;
        var_out[seq * 2 + 0] = P(vx != vx ? vx : fmax(vx, 0.0));
        var_out[seq * 2 + 1] = P(vy != vy ? vy : fmax(vy, 0.0));
    }
}

// frames per tile: the largest of {64,32,16,8} whose staged tile (M seed chunks + output tile) fits the
// shared-memory budget that still allows ~4 CTAs per SM; 0 -> use the direct kernel (64 frames per tile)
static int staged_tile_frames(int M, int K, int in_bytes, int out_bytes) {
    const int cand[4] = {64, 32, 16, 8};
    for (int i = 0; i < 4; ++i) {
        const int TT = cand[i];
        const size_t need = (size_t)M * (((size_t)TT * K * 3 * in_bytes + 15) / 16 * 16) +
                            (size_t)5 * K * (TT + 1) * out_bytes;
        if (need <= 56 * 1024) return TT;
    }
    return 0;
}

template <class Tin, class P>
int launch_ensemble(const Tin* raw, long long raw_sess_stride, int S, int M, int V, int T, int K, int avg_median,
                    int var_mode, double nan_repl, P* out, const EnsOut& eo, double* partials, cudaStream_t st) {
    EKS_REQUIRE(M <= 32, "ensemble: at most 32 seeds supported, got %d", M);
    const int TTs = staged_tile_frames(M, K, (int)sizeof(Tin), (int)sizeof(P));
    const bool staged = TTs > 0;
    const int TT = staged ? TTs : 64;
    const int ntiles = (T + TT - 1) / TT;
    // (a CTA sized to the tile's cell count -- 320 threads for 16 frames x 20 keypoints -- was measured slower:
    // 6.1 ms vs 5.3 ms for the stage on the c5 bench; other CTAs on the SM already fill the uneven second round)
    // threads per CTA: EKS_ENS_THREADS overrides (experiments); 256 by default
    static const int nt_env = [] { const char* e = getenv("EKS_ENS_THREADS"); return e ? atoi(e) : 256; }();
    dim3 grid(ntiles, V, S), block((nt_env >= 64 && nt_env <= 256 && nt_env % 32 == 0) ? nt_env : 256);
    size_t smem = (size_t)5 * K * (TT + 1) * sizeof(P);
    if (staged) smem += (size_t)M * (((size_t)TT * K * 3 * sizeof(Tin) + 15) / 16 * 16);
    EKS_REQUIRE(smem <= 200 * 1024, "ensemble: K=%d too large for the shared-memory tile", K);
    int log2TT = 0;
    while ((1 << log2TT) < TT) ++log2TT;
    // e / K == umulhi(e, ceil(2^32 / K)) for e * K < 2^32 (cells per tile are far below that)
    const unsigned invK = (unsigned)((((unsigned long long)1 << 32) + (unsigned long long)K - 1) / (unsigned long long)K);
#define EKS_ENS_STAGED(MAXM, EXACT)                                                                           \
    do {                                                                                                      \
        auto kern = ensemble_staged_kernel<Tin, P, MAXM, EXACT>;                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<grid, block, smem, st>>>(raw, raw_sess_stride, M, V, T, K, avg_median, var_mode, (P)nan_repl, out, \
                                        eo, partials, TT, log2TT, invK);                                      \
    } while (0)
#define EKS_ENS_DIRECT(MAXM)                                                                                  \
    do {                                                                                                      \
        auto kern = ensemble_kernel<Tin, P, MAXM>;                                                            \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<grid, block, smem, st>>>(raw, raw_sess_stride, M, V, T, K, avg_median, var_mode, (P)nan_repl, out, \
                                        eo, partials, TT);                                                    \
    } while (0)
    if (staged) {
        switch (M) {  // exact-size networks for the usual ensemble sizes
            case 1: EKS_ENS_STAGED(1, true); break;
            case 2: EKS_ENS_STAGED(2, true); break;
            case 3: EKS_ENS_STAGED(3, true); break;
            case 4: EKS_ENS_STAGED(4, true); break;
            case 5: EKS_ENS_STAGED(5, true); break;
            case 6: EKS_ENS_STAGED(6, true); break;
            case 7: EKS_ENS_STAGED(7, true); break;
            case 8: EKS_ENS_STAGED(8, true); break;
            case 9: EKS_ENS_STAGED(9, true); break;
            case 10: EKS_ENS_STAGED(10, true); break;
            default:
                if (M <= 16) EKS_ENS_STAGED(16, false);
                else EKS_ENS_STAGED(32, false);
        }
    } else {
        if (M <= 16) EKS_ENS_DIRECT(16);
        else EKS_ENS_DIRECT(32);
    }
#undef EKS_ENS_STAGED
#undef EKS_ENS_DIRECT
    return check_launch("ensemble_kernel");
}

}  // namespace eks

using namespace eks;

extern "C" int eks_ensemble_tile_frames(int M, int K, int raw_dtype, int out_dtype) {
    const int TT = staged_tile_frames(M, K, raw_dtype == EKS_F32 ? 4 : 8, out_dtype == EKS_F32 ? 4 : 8);
    return TT > 0 ? TT : 64;
}

extern "C" int eks_ensemble_stats(const void* raw, int raw_dtype, long long raw_sess_stride, int n_sessions, int M,
                                  int V, int T, int K, int avg_median, int var_mode, double nan_replacement,
                                  void* out, int out_dtype, long long sess_stride, long long cam_stride,
                                  long long kp_stride, const long long* plane_off, double* moment_partials,
                                  void* stream) {
    EKS_REQUIRE(raw && out && plane_off, "ensemble: null pointer");
    EKS_REQUIRE(M >= 1 && V >= 1 && T >= 1 && K >= 1 && n_sessions >= 1, "ensemble: bad dims");
    EKS_REQUIRE(!(raw_dtype == EKS_F32 && out_dtype == EKS_F64), "ensemble: f32 input with f64 output unsupported");
    EnsOut eo;
    eo.sess_stride = sess_stride; eo.cam_stride = cam_stride; eo.kp_stride = kp_stride;
    for (int i = 0; i < 5; ++i) eo.plane_off[i] = plane_off[i];
    cudaStream_t st = (cudaStream_t)stream;
    if (raw_dtype == EKS_F32 && out_dtype == EKS_F32)
        return launch_ensemble<float, float>((const float*)raw, raw_sess_stride, n_sessions, M, V, T, K, avg_median,
                                             var_mode, nan_replacement, (float*)out, eo, moment_partials, st);
    if (raw_dtype == EKS_F64 && out_dtype == EKS_F32)
        return launch_ensemble<double, float>((const double*)raw, raw_sess_stride, n_sessions, M, V, T, K,
                                              avg_median, var_mode, nan_replacement, (float*)out, eo,
                                              moment_partials, st);
    if (raw_dtype == EKS_F64 && out_dtype == EKS_F64)
        return launch_ensemble<double, double>((const double*)raw, raw_sess_stride, n_sessions, M, V, T, K,
                                               avg_median, var_mode, nan_replacement, (double*)out, eo,
                                               moment_partials, st);
    EKS_REQUIRE(false, "ensemble: unsupported dtype combination");
}

extern "C" int eks_center_moments(const double* moment_partials, int n_seq, int n_tiles, int T, void* mean_out,
                                  void* var_out, int dtype, void* stream) {
    EKS_REQUIRE(moment_partials && mean_out && var_out, "center_moments: null pointer");
    const int ntiles = n_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        moments_finalize_kernel<float><<<n_seq, 1024, 0, st>>>(moment_partials, n_seq, ntiles, T, (float*)mean_out,
                                                              (float*)var_out);
    else
        moments_finalize_kernel<double><<<n_seq, 1024, 0, st>>>(moment_partials, n_seq, ntiles, T, (double*)mean_out,
                                                               (double*)var_out);
    return check_launch("moments_finalize_kernel");
}
