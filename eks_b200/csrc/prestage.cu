// prestage.cu -- small per-sequence reductions that feed the s-optimiser:
//   * initial smoothing-parameter guess  (reference eks/core.py:104-133, caller :233-236, :612-613)
//   * constant observation noise for the loss path: nanmedian over (cropped) time of the
//     ensemble variances with a floor (reference eks/core.py:702-709; crop eks/utils.py:235-290)
// The median is an exact radix select on order-preserving integer keys (no sort, no sampling).
#include "common.cuh"
#include "../../include/eks_b200.h"

namespace eks {

constexpr int MAX_SPANS = 16;
struct Spans {
    int n;
    int start[MAX_SPANS];
    int cum[MAX_SPANS + 1];  // cum[i] = number of cropped frames before span i
};

__device__ inline int span_to_frame(const Spans& sp, int i) {
    int j = 0;
#pragma unroll 1
    while (j + 1 < sp.n && i >= sp.cum[j + 1]) ++j;
    return sp.start[j] + (i - sp.cum[j]);
}

// ------------------------------------------------------------------ initial guess
template <class P>
__global__ void __launch_bounds__(256) guess_kernel(PlaneView var, int B, int O, int T, double* __restrict__ guess,
                                                    P* __restrict__ s_log0) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const int n = min(T, 2000);
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride;
    // nanstd (ddof 0) of all frame-to-frame differences of the first <=2000 frames, all channels
    double sum = 0, cnt = 0;
    for (int i = threadIdx.x; i < (n - 1) * O; i += blockDim.x) {
        const int o = i / (n - 1), t = i - o * (n - 1);
        const P* p = base + var.chan_off[o] + t;
        const P d = p[1] - p[0];  // difference formed in the storage precision, as numpy does
        if (!isnan(d)) { sum += (double)d; cnt += 1.0; }
    }
    sum = block_sum(sum, scratch);
    cnt = block_sum(cnt, scratch);
    const double mean = sum / cnt;
    double ss = 0;
    for (int i = threadIdx.x; i < (n - 1) * O; i += blockDim.x) {
        const int o = i / (n - 1), t = i - o * (n - 1);
        const P* p = base + var.chan_off[o] + t;
        const P d = p[1] - p[0];
        if (!isnan(d)) { const double e = (double)d - mean; ss += e * e; }
    }
    ss = block_sum(ss, scratch);
    if (threadIdx.x == 0) {
        double g = sqrt(ss / cnt);
        if (sizeof(P) == 4) g = (double)(float)g;  // numpy returns float32 for float32 input
        g = rint(g * 1e5) / 1e5;                   // round(., 5)
        if (sizeof(P) == 4) g = (double)(float)g;
        if (!(g > 0.0) || !isfinite(g)) g = 2.0;   // `or 2.0` and the non-finite / <=0 fallback
        guess[b] = g;
        if (s_log0) {
            const double s0 = fmin(fmax(g, 1e-6), 1e3);
            s_log0[b] = P((float)log(s0));         // float32 seed (core.py:622)
        }
    }
}

// ------------------------------------------------------------------ radix select (nanmedian)
template <class P> struct KeyT;
template <> struct KeyT<float> {
    using type = unsigned int;
    static constexpr int nlevels = 3;
    __device__ static type key(float x) {
        if (isnan(x)) return 0xFFFFFFFFu;
        unsigned int u = __float_as_uint(x);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    __device__ static float unkey(type k) {
        unsigned int u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
        return __uint_as_float(u);
    }
    __host__ __device__ static int shift(int level) { return level == 0 ? 21 : (level == 1 ? 10 : 0); }
    __host__ __device__ static int bits(int level) { return level == 2 ? 10 : 11; }
};
template <> struct KeyT<double> {
    using type = unsigned long long;
    static constexpr int nlevels = 6;
    __device__ static type key(double x) {
        if (isnan(x)) return 0xFFFFFFFFFFFFFFFFull;
        unsigned long long u = (unsigned long long)__double_as_longlong(x);
        return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    }
    __device__ static double unkey(type k) {
        unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
        return __longlong_as_double((long long)u);
    }
    __host__ __device__ static int shift(int level) {
        const int s[6] = {53, 42, 31, 20, 10, 0};
        return s[level];
    }
    __host__ __device__ static int bits(int level) { return level >= 4 ? 10 : 11; }
};

struct SelState {           // per problem (sequence, channel)
    unsigned long long prefix[2];
    int rank[2];
    int n_valid;
    int pad;
};

constexpr int NBINS = 2048;
constexpr int SEL_CHUNK = 16384;

// extra workspace of the one-pass median (states, fail flags, candidate keys): see run_const_R
static size_t med_extra_bytes(int nprob, int n_total, size_t key_bytes, size_t state_bytes) {
    const size_t cap = (size_t)n_total / 8 + 4096;
    return 256 + (((size_t)nprob * state_bytes + 255) & ~(size_t)255) + (((size_t)nprob * sizeof(int) + 255) & ~(size_t)255) +
           (size_t)nprob * cap * key_bytes + 256;
}

template <class P>
__global__ void __launch_bounds__(256) select_hist_kernel(PlaneView var, int O, Spans sp, int n_total, int level,
                                                          const SelState* __restrict__ state,
                                                          int* __restrict__ hist /*[prob][2][NBINS]*/,
                                                          const int* __restrict__ only = nullptr) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ int sh[2][NBINS];
    const int prob = blockIdx.x, b = prob / O, o = prob - b * O;   // problems on grid x: no 65535 limit
    if (only && !only[prob]) return;
    for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride + var.chan_off[o];
    const int shift = KT::shift(level), nb = KT::bits(level);
    const int hshift = shift + nb;  // bits above this are already determined
    key_t pre0 = 0, pre1 = 0;
    if (level > 0) { pre0 = (key_t)state[prob].prefix[0]; pre1 = (key_t)state[prob].prefix[1]; }
    const bool same = (level == 0) || (pre0 == pre1);
    const int i0 = blockIdx.y * SEL_CHUNK, i1 = min(n_total, i0 + SEL_CHUNK);
    // (warp-aggregating the atomics with match.any was measured slower -- 3.3 ms vs 2.3 ms for this stage on the
    // c5 bench -- the bins hit by one warp are too many for the aggregation loop to pay off)
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int t = (sp.n == 1) ? sp.start[0] + i : span_to_frame(sp, i);
        const P x = base[t];
        if (isnan(x)) continue;
        const key_t k = KT::key(x);
        const int bin = (int)((k >> shift) & (key_t)((1 << nb) - 1));
        const key_t top = (level == 0) ? 0 : (k >> hshift);
        if (level == 0 || top == pre0) atomicAdd(&sh[0][bin], 1);
        if (!same && top == pre1) atomicAdd(&sh[1][bin], 1);
    }
    __syncthreads();
    int* gh = hist + (long long)prob * 2 * NBINS;
    for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) {
        const int c = (&sh[0][0])[i];
        if (c) atomicAdd(gh + i, c);
    }
}

// numpy's virtual index of the 'linear' percentile method, evaluated in the array's precision exactly as
// numpy does for floating-point input and a Python-float q: (n - 1) * (q / 100), both operations in P
// (numpy/lib/_function_base_impl.py: percentile -> true_divide(q, dtype(100)); _QuantileMethods['linear'])
template <class P>
__device__ inline P virtual_index(int n, double q) {
    return (P)(n - 1) * ((P)q / P(100));
}

template <class P>
__global__ void __launch_bounds__(256) select_scan_kernel(int level, SelState* __restrict__ state,
                                                          int* __restrict__ hist, double floor_lo, double floor_hi,
                                                          double q /* percent; < 0: nanmedian */,
                                                          P* __restrict__ out, const int* __restrict__ only = nullptr) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ int sh[2][NBINS];
    const int prob = blockIdx.x;
    if (only && !only[prob]) return;
    int* gh = hist + (long long)prob * 2 * NBINS;
    for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) {
        (&sh[0][0])[i] = gh[i];
        gh[i] = 0;  // ready for the next level
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        SelState st = state[prob];
        const int nb = KT::bits(level);
        bool same = true;
        if (level == 0) {
            int n = 0;
            for (int i = 0; i < NBINS; ++i) n += sh[0][i];
            st.n_valid = n;
            if (q < 0) {
                st.rank[0] = (n - 1) / 2;
                st.rank[1] = n / 2;
            } else {  // np.percentile, method 'linear': virtual index (n - 1) q / 100 between two order statistics
                const P pos = virtual_index<P>(n, q);
                const int lo = (int)floor((double)pos);
                if (pos >= (P)(n - 1)) st.rank[0] = st.rank[1] = n - 1;       // _get_indexes: above bounds
                else if (pos < P(0)) st.rank[0] = st.rank[1] = 0;
                else { st.rank[0] = lo; st.rank[1] = lo + 1; }
            }
            st.prefix[0] = st.prefix[1] = 0;
        } else {
            same = (st.prefix[0] == st.prefix[1]);
        }
        if (st.n_valid > 0) {
            for (int r = 0; r < 2; ++r) {
                const int* h = sh[(same ? 0 : r)];
                int cum = 0, bin = 0;
                for (bin = 0; bin < (1 << nb); ++bin) {
                    if (cum + h[bin] > st.rank[r]) break;
                    cum += h[bin];
                }
                st.rank[r] -= cum;
                st.prefix[r] = (st.prefix[r] << nb) | (unsigned long long)bin;
            }
        }
        state[prob] = st;
        if (level == KT::nlevels - 1) {
            P med;
            if (st.n_valid == 0) med = P(nan(""));
            else {
                const P a = KT::unkey((key_t)st.prefix[0]), c = KT::unkey((key_t)st.prefix[1]);
                if (q < 0) {
                    med = (a + c) * P(0.5);
                    // floors: build_R_from_vars clip (1e-12) then min_R_var (np.clip keeps NaN)
                    med = (P)fmax((double)med, floor_lo);
                    med = (P)fmax((double)med, floor_hi);
                } else {  // numpy _lerp(a, c, gamma) in the array's precision
                    const P pos = virtual_index<P>(st.n_valid, q);
                    const P g = pos - (P)floor((double)pos);
                    const P d = c - a;
                    med = (g >= P(0.5)) ? c - d * (P(1) - g) : a + d * g;
                }
            }
            out[prob] = med;
        }
    }
}

// ------------------------------------------------------------------ nanmedian in ONE pass over the data
// The three radix passes above read every variance plane three times.  For long sequences the median is bracketed
// first: med_sample_kernel sorts MED_NS jittered samples of a problem and takes the sample quantiles 0.5 -+ MED_DELTA
// (5 sigma of the sample-quantile rank error) as a key bracket [lo, hi]; med_count_kernel then streams the plane ONCE,
// counting the NaNs and the keys below the bracket and compacting the ~8 % of the keys inside it; med_final_kernel
// runs the exact radix select on those candidates with the ranks shifted by the count below.  The result is the same
// order statistic(s) as the three-pass select -- bit exact -- whenever the true median ranks fall inside the bracket
// and the candidate buffers did not overflow; otherwise the problem is flagged and the three-pass select redoes it
// (its CTAs return immediately for unflagged problems).
constexpr int MED_NS = 4096;
constexpr float MED_DELTA = 0.04f;
constexpr int MED_CHUNK = 16384;          // frames per CTA of med_count_kernel
constexpr int MED_SCAP = 4096;            // candidates a CTA can hold (25 % of its chunk)
constexpr int MED_MIN_FRAMES = 1 << 17;   // below this the three-pass select is used directly

template <class P>
struct MedState {
    typename KeyT<P>::type lo, hi;
    int n_below, n_nan, n_cand, fail;
};

template <class P>
__global__ void __launch_bounds__(512) med_sample_kernel(PlaneView var, int O, Spans sp, int n_total,
                                                         MedState<P>* __restrict__ state) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ key_t keys[MED_NS];
    __shared__ int nvalid_sh;
    const int prob = blockIdx.x, b = prob / O, o = prob - b * O;
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride + var.chan_off[o];
    if (threadIdx.x == 0) nvalid_sh = 0;
    __syncthreads();
    const int stride = n_total / MED_NS;          // >= 32 (MED_MIN_FRAMES)
    int nv = 0;
    for (int j = threadIdx.x; j < MED_NS; j += blockDim.x) {
        const unsigned h = (unsigned)j * 2654435761u;
        const int i = j * stride + (int)((h >> 8) % (unsigned)stride);      // jittered position inside cell j
        const int t = (sp.n == 1) ? sp.start[0] + i : span_to_frame(sp, i);
        const P x = base[t];
        keys[j] = KT::key(x);                     // NaN -> all ones: sorted to the end
        nv += isnan(x) ? 0 : 1;
    }
    nv = warp_sum(nv);
    if ((threadIdx.x & 31) == 0) atomicAdd(&nvalid_sh, nv);
    __syncthreads();
    for (int k = 2; k <= MED_NS; k <<= 1)          // bitonic sort, ascending
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = threadIdx.x; i < MED_NS; i += blockDim.x) {
                const int l = i ^ jj;
                if (l > i) {
                    const key_t a = keys[i], c = keys[l];
                    const bool up = ((i & k) == 0);
                    if ((a > c) == up) { keys[i] = c; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        MedState<P> st;
        const int n = nvalid_sh;
        st.n_below = 0; st.n_nan = 0; st.n_cand = 0;
        st.fail = (n < 256) ? 1 : 0;               // mostly NaN: no usable bracket
        const int il = max(0, (int)floorf((float)n * (0.5f - MED_DELTA)) - 1);
        const int ih = min(max(n - 1, 0), (int)ceilf((float)n * (0.5f + MED_DELTA)) + 1);
        st.lo = keys[il];
        st.hi = keys[ih];
        state[prob] = st;
    }
}

template <class P>
__global__ void __launch_bounds__(256) med_count_kernel(PlaneView var, int O, Spans sp, int n_total, int cap,
                                                        MedState<P>* __restrict__ state,
                                                        typename KeyT<P>::type* __restrict__ cand) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    // candidates a thread can hold: its 64 frames contain Poisson(~5.3) of them.  16 overflowed in ~2e-5 of the threads,
    // i.e. in ~25 % of the 10^6-frame problems (15 600 threads each), which then took the three-pass fallback (ncu launch
    // list of round 2: select_hist / select_scan doing real work); with 24 the overflow probability per problem is < 1e-5
    constexpr int LCAP = 24;
    __shared__ key_t buf[MED_SCAP];
    __shared__ int wsum[8];
    __shared__ int sh_base, sh_fail;
    const int prob = blockIdx.x, b = prob / O, o = prob - b * O;
    const key_t lo = state[prob].lo, hi = state[prob].hi;
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride + var.chan_off[o];
    if (threadIdx.x == 0) sh_fail = state[prob].fail;
    __syncthreads();
    if (sh_fail) return;      // (read once: other CTAs of this problem may set the flag concurrently)
    const int i0 = blockIdx.y * MED_CHUNK, i1 = min(n_total, i0 + MED_CHUNK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int below = 0, nans = 0, nloc = 0;
    key_t loc[LCAP];
    constexpr int UN = 4;       // independent loads in flight per thread
    const bool one_span = (sp.n == 1);
    const P* src = base + (one_span ? sp.start[0] : 0);
    const key_t NANKEY = ~(key_t)0;
    for (int ib = i0 + threadIdx.x; ib < i1; ib += UN * blockDim.x) {
        P x[UN];
        if (one_span && ib + (UN - 1) * (int)blockDim.x < i1) {
            // full batch of one contiguous span (all but the last batch of a chunk): no bounds checks, no span lookup,
            // branch-free counting -- the checked path below cost ~41 instructions per element (ncu r2f: ISETP + BRA +
            // IMAD + BSSY/BSYNC 60 % of the kernel), which kept this pass at half the HBM rate
#pragma unroll
            for (int u = 0; u < UN; ++u) x[u] = src[ib + u * blockDim.x];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const key_t k = KT::key(x[u]);                 // NaN -> all ones
                const bool isn = (k == NANKEY);
                nans += isn ? 1 : 0;
                below += (k < lo) ? 1 : 0;
                if (k >= lo && k <= hi && !isn) {
                    if (nloc < LCAP) loc[nloc] = k;
                    ++nloc;
                }
            }
            continue;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = ib + u * blockDim.x;
            x[u] = P(0);
            if (i < i1) x[u] = base[one_span ? sp.start[0] + i : span_to_frame(sp, i)];
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = ib + u * blockDim.x;
            if (i < i1) {
                if (isnan(x[u])) ++nans;
                else {
                    const key_t k = KT::key(x[u]);
                    if (k < lo) ++below;
                    else if (k <= hi) {
                        if (nloc < LCAP) loc[nloc] = k;
                        ++nloc;
                    }
                }
            }
        }
    }
    // block-wide exclusive scan of the per-thread candidate counts (no atomics in the streaming loop)
    const int over = __syncthreads_or(nloc > LCAP ? 1 : 0);
    int incl = nloc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    below = warp_sum(below);
    nans = warp_sum(nans);
    __syncthreads();
    int woff = 0, n = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) woff += wsum[w]; n += wsum[w]; }
    const int excl = woff + incl - nloc;
    if (lane == 0) {
        if (below) atomicAdd(&state[prob].n_below, below);
        if (nans) atomicAdd(&state[prob].n_nan, nans);
    }
    if (threadIdx.x == 0) {
        int gbase = 0;
        if (over || n > MED_SCAP) { state[prob].fail = 1; gbase = cap; }   // a thread's or the CTA's buffer overflowed
        else if (n > 0) gbase = atomicAdd(&state[prob].n_cand, n);
        sh_base = gbase;
    }
    if (!over && n <= MED_SCAP)
        for (int j = 0; j < nloc; ++j) buf[excl + j] = loc[j];
    __syncthreads();
    const int gbase = sh_base;
    if (gbase + n > cap) {                                                  // the problem's buffer overflowed
        if (threadIdx.x == 0) state[prob].fail = 1;
        return;
    }
    key_t* dst = cand + (long long)prob * cap + gbase;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
}

template <class P>
__global__ void __launch_bounds__(1024) med_final_kernel(int n_total, int cap, MedState<P>* __restrict__ state,
                                                         const typename KeyT<P>::type* __restrict__ cand,
                                                         double floor_lo, double floor_hi, P* __restrict__ out,
                                                         int* __restrict__ fail_out) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ int hist[2][NBINS];
    __shared__ key_t prefix[2];
    __shared__ int rank[2];
    __shared__ int ok;
    const int prob = blockIdx.x;
    const MedState<P> st = state[prob];
    const int n_valid = n_total - st.n_nan;
    if (threadIdx.x == 0) {
        rank[0] = (n_valid - 1) / 2 - st.n_below;
        rank[1] = n_valid / 2 - st.n_below;
        prefix[0] = prefix[1] = 0;
        ok = !st.fail && n_valid > 0 && rank[0] >= 0 && rank[1] < st.n_cand && st.n_cand <= cap;
        fail_out[prob] = ok ? 0 : 1;
    }
    __syncthreads();
    if (!ok) return;
    const key_t* kc = cand + (long long)prob * cap;
    const int nc = st.n_cand;
    // Radix select on the keys RELATIVE to the bracket, k - lo in [0, hi - lo], most significant digit first.  (On the raw
    // keys every candidate shares the leading digits: ~10^5 shared-memory atomics on one or two bins, measured 190 us.)
    const key_t span = st.hi - st.lo;
    int nbits = 0;
    while (nbits < (int)(8 * sizeof(key_t)) && (span >> nbits) != 0) ++nbits;      // bit length of span
    for (int top = nbits; top > 0; top -= 11) {
        const int nb = min(11, top), shift = top - nb;
        for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) (&hist[0][0])[i] = 0;
        __syncthreads();
        const key_t pre0 = prefix[0], pre1 = prefix[1];
        const bool first = (top == nbits);
        const bool same = first || (pre0 == pre1);
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const key_t k = kc[i] - st.lo;
            const int bin = (int)((k >> shift) & (key_t)((1 << nb) - 1));
            const key_t hi_part = first ? 0 : (k >> top);
            if (first || hi_part == pre0) atomicAdd(&hist[0][bin], 1);
            if (!same && hi_part == pre1) atomicAdd(&hist[1][bin], 1);
        }
        __syncthreads();
        if (threadIdx.x < 64) {     // warp r finds the bin holding rank[r]: 64 bins per lane, warp prefix sum, local scan
            const int r = threadIdx.x >> 5, ln = threadIdx.x & 31;
            const int* h = hist[same ? 0 : r];
            const int per = NBINS / 32;
            int mine = 0;
            for (int q = 0; q < per; ++q) mine += h[ln * per + q];
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (ln >= d) incl += v;
            }
            const int target = rank[r];
            const unsigned hit = __ballot_sync(0xffffffffu, incl > target);     // first lane whose cumulative count passes
            const int owner = hit ? __ffs(hit) - 1 : 31;
            __syncwarp();
            if (ln == owner) {
                int cum = incl - mine, bin = ln * per;
                for (; bin < ln * per + per - 1; ++bin) {
                    if (cum + h[bin] > target) break;
                    cum += h[bin];
                }
                rank[r] = target - cum;
                prefix[r] = (prefix[r] << nb) | (key_t)bin;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const P a = KT::unkey(st.lo + prefix[0]), c = KT::unkey(st.lo + prefix[1]);
        P med = (a + c) * P(0.5);
        med = (P)fmax((double)med, floor_lo);      // build_R_from_vars clip (1e-12) then min_R_var, as select_scan_kernel
        med = (P)fmax((double)med, floor_hi);
        out[prob] = med;
    }
}

template <class P>
int run_const_R(const PlaneView& var, int B, int O, const Spans& sp, int n_total, double min_var, P* out,
                void* workspace, size_t workspace_bytes, cudaStream_t st, double q = -1.0, double floor_lo = 1e-12) {
    const int nprob = B * O;
    const size_t need = (size_t)nprob * (sizeof(SelState) + 2 * NBINS * sizeof(int));
    EKS_REQUIRE(workspace && workspace_bytes >= need, "const_R_median: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    SelState* state = reinterpret_cast<SelState*>(workspace);
    int* hist = reinterpret_cast<int*>(state + nprob);
    cudaMemsetAsync(workspace, 0, need, st);
    const int nchunks = (n_total + SEL_CHUNK - 1) / SEL_CHUNK;
    const int* only = nullptr;
    int launches = 0;
    if (q < 0 && n_total >= MED_MIN_FRAMES) {      // nanmedian of a long sequence: bracket + ONE pass, three-pass fallback
        using key_t = typename KeyT<P>::type;
        const int cap = n_total / 8 + MED_SCAP;
        const size_t extra = med_extra_bytes(nprob, n_total, sizeof(key_t), sizeof(MedState<P>));
        EKS_REQUIRE(workspace_bytes >= need + extra, "const_R_median: workspace too small (%zu < %zu)", workspace_bytes,
                    need + extra);
        unsigned char* w = (unsigned char*)workspace + ((need + 255) & ~(size_t)255);
        MedState<P>* ms = (MedState<P>*)w; w += ((size_t)nprob * sizeof(MedState<P>) + 255) & ~(size_t)255;
        int* fail = (int*)w; w += ((size_t)nprob * sizeof(int) + 255) & ~(size_t)255;
        key_t* cand = (key_t*)w;
        med_sample_kernel<P><<<nprob, 512, 0, st>>>(var, O, sp, n_total, ms);
        med_count_kernel<P><<<dim3(nprob, (n_total + MED_CHUNK - 1) / MED_CHUNK), 256, 0, st>>>(var, O, sp, n_total, cap, ms,
                                                                                                 cand);
        med_final_kernel<P><<<nprob, 1024, 0, st>>>(n_total, cap, ms, cand, floor_lo, min_var, out, fail);
        const int rc = check_launch("median bracket kernels");
        if (rc) return rc;
        only = fail;
        launches = 3;
    }
    for (int level = 0; level < KeyT<P>::nlevels; ++level) {
        select_hist_kernel<P><<<dim3(nprob, nchunks), 256, 0, st>>>(var, O, sp, n_total, level, state, hist, only);
        select_scan_kernel<P><<<nprob, 256, 0, st>>>(level, state, hist, floor_lo, min_var, q, out, only);
    }
    note_launches(launches + 2 * KeyT<P>::nlevels);
    return check_launch("select kernels");
}


// ====================================================================================================
// Multi-camera pre-stage: centring (eks/utils.py:293-365), the data passes of the per-keypoint
// PCA fit (eks/stats.py:9-64 -> sklearn PCA: mean + covariance of the variance-filtered frames; the tiny
// eigen-decomposition stays on the host) and the latent-space initialisation (eks/multicam_smoother.py:554-597).
// A "problem" b = (session, keypoint) owns O = 2 * cameras channel planes.  All sums are fp64, reduced in a fixed
// order (per-chunk partials, then one warp per problem), so results are run-to-run deterministic.
// ====================================================================================================
constexpr int MC_CHUNK = 1024;
constexpr int MC_NT = 256;
constexpr int MC_SUB = MC_CHUNK / MC_NT;

struct McWork {
    void* maxvar;      // [B][T]   max over channels of the ensemble variance
    void* thr;         // [B]      percentile threshold
    int* cnt;          // [B][nchunk] good frames per chunk
    int* last;         // [B][nchunk] last good frame of the chunk (-1: none)
    int* pre;          // [B][nchunk] good frames before the chunk
    int* lastbefore;   // [B][nchunk] last good frame before the chunk (-1: none)
    int* ngood;        // [B]
    int* minf;         // [B] min over the session's keypoints of ngood
    double* part;      // [B][nchunk][MC_PART_MAX]
    void* sel;         // radix-select state + histograms for B problems
    int nchunk;
};
constexpr int MC_PART_MAX = MAX_CHAN + MAX_CHAN * (MAX_CHAN + 1) / 2;   // 152 >= 2 + 3 L + L (L + 1) / 2

static size_t mc_align(size_t x) { return (x + 255) / 256 * 256; }

static size_t mc_carve(int dtype, int B, int T, void* workspace, McWork* w) {
    const size_t real = dtype == EKS_F32 ? 4 : 8;
    const int nchunk = (T + MC_CHUNK - 1) / MC_CHUNK;
    unsigned char* p = (unsigned char*)workspace;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* q = p ? p + off : nullptr; off += mc_align(bytes); return q; };
    void* maxvar = take((size_t)B * T * real);
    void* thr = take((size_t)B * real);
    int* cnt = (int*)take((size_t)B * nchunk * 4);
    int* last = (int*)take((size_t)B * nchunk * 4);
    int* pre = (int*)take((size_t)B * nchunk * 4);
    int* lastbefore = (int*)take((size_t)B * nchunk * 4);
    int* ngood = (int*)take((size_t)B * 4);
    int* minf = (int*)take((size_t)B * 4);
    double* part = (double*)take((size_t)B * nchunk * MC_PART_MAX * 8);
    void* sel = take((size_t)B * (sizeof(SelState) + 2 * NBINS * sizeof(int)));
    if (w) *w = McWork{maxvar, thr, cnt, last, pre, lastbefore, ngood, minf, part, sel, nchunk};
    return off;
}

__device__ inline int block_sum_int(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = 0;
    for (int i = 0; i < nwarp; ++i) r += scratch[i];
    return r;
}
__device__ inline int block_max_int(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = __reduce_max_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = scratch[0];
    for (int i = 1; i < nwarp; ++i) r = max(r, scratch[i]);
    return r;
}

template <class P>
__global__ void __launch_bounds__(MC_NT) mc_maxvar_kernel(PlaneView var, int O, int T, P* __restrict__ mv) {
    const int b = blockIdx.x;
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride;
    const int t1 = min(T, (int)(blockIdx.y + 1) * MC_CHUNK);
    for (int t = blockIdx.y * MC_CHUNK + threadIdx.x; t < t1; t += MC_NT) {
        P m = base[var.chan_off[0] + t];
        for (int o = 1; o < O; ++o) {
            const P v = base[var.chan_off[o] + t];
            m = (v > m || isnan(v)) ? v : m;      // np.max: NaN propagates
        }
        mv[(long long)b * T + t] = m;
    }
}

template <class P>
__global__ void __launch_bounds__(MC_NT) mc_count_kernel(const P* __restrict__ mv, const P* __restrict__ thr, int T,
                                                        int nchunk, int* __restrict__ cnt, int* __restrict__ last) {
    __shared__ int scratch[32];
    const int b = blockIdx.x, c = blockIdx.y;
    const P th = thr[b];
    const int t1 = min(T, (c + 1) * MC_CHUNK);
    int n = 0, l = -1;
    for (int t = c * MC_CHUNK + threadIdx.x; t < t1; t += MC_NT)
        if (mv[(long long)b * T + t] <= th) { ++n; l = t; }
    n = block_sum_int(n, scratch);
    l = block_max_int(l, scratch);
    if (threadIdx.x == 0) { cnt[(long long)b * nchunk + c] = n; last[(long long)b * nchunk + c] = l; }
}

// one CTA per session: chunk prefixes per keypoint, then the session's minimum good-frame count
__global__ void __launch_bounds__(MC_NT) mc_scan_kernel(int K, int nchunk, const int* __restrict__ cnt,
                                                       const int* __restrict__ last, int* __restrict__ pre,
                                                       int* __restrict__ lastbefore, int* __restrict__ ngood,
                                                       int* __restrict__ minf) {
    __shared__ int ssum[MC_NT], smax[MC_NT];
    __shared__ int carry_sum, carry_max, s_min;
    const int s = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_min = 0x7fffffff;
    for (int k = 0; k < K; ++k) {
        const long long b = (long long)s * K + k;
        if (tid == 0) { carry_sum = 0; carry_max = -1; }
        __syncthreads();
        for (int base = 0; base < nchunk; base += MC_NT) {
            const int i = base + tid;
            const int v = i < nchunk ? cnt[b * nchunk + i] : 0;
            const int m = i < nchunk ? last[b * nchunk + i] : -1;
            ssum[tid] = v; smax[tid] = m;
            __syncthreads();
            for (int d = 1; d < MC_NT; d <<= 1) {          // inclusive Hillis-Steele scans
                const int a = tid >= d ? ssum[tid - d] : 0;
                const int x = tid >= d ? smax[tid - d] : -1;
                __syncthreads();
                ssum[tid] += a; smax[tid] = max(smax[tid], x);
                __syncthreads();
            }
            if (i < nchunk) {
                pre[b * nchunk + i] = carry_sum + ssum[tid] - v;
                lastbefore[b * nchunk + i] = max(carry_max, tid > 0 ? smax[tid - 1] : -1);
            }
            __syncthreads();
            if (tid == MC_NT - 1) { carry_sum += ssum[tid]; carry_max = max(carry_max, smax[tid]); }
            __syncthreads();
        }
        if (tid == 0) { ngood[b] = carry_sum; s_min = min(s_min, carry_sum); }
        __syncthreads();
    }
    if (tid < K) minf[(long long)s * K + tid] = s_min;
    for (int k = MC_NT; k < K; k += MC_NT) if (k + tid < K) minf[(long long)s * K + k + tid] = s_min;
}

// in-chunk bookkeeping shared by the three data passes: good flag, rank among the good frames, previous good frame
struct McSel {
    int run_cnt, run_last;   // carried across the sub-tiles of the chunk (uniform)
};
template <class P>
__device__ inline void mc_flags(const P* __restrict__ mv_b, P th, int t, int T, McSel& st, int* wcnt, int* wlast,
                                bool& good, int& rank, int& prev) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    good = (t < T) && (mv_b[t] <= th);
    const unsigned bal = __ballot_sync(0xffffffffu, good);
    const unsigned lower = bal & ((1u << lane) - 1u);
    __syncthreads();                                      // previous sub-tile's readers are done
    if (lane == 0) { wcnt[warp] = __popc(bal); wlast[warp] = bal ? (t - lane) + (31 - __clz(bal)) : -1; }
    __syncthreads();
    int off = st.run_cnt, pl = st.run_last;
    for (int w = 0; w < warp; ++w) { off += wcnt[w]; pl = max(pl, wlast[w]); }
    rank = off + __popc(lower);
    prev = lower ? (t - lane) + (31 - __clz(lower)) : pl;
    int tot = 0, tl = -1;
    for (int w = 0; w < MC_NT / 32; ++w) { tot += wcnt[w]; tl = max(tl, wlast[w]); }
    st.run_cnt += tot;
    st.run_last = max(st.run_last, tl);
}

// pass 1: sum of the selected frames (first minf good frames) per channel
template <class P>
__global__ void __launch_bounds__(MC_NT) mc_mean_kernel(PlaneView y, int O, int T, McWork w) {
    __shared__ int wcnt[MC_NT / 32], wlast[MC_NT / 32];
    __shared__ double scratch[32];
    const int b = blockIdx.x, c = blockIdx.y;
    const P* yb = reinterpret_cast<const P*>(y.base) + (long long)b * y.seq_stride;
    const P* mv_b = reinterpret_cast<const P*>(w.maxvar) + (long long)b * T;
    const P th = reinterpret_cast<const P*>(w.thr)[b];
    const int minf = w.minf[b];
    McSel st{w.pre[(long long)b * w.nchunk + c], w.lastbefore[(long long)b * w.nchunk + c]};
    double acc[MAX_CHAN];
    for (int o = 0; o < MAX_CHAN; ++o) acc[o] = 0;
    for (int j = 0; j < MC_SUB; ++j) {
        const int t = c * MC_CHUNK + j * MC_NT + threadIdx.x;
        bool good; int rank, prev;
        mc_flags<P>(mv_b, th, t, T, st, wcnt, wlast, good, rank, prev);
        if (good && rank < minf)
            for (int o = 0; o < O; ++o) acc[o] += (double)yb[y.chan_off[o] + t];
    }
    double* out = w.part + ((long long)b * w.nchunk + c) * MC_PART_MAX;
    for (int o = 0; o < O; ++o) {
        const double v = block_sum(acc[o], scratch);
        if (threadIdx.x == 0) out[o] = v;
    }
}

// one warp per problem: fixed-order sum of the per-chunk partials
__device__ inline double mc_part_sum(const McWork& w, int b, int i) {
    double v = 0;
    for (int c = threadIdx.x & 31; c < w.nchunk; c += 32) v += w.part[((long long)b * w.nchunk + c) * MC_PART_MAX + i];
    return warp_sum(v);
}

template <class P>
__global__ void __launch_bounds__(32) mc_mean_final_kernel(int O, McWork w, P* __restrict__ ymean,
                                                           int* __restrict__ n_good_out) {
    const int b = blockIdx.x;
    const int n = w.minf[b];
    for (int o = 0; o < O; ++o) {
        const double v = mc_part_sum(w, b, o);
        if (threadIdx.x == 0) ymean[(long long)b * O + o] = (P)(v / (double)n);
    }
    if (threadIdx.x == 0 && n_good_out) { n_good_out[2 * b] = w.ngood[b]; n_good_out[2 * b + 1] = n; }
}

// pass 2: first and second moments of the centred selected frames (the PCA fit's data pass)
template <class P, int OC>
__global__ void __launch_bounds__(MC_NT) mc_cov_kernel(PlaneView y, int O, int T, const P* __restrict__ ymean, McWork w) {
    constexpr int NACC = OC + OC * (OC + 1) / 2;
    __shared__ int wcnt[MC_NT / 32], wlast[MC_NT / 32];
    __shared__ double scratch[32];
    const int b = blockIdx.x, c = blockIdx.y;
    const P* yb = reinterpret_cast<const P*>(y.base) + (long long)b * y.seq_stride;
    const P* mv_b = reinterpret_cast<const P*>(w.maxvar) + (long long)b * T;
    const P th = reinterpret_cast<const P*>(w.thr)[b];
    const int minf = w.minf[b];
    McSel st{w.pre[(long long)b * w.nchunk + c], w.lastbefore[(long long)b * w.nchunk + c]};
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    for (int j = 0; j < MC_SUB; ++j) {
        const int t = c * MC_CHUNK + j * MC_NT + threadIdx.x;
        bool good; int rank, prev;
        mc_flags<P>(mv_b, th, t, T, st, wcnt, wlast, good, rank, prev);
        if (good && rank < minf) {
            double x[OC];
#pragma unroll
            for (int o = 0; o < OC; ++o) x[o] = o < O ? (double)(yb[y.chan_off[o] + t] - ymean[(long long)b * O + o]) : 0.0;
            int q = OC;
#pragma unroll
            for (int i = 0; i < OC; ++i) {
                acc[i] += x[i];
#pragma unroll
                for (int k = i; k < OC; ++k) acc[q++] += x[i] * x[k];
            }
        }
    }
    double* out = w.part + ((long long)b * w.nchunk + c) * MC_PART_MAX;
#pragma unroll 1
    for (int i = 0; i < NACC; ++i) {
        const double v = block_sum(acc[i], scratch);
        if (threadIdx.x == 0) out[i] = v;
    }
}

// moments_out [B][1 + O + O*O]: n, sum x_o, sum x_i x_j (full symmetric matrix)
template <int OC>
__global__ void __launch_bounds__(32) mc_cov_final_kernel(int O, McWork w, double* __restrict__ moments_out) {
    const int b = blockIdx.x;
    double* out = moments_out + (long long)b * (1 + O + O * O);
    if (threadIdx.x == 0) out[0] = (double)w.minf[b];
    for (int i = 0; i < O; ++i) {
        const double v = mc_part_sum(w, b, i);
        if (threadIdx.x == 0) out[1 + i] = v;
    }
    int q = OC;
    for (int i = 0; i < OC; ++i)
        for (int k = i; k < OC; ++k, ++q) {
            if (i >= O || k >= O) continue;
            const double v = mc_part_sum(w, b, q);
            if (threadIdx.x == 0) { out[1 + O + i * O + k] = v; out[1 + O + k * O + i] = v; }
        }
}

// pass 3: latent coordinates of ALL good frames: their variance and the covariance of consecutive differences
template <class P, int LC>
__global__ void __launch_bounds__(MC_NT) mc_latent_kernel(PlaneView y, int O, int L, int T, const P* __restrict__ ymean,
                                                         const P* __restrict__ pca_mean, const P* __restrict__ comps,
                                                         McWork w) {
    constexpr int NACC = 2 + 3 * LC + LC * (LC + 1) / 2;
    __shared__ int wcnt[MC_NT / 32], wlast[MC_NT / 32];
    __shared__ double scratch[32];
    __shared__ P sC[MAX_CHAN * LC], sMu[MAX_CHAN], sPm[MAX_CHAN];
    const int b = blockIdx.x, c = blockIdx.y;
    for (int i = threadIdx.x; i < O * LC; i += MC_NT) {
        const int o = i / LC, l = i - o * LC;
        sC[i] = l < L ? comps[((long long)b * O + o) * L + l] : P(0);
    }
    for (int i = threadIdx.x; i < O; i += MC_NT) { sMu[i] = ymean[(long long)b * O + i]; sPm[i] = pca_mean[(long long)b * O + i]; }
    __syncthreads();
    const P* yb = reinterpret_cast<const P*>(y.base) + (long long)b * y.seq_stride;
    const P* mv_b = reinterpret_cast<const P*>(w.maxvar) + (long long)b * T;
    const P th = reinterpret_cast<const P*>(w.thr)[b];
    McSel st{w.pre[(long long)b * w.nchunk + c], w.lastbefore[(long long)b * w.nchunk + c]};
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    auto latent = [&](int t, P* z) {   // pca.transform of the centred prediction: ((y - mean) - pca.mean_) @ components^T
#pragma unroll
        for (int l = 0; l < LC; ++l) z[l] = P(0);
        for (int o = 0; o < O; ++o) {
            const P x = (yb[y.chan_off[o] + t] - sMu[o]) - sPm[o];
#pragma unroll
            for (int l = 0; l < LC; ++l) z[l] += x * sC[o * LC + l];
        }
    };
    for (int j = 0; j < MC_SUB; ++j) {
        const int t = c * MC_CHUNK + j * MC_NT + threadIdx.x;
        bool good; int rank, prev;
        mc_flags<P>(mv_b, th, t, T, st, wcnt, wlast, good, rank, prev);
        if (good) {
            P z[LC];
            latent(t, z);
            acc[0] += 1.0;
#pragma unroll
            for (int l = 0; l < LC; ++l) { acc[2 + l] += (double)z[l]; acc[2 + LC + l] += (double)z[l] * (double)z[l]; }
            if (prev >= 0) {
                P zp[LC];
                latent(prev, zp);
                double d[LC];
#pragma unroll
                for (int l = 0; l < LC; ++l) d[l] = (double)(z[l] - zp[l]);
                acc[1] += 1.0;
                int q = 2 + 3 * LC;
#pragma unroll
                for (int i = 0; i < LC; ++i) {
                    acc[2 + 2 * LC + i] += d[i];
#pragma unroll
                    for (int k = i; k < LC; ++k) acc[q++] += d[i] * d[k];
                }
            }
        }
    }
    double* out = w.part + ((long long)b * w.nchunk + c) * MC_PART_MAX;
#pragma unroll 1
    for (int i = 0; i < NACC; ++i) {
        const double v = block_sum(acc[i], scratch);
        if (threadIdx.x == 0) out[i] = v;
    }
}

// S0 = diag(np.var(good_pcs)), Q = np.cov(diff(good_pcs).T) / max|.|  (eks/multicam_smoother.py:574-588)
template <class P, int LC>
__global__ void __launch_bounds__(32) mc_latent_final_kernel(int L, McWork w, P* __restrict__ S0, P* __restrict__ Q) {
    constexpr int NACC = 2 + 3 * LC + LC * (LC + 1) / 2;
    const int b = blockIdx.x;
    double a[NACC];
    for (int i = 0; i < NACC; ++i) a[i] = mc_part_sum(w, b, i);
    if (threadIdx.x != 0) return;
    const double n = a[0], nd = a[1];
    P* S = S0 + (long long)b * L * L;
    P* Qb = Q + (long long)b * L * L;
    for (int i = 0; i < L * L; ++i) { S[i] = P(0); Qb[i] = P(0); }
    for (int l = 0; l < L; ++l) {
        const double mu = a[2 + l] / n;
        S[l * L + l] = (P)(a[2 + LC + l] / n - mu * mu);
    }
    double cov[LC * LC], mx = 0;
    int q = 2 + 3 * LC;
    for (int i = 0; i < LC; ++i)
        for (int k = i; k < LC; ++k, ++q) {
            const double v = (a[q] - a[2 + 2 * LC + i] * a[2 + 2 * LC + k] / nd) / (nd - 1.0);
            cov[i * LC + k] = cov[k * LC + i] = v;
            if (i < L && k < L) mx = fmax(mx, fabs(v));
        }
    for (int i = 0; i < L; ++i)
        for (int k = 0; k < L; ++k) Qb[i * L + k] = (P)(mx > 0 ? cov[i * LC + k] / mx : cov[i * LC + k]);
}

template <class P>
static int mc_center_run(int S, int K, int O, int T, const PlaneView& y, const PlaneView& var, double q, P* ymean,
                         int* n_good_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int B = S * K;
    const int dtype = sizeof(P) == 4 ? EKS_F32 : EKS_F64;
    McWork w;
    const size_t need = mc_carve(dtype, B, T, workspace, &w);
    EKS_REQUIRE(workspace && workspace_bytes >= need, "mc_center: workspace too small (%zu < %zu)", workspace_bytes, need);
    const dim3 grid(B, w.nchunk);   // problems on grid x (no 65535 limit), chunks on y
    mc_maxvar_kernel<P><<<grid, MC_NT, 0, st>>>(var, O, T, (P*)w.maxvar);
    PlaneView mvv;
    mvv.base = w.maxvar; mvv.seq_stride = T;
    for (int i = 0; i < MAX_CHAN; ++i) mvv.chan_off[i] = 0;
    Spans sp; sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = T;
    const size_t sel_bytes = (size_t)B * (sizeof(SelState) + 2 * NBINS * sizeof(int));
    if (int rc = run_const_R<P>(mvv, B, 1, sp, T, 0.0, (P*)w.thr, w.sel, sel_bytes, st, q)) return rc;
    mc_count_kernel<P><<<grid, MC_NT, 0, st>>>((const P*)w.maxvar, (const P*)w.thr, T, w.nchunk, w.cnt, w.last);
    mc_scan_kernel<<<S, MC_NT, 0, st>>>(K, w.nchunk, w.cnt, w.last, w.pre, w.lastbefore, w.ngood, w.minf);
    mc_mean_kernel<P><<<grid, MC_NT, 0, st>>>(y, O, T, w);
    mc_mean_final_kernel<P><<<B, 32, 0, st>>>(O, w, ymean, n_good_out);
    return check_launch("multicam centring kernels");
}

template <class P, int OC>
static int mc_cov_run(int B, int O, int T, const PlaneView& y, const P* ymean, double* moments_out, const McWork& w,
                      cudaStream_t st) {
    mc_cov_kernel<P, OC><<<dim3(B, w.nchunk), MC_NT, 0, st>>>(y, O, T, ymean, w);
    mc_cov_final_kernel<OC><<<B, 32, 0, st>>>(O, w, moments_out);
    return check_launch("multicam PCA moment kernels");
}

template <class P, int LC>
static int mc_latent_run(int B, int O, int L, int T, const PlaneView& y, const P* ymean, const P* pca_mean,
                         const P* comps, P* S0, P* Q, const McWork& w, cudaStream_t st) {
    mc_latent_kernel<P, LC><<<dim3(B, w.nchunk), MC_NT, 0, st>>>(y, O, L, T, ymean, pca_mean, comps, w);
    mc_latent_final_kernel<P, LC><<<B, 32, 0, st>>>(L, w, S0, Q);
    return check_launch("multicam latent initialisation kernels");
}


// ----------------------------------------------------------------------------------------------------
// Mahalanobis variance inflation (eks/stats.py:67-157 compute_mahalanobis, eks/multicam_smoother.py:653-764).
// One iteration of the reference's while-loop is: (1) rows with max-variance < percentile (and, optionally, min
// likelihood >= threshold) feed a FactorAnalysis fit -- the device reduces their moments, the O x O fit is host
// work; (2) per frame: posterior B = (W^T diag(1/(v+eps)) W)^-1, reconstruction, per-view Mahalanobis distance,
// variances of flagged views times `scalar` (in place), a flag per problem if anything was inflated.
// The per-frame algebra runs in fp64 like the reference's NumPy code.
// ----------------------------------------------------------------------------------------------------
template <class P, int OC>
__global__ void __launch_bounds__(MC_NT) mc_valid_moments_kernel(PlaneView y, PlaneView lik, int V, int O, int T,
                                                                 const P* __restrict__ ymean, double lik_thr,
                                                                 int use_q, const int* __restrict__ active, McWork w) {
    constexpr int NACC = 1 + OC + OC * (OC + 1) / 2;
    __shared__ double scratch[32];
    const int b = blockIdx.x, c = blockIdx.y;
    if (active && !active[b]) return;
    const P* yb = reinterpret_cast<const P*>(y.base) + (long long)b * y.seq_stride;
    const P* lb = lik.base ? reinterpret_cast<const P*>(lik.base) + (long long)b * lik.seq_stride : nullptr;
    const P* mv_b = reinterpret_cast<const P*>(w.maxvar) + (long long)b * T;
    const P th = reinterpret_cast<const P*>(w.thr)[b];
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    const int t1 = min(T, (c + 1) * MC_CHUNK);
    for (int t = c * MC_CHUNK + threadIdx.x; t < t1; t += MC_NT) {
        bool valid = use_q ? (mv_b[t] < th) : true;          // strict, stats.py:119
        if (lb) {
            P ml = lb[lik.chan_off[0] + t];
            for (int v = 1; v < V; ++v) { const P x = lb[lik.chan_off[v] + t]; ml = (x < ml || isnan(x)) ? x : ml; }
            valid = valid && ((double)ml >= lik_thr);
        }
        if (!valid) continue;
        double x[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) x[o] = o < O ? (double)(yb[y.chan_off[o] + t] - ymean[(long long)b * O + o]) : 0.0;
        acc[0] += 1.0;
        int q = 1 + OC;
#pragma unroll
        for (int i = 0; i < OC; ++i) {
            acc[1 + i] += x[i];
#pragma unroll
            for (int k = i; k < OC; ++k) acc[q++] += x[i] * x[k];
        }
    }
    double* out = w.part + ((long long)b * w.nchunk + c) * MC_PART_MAX;
#pragma unroll 1
    for (int i = 0; i < NACC; ++i) {
        const double v = block_sum(acc[i], scratch);
        if (threadIdx.x == 0) out[i] = v;
    }
}

template <int OC>
__global__ void __launch_bounds__(32) mc_valid_moments_final_kernel(int O, McWork w, const int* __restrict__ active,
                                                                    double* __restrict__ moments_out) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    double* out = moments_out + (long long)b * (1 + O + O * O);
    const double n = mc_part_sum(w, b, 0);
    if (threadIdx.x == 0) out[0] = n;
    for (int i = 0; i < O; ++i) {
        const double v = mc_part_sum(w, b, 1 + i);
        if (threadIdx.x == 0) out[1 + i] = v;
    }
    int q = 1 + OC;
    for (int i = 0; i < OC; ++i)
        for (int k = i; k < OC; ++k, ++q) {
            if (i >= O || k >= O) continue;
            const double v = mc_part_sum(w, b, q);
            if (threadIdx.x == 0) { out[1 + O + i * O + k] = v; out[1 + O + k * O + i] = v; }
        }
}

// in-place inverse of a symmetric positive definite L x L matrix (L <= 6) by Gauss-Jordan with no pivoting;
// returns false if a pivot is not positive/finite
template <int LC>
__device__ inline bool spd_inverse(double* Mx, int L) {
    double inv[LC * LC];
    for (int i = 0; i < L; ++i) for (int j = 0; j < L; ++j) inv[i * L + j] = i == j ? 1.0 : 0.0;
    for (int c = 0; c < L; ++c) {
        const double piv = Mx[c * L + c];
        if (!(fabs(piv) > 0) || !isfinite(piv)) return false;
        const double ip = 1.0 / piv;
        for (int j = 0; j < L; ++j) { Mx[c * L + j] *= ip; inv[c * L + j] *= ip; }
        for (int i = 0; i < L; ++i) {
            if (i == c) continue;
            const double f = Mx[i * L + c];
            for (int j = 0; j < L; ++j) { Mx[i * L + j] -= f * Mx[c * L + j]; inv[i * L + j] -= f * inv[c * L + j]; }
        }
    }
    for (int i = 0; i < L * L; ++i) Mx[i] = inv[i];
    return true;
}

template <class P, int OC, int LC>
__global__ void __launch_bounds__(MC_NT) mc_inflate_kernel(PlaneView y, PlaneView var, int V, int O, int L, int T,
                                                           const P* __restrict__ ymean, const double* __restrict__ Wm,
                                                           const double* __restrict__ mu, double eps, double threshold,
                                                           double scalar, const int* __restrict__ active,
                                                           int* __restrict__ flags) {
    __shared__ double sW[OC * LC], sMu[OC];
    __shared__ int any_block;
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    for (int i = threadIdx.x; i < O * L; i += MC_NT) sW[i] = Wm[(long long)b * O * L + i];
    for (int i = threadIdx.x; i < O; i += MC_NT) sMu[i] = mu[(long long)b * O + i];
    if (threadIdx.x == 0) any_block = 0;
    __syncthreads();
    const P* yb = reinterpret_cast<const P*>(y.base) + (long long)b * y.seq_stride;
    P* vb = const_cast<P*>(reinterpret_cast<const P*>(var.base)) + (long long)b * var.seq_stride;
    bool any = false;
    const int t1 = min(T, (int)(blockIdx.y + 1) * MC_CHUNK);
    for (int t = blockIdx.y * MC_CHUNK + threadIdx.x; t < t1; t += MC_NT) {
        double x[OC], v[OC], iv[OC];
        for (int o = 0; o < O; ++o) {
            x[o] = (double)(yb[y.chan_off[o] + t] - ymean[(long long)b * O + o]) - sMu[o];
            v[o] = (double)vb[var.chan_off[o] + t];
            iv[o] = 1.0 / (v[o] + eps);
        }
        double Bm[LC * LC], rhs[LC];
        for (int i = 0; i < L; ++i) {
            double r = 0;
            for (int o = 0; o < O; ++o) r += sW[o * L + i] * iv[o] * x[o];
            rhs[i] = r;
            for (int j = 0; j < L; ++j) {
                double a2 = 0;
                for (int o = 0; o < O; ++o) a2 += sW[o * L + i] * iv[o] * sW[o * L + j];
                Bm[i * L + j] = a2;
            }
        }
        const bool okB = spd_inverse<LC>(Bm, L);
        double z[LC];
        for (int i = 0; i < L; ++i) { double r = 0; for (int j = 0; j < L; ++j) r += Bm[i * L + j] * rhs[j]; z[i] = r; }
        bool flag[OC / 2];
        bool anyv = false;
        for (int c = 0; c < V; ++c) {
            double d[2], WB[2][LC], Q[2][2];
            for (int a2 = 0; a2 < 2; ++a2) {
                const int o = 2 * c + a2;
                double xh = 0;
                for (int i = 0; i < L; ++i) xh += sW[o * L + i] * z[i];
                d[a2] = x[o] - xh;
                for (int j = 0; j < L; ++j) { double r = 0; for (int i = 0; i < L; ++i) r += sW[o * L + i] * Bm[i * L + j]; WB[a2][j] = r; }
            }
            for (int a2 = 0; a2 < 2; ++a2)
                for (int e2 = 0; e2 < 2; ++e2) {
                    double r = 0;
                    for (int j = 0; j < L; ++j) r += WB[a2][j] * sW[(2 * c + e2) * L + j];
                    Q[a2][e2] = r + (a2 == e2 ? v[2 * c + a2] : 0.0);
                }
            const double det = Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0];
            const double m = (d[0] * (Q[1][1] * d[0] - Q[0][1] * d[1]) + d[1] * (Q[0][0] * d[1] - Q[1][0] * d[0])) / det;
            flag[c] = okB && (m > threshold);        // NaN compares false, like numpy
            anyv = anyv || flag[c];
        }
        if (V == 2 && anyv) { flag[0] = true; flag[1] = true; }     // multicam_smoother.py:755-757
        if (anyv) {
            any = true;
            for (int c = 0; c < V; ++c)
                if (flag[c]) {
                    vb[var.chan_off[2 * c] + t] = (P)((double)vb[var.chan_off[2 * c] + t] * scalar);
                    vb[var.chan_off[2 * c + 1] + t] = (P)((double)vb[var.chan_off[2 * c + 1] + t] * scalar);
                }
        }
    }
    if (any) any_block = 1;
    __syncthreads();
    if (threadIdx.x == 0 && any_block) atomicExch(flags + b, 1);
}

}  // namespace eks

using namespace eks;

static int make_spans(int T, int n_spans, const int* span_start, const int* span_end, Spans& sp, int& n_total) {
    if (n_spans <= 0) {
        sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = T; n_total = T;
        return 0;
    }
    EKS_REQUIRE(n_spans <= MAX_SPANS, "at most %d frame spans supported on device", MAX_SPANS);
    sp.n = n_spans; sp.cum[0] = 0;
    for (int i = 0; i < n_spans; ++i) {
        EKS_REQUIRE(span_start[i] >= 0 && span_end[i] <= T && span_start[i] < span_end[i], "bad span %d", i);
        sp.start[i] = span_start[i];
        sp.cum[i + 1] = sp.cum[i] + (span_end[i] - span_start[i]);
    }
    n_total = sp.cum[n_spans];
    return 0;
}

static PlaneView make_view(const void* base, long long seq_stride, const long long* chan_off, int O) {
    PlaneView v;
    v.base = base; v.seq_stride = seq_stride;
    for (int i = 0; i < MAX_CHAN; ++i) v.chan_off[i] = i < O ? chan_off[i] : 0;
    return v;
}

extern "C" int eks_initial_guess(const void* var_base, long long seq_stride, const long long* chan_off, int dtype,
                                 int B, int O, int T, double* guess_out, void* s_log0_out, void* stream) {
    EKS_REQUIRE(var_base && chan_off && guess_out, "initial_guess: null pointer");
    EKS_REQUIRE(O >= 1 && O <= MAX_CHAN, "initial_guess: bad channel count %d", O);
    EKS_REQUIRE(T >= 2, "Not enough frames to compute temporal differences.");
    PlaneView v = make_view(var_base, seq_stride, chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32) guess_kernel<float><<<B, 256, 0, st>>>(v, B, O, T, guess_out, (float*)s_log0_out);
    else guess_kernel<double><<<B, 256, 0, st>>>(v, B, O, T, guess_out, (double*)s_log0_out);
    return check_launch("guess_kernel");
}

namespace eks {
// ====================================================================================================
// Geometric initialisation of the calibrated model (eks/multicam_smoother.py:600-650) on the device.
// tri [B][T][3] float64 = triangulated ensemble means.  Per (sequence b, dimension d):
//   m0   = mean of the first min(10, T) frames                                   (:618)
//   S0   = nanvar over all frames (ddof 0, two passes like numpy) + 1e-4         (:619-625)
//   Q    = max((1.4826 (median |dx - median dx| + 1e-12))^2, 1e-8), dx = lag-1 differences   (:632-642)
// np.median is not NaN-aware: any NaN in the column makes Q NaN, which is reproduced.  The two medians are the exact
// radix select of this file on float64 keys (one-pass bracketed form for long sequences).
constexpr int GEO_NT = 1024;

__global__ void __launch_bounds__(GEO_NT) geo_moments_kernel(const double* __restrict__ tri, int T,
                                                             double* __restrict__ m0, double* __restrict__ S0d,
                                                             int* __restrict__ n_nan) {
    __shared__ double scratch[32];
    const int prob = blockIdx.x, b = prob / 3, d = prob - b * 3;
    const double* x = tri + (long long)b * T * 3 + d;
    double s = 0, c = 0;
    for (int t = threadIdx.x; t < T; t += GEO_NT) {
        const double v = x[(long long)t * 3];
        if (!isnan(v)) { s += v; c += 1.0; }
    }
    s = block_sum(s, scratch);
    c = block_sum(c, scratch);
    const double mean = s / c;
    double ss = 0;
    for (int t = threadIdx.x; t < T; t += GEO_NT) {
        const double v = x[(long long)t * 3];
        if (!isnan(v)) { const double e = v - mean; ss = fma(e, e, ss); }
    }
    ss = block_sum(ss, scratch);
    if (threadIdx.x == 0) {
        const int nh = min(10, T);
        double h = 0;
        for (int t = 0; t < nh; ++t) h += x[(long long)t * 3];
        m0[prob] = h / (double)nh;
        S0d[prob] = ss / c + 1e-4;
        n_nan[prob] = T - (int)c;
    }
}

__global__ void __launch_bounds__(256) geo_diff_kernel(const double* __restrict__ tri, int T, double* __restrict__ dx) {
    const int prob = blockIdx.y, b = prob / 3, d = prob - b * 3;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= T - 1) return;
    const double* x = tri + (long long)b * T * 3 + d;
    dx[(long long)prob * (T - 1) + t] = x[(long long)(t + 1) * 3] - x[(long long)t * 3];
}

__global__ void __launch_bounds__(256) geo_absdev_kernel(int n, const double* __restrict__ med, double* __restrict__ dx) {
    const int prob = blockIdx.y;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    double* p = dx + (long long)prob * n + t;
    *p = fabs(*p - med[prob]);
}

__global__ void geo_final_kernel(int nprob, const double* __restrict__ mad, const int* __restrict__ n_nan,
                                 double* __restrict__ Qd) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= nprob) return;
    const double sigma = 1.4826 * (mad[prob] + 1e-12);
    const double v = sigma * sigma;
    Qd[prob] = (n_nan[prob] > 0 || isnan(v)) ? nan("") : fmax(v, 1e-8);      // np.maximum propagates NaN
}

}  // namespace eks

static size_t geo_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t eks_geometric_init_workspace_bytes(int B, int T) {
    const size_t np3 = (size_t)B * 3, n = (size_t)(T > 1 ? T - 1 : 1);
    return geo_align(np3 * n * sizeof(double)) + 3 * geo_align(np3 * sizeof(double)) + geo_align(np3 * sizeof(int)) +
           geo_align(eks_const_R_median_workspace_bytes(EKS_F64, B, 3, (int)n)) + 256;
}

extern "C" int eks_geometric_init(int B, int T, const void* tri, void* m0_out, void* S0_diag_out, void* Q_diag_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(tri && m0_out && S0_diag_out && Q_diag_out, "geometric_init: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 2, "geometric_init: need at least two frames (B=%d, T=%d)", B, T);
    EKS_REQUIRE(workspace && workspace_bytes >= eks_geometric_init_workspace_bytes(B, T),
                "geometric_init: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int np3 = B * 3, n = T - 1;
    unsigned char* w = (unsigned char*)workspace;
    double* dx = (double*)w; w += geo_align((size_t)np3 * n * sizeof(double));
    double* med = (double*)w; w += geo_align((size_t)np3 * sizeof(double));
    double* mad = (double*)w; w += geo_align((size_t)np3 * sizeof(double));
    w += geo_align((size_t)np3 * sizeof(double));
    int* n_nan = (int*)w; w += geo_align((size_t)np3 * sizeof(int));
    const size_t sel_bytes = eks_const_R_median_workspace_bytes(EKS_F64, B, 3, n);
    geo_moments_kernel<<<np3, GEO_NT, 0, st>>>((const double*)tri, T, (double*)m0_out, (double*)S0_diag_out, n_nan);
    const dim3 grid((n + 255) / 256, np3);
    geo_diff_kernel<<<grid, 256, 0, st>>>((const double*)tri, T, dx);
    PlaneView v;
    v.base = dx; v.seq_stride = 3LL * n;
    for (int i = 0; i < MAX_CHAN; ++i) v.chan_off[i] = i < 3 ? (long long)i * n : 0;
    Spans sp; sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = n;
    const double NOFLOOR = -HUGE_VAL;
    if (int rc = run_const_R<double>(v, B, 3, sp, n, NOFLOOR, med, w, sel_bytes, st, -1.0, NOFLOOR)) return rc;
    int launches = 2 + eks_last_launch_count();
    geo_absdev_kernel<<<grid, 256, 0, st>>>(n, med, dx);
    if (int rc = run_const_R<double>(v, B, 3, sp, n, NOFLOOR, mad, w, sel_bytes, st, -1.0, NOFLOOR)) return rc;
    launches += 2 + eks_last_launch_count();
    geo_final_kernel<<<(np3 + 127) / 128, 128, 0, st>>>(np3, mad, n_nan, (double*)Q_diag_out);
    note_launches(launches);
    return check_launch("geometric initialisation kernels");
}

extern "C" size_t eks_const_R_median_workspace_bytes(int dtype, int B, int O, int T) {
    const size_t base = (size_t)B * O * (sizeof(SelState) + 2 * NBINS * sizeof(int));
    if (T < MED_MIN_FRAMES) return base;
    return base + 256 + (dtype == EKS_F32 ? med_extra_bytes(B * O, T, 4, sizeof(MedState<float>))
                                          : med_extra_bytes(B * O, T, 8, sizeof(MedState<double>)));
}

extern "C" int eks_const_R_median(const void* var_base, long long seq_stride, const long long* chan_off, int dtype,
                                  int B, int O, int T, int n_spans, const int* span_start, const int* span_end,
                                  double min_var, void* Rconst_out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    EKS_REQUIRE(var_base && chan_off && Rconst_out, "const_R_median: null pointer");
    EKS_REQUIRE(O >= 1 && O <= MAX_CHAN, "const_R_median: bad channel count %d", O);
    Spans sp; int n_total = 0;
    if (make_spans(T, n_spans, span_start, span_end, sp, n_total)) return -1;
    PlaneView v = make_view(var_base, seq_stride, chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        return run_const_R<float>(v, B, O, sp, n_total, min_var, (float*)Rconst_out, workspace, workspace_bytes, st);
    return run_const_R<double>(v, B, O, sp, n_total, min_var, (double*)Rconst_out, workspace, workspace_bytes, st);
}


extern "C" size_t eks_mc_prestage_workspace_bytes(int dtype, int B, int O, int T) {
    (void)O;
    return mc_carve(dtype, B, T, nullptr, nullptr);
}

extern "C" int eks_mc_center(int dtype, int S, int K, int O, int T, const void* y_base, long long y_seq_stride,
                             const long long* y_chan_off, const void* var_base, long long var_seq_stride,
                             const long long* var_chan_off, double quantile_keep, void* ymean_out, int* n_good_out,
                             void* workspace, size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(y_base && y_chan_off && var_base && var_chan_off && ymean_out, "mc_center: null pointer");
    EKS_REQUIRE(S >= 1 && K >= 1 && T >= 1 && O >= 1 && O <= MAX_CHAN, "mc_center: bad dims");
    EKS_REQUIRE(quantile_keep >= 0.0 && quantile_keep <= 100.0, "Percentiles must be in the range [0, 100]");
    const PlaneView y = make_view(y_base, y_seq_stride, y_chan_off, O);
    const PlaneView v = make_view(var_base, var_seq_stride, var_chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == EKS_F32)
        return mc_center_run<float>(S, K, O, T, y, v, quantile_keep, (float*)ymean_out, n_good_out, workspace,
                                    workspace_bytes, st);
    return mc_center_run<double>(S, K, O, T, y, v, quantile_keep, (double*)ymean_out, n_good_out, workspace,
                                 workspace_bytes, st);
}

extern "C" int eks_mc_pca_moments(int dtype, int B, int O, int T, const void* y_base, long long y_seq_stride,
                                  const long long* y_chan_off, const void* ymean, double* moments_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(y_base && y_chan_off && ymean && moments_out, "mc_pca_moments: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && O >= 1 && O <= MAX_CHAN, "mc_pca_moments: bad dims");
    McWork w;
    const size_t need = mc_carve(dtype, B, T, workspace, &w);
    EKS_REQUIRE(workspace && workspace_bytes >= need, "mc_pca_moments: workspace too small");
    const PlaneView y = make_view(y_base, y_seq_stride, y_chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
#define EKS_MC_COV(PT)                                                                                  \
    if (O <= 4) return mc_cov_run<PT, 4>(B, O, T, y, (const PT*)ymean, moments_out, w, st);             \
    if (O <= 6) return mc_cov_run<PT, 6>(B, O, T, y, (const PT*)ymean, moments_out, w, st);             \
    if (O <= 8) return mc_cov_run<PT, 8>(B, O, T, y, (const PT*)ymean, moments_out, w, st);             \
    return mc_cov_run<PT, MAX_CHAN>(B, O, T, y, (const PT*)ymean, moments_out, w, st);
    if (dtype == EKS_F32) { EKS_MC_COV(float) }
    EKS_MC_COV(double)
#undef EKS_MC_COV
}

extern "C" int eks_mc_latent_init(int dtype, int B, int O, int L, int T, const void* y_base, long long y_seq_stride,
                                  const long long* y_chan_off, const void* ymean, const void* pca_mean,
                                  const void* components, void* S0_out, void* Q_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    EKS_REQUIRE(y_base && y_chan_off && ymean && pca_mean && components && S0_out && Q_out,
                "mc_latent_init: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && O >= 1 && O <= MAX_CHAN && L >= 1 && L <= EKS_MAX_STATE, "mc_latent_init: bad dims");
    McWork w;
    const size_t need = mc_carve(dtype, B, T, workspace, &w);
    EKS_REQUIRE(workspace && workspace_bytes >= need, "mc_latent_init: workspace too small");
    const PlaneView y = make_view(y_base, y_seq_stride, y_chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
#define EKS_MC_LAT(PT)                                                                                             \
    if (L <= 3)                                                                                                    \
        return mc_latent_run<PT, 3>(B, O, L, T, y, (const PT*)ymean, (const PT*)pca_mean, (const PT*)components,   \
                                    (PT*)S0_out, (PT*)Q_out, w, st);                                               \
    return mc_latent_run<PT, EKS_MAX_STATE>(B, O, L, T, y, (const PT*)ymean, (const PT*)pca_mean,                  \
                                            (const PT*)components, (PT*)S0_out, (PT*)Q_out, w, st);
    if (dtype == EKS_F32) { EKS_MC_LAT(float) }
    EKS_MC_LAT(double)
#undef EKS_MC_LAT
}


extern "C" int eks_mc_valid_moments(int dtype, int B, int V, int T, const void* y_base, long long y_seq_stride,
                                    const long long* y_chan_off, const void* ymean, const void* var_base,
                                    long long var_seq_stride, const long long* var_chan_off, const void* lik_base,
                                    long long lik_seq_stride, const long long* lik_chan_off, double lik_threshold,
                                    double v_quantile, const int* active, double* moments_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    const int O = 2 * V;
    EKS_REQUIRE(y_base && y_chan_off && ymean && var_base && var_chan_off && moments_out, "mc_valid_moments: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && V >= 1 && O <= MAX_CHAN, "mc_valid_moments: bad dims");
    McWork w;
    const size_t need = mc_carve(dtype, B, T, workspace, &w);
    EKS_REQUIRE(workspace && workspace_bytes >= need, "mc_valid_moments: workspace too small");
    const PlaneView y = make_view(y_base, y_seq_stride, y_chan_off, O);
    const PlaneView var = make_view(var_base, var_seq_stride, var_chan_off, O);
    PlaneView lik;
    lik.base = nullptr; lik.seq_stride = 0;
    if (lik_base) lik = make_view(lik_base, lik_seq_stride, lik_chan_off, V);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(B, w.nchunk);   // problems on grid x (no 65535 limit), chunks on y
    const int use_q = v_quantile >= 0.0;
    const size_t sel_bytes = (size_t)B * (sizeof(SelState) + 2 * NBINS * sizeof(int));
    Spans sp; sp.n = 1; sp.start[0] = 0; sp.cum[0] = 0; sp.cum[1] = T;
    PlaneView mvv;
    mvv.base = w.maxvar; mvv.seq_stride = T;
    for (int i = 0; i < MAX_CHAN; ++i) mvv.chan_off[i] = 0;
#define EKS_MC_VM(PT, OCV)                                                                                          \
    {                                                                                                               \
        if (use_q) {                                                                                                \
            mc_maxvar_kernel<PT><<<grid, MC_NT, 0, st>>>(var, O, T, (PT*)w.maxvar);                                 \
            if (int rc = run_const_R<PT>(mvv, B, 1, sp, T, 0.0, (PT*)w.thr, w.sel, sel_bytes, st, v_quantile)) return rc; \
        }                                                                                                           \
        mc_valid_moments_kernel<PT, OCV><<<grid, MC_NT, 0, st>>>(y, lik, V, O, T, (const PT*)ymean, lik_threshold,  \
                                                                 use_q, active, w);                                 \
        mc_valid_moments_final_kernel<OCV><<<B, 32, 0, st>>>(O, w, active, moments_out);                            \
        return check_launch("Mahalanobis moment kernels");                                                          \
    }
#define EKS_MC_VM_O(PT)                      \
    if (O <= 4) EKS_MC_VM(PT, 4)             \
    if (O <= 6) EKS_MC_VM(PT, 6)             \
    if (O <= 8) EKS_MC_VM(PT, 8)             \
    EKS_MC_VM(PT, MAX_CHAN)
    if (dtype == EKS_F32) { EKS_MC_VM_O(float) }
    EKS_MC_VM_O(double)
#undef EKS_MC_VM_O
#undef EKS_MC_VM
}

extern "C" int eks_mc_inflate_step(int dtype, int B, int V, int L, int T, const void* y_base, long long y_seq_stride,
                                   const long long* y_chan_off, const void* ymean, void* var_base,
                                   long long var_seq_stride, const long long* var_chan_off, const double* loading,
                                   const double* mean, double epsilon, double threshold, double scalar,
                                   const int* active, int* flags_out, void* stream) {
    const int O = 2 * V;
    EKS_REQUIRE(y_base && y_chan_off && ymean && var_base && var_chan_off && loading && mean && flags_out,
                "mc_inflate_step: null pointer");
    EKS_REQUIRE(B >= 1 && T >= 1 && V >= 2 && O <= MAX_CHAN && L >= 1 && L <= EKS_MAX_STATE,
                "must have >=2 views to inflate variance");
    const PlaneView y = make_view(y_base, y_seq_stride, y_chan_off, O);
    const PlaneView var = make_view(var_base, var_seq_stride, var_chan_off, O);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(flags_out, 0, (size_t)B * sizeof(int), st);
    const dim3 grid(B, (T + MC_CHUNK - 1) / MC_CHUNK);
    if (dtype == EKS_F32)
        mc_inflate_kernel<float, MAX_CHAN, EKS_MAX_STATE><<<grid, MC_NT, 0, st>>>(
            y, var, V, O, L, T, (const float*)ymean, loading, mean, epsilon, threshold, scalar, active, flags_out);
    else
        mc_inflate_kernel<double, MAX_CHAN, EKS_MAX_STATE><<<grid, MC_NT, 0, st>>>(
            y, var, V, O, L, T, (const double*)ymean, loading, mean, epsilon, threshold, scalar, active, flags_out);
    return check_launch("Mahalanobis inflation kernel");
}
