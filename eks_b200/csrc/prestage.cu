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

template <class P>
__global__ void __launch_bounds__(256) select_hist_kernel(PlaneView var, int O, Spans sp, int n_total, int level,
                                                          const SelState* __restrict__ state,
                                                          int* __restrict__ hist /*[prob][2][NBINS]*/) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ int sh[2][NBINS];
    const int prob = blockIdx.y, b = prob / O, o = prob - b * O;
    for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const P* base = reinterpret_cast<const P*>(var.base) + (long long)b * var.seq_stride + var.chan_off[o];
    const int shift = KT::shift(level), nb = KT::bits(level);
    const int hshift = shift + nb;  // bits above this are already determined
    key_t pre0 = 0, pre1 = 0;
    if (level > 0) { pre0 = (key_t)state[prob].prefix[0]; pre1 = (key_t)state[prob].prefix[1]; }
    const bool same = (level == 0) || (pre0 == pre1);
    const int i0 = blockIdx.x * SEL_CHUNK, i1 = min(n_total, i0 + SEL_CHUNK);
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

template <class P>
__global__ void __launch_bounds__(256) select_scan_kernel(int level, SelState* __restrict__ state,
                                                          int* __restrict__ hist, double floor_lo, double floor_hi,
                                                          P* __restrict__ out) {
    using KT = KeyT<P>;
    using key_t = typename KT::type;
    __shared__ int sh[2][NBINS];
    const int prob = blockIdx.x;
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
            st.rank[0] = (n - 1) / 2;
            st.rank[1] = n / 2;
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
                med = (a + c) * P(0.5);
                // floors: build_R_from_vars clip (1e-12) then min_R_var (np.clip keeps NaN)
                med = (P)fmax((double)med, floor_lo);
                med = (P)fmax((double)med, floor_hi);
            }
            out[prob] = med;
        }
    }
}

template <class P>
int run_const_R(const PlaneView& var, int B, int O, const Spans& sp, int n_total, double min_var, P* out,
                void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int nprob = B * O;
    const size_t need = (size_t)nprob * (sizeof(SelState) + 2 * NBINS * sizeof(int));
    EKS_REQUIRE(workspace && workspace_bytes >= need, "const_R_median: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    SelState* state = reinterpret_cast<SelState*>(workspace);
    int* hist = reinterpret_cast<int*>(state + nprob);
    cudaMemsetAsync(workspace, 0, need, st);
    const int nchunks = (n_total + SEL_CHUNK - 1) / SEL_CHUNK;
    for (int level = 0; level < KeyT<P>::nlevels; ++level) {
        select_hist_kernel<P><<<dim3(nchunks, nprob), 256, 0, st>>>(var, O, sp, n_total, level, state, hist);
        select_scan_kernel<P><<<nprob, 256, 0, st>>>(level, state, hist, 1e-12, min_var, out);
    }
    return check_launch("select kernels");
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

extern "C" size_t eks_const_R_median_workspace_bytes(int B, int O) {
    return (size_t)B * O * (sizeof(SelState) + 2 * NBINS * sizeof(int));
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
