// Inference-side pieces of the path (models/u2.py:221-317, nets/ctc.py:25-26):
//   lasr_logsoftmax_topk        (device) per frame: log-sum-exp, optional full log-softmax, and the K best classes in
//                               (log-prob descending, index ascending) order -- the `torch.topk(logp, beam)` prune of the
//                               CTC prefix beam search (u2.py:230) and, with K = 1, greedy CTC's argmax.
//   lasr_gather_logp            (device) logp[row, tok[row]] = logits[row, tok] - lse[row]: the attention-rescoring lookups
//                               (u2.py:306-310) without materialising the (beam, L, V) log-softmax.
//   lasr_ctc_prefix_beam_search (HOST)   the prefix search itself (u2.py:224-261) on the pruned (frames, K) matrix.  It is a
//                               sequential dictionary algorithm in float64 Python arithmetic in the reference; it is restated
//                               here in C++ with the same operation order (libm exp/log, left-to-right sums, insertion-ordered
//                               map, stable sort), so n-best lists and scores are bit-identical given the same log-probs.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace lasr {

// ---------------------------------------------------------------------------------------------
// log-softmax + top-K: one CTA (256 threads) per row
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) logsoftmax_topk_kernel(const T* __restrict__ logits, long ld, int V, int K,
                                                              float* __restrict__ lse_out, float* __restrict__ logp_full, long ldf,
                                                              float* __restrict__ top_val, int* __restrict__ top_idx) {
    __shared__ float scratch[32];
    __shared__ float sv[8];
    __shared__ int si[8];
    const long row = blockIdx.x;
    const T* x = logits + row * ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < V; c += 256) mx = fmaxf(mx, to_f32<T>(x[c]));
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int c = threadIdx.x; c < V; c += 256) sum += expf(to_f32<T>(x[c]) - mx);
    sum = block_sum(sum, scratch);
    const float lsum = logf(sum);
    if (threadIdx.x == 0 && lse_out) lse_out[row] = mx + lsum;
    // log-probabilities are formed as (x - max) - log(sum), the rounding sequence of torch.log_softmax: the reference ranks the
    // ROUNDED fp32 log-probs (argmax / torch.topk of log_softmax), where two classes whose logits differ in the last bits can tie
    if (logp_full)
        for (int c = threadIdx.x; c < V; c += 256) logp_full[row * ldf + c] = (to_f32<T>(x[c]) - mx) - lsum;
    // K rounds of arg-max in the total order (log-prob descending, index ascending); round r only considers elements that come
    // strictly after the previous winner in that order, so no "taken" set is needed
    float last_v = INFINITY;
    int last_i = -1;
    for (int r = 0; r < K; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < V; c += 256) {
            const float v = (to_f32<T>(x[c]) - mx) - lsum;
            const bool after = (v < last_v) || (v == last_v && c > last_i);
            if (after && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        bv = sv[0]; bi = si[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        if (threadIdx.x == 0) {
            top_val[row * K + r] = (bi == 0x7fffffff) ? -INFINITY : bv;
            top_idx[row * K + r] = (bi == 0x7fffffff) ? -1 : bi;
        }
        last_v = bv;
        last_i = bi;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_logp_kernel(const T* __restrict__ logits, long ld, const float* __restrict__ lse,
                                                          const int64_t* __restrict__ tok, float* __restrict__ out, long rows, int V) {
    const long r = (long)blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    const long t = tok[r];
    out[r] = (t >= 0 && t < V) ? to_f32<T>(logits[r * ld + t]) - lse[r] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// host: CTC prefix beam search
// ---------------------------------------------------------------------------------------------
static const double NINF = -std::numeric_limits<double>::infinity();

// models/u2.py:367-375
static double log_add(const double* a, int n) {
    bool all_inf = true;
    for (int i = 0; i < n; ++i) all_inf = all_inf && (a[i] == NINF);
    if (all_inf) return NINF;
    double a_max = a[0];
    for (int i = 1; i < n; ++i) a_max = std::max(a_max, a[i]);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += std::exp(a[i] - a_max);
    return a_max + std::log(s);
}

struct Hyp {
    std::vector<int32_t> prefix;
    double pb, pnb;
};

static std::string key_of(const std::vector<int32_t>& p) {
    return std::string(reinterpret_cast<const char*>(p.data()), p.size() * sizeof(int32_t));
}

}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_logsoftmax_topk(const void* logits, int dtype, int64_t ld, int64_t rows, int V, int K, float* lse, float* logp_full,
                         int64_t ldf, float* top_val, int32_t* top_idx, void* stream) {
    LASR_REQUIRE(logits && rows > 0 && V > 0 && K >= 0 && (K == 0 || (top_val && top_idx)) && K <= V, "logsoftmax_topk: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) logsoftmax_topk_kernel<float><<<(unsigned)rows, 256, 0, st>>>((const float*)logits, ld, V, K, lse, logp_full, ldf, top_val, top_idx);
    else if (dtype == LASR_BF16) logsoftmax_topk_kernel<bf16><<<(unsigned)rows, 256, 0, st>>>((const bf16*)logits, ld, V, K, lse, logp_full, ldf, top_val, top_idx);
    else { set_error("logsoftmax_topk: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("logsoftmax_topk");
}

int lasr_gather_logp(const void* logits, int dtype, int64_t ld, const float* lse, const int64_t* tokens, float* out, int64_t rows, int V,
                     void* stream) {
    LASR_REQUIRE(logits && lse && tokens && out && rows > 0 && V > 0, "gather_logp: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) gather_logp_kernel<float><<<ceil_div(rows, 256), 256, 0, st>>>((const float*)logits, ld, lse, tokens, out, rows, V);
    else if (dtype == LASR_BF16) gather_logp_kernel<bf16><<<ceil_div(rows, 256), 256, 0, st>>>((const bf16*)logits, ld, lse, tokens, out, rows, V);
    else { set_error("gather_logp: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("gather_logp");
}

/* HOST function: topk_logp / topk_idx are host arrays (frames, K), row-major, classes in the order the device kernel emits
 * (log-prob descending).  Writes up to `beam` hypotheses best first: tokens (beam, max_len) padded with -1, lengths, scores
 * (= log_add(p_blank, p_nonblank) in float64).  *n_out = number of hypotheses. */
int lasr_ctc_prefix_beam_search(const float* topk_logp, const int32_t* topk_idx, int frames, int K, int beam, int blank,
                                int32_t* out_tokens, int32_t* out_lens, double* out_scores, int max_len, int32_t* n_out) {
    LASR_REQUIRE(topk_logp && topk_idx && out_tokens && out_lens && out_scores && n_out && frames >= 0 && K > 0 && beam > 0 && max_len >= 0,
                 "ctc_prefix_beam_search: bad args");
    std::vector<Hyp> cur(1);
    cur[0].pb = 0.0;
    cur[0].pnb = NINF;
    std::vector<Hyp> next;
    std::unordered_map<std::string, int> index;
    std::vector<std::pair<double, int>> order;
    auto slot = [&](const std::vector<int32_t>& p) -> Hyp& {  // defaultdict(lambda: (-inf, -inf)) in insertion order
        const std::string k = key_of(p);
        auto it = index.find(k);
        if (it != index.end()) return next[it->second];
        index.emplace(k, (int)next.size());
        next.push_back(Hyp{p, NINF, NINF});
        return next.back();
    };
    std::vector<int32_t> ext;
    for (int t = 0; t < frames; ++t) {
        next.clear();
        index.clear();
        for (int j = 0; j < K; ++j) {
            const int s = topk_idx[(long)t * K + j];
            if (s < 0) continue;
            const double ps = (double)topk_logp[(long)t * K + j];
            // NOTE: `cur` is not modified inside this loop; `slot` may reallocate `next` only
            for (size_t h = 0; h < cur.size(); ++h) {
                const std::vector<int32_t>& prefix = cur[h].prefix;
                const double pb = cur[h].pb, pnb = cur[h].pnb;
                const bool has_last = !prefix.empty();
                if (s == blank) {
                    Hyp& e = slot(prefix);
                    const double a[3] = {e.pb, pb + ps, pnb + ps};
                    e.pb = log_add(a, 3);
                } else if (has_last && s == prefix.back()) {
                    {
                        Hyp& e = slot(prefix);  // *ss -> *s
                        const double a[2] = {e.pnb, pnb + ps};
                        e.pnb = log_add(a, 2);
                    }
                    ext = prefix;
                    ext.push_back(s);
                    Hyp& e2 = slot(ext);        // *s-s -> *ss
                    const double a2[2] = {e2.pnb, pb + ps};
                    e2.pnb = log_add(a2, 2);
                } else {
                    ext = prefix;
                    ext.push_back(s);
                    Hyp& e = slot(ext);
                    const double a[3] = {e.pnb, pb + ps, pnb + ps};
                    e.pnb = log_add(a, 3);
                }
            }
        }
        order.clear();
        for (size_t i = 0; i < next.size(); ++i) {
            const double a[2] = {next[i].pb, next[i].pnb};
            order.emplace_back(log_add(a, 2), (int)i);
        }
        std::stable_sort(order.begin(), order.end(), [](const std::pair<double, int>& x, const std::pair<double, int>& y) { return x.first > y.first; });
        const size_t keep = std::min((size_t)beam, order.size());
        std::vector<Hyp> nc;
        nc.reserve(keep);
        for (size_t i = 0; i < keep; ++i) nc.push_back(std::move(next[order[i].second]));
        cur.swap(nc);
    }
    const int n = (int)std::min((size_t)beam, cur.size());
    for (int i = 0; i < n; ++i) {
        const int len = (int)cur[i].prefix.size();
        LASR_REQUIRE(len <= max_len, "ctc_prefix_beam_search: hypothesis of %d tokens exceeds max_len %d", len, max_len);
        for (int j = 0; j < max_len; ++j) out_tokens[(long)i * max_len + j] = j < len ? cur[i].prefix[j] : -1;
        out_lens[i] = len;
        const double a[2] = {cur[i].pb, cur[i].pnb};
        out_scores[i] = log_add(a, 2);
    }
    *n_out = n;
    return LASR_OK;
}

}  // extern "C"
