// Label-smoothed KL loss on the decoder logits, forward + gradient fused
// (criterions/hybrid_ctc_attn.py:49-64), and the hybrid mix (:78).
//   target(b,l) = ys[b,l] (l < ylens[b]) | eos (l == ylens[b]) | ignore (models/u2.py:323-328)
//   q = eps/(V-1) off target, 1-eps on target;  row loss = sum_v q (log q - logp)   (0 for ignored rows)
//   dlogits = scale * upstream * (softmax - q)   (0 for ignored rows);  sum / B done by the combine kernel.
// One CTA per token row; V <= 8192 is read once per pass through L1/L2 (rows are <= 20 KB).
#include "common.cuh"

namespace lasr {

template <typename TD>
__global__ void __launch_bounds__(256) lsmooth_kl_kernel(const TD* __restrict__ logits, long ldl, const int64_t* __restrict__ ys,
                                                         const int64_t* __restrict__ ylens, int lmax, int V, float eps,
                                                         float grad_scale, const float* __restrict__ upstream,
                                                         float* __restrict__ row_loss, TD* __restrict__ grad, long ldg) {
    __shared__ float scratch[32];
    const int row = blockIdx.x, L1 = lmax + 1;
    const int b = row / L1, l = row % L1;
    const int yl = (int)ylens[b];
    long tgt = -1;
    if (l < yl) tgt = ys[(long)b * lmax + l];
    else if (l == yl) tgt = V - 1;
    const TD* x = logits + (long)row * ldl;
    TD* g = grad + (long)row * ldg;
    if (tgt < 0) {
        for (int c = threadIdx.x; c < V; c += 256) g[c] = from_f32<TD>(0.f);
        if (threadIdx.x == 0) row_loss[row] = 0.f;
        return;
    }
    float s = grad_scale;
    if (upstream) s *= __ldg(upstream);
    float mx = -INFINITY, tot = 0.f;
    for (int c = threadIdx.x; c < V; c += 256) {
        const float v = to_f32<TD>(x[c]);
        mx = fmaxf(mx, v);
        tot += v;
    }
    mx = block_max(mx, scratch);
    tot = block_sum(tot, scratch);
    float se = 0.f;
    for (int c = threadIdx.x; c < V; c += 256) se += expf(to_f32<TD>(x[c]) - mx);
    se = block_sum(se, scratch);
    const float lse = mx + logf(se);
    const float qo = eps / (float)(V - 1), qt = 1.f - eps;
    const float inv = 1.f / se;
    for (int c = threadIdx.x; c < V; c += 256) {
        const float p = expf(to_f32<TD>(x[c]) - mx) * inv;
        g[c] = from_f32<TD>(s * (p - (c == tgt ? qt : qo)));
    }
    if (threadIdx.x == 0) {
        const float lpt = to_f32<TD>(x[tgt]) - lse;
        const float sum_lp = tot - (float)V * lse;  // sum_v logp_v
        float loss = -(qt * lpt + qo * (sum_lp - lpt));
        if (qt > 0.f) loss += qt * logf(qt);
        if (qo > 0.f) loss += (float)(V - 1) * qo * logf(qo);
        row_loss[row] = loss;
    }
}

// loss = w * sum(nll)/B + (1-w) * sum(row_kl)/B     (single CTA, fixed order -> deterministic)
__global__ void __launch_bounds__(256) hybrid_combine_kernel(const float* __restrict__ nll, int B, const float* __restrict__ row_kl,
                                                             int M, float w, float* __restrict__ out) {
    __shared__ float scratch[32];
    float a = 0.f, k = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) a += nll[i];
    for (int i = threadIdx.x; i < M; i += 256) k += row_kl[i];
    a = block_sum(a, scratch);
    k = block_sum(k, scratch);
    if (threadIdx.x == 0) {
        const float lc = a / (float)B, la = k / (float)B;
        out[0] = w * lc + (1.f - w) * la;
        out[1] = lc;
        out[2] = la;
    }
}

}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_lsmooth_kl_fwdbwd(const void* logits, int dtype, int64_t ldl, const int64_t* ys, const int64_t* ylens, int B, int lmax, int V,
                           float smoothing, float grad_scale, const float* upstream, float* row_loss, void* grad, int64_t ldg,
                           void* stream) {
    LASR_REQUIRE(logits && ys && ylens && row_loss && grad && B > 0 && lmax >= 1 && V > 1, "lsmooth_kl: bad args");
    const int rows = B * (lmax + 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32)
        lsmooth_kl_kernel<float><<<rows, 256, 0, st>>>((const float*)logits, ldl, ys, ylens, lmax, V, smoothing, grad_scale, upstream, row_loss, (float*)grad, ldg);
    else if (dtype == LASR_BF16)
        lsmooth_kl_kernel<bf16><<<rows, 256, 0, st>>>((const bf16*)logits, ldl, ys, ylens, lmax, V, smoothing, grad_scale, upstream, row_loss, (bf16*)grad, ldg);
    else { set_error("lsmooth_kl: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("lsmooth_kl");
}

/* out[0] = ctc_weight * sum(nll)/B + (1 - ctc_weight) * sum(row_kl)/B ; out[1] = ctc term ; out[2] = attention term */
int lasr_hybrid_combine(const float* nll, int B, const float* row_kl, int M, float ctc_weight, float* out, void* stream) {
    LASR_REQUIRE(nll && row_kl && out && B > 0 && M > 0, "hybrid_combine: bad args");
    hybrid_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(nll, B, row_kl, M, ctc_weight, out);
    return check_launch("hybrid_combine");
}

}  // extern "C"
