// Attention score post-processing, fused: legacy rel_shift + scale + key-padding / causal mask + softmax
// (nets/attention.py:46-59,99-118,145-152) and the matching backward.
//
// The dense contractions (Q.K^T, Q.P^T, probs.V and their transposes) run on the tcgen05 GEMM; these kernels
// do everything between them in ONE pass per direction, one warp per score row held in registers:
//   fwd : s[i,j] = (ac[i,j] + bdshift[i,j]) * scale ; masked -> -1e38 ; p = softmax_j(s)   (bf16/fp32 out)
//         bdshift[i,j] = bd[i, Tk-1-(i-j)] (j <= i) | 0 (j == i+1) | bd[i+1, j-i-2] (j > i+1)      (quirk Q1)
//   bwd : ds = p * (dp - sum_j p*dp) * scale  -> dac ; dbd = inverse-shift gather of ds (second tiny kernel)
// Key mask: key j of batch b is valid iff j < klen(b);  klen mode: 0 none, 1 lens[b], 2 lens[b]+1,
// 3 #{j : 4j < lens[b]} (the encoder's mask[:, :-2:2][:, :-2:2] of a padding mask, nets/transformer_encoder.py:118).
// Masked scores are -1e38 exactly like masked_fill (no -inf, no post-softmax zeroing: quirk Q4).
#include "common.cuh"

namespace lasr {

constexpr int SM_MAXE = 32;  // elements per lane -> Tk <= 1024
constexpr float MASK_FILL = -1e38f;

__device__ __forceinline__ int key_len(const int64_t* lens, int mode, int b, int Tk) {
    if (mode == 0 || !lens) return Tk;
    const long l = lens[b];
    long k = l;
    if (mode == 2) k = l + 1;
    else if (mode == 3) k = (l + 3) / 4;
    return (int)(k < Tk ? (k < 0 ? 0 : k) : Tk);
}

template <typename TP>
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(const float* __restrict__ ac, const float* __restrict__ bd,
                                                               TP* __restrict__ probs, const int64_t* __restrict__ lens,
                                                               int mask_mode, int causal, float scale, int B, int H, int Tq,
                                                               int Tk, int ld) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long)B * H * Tq) return;
    const int i = (int)(row % Tq);
    const int b = (int)(row / ((long)H * Tq));
    const int klen = key_len(lens, mask_mode, b, Tk);
    const float* ar = ac + row * ld;
    const float* br = bd ? bd + row * ld : nullptr;  // bd row i of the same (b,h); row i+1 is br + ld
    float s[SM_MAXE];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < SM_MAXE; ++e) {
        const int j = lane + 32 * e;
        s[e] = -INFINITY;
        if (j < Tk) {
            float v = ar[j];
            if (br) {
                float sh = 0.f;
                if (j <= i) sh = br[Tk - 1 - i + j];
                else if (j > i + 1) sh = br[ld + j - i - 2];
                v += sh;
            }
            v *= scale;
            if (j >= klen || (causal && j > i)) v = MASK_FILL;
            s[e] = v;
            mx = fmaxf(mx, v);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < SM_MAXE; ++e) {
        const int j = lane + 32 * e;
        if (j < Tk) {
            s[e] = expf(s[e] - mx);
            sum += s[e];
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    TP* pr = probs + row * ld;
#pragma unroll
    for (int e = 0; e < SM_MAXE; ++e) {
        const int j = lane + 32 * e;
        if (j < ld) pr[j] = from_f32<TP>(j < Tk ? s[e] * inv : 0.f);
    }
}

template <typename TP>
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(const TP* __restrict__ probs, const float* __restrict__ dprobs,
                                                               TP* __restrict__ dsc, float scale, long rows, int Tk, int ld) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const TP* pr = probs + row * ld;
    const float* dr = dprobs + row * ld;
    float p[SM_MAXE], g[SM_MAXE];
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < SM_MAXE; ++e) {
        const int j = lane + 32 * e;
        p[e] = 0.f; g[e] = 0.f;
        if (j < Tk) {
            p[e] = to_f32<TP>(pr[j]);
            g[e] = dr[j];
            dot += p[e] * g[e];
        }
    }
    dot = warp_sum(dot);
    TP* o = dsc + row * ld;
#pragma unroll
    for (int e = 0; e < SM_MAXE; ++e) {
        const int j = lane + 32 * e;
        if (j < ld) o[j] = from_f32<TP>(j < Tk ? p[e] * (g[e] - dot) * scale : 0.f);
    }
}

// dbd[r, m] = ds[r-1, r+m+1] if m < T-1-r (0 when r == 0) else ds[r, m-(T-1-r)]      (inverse of the legacy shift)
template <typename TP>
__global__ void __launch_bounds__(256) rel_shift_bwd_kernel(const TP* __restrict__ ds, TP* __restrict__ dbd, long nmat, int T,
                                                            int ld) {
    const long total = nmat * T * ld;
    for (long idx = (long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long)gridDim.x * 256) {
        const int m = (int)(idx % ld);
        const long rr = idx / ld;
        const int r = (int)(rr % T);
        const long mat = rr / T;
        TP v = from_f32<TP>(0.f);
        if (m < T) {
            const TP* base = ds + mat * T * ld;
            if (m < T - 1 - r) {
                if (r >= 1) v = base[(long)(r - 1) * ld + r + m + 1];
            } else {
                v = base[(long)r * ld + m - (T - 1 - r)];
            }
        }
        dbd[idx] = v;
    }
}

}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_attn_softmax_fwd(const float* ac, const float* bd, void* probs, int p_dtype, const int64_t* lens, int mask_mode, int causal,
                          float scale, int B, int H, int Tq, int Tk, int ld, void* stream) {
    LASR_REQUIRE(ac && probs && B > 0 && H > 0 && Tq > 0 && Tk > 0 && ld >= Tk && ld <= 32 * SM_MAXE, "attn_softmax_fwd: bad args (Tk<=1024)");
    LASR_REQUIRE(!bd || Tq == Tk, "attn_softmax_fwd: rel_shift needs Tq == Tk");
    LASR_REQUIRE(mask_mode >= 0 && mask_mode <= 3 && (mask_mode == 0 || lens), "attn_softmax_fwd: bad mask mode");
    const long rows = (long)B * H * Tq;
    cudaStream_t st = (cudaStream_t)stream;
    if (p_dtype == LASR_F32) attn_softmax_fwd_kernel<float><<<ceil_div(rows, 8), 256, 0, st>>>(ac, bd, (float*)probs, lens, mask_mode, causal, scale, B, H, Tq, Tk, ld);
    else if (p_dtype == LASR_BF16) attn_softmax_fwd_kernel<bf16><<<ceil_div(rows, 8), 256, 0, st>>>(ac, bd, (bf16*)probs, lens, mask_mode, causal, scale, B, H, Tq, Tk, ld);
    else { set_error("attn_softmax_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("attn_softmax_fwd");
}

int lasr_attn_softmax_bwd(const void* probs, const float* dprobs, void* dscores, void* dbd, int dtype, float scale, int B, int H, int Tq,
                          int Tk, int ld, void* stream) {
    LASR_REQUIRE(probs && dprobs && dscores && B > 0 && H > 0 && Tq > 0 && Tk > 0 && ld >= Tk && ld <= 32 * SM_MAXE, "attn_softmax_bwd: bad args");
    LASR_REQUIRE(!dbd || Tq == Tk, "attn_softmax_bwd: rel_shift needs Tq == Tk");
    const long rows = (long)B * H * Tq;
    cudaStream_t st = (cudaStream_t)stream;
    int grid2 = ceil_div(rows * ld, 256);
    if (grid2 > 148 * 32) grid2 = 148 * 32;
    if (dtype == LASR_F32) {
        attn_softmax_bwd_kernel<float><<<ceil_div(rows, 8), 256, 0, st>>>((const float*)probs, dprobs, (float*)dscores, scale, rows, Tk, ld);
        if (dbd && check_launch("attn_softmax_bwd")) return LASR_ERR_CUDA;
        if (dbd) rel_shift_bwd_kernel<float><<<grid2, 256, 0, st>>>((const float*)dscores, (float*)dbd, (long)B * H, Tq, ld);
    } else if (dtype == LASR_BF16) {
        attn_softmax_bwd_kernel<bf16><<<ceil_div(rows, 8), 256, 0, st>>>((const bf16*)probs, dprobs, (bf16*)dscores, scale, rows, Tk, ld);
        if (dbd && check_launch("attn_softmax_bwd")) return LASR_ERR_CUDA;
        if (dbd) rel_shift_bwd_kernel<bf16><<<grid2, 256, 0, st>>>((const bf16*)dscores, (bf16*)dbd, (long)B * H, Tq, ld);
    } else { set_error("attn_softmax_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("attn_softmax_bwd");
}

}  // extern "C"
