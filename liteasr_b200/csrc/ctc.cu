// Fused CTC forward-backward + log-softmax backward  (replaces criterions/hybrid_ctc_attn.py:67-75).
//
// Three launches, all on the caller's stream:
//   K_A  ctc_softmax_gather  (all SMs, HBM-bound): one read of the logits, one write of the gradient.
//        per row (t,b): lse = logsumexp(x); grad = s * softmax(x)   (0 and NO read for t >= in_len[b]);
//        lp_ext[b][t][0] = x[blank]-lse, lp_ext[b][t][k+1] = x[label_k]-lse  (compact lattice inputs).
//   K_B  ctc_lattice (one CTA per utterance, latency-bound): thread k owns the state pair
//        (blank 2k, label 2k+1); the only neighbour traffic per step is one warp shuffle (+ one
//        shared-memory edge value per warp); alpha stored once, then overwritten in place by the
//        occupancy exp(alpha+beta-lp+nll) during the beta sweep.  Inputs are register-prefetched one
//        8-step block ahead so global latency never sits on the recursion's critical path.
//   K_C  ctc_scatter (all SMs): grad[t,b,blank] -= s * sum_k occ_blank, grad[t,b,label_k] -= s * occ_label
//        (red.add; only rows t < in_len[b], only L+1 addresses per row).
// Physical traffic is therefore ~1 read + 1 write of the (T,B,V) tensor plus O(T*B*L) lattice state,
// instead of the ~7 dense passes of log_softmax + ctc_loss + their backward kernels.
#include "common.cuh"

namespace lasr {

__device__ __forceinline__ float lse2f(float a, float b) {
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log1pf(expf(fminf(a, b) - m));
}
__device__ __forceinline__ float lse3f(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// ---------------------------------------------------------------------------------------------
// K_A
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    typedef float4 type;
    static __device__ __forceinline__ void unpack(const float4& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    static __device__ __forceinline__ float4 pack(const float* f) { return make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec4<bf16> {
    typedef uint2 type;
    static __device__ __forceinline__ void unpack(const uint2& v, float* f) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x), b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
        f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
    }
    static __device__ __forceinline__ uint2 pack(const float* f) {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
        return u;
    }
};

struct CtcDenseParams {
    const void* logits;
    void* grad;
    long st, sb, gst, gsb;
    const int64_t* targets;
    const int64_t* in_len;
    const int64_t* tgt_len;
    float* lp_ext;
    int T, B, V, lmax, blank;
    float grad_scale;
    const float* upstream;
};

// GROUP threads cooperate on one row; each holds NI chunks of 4 consecutive classes in registers.
template <typename T, int GROUP, int NI>
__global__ void __launch_bounds__(256) ctc_softmax_gather_kernel(const CtcDenseParams p) {
    constexpr int ROWS = 256 / GROUP;
    __shared__ float red[ROWS][8];
    const int g = threadIdx.x / GROUP, gl = threadIdx.x % GROUP;
    const long row = (long)blockIdx.x * ROWS + g;
    const bool row_ok = row < (long)p.T * p.B;
    const int t = row_ok ? (int)(row / p.B) : 0, b = row_ok ? (int)(row % p.B) : 0;
    const int Tb = (int)p.in_len[b];
    const bool live = row_ok && t < Tb;
    const T* x = reinterpret_cast<const T*>(p.logits) + (long)t * p.st + (long)b * p.sb;
    T* gout = reinterpret_cast<T*>(p.grad) + (long)t * p.gst + (long)b * p.gsb;
    typedef typename Vec4<T>::type V4;
    float s = p.grad_scale;
    if (p.upstream) s *= __ldg(p.upstream);

    float v[NI][4];
    float mx = -INFINITY;
    if (live) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = (i * GROUP + gl) * 4;
            if (c + 3 < p.V) {
                const V4 raw = *reinterpret_cast<const V4*>(x + c);
                Vec4<T>::unpack(raw, v[i]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] = (c + j < p.V) ? to_f32<T>(x[c + j]) : -INFINITY;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mx = fmaxf(mx, v[i][j]);
        }
    }
    // group max
    if (GROUP == 32) {
        mx = warp_max(mx);
    } else {
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = red[0][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[0][w]);
        __syncthreads();
    }
    float sum = 0.f;
    if (live) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[i][j] = expf(v[i][j] - mx);  // exp(-inf) = 0 for the padded tail
                sum += v[i][j];
            }
    }
    if (GROUP == 32) {
        sum = warp_sum(sum);
    } else {
        sum = warp_sum(sum);
        if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[0][w];
    }
    if (!row_ok) return;
    if (live) {
        const float k = s / sum;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = (i * GROUP + gl) * 4;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = v[i][j] * k;
            if (c + 3 < p.V) {
                *reinterpret_cast<V4*>(gout + c) = Vec4<T>::pack(o);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < p.V) gout[c + j] = from_f32<T>(o[j]);
            }
        }
        // compact lattice inputs (re-read hits L1/L2: this CTA just streamed the row)
        const float lse = mx + logf(sum);
        const int L = (int)p.tgt_len[b];
        float* le = p.lp_ext + ((long)b * p.T + t) * (p.lmax + 1);
        const int64_t* tg = p.targets + (long)b * p.lmax;
        const float xb = to_f32<T>(x[p.blank]);
        for (int k2 = gl; k2 <= L; k2 += GROUP)
            le[k2] = (k2 == 0) ? (xb - lse) : (to_f32<T>(x[(int)tg[k2 - 1]]) - xb);
    } else {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = (i * GROUP + gl) * 4;
            if (c + 3 < p.V) {
                *reinterpret_cast<V4*>(gout + c) = Vec4<T>::pack(z);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < p.V) gout[c + j] = from_f32<T>(0.f);
            }
        }
    }
}

// generic fallback: any V / alignment, one CTA per row, re-reads through L1/L2
template <typename T>
__global__ void __launch_bounds__(256) ctc_softmax_gather_generic(const CtcDenseParams p) {
    __shared__ float scratch[32];
    const long row = blockIdx.x;
    const int t = (int)(row / p.B), b = (int)(row % p.B);
    const int Tb = (int)p.in_len[b];
    const T* x = reinterpret_cast<const T*>(p.logits) + (long)t * p.st + (long)b * p.sb;
    T* gout = reinterpret_cast<T*>(p.grad) + (long)t * p.gst + (long)b * p.gsb;
    if (t >= Tb) {
        for (int c = threadIdx.x; c < p.V; c += 256) gout[c] = from_f32<T>(0.f);
        return;
    }
    float s = p.grad_scale;
    if (p.upstream) s *= __ldg(p.upstream);
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < p.V; c += 256) mx = fmaxf(mx, to_f32<T>(x[c]));
    mx = block_max(mx, scratch);
    float sum = 0.f;
    for (int c = threadIdx.x; c < p.V; c += 256) sum += expf(to_f32<T>(x[c]) - mx);
    sum = block_sum(sum, scratch);
    const float k = s / sum;
    for (int c = threadIdx.x; c < p.V; c += 256) gout[c] = from_f32<T>(expf(to_f32<T>(x[c]) - mx) * k);
    const float lse = mx + logf(sum);
    const int L = (int)p.tgt_len[b];
    float* le = p.lp_ext + ((long)b * p.T + t) * (p.lmax + 1);
    const int64_t* tg = p.targets + (long)b * p.lmax;
    const float xb = to_f32<T>(x[p.blank]);
    for (int k2 = threadIdx.x; k2 <= L; k2 += 256)
        le[k2] = (k2 == 0) ? (xb - lse) : (to_f32<T>(x[(int)tg[k2 - 1]]) - xb);
}

// ---------------------------------------------------------------------------------------------
// K_B  lattice: thread k owns states (2k = blank, 2k+1 = label k)
// ---------------------------------------------------------------------------------------------
constexpr int PF = 8;  // prefetch block (time steps)

template <int MAXT>
__global__ void __launch_bounds__(MAXT) ctc_lattice_kernel(const float* __restrict__ lp_ext, float2* __restrict__ ab,
                                                           const int64_t* __restrict__ targets,
                                                           const int64_t* __restrict__ in_len,
                                                           const int64_t* __restrict__ tgt_len, float* __restrict__ nll,
                                                           int T, int lmax) {
    __shared__ float edge[2][2][32];  // [parity][value 0/1][warp]
    __shared__ float fin[2];
    const int b = blockIdx.x, k = threadIdx.x, lane = k & 31, warp = k >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    const int Tb = min((int)in_len[b], T), L = min((int)tgt_len[b], lmax);
    const int W = lmax + 1;
    if (Tb <= 0) {
        if (k == 0) nll[b] = (L == 0) ? 0.f : INFINITY;
        return;
    }
    const int64_t* tg = targets + (long)b * lmax;
    const long my = (k < L) ? tg[k] : -1;
    const bool skip_a = (k >= 1 && k < L) && (tg[k - 1] != my);      // alpha: 2k-1 -> 2k+1 allowed
    const bool skip_b = (k + 1 < L) && (tg[k + 1] != my);            // beta : 2k+1 -> 2k+3 allowed
    const bool v_bl = k <= L, v_lb = k < L;
    const float* lpb = lp_ext + (long)b * T * W;
    float2* abb = ab + (long)b * T * W;
    const int kk = min(k, lmax);  // clamp so idle threads read valid memory
    const float NEG = -INFINITY;

    // Normalised recursion: a~_t(s) = alpha_t(s) - sum_{tau<=t} lp_tau(blank) (and the mirror image for
    // beta).  Blank states then carry no emission term, label states add d_t(k) = x[label_k]-x[blank],
    // and the big common offset C = sum_t lp_t(blank) cancels analytically in the occupancy:
    //   occ_t(s) = exp(a~_t(s) + b~_t(s) - e_t(s) - tot~),  e = 0 (blank) | d_t(k) (label),  nll = -(tot~ + C).
    // This keeps |a~| small (fp32 ulp ~1e-6) where raw log-alpha reaches T*log(V) ~ 1e4 (ulp ~1e-3).
    // ---------------- alpha sweep ----------------
    float a_bl, a_lb;
    double csum = 0.0;
    {
        const float l0b = lpb[0], l0d = lpb[min(kk + 1, lmax)];
        csum = (double)l0b;
        a_bl = (k == 0) ? 0.f : NEG;
        a_lb = (k == 0 && L > 0) ? l0d : NEG;
        if (k < W) abb[k] = make_float2(a_bl, a_lb);
        if (lane == 31) edge[0][0][warp] = a_lb;
    }
    __syncthreads();
    float cb[PF], cl[PF], nb_[PF], nl_[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
        const int t = min(1 + i, Tb - 1);
        cb[i] = lpb[(long)t * W];
        cl[i] = lpb[(long)t * W + min(kk + 1, lmax)];
    }
    for (int t0 = 1; t0 < Tb; t0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {  // prefetch next block (clamped; unused values are harmless)
            const int t = min(t0 + PF + i, Tb - 1);
            nb_[i] = lpb[(long)t * W];
            nl_[i] = lpb[(long)t * W + min(kk + 1, lmax)];
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int t = t0 + i;
            if (t < Tb) {  // uniform over the CTA
                float nbv = __shfl_up_sync(0xffffffffu, a_lb, 1);
                if (lane == 0) nbv = (warp > 0) ? edge[(t - 1) & 1][0][warp - 1] : NEG;
                const float n_bl = lse2f(a_bl, nbv);
                const float n_lb = cl[i] + lse3f(a_lb, a_bl, skip_a ? nbv : NEG);
                a_bl = v_bl ? n_bl : NEG;
                a_lb = v_lb ? n_lb : NEG;
                csum += (double)cb[i];
                if (lane == 31) edge[t & 1][0][warp] = a_lb;
                if (k < W) abb[(long)t * W + k] = make_float2(a_bl, a_lb);
                __syncthreads();
            }
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) { cb[i] = nb_[i]; cl[i] = nl_[i]; }
    }
    if (k == L) fin[0] = a_bl;
    if (L > 0 && k == L - 1) fin[1] = a_lb;
    if (L == 0 && k == 0) fin[1] = NEG;
    __syncthreads();
    const float tot = lse2f(fin[0], fin[1]);
    if (k == 0) nll[b] = (float)(-((double)tot + csum));
    const bool feasible = (tot > NEG);
    const float qnan = __int_as_float(0x7fc00000);

    // ---------------- beta sweep (t = Tb-1 handled with alpha still in registers) ----------------
    float b_bl, b_lb;
    {
        const int t = Tb - 1;
        const float ld_ = lpb[(long)t * W + min(kk + 1, lmax)];
        b_bl = (k == L) ? 0.f : NEG;
        b_lb = (L > 0 && k == L - 1) ? ld_ : NEG;
        float o_bl = (a_bl > NEG && b_bl > NEG) ? expf(a_bl + b_bl - tot) : 0.f;
        float o_lb = (a_lb > NEG && b_lb > NEG) ? expf(a_lb + b_lb - ld_ - tot) : 0.f;
        if (!feasible) { o_bl = qnan; o_lb = qnan; }
        if (k < W) abb[(long)t * W + k] = make_float2(o_bl, o_lb);
        if (lane == 0) { edge[t & 1][0][warp] = b_bl; edge[t & 1][1][warp] = b_lb; }
    }
    __syncthreads();
    float2 ca[PF], na[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
        const int t = max(Tb - 2 - i, 0);
        cl[i] = lpb[(long)t * W + min(kk + 1, lmax)];
        ca[i] = abb[(long)t * W + kk];
    }
    for (int t0 = Tb - 2; t0 >= 0; t0 -= PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int t = max(t0 - PF - i, 0);
            nl_[i] = lpb[(long)t * W + min(kk + 1, lmax)];
            na[i] = abb[(long)t * W + kk];
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int t = t0 - i;
            if (t >= 0) {
                float n1 = __shfl_down_sync(0xffffffffu, b_bl, 1);  // beta_{t+1}(2k+2)
                float n2 = __shfl_down_sync(0xffffffffu, b_lb, 1);  // beta_{t+1}(2k+3)
                if (lane == 31) {
                    const bool has = warp + 1 < nwarp;
                    n1 = has ? edge[(t + 1) & 1][0][warp + 1] : NEG;
                    n2 = has ? edge[(t + 1) & 1][1][warp + 1] : NEG;
                }
                const float n_bl = lse2f(b_bl, b_lb);
                const float n_lb = cl[i] + lse3f(b_lb, n1, skip_b ? n2 : NEG);
                b_bl = v_bl ? n_bl : NEG;
                b_lb = v_lb ? n_lb : NEG;
                if (lane == 0) { edge[t & 1][0][warp] = b_bl; edge[t & 1][1][warp] = b_lb; }
                const float al_bl = ca[i].x, al_lb = ca[i].y;
                float o_bl = (al_bl > NEG && b_bl > NEG) ? expf(al_bl + b_bl - tot) : 0.f;
                float o_lb = (al_lb > NEG && b_lb > NEG) ? expf(al_lb + b_lb - cl[i] - tot) : 0.f;
                if (!feasible) { o_bl = qnan; o_lb = qnan; }
                if (k < W) abb[(long)t * W + k] = make_float2(o_bl, o_lb);
                __syncthreads();
            }
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) { cl[i] = nl_[i]; ca[i] = na[i]; }
    }
}

// ---------------------------------------------------------------------------------------------
// K_C  scatter: one warp per live row (t,b)
// ---------------------------------------------------------------------------------------------
template <typename GT>
__global__ void __launch_bounds__(256) ctc_scatter_kernel(const float2* __restrict__ occ, void* grad, long gst, long gsb,
                                                          const int64_t* __restrict__ targets,
                                                          const int64_t* __restrict__ in_len,
                                                          const int64_t* __restrict__ tgt_len, int T, int B, int lmax,
                                                          int blank, float grad_scale, const float* upstream) {
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long)T * B) return;
    const int lane = threadIdx.x & 31;
    const int t = (int)(row / B), b = (int)(row % B);
    if (t >= (int)in_len[b]) return;
    const int L = (int)tgt_len[b], W = lmax + 1;
    float s = grad_scale;
    if (upstream) s *= __ldg(upstream);
    const float2* o = occ + ((long)b * T + t) * W;
    const int64_t* tg = targets + (long)b * lmax;
    GT* g = reinterpret_cast<GT*>(grad) + (long)t * gst + (long)b * gsb;
    float bsum = 0.f;
    for (int k = lane; k <= L; k += 32) {
        const float2 v = o[k];
        bsum += v.x;
        if (k < L) {
            const int c = (int)tg[k];
            if constexpr (sizeof(GT) == 4) atomicAdd(reinterpret_cast<float*>(g) + c, -s * v.y);
            else atomicAdd(reinterpret_cast<bf16*>(g) + c, __float2bfloat16_rn(-s * v.y));
        }
    }
    bsum = warp_sum(bsum);
    if (lane == 0) {
        if constexpr (sizeof(GT) == 4) atomicAdd(reinterpret_cast<float*>(g) + blank, -s * bsum);
        else atomicAdd(reinterpret_cast<bf16*>(g) + blank, __float2bfloat16_rn(-s * bsum));
    }
}

template <typename T>
static int ctc_launch(const CtcDenseParams& p, float2* ab, float* nll, cudaStream_t st) {
    const long rows = (long)p.T * p.B;
    const bool vec = ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.grad)) % (4 * sizeof(T)) == 0) &&
                     p.st % 4 == 0 && p.sb % 4 == 0 && p.gst % 4 == 0 && p.gsb % 4 == 0;
    bool done = false;
    if (vec) {
        const int chunks = (p.V + 3) / 4;
#define LASR_CTC_CASE(GROUP, NI)                                                                        \
    if (!done && chunks <= GROUP * NI) {                                                                \
        ctc_softmax_gather_kernel<T, GROUP, NI><<<ceil_div(rows, 256 / GROUP), 256, 0, st>>>(p);         \
        done = true;                                                                                    \
    }
        LASR_CTC_CASE(32, 1) LASR_CTC_CASE(32, 2) LASR_CTC_CASE(32, 4) LASR_CTC_CASE(32, 8)
        LASR_CTC_CASE(256, 2) LASR_CTC_CASE(256, 4) LASR_CTC_CASE(256, 6) LASR_CTC_CASE(256, 8)
#undef LASR_CTC_CASE
    }
    if (!done) ctc_softmax_gather_generic<T><<<(unsigned)rows, 256, 0, st>>>(p);
    int rc = check_launch("ctc_softmax_gather");
    if (rc) return rc;
    const int threads = ((p.lmax + 1 + 31) / 32) * 32;
    if (threads <= 256)
        ctc_lattice_kernel<256><<<p.B, threads, 0, st>>>(p.lp_ext, ab, p.targets, p.in_len, p.tgt_len, nll, p.T, p.lmax);
    else
        ctc_lattice_kernel<1024><<<p.B, threads, 0, st>>>(p.lp_ext, ab, p.targets, p.in_len, p.tgt_len, nll, p.T, p.lmax);
    rc = check_launch("ctc_lattice");
    if (rc) return rc;
    ctc_scatter_kernel<T><<<ceil_div(rows, 8), 256, 0, st>>>(ab, p.grad, p.gst, p.gsb, p.targets, p.in_len, p.tgt_len,
                                                            p.T, p.B, p.lmax, p.blank, p.grad_scale, p.upstream);
    return check_launch("ctc_scatter");
}

}  // namespace lasr

extern "C" {

size_t lasr_ctc_workspace_bytes(int T, int B, int lmax) {
    const size_t w = (size_t)(lmax < 1 ? 1 : lmax) + 1;
    return (size_t)T * B * w * (sizeof(float) + sizeof(float2)) + 512;
}

int lasr_ctc_fwdbwd(const void* logits, int dtype, int64_t st, int64_t sb, const int64_t* targets, const int64_t* in_len,
                    const int64_t* tgt_len, int T, int B, int V, int lmax, int blank, float grad_scale,
                    const float* upstream, float* nll, void* grad, int64_t gst, int64_t gsb, void* workspace,
                    size_t ws_bytes, void* stream) {
    using namespace lasr;
    LASR_REQUIRE(logits && targets && in_len && tgt_len && nll && grad && workspace, "ctc: null pointer");
    LASR_REQUIRE(T > 0 && B > 0 && V > 1 && lmax >= 1, "ctc: bad shape T=%d B=%d V=%d lmax=%d", T, B, V, lmax);
    LASR_REQUIRE(lmax + 1 <= 1024, "ctc: lmax=%d exceeds the 1023-label lattice CTA", lmax);
    LASR_REQUIRE(blank >= 0 && blank < V, "ctc: bad blank");
    LASR_REQUIRE(ws_bytes >= lasr_ctc_workspace_bytes(T, B, lmax), "ctc: workspace too small");
    CtcDenseParams p;
    p.logits = logits; p.grad = grad; p.st = st; p.sb = sb; p.gst = gst; p.gsb = gsb;
    p.targets = targets; p.in_len = in_len; p.tgt_len = tgt_len;
    const size_t w = (size_t)lmax + 1;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
    float2* ab = reinterpret_cast<float2*>(base);
    p.lp_ext = reinterpret_cast<float*>(base + (size_t)T * B * w * sizeof(float2));
    p.T = T; p.B = B; p.V = V; p.lmax = lmax; p.blank = blank; p.grad_scale = grad_scale; p.upstream = upstream;
    if (dtype == LASR_F32) return ctc_launch<float>(p, ab, nll, (cudaStream_t)stream);
    if (dtype == LASR_BF16) return ctc_launch<bf16>(p, ab, nll, (cudaStream_t)stream);
    set_error("ctc: unsupported dtype %d", dtype);
    return LASR_ERR_UNSUPPORTED;
}

}  // extern "C"
