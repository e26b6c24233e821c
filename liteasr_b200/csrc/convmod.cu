// Conformer convolution-module middle (nets/conformer_convolution.py:49-53 and its backward):
//   GLU -> depthwise Conv1d(k=15, pad 7, groups=d) -> BatchNorm1d (train: batch stats over ALL B*T'
//   frames, padding included -- quirk Q2) -> Swish.
// Layout is (B, T', C) with channels contiguous (what the GEMMs produce/consume), so the reference's
// transposes vanish and every access is coalesced over channels.  Each thread owns one channel pair
// and slides a 15-deep register window along time (no shared memory needed: neighbouring time steps
// are re-used from registers, neighbouring channels never interact).
//   fwd 1: z = dwconv(glu(y2)) + per-(batch,chunk) partial sum / sum-of-squares   (deterministic)
//   fwd 2: finalize stats (double), running-stat update (momentum 0.1, unbiased var)
//   fwd 3: a = swish(gamma * (z - mean) * rstd + beta)
//   bwd 1: partial sums of du and du*zhat  (du = da * swish'(u))  -> also dgamma / dbeta
//   bwd 2: dz on the fly -> depthwise dgrad + wgrad -> GLU backward, one pass
#include "common.cuh"

namespace lasr {

constexpr int KW = 15, HALO = 7, TCH = 32;

template <typename TD> __device__ __forceinline__ float2 ld2(const TD* p);
template <> __device__ __forceinline__ float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld2<bf16>(const bf16* p) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
    return make_float2(__low2float(h), __high2float(h));
}
template <typename TD> __device__ __forceinline__ void st2(TD* p, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void st2<bf16>(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// ---------------------------------------------------------------- fwd 1
template <typename TD>
__global__ void __launch_bounds__(128) glu_dwconv_fwd_kernel(const TD* __restrict__ y2, long ldy, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ z,
                                                             float* __restrict__ partial, int T, int d) {
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
    const int t0 = chunk * TCH, t1 = min(T, t0 + TCH);
    float* part = partial + ((long)(b * nchunk + chunk)) * 2 * d;
    for (int c = threadIdx.x * 2; c < d; c += blockDim.x * 2) {
        float w0[KW], w1[KW], r0[KW], r1[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) { w0[k] = w[c * KW + k]; w1[k] = w[(c + 1) * KW + k]; r0[k] = 0.f; r1[k] = 0.f; }
        const float b0 = bias[c], b1 = bias[c + 1];
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (int tt = t0 - HALO; tt < t1 + HALO; ++tt) {
            float g0 = 0.f, g1 = 0.f;
            if (tt >= 0 && tt < T) {
                const TD* row = y2 + ((long)b * T + tt) * ldy;
                const float2 v = ld2<TD>(row + c), gt = ld2<TD>(row + d + c);
                g0 = v.x * sigmoidf_(gt.x);
                g1 = v.y * sigmoidf_(gt.y);
            }
#pragma unroll
            for (int k = 0; k < KW - 1; ++k) { r0[k] = r0[k + 1]; r1[k] = r1[k + 1]; }
            r0[KW - 1] = g0; r1[KW - 1] = g1;
            const int t = tt - HALO;  // window now holds g[t-7 .. t+7]
            if (t >= t0) {
                float a0 = b0, a1 = b1;
#pragma unroll
                for (int k = 0; k < KW; ++k) { a0 = fmaf(w0[k], r0[k], a0); a1 = fmaf(w1[k], r1[k], a1); }
                *reinterpret_cast<float2*>(z + ((long)b * T + t) * d + c) = make_float2(a0, a1);
                s0 += a0; s1 += a1; q0 += a0 * a0; q1 += a1 * a1;
            }
        }
        part[c] = s0; part[c + 1] = s1;
        part[d + c] = q0; part[d + c + 1] = q1;
    }
}

// ---------------------------------------------------------------- fwd 2 (also used for eval: running stats)
__global__ void __launch_bounds__(128) bn_finalize_kernel(const float* __restrict__ partial, int nblk, int d, long count,
                                                          float eps, float momentum, float* __restrict__ mean,
                                                          float* __restrict__ rstd, float* __restrict__ running_mean,
                                                          float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                                          int training) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= d) return;
    if (!training) {
        mean[c] = running_mean[c];
        rstd[c] = rsqrtf(running_var[c] + eps);
        return;
    }
    double s = 0.0, q = 0.0;
    for (int i = 0; i < nblk; ++i) {
        s += (double)partial[(long)i * 2 * d + c];
        q += (double)partial[(long)i * 2 * d + d + c];
    }
    const double mu = s / (double)count;
    double var = q / (double)count - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
        if (c == 0 && nbt) *nbt += 1;
    }
}

// ---------------------------------------------------------------- fwd 3
template <typename TD>
__global__ void __launch_bounds__(256) bn_swish_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, TD* __restrict__ a, long n, int d) {
    const long i = ((long)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= n) return;
    const int c = (int)(i % d);
    const float2 v = *reinterpret_cast<const float2*>(z + i);
    const float u0 = (v.x - mean[c]) * rstd[c] * gamma[c] + beta[c];
    const float u1 = (v.y - mean[c + 1]) * rstd[c + 1] * gamma[c + 1] + beta[c + 1];
    st2<TD>(a + i, swishf_(u0), swishf_(u1));
}

// ---------------------------------------------------------------- bwd 1
template <typename TD>
__global__ void __launch_bounds__(128) bn_swish_bwd_stats_kernel(const TD* __restrict__ da, const float* __restrict__ z,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 float* __restrict__ partial, long rows, int d) {
    const long r0 = (long)blockIdx.x * TCH, r1 = min(rows, r0 + TCH);
    float* part = partial + (long)blockIdx.x * 2 * d;
    for (int c = threadIdx.x * 2; c < d; c += blockDim.x * 2) {
        const float m0 = mean[c], m1 = mean[c + 1], rs0 = rstd[c], rs1 = rstd[c + 1];
        const float g0 = gamma[c], g1 = gamma[c + 1], be0 = beta[c], be1 = beta[c + 1];
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (long r = r0; r < r1; ++r) {
            const float2 zz = *reinterpret_cast<const float2*>(z + r * d + c);
            const float2 g = ld2<TD>(da + r * d + c);
            const float zh0 = (zz.x - m0) * rs0, zh1 = (zz.y - m1) * rs1;
            const float du0 = g.x * dswishf_(g0 * zh0 + be0), du1 = g.y * dswishf_(g1 * zh1 + be1);
            s0 += du0; s1 += du1; q0 += du0 * zh0; q1 += du1 * zh1;
        }
        part[c] = s0; part[c + 1] = s1;
        part[d + c] = q0; part[d + c + 1] = q1;
    }
}

__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int d,
                                                              float* __restrict__ sums, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= d) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < nblk; ++i) {
        s += (double)partial[(long)i * 2 * d + c];
        q += (double)partial[(long)i * 2 * d + d + c];
    }
    sums[c] = (float)s;
    sums[d + c] = (float)q;
    dgamma[c] += (float)q;
    dbeta[c] += (float)s;
}

// ---------------------------------------------------------------- bwd 2
template <typename TD>
__global__ void __launch_bounds__(128) dwconv_glu_bwd_kernel(const TD* __restrict__ da, const float* __restrict__ z,
                                                             const TD* __restrict__ y2, long ldy, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ sums,
                                                             const float* __restrict__ w, TD* __restrict__ dy2, long lddy,
                                                             float* __restrict__ dw, float* __restrict__ dbias, int T, int d,
                                                             float inv_count) {
    const int b = blockIdx.y, t0 = blockIdx.x * TCH, t1 = min(T, t0 + TCH);
    for (int c = threadIdx.x * 2; c < d; c += blockDim.x * 2) {
        float w0[KW], w1[KW], z0[KW], z1[KW], g0r[KW], g1r[KW], aw0[KW], aw1[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            w0[k] = w[c * KW + k]; w1[k] = w[(c + 1) * KW + k];
            z0[k] = z1[k] = g0r[k] = g1r[k] = aw0[k] = aw1[k] = 0.f;
        }
        const float m0 = mean[c], m1 = mean[c + 1], rs0 = rstd[c], rs1 = rstd[c + 1];
        const float ga0 = gamma[c], ga1 = gamma[c + 1], be0 = beta[c], be1 = beta[c + 1];
        const float ms0 = sums[c] * inv_count, ms1 = sums[c + 1] * inv_count;          // mean(du)
        const float mq0 = sums[d + c] * inv_count, mq1 = sums[d + c + 1] * inv_count;  // mean(du * zhat)
        float ab0 = 0.f, ab1 = 0.f;
        for (int tt = t0 - HALO; tt < t1 + HALO; ++tt) {
            float dz0 = 0.f, dz1 = 0.f, g0 = 0.f, g1 = 0.f;
            if (tt >= 0 && tt < T) {
                const long r = (long)b * T + tt;
                const float2 zz = *reinterpret_cast<const float2*>(z + r * d + c);
                const float2 gd = ld2<TD>(da + r * d + c);
                const float zh0 = (zz.x - m0) * rs0, zh1 = (zz.y - m1) * rs1;
                const float du0 = gd.x * dswishf_(ga0 * zh0 + be0), du1 = gd.y * dswishf_(ga1 * zh1 + be1);
                dz0 = ga0 * rs0 * (du0 - ms0 - zh0 * mq0);
                dz1 = ga1 * rs1 * (du1 - ms1 - zh1 * mq1);
                const TD* row = y2 + r * ldy;
                const float2 v = ld2<TD>(row + c), gt = ld2<TD>(row + d + c);
                g0 = v.x * sigmoidf_(gt.x);
                g1 = v.y * sigmoidf_(gt.y);
            }
#pragma unroll
            for (int k = 0; k < KW - 1; ++k) { z0[k] = z0[k + 1]; z1[k] = z1[k + 1]; g0r[k] = g0r[k + 1]; g1r[k] = g1r[k + 1]; }
            z0[KW - 1] = dz0; z1[KW - 1] = dz1; g0r[KW - 1] = g0; g1r[KW - 1] = g1;
            const int t = tt - HALO;  // windows hold dz / g at times t-7 .. t+7 (index j <-> t-7+j)
            if (t >= t0) {
                // z[t'] = sum_k w[k] g[t'+k-7]  =>  dg[t] = sum_k w[k] dz[t-k+7] (index 14-k)
                float dg0 = 0.f, dg1 = 0.f;
#pragma unroll
                for (int k = 0; k < KW; ++k) { dg0 = fmaf(w0[k], z0[KW - 1 - k], dg0); dg1 = fmaf(w1[k], z1[KW - 1 - k], dg1); }
                // dw[k] += dz[t] * g[t+k-7] ; db += dz[t]
                const float c0 = z0[HALO], c1 = z1[HALO];
#pragma unroll
                for (int k = 0; k < KW; ++k) { aw0[k] = fmaf(c0, g0r[k], aw0[k]); aw1[k] = fmaf(c1, g1r[k], aw1[k]); }
                ab0 += c0; ab1 += c1;
                // GLU backward at t
                const long r = (long)b * T + t;
                const TD* row = y2 + r * ldy;
                const float2 v = ld2<TD>(row + c), gt = ld2<TD>(row + d + c);
                const float sg0 = sigmoidf_(gt.x), sg1 = sigmoidf_(gt.y);
                TD* orow = dy2 + r * lddy;
                st2<TD>(orow + c, dg0 * sg0, dg1 * sg1);
                st2<TD>(orow + d + c, dg0 * v.x * sg0 * (1.f - sg0), dg1 * v.y * sg1 * (1.f - sg1));
            }
        }
#pragma unroll
        for (int k = 0; k < KW; ++k) { atomicAdd(dw + c * KW + k, aw0[k]); atomicAdd(dw + (c + 1) * KW + k, aw1[k]); }
        atomicAdd(dbias + c, ab0);
        atomicAdd(dbias + c + 1, ab1);
    }
}

}  // namespace lasr

extern "C" {
using namespace lasr;

/* partial: B * ceil(T/32) * 2 * d floats */
int lasr_glu_dwconv_fwd(const void* y2, int dtype, int64_t ldy, const float* w, const float* bias, float* z, float* partial,
                        int B, int T, int d, void* stream) {
    LASR_REQUIRE(y2 && w && bias && z && partial && B > 0 && T > 0 && d > 0 && d % 2 == 0 && ldy % 2 == 0, "glu_dwconv_fwd: bad args");
    dim3 grid(ceil_div(T, TCH), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) glu_dwconv_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)y2, ldy, w, bias, z, partial, T, d);
    else if (dtype == LASR_BF16) glu_dwconv_fwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)y2, ldy, w, bias, z, partial, T, d);
    else { set_error("glu_dwconv_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("glu_dwconv_fwd");
}

int lasr_bn_finalize(const float* partial, int nblk, int d, int64_t count, float eps, float momentum, float* mean, float* rstd,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, int training, void* stream) {
    LASR_REQUIRE(mean && rstd && d > 0 && (training ? (partial && nblk > 0 && count > 0) : (running_mean && running_var)), "bn_finalize: bad args");
    bn_finalize_kernel<<<ceil_div(d, 128), 128, 0, (cudaStream_t)stream>>>(partial, nblk, d, count, eps, momentum, mean, rstd,
                                                                          running_mean, running_var, num_batches_tracked, training);
    return check_launch("bn_finalize");
}

int lasr_bn_swish_fwd(const float* z, const float* mean, const float* rstd, const float* gamma, const float* beta, void* a,
                      int dtype, int64_t rows, int d, void* stream) {
    LASR_REQUIRE(z && mean && rstd && gamma && beta && a && rows > 0 && d % 2 == 0, "bn_swish_fwd: bad args");
    const long n = rows * d;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) bn_swish_fwd_kernel<float><<<ceil_div(n / 2, 256), 256, 0, st>>>(z, mean, rstd, gamma, beta, (float*)a, n, d);
    else if (dtype == LASR_BF16) bn_swish_fwd_kernel<bf16><<<ceil_div(n / 2, 256), 256, 0, st>>>(z, mean, rstd, gamma, beta, (bf16*)a, n, d);
    else { set_error("bn_swish_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("bn_swish_fwd");
}

/* partial: ceil(rows/32) * 2 * d floats; sums: 2 * d floats */
int lasr_bn_swish_bwd_stats(const void* da, int dtype, const float* z, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, float* partial, float* sums, float* dgamma, float* dbeta, int64_t rows, int d,
                            void* stream) {
    LASR_REQUIRE(da && z && mean && rstd && gamma && beta && partial && sums && dgamma && dbeta && rows > 0 && d % 2 == 0, "bn_swish_bwd_stats: bad args");
    const int nblk = ceil_div(rows, TCH);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) bn_swish_bwd_stats_kernel<float><<<nblk, 128, 0, st>>>((const float*)da, z, mean, rstd, gamma, beta, partial, rows, d);
    else if (dtype == LASR_BF16) bn_swish_bwd_stats_kernel<bf16><<<nblk, 128, 0, st>>>((const bf16*)da, z, mean, rstd, gamma, beta, partial, rows, d);
    else { set_error("bn_swish_bwd_stats: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    int rc = check_launch("bn_swish_bwd_stats");
    if (rc) return rc;
    bn_bwd_finalize_kernel<<<ceil_div(d, 128), 128, 0, st>>>(partial, nblk, d, sums, dgamma, dbeta);
    return check_launch("bn_bwd_finalize");
}

int lasr_dwconv_glu_bwd(const void* da, const float* z, const void* y2, int dtype, int64_t ldy, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, const float* sums, const float* w, void* dy2, int64_t lddy, float* dw,
                        float* dbias, int B, int T, int d, void* stream) {
    LASR_REQUIRE(da && z && y2 && mean && rstd && gamma && beta && sums && w && dy2 && dw && dbias && B > 0 && T > 0 && d % 2 == 0 &&
                     ldy % 2 == 0 && lddy % 2 == 0, "dwconv_glu_bwd: bad args");
    dim3 grid(ceil_div(T, TCH), B);
    const float inv = 1.f / (float)((long)B * T);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32)
        dwconv_glu_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)da, z, (const float*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                          (float*)dy2, lddy, dw, dbias, T, d, inv);
    else if (dtype == LASR_BF16)
        dwconv_glu_bwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)da, z, (const bf16*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                         (bf16*)dy2, lddy, dw, dbias, T, d, inv);
    else { set_error("dwconv_glu_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("dwconv_glu_bwd");
}

}  // extern "C"
