// Conformer convolution-module middle (nets/conformer_convolution.py:49-53 and its backward):
//   GLU -> depthwise Conv1d(k=15, pad 7, groups=d) -> BatchNorm1d (train: batch stats over ALL B*T'
//   frames, padding included -- quirk Q2) -> Swish.
// Layout is (B, T', C) with channels contiguous (what the GEMMs produce/consume), so the reference's
// transposes vanish and every access is coalesced over channels.
// Fast kernels (d % 4 == 0): one CTA per (32-step time chunk, utterance, 128-channel block).  Phase 1 stages the
// 46-row halo tile in shared memory with 64/128-bit loads (all loads of a CTA in flight at once), applying the
// pointwise math (GLU; BatchNorm + Swish backward) on the way in; phase 2 gives each thread one channel pair and
// 8 time steps and slides the 15-tap window through registers (static indexing, conflict-free float2 LDS).
// The *_generic kernels (any even d) keep the register-window-only scheme.
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
__global__ void __launch_bounds__(128) glu_dwconv_fwd_generic(const TD* __restrict__ y2, long ldy, const float* __restrict__ w,
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
__global__ void __launch_bounds__(128) bn_swish_bwd_stats_generic(const TD* __restrict__ da, const float* __restrict__ z,
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

// ---------------------------------------------------------------- bwd 2
template <typename TD>
__global__ void __launch_bounds__(128) dwconv_glu_bwd_generic(const TD* __restrict__ da, const float* __restrict__ z,
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


// =================================================================================================
// fast kernels
// =================================================================================================
constexpr int CB = 128, WIN = TCH + 2 * HALO, TQ = TCH / 4;  // channel block, halo window rows, time steps per thread

template <typename TD> __device__ __forceinline__ float4 ld4c(const TD* p);
template <> __device__ __forceinline__ float4 ld4c<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4c<bf16>(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// ---------------------------------------------------------------- fwd 1 (fast)
template <typename TD>
__global__ void __launch_bounds__(256) glu_dwconv_fwd_kernel(const TD* __restrict__ y2, long ldy, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ z,
                                                             float* __restrict__ partial, int T, int d) {
    __shared__ __align__(16) float gt[WIN][CB];
    __shared__ __align__(16) float red[2][4][CB];
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x, c0 = blockIdx.z * CB;
    const int t0 = chunk * TCH;
    for (int idx = threadIdx.x; idx < WIN * (CB / 4); idx += 256) {
        const int r = idx / (CB / 4), cv = (idx % (CB / 4)) * 4, tt = t0 - HALO + r;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tt >= 0 && tt < T && c0 + cv < d) {
            const TD* row = y2 + ((long)b * T + tt) * ldy + c0 + cv;
            const float4 v = ld4c<TD>(row), gate = ld4c<TD>(row + d);
            g = make_float4(v.x * sigmoidf_(gate.x), v.y * sigmoidf_(gate.y), v.z * sigmoidf_(gate.z), v.w * sigmoidf_(gate.w));
        }
        *reinterpret_cast<float4*>(&gt[r][cv]) = g;
    }
    __syncthreads();
    const int cp = threadIdx.x & 63, qtr = threadIdx.x >> 6, c = 2 * cp, cg = c0 + c;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (cg < d) {
        float w0[KW], w1[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) { w0[k] = w[cg * KW + k]; w1[k] = w[(cg + 1) * KW + k]; }
        const float b0 = bias[cg], b1 = bias[cg + 1];
        const int tb = qtr * TQ;
        float2 win[TQ + KW - 1];
#pragma unroll
        for (int j = 0; j < TQ + KW - 1; ++j) win[j] = *reinterpret_cast<const float2*>(&gt[tb + j][c]);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const int t = t0 + tb + i;
            if (t < T) {
                float a0 = b0, a1 = b1;
#pragma unroll
                for (int k = 0; k < KW; ++k) { a0 = fmaf(w0[k], win[i + k].x, a0); a1 = fmaf(w1[k], win[i + k].y, a1); }
                *reinterpret_cast<float2*>(z + ((long)b * T + t) * d + cg) = make_float2(a0, a1);
                s0 += a0; s1 += a1; q0 += a0 * a0; q1 += a1 * a1;
            }
        }
    }
    red[0][qtr][c] = s0; red[0][qtr][c + 1] = s1;
    red[1][qtr][c] = q0; red[1][qtr][c + 1] = q1;
    __syncthreads();
    if (threadIdx.x < CB && c0 + threadIdx.x < d) {
        const int cc = threadIdx.x;
        float* part = partial + ((long)(b * nchunk + chunk)) * 2 * d;
        part[c0 + cc] = (red[0][0][cc] + red[0][1][cc]) + (red[0][2][cc] + red[0][3][cc]);
        part[d + c0 + cc] = (red[1][0][cc] + red[1][1][cc]) + (red[1][2][cc] + red[1][3][cc]);
    }
}

// ---------------------------------------------------------------- column reduction of [nblk][2][d] partials (double)
// block (32 channels, 8 row groups); FIN = 0: BatchNorm statistics + running stats, FIN = 1: backward sums + dgamma/dbeta
template <int FIN>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ partial, int nblk, int d, long count, float eps,
                                                        float momentum, float* __restrict__ o0, float* __restrict__ o1,
                                                        float* __restrict__ r0, float* __restrict__ r1, int64_t* __restrict__ nbt) {
    __shared__ double sh[2][8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5, c = blockIdx.x * 32 + cx;
    double s = 0.0, q = 0.0;
    if (c < d) {
        for (int i = ry; i < nblk; i += 8) {
            s += (double)partial[(long)i * 2 * d + c];
            q += (double)partial[(long)i * 2 * d + d + c];
        }
    }
    sh[0][ry][cx] = s; sh[1][ry][cx] = q;
    __syncthreads();
    if (ry != 0 || c >= d) return;
#pragma unroll
    for (int j = 1; j < 8; ++j) { s += sh[0][j][cx]; q += sh[1][j][cx]; }
    if (FIN == 0) {
        const double mu = s / (double)count;
        double var = q / (double)count - mu * mu;
        if (var < 0.0) var = 0.0;
        o0[c] = (float)mu;
        o1[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (r0) {
            const double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
            r0[c] = (float)((1.0 - momentum) * (double)r0[c] + momentum * mu);
            r1[c] = (float)((1.0 - momentum) * (double)r1[c] + momentum * unb);
            if (c == 0 && nbt) *nbt += 1;
        }
    } else {
        o0[c] = (float)s;       // sums[0:d]  = sum du
        o0[d + c] = (float)q;   // sums[d:2d] = sum du * zhat
        r0[c] += (float)q;      // dgamma
        r1[c] += (float)s;      // dbeta
    }
}

__global__ void __launch_bounds__(128) bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                                            float eps, float* __restrict__ mean, float* __restrict__ rstd, int d) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= d) return;
    mean[c] = running_mean[c];
    rstd[c] = rsqrtf(running_var[c] + eps);
}

// ---------------------------------------------------------------- bwd 1 (fast): one thread = 4 channels x (32 / groups) rows
template <typename TD>
__global__ void __launch_bounds__(256) bn_swish_bwd_stats_kernel(const TD* __restrict__ da, const float* __restrict__ z,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 float* __restrict__ partial, long rows, int d) {
    __shared__ __align__(16) float4 red[2][256];
    const int vpr = d >> 2;                                 // float4 chunks per row
    const long r0 = (long)blockIdx.x * TCH, r1 = min(rows, r0 + TCH);
    float* part = partial + (long)blockIdx.x * 2 * d;
    for (int cb = 0; cb < vpr; cb += 256) {                 // d <= 1024: one pass
        const int per = min(vpr - cb, 256);                 // chunks handled in this pass (power of two for the supported d)
        const int groups = 256 / per;                       // row groups
        const int cv = cb + (threadIdx.x % per), grp = threadIdx.x / per;
        float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
        if (grp < groups) {
            const int c = cv * 4;
            const float4 m = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
            const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
#pragma unroll 4
            for (long r = r0 + grp; r < r1; r += groups) {
                const float4 zz = *reinterpret_cast<const float4*>(z + r * d + c);
                const float4 g = ld4c<TD>(da + r * d + c);
                const float zh0 = (zz.x - m.x) * rs.x, zh1 = (zz.y - m.y) * rs.y, zh2 = (zz.z - m.z) * rs.z, zh3 = (zz.w - m.w) * rs.w;
                const float du0 = g.x * dswishf_(ga.x * zh0 + be.x), du1 = g.y * dswishf_(ga.y * zh1 + be.y);
                const float du2 = g.z * dswishf_(ga.z * zh2 + be.z), du3 = g.w * dswishf_(ga.w * zh3 + be.w);
                s.x += du0; s.y += du1; s.z += du2; s.w += du3;
                q.x += du0 * zh0; q.y += du1 * zh1; q.z += du2 * zh2; q.w += du3 * zh3;
            }
        }
        red[0][threadIdx.x] = s; red[1][threadIdx.x] = q;
        __syncthreads();
        if (grp == 0) {
            for (int g2 = 1; g2 < groups; ++g2) {
                const float4 a = red[0][threadIdx.x + g2 * per], b2 = red[1][threadIdx.x + g2 * per];
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
                q.x += b2.x; q.y += b2.y; q.z += b2.z; q.w += b2.w;
            }
            *reinterpret_cast<float4*>(part + cv * 4) = s;
            *reinterpret_cast<float4*>(part + d + cv * 4) = q;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- bwd 2 (fast)
// shared: dz tile [WIN][CB], g tile [WIN][CB], sigmoid(gate) tile [TCH][CB]; the dz tile is reused as the reduction buffer
struct BwdSmem {
    float dz[WIN][CB];
    float g[WIN][CB];
    float sg[TCH][CB];
};

template <typename TD>
__global__ void __launch_bounds__(256) dwconv_glu_bwd_kernel(const TD* __restrict__ da, const float* __restrict__ z,
                                                             const TD* __restrict__ y2, long ldy, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ sums,
                                                             const float* __restrict__ w, TD* __restrict__ dy2, long lddy,
                                                             float* __restrict__ dw, float* __restrict__ dbias,
                                                             float* __restrict__ colsum, float* __restrict__ wpartial, int T, int d,
                                                             float inv_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int b = blockIdx.y, c0 = blockIdx.z * CB, t0 = blockIdx.x * TCH;
    // phase 1: dz (BatchNorm + Swish backward) and g (GLU output) for the halo window, sigmoid(gate) for the centre rows
    for (int idx = threadIdx.x; idx < WIN * (CB / 4); idx += 256) {
        const int r = idx / (CB / 4), cv = (idx % (CB / 4)) * 4, tt = t0 - HALO + r;
        float4 dzv = make_float4(0.f, 0.f, 0.f, 0.f), gv = dzv, sgv = dzv;
        if (tt >= 0 && tt < T && c0 + cv < d) {
            const int c = c0 + cv;
            const long rr = (long)b * T + tt;
            const float4 zz = *reinterpret_cast<const float4*>(z + rr * d + c);
            const float4 gd = ld4c<TD>(da + rr * d + c);
            const float4 m = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
            const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
            const float4 su = *reinterpret_cast<const float4*>(sums + c), sq = *reinterpret_cast<const float4*>(sums + d + c);
            const float zh0 = (zz.x - m.x) * rs.x, zh1 = (zz.y - m.y) * rs.y, zh2 = (zz.z - m.z) * rs.z, zh3 = (zz.w - m.w) * rs.w;
            dzv.x = ga.x * rs.x * (gd.x * dswishf_(ga.x * zh0 + be.x) - su.x * inv_count - zh0 * sq.x * inv_count);
            dzv.y = ga.y * rs.y * (gd.y * dswishf_(ga.y * zh1 + be.y) - su.y * inv_count - zh1 * sq.y * inv_count);
            dzv.z = ga.z * rs.z * (gd.z * dswishf_(ga.z * zh2 + be.z) - su.z * inv_count - zh2 * sq.z * inv_count);
            dzv.w = ga.w * rs.w * (gd.w * dswishf_(ga.w * zh3 + be.w) - su.w * inv_count - zh3 * sq.w * inv_count);
            const TD* row = y2 + rr * ldy + c;
            const float4 v = ld4c<TD>(row), gate = ld4c<TD>(row + d);
            sgv = make_float4(sigmoidf_(gate.x), sigmoidf_(gate.y), sigmoidf_(gate.z), sigmoidf_(gate.w));
            gv = make_float4(v.x * sgv.x, v.y * sgv.y, v.z * sgv.z, v.w * sgv.w);
        }
        *reinterpret_cast<float4*>(&sm.dz[r][cv]) = dzv;
        *reinterpret_cast<float4*>(&sm.g[r][cv]) = gv;
        if (r >= HALO && r < HALO + TCH) *reinterpret_cast<float4*>(&sm.sg[r - HALO][cv]) = sgv;
    }
    __syncthreads();
    // phase 2
    const int cp = threadIdx.x & 63, qtr = threadIdx.x >> 6, c = 2 * cp, cg = c0 + c;
    const bool act = cg < d;
    float aw0[KW], aw1[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) { aw0[k] = 0.f; aw1[k] = 0.f; }
    float ab0 = 0.f, ab1 = 0.f, cv0 = 0.f, cv1 = 0.f, cg0 = 0.f, cg1 = 0.f;
    if (act) {
        const int tb = qtr * TQ;
        {
            float w0[KW], w1[KW];
#pragma unroll
            for (int k = 0; k < KW; ++k) { w0[k] = w[cg * KW + k]; w1[k] = w[(cg + 1) * KW + k]; }
            float2 dzw[TQ + KW - 1];
#pragma unroll
            for (int j = 0; j < TQ + KW - 1; ++j) dzw[j] = *reinterpret_cast<const float2*>(&sm.dz[tb + j][c]);
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                const int t = t0 + tb + i;
                // z[t'] = sum_k w[k] g[t'+k-7]  =>  dg[t] = sum_k w[k] dz[t+7-k]   (window index i + 14 - k)
                float dg0 = 0.f, dg1 = 0.f;
#pragma unroll
                for (int k = 0; k < KW; ++k) { dg0 = fmaf(w0[k], dzw[i + KW - 1 - k].x, dg0); dg1 = fmaf(w1[k], dzw[i + KW - 1 - k].y, dg1); }
                if (t < T) {
                    const float2 sg = *reinterpret_cast<const float2*>(&sm.sg[tb + i][c]);
                    const float2 gg = *reinterpret_cast<const float2*>(&sm.g[tb + i + HALO][c]);
                    const float v0 = dg0 * sg.x, v1 = dg1 * sg.y;                              // d(value half)
                    const float g0 = dg0 * gg.x * (1.f - sg.x), g1 = dg1 * gg.y * (1.f - sg.y);  // d(gate half): v*s*(1-s) = g*(1-s)
                    TD* orow = dy2 + ((long)b * T + t) * lddy + cg;
                    st2<TD>(orow, v0, v1);
                    st2<TD>(orow + d, g0, g1);
                    cv0 += v0; cv1 += v1; cg0 += g0; cg1 += g1;
                }
            }
        }
        {
            // dw[k] += dz[t] * g[t+k-7] ; db += dz[t]     (dz[t] = 0 beyond the utterance end)
            float2 gw[TQ + KW - 1];
#pragma unroll
            for (int j = 0; j < TQ + KW - 1; ++j) gw[j] = *reinterpret_cast<const float2*>(&sm.g[tb + j][c]);
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                const float2 dzc = *reinterpret_cast<const float2*>(&sm.dz[tb + i + HALO][c]);
#pragma unroll
                for (int k = 0; k < KW; ++k) { aw0[k] = fmaf(dzc.x, gw[i + k].x, aw0[k]); aw1[k] = fmaf(dzc.y, gw[i + k].y, aw1[k]); }
                ab0 += dzc.x; ab1 += dzc.y;
            }
        }
    }
    __syncthreads();  // every thread is done with the tiles: reuse sm.dz as the reduction buffer [4][NRED][CB]
    constexpr int NRED = KW + 3;
    float* red = &sm.dz[0][0];
    {
        float* mine = red + (long)qtr * NRED * CB;
#pragma unroll
        for (int k = 0; k < KW; ++k) { mine[k * CB + c] = aw0[k]; mine[k * CB + c + 1] = aw1[k]; }
        mine[KW * CB + c] = ab0; mine[KW * CB + c + 1] = ab1;
        mine[(KW + 1) * CB + c] = cv0; mine[(KW + 1) * CB + c + 1] = cv1;
        mine[(KW + 2) * CB + c] = cg0; mine[(KW + 2) * CB + c + 1] = cg1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NRED * CB; idx += 256) {
        const int k = idx / CB, cc = idx % CB;
        if (c0 + cc >= d) continue;
        const float v = (red[idx] + red[NRED * CB + idx]) + (red[2 * NRED * CB + idx] + red[3 * NRED * CB + idx]);
        if (wpartial) {  // deterministic two-stage reduction: [CTA (b, chunk)][NRED][d], summed by dwconv_reduce_kernel
            wpartial[(((long)b * gridDim.x + blockIdx.x) * NRED + k) * d + c0 + cc] = v;
            continue;
        }
        if (k < KW) atomicAdd(dw + (long)(c0 + cc) * KW + k, v);
        else if (k == KW) atomicAdd(dbias + c0 + cc, v);
        else if (colsum) atomicAdd(colsum + (k == KW + 1 ? 0 : d) + c0 + cc, v);
    }
}

// second stage of the depthwise-weight / bias / pointwise-bias gradient: block (32 channels, 8 row groups), grid (d/32, KW+3)
__global__ void __launch_bounds__(256) dwconv_reduce_kernel(const float* __restrict__ wpartial, int nblk, int d, float* __restrict__ dw,
                                                            float* __restrict__ dbias, float* __restrict__ colsum) {
    __shared__ float sh[8][33];
    constexpr int NRED = KW + 3;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5, c = blockIdx.x * 32 + cx, k = blockIdx.y;
    float s = 0.f;
    if (c < d)
        for (int i = ry; i < nblk; i += 8) s += wpartial[((long)i * NRED + k) * d + c];
    sh[ry][cx] = s;
    __syncthreads();
    if (ry != 0 || c >= d) return;
#pragma unroll
    for (int j = 1; j < 8; ++j) s += sh[j][cx];
    if (k < KW) dw[(long)c * KW + k] += s;
    else if (k == KW) dbias[c] += s;
    else if (colsum) colsum[(k == KW + 1 ? 0 : d) + c] += s;
}

}  // namespace lasr

extern "C" {
using namespace lasr;

static inline bool fast_d(int d) { return d % 4 == 0; }
static inline bool pow2_rows(int d) { const int v = d / 4; return d % 4 == 0 && v >= 1 && v <= 256 && (v & (v - 1)) == 0; }

/* partial: B * ceil(T/32) * 2 * d floats */
int lasr_glu_dwconv_fwd(const void* y2, int dtype, int64_t ldy, const float* w, const float* bias, float* z, float* partial,
                        int B, int T, int d, void* stream) {
    LASR_REQUIRE(y2 && w && bias && z && partial && B > 0 && T > 0 && d > 0 && d % 2 == 0 && ldy % 2 == 0, "glu_dwconv_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast = fast_d(d) && ldy % 4 == 0 && ((uintptr_t)y2 & 15) == 0 && ((uintptr_t)z & 15) == 0;
    if (fast) {
        dim3 grid(ceil_div(T, TCH), B, ceil_div(d, CB));
        if (dtype == LASR_F32) glu_dwconv_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)y2, ldy, w, bias, z, partial, T, d);
        else if (dtype == LASR_BF16) glu_dwconv_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)y2, ldy, w, bias, z, partial, T, d);
        else { set_error("glu_dwconv_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    } else {
        dim3 grid(ceil_div(T, TCH), B);
        if (dtype == LASR_F32) glu_dwconv_fwd_generic<float><<<grid, 128, 0, st>>>((const float*)y2, ldy, w, bias, z, partial, T, d);
        else if (dtype == LASR_BF16) glu_dwconv_fwd_generic<bf16><<<grid, 128, 0, st>>>((const bf16*)y2, ldy, w, bias, z, partial, T, d);
        else { set_error("glu_dwconv_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    }
    return check_launch("glu_dwconv_fwd");
}

int lasr_bn_finalize(const float* partial, int nblk, int d, int64_t count, float eps, float momentum, float* mean, float* rstd,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, int training, void* stream) {
    LASR_REQUIRE(mean && rstd && d > 0 && (training ? (partial && nblk > 0 && count > 0) : (running_mean && running_var)), "bn_finalize: bad args");
    LASR_REQUIRE(!training || ((running_mean != nullptr) == (running_var != nullptr)), "bn_finalize: running stats come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    if (!training) bn_eval_stats_kernel<<<ceil_div(d, 128), 128, 0, st>>>(running_mean, running_var, eps, mean, rstd, d);
    else bn_reduce_kernel<0><<<ceil_div(d, 32), 256, 0, st>>>(partial, nblk, d, count, eps, momentum, mean, rstd, running_mean, running_var,
                                                             num_batches_tracked);
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
    const bool fast = pow2_rows(d) && ((uintptr_t)da & 15) == 0 && ((uintptr_t)z & 15) == 0 && ((uintptr_t)partial & 15) == 0 &&
                      (((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
    if (dtype != LASR_F32 && dtype != LASR_BF16) { set_error("bn_swish_bwd_stats: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    if (fast) {
        if (dtype == LASR_F32) bn_swish_bwd_stats_kernel<float><<<nblk, 256, 0, st>>>((const float*)da, z, mean, rstd, gamma, beta, partial, rows, d);
        else bn_swish_bwd_stats_kernel<bf16><<<nblk, 256, 0, st>>>((const bf16*)da, z, mean, rstd, gamma, beta, partial, rows, d);
    } else {
        if (dtype == LASR_F32) bn_swish_bwd_stats_generic<float><<<nblk, 128, 0, st>>>((const float*)da, z, mean, rstd, gamma, beta, partial, rows, d);
        else bn_swish_bwd_stats_generic<bf16><<<nblk, 128, 0, st>>>((const bf16*)da, z, mean, rstd, gamma, beta, partial, rows, d);
    }
    int rc = check_launch("bn_swish_bwd_stats");
    if (rc) return rc;
    bn_reduce_kernel<1><<<ceil_div(d, 32), 256, 0, st>>>(partial, nblk, d, 0, 0.f, 0.f, sums, nullptr, dgamma, dbeta, nullptr);
    return check_launch("bn_bwd_finalize");
}

int lasr_dwconv_glu_bwd(const void* da, const float* z, const void* y2, int dtype, int64_t ldy, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, const float* sums, const float* w, void* dy2, int64_t lddy, float* dw,
                        float* dbias, float* colsum, float* wpartial, int B, int T, int d, void* stream) {
    LASR_REQUIRE(da && z && y2 && mean && rstd && gamma && beta && sums && w && dy2 && dw && dbias && B > 0 && T > 0 && d % 2 == 0 &&
                     ldy % 2 == 0 && lddy % 2 == 0, "dwconv_glu_bwd: bad args");
    const float inv = 1.f / (float)((long)B * T);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != LASR_F32 && dtype != LASR_BF16) { set_error("dwconv_glu_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    const bool fast = fast_d(d) && ldy % 4 == 0 && lddy % 2 == 0 &&
                      (((uintptr_t)da | (uintptr_t)z | (uintptr_t)y2 | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma | (uintptr_t)beta |
                        (uintptr_t)sums) & 15) == 0;
    if (fast) {
        dim3 grid(ceil_div(T, TCH), B, ceil_div(d, CB));
        const int smem = (int)sizeof(BwdSmem);
        static bool configured = false;
        if (!configured) {
            if (cudaFuncSetAttribute(dwconv_glu_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
                cudaFuncSetAttribute(dwconv_glu_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
                return check_launch("dwconv_glu_bwd smem attr");
            configured = true;
        }
        if (dtype == LASR_F32)
            dwconv_glu_bwd_kernel<float><<<grid, 256, smem, st>>>((const float*)da, z, (const float*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                                 (float*)dy2, lddy, dw, dbias, colsum, wpartial, T, d, inv);
        else
            dwconv_glu_bwd_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)da, z, (const bf16*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                                (bf16*)dy2, lddy, dw, dbias, colsum, wpartial, T, d, inv);
        int rc = check_launch("dwconv_glu_bwd");
        if (rc || !wpartial) return rc;
        dwconv_reduce_kernel<<<dim3(ceil_div(d, 32), KW + 3), 256, 0, st>>>(wpartial, B * ceil_div(T, TCH), d, dw, dbias, colsum);
        return check_launch("dwconv_reduce");
    }
    dim3 grid(ceil_div(T, TCH), B);
    if (dtype == LASR_F32)
        dwconv_glu_bwd_generic<float><<<grid, 128, 0, st>>>((const float*)da, z, (const float*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                           (float*)dy2, lddy, dw, dbias, T, d, inv);
    else
        dwconv_glu_bwd_generic<bf16><<<grid, 128, 0, st>>>((const bf16*)da, z, (const bf16*)y2, ldy, mean, rstd, gamma, beta, sums, w,
                                                          (bf16*)dy2, lddy, dw, dbias, T, d, inv);
    int rc = check_launch("dwconv_glu_bwd");
    if (rc || !colsum) return rc;
    return lasr_act_bwd(dy2, lddy, nullptr, 0, nullptr, 0, colsum, B * T, 2 * d, LASR_ACT_NONE, 1.f, dtype, stream);
}

}  // extern "C"
