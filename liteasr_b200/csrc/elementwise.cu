// Bandwidth kernels around the GEMMs: casts / relayouts, activation backward + bias gradient,
// rel-pos bias add (nets/attention.py:135-139), decoder embedding + positional encoding
// (nets/transformer_decoder.py:77-78, nets/positional_encoding.py:49-56).
// All are coalesced along the feature dimension with 128-bit (fp32) / 64-bit (bf16) accesses.
#include "common.cuh"

namespace lasr {

template <typename TD>
__device__ __forceinline__ float4 ld4(const TD* p) {
    if constexpr (sizeof(TD) == 4) {
        return *reinterpret_cast<const float4*>(p);
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    }
}
template <typename TD>
__device__ __forceinline__ void st4(TD* p, float4 v) {
    if constexpr (sizeof(TD) == 4) {
        *reinterpret_cast<float4*>(p) = v;
    } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
}

// ------------------------------------------------------------------ cast fp32 -> bf16 (flat)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
    LASR_PDL_SYNC();
    const long i4 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        st4<bf16>(dst + i4, *reinterpret_cast<const float4*>(src + i4));
    } else {
        for (long i = i4; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
    }
}

// ------------------------------------------------------------------ generic 4-D strided copy/cast
struct Perm4 {
    long n[4], ss[4], ds[4];
};
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) permute4d_kernel(const TS* __restrict__ src, TD* __restrict__ dst, Perm4 p,
                                                        int accumulate) {
    const long total = p.n[0] * p.n[1] * p.n[2] * p.n[3];
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        long r = i;
        const long i3 = r % p.n[3]; r /= p.n[3];
        const long i2 = r % p.n[2]; r /= p.n[2];
        const long i1 = r % p.n[1]; r /= p.n[1];
        const long i0 = r;
        const long so = i0 * p.ss[0] + i1 * p.ss[1] + i2 * p.ss[2] + i3 * p.ss[3];
        const long dO = i0 * p.ds[0] + i1 * p.ds[1] + i2 * p.ds[2] + i3 * p.ds[3];
        float v = to_f32<TS>(src[so]);
        if (accumulate) v += to_f32<TD>(dst[dO]);
        dst[dO] = from_f32<TD>(v);
    }
}

// ------------------------------------------------------------------ dH = dA * act'(.), dbias += colsum(dH)
// saved: pre-activation H for swish, activation output A for relu (relu'(h) = a > 0), unused for none.
template <typename TD>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TD* __restrict__ da, long ldda, const TD* __restrict__ saved,
                                                      long lds, TD* __restrict__ dh, long lddh,
                                                      float* __restrict__ dbias, int rows, int cols, int act, float scale) {
    LASR_PDL_SYNC();
    __shared__ float4 red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int c = (blockIdx.x * 64 + tx) * 4;
    const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
    const bool full = c + 3 < cols;  // ragged tail (cols % 4 != 0, e.g. V = 4233) goes element-wise
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < cols) {
        for (int r = r0 + ty; r < r1; r += 4) {
            float g[4], sv[4] = {0.f, 0.f, 0.f, 0.f};
            if (full) {
                const float4 t = ld4<TD>(da + (long)r * ldda + c);
                g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
                if (act != LASR_ACT_NONE) {
                    const float4 u = ld4<TD>(saved + (long)r * lds + c);
                    sv[0] = u.x; sv[1] = u.y; sv[2] = u.z; sv[3] = u.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = (c + j < cols) ? to_f32<TD>(da[(long)r * ldda + c + j]) : 0.f;
                    if (act != LASR_ACT_NONE && c + j < cols) sv[j] = to_f32<TD>(saved[(long)r * lds + c + j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                g[j] *= scale;
                if (act == LASR_ACT_SWISH) g[j] *= dswishf_(sv[j]);
                else if (act == LASR_ACT_RELU) g[j] = sv[j] > 0.f ? g[j] : 0.f;
                acc[j] += g[j];
            }
            if (dh) {
                if (full) st4<TD>(dh + (long)r * lddh + c, make_float4(g[0], g[1], g[2], g[3]));
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c + j < cols) dh[(long)r * lddh + c + j] = from_f32<TD>(g[j]);
                }
            }
        }
    }
    if (!dbias) return;
    red[ty][tx] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (ty == 0 && c < cols) {
        float4 t = red[0][tx];
#pragma unroll
        for (int w = 1; w < 4; ++w) { t.x += red[w][tx].x; t.y += red[w][tx].y; t.z += red[w][tx].z; t.w += red[w][tx].w; }
        atomicAdd(dbias + c, t.x);
        if (c + 1 < cols) atomicAdd(dbias + c + 1, t.y);
        if (c + 2 < cols) atomicAdd(dbias + c + 2, t.z);
        if (c + 3 < cols) atomicAdd(dbias + c + 3, t.w);
    }
}

// ------------------------------------------------------------------ q + pos_bias_u / q + pos_bias_v
template <typename TD>
__global__ void __launch_bounds__(256) pos_bias_fwd_kernel(const TD* __restrict__ q, long ldq, const float* __restrict__ u,
                                                           const float* __restrict__ v, TD* __restrict__ qu,
                                                           TD* __restrict__ qv, long ldo, int rows, int d) {
    LASR_PDL_SYNC();
    const int per_row = d >> 2;
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long)rows * per_row) return;
    const long r = i / per_row;
    const int c = (int)(i % per_row) * 4;
    const float4 x = ld4<TD>(q + r * ldq + c);
    const float4 bu = *reinterpret_cast<const float4*>(u + c), bv = *reinterpret_cast<const float4*>(v + c);
    st4<TD>(qu + r * ldo + c, make_float4(x.x + bu.x, x.y + bu.y, x.z + bu.z, x.w + bu.w));
    st4<TD>(qv + r * ldo + c, make_float4(x.x + bv.x, x.y + bv.y, x.z + bv.z, x.w + bv.w));
}

// dq = dqu + dqv ; du += colsum(dqu) ; dv += colsum(dqv)
template <typename TD>
__global__ void __launch_bounds__(256) pos_bias_bwd_kernel(const TD* __restrict__ dqu, const TD* __restrict__ dqv, long ldi,
                                                           TD* __restrict__ dq, long ldq, float* __restrict__ du,
                                                           float* __restrict__ dv, float* __restrict__ dqb, int rows, int d) {
    LASR_PDL_SYNC();
    __shared__ float4 red[2][4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int c = (blockIdx.x * 64 + tx) * 4;
    const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
    float4 au = make_float4(0, 0, 0, 0), av = make_float4(0, 0, 0, 0);
    if (c < d) {
        // 8 rows per trip with all 16 loads issued before the first use (the kernel ran at 1.7 TB/s with one row in flight)
        int r = r0 + ty;
        for (; r + 28 < r1; r += 32) {
            float4 a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = ld4<TD>(dqu + (long)(r + 4 * i) * ldi + c);
                b[i] = ld4<TD>(dqv + (long)(r + 4 * i) * ldi + c);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                st4<TD>(dq + (long)(r + 4 * i) * ldq + c, make_float4(a[i].x + b[i].x, a[i].y + b[i].y, a[i].z + b[i].z, a[i].w + b[i].w));
                au.x += a[i].x; au.y += a[i].y; au.z += a[i].z; au.w += a[i].w;
                av.x += b[i].x; av.y += b[i].y; av.z += b[i].z; av.w += b[i].w;
            }
        }
        for (; r < r1; r += 4) {
            const float4 a = ld4<TD>(dqu + (long)r * ldi + c), b = ld4<TD>(dqv + (long)r * ldi + c);
            st4<TD>(dq + (long)r * ldq + c, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
            au.x += a.x; au.y += a.y; au.z += a.z; au.w += a.w;
            av.x += b.x; av.y += b.y; av.z += b.z; av.w += b.w;
        }
    }
    red[0][ty][tx] = au;
    red[1][ty][tx] = av;
    __syncthreads();
    if (ty == 0 && c < d) {
        float4 t = red[0][0][tx], s = red[1][0][tx];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
            t.x += red[0][w][tx].x; t.y += red[0][w][tx].y; t.z += red[0][w][tx].z; t.w += red[0][w][tx].w;
            s.x += red[1][w][tx].x; s.y += red[1][w][tx].y; s.z += red[1][w][tx].z; s.w += red[1][w][tx].w;
        }
        atomicAdd(du + c, t.x); atomicAdd(du + c + 1, t.y); atomicAdd(du + c + 2, t.z); atomicAdd(du + c + 3, t.w);
        atomicAdd(dv + c, s.x); atomicAdd(dv + c + 1, s.y); atomicAdd(dv + c + 2, s.z); atomicAdd(dv + c + 3, s.w);
        if (dqb) {  // bias gradient of linear_q: colsum(dq) = colsum(dqu) + colsum(dqv)
            atomicAdd(dqb + c, t.x + s.x); atomicAdd(dqb + c + 1, t.y + s.y); atomicAdd(dqb + c + 2, t.z + s.z); atomicAdd(dqb + c + 3, t.w + s.w);
        }
    }
}

// bf16, d % 8 == 0: 16-byte loads (8 columns per thread), ~256 rows per CTA so that a launch issues one set of column-sum atomics
// per CTA of a single wave instead of one per 64 rows (589 CTAs x 768 scalar atomics on 768 addresses at C2 / B = 126: 32 us for
// 58 MB = 1.8 TB/s)
__device__ __forceinline__ void acc_bf16x8(float (&acc)[8], const uint4& u) {
    acc[0] += __uint_as_float(u.x << 16); acc[1] += __uint_as_float(u.x & 0xffff0000u);
    acc[2] += __uint_as_float(u.y << 16); acc[3] += __uint_as_float(u.y & 0xffff0000u);
    acc[4] += __uint_as_float(u.z << 16); acc[5] += __uint_as_float(u.z & 0xffff0000u);
    acc[6] += __uint_as_float(u.w << 16); acc[7] += __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(a << 16) + __uint_as_float(b << 16),
                                                   __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u));
    return *reinterpret_cast<const uint32_t*>(&h);
}
__global__ void __launch_bounds__(256) pos_bias_bwd_wide_kernel(const bf16* __restrict__ dqu, const bf16* __restrict__ dqv, long ldi,
                                                                bf16* __restrict__ dq, long ldq, float* __restrict__ du,
                                                                float* __restrict__ dv, float* __restrict__ dqb, int rows, int d,
                                                                int rows_per_cta) {
    LASR_PDL_SYNC();
    __shared__ float red[2][8][32][9];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 8;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float au[8], av[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { au[e] = 0.f; av[e] = 0.f; }
    if (c < d) {
        int r = r0 + ty;
        for (; r + 24 < r1; r += 32) {  // 4 rows per trip, all 8 loads issued before the first use
            uint4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = *reinterpret_cast<const uint4*>(dqu + (long)(r + 8 * i) * ldi + c);
                b[i] = *reinterpret_cast<const uint4*>(dqv + (long)(r + 8 * i) * ldi + c);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 o;
                o.x = add_bf16x2(a[i].x, b[i].x); o.y = add_bf16x2(a[i].y, b[i].y); o.z = add_bf16x2(a[i].z, b[i].z); o.w = add_bf16x2(a[i].w, b[i].w);
                *reinterpret_cast<uint4*>(dq + (long)(r + 8 * i) * ldq + c) = o;
                acc_bf16x8(au, a[i]);
                acc_bf16x8(av, b[i]);
            }
        }
        for (; r < r1; r += 8) {
            const uint4 a = *reinterpret_cast<const uint4*>(dqu + (long)r * ldi + c), b = *reinterpret_cast<const uint4*>(dqv + (long)r * ldi + c);
            uint4 o;
            o.x = add_bf16x2(a.x, b.x); o.y = add_bf16x2(a.y, b.y); o.z = add_bf16x2(a.z, b.z); o.w = add_bf16x2(a.w, b.w);
            *reinterpret_cast<uint4*>(dq + (long)r * ldq + c) = o;
            acc_bf16x8(au, a);
            acc_bf16x8(av, b);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[0][ty][tx][e] = au[e]; red[1][ty][tx][e] = av[e]; }
    __syncthreads();
    // thread -> one column of the CTA's 256: sum the 8 row lanes, then one atomic per column and tensor
    const int col = threadIdx.x, cx = col >> 3, ce = col & 7, cg = blockIdx.x * 256 + col;
    if (cg < d) {
        float t = 0.f, u = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { t += red[0][w][cx][ce]; u += red[1][w][cx][ce]; }
        atomicAdd(du + cg, t);
        atomicAdd(dv + cg, u);
        if (dqb) atomicAdd(dqb + cg, t + u);  // bias gradient of linear_q: colsum(dq) = colsum(dqu) + colsum(dqv)
    }
}

// ------------------------------------------------------------------ decoder embedding + PE
// out[b,l] = emb[tokens[b,l]] * scale + pe[l]   (nets/transformer_decoder.py:77-78, positional_encoding.py:49-56)
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t* __restrict__ tokens, int L, const float* __restrict__ emb,
                                                        const float* __restrict__ pe, float* __restrict__ out, int B, int d,
                                                        float scale) {
    const int per_row = d >> 2;
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long)B * L * per_row) return;
    const long row = i / per_row;
    const int c = (int)(i % per_row) * 4, l = (int)(row % L);
    const long tok = tokens[row];
    const float4 e = *reinterpret_cast<const float4*>(emb + tok * d + c);
    const float4 p = *reinterpret_cast<const float4*>(pe + (long)l * d + c);
    *reinterpret_cast<float4*>(out + row * d + c) =
        make_float4(e.x * scale + p.x, e.y * scale + p.y, e.z * scale + p.z, e.w * scale + p.w);
}
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ dout,
                                                        float* __restrict__ demb, long rows, int d, float scale) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * d) return;
    const long row = i / d;
    const int c = (int)(i % d);
    atomicAdd(demb + tokens[row] * d + c, scale * dout[i]);
}

// x[i] *= *scalar   (autograd's upstream gradient applied to a loss-gradient buffer without a host sync)
template <typename TD>
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(TD* __restrict__ x, long n, const float* __restrict__ scalar) {
    const float s = __ldg(scalar);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) x[i] = from_f32<TD>(to_f32<TD>(x[i]) * s);
}

}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
    LASR_REQUIRE(src && dst && n > 0, "cast: bad args");
    LASR_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "cast: unaligned");
    launch_pdl(cast_bf16_kernel, ceil_div(n, 1024), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n);
    return check_launch("cast_f32_bf16");
}

int lasr_permute4d(const void* src, int src_dtype, void* dst, int dst_dtype, const int64_t* n, const int64_t* src_strides,
                   const int64_t* dst_strides, int accumulate, void* stream) {
    LASR_REQUIRE(src && dst && n && src_strides && dst_strides, "permute4d: null pointer");
    Perm4 p;
    long total = 1;
    for (int i = 0; i < 4; ++i) { p.n[i] = n[i]; p.ss[i] = src_strides[i]; p.ds[i] = dst_strides[i]; total *= n[i]; }
    LASR_REQUIRE(total > 0, "permute4d: empty");
    int grid = ceil_div(total, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == LASR_F32 && dst_dtype == LASR_F32) permute4d_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, p, accumulate);
    else if (src_dtype == LASR_F32 && dst_dtype == LASR_BF16) permute4d_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, (bf16*)dst, p, accumulate);
    else if (src_dtype == LASR_BF16 && dst_dtype == LASR_F32) permute4d_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, (float*)dst, p, accumulate);
    else permute4d_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, (bf16*)dst, p, accumulate);
    return check_launch("permute4d");
}

int lasr_act_bwd(const void* da, int64_t ldda, const void* saved, int64_t lds, void* dh, int64_t lddh, float* dbias, int rows,
                 int cols, int act, float scale, int dtype, void* stream) {
    LASR_REQUIRE(da && rows > 0 && cols > 0, "act_bwd: bad args");
    LASR_REQUIRE(act == LASR_ACT_NONE || saved, "act_bwd: saved tensor required");
    LASR_REQUIRE(ldda % 4 == 0 && lds % 4 == 0 && lddh % 4 == 0, "act_bwd: strides must be multiples of 4");
    dim3 grid(ceil_div(cols, 256), ceil_div(rows, 64));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) launch_pdl(act_bwd_kernel<float>, grid, 256, 0, st, (const float*)da, ldda, (const float*)saved, lds, (float*)dh, lddh, dbias, rows, cols, act, scale);
    else if (dtype == LASR_BF16) launch_pdl(act_bwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)da, ldda, (const bf16*)saved, lds, (bf16*)dh, lddh, dbias, rows, cols, act, scale);
    else { set_error("act_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("act_bwd");
}

int lasr_pos_bias_fwd(const void* q, int64_t ldq, const float* u, const float* v, void* qu, void* qv, int64_t ldo, int rows,
                      int d, int dtype, void* stream) {
    LASR_REQUIRE(q && u && v && qu && qv && rows > 0 && d % 4 == 0 && ldq % 4 == 0 && ldo % 4 == 0, "pos_bias_fwd: bad args");
    const int grid = ceil_div((long)rows * (d / 4), 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_F32) launch_pdl(pos_bias_fwd_kernel<float>, grid, 256, 0, st, (const float*)q, ldq, u, v, (float*)qu, (float*)qv, ldo, rows, d);
    else if (dtype == LASR_BF16) launch_pdl(pos_bias_fwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)q, ldq, u, v, (bf16*)qu, (bf16*)qv, ldo, rows, d);
    else { set_error("pos_bias_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("pos_bias_fwd");
}

int lasr_pos_bias_bwd(const void* dqu, const void* dqv, int64_t ldi, void* dq, int64_t ldq, float* du, float* dv, float* dqbias,
                      int rows, int d, int dtype, void* stream) {
    LASR_REQUIRE(dqu && dqv && dq && du && dv && rows > 0 && d % 4 == 0 && ldi % 4 == 0 && ldq % 4 == 0, "pos_bias_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LASR_BF16 && d % 8 == 0 && ldi % 8 == 0 && ldq % 8 == 0 && (((uintptr_t)dqu | (uintptr_t)dqv | (uintptr_t)dq) & 15) == 0) {
        const int cx = ceil_div(d, 256);
        int per = ceil_div(rows, ceil_div(2 * 148, cx));  // about two CTAs per SM
        per = ceil_div(per < 32 ? 32 : per, 8) * 8;
        launch_pdl(pos_bias_bwd_wide_kernel, dim3(cx, ceil_div(rows, per)), 256, 0, st, (const bf16*)dqu, (const bf16*)dqv, ldi, (bf16*)dq, ldq, du, dv,
                   dqbias, rows, d, per);
        return check_launch("pos_bias_bwd");
    }
    dim3 grid(ceil_div(d, 256), ceil_div(rows, 64));
    if (dtype == LASR_F32) launch_pdl(pos_bias_bwd_kernel<float>, grid, 256, 0, st, (const float*)dqu, (const float*)dqv, ldi, (float*)dq, ldq, du, dv, dqbias, rows, d);
    else if (dtype == LASR_BF16) launch_pdl(pos_bias_bwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)dqu, (const bf16*)dqv, ldi, (bf16*)dq, ldq, du, dv, dqbias, rows, d);
    else { set_error("pos_bias_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("pos_bias_bwd");
}

int lasr_embed_fwd(const int64_t* tokens, int L, const float* emb, const float* pe, float* out, int B, int d, float scale,
                   void* stream) {
    LASR_REQUIRE(tokens && emb && pe && out && B > 0 && L > 0 && d % 4 == 0, "embed_fwd: bad args");
    const long n = (long)B * L * (d / 4);
    embed_fwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(tokens, L, emb, pe, out, B, d, scale);
    return check_launch("embed_fwd");
}

int lasr_embed_bwd(const int64_t* tokens, const float* dout, float* demb, int64_t rows, int d, float scale, void* stream) {
    LASR_REQUIRE(tokens && dout && demb && rows > 0 && d > 0, "embed_bwd: bad args");
    embed_bwd_kernel<<<ceil_div(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(tokens, dout, demb, rows, d, scale);
    return check_launch("embed_bwd");
}

int lasr_scale_by_scalar(void* x, int dtype, int64_t n, const float* scalar, void* stream) {
    LASR_REQUIRE(x && scalar && n > 0, "scale_by_scalar: bad args");
    int grid = ceil_div(n, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (dtype == LASR_F32) scale_by_scalar_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, n, scalar);
    else if (dtype == LASR_BF16) scale_by_scalar_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((bf16*)x, n, scalar);
    else { set_error("scale_by_scalar: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("scale_by_scalar");
}

}  // extern "C"
