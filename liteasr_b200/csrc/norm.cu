// LayerNorm forward / backward over the last dim (replaces nets/layer_norm.py:8-29, eps = 1e-12,
// and its autograd backward).  HBM-bound: one warp per row, the row lives in registers (float4
// loads), fp32 statistics, two-pass variance (mean first) like ATen's kernel.
//   fwd: y = (x - mean) * rstd * gamma + beta            (y in fp32 or bf16; mean/rstd saved)
//   bwd: dx (+)= rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//        dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy   (per-CTA partials -> red.add)
// The backward's "+=" mode implements the residual-stream gradient add of the pre-norm blocks
// (nets/conformer_layer.py:37-66) without an extra pass.
#include "common.cuh"
#include "philox.cuh"

namespace lasr {

constexpr int LN_MAXV = 8;  // float4 chunks per lane -> d <= 1024

template <typename TY, int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long ldx,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, TY* __restrict__ y, long ldy,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                            int d, float eps) {
    LASR_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    const int nchunk = d >> 2;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
            v[i] = *reinterpret_cast<const float4*>(xr + 4 * c);
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
            const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, e = v[i].w - mu;
            q += a * a + b * b + cc * cc + e * e;
        }
    }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
        if (mean) mean[row] = mu;
        if (rstd) rstd[row] = rs;
    }
    TY* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
            const float4 b = *reinterpret_cast<const float4*>(beta + 4 * c);
            const float o0 = (v[i].x - mu) * rs * g.x + b.x, o1 = (v[i].y - mu) * rs * g.y + b.y;
            const float o2 = (v[i].z - mu) * rs * g.z + b.z, o3 = (v[i].w - mu) * rs * g.w + b.w;
            if constexpr (sizeof(TY) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(yr) + 4 * c) = make_float4(o0, o1, o2, o3);
            } else {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
                uint2 u; u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(yr) + 4 * c) = u;
            }
        }
    }
}

template <typename TD>
__device__ __forceinline__ float4 load4(const TD* p) {
    if constexpr (sizeof(TD) == 4) {
        return *reinterpret_cast<const float4*>(p);
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    }
}

// NV = float4 chunks per lane (d <= 128 * NV).  Each warp walks rows with stride (grid * 8); the per-column partials
// (dgamma, dbeta and the optional column sum of the updated dx) stay in registers until the end of the kernel.
// Optional fused outputs for the pre-norm residual blocks (nets/conformer_layer.py:37-66 backward):
//   dx_lo  : bf16 copy of the final dx (the A operand of the next block's dgrad / wgrad GEMMs)
//   colsum : += cs_scale * sum_rows dx  (the bias gradient of the Linear whose output was added to this residual stream)
// Dropout (drop.thr != 0): the block that consumes dx_lo / colsum sits behind a dropout on its OUTPUT in the forward pass
// (x + drop(f(LN x)), nets/conformer_layer.py:37-66), so the gradient it must see is keep * scale * dx: the mask of that site is
// regenerated here (philox.cuh) and applied to dx_lo and to the column sums; dx itself (the residual-stream gradient) is not
// masked.  TLO = bf16 (tensor-core mode) or float (fp32 parity mode, where the masked copy is a second fp32 tensor).
template <typename TD, typename TLO, int NV>
__global__ void __launch_bounds__(256, (NV <= 2 ? 3 : (NV <= 4 ? 2 : 1))) layernorm_bwd_kernel(const TD* __restrict__ dy, long lddy,
                                                            const float* __restrict__ x, long ldx,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            long lddx, int accumulate, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int rows, int d,
                                                            TLO* __restrict__ dx_lo, long lddxlo,
                                                            float* __restrict__ colsum, float cs_scale, DropCfg drop) {
    LASR_PDL_SYNC();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = d >> 2;
    const bool drop_on = drop.thr != 0;
    DropKey dk = {};
    if (drop_on) dk = drop_key(drop);
    // gamma lives in shared memory, not in registers: at 3 CTAs per SM (80 registers) every register held across the row loop is a
    // spilled one; a conflict-free LDS.128 per chunk and row is cheaper than the local-memory traffic it replaces
    __shared__ float4 gam_s[NV][32];
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            gam_s[i][lane] = (c < nchunk) ? *reinterpret_cast<const float4*>(gamma + 4 * c) : make_float4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    float4 ag[NV], abt[NV], acs[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        ag[i] = make_float4(0, 0, 0, 0);
        abt[i] = make_float4(0, 0, 0, 0);
        acs[i] = make_float4(0, 0, 0, 0);
    }
    const long nwarps = (long)gridDim.x * 8;
    for (long row = (long)blockIdx.x * 8 + warp; row < rows; row += nwarps) {
        const float mu = mean[row], rs = rstd[row];
        const TD* dyr = dy + row * lddy;
        const float* xr = x + row * ldx;
        float* dxr = dx + row * lddx;
        float4 g[NV], xh[NV], old[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            old[i] = make_float4(0, 0, 0, 0);
            if (c < nchunk) {
                const float4 dv = load4<TD>(dyr + 4 * c);
                const float4 xv = *reinterpret_cast<const float4*>(xr + 4 * c);
                if (accumulate) old[i] = *reinterpret_cast<const float4*>(dxr + 4 * c);
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                const float4 gm = gam_s[i][lane];
                g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
                s1 += g[i].x + g[i].y + g[i].z + g[i].w;
                s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
                ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
                abt[i].x += dv.x; abt[i].y += dv.y; abt[i].z += dv.z; abt[i].w += dv.w;
            }
        }
        const float m1 = warp_sum(s1) / d, m2 = warp_sum(s2) / d;
        // keep bits of this lane's chunks.  One Philox call covers 16 columns = the chunks of a lane QUAD (chunk lane + 32 i is
        // nibble lane & 3 of group (lane >> 2) + 8 i), so the quad splits the calls: for chunk slots (i, i + 1) the even lanes
        // evaluate the group of slot i, the odd lanes the group of slot i + 1, and two shuffles hand every lane its nibbles.
        uint32_t k4[NV];
        if (drop_on) {
            if constexpr (NV >= 2) {
                const int q0 = lane & ~3, sh = 4 * (lane & 3);
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    const uint32_t mine = drop_keep16(dk, (uint32_t)row, (uint32_t)((lane >> 2) + 8 * (i + (lane & 1))));
                    k4[i] = (__shfl_sync(0xffffffffu, mine, q0) >> sh) & 15u;
                    k4[i + 1] = (__shfl_sync(0xffffffffu, mine, q0 + 1) >> sh) & 15u;
                }
            } else {
                k4[0] = drop_keep4(dk, (uint32_t)row, (uint32_t)(4 * lane));
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                const float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2) + old[i].x, rs * (g[i].y - m1 - xh[i].y * m2) + old[i].y,
                                             rs * (g[i].z - m1 - xh[i].z * m2) + old[i].z, rs * (g[i].w - m1 - xh[i].w * m2) + old[i].w);
                *reinterpret_cast<float4*>(dxr + 4 * c) = o;
                float4 om = o;
                if (drop_on) {
                    om.x = (k4[i] & 1u) ? o.x * dk.scale : 0.f; om.y = (k4[i] & 2u) ? o.y * dk.scale : 0.f;
                    om.z = (k4[i] & 4u) ? o.z * dk.scale : 0.f; om.w = (k4[i] & 8u) ? o.w * dk.scale : 0.f;
                }
                if (dx_lo) {
                    if constexpr (sizeof(TLO) == 4) {
                        *reinterpret_cast<float4*>(dx_lo + row * lddxlo + 4 * c) = om;
                    } else {
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(om.x, om.y), h1 = __floats2bfloat162_rn(om.z, om.w);
                        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                        *reinterpret_cast<uint2*>(dx_lo + row * lddxlo + 4 * c) = u;
                    }
                }
                acs[i].x += om.x; acs[i].y += om.y; acs[i].z += om.z; acs[i].w += om.w;
            }
        }
    }
    // cross-warp reduce of the per-lane column partials, then one red.add per column per CTA.  (Reducing across a thread-block
    // cluster of 8 CTAs through distributed shared memory first -- 8x fewer atomics on the same addresses -- measured much SLOWER:
    // 54 vs 32 us at C2 / B = 126; the cluster launch constrains where the 444 resident CTAs may run.)
    __shared__ float4 red[3][8][32];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (32 * i >= nchunk) break;  // uniform
        red[0][warp][lane] = ag[i];
        red[1][warp][lane] = abt[i];
        red[2][warp][lane] = acs[i];
        __syncthreads();
        if (warp < 3 && c < nchunk && (warp < 2 || colsum)) {
            float4 t = red[warp][0][lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) { t.x += red[warp][w][lane].x; t.y += red[warp][w][lane].y; t.z += red[warp][w][lane].z; t.w += red[warp][w][lane].w; }
            float* dst = (warp == 0 ? dgamma : (warp == 1 ? dbeta : colsum)) + 4 * c;
            if (warp == 2) { t.x *= cs_scale; t.y *= cs_scale; t.z *= cs_scale; t.w *= cs_scale; }
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
        }
        __syncthreads();
    }
}

// CTAs of 256 threads of `kern` that are resident at once on the device (cached per instantiation), capped by the row count
template <typename K>
static int resident_grid(K kern, int rows) {
    static int cap = 0;
    if (!cap) {
        int dev = 0, sms = 148, occ = 2;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ <= 0) occ = 2;
        cap = sms * occ;
    }
    const int need = ceil_div(rows, 8);
    return need < cap ? need : cap;
}

}  // namespace lasr

extern "C" {

int lasr_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int y_dtype,
                       int64_t ldy, float* mean, float* rstd, int rows, int d, float eps, void* stream) {
    using namespace lasr;
    LASR_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null pointer");
    LASR_REQUIRE(rows > 0 && d > 0 && d % 4 == 0 && d <= 128 * LN_MAXV, "layernorm_fwd: d=%d unsupported (d%%4==0, d<=1024)", d);
    LASR_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "layernorm_fwd: row strides must be multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(rows, 8);
#define LASR_LNF(TY, NV) launch_pdl(layernorm_fwd_kernel<TY, NV>, grid, 256, 0, st, x, ldx, gamma, beta, (TY*)y, ldy, mean, rstd, rows, d, eps)
#define LASR_LNF_D(TY)                      \
    do {                                    \
        if (d <= 128) LASR_LNF(TY, 1);      \
        else if (d <= 256) LASR_LNF(TY, 2); \
        else if (d <= 512) LASR_LNF(TY, 4); \
        else LASR_LNF(TY, 8);               \
    } while (0)
    if (y_dtype == LASR_F32) LASR_LNF_D(float);
    else if (y_dtype == LASR_BF16) LASR_LNF_D(bf16);
    else { set_error("layernorm_fwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
#undef LASR_LNF_D
#undef LASR_LNF
    return check_launch("layernorm_fwd");
}

int lasr_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                       const float* rstd, const float* gamma, float* dx, int64_t lddx, int accumulate, float* dgamma,
                       float* dbeta, int rows, int d, void* dx_lo, int64_t lddxlo, float* colsum, float colsum_scale,
                       void* stream) {
    return lasr_layernorm_bwd_drop(dy, dy_dtype, lddy, x, ldx, mean, rstd, gamma, dx, lddx, accumulate, dgamma, dbeta, rows, d, dx_lo,
                                   LASR_BF16, lddxlo, colsum, colsum_scale, nullptr, 0u, 0u, 1.f, stream);
}

int lasr_layernorm_bwd_drop(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                            const float* rstd, const float* gamma, float* dx, int64_t lddx, int accumulate, float* dgamma,
                            float* dbeta, int rows, int d, void* dx_lo, int lo_dtype, int64_t lddxlo, float* colsum,
                            float colsum_scale, const void* drop_state, uint32_t drop_site, uint32_t drop_thr, float drop_scale,
                            void* stream) {
    using namespace lasr;
    LASR_REQUIRE(drop_thr == 0 || (drop_state && drop_thr <= LASR_DROP_THR_MAX), "layernorm_bwd: dropout needs drop_state and thr <= 0x7c00");
    LASR_REQUIRE(lo_dtype == LASR_BF16 || lo_dtype == LASR_F32, "layernorm_bwd: bad dx_lo dtype");
    DropCfg drop;
    drop.state = (const unsigned long long*)drop_state; drop.site = drop_site; drop.thr = drop_thr; drop.scale = drop_scale;
    LASR_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
    LASR_REQUIRE(rows > 0 && d > 0 && d % 4 == 0 && d <= 128 * LN_MAXV, "layernorm_bwd: d=%d unsupported", d);
    LASR_REQUIRE(ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && lddxlo % 4 == 0, "layernorm_bwd: row strides must be multiples of 4");
    LASR_REQUIRE(((uintptr_t)dgamma & 15) == 0 && ((uintptr_t)dbeta & 15) == 0 && ((uintptr_t)colsum & 15) == 0,
                 "layernorm_bwd: dgamma/dbeta/colsum must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // the kernel walks the rows with a grid stride and keeps its column partials in registers: launch exactly the CTAs that are
    // resident at once (444 CTAs on 296 slots ran as one and a half waves, the second at half occupancy)
#define LASR_LNB(TD, NV)                                                                                                            \
    do {                                                                                                                            \
        if (lo_dtype == LASR_F32) {                                                                                                 \
            const int grid = resident_grid(layernorm_bwd_kernel<TD, float, NV>, rows);                                              \
            launch_pdl(layernorm_bwd_kernel<TD, float, NV>, grid, 256, 0, st, (const TD*)dy, lddy, x, ldx, mean, rstd, gamma, dx, lddx,  \
                       accumulate, dgamma, dbeta, rows, d, (float*)dx_lo, lddxlo, colsum, colsum_scale, drop);                      \
        } else {                                                                                                                    \
            const int grid = resident_grid(layernorm_bwd_kernel<TD, bf16, NV>, rows);                                               \
            launch_pdl(layernorm_bwd_kernel<TD, bf16, NV>, grid, 256, 0, st, (const TD*)dy, lddy, x, ldx, mean, rstd, gamma, dx, lddx,   \
                       accumulate, dgamma, dbeta, rows, d, (bf16*)dx_lo, lddxlo, colsum, colsum_scale, drop);                       \
        }                                                                                                                           \
    } while (0)
#define LASR_LNB_D(TD)                  \
    do {                                \
        if (d <= 128) LASR_LNB(TD, 1);  \
        else if (d <= 256) LASR_LNB(TD, 2); \
        else if (d <= 512) LASR_LNB(TD, 4); \
        else LASR_LNB(TD, 8);           \
    } while (0)
    if (dy_dtype == LASR_F32) LASR_LNB_D(float);
    else if (dy_dtype == LASR_BF16) LASR_LNB_D(bf16);
    else { set_error("layernorm_bwd: bad dtype"); return LASR_ERR_UNSUPPORTED; }
#undef LASR_LNB_D
#undef LASR_LNB
    return check_launch("layernorm_bwd");
}

}  // extern "C"
