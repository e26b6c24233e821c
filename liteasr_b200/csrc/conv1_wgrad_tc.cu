// Weight / bias gradient of the first sub-sampling convolution (Conv2d(1 -> d, 3x3, stride 2), nets/subsampling.py:33) on the
// tensor cores:
//
//   dW[c][tap] = sum_{b, t1, f1} dh1[b][t1][f1][c] * x[b][2 t1 + kh][2 f1 + kw]          dbias[c] = sum dh1[b][t1][f1][c]
//
// is a GEMM with M = d channels, N = taps and K = every output position of the batch (2.9 M at C2/B = 126), whose only large
// operand is dh1 (1.5 GB of bf16, already in parity planes and already ReLU-masked by conv2's input-gradient epilogue).  The SIMT
// kernel it replaces spends 9 FMAs per gradient element and runs at 1.9 TB/s; here dh1 streams through TMA exactly once
// (MN-major A operand: the planes are (rows, channels) row-major, no transpose) and the 9 FMAs become one tcgen05.mma per 16 rows.
//
//   A (M = 128 channels x MH, K = 64 rows)  TMA boxes {64 channels, 64 rows} straight from the planes
//   B (N = 32, K = 64 rows)                 built in shared memory by four warps from x (fp32, L2-resident): column n < 9 = the
//                                           bf16 head of x at tap n, column 9 = 1 (-> dbias for free), column 16 + n = the bf16
//                                           tail x - head (so the product is exact to ~2^-17 although the MMA operands are bf16)
//   D (128 x 32 fp32 per channel half)      TMEM, accumulated over ALL the row blocks of the CTA; one atomicAdd per value at the end
//
//   warp 0 : TMA producer   warp 1 : MMA issuer   warps 2..5 : B builders, then the epilogue (one TMEM lane quarter each)
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lasr {
namespace c1w {

constexpr int KB = 64;       // rows per block
constexpr int NB = 32;       // B rows (taps: 9 heads, the ones column, 9 tails)
constexpr int B_BYTES = NB * 128;
constexpr int MAX_STAGES = 8;
constexpr int THREADS = 192;

struct Params {
    const float* x;
    float* dw;
    float* dbias;
    int B, T, F, T1, F1, U, V, d, mh;  // mh = d / 128
    int rblocks;                        // row blocks per plane
    long units;                         // B * 4 * rblocks
    int stages, a_bytes;
};

__global__ void __launch_bounds__(THREADS, 1) conv1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stage_bytes = p.a_bytes + B_BYTES;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* full_b = full_a + MAX_STAGES;
    uint64_t* empty = full_b + MAX_STAGES;
    uint64_t* done = empty + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = (p.mh * NB <= 32) ? 32u : (p.mh * NB <= 64 ? 64u : 128u);

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_a) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(full_a + s, 1);
            mbar_init(full_b + s, 4);  // one arrival per builder warp
            mbar_init(empty + s, 1);
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    LASR_PDL_SYNC();

    // contiguous share of the (utterance, plane, row block) list
    const long per = (p.units + gridDim.x - 1) / gridDim.x;
    const long u0 = (long)blockIdx.x * per, u1 = (u0 + per < p.units) ? u0 + per : p.units;
    const int PR = p.U * p.V;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long u = u0; u < u1; ++u) {
                const int rb = (int)(u % p.rblocks);
                const int bp = (int)(u / p.rblocks);  // utterance * 4 + plane
                mbar_wait(empty + s, ph ^ 1);
                mbar_arrive_expect_tx(full_a + s, (uint32_t)p.a_bytes);
                uint8_t* sa = smem + s * stage_bytes;
                for (int j = 0; j < p.d / 64; ++j) tma_load_4d(sa + j * 8192, &tma_a, full_a + s, 64 * j, rb * KB, bp, 0);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // c = f32, a = b = bf16, A MN-major, B K-major, N = 32, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (long u = u0; u < u1; ++u) {
                mbar_wait(full_a + s, ph);
                mbar_wait(full_b + s, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * stage_bytes), sb = sa + (uint32_t)p.a_bytes;
                for (int mh = 0; mh < p.mh; ++mh)
#pragma unroll
                    for (int kk = 0; kk < KB / 16; ++kk)
                        tc_mma_bf16(tmem_base + (uint32_t)(mh * NB), umma_desc(sa + mh * 16384 + kk * 2048, 8192, 1024),
                                    umma_desc(sb + kk * 32, 16, 1024), idesc, (u > u0 || kk > 0) ? 1u : 0u);
                tc_commit(empty + s);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            tc_commit(done);
        }
    } else {
        // ---- B builders: thread = (tap row n, 16-byte chunk ch of 8 rows)
        const int g = threadIdx.x - 64;
        const int n = g >> 3, ch = g & 7;
        const int kh = n / 3, kw = n - 3 * kh;
        int s = 0;
        uint32_t ph = 0;
        for (long u = u0; u < u1; ++u) {
            const int rb = (int)(u % p.rblocks);
            const int bp = (int)(u / p.rblocks);
            const int b = bp >> 2, pt = (bp >> 1) & 1, pf = bp & 1;
            uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
            if (n <= 9) {
                const float* xb = p.x + (long)b * p.T * p.F;
                const int r0 = rb * KB + ch * 8;
                int uu = r0 / p.V, v = r0 - uu * p.V;
                float xv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int t1 = 2 * uu + pt, f1 = 2 * v + pf;
                    const bool ok = (r0 + e < PR) && t1 < p.T1 && f1 < p.F1;  // slots without a (t1, f1) contribute nothing
                    if (n < 9) xv[e] = ok ? __ldg(xb + (long)(2 * t1 + kh) * p.F + 2 * f1 + kw) : 0.f;
                    else xv[e] = ok ? 1.f : 0.f;  // the bias-gradient column
                    if (++v == p.V) { v = 0; ++uu; }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(xv[2 * e], xv[2 * e + 1]);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(xv[2 * e] - __low2float(h2), xv[2 * e + 1] - __high2float(h2));
                    hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
                    lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            mbar_wait(empty + s, ph ^ 1);
            uint8_t* sb = smem + s * stage_bytes + p.a_bytes;
            // K-major SWIZZLE_128B: row n at n * 128, 16-byte chunk ch at (ch ^ (n & 7)); rows n and n + 16 share n & 7
            *reinterpret_cast<uint4*>(sb + n * 128 + ((ch ^ (n & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(sb + (n + 16) * 128 + ((ch ^ (n & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_b + s);
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        // ---- epilogue: lane = channel, columns = taps
        if (u1 > u0) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int q = warp & 3;
            for (int mh = 0; mh < p.mh; ++mh) {
                float v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * NB), v);
                const int c = mh * 128 + q * 32 + lane;
#pragma unroll
                for (int t = 0; t < 9; ++t) atomicAdd(p.dw + c * 9 + t, v[t] + v[16 + t]);
                atomicAdd(p.dbias + c, v[9]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

}  // namespace c1w

// returns LASR_ERR_UNSUPPORTED (without setting an error) when the shape is outside the kernel's range: the caller keeps the SIMT path
int conv1_wgrad_tc_dispatch(const float* x, const void* dh1p, float* dw, float* dbias, int B, int T, int F, int d, cudaStream_t st) {
    using namespace c1w;
    if (d % 128 != 0 || d > 512) return LASR_ERR_UNSUPPORTED;
    auto enc = encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(dh1p) & 15)) return LASR_ERR_UNSUPPORTED;
    Params p;
    p.x = x; p.dw = dw; p.dbias = dbias;
    p.B = B; p.T = T; p.F = F; p.d = d; p.mh = d / 128;
    p.T1 = (T - 3) / 2 + 1; p.F1 = (F - 3) / 2 + 1; p.U = (p.T1 + 1) / 2; p.V = (p.F1 + 1) / 2;
    const long PR = (long)p.U * p.V;
    p.rblocks = (int)((PR + KB - 1) / KB);
    p.units = (long)B * 4 * p.rblocks;
    p.a_bytes = d * KB * 2;
    const int stage_bytes = p.a_bytes + B_BYTES;
    p.stages = (232448 - 2048) / stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    if (p.stages < 2) return LASR_ERR_UNSUPPORTED;
    CUtensorMap ma;
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)PR, (cuuint64_t)B * 4, 1};
    cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)PR * d * 2, (cuuint64_t)PR * d * 2};
    cuuint32_t box[4] = {64, KB, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dh1p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return LASR_ERR_UNSUPPORTED;
    const int smem = p.stages * stage_bytes + 512 + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(conv1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess)
            return check_launch("conv1_wgrad_tc smem attr");
        configured = true;
    }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(p.units < sms ? p.units : sms);
    launch_pdl(conv1_wgrad_tc_kernel, dim3((unsigned)grid), dim3(THREADS), (size_t)smem, st, ma, p);
    return check_launch("conv1_wgrad_tc");
}

}  // namespace lasr
