// Fused backward of a position-wise feed-forward block's two inner GEMMs (nets/feed_forward.py:18-19 backward, the block the
// encoder runs twice per Conformer layer: nets/conformer_layer.py:37-47,58-66):
//
//     dh  = alpha * (dy . W2) * g          g = act'(fc1 pre-activation), saved by the forward pass (lasr_gemm aux_deriv);
//                                          0 where the inner dropout dropped the activation
//     db1 += colsum(dh)                    fc1's bias gradient
//     dln = dh . W1                        gradient wrt the block's LayerNorm output
//
// in ONE tcgen05 kernel.  Unfused, dh (rows x f, 154 MB at C2 / B = 126) is written by the first GEMM's epilogue and read back by
// the second GEMM through HBM (102 us + 43 us per block); here a 128-row tile of dh lives only as 64-column chunks that go
// TMEM -> registers (x g) -> a shared-memory slab that is at the same time the A operand of the second MMA and the source of
// the bulk tensor store of dh (which the fc1 weight-gradient GEMM still needs).  HBM sees dy and g read once, dh and dln written
// once.
//
// Persistent, one CTA per SM, one 128-row tile at a time, f walked in chunks of 64 columns:
//   warp 0       TMA producer: dy tile (resident for the tile), W2 / W1 chunk rings (2 stages each)
//   warp 1       MMA issuer:   acc1[chunk & 1] (128 x 64 fp32, TMEM) = dy . W2[:, chunk]      (K = d)
//                              acc2 (128 x d fp32, TMEM)            += dh_chunk . W1[chunk, :]  (K = 64)
//                              + the bulk tensor store of every finished dh slab
//   warps 2..17  epilogue:     acc1 -> x alpha x g -> bf16 slab (+ column sums), per tile acc2 -> bf16 -> dln
// d <= 256 (acc2 takes d of the 512 TMEM columns, acc1 2 x 64), d % 64 == 0, f % 64 == 0.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lasr {
namespace ffn {

constexpr int TM = 128;   // rows per tile (MMA M)
constexpr int CN = 64;    // f columns per chunk
constexpr int EPI_W = 16;
constexpr int THREADS = 64 + 32 * EPI_W;
constexpr int DMAX = 256;

constexpr int OFF_DY = 0;                        // d/64 k-blocks of 128 rows x 64 bf16 (16 KB each)
constexpr int OFF_W2 = OFF_DY + 4 * 16384;       // 2 stages x (d/64 boxes of 64 k x 64 n, 8 KB each)
constexpr int OFF_W1 = OFF_W2 + 2 * 32768;       // 2 stages x (d/64 boxes of 64 k x 64 n)
constexpr int OFF_DH = OFF_W1 + 2 * 32768;       // 2 slabs of 128 rows x 64 bf16
constexpr int OFF_BAR = OFF_DH + 2 * 16384;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared-memory budget");

enum { DY_FULL = 0, DY_EMPTY, W2_FULL, W2_EMPTY = W2_FULL + 2, W1_FULL = W2_EMPTY + 2, W1_EMPTY = W1_FULL + 2, ACC1_FULL = W1_EMPTY + 2,
       ACC1_EMPTY = ACC1_FULL + 2, DH_FULL = ACC1_EMPTY + 2, DH_EMPTY = DH_FULL + 2, ACC2_FULL = DH_EMPTY + 2, ACC2_EMPTY, NBARS };

struct Params {
    const bf16* g;   // (M, F) saved activation derivative
    long ldg;
    bf16* dln;       // (M, D)
    long lddln;
    float* colsum;   // (F) += column sums of dh, or nullptr
    float alpha;
    int M, D, F;
    int tiles;
};

__device__ __forceinline__ uint4 ldg_pred_u4(const void* p, bool pred) {
    uint4 v;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
        "@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// column sums of a warp's 32 rows x 16 columns: after the halving exchanges lane l holds the sum of column
// 8 b4 + 4 b3 + 2 b2 + b1 (b_i = bit i of l); lanes differing only in bit 0 hold the same column
__device__ __forceinline__ float col_sums_16(float (&v)[16], int lane) {
#pragma unroll
    for (int s = 16, n = 8; s >= 2; s >>= 1, n >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__global__ void __launch_bounds__(THREADS, 1)
ffn_bwd_kernel(const __grid_constant__ CUtensorMap m_dy, const __grid_constant__ CUtensorMap m_w2,
               const __grid_constant__ CUtensorMap m_w1, const __grid_constant__ CUtensorMap m_dh, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * NBARS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KBD = p.D >> 6;      // 64-wide blocks of d
    const int NC = p.F >> 6;       // chunks per tile

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_dy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_w2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_dh) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NBARS; ++i) {
            uint32_t cnt = 1;
            if (i == ACC1_EMPTY || i == ACC1_EMPTY + 1 || i == DH_FULL || i == DH_FULL + 1 || i == ACC2_EMPTY) cnt = EPI_W;
            if (i == DH_EMPTY || i == DH_EMPTY + 1) cnt = 2;  // the second MMA has read the slab AND its bulk store has
            mbar_init(bars + i, cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            uint32_t c = 0, t = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++t) {
                const int m0 = tile * TM;
                mbar_wait(bars + DY_EMPTY, (t & 1u) ^ 1u);
                mbar_arrive_expect_tx(bars + DY_FULL, (uint32_t)(KBD * 16384));
                for (int kb = 0; kb < KBD; ++kb) tma_load_4d(smem + OFF_DY + kb * 16384, &m_dy, bars + DY_FULL, kb * 64, m0, 0, 0);
                for (int j = 0; j < NC; ++j, ++c) {
                    const uint32_t s = c & 1u, ph = (c >> 1) & 1u;
                    const int f0 = j * CN;
                    mbar_wait(bars + W2_EMPTY + s, ph ^ 1u);
                    mbar_arrive_expect_tx(bars + W2_FULL + s, (uint32_t)(KBD * 8192));
                    for (int kb = 0; kb < KBD; ++kb)  // W2 (d, f) row-major = (K, N): box {64 n, 64 k}
                        tma_load_4d(smem + OFF_W2 + s * 32768 + kb * 8192, &m_w2, bars + W2_FULL + s, f0, kb * 64, 0, 0);
                    mbar_wait(bars + W1_EMPTY + s, ph ^ 1u);
                    mbar_arrive_expect_tx(bars + W1_FULL + s, (uint32_t)(KBD * 8192));
                    for (int nb = 0; nb < KBD; ++nb)  // W1 (f, d) row-major = (K, N): box {64 n, 64 k}
                        tma_load_4d(smem + OFF_W1 + s * 32768 + nb * 8192, &m_w1, bars + W1_FULL + s, nb * 64, f0, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptors: c = f32, a = b = bf16, A K-major, B MN-major, N >> 3, M >> 4
            const uint32_t id1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(CN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t id2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(p.D >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t s_dy = smem_u32(smem + OFF_DY), s_w2 = smem_u32(smem + OFF_W2), s_w1 = smem_u32(smem + OFF_W1),
                           s_dh = smem_u32(smem + OFF_DH);
            uint32_t c1 = 0, c2 = 0, t = 0;
            auto mma1 = [&]() {  // acc1[c1 & 1] = dy . W2[:, chunk]
                const uint32_t s = c1 & 1u, ph = (c1 >> 1) & 1u;
                mbar_wait(bars + W2_FULL + s, ph);
                mbar_wait(bars + ACC1_EMPTY + s, ph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < KBD; ++kb)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_bf16(tmem_base + s * CN, umma_desc(s_dy + kb * 16384 + kk * 32, 16, 1024),
                                    umma_desc(s_w2 + s * 32768 + kb * 8192 + kk * 2048, 8192, 1024), id1, (kb | kk) ? 1u : 0u);
                tc_commit(bars + W2_EMPTY + s);
                tc_commit(bars + ACC1_FULL + s);
                ++c1;
            };
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++t) {
                const int m0 = tile * TM;
                mbar_wait(bars + DY_FULL, t & 1u);
                tc_fence_after();
                mma1();
                for (int j = 0; j < NC; ++j, ++c2) {
                    if (j + 1 < NC) mma1();
                    else tc_commit(bars + DY_EMPTY);  // every MMA reading this tile's dy has been issued
                    const uint32_t s = c2 & 1u, ph = (c2 >> 1) & 1u;
                    mbar_wait(bars + W1_FULL + s, ph);
                    mbar_wait(bars + DH_FULL + s, ph);
                    if (j == 0) mbar_wait(bars + ACC2_EMPTY, (t & 1u) ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)  // acc2 += dh_chunk (A, K-major slab) . W1[chunk, :]
                        tc_mma_bf16(tmem_base + 2 * CN, umma_desc(s_dh + s * 16384 + kk * 32, 16, 1024),
                                    umma_desc(s_w1 + s * 32768 + kk * 2048, 8192, 1024), id2, (j | kk) ? 1u : 0u);
                    tc_commit(bars + W1_EMPTY + s);
                    tc_commit(bars + DH_EMPTY + s);
                    // dh goes to HBM straight from the MMA operand slab; the slab of the PREVIOUS chunk is free once its store
                    // has finished reading shared memory (all bulk groups but the newest)
                    tma_store_4d(&m_dh, smem + OFF_DH + s * 16384, j * CN, m0, 0, 0);
                    bulk_commit();
                    bulk_wait_read<1>();
                    if (c2 > 0) mbar_arrive(bars + DH_EMPTY + (s ^ 1u));
                }
                tc_commit(bars + ACC2_FULL);
            }
            bulk_wait_read<0>();
        }
    } else {
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int part = (warp - 2) >> 2;  // 16 of a chunk's 64 columns; 64 of dln's columns
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const float alpha = p.alpha;
        uint32_t c = 0, t = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++t) {
            const int row = tile * TM + r;
            const bool row_ok = row < p.M;
            const bf16* gp = p.g + (long)row * p.ldg + 16 * part;
            uint4 g0 = ldg_pred_u4(gp, row_ok), g1 = ldg_pred_u4(gp + 8, row_ok);
            for (int j = 0; j < NC; ++j, ++c) {
                const uint32_t s = c & 1u, ph = (c >> 1) & 1u;
                // next chunk's factors: in flight while this chunk is processed
                const bool more = j + 1 < NC;
                const uint4 n0 = ldg_pred_u4(gp + (j + 1) * CN, row_ok && more), n1 = ldg_pred_u4(gp + (j + 1) * CN + 8, row_ok && more);
                mbar_wait(bars + ACC1_FULL + s, ph);
                tc_fence_after();
                float v[16];
                tc_ld16(lane_addr + s * CN + 16 * part, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + ACC1_EMPTY + s);
                const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[2 * e] *= alpha * bf_lo(gw[e]);
                    v[2 * e + 1] *= alpha * bf_hi(gw[e]);
                }
                uint4 u0, u1;
                u0.x = pack2(v[0], v[1]); u0.y = pack2(v[2], v[3]); u0.z = pack2(v[4], v[5]); u0.w = pack2(v[6], v[7]);
                u1.x = pack2(v[8], v[9]); u1.y = pack2(v[10], v[11]); u1.z = pack2(v[12], v[13]); u1.w = pack2(v[14], v[15]);
                mbar_wait(bars + DH_EMPTY + s, ph ^ 1u);
                uint8_t* rowp = smem + OFF_DH + s * 16384 + r * 128;
                *reinterpret_cast<uint4*>(rowp + (((2 * part) ^ (r & 7)) << 4)) = u0;
                *reinterpret_cast<uint4*>(rowp + (((2 * part + 1) ^ (r & 7)) << 4)) = u1;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + DH_FULL + s);
                if (p.colsum) {  // fc1's bias gradient, from the fp32 values (before the bf16 rounding of the stored copy)
                    const float cs = col_sums_16(v, lane);
                    if ((lane & 1) == 0) atomicAdd(p.colsum + j * CN + 16 * part + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1), cs);
                }
                g0 = n0;
                g1 = n1;
            }
            // dln tile: acc2 -> bf16 -> global (this warp: its 32 rows x 64 columns)
            mbar_wait(bars + ACC2_FULL, t & 1u);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = part * 64 + i * 16;
                if (cc < p.D) {  // warp-uniform
                    float v[16];
                    tc_ld16(lane_addr + 2 * CN + cc, v);
                    if (row_ok) {
                        uint4* dst = reinterpret_cast<uint4*>(p.dln + (long)row * p.lddln + cc);
                        uint4 u;
                        u.x = pack2(v[0], v[1]); u.y = pack2(v[2], v[3]); u.z = pack2(v[4], v[5]); u.w = pack2(v[6], v[7]);
                        dst[0] = u;
                        u.x = pack2(v[8], v[9]); u.y = pack2(v[10], v[11]); u.z = pack2(v[12], v[13]); u.w = pack2(v[14], v[15]);
                        dst[1] = u;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + ACC2_EMPTY);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

// bf16 matrix (rows, inner) with row stride ld; box {64 inner, box_rows}, SWIZZLE_128B
static int make_map(CUtensorMap* map, const void* base, long inner, long rows, long ld, int box_rows, bool store) {
    auto enc = encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LASR_ERR_DRIVER; }
    const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, 1, 1};
    cuuint64_t strides[3] = {row_bytes, row_bytes, row_bytes};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_bytes & 15)) {
        set_error("ffn_bwd: operand base / row stride must be 16-byte aligned (base=%p ld=%ld)", base, ld);
        return LASR_ERR_BAD_ARG;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     store ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("ffn_bwd: cuTensorMapEncodeTiled failed (%d): inner=%ld rows=%ld ld=%ld", (int)r, inner, rows, ld);
        return LASR_ERR_DRIVER;
    }
    return LASR_OK;
}

}  // namespace ffn
}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_ffn_bwd_supported(int d, int f) { return (d >= 64 && d <= ffn::DMAX && d % 64 == 0 && f >= 64 && f % 64 == 0) ? 1 : 0; }

int lasr_ffn_bwd(const void* dy, int64_t lddy, const void* g, int64_t ldg, const void* w2, int64_t ldw2, const void* w1, int64_t ldw1,
                 void* dh, int64_t lddh, void* dln, int64_t lddln, float* colsum, float alpha, int M, int d, int f, void* stream) {
    LASR_REQUIRE(dy && g && w2 && w1 && dh && dln && M > 0, "ffn_bwd: null operand or empty problem");
    if (!lasr_ffn_bwd_supported(d, f)) {
        set_error("ffn_bwd: needs d %% 64 == 0, 64 <= d <= %d, f %% 64 == 0 (got d=%d f=%d)", ffn::DMAX, d, f);
        return LASR_ERR_UNSUPPORTED;
    }
    LASR_REQUIRE(ldg % 8 == 0 && lddln % 8 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (reinterpret_cast<uintptr_t>(dln) & 15) == 0,
                 "ffn_bwd: g and dln must be 16-byte aligned with row strides that are multiples of 8");
    CUtensorMap m_dy, m_w2, m_w1, m_dh;
    int rc;
    if ((rc = ffn::make_map(&m_dy, dy, d, M, lddy, ffn::TM, false)) != LASR_OK) return rc;
    if ((rc = ffn::make_map(&m_w2, w2, f, d, ldw2, 64, false)) != LASR_OK) return rc;   // (d, f): inner = f (N), rows = d (K)
    if ((rc = ffn::make_map(&m_w1, w1, d, f, ldw1, 64, false)) != LASR_OK) return rc;   // (f, d): inner = d (N), rows = f (K)
    if ((rc = ffn::make_map(&m_dh, dh, f, M, lddh, ffn::TM, true)) != LASR_OK) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ffn::ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn::SMEM_BYTES) != cudaSuccess)
            return check_launch("ffn_bwd smem attr");
        configured = true;
    }
    ffn::Params p;
    p.g = reinterpret_cast<const bf16*>(g); p.ldg = ldg;
    p.dln = reinterpret_cast<bf16*>(dln); p.lddln = lddln;
    p.colsum = colsum; p.alpha = alpha;
    p.M = M; p.D = d; p.F = f;
    p.tiles = (M + ffn::TM - 1) / ffn::TM;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int grid = p.tiles < sms ? p.tiles : sms;
    launch_pdl(ffn::ffn_bwd_kernel, dim3((unsigned)grid), dim3(ffn::THREADS), (size_t)ffn::SMEM_BYTES, (cudaStream_t)stream, m_dy, m_w2, m_w1,
               m_dh, p);
    return check_launch("ffn_bwd");
}

}  // extern "C"
