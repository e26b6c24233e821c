// Weight-gradient GEMM on CTA PAIRS (tcgen05.mma.cta_group::2):   C (M x N, fp32) += alpha * A^T . B
//     A : (K, M) row-major bf16  (dy: rows = the K = batch * time reduction dimension)       -> MN-major A operand
//     B : (K, N) row-major bf16  (the layer input x)                                          -> MN-major B operand
// i.e. dW += dy^T x of every Linear (trainer-side: nets/feed_forward.py:18-19, nets/attention.py:35-37 ... backward).
//
// Why a second kernel: the single-CTA kernel (gemm_tc.cu) runs these shapes at ~900 TFLOP/s in isolation, 54 % of the tensor pipe in
// the step -- it stages 48 KB of operands (A 128 x 64, B 256 x 64) per 512 cycles of MMA, and ~50 B/clk per SM is what the L2 ->
// shared-memory path delivers.  A CTA pair computes a 256 x 256 tile with each SM staging only ITS half of both operands (A: its
// 128 rows of M, B: 128 of the 256 columns of N; the tensor cores read the partner's half of B through the pair): 32 KB per SM for
// the same 512 cycles of MMA.
//
// One 256 x 256 x K/S work unit per cluster of 2 CTAs (non-persistent; M / 256 x N / 256 x S units, S = split-K), per CTA:
//   warp 0  TMA producer: 2 + 2 boxes {64 mn, 64 k} per stage, cp.async.bulk.tensor ... .cta_group::2 -- the transaction bytes of BOTH
//           CTAs land on the LEADER's full barrier (the barrier address is mapped to CTA 0 of the pair)
//   warp 1  MMA issuer (leader CTA only): tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16; tcgen05.commit ... multicast releases
//           the stage in both CTAs; the last commit signals the accumulator barrier of both
//   warps 2..9 epilogue (both CTAs, each its own 128 accumulator rows): TMEM -> swizzled 32 x 32 fp32 staging box -> bulk tensor
//           reduce-add into C
// Requires M % 256 == 0, N % 256 == 0 (every projection of the d = 256 / 512 models); other shapes stay on gemm_tc.cu.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lasr {
namespace g2 {

constexpr int STAGES = 6, STAGE_BYTES = 32768;   // per CTA: A 2 boxes + B 2 boxes of 64 x 64 bf16
constexpr int EPI_W = 8, THREADS = 64 + 32 * EPI_W;
constexpr int OFF_RING = 0;
constexpr int OFF_STAGE = OFF_RING + STAGES * STAGE_BYTES;   // EPI_W x 4 KB staging boxes
constexpr int OFF_BAR = OFF_STAGE + EPI_W * 4096;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared-memory budget");
constexpr int ACC_COLS = 256;

struct Params {
    int K, tiles_m, tiles_n, splits;
    float alpha;
    // implicit-GEMM weight gradient of the 3x3 stride-2 convolution over parity planes (gemm_tc.cu::conv2_tc_dispatch, mode 3): the
    // reduction dimension runs over (utterance, padded output row), the 256-wide N tile selects the tap = (row shift, parity plane)
    int conv;          // 0 plain | 1 convolution
    int kbpb;          // k-blocks per utterance
    int conv_d;        // channels
    int conv_off[9], conv_plane[9];
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// TMA load of this CTA's share of a pair's operand tile; the bytes are counted on the barrier `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void commit2_both(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap m_a, const __grid_constant__ CUtensorMap m_b, const __grid_constant__ CUtensorMap m_c,
              const Params p) {
    extern __shared__ uint8_t smem_raw[];
    // the same offset in both CTAs (the pair addresses its partner's memory by offset): the dynamic window starts at the same
    // shared::cta address in every CTA of a launch, so aligning by the address is the same adjustment in both
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int unit = blockIdx.x >> 1;
    const int nt = unit % p.tiles_n, mt = (unit / p.tiles_n) % p.tiles_m, split = unit / (p.tiles_n * p.tiles_m);
    const int total_kb = p.conv ? p.K * p.kbpb : (p.K + 63) / 64, per = (total_kb + p.splits - 1) / p.splits;  // conv: K = utterances
    const int kb0 = split * per, nkb = min(total_kb, kb0 + per) - kb0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before either touches the partner's
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (nkb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const int m0 = mt * 256 + (int)rank * 128;
                int n0 = nt * 256 + (int)rank * 128, boff = 0, bplane = 0;
                if (p.conv) {  // the N tile lies inside one tap: columns [tap * d, (tap + 1) * d) of the (co, tap, ci) gradient
                    const int tp = (nt * 256) / p.conv_d;
                    n0 -= tp * p.conv_d;
                    boff = p.conv_off[tp];
                    bplane = p.conv_plane[tp];
                }
                uint32_t s = 0, ph = 0;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(empty_bar + s, ph ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(full_bar + s, (uint32_t)(2 * STAGE_BYTES));  // both CTAs' bytes
                    const uint32_t bar = map_to_rank(smem_u32(full_bar + s), 0);
                    uint8_t* st = smem + OFF_RING + s * STAGE_BYTES;
                    int k0 = (kb0 + i) * 64, bi = 0;
                    if (p.conv) { bi = (kb0 + i) / p.kbpb; k0 = (kb0 + i - bi * p.kbpb) * 64; }  // rows >= YR of dY read as zeros
                    tma_load_2sm(st, &m_a, bar, m0, k0, 0, bi);               // A: box {64 m, 64 k} x 2
                    tma_load_2sm(st + 8192, &m_a, bar, m0 + 64, k0, 0, bi);
                    tma_load_2sm(st + 16384, &m_b, bar, n0, k0 + boff, bplane, bi);       // B: this CTA's 128 of the tile's 256 columns
                    tma_load_2sm(st + 24576, &m_b, bar, n0 + 64, k0 + boff, bplane, bi);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0 && rank == 0) {
                // c = f32, a = b = bf16, A and B MN-major, N = 256, M = 256 (128 rows per CTA)
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                uint32_t s = 0, ph = 0;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(full_bar + s, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + OFF_RING + s * STAGE_BYTES), sb = sa + 16384;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        mma2_bf16(tmem_base, umma_desc(sa + kk * 2048, 8192, 1024), umma_desc(sb + kk * 2048, 8192, 1024), idesc, (i | kk) ? 1u : 0u);
                    commit2_both(empty_bar + s);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                commit2_both(acc_full);
            }
        } else {
            const int ew = warp - 2, q = warp & 3, half = ew >> 2;  // TMEM lane quarter; 128 of the 256 columns
            uint8_t* stage = smem + OFF_STAGE + ew * 4096;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const int row0 = mt * 256 + (int)rank * 128 + q * 32;
            mbar_wait(acc_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int cc = half * 128 + c * 32;
                float v[32];
                tc_ld32(lane_addr + (uint32_t)cc, v);
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(p.alpha * v[4 * j], p.alpha * v[4 * j + 1], p.alpha * v[4 * j + 2], p.alpha * v[4 * j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_reduce_add_2d(&m_c, stage, nt * 256 + cc, row0);
                    bulk_commit();
                }
            }
            if (lane == 0) bulk_wait_read<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA frees tensor memory (or exits) while its partner may still read its shared memory
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ACC_COLS) : "memory");
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

// bf16 operand, 4-D (dims / strides innermost first, strides in elements for dims 1..3), box {64, 64, 1, 1}, SWIZZLE_128B
static int make_map4(CUtensorMap* map, const void* base, const long dims_[4], const long strides_[3]) {
    auto enc = encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LASR_ERR_DRIVER; }
    cuuint64_t dims[4], strides[3];
    for (int i = 0; i < 4; ++i) dims[i] = (cuuint64_t)(dims_[i] > 0 ? dims_[i] : 1);
    for (int i = 0; i < 3; ++i) strides[i] = (cuuint64_t)strides_[i] * 2;
    cuuint32_t box[4] = {64, 64, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        set_error("wgrad2: operand base / strides must be 16-byte aligned");
        return LASR_ERR_BAD_ARG;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad2: cuTensorMapEncodeTiled failed (%d)", (int)r); return LASR_ERR_DRIVER; }
    return LASR_OK;
}

static int make_map2(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int es, long inner, long rows, long ld, int box_inner, int box_rows,
                     CUtensorMapSwizzle sw) {
    auto enc = encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LASR_ERR_DRIVER; }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15)) { set_error("wgrad2: operand base / row stride must be 16-byte aligned"); return LASR_ERR_BAD_ARG; }
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad2: cuTensorMapEncodeTiled failed (%d)", (int)r); return LASR_ERR_DRIVER; }
    return LASR_OK;
}

}  // namespace g2
}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_wgrad2_supported(int m, int n) { return (m > 0 && n > 0 && m % 256 == 0 && n % 256 == 0) ? 1 : 0; }

int lasr_wgrad2(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc, float alpha, int m, int n, int k, int split_k,
                void* stream) {
    LASR_REQUIRE(a && b && c && k > 0, "wgrad2: null operand or empty problem");
    if (!lasr_wgrad2_supported(m, n)) { set_error("wgrad2: needs M %% 256 == 0 and N %% 256 == 0 (got %d x %d)", m, n); return LASR_ERR_UNSUPPORTED; }
    CUtensorMap m_a, m_b, m_c;
    int rc;
    {
        const long da[4] = {m, k, 1, 1}, sa[3] = {lda, lda, lda}, db[4] = {n, k, 1, 1}, sb[3] = {ldb, ldb, ldb};
        if ((rc = g2::make_map4(&m_a, a, da, sa)) != LASR_OK) return rc;
        if ((rc = g2::make_map4(&m_b, b, db, sb)) != LASR_OK) return rc;
    }
    if ((rc = g2::make_map2(&m_c, c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, n, m, ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) != LASR_OK) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(g2::wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::SMEM_BYTES) != cudaSuccess)
            return check_launch("wgrad2 smem attr");
        configured = true;
    }
    g2::Params p;
    memset(&p, 0, sizeof(p));
    p.K = k;
    p.tiles_m = m / 256;
    p.tiles_n = n / 256;
    const int total_kb = (k + 63) / 64;
    int s = split_k < 1 ? 1 : split_k;
    if (s > total_kb) s = total_kb;
    p.splits = s;
    p.alpha = alpha;
    const long units = (long)p.tiles_m * p.tiles_n * s;
    launch_pdl(g2::wgrad2_kernel, dim3((unsigned)(2 * units)), dim3(g2::THREADS), (size_t)g2::SMEM_BYTES, (cudaStream_t)stream, m_a, m_b, m_c, p);
    return check_launch("wgrad2");
}

}  // extern "C"

namespace lasr {
// Weight gradient of the sub-sampling front end's second convolution (nets/subsampling.py:33-34 backward) on CTA pairs: the same
// implicit GEMM as gemm_tc.cu::conv2_tc_dispatch(mode 3) -- dW[co][tap][ci] += sum_b sum_r dY[b][r][co] h1p[b][plane][r + off][ci] --
// for d % 256 == 0.  Returns LASR_ERR_UNSUPPORTED for other d (the caller keeps the single-CTA kernel).
int conv2_wgrad2_dispatch(const void* h1p, const void* dy, float* out, int B, int U, int V, int T2, int d, cudaStream_t st) {
    if (d % 256 != 0 || B < 1 || U < 1 || V < 1 || T2 < 1) return LASR_ERR_UNSUPPORTED;
    const long PR = (long)U * V, YR = (long)T2 * V;
    CUtensorMap m_a, m_b, m_c;
    int rc;
    {
        const long da[4] = {d, YR, 1, B}, sa[3] = {d, YR * d, YR * d};
        const long db[4] = {d, PR, 4, B}, sb[3] = {d, PR * d, 4 * PR * d};
        if ((rc = g2::make_map4(&m_a, dy, da, sa)) != LASR_OK) return rc;
        if ((rc = g2::make_map4(&m_b, h1p, db, sb)) != LASR_OK) return rc;
    }
    if ((rc = g2::make_map2(&m_c, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 9L * d, d, 9L * d, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) != LASR_OK) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(g2::wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::SMEM_BYTES) != cudaSuccess)
            return check_launch("conv2_wgrad2 smem attr");
        configured = true;
    }
    g2::Params p;
    memset(&p, 0, sizeof(p));
    p.conv = 1;
    p.K = B;
    p.kbpb = (int)((YR + 63) / 64);
    p.conv_d = d;
    for (int kh = 0, t = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw, ++t) {
            p.conv_off[t] = (kh >> 1) * V + (kw >> 1);
            p.conv_plane[t] = (kh & 1) * 2 + (kw & 1);
        }
    p.tiles_m = d / 256;
    p.tiles_n = 9 * d / 256;
    const int tiles = p.tiles_m * p.tiles_n;
    const long total_kb = (long)B * p.kbpb;
    long s = tiles <= 74 ? 74 / tiles : 1;
    if (s > total_kb) s = total_kb;
    p.splits = (int)(s < 1 ? 1 : s);
    p.alpha = 1.f;
    launch_pdl(g2::wgrad2_kernel, dim3((unsigned)(2L * tiles * p.splits)), dim3(g2::THREADS), (size_t)g2::SMEM_BYTES, st, m_a, m_b, m_c, p);
    return check_launch("conv2_wgrad2");
}
}  // namespace lasr
