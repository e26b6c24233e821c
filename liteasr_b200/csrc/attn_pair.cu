// Attention backward: the TWO contractions that consume one (B, H, Tq, Tk) gradient tensor in ONE tcgen05 kernel, so that the
// tensor crosses HBM once instead of twice (nets/attention.py:46-59,120-154 backward, per head):
//
//     dL[i, :] = sum_j X[i, j] R[j, :]        X = dS:  R = K,   dL = d(q + u)        X = dbd:  R = P (pos. projection), dL = d(q + v)
//     dR[j, :] = sum_i X[i, j] L[i, :]        X = dS:  L = q+u, dR = dK              X = dbd:  L = q + v,  dR = dP (summed over the batch)
//
// Unfused these are two batched lasr_gemm launches that each stream the 92 MB tensor (C2, per-GPU batch 126) at ~30 us; both are
// HBM-bound on X (K = 64 / 299 contractions at 100-190 TFLOP/s), so reading X once halves their cost.
//
// Work unit = (head, utterance); a CTA owns a contiguous range of units (head-major), one 128-row query tile of X at a time:
//   * the X tile lives in shared memory as 64-key boxes of 128 rows x 128 B (SWIZZLE_128B).  The SAME bytes are the K-major A
//     operand of the first contraction (M = 128 queries, K = keys) and the MN-major A operand of the second (M = 128 keys = two
//     boxes, K = 128 queries): only the UMMA descriptor differs (LBO = box pitch), nothing is transposed or copied;
//   * R (all keys of the unit, MN-major B of the first contraction) is loaded once per unit and double-buffered across units,
//     the L tile (MN-major B of the second contraction) per query tile;
//   * accumulators in tensor memory (all 512 columns): dL tile 2 x 64 columns, dR 2 x (3 x 64) columns (three 128-key M tiles), both
//     double-buffered against the epilogue; dR is accumulated over the unit's query tiles -- and, in the batch-reduced mode (dP), over the units of the
//     same head the CTA owns, so that the fp32 red.add traffic is one flush per (CTA, head) instead of one per utterance;
//   * box pair p of the X buffer is refilled for the next tile as soon as the second contraction has consumed it
//     (tcgen05.commit -> mbarrier), so loads of tile n + 1 overlap the MMAs of tile n without a second 80 KB buffer.
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue (TMEM -> bf16 rows / fp32 red.add, optional column sums).
// Needs dk == 64, Tk <= 320.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lasr {
namespace apair {

constexpr int TM = 128, DK = 64, MAXBOX = 5, MAXPAIR = 3;
constexpr int EPI_W = 8, THREADS = 64 + 32 * EPI_W;
constexpr int BOX = TM * 128;                        // one X box: 128 rows x 64 keys bf16 = 16 KB
constexpr int OFF_X = 0;                             // MAXBOX + 1 boxes (the last one stays zero: partner of an odd last box)
constexpr int OFF_R = OFF_X + (MAXBOX + 1) * BOX;    // 2 buffers x MAXBOX boxes of 64 keys x 128 B
constexpr int RBUF = MAXBOX * 8192;
constexpr int OFF_L = OFF_R + 2 * RBUF;              // 2 buffers x 128 rows x 128 B
constexpr int OFF_BAR = OFF_L + 2 * BOX;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared-memory budget");
constexpr int TM_DL = 0, TM_DR = 128, DR_COLS = MAXPAIR * DK;  // TMEM columns: dL 2 x 64, dR 2 x (3 x 64): both double-buffered against the epilogue

enum { R_FULL = 0, R_EMPTY = 2, L_FULL = 4, L_EMPTY = 6, X_FULL = 8, X_EMPTY = X_FULL + MAXPAIR, DL_FULL = X_EMPTY + MAXPAIR, DL_EMPTY = DL_FULL + 2,
       DR_FULL = DL_EMPTY + 2, DR_EMPTY = DR_FULL + 2, NBARS = DR_EMPTY + 2 };

struct Params {
    bf16* dl;        // (B * Tq, *) rows, head h at column h * 64
    long lddl;
    void* dr;        // bf16 (B * Tk, *) rows, or fp32 (Tk, *) accumulated over the batch (reduce_b)
    long lddr;
    float* colsum;   // (H * 64) += column sums of dR over (b, key), or nullptr
    int B, H, Tq, Tk;
    int r_batched;   // 0: R is one (Tk, *) matrix shared by every utterance (the positional projection)
    int reduce_b;    // 1: dR is fp32, summed over the batch with red.add
    int nbox, npair, ntile, units;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float col_sums_32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap m_x, const __grid_constant__ CUtensorMap m_r, const __grid_constant__ CUtensorMap m_l,
                 const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * NBARS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's contiguous range of units (unit = h * B + b: the units of a range share their head except at one boundary)
    const int u0 = (int)((long)blockIdx.x * p.units / gridDim.x), u1 = (int)((long)(blockIdx.x + 1) * p.units / gridDim.x);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_r) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&m_l) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NBARS; ++i) {
            const bool epi = (i == DL_EMPTY || i == DL_EMPTY + 1 || i == DR_EMPTY || i == DR_EMPTY + 1);
            mbar_init(bars + i, epi ? EPI_W : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && (p.nbox & 1)) {  // the zero partner of the odd last box (read by the second contraction's last M tile)
        uint4* z = reinterpret_cast<uint4*>(smem + OFF_X + p.nbox * BOX);
        for (int i = threadIdx.x - 64; i < BOX / 16; i += 32 * EPI_W) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;  // query-tile counter of this CTA
            for (int u = u0; u < u1; ++u) {
                const int h = u / p.B, b = u - h * p.B;
                const uint32_t ru = (uint32_t)(u - u0), rs = ru & 1u;
                mbar_wait(bars + R_EMPTY + rs, ((ru >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bars + R_FULL + rs, (uint32_t)(p.nbox * 8192));
                for (int i = 0; i < p.nbox; ++i)  // box {64 dk, 64 keys}: keys >= Tk are zero-filled
                    tma_load_4d(smem + OFF_R + rs * RBUF + i * 8192, &m_r, bars + R_FULL + rs, h * DK, i * 64, p.r_batched ? b : 0, 0);
                for (int t = 0; t < p.ntile; ++t, ++g) {
                    const uint32_t ls = g & 1u;
                    mbar_wait(bars + L_EMPTY + ls, ((g >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bars + L_FULL + ls, (uint32_t)BOX);
                    tma_load_4d(smem + OFF_L + ls * BOX, &m_l, bars + L_FULL + ls, h * DK, t * TM, b, 0);  // rows >= Tq: zeros
                    for (int pr = 0; pr < p.npair; ++pr) {
                        const int nb = min(2, p.nbox - 2 * pr);
                        mbar_wait(bars + X_EMPTY + pr, (g & 1u) ^ 1u);
                        mbar_arrive_expect_tx(bars + X_FULL + pr, (uint32_t)(nb * BOX));
                        for (int i = 0; i < nb; ++i)  // box {64 keys, 128 queries}: keys >= Tk and queries >= Tq are zero-filled
                            tma_load_4d(smem + OFF_X + (2 * pr + i) * BOX, &m_x, bars + X_FULL + pr, (2 * pr + i) * 64, t * TM, h, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // c = f32, a = b = bf16, N = 64, M = 128; first contraction: A K-major, B MN-major; second: A and B MN-major
            const uint32_t idc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DK >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t id1 = idc | (1u << 16), id2 = idc | (1u << 15) | (1u << 16);
            const uint64_t d_xk = umma_desc(smem_u32(smem + OFF_X), 16, 1024);       // K-major view of a box (+ 32 B per 16 keys)
            const uint64_t d_xm = umma_desc(smem_u32(smem + OFF_X), BOX, 1024);      // MN-major view of a box pair (LBO = box pitch)
            const uint64_t d_r = umma_desc(smem_u32(smem + OFF_R), 8192, 1024), d_l = umma_desc(smem_u32(smem + OFF_L), 8192, 1024);
            uint32_t g = 0, flushes = 0;
            bool fresh = true;  // the dR accumulators hold nothing yet (start, or just flushed)
            for (int u = u0; u < u1; ++u) {
                const int h = u / p.B;
                const uint32_t ru = (uint32_t)(u - u0), rs = ru & 1u;
                mbar_wait(bars + R_FULL + rs, (ru >> 1) & 1u);
                const uint32_t fb = flushes & 1u;  // dR accumulator buffer of this flush group
                if (fresh) mbar_wait(bars + DR_EMPTY + fb, ((flushes >> 1) & 1u) ^ 1u);  // the epilogue has read this buffer's previous flush
                for (int t = 0; t < p.ntile; ++t, ++g) {
                    const uint32_t ls = g & 1u, lph = (g >> 1) & 1u;
                    mbar_wait(bars + L_FULL + ls, lph);
                    mbar_wait(bars + DL_EMPTY + ls, lph ^ 1u);
                    tc_fence_after();
                    const uint32_t t_dl = tmem_base + TM_DL + ls * DK;
                    const uint64_t dl_b = d_l + (uint64_t)(ls * (BOX >> 4)), dr_b = d_r + (uint64_t)(rs * (RBUF >> 4));
                    for (int pr = 0; pr < p.npair; ++pr) {
                        const int nb = min(2, p.nbox - 2 * pr);
                        mbar_wait(bars + X_FULL + pr, g & 1u);
                        tc_fence_after();
                        for (int i = 0; i < nb; ++i) {  // dL tile += X[:, box] . R[box, :]
                            const int bx = 2 * pr + i;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_bf16(t_dl, d_xk + (uint64_t)((bx * BOX + kk * 32) >> 4), dr_b + (uint64_t)((bx * 8192 + kk * 2048) >> 4), id1,
                                            (bx | kk) ? 1u : 0u);
                        }
                        // dR[keys of the pair, :] += X[:, pair]^T . L tile
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            tc_mma_bf16(tmem_base + TM_DR + fb * DR_COLS + pr * DK, d_xm + (uint64_t)((2 * pr * BOX + kk * 2048) >> 4), dl_b + (uint64_t)((kk * 2048) >> 4), id2,
                                        (!fresh || t > 0 || kk > 0) ? 1u : 0u);
                        tc_commit(bars + X_EMPTY + pr);
                    }
                    tc_commit(bars + DL_FULL + ls);
                    tc_commit(bars + L_EMPTY + ls);
                }
                tc_commit(bars + R_EMPTY + rs);
                fresh = false;
                const bool flush = !p.reduce_b || u + 1 == u1 || (u + 1) / p.B != h;
                if (flush) {
                    tc_commit(bars + DR_FULL + fb);
                    ++flushes;
                    fresh = true;
                }
            }
        }
    } else {
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;  // TMEM lane quarter of this warp; 32 of the 64 columns
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t g = 0, flushes = 0;
        float cs_acc = 0.f;  // this lane's column sum of dR over the units of the current head (one atomic per warp, head and CTA:
                             // 1512 same-address atomics per launch and column otherwise, ~27 cycles each in the L2)
        for (int u = u0; u < u1; ++u) {
            const int h = u / p.B, b = u - h * p.B;
            for (int t = 0; t < p.ntile; ++t, ++g) {
                const uint32_t ls = g & 1u;
                mbar_wait(bars + DL_FULL + ls, (g >> 1) & 1u);
                tc_fence_after();
                float v[32];
                tc_ld32(lane_addr + TM_DL + ls * DK + 32 * half, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + DL_EMPTY + ls);
                const int row = t * TM + r;
                if (row < p.Tq) {
                    uint4* dst = reinterpret_cast<uint4*>(p.dl + ((long)b * p.Tq + row) * p.lddl + h * DK + 32 * half);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 w;
                        w.x = pack2(v[8 * i], v[8 * i + 1]); w.y = pack2(v[8 * i + 2], v[8 * i + 3]);
                        w.z = pack2(v[8 * i + 4], v[8 * i + 5]); w.w = pack2(v[8 * i + 6], v[8 * i + 7]);
                        dst[i] = w;
                    }
                }
            }
            const bool flush = !p.reduce_b || u + 1 == u1 || (u + 1) / p.B != h;
            if (!flush) continue;
            const bool last_of_head = u + 1 == u1 || (u + 1) / p.B != h;
            const uint32_t fb = flushes & 1u;
            mbar_wait(bars + DR_FULL + fb, (flushes >> 1) & 1u);
            ++flushes;
            tc_fence_after();
            for (int pr = 0; pr < p.npair; ++pr) {
                float v[32];
                tc_ld32(lane_addr + TM_DR + fb * DR_COLS + pr * DK + 32 * half, v);
                const int key = pr * TM + r;
                if (key < p.Tk) {
                    if (p.reduce_b) {
                        float* dst = reinterpret_cast<float*>(p.dr) + (long)key * p.lddr + h * DK + 32 * half;
#pragma unroll
                        for (int i = 0; i < 8; ++i) red_add_f32x4(dst + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                    } else {
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.dr) + ((long)b * p.Tk + key) * p.lddr + h * DK + 32 * half);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 w;
                            w.x = pack2(v[8 * i], v[8 * i + 1]); w.y = pack2(v[8 * i + 2], v[8 * i + 3]);
                            w.z = pack2(v[8 * i + 4], v[8 * i + 5]); w.w = pack2(v[8 * i + 6], v[8 * i + 7]);
                            dst[i] = w;
                        }
                    }
                }
                if (p.colsum) cs_acc += col_sums_32(v, lane);  // bias gradient of the projection that produced R: keys >= Tk hold exact zeros
            }
            if (p.colsum && last_of_head) {
                atomicAdd(p.colsum + h * DK + 32 * half + lane, cs_acc);
                cs_acc = 0.f;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + DR_EMPTY + fb);
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

// bf16 tensor, up to 4-D, dims / strides innermost first (strides in elements, for dims 1..3); SWIZZLE_128B, box {64, rows, 1, 1}
static int make_map(CUtensorMap* map, const void* base, const long dims_[4], const long strides_[3], int box_rows, const char* what) {
    auto enc = encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return LASR_ERR_DRIVER; }
    cuuint64_t dims[4], strides[3];
    for (int i = 0; i < 4; ++i) dims[i] = (cuuint64_t)(dims_[i] > 0 ? dims_[i] : 1);
    for (int i = 0; i < 3; ++i) strides[i] = (cuuint64_t)strides_[i] * 2;
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        set_error("attn_bwd_pair: %s base / strides must be 16-byte aligned", what);
        return LASR_ERR_BAD_ARG;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("attn_bwd_pair: cuTensorMapEncodeTiled failed (%d) for %s", (int)r, what);
        return LASR_ERR_DRIVER;
    }
    return LASR_OK;
}

}  // namespace apair
}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_attn_bwd_pair_supported(int Tk, int dk) { return (dk == apair::DK && Tk >= 1 && Tk <= 64 * apair::MAXBOX) ? 1 : 0; }

int lasr_attn_bwd_pair(const void* x, int64_t ldx, const void* r, int64_t ldr, int r_batched, const void* l, int64_t ldl, void* dl, int64_t lddl,
                       void* dr, int64_t lddr, int reduce_b, float* colsum, int B, int H, int Tq, int Tk, int dk, void* stream) {
    LASR_REQUIRE(x && r && l && dl && dr && B > 0 && H > 0 && Tq > 0 && Tk > 0, "attn_bwd_pair: null operand or empty problem");
    if (!lasr_attn_bwd_pair_supported(Tk, dk)) {
        set_error("attn_bwd_pair: needs dk == 64 and Tk <= %d (got dk=%d Tk=%d)", 64 * apair::MAXBOX, dk, Tk);
        return LASR_ERR_UNSUPPORTED;
    }
    LASR_REQUIRE(ldx >= Tk && ldx % 8 == 0 && ldr % 8 == 0 && ldl % 8 == 0 && lddl % 8 == 0 && (reinterpret_cast<uintptr_t>(dl) & 15) == 0,
                 "attn_bwd_pair: row strides must be multiples of 8 elements, dl 16-byte aligned");
    LASR_REQUIRE(reduce_b ? (lddr % 4 == 0 && (reinterpret_cast<uintptr_t>(dr) & 15) == 0) : (lddr % 8 == 0 && (reinterpret_cast<uintptr_t>(dr) & 15) == 0),
                 "attn_bwd_pair: dr must be 16-byte aligned with an aligned row stride");
    LASR_REQUIRE(!colsum || !reduce_b, "attn_bwd_pair: column sums are for the per-utterance mode");
    CUtensorMap m_x, m_r, m_l;
    int rc;
    {   // X (B, H, Tq, ld): dims (Tk, Tq, H, B) -- the inner extent is Tk, so the padding columns [Tk, ld) read as zeros
        const long dims[4] = {Tk, Tq, H, B}, st[3] = {ldx, (long)Tq * ldx, (long)H * Tq * ldx};
        if ((rc = apair::make_map(&m_x, x, dims, st, apair::TM, "X")) != LASR_OK) return rc;
    }
    {   // R: (B * Tk, *) rows (or (Tk, *)), head h at column 64 h: dims (64 H, Tk, B)
        const long dims[4] = {64L * H, Tk, r_batched ? B : 1, 1}, st[3] = {ldr, (long)Tk * ldr, (long)Tk * ldr};
        if ((rc = apair::make_map(&m_r, r, dims, st, 64, "R")) != LASR_OK) return rc;
    }
    {   // L: (B * Tq, *) rows: dims (64 H, Tq, B)
        const long dims[4] = {64L * H, Tq, B, 1}, st[3] = {ldl, (long)Tq * ldl, (long)Tq * ldl};
        if ((rc = apair::make_map(&m_l, l, dims, st, apair::TM, "L")) != LASR_OK) return rc;
    }
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(apair::attn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, apair::SMEM_BYTES) != cudaSuccess)
            return check_launch("attn_bwd_pair smem attr");
        configured = true;
    }
    apair::Params p;
    p.dl = reinterpret_cast<bf16*>(dl); p.lddl = lddl;
    p.dr = dr; p.lddr = lddr;
    p.colsum = colsum;
    p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk;
    p.r_batched = r_batched; p.reduce_b = reduce_b;
    p.nbox = (Tk + 63) / 64;
    p.npair = (p.nbox + 1) / 2;
    p.ntile = (Tq + apair::TM - 1) / apair::TM;
    p.units = B * H;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int grid = p.units < sms ? p.units : sms;
    launch_pdl(apair::attn_pair_kernel, dim3((unsigned)grid), dim3(apair::THREADS), (size_t)apair::SMEM_BYTES, (cudaStream_t)stream, m_x, m_r, m_l, p);
    return check_launch("attn_bwd_pair");
}

}  // extern "C"
