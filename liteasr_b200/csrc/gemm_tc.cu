// bf16 GEMM on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, TMA-fed smem ring.
//
//   C[b1,b2] (M,N) = alpha * act(A . B^T + bias) (+ res)        (or C += alpha * A.B^T with red.add)
//
// One 128 x BN output tile per CTA, K streamed in 64-wide slabs through a STAGES-deep
// TMA -> mbarrier -> tcgen05.mma pipeline (warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue, one per TMEM lane quarter).  Both operands may be K-major (row-major
// (rows,K)) or MN-major (row-major (K,rows)); the difference is only in the TMA box shape, the
// UMMA shared-memory descriptor (LBO/SBO) and the instruction-descriptor major bits, so backward
// GEMMs (dgrad: B MN-major, wgrad: A and B MN-major) need no transposed copies.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace lasr {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_THREADS = 192;
constexpr int UMMA_K = 16;

struct TcParams {
    void* c;
    const float* bias;
    const float* res;
    void* aux;
    int m, n, k;
    int c_dtype;
    long ldc, ldres;
    int batch2;
    long sc1, sc2;
    int a_b1, a_b2, b_b1, b_b2;  // 1 if the operand really advances along that batch level
    float alpha;
    int act;
    int accumulate;
    int split_k;
    int vec_ok;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > LASR_DEVICE_TIMEOUT_CYCLES) __trap();  // never hang the box
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SW128)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN>
struct TileCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <typename CT>
__device__ __forceinline__ void epilogue_store(const TcParams& p, float* v, long row_off, int m, int nbase) {
    // v[0..31] are accumulators for columns nbase..nbase+31 of row m
    CT* crow = reinterpret_cast<CT*>(p.c) + row_off;
    CT* arow = p.aux ? reinterpret_cast<CT*>(p.aux) + row_off : nullptr;
    const int nvalid = min(32, p.n - nbase);
    if (p.accumulate) {
        float* cf = reinterpret_cast<float*>(p.c) + row_off;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) atomicAdd(cf + nbase + j, p.alpha * v[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = v[j];
        if (p.bias && j < nvalid) x += __ldg(p.bias + nbase + j);
        v[j] = x;
    }
    if (arow) {
        if (p.vec_ok && nvalid == 32) {
            if constexpr (sizeof(CT) == 4) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(arow) + nbase + j) =
                        make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(arow) + nbase + j) = u;
                }
            }
        } else {
            for (int j = 0; j < nvalid; ++j) arow[nbase + j] = from_f32<CT>(v[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = p.alpha * apply_act(v[j], p.act);
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const TcParams p) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int split = blockIdx.z % p.split_k, batch = blockIdx.z / p.split_k;
    const int b2 = batch % p.batch2, b1 = batch / p.batch2;
    const int total_kb = (p.k + BK - 1) / BK;
    const int kb_per = (total_kb + p.split_k - 1) / p.split_k;
    const int kb_begin = split * kb_per;
    const int num_kb = min(total_kb, kb_begin + kb_per) - kb_begin;
    if (num_kb <= 0) return;  // uniform over the CTA (only possible for trailing splits)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            const int ab1 = b1 * p.a_b1, ab2 = b2 * p.a_b2, bb1 = b1 * p.b_b1, bb2 = b2 * p.b_b2;
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_arrive_expect_tx(full_bar + s, Cfg::STAGE_BYTES);
                uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                const int k0 = (kb_begin + i) * BK;
                if constexpr (!A_MN) {
                    tma_load_4d(sa, &tma_a, full_bar + s, k0, m0, ab2, ab1);  // box {64 k, 128 m}
                } else {
#pragma unroll
                    for (int j = 0; j < BM / 64; ++j)  // box {64 m, 64 k}
                        tma_load_4d(sa + j * 8192, &tma_a, full_bar + s, m0 + 64 * j, k0, ab2, ab1);
                }
                if constexpr (!B_MN) {
                    tma_load_4d(sb, &tma_b, full_bar + s, k0, n0, bb2, bb1);  // box {64 k, BN n}
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_4d(sb + j * 8192, &tma_b, full_bar + s, n0 + 64 * j, k0, bb2, bb1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a=b=bf16, majors, N>>3, M>>4
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                       ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(full_bar + s, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    // K-major: +32 B per UMMA_K inside the 128 B swizzle row; SBO = 8 rows * 128 B.
                    // MN-major: +16 K-rows * 128 B; LBO = next 64-wide MN atom (64 K-rows * 128 B), SBO = 8 K-rows.
                    const uint64_t da = A_MN ? umma_desc(sa + kk * 2048, 8192, 1024) : umma_desc(sa + kk * 32, 16, 1024);
                    const uint64_t db = B_MN ? umma_desc(sb + kk * 2048, 8192, 1024) : umma_desc(sb + kk * 32, 16, 1024);
                    tc_mma_bf16(tmem_base, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                }
                tc_commit(empty_bar + s);  // frees the smem slot once these MMAs retire
            }
            tc_commit(tmem_full);
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int m = m0 + q * 32 + lane;
        const long boff = (long)b1 * p.sc1 + (long)b2 * p.sc2;
        const long row_off = boff + (long)m * p.ldc;
        const long res_off = boff + (long)m * p.ldres;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            const int nbase = n0 + c0;
            if (nbase >= p.n) break;
            float v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (m >= p.m) continue;
            const int nvalid = min(32, p.n - nbase);
            if (p.c_dtype == LASR_F32) epilogue_store<float>(p, v, row_off, m, nbase);
            else epilogue_store<bf16>(p, v, row_off, m, nbase);
            if (p.accumulate) continue;
            if (p.res) {
                const float* rr = p.res + res_off + nbase;
                if (p.vec_ok && nvalid == 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rr + j);
                        v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                    }
                } else {
                    for (int j = 0; j < nvalid; ++j) v[j] += rr[j];
                }
            }
            if (p.c_dtype == LASR_F32) {
                float* cr = reinterpret_cast<float*>(p.c) + row_off + nbase;
                if (p.vec_ok && nvalid == 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(cr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < nvalid; ++j) cr[j] = v[j];
                }
            } else {
                bf16* cr = reinterpret_cast<bf16*>(p.c) + row_off + nbase;
                if (p.vec_ok && nvalid == 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                        uint4 u;
                        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(cr + j) = u;
                    }
                } else {
                    for (int j = 0; j < nvalid; ++j) cr[j] = __float2bfloat16_rn(v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// operand stored row-major as (outer rows, inner contiguous) with row stride ld (elements), two batch levels.
static int make_map(CUtensorMap* map, const void* base, long inner, long rows, long ld, int nb2, long s2, int nb1, long s1,
                    int box_inner, int box_rows) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return LASR_ERR_DRIVER;
    }
    const bool use2 = (s2 != 0 && nb2 > 1), use1 = (s1 != 0 && nb1 > 1);
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)(use2 ? nb2 : 1), (cuuint64_t)(use1 ? nb1 : 1)};
    const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
    cuuint64_t strides[3] = {row_bytes, use2 ? (cuuint64_t)s2 * 2 : row_bytes, use1 ? (cuuint64_t)s1 * 2 : row_bytes};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        set_error("gemm_tc: operand base/strides must be 16-byte aligned (base=%p ld=%ld s2=%ld s1=%ld)", base, ld, s2, s1);
        return LASR_ERR_BAD_ARG;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): inner=%ld rows=%ld ld=%ld", (int)r, inner, rows, ld);
        return LASR_ERR_DRIVER;
    }
    return LASR_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int nbatch, cudaStream_t st) {
    using Cfg = TileCfg<BN>;
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return check_launch("gemm_tc smem attr");
        configured = true;
    }
    dim3 grid(ceil_div(p.m, BM), ceil_div(p.n, BN), nbatch * p.split_k);
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
    return check_launch("gemm_tc");
}

int gemm_tc_dispatch(const lasr_gemm_args* a, cudaStream_t st) {
    const int bn = (a->n > 128 && (a->n % 256 == 0 || a->n > 512)) ? 256 : (a->n > 64 ? 128 : 64);
    CUtensorMap ma, mb;
    int rc;
    if (!a->trans_a) rc = make_map(&ma, a->a, a->k, a->m, a->lda, a->batch2, a->sa2, a->batch1, a->sa1, BK, BM);
    else rc = make_map(&ma, a->a, a->m, a->k, a->lda, a->batch2, a->sa2, a->batch1, a->sa1, 64, BK);
    if (rc) return rc;
    if (!a->trans_b) rc = make_map(&mb, a->b, a->k, a->n, a->ldb, a->batch2, a->sb2, a->batch1, a->sb1, BK, bn);
    else rc = make_map(&mb, a->b, a->n, a->k, a->ldb, a->batch2, a->sb2, a->batch1, a->sb1, 64, BK);
    if (rc) return rc;

    TcParams p;
    p.c = a->c; p.bias = a->bias; p.res = a->res; p.aux = a->aux;
    p.m = a->m; p.n = a->n; p.k = a->k; p.c_dtype = a->c_dtype;
    p.ldc = a->ldc; p.ldres = a->ldres; p.batch2 = a->batch2; p.sc1 = a->sc1; p.sc2 = a->sc2;
    p.a_b1 = (a->sa1 != 0 && a->batch1 > 1); p.a_b2 = (a->sa2 != 0 && a->batch2 > 1);
    p.b_b1 = (a->sb1 != 0 && a->batch1 > 1); p.b_b2 = (a->sb2 != 0 && a->batch2 > 1);
    p.alpha = a->alpha; p.act = a->act; p.accumulate = a->accumulate; p.split_k = a->split_k < 1 ? 1 : a->split_k;
    const long esz = a->c_dtype == LASR_F32 ? 4 : 2;
    auto al16 = [&](long elems, long es) { return ((elems * es) & 15) == 0; };
    p.vec_ok = ((reinterpret_cast<uintptr_t>(a->c) & 15) == 0) && al16(a->ldc, esz) && al16(a->sc1, esz) && al16(a->sc2, esz);
    if (a->aux) p.vec_ok = p.vec_ok && ((reinterpret_cast<uintptr_t>(a->aux) & 15) == 0);
    if (a->res) p.vec_ok = p.vec_ok && ((reinterpret_cast<uintptr_t>(a->res) & 15) == 0) && al16(a->ldres, 4) && al16(a->sc1, 4) && al16(a->sc2, 4);
    const int nb = a->batch1 * a->batch2;

#define LASR_TC_CASE(BN_)                                                                     \
    if (bn == BN_) {                                                                          \
        if (!a->trans_a && !a->trans_b) return launch_tc<BN_, false, false>(ma, mb, p, nb, st); \
        if (!a->trans_a && a->trans_b) return launch_tc<BN_, false, true>(ma, mb, p, nb, st);   \
        if (a->trans_a && !a->trans_b) return launch_tc<BN_, true, false>(ma, mb, p, nb, st);   \
        return launch_tc<BN_, true, true>(ma, mb, p, nb, st);                                 \
    }
    LASR_TC_CASE(64)
    LASR_TC_CASE(128)
    LASR_TC_CASE(256)
#undef LASR_TC_CASE
    set_error("gemm_tc: no tile config");
    return LASR_ERR_UNSUPPORTED;
}

}  // namespace lasr
