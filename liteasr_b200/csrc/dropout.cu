// Dropout as a stand-alone pass, for the sites that do not sit behind a GEMM epilogue or a LayerNorm backward (philox.cuh has
// the generator and the indexing): the positional-embedding table (nets/positional_encoding.py:75), the CTC head's input
// (nets/ctc.py:29, fused with the fp32 -> operand-dtype cast), the decoder's embedded input and its gradient
// (nets/positional_encoding.py:55), and attention probabilities / their gradient when an attention dropout rate is non-zero
// (nets/attention.py:55; 0.0 in the reference's shipped config, so that path keeps the unfused attention sequence).
//   y[r, c] = keep(r, c) * scale * x[r, c]      (x fp32|bf16 -> y fp32|bf16, in place allowed when the dtypes agree)
// plus the step counter that makes a captured CUDA graph draw fresh masks on every replay.
#include "common.cuh"
#include "philox.cuh"

namespace lasr {

__global__ void rng_advance_kernel(unsigned long long* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += 1ULL;
}

// known-answer hook: one Philox4x32 block with a run-time round count (7 = the masks, 10 = the Random123 vectors)
__global__ void philox_raw_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int rounds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t c0 = in[0], c1 = in[1], c2 = in[2], c3 = in[3];
    if (rounds == 10) philox4x32<10>(c0, c1, c2, c3, in[4], in[5]);
    else philox4x32<LASR_PHILOX_ROUNDS>(c0, c1, c2, c3, in[4], in[5]);
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) dropout_kernel(const TX* __restrict__ x, long ldx, TY* __restrict__ y, long ldy, long rows, int cols,
                                                      DropCfg cfg) {
    LASR_PDL_SYNC();
    const int groups = (cols + 15) >> 4;
    const long total = rows * groups;
    const DropKey dk = drop_key(cfg);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = i / groups;
        const int g = (int)(i - r * groups);
        const uint32_t keep = drop_keep16(dk, (uint32_t)r, (uint32_t)g);
        const TX* xr = x + r * ldx + 16 * g;
        TY* yr = y + r * ldy + 16 * g;
        const int n = min(16, cols - 16 * g);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            if (e < n) {
                const float v = to_f32<TX>(xr[e]);
                yr[e] = from_f32<TY>(((keep >> e) & 1u) ? v * dk.scale : 0.f);
            }
        }
    }
}

}  // namespace lasr

extern "C" {
using namespace lasr;

int lasr_rng_advance(void* state, void* stream) {
    LASR_REQUIRE(state, "rng_advance: null state");
    rng_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long*)state);
    return check_launch("rng_advance");
}

int lasr_philox_raw(const uint32_t* ctr_key, uint32_t* out, int rounds, void* stream) {
    LASR_REQUIRE(ctr_key && out && (rounds == 10 || rounds == LASR_PHILOX_ROUNDS), "philox_raw: rounds must be 10 or %d", LASR_PHILOX_ROUNDS);
    philox_raw_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctr_key, out, rounds);
    return check_launch("philox_raw");
}

int lasr_dropout(const void* x, int x_dtype, int64_t ldx, void* y, int y_dtype, int64_t ldy, int64_t rows, int cols,
                 const void* state, uint32_t site, uint32_t thr, float scale, void* stream) {
    LASR_REQUIRE(x && y && state && rows > 0 && cols > 0 && thr <= LASR_DROP_THR_MAX, "dropout: bad args (thr = round(p * 32768) <= 0x7c00)");
    LASR_REQUIRE(rows <= 0xffffffffLL, "dropout: more than 2^32 rows");
    DropCfg cfg;
    cfg.state = (const unsigned long long*)state; cfg.site = site; cfg.thr = thr; cfg.scale = scale;
    const long total = rows * ((cols + 15) / 16);
    int grid = ceil_div(total, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == LASR_F32 && y_dtype == LASR_F32) launch_pdl(dropout_kernel<float, float>, grid, 256, 0, st, (const float*)x, ldx, (float*)y, ldy, rows, cols, cfg);
    else if (x_dtype == LASR_F32 && y_dtype == LASR_BF16) launch_pdl(dropout_kernel<float, bf16>, grid, 256, 0, st, (const float*)x, ldx, (bf16*)y, ldy, rows, cols, cfg);
    else if (x_dtype == LASR_BF16 && y_dtype == LASR_BF16) launch_pdl(dropout_kernel<bf16, bf16>, grid, 256, 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, rows, cols, cfg);
    else if (x_dtype == LASR_BF16 && y_dtype == LASR_F32) launch_pdl(dropout_kernel<bf16, float>, grid, 256, 0, st, (const bf16*)x, ldx, (float*)y, ldy, rows, cols, cfg);
    else { set_error("dropout: bad dtype"); return LASR_ERR_UNSUPPORTED; }
    return check_launch("dropout");
}

}  // extern "C"
