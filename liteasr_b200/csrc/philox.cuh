// Dropout masks from a counter-based generator: Philox4x32 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11; Random123 constants) with LASR_PHILOX_ROUNDS = 7 rounds -- the paper's (and Random123's philox4x32_R<7>)
// minimum Crush-resistant round count; the 10-round default only adds a safety margin that dropout masks do not need, and the
// rounds are the dominant cost of dropout here (integer instructions in issue-bound GEMM epilogues: 7 rounds instead of 10
// removes 30 % of them).  The round function is pinned by the Random123 known-answer vectors at 10 rounds
// (oracle/philox_oracle.py, tests/test_philox_cpu.py, and lasr_philox_raw on the GPU).  Masks are evaluated INSIDE the kernels that produce or consume the dropped tensor -- no mask
// tensor exists, the backward pass regenerates (or, for the FFN inner site, reads back from the saved pre-activation) the mask
// the forward pass applied.
//
// The reference applies nn.Dropout / F.dropout at ten kinds of sites (nets/positional_encoding.py:55,75, nets/attention.py:55,
// nets/feed_forward.py:19, nets/conformer_layer.py:42,54,63,125, nets/transformer_layer.py:48,58,174, nets/ctc.py:29) with masks
// drawn from torch's global generator; that stream cannot be replayed by fused kernels, so the stream here is the
// library's own:
//
//   keep(row, col) of a logical row-major (rows, n) tensor at dropout site `site` in optimizer step `step`:
//       w[0..3] = philox4x32<7>(counter = (col >> 3, row, site, step), key = (seed_lo, seed_hi))
//       u16     = 16-bit lane (col & 7) of w  (lane e = word e >> 1, low half first)
//       keep  <=>  u16 >= thr,   thr = round(p * 65536),   kept values are multiplied by scale = 65536 / (65536 - thr)
//
// One Philox call serves 8 consecutive columns.  `state` is a device array {seed, step} (two uint64) so that a captured CUDA
// graph draws fresh masks on every replay: lasr_rng_advance increments `step` inside the graph.
// oracle/philox_oracle.py restates this in numpy (pinned by the Random123 known-answer vectors) for the parity tests.
#pragma once
#include <stdint.h>

namespace lasr {

struct DropCfg {
    const unsigned long long* state;  // {seed, step} in device memory; nullptr = dropout off
    uint32_t site;
    uint32_t thr;                     // 0 = off
    float scale;
};

struct DropKey {
    uint32_t k0, k1, site, step, thr;
    float scale;
};

__device__ __forceinline__ DropKey drop_key(const DropCfg& c) {
    DropKey k;
    const unsigned long long seed = __ldg(c.state), step = __ldg(c.state + 1);
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    k.site = c.site;
    k.step = (uint32_t)step;
    k.thr = c.thr;
    k.scale = c.scale;
    return k;
}

#define LASR_PHILOX_ROUNDS 7

template <int ROUNDS = LASR_PHILOX_ROUNDS>
__device__ __forceinline__ void philox4x32(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// keep bits of the 8 columns [8g, 8g + 8) of `row`: bit e set <=> column 8g + e is kept
__device__ __forceinline__ uint32_t drop_keep8(const DropKey& k, uint32_t row, uint32_t g) {
    uint32_t c0 = g, c1 = row, c2 = k.site, c3 = k.step;
    philox4x32<>(c0, c1, c2, c3, k.k0, k.k1);
    uint32_t bits = 0;
    bits |= ((c0 & 0xffffu) >= k.thr) ? 1u : 0u;
    bits |= ((c0 >> 16) >= k.thr) ? 2u : 0u;
    bits |= ((c1 & 0xffffu) >= k.thr) ? 4u : 0u;
    bits |= ((c1 >> 16) >= k.thr) ? 8u : 0u;
    bits |= ((c2 & 0xffffu) >= k.thr) ? 16u : 0u;
    bits |= ((c2 >> 16) >= k.thr) ? 32u : 0u;
    bits |= ((c3 & 0xffffu) >= k.thr) ? 64u : 0u;
    bits |= ((c3 >> 16) >= k.thr) ? 128u : 0u;
    return bits;
}

// keep bits of the 32 columns [col0, col0 + 32), col0 % 8 == 0
__device__ __forceinline__ uint32_t drop_keep32(const DropKey& k, uint32_t row, uint32_t col0) {
    const uint32_t g = col0 >> 3;
    return drop_keep8(k, row, g) | (drop_keep8(k, row, g + 1) << 8) | (drop_keep8(k, row, g + 2) << 16) | (drop_keep8(k, row, g + 3) << 24);
}

// keep bits (low 4) of the 4 columns [col, col + 4), col % 4 == 0
__device__ __forceinline__ uint32_t drop_keep4(const DropKey& k, uint32_t row, uint32_t col) {
    return (drop_keep8(k, row, col >> 3) >> (col & 4)) & 15u;
}

__device__ __forceinline__ bool drop_keep1(const DropKey& k, uint32_t row, uint32_t col) {
    return (drop_keep8(k, row, col >> 3) >> (col & 7)) & 1u;
}

// value written to the saved pre-activation of a DROPPED element of the FFN inner site: act'(.) of it is exactly 0 in both
// the tanh-based (gemm_tc.cu::dswish_scaled) and the exp-based (common.cuh::dswishf_) Swish derivative, so the activation-
// backward epilogue needs no mask at all
#define LASR_DROP_MARK (-1.0e30f)

}  // namespace lasr
