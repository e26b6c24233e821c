// Dropout masks from a counter-based generator: Philox4x32 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11; Random123 constants) with LASR_PHILOX_ROUNDS = 7 rounds -- the paper's (and Random123's philox4x32_R<7>)
// minimum Crush-resistant round count; the 10-round default only adds a safety margin that dropout masks do not need, and the
// rounds are the dominant cost of dropout here (integer instructions in issue-bound GEMM epilogues).  The round function is
// pinned by the Random123 known-answer vectors at 10 rounds (oracle/philox_oracle.py, tests/test_philox_cpu.py, and
// lasr_philox_raw on the GPU).  Masks are evaluated INSIDE the kernels that produce or consume the dropped tensor -- no mask
// tensor exists, the backward pass regenerates (or, for the FFN inner site, reads back from the saved activation derivative)
// the mask the forward pass applied.
//
// The reference applies nn.Dropout / F.dropout at ten kinds of sites (nets/positional_encoding.py:55,75, nets/attention.py:55,
// nets/feed_forward.py:19, nets/conformer_layer.py:42,54,63,125, nets/transformer_layer.py:48,58,174, nets/ctc.py:29) with masks
// drawn from torch's global generator; that stream cannot be replayed by fused kernels, so the stream here is the
// library's own:
//
//   keep(row, col) of a logical row-major (rows, n) tensor at dropout site `site` in optimizer step `step`:
//       w[0..3] = philox4x32<7>(counter = (col >> 4, row, site, step), key = (seed_lo, seed_hi))       one call = 16 columns
//       e = col & 15,  i = e >> 2,  j = e & 3
//       u15     = ((byte j of w[i]) << 8 | (byte j of w[i ^ 1])) & 0x7fff        15 uniform bits: two DIFFERENT bytes of the block
//       keep  <=>  u15 >= thr,   thr = round(p * 32768) <= 0x7c00,   kept values are multiplied by scale = 32768 / (32768 - thr)
//
// Why this shape: every element costs Philox instructions, so one call serves 16 columns instead of 8; a column's 15 bits are
// the byte it "owns" plus the same byte position of the neighbouring word (a byte is the high part of one column and the low
// part of another: the low part matters for 1 column in 128, so the induced correlation is invisible, and the pair is uniform
// over 2^15, so the keep rate is exactly 1 - thr / 32768); and 15-bit lanes in packed pairs are non-negative fp16 bit patterns
// whose order is the integer order, so ONE half2 compare (HSET2.GEU: NaN patterns, which lie above 0x7c00 >= thr, compare true)
// yields the 0xffff / 0x0000 lane masks of two columns -- the form a packed bf16x2 result needs (one AND per two elements).
//
// `state` is a device array {seed, step} (two uint64) so that a captured CUDA graph draws fresh masks on every replay:
// lasr_rng_advance increments `step` inside the graph.
// oracle/philox_oracle.py restates this in numpy (pinned by the Random123 known-answer vectors) for the parity tests.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace lasr {

#define LASR_DROP_THR_MAX 0x7c00u  // thr = round(p * 32768) must not exceed the fp16 +inf pattern (p <= 0.96875)

struct DropCfg {
    const unsigned long long* state;  // {seed, step} in device memory; nullptr = dropout off
    uint32_t site;
    uint32_t thr;                     // 0 = off
    float scale;
};

struct DropKey {
    uint32_t k0, k1, site, step, thr2;  // thr2 = thr | thr << 16
    float scale;
};

__device__ __forceinline__ DropKey drop_key(const DropCfg& c) {
    DropKey k;
    const unsigned long long seed = __ldg(c.state), step = __ldg(c.state + 1);
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    k.site = c.site;
    k.step = (uint32_t)step;
    k.thr2 = c.thr | (c.thr << 16);
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

__device__ __forceinline__ uint32_t drop_lane_mask2(uint32_t hi_word, uint32_t lo_word, uint32_t sel, uint32_t thr2) {
    const uint32_t u = __byte_perm(hi_word, lo_word, sel) & 0x7fff7fffu;
    return __hgeu2_mask(*reinterpret_cast<const __half2*>(&u), *reinterpret_cast<const __half2*>(&thr2));
}

// packed keep masks of the 16 columns [16g, 16g + 16) of `row`: m[t] covers columns 16g + 2t (low 16 bits) and 16g + 2t + 1
// (high 16 bits), 0xffff = kept, 0 = dropped -- AND it into a packed bf16x2 pair
__device__ __forceinline__ void drop_masks16(const DropKey& k, uint32_t row, uint32_t g, uint32_t (&m)[8]) {
    uint32_t w0 = g, w1 = row, w2 = k.site, w3 = k.step;
    philox4x32<>(w0, w1, w2, w3, k.k0, k.k1);
    m[0] = drop_lane_mask2(w0, w1, 0x1504u, k.thr2); m[1] = drop_lane_mask2(w0, w1, 0x3726u, k.thr2);
    m[2] = drop_lane_mask2(w1, w0, 0x1504u, k.thr2); m[3] = drop_lane_mask2(w1, w0, 0x3726u, k.thr2);
    m[4] = drop_lane_mask2(w2, w3, 0x1504u, k.thr2); m[5] = drop_lane_mask2(w2, w3, 0x3726u, k.thr2);
    m[6] = drop_lane_mask2(w3, w2, 0x1504u, k.thr2); m[7] = drop_lane_mask2(w3, w2, 0x3726u, k.thr2);
}

// full 32-bit masks of the two columns of a packed pair (for fp32 results: AND into the float's bits)
__device__ __forceinline__ uint32_t drop_mask_lo(uint32_t m2) { return __byte_perm(m2, 0u, 0x1100u); }
__device__ __forceinline__ uint32_t drop_mask_hi(uint32_t m2) { return __byte_perm(m2, 0u, 0x3322u); }
__device__ __forceinline__ float drop_and(float x, uint32_t mask32) { return __uint_as_float(__float_as_uint(x) & mask32); }

// keep bits of the 16 columns [16g, 16g + 16): bit e set <=> column 16g + e is kept
__device__ __forceinline__ uint32_t drop_keep16(const DropKey& k, uint32_t row, uint32_t g) {
    uint32_t m[8];
    drop_masks16(k, row, g, m);
    uint32_t bits = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) bits |= ((m[t] & 1u) | ((m[t] >> 15) & 2u)) << (2 * t);
    return bits;
}

// keep bits of the 32 columns [col0, col0 + 32), col0 % 16 == 0
__device__ __forceinline__ uint32_t drop_keep32(const DropKey& k, uint32_t row, uint32_t col0) {
    const uint32_t g = col0 >> 4;
    return drop_keep16(k, row, g) | (drop_keep16(k, row, g + 1) << 16);
}

// keep bits (low 4) of the 4 columns [col, col + 4), col % 4 == 0
__device__ __forceinline__ uint32_t drop_keep4(const DropKey& k, uint32_t row, uint32_t col) {
    return (drop_keep16(k, row, col >> 4) >> (col & 12u)) & 15u;
}

__device__ __forceinline__ bool drop_keep1(const DropKey& k, uint32_t row, uint32_t col) {
    return (drop_keep16(k, row, col >> 4) >> (col & 15u)) & 1u;
}

// value written to the saved pre-activation of a DROPPED element of the FFN inner site: act'(.) of it is exactly 0 in both
// the tanh-based (gemm_tc.cu::dswish_scaled) and the exp-based (common.cuh::dswishf_) Swish derivative, so the activation-
// backward epilogue needs no mask at all
#define LASR_DROP_MARK (-1.0e30f)

}  // namespace lasr
