// fp32 SIMT GEMM ("fp32 parity mode"): same contract as the tcgen05 kernel, arbitrary strides.
//   C[b1,b2] (M,N) = alpha * act(A . B^T + bias) (+ res)     or   C += alpha * A.B^T (atomics)
// 64x64 tile, BK = 16, 256 threads x (4x4) register micro-tile, smem double-buffer-free classic loop.
// Not a roofline kernel: it exists so fp32 runs (loss/grad parity, bit-exact greedy CTC) see fp32 operands and
// fp32 results like the reference's torch fp32 path (dot products are accumulated in fp64 and rounded once).
#include "common.cuh"
#include "philox.cuh"

namespace lasr {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtParams {
    const float* a;
    const float* b;
    float* c;
    const float* bias;
    const float* res;
    float* aux;
    int m, n, k;
    long sam, sak, sbn, sbk;  // element strides of A(m,k) and B(n,k)
    long ldc, ldres;
    int batch2;
    long sa1, sa2, sb1, sb2, sc1, sc2;
    float alpha;
    int act, accumulate, split_k;
    const float* dact;
    long lddact;
    float* colsum;
    long cs1, cs2;
    int n_store;
    DropCfg drop;
    int drop_mark;
    int aux_deriv;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int split = blockIdx.z % p.split_k, batch = blockIdx.z / p.split_k;
    const int b2 = batch % p.batch2, b1 = batch / p.batch2;
    const float* A = p.a + (long)b1 * p.sa1 + (long)b2 * p.sa2;
    const float* B = p.b + (long)b1 * p.sb1 + (long)b2 * p.sb2;
    const int kchunk = ((p.k + p.split_k - 1) / p.split_k + SBK - 1) / SBK * SBK;
    const int kbeg = split * kchunk, kend = min(p.k, kbeg + kchunk);
    if (kbeg >= kend) return;

    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4
    // products of two fp32 values are exact in fp64; accumulating there makes every output a correctly rounded fp32 dot
    // product, so ReLU masks / argmax decisions flip no more often than in an fp64 run (this kernel is the parity mode)
    double acc[4][4] = {};
    // loader mapping: choose the index that walks the contiguous dimension fastest
    const bool a_kfast = (p.sak == 1), b_kfast = (p.sbk == 1);
    for (int k0 = kbeg; k0 < kend; k0 += SBK) {
#pragma unroll
        for (int i = 0; i < (SBM * SBK) / 256; ++i) {
            const int e = tid + i * 256;
            int mm, kk;
            if (a_kfast) { kk = e % SBK; mm = e / SBK; } else { mm = e % SBM; kk = e / SBM; }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < p.m && gk < kend) ? A[(long)gm * p.sam + (long)gk * p.sak] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (SBN * SBK) / 256; ++i) {
            const int e = tid + i * 256;
            int nn, kk;
            if (b_kfast) { kk = e % SBK; nn = e / SBK; } else { nn = e % SBN; kk = e / SBN; }
            const int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < p.n && gk < kend) ? B[(long)gn * p.sbn + (long)gk * p.sbk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma((double)av[i], (double)bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const long boff = (long)b1 * p.sc1 + (long)b2 * p.sc2;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    const bool drop_on = p.drop.thr != 0;
    DropKey dk = {};
    if (drop_on) dk = drop_key(p.drop);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.m) continue;
        // the thread's 4 columns n0 + 4 tx .. + 3 share one Philox call (same 8-column group)
        const uint32_t keep4 = drop_on ? drop_keep4(dk, (uint32_t)m, (uint32_t)(n0 + tx * 4)) : 15u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.n_store) continue;
            const long off = boff + (long)m * p.ldc + n;
            if (p.accumulate) {
                atomicAdd(p.c + off, p.alpha * (float)acc[i][j]);
                continue;
            }
            float v = (float)acc[i][j];
            const bool keep = (keep4 >> j) & 1u;
            if (p.dact) {
                const float sv = p.dact[boff + (long)m * p.lddact + n];
                v *= p.alpha * (p.act == LASR_ACT_SWISH ? dswishf_(sv) : (p.act == LASR_ACT_MUL ? sv : (sv > 0.f ? 1.f : 0.f)));
            } else {
                if (p.bias) v += p.bias[n];
                if (p.aux) {
                    if (p.aux_deriv) p.aux[off] = (keep || !p.drop_mark) ? dswishf_(v) : 0.f;   // act'(pre-activation); 0 where dropped
                    else p.aux[off] = (keep || !p.drop_mark) ? v : LASR_DROP_MARK;
                }
                v = p.alpha * apply_act(v, p.act);
                if (drop_on) v = keep ? v * dk.scale : 0.f;
                if (p.res) v += p.res[boff + (long)m * p.ldres + n];
            }
            p.c[off] = v;
            cs[j] += v;
        }
    }
    if (p.colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < p.n) atomicAdd(p.colsum + (long)b1 * p.cs1 + (long)b2 * p.cs2 + n, cs[j]);
        }
    }
}

int gemm_simt_dispatch(const lasr_gemm_args* a, cudaStream_t st) {
    SimtParams p;
    p.a = (const float*)a->a; p.b = (const float*)a->b; p.c = (float*)a->c;
    p.bias = a->bias; p.res = a->res; p.aux = (float*)a->aux;
    p.m = a->m; p.n = a->n; p.k = a->k;
    p.sam = a->trans_a ? 1 : a->lda; p.sak = a->trans_a ? a->lda : 1;
    p.sbn = a->trans_b ? 1 : a->ldb; p.sbk = a->trans_b ? a->ldb : 1;
    p.ldc = a->ldc; p.ldres = a->ldres; p.batch2 = a->batch2;
    p.sa1 = a->sa1; p.sa2 = a->sa2; p.sb1 = a->sb1; p.sb2 = a->sb2; p.sc1 = a->sc1; p.sc2 = a->sc2;
    p.alpha = a->alpha; p.act = a->act; p.accumulate = a->accumulate; p.split_k = a->split_k < 1 ? 1 : a->split_k;
    p.n_store = a->n_store ? a->n_store : a->n;
    p.dact = (const float*)a->dact; p.lddact = a->lddact; p.colsum = a->colsum; p.cs1 = a->cs1; p.cs2 = a->cs2;
    p.drop.state = (const unsigned long long*)a->drop_state; p.drop.site = a->drop_site; p.drop.thr = a->drop_thr; p.drop.scale = a->drop_scale;
    p.drop_mark = a->drop_mark_aux;
    p.aux_deriv = a->aux_deriv;
    dim3 grid(ceil_div(a->m, SBM), ceil_div(p.n_store, SBN), a->batch1 * a->batch2 * p.split_k);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("gemm_simt");
}

}  // namespace lasr
