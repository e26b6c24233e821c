// SpecAugment on the GPU (utils/transform/spec_augment.py:19-125 of the reference).
// The random decisions are taken on the host in the reference's RNG order (liteasr_b200/utils/transform/spec_augment.py); this
// kernel applies them to a padded (B, Tmax, F) fp32 batch, one CTA per utterance:
//   1. time warp: rows [0, center) -> warped rows and rows [center, T) -> T - warped rows, each with Pillow's BICUBIC resampling
//      along time (support 2 x max(in/out, 1), Keys cubic a = -0.5, window int(c -/+ support + 0.5) clipped, coefficients
//      normalised by their sum, products and sums in double WITHOUT contraction, cast to float32) -- bit-identical to
//      PIL.Image.resize on mode "F";
//   2. frequency masks, then time masks, in order, each filled with the mean of the CURRENT array (or zero).
#include "common.cuh"

namespace lasr {

constexpr int SA_MAXM = 8;                    // masks per kind
constexpr int SA_NPAR = 5 + 4 * SA_MAXM;      // T, center, warped, n_freq, n_time, (lo,hi) x 8 freq, (lo,hi) x 8 time

__device__ __forceinline__ double cubic_keys(double x) {
    const double a = -0.5;
    x = fabs(x);
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(a + 2.0, x), -(a + 3.0)), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), a);
    return 0.0;
}

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += scratch[i];  // fixed order: deterministic
    return r;
}

__global__ void __launch_bounds__(512) spec_augment_kernel(const float* __restrict__ x, float* __restrict__ y, long ldb, int F,
                                                           const int* __restrict__ params, int replace_with_zero) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const int* p = params + (long)b * SA_NPAR;
    const int T = p[0], center = p[1], warped = p[2], nf = min(p[3], SA_MAXM), nt = min(p[4], SA_MAXM);
    const float* xb = x + (long)b * ldb;
    float* yb = y + (long)b * ldb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // ---- 1. time warp (or copy)
    for (int t = warp; t < T; t += nwarps) {
        if (center < 0) {
            for (int f = lane; f < F; f += 32) yb[(long)t * F + f] = xb[(long)t * F + f];
            continue;
        }
        const bool left = t < warped;
        const int in0 = left ? 0 : center, in_rows = left ? center : T - center;
        const int out_rows = left ? warped : T - warped, yy = left ? t : t - warped;
        const double scale = (double)in_rows / (double)out_rows;
        const double filterscale = scale > 1.0 ? scale : 1.0;
        const double support = 2.0 * filterscale, ss = 1.0 / filterscale;
        const double c = __dmul_rn((double)yy + 0.5, scale);
        int ymin = (int)(c - support + 0.5);
        if (ymin < 0) ymin = 0;
        int ymax = (int)(c + support + 0.5);
        if (ymax > in_rows) ymax = in_rows;
        ymax -= ymin;
        double ww = 0.0;
        for (int k = 0; k < ymax; ++k) ww = __dadd_rn(ww, cubic_keys(__dmul_rn(__dadd_rn(__dadd_rn((double)(k + ymin), -c), 0.5), ss)));
        for (int f = lane; f < F; f += 32) {
            double acc = 0.0;
            for (int k = 0; k < ymax; ++k) {
                const double wk = cubic_keys(__dmul_rn(__dadd_rn(__dadd_rn((double)(k + ymin), -c), 0.5), ss)) / ww;
                acc = __dadd_rn(acc, __dmul_rn((double)xb[(long)(in0 + ymin + k) * F + f], wk));
            }
            yb[(long)t * F + f] = (float)acc;
        }
    }
    __syncthreads();
    // ---- 2. masks, in the reference's order; the fill value is the mean of the array as it is at that moment
    const long n = (long)T * F;
    for (int m = 0; m < nf + nt; ++m) {
        const bool is_f = m < nf;
        const int lo = is_f ? p[5 + 2 * m] : p[5 + 2 * SA_MAXM + 2 * (m - nf)];
        int hi = is_f ? p[6 + 2 * m] : p[6 + 2 * SA_MAXM + 2 * (m - nf)];
        hi = min(hi, is_f ? F : T);
        float fill = 0.f;
        if (!replace_with_zero) {
            double s = 0.0;
            for (long i = threadIdx.x; i < n; i += blockDim.x) s += (double)yb[i];
            fill = (float)(block_sum_d(s, scratch) / (double)n);
        }
        if (is_f) {
            const int wdt = hi - lo;
            if (wdt > 0)
                for (long i = threadIdx.x; i < (long)T * wdt; i += blockDim.x) yb[(i / wdt) * F + lo + (i % wdt)] = fill;
        } else {
            for (long i = (long)lo * F + threadIdx.x; i < (long)hi * F; i += blockDim.x) yb[i] = fill;
        }
        __syncthreads();
    }
}

}  // namespace lasr

extern "C" {

int lasr_spec_augment_npar(void) { return lasr::SA_NPAR; }

int lasr_spec_augment(const float* x, float* y, int64_t ld_batch, int F, const int32_t* params, int B, int replace_with_zero,
                      void* stream) {
    using namespace lasr;
    LASR_REQUIRE(x && y && params && x != y && B > 0 && F > 0 && ld_batch >= F, "spec_augment: bad args (out of place only)");
    spec_augment_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(x, y, ld_batch, F, params, replace_with_zero);
    return check_launch("spec_augment");
}

}  // extern "C"
