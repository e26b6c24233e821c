// Fused optimizer tail: global-norm clip + NaN skip + Noam LR + Adam over the FLAT
// parameter / gradient buffers, no host synchronisation (trainer.py:153-171, optims/noam.py:33-46,
// optims/adam.py:27-34).  The reference syncs the host on math.isnan(grad_norm); here the step counter
// lives on the device and only advances when the gradient norm is finite, so the skip decision, the
// learning rate and the update are all taken on the GPU.
#include "common.cuh"

namespace lasr {

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long)gridDim.x * 1024) {
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(g + i);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        } else {
            for (long j = i; j < n; ++j) s += g[j] * g[j];
        }
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// state[0] = step (as float, advanced only on finite norm), state[1] = last grad norm, state[2] = last lr, state[3] = skipped flag
__global__ void __launch_bounds__(256) optim_prepare_kernel(const float* __restrict__ partial, int nblk, float grad_mult,
                                                            float max_norm, float noam_factor, float model_dim, float warmup,
                                                            float fixed_lr, float* __restrict__ state) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 256) s += (double)partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(sh[0]) * fabsf(grad_mult);
        const bool ok = isfinite(norm);
        float step = state[0];
        if (ok) step += 1.f;
        float lr = fixed_lr;
        if (noam_factor > 0.f) lr = noam_factor * rsqrtf(model_dim) * fminf(rsqrtf(step), step * powf(warmup, -1.5f));
        float coef = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        if (coef > 1.f) coef = 1.f;
        state[0] = step;
        state[1] = norm;
        state[2] = lr;
        state[3] = ok ? 0.f : 1.f;
        state[4] = coef * grad_mult;
    }
}

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long n, float beta1, float beta2, float eps,
                                                        float weight_decay, const float* __restrict__ state) {
    if (state[3] != 0.f) return;  // NaN/inf gradient norm: skip the step (trainer.py:157-169)
    const float step = state[0], lr = state[2], gs = state[4];
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    const float step_size = lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
    for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long)gridDim.x * 1024) {
        if (i + 3 < n) {
            float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gr = ga[j] * gs + weight_decay * pa[j];
                ma[j] = beta1 * ma[j] + (1.f - beta1) * gr;
                va[j] = beta2 * va[j] + (1.f - beta2) * gr * gr;
                pa[j] -= step_size * ma[j] / (sqrtf(va[j]) * inv_bc2_sqrt + eps);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long j = i; j < n; ++j) {
                float gr = g[j] * gs + weight_decay * p[j];
                m[j] = beta1 * m[j] + (1.f - beta1) * gr;
                v[j] = beta2 * v[j] + (1.f - beta2) * gr * gr;
                p[j] -= step_size * m[j] / (sqrtf(v[j]) * inv_bc2_sqrt + eps);
            }
        }
    }
}

}  // namespace lasr

extern "C" {
using namespace lasr;

/* workspace: 1024 floats of partial sums.  state: 8 floats (see optim.cu).  grad_mult scales the gradient
 * before the norm (1/world_size after an allreduce-sum, or 1).  noam_factor > 0 selects the Noam schedule
 * lr = factor * model_dim^-0.5 * min(step^-0.5, step * warmup^-1.5), else fixed_lr. */
int lasr_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_mult,
                        float max_norm, float beta1, float beta2, float eps, float weight_decay, float noam_factor, float model_dim,
                        float warmup, float fixed_lr, float* state, float* workspace, void* stream) {
    LASR_REQUIRE(params && grads && exp_avg && exp_avg_sq && state && workspace && n > 0, "clip_adam_step: bad args");
    LASR_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "clip_adam_step: unaligned");
    cudaStream_t st = (cudaStream_t)stream;
    int nblk = ceil_div(n, 1024 * 8);
    if (nblk > 1024) nblk = 1024;
    if (nblk < 1) nblk = 1;
    sumsq_partial_kernel<<<nblk, 256, 0, st>>>(grads, n, workspace);
    if (int rc = check_launch("sumsq_partial")) return rc;
    optim_prepare_kernel<<<1, 256, 0, st>>>(workspace, nblk, grad_mult, max_norm, noam_factor, model_dim, warmup, fixed_lr, state);
    if (int rc = check_launch("optim_prepare")) return rc;
    int grid = ceil_div(n, 1024 * 4);
    if (grid > 148 * 8) grid = 148 * 8;
    adam_step_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, beta1, beta2, eps, weight_decay, state);
    return check_launch("clip_adam_step");
}

int lasr_zero(void* ptr, size_t bytes, void* stream) {
    LASR_REQUIRE(ptr || bytes == 0, "zero: null pointer");
    if (bytes == 0) return LASR_OK;
    if (cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream) != cudaSuccess) return check_launch("zero");
    return LASR_OK;
}

}  // extern "C"
