// Shared device/host helpers for liblasr (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/lasr.h"

#ifndef LASR_DEVICE_TIMEOUT_CYCLES
#define LASR_DEVICE_TIMEOUT_CYCLES (4000000000LL)  // ~2 s at 1.9 GHz: a broken pipeline traps instead of hanging the box
#endif

namespace lasr {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define LASR_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            lasr::set_error(__VA_ARGS__); \
            return LASR_ERR_BAD_ARG;     \
        }                                \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl may be scheduled while its predecessor in the
// stream is still draining; LASR_PDL_SYNC() -- the FIRST statement of every such kernel, before any global-memory access --
// lets the successor do the same and then blocks until the predecessor has completed and its writes are visible.  Launched
// with plain <<< >>> the two instructions are no-ops.  ~1 us per launch on the ~1000-launch training step.
#define LASR_PDL_SYNC() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")

template <typename... Exp, typename... Act>
static inline void launch_pdl(void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static int pdl = -1;  // LASR_PDL=0: developer switch (plain stream-ordered launches)
    if (pdl < 0) { const char* e = getenv("LASR_PDL"); pdl = e ? atoi(e) : 1; }
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<Exp>(args)...);  // failures surface through check_launch()
}

// ---------------------------------------------------------------------------------------------
// load/store with dtype conversion (fp32 math everywhere)
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x <= 1024 (scratch: 32 floats of shared memory)
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// ex2 + rcp on the MUFU pipe (__fdividef(1, y) = rcp.approx, <= 1 ulp): the IEEE division expands to ~8 instructions per element
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dswishf_(float x) {
    const float s = sigmoidf_(x);
    return s * (1.f + x * (1.f - s));
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == LASR_ACT_RELU) return fmaxf(v, 0.f);
    if (act == LASR_ACT_SWISH) return swishf_(v);
    return v;
}

}  // namespace lasr
