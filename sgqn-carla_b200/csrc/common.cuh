// Shared device helpers for the SGSAC update kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>
#include <utility>

#define SGQN_CHECK_LAUNCH() (int)cudaGetLastError()

// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may start -- block scheduling, shared /
// tensor memory allocation, barrier initialisation -- as soon as every CTA of the previous kernel on the stream has
// executed pdl_launch() or exited, i.e. under that kernel's last tiles instead of after its full completion + launch
// latency.  Rules every such kernel follows: (1) pdl_wait() before the first global-memory access that could depend on
// ANY earlier kernel (it returns when the previous kernel has completed and its writes are visible); (2) pdl_launch()
// only AFTER its own pdl_wait() and after its own tensor-memory allocation (a dependent CTA that grabbed TMEM first
// would otherwise spin in its wait while we spin in the allocation).  Without the launch attribute both are no-ops.
// SGQN_PDL=0 in the environment turns the attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static inline int pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SGQN_PDL"); v = e ? atoi(e) : 1; }
    return v;
}
template <typename... KArgs, typename... Args>
static inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

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
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}
