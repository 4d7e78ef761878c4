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

// Zero-fill of `batch` strided [rows][cols] blocks as ONE kernel node with a programmatic edge on both sides (a memset node per
// block in front of every split-K GEMM cost the update's graph a node hop each, and cannot be a PDL predecessor).
static __global__ void zero2d_kernel(float* __restrict__ p, int ld, long long bs, int rows, int cols, int vec) {
    pdl_wait();
    pdl_launch();
    float* base = p + (long long)blockIdx.y * bs;
    if (vec) {
        const int c4 = cols >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * c4; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / c4), c = (int)(i - (long long)r * c4);
            reinterpret_cast<float4*>(base + (long long)r * ld)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * cols; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
            base[(long long)r * ld + c] = 0.f;
        }
    }
}
static inline int zero2d(float* p, int ld, long long bs, int rows, int cols, int batch, void* stream) {
    if (rows <= 0 || cols <= 0 || batch <= 0) return 0;
    const int vec = !((size_t)p & 15) && !(ld & 3) && !(bs & 3) && !(cols & 3);
    long long n = (long long)rows * (vec ? cols >> 2 : cols);
    int gx = (int)((n + 255) / 256);
    if (gx > 592) gx = 592;
    return launch_pdl(zero2d_kernel, dim3(gx, batch), dim3(256), 0, stream, p, ld, bs, rows, cols, vec);
}
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
