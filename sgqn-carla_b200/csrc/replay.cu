// Replay sampling (utils.py:124-135,185-198): gather 3-frame uint8 stacks of the sampled transitions out of the
// device-resident frame ring and emit fp32 NCHW batches, with random_crop (augmentations.py:236-264) or
// random_shift (augmentations.py:229-233 == clamp-gather) fused into the read.  HBM-bound byte work.
#include "common.cuh"
#include "../../include/sgqn_b200.h"

// grid: (6*B) blocks; block -> (sample b, which in {obs,next}, frame j); 3 channels x Ho x Ho outputs per block
__global__ void __launch_bounds__(256)
replay_gather_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ fidx, const int64_t* __restrict__ idxs,
                     const int32_t* __restrict__ offs, float* __restrict__ obs, float* __restrict__ next_obs, int B, int Hs,
                     int Ho, int mode, int pad) {
    pdl_wait();
    pdl_launch();
    int blk = blockIdx.x;
    int j = blk % 3; int t = blk / 3; int which = t & 1; int b = t >> 1;
    long long tr = idxs[b];
    int slot = fidx[tr * 6 + which * 3 + j];
    const uint8_t* src = frames + (size_t)slot * 3 * Hs * Hs;
    float* dst = (which ? next_obs : obs) + ((size_t)b * 9 + 3 * j) * Ho * Ho;
    int oy = 0, ox = 0;
    if (offs) { oy = offs[(which * B + b) * 2 + 0]; ox = offs[(which * B + b) * 2 + 1]; }
    if (mode == 1) { oy -= pad; ox -= pad; }
    const int W4 = Ho >> 2;                    // Ho % 4 == 0 (84)
    const int n4 = 3 * Ho * W4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        int x4 = i % W4; int r = i / W4; int y = r % Ho; int c = r / Ho;
        int sy = y + oy;
        if (mode == 1) sy = min(max(sy, 0), Hs - 1);
        const uint8_t* row = src + ((size_t)c * Hs + sy) * Hs;
        int sx = x4 * 4 + ox;
        float4 v;
        if (mode == 1) {
            v.x = (float)__ldg(row + min(max(sx, 0), Hs - 1));
            v.y = (float)__ldg(row + min(max(sx + 1, 0), Hs - 1));
            v.z = (float)__ldg(row + min(max(sx + 2, 0), Hs - 1));
            v.w = (float)__ldg(row + min(max(sx + 3, 0), Hs - 1));
        } else if ((((uintptr_t)(row + sx)) & 3) == 0) {
            uchar4 u = __ldg(reinterpret_cast<const uchar4*>(row + sx));
            v = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
        } else {
            v.x = (float)__ldg(row + sx); v.y = (float)__ldg(row + sx + 1);
            v.z = (float)__ldg(row + sx + 2); v.w = (float)__ldg(row + sx + 3);
        }
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

extern "C" int sgqn_replay_gather(const uint8_t* frames, const int32_t* fidx, const int64_t* idxs, const int32_t* offs,
                                  float* obs, float* next_obs, int B, int Hs, int Ho, int mode, int pad, void* stream) {
    if (B <= 0) return 0;
    if (Ho & 3) return (int)cudaErrorInvalidValue;
    if (mode == 0 && Ho > Hs) return (int)cudaErrorInvalidValue;
    { int rc_ = launch_pdl(replay_gather_kernel, dim3(6 * B), dim3(256), 0, (cudaStream_t)stream, frames, fidx, idxs, offs, obs, next_obs, B, Hs, Ho, mode, pad); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// Prefetch of the NEXT update's batch from a host-resident (pinned, zero-copy) frame ring: the six frames of every sampled
// transition are copied as raw bytes into a device staging ring [B][6][fbytes], which sgqn_replay_gather then reads with
// fidx = arange(6B), idxs = arange(B).  Runs on a side stream under the current update, so the PCIe transfer of step t+1
// (16 MB at B = 128) is off the critical path of step t.
// Few, fat CTAs (grid-stride over the 6B frames, four 16-byte loads in flight per thread): enough outstanding PCIe reads to
// saturate the link (~47 GB/s measured) while leaving the SMs' block slots to the update's kernels it runs beside.
__global__ void __launch_bounds__(512)
frames_copy_kernel(const uint4* __restrict__ frames, const int32_t* __restrict__ fidx, const int64_t* __restrict__ idxs,
                   uint4* __restrict__ dst, int n16, int nframes) {
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        const int b = f / 6, j = f - b * 6;
        const uint4* src = frames + (size_t)fidx[idxs[b] * 6 + j] * n16;
        uint4* d = dst + (size_t)f * n16;
        int i = threadIdx.x;
        for (; i + 3 * 512 < n16; i += 4 * 512) {
            const uint4 v0 = src[i], v1 = src[i + 512], v2 = src[i + 1024], v3 = src[i + 1536];
            d[i] = v0; d[i + 512] = v1; d[i + 1024] = v2; d[i + 1536] = v3;
        }
        for (; i < n16; i += 512) d[i] = src[i];
    }
}
extern "C" int sgqn_frames_copy(const uint8_t* frames, const int32_t* fidx, const int64_t* idxs, uint8_t* dst, int B, int fbytes,
                                void* stream) {
    if (B <= 0) return 0;
    if ((fbytes & 15) || (((size_t)frames | (size_t)dst) & 15)) return (int)cudaErrorInvalidValue;
    static int ctas = 0;
    if (!ctas) { const char* e = getenv("SGQN_COPY_CTAS"); ctas = e ? atoi(e) : 16; if (ctas < 1) ctas = 1; }
    int grid = 6 * B < ctas ? 6 * B : ctas;
    frames_copy_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>((const uint4*)frames, fidx, idxs, (uint4*)dst, fbytes / 16, 6 * B);
    return SGQN_CHECK_LAUNCH();
}

// actions / rewards / not_dones rows of the sampled transitions
__global__ void take_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idxs, float* __restrict__ dst,
                                 int B, int width) {
    pdl_wait();
    pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * width) return;
    int b = i / width, c = i - b * width;
    dst[i] = __ldg(src + (size_t)idxs[b] * width + c);
}

extern "C" int sgqn_take_rows(const float* src, const int64_t* idxs, float* dst, int B, int width, void* stream) {
    if (B * width <= 0) return 0;
    { int rc_ = launch_pdl(take_rows_kernel, dim3(cdiv(B * width, 256)), dim3(256), 0, (cudaStream_t)stream, src, idxs, dst, B, width); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// crop / shift of an already materialised fp32 batch (public augmentations.random_crop / random_shift entry points)
__global__ void crop_shift_kernel(const float* __restrict__ x, const int32_t* __restrict__ offs, float* __restrict__ y, int B,
                                  int C, int Hs, int Ho, int mode, int pad, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int xo = (int)(i % Ho); long long t = i / Ho; int yo = (int)(t % Ho); t /= Ho; int c = (int)(t % C); int b = (int)(t / C);
    int sy = yo + offs[b * 2], sx = xo + offs[b * 2 + 1];
    if (mode == 1) { sy = min(max(sy - pad, 0), Hs - 1); sx = min(max(sx - pad, 0), Hs - 1); }
    y[i] = __ldg(x + (((size_t)b * C + c) * Hs + sy) * Hs + sx);
}

extern "C" int sgqn_crop_shift(const float* x, const int32_t* offs, float* y, int B, int C, int Hs, int Ho, int mode, int pad,
                               void* stream) {
    long long total = (long long)B * C * Ho * Ho;
    if (total <= 0) return 0;
    crop_shift_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(x, offs, y, B, C, Hs, Ho, mode, pad, total);
    return SGQN_CHECK_LAUNCH();
}

extern "C" int sgqn_zero(void* p, long long bytes, void* stream) {
    if (bytes <= 0) return 0;
    if (!(bytes & 3) && !((size_t)p & 3) && bytes < (1ll << 32))      // a kernel node with programmatic edges (common.cuh) instead of a memset node
        return zero2d((float*)p, 0, 0, 1, (int)(bytes >> 2), 1, stream);
    return (int)cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
}
