// Saliency masking (rl_utils.py:76-82 compute_attribution_mask; sgsac.py:67-70 mask application) and
// overlay augmentation (augmentations.py:79-99).  The per-(sample,frame) torch.quantile threshold is a
// CTA-level radix-select of the two order statistics that bracket the rank (no sort), fused with the
// abs-max over the 3 channels of a frame, the >= compare and the fill of masked-out pixels.
#include "common.cuh"
#include "../../include/sgqn_b200.h"

// ---------------------------------------------------------------- global min / max of the obs batch (sgsac.py:68-69)
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float4* __restrict__ x, long long n4, float* __restrict__ part) {
    pdl_wait();
    pdl_launch();
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(x + i);
        lo = fminf(fminf(lo, v.x), fminf(fminf(v.y, v.z), v.w));
        hi = fmaxf(fmaxf(hi, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
    }
    __shared__ float slo[8], shi[8];
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { lo = fminf(lo, slo[i]); hi = fmaxf(hi, shi[i]); }
        part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi;
    }
}
__global__ void minmax_final_kernel(const float* __restrict__ part, int nblk, float* __restrict__ out) {
    pdl_wait();
    pdl_launch();
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < nblk; i += 32) { lo = fminf(lo, part[2 * i]); hi = fmaxf(hi, part[2 * i + 1]); }
    lo = warp_min(lo); hi = warp_max(hi);
    // out[2..3] = {-min, max}: the form a single MAX all-reduce makes global (data-parallel shards, dist.py)
    if (threadIdx.x == 0) { out[0] = lo; out[1] = hi; out[2] = -lo; out[3] = hi; }
}

extern "C" int sgqn_minmax(const float* x, long long n, float* scratch, float* out4, void* stream) {
    if (n <= 0 || (n & 3)) return (int)cudaErrorInvalidValue;
    int nblk = 296;
    { int rc_ = launch_pdl(minmax_partial_kernel, dim3(nblk), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, n / 4, scratch); if (rc_) return rc_; }
    { int rc_ = launch_pdl(minmax_final_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, scratch, nblk, out4); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- quantile mask
// torch.quantile('linear') restated: rank = fp32(q)*(n-1); lo = floor(rank); w = rank - lo;
// thr = w < 0.5 ? s[lo] + w*(s[hi]-s[lo]) : s[hi] - (s[hi]-s[lo])*(1-w)   (no FMA contraction)
__device__ __forceinline__ float quantile_lerp(float x0, float x1, float w) {
    float d = __fsub_rn(x1, x0);
    if (w < 0.5f) return __fadd_rn(x0, __fmul_rn(w, d));
    return __fsub_rn(x1, __fmul_rn(d, __fsub_rn(1.0f, w)));
}

// one CTA per (sample, frame); dynamic smem: HW uint32 keys
__global__ void __launch_bounds__(256)
attribution_mask_kernel(const float* __restrict__ grad, const float* __restrict__ obs, const float* __restrict__ mm,
                        const float* __restrict__ u, float quantile, uint8_t* __restrict__ mask, float* __restrict__ masked,
                        int HW, int mm_neg) {
    pdl_wait();
    pdl_launch();
    extern __shared__ uint32_t keys[];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_krem;
    __shared__ uint32_t s_red[2][8];
    const int tid = threadIdx.x;
    const int b = blockIdx.x / 3, f = blockIdx.x % 3;
    const float* g0 = grad + ((size_t)b * 9 + 3 * f) * HW;

    // a = max_c |g|  (bit pattern of a non-negative float is order preserving as uint32)
    for (int i = tid; i < HW; i += 256) {
        uint32_t k0 = __float_as_uint(__ldg(g0 + i)) & 0x7fffffffu;
        uint32_t k1 = __float_as_uint(__ldg(g0 + HW + i)) & 0x7fffffffu;
        uint32_t k2 = __float_as_uint(__ldg(g0 + 2 * HW + i)) & 0x7fffffffu;
        keys[i] = max(k0, max(k1, k2));
    }
    const float rank = __fmul_rn(quantile, (float)(HW - 1));
    const float flo = floorf(rank);
    const float w = __fsub_rn(rank, flo);
    const uint32_t lo = (uint32_t)flo;
    if (tid == 0) { s_prefix = 0u; s_krem = lo; }
    __syncthreads();

    // radix select of the element of ascending rank `lo` (0-based), 8 bits per pass, MSB first
    uint32_t known = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (int i = tid; i < HW; i += 256) {
            uint32_t k = keys[i];
            if ((k & known) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            const uint32_t krem = s_krem;                  // read before the shuffles (they re-converge the warp)
            uint32_t c[8], s = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = hist[tid * 8 + q]; s += c[q]; }
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
            uint32_t excl = incl - s;
            if (krem >= excl && krem < incl) {             // exactly one lane
                uint32_t acc = excl;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (krem < acc + c[q]) { s_prefix = prefix | ((uint32_t)(tid * 8 + q) << shift); s_krem = krem - acc; break; }
                    acc += c[q];
                }
            }
        }
        known |= (255u << shift);
        __syncthreads();
    }
    const uint32_t v0 = s_prefix;                          // bits of s[lo]

    // s[lo+1]: v0 again when enough elements are <= v0, else the smallest key above v0
    uint32_t cnt_le = 0u, min_gt = 0xffffffffu;
    for (int i = tid; i < HW; i += 256) {
        uint32_t k = keys[i];
        if (k <= v0) ++cnt_le; else min_gt = min(min_gt, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o);
        min_gt = min(min_gt, __shfl_xor_sync(0xffffffffu, min_gt, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = cnt_le; s_red[1][tid >> 5] = min_gt; }
    __syncthreads();
    cnt_le = 0u; min_gt = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 8; ++i) { cnt_le += s_red[0][i]; min_gt = min(min_gt, s_red[1][i]); }
    const float x0 = __uint_as_float(v0);
    float x1 = x0;
    if (w != 0.0f && cnt_le < lo + 2u) x1 = __uint_as_float(min_gt);   // w == 0 -> ceil(rank) == lo
    const float thr = quantile_lerp(x0, x1, w);

    // mask + fill: masked = mask ? obs : lo + (hi - lo) * u   (sgsac.py:67-70)
    float fill = 0.f;
    if (masked) {
        const float lo_v = mm_neg ? -mm[0] : mm[0];          // mm_neg: {-min, max} (the all-reduced exchange form of sgqn_minmax)
        fill = __fadd_rn(lo_v, __fmul_rn(__fsub_rn(mm[1], lo_v), u[0]));
    }
    uint8_t* mrow = mask + ((size_t)b * 3 + f) * HW;
    const float* o0 = obs + ((size_t)b * 9 + 3 * f) * HW;
    float* d0 = masked ? masked + ((size_t)b * 9 + 3 * f) * HW : nullptr;
    for (int i = tid; i < HW; i += 256) {
        bool keep = __uint_as_float(keys[i]) >= thr;
        mrow[i] = keep ? 1 : 0;
        if (masked) {
            d0[i] = keep ? __ldg(o0 + i) : fill;
            d0[HW + i] = keep ? __ldg(o0 + HW + i) : fill;
            d0[2 * HW + i] = keep ? __ldg(o0 + 2 * HW + i) : fill;
        }
    }
}

extern "C" int sgqn_attribution_mask(const float* grad, const float* obs, const float* minmax, const float* u, float quantile,
                                     uint8_t* mask, float* masked_obs, int B, int HW, int minmax_neg, void* stream) {
    if (B <= 0) return 0;
    size_t smem = (size_t)HW * sizeof(uint32_t);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attribution_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    { int rc_ = launch_pdl(attribution_mask_kernel, dim3(3 * B), dim3(256), smem, (cudaStream_t)stream, grad, obs, minmax, u, quantile, mask, masked_obs, HW, minmax_neg); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- overlay: ((1-a)*(x/255) + a*(img/255))*255
__global__ void overlay_u8_kernel(const float* __restrict__ obs, const uint8_t* __restrict__ pool, const int64_t* __restrict__ ids,
                                  float one_minus_alpha, float alpha, float* __restrict__ out, int HW, long long total) {
    pdl_wait();
    pdl_launch();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int p = (int)(i % HW); long long t = i / HW; int c = (int)(t % 9); int b = (int)(t / 9);
    float img = (float)__ldg(pool + ((size_t)ids[b] * 3 + (c % 3)) * HW + p);
    float x = __fdiv_rn(__ldg(obs + i), 255.0f);
    float y = __fdiv_rn(img, 255.0f);
    out[i] = __fmul_rn(__fadd_rn(__fmul_rn(one_minus_alpha, x), __fmul_rn(alpha, y)), 255.0f);
}
__global__ void overlay_f32_kernel(const float* __restrict__ obs, const float* __restrict__ imgs, const int64_t* __restrict__ ids,
                                   float one_minus_alpha, float alpha, float* __restrict__ out, int HW, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int p = (int)(i % HW); long long t = i / HW; int c = (int)(t % 9); int b = (int)(t / 9);
    const size_t img = ids ? (size_t)ids[b] : (size_t)b;        // ids: rows of a device-resident image pool
    float y = __ldg(imgs + (img * 3 + (c % 3)) * HW + p);
    float x = __fdiv_rn(__ldg(obs + i), 255.0f);
    out[i] = __fmul_rn(__fadd_rn(__fmul_rn(one_minus_alpha, x), __fmul_rn(alpha, y)), 255.0f);
}

extern "C" int sgqn_overlay_u8(const float* obs, const uint8_t* pool, const int64_t* ids, float one_minus_alpha, float alpha,
                               float* out, int B, int HW, void* stream) {
    long long total = (long long)B * 9 * HW;
    if (total <= 0) return 0;
    { int rc_ = launch_pdl(overlay_u8_kernel, dim3((unsigned)cdivll(total, 256)), dim3(256), 0, (cudaStream_t)stream, obs, pool, ids, one_minus_alpha, alpha, out, HW, total); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}
extern "C" int sgqn_overlay_f32(const float* obs, const float* imgs, const int64_t* ids, float one_minus_alpha, float alpha,
                                float* out, int B, int HW, void* stream) {
    long long total = (long long)B * 9 * HW;
    if (total <= 0) return 0;
    overlay_f32_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(obs, imgs, ids, one_minus_alpha, alpha, out, HW, total);
    return SGQN_CHECK_LAUNCH();
}
