// First encoder conv (NormalizeImg + Conv2d(9, 32, 3, stride 2), modules.py:86-93,143) + the ReLU that follows it, as ONE
// tcgen05 kernel: the im2col matrix is built in shared memory, never in HBM.
//
// Per-position GEMM out[pix][32] = A[pix][96] * W[32][96]^T with A[pix][c*9 + ky*3 + kx] = tf32(obs[b][c][2y+ky][2x+kx] / 255)
// (81 real columns).  TMA cannot express the stride-2 window gather over an NCHW fp32 image, so the A tile of 128 output
// pixels is produced by sixteen producer warps (two groups of eight alternate tiles; the producers are latency-bound, so
// what matters is loads in flight): two threads per pixel gather its 81 values straight from the observation (k < 48 /
// k >= 48; one float2 + one float load per (channel, ky), consecutive lanes = consecutive 8-byte words), scale / round
// them and write their half of the row as 12 16-byte units in the SWIZZLE_128B K-major layout the UMMA
// descriptor expects (unit u of row r at ((u ^ (r & 7)) << 4): every quarter-warp store covers all 32 banks),
// fence.proxy.async, mbarrier arrive.  One thread issues the 12 MMAs (3 chunks x 4 k-steps of 128x32x8, TF32) per tile into
// one of two TMEM accumulators; four epilogue warps add bias, apply ReLU, round to TF32 (the next layer's operand format)
// and store the pitch-linear activation [n][43][41][32].
// The materialised version moved 384 B per output pixel through HBM twice (im2col write + GEMM read: 165 MB each way at
// 256 samples, 120 us); this one reads the observation once (65 MB) and writes the activation (58 MB).
// The weight gradient of the layer still wants A as a matrix: when `col` is given, the finished shared-memory tile is
// ALSO copied out by TMA (cp.async.bulk.tensor store, no thread instructions) for the tiles at or after `col_row0`.
#include "tc_common.cuh"
#include "../../include/sgqn_b200.h"

using namespace tc;

namespace {

constexpr int kStages = 4;                        // A stages; producer group g owns stages g and g + 2
constexpr int kChunkBytes = 128 * 128;            // [128 rows][32 floats]
constexpr int kABytes = 3 * kChunkBytes;          // 48 KB
constexpr int kWBytes = 3 * 32 * 128;             // 12 KB: three [32 co][32 k] chunks
constexpr int kSmem = kStages * kABytes + kWBytes + kChunkBytes + 1024 + 256;
constexpr int kThreads = 32 * 21;                 // warp 0 MMA, warps 1-4 epilogue, warps 5-20 producers
constexpr int kPix = 41 * 41;
constexpr int kTilesPerSample = (kPix + 127) / 128;    // 14: tiles never straddle two samples (the last one of a sample has 17 rows)

struct C1Params {
    const float* obs; const float* bias; float* out;
    int n_pix, Hin, crop, num_tiles, col_first_tile, write_col, B;
};

// cvt.rna.tf32.f32 for finite inputs (round to nearest, ties away from zero, on the 13 dropped bits) with two integer
// instructions instead of the conversion pipe (16 lanes / clk / SM: 113 conversions per pixel kept it half busy)
__device__ __forceinline__ float rna_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((unsigned long long)tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// One producer thread's share of an A row: (channel, ky) pairs [PAIR0, PAIR0 + NPAIR) = k in [3 PAIR0, 3 (PAIR0 + NPAIR)),
// zero padded to 48 values, stored as the 12 swizzled 16-byte units starting at unit UNIT0.
template <int PAIR0, int NPAIR, int UNIT0>
__device__ __forceinline__ void build_half_row(const float* __restrict__ src, size_t plane, int Hin, uint32_t srow, int r,
                                               bool own_third) {
    float a[48];
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
        const int c = (PAIR0 + j) / 3, ky = (PAIR0 + j) % 3;
        const float* pr = src + c * plane + ky * Hin;
        // pixel x needs columns 2x, 2x+1, 2x+2: the third is the next lane's first (the next pixel of the same output row);
        // only the last lane / the last pixel of a row loads it itself -- halves the L1 requests of the gather
        const float2 v01 = __ldg(reinterpret_cast<const float2*>(pr));
        float v2 = __shfl_down_sync(0xffffffffu, v01.x, 1);
        if (own_third) v2 = __ldg(pr + 2);
        a[3 * j] = v01.x; a[3 * j + 1] = v01.y; a[3 * j + 2] = v2;
    }
    // x / 255 correctly rounded without the division subroutine (~40 instructions, it made the producers the bottleneck):
    // q = x * rcp, r = x - 255 q exactly (FMA), q + r * rcp rounds to RN(x / 255) (Markstein; rcp = RN(1/255), 255's
    // significand is not all ones; no overflow / underflow for pixel-range inputs)
    const float rcp = 1.0f / 255.0f;
#pragma unroll
    for (int k = 0; k < 3 * NPAIR; ++k) {
        const float q0 = a[k] * rcp;
        const float rem = __fmaf_rn(-q0, 255.0f, a[k]);
        a[k] = rna_tf32(__fmaf_rn(rem, rcp, q0));
    }
#pragma unroll
    for (int k = 3 * NPAIR; k < 48; ++k) a[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const int u = UNIT0 + j;
        const uint32_t addr = srow + (u >> 3) * kChunkBytes + (((u & 7) ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a[4 * j]), "f"(a[4 * j + 1]), "f"(a[4 * j + 2]),
                     "f"(a[4 * j + 3]) : "memory");
    }
}

__global__ void __launch_bounds__(kThreads, 1)
conv1_fused_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmCol,
                      const __grid_constant__ CUtensorMap tmOut, C1Params p) {
    constexpr uint32_t kIdesc = idesc_tf32(32, false, false);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;
    const uint32_t w_sm = a_sm + kStages * kABytes;
    const uint32_t stg_sm = w_sm + kWBytes;                       // epilogue staging tile [128][32] (1024-byte aligned)
    const uint32_t bars = stg_sm + kChunkBytes;
    const uint32_t full0 = bars, empty0 = full0 + 8 * kStages, wbar = empty0 + 8 * kStages;
    const uint32_t tfull0 = wbar + 8, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmOut) : "memory");
        if (p.write_col) asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmCol) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 256); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(wbar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, kWBytes);
            for (int c = 0; c < 3; ++c) tma_load_2d(&tmW, wbar, w_sm + c * 4096, c * 32, 0);
            mbar_wait(wbar, 0);
            tc_fence_after();
            int i = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
                const int stage = i & (kStages - 1), acc = i & 1;
                mbar_wait(tempty0 + 8 * acc, ((uint32_t)(i >> 1) & 1u) ^ 1u);
                mbar_wait(full0 + 8 * stage, (uint32_t)(i / kStages) & 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint64_t ad = make_desc_sw128(a_sm + stage * kABytes + c * kChunkBytes);
                    const uint64_t bd = make_desc_sw128(w_sm + c * 4096);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_tf32(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc, (c | k) != 0);
                }
                tc_commit(empty0 + 8 * stage);
                tc_commit(tfull0 + 8 * acc);
            }
        }
    } else if (warp <= 4) {
        // ---- epilogue: TMEM lane quarter (warp & 3) -> bias, ReLU, TF32 round -> out[((b*43 + y)*41 + x)*32 ...]
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const bool e_elected = threadIdx.x == 32;
        float bv[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bv[c] = __ldg(p.bias + c);
        int i = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
            const int acc = i & 1;
            mbar_wait(tfull0 + 8 * acc, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * acc);
            float o[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = rna_tf32(fmaxf(__uint_as_float(v[c]) + bv[c], 0.f));
            // row-per-thread global stores touch 32 lines per instruction (3.4 M half-sector writes at 256 samples, a
            // quarter of the kernel): stage the tile in the swizzled layout and let TMA write whole rows.  The output
            // tensor is seen as [B][1681][32] with a sample stride of 43*41 rows; the last tile of a sample is clipped by
            // the tensor bounds.
            if (e_elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 4, 128;" ::: "memory");
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const uint32_t addr = stg_sm + row * 128 + ((c4 ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[4 * c4]), "f"(o[4 * c4 + 1]), "f"(o[4 * c4 + 2]),
                             "f"(o[4 * c4 + 3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 4, 128;" ::: "memory");
            if (e_elected) {
                const int b0 = tile / kTilesPerSample, p0 = (tile - b0 * kTilesPerSample) * 128;
                tma_store_3d(&tmOut, stg_sm, 0, p0, b0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (e_elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // ---- producers: group g = tiles i with i % 2 == g; threads (r, half) build the two halves of row r of the A tile
        const int g = (warp - 5) >> 3;
        const int pt = ((warp - 5) & 7) * 32 + lane;
        const int r = pt & 127, half = pt >> 7;
        const bool elected = pt == 0;
        const int Hin = p.Hin;
        const size_t plane = (size_t)Hin * Hin;
        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, i += 2) {
            const int stage = i & (kStages - 1);
            const uint32_t srow = a_sm + stage * kABytes + r * 128;
            mbar_wait(empty0 + 8 * stage, ((uint32_t)(i / kStages) & 1u) ^ 1u);
            if (p.write_col) {                       // the TMA store that read this stage two group-tiles ago must be done
                if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");
            }
            const int b = tile / kTilesPerSample, p0 = (tile - b * kTilesPerSample) * 128;
            {
                const int rem = min(p0 + r, kPix - 1);           // rows past the sample's end repeat its last pixel (never stored)
                const int y = rem / 41, x = rem - y * 41;
                const float* src = p.obs + (size_t)b * 9 * plane + (size_t)(2 * y + p.crop) * Hin + 2 * x + p.crop;
                const bool own_third = lane == 31 || x == 40;
                if (half == 0) build_half_row<0, 16, 0>(src, plane, Hin, srow, r, own_third);
                else build_half_row<16, 11, 12>(src, plane, Hin, srow, r, own_third);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full0 + 8 * stage);
            if (p.write_col) {
                asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");     // every row written and fenced
                if (elected) {
                    if (tile >= p.col_first_tile)
                        for (int c = 0; c < 3; ++c) tma_store_3d(&tmCol, a_sm + stage * kABytes + c * kChunkBytes, c * 32, p0, b);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (p.write_col && elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
    }
}

}  // namespace

// obs: (B, 9, Hin, Hin) fp32 NCHW with values 0..255 (Hin = 84, or 100 = centre crop, modules.py:70-83); w1p: TF32 operand
// copy [32][96] of the conv weight (sgqn_conv1_weights_prep); out: relu(conv1(obs / 255)) rounded to TF32, pitch-linear
// [B][43][41][32] (rows 41, 42 of every sample are never written).  col (optional): im2col matrix [B*1681][96] for the
// weight gradient, written for the samples >= col_row0 (earlier rows of col are not touched).
extern "C" int sgqn_conv1_fused_tc(const float* obs, const float* w1p, const float* bias, float* out, float* col, int B, int Hin,
                                   int col_row0, void* stream) {
    if (B <= 0) return 0;
    if (Hin < 84 || ((Hin - 84) & 3) || !bias) return (int)cudaErrorInvalidValue;
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv1_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    C1Params p;
    p.obs = obs; p.bias = bias; p.out = out;
    p.n_pix = B * kPix; p.Hin = Hin; p.crop = (Hin - 84) / 2;
    p.num_tiles = B * kTilesPerSample;
    p.write_col = col != nullptr && col_row0 < B;
    p.B = B;
    p.col_first_tile = p.write_col ? col_row0 * kTilesPerSample : p.num_tiles;
    CUtensorMap tmW, tmCol, tmOut;
    int rc = make_map_2d(&tmW, w1p, 96, 32, 32, 32);
    if (rc) return rc;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    auto map3 = [&](CUtensorMap* tm, const float* ptr, int cols, uint64_t sample_stride_bytes) -> int {      // [B][1681][cols], box {32, 128, 1}
        cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)kPix, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)cols * 4, sample_stride_bytes};
        cuuint32_t box[3] = {32, 128, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS ? 0 : 900 + (int)r;
    };
    rc = map3(&tmOut, out, 32, (uint64_t)43 * 41 * 128);
    if (rc) return rc;
    rc = map3(&tmCol, p.write_col ? col : out, p.write_col ? 96 : 32, p.write_col ? (uint64_t)kPix * 384 : (uint64_t)43 * 41 * 128);
    if (rc) return rc;
    int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    return launch_pdl(conv1_fused_tc_kernel, dim3(grid), dim3(kThreads), kSmem, stream, tmW, tmCol, tmOut, p);
}

// =====================================================================================================================
// Data gradient of the first conv to the observation (the last step of compute_attribution, rl_utils.py:35-39,57-62) as
// ONE kernel: dcol[pix][96] = d(act_0)[pix][32] * W1 on tcgen05, gathered back to the NCHW observation gradient through
// shared memory -- the 82.6 MB dcol matrix (B = 128) is never written to / re-read from HBM.
//
// Tile = output rows y0 = 2t, y0+1, y0+2 of one sample (123 pixels = one TMA box of d(act_0), rows past the sample's end are
// zero-filled by TMA: their dcol rows are exact zeros).  Observation row Y collects ky = 1 from output row (Y-1)/2 (Y odd) or
// ky = 0 from row Y/2 and ky = 2 from row Y/2 - 1 (Y even), so the tile owns Y = 2y0+1 .. 2y0+4 (and Y = 0 for t = 0): every
// observation pixel is written exactly once, no atomics.  Same summation order and exact /255 as conv1_col2im_rows_kernel.
namespace {

constexpr int kDgStages = 3;
constexpr int kDgABytes = 128 * 128;                 // d(act_0) tile: [128 px][32 ch]
constexpr int kDgWBytes = 96 * 128;                  // W1d: [96 k][32 co]
constexpr int kDgPitch = 97;                         // dcol staging pitch (floats): consecutive pixels hit consecutive banks
constexpr int kDgStgBytes = 123 * kDgPitch * 4;
constexpr int kDgSmem = kDgStages * kDgABytes + kDgWBytes + kDgStgBytes + 1024 + 256;
constexpr int kDgTilesPerSample = 21;

struct DgParams { float* dobs; int B, num_tiles; };

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(320, 2)
conv1_dgrad_fused_tc_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmW, DgParams p) {
    constexpr uint32_t kIdesc = idesc_tf32(96, false, false);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;
    const uint32_t w_sm = a_sm + kDgStages * kDgABytes;
    const uint32_t stg_off = (w_sm + kDgWBytes) - smem_u32(smem_raw);
    const uint32_t bars = w_sm + kDgWBytes + ((kDgStgBytes + 15) & ~15);
    const uint32_t full0 = bars, empty0 = full0 + 8 * kDgStages, wbar = empty0 + 8 * kDgStages;
    const uint32_t tfull0 = wbar + 8, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmD) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmW) : "memory");
        for (int s = 0; s < kDgStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(wbar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, kDgWBytes);
            tma_load_2d(&tmW, wbar, w_sm, 0, 0);
            int i = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
                const int stage = i % kDgStages;
                const int b = tile / kDgTilesPerSample, t = tile - b * kDgTilesPerSample;
                mbar_wait(empty0 + 8 * stage, ((uint32_t)(i / kDgStages) & 1u) ^ 1u);
                mbar_expect_tx(full0 + 8 * stage, kDgABytes);
                tma_load_3d(&tmD, full0 + 8 * stage, a_sm + stage * kDgABytes, 0, 2 * t * 41, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(wbar, 0);
            tc_fence_after();
            int i = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
                const int stage = i % kDgStages, acc = i & 1;
                mbar_wait(tempty0 + 8 * acc, ((uint32_t)(i >> 1) & 1u) ^ 1u);
                mbar_wait(full0 + 8 * stage, (uint32_t)(i / kDgStages) & 1u);
                tc_fence_after();
                const uint64_t ad = make_desc_sw128(a_sm + stage * kDgABytes);
                const uint64_t bd = make_desc_sw128(w_sm);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_tf32(tmem_base + (uint32_t)(acc * 128), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc, k != 0);
                tc_commit(empty0 + 8 * stage);
                tc_commit(tfull0 + 8 * acc);
            }
        }
    } else {
        const int et = threadIdx.x - 64;                 // 0..255
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        float* stg = reinterpret_cast<float*>(smem_raw + stg_off);
        const int xg = et / 84, X = et - xg * 84;        // gather role: column X of observation rows dealt to group xg (xg == 3: idle)
        const int x0 = X >> 1;
        const bool xodd = X & 1, v0 = x0 < 41, v1 = x0 >= 1;
        const int oA = x0 * kDgPitch + (xodd ? 1 : 0);   // X odd: tap kx = 1 of pixel x0; X even: tap kx = 0 of pixel x0 ...
        const int oB = (x0 - 1) * kDgPitch + 2;          // ... and tap kx = 2 of pixel x0 - 1
        int i = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
            const int acc = i & 1;
            const int b = tile / kDgTilesPerSample, t = tile - b * kDgTilesPerSample;
            mbar_wait(tfull0 + 8 * acc, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            // ---- drain: thread (row, half) moves 48 of the pixel's 96 dcol values to the staging tile
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128 + half * 48 + g * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < 123) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) stg[row * kDgPitch + half * 48 + g * 16 + e] = __uint_as_float(v[e]);
                }
            }
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * acc);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- gather: observation rows Y = 2 y0 + 1 .. 2 y0 + 4 (and Y = 0 for the first tile), 9 channels, 84 columns
            const int y0 = 2 * t;
            const int Ylo = t == 0 ? 0 : 2 * y0 + 1, Yhi = min(2 * y0 + 4, 83);
            const int nrows = Yhi - Ylo + 1;
            // three groups of 84 threads, one observation column X each (its tap offsets and bounds are loop invariants); the
            // (row Y, channel) pairs are dealt round-robin to the groups and unrolled for independent shared-memory loads in
            // flight -- a flat loop over the 3024 outputs with per-output index arithmetic was 80 % of the kernel
            if (xg < 3) {
                const int npairs = nrows * 9;
#pragma unroll 4
                for (int pr = xg; pr < npairs; pr += 3) {
                    const int yi = pr / 9, ci = pr - yi * 9, Y = Ylo + yi;
                    float a = 0.f;
                    if (Y & 1) {                             // ky = 1 from output row (Y-1)/2
                        const float* s0 = stg + (((Y - 1) >> 1) - y0) * 41 * kDgPitch + ci * 9 + 3;
                        if (v0) a += s0[oA];
                        if (!xodd && v1) a += s0[oB];
                    } else {                                 // ky = 0 from row Y/2, ky = 2 from row Y/2 - 1 (zero above the image)
                        const int r0 = (Y >> 1) - y0;
                        const float* s0 = stg + r0 * 41 * kDgPitch + ci * 9;
                        const float* s1 = s0 - 41 * kDgPitch + 6;
                        const bool up = r0 >= 1;             // r0 == 0 only for Y = 0 (t = 0): no row above
                        if (v0) a += s0[oA] + (up ? s1[oA] : 0.f);
                        if (!xodd && v1) a += s0[oB] + (up ? s1[oB] : 0.f);
                    }
                    const float rcp = 1.0f / 255.0f;         // exact a / 255 (see build_half_row)
                    const float q0 = a * rcp;
                    p.dobs[((size_t)(b * 9 + ci) * 84 + Y) * 84 + X] = __fmaf_rn(__fmaf_rn(-q0, 255.0f, a), rcp, q0);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile free again
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

}  // namespace

// d: d(act_0) compact [B][41][41][32] (TF32-rounded); w1d: [96][32] transposed TF32 operand copy of the conv weight
// (sgqn_conv1_weights_prep); dobs: (B, 9, 84, 84) fp32 gradient w.r.t. the 0..255 observation (includes NormalizeImg's 1/255),
// every element written (row / column 83 are structurally zero).
extern "C" int sgqn_conv1_dgrad_fused_tc(const float* d, const float* w1d, float* dobs, int B, void* stream) {
    if (B <= 0) return 0;
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv1_dgrad_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDgSmem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    DgParams p;
    p.dobs = dobs; p.B = B; p.num_tiles = B * kDgTilesPerSample;
    CUtensorMap tmD, tmW;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    {
        cuuint64_t dims[3] = {32, (cuuint64_t)kPix, (cuuint64_t)B};
        cuuint64_t strides[2] = {128, (cuuint64_t)kPix * 128};
        cuuint32_t box[3] = {32, 128, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 900 + (int)r;
    }
    int rc = make_map_2d(&tmW, w1d, 32, 96, 32, 96);
    if (rc) return rc;
    int grid = p.num_tiles < 2 * num_sms ? p.num_tiles : 2 * num_sms;
    return launch_pdl(conv1_dgrad_fused_tc_kernel, dim3(grid), dim3(320), kDgSmem, stream, tmD, tmW, p);
}
