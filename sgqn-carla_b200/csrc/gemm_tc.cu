// Dense layers of the heads (RLProjection, Q trunks, actor MLP -- modules.py:102-113,235-261,190-216) on tcgen05 with
// fp32-grade accuracy: "3xTF32" split precision.
//
// The reference runs nn.Linear in fp32 (torch.backends.cuda.matmul.allow_tf32 is False by default), so a plain TF32
// MMA would lose 13 mantissa bits.  Each fp32 operand x is split into big = tf32(x) and small = tf32(x - big)
// (x = big + small up to 2^-22 |x|), and the product is accumulated in fp32 TMEM as
//         A.B  ~=  A_small.B_big + A_big.B_small + A_big.B_big            (the small.small term is below fp32 eps)
// which is the error-compensated scheme cuBLAS / CUTLASS offer as fp32 emulation.  Measured against fp64 the result is
// as close as an fp32 FMA chain (tests/test_kernels_gpu.py::test_gemm_tc_*).
//
// One kernel serves forward, data gradient and weight gradient:  C[M][N] (+)= sum_k A[m][k] B[n][k],  where each
// operand is either "K-major" (stored [rows][k], k contiguous: SWIZZLE_128B tiles of 128 rows x 32 k) or "MN-major"
// (stored [k][rows], rows contiguous: SWIZZLE_128B_BASE32B atoms of 32 k x 32 rows) -- only the TMA box and the UMMA
// descriptor differ.  Pipeline per 32-wide k-block (3 stages of 64 KB):
//     warp 0      TMA producer: raw fp32 tiles of A (128 rows) and B (bn <= 128 rows)
//     warps 2-5   converter: in place big = tf32(relu?(x)), small copy next to it (elementwise, layout-agnostic),
//                 fence.proxy.async, arrive
//     warp 1      MMA issuer: 4 k-steps x 3 MMAs (128 x bn x 8) into one TMEM accumulator
//     warps 6-9   epilogue after the last k-block: TMEM -> registers -> bias / ReLU-mask -> store or red.add (split-K)
// Grid = (M tiles x N tiles, K splits, batch): the layers are small (M = 128..256 rows), so split-K spreads them over
// the 148 SMs; partial sums meet in global memory through vector red.add.
#include "tc_common.cuh"
#include "../../include/sgqn_b200.h"

using namespace tc;

namespace {

constexpr int kGtStages = 3;
constexpr int kGtTile = 128 * 128;                     // 16 KB: 128 rows x 32 fp32 (or 4 MN atoms of 32 k x 32 rows)
constexpr int kGtStage = 4 * kGtTile;                  // A | B | A_small | B_small
constexpr int kGtSmem = kGtStages * kGtStage + 1024 + 256;
constexpr int kGtThreads = 320;

struct GtParams {
    int M, N;                  // output extent
    int bn;                    // N tile: multiple of 32, <= 128
    int tiles_n;
    int a_mn, b_mn, relu_a, relu_b;
    int kb_total, kb_per_split;
    uint32_t idesc;
    float* c; int ldc; long long cbs;
    const float* bias; long long bbs;
    const float* mask; int ldm; long long mbs; int mode;      // mode 1: v = mask > 0 ? v : 0
    int atomic;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 32-row atoms along M / N
    d |= (uint64_t)(512 >> 4) << 32;                    // stride between 4-k atoms
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kGtThreads, 1)
gemm3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + kGtStages * kGtStage;
    const uint32_t full0 = bars, conv0 = bars + 8 * kGtStages, empty0 = bars + 16 * kGtStages, done_bar = bars + 24 * kGtStages;
    const uint32_t tmem_slot = done_bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x / p.tiles_n, tile_n = blockIdx.x - tile_m * p.tiles_n;
    const int m0 = tile_m * 128, n0 = tile_n * p.bn;
    const int split = blockIdx.y, batch = blockIdx.z;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
    const uint32_t b_bytes = (uint32_t)p.bn * 128u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmB) : "memory");
        for (int s = 0; s < kGtStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(conv0 + 8 * s, 128); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();                                        // operands / C come from the kernels before (common.cuh: PDL rules)
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const uint32_t sb = base + stage * kGtStage;
                const int k0 = kb * 32;
                mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                mbar_expect_tx(full0 + 8 * stage, kGtTile + b_bytes);
                if (p.a_mn) {
                    for (int a = 0; a < 4; ++a) tma_load_3d(&tmA, full0 + 8 * stage, sb + a * 4096, m0 + 32 * a, k0, batch);
                } else {
                    tma_load_3d(&tmA, full0 + 8 * stage, sb, k0, m0, batch);
                }
                if (p.b_mn) {
                    for (int a = 0; a < p.bn / 32; ++a) tma_load_3d(&tmB, full0 + 8 * stage, sb + kGtTile + a * 4096, n0 + 32 * a, k0, batch);
                } else {
                    tma_load_3d(&tmB, full0 + 8 * stage, sb + kGtTile, k0, n0, batch);
                }
                if (++stage == kGtStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            // descriptors of stage 0 and what one k-step / one stage adds to their address field (>> 4): nothing is rebuilt per MMA --
            // the issuing thread is on the critical path of these 4..8 k-block launches
            const uint64_t ab0 = p.a_mn ? make_desc_mn32(base, 4096) : make_desc_sw128(base);
            const uint64_t bb0 = p.b_mn ? make_desc_mn32(base + kGtTile, 4096) : make_desc_sw128(base + kGtTile);
            const uint64_t ka = p.a_mn ? 64u : 2u, kbi = p.b_mn ? 64u : 2u;
            constexpr uint64_t kSmall = (2 * kGtTile) >> 4, kStageInc = kGtStage >> 4;
            for (int kb = kb0; kb < kb1; ++kb) {
                const uint64_t ab_s = ab0 + (uint64_t)stage * kStageInc, bb_s = bb0 + (uint64_t)stage * kStageInc;
                mbar_wait(conv0 + 8 * stage, phase);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ab = ab_s + (uint64_t)k * ka, as = ab + kSmall;
                    const uint64_t bb = bb_s + (uint64_t)k * kbi, bs = bb + kSmall;
                    tc_mma_tf32(tmem_base, as, bb, p.idesc, (kb != kb0 || k != 0) ? 1u : 0u);     // small terms first
                    tc_mma_tf32_acc(tmem_base, ab, bs, p.idesc);
                    tc_mma_tf32_acc(tmem_base, ab, bb, p.idesc);
                }
                tc_commit(empty0 + 8 * stage);
                if (++stage == kGtStages) { stage = 0; phase ^= 1u; }
            }
            tc_commit(done_bar);
        }
    } else if (warp < 6) {
        // converter (then helps with the second epilogue phase): x -> (big, small) TF32 pair, elementwise on the raw tiles (whatever their swizzle)
        const int t = threadIdx.x - 64;
        const int nb4 = p.bn * 8;
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            float4* ta = reinterpret_cast<float4*>(gbase + stage * kGtStage);
            float4* tb = ta + kGtTile / 16;
            mbar_wait(full0 + 8 * stage, phase);
#pragma unroll 4
            for (int i = t; i < 1024; i += 128) {
                float4 v = ta[i];
                if (p.relu_a) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                const float4 b = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
                ta[i] = b;
                ta[i + 2 * kGtTile / 16] = make_float4(round_tf32(v.x - b.x), round_tf32(v.y - b.y), round_tf32(v.z - b.z), round_tf32(v.w - b.w));
            }
#pragma unroll 4
            for (int i = t; i < nb4; i += 128) {
                float4 v = tb[i];
                if (p.relu_b) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                const float4 b = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
                tb[i] = b;
                tb[i + 2 * kGtTile / 16] = make_float4(round_tf32(v.x - b.x), round_tf32(v.y - b.y), round_tf32(v.z - b.z), round_tf32(v.w - b.w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA
            mbar_arrive(conv0 + 8 * stage);
            if (++stage == kGtStages) { stage = 0; phase ^= 1u; }
        }
    } else if (kb1 > kb0) {
        // epilogue phase 1: TMEM (row per thread) -> shared staging tile (the operand stages are idle once done_bar fired) ->
        // row-per-warp, 4 columns per lane: bias / mask loads and the stores or red.adds are whole 512-byte rows
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        float* stg = reinterpret_cast<float*>(gbase);
        const int pitch = p.bn + 4;                         // floats; (bn + 4) / 4 odd => conflict-free float4 rows
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int g = 0; g < p.bn / 32; ++g) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 32);
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
            uint4* d = reinterpret_cast<uint4*>(stg + row * pitch + g * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) d[c] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
    }
    if (warp >= 2 && kb1 > kb0) {
        // epilogue phase 2 (converter + epilogue warps): one output row per warp instruction, 4 columns per lane -- bias
        // and mask loads and the stores / red.adds are whole 512-byte rows
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int ew = warp - 2;
        const float* stg = reinterpret_cast<const float*>(gbase);
        const int pitch = p.bn + 4;
        const int col = n0 + lane * 4;
        const bool col_ok = lane * 4 < p.bn && col < p.N;
        const float* bias = (p.bias && split == 0) ? p.bias + batch * p.bbs : nullptr;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias && col_ok)
            for (int e = 0; e < 4; ++e) if (col + e < p.N) bv[e] = __ldg(bias + col + e);
        float* cb = p.c + batch * p.cbs;
        const float* mb = p.mode ? p.mask + batch * p.mbs : nullptr;
        const bool vec = ((p.ldc & 3) == 0) && ((p.N & 3) == 0) && (((size_t)cb & 15) == 0) &&
                         (!p.mode || (((p.ldm & 3) == 0) && (((size_t)mb & 15) == 0)));
        const int rows = min(128, p.M - m0);
        for (int r0 = ew; r0 < rows; r0 += 64) {            // 8 rows per batch: their loads are issued together
            float4 a[8], mk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r0 + 8 * j;
                if (r < rows && col_ok) {
                    a[j] = *reinterpret_cast<const float4*>(stg + r * pitch + lane * 4);
                    if (p.mode) {
                        const float* mp = mb + (size_t)(m0 + r) * p.ldm + col;
                        mk[j] = vec ? __ldg(reinterpret_cast<const float4*>(mp))
                                    : make_float4(__ldg(mp), col + 1 < p.N ? __ldg(mp + 1) : 0.f, col + 2 < p.N ? __ldg(mp + 2) : 0.f,
                                                  col + 3 < p.N ? __ldg(mp + 3) : 0.f);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r0 + 8 * j;
                if (r >= rows || !col_ok) continue;
                float o[4] = {a[j].x + bv[0], a[j].y + bv[1], a[j].z + bv[2], a[j].w + bv[3]};
                if (p.mode) {
                    o[0] = mk[j].x > 0.f ? o[0] : 0.f; o[1] = mk[j].y > 0.f ? o[1] : 0.f;
                    o[2] = mk[j].z > 0.f ? o[2] : 0.f; o[3] = mk[j].w > 0.f ? o[3] : 0.f;
                }
                float* dst = cb + (size_t)(m0 + r) * p.ldc + col;
                if (vec) {
                    if (p.atomic) red_add_v4(dst, o[0], o[1], o[2], o[3]);
                    else *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (col + e < p.N) { if (p.atomic) atomicAdd(dst + e, o[e]); else dst[e] = o[e]; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

// 3-D tensor map of one operand: K-major = [batch][rows][ld] with the contraction index contiguous (box 32 k x box_rows);
// MN-major = [batch][k][ld] with the row index contiguous (box 32 rows x 32 k, BASE32B swizzle).
int make_operand_map(CUtensorMap* tm, const float* ptr, int rows, int kdim, int ld, long long bs, int batch, int mn, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    if ((ld & 3) || (bs & 3) || ((size_t)ptr & 15)) return (int)cudaErrorInvalidValue;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    dims[0] = (cuuint64_t)(mn ? rows : kdim);
    dims[1] = (cuuint64_t)(mn ? kdim : rows);
    dims[2] = (cuuint64_t)batch;
    strides[0] = (cuuint64_t)ld * 4;
    strides[1] = batch > 1 ? (cuuint64_t)bs * 4 : strides[0] * dims[1];
    box[0] = 32; box[1] = mn ? 32 : (cuuint32_t)box_rows; box[2] = 1;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 900 + (int)r;
}

struct Operand { const float* ptr; int ld; long long bs; int mn; int relu; };


// C[batch][M][N] (+)= sum_k A[m][k] B[n][k].  accumulate 1: add onto what C holds (red.add); 0: overwrite -- with split-K
// (max_split > 1 and enough k-blocks to fill the SMs) C is zero-filled here and the splits meet through red.add,
// otherwise plain stores.
int launch_gemm3(const Operand& A, const Operand& B, int M, int N, int K, int batch, float* c, int ldc, long long cbs,
                 const float* bias, long long bbs, const float* mask, int ldm, long long mbs, int mode, int accumulate, int max_split,
                 cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return 0;
    static int smem_set = 0, num_sms = 0;
    if (!smem_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGtSmem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        smem_set = 1;
    }
    GtParams p;
    p.M = M; p.N = N;
    int bn = N >= 128 ? 128 : (N + 31) / 32 * 32;      // multiple of 32 covers both operand layouts
    p.bn = bn;
    p.tiles_n = (N + bn - 1) / bn;
    int tiles_m = (M + 127) / 128;
    p.a_mn = A.mn; p.b_mn = B.mn; p.relu_a = A.relu; p.relu_b = B.relu;
    p.kb_total = (K + 31) / 32;
    long long tiles = (long long)tiles_m * p.tiles_n * batch;
    int nsplit = 1;
    if (max_split > 1) {
        nsplit = (int)((num_sms + tiles - 1) / tiles);
        if (nsplit > p.kb_total / 2) nsplit = p.kb_total / 2;       // at least two k-blocks per split
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
    }
    p.kb_per_split = (p.kb_total + nsplit - 1) / nsplit;
    nsplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.idesc = idesc_tf32(bn, A.mn != 0, B.mn != 0);
    p.c = c; p.ldc = ldc; p.cbs = cbs; p.bias = bias; p.bbs = bbs; p.mask = mask; p.ldm = ldm; p.mbs = mbs; p.mode = mode;
    p.atomic = (accumulate || nsplit > 1) ? 1 : 0;
    if (!accumulate && nsplit > 1) { int rc = zero2d(c, ldc, cbs, M, N, batch, st); if (rc) return rc; }
    CUtensorMap tmA, tmB;
    int rc = make_operand_map(&tmA, A.ptr, M, K, A.ld, A.bs, batch, A.mn, 128);
    if (rc) return rc;
    rc = make_operand_map(&tmB, B.ptr, N, K, B.ld, B.bs, batch, B.mn, bn);
    if (rc) return rc;
    dim3 grid(tiles_m * p.tiles_n, nsplit, batch);
    return launch_pdl(gemm3_tc_kernel, grid, dim3(kGtThreads), kGtSmem, st, tmA, tmB, p);
}

__global__ void guided_mask_kernel(float* __restrict__ dx, int lddx, long long dxbs, const float* __restrict__ z, int ldm, long long mbs,
                                   int M, int K) {
    pdl_wait();
    pdl_launch();
    const int b = blockIdx.z;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)M * K; i += (long long)gridDim.x * blockDim.x) {
        int m = (int)(i / K), k = (int)(i - (long long)m * K);
        float* d = dx + b * dxbs + (size_t)m * lddx + k;
        float v = fmaxf(*d, 0.f);
        *d = __ldg(z + b * mbs + (size_t)m * ldm + k) > 0.f ? v : 0.f;
    }
}

}  // namespace

// y[M,N] = act(x)[M,K] W[N,K]^T + b   (sgqn_linear_fwd semantics; splitk 0: plain store, 1: accumulate onto y, 2: zero + accumulate)
extern "C" int sgqn_linear_fwd_tc(const float* x, int ldx, long long xbs, const float* w, long long wbs, const float* bias,
                                  long long bbs, float* y, int ldy, long long ybs, int M, int N, int K, int relu_in, int batch,
                                  int splitk, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Operand A{x, ldx, xbs, 0, relu_in}, B{w, K, wbs, 0, 0};
    return launch_gemm3(A, B, M, N, K, batch, y, ldy, ybs, bias, bbs, nullptr, 0, 0, 0, splitk == 1 ? 1 : 0, splitk ? 64 : 1, st);
}

// dx[M,K] = dy[M,N] W[N,K], masked by zmask (mode 1: plain ReLU backward, 2: guided backprop); accumulate as in sgqn_linear_dgrad
extern "C" int sgqn_linear_dgrad_tc(const float* dy, int lddy, long long dybs, const float* w, long long wbs, const float* zmask,
                                    int ldm, long long mbs, float* dx, int lddx, long long dxbs, int M, int N, int K, int mode,
                                    int accumulate, int batch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 2 && accumulate == 1) return (int)cudaErrorInvalidValue;     // guided mask does not distribute over a sum
    Operand A{dy, lddy, dybs, 0, 0}, B{w, K, wbs, 1, 0};
    int rc = launch_gemm3(A, B, M, K, N, batch, dx, lddx, dxbs, nullptr, 0, zmask, ldm, mbs, mode == 1 ? 1 : 0, accumulate == 1 ? 1 : 0,
                          64, st);
    if (rc || mode != 2) return rc;
    long long n = (long long)M * K;
    { int rc_ = launch_pdl(guided_mask_kernel, dim3(dim3((unsigned)((n + 255) / 256), 1, batch)), dim3(256), 0, st, dx, lddx, dxbs, zmask, ldm, mbs, M, K); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// dw[N,K] += dy^T[N,M] act(x)[M,K];  db[N] += colsum(dy)
extern "C" int sgqn_linear_wgrad_tc(const float* x, int ldx, long long xbs, const float* dy, int lddy, long long dybs, float* dw,
                                    long long dwbs, float* db, long long dbbs, int M, int N, int K, int relu_in, int batch,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Operand A{dy, lddy, dybs, 1, 0}, B{x, ldx, xbs, 1, relu_in};
    int rc = launch_gemm3(A, B, N, K, M, batch, dw, K, dwbs, nullptr, 0, nullptr, 0, 0, 0, 1, 64, st);
    if (rc) return rc;
    if (db)
        for (int bi = 0; bi < batch; ++bi) {
            rc = sgqn_colsum(dy + bi * dybs, lddy, M, N, db + bi * dbbs, stream);
            if (rc) return rc;
        }
    return 0;
}
