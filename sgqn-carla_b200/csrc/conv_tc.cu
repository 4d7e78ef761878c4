// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution (32 -> 32 channels, stride 1, valid) for the SharedCNN layers
// 2..11 (modules.py:144-146): forward and data-gradient.  TF32 operands, fp32 accumulation in TMEM.
//
// Formulation ("pitch-linear implicit GEMM"): activations are NHWC fp32, so one pixel = 32 channels = 128 B = one
// SWIZZLE_128B row.  Number the output positions in INPUT pitch coordinates, q = (b*Hp + y)*Wp + x; then for filter
// tap (ky,kx) the A-operand row of output q is input row q + ky*Wp + kx -- for a tile of 128 consecutive q the A tile of
// each tap is 128 CONSECUTIVE input rows, i.e. one plain 2-D TMA box {32 ch, 128 rows}.  Positions with x >= Wp-2 or
// y >= Hp-2 are computed and dropped in the epilogue (5-17 % of the rows).
//
// N = 32 output channels is a bad tcgen05 shape: a 128x32x8 TF32 MMA retires every ~83 cycles whatever N is (operand
// fetch), so 9 taps x 4 k-steps cost ~3000 cycles per tile.  The kernel therefore widens N to 96 = (kx, cout): for
// each ky ONE MMA chain multiplies the tile shifted by ky*Wp rows with the three kx taps' weights side by side,
//     D_kx[r] = sum_ky A[q0 + r + ky*Wp] . W[ky][kx]          (12 MMAs of 128x96x8 instead of 36 of 128x32x8)
// and the kx shift moves to the epilogue:  out[q0 + r] = D_0[r] + D_1[r+1] + D_2[r+2]  -- row offsets in the
// shared-memory staging tile the epilogue goes through anyway to turn TMEM's row-per-thread layout into coalesced
// stores.  A tile of 128 A rows yields 126 outputs.  Two 96-column accumulators in TMEM overlap the epilogue with the
// next tile's MMAs.
// The data-gradient is the same kernel on a zero-bordered (pad 2) gradient buffer with flipped / transposed weights
// and a ReLU-mask (plain or guided, rl_utils.py:35-39) epilogue writing into the interior of the next padded buffer.
//
// Warp roles (576 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..17 = epilogue (TMEM -> shared staging -> kx shift + bias / ReLU / mask / TF32 round -> global).
#include "tc_common.cuh"
#include "../../include/sgqn_b200.h"

using namespace tc;

namespace {

constexpr int kMaxStages = 6;
constexpr int kTileM = 128;                    // A rows (TMEM lanes) per tile
constexpr int kTileOut = 126;                  // outputs per tile: rows r with r + 2 < 128
constexpr int kAccCols = 128;                  // TMEM column stride between the two accumulators (96 used)
constexpr int kStgPitch = 400;                 // staging row: 96 floats + 16 B, so that 8 consecutive rows hit 8 different bank groups
constexpr int kStgBytes = kTileM * kStgPitch;  // 50 KB
constexpr int kEpiBytes = kStgBytes + 2048 + 128;   // staging tile, rowinfo[2][128], bias
constexpr int kWBytes = 9 * 32 * 128;          // 36 KB: 9 taps x [32 n][32 k] fp32
constexpr int kSmemBudget = 200 * 1024;        // dynamic shared memory we ask for

struct TcParams {
    int total_q;            // B * Hr * Wp virtual output positions (input-pitch coordinates)
    int Hr, Wp;             // rows per sample and row pitch of the INPUT buffer
    int Hv, Wv;             // valid output extent: positions with y < Hv and x < Wv are written
    int shift;              // input row of tap (ky,kx) for position q = q + ky*Wp + kx + shift
    int Hq, Wq, oy, ox;     // output buffer: rows per sample, pitch, offset of output (0,0)
    int Hm, Wm;             // mask buffer: rows per sample, pitch (mask of output (y,x) at row y, col x)
    int num_tiles;
    int halo_rows, stage_bytes, stages;   // A stage = one halo tile: rows [q0+shift, q0+shift+halo_rows) serve all 9 taps
    const float* bias;
    const float* mask;
    float* out;
    float* dbias;            // optional: += per-channel sum of the written outputs (bias gradient of the layer below)
    int relu_out, round_out, mask_mode, early_w;
};

// kind::tf32, D fp32, A/B TF32 K-major, M = 128, N = 96 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

// kEpiWarps = 16 epilogue warps (forward and data gradient)
template <int kEpiWarps>
__global__ void __launch_bounds__(64 + kEpiWarps * 32, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, TcParams p) {
    constexpr int kEpiThreads = kEpiWarps * 32;
    constexpr int kChPerThread = 128 / kEpiWarps;  // phase-1 channels per thread: the warps of a TMEM quarter share its 32 channels
    constexpr int kItems = 1024 / kEpiThreads;     // phase-2 (row, 16-byte chunk) items per thread
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int kStages = p.stages;
    const uint32_t a_sm = base;
    const uint32_t w_sm = base + kStages * p.stage_bytes;
    const uint32_t bars = w_sm + kWBytes;                         // 8-byte mbarriers
    const uint32_t full0 = bars, empty0 = bars + 8 * kMaxStages, wbar = bars + 16 * kMaxStages;
    const uint32_t tfull0 = wbar + 8, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    const uint32_t stg_off = (bars - smem_u32(smem_raw)) + 256;   // byte offset of the epilogue staging area
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmW) : "memory");
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(wbar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiThreads); }
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
    // The 36 KB of weights are fetched BEFORE the grid-dependency wait, under the previous kernel's last tiles: the operand copy
    // was written by sgqn_conv_weights_prep at least two launches earlier on this stream (or on another stream, joined by an
    // event), so it is complete and visible by the time the previous kernel let this one start (contract, sgqn_b200.h).
    // Callers that cannot promise this leave flags bit 4 clear: the weights are then fetched after the wait.
    if (warp == 0 && lane == 0 && p.early_w) {
        mbar_expect_tx(wbar, kWBytes);
        for (int t = 0; t < 9; ++t) tma_load_2d(&tmW, wbar, w_sm + t * 4096, t * 32, 0);
    }
    // everything above overlapped the previous kernel's last tiles (PDL, common.cuh); activations / gradients from here on
    pdl_wait();
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            if (!p.early_w) {
                mbar_expect_tx(wbar, kWBytes);
                for (int t = 0; t < 9; ++t) tma_load_2d(&tmW, wbar, w_sm + t * 4096, t * 32, 0);
            }
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                // one halo tile per output tile: the 9 tap operands are row-shifted views of it (9x less L2->SM traffic)
                mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                mbar_expect_tx(full0 + 8 * stage, p.stage_bytes);
                tma_load_2d(&tmA, full0 + 8 * stage, a_sm + stage * p.stage_bytes, 0, tile * kTileOut + p.shift);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(wbar, 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                for (int ky = 0; ky < 3; ++ky) {
                    // filter row ky = the halo tile shifted by ky*Wp rows of 128 B; the 128B swizzle is a function of
                    // the absolute shared-memory address, so a row-shifted start address reads what TMA wrote.  B = the
                    // three kx taps of this ky: 96 rows (kx, cout) x 32 cin, contiguous in the resident weights.
                    const uint64_t ad = make_desc_sw128(a_sm + stage * p.stage_bytes + ky * p.Wp * 128);
                    const uint64_t bd = make_desc_sw128(w_sm + ky * 3 * 4096);
#pragma unroll
                    for (int k = 0; k < 4; ++k)          // advance 8 TF32 = 32 B inside the swizzle atom: +2 in the >>4 address field
                        tc_mma_tf32(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc, (ky | k) != 0);
                }
                tc_commit(empty0 + 8 * stage);           // smem slot free once these MMAs retire
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
                tc_commit(tfull0 + 8 * acc);             // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // 16 epilogue warps, two phases per tile around a shared-memory staging tile (row pitch 400 B: conflict-free):
        //  1. row-per-thread (the only way to read TMEM): warp w drains TMEM lane quarter (w & 3), 8 of the 32 channels
        //     of D_0 | D_1 | D_2 -> stg[row][96]; the accumulator is released right after;
        //  2. chunk-per-thread: thread t owns the 16-byte channel chunk t % 8 of rows t/8 + 64 j; it adds the three
        //     kx-shifted partial sums  out[r] = D_0[r] + D_1[r+1] + D_2[r+2]  (plain row offsets in shared memory),
        //     applies bias / ReLU / mask / TF32 rounding and stores -- 8 lanes cover one 128-byte pixel, so global
        //     stores and mask loads are whole lines (a row-per-thread store touches 32 lines per instruction).
        // Output / mask offsets of the tile's rows are computed once per row (rowinfo), one tile ahead, so that the
        // mask loads of a tile are in flight while its accumulator is still being computed.
        const int et = threadIdx.x - 64;
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which kChPerThread output channels (phase 1)
        const int row = quarter * 32 + lane;             // phase-1 row
        const int chunk = et & 7, rslot = et >> 3;       // phase-2 chunk / first row
        int acc = 0; uint32_t acc_phase = 0;
        const int HW = p.Hr * p.Wp;
        uint8_t* stg = smem_raw + stg_off;
        int2* rowinfo = reinterpret_cast<int2*>(smem_raw + stg_off + kStgBytes);          // [2][128]: {out, mask} offsets in float4
        const float4 bias4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        float csum[4] = {0.f, 0.f, 0.f, 0.f};
        auto fill_rowinfo = [&](int tile, int par) {
            if (half != 0) return;
            int2 info = make_int2(-1, 0);
            const int q = tile * kTileOut + row;
            if (tile < p.num_tiles && row < kTileOut && q < p.total_q) {
                const int b = q / HW; const int r2 = q - b * HW; const int y = r2 / p.Wp; const int x = r2 - y * p.Wp;
                if (y < p.Hv && x < p.Wv)
                    info = make_int2(((b * p.Hq + y + p.oy) * p.Wq + x + p.ox) * 8, ((b * p.Hm + y) * p.Wm + x) * 8);
            }
            rowinfo[par * 128 + row] = info;
        };
        fill_rowinfo(blockIdx.x, 0);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        int par = 0;
        const float4* mask4 = reinterpret_cast<const float4*>(p.mask);
        float4* out4 = reinterpret_cast<float4*>(p.out);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, par ^= 1) {
            // mask loads of this tile's four phase-2 items first: in flight during the accumulator wait
            int2 info[kItems];
            float4 mk[kItems];
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int r = rslot + (kEpiThreads / 8) * j;
                info[j] = r < kTileOut ? rowinfo[par * 128 + r] : make_int2(-1, 0);
                if (p.mask_mode && info[j].x >= 0) mk[j] = __ldg(mask4 + info[j].y + chunk);
            }
            fill_rowinfo(tile + gridDim.x, par ^ 1);
            // ---- phase 1: TMEM -> staging
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            {
                uint32_t v[3 * kChPerThread];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols + half * kChPerThread);
                static_assert(kChPerThread == 8 || kChPerThread == 16, "phase-1 TMEM loads: 8 or 16 channels per thread");
#define SGQN_TMEM_LD8(O, ADDR)                                                                                            \
                asm volatile(                                                                                            \
                    "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                      \
                    : "=r"(v[O + 0]), "=r"(v[O + 1]), "=r"(v[O + 2]), "=r"(v[O + 3]), "=r"(v[O + 4]), "=r"(v[O + 5]),    \
                      "=r"(v[O + 6]), "=r"(v[O + 7])                                                                     \
                    : "r"(ADDR) : "memory")
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int h = 0; h < kChPerThread / 8; ++h) SGQN_TMEM_LD8(kx * kChPerThread + 8 * h, taddr + (uint32_t)(32 * kx + 8 * h));
#undef SGQN_TMEM_LD8
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * acc);          // accumulator drained: the MMA warp may reuse it
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                uint4* srow = reinterpret_cast<uint4*>(stg + row * kStgPitch + half * (kChPerThread * 4));
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int c = 0; c < kChPerThread / 4; ++c)
                        srow[kx * 8 + c] = make_uint4(v[kx * kChPerThread + 4 * c], v[kx * kChPerThread + 4 * c + 1],
                                                      v[kx * kChPerThread + 4 * c + 2], v[kx * kChPerThread + 4 * c + 3]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            // ---- phase 2: shifted sum, epilogue math, coalesced store
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                if (info[j].x < 0) continue;
                const int r = rslot + (kEpiThreads / 8) * j;
                const float4 a0 = *reinterpret_cast<const float4*>(stg + r * kStgPitch + chunk * 16);
                const float4 a1 = *reinterpret_cast<const float4*>(stg + (r + 1) * kStgPitch + 128 + chunk * 16);
                const float4 a2 = *reinterpret_cast<const float4*>(stg + (r + 2) * kStgPitch + 256 + chunk * 16);
                float o[4] = {a0.x + a1.x + a2.x + bias4.x, a0.y + a1.y + a2.y + bias4.y, a0.z + a1.z + a2.z + bias4.z,
                              a0.w + a1.w + a2.w + bias4.w};
                if (p.relu_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                if (p.mask_mode) {
                    const float mm[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (p.mask_mode == 2) o[e] = fmaxf(o[e], 0.f);
                        o[e] = mm[e] > 0.f ? o[e] : 0.f;
                    }
                }
                if (p.round_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = round_tf32(o[e]);
                }
                out4[info[j].x + chunk] = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) csum[e] += o[e];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // staging tile and rowinfo[par] are free again
        }
        if (p.dbias) {
            // lanes with equal chunk hold the same channels: 2 shuffles, then across the warps through the (now free) staging
            // tile, then ONE atomic per channel per CTA (per-lane atomics from every warp of every CTA onto the same 32
            // addresses serialised in L2 and cost ~10 us per launch)
            float* red = reinterpret_cast<float*>(stg);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = csum[e];
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                if (lane < 8) red[(warp - 2) * 32 + chunk * 4 + e] = t;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (et < 32) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < kEpiWarps; ++w) t += red[w * 32 + et];
                atomicAdd(p.dbias + et, t);
            }
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

// x: [B][Hr][Wp][32]; w: [32 n][9 taps][32 k] (TF32-rounded copy).  Output (b,y,x), y < Hv, x < Wv, is the 3x3 window sum
// over input rows q + ky*Wp + kx + shift (q = (b*Hr + y)*Wp + x) and goes to out[((b*Hq + y+oy)*Wq + x+ox)*32].
// flags: bit0 ReLU on the output, bit1 round the output to TF32 (it feeds another TF32 conv), bits 2-3 mask mode
// (1 plain ReLU backward, 2 guided) with the mask of (b,y,x) at mask[((b*Hm + y)*Wm + x)*32].
// dbias (optional): dbias[32] += sum over the written outputs (atomic) -- in the data-gradient chain this is the bias
// gradient of the layer below, for free.
extern "C" int sgqn_conv_tc(const float* x, const float* w, const float* bias, const float* mask, float* out, float* dbias, int B,
                            int Hr, int Wp, int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm, int Wm, int flags,
                            void* stream) {
    if (B <= 0) return 0;
    static int smem_set = 0;
    static int num_sms = 0;
    if (!smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        smem_set = 1;
    }
    TcParams p;
    p.total_q = B * Hr * Wp; p.Hr = Hr; p.Wp = Wp; p.Hv = Hv; p.Wv = Wv; p.shift = shift;
    p.Hq = Hq; p.Wq = Wq; p.oy = oy; p.ox = ox; p.Hm = Hm; p.Wm = Wm;
    p.num_tiles = (p.total_q + kTileOut - 1) / kTileOut;
    p.bias = bias; p.mask = mask; p.out = out; p.dbias = dbias;
    p.relu_out = flags & 1; p.round_out = (flags >> 1) & 1; p.mask_mode = (flags >> 2) & 3; p.early_w = (flags >> 4) & 1;
    if (p.mask_mode && !mask) return (int)cudaErrorInvalidValue;
    p.halo_rows = (kTileM + 2 * Wp + 7) / 8 * 8;                   // whole 1024-byte swizzle atoms
    if (p.halo_rows > 256) return (int)cudaErrorInvalidValue;      // TMA box limit
    p.stage_bytes = p.halo_rows * 128;
    p.stages = (kSmemBudget - kWBytes - 1024 - 256 - kEpiBytes) / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) return (int)cudaErrorInvalidValue;
    CUtensorMap tmA, tmW;
    int rc = make_map_2d(&tmA, x, 32, (uint64_t)p.total_q, 32, (uint32_t)p.halo_rows);
    if (rc) return rc;
    rc = make_map_2d(&tmW, w, 288, 32, 32, 32);
    if (rc) return rc;
    int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    // 16 epilogue warps for the masked data gradient too: 8 were better only while the bias-gradient atomics dominated its
    // epilogue (41.5 -> 37.0 us at 256 samples, 41x41)
    return launch_pdl(conv3x3_tc_kernel<16>, dim3(grid), dim3(64 + 16 * 32), kSmemBudget, stream, tmA, tmW, p);
}

// TF32-rounded operand copies of the 32->32 conv weights, refreshed after every optimiser step that touches them:
// wf[l][n=co][t][k=ci] = rna(w[l][co][t][ci]) (forward), wd[l][n=ci][t][k=co] = rna(w[l][co][8-t][ci]) (data gradient:
// flipped taps, transposed channels).  `lstride` = distance between consecutive layers' weights in `w` (floats).
__global__ void conv_weights_prep_kernel(const float* __restrict__ w, long long lstride, float* __restrict__ wf,
                                         float* __restrict__ wd, int n_layers) {
    pdl_wait();
    pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_layers * 9216) return;
    int l = i / 9216, r = i - l * 9216;
    int co = r / 288, t = (r / 32) % 9, ci = r & 31;
    float v = round_tf32(w[(size_t)l * lstride + r]);
    wf[i] = v;
    wd[(size_t)l * 9216 + ci * 288 + (8 - t) * 32 + co] = v;
}
extern "C" int sgqn_conv_weights_prep(const float* w, long long lstride, float* wf, float* wd, int n_layers, void* stream) {
    int n = n_layers * 9216;
    if (n <= 0) return 0;
    { int rc_ = launch_pdl(conv_weights_prep_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, w, lstride, wf, wd, n_layers); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// compact [B][H][W][C] -> rows [oy, oy+H), cols [ox, ox+W) of a zero-initialised [B][Hq][Wq][C] buffer
__global__ void pad_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int H, int W, int C4, int Hq, int Wq,
                                int oy, int ox, int round_out, long long total) {
    pdl_wait();
    pdl_launch();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % C4); long long t = i / C4;
    int x = (int)(t % W); t /= W; int y = (int)(t % H); int b = (int)(t / H);
    float4 v = __ldg(src + i);
    if (round_out & 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (round_out & 1) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    dst[(((size_t)b * Hq + y + oy) * Wq + x + ox) * C4 + c] = v;
}
extern "C" int sgqn_pad_copy(const float* src, float* dst, int B, int H, int W, int C, int Hq, int Wq, int oy, int ox,
                             int round_out, void* stream) {
    if (C & 3) return (int)cudaErrorInvalidValue;
    long long total = (long long)B * H * W * (C / 4);
    if (total <= 0) return 0;
    { int rc_ = launch_pdl(pad_copy_kernel, dim3((unsigned)cdivll(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)src, (float4*)dst, H, W,
                                                                                   C / 4, Hq, Wq, oy, ox, round_out, total); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// =====================================================================================================================
// Weight gradient of the 32->32 convs on tcgen05:  dW[co][ky][kx][ci] += sum_q dY[q][co] * X[q + (ky-2)*Wp + kx][ci].
// X (post-ReLU activations of the previous layer) and dY (zero-bordered gradient of this layer's output) share one
// pitch-linear geometry [B][Hr][Wp][32].  Substituting q' = q + (ky-2)*Wp moves the ky shift onto dY:
//         dW[co][ky][kx][ci] = sum_q' X[q' + kx][ci] * dY[q' + (2-ky)*Wp][co]
// so ONE MMA chain per block of 128 consecutive q' produces all nine taps:  D[M = (kx, ci)][N = (2-ky, co)], reduction
// over pixels.  Both operands are "MN-major" (the pixel index runs over the 128-byte rows of the tile, channels are
// contiguous: SWIZZLE_128B_BASE32B, the only MN-major layout for TF32, UMMA_K = 8 pixels = two 512-byte atoms) and both
// are HALO tiles loaded once per block: the four 32-row M atoms are the X tile shifted by kx = 0..3 rows (leading byte
// offset 128 B; kx = 3 is computed and dropped), the three 32-column N atoms are the dY tile shifted by 0, Wp, 2Wp
// rows (leading byte offset Wp*128 B).  Per 128 pixels the kernel moves 45 KB through shared memory (1.4x the
// operands) and issues 16 MMAs of 128x96x8.  Each CTA reduces a contiguous range of pixel blocks in TMEM and adds its
// 9216 partial sums to dW with red.global.add.f32 once at the end.
namespace {

constexpr int kWgRows = 128;                           // pixels per k-block
constexpr int kWgXRows = kWgRows + 8;                  // X halo: + kx <= 3, whole 8-row swizzle atoms
constexpr int kWgXBytes = kWgXRows * 128;
constexpr int kWgMaxStages = 4;
constexpr int kWgSmemBudget = 200 * 1024;
// kind::tf32, D fp32, M = 128, N = 96, A and B MN-major
constexpr uint32_t kIdescMN = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

// MN-major TF32 operands only exist in the SWIZZLE_128B_BASE32B layout (cute::UMMA::Layout_MN_SW128_32B_Atom: rows of
// 128 B = 32 channels, 32-byte chunks XOR-ed with row % 4, K atom = 4 rows = 512 B); TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The swizzle is a function of the absolute shared-memory address, so atoms that
// start at any multiple of 128 B (row-shifted views of one tile) read what TMA wrote.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 32-channel atoms along M / N
    d |= (uint64_t)(512 >> 4) << 32;                    // stride between 4-pixel atoms along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
    return d;
}

struct WgParams { int total_q, Wp, kb_total, kb_per_cta, drows, stage_bytes, stages; float* dw; };

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + p.stages * p.stage_bytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * kWgMaxStages, done_bar = bars + 16 * kWgMaxStages, tmem_slot = done_bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = blockIdx.x * p.kb_per_cta;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_cta);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmD) : "memory");
        for (int s = 0; s < kWgMaxStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
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
    pdl_wait();
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int q0 = kb * kWgRows;
                const uint32_t sb = base + stage * p.stage_bytes;
                mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                mbar_expect_tx(full0 + 8 * stage, p.stage_bytes);
                tma_load_2d(&tmX, full0 + 8 * stage, sb, 0, q0);                    // X rows q0 .. q0 + 135
                tma_load_2d(&tmD, full0 + 8 * stage, sb + kWgXBytes, 0, q0);        // dY rows q0 .. q0 + 127 + 2 Wp
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t lbo_d = (uint32_t)p.Wp * 128u;
            for (int kb = kb0; kb < kb1; ++kb) {
                const uint32_t sb = base + stage * p.stage_bytes;
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint64_t ad0 = make_desc_mn_sw128(sb, 128);                               // atoms: kx = 0..3 row shifts
                const uint64_t bd0 = make_desc_mn_sw128(sb + kWgXBytes, lbo_d);                 // atoms: 0, Wp, 2 Wp row shifts
                const uint32_t first = kb != kb0 ? 1u : 0u;
#pragma unroll
                for (int j = 0; j < kWgRows / 8; ++j)            // k-step j: +1024 bytes = +64 in the descriptors' address field
                    tc_mma_tf32(tmem_base, ad0 + (uint64_t)(j * 64), bd0 + (uint64_t)(j * 64), kIdescMN, j ? 1u : first);
                tc_commit(empty0 + 8 * stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            tc_commit(done_bar);
        }
    } else if (kb1 > kb0) {
        const int kx = warp & 3;                           // TMEM lane quarter = rows (kx, ci = lane)
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int g = 0; g < 3; ++g) {                      // column group g = dY shifted by g*Wp rows = filter row ky = 2 - g
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(kx * 32) << 16) + (uint32_t)(g * 32);
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
            if (kx < 3) {
                float* dst = p.dw + ((2 - g) * 3 + kx) * 32 + lane;
#pragma unroll
                for (int co = 0; co < 32; ++co) atomicAdd(dst + co * 288, __uint_as_float(v[co]));
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

}  // namespace

// x, dy: [B][Hr][Wp][32] (same geometry; dy zero outside its valid region, x finite everywhere);
// dw[32 co][9][32 ci] += sum_q dy[q][co] * x[q + (ky-2)*Wp + kx][ci]   (atomic; caller zero-fills)
extern "C" int sgqn_conv_wgrad_tc(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, void* stream) {
    if (B <= 0) return 0;
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBudget);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    WgParams p;
    p.total_q = B * Hr * Wp; p.Wp = Wp; p.dw = dw;
    p.drows = (kWgRows + 2 * Wp + 7) / 8 * 8;                       // dY halo, whole swizzle atoms
    if (p.drows > 256) return (int)cudaErrorInvalidValue;           // TMA box limit
    p.stage_bytes = kWgXBytes + p.drows * 128;
    p.stages = (kWgSmemBudget - 1024 - 256) / p.stage_bytes;
    if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
    if (p.stages < 2) return (int)cudaErrorInvalidValue;
    p.kb_total = (p.total_q + kWgRows - 1) / kWgRows;
    int grid = p.kb_total < num_sms ? p.kb_total : num_sms;
    p.kb_per_cta = (p.kb_total + grid - 1) / grid;
    grid = (p.kb_total + p.kb_per_cta - 1) / p.kb_per_cta;
    CUtensorMap tmX, tmD;
    int rc = make_map_2d(&tmX, x, 32, (uint64_t)p.total_q, 32, kWgXRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = make_map_2d(&tmD, dy, 32, (uint64_t)p.total_q, 32, (uint32_t)p.drows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    return launch_pdl(conv3x3_wgrad_tc_kernel, dim3(grid), dim3(192), kWgSmemBudget, stream, tmX, tmD, p);
}
