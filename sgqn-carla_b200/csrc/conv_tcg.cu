// Generalised tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution: Cin = 32*KC input channels, N = Cout in {32, 64, 128},
// stride 1, TF32 operands, fp32 accumulation.  Used for the AttributionDecoder convs (modules.py:319-326), forward and
// data gradient: conv1 32->128 @21x21, conv2 128->64 @42x42 (input = nearest x2 upsample of relu(conv1)), conv3 64->9(32)
// @84x84 (input = nearest x2 upsample of relu(conv2)); padding 1 is a zero border in the pitch-linear input buffer.
//
// Same pitch-linear formulation as conv_tc.cu (one 2-D TMA halo tile per 128 consecutive output positions serves all 9
// taps through row-shifted UMMA descriptors), with two differences: the K loop runs over channel chunks of 32 (one halo
// tile per chunk, ring "A"), and the weights do not fit in shared memory, so the [N][32] weight tile of every
// (chunk, tap) streams through its own TMA ring "W".  The epilogue can scatter every output pixel to the 2x2 block of
// the next layer's zero-bordered input (ReLU + nearest upsample + TF32 rounding fused into the producer).
#include <type_traits>

#include "tc_common.cuh"
#include "../../include/sgqn_b200.h"

using namespace tc;

extern "C" int sgqn_conv_tcg_taps(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int,
                                  int, int, int, int, int, int, int, int, void*);
extern "C" int sgqn_gemm_wgrad_tcg(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, void*);

namespace {

constexpr int kTileM = 128;
constexpr int kSmemBudget = 208 * 1024;
constexpr int kMaxA = 4, kMaxW = 8;

struct GParams {
    int total_q, Hr, Wp, Hv, Wv, shift;
    int Hq, Wq, oy, ox, Hm, Wm;
    int num_tiles, kc, cin, ntaps;
    int a_bytes, piece_rows, pieces, a_stages, w_stages;
    const float* bias;
    const float* mask;
    float* out;
    int relu_out, round_out, mask_mode, upsample;
    int pair, num_iters;    // pair: TWO position tiles (M = 256, two accumulators) per pass over the weight tiles; iterations = tiles / (1 + pair)
    int w_resident;         // every (chunk, tap) weight tile stays in shared memory for the whole launch (<= ~150 KB of weights)
    int shuffle;            // 0 none, 1 depth-to-space (phase channels -> 2x2 pixels), 2 space-to-depth (pixel -> phase channels)
};

template <int N, int T>
__global__ void __launch_bounds__(320, 1)
conv3x3_tcg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, GParams p) {
    constexpr int kWBytes = N * 128;
    constexpr uint32_t kIdescN = idesc_tf32(N, false, false);
    // accumulators: T position tiles per iteration (T = 2: "pair" mode), double-buffered while 2 T N columns fit the 512 of TMEM
    constexpr int nbuf = 2 * T * N <= 512 ? 2 : 1;
    constexpr int acc_cols = nbuf * T * N;
    constexpr uint32_t kTmemCols = acc_cols <= 64 ? 64 : (acc_cols <= 128 ? 128 : (acc_cols <= 256 ? 256 : 512));
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;
    const uint32_t w_sm = a_sm + p.a_stages * p.a_bytes;
    const uint32_t bias_sm = w_sm + p.w_stages * kWBytes;
    const uint32_t bars = bias_sm + N * 4;
    const uint32_t afull0 = bars, aempty0 = afull0 + 8 * kMaxA, wfull0 = aempty0 + 8 * kMaxA, wempty0 = wfull0 + 8 * kMaxW;
    const uint32_t tfull0 = wempty0 + 8 * kMaxW, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmW) : "memory");
        for (int s = 0; s < kMaxA; ++s) { mbar_init(afull0 + 8 * s, 1); mbar_init(aempty0 + 8 * s, 1); }
        for (int s = 0; s < kMaxW; ++s) { mbar_init(wfull0 + 8 * s, 1); mbar_init(wempty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();                                        // (the bias below is the first read of memory an earlier kernel wrote)
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < N; i += 256) {
            float bv = p.bias ? __ldg(p.bias + i) : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_sm + 4 * i), "f"(bv) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_launch();

    if (warp == 0) {
        if (lane == 0) {
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0;
            if (p.w_resident) {                              // the whole weight block once (re-streaming it per tile made the small
                mbar_expect_tx(wfull0, p.kc * p.ntaps * kWBytes);   // decoder convs L2-bound: 147 KB of weights per 28 KB halo tile)
                for (int c = 0; c < p.kc; ++c)
                    for (int t = 0; t < p.ntaps; ++t)
                        tma_load_2d(&tmW, wfull0, w_sm + (c * p.ntaps + t) * kWBytes, t * p.cin + c * 32, 0);
            }
            for (int it = blockIdx.x; it < p.num_iters; it += gridDim.x) {
                const int r0 = it * T * kTileM + p.shift;
                for (int c = 0; c < p.kc; ++c) {
                    mbar_wait(aempty0 + 8 * as, aph ^ 1u);
                    mbar_expect_tx(afull0 + 8 * as, p.a_bytes);
                    for (int pc = 0; pc < p.pieces; ++pc)
                        tma_load_2d(&tmA, afull0 + 8 * as, a_sm + as * p.a_bytes + pc * p.piece_rows * 128, c * 32, r0 + pc * p.piece_rows);
                    if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                    if (p.w_resident) continue;
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(wempty0 + 8 * ws, wph ^ 1u);
                        mbar_expect_tx(wfull0 + 8 * ws, kWBytes);
                        tma_load_2d(&tmW, wfull0 + 8 * ws, w_sm + ws * kWBytes, t * p.cin + c * 32, 0);
                        if (++ws == p.w_stages) { ws = 0; wph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0; int acc = 0; uint32_t acc_phase = 0;
            if (p.w_resident) { mbar_wait(wfull0, 0); tc_fence_after(); }
            for (int it = blockIdx.x; it < p.num_iters; it += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T * N);
                for (int c = 0; c < p.kc; ++c) {
                    mbar_wait(afull0 + 8 * as, aph);
                    tc_fence_after();
                    const uint64_t a0 = make_desc_sw128(a_sm + as * p.a_bytes);
                    const uint32_t rowq = (uint32_t)p.Wp * 8u;
                    if (p.ntaps == 9) {
                        // the nine taps as straight-line code, one copy per weight mode: the issuing thread, not the tensor pipe, paces
                        // 36-cycle MMAs when every tap costs a loop trip with two divisions and a mode test (conv3 forward 70 -> 44 us,
                        // conv1 data gradient 40 -> 26 us, conv2 data gradient 94 -> 86 us)
                        const uint64_t b0 = make_desc_sw128(w_sm + c * 9 * kWBytes);
                        auto taps = [&](auto resident_tag) {
                            constexpr bool kResident = decltype(resident_tag)::value;
#pragma unroll
                            for (int t = 0; t < 9; ++t) {
                                uint64_t bd;
                                if constexpr (kResident) {
                                    bd = b0 + (uint64_t)(t * (kWBytes >> 4));
                                } else {
                                    mbar_wait(wfull0 + 8 * ws, wph);
                                    tc_fence_after();
                                    bd = make_desc_sw128(w_sm + ws * kWBytes);
                                }
                                const uint64_t ad = a0 + (uint64_t)((uint32_t)(t / 3) * rowq + (uint32_t)(t % 3) * 8u);
#pragma unroll
                                for (int h = 0; h < T; ++h) {    // the weight tile serves every position tile of the iteration
                                    const uint64_t adh = ad + (uint64_t)(h * (kTileM * 128 >> 4));
                                    const uint32_t dh = d_tmem + (uint32_t)(h * N);
                                    if (t == 0 && c == 0) tc_mma_tf32_zero(dh, adh, bd, kIdescN);
                                    else tc_mma_tf32_acc(dh, adh, bd, kIdescN);
#pragma unroll
                                    for (int k = 1; k < 4; ++k) tc_mma_tf32_acc(dh, adh + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdescN);
                                }
                                if constexpr (!kResident) {
                                    tc_commit(wempty0 + 8 * ws);
                                    if (++ws == p.w_stages) { ws = 0; wph ^= 1u; }
                                }
                            }
                        };
                        if (p.w_resident) taps(std::true_type{});
                        else taps(std::false_type{});
                        tc_commit(aempty0 + 8 * as);
                        if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                        continue;
                    }
                    for (int t = 0; t < p.ntaps; ++t) {
                        if (!p.w_resident) {
                            mbar_wait(wfull0 + 8 * ws, wph);
                            tc_fence_after();
                        }
                        const uint64_t ad = a0 + (uint64_t)(p.ntaps == 9 ? (t / 3) * rowq + (t % 3) * 8u : 0u);
                        const uint64_t bd = make_desc_sw128(w_sm + (p.w_resident ? c * p.ntaps + t : ws) * kWBytes);
#pragma unroll
                        for (int h = 0; h < T; ++h) {            // the weight tile serves every position tile of the iteration
                            const uint64_t adh = ad + (uint64_t)(h * (kTileM * 128 >> 4));
                            const uint32_t dh = d_tmem + (uint32_t)(h * N);
                            if ((c | t) == 0) tc_mma_tf32_zero(dh, adh, bd, kIdescN);
                            else tc_mma_tf32_acc(dh, adh, bd, kIdescN);
#pragma unroll
                            for (int k = 1; k < 4; ++k) tc_mma_tf32_acc(dh, adh + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdescN);
                        }
                        if (p.w_resident) continue;
                        tc_commit(wempty0 + 8 * ws);
                        if (++ws == p.w_stages) { ws = 0; wph ^= 1u; }
                    }
                    tc_commit(aempty0 + 8 * as);
                    if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                }
                tc_commit(tfull0 + 8 * acc);
                if (++acc == nbuf) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        if constexpr (N > 128) {
            // N = 256 (phase convs, 4 x 64 channels): 128 accumulator columns per thread, drained 32 at a time.  With
            // depth-to-space (p.shuffle == 1) channel block ph = 2a + b' goes to pixel (2y+a, 2x+b') of a 64-channel output.
            const int quarter = warp & 3;
            const int half = (warp - 2) >> 2;
            const int row = quarter * 32 + lane;
            int acc = 0; uint32_t acc_phase = 0;
            const int HW = p.Hr * p.Wp;
            constexpr int NG = N / 4;                    // channels per phase
            for (int it = blockIdx.x; it < p.num_iters; it += gridDim.x) {
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
#pragma unroll 1
              for (int h = 0; h < T; ++h) {
                const int q = (it * T + h) * kTileM + row;
                const int b = q / HW; const int r2 = q - b * HW; const int y = r2 / p.Wp; const int x = r2 - y * p.Wp;
                const bool valid = q < p.total_q && y < p.Hv && x < p.Wv;
#pragma unroll 1
                for (int g = 0; g < N / 2 / 32; ++g) {
                    const int c0 = half * (N / 2) + g * 32;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * T + h) * N + c0);
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
                    if (!valid) continue;
                    float* dst;
                    if (p.shuffle == 1) {
                        const int ph = c0 / NG, cc = c0 - ph * NG;
                        dst = p.out + ((size_t)(b * p.Hq + 2 * y + (ph >> 1) + p.oy) * p.Wq + 2 * x + (ph & 1) + p.ox) * NG + cc;
                    } else {
                        dst = p.out + ((size_t)(b * p.Hq + y + p.oy) * p.Wq + x + p.ox) * N + c0;
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float bv;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_sm + 4 * (c0 + 4 * c4 + e)));
                            float f = __uint_as_float(v[4 * c4 + e]) + bv;
                            if (p.relu_out) f = fmaxf(f, 0.f);
                            if (p.round_out) f = round_tf32(f);
                            o[e] = f;
                        }
                        reinterpret_cast<float4*>(dst)[c4] = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
              }
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * acc);
                if (++acc == nbuf) { acc = 0; acc_phase ^= 1u; }
            }
        } else {
        constexpr int CPW = N / 2;                       // accumulator columns per epilogue warp
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        const int HW = p.Hr * p.Wp;
        for (int it = blockIdx.x; it < p.num_iters; it += gridDim.x) {
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
#pragma unroll
          for (int h = 0; h < T; ++h) {
            const int q = (it * T + h) * kTileM + row;
            const int b = q / HW; const int r2 = q - b * HW; const int y = r2 / p.Wp; const int x = r2 - y * p.Wp;
            const bool valid = q < p.total_q && y < p.Hv && x < p.Wv;
            uint32_t v[CPW];
#pragma unroll
            for (int g = 0; g < CPW / 16; ++g) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * T + h) * N + half * CPW + g * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[16 * g + 0]), "=r"(v[16 * g + 1]), "=r"(v[16 * g + 2]), "=r"(v[16 * g + 3]), "=r"(v[16 * g + 4]),
                      "=r"(v[16 * g + 5]), "=r"(v[16 * g + 6]), "=r"(v[16 * g + 7]), "=r"(v[16 * g + 8]), "=r"(v[16 * g + 9]),
                      "=r"(v[16 * g + 10]), "=r"(v[16 * g + 11]), "=r"(v[16 * g + 12]), "=r"(v[16 * g + 13]), "=r"(v[16 * g + 14]),
                      "=r"(v[16 * g + 15])
                    : "r"(taddr) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (h == T - 1) {                              // every accumulator of the iteration is in registers: release the set
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * acc);
            }
            if (!valid) continue;
            const int c0 = half * CPW;
            const float* mk = p.mask_mode ? p.mask + ((size_t)(b * p.Hm + y) * p.Wm + x) * N + c0 : nullptr;
            float* dst0;
            if (p.upsample) dst0 = p.out + ((size_t)(b * p.Hq + 2 * y + p.oy) * p.Wq + 2 * x + p.ox) * N + c0;
            else if (p.shuffle == 2)      // space-to-depth: pixel (y,x) -> low-res pixel (y/2, x/2), channel block 2(y&1) + (x&1) of 4N
                dst0 = p.out + ((size_t)(b * p.Hq + (y >> 1) + p.oy) * p.Wq + (x >> 1) + p.ox) * (4 * N) + (2 * (y & 1) + (x & 1)) * N + c0;
            else dst0 = p.out + ((size_t)(b * p.Hq + y + p.oy) * p.Wq + x + p.ox) * N + c0;
#pragma unroll
            for (int c4 = 0; c4 < CPW / 4; ++c4) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float bv;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_sm + 4 * (c0 + 4 * c4 + e)));
                    float f = __uint_as_float(v[4 * c4 + e]) + bv;
                    if (p.relu_out) f = fmaxf(f, 0.f);
                    o[e] = f;
                }
                if (p.mask_mode) {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(mk) + c4);
                    const float mm[4] = {m.x, m.y, m.z, m.w};
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
                const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(dst0)[c4] = ov;
                if (p.upsample) {                        // nearest x2: the same value at (2y,2x+1), (2y+1,2x), (2y+1,2x+1)
                    reinterpret_cast<float4*>(dst0 + N)[c4] = ov;
                    reinterpret_cast<float4*>(dst0 + (size_t)p.Wq * N)[c4] = ov;
                    reinterpret_cast<float4*>(dst0 + (size_t)p.Wq * N + N)[c4] = ov;
                }
            }
          }
            if (++acc == nbuf) { acc = 0; acc_phase ^= 1u; }
        }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int N, int T>
int launch_tcg(const CUtensorMap& tmA, const CUtensorMap& tmW, const GParams& p, int smem, cudaStream_t st) {
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tcg_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    int grid = p.num_iters < num_sms ? p.num_iters : num_sms;
    return launch_pdl(conv3x3_tcg_kernel<N, T>, dim3(grid), dim3(320), smem, st, tmA, tmW, p);
}

}  // namespace

// x: [B][Hr][Wp][Cin] pitch-linear (zero border where the conv pads); wop: TF32 operand copy [Cout][9][Cin] (forward) or
// [Cin_orig][9 flipped][Cout_orig] (data gradient), made by sgqn_conv_weights_prep_g.  Output (b,y,x), y < Hv, x < Wv, is the
// 3x3 window sum over input rows q + ky*Wp + kx + shift, q = (b*Hr + y)*Wp + x, and goes to
// out[((b*Hq + y+oy)*Wq + x+ox)*Cout]; with flags bit4 it goes to the 2x2 block at (2y+oy, 2x+ox) (nearest upsample).
// flags: bit0 ReLU, bit1 TF32 round, bits 2-3 mask mode against mask[((b*Hm + y)*Wm + x)*Cout], bit4 upsample,
// bits 5-6: 1 = depth-to-space (Cout = 256 = 4 phases x 64: phase 2a+b' of (y,x) -> pixel (2y+a+oy, 2x+b'+ox) of a 64-channel
// buffer), 2 = space-to-depth (pixel (y,x) -> channel block 2(y&1)+(x&1) of low-res pixel (y/2+oy, x/2+ox) in a 4*Cout-channel buffer).
extern "C" int sgqn_conv_tcg(const float* x, const float* wop, const float* bias, const float* mask, float* out, int B, int Hr,
                             int Wp, int Cin, int Cout, int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm, int Wm,
                             int flags, void* stream) {
    return sgqn_conv_tcg_taps(x, wop, bias, mask, out, B, Hr, Wp, Cin, Cout, Hv, Wv, shift, Hq, Wq, oy, ox, Hm, Wm, flags, 9, stream);
}

// ntaps = 9: 3x3 conv (above); ntaps = 1: per-position GEMM out[q][Cout] = x[q][Cin] * wop[Cout][Cin]^T (1x1 conv) -- the
// first encoder conv on its im2col matrix (col[q][96], 81 real columns).
extern "C" int sgqn_conv_tcg_taps(const float* x, const float* wop, const float* bias, const float* mask, float* out, int B, int Hr,
                                  int Wp, int Cin, int Cout, int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm,
                                  int Wm, int flags, int ntaps, void* stream) {
    if (B <= 0) return 0;
    if ((Cin & 31) || (Cout != 32 && Cout != 64 && Cout != 96 && Cout != 128 && Cout != 256)) return (int)cudaErrorInvalidValue;
    GParams p;
    p.total_q = B * Hr * Wp; p.Hr = Hr; p.Wp = Wp; p.Hv = Hv; p.Wv = Wv; p.shift = shift;
    p.Hq = Hq; p.Wq = Wq; p.oy = oy; p.ox = ox; p.Hm = Hm; p.Wm = Wm;
    p.num_tiles = (p.total_q + kTileM - 1) / kTileM;
    p.kc = Cin / 32; p.cin = Cin; p.ntaps = ntaps;
    if (ntaps != 1 && ntaps != 9) return (int)cudaErrorInvalidValue;
    p.bias = bias; p.mask = mask; p.out = out;
    p.relu_out = flags & 1; p.round_out = (flags >> 1) & 1; p.mask_mode = (flags >> 2) & 3; p.upsample = (flags >> 4) & 1;
    p.shuffle = (flags >> 5) & 3;
    if ((p.shuffle == 1) != (Cout == 256) || (Cout == 256 && (p.mask_mode || p.upsample)) || (p.shuffle == 2 && p.upsample) || p.shuffle == 3)
        return (int)cudaErrorInvalidValue;      // N = 256 exists for the depth-to-space phase conv only
    if (p.mask_mode && !mask) return (int)cudaErrorInvalidValue;
    static int num_sms_h = 0;
    if (!num_sms_h) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms_h, cudaDevAttrMultiProcessorCount, dev); }
    const int wbytes_all = (Cin / 32) * ntaps * Cout * 128;
    // Weights too large to stay resident are re-streamed per iteration (L2 -> SM bound at 16 KB per 4 MMAs): two position tiles
    // per weight tile halve that traffic.  N = 128 keeps two accumulator SETS (4 x 128 columns): conv2's data gradient 122 -> 94 us.
    // At N = 256 the pair fills TMEM, the epilogue is exposed and the launch gets slower (68 -> 77 us, measured): not used there.
    p.pair = (ntaps == 9 && Cout == 128 && wbytes_all > 160 * 1024 && p.num_tiles > num_sms_h) ? 1 : 0;
    p.num_iters = (p.num_tiles + p.pair) / (1 + p.pair);
    int halo = ntaps == 9 ? (1 + p.pair) * kTileM + 2 * Wp + 2 : kTileM;
    p.pieces = (halo + 255) / 256;
    p.piece_rows = ((halo + p.pieces - 1) / p.pieces + 7) / 8 * 8;
    if (p.piece_rows > 256) return (int)cudaErrorInvalidValue;
    p.a_bytes = p.pieces * p.piece_rows * 128;
    const int wb = Cout * 128;
    p.w_resident = kSmemBudget - 4096 - p.kc * ntaps * wb >= 2 * p.a_bytes;
    p.w_stages = p.w_resident ? p.kc * ntaps : kMaxW;
    while (p.w_stages > 2 && kSmemBudget - 4096 - p.w_stages * wb < 2 * p.a_bytes) --p.w_stages;
    p.a_stages = (kSmemBudget - 4096 - p.w_stages * wb) / p.a_bytes;
    if (p.a_stages > kMaxA) p.a_stages = kMaxA;
    if (p.a_stages < 2) return (int)cudaErrorInvalidValue;
    int smem = p.a_stages * p.a_bytes + p.w_stages * wb + Cout * 4 + 512 + 1024;
    CUtensorMap tmA, tmW;
    int rc = make_map_2d(&tmA, x, (uint64_t)Cin, (uint64_t)p.total_q, 32, (uint32_t)p.piece_rows);
    if (rc) return rc;
    rc = make_map_2d(&tmW, wop, (uint64_t)ntaps * Cin, (uint64_t)Cout, 32, (uint32_t)Cout);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 32) return launch_tcg<32, 1>(tmA, tmW, p, smem, st);
    if (Cout == 64) return launch_tcg<64, 1>(tmA, tmW, p, smem, st);
    if (Cout == 96) return launch_tcg<96, 1>(tmA, tmW, p, smem, st);
    if (Cout == 256) return launch_tcg<256, 1>(tmA, tmW, p, smem, st);
    return p.pair ? launch_tcg<128, 2>(tmA, tmW, p, smem, st) : launch_tcg<128, 1>(tmA, tmW, p, smem, st);
}

// TF32 operand copies of a [Cout][9][Cin] conv weight (Cout_real <= Cout rows are real, the rest of wf/wd is zero-filled):
// wf[co][t][ci] = rna(w[co][t][ci]);  wd[ci][t][co] = rna(w[co][8-t][ci])
__global__ void conv_weights_prep_g_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int Cout,
                                           int Cin, int Cout_real) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * 9 * Cin) return;
    int co = i / (9 * Cin), r = i - co * 9 * Cin, t = r / Cin, ci = r - t * Cin;
    float v = co < Cout_real ? round_tf32(w[i]) : 0.f;
    wf[i] = v;
    wd[((size_t)ci * 9 + (8 - t)) * Cout + co] = v;
}
extern "C" int sgqn_conv_weights_prep_g(const float* w, float* wf, float* wd, int Cout, int Cin, int Cout_real, void* stream) {
    int n = Cout * 9 * Cin;
    if (n <= 0) return 0;
    conv_weights_prep_g_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, wf, wd, Cout, Cin, Cout_real);
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------------------------------
// Sub-pixel ("phase") form of  conv3x3(pad 1) o nearest-x2-upsample  (modules.py:327-337: F.upsample then the next conv).
// Output pixel (2y+a, 2x+b) of the conv over the upsampled image only ever sees the 3x3 LOW-resolution neighbourhood of
// (y,x): upsampled rows 2y+a-1 .. 2y+a+1 are low-res rows {y-1, y, y} (a = 0) or {y, y, y+1} (a = 1), same for columns.  So
//     Y[2y+a][2x+b][co] = sum_{dy,dx in -1..1} Wphi[(a,b,co)][dy][dx][:] . X[y+dy][x+dx][:]
// with Wphi[(a,b,co)][dy][dx] = sum of the W[co][ky][kx] whose (ky,kx) land on (dy,dx): a plain 3x3 pad-1 conv at LOW
// resolution with 4*Cg output channels (phase p = 2a+b owns channels [p*Cg, p*Cg + Cout_real)), whose output is the
// space-to-depth view of Y.  The upsampled tensor (4x the bytes, read by forward, weight gradient and written by the data
// gradient) is never materialised, the MMA N grows 4x (32 -> 64 for conv3: a 128x32x8 TF32 MMA costs as much as 128x64x8)
// and the data gradient lands directly on the low-res tensor (the 2x2 sum-pool of the upsample backward is the sum over
// phases inside the GEMM).  The zero ring of the upsampled image is the zero ring of the low-res one (row -1 <-> row -1,
// row 2H <-> row H).
__device__ __forceinline__ int phase_tap(int a, int k) { return a == 0 ? (k == 0 ? -1 : 0) : (k == 2 ? 1 : 0); }

// w: [>= Cout_real][9][Cin] fp32 (reference taps); wf: [4*Cg][9][Cin] = rna(Wphi) (forward operand), wd: [Cin][9 flipped][4*Cg]
// (data-gradient operand), bphi: [4*Cg] = bias replicated per phase (0 in the padding channels).
__global__ void conv_weights_prep_phase_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ wf,
                                               float* __restrict__ wd, float* __restrict__ bphi, int Cin, int Cout_real, int Cg) {
    pdl_wait();
    pdl_launch();
    const int Np = 4 * Cg;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np * 9 * Cin) return;
    int row = i / (9 * Cin), r = i - row * 9 * Cin, t = r / Cin, ci = r - t * Cin;
    int ph = row / Cg, co = row - ph * Cg, a = ph >> 1, b = ph & 1, dy = t / 3 - 1, dx = t % 3 - 1;
    float v = 0.f;
    if (co < Cout_real)
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                if (phase_tap(a, ky) == dy && phase_tap(b, kx) == dx) v += w[((size_t)co * 9 + ky * 3 + kx) * Cin + ci];
    v = round_tf32(v);
    wf[i] = v;
    wd[((size_t)ci * 9 + (8 - t)) * Np + row] = v;
    if (i < Np) bphi[i] = (i % Cg) < Cout_real ? bias[i % Cg] : 0.f;
}
extern "C" int sgqn_conv_weights_prep_phase(const float* w, const float* bias, float* wf, float* wd, float* bphi, int Cin,
                                            int Cout_real, int Cg, void* stream) {
    int n = 4 * Cg * 9 * Cin;
    if (n <= 0 || Cout_real > Cg) return (int)cudaErrorInvalidValue;
    { int rc_ = launch_pdl(conv_weights_prep_phase_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, w, bias, wf, wd, bphi, Cin, Cout_real, Cg); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// Chain rule of the map above: dW[co][ky][kx][ci] += sum_{a,b} dWphi[(a,b,co)][tap(a,ky)][tap(b,kx)][ci];
// db[co] += sum_p dbphi[p*Cg + co].  Single writer per element (plain +=).
__global__ void conv_phase_fold_kernel(const float* __restrict__ dwphi, const float* __restrict__ dbphi, float* __restrict__ dw,
                                       float* __restrict__ db, int Cin, int Cout_real, int Cg) {
    pdl_wait();
    pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout_real * 9 * Cin) return;
    int co = i / (9 * Cin), r = i - co * 9 * Cin, t = r / Cin, ci = r - t * Cin, ky = t / 3, kx = t % 3;
    float s = 0.f;
    for (int ph = 0; ph < 4; ++ph) {
        int tp = (phase_tap(ph >> 1, ky) + 1) * 3 + phase_tap(ph & 1, kx) + 1;
        s += dwphi[((size_t)(ph * Cg + co) * 9 + tp) * Cin + ci];
    }
    dw[i] += s;
    if (i < Cout_real && db) db[i] += (dbphi[i] + dbphi[Cg + i]) + (dbphi[2 * Cg + i] + dbphi[3 * Cg + i]);
}
extern "C" int sgqn_conv_phase_fold(const float* dwphi, const float* dbphi, float* dw, float* db, int Cin, int Cout_real, int Cg,
                                    void* stream) {
    int n = Cout_real * 9 * Cin;
    if (n <= 0 || Cout_real > Cg) return (int)cudaErrorInvalidValue;
    { int rc_ = launch_pdl(conv_phase_fold_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, dwphi, dbphi, dw, db, Cin, Cout_real, Cg); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// =====================================================================================================================
// Generalised weight gradient:  dW[co][ky*3+kx][c*32+ci] += sum_q dY[q][co] * X[q + (ky+ta)*Wp + kx + tb][c*32+ci]
// X: [rows][Cin = 32*KC], dY: [rows][N = 32*NA], both pitch-linear over the same position index q.
// grid = (pixel ranges, 3): blockIdx.y = ky.  Per 64-position block a CTA loads, per channel chunk c, ONE X tile of 72 rows
// starting at q0 + (ky+ta)*Wp + tb, and the NA dY tiles.  MN-major TF32 operands (SWIZZLE_128B_BASE32B).  One M=128
// accumulator per chunk: its four 32-row atoms are the SAME tile shifted by kx = 0,1,2,(3 unused) rows -- the UMMA
// descriptor's leading-byte-offset is simply 128 B.  N = Cout columns per accumulator (KC*N <= 512 TMEM columns).
namespace {

constexpr int kGwRows = 64;
constexpr int kGwXRows = 72;
constexpr int kGwStagesMax = 4;

struct GwParams { int total_q, Wp, ta, tb, kc, cin, na, kb_total, kb_per_cta, stages, stage_bytes, ntaps, kvalid, fuse, xrows; float* dw; };

__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
    return d;
}

template <int N>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_tcg_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, GwParams p) {
    constexpr uint32_t kIdescMN = idesc_tf32(N, true, true);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + p.stages * p.stage_bytes + 1024;      // +1 KB: the unused 4th atom reads past the last X tile
    const uint32_t full0 = bars, empty0 = bars + 8 * kGwStagesMax, done_bar = empty0 + 8 * kGwStagesMax, tmem_slot = done_bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = blockIdx.y;
    const int kb0 = blockIdx.x * p.kb_per_cta;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_cta);
    // fuse: all three filter rows in one CTA -- ONE X halo tile (72 + 2 Wp rows) serves ky = 0..2 as row-shifted views and dY is
    // loaded once instead of three times (the ky-split version was L2 -> SM bound: 3 x 34 KB per 64 pixels at Cin = N = 64)
    const int nacc = (p.fuse ? 3 : 1) * p.kc * N;
    const int tmem_cols = nacc < 32 ? 32 : (nacc <= 64 ? 64 : (nacc <= 128 ? 128 : (nacc <= 256 ? 256 : 512)));
    const uint32_t xtile = (uint32_t)p.xrows * 128, dtile = kGwRows * 128;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmD) : "memory");
        for (int s = 0; s < kGwStagesMax; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
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
                const int q0 = kb * kGwRows;
                const uint32_t sb = base + stage * p.stage_bytes;
                mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                mbar_expect_tx(full0 + 8 * stage, p.kc * xtile + p.na * dtile);
                for (int c = 0; c < p.kc; ++c)
                    tma_load_2d(&tmX, full0 + 8 * stage, sb + c * xtile, c * 32, q0 + (ky + p.ta) * p.Wp + p.tb);
                for (int a = 0; a < p.na; ++a)
                    tma_load_2d(&tmD, full0 + 8 * stage, sb + p.kc * xtile + a * dtile, a * 32, q0);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const uint32_t sb = base + stage * p.stage_bytes;
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                // One descriptor pair per accumulator, then the kGwRows / 8 k-steps as straight-line code (+1024 bytes = +64 in the
                // descriptors' address field): the issuing thread paces these short MMAs, so nothing is recomputed per MMA.  Every
                // accumulator still receives its k-steps in the same order as before.
                const uint64_t bd0 = make_desc_mn32(sb + p.kc * xtile, dtile);
                const uint32_t first = kb != kb0 ? 1u : 0u;
                if (p.fuse) {
                    for (int f = 0; f < 3; ++f)
                        for (int c = 0; c < p.kc; ++c) {
                            const uint64_t ad0 = make_desc_mn32(sb + c * xtile + (uint32_t)(f * p.Wp) * 128u, 128);
                            const uint32_t d = tmem_base + (uint32_t)((f * p.kc + c) * N);
#pragma unroll
                            for (int j = 0; j < kGwRows / 8; ++j)
                                tc_mma_tf32(d, ad0 + (uint64_t)(j * 64), bd0 + (uint64_t)(j * 64), kIdescMN, j ? 1u : first);
                        }
                } else if (p.ntaps == 1) {
                    // per-position GEMM: no row shifts, so the four 32-row M atoms are the channel chunks themselves (their
                    // tiles are xtile bytes apart; a 4th atom past kc = 3 reads the dY tile: finite garbage in accumulator
                    // rows nobody reads) -- one MMA per k-step instead of one per chunk with three quarters of M wasted
                    const uint64_t ad0 = make_desc_mn32(sb, xtile);
#pragma unroll
                    for (int j = 0; j < kGwRows / 8; ++j)
                        tc_mma_tf32(tmem_base, ad0 + (uint64_t)(j * 64), bd0 + (uint64_t)(j * 64), kIdescMN, j ? 1u : first);
                } else {
                    for (int c = 0; c < p.kc; ++c) {
                        const uint64_t ad0 = make_desc_mn32(sb + c * xtile, 128);    // atoms = row shifts kx = 0..3
                        const uint32_t d = tmem_base + (uint32_t)(c * N);
#pragma unroll
                        for (int j = 0; j < kGwRows / 8; ++j)
                            tc_mma_tf32(d, ad0 + (uint64_t)(j * 64), bd0 + (uint64_t)(j * 64), kIdescMN, j ? 1u : first);
                    }
                }
                tc_commit(empty0 + 8 * stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            tc_commit(done_bar);
        }
    } else if (kb1 > kb0) {
        const int kx = warp & 3;                          // TMEM lane quarter = accumulator rows of tap kx
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int f = 0; f < (p.fuse ? 3 : 1); ++f) {
        const int kyf = p.fuse ? f : ky;
        for (int c = 0; c < (p.ntaps == 1 ? 1 : p.kc); ++c) {
#pragma unroll 1
            for (int g = 0; g < N / 16; ++g) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(kx * 32) << 16) + (uint32_t)((f * p.kc + c) * N + g * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p.ntaps == 9) {
                    if (kx < 3) {
                        float* dst = p.dw + (size_t)(kyf * 3 + kx) * p.cin + c * 32 + lane;
#pragma unroll
                        for (int e = 0; e < 16; ++e) atomicAdd(dst + (size_t)(g * 16 + e) * 9 * p.cin, __uint_as_float(v[e]));
                    }
                } else if (kx < p.kc && kx * 32 + lane < p.kvalid) {   // per-position GEMM: dw[co][kvalid]; lane quarter = channel chunk
                    float* dst = p.dw + kx * 32 + lane;
#pragma unroll
                    for (int e = 0; e < 16; ++e) atomicAdd(dst + (size_t)(g * 16 + e) * p.kvalid, __uint_as_float(v[e]));
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

template <int N>
int launch_wgrad_tcg(const CUtensorMap& tmX, const CUtensorMap& tmD, GwParams& p, cudaStream_t st) {
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tcg_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    const int gy = (p.ntaps == 9 && !p.fuse) ? 3 : 1;
    int gx = (num_sms + gy - 1) / gy;
    if (gx > p.kb_total) gx = p.kb_total;
    p.kb_per_cta = (p.kb_total + gx - 1) / gx;
    gx = (p.kb_total + p.kb_per_cta - 1) / p.kb_per_cta;
    int smem = p.stages * p.stage_bytes + 1024 + 1024 + 256;
    return launch_pdl(conv3x3_wgrad_tcg_kernel<N>, dim3(gx, gy), dim3(192), smem, st, tmX, tmD, p);
}

}  // namespace

// x: [rows][Cin], dy: [rows][Cout] over the same B*Hr*Wp positions; dw[Cout][9][Cin] += ... (atomic; caller zero-fills).
// (ta, tb): the activation row paired with dy row q for tap (ky,kx) is q + (ky+ta)*Wp + kx + tb.
extern "C" int sgqn_conv_wgrad_tcg(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta,
                                   int tb, void* stream) {
    return sgqn_gemm_wgrad_tcg(x, dy, dw, B, Hr, Wp, Cin, Cout, ta, tb, 9, 0, stream);
}

// ntaps = 1: dw[Cout][kvalid] += dy^T[Cout][rows] * x[rows][:kvalid]  (weight gradient of a per-position GEMM; Wp/ta/tb unused)
static int wgrad_tcg_impl(const float* x, const float* dy, int ldy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta,
                          int tb, int ntaps, int kvalid, void* stream);
extern "C" int sgqn_gemm_wgrad_tcg(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta,
                                   int tb, int ntaps, int kvalid, void* stream) {
    return wgrad_tcg_impl(x, dy, Cout, dw, B, Hr, Wp, Cin, Cout, ta, tb, ntaps, kvalid, stream);
}
// The 3x3 weight gradient against a column block of a wider dy: dy rows are ldy floats apart, the Cout columns starting at
// `dy` are used (dw[Cout][9][Cin] of that block).  The phase form of conv2 has 256 output channels = two blocks of 128.
extern "C" int sgqn_conv_wgrad_tcg_ld(const float* x, const float* dy, int ldy, float* dw, int B, int Hr, int Wp, int Cin, int Cout,
                                      int ta, int tb, void* stream) {
    if (ldy < Cout || (ldy & 3)) return (int)cudaErrorInvalidValue;
    return wgrad_tcg_impl(x, dy, ldy, dw, B, Hr, Wp, Cin, Cout, ta, tb, 9, 0, stream);
}
static int wgrad_tcg_impl(const float* x, const float* dy, int ldy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta,
                          int tb, int ntaps, int kvalid, void* stream) {
    if (B <= 0) return 0;
    if (ntaps != 1 && ntaps != 9) return (int)cudaErrorInvalidValue;
    if ((Cin & 31) || (Cout != 32 && Cout != 64 && Cout != 128) || (Cin / 32) * Cout > 512) return (int)cudaErrorInvalidValue;
    if (ntaps == 1 && Cin > 128) return (int)cudaErrorInvalidValue;      // the channel chunks are the four M atoms of one MMA
    GwParams p;
    p.total_q = B * Hr * Wp; p.Wp = ntaps == 9 ? Wp : 0; p.ta = ntaps == 9 ? ta : 0; p.tb = ntaps == 9 ? tb : 0;
    p.kc = Cin / 32; p.cin = Cin; p.na = Cout / 32; p.dw = dw; p.ntaps = ntaps; p.kvalid = kvalid;
    p.kb_total = (p.total_q + kGwRows - 1) / kGwRows;
    // all three filter rows in one CTA when their accumulators fit TMEM and two stages of the taller X halo tile fit shared memory
    p.xrows = kGwXRows; p.fuse = 0;
    if (ntaps == 9 && 3 * p.kc * Cout <= 512) {
        int xr = (kGwXRows + 2 * Wp + 7) / 8 * 8;
        if (xr <= 256 && 2 * (p.kc * xr * 128 + p.na * kGwRows * 128) <= kSmemBudget - 4096) { p.fuse = 1; p.xrows = xr; }
    }
    p.stage_bytes = p.kc * p.xrows * 128 + p.na * kGwRows * 128;
    p.stages = (kSmemBudget - 4096) / p.stage_bytes;
    if (p.stages > kGwStagesMax) p.stages = kGwStagesMax;
    if (p.stages < 2) return (int)cudaErrorInvalidValue;
    CUtensorMap tmX, tmD;
    int rc = make_map_2d(&tmX, x, (uint64_t)Cin, (uint64_t)p.total_q, 32, (uint32_t)p.xrows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = make_map_2d(&tmD, dy, (uint64_t)ldy, (uint64_t)p.total_q, 32, kGwRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 32) return launch_wgrad_tcg<32>(tmX, tmD, p, st);
    if (Cout == 64) return launch_wgrad_tcg<64>(tmX, tmD, p, st);
    return launch_wgrad_tcg<128>(tmX, tmD, p, st);
}

// backward of "nearest x2 upsample of relu(x)": dst(b,y,x) = (sum of the 2x2 block of dup) * 1[src(2y,2x) > 0]
// dup compact [B][2H][2W][C]; src = the upsampled, zero-bordered forward buffer [B][2H+2][2W+2][C] (value of (Y,X) at row Y+1,
// col X); dst = zero-bordered gradient buffer [B][H+2][W+2][C] (value of (y,x) at row y+1, col x), TF32-rounded.
__global__ void pool2_bwd_kernel(const float4* __restrict__ dup, const float4* __restrict__ src, float4* __restrict__ dst, int H, int W,
                                 int C4, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % C4); long long t = i / C4;
    int x = (int)(t % W); t /= W; int y = (int)(t % H); int b = (int)(t / H);
    const int W2 = 2 * W;
    size_t d0 = (((size_t)b * 2 * H + 2 * y) * W2 + 2 * x) * C4 + c;
    float4 a0 = __ldg(dup + d0), a1 = __ldg(dup + d0 + C4), a2 = __ldg(dup + d0 + (size_t)W2 * C4), a3 = __ldg(dup + d0 + (size_t)W2 * C4 + C4);
    float4 m = __ldg(src + (((size_t)b * (2 * H + 2) + 2 * y + 1) * (W2 + 2) + 2 * x) * C4 + c);
    float4 r;
    r.x = m.x > 0.f ? round_tf32((a0.x + a1.x) + (a2.x + a3.x)) : 0.f;
    r.y = m.y > 0.f ? round_tf32((a0.y + a1.y) + (a2.y + a3.y)) : 0.f;
    r.z = m.z > 0.f ? round_tf32((a0.z + a1.z) + (a2.z + a3.z)) : 0.f;
    r.w = m.w > 0.f ? round_tf32((a0.w + a1.w) + (a2.w + a3.w)) : 0.f;
    dst[(((size_t)b * (H + 2) + y + 1) * (W + 2) + x) * C4 + c] = r;
}
extern "C" int sgqn_pool2_bwd(const float* dup, const float* src, float* dst, int B, int H, int W, int C, void* stream) {
    if (C & 3) return (int)cudaErrorInvalidValue;
    long long total = (long long)B * H * W * (C / 4);
    if (total <= 0) return 0;
    pool2_bwd_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)dup, (const float4*)src,
                                                                                    (float4*)dst, H, W, C / 4, total);
    return SGQN_CHECK_LAUNCH();
}
