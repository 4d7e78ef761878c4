// A whole chain of 32 -> 32 3x3 convolutions (SharedCNN layers 2..11 forward, modules.py:144-146, or their data gradients
// in reverse order) as ONE persistent tcgen05 / TMEM / TMA kernel.
//
// Per-tile arithmetic is that of conv3x3_tc_kernel (conv_tc.cu: pitch-linear implicit GEMM, N = 96 = (kx, cout), the kx shift
// in a staged epilogue) and gives bit-identical results.  What changes is the schedule: one launch per LAYER left every
// layer with its own pipeline fill (TMEM allocation, 36 KB of weights, first halo tile) and drain (last epilogue, a ragged
// last wave of tiles) -- 4-6 us of a 10-20 us launch, ten times per chain.  Here the tiles of all layers form one global
// list (layer-major); CTAs take tickets from a global counter and run the list in order; a tile of layer l waits -- per
// tile, not per layer -- for the 2-4 tiles of layer l-1 that produce the rows its halo tile reads (release / acquire
// flags in global memory), so layer l+1 starts on the first samples while layer l is still finishing the last ones, and
// the activations a layer reads were written microseconds earlier (L2 hits, 126 MB).
//
// Deadlock freedom does not depend on how many CTAs are resident: a ticket's dependencies are all LOWER tickets, a lower
// ticket is owned by a CTA that is already running, and that CTA in turn only waits on lower tickets.  (A static
// round-robin assignment would need every CTA of the grid co-resident -- not guaranteed beside kernels of other streams.)
//
// Warp roles (736 threads, one CTA per SM): warp 0 = scheduler (one lane: tickets -- two requests always in flight --, layer
// lookup, the next layer's weights into the other weight buffer at a layer change, the tile list of this CTA), warps 20..22 =
// fetch warps, one per pipeline stage (poll the producer tiles' flags, one flag per lane; cross-proxy fence; halo TMA load:
// ~2 600 cycles of latency per tile, measured -- as part of ONE producer thread they capped the CTA at a tile per 3 900
// cycles, three in parallel keep ahead of the 1 700-cycle tile), warp 1 = MMA issuer, warps 2..17 =
// epilogue, warp 18 = publisher (waits for the epilogue's stores of a tile, fences, releases the tile's flag: the
// membar.gpu round trip stays out of the epilogue's tile loop), warp 19 = row offsets (output / mask addresses of a tile's
// 128 rows; measured: as part of the scheduler's per-tile work they made the scheduler the bottleneck, 2 500 cycles / tile).
#include "tc_common.cuh"
#include "../../include/sgqn_b200.h"

using namespace tc;

namespace {

constexpr int kMaxLayers = 10;
constexpr int kStages = 3;
constexpr int kTileM = 128, kTileOut = 126, kAccCols = 128;
constexpr int kStgPitch = 400;
constexpr int kStgBytes = kTileM * kStgPitch;          // 51 200
constexpr int kWBytes = 9 * 32 * 128;                  // 36 864 per layer, two buffers
constexpr int kMaxHalo = 224;                          // rows: 128 + 2 * Wp (Wp <= 48), whole 8-row swizzle atoms
constexpr int kStageBytes = kMaxHalo * 128;            // 28 672
constexpr int kInfoDepth = 16, kRowDepth = 16;
constexpr int kSchedLead = 12;                        // tiles the scheduler may publish ahead of the publisher (< ring depths)
constexpr int kEpiWarps = 16, kEpiThreads = kEpiWarps * 32;
constexpr int kFetchWarp0 = 2 + kEpiWarps + 2;               // warps 20..22: one fetch warp per stage
constexpr int kThreads = 64 + kEpiThreads + 64 + 32 * kStages;
constexpr int kDoneDepth = 4;                          // publisher ring: the epilogue may lead it by at most 3 tiles
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

struct LayerP {
    int total_q, Hr, Wp, Hv, Wv, shift;
    int Hq, Wq, oy, ox, Hm, Wm;
    int num_tiles, tile0, halo_rows;
    int relu_out, round_out, mask_mode;
    int magic_wp;               // (x * magic_wp) >> 16 == x / Wp for x < 256
    float r_hw, r_wp;           // 1 / (Hr * Wp), 1 / Wp: quotients by one multiply + a +-1 correction (operands < 2^24)
    const float* bias;
    const float* mask;
    float* out;
    float* dbias;
};
struct ChainParams {
    int n_layers, total_tiles, lead;
    int* sync;                  // [0] ticket counter, [1] CTAs that have finished, [2] launch epoch
    int* flags;                 // [total_tiles]: == epoch + 1 once the tile's outputs are visible
    LayerP L[kMaxLayers];
};
struct alignas(64) ChainMaps {
    CUtensorMap a[kMaxLayers];
    CUtensorMap w[kMaxLayers];
};

struct TileInfo { int layer, tile, gtile, wsel_new; };   // wsel_new: bit0 = weight buffer, bit1 = first tile of this layer in this CTA

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// floor(n / d) for 0 <= n < 2^24 with rd = 1 / d: one multiply and a correction instead of the ~60-instruction integer division
// (the scheduler and the row-offset warp are serial per tile)
__device__ __forceinline__ int div_small(int n, int d, float rd) {
    int q = (int)((float)n * rd);
    const int r = n - q * d;
    if (r < 0) --q; else if (r >= d) ++q;
    return q;
}

// position (in the PRODUCER layer's coordinates) of the output that lands on row r of the consumer's input buffer, rounded
// to the nearest written row in the same sample (monotone in r: rows nobody writes stay zero and need no producer)
__device__ __forceinline__ int producer_pos(const LayerP& P, int r, int HWc, int Wpc, float rHWc, float rWpc) {
    const int b = div_small(r, HWc, rHWc); const int rem = r - b * HWc; const int y = div_small(rem, Wpc, rWpc); const int x = rem - y * Wpc;
    int yy = y - P.oy, xx = x - P.ox;
    if (yy < 0) { yy = 0; xx = 0; }
    else if (yy >= P.Hv) { yy = P.Hv - 1; xx = P.Wv - 1; }
    else xx = min(max(xx, 0), P.Wv - 1);
    return (b * P.Hr + yy) * P.Wp + xx;
}

__device__ __forceinline__ uint8_t* sm0_of(uint8_t* raw, uint32_t base) { return raw + (base - smem_u32(raw)); }

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
    constexpr int kChPerThread = 128 / kEpiWarps;      // 8
    constexpr int kItems = 1024 / kEpiThreads;         // 2
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_sm = base;
    const uint32_t w_sm = a_sm + kStages * kStageBytes;
    const uint32_t stg_sm = w_sm + 2 * kWBytes;
    const uint32_t row_sm = stg_sm + kStgBytes;                         // int2 [kRowDepth][128]
    const uint32_t info_sm = row_sm + kRowDepth * 128 * 8;              // TileInfo [kInfoDepth]
    const uint32_t lay_sm = info_sm + kInfoDepth * 16;                  // LayerP [kMaxLayers]
    const uint32_t bars = (lay_sm + kMaxLayers * (uint32_t)sizeof(LayerP) + 15u) & ~15u;
    const uint32_t full0 = bars, empty0 = full0 + 8 * kStages, wfull0 = empty0 + 8 * kStages, wempty0 = wfull0 + 16;
    const uint32_t tfull0 = wempty0 + 16, tempty0 = tfull0 + 16, tdone0 = tempty0 + 16, ifull0 = tdone0 + 8 * kDoneDepth;
    const uint32_t rfull0 = ifull0 + 8 * kInfoDepth, tmem_slot = rfull0 + 8 * kRowDepth;
    uint8_t* const sm0 = sm0_of(smem_raw, base);                        // generic pointer to `base`
    volatile int* const pub_seq = reinterpret_cast<volatile int*>(sm0 + (tmem_slot + 8 - base));   // tiles published so far
    uint8_t* const stg = sm0 + (stg_sm - base);
    int2* const rowring = reinterpret_cast<int2*>(sm0 + (row_sm - base));
    TileInfo* const info = reinterpret_cast<TileInfo*>(sm0 + (info_sm - base));
    LayerP* const Ls = reinterpret_cast<LayerP*>(sm0 + (lay_sm - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < p.n_layers * (int)(sizeof(LayerP) / 4); i += blockDim.x)
        reinterpret_cast<int*>(Ls)[i] = reinterpret_cast<const int*>(p.L)[i];
    if (warp == 0 && lane == 0) {
        for (int l = 0; l < p.n_layers; ++l) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&maps.a[l]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&maps.w[l]) : "memory");
        }
        for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(wfull0 + 8 * b, 1); mbar_init(wempty0 + 8 * b, 1);
            mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, kEpiWarps);
        }
        for (int b = 0; b < kDoneDepth; ++b) mbar_init(tdone0 + 8 * b, kEpiWarps);
        for (int b = 0; b < kInfoDepth; ++b) mbar_init(ifull0 + 8 * b, 1);
        for (int b = 0; b < kRowDepth; ++b) mbar_init(rfull0 + 8 * b, 1);
        *pub_seq = 0;
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
        // ------------------------------------------------------------------ scheduler: tickets -> tile list of this CTA
        // One lane; per tile it only takes a ticket (two requests are always in flight in two registers, so the atomic's round
        // trip is hidden), finds the layer, loads the next layer's weights at a layer change, waits for the tile's stage and
        // publishes the tile (info ring).  Everything with a long latency -- polling the producer tiles' flags, the cross-proxy
        // fence, the halo load -- happens in the fetch warp that owns the stage, three tiles in parallel.
        if (lane == 0) {
            int seq = 0, cur = -1, wsel = 1, wl0 = 0, wl1 = 0;
            int gA = atomicAdd(p.sync, 1), gB = atomicAdd(p.sync, 1);
            bool done = false;
            auto take = [&](int& gx) {
                const int g = gx;
                const bool last = g >= p.total_tiles;
                if (!last) gx = atomicAdd(p.sync, 1);            // consumed two tiles from now
                int l = -1, t = 0, flagsw = 0;
                if (!last) {
                    l = cur < 0 ? 0 : cur;
                    while (g >= Ls[l].tile0 + Ls[l].num_tiles) ++l;
                    t = g - Ls[l].tile0;
                    if (l != cur) {
                        // next layer's weights into the other buffer, once the MMAs of the layer before last have retired
                        wsel ^= 1; flagsw = 2;
                        const int k = wsel ? wl1++ : wl0++;
                        mbar_wait(wempty0 + 8 * wsel, (uint32_t)((k & 1) ^ 1));
                        mbar_expect_tx(wfull0 + 8 * wsel, kWBytes);
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_2d(&maps.w[l], wfull0 + 8 * wsel, w_sm + wsel * kWBytes + tap * 4096, tap * 32, 0);
                    }
                    flagsw |= wsel;
                    cur = l;
                }
                {   // stay within the rings: at most kSchedLead tiles ahead of what the publisher has retired
                    uint32_t it = 0;
                    while (seq - *pub_seq >= p.lead) {
                        __nanosleep(64);
                        if (++it > (1u << 24)) { printf("sgqn conv_chain: scheduler stalled (block %d)\n", blockIdx.x); __trap(); }
                    }
                }
                TileInfo ti; ti.layer = l; ti.tile = t; ti.gtile = g; ti.wsel_new = flagsw;
                info[seq & (kInfoDepth - 1)] = ti;
                mbar_arrive(ifull0 + 8 * (seq & (kInfoDepth - 1)));         // -> fetch warp of this stage, row-offset warp
                ++seq;
                if (last) {
                    // end of the list: the other two fetch warps get a "stop" entry of their own (nothing downstream reads it)
                    for (int k = 0; k < kStages - 1; ++k) {
                        TileInfo te; te.layer = -2; te.tile = 0; te.gtile = 0; te.wsel_new = 0;
                        info[seq & (kInfoDepth - 1)] = te;
                        mbar_arrive(ifull0 + 8 * (seq & (kInfoDepth - 1)));
                        ++seq;
                    }
                    done = true;
                }
            };
            while (!done) {
                take(gA);
                if (!done) take(gB);
            }
        }
    } else if (warp >= kFetchWarp0) {
        // ------------------------------------------------------------------ fetch warps: one per stage (tiles seq = stage mod 3)
        const int stage = warp - kFetchWarp0;
        const int want = *reinterpret_cast<volatile int*>(p.sync + 2) + 1;
        uint32_t phase = 0;
        for (int seq = stage;; seq += kStages, phase ^= 1u) {
            mbar_wait(ifull0 + 8 * (seq & (kInfoDepth - 1)), (uint32_t)((seq / kInfoDepth) & 1));
            const TileInfo ti = info[seq & (kInfoDepth - 1)];
            const int l = ti.layer, t = ti.tile;
            if (l < 0) {
                if (l == -1 && lane == 0) {              // pass the end marker on to the MMA warp
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    mbar_arrive(full0 + 8 * stage);
                }
                break;
            }
            const LayerP& P = Ls[l];
            if (l > 0) {
                // the producer tiles (layer l-1) of the rows [r0, r1) this tile's halo load reads: one flag per lane, polled together
                const LayerP& Q = Ls[l - 1];
                const int r0 = max(t * kTileOut + P.shift, 0);
                const int r1 = min(t * kTileOut + P.shift + P.halo_rows, P.total_q);
                if (r1 > r0) {
                    const int HWc = P.Hr * P.Wp;
                    const int t_lo = div_small(producer_pos(Q, r0, HWc, P.Wp, P.r_hw, P.r_wp), kTileOut, 1.0f / kTileOut);
                    const int t_hi = min(div_small(producer_pos(Q, r1 - 1, HWc, P.Wp, P.r_hw, P.r_wp), kTileOut, 1.0f / kTileOut), Q.num_tiles - 1);
                    for (int j0 = t_lo; j0 <= t_hi; j0 += 32) {
                        const int j = j0 + lane;
                        const int* f = p.flags + Q.tile0 + j;
                        uint32_t it = 0;
                        for (;;) {
                            const int v = j <= t_hi ? ld_acquire(f) : want;
                            if (__all_sync(0xffffffffu, v == want)) break;
                            __nanosleep(32);
                            if (++it > (1u << 24)) {
                                if (v != want) printf("sgqn conv_chain: flag timeout (block %d layer %d tile %d waits for %d)\n", blockIdx.x, l, t, j);
                                __trap();
                            }
                        }
                    }
                    __syncwarp();                        // lane 0 issues the TMA: after every lane's acquire (warp barrier = memory order)
                    if (lane == 0) fence_proxy_async_all();     // ... and the async-proxy read after these generic-proxy observations
                }
            }
            if (lane == 0) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1u);       // the stage's previous tile has been consumed by the MMAs
                mbar_expect_tx(full0 + 8 * stage, P.halo_rows * 128);
                tma_load_2d(&maps.a[l], full0 + 8 * stage, a_sm + stage * kStageBytes, 0, t * kTileOut + P.shift);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0; int seq = 0, wf0 = 0, wf1 = 0, wcur = -1;
            for (;;) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const TileInfo ti = info[seq & (kInfoDepth - 1)];
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                if (ti.layer < 0) break;                 // end of the list
                const int wsel = ti.wsel_new & 1;
                if (ti.wsel_new & 2) {
                    if (wcur >= 0) tc_commit(wempty0 + 8 * wcur);       // every MMA on the previous layer's weights has been issued
                    const int k = wsel ? wf1++ : wf0++;
                    mbar_wait(wfull0 + 8 * wsel, (uint32_t)(k & 1));
                    tc_fence_after();
                    wcur = wsel;
                }
                const int Wp = Ls[ti.layer].Wp;
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                for (int ky = 0; ky < 3; ++ky) {
                    const uint64_t ad = make_desc_sw128(a_sm + stage * kStageBytes + ky * Wp * 128);
                    const uint64_t bd = make_desc_sw128(w_sm + wsel * kWBytes + ky * 3 * 4096);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_tf32(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc, (ky | k) != 0);
                }
                tc_commit(empty0 + 8 * stage);
                tc_commit(tfull0 + 8 * acc);
                ++seq;
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp == 2 + kEpiWarps) {
        // ------------------------------------------------------------------ publisher: tile done -> flag visible to the other SMs
        if (lane == 0) {
            const int want = *reinterpret_cast<volatile int*>(p.sync + 2) + 1;
            for (int seq = 0;; ++seq) {
                mbar_wait(tdone0 + 8 * (seq & (kDoneDepth - 1)), (uint32_t)((seq / kDoneDepth) & 1));
                const TileInfo ti = info[seq & (kInfoDepth - 1)];
                if (ti.layer < 0) break;
                st_release(p.flags + ti.gtile, want);    // release.gpu: cumulative over the epilogue warps' stores (their mbarrier arrivals)
                *pub_seq = seq + 1;
            }
        }
    } else if (warp == 3 + kEpiWarps) {
        // ------------------------------------------------------------------ row offsets of every tile, for the epilogue
        for (int seq = 0;; ++seq) {
            mbar_wait(ifull0 + 8 * (seq & (kInfoDepth - 1)), (uint32_t)((seq / kInfoDepth) & 1));
            const TileInfo ti = info[seq & (kInfoDepth - 1)];
            if (ti.layer < 0) break;
            const LayerP& P = Ls[ti.layer];
            // output / mask offsets (in float4 units) of the tile's 128 rows: (b, y, x) of the tile's first position by division
            // (warp-uniform), every row from it by carries (x' < 256: (x' * magic) >> 16 == x' / Wp)
            const int HW = P.Hr * P.Wp;
            const int q0 = ti.tile * kTileOut;
            const int b0 = div_small(q0, HW, P.r_hw); const int r20 = q0 - b0 * HW; const int y0 = div_small(r20, P.Wp, P.r_wp);
            const int x0 = r20 - y0 * P.Wp;
            const int magic = P.magic_wp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = lane + 32 * j;
                int2 ri = make_int2(-1, 0);
                if (row < kTileOut && q0 + row < P.total_q) {
                    const int xs = x0 + row;
                    const int dy = (xs * magic) >> 16;
                    const int x = xs - dy * P.Wp;
                    int y = y0 + dy, b = b0;
                    if (y >= P.Hr) { y -= P.Hr; ++b; }
                    if (y < P.Hv && x < P.Wv)
                        ri = make_int2(((b * P.Hq + y + P.oy) * P.Wq + x + P.ox) * 8, ((b * P.Hm + y) * P.Wm + x) * 8);
                }
                rowring[(seq & (kRowDepth - 1)) * 128 + row] = ri;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rfull0 + 8 * (seq & (kRowDepth - 1)));
        }
    } else {
        // ------------------------------------------------------------------ epilogue (conv_tc.cu's, driven by the tile list)
        const int et = threadIdx.x - 64;
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int chunk = et & 7, rslot = et >> 3;
        int acc = 0; uint32_t acc_phase = 0; int seq = 0, cur = -1;
        float csum[4] = {0.f, 0.f, 0.f, 0.f};
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto flush_dbias = [&](int l) {
            // per-channel sum of what this CTA wrote for layer l = its share of the bias gradient of the layer below
            float* dst = Ls[l].dbias;
            if (dst) {
                float* red = reinterpret_cast<float*>(stg);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float tsum = csum[e];
                    tsum += __shfl_xor_sync(0xffffffffu, tsum, 8);
                    tsum += __shfl_xor_sync(0xffffffffu, tsum, 16);
                    if (lane < 8) red[(warp - 2) * 32 + chunk * 4 + e] = tsum;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (et < 32) {
                    float tsum = 0.f;
#pragma unroll
                    for (int w = 0; w < kEpiWarps; ++w) tsum += red[w * 32 + et];
                    atomicAdd(dst + et, tsum);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) csum[e] = 0.f;
        };
        for (;;) {
            // the tile's identity, row offsets and mask values are known long before its accumulator is complete
            mbar_wait(ifull0 + 8 * (seq & (kInfoDepth - 1)), (uint32_t)((seq / kInfoDepth) & 1));
            const TileInfo ti = info[seq & (kInfoDepth - 1)];
            if (ti.layer < 0) {
                if (lane == 0) mbar_arrive(tdone0 + 8 * (seq & (kDoneDepth - 1)));
                break;
            }
            if (et == 0) {                               // the publisher ring slot of this tile must be free again (it rarely is not)
                uint32_t it = 0;
                while (*pub_seq < seq - (kDoneDepth - 1)) {
                    __nanosleep(32);
                    if (++it > (1u << 24)) { printf("sgqn conv_chain: publisher stalled (block %d)\n", blockIdx.x); __trap(); }
                }
            }
            if (ti.layer != cur) {
                if (cur >= 0) flush_dbias(cur);
                cur = ti.layer;
                const float* bp = Ls[cur].bias;
                bias4 = bp ? __ldg(reinterpret_cast<const float4*>(bp) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const LayerP& P = Ls[cur];
            const int mask_mode = P.mask_mode, relu_out = P.relu_out, round_out = P.round_out;
            const float4* mask4 = reinterpret_cast<const float4*>(P.mask);
            float4* out4 = reinterpret_cast<float4*>(P.out);
            mbar_wait(rfull0 + 8 * (seq & (kRowDepth - 1)), (uint32_t)((seq / kRowDepth) & 1));    // (long complete)
            const int2* rowinfo = rowring + (seq & (kRowDepth - 1)) * 128;
            int2 inf[kItems];
            float4 mk[kItems];
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const int r = rslot + (kEpiThreads / 8) * j;
                inf[j] = r < kTileOut ? rowinfo[r] : make_int2(-1, 0);
                if (mask_mode && inf[j].x >= 0) mk[j] = __ldg(mask4 + inf[j].y + chunk);
            }
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            // ---- phase 1: TMEM -> staging
            {
                uint32_t v[3 * kChPerThread];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols + half * kChPerThread);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(v[kx * 8 + 0]), "=r"(v[kx * 8 + 1]), "=r"(v[kx * 8 + 2]), "=r"(v[kx * 8 + 3]), "=r"(v[kx * 8 + 4]),
                                   "=r"(v[kx * 8 + 5]), "=r"(v[kx * 8 + 6]), "=r"(v[kx * 8 + 7])
                                 : "r"(taddr + (uint32_t)(32 * kx)) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);         // one arrival per warp: the accumulator is drained
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                uint4* srow = reinterpret_cast<uint4*>(stg + row * kStgPitch + half * (kChPerThread * 4));
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int c = 0; c < kChPerThread / 4; ++c)
                        srow[kx * 8 + c] = make_uint4(v[kx * 8 + 4 * c], v[kx * 8 + 4 * c + 1], v[kx * 8 + 4 * c + 2], v[kx * 8 + 4 * c + 3]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            // ---- phase 2: shifted sum, epilogue math, coalesced store
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                if (inf[j].x < 0) continue;
                const int r = rslot + (kEpiThreads / 8) * j;
                const float4 a0 = *reinterpret_cast<const float4*>(stg + r * kStgPitch + chunk * 16);
                const float4 a1 = *reinterpret_cast<const float4*>(stg + (r + 1) * kStgPitch + 128 + chunk * 16);
                const float4 a2 = *reinterpret_cast<const float4*>(stg + (r + 2) * kStgPitch + 256 + chunk * 16);
                float o[4] = {a0.x + a1.x + a2.x + bias4.x, a0.y + a1.y + a2.y + bias4.y, a0.z + a1.z + a2.z + bias4.z,
                              a0.w + a1.w + a2.w + bias4.w};
                if (relu_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                if (mask_mode) {
                    const float mm[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (mask_mode == 2) o[e] = fmaxf(o[e], 0.f);
                        o[e] = mm[e] > 0.f ? o[e] : 0.f;
                    }
                }
                if (round_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = round_tf32(o[e]);
                }
                out4[inf[j].x + chunk] = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) csum[e] += o[e];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tdone0 + 8 * (seq & (kDoneDepth - 1)));   // this warp's stores of the tile are issued -> publisher
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // staging is free
            ++seq;
        }
        if (cur >= 0) flush_dbias(cur);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
    if (threadIdx.x == 0) {
        // the last CTA out re-arms the ticket counter and advances the epoch for the next launch on this workspace
        __threadfence();
        if (atomicAdd(p.sync + 1, 1) == (int)gridDim.x - 1) {
            p.sync[0] = 0; p.sync[1] = 0;
            __threadfence();
            atomicAdd(p.sync + 2, 1);
        }
    }
}

constexpr int kChainSmem = 1024 + kStages * kStageBytes + 2 * kWBytes + kStgBytes + kRowDepth * 128 * 8 + kInfoDepth * 16 +
                           kMaxLayers * (int)sizeof(LayerP) + 640;

}  // namespace

// layers[i]: one call of sgqn_conv_tc (same fields, same meaning); layer i+1 must read what layer i writes (x == out of
// the layer before, with that layer's output geometry as its input geometry).  ws: int workspace of >= 4 + (total tiles)
// entries, zero-initialised ONCE by the caller and then owned by the launches of one stream (two chains that may run
// concurrently need two workspaces).
extern "C" int sgqn_conv_chain(const sgqn_conv_layer* layers, int n_layers, int* ws, long long ws_ints, void* stream) {
    if (n_layers <= 0) return 0;
    if (n_layers > kMaxLayers || !ws) return (int)cudaErrorInvalidValue;
    static int inited = 0, num_sms = 0;
    if (!inited) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        inited = 1;
    }
    ChainParams p;
    ChainMaps maps;
    p.n_layers = n_layers;
    { const char* e = getenv("SGQN_CHAIN_LEAD"); p.lead = e ? atoi(e) : kSchedLead; if (p.lead < 1 || p.lead > kSchedLead) p.lead = kSchedLead; }
    int tiles = 0;
    for (int i = 0; i < n_layers; ++i) {
        const sgqn_conv_layer& s = layers[i];
        LayerP& L = p.L[i];
        if (s.B <= 0) return (int)cudaErrorInvalidValue;
        L.total_q = s.B * s.Hr * s.Wp; L.Hr = s.Hr; L.Wp = s.Wp; L.Hv = s.Hv; L.Wv = s.Wv; L.shift = s.shift;
        L.Hq = s.Hq; L.Wq = s.Wq; L.oy = s.oy; L.ox = s.ox; L.Hm = s.Hm; L.Wm = s.Wm;
        L.num_tiles = (L.total_q + kTileOut - 1) / kTileOut; L.tile0 = tiles;
        tiles += L.num_tiles;
        L.halo_rows = (kTileM + 2 * s.Wp + 7) / 8 * 8;
        if (L.halo_rows > kMaxHalo) return (int)cudaErrorInvalidValue;
        L.magic_wp = (65536 + s.Wp - 1) / s.Wp; L.r_hw = 1.0f / (float)(s.Hr * s.Wp); L.r_wp = 1.0f / (float)s.Wp;
        if (L.total_q >= (1 << 24)) return (int)cudaErrorInvalidValue;
        L.relu_out = s.flags & 1; L.round_out = (s.flags >> 1) & 1; L.mask_mode = (s.flags >> 2) & 3;
        L.bias = s.bias; L.mask = s.mask; L.out = s.out; L.dbias = s.dbias;
        if (L.mask_mode && !s.mask) return (int)cudaErrorInvalidValue;
        if (i > 0) {
            const sgqn_conv_layer& q = layers[i - 1];
            if (s.x != q.out || s.B != q.B || s.Hr != q.Hq || s.Wp != q.Wq) return (int)cudaErrorInvalidValue;
        }
        int rc = make_map_2d(&maps.a[i], s.x, 32, (uint64_t)L.total_q, 32, (uint32_t)L.halo_rows);
        if (rc) return rc;
        rc = make_map_2d(&maps.w[i], s.w, 288, 32, 32, 32);
        if (rc) return rc;
    }
    for (int i = n_layers; i < kMaxLayers; ++i) { p.L[i] = p.L[0]; maps.a[i] = maps.a[0]; maps.w[i] = maps.w[0]; }
    p.total_tiles = tiles;
    if (ws_ints < 4 + (long long)tiles) return (int)cudaErrorInvalidValue;
    p.sync = ws; p.flags = ws + 4;
    const int grid = tiles < num_sms ? tiles : num_sms;
    return launch_pdl(conv3x3_chain_kernel, dim3(grid), dim3(kThreads), kChainSmem, stream, maps, p);
}
