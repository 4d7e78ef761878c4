// Gradient exchange of the batch-sharded update over NVLink peer memory, without NCCL kernels.
//
// Every rank owns one SYMMETRIC arena (same layout on every GPU, mapped into every peer's address space; the host side
// passes the `world` base pointers):   [ flags | control | small slots | ... gradient arena ... ].
// Why not NCCL: the update's kernels are persistent, one CTA per SM with ~all of its shared memory.  An NCCL CTA cannot
// share an SM with them, so every collective that overlaps the backward pass takes SMs away and the persistent kernels
// finish in two waves.  These kernels use no shared memory, 128 threads and <= 64 registers per thread: their CTAs fit
// BESIDE a resident conv / GEMM CTA (one per SM: 148 x 128 threads keep ~1 MB of peer loads in flight), and the tiny
// exchanges (min/max, loss vector, alpha gradient) are one 64-thread CTA.
//
// sgqn_p2p_allreduce_sum: in place, two-shot.  Rank r owns slice r of the range: after a cross-GPU barrier (every peer's
// gradients are complete) it reads slice r from every rank in rank order (so every replica ends up with bit-identical
// sums), adds, and writes the result into slice r of EVERY rank's arena; a second barrier publishes it.
// sgqn_p2p_small: one-shot.  Every rank stores its n <= 32 values into slot [rank] of every peer (double-buffered by call
// parity), one barrier, then reduces the `world` slots locally (sum / max, fp32 or fp64).
// Barriers are per CTA index: CTA b of rank r stores a monotonically increasing ticket into flag [slot][b][r] of every
// peer (st.release.sys after a system fence) and spins on its own flags (ld.acquire.sys).  The ticket is derived from a
// per-slot call counter kept in the rank's own control block, so CUDA-graph replays need no host-side state.  A slot is
// used from ONE stream per rank and every rank issues a slot's collectives in the same order (dist.py).
#include "common.cuh"
#include "../../include/sgqn_b200.h"

namespace {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 160;
constexpr int kSlots = 8;
constexpr int kUnroll = 4;                         // 16-byte loads in flight per thread and peer

struct Peers { char* base[kMaxWorld]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// CTA b of every rank meets CTA b of every other rank at ticket v.
__device__ __forceinline__ void xbarrier(const Peers& P, int rank, int world, long long flags_off, int slot, uint32_t v, uint32_t* err) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        const size_t row = ((size_t)slot * kMaxCtas + blockIdx.x) * kMaxWorld;
        st_release_sys(reinterpret_cast<uint32_t*>(P.base[threadIdx.x] + flags_off) + row + rank, v);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(P.base[rank] + flags_off) + row + threadIdx.x;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - v) < 0) {
            if (++spins > (1u << 26)) {                 // a peer never arrived (~1 min): record it and carry on instead of hanging the GPU
                atomicAdd(err, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128, 8)            // <= 64 registers x 128 threads: a CTA fits beside a resident 576-thread conv CTA
p2p_allreduce_kernel(Peers P, int rank, int world, long long flags_off, long long ctl_off, int slot, long long data_off, long long n4) {
    pdl_wait();                                        // (no early launch_dependents: the next kernel reads what the last barrier publishes)
    uint32_t* ctl = reinterpret_cast<uint32_t*>(P.base[rank] + ctl_off) + slot * 4;     // {calls so far, CTAs done, barrier time-outs}
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
    xbarrier(P, rank, world, flags_off, slot, 2u * e - 1u, ctl + 2);
    const long long per = (n4 + world - 1) / world;
    const long long s0 = rank * per, s1 = s0 + per < n4 ? s0 + per : n4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += kUnroll * stride) {
        float4 acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {
            const float4* src = reinterpret_cast<const float4*>(P.base[p] + data_off);
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (i + u * stride < s1) v[u] = ld_volatile4(src + i + u * stride);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (i + u * stride < s1) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
        for (int p = 0; p < world; ++p) {
            float4* dst = reinterpret_cast<float4*>(P.base[p] + data_off);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (i + u * stride < s1) dst[i + u * stride] = acc[u];
        }
    }
    xbarrier(P, rank, world, flags_off, slot, 2u * e, ctl + 2);
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(ctl + 1, 1u);
        if (prev == gridDim.x - 1) {                    // every CTA of this call has read the counter: advance it
            ctl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(ctl) = e;
        }
    }
}

// Small ranges (the 0.38 MB SharedCNN gradients that the optimiser step waits for): ONE barrier instead of two.  Every rank
// pushes its whole range into slot [parity][rank] of every peer's staging area, one barrier, then every rank adds the `world`
// copies in rank order (bit-identical replicas again).  Staging is double-buffered by call parity: a rank can only be one call
// ahead of a peer (its next barrier needs that peer's ticket), so the copy a slow peer is still reading is never overwritten.
__global__ void __launch_bounds__(128, 8)
p2p_allreduce_push_kernel(Peers P, int rank, int world, long long flags_off, long long ctl_off, int slot, long long data_off, long long n4,
                          long long stage_off, long long stage_stride) {
    pdl_wait();
    uint32_t* ctl = reinterpret_cast<uint32_t*>(P.base[rank] + ctl_off) + slot * 4;
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
    const long long par_off = stage_off + (long long)(e & 1u) * kMaxWorld * stage_stride;
    float4* mine = reinterpret_cast<float4*>(P.base[rank] + data_off);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = mine[i];
        for (int p = 0; p < world; ++p)
            if (p != rank) reinterpret_cast<float4*>(P.base[p] + par_off + rank * stage_stride)[i] = v;
    }
    xbarrier(P, rank, world, flags_off, slot, e, ctl + 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {
            const float4 v = p == rank ? mine[i] : ld_volatile4(reinterpret_cast<const float4*>(P.base[rank] + par_off + p * stage_stride) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        mine[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(ctl + 1, 1u);
        if (prev == gridDim.x - 1) {
            ctl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(ctl) = e;
        }
    }
}

// op 0: fp32 sum, 1: fp32 max, 2: fp64 sum (n counts elements; a slot holds 128 bytes per rank)
__global__ void __launch_bounds__(64)
p2p_small_kernel(Peers P, int rank, int world, long long flags_off, long long ctl_off, long long small_off, int slot, const void* src,
                 void* dst, int n, int op) {
    pdl_wait();
    uint32_t* ctl = reinterpret_cast<uint32_t*>(P.base[rank] + ctl_off) + slot * 4;
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
    const size_t area = (size_t)small_off + ((size_t)(slot * 2 + (e & 1u)) * kMaxWorld) * 128;
    const int t = threadIdx.x;
    if (t < n) {
        for (int p = 0; p < world; ++p) {
            char* slot_p = P.base[p] + area + (size_t)rank * 128;
            if (op == 2) reinterpret_cast<volatile double*>(slot_p)[t] = reinterpret_cast<const double*>(src)[t];
            else reinterpret_cast<volatile float*>(slot_p)[t] = reinterpret_cast<const float*>(src)[t];
        }
    }
    xbarrier(P, rank, world, flags_off, slot, e, ctl + 2);
    if (t < n) {
        const char* mine = P.base[rank] + area;
        if (op == 2) {
            double acc = 0.0;
            for (int p = 0; p < world; ++p) acc += reinterpret_cast<const volatile double*>(mine + (size_t)p * 128)[t];
            reinterpret_cast<double*>(dst)[t] = acc;
        } else {
            float acc = reinterpret_cast<const volatile float*>(mine)[t];
            for (int p = 1; p < world; ++p) {
                const float v = reinterpret_cast<const volatile float*>(mine + (size_t)p * 128)[t];
                acc = op == 1 ? fmaxf(acc, v) : acc + v;
            }
            reinterpret_cast<float*>(dst)[t] = acc;
        }
    }
    __syncthreads();
    if (t == 0) *reinterpret_cast<volatile uint32_t*>(ctl) = e;
}

int fill_peers(Peers* P, const void* const* bases, int rank, int world) {
    if (!bases || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return (int)cudaErrorInvalidValue;
    for (int i = 0; i < kMaxWorld; ++i) P->base[i] = i < world ? (char*)bases[i] : nullptr;
    return 0;
}

}  // namespace

// Arena header the host side must reserve: out3 (HOST pointer) = {bytes of the flag block, of the control block, of the small slots}.
extern "C" int sgqn_p2p_layout(long long* out3) {
    if (!out3) return (int)cudaErrorInvalidValue;
    out3[0] = (long long)kSlots * kMaxCtas * kMaxWorld * 4;
    out3[1] = (long long)kSlots * 16;
    out3[2] = (long long)kSlots * 2 * kMaxWorld * 128;
    return 0;
}

// In-place sum over the ranks of arena[data_off : data_off + 4 n] (bytes; data_off and n*4 multiples of 16).  stage_off >= 0
// names a staging area of 2 * 8 * stage_stride bytes inside the arena: ranges of at most stage_stride bytes then take the
// one-barrier push form, larger ones the two-shot form.
extern "C" int sgqn_p2p_allreduce_sum(const void* const* bases, int rank, int world, long long flags_off, long long ctl_off, int slot,
                                      long long data_off, long long n, int ctas, long long stage_off, long long stage_stride,
                                      void* stream) {
    if (n <= 0) return 0;
    Peers P;
    int rc = fill_peers(&P, bases, rank, world);
    if (rc) return rc;
    if ((n & 3) || (data_off & 15) || slot < 0 || slot >= kSlots || ctas < 1 || ctas > kMaxCtas) return (int)cudaErrorInvalidValue;
    if (stage_off >= 0 && 4 * n <= stage_stride) {
        if ((stage_off & 15) || (stage_stride & 15)) return (int)cudaErrorInvalidValue;
        return launch_pdl(p2p_allreduce_push_kernel, dim3(ctas), dim3(128), 0, stream, P, rank, world, flags_off, ctl_off, slot, data_off,
                          n / 4, stage_off, stage_stride);
    }
    return launch_pdl(p2p_allreduce_kernel, dim3(ctas), dim3(128), 0, stream, P, rank, world, flags_off, ctl_off, slot, data_off, n / 4);
}

// dst[0:n] = reduce over the ranks of src[0:n] (device pointers local to this rank, may alias); op 0 fp32 sum, 1 fp32 max, 2 fp64 sum.
extern "C" int sgqn_p2p_small(const void* const* bases, int rank, int world, long long flags_off, long long ctl_off, long long small_off,
                              int slot, const void* src, void* dst, int n, int op, void* stream) {
    Peers P;
    int rc = fill_peers(&P, bases, rank, world);
    if (rc) return rc;
    if (n < 1 || n > (op == 2 ? 16 : 32) || op < 0 || op > 2 || slot < 0 || slot >= kSlots) return (int)cudaErrorInvalidValue;
    return launch_pdl(p2p_small_kernel, dim3(1), dim3(64), 0, stream, P, rank, world, flags_off, ctl_off, small_off, slot, src, dst, n, op);
}
