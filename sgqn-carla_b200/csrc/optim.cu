// Multi-tensor Adam over a flat parameter range with the soft target update (EMA) fused in
// (torch.optim.Adam single-tensor math, sac.py:60-68 / sgsac.py:35-39; utils.py:31-33 soft_update_params),
// the fp64 log_alpha Adam (sac.py:56,66-68) and the per-step device RNG (Philox4x32-10).
#include "common.cuh"
#include "../../include/sgqn_b200.h"

// state[0] = step count (as float bits of int), bc[0] = 1 - b1^t, bc[1] = sqrt(1 - b2^t)
__global__ void adam_prep_kernel(int* __restrict__ step, float* __restrict__ bc, double b1, double b2) {
    pdl_wait();
    pdl_launch();
    int t = *step + 1;
    *step = t;
    bc[0] = (float)(1.0 - pow(b1, (double)t));
    bc[1] = (float)sqrt(1.0 - pow(b2, (double)t));
}

extern "C" int sgqn_adam_prep(int* step, float* bc, double b1, double b2, void* stream) {
    { int rc_ = launch_pdl(adam_prep_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step, bc, b1, b2); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// p,g,m,v: n floats (n % 4 == 0, 16B aligned).  target != null: t = tau*p_new + (1-tau)*t with tau0 for
// elements [0, n_tau0) and tau1 for the rest.
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long long n4,
            const float* __restrict__ bc, float lr, float omb1, float b2, float omb2, float eps, float4* __restrict__ target,
            long long n4_tau0, float tau0, float tau1, float wd) {
    pdl_wait();
    pdl_launch();
    const float bc1 = bc[0], bc2s = bc[1];
    const float step_size = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = p[i], gv = __ldg(g + i), mv = m[i], vv = v[i];
        float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
        float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (wd != 0.f) gg[k] = gg[k] + wd * pp[k];                    // weight_decay: grad.add(param, alpha=wd)  (sac.py:63-65)
            mm[k] = mm[k] + omb1 * (gg[k] - mm[k]);                       // exp_avg.lerp_(grad, 1-beta1)
            vq[k] = vq[k] * b2 + omb2 * gg[k] * gg[k];                    // mul_(beta2).addcmul_(g, g, 1-beta2)
            float denom = sqrtf(vq[k]) / bc2s + eps;
            pp[k] = pp[k] - step_size * (mm[k] / denom);                  // addcdiv_(exp_avg, denom, -step_size)
        }
        p[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        v[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
        if (target) {
            float tau = i < n4_tau0 ? tau0 : tau1;
            float4 tv = target[i];
            tv.x = tau * pp[0] + (1.f - tau) * tv.x; tv.y = tau * pp[1] + (1.f - tau) * tv.y;
            tv.z = tau * pp[2] + (1.f - tau) * tv.z; tv.w = tau * pp[3] + (1.f - tau) * tv.w;
            target[i] = tv;
        }
    }
}

extern "C" int sgqn_adam(float* p, const float* g, float* m, float* v, long long n, const float* bc, float lr, float omb1,
                         float b2, float omb2, float eps, float* target, long long n_tau0, float tau0, float tau1, float weight_decay,
                         void* stream) {
    if (n <= 0) return 0;
    if ((n & 3) || (n_tau0 & 3)) return (int)cudaErrorInvalidValue;
    long long n4 = n / 4;
    int grid = (int)(cdivll(n4, 256) < 148 * 8 ? cdivll(n4, 256) : 148 * 8);
    { int rc_ = launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n4, bc, lr, omb1, b2, omb2,
                                                        eps, (float4*)target, n_tau0 / 4, tau0, tau1, weight_decay); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// EMA only (soft_update_params when it cannot ride on an Adam launch)
__global__ void __launch_bounds__(256)
ema_kernel(const float4* __restrict__ p, float4* __restrict__ target, long long n4, long long n4_tau0, float tau0, float tau1) {
    pdl_wait();
    pdl_launch();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float tau = i < n4_tau0 ? tau0 : tau1;
        float4 pv = __ldg(p + i), tv = target[i];
        tv.x = tau * pv.x + (1.f - tau) * tv.x; tv.y = tau * pv.y + (1.f - tau) * tv.y;
        tv.z = tau * pv.z + (1.f - tau) * tv.z; tv.w = tau * pv.w + (1.f - tau) * tv.w;
        target[i] = tv;
    }
}
extern "C" int sgqn_ema(const float* p, float* target, long long n, long long n_tau0, float tau0, float tau1, void* stream) {
    if (n <= 0) return 0;
    if ((n & 3) || (n_tau0 & 3)) return (int)cudaErrorInvalidValue;
    long long n4 = n / 4;
    int grid = (int)(cdivll(n4, 256) < 148 * 8 ? cdivll(n4, 256) : 148 * 8);
    { int rc_ = launch_pdl(ema_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)p, (float4*)target, n4, n_tau0 / 4, tau0, tau1); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// fp64 scalar Adam for log_alpha; st = {m, v}, step = device int
__global__ void alpha_adam_kernel(double* __restrict__ log_alpha, const double* __restrict__ grad, double* __restrict__ st,
                                  int* __restrict__ step, double lr, double b1, double b2, double eps) {
    pdl_wait();
    pdl_launch();
    int t = *step + 1; *step = t;
    double g = *grad;
    double m = st[0] + (1.0 - b1) * (g - st[0]);
    double v = st[1] * b2 + (1.0 - b2) * g * g;
    st[0] = m; st[1] = v;
    double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    double denom = sqrt(v) / sqrt(bc2) + eps;
    *log_alpha = *log_alpha - (lr / bc1) * (m / denom);
}
extern "C" int sgqn_alpha_adam(double* log_alpha, const double* grad, double* st, int* step, double lr, double b1, double b2,
                               double eps, void* stream) {
    { int rc_ = launch_pdl(alpha_adam_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, log_alpha, grad, st, step, lr, b1, b2, eps); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- per-step randomness
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

// One launch draws everything one update consumes.  Layout of the draw id space (one Philox block per id):
//   [0,B) idxs in [0,*n_valid) | [B,2B) overlay ids in [0,pool_n) | [2B,3B) offsets (4 ints in [0,off_n)) |
//   [3B, 3B+nA) normal pairs for noise_next / noise_pi | last: u
__global__ void rng_step_kernel(unsigned long long seed, unsigned long long* __restrict__ counter, const int* __restrict__ n_valid,
                                int64_t* __restrict__ idxs, int64_t* __restrict__ overlay_ids, int pool_n, int32_t* __restrict__ offs,
                                int off_n, float* __restrict__ noise_next, float* __restrict__ noise_pi, float* __restrict__ u,
                                int B, int A, unsigned long long seed_u) {
    pdl_wait();
    pdl_launch();
    const int nA = B * A;
    const int total = 3 * B + nA + 1;
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    unsigned long long ctr = *counter;
    uint32_t r[4];
    philox4x32_10((uint32_t)id, (uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5367514eu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    if (id < B) {
        unsigned long long x = ((unsigned long long)r[0] << 32) | r[1];
        if (idxs) idxs[id] = (int64_t)(x % (unsigned long long)max(*n_valid, 1));
    } else if (id < 2 * B) {
        if (overlay_ids) overlay_ids[id - B] = (int64_t)(r[0] % (uint32_t)max(pool_n, 1));
    } else if (id < 3 * B) {
        if (offs) {
            int b = id - 2 * B;                           // offs layout [2][B][2]
            offs[(0 * B + b) * 2 + 0] = (int)(r[0] % (uint32_t)max(off_n, 1));
            offs[(0 * B + b) * 2 + 1] = (int)(r[1] % (uint32_t)max(off_n, 1));
            offs[(1 * B + b) * 2 + 0] = (int)(r[2] % (uint32_t)max(off_n, 1));
            offs[(1 * B + b) * 2 + 1] = (int)(r[3] % (uint32_t)max(off_n, 1));
        }
    } else if (id < 3 * B + nA) {
        int i = id - 3 * B;
        float rad = sqrtf(-2.0f * logf(u01(r[0]))), ang = 6.283185307179586f * u01(r[1]);
        if (noise_next) noise_next[i] = rad * cosf(ang);
        float rad2 = sqrtf(-2.0f * logf(u01(r[2]))), ang2 = 6.283185307179586f * u01(r[3]);
        if (noise_pi) noise_pi[i] = rad2 * cosf(ang2);
    } else if (u) {
        // the fill scalar is ONE draw per global batch (sgsac.py:68-70): data-parallel ranks share seed_u and the counter
        philox4x32_10((uint32_t)id, (uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5367514eu, (uint32_t)seed_u, (uint32_t)(seed_u >> 32), r);
        *u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);   // [0,1)
    }
}
__global__ void rng_advance_kernel(unsigned long long* counter) {
    pdl_wait();
    pdl_launch();
    *counter += 1ull;
}

extern "C" int sgqn_rng_step(unsigned long long seed, unsigned long long* counter, const int* n_valid, int64_t* idxs,
                             int64_t* overlay_ids, int pool_n, int32_t* offs, int off_n, float* noise_next, float* noise_pi,
                             float* u, int B, int A, unsigned long long seed_u, void* stream) {
    int total = 3 * B + B * A + 1;
    { int rc_ = launch_pdl(rng_step_kernel, dim3(cdiv(total, 128)), dim3(128), 0, (cudaStream_t)stream, seed, counter, n_valid, idxs, overlay_ids, pool_n, offs, off_n,
                                                                     noise_next, noise_pi, u, B, A, seed_u); if (rc_) return rc_; }
    { int rc_ = launch_pdl(rng_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, counter); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}
