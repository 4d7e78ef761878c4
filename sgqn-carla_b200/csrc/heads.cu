// Head and loss kernels of the SGSAC update: RLProjection LayerNorm+Tanh (modules.py:102-113), the Actor's
// squashed-Gaussian head (modules.py:20-33,212-232), critic TD / consistency loss (sgsac.py:52-74, svea.py:19-47),
// actor + alpha loss (sac.py:125-151) and the mask BCE (sgsac.py:163-167).  Tiny tensors: latency bound.
#include "common.cuh"
#include "../../include/sgqn_b200.h"

// ---------------------------------------------------------------- LayerNorm + tanh, one warp per row (P <= 1024)
__global__ void __launch_bounds__(256)
ln_tanh_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ h, int ldh, int M, int P) {
    pdl_wait();
    pdl_launch();
    int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* zr = z + (size_t)row * P;
    float s = 0.f;
    for (int i = lane; i < P; i += 32) s += zr[i];
    float mean = warp_sum(s) / (float)P;
    float v = 0.f;
    for (int i = lane; i < P; i += 32) { float d = zr[i] - mean; v += d * d; }
    float rstd = rsqrtf(warp_sum(v) / (float)P + 1e-5f);
    for (int i = lane; i < P; i += 32)
        h[(size_t)row * ldh + i] = tanhf((zr[i] - mean) * rstd * gamma[i] + beta[i]);
}

extern "C" int sgqn_ln_tanh_fwd(const float* z, const float* gamma, const float* beta, float* h, int ldh, int M, int P,
                                void* stream) {
    if (M <= 0) return 0;
    { int rc_ = launch_pdl(ln_tanh_fwd_kernel, dim3(cdiv(M, 8)), dim3(256), 0, (cudaStream_t)stream, z, gamma, beta, h, ldh, M, P); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// dz = LN^T( dh * (1 - h^2) ), dgamma / dbeta accumulated with atomics (pass nullptr to skip)
__global__ void __launch_bounds__(256)
ln_tanh_bwd_kernel(const float* __restrict__ dh, int lddh, const float* __restrict__ z, const float* __restrict__ h, int ldh,
                   const float* __restrict__ gamma, float* __restrict__ dz, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, int M, int P) {
    pdl_wait();
    pdl_launch();
    int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* zr = z + (size_t)row * P;
    float s = 0.f;
    for (int i = lane; i < P; i += 32) s += zr[i];
    float mean = warp_sum(s) / (float)P;
    float v = 0.f;
    for (int i = lane; i < P; i += 32) { float d = zr[i] - mean; v += d * d; }
    float rstd = rsqrtf(warp_sum(v) / (float)P + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < P; i += 32) {
        float hv = h[(size_t)row * ldh + i];
        float dt = dh[(size_t)row * lddh + i] * (1.f - hv * hv);
        float xh = (zr[i] - mean) * rstd;
        float dxh = dt * gamma[i];
        s1 += dxh; s2 += dxh * xh;
        if (dgamma) { atomicAdd(dgamma + i, dt * xh); atomicAdd(dbeta + i, dt); }
    }
    s1 = warp_sum(s1) / (float)P; s2 = warp_sum(s2) / (float)P;
    for (int i = lane; i < P; i += 32) {
        float hv = h[(size_t)row * ldh + i];
        float dt = dh[(size_t)row * lddh + i] * (1.f - hv * hv);
        float xh = (zr[i] - mean) * rstd;
        dz[(size_t)row * P + i] = rstd * (dt * gamma[i] - s1 - xh * s2);
    }
}

extern "C" int sgqn_ln_tanh_bwd(const float* dh, int lddh, const float* z, const float* h, int ldh, const float* gamma,
                                float* dz, float* dgamma, float* dbeta, int M, int P, void* stream) {
    if (M <= 0) return 0;
    { int rc_ = launch_pdl(ln_tanh_bwd_kernel, dim3(cdiv(M, 8)), dim3(256), 0, (cudaStream_t)stream, dh, lddh, z, h, ldh, gamma, dz, dgamma, dbeta, M, P); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// dst[:, col0:col0+n] = src[:, :n]
__global__ void set_cols_kernel(float* __restrict__ dst, int ld, int col0, const float* __restrict__ src, int lds, int M, int n) {
    pdl_wait();
    pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * n) return;
    int r = i / n, c = i - r * n;
    dst[(size_t)r * ld + col0 + c] = src[(size_t)r * lds + c];
}
extern "C" int sgqn_set_cols(float* dst, int ld, int col0, const float* src, int lds, int M, int n, void* stream) {
    if (M * n <= 0) return 0;
    { int rc_ = launch_pdl(set_cols_kernel, dim3(cdiv(M * n, 256)), dim3(256), 0, (cudaStream_t)stream, dst, ld, col0, src, lds, M, n); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- Actor head
// raw (M, 2A) = [mu | log_std_raw]; outputs tanh(mu), tanh(pi), log_pi, log_std   (modules.py:212-232)
__global__ void actor_head_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ noise, float lmin, float lmax,
                                      float* __restrict__ mu_t, float* __restrict__ pi_t, int ldpi, float* __restrict__ log_pi,
                                      float* __restrict__ log_std, int M, int A) {
    pdl_wait();
    pdl_launch();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    float resid = 0.f, corr = 0.f;
    for (int a = 0; a < A; ++a) {
        float mu = raw[(size_t)r * 2 * A + a];
        float ls = tanhf(raw[(size_t)r * 2 * A + A + a]);
        ls = lmin + 0.5f * (lmax - lmin) * (ls + 1.f);
        if (log_std) log_std[(size_t)r * A + a] = ls;
        if (mu_t) mu_t[(size_t)r * A + a] = tanhf(mu);
        if (noise) {
            float n = noise[(size_t)r * A + a];
            float p = tanhf(mu + n * expf(ls));
            if (pi_t) pi_t[(size_t)r * ldpi + a] = p;
            resid += -0.5f * n * n - ls;
            corr += logf(fmaxf(1.f - p * p, 0.f) + 1e-6f);
        }
    }
    if (log_pi && noise) log_pi[r] = resid - 0.5f * 1.8378770664093453f * (float)A - corr;
}

extern "C" int sgqn_actor_head_fwd(const float* raw, const float* noise, float lmin, float lmax, float* mu_t, float* pi_t,
                                   int ldpi, float* log_pi, float* log_std, int M, int A, void* stream) {
    if (M <= 0) return 0;
    { int rc_ = launch_pdl(actor_head_fwd_kernel, dim3(cdiv(M, 128)), dim3(128), 0, (cudaStream_t)stream, raw, noise, lmin, lmax, mu_t, pi_t, ldpi, log_pi, log_std, M, A); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// d raw given dL/d pi_t (from the Q heads) and dL/d log_pi = alpha / Bg  (sac.py:129-130; Bg = the GLOBAL batch the
// mean runs over: a data-parallel shard of M rows passes the global size, like the loss kernels, and gradients are summed)
__global__ void actor_head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                      const float* __restrict__ dpi, int lddpi, const double* __restrict__ log_alpha, float lmin,
                                      float lmax, float* __restrict__ draw, int M, int A, int Bg) {
    pdl_wait();
    pdl_launch();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    float glp = (float)exp(*log_alpha) / (float)Bg;
    for (int a = 0; a < A; ++a) {
        float mu = raw[(size_t)r * 2 * A + a];
        float t = tanhf(raw[(size_t)r * 2 * A + A + a]);
        float ls = lmin + 0.5f * (lmax - lmin) * (t + 1.f);
        float sd = expf(ls), n = noise[(size_t)r * A + a];
        float p = tanhf(mu + n * sd);
        float om = 1.f - p * p;
        float g = dpi[(size_t)r * lddpi + a];
        if (om > 0.f) g += glp * 2.f * p / (om + 1e-6f);          // - d/dpi log(relu(1-pi^2)+1e-6)
        float gpre = g * om;                                       // tanh'
        draw[(size_t)r * 2 * A + a] = gpre;
        float gls = gpre * n * sd - glp;                           // via pi and via gaussian_logprob
        draw[(size_t)r * 2 * A + A + a] = gls * 0.5f * (lmax - lmin) * (1.f - t * t);
    }
}

extern "C" int sgqn_actor_head_bwd(const float* raw, const float* noise, const float* dpi, int lddpi, const double* log_alpha,
                                   float lmin, float lmax, float* draw, int M, int A, int Bg, void* stream) {
    if (M <= 0) return 0;
    { int rc_ = launch_pdl(actor_head_bwd_kernel, dim3(cdiv(M, 128)), dim3(128), 0, (cudaStream_t)stream, raw, noise, dpi, lddpi, log_alpha, lmin, lmax, draw, M, A,
                                                                          Bg > 0 ? Bg : M); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- critic loss
// q: [2 heads][R] (head stride qs).  Bg = global batch the means are taken over (data-parallel shards pass the
// global B and sum gradients across ranks).
// mode 0 (sac.py:114-117)   R = B      : mse(q1,tq)+mse(q2,tq)
// mode 1 (sgsac.py:59-74)   R = 2B     : rows [clean | masked]: + 0.5*(mse(q1,mq1)+mse(q2,mq2)), grads to both branches
// mode 2 (svea.py:25-45)    R = 2B     : rows [obs | aug]: wa*(mse on obs rows) + wb*(mse on aug rows)
__global__ void __launch_bounds__(256)
critic_loss_kernel(const float* __restrict__ q, long long qs, const float* __restrict__ tq1, const float* __restrict__ tq2,
                   const float* __restrict__ next_log_pi, const float* __restrict__ reward, const float* __restrict__ not_done,
                   const double* __restrict__ log_alpha, float discount, int mode, float wa, float wb, float* __restrict__ target_q,
                   float* __restrict__ dq, float* __restrict__ loss, int B, int Bg) {
    pdl_wait();
    pdl_launch();
    __shared__ float sh[33];
    float alpha = (float)exp(*log_alpha);
    float acc = 0.f;
    const float invB = 1.f / (float)Bg;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float tv = fminf(tq1[b], tq2[b]) - alpha * next_log_pi[b];
        float tq = reward[b] + not_done[b] * discount * tv;
        target_q[b] = tq;
        for (int hd = 0; hd < 2; ++hd) {
            const float* qh = q + hd * qs; float* dqh = dq + hd * qs;
            float e = qh[b] - tq;
            if (mode == 0) { acc += e * e; dqh[b] = 2.f * invB * e; }
            else if (mode == 1) {
                float c = qh[b] - qh[B + b];
                acc += e * e + 0.5f * c * c;
                dqh[b] = 2.f * invB * e + invB * c;
                dqh[B + b] = -invB * c;
            } else {
                float e2 = qh[B + b] - tq;
                acc += wa * e * e + wb * e2 * e2;
                dqh[b] = 2.f * invB * wa * e; dqh[B + b] = 2.f * invB * wb * e2;
            }
        }
    }
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) *loss = tot * invB;
}

extern "C" int sgqn_critic_loss(const float* q, long long qs, const float* tq1, const float* tq2, const float* next_log_pi,
                                const float* reward, const float* not_done, const double* log_alpha, float discount, int mode,
                                float wa, float wb, float* target_q, float* dq, float* loss, int B, int Bg, void* stream) {
    { int rc_ = launch_pdl(critic_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, q, qs, tq1, tq2, next_log_pi, reward, not_done, log_alpha, discount,
                                                            mode, wa, wb, target_q, dq, loss, B, Bg); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- actor + alpha loss (sac.py:125-151)
// out[0] = actor_loss, out[1] = alpha_loss, out[2] = alpha ; alpha_grad (fp64) = d alpha_loss / d log_alpha
__global__ void __launch_bounds__(256)
actor_loss_kernel(const float* __restrict__ q, long long qs, const float* __restrict__ log_pi, const double* __restrict__ log_alpha,
                  float target_entropy, float* __restrict__ dq, float* __restrict__ out, double* __restrict__ alpha_grad, int B,
                  int Bg) {
    pdl_wait();
    pdl_launch();
    __shared__ float sh[33];
    double alpha_d = exp(*log_alpha);
    float alpha = (float)alpha_d;
    const float invB = 1.f / (float)Bg;
    float a1 = 0.f, a2 = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float q1 = q[b], q2 = q[qs + b];
        a1 += alpha * log_pi[b] - fminf(q1, q2);
        a2 += -log_pi[b] - target_entropy;
        float g1 = q1 < q2 ? 1.f : (q1 == q2 ? 0.5f : 0.f);
        dq[b] = -invB * g1; dq[qs + b] = -invB * (1.f - g1);
    }
    float s1 = block_sum(a1, sh);
    float s2 = block_sum(a2, sh);
    if (threadIdx.x == 0) {
        out[0] = s1 * invB;
        out[1] = alpha * s2 * invB;
        out[2] = alpha;
        *alpha_grad = (double)(s2 * invB) * alpha_d;
    }
}

extern "C" int sgqn_actor_loss(const float* q, long long qs, const float* log_pi, const double* log_alpha, float target_entropy,
                               float* dq, float* out, double* alpha_grad, int B, int Bg, void* stream) {
    { int rc_ = launch_pdl(actor_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, q, qs, log_pi, log_alpha, target_entropy, dq, out, alpha_grad, B, Bg); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- BCE-with-logits against the attribution mask (sgsac.py:163-167)
// logits NHWC with Cs >= 9 stored channels (extra ones are padding) in a [B][Hq][Wq][Cs] buffer, pixel (y,x) at row y+oy,
// col x+ox; mask [B][3][H*W] uint8 (frame f covers channels 3f..3f+2).  dlogits has the logits' layout.
__global__ void __launch_bounds__(256)
bce_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, float* __restrict__ loss,
           float* __restrict__ dlogits, int H, int W, int Hq, int Wq, int oy, int ox, int Cs, long long npix, float inv_n,
           int round_out) {
    __shared__ float sh[33];
    float acc = 0.f;
    const int HW = H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        int b = (int)(p / HW), i = (int)(p - (long long)b * HW);
        int y = i / W, xx = i - y * W;
        size_t o = (((size_t)b * Hq + y + oy) * Wq + xx + ox) * Cs;
        const float* x = logits + o;
        float* d = dlogits + o;
        for (int c = 0; c < Cs; ++c) {
            if (c >= 9) { d[c] = 0.f; continue; }
            float yv = (float)mask[((size_t)b * 3 + c / 3) * HW + i];
            float v = x[c];
            acc += fmaxf(v, 0.f) - v * yv + log1pf(expf(-fabsf(v)));
            float g = (1.f / (1.f + expf(-v)) - yv) * inv_n;
            if (round_out) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(g)); g = __uint_as_float(r); }
            d[c] = g;
        }
    }
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss, tot * inv_n);
}

// Vectorised variant for Cs % 4 == 0, Cs >= 12: three lanes per pixel, one float4 chunk (channels 4k..4k+3) each, so the
// 9 real channels are read and written as 48 contiguous bytes per pixel.  Channels >= 12 of dlogits are NOT written: the
// caller keeps them zero (they are zero from allocation on and nothing else writes the buffer).
__global__ void __launch_bounds__(256)
bce_vec_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, float* __restrict__ loss,
               float* __restrict__ dlogits, int H, int W, int Hq, int Wq, int oy, int ox, int Cs, long long nitems, float inv_n,
               int round_out) {
    __shared__ float sh[33];
    float acc = 0.f;
    const int HW = H * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nitems; t += (long long)gridDim.x * blockDim.x) {
        const long long p = t / 3; const int k = (int)(t - p * 3);
        int b = (int)(p / HW), i = (int)(p - (long long)b * HW);
        int y = i / W, xx = i - y * W;
        size_t o = (((size_t)b * Hq + y + oy) * Wq + xx + ox) * Cs + 4 * k;
        const float4 xv = __ldg(reinterpret_cast<const float4*>(logits + o));
        const float x[4] = {xv.x, xv.y, xv.z, xv.w};
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * k + e;
            g[e] = 0.f;
            if (c < 9) {
                float yv = (float)mask[((size_t)b * 3 + c / 3) * HW + i];
                float v = x[e];
                acc += fmaxf(v, 0.f) - v * yv + log1pf(expf(-fabsf(v)));
                float gg = (1.f / (1.f + expf(-v)) - yv) * inv_n;
                if (round_out) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(gg)); gg = __uint_as_float(r); }
                g[e] = gg;
            }
        }
        *reinterpret_cast<float4*>(dlogits + o) = make_float4(g[0], g[1], g[2], g[3]);
    }
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss, tot * inv_n);
}

// The same loss on the sub-pixel ("phase") layout of the logits (conv_tcg.cu, conv_weights_prep_phase): buffers are
// [B][Hq][Wq][4 phases][16] at LOW resolution (H/2 x W/2 pixels, pixel (y,x) at row y+oy, col x+ox); phase p = 2a+b,
// channel c < 9 of low-res pixel (y,x) is logit c of output pixel (2y+a, 2x+b).  One 16-byte chunk per work item, 12 of the
// 16 chunks of a pixel (channels 12..15 of each phase are padding: never read, never written, zero in dlogits).
__global__ void __launch_bounds__(256)
bce_phase_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, float* __restrict__ loss,
                 float* __restrict__ dlogits, int H, int W, int Hq, int Wq, int oy, int ox, long long nitems, float inv_n, int round_out) {
    pdl_wait();
    pdl_launch();
    __shared__ float sh[33];
    float acc = 0.f;
    const int Hl = H >> 1, Wl = W >> 1, HW = H * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nitems; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % 3); long long u = t / 3; const int ph = (int)(u & 3); u >>= 2;
        const int xl = (int)(u % Wl); u /= Wl; const int yl = (int)(u % Hl); const int b = (int)(u / Hl);
        const int i = (2 * yl + (ph >> 1)) * W + 2 * xl + (ph & 1);
        const size_t o = (((size_t)b * Hq + yl + oy) * Wq + xl + ox) * 64 + 16 * ph + 4 * k;
        const float4 xv = __ldg(reinterpret_cast<const float4*>(logits + o));
        const float x[4] = {xv.x, xv.y, xv.z, xv.w};
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * k + e;
            g[e] = 0.f;
            if (c < 9) {
                float yv = (float)mask[((size_t)b * 3 + c / 3) * HW + i];
                float v = x[e];
                acc += fmaxf(v, 0.f) - v * yv + log1pf(expf(-fabsf(v)));
                float gg = (1.f / (1.f + expf(-v)) - yv) * inv_n;
                if (round_out) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(gg)); gg = __uint_as_float(r); }
                g[e] = gg;
            }
        }
        *reinterpret_cast<float4*>(dlogits + o) = make_float4(g[0], g[1], g[2], g[3]);
    }
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss, tot * inv_n);
}

extern "C" int sgqn_bce_phase(const float* logits, const uint8_t* mask, float* loss, float* dlogits, int B, int H, int W, int Hq,
                              int Wq, int oy, int ox, int Bg, int round_out, void* stream) {
    if ((H | W) & 1) return (int)cudaErrorInvalidValue;
    long long nitems = (long long)B * (H / 2) * (W / 2) * 12;
    if (nitems <= 0) return 0;
    float inv_n = 1.0f / ((float)Bg * 9.0f * (float)(H * W));
    int grid = (int)(cdivll(nitems, 256) < 4736 ? cdivll(nitems, 256) : 4736);
    { int rc_ = launch_pdl(bce_phase_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, logits, mask, loss, dlogits, H, W, Hq, Wq, oy, ox, nitems, inv_n, round_out); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

extern "C" int sgqn_bce(const float* logits, const uint8_t* mask, float* loss, float* dlogits, int B, int H, int W, int Hq, int Wq,
                        int oy, int ox, int Cs, int Bg, int round_out, void* stream) {
    long long npix = (long long)B * H * W;
    if (npix <= 0) return 0;
    float inv_n = 1.0f / ((float)Bg * 9.0f * (float)(H * W));
    if (Cs >= 12 && (Cs & 3) == 0 && (((size_t)logits | (size_t)dlogits) & 15) == 0) {
        long long nitems = npix * 3;
        int grid = (int)(cdivll(nitems, 256) < 4736 ? cdivll(nitems, 256) : 4736);
        bce_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, mask, loss, dlogits, H, W, Hq, Wq, oy, ox, Cs, nitems, inv_n, round_out);
        return SGQN_CHECK_LAUNCH();
    }
    int grid = (int)(cdivll(npix, 256) < 1184 ? cdivll(npix, 256) : 1184);
    bce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, mask, loss, dlogits, H, W, Hq, Wq, oy, ox, Cs, npix, inv_n, round_out);
    return SGQN_CHECK_LAUNCH();
}


// ---------------------------------------------------------------- CURL contrastive loss (curl.py:35-37, modules.py:270-281)
// logits [B][ld] = z_a W z_pos^T; label of row i = i.  loss += sum_i (logsumexp_j l_ij - l_ii) / Bg;  dlogits = (softmax - I) / Bg.
// (The reference subtracts the row maximum first; the cross entropy does not change under a per-row shift, nor does its gradient.)
// One warp per row.
__global__ void __launch_bounds__(128) ce_diag_kernel(const float* __restrict__ logits, int ld, float* __restrict__ loss,
                                                      float* __restrict__ dlogits, int lddl, int B, float inv_n) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* l = logits + (size_t)row * ld;
    float m = -INFINITY;
    for (int j = lane; j < B; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < B; j += 32) s += expf(l[j] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    float* d = dlogits + (size_t)row * lddl;
    for (int j = lane; j < B; j += 32) d[j] = (expf(l[j] - lse) - (j == row ? 1.f : 0.f)) * inv_n;
    if (lane == 0) atomicAdd(loss, (lse - l[row]) * inv_n);
}
extern "C" int sgqn_ce_diag(const float* logits, int ld, float* loss, float* dlogits, int lddl, int B, int Bg, void* stream) {
    if (B <= 0) return 0;
    ce_diag_kernel<<<cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(logits, ld, loss, dlogits, lddl, B, 1.0f / (float)Bg);
    return SGQN_CHECK_LAUNCH();
}


// ---------------------------------------------------------------- mean-squared error (PAD inverse dynamics, pad.py:42-43)
// *loss += sum (pred - target)^2 * inv_n ; dpred = 2 (pred - target) * inv_n ; inv_n = 1 / (global rows * width)
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       float* __restrict__ loss, float* __restrict__ dpred, int n, float inv_n) {
    __shared__ float sh[33];
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float d = pred[i] - target[i];
        acc += d * d;
        dpred[i] = 2.f * d * inv_n;
    }
    const float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss, tot * inv_n);
}
extern "C" int sgqn_mse_loss(const float* pred, const float* target, float* loss, float* dpred, int rows, int width, int rows_global,
                             void* stream) {
    const int n = rows * width;
    if (n <= 0) return 0;
    const int grid = cdiv(n, 256) < 64 ? cdiv(n, 256) : 64;
    mse_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, loss, dpred, n, 1.0f / ((float)rows_global * (float)width));
    return SGQN_CHECK_LAUNCH();
}


// ---------------------------------------------------------------- SODA (soda.py:41-49, modules.py:116-129)
// BatchNorm1d in training mode over x[M][P] (+ ReLU): y = relu(gamma * (x - mean) * rstd + beta), biased batch variance, eps 1e-5;
// stats[0..P) = mean, stats[P..2P) = rstd for the backward pass.  One warp per feature (M <= a few thousand rows: tiny).
__global__ void __launch_bounds__(128) bn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ y,
                                                          float* __restrict__ stats, int M, int P) {
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= P) return;
    float s = 0.f;
    for (int r = lane; r < M; r += 32) s += x[(size_t)r * P + f];
    const float mean = warp_sum(s) / (float)M;
    float v = 0.f;
    for (int r = lane; r < M; r += 32) { const float d = x[(size_t)r * P + f] - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / (float)M + 1e-5f);
    const float g = gamma[f], b = beta[f];
    for (int r = lane; r < M; r += 32) y[(size_t)r * P + f] = fmaxf((x[(size_t)r * P + f] - mean) * rstd * g + b, 0.f);
    if (lane == 0) { stats[f] = mean; stats[P + f] = rstd; }
}
// backward of the above: dy arrives for the ReLU output; dx, dgamma += , dbeta += (caller zero-fills the two)
__global__ void __launch_bounds__(128) bn_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                          const float* __restrict__ y, const float* __restrict__ gamma,
                                                          const float* __restrict__ stats, float* __restrict__ dx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int P) {
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= P) return;
    const float mean = stats[f], rstd = stats[P + f], g = gamma[f];
    float s1 = 0.f, s2 = 0.f;
    for (int r = lane; r < M; r += 32) {
        const size_t i = (size_t)r * P + f;
        const float d = y[i] > 0.f ? dy[i] : 0.f;
        const float xh = (x[i] - mean) * rstd;
        s1 += d; s2 += d * xh;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    for (int r = lane; r < M; r += 32) {
        const size_t i = (size_t)r * P + f;
        const float d = y[i] > 0.f ? dy[i] : 0.f;
        const float xh = (x[i] - mean) * rstd;
        dx[i] = g * rstd * (d - s1 / (float)M - xh * s2 / (float)M);
    }
    if (lane == 0) { atomicAdd(dgamma + f, s2); atomicAdd(dbeta + f, s1); }
}
extern "C" int sgqn_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int M, int P, void* stream) {
    if (M <= 0 || P <= 0) return 0;
    bn_relu_fwd_kernel<<<cdiv(P, 4), 128, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, stats, M, P);
    return SGQN_CHECK_LAUNCH();
}
extern "C" int sgqn_bn_relu_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats, float* dx,
                                float* dgamma, float* dbeta, int M, int P, void* stream) {
    if (M <= 0 || P <= 0) return 0;
    bn_relu_bwd_kernel<<<cdiv(P, 4), 128, 0, (cudaStream_t)stream>>>(dy, x, y, gamma, stats, dx, dgamma, dbeta, M, P);
    return SGQN_CHECK_LAUNCH();
}

// loss = mse(normalize(h0), normalize(h1)) (F.normalize p=2, eps 1e-12; mean over M*P), dh0 = d loss / d h0.  One warp per row.
__global__ void __launch_bounds__(128) soda_loss_kernel(const float* __restrict__ h0, const float* __restrict__ h1,
                                                        float* __restrict__ loss, float* __restrict__ dh0, int M, int P, float inv_n) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* a = h0 + (size_t)row * P;
    const float* b = h1 + (size_t)row * P;
    float na = 0.f, nb = 0.f;
    for (int i = lane; i < P; i += 32) { na += a[i] * a[i]; nb += b[i] * b[i]; }
    na = fmaxf(sqrtf(warp_sum(na)), 1e-12f); nb = fmaxf(sqrtf(warp_sum(nb)), 1e-12f);
    float l = 0.f, dot = 0.f;                          // dot = <a_hat, dL/da_hat>
    for (int i = lane; i < P; i += 32) {
        const float ah = a[i] / na, d = ah - b[i] / nb;
        l += d * d; dot += ah * (2.f * d * inv_n);
    }
    l = warp_sum(l); dot = warp_sum(dot);
    for (int i = lane; i < P; i += 32) {
        const float ah = a[i] / na, d = ah - b[i] / nb;
        dh0[(size_t)row * P + i] = (2.f * d * inv_n - ah * dot) / na;
    }
    if (lane == 0) atomicAdd(loss, l * inv_n);
}
extern "C" int sgqn_soda_loss(const float* h0, const float* h1, float* loss, float* dh0, int M, int P, int M_global, void* stream) {
    if (M <= 0) return 0;
    soda_loss_kernel<<<cdiv(M, 4), 128, 0, (cudaStream_t)stream>>>(h0, h1, loss, dh0, M, P, 1.0f / ((float)M_global * (float)P));
    return SGQN_CHECK_LAUNCH();
}
