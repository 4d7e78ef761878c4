// Dense ops of the SGSAC update on the fp32 CUDA-core tile GEMM (gemm_simt.cuh):
// Linear fwd/dgrad/wgrad (RLProjection, Actor/Critic MLPs, decoder proj: modules.py:102-113,187-261,315-341),
// 3x3 conv fwd/dgrad/wgrad on NHWC activations (SharedCNN layers 2..11 and the AttributionDecoder
// convs: modules.py:132-152,315-341) and the stride-2 first conv on NCHW observations.
#include "gemm_simt.cuh"
#include "../../include/sgqn_b200.h"

using namespace sgqn;

static inline int aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// ---------------------------------------------------------------- column sums (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ x, int ld, int M, int N, float* __restrict__ out, int rows_per_block) {
    pdl_wait();
    pdl_launch();
    // block: 256 threads = 8 row-lanes x 32 columns
    __shared__ float sh[8][33];
    int col = blockIdx.y * 32 + (threadIdx.x & 31);
    int rl = threadIdx.x >> 5;
    int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (col < N)
        for (int r = r0 + rl; r < r1; r += 8) s += __ldg(x + (size_t)r * ld + col);
    sh[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        atomicAdd(out + col, t);
    }
}

static int launch_colsum(const float* x, int ld, int M, int N, float* out, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    int ctas_y = cdiv(N, 32);
    int want = cdiv(296, ctas_y);
    int rpb = cdiv(M, want);
    if (rpb < 64) rpb = 64;
    dim3 grid(cdiv(M, rpb), ctas_y);
    { int rc_ = launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, st, x, ld, M, N, out, rpb); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

extern "C" int sgqn_colsum(const float* x, int ld, int M, int N, float* out, void* stream) {
    return launch_colsum(x, ld, M, N, out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- Linear
extern "C" int sgqn_linear_fwd(const float* x, int ldx, long long xbs, const float* w, long long wbs, const float* bias,
                               long long bbs, float* y, int ldy, long long ybs, int M, int N, int K, int relu_in,
                               int batch, int splitk, void* stream) {
    RowMajorC a{x, ldx, xbs, relu_in, aligned16(x) && (ldx % 4 == 0) && (xbs % 4 == 0)};
    RowMajorC b{w, K, wbs, 0, aligned16(w) && (K % 4 == 0) && (wbs % 4 == 0)};
    EpStore ep{y, ldy, ybs, bias, bbs, nullptr, 0, 0, 0, splitk ? 1 : 0, 1.0f, 0, 0, 0};
    if (splitk == 2) {               // zero-fill here (strided rows, per batch), then split-K with atomics
        int rc = zero2d(y, ldy, ybs, M, N, batch, stream);
        if (rc) return rc;
    }
    return launch_gemm<64, 64, 16, 4, 4>(a, b, ep, M, N, K, batch, splitk ? 64 : 1, (cudaStream_t)stream);
}

extern "C" int sgqn_linear_dgrad(const float* dy, int lddy, long long dybs, const float* w, long long wbs,
                                 const float* zmask, int ldm, long long mbs, float* dx, int lddx, long long dxbs, int M,
                                 int N, int K, int mode, int accumulate, int batch, void* stream) {
    // dx[M,K] = dy[M,N] * W[N,K]   (contraction over the N output features)
    RowMajorC a{dy, lddy, dybs, 0, aligned16(dy) && (lddy % 4 == 0) && (dybs % 4 == 0)};
    ColMajorR b{w, K, wbs, 0, aligned16(w) && (K % 4 == 0) && (wbs % 4 == 0)};
    EpStore ep{dx, lddx, dxbs, nullptr, 0, zmask, ldm, mbs, mode, accumulate ? 1 : 0, 1.0f, 0, 0, 0};
    if (accumulate == 2) {           // zero-fill here, then split-K with atomics (the plain ReLU mask distributes over the sum)
        if (mode == 2) return (int)cudaErrorInvalidValue;
        int rc = zero2d(dx, lddx, dxbs, M, K, batch, stream);
        if (rc) return rc;
    }
    return launch_gemm<64, 64, 16, 4, 4>(a, b, ep, M, K, N, batch, accumulate ? 64 : 1, (cudaStream_t)stream);
}

extern "C" int sgqn_linear_wgrad(const float* x, int ldx, long long xbs, const float* dy, int lddy, long long dybs,
                                 float* dw, long long dwbs, float* db, long long dbbs, int M, int N, int K, int relu_in,
                                 int batch, void* stream) {
    // dw[N,K] += dy^T[N,M] * act(x)[M,K] ; db[N] += colsum(dy)
    ColMajorR a{dy, lddy, dybs, 0, aligned16(dy) && (lddy % 4 == 0) && (dybs % 4 == 0)};
    ColMajorR b{x, ldx, xbs, relu_in, aligned16(x) && (ldx % 4 == 0) && (xbs % 4 == 0)};
    EpStore ep{dw, K, dwbs, nullptr, 0, nullptr, 0, 0, 0, 1, 1.0f, 0, 0, 0};
    int rc = launch_gemm<64, 64, 16, 4, 4>(a, b, ep, N, K, M, batch, 64, (cudaStream_t)stream);
    if (rc) return rc;
    if (db)
        for (int bi = 0; bi < batch; ++bi) {
            rc = launch_colsum(dy + bi * dybs, lddy, M, N, db + bi * dbbs, (cudaStream_t)stream);
            if (rc) return rc;
        }
    return 0;
}

// ---------------------------------------------------------------- 3x3 conv, NHWC, weights [Cout][9][Cin]
extern "C" int sgqn_conv_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Hs, int Ws, int Cin,
                             int Cout, int pad, int up, int relu_in, int flags, void* stream) {
    if ((Cin & 3) || (Cout & 3) || (up != 1 && up != 2)) return (int)cudaErrorInvalidValue;
    int Ho = Hs * up + 2 * pad - 2, Wo = Ws * up + 2 * pad - 2;
    ConvGeom g{Hs, Ws, up, Ho, Wo, 1, pad, 0, Cin};
    ConvPixC a{x, g, relu_in};
    RowMajorC b{w, 9 * Cin, 0, 0, aligned16(w)};
    EpStore ep{y, Cout, 0, bias, 0, nullptr, 0, 0, 0, 0, 1.0f, flags & 3, 0, 0};
    int M = B * Ho * Wo, K = 9 * Cin;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout <= 16) return launch_gemm<128, 16, 32, 4, 4>(a, b, ep, M, Cout, K, 1, 1, st);
    if (Cout <= 32) return launch_gemm<128, 32, 32, 8, 4>(a, b, ep, M, Cout, K, 1, 1, st);
    return launch_gemm<128, 64, 32, 8, 4>(a, b, ep, M, Cout, K, 1, 1, st);
}

extern "C" int sgqn_conv_dgrad(const float* dy, const float* w, const float* mask, float* dx, int B, int Hl, int Wl,
                               int Cin, int Cout, int pad, int mode, void* stream) {
    // dx[B][Hl][Wl][Cin]: gradient w.r.t. the conv's logical input (after ReLU / upsample), optionally
    // masked by `mask` (same shape): mode 1 plain ReLU backward, mode 2 guided (captum GuidedBackprop).
    if ((Cin & 3) || (Cout & 3)) return (int)cudaErrorInvalidValue;
    int Ho = Hl + 2 * pad - 2, Wo = Wl + 2 * pad - 2;
    ConvGeom g{Ho, Wo, 1, Hl, Wl, 1, pad, 1, Cout};
    ConvPixC a{dy, g, 0};
    ConvWdgradR b{w, Cin, Cout};
    EpStore ep{dx, Cin, 0, nullptr, 0, mask, Cin, 0, mode, 0, 1.0f, 0, 0, 0};
    int M = B * Hl * Wl, K = 9 * Cout;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin <= 32) return launch_gemm<128, 32, 32, 8, 4>(a, b, ep, M, Cin, K, 1, 1, st);
    return launch_gemm<128, 64, 32, 8, 4>(a, b, ep, M, Cin, K, 1, 1, st);
}

extern "C" int sgqn_conv_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int Hs, int Ws, int Cin,
                               int Cout, int pad, int up, int relu_in, int dy_border, void* stream) {
    // dw[Cout][9][Cin] += sum_pix dy[pix][co] * act(x)[src(pix,tap)][ci] ; db[Cout] += sum_pix dy
    if ((Cin & 3) || (Cout & 3)) return (int)cudaErrorInvalidValue;
    int Ho = Hs * up + 2 * pad - 2, Wo = Ws * up + 2 * pad - 2;
    ConvGeom g{Hs, Ws, up, Ho, Wo, 1, pad, 0, Cin};
    int P = B * Ho * Wo;
    ConvPixR b{x, g, relu_in};
    EpStore ep{dw, 9 * Cin, 0, nullptr, 0, nullptr, 0, 0, 0, 1, 1.0f, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (dy_border > 0) {          // dy stored [B][Ho+2b][Wo+2b][Cout] with a zero border (tcgen05 dgrad layout)
        PixRowsR a{dy, Cout, Ho, Wo, Ho + 2 * dy_border, Wo + 2 * dy_border, dy_border};
        if (Cout <= 32) rc = launch_gemm<32, 64, 16, 4, 4>(a, b, ep, Cout, 9 * Cin, P, 1, 4096, st);
        else rc = launch_gemm<64, 64, 16, 4, 4>(a, b, ep, Cout, 9 * Cin, P, 1, 4096, st);
        if (rc) return rc;      // the border is zero, so the bias gradient may sum every stored row
        return db ? launch_colsum(dy, Cout, B * (Ho + 2 * dy_border) * (Wo + 2 * dy_border), Cout, db, st) : 0;
    }
    ColMajorR a{dy, Cout, 0, 0, aligned16(dy)};
    if (Cout <= 32) rc = launch_gemm<32, 64, 16, 4, 4>(a, b, ep, Cout, 9 * Cin, P, 1, 4096, st);
    else rc = launch_gemm<64, 64, 16, 4, 4>(a, b, ep, Cout, 9 * Cin, P, 1, 4096, st);
    if (rc) return rc;
    return db ? launch_colsum(dy, Cout, P, Cout, db, st) : 0;
}

// ---------------------------------------------------------------- first encoder conv (modules.py:142): NCHW obs, 9 -> 32, stride 2
extern "C" int sgqn_conv1_fwd(const float* obs, const float* w, const float* bias, float* y, int B, int Hin, int Cin,
                              int Cout, int flags, void* stream) {
    int crop = (Hin - 84) / 2;                               // CenterCrop(84), modules.py:70-83
    int Ho = (84 - 3) / 2 + 1;
    Conv1ObsC a{{obs, Hin, Ho, crop, Cin}};
    RowMajorC b{w, 9 * Cin, 0, 0, 0};
    EpStore ep{y, Cout, 0, bias, 0, nullptr, 0, 0, 0, 0, 1.0f, flags & 3, (flags & 4) ? Ho * Ho : 0, 2 * Ho};
    return launch_gemm<128, 32, 16, 8, 4>(a, b, ep, B * Ho * Ho, Cout, 9 * Cin, 1, 1, (cudaStream_t)stream);
}

extern "C" int sgqn_conv1_wgrad(const float* obs, const float* dy, float* dw, float* db, int B, int Hin, int Cin, int Cout,
                                void* stream) {
    int crop = (Hin - 84) / 2, Ho = 41, P = B * Ho * Ho;
    ColMajorR a{dy, Cout, 0, 0, aligned16(dy) && (Cout % 4 == 0)};
    Conv1ObsR b{{obs, Hin, Ho, crop, Cin}};
    EpStore ep{dw, 9 * Cin, 0, nullptr, 0, nullptr, 0, 0, 0, 1, 1.0f, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_gemm<32, 64, 16, 4, 4>(a, b, ep, Cout, 9 * Cin, P, 1, 4096, st);
    if (rc) return rc;
    return db ? launch_colsum(dy, Cout, P, Cout, db, st) : 0;
}

extern "C" int sgqn_conv1_dgrad(const float* dy, const float* w, float* dobs, int B, int Cin, int Cout, void* stream) {
    // d obs (NCHW, 84x84) = conv1^T(dy) / 255 ; only the attribution path needs it (rl_utils.py:35-39)
    ConvGeom g{41, 41, 1, 84, 84, 2, 0, 1, Cout};
    ConvPixC a{dy, g, 0};
    Conv1WdgradR b{w, Cin, Cout};
    EpObsGrad ep{dobs, 84, Cin};
    return launch_gemm<128, 16, 32, 4, 4>(a, b, ep, B * 84 * 84, Cin, 9 * Cout, 1, 1, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- nearest-upsample backward (+ ReLU mask of the pre-upsample activation)
__global__ void upsample2_bwd_kernel(const float4* __restrict__ dup, const float4* __restrict__ act, float4* __restrict__ dx,
                                     int Hs, int Ws, int C4, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % C4); long long t = i / C4;
    int x = (int)(t % Ws); t /= Ws;
    int y = (int)(t % Hs); int b = (int)(t / Hs);
    int Wl = 2 * Ws;
    size_t base = (((size_t)b * 2 * Hs + 2 * y) * Wl + 2 * x) * C4 + c;
    float4 a0 = __ldg(dup + base), a1 = __ldg(dup + base + C4), a2 = __ldg(dup + base + (size_t)Wl * C4),
           a3 = __ldg(dup + base + (size_t)Wl * C4 + C4);
    float4 m = __ldg(act + i);
    float4 r;
    r.x = m.x > 0.f ? (a0.x + a1.x) + (a2.x + a3.x) : 0.f;
    r.y = m.y > 0.f ? (a0.y + a1.y) + (a2.y + a3.y) : 0.f;
    r.z = m.z > 0.f ? (a0.z + a1.z) + (a2.z + a3.z) : 0.f;
    r.w = m.w > 0.f ? (a0.w + a1.w) + (a2.w + a3.w) : 0.f;
    dx[i] = r;
}

extern "C" int sgqn_upsample2_bwd(const float* dup, const float* act, float* dx, int B, int Hs, int Ws, int C, void* stream) {
    if (C & 3) return (int)cudaErrorInvalidValue;
    long long total = (long long)B * Hs * Ws * (C / 4);
    upsample2_bwd_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dup, (const float4*)act, (float4*)dx, Hs, Ws, C / 4, total);
    return SGQN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- first conv through a materialised im2col matrix
// col[pix][84]: columns 0..80 = obs[b][ci][2y+ky+crop][2x+kx+crop] / 255 (c = ci*9 + ky*3 + kx), 81..83 = 0.
// Forward, weight gradient and the attribution's data gradient of conv1 then are plain GEMMs over `col` (the
// per-element index arithmetic of the direct kernels above is paid once per observation batch instead of per use).
__global__ void conv1_im2col_kernel(const float* __restrict__ obs, float4* __restrict__ col, int Hin, int crop, long long total,
                                    int nc4, int round_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c4 = (int)(i % nc4); long long pix = i / nc4;
    int x = (int)(pix % 41); long long t = pix / 41; int y = (int)(t % 41); int b = (int)(t / 41);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int c = c4 * 4 + e;
        if (c < 81) {
            int ci = c / 9, r = c - ci * 9, ky = r / 3, kx = r - ky * 3;
            v[e] = __fdiv_rn(__ldg(obs + ((size_t)(b * 9 + ci) * Hin + (2 * y + ky + crop)) * Hin + (2 * x + kx + crop)), 255.0f);
        } else v[e] = 0.f;
        if (round_out) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v[e])); v[e] = __uint_as_float(r); }
    }
    col[i] = make_float4(v[0], v[1], v[2], v[3]);
}

extern "C" int sgqn_conv1_im2col(const float* obs, float* col, int B, int Hin, void* stream) {
    long long total = (long long)B * 1681 * 21;
    if (total <= 0) return 0;
    conv1_im2col_kernel<<<(unsigned)cdivll(total, 256), 256, 0, (cudaStream_t)stream>>>(obs, (float4*)col, Hin, (Hin - 84) / 2, total, 21, 0);
    return SGQN_CHECK_LAUNCH();
}

// tcgen05 variant: col[pix][96] (three 32-channel chunks), values rounded to TF32.  One CTA per output row (b, y):
// the 27 input rows (9 channels x 3 ky) it needs are staged in shared memory with coalesced loads, scaled and
// rounded once, and the 41 x 96 output row is written as one contiguous 15.7 KB run -- the kernel is bound by
// that write (82.6 MB at B=128), not by the stride-2 gather.
__global__ void __launch_bounds__(256) conv1_im2col96_rows_kernel(const float* __restrict__ obs, float4* __restrict__ col, int Hin, int crop) {
    __shared__ float s[27][84];
    const int y = blockIdx.x % 41, b = blockIdx.x / 41;
    for (int i = threadIdx.x; i < 27 * 84; i += 256) {
        int r = i / 84, c = i - r * 84, ci = r / 3, ky = r - ci * 3;
        float v = __fdiv_rn(__ldg(obs + ((size_t)(b * 9 + ci) * Hin + (2 * y + ky + crop)) * Hin + c + crop), 255.0f);
        uint32_t t; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
        s[r][c] = __uint_as_float(t);
    }
    __syncthreads();
    float4* dst = col + (size_t)(b * 1681 + y * 41) * 24;
    for (int i = threadIdx.x; i < 41 * 24; i += 256) {
        int x = i / 24, c4 = i - x * 24;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int c = c4 * 4 + e, r = c / 3, kx = c - r * 3;          // c = ci*9 + ky*3 + kx  =>  r = ci*3 + ky
            v[e] = c < 81 ? s[r][2 * x + kx] : 0.f;
        }
        dst[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}
extern "C" int sgqn_conv1_im2col96(const float* obs, float* col, int B, int Hin, void* stream) {
    if (B <= 0) return 0;
    conv1_im2col96_rows_kernel<<<(unsigned)(B * 41), 256, 0, (cudaStream_t)stream>>>(obs, (float4*)col, Hin, (Hin - 84) / 2);
    return SGQN_CHECK_LAUNCH();
}

// TF32 operand copy of the first conv's weights for the per-position GEMM: wp[32][96] = rna(w[32][81]), zero padded
// ... and, optionally, its transpose wd[96][32] = the operand of the data gradient dcol[pix][96] = d(act_0)[pix][32] * W
__global__ void conv1_weights_prep_kernel(const float* __restrict__ w, float* __restrict__ wp, float* __restrict__ wd) {
    pdl_wait();
    pdl_launch();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * 96) return;
    int co = i / 96, c = i - co * 96;
    float v = c < 81 ? w[co * 81 + c] : 0.f;
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    wp[i] = __uint_as_float(r);
    if (wd) wd[c * 32 + co] = __uint_as_float(r);
}
extern "C" int sgqn_conv1_weights_prep(const float* w, float* wp, float* wd, void* stream) {
    { int rc_ = launch_pdl(conv1_weights_prep_kernel, dim3(12), dim3(256), 0, (cudaStream_t)stream, w, wp, wd); if (rc_) return rc_; }
    return SGQN_CHECK_LAUNCH();
}

extern "C" int sgqn_conv1_fwd_col(const float* col, const float* w, const float* bias, float* y, int B, int flags, void* stream) {
    RowMajorC a{col, 84, 0, 0, aligned16(col)};
    RowMajorC b{w, 81, 0, 0, 0};
    EpStore ep{y, 32, 0, bias, 0, nullptr, 0, 0, 0, 0, 1.0f, flags & 3, (flags & 4) ? 1681 : 0, 82};
    return launch_gemm<128, 32, 16, 8, 4>(a, b, ep, B * 1681, 32, 81, 1, 1, (cudaStream_t)stream);
}

extern "C" int sgqn_conv1_wgrad_col(const float* col, const float* dy, float* dw, float* db, int B, void* stream) {
    int P = B * 1681;
    ColMajorR a{dy, 32, 0, 0, aligned16(dy)};
    ColMajorR b{col, 84, 0, 0, aligned16(col)};
    EpStore ep{dw, 81, 0, nullptr, 0, nullptr, 0, 0, 0, 1, 1.0f, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_gemm<32, 64, 16, 4, 4>(a, b, ep, 32, 81, P, 1, 4096, st);
    if (rc) return rc;
    return db ? launch_colsum(dy, 32, P, 32, db, st) : 0;
}

// d obs[b][ci][Y][X] = (1/255) * sum over windows (y,x,ky,kx) with 2y+ky = Y, 2x+kx = X of dcol[(b,y,x)][ci*9+ky*3+kx]
__global__ void conv1_col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dobs, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int X = (int)(i % 84); long long t = i / 84; int Y = (int)(t % 84); t /= 84; int ci = (int)(t % 9); int b = (int)(t / 9);
    float s = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        int yy = Y - ky;
        if (yy < 0 || (yy & 1) || (yy >> 1) >= 41) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            int xx = X - kx;
            if (xx < 0 || (xx & 1) || (xx >> 1) >= 41) continue;
            s += __ldg(dcol + ((size_t)(b * 41 + (yy >> 1)) * 41 + (xx >> 1)) * 84 + ci * 9 + ky * 3 + kx);
        }
    }
    dobs[i] = __fdiv_rn(s, 255.0f);
}

// The same gather with the two dcol rows an output row pair needs staged in shared memory: CTA (b, y) reads dcol rows
// y and y-1 (coalesced, 15.7 KB each at pitch 96) and writes observation rows Y = 2y (taps ky = 0 from row y, ky = 2 from
// row y-1) and Y = 2y+1 (ky = 1 from row y), all 9 channels.  Pixel pitch 97 floats in shared memory: consecutive x hit
// consecutive banks.
__global__ void __launch_bounds__(256) conv1_col2im_rows_kernel(const float* __restrict__ dcol, int pitch, float* __restrict__ dobs) {
    __shared__ float s[2][41 * 97];
    const int y = blockIdx.x % 42, b = blockIdx.x / 42;
    for (int r = 0; r < 2; ++r) {
        const int yy = y - r;
        const bool ok = yy >= 0 && yy < 41;
        const float4* src = reinterpret_cast<const float4*>(dcol + ((size_t)(b * 41 + (ok ? yy : 0)) * 41) * pitch);
        const int p4 = pitch / 4;
        for (int i = threadIdx.x; i < 41 * 21; i += 256) {          // 84 of the columns are enough (81 real)
            int x = i / 21, c4 = i - x * 21;
            float4 v = ok ? __ldg(src + x * p4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float* d = &s[r][x * 97 + c4 * 4];
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 9 * 84; i += 256) {
        const int X = i % 84; int t = i / 84; const int ci = t % 9; const int yo = t / 9;
        const int Y = 2 * y + yo;
        if (Y >= 84) continue;
        float acc = 0.f;
        // column taps: X even -> kx = 0 (x = X/2), kx = 2 (x = X/2 - 1); X odd -> kx = 1 (x = (X-1)/2)
        const int x0 = X >> 1;
        if (yo == 1) {                                   // ky = 1, row y
            if (X & 1) { if (x0 < 41) acc += s[0][x0 * 97 + ci * 9 + 3 + 1]; }
            else {
                if (x0 < 41) acc += s[0][x0 * 97 + ci * 9 + 3 + 0];
                if (x0 >= 1) acc += s[0][(x0 - 1) * 97 + ci * 9 + 3 + 2];
            }
        } else {                                         // ky = 0 from row y, ky = 2 from row y-1
            if (X & 1) { if (x0 < 41) acc += s[0][x0 * 97 + ci * 9 + 1] + s[1][x0 * 97 + ci * 9 + 6 + 1]; }
            else {
                if (x0 < 41) acc += s[0][x0 * 97 + ci * 9 + 0] + s[1][x0 * 97 + ci * 9 + 6 + 0];
                if (x0 >= 1) acc += s[0][(x0 - 1) * 97 + ci * 9 + 2] + s[1][(x0 - 1) * 97 + ci * 9 + 6 + 2];
            }
        }
        dobs[((size_t)(b * 9 + ci) * 84 + Y) * 84 + X] = __fdiv_rn(acc, 255.0f);
    }
}
// dobs[B][9][84][84] from dcol[B*1681][pitch] (pitch a multiple of 4, >= 84)
extern "C" int sgqn_conv1_col2im(const float* dcol, int pitch, float* dobs, int B, void* stream) {
    if (B <= 0) return 0;
    if ((pitch & 3) || pitch < 84) return (int)cudaErrorInvalidValue;
    conv1_col2im_rows_kernel<<<(unsigned)(B * 42), 256, 0, (cudaStream_t)stream>>>(dcol, pitch, dobs);
    return SGQN_CHECK_LAUNCH();
}

extern "C" int sgqn_conv1_dgrad_col(const float* dy, const float* w, float* dcol, float* dobs, int B, void* stream) {
    // dcol[pix][84] = dy[pix][32] * W1[32][81]  (cols 81..83 untouched), then gather back to the NCHW observation gradient
    RowMajorC a{dy, 32, 0, 0, aligned16(dy)};
    ColMajorR b{w, 81, 0, 0, 0};
    EpStore ep{dcol, 84, 0, nullptr, 0, nullptr, 0, 0, 0, 0, 1.0f, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_gemm<128, 32, 16, 8, 4>(a, b, ep, B * 1681, 81, 32, 1, 1, st);
    if (rc) return rc;
    long long total = (long long)B * 9 * 84 * 84;
    conv1_col2im_kernel<<<(unsigned)cdivll(total, 256), 256, 0, st>>>(dcol, dobs, total);
    return SGQN_CHECK_LAUNCH();
}
