// Generic fp32 CUDA-core tile GEMM  C[i][j] (+)= sum_c A(i,c) * B(j,c)  with pluggable operand
// loaders and epilogue.  Every dense op of the SGSAC update that is not (yet) on the tcgen05
// path is an instantiation of this core: Linear fwd/dgrad/wgrad, implicit-GEMM conv
// fwd/dgrad/wgrad (NHWC activations), the stride-2 first conv on NCHW observations.
//
// Loader concept:
//   static constexpr bool kRowContig;
//   __device__ float4 fetch4(int row, int c, int rows, int cend, int batch) const;
//     kRowContig == false : returns (row, c..c+3)      ("C" loaders: contraction index contiguous)
//     kRowContig == true  : returns (row..row+3, c)    ("R" loaders: row index contiguous)
//   out-of-range elements must come back as 0.
#pragma once
#include "common.cuh"

namespace sgqn {

// ------------------------------------------------------------------ plain matrices
struct RowMajorC {            // element (row,c) at ptr[row*ld + c]
    static constexpr bool kRowContig = false;
    const float* ptr; int ld; long long bstride; int relu; int vec;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int batch) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        const float* p = ptr + (size_t)batch * bstride + (size_t)row * ld + c;
        if (vec && c + 3 < cend) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
            v.x = __ldg(p);
            if (c + 1 < cend) v.y = __ldg(p + 1);
            if (c + 2 < cend) v.z = __ldg(p + 2);
            if (c + 3 < cend) v.w = __ldg(p + 3);
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        return v;
    }
};

struct ColMajorR {            // element (row,c) at ptr[c*ld + row]
    static constexpr bool kRowContig = true;
    const float* ptr; int ld; long long bstride; int relu; int vec;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int batch) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        const float* p = ptr + (size_t)batch * bstride + (size_t)c * ld + row;
        if (vec && row + 3 < rows) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
            v.x = __ldg(p);
            if (row + 1 < rows) v.y = __ldg(p + 1);
            if (row + 2 < rows) v.z = __ldg(p + 2);
            if (row + 3 < rows) v.w = __ldg(p + 3);
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        return v;
    }
};

// element (row = channel, c = compact pixel of [B][Ho][Wo]) of a pixel-major tensor stored with its own pitch:
// ptr[((b*Hq + y + off)*Wq + x + off)*C + row]   (conv wgrad A operand when dY lives in a zero-bordered buffer)
struct PixRowsR {
    static constexpr bool kRowContig = true;
    const float* ptr; int C, Ho, Wo, Hq, Wq, off;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        int x = c % Wo; int t = c / Wo; int y = t % Ho; int b = t / Ho;
        const float* p = ptr + ((size_t)(b * Hq + y + off) * Wq + x + off) * C + row;
        if (row + 3 < rows) return __ldg(reinterpret_cast<const float4*>(p));
        v.x = __ldg(p);
        if (row + 1 < rows) v.y = __ldg(p + 1);
        if (row + 2 < rows) v.z = __ldg(p + 2);
        return v;
    }
};

// ------------------------------------------------------------------ 3x3 conv geometry (NHWC, C % 4 == 0)
struct ConvGeom {
    int Hs, Ws;      // stored source tensor [B][Hs][Ws][C]
    int up;          // logical source = stored * up (nearest upsample fused into the read), 1 or 2
    int Hr, Wr;      // pixel grid the GEMM index runs over, [B][Hr][Wr]
    int stride;      // forward stride (1, or 2 for the first encoder conv)
    int pad;
    int dgrad;       // 0: src = r*stride + k - pad ; 1: src = (r + pad - k) / stride (exact division only)
    int C;           // channels of the source tensor
    __device__ __forceinline__ bool src(int pix, int tap, int& off) const {
        int x = pix % Wr; int t = pix / Wr; int y = t % Hr; int b = t / Hr;
        int ky = tap / 3, kx = tap - ky * 3;
        int sy, sx;
        if (!dgrad) { sy = y * stride + ky - pad; sx = x * stride + kx - pad; }
        else {
            sy = y + pad - ky; sx = x + pad - kx;
            if (stride == 2) { if ((sy | sx) & 1) return false; sy >>= 1; sx >>= 1; }
        }
        if (sy < 0 || sx < 0 || sy >= Hs * up || sx >= Ws * up) return false;
        if (up == 2) { sy >>= 1; sx >>= 1; }
        off = ((b * Hs + sy) * Ws + sx) * C;
        return true;
    }
};

struct ConvPixC {             // A(row = pixel, c = tap*C + ch): conv fwd (src = input) / dgrad (src = dY)
    static constexpr bool kRowContig = false;
    const float* ptr; ConvGeom g; int relu;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        int tap = c / g.C, ch = c - tap * g.C, off;
        if (!g.src(row, tap, off)) return v;
        v = __ldg(reinterpret_cast<const float4*>(ptr + (size_t)off + ch));
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        return v;
    }
};

struct ConvPixR {             // B(row = tap*C + ch, c = pixel): conv wgrad (src = input activations)
    static constexpr bool kRowContig = true;
    const float* ptr; ConvGeom g; int relu;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        int tap = row / g.C, ch = row - tap * g.C, off;
        if (!g.src(c, tap, off)) return v;
        v = __ldg(reinterpret_cast<const float4*>(ptr + (size_t)off + ch));
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        return v;
    }
};

struct ConvWdgradR {          // B(row = ci, c = tap*Cout + co) = W[co][tap][ci]   (weights stored [Cout][9][Cin])
    static constexpr bool kRowContig = true;
    const float* w; int Cin, Cout;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows || c >= cend) return v;
        int tap = c / Cout, co = c - tap * Cout;
        return __ldg(reinterpret_cast<const float4*>(w + ((size_t)co * 9 + tap) * Cin + row));
    }
};

// ------------------------------------------------------------------ first encoder conv (NCHW fp32 obs, Cin = 9, stride 2)
struct Conv1Obs {             // value(pixel (b,y,x) on Ho x Ho, c = ci*9 + ky*3 + kx) = obs[b][ci][2y+ky+crop][2x+kx+crop] / 255
    const float* obs; int Hin, Ho, crop, Cin;
    __device__ __forceinline__ float val(int pix, int c) const {
        int x = pix % Ho; int t = pix / Ho; int y = t % Ho; int b = t / Ho;
        int ci = c / 9, r = c - ci * 9, ky = r / 3, kx = r - ky * 3;
        float v = __ldg(obs + ((size_t)(b * Cin + ci) * Hin + (2 * y + ky + crop)) * Hin + (2 * x + kx + crop));
        return __fdiv_rn(v, 255.0f);                      // modules.py:86-91 NormalizeImg: x / 255.0
    }
};
struct Conv1ObsC {            // A(row = pixel, c in [0, 9*Cin))
    static constexpr bool kRowContig = false;
    Conv1Obs o;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= rows) return v;
        if (c < cend) v.x = o.val(row, c);
        if (c + 1 < cend) v.y = o.val(row, c + 1);
        if (c + 2 < cend) v.z = o.val(row, c + 2);
        if (c + 3 < cend) v.w = o.val(row, c + 3);
        return v;
    }
};
struct Conv1ObsR {            // B(row = c81, c = pixel)
    static constexpr bool kRowContig = true;
    Conv1Obs o;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= cend) return v;
        if (row < rows) v.x = o.val(c, row);
        if (row + 1 < rows) v.y = o.val(c, row + 1);
        if (row + 2 < rows) v.z = o.val(c, row + 2);
        if (row + 3 < rows) v.w = o.val(c, row + 3);
        return v;
    }
};
struct Conv1WdgradR {         // B(row = ci, c = tap*Cout + co) = W1[co][ci][tap]   (reference layout [Cout][Cin][3][3])
    static constexpr bool kRowContig = true;
    const float* w; int Cin, Cout;
    __device__ __forceinline__ float4 fetch4(int row, int c, int rows, int cend, int) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= cend) return v;
        int tap = c / Cout, co = c - tap * Cout;
        const float* p = w + (size_t)co * Cin * 9 + tap;
        if (row < rows) v.x = __ldg(p + row * 9);
        if (row + 1 < rows) v.y = __ldg(p + (row + 1) * 9);
        if (row + 2 < rows) v.z = __ldg(p + (row + 2) * 9);
        if (row + 3 < rows) v.w = __ldg(p + (row + 3) * 9);
        return v;
    }
};

// ------------------------------------------------------------------ epilogues
struct EpStore {
    float* C; int ldc; long long bstride;
    const float* bias; long long bias_bstride;
    const float* mask; int ldm; long long mask_bstride;   // same index space as C
    int mode;        // 0 none, 1 v *= (mask > 0), 2 guided: v = (mask > 0) ? max(v, 0) : 0
    int atomic;      // 1: atomicAdd into C (caller zero-fills / accumulates)
    float scale;
    int post;        // bit0: ReLU on the output, bit1: round the output to TF32 (it feeds a tcgen05 TF32 conv)
    int grp, gpad;   // grp > 0: output row r is stored at row r + (r / grp) * gpad (extra rows per sample in C)
    __device__ __forceinline__ float finish(float v) const {
        if (post & 1) v = fmaxf(v, 0.f);
        if (post & 2) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
        return v;
    }
    template <int TM, int TN>
    __device__ __forceinline__ void store(float (&acc)[TM][TN], int i, int j, int M, int N, int batch, int split) const {
        float* Cb = C + (size_t)batch * bstride;
        const float* mb = mask ? mask + (size_t)batch * mask_bstride : nullptr;
        const float* bb = (bias && split == 0) ? bias + (size_t)batch * bias_bstride : nullptr;
        const bool vec = !atomic && TN == 4 && (j + 3 < N) && ((ldc & 3) == 0) && ((((uintptr_t)Cb) & 15) == 0) &&
                         (!mode || (((ldm & 3) == 0) && ((((uintptr_t)mb) & 15) == 0)));
#pragma unroll
        for (int m = 0; m < TM; ++m) {
            int row = i + m;
            if (row >= M) break;
            if (grp > 0) row += (row / grp) * gpad;
            if (vec) {
                float v[4];
#pragma unroll
                for (int n = 0; n < 4; ++n) v[n] = finish(acc[m][n] + (bb ? __ldg(bb + j + n) : 0.f));
                if (mode) {
                    float4 mk = __ldg(reinterpret_cast<const float4*>(mb + (size_t)row * ldm + j));
                    float mm[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        if (mode == 2) v[n] = fmaxf(v[n], 0.f);
                        v[n] = mm[n] > 0.f ? v[n] : 0.f;
                    }
                }
                *reinterpret_cast<float4*>(Cb + (size_t)row * ldc + j) =
                    make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
                continue;
            }
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                int col = j + n;
                if (col >= N) continue;
                float v = acc[m][n];
                if (bb) v += __ldg(bb + col);
                v = finish(v);
                if (mode) {
                    float mk = __ldg(mb + (size_t)row * ldm + col);
                    if (mode == 2) v = fmaxf(v, 0.f);
                    v = mk > 0.f ? v : 0.f;
                }
                v *= scale;
                float* dst = Cb + (size_t)row * ldc + col;
                if (atomic) atomicAdd(dst, v); else *dst = v;
            }
        }
    }
};

struct EpObsGrad {            // conv1 dgrad: rows (b,Y,X) on H x H, cols ci -> NCHW obs gradient, / 255
    float* out; int H, Cin;
    template <int TM, int TN>
    __device__ __forceinline__ void store(float (&acc)[TM][TN], int i, int j, int M, int N, int, int) const {
#pragma unroll
        for (int m = 0; m < TM; ++m) {
            int row = i + m;
            if (row >= M) break;
            int hw = H * H, b = row / hw, r = row - b * hw;
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                int col = j + n;
                if (col >= N) continue;
                out[((size_t)b * Cin + col) * hw + r] = __fdiv_rn(acc[m][n], 255.0f);
            }
        }
    }
};

// ------------------------------------------------------------------ the core
// One shared-memory buffer per operand + register prefetch: the global loads of tile t+1 are in flight while tile t is
// being multiplied (these kernels run at low occupancy, so the overlap has to come from inside the thread).
template <int BR, int BK, int NT>
struct TileRegs { static constexpr int G = (BR * BK / 4 + NT - 1) / NT; float4 v[G]; };

template <int BR, int BK, int NT, class L>
__device__ __forceinline__ void fetch_tile(TileRegs<BR, BK, NT>& t, const L& ld, int r0, int c0, int rows, int cend, int batch, int tid) {
    constexpr int G = BR * BK / 4;
#pragma unroll
    for (int u = 0; u < TileRegs<BR, BK, NT>::G; ++u) {
        int g = tid + u * NT;
        if (G % NT != 0 && g >= G) break;
        if constexpr (!L::kRowContig) {
            int r = g / (BK / 4), cg = g - r * (BK / 4);
            t.v[u] = ld.fetch4(r0 + r, c0 + cg * 4, rows, cend, batch);
        } else {
            int c = g / (BR / 4), rg = g - c * (BR / 4);
            t.v[u] = ld.fetch4(r0 + rg * 4, c0 + c, rows, cend, batch);
        }
    }
}

template <int BR, int BK, int NT, class L>
__device__ __forceinline__ void store_tile(float (*tile)[BR + 4], const TileRegs<BR, BK, NT>& t, int tid) {
    constexpr int G = BR * BK / 4;
#pragma unroll
    for (int u = 0; u < TileRegs<BR, BK, NT>::G; ++u) {
        int g = tid + u * NT;
        if (G % NT != 0 && g >= G) break;
        const float4 v = t.v[u];
        if constexpr (!L::kRowContig) {
            int r = g / (BK / 4), cg = g - r * (BK / 4);
            tile[cg * 4 + 0][r] = v.x; tile[cg * 4 + 1][r] = v.y; tile[cg * 4 + 2][r] = v.z; tile[cg * 4 + 3][r] = v.w;
        } else {
            int c = g / (BR / 4), rg = g - c * (BR / 4);
            *reinterpret_cast<float4*>(&tile[c][rg * 4]) = v;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN, class AL, class BL, class EP>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(AL al, BL bl, EP ep, int M, int N, int K, int nsplit, int klen) {
    constexpr int NT = (BM / TM) * (BN / TN);
    static_assert(TM % 4 == 0 && TN % 4 == 0 && BK % 4 == 0, "tile shape");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    pdl_wait();
    pdl_launch();
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int batch = blockIdx.z / nsplit, split = blockIdx.z - batch * nsplit;
    const int kb = split * klen;
    const int ke = min(K, kb + klen);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;

    TileRegs<BM, BK, NT> ra;
    TileRegs<BN, BK, NT> rb;
    if (kb < ke) {
        fetch_tile<BM, BK, NT, AL>(ra, al, i0, kb, M, ke, batch, tid);
        fetch_tile<BN, BK, NT, BL>(rb, bl, j0, kb, N, ke, batch, tid);
    }
    for (int c0 = kb; c0 < ke; c0 += BK) {
        store_tile<BM, BK, NT, AL>(As, ra, tid);
        store_tile<BN, BK, NT, BL>(Bs, rb, tid);
        __syncthreads();
        if (c0 + BK < ke) {
            fetch_tile<BM, BK, NT, AL>(ra, al, i0, c0 + BK, M, ke, batch, tid);
            fetch_tile<BN, BK, NT, BL>(rb, bl, j0, c0 + BK, N, ke, batch, tid);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int t = 0; t < TM / 4; ++t) {
                float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4 * t]);
                a[4 * t] = v.x; a[4 * t + 1] = v.y; a[4 * t + 2] = v.z; a[4 * t + 3] = v.w;
            }
#pragma unroll
            for (int t = 0; t < TN / 4; ++t) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + 4 * t]);
                b[4 * t] = v.x; b[4 * t + 1] = v.y; b[4 * t + 2] = v.z; b[4 * t + 3] = v.w;
            }
#pragma unroll
            for (int m = 0; m < TM; ++m)
#pragma unroll
                for (int n = 0; n < TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
        }
        __syncthreads();
    }
    ep.template store<TM, TN>(acc, i0 + ty * TM, j0 + tx * TN, M, N, batch, split);
}

// host-side launcher: picks the split count so the grid covers the SMs (~2 waves of 148)
template <int BM, int BN, int BK, int TM, int TN, class AL, class BL, class EP>
static int launch_gemm(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int batch, int max_split,
                       cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return 0;
    int tm = cdiv(M, BM), tn = cdiv(N, BN);
    long long tiles = (long long)tm * tn * batch;
    int nsplit = 1;
    if (max_split > 1) {
        nsplit = (int)cdivll(592, tiles);            // ~4 CTAs per SM: these kernels hide latency with occupancy only
        int kchunks = cdiv(K, BK * 2);                 // at least 2 BK steps per split
        if (nsplit > kchunks) nsplit = kchunks;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
    }
    int klen = cdiv(cdiv(K, nsplit), BK) * BK;
    nsplit = cdiv(K, klen);
    dim3 grid(tm, tn, batch * nsplit);
    return launch_pdl(gemm_simt_kernel<BM, BN, BK, TM, TN, AL, BL, EP>, grid, dim3((BM / TM) * (BN / TN)), 0, st, al, bl, ep, M, N, K,
                      nsplit, klen);
}

}  // namespace sgqn
