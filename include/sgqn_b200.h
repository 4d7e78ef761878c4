/* sgqn_b200.h -- C ABI of libsgqn_b200.so: the sm_100a kernels behind the SGSAC update path.
 *
 * The reference (gferraro2019/SGQN-CARLA) has no FFI: its boundary is the Python agent API
 * (`make_agent(...)` -> `update(replay_buffer, L, step)`, `select_action`, `sample_action`;
 * src/algorithms/factory.py:22-23, sac.py:95-105,160-169, sgsac.py:169-185).  The host-side mirror of
 * that API lives in `sgqn-carla_b200/` (Python, like the reference) and drives these entry points
 * through ctypes.  Every function:
 *   - takes raw DEVICE pointers, sizes and scalars only (no torch types), plus a `cudaStream_t` as `void*`;
 *   - launches asynchronously on that stream, allocates nothing, never synchronises the host;
 *   - returns 0 or the `cudaError_t` of the launch.
 * Citations `file:line` are relative to /root/reference/src and name what each entry point replaces.
 *
 * Layout conventions: observations are fp32 NCHW (B,9,H,W) with values 0..255 (what
 * ReplayBuffer.sample returns, utils.py:185-198); every feature map after the first conv is fp32 NHWC;
 * 3x3 conv weights (layers 2..11 of SharedCNN and the decoder convs) are stored [Cout][ky][kx][Cin]; the first
 * conv keeps the reference layout [Cout][Cin][ky][kx]; Linear weights keep the reference layout [out][in].
 */
#ifndef SGQN_B200_H
#define SGQN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int sgqn_abi_version(void);

/* ---- replay sampling: ReplayBuffer.sample / sample_drq (utils.py:124-135,158-171,185-198) with
 *      augmentations.random_crop (augmentations.py:236-264, mode 0) or random_shift (:229-233, mode 1) fused.
 *      frames: uint8 ring [F][3][Hs][Hs]; fidx: int32 [capacity][6] = frame slots of (obs f0..f2, next f0..f2);
 *      idxs: int64 [B]; offs: int32 [2][B][2] (row, col) offsets for obs / next_obs, or NULL. */
int sgqn_replay_gather(const uint8_t* frames, const int32_t* fidx, const int64_t* idxs, const int32_t* offs, float* obs,
                       float* next_obs, int B, int Hs, int Ho, int mode, int pad, void* stream);
/* raw copy of the six uint8 frames of every sampled transition into a device staging ring [B][6][fbytes] (prefetch of the next
 * batch from a pinned-host frame ring; the staging ring is then read with fidx = arange(6B), idxs = arange(B)) */
int sgqn_frames_copy(const uint8_t* frames, const int32_t* fidx, const int64_t* idxs, uint8_t* dst, int B, int fbytes, void* stream);
int sgqn_take_rows(const float* src, const int64_t* idxs, float* dst, int B, int width, void* stream);
/* random_crop / random_shift on a materialised fp32 batch; offs int32 [B][2] */
int sgqn_crop_shift(const float* x, const int32_t* offs, float* y, int B, int C, int Hs, int Ho, int mode, int pad, void* stream);
int sgqn_zero(void* p, long long bytes, void* stream);

/* ---- Linear layers (modules.py:107,194-198,239-243,318).  Strides (ld*, *bs) in elements; `batch` runs
 *      independent problems (the Q1/Q2 pair) through one launch.
 *      fwd:   y[M,N] = act(x)[M,K] * w[N,K]^T + bias      (relu_in: ReLU applied to x while loading;
 *             splitk 1: split-K with atomic accumulation into a caller-zeroed y; 2: y is zero-filled here first)
 *      dgrad: dx[M,K] = dy[M,N] * w[N,K], optionally masked by zmask (mode 1: *1[z>0]; mode 2 guided:
 *             relu(.)*1[z>0], captum GuidedBackprop, rl_utils.py:35-39); accumulate 1: atomic += (split-K allowed);
 *             2: dx zero-filled here, then split-K atomics (modes 0/1 only)
 *      wgrad: dw[N,K] += dy^T * act(x);  db[N] += colsum(dy)   (atomic; caller zero-fills) */
int sgqn_linear_fwd(const float* x, int ldx, long long xbs, const float* w, long long wbs, const float* bias, long long bbs,
                    float* y, int ldy, long long ybs, int M, int N, int K, int relu_in, int batch, int splitk, void* stream);
int sgqn_linear_dgrad(const float* dy, int lddy, long long dybs, const float* w, long long wbs, const float* zmask, int ldm,
                      long long mbs, float* dx, int lddx, long long dxbs, int M, int N, int K, int mode, int accumulate,
                      int batch, void* stream);
int sgqn_linear_wgrad(const float* x, int ldx, long long xbs, const float* dy, int lddy, long long dybs, float* dw,
                      long long dwbs, float* db, long long dbbs, int M, int N, int K, int relu_in, int batch, void* stream);
int sgqn_colsum(const float* x, int ld, int M, int N, float* out, void* stream);
/* The same three operations on tcgen05 with split-precision ("3xTF32") operands: x = tf32(x) + tf32(x - tf32(x)),
 * A.B ~= As.Bb + Ab.Bs + Ab.Bb accumulated in fp32 -- fp32-grade accuracy (the reference's nn.Linear runs in fp32) at
 * tensor-core speed.  Same arguments and semantics; every leading dimension / batch stride must be a multiple of 4
 * floats and every base pointer 16-byte aligned (TMA), otherwise cudaErrorInvalidValue. */
int sgqn_linear_fwd_tc(const float* x, int ldx, long long xbs, const float* w, long long wbs, const float* bias, long long bbs,
                       float* y, int ldy, long long ybs, int M, int N, int K, int relu_in, int batch, int splitk, void* stream);
int sgqn_linear_dgrad_tc(const float* dy, int lddy, long long dybs, const float* w, long long wbs, const float* zmask, int ldm,
                         long long mbs, float* dx, int lddx, long long dxbs, int M, int N, int K, int mode, int accumulate,
                         int batch, void* stream);
int sgqn_linear_wgrad_tc(const float* x, int ldx, long long xbs, const float* dy, int lddy, long long dybs, float* dw,
                         long long dwbs, float* db, long long dbbs, int M, int N, int K, int relu_in, int batch, void* stream);

/* ---- 3x3 convolutions on NHWC fp32 (SharedCNN layers 2..11: modules.py:144-146, valid, stride 1; decoder
 *      convs: modules.py:319-326, pad 1, nearest x2 upsample of the input fused via up=2).
 *      fwd:   y[B][Ho][Wo][Cout] = conv(act(up(x))) + bias,  Ho = Hs*up + 2*pad - 2
 *      dgrad: dx[B][Hl][Wl][Cin] gradient w.r.t. the conv's logical input, masked per `mode` by `mask`
 *      wgrad: dw[Cout][3][3][Cin] += ..., db[Cout] += ...     (atomic; caller zero-fills) */
int sgqn_conv_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Hs, int Ws, int Cin, int Cout,
                  int pad, int up, int relu_in, int flags /* bit0: ReLU on the output, bit1: round it to TF32 */, void* stream);
int sgqn_conv_dgrad(const float* dy, const float* w, const float* mask, float* dx, int B, int Hl, int Wl, int Cin, int Cout,
                    int pad, int mode, void* stream);
int sgqn_conv_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int Hs, int Ws, int Cin, int Cout, int pad,
                    int up, int relu_in, int dy_border, void* stream);
/* first encoder conv (modules.py:139-142: CenterCrop(84) -> x/255 -> Conv2d(Cin,Cout,3,stride=2)) on NCHW obs */
int sgqn_conv1_fwd(const float* obs, const float* w, const float* bias, float* y, int B, int Hin, int Cin, int Cout,
                   int flags /* bit0 ReLU, bit1 TF32 round, bit2: 2 extra (untouched) rows per sample in y */, void* stream);
int sgqn_conv1_wgrad(const float* obs, const float* dy, float* dw, float* db, int B, int Hin, int Cin, int Cout, void* stream);
int sgqn_conv1_dgrad(const float* dy, const float* w, float* dobs, int B, int Cin, int Cout, void* stream);
/*      the same three through a materialised im2col matrix col[B*41*41][84] (81 real columns, /255 applied): the index
 *      arithmetic is paid once per observation batch, forward / weight gradient / data gradient become plain GEMMs */
int sgqn_conv1_im2col(const float* obs, float* col, int B, int Hin, void* stream);
int sgqn_conv1_im2col96(const float* obs, float* col, int B, int Hin, void* stream);   /* col[.][96], TF32-rounded (tcgen05 path) */
int sgqn_conv1_weights_prep(const float* w, float* wp /* [32][96] */, float* wd /* optional transpose [96][32] */, void* stream);
/* observation gradient dobs[B][9][84][84] gathered from dcol[B*1681][pitch] = d(act_0) * W (col index ci*9+ky*3+kx), / 255 */
int sgqn_conv1_col2im(const float* dcol, int pitch, float* dobs, int B, void* stream);
int sgqn_conv1_fwd_col(const float* col, const float* w, const float* bias, float* y, int B, int flags, void* stream);
int sgqn_conv1_wgrad_col(const float* col, const float* dy, float* dw, float* db, int B, void* stream);
int sgqn_conv1_dgrad_col(const float* dy, const float* w, float* dcol, float* dobs, int B, void* stream);
/* ---- tcgen05 / TMEM / TMA implicit-GEMM 3x3 conv, 32 -> 32 channels, TF32 (SharedCNN layers 2..11, forward and data
 *      gradient; conv_tc.cu).  x [B][Hr][Wp][32] pitch-linear; w: TF32-rounded operand copy [32][9][32] made by
 *      sgqn_conv_weights_prep (wf: forward, wd: flipped+transposed for the data gradient).  Output (b,y,x), y < Hv,
 *      x < Wv = sum over taps of x-row q + ky*Wp + kx + shift, q = (b*Hr + y)*Wp + x; it goes to
 *      out[((b*Hq + y+oy)*Wq + x+ox)*32].  flags: bit0 ReLU, bit1 round output to TF32, bits 2-3 mask mode (1 plain ReLU
 *      backward, 2 guided) with the mask value of (b,y,x) at mask[((b*Hm + y)*Wm + x)*32].
 *      flags bit 4 ("old weights"): the kernel is a programmatic dependent launch; with this bit it fetches `w` BEFORE
 *      waiting for the previous kernel of the stream (under that kernel's last tiles).  The caller thereby guarantees that
 *      `w` was written at least two launches earlier on the stream, or on another stream joined by an event -- true for the
 *      operand copies of sgqn_conv_weights_prep, refreshed once per optimiser step, never by the producer of `x`. */
int sgqn_conv_tc(const float* x, const float* w, const float* bias, const float* mask, float* out,
                 float* dbias /* optional: dbias[32] += per-channel sum of the outputs written (atomic) */, int B, int Hr, int Wp,
                 int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm, int Wm, int flags, void* stream);
/*      A CHAIN of such convs -- the ten SharedCNN layers 2..11 of one encoder pass (modules.py:144-146), or their data
 *      gradients in reverse order -- as ONE persistent launch (conv_chain.cu): the tiles of all layers form one ticketed
 *      list and a tile of layer i+1 waits, per tile, for the tiles of layer i that produce its input rows, so there is no
 *      pipeline fill / drain between layers and a layer's input is still in L2.  layers[i] has the fields of one
 *      sgqn_conv_tc call; layers[i+1].x must be layers[i].out with layers[i]'s output geometry (Hq, Wq) as its input
 *      geometry (Hr, Wp).  Results are bit-identical to n_layers sgqn_conv_tc calls (dbias: up to summation order).
 *      ws: >= 4 + (number of tiles) ints, zero-filled once by the caller, owned by the launches of ONE stream. */
typedef struct sgqn_conv_layer {
    const float* x; const float* w; const float* bias; const float* mask; float* out; float* dbias;
    int B, Hr, Wp, Hv, Wv, shift, Hq, Wq, oy, ox, Hm, Wm, flags;
} sgqn_conv_layer;
int sgqn_conv_chain(const sgqn_conv_layer* layers, int n_layers, int* ws, long long ws_ints, void* stream);
/*      weight gradient of the same convs on tcgen05 (MN-major TF32 operands, reduction over pixels, one
 *      red.global.add per CTA and element): x, dy [B][Hr][Wp][32] share one geometry, dy zero outside its valid region */
int sgqn_conv_wgrad_tc(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, void* stream);
int sgqn_conv_weights_prep(const float* w, long long lstride, float* wf, float* wd, int n_layers, void* stream);
int sgqn_pad_copy(const float* src, float* dst, int B, int H, int W, int C, int Hq, int Wq, int oy, int ox,
                  int flags /* bit0 TF32 round, bit1 ReLU */, void* stream);
/* ---- generalised tcgen05 convs for the AttributionDecoder (modules.py:319-326; conv_tcg.cu): Cin = 32*k, Cout in {32,64,128},
 *      weights streamed per (channel chunk, tap); forward and data gradient through sgqn_conv_tcg (flags bit4: scatter every
 *      output to the 2x2 block of the next layer's zero-bordered input = fused ReLU + nearest x2 upsample), weight gradient
 *      through sgqn_conv_wgrad_tcg; sgqn_pool2_bwd is the backward of the fused upsample + ReLU. */
int sgqn_conv_tcg(const float* x, const float* wop, const float* bias, const float* mask, float* out, int B, int Hr, int Wp, int Cin,
                  int Cout, int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm, int Wm, int flags, void* stream);
int sgqn_conv_tcg_taps(const float* x, const float* wop, const float* bias, const float* mask, float* out, int B, int Hr, int Wp,
                       int Cin, int Cout, int Hv, int Wv, int shift, int Hq, int Wq, int oy, int ox, int Hm, int Wm, int flags,
                       int ntaps /* 9: 3x3 conv, 1: per-position GEMM (the first conv on its im2col matrix) */, void* stream);
int sgqn_gemm_wgrad_tcg(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta, int tb,
                        int ntaps, int kvalid, void* stream);
/* NormalizeImg + first conv (stride 2) + ReLU as one tcgen05 kernel whose im2col tile is built in shared memory (modules.py:86-93,
 * 143-144): obs (B,9,Hin,Hin) fp32 0..255 -> out [B][43][41][32] pitch-linear, TF32-rounded.  w1p from sgqn_conv1_weights_prep.
 * col (optional, for the weight gradient): the im2col matrix [B*1681][96] of the samples >= col_row0, written by TMA. */
int sgqn_conv1_fused_tc(const float* obs, const float* w1p, const float* bias, float* out, float* col, int B, int Hin, int col_row0,
                        void* stream);
/* ... and its data gradient to the observation (last step of compute_attribution, rl_utils.py:57-62): d(act_0) compact
 * [B][41][41][32] x w1d [96][32] on tcgen05, gathered to dobs (B,9,84,84) through shared memory (no dcol matrix in HBM). */
int sgqn_conv1_dgrad_fused_tc(const float* d, const float* w1d, float* dobs, int B, void* stream);
int sgqn_conv_weights_prep_g(const float* w, float* wf, float* wd, int Cout, int Cin, int Cout_real, void* stream);
int sgqn_conv_wgrad_tcg(const float* x, const float* dy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta, int tb,
                        void* stream);
int sgqn_conv_wgrad_tcg_ld(const float* x, const float* dy, int ldy, float* dw, int B, int Hr, int Wp, int Cin, int Cout, int ta,
                           int tb, void* stream);   /* dy = a Cout-column block of rows that are ldy floats apart */
int sgqn_pool2_bwd(const float* dup, const float* src, float* dst, int B, int H, int W, int C, void* stream);
/* Sub-pixel ("phase") form of conv3x3(pad 1) after F.upsample(x, 2) (modules.py:333-337 then :324-326): a 3x3 pad-1 conv at LOW
 * resolution with 4*Cg output channels -- phase p = 2a+b, channels [p*Cg, p*Cg+Cout_real) = output pixel (2y+a, 2x+b) -- run
 * through sgqn_conv_tcg / sgqn_conv_wgrad_tcg; the upsampled tensor is never materialised.  prep: w [>=Cout_real][9][Cin] ->
 * wf [4Cg][9][Cin] (forward operand), wd [Cin][9][4Cg] (data-gradient operand), bphi [4Cg]; fold: the chain rule back to the
 * reference's dW [Cout_real][9][Cin] / db (+=). */
int sgqn_conv_weights_prep_phase(const float* w, const float* bias, float* wf, float* wd, float* bphi, int Cin, int Cout_real,
                                 int Cg, void* stream);
int sgqn_conv_phase_fold(const float* dwphi, const float* dbphi, float* dw, float* db, int Cin, int Cout_real, int Cg, void* stream);
/* backward of F.upsample(x, 2) followed by ReLU mask of the pre-upsample activation (modules.py:333-337) */
int sgqn_upsample2_bwd(const float* dup, const float* act, float* dx, int B, int Hs, int Ws, int C, void* stream);

/* ---- saliency: compute_attribution_mask (rl_utils.py:76-82) fused with the mask application of
 *      update_critic (sgsac.py:67-70); mask uint8 [B][3][HW] (frame mask, the reference repeats it over the 3
 *      channels of a frame); masked_obs may be NULL (mask only).  u: 1 float (device).
 *      sgqn_minmax writes out4 = {min, max, -min, max}: the second pair is what ONE max-all-reduce turns into the global
 *      batch's pair in the data-parallel configuration; sgqn_attribution_mask reads minmax[0..1] as {min, max}, or as
 *      {-min, max} when minmax_neg != 0. */
int sgqn_minmax(const float* x, long long n, float* scratch /* >= 592 floats */, float* out4, void* stream);
int sgqn_attribution_mask(const float* grad, const float* obs, const float* minmax, const float* u, float quantile,
                          uint8_t* mask, float* masked_obs, int B, int HW, int minmax_neg, void* stream);
/* random_overlay (augmentations.py:79-99): 'carla' pool of uint8 frames [N][3][HW]; float places images in [0,1]: a batch
 * [B][3][HW] (ids NULL) or rows ids[b] of a device-resident pool [N][3][HW] */
int sgqn_overlay_u8(const float* obs, const uint8_t* pool, const int64_t* ids, float one_minus_alpha, float alpha, float* out,
                    int B, int HW, void* stream);
int sgqn_overlay_f32(const float* obs, const float* imgs, const int64_t* ids, float one_minus_alpha, float alpha, float* out,
                     int B, int HW, void* stream);

/* ---- heads and losses */
int sgqn_ln_tanh_fwd(const float* z, const float* gamma, const float* beta, float* h, int ldh, int M, int P, void* stream);
int sgqn_ln_tanh_bwd(const float* dh, int lddh, const float* z, const float* h, int ldh, const float* gamma, float* dz,
                     float* dgamma, float* dbeta, int M, int P, void* stream);
int sgqn_set_cols(float* dst, int ld, int col0, const float* src, int lds, int M, int n, void* stream);
int sgqn_actor_head_fwd(const float* raw, const float* noise, float lmin, float lmax, float* mu_t, float* pi_t, int ldpi,
                        float* log_pi, float* log_std, int M, int A, void* stream);
/* Bg: the global batch the actor-loss mean runs over (M = this shard's rows; Bg <= 0 means M) */
int sgqn_actor_head_bwd(const float* raw, const float* noise, const float* dpi, int lddpi, const double* log_alpha, float lmin,
                        float lmax, float* draw, int M, int A, int Bg, void* stream);
int sgqn_critic_loss(const float* q, long long qs, const float* tq1, const float* tq2, const float* next_log_pi,
                     const float* reward, const float* not_done, const double* log_alpha, float discount, int mode, float wa,
                     float wb, float* target_q, float* dq, float* loss, int B, int Bg, void* stream);
int sgqn_actor_loss(const float* q, long long qs, const float* log_pi, const double* log_alpha, float target_entropy, float* dq,
                    float* out3, double* alpha_grad, int B, int Bg, void* stream);
/* BCE-with-logits of the 9 real channels against the attribution mask (sgsac.py:163-167); logits / dlogits are NHWC with
 * Cs >= 9 stored channels.  Channels 9..11 of dlogits are written as zero; with Cs >= 12 the padding channels >= 12 are not
 * touched (keep them zero). */
int sgqn_bce(const float* logits, const uint8_t* mask, float* loss, float* dlogits, int B, int H, int W, int Hq, int Wq, int oy,
             int ox, int Cs, int Bg, int round_out, void* stream);
/* the same on the phase layout of the logits: [B][Hq][Wq][4][16] at H/2 x W/2 pixels (see sgqn_conv_weights_prep_phase) */
int sgqn_bce_phase(const float* logits, const uint8_t* mask, float* loss, float* dlogits, int B, int H, int W, int Hq, int Wq,
                   int oy, int ox, int Bg, int round_out, void* stream);

/* CURL (curl.py:35-37, modules.py:270-281): cross entropy of the (B,B) logits z_a W z_pos^T against the diagonal;
 * *loss += mean (caller zero-fills), dlogits = (softmax - I) / Bg */
int sgqn_ce_diag(const float* logits, int ld, float* loss, float* dlogits, int lddl, int B, int Bg, void* stream);

/* PAD (pad.py:42-43): F.mse_loss(pred, target) over (rows_global x width) elements; *loss += (caller zero-fills), dpred */
int sgqn_mse_loss(const float* pred, const float* target, float* loss, float* dpred, int rows, int width, int rows_global,
                  void* stream);

/* SODA (soda.py:41-49; SODAMLP, modules.py:116-129): BatchNorm1d in training mode (batch mean / biased variance, eps 1e-5)
 * fused with the ReLU that follows it, forward (stats = {mean[P], rstd[P]}) and backward (dgamma / dbeta +=, caller zero-fills);
 * mse(normalize(h0), normalize(h1)) with its gradient w.r.t. h0 */
int sgqn_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int M, int P, void* stream);
int sgqn_bn_relu_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stats, float* dx,
                     float* dgamma, float* dbeta, int M, int P, void* stream);
int sgqn_soda_loss(const float* h0, const float* h1, float* loss, float* dh0, int M, int P, int M_global, void* stream);

/* ---- optimiser: torch.optim.Adam (sac.py:60-68, sgsac.py:35-39) over a flat range, soft target update
 *      (utils.py:31-33, sac.py:153-158) fused when target != NULL; weight_decay = torch's L2 form (grad += wd * p;
 *      critic_weight_decay, sac.py:63-65) */
int sgqn_adam_prep(int* step, float* bc, double b1, double b2, void* stream);
int sgqn_adam(float* p, const float* g, float* m, float* v, long long n, const float* bc, float lr, float one_minus_b1, float b2,
              float one_minus_b2, float eps, float* target, long long n_tau0, float tau0, float tau1, float weight_decay,
              void* stream);
int sgqn_ema(const float* p, float* target, long long n, long long n_tau0, float tau0, float tau1, void* stream);
int sgqn_alpha_adam(double* log_alpha, const double* grad, double* st, int* step, double lr, double b1, double b2, double eps,
                    void* stream);
/* all random draws of one update (numpy idxs utils.py:127, python random sgsac.py:68 / augmentations.py:70, torch
 * randn_like modules.py:219, crop offsets augmentations.py:255-256) from one Philox launch.  seed_u keys the fill scalar u
 * alone: data-parallel ranks draw their own indices / noise (seed) but ONE u per global batch (sgsac.py:68-70) */
int sgqn_rng_step(unsigned long long seed, unsigned long long* counter, const int* n_valid, int64_t* idxs, int64_t* overlay_ids,
                  int pool_n, int32_t* offs, int off_n, float* noise_next, float* noise_pi, float* u, int B, int A,
                  unsigned long long seed_u, void* stream);

/* ---- gradient exchange of the batch-sharded update over NVLink peer memory (new functionality; the reference is
 *      single-GPU and has no counterpart).  bases: HOST array of `world` device pointers = the base of every rank's symmetric arena
 *      (same layout everywhere, peer-mapped); the arena starts with a header of flags | control words | small slots whose sizes the
 *      sgqn_p2p_layout() reports (zero-filled once by the host); offsets are bytes from the arena base.  A slot (0..7) is used
 *      from one stream per rank and every rank issues a slot's calls in the same order.
 *      allreduce_sum: in place over arena[data_off : data_off + 4n], two-shot (rank r reduces slice r in rank order and writes it to
 *      every rank), `ctas` (<= 160) CTAs of 128 threads, no shared memory; with a staging area (stage_off >= 0: 2 x 8 x stage_stride
 *      bytes inside the arena) ranges of <= stage_stride bytes are pushed to every peer and summed locally behind ONE barrier.
 *      small: dst[0:n] = reduction over ranks of src[0:n]
 *      (op 0 fp32 sum, 1 fp32 max, 2 fp64 sum; n <= 32 / 16), one 64-thread CTA. */
int sgqn_p2p_layout(long long* out3);     /* HOST pointer: {flag block, control block, small slots} sizes in bytes */
int sgqn_p2p_allreduce_sum(const void* const* bases, int rank, int world, long long flags_off, long long ctl_off, int slot,
                           long long data_off, long long n, int ctas, long long stage_off, long long stage_stride, void* stream);
int sgqn_p2p_small(const void* const* bases, int rank, int world, long long flags_off, long long ctl_off, long long small_off, int slot,
                   const void* src, void* dst, int n, int op, void* stream);

#ifdef __cplusplus
}
#endif
#endif
