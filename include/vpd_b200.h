/* vpd_b200 - C ABI of the B200-native VPD student hot path.
 *
 * The reference (jhong93/vpd) is pure Python/PyTorch and has no FFI layer; its
 * boundary for this path is the Python object API of models/rgb.py:46-86
 * (RGBF_EmbeddingModel), train_vpd_model.py:53-112 (ModelTrainer) and
 * apply_vpd_model.py:152-178. `vpd_b200/` mirrors that API in Python and binds
 * the functions below with ctypes (vpd_b200/_lib.py); INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer unless marked "host";
 *   - `stream` is a cudaStream_t passed as void*; no function synchronises it;
 *   - return value 0 = ok, non-zero = error, message via vpd_last_error()
 *     (thread-local, valid until the next failing call on that thread);
 *   - nothing here allocates device memory: callers own every buffer
 *     (vpd_net_workspace_bytes tells how much scratch a network needs);
 *   - activations inside the network are NHWC bf16, parameters and gradients
 *     are fp32 in the reference's own state_dict layout (OIHW conv weights).
 */
#ifndef VPD_B200_H_
#define VPD_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define VPD_API __attribute__((visibility("default")))
#else
#define VPD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

VPD_API const char* vpd_last_error(void);
VPD_API int vpd_abi_version(void);

/* Per-channel statistics accumulator (BatchNorm batch sums, BN-backward sums): two 64-bit
 * INTEGER limbs, value = hi * 2^-4 + lo * 2^-52. Kernels add to it with integer atomics,
 * which are associative: the totals - and every activation / gradient downstream of the
 * batch statistics - are bit-reproducible from run to run (fp64 atomics are not).
 * Buffers are arrays of these, zeroed by the caller where a function accumulates (+=);
 * vpd_b200/_lib.py has acc_zeros / acc_from_f64 / acc_to_f64 for the Python side. */
typedef struct vpd_stat_acc {
  int64_t hi, lo;
} vpd_stat_acc;

/* ---- K1: frame-batch assembly ------------------------------------------------
 * Replaces vpd_dataset/common.py:52-69 (_load_image/_load_flow arithmetic),
 * vpd_dataset/single_frame.py:168-206 (GenericDataset.__getitem__, deterministic
 * part: teacher-row select by flip bit, RGB+flow stack, horizontal flip with
 * flow-x negation) and :373-400 (FrameDataset.__getitem__, [orig, flipped]).
 *   rgb      uint8 [pool][H][W][3]  (RGB order, i.e. after cv2 BGR2RGB)
 *   flow     uint8 [pool][H][W][flow_channels] or NULL (RGB-only model)
 *   index    int32 [B] pool index per output frame, or NULL (frame b = pool b)
 *   flip     uint8 [B] flip bit per frame, or NULL (k == 1 only)
 *   teacher  fp32  [pool][teacher_rows][tdim] or NULL; row = flip bit
 *   mean,std host fp32 [3]
 *   k        1: one image per frame (flipped iff flip[b]); 2: [orig, flipped]
 * vpd_assemble_nchw writes the reference layout fp32 [B][k][C][H][W] (bit-exact);
 * vpd_assemble_stem writes the network's own input layout, bf16
 * [B*k][(H+7)/2][(W+9)/4][64]: the image zero-padded by 3 rows / columns on the top / left,
 * 8 channel slots per pixel, stored space-to-depth 2 x 4 (cell (i, j), element (a*4+q)*8+c =
 * padded pixel (2i+a, 4j+q), channel c; the same bytes per frame as a padded [H+6][W+8][8]
 * image - see DESIGN.md), for the fused device-resident pipeline.
 */
VPD_API int vpd_assemble_nchw(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      float* out_img, float* out_tgt, int B, int H, int W, int k, void* stream);
/* Host only (no GPU, no stream): K1's lookup tables, lut[c*256 + u] = the reference's fp32
 * value of byte u in channel c (c < 3: ((u/255) - mean[c]) / std[c] with three fp32 roundings,
 * vpd_dataset/common.py:52-58; c = 3, 4: float32(u/255.0 - 0.5), common.py:61-69), and the
 * constants of the stem-layout kernel's arithmetic path. Returns 1 when
 * bf16(fmaf(u, scale[c], shift[c])) == bf16(lut[c*256 + u]) for all 5 x 256 entries (the kernel
 * then computes instead of looking up), 0 when it keeps the tables. Value-returning, never fails. */
VPD_API int vpd_assemble_tables(const float* mean, const float* std, float* lut, float* scale,
                        float* shift);
VPD_API int vpd_assemble_stem(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      void* out_stem_bf16, float* out_tgt, int B, int H, int W, int k,
                      void* stream);
/* Training batch with the reference's masked-noise augmentation (single_frame.py:179-191:
 * with probability 0.5 per frame, Gaussian noise of sd sqrt(0.05) is added to the normalised
 * RGB planes at the pixels where `<n>.mask.png`'s first channel is NOT 0, before the flip).
 *   mask      uint8 [pool][H][W]   first channel of the mask PNG
 *   noise_on  uint8 [B]            the per-frame coin (host-drawn like `flip`), NULL = all
 *   noise     fp32 [B][3][H][W]    explicit noise in source orientation (bit-exact parity
 *                                  with a host generator), or NULL: counter-based Philox
 *                                  normals from `seed` scaled by `noise_sd` */
VPD_API int vpd_assemble_nchw_noise(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                            const int32_t* index, const uint8_t* flip, const float* teacher,
                            int teacher_rows, int tdim, const float* mean, const float* std,
                            float* out_img, float* out_tgt, int B, int H, int W,
                            const uint8_t* mask, const uint8_t* noise_on, const float* noise,
                            float noise_sd, uint64_t seed, void* stream);
VPD_API int vpd_assemble_stem_noise(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                            const int32_t* index, const uint8_t* flip, const float* teacher,
                            int teacher_rows, int tdim, const float* mean, const float* std,
                            void* out_stem_bf16, float* out_tgt, int B, int H, int W,
                            const uint8_t* mask, const uint8_t* noise_on, const float* noise,
                            float noise_sd, uint64_t seed, void* stream);
/* Training batch with the reference's full `augment=True` pipeline (single_frame.py:168-206):
 * u8/255 -> transforms.ColorJitter (vpd_dataset/common.py:11-12,88-92) -> Normalize -> masked
 * noise (:179-191) -> flow stacked -> horizontal flip (:199-203) -> transforms.RandomResizedCrop
 * (common.py:49-50,79-80; antialiased bilinear resize of the crop back to H x W). The DRAWS are
 * made by the caller in the reference's order (vpd_b200/augment.py); the pixel arithmetic
 * follows torchvision / ATen rounding for rounding.
 *   jitter_order  uint8 [B][4]  permutation of 0 brightness, 1 contrast, 2 saturation, 3 hue
 *                               (a value > 3 skips that slot); NULL = no ColorJitter
 *   jitter_factor fp32 [B][8]   {b, c, 1-c, s, 1-s, hue, 0, 0}, each rounded from the double
 *   crop          int32 [B][4]  (i, j, h, w) inside the frame; NULL = no crop
 *   mask / noise_on / noise / noise_sd / seed: as vpd_assemble_nchw_noise (mask NULL = none)
 * Output fp32 [B][C][H][W]. One CTA per frame; H*W*12 bytes of shared memory (<= 227 KB). */
VPD_API int vpd_assemble_nchw_aug(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                          const int32_t* index, const uint8_t* flip, const float* teacher,
                          int teacher_rows, int tdim, const float* mean, const float* std,
                          float* out_img, float* out_tgt, int B, int H, int W,
                          const uint8_t* jitter_order, const float* jitter_factor,
                          const int32_t* crop, const uint8_t* mask, const uint8_t* noise_on,
                          const float* noise, float noise_sd, uint64_t seed, void* stream);
/* fp32 NCHW batch (the reference's batch['img']) -> network input layout */
VPD_API int vpd_nchw_to_stem(const float* x, void* out_stem_bf16, int B, int C, int H, int W,
                     void* stream);

/* ---- K5: fused AdamW ------------------------------------------------------------
 * Replaces torch.optim.AdamW.step over all tensors (train_vpd_model.py:100-105,
 * models/util.py:50-58). Flat fp32 arenas of n elements; `step` is 1-based. */
VPD_API int vpd_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
              double lr, double beta1, double beta2, double eps, double weight_decay, int step,
              float grad_scale, void* stream);
/* fused SGD: torch.optim.SGD over the same flat arenas (the reference trains with AdamW; this
 * is the other optimiser BASELINE.json's north_star names). momentum_buf may be NULL when
 * momentum == 0; first_step != 0 initialises the buffer with the gradient like torch does. */
VPD_API int vpd_sgd(float* params, const float* grads, float* momentum_buf, int64_t n, double lr,
            double momentum, double dampening, double weight_decay, int nesterov, int first_step,
            float grad_scale, void* stream);

/* ---- K2: single convolution ops (NHWC bf16), used by the tests and the network ---
 * Replace ATen conv2d forward / backward as called by torchvision BasicBlock
 * (models/rgb.py:68-70). Weights: vpd_pack_conv_weight turns fp32 OIHW master
 * weights into bf16 [k*k][Cout][Cin] (forward) and [k*k][Cin][Cout] (dgrad). */
VPD_API int vpd_pack_conv_weight(const float* w_oihw, void* w_tap_bf16, void* wT_tap_bf16, int Cout,
                         int Cin, int k, void* stream);
/* stem (7x7/2, Cimg <= 8 input channels): fp32 [64][Cimg][7][7] -> the operand mirrors of the
 * space-to-depth stem kernel, 20 x 64 x 64 bf16 (8 + 12 tap tiles of the two output-column
 * classes; opaque) */
VPD_API int vpd_pack_stem_weight(const float* w_oihw, void* w_stem_bf16, int Cimg, void* stream);
/* y = conv(x); optional epilogue: y*scale[c]+shift[c], + residual, ReLU; optional
 * per-channel (sum, sumsq) accumulation into stats[2][Cout] (training BN). */
VPD_API int vpd_conv2d_fwd(const void* x, const void* w_tap, void* y, int N, int H, int W, int Cin,
                   int Cout, int k, int stride, int pad, const float* scale, const float* shift,
                   const void* residual, int relu, vpd_stat_acc* stats, void* stream);
/* 7x7/2 pad-3 stem on the padded input layout written by vpd_assemble_stem */
VPD_API int vpd_stem_conv_fwd(const void* x_stem, const void* w_stem, void* y, int N, int H, int W,
                      const float* scale, const float* shift, int relu, vpd_stat_acc* stats,
                      void* stream);
/* dx = conv_transpose(dy) (+ residual); for stride 2 the optional 1x1/2
 * downsample branch (dy_ds, wT_ds) is accumulated in the same pass. */
VPD_API int vpd_conv2d_dgrad(const void* dy, const void* wT_tap, void* dx, int N, int H, int W, int Cin,
                     int Cout, int k, int stride, int pad, const void* residual,
                     const void* dy_ds, const void* wT_ds, int cout_ds, void* stream);

/* Same as vpd_conv2d_dgrad (stride 1 or 2, no downsample branch) with the BatchNorm-
 * backward reduction of the consuming ReLU->BN stage folded into the epilogue: dx is
 * stored already masked, g = dx * 1[z > 0], and sums[0..Cin) += sum g,
 * sums[Cin..2Cin) += sum g * (y - mean) * rstd  (accumulated). relu_mask = 1[z > 0] as one
 * bit per element, uint8 [N][H][W][Cin/8], bit j of byte g = channel 8g + j (written by
 * vpd_bn_act_fwd, or by vpd_relu_bitmask from a tensor z): the kernel reads 1/16 of z's bytes. */
VPD_API int vpd_conv2d_dgrad_bnfused(const void* dy, const void* wT_tap, void* dx, int N, int H, int W,
                             int Cin, int Cout, int k, int stride, int pad, const void* residual,
                             const uint8_t* relu_mask, const void* y, const float* mean,
                             const float* rstd, vpd_stat_acc* sums, void* stream);
/* dw[k*k][Cout][Cin] (fp32, tap-major: the gradient arena's native conv layout)
 * += sum over pixels dy (x) x. ACCUMULATES into dw (zero it first). */
VPD_API int vpd_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin,
                     int Cout, int k, int stride, int pad, void* stream);
/* stem weight gradient, dw[7][64][64] fp32 in the packed stem layout (kh, cout, kw*8+c) */
VPD_API int vpd_stem_conv_wgrad(const void* x_stem, const void* dy, float* dw, int N, int H, int W,
                        void* stream);

/* ---- BatchNorm / ReLU / pooling / head ops (NHWC bf16 activations) ----------------
 * Replace torchvision BasicBlock's bn/relu/add, nn.MaxPool2d(3,2,1), the avgpool+fc
 * head, FCNet and F.mse_loss(sum) with their autograd (models/rgb.py:68-70,
 * models/module.py:133-156, train_vpd_model.py:87). Exposed for tests and reuse. */
/* train-mode BN (+ optional residual, itself optionally batch-normalised) + ReLU.
 * stats/res_stats: [2][C] per-channel (sum, sumsq) of y / res as produced by
 * vpd_conv2d_fwd; running buffers are updated (momentum .1), save_* receive batch
 * mean and 1/sqrt(var+eps). res_stats == NULL: residual added as is. relu_mask (may be
 * NULL): uint8 [M][C/8] out, bit j of byte (row, g) = 1[z[row][8g + j] > 0] - what
 * vpd_conv2d_dgrad_bnfused reads in the backward pass instead of z. */
VPD_API int vpd_bn_act_fwd(const void* y, const void* res, void* z, int64_t M, int C, int relu,
                   const vpd_stat_acc* stats, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches,
                   float* save_mean, float* save_rstd, const vpd_stat_acc* res_stats,
                   const float* res_gamma, const float* res_beta, float* res_running_mean,
                   float* res_running_var, int64_t* res_num_batches, float* res_save_mean,
                   float* res_save_rstd, uint8_t* relu_mask, void* stream);
/* relu_mask of an existing bf16 tensor z [M][C] (same bit layout) */
VPD_API int vpd_relu_bitmask(const void* z, uint8_t* mask, int64_t M, int C, void* stream);
/* gradient of the stage above: g = dz * 1[z > 0] (z may be NULL: no mask); dy = BN
 * backward of g through (y, save_mean, save_rstd, gamma); optional second branch
 * (y2...) fed by the same g; dmask (may alias dz) receives g; sums = [2][C]
 * scratch per branch, zeroed by the caller. */
VPD_API int vpd_bn_act_bwd(const void* dz, const void* z, void* dmask, int64_t M, int C,
                   const void* y, void* dy, const float* gamma, const float* save_mean,
                   const float* save_rstd, vpd_stat_acc* sums, float* dgamma, float* dbeta,
                   const void* y2, void* dy2, const float* gamma2, const float* save_mean2,
                   const float* save_rstd2, vpd_stat_acc* sums2, float* dgamma2, float* dbeta2,
                   void* stream);
/* stem: train-mode BN + ReLU + maxpool 3x3/2 pad 1 (argmax: uint8 window index; ysel, may be
 * NULL: bf16 [N][H/2][W/2][C], the pre-BN value at the argmax - handing it to the backward
 * call lets its reduction read two pooled tensors instead of gathering from y again) */
VPD_API int vpd_stem_bn_pool_fwd(const void* y, void* z, uint8_t* argmax, void* ysel, int N, int H, int W,
                         int C, const vpd_stat_acc* stats, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, int64_t* num_batches,
                         float* save_mean, float* save_rstd, void* stream);
VPD_API int vpd_stem_bn_pool_bwd(const void* dpool, const uint8_t* argmax, const void* y,
                         const void* ysel, void* dy, int N, int H, int W, int C, const float* gamma, const float* beta,
                         const float* save_mean, const float* save_rstd, vpd_stat_acc* sums,
                         float* dgamma, float* dbeta, void* stream);
/* K4: avgpool -> fc -> [FCNet] -> sum-squared-error loss and backward. params/grads:
 * fp32 arrays laid out [fc_w D*F][fc_b D][w0 Hd*D][b0 Hd][w2 Hd*Hd][b2 Hd][w5 T*Hd][b5 T]
 * (decoder part only when motion != 0, Hd = 128). target NULL: forward only.
 * ws: fp32 scratch of B*(F + 2D + 4*Hd + T). */
VPD_API int vpd_head_fwd_bwd(const void* z, int B, int HW, int F, int D, int T, int motion,
                     const float* params, const float* target, float* emb_out, float* out,
                     double* loss_sum, void* dz, float* ws, float* grads, void* stream);

/* ---- keypoint (VIPE*) teacher encoder: row-matrix helpers -----------------------------
 * The teacher whose embeddings the student regresses is an MLP (models/module.py:159-204
 * FcResidualBlock / FCResNet; eval forward models/keypoint.py:128-160 `_predict`,
 * apply_vipe_model.py:165-204). Its hidden x hidden Linear layers run on vpd_conv2d_fwd as
 * 1x1 convolutions over an [n][1][1][C] tensor (Linear bias + eval BatchNorm1d folded into
 * the epilogue's scale/shift by vpd_bn_fold); these cover the rest:
 *   vpd_rows_to_bf16    fp32 [M][C] -> bf16 [M][Cpad], zero padded (39 pose values -> 64)
 *   vpd_axpby_bf16      out = alpha*a + beta*b over n bf16 values (b may be NULL): `x2 - x`
 *   vpd_bn_fold         scale = gamma/sqrt(var+eps), shift = (bias-mean)*scale + beta
 *                       (bias may be NULL)
 *   vpd_linear_rows_f32 out fp32 [M][D] = x bf16 [M][K] . w fp32 [D][K]^T + bias; D <= 64,
 *                       K % 64 == 0 (the last Linear: the embedding is never rounded to bf16) */
VPD_API int vpd_rows_to_bf16(const float* x, void* out_bf16, int64_t M, int C, int Cpad, void* stream);
VPD_API int vpd_axpby_bf16(const void* a, float alpha, const void* b, float beta, void* out, int64_t n,
                   void* stream);
VPD_API int vpd_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                const float* running_var, const float* bias, float eps, float* scale,
                float* shift, int C, void* stream);
VPD_API int vpd_linear_rows_f32(const void* x_bf16, const float* w, const float* bias, float* out,
                        int64_t M, int K, int D, void* stream);

/* Training side of the same encoder (models/module.py:159-177 in train mode,
 * models/keypoint.py:38-126). Rows are bf16 [M][C], C % 8 == 0, C <= 2048.
 *   vpd_dropout_mask   keep[i] = 1 with probability 1 - p_drop (Philox4x32-10, keyed by seed
 *                      [+ *seed_add, a device-side step counter that lets a captured CUDA
 *                      graph draw fresh masks on every replay; may be NULL] and stream_id);
 *                      n % 4 == 0
 *   vpd_bn1d_fwd       out = keep * relu(BatchNorm1d_train(a)) / (1 - p_drop) [- res].
 *                      `a` is the Linear output WITHOUT its bias (a bias in front of a batch-
 *                      statistics BN cancels; lin_bias only enters running_mean); stats =
 *                      [2][C] column sums / sums of squares of `a` as vpd_conv2d_fwd
 *                      accumulates them; running stats (momentum .1, unbiased variance) and
 *                      num_batches are updated; save_mean / save_rstd for the backward
 *   vpd_bn1d_bwd       da = BN backward of g = dz * keep / (1-p) * 1[bn(a) > 0];
 *                      dgamma += sum g*xhat, dbeta += sum g (ACCUMULATED: the three encoder
 *                      passes of a step share weights); sums = [2][C] accumulator scratch
 *                      `groups`: the weight-sharing encoder passes of a step are stacked along
 *                      the rows - `groups` consecutive blocks of M rows, each its own BatchNorm
 *                      batch (stats / save_mean / save_rstd / sums are [groups][...]); running
 *                      statistics are updated group after group like consecutive calls
 *   vpd_colstats_bf16  stats[g] = column sums and sums of squares of row group g
 *   vpd_relu_mask_bf16 out = d * 1[z > 0]
 *   vpd_colsum_bf16    out[c] += sum_r x[r][c]   (Linear bias gradients)
 *   vpd_vipe_loss      the loss head: contra = ||e1-e2|| + valid * max(0, 1 - ||e1-en||)
 *                      (F.hinge_embedding_loss, targets +1 / -1), loss = contra + w3d *
 *                      (SSE(pred1, true3d) + SSE(pred2, true3d)); sums[0] += contra,
 *                      sums[1] += loss; gradients times gscale (1 / batch size): de* fp32
 *                      [n][D], dpred* bf16 [n][Tpad] (pad columns zero). e2 / en / true3d /
 *                      pred2 may be NULL (datasets without those entries). */
VPD_API int vpd_dropout_mask(uint8_t* keep, int64_t n, float p_drop, uint64_t seed,
                     const uint64_t* seed_add, int stream_id, void* stream);
VPD_API int vpd_bn1d_fwd(const void* a, const vpd_stat_acc* stats, const float* gamma, const float* beta,
                 const float* lin_bias, float* running_mean, float* running_var,
                 int64_t* num_batches, float* save_mean, float* save_rstd, const uint8_t* keep,
                 float p_drop, const void* res, void* out, int64_t M, int C, int groups,
                 void* stream);
VPD_API int vpd_bn1d_bwd(const void* dz, const void* a, const uint8_t* keep, float p_drop,
                 const float* gamma, const float* beta, const float* save_mean,
                 const float* save_rstd, vpd_stat_acc* sums, void* da, float* dgamma, float* dbeta,
                 int64_t M, int C, int groups, void* stream);
VPD_API int vpd_colstats_bf16(const void* x, vpd_stat_acc* stats, int64_t M, int C, int groups, void* stream);
VPD_API int vpd_relu_mask_bf16(const void* d, const void* z, void* out, int64_t n, void* stream);
VPD_API int vpd_colsum_bf16(const void* x, float* out, int64_t M, int C, void* stream);
VPD_API int vpd_vipe_loss(const float* e1, const float* e2, const float* en, const float* valid,
                  const void* pred1, const void* pred2, const float* true3d, float* de1,
                  float* de2, float* den, void* dpred1, void* dpred2, double* sums, int64_t n,
                  int D, int T, int Tpad, float w3d, float gscale, void* stream);

/* ---- the student network ---------------------------------------------------------
 * Replaces RGBF_EmbeddingModel.forward/embed (models/rgb.py:68-86), the body of
 * ModelTrainer.epoch (train_vpd_model.py:67-98: forward, FCNet decoder,
 * F.mse_loss(reduction='sum')) and loss.backward() (models/util.py:50-58).
 *
 * A vpd_net is a host-side execution plan (TMA descriptors, launch parameters);
 * it owns no device memory. The caller binds:
 *   params / grads   fp32 [vpd_net_param_count]   (grads may be NULL: inference only)
 *   buffers          fp32 [vpd_net_buffer_count]  BN running means then variances
 *   nbt              int64 [vpd_net_num_bn]       BN num_batches_tracked
 *   workspace        bytes [vpd_net_workspace_bytes]
 * vpd_net_tensor_info maps every reference state_dict entry ("resnet.*", plus
 * "decoder.layers.*" for the FCNet) to (arena, offset, layout):
 *   arena  0 params/grads, 1 buffers, 2 nbt
 *   layout 0 plain row-major in the reference shape
 *          1 conv weight stored tap-major [kh*kw][Cout][Cin]
 *          2 stem conv stored [7 kh][64 cout][kw*8 + c] (kw == 7 / c >= Cin zero)
 * Inputs: either x_nchw (fp32 [B][C][H][W], the reference's batch['img']) or
 * x_stem (the layout vpd_assemble_stem writes; may be vpd_net_stem_input itself).
 */
typedef struct vpd_net vpd_net;
VPD_API vpd_net* vpd_net_create(const char* arch, int emb_dim, int in_channels, int H, int W,
                        int max_batch, int motion);
VPD_API void vpd_net_destroy(vpd_net* net);
VPD_API int64_t vpd_net_param_count(vpd_net* net);
VPD_API int64_t vpd_net_conv_param_count(vpd_net* net);
VPD_API int64_t vpd_net_buffer_count(vpd_net* net);
VPD_API int vpd_net_num_bn(vpd_net* net);
VPD_API int64_t vpd_net_workspace_bytes(vpd_net* net);
VPD_API int vpd_net_num_tensors(vpd_net* net);
VPD_API int vpd_net_tensor_info(vpd_net* net, int i, char* name, int name_cap, int* arena,
                        int64_t* offset, int* layout, int* ndim, int64_t* shape4);
VPD_API int vpd_net_bind(vpd_net* net, float* params, float* grads, float* buffers, int64_t* nbt,
                 void* workspace, int64_t workspace_bytes);
/* vpd_adamw over the arenas bound to `net` (same arithmetic, bit-identical parameters and
 * moments) that also writes the bf16 tensor-core mirrors of the updated conv weights, so the
 * next vpd_net_train_step / vpd_net_forward starts without a weight-packing pass
 * (train_vpd_model.py:100-105, models/util.py:50-58). exp_avg / exp_avg_sq: fp32 arenas of
 * vpd_net_param_count elements. */
VPD_API int vpd_net_adamw(vpd_net* net, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int step, float grad_scale,
                  void* stream);
/* The same update for the arena range [offset, offset + count) only - one gradient bucket as
 * handed out by the vpd_net_set_bucket_callback hook - so the optimizer can run bucket by
 * bucket on another stream (after that bucket's all-reduce when data-parallel) while the
 * backward pass of the earlier layers is still running. The ranges of a step must tile the
 * arena; finish != 0 on the last one. */
VPD_API int vpd_net_adamw_range(vpd_net* net, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                        double beta2, double eps, double weight_decay, int step, float grad_scale,
                        int64_t offset, int64_t count, int finish, void* stream);
/* call after writing the parameter arena from outside (load_state_dict, optimizer) */
VPD_API int vpd_net_params_changed(vpd_net* net);
VPD_API void* vpd_net_stem_input(vpd_net* net);
/* Data-parallel hook: during vpd_net_train_step, `fn(user, offset, count)` is called on
 * the host each time a contiguous range of the gradient arena has been fully enqueued on
 * `stream` (last layers first; the ranges partition the arena). The caller records an
 * event and starts its all-reduce of that range on another stream, overlapping the rest
 * of the backward pass. fn == NULL disables the hook. */
typedef void (*vpd_bucket_fn)(void* user, int64_t offset, int64_t count);
VPD_API int vpd_net_set_bucket_callback(vpd_net* net, vpd_bucket_fn fn, void* user);
/* eval-mode encoder: emb_out fp32 [B][emb_dim] */
VPD_API int vpd_net_forward(vpd_net* net, const float* x_nchw, const void* x_stem, int B,
                    float* emb_out, void* stream);
/* training-mode encoder forward only (models/rgb.py:68-70 on a module in train() mode):
 * BatchNorm uses the batch statistics and updates its running buffers / counters;
 * emb_out fp32 [B][emb_dim]. No loss, no gradients (those come from vpd_net_train_step). */
VPD_API int vpd_net_forward_train(vpd_net* net, const float* x_nchw, const void* x_stem, int B,
                          float* emb_out, void* stream);
/* eval-mode forward + decoder + sum-squared-error; *loss_sum (device fp64) += loss;
 * out (optional) fp32 [B][target_dim] */
VPD_API int vpd_net_eval_loss(vpd_net* net, const float* x_nchw, const void* x_stem,
                      const float* target, int B, double* loss_sum, float* out, void* stream);
/* train-mode forward (batch-stat BN, running stats updated) + loss + backward;
 * gradients of every parameter are written to the grads arena */
VPD_API int vpd_net_train_step(vpd_net* net, const float* x_nchw, const void* x_stem,
                       const float* target, int B, double* loss_sum, void* stream);

/* Test/debug access to the bf16 NHWC activation buffers left by the last step:
 * block -1 = stem (which 0: conv1 output, 4: pooled), block >= 0 = BasicBlock index
 * (which 0 conv1 out, 1 post-bn1-relu, 2 conv2 out, 3 downsample conv out, 4 block out;
 * 5 / 6: the ReLU bit masks of 1 / 4 left by a TRAINING step - uint8, numel = bytes, one bit
 * per element as in vpd_bn_act_fwd) */
VPD_API int vpd_net_activation(vpd_net* net, int block, int which, int B, void** ptr,
                       int64_t* numel);
/* device-to-device copy on `stream` (lets tests snapshot the buffers above) */
VPD_API int vpd_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream);
/* Per-launch timing of vpd_net_train_step (CUDA events around every kernel),
 * summed into ms[8 kinds][8 stages] / counts[8][8]; kinds: 0 conv fwd, 1 BN/ReLU/pool
 * fwd, 2 conv dgrad, 3 conv wgrad, 4 BN/ReLU/pool bwd, 5 head+loss, 6 pack/convert,
 * 7 other; stages: 0 stem, 1..4 residual stages, 5 head/global. Synchronise the
 * stream before reading. Used by bench.py's roofline pass. */
VPD_API int vpd_net_profile_enable(vpd_net* net, int on);
VPD_API int vpd_net_profile_read(vpd_net* net, float* ms_host, int* counts_host);
/* Test-only hardware probe: D[128][64] = A x I where A is a K-major SWIZZLE_128B UMMA
 * operand whose descriptor starts at row `row_start` of a swizzled [rows][64] bf16 patch
 * in shared memory, 8-row groups `sbo_bytes` apart; base_offset_mode 1 sets the
 * descriptor's base-offset field to (start_addr >> 7) & 7. */
VPD_API int vpd_umma_probe(const void* src_bf16, int rows, int row_start, int sbo_bytes,
                   int base_offset_mode, float* out, void* stream);
/* Debug: while dev_i64 is non-null, every generic implicit-GEMM conv launch writes 8 int64
 * per CTA into it (globaltimer at entry, then SM clock at entry / dependencies resolved /
 * first operands landed / last MMA issued / first accumulator ready / epilogue done /
 * exit). The buffer must hold 8 * 148 entries. */
VPD_API int vpd_conv_trace(void* dev_i64);
/* number of kernel launches issued by this library since it was loaded */
VPD_API int64_t vpd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VPD_B200_H_ */
