// extern "C" surface declared in include/vpd_b200.h
#include "../../include/vpd_b200.h"

#include <string.h>

#include "conv.h"
#include "elementwise.cuh"
#include "head.h"
#include "net.h"
#include "ops.h"

using namespace vpd;
typedef __nv_bfloat16 bf16;

extern "C" {

const char* vpd_last_error(void) { return get_error(); }
int vpd_abi_version(void) { return 3; }   // 3: ReLU bit masks (vpd_bn_act_fwd, vpd_conv2d_dgrad_bnfused), vpd_relu_bitmask, vpd_assemble_tables

int vpd_assemble_nchw(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      float* out_img, float* out_tgt, int B, int H, int W, int k, void* stream) {
  return assemble_nchw(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, out_img, out_tgt, B, H, W, k, (cudaStream_t)stream);
}

int vpd_assemble_tables(const float* mean, const float* std, float* lut, float* scale, float* shift) {
  return assemble_tables(mean, std, lut, scale, shift);
}

int vpd_assemble_stem(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      void* out_stem_bf16, float* out_tgt, int B, int H, int W, int k,
                      void* stream) {
  return assemble_pad8(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, (bf16*)out_stem_bf16, out_tgt, B, H, W, k, (cudaStream_t)stream);
}

int vpd_assemble_nchw_noise(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                            const int32_t* index, const uint8_t* flip, const float* teacher,
                            int teacher_rows, int tdim, const float* mean, const float* std,
                            float* out_img, float* out_tgt, int B, int H, int W,
                            const uint8_t* mask, const uint8_t* noise_on, const float* noise,
                            float noise_sd, uint64_t seed, void* stream) {
  AsmNoise nz;
  nz.mask = mask;
  nz.noise_on = noise_on;
  nz.noise = noise;
  nz.noise_sd = noise_sd;
  nz.seed = seed;
  return assemble_nchw(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, out_img, out_tgt, B, H, W, 1, (cudaStream_t)stream, &nz);
}

int vpd_assemble_stem_noise(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                            const int32_t* index, const uint8_t* flip, const float* teacher,
                            int teacher_rows, int tdim, const float* mean, const float* std,
                            void* out_stem_bf16, float* out_tgt, int B, int H, int W,
                            const uint8_t* mask, const uint8_t* noise_on, const float* noise,
                            float noise_sd, uint64_t seed, void* stream) {
  AsmNoise nz;
  nz.mask = mask;
  nz.noise_on = noise_on;
  nz.noise = noise;
  nz.noise_sd = noise_sd;
  nz.seed = seed;
  return assemble_pad8(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, (bf16*)out_stem_bf16, out_tgt, B, H, W, 1, (cudaStream_t)stream, &nz);
}

int vpd_assemble_nchw_aug(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                          const int32_t* index, const uint8_t* flip, const float* teacher,
                          int teacher_rows, int tdim, const float* mean, const float* std,
                          float* out_img, float* out_tgt, int B, int H, int W,
                          const uint8_t* jitter_order, const float* jitter_factor,
                          const int32_t* crop, const uint8_t* mask, const uint8_t* noise_on,
                          const float* noise, float noise_sd, uint64_t seed, void* stream) {
  AsmNoise nz;
  nz.mask = mask;
  nz.noise_on = noise_on;
  nz.noise = noise;
  nz.noise_sd = noise_sd;
  nz.seed = seed;
  return assemble_aug(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                      std, out_img, out_tgt, B, H, W, jitter_order, jitter_factor, crop,
                      (cudaStream_t)stream, &nz);
}

int vpd_rows_to_bf16(const float* x, void* out_bf16, int64_t M, int C, int Cpad, void* stream) {
  return rows_to_bf16(x, (bf16*)out_bf16, M, C, Cpad, (cudaStream_t)stream);
}
int vpd_axpby_bf16(const void* a, float alpha, const void* b, float beta, void* out, int64_t n,
                   void* stream) {
  return axpby_bf16((const bf16*)a, alpha, (const bf16*)b, beta, (bf16*)out, n,
                    (cudaStream_t)stream);
}
int vpd_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                const float* running_var, const float* bias, float eps, float* scale,
                float* shift, int C, void* stream) {
  return bn_fold(gamma, beta, running_mean, running_var, bias, eps, scale, shift, C,
                 (cudaStream_t)stream);
}
int vpd_linear_rows_f32(const void* x_bf16, const float* w, const float* bias, float* out,
                        int64_t M, int K, int D, void* stream) {
  return linear_rows_f32((const bf16*)x_bf16, w, bias, out, M, K, D, (cudaStream_t)stream);
}

int vpd_dropout_mask(uint8_t* keep, int64_t n, float p_drop, uint64_t seed,
                     const uint64_t* seed_add, int stream_id, void* stream) {
  return dropout_mask(keep, n, p_drop, seed, (const unsigned long long*)seed_add,
                      (unsigned int)stream_id, (cudaStream_t)stream);
}
int vpd_bn1d_fwd(const void* a, const vpd_stat_acc* stats, const float* gamma, const float* beta,
                 const float* lin_bias, float* running_mean, float* running_var,
                 int64_t* num_batches, float* save_mean, float* save_rstd, const uint8_t* keep,
                 float p_drop, const void* res, void* out, int64_t M, int C, int groups,
                 void* stream) {
  return bn1d_fwd((const bf16*)a, (const StatAcc*)stats, gamma, beta, lin_bias, running_mean, running_var,
                  (long long*)num_batches, save_mean, save_rstd, keep, p_drop, (const bf16*)res,
                  (bf16*)out, M, C, groups, (cudaStream_t)stream);
}
int vpd_bn1d_bwd(const void* dz, const void* a, const uint8_t* keep, float p_drop,
                 const float* gamma, const float* beta, const float* save_mean,
                 const float* save_rstd, vpd_stat_acc* sums, void* da, float* dgamma, float* dbeta,
                 int64_t M, int C, int groups, void* stream) {
  return bn1d_bwd((const bf16*)dz, (const bf16*)a, keep, p_drop, gamma, beta, save_mean, save_rstd,
                  (StatAcc*)sums, (bf16*)da, dgamma, dbeta, M, C, groups, (cudaStream_t)stream);
}
int vpd_colstats_bf16(const void* x, vpd_stat_acc* stats, int64_t M, int C, int groups, void* stream) {
  return colstats_bf16((const bf16*)x, (StatAcc*)stats, M, C, groups, (cudaStream_t)stream);
}
int vpd_relu_mask_bf16(const void* d, const void* z, void* out, int64_t n, void* stream) {
  return relu_mask_bf16((const bf16*)d, (const bf16*)z, (bf16*)out, n, (cudaStream_t)stream);
}
int vpd_colsum_bf16(const void* x, float* out, int64_t M, int C, void* stream) {
  return colsum_bf16((const bf16*)x, out, M, C, (cudaStream_t)stream);
}
int vpd_vipe_loss(const float* e1, const float* e2, const float* en, const float* valid,
                  const void* pred1, const void* pred2, const float* true3d, float* de1,
                  float* de2, float* den, void* dpred1, void* dpred2, double* sums, int64_t n,
                  int D, int T, int Tpad, float w3d, float gscale, void* stream) {
  return vipe_loss(e1, e2, en, valid, (const bf16*)pred1, (const bf16*)pred2, true3d, de1, de2,
                   den, (bf16*)dpred1, (bf16*)dpred2, sums, n, D, T, Tpad, w3d, gscale,
                   (cudaStream_t)stream);
}

int vpd_nchw_to_stem(const float* x, void* out_stem_bf16, int B, int C, int H, int W,
                     void* stream) {
  return nchw_to_pad8(x, (bf16*)out_stem_bf16, B, C, H, W, (cudaStream_t)stream);
}

int vpd_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
              double lr, double beta1, double beta2, double eps, double weight_decay, int step,
              float grad_scale, void* stream) {
  return adamw_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                    step, grad_scale, (cudaStream_t)stream);
}

int vpd_sgd(float* params, const float* grads, float* momentum_buf, int64_t n, double lr,
            double momentum, double dampening, double weight_decay, int nesterov, int first_step,
            float grad_scale, void* stream) {
  return sgd_step(params, grads, momentum_buf, n, lr, momentum, dampening, weight_decay, nesterov,
                  first_step, grad_scale, (cudaStream_t)stream);
}

int vpd_pack_conv_weight(const float* w_oihw, void* w_tap_bf16, void* wT_tap_bf16, int Cout,
                         int Cin, int k, void* stream) {
  return pack_conv_weight(w_oihw, (bf16*)w_tap_bf16, (bf16*)wT_tap_bf16, Cout, Cin, k,
                          (cudaStream_t)stream);
}

int vpd_pack_stem_weight(const float* w_oihw, void* w_stem_bf16, int Cimg, void* stream) {
  return pack_stem_weight(w_oihw, (bf16*)w_stem_bf16, Cimg, (cudaStream_t)stream);
}

int vpd_conv2d_fwd(const void* x, const void* w_tap, void* y, int N, int H, int W, int Cin,
                   int Cout, int k, int stride, int pad, const float* scale, const float* shift,
                   const void* residual, int relu, vpd_stat_acc* stats, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  ConvEpilogue e;
  e.scale = scale;
  e.shift = shift;
  e.residual = (const bf16*)residual;
  e.relu = relu;
  e.stats = (StatAcc*)stats;
  ConvLaunch L;
  if (plan_conv_fwd(&L, g, (const bf16*)x, (const bf16*)w_tap, (bf16*)y, e)) return -1;
  return launch_conv(L, (cudaStream_t)stream);
}

int vpd_stem_conv_fwd(const void* x_stem, const void* w_stem, void* y, int N, int H, int W,
                      const float* scale, const float* shift, int relu, vpd_stat_acc* stats,
                      void* stream) {
  ConvEpilogue e;
  e.scale = scale;
  e.shift = shift;
  e.relu = relu;
  e.stats = (StatAcc*)stats;
  ConvLaunch L[2];
  if (plan_stem_fwd(L, N, H, W, (const bf16*)x_stem, (const bf16*)w_stem, (bf16*)y, e)) return -1;
  if (launch_conv(L[0], (cudaStream_t)stream)) return -1;
  return launch_conv(L[1], (cudaStream_t)stream);
}

int vpd_conv2d_dgrad(const void* dy, const void* wT_tap, void* dx, int N, int H, int W, int Cin,
                     int Cout, int k, int stride, int pad, const void* residual,
                     const void* dy_ds, const void* wT_ds, int cout_ds, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  ConvLaunch L[4];
  int count = 0;
  if (plan_conv_dgrad(L, &count, g, (const bf16*)dy, (const bf16*)wT_tap, (bf16*)dx,
                      (const bf16*)residual, (const bf16*)dy_ds, (const bf16*)wT_ds, cout_ds))
    return -1;
  for (int i = 0; i < count; ++i)
    if (launch_conv(L[i], (cudaStream_t)stream)) return -1;
  return 0;
}

int vpd_conv2d_dgrad_bnfused(const void* dy, const void* wT_tap, void* dx, int N, int H, int W,
                             int Cin, int Cout, int k, int stride, int pad, const void* residual,
                             const uint8_t* relu_mask, const void* y, const float* mean,
                             const float* rstd, vpd_stat_acc* sums, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  VPD_REQUIRE(relu_mask != nullptr && y != nullptr, "dgrad_bnfused: relu_mask and y are required");
  ConvBwdFuse f;
  f.nb = 1;
  f.mask = relu_mask;
  f.y[0] = (const bf16*)y;
  f.mean[0] = mean;
  f.rstd[0] = rstd;
  f.sums[0] = (StatAcc*)sums;
  ConvLaunch L[4];
  int count = 0;
  if (plan_conv_dgrad(L, &count, g, (const bf16*)dy, (const bf16*)wT_tap, (bf16*)dx,
                      (const bf16*)residual, nullptr, nullptr, 0, &f))
    return -1;
  for (int i = 0; i < count; ++i)
    if (launch_conv(L[i], (cudaStream_t)stream)) return -1;
  return 0;
}

int vpd_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin,
                     int Cout, int k, int stride, int pad, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  WgradLaunch L;
  if (plan_conv_wgrad(&L, g, (const bf16*)x, (const bf16*)dy, dw)) return -1;
  return launch_wgrad(L, (cudaStream_t)stream);
}

int vpd_stem_conv_wgrad(const void* x_stem, const void* dy, float* dw, int N, int H, int W,
                        void* stream) {
  WgradLaunch L[2];
  if (plan_stem_wgrad(L, N, H, W, (const bf16*)x_stem, (const bf16*)dy, dw)) return -1;
  if (launch_wgrad(L[0], (cudaStream_t)stream)) return -1;
  return launch_wgrad(L[1], (cudaStream_t)stream);
}

static BnLayer make_bn(const vpd_stat_acc* stats, const float* gamma, const float* beta, float* rm,
                       float* rv, int64_t* nbt, float* sm, float* sr, long long count) {
  BnLayer L;
  L.stats = (const StatAcc*)stats;
  L.gamma = gamma;
  L.beta = beta;
  L.running_mean = rm;
  L.running_var = rv;
  L.num_batches = (long long*)nbt;
  L.save_mean = sm;
  L.save_rstd = sr;
  L.count = (float)count;
  L.inv_count = 1.0 / (double)count;
  L.momentum = 0.1f;
  L.eps = 1e-5f;
  L.update_running = rm != nullptr ? 1 : 0;
  return L;
}

int vpd_bn_act_fwd(const void* y, const void* res, void* z, int64_t M, int C, int relu,
                   const vpd_stat_acc* stats, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* num_batches,
                   float* save_mean, float* save_rstd, const vpd_stat_acc* res_stats,
                   const float* res_gamma, const float* res_beta, float* res_running_mean,
                   float* res_running_var, int64_t* res_num_batches, float* res_save_mean,
                   float* res_save_rstd, uint8_t* relu_mask, void* stream) {
  BnApplyParams a;
  memset(&a, 0, sizeof(a));
  a.y = (const bf16*)y;
  a.res = (const bf16*)res;
  a.z = (bf16*)z;
  a.mask = relu_mask;
  a.M = M;
  a.C = C;
  a.relu = relu;
  a.bn = make_bn(stats, gamma, beta, running_mean, running_var, num_batches, save_mean, save_rstd, M);
  if (res_stats != nullptr) {
    a.has_res_bn = 1;
    a.res_bn = make_bn(res_stats, res_gamma, res_beta, res_running_mean, res_running_var,
                       res_num_batches, res_save_mean, res_save_rstd, M);
  }
  return launch_bn_apply(a, (cudaStream_t)stream);
}

int vpd_relu_bitmask(const void* z, uint8_t* mask, int64_t M, int C, void* stream) {
  return launch_relu_mask((const bf16*)z, mask, M, C, (cudaStream_t)stream);
}

int vpd_bn_act_bwd(const void* dz, const void* z, void* dmask, int64_t M, int C,
                   const void* y, void* dy, const float* gamma, const float* save_mean,
                   const float* save_rstd, vpd_stat_acc* sums, float* dgamma, float* dbeta,
                   const void* y2, void* dy2, const float* gamma2, const float* save_mean2,
                   const float* save_rstd2, vpd_stat_acc* sums2, float* dgamma2, float* dbeta2,
                   void* stream) {
  BnBwdParams q;
  memset(&q, 0, sizeof(q));
  q.dz = (const bf16*)dz;
  q.z = (const bf16*)z;
  q.dmask = (bf16*)dmask;
  q.M = M;
  q.C = C;
  q.nbranch = y2 != nullptr ? 2 : 1;
  q.y[0] = (const bf16*)y; q.dy[0] = (bf16*)dy; q.gamma[0] = gamma; q.save_mean[0] = save_mean;
  q.save_rstd[0] = save_rstd; q.sums[0] = (StatAcc*)sums; q.dgamma[0] = dgamma; q.dbeta[0] = dbeta;
  q.y[1] = (const bf16*)y2; q.dy[1] = (bf16*)dy2; q.gamma[1] = gamma2; q.save_mean[1] = save_mean2;
  q.save_rstd[1] = save_rstd2; q.sums[1] = (StatAcc*)sums2; q.dgamma[1] = dgamma2; q.dbeta[1] = dbeta2;
  return launch_bn_bwd(q, (cudaStream_t)stream);
}

int vpd_stem_bn_pool_fwd(const void* y, void* z, uint8_t* argmax, void* ysel, int N, int H, int W,
                         int C, const vpd_stat_acc* stats, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, int64_t* num_batches,
                         float* save_mean, float* save_rstd, void* stream) {
  PoolParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.y = (const bf16*)y;
  pp.z = (bf16*)z;
  pp.argmax = argmax;
  pp.ysel = (bf16*)ysel;
  pp.N = N; pp.H = H; pp.W = W; pp.C = C;
  pp.bn = make_bn(stats, gamma, beta, running_mean, running_var, num_batches, save_mean, save_rstd,
                  (long long)N * H * W);
  return launch_bn_pool(pp, (cudaStream_t)stream);
}

int vpd_stem_bn_pool_bwd(const void* dpool, const uint8_t* argmax, const void* y,
                         const void* ysel, void* dy, int N, int H, int W, int C, const float* gamma, const float* beta,
                         const float* save_mean, const float* save_rstd, vpd_stat_acc* sums,
                         float* dgamma, float* dbeta, void* stream) {
  StemBwdParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.dpool = (const bf16*)dpool;
  sp.argmax = argmax;
  sp.y = (const bf16*)y;
  sp.ysel = (const bf16*)ysel;
  sp.dy = (bf16*)dy;
  sp.N = N; sp.H = H; sp.W = W; sp.C = C;
  sp.gamma = gamma; sp.beta = beta; sp.save_mean = save_mean; sp.save_rstd = save_rstd;
  sp.sums = (StatAcc*)sums; sp.dgamma = dgamma; sp.dbeta = dbeta;
  return launch_stem_bwd(sp, (cudaStream_t)stream);
}

int vpd_head_fwd_bwd(const void* z, int B, int HW, int F, int D, int T, int motion,
                     const float* params, const float* target, float* emb_out, float* out,
                     double* loss_sum, void* dz, float* ws, float* grads, void* stream) {
  const int Hd = 128;
  HeadParams h;
  memset(&h, 0, sizeof(h));
  h.z = (const bf16*)z;
  h.B = B; h.HW = HW; h.F = F; h.D = D; h.T = T; h.Hd = Hd; h.motion = motion;
  const float* p = params;
  float* g = grads;
  HeadGrads hg;
  memset(&hg, 0, sizeof(hg));
  h.fc_w = p; p += (size_t)D * F; h.fc_b = p; p += D;
  if (g) { hg.fc_w = g; g += (size_t)D * F; hg.fc_b = g; g += D; }
  if (motion) {
    h.w0 = p; p += Hd * D; h.b0 = p; p += Hd; h.w2 = p; p += Hd * Hd; h.b2 = p; p += Hd;
    h.w5 = p; p += T * Hd; h.b5 = p;
    if (g) {
      hg.w0 = g; g += Hd * D; hg.b0 = g; g += Hd; hg.w2 = g; g += Hd * Hd; hg.b2 = g; g += Hd;
      hg.w5 = g; g += T * Hd; hg.b5 = g;
    }
  }
  h.target = target;
  h.emb_out = emb_out;
  h.out = out;
  h.loss = loss_sum;
  h.dz = (bf16*)dz;
  h.ws = ws;
  return launch_head(h, grads ? &hg : nullptr, (cudaStream_t)stream);
}

vpd_net* vpd_net_create(const char* arch, int emb_dim, int in_channels, int H, int W,
                        int max_batch, int motion) {
  return (vpd_net*)net_create(arch, emb_dim, in_channels, H, W, max_batch, motion);
}
void vpd_net_destroy(vpd_net* net) { net_destroy((Net*)net); }
int64_t vpd_net_param_count(vpd_net* net) { return net_param_count((Net*)net); }
int64_t vpd_net_conv_param_count(vpd_net* net) { return net_conv_section_len((Net*)net); }
int64_t vpd_net_buffer_count(vpd_net* net) { return net_buffer_count((Net*)net); }
int vpd_net_num_bn(vpd_net* net) { return net_num_bn((Net*)net); }
int64_t vpd_net_workspace_bytes(vpd_net* net) { return net_workspace_bytes((Net*)net); }
int vpd_net_num_tensors(vpd_net* net) { return net_num_tensors((Net*)net); }
int vpd_net_tensor_info(vpd_net* net, int i, char* name, int name_cap, int* arena,
                        int64_t* offset, int* layout, int* ndim, int64_t* shape4) {
  return net_tensor_info((Net*)net, i, name, name_cap, arena, (long long*)offset, layout, ndim,
                         (long long*)shape4);
}
int vpd_net_bind(vpd_net* net, float* params, float* grads, float* buffers, int64_t* nbt,
                 void* workspace, int64_t workspace_bytes) {
  return net_bind((Net*)net, params, grads, buffers, (long long*)nbt, workspace, workspace_bytes);
}
int vpd_net_params_changed(vpd_net* net) {
  net_params_changed((Net*)net);
  return 0;
}
int vpd_net_set_bucket_callback(vpd_net* net, vpd_bucket_fn fn, void* user) {
  net_set_bucket_callback((Net*)net, (void (*)(void*, long long, long long))fn, user);
  return 0;
}
void* vpd_net_stem_input(vpd_net* net) { return net_stem_input((Net*)net); }
int vpd_net_forward(vpd_net* net, const float* x_nchw, const void* x_stem, int B,
                    float* emb_out, void* stream) {
  return net_forward((Net*)net, x_nchw, x_stem, B, emb_out, (cudaStream_t)stream);
}
int vpd_net_forward_train(vpd_net* net, const float* x_nchw, const void* x_stem, int B,
                          float* emb_out, void* stream) {
  return net_forward_train((Net*)net, x_nchw, x_stem, B, emb_out, (cudaStream_t)stream);
}
int vpd_net_eval_loss(vpd_net* net, const float* x_nchw, const void* x_stem,
                      const float* target, int B, double* loss_sum, float* out, void* stream) {
  return net_eval_loss((Net*)net, x_nchw, x_stem, target, B, loss_sum, out, (cudaStream_t)stream);
}
int vpd_net_train_step(vpd_net* net, const float* x_nchw, const void* x_stem,
                       const float* target, int B, double* loss_sum, void* stream) {
  return net_train_step((Net*)net, x_nchw, x_stem, target, B, loss_sum, (cudaStream_t)stream);
}

int vpd_net_adamw(vpd_net* net, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int step, float grad_scale,
                  void* stream) {
  return net_adamw((Net*)net, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step,
                   grad_scale, (cudaStream_t)stream);
}

int vpd_net_adamw_range(vpd_net* net, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                        double beta2, double eps, double weight_decay, int step, float grad_scale,
                        int64_t offset, int64_t count, int finish, void* stream) {
  return net_adamw_range((Net*)net, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step,
                         grad_scale, offset, count, finish, (cudaStream_t)stream);
}

int vpd_net_activation(vpd_net* net, int block, int which, int B, void** ptr, int64_t* numel) {
  return net_activation((Net*)net, block, which, B, ptr, (long long*)numel);
}
int vpd_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream) {
  VPD_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
  return 0;
}
int vpd_net_profile_enable(vpd_net* net, int on) {
  net_profile_enable((Net*)net, on);
  return 0;
}
int vpd_net_profile_read(vpd_net* net, float* ms_host, int* counts_host) {
  return net_profile_read((Net*)net, ms_host, counts_host);
}
int64_t vpd_launch_count(void) { return launch_count(); }
int vpd_conv_trace(void* dev_i64) {
  set_conv_trace((long long*)dev_i64);
  return 0;
}
int vpd_umma_probe(const void* src_bf16, int rows, int row_start, int sbo_bytes,
                   int base_offset_mode, float* out, void* stream) {
  return umma_probe((const bf16*)src_bf16, rows, row_start, sbo_bytes, base_offset_mode, out,
                    (cudaStream_t)stream);
}

}  // extern "C"
