// extern "C" surface declared in include/vpd_b200.h
#include "../../include/vpd_b200.h"

#include "conv.h"
#include "net.h"
#include "ops.h"

using namespace vpd;
typedef __nv_bfloat16 bf16;

extern "C" {

const char* vpd_last_error(void) { return get_error(); }
int vpd_abi_version(void) { return 1; }

int vpd_assemble_nchw(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      float* out_img, float* out_tgt, int B, int H, int W, int k, void* stream) {
  return assemble_nchw(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, out_img, out_tgt, B, H, W, k, (cudaStream_t)stream);
}

int vpd_assemble_stem(const uint8_t* rgb, const uint8_t* flow, int flow_channels,
                      const int32_t* index, const uint8_t* flip, const float* teacher,
                      int teacher_rows, int tdim, const float* mean, const float* std,
                      void* out_stem_bf16, float* out_tgt, int B, int H, int W, int k,
                      void* stream) {
  return assemble_pad8(rgb, flow, flow_channels, index, flip, teacher, teacher_rows, tdim, mean,
                       std, (bf16*)out_stem_bf16, out_tgt, B, H, W, k, (cudaStream_t)stream);
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
                   const void* residual, int relu, double* stats, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  ConvEpilogue e;
  e.scale = scale;
  e.shift = shift;
  e.residual = (const bf16*)residual;
  e.relu = relu;
  e.stats = stats;
  ConvLaunch L;
  if (plan_conv_fwd(&L, g, (const bf16*)x, (const bf16*)w_tap, (bf16*)y, e)) return -1;
  return launch_conv(L, (cudaStream_t)stream);
}

int vpd_stem_conv_fwd(const void* x_stem, const void* w_stem, void* y, int N, int H, int W,
                      const float* scale, const float* shift, int relu, double* stats,
                      void* stream) {
  ConvEpilogue e;
  e.scale = scale;
  e.shift = shift;
  e.relu = relu;
  e.stats = stats;
  ConvLaunch L;
  if (plan_stem_fwd(&L, N, H, W, (const bf16*)x_stem, (const bf16*)w_stem, (bf16*)y, e)) return -1;
  return launch_conv(L, (cudaStream_t)stream);
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
                             const void* z, const void* y, const float* mean, const float* rstd,
                             double* sums, void* stream) {
  ConvGeom g{N, H, W, Cin, Cout, k, stride, pad};
  ConvBwdFuse f;
  f.nb = 1;
  f.z = (const bf16*)z;
  f.y[0] = (const bf16*)y;
  f.mean[0] = mean;
  f.rstd[0] = rstd;
  f.sums[0] = sums;
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
  WgradLaunch L;
  if (plan_stem_wgrad(&L, N, H, W, (const bf16*)x_stem, (const bf16*)dy, dw)) return -1;
  return launch_wgrad(L, (cudaStream_t)stream);
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
void* vpd_net_stem_input(vpd_net* net) { return net_stem_input((Net*)net); }
int vpd_net_forward(vpd_net* net, const float* x_nchw, const void* x_stem, int B,
                    float* emb_out, void* stream) {
  return net_forward((Net*)net, x_nchw, x_stem, B, emb_out, (cudaStream_t)stream);
}
int vpd_net_eval_loss(vpd_net* net, const float* x_nchw, const void* x_stem,
                      const float* target, int B, double* loss_sum, float* out, void* stream) {
  return net_eval_loss((Net*)net, x_nchw, x_stem, target, B, loss_sum, out, (cudaStream_t)stream);
}
int vpd_net_train_step(vpd_net* net, const float* x_nchw, const void* x_stem,
                       const float* target, int B, double* loss_sum, void* stream) {
  return net_train_step((Net*)net, x_nchw, x_stem, target, B, loss_sum, (cudaStream_t)stream);
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
int vpd_umma_probe(const void* src_bf16, int rows, int row_start, int sbo_bytes,
                   int base_offset_mode, float* out, void* stream) {
  return umma_probe((const bf16*)src_bf16, rows, row_start, sbo_bytes, base_offset_mode, out,
                    (cudaStream_t)stream);
}

}  // extern "C"
