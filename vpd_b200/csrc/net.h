#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpd {

struct Net;

Net* net_create(const char* arch, int emb_dim, int in_channels, int H, int W, int max_batch,
                int motion);
void net_destroy(Net* n);
long long net_workspace_bytes(Net* n);
int net_bind(Net* n, float* params, float* grads, float* buffers, long long* nbt, void* ws,
             long long ws_bytes);
int net_forward(Net* n, const float* x_nchw, const void* x_stem, int B, float* emb_out,
                cudaStream_t s);
int net_forward_train(Net* n, const float* x_nchw, const void* x_stem, int B, float* emb_out,
                      cudaStream_t s);
int net_eval_loss(Net* n, const float* x_nchw, const void* x_stem, const float* target, int B,
                  double* loss_sum, float* out, cudaStream_t s);
int net_train_step(Net* n, const float* x_nchw, const void* x_stem, const float* target, int B,
                   double* loss_sum, cudaStream_t s);

int net_adamw(Net* n, float* exp_avg, float* exp_avg_sq, double lr, double b1, double b2,
              double eps, double wd, int step, float grad_scale, cudaStream_t s);

int net_adamw_range(Net* n, float* exp_avg, float* exp_avg_sq, double lr, double b1, double b2,
                    double eps, double wd, int step, float grad_scale, long long offset,
                    long long count, int finish, cudaStream_t s);

// introspection (implemented in net.cu)
long long net_param_count(Net* n);
long long net_buffer_count(Net* n);
int net_num_bn(Net* n);
int net_num_tensors(Net* n);
int net_tensor_info(Net* n, int i, char* name, int name_cap, int* arena, long long* offset,
                    int* layout, int* ndim, long long* shape4);
void net_params_changed(Net* n);
void net_set_bucket_callback(Net* n, void (*fn)(void*, long long, long long), void* user);
void* net_stem_input(Net* n);
long long net_conv_section_len(Net* n);
int net_activation(Net* n, int block, int which, int B, void** ptr, long long* numel);
void net_profile_enable(Net* n, int on);
int net_profile_read(Net* n, float* ms, int* counts);

}  // namespace vpd
