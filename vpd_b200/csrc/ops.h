// Internal (C++) declarations of the op launchers implemented across csrc/*.cu
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace vpd {

// optional masked-noise augmentation of the assembled training batch (see assemble.cu)
struct AsmNoise {
  const uint8_t* mask = nullptr;      // [pool][H][W] first channel of <n>.mask.png
  const uint8_t* noise_on = nullptr;  // [B] per-frame coin, or null (all frames)
  const float* noise = nullptr;       // [B][3][H][W] explicit noise, or null (device Philox)
  float noise_sd = 0.f;
  unsigned long long seed = 0;
};

int assemble_nchw(const uint8_t* rgb, const uint8_t* flow, int flow_channels, const int* index,
                  const uint8_t* flip, const float* teacher, int teacher_rows, int tdim,
                  const float* mean, const float* stdv, float* out_img, float* out_tgt, int B,
                  int H, int W, int k, cudaStream_t stream, const AsmNoise* nz = nullptr);
// host only: K1's 5 x 256 lookup tables and the verified constants of its arithmetic path
// (returns 1 when fmaf(u, sc[c], sh[c]) reproduces every table entry in bf16)
int assemble_tables(const float* mean, const float* stdv, float* lut_out, float* sc_out,
                    float* sh_out);
int assemble_pad8(const uint8_t* rgb, const uint8_t* flow, int flow_channels, const int* index,
                  const uint8_t* flip, const float* teacher, int teacher_rows, int tdim,
                  const float* mean, const float* stdv, __nv_bfloat16* out_pad, float* out_tgt,
                  int B, int H, int W, int k, cudaStream_t stream, const AsmNoise* nz = nullptr);
// training batch with ColorJitter + RandomResizedCrop (+ noise, flip) applied, fp32 NCHW
int assemble_aug(const uint8_t* rgb, const uint8_t* flow, int flow_channels, const int* index,
                 const uint8_t* flip, const float* teacher, int teacher_rows, int tdim,
                 const float* mean, const float* stdv, float* out_img, float* out_tgt, int B,
                 int H, int W, const uint8_t* jitter_order, const float* jitter_factor,
                 const int* crop, cudaStream_t stream, const AsmNoise* nz = nullptr);
int nchw_to_pad8(const float* x, __nv_bfloat16* out, int B, int C, int H, int W,
                 cudaStream_t stream);
int adamw_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
               double b2, double eps, double wd, int step, float grad_scale,
               cudaStream_t stream);

// AdamW whose conv-weight section also refreshes the bf16 operand mirrors (adamw.cu);
// `table` = one (matrix offset, rows, cols, 32-row tile, 64-column tile) entry per CTA
int adamw_step_mirrored(float* p, const float* g, float* m, float* v, long long n,
                        long long conv_off, long long conv_len, __nv_bfloat16* w_tap,
                        __nv_bfloat16* wT, const int* table, int tiles, double lr, double b1,
                        double b2, double eps, double wd, int step, float grad_scale,
                        cudaStream_t stream);

int adamw_tiles(float* p_conv, const float* g_conv, float* m_conv, float* v_conv,
                __nv_bfloat16* w_tap, __nv_bfloat16* wT, const int* table, int tiles, double lr,
                double b1, double b2, double eps, double wd, int step, float grad_scale,
                cudaStream_t stream);

int sgd_step(float* p, const float* g, float* buf, long long n, double lr, double momentum,
             double dampening, double wd, int nesterov, int first_step, float grad_scale,
             cudaStream_t stream);

// row-matrix helpers of the keypoint (VIPE*) encoder, mlp.cu
int rows_to_bf16(const float* x, __nv_bfloat16* out, long long M, int C, int Cpad,
                 cudaStream_t stream);
int axpby_bf16(const __nv_bfloat16* a, float alpha, const __nv_bfloat16* b, float beta,
               __nv_bfloat16* out, long long n, cudaStream_t stream);
int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
            const float* bias, float eps, float* scale, float* shift, int C, cudaStream_t stream);
int linear_rows_f32(const __nv_bfloat16* x, const float* w, const float* bias, float* out,
                    long long M, int K, int D, cudaStream_t stream);

int dropout_mask(uint8_t* keep, long long n, float p_drop, unsigned long long seed,
                 const unsigned long long* seed_add, unsigned int stream_id, cudaStream_t stream);
int bn1d_fwd(const __nv_bfloat16* a, const StatAcc* stats, const float* gamma, const float* beta,
             const float* lin_bias, float* running_mean, float* running_var, long long* num_batches,
             float* save_mean, float* save_rstd, const uint8_t* keep, float p_drop,
             const __nv_bfloat16* res, __nv_bfloat16* out, long long M, int C, int groups,
             cudaStream_t stream);
int bn1d_bwd(const __nv_bfloat16* dz, const __nv_bfloat16* a, const uint8_t* keep, float p_drop,
             const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
             StatAcc* sums, __nv_bfloat16* da, float* dgamma, float* dbeta, long long M, int C,
             int groups, cudaStream_t stream);
int colstats_bf16(const __nv_bfloat16* x, StatAcc* stats, long long M, int C, int groups,
                  cudaStream_t stream);
int relu_mask_bf16(const __nv_bfloat16* d, const __nv_bfloat16* z, __nv_bfloat16* out, long long n,
                   cudaStream_t stream);
int colsum_bf16(const __nv_bfloat16* x, float* out, long long M, int C, cudaStream_t stream);
int vipe_loss(const float* e1, const float* e2, const float* en, const float* valid,
              const __nv_bfloat16* pred1, const __nv_bfloat16* pred2, const float* true3d,
              float* de1, float* de2, float* den, __nv_bfloat16* dpred1, __nv_bfloat16* dpred2,
              double* sums, long long n, int D, int T, int Tpad, float w3d, float gscale,
              cudaStream_t stream);

int umma_probe(const __nv_bfloat16* src, int rows, int row_start, int sbo_bytes,
               int base_offset_mode, float* out, cudaStream_t stream);

}  // namespace vpd
