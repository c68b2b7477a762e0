// BatchNorm / ReLU / residual / pooling kernels around the tensor-core convs
// (torchvision BasicBlock.forward and its autograd).
// All activations are NHWC bf16 viewed as [M = N*H*W rows][C channels]; every
// thread owns a fixed group of 8 channels (one 16-byte vector) and walks rows,
// so loads/stores are 128-bit and coalesced and per-channel reductions need no
// cross-thread traffic until the final block-level combine. HBM-bound.
#pragma once
#include "common.cuh"

namespace vpd {

// Per-BatchNorm-layer pointers. Training: `stats` holds the fp64 sum / sumsq
// produced by the conv epilogue; evaluation: stats == nullptr and the running
// buffers are used.
struct BnLayer {
  const double* stats;     // [2][C] or null (eval)
  const float* gamma;      // [C]
  const float* beta;       // [C]
  float* running_mean;     // [C]
  float* running_var;      // [C]
  long long* num_batches;  // scalar
  float* save_mean;        // [C] out (train) - batch mean
  float* save_rstd;        // [C] out (train) - 1/sqrt(var+eps)
  float count;             // N*H*W
  double inv_count;        // 1 / count
  float momentum, eps;
  int update_running;      // train: block 0 updates running stats
};

struct BnApplyParams {
  const __nv_bfloat16* y;    // [M][C] conv output (pre-BN)
  const __nv_bfloat16* res;  // [M][C] residual input or null
  __nv_bfloat16* z;          // [M][C] out
  long long M;
  int C;
  int relu;
  int has_res_bn;            // residual goes through its own BN (downsample branch)
  BnLayer bn, res_bn;
};

struct PoolParams {          // stem: BN + ReLU + maxpool 3x3/2 pad 1
  const __nv_bfloat16* y;    // [N][H][W][C]
  __nv_bfloat16* z;          // [N][H/2][W/2][C]
  uint8_t* argmax;           // [N][H/2][W/2][C] window index 0..8, or null
  int N, H, W, C;
  BnLayer bn;
};

struct BnBwdParams {
  // upstream gradient g = dz * 1[z > 0] (mask skipped when z == null)
  const __nv_bfloat16* dz;   // [M][C]
  const __nv_bfloat16* z;    // [M][C] post-ReLU output of this stage, or null
  __nv_bfloat16* dmask;      // [M][C] out: masked gradient (identity branch), or null
  long long M;
  int C;
  int nbranch;               // 1 or 2 BN branches fed by the same g
  int sums_ready;            // 1: sums were accumulated by the producer (fused dgrad epilogue)
  const __nv_bfloat16* y[2]; // pre-BN conv outputs
  __nv_bfloat16* dy[2];      // out: gradient wrt conv outputs
  const float* gamma[2];
  const float* save_mean[2];
  const float* save_rstd[2];
  double* sums[2];           // [2][C] scratch: sum g, sum g*xhat (zeroed by caller)
  float* dgamma[2];          // [C] out
  float* dbeta[2];           // [C] out
};

struct StemBwdParams {       // maxpool + ReLU + BN backward of the stem
  const __nv_bfloat16* dpool;  // [N][H/2][W/2][C]
  const uint8_t* argmax;       // [N][H/2][W/2][C]
  const __nv_bfloat16* y;      // [N][H][W][C]
  __nv_bfloat16* dy;           // [N][H][W][C]
  int N, H, W, C;
  const float* gamma;
  const float* save_mean;
  const float* save_rstd;
  const float* beta;
  double* sums;                // [2][C]
  float* dgamma;
  float* dbeta;
};

int launch_bn_apply(const BnApplyParams& p, cudaStream_t s);
int launch_channel_stats(const __nv_bfloat16* y, long long M, int C, double* stats, cudaStream_t s);
int launch_bn_pool(const PoolParams& p, cudaStream_t s);
int launch_bn_bwd(const BnBwdParams& p, cudaStream_t s);    // reduce + apply
int launch_stem_bwd(const StemBwdParams& p, cudaStream_t s);
int launch_maxpool(const __nv_bfloat16* x, __nv_bfloat16* z, int N, int H, int W, int C,
                   cudaStream_t s);
// scale/shift for eval-mode folding: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale
int launch_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                   float eps, float* scale, float* shift, int C, cudaStream_t s);

}  // namespace vpd
