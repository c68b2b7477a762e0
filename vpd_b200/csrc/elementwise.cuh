// BatchNorm / ReLU / residual / pooling kernels around the tensor-core convs
// (torchvision BasicBlock.forward and its autograd).
// All activations are NHWC bf16 viewed as [M = N*H*W rows][C channels]; every
// thread owns a fixed group of 8 channels (one 16-byte vector) and walks rows,
// so loads/stores are 128-bit and coalesced and per-channel reductions need no
// cross-thread traffic until the final block-level combine. HBM-bound.
#pragma once
#include "common.cuh"

namespace vpd {

// Per-BatchNorm-layer pointers. Training: `stats` holds the sum / sumsq produced by
// the conv epilogue (order-independent integer accumulators, see StatAcc);
// evaluation: stats == nullptr and the running buffers are used.
struct BnLayer {
  const StatAcc* stats;    // [2][C] or null (eval)
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

// Batch (train) or running (eval) statistics -> fp32 mean / rstd of channel c.
// Only fp64 multiplies/FMAs (full rate) - the mean-square subtraction is the one
// place that needs the extra bits; the reciprocal square root is fp32.
// kCoherent: the sums were accumulated earlier in THIS kernel by other CTAs (conv kernels
// with the BatchNorm apply fused behind a grid barrier): read them through L2, not the
// non-coherent path.
template <bool kCoherent = false>
VPD_DEVINL void bn_mean_rstd(const BnLayer& bn, int c, int C, float& mean, float& rstd,
                             float& var_biased) {
  if (bn.stats != nullptr) {
    const double inv = bn.inv_count;
    const double s0 = kCoherent ? stat_read_cg(bn.stats + c) : stat_read(bn.stats + c);
    const double s1 = kCoherent ? stat_read_cg(bn.stats + C + c) : stat_read(bn.stats + C + c);
    const double m = s0 * inv;
    double v = fma(s1, inv, -m * m);
    if (v < 0.0) v = 0.0;
    mean = static_cast<float>(m);
    var_biased = static_cast<float>(v);
    rstd = rsqrtf(var_biased + bn.eps);
  } else {
    mean = bn.running_mean[c];
    var_biased = bn.running_var[c];
    rstd = rsqrtf(var_biased + bn.eps);
  }
}
// The affine every kernel (forward and backward) derives from (mean, rstd).
VPD_DEVINL void bn_affine(float gamma, float beta, float mean, float rstd, float& scale,
                          float& shift) {
  scale = gamma * rstd;
  shift = beta - mean * scale;
}
// Persist the batch statistics of channel c and update the running buffers like
// nn.BatchNorm2d (momentum 0.1, unbiased variance for the running estimate).
VPD_DEVINL void bn_channel_side_effects(const BnLayer& bn, int c, float mean, float rstd,
                                        float var) {
  if (bn.save_mean) bn.save_mean[c] = mean;
  if (bn.save_rstd) bn.save_rstd[c] = rstd;
  if (bn.update_running) {
    const float unbias = bn.count > 1.f ? bn.count / (bn.count - 1.f) : 1.f;
    bn.running_mean[c] = (1.f - bn.momentum) * bn.running_mean[c] + bn.momentum * mean;
    bn.running_var[c] = (1.f - bn.momentum) * bn.running_var[c] + bn.momentum * var * unbias;
  }
}

struct BnApplyParams {
  const __nv_bfloat16* y;    // [M][C] conv output (pre-BN)
  const __nv_bfloat16* res;  // [M][C] residual input or null
  __nv_bfloat16* z;          // [M][C] out
  uint8_t* mask;             // [M][C/8] out or null: bit j of byte (row, g) = 1[z[row][8g + j] > 0],
                             // what the fused BN-backward reduction of the data-gradient kernels
                             // reads instead of z (1/16 of its bytes)
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
  __nv_bfloat16* ysel;       // [N][H/2][W/2][C] out: the PRE-BN value at the argmax, or null
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
  StatAcc* sums[2];          // [2][C] scratch: sum g, sum g*xhat (zeroed by caller)
  float* dgamma[2];          // [C] out
  float* dbeta[2];           // [C] out
};

struct StemBwdParams {       // maxpool + ReLU + BN backward of the stem
  const __nv_bfloat16* dpool;  // [N][H/2][W/2][C]
  const uint8_t* argmax;       // [N][H/2][W/2][C]
  const __nv_bfloat16* y;      // [N][H][W][C]
  const __nv_bfloat16* ysel;   // [N][H/2][W/2][C] from the forward pool kernel, or null: the
                               // reduction then reads two pooled tensors instead of gathering
                               // every window from y again
  __nv_bfloat16* dy;           // [N][H][W][C]
  int N, H, W, C;
  const float* gamma;
  const float* save_mean;
  const float* save_rstd;
  const float* beta;
  StatAcc* sums;               // [2][C]
  float* dgamma;
  float* dbeta;
};

int launch_bn_apply(const BnApplyParams& p, cudaStream_t s);
// mask[row][g] bit j = 1[z[row][8g + j] > 0] for a bf16 tensor produced elsewhere
int launch_relu_mask(const __nv_bfloat16* z, uint8_t* mask, long long M, int C, cudaStream_t s);
int launch_channel_stats(const __nv_bfloat16* y, long long M, int C, StatAcc* stats, cudaStream_t s);
int launch_bn_pool(const PoolParams& p, cudaStream_t s);
int launch_bn_bwd(const BnBwdParams& p, cudaStream_t s);    // reduce + apply
int launch_stem_bwd(const StemBwdParams& p, cudaStream_t s);
int launch_maxpool(const __nv_bfloat16* x, __nv_bfloat16* z, int N, int H, int W, int C,
                   cudaStream_t s);
// scale/shift for eval-mode folding: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale
int launch_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                   float eps, float* scale, float* shift, int C, cudaStream_t s);

}  // namespace vpd
