#include "elementwise.cuh"

#include "tma_host.h"

namespace vpd {

constexpr int kEwThreads = 256;

VPD_DEVINL void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
VPD_DEVINL uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// Batch (train) or running (eval) statistics -> fp32 mean / rstd of channel c.
VPD_DEVINL void bn_mean_rstd(const BnLayer& bn, int c, int C, float& mean, float& rstd,
                             float& var_biased) {
  if (bn.stats != nullptr) {
    const double inv = 1.0 / static_cast<double>(bn.count);
    const double m = bn.stats[c] * inv;
    double v = bn.stats[C + c] * inv - m * m;
    if (v < 0.0) v = 0.0;
    mean = static_cast<float>(m);
    var_biased = static_cast<float>(v);
    rstd = static_cast<float>(1.0 / sqrt(v + static_cast<double>(bn.eps)));
  } else {
    mean = bn.running_mean[c];
    var_biased = bn.running_var[c];
    rstd = 1.0f / sqrtf(var_biased + bn.eps);
  }
}
// The affine every kernel (forward and backward) derives from (mean, rstd).
VPD_DEVINL void bn_affine(float gamma, float beta, float mean, float rstd, float& scale,
                          float& shift) {
  scale = gamma * rstd;
  shift = beta - mean * scale;
}

// Block 0: persist batch statistics and update the running buffers like
// nn.BatchNorm2d (momentum 0.1, unbiased variance for the running estimate).
VPD_DEVINL void bn_side_effects(const BnLayer& bn, int C) {
  if (bn.stats == nullptr) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd, var;
    bn_mean_rstd(bn, c, C, mean, rstd, var);
    if (bn.save_mean) bn.save_mean[c] = mean;
    if (bn.save_rstd) bn.save_rstd[c] = rstd;
    if (bn.update_running) {
      const float unbias = bn.count > 1.f ? bn.count / (bn.count - 1.f) : 1.f;
      bn.running_mean[c] = (1.f - bn.momentum) * bn.running_mean[c] + bn.momentum * mean;
      bn.running_var[c] = (1.f - bn.momentum) * bn.running_var[c] + bn.momentum * var * unbias;
    }
  }
  if (bn.update_running && threadIdx.x == 0 && bn.num_batches) *bn.num_batches += 1;
}

// ------------------------------------------------------------------ BN apply
__global__ void __launch_bounds__(kEwThreads) bn_apply_kernel(const BnApplyParams p) {
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int r0 = threadIdx.x / groups;
  const int rstep = kEwThreads / groups;
  float sc[8], sh[8], rsc[8], rsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    float mean, rstd, var;
    bn_mean_rstd(p.bn, c, p.C, mean, rstd, var);
    bn_affine(p.bn.gamma[c], p.bn.beta[c], mean, rstd, sc[j], sh[j]);
    rsc[j] = 1.f;
    rsh[j] = 0.f;
    if (p.has_res_bn) {
      bn_mean_rstd(p.res_bn, c, p.C, mean, rstd, var);
      bn_affine(p.res_bn.gamma[c], p.res_bn.beta[c], mean, rstd, rsc[j], rsh[j]);
    }
  }
  const long long chunk = (p.M + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * chunk;
  const long long end = beg + chunk < p.M ? beg + chunk : p.M;
  for (long long row = beg + r0; row < end; row += rstep) {
    const size_t off = (size_t)row * p.C + g * 8;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.y + off)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    if (p.res != nullptr) {
      float r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.res + off)), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += fmaf(r[j], rsc[j], rsh[j]);
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    stg_v4(p.z + off, pack8(f));
  }
  // every block has consumed the statistics above before block 0 may touch
  // the running buffers it also reads in eval mode; in train mode the inputs
  // (stats) are never modified here, so no ordering issue.
  if (blockIdx.x == 0) {
    bn_side_effects(p.bn, p.C);
    if (p.has_res_bn) bn_side_effects(p.res_bn, p.C);
  }
}

static int ew_grid(long long vectors) {
  long long blocks = (vectors + kEwThreads * 8 - 1) / (kEwThreads * 8);
  const long long cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int launch_bn_apply(const BnApplyParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 2048 && kEwThreads % (p.C / 8) == 0,
              "bn_apply: unsupported channel count %d", p.C);
  if (p.M == 0) return 0;
  bn_apply_kernel<<<ew_grid(p.M * (p.C / 8)), kEwThreads, 0, s>>>(p);
  VPD_LAUNCHED(1);
  return 0;
}

// ---------------------------------------------- stem: BN + ReLU + maxpool 3x3/2
__global__ void __launch_bounds__(kEwThreads) bn_pool_kernel(const PoolParams p) {
  const int groups = p.C >> 3;
  const int Ho = p.H / 2, Wo = p.W / 2;
  const long long total = (long long)p.N * Ho * Wo * groups;
  // channel group is fixed per thread when the grid stride is a multiple of `groups`
  const long long stride = (long long)gridDim.x * kEwThreads;
  const int g = threadIdx.x % groups;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    float mean, rstd, var;
    bn_mean_rstd(p.bn, c, p.C, mean, rstd, var);
    bn_affine(p.bn.gamma[c], p.bn.beta[c], mean, rstd, sc[j], sh[j]);
  }
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += stride) {
    long long pix = i / groups;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float best[8];
    int idx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      idx[j] = 0;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
      if (h < 0 || h >= p.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo - 1 + kw;
        if (w < 0 || w >= p.W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(
                    p.y + (((size_t)n * p.H + h) * p.W + w) * p.C + g * 8)),
                f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
          if (v > best[j]) {
            best[j] = v;
            idx[j] = kh * 3 + kw;
          }
        }
      }
    }
    const size_t o = (((size_t)n * Ho + ho) * Wo + wo) * p.C + g * 8;
    stg_v4(p.z + o, pack8(best));
    if (p.argmax != nullptr) {
      uint2 a;
      a.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
      a.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
      *reinterpret_cast<uint2*>(p.argmax + o) = a;
    }
  }
  if (blockIdx.x == 0) bn_side_effects(p.bn, p.C);
}

int launch_bn_pool(const PoolParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && kEwThreads % (p.C / 8) == 0, "bn_pool: unsupported C=%d", p.C);
  VPD_REQUIRE(p.H % 2 == 0 && p.W % 2 == 0, "bn_pool: odd spatial dims");
  if (p.N == 0) return 0;
  const long long total = (long long)p.N * (p.H / 2) * (p.W / 2) * (p.C / 8);
  long long blocks = (total + kEwThreads - 1) / kEwThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_pool_kernel<<<(int)blocks, kEwThreads, 0, s>>>(p);
  VPD_LAUNCHED(1);
  return 0;
}

// plain maxpool 3x3/2 pad 1 (eval path: input already activated)
__global__ void __launch_bounds__(kEwThreads)
maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ z, int N, int H,
               int W, int C) {
  const int groups = C >> 3;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kEwThreads) {
    const int g = (int)(i % groups);
    long long pix = i / groups;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
      if (h < 0 || h >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo - 1 + kw;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + h) * W + w) * C + g * 8)),
                f);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], f[j]);
      }
    }
    stg_v4(z + (((size_t)n * Ho + ho) * Wo + wo) * C + g * 8, pack8(best));
  }
}

int launch_maxpool(const __nv_bfloat16* x, __nv_bfloat16* z, int N, int H, int W, int C,
                   cudaStream_t s) {
  VPD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: unsupported shape");
  if (N == 0) return 0;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  long long blocks = (total + kEwThreads - 1) / kEwThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxpool_kernel<<<(int)blocks, kEwThreads, 0, s>>>(x, z, N, H, W, C);
  VPD_LAUNCHED(1);
  return 0;
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm,
                               const float* rv, float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(rv[c] + eps);
  bn_affine(gamma[c], beta[c], rm[c], rstd, scale[c], shift[c]);
}

int launch_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                   float eps, float* scale, float* shift, int C, cudaStream_t s) {
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(gamma, beta, rm, rv, eps, scale, shift, C);
  VPD_LAUNCHED(1);
  return 0;
}

// --------------------------------------------------------------- BN backward
// pass 1: per-channel sum(g), sum(g * xhat_b);  pass 2: dy_b, dgamma, dbeta.
template <bool kApply>
__global__ void __launch_bounds__(kEwThreads) bn_bwd_kernel(const BnBwdParams p) {
  __shared__ float s_g[512];
  __shared__ float s_gx[2][512];
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int r0 = threadIdx.x / groups;
  const int rstep = kEwThreads / groups;
  float mean[2][8], rstd[2][8], k0[2][8], k1[2][8], k2[2][8];
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    if (b >= p.nbranch) break;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      mean[b][j] = p.save_mean[b][c];
      rstd[b][j] = p.save_rstd[b][c];
      if (kApply) {
        const float invM = 1.0f / static_cast<float>(p.M);
        const float sg = static_cast<float>(p.sums[b][c]);
        const float sgx = static_cast<float>(p.sums[b][p.C + c]);
        k0[b][j] = p.gamma[b][c] * rstd[b][j];  // dy = k0 * (g - k1 - xhat * k2)
        k1[b][j] = sg * invM;
        k2[b][j] = sgx * invM;
      }
    }
  }
  float acc_g[8], acc_gx[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_g[j] = acc_gx[0][j] = acc_gx[1][j] = 0.f;

  const long long chunk = (p.M + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * chunk;
  const long long end = beg + chunk < p.M ? beg + chunk : p.M;
  for (long long row = beg + r0; row < end; row += rstep) {
    const size_t off = (size_t)row * p.C + g * 8;
    float gr[8];
    unpack8(*reinterpret_cast<const uint4*>(p.dz + off), gr);  // may alias dmask/dy
    if (p.z != nullptr) {
      float zz[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.z + off)), zz);
#pragma unroll
      for (int j = 0; j < 8; ++j) gr[j] = zz[j] > 0.f ? gr[j] : 0.f;
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (b >= p.nbranch) break;
      float yy[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.y[b] + off)), yy);
      if (kApply) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (yy[j] - mean[b][j]) * rstd[b][j];
          o[j] = k0[b][j] * (gr[j] - k1[b][j] - xh * k2[b][j]);
        }
        stg_v4(p.dy[b] + off, pack8(o));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (yy[j] - mean[b][j]) * rstd[b][j];
          acc_gx[b][j] += gr[j] * xh;
        }
      }
    }
    if (kApply) {
      if (p.dmask != nullptr) stg_v4(p.dmask + off, pack8(gr));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_g[j] += gr[j];
    }
  }
  if (kApply) {
    if (blockIdx.x == 0) {
      for (int c = threadIdx.x; c < p.C; c += kEwThreads)
        for (int b = 0; b < p.nbranch; ++b) {
          p.dbeta[b][c] = static_cast<float>(p.sums[b][c]);
          p.dgamma[b][c] = static_cast<float>(p.sums[b][p.C + c]);
        }
    }
  } else {
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) s_g[c] = s_gx[0][c] = s_gx[1][c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_g[g * 8 + j], acc_g[j]);
      atomicAdd(&s_gx[0][g * 8 + j], acc_gx[0][j]);
      if (p.nbranch > 1) atomicAdd(&s_gx[1][g * 8 + j], acc_gx[1][j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += kEwThreads)
      for (int b = 0; b < p.nbranch; ++b) {
        atomicAdd(&p.sums[b][c], static_cast<double>(s_g[c]));
        atomicAdd(&p.sums[b][p.C + c], static_cast<double>(s_gx[b][c]));
      }
  }
}

int launch_bn_bwd(const BnBwdParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 512 && kEwThreads % (p.C / 8) == 0,
              "bn_bwd: unsupported channel count %d", p.C);
  VPD_REQUIRE(p.nbranch == 1 || p.nbranch == 2, "bn_bwd: nbranch");
  if (p.M == 0) return 0;
  const int grid = ew_grid(p.M * (p.C / 8));
  bn_bwd_kernel<false><<<grid, kEwThreads, 0, s>>>(p);
  bn_bwd_kernel<true><<<grid, kEwThreads, 0, s>>>(p);
  VPD_LAUNCHED(2);
  return 0;
}

// ------------------------------------------- stem backward (pool + ReLU + BN)
template <bool kApply>
__global__ void __launch_bounds__(kEwThreads) stem_bwd_kernel(const StemBwdParams p) {
  __shared__ float s_g[512];
  __shared__ float s_gx[512];
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int Ho = p.H / 2, Wo = p.W / 2;
  const long long M = (long long)p.N * p.H * p.W;
  float mean[8], rstd[8], sc[8], sh[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    mean[j] = p.save_mean[c];
    rstd[j] = p.save_rstd[c];
    bn_affine(p.gamma[c], p.beta[c], mean[j], rstd[j], sc[j], sh[j]);
    if (kApply) {
      const float invM = 1.0f / static_cast<float>(M);
      k0[j] = p.gamma[c] * rstd[j];
      k1[j] = static_cast<float>(p.sums[c]) * invM;
      k2[j] = static_cast<float>(p.sums[p.C + c]) * invM;
    }
  }
  float acc_g[8], acc_gx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_g[j] = acc_gx[j] = 0.f;
  const long long total = M * groups;
  const long long stride = (long long)gridDim.x * kEwThreads;  // multiple of groups
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += stride) {
    long long pix = i / groups;
    const int w = (int)(pix % p.W);
    pix /= p.W;
    const int h = (int)(pix % p.H);
    const int n = (int)(pix / p.H);
    const size_t off = (((size_t)n * p.H + h) * p.W + w) * p.C + g * 8;
    float yy[8], gr[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.y + off)), yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) gr[j] = 0.f;
    // pooled windows (i, j) that contain (h, w): rows 2i-1 .. 2i+1
    const int i_lo = h >> 1, i_hi = (h + 1) >> 1;  // i_lo == i_hi when h is even
    const int j_lo = w >> 1, j_hi = (w + 1) >> 1;
    for (int pi = i_lo; pi <= i_hi; ++pi) {
      if (pi >= Ho) continue;
      for (int pj = j_lo; pj <= j_hi; ++pj) {
        if (pj >= Wo) continue;
        const int kidx = (h - (2 * pi - 1)) * 3 + (w - (2 * pj - 1));
        const size_t po = (((size_t)n * Ho + pi) * Wo + pj) * p.C + g * 8;
        const uint2 am = __ldg(reinterpret_cast<const uint2*>(p.argmax + po));
        float dp[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.dpool + po)), dp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t word = j < 4 ? am.x : am.y;
          const int a = (word >> (8 * (j & 3))) & 0xFF;
          if (a == kidx) gr[j] += dp[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) gr[j] = fmaf(yy[j], sc[j], sh[j]) > 0.f ? gr[j] : 0.f;
    if (kApply) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (yy[j] - mean[j]) * rstd[j];
        o[j] = k0[j] * (gr[j] - k1[j] - xh * k2[j]);
      }
      stg_v4(p.dy + off, pack8(o));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (yy[j] - mean[j]) * rstd[j];
        acc_g[j] += gr[j];
        acc_gx[j] += gr[j] * xh;
      }
    }
  }
  if (kApply) {
    if (blockIdx.x == 0)
      for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
        p.dbeta[c] = static_cast<float>(p.sums[c]);
        p.dgamma[c] = static_cast<float>(p.sums[p.C + c]);
      }
  } else {
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) s_g[c] = s_gx[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_g[g * 8 + j], acc_g[j]);
      atomicAdd(&s_gx[g * 8 + j], acc_gx[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
      atomicAdd(&p.sums[c], static_cast<double>(s_g[c]));
      atomicAdd(&p.sums[p.C + c], static_cast<double>(s_gx[c]));
    }
  }
}

int launch_stem_bwd(const StemBwdParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 512 && kEwThreads % (p.C / 8) == 0, "stem_bwd: C=%d", p.C);
  if (p.N == 0) return 0;
  const long long total = (long long)p.N * p.H * p.W * (p.C / 8);
  long long blocks = (total + kEwThreads * 4 - 1) / (kEwThreads * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  stem_bwd_kernel<false><<<(int)blocks, kEwThreads, 0, s>>>(p);
  stem_bwd_kernel<true><<<(int)blocks, kEwThreads, 0, s>>>(p);
  VPD_LAUNCHED(2);
  return 0;
}

}  // namespace vpd
