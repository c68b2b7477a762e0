// Row-matrix helpers for the keypoint (VIPE*) encoder, an MLP whose wide layers run on the
// implicit-GEMM kernels as 1x1 convolutions over an [n][1][1][C] "image" (SURVEY §8f 1;
// reference models/module.py:159-204 FcResidualBlock / FCResNet, models/keypoint.py:128-160
// Keypoint_EmbeddingModel._predict). What the conv kernels cannot express lives here:
//   rows_to_bf16    fp32 [M][C] -> bf16 [M][Cpad], zero padded (39 pose values -> 64 channels)
//   axpby_bf16      out = alpha * a + beta * b on bf16 rows (the block's `x2 - x`)
//   bn_fold         eval-mode BatchNorm1d + the preceding Linear bias -> per-column scale/shift
//                   for the conv epilogue: y = (a + bias - mean) * gamma / sqrt(var + eps) + beta
//   linear_rows_f32 the last Linear (hidden -> emb_dim <= 64): fp32 weights, fp32 accumulation,
//                   fp32 output - the embedding itself is never rounded to bf16
// All HBM-bound / tiny; 16-byte vector accesses.
#include "common.cuh"
#include "ops.h"
#include "tma_host.h"

namespace vpd {

__global__ void __launch_bounds__(256)
rows_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long M,
                    int C, int Cpad) {
  pdl_trigger();
  pdl_wait();
  const int groups = Cpad >> 3;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? __ldg(x + r * C + c0 + j) : 0.f;
    stg_v4(out + r * Cpad + c0, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                           pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
  }
}

__global__ void __launch_bounds__(256)
axpby_bf16_kernel(const __nv_bfloat16* __restrict__ a, float alpha,
                  const __nv_bfloat16* __restrict__ b, float beta, __nv_bfloat16* out,
                  long long n8) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 va = ldg_nc_v4(a + i * 8);
    uint4 vb = make_uint4(0, 0, 0, 0);
    if (b != nullptr) vb = ldg_nc_v4(b + i * 8);
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(__fadd_rn(__fmul_rn(alpha, bf16_lo(wa[j])), __fmul_rn(beta, bf16_lo(wb[j]))),
                         __fadd_rn(__fmul_rn(alpha, bf16_hi(wa[j])), __fmul_rn(beta, bf16_hi(wb[j]))));
    stg_v4(out + i * 8, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

__global__ void __launch_bounds__(256)
bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ mean, const float* __restrict__ var,
               const float* __restrict__ bias, float eps, float* __restrict__ scale,
               float* __restrict__ shift, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] / sqrtf(var[c] + eps);
  const float b = bias != nullptr ? bias[c] : 0.f;
  scale[c] = s;
  shift[c] = (b - mean[c]) * s + beta[c];
}

// out[M][D] = x[M][K] (bf16) . w[D][K]^T (fp32) + bias[D]; D <= 64, K % 64 == 0.
// CTA = 32 rows x 256 threads; thread (r = tid / 8, q = tid % 8) owns outputs d = q, q+8, ...
constexpr int kLinRows = 32;
constexpr int kLinK = 64;
constexpr int kLinStride = kLinK + 4;   // rows stay 16-byte aligned; float4 reads conflict-free
__global__ void __launch_bounds__(256)
linear_rows_f32_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, float* __restrict__ out, long long M,
                       int K, int D) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_w[64][kLinStride];
  __shared__ __align__(16) float s_x[kLinRows][kLinStride];
  const int tid = threadIdx.x;
  const int r = tid >> 3, q = tid & 7;
  const long long row0 = (long long)blockIdx.x * kLinRows;
  const int nj = (D + 7) >> 3;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kLinK) {
    for (int i = tid; i < 64 * (kLinK / 4); i += 256) {
      const int d = i / (kLinK / 4), k4 = i - d * (kLinK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < D) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)d * K + k0) + k4);
      *reinterpret_cast<float4*>(&s_w[d][k4 * 4]) = v;
    }
    for (int i = tid; i < kLinRows * (kLinK / 8); i += 256) {
      const int rr = i / (kLinK / 8), g = i - rr * (kLinK / 8);
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row0 + rr < M) {
        const uint4 u = ldg_nc_v4(x + (size_t)(row0 + rr) * K + k0 + g * 8);
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
      }
      *reinterpret_cast<float4*>(&s_x[rr][g * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s_x[rr][g * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kLinK; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(&s_x[r][k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nj) {   // uniform: rows >= D of s_w are zero padding
          const float4 wv = *reinterpret_cast<const float4*>(&s_w[q + 8 * j][k]);
          acc[j] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[j]))));
        }
      }
    }
    __syncthreads();
  }
  if (row0 + r < M) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = q + 8 * j;
      if (d < D) out[(size_t)(row0 + r) * D + d] = acc[j] + (bias != nullptr ? bias[d] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Training-side pieces of the keypoint encoder (models/module.py:159-177 in train mode,
// models/keypoint.py:38-126): BatchNorm1d with batch statistics + ReLU + Dropout (+ `x2 - x`)
// forward and backward over bf16 rows [M][C], the ReLU mask, column sums (Linear bias
// gradients) and the hinge + MSE loss head. C % 8 == 0; a thread owns 8 adjacent columns
// (one 16-byte vector) of a set of rows.

// keep[i] = 1 with probability 1 - p_drop (counter-based: Philox4x32-10 keyed by the seed,
// one call per 4 elements)
__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ keep, long long n4, float p_drop,
                    unsigned long long seed, const unsigned long long* __restrict__ seed_add,
                    unsigned int stream_id) {
  pdl_trigger();
  pdl_wait();
  if (seed_add != nullptr) seed += *seed_add;   // device-side step counter (CUDA-graph replays)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned int c0 = (unsigned int)i, c1 = (unsigned int)(i >> 32), c2 = stream_id, c3 = 0x5eedu;
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ k1;
      c3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const unsigned int w[4] = {c0, c1, c2, c3};
    unsigned int out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = (static_cast<float>(w[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      out |= (u >= p_drop ? 1u : 0u) << (8 * j);
    }
    reinterpret_cast<unsigned int*>(keep)[i] = out;
  }
}

struct Bn1dParams {
  const __nv_bfloat16* a;      // [M][C] pre-BN Linear output WITHOUT its bias
  const StatAcc* stats;        // [2][C] sum, sum of squares of `a` over the M rows
  const float* gamma;
  const float* beta;
  const float* lin_bias;       // the Linear's bias: only shifts the batch mean (running_mean)
  float* running_mean;
  float* running_var;
  long long* num_batches;
  float* save_mean;            // [C] batch mean of `a` (bias excluded)
  float* save_rstd;            // [C]
  const uint8_t* keep;         // [M][C] dropout keep mask or null
  float keep_scale;            // 1 / (1 - p)
  const __nv_bfloat16* res;    // [M][C] subtracted from the result (`x2 - x`) or null
  __nv_bfloat16* out;          // [M][C]
  long long M;
  int C;
  float eps, momentum;
};

// gridDim.y = number of row groups (the weight-sharing encoder passes of one step are stacked
// along the rows; each group of p.M rows is its own BatchNorm batch): group g owns rows
// [g*M, (g+1)*M), stats[g], save_mean[g], save_rstd[g]. Running statistics are updated group
// after group, as the reference's consecutive encoder calls do.
__global__ void __launch_bounds__(256)
bn1d_fwd_kernel(const Bn1dParams p) {
  pdl_trigger();
  pdl_wait();
  const int groups = p.C >> 3;                    // host: groups <= 256
  const int rstep = 256 / groups;                 // rows per CTA pass; threads beyond are idle
  const int gg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int grp = blockIdx.y;
  const StatAcc* stats = p.stats + (size_t)grp * 2 * p.C;
  const size_t base = (size_t)grp * p.M * p.C;
  if (rl < rstep) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = gg * 8 + j;
      const double mean = stat_read(stats + c) / static_cast<double>(p.M);
      double var = stat_read(stats + p.C + c) / static_cast<double>(p.M) - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
      sc[j] = p.gamma[c] * rstd;
      sh[j] = p.beta[c] - static_cast<float>(mean) * sc[j];
      if (blockIdx.x == 0 && rl == 0) {
        p.save_mean[(size_t)grp * p.C + c] = static_cast<float>(mean);
        p.save_rstd[(size_t)grp * p.C + c] = rstd;
        if (grp == 0) {
          const float bias = p.lin_bias ? p.lin_bias[c] : 0.f;
          float rm = p.running_mean[c], rv = p.running_var[c];
          for (int q = 0; q < (int)gridDim.y; ++q) {
            const double m = stat_read(p.stats + (size_t)q * 2 * p.C + c) / static_cast<double>(p.M);
            double v = stat_read(p.stats + (size_t)q * 2 * p.C + p.C + c) / static_cast<double>(p.M) - m * m;
            if (v < 0.0) v = 0.0;
            const double unbiased = p.M > 1 ? v * static_cast<double>(p.M) / static_cast<double>(p.M - 1) : v;
            rm = (1.f - p.momentum) * rm + p.momentum * (static_cast<float>(m) + bias);
            rv = (1.f - p.momentum) * rv + p.momentum * static_cast<float>(unbiased);
          }
          p.running_mean[c] = rm;
          p.running_var[c] = rv;
        }
      }
    }
    const long long r0 = (long long)blockIdx.x * rstep + rl;
    for (long long r = r0; r < p.M; r += (long long)gridDim.x * rstep) {
      const size_t off = base + (size_t)r * p.C + gg * 8;
      const uint4 va = ldg_nc_v4(p.a + off);
      float v[8] = {bf16_lo(va.x), bf16_hi(va.x), bf16_lo(va.y), bf16_hi(va.y),
                    bf16_lo(va.z), bf16_hi(va.z), bf16_lo(va.w), bf16_hi(va.w)};
      unsigned long long kp = 0x0101010101010101ull;
      if (p.keep) kp = *reinterpret_cast<const unsigned long long*>(p.keep + off);
      float rr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.res) {
        const uint4 vr = ldg_nc_v4(p.res + off);
        rr[0] = bf16_lo(vr.x); rr[1] = bf16_hi(vr.x); rr[2] = bf16_lo(vr.y); rr[3] = bf16_hi(vr.y);
        rr[4] = bf16_lo(vr.z); rr[5] = bf16_hi(vr.z); rr[6] = bf16_lo(vr.w); rr[7] = bf16_hi(vr.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float u = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
        u = ((kp >> (8 * j)) & 0xff) ? u * p.keep_scale : 0.f;
        v[j] = u - rr[j];
      }
      stg_v4(p.out + off, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                     pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && p.num_batches)
    *p.num_batches += gridDim.y;
}

struct Bn1dBwdParams {
  const __nv_bfloat16* dz;     // [M][C] gradient of the block output path (before `- x`)
  const __nv_bfloat16* a;      // [M][C] saved pre-BN values
  const uint8_t* keep;         // or null
  float keep_scale;
  const float* gamma;
  const float* beta;
  const float* save_mean;
  const float* save_rstd;
  StatAcc* sums;               // [2][C] scratch: sum g, sum g * xhat (zeroed by the caller)
  __nv_bfloat16* da;           // [M][C]
  float* dgamma;               // += (weights are shared by the three encoder passes)
  float* dbeta;                // +=
  long long M;
  int C;
};

// g = dz * keep * keep_scale * 1[gamma * xhat + beta > 0]; gridDim.y = row groups (see forward)
template <bool kApply>
__global__ void __launch_bounds__(256)
bn1d_bwd_kernel(const Bn1dBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int groups = p.C >> 3;
  const int rstep = 256 / groups;
  const int gg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int grp = blockIdx.y;
  StatAcc* sums = p.sums + (size_t)grp * 2 * p.C;
  const size_t base = (size_t)grp * p.M * p.C;
  if (rl < rstep) {
    float mean[8], rstd[8], ga[8], be[8], k1[8], k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = gg * 8 + j;
      mean[j] = p.save_mean[(size_t)grp * p.C + c];
      rstd[j] = p.save_rstd[(size_t)grp * p.C + c];
      ga[j] = p.gamma[c];
      be[j] = p.beta[c];
      k1[j] = kApply ? static_cast<float>(stat_read(sums + c) / static_cast<double>(p.M)) : 0.f;
      k2[j] = kApply ? static_cast<float>(stat_read(sums + p.C + c) / static_cast<double>(p.M)) : 0.f;
    }
    float sg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sgx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long r0 = (long long)blockIdx.x * rstep + rl;
    for (long long r = r0; r < p.M; r += (long long)gridDim.x * rstep) {
      const size_t off = base + (size_t)r * p.C + gg * 8;
      const uint4 vd = ldg_nc_v4(p.dz + off), va = ldg_nc_v4(p.a + off);
      const float d[8] = {bf16_lo(vd.x), bf16_hi(vd.x), bf16_lo(vd.y), bf16_hi(vd.y),
                          bf16_lo(vd.z), bf16_hi(vd.z), bf16_lo(vd.w), bf16_hi(vd.w)};
      const float av[8] = {bf16_lo(va.x), bf16_hi(va.x), bf16_lo(va.y), bf16_hi(va.y),
                           bf16_lo(va.z), bf16_hi(va.z), bf16_lo(va.w), bf16_hi(va.w)};
      unsigned long long kp = 0x0101010101010101ull;
      if (p.keep) kp = *reinterpret_cast<const unsigned long long*>(p.keep + off);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sc = ga[j] * rstd[j];
        const float pre = fmaf(av[j], sc, be[j] - mean[j] * sc);   // same expression as forward
        const float xh = (av[j] - mean[j]) * rstd[j];
        float gr = (pre > 0.f && ((kp >> (8 * j)) & 0xff)) ? d[j] * p.keep_scale : 0.f;
        if (kApply) {
          o[j] = sc * (gr - k1[j] - xh * k2[j]);
        } else {
          sg[j] += gr;
          sgx[j] = fmaf(gr, xh, sgx[j]);
        }
      }
      if (kApply)
        stg_v4(p.da + off, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                      pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
    }
    if (!kApply) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        stat_add(&sums[gg * 8 + j], static_cast<double>(sg[j]));
        stat_add(&sums[p.C + gg * 8 + j], static_cast<double>(sgx[j]));
      }
    } else if (blockIdx.x == 0 && rl == 0) {
      // the groups (and earlier launches of the step) accumulate into the same dgamma / dbeta
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&p.dbeta[gg * 8 + j], static_cast<float>(stat_read(sums + gg * 8 + j)));
        atomicAdd(&p.dgamma[gg * 8 + j], static_cast<float>(stat_read(sums + p.C + gg * 8 + j)));
      }
    }
  }
}

// Column reductions share one mapping: a CTA owns `ccta` (<= 256) adjacent columns of one row
// group (blockIdx.z = column chunk, blockIdx.y = row group) and 256 / (ccta / 8) row lanes, so
// even a 1024-wide tensor keeps >= 8 rows per CTA in flight; the row lanes are combined in
// shared memory and each column ends in ONE atomic per CTA.
struct RedMap {
  int cg, rstep, gl, rl, gg;
};
VPD_DEVINL RedMap red_map(int ccta) {
  RedMap m;
  m.cg = ccta >> 3;
  m.rstep = 256 / m.cg;
  m.gl = threadIdx.x % m.cg;
  m.rl = threadIdx.x / m.cg;
  m.gg = blockIdx.z * m.cg + m.gl;
  return m;
}
// combine the row lanes' partials (a[8], b[8] per thread) and hand column c's totals to f(c, sa, sb)
template <typename F>
VPD_DEVINL void red_finish(const RedMap& m, int ccta, const float* a, const float* b, bool two, F f) {
  __shared__ float s_a[2048], s_b[2048];
  if (m.rl < m.rstep) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_a[m.rl * ccta + m.gl * 8 + j] = a[j];
      if (two) s_b[m.rl * ccta + m.gl * 8 + j] = b[j];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < ccta) {
    float sa = 0.f, sb = 0.f;
    for (int l = 0; l < m.rstep; ++l) {
      sa += s_a[l * ccta + threadIdx.x];
      if (two) sb += s_b[l * ccta + threadIdx.x];
    }
    f(blockIdx.z * ccta + threadIdx.x, sa, sb);
  }
}

__global__ void __launch_bounds__(256)
bn1d_bwd_reduce_kernel(const Bn1dBwdParams p, int ccta) {
  pdl_trigger();
  pdl_wait();
  const RedMap m = red_map(ccta);
  const int grp = blockIdx.y;
  StatAcc* sums = p.sums + (size_t)grp * 2 * p.C;
  const size_t base = (size_t)grp * p.M * p.C;
  float sg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sgx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m.rl < m.rstep) {
    float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = m.gg * 8 + j;
      mean[j] = p.save_mean[(size_t)grp * p.C + c];
      rstd[j] = p.save_rstd[(size_t)grp * p.C + c];
      ga[j] = p.gamma[c];
      be[j] = p.beta[c];
    }
    for (long long r = (long long)blockIdx.x * m.rstep + m.rl; r < p.M; r += (long long)gridDim.x * m.rstep) {
      const size_t off = base + (size_t)r * p.C + m.gg * 8;
      const uint4 vd = ldg_nc_v4(p.dz + off), va = ldg_nc_v4(p.a + off);
      const float d[8] = {bf16_lo(vd.x), bf16_hi(vd.x), bf16_lo(vd.y), bf16_hi(vd.y),
                          bf16_lo(vd.z), bf16_hi(vd.z), bf16_lo(vd.w), bf16_hi(vd.w)};
      const float av[8] = {bf16_lo(va.x), bf16_hi(va.x), bf16_lo(va.y), bf16_hi(va.y),
                           bf16_lo(va.z), bf16_hi(va.z), bf16_lo(va.w), bf16_hi(va.w)};
      unsigned long long kp = 0x0101010101010101ull;
      if (p.keep) kp = *reinterpret_cast<const unsigned long long*>(p.keep + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sc = ga[j] * rstd[j];
        const float pre = fmaf(av[j], sc, be[j] - mean[j] * sc);   // same expression as forward
        const float xh = (av[j] - mean[j]) * rstd[j];
        const float gr = (pre > 0.f && ((kp >> (8 * j)) & 0xff)) ? d[j] * p.keep_scale : 0.f;
        sg[j] += gr;
        sgx[j] = fmaf(gr, xh, sgx[j]);
      }
    }
  }
  const int C = p.C;
  red_finish(m, ccta, sg, sgx, true, [sums, C](int c, float a, float b) {
    stat_add(&sums[c], static_cast<double>(a));
    stat_add(&sums[C + c], static_cast<double>(b));
  });
}

// stats[g][0][c] += sum_r x, stats[g][1][c] += sum_r x^2 over the rows of group g = blockIdx.y
__global__ void __launch_bounds__(256)
colstats_bf16_kernel(const __nv_bfloat16* __restrict__ x, StatAcc* stats, long long M, int C, int ccta) {
  pdl_trigger();
  pdl_wait();
  const RedMap m = red_map(ccta);
  const int grp = blockIdx.y;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m.rl < m.rstep) {
    for (long long r = (long long)blockIdx.x * m.rstep + m.rl; r < M; r += (long long)gridDim.x * m.rstep) {
      const uint4 v = ldg_nc_v4(x + ((size_t)grp * M + r) * C + m.gg * 8);
      const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                          bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
  }
  StatAcc* st = stats + (size_t)grp * 2 * C;
  red_finish(m, ccta, s, q, true, [st, C](int c, float a, float b) {
    stat_add(&st[c], static_cast<double>(a));
    stat_add(&st[C + c], static_cast<double>(b));
  });
}

// out = d * 1[z > 0]
__global__ void __launch_bounds__(256)
relu_mask_kernel(const __nv_bfloat16* __restrict__ d, const __nv_bfloat16* __restrict__ z,
                 __nv_bfloat16* out, long long n8) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 vd = ldg_nc_v4(d + i * 8), vz = ldg_nc_v4(z + i * 8);
    stg_v4(out + i * 8, make_uint4(vd.x & bf16x2_gt0_mask(vz.x), vd.y & bf16x2_gt0_mask(vz.y),
                                   vd.z & bf16x2_gt0_mask(vz.z), vd.w & bf16x2_gt0_mask(vz.w)));
  }
}

// out[c] += sum over rows of x[r][c]   (x bf16 [M][C], C % 8 == 0)
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* out, long long M, int C, int ccta) {
  pdl_trigger();
  pdl_wait();
  const RedMap m = red_map(ccta);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m.rl < m.rstep) {
    for (long long r = (long long)blockIdx.x * m.rstep + m.rl; r < M; r += (long long)gridDim.x * m.rstep) {
      const uint4 v = ldg_nc_v4(x + (size_t)r * C + m.gg * 8);
      s[0] += bf16_lo(v.x); s[1] += bf16_hi(v.x); s[2] += bf16_lo(v.y); s[3] += bf16_hi(v.y);
      s[4] += bf16_lo(v.z); s[5] += bf16_hi(v.z); s[6] += bf16_lo(v.w); s[7] += bf16_hi(v.w);
    }
  }
  red_finish(m, ccta, s, s, false, [out](int c, float a, float) { atomicAdd(out + c, a); });
}

// Loss head of Keypoint_EmbeddingModel.epoch (models/keypoint.py:58-104), one warp per sample:
//   contra  = ||e1 - e2||  (hinge_embedding_loss, target +1; only when e2 is given)
//           + valid * max(0, 1 - ||e1 - en||)  (target -1, margin 1; only when en is given)
//   loss    = contra + w3d * (sum (pred1 - true)^2 + sum (pred2 - true)^2)
// and its gradient scaled by `gscale` (= 1 / batch_n: the reference divides the summed loss by
// the number of samples before backward): de* fp32 [n][D], dpred* bf16 [n][Tpad].
struct VipeLossParams {
  const float* e1;
  const float* e2;
  const float* en;            // [n][D] (e2 / en may be null)
  const float* valid;         // [n] or null (all valid)
  const __nv_bfloat16* pred1;
  const __nv_bfloat16* pred2; // [n][Tpad] or null
  const float* true3d;        // [n][T] or null
  float* de1;
  float* de2;
  float* den;                 // [n][D]
  __nv_bfloat16* dpred1;
  __nv_bfloat16* dpred2;      // [n][Tpad]
  double* sums;               // [2]: contra, total (+=)
  long long n;
  int D, T, Tpad;
  float w3d, gscale;
};

__global__ void __launch_bounds__(256)
vipe_loss_kernel(const VipeLossParams p) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double contra_acc = 0.0, total_acc = 0.0;
  for (long long r = warp; r < p.n; r += nwarps) {
    float contra = 0.f, mse = 0.f;
    // embeddings: lane owns components lane, lane + 32
    float a[2] = {0.f, 0.f}, d12[2] = {0.f, 0.f}, d1n[2] = {0.f, 0.f};
    float s12 = 0.f, s1n = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane + 32 * j;
      if (c < p.D) {
        a[j] = p.e1[r * p.D + c];
        if (p.e2) { d12[j] = a[j] - p.e2[r * p.D + c]; s12 = fmaf(d12[j], d12[j], s12); }
        if (p.en) { d1n[j] = a[j] - p.en[r * p.D + c]; s1n = fmaf(d1n[j], d1n[j], s1n); }
      }
    }
    s12 = warp_sum(s12);
    s1n = warp_sum(s1n);
    const float n12 = sqrtf(s12), n1n = sqrtf(s1n);
    float c12 = 0.f, c1n = 0.f;                 // d loss / d (e1 - e2), d loss / d (e1 - en)
    if (p.e2) {
      contra += n12;
      c12 = n12 > 0.f ? 1.f / n12 : 0.f;
    }
    if (p.en) {
      const float v = p.valid ? p.valid[r] : 1.f;
      const float h = 1.f - n1n;
      if (h > 0.f) {
        contra += v * h;
        c1n = n1n > 0.f ? -v / n1n : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane + 32 * j;
      if (c < p.D) {
        const float g12 = c12 * d12[j] * p.gscale, g1n = c1n * d1n[j] * p.gscale;
        p.de1[r * p.D + c] = g12 + g1n;
        if (p.de2) p.de2[r * p.D + c] = -g12;
        if (p.den) p.den[r * p.D + c] = -g1n;
      }
    }
    if (p.true3d) {
      for (int c = lane; c < p.Tpad; c += 32) {
        const float t = c < p.T ? p.true3d[r * p.T + c] : 0.f;
        const float e1 = c < p.T ? __bfloat162float(p.pred1[r * p.Tpad + c]) - t : 0.f;
        mse = fmaf(e1, e1, mse);
        p.dpred1[r * p.Tpad + c] = __float2bfloat16_rn(2.f * p.w3d * e1 * p.gscale);
        if (p.pred2) {
          const float e2 = c < p.T ? __bfloat162float(p.pred2[r * p.Tpad + c]) - t : 0.f;
          mse = fmaf(e2, e2, mse);
          p.dpred2[r * p.Tpad + c] = __float2bfloat16_rn(2.f * p.w3d * e2 * p.gscale);
        }
      }
      mse = warp_sum(mse);
    }
    contra_acc += static_cast<double>(contra);
    total_acc += static_cast<double>(contra) + static_cast<double>(p.w3d) * static_cast<double>(mse);
  }
  if (lane == 0 && (contra_acc != 0.0 || total_acc != 0.0)) {
    atomicAdd(&p.sums[0], contra_acc);
    atomicAdd(&p.sums[1], total_acc);
  }
}

static unsigned grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

int rows_to_bf16(const float* x, __nv_bfloat16* out, long long M, int C, int Cpad,
                 cudaStream_t stream) {
  VPD_REQUIRE(M >= 0 && C >= 1 && Cpad >= C && Cpad % 8 == 0,
              "rows_to_bf16: need 1 <= C <= Cpad, Cpad %% 8 == 0 (C=%d, Cpad=%d)", C, Cpad);
  if (M == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(rows_to_bf16_kernel, dim3(grid_for(M * (Cpad / 8), 256)), dim3(256),
                               0, stream, x, out, M, C, Cpad));
  VPD_LAUNCHED(1);
  return 0;
}

int axpby_bf16(const __nv_bfloat16* a, float alpha, const __nv_bfloat16* b, float beta,
               __nv_bfloat16* out, long long n, cudaStream_t stream) {
  VPD_REQUIRE(n >= 0 && n % 8 == 0, "axpby_bf16: element count must be a multiple of 8");
  VPD_REQUIRE(a != nullptr && out != nullptr, "axpby_bf16: null operand");
  if (n == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(axpby_bf16_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, stream,
                               a, alpha, b, beta, out, n / 8));
  VPD_LAUNCHED(1);
  return 0;
}

int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
            const float* bias, float eps, float* scale, float* shift, int C, cudaStream_t stream) {
  VPD_REQUIRE(C >= 1, "bn_fold: C must be positive");
  VPD_CHECK_CUDA(launch_kernel(bn_fold_kernel, dim3((C + 255) / 256), dim3(256), 0, stream, gamma,
                               beta, mean, var, bias, eps, scale, shift, C));
  VPD_LAUNCHED(1);
  return 0;
}

int linear_rows_f32(const __nv_bfloat16* x, const float* w, const float* bias, float* out,
                    long long M, int K, int D, cudaStream_t stream) {
  VPD_REQUIRE(D >= 1 && D <= 64, "linear_rows_f32: 1 <= D <= 64 (got %d)", D);
  VPD_REQUIRE(K >= kLinK && K % kLinK == 0, "linear_rows_f32: K must be a multiple of %d", kLinK);
  if (M <= 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(linear_rows_f32_kernel, dim3((unsigned)((M + kLinRows - 1) / kLinRows)),
                               dim3(256), 0, stream, x, w, bias, out, M, K, D));
  VPD_LAUNCHED(1);
  return 0;
}

__global__ void __launch_bounds__(256) zero_acc_kernel(StatAcc* p, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = StatAcc{0, 0};
}

static unsigned rows_grid(long long M, int C, long long cap = 148 * 4) {
  const int rstep = 256 / (C / 8);
  long long blocks = (M + rstep - 1) / rstep;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
// column-reduction launches: columns per CTA (<= 256, divides C) and the (x, y, z) grid
static int red_ccta(int C) {
  if (C <= 256) return C;
  for (int c = 256; c >= 8; c -= 8)
    if (C % c == 0) return c;
  return 8;
}
static dim3 red_grid(long long M, int C, int groups) {
  const int ccta = red_ccta(C), nz = C / ccta, rstep = 256 / (ccta / 8);
  long long nx = 296 / ((long long)groups * nz);
  const long long need = (M + rstep - 1) / rstep;
  if (nx > need) nx = need;
  if (nx < 1) nx = 1;
  return dim3((unsigned)nx, (unsigned)groups, (unsigned)nz);
}
#define VPD_ROWS_OK(C) VPD_REQUIRE((C) >= 8 && (C) % 8 == 0 && (C) <= 2048, "row kernels: C %% 8 == 0, C <= 2048 (got %d)", (C))

int dropout_mask(uint8_t* keep, long long n, float p_drop, unsigned long long seed,
                 const unsigned long long* seed_add, unsigned int stream_id, cudaStream_t stream) {
  VPD_REQUIRE(n >= 0 && n % 4 == 0, "dropout_mask: element count must be a multiple of 4");
  VPD_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout_mask: p must be in [0, 1)");
  if (n == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(dropout_mask_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, stream,
                               keep, n / 4, p_drop, seed, seed_add, stream_id));
  VPD_LAUNCHED(1);
  return 0;
}

int bn1d_fwd(const __nv_bfloat16* a, const StatAcc* stats, const float* gamma, const float* beta,
             const float* lin_bias, float* running_mean, float* running_var, long long* num_batches,
             float* save_mean, float* save_rstd, const uint8_t* keep, float p_drop,
             const __nv_bfloat16* res, __nv_bfloat16* out, long long M, int C, int groups,
             cudaStream_t stream) {
  VPD_ROWS_OK(C);
  VPD_REQUIRE(M >= 1 && groups >= 1 && groups <= 64, "bn1d_fwd: empty batch / bad group count");
  VPD_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "bn1d_fwd: p must be in [0, 1)");
  Bn1dParams p;
  p.a = a; p.stats = stats; p.gamma = gamma; p.beta = beta; p.lin_bias = lin_bias;
  p.running_mean = running_mean; p.running_var = running_var; p.num_batches = num_batches;
  p.save_mean = save_mean; p.save_rstd = save_rstd; p.keep = keep;
  p.keep_scale = 1.f / (1.f - p_drop);
  p.res = res; p.out = out; p.M = M; p.C = C; p.eps = 1e-5f; p.momentum = 0.1f;
  VPD_CHECK_CUDA(launch_kernel(bn1d_fwd_kernel, dim3(rows_grid(M, C, 148 * 4 / groups + 1), groups), dim3(256), 0, stream, p));
  VPD_LAUNCHED(1);
  return 0;
}

int bn1d_bwd(const __nv_bfloat16* dz, const __nv_bfloat16* a, const uint8_t* keep, float p_drop,
             const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
             StatAcc* sums, __nv_bfloat16* da, float* dgamma, float* dbeta, long long M, int C,
             int groups, cudaStream_t stream) {
  VPD_ROWS_OK(C);
  VPD_REQUIRE(M >= 1 && groups >= 1 && groups <= 64, "bn1d_bwd: empty batch / bad group count");
  Bn1dBwdParams p;
  p.dz = dz; p.a = a; p.keep = keep; p.keep_scale = 1.f / (1.f - p_drop);
  p.gamma = gamma; p.beta = beta; p.save_mean = save_mean; p.save_rstd = save_rstd;
  p.sums = sums; p.da = da; p.dgamma = dgamma; p.dbeta = dbeta; p.M = M; p.C = C;
  // a kernel, not a memset node: the step is captured into a CUDA graph with programmatic edges
  VPD_CHECK_CUDA(launch_kernel(zero_acc_kernel, dim3((2 * C * groups + 255) / 256), dim3(256), 0, stream, sums, 2 * C * groups));
  VPD_CHECK_CUDA(launch_kernel(bn1d_bwd_reduce_kernel, red_grid(M, C, groups), dim3(256), 0, stream, p, red_ccta(C)));
  VPD_CHECK_CUDA(launch_kernel(bn1d_bwd_kernel<true>, dim3(rows_grid(M, C, 148 * 4 / groups + 1), groups), dim3(256), 0, stream, p));
  VPD_LAUNCHED(3);
  return 0;
}

int colstats_bf16(const __nv_bfloat16* x, StatAcc* stats, long long M, int C, int groups,
                  cudaStream_t stream) {
  VPD_ROWS_OK(C);
  VPD_REQUIRE(M >= 1 && groups >= 1 && groups <= 64, "colstats: empty batch / bad group count");
  VPD_CHECK_CUDA(launch_kernel(zero_acc_kernel, dim3((2 * C * groups + 255) / 256), dim3(256), 0, stream, stats, 2 * C * groups));
  VPD_CHECK_CUDA(launch_kernel(colstats_bf16_kernel, red_grid(M, C, groups), dim3(256), 0, stream, x, stats, M, C, red_ccta(C)));
  VPD_LAUNCHED(2);
  return 0;
}

int relu_mask_bf16(const __nv_bfloat16* d, const __nv_bfloat16* z, __nv_bfloat16* out, long long n,
                   cudaStream_t stream) {
  VPD_REQUIRE(n >= 0 && n % 8 == 0, "relu_mask: element count must be a multiple of 8");
  if (n == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(relu_mask_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, stream, d, z,
                               out, n / 8));
  VPD_LAUNCHED(1);
  return 0;
}

int colsum_bf16(const __nv_bfloat16* x, float* out, long long M, int C, cudaStream_t stream) {
  VPD_ROWS_OK(C);
  if (M <= 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(colsum_bf16_kernel, red_grid(M, C, 1), dim3(256), 0, stream, x, out, M, C, red_ccta(C)));
  VPD_LAUNCHED(1);
  return 0;
}

int vipe_loss(const float* e1, const float* e2, const float* en, const float* valid,
              const __nv_bfloat16* pred1, const __nv_bfloat16* pred2, const float* true3d,
              float* de1, float* de2, float* den, __nv_bfloat16* dpred1, __nv_bfloat16* dpred2,
              double* sums, long long n, int D, int T, int Tpad, float w3d, float gscale,
              cudaStream_t stream) {
  VPD_REQUIRE(D >= 1 && D <= 64, "vipe_loss: 1 <= D <= 64 (got %d)", D);
  VPD_REQUIRE(e1 != nullptr && de1 != nullptr && sums != nullptr, "vipe_loss: null e1 / de1 / sums");
  VPD_REQUIRE((e2 == nullptr) == (de2 == nullptr) && (en == nullptr) == (den == nullptr),
              "vipe_loss: every given embedding needs its gradient buffer");
  VPD_REQUIRE(true3d == nullptr || (pred1 != nullptr && dpred1 != nullptr && T >= 1 && Tpad >= T),
              "vipe_loss: 3-D targets need pred1 / dpred1 and 1 <= T <= Tpad");
  VPD_REQUIRE((pred2 == nullptr) == (dpred2 == nullptr), "vipe_loss: pred2 and dpred2 come together");
  if (n <= 0) return 0;
  VipeLossParams p;
  p.e1 = e1; p.e2 = e2; p.en = en; p.valid = valid; p.pred1 = pred1; p.pred2 = pred2;
  p.true3d = true3d; p.de1 = de1; p.de2 = de2; p.den = den; p.dpred1 = dpred1; p.dpred2 = dpred2;
  p.sums = sums; p.n = n; p.D = D; p.T = T; p.Tpad = Tpad; p.w3d = w3d; p.gscale = gscale;
  VPD_CHECK_CUDA(launch_kernel(vipe_loss_kernel, dim3(grid_for(n * 32, 256)), dim3(256), 0, stream, p));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
