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
__global__ void __launch_bounds__(256)
linear_rows_f32_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, float* __restrict__ out, long long M,
                       int K, int D) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_w[64][kLinK + 1];
  __shared__ float s_x[kLinRows][kLinK + 1];
  const int tid = threadIdx.x;
  const int r = tid >> 3, q = tid & 7;
  const long long row0 = (long long)blockIdx.x * kLinRows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kLinK) {
    for (int i = tid; i < D * kLinK; i += 256) {
      const int d = i / kLinK, k = i - d * kLinK;
      s_w[d][k] = __ldg(w + (size_t)d * K + k0 + k);
    }
    for (int i = tid; i < kLinRows * (kLinK / 8); i += 256) {
      const int rr = i / (kLinK / 8), g = i - rr * (kLinK / 8);
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row0 + rr < M) {
        const uint4 u = ldg_nc_v4(x + (size_t)(row0 + rr) * K + k0 + g * 8);
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_x[rr][g * 8 + j] = v[j];
    }
    __syncthreads();
    for (int k = 0; k < kLinK; ++k) {
      const float xv = s_x[r][k];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (q + 8 * j < D) acc[j] = fmaf(xv, s_w[q + 8 * j][k], acc[j]);
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

}  // namespace vpd
