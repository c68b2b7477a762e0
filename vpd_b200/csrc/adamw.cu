// K5: fused AdamW over the flat parameter arena.
//
// One launch updates every parameter of the encoder and decoder
// (torch.optim.AdamW defaults as built at train_vpd_model.py:100-105: decoupled
// weight decay applied to every tensor). The fp32 rounding sequence follows
// torch's single-tensor CPU path op for op (see oracle/student_ref.py
// `adamw_step_numpy`), so with identical gradients the moments are bit-exact and
// the parameters differ only where torch's vectorised sqrt is itself 1 ulp off:
//     p1  = p * (1 - lr*wd)
//     m1  = fma(1-b1, g - m, m)
//     v1  = fma((1-b2) * g, g, v * b2)
//     den = sqrt(v1) / sqrt(bc2) + eps
//     p2  = p1 + (-(lr/bc1) * m1) / den
// HBM-bound: 16 B read (p,g,m,v) + 12 B written (p,m,v) per parameter, 128-bit
// accesses, grid-stride over a multiple of the SM count.
#include <math.h>

#include "common.cuh"
#include "tma_host.h"

namespace vpd {

struct AdamScalars {
  float decay;     // 1 - lr*wd
  float w1;        // 1 - b1
  float b2, w2;    // b2, 1 - b2
  float bc2_sqrt;  // sqrt(1 - b2^t)
  float eps;
  float neg_step;  // -(lr / (1 - b1^t))
  float gscale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v,
                                         const AdamScalars& s) {
  if (s.gscale != 1.f) g = __fmul_rn(g, s.gscale);
  const float p1 = __fmul_rn(p, s.decay);
  const float m1 = __fmaf_rn(s.w1, __fsub_rn(g, m), m);
  const float v1 = __fmaf_rn(__fmul_rn(s.w2, g), g, __fmul_rn(v, s.b2));
  const float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v1), s.bc2_sqrt), s.eps);
  p = __fadd_rn(p1, __fdiv_rn(__fmul_rn(s.neg_step, m1), den));
  m = m1;
  v = v1;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, long long n, const AdamScalars s) {
  pdl_trigger();
  pdl_wait();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, s);
    adam_one(pp.y, gg.y, mm.y, vv.y, s);
    adam_one(pp.z, gg.z, mm.z, vv.z, s);
    adam_one(pp.w, gg.w, mm.w, vv.w, s);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n % 4 elements)
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) adam_one(p[t], g[t], m[t], v[t], s);
}

static AdamScalars adam_scalars(double lr, double b1, double b2, double eps, double wd, int step,
                                float grad_scale);

int adamw_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
               double b2, double eps, double wd, int step, float grad_scale,
               cudaStream_t stream) {
  VPD_REQUIRE(step >= 1, "adamw: step must be >= 1");
  VPD_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
              "adamw: arenas must be 16-byte aligned");
  if (n == 0) return 0;
  const AdamScalars s = adam_scalars(lr, b1, b2, eps, wd, step, grad_scale);
  long long blocks = ((n >> 2) + 255) / 256;
  const long long cap = (long long)148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  VPD_CHECK_CUDA(launch_kernel(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, p, g, m, v, n, s));
  VPD_LAUNCHED(1);
  return 0;
}

// ------------------------------------------------- AdamW that also writes the bf16 mirrors
// The tensor-core kernels read the conv weights as bf16 in two pre-tiled layouts (forward
// operand [tap][Cout][Cin] and its transpose for the data gradients, see wtile_offset). They
// used to be refreshed from the fp32 masters by a separate pass at the start of every step
// (85 MB re-read + 85 MB written); here the optimizer writes them while the updated weights
// are still in registers: 32 B per conv parameter instead of 28 + 12.
// One CTA owns a 32 (cout) x 64 (cin) tile of one tap's matrix: float4 accesses along cin
// (256-byte runs per row), the transpose goes through shared memory. Same rounding sequence
// as adamw_kernel (bit-identical parameters and moments).
struct MirrorTile {
  long long base;   // element offset of the tap's [rows][cols] matrix inside the conv section
  int rows, cols;   // Cout, Cin
  int r0, c0;       // first row / column of this CTA's tile
};

__global__ void __launch_bounds__(256)
adamw_mirror_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                    float* __restrict__ v, __nv_bfloat16* __restrict__ w_tap,
                    __nv_bfloat16* __restrict__ wT, const int* __restrict__ table,
                    const AdamScalars s) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][65];
  const int* e = table + blockIdx.x * 5;
  const long long base = e[0];
  const int rows = e[1], cols = e[2], r0 = e[3] * 32, c0 = e[4] * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 float4 columns x 16 rows
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int r = ty + it * 16;
    const long long idx = base + (long long)(r0 + r) * cols + c0 + tx * 4;
    float4 pp = *reinterpret_cast<const float4*>(p + idx);
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g + idx));
    float4 mm = *reinterpret_cast<const float4*>(m + idx);
    float4 vv = *reinterpret_cast<const float4*>(v + idx);
    adam_one(pp.x, gg.x, mm.x, vv.x, s);
    adam_one(pp.y, gg.y, mm.y, vv.y, s);
    adam_one(pp.z, gg.z, mm.z, vv.z, s);
    adam_one(pp.w, gg.w, mm.w, vv.w, s);
    *reinterpret_cast<float4*>(p + idx) = pp;
    *reinterpret_cast<float4*>(m + idx) = mm;
    *reinterpret_cast<float4*>(v + idx) = vv;
    uint2 b;
    b.x = pack_bf16x2(pp.x, pp.y);
    b.y = pack_bf16x2(pp.z, pp.w);
    *reinterpret_cast<uint2*>(w_tap + base + wtile_offset(r0 + r, c0 + tx * 4, rows)) = b;
    tile[r][tx * 4 + 0] = pp.x;
    tile[r][tx * 4 + 1] = pp.y;
    tile[r][tx * 4 + 2] = pp.z;
    tile[r][tx * 4 + 3] = pp.w;
  }
  __syncthreads();
  // transposed operand: rows' = cin, k' = cout. Thread -> (cin row c, 8 consecutive couts):
  // 64 cin rows x 4 groups of 8 couts = 256 threads, one 16-byte store each
  const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
  uint4 o;
  o.x = pack_bf16x2(tile[q * 8 + 0][c], tile[q * 8 + 1][c]);
  o.y = pack_bf16x2(tile[q * 8 + 2][c], tile[q * 8 + 3][c]);
  o.z = pack_bf16x2(tile[q * 8 + 4][c], tile[q * 8 + 5][c]);
  o.w = pack_bf16x2(tile[q * 8 + 6][c], tile[q * 8 + 7][c]);
  *reinterpret_cast<uint4*>(wT + base + wtile_offset(c0 + c, r0 + q * 8, cols)) = o;
}

static AdamScalars adam_scalars(double lr, double b1, double b2, double eps, double wd, int step,
                                float grad_scale) {
  AdamScalars s;
  const double b1t = pow(b1, (double)step), b2t = pow(b2, (double)step);
  const double bc1 = 1.0 - b1t, bc2 = 1.0 - b2t;
  s.decay = (float)(1.0 - lr * wd);
  s.w1 = (float)(1.0 - b1);
  s.b2 = (float)b2;
  s.w2 = (float)(1.0 - b2);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.eps = (float)eps;
  s.neg_step = (float)(-(lr / bc1));
  s.gscale = grad_scale;
  return s;
}

// the mirror-writing kernel over `tiles` table entries (pointers at the start of the conv section)
int adamw_tiles(float* p_conv, const float* g_conv, float* m_conv, float* v_conv,
                __nv_bfloat16* w_tap, __nv_bfloat16* wT, const int* table, int tiles, double lr,
                double b1, double b2, double eps, double wd, int step, float grad_scale,
                cudaStream_t stream) {
  VPD_REQUIRE(step >= 1, "adamw: step must be >= 1");
  if (tiles <= 0) return 0;
  const AdamScalars s = adam_scalars(lr, b1, b2, eps, wd, step, grad_scale);
  VPD_CHECK_CUDA(launch_kernel(adamw_mirror_kernel, dim3((unsigned)tiles), dim3(256), 0, stream,
                               p_conv, g_conv, m_conv, v_conv, w_tap, wT, table, s));
  VPD_LAUNCHED(1);
  return 0;
}

// conv section [conv_off, conv_off + conv_len) of the arenas: tile kernel + mirrors; the rest
// (BN affine, fc, decoder: a few 10^4 parameters) through the flat kernel.
int adamw_step_mirrored(float* p, const float* g, float* m, float* v, long long n,
                        long long conv_off, long long conv_len, __nv_bfloat16* w_tap,
                        __nv_bfloat16* wT, const int* table, int tiles, double lr, double b1,
                        double b2, double eps, double wd, int step, float grad_scale,
                        cudaStream_t stream) {
  VPD_REQUIRE(step >= 1, "adamw: step must be >= 1");
  VPD_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0 &&
                  conv_off % 4 == 0 && conv_len % 4 == 0,
              "adamw: arenas must be 16-byte aligned");
  const AdamScalars s = adam_scalars(lr, b1, b2, eps, wd, step, grad_scale);
  VPD_CHECK_CUDA(launch_kernel(adamw_mirror_kernel, dim3((unsigned)tiles), dim3(256), 0, stream,
                               p + conv_off, g + conv_off, m + conv_off, v + conv_off, w_tap, wT,
                               table, s));
  int launched = 1;
  const long long lo[2] = {0, conv_off + conv_len}, len[2] = {conv_off, n - conv_off - conv_len};
  for (int i = 0; i < 2; ++i) {
    if (len[i] <= 0) continue;
    long long blocks = ((len[i] >> 2) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    VPD_CHECK_CUDA(launch_kernel(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, stream,
                                 p + lo[i], g + lo[i], m + lo[i], v + lo[i], len[i], s));
    ++launched;
  }
  VPD_LAUNCHED(launched);
  return 0;
}

// ---------------------------------------------------------------- fused SGD
// torch.optim.SGD (single-tensor path, torch/optim/sgd.py::_single_tensor_sgd) over the flat
// arena, one launch:
//     g1  = g + wd * p                       (weight_decay != 0)
//     buf = g1                 (first step)  |  buf = momentum * buf + (1 - dampening) * g1
//     g2  = nesterov ? g1 + momentum * buf : buf          (momentum != 0)
//     p   = p - lr * g2
// every `a + alpha * b` is one fma, which is what ATen's vectorised add(alpha) does.
struct SgdScalars {
  float lr, momentum, one_minus_damp, wd, gscale;
  int nesterov, first;
};

__device__ __forceinline__ void sgd_one(float& p, float g, float* buf, const SgdScalars& s) {
  if (s.gscale != 1.f) g = __fmul_rn(g, s.gscale);
  if (s.wd != 0.f) g = __fmaf_rn(s.wd, p, g);
  if (s.momentum != 0.f) {
    float b;
    if (s.first) b = g;
    else b = __fmaf_rn(s.one_minus_damp, g, __fmul_rn(*buf, s.momentum));
    *buf = b;
    g = s.nesterov ? __fmaf_rn(s.momentum, b, g) : b;
  }
  p = __fmaf_rn(-s.lr, g, p);
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
           const SgdScalars s) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pp = p[i];
    sgd_one(pp, __ldcs(g + i), buf ? buf + i : nullptr, s);
    p[i] = pp;
  }
}

int sgd_step(float* p, const float* g, float* buf, long long n, double lr, double momentum,
             double dampening, double wd, int nesterov, int first_step, float grad_scale,
             cudaStream_t stream) {
  VPD_REQUIRE(momentum == 0.0 || buf != nullptr, "sgd: momentum needs a buffer");
  VPD_REQUIRE(!nesterov || (momentum > 0.0 && dampening == 0.0),
              "sgd: nesterov needs momentum > 0 and zero dampening");
  if (n == 0) return 0;
  SgdScalars s;
  s.lr = (float)lr;
  s.momentum = (float)momentum;
  s.one_minus_damp = (float)(1.0 - dampening);
  s.wd = (float)wd;
  s.gscale = grad_scale;
  s.nesterov = nesterov;
  s.first = first_step;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)148 * 8;
  if (blocks > cap) blocks = cap;
  VPD_CHECK_CUDA(launch_kernel(sgd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, p, g, buf, n, s));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
