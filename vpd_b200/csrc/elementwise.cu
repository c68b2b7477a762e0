#include "elementwise.cuh"

#include "tma_host.h"

namespace vpd {

constexpr int kEwThreads = 256;

// Deterministic block-level combine of per-thread column partials. Thread t = r0 * groups + g
// holds eight partial sums for channels g*8 .. g*8+7 of its row lane r0; they go to
// scr[r0][C] and channel c is then added up over the row lanes IN LANE ORDER by one thread.
// (Shared-memory float atomics would add them in arrival order, so the rounding - and through
// the BatchNorm statistics every activation downstream - would change from run to run.)
constexpr int kColScratch = kEwThreads * 8;   // floats per quantity
VPD_DEVINL void colsum_put(const float* acc, float* scr, int C, int g, int r0) {
  float4* d = reinterpret_cast<float4*>(scr + r0 * C + g * 8);
  d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
VPD_DEVINL float colsum_get(const float* scr, int C, int lanes, int c) {
  float s = 0.f;
  for (int r = 0; r < lanes; ++r) s += scr[r * C + c];
  return s;
}

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
// bit j = 1[element j of the packed bf16 vector > 0] (the stored, rounded values decide)
// The vector holds ReLU outputs: no negative values and no -0 (fmaxf(x, 0.f) returns +0), so
// "> 0" is "bit pattern != 0" and min(half, 1) is the flag of a half (one SIMD instruction per
// packed pair); NaNs (never produced by a finite network) would count as positive.
VPD_DEVINL uint32_t gt0_bits_relu(const uint4& v) {
  const uint32_t x = __vminu2(v.x, 0x00010001u) + (__vminu2(v.y, 0x00010001u) << 2) +
                     (__vminu2(v.z, 0x00010001u) << 4) + (__vminu2(v.w, 0x00010001u) << 6);
  return (x & 0x55u) | ((x >> 15) & 0xAAu);   // even channels sit at bits 0,2,4,6, odd at 16,18,..
}
// any bf16 vector (negative values, -0, NaN -> 0)
VPD_DEVINL uint32_t gt0_bits(const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t b = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t m = bf16x2_gt0_mask(w[k]);
    b |= ((m & 1u) << (2 * k)) | ((m >> 31) << (2 * k + 1));
  }
  return b;
}

// Block 0: persist batch statistics and update the running buffers like
// nn.BatchNorm2d (momentum 0.1, unbiased variance for the running estimate).
VPD_DEVINL void bn_side_effects(const BnLayer& bn, int C) {
  if (bn.stats == nullptr) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd, var;
    bn_mean_rstd(bn, c, C, mean, rstd, var);
    bn_channel_side_effects(bn, c, mean, rstd, var);
  }
  if (bn.update_running && threadIdx.x == 0 && bn.num_batches) *bn.num_batches += 1;
}

// The BatchNorm affine of channel c is derived ONCE per CTA (by thread c, through the fp64
// mean / variance path on the integer statistics accumulators) and shared through shared
// memory, instead of by every thread that owns the channel: these kernels are latency-bound
// and the fp64 prologue was a measurable part of every launch.
constexpr int kMaxEwC = 2048;
VPD_DEVINL void bn_stage_affine(const BnLayer& bn, int C, float* s_sc, float* s_sh) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd, var, sc, sh;
    bn_mean_rstd(bn, c, C, mean, rstd, var);
    bn_affine(bn.gamma[c], bn.beta[c], mean, rstd, sc, sh);
    s_sc[c] = sc;
    s_sh[c] = sh;
  }
}
VPD_DEVINL void load8(const float* s, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(s);
  const float4 b = *reinterpret_cast<const float4*>(s + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w;
  f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

// ------------------------------------------------------------------ BN apply
template <int RES>  // 0: no residual, 1: plain residual, 2: residual through its own BN
__global__ void __launch_bounds__(kEwThreads, RES == 0 ? 4 : 3) bn_apply_kernel(const BnApplyParams p) {
  pdl_trigger();
  pdl_wait();
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int r0 = threadIdx.x / groups;
  const int rstep = kEwThreads / groups;
  __shared__ __align__(16) float s_co[(RES == 2 ? 4 : 2) * kMaxEwC];
  bn_stage_affine(p.bn, p.C, s_co, s_co + kMaxEwC);
  if (RES == 2) bn_stage_affine(p.res_bn, p.C, s_co + 2 * kMaxEwC, s_co + 3 * kMaxEwC);
  __syncthreads();
  float sc[8], sh[8], rsc[8], rsh[8];
  load8(s_co + g * 8, sc);
  load8(s_co + kMaxEwC + g * 8, sh);
  if (RES == 2) {
    load8(s_co + 2 * kMaxEwC + g * 8, rsc);
    load8(s_co + 3 * kMaxEwC + g * 8, rsh);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) rsc[j] = 1.f, rsh[j] = 0.f;
  }
  const long long chunk = (p.M + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * chunk;
  const long long end = beg + chunk < p.M ? beg + chunk : p.M;
  constexpr int U = 4;  // independent rows in flight per thread
  for (long long row = beg + r0; row < end; row += (long long)rstep * U) {
    uint4 vy[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r < end) {
        const size_t off = (size_t)r * p.C + g * 8;
        vy[u] = ldg_nc_v4(p.y + off);
        if (RES != 0) vr[u] = ldg_nc_v4(p.res + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r >= end) break;
      const size_t off = (size_t)r * p.C + g * 8;
      float f[8];
      unpack8(vy[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
      if (RES != 0) {
        float rr[8];
        unpack8(vr[u], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += RES == 2 ? fmaf(rr[j], rsc[j], rsh[j]) : rr[j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      const uint4 zv = pack8(f);
      stg_v4(p.z + off, zv);
      if (p.mask != nullptr)
        p.mask[(size_t)r * groups + g] = static_cast<uint8_t>(p.relu ? gt0_bits_relu(zv) : gt0_bits(zv));
    }
  }
  // every block has consumed the statistics above before block 0 may touch
  // the running buffers it also reads in eval mode; in train mode the inputs
  // (stats) are never modified here, so no ordering issue.
  if (blockIdx.x == 0) {
    bn_side_effects(p.bn, p.C);
    if (RES == 2) bn_side_effects(p.res_bn, p.C);
  }
}

// One wave at most: `per_thread` vectors per thread, never more CTAs than
// `blocks_per_sm` (the kernel's real residency) x SMs - these kernels are latency
// bound on small tensors, so a second wave costs a full prologue + memory round trip.
static int ew_grid(long long vectors, int per_thread, int blocks_per_sm) {
  long long blocks = (vectors + (long long)kEwThreads * per_thread - 1) / (kEwThreads * per_thread);
  const long long cap = 148LL * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int launch_bn_apply(const BnApplyParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 2048 && kEwThreads % (p.C / 8) == 0,
              "bn_apply: unsupported channel count %d", p.C);
  if (p.M == 0) return 0;
  const long long vectors = p.M * (p.C / 8);
  if (p.res == nullptr)
    VPD_CHECK_CUDA(launch_kernel(bn_apply_kernel<0>, dim3(ew_grid(vectors, 4, 4)), dim3(kEwThreads), 0, s, p));
  else if (!p.has_res_bn)
    VPD_CHECK_CUDA(launch_kernel(bn_apply_kernel<1>, dim3(ew_grid(vectors, 4, 3)), dim3(kEwThreads), 0, s, p));
  else
    VPD_CHECK_CUDA(launch_kernel(bn_apply_kernel<2>, dim3(ew_grid(vectors, 4, 3)), dim3(kEwThreads), 0, s, p));
  VPD_LAUNCHED(1);
  return 0;
}

__global__ void __launch_bounds__(kEwThreads) relu_mask_kernel(const __nv_bfloat16* __restrict__ z,
                                                               uint8_t* __restrict__ mask, long long vectors) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < vectors;
       i += (long long)gridDim.x * kEwThreads)
    mask[i] = static_cast<uint8_t>(gt0_bits(ldg_nc_v4(z + i * 8)));
}
int launch_relu_mask(const __nv_bfloat16* z, uint8_t* mask, long long M, int C, cudaStream_t s) {
  VPD_REQUIRE(C % 8 == 0, "relu_mask: C=%d must be a multiple of 8", C);
  const long long vectors = M * (C / 8);
  if (vectors == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(relu_mask_kernel, dim3(ew_grid(vectors, 4, 8)), dim3(kEwThreads), 0, s, z, mask, vectors));
  VPD_LAUNCHED(1);
  return 0;
}

// ---------------------------------------------- stem: BN + ReLU + maxpool 3x3/2
// The kernel is bound by instruction issue (ncu: 64 % issue-active at 24 % of DRAM
// throughput, 960 instructions per thread in the first version), so the inner loop is pared
// down: 32-bit index arithmetic, the maximum is taken over the BN outputs BEFORE the ReLU
// (max relu(a) = relu(max a): nine fmax fewer per channel), and the running maximum carries
// its payload - the pre-BN bf16 value in the upper half, the window index in the lower - in ONE
// register, so a new maximum costs two selects instead of three. The window index follows
// torch's max_pool2d on the ReLU output: the first position that attains the maximum, which is
// the first valid position when nothing in the window is positive.
template <bool kInterior>
VPD_DEVINL void pool_window(const __nv_bfloat16* __restrict__ y0, int rs, int C, int ho, int wo,
                            const float (&sc)[8], const float (&sh)[8], float (&best)[8],
                            uint32_t (&pay)[8]) {
  // y0 -> window position (0, 0); rs = row stride in elements
  uint4 win[9];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const bool ok = kInterior || ((kh > 0 || ho > 0) && (kw > 0 || wo > 0));
      win[kh * 3 + kw] = ok ? __ldg(reinterpret_cast<const uint4*>(y0 + kh * rs + kw * C))
                            : make_uint4(0, 0, 0, 0);
    }
  const int kf = kInterior ? 0 : (ho == 0 ? 3 : 0) + (wo == 0 ? 1 : 0);   // first valid position
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    best[j] = -INFINITY;
    pay[j] = 0u;
  }
  uint32_t first[8];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const bool ok = kInterior || ((k >= 3 || ho > 0) && (k % 3 > 0 || wo > 0));
    const uint32_t w4[4] = {win[k].x, win[k].y, win[k].z, win[k].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t fb = (j & 1) ? (w4[j >> 1] & 0xFFFF0000u) : (w4[j >> 1] << 16);
      const float a = fmaf(__uint_as_float(fb), sc[j], sh[j]);
      const bool up = ok && a > best[j];
      best[j] = up ? a : best[j];
      pay[j] = up ? (fb | static_cast<uint32_t>(k)) : pay[j];
      if (kInterior ? k == 0 : k == kf) first[j] = fb | static_cast<uint32_t>(k);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (!(best[j] > 0.f)) {   // nothing positive: ReLU output 0 everywhere, first position wins
      best[j] = 0.f;
      pay[j] = first[j];
    }
  }
}

__global__ void __launch_bounds__(kEwThreads, 2) bn_pool_kernel(const PoolParams p) {
  pdl_trigger();
  pdl_wait();
  const int groups = p.C >> 3;
  const int Ho = p.H / 2, Wo = p.W / 2;
  const unsigned npix = (unsigned)p.N * Ho * Wo;
  const int g = threadIdx.x % groups;
  const unsigned ppc = kEwThreads / groups;      // pooled pixels per CTA and pass
  __shared__ __align__(16) float s_co[2 * 512];
  bn_stage_affine(p.bn, p.C, s_co, s_co + 512);
  __syncthreads();
  float sc[8], sh[8];
  load8(s_co + g * 8, sc);
  load8(s_co + 512 + g * 8, sh);
  const int rs = p.W * p.C;
  for (unsigned pix = blockIdx.x * ppc + threadIdx.x / groups; pix < npix; pix += gridDim.x * ppc) {
    const unsigned t = pix / Wo;
    const int wo = (int)(pix - t * Wo);
    const unsigned n = t / Ho;
    const int ho = (int)(t - n * Ho);
    const __nv_bfloat16* y0 =
        p.y + (((long long)n * p.H + 2 * ho - 1) * p.W + 2 * wo - 1) * p.C + g * 8;
    float best[8];
    uint32_t pay[8];
    if (ho > 0 && wo > 0) pool_window<true>(y0, rs, p.C, ho, wo, sc, sh, best, pay);
    else pool_window<false>(y0, rs, p.C, ho, wo, sc, sh, best, pay);
    const size_t o = (size_t)pix * p.C + g * 8;
    stg_v4(p.z + o, pack8(best));
    if (p.argmax != nullptr) {
      uint2 a;
      a.x = (pay[0] & 0xFFu) | ((pay[1] & 0xFFu) << 8) | ((pay[2] & 0xFFu) << 16) | (pay[3] << 24);
      a.y = (pay[4] & 0xFFu) | ((pay[5] & 0xFFu) << 8) | ((pay[6] & 0xFFu) << 16) | (pay[7] << 24);
      *reinterpret_cast<uint2*>(p.argmax + o) = a;
    }
    if (p.ysel != nullptr)   // bf16 in, bf16 out: exact
      stg_v4(p.ysel + o, make_uint4((pay[0] >> 16) | (pay[1] & 0xFFFF0000u),
                                    (pay[2] >> 16) | (pay[3] & 0xFFFF0000u),
                                    (pay[4] >> 16) | (pay[5] & 0xFFFF0000u),
                                    (pay[6] >> 16) | (pay[7] & 0xFFFF0000u)));
  }
  if (blockIdx.x == 0) bn_side_effects(p.bn, p.C);
}

int launch_bn_pool(const PoolParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 512 && kEwThreads % (p.C / 8) == 0, "bn_pool: unsupported C=%d", p.C);
  VPD_REQUIRE(p.H % 2 == 0 && p.W % 2 == 0, "bn_pool: odd spatial dims");
  if (p.N == 0) return 0;
  const long long total = (long long)p.N * (p.H / 2) * (p.W / 2) * (p.C / 8);
  VPD_REQUIRE(total < (1LL << 31), "bn_pool: tensor too large for 32-bit indexing");
  long long blocks = (total + kEwThreads - 1) / kEwThreads;
  if (blocks > 148 * 2) blocks = 148 * 2;  // one resident wave
  VPD_CHECK_CUDA(launch_kernel(bn_pool_kernel, dim3((int)blocks), dim3(kEwThreads), 0, s, p));
  VPD_LAUNCHED(1);
  return 0;
}

// plain maxpool 3x3/2 pad 1 (eval path: input already activated). All nine window vectors are
// requested before any is used (the first version walked the window in nested loops with
// early-outs, one dependent load at a time: 209 us per 1000 images at 3.1 TB/s), and the
// maximum is taken on the packed bf16 pairs (exact, no conversion).
VPD_DEVINL uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                   *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__global__ void __launch_bounds__(kEwThreads)
maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ z, int N, int H,
               int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int groups = C >> 3;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * groups;
  const uint32_t ninf = 0xFF80FF80u;   // (-inf, -inf)
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kEwThreads) {
    const int g = (int)(i % groups);
    long long pix = i / groups;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    uint4 win[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo - 1 + kw;
        if (h >= 0 && h < H && w >= 0 && w < W)
          win[kh * 3 + kw] = ldg_nc_v4(x + (((size_t)n * H + h) * W + w) * C + g * 8);
        else
          win[kh * 3 + kw] = make_uint4(ninf, ninf, ninf, ninf);
      }
    }
    uint4 best = win[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) {
      best.x = bf16x2_max(best.x, win[k].x);
      best.y = bf16x2_max(best.y, win[k].y);
      best.z = bf16x2_max(best.z, win[k].z);
      best.w = bf16x2_max(best.w, win[k].w);
    }
    stg_v4(z + (((size_t)n * Ho + ho) * Wo + wo) * C + g * 8, best);
  }
}

int launch_maxpool(const __nv_bfloat16* x, __nv_bfloat16* z, int N, int H, int W, int C,
                   cudaStream_t s) {
  VPD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: unsupported shape");
  if (N == 0) return 0;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  long long blocks = (total + kEwThreads - 1) / kEwThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  VPD_CHECK_CUDA(launch_kernel(maxpool_kernel, dim3((int)blocks), dim3(kEwThreads), 0, s, x, z, N, H, W, C));
  VPD_LAUNCHED(1);
  return 0;
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm,
                               const float* rv, float eps, float* scale, float* shift, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(rv[c] + eps);
  bn_affine(gamma[c], beta[c], rm[c], rstd, scale[c], shift[c]);
}

int launch_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                   float eps, float* scale, float* shift, int C, cudaStream_t s) {
  VPD_CHECK_CUDA(launch_kernel(bn_fold_kernel, dim3((C + 127) / 128), dim3(128), 0, s, gamma, beta, rm, rv, eps, scale, shift, C));
  VPD_LAUNCHED(1);
  return 0;
}

// ------------------------------------------------- standalone channel statistics
// sum / sum-of-squares per channel of a [M][C] bf16 tensor (fp64 atomics at the end).
// Used for the 64-channel layers, whose conv epilogue is the bottleneck: moving the
// reduction out of it is cheaper than the extra (L2-resident) read.
__global__ void __launch_bounds__(kEwThreads, 4)
channel_stats_kernel(const __nv_bfloat16* __restrict__ y, long long M, int C, StatAcc* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_a[kColScratch];
  __shared__ __align__(16) float s_b[kColScratch];
  const int groups = C >> 3;
  const int g = threadIdx.x % groups;
  const int r0 = threadIdx.x / groups;
  const int rstep = kEwThreads / groups;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  const long long chunk = (M + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * chunk;
  const long long end = beg + chunk < M ? beg + chunk : M;
  constexpr int U = 4;
  for (long long row = beg + r0; row < end; row += (long long)rstep * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r < end) v[u] = ldg_nc_v4(y + (size_t)r * C + g * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r >= end) break;
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += f[j];
        b[j] = fmaf(f[j], f[j], b[j]);
      }
    }
  }
  colsum_put(a, s_a, C, g, r0);
  colsum_put(b, s_b, C, g, r0);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kEwThreads) {
    stat_add(&stats[c], static_cast<double>(colsum_get(s_a, C, rstep, c)));
    stat_add(&stats[C + c], static_cast<double>(colsum_get(s_b, C, rstep, c)));
  }
}

int launch_channel_stats(const __nv_bfloat16* y, long long M, int C, StatAcc* stats, cudaStream_t s) {
  VPD_REQUIRE(C % 64 == 0 && C <= 512 && kEwThreads % (C / 8) == 0, "channel_stats: C=%d", C);
  if (M == 0) return 0;
  VPD_CHECK_CUDA(launch_kernel(channel_stats_kernel, dim3(ew_grid(M * (C / 8), 4, 4)),
                               dim3(kEwThreads), 0, s, y, M, C, stats));
  VPD_LAUNCHED(1);
  return 0;
}

// --------------------------------------------------------------- BN backward
// pass 1 (kApply = false): per-channel sum(g), sum(g * xhat_b), g = dz * 1[z > 0]
// pass 2 (kApply = true) : dy_b = A_b*g + B_b*y_b + C_b  (the usual
//          gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) with the constants folded),
//          masked gradient written back for the identity branch, dgamma/dbeta.
// Four rows per thread are loaded before any is used (memory-level parallelism).
template <bool kApply, int NB>
__global__ void __launch_bounds__(kEwThreads, NB == 1 ? 3 : 2) bn_bwd_kernel(const BnBwdParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_g[kApply ? 4 : kColScratch];
  __shared__ __align__(16) float s_gx[NB][kApply ? 4 : kColScratch];
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int r0 = threadIdx.x / groups;
  const int rstep = kEwThreads / groups;
  // reduce: c0 = mean, c1 = rstd.  apply: dy = c0*g + c1*y + c2 (derived once per CTA and
  // channel from the integer sums, then shared: see bn_stage_affine)
  float c0[NB][8], c1[NB][8], c2[NB][8];
  if (kApply) {
    __shared__ __align__(16) float s_co[kApply ? NB * 3 * 512 : 4];
    for (int i = threadIdx.x; i < NB * p.C; i += kEwThreads) {
      const int b = i / p.C, c = i - b * p.C;
      const float mean = __ldg(p.save_mean[b] + c), rstd = __ldg(p.save_rstd[b] + c);
      const float invM = 1.0f / static_cast<float>(p.M);
      const float k0 = __ldg(p.gamma[b] + c) * rstd;
      const float k1 = static_cast<float>(stat_read(p.sums[b] + c)) * invM;
      const float k2 = static_cast<float>(stat_read(p.sums[b] + p.C + c)) * invM;
      s_co[(b * 3 + 0) * 512 + c] = k0;
      s_co[(b * 3 + 1) * 512 + c] = -k0 * k2 * rstd;
      s_co[(b * 3 + 2) * 512 + c] = -k0 * k1 + k0 * k2 * rstd * mean;
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      load8(s_co + (b * 3 + 0) * 512 + g * 8, c0[b]);
      load8(s_co + (b * 3 + 1) * 512 + g * 8, c1[b]);
      load8(s_co + (b * 3 + 2) * 512 + g * 8, c2[b]);
    }
  } else {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        c0[b][j] = __ldg(p.save_mean[b] + c);
        c1[b][j] = __ldg(p.save_rstd[b] + c);
        c2[b][j] = 0.f;
      }
    }
  }
  float acc_g[8], acc_gx[NB][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc_g[j] = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) acc_gx[b][j] = 0.f;
  }

  const long long chunk = (p.M + gridDim.x - 1) / gridDim.x;
  const long long beg = (long long)blockIdx.x * chunk;
  const long long end = beg + chunk < p.M ? beg + chunk : p.M;
  constexpr int U = 2;
  for (long long row = beg + r0; row < end; row += (long long)rstep * U) {
    uint4 vd[U], vz[U], vy[NB][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r < end) {
        const size_t off = (size_t)r * p.C + g * 8;
        vd[u] = *reinterpret_cast<const uint4*>(p.dz + off);  // may alias dmask / dy
        if (p.z != nullptr) vz[u] = ldg_nc_v4(p.z + off);
#pragma unroll
        for (int b = 0; b < NB; ++b) vy[b][u] = ldg_nc_v4(p.y[b] + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = row + (long long)u * rstep;
      if (r >= end) break;
      const size_t off = (size_t)r * p.C + g * 8;
      float gr[8];
      unpack8(vd[u], gr);
      if (p.z != nullptr) {
        float zz[8];
        unpack8(vz[u], zz);
#pragma unroll
        for (int j = 0; j < 8; ++j) gr[j] = zz[j] > 0.f ? gr[j] : 0.f;
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float yy[8];
        unpack8(vy[b][u], yy);
        if (kApply) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(c0[b][j], gr[j], fmaf(c1[b][j], yy[j], c2[b][j]));
          stg_v4(p.dy[b] + off, pack8(o));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            acc_gx[b][j] = fmaf(gr[j], (yy[j] - c0[b][j]) * c1[b][j], acc_gx[b][j]);
        }
      }
      if (kApply) {
        if (p.dmask != nullptr) stg_v4(p.dmask + off, pack8(gr));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_g[j] += gr[j];
      }
    }
  }
  if (kApply) {
    if (blockIdx.x == 0) {
      for (int c = threadIdx.x; c < p.C; c += kEwThreads)
        for (int b = 0; b < NB; ++b) {
          p.dbeta[b][c] = static_cast<float>(stat_read(p.sums[b] + c));
          p.dgamma[b][c] = static_cast<float>(stat_read(p.sums[b] + p.C + c));
        }
    }
  } else {
    colsum_put(acc_g, s_g, p.C, g, r0);
#pragma unroll
    for (int b = 0; b < NB; ++b) colsum_put(acc_gx[b], s_gx[b], p.C, g, r0);
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
      const double sg = static_cast<double>(colsum_get(s_g, p.C, rstep, c));
      for (int b = 0; b < NB; ++b) {
        stat_add(&p.sums[b][c], sg);
        stat_add(&p.sums[b][p.C + c], static_cast<double>(colsum_get(s_gx[b], p.C, rstep, c)));
      }
    }
  }
}

int launch_bn_bwd(const BnBwdParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 512 && kEwThreads % (p.C / 8) == 0,
              "bn_bwd: unsupported channel count %d", p.C);
  VPD_REQUIRE(p.nbranch == 1 || p.nbranch == 2, "bn_bwd: nbranch");
  if (p.M == 0) return 0;
  const long long vectors = p.M * (p.C / 8);
  const int occ = p.nbranch == 1 ? 3 : 2;
  const int grid_r = ew_grid(vectors, 2, occ);
  const int grid_a = ew_grid(vectors, 2, occ);
  if (p.nbranch == 1) {
    if (!p.sums_ready)
      VPD_CHECK_CUDA(launch_kernel(bn_bwd_kernel<false, 1>, dim3(grid_r), dim3(kEwThreads), 0, s, p));
    VPD_CHECK_CUDA(launch_kernel(bn_bwd_kernel<true, 1>, dim3(grid_a), dim3(kEwThreads), 0, s, p));
  } else {
    if (!p.sums_ready)
      VPD_CHECK_CUDA(launch_kernel(bn_bwd_kernel<false, 2>, dim3(grid_r), dim3(kEwThreads), 0, s, p));
    VPD_CHECK_CUDA(launch_kernel(bn_bwd_kernel<true, 2>, dim3(grid_a), dim3(kEwThreads), 0, s, p));
  }
  VPD_LAUNCHED(p.sums_ready ? 1 : 2);
  return 0;
}

// ------------------------------------------- stem backward (pool + ReLU + BN)
// pass 1 runs at POOLED resolution: the gradient of the pre-pool activation is
// non-zero only at each window's argmax, so sum(g) and sum(g*xhat) are sums over
// pooled elements of dpool * 1[relu'(y_argmax)] (* xhat(y_argmax)).
__global__ void __launch_bounds__(kEwThreads) stem_bwd_reduce_kernel(const StemBwdParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_g[kColScratch];
  __shared__ __align__(16) float s_gx[kColScratch];
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int Ho = p.H / 2, Wo = p.W / 2;
  float mean[8], rstd[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    mean[j] = __ldg(p.save_mean + c);
    rstd[j] = __ldg(p.save_rstd + c);
    bn_affine(__ldg(p.gamma + c), __ldg(p.beta + c), mean[j], rstd[j], sc[j], sh[j]);
  }
  float acc_g[8], acc_gx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_g[j] = acc_gx[j] = 0.f;
  const long long total = (long long)p.N * Ho * Wo * groups;
  const long long stride = (long long)gridDim.x * kEwThreads;  // multiple of groups
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += stride) {
    long long pix = i / groups;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const size_t po = (((size_t)n * Ho + ho) * Wo + wo) * p.C + g * 8;
    const uint2 am = __ldg(reinterpret_cast<const uint2*>(p.argmax + po));
    float dp[8], ysel[8];
    unpack8(ldg_nc_v4(p.dpool + po), dp);
    uint4 win[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo - 1 + kw;
        if (h >= 0 && h < p.H && w >= 0 && w < p.W)
          win[kh * 3 + kw] = __ldg(reinterpret_cast<const uint4*>(
              p.y + (((size_t)n * p.H + h) * p.W + w) * p.C + g * 8));
        else
          win[kh * 3 + kw] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ysel[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float f[8];
      unpack8(win[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t word = j < 4 ? am.x : am.y;
        const int a = (word >> (8 * (j & 3))) & 0xFF;
        if (a == k) ysel[j] = f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gq = fmaf(ysel[j], sc[j], sh[j]) > 0.f ? dp[j] : 0.f;
      acc_g[j] += gq;
      acc_gx[j] = fmaf(gq, (ysel[j] - mean[j]) * rstd[j], acc_gx[j]);
    }
  }
  colsum_put(acc_g, s_g, p.C, g, threadIdx.x / groups);
  colsum_put(acc_gx, s_gx, p.C, g, threadIdx.x / groups);
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
    stat_add(&p.sums[c], static_cast<double>(colsum_get(s_g, p.C, kEwThreads / groups, c)));
    stat_add(&p.sums[p.C + c], static_cast<double>(colsum_get(s_gx, p.C, kEwThreads / groups, c)));
  }
}

// pass 1 when the forward kept the pre-BN value at every window's argmax (PoolParams::ysel):
// two pooled tensors (dpool, ysel: 2 x 33 MB at batch 256) instead of gathering the 3x3
// window of every pooled element from y again (134 MB + argmax).
__global__ void __launch_bounds__(kEwThreads, 3) stem_bwd_reduce_sel_kernel(const StemBwdParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_g[kColScratch];
  __shared__ __align__(16) float s_gx[kColScratch];
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  float mean[8], rstd[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    mean[j] = __ldg(p.save_mean + c);
    rstd[j] = __ldg(p.save_rstd + c);
    bn_affine(__ldg(p.gamma + c), __ldg(p.beta + c), mean[j], rstd[j], sc[j], sh[j]);
  }
  float acc_g[8], acc_gx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_g[j] = acc_gx[j] = 0.f;
  const long long total = (long long)p.N * (p.H / 2) * (p.W / 2) * groups;   // 16-byte vectors
  const long long stride = (long long)gridDim.x * kEwThreads;                // multiple of groups
  constexpr int U = 2;
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += stride * U) {
    uint4 vd[U], vs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = i + u * stride;
      if (k < total) {
        vd[u] = ldg_nc_v4(p.dpool + k * 8);
        vs[u] = ldg_nc_v4(p.ysel + k * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= total) break;
      float dp[8], ys[8];
      unpack8(vd[u], dp);
      unpack8(vs[u], ys);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gq = fmaf(ys[j], sc[j], sh[j]) > 0.f ? dp[j] : 0.f;
        acc_g[j] += gq;
        acc_gx[j] = fmaf(gq, (ys[j] - mean[j]) * rstd[j], acc_gx[j]);
      }
    }
  }
  colsum_put(acc_g, s_g, p.C, g, threadIdx.x / groups);
  colsum_put(acc_gx, s_gx, p.C, g, threadIdx.x / groups);
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
    stat_add(&p.sums[c], static_cast<double>(colsum_get(s_g, p.C, kEwThreads / groups, c)));
    stat_add(&p.sums[p.C + c], static_cast<double>(colsum_get(s_gx, p.C, kEwThreads / groups, c)));
  }
}

// pass 2 at input resolution. A thread owns a 2x2 pixel quad (rows 2a, 2a+1; columns 2b,
// 2b+1) of one 8-channel group: the only pooling windows that contain any of its pixels are
// (a | a+1, b | b+1), so four dpool / argmax vectors serve four outputs (a per-pixel
// gather needs up to four per output).
__global__ void __launch_bounds__(kEwThreads, 2) stem_bwd_apply_kernel(const StemBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  const int Ho = p.H / 2, Wo = p.W / 2;
  const long long M = (long long)p.N * p.H * p.W;
  float sc[8], sh[8], a0[8], a1[8], a2[8];
  __shared__ __align__(16) float s_co[5 * 512];
  for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
    const float mean = __ldg(p.save_mean + c), rstd = __ldg(p.save_rstd + c);
    const float gamma = __ldg(p.gamma + c);
    float scc, shc;
    bn_affine(gamma, __ldg(p.beta + c), mean, rstd, scc, shc);
    const float invM = 1.0f / static_cast<float>(M);
    const float k0 = gamma * rstd;
    const float k1 = static_cast<float>(stat_read(p.sums + c)) * invM;
    const float k2 = static_cast<float>(stat_read(p.sums + p.C + c)) * invM;
    s_co[c] = scc;
    s_co[512 + c] = shc;
    s_co[2 * 512 + c] = k0;
    s_co[3 * 512 + c] = -k0 * k2 * rstd;
    s_co[4 * 512 + c] = -k0 * k1 + k0 * k2 * rstd * mean;
  }
  __syncthreads();
  load8(s_co + g * 8, sc);
  load8(s_co + 512 + g * 8, sh);
  load8(s_co + 2 * 512 + g * 8, a0);
  load8(s_co + 3 * 512 + g * 8, a1);
  load8(s_co + 4 * 512 + g * 8, a2);
  const long long total = (long long)p.N * Ho * Wo * groups;   // quads x channel groups
  const long long stride = (long long)gridDim.x * kEwThreads;
  for (long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += stride) {
    long long q = i / groups;
    const int b = (int)(q % Wo);
    q /= Wo;
    const int a = (int)(q % Ho);
    const int n = (int)(q / Ho);
    // the four windows (a + da, b + db); the second ones may fall off the edge
    const bool has_a1 = a + 1 < Ho, has_b1 = b + 1 < Wo;
    uint4 vdp[4];
    uint2 vam[4];
#pragma unroll
    for (int da = 0; da < 2; ++da)
#pragma unroll
      for (int db = 0; db < 2; ++db) {
        const int k = da * 2 + db;
        if ((da == 0 || has_a1) && (db == 0 || has_b1)) {
          const size_t po = (((size_t)n * Ho + a + da) * Wo + b + db) * p.C + g * 8;
          vam[k] = __ldg(reinterpret_cast<const uint2*>(p.argmax + po));
          vdp[k] = ldg_nc_v4(p.dpool + po);
        } else {
          vam[k] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);   // matches no window position
          vdp[k] = make_uint4(0, 0, 0, 0);
        }
      }
    uint4 vy[4];
    size_t off[4];
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        off[dh * 2 + dw] = (((size_t)n * p.H + 2 * a + dh) * p.W + 2 * b + dw) * p.C + g * 8;
        vy[dh * 2 + dw] = ldg_nc_v4(p.y + off[dh * 2 + dw]);
      }
    float dp[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8(vdp[k], dp[k]);
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        // pixel (2a+dh, 2b+dw) sits at window position (kh, kw) = (2a+dh - (2(a+da)-1), ...)
        //   = (dh + 1 - 2da, dw + 1 - 2db); valid when both are in 0..2
        float yy[8], gr[8];
        unpack8(vy[dh * 2 + dw], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) gr[j] = 0.f;
#pragma unroll
        for (int da = 0; da < 2; ++da)
#pragma unroll
          for (int db = 0; db < 2; ++db) {
            const int kh = dh + 1 - 2 * da, kw = dw + 1 - 2 * db;
            if (kh < 0 || kw < 0) continue;   // compile-time after unrolling
            const int k = da * 2 + db;
            const int kidx = kh * 3 + kw;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t word = j < 4 ? vam[k].x : vam[k].y;
              const int am = (word >> (8 * (j & 3))) & 0xFF;
              if (am == kidx) gr[j] += dp[k][j];
            }
          }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gq = fmaf(yy[j], sc[j], sh[j]) > 0.f ? gr[j] : 0.f;
          o[j] = fmaf(a0[j], gq, fmaf(a1[j], yy[j], a2[j]));
        }
        stg_cs_v4(p.dy + off[dh * 2 + dw], pack8(o));
      }
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
      p.dbeta[c] = static_cast<float>(stat_read(p.sums + c));
      p.dgamma[c] = static_cast<float>(stat_read(p.sums + p.C + c));
    }
}

int launch_stem_bwd(const StemBwdParams& p, cudaStream_t s) {
  VPD_REQUIRE(p.C % 64 == 0 && p.C <= 512 && kEwThreads % (p.C / 8) == 0, "stem_bwd: C=%d", p.C);
  if (p.N == 0) return 0;
  const long long pooled = (long long)p.N * (p.H / 2) * (p.W / 2) * (p.C / 8);
  if (p.ysel != nullptr)
    VPD_CHECK_CUDA(launch_kernel(stem_bwd_reduce_sel_kernel, dim3(ew_grid(pooled, 2, 3)), dim3(kEwThreads), 0, s, p));
  else
    VPD_CHECK_CUDA(launch_kernel(stem_bwd_reduce_kernel, dim3(ew_grid(pooled, 2, 2)), dim3(kEwThreads), 0, s, p));
  VPD_CHECK_CUDA(launch_kernel(stem_bwd_apply_kernel, dim3(ew_grid(pooled, 2, 2)), dim3(kEwThreads), 0, s, p));
  VPD_LAUNCHED(2);
  return 0;
}

}  // namespace vpd
