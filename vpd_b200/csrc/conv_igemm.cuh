// K2: implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves the forward convolutions
// (torchvision BasicBlock convs called from models/rgb.py:68-70)
// and their data gradients (A9). GEMM view, per output tile:
//     D[128 pixels, BLOCK_N channels] = sum over taps t, 64-channel chunks kc of
//         A_t,kc[128 pixels, 64]  x  B_t,kc[BLOCK_N, 64]^T
//   * activations live in HBM as NHWC bf16; the A tile of tap t is ONE 5-D TMA
//     box {64 ch, tw, 1, th, tn} fetched at the tap's spatial offset from a
//     tensor-map *view* of the same buffer (stride-1: (C,W,1,H,N); stride-2:
//     (2C,W/2,2,H/2,N) so that even/odd pixels become separate coordinates; the
//     7x7 stem uses overlapping 8-pixel windows). Out-of-range coordinates
//     are zero-filled by TMA, which implements the conv padding.
//   * weights are bf16 [tap][Cout][Cin] (K-major), box {64, BLOCK_N, 1}.
//   * both land in shared memory 128B-swizzled, K-major; a single thread
//     issues tcgen05.mma (M=128, N=BLOCK_N, K=16) accumulating in TMEM
//     (fp32). Two TMEM accumulator stages let the epilogue of tile i overlap
//     the MMAs of tile i+1.
//   * epilogue warps read TMEM (tcgen05.ld), optionally apply a per-channel
//     affine (folded eval-mode BN) + residual + ReLU, write bf16 NHWC, and in
//     training mode accumulate the per-channel sum / sum-of-squares that
//     BatchNorm needs (warp transpose-reduce -> smem -> fp64 global atomics).
#pragma once
#include "common.cuh"

namespace vpd {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 12;
constexpr int kEpiWarps = 4;                       // epilogue warps (one per TMEM lane quarter; 8 was slower)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreads = 64 + kEpiThreads;     // TMA warp + MMA warp + epilogue warps
constexpr int kWgradThreads = 192;

struct ConvTap {
  int c0;       // offset added to the innermost (channel) coordinate
  int d1, d2, d3;  // offsets for the w, parity and h coordinates
  int src;      // which (A,B) tensor-map pair (0/1)
  int btap;     // index on the tap axis of the weight tensor
  int kchunks;  // number of 64-wide K chunks for this tap
  int pad_;
};

struct ConvParams {
  int tw, th, tn;  // tile = tw x th pixels x tn images, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;     // Cout / BLOCK_N
  int num_taps;
  ConvTap taps[kMaxTaps];
  int batch, out_h, out_w;  // valid extents of the (n, h, w) tile coordinates
  int cout;
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;      // same addressing as out, or null
  long long out_sn, out_sh, out_sw;   // element strides of out/residual
  const float* scale;                 // [cout] or null
  const float* shift;                 // [cout] or null
  int relu;
  double* stats;                      // [2][cout] (sum, sumsq) or null
  // pre-tiled bf16 weights (see wtile_offset) for tap.src 0 / 1: element offset of the
  // [BLOCK_N x 64] tile (tap t, chunk kc, rows n0..) = ((t*w_kc + kc)*w_rb + n0/64) * 4096
  const __nv_bfloat16* w[2];
  int w_kc[2];                        // K chunks per tap
  int w_rb[2];                        // 64-row blocks (N dimension / 64)
  // Fused BatchNorm-backward reduction (dgrad launches): the tile just computed is
  // dz of a ReLU->BN stage; mask it with 1[z > 0] before it is stored (so `out`
  // holds g) and accumulate sum(g), sum(g * xhat_b) for up to two BN branches.
  int bnb;                            // 0 = off, else number of branches (1 or 2)
  const __nv_bfloat16* bz;            // post-ReLU output of that stage (same addressing as out)
  const __nv_bfloat16* by[2];         // its pre-BN conv outputs
  const float* bmean[2];              // [cout] saved batch mean
  const float* brstd[2];              // [cout] saved 1/sqrt(var+eps)
  double* bsums[2];                   // [2][cout]: sum g, sum g*xhat
  // debug (vpd_conv_trace): 8 int64 per CTA - globaltimer at entry, then clock64 at
  // entry / after the dependency wait / first operands landed / last MMA issued /
  // first accumulator ready / epilogue done / exit
  long long* trace;
};
VPD_DEVINL void trace_mark(const ConvParams& p, int slot) {
  if (p.trace != nullptr) p.trace[blockIdx.x * 8 + slot] = clock64();
}

// called by the 128 epilogue threads (threadIdx.x in [64,192))
template <int BLOCK_N>
VPD_DEVINL void flush_channel_sums(const ConvParams& p, int ntile, float* s_sum, float* s_sq,
                                   float* s_x2, bool clear) {
  for (int i = threadIdx.x - 64; i < BLOCK_N; i += kEpiThreads) {
    const int c = ntile * BLOCK_N + i;
    if (p.stats != nullptr) {
      atomicAdd(&p.stats[c], static_cast<double>(s_sum[i]));
      atomicAdd(&p.stats[p.cout + c], static_cast<double>(s_sq[i]));
    } else {
      atomicAdd(&p.bsums[0][c], static_cast<double>(s_sum[i]));
      atomicAdd(&p.bsums[0][p.cout + c], static_cast<double>(s_sq[i]));
      if (p.bnb > 1) {
        atomicAdd(&p.bsums[1][c], static_cast<double>(s_sum[i]));
        atomicAdd(&p.bsums[1][p.cout + c], static_cast<double>(s_x2[i]));
      }
    }
    if (clear) {
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
      s_x2[i] = 0.f;
    }
  }
}

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 64 ? 6 : (BLOCK_N == 128 ? 5 : 4);
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 7 * BLOCK_N * 4 + 1024;
};

// Warp transpose-reduce of two 32-wide per-lane vectors: afterwards lane j holds in
// a[0] / b[0] the totals of column j over the warp's 32 rows (31 shuffles each).
VPD_DEVINL void warp_colsum2(float (&a)[32], float (&b)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float ka = up ? a[i + o] : a[i];
      const float sa = up ? a[i] : a[i + o];
      a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, o);
      const float kb = up ? b[i + o] : b[i];
      const float sb = up ? b[i] : b[i + o];
      b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, o);
    }
  }
}
struct ConvParams;
template <int BLOCK_N>
VPD_DEVINL void flush_channel_sums(const ConvParams& p, int ntile, float* s_sum, float* s_sq,
                                   float* s_x2, bool clear);

VPD_DEVINL void warp_colsum1(float (&a)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float ka = up ? a[i + o] : a[i];
      const float sa = up ? a[i] : a[i + o];
      a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, o);
    }
  }
}

// Epilogue shared by the implicit-GEMM kernels: executed by warps 2..5 (threads
// 64..191); drains the TMEM accumulator stages tile by tile.
template <int BLOCK_N, int CS>
VPD_DEVINL void conv_epilogue(const ConvParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                              uint64_t* tempty_bar, float* s_sum, float* s_sq, float* s_x2,
                              float* s_bn, int rank, int first_item, int item_stride,
                              int total_tiles, int warp, int lane) {
  // ---------------------------------------------------------------- epilogue
  // pair mode (CS == 2): the leader's MMA thread owns the accumulator hand-shake, so the
  // peer's epilogue warps release the TMEM stage on the LEADER's barrier
  const uint32_t tempty_remote0 =
      (CS == 2 && rank == 1) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
  const int q = warp & 3;       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;  // row of the 128-row tile
  int as = 0;
  uint32_t aphase = 0;
  int cur_ntile = -1;
  for (int tile = first_item; tile < total_tiles; tile += item_stride) {
    const int n_tile = tile % p.n_tiles;
    int mt = (tile / p.n_tiles) * CS + rank;
    const int w0 = (mt % p.tiles_w) * p.tw;
    mt /= p.tiles_w;
    const int h0 = (mt % p.tiles_h) * p.th;
    const int b0 = (mt / p.tiles_h) * p.tn;
    const int w = w0 + r % p.tw;
    const int h = h0 + (r / p.tw) % p.th;
    const int n = b0 + r / (p.tw * p.th);
    const bool valid = (n < p.batch) && (h < p.out_h) && (w < p.out_w);
    const long long off = n * p.out_sn + h * p.out_sh + w * p.out_sw + n_tile * BLOCK_N;

    if ((p.stats != nullptr || p.bnb > 0) && cur_ntile != n_tile) {
      // flush per-CTA channel sums when the channel block changes
      // (named barrier over the 4 epilogue warps only)
      if (cur_ntile >= 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        flush_channel_sums<BLOCK_N>(p, cur_ntile, s_sum, s_sq, s_x2, true);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      cur_ntile = n_tile;
      if (p.bnb > 0) {
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += kEpiThreads) {
          const int ch = n_tile * BLOCK_N + i;
          s_bn[i] = __ldg(p.bmean[0] + ch);
          s_bn[BLOCK_N + i] = __ldg(p.brstd[0] + ch);
          if (p.bnb > 1) {
            s_bn[2 * BLOCK_N + i] = __ldg(p.bmean[1] + ch);
            s_bn[3 * BLOCK_N + i] = __ldg(p.brstd[1] + ch);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
    }

    mbar_wait(&tfull_bar[as], aphase);
    if (tile == first_item && threadIdx.x == 64) trace_mark(p, 5);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c * 32, v);
      // issue every global load of this chunk before waiting on anything, so the
      // TMEM read and the (up to four) 64-byte row segments are all in flight together
      const bool do_res = p.residual != nullptr && valid;
      const bool do_bn = p.bnb > 0 && valid;
      uint4 rres[4], rz[4], ry0[4], ry1[4];
      if (do_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) rres[j] = rp[j];  // plain load: residual may alias out
      }
      if (do_bn) {
        const uint4* zp = reinterpret_cast<const uint4*>(p.bz + off + c * 32);
        const uint4* yp = reinterpret_cast<const uint4*>(p.by[0] + off + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) rz[j] = __ldg(zp + j);
#pragma unroll
        for (int j = 0; j < 4; ++j) ry0[j] = __ldg(yp + j);
        if (p.bnb > 1) {
          const uint4* y1p = reinterpret_cast<const uint4*>(p.by[1] + off + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) ry1[j] = __ldg(y1p + j);
        }
      }
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const int ch0 = n_tile * BLOCK_N + c * 32;
      if (p.scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + ch0 + j));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + ch0 + j));
          f[j + 0] = fmaf(f[j + 0], sc.x, sh.x);
          f[j + 1] = fmaf(f[j + 1], sc.y, sh.y);
          f[j + 2] = fmaf(f[j + 2], sc.z, sh.z);
          f[j + 3] = fmaf(f[j + 3], sc.w, sh.w);
        }
      }
      if (do_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[8 * j + 0] += bf16_lo(rres[j].x);
          f[8 * j + 1] += bf16_hi(rres[j].x);
          f[8 * j + 2] += bf16_lo(rres[j].y);
          f[8 * j + 3] += bf16_hi(rres[j].y);
          f[8 * j + 4] += bf16_lo(rres[j].z);
          f[8 * j + 5] += bf16_hi(rres[j].z);
          f[8 * j + 6] += bf16_lo(rres[j].w);
          f[8 * j + 7] += bf16_hi(rres[j].w);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (do_bn) {  // g = dz * 1[z > 0]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[8 * j + 0] = bf16_lo(rz[j].x) > 0.f ? f[8 * j + 0] : 0.f;
          f[8 * j + 1] = bf16_hi(rz[j].x) > 0.f ? f[8 * j + 1] : 0.f;
          f[8 * j + 2] = bf16_lo(rz[j].y) > 0.f ? f[8 * j + 2] : 0.f;
          f[8 * j + 3] = bf16_hi(rz[j].y) > 0.f ? f[8 * j + 3] : 0.f;
          f[8 * j + 4] = bf16_lo(rz[j].z) > 0.f ? f[8 * j + 4] : 0.f;
          f[8 * j + 5] = bf16_hi(rz[j].z) > 0.f ? f[8 * j + 5] : 0.f;
          f[8 * j + 6] = bf16_lo(rz[j].w) > 0.f ? f[8 * j + 6] : 0.f;
          f[8 * j + 7] = bf16_hi(rz[j].w) > 0.f ? f[8 * j + 7] : 0.f;
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
      if (valid) {
        uint4* op = reinterpret_cast<uint4*>(p.out + off + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg_v4(op + j, make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
      }
      if (p.stats != nullptr) {
        // statistics of the values as stored (bf16-rounded); invalid rows = 0
        float s[32], s2[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = valid ? bf16_lo(pk[j]) : 0.f;
          const float b = valid ? bf16_hi(pk[j]) : 0.f;
          s[2 * j] = a;
          s[2 * j + 1] = b;
          s2[2 * j] = a * a;
          s2[2 * j + 1] = b * b;
        }
        warp_colsum2(s, s2, lane);
        red_shared_add(smem_u32(s_sum + c * 32 + lane), s[0]);
        red_shared_add(smem_u32(s_sq + c * 32 + lane), s2[0]);
      }
      if (p.bnb > 0) {
        // sum g and sum g*xhat of the stored (bf16-rounded) masked gradient
        float g[32], gx[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          g[2 * j] = valid ? bf16_lo(pk[j]) : 0.f;
          g[2 * j + 1] = valid ? bf16_hi(pk[j]) : 0.f;
        }
        const uint32_t sbn = smem_u32(s_bn + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 m0 = __uint4_as_float4(lds_v4(sbn + 32 * j));
          const float4 m1 = __uint4_as_float4(lds_v4(sbn + 32 * j + 16));
          const float4 r0 = __uint4_as_float4(lds_v4(sbn + 4 * BLOCK_N + 32 * j));
          const float4 r1 = __uint4_as_float4(lds_v4(sbn + 4 * BLOCK_N + 32 * j + 16));
          gx[8 * j + 0] = do_bn ? g[8 * j + 0] * ((bf16_lo(ry0[j].x) - m0.x) * r0.x) : 0.f;
          gx[8 * j + 1] = do_bn ? g[8 * j + 1] * ((bf16_hi(ry0[j].x) - m0.y) * r0.y) : 0.f;
          gx[8 * j + 2] = do_bn ? g[8 * j + 2] * ((bf16_lo(ry0[j].y) - m0.z) * r0.z) : 0.f;
          gx[8 * j + 3] = do_bn ? g[8 * j + 3] * ((bf16_hi(ry0[j].y) - m0.w) * r0.w) : 0.f;
          gx[8 * j + 4] = do_bn ? g[8 * j + 4] * ((bf16_lo(ry0[j].z) - m1.x) * r1.x) : 0.f;
          gx[8 * j + 5] = do_bn ? g[8 * j + 5] * ((bf16_hi(ry0[j].z) - m1.y) * r1.y) : 0.f;
          gx[8 * j + 6] = do_bn ? g[8 * j + 6] * ((bf16_lo(ry0[j].w) - m1.z) * r1.z) : 0.f;
          gx[8 * j + 7] = do_bn ? g[8 * j + 7] * ((bf16_hi(ry0[j].w) - m1.w) * r1.w) : 0.f;
        }
        if (p.bnb > 1) {
          float gx1[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 m0 = __uint4_as_float4(lds_v4(sbn + 8 * BLOCK_N + 32 * j));
            const float4 m1 = __uint4_as_float4(lds_v4(sbn + 8 * BLOCK_N + 32 * j + 16));
            const float4 r0 = __uint4_as_float4(lds_v4(sbn + 12 * BLOCK_N + 32 * j));
            const float4 r1 = __uint4_as_float4(lds_v4(sbn + 12 * BLOCK_N + 32 * j + 16));
            gx1[8 * j + 0] = do_bn ? g[8 * j + 0] * ((bf16_lo(ry1[j].x) - m0.x) * r0.x) : 0.f;
            gx1[8 * j + 1] = do_bn ? g[8 * j + 1] * ((bf16_hi(ry1[j].x) - m0.y) * r0.y) : 0.f;
            gx1[8 * j + 2] = do_bn ? g[8 * j + 2] * ((bf16_lo(ry1[j].y) - m0.z) * r0.z) : 0.f;
            gx1[8 * j + 3] = do_bn ? g[8 * j + 3] * ((bf16_hi(ry1[j].y) - m0.w) * r0.w) : 0.f;
            gx1[8 * j + 4] = do_bn ? g[8 * j + 4] * ((bf16_lo(ry1[j].z) - m1.x) * r1.x) : 0.f;
            gx1[8 * j + 5] = do_bn ? g[8 * j + 5] * ((bf16_hi(ry1[j].z) - m1.y) * r1.y) : 0.f;
            gx1[8 * j + 6] = do_bn ? g[8 * j + 6] * ((bf16_lo(ry1[j].w) - m1.z) * r1.z) : 0.f;
            gx1[8 * j + 7] = do_bn ? g[8 * j + 7] * ((bf16_hi(ry1[j].w) - m1.w) * r1.w) : 0.f;
          }
          warp_colsum1(gx1, lane);
          red_shared_add(smem_u32(s_x2 + c * 32 + lane), gx1[0]);
        }
        warp_colsum2(g, gx, lane);
        red_shared_add(smem_u32(s_sum + c * 32 + lane), g[0]);
        red_shared_add(smem_u32(s_sq + c * 32 + lane), gx[0]);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CS == 2 && rank == 1) mbar_arrive_remote(tempty_remote0 + as * 8);
      else mbar_arrive(&tempty_bar[as]);
    }
    if (++as == 2) {
      as = 0;
      aphase ^= 1;
    }
  }
  if ((p.stats != nullptr || p.bnb > 0) && cur_ntile >= 0) {
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    flush_channel_sums<BLOCK_N>(p, cur_ntile, s_sum, s_sq, s_x2, false);
  }
}

// CS = 2: PAIR mode (tcgen05 cta_group::2). The two CTAs of a cluster own two adjacent
// pixel tiles of the same channel block and act as ONE 256 x BLOCK_N MMA unit: each loads
// its own A tile and HALF of the weight tile; the leader issues the MMAs, which read A
// from both CTAs' shared memory and the weight halves from both, and write each CTA's
// 128 rows to its own TMEM. A 128x128 single-CTA tile needs 128 B/cycle of operand reads
// plus as much TMA write traffic - twice the 128 B/cycle the shared memory can move -
// which is what pins the single-CTA kernel near half of the tensor peak; in pair mode a
// CTA reads/writes 3/4 (N=128) or 1/2 (N=256) as many bytes per MAC.
template <int BLOCK_N, int CS>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* pfull_bar = tempty_bar + 2;  // leader only: "the peer's stage has landed"
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pfull_bar + Cfg::kStages);
  float* s_sum = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;  // second BN branch (fused backward reduction)
  float* s_bn = s_x2 + BLOCK_N;  // [mean0 | rstd0 | mean1 | rstd1] of the current channel block

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if (CS == 2) mbar_init(&pfull_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CS * kEpiWarps);  // pair mode: both CTAs' epilogue warps
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
  }
  if (warp == 1) {
    if (CS == 2) {
      tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  for (int i = threadIdx.x; i < 3 * BLOCK_N; i += kConvThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // peers' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0 && p.trace != nullptr) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.trace[blockIdx.x * 8] = static_cast<long long>(g);
    trace_mark(p, 1);
  }
  pdl_wait();  // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) trace_mark(p, 2);

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  // work items are (group of CS pixel tiles, channel block); a CTA takes the pixel
  // tile `group*CS + rank` (possibly past the end: loads zero-fill, stores masked)
  const int rank = CS > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int total_tiles = ((m_tiles + CS - 1) / CS) * p.n_tiles;
  const int first_item = blockIdx.x / CS;
  const int item_stride = gridDim.x / CS;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        const int n_tile = tile % p.n_tiles;
        int mt = (tile / p.n_tiles) * CS + rank;
        const int w0 = (mt % p.tiles_w) * p.tw;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.th;
        const int b0 = (mt / p.tiles_h) * p.tn;
        for (int t = 0; t < p.num_taps; ++t) {
          const ConvTap tap = p.taps[t];
          const CUtensorMap* ma = tap.src ? &tmA1 : &tmA0;
          const CUtensorMap* mb = tap.src ? &tmB1 : &tmB0;
          for (int kc = 0; kc < tap.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes / CS);
            tma_load_5d(sa, ma, &full_bar[stage], tap.c0 + kc * kBlockK, w0 + tap.d1, tap.d2,
                        h0 + tap.d3, b0);
            // this CTA's share of the weight tile: BLOCK_N / CS rows
            bulk_load(sb,
                      p.w[tap.src] + ((size_t)(tap.btap * p.w_kc[tap.src] + kc) * p.w_rb[tap.src] +
                                      n_tile * (BLOCK_N / 64) + rank * (BLOCK_N / 64 / CS)) * 4096,
                      Cfg::kBBytes / CS, &full_bar[stage]);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    const bool issuer = elect_one();  // executed by the whole (converged) warp
    if (issuer && (CS == 1 || rank == 0)) {
      constexpr uint32_t idesc = make_idesc_bf16(CS * kBlockM, BLOCK_N, 0, 0);
      int total_kb = 0;
      for (int t = 0; t < p.num_taps; ++t) total_kb += p.taps[t].kchunks;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (CS == 2) mbar_wait(&pfull_bar[stage], phase);
          if (kb == 0 && tile == first_item) trace_mark(p, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 B (16 bf16) along K inside the 128 B swizzled row
            if (CS == 2) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (CS == 2) umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs
          else umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CS == 2) umma_commit_pair(&tfull_bar[as]);
        else umma_commit(&tfull_bar[as]);
        trace_mark(p, 4);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    } else if (issuer && CS == 2) {
      // peer CTA: relay "my stage has landed" to the leader's MMA thread
      const uint32_t pfull_remote0 = mapa_shared(smem_u32(&pfull_bar[0]), 0);
      int total_kb = 0;
      for (int t = 0; t < p.num_taps; ++t) total_kb += p.taps[t].kchunks;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          mbar_arrive_remote(pfull_remote0 + stage * 8);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    conv_epilogue<BLOCK_N, CS>(p, tmem_base, tfull_bar, tempty_bar, s_sum, s_sq, s_x2, s_bn, rank,
                               first_item, item_stride, total_tiles, warp, lane);
  }

  if (threadIdx.x == 64) trace_mark(p, 6);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p, 7);
  if (CS > 1) cluster_sync_all();  // no CTA exits while a peer may still signal / multicast to it
  tc_fence_after();
  if (warp == 1) {
    if (CS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ===========================================================================
// K2h: 3x3 stride-1 convolution (forward or dgrad) with HALO REUSE.
//
// The generic kernel above fetches a separate 128-pixel A tile per tap, i.e. it pulls
// every activation nine times through the L2->SM path, which is what bounds it
// (~40 B/cycle/SM measured). Here a pixel tile is 8 wide x 16 high, and ONE TMA box
// {64 ch, 10, 1, 18, 1} brings the tile plus its 1-pixel halo (180 rows of 128 B)
// into shared memory. Each of the nine taps is then just a different UMMA
// descriptor into that patch: start row (1+dy)*10 + (1+dx), 8-row groups 10 rows
// (1280 B) apart - the tensor core's 128B-swizzle is a function of absolute
// shared-memory address bits (probed in tests/diag_umma_probe.py), so any start
// row / group stride that is a multiple of 128 B reads back exactly what TMA wrote.
// The weights of the CTA's channel block (9 taps x CHUNKS x 64 x 64 bf16) are
// loaded once and stay resident. L2->SM traffic per tile drops from
// 9*CHUNKS*(16+8) KB to CHUNKS*22.5 KB.
// ===========================================================================
template <int CHUNKS>
struct HaloCfg {
  static constexpr int kBlockN = 64;
  static constexpr int kPatchRows = 18 * 10;
  static constexpr int kPatchBytes = kPatchRows * 128;               // 23040
  static constexpr int kPatchSlot = (kPatchBytes + 1023) & ~1023;    // 23552
  static constexpr int kWBytes = 9 * CHUNKS * kBlockN * 128;         // resident weights
  static constexpr int kSlots = CHUNKS == 1 ? 5 : 3;                 // patch ring (per chunk)
  static constexpr int kTmemCols = 2 * kBlockN;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kWBytes + kSlots * kPatchSlot + kBarBytes + 7 * kBlockN * 4 + 1024;
};

template <int CHUNKS>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ ConvParams p) {
  using Cfg = HaloCfg<CHUNKS>;
  constexpr int BLOCK_N = Cfg::kBlockN;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                         // [chunk][tap][64 cout][64 cin] bf16, swizzled
  uint8_t* s_patch = smem + Cfg::kWBytes;      // kSlots x kPatchSlot
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_patch + Cfg::kSlots * Cfg::kPatchSlot);
  uint64_t* empty_bar = full_bar + Cfg::kSlots;
  uint64_t* tfull_bar = empty_bar + Cfg::kSlots;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_sum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;
  float* s_bn = s_x2 + BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * BLOCK_N; i += kConvThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles;
  // the channel block of a CTA is fixed (host: gridDim.x % n_tiles == 0)
  const int my_ntile = blockIdx.x % p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      // resident weights of this CTA's channel block: CHUNKS x 9 boxes {64 cin, 64 cout, 1}
      mbar_expect_tx(w_bar, Cfg::kWBytes);
      for (int kc = 0; kc < CHUNKS; ++kc)
        for (int t = 0; t < 9; ++t)
          bulk_load(s_w + (kc * 9 + t) * (BLOCK_N * 128),
                    p.w[0] + ((size_t)(p.taps[t].btap * p.w_kc[0] + kc) * p.w_rb[0] + my_ntile) * 4096,
                    BLOCK_N * 128, w_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.tw;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.th;
        const int b0 = mt / p.tiles_h;
        for (int kc = 0; kc < CHUNKS; ++kc) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          mbar_expect_tx(&full_bar[slot], Cfg::kPatchBytes);
          tma_load_5d(s_patch + slot * Cfg::kPatchSlot, &tmA, &full_bar[slot], kc * 64, w0 - 1, 0,
                      h0 - 1, b0);
          if (++slot == Cfg::kSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      mbar_wait(w_bar, 0);
      int slot = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kc = 0; kc < CHUNKS; ++kc) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint32_t patch = smem_u32(s_patch + slot * Cfg::kPatchSlot);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            // tap (dy, dx): patch rows start at (1+dy)*10 + (1+dx); tile row h -> +10 rows
            const int start_row = (1 + p.taps[t].d3) * 10 + (1 + p.taps[t].d1);
            const uint64_t adesc = make_smem_desc(patch + start_row * 128, 16, 1280);
            const uint64_t bdesc =
                make_smem_desc(smem_u32(s_w + (kc * 9 + t) * (BLOCK_N * 128)), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | t | k) != 0);
          }
          umma_commit(&empty_bar[slot]);
          if (++slot == Cfg::kSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    conv_epilogue<BLOCK_N, 1>(p, tmem_base, tfull_bar, tempty_bar, s_sum, s_sq, s_x2, s_bn, 0,
                              blockIdx.x, gridDim.x, total_tiles, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------
// Halo-reuse variant with STREAMED weights (channel blocks too large to keep
// resident, e.g. 128 -> 128): one patch per 64-channel chunk as above, and the
// (chunk, tap) weight tiles [BLOCK_N][64] flow through their own TMA ring.
// ---------------------------------------------------------------------------
template <int BLOCK_N>
struct HaloStreamCfg {
  static constexpr int kPatchBytes = 18 * 10 * 128;
  static constexpr int kPatchSlot = (kPatchBytes + 1023) & ~1023;
  static constexpr int kSlots = 3;
  static constexpr int kWTile = BLOCK_N * 128;
  static constexpr int kWStages = 5;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes =
      kWStages * kWTile + kSlots * kPatchSlot + kBarBytes + 7 * BLOCK_N * 4 + 1024;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_halo_stream_kernel(const __grid_constant__ CUtensorMap tmA,
                           const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ ConvParams p) {
  using Cfg = HaloStreamCfg<BLOCK_N>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                                   // kWStages x [BLOCK_N][64]
  uint8_t* s_patch = smem + Cfg::kWStages * Cfg::kWTile;  // kSlots x kPatchSlot
  uint64_t* pfull = reinterpret_cast<uint64_t*>(s_patch + Cfg::kSlots * Cfg::kPatchSlot);
  uint64_t* pempty = pfull + Cfg::kSlots;
  uint64_t* wfull = pempty + Cfg::kSlots;
  uint64_t* wempty = wfull + Cfg::kWStages;
  uint64_t* tfull_bar = wempty + Cfg::kWStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_sum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pfull) + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;
  float* s_bn = s_x2 + BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kSlots; ++s) {
      mbar_init(&pfull[s], 1);
      mbar_init(&pempty[s], 1);
    }
    for (int s = 0; s < Cfg::kWStages; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * BLOCK_N; i += kConvThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles;
  const int chunks = p.taps[0].kchunks;

  if (warp == 0) {
    if (elect_one()) {
      int slot = 0, ws = 0;
      uint32_t pphase = 0, wphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.tw;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.th;
        const int b0 = mt / p.tiles_h;
        for (int kc = 0; kc < chunks; ++kc) {
          mbar_wait(&pempty[slot], pphase ^ 1);
          mbar_expect_tx(&pfull[slot], Cfg::kPatchBytes);
          tma_load_5d(s_patch + slot * Cfg::kPatchSlot, &tmA, &pfull[slot], kc * 64, w0 - 1, 0,
                      h0 - 1, b0);
          if (++slot == Cfg::kSlots) {
            slot = 0;
            pphase ^= 1;
          }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&wempty[ws], wphase ^ 1);
            mbar_expect_tx(&wfull[ws], Cfg::kWTile);
            bulk_load(s_w + ws * Cfg::kWTile,
                      p.w[0] + ((size_t)(p.taps[t].btap * p.w_kc[0] + kc) * p.w_rb[0] +
                                n_tile * (BLOCK_N / 64)) * 4096,
                      Cfg::kWTile, &wfull[ws]);
            if (++ws == Cfg::kWStages) {
              ws = 0;
              wphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      int slot = 0, ws = 0;
      uint32_t pphase = 0, wphase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kc = 0; kc < chunks; ++kc) {
          mbar_wait(&pfull[slot], pphase);
          tc_fence_after();
          const uint32_t patch = smem_u32(s_patch + slot * Cfg::kPatchSlot);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&wfull[ws], wphase);
            tc_fence_after();
            const int start_row = (1 + p.taps[t].d3) * 10 + (1 + p.taps[t].d1);
            const uint64_t adesc = make_smem_desc(patch + start_row * 128, 16, 1280);
            const uint64_t bdesc = make_smem_desc(smem_u32(s_w + ws * Cfg::kWTile), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | t | k) != 0);
            umma_commit(&wempty[ws]);
            if (++ws == Cfg::kWStages) {
              ws = 0;
              wphase ^= 1;
            }
          }
          umma_commit(&pempty[slot]);
          if (++slot == Cfg::kSlots) {
            slot = 0;
            pphase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    conv_epilogue<BLOCK_N, 1>(p, tmem_base, tfull_bar, tempty_bar, s_sum, s_sq, s_x2, s_bn, 0,
                              blockIdx.x, gridDim.x, total_tiles, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ===========================================================================
// K2c: weight gradient, dW[tap][co][ci] += sum over pixels dY[p][co] * X[p+tap][ci]
//
// GEMM view: the reduction runs over PIXELS, which is the slow axis of both NHWC
// operands, so both are fed to the tensor core MN-major straight from the same
// TMA boxes the forward uses (no transposed copies):
//     D[128 = two (tap, 64-channel) units of X][BLOCK_N couts] +=
//         A[128 pixels, 128]^T  x  B[128 pixels, BLOCK_N]
// A work item is (unit pair, cout tile, pixel split); it walks its share of the
// pixel tiles accumulating in TMEM and finally adds its partial result into the
// fp32 gradient arena (tap-major [tap][Cout][Cin], the arena's native layout)
// with coalesced red.global.add.f32.
// ===========================================================================
struct WgradParams {
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;          // Cout / BLOCK_N
  int num_units;        // taps * (Cin / 64)
  int num_pairs;        // ceil(num_units / 2)
  int splits;           // pixel-range splits per (pair, n_tile)
  int kchunks;          // Cin / 64
  int num_taps;
  ConvTap taps[kMaxTaps];
  int cin, cout;
  int row_limit;        // rows (ci within unit) >= row_limit are not written (stem pad)
  float* dw;            // [taps][cout][cin] fp32, accumulated
};

template <int BLOCK_N>
struct WgradCfg {
  static constexpr int kABytes = 2 * kBlockM * 64 * 2;        // two unit boxes
  static constexpr int kBBytes = (BLOCK_N / 64) * kBlockM * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 64 ? 4 : 3;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<BLOCK_N>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail

  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per_split = (pix_tiles + p.splits - 1) / p.splits;
  const int total_items = p.num_pairs * p.n_tiles * p.splits;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item % p.splits;
        const int n_tile = (item / p.splits) % p.n_tiles;
        const int pair = item / (p.splits * p.n_tiles);
        int u[2] = {2 * pair, 2 * pair + 1};
        if (u[1] >= p.num_units) u[1] = u[0];  // dummy second unit (result discarded)
        const int pt_end = min(pix_tiles, (split + 1) * per_split);
        for (int pt = split * per_split; pt < pt_end; ++pt) {
          int mt = pt;
          const int w0 = (mt % p.tiles_w) * p.tw;
          mt /= p.tiles_w;
          const int h0 = (mt % p.tiles_h) * p.th;
          const int b0 = (mt / p.tiles_h) * p.tn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const ConvTap tap = p.taps[u[j] / p.kchunks];
            const int kc = u[j] % p.kchunks;
            tma_load_5d(sa + j * (kBlockM * 128), &tmX, &full_bar[stage], tap.c0 + kc * 64,
                        w0 + tap.d1, tap.d2, h0 + tap.d3, b0);
          }
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_5d(sb + j * (kBlockM * 128), &tmDY, &full_bar[stage],
                        n_tile * BLOCK_N + j * 64, w0, 0, h0, b0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item % p.splits;
        const int pt_beg = split * per_split;
        const int pt_end = min(pix_tiles, (split + 1) * per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int pt = pt_beg; pt < pt_end; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          // MN-major: LBO = bytes between 64-wide M/N blocks (one TMA box),
          // SBO = bytes between 8-pixel groups along K
          const uint64_t adesc = make_smem_desc(sa, kBlockM * 128, 1024);
          const uint64_t bdesc = make_smem_desc(sb, kBlockM * 128, 1024);
#pragma unroll
          for (int k = 0; k < kBlockM / kUmmaK; ++k) {
            // 16 pixels (rows of 128 B) per MMA -> 2048 B -> +128 in the address field
            umma_bf16(d_tmem, adesc + 128 * k, bdesc + 128 * k, idesc,
                      (pt != pt_beg || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int split = item % p.splits;
      const int n_tile = (item / p.splits) % p.n_tiles;
      const int pair = item / (p.splits * p.n_tiles);
      const int unit = 2 * pair + (r >> 6);
      const int rl = r & 63;
      const bool has_work = split * per_split < pix_tiles;
      const bool valid = has_work && unit < p.num_units && rl < p.row_limit;
      const int tap_i = valid ? unit / p.kchunks : 0;
      const int kc = valid ? unit % p.kchunks : 0;
      float* dst = p.dw + ((size_t)p.taps[tap_i].btap * p.cout + n_tile * BLOCK_N) * p.cin +
                   kc * 64 + rl;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(dst + (size_t)(c * 32 + j) * p.cin, __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace vpd
