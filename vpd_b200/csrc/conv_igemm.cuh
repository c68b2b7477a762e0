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
//     BatchNorm needs (warp transpose-reduce -> smem -> order-independent integer global atomics, see StatAcc).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace vpd {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 12;
constexpr int kEpiWarps = 4;                       // epilogue warps (one per TMEM lane quarter; 8 was slower)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kStatWarps = 4;                      // reduction warps (one per 32 tile rows)
constexpr int kStatThreads = kStatWarps * 32;
// Three aligned warpgroups so that registers can be re-balanced with setmaxnreg:
//   warps 0-3: TMA producer (0), MMA issuer (1), two idle warps; warps 4-7: epilogue;
//   warps 8-11: statistics / store
constexpr int kEpiWarp0 = 4;
constexpr int kStatWarp0 = 8;
constexpr int kEpiThread0 = kEpiWarp0 * 32;
constexpr int kStatThread0 = kStatWarp0 * 32;
constexpr int kConvThreads = kStatThread0 + kStatThreads;
constexpr int kRegsCtl = 64, kRegsEpi = 152, kRegsStat = 232;
constexpr int kWgradThreads = 192;

// Division by a launch constant without the ~30-instruction signed-division sequence
// (Granlund-Montgomery, unsigned n < 2^32): q = (t + ((n - t) >> s1)) >> s2, t = umulhi(m, n).
// The per-tile coordinate decode sits on the critical path of the statistics warps: with four
// runtime divisions it was 120 of their ~880 instructions per tile.
struct FastDiv {
  uint32_t d, m, s1, s2;
};
inline FastDiv fd_make(uint32_t d) {
  FastDiv f;
  if (d == 0) d = 1;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.d = d;
  f.m = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l > 0 ? l - 1 : 0;
  return f;
}
__host__ __device__ __forceinline__ uint32_t fd_div(const FastDiv& f, uint32_t n) {
#ifdef __CUDA_ARCH__
  const uint32_t t = __umulhi(f.m, n);
#else
  const uint32_t t = static_cast<uint32_t>((static_cast<uint64_t>(f.m) * n) >> 32);   // host: tests
#endif
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

struct ConvTap {
  int c0;       // offset added to the innermost (channel) coordinate
  int d1, d2, d3;  // offsets for the w, parity and h coordinates
  int src;      // which (A,B) tensor-map pair (0/1)
  int btap;     // index on the tap axis of the weight tensor
  int kchunks;  // number of 64-wide K chunks for this tap
  int kskip;    // halo kernels: bit k set = K step k (16 channels) of this tap has all-zero weights
                // and is not issued (the stem's taps cover 8 x 8 positions of a 7 x 7 kernel)
};

struct ConvParams {
  int tw, th, tn;  // tile = tw x th pixels x tn images, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;     // Cout / BLOCK_N
  int num_taps;
  ConvTap taps[kMaxTaps];
  int batch, out_h, out_w;  // valid extents of the (n, h, w) tile coordinates
  int cout;
  __nv_bfloat16* out;                 // written through the output tensor map (TMA store)
  // Output classes. Ordinary convolutions have one; the data gradient of a stride-2 conv has
  // four (one per output-pixel parity), each with its own tap subset, its own coordinate
  // offsets in the strided output view and its own base offset for `out`-addressed operands.
  // They run as ONE launch: tile index = class * tiles_per_class + (pixel tile, channel block).
  int num_classes;
  struct OutClass {
    int tap0, ntaps;       // taps[tap0 .. tap0 + ntaps)
    int out_c0, out_d2;    // channel / parity coordinate offsets of the output view
    long long base;        // element offset of the class's first output pixel
  } cls[4];
  const __nv_bfloat16* residual;      // same addressing as out, or null
  long long out_sn, out_sh, out_sw;   // element strides of out/residual
  const float* scale;                 // [cout] or null
  const float* shift;                 // [cout] or null
  int relu;
  StatAcc* stats;                     // [2][cout] (sum, sumsq) or null
  // pre-tiled bf16 weights (see wtile_offset) for tap.src 0 / 1: element offset of the
  // [BLOCK_N x 64] tile (tap t, chunk kc, rows n0..) = ((t*w_kc + kc)*w_rb + n0/64) * 4096
  const __nv_bfloat16* w[2];
  int w_kc[2];                        // K chunks per tap
  int w_rb[2];                        // 64-row blocks (N dimension / 64)
  // Fused BatchNorm-backward reduction (dgrad launches): the tile just computed is
  // dz of a ReLU->BN stage; mask it with 1[z > 0] before it is stored (so `out`
  // holds g) and accumulate sum(g), sum(g * xhat_b) for up to two BN branches.
  int bnb;                            // 0 = off, else number of branches (1 or 2)
  // 1[z > 0] of that stage's post-ReLU output as ONE BIT per element, written by the forward
  // BatchNorm kernel: byte (pixel, g) covers channels 8g .. 8g+7 (element offset / 8 in out's
  // addressing). Reading z itself cost as many bytes as the gradient tile: the stage-1 data
  // gradients were bound by HBM traffic (dy + z + y in, g out)
  const uint8_t* bmask;
  const __nv_bfloat16* by[2];         // its pre-BN conv outputs
  const float* bmean[2];              // [cout] saved batch mean
  const float* brstd[2];              // [cout] saved 1/sqrt(var+eps)
  StatAcc* bsums[2];                  // [2][cout]: sum g, sum g*xhat
  // debug (vpd_conv_trace): 8 int64 per CTA - globaltimer at entry, then clock64 at
  // entry / after the dependency wait / first operands landed / last MMA issued /
  // first accumulator ready / epilogue done / exit
  long long* trace;
  // weights of the NEXT convolution in the stream: every CTA pulls one slice into L2 while
  // this kernel runs (they come from HBM once per step and nothing else would hide that)
  const __nv_bfloat16* pf_ptr;
  long long pf_bytes;
  // halo-reuse kernels: the patch box (tile + halo) starts at tile origin + (patch_dx, patch_dy)
  // and is patch_bytes long; tap t reads it from patch row (1 + d3) * 10 + (1 + d1)
  int patch_dx, patch_dy, patch_bytes;
  // generic kernel, every CTA has at most ONE tile: once its accumulator is complete the operand
  // ring is dead, so every output slab gets its own 16 KB of it - the epilogue warps drain TMEM
  // without waiting for slab slots and the statistics warps follow one slab behind (with the
  // one- or two-slot ring TMEM drain, statistics and store of a slab ran back to back)
  int single_tile;
  // decode_tile's divisors (launch_conv fills them): work items per output class, channel
  // blocks, pixel tiles per row / column
  FastDiv fd_class, fd_ntiles, fd_tw, fd_th;
  int dbg;  // diagnostics (VPD_DBG_SKIP): bit0 skip the A loads, bit1 skip the B loads, bit2 skip the epilogue body
};
struct TileCoord {
  int cls, n_tile, w0, h0, b0;
};
// tile index -> (class, channel block, first pixel); CS = CTAs per cluster, rank = CTA in it
template <int CS>
VPD_DEVINL TileCoord decode_tile(const ConvParams& p, int tile, int rank) {
  TileCoord t;
  const uint32_t u = static_cast<uint32_t>(tile);
  const uint32_t c = p.num_classes > 1 ? fd_div(p.fd_class, u) : 0u;
  const uint32_t inner = u - c * p.fd_class.d;
  const uint32_t q = fd_div(p.fd_ntiles, inner);
  const uint32_t mt = q * CS + rank;
  const uint32_t r1 = fd_div(p.fd_tw, mt);
  const uint32_t r2 = fd_div(p.fd_th, r1);
  t.cls = static_cast<int>(c);
  t.n_tile = static_cast<int>(inner - q * p.fd_ntiles.d);
  t.w0 = static_cast<int>(mt - r1 * p.fd_tw.d) * p.tw;
  t.h0 = static_cast<int>(r1 - r2 * p.fd_th.d) * p.th;
  t.b0 = static_cast<int>(r2) * p.tn;
  return t;
}
// one thread per CTA: this CTA's slice of the next convolution's weights -> L2
VPD_DEVINL void conv_prefetch_next(const ConvParams& p) {
  if (p.pf_ptr == nullptr) return;
  const long long per = ((p.pf_bytes / gridDim.x) + 15) & ~15LL;
  const long long beg = per * blockIdx.x;
  long long n = p.pf_bytes - beg;
  if (n > per) n = per;
  if (n >= 16)
    bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.pf_ptr) + beg,
                     static_cast<uint32_t>(n & ~15LL));
}
VPD_DEVINL void trace_mark(const ConvParams& p, int slot) {
  if (p.trace != nullptr) p.trace[blockIdx.x * 16 + slot] = clock64();
}

// Output staging: the epilogue warps write finished 128-pixel x 64-channel SLABS (bf16,
// 128 B per pixel row, 16-byte chunks XOR-swizzled by row % 8 = the TMA SWIZZLE_128B image)
// into a small shared-memory ring; the statistics warps reduce / mask them there and one
// thread sends each slab to global memory with a single TMA tile store.
constexpr int kSlabBytes = kBlockM * 128;
constexpr int kStatScratchBytes = 2 * 4 * 3 * 64 * 4;  // conv_stats cross-warp scratch
template <int BLOCK_N>
struct StageCfg {
  static constexpr int kSlabs = BLOCK_N / 64;          // slabs per tile
  static constexpr int kSlots = BLOCK_N == 256 ? 1 : 2;  // ring depth (shared-memory budget)
  static constexpr int kBytes = kSlots * kSlabBytes;
};

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 64 ? 6 : (BLOCK_N == 128 ? 5 : 4);
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + StageCfg<BLOCK_N>::kBytes + kBarBytes + 3 * BLOCK_N * 4 +
      kStatScratchBytes + 1024;
};

// Epilogue, warps 2..5 (threads 64..191): drains the TMEM accumulator stages tile by tile -
// optional folded-BN affine, residual, ReLU - and writes bf16 slabs to the staging ring.
// It performs no reductions and no global stores.
template <int BLOCK_N, int CS, int MODE>
VPD_DEVINL void conv_epilogue(const ConvParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                              uint64_t* tempty_bar, uint8_t* slabs, uint64_t* sfull,
                              uint64_t* sempty, float* s_scale, float* s_shift, uint8_t* ring,
                              uint64_t* zfull, int rank,
                              int first_item, int item_stride, int total_tiles, int warp,
                              int lane) {
  using SC = StageCfg<BLOCK_N>;
  // launches whose CTAs have a single tile (ConvParams::single_tile): once the accumulator is
  // complete the operand ring is dead, so every slab gets its own 16 KB of it (no slot hand-shake)
  const bool direct = ring != nullptr && p.single_tile != 0;
  // pair mode (CS == 2): the leader's MMA thread owns the accumulator hand-shake, so the
  // peer's epilogue warps release the TMEM stage on the LEADER's barrier
  const uint32_t tempty_remote0 =
      (CS == 2 && rank == 1) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
  const int q = warp & 3;       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;  // row of the 128-row tile
  const uint32_t row_addr = smem_u32(slabs) + r * 128;
  const uint32_t rsw = static_cast<uint32_t>(r & 7);
  // position of this row inside a tile (tile extents are powers of two)
  const int ltw = __ffs(p.tw) - 1, lth = __ffs(p.th) - 1;
  const int r_w = r & (p.tw - 1), r_h = (r >> ltw) & (p.th - 1), r_n = r >> (ltw + lth);
  const long long r_off = r_n * p.out_sn + r_h * p.out_sh + r_w * p.out_sw;
  int as = 0;
  uint32_t aphase = 0;
  int slot = 0;
  uint32_t sphase = 0;
  int cached_ntile = -1;
  for (int tile = first_item; tile < total_tiles; tile += item_stride) {
    const TileCoord tc = decode_tile<CS>(p, tile, rank);
    const int n_tile = tc.n_tile;
    const bool valid = (tc.b0 + r_n < p.batch) && (tc.h0 + r_h < p.out_h) && (tc.w0 + r_w < p.out_w);
    const long long off = p.cls[tc.cls].base + tc.b0 * p.out_sn + tc.h0 * p.out_sh +
                          tc.w0 * p.out_sw + r_off + n_tile * BLOCK_N;

    // fused BN backward (MODE 2 / 3): this row's ReLU mask bits, 32 channels per word, requested
    // before the wait for the accumulator (a whole main loop hides the latency)
    uint32_t mw[2 * SC::kSlabs];
    if (MODE == 2 || MODE == 3 || (MODE < 0 && p.bnb > 0)) {
#pragma unroll
      for (int c = 0; c < 2 * SC::kSlabs; ++c)
        mw[c] = valid ? ldg_nc_u32(p.bmask + ((off + c * 32) >> 3)) : 0u;
    }
    // the residual rows of the FIRST slab likewise (a row's residual may alias its output, but
    // that is written by this tile's own store, after this read); later slabs load in place
    const __nv_bfloat16* resp = MODE == 1 ? nullptr : p.residual;
    const bool do_res = resp != nullptr && valid;
    uint4 rpre[2][4];
    if (do_res) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const uint4* rp = reinterpret_cast<const uint4*>(resp + off + cc * 32);
#pragma unroll
        for (int k = 0; k < 4; ++k) rpre[cc][k] = rp[k];
      }
    }
    // folded BatchNorm (eval): this channel block's scale / shift through shared memory (every
    // thread of a warp reads the same 16 bytes: one broadcast instead of 16 global loads per
    // chunk and thread)
    if (MODE != 1 && p.scale != nullptr && n_tile != cached_ntile) {
      asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");   // the old block is no longer read
      for (int i = threadIdx.x - kEpiThread0; i < BLOCK_N; i += kEpiThreads) {
        s_scale[i] = __ldg(p.scale + n_tile * BLOCK_N + i);
        s_shift[i] = __ldg(p.shift + n_tile * BLOCK_N + i);
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
      cached_ntile = n_tile;
    }
    mbar_wait(&tfull_bar[as], aphase);
    if (tile == first_item && threadIdx.x == kEpiThread0) trace_mark(p, 5);
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < SC::kSlabs; ++j) {
      if (!direct) mbar_wait(&sempty[slot], sphase ^ 1);  // the slab's previous contents have been stored
      const uint32_t dst = direct ? smem_u32(ring) + j * kSlabBytes + r * 128
                                  : row_addr + slot * kSlabBytes;
#pragma unroll 1
      for (int cc = 0; cc < ((p.dbg & 4) ? 0 : 2); ++cc) {
        const int c = 2 * j + cc;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c * 32, v);
        // training-forward kernels (MODE 1): a plain conversion
        uint4 rres[4];
        if (do_res) {
          if (j == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) rres[k] = cc == 0 ? rpre[0][k] : rpre[1][k];
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(resp + off + c * 32);
#pragma unroll
            for (int k = 0; k < 4; ++k) rres[k] = rp[k];  // plain load: residual may alias out
          }
        }
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
        const int ch0 = n_tile * BLOCK_N + c * 32;
        if (MODE != 1 && p.scale != nullptr) {
          const uint32_t sc_a = smem_u32(s_scale + c * 32), sh_a = smem_u32(s_shift + c * 32);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 sc = __uint4_as_float4(lds_v4(sc_a + k * 4));
            const float4 sh = __uint4_as_float4(lds_v4(sh_a + k * 4));
            f[k + 0] = fmaf(f[k + 0], sc.x, sh.x);
            f[k + 1] = fmaf(f[k + 1], sc.y, sh.y);
            f[k + 2] = fmaf(f[k + 2], sc.z, sh.z);
            f[k + 3] = fmaf(f[k + 3], sc.w, sh.w);
          }
        }
        if (do_res) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            f[8 * k + 0] += bf16_lo(rres[k].x);
            f[8 * k + 1] += bf16_hi(rres[k].x);
            f[8 * k + 2] += bf16_lo(rres[k].y);
            f[8 * k + 3] += bf16_hi(rres[k].y);
            f[8 * k + 4] += bf16_lo(rres[k].z);
            f[8 * k + 5] += bf16_hi(rres[k].z);
            f[8 * k + 6] += bf16_lo(rres[k].w);
            f[8 * k + 7] += bf16_hi(rres[k].w);
          }
        }
        if (MODE != 1 && p.relu) {
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = fmaxf(f[k], 0.f);
        }
        if (!valid) {  // rows outside the tensor: clipped by the store, zero for the sums
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = 0.f;
        }
        const bool masked = MODE == 2 || MODE == 3 || (MODE < 0 && p.bnb > 0);
        // chunks come in order: take the head of the queue and shift it (static register indices)
        const uint32_t mword = masked ? mw[0] : 0u;
        if (masked) {
#pragma unroll
          for (int i = 0; i + 1 < 2 * SC::kSlabs; ++i) mw[i] = mw[i + 1];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 o = make_uint4(pack_bf16x2(f[8 * k + 0], f[8 * k + 1]),
                               pack_bf16x2(f[8 * k + 2], f[8 * k + 3]),
                               pack_bf16x2(f[8 * k + 4], f[8 * k + 5]),
                               pack_bf16x2(f[8 * k + 6], f[8 * k + 7]));
          if (masked) {
            // g = dz * 1[z > 0]. Bits 0-3 / 4-7 of the byte -> the sign bits of four bytes (no
            // carries: the partial products of the multiplier do not overlap), then byte-wise
            // sign replication gives 0xFFFF per selected bf16
            const uint32_t b = (mword >> (8 * k)) & 0xFFu;
            const uint32_t s03 = ((b & 0xFu) * 0x10204080u) & 0x80808080u;
            const uint32_t s47 = ((b >> 4) * 0x10204080u) & 0x80808080u;
            o.x &= prmt(s03, 0u, 0x9988u);
            o.y &= prmt(s03, 0u, 0xBBAAu);
            o.z &= prmt(s47, 0u, 0x9988u);
            o.w &= prmt(s47, 0u, 0xBBAAu);
          }
          sts_v4(dst + (((static_cast<uint32_t>(cc * 4 + k)) ^ rsw) << 4), o);
        }
      }
      if (j == SC::kSlabs - 1) {
        // accumulator fully read: hand the TMEM stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CS == 2 && rank == 1) mbar_arrive_remote(tempty_remote0 + as * 8);
          else mbar_arrive(&tempty_bar[as]);
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
      if (direct) {
        mbar_arrive(&zfull[j]);
      } else {
        mbar_arrive(&sfull[slot]);
        if (++slot == SC::kSlots) {
          slot = 0;
          sphase ^= 1;
        }
      }
    }
    if (++as == 2) {
      as = 0;
      aphase ^= 1;
    }
  }
}

// Statistics / store warps (warps 6..9, threads 192..319). For every staged slab:
//   forward train (p.stats):  sum y, sum y^2 per channel  -> BatchNorm batch statistics
//   dgrad (p.bnb):  the slab holds g = dz * 1[z > 0] of a ReLU->BN stage (the epilogue warps
//                   applied the bit mask ConvParams::bmask): accumulate sum g and sum g*(y - mean)
//                   for up to two BN branches               -> fused BN-backward reduction
// always on the stored (bf16-rounded) values; then ONE thread issues the slab's TMA tile
// store. A lane owns eight consecutive channels and the four row groups of a warp
// instruction cover whole 128-byte rows, so every shared / global access is fully
// coalesced; the z / y rows of the NEXT slab are requested before the current one is
// reduced, so their latency hides behind a whole slab period. Cross-warp totals are
// combined through a small double-buffered scratch with exclusive owners (no shared-memory
// float atomics: those are CAS loops) and reach the fp64 global accumulators when the
// channel block changes and at the end.
// MODE: 0 store only (eval / plain dgrad), 1 BatchNorm statistics (training forward),
// 2 / 3 fused BN-backward reduction with one / two BN branches (dgrad), -1 decided at run time. The specialised kernels
// carry only their own reduction code: these kernels are large enough for instruction-cache
// misses to show (keeping two epilogue flavours in one kernel cost ~10 % on the deep layers).
template <int BLOCK_N, int CS, int MODE>
VPD_DEVINL void conv_stats(const ConvParams& p, const CUtensorMap* tm_out,
                           uint8_t* ring,
                           uint64_t* zfull, uint8_t* slabs,
                           uint64_t* sfull, uint64_t* sempty, float* s_sum, float* s_sq,
                           float* s_x2, float* s_scr, int rank, int first_item, int item_stride,
                           int total_tiles, int sw, int lane) {
  using SC = StageCfg<BLOCK_N>;
  const int cg = lane & 7;     // 16-byte chunk (8 channels) of the 128-byte slab row
  const int rsub = lane >> 3;  // row within the 4 rows one warp instruction covers
  const int ltw = __ffs(p.tw) - 1, lth = __ffs(p.th) - 1;  // tile extents are powers of two
  const int nbr = MODE < 0 ? p.bnb : (MODE == 2 ? 1 : (MODE == 3 ? 2 : 0));
  const bool fwd_stats = MODE < 0 ? (p.stats != nullptr) : (MODE == 1);
  const bool sums = fwd_stats || nbr > 0;
  const bool live = !(p.dbg & (4 | 16));
  const int st = threadIdx.x - kStatThread0;  // 0..127
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a0[k] = a1[k] = a2[k] = 0.f;

  // this warp's register sums -> scratch[par][sw][q][64] (lanes 8 apart hold the same channels)
  auto regs_to_scratch = [&](int par) {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a0[k] += __shfl_xor_sync(0xffffffffu, a0[k], o);
        a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o);
        if (nbr > 1) a2[k] += __shfl_xor_sync(0xffffffffu, a2[k], o);
      }
    }
    if (lane < 8) {
      const uint32_t base = smem_u32(s_scr + ((par * kStatWarps + sw) * 3) * 64 + cg * 8);
      sts_v4(base, make_uint4(__float_as_uint(a0[0]), __float_as_uint(a0[1]),
                              __float_as_uint(a0[2]), __float_as_uint(a0[3])));
      sts_v4(base + 16, make_uint4(__float_as_uint(a0[4]), __float_as_uint(a0[5]),
                                   __float_as_uint(a0[6]), __float_as_uint(a0[7])));
      sts_v4(base + 256, make_uint4(__float_as_uint(a1[0]), __float_as_uint(a1[1]),
                                    __float_as_uint(a1[2]), __float_as_uint(a1[3])));
      sts_v4(base + 256 + 16, make_uint4(__float_as_uint(a1[4]), __float_as_uint(a1[5]),
                                         __float_as_uint(a1[6]), __float_as_uint(a1[7])));
      if (nbr > 1) {
        sts_v4(base + 512, make_uint4(__float_as_uint(a2[0]), __float_as_uint(a2[1]),
                                      __float_as_uint(a2[2]), __float_as_uint(a2[3])));
        sts_v4(base + 512 + 16, make_uint4(__float_as_uint(a2[4]), __float_as_uint(a2[5]),
                                           __float_as_uint(a2[6]), __float_as_uint(a2[7])));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a0[k] = a1[k] = a2[k] = 0.f;
  };
  // after a barrier: scratch[par] of the four warps -> per-CTA sums of slab j. Thread st
  // owns channel st % 64 of quantity 0 / 2 (st < 64) or 1 (st >= 64): plain read-modify-write
  auto combine = [&](int j, int par) {
    const int c = st & 63;
    const float* scr = s_scr + par * kStatWarps * 3 * 64 + c;
    if (st < 64) {
      s_sum[j * 64 + c] += scr[0] + scr[3 * 64] + scr[6 * 64] + scr[9 * 64];
      if (nbr > 1)
        s_x2[j * 64 + c] += scr[2 * 64] + scr[5 * 64] + scr[8 * 64] + scr[11 * 64];
    } else {
      s_sq[j * 64 + c] += scr[64] + scr[4 * 64] + scr[7 * 64] + scr[10 * 64];
    }
  };
  // per-CTA sums -> fp64 global accumulators (N == 64 keeps its sums in registers until here)
  auto global_flush = [&](int ntile) {
    if (BLOCK_N == 64) {
      regs_to_scratch(0);
      asm volatile("bar.sync 2, %0;" ::"n"(kStatThreads) : "memory");
      combine(0, 0);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kStatThreads) : "memory");
    for (int i = st; i < BLOCK_N; i += kStatThreads) {
      const int c = ntile * BLOCK_N + i;
      if (fwd_stats) {
        stat_add(&p.stats[c], static_cast<double>(s_sum[i]));
        stat_add(&p.stats[p.cout + c], static_cast<double>(s_sq[i]));
      } else {
        // sum g * xhat = rstd * (sum g * y - mean * sum g)
        const double sg = static_cast<double>(s_sum[i]);
        stat_add(&p.bsums[0][c], sg);
        stat_add(&p.bsums[0][p.cout + c],
                 (static_cast<double>(s_sq[i]) - static_cast<double>(__ldg(p.bmean[0] + c)) * sg) *
                     static_cast<double>(__ldg(p.brstd[0] + c)));
        if (nbr > 1) {
          stat_add(&p.bsums[1][c], sg);
          stat_add(&p.bsums[1][p.cout + c],
                   (static_cast<double>(s_x2[i]) - static_cast<double>(__ldg(p.bmean[1] + c)) * sg) *
                       static_cast<double>(__ldg(p.brstd[1] + c)));
        }
      }
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
      s_x2[i] = 0.f;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kStatThreads) : "memory");
  };

  // ---- prefetched operands of the fused BN backward: this warp's 32 rows of the mask, y0 (, y1).
  // All offsets are 32-bit element counts (the planner rejects tensors of 2^31 elements or more)
  const int sn = static_cast<int>(p.out_sn), sh = static_cast<int>(p.out_sh),
            sw_ = static_cast<int>(p.out_sw);
  int rel[8];   // element offset of row i of this lane inside a tile (tile-invariant)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = sw * 32 + i * 4 + rsub;
    rel[i] = (r >> (ltw + lth)) * sn + ((r >> ltw) & (p.th - 1)) * sh + (r & (p.tw - 1)) * sw_ + cg * 8;
  }
  uint4 y0[8], y1[8];
  int cls = 0;  // output class of the tile the cursor is on
  auto issue_loads = [&](int n_tile, int w0, int h0, int b0, int j) {
    const int cbase = static_cast<int>(p.cls[cls].base) + n_tile * BLOCK_N + j * 64;
    const int base = cbase + b0 * sn + h0 * sh + w0 * sw_;
    const bool full = (b0 + p.tn <= p.batch) && (h0 + p.th <= p.out_h) && (w0 + p.tw <= p.out_w);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bool ok = full;
      if (!full) {
        const int r = sw * 32 + i * 4 + rsub;
        ok = (b0 + (r >> (ltw + lth)) < p.batch) && (h0 + ((r >> ltw) & (p.th - 1)) < p.out_h) &&
             (w0 + (r & (p.tw - 1)) < p.out_w);
      }
      // rows outside the tensor hold g = 0 in the slab: any finite operand will do
      const uint32_t off = static_cast<uint32_t>(ok ? base + rel[i] : cbase + cg * 8);
      y0[i] = ldg_nc_v4(p.by[0] + off);
      if (nbr > 1) y1[i] = ldg_nc_v4(p.by[1] + off);
    }
  };
  auto tile_coords = [&](int tile, int& n_tile, int& w0, int& h0, int& b0) {
    const TileCoord tc = decode_tile<CS>(p, tile, rank);
    cls = tc.cls;
    n_tile = tc.n_tile;
    w0 = tc.w0;
    h0 = tc.h0;
    b0 = tc.b0;
  };
  // shared-memory address of this lane's 16-byte chunk in row i (relative to the slab)
  uint32_t srow[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = sw * 32 + i * 4 + rsub;
    srow[i] = r * 128 + ((cg ^ (r & 7)) << 4);
  }

  int cur_ntile = -1;
  int slot = 0, prev_slot = -1, par = 0;
  uint32_t sphase = 0;
  int tile = first_item, j = 0;
  bool have = tile < total_tiles;
  int n_tile = 0, w0 = 0, h0 = 0, b0 = 0;
  if (have) {
    tile_coords(tile, n_tile, w0, h0, b0);
    if (nbr > 0 && live && !(p.dbg & 32)) issue_loads(n_tile, w0, h0, b0, 0);
    if (nbr > 0 && live && SC::kSlabs > 1 && p.single_tile != 0 && !(p.dbg & 128)) {
      // One tile per CTA: the operands of slabs 1.. are requested only when the slab before
      // them has been reduced, i.e. after the main loop, and they were written a whole
      // forward pass ago. Pull them into L2 now, under the main loop (thread st = tile row st).
      const int pn = b0 + (st >> (ltw + lth)), ph = h0 + ((st >> ltw) & (p.th - 1)),
                pw = w0 + (st & (p.tw - 1));
      if (pn < p.batch && ph < p.out_h && pw < p.out_w) {
        const uint32_t off = static_cast<uint32_t>(static_cast<int>(p.cls[cls].base) +
                                                   n_tile * BLOCK_N + pn * sn + ph * sh + pw * sw_);
#pragma unroll
        for (int jj = 1; jj < SC::kSlabs; ++jj) {
          prefetch_l2(p.by[0] + off + jj * 64);
          if (nbr > 1) prefetch_l2(p.by[1] + off + jj * 64);
        }
      }
    }
  }
  while (have) {
    if (sums && cur_ntile != n_tile) {
      if (cur_ntile >= 0) global_flush(cur_ntile);
      cur_ntile = n_tile;
    }
    const bool direct = ring != nullptr && p.single_tile != 0;   // slab j lives in the dead operand ring
    if (direct) mbar_wait(&zfull[j], 0);
    else mbar_wait(&sfull[slot], sphase);
    const uint8_t* slab_ptr = direct ? ring + j * kSlabBytes : slabs + slot * kSlabBytes;
    const uint32_t slab = smem_u32(slab_ptr);
    if (fwd_stats && live) {
      uint4 g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = lds_v4(slab + srow[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t gw[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = bf16_lo(gw[k]), hi = bf16_hi(gw[k]);
          a0[2 * k] += lo;
          a0[2 * k + 1] += hi;
          a1[2 * k] = fmaf(lo, lo, a1[2 * k]);
          a1[2 * k + 1] = fmaf(hi, hi, a1[2 * k + 1]);
        }
      }
    } else if (nbr > 0 && live) {
      uint4 g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = lds_v4(slab + srow[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // the slab already holds g = dz * 1[z > 0]: the epilogue warps applied the mask
        const uint32_t gw[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
        const uint32_t yw[4] = {y0[i].x, y0[i].y, y0[i].z, y0[i].w};
        const uint32_t y1w[4] = {y1[i].x, y1[i].y, y1[i].z, y1[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = bf16_lo(gw[k]), hi = bf16_hi(gw[k]);
          a0[2 * k] += lo;
          a0[2 * k + 1] += hi;
          // sum g*y here, sum g*(y - mean) = sum g*y - mean * sum g when the CTA's sums leave
          a1[2 * k] = fmaf(lo, bf16_lo(yw[k]), a1[2 * k]);
          a1[2 * k + 1] = fmaf(hi, bf16_hi(yw[k]), a1[2 * k + 1]);
          if (nbr > 1) {
            a2[2 * k] = fmaf(lo, bf16_lo(y1w[k]), a2[2 * k]);
            a2[2 * k + 1] = fmaf(hi, bf16_hi(y1w[k]), a2[2 * k + 1]);
          }
        }
      }
    }
    // store coordinates of the slab just processed
    const int st_c = p.cls[cls].out_c0 + n_tile * BLOCK_N + j * 64, st_w = w0, st_h = h0, st_b = b0;
    const int st_d2 = p.cls[cls].out_d2;
    const int cur_j = j;
    // advance, and request the next slab's operands before anything else
    if (++j == SC::kSlabs) {
      j = 0;
      tile += item_stride;
      have = tile < total_tiles;
      if (have) tile_coords(tile, n_tile, w0, h0, b0);
    }
    if (have && nbr > 0 && live && !(p.dbg & 32)) issue_loads(n_tile, w0, h0, b0, j);
    if (sums && BLOCK_N > 64) regs_to_scratch(par);
    // every statistics thread is done with the slab (and its masked rewrite is visible)
    asm volatile("bar.sync 2, %0;" ::"n"(kStatThreads) : "memory");
    if (sums && BLOCK_N > 64) {
      combine(cur_j, par);
      par ^= 1;
    }
    if (st == 0 && direct) {
      tma_store_5d(tm_out, slab_ptr, st_c, st_w, st_d2, st_h, st_b);
      bulk_commit_group();
    } else if (st == 0) {
      if (!(p.dbg & (4 | 8)))
        tma_store_5d(tm_out, slabs + slot * kSlabBytes, st_c, st_w, st_d2, st_h, st_b);
      bulk_commit_group();
      if (SC::kSlots > 1 && !(p.dbg & 64)) {
        // Release THIS slab's slot as soon as its store has finished reading shared memory.
        // (Releasing the previous slab's slot here instead - one group still pending - costs
        // this thread nothing, but the slot then stays busy for the whole processing time of
        // the next slab: with two slots the epilogue warps could refill it only after the
        // statistics warps had finished the slab in between, i.e. TMEM drain and statistics
        // ran back to back instead of overlapped - the ncu PC samples of the data-gradient
        // kernels sat in exactly that wait.)
        bulk_wait_read0();
        mbar_arrive(&sempty[slot]);
      } else if (SC::kSlots > 1) {
        bulk_wait_read1();
        if (prev_slot >= 0) mbar_arrive(&sempty[prev_slot]);
        prev_slot = slot;
      } else {
        bulk_wait_read0();
        mbar_arrive(&sempty[slot]);
      }
    }
    if (!direct && ++slot == SC::kSlots) {
      slot = 0;
      sphase ^= 1;
    }
  }
  if (sums && cur_ntile >= 0) global_flush(cur_ntile);
  // the last stores only need to have READ their slabs before the CTA (and its shared memory)
  // goes away; grid completion makes the writes visible to the dependent kernel
  if (st == 0) bulk_wait_read0();
}

// CS = 2: PAIR mode (tcgen05 cta_group::2). The two CTAs of a cluster own two adjacent
// pixel tiles of the same channel block and act as ONE 256 x BLOCK_N MMA unit: each loads
// its own A tile and HALF of the weight tile; the leader issues the MMAs, which read A
// from both CTAs' shared memory and the weight halves from both, and write each CTA's
// 128 rows to its own TMEM. A 128x128 single-CTA tile needs 128 B/cycle of operand reads
// plus as much TMA write traffic - twice the 128 B/cycle the shared memory can move -
// which is what pins the single-CTA kernel near half of the tensor peak; in pair mode a
// CTA reads/writes 3/4 (N=128) or 1/2 (N=256) as many bytes per MAC.
template <int BLOCK_N, int CS, int MODE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* slabs = smem + Cfg::kStages * Cfg::kStageBytes;  // output staging ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + StageCfg<BLOCK_N>::kBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* pfull_bar = tempty_bar + 2;  // leader only: "the peer's stage has landed"
  uint64_t* sfull_bar = pfull_bar + Cfg::kStages;  // epilogue -> statistics warps: slab staged
  uint64_t* sempty_bar = sfull_bar + 2;            // slab stored, slot free
  uint64_t* zfull_bar = sempty_bar + 2;            // single-tile launches: slab j staged in the ring
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(zfull_bar + 8);
  float* s_sum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;  // second BN branch (fused backward reduction)
  float* s_scr = s_x2 + BLOCK_N;

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
      mbar_init(&sfull_bar[s], kEpiThreads);
      mbar_init(&sempty_bar[s], 1);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&zfull_bar[s], kEpiThreads);
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmOut);
    conv_prefetch_next(p);
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
    p.trace[blockIdx.x * 16] = static_cast<long long>(g);
    trace_mark(p, 1);
  }
  pdl_wait();  // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) trace_mark(p, 2);

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  // work items are (group of CS pixel tiles, channel block); a CTA takes the pixel
  // tile `group*CS + rank` (possibly past the end: loads zero-fill, stores masked)
  const int rank = CS > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int total_tiles = ((m_tiles + CS - 1) / CS) * p.n_tiles * p.num_classes;
  const int first_item = blockIdx.x / CS;
  const int item_stride = gridDim.x / CS;

  if (warp >= kStatWarp0) {
    setmaxnreg_inc<kRegsStat>();
    conv_stats<BLOCK_N, CS, MODE>(p, &tmOut, smem, zfull_bar, slabs, sfull_bar,
                                  sempty_bar, s_sum, s_sq, s_x2, s_scr, rank, first_item,
                                  item_stride, total_tiles, warp - kStatWarp0, lane);
  } else if (warp >= kEpiWarp0) {
    setmaxnreg_dec<kRegsEpi>();
    conv_epilogue<BLOCK_N, CS, MODE>(p, tmem_base, tfull_bar, tempty_bar, slabs, sfull_bar,
                                     sempty_bar, s_sum, s_sq, smem, zfull_bar, rank,
                                     first_item, item_stride, total_tiles, warp, lane);
  } else {
   setmaxnreg_dec<kRegsCtl>();
   if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        const TileCoord tc = decode_tile<CS>(p, tile, rank);
        const int n_tile = tc.n_tile, w0 = tc.w0, h0 = tc.h0, b0 = tc.b0;
        const int tap_end = p.cls[tc.cls].tap0 + p.cls[tc.cls].ntaps;
        for (int t = p.cls[tc.cls].tap0; t < tap_end; ++t) {
          const ConvTap tap = p.taps[t];
          const CUtensorMap* ma = tap.src ? &tmA1 : &tmA0;
          for (int kc = 0; kc < tap.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            mbar_expect_tx(&full_bar[stage], ((p.dbg & 1) ? 0 : Cfg::kABytes) +
                                                 ((p.dbg & 2) ? 0 : Cfg::kBBytes / CS));
            if (!(p.dbg & 1))
              tma_load_5d(sa, ma, &full_bar[stage], tap.c0 + kc * kBlockK, w0 + tap.d1, tap.d2,
                          h0 + tap.d3, b0);
            // this CTA's share of the weight tile: BLOCK_N / CS rows
            if (!(p.dbg & 2))
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
      int kb_cls[4] = {0, 0, 0, 0};
      for (int c = 0; c < p.num_classes; ++c)
        for (int t = 0; t < p.cls[c].ntaps; ++t) kb_cls[c] += p.taps[p.cls[c].tap0 + t].kchunks;
      const int per_class = total_tiles / p.num_classes;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        const int c_idx = p.num_classes > 1 ? tile / per_class : 0;
        const int total_kb = c_idx == 0 ? kb_cls[0] : (c_idx == 1 ? kb_cls[1] : (c_idx == 2 ? kb_cls[2] : kb_cls[3]));
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
      int kb_cls[4] = {0, 0, 0, 0};
      for (int c = 0; c < p.num_classes; ++c)
        for (int t = 0; t < p.cls[c].ntaps; ++t) kb_cls[c] += p.taps[p.cls[c].tap0 + t].kchunks;
      const int per_class = total_tiles / p.num_classes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        const int c_idx = p.num_classes > 1 ? tile / per_class : 0;
        const int total_kb = c_idx == 0 ? kb_cls[0] : (c_idx == 1 ? kb_cls[1] : (c_idx == 2 ? kb_cls[2] : kb_cls[3]));
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
   }
  }

  if (threadIdx.x == kEpiThread0) trace_mark(p, 6);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p, 7);
  if (threadIdx.x == 0 && p.trace != nullptr) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.trace[blockIdx.x * 16 + 8] = static_cast<long long>(g);
  }
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
// NTAPS: taps with resident weights. 9 = the 3x3 layers (18-row patches); 12 = the stem on its
// space-to-depth input (4 x 2 / 4 x 3 taps over cells, 19-row patches, see plan_stem_fwd).
template <int CHUNKS, int NTAPS = 9>
struct HaloCfg {
  static constexpr int kBlockN = 64;
  static constexpr int kPatchRows = (NTAPS > 9 ? 19 : 18) * 10;
  static constexpr int kPatchSlot = (kPatchRows * 128 + 1023) & ~1023;   // 23552 / 24576
  static constexpr int kWBytes = NTAPS * CHUNKS * kBlockN * 128;     // resident weights
  static constexpr int kSlots = (CHUNKS == 1 && NTAPS <= 9) ? 4 : 3; // patch ring (per chunk)
  static constexpr int kTmemCols = 2 * kBlockN;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes =
      kWBytes + kSlots * kPatchSlot + StageCfg<kBlockN>::kBytes + kBarBytes + 3 * kBlockN * 4 +
      kStatScratchBytes + 1024;
};

template <int CHUNKS, int MODE, int NTAPS = 9>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ ConvParams p) {
  using Cfg = HaloCfg<CHUNKS, NTAPS>;
  constexpr int BLOCK_N = Cfg::kBlockN;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                         // [chunk][tap][64 cout][64 cin] bf16, swizzled
  uint8_t* s_patch = smem + Cfg::kWBytes;      // kSlots x kPatchSlot
  uint8_t* slabs = s_patch + Cfg::kSlots * Cfg::kPatchSlot;  // output staging ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + StageCfg<BLOCK_N>::kBytes);
  uint64_t* empty_bar = full_bar + Cfg::kSlots;
  uint64_t* tfull_bar = empty_bar + Cfg::kSlots;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint64_t* sfull_bar = w_bar + 1;
  uint64_t* sempty_bar = sfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sempty_bar + 2);
  float* s_sum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;
  float* s_scr = s_x2 + BLOCK_N;

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
      mbar_init(&sfull_bar[s], kEpiThreads);
      mbar_init(&sempty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmOut);
    conv_prefetch_next(p);
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

  if (warp >= kStatWarp0) {
    setmaxnreg_inc<kRegsStat>();
    conv_stats<BLOCK_N, 1, MODE>(p, &tmOut, nullptr, nullptr, slabs, sfull_bar, sempty_bar, s_sum, s_sq, s_x2, s_scr, 0,
                           blockIdx.x, gridDim.x, total_tiles, warp - kStatWarp0, lane);
  } else if (warp >= kEpiWarp0) {
    setmaxnreg_dec<kRegsEpi>();
    conv_epilogue<BLOCK_N, 1, MODE>(p, tmem_base, tfull_bar, tempty_bar, slabs, sfull_bar, sempty_bar,
                              s_sum, s_sq, nullptr, nullptr, 0, blockIdx.x, gridDim.x, total_tiles, warp, lane);
  } else {
   setmaxnreg_dec<kRegsCtl>();
   if (warp == 0) {
    if (elect_one()) {
      // resident weights of this CTA's channel block: CHUNKS x 9 boxes {64 cin, 64 cout, 1}
      mbar_expect_tx(w_bar, p.num_taps * CHUNKS * (BLOCK_N * 128));
      for (int kc = 0; kc < CHUNKS; ++kc)
        for (int t = 0; t < p.num_taps; ++t)
          bulk_load(s_w + (kc * NTAPS + t) * (BLOCK_N * 128),
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
          mbar_expect_tx(&full_bar[slot], p.patch_bytes);
          tma_load_5d(s_patch + slot * Cfg::kPatchSlot, &tmA, &full_bar[slot], kc * 64,
                      w0 + p.patch_dx, 0, h0 + p.patch_dy, b0);
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
        uint32_t accum = 0u;   // the tile's first MMA overwrites the accumulator
        for (int kc = 0; kc < CHUNKS; ++kc) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint32_t patch = smem_u32(s_patch + slot * Cfg::kPatchSlot);
#pragma unroll 1
          for (int t = 0; t < p.num_taps; ++t) {
            // tap (dy, dx): patch rows start at (1+dy)*10 + (1+dx); tile row h -> +10 rows
            const int start_row = (1 + p.taps[t].d3) * 10 + (1 + p.taps[t].d1);
            const uint64_t adesc = make_smem_desc(patch + start_row * 128, 16, 1280);
            const uint64_t bdesc =
                make_smem_desc(smem_u32(s_w + (kc * NTAPS + t) * (BLOCK_N * 128)), 16, 1024);
            const int kskip = NTAPS > 9 ? p.taps[t].kskip : 0;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              if (NTAPS > 9 && ((kskip >> k) & 1)) continue;
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
              accum = 1u;
            }
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
   }
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
      kWStages * kWTile + kSlots * kPatchSlot + StageCfg<BLOCK_N>::kBytes + kBarBytes +
      3 * BLOCK_N * 4 + kStatScratchBytes + 1024;
};

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_halo_stream_kernel(const __grid_constant__ CUtensorMap tmA,
                           const __grid_constant__ CUtensorMap tmOut,
                           const __grid_constant__ ConvParams p) {
  using Cfg = HaloStreamCfg<BLOCK_N>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                                   // kWStages x [BLOCK_N][64]
  uint8_t* s_patch = smem + Cfg::kWStages * Cfg::kWTile;  // kSlots x kPatchSlot
  uint8_t* slabs = s_patch + Cfg::kSlots * Cfg::kPatchSlot;  // output staging ring
  uint64_t* pfull = reinterpret_cast<uint64_t*>(slabs + StageCfg<BLOCK_N>::kBytes);
  uint64_t* pempty = pfull + Cfg::kSlots;
  uint64_t* wfull = pempty + Cfg::kSlots;
  uint64_t* wempty = wfull + Cfg::kWStages;
  uint64_t* tfull_bar = wempty + Cfg::kWStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sfull_bar = tempty_bar + 2;
  uint64_t* sempty_bar = sfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sempty_bar + 2);
  float* s_sum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pfull) + Cfg::kBarBytes);
  float* s_sq = s_sum + BLOCK_N;
  float* s_x2 = s_sq + BLOCK_N;
  float* s_scr = s_x2 + BLOCK_N;

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
      mbar_init(&sfull_bar[s], kEpiThreads);
      mbar_init(&sempty_bar[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmOut);
    conv_prefetch_next(p);
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

  if (warp >= kStatWarp0) {
    setmaxnreg_inc<kRegsStat>();
    conv_stats<BLOCK_N, 1, MODE>(p, &tmOut, nullptr, nullptr, slabs, sfull_bar, sempty_bar, s_sum, s_sq, s_x2, s_scr, 0,
                           blockIdx.x, gridDim.x, total_tiles, warp - kStatWarp0, lane);
  } else if (warp >= kEpiWarp0) {
    setmaxnreg_dec<kRegsEpi>();
    conv_epilogue<BLOCK_N, 1, MODE>(p, tmem_base, tfull_bar, tempty_bar, slabs, sfull_bar, sempty_bar,
                              s_sum, s_sq, nullptr, nullptr, 0, blockIdx.x, gridDim.x, total_tiles, warp, lane);
  } else {
   setmaxnreg_dec<kRegsCtl>();
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
          tma_load_5d(s_patch + slot * Cfg::kPatchSlot, &tmA, &pfull[slot], kc * 64,
                      w0 + p.patch_dx, 0, h0 + p.patch_dy, b0);
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
   }
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
  int num_pairs;        // unit groups: ceil(num_units / 2), or / 4 for the 256-wide kernel
  int splits;           // pixel-range splits per (unit group, n_tile)
  int kchunks;          // Cin / 64
  int num_taps;
  ConvTap taps[kMaxTaps];
  int cin, cout;
  int row_limit;        // rows (ci within unit) >= row_limit are not written (stem pad)
  float* dw;            // [taps][cout][cin] fp32, accumulated
  int dbg;              // diagnostics (VPD_WGRAD_DBG): bit 0 skip the atomics
};

// BLOCK_N = 256 (256 / 512 output channels): per unit pair the dY tile is read once for 256
// output channels instead of once per 128 (48 KB instead of 64 KB of operands per 64 pixels x
// 256 channels). Its pipeline stages hold 64 pixels (K = 64) instead of 128: 48 KB x 4 stages.
// kUnits = 4 (two M = 128 accumulators filling all 512 TMEM columns, 64 KB per 8 MMAs) is
// implemented by the loops below and measured SLOWER (21.2 vs 17.4 us per 256 -> 256 layer): a
// CTA has one work item, so the fp32 red.global epilogue - 148 x the accumulator size per
// launch, whatever the split - is exposed, and it doubles with the accumulator.
template <int BLOCK_N>
struct WgradCfg {
  static constexpr int kUnits = 2;                            // (tap, chunk) units per work item
  static constexpr int kPix = BLOCK_N == 256 ? 64 : kBlockM;  // pixels (K) per pipeline stage
  static constexpr int kBoxBytes = kPix * 64 * 2;             // one {64 ch, kPix pixels} TMA box
  static constexpr int kABytes = kUnits * kBoxBytes;
  static constexpr int kBBytes = (BLOCK_N / 64) * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 128 ? 3 : 4;
  // accumulator stages: two (the epilogue of an item overlaps the next item's MMAs) when they fit
  static constexpr int kAccCols = (kUnits / 2) * BLOCK_N;
  static constexpr int kAccStages = 2 * kAccCols <= 512 ? 2 : 1;
  static constexpr int kTmemCols = kAccStages * kAccCols;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgradThreads, 3)
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
        int u[Cfg::kUnits];
#pragma unroll
        for (int j = 0; j < Cfg::kUnits; ++j) {
          u[j] = Cfg::kUnits * pair + j;
          if (u[j] >= p.num_units) u[j] = Cfg::kUnits * pair;  // dummy unit (result discarded)
        }
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
          for (int j = 0; j < Cfg::kUnits; ++j) {
            const ConvTap tap = p.taps[u[j] / p.kchunks];
            const int kc = u[j] % p.kchunks;
            tma_load_5d(sa + j * Cfg::kBoxBytes, &tmX, &full_bar[stage], tap.c0 + kc * 64,
                        w0 + tap.d1, tap.d2, h0 + tap.d3, b0);
          }
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_5d(sb + j * Cfg::kBoxBytes, &tmDY, &full_bar[stage],
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
        const uint32_t d_tmem = tmem_base + as * Cfg::kAccCols;
        for (int pt = pt_beg; pt < pt_end; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          // MN-major: LBO = bytes between 64-wide M/N blocks (one TMA box),
          // SBO = bytes between 8-pixel groups along K
          const uint64_t bdesc = make_smem_desc(sb, Cfg::kBoxBytes, 1024);
#pragma unroll
          for (int h = 0; h < Cfg::kUnits / 2; ++h) {   // unit pairs = M = 128 accumulators
            const uint64_t adesc = make_smem_desc(sa + h * 2 * Cfg::kBoxBytes, Cfg::kBoxBytes, 1024);
#pragma unroll
            for (int k = 0; k < Cfg::kPix / kUmmaK; ++k) {
              // 16 pixels (rows of 128 B) per MMA -> 2048 B -> +128 in the address field
              umma_bf16(d_tmem + h * BLOCK_N, adesc + 128 * k, bdesc + 128 * k, idesc,
                        (pt != pt_beg || k != 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == Cfg::kAccStages) {
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
      const int rl = r & 63;
      const bool has_work = split * per_split < pix_tiles;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < Cfg::kUnits / 2; ++h) {
        const int unit = Cfg::kUnits * pair + 2 * h + (r >> 6);
        const bool valid = has_work && unit < p.num_units && rl < p.row_limit && !(p.dbg & 1);
        const int tap_i = valid ? unit / p.kchunks : 0;
        const int kc = valid ? unit % p.kchunks : 0;
        float* dst = p.dw + ((size_t)p.taps[tap_i].btap * p.cout + n_tile * BLOCK_N) * p.cin +
                     kc * 64 + rl;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::kAccCols +
                        h * BLOCK_N + c * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              atomicAdd(dst + (size_t)(c * 32 + j) * p.cin, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == Cfg::kAccStages) {
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

// ===========================================================================
// K2c-h: weight gradient of a stride-1 3x3 convolution with HALO REUSE.
//
// conv_wgrad_kernel above re-fetches the activation box for every (tap, chunk) unit and the
// dY box for every unit pair: per 128-pixel tile 9 x 16 KB + 5 x 16 KB of L2 -> SM traffic for
// 1280 MMA cycles at N = 64, four times what the ~50 B/cycle/SM delivery path sustains (the
// six 64 -> 64 layers ran at 460 TFLOP/s). Here, as in the forward halo kernel, ONE TMA
// patch {64 ch, 10, th+2 rows, tn images} brings the tile and its 1-pixel halo, and every tap
// is a shifted operand descriptor into it. Both operands stay MN-major (the reduction runs
// over pixels): A = patch rows, M = 2 taps x 64 input channels - the second tap is simply
// LBO = (its first patch row - the first tap's) x 128 B away; an MMA's K = 16 pixels are two
// 8-pixel row groups SBO = 10 patch rows (1280 B) apart; B = the dY tile [128 pixels][64 co].
// All nine taps of a (64-channel chunk, 64-cout block) accumulate side by side in TMEM
// (5 tap pairs x 64 columns) over the work item's pixel range and leave through coalesced
// red.global.add.f32 into the tap-major gradient arena. Per tile: 39 KB loaded for 40 MMAs.
// ===========================================================================
struct WgradHaloParams {
  int th, tn;                 // tile = 8 wide x th high x tn images, th * tn == 16
  int tiles_w, tiles_h, tiles_b;
  int kchunks, n_tiles;       // Cin / 64, Cout / 64
  int splits;                 // pixel-range splits per (chunk, cout block)
  int cin, cout;
  int patch_bytes;            // (th + 2) * 10 * tn * 128
  int kstep16[8];             // K step k (16 pixels): patch offset of its first row, in 16-byte units
  float* dw;                  // [9][cout][cin] fp32, accumulated
  int dbg;                    // diagnostics (VPD_WGRAD_DBG): 1 skip the atomics, 2 skip the MMAs, 4 skip the dY loads
  // taps: first patch row of tap t (3x3: kh * 10 + kw); processed two at a time
  int num_taps, num_pairs;
  int tap_row[12];
  int patch_dx, patch_dy;     // patch origin relative to the tile origin (-1, -1 for 3x3 'same')
  int dy_c0;                  // channel-coordinate offset of the dY view (stem: column class * 64)
  // stem (space-to-depth input, see plan_stem_fwd): stem_ntx = horizontal taps of the class
  // (2 or 3), else 0. Accumulator row e = (a*4 + q)*8 + c of tap (dy, dxb) is the gradient of
  // weight (kh = 2 dy + a, kw = 4 dxb + q - 2 cls, c): written to dw[kh][co][kw*8 + c], skipped
  // where that lies outside the 7 x 7 kernel
  int stem_ntx, stem_cls;
};

struct WgradHaloCfg {
  static constexpr int kPatchSlot = 25600;      // 200 rows (8x8 images, two per tile): max
  static constexpr int kDyBytes = kBlockM * 128;
  static constexpr int kStageBytes = kPatchSlot + kDyBytes;
  static constexpr int kStages = 4;
  static constexpr int kTmemCols = 512;         // <= 6 tap pairs x 64 columns, power of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
};

template <int NCO>   // couts per work item (accumulator columns per tap pair)
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const __grid_constant__ WgradHaloParams p) {
  using Cfg = WgradHaloCfg;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
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
  pdl_wait();

  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int per_split = (pix_tiles + p.splits - 1) / p.splits;
  const int total_items = p.kchunks * p.n_tiles * p.splits;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item % p.splits;
        const int n_tile = (item / p.splits) % p.n_tiles;
        const int kc = item / (p.splits * p.n_tiles);
        const int pt_end = min(pix_tiles, (split + 1) * per_split);
        for (int pt = split * per_split; pt < pt_end; ++pt) {
          int mt = pt;
          const int w0 = (mt % p.tiles_w) * 8;
          mt /= p.tiles_w;
          const int h0 = (mt % p.tiles_h) * p.th;
          const int b0 = (mt / p.tiles_h) * p.tn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sp = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], p.patch_bytes + ((p.dbg & 4) ? 0 : Cfg::kDyBytes));
          tma_load_5d(sp, &tmX, &full_bar[stage], kc * 64, w0 + p.patch_dx, 0, h0 + p.patch_dy, b0);
          if (!(p.dbg & 4))
            tma_load_5d(sp + Cfg::kPatchSlot, &tmDY, &full_bar[stage], p.dy_c0 + n_tile * 64, w0, 0,
                        h0, b0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      static_assert(NCO == 64, "5 tap pairs x NCO accumulator columns must fit 512 TMEM columns");
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, NCO, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item % p.splits;
        const int pt_beg = split * per_split;
        const int pt_end = min(pix_tiles, (split + 1) * per_split);
        mbar_wait(tempty_bar, tphase ^ 1);   // the previous item's accumulators have been drained
        tc_fence_after();
        for (int pt = pt_beg; pt < pt_end; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t patch = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t bdesc = make_smem_desc(patch + Cfg::kPatchSlot, kBlockM * 128, 1024);
#pragma unroll 1
          for (int j = 0; j < ((p.dbg & 2) ? 0 : p.num_pairs); ++j) {
            // taps 2j, 2j+1 in patch-row order; an odd tap count pairs the last one with the
            // row after it - that half of the accumulator is discarded
            const int row_a = p.tap_row[2 * j];
            const int row_b = 2 * j + 1 < p.num_taps ? p.tap_row[2 * j + 1] : row_a + 1;
            const uint64_t adesc =
                make_smem_desc(patch + row_a * 128, (row_b - row_a) * 128, 1280);
            const uint32_t d_tmem = tmem_base + j * 64;
#pragma unroll
            for (int k = 0; k < kBlockM / kUmmaK; ++k)
              umma_bf16(d_tmem, adesc + static_cast<uint64_t>(p.kstep16[k]), bdesc + 128 * k, idesc,
                        (pt != pt_beg || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull_bar);
        tphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;       // TMEM lane: rows 0-63 first tap of a pair, 64-127 second
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int split = item % p.splits;
      const int n_tile = (item / p.splits) % p.n_tiles;
      const int kc = item / (p.splits * p.n_tiles);
      const bool has_work = split * per_split < pix_tiles;
      mbar_wait(tfull_bar, tphase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < p.num_pairs; ++j) {
        const int tap = 2 * j + (r >> 6);
        bool valid = has_work && tap < p.num_taps && !(p.dbg & 1);
        float* dst = p.dw + ((size_t)(valid ? tap : 0) * p.cout + n_tile * 64) * p.cin + kc * 64 +
                     (r & 63);
        if (p.stem_ntx > 0) {
          const int e = r & 63;
          const int kh = 2 * (tap / p.stem_ntx) + (e >> 5);
          const int kw = 4 * (tap % p.stem_ntx) + ((e >> 3) & 3) - 2 * p.stem_cls;
          valid = valid && kh < 7 && kw >= 0 && kw < 7;
          dst = p.dw + (valid ? kh * 4096 + kw * 8 + (e & 7) : 0);
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 64 + c * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              atomicAdd(dst + (size_t)(c * 32 + i) * p.cin, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


}  // namespace vpd
