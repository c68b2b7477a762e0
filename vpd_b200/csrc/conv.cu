#include "conv.h"

#include <stdlib.h>
#include <string.h>

namespace vpd {

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static int floor_pow2(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// 128 output pixels per tile as tw x th x tn
static void tile_geometry(int Ho, int Wo, int N, ConvParams* p) {
  int tw = floor_pow2(Wo < 128 ? Wo : 128);
  int th = floor_pow2(Ho);
  if (th > 128 / tw) th = 128 / tw;
  int tn = 128 / (tw * th);
  p->tw = tw;
  p->th = th;
  p->tn = tn;
  p->tiles_w = (Wo + tw - 1) / tw;
  p->tiles_h = (Ho + th - 1) / th;
  p->tiles_b = (N + tn - 1) / tn;
  p->batch = N;
  p->out_h = Ho;
  p->out_w = Wo;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Channel-block width. The kernels are bound by L2->SM operand traffic
// (~42 B/cycle/SM), so wider tiles are better as long as enough tiles remain to
// fill the SMs.
static int pick_block_n(int cout, int m_tiles) {
  if (cout % 128 != 0) return 64;
  static const int allow256 = env_int("VPD_BN256", 1);
  if (allow256 && cout % 256 == 0 && (long long)m_tiles * (cout / 256) >= 96) return 256;
  return 128;
}

static int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }
static int mod2(int v) { return ((v % 2) + 2) % 2; }

static void finish_launch(ConvLaunch* L, int cout, bool stats) {
  const int m_tiles = L->p.tiles_w * L->p.tiles_h * L->p.tiles_b;
  L->block_n = pick_block_n(cout, m_tiles);
  L->p.n_tiles = cout / L->block_n;
  L->p.cout = cout;
  // pair mode (tcgen05 cta_group::2): two CTAs share one 256-row MMA and each stages only
  // half of the weight tile, halving the shared-memory traffic per MAC
  static const int want_pair = env_int("VPD_PAIR", 0);
  L->cluster = (want_pair && L->block_n >= 128 && m_tiles >= 2) ? 2 : 1;
  const int cs = L->cluster;
  const int ncls = L->p.num_classes > 1 ? L->p.num_classes : 1;
  const int items = ((m_tiles + cs - 1) / cs) * L->p.n_tiles * ncls;  // cluster-level work items
  int clusters = device_sm_count() / cs;
  if (clusters > items) clusters = items;
  // keep the channel block fixed per CTA so BN statistics stay in shared memory
  if (stats && clusters >= L->p.n_tiles) clusters -= clusters % L->p.n_tiles;
  L->grid = clusters * cs;
  static const int direct_on = env_int("VPD_DIRECT_SLABS", 1);
  L->p.single_tile = (direct_on && clusters >= items) ? 1 : 0;
}

// weights [taps][rows][kdim] bf16 -> 3-D map, box {64, block_n, 1}
static int weight_map(CUtensorMap* m, const __nv_bfloat16* w, int taps, int rows, int kdim,
                      int box_rows) {
  uint64_t dims[3] = {(uint64_t)kdim, (uint64_t)rows, (uint64_t)taps};
  uint64_t str[3] = {2, (uint64_t)kdim * 2, (uint64_t)rows * kdim * 2};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return encode_tmap_bf16(m, w, 3, dims, str, box, true);
}

// NHWC activation [N][H][W][C] viewed for a conv of the given stride
static int act_map(CUtensorMap* m, const __nv_bfloat16* x, int N, int H, int W, int C, int stride,
                   const ConvParams& p) {
  uint32_t box[5] = {64, (uint32_t)p.tw, 1, (uint32_t)p.th, (uint32_t)p.tn};
  if (stride == 1) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, 1, (uint64_t)H, (uint64_t)N};
    uint64_t str[5] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)W * C * 2,
                       (uint64_t)H * W * C * 2};
    return encode_tmap_bf16(m, x, 5, dims, str, box, true);
  }
  // stride 2: even/odd columns fold into the channel axis, even/odd rows get
  // their own axis, so every tap is a dense box
  uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)N};
  uint64_t str[5] = {2, (uint64_t)2 * C * 2, (uint64_t)W * C * 2, (uint64_t)2 * W * C * 2,
                     (uint64_t)H * W * C * 2};
  return encode_tmap_bf16(m, x, 5, dims, str, box, true);
}

// view of the output tensor [N][H][W][C] for the slab stores: same box as the A loads
// (64 channels x the tile's pixels); `stride` 2 addresses one pixel-parity class
static int out_map(ConvLaunch* L, const __nv_bfloat16* out, int N, int H, int W, int C, int stride,
                   int c0, int d2) {
  if (L->p.num_classes <= 1) {   // ordinary convolution: one output class over all taps
    L->p.num_classes = 1;
    L->p.cls[0].tap0 = 0;
    L->p.cls[0].ntaps = L->p.num_taps;
    L->p.cls[0].out_c0 = c0;
    L->p.cls[0].out_d2 = d2;
    L->p.cls[0].base = 0;
  }
  if (act_map(&L->o, out, N, H, W, C, stride, L->p)) return -1;
  return 0;
}

static void set_weights(ConvParams* p, int src, const __nv_bfloat16* w, int rows, int kdim) {
  p->w[src] = w;
  p->w_kc[src] = kdim / 64;
  p->w_rb[src] = rows / 64;
}

// Re-plan a stride-1 3x3 'same' convolution (forward or dgrad) for the halo-reuse
// kernel: 8x16 pixel tiles, 64-wide channel blocks with resident weights.
// `x`: the tensor the taps read ([N][H][W][Cin_k]); `wt`: its tap-major weights
// [9][Cout_k][Cin_k]. Returns false when the shape does not qualify.
static thread_local bool g_plan_no_halo = false;  // set while planning wgrad helpers

static bool try_halo(ConvLaunch* L, const __nv_bfloat16* x, const __nv_bfloat16* wt, int N, int H,
                     int W, int cin_k, int cout_k, bool stats) {
  static const int enable = env_int("VPD_HALO", 1);
  if (g_plan_no_halo) return false;
  const int chunks = cin_k / 64;
  if (!enable || cin_k % 64 != 0 || cout_k % 64 != 0) return false;
  if (H % 16 != 0 || W % 8 != 0) return false;
  // 64 -> 64 layers keep their 72 KB of weights resident; wider layers stream the
  // (chunk, tap) weight tiles through a TMA ring with 128-wide channel blocks
  static const int stream_on = env_int("VPD_HALO_STREAM", 1);  // 128-channel 16x16 layers: -0.04 ms/step
  const bool resident = chunks == 1 && cout_k == 64;
  if (!resident && !(stream_on && cout_k % 128 == 0 && chunks <= 8)) return false;
  ConvParams& p = L->p;
  p.tw = 8;
  p.th = 16;
  p.tn = 1;
  p.tiles_w = W / 8;
  p.tiles_h = H / 16;
  p.tiles_b = N;
  p.batch = N;
  p.out_h = H;
  p.out_w = W;
  L->block_n = resident ? 64 : 128;
  L->cluster = 1;
  L->halo = resident ? 1 : 3;   // 1: resident-weight kernel, 3: streamed weights
  p.n_tiles = cout_k / L->block_n;
  p.cout = cout_k;
  for (int t = 0; t < 9; ++t) p.taps[t].kchunks = chunks;
  p.patch_dx = -1;
  p.patch_dy = -1;
  p.patch_bytes = 18 * 10 * 128;
  const int total = p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;
  int grid = device_sm_count();
  if (grid > total) grid = total;
  grid -= grid % p.n_tiles;  // fixed channel block per CTA (resident weights)
  if (grid <= 0) return false;
  L->grid = grid;
  (void)stats;
  uint64_t dims[5] = {(uint64_t)cin_k, (uint64_t)W, 1, (uint64_t)H, (uint64_t)N};
  uint64_t str[5] = {2, (uint64_t)cin_k * 2, (uint64_t)W * cin_k * 2, (uint64_t)W * cin_k * 2,
                     (uint64_t)H * W * cin_k * 2};
  uint32_t box[5] = {64, 10, 1, 18, 1};
  if (encode_tmap_bf16(&L->a0, x, 5, dims, str, box, true)) return false;
  set_weights(&p, 0, wt, cout_k, cin_k);
  set_weights(&p, 1, wt, cout_k, cin_k);
  L->a1 = L->a0;
  return true;
}

static void set_epilogue(ConvParams* p, const ConvEpilogue& e) {
  p->scale = e.scale;
  p->shift = e.shift;
  p->residual = e.residual;
  p->relu = e.relu;
  p->stats = e.stats;
}

int plan_conv_fwd(ConvLaunch* L, const ConvGeom& g, const __nv_bfloat16* x,
                  const __nv_bfloat16* w_tap, __nv_bfloat16* y, const ConvEpilogue& e) {
  memset(L, 0, sizeof(*L));
  L->w_ptr = w_tap;
  L->w_bytes = (long long)g.k * g.k * g.Cout * g.Cin * 2;
  VPD_REQUIRE(g.Cin % 64 == 0 && g.Cout % 64 == 0, "conv: Cin/Cout must be multiples of 64 (%d,%d)",
              g.Cin, g.Cout);
  VPD_REQUIRE(g.stride == 1 || g.stride == 2, "conv: stride %d unsupported", g.stride);
  VPD_REQUIRE(g.k * g.k <= kMaxTaps, "conv: kernel %d too large", g.k);
  if (g.stride == 2) VPD_REQUIRE(g.H % 2 == 0 && g.W % 2 == 0, "conv: stride 2 needs even H, W");
  const int Ho = g.Ho(), Wo = g.Wo();
  ConvParams& p = L->p;
  tile_geometry(Ho, Wo, g.N, &p);
  p.num_taps = g.k * g.k;
  for (int kh = 0; kh < g.k; ++kh)
    for (int kw = 0; kw < g.k; ++kw) {
      ConvTap& t = p.taps[kh * g.k + kw];
      const int u = kw - g.pad, v = kh - g.pad;
      if (g.stride == 1) {
        t.c0 = 0;
        t.d1 = u;
        t.d2 = 0;
        t.d3 = v;
      } else {
        t.c0 = mod2(u) * g.Cin;
        t.d1 = floordiv2(u);
        t.d2 = mod2(v);
        t.d3 = floordiv2(v);
      }
      t.src = 0;
      t.btap = kh * g.k + kw;
      t.kchunks = g.Cin / 64;
    }
  p.out = y;
  p.out_sn = (long long)Ho * Wo * g.Cout;
  p.out_sh = (long long)Wo * g.Cout;
  p.out_sw = g.Cout;
  set_epilogue(&p, e);
  if (g.k == 3 && g.stride == 1 && g.pad == 1 &&
      try_halo(L, x, w_tap, g.N, g.H, g.W, g.Cin, g.Cout, e.stats != nullptr))
    return out_map(L, y, g.N, Ho, Wo, g.Cout, 1, 0, 0);
  finish_launch(L, g.Cout, e.stats != nullptr);
  if (act_map(&L->a0, x, g.N, g.H, g.W, g.Cin, g.stride, p)) return -1;
  set_weights(&p, 0, w_tap, g.Cout, g.Cin);
  set_weights(&p, 1, w_tap, g.Cout, g.Cin);
  L->a1 = L->a0;
  if (out_map(L, y, g.N, Ho, Wo, g.Cout, 1, 0, 0)) return -1;
  return 0;
}

// The 7x7 / stride-2 / pad-3 stem on the space-to-depth input layout (common.cuh::
// stem_pixel_offset): output pixel (ho, wo = 2m + cls) reads padded rows 2ho + kh = cell rows
// ho + (kh >> 1) and padded columns 2wo + kw = 4m + 2 cls + kw = cell columns m + ((kw + 2 cls)
// >> 2), so each column-parity class `cls` is a STRIDE-1 convolution over cells with 4 x (2 +
// cls) taps of K = 64 (cell elements (a*4 + q)*8 + c <-> kh = 2 dy + a, kw = 4 dxb + q - 2 cls;
// combinations that fall outside the 7x7 kernel carry zero weights). Two launches of the
// resident-weight halo kernel (one aligned 19 x 10-cell patch per 16 x 8-pixel tile, every tap
// a shifted descriptor into it), writing the class's columns through a (2C, W/2) view of y.
// w_s2d: bf16 [8 taps of class 0][12 taps of class 1] 64 x 64 tiles (pack_stem_weight).
int stem_taps(int cls) { return 4 * (2 + cls); }

int plan_stem_fwd(ConvLaunch* L2, int N, int H, int W, const __nv_bfloat16* x_s2d,
                  const __nv_bfloat16* w_s2d, __nv_bfloat16* y, const ConvEpilogue& e) {
  VPD_REQUIRE(H % 32 == 0 && W % 32 == 0, "stem: H and W must be multiples of 32 (%d,%d)", H, W);
  const int Ho = H / 2, Wo = W / 2, Hs = stem_cells_h(H), Ws = stem_cells_w(W);
  for (int cls = 0; cls < 2; ++cls) {
    ConvLaunch* L = L2 + cls;
    memset(L, 0, sizeof(*L));
    const __nv_bfloat16* wc = w_s2d + (cls == 0 ? 0 : stem_taps(0) * 4096);
    L->w_ptr = wc;
    L->w_bytes = (long long)stem_taps(cls) * 4096 * 2;
    ConvParams& p = L->p;
    p.tw = 8;
    p.th = 16;
    p.tn = 1;
    p.tiles_w = (Wo / 2) / 8;
    p.tiles_h = Ho / 16;
    p.tiles_b = N;
    p.batch = N;
    p.out_h = Ho;
    p.out_w = Wo / 2;
    const int ntx = 2 + cls;
    p.num_taps = stem_taps(cls);
    for (int t = 0; t < p.num_taps; ++t) {
      ConvTap& tp = p.taps[t];
      tp.c0 = 0;
      tp.d3 = t / ntx - 1;   // the kernel reads patch row (1 + d3) * 10 + (1 + d1) = dy * 10 + dxb
      tp.d1 = t % ntx - 1;
      tp.d2 = 0;
      tp.src = 0;
      tp.btap = t;
      tp.kchunks = 1;
      // K step ks of a cell = row a = ks >> 1, column pair q = 2 (ks & 1), + 1: weight
      // (kh = 2 dy + a, kw = 4 dxb + q - 2 cls) lies outside the 7 x 7 kernel for BOTH columns
      // (or for the row) -> all-zero slice, not issued: 28 of 32 / 28 of 48 MMAs remain
      const int dy = t / ntx, dxb = t % ntx;
      tp.kskip = 0;
      for (int ks = 0; ks < 4; ++ks) {
        const int kh = 2 * dy + (ks >> 1);
        bool any = false;
        for (int q = 2 * (ks & 1); q < 2 * (ks & 1) + 2; ++q) {
          const int kw = 4 * dxb + q - 2 * cls;
          any = any || (kh < 7 && kw >= 0 && kw < 7);
        }
        if (!any) tp.kskip |= 1 << ks;
      }
    }
    p.patch_dx = 0;
    p.patch_dy = 0;
    p.patch_bytes = 19 * 10 * 128;
    p.out = y;
    p.out_sn = (long long)Ho * Wo * 64;
    p.out_sh = (long long)Wo * 64;
    p.out_sw = 128;                       // consecutive pixels of a class are two columns apart
    set_epilogue(&p, e);
    L->block_n = 64;
    L->cluster = 1;
    L->halo = 4;                          // resident-weight halo kernel, 12-tap configuration
    p.n_tiles = 1;
    p.cout = 64;
    p.num_classes = 1;
    p.cls[0].tap0 = 0;
    p.cls[0].ntaps = p.num_taps;
    p.cls[0].out_c0 = cls * 64;
    p.cls[0].out_d2 = 0;
    p.cls[0].base = cls * 64;
    const int total = p.tiles_w * p.tiles_h * p.tiles_b;
    L->grid = device_sm_count() < total ? device_sm_count() : total;
    uint64_t dims[5] = {64, (uint64_t)Ws, 1, (uint64_t)Hs, (uint64_t)N};
    uint64_t str[5] = {2, 128, (uint64_t)Ws * 128, (uint64_t)Ws * 128, (uint64_t)Hs * Ws * 128};
    uint32_t box[5] = {64, 10, 1, 19, 1};
    if (encode_tmap_bf16(&L->a0, x_s2d, 5, dims, str, box, true)) return -1;
    L->a1 = L->a0;
    set_weights(&p, 0, wc, 64, 64);
    set_weights(&p, 1, wc, 64, 64);
    // y [N][Ho][Wo][64] seen as [N][Ho][Wo/2][2 x 64]: channel coordinate cls * 64 + c
    uint64_t odims[5] = {128, (uint64_t)Wo / 2, 1, (uint64_t)Ho, (uint64_t)N};
    uint64_t ostr[5] = {2, 256, (uint64_t)Wo * 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128};
    uint32_t obox[5] = {64, 8, 1, 16, 1};
    if (encode_tmap_bf16(&L->o, y, 5, odims, ostr, obox, true)) return -1;
    }
  return 0;
}

int plan_conv_dgrad(ConvLaunch* Ls, int* count, const ConvGeom& g, const __nv_bfloat16* dy,
                    const __nv_bfloat16* wT_tap, __nv_bfloat16* dx,
                    const __nv_bfloat16* residual, const __nv_bfloat16* dy_ds,
                    const __nv_bfloat16* wT_ds, int cout_ds, const ConvBwdFuse* fuse) {
  VPD_REQUIRE(g.Cin % 64 == 0 && g.Cout % 64 == 0, "dgrad: channels must be multiples of 64");
  // attach the fused BN-backward reduction; `base` = element offset of this launch's
  // first output pixel (stride-2 parity classes)
  auto set_fuse = [&](ConvParams* p, long long base) {
    p->bnb = 0;
    if (fuse == nullptr || fuse->nb == 0) return;
    p->bnb = fuse->nb;
    p->bmask = fuse->mask + base / 8;
    for (int b = 0; b < fuse->nb; ++b) {
      p->by[b] = fuse->y[b] + base;
      p->bmean[b] = fuse->mean[b];
      p->brstd[b] = fuse->rstd[b];
      p->bsums[b] = fuse->sums[b];
    }
  };
  const int Ho = g.Ho(), Wo = g.Wo();
  *count = 0;
  VPD_REQUIRE(fuse == nullptr || fuse->nb == 0 || (long long)g.N * g.H * g.W * g.Cin < (1LL << 31),
              "dgrad: the fused BN-backward reduction indexes with 32 bits (%lld elements)",
              (long long)g.N * g.H * g.W * g.Cin);
  if (g.stride == 1) {
    VPD_REQUIRE(g.k == 2 * g.pad + 1, "dgrad: stride-1 conv must be 'same' padded");
    ConvLaunch* L = &Ls[0];
    memset(L, 0, sizeof(*L));
    L->w_ptr = wT_tap;
    L->w_bytes = (long long)g.k * g.k * g.Cout * g.Cin * 2;
    ConvParams& p = L->p;
    tile_geometry(g.H, g.W, g.N, &p);
    p.num_taps = g.k * g.k;
    for (int kh = 0; kh < g.k; ++kh)
      for (int kw = 0; kw < g.k; ++kw) {
        ConvTap& t = p.taps[kh * g.k + kw];
        t.c0 = 0;
        t.d1 = g.pad - kw;
        t.d2 = 0;
        t.d3 = g.pad - kh;
        t.src = 0;
        t.btap = kh * g.k + kw;
        t.kchunks = g.Cout / 64;
      }
    p.out = dx;
    p.residual = residual;
    p.out_sn = (long long)g.H * g.W * g.Cin;
    p.out_sh = (long long)g.W * g.Cin;
    p.out_sw = g.Cin;
    set_fuse(&p, 0);
    if (g.k == 3 && try_halo(L, dy, wT_tap, g.N, Ho, Wo, g.Cout, g.Cin, p.bnb > 0)) {
      *count = 1;
      return out_map(L, dx, g.N, g.H, g.W, g.Cin, 1, 0, 0);
    }
    finish_launch(L, g.Cin, p.bnb > 0);
    if (act_map(&L->a0, dy, g.N, Ho, Wo, g.Cout, 1, p)) return -1;
    set_weights(&p, 0, wT_tap, g.Cin, g.Cout);
    set_weights(&p, 1, wT_tap, g.Cin, g.Cout);
    L->a1 = L->a0;
    *count = 1;
    return out_map(L, dx, g.N, g.H, g.W, g.Cin, 1, 0, 0);
  }
  VPD_REQUIRE(g.stride == 2 && g.k == 3 && g.pad == 1, "dgrad: only 3x3/2 pad 1 strided convs");
  VPD_REQUIRE(g.H == 2 * Ho && g.W == 2 * Wo, "dgrad: stride 2 needs even input dims");
  // dx[2i+a, 2j+b] = sum over kh = a+1 (mod 2), kw = b+1 (mod 2) of
  //                  dy[i + (a+1-kh)/2, j + (b+1-kw)/2] * w[kh,kw]
  // The four output-parity classes (a, b) share everything but their tap subsets and their
  // position in dx, so they are four OUTPUT CLASSES of a single launch (four separate
  // launches of 1-4 taps each spent most of their time in launch/drain overhead).
  ConvLaunch* L = &Ls[0];
  memset(L, 0, sizeof(*L));
  L->w_ptr = wT_tap;
  L->w_bytes = (long long)g.k * g.k * g.Cout * g.Cin * 2;
  ConvParams& p = L->p;
  tile_geometry(Ho, Wo, g.N, &p);
  int nt = 0, nc = 0;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      ConvParams::OutClass& c = p.cls[nc++];
      c.tap0 = nt;
      for (int kh = 0; kh < 3; ++kh) {
        if ((a + 1 - kh) % 2 != 0) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if ((b + 1 - kw) % 2 != 0) continue;
          ConvTap& t = p.taps[nt++];
          t.c0 = 0;
          t.d1 = (b + 1 - kw) / 2;
          t.d2 = 0;
          t.d3 = (a + 1 - kh) / 2;
          t.src = 0;
          t.btap = kh * 3 + kw;
          t.kchunks = g.Cout / 64;
        }
      }
      if (a == 0 && b == 0 && dy_ds != nullptr) {   // 1x1/2 downsample branch: extra tap of (0,0)
        ConvTap& t = p.taps[nt++];
        t.c0 = 0;
        t.d1 = 0;
        t.d2 = 0;
        t.d3 = 0;
        t.src = 1;
        t.btap = 0;
        t.kchunks = cout_ds / 64;
      }
      c.ntaps = nt - c.tap0;
      // class (a, b) of dx through the stride-2 view: channel coordinate b*Cin + ch,
      // parity coordinate a, tile coordinates in units of the half-resolution grid
      c.out_c0 = b * g.Cin;
      c.out_d2 = a;
      c.base = ((long long)a * g.W + b) * g.Cin;
    }
  VPD_REQUIRE(nt <= kMaxTaps, "dgrad: too many taps (%d)", nt);
  p.num_taps = nt;
  p.num_classes = 4;
  p.out = dx;
  p.residual = residual;
  p.out_sn = (long long)g.H * g.W * g.Cin;
  p.out_sh = (long long)2 * g.W * g.Cin;
  p.out_sw = (long long)2 * g.Cin;
  set_fuse(&p, 0);
  finish_launch(L, g.Cin, p.bnb > 0);
  if (act_map(&L->a0, dy, g.N, Ho, Wo, g.Cout, 1, p)) return -1;
  set_weights(&p, 0, wT_tap, g.Cin, g.Cout);
  set_weights(&p, 1, wT_tap, g.Cin, g.Cout);
  if (dy_ds != nullptr) {
    if (act_map(&L->a1, dy_ds, g.N, Ho, Wo, cout_ds, 1, p)) return -1;
    set_weights(&p, 1, wT_ds, g.Cin, cout_ds);
  } else {
    L->a1 = L->a0;
  }
  if (out_map(L, dx, g.N, g.H, g.W, g.Cin, 2, 0, 0)) return -1;
  *count = 1;
  return 0;
}

static int conv_mode(const ConvLaunch& L) {
  return L.p.stats != nullptr ? 1 : (L.p.bnb == 1 ? 2 : (L.p.bnb == 2 ? 3 : 0));
}

template <int BN, int CS, int MODE>
static int launch_bn_impl(const ConvLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VPD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, CS, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        ConvCfg<BN>::kSmemBytes));
    attr_set = true;
  }
  VPD_CHECK_CUDA(launch_kernel_cluster(CS, conv_igemm_kernel<BN, CS, MODE>, dim3(L.grid),
                                       dim3(kConvThreads), ConvCfg<BN>::kSmemBytes, stream, L.a0,
                                       L.a1, L.o, L.p));
  VPD_LAUNCHED(1);
  return 0;
}

template <int BN, int CS>
static int launch_bn(const ConvLaunch& L, cudaStream_t stream) {
  if (CS != 1) return launch_bn_impl<BN, CS, -1>(L, stream);   // opt-in pair mode: generic
  switch (conv_mode(L)) {
    case 1: return launch_bn_impl<BN, 1, 1>(L, stream);
    case 2: return launch_bn_impl<BN, 1, 2>(L, stream);
    case 3: return launch_bn_impl<BN, 1, 3>(L, stream);
    default: return launch_bn_impl<BN, 1, 0>(L, stream);
  }
}

template <int CHUNKS, int MODE, int NTAPS>
static int launch_halo_impl(const ConvLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VPD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<CHUNKS, MODE, NTAPS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        HaloCfg<CHUNKS, NTAPS>::kSmemBytes));
    attr_set = true;
  }
  VPD_CHECK_CUDA(launch_kernel(conv3x3_halo_kernel<CHUNKS, MODE, NTAPS>, dim3(L.grid),
                               dim3(kConvThreads), HaloCfg<CHUNKS, NTAPS>::kSmemBytes, stream, L.a0,
                               L.o, L.p));
  VPD_LAUNCHED(1);
  return 0;
}
template <int CHUNKS, int NTAPS = 9>
static int launch_halo(const ConvLaunch& L, cudaStream_t stream) {
  if (NTAPS > 9)   // the stem: forward only (statistics in training, folded BN + ReLU in eval)
    return conv_mode(L) == 1 ? launch_halo_impl<CHUNKS, 1, NTAPS>(L, stream)
                             : launch_halo_impl<CHUNKS, 0, NTAPS>(L, stream);
  switch (conv_mode(L)) {
    case 1: return launch_halo_impl<CHUNKS, 1, 9>(L, stream);
    case 2: return launch_halo_impl<CHUNKS, 2, 9>(L, stream);
    case 3: return launch_halo_impl<CHUNKS, 3, 9>(L, stream);
    default: return launch_halo_impl<CHUNKS, 0, 9>(L, stream);
  }
}
template <int MODE>
static int launch_halo_stream_impl(const ConvLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VPD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_stream_kernel<128, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        HaloStreamCfg<128>::kSmemBytes));
    attr_set = true;
  }
  VPD_CHECK_CUDA(launch_kernel(conv3x3_halo_stream_kernel<128, MODE>, dim3(L.grid),
                               dim3(kConvThreads), HaloStreamCfg<128>::kSmemBytes, stream, L.a0,
                               L.o, L.p));
  VPD_LAUNCHED(1);
  return 0;
}

static long long* g_conv_trace = nullptr;
void set_conv_trace(long long* dev_buf) { g_conv_trace = dev_buf; }

int launch_conv(const ConvLaunch& L0, cudaStream_t stream) {
  if (L0.grid <= 0) return 0;
  ConvLaunch traced = L0;
  if (g_conv_trace != nullptr) traced.p.trace = g_conv_trace;
  static const int dbg_skip = env_int("VPD_DBG_SKIP", 0);
  if (dbg_skip != 0) traced.p.dbg = dbg_skip;
  {   // divisors of the tile decode (decode_tile)
    ConvParams& q = traced.p;
    const int cs = L0.cluster > 1 ? L0.cluster : 1;
    const int m_tiles = q.tiles_w * q.tiles_h * q.tiles_b;
    q.fd_class = fd_make((uint32_t)(((m_tiles + cs - 1) / cs) * q.n_tiles));
    q.fd_ntiles = fd_make((uint32_t)q.n_tiles);
    q.fd_tw = fd_make((uint32_t)q.tiles_w);
    q.fd_th = fd_make((uint32_t)q.tiles_h);
  }
  const ConvLaunch& L = traced;
  if (L.halo == 1) return launch_halo<1>(L, stream);
  if (L.halo == 4) return launch_halo<1, 12>(L, stream);
  if (L.halo == 3) {
    switch (conv_mode(L)) {
      case 1: return launch_halo_stream_impl<1>(L, stream);
      case 2: return launch_halo_stream_impl<2>(L, stream);
      case 3: return launch_halo_stream_impl<3>(L, stream);
      default: return launch_halo_stream_impl<0>(L, stream);
    }
  }
  if (L.cluster == 2) {
    switch (L.block_n) {
      case 128: return launch_bn<128, 2>(L, stream);
      case 256: return launch_bn<256, 2>(L, stream);
    }
  } else {
    switch (L.block_n) {
      case 64: return launch_bn<64, 1>(L, stream);
      case 128: return launch_bn<128, 1>(L, stream);
      case 256: return launch_bn<256, 1>(L, stream);
    }
  }
  set_error("launch_conv: bad block_n %d", L.block_n);
  return -1;
}

// ------------------------------------------------------------------------ wgrad
static void finish_wgrad(WgradLaunch* L, int cin, int cout, int num_taps, int kchunks,
                         int row_limit, float* dw) {
  WgradParams& p = L->p;
  if (L->block_n != 256) L->block_n = cout % 128 == 0 ? 128 : 64;
  p.n_tiles = cout / L->block_n;
  p.kchunks = kchunks;
  p.num_taps = num_taps;
  p.num_units = num_taps * kchunks;
  const int group = 2;                           // WgradCfg::kUnits
  p.dbg = env_int("VPD_WGRAD_DBG", 0);
  p.num_pairs = (p.num_units + group - 1) / group;
  p.cin = cin;
  p.cout = cout;
  p.row_limit = row_limit;
  p.dw = dw;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  int splits = device_sm_count() / (p.num_pairs * p.n_tiles);
  if (splits < 1) splits = 1;
  if (splits > pix_tiles) splits = pix_tiles;
  // no empty splits: per_split = ceil(pix/splits) must leave the last one non-empty
  while (splits > 1 && (splits - 1) * ((pix_tiles + splits - 1) / splits) >= pix_tiles) --splits;
  p.splits = splits;
  const int items = p.num_pairs * p.n_tiles * splits;
  L->grid = items < device_sm_count() ? items : device_sm_count();
}

static void copy_geometry(WgradParams* w, const ConvParams& c) {
  w->tw = c.tw;
  w->th = c.th;
  w->tn = c.tn;
  w->tiles_w = c.tiles_w;
  w->tiles_h = c.tiles_h;
  w->tiles_b = c.tiles_b;
  for (int i = 0; i < kMaxTaps; ++i) w->taps[i] = c.taps[i];
}

// Halo-reuse plan for the weight gradient of a stride-1 3x3 'same' convolution (see
// conv_wgrad_halo_kernel): images of 8k x 16m pixels use 8 x 16 tiles, 8 x 8 images two per tile.
static bool try_wgrad_halo(WgradLaunch* L, const ConvGeom& g, const __nv_bfloat16* x,
                           const __nv_bfloat16* dy, float* dw) {
  // Opt-in (VPD_WGRAD_HALO=1): parity-tested, but measured no faster than conv_wgrad_kernel at
  // batch 256 (tests/diag_wgrad.py: 64->64 29.3 vs 29.0 us, 128->128 26.5 vs 19.0, 256->256
  // 26.2 vs 18.0). The L2 -> SM traffic does fall 6x, but with N = 64 accumulator columns per
  // tap pair (5 pairs must share the 512 TMEM columns) every MMA reads 6 KB of operands for 32
  // cycles of math - the shared-memory bandwidth bound (22 us without the atomics) - and a work
  // item now covers all nine taps of 1/148 of the pixels, so 5x as many fp32 atomics leave
  // through L2 (4-7 us, nothing to overlap them with).
  static const int enable = env_int("VPD_WGRAD_HALO", 0);
  if (!enable || g.k != 3 || g.stride != 1 || g.pad != 1) return false;
  if (g.Cin % 64 != 0 || g.Cout % 64 != 0 || g.W % 8 != 0) return false;
  WgradHaloParams& p = L->hp;
  if (g.H % 16 == 0) {
    p.th = 16;
    p.tn = 1;
  } else if (g.H == 8) {
    p.th = 8;
    p.tn = 2;
  } else {
    return false;
  }
  p.tiles_w = g.W / 8;
  p.tiles_h = g.H / p.th;
  p.tiles_b = (g.N + p.tn - 1) / p.tn;
  p.kchunks = g.Cin / 64;
  p.n_tiles = g.Cout / 64;
  p.cin = g.Cin;
  p.cout = g.Cout;
  p.dw = dw;
  p.dbg = env_int("VPD_WGRAD_DBG", 0);
  p.num_taps = 9;
  p.num_pairs = 5;
  for (int t = 0; t < 9; ++t) p.tap_row[t] = (t / 3) * 10 + t % 3;
  p.patch_dx = -1;
  p.patch_dy = -1;
  p.dy_c0 = 0;
  p.stem_ntx = 0;
  p.stem_cls = 0;
  const int img_rows = (p.th + 2) * 10;                 // patch rows per image
  p.patch_bytes = img_rows * p.tn * 128;
  if (p.patch_bytes > WgradHaloCfg::kPatchSlot) return false;
  for (int k = 0; k < 8; ++k) {
    // pixels 16k .. 16k+15 of the tile = image k / (th/2), rows 2 (k % (th/2)) and the next
    const int per_img = p.th / 2;
    const int row = (k / per_img) * img_rows + 2 * (k % per_img) * 10;
    p.kstep16[k] = row * 128 / 16;
  }
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int combos = p.kchunks * p.n_tiles;
  int splits = device_sm_count() / combos;
  const int cap = env_int("VPD_WGRAD_SPLITS", 0);      // experiments: fewer, longer work items
  if (cap > 0 && splits > cap) splits = cap;
  if (splits < 1) splits = 1;
  if (splits > pix_tiles) splits = pix_tiles;
  while (splits > 1 && (splits - 1) * ((pix_tiles + splits - 1) / splits) >= pix_tiles) --splits;
  p.splits = splits;
  const int items = combos * splits;
  L->grid = items < device_sm_count() ? items : device_sm_count();
  uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.W, 1, (uint64_t)g.H, (uint64_t)g.N};
  uint64_t str[5] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.W * g.Cin * 2,
                     (uint64_t)g.H * g.W * g.Cin * 2};
  uint32_t box[5] = {64, 10, 1, (uint32_t)(p.th + 2), (uint32_t)p.tn};
  if (encode_tmap_bf16(&L->x, x, 5, dims, str, box, true)) return false;
  uint64_t ddims[5] = {(uint64_t)g.Cout, (uint64_t)g.W, 1, (uint64_t)g.H, (uint64_t)g.N};
  uint64_t dstr[5] = {2, (uint64_t)g.Cout * 2, (uint64_t)g.W * g.Cout * 2,
                      (uint64_t)g.W * g.Cout * 2, (uint64_t)g.H * g.W * g.Cout * 2};
  uint32_t dbox[5] = {64, 8, 1, (uint32_t)p.th, (uint32_t)p.tn};
  if (encode_tmap_bf16(&L->dy, dy, 5, ddims, dstr, dbox, true)) return false;
  L->halo = 1;
  L->block_n = 64;
  return true;
}

int plan_conv_wgrad(WgradLaunch* L, const ConvGeom& g, const __nv_bfloat16* x,
                    const __nv_bfloat16* dy, float* dw) {
  memset(L, 0, sizeof(*L));
  if (try_wgrad_halo(L, g, x, dy, dw)) return 0;
  memset(L, 0, sizeof(*L));
  // reuse the forward plan for tile geometry, taps and the X tensor-map view
  ConvLaunch f;
  ConvEpilogue e;
  g_plan_no_halo = true;
  const int rc = plan_conv_fwd(&f, g, x, reinterpret_cast<const __nv_bfloat16*>(dw),
                               const_cast<__nv_bfloat16*>(dy), e);
  g_plan_no_halo = false;
  if (rc) return -1;
  // 256-wide output blocks (conv_wgrad_kernel<256>) work on 64-pixel tiles: halve the tile
  // along the image (or row) axis and rebuild the X view with the smaller box
  static const int wide_on = env_int("VPD_WGRAD256", 1);
  if (wide_on && g.Cout % 256 == 0 && (f.p.tn % 2 == 0 || f.p.th % 2 == 0)) {
    if (f.p.tn % 2 == 0) {
      f.p.tn /= 2;
      f.p.tiles_b = (g.N + f.p.tn - 1) / f.p.tn;
    } else {
      f.p.th /= 2;
      f.p.tiles_h = (g.Ho() + f.p.th - 1) / f.p.th;
    }
    if (act_map(&f.a0, x, g.N, g.H, g.W, g.Cin, g.stride, f.p)) return -1;
    L->block_n = 256;
  }
  copy_geometry(&L->p, f.p);
  L->x = f.a0;
  if (act_map(&L->dy, dy, g.N, g.Ho(), g.Wo(), g.Cout, 1, f.p)) return -1;
  finish_wgrad(L, g.Cin, g.Cout, g.k * g.k, g.Cin / 64, 64, dw);
  return 0;
}

// Stem weight gradient on the space-to-depth input: per column class the halo-reuse wgrad
// kernel (one aligned cell patch per 16 x 8-pixel tile feeds all 8 / 12 taps; accumulator rows
// are cell elements, mapped back to (kh, kw, c) when they leave for the arena).
int plan_stem_wgrad(WgradLaunch* L2, int N, int H, int W, const __nv_bfloat16* x_s2d,
                    const __nv_bfloat16* dy, float* dw) {
  VPD_REQUIRE(H % 32 == 0 && W % 32 == 0, "stem: H and W must be multiples of 32 (%d,%d)", H, W);
  const int Ho = H / 2, Wo = W / 2, Hs = stem_cells_h(H), Ws = stem_cells_w(W);
  for (int cls = 0; cls < 2; ++cls) {
    WgradLaunch* L = L2 + cls;
    memset(L, 0, sizeof(*L));
    WgradHaloParams& p = L->hp;
    p.th = 16;
    p.tn = 1;
    p.tiles_w = (Wo / 2) / 8;
    p.tiles_h = Ho / 16;
    p.tiles_b = N;
    p.kchunks = 1;
    p.n_tiles = 1;
    p.cin = 64;
    p.cout = 64;
    p.dw = dw;
    p.dbg = 0;
    p.patch_bytes = 19 * 10 * 128;
    for (int k = 0; k < 8; ++k) p.kstep16[k] = 2 * k * 10 * 128 / 16;
    const int ntx = 2 + cls;
    p.num_taps = stem_taps(cls);
    p.num_pairs = (p.num_taps + 1) / 2;
    for (int t = 0; t < p.num_taps; ++t) p.tap_row[t] = (t / ntx) * 10 + t % ntx;
    p.patch_dx = 0;
    p.patch_dy = 0;
    p.dy_c0 = cls * 64;
    p.stem_ntx = ntx;
    p.stem_cls = cls;
    const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    int splits = device_sm_count();
    if (splits > pix_tiles) splits = pix_tiles;
    while (splits > 1 && (splits - 1) * ((pix_tiles + splits - 1) / splits) >= pix_tiles) --splits;
    p.splits = splits;
    L->grid = splits;
    uint64_t dims[5] = {64, (uint64_t)Ws, 1, (uint64_t)Hs, (uint64_t)N};
    uint64_t str[5] = {2, 128, (uint64_t)Ws * 128, (uint64_t)Ws * 128, (uint64_t)Hs * Ws * 128};
    uint32_t box[5] = {64, 10, 1, 19, 1};
    if (encode_tmap_bf16(&L->x, x_s2d, 5, dims, str, box, true)) return -1;
    uint64_t ddims[5] = {128, (uint64_t)Wo / 2, 1, (uint64_t)Ho, (uint64_t)N};
    uint64_t dstr[5] = {2, 256, (uint64_t)Wo * 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128};
    uint32_t dbox[5] = {64, 8, 1, 16, 1};
    if (encode_tmap_bf16(&L->dy, dy, 5, ddims, dstr, dbox, true)) return -1;
    L->halo = 1;
    L->block_n = 64;
  }
  return 0;
}

template <int BN>
static int launch_wgrad_bn(const WgradLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VPD_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        WgradCfg<BN>::kSmemBytes));
    attr_set = true;
  }
  VPD_CHECK_CUDA(launch_kernel(conv_wgrad_kernel<BN>, dim3(L.grid), dim3(kWgradThreads), WgradCfg<BN>::kSmemBytes, stream, L.x, L.dy, L.p));
  VPD_LAUNCHED(1);
  return 0;
}

static int launch_wgrad_halo(const WgradLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VPD_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<64>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        WgradHaloCfg::kSmemBytes));
    attr_set = true;
  }
  VPD_CHECK_CUDA(launch_kernel(conv_wgrad_halo_kernel<64>, dim3(L.grid), dim3(kWgradThreads),
                               WgradHaloCfg::kSmemBytes, stream, L.x, L.dy, L.hp));
  VPD_LAUNCHED(1);
  return 0;
}

int launch_wgrad(const WgradLaunch& L, cudaStream_t stream) {
  if (L.grid <= 0) return 0;
  static const int skip = env_int("VPD_DBG_SKIP_WGRAD", 0);  // timing experiments only
  if (skip) return 0;
  if (L.halo) return launch_wgrad_halo(L, stream);
  if (L.block_n == 64) return launch_wgrad_bn<64>(L, stream);
  if (L.block_n == 256) return launch_wgrad_bn<256>(L, stream);
  return launch_wgrad_bn<128>(L, stream);
}

// ---------------------------------------------------------------- weight packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_tap,
                                        __nv_bfloat16* __restrict__ wT_tap, int Cout, int Cin,
                                        int kk) {
  pdl_trigger();
  pdl_wait();
  // one thread per (co, ci): reads kk contiguous floats, scatters to both layouts
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin) return;
  const int ci = (int)(i % Cin), co = (int)(i / Cin);
  const float* src = w + i * kk;
  for (int t = 0; t < kk; ++t) {
    const __nv_bfloat16 v = __float2bfloat16_rn(src[t]);
    if (w_tap) w_tap[(long long)t * Cout * Cin + wtile_offset(co, ci, Cout)] = v;
    if (wT_tap) wT_tap[(long long)t * Cout * Cin + wtile_offset(ci, co, Cin)] = v;
  }
}

int pack_conv_weight(const float* w_oihw, __nv_bfloat16* w_tap, __nv_bfloat16* wT_tap, int Cout,
                     int Cin, int k, cudaStream_t stream) {
  const long long n = (long long)Cout * Cin;
  VPD_CHECK_CUDA(launch_kernel(pack_conv_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, w_oihw, w_tap, wT_tap,
                                                                          Cout, Cin, k * k));
  VPD_LAUNCHED(1);
  return 0;
}

// Stem operand mirrors for the space-to-depth kernel (see plan_stem_fwd): per column-parity
// class and tap (dy, dxb) a 64 (cout) x 64 (cell element e = (a*4 + q)*8 + c) tile in the
// pre-tiled operand layout; the entry is weight (kh = 2 dy + a, kw = 4 dxb + q - 2 cls, c) or
// zero where that falls outside the 7 x 7 x Cimg kernel. kArena: the source is the parameter
// arena's packed stem block [kh][co][kw*8 + c] (fp32), else the reference's OIHW tensor.
template <bool kArena>
__global__ void pack_stem_s2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws,
                                     int Cimg) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 20 * 4096) return;
  const int e = i % 64, co = (i / 64) % 64, tg = i / 4096;       // tg: 0..7 class 0, 8..19 class 1
  const int cls = tg >= 8 ? 1 : 0, t = tg - 8 * cls, ntx = 2 + cls;
  const int dy = t / ntx, dxb = t % ntx;
  const int a = e >> 5, q = (e >> 3) & 3, c = e & 7;
  const int kh = 2 * dy + a, kw = 4 * dxb + q - 2 * cls;
  float v = 0.f;
  if (kh < 7 && kw >= 0 && kw < 7 && c < Cimg)
    v = kArena ? w[kh * 4096 + co * 64 + kw * 8 + c] : w[((co * Cimg + c) * 7 + kh) * 7 + kw];
  ws[(long long)tg * 4096 + wtile_offset(co, e, 64)] = __float2bfloat16_rn(v);
}

int pack_stem_weight(const float* w_oihw, __nv_bfloat16* w_stem, int Cimg, cudaStream_t stream) {
  VPD_REQUIRE(Cimg >= 1 && Cimg <= 8, "stem: %d input channels unsupported", Cimg);
  VPD_CHECK_CUDA(launch_kernel(pack_stem_s2d_kernel<false>, dim3((20 * 4096 + 255) / 256), dim3(256),
                               0, stream, w_oihw, w_stem, Cimg));
  VPD_LAUNCHED(1);
  return 0;
}

int pack_stem_weight_arena(const float* w_arena, __nv_bfloat16* w_stem, cudaStream_t stream) {
  VPD_CHECK_CUDA(launch_kernel(pack_stem_s2d_kernel<true>, dim3((20 * 4096 + 255) / 256), dim3(256),
                               0, stream, w_arena, w_stem, 8));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
