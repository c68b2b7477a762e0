// K1a: the augmented training batch (the stochastic half).
//
// Reference per frame (vpd_dataset/single_frame.py:168-206 with augment=True):
//   x = u8 / 255 -> ColorJitter (common.py:88-92; torchvision _functional_tensor.py:
//   _blend, rgb_to_grayscale, adjust_contrast, _rgb2hsv / _hsv2rgb) -> Normalize ->
//   masked Gaussian noise (:179-191) -> cat flow -> horizontal flip (+ flow-x sign) ->
//   RandomResizedCrop (common.py:49-50: crop (i, j, h, w), antialiased bilinear resize back
//   to H x W; ATen UpSampleKernel.cpp `_compute_indices_min_size_weights_aa`, horizontal then
//   vertical pass, fused multiply-add chains).
// The random DRAWS (jitter order and factors, crop box, flip, noise coin) come from the host
// (vpd_b200/augment.py draws them in the reference's order from the same generators); the
// pixel arithmetic runs here with the reference's rounding sequence: every torchvision tensor
// op rounds to fp32 on its own, so each step is an explicit _rn intrinsic (nvcc would contract
// a*b+c otherwise), and the resize uses fmaf exactly where ATen's vectorised loop does. The
// one value that cannot match bit for bit is adjust_contrast's grayscale mean (torch.mean's
// summation order depends on the CPU's vector width): it is the fp64 sum over the frame.
//
// One CTA per frame. The three jittered / normalised RGB planes of the frame live in shared
// memory as fp32 (3 * H * W * 4 = 192 KB at 128 x 128) because the resize gathers from them
// at up to 3 x 3 taps per output; the two flow planes reuse the same storage afterwards.
// HBM traffic is the algorithmic minimum (uint8 in, fp32 NCHW out, coalesced).
#include "common.cuh"
#include "ops.h"
#include "rng.cuh"
#include "tma_host.h"

namespace vpd {

constexpr int kAugThreads = 1024;   // 32 warps: the jitter chains (IEEE divisions) need the latency hiding

struct AugParams {
  const uint8_t* rgb;     // [pool][H][W][3]
  const uint8_t* flow;    // [pool][H][W][fc] or null
  const int* index;       // [B] or null
  const uint8_t* flip;    // [B] or null
  const float* teacher;   // [pool][rows][tdim] or null
  float mean[3], stdv[3];
  int B, H, W, fc, teacher_rows, tdim;
  const uint8_t* jorder;  // [B][4] op order (0 brightness 1 contrast 2 saturation 3 hue, >3 skip)
  const float* jfactor;   // [B][8] {b, c, 1-c, s, 1-s, hue, -, -} rounded from the host doubles
  const int* crop;        // [B][4] (i, j, h, w) or null
  const uint8_t* mask;
  const uint8_t* noise_on;
  const float* noise;
  float noise_sd;
  unsigned long long seed;
  float* out_img;         // [B][C][H][W]
  float* out_tgt;         // [B][tdim]
};

VPD_DEVINL float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
// _blend(img1, img2, ratio) with r = fp32(ratio), q = fp32(1 - ratio)
VPD_DEVINL float blend(float a, float b, float r, float q) {
  return clamp01(__fadd_rn(__fmul_rn(r, a), __fmul_rn(q, b)));
}
VPD_DEVINL float gray_of(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}

// adjust_hue: _rgb2hsv, (h + f) % 1, _hsv2rgb
VPD_DEVINL void hue_shift(float& r, float& g, float& b, float hf) {
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
  const float crd = eqc ? 1.f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), crd);
  const float gc = __fdiv_rn(__fsub_rn(maxc, g), crd);
  const float bc = __fdiv_rn(__fsub_rn(maxc, b), crd);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
  const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
  const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
  float m = fmodf(__fadd_rn(h, hf), 1.f);          // torch remainder: sign of the divisor
  if (m != 0.f && m < 0.f) m = __fadd_rn(m, 1.f);
  const float h6 = __fmul_rn(m, 6.f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  const int i = static_cast<int>(fi) % 6;
  const float v = maxc;
  const float p = clamp01(__fmul_rn(v, __fsub_rn(1.f, s)));
  const float q = clamp01(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f))));
  const float t = clamp01(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f)))));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// one pointwise ColorJitter op (everything but contrast, which needs the frame mean)
VPD_DEVINL void jitter_op(int op, const float* jf, float& r, float& g, float& b) {
  if (op == 0) {          // brightness: blend with zeros
    r = clamp01(__fmul_rn(jf[0], r));
    g = clamp01(__fmul_rn(jf[0], g));
    b = clamp01(__fmul_rn(jf[0], b));
  } else if (op == 2) {   // saturation: blend with the pixel's gray value
    const float l = gray_of(r, g, b);
    r = blend(r, l, jf[3], jf[4]);
    g = blend(g, l, jf[3], jf[4]);
    b = blend(b, l, jf[3], jf[4]);
  } else if (op == 3) {
    hue_shift(r, g, b, jf[5]);
  }
}

// ATen _compute_indices_min_size_weights_aa for output index i, support 1 (in <= out)
VPD_DEVINL void aa_weights(int in_size, int out_size, int i, int* xmin, int* xsize, float* w) {
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const float center = static_cast<float>(__dmul_rn(static_cast<double>(scale), static_cast<double>(i) + 0.5));
  long long lo = static_cast<long long>(__dadd_rn(static_cast<double>(__fsub_rn(center, 1.f)), 0.5));
  if (lo < 0) lo = 0;
  long long hi = static_cast<long long>(__dadd_rn(static_cast<double>(__fadd_rn(center, 1.f)), 0.5));
  if (hi > in_size) hi = in_size;
  int n = static_cast<int>(hi - lo);
  n = n < 0 ? 0 : (n > 3 ? 3 : n);
  float total = 0.f;
  float ww[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < n; ++j) {
    const float d = __fsub_rn(static_cast<float>(j + lo), center);
    float x = static_cast<float>(__dadd_rn(static_cast<double>(d), 0.5));
    x = fabsf(x);
    ww[j] = x < 1.f ? __fsub_rn(1.f, x) : 0.f;
    total = __fadd_rn(total, ww[j]);
  }
  for (int j = 0; j < 3; ++j) w[j] = (j < n && total != 0.f) ? __fdiv_rn(ww[j], total) : ww[j];
  *xmin = static_cast<int>(lo);
  *xsize = n;
}

__global__ void __launch_bounds__(kAugThreads)
assemble_aug_kernel(const AugParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t sm[];
  const int H = p.H, W = p.W, HW = H * W;
  float* plane = reinterpret_cast<float*>(sm);                 // [3][HW]
  double* red = reinterpret_cast<double*>(plane + 3 * HW);     // [40]: 32 warp sums, [32] total
  int* xmin = reinterpret_cast<int*>(red + 40);                // [W]
  int* xsize = xmin + W;                                       // [W]
  int* ymin = xsize + W;                                       // [H]
  int* ysize = ymin + H;                                       // [H]
  float* wx = reinterpret_cast<float*>(ysize + H);             // [W][3]
  float* wy = wx + 3 * W;                                      // [H][3]
  float* flut = wy + 3 * H;                                    // [256] flow byte -> value
  float* ulut = flut + 256;                                    // [256] byte / 255
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int src = p.index ? p.index[b] : b;
  const bool fl = p.flip && p.flip[b];
  const int C = p.flow ? 5 : 3;
  int ci = 0, cj = 0, ch = H, cw = W;
  if (p.crop) {
    ci = p.crop[b * 4 + 0];
    cj = p.crop[b * 4 + 1];
    ch = p.crop[b * 4 + 2];
    cw = p.crop[b * 4 + 3];
    // a box outside the frame would read outside the planes: clamp (the host validates too)
    ch = min(max(ch, 1), H);
    cw = min(max(cw, 1), W);
    ci = min(max(ci, 0), H - ch);
    cj = min(max(cj, 0), W - cw);
  }

  // ---- tables: resize weights of this frame's crop, flow byte -> value ----
  for (int i = tid; i < W + H; i += kAugThreads) {
    if (i < W) aa_weights(cw, W, i, &xmin[i], &xsize[i], &wx[3 * i]);
    else aa_weights(ch, H, i - W, &ymin[i - W], &ysize[i - W], &wy[3 * (i - W)]);
  }
  for (int u = tid; u < 256; u += kAugThreads)
  {
    flut[u] = static_cast<float>(__dsub_rn(__ddiv_rn(static_cast<double>(u), 255.0), 0.5));
    ulut[u] = __fdiv_rn(static_cast<float>(u), 255.f);
  }
  if (p.teacher && p.out_tgt) {
    const int row = (p.teacher_rows > 1 && fl) ? 1 : 0;
    const float* t = p.teacher + ((size_t)src * p.teacher_rows + row) * p.tdim;
    for (int i = tid; i < p.tdim; i += kAugThreads) p.out_tgt[(size_t)b * p.tdim + i] = t[i];
  }

  // ---- RGB: u8/255 -> ColorJitter -> Normalize -> masked noise, into the smem planes ----
  int ops[4] = {9, 9, 9, 9};
  float jf[6] = {1.f, 1.f, 0.f, 1.f, 0.f, 0.f};
  int cpos = 4;  // position of the contrast op (needs the frame's gray mean)
  if (p.jorder) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ops[k] = p.jorder[b * 4 + k];
      if (ops[k] == 1) cpos = k;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) jf[k] = p.jfactor[b * 8 + k];
  }
  const bool noisy = p.mask != nullptr && (p.noise_on == nullptr || p.noise_on[b] != 0);
  auto finish = [&](int px, float r, float g, float bl) {
    float v[3] = {r, g, bl};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float o = __fdiv_rn(__fsub_rn(v[c], p.mean[c]), p.stdv[c]);
      if (noisy && p.mask[(size_t)src * HW + px] != 0) {
        const size_t e = (size_t)c * HW + px;
        const float nz = p.noise ? p.noise[(size_t)b * 3 * HW + e]
                                 : p.noise_sd * philox_normal(p.seed, static_cast<unsigned int>(e),
                                                              static_cast<unsigned int>(b));
        o = __fadd_rn(o, nz);
      }
      plane[c * HW + px] = o;
    }
  };

  __syncthreads();   // tables
  const uint32_t* rgb4 = reinterpret_cast<const uint32_t*>(p.rgb + (size_t)src * HW * 3);
  double gsum = 0.0;
  for (int q4 = tid; q4 < HW / 4; q4 += kAugThreads) {   // 4 pixels = 12 bytes = 3 words
    const uint32_t w0 = __ldg(rgb4 + 3 * q4), w1 = __ldg(rgb4 + 3 * q4 + 1),
                   w2 = __ldg(rgb4 + 3 * q4 + 2);
    const uint32_t by[12] = {w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, w0 >> 24,
                             w1 & 255, (w1 >> 8) & 255, (w1 >> 16) & 255, w1 >> 24,
                             w2 & 255, (w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int px = q4 * 4 + j;
      float r = ulut[by[3 * j]], g = ulut[by[3 * j + 1]], bl = ulut[by[3 * j + 2]];
      for (int k = 0; k < cpos; ++k) jitter_op(ops[k], jf, r, g, bl);
      if (cpos < 4) {
        gsum += static_cast<double>(gray_of(r, g, bl));
        plane[px] = r;
        plane[HW + px] = g;
        plane[2 * HW + px] = bl;
      } else {
        finish(px, r, g, bl);
      }
    }
  }
  if (cpos < 4) {
    // frame mean of the gray image: fp64 sum / count, rounded to fp32
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
    if ((tid & 31) == 0) red[tid >> 5] = gsum;
    __syncthreads();
    if (tid < 32) {
      double v = tid < kAugThreads / 32 ? red[tid] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (tid == 0) red[32] = v;
    }
    __syncthreads();
    const float gmean = static_cast<float>(__ddiv_rn(red[32], static_cast<double>(HW)));
    for (int px = tid; px < HW; px += kAugThreads) {
      float r = plane[px], g = plane[HW + px], bl = plane[2 * HW + px];
      r = blend(r, gmean, jf[1], jf[2]);
      g = blend(g, gmean, jf[1], jf[2]);
      bl = blend(bl, gmean, jf[1], jf[2]);
      for (int k = cpos + 1; k < 4; ++k) jitter_op(ops[k], jf, r, g, bl);
      finish(px, r, g, bl);
    }
  }
  __syncthreads();

  // ---- crop + antialiased bilinear resize out of the smem planes ----
  auto resize_planes = [&](int nplanes, int c0, bool negate_first) {
    for (int idx = tid; idx < nplanes * HW; idx += kAugThreads) {
      const int c = idx / HW, rem = idx - c * HW;
      const int oy = rem / W, ox = rem - oy * W;
      const float* pl = plane + c * HW;
      const int x0 = cj + xmin[ox], nx = xsize[ox];
      const int y0 = ci + ymin[oy], ny = ysize[oy];
      const float* wxp = wx + 3 * ox;
      const float* wyp = wy + 3 * oy;
      float acc = 0.f;
      for (int jy = 0; jy < ny; ++jy) {
        const float* rowp = pl + (y0 + jy) * W;
        float t = 0.f;
        for (int jx = 0; jx < nx; ++jx) {
          const int colf = x0 + jx;
          const float v = rowp[fl ? (W - 1 - colf) : colf];
          t = jx == 0 ? __fmul_rn(v, wxp[0]) : __fmaf_rn(v, wxp[jx], t);
        }
        acc = jy == 0 ? __fmul_rn(t, wyp[0]) : __fmaf_rn(t, wyp[jy], acc);
      }
      if (negate_first && c == 0) acc = -acc;
      p.out_img[((size_t)b * C + c0 + c) * HW + rem] = acc;
    }
  };
  resize_planes(3, 0, false);

  if (p.flow) {
    __syncthreads();
    const uint8_t* fsrc = p.flow + (size_t)src * HW * p.fc;
    for (int px = tid; px < HW; px += kAugThreads) {
      plane[px] = flut[fsrc[(size_t)px * p.fc]];
      plane[HW + px] = flut[fsrc[(size_t)px * p.fc + 1]];
    }
    __syncthreads();
    resize_planes(2, 3, fl);   // flipped frames: flow-x changes sign (exact, commutes with fma)
  }
}

int assemble_aug(const uint8_t* rgb, const uint8_t* flow, int flow_channels, const int* index,
                 const uint8_t* flip, const float* teacher, int teacher_rows, int tdim,
                 const float* mean, const float* stdv, float* out_img, float* out_tgt, int B,
                 int H, int W, const uint8_t* jitter_order, const float* jitter_factor,
                 const int* crop, cudaStream_t stream, const AsmNoise* nz) {
  VPD_REQUIRE(B >= 0, "assemble_aug: negative batch");
  VPD_REQUIRE(rgb != nullptr && out_img != nullptr, "assemble_aug: null rgb / output");
  VPD_REQUIRE((H * W) % 4 == 0, "assemble_aug: H*W must be a multiple of 4 (%d x %d)", H, W);
  VPD_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 3) == 0, "assemble_aug: rgb must be 4-byte aligned");
  VPD_REQUIRE(flow == nullptr || flow_channels >= 2, "assemble_aug: flow needs >= 2 channels");
  VPD_REQUIRE((jitter_order == nullptr) == (jitter_factor == nullptr),
              "assemble_aug: jitter order and factors come together");
  if (B == 0) return 0;
  AugParams p;
  p.rgb = rgb;
  p.flow = flow;
  p.index = index;
  p.flip = flip;
  p.teacher = teacher;
  for (int i = 0; i < 3; ++i) {
    p.mean[i] = mean[i];
    p.stdv[i] = stdv[i];
  }
  p.B = B;
  p.H = H;
  p.W = W;
  p.fc = flow_channels;
  p.teacher_rows = teacher_rows;
  p.tdim = tdim;
  p.jorder = jitter_order;
  p.jfactor = jitter_factor;
  p.crop = crop;
  p.mask = nullptr;
  p.noise_on = nullptr;
  p.noise = nullptr;
  p.noise_sd = 0.f;
  p.seed = 0;
  if (nz != nullptr && nz->mask != nullptr) {
    p.mask = nz->mask;
    p.noise_on = nz->noise_on;
    p.noise = nz->noise;
    p.noise_sd = nz->noise_sd;
    p.seed = nz->seed;
  }
  p.out_img = out_img;
  p.out_tgt = out_tgt;
  const size_t smem = (size_t)3 * H * W * 4 + 40 * 8 + (size_t)(2 * W + 2 * H) * 4 +
                      (size_t)(3 * W + 3 * H) * 4 + 2 * 256 * 4;
  VPD_REQUIRE(smem <= 227 * 1024, "assemble_aug: frame too large for shared memory (%d x %d)", H, W);
  if (smem > 48 * 1024)
    VPD_CHECK_CUDA(cudaFuncSetAttribute(assemble_aug_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VPD_CHECK_CUDA(launch_kernel(assemble_aug_kernel, dim3(B), dim3(kAugThreads), smem, stream, p));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
