// Host-side planning for the implicit-GEMM convolution kernels (K2/K2b/K2c).
#pragma once
#include "conv_igemm.cuh"
#include "tma_host.h"

namespace vpd {

struct ConvLaunch {
  CUtensorMap a0, a1;  // activation views read by the taps (tap.src 0 / 1)
  CUtensorMap o;       // output view written by the TMA tile stores
  const void* w_ptr;   // this launch's (main) weight block, for the previous launch's L2 prefetch
  long long w_bytes;
  ConvParams p;
  int block_n;
  int cluster;  // CTAs per cluster along M (weight multicast)
  int halo;     // 0 = generic kernel, else number of 64-channel chunks (halo-reuse kernel)
  int grid;
};

// debug: generic-kernel launches record 8 int64 per CTA into dev_buf (null = off)
void set_conv_trace(long long* dev_buf);

struct ConvGeom {
  int N, H, W;     // input batch / height / width (of x; for dgrad: of dx)
  int Cin, Cout;
  int k, stride, pad;
  int Ho() const { return (H + 2 * pad - k) / stride + 1; }
  int Wo() const { return (W + 2 * pad - k) / stride + 1; }
};

struct ConvEpilogue {
  const float* scale = nullptr;
  const float* shift = nullptr;
  const __nv_bfloat16* residual = nullptr;
  int relu = 0;
  StatAcc* stats = nullptr;
};

// Fused BN-backward reduction for a dgrad launch (see ConvParams::bnb)
struct ConvBwdFuse {
  int nb = 0;
  const uint8_t* mask = nullptr;   // 1[z > 0], one bit per element (BnApplyParams::mask)
  const __nv_bfloat16* y[2] = {nullptr, nullptr};
  const float* mean[2] = {nullptr, nullptr};
  const float* rstd[2] = {nullptr, nullptr};
  StatAcc* sums[2] = {nullptr, nullptr};
};

int device_sm_count();

// y[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin], w) ; w_tap = bf16 [k*k][Cout][Cin]
int plan_conv_fwd(ConvLaunch* L, const ConvGeom& g, const __nv_bfloat16* x,
                  const __nv_bfloat16* w_tap, __nv_bfloat16* y, const ConvEpilogue& e);

// 7x7/2 pad 3 stem on the space-to-depth input (common.cuh::stem_pixel_offset): TWO launches
// (even / odd output columns) into L[0], L[1]; w_s2d = kStemMirrorElems bf16 (pack_stem_weight)
constexpr int kStemMirrorElems = 20 * 4096;
int plan_stem_fwd(ConvLaunch* L, int N, int H, int W, const __nv_bfloat16* x_s2d,
                  const __nv_bfloat16* w_s2d, __nv_bfloat16* y, const ConvEpilogue& e);

// dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], w) (+ residual)
// wT_tap = bf16 [k*k][Cin][Cout]. For stride 2 this produces 4 launches (one
// per output-pixel parity class); the 1x1/2 downsample branch (dy_ds, wT_ds
// = bf16 [1][Cin][Cout_ds]) is fused into class (0,0).
int plan_conv_dgrad(ConvLaunch* L, int* count, const ConvGeom& g, const __nv_bfloat16* dy,
                    const __nv_bfloat16* wT_tap, __nv_bfloat16* dx,
                    const __nv_bfloat16* residual, const __nv_bfloat16* dy_ds,
                    const __nv_bfloat16* wT_ds, int cout_ds, const ConvBwdFuse* fuse = nullptr);

int launch_conv(const ConvLaunch& L, cudaStream_t stream);
// `prev` pulls the weights of `next` into L2 while it runs
inline void chain_weight_prefetch(ConvLaunch* prev, const ConvLaunch& next) {
  prev->p.pf_ptr = static_cast<const __nv_bfloat16*>(next.w_ptr);
  prev->p.pf_bytes = next.w_bytes;
}

struct WgradLaunch {
  CUtensorMap x, dy;
  WgradParams p;
  int block_n;
  int grid;
  int halo;              // 1: conv_wgrad_halo_kernel (stride-1 3x3 on 8-pixel-wide tiles), hp valid
  WgradHaloParams hp;
};
// dw[k*k][Cout][Cin] (fp32, tap-major) += sum over pixels dy * x  (accumulates!)
int plan_conv_wgrad(WgradLaunch* L, const ConvGeom& g, const __nv_bfloat16* x,
                    const __nv_bfloat16* dy, float* dw);
// stem: dw[7][64][64] (kh, cout, kw*8+c; entries with kw == 7 are left untouched); TWO launches
// (even / odd output columns) into L[0], L[1]
int plan_stem_wgrad(WgradLaunch* L, int N, int H, int W, const __nv_bfloat16* x_s2d,
                    const __nv_bfloat16* dy, float* dw);
int launch_wgrad(const WgradLaunch& L, cudaStream_t stream);

// fp32 OIHW master weights -> bf16 tap-major copies used by the kernels
int pack_conv_weight(const float* w_oihw, __nv_bfloat16* w_tap, __nv_bfloat16* wT_tap, int Cout,
                     int Cin, int k, cudaStream_t stream);
int pack_stem_weight(const float* w_oihw, __nv_bfloat16* w_stem, int Cimg, cudaStream_t stream);
// same mirrors from the parameter arena's packed stem block [kh][co][kw*8+c] (fp32)
int pack_stem_weight_arena(const float* w_arena, __nv_bfloat16* w_stem, cudaStream_t stream);

}  // namespace vpd
