#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vpd {

struct HeadParams {
  const __nv_bfloat16* z;  // [B][HW][F] final feature map (post-ReLU), NHWC
  int B, HW, F, D, T, Hd, motion;
  const float* fc_w;       // [D][F]
  const float* fc_b;       // [D]
  const float* w0;         // decoder layers.0 [Hd][D]   (motion only)
  const float* b0;
  const float* w2;         // decoder layers.2 [Hd][Hd]
  const float* b2;
  const float* w5;         // decoder layers.5 [T][Hd]
  const float* b5;
  const float* target;     // [B][T] or null (no loss)
  float* emb_out;          // [B][D] encoder output, or null
  float* out;              // [B][T] decoder output, or null
  double* loss;            // += sum of squared errors
  __nv_bfloat16* dz;       // [B][HW][F] gradient of the feature map, or null (no backward)
  float* ws;               // [B][head_ws_stride] saved vectors (backward only)
};

struct HeadGrads {
  float *fc_w, *fc_b, *w0, *b0, *w2, *b2, *w5, *b5;
};

__host__ __device__ inline int head_ws_stride(int F, int D, int Hd, int T) {
  return F + 2 * D + 4 * Hd + T;
}

int launch_head(const HeadParams& p, const HeadGrads* grads, cudaStream_t stream);

}  // namespace vpd
