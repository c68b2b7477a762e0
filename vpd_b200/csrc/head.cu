// K4: embedding head + distillation loss + its gradient (SURVEY §8 rows A5 tail,
// A7, A8): global average pool -> fc (F -> D) -> [FCNet decoder D -> 128 -> 128 ->
// T, models/module.py:133-156] -> sum-of-squares loss against the teacher row
// (F.mse_loss(reduction='sum'), train_vpd_model.py:87) and, in the same kernel,
// the backward pass down to the gradient of the final feature map. One CTA per
// frame; dot products are warp-shuffle reductions / thread-per-output loops over
// shared memory; everything is fp32. A second small kernel turns the saved
// per-frame vectors into the weight gradients (outer-product sums over the batch).
#include "common.cuh"
#include "head.h"
#include "tma_host.h"

namespace vpd {

constexpr int kHeadThreads = 512;

__global__ void __launch_bounds__(kHeadThreads) head_kernel(const HeadParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sh[];
  float* pooled = sh;                 // F
  float* e = pooled + p.F;            // D
  float* h1 = e + p.D;                // Hd
  float* h2 = h1 + p.Hd;              // Hd
  float* o = h2 + p.Hd;               // T   (later: dO)
  float* dh2 = o + p.T;               // Hd
  float* dh1 = dh2 + p.Hd;            // Hd
  float* de = dh1 + p.Hd;             // D
  __shared__ float red[kHeadThreads / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inv_hw = 1.0f / static_cast<float>(p.HW);

  // 1. global average pool (torch: sum / HW in fp32)
  const __nv_bfloat16* zb = p.z + (size_t)b * p.HW * p.F;
  for (int c = tid; c < p.F; c += kHeadThreads) {
    float s = 0.f;
    for (int r = 0; r < p.HW; ++r) s += __bfloat162float(zb[(size_t)r * p.F + c]);
    pooled[c] = s * inv_hw;
  }
  __syncthreads();
  // 2. fc: one warp per output
  for (int d = warp; d < p.D; d += kHeadThreads / 32) {
    const float* w = p.fc_w + (size_t)d * p.F;
    float s = 0.f;
    for (int c = lane; c < p.F; c += 32) s = fmaf(w[c], pooled[c], s);
    s = warp_sum(s);
    if (lane == 0) e[d] = s + p.fc_b[d];
  }
  __syncthreads();
  if (p.emb_out != nullptr)
    for (int d = tid; d < p.D; d += kHeadThreads) p.emb_out[(size_t)b * p.D + d] = e[d];
  // 3. decoder: one warp per output row, lanes stride the input (coalesced weight reads)
  if (p.motion) {
    for (int j = warp; j < p.Hd; j += kHeadThreads / 32) {
      float s = 0.f;
      for (int i = lane; i < p.D; i += 32) s = fmaf(p.w0[j * p.D + i], e[i], s);
      s = warp_sum(s);
      if (lane == 0) h1[j] = fmaxf(s + p.b0[j], 0.f);
    }
    __syncthreads();
    for (int j = warp; j < p.Hd; j += kHeadThreads / 32) {
      float s = 0.f;
      for (int i = lane; i < p.Hd; i += 32) s = fmaf(p.w2[j * p.Hd + i], h1[i], s);
      s = warp_sum(s);
      if (lane == 0) h2[j] = fmaxf(s + p.b2[j], 0.f);
    }
    __syncthreads();
    for (int t = warp; t < p.T; t += kHeadThreads / 32) {
      float s = 0.f;
      for (int i = lane; i < p.Hd; i += 32) s = fmaf(p.w5[t * p.Hd + i], h2[i], s);
      s = warp_sum(s);
      if (lane == 0) o[t] = s + p.b5[t];
    }
  } else {
    for (int t = tid; t < p.T; t += kHeadThreads) o[t] = e[t];
  }
  __syncthreads();
  if (p.out != nullptr)
    for (int t = tid; t < p.T; t += kHeadThreads) p.out[(size_t)b * p.T + t] = o[t];
  if (p.target == nullptr) return;

  // 4. loss = sum (o - t)^2 ; dO = 2 (o - t)
  float part = 0.f;
  for (int t = tid; t < p.T; t += kHeadThreads) {
    const float diff = o[t] - p.target[(size_t)b * p.T + t];
    part = fmaf(diff, diff, part);
    o[t] = 2.f * diff;
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < kHeadThreads / 32; ++i) s += red[i];
    atomicAdd(p.loss, static_cast<double>(s));
  }
  if (p.dz == nullptr) return;

  // 5. backward through the decoder and fc
  if (p.motion) {
    for (int j = tid; j < p.Hd; j += kHeadThreads) {
      float s = 0.f;
      for (int t = 0; t < p.T; ++t) s = fmaf(p.w5[t * p.Hd + j], o[t], s);
      dh2[j] = h2[j] > 0.f ? s : 0.f;
    }
    __syncthreads();
    for (int j = tid; j < p.Hd; j += kHeadThreads) {
      float s = 0.f;
      for (int k = 0; k < p.Hd; ++k) s = fmaf(p.w2[k * p.Hd + j], dh2[k], s);
      dh1[j] = h1[j] > 0.f ? s : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < p.D; i += kHeadThreads) {
      float s = 0.f;
      for (int j = 0; j < p.Hd; ++j) s = fmaf(p.w0[j * p.D + i], dh1[j], s);
      de[i] = s;
    }
  } else {
    for (int i = tid; i < p.D; i += kHeadThreads) de[i] = o[i];
  }
  __syncthreads();
  // d pooled -> broadcast over the HW positions of the final feature map
  __nv_bfloat16* dzb = p.dz + (size_t)b * p.HW * p.F;
  for (int c = tid; c < p.F; c += kHeadThreads) {
    float s = 0.f;
    for (int d = 0; d < p.D; ++d) s = fmaf(p.fc_w[(size_t)d * p.F + c], de[d], s);
    const __nv_bfloat16 v = __float2bfloat16_rn(s * inv_hw);
    for (int r = 0; r < p.HW; ++r) dzb[(size_t)r * p.F + c] = v;
  }
  // 6. save the per-frame vectors the weight-gradient kernel needs
  float* w = p.ws + (size_t)b * head_ws_stride(p.F, p.D, p.Hd, p.T);
  for (int c = tid; c < p.F; c += kHeadThreads) w[c] = pooled[c];
  w += p.F;
  for (int i = tid; i < p.D; i += kHeadThreads) {
    w[i] = e[i];
    w[p.D + i] = de[i];
  }
  w += 2 * p.D;
  for (int j = tid; j < p.Hd; j += kHeadThreads) {
    w[j] = h1[j];
    w[p.Hd + j] = h2[j];
    w[2 * p.Hd + j] = dh1[j];
    w[3 * p.Hd + j] = dh2[j];
  }
  w += 4 * p.Hd;
  for (int t = tid; t < p.T; t += kHeadThreads) w[t] = o[t];
}

// dW[o][i] = sum_b dout[b][o] * in[b][i],  db[o] = sum_b dout[b][o]
struct OuterSeg {
  const float* dout;
  const float* in;
  float* dW;
  float* db;
  int O, I;
};
struct OuterParams {
  OuterSeg seg[4];
  int nseg;
  int B;
  int stride;  // floats between consecutive frames in the workspace
};

// One CTA per 8 x 32 tile of one weight matrix (176 CTAs for the default head): the two
// operand tiles of up to 256 frames (dout[b][8], in[b][32]) are staged in shared memory with
// coalesced loads, all in flight at once, and every thread accumulates ONE weight element
// over the frames in frame order (a fixed, run-to-run identical reduction). The first version
// gave 8 lanes to every weight element, each walking every 8th frame straight from global
// memory: a warp instruction touched eight 16-byte pieces of eight different rows, 57 us per
// step at batch 256 for 23 MFLOP.
constexpr int kOwTileO = 8, kOwTileI = 32, kOwChunk = 256;
__global__ void __launch_bounds__(256) head_wgrad_kernel(const OuterParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_do[kOwChunk][kOwTileO];
  __shared__ float s_in[kOwChunk][kOwTileI];
  int tile = blockIdx.x, seg = 0, tiles_i = 1;
  for (; seg < p.nseg; ++seg) {
    tiles_i = (p.seg[seg].I + kOwTileI - 1) / kOwTileI;
    const int n = ((p.seg[seg].O + kOwTileO - 1) / kOwTileO) * tiles_i;
    if (tile < n) break;
    tile -= n;
  }
  if (seg >= p.nseg) return;
  const OuterSeg g = p.seg[seg];
  const int o0 = (tile / tiles_i) * kOwTileO, i0 = (tile % tiles_i) * kOwTileI;
  const int to = threadIdx.x >> 5, ti = threadIdx.x & 31;
  float acc = 0.f, accb = 0.f;
  for (int b0 = 0; b0 < p.B; b0 += kOwChunk) {
    const int nb = min(kOwChunk, p.B - b0);
    // dout tile: 8 columns x nb rows, in tile: 32 columns x nb rows (zero outside the matrix)
#pragma unroll
    for (int e = threadIdx.x; e < kOwChunk * kOwTileO; e += 256) {
      const int r = e / kOwTileO, c = e % kOwTileO;
      s_do[r][c] = (r < nb && o0 + c < g.O) ? __ldg(g.dout + (size_t)(b0 + r) * p.stride + o0 + c) : 0.f;
    }
#pragma unroll 8
    for (int e = threadIdx.x; e < kOwChunk * kOwTileI; e += 256) {
      const int r = e / kOwTileI, c = e % kOwTileI;
      s_in[r][c] = (r < nb && i0 + c < g.I) ? __ldg(g.in + (size_t)(b0 + r) * p.stride + i0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kOwChunk; ++r) {
      const float d = s_do[r][to];
      acc = fmaf(d, s_in[r][ti], acc);
      accb += d;
    }
    __syncthreads();
  }
  const int o = o0 + to;
  if (o < g.O) {
    if (i0 + ti < g.I) g.dW[(size_t)o * g.I + i0 + ti] = acc;
    if (i0 == 0 && ti == 0) g.db[o] = accb;
  }
}

int launch_head(const HeadParams& p, const HeadGrads* grads, cudaStream_t stream) {
  VPD_REQUIRE(p.F <= 4096 && p.D <= 512 && p.T <= 1024 && p.Hd <= 512, "head: dims too large");
  VPD_REQUIRE(p.motion || p.T == p.D, "head: target dim must equal emb_dim without decoder");
  if (p.B == 0) return 0;
  const int smem = (p.F + 2 * p.D + 4 * p.Hd + p.T) * sizeof(float);
  VPD_CHECK_CUDA(launch_kernel(head_kernel, dim3(p.B), dim3(kHeadThreads), smem, stream, p));
  VPD_LAUNCHED(1);
  if (grads == nullptr || p.dz == nullptr) return 0;
  // workspace layout per frame: pooled[F] e[D] de[D] h1[Hd] h2[Hd] dh1[Hd] dh2[Hd] dO[T]
  OuterParams op;
  op.B = p.B;
  op.stride = head_ws_stride(p.F, p.D, p.Hd, p.T);
  const float* w = p.ws;
  const float* pooled = w;
  const float* e = w + p.F;
  const float* de = e + p.D;
  const float* h1 = de + p.D;
  const float* h2 = h1 + p.Hd;
  const float* dh1 = h2 + p.Hd;
  const float* dh2 = dh1 + p.Hd;
  const float* dO = dh2 + p.Hd;
  int n = 0, total = 0;
  op.seg[n++] = OuterSeg{de, pooled, grads->fc_w, grads->fc_b, p.D, p.F};
  if (p.motion) {
    op.seg[n++] = OuterSeg{dh1, e, grads->w0, grads->b0, p.Hd, p.D};
    op.seg[n++] = OuterSeg{dh2, h1, grads->w2, grads->b2, p.Hd, p.Hd};
    op.seg[n++] = OuterSeg{dO, h2, grads->w5, grads->b5, p.T, p.Hd};
  }
  op.nseg = n;
  for (int i = 0; i < n; ++i)
    total += ((op.seg[i].O + kOwTileO - 1) / kOwTileO) * ((op.seg[i].I + kOwTileI - 1) / kOwTileI);
  VPD_CHECK_CUDA(launch_kernel(head_wgrad_kernel, dim3(total), dim3(256), 0, stream, op));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
