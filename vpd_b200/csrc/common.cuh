// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM
// PTX wrappers and small vector utilities. Everything here is hand-written
// inline PTX (no CUTLASS/CuTe dependency); the descriptor bit layouts follow
// the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define VPD_DEVINL __device__ __forceinline__

namespace vpd {

// ------------------------------------------- programmatic dependent launch (PDL)
// Every kernel triggers its dependents at entry (the next kernel in the stream may
// be scheduled as soon as all CTAs of this one have started, hiding its launch
// latency and prologue) and waits for the previous kernel's completion + memory
// flush before it touches any global data.
VPD_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
VPD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// --------------------------------------------- order-independent statistics accumulators
// Per-channel sums that many CTAs add to (BatchNorm batch statistics, BN-backward sums) are
// kept as TWO 64-bit INTEGER limbs: value = hi * 2^-4 + lo * 2^-52. Integer atomic adds are
// associative, so the total does not depend on the order in which the CTAs arrive: with the
// fixed tile -> CTA assignment of the kernels the statistics, and everything downstream of
// them, are bit-reproducible from run to run (fp64 atomics are not: a last-bit change of a
// mean flips bf16 roundings that the quantised network then amplifies to its noise floor).
// A contribution x (an fp32 partial sum widened to fp64) is split exactly: hi = rint(16 x),
// the remainder (< 2^-5, exact in fp64) is quantised to 2^-52. Range per contribution
// |x| < 2^58; resolution 2.2e-16 absolute, i.e. an fp32 partial >= 4e-9 keeps all its bits.
struct __align__(16) StatAcc {
  long long hi, lo;
};
VPD_DEVINL void stat_add(StatAcc* a, double x) {
  const long long hi = __double2ll_rn(x * 16.0);
  const double rem = x - static_cast<double>(hi) * 0.0625;
  const long long lo = __double2ll_rn(rem * 4503599627370496.0);
  atomicAdd(reinterpret_cast<unsigned long long*>(&a->hi), static_cast<unsigned long long>(hi));
  atomicAdd(reinterpret_cast<unsigned long long*>(&a->lo), static_cast<unsigned long long>(lo));
}
VPD_DEVINL double stat_value(longlong2 v) {
  return fma(static_cast<double>(v.y), 2.220446049250313e-16, static_cast<double>(v.x) * 0.0625);
}
// read-only path (sums produced by an EARLIER kernel)
VPD_DEVINL double stat_read(const StatAcc* a) {
  return stat_value(__ldg(reinterpret_cast<const longlong2*>(a)));
}
// through L2 (sums produced earlier in THIS kernel by other CTAs, behind a grid barrier)
VPD_DEVINL double stat_read_cg(const StatAcc* a) {
  return stat_value(__ldcg(reinterpret_cast<const longlong2*>(a)));
}

// ---------------------------------------------------------------- addresses
VPD_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------- mbarrier
VPD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
VPD_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
VPD_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
VPD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
VPD_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VPD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
VPD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------------- TMA
VPD_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
VPD_DEVINL void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
VPD_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// shared -> global tile store (bulk async-group completion); out-of-range box elements are
// clipped by the hardware
VPD_DEVINL void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::
          "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
VPD_DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING their shared-memory source
VPD_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
VPD_DEVINL void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// ... have completed (writes performed)
VPD_DEVINL void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// one cache line -> L2
VPD_DEVINL void prefetch_l2(const void* g) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
}
// asynchronous L2 prefetch of a contiguous global range (bytes % 16 == 0)
VPD_DEVINL void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gsrc)),
               "r"(bytes)
               : "memory");
}

// 1-D bulk copy global -> shared (contiguous bytes, size % 16 == 0), completing on an mbarrier
VPD_DEVINL void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes),
      "r"(smem_u32(bar))
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative smem offset of every CTA
// in `cta_mask`, and each destination's mbarrier (same offset) gets the complete_tx
VPD_DEVINL void tma_load_3d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                  int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "h"(cta_mask)
      : "memory");
}

// One lane of a converged warp. Unlike `lane == 0`, the compiler knows the guarded region
// runs on a single thread, so uniform-datapath instructions (UTCHMMA, UTMALDG, UTCBAR)
// issue directly instead of inside a per-active-lane broadcast loop.
VPD_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Per-warpgroup register re-allocation (all 4 warps of an aligned warpgroup must execute it)
template <int N>
VPD_DEVINL void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
VPD_DEVINL void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ------------------------------------------------------------------ clusters
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
VPD_DEVINL uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive on an mbarrier that lives in another CTA of the cluster
VPD_DEVINL void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
VPD_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
VPD_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------ tcgen05 / TMEM
VPD_DEVINL void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
VPD_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VPD_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// ---- cta_group::2: the two CTAs of a cluster pair act as one 256-row MMA unit
VPD_DEVINL void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
VPD_DEVINL void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
VPD_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// issued by the leader CTA only: D[256 x N] += A[256 x 16] * B[N x 16]^T, rows 0-127 of A/D
// and the first N/2 rows of B in the leader's smem/TMEM, the rest in the peer's (same offsets)
VPD_DEVINL void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (when the pair's MMAs retire) on the barrier at this offset in both CTAs
VPD_DEVINL void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

VPD_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
VPD_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, one CTA.
VPD_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread retire.
VPD_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `cta_mask`
VPD_DEVINL void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base_lane+i).
VPD_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
VPD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, 128-byte swizzle, dense [rows][128 B] tile.
//   K-major  operand: rows = M/N index, 64 bf16 of K per row; SBO = 1024 B
//                     between 8-row groups; LBO unused (set to 1).
//   MN-major operand: rows = K index, 64 bf16 of M/N per row; SBO = 1024 B
//                     between 8-row K groups; LBO = byte distance between
//                     consecutive 64-wide M/N blocks.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
//       [61,64) layout (2 = SWIZZLE_128B)
VPD_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9,
// 10-12 = 1), a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major),
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// Weight mirrors are stored PRE-TILED in the exact shared-memory image the tensor core
// reads: per tap, per 64-wide K chunk, per 64-row block: [64 rows][64 k] bf16 with the
// 16-byte chunks of a row XOR-swizzled by (row % 8). A [BLOCK_N rows x 64 k] operand tile
// is then one contiguous run of BLOCK_N*128 bytes (a single bulk copy instead of BLOCK_N
// strided 128-byte TMA rows). Offset (in elements) inside one tap's [rows][cols] block:
__host__ __device__ inline long long wtile_offset(int row, int col, int rows) {
  const int kc = col >> 6, cc = col & 63, rb = row >> 6, rr = row & 63;
  const int chunk = (cc >> 3) ^ (rr & 7);
  return ((long long)(kc * (rows >> 6) + rb) * 64 + rr) * 64 + chunk * 8 + (cc & 7);
}

// Network input ("stem") layout: the zero-padded image (3 rows / columns of conv padding on
// the top / left, 8 channel slots per pixel) stored SPACE-TO-DEPTH 2 x 4: cell (i, j) holds
// padded rows 2i, 2i+1 x padded columns 4j .. 4j+3 as 64 contiguous bf16 (128 bytes),
// element (a*4 + q)*8 + c. [frame][Hs = (H+6+1)/2][Ws = (W+6+3)/4][64]; for 128 x 128 crops
// 67 x 34 cells = the same 291,584 bytes per frame as a plain padded [H+6][W+8][8] image.
// Why: the 7x7 / stride-2 stem then is two STRIDE-1 convolutions over cells (even / odd
// output columns, 4 x 2 and 4 x 3 taps of K = 64), i.e. exactly what the halo-reuse kernels
// do - one aligned TMA patch per tile instead of seven 128-row tiles of 32-byte-misaligned,
// overlapping windows (which ran at a quarter of the kernel's own bound).
__host__ __device__ inline int stem_cells_h(int H) { return (H + 6 + 1) / 2; }
__host__ __device__ inline int stem_cells_w(int W) { return (W + 6 + 3) / 4; }
// element offset of channel slot 0 of padded pixel (ph, pw) inside one frame
__host__ __device__ inline long long stem_pixel_offset(int ph, int pw, int Ws) {
  return ((long long)(ph >> 1) * Ws + (pw >> 2)) * 64 + ((ph & 1) * 4 + (pw & 3)) * 8;
}

// ------------------------------------------------------------- small helpers
VPD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
VPD_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
VPD_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
// 0xFFFF in each half whose bf16 value is > 0 (NaN -> 0), for masking a packed pair
VPD_DEVINL uint32_t bf16x2_gt0_mask(uint32_t v) {
  const __nv_bfloat162 z = *reinterpret_cast<const __nv_bfloat162*>(&v);
  return __hgt2_mask(z, __floats2bfloat162_rn(0.f, 0.f));
}
VPD_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// PTX prmt (default mode): selector nibble bits 0-2 pick a byte of {b, a}, bit 3 replicates
// that byte's sign bit over the whole output byte (__byte_perm documents only the low 3 bits)
VPD_DEVINL uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
VPD_DEVINL uint32_t ldg_nc_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
VPD_DEVINL uint32_t ldg_nc_u8(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
VPD_DEVINL uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// coherent 16-byte global load (data written earlier in this kernel by other warps of the CTA)
VPD_DEVINL uint4 ldg_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
VPD_DEVINL void stg_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
VPD_DEVINL uint2 ldg_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
VPD_DEVINL void stg_v2(void* p, uint2 v) {
  asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
VPD_DEVINL void stg_cs_v2(void* p, uint2 v) {  // streaming store (evict-first)
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
VPD_DEVINL void stg_cs_v4(void* p, uint4 v) {  // streaming store (evict-first)
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Explicit shared-space accesses. Pointers carved out of the dynamic smem buffer via
// integer arithmetic are GENERIC to the compiler: plain loads/stores become LD.E/ST.E
// and atomicAdd becomes a CAS loop; these wrappers emit LDS/STS/RED.shared instead.
VPD_DEVINL void sts_v4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
VPD_DEVINL uint4 lds_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(saddr)
               : "memory");
  return r;
}
VPD_DEVINL float4 __uint4_as_float4(uint4 v) {
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
VPD_DEVINL float lds_f32(uint32_t saddr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(saddr) : "memory");
  return r;
}
VPD_DEVINL void red_shared_add(uint32_t saddr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

VPD_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vpd
