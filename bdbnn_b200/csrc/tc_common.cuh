// Shared tcgen05 / TMA / mbarrier PTX wrappers and tensor-map helpers (sm_100a).
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace bdbnn {

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >>4
// in [0,14), LBO>>4 in [16,30), SBO>>4 in [32,46), version=1 in [46,48), layout type in [61,64).
// For swizzled K-major tiles LBO is unused; SBO = bytes between 8-row groups = 8 * row_bytes.
// NOTE (measured on B200): a start address that is a whole number of 128-byte rows into a
// 1024B-aligned SWIZZLE_128B tile needs base_offset = 0 — the hardware applies the swizzle XOR to the
// absolute shared-memory address bits, so row-shifted views of one TMA-written tile are valid operands.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);  // SW128 / SW64 / SW32
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;                          // LBO (ignored for swizzled K-major)
  d |= uint64_t((8u * row_bytes) >> 4) << 32;      // SBO
  d |= uint64_t(1) << 46;                          // descriptor version (Blackwell)
  d |= uint64_t(layout) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// K-major A and B (0) @15/@16, N>>3 @17, M>>4 @24.
// fmt: BDBNN_FMT_FP16 (0) or BDBNN_FMT_BF16 (1) for both A and B (kind::f16).
__host__ __device__ inline uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t fmt = 1u) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

constexpr int kTcThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;
constexpr int kMaxTaps = 49;

// Four K=16 steps of one 64-element (128-byte, SWIZZLE_128B K-major) K block in a single asm block:
// the MMA-issuing thread is a lone scalar thread, so per-MMA instruction count is what bounds the
// issue rate (a generic descriptor rebuild per MMA costs ~80 clk, more than a 128x64x16 MMA takes).
// a_lo/b_lo are the low descriptor words (address>>4 | LBO), desc_hi the shared high word.
__device__ __forceinline__ void umma_bf16_k4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                             uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 al, %1, 2;\n\t"
      "add.u32 bl, %2, 2;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
      "add.u32 al, %1, 4;\n\t"
      "add.u32 bl, %2, 4;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
      "add.u32 al, %1, 6;\n\t"
      "add.u32 bl, %2, 6;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// One MMA from pre-split descriptor words (cheap to advance: only the low words change per K step).
__device__ __forceinline__ void umma_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp8 (e4m3 x e4m3 -> f32) MMA, K = 32 per instruction; +-1 is exact in e4m3.
__device__ __forceinline__ void umma_f8_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// e4m3 A and B (format code 0), f32 accumulate, K-major
__host__ __device__ inline uint32_t make_idesc_f8(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Low / high descriptor words of a swizzled K-major tile with rows of `row_bytes` (64 or 128)
__device__ __forceinline__ uint32_t kmajor_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t kmajor_hi(uint32_t row_bytes) {
  return ((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
}
// Low / high words of a SWIZZLE_128B K-major descriptor (see make_kmajor_desc).
__device__ __forceinline__ uint32_t kmajor128_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t kmajor128_hi() { return (1024u >> 4) | (1u << 14) | (2u << 29); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) helpers: two CTAs of a 2-CTA cluster (same TPC) execute ONE tcgen05.mma of M = 256 —
// each SM multiplies its own 128 A rows (its own shared memory, same descriptor offsets in both CTAs) with the
// full B tile, of which each CTA holds half the rows (N/2) — so every SM fetches half as many B bytes from its
// shared memory per unit of work.  Only the leader CTA (cluster rank 0) issues MMAs and commits; barriers that the
// MMA thread waits on live in the leader, and the peer signals them through shared::cluster addresses.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the PEER CTA (the leader's full barrier)
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(uint16_t(3))
      : "memory");
}
// One M=256 MMA of the pair from pre-split descriptor words; KIND 0 = f16/bf16 (K = 16), 1 = f8f6f4 (K = 32)
template <int KIND>
__device__ __forceinline__ void umma_split_cg2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side: tensor-map construction and launch
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 NHWC activation tensor [N][H][W][C] -> 4-D map, box [BNI][BH][BW][KB].
// esize = 2: 16-bit elements (bf16/fp16 bit patterns), esize = 1: fp8 bytes.
inline int make_act_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int KB, int BW,
                        int BH, int BNI, int step = 1, int esize = 2) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return BDBNN_ERR_CUDA; }
  cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  cuuint64_t strides[3] = {cuuint64_t(C) * esize, cuuint64_t(W) * C * esize, cuuint64_t(H) * W * C * esize};
  // element stride `step` in w/h: TMA loads ceil(box/step) elements, so box = loaded * step
  cuuint32_t box[4] = {cuuint32_t(KB), cuuint32_t(BW * step), cuuint32_t(BH * step), cuuint32_t(BNI)};
  cuuint32_t estr[4] = {1, cuuint32_t(step), cuuint32_t(step), 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(KB * esize),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act) failed: %d", int(r)); return BDBNN_ERR_CUDA; }
  return BDBNN_OK;
}

// Overlapping-window view of a zero-padded 16-bit NHWC image (stem conv, stem.cu): element
// (j, wo, hp, n) lives at base + n*img_stride + hp*row_stride + wo*win_stride + 2*j (bytes), i.e. window
// `wo` of a padded row is the KB consecutive values starting win_stride bytes after window wo-1 —
// consecutive windows overlap; TMA only needs every stride to be a multiple of 16 bytes.
// Box [BNI][BH (every hstep-th row)][BW][KB].
inline int make_window_map(CUtensorMap* map, const void* base, int N, int HP, int WO, int KB,
                           uint64_t win_stride, uint64_t row_stride, uint64_t img_stride, int BW, int BH,
                           int BNI, int hstep) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return BDBNN_ERR_CUDA; }
  cuuint64_t dims[4] = {cuuint64_t(KB), cuuint64_t(WO), cuuint64_t(HP), cuuint64_t(N)};
  cuuint64_t strides[3] = {win_stride, row_stride, img_stride};
  cuuint32_t box[4] = {cuuint32_t(KB), cuuint32_t(BW), cuuint32_t(BH * hstep), cuuint32_t(BNI)};
  cuuint32_t estr[4] = {1, 1, cuuint32_t(hstep), 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(KB * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(window) failed: %d", int(r)); return BDBNN_ERR_CUDA; }
  return BDBNN_OK;
}

// bf16 K-major weight matrix [rows][cols] -> 2-D map, box [BN rows][KB cols].
inline int make_weight_map(CUtensorMap* map, const void* base, int rows, int cols, int KB, int BN, int esize = 2) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return BDBNN_ERR_CUDA; }
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(cols) * esize};
  cuuint32_t box[2] = {cuuint32_t(KB), cuuint32_t(BN)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(KB * esize),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weight) failed: %d", int(r)); return BDBNN_ERR_CUDA; }
  return BDBNN_OK;
}


// One implicit-GEMM launch: D[NIMG x OH x OW pixels, Nout] = sum_taps A[pixel*in_step + tap offset] * B.
struct TcConvLaunch {
  const uint16_t* A; int IH, IW, Kc, a_halves, in_step;
  const uint16_t* B; int b_taps, Nout;
  int NIMG, OH, OW;
  int n_taps; int8_t dh[kMaxTaps], dw[kMaxTaps]; uint8_t tb[kMaxTaps];
  int out_step, off_h, off_w, OHf, OWf;
  const float* alpha; const uint32_t* mask; float* out;
  int16_t* out_i16;              // forward (persistent kernel only): write the integer accumulator as int16 instead of out
  int fmt;                       // operand format (BDBNN_FMT_*); -1 = fp8 e4m3 bytes (forward only)
  const uint32_t* amax_bits;     // FP16S gradient: device word with max|A|; epilogue multiplies by 2^-e
  const float* add;              // dgrad: optional tensor added to the result (shortcut gradient), or NULL
  // forward only: per-output-channel BatchNorm statistics of the result, accumulated by the epilogue
  // (sum y, sum y^2 as doubles [2*Nout]; max|y| bits [Nout]); NULL = off.  Zeroed by the caller.
  double* bn_sums;
  uint32_t* bn_ymax;
  // dgrad only (persistent kernel, stride 1): bn_sums / bn_ymax receive the BACKWARD statistics of the BatchNorm
  // unit that produced this conv's input — sum gz, sum gz*yhat, max|gz| with gz = this launch's result — using
  // that unit's int16 conv result st_y [pixels][Nout] and per-channel alpha / mean / invstd
  const int16_t* st_y; const float* st_alpha; const float* st_mean; const float* st_invstd;
  // stem conv: A is an overlapping-window view (make_window_map) instead of a dense NHWC tensor
  int win;                       // 0 = dense NHWC
  uint64_t win_stride, win_row_stride, win_img_stride;
};

// Geometry of the stem's packed input (stem.cu): fp16 [N][HP][WP][4] zero-padded by 3 on every side,
// windows of 8 pixels (32 values) every 2 pixels.
struct StemGeom {
  int N, H, W, Ho, Wo, HP, WP;
  uint64_t win_stride, row_stride, img_stride;   // bytes
};
inline bool stem_geom(int N, int H, int W, StemGeom* g) {
  if (N < 1 || H < 7 || W < 7) return false;
  g->N = N; g->H = H; g->W = W;
  g->Ho = (H + 6 - 7) / 2 + 1; g->Wo = (W + 6 - 7) / 2 + 1;
  if (g->Wo > kTileM || g->Wo < 1) return false;
  g->HP = H + 6;
  int wp = W + 6 > 2 * g->Wo + 6 ? W + 6 : 2 * g->Wo + 6;
  g->WP = (wp + 7) & ~7;
  g->win_stride = 16;                               // 2 pixels x 4 halves x 2 bytes
  g->row_stride = uint64_t(g->WP) * 8;
  g->img_stride = g->row_stride * uint64_t(g->HP);
  return true;
}
constexpr int kStemCout = 64, kStemTaps = 7, kStemWin = 32;
int launch_stem_fwd(const uint16_t* xw, const uint16_t* wf, const float* alpha, float* y, const StemGeom& g,
                    double* bn_sums, uint32_t* bn_ymax, cudaStream_t st);
// zero the statistics buffers of a conv launch (no-op when off)
inline int bn_stats_zero(double* sums, uint32_t* ymax, int C, cudaStream_t st) {
  if (!sums) return BDBNN_OK;
  BDBNN_CUDA(cudaMemsetAsync(sums, 0, size_t(2 * C) * sizeof(double), st));
  BDBNN_CUDA(cudaMemsetAsync(ymax, 0, size_t(C) * sizeof(uint32_t), st));
  return BDBNN_OK;
}
size_t stem_wgrad_workspace_bytes(const StemGeom& g);
int launch_stem_wgrad(const uint16_t* gys, const uint16_t* xw, float* ws, size_t ws_bytes, int* ksplit_out,
                      const StemGeom& g, cudaStream_t st);

// Persistent multi-accumulator kernel (tc_conv2.cu). Returns BDBNN_ERR_UNSUPPORTED if the geometry
// does not qualify (caller falls back to the one-tile-per-CTA kernel in tc_conv.cu).
int launch_tc_conv2(const TcConvLaunch& L, int mode, cudaStream_t st);
// Pixel-N kernel for the 64-channel layers (tc_conv64.cu): forward and stride-1 dgrad with Kc = Nout = 64 on images
// of more than 256 pixels; BDBNN_ERR_UNSUPPORTED otherwise (callers then use launch_tc_conv2).
int launch_tc_conv64(const TcConvLaunch& L, int mode, cudaStream_t st);
bool tc_conv64_eligible(const TcConvLaunch& L);
int launch_tc_conv64_stem(const TcConvLaunch& L, cudaStream_t st);
void set_tc_trace(long long* buf);
void set_conv_plan_sink(int32_t* sink);   // host-only planning probe, see tc_conv2.cu
long long* get_tc_trace();
bool wgrad_tc_ok(const bdbnn_conv_shape* s);

}  // namespace bdbnn
