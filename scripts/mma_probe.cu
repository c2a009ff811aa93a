// Microbenchmark: what does ONE tcgen05.mma cost in SS mode on this GPU, free of TMA / epilogue / barriers?
// One CTA per SM, operands resident in shared memory (SWIZZLE_128B K-major tiles, never reloaded), a single thread
// issues `reps` x `chain` MMAs round-robin over `nacc` TMEM accumulators and commits once; clock64 around issue+drain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bdbnn_b200/csrc -I include scripts/mma_probe.cu -o /tmp/mma_probe
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace bdbnn;
namespace bdbnn { void set_error(const char*, ...) {} }

// converged issue: the whole warp runs the loop, one elected lane issues (operands stay warp-uniform -> uniform
// registers, no per-MMA ELECT / R2UR.BROADCAST waterfall loop)
template <int KIND>
__device__ __forceinline__ void umma_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0)
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(int N, int nacc, int reps, int f8, int a_rows_shift, long long* out, int converged) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 128 rows (+ shift) x 128 B; B: N rows x 128 B (one 64-element fp16 / 128-element fp8 K block)
  const uint32_t a_addr = base, b_addr = base + 64 * 1024;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&done_bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (converged && threadIdx.x < 32) {
    const uint32_t idesc = f8 ? make_idesc_f8(128, uint32_t(N)) : make_idesc_bf16(128, uint32_t(N), 0u);
    const uint32_t hi = kmajor_hi(128u);
    const uint32_t a_lo = kmajor_lo(a_addr) + uint32_t(a_rows_shift) * 8u, b_lo = kmajor_lo(b_addr);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int j = 0; j < nacc; ++j) {
        const uint32_t acc = tmem_d + uint32_t(j * N);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (f8) umma_elect<1>(acc, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (r | k) ? 1u : 0u);
          else    umma_elect<0>(acc, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (r | k) ? 1u : 0u);
        }
      }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) umma_commit(smem_u32(&done_bar));
    mbar_wait(smem_u32(&done_bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (!converged && threadIdx.x == 0) {
    const uint32_t idesc = f8 ? make_idesc_f8(128, uint32_t(N)) : make_idesc_bf16(128, uint32_t(N), 0u);
    const uint32_t hi = kmajor_hi(128u);
    const uint32_t a_lo = kmajor_lo(a_addr) + uint32_t(a_rows_shift) * 8u, b_lo = kmajor_lo(b_addr);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int j = 0; j < nacc; ++j) {
        const uint32_t acc = tmem_d + uint32_t(j * N);
        for (int k = 0; k < 4; ++k) {
          if (f8) umma_f8_split(acc, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (r | k) ? 1u : 0u);
          else    umma_bf16_split(acc, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (r | k) ? 1u : 0u);
        }
      }
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&done_bar));
    mbar_wait(smem_u32(&done_bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const size_t smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("# per-MMA cost in clocks (issue-only / issue+drain), 128xNx16 fp16 (or x32 fp8), SS mode, %d CTAs\n", sms);
  printf("%-6s %-4s %-5s %-6s %-5s %10s %10s %12s\n", "kind", "N", "nacc", "shift", "conv", "clk_issue", "clk_total", "MAC/clk/SM");
  for (int f8 = 0; f8 < 2; ++f8)
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2, 4})
        for (int shift : {0, 1}) {
          if (nacc * N > 512) continue;
          const int reps = 512;
          for (int conv : {0, 1}) {
            const int grid = sms;
            probe_kernel<<<grid, 128, smem>>>(N, nacc, reps, f8, shift, d, conv);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double n_mma = double(reps) * nacc * 4;
            const double mac = 128.0 * N * (f8 ? 32 : 16);
            printf("%-6s %-4d %-5d %-6d %-5d %10.1f %10.1f %12.1f\n", f8 ? "fp8" : "fp16", N, nacc, shift, conv,
                   h[0] / n_mma, h[1] / n_mma, mac / (h[1] / n_mma));
          }
        }
  return 0;
}
