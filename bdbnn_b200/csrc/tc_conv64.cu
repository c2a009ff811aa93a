// Pixel-N implicit-GEMM conv for the 64-channel 3x3 layers (ResNet layer1 forward and dgrad): the PIXELS are the
// N = 256 dimension of the MMA and the tap weights are a RESIDENT A operand.
//
// Why (measured, DESIGN.md section 5 finding 17, scripts/mma_probe.cu): one SS-mode tcgen05.mma of 128 x N x 16 costs
// ~104 clk for N = 64 and for N = 128 (the 4 KB A operand is fetched at ~39 B/clk) and 128 clk for N = 256, where it
// is tensor-bound.  The pixel-M kernels (tc_conv2.cu: A = 128 pixels, B = weights, N = Cout) therefore top out at
// 30 % of the tensor peak on the Cout = 64 layers — the layers with the most HBM traffic of the network.  Here
//     D[o, p] = sum over taps t, channels c of  W_t[o, c] * x[p + shift_t, c]
//   A = W_t  : [64 output-channel rows] x 64 channels, K-major SWIZZLE_128B, ALL taps loaded once per CTA (TMA) and
//              kept in shared memory for the whole kernel (9 x 8 KB); the M = 128 descriptor of tap t also covers the
//              64 rows of tap t+1, whose results land in TMEM lanes 64..127 and are ignored
//   B = x    : the halo patch of a unit (256 output pixels in padded-width raster order plus halo rows, one TMA load
//              per 64-channel K block, double-buffered); tap (dh,dw) is the patch viewed from row
//              (dh-dh_min)*PW + (dw-dw_min) — the same absolute-address-swizzle trick as tc_conv2.cu, on the B side
//   D        : [64 channels (TMEM lanes 0..63)] x [256 pixels (columns)] fp32, two accumulators (512 columns) so that
//              the epilogue of unit i overlaps the MMAs of unit i+1
// 36 MMAs of 128 clk per 256 pixels instead of 72 of >= 104 (157 measured): the MMA time of a ResNet-18 layer1 launch
// (N = 256) drops from ~150 us to ~50 us, the neighbourhood of its 47 us HBM time.
// The epilogue has lane = channel and register = pixel: every store / load instruction of a warp touches 32
// consecutive channels of one pixel (one 128-byte NHWC segment), per-channel BatchNorm statistics are plain per-lane
// accumulators, the STE mask word of a pixel is one broadcast load.
// Warp roles (448 threads): warps 0,1 4,5 8,9 12,13 = epilogue (TMEM lane quadrant = warp & 3 in {0,1} — only lanes
// 0..63 carry channels — and column-block group = warp >> 2), warp 2 = TMA producer (one lane), warp 3 = MMA issuer
// (converged warp, elected lane) + TMEM owner; warps 6,7,10,11 only exist to give the epilogue warps their ids.
#include "tc_common.cuh"

namespace bdbnn {

constexpr int kC64Threads = 448;    // 14 warps: see the role list above
constexpr int kC64N = 256;          // pixels (MMA N) per unit
constexpr int kC64MaxTaps = 9;
constexpr uint32_t kC64TapBytes = 64u * 128u;   // one tap's weights: 64 rows x 128 B (64-byte rows in the stem variant)

struct TcConv64Params {
  int32_t OW, OH, NIMG;
  int32_t PW, PH, SH, dh_min, dw_min, units_per_img, n_units;
  uint32_t patch_bytes;
  int32_t n_taps;
  int8_t tap_dh[kC64MaxTaps], tap_dw[kC64MaxTaps];
  uint8_t tap_b[kC64MaxTaps];
  // STEM variant (7x7/2 stem conv over the packed window image, stem.cu): K = 32 halves per tap (64-byte rows), the
  // unit is UH = 2 output rows x 128 windows, and the patch has two row-parity planes (even / odd window rows) so that
  // tap r is plane r&1 viewed from row (r>>1)*PW
  int32_t plane_rows;        // window rows per plane box
  uint32_t plane_bytes;      // smem bytes of one plane
  int32_t a_halves;          // K blocks of the activation operand (2 = bf16 hi|lo gradient)
  int32_t fmt;
  int32_t dbg;               // BDBNN_TC_DBG experiment bits: 1 = no global stores / loads in the epilogue, 2 = no TMEM loads, 4 = no MMAs
  const uint32_t* amax_bits;
  const float* add;
  const float* alpha;
  const uint32_t* mask;
  float* out;
  int16_t* out_i16;
  double* bn_sums;
  uint32_t* bn_ymax;
  const int16_t* st_y;
  const float* st_alpha;
  const float* st_mean;
  const float* st_invstd;
};

// one MMA issued by an elected lane of a converged warp (operands stay in uniform registers)
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}

template <int MODE, bool BST, bool I16, bool STEM = false>
__global__ void __launch_bounds__(kC64Threads, 1)
tc_conv64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const TcConv64Params p) {
  constexpr uint32_t RB = STEM ? 64u : 128u;          // bytes per operand row of one K block
  constexpr uint32_t kTapBytes = 64u * RB;            // one tap's weights: 64 rows
  constexpr int kKSteps = int(RB / 32u);              // K = 16 MMAs per tap and K block
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t pfull_bar[2], pempty_bar[2];
  __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t tap_shift_rows[kC64MaxTaps];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;                                              // (n_taps + 1) tap tiles
  const uint32_t patch_base = smem_base + ((uint32_t(p.n_taps + 1) * kTapBytes + 1023u) & ~1023u);  // two patch buffers
  if (threadIdx.x < p.n_taps) {
    // offset of tap t's view inside a patch buffer, in 16-byte descriptor units
    if (STEM)      // tap r = window row 2*oh + r: plane r & 1, row shift (r >> 1) * PW
      tap_shift_rows[threadIdx.x] = (uint32_t(threadIdx.x & 1) * p.plane_bytes +
                                     uint32_t(threadIdx.x >> 1) * uint32_t(p.PW) * RB) >> 4;
    else
      tap_shift_rows[threadIdx.x] = uint32_t((p.tap_dh[threadIdx.x] - p.dh_min) * p.PW +
                                             (p.tap_dw[threadIdx.x] - p.dw_min)) * (RB >> 4);
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_bar), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&pfull_bar[s]), 1);
      mbar_init(smem_u32(&pempty_bar[s]), 1);
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 8);     // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  // the tile behind the last tap is read by that tap's M = 128 descriptor (rows 64..127, results ignored): keep it finite
  {
    uint8_t* tail = smem_raw + (w_base - smem_u32(smem_raw)) + size_t(p.n_taps) * kTapBytes;
    for (int i = threadIdx.x; i < int(kTapBytes / 16); i += blockDim.x)
      reinterpret_cast<uint4*>(tail)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (warp == 2 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
  }
  if (warp == 3) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int kb_total = p.a_halves;        // one 64-channel K block per operand half

  if (warp == 2) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // all tap weights, once
      const uint32_t wb = smem_u32(&w_bar);
      mbar_expect_tx(wb, uint32_t(p.n_taps) * kTapBytes);
      for (int t = 0; t < p.n_taps; ++t)
        tma_load_2d(w_base + uint32_t(t) * kTapBytes, &tmW, wb, int(p.tap_b[t]) * int(RB / 2u), 0);
      uint32_t pcount = 0;
      const uint32_t patch_tx = STEM ? 2u * uint32_t(p.PW * p.plane_rows) * RB : uint32_t(p.PW * p.PH) * RB;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int n0 = u / p.units_per_img, h0 = (u - n0 * p.units_per_img) * p.SH;
        for (int kb = 0; kb < kb_total; ++kb, ++pcount) {
          const uint32_t pa = pcount & 1u;
          mbar_wait(smem_u32(&pempty_bar[pa]), ((pcount >> 1) & 1u) ^ 1u);
          const uint32_t pb = smem_u32(&pfull_bar[pa]);
          mbar_expect_tx(pb, patch_tx);
          const uint32_t dst = patch_base + pa * p.patch_bytes;
          if (STEM) {     // even and odd window rows of the unit (the map steps two rows per box row)
            tma_load_4d(dst, &tmX, pb, 0, 0, 2 * h0, n0);
            tma_load_4d(dst + p.plane_bytes, &tmX, pb, 0, 0, 2 * h0 + 1, n0);
          } else {
            tma_load_4d(dst, &tmX, pb, kb * 64, p.dw_min, h0 + p.dh_min, n0);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ================================ MMA issuer (converged warp) ================================
    const uint32_t idesc = make_idesc_bf16(128u, uint32_t(kC64N), uint32_t(p.fmt));
    const uint32_t desc_hi = kmajor_hi(RB);
    mbar_wait(smem_u32(&w_bar), 0);
    tc_fence_after();
    uint32_t pcount = 0, ucount = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ucount) {
      const uint32_t buf = ucount & 1u;
      mbar_wait(smem_u32(&tempty_bar[buf]), ((ucount >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_d + buf * uint32_t(kC64N);
      uint32_t first = 0u;        // 0 for the very first MMA of the unit (overwrite), 1 afterwards
      for (int kb = 0; kb < kb_total; ++kb, ++pcount) {
        const uint32_t pa = pcount & 1u;
        mbar_wait(smem_u32(&pfull_bar[pa]), (pcount >> 1) & 1u);
        tc_fence_after();
        const uint32_t patch_lo = kmajor_lo(patch_base + pa * p.patch_bytes);
        for (int t = 0; t < p.n_taps; ++t) {
          const uint32_t a_lo = kmajor_lo(w_base + uint32_t(t) * kTapBytes);
          const uint32_t b_lo = patch_lo + tap_shift_rows[t];
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < kKSteps; ++k) {
              umma_f16_elect(acc, a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, idesc, first);
              first = 1u;
            }
          }
        }
        umma_commit_elect(smem_u32(&pempty_bar[pa]));
      }
      umma_commit_elect(smem_u32(&tfull_bar[buf]));
    }
  } else if ((warp & 3) < 2) {
    // ================================ epilogue (warps 0,1 4,5 8,9 12,13) ================================
    const int quad = warp & 3;          // TMEM lanes quad*32 .. +31 = channels quad*32 + lane
    const int grp = warp >> 2;          // handles the 32-pixel column blocks cb with (cb & 3) == grp
    const int ch = quad * 32 + lane;
    const float post = (MODE == 1 && p.amax_bits) ? amax_pow2_scale(__ldg(p.amax_bits), true) : 1.0f;
    const float alpha = MODE == 0 ? __ldg(p.alpha + ch) : 1.0f;
    const bool do_stats = p.bn_sums != nullptr;
    float st_a = 0.f, st_b = 0.f;       // BST: yhat = y_int * st_a - st_b
    if (BST && do_stats) {
      const float is = __ldg(p.st_invstd + ch);
      st_a = __ldg(p.st_alpha + ch) * is;
      st_b = __ldg(p.st_mean + ch) * is;
    }
    // per-channel statistics of this lane.  Forward: the conv result is an exact integer, so sum y_int and
    // sum y_int^2 are accumulated EXACTLY in integers (|y_int| <= 576: 64-bit sums cannot overflow) and scaled by
    // alpha / alpha^2 once at the end.  Backward (BST): fp32 within a unit, fp64 across units.
    long long i_sum = 0;
    unsigned long long i_sq = 0;
    int i_max = 0;
    double d_sum = 0.0, d_sq = 0.0;
    float f_max = 0.f;
    uint32_t ucount = 0;
    const int row_elems = p.OW * 64;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ucount) {
      const int n0 = u / p.units_per_img, h0 = (u - n0 * p.units_per_img) * p.SH;
      const uint32_t buf = ucount & 1u;
      mbar_wait(smem_u32(&tfull_bar[buf]), (ucount >> 1) & 1u);
      tc_fence_after();
      const uint32_t tbase = tmem_d + (uint32_t(quad * 32) << 16) + buf * uint32_t(kC64N);
      const int rows_ok = (p.dbg & 1) ? 0 : min(p.SH, p.OH - h0);  // output rows of this unit
      int u_isum = 0;                  // forward: exact 32-bit partials of one unit (<= 256 pixels of |y_int| <= 576)
      uint32_t u_isq = 0;
      float u_sum = 0.f, u_sq = 0.f;   // backward statistics of one unit
      for (int cbi = 0; cbi < kC64N / 128; ++cbi) {
        if (p.dbg & 2) break;
        const int c0 = (4 * cbi + grp) * 32;                       // first pixel column of this block
        // The padded width PW is a multiple of 32, so a block of 32 raster columns lies inside ONE unit row:
        // image row h0 + hi, image columns wi0 .. wi0 + 31, of which the first nv are outputs.
        const int hi = c0 / p.PW, wi0 = c0 - hi * p.PW;
        const int nv = hi < rows_ok ? min(32, max(0, p.OW - wi0)) : 0;
        if (nv == 0) continue;
        // element offset (32-bit, checked on the host) of pixel 0 of the block for this lane's channel
        const int off = ((n0 * p.OH + h0 + hi) * p.OW + wi0) * 64 + ch;
        uint32_t mword = 0;
        if (MODE == 1)   // lane j fetches the STE-mask word of pixel j (this warp's 32 channels); broadcast by shuffle below
          mword = lane < nv ? __ldg(p.mask + int64_t((off - ch) >> 6) * 2 + lane * 2 + quad) : 0u;
        // pixels per pass: 32, or 16 with the BatchNorm backward sums (v + add + y of 32 pixels would not fit the
        // register budget of 448 threads: the first version spilled and ran at 0.4 ms per launch)
        constexpr int NPB = BST ? 16 : 32;
#pragma unroll
        for (int hb = 0; hb < 32 / NPB; ++hb) {
          const int p0 = hb * NPB;                                 // first pixel of this pass inside the block
          if (p0 >= nv) break;                                     // warp-uniform
          uint32_t v[NPB];
          if constexpr (NPB == 32) tmem_ld32(tbase + uint32_t(c0), v);
          else tmem_ld16(tbase + uint32_t(c0 + p0), v);
          float addv[NPB];
          uint32_t yv[BST ? NPB : 1];
          if (MODE == 1) {
            if (p.add != nullptr) {
#pragma unroll
              for (int j = 0; j < NPB; ++j) addv[j] = p0 + j < nv ? p.add[off + (p0 + j) * 64] : 0.f;   // may alias out
            }
            if (BST) {
              // plain ld.global, NOT ld.global.nc / __ldg: a non-coherent load may be moved past the stores below, and
              // ptxas sank each one next to its use — one exposed DRAM latency per pixel, 0.38 ms per launch.  An
              // ordinary load may alias the stores, so all NPB of them are issued here, ahead of the TMEM wait.
#pragma unroll
              for (int j = 0; j < NPB; ++j) {
                uint16_t t = 0;
                if (p0 + j < nv)
                  asm volatile("ld.global.u16 %0, [%1];" : "=h"(t) : "l"(p.st_y + off + (p0 + j) * 64) : "memory");
                yv[j] = t;
              }
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < NPB; ++jj) {
            const int j = p0 + jj;
            const bool ok = j < nv;                                  // warp-uniform
            float o = __uint_as_float(v[jj]);
            if (MODE == 0 && STEM) {        // real-valued result: fp32 statistics per unit, fp64 across units
              const float yv_ = ok ? o * alpha : 0.f;
              if (ok) p.out[off + j * 64] = yv_;
              u_sum += yv_; u_sq = fmaf(yv_, yv_, u_sq); f_max = fmaxf(f_max, fabsf(yv_));
            } else if (MODE == 0) {
              const int yi = ok ? int(o) : 0;
              if (ok) {
                if (I16) p.out_i16[off + j * 64] = int16_t(yi);
                else p.out[off + j * 64] = o * alpha;
              }
              u_isum += yi;
              u_isq += uint32_t(yi * yi);
              i_max = max(i_max, abs(yi));
            } else {
              const uint32_t wordj = __shfl_sync(0xffffffffu, mword, j);
              o = ((wordj >> lane) & 1u) ? o * post : 0.f;
              if (p.add != nullptr) o += addv[jj];
              if (ok) p.out[off + j * 64] = o;
              if (BST) {
                const float g = ok ? o : 0.f;
                const float yh = fmaf(float(int16_t(yv[jj])), st_a, -st_b);
                u_sum += g; u_sq = fmaf(g, yh, u_sq); f_max = fmaxf(f_max, fabsf(g));
              }
            }
          }
        }
      }
      if (MODE == 0 && !STEM) { i_sum += u_isum; i_sq += u_isq; }
      if (BST || STEM) { d_sum += double(u_sum); d_sq += double(u_sq); }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
    }
    if (do_stats) {
      if (MODE == 0 && STEM) {
        if (f_max != 0.f) {
          atomicAdd(p.bn_sums + ch, d_sum);
          atomicAdd(p.bn_sums + 64 + ch, d_sq);
          atomicMax(p.bn_ymax + ch, __float_as_uint(f_max));
        }
      } else if (MODE == 0) {
        if (i_max != 0) {
          const double al = double(alpha);
          atomicAdd(p.bn_sums + ch, al * double(i_sum));
          atomicAdd(p.bn_sums + 64 + ch, al * al * double(i_sq));
          atomicMax(p.bn_ymax + ch, __float_as_uint(fabsf(alpha) * float(i_max)));
        }
      } else if (BST && (f_max != 0.f || d_sum != 0.0 || d_sq != 0.0)) {
        atomicAdd(p.bn_sums + ch, d_sum);
        atomicAdd(p.bn_sums + 64 + ch, d_sq);
        atomicMax(p.bn_ymax + ch, __float_as_uint(f_max));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 512);
  }
}

static int c64_enabled() {
  static const int enabled = [] { const char* e = getenv("BDBNN_TC_C64"); return e ? atoi(e) : 1; }();
  return enabled;
}

// Host-only: would launch_tc_conv64 take this launch (ignoring the statistics option)?
bool tc_conv64_eligible(const TcConvLaunch& L) {
  if (!c64_enabled() || L.win || L.fmt < 0) return false;
  if (L.Kc != 64 || L.Nout != 64 || L.in_step != 1 || L.out_step != 1 || L.off_h != 0 || L.off_w != 0) return false;
  if (L.n_taps < 1 || L.n_taps > kC64MaxTaps || (L.a_halves != 1 && L.a_halves != 2)) return false;
  if (L.OHf != L.OH || L.OWf != L.OW) return false;
  int dw0 = 127, dw1 = -127;
  for (int i = 0; i < L.n_taps; ++i) { dw0 = min(dw0, int(L.dw[i])); dw1 = max(dw1, int(L.dw[i])); }
  const int PW = (L.OW + (dw1 - dw0) + 31) & ~31;
  if (PW > 128 || L.OH * L.OW <= kC64N) return false;
  return int64_t(L.NIMG) * L.OH * L.OW * 64 < (int64_t(1) << 31);
}

// Returns BDBNN_ERR_UNSUPPORTED when the geometry does not qualify (the caller falls back to tc_conv2.cu).
int launch_tc_conv64(const TcConvLaunch& L, int mode, cudaStream_t st) {
  const int enabled = c64_enabled();
  if (!tc_conv64_eligible(L)) return BDBNN_ERR_UNSUPPORTED;
  if (mode == 0 && L.out_i16 != nullptr && L.n_taps * 64 > 32767) return BDBNN_ERR_UNSUPPORTED;
  TcConv64Params p;
  memset(&p, 0, sizeof(p));
  int dh0 = 127, dh1 = -127, dw0 = 127, dw1 = -127;
  for (int i = 0; i < L.n_taps; ++i) {
    p.tap_dh[i] = L.dh[i]; p.tap_dw[i] = L.dw[i]; p.tap_b[i] = L.tb[i];
    dh0 = min(dh0, int(L.dh[i])); dh1 = max(dh1, int(L.dh[i]));
    dw0 = min(dw0, int(L.dw[i])); dw1 = max(dw1, int(L.dw[i]));
  }
  const int dh_span = dh1 - dh0, dw_span = dw1 - dw0;
  // padded unit width: a multiple of 32 raster columns, so that every 32-column epilogue block lies inside one row
  const int PW = (L.OW + dw_span + 31) & ~31;
  if (PW > 128 || L.OH * L.OW <= kC64N) return BDBNN_ERR_UNSUPPORTED;     // small images stay on the pixel-M kernel
  if (int64_t(L.NIMG) * L.OH * L.OW * 64 >= (int64_t(1) << 31)) return BDBNN_ERR_UNSUPPORTED;   // 32-bit element offsets
  p.OW = L.OW; p.OH = L.OH; p.NIMG = L.NIMG;
  p.PW = PW; p.dh_min = dh0; p.dw_min = dw0;
  p.SH = kC64N / PW;
  if (p.SH < 1) return BDBNN_ERR_UNSUPPORTED;
  if (p.SH > L.OH) p.SH = L.OH;
  p.PH = p.SH + dh_span;
  if (p.PH > 256) return BDBNN_ERR_UNSUPPORTED;
  p.units_per_img = (L.OH + p.SH - 1) / p.SH;
  p.n_units = L.NIMG * p.units_per_img;
  // rows a shifted 256-row view can touch: 256 + dh_span*PW + dw_span; the TMA box writes PH*PW of them
  const uint32_t view_rows = uint32_t(kC64N + dh_span * PW + dw_span);
  const uint32_t box_rows = uint32_t(p.PH * PW);
  p.patch_bytes = ((view_rows > box_rows ? view_rows : box_rows) * 128u + 1023u) & ~1023u;
  p.n_taps = L.n_taps; p.a_halves = L.a_halves; p.fmt = L.fmt;
  static const int dbg_env = [] { const char* e = getenv("BDBNN_TC_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_env;
  p.amax_bits = L.amax_bits; p.add = L.add; p.alpha = L.alpha; p.mask = L.mask; p.out = L.out;
  p.out_i16 = mode == 0 ? L.out_i16 : nullptr;
  p.bn_sums = L.bn_sums; p.bn_ymax = L.bn_ymax;
  p.st_y = L.st_y; p.st_alpha = L.st_alpha; p.st_mean = L.st_mean; p.st_invstd = L.st_invstd;
  const bool bst = mode == 1 && L.bn_sums != nullptr;
  if (bst && !(L.st_y && L.st_alpha && L.st_mean && L.st_invstd)) return BDBNN_ERR_UNSUPPORTED;
  // Measured (ResNet-18 N=256 step): with the producing unit's BatchNorm sums in the epilogue (three more loads and
  // ~6 more instructions per pixel on two SM sub-partitions) this kernel takes 0.4 ms per layer1 launch, more than
  // the plain form (0.13 ms) plus the separate reduction pass (0.06 ms): bdbnn_binconv_dgrad_tc_stats declines
  // these shapes (the caller then runs the plain dgrad here + bn_reduce) unless BDBNN_TC_C64=2.
  if (bst && enabled < 2) return BDBNN_ERR_UNSUPPORTED;
  const size_t smem = ((size_t(L.n_taps + 1) * kC64TapBytes + 1023) & ~size_t(1023)) + 2 * size_t(p.patch_bytes) + 1024;
  if (smem > 226u * 1024u) return BDBNN_ERR_UNSUPPORTED;
  CUtensorMap tmX, tmW;
  int rc = make_act_map(&tmX, L.A, L.NIMG, L.IH, L.IW, L.Kc * L.a_halves, 64, p.PW, p.PH, 1, 1, 2);
  if (rc) return rc;
  rc = make_weight_map(&tmW, L.B, L.Nout, L.b_taps * L.Kc, 64, 64, 2);
  if (rc) return rc;
  int grid = num_sms();
  if (grid > p.n_units) grid = p.n_units;
  auto launch = [&](auto kern) -> int {
    BDBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kC64Threads, smem, st>>>(tmX, tmW, p);
    return check_launch("tc_conv64_kernel");
  };
  if (mode == 0) return p.out_i16 ? launch(tc_conv64_kernel<0, false, true>) : launch(tc_conv64_kernel<0, false, false>);
  return bst ? launch(tc_conv64_kernel<1, true, false>) : launch(tc_conv64_kernel<1, false, false>);
}

// Stem conv forward (7 vertical taps of K = 32 over the packed window image) on the pixel-N kernel.
// L is the launch launch_stem_fwd builds (win = 1, Kc = 32, in_step = 2, Nout = 64).
int launch_tc_conv64_stem(const TcConvLaunch& L, cudaStream_t st) {
  static const int enabled = [] { const char* e = getenv("BDBNN_TC_C64_STEM"); return e ? atoi(e) : 1; }();
  if (!enabled || !c64_enabled() || !L.win || L.Kc != 32 || L.Nout != 64 || L.in_step != 2 || L.n_taps != 7 ||
      L.fmt < 0 || L.a_halves != 1)
    return BDBNN_ERR_UNSUPPORTED;
  for (int r = 0; r < 7; ++r)
    if (L.dh[r] != r || L.dw[r] != 0 || L.tb[r] != r) return BDBNN_ERR_UNSUPPORTED;
  if (L.OW > 128 || L.OW <= 32 || L.OH < 2) return BDBNN_ERR_UNSUPPORTED;
  if (int64_t(L.NIMG) * L.OH * L.OW * 64 >= (int64_t(1) << 31)) return BDBNN_ERR_UNSUPPORTED;
  TcConv64Params p;
  memset(&p, 0, sizeof(p));
  p.OW = L.OW; p.OH = L.OH; p.NIMG = L.NIMG;
  p.PW = 128;                                   // windows per unit row (zero-filled beyond OW)
  p.SH = kC64N / p.PW;                          // 2 output rows per unit
  p.plane_rows = p.SH + 3;                      // window rows oh0+k, k = 0 .. SH-1+3, of one parity
  p.plane_bytes = uint32_t(p.plane_rows * p.PW) * 64u;          // 40 KB: also covers every 256-row view (k <= 3)
  p.patch_bytes = 2u * p.plane_bytes;
  p.units_per_img = (L.OH + p.SH - 1) / p.SH;
  p.n_units = L.NIMG * p.units_per_img;
  p.n_taps = 7; p.a_halves = 1; p.fmt = L.fmt;
  for (int r = 0; r < 7; ++r) p.tap_b[r] = uint8_t(r);
  p.alpha = L.alpha; p.out = L.out; p.bn_sums = L.bn_sums; p.bn_ymax = L.bn_ymax;
  static const int dbg_env = [] { const char* e = getenv("BDBNN_TC_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_env;
  const size_t smem = ((size_t(7 + 1) * 64 * 64 + 1023) & ~size_t(1023)) + 2 * size_t(p.patch_bytes) + 1024;
  if (smem > 226u * 1024u) return BDBNN_ERR_UNSUPPORTED;
  CUtensorMap tmX, tmW;
  // window image: box = [1 image][plane_rows rows, every 2nd window row][128 windows][32 halves]
  int rc = make_window_map(&tmX, L.A, L.NIMG, L.IH, L.IW, 32, L.win_stride, L.win_row_stride, L.win_img_stride, p.PW,
                           p.plane_rows, 1, 2);
  if (rc) return rc;
  rc = make_weight_map(&tmW, L.B, L.Nout, L.b_taps * L.Kc, 32, 64, 2);
  if (rc) return rc;
  int grid = num_sms();
  if (grid > p.n_units) grid = p.n_units;
  auto kern = tc_conv64_kernel<0, false, false, true>;
  BDBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  kern<<<grid, kC64Threads, smem, st>>>(tmX, tmW, p);
  return check_launch("tc_conv64_kernel(stem)");
}

}  // namespace bdbnn
