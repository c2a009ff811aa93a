// tcgen05 / TMA implicit-GEMM kernels for the binary convolution (sm_100a only).
//
//   tc_conv_kernel  : D[pix, n] = sum_{tap t} sum_{k} A[pix shifted by t, k] * B[n, t*Kc + k]
//                     forward : A = sign(x)  (+-1 bf16, NHWC), B = sign(W)  [Cout][T][Cin]; epilogue * alpha[n]
//                     dgrad   : A = gy*gscale (bf16,  NHWC), B = sign(W)^T [Cin][T'][Cout]; epilogue * STE mask bit
//
// Dataflow per CTA (one 128-row output tile, one N tile):
//   warp 4 lane 0 : TMA producer. For every (tap, K-block) loads the activation box
//                   [BNI images][BH rows][BW=OW cols][KB channels] at coordinates shifted by the tap —
//                   TMA zero-fills out-of-image coordinates, which IS the conv's zero padding — and the
//                   matching [BN x KB] weight box, into a `stages`-deep ring of 128B-swizzled K-major tiles.
//   warp 5 lane 0 : issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), M=128, N=BN, K=16 per
//                   instruction; tcgen05.commit releases ring slots / signals the epilogue.
//   warps 0..3    : epilogue. tcgen05.ld the accumulator (warp w owns TMEM lanes 32w..32w+31 = tile rows),
//                   apply alpha / STE mask, store fp32 NHWC rows (each thread writes whole 64-byte runs).
// +-1 operands and fp32 accumulation make the forward integer-exact (|sum| <= 9*512 < 2^24).
#include "tc_common.cuh"

namespace bdbnn {


struct TcConvParams {
  int32_t OW, OH, NIMG;      // output pixel grid (one GEMM row per output pixel)
  int32_t BW, BH, BNI;       // tile = BNI images x BH rows x BW(=OW) cols  (<= 128 pixels)
  int32_t tiles_h;           // tiles per image group along h
  int32_t Kc, KB, n_kb;      // contraction channels, K-block elements (16/32/64), Kc/KB
  int32_t a_halves;          // 1: A has Kc channels; 2: A = [hi | lo] bf16 split, 2*Kc channels, B reused
  int32_t in_step;           // input coordinate = out coordinate * in_step + tap offset (2 for stride-2 fwd)
  int32_t n_taps;            // taps actually visited (subset for the stride-2 dgrad phases)
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];   // input offset of tap i
  uint8_t tap_b[kMaxTaps];   // K-block row of B for tap i (B column = tap_b * Kc + k)
  int32_t out_step, out_off_h, out_off_w, OHf, OWf;  // out pixel = (oh*out_step+off_h, ow*out_step+off_w) in OHf x OWf
  int32_t Nout, BN;          // GEMM N total / per CTA
  // halo mode: ONE activation patch [PH][PW][64ch] per K block serves all taps; tap (dh,dw) is the same
  // swizzled patch read from row offset (dh-dh_min)*PW + (dw-dw_min) (descriptor start + 128*shift)
  int32_t halo, PW, PH, dh_min, dw_min, patch_bytes;
  int32_t stages;
  int32_t row_bytes;         // KB * 2 = swizzle span (32/64/128)
  int32_t fmt;               // operand format
  const uint32_t* amax_bits; // FP16S: post-scale 2^-e (NULL otherwise)
  const float* add;          // dgrad: optional additive tensor
  const float* alpha;        // MODE 0: per-output-channel scale
  const uint32_t* mask;      // MODE 1: STE mask words [pix][Nout/32 (ceil)]
  float* out;                // [pix][Nout] fp32
};

template <int MODE>
__global__ void __launch_bounds__(kTcThreads)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ __align__(8) uint64_t pfull_bar[2], pempty_bar[2];   // halo patch ring
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t tiles_base = smem_base + (p.halo ? 2u * uint32_t(p.patch_bytes) : 0u);
  const uint32_t a_bytes = p.halo ? 0u : kTileM * p.row_bytes;   // ring slot: [A tile] then B tile
  const uint32_t b_bytes = p.BN * p.row_bytes;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  const uint32_t tmem_cols = p.BN < 32 ? 32u : uint32_t(p.BN);   // power of two >= 32

  const int tile_n = blockIdx.x / p.tiles_h, tile_h = blockIdx.x - tile_n * p.tiles_h;
  const int n0 = tile_n * p.BNI, h0 = tile_h * p.BH;
  const int nn0 = blockIdx.y * p.BN;
  const int kb_total = p.n_kb * p.a_halves;
  const int n_iters = p.n_taps * kb_total;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&pfull_bar[s]), 1);
      mbar_init(smem_u32(&pempty_bar[s]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 5) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (p.halo) {
    if (warp == 4) {
      if (lane == 0) {
        int it = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pa = kb & 1;
          mbar_wait(smem_u32(&pempty_bar[pa]), (uint32_t(kb >> 1) & 1u) ^ 1u);
          const uint32_t pb = smem_u32(&pfull_bar[pa]);
          mbar_expect_tx(pb, uint32_t(p.PW * p.PH) * 128u);
          tma_load_4d(smem_base + pa * p.patch_bytes, &tmA, pb, kb * p.KB, p.dw_min, h0 + p.dh_min, n0);
          const int kbb = kb >= p.n_kb ? kb - p.n_kb : kb;
          for (int ti = 0; ti < p.n_taps; ++ti, ++it) {
            const int stage = it % p.stages;
            const uint32_t phase = uint32_t(it / p.stages) & 1u;
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, uint32_t(p.BN) * 128u);
            tma_load_2d(tiles_base + stage * stage_bytes, &tmB, fb, p.tap_b[ti] * p.Kc + kbb * p.KB, nn0);
          }
        }
      }
    } else if (warp == 5) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN), uint32_t(p.fmt));
        int it = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int pa = kb & 1;
          mbar_wait(smem_u32(&pfull_bar[pa]), uint32_t(kb >> 1) & 1u);
          tc_fence_after();
          const uint32_t patch = smem_base + pa * p.patch_bytes;
          for (int ti = 0; ti < p.n_taps; ++ti, ++it) {
            const int stage = it % p.stages;
            const uint32_t phase = uint32_t(it / p.stages) & 1u;
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t shift = uint32_t((p.tap_dh[ti] - p.dh_min) * p.PW + (p.tap_dw[ti] - p.dw_min));
            const uint32_t b_src = tiles_base + stage * stage_bytes;
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_kmajor_desc(patch + shift * 128u + k * 32, 128);
              const uint64_t bd = make_kmajor_desc(b_src + k * 32, 128);
              umma_bf16(tmem_d, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          umma_commit(smem_u32(&pempty_bar[pa]));
        }
        umma_commit(smem_u32(&accum_bar));
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t tx = uint32_t(p.BNI * p.BH * p.BW + p.BN) * uint32_t(p.row_bytes);
      int it = 0;
      for (int ti = 0; ti < p.n_taps; ++ti) {
        const int dh = p.tap_dh[ti], dw = p.tap_dw[ti], tb = p.tap_b[ti];
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const int stage = it % p.stages;
          const uint32_t phase = uint32_t(it / p.stages) & 1u;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, tx);
          const uint32_t a_dst = tiles_base + stage * stage_bytes;
          const int kbb = kb >= p.n_kb ? kb - p.n_kb : kb;       // hi and lo halves share B
          tma_load_4d(a_dst, &tmA, fb, kb * p.KB, dw, h0 * p.in_step + dh, n0);
          tma_load_2d(a_dst + a_bytes, &tmB, fb, tb * p.Kc + kbb * p.KB, nn0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN), uint32_t(p.fmt));
      const int k_steps = p.KB / 16;                  // UMMA K = 16 bf16 = 32 bytes
      for (int it = 0; it < n_iters; ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = uint32_t(it / p.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t a_src = tiles_base + stage * stage_bytes;
        for (int k = 0; k < k_steps; ++k) {
          const uint64_t ad = make_kmajor_desc(a_src + k * 32, p.row_bytes);
          const uint64_t bd = make_kmajor_desc(a_src + a_bytes + k * 32, p.row_bytes);
          umma_bf16(tmem_d, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[stage]));     // slot reusable once these MMAs retire
      }
      umma_commit(smem_u32(&accum_bar));
    }
  }
  if (warp < 4) {
    // ---- epilogue: warp w <-> TMEM lanes [32w, 32w+32) <-> tile rows ----
    const int m = warp * 32 + lane;
    const int rw = p.halo ? p.PW : p.BW;               // halo tiles live in padded-width pixel space
    const int wi = m % rw;
    const int q = m / rw;
    const int hi = q % p.BH, ni = (p.halo && wi >= p.BW) ? p.BNI : q / p.BH;
    const int oh = (h0 + hi) * p.out_step + p.out_off_h, ow = wi * p.out_step + p.out_off_w;
    const bool valid = (ni < p.BNI) && (h0 + hi < p.OH) && (n0 + ni < p.NIMG) && oh < p.OHf && ow < p.OWf;
    const int64_t pix = (int64_t(n0 + ni) * p.OHf + oh) * p.OWf + ow;
    float* orow = p.out + pix * p.Nout + nn0;
    const int mask_words = (p.Nout + 31) >> 5;

    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const float post = p.amax_bits ? amax_pow2_scale(__ldg(p.amax_bits), true) : 1.0f;
    const uint32_t lane_base = tmem_d + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(lane_base + uint32_t(c0), v);
      tmem_ld_wait();
      if (valid) {
        float f[16];
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * __ldg(p.alpha + nn0 + c0 + j);
        } else {
          const int col = nn0 + c0;
          const uint32_t word = __ldg(p.mask + pix * mask_words + (col >> 5)) >> (col & 31);
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = ((word >> j) & 1u) ? __uint_as_float(v[j]) * post : 0.0f;
          if (p.add != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += p.add[pix * p.Nout + nn0 + c0 + j];   // plain load: add may alias out
          }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

static bool chan_ok(int c) { return c == 16 || c == 32 || c == 64 || (c >= 128 && c % 128 == 0); }
static int pick_bn(int n) { return n >= 128 ? 128 : n; }   // n in {16,32,64} or multiple of 128

static bool tc_shape_ok(const bdbnn_conv_shape* s) {
  if (!s || (s->stride != 1 && s->stride != 2)) return false;
  if (!chan_ok(s->Cin) || !chan_ok(s->Cout)) return false;   // K blocks of 16/32/64, N tiles <= 128
  if (s->Wo > 128 || s->W > 128 * s->stride || s->W < 1) return false;
  if (s->kh != s->kw || s->kh > 7) return false;
  if (s->pad > s->kh - 1) return false;
  return true;
}

template <int MODE>
static int launch_tc_conv(const TcConvLaunch& L, cudaStream_t st) {
  // Persistent multi-accumulator kernel first (tc_conv2.cu); BDBNN_TC_V2=0 forces the simple kernel.
  static const int v2_env = [] { const char* e = getenv("BDBNN_TC_V2"); return e ? atoi(e) : 1; }();
  if (v2_env) {
    const int rc64 = launch_tc_conv64(L, MODE, st);        // 64-channel layers: pixels as the N = 256 dimension
    if (rc64 != BDBNN_ERR_UNSUPPORTED) return rc64;
    const int rc2 = launch_tc_conv2(L, MODE, st);
    if (rc2 != BDBNN_ERR_UNSUPPORTED) return rc2;
  }
  if (L.out_i16) return BDBNN_ERR_UNSUPPORTED;      // int16 output exists in the persistent kernel only
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  p.OW = L.OW; p.OH = L.OH; p.NIMG = L.NIMG;
  p.BW = L.OW;
  if (L.OH * L.OW <= kTileM) {
    p.BH = L.OH;
    p.BNI = kTileM / (L.OH * L.OW);
    if (p.BNI > L.NIMG) p.BNI = L.NIMG;
    if (p.BNI > 256) p.BNI = 256;
  } else {
    p.BH = kTileM / L.OW;
    p.BNI = 1;
  }
  p.tiles_h = (L.OH + p.BH - 1) / p.BH;
  const int tiles_n = (L.NIMG + p.BNI - 1) / p.BNI;
  p.Kc = L.Kc;
  p.KB = L.Kc >= 64 ? 64 : L.Kc;
  p.n_kb = L.Kc / p.KB;
  p.a_halves = L.a_halves;
  p.in_step = L.in_step;
  p.row_bytes = p.KB * 2;
  p.n_taps = L.n_taps;
  memcpy(p.tap_dh, L.dh, sizeof(p.tap_dh));
  memcpy(p.tap_dw, L.dw, sizeof(p.tap_dw));
  memcpy(p.tap_b, L.tb, sizeof(p.tap_b));
  p.out_step = L.out_step; p.out_off_h = L.off_h; p.out_off_w = L.off_w; p.OHf = L.OHf; p.OWf = L.OWf;
  p.Nout = L.Nout;
  p.BN = pick_bn(L.Nout);
  p.alpha = L.alpha; p.mask = L.mask; p.out = L.out;
  p.fmt = L.fmt; p.amax_bits = L.amax_bits; p.add = L.add;
  // Halo mode (stride-1 launches with >128-pixel images and 64-channel K blocks): see TcConvParams.
  static const int halo_env = [] { const char* e = getenv("BDBNN_TC_HALO"); return e ? atoi(e) : 1; }();
  if (halo_env > 0 && !L.win && L.in_step == 1 && p.KB == 64 && L.OH * L.OW > kTileM && p.n_taps > 0) {
    int dh0 = 127, dh1 = -127, dw0 = 127, dw1 = -127;
    for (int i = 0; i < p.n_taps; ++i) {
      dh0 = min(dh0, int(p.tap_dh[i])); dh1 = max(dh1, int(p.tap_dh[i]));
      dw0 = min(dw0, int(p.tap_dw[i])); dw1 = max(dw1, int(p.tap_dw[i]));
    }
    const int PW = L.OW + (dw1 - dw0);
    if (PW <= kTileM) {
      p.halo = 1;
      p.PW = PW; p.dh_min = dh0; p.dw_min = dw0;
      p.BNI = 1;
      p.BH = kTileM / PW;
      p.PH = p.BH + (dh1 - dh0);
      p.tiles_h = (L.OH + p.BH - 1) / p.BH;
      const int rows = kTileM + (dh1 - dh0) * PW + (dw1 - dw0);
      p.patch_bytes = int((uint32_t(rows) * 128u + 1023u) & ~1023u);
    }
  }
  const int tiles_n_eff = (L.NIMG + p.BNI - 1) / p.BNI;
  const uint32_t stage_bytes = p.halo ? ((uint32_t(p.BN) * 128u + 1023u) & ~1023u)
                                      : ((uint32_t(kTileM + p.BN) * p.row_bytes + 1023u) & ~1023u);
  const int n_iters = p.n_taps * p.n_kb * p.a_halves;
  const uint32_t ring_budget = 96u * 1024u - (p.halo ? 2u * uint32_t(p.patch_bytes) : 0u);
  int stages = int(ring_budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > n_iters) stages = n_iters;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = size_t(stages) * stage_bytes + (p.halo ? 2u * size_t(p.patch_bytes) : 0u) + 1024;

  CUtensorMap tmA, tmB;
  int rc;
  if (L.win)   // stem: overlapping windows, row step in_step, window step 1
    rc = make_window_map(&tmA, L.A, L.NIMG, L.IH, L.IW, p.KB, L.win_stride, L.win_row_stride, L.win_img_stride,
                         p.BW, p.BH, p.BNI, L.in_step);
  else
    rc = p.halo ? make_act_map(&tmA, L.A, L.NIMG, L.IH, L.IW, L.Kc * L.a_halves, 64, p.PW, p.PH, 1, 1)
                : make_act_map(&tmA, L.A, L.NIMG, L.IH, L.IW, L.Kc * L.a_halves, p.KB, p.BW, p.BH, p.BNI,
                               L.in_step);
  if (rc) return rc;
  rc = make_weight_map(&tmB, L.B, L.Nout, L.b_taps * L.Kc, p.KB, p.BN);
  if (rc) return rc;
  auto kern = tc_conv_kernel<MODE>;
  BDBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dim3 grid(unsigned(p.tiles_h * tiles_n_eff), unsigned(L.Nout / p.BN));
  kern<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
  rc = check_launch("tc_conv_kernel");
  if (rc) return rc;
  // this kernel has no statistics epilogue: one extra pass over the result when they were asked for
  if (MODE == 0 && L.bn_sums != nullptr)
    return bn_stats_launch(L.out, int64_t(L.NIMG) * L.OHf * L.OWf, L.Nout, L.bn_sums, L.bn_ymax, st);
  return BDBNN_OK;
}

// Stem conv forward (7x7 / stride 2 / pad 3 on 3 channels) as a 7-tap implicit GEMM over the packed
// window view: tap r reads window row 2*oh + r, K = 32 values (8 pixels x 4 halves) per tap.
int launch_stem_fwd(const uint16_t* xw, const uint16_t* wf, const float* alpha, float* y, const StemGeom& g,
                    double* bn_sums, uint32_t* bn_ymax, cudaStream_t st) {
  TcConvLaunch L;
  memset(&L, 0, sizeof(L));
  L.A = xw; L.IH = g.HP; L.IW = g.Wo; L.Kc = kStemWin; L.a_halves = 1; L.in_step = 2;
  L.win = 1; L.win_stride = g.win_stride; L.win_row_stride = g.row_stride; L.win_img_stride = g.img_stride;
  L.B = wf; L.b_taps = kStemTaps; L.Nout = kStemCout;
  L.NIMG = g.N; L.OH = g.Ho; L.OW = g.Wo;
  for (int r = 0; r < kStemTaps; ++r) { L.dh[r] = int8_t(r); L.dw[r] = 0; L.tb[r] = uint8_t(r); }
  L.n_taps = kStemTaps;
  L.out_step = 1; L.OHf = g.Ho; L.OWf = g.Wo;
  L.alpha = alpha; L.out = y; L.fmt = BDBNN_FMT_FP16;
  L.bn_sums = bn_sums; L.bn_ymax = bn_ymax;
  int rc = bn_stats_zero(bn_sums, bn_ymax, kStemCout, st);
  if (rc) return rc;
  rc = launch_tc_conv64_stem(L, st);            // pixel-N kernel (tc_conv64.cu); falls back to the pixel-M kernels
  if (rc != BDBNN_ERR_UNSUPPORTED) return rc;
  return launch_tc_conv<0>(L, st);
}

}  // namespace bdbnn

using namespace bdbnn;

extern "C" int bdbnn_debug_trace(long long* device_buf) {
  set_tc_trace(device_buf);
  return BDBNN_OK;
}

// Host-only: the persistent kernel's tiling for the forward (mode 0: 16-bit operands, 2: fp8) or the first
// dgrad phase (mode 1) of a shape.  out[12] = {planned, halo, TS, NB, BN, n_tiles, supers, stages, dynamic smem,
// grid, stage bytes, patch bytes}; planned = 0 when the shape goes to the one-tile kernel.  No CUDA call.
extern "C" int bdbnn_debug_conv_plan(const bdbnn_conv_shape* s, int32_t mode, int32_t grad_halves, int32_t* out,
                                     int32_t n_out) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(out && n_out >= 12 && mode >= 0 && mode <= 2 && (grad_halves == 1 || grad_halves == 2),
                "debug_conv_plan: bad arguments");
  memset(out, 0, 12 * sizeof(int32_t));
  if (!tc_shape_ok(s)) return BDBNN_OK;
  TcConvLaunch L;
  memset(&L, 0, sizeof(L));
  const int T = s->kh * s->kw;
  if (mode == 1) {
    const int sd = s->stride;            // phase (0,0) of the stride decomposition
    for (int r = 0; r < s->kh; ++r) {
      if ((s->pad - r) % sd != 0) continue;
      for (int q = 0; q < s->kw; ++q) {
        if ((s->pad - q) % sd != 0) continue;
        L.dh[L.n_taps] = int8_t((s->pad - r) / sd); L.dw[L.n_taps] = int8_t((s->pad - q) / sd);
        L.tb[L.n_taps] = uint8_t(T - 1 - (r * s->kw + q));
        ++L.n_taps;
      }
    }
    L.OH = (s->H + sd - 1) / sd; L.OW = (s->W + sd - 1) / sd;
    L.IH = s->Ho; L.IW = s->Wo; L.Kc = s->Cout; L.a_halves = grad_halves; L.in_step = 1;
    L.b_taps = T; L.Nout = s->Cin; L.NIMG = s->N;
    L.out_step = sd; L.OHf = s->H; L.OWf = s->W;
    L.fmt = grad_halves == 1 ? BDBNN_FMT_FP16 : BDBNN_FMT_BF16;
  } else {
    L.IH = s->H; L.IW = s->W; L.Kc = s->Cin; L.a_halves = 1; L.in_step = s->stride;
    L.b_taps = T; L.Nout = s->Cout; L.NIMG = s->N; L.OH = s->Ho; L.OW = s->Wo;
    for (int r = 0; r < s->kh; ++r)
      for (int q = 0; q < s->kw; ++q) {
        const int t = r * s->kw + q;
        L.dh[t] = int8_t(r - s->pad); L.dw[t] = int8_t(q - s->pad); L.tb[t] = uint8_t(t);
      }
    L.n_taps = T;
    L.out_step = 1; L.OHf = s->Ho; L.OWf = s->Wo;
    L.fmt = mode == 2 ? -1 : BDBNN_FMT_FP16;
    if (mode == 2 && s->Cin % 128 != 0) return BDBNN_OK;
  }
  if (L.n_taps == 0) return BDBNN_OK;
  set_conv_plan_sink(out);
  rc = launch_tc_conv2(L, mode == 1 ? 1 : 0, nullptr);
  set_conv_plan_sink(nullptr);
  if (rc == BDBNN_ERR_UNSUPPORTED) { memset(out, 0, 12 * sizeof(int32_t)); return BDBNN_OK; }
  return rc;
}

extern "C" int bdbnn_tc_supported(const bdbnn_conv_shape* s) {
  if (!tc_shape_ok(s)) return 0;
  // fp8 forward only with 128-byte K rows (Cin % 128 == 0): with 64-channel (64-byte, SWIZZLE_64B) rows
  // each K=32 MMA took ~240 clk on B200 (vs ~105 clk for the 16-bit K=16 MMA), i.e. no gain for layer1
  const bool v2 = s->Cin % 128 == 0 && (s->Cout == 64 || s->Cout % 128 == 0);
  // int16 forward output: the persistent kernel must take the 16-bit-operand forward of this shape, and the
  // integer result must fit (|y_int| <= kh*kw*Cin)
  bool i16 = s->kh * s->kw * s->Cin <= 32767 && s->Cout <= 512;
  if (i16) {
    int32_t plan[12] = {0};
    i16 = bdbnn_debug_conv_plan(s, 0, 1, plan, 12) == BDBNN_OK && plan[0] == 1;
    if (i16 && v2) i16 = bdbnn_debug_conv_plan(s, 2, 1, plan, 12) == BDBNN_OK && plan[0] == 1;   // fp8 forward
  }
  return BDBNN_TC_FWD | BDBNN_TC_DGRAD | (wgrad_tc_ok(s) ? BDBNN_TC_WGRAD : 0) | (v2 ? BDBNN_TC_FWD8 : 0) |
         (i16 ? BDBNN_TC_FWD_I16 : 0);
}

// Forward with the result stored as the exact integer accumulator (int16) — see include/bdbnn.h.
extern "C" int bdbnn_binconv_fwd_tc_i16(const void* xb, const void* wf, int32_t fmt, const float* alpha,
                                        int16_t* y_int, const bdbnn_conv_shape* s, double* bn_sums,
                                        uint32_t* bn_ymax, void* stream) {
  BDBNN_REQUIRE(fmt == BDBNN_FMT_FP16 || fmt == BDBNN_FMT_BF16 || fmt == -1, "binconv_fwd_tc_i16: bad operand format");
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(xb && wf && alpha && y_int, "binconv_fwd_tc_i16: NULL pointer");
  if (!tc_shape_ok(s) || (fmt == -1 && s->Cin % 128 != 0) || s->kh * s->kw * s->Cin > 32767) {
    set_error("binconv_fwd_tc_i16: shape not supported");
    return BDBNN_ERR_UNSUPPORTED;
  }
  TcConvLaunch L;
  memset(&L, 0, sizeof(L));
  L.A = static_cast<const uint16_t*>(xb); L.IH = s->H; L.IW = s->W; L.Kc = s->Cin; L.a_halves = 1; L.in_step = s->stride;
  L.B = static_cast<const uint16_t*>(wf); L.b_taps = s->kh * s->kw; L.Nout = s->Cout;
  L.NIMG = s->N; L.OH = s->Ho; L.OW = s->Wo;
  for (int r = 0; r < s->kh; ++r)
    for (int q = 0; q < s->kw; ++q) {
      const int t = r * s->kw + q;
      L.dh[t] = int8_t(r - s->pad); L.dw[t] = int8_t(q - s->pad); L.tb[t] = uint8_t(t);
    }
  L.n_taps = s->kh * s->kw;
  L.out_step = 1; L.OHf = s->Ho; L.OWf = s->Wo;
  L.alpha = alpha; L.out = nullptr; L.out_i16 = y_int; L.fmt = fmt;
  BDBNN_REQUIRE((bn_sums == nullptr) == (bn_ymax == nullptr), "binconv_fwd_tc_i16: bn_sums and bn_ymax go together");
  BDBNN_REQUIRE(bn_sums == nullptr || s->Cout <= 512, "binconv_fwd_tc_i16: statistics need Cout <= 512");
  L.bn_sums = bn_sums; L.bn_ymax = bn_ymax;
  rc = bn_stats_zero(bn_sums, bn_ymax, s->Cout, cudaStream_t(stream));
  if (rc) return rc;
  rc = launch_tc_conv64(L, 0, cudaStream_t(stream));
  if (rc == BDBNN_ERR_UNSUPPORTED) rc = launch_tc_conv2(L, 0, cudaStream_t(stream));
  if (rc == BDBNN_ERR_UNSUPPORTED) set_error("binconv_fwd_tc_i16: geometry not supported by the persistent kernel");
  return rc;
}

extern "C" int bdbnn_binconv_fwd_tc(const uint16_t* xb_bf16, const uint16_t* wf_bf16, int32_t fmt,
                                    const float* alpha, float* y, const bdbnn_conv_shape* s, double* bn_sums,
                                    uint32_t* bn_ymax, void* stream) {
  BDBNN_REQUIRE(fmt == BDBNN_FMT_FP16 || fmt == BDBNN_FMT_BF16, "binconv_fwd_tc: bad operand format");
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(xb_bf16 && wf_bf16 && alpha && y, "binconv_fwd_tc: NULL pointer");
  if (!tc_shape_ok(s)) { set_error("binconv_fwd_tc: shape not supported by the tcgen05 path"); return BDBNN_ERR_UNSUPPORTED; }
  TcConvLaunch L;
  memset(&L, 0, sizeof(L));
  L.A = xb_bf16; L.IH = s->H; L.IW = s->W; L.Kc = s->Cin; L.a_halves = 1; L.in_step = s->stride;
  L.B = wf_bf16; L.b_taps = s->kh * s->kw; L.Nout = s->Cout;
  L.NIMG = s->N; L.OH = s->Ho; L.OW = s->Wo;
  for (int r = 0; r < s->kh; ++r)
    for (int q = 0; q < s->kw; ++q) {
      const int t = r * s->kw + q;
      L.dh[t] = int8_t(r - s->pad); L.dw[t] = int8_t(q - s->pad); L.tb[t] = uint8_t(t);
    }
  L.n_taps = s->kh * s->kw;
  L.out_step = 1; L.OHf = s->Ho; L.OWf = s->Wo;
  L.alpha = alpha; L.out = y; L.fmt = fmt;
  BDBNN_REQUIRE((bn_sums == nullptr) == (bn_ymax == nullptr), "binconv_fwd_tc: bn_sums and bn_ymax go together");
  L.bn_sums = bn_sums; L.bn_ymax = bn_ymax;
  rc = bn_stats_zero(bn_sums, bn_ymax, s->Cout, cudaStream_t(stream));
  if (rc) return rc;
  return launch_tc_conv<0>(L, cudaStream_t(stream));
}

extern "C" int bdbnn_binconv_fwd_tc8(const uint8_t* xb_fp8, const uint8_t* wf_fp8, const float* alpha, float* y,
                                     const bdbnn_conv_shape* s, double* bn_sums, uint32_t* bn_ymax,
                                     void* stream) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(xb_fp8 && wf_fp8 && alpha && y, "binconv_fwd_tc8: NULL pointer");
  if (!tc_shape_ok(s) || s->Cin % 128 != 0) {
    set_error("binconv_fwd_tc8: shape not supported by the fp8 tcgen05 path");
    return BDBNN_ERR_UNSUPPORTED;
  }
  TcConvLaunch L;
  memset(&L, 0, sizeof(L));
  L.A = reinterpret_cast<const uint16_t*>(xb_fp8); L.IH = s->H; L.IW = s->W; L.Kc = s->Cin; L.a_halves = 1;
  L.in_step = s->stride;
  L.B = reinterpret_cast<const uint16_t*>(wf_fp8); L.b_taps = s->kh * s->kw; L.Nout = s->Cout;
  L.NIMG = s->N; L.OH = s->Ho; L.OW = s->Wo;
  for (int r = 0; r < s->kh; ++r)
    for (int q = 0; q < s->kw; ++q) {
      const int t = r * s->kw + q;
      L.dh[t] = int8_t(r - s->pad); L.dw[t] = int8_t(q - s->pad); L.tb[t] = uint8_t(t);
    }
  L.n_taps = s->kh * s->kw;
  L.out_step = 1; L.OHf = s->Ho; L.OWf = s->Wo;
  L.alpha = alpha; L.out = y; L.fmt = -1;
  BDBNN_REQUIRE((bn_sums == nullptr) == (bn_ymax == nullptr), "binconv_fwd_tc8: bn_sums and bn_ymax go together");
  BDBNN_REQUIRE(bn_sums == nullptr || s->Cout <= 512, "binconv_fwd_tc8: statistics need Cout <= 512");
  L.bn_sums = bn_sums; L.bn_ymax = bn_ymax;
  rc = bn_stats_zero(bn_sums, bn_ymax, s->Cout, cudaStream_t(stream));
  if (rc) return rc;
  rc = launch_tc_conv2(L, 0, cudaStream_t(stream));
  if (rc == BDBNN_ERR_UNSUPPORTED) set_error("binconv_fwd_tc8: geometry not supported by the persistent kernel");
  return rc;
}

struct DgradStats {      // backward statistics of the BatchNorm unit that produced this conv's input (see tc_common.cuh)
  const int16_t* y_int; const float* alpha; const float* mean; const float* invstd;
  double* sums; uint32_t* gmax;
};

static int dgrad_tc_impl(const uint16_t* gys_bf16, int32_t grad_mode, const uint32_t* amax_bits,
                         const uint16_t* wt_bf16, const uint32_t* mask_bits, const float* add, float* gx,
                         const bdbnn_conv_shape* s, const DgradStats* stats, void* stream) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(gys_bf16 && wt_bf16 && mask_bits && gx, "binconv_dgrad_tc: NULL pointer");
  BDBNN_REQUIRE(grad_mode >= BDBNN_GRAD_BF16 && grad_mode <= BDBNN_GRAD_FP16S, "binconv_dgrad_tc: bad grad_mode");
  BDBNN_REQUIRE(grad_mode != BDBNN_GRAD_FP16S || amax_bits, "binconv_dgrad_tc: FP16S needs amax_bits");
  const int grad_halves = grad_mode == BDBNN_GRAD_BF16X2 ? 2 : 1;
  if (!tc_shape_ok(s)) { set_error("binconv_dgrad_tc: shape not supported by the tcgen05 path"); return BDBNN_ERR_UNSUPPORTED; }
  cudaStream_t st = cudaStream_t(stream);
  const int T = s->kh * s->kw, sd = s->stride;
  // gx[h,w,c] = sum over taps (r,q) with (h+p-r) % sd == 0 (same in w) of
  //             gys[(h+p-r)/sd, (w+p-q)/sd, o] * wt[c][T-1-t][o]; one launch per (h%sd, w%sd) phase.
  TcConvLaunch Ls[4];
  int n_launch = 0;
  bool empty_phase = false;
  for (int a = 0; a < sd; ++a)
    for (int b = 0; b < sd; ++b) {
      TcConvLaunch& L = Ls[n_launch];
      memset(&L, 0, sizeof(L));
      for (int r = 0; r < s->kh; ++r) {
        const int nh = a + s->pad - r;
        if (nh % sd != 0) continue;
        for (int q = 0; q < s->kw; ++q) {
          const int nw = b + s->pad - q;
          if (nw % sd != 0) continue;
          L.dh[L.n_taps] = int8_t(nh / sd); L.dw[L.n_taps] = int8_t(nw / sd);
          L.tb[L.n_taps] = uint8_t(T - 1 - (r * s->kw + q));
          ++L.n_taps;
        }
      }
      L.OH = (s->H - a + sd - 1) / sd; L.OW = (s->W - b + sd - 1) / sd;
      if (L.OH <= 0 || L.OW <= 0) continue;
      if (L.n_taps == 0) { empty_phase = true; continue; }
      L.A = gys_bf16; L.IH = s->Ho; L.IW = s->Wo; L.Kc = s->Cout; L.a_halves = grad_halves; L.in_step = 1;
      L.B = wt_bf16; L.b_taps = T; L.Nout = s->Cin;
      L.NIMG = s->N;
      L.out_step = sd; L.off_h = a; L.off_w = b; L.OHf = s->H; L.OWf = s->W;
      L.mask = mask_bits; L.out = gx;
      L.fmt = grad_mode == BDBNN_GRAD_FP16S ? BDBNN_FMT_FP16 : BDBNN_FMT_BF16;
      L.amax_bits = grad_mode == BDBNN_GRAD_FP16S ? amax_bits : nullptr;
      L.add = add;
      if (stats) {
        L.bn_sums = stats->sums; L.bn_ymax = stats->gmax;
        L.st_y = stats->y_int; L.st_alpha = stats->alpha; L.st_mean = stats->mean; L.st_invstd = stats->invstd;
      }
      ++n_launch;
    }
  if (empty_phase) {
    const size_t bytes = size_t(s->N) * s->H * s->W * s->Cin * sizeof(float);
    // add == gx: accumulate in place (positions no phase writes keep their value, nothing to initialise)
    if (add != nullptr && add != gx) BDBNN_CUDA(cudaMemcpyAsync(gx, add, bytes, cudaMemcpyDeviceToDevice, st));
    else if (add == nullptr) BDBNN_CUDA(cudaMemsetAsync(gx, 0, bytes, st));
  }
  if (stats) {
    // one phase (stride 1), persistent kernel only: its epilogue owns the statistics
    if (n_launch != 1 || empty_phase) { set_error("binconv_dgrad_tc_stats: needs stride 1"); return BDBNN_ERR_UNSUPPORTED; }
    static const int c64_env = [] { const char* e = getenv("BDBNN_TC_C64"); return e ? atoi(e) : 1; }();
    if (c64_env < 2 && tc_conv64_eligible(Ls[0])) {
      // 64-channel layers: the plain pixel-N dgrad + the separate reduction pass is faster (see tc_conv64.cu)
      set_error("binconv_dgrad_tc_stats: declined for this shape (pixel-N kernel without statistics is faster)");
      return BDBNN_ERR_UNSUPPORTED;
    }
    rc = bn_stats_zero(stats->sums, stats->gmax, s->Cin, st);
    if (rc) return rc;
    rc = launch_tc_conv64(Ls[0], 1, st);
    if (rc == BDBNN_ERR_UNSUPPORTED) rc = launch_tc_conv2(Ls[0], 1, st);
    if (rc == BDBNN_ERR_UNSUPPORTED) set_error("binconv_dgrad_tc_stats: geometry not supported by the persistent kernel");
    return rc;
  }
  for (int i = 0; i < n_launch; ++i) {
    rc = launch_tc_conv<1>(Ls[i], st);
    if (rc) return rc;
  }
  return BDBNN_OK;
}

extern "C" int bdbnn_binconv_dgrad_tc(const uint16_t* gys_bf16, int32_t grad_mode, const uint32_t* amax_bits,
                                      const uint16_t* wt_bf16, const uint32_t* mask_bits, const float* add,
                                      float* gx, const bdbnn_conv_shape* s, void* stream) {
  return dgrad_tc_impl(gys_bf16, grad_mode, amax_bits, wt_bf16, mask_bits, add, gx, s, nullptr, stream);
}

extern "C" int bdbnn_binconv_dgrad_tc_stats(const uint16_t* gys_bf16, int32_t grad_mode, const uint32_t* amax_bits,
                                            const uint16_t* wt_bf16, const uint32_t* mask_bits, const float* add,
                                            float* gx, const bdbnn_conv_shape* s, const int16_t* prod_y_int,
                                            const float* prod_alpha, const float* prod_mean,
                                            const float* prod_invstd, double* prod_sums, uint32_t* prod_gmax,
                                            void* stream) {
  BDBNN_REQUIRE(prod_y_int && prod_alpha && prod_mean && prod_invstd && prod_sums && prod_gmax,
                "binconv_dgrad_tc_stats: NULL statistics pointer");
  BDBNN_REQUIRE(s && s->stride == 1 && s->Cin <= 512, "binconv_dgrad_tc_stats: needs stride 1 and Cin <= 512");
  const DgradStats stats = {prod_y_int, prod_alpha, prod_mean, prod_invstd, prod_sums, prod_gmax};
  return dgrad_tc_impl(gys_bf16, grad_mode, amax_bits, wt_bf16, mask_bits, add, gx, s, &stats, stream);
}

