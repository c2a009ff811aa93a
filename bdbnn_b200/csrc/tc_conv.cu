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
        const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN));
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
      const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN));
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
          for (int j = 0; j < 16; ++j) f[j] = ((word >> j) & 1u) ? __uint_as_float(v[j]) : 0.0f;
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
    const int rc2 = launch_tc_conv2(L, MODE, st);
    if (rc2 != BDBNN_ERR_UNSUPPORTED) return rc2;
  }
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
  // Halo mode (stride-1 launches with >128-pixel images and 64-channel K blocks): see TcConvParams.
  static const int halo_env = [] { const char* e = getenv("BDBNN_TC_HALO"); return e ? atoi(e) : 1; }();
  if (halo_env > 0 && L.in_step == 1 && p.KB == 64 && L.OH * L.OW > kTileM && p.n_taps > 0) {
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
  int rc = p.halo ? make_act_map(&tmA, L.A, L.NIMG, L.IH, L.IW, L.Kc * L.a_halves, 64, p.PW, p.PH, 1, 1)
                  : make_act_map(&tmA, L.A, L.NIMG, L.IH, L.IW, L.Kc * L.a_halves, p.KB, p.BW, p.BH, p.BNI,
                                 L.in_step);
  if (rc) return rc;
  rc = make_weight_map(&tmB, L.B, L.Nout, L.b_taps * L.Kc, p.KB, p.BN);
  if (rc) return rc;
  auto kern = tc_conv_kernel<MODE>;
  BDBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dim3 grid(unsigned(p.tiles_h * tiles_n_eff), unsigned(L.Nout / p.BN));
  kern<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
  return check_launch("tc_conv_kernel");
}

// ===================================================================================================
// Weight gradient on tensor cores.
//
//   D[(t,c), o] = sum_pix  xb[pix shifted by tap t, c] * gys[pix, o]          (K = output pixels)
//
// Both operands are "MN-major" for this GEMM: for a fixed pixel the M index (c) / N index (o) is the
// contiguous one in NHWC memory, so a TMA box [pixels][64 channels] lands in shared memory exactly in
// the canonical MN-major SWIZZLE_128B layout (8 pixel rows x 128 B per swizzle atom; SBO = 1024 B
// between 8-row K groups; LBO = distance between 64-channel chunks = one box).  An M tile of 128 rows
// is two 64-channel "units" (tap, channel-chunk) — for Cin = 64 that is two taps side by side.
// Each CTA owns up to G accumulators (G*BN <= 512 TMEM columns) that share the gys tile, walks a
// contiguous range of pixel boxes (split-K), and finally reduces its partial tile into the fp32
// workspace with red.global.add; a small finalize kernel applies 1/gscale[o], the |W|<=1 STE mask and
// the [t][c][o] -> OIHW transposition.
// ===================================================================================================
struct TcWgradParams {
  int32_t OW, OH, NIMG;          // gy pixel grid
  int32_t BW, BH, BNI, tiles_h;  // K box = BNI x BH x BW output pixels
  int32_t rows_box, k_stage;     // valid pixel rows per box, rounded up to 16
  int32_t n_kboxes, kboxes_per_cta;
  int32_t Cin, Cout, kh, kw, pad, stride;
  int32_t g_halves;              // 1: gys = bf16(g); 2: gys = [hi | lo] split, both accumulated
  int32_t chunks_per_tap;        // Cin / 64
  int32_t n_units, G;            // (tap, chunk) units; M tiles (accumulators) per CTA
  int32_t BN;                    // N tile (output channels per CTA)
  int32_t stages;
  float* ws;                     // [T*Cin][Cout] fp32, zero-initialised
};

__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;   // LBO: stride between 64-element MN chunks
  d |= uint64_t(1024u >> 4) << 32;                   // SBO: stride between 8-row K groups
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;                            // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kTcThreads)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                const TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t box_bytes = uint32_t(p.k_stage) * 128u;          // one [k_stage][64ch] bf16 box
  const int mt0 = blockIdx.y * p.G;                               // first M tile of this CTA
  const int n_mtiles = (p.n_units + 1) / 2;
  const int g_cta = min(p.G, n_mtiles - mt0);
  const int nb_chunks = p.BN / 64;                                // 64-channel chunks of the N tile
  const int nb_boxes = nb_chunks * p.g_halves;                    // hi (and lo) boxes of gys
  const uint32_t a_bytes = uint32_t(p.G) * 2u * box_bytes;
  const uint32_t stage_bytes = a_bytes + uint32_t(nb_boxes) * box_bytes;
  const int nn0 = blockIdx.z * p.BN;
  const int kb_begin = blockIdx.x * p.kboxes_per_cta;
  const int kb_end = min(kb_begin + p.kboxes_per_cta, p.n_kboxes);
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(p.G * p.BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    fence_barrier_init();
  }
  // Zero the K-tail rows [rows_box, k_stage) of every box once: TMA never writes them, so they
  // contribute 0 * 0 to every MMA.
  if (p.k_stage > p.rows_box) {
    const int tail_u4 = (p.k_stage - p.rows_box) * 8;             // 16-byte words per box tail
    const int boxes_per_stage = p.G * 2 + nb_boxes;
    const int total = p.stages * boxes_per_stage * tail_u4;
    uint8_t* base_generic = smem_raw + (tiles_base - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int w = i % tail_u4, b = (i / tail_u4) % boxes_per_stage, s = i / (tail_u4 * boxes_per_stage);
      uint4* dst = reinterpret_cast<uint4*>(base_generic + size_t(s) * stage_bytes + size_t(b) * box_bytes +
                                            size_t(p.rows_box) * 128u) + w;
      *dst = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
  }
  if (warp == 5) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t tx = uint32_t(g_cta * 2 + nb_boxes) * uint32_t(p.rows_box) * 128u;
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = uint32_t(it / p.stages) & 1u;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, tx);
        const int tile_n = kb / p.tiles_h, tile_h = kb - tile_n * p.tiles_h;
        const int n0 = tile_n * p.BNI, h0 = tile_h * p.BH;
        const uint32_t dst0 = tiles_base + stage * stage_bytes;
        for (int i = 0; i < g_cta * 2; ++i) {
          int u = mt0 * 2 + i;
          if (u >= p.n_units) u = p.n_units - 1;                  // odd unit count: duplicate, rows ignored
          const int t = u / p.chunks_per_tap, j = u - t * p.chunks_per_tap;
          const int r = t / p.kw, s = t - r * p.kw;
          tma_load_4d(dst0 + i * box_bytes, &tmX, fb, j * 64, s - p.pad, h0 * p.stride + r - p.pad, n0);
        }
        for (int hf = 0; hf < p.g_halves; ++hf)
          for (int jb = 0; jb < nb_chunks; ++jb)
            tma_load_4d(dst0 + a_bytes + (hf * nb_chunks + jb) * box_bytes, &tmG, fb,
                        hf * p.Cout + nn0 + jb * 64, 0, h0, n0);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // M=128, N=BN, A and B MN-major (bits 15/16)
      const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN)) | (1u << 15) | (1u << 16);
      const int k_steps = p.k_stage / 16;
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = uint32_t(it / p.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t a0 = tiles_base + stage * stage_bytes;
        const uint32_t b0 = a0 + a_bytes;
        for (int g = 0; g < g_cta; ++g) {
          for (int hf = 0; hf < p.g_halves; ++hf) {
            for (int k = 0; k < k_steps; ++k) {
              const uint64_t ad = make_mnmajor_desc(a0 + g * 2 * box_bytes + k * 2048, box_bytes);
              const uint64_t bd = make_mnmajor_desc(b0 + hf * nb_chunks * box_bytes + k * 2048, box_bytes);
              umma_bf16(tmem_d + uint32_t(g * p.BN), ad, bd, idesc, (it > 0 || k > 0 || hf > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      umma_commit(smem_u32(&accum_bar));
    }
  } else if (kb_end > kb_begin) {
    const int m = warp * 32 + lane;
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_d + (uint32_t(warp * 32) << 16);
    for (int g = 0; g < g_cta; ++g) {
      const int u = (mt0 + g) * 2 + (m >> 6);
      const bool valid = u < p.n_units;
      float* wrow = p.ws + (int64_t(u) * 64 + (m & 63)) * p.Cout + nn0;   // row (t*Cin + c)
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + uint32_t(g * p.BN + c0), v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(wrow + c0 + j, __uint_as_float(v[j]));
        }
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

// gW[o][c][t] = wmask ? ws[(t*Cin + c)*Cout + o] * inv_gscale[o] : 0
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ ws, const uint32_t* __restrict__ wmask,
                      const float* __restrict__ inv_gscale, float* __restrict__ gW, int Cout, int Cin,
                      int T) {
  const int64_t n = int64_t(Cout) * Cin * T;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(e % T);
    const int64_t oc = e / T;
    const int c = int(oc % Cin), o = int(oc / Cin);
    const bool pass = (wmask[e >> 5] >> (e & 31)) & 1u;
    gW[e] = pass ? ws[(int64_t(t) * Cin + c) * Cout + o] * inv_gscale[o] : 0.0f;
  }
}

struct WgradPlan {
  TcWgradParams p;
  int ksplit, mgroups, ntiles;
  size_t smem;
  bool ok;
};

static WgradPlan plan_wgrad(const bdbnn_conv_shape* s, int halves = 2) {
  WgradPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!s || (s->stride != 1 && s->stride != 2) || s->kh != s->kw || s->kh > 7 || s->pad > s->kh - 1) return pl;
  if (s->Cin % 64 != 0 || s->Cout % 64 != 0) return pl;
  if (s->Cout > 128 && s->Cout % 256 != 0) return pl;
  if (s->Wo > 128 || s->W > 128 * s->stride) return pl;
  TcWgradParams& p = pl.p;
  const int T = s->kh * s->kw;
  p.OW = s->Wo; p.OH = s->Ho; p.NIMG = s->N;
  p.Cin = s->Cin; p.Cout = s->Cout; p.kh = s->kh; p.kw = s->kw; p.pad = s->pad; p.stride = s->stride;
  p.g_halves = halves;
  p.chunks_per_tap = s->Cin / 64;
  p.n_units = T * p.chunks_per_tap;
  const int n_mtiles = (p.n_units + 1) / 2;
  p.BN = s->Cout >= 256 ? 256 : s->Cout;
  p.G = 512 / p.BN;
  if (p.G > n_mtiles) p.G = n_mtiles;
  if (p.G > 5) p.G = 5;
  p.stages = 2;
  // choose the K box: largest pixel-row count whose ring fits ~200 KB, best utilisation first
  const int row_bytes_all = p.G * 256 + p.BN * 2 * halves;      // smem bytes per pixel row per stage
  const int k_cap = int((200u * 1024u) / (unsigned(p.stages) * unsigned(row_bytes_all))) & ~15;
  if (k_cap < 16) return pl;
  const int kmax = k_cap > 128 ? 128 : k_cap;
  p.BW = s->Wo;
  double best = -1.0;
  int bBH = 0, bBNI = 0;
  if (s->Ho * s->Wo <= kmax) {
    for (int ni = 1; ni * s->Ho * s->Wo <= kmax && ni <= s->N && ni <= 256; ++ni) {
      const int rows = ni * s->Ho * s->Wo, ks = (rows + 15) & ~15;
      const double eff = double(rows) / ks * double(s->N) / (double((s->N + ni - 1) / ni) * ni);
      if (eff * (1.0 + 0.02 * ni) > best) { best = eff * (1.0 + 0.02 * ni); bBH = s->Ho; bBNI = ni; }
    }
  } else {
    for (int bh = 1; bh * s->Wo <= kmax && bh <= s->Ho; ++bh) {
      const int rows = bh * s->Wo, ks = (rows + 15) & ~15;
      const double eff = double(rows) / ks * double(s->Ho) / (double((s->Ho + bh - 1) / bh) * bh);
      if (eff * (1.0 + 0.01 * bh) > best) { best = eff * (1.0 + 0.01 * bh); bBH = bh; bBNI = 1; }
    }
  }
  if (bBH == 0) return pl;
  p.BH = bBH; p.BNI = bBNI;
  p.rows_box = p.BNI * p.BH * p.BW;
  p.k_stage = (p.rows_box + 15) & ~15;
  p.tiles_h = (s->Ho + p.BH - 1) / p.BH;
  p.n_kboxes = p.tiles_h * ((s->N + p.BNI - 1) / p.BNI);
  pl.mgroups = (n_mtiles + p.G - 1) / p.G;
  pl.ntiles = s->Cout / p.BN;
  int ks = (num_sms() + pl.mgroups * pl.ntiles - 1) / (pl.mgroups * pl.ntiles);
  if (ks < 1) ks = 1;
  if (ks > p.n_kboxes) ks = p.n_kboxes;
  p.kboxes_per_cta = (p.n_kboxes + ks - 1) / ks;
  pl.ksplit = (p.n_kboxes + p.kboxes_per_cta - 1) / p.kboxes_per_cta;
  pl.smem = size_t(p.stages) * size_t(row_bytes_all) * size_t(p.k_stage) + 1024;
  pl.ok = pl.smem <= 227u * 1024u;
  return pl;
}

}  // namespace bdbnn

using namespace bdbnn;

extern "C" int bdbnn_debug_trace(long long* device_buf) {
  set_tc_trace(device_buf);
  return BDBNN_OK;
}

extern "C" int bdbnn_tc_supported(const bdbnn_conv_shape* s) {
  if (!tc_shape_ok(s)) return 0;
  return BDBNN_TC_FWD | BDBNN_TC_DGRAD | (plan_wgrad(s, 2).ok ? BDBNN_TC_WGRAD : 0);
}

extern "C" int bdbnn_binconv_fwd_tc(const uint16_t* xb_bf16, const uint16_t* wf_bf16, const float* alpha,
                                    float* y, const bdbnn_conv_shape* s, void* stream) {
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
  L.alpha = alpha; L.out = y;
  return launch_tc_conv<0>(L, cudaStream_t(stream));
}

extern "C" int bdbnn_binconv_dgrad_tc(const uint16_t* gys_bf16, int32_t grad_halves,
                                      const uint16_t* wt_bf16, const uint32_t* mask_bits, float* gx,
                                      const bdbnn_conv_shape* s, void* stream) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(gys_bf16 && wt_bf16 && mask_bits && gx, "binconv_dgrad_tc: NULL pointer");
  BDBNN_REQUIRE(grad_halves == 1 || grad_halves == 2, "binconv_dgrad_tc: grad_halves must be 1 or 2");
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
      ++n_launch;
    }
  if (empty_phase)
    BDBNN_CUDA(cudaMemsetAsync(gx, 0, size_t(s->N) * s->H * s->W * s->Cin * sizeof(float), st));
  for (int i = 0; i < n_launch; ++i) {
    rc = launch_tc_conv<1>(Ls[i], st);
    if (rc) return rc;
  }
  return BDBNN_OK;
}

extern "C" size_t bdbnn_wgrad_tc_workspace_bytes(const bdbnn_conv_shape* s) {
  if (!s || !plan_wgrad(s, 2).ok) return 0;
  return size_t(s->kh) * s->kw * s->Cin * s->Cout * sizeof(float);
}

extern "C" int bdbnn_binconv_wgrad_tc(const uint16_t* gys_bf16, int32_t grad_halves, const uint16_t* xb_bf16,
                                      const uint32_t* wmask_bits, const float* inv_gscale, float* gW,
                                      const bdbnn_conv_shape* s, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(gys_bf16 && xb_bf16 && wmask_bits && inv_gscale && gW && workspace,
                "binconv_wgrad_tc: NULL pointer");
  BDBNN_REQUIRE(grad_halves == 1 || grad_halves == 2, "binconv_wgrad_tc: grad_halves must be 1 or 2");
  WgradPlan pl = plan_wgrad(s, grad_halves);
  if (!pl.ok) { set_error("binconv_wgrad_tc: shape not supported by the tcgen05 path"); return BDBNN_ERR_UNSUPPORTED; }
  const size_t need = bdbnn_wgrad_tc_workspace_bytes(s);
  if (workspace_bytes < need) {
    set_error("binconv_wgrad_tc: workspace %zu B < required %zu B", workspace_bytes, need);
    return BDBNN_ERR_WORKSPACE;
  }
  cudaStream_t st = cudaStream_t(stream);
  pl.p.ws = static_cast<float*>(workspace);
  BDBNN_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  CUtensorMap tmX, tmG;
  rc = make_act_map(&tmX, xb_bf16, s->N, s->H, s->W, s->Cin, 64, pl.p.BW, pl.p.BH, pl.p.BNI, s->stride);
  if (rc) return rc;
  rc = make_act_map(&tmG, gys_bf16, s->N, s->Ho, s->Wo, s->Cout * grad_halves, 64, pl.p.BW, pl.p.BH, pl.p.BNI);
  if (rc) return rc;
  BDBNN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
  dim3 grid(unsigned(pl.ksplit), unsigned(pl.mgroups), unsigned(pl.ntiles));
  tc_wgrad_kernel<<<grid, kTcThreads, pl.smem, st>>>(tmX, tmG, pl.p);
  rc = check_launch("tc_wgrad_kernel");
  if (rc) return rc;
  const int64_t n = int64_t(s->Cout) * s->Cin * s->kh * s->kw;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  wgrad_finalize_kernel<<<unsigned(blocks), 256, 0, st>>>(pl.p.ws, wmask_bits, inv_gscale, gW, s->Cout,
                                                          s->Cin, s->kh * s->kw);
  return check_launch("wgrad_finalize_kernel");
}
