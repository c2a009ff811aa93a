// Weight gradient on tensor cores (tcgen05 + TMA, sm_100a).
//
//   D[(t,c), o] = sum_pix  xb[pix shifted by tap t, c] * gys[pix, o]          (K = output pixels)
//
// Both operands are "MN-major" for this GEMM: for a fixed pixel the M index (c) / N index (o) is the
// contiguous one in NHWC memory, so a TMA box [pixels][64 channels] lands in shared memory exactly in
// the canonical MN-major SWIZZLE_128B layout (8 pixel rows x 128 B per swizzle atom; SBO = 1024 B
// between 8-row K groups; LBO = distance between the two 64-channel chunks of an M tile).
// An M tile of 128 rows is two 64-channel "units" (tap, channel-chunk); unit u = tap * (Cin/64) + chunk.
//
// HALO mode (stride 1, images larger than 128 pixels): per K stage ONE activation patch per
// 64-channel chunk is loaded — BH(+halo) image rows in padded-width raster order — and every tap is
// that patch viewed from row offset (dh-dh_min)*PW + (dw-dw_min) (descriptor start + 128 B * shift;
// the swizzle XOR is taken from absolute smem address bits, so row-shifted views are valid operands).
// The gys box uses the same padded raster: its columns >= Wo fall outside the tensor and TMA zero-fills
// them, so the padded K rows contribute exactly 0.  L2->SM traffic drops ~3x versus one box per tap.
// BOX mode (small images, stride 2): one TMA box per unit, as before.
//
// Each CTA owns up to G accumulators (G*BN <= 512 TMEM columns) that share the gys stage, walks a
// contiguous range of K stages (split-K across CTAs) and writes its partial tiles to its own slice of
// an fp32 workspace [ksplit][T*Cin][Cout] (plain stores: a red.global.add flush of 148 CTAs onto the
// same 37 K addresses cost 57 K cycles, 30 % of the layer-1 kernel); the finalize kernel sums the
// slices in a fixed order (run-to-run deterministic), applies 1/gscale[o], the FP16S 2^-e, the
// |W|<=1 STE mask and the [t][c][o] -> OIHW transposition.
#include "tc_common.cuh"

namespace bdbnn {

constexpr int kMaxUnits = 16;      // units per CTA (G <= 5 M tiles of 2 units, or 2 tiles of 8 16-wide units)

struct TcWgradParams {
  int32_t OW, OH, NIMG;          // gy pixel grid
  int32_t halo;
  int32_t BW, BH, BNI, tiles_h;  // K stage = BNI x BH x (BW | PW) output pixels
  int32_t PW, PH, dh_min, dw_min;
  int32_t rows_a, rows_b;        // rows TMA writes per A box/patch and per B box
  int32_t k_stage;               // K rows consumed per stage (multiple of 16)
  uint32_t a_box_bytes;          // smem footprint of one A box/patch (1024-aligned)
  uint32_t b_box_bytes;
  int32_t n_a_boxes;             // A boxes/patches per stage
  int32_t n_kboxes, kboxes_per_cta;
  int32_t Cin, Cout, kh, kw, pad, stride;
  int32_t g_halves;              // 1: gys = bf16(g); 2: gys = [hi | lo] split, both accumulated
  int32_t UW;                    // unit width in channels: 64 (128-byte rows, SWIZZLE_128B) or 32 (stem
                                 // windows: 64-byte rows, SWIZZLE_64B); an M tile is 128/UW units
  int32_t chunks_per_tap;        // Cin / 64
  int32_t n_units, G;            // (tap, chunk) units; M tiles (accumulators) per CTA
  int32_t BN;                    // N tile (output channels per CTA)
  int32_t stages;
  int32_t fmt;                   // operand format (BDBNN_FMT_*)
  float* ws;                     // [ksplit][T*Cin][Cout] fp32 partials (every valid entry is written)
  long long* trace;              // optional clock64 trace of CTA (0,0,0) (bdbnn_debug_trace)
};

#define BDBNN_WTR(r, ev)                                                                         \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tr_n < 1023) { \
      p.trace[(r) * 2048 + 2 * tr_n] = (ev);                                                     \
      p.trace[(r) * 2048 + 2 * tr_n + 1] = clock64();                                            \
      ++tr_n;                                                                                    \
    }                                                                                            \
  } while (0)

// row_bytes 128: SWIZZLE_128B atoms of 8 K rows x 64 MN elements; 64: SWIZZLE_64B, 8 x 32; 32: SWIZZLE_32B, 8 x 16.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t row_bytes = 128u) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;   // LBO: stride between MN chunks (one swizzle row each)
  d |= uint64_t((8u * row_bytes) >> 4) << 32;        // SBO: stride between 8-row K groups
  d |= uint64_t(1) << 46;
  d |= uint64_t(row_bytes == 128u ? 2u : (row_bytes == 64u ? 4u : 6u)) << 61;  // SWIZZLE_128B / _64B / _32B
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
  __shared__ uint32_t unit_off[kMaxUnits];   // byte offset of unit i's operand view inside a stage
  __shared__ int32_t unit_c[kMaxUnits], unit_dw[kMaxUnits], unit_dh[kMaxUnits];   // box mode: TMA coordinates of unit i

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int mt0 = blockIdx.y * p.G;                               // first M tile of this CTA
  const int upt = kTileM / p.UW;                                  // units per M tile (2 or 4)
  const uint32_t a_row = uint32_t(p.UW) * 2u;                     // bytes per A row
  const int n_mtiles = (p.n_units + upt - 1) / upt;
  const int g_cta = min(p.G, n_mtiles - mt0);
  const int nb_chunks = p.BN >= 64 ? p.BN / 64 : 1;               // 64-channel chunks of the N tile (or one narrow one)
  const uint32_t b_row = p.BN >= 64 ? 128u : uint32_t(p.BN) * 2u;  // bytes per gys row of one chunk
  const int nb_boxes = nb_chunks * p.g_halves;                    // hi (and lo) boxes of gys
  const uint32_t a_bytes = uint32_t(p.n_a_boxes) * p.a_box_bytes;
  const uint32_t stage_bytes = a_bytes + uint32_t(nb_boxes) * p.b_box_bytes;
  const int nn0 = blockIdx.z * p.BN;
  const int kb_begin = blockIdx.x * p.kboxes_per_cta;
  const int kb_end = min(kb_begin + p.kboxes_per_cta, p.n_kboxes);
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(p.G * p.BN)) tmem_cols <<= 1;

  if (threadIdx.x < g_cta * upt) {
    int u = mt0 * upt + threadIdx.x;
    bool dup = false;
    if (u >= p.n_units) { u = p.n_units - 1; dup = true; }        // odd unit count: rows ignored
    const int t = u / p.chunks_per_tap, j = u - t * p.chunks_per_tap;
    uint32_t off;
    if (p.halo) {
      const int r = t / p.kw, s = t - r * p.kw;
      off = uint32_t(j) * p.a_box_bytes +
            uint32_t((r - p.pad - p.dh_min) * p.PW + (s - p.pad - p.dw_min)) * 128u;
      if (dup) off += 128u;                                        // keep LBO > 0
    } else {
      off = threadIdx.x * p.a_box_bytes;
    }
    unit_off[threadIdx.x] = off;
    // the producer is ONE thread: everything per unit that needs a division is computed here, once (ncu of the
    // layer4 launch showed the per-stage address arithmetic of 8 loads — ~750 dependent instructions — at 4.4 K
    // cycles per stage, four times the MMA time)
    const int r = t / p.kw, sx = t - r * p.kw;
    unit_c[threadIdx.x] = j * 64;
    unit_dw[threadIdx.x] = sx - p.pad;
    unit_dh[threadIdx.x] = r - p.pad;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    fence_barrier_init();
  }
  // Zero every row TMA never writes (A rows >= rows_a, B rows >= rows_b): they are read as K rows.
  {
    uint8_t* base_generic = smem_raw + (tiles_base - smem_u32(smem_raw));
    const int a_w16 = int(a_row / 16u);                            // 16-byte words per A row
    const int a_tail = int(p.a_box_bytes / 16) - p.rows_a * a_w16;
    const int b_w16 = int(b_row / 16u);
    const int b_tail = int(p.b_box_bytes / 16) - p.rows_b * b_w16;
    const int per_stage = p.n_a_boxes * a_tail + nb_boxes * b_tail;
    const int total = p.stages * per_stage;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int s = i / per_stage;
      int r = i - s * per_stage;
      uint8_t* dst;
      if (r < p.n_a_boxes * a_tail) {
        const int b = r / a_tail, w = r - b * a_tail;
        dst = base_generic + size_t(s) * stage_bytes + size_t(b) * p.a_box_bytes + size_t(p.rows_a) * a_row +
              size_t(w) * 16u;
      } else {
        r -= p.n_a_boxes * a_tail;
        const int b = r / b_tail, w = r - b * b_tail;
        dst = base_generic + size_t(s) * stage_bytes + a_bytes + size_t(b) * p.b_box_bytes +
              size_t(p.rows_b) * b_row + size_t(w) * 16u;
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
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
      const int n_a_loads = p.halo ? p.n_a_boxes : g_cta * upt;
      const uint32_t tx = uint32_t(n_a_loads) * uint32_t(p.rows_a) * a_row +
                          uint32_t(nb_boxes) * uint32_t(p.rows_b) * b_row;
      int tr_n = 0;
      int stage = 0;
      uint32_t phase = 0;
      int tile_n = kb_begin / p.tiles_h, tile_h = kb_begin - tile_n * p.tiles_h;   // advanced incrementally
      const int n_units_cta = g_cta * upt;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        BDBNN_WTR(0, 0);
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        BDBNN_WTR(0, 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, tx);
        const int n0 = tile_n * p.BNI, h0 = tile_h * p.BH;
        const uint32_t dst0 = tiles_base + stage * stage_bytes;
        if (p.halo) {
          for (int j = 0; j < p.n_a_boxes; ++j)
            tma_load_4d(dst0 + j * p.a_box_bytes, &tmX, fb, j * 64, p.dw_min, h0 + p.dh_min, n0);
        } else {
          const int hs = h0 * p.stride;
          for (int i = 0; i < n_units_cta; ++i)
            tma_load_4d(dst0 + i * p.a_box_bytes, &tmX, fb, unit_c[i], unit_dw[i], hs + unit_dh[i], n0);
        }
        uint32_t bdst = dst0 + a_bytes;
        for (int hf = 0; hf < p.g_halves; ++hf)
          for (int jb = 0; jb < nb_chunks; ++jb, bdst += p.b_box_bytes)
            tma_load_4d(bdst, &tmG, fb, hf * p.Cout + nn0 + jb * 64, 0, h0, n0);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        if (++tile_h == p.tiles_h) { tile_h = 0; ++tile_n; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // M=128, N=BN, A and B MN-major (bits 15/16)
      const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.BN), uint32_t(p.fmt)) | (1u << 15) | (1u << 16);
      const int k_steps = p.k_stage / 16;
      int it = 0, tr_n = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        BDBNN_WTR(1, 0);
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        BDBNN_WTR(1, 1);
        const uint32_t a0 = tiles_base + stage * stage_bytes;
        const uint32_t b0 = a0 + a_bytes;
        for (int g = 0; g < g_cta; ++g) {
          const uint32_t ua = unit_off[upt * g], ub = unit_off[upt * g + 1];
          const uint64_t ad0 = make_mnmajor_desc(a0 + ua, ub - ua, a_row);
          const uint32_t a_kstep = a_row;                           // 16 K rows x a_row bytes, >> 4
          const uint32_t a_lo0 = uint32_t(ad0), a_hi = uint32_t(ad0 >> 32);
          const uint32_t acc = tmem_d + uint32_t(g * p.BN);
          for (int hf = 0; hf < p.g_halves; ++hf) {
            const uint64_t bd0 = make_mnmajor_desc(b0 + hf * nb_chunks * p.b_box_bytes, p.b_box_bytes, b_row);
            const uint32_t b_kstep = b_row;                           // 16 K rows x b_row bytes, >> 4
            const uint32_t b_lo0 = uint32_t(bd0), b_hi = uint32_t(bd0 >> 32);
            // 16 pixel rows per K step: +2048 B (128-byte rows) = +128 in the low word's address field
            umma_bf16_split(acc, a_lo0, a_hi, b_lo0, b_hi, idesc, (it > 0 || hf > 0) ? 1u : 0u);
#pragma unroll 4
            for (int k = 1; k < k_steps; ++k)
              umma_bf16_split(acc, a_lo0 + uint32_t(k) * a_kstep, a_hi, b_lo0 + uint32_t(k) * b_kstep, b_hi, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        BDBNN_WTR(1, 4);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(&accum_bar));
    }
  } else if (kb_end > kb_begin) {
    const int m = warp * 32 + lane;
    int tr_n = threadIdx.x == 0 ? 0 : 100000;
    BDBNN_WTR(2, 0);
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    BDBNN_WTR(2, 1);
    const uint32_t lane_base = tmem_d + (uint32_t(warp * 32) << 16);
    for (int g = 0; g < g_cta; ++g) {
      const int u = (mt0 + g) * upt + m / p.UW;
      const bool valid = u < p.n_units;
      float* wrow = p.ws + int64_t(blockIdx.x) * (int64_t(p.n_units) * p.UW * p.Cout) +
                    (int64_t(u) * p.UW + (m % p.UW)) * p.Cout + nn0;       // slice, row (t*Cin + c)
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + uint32_t(g * p.BN + c0), v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(wrow + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
        }
      }
    }
    BDBNN_WTR(2, 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// gW[o][c][t] = wmask ? inv_gscale[o] * 2^-e * sum_k ws[k][(t*Cin + c)*Cout + o] : 0
// One thread per [t*Cin+c][o] entry: the ksplit partials are read coalesced and added in slice order.
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ ws, int ksplit, const uint32_t* __restrict__ wmask,
                      const float* __restrict__ inv_gscale, const uint32_t* __restrict__ amax_bits,
                      float* __restrict__ gW, int Cout, int Cin, int T) {
  const int64_t n = int64_t(Cout) * Cin * T;
  const float post = amax_bits ? amax_pow2_scale(__ldg(amax_bits), true) : 1.0f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int o = int(i % Cout);
    const int64_t tc = i / Cout;
    const int c = int(tc % Cin), t = int(tc / Cin);
    const int64_t e = (int64_t(o) * Cin + c) * T + t;              // OIHW flat index
    float acc = 0.f;
    if ((wmask[e >> 5] >> (e & 31)) & 1u) {
      // fixed summation order (deterministic); four independent chains hide the load latency
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 3 < ksplit; k += 4) {
        a0 += ws[int64_t(k) * n + i];
        a1 += ws[int64_t(k + 1) * n + i];
        a2 += ws[int64_t(k + 2) * n + i];
        a3 += ws[int64_t(k + 3) * n + i];
      }
      for (; k < ksplit; ++k) a0 += ws[int64_t(k) * n + i];
      acc = ((a0 + a1) + (a2 + a3)) * post * inv_gscale[o];
    }
    gW[e] = acc;
  }
}

struct WgradPlan {
  TcWgradParams p;
  int ksplit, ks_cap, mgroups, ntiles;
  size_t smem;
  bool ok;
};

static WgradPlan plan_wgrad_small(const bdbnn_conv_shape* s, int halves);

static WgradPlan plan_wgrad(const bdbnn_conv_shape* s, int halves) {
  WgradPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!s || (s->stride != 1 && s->stride != 2) || s->kh != s->kw || s->kh > 7 || s->pad > s->kh - 1) return pl;
  if (s->Cin == 16 || s->Cin == 32 || s->Cout == 16 || s->Cout == 32) return plan_wgrad_small(s, halves);
  if (s->Cin % 64 != 0 || s->Cout % 64 != 0) return pl;
  if (s->Cout > 128 && s->Cout % 256 != 0) return pl;
  if (s->Wo > 128 || s->W > 128 * s->stride) return pl;
  static const int halo_env = [] { const char* e = getenv("BDBNN_WGRAD_HALO"); return e ? atoi(e) : 1; }();
  TcWgradParams& p = pl.p;
  const int T = s->kh * s->kw;
  p.OW = s->Wo; p.OH = s->Ho; p.NIMG = s->N;
  p.Cin = s->Cin; p.Cout = s->Cout; p.kh = s->kh; p.kw = s->kw; p.pad = s->pad; p.stride = s->stride;
  p.g_halves = halves;
  p.UW = 64;
  p.chunks_per_tap = s->Cin / 64;
  p.n_units = T * p.chunks_per_tap;
  const int n_mtiles = (p.n_units + 1) / 2;
  p.BN = s->Cout >= 256 ? 256 : s->Cout;
  p.G = 512 / p.BN;
  if (p.G > n_mtiles) p.G = n_mtiles;
  if (p.G > 5) p.G = 5;
  {  // balance the M groups (e.g. 9 M tiles with G<=4 -> 3 groups of 3, not 4+4+1)
    const int groups = (n_mtiles + p.G - 1) / p.G;
    p.G = (n_mtiles + groups - 1) / groups;
  }
  p.stages = 2;
  const uint32_t budget = 222u * 1024u;
  const int span = s->kh - 1;                                     // tap offsets span [-pad, k-1-pad]
  const int PW = s->Wo + span;
  bool planned = false;
  if (halo_env && s->stride == 1 && s->Ho * s->Wo > kTileM && PW <= 256) {
    // halo mode: patches hold (BH + span) x PW pixels; pick the largest BH whose 2-stage ring fits
    for (int bh = (s->Ho < 16 ? s->Ho : 16); bh >= 1 && !planned; --bh) {
      const int k_stage = (bh * PW + 15) & ~15;
      if (k_stage > 512) continue;
      const int rows_a = (bh + span) * PW;
      const int view_rows = k_stage + span * PW + span + 1;       // rows a shifted view can touch
      const uint32_t a_box = (uint32_t(max(rows_a, view_rows)) * 128u + 1023u) & ~1023u;
      const uint32_t b_box = uint32_t(k_stage) * 128u;
      const uint32_t stage = uint32_t(p.chunks_per_tap) * a_box + uint32_t(p.BN / 64 * halves) * b_box;
      if (2u * stage > budget) continue;
      // prefer row counts that waste little of the last stage of an image
      const int tiles_h = (s->Ho + bh - 1) / bh;
      if (tiles_h * bh - s->Ho > bh / 2 && bh > 1) continue;
      p.halo = 1; p.PW = PW; p.PH = bh + span; p.dh_min = -s->pad; p.dw_min = -s->pad;
      p.BW = PW; p.BH = bh; p.BNI = 1; p.tiles_h = tiles_h;
      p.rows_a = rows_a; p.rows_b = bh * PW; p.k_stage = k_stage;
      p.a_box_bytes = a_box; p.b_box_bytes = b_box; p.n_a_boxes = p.chunks_per_tap;
      planned = true;
    }
  }
  if (!planned) {
    // box mode: one [rows][64ch] box per unit
    const int row_bytes_all = p.G * 256 + p.BN * 2 * halves;      // smem bytes per pixel row per stage
    const int k_cap = int(budget / (unsigned(p.stages) * unsigned(row_bytes_all))) & ~15;
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
    p.rows_a = p.rows_b = p.BNI * p.BH * p.BW;
    p.k_stage = (p.rows_a + 15) & ~15;
    p.a_box_bytes = p.b_box_bytes = uint32_t(p.k_stage) * 128u;
    p.n_a_boxes = p.G * 2;
    p.tiles_h = (s->Ho + p.BH - 1) / p.BH;
  }
  p.n_kboxes = p.tiles_h * ((s->N + p.BNI - 1) / p.BNI);
  pl.mgroups = (n_mtiles + p.G - 1) / p.G;
  pl.ntiles = s->Cout / p.BN;
  // split K so that the whole grid is ONE wave (<= #SMs CTAs): 1 CTA/SM by shared-memory size, and a
  // second partial wave would double the kernel time
  int ks = num_sms() / (pl.mgroups * pl.ntiles);
  if (ks < 1) ks = 1;
  pl.ks_cap = ks;                       // upper bound of ksplit for any gradient mode (workspace sizing)
  if (ks > p.n_kboxes) ks = p.n_kboxes;
  p.kboxes_per_cta = (p.n_kboxes + ks - 1) / ks;
  pl.ksplit = (p.n_kboxes + p.kboxes_per_cta - 1) / p.kboxes_per_cta;
  const size_t stage_bytes = size_t(p.n_a_boxes) * p.a_box_bytes + size_t(p.BN / 64 * halves) * p.b_box_bytes;
  // Box-mode stages are small (32-64 K rows) and the loop is bound by the TMA round trip, not by bytes: use every
  // stage the shared memory holds (BDBNN_WG_STAGES=2 restores the two-stage ring for A/B runs).
  static const int stages_env = [] { const char* e = getenv("BDBNN_WG_STAGES"); return e ? atoi(e) : 0; }();
  int stages = stages_env > 0 ? stages_env : int(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.kboxes_per_cta) stages = p.kboxes_per_cta;
  if (stages < 2) stages = 2;
  if (size_t(stages) * stage_bytes > budget) stages = 2;
  p.stages = stages;
  pl.smem = size_t(p.stages) * stage_bytes + 1024;
  pl.ok = pl.smem <= 227u * 1024u;
  return pl;
}

// CIFAR-sized layers (ResNet-20: 16/32/64 channels): box mode with narrow units — A rows of Cin*2 bytes
// (16 channels: SWIZZLE_32B, 32: SWIZZLE_64B, 64: SWIZZLE_128B), one gys box of Cout channels as the N
// tile.  All M tiles of the layer live in one CTA when they fit (T*Cin/128 accumulators of Cout columns).
static WgradPlan plan_wgrad_small(const bdbnn_conv_shape* s, int halves) {
  WgradPlan pl;
  memset(&pl, 0, sizeof(pl));
  auto ok_c = [](int c) { return c == 16 || c == 32 || c == 64; };
  if (!ok_c(s->Cin) || !ok_c(s->Cout)) return pl;
  if (s->Wo > 128 || s->W > 128 * s->stride) return pl;
  TcWgradParams& p = pl.p;
  const int T = s->kh * s->kw;
  p.OW = s->Wo; p.OH = s->Ho; p.NIMG = s->N;
  p.Cin = s->Cin; p.Cout = s->Cout; p.kh = s->kh; p.kw = s->kw; p.pad = s->pad; p.stride = s->stride;
  p.g_halves = halves;
  p.UW = s->Cin;
  p.chunks_per_tap = 1;
  p.n_units = T;
  const int upt = kTileM / p.UW;
  const int n_mtiles = (p.n_units + upt - 1) / upt;
  p.BN = s->Cout;
  p.G = kMaxUnits / upt;
  if (p.G > n_mtiles) p.G = n_mtiles;
  if (p.G * p.BN > 512) p.G = 512 / p.BN;
  p.halo = 0;
  p.BW = s->Wo;
  const int a_row = p.UW * 2, b_row = p.BN >= 64 ? 128 : p.BN * 2;
  // K stage: whole images when they are small, else rows of one image; <= 128 pixel rows
  if (s->Ho * s->Wo <= kTileM) {
    p.BH = s->Ho;
    p.BNI = kTileM / (s->Ho * s->Wo);
    if (p.BNI > s->N) p.BNI = s->N;
  } else {
    p.BH = kTileM / s->Wo;
    p.BNI = 1;
  }
  p.rows_a = p.rows_b = p.BNI * p.BH * p.BW;
  p.k_stage = (p.rows_a + 15) & ~15;
  p.a_box_bytes = (uint32_t(p.k_stage) * uint32_t(a_row) + 1023u) & ~1023u;
  p.b_box_bytes = (uint32_t(p.k_stage) * uint32_t(b_row) + 1023u) & ~1023u;
  p.tiles_h = (s->Ho + p.BH - 1) / p.BH;
  p.n_kboxes = p.tiles_h * ((s->N + p.BNI - 1) / p.BNI);
  // fewer M tiles per CTA until a 2-stage ring fits in shared memory
  size_t stage_bytes = 0;
  for (;; --p.G) {
    p.n_a_boxes = p.G * upt;
    stage_bytes = size_t(p.n_a_boxes) * p.a_box_bytes + size_t(halves) * p.b_box_bytes;
    p.stages = int((226u * 1024u) / stage_bytes);
    if (p.stages >= 2 || p.G == 1) break;
  }
  if (p.stages > 4) p.stages = 4;
  pl.mgroups = (n_mtiles + p.G - 1) / p.G;
  pl.ntiles = 1;
  int ks = num_sms() / pl.mgroups;
  if (ks < 1) ks = 1;
  pl.ks_cap = ks;
  if (ks > p.n_kboxes) ks = p.n_kboxes;
  p.kboxes_per_cta = (p.n_kboxes + ks - 1) / ks;
  pl.ksplit = (p.n_kboxes + p.kboxes_per_cta - 1) / p.kboxes_per_cta;
  pl.smem = size_t(p.stages) * stage_bytes + 1024;
  pl.ok = p.stages >= 2 && pl.smem <= 227u * 1024u;
  return pl;
}

bool wgrad_tc_ok(const bdbnn_conv_shape* s) { return s && plan_wgrad(s, 2).ok; }

// ---- stem conv weight gradient ------------------------------------------------------------------
// D[(r, j), o] = sum_pix xw[window(pix) of row 2*oh + r][j] * gys[pix, o]: 7 units (rows r of the 7x7
// filter) of 32 window values (8 pixels x 4 halves; j = s*4 + c), box mode, one K stage per output row
// group.  Both M tiles (units 0-3, 4-6 + one duplicate) live in every CTA, so gys is fetched once.
static WgradPlan plan_stem_wgrad(const StemGeom& g) {
  WgradPlan pl;
  memset(&pl, 0, sizeof(pl));
  TcWgradParams& p = pl.p;
  p.OW = g.Wo; p.OH = g.Ho; p.NIMG = g.N;
  p.Cin = kStemWin; p.Cout = kStemCout; p.kh = kStemTaps; p.kw = 1; p.pad = 0; p.stride = 2;
  p.g_halves = 1;
  p.UW = kStemWin;
  p.chunks_per_tap = 1;
  p.n_units = kStemTaps;
  p.BN = kStemCout;
  p.G = 2;
  p.stages = 3;
  p.BW = g.Wo;
  // K stage: BH output rows (or BNI whole images) of Wo pixels, <= 128 rows
  if (g.Ho * g.Wo <= kTileM) {
    p.BH = g.Ho;
    p.BNI = kTileM / (g.Ho * g.Wo);
    if (p.BNI > g.N) p.BNI = g.N;
  } else {
    p.BH = kTileM / g.Wo;
    p.BNI = 1;
  }
  p.rows_a = p.rows_b = p.BNI * p.BH * p.BW;
  p.k_stage = (p.rows_a + 15) & ~15;
  p.a_box_bytes = (uint32_t(p.k_stage) * 64u + 1023u) & ~1023u;
  p.b_box_bytes = uint32_t(p.k_stage) * 128u;
  p.n_a_boxes = p.G * (kTileM / p.UW);
  p.tiles_h = (g.Ho + p.BH - 1) / p.BH;
  p.n_kboxes = p.tiles_h * ((g.N + p.BNI - 1) / p.BNI);
  pl.mgroups = 1;
  pl.ntiles = 1;
  int ks = num_sms();
  pl.ks_cap = ks;
  if (ks > p.n_kboxes) ks = p.n_kboxes;
  p.kboxes_per_cta = (p.n_kboxes + ks - 1) / ks;
  pl.ksplit = (p.n_kboxes + p.kboxes_per_cta - 1) / p.kboxes_per_cta;
  const size_t stage_bytes = size_t(p.n_a_boxes) * p.a_box_bytes + p.b_box_bytes;
  p.stages = int((226u * 1024u) / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  pl.smem = size_t(p.stages) * stage_bytes + 1024;
  pl.ok = p.stages >= 2 && pl.smem <= 227u * 1024u;
  return pl;
}

size_t stem_wgrad_workspace_bytes(const StemGeom& g) {
  const WgradPlan pl = plan_stem_wgrad(g);
  return pl.ok ? size_t(pl.ks_cap) * kStemTaps * kStemWin * kStemCout * sizeof(float) : 0;
}

int launch_stem_wgrad(const uint16_t* gys, const uint16_t* xw, float* ws, size_t ws_bytes, int* ksplit_out,
                      const StemGeom& g, cudaStream_t st) {
  WgradPlan pl = plan_stem_wgrad(g);
  if (!pl.ok) { set_error("stem_conv_wgrad: geometry not supported"); return BDBNN_ERR_UNSUPPORTED; }
  if (ws_bytes < stem_wgrad_workspace_bytes(g)) {
    set_error("stem_conv_wgrad: workspace %zu B too small", ws_bytes);
    return BDBNN_ERR_WORKSPACE;
  }
  pl.p.fmt = BDBNN_FMT_FP16;
  pl.p.ws = ws;
  pl.p.trace = get_tc_trace();
  CUtensorMap tmX, tmG;
  int rc = make_window_map(&tmX, xw, g.N, g.HP, g.Wo, kStemWin, g.win_stride, g.row_stride, g.img_stride, pl.p.BW,
                           pl.p.BH, pl.p.BNI, 2);
  if (rc) return rc;
  rc = make_act_map(&tmG, gys, g.N, g.Ho, g.Wo, kStemCout, 64, pl.p.BW, pl.p.BH, pl.p.BNI);
  if (rc) return rc;
  BDBNN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
  dim3 grid(unsigned(pl.ksplit), 1, 1);
  tc_wgrad_kernel<<<grid, kTcThreads, pl.smem, st>>>(tmX, tmG, pl.p);
  *ksplit_out = pl.ksplit;
  return check_launch("tc_wgrad_kernel(stem)");
}

}  // namespace bdbnn

using namespace bdbnn;

// Host-only view of the wgrad tiling decision (no CUDA call): lets the CPU test-suite sweep shapes.
extern "C" int bdbnn_debug_wgrad_plan(const bdbnn_conv_shape* s, int32_t grad_halves, int32_t* out, int32_t n_out) {
  BDBNN_REQUIRE(s && out && n_out >= 12, "debug_wgrad_plan: need a shape and room for 12 ints");
  BDBNN_REQUIRE(grad_halves == 1 || grad_halves == 2, "debug_wgrad_plan: grad_halves must be 1 or 2");
  const WgradPlan pl = plan_wgrad(s, grad_halves);
  const int32_t v[12] = {pl.ok ? 1 : 0, pl.ksplit, pl.ks_cap, pl.mgroups, pl.ntiles, int32_t(pl.smem), pl.p.G, pl.p.BN,
                         pl.p.UW, pl.p.halo, pl.p.k_stage, pl.p.stages};
  memcpy(out, v, sizeof(v));
  return BDBNN_OK;
}

extern "C" size_t bdbnn_wgrad_tc_workspace_bytes(const bdbnn_conv_shape* s) {
  if (!s) return 0;
  // the split count may depend on the gradient mode (narrow-channel plan): size for the larger one
  const WgradPlan p2 = plan_wgrad(s, 2), p1 = plan_wgrad(s, 1);
  if (!p2.ok) return 0;
  const int cap = p1.ok && p1.ks_cap > p2.ks_cap ? p1.ks_cap : p2.ks_cap;
  return size_t(cap) * s->kh * s->kw * s->Cin * s->Cout * sizeof(float);
}

extern "C" int bdbnn_binconv_wgrad_tc(const uint16_t* gys_bf16, int32_t grad_mode, const uint32_t* amax_bits,
                                      const uint16_t* xb_bf16,
                                      const uint32_t* wmask_bits, const float* inv_gscale, float* gW,
                                      const bdbnn_conv_shape* s, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  int rc = validate_shape(s);
  if (rc) return rc;
  BDBNN_REQUIRE(gys_bf16 && xb_bf16 && wmask_bits && inv_gscale && gW && workspace,
                "binconv_wgrad_tc: NULL pointer");
  BDBNN_REQUIRE(grad_mode >= BDBNN_GRAD_BF16 && grad_mode <= BDBNN_GRAD_FP16S, "binconv_wgrad_tc: bad grad_mode");
  BDBNN_REQUIRE(grad_mode != BDBNN_GRAD_FP16S || amax_bits, "binconv_wgrad_tc: FP16S needs amax_bits");
  const int grad_halves = grad_mode == BDBNN_GRAD_BF16X2 ? 2 : 1;
  WgradPlan pl = plan_wgrad(s, grad_halves);
  pl.p.fmt = grad_mode == BDBNN_GRAD_FP16S ? BDBNN_FMT_FP16 : BDBNN_FMT_BF16;
  if (!pl.ok) { set_error("binconv_wgrad_tc: shape not supported by the tcgen05 path"); return BDBNN_ERR_UNSUPPORTED; }
  const size_t need = bdbnn_wgrad_tc_workspace_bytes(s);
  if (workspace_bytes < need) {
    set_error("binconv_wgrad_tc: workspace %zu B < required %zu B", workspace_bytes, need);
    return BDBNN_ERR_WORKSPACE;
  }
  cudaStream_t st = cudaStream_t(stream);
  pl.p.ws = static_cast<float*>(workspace);
  pl.p.trace = get_tc_trace();
  CUtensorMap tmX, tmG;
  if (pl.p.halo)
    rc = make_act_map(&tmX, xb_bf16, s->N, s->H, s->W, s->Cin, 64, pl.p.PW, pl.p.PH, 1, 1);
  else
    rc = make_act_map(&tmX, xb_bf16, s->N, s->H, s->W, s->Cin, pl.p.UW, pl.p.BW, pl.p.BH, pl.p.BNI, s->stride);
  if (rc) return rc;
  rc = make_act_map(&tmG, gys_bf16, s->N, s->Ho, s->Wo, s->Cout * grad_halves, pl.p.BN < 64 ? pl.p.BN : 64, pl.p.BW,
                    pl.p.BH, pl.p.BNI);
  if (rc) return rc;
  BDBNN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
  dim3 grid(unsigned(pl.ksplit), unsigned(pl.mgroups), unsigned(pl.ntiles));
  tc_wgrad_kernel<<<grid, kTcThreads, pl.smem, st>>>(tmX, tmG, pl.p);
  rc = check_launch("tc_wgrad_kernel");
  if (rc) return rc;
  const int64_t n = int64_t(s->Cout) * s->Cin * s->kh * s->kw;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  wgrad_finalize_kernel<<<unsigned(blocks), 256, 0, st>>>(
      pl.p.ws, pl.ksplit, wmask_bits, inv_gscale, grad_mode == BDBNN_GRAD_FP16S ? amax_bits : nullptr, gW, s->Cout, s->Cin,
      s->kh * s->kw);
  return check_launch("wgrad_finalize_kernel");
}
