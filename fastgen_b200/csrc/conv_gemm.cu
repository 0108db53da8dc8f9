// Implicit-GEMM convolution / linear layer for sm_100a:
//   TMA (tile mode, 128B swizzle) -> shared memory ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
//   -> tcgen05.ld epilogue with the fused bias / per-sample vector / activation / gate / residual(+resample) / scale
//   tail -> 128B-swizzled shared-memory slab -> TMA store.
//
// GEMM view: M = N*H*W output pixels (128 per tile = one TMA box of BW x BH x BNI pixels),
//            N = cout (BLOCK_N per tile), K = sum over segments of taps * cin (64 per pipeline stage).
// A 3x3 tap is the same TMA box shifted by (dx, dy); TMA's out-of-bounds zero fill IS the padding.
//
// Warp roles (320 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quarter = warp_idx % 4; the two warps of a quarter -- "twins" -- split the columns of
// every output slab: with one epilogue warp per scheduler the tail of an epilogue-bound layer was dependency-latency bound,
// profiles/conv_epilogue_r02.md).  Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// Epilogue: a ring of four 16 KB slabs (128 rows x 128 B).  The residual of a slab is prefetched into the ring by
// TMA two slabs ahead (FG_RES_SAME: into the slab itself, added in place; FG_RES_UP2: the quarter-size source box),
// the 128 epilogue threads add their accumulator row and the slab leaves through a TMA store: every global access
// of the tail is a full 128-byte line.  The hot loop is specialised at compile time on (output type, residual
// kind); measured on B200 the runtime-generic version of the same loop was 3x slower (instruction issue bound).
//
// Replaces: fastgen/networks/EDM/network.py:93-126 (Conv2d.forward), :47-51 (Linear.forward), the elementwise tail
// of UNetBlock.forward (:283, :286-287, :297-298); and, for the transformer networks, the Linear + bias + GELU /
// gate * x + residual tails of the Wan / DiT blocks (fastgen/networks/Wan/network.py:116-151, DiT/network.py:184-201).
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kConvThreads = 320;   // producer warp + MMA warp + 8 epilogue warps (two per TMEM lane quarter)

// epilogue specialisations
constexpr int kEpiGeneric = -1;  // no TMA store: direct (vector or scalar) global accesses, every feature at run time
constexpr int kResNone = 0;      // no residual, or one the TMA path cannot take (read with plain loads)
constexpr int kResSame = 1;
constexpr int kResUp2 = 2;
__host__ __device__ constexpr int epi_mode(bool out_bf16, int res) { return (out_bf16 ? 1 : 0) | (res << 1); }
// FG_ACT_GEGLU: bf16 output of half the accumulator width, out[j] = v[2j] * gelu_erf(v[2j+1]), no residual
constexpr int kEpiGeglu = 8 | 1;

// halo tiles (HALO): a tile is 16 rows x 8 pixels of one image; ONE TMA box of (16 + 2) x (8 + 2) pixel rows per 64-channel
// K chunk serves all nine taps of a 3x3 segment (tap (dy, dx) = the same buffer read from row (1 + dy) * 10 + (1 + dx) with
// a 10-row stride between the 8-pixel tile rows), so A traffic per tap drops from 16 KB to 2.5 KB
constexpr int kHaloW = 8, kHaloH = 16;
constexpr int kHaloRows = (kHaloH + 2) * (kHaloW + 2);       // 180 pixel rows of 128 bytes
constexpr int kHaloTxBytes = kHaloRows * 128;                // bytes one halo box load delivers
constexpr int kHaloSbo = (kHaloW + 2) * 128;                 // descriptor stride between 8-pixel tile rows

template <int BLOCK_N, int CG = 1, bool HALO = false>   // CG = 2: CTA pair (cta_group::2), each CTA stages half of the B tile
struct ConvCfg {
  static constexpr int A_BYTES = kBlockM * kBlockK * 2;
  static constexpr int B_BYTES = (BLOCK_N / CG) * kBlockK * 2;
  static constexpr int HALO_BYTES = ((kHaloTxBytes + 1023) / 1024) * 1024;
  static constexpr int HALO_BUFS = 2;
  static constexpr int A_REGION = HALO ? HALO_BUFS * HALO_BYTES : 0;      // halo ring in front of the B ring
  static constexpr int STAGE_BYTES = HALO ? B_BYTES : A_BYTES + B_BYTES;  // HALO: a stage holds one tap's B tile
  static constexpr int OUT_SLAB_BYTES = kBlockM * 128;  // epilogue staging slab: 128 rows x 128 B (TMA box)
  static constexpr int OUT_SLABS = 4;                   // ring: residual prefetch (TMA load) -> in-place add -> TMA store
  static constexpr int RES_UP_BYTES = OUT_SLAB_BYTES / 4;  // FG_RES_UP2: the (BW/2 x BH/2) source box of a slab
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_RAW = (SMEM_MAX - 1024 - BAR_BYTES - OUT_SLABS * OUT_SLAB_BYTES - A_REGION) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > (HALO ? 10 : 8) ? (HALO ? 10 : 8) : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static constexpr int SMEM_BYTES =
      A_REGION + STAGES * STAGE_BYTES + OUT_SLABS * OUT_SLAB_BYTES + 1024 /*align slack*/ + BAR_BYTES /*barriers*/;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024B alignment for the 128B swizzle");
};

#ifdef FG_CONV_TIMING
// Experiment build only (profiles/experiments/conv_epilogue_timing.py): SM clocks per epilogue phase of CTA 0, summed over
// its slabs, for the store thread (warp 2, lane 0) and for a plain epilogue thread (warp 7, lane 0) --
// [who][wait accumulator, store-thread prologue, residual wait, TMEM read, tail values, smem writes + fence, barrier,
//  store issue, statistics, slabs].
__device__ unsigned long long g_conv_timing[2][12];
#define FG_CT_DECL long long ct_prev = clock64(); unsigned long long ct_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define FG_CT(i) do { const long long ct_now = clock64(); ct_acc[i] += ct_now - ct_prev; ct_prev = ct_now; } while (0)
#define FG_XDBG(bit) (p.dbg & (bit))   // FG_CONV_DBG bits of the experiment build: 16 no statistics atomics, 32 no shuffles, 64 no column sums, 128 no bias / row-vector loads
#else
#define FG_CT_DECL
#define FG_CT(i)
#define FG_XDBG(bit) false
#endif

struct ConvKParams {
  CUtensorMap tmA[FG_MAX_SEG];
  CUtensorMap tmW[FG_MAX_SEG];
  CUtensorMap tmOut;  // output tile store: box (32 fp32 | 64 bf16 channels, BW, BH, BNI)
  CUtensorMap tmRes;  // residual prefetch: same box (FG_RES_SAME) or (.., BW/2, BH/2, BNI) of the half-size tensor (UP2)
  int32_t seg_taps[FG_MAX_SEG];
  int32_t seg_kchunks[FG_MAX_SEG];
  int32_t nseg;
  int32_t N, H, W, Cout;
  int32_t BW, BH, BNI;  // TMA box in pixels, BW*BH*BNI == 128
  int32_t tiles_w, tiles_h, tiles_img, tiles_n, total_tiles;
  uint32_t magic_tiles_n, magic_tiles_w, magic_tiles_h;   // fast_div multipliers of decode_tile (0: divide)
  int32_t k_iters;
  int32_t vec;  // generic epilogue: every access may use 4-element vectors
  int32_t dbg;  // FG_CONV_DBG bit mask (profiling experiments only; results are wrong when non-zero)
  const float* bias;
  const float* rowvec;
  int64_t rowvec_ld;
  int32_t act;
  const float* gate;
  int64_t gate_ld;
  int32_t gate_rows;
  const void* res;
  int32_t res_dtype, res_mode;
  int64_t res_ld;
  float out_scale;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
  // fused GroupNorm statistics of the output (fp32 TMA-store path): per (image, group) sum / sum of squares
  double* stats;
  int32_t stats_groups, stats_cg, stats_ch0;
  double* stats2;  // second consumer of the same output (the decoder GroupNorm that reads it inside a skip concatenation)
  int32_t stats2_groups, stats2_cg, stats2_ch0;
  // group index of a channel without an integer division: g = umulhi(ch0 + ch, magic), magic = ceil(2^32 / cg) (exact for
  // channel numbers below 2^16); span = lanes a group can cover inside a warp (bounds the segmented shuffle reduction)
  uint32_t stats_magic, stats2_magic;
  int32_t stats_span, stats2_span;
  // lean slab loop of the TMA-store epilogues: no activation / gate / plain-load residual / ragged last slab, so a slab is
  // TMEM read -> + bias (requested one slab ahead) [+ row vector] [+ staged residual] -> scale -> shared memory
  int32_t lean;
  uint32_t magic_gate_rows;   // fast_div multiplier of gate_rows (lean path; 0: divide)
};

struct TileCoord {
  int w0, h0, img0, co0;
};

// x / d through one multiplication: magic = ceil(2^32 / d), exact while x * d < 2^32 (the host passes 0 -- plain division
// -- when a launch cannot guarantee that, and for d == 1)
__device__ __forceinline__ int fast_div(int x, int d, uint32_t magic) {
  return magic ? static_cast<int>(__umulhi(static_cast<uint32_t>(x), magic)) : d == 1 ? x : x / d;
}

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile, int block_n) {
  TileCoord t;
  int m_tile = fast_div(tile, p.tiles_n, p.magic_tiles_n);
  const int n_tile = tile - m_tile * p.tiles_n;
  t.co0 = n_tile * block_n;
  int q = fast_div(m_tile, p.tiles_w, p.magic_tiles_w);
  t.w0 = (m_tile - q * p.tiles_w) * p.BW;
  m_tile = q;
  q = fast_div(m_tile, p.tiles_h, p.magic_tiles_h);
  t.h0 = (m_tile - q * p.tiles_h) * p.BH;
  t.img0 = q * p.BNI;
  return t;
}

__device__ __forceinline__ float4 ld4(const void* base, int dtype, int64_t idx) {
  if (dtype == FG_F32) {
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx));
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + idx));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}
__device__ __forceinline__ float ld1(const void* base, int dtype, int64_t idx) {
  if (dtype == FG_F32) return __ldg(static_cast<const float*>(base) + idx);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx]);
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf GELU through one MUFU: 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with the odd polynomial fitted (minimax over
// |x| <= 8, argument clamped there: tanh has saturated) to atanh(erf(x / sqrt 2)).  Max |error| against the exact
// 0.5 x (1 + erf(x / sqrt 2)) is 2.5e-5 (the textbook tanh GELU: 4.7e-4), 80x below the bf16 rounding of the result;
// erff() in this tail made the GEGLU GEMM epilogue-bound (834 vs 1314 TFLOP/s, profiles/sdxl_r01.md).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  const float u = xc * fmaf(x2, fmaf(x2, -0.000351519786667389f, 0.037005665847143085f), 0.7975078607596384f);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_fast(u), hx);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == FG_ACT_GELU_TANH) {
    const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    return 0.5f * x * (1.0f + tanh_fast(u));
  }
  if (act == FG_ACT_SILU) return silu_f(x);
  return x;
}

// residual source of one output pixel, for plain-load access
struct ResAddr {
  int64_t rpix, step_w, step_h;
};
__device__ __forceinline__ ResAddr res_addr(const ConvKParams& p, int img, int h, int w) {
  ResAddr r;
  r.rpix = (static_cast<int64_t>(img) * p.H + h) * p.W + w;
  r.step_w = 0, r.step_h = 0;
  if (p.res_mode == FG_RES_UP2) {
    r.rpix = (static_cast<int64_t>(img) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
  } else if (p.res_mode == FG_RES_DOWN2) {
    const int64_t W2 = static_cast<int64_t>(p.W) * 2;
    r.rpix = (static_cast<int64_t>(img) * (p.H * 2) + 2 * h) * W2 + 2 * w;
    r.step_w = p.res_ld;
    r.step_h = W2 * p.res_ld;
  }
  return r;
}
__device__ __forceinline__ float4 res_load4(const ConvKParams& p, const ResAddr& ra, int c) {
  const int64_t o = ra.rpix * p.res_ld + c;
  if (p.res_mode != FG_RES_DOWN2) return ld4(p.res, p.res_dtype, o);
  const float4 b0 = ld4(p.res, p.res_dtype, o), b1 = ld4(p.res, p.res_dtype, o + ra.step_w);
  const float4 b2 = ld4(p.res, p.res_dtype, o + ra.step_h), b3 = ld4(p.res, p.res_dtype, o + ra.step_h + ra.step_w);
  return make_float4(0.25f * ((b0.x + b1.x) + (b2.x + b3.x)), 0.25f * ((b0.y + b1.y) + (b2.y + b3.y)),
                     0.25f * ((b0.z + b1.z) + (b2.z + b3.z)), 0.25f * ((b0.w + b1.w) + (b2.w + b3.w)));
}
__device__ __forceinline__ float res_load1(const ConvKParams& p, const ResAddr& ra, int c) {
  const int64_t o = ra.rpix * p.res_ld + c;
  if (p.res_mode != FG_RES_DOWN2) return ld1(p.res, p.res_dtype, o);
  return 0.25f * ((ld1(p.res, p.res_dtype, o) + ld1(p.res, p.res_dtype, o + ra.step_w)) +
                  (ld1(p.res, p.res_dtype, o + ra.step_h) + ld1(p.res, p.res_dtype, o + ra.step_h + ra.step_w)));
}

// Generic tail for CH accumulator columns of one output pixel, written straight to global memory.
//   v = acc + bias + rowvec;  v = act(v);  v *= gate;  v += residual;  v *= out_scale
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const ConvKParams& p, const uint32_t (&v)[CH], int c_base, int img,
                                               int h, int w) {
  const int64_t pix = (static_cast<int64_t>(img) * p.H + h) * p.W + w;
  const ResAddr ra = res_addr(p, img, h, w);
  const float* grow = p.gate ? p.gate + (pix / p.gate_rows) * p.gate_ld : nullptr;
#pragma unroll
  for (int j = 0; j < CH; j += 4) {
    const int c = c_base + j;
    if (p.vec) {
      if (c < p.Cout) {
        float4 a = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
        if (p.bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        if (p.rowvec) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowvec + img * p.rowvec_ld + c));
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        if (p.act) {
          a.x = apply_act(a.x, p.act), a.y = apply_act(a.y, p.act);
          a.z = apply_act(a.z, p.act), a.w = apply_act(a.w, p.act);
        }
        if (grow) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(grow + c));
          a.x *= g.x, a.y *= g.y, a.z *= g.z, a.w *= g.w;
        }
        if (p.res_mode != FG_RES_NONE) {
          const float4 b = res_load4(p, ra, c);
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        a.x *= p.out_scale, a.y *= p.out_scale, a.z *= p.out_scale, a.w *= p.out_scale;
        if (p.out_dtype == FG_F32) {
          *reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * p.out_ld + c) = a;
        } else {
          uint2 u;
          u.x = pack_bf16x2(a.x, a.y);
          u.y = pack_bf16x2(a.z, a.w);
          *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + c) = u;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cc = c + i;
        if (cc < p.Cout) {
          float a = __uint_as_float(v[j + i]);
          if (p.bias) a += __ldg(p.bias + cc);
          if (p.rowvec) a += __ldg(p.rowvec + img * p.rowvec_ld + cc);
          a = apply_act(a, p.act);
          if (grow) a *= __ldg(grow + cc);
          if (p.res_mode != FG_RES_NONE) a += res_load1(p, ra, cc);
          a *= p.out_scale;
          if (p.out_dtype == FG_F32) {
            static_cast<float*>(p.out)[pix * p.out_ld + cc] = a;
          } else {
            static_cast<__nv_bfloat16*>(p.out)[pix * p.out_ld + cc] = __float2bfloat16_rn(a);
          }
        }
      }
    }
  }
}

// Specialised tail of the TMA-store path for 32 accumulator columns: everything except the residual that arrives
// through shared memory and the final scale.  Branches are warp-uniform; the rare plain-load residual (FG_RES_DOWN2,
// or a residual whose type differs from the output's) stays behind one of them.
template <int RES, int CH = 32>
__device__ __forceinline__ void tail_values(const ConvKParams& p, const uint32_t (&v)[CH], int c_base, int img, int h,
                                            int w, bool valid, float (&a)[CH]) {
#pragma unroll
  for (int j = 0; j < CH; ++j) a[j] = __uint_as_float(v[j]);
  if (p.bias && !FG_XDBG(128)) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      if (c_base + j < p.Cout) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c_base + j));
        a[j] += b.x, a[j + 1] += b.y, a[j + 2] += b.z, a[j + 3] += b.w;
      }
    }
  }
  if (p.rowvec && valid && !FG_XDBG(128)) {
    const float* rv = p.rowvec + img * p.rowvec_ld + c_base;
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      if (c_base + j < p.Cout) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(rv + j));
        a[j] += b.x, a[j + 1] += b.y, a[j + 2] += b.z, a[j + 3] += b.w;
      }
    }
  }
  if (p.act == FG_ACT_GELU_TANH) {  // (the switch is hoisted: one warp-uniform branch per 32 columns)
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const float x = a[j];
      const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
      const float hx = 0.5f * x;
      a[j] = fmaf(hx, tanh_fast(u), hx);
    }
  } else if (p.act == FG_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < CH; ++j) a[j] = silu_f(a[j]);
  }
  if (p.gate && valid) {
    const int64_t pix = (static_cast<int64_t>(img) * p.H + h) * p.W + w;
    const float* g = p.gate + (pix / p.gate_rows) * p.gate_ld + c_base;
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      if (c_base + j < p.Cout) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(g + j));
        a[j] *= b.x, a[j + 1] *= b.y, a[j + 2] *= b.z, a[j + 3] *= b.w;
      }
    }
  }
  if constexpr (RES == kResNone) {
    if (p.res_mode != FG_RES_NONE && valid) {
      const ResAddr ra = res_addr(p, img, h, w);
#pragma unroll
      for (int j = 0; j < CH; j += 4) {  // (unrolled: a dynamic index would push a[] into local memory)
        if (c_base + j < p.Cout) {
          const float4 b = res_load4(p, ra, c_base + j);
          a[j] += b.x, a[j + 1] += b.y, a[j + 2] += b.z, a[j + 3] += b.w;
        }
      }
    }
  }
}

// CG = 2: a cluster of two CTAs owns two consecutive M tiles of one N tile.  Work items ("items") are tiles for CG = 1 and
// tile pairs for CG = 2; item -> this CTA's tile index goes through item_tile().  An odd M-tile count leaves the second
// CTA of the last pair a tile past the end: its TMA boxes are entirely out of bounds (zero-filled loads, clipped stores),
// its statistics and plain accesses are masked by `img < N`.
template <int CG>
__device__ __forceinline__ int item_tile(const ConvKParams& p, int item, int cta_rank) {
  if constexpr (CG == 1) return item;
  const int mp = fast_div(item, p.tiles_n, p.magic_tiles_n);
  return (2 * mp + cta_rank) * p.tiles_n + (item - mp * p.tiles_n);
}

template <int BLOCK_N, int EPI, int CG = 1, bool HALO = false>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BLOCK_N, CG, HALO>;
  static_assert(CG == 1 || (EPI != kEpiGeneric && BLOCK_N >= 128), "CTA pairs: TMA-store epilogues, tiles >= 128 wide");
  static_assert(!HALO || CG == 2, "halo tiles are built for the CTA-pair kernels");
  const int cr = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;   // rank in the pair; 0 = leader (issues the MMAs)
  const int item0 = static_cast<int>(blockIdx.x) / CG, item_step = static_cast<int>(gridDim.x) / CG;
  constexpr bool kTmaStore = EPI != kEpiGeneric;
  constexpr bool kOutBf16 = kTmaStore && (EPI & 1);
  constexpr int kRes = kTmaStore ? ((EPI >> 1) & 3) : kResNone;
  constexpr bool kGeglu = kTmaStore && (EPI & 8);
  static_assert(!kGeglu || (kOutBf16 && kRes == kResNone && BLOCK_N % 128 == 0), "GEGLU tail: bf16, no residual");
  static_assert(!kTmaStore || BLOCK_N >= 64, "the TMA-store epilogue needs tiles of at least one bf16 slab");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + Cfg::A_REGION;     // HALO: [halo buffers][B ring]; else [A | B] stages
  uint8_t* out_slabs = ring + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_slabs + Cfg::OUT_SLABS * Cfg::OUT_SLAB_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_full = tempty + 2;  // [OUT_SLABS]
  uint64_t* a_full = res_full + Cfg::OUT_SLABS;   // [HALO_BUFS] (HALO only)
  uint64_t* a_empty = a_full + Cfg::HALO_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + Cfg::HALO_BUFS);
  static_assert((2 * Cfg::STAGES + 4 + Cfg::OUT_SLABS + 2 * Cfg::HALO_BUFS) * 8 + 4 <= Cfg::BAR_BYTES, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmW[s]);
    }
    if (kTmaStore) tma_prefetch_desc(&p.tmOut);
    if (kRes != kResNone) tma_prefetch_desc(&p.tmRes);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 8 * CG);   // 8 epilogue warps; CG = 2: the leader's barrier collects those of both CTAs
      }
      for (int i = 0; i < Cfg::OUT_SLABS; ++i) mbar_init(&res_full[i], 1);
      for (int i = 0; i < Cfg::HALO_BUFS; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CG == 2) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);   // collective over the pair: the same warp of both CTAs
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers exist before anything signals them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (HALO) {
        int ab = 0;
        uint32_t aphase = 0;
        for (int item = item0; item < p.total_tiles; item += item_step) {
          const TileCoord t = decode_tile(p, item_tile<CG>(p, item, cr), BLOCK_N);
          for (int s = 0; s < p.nseg; ++s) {
            const int taps = p.seg_taps[s];
            const int kch = p.seg_kchunks[s];
            for (int kc = 0; kc < kch; ++kc) {
              // one halo box per 64-channel chunk (a 1x1 segment reads only its centre); both CTAs' boxes are counted on
              // the leader's barrier
              mbar_wait(&a_empty[ab], aphase ^ 1u);
              // 3x3 segment: the (16 + 2) x (8 + 2) halo box; 1x1 segment (a shortcut projection riding along): the plain
              // 16 x 8 tile box of its tensor map, read back with the standard 8-row stride
              const int halo = taps == 9 ? 1 : 0;
              if (cr == 0) mbar_arrive_expect_tx(&a_full[ab], 2 * (halo ? kHaloTxBytes : Cfg::A_BYTES));
              tma_load_4d_2sm(smem + ab * Cfg::HALO_BYTES, &p.tmA[s], mapa_u32(&a_full[ab], 0), kc * kBlockK, t.w0 - halo,
                              t.h0 - halo, t.img0);
              ab ^= 1;
              if (ab == 0) aphase ^= 1u;
              for (int tap = 0; tap < taps; ++tap) {
                mbar_wait(&empty[stage], phase ^ 1u);
                if (cr == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::B_BYTES);
                tma_load_3d_2sm(ring + stage * Cfg::STAGE_BYTES, &p.tmW[s], mapa_u32(&full[stage], 0), kc * kBlockK, tap,
                                t.co0 + cr * (BLOCK_N / 2));
                if (++stage == Cfg::STAGES) {
                  stage = 0;
                  phase ^= 1u;
                }
              }
            }
          }
        }
      } else
      for (int item = item0; item < p.total_tiles && !(p.dbg & 4); item += item_step) {
        const TileCoord t = decode_tile(p, item_tile<CG>(p, item, cr), BLOCK_N);
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int kch = p.seg_kchunks[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            for (int kc = 0; kc < kch; ++kc) {
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
              uint8_t* sb = sa + Cfg::A_BYTES;
              if constexpr (CG == 2) {
                // both CTAs' bytes are counted on the leader's barrier; this CTA stages its A rows and its half of B
                if (cr == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                const uint32_t lead_full = mapa_u32(&full[stage], 0);
                tma_load_4d_2sm(sa, &p.tmA[s], lead_full, kc * kBlockK, t.w0 + dx, t.h0 + dy, t.img0);
                tma_load_3d_2sm(sb, &p.tmW[s], lead_full, kc * kBlockK, tap, t.co0 + cr * (BLOCK_N / 2));
              } else {
                mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                tma_load_4d(sa, &p.tmA[s], &full[stage], kc * kBlockK, t.w0 + dx, t.h0 + dy, t.img0);
                tma_load_3d(sb, &p.tmW[s], &full[stage], kc * kBlockK, tap, t.co0);
              }
              if (++stage == Cfg::STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp walks the loop (converged: descriptors stay in uniform registers, no per-MMA broadcast loop -- see
    // umma_bf16_ss_elect), one elected lane issues
    if (cr == 0) {
      const uint32_t elected = elect_one_sync();
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      [[maybe_unused]] int hab = 0;
      [[maybe_unused]] uint32_t haphase = 0;
      for (int item = item0; item < p.total_tiles; item += item_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        if constexpr (HALO) {
          uint32_t started = 0;
          for (int s = 0; s < p.nseg; ++s) {
            const int taps = p.seg_taps[s];
            const int kch = p.seg_kchunks[s];
            for (int kc = 0; kc < kch; ++kc) {
              mbar_wait(&a_full[hab], haphase);
              tc_fence_after();
              const uint32_t halo_addr = smem_u32(smem + hab * Cfg::HALO_BYTES);
              for (int tap = 0; tap < taps; ++tap) {
                // tap (dy, dx): the tile's pixel (y, x) sits at halo row (y + 1 + dy) * 10 + (x + 1 + dx)
                const uint32_t row0 = taps == 9 ? static_cast<uint32_t>((tap / 3) * (kHaloW + 2) + tap % 3) : 0u;
                const uint32_t sbo = taps == 9 ? kHaloSbo : 1024u;
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(ring + stage * Cfg::STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16_ss2_elect(elected, d_tmem, umma_desc_sw128_sbo(halo_addr + row0 * 128 + k * 32, sbo),
                                umma_desc_sw128(b_addr + k * 32), idesc, (started | k) != 0 ? 1u : 0u);
                }
                started = 1;
                umma_commit2_elect(elected, &empty[stage], 3);
                if (++stage == Cfg::STAGES) {
                  stage = 0;
                  phase ^= 1u;
                }
              }
              umma_commit2_elect(elected, &a_empty[hab], 3);   // the halo buffer is free once the last tap's MMAs retire (both CTAs)
              hab ^= 1;
              if (hab == 0) haphase ^= 1u;
            }
          }
        } else
        for (int it = 0; it < p.k_iters && !(p.dbg & 4); ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if constexpr (CG == 2)
              umma_bf16_ss2_elect(elected, d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                            (it | k) != 0 ? 1u : 0u);
            else
              umma_bf16_ss_elect(elected, d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                           (it | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot once these MMAs retire (CG = 2: in both CTAs -- each producer waits on its own barrier)
          if constexpr (CG == 2) umma_commit2_elect(elected, &empty[stage], 3);
          else umma_commit_elect(elected, &empty[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if constexpr (CG == 2) umma_commit2_elect(elected, &tfull[acc], 3);  // accumulator ready for the epilogue warps of both CTAs
        else umma_commit_elect(elected, &tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int twin = (warp - 2) >> 2;   // which half of every slab's columns this warp handles
    const int row = q * 32 + lane;
    const int wl = row % p.BW;
    const int hl = (row / p.BW) % p.BH;
    const int il = row / (p.BW * p.BH);
    int acc = 0;
    uint32_t acc_phase = 0;
    // slab ring bookkeeping (slab_idx counts the slabs of this CTA across all its tiles)
    constexpr uint32_t kResAhead = 2;  // residual slabs in flight ahead of the one being processed
    constexpr int kSlabCols = kGeglu ? 128 : kOutBf16 ? 64 : 32;  // accumulator columns behind one 128-byte slab row
    // FG_RES_UP2 prefetch: three output slabs, the fourth buffer holds the four quarter-size residual boxes
    constexpr uint32_t kOutRing = (kRes == kResUp2) ? Cfg::OUT_SLABS - 1 : Cfg::OUT_SLABS;
    const bool is_store_thread = (warp == 2 && lane == 0);
    uint8_t* res_up = out_slabs + (Cfg::OUT_SLABS - 1) * Cfg::OUT_SLAB_BYTES;
    uint32_t slab_idx = 0;
    uint32_t ld_idx = 0;  // store thread only: next residual slab to request, its tile and channel offset
    int ld_item = item0, ld_c0 = 0;
    TileCoord lt = {0, 0, 0, 0};   // decoded tile of ld_item
    int lt_item = -1;
    const uint32_t wkey = row & 7;
    // residual row inside a quarter-size UP2 box
    const int rs_up = (il * (p.BH >> 1) + (hl >> 1)) * (p.BW >> 1) + (wl >> 1);

    const uint32_t lead_tempty[2] = {CG == 2 ? mapa_u32(&tempty[0], 0) : 0u, CG == 2 ? mapa_u32(&tempty[1], 0) : 0u};
    const int stat_img = ((row >> 5) * 32) / (p.BW * p.BH);   // image (within the tile) of this thread's statistics rows
    constexpr bool kCanLean = kTmaStore && !kGeglu && kRes != kResUp2;
    constexpr int kLeanCols = kOutBf16 ? 32 : 16;             // accumulator columns per thread and slab
    const bool lean = kCanLean && p.lean != 0;
    float nb[kLeanCols];                                       // lean: bias of the next slab, requested one slab ahead
    auto lean_bias = [&](int cbase) {
#pragma unroll
      for (int j = 0; j < kLeanCols; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
        nb[j] = b.x, nb[j + 1] = b.y, nb[j + 2] = b.z, nb[j + 3] = b.w;
      }
    };
#pragma unroll
    for (int j = 0; j < kLeanCols; ++j) nb[j] = 0.f;
    FG_CT_DECL;
    for (int item = item0; item < p.total_tiles; item += item_step) {
      const TileCoord t = decode_tile(p, item_tile<CG>(p, item, cr), BLOCK_N);
      const int img = t.img0 + il, h = t.h0 + hl, w = t.w0 + wl;
      const bool valid = img < p.N;
      const float* rv_row = nullptr;   // lean: this row's slice of the per-image row vector (rows past the last image
      if (lean) {                      // read the last image's: the TMA store clips them)
        if (p.bias) lean_bias(t.co0 + twin * kLeanCols);
        if (p.rowvec) rv_row = p.rowvec + static_cast<int64_t>(valid ? img : p.N - 1) * p.rowvec_ld;
        // ... or of the multiplicative AdaLN gate (one gate row per gate_rows consecutive pixels); a launch has one or the other
        if (p.gate) {
          const int pix = valid ? (img * p.H + h) * p.W + w : 0;
          rv_row = p.gate + static_cast<int64_t>(fast_div(pix, p.gate_rows, p.magic_gate_rows)) * p.gate_ld;
        }
      }
      FG_CT(11);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      FG_CT(0);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      if (p.dbg & 8) {
      } else if constexpr (kTmaStore) {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += kSlabCols, ++slab_idx) {
          if (t.co0 + c0 >= p.Cout) break;  // warp-uniform: nothing left to store in this tile
          const uint32_t ob = slab_idx % kOutRing;        // output slab buffer
          const uint32_t rb = slab_idx % Cfg::OUT_SLABS;  // residual buffer / barrier
          uint8_t* slab = out_slabs + ob * Cfg::OUT_SLAB_BYTES;
          if (is_store_thread) {
            // every store but the most recent one has finished reading shared memory: the buffer that slab
            // slab_idx + kResAhead lands in (last stored from two slabs ago) is free again
            tma_store_wait_read<1>();
            if constexpr (kRes != kResNone) {
              while (ld_idx <= slab_idx + kResAhead && ld_item < p.total_tiles) {
                if (ld_item != lt_item) {   // (the divisions of decode_tile once per tile, not once per slab)
                  lt = decode_tile(p, item_tile<CG>(p, ld_item, cr), BLOCK_N);
                  lt_item = ld_item;
                }
                const uint32_t lrb = ld_idx % Cfg::OUT_SLABS;
                if constexpr (kRes == kResSame) {
                  mbar_arrive_expect_tx(&res_full[lrb], Cfg::OUT_SLAB_BYTES);
                  tma_load_4d(out_slabs + lrb * Cfg::OUT_SLAB_BYTES, &p.tmRes, &res_full[lrb], lt.co0 + ld_c0, lt.w0,
                              lt.h0, lt.img0);
                } else {
                  mbar_arrive_expect_tx(&res_full[lrb], Cfg::RES_UP_BYTES);
                  tma_load_4d(res_up + lrb * Cfg::RES_UP_BYTES, &p.tmRes, &res_full[lrb], lt.co0 + ld_c0, lt.w0 >> 1,
                              lt.h0 >> 1, lt.img0);
                }
                ++ld_idx;
                ld_c0 += kSlabCols;
                if (ld_c0 >= BLOCK_N || lt.co0 + ld_c0 >= p.Cout) {
                  ld_c0 = 0;
                  ld_item += item_step;
                }
              }
            }
          }
          FG_CT(1);
          const uint8_t* rrow = nullptr;
          uint32_t rkey = 0;
          if constexpr (kRes == kResSame) {
            mbar_wait(&res_full[rb], (slab_idx / Cfg::OUT_SLABS) & 1u);
            rrow = slab + row * 128;
            rkey = wkey;
          } else if constexpr (kRes == kResUp2) {
            mbar_wait(&res_full[rb], (slab_idx / Cfg::OUT_SLABS) & 1u);
            rrow = res_up + rb * Cfg::RES_UP_BYTES + rs_up * 128;
            rkey = rs_up & 7;
          }
          uint8_t* my_row = slab + row * 128;
          const float sc = p.out_scale;
          FG_CT(2);
          if (lean) {
            if constexpr (kCanLean) {
              const int cb = t.co0 + c0 + twin * kLeanCols;   // first of this thread's columns
              const uint32_t my_row_s = smem_u32(my_row);
              const uint32_t rrow_s = kRes == kResSame ? smem_u32(rrow) : 0u;
              float rv[kLeanCols];
              if (rv_row) {
#pragma unroll
                for (int j = 0; j < kLeanCols; j += 4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(rv_row + cb + j));
                  rv[j] = b.x, rv[j + 1] = b.y, rv[j + 2] = b.z, rv[j + 3] = b.w;
                }
              }
              uint32_t v[kLeanCols];
              if constexpr (kOutBf16) tmem_ld_32x32(taddr + c0 + twin * 32, v);
              else tmem_ld_32x16(taddr + c0 + twin * 16, v);
              tmem_ld_wait();
              FG_CT(3);
              float a[kLeanCols];
#pragma unroll
              for (int j = 0; j < kLeanCols; ++j) a[j] = __uint_as_float(v[j]) + nb[j];
              if (rv_row && !p.gate) {
#pragma unroll
                for (int j = 0; j < kLeanCols; ++j) a[j] += rv[j];
              }
              if (p.act == FG_ACT_GELU_TANH) {   // (warp-uniform; same arithmetic as tail_values)
#pragma unroll
                for (int j = 0; j < kLeanCols; ++j) {
                  const float x = a[j];
                  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
                  const float hx = 0.5f * x;
                  a[j] = fmaf(hx, tanh_fast(u), hx);
                }
              } else if (p.act == FG_ACT_SILU) {
#pragma unroll
                for (int j = 0; j < kLeanCols; ++j) a[j] = silu_f(a[j]);
              }
              if (rv_row && p.gate) {
#pragma unroll
                for (int j = 0; j < kLeanCols; ++j) a[j] *= rv[j];
              }
              // the next slab's bias: in flight during the stores, the barrier and the statistics of this one
              if (p.bias && c0 + kSlabCols < BLOCK_N && t.co0 + c0 + kSlabCols < p.Cout) lean_bias(cb + kSlabCols);
              FG_CT(4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t chunk = static_cast<uint32_t>(twin * 4 + j);   // 16-byte chunk of the 128-byte slab row
                if constexpr (!kOutBf16) {
                  float4 o = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                  if constexpr (kRes == kResSame) {
                    const float4 r = lds_v4f(rrow_s + ((chunk ^ rkey) << 4));
                    o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
                  }
                  o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
                  sts_v4f(my_row_s + ((chunk ^ wkey) << 4), o);
                } else {
                  if constexpr (kRes == kResSame) {
                    const uint4 r = lds_v4u(rrow_s + ((chunk ^ rkey) << 4));
                    const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y), r2 = unpack_bf16x2(r.z),
                                 r3 = unpack_bf16x2(r.w);
                    a[8 * j] += r0.x, a[8 * j + 1] += r0.y, a[8 * j + 2] += r1.x, a[8 * j + 3] += r1.y;
                    a[8 * j + 4] += r2.x, a[8 * j + 5] += r2.y, a[8 * j + 6] += r3.x, a[8 * j + 7] += r3.y;
                  }
                  uint4 u;
                  u.x = pack_bf16x2(a[8 * j] * sc, a[8 * j + 1] * sc), u.y = pack_bf16x2(a[8 * j + 2] * sc, a[8 * j + 3] * sc);
                  u.z = pack_bf16x2(a[8 * j + 4] * sc, a[8 * j + 5] * sc);
                  u.w = pack_bf16x2(a[8 * j + 6] * sc, a[8 * j + 7] * sc);
                  sts_v4u(my_row_s + ((chunk ^ wkey) << 4), u);
                }
              }
            }
          } else if constexpr (!kOutBf16 && !kGeglu) {
            // fp32 slab = 32 accumulator columns: each twin takes 16 of them (four 16-byte chunks of the slab row)
            uint32_t v[16];
            float a[16];
            tmem_ld_32x16(taddr + c0 + twin * 16, v);
            tmem_ld_wait();
            FG_CT(3);
            tail_values<kRes, 16>(p, v, t.co0 + c0 + twin * 16, img, h, w, valid, a);
            FG_CT(4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int chunk = twin * 4 + j;
              float4 o = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
              if constexpr (kRes != kResNone) {
                const float4 r = *reinterpret_cast<const float4*>(rrow + ((chunk ^ rkey) << 4));
                o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
              }
              o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
              *reinterpret_cast<float4*>(my_row + ((chunk ^ wkey) << 4)) = o;
            }
          }
#pragma unroll
          for (int cc = twin * (kSlabCols / 2); cc < (twin + 1) * (kSlabCols / 2) && (kOutBf16 || kGeglu) && !lean; cc += 32) {
            uint32_t v[32];
            float a[32];
            if (!(p.dbg & 2)) {
              tmem_ld_32x32(taddr + c0 + cc, v);
              tmem_ld_wait();
              FG_CT(3);
              tail_values<kRes>(p, v, t.co0 + c0 + cc, img, h, w, valid, a);
              FG_CT(4);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) a[j] = 1.0f;
            }
            if constexpr (kGeglu) {
              // columns (2j, 2j+1) = (hidden_j, gate_j): 32 accumulator columns -> 16 outputs = two 16-byte chunks
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float hid = a[16 * j + 2 * k], g = a[16 * j + 2 * k + 1];
                  o[k] = hid * gelu_erf_fast(g);
                }
                const int chunk = (cc >> 4) + j;
                uint4 u;
                u.x = pack_bf16x2(o[0] * sc, o[1] * sc), u.y = pack_bf16x2(o[2] * sc, o[3] * sc);
                u.z = pack_bf16x2(o[4] * sc, o[5] * sc), u.w = pack_bf16x2(o[6] * sc, o[7] * sc);
                *reinterpret_cast<uint4*>(my_row + ((chunk ^ wkey) << 4)) = u;
              }
            } else if constexpr (!kOutBf16) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                if constexpr (kRes != kResNone) {
                  const float4 r = *reinterpret_cast<const float4*>(rrow + ((j ^ rkey) << 4));
                  o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
                }
                o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
                *reinterpret_cast<float4*>(my_row + ((j ^ wkey) << 4)) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int chunk = (cc >> 3) + j;
                if constexpr (kRes != kResNone) {
                  const uint4 r = *reinterpret_cast<const uint4*>(rrow + ((chunk ^ rkey) << 4));
                  const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y), r2 = unpack_bf16x2(r.z),
                               r3 = unpack_bf16x2(r.w);
                  a[8 * j] += r0.x, a[8 * j + 1] += r0.y, a[8 * j + 2] += r1.x, a[8 * j + 3] += r1.y;
                  a[8 * j + 4] += r2.x, a[8 * j + 5] += r2.y, a[8 * j + 6] += r3.x, a[8 * j + 7] += r3.y;
                }
                uint4 u;
                u.x = pack_bf16x2(a[8 * j] * sc, a[8 * j + 1] * sc), u.y = pack_bf16x2(a[8 * j + 2] * sc, a[8 * j + 3] * sc);
                u.z = pack_bf16x2(a[8 * j + 4] * sc, a[8 * j + 5] * sc);
                u.w = pack_bf16x2(a[8 * j + 6] * sc, a[8 * j + 7] * sc);
                *reinterpret_cast<uint4*>(my_row + ((chunk ^ wkey) << 4)) = u;
              }
            }
          }
          if (!FG_XDBG(256)) fence_proxy_async_smem();
          FG_CT(5);
          named_bar_sync(1, 256);
          FG_CT(6);
          if (is_store_thread && !(p.dbg & 1)) {
            tma_store_4d(&p.tmOut, slab, kGeglu ? ((t.co0 + c0) >> 1) : (t.co0 + c0), t.w0, t.h0, t.img0);
            tma_store_commit();
          }
          FG_CT(7);
          if constexpr (!kGeglu) {
            if (p.stats) {
              // GroupNorm statistics of the finished slab (the next consumer's fg_gn_stats pass disappears): thread
              // (row block rb, column, twin) sums 16 rows of one column out of shared memory (the twins halve the
              // dependent load chain), runs of lanes that fall into the same group are combined with a segmented
              // shuffle reduction, one fp64 atomic pair per run and twin.  A bf16 slab holds 64 columns: a thread then
              // owns a PAIR of adjacent columns (one 4-byte word per row; the host checks that a pair never straddles
              // two groups) and sums the rounded values, i.e. exactly what the consumer's fg_gn_apply will read.
              const int col = row & 31, rb = row >> 5;
              const int ch = t.co0 + c0 + col * (kOutBf16 ? 2 : 1);
              const int simg = t.img0 + stat_img;
              float s1 = 0.f, s2 = 0.f;
              if (ch < p.Cout && simg < p.N && !FG_XDBG(64)) {
                const uint32_t base = smem_u32(slab) + (rb * 32 + twin * 16) * 128 + (col & 3) * 4;
                const uint32_t cchunk = static_cast<uint32_t>(col >> 2);
                float t1 = 0.f, t2 = 0.f;   // two independent chains
#pragma unroll
                for (int r = 0; r < 16; r += 2) {   // (twin * 16 + r) & 7 == r & 7: the swizzle key of the row
                  if constexpr (kOutBf16) {
                    const float2 v = unpack_bf16x2(lds_u(base + r * 128 + ((cchunk ^ (r & 7)) << 4)));
                    const float2 u = unpack_bf16x2(lds_u(base + (r + 1) * 128 + ((cchunk ^ ((r + 1) & 7)) << 4)));
                    s1 += v.x + v.y, t1 += u.x + u.y;
                    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2)), t2 = fmaf(u.x, u.x, fmaf(u.y, u.y, t2));
                  } else {
                    const float v = lds_f(base + r * 128 + ((cchunk ^ (r & 7)) << 4));
                    const float u = lds_f(base + (r + 1) * 128 + ((cchunk ^ ((r + 1) & 7)) << 4));
                    s1 += v, t1 += u;
                    s2 = fmaf(v, v, s2), t2 = fmaf(u, u, t2);
                  }
                }
                s1 += t1, s2 += t2;
              }
              // the column sums feed up to two group layouts: this tensor's own next norm and, for a tensor that is also
              // a skip connection, the decoder norm that reads it at channel offset stats2_ch0 of a concatenation
              auto reduce_groups = [&](double* base, int groups, uint32_t magic, int span, int ch0) {
                float r1 = s1, r2 = s2;
                const int g = (ch < p.Cout) ? static_cast<int>(__umulhi(static_cast<uint32_t>(ch0 + ch), magic))
                                            : -1 - col;  // invalid lanes: singleton runs
#pragma unroll 1
                for (int off = 1; off < span && !FG_XDBG(32); off <<= 1) {   // lanes >= span apart never share a group
                  const float o1 = __shfl_down_sync(0xffffffffu, r1, off);
                  const float o2 = __shfl_down_sync(0xffffffffu, r2, off);
                  const int og = __shfl_down_sync(0xffffffffu, g, off);
                  if (lane + off < 32 && og == g) r1 += o1, r2 += o2;
                }
                const int pg = __shfl_up_sync(0xffffffffu, g, 1);
                if ((lane == 0 || pg != g) && g >= 0 && simg < p.N && !FG_XDBG(16)) {
                  double* dst = base + (static_cast<int64_t>(simg) * groups + g) * 2;
                  atomicAdd(dst, static_cast<double>(r1));
                  atomicAdd(dst + 1, static_cast<double>(r2));
                }
              };
              reduce_groups(p.stats, p.stats_groups, p.stats_magic, p.stats_span, p.stats_ch0);
              if (p.stats2) reduce_groups(p.stats2, p.stats2_groups, p.stats2_magic, p.stats2_span, p.stats2_ch0);
            }
          }
          FG_CT(8);
#ifdef FG_CONV_TIMING
          ++ct_acc[9];
#endif
        }
      } else if constexpr (BLOCK_N >= 32) {
#pragma unroll 1
        for (int c0 = twin * 32; c0 < BLOCK_N; c0 += 64) {   // the twins alternate 32-column chunks
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          if (valid && t.co0 + c0 < p.Cout) epilogue_chunk<32>(p, v, t.co0 + c0, img, h, w);
        }
      } else {
        if (twin == 0) {
          uint32_t v[16];
          tmem_ld_32x16(taddr, v);
          tmem_ld_wait();
          if (valid) epilogue_chunk<16>(p, v, t.co0, img, h, w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(lead_tempty[acc]);
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (kTmaStore && is_store_thread) tma_store_wait<0>();
#ifdef FG_CONV_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 7)) {
      for (int i = 0; i < 12; ++i) g_conv_timing[warp == 2 ? 0 : 1][i] = ct_acc[i];
    }
#endif
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while the pair is still in use
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N, int EPI, int CG = 1, bool HALO = false>
static int launch_conv(const ConvKParams& kp, int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, CG, HALO>;
  static PerDeviceFlag configured;
  const int cfg_dev = current_device();
  if (!configured.test(cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, EPI, CG, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_gemm: cudaFuncSetAttribute(%d B smem) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
    configured.set(cfg_dev);
  }
  if constexpr (CG == 2) {
    // kp.total_tiles counts tile PAIRS: one cluster of two CTAs (the two SMs of a TPC) per pair, persistent over pairs
    const int clusters = kp.total_tiles < num_sms / 2 ? kp.total_tiles : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(kConvThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, EPI, CG, HALO>, kp);
    if (e != cudaSuccess) {
      set_error("conv_gemm: cluster launch failed: %s", cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
  } else {
    const int grid = kp.total_tiles < num_sms ? kp.total_tiles : num_sms;
    conv_gemm_kernel<BLOCK_N, EPI, CG, HALO><<<grid, kConvThreads, Cfg::SMEM_BYTES, stream>>>(kp);
  }
  FG_CHECK_LAUNCH("conv_gemm");
  return FG_OK;
}

template <int BLOCK_N, int CG = 1, bool HALO = false>
static int dispatch_epi(const ConvKParams& kp, int epi, int num_sms, cudaStream_t stream) {
  if constexpr (BLOCK_N >= 64) {
    switch (epi) {
      case epi_mode(false, kResNone): return launch_conv<BLOCK_N, epi_mode(false, kResNone), CG, HALO>(kp, num_sms, stream);
      case epi_mode(true, kResNone): return launch_conv<BLOCK_N, epi_mode(true, kResNone), CG, HALO>(kp, num_sms, stream);
      case epi_mode(false, kResSame): return launch_conv<BLOCK_N, epi_mode(false, kResSame), CG, HALO>(kp, num_sms, stream);
      case epi_mode(true, kResSame): return launch_conv<BLOCK_N, epi_mode(true, kResSame), CG, HALO>(kp, num_sms, stream);
      case epi_mode(false, kResUp2): return launch_conv<BLOCK_N, epi_mode(false, kResUp2), CG, HALO>(kp, num_sms, stream);
      case epi_mode(true, kResUp2): return launch_conv<BLOCK_N, epi_mode(true, kResUp2), CG, HALO>(kp, num_sms, stream);
      case kEpiGeglu:
        if constexpr (BLOCK_N % 128 == 0) return launch_conv<BLOCK_N, kEpiGeglu, CG, HALO>(kp, num_sms, stream);
        break;
      default: break;
    }
  }
  if constexpr (CG == 2) {
    set_error("conv_gemm: the CTA-pair kernels need a TMA-store epilogue");
    return FG_ERR_UNSUPPORTED;
  } else {
    return launch_conv<BLOCK_N, kEpiGeneric>(kp, num_sms, stream);
  }
}

// Tile choice.  Cost of a candidate (tile width bn, CTAs per MMA cg) = waves of work items over the SMs (pairs of SMs)
// x clocks per 64-deep K stage, the larger of
//   * the tensor pipe: 2 bn clocks (an M = 128 (x cg) by bn by 16 MMA takes bn / 8 clocks per SM), and
//   * the L2 -> SM port of 64 B/clk/SM feeding the stage: (16 KB of A + bn / cg * 128 B of B) / 64.
// Measured on B200 (profiles/conv_2cta_r02.md): one-CTA tiles run at 50 % (bn = 128) and 67-69 % (bn = 256) of the tensor
// pipe exactly as the port predicts; a CTA pair halves the B bytes per CTA (the peer's half is read over the pair link),
// so a 256-wide pair tile is no longer port-bound.  Ties go to the wider tile / the pair.
struct TileChoice {
  int bn, cg;
};
static TileChoice pick_tile(int cout, int m_tiles, int k_iters, int num_sms, bool geglu, bool pair_ok, bool halo_ok) {
  static const int forced = getenv("FG_CONV_BN") ? atoi(getenv("FG_CONV_BN")) : 0;  // tuning experiments only
  static const int forced_cg = getenv("FG_CONV_CG") ? atoi(getenv("FG_CONV_CG")) : 0;
  if (!geglu) {
    if (cout <= 16) return {16, 1};
    if (cout <= 64) return {64, 1};
  }
  // a short K loop is epilogue-bound, and the accumulator hand-back of a pair couples the epilogues of two SMs: measured
  // 586 vs 711 TFLOP/s (K = 384, N = 1152) and 471 vs 523 (K = 384, N = 384) for pair vs single CTA, against 1291 vs 1236
  // at K = 3456 and 1182 vs 1130 at K = 1728 (profiles/conv_2cta_r02.md)
  static const int min_k_iters = getenv("FG_CONV_CG_MIN_K") ? atoi(getenv("FG_CONV_CG_MIN_K")) : 16;
  if (forced_cg == 1 || m_tiles < 2 || (forced_cg != 2 && k_iters < min_k_iters)) pair_ok = false;
  TileChoice best = {128, 1};
  int64_t best_cost = INT64_MAX;
  const int cands[3] = {128, 192, 256};
  for (int cg = 1; cg <= (pair_ok ? 2 : 1); ++cg) {
    if (forced_cg == 2 && pair_ok && cg == 1) continue;
    for (int i = 0; i < 3; ++i) {
      const int bn = cands[i];
      if (geglu && bn % 128) continue;  // a GEGLU slab is 128 accumulator columns
      if (forced && bn != forced && cout > 64 && !(geglu && forced % 128)) continue;
      if (cout <= 128 && bn > 128) continue;
      const int n_tiles = (cout + bn - 1) / bn;
      const int64_t items = static_cast<int64_t>((m_tiles + cg - 1) / cg) * n_tiles;
      const int64_t slots = num_sms / cg;
      const int64_t waves = (items + slots - 1) / slots;
      const int mma_clk = 2 * bn;
      // halo tiles (pairs only): 23 KB of A per nine tap stages instead of 16 KB per stage
      const int a_bytes = (halo_ok && cg == 2) ? kHaloTxBytes / 9 + 512 : 16384;
      const int port_clk = (a_bytes + bn / cg * 128) / 64;
      // + the fixed cost of a pipeline stage (barrier wait, descriptor set-up, commit): a 256-clock stage of a 128-wide
      // tile pays it on every 64 channels -- measured 740-880 vs 1030-1150 TFLOP/s when the model first picked 128-wide
      // pair tiles for the N = 384 layers
      const int stage_clk = (mma_clk > port_clk ? mma_clk : port_clk) + 64;
      // + the part of the per-tile tail that the next tile's main loop cannot hide (two accumulator stages overlap them:
      // only a K loop shorter than the epilogue shows it)
      const int64_t epi_clk = 12 * bn;
      const int64_t main_clk = static_cast<int64_t>(k_iters) * stage_clk;
      const int64_t cost = waves * (main_clk > epi_clk ? main_clk : epi_clk) + main_clk / 8;
      if (cost < best_cost || (cost == best_cost && (bn > best.bn || cg > best.cg))) {
        best_cost = cost;
        best = {bn, cg};
      }
    }
  }
  return best;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fg

namespace fg {
// What fg_conv_gemm_prepare leaves in the caller's blob: the kernel parameter block (tile geometry, the 2 nseg + 2 encoded
// TMA descriptors, epilogue pointers) and the kernel choice.  fg_conv_gemm_run only launches it.
struct alignas(64) ConvPrepared {
  uint64_t magic;
  int32_t block_n, cg, halo, epi, num_sms, device;
  ConvKParams kp;
};
constexpr uint64_t kConvPreparedMagic = 0x46474350524550ull;  // "FGCPREP"

static int conv_launch_prepared(const ConvPrepared& pp, cudaStream_t stream) {
  if (pp.cg == 2 && pp.halo) {
    switch (pp.block_n) {
      case 128: return dispatch_epi<128, 2, true>(pp.kp, pp.epi, pp.num_sms, stream);
      case 192: return dispatch_epi<192, 2, true>(pp.kp, pp.epi, pp.num_sms, stream);
      case 256: return dispatch_epi<256, 2, true>(pp.kp, pp.epi, pp.num_sms, stream);
    }
    set_error("conv_gemm: no halo kernel for block_n=%d", pp.block_n);
    return FG_ERR_UNSUPPORTED;
  }
  if (pp.cg == 2) {
    switch (pp.block_n) {
      case 128: return dispatch_epi<128, 2>(pp.kp, pp.epi, pp.num_sms, stream);
      case 192: return dispatch_epi<192, 2>(pp.kp, pp.epi, pp.num_sms, stream);
      case 256: return dispatch_epi<256, 2>(pp.kp, pp.epi, pp.num_sms, stream);
    }
    set_error("conv_gemm: no CTA-pair kernel for block_n=%d", pp.block_n);
    return FG_ERR_UNSUPPORTED;
  }
  switch (pp.block_n) {
    case 16: return dispatch_epi<16>(pp.kp, pp.epi, pp.num_sms, stream);
    case 64: return dispatch_epi<64>(pp.kp, pp.epi, pp.num_sms, stream);
    case 128: return dispatch_epi<128>(pp.kp, pp.epi, pp.num_sms, stream);
    case 192: return dispatch_epi<192>(pp.kp, pp.epi, pp.num_sms, stream);
    case 256: return dispatch_epi<256>(pp.kp, pp.epi, pp.num_sms, stream);
  }
  set_error("conv_gemm: no kernel for block_n=%d", pp.block_n);
  return FG_ERR_UNSUPPORTED;
}
static int conv_prepare(const fg_conv_params* p, ConvPrepared& pp);
}  // namespace fg

extern "C" int64_t fg_conv_gemm_prepared_size(void) { return static_cast<int64_t>(sizeof(fg::ConvPrepared)); }

extern "C" int fg_conv_gemm_prepare(const fg_conv_params* p, void* blob, int64_t blob_bytes) {
  using namespace fg;
  FG_CHECK_ARG(blob != nullptr && blob_bytes >= static_cast<int64_t>(sizeof(ConvPrepared)) &&
                   (reinterpret_cast<uintptr_t>(blob) & 63u) == 0,
               "conv_gemm_prepare: the blob must be 64-byte aligned and hold fg_conv_gemm_prepared_size() = %lld bytes",
               (long long)sizeof(ConvPrepared));
  ConvPrepared* pp = static_cast<ConvPrepared*>(blob);
  pp->magic = 0;
  const int r = conv_prepare(p, *pp);
  if (r == FG_OK) pp->magic = kConvPreparedMagic;
  return r;
}

extern "C" int fg_conv_gemm_run(const void* blob, fg_stream_t stream_) {
  using namespace fg;
  const ConvPrepared* pp = static_cast<const ConvPrepared*>(blob);
  FG_CHECK_ARG(pp != nullptr && (reinterpret_cast<uintptr_t>(blob) & 63u) == 0 && pp->magic == kConvPreparedMagic,
               "conv_gemm_run: not a blob filled by fg_conv_gemm_prepare");
  return conv_launch_prepared(*pp, static_cast<cudaStream_t>(stream_));
}

extern "C" int fg_conv_gemm(const fg_conv_params* p, fg_stream_t stream_) {
  using namespace fg;
  ConvPrepared pp;
  const int r = conv_prepare(p, pp);
  if (r != FG_OK) return r;
  return conv_launch_prepared(pp, static_cast<cudaStream_t>(stream_));
}

static int fg::conv_prepare(const fg_conv_params* p, ConvPrepared& pp) {
  FG_CHECK_ARG(p != nullptr, "conv_gemm: null params");
  FG_CHECK_ARG(p->n > 0 && p->h > 0 && p->w > 0 && p->cout > 0, "conv_gemm: empty problem n=%d h=%d w=%d cout=%d",
               p->n, p->h, p->w, p->cout);
  FG_CHECK_ARG(p->nseg >= 1 && p->nseg <= FG_MAX_SEG, "conv_gemm: nseg=%d out of range", p->nseg);
  FG_CHECK_ARG(p->out != nullptr, "conv_gemm: null output");
  FG_CHECK_ARG(p->out_dtype == FG_F32 || p->out_dtype == FG_BF16, "conv_gemm: bad out_dtype %d", p->out_dtype);
  FG_CHECK_ARG(p->act >= FG_ACT_NONE && p->act <= FG_ACT_GEGLU, "conv_gemm: bad act %d", p->act);
  const bool geglu = p->act == FG_ACT_GEGLU;
  const int out_cols = geglu ? p->cout / 2 : p->cout;  // GEGLU writes one column per (hidden, gate) column pair
  FG_CHECK_ARG(p->out_ld >= out_cols, "conv_gemm: out_ld %lld < %d output columns", (long long)p->out_ld, out_cols);
  if (geglu)
    FG_CHECK_ARG(p->cout % 16 == 0 && p->out_dtype == FG_BF16 && !p->res && !p->gate && !p->stats,
                 "conv_gemm: FG_ACT_GEGLU needs cout %% 16 == 0, a bf16 output, no residual / gate / statistics");
  if (p->gate)
    FG_CHECK_ARG(p->gate_rows > 0 && p->gate_ld >= p->cout, "conv_gemm: gate needs gate_rows > 0 and gate_ld >= cout");

  ConvKParams& kp = pp.kp;
  memset(&kp, 0, sizeof(kp));
  kp.N = p->n, kp.H = p->h, kp.W = p->w, kp.Cout = p->cout;
  {
    static const int dbg = getenv("FG_CONV_DBG") ? atoi(getenv("FG_CONV_DBG")) : 0;
    kp.dbg = dbg;
  }
  if (p->w >= 128) {
    FG_CHECK_ARG(p->w % 128 == 0, "conv_gemm: w=%d must be a power of two or a multiple of 128", p->w);
    kp.BW = 128, kp.BH = 1, kp.BNI = 1;
  } else {
    FG_CHECK_ARG(is_pow2(p->w), "conv_gemm: w=%d must be a power of two or a multiple of 128", p->w);
    kp.BW = p->w;
    const int rows = 128 / p->w;
    if (p->h >= rows) {
      FG_CHECK_ARG(p->h % rows == 0, "conv_gemm: h=%d not a multiple of the %d-row tile", p->h, rows);
      kp.BH = rows, kp.BNI = 1;
    } else {
      FG_CHECK_ARG(is_pow2(p->h), "conv_gemm: h=%d must be a power of two", p->h);
      kp.BH = p->h, kp.BNI = rows / p->h;
    }
  }
  kp.tiles_w = p->w / kp.BW;
  kp.tiles_h = p->h / kp.BH;
  kp.tiles_img = (p->n + kp.BNI - 1) / kp.BNI;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();  // no device (CPU-only argument check)
  const int num_sms = device_sm_count(dev);
  const int res_mode_eff = p->res ? p->res_mode : FG_RES_NONE;
  kp.vec = (p->cout % 4 == 0) && (p->out_ld % 4 == 0) && aligned16(p->out) && (!p->bias || aligned16(p->bias)) &&
           (!p->rowvec || (aligned16(p->rowvec) && p->rowvec_ld % 4 == 0)) &&
           (!p->gate || (aligned16(p->gate) && p->gate_ld % 4 == 0)) &&
           (res_mode_eff == FG_RES_NONE || (aligned16(p->res) && p->res_ld % 4 == 0));
  static const bool allow_tma_store = !(getenv("FG_CONV_TMA_STORE") && getenv("FG_CONV_TMA_STORE")[0] == '0');
  // TMA-store epilogue whenever the output rows are 16-byte addressable and the tile is at least one slab wide
  const bool tma_store_ok = allow_tma_store && kp.vec && (p->out_ld * (p->out_dtype == FG_F32 ? 4 : 2)) % 16 == 0;
  const int m_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_img;
  int k_iters_total = 0;
  for (int s = 0; s < p->nseg; ++s)
    if (p->seg[s].cin > 0) k_iters_total += p->seg[s].taps * ((p->seg[s].cin + kBlockK - 1) / kBlockK);
  // halo tiles: every segment 3x3 or 1x1, at least one 3x3, the image tiles into 16 x 8 pixel blocks
  static const bool allow_halo = !(getenv("FG_CONV_HALO") && getenv("FG_CONV_HALO")[0] == '0');
  bool halo_ok = allow_halo && tma_store_ok && p->w % kHaloW == 0 && p->h % kHaloH == 0;
  {
    bool any9 = false;
    for (int s = 0; s < p->nseg; ++s) any9 = any9 || p->seg[s].taps == 9;
    halo_ok = halo_ok && any9;
  }
  const TileChoice tc = pick_tile(p->cout, m_tiles, k_iters_total, num_sms, geglu, /*pair_ok=*/tma_store_ok, halo_ok);
  const int block_n = tc.bn, cg = tc.cg;
  const bool halo = halo_ok && cg == 2;
  if (halo) {   // same tile count (128 pixels each), different shape: 16 rows x 8 pixels of one image
    kp.BW = kHaloW, kp.BH = kHaloH, kp.BNI = 1;
    kp.tiles_w = p->w / kp.BW, kp.tiles_h = p->h / kp.BH, kp.tiles_img = p->n;
  }
  kp.tiles_n = (p->cout + block_n - 1) / block_n;
  kp.total_tiles = ((m_tiles + cg - 1) / cg) * kp.tiles_n;   // work items: tiles, or tile pairs of a CTA pair
  {
    // decode_tile without integer divisions (largest dividend: a tile index of the padded tile grid)
    const int64_t max_tile = static_cast<int64_t>(m_tiles + cg) * kp.tiles_n;
    auto magic_of = [&](int d) -> uint32_t {
      if (d <= 1 || max_tile * d >= (1ll << 32)) return 0u;
      return static_cast<uint32_t>(((1ull << 32) + d - 1) / d);
    };
    kp.magic_tiles_n = magic_of(kp.tiles_n), kp.magic_tiles_w = magic_of(kp.tiles_w), kp.magic_tiles_h = magic_of(kp.tiles_h);
  }

  kp.nseg = p->nseg;
  kp.k_iters = 0;
  for (int s = 0; s < p->nseg; ++s) {
    const fg_conv_seg& sg = p->seg[s];
    FG_CHECK_ARG(sg.a && sg.w, "conv_gemm: segment %d has a null pointer", s);
    FG_CHECK_ARG(sg.taps == 1 || sg.taps == 9, "conv_gemm: segment %d taps=%d (1 or 9)", s, sg.taps);
    FG_CHECK_ARG(sg.cin > 0 && sg.cin % 8 == 0, "conv_gemm: segment %d cin=%d must be a positive multiple of 8", s,
                 sg.cin);
    FG_CHECK_ARG(sg.a_ld >= sg.cin && sg.a_ld % 8 == 0, "conv_gemm: segment %d a_ld=%lld (>= cin, %% 8 == 0)", s,
                 (long long)sg.a_ld);
    FG_CHECK_ARG(aligned16(sg.a) && aligned16(sg.w), "conv_gemm: segment %d pointers must be 16-byte aligned", s);
    kp.seg_taps[s] = sg.taps;
    kp.seg_kchunks[s] = (sg.cin + kBlockK - 1) / kBlockK;
    kp.k_iters += sg.taps * kp.seg_kchunks[s];
    {
      const uint64_t dims[4] = {(uint64_t)sg.cin, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
      const uint64_t ldb = (uint64_t)sg.a_ld * 2;
      const uint64_t strides[4] = {2, ldb, ldb * p->w, ldb * p->w * p->h};
      const bool hbox = halo && sg.taps == 9;   // (a 1x1 segment of a halo launch keeps the plain tile box)
      const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(hbox ? kHaloW + 2 : kp.BW), (uint32_t)(hbox ? kHaloH + 2 : kp.BH),
                               (uint32_t)kp.BNI};
      int r = encode_tmap_bf16(&kp.tmA[s], sg.a, 4, dims, strides, box);
      if (r != FG_OK) return r;
    }
    {
      const uint64_t dims[3] = {(uint64_t)sg.cin, (uint64_t)sg.taps, (uint64_t)p->cout};
      const uint64_t strides[3] = {2, (uint64_t)sg.cin * 2, (uint64_t)sg.cin * sg.taps * 2};
      const uint32_t box[3] = {(uint32_t)kBlockK, 1u, (uint32_t)(block_n / cg)};   // a CTA of a pair stages half the tile
      int r = encode_tmap_bf16(&kp.tmW[s], sg.w, 3, dims, strides, box);
      if (r != FG_OK) return r;
    }
  }

  kp.bias = p->bias;
  kp.rowvec = p->rowvec;
  kp.rowvec_ld = p->rowvec_ld;
  kp.act = p->act;
  kp.gate = p->gate;
  kp.gate_ld = p->gate_ld;
  kp.gate_rows = p->gate ? p->gate_rows : 1;
  kp.res = p->res;
  kp.res_dtype = p->res_dtype;
  kp.res_mode = p->res ? p->res_mode : FG_RES_NONE;
  kp.res_ld = p->res_ld;
  kp.out_scale = p->out_scale;
  kp.out = p->out;
  kp.out_dtype = p->out_dtype;
  kp.out_ld = p->out_ld;
  FG_CHECK_ARG(!p->stats2 || p->stats, "conv_gemm: stats2 without stats (use stats for a single consumer)");
  if (p->stats) {
    FG_CHECK_ARG(p->stats_groups > 0 && p->stats_cg > 0 && p->stats_ch0 >= 0,
                 "conv_gemm: fused GroupNorm statistics need stats_groups > 0, stats_cg > 0, stats_ch0 >= 0");
    // bf16 output: the epilogue sums pairs of adjacent columns, which must fall into one group
    FG_CHECK_ARG(p->out_dtype == FG_F32 || (p->act != FG_ACT_GEGLU && p->stats_cg % 2 == 0 && p->stats_ch0 % 2 == 0 &&
                                            (!p->stats2 || (p->stats2_cg % 2 == 0 && p->stats2_ch0 % 2 == 0))),
                 "conv_gemm: fused GroupNorm statistics of a bf16 output need even group widths and channel offsets");
    FG_CHECK_ARG(kp.BW * kp.BH >= 32, "conv_gemm: fused GroupNorm statistics need images of at least 32 pixels");
    kp.stats = p->stats, kp.stats_groups = p->stats_groups, kp.stats_cg = p->stats_cg, kp.stats_ch0 = p->stats_ch0;
    // group index by multiplication (see ConvKParams::stats_magic) and the lanes a group spans in the epilogue's warps
    auto magic_of = [](int cg) { return static_cast<uint32_t>(((1ull << 32) + cg - 1) / cg); };
    auto span_of = [&](int cg) { const int l = p->out_dtype == FG_F32 ? cg : cg / 2; return l > 32 ? 32 : l; };
    FG_CHECK_ARG(p->stats_cg >= 2 && (!p->stats2 || p->stats2_cg >= 2), "conv_gemm: fused GroupNorm statistics need groups of at least 2 channels");
    FG_CHECK_ARG(static_cast<int64_t>(p->stats_ch0 + p->cout) * p->stats_cg < (1ll << 32) &&
                     (!p->stats2 || static_cast<int64_t>(p->stats2_ch0 + p->cout) * p->stats2_cg < (1ll << 32)),
                 "conv_gemm: fused GroupNorm statistics: channel range too large");
    kp.stats_magic = magic_of(p->stats_cg), kp.stats_span = span_of(p->stats_cg);
    if (p->stats2) {
      FG_CHECK_ARG(p->stats2_groups > 0 && p->stats2_cg > 0 && p->stats2_ch0 >= 0,
                   "conv_gemm: stats2 needs stats2_groups > 0, stats2_cg > 0, stats2_ch0 >= 0");
      kp.stats2 = p->stats2, kp.stats2_groups = p->stats2_groups, kp.stats2_cg = p->stats2_cg, kp.stats2_ch0 = p->stats2_ch0;
      kp.stats2_magic = magic_of(p->stats2_cg), kp.stats2_span = span_of(p->stats2_cg);
    }
  }
  FG_CHECK_ARG(kp.res_mode >= FG_RES_NONE && kp.res_mode <= FG_RES_DOWN2, "conv_gemm: bad res_mode %d", kp.res_mode);
  if (kp.res_mode != FG_RES_NONE) {
    FG_CHECK_ARG(p->res_dtype == FG_F32 || p->res_dtype == FG_BF16, "conv_gemm: bad res_dtype %d", p->res_dtype);
    FG_CHECK_ARG(p->res_ld >= p->cout, "conv_gemm: res_ld %lld < cout", (long long)p->res_ld);
    if (kp.res_mode == FG_RES_UP2)
      FG_CHECK_ARG(p->h % 2 == 0 && p->w % 2 == 0, "conv_gemm: FG_RES_UP2 needs even h,w");
  }
  // the residual rides the slab ring of the TMA-store epilogue when it has the output's element type (FG_RES_SAME / UP2)
  int epi = kEpiGeneric;
  {
    const int esz = p->out_dtype == FG_F32 ? 4 : 2;
    if (tma_store_ok && block_n >= 64) {
      const uint64_t dims[4] = {(uint64_t)out_cols, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
      const uint64_t ldb = (uint64_t)p->out_ld * esz;
      const uint64_t strides[4] = {(uint64_t)esz, ldb, ldb * p->w, ldb * p->w * p->h};
      const uint32_t box[4] = {(uint32_t)(128 / esz), (uint32_t)kp.BW, (uint32_t)kp.BH, (uint32_t)kp.BNI};
      int r = encode_tmap(&kp.tmOut, p->out, p->out_dtype, 4, dims, strides, box);
      if (r != FG_OK) return r;
      int res_kind = kResNone;
      static const bool allow_tma_res = !(getenv("FG_CONV_TMA_RES") && getenv("FG_CONV_TMA_RES")[0] == '0');
      const bool up = kp.res_mode == FG_RES_UP2;
      if (allow_tma_res && (kp.res_mode == FG_RES_SAME || (up && kp.BW % 2 == 0 && kp.BH % 2 == 0)) &&
          p->res_dtype == p->out_dtype && (p->res_ld * esz) % 16 == 0) {
        const int sh = up ? 1 : 0;
        const uint64_t rdims[4] = {(uint64_t)p->cout, (uint64_t)(p->w >> sh), (uint64_t)(p->h >> sh), (uint64_t)p->n};
        const uint64_t rldb = (uint64_t)p->res_ld * esz;
        const uint64_t rstrides[4] = {(uint64_t)esz, rldb, rldb * (p->w >> sh), rldb * (p->w >> sh) * (p->h >> sh)};
        const uint32_t rbox[4] = {(uint32_t)(128 / esz), (uint32_t)(kp.BW >> sh), (uint32_t)(kp.BH >> sh),
                                  (uint32_t)kp.BNI};
        r = encode_tmap(&kp.tmRes, p->res, p->res_dtype, 4, rdims, rstrides, rbox);
        if (r != FG_OK) return r;
        res_kind = up ? kResUp2 : kResSame;
      }
      epi = geglu ? kEpiGeglu : epi_mode(p->out_dtype == FG_BF16, res_kind);
      static const bool allow_lean = !(getenv("FG_CONV_LEAN") && getenv("FG_CONV_LEAN")[0] == '0');
      const int64_t pixels = static_cast<int64_t>(p->n) * p->h * p->w;
      kp.lean = allow_lean && !geglu && res_kind != kResUp2 &&
                (kp.res_mode == FG_RES_NONE || res_kind == kResSame) && p->cout % (128 / esz) == 0 &&
                (!p->bias || aligned16(p->bias)) && (!p->rowvec || (aligned16(p->rowvec) && p->rowvec_ld % 4 == 0)) &&
                (!p->gate || (!p->rowvec && aligned16(p->gate) && p->gate_ld % 4 == 0 && p->gate_rows > 0 &&
                              pixels < (1ll << 31)));
      kp.magic_gate_rows = 0;
      if (p->gate && p->gate_rows > 1 && pixels * p->gate_rows < (1ll << 32))
        kp.magic_gate_rows = static_cast<uint32_t>(((1ull << 32) + p->gate_rows - 1) / p->gate_rows);
    }
  }
  if (geglu) FG_CHECK_ARG(epi == kEpiGeglu, "conv_gemm: FG_ACT_GEGLU needs the TMA-store epilogue (16-byte aligned bf16 output rows)");
  if (kp.stats) FG_CHECK_ARG(epi != kEpiGeneric, "conv_gemm: fused GroupNorm statistics need the TMA-store epilogue (cout %% 4 == 0, cout > 16, 16-byte aligned rows)");

  pp.block_n = block_n, pp.cg = cg, pp.halo = halo ? 1 : 0, pp.epi = epi, pp.num_sms = num_sms, pp.device = dev;
  return FG_OK;
}

#ifdef FG_CONV_TIMING
extern "C" int fg_debug_conv_timing(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, fg::g_conv_timing, sizeof(unsigned long long) * 24) == cudaSuccess ? 0 : -1;
}
#endif
