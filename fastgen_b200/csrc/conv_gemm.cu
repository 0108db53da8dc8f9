// Implicit-GEMM convolution / linear layer for sm_100a:
//   TMA (tile mode, 128B swizzle) -> shared memory ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
//   -> tcgen05.ld epilogue with the fused bias / per-sample vector / residual(+resample) / scale tail.
//
// GEMM view: M = N*H*W output pixels (128 per tile = one TMA box of BW x BH x BNI pixels),
//            N = cout (BLOCK_N per tile), K = sum over segments of taps * cin (64 per pipeline stage).
// A 3x3 tap is the same TMA box shifted by (dx, dy); TMA's out-of-bounds zero fill IS the padding.
//
// Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).  Two TMEM accumulator stages let the epilogue of
// tile i overlap the main loop of tile i+1.
//
// Replaces: fastgen/networks/EDM/network.py:93-126 (Conv2d.forward), :47-51 (Linear.forward) and the
// elementwise tail of UNetBlock.forward (:283, :286-287, :297-298).
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kConvThreads = 192;

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int A_BYTES = kBlockM * kBlockK * 2;
  static constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_SLAB_BYTES = kBlockM * 128;  // epilogue staging slab: 128 rows x 128 B (TMA store box)
  static constexpr int OUT_SLABS = 2;
  static constexpr int STAGES_RAW = (196 * 1024 - OUT_SLABS * OUT_SLAB_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + OUT_SLABS * OUT_SLAB_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024B alignment for the 128B swizzle");
};

struct ConvKParams {
  CUtensorMap tmA[FG_MAX_SEG];
  CUtensorMap tmW[FG_MAX_SEG];
  CUtensorMap tmOut;  // output tile store: box (32 fp32 | 64 bf16 channels, BW, BH, BNI)
  int32_t tma_store;  // 1: epilogue stages 128B-row slabs in smem and stores them with TMA
  int32_t seg_taps[FG_MAX_SEG];
  int32_t seg_kchunks[FG_MAX_SEG];
  int32_t nseg;
  int32_t N, H, W, Cout;
  int32_t BW, BH, BNI;  // TMA box in pixels, BW*BH*BNI == 128
  int32_t tiles_w, tiles_h, tiles_img, tiles_n, total_tiles;
  int32_t k_iters;
  int32_t vec;  // 1: every epilogue access may use 4-element vectors
  const float* bias;
  const float* rowvec;
  int64_t rowvec_ld;
  const void* res;
  int32_t res_dtype, res_mode;
  int64_t res_ld;
  float out_scale;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
};

struct TileCoord {
  int w0, h0, img0, co0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile, int block_n) {
  TileCoord t;
  const int n_tile = tile % p.tiles_n;
  int m_tile = tile / p.tiles_n;
  t.co0 = n_tile * block_n;
  t.w0 = (m_tile % p.tiles_w) * p.BW;
  m_tile /= p.tiles_w;
  t.h0 = (m_tile % p.tiles_h) * p.BH;
  t.img0 = (m_tile / p.tiles_h) * p.BNI;
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

// Fused tail for CH accumulator columns of one output pixel.
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const ConvKParams& p, const uint32_t (&v)[CH], int c_base, int img,
                                               int h, int w) {
  const int64_t pix = (static_cast<int64_t>(img) * p.H + h) * p.W + w;
  // residual source pixel(s)
  int64_t rpix = pix;
  int64_t rstep_w = 0, rstep_h = 0;
  if (p.res_mode == FG_RES_UP2) {
    rpix = (static_cast<int64_t>(img) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
  } else if (p.res_mode == FG_RES_DOWN2) {
    const int64_t W2 = static_cast<int64_t>(p.W) * 2;
    rpix = (static_cast<int64_t>(img) * (p.H * 2) + 2 * h) * W2 + 2 * w;
    rstep_w = p.res_ld;
    rstep_h = W2 * p.res_ld;
  }
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
        if (p.res_mode == FG_RES_SAME || p.res_mode == FG_RES_UP2) {
          const float4 b = ld4(p.res, p.res_dtype, rpix * p.res_ld + c);
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        } else if (p.res_mode == FG_RES_DOWN2) {
          const int64_t o = rpix * p.res_ld + c;
          const float4 b0 = ld4(p.res, p.res_dtype, o), b1 = ld4(p.res, p.res_dtype, o + rstep_w);
          const float4 b2 = ld4(p.res, p.res_dtype, o + rstep_h), b3 = ld4(p.res, p.res_dtype, o + rstep_h + rstep_w);
          a.x += 0.25f * ((b0.x + b1.x) + (b2.x + b3.x));
          a.y += 0.25f * ((b0.y + b1.y) + (b2.y + b3.y));
          a.z += 0.25f * ((b0.z + b1.z) + (b2.z + b3.z));
          a.w += 0.25f * ((b0.w + b1.w) + (b2.w + b3.w));
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
          if (p.res_mode == FG_RES_SAME || p.res_mode == FG_RES_UP2) {
            a += ld1(p.res, p.res_dtype, rpix * p.res_ld + cc);
          } else if (p.res_mode == FG_RES_DOWN2) {
            const int64_t o = rpix * p.res_ld + cc;
            a += 0.25f * ((ld1(p.res, p.res_dtype, o) + ld1(p.res, p.res_dtype, o + rstep_w)) +
                          (ld1(p.res, p.res_dtype, o + rstep_h) + ld1(p.res, p.res_dtype, o + rstep_h + rstep_w)));
          }
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

// Fused tail evaluated into registers (vector path only): a[j] = out_scale * (acc + bias + rowvec + residual).
template <int CH>
__device__ __forceinline__ void epilogue_values(const ConvKParams& p, const uint32_t (&v)[CH], int c_base, int img,
                                                int h, int w, bool valid, float (&a)[CH]) {
  const int64_t pix = (static_cast<int64_t>(img) * p.H + h) * p.W + w;
  int64_t rpix = pix;
  int64_t rstep_w = 0, rstep_h = 0;
  if (p.res_mode == FG_RES_UP2) {
    rpix = (static_cast<int64_t>(img) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
  } else if (p.res_mode == FG_RES_DOWN2) {
    const int64_t W2 = static_cast<int64_t>(p.W) * 2;
    rpix = (static_cast<int64_t>(img) * (p.H * 2) + 2 * h) * W2 + 2 * w;
    rstep_w = p.res_ld;
    rstep_h = W2 * p.res_ld;
  }
#pragma unroll
  for (int j = 0; j < CH; j += 4) {
    const int c = c_base + j;
    float4 r = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                           __uint_as_float(v[j + 3]));
    if (valid && c < p.Cout) {
      if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        r.x += b.x, r.y += b.y, r.z += b.z, r.w += b.w;
      }
      if (p.rowvec) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowvec + img * p.rowvec_ld + c));
        r.x += b.x, r.y += b.y, r.z += b.z, r.w += b.w;
      }
      if (p.res_mode == FG_RES_SAME || p.res_mode == FG_RES_UP2) {
        const float4 b = ld4(p.res, p.res_dtype, rpix * p.res_ld + c);
        r.x += b.x, r.y += b.y, r.z += b.z, r.w += b.w;
      } else if (p.res_mode == FG_RES_DOWN2) {
        const int64_t o = rpix * p.res_ld + c;
        const float4 b0 = ld4(p.res, p.res_dtype, o), b1 = ld4(p.res, p.res_dtype, o + rstep_w);
        const float4 b2 = ld4(p.res, p.res_dtype, o + rstep_h), b3 = ld4(p.res, p.res_dtype, o + rstep_h + rstep_w);
        r.x += 0.25f * ((b0.x + b1.x) + (b2.x + b3.x));
        r.y += 0.25f * ((b0.y + b1.y) + (b2.y + b3.y));
        r.z += 0.25f * ((b0.z + b1.z) + (b2.z + b3.z));
        r.w += 0.25f * ((b0.w + b1.w) + (b2.w + b3.w));
      }
    }
    a[j] = r.x * p.out_scale, a[j + 1] = r.y * p.out_scale, a[j + 2] = r.z * p.out_scale, a[j + 3] = r.w * p.out_scale;
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_slabs = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_slabs + Cfg::OUT_SLABS * Cfg::OUT_SLAB_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmW[s]);
    }
    if (p.tma_store) tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile, BLOCK_N);
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int kch = p.seg_kchunks[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            for (int kc = 0; kc < kch; ++kc) {
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              uint8_t* sb = sa + Cfg::A_BYTES;
              mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
              tma_load_4d(sa, &p.tmA[s], &full[stage], kc * kBlockK, t.w0 + dx, t.h0 + dy, t.img0);
              tma_load_3d(sb, &p.tmW[s], &full[stage], kc * kBlockK, tap, t.co0);
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
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                         (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wl = row % p.BW;
    const int hl = (row / p.BW) % p.BH;
    const int il = row / (p.BW * p.BH);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_idx = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile, BLOCK_N);
      const int img = t.img0 + il, h = t.h0 + hl, w = t.w0 + wl;
      const bool valid = img < p.N;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      if (p.tma_store) {
        // ---- staged epilogue: TMEM -> registers -> 128B-swizzled smem slab -> TMA store (full-line writes,
        //      out-of-range rows / channels clipped by the tensor map) ----
        if constexpr (BLOCK_N >= 32) {
          const bool is_store_thread = (warp == 2 && lane == 0);
          const int slab_cols = (p.out_dtype == FG_F32) ? 32 : 64;
          uint8_t* my_row_base = out_slabs + row * 128;
#pragma unroll 1
          for (int c0 = 0; c0 < BLOCK_N; c0 += slab_cols) {
            if (t.co0 + c0 >= p.Cout) break;  // warp-uniform: nothing left to store in this tile
            uint8_t* slab = my_row_base + (slab_idx & 1) * Cfg::OUT_SLAB_BYTES;
            if (is_store_thread) tma_store_wait_read<1>();  // the store that used this slab two slabs ago is done
            named_bar_sync(1, 128);
#pragma unroll 1
            for (int cc = 0; cc < slab_cols; cc += 32) {
              uint32_t v[32];
              float a[32];
              tmem_ld_32x32(taddr + c0 + cc, v);
              tmem_ld_wait();
              epilogue_values<32>(p, v, t.co0 + c0 + cc, img, h, w, valid, a);
              if (p.out_dtype == FG_F32) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(slab + ((j ^ (row & 7)) << 4)) =
                      make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 u;
                  u.x = pack_bf16x2(a[8 * j], a[8 * j + 1]), u.y = pack_bf16x2(a[8 * j + 2], a[8 * j + 3]);
                  u.z = pack_bf16x2(a[8 * j + 4], a[8 * j + 5]), u.w = pack_bf16x2(a[8 * j + 6], a[8 * j + 7]);
                  *reinterpret_cast<uint4*>(slab + ((((cc >> 3) + j) ^ (row & 7)) << 4)) = u;
                }
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (is_store_thread) {
              tma_store_4d(&p.tmOut, out_slabs + (slab_idx & 1) * Cfg::OUT_SLAB_BYTES, t.co0 + c0, t.w0, t.h0, t.img0);
              tma_store_commit();
            }
            ++slab_idx;
          }
        }
      } else if constexpr (BLOCK_N >= 32) {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          if (valid && t.co0 + c0 < p.Cout) epilogue_chunk<32>(p, v, t.co0 + c0, img, h, w);
        }
      } else {
        uint32_t v[16];
        tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        if (valid) epilogue_chunk<16>(p, v, t.co0, img, h, w);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (p.tma_store && warp == 2 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N>
static int launch_conv(const ConvKParams& kp, int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_gemm: cudaFuncSetAttribute(%d B smem) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
    configured = true;
  }
  const int grid = kp.total_tiles < num_sms ? kp.total_tiles : num_sms;
  conv_gemm_kernel<BLOCK_N><<<grid, kConvThreads, Cfg::SMEM_BYTES, stream>>>(kp);
  FG_CHECK_LAUNCH("conv_gemm");
  return FG_OK;
}

static int pick_block_n(int cout) {
  if (cout <= 16) return 16;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  int best = 128, best_cost = 1 << 30;
  const int cands[3] = {128, 192, 256};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int cost = ((cout + bn - 1) / bn) * bn;
    if (cost <= best_cost) {  // ties go to the larger tile (fewer A re-reads)
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fg

extern "C" int fg_conv_gemm(const fg_conv_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p != nullptr, "conv_gemm: null params");
  FG_CHECK_ARG(p->n > 0 && p->h > 0 && p->w > 0 && p->cout > 0, "conv_gemm: empty problem n=%d h=%d w=%d cout=%d",
               p->n, p->h, p->w, p->cout);
  FG_CHECK_ARG(p->nseg >= 1 && p->nseg <= FG_MAX_SEG, "conv_gemm: nseg=%d out of range", p->nseg);
  FG_CHECK_ARG(p->out != nullptr, "conv_gemm: null output");
  FG_CHECK_ARG(p->out_dtype == FG_F32 || p->out_dtype == FG_BF16, "conv_gemm: bad out_dtype %d", p->out_dtype);
  FG_CHECK_ARG(p->out_ld >= p->cout, "conv_gemm: out_ld %lld < cout %d", (long long)p->out_ld, p->cout);

  ConvKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.N = p->n, kp.H = p->h, kp.W = p->w, kp.Cout = p->cout;
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
  const int block_n = pick_block_n(p->cout);
  kp.tiles_n = (p->cout + block_n - 1) / block_n;
  kp.total_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_img * kp.tiles_n;

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
      const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kp.BW, (uint32_t)kp.BH, (uint32_t)kp.BNI};
      int r = encode_tmap_bf16(&kp.tmA[s], sg.a, 4, dims, strides, box);
      if (r != FG_OK) return r;
    }
    {
      const uint64_t dims[3] = {(uint64_t)sg.cin, (uint64_t)sg.taps, (uint64_t)p->cout};
      const uint64_t strides[3] = {2, (uint64_t)sg.cin * 2, (uint64_t)sg.cin * sg.taps * 2};
      const uint32_t box[3] = {(uint32_t)kBlockK, 1u, (uint32_t)block_n};
      int r = encode_tmap_bf16(&kp.tmW[s], sg.w, 3, dims, strides, box);
      if (r != FG_OK) return r;
    }
  }

  kp.bias = p->bias;
  kp.rowvec = p->rowvec;
  kp.rowvec_ld = p->rowvec_ld;
  kp.res = p->res;
  kp.res_dtype = p->res_dtype;
  kp.res_mode = p->res ? p->res_mode : FG_RES_NONE;
  kp.res_ld = p->res_ld;
  kp.out_scale = p->out_scale;
  kp.out = p->out;
  kp.out_dtype = p->out_dtype;
  kp.out_ld = p->out_ld;
  FG_CHECK_ARG(kp.res_mode >= FG_RES_NONE && kp.res_mode <= FG_RES_DOWN2, "conv_gemm: bad res_mode %d", kp.res_mode);
  if (kp.res_mode != FG_RES_NONE) {
    FG_CHECK_ARG(p->res_dtype == FG_F32 || p->res_dtype == FG_BF16, "conv_gemm: bad res_dtype %d", p->res_dtype);
    FG_CHECK_ARG(p->res_ld >= p->cout, "conv_gemm: res_ld %lld < cout", (long long)p->res_ld);
    if (kp.res_mode == FG_RES_UP2)
      FG_CHECK_ARG(p->h % 2 == 0 && p->w % 2 == 0, "conv_gemm: FG_RES_UP2 needs even h,w");
  }
  kp.vec = (p->cout % 4 == 0) && (p->out_ld % 4 == 0) && aligned16(p->out) && (!p->bias || aligned16(p->bias)) &&
           (!p->rowvec || (aligned16(p->rowvec) && p->rowvec_ld % 4 == 0)) &&
           (kp.res_mode == FG_RES_NONE || (aligned16(p->res) && p->res_ld % 4 == 0));

  // TMA-store epilogue whenever the output rows are 16-byte addressable and the tile is at least 32 wide
  kp.tma_store = 0;
  {
    const int esz = p->out_dtype == FG_F32 ? 4 : 2;
    static const bool allow_tma_store = !(getenv("FG_CONV_TMA_STORE") && getenv("FG_CONV_TMA_STORE")[0] == '0');
    if (allow_tma_store && kp.vec && block_n >= 64 && (p->out_ld * esz) % 16 == 0) {
      const uint64_t dims[4] = {(uint64_t)p->cout, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
      const uint64_t ldb = (uint64_t)p->out_ld * esz;
      const uint64_t strides[4] = {(uint64_t)esz, ldb, ldb * p->w, ldb * p->w * p->h};
      const uint32_t box[4] = {(uint32_t)(128 / esz), (uint32_t)kp.BW, (uint32_t)kp.BH, (uint32_t)kp.BNI};
      int r = encode_tmap(&kp.tmOut, p->out, p->out_dtype, 4, dims, strides, box);
      if (r != FG_OK) return r;
      kp.tma_store = 1;
    }
  }

  int dev = 0, num_sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("conv_gemm: cannot query device");
    return FG_ERR_CUDA;
  }
  switch (block_n) {
    case 16: return launch_conv<16>(kp, num_sms, stream);
    case 64: return launch_conv<64>(kp, num_sms, stream);
    case 128: return launch_conv<128>(kp, num_sms, stream);
    case 192: return launch_conv<192>(kp, num_sms, stream);
    case 256: return launch_conv<256>(kp, num_sms, stream);
  }
  set_error("conv_gemm: no kernel for block_n=%d", block_n);
  return FG_ERR_UNSUPPORTED;
}
