// GroupNorm for channels-last activations, split into a statistics pass and a fused apply pass.
//   fg_gn_stats : per (sample, group) sum / sum-of-squares, fp32 per-thread partials merged in fp64.
//   fg_gn_apply : y = [silu]( (x-mean)*rstd*gamma+beta ) * (1+scale) + shift ), written as the bf16 operand
//                 of the next convolution, optionally 2x up/down-sampled on the fly, optionally with a bf16
//                 low part (split precision) and a bf16 copy of the raw input for the 1x1 skip projection.
// Both kernels are HBM-bound streaming kernels: 16-byte vector accesses, a warp covers 512 contiguous bytes.
//
// Replaces: fastgen/networks/EDM/network.py:141-149 (GroupNorm.forward), silu/addcmul at :276-283 and the
// depthwise resample convolutions at :114-121.
#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kGnThreads = 256;

template <typename T>
struct Vec;  // 16-byte vector of T
template <>
struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    float2 f;
    f = unpack_bf16x2(a.x), v[0] = f.x, v[1] = f.y;
    f = unpack_bf16x2(a.y), v[2] = f.x, v[3] = f.y;
    f = unpack_bf16x2(a.z), v[4] = f.x, v[5] = f.y;
    f = unpack_bf16x2(a.w), v[6] = f.x, v[7] = f.y;
  }
};

// grid (slabs, n).  A warp owns a pixel row subset; lane l handles channel vectors l, l+32, ...
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const T* __restrict__ x, int64_t ld, int hw, int c,
                                                              int cg, int groups, int pix_per_slab,
                                                              double* __restrict__ stats) {
  constexpr int V = Vec<T>::N;
  extern __shared__ double s_acc[];  // [groups][2]
  for (int i = threadIdx.x; i < groups * 2; i += kGnThreads) s_acc[i] = 0.0;
  __syncthreads();

  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_slab;
  const int p1 = min(hw, p0 + pix_per_slab);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kGnThreads / 32;
  const int nvec = c / V;
  const T* base = x + static_cast<int64_t>(n) * hw * ld;

  for (int vi = lane; vi < nvec; vi += 32) {
    float s[V], ss[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s[k] = 0.f, ss[k] = 0.f;
    int pix = p0 + warp;
    // 4 independent loads in flight per thread
    for (; pix + 3 * kWarps < p1; pix += 4 * kWarps) {
      float v0[V], v1[V], v2[V], v3[V];
      Vec<T>::load(base + static_cast<int64_t>(pix) * ld + vi * V, v0);
      Vec<T>::load(base + static_cast<int64_t>(pix + kWarps) * ld + vi * V, v1);
      Vec<T>::load(base + static_cast<int64_t>(pix + 2 * kWarps) * ld + vi * V, v2);
      Vec<T>::load(base + static_cast<int64_t>(pix + 3 * kWarps) * ld + vi * V, v3);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        s[k] += (v0[k] + v1[k]) + (v2[k] + v3[k]);
        ss[k] += (v0[k] * v0[k] + v1[k] * v1[k]) + (v2[k] * v2[k] + v3[k] * v3[k]);
      }
    }
    for (; pix < p1; pix += kWarps) {
      float v0[V];
      Vec<T>::load(base + static_cast<int64_t>(pix) * ld + vi * V, v0);
#pragma unroll
      for (int k = 0; k < V; ++k) s[k] += v0[k], ss[k] += v0[k] * v0[k];
    }
    // fold the V channels of this thread into their groups (a vector may straddle a group boundary)
    int g_prev = (vi * V) / cg;
    double ds = 0.0, dss = 0.0;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int g = (vi * V + k) / cg;
      if (g != g_prev) {
        atomicAdd(&s_acc[2 * g_prev], ds);
        atomicAdd(&s_acc[2 * g_prev + 1], dss);
        ds = 0.0, dss = 0.0, g_prev = g;
      }
      ds += static_cast<double>(s[k]);
      dss += static_cast<double>(ss[k]);
    }
    atomicAdd(&s_acc[2 * g_prev], ds);
    atomicAdd(&s_acc[2 * g_prev + 1], dss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += kGnThreads)
    atomicAdd(&stats[static_cast<int64_t>(n) * groups * 2 + i], s_acc[i]);
}

struct GnApplyK {
  const void* x;
  int64_t x_ld;
  int n, h, w, c, groups, cg;
  const double* stats;
  float eps;
  const float* gamma;
  const float* beta;
  const float* scale;
  const float* shift;
  int64_t mod_ld;
  int silu, resample;
  __nv_bfloat16* y;
  int64_t y_ld;
  __nv_bfloat16* y_lo;
  __nv_bfloat16* raw;
  int64_t raw_ld;
  __nv_bfloat16* raw_lo;
  float* raw_f32;    // FG_RES_DOWN2 only: the 2x2 mean of x itself in fp32 (the block's resampled identity skip)
  int64_t raw_f32_ld;
  int pix_per_slab;  // in OUTPUT pixels for DOWN2, input pixels otherwise
  int vpp;           // 4-channel vectors per pass over the channels (= threads per pixel row)
};

__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void split8(const float (&v)[8], float (&lo)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) lo[k] = v[k] - __bfloat162float(__float2bfloat16_rn(v[k]));
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  Vec<__nv_bfloat16>::load(p, v);
}

// grid (slabs, n).  A thread owns ONE 4-channel vector (its 8 coefficients of y = a*x + b live in registers, computed
// by the thread itself from the fp64 statistics: no shared memory, no block barrier) and walks the pixels of the slab
// with four independent 16-byte loads in flight.  Small register footprint (launch bounds: 4 blocks of up to 288 threads
// per SM) so that enough warps are resident to hide the SiLU's MUFU latency -- profiles/ncu_r01.md has the capture of
// the previous version (89 registers, 14 warps per SM, 2-4 waves of big blocks: 2.2-3.7 TB/s).
constexpr int kGnApplyMaxThreads = 288;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 f0 = unpack_bf16x2(a.x), f1 = unpack_bf16x2(a.y);
  v[0] = f0.x, v[1] = f0.y, v[2] = f1.x, v[3] = f1.y;
}
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void split4(const float (&v)[4], float (&lo)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) lo[k] = v[k] - __bfloat162float(__float2bfloat16_rn(v[k]));
}

template <typename T, int MODE, bool LO>  // MODE: FG_RES_NONE / FG_RES_UP2 / FG_RES_DOWN2; LO: split-precision outputs
__global__ void __launch_bounds__(kGnApplyMaxThreads, 4) gn_apply_kernel(const GnApplyK p) {
  const int n = blockIdx.y;
  const int hw = p.h * p.w;
  const T* xb = static_cast<const T*>(p.x) + static_cast<int64_t>(n) * hw * p.x_ld;
  const int nvec = p.c >> 2;
  const int vpp = p.vpp;             // vectors per pass (<= blockDim.x, host-chosen)
  const int ppi = blockDim.x / vpp;  // pixels per iteration of the block
  const int prow = threadIdx.x / vpp;
  if (prow >= ppi) return;
  constexpr bool up = MODE == FG_RES_UP2;
  constexpr bool down = MODE == FG_RES_DOWN2;
  const int ow = down ? (p.w >> 1) : p.w;
  const int npix = down ? (p.h >> 1) * ow : hw;  // pixels this kernel iterates over (output pixels for DOWN2)
  const int p0 = blockIdx.x * p.pix_per_slab;
  const int p1 = min(npix, p0 + p.pix_per_slab);
  const double inv_cnt = 1.0 / (static_cast<double>(hw) * p.cg);

  for (int vi = threadIdx.x % vpp; vi < nvec; vi += vpp) {
    const int co = vi * 4;
    float a4[4], b4[4];
    {
      int g_prev = -1;
      float meanf = 0.f, rstd = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ch = co + k;
        const int g = ch / p.cg;
        if (g != g_prev) {  // (a 4-channel vector touches one or two groups, rarely more)
          const double* st = p.stats + (static_cast<int64_t>(n) * p.groups + g) * 2;
          const double mean = st[0] * inv_cnt;
          double var = st[1] * inv_cnt - mean * mean;
          var = var > 0.0 ? var : 0.0;
          rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
          meanf = static_cast<float>(mean);
          g_prev = g;
        }
        float a = rstd * __ldg(p.gamma + ch);
        float b = __ldg(p.beta + ch) - meanf * a;
        if (p.scale) {
          const float sc = 1.0f + __ldg(p.scale + n * p.mod_ld + ch);
          a *= sc;
          b *= sc;
        }
        if (p.shift) b += __ldg(p.shift + n * p.mod_ld + ch);
        a4[k] = a, b4[k] = b;
      }
    }

    if constexpr (down) {
      for (int opix = p0 + prow; opix < p1; opix += ppi) {
        const int oy = opix / ow, ox = opix % ow;
        float v[4][4];
#pragma unroll
        for (int d = 0; d < 4; ++d)
          load4<T>(xb + static_cast<int64_t>((2 * oy + (d >> 1)) * p.w + 2 * ox + (d & 1)) * p.x_ld + co, v[d]);
        float acc[4], racc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float y0 = fmaf(a4[k], v[0][k], b4[k]), y1 = fmaf(a4[k], v[1][k], b4[k]);
          float y2 = fmaf(a4[k], v[2][k], b4[k]), y3 = fmaf(a4[k], v[3][k], b4[k]);
          if (p.silu) y0 = silu_f(y0), y1 = silu_f(y1), y2 = silu_f(y2), y3 = silu_f(y3);
          acc[k] = 0.25f * ((y0 + y1) + (y2 + y3));
          racc[k] = 0.25f * ((v[0][k] + v[1][k]) + (v[2][k] + v[3][k]));
        }
        const int64_t o = static_cast<int64_t>(n) * npix + opix;
        store4_bf16(p.y + o * p.y_ld + co, acc);
        if (LO && p.y_lo) {
          float lo[4];
          split4(acc, lo);
          store4_bf16(p.y_lo + o * p.y_ld + co, lo);
        }
        if (p.raw_f32)
          *reinterpret_cast<float4*>(p.raw_f32 + o * p.raw_f32_ld + co) = make_float4(racc[0], racc[1], racc[2], racc[3]);
        if (p.raw) {
          store4_bf16(p.raw + o * p.raw_ld + co, racc);
          if (LO && p.raw_lo) {
            float lo[4];
            split4(racc, lo);
            store4_bf16(p.raw_lo + o * p.raw_ld + co, lo);
          }
        }
      }
      continue;
    }

    auto emit = [&](int pix, const float (&v)[4]) {
      float yv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = fmaf(a4[k], v[k], b4[k]);
        yv[k] = p.silu ? silu_f(t) : t;
      }
      float ylo[4], rlo[4];
      if (LO && p.y_lo) split4(yv, ylo);
      if (LO && p.raw_lo) split4(v, rlo);
      if constexpr (up) {
        const int iy = pix / p.w, ix = pix % p.w;
        const int W2 = p.w * 2;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int64_t o =
              static_cast<int64_t>(n) * hw * 4 + static_cast<int64_t>(2 * iy + (d >> 1)) * W2 + 2 * ix + (d & 1);
          store4_bf16(p.y + o * p.y_ld + co, yv);
          if (LO && p.y_lo) store4_bf16(p.y_lo + o * p.y_ld + co, ylo);
          if (p.raw) store4_bf16(p.raw + o * p.raw_ld + co, v);
          if (LO && p.raw_lo) store4_bf16(p.raw_lo + o * p.raw_ld + co, rlo);
        }
      } else {
        const int64_t o = static_cast<int64_t>(n) * hw + pix;
        store4_bf16(p.y + o * p.y_ld + co, yv);
        if (LO && p.y_lo) store4_bf16(p.y_lo + o * p.y_ld + co, ylo);
        if (p.raw) store4_bf16(p.raw + o * p.raw_ld + co, v);
        if (LO && p.raw_lo) store4_bf16(p.raw_lo + o * p.raw_ld + co, rlo);
      }
    };

    int pix = p0 + prow;
    if constexpr (sizeof(T) == 2 && !LO) {
      // bf16 input: 8-byte loads, so eight of them in flight per thread keep as many bytes on the way as the four 16-byte
      // loads of the fp32 form (with four, the bf16-input launches ran at the fp32 launches' time: latency-, not byte-bound)
      for (; pix + 7 * ppi < p1; pix += 8 * ppi) {
        uint2 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          r[u] = __ldg(reinterpret_cast<const uint2*>(xb + static_cast<int64_t>(pix + u * ppi) * p.x_ld + co));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float2 f0 = unpack_bf16x2(r[u].x), f1 = unpack_bf16x2(r[u].y);
          const float v[4] = {f0.x, f0.y, f1.x, f1.y};
          emit(pix + u * ppi, v);
        }
      }
    }
    for (; pix + 3 * ppi < p1; pix += 4 * ppi) {
      float v0[4], v1[4], v2[4], v3[4];
      load4<T>(xb + static_cast<int64_t>(pix) * p.x_ld + co, v0);
      load4<T>(xb + static_cast<int64_t>(pix + ppi) * p.x_ld + co, v1);
      load4<T>(xb + static_cast<int64_t>(pix + 2 * ppi) * p.x_ld + co, v2);
      load4<T>(xb + static_cast<int64_t>(pix + 3 * ppi) * p.x_ld + co, v3);
      emit(pix, v0);
      emit(pix + ppi, v1);
      emit(pix + 2 * ppi, v2);
      emit(pix + 3 * ppi, v3);
    }
    for (; pix < p1; pix += ppi) {
      float v0[4];
      load4<T>(xb + static_cast<int64_t>(pix) * p.x_ld + co, v0);
      emit(pix, v0);
    }
  }
}

// y = bf16(x) for a channels-last tensor, optionally nearest-upsampled 2x: the raw (un-normalised) conv operands of the
// SDXL UNet's Downsample2D / Upsample2D (diffusers; wrapped by fastgen/networks/SDXL/network.py:210-230).
template <typename T, bool UP>
__global__ void __launch_bounds__(kGnThreads) cast_resample_kernel(const T* __restrict__ x, int64_t x_ld, int n, int h,
                                                                   int w, int c, __nv_bfloat16* __restrict__ y,
                                                                   int64_t y_ld) {
  const int nvec = c >> 3;
  const int64_t total = static_cast<int64_t>(n) * h * w * nvec;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kGnThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kGnThreads) {
    const int64_t pix = i / nvec;
    const int co = static_cast<int>(i % nvec) * 8;
    float v[8];
    load8<T>(x + pix * x_ld + co, v);
    if constexpr (UP) {
      const int ix = static_cast<int>(pix % w);
      const int64_t t = pix / w;
      const int iy = static_cast<int>(t % h);
      const int64_t img = t / h;
      const int W2 = 2 * w;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t o = (img * 2 * h + 2 * iy + (d >> 1)) * W2 + 2 * ix + (d & 1);
        store8_bf16(y + o * y_ld + co, v);
      }
    } else {
      store8_bf16(y + pix * y_ld + co, v);
    }
  }
}

// fg_gn_apply has no accumulation, so its slab may depend on the batch.  ~24-96K elements per CTA keep the per-thread
// prologue (coefficients from the fp64 statistics) at a few percent of the work and give ~8 waves of 4 blocks per SM
// at the bench batch, so the tail is short.
static int pick_slab_apply(int n, int pixels, int c, int* pix_per_slab) {
  const int64_t total = static_cast<int64_t>(n) * pixels * c;
  int64_t per_cta = total / (148 * 4 * 8);
  if (per_cta < 24576) per_cta = 24576;
  if (per_cta > 98304) per_cta = 98304;
  int64_t slabs = (static_cast<int64_t>(pixels) * c + per_cta - 1) / per_cta;
  if (slabs > pixels) slabs = pixels;
  if (slabs < 1) slabs = 1;
  *pix_per_slab = static_cast<int>((pixels + slabs - 1) / slabs);
  return static_cast<int>((pixels + *pix_per_slab - 1) / *pix_per_slab);
}

static int pick_slab(int n, int pixels, int c, int* pix_per_slab) {
  // ~32K elements per CTA.  Deliberately independent of the batch size: a sample's partial sums are then
  // accumulated in the same order whatever batch it is part of.
  (void)n;
  int64_t slabs = (static_cast<int64_t>(pixels) * c + 32767) / 32768;
  // at least 64 pixels per CTA (8 per warp): with very wide tensors (SDXL concatenations, c = 1920) a 32K-element slab
  // is only 17 pixels and the per-vector merges into shared memory dominate (measured 1.1 TB/s)
  if (slabs > (pixels + 63) / 64) slabs = (pixels + 63) / 64;
  if (slabs < 1) slabs = 1;
  *pix_per_slab = static_cast<int>((pixels + slabs - 1) / slabs);
  return static_cast<int>((pixels + *pix_per_slab - 1) / *pix_per_slab);
}

}  // namespace fg

extern "C" int fg_gn_stats(const void* x, int32_t x_dtype, int64_t x_ld, int32_t n, int32_t hw, int32_t c,
                           int32_t groups, double* stats, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && stats, "gn_stats: null pointer");
  FG_CHECK_ARG(n > 0 && hw > 0 && c > 0 && groups > 0 && c % groups == 0, "gn_stats: bad shape n=%d hw=%d c=%d g=%d", n,
               hw, c, groups);
  FG_CHECK_ARG(x_dtype == FG_F32 || x_dtype == FG_BF16, "gn_stats: bad dtype %d", x_dtype);
  const int v = x_dtype == FG_F32 ? 4 : 8;
  FG_CHECK_ARG(c % v == 0 && x_ld % v == 0 && x_ld >= c, "gn_stats: c=%d / ld=%lld must be multiples of %d", c,
               (long long)x_ld, v);
  FG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "gn_stats: x must be 16-byte aligned");
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * n * groups, stream) != cudaSuccess) {
    set_error("gn_stats: memset failed");
    return FG_ERR_CUDA;
  }
  int pps = 0;
  const int slabs = pick_slab(n, hw, c, &pps);
  const dim3 grid(slabs, n);
  const size_t smem = sizeof(double) * 2 * groups;
  if (x_dtype == FG_F32)
    gn_stats_kernel<float><<<grid, kGnThreads, smem, stream>>>(static_cast<const float*>(x), x_ld, hw, c, c / groups,
                                                               groups, pps, stats);
  else
    gn_stats_kernel<__nv_bfloat16><<<grid, kGnThreads, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, hw,
                                                                       c, c / groups, groups, pps, stats);
  FG_CHECK_LAUNCH("gn_stats");
  return FG_OK;
}

extern "C" int fg_gn_apply(const fg_gn_apply_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p && p->x && p->stats && p->gamma && p->beta && p->y, "gn_apply: null pointer");
  FG_CHECK_ARG(p->n > 0 && p->h > 0 && p->w > 0 && p->c > 0 && p->groups > 0 && p->c % p->groups == 0,
               "gn_apply: bad shape");
  FG_CHECK_ARG(p->x_dtype == FG_F32 || p->x_dtype == FG_BF16, "gn_apply: bad dtype %d", p->x_dtype);
  FG_CHECK_ARG(p->c % 8 == 0 && p->x_ld % 8 == 0 && p->y_ld % 8 == 0 && p->x_ld >= p->c && p->y_ld >= p->c,
               "gn_apply: c=%d, x_ld=%lld, y_ld=%lld must be multiples of 8 and ld >= c", p->c, (long long)p->x_ld,
               (long long)p->y_ld);
  FG_CHECK_ARG(p->resample >= FG_RES_NONE && p->resample <= FG_RES_DOWN2 && p->resample != FG_RES_SAME,
               "gn_apply: bad resample %d", p->resample);
  if (p->resample == FG_RES_DOWN2) FG_CHECK_ARG(p->h % 2 == 0 && p->w % 2 == 0, "gn_apply: DOWN2 needs even h,w");
  if (p->raw) FG_CHECK_ARG(p->raw_ld % 8 == 0 && p->raw_ld >= p->c, "gn_apply: bad raw_ld");
  if (p->scale || p->shift) FG_CHECK_ARG(p->mod_ld >= p->c, "gn_apply: mod_ld < c");
  GnApplyK k;
  k.x = p->x, k.x_ld = p->x_ld, k.n = p->n, k.h = p->h, k.w = p->w, k.c = p->c, k.groups = p->groups;
  k.cg = p->c / p->groups;
  k.stats = p->stats, k.eps = p->eps, k.gamma = p->gamma, k.beta = p->beta;
  k.scale = p->scale, k.shift = p->shift, k.mod_ld = p->mod_ld, k.silu = p->silu, k.resample = p->resample;
  k.y = static_cast<__nv_bfloat16*>(p->y), k.y_ld = p->y_ld, k.y_lo = static_cast<__nv_bfloat16*>(p->y_lo);
  k.raw = static_cast<__nv_bfloat16*>(p->raw), k.raw_ld = p->raw_ld, k.raw_lo = static_cast<__nv_bfloat16*>(p->raw_lo);
  if (p->raw_f32)
    FG_CHECK_ARG(p->resample == FG_RES_DOWN2 && p->raw_f32_ld % 4 == 0 && p->raw_f32_ld >= p->c &&
                     (reinterpret_cast<uintptr_t>(p->raw_f32) & 15u) == 0,
                 "gn_apply: raw_f32 needs FG_RES_DOWN2, 16-byte aligned rows, raw_f32_ld >= c");
  k.raw_f32 = static_cast<float*>(p->raw_f32), k.raw_f32_ld = p->raw_f32_ld;
  const int pixels = p->resample == FG_RES_DOWN2 ? (p->h / 2) * (p->w / 2) : p->h * p->w;
  const int slabs = pick_slab_apply(p->n, pixels, p->c, &k.pix_per_slab);
  const dim3 grid(slabs, p->n);
  // threads: a multiple of the vectors-per-pass close to 256 (<= 288); very wide tensors take several passes per thread
  const int nvec = p->c / 4;
  const int passes = (nvec + kGnApplyMaxThreads - 1) / kGnApplyMaxThreads;
  k.vpp = (nvec + passes - 1) / passes;
  int rows_per_iter = (256 + k.vpp / 2) / k.vpp;
  if (rows_per_iter < 1) rows_per_iter = 1;
  while (rows_per_iter > 1 && rows_per_iter * k.vpp > kGnApplyMaxThreads) --rows_per_iter;
  const int threads = rows_per_iter * k.vpp;
  const size_t smem = 0;
  const bool lo = p->y_lo || p->raw_lo;
#define FG_GN_LAUNCH(T, MODE)                                             \
  do {                                                                    \
    if (lo)                                                               \
      gn_apply_kernel<T, MODE, true><<<grid, threads, smem, stream>>>(k); \
    else                                                                  \
      gn_apply_kernel<T, MODE, false><<<grid, threads, smem, stream>>>(k); \
  } while (0)
#define FG_GN_DISPATCH(T)                                        \
  do {                                                           \
    if (p->resample == FG_RES_UP2)                               \
      FG_GN_LAUNCH(T, FG_RES_UP2);                               \
    else if (p->resample == FG_RES_DOWN2)                        \
      FG_GN_LAUNCH(T, FG_RES_DOWN2);                             \
    else                                                         \
      FG_GN_LAUNCH(T, FG_RES_NONE);                              \
  } while (0)
  if (p->x_dtype == FG_F32)
    FG_GN_DISPATCH(float);
  else
    FG_GN_DISPATCH(__nv_bfloat16);
#undef FG_GN_DISPATCH
#undef FG_GN_LAUNCH
  FG_CHECK_LAUNCH("gn_apply");
  return FG_OK;
}

extern "C" int fg_cast_resample(const void* x, int32_t x_dtype, int64_t x_ld, int32_t n, int32_t h, int32_t w, int32_t c,
                                int32_t resample, void* y, int64_t y_ld, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0, "cast_resample: bad arguments");
  FG_CHECK_ARG(x_dtype == FG_F32 || x_dtype == FG_BF16, "cast_resample: bad dtype %d", x_dtype);
  FG_CHECK_ARG(resample == FG_RES_NONE || resample == FG_RES_UP2, "cast_resample: resample must be NONE or UP2");
  FG_CHECK_ARG(c % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && x_ld >= c && y_ld >= c,
               "cast_resample: c=%d, x_ld=%lld, y_ld=%lld must be multiples of 8 and ld >= c", c, (long long)x_ld,
               (long long)y_ld);
  FG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0,
               "cast_resample: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 8);
  int64_t g = (total + kGnThreads - 1) / kGnThreads;
  if (g > 148 * 16) g = 148 * 16;
  const int grid = static_cast<int>(g);
  auto* yo = static_cast<__nv_bfloat16*>(y);
  const bool up = resample == FG_RES_UP2;
  if (x_dtype == FG_F32) {
    const float* xi = static_cast<const float*>(x);
    if (up) cast_resample_kernel<float, true><<<grid, kGnThreads, 0, stream>>>(xi, x_ld, n, h, w, c, yo, y_ld);
    else cast_resample_kernel<float, false><<<grid, kGnThreads, 0, stream>>>(xi, x_ld, n, h, w, c, yo, y_ld);
  } else {
    const __nv_bfloat16* xi = static_cast<const __nv_bfloat16*>(x);
    if (up) cast_resample_kernel<__nv_bfloat16, true><<<grid, kGnThreads, 0, stream>>>(xi, x_ld, n, h, w, c, yo, y_ld);
    else cast_resample_kernel<__nv_bfloat16, false><<<grid, kGnThreads, 0, stream>>>(xi, x_ld, n, h, w, c, yo, y_ld);
  }
  FG_CHECK_LAUNCH("cast_resample");
  return FG_OK;
}
