// The steps either side of the sampler: what the reference's encoder / decoder WRAPPERS compute around the
// external VAE and text-encoder networks -- per-channel latent (de)normalisation, the clamp of the decoded pixels and
// the zero padding of the prompt embeddings behind each prompt's last token
// (fastgen/networks/Wan/network.py:431-539, SDXL/network.py:46-156).  All three are one streaming pass at HBM speed
// where the reference runs three to five eager tensor ops; their rounding sequence is the reference's (every eager op
// rounds to the tensor's dtype), so the results are bit-identical.
#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kHoThreads = 256;

__device__ __forceinline__ float round_as(float v, bool bf16) { return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v; }

// out[b, c, i] = OP(x[b, c, i]; a[c], b[c]), i < inner.  Arithmetic in fp32 with a rounding to x's dtype after every
// reference op (a, b arrive already rounded to that dtype):
//   mode 0 (decode, Wan/network.py:517-528):  (x / a) + b          a = 1 / std, b = mean
//   mode 1 (encode, :496-513):                (x - b) * a
//   mode 2 (SDXL/network.py:141-149):         x * a                a = scaling factor or its reciprocal, per tensor
template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(kHoThreads) latent_affine_kernel(const TI* __restrict__ x, TO* __restrict__ out,
                                                                   const float* __restrict__ a, const float* __restrict__ b,
                                                                   int64_t outer, int c, int64_t inner) {
  constexpr bool kBf = sizeof(TI) == 2;
  constexpr int V = 16 / sizeof(TI);            // elements per 16-byte vector of the input
  const int64_t row = blockIdx.y;               // (batch, channel) row
  if (row >= outer * c) return;
  const int ch = static_cast<int>(row % c);
  const float av = a[ch], bv = b ? b[ch] : 0.f;
  const TI* xr = x + row * inner;
  TO* orow = out + row * inner;
  auto op = [&](float v) {
    if (MODE == 0) return round_as(round_as(__fdiv_rn(v, av), kBf) + bv, kBf);
    if (MODE == 1) return round_as(round_as(v - bv, kBf) * av, kBf);
    return round_as(v * av, kBf);
  };
  const bool vec = (inner % V == 0) && ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(orow) & (V * sizeof(TO) - 1)) == 0);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kHoThreads;
  if (vec) {
    for (int64_t i = (blockIdx.x * static_cast<int64_t>(kHoThreads) + threadIdx.x) * V; i < inner; i += stride * V) {
      float v[V];
      if constexpr (kBf) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + i);
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        v[0] = f0.x, v[1] = f0.y, v[2] = f1.x, v[3] = f1.y, v[4] = f2.x, v[5] = f2.y, v[6] = f3.x, v[7] = f3.y;
      } else {
        const float4 u = *reinterpret_cast<const float4*>(xr + i);
        v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w;
      }
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = op(v[k]);
      if constexpr (sizeof(TO) == 2) {
#pragma unroll
        for (int k = 0; k < V; k += 2)
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(orow) + i + k) = pack_bf16x2(v[k], v[k + 1]);
      } else {
#pragma unroll
        for (int k = 0; k < V; k += 4)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + i + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
  } else {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kHoThreads) + threadIdx.x; i < inner; i += stride) {
      float v;
      if constexpr (kBf) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xr)[i]);
      else v = reinterpret_cast<const float*>(xr)[i];
      v = op(v);
      if constexpr (sizeof(TO) == 2) reinterpret_cast<__nv_bfloat16*>(orow)[i] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(orow)[i] = v;
    }
  }
}

// x <- min(max(x, lo), hi) in place (the `.clip_(-1, 1)` behind both VAE decoders); NaN stays NaN like torch.clamp
template <typename T>
__global__ void __launch_bounds__(kHoThreads) clamp_kernel(T* __restrict__ x, int64_t total, float lo, float hi) {
  constexpr int V = 16 / sizeof(T);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kHoThreads;
  const int64_t nvec = total / V;
  auto cl = [&](float v) { return v != v ? v : fminf(fmaxf(v, lo), hi); };
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kHoThreads) + threadIdx.x; i < nvec; i += stride) {
    if constexpr (sizeof(T) == 2) {
      uint4 u = reinterpret_cast<uint4*>(x)[i];
      uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        w[k] = pack_bf16x2(cl(f.x), cl(f.y));
      }
      reinterpret_cast<uint4*>(x)[i] = u;
    } else {
      float4 u = reinterpret_cast<float4*>(x)[i];
      u.x = cl(u.x), u.y = cl(u.y), u.z = cl(u.z), u.w = cl(u.w);
      reinterpret_cast<float4*>(x)[i] = u;
    }
  }
  for (int64_t i = nvec * V + blockIdx.x * static_cast<int64_t>(kHoThreads) + threadIdx.x; i < total; i += stride) {
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(x) + i;
      *p = __float2bfloat16_rn(cl(__bfloat162float(*p)));
    } else {
      float* p = reinterpret_cast<float*>(x) + i;
      *p = cl(*p);
    }
  }
}

// out[b, t, :] = t < len[b] ? cast(x[b, t, :]) : 0  -- "u[:v] then pad with zeros to max_length" of WanTextEncoder.encode.
// VEC: four elements per thread and iteration (rows of a multiple of 4 elements, 16-byte aligned input rows); rows behind a
// prompt's last token are written without being read.
template <typename TI, typename TO, bool VEC>
__global__ void __launch_bounds__(kHoThreads) zero_tail_rows_kernel(const TI* __restrict__ x, int64_t x_ld, int64_t x_bs,
                                                                    const int64_t* __restrict__ len, int b, int rows, int d,
                                                                    TO* __restrict__ out) {
  constexpr int V = VEC ? 4 : 1;
  const int dv = d / V;
  const int64_t total = static_cast<int64_t>(b) * rows * dv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kHoThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kHoThreads) {
    const int col = static_cast<int>(i % dv) * V;
    const int64_t r = i / dv;
    const int t = static_cast<int>(r % rows), s = static_cast<int>(r / rows);
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.f;
    if (t < len[s]) {
      const TI* src = x + s * x_bs + t * x_ld + col;
      if constexpr (VEC && sizeof(TI) == 4) {
        const float4 u = *reinterpret_cast<const float4*>(src);
        v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w;
      } else if constexpr (VEC) {
        const uint2 u = *reinterpret_cast<const uint2*>(src);
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y);
        v[0] = f0.x, v[1] = f0.y, v[2] = f1.x, v[3] = f1.y;
      } else if constexpr (sizeof(TI) == 2) {
        v[0] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(src));
      } else {
        v[0] = *reinterpret_cast<const float*>(src);
      }
    }
    TO* dst = out + r * d + col;
    if constexpr (VEC && sizeof(TO) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (VEC) {
      uint2 u;
      u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(dst) = u;
    } else if constexpr (sizeof(TO) == 2) {
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v[0]);
    } else {
      *reinterpret_cast<float*>(dst) = v[0];
    }
  }
}

static int ho_grid(int64_t work_items) {
  int64_t g = (work_items + kHoThreads - 1) / kHoThreads;
  const int64_t cap = 148 * 8;
  return static_cast<int>(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace fg

extern "C" int fg_latent_affine(const void* x, int32_t x_dtype, void* out, int32_t out_dtype, const float* a, const float* b,
                                int64_t outer, int32_t c, int64_t inner, int32_t mode, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && out && a && outer > 0 && c > 0 && inner > 0, "latent_affine: bad arguments");
  FG_CHECK_ARG(x_dtype == FG_F32 || x_dtype == FG_BF16, "latent_affine: bad x dtype %d", x_dtype);
  FG_CHECK_ARG(out_dtype == FG_F32 || out_dtype == FG_BF16, "latent_affine: bad out dtype %d", out_dtype);
  FG_CHECK_ARG(mode >= 0 && mode <= 2 && (mode == 2 || b), "latent_affine: bad mode %d / missing b", mode);
  FG_CHECK_ARG(outer * c < 65536, "latent_affine: more than 65535 (batch, channel) rows");
  const int v = x_dtype == FG_F32 ? 4 : 8;
  int gx = ho_grid((inner + v - 1) / v);
  const int64_t rows = outer * c;
  while (gx > 1 && gx * rows > 148 * 16) gx = (gx + 1) / 2;   // enough CTAs to fill the device, no more
  const dim3 grid(gx, static_cast<unsigned>(rows));
#define FG_LA(TI, TO, M) latent_affine_kernel<TI, TO, M><<<grid, kHoThreads, 0, stream>>>( \
      static_cast<const TI*>(x), static_cast<TO*>(out), a, b, outer, c, inner)
#define FG_LA_MODE(TI, TO)                 \
  do {                                     \
    if (mode == 0) FG_LA(TI, TO, 0);       \
    else if (mode == 1) FG_LA(TI, TO, 1);  \
    else FG_LA(TI, TO, 2);                 \
  } while (0)
  if (x_dtype == FG_F32 && out_dtype == FG_F32) FG_LA_MODE(float, float);
  else if (x_dtype == FG_F32) FG_LA_MODE(float, __nv_bfloat16);
  else if (out_dtype == FG_F32) FG_LA_MODE(__nv_bfloat16, float);
  else FG_LA_MODE(__nv_bfloat16, __nv_bfloat16);
#undef FG_LA_MODE
#undef FG_LA
  FG_CHECK_LAUNCH("latent_affine");
  return FG_OK;
}

extern "C" int fg_clamp(void* x, int32_t dtype, int64_t total, float lo, float hi, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && total > 0 && lo <= hi, "clamp: bad arguments");
  FG_CHECK_ARG(dtype == FG_F32 || dtype == FG_BF16, "clamp: bad dtype %d", dtype);
  FG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "clamp: tensor not 16-byte aligned");
  const int grid = ho_grid(total / (dtype == FG_F32 ? 4 : 8) + 1);
  if (dtype == FG_F32) clamp_kernel<float><<<grid, kHoThreads, 0, stream>>>(static_cast<float*>(x), total, lo, hi);
  else clamp_kernel<__nv_bfloat16><<<grid, kHoThreads, 0, stream>>>(static_cast<__nv_bfloat16*>(x), total, lo, hi);
  FG_CHECK_LAUNCH("clamp");
  return FG_OK;
}

extern "C" int fg_zero_tail_rows(const void* x, int32_t x_dtype, int64_t x_ld, int64_t x_bs, const int64_t* len, int32_t b,
                                 int32_t rows, int32_t d, void* out, int32_t out_dtype, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && out && len && b > 0 && rows > 0 && d > 0 && x_ld >= d, "zero_tail_rows: bad arguments");
  FG_CHECK_ARG((x_dtype == FG_F32 || x_dtype == FG_BF16) && (out_dtype == FG_F32 || out_dtype == FG_BF16),
               "zero_tail_rows: bad dtype");
  const int esz = x_dtype == FG_F32 ? 4 : 2;
  const bool vec = d % 4 == 0 && x_ld % 4 == 0 && x_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & (4 * esz - 1)) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  const int grid = ho_grid(static_cast<int64_t>(b) * rows * (vec ? d / 4 : d));
#define FG_ZT(TI, TO)                                                                                                  \
  do {                                                                                                                 \
    if (vec)                                                                                                           \
      zero_tail_rows_kernel<TI, TO, true><<<grid, kHoThreads, 0, stream>>>(static_cast<const TI*>(x), x_ld, x_bs, len, b, \
                                                                           rows, d, static_cast<TO*>(out));            \
    else                                                                                                               \
      zero_tail_rows_kernel<TI, TO, false><<<grid, kHoThreads, 0, stream>>>(static_cast<const TI*>(x), x_ld, x_bs, len, \
                                                                            b, rows, d, static_cast<TO*>(out));        \
  } while (0)
  if (x_dtype == FG_F32 && out_dtype == FG_F32) FG_ZT(float, float);
  else if (x_dtype == FG_F32) FG_ZT(float, __nv_bfloat16);
  else if (out_dtype == FG_F32) FG_ZT(__nv_bfloat16, float);
  else FG_ZT(__nv_bfloat16, __nv_bfloat16);
#undef FG_ZT
  FG_CHECK_LAUNCH("zero_tail_rows");
  return FG_OK;
}
