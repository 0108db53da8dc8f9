// Sampler-step kernels: EDM preconditioning, x0 update, re-noising and the fp64 "one rounding" helpers of the
// noise schedules, each fused into a single vectorised streaming pass (HBM-bound).
//
// Rounding contract (what makes these bit-exact against the reference given identical inputs):
//   * per-sample coefficients arrive as fp64 device arrays computed by the host mirror of the schedule with the
//     same torch ops as the reference, and are cast to the tensor dtype exactly where the reference casts them;
//   * products/sums the reference evaluates in the tensor dtype use explicit round-to-nearest intrinsics with no
//     FMA contraction (EDM/network.py:803-805  `c_skip * x_t + c_out * out`);
//   * everything the reference evaluates in float64 (noise_schedule.py:441-449, 559-574, 992-1004) is evaluated
//     in float64 here with __dmul_rn/__dadd_rn/__ddiv_rn and rounded once to the tensor dtype.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"
#include "sampler_vec.cuh"

namespace fg {

constexpr int kEwThreads = 256;

// vectorised bodies (sampler_vec.cuh; index arithmetic and alignment checked on the host by
// tests/test_sampler_vec_host_cpu.py): 4 pixels x all channels per thread for the EDM step, 16 bytes per thread and
// iteration for the fp64-coefficient helpers
template <typename T>
__global__ void __launch_bounds__(kEwThreads) edm_step_vec_kernel(const vec::EdmStepV p) {
  vec::edm_step_vec_body<T>(p, blockIdx.x * static_cast<int64_t>(kEwThreads) + threadIdx.x,
                            static_cast<int64_t>(gridDim.x) * kEwThreads);
}
template <typename T, int MODE>
__global__ void __launch_bounds__(kEwThreads) coef_vec_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t total,
                                                              int64_t outer, int64_t mid, int64_t inner,
                                                              const double* __restrict__ c0,
                                                              const double* __restrict__ c1, T* __restrict__ out) {
  vec::coef_vec_body<T, MODE>(a, b, total, outer, mid, inner, c0, c1, out,
                              blockIdx.x * static_cast<int64_t>(kEwThreads) + threadIdx.x,
                              static_cast<int64_t>(gridDim.x) * kEwThreads);
}

// The vector kernels are the default (B200, round 2: bit-identical to the scalar forms on every tested shape, 5.97 TB/s
// for the EDM step and >= 5.1 TB/s for the fp64-coefficient helpers at 1 GB against 4.2 / 1.3 TB/s).  The scalar kernels
// below only take the shapes the vector forms cannot (a row length that is not a multiple of the vector, unaligned
// views, more than 4 channels) and stay switchable for the equality tests (fg_sampler_vec(0), FASTGEN_B200_SAMPLER_VEC=0).
static int g_sampler_vec = -1;  // -1: not yet read from the environment
static bool sampler_vec_enabled() {
  if (g_sampler_vec < 0) {
    const char* e = getenv("FASTGEN_B200_SAMPLER_VEC");
    g_sampler_vec = (e && e[0] == '0') ? 0 : 1;
  }
  return g_sampler_vec != 0;
}
static bool aligned_to(const void* p, uintptr_t bytes) { return (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) == 0; }

template <typename T>
struct Cast;
template <>
struct Cast<float> {
  static __device__ __forceinline__ float from_double(double v) { return __double2float_rn(v); }
  static __device__ __forceinline__ float to_float(float v) { return v; }
  static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <>
struct Cast<__nv_bfloat16> {
  // torch: double -> bfloat16 goes through float
  static __device__ __forceinline__ __nv_bfloat16 from_double(double v) {
    return __float2bfloat16_rn(__double2float_rn(v));
  }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Cast<double> {
  static __device__ __forceinline__ double from_double(double v) { return v; }
  static __device__ __forceinline__ float to_float(double v) { return static_cast<float>(v); }
  static __device__ __forceinline__ double from_float(float v) { return static_cast<double>(v); }
};
template <typename T>
__device__ __forceinline__ double to_double(T v) {
  return static_cast<double>(Cast<T>::to_float(v));
}
template <>
__device__ __forceinline__ double to_double<double>(double v) {
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// fg_edm_input: NCHW latent -> channels-last bf16 operand, scaled by cast(c_in[n])   (EDM/network.py:771-773)
// one thread per pixel; reads are coalesced per channel plane, the write is one 16B (c_pad = 8) store
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kEwThreads) edm_input_kernel(const T* __restrict__ x, int n, int c, int hw,
                                                               const double* __restrict__ c_in,
                                                               __nv_bfloat16* __restrict__ xin,
                                                               __nv_bfloat16* __restrict__ xin_lo, int c_pad) {
  const int64_t total = static_cast<int64_t>(n) * hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kEwThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kEwThreads) {
    const int img = static_cast<int>(i / hw);
    const int pix = static_cast<int>(i % hw);
    const float ci = c_in ? Cast<T>::to_float(Cast<T>::from_double(c_in[img])) : 1.0f;
    for (int c0 = 0; c0 < c_pad; c0 += 8) {
      float v[8], lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = c0 + k;
        float val = 0.f;
        if (ch < c) {
          const float xv = Cast<T>::to_float(x[(static_cast<int64_t>(img) * c + ch) * hw + pix]);
          // product evaluated in the tensor dtype, like `c_in * x_t`
          val = c_in ? Cast<T>::to_float(Cast<T>::from_float(__fmul_rn(ci, xv))) : xv;
        }
        v[k] = val;
        lo[k] = val - __bfloat162float(__float2bfloat16_rn(val));
      }
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]), u.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(xin + i * c_pad + c0) = u;
      if (xin_lo) {
        u.x = pack_bf16x2(lo[0], lo[1]), u.y = pack_bf16x2(lo[2], lo[3]);
        u.z = pack_bf16x2(lo[4], lo[5]), u.w = pack_bf16x2(lo[6], lo[7]);
        *reinterpret_cast<uint4*>(xin_lo + i * c_pad + c0) = u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// fg_edm_step: precond_output (+ optional ODE eps) + forward_process, one pass, one thread per pixel
// ---------------------------------------------------------------------------------------------------------
struct EdmStepK {
  const float* f;
  int64_t f_ld;
  const void* x_t;
  const void* eps;
  int n, c, hw;
  const double* c_skip;
  const double* c_out;
  int ode;
  const double* alpha_cur;
  const double* sigma_cur;
  const double* alpha_next;
  const double* sigma_next;
  void* x0;
  void* x_next;
};

template <typename T>
__global__ void __launch_bounds__(kEwThreads) edm_step_kernel(const EdmStepK p) {
  const T* __restrict__ xt = static_cast<const T*>(p.x_t);
  const T* __restrict__ eps = static_cast<const T*>(p.eps);
  T* __restrict__ x0o = static_cast<T*>(p.x0);
  T* __restrict__ xno = static_cast<T*>(p.x_next);
  const int64_t total = static_cast<int64_t>(p.n) * p.hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kEwThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kEwThreads) {
    const int img = static_cast<int>(i / p.hw);
    const int pix = static_cast<int>(i % p.hw);
    float cs = 0.f, co = 1.f;
    if (p.c_skip) {
      cs = Cast<T>::to_float(Cast<T>::from_double(p.c_skip[img]));
      co = Cast<T>::to_float(Cast<T>::from_double(p.c_out[img]));
    }
    const bool renoise = p.x_next != nullptr;
    double a_n = 0.0, s_n = 0.0, a_c = 0.0, s_c = 1.0;
    if (renoise) {
      a_n = p.alpha_next[img], s_n = p.sigma_next[img];
      if (p.ode) a_c = p.alpha_cur[img], s_c = p.sigma_cur[img];
    }
    for (int ch = 0; ch < p.c; ++ch) {
      const int64_t o = (static_cast<int64_t>(img) * p.c + ch) * p.hw + pix;
      const T fv = Cast<T>::from_float(p.f[i * p.f_ld + ch]);  // network output in the tensor dtype
      const T xv = xt[o];
      T x0v;
      if (p.c_skip) {
        const T t1 = Cast<T>::from_float(__fmul_rn(cs, Cast<T>::to_float(xv)));
        const T t2 = Cast<T>::from_float(__fmul_rn(co, Cast<T>::to_float(fv)));
        x0v = Cast<T>::from_float(__fadd_rn(Cast<T>::to_float(t1), Cast<T>::to_float(t2)));
      } else {
        x0v = fv;
      }
      x0o[o] = x0v;
      if (renoise) {
        T ev;
        if (p.ode) {
          const double num = __dsub_rn(to_double<T>(xv), __dmul_rn(to_double<T>(x0v), a_c));
          ev = Cast<T>::from_double(__ddiv_rn(num, s_c));
        } else {
          ev = eps[o];
        }
        const double r = __dadd_rn(__dmul_rn(to_double<T>(x0v), a_n), __dmul_rn(to_double<T>(ev), s_n));
        xno[o] = Cast<T>::from_double(r);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// generic fp64 helpers over [n][inner] tensors with per-sample fp64 coefficients
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MODE>  // MODE 0: a*ca + b*cb ; MODE 1: (a - b*cb)/cd
__global__ void __launch_bounds__(kEwThreads) coef_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                          int64_t total, int64_t outer, int64_t mid, int64_t inner,
                                                          const double* __restrict__ c0, const double* __restrict__ c1,
                                                          T* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kEwThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kEwThreads) {
    const int64_t img = (i / outer) * mid + (i / inner) % mid;
    double r;
    if (MODE == 0) {
      r = __dmul_rn(to_double<T>(a[i]), c0[img]);
      if (b) r = __dadd_rn(r, __dmul_rn(to_double<T>(b[i]), c1[img]));
    } else {
      r = __ddiv_rn(__dsub_rn(to_double<T>(a[i]), __dmul_rn(to_double<T>(b[i]), c0[img])), c1[img]);
    }
    out[i] = Cast<T>::from_double(r);
  }
}

// PositionalEmbedding (EDM/network.py:313-319): fp32 freqs = (1/max_positions) ** (i / (half - endpoint))
__global__ void pos_embedding_kernel(const float* __restrict__ t, int n, int channels, float max_positions,
                                     int endpoint, int swap, float* __restrict__ out, int64_t out_ld) {
  const int half = channels / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int row = idx / half, j = idx % half;
  float fr = static_cast<float>(j) / static_cast<float>(half - (endpoint ? 1 : 0));
  fr = powf(1.0f / max_positions, fr);
  const float a = t[row] * fr;
  const float cv = cosf(a), sv = sinf(a);
  float* o = out + row * out_ld;
  if (swap) {
    o[j] = sv;
    o[half + j] = cv;
  } else {
    o[j] = cv;
    o[half + j] = sv;
  }
}

// y = act(x + add) -> bf16 hi / lo (+ optional fp32 copy)
__global__ void split_bf16_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ add,
                                  int64_t add_ld, int n, int c, int act, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int64_t out_ld, float* __restrict__ y, int64_t y_ld) {
  const int64_t total = static_cast<int64_t>(n) * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / c), col = static_cast<int>(i % c);
    float v = x[row * x_ld + col];
    if (add) v += add[row * add_ld + col];
    if (act == 1) v = v / (1.0f + expf(-v));
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[row * out_ld + col] = h;
    if (lo) lo[row * out_ld + col] = __float2bfloat16_rn(v - __bfloat162float(h));
    if (y) y[row * y_ld + col] = v;
  }
}

static int ew_grid(int64_t total) {
  int64_t g = (total + kEwThreads - 1) / kEwThreads;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace fg

extern "C" int fg_sampler_vec(int enable) {
  const int prev = fg::sampler_vec_enabled() ? 1 : 0;
  if (enable >= 0) fg::g_sampler_vec = enable ? 1 : 0;
  return prev;
}

extern "C" int fg_edm_input(const void* x, int32_t x_dtype, int32_t n, int32_t c, int32_t hw, const double* c_in,
                            void* xin, void* xin_lo, int32_t c_pad, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && xin, "edm_input: null pointer");
  FG_CHECK_ARG(n > 0 && c > 0 && hw > 0 && c_pad >= c && c_pad % 8 == 0, "edm_input: bad shape n=%d c=%d hw=%d c_pad=%d",
               n, c, hw, c_pad);
  const int grid = ew_grid(static_cast<int64_t>(n) * hw);
  auto* o = static_cast<__nv_bfloat16*>(xin);
  auto* ol = static_cast<__nv_bfloat16*>(xin_lo);
  if (x_dtype == FG_F32)
    edm_input_kernel<float><<<grid, kEwThreads, 0, stream>>>(static_cast<const float*>(x), n, c, hw, c_in, o, ol, c_pad);
  else if (x_dtype == FG_BF16)
    edm_input_kernel<__nv_bfloat16><<<grid, kEwThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), n, c, hw,
                                                                     c_in, o, ol, c_pad);
  else
    FG_CHECK_ARG(false, "edm_input: unsupported dtype %d", x_dtype);
  FG_CHECK_LAUNCH("edm_input");
  return FG_OK;
}

extern "C" int fg_edm_step(const fg_edm_step_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p && p->f && p->x_t && p->x0, "edm_step: null pointer");
  FG_CHECK_ARG(p->n > 0 && p->c > 0 && p->hw > 0 && p->f_ld >= p->c, "edm_step: bad shape");
  FG_CHECK_ARG((p->c_skip == nullptr) == (p->c_out == nullptr), "edm_step: c_skip and c_out go together");
  if (p->x_next) {
    FG_CHECK_ARG(p->alpha_next && p->sigma_next, "edm_step: re-noising needs alpha_next/sigma_next");
    if (p->ode)
      FG_CHECK_ARG(p->alpha_cur && p->sigma_cur, "edm_step: ode needs alpha_cur/sigma_cur");
    else
      FG_CHECK_ARG(p->eps != nullptr, "edm_step: sde re-noising needs eps");
  }
  EdmStepK k;
  k.f = static_cast<const float*>(p->f), k.f_ld = p->f_ld, k.x_t = p->x_t, k.eps = p->eps;
  k.n = p->n, k.c = p->c, k.hw = p->hw, k.c_skip = p->c_skip, k.c_out = p->c_out, k.ode = p->ode;
  k.alpha_cur = p->alpha_cur, k.sigma_cur = p->sigma_cur, k.alpha_next = p->alpha_next, k.sigma_next = p->sigma_next;
  k.x0 = p->x0, k.x_next = p->x_next;
  if (p->x_dtype != FG_F32 && p->x_dtype != FG_BF16) FG_CHECK_ARG(false, "edm_step: unsupported dtype %d", p->x_dtype);
  const uintptr_t xa = p->x_dtype == FG_F32 ? 16 : 8;  // four elements of the latent dtype
  if (sampler_vec_enabled() && p->hw % 4 == 0 && p->c <= 4 && p->f_ld == 4 && aligned_to(p->f, 16) && aligned_to(p->x_t, xa) &&
      aligned_to(p->x0, xa) && aligned_to(p->eps, xa) && aligned_to(p->x_next, xa)) {
    vec::EdmStepV v;
    v.f = k.f, v.x_t = k.x_t, v.eps = k.eps, v.n = k.n, v.c = k.c, v.hw = k.hw, v.c_skip = k.c_skip, v.c_out = k.c_out;
    v.ode = k.ode, v.alpha_cur = k.alpha_cur, v.sigma_cur = k.sigma_cur, v.alpha_next = k.alpha_next;
    v.sigma_next = k.sigma_next, v.x0 = k.x0, v.x_next = k.x_next;
    const int vgrid = ew_grid(static_cast<int64_t>(p->n) * p->hw / 4);
    if (p->x_dtype == FG_F32)
      edm_step_vec_kernel<float><<<vgrid, kEwThreads, 0, stream>>>(v);
    else
      edm_step_vec_kernel<__nv_bfloat16><<<vgrid, kEwThreads, 0, stream>>>(v);
    FG_CHECK_LAUNCH("edm_step");
    return FG_OK;
  }
  const int grid = ew_grid(static_cast<int64_t>(p->n) * p->hw);
  if (p->x_dtype == FG_F32)
    edm_step_kernel<float><<<grid, kEwThreads, 0, stream>>>(k);
  else if (p->x_dtype == FG_BF16)
    edm_step_kernel<__nv_bfloat16><<<grid, kEwThreads, 0, stream>>>(k);
  else
    FG_CHECK_ARG(false, "edm_step: unsupported dtype %d", p->x_dtype);
  FG_CHECK_LAUNCH("edm_step");
  return FG_OK;
}

template <int MODE>
static int launch_coef(const void* a, const void* b, int32_t dtype, int64_t total, int64_t outer, int64_t mid,
                       int64_t inner, const double* c0, const double* c1, void* out, cudaStream_t stream,
                       const char* name) {
  using namespace fg;
  FG_CHECK_ARG(a && out && c0, "%s: null pointer", name);
  FG_CHECK_ARG(total > 0 && outer > 0 && mid > 0 && inner > 0 && total % outer == 0 && outer % (mid * inner) == 0,
               "%s: bad extents total=%lld outer=%lld mid=%lld inner=%lld", name, (long long)total, (long long)outer,
               (long long)mid, (long long)inner);
  if (MODE == 0) FG_CHECK_ARG((b == nullptr) || c1, "%s: b given without its coefficient", name);
  if (MODE == 1) FG_CHECK_ARG(b && c1, "%s: null pointer", name);
  const int64_t esz = dtype == FG_F32 ? 4 : dtype == FG_BF16 ? 2 : 8;
  if (sampler_vec_enabled() && (dtype == FG_F32 || dtype == FG_BF16 || dtype == FG_F64) && inner % (16 / esz) == 0 &&
      aligned_to(a, 16) && aligned_to(b, 16) && aligned_to(out, 16)) {
    const int vgrid = ew_grid(total / (16 / esz));
    if (dtype == FG_F32)
      coef_vec_kernel<float, MODE><<<vgrid, kEwThreads, 0, stream>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                                     total, outer, mid, inner, c0, c1, static_cast<float*>(out));
    else if (dtype == FG_BF16)
      coef_vec_kernel<__nv_bfloat16, MODE><<<vgrid, kEwThreads, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), total, outer, mid, inner, c0, c1,
          static_cast<__nv_bfloat16*>(out));
    else
      coef_vec_kernel<double, MODE><<<vgrid, kEwThreads, 0, stream>>>(static_cast<const double*>(a),
                                                                      static_cast<const double*>(b), total, outer, mid,
                                                                      inner, c0, c1, static_cast<double*>(out));
    FG_CHECK_LAUNCH(name);
    return FG_OK;
  }
  const int grid = ew_grid(total);
  if (dtype == FG_F32)
    coef_kernel<float, MODE><<<grid, kEwThreads, 0, stream>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                              total, outer, mid, inner, c0, c1,
                                                              static_cast<float*>(out));
  else if (dtype == FG_BF16)
    coef_kernel<__nv_bfloat16, MODE><<<grid, kEwThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(a),
                                                                      static_cast<const __nv_bfloat16*>(b), total, outer,
                                                                      mid, inner, c0, c1,
                                                                      static_cast<__nv_bfloat16*>(out));
  else if (dtype == FG_F64)
    coef_kernel<double, MODE><<<grid, kEwThreads, 0, stream>>>(static_cast<const double*>(a),
                                                               static_cast<const double*>(b), total, outer, mid, inner,
                                                               c0, c1, static_cast<double*>(out));
  else
    FG_CHECK_ARG(false, "%s: unsupported dtype %d", name, dtype);
  FG_CHECK_LAUNCH(name);
  return FG_OK;
}

extern "C" int fg_axpby(const void* a, const void* b, int32_t dtype, int64_t total, int64_t outer, int64_t mid,
                        int64_t inner, const double* ca, const double* cb, void* out, fg_stream_t stream) {
  return launch_coef<0>(a, b, dtype, total, outer, mid, inner, ca, cb, out, static_cast<cudaStream_t>(stream), "axpby");
}
extern "C" int fg_sub_div(const void* a, const void* b, int32_t dtype, int64_t total, int64_t outer, int64_t mid,
                          int64_t inner, const double* cb, const double* cd, void* out, fg_stream_t stream) {
  return launch_coef<1>(a, b, dtype, total, outer, mid, inner, cb, cd, out, static_cast<cudaStream_t>(stream),
                        "sub_div");
}

extern "C" int fg_pos_embedding(const float* t, int32_t n, int32_t channels, int32_t max_positions, int32_t endpoint,
                                int32_t swap_sin_cos, float* out, int64_t out_ld, fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(t && out && n > 0 && channels >= 2 && channels % 2 == 0 && out_ld >= channels, "pos_embedding: bad args");
  const int total = n * (channels / 2);
  pos_embedding_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      t, n, channels, static_cast<float>(max_positions), endpoint, swap_sin_cos, out, out_ld);
  FG_CHECK_LAUNCH("pos_embedding");
  return FG_OK;
}

extern "C" int fg_split_bf16(const float* x, int64_t x_ld, const float* add, int64_t add_ld, int32_t n, int32_t c,
                             int32_t act, void* hi, void* lo, int64_t out_ld, float* y_f32, int64_t y_ld,
                             fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(x && hi && n > 0 && c > 0 && x_ld >= c && out_ld >= c, "split_bf16: bad args");
  FG_CHECK_ARG(act == 0 || act == 1, "split_bf16: bad act %d", act);
  const int grid = ew_grid(static_cast<int64_t>(n) * c);
  split_bf16_kernel<<<grid, kEwThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, x_ld, add, add_ld, n, c, act, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), out_ld, y_f32,
      y_ld);
  FG_CHECK_LAUNCH("split_bf16");
  return FG_OK;
}

extern "C" int fg_zero(void* ptr, int64_t bytes, fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(ptr && bytes >= 0, "zero: bad arguments");
  if (cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream_)) != cudaSuccess) {
    set_error("zero: cudaMemsetAsync failed");
    return FG_ERR_CUDA;
  }
  count_launch();
  return FG_OK;
}
