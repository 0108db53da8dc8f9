// Vectorised bodies of the sampler-step kernels (16-byte / 8-byte accesses, several elements per thread).
//
// The thread bodies are plain functions of (params, global thread id, global thread count) so that the SAME source runs
//   * on the GPU, called from the __global__ wrappers in sampler.cu, and
//   * on the host, called in a loop over thread ids by tests/host_emu/sampler_emu.cpp (compiled with g++ by the CPU test
//     tests/test_sampler_vec_host_cpu.py), which checks the index arithmetic, tails and alignment of every access against
//     torch on the CPU -- test infrastructure only, never part of libfastgen_b200.so.
// The per-element arithmetic is the rounding contract documented at the top of sampler.cu (fp64 "one rounding" formulas,
// tensor-dtype products without FMA contraction); on the host the same operations are plain IEEE arithmetic
// (-ffp-contract=off), which is what makes the host run bit-comparable.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FG_UNROLL _Pragma("unroll")
#else
#define FG_UNROLL
#endif

#ifndef FG_EMU_ACCESS
#define FG_EMU_ACCESS(ptr, bytes, is_store) ((void)0)  // the host emulation records / checks every vector access here
#endif

namespace fg {
namespace vec {

#if defined(__CUDA_ARCH__)
FG_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
FG_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
FG_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
FG_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
FG_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
FG_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
FG_HD float d2f(double v) { return __double2float_rn(v); }
#else
FG_HD double dmul(double a, double b) { return a * b; }
FG_HD double dadd(double a, double b) { return a + b; }
FG_HD double dsub(double a, double b) { return a - b; }
FG_HD double ddiv(double a, double b) { return a / b; }
FG_HD float fmul(float a, float b) { return a * b; }
FG_HD float fadd(float a, float b) { return a + b; }
FG_HD float d2f(double v) { return static_cast<float>(v); }
#endif

template <typename T>
struct Num;
template <>
struct Num<float> {
  static FG_HD float from_double(double v) { return d2f(v); }
  static FG_HD float to_float(float v) { return v; }
  static FG_HD float from_float(float v) { return v; }
  static FG_HD double to_double(float v) { return static_cast<double>(v); }
};
template <>
struct Num<__nv_bfloat16> {
  static FG_HD __nv_bfloat16 from_double(double v) { return __float2bfloat16_rn(d2f(v)); }  // torch: via float
  static FG_HD float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static FG_HD __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
  // (an integer re-bias of exponent and mantissa instead of the two conversions was tried: 2.3 TB/s against 5.1 TB/s --
  // the special-case branch costs more than the conversion unit saves)
  static FG_HD double to_double(__nv_bfloat16 v) { return static_cast<double>(__bfloat162float(v)); }
};
template <>
struct Num<double> {
  static FG_HD double from_double(double v) { return v; }
  static FG_HD double to_double(double v) { return v; }
};

// N elements moved as one aligned access of N * sizeof(T) bytes (16 bytes, or 8 for four bf16)
template <typename T, int N>
struct alignas(sizeof(T) * N) Pack {
  T v[N];
};
template <typename T, int N>
FG_HD Pack<T, N> load_pack(const T* p) {
  FG_EMU_ACCESS(p, sizeof(T) * N, 0);
  return *reinterpret_cast<const Pack<T, N>*>(p);
}
template <typename T, int N>
FG_HD void store_pack(T* p, const Pack<T, N>& v) {
  FG_EMU_ACCESS(p, sizeof(T) * N, 1);
  *reinterpret_cast<Pack<T, N>*>(p) = v;
}

// ---------------------------------------------------------------------------------------------------------
// fg_edm_step, vector form: one thread = 4 consecutive pixels of one image, all channels.
// Preconditions (checked by the launcher, which otherwise takes the scalar kernel): hw % 4 == 0, c <= 4, f_ld == 4,
// f 16-byte aligned, x_t / eps / x0 / x_next aligned to 4 elements.
// ---------------------------------------------------------------------------------------------------------
struct EdmStepV {
  const float* f;  // [n*hw][4] fp32
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
FG_HD void edm_step_vec_body(const EdmStepV& p, int64_t tid, int64_t nthreads) {
  const T* xt = static_cast<const T*>(p.x_t);
  const T* eps = static_cast<const T*>(p.eps);
  T* x0o = static_cast<T*>(p.x0);
  T* xno = static_cast<T*>(p.x_next);
  const int64_t quads = static_cast<int64_t>(p.n) * p.hw / 4;
  const bool renoise = p.x_next != nullptr;
  for (int64_t q = tid; q < quads; q += nthreads) {
    const int64_t i = q * 4;  // first pixel (flat over [n][hw]); the quad never straddles an image: hw % 4 == 0
    const int img = static_cast<int>(i / p.hw);
    const int pix = static_cast<int>(i - static_cast<int64_t>(img) * p.hw);
    float cs = 0.f, co = 1.f;
    if (p.c_skip) {
      cs = Num<T>::to_float(Num<T>::from_double(p.c_skip[img]));
      co = Num<T>::to_float(Num<T>::from_double(p.c_out[img]));
    }
    double a_n = 0.0, s_n = 0.0, a_c = 0.0, s_c = 1.0;
    if (renoise) {
      a_n = p.alpha_next[img], s_n = p.sigma_next[img];
      if (p.ode) a_c = p.alpha_cur[img], s_c = p.sigma_cur[img];
    }
    Pack<float, 4> fr[4];  // the 4 pixels' network outputs: 64 contiguous bytes
FG_UNROLL
    for (int j = 0; j < 4; ++j) fr[j] = load_pack<float, 4>(p.f + (i + j) * 4);
    for (int ch = 0; ch < p.c; ++ch) {
      const int64_t o = (static_cast<int64_t>(img) * p.c + ch) * p.hw + pix;
      const Pack<T, 4> xv = load_pack<T, 4>(xt + o);
      Pack<T, 4> ev, x0v, xnv;
      if (renoise && !p.ode) ev = load_pack<T, 4>(eps + o);
FG_UNROLL
      for (int j = 0; j < 4; ++j) {
        const T fv = Num<T>::from_float(fr[j].v[ch]);  // network output in the tensor dtype
        if (p.c_skip) {
          const T t1 = Num<T>::from_float(fmul(cs, Num<T>::to_float(xv.v[j])));
          const T t2 = Num<T>::from_float(fmul(co, Num<T>::to_float(fv)));
          x0v.v[j] = Num<T>::from_float(fadd(Num<T>::to_float(t1), Num<T>::to_float(t2)));
        } else {
          x0v.v[j] = fv;
        }
        if (renoise) {
          T e;
          if (p.ode) {
            const double num = dsub(Num<T>::to_double(xv.v[j]), dmul(Num<T>::to_double(x0v.v[j]), a_c));
            e = Num<T>::from_double(ddiv(num, s_c));
          } else {
            e = ev.v[j];
          }
          xnv.v[j] = Num<T>::from_double(dadd(dmul(Num<T>::to_double(x0v.v[j]), a_n), dmul(Num<T>::to_double(e), s_n)));
        }
      }
      store_pack<T, 4>(x0o + o, x0v);
      if (renoise) store_pack<T, 4>(xno + o, xnv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// fg_axpby / fg_sub_div, vector form: one thread = N = 16 / sizeof(T) consecutive elements per iteration; the
// coefficient index is evaluated once per vector (inner % N == 0: a vector never straddles two coefficients).
// MODE 0: a*c0 + b*c1 (b optional); MODE 1: (a - b*c0) / c1.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MODE>
FG_HD Pack<T, 16 / sizeof(T)> coef_vec_eval(const Pack<T, 16 / sizeof(T)>& av, const Pack<T, 16 / sizeof(T)>& bv, bool has_b,
                                            double k0, double k1) {
  constexpr int N = 16 / sizeof(T);
  Pack<T, N> r;
FG_UNROLL
  for (int j = 0; j < N; ++j) {
    double x;
    if (MODE == 0) {
      x = dmul(Num<T>::to_double(av.v[j]), k0);
      if (has_b) x = dadd(x, dmul(Num<T>::to_double(bv.v[j]), k1));
    } else {
      x = ddiv(dsub(Num<T>::to_double(av.v[j]), dmul(Num<T>::to_double(bv.v[j]), k0)), k1);
    }
    r.v[j] = Num<T>::from_double(x);
  }
  return r;
}

// fp32 / fp64 (4 or 2 elements per 16-byte vector): two vectors per thread and iteration (v and v + nthreads: both
// coalesced), all four loads issued before the arithmetic -- 6.07 TB/s at the 1.6 GB asymptote (0.93 of the copy rate).
// bf16 (8 elements per vector, each through the fp64 pipe) keeps one vector per iteration: with two, the sixteen fp64
// temporaries per thread cost more occupancy than the extra loads in flight gain (3.5 against 5.1 TB/s measured).
template <typename T, int MODE>
FG_HD void coef_vec_body(const T* a, const T* b, int64_t total, int64_t outer, int64_t mid, int64_t inner, const double* c0,
                         const double* c1, T* out, int64_t tid, int64_t nthreads) {
  constexpr int N = 16 / sizeof(T);
  constexpr bool kTwo = sizeof(T) >= 4;
  const int64_t vecs = total / N;
  const bool has_b = MODE == 1 || b != nullptr;
  const int64_t step = kTwo ? 2 * nthreads : nthreads;
  for (int64_t v = tid; v < vecs; v += step) {
    const int64_t i0 = v * N, v1 = v + nthreads;
    const bool two = kTwo && v1 < vecs;
    const int64_t i1 = two ? v1 * N : i0;
    const Pack<T, N> a0 = load_pack<T, N>(a + i0);
    Pack<T, N> a1 = a0, b0 = a0, b1 = a0;
    if (two) a1 = load_pack<T, N>(a + i1);
    if (has_b) {
      b0 = load_pack<T, N>(b + i0);
      if (two) b1 = load_pack<T, N>(b + i1);
    }
    const int64_t k_0 = (i0 / outer) * mid + (i0 / inner) % mid;
    const double k00 = c0[k_0], k01 = has_b ? c1[k_0] : 0.0;
    store_pack<T, N>(out + i0, coef_vec_eval<T, MODE>(a0, b0, has_b, k00, k01));
    if (two) {
      const int64_t k_1 = (i1 / outer) * mid + (i1 / inner) % mid;
      const double k10 = c0[k_1], k11 = has_b ? c1[k_1] : 0.0;
      store_pack<T, N>(out + i1, coef_vec_eval<T, MODE>(a1, b1, has_b, k10, k11));
    }
  }
}

}  // namespace vec
}  // namespace fg
