// HBM-bound row kernels of the transformer (Wan / DiT) blocks: one warp owns one token row, every global access is a
// 16-byte vector, a warp instruction covers 512 contiguous bytes, statistics stay in fp32 registers (true two-pass
// variance: the row is held in registers, nothing is re-read).
//
//   fg_ln_modulate   y = LayerNorm_fp32(x) * (1 + scale[g]) + shift[g]   (AdaLN, per sample or per frame)
//                    or  LayerNorm_fp32(x) * gamma + beta                 (norm2, affine)           -> bf16
//   fg_qk_norm_rope  q, k <- RoPE(RMSNorm_across_heads(q | k) * w), in place; optionally appends the roped k and the
//                    raw v to the KV cache (store_kv)                                               -> bf16
//   fg_patchify / fg_unpatchify   latent [B,C,F,H,W] <-> token rows [B*F'*H'*W', C*pt*ph*pw]
//   fg_bcast_add     out[i][g][:] = a[i][:] + b[g][:]   (scale_shift_table + timestep projection, all blocks at once)
//
// Replaces: FP32LayerNorm + the modulate / type_as chains of fastgen/networks/Wan/network.py:116-151 and
// network_causal.py:503-545, torch.nn.RMSNorm + apply_rotary_emb of network_causal.py:262-291, the KV-cache slice
// assignments of :365-368, Conv3d patch embedding / unpatchify permutes of Wan/network.py:300-301, 262-266.
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kRowThreads = 128;  // 4 warps = 4 rows per CTA

__device__ __forceinline__ void unpack8(const uint4& a, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(a.x), v[0] = f.x, v[1] = f.y;
  f = unpack_bf16x2(a.y), v[2] = f.x, v[3] = f.y;
  f = unpack_bf16x2(a.z), v[4] = f.x, v[5] = f.y;
  f = unpack_bf16x2(a.w), v[6] = f.x, v[7] = f.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]), u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (+ modulate | + affine)
// ---------------------------------------------------------------------------------------------------------
struct LnK {
  const void* x;  // bf16 or fp32 rows
  int64_t x_ld;
  int rows, d;
  float eps;
  const float* gamma;
  const float* beta;
  const float* scale;
  const float* shift;
  int64_t mod_ld;
  int rows_per_group;
  __nv_bfloat16* y;
  int64_t y_ld;
};

template <int NV, typename TX>  // NV 8-element vectors per lane: d <= NV * 256; TX = input row type
__global__ void __launch_bounds__(kRowThreads) ln_modulate_kernel(const LnK p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = p.d >> 3;
  const float inv_d = 1.0f / static_cast<float>(p.d);
  for (int row = blockIdx.x * (kRowThreads / 32) + warp; row < p.rows; row += gridDim.x * (kRowThreads / 32)) {
    const TX* xr = static_cast<const TX*>(p.x) + static_cast<int64_t>(row) * p.x_ld;
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        if constexpr (sizeof(TX) == 2) {
          unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + vi), v[i]);
        } else {
          const float4 a = __ldg(reinterpret_cast<const float4*>(xr) + 2 * vi);
          const float4 b = __ldg(reinterpret_cast<const float4*>(xr) + 2 * vi + 1);
          v[i][0] = a.x, v[i][1] = a.y, v[i][2] = a.z, v[i][3] = a.w;
          v[i][4] = b.x, v[i][5] = b.y, v[i][6] = b.z, v[i][7] = b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[i][k];
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dlt = v[i][k] - mean;
          ss = fmaf(dlt, dlt, ss);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + p.eps);
    const int g = row / p.rows_per_group;
    const float* sc = p.scale ? p.scale + static_cast<int64_t>(g) * p.mod_ld : nullptr;
    const float* sh = p.shift ? p.shift + static_cast<int64_t>(g) * p.mod_ld : nullptr;
    __nv_bfloat16* yr = p.y + static_cast<int64_t>(row) * p.y_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd;
        if (p.gamma) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * vi);
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * vi + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * vi);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * vi + 1);
          o[0] = fmaf(o[0], g0.x, b0.x), o[1] = fmaf(o[1], g0.y, b0.y), o[2] = fmaf(o[2], g0.z, b0.z);
          o[3] = fmaf(o[3], g0.w, b0.w), o[4] = fmaf(o[4], g1.x, b1.x), o[5] = fmaf(o[5], g1.y, b1.y);
          o[6] = fmaf(o[6], g1.z, b1.z), o[7] = fmaf(o[7], g1.w, b1.w);
        }
        if (sc) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc) + 2 * vi);
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(sc) + 2 * vi + 1);
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh) + 2 * vi);
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(sh) + 2 * vi + 1);
          o[0] = fmaf(o[0], 1.0f + s0.x, h0.x), o[1] = fmaf(o[1], 1.0f + s0.y, h0.y);
          o[2] = fmaf(o[2], 1.0f + s0.z, h0.z), o[3] = fmaf(o[3], 1.0f + s0.w, h0.w);
          o[4] = fmaf(o[4], 1.0f + s1.x, h1.x), o[5] = fmaf(o[5], 1.0f + s1.y, h1.y);
          o[6] = fmaf(o[6], 1.0f + s1.z, h1.z), o[7] = fmaf(o[7], 1.0f + s1.w, h1.w);
        }
        *(reinterpret_cast<uint4*>(yr) + vi) = pack8(o);
      }
    }
  }
}

// Streaming variant: one warp per row, nothing of the row kept in registers.  Pass 1 accumulates the sum and the sum of
// squares of (x - K) with K = the row's first element (shifted single-pass variance: no cancellation unless the row mean is
// far from x[0] in units of the row's spread), pass 2 re-reads the row (L1 / L2: it was just read) and writes the output.
// <= 64 registers, so 8+ CTAs (32+ warps) per SM are resident; the register-resident kernel above needs 101-255 registers (2-5 CTAs) and
// measured 0.25 (d = 5120) to 0.65 of the copy bandwidth (profiles/membound_micro_r01.txt).
template <typename TX>
__device__ __forceinline__ void ln_load8(const TX* xr, int vi, float (&v)[8]) {
  if constexpr (sizeof(TX) == 2) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + vi), v);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(xr) + 2 * vi);
    const float4 b = __ldg(reinterpret_cast<const float4*>(xr) + 2 * vi + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
}

template <typename TX>
__global__ void __launch_bounds__(kRowThreads, 8) ln_modulate_stream_kernel(const LnK p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = p.d >> 3;
  const float inv_d = 1.0f / static_cast<float>(p.d);
  for (int row = blockIdx.x * (kRowThreads / 32) + warp; row < p.rows; row += gridDim.x * (kRowThreads / 32)) {
    const TX* xr = static_cast<const TX*>(p.x) + static_cast<int64_t>(row) * p.x_ld;
    float kshift;
    if constexpr (sizeof(TX) == 2)
      kshift = __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(xr)));
    else
      kshift = __ldg(reinterpret_cast<const float*>(xr));
    float s = 0.f, ss = 0.f;
    int vi = lane;
    for (; vi + 96 < nvec; vi += 128) {  // four independent 16/32-byte loads in flight per lane
      float v0[8], v1[8], v2[8], v3[8];
      ln_load8<TX>(xr, vi, v0);
      ln_load8<TX>(xr, vi + 32, v1);
      ln_load8<TX>(xr, vi + 64, v2);
      ln_load8<TX>(xr, vi + 96, v3);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d0 = v0[k] - kshift, d1 = v1[k] - kshift, d2 = v2[k] - kshift, d3 = v3[k] - kshift;
        s += (d0 + d1) + (d2 + d3);
        ss = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, ss))));
      }
    }
    for (; vi < nvec; vi += 32) {
      float v0[8];
      ln_load8<TX>(xr, vi, v0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d0 = v0[k] - kshift;
        s += d0;
        ss = fmaf(d0, d0, ss);
      }
    }
    const float ms = warp_sum(s) * inv_d;  // mean - K
    float var = warp_sum(ss) * inv_d - ms * ms;
    var = var > 0.f ? var : 0.f;
    const float rstd = rsqrtf(var + p.eps);
    const float mean = kshift + ms;
    const int g = row / p.rows_per_group;
    const float* sc = p.scale ? p.scale + static_cast<int64_t>(g) * p.mod_ld : nullptr;
    const float* sh = p.shift ? p.shift + static_cast<int64_t>(g) * p.mod_ld : nullptr;
    __nv_bfloat16* yr = p.y + static_cast<int64_t>(row) * p.y_ld;
    for (vi = lane; vi < nvec; vi += 32) {
      float o[8];
      ln_load8<TX>(xr, vi, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (o[k] - mean) * rstd;
      if (p.gamma) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * vi);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * vi + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * vi);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * vi + 1);
        o[0] = fmaf(o[0], g0.x, b0.x), o[1] = fmaf(o[1], g0.y, b0.y), o[2] = fmaf(o[2], g0.z, b0.z);
        o[3] = fmaf(o[3], g0.w, b0.w), o[4] = fmaf(o[4], g1.x, b1.x), o[5] = fmaf(o[5], g1.y, b1.y);
        o[6] = fmaf(o[6], g1.z, b1.z), o[7] = fmaf(o[7], g1.w, b1.w);
      }
      if (sc) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc) + 2 * vi);
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(sc) + 2 * vi + 1);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh) + 2 * vi);
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(sh) + 2 * vi + 1);
        o[0] = fmaf(o[0], 1.0f + s0.x, h0.x), o[1] = fmaf(o[1], 1.0f + s0.y, h0.y);
        o[2] = fmaf(o[2], 1.0f + s0.z, h0.z), o[3] = fmaf(o[3], 1.0f + s0.w, h0.w);
        o[4] = fmaf(o[4], 1.0f + s1.x, h1.x), o[5] = fmaf(o[5], 1.0f + s1.y, h1.y);
        o[6] = fmaf(o[6], 1.0f + s1.z, h1.z), o[7] = fmaf(o[7], 1.0f + s1.w, h1.w);
      }
      *(reinterpret_cast<uint4*>(yr) + vi) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// RMSNorm across heads + rotary embedding (+ KV-cache append)
// ---------------------------------------------------------------------------------------------------------
struct QkK {
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  const __nv_bfloat16* v;
  int64_t q_ld, k_ld, v_ld;
  int rows, heads, d;
  float eps;
  const __nv_bfloat16* wq;
  const __nv_bfloat16* wk;
  const float* rope_cos;  // [max_pos][d/2]: columns [0,nt) temporal, [nt,nt+nh) height, rest width
  const float* rope_sin;
  int rope_max_pos, nt, nh;
  int rows_per_batch, ppf, pph, ppw, frame0;
  int frame_rows, tok0;    // rows of one frame held here and the in-frame index of the first (sequence-parallel shard)
  __nv_bfloat16* k_cache;  // [batch][cache_bs]: row (cache_row0 + t) * cache_ld
  __nv_bfloat16* v_cache;
  int64_t cache_ld, cache_bs;
  int cache_row0;
};

template <int NV>
__device__ __forceinline__ void norm_rope_row(const QkK& p, __nv_bfloat16* xr, const __nv_bfloat16* w, int lane,
                                              int nvec, int pf, int ph, int pw, __nv_bfloat16* dup) {
  float v[NV][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      unpack8(*(reinterpret_cast<const uint4*>(xr) + vi), v[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) ss = fmaf(v[i][k], v[i][k], ss);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(nvec * 8) + p.eps);
  const int vec_per_head = p.d >> 3;
  const int half = p.d >> 1;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float wv[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + vi), wv);
      // torch.nn.RMSNorm on bf16: normalise in fp32, round to bf16, multiply by the bf16 weight, round
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = bf16_round(bf16_round(v[i][k] * rstd) * wv[k]);
      if (p.rope_cos) {
        const int pair0 = (vi % vec_per_head) * 4;  // first of the 4 (even, odd) pairs of this vector
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pi = pair0 + j;
          const int pos = pi < p.nt ? pf : (pi < p.nt + p.nh ? ph : pw);
          const float c = __ldg(p.rope_cos + static_cast<int64_t>(pos) * half + pi);
          const float s = __ldg(p.rope_sin + static_cast<int64_t>(pos) * half + pi);
          const float x1 = o[2 * j], x2 = o[2 * j + 1];
          o[2 * j] = x1 * c - x2 * s;
          o[2 * j + 1] = x1 * s + x2 * c;
        }
      }
      const uint4 u = pack8(o);
      *(reinterpret_cast<uint4*>(xr) + vi) = u;
      if (dup) *(reinterpret_cast<uint4*>(dup) + vi) = u;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads) qk_norm_rope_kernel(const QkK p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = (p.heads * p.d) >> 3;
  for (int row = blockIdx.x * (kRowThreads / 32) + warp; row < p.rows; row += gridDim.x * (kRowThreads / 32)) {
    const int b = row / p.rows_per_batch, t = row % p.rows_per_batch;
    int pf = 0, ph = 0, pw = 0;
    if (p.rope_cos) {
      pf = min(p.frame0 + t / p.frame_rows, p.rope_max_pos - 1);  // clamp like the reference pads with the last row
      const int tf = p.tok0 + t % p.frame_rows;                     // token index inside its frame
      ph = tf / p.ppw;
      pw = tf % p.ppw;
    }
    __nv_bfloat16* kc = nullptr;
    if (p.k_cache) kc = p.k_cache + static_cast<int64_t>(b) * p.cache_bs + static_cast<int64_t>(p.cache_row0 + t) * p.cache_ld;
    if (p.q) norm_rope_row<NV>(p, p.q + static_cast<int64_t>(row) * p.q_ld, p.wq, lane, nvec, pf, ph, pw, nullptr);
    if (p.k) norm_rope_row<NV>(p, p.k + static_cast<int64_t>(row) * p.k_ld, p.wk, lane, nvec, pf, ph, pw, kc);
    if (p.v_cache) {
      const uint4* vr = reinterpret_cast<const uint4*>(p.v + static_cast<int64_t>(row) * p.v_ld);
      uint4* vc = reinterpret_cast<uint4*>(p.v_cache + static_cast<int64_t>(b) * p.cache_bs +
                                           static_cast<int64_t>(p.cache_row0 + t) * p.cache_ld);
      for (int vi = lane; vi < nvec; vi += 32) vc[vi] = vr[vi];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// patchify / unpatchify  (patch (1, ph, pw); token order (f, h', w'))
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, int B, int C, int F, int H, int W, int ph, int pw,
                                __nv_bfloat16* __restrict__ a, int64_t a_ld) {
  // one thread per (token, c, i): writes pw consecutive k entries; k = (c*ph + i)*pw + j (Conv3d weight order)
  const int Hp = H / ph, Wp = W / pw;
  const int64_t total = static_cast<int64_t>(B) * F * Hp * Wp * C * ph;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // fastest: w' (coalesced reads along W), then i, c, h', f, b
    int64_t r = idx;
    const int wp = static_cast<int>(r % Wp);
    r /= Wp;
    const int i = static_cast<int>(r % ph);
    r /= ph;
    const int c = static_cast<int>(r % C);
    r /= C;
    const int hp = static_cast<int>(r % Hp);
    r /= Hp;
    const int f = static_cast<int>(r % F);
    const int b = static_cast<int>(r / F);
    const T* src = x + (((static_cast<int64_t>(b) * C + c) * F + f) * H + hp * ph + i) * W + wp * pw;
    const int64_t tok = ((static_cast<int64_t>(b) * F + f) * Hp + hp) * Wp + wp;
    __nv_bfloat16* dst = a + tok * a_ld + (c * ph + i) * pw;
    for (int j = 0; j < pw; ++j) dst[j] = __float2bfloat16_rn(static_cast<float>(src[j]));
  }
}

template <typename T>
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ y, int64_t y_ld, int B, int C, int F, int H, int W,
                                  int ph, int pw, T* __restrict__ out) {
  // out[b, c, f, h, w] = y[token(b, f, h/ph, w/pw)][((h%ph)*pw + (w%pw))*C + c]   (reference permute (0,7,1,4,2,5,3,6))
  const int Hp = H / ph, Wp = W / pw;
  const int64_t total = static_cast<int64_t>(B) * C * F * H * W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = idx;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H);
    r /= H;
    const int f = static_cast<int>(r % F);
    r /= F;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const int64_t tok = ((static_cast<int64_t>(b) * F + f) * Hp + h / ph) * Wp + w / pw;
    const float v = __bfloat162float(y[tok * y_ld + ((h % ph) * pw + (w % pw)) * C + c]);
    out[idx] = static_cast<T>(v);
  }
}
template <>
__global__ void unpatchify_kernel<__nv_bfloat16>(const __nv_bfloat16* __restrict__ y, int64_t y_ld, int B, int C, int F,
                                                 int H, int W, int ph, int pw, __nv_bfloat16* __restrict__ out) {
  const int Hp = H / ph, Wp = W / pw;
  const int64_t total = static_cast<int64_t>(B) * C * F * H * W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = idx;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H);
    r /= H;
    const int f = static_cast<int>(r % F);
    r /= F;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const int64_t tok = ((static_cast<int64_t>(b) * F + f) * Hp + h / ph) * Wp + w / pw;
    out[idx] = y[tok * y_ld + ((h % ph) * pw + (w % pw)) * C + c];
  }
}

// out[(i0, i1, i2, i3)] <- in[(i0, i1, i2, i3)] for a 4-D index space, 16-byte vectors, arbitrary strides on both sides
struct CopyK {
  const uint4* src;
  uint4* dst;
  int n[4];
  int64_t ss[4], ds[4];  // strides in 16-byte units
  int vec;               // 16-byte vectors per innermost item
};
__global__ void strided_copy_kernel(const CopyK p) {
  const int64_t items = static_cast<int64_t>(p.n[0]) * p.n[1] * p.n[2] * p.n[3];
  const int64_t total = items * p.vec;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % p.vec);
    int64_t r = idx / p.vec;
    const int i3 = static_cast<int>(r % p.n[3]);
    r /= p.n[3];
    const int i2 = static_cast<int>(r % p.n[2]);
    r /= p.n[2];
    const int i1 = static_cast<int>(r % p.n[1]);
    const int i0 = static_cast<int>(r / p.n[1]);
    p.dst[i0 * p.ds[0] + i1 * p.ds[1] + i2 * p.ds[2] + i3 * p.ds[3] + v] =
        p.src[i0 * p.ss[0] + i1 * p.ss[1] + i2 * p.ss[2] + i3 * p.ss[3] + v];
  }
}

// GEGLU: y[r][j] = h[r][j] * gelu(h[r][inner + j]) with the exact (erf) GELU of torch.nn.functional.gelu
__global__ void __launch_bounds__(256) geglu_kernel(const __nv_bfloat16* __restrict__ h, int64_t h_ld, int rows,
                                                    int inner, __nv_bfloat16* __restrict__ y, int64_t y_ld) {
  const int vec_per_row = inner >> 3;
  const int64_t total = static_cast<int64_t>(rows) * vec_per_row;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / vec_per_row;
    const int v = static_cast<int>(idx % vec_per_row);
    const __nv_bfloat16* hr = h + r * h_ld;
    float a[8], g[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(hr) + v), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(hr + inner) + v), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // diffusers GEGLU: hidden * gelu(gate), the gelu evaluated and rounded in the activation dtype (bf16)
      const float ge = bf16_round(0.5f * g[k] * (1.0f + erff(g[k] * 0.70710678118654752f)));
      o[k] = a[k] * ge;
    }
    *(reinterpret_cast<uint4*>(y + r * y_ld) + v) = pack8(o);
  }
}

__global__ void bcast_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int na, int nb, int d,
                                 float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(na) * nb * d;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d);
    const int g = static_cast<int>((idx / d) % nb);
    const int i = static_cast<int>(idx / (static_cast<int64_t>(d) * nb));
    out[idx] = a[static_cast<int64_t>(i) * d + c] + b[static_cast<int64_t>(g) * d + c];
  }
}

static int row_grid(int rows) {
  int64_t g = (static_cast<int64_t>(rows) + (kRowThreads / 32) - 1) / (kRowThreads / 32);
  const int64_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}
static int ew_grid2(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fg

extern "C" int fg_ln_modulate(const fg_ln_mod_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p && p->x && p->y, "ln_modulate: null pointer");
  FG_CHECK_ARG(p->rows > 0 && p->d > 0 && p->d % 8 == 0 && p->d <= 8192, "ln_modulate: bad shape rows=%d d=%d (d %% 8 == 0, <= 8192)",
               p->rows, p->d);
  FG_CHECK_ARG(p->x_ld >= p->d && p->x_ld % 8 == 0 && p->y_ld >= p->d && p->y_ld % 8 == 0 && al16(p->x) && al16(p->y),
               "ln_modulate: rows must be 16-byte aligned (ld %% 8 == 0)");
  FG_CHECK_ARG((p->gamma == nullptr) == (p->beta == nullptr), "ln_modulate: gamma and beta go together");
  FG_CHECK_ARG((p->scale == nullptr) == (p->shift == nullptr), "ln_modulate: scale and shift go together");
  if (p->scale)
    FG_CHECK_ARG(p->rows_per_group > 0 && p->mod_ld >= p->d && p->mod_ld % 4 == 0 && al16(p->scale) && al16(p->shift),
                 "ln_modulate: modulation needs rows_per_group > 0, mod_ld >= d, 16-byte aligned rows");
  if (p->gamma) FG_CHECK_ARG(al16(p->gamma) && al16(p->beta), "ln_modulate: gamma/beta must be 16-byte aligned");
  LnK k;
  FG_CHECK_ARG(p->x_dtype == FG_BF16 || p->x_dtype == FG_F32, "ln_modulate: x_dtype %d must be FG_BF16 or FG_F32", p->x_dtype);
  k.x = p->x, k.x_ld = p->x_ld, k.rows = p->rows, k.d = p->d, k.eps = p->eps;
  k.gamma = p->gamma, k.beta = p->beta, k.scale = p->scale, k.shift = p->shift, k.mod_ld = p->mod_ld;
  k.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : p->rows;
  k.y = static_cast<__nv_bfloat16*>(p->y), k.y_ld = p->y_ld;
  const int grid = row_grid(p->rows);
  const int nv = (p->d / 8 + 31) / 32;
  // rows wider than 1536 elements stream (the register-resident kernel would need 168-255 registers); FG_LN_STREAM=0|1
  // forces one kernel for A/B runs
  static const int stream_mode = getenv("FG_LN_STREAM") ? atoi(getenv("FG_LN_STREAM")) : -1;
  const bool use_stream = stream_mode >= 0 ? stream_mode != 0 : true;
  if (use_stream) {
    const int sgrid = static_cast<int>(std::min<int64_t>((static_cast<int64_t>(p->rows) + 3) / 4, 148 * 8 * 4));
    if (p->x_dtype == FG_BF16)
      ln_modulate_stream_kernel<__nv_bfloat16><<<sgrid, kRowThreads, 0, stream>>>(k);
    else
      ln_modulate_stream_kernel<float><<<sgrid, kRowThreads, 0, stream>>>(k);
    FG_CHECK_LAUNCH("ln_modulate");
    return FG_OK;
  }
#define FG_LN_LAUNCH(TX)                                                \
  do {                                                                  \
    if (nv <= 6)                                                        \
      ln_modulate_kernel<6, TX><<<grid, kRowThreads, 0, stream>>>(k);   \
    else if (nv <= 12)                                                  \
      ln_modulate_kernel<12, TX><<<grid, kRowThreads, 0, stream>>>(k);  \
    else if (nv <= 20)                                                  \
      ln_modulate_kernel<20, TX><<<grid, kRowThreads, 0, stream>>>(k);  \
    else                                                                \
      ln_modulate_kernel<32, TX><<<grid, kRowThreads, 0, stream>>>(k);  \
  } while (0)
  if (p->x_dtype == FG_BF16)
    FG_LN_LAUNCH(__nv_bfloat16);
  else
    FG_LN_LAUNCH(float);
#undef FG_LN_LAUNCH
  FG_CHECK_LAUNCH("ln_modulate");
  return FG_OK;
}

extern "C" int fg_qk_norm_rope(const fg_qk_norm_rope_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p && (p->q || p->k), "qk_norm_rope: null pointer");
  const int dm = p->heads * p->d;
  FG_CHECK_ARG(p->rows > 0 && p->heads > 0 && p->d > 0 && p->d % 8 == 0 && dm <= 8192, "qk_norm_rope: bad shape");
  if (p->q) FG_CHECK_ARG(p->wq && p->q_ld >= dm && p->q_ld % 8 == 0 && al16(p->q) && al16(p->wq), "qk_norm_rope: bad q");
  if (p->k) FG_CHECK_ARG(p->wk && p->k_ld >= dm && p->k_ld % 8 == 0 && al16(p->k) && al16(p->wk), "qk_norm_rope: bad k");
  FG_CHECK_ARG(p->rows_per_batch > 0, "qk_norm_rope: rows_per_batch must be positive");
  if (p->rope_cos) {
    FG_CHECK_ARG(p->rope_sin && p->rope_max_pos > 0 && p->pph > 0 && p->ppw > 0 && p->nt >= 0 && p->nh >= 0 &&
                     p->nt + p->nh <= p->d / 2 && p->frame0 >= 0,
                 "qk_norm_rope: bad rotary table description");
    FG_CHECK_ARG(p->pph <= p->rope_max_pos && p->ppw <= p->rope_max_pos, "qk_norm_rope: grid exceeds the rotary table");
    FG_CHECK_ARG(p->frame_rows >= 0 && p->tok0 >= 0 &&
                     p->tok0 + (p->frame_rows > 0 ? p->frame_rows : p->pph * p->ppw) <= p->pph * p->ppw,
                 "qk_norm_rope: frame shard [tok0, tok0 + frame_rows) exceeds the frame");
  }
  if (p->k_cache || p->v_cache) {
    FG_CHECK_ARG(p->k && p->k_cache && p->v_cache && p->v && p->v_ld >= dm && p->v_ld % 8 == 0 && al16(p->v) &&
                     al16(p->k_cache) && al16(p->v_cache) && p->cache_ld >= dm && p->cache_ld % 8 == 0 &&
                     p->cache_bs % 8 == 0 && p->cache_row0 >= 0,
                 "qk_norm_rope: bad KV-cache description");
  }
  QkK k;
  k.q = static_cast<__nv_bfloat16*>(p->q), k.k = static_cast<__nv_bfloat16*>(p->k);
  k.v = static_cast<const __nv_bfloat16*>(p->v);
  k.q_ld = p->q_ld, k.k_ld = p->k_ld, k.v_ld = p->v_ld;
  k.rows = p->rows, k.heads = p->heads, k.d = p->d, k.eps = p->eps;
  k.wq = static_cast<const __nv_bfloat16*>(p->wq), k.wk = static_cast<const __nv_bfloat16*>(p->wk);
  k.rope_cos = p->rope_cos, k.rope_sin = p->rope_sin, k.rope_max_pos = p->rope_max_pos, k.nt = p->nt, k.nh = p->nh;
  k.rows_per_batch = p->rows_per_batch, k.ppf = p->ppf, k.pph = p->pph, k.ppw = p->ppw, k.frame0 = p->frame0;
  k.frame_rows = p->frame_rows > 0 ? p->frame_rows : p->pph * p->ppw;
  k.tok0 = p->tok0;
  k.k_cache = static_cast<__nv_bfloat16*>(p->k_cache), k.v_cache = static_cast<__nv_bfloat16*>(p->v_cache);
  k.cache_ld = p->cache_ld, k.cache_bs = p->cache_bs, k.cache_row0 = p->cache_row0;
  const int grid = row_grid(p->rows);
  const int nv = (dm / 8 + 31) / 32;
  if (nv <= 6)
    qk_norm_rope_kernel<6><<<grid, kRowThreads, 0, stream>>>(k);
  else if (nv <= 12)
    qk_norm_rope_kernel<12><<<grid, kRowThreads, 0, stream>>>(k);
  else if (nv <= 20)
    qk_norm_rope_kernel<20><<<grid, kRowThreads, 0, stream>>>(k);
  else
    qk_norm_rope_kernel<32><<<grid, kRowThreads, 0, stream>>>(k);
  FG_CHECK_LAUNCH("qk_norm_rope");
  return FG_OK;
}

extern "C" int fg_patchify(const void* x, int32_t x_dtype, int32_t b, int32_t c, int32_t f, int32_t h, int32_t w,
                           int32_t ph, int32_t pw, void* a, int64_t a_ld, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(x && a, "patchify: null pointer");
  FG_CHECK_ARG(b > 0 && c > 0 && f > 0 && h > 0 && w > 0 && ph > 0 && pw > 0 && h % ph == 0 && w % pw == 0,
               "patchify: bad shape");
  FG_CHECK_ARG(a_ld >= static_cast<int64_t>(c) * ph * pw, "patchify: a_ld too small");
  const int64_t total = static_cast<int64_t>(b) * f * (h / ph) * (w / pw) * c * ph;
  const int grid = ew_grid2(total);
  auto* o = static_cast<__nv_bfloat16*>(a);
  if (x_dtype == FG_F32)
    patchify_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), b, c, f, h, w, ph, pw, o, a_ld);
  else if (x_dtype == FG_BF16)
    patchify_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), b, c, f, h, w, ph, pw,
                                                             o, a_ld);
  else
    FG_CHECK_ARG(false, "patchify: unsupported dtype %d", x_dtype);
  FG_CHECK_LAUNCH("patchify");
  return FG_OK;
}

extern "C" int fg_unpatchify(const void* y, int64_t y_ld, int32_t b, int32_t c, int32_t f, int32_t h, int32_t w,
                             int32_t ph, int32_t pw, void* out, int32_t out_dtype, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(y && out, "unpatchify: null pointer");
  FG_CHECK_ARG(b > 0 && c > 0 && f > 0 && h > 0 && w > 0 && ph > 0 && pw > 0 && h % ph == 0 && w % pw == 0,
               "unpatchify: bad shape");
  FG_CHECK_ARG(y_ld >= static_cast<int64_t>(c) * ph * pw, "unpatchify: y_ld too small");
  const int64_t total = static_cast<int64_t>(b) * c * f * h * w;
  const int grid = ew_grid2(total);
  auto* src = static_cast<const __nv_bfloat16*>(y);
  if (out_dtype == FG_F32)
    unpatchify_kernel<float><<<grid, 256, 0, stream>>>(src, y_ld, b, c, f, h, w, ph, pw, static_cast<float*>(out));
  else if (out_dtype == FG_BF16)
    unpatchify_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, y_ld, b, c, f, h, w, ph, pw,
                                                               static_cast<__nv_bfloat16*>(out));
  else
    FG_CHECK_ARG(false, "unpatchify: unsupported dtype %d", out_dtype);
  FG_CHECK_LAUNCH("unpatchify");
  return FG_OK;
}

extern "C" int fg_bcast_add(const float* a, const float* b, int32_t na, int32_t nb, int32_t d, float* out,
                            fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(a && b && out && na > 0 && nb > 0 && d > 0, "bcast_add: bad arguments");
  const int grid = ew_grid2(static_cast<int64_t>(na) * nb * d);
  bcast_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(a, b, na, nb, d, out);
  FG_CHECK_LAUNCH("bcast_add");
  return FG_OK;
}

extern "C" int fg_strided_copy(const void* src, void* dst, const int32_t* n, const int64_t* src_strides,
                               const int64_t* dst_strides, int32_t item_bytes, fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(src && dst && n && src_strides && dst_strides, "strided_copy: null pointer");
  FG_CHECK_ARG(item_bytes > 0 && item_bytes % 16 == 0 && al16(src) && al16(dst), "strided_copy: items must be 16-byte multiples, 16-byte aligned");
  CopyK k;
  k.src = static_cast<const uint4*>(src), k.dst = static_cast<uint4*>(dst), k.vec = item_bytes / 16;
  int64_t items = 1;
  for (int i = 0; i < 4; ++i) {
    FG_CHECK_ARG(n[i] > 0 && src_strides[i] % 16 == 0 && dst_strides[i] % 16 == 0, "strided_copy: extents must be positive, strides (bytes) multiples of 16");
    k.n[i] = n[i], k.ss[i] = src_strides[i] / 16, k.ds[i] = dst_strides[i] / 16;
    items *= n[i];
  }
  strided_copy_kernel<<<ew_grid2(items * k.vec), 256, 0, static_cast<cudaStream_t>(stream_)>>>(k);
  FG_CHECK_LAUNCH("strided_copy");
  return FG_OK;
}

extern "C" int fg_geglu(const void* h, int64_t h_ld, int32_t rows, int32_t inner, void* y, int64_t y_ld,
                        fg_stream_t stream_) {
  using namespace fg;
  FG_CHECK_ARG(h && y && rows > 0 && inner > 0 && inner % 8 == 0, "geglu: bad arguments (inner %% 8 == 0)");
  FG_CHECK_ARG(h_ld >= 2 * inner && h_ld % 8 == 0 && y_ld >= inner && y_ld % 8 == 0 && al16(h) && al16(y),
               "geglu: rows must be 16-byte aligned, h_ld >= 2 * inner");
  geglu_kernel<<<ew_grid2(static_cast<int64_t>(rows) * (inner / 8)), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(h), h_ld, rows, inner, static_cast<__nv_bfloat16*>(y), y_ld);
  FG_CHECK_LAUNCH("geglu");
  return FG_OK;
}
