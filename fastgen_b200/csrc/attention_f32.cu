// fp32 UNet self-attention for the split-precision ("bf16x3", fp32-parity) path of the EDM networks.
//
// The tensor-core attention kernels round q, k, v, the softmax weights and the output to bf16, which is what keeps
// `EDMPrecond(compute="bf16x3")` at ~1e-3 relative L2 of the fp32 reference while every convolution of that path already
// carries ~16 mantissa bits (profiles/x3_emulation.py: 1e-3 -> 1e-5 once the attention is evaluated in fp32).  This
// kernel evaluates softmax(q.k / sqrt(d)) v entirely in fp32 on the CUDA cores, flash-attention style (no [T,T] matrix
// in memory), and writes the result as a bf16 hi + lo pair, the operand format of the three-segment projection GEMM.
// It is the parity path, not the throughput path: the attention is 6 % of the FLOPs of the ImageNet-64 network.
//
// Replaces: fastgen/networks/EDM/network.py:160-168, 290-296 (AttentionOp: fp32 `einsum("ncq,nck->nqk")`, softmax(dim=2),
// `einsum("nqk,nck->ncq")`).
//
// One CTA = 64 queries of one (sample, head); 256 threads as a 16 x 16 grid.  Per 64-key tile:
//   S (64 x 64): thread (ty, tx) owns rows ty + 16 i, columns tx + 16 j (4 x 4), Q / K rows read as float4 along d;
//   online softmax in registers (row max through 4 xor-shuffles over the 16 tx lanes, partial row sums per thread);
//   P -> shared memory; O (64 x D): thread owns the same 4 rows and the float4 columns 4 tx + 64 jj.
// K and V share one shared-memory buffer (loaded one after the other); rows are padded by 4 floats so that the float4
// reads of 8 consecutive rows fall into distinct banks.
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kF32Threads = 256;
constexpr int kF32Tile = 64;

template <int D>
struct AttnF32Cfg {
  static constexpr int LDD = D + 4;         // padded row of Q / K / V
  static constexpr int LDP = kF32Tile + 4;  // padded row of P
  static constexpr int SMEM_BYTES = (2 * kF32Tile * LDD + kF32Tile * LDP) * 4;
};

// 64 rows x D floats of one head, rows [row0, row0 + 64) of `src` (row stride ld floats), zero beyond `rows`
template <int D>
__device__ __forceinline__ void load_tile_f32(float* dst, const float* __restrict__ src, int64_t ld, int row0, int rows) {
  constexpr int V = D / 4;  // float4 per row
  for (int i = threadIdx.x; i < kF32Tile * V; i += kF32Threads) {
    const int r = i / V, c = i - r * V;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows) v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(row0 + r) * ld) + c);
    *reinterpret_cast<float4*>(dst + r * AttnF32Cfg<D>::LDD + 4 * c) = v;
  }
}

template <int D>
__global__ void __launch_bounds__(kF32Threads) attention_f32_kernel(const float* __restrict__ qkv, int64_t qkv_ld,
                                                                    int tokens, int heads, float scale_log2e,
                                                                    __nv_bfloat16* __restrict__ out_hi,
                                                                    __nv_bfloat16* __restrict__ out_lo, int64_t out_ld) {
  using Cfg = AttnF32Cfg<D>;
  constexpr int LDD = Cfg::LDD, LDP = Cfg::LDP;
  constexpr int DV = D / 64;  // float4 column groups of O per thread
  extern __shared__ __align__(16) float smem_f32[];
  float* Qs = smem_f32;
  float* KVs = Qs + kF32Tile * LDD;
  float* Ps = KVs + kF32Tile * LDD;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int q0 = blockIdx.x * kF32Tile, head = blockIdx.y, sample = blockIdx.z;
  const float* base = qkv + static_cast<int64_t>(sample) * tokens * qkv_ld;
  const float* qp = base + head * D;
  const float* kp = base + (heads + head) * D;
  const float* vp = base + (2 * heads + head) * D;

  load_tile_f32<D>(Qs, qp, qkv_ld, q0, tokens);

  float m[4], l[4], o[4][DV][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY, l[i] = 0.f;
#pragma unroll
    for (int jj = 0; jj < DV; ++jj)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[i][jj][e] = 0.f;
  }

  for (int k0 = 0; k0 < tokens; k0 += kF32Tile) {
    __syncthreads();  // the previous tile's P.V has finished reading KVs / Ps (first pass: orders nothing but is harmless)
    load_tile_f32<D>(KVs, kp, qkv_ld, k0, tokens);
    __syncthreads();  // K (and, on the first pass, Q) visible

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < D; k += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * LDD + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(KVs + (tx + 16 * j) * LDD + k);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(a[i].x, b[j].x, s[i][j]);
          s[i][j] = fmaf(a[i].y, b[j].y, s[i][j]);
          s[i][j] = fmaf(a[i].z, b[j].z, s[i][j]);
          s[i][j] = fmaf(a[i].w, b[j].w, s[i][j]);
        }
    }

    // online softmax in base 2: s <- s * scale * log2(e); keys beyond the sequence are masked out
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx + 16 * j < tokens) ? s[i][j] * scale_log2e : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int sh = 1; sh < 16; sh <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
      const float m_new = fmaxf(m[i], mx);  // finite: key k0 + 0 is always valid, so lane tx = 0 holds a finite score
      const float alpha = exp2f(m[i] - m_new);  // exp2f(-inf) = 0 on the first tile
      m[i] = m_new;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = exp2f(s[i][j] - m_new);
        rs += p;
        Ps[(ty + 16 * i) * LDP + tx + 16 * j] = p;
      }
      l[i] = l[i] * alpha + rs;
#pragma unroll
      for (int jj = 0; jj < DV; ++jj)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[i][jj][e] *= alpha;
    }
    __syncthreads();  // every thread is done with K; P visible
    load_tile_f32<D>(KVs, vp, qkv_ld, k0, tokens);
    __syncthreads();  // V visible

#pragma unroll 2
    for (int c = 0; c < kF32Tile; c += 4) {
      float4 p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * LDP + c);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
        for (int jj = 0; jj < DV; ++jj) {
          const float4 v = *reinterpret_cast<const float4*>(KVs + (c + cc) * LDD + 4 * tx + 64 * jj);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float pv = cc == 0 ? p[i].x : cc == 1 ? p[i].y : cc == 2 ? p[i].z : p[i].w;
            o[i][jj][0] = fmaf(pv, v.x, o[i][jj][0]);
            o[i][jj][1] = fmaf(pv, v.y, o[i][jj][1]);
            o[i][jj][2] = fmaf(pv, v.z, o[i][jj][2]);
            o[i][jj][3] = fmaf(pv, v.w, o[i][jj][3]);
          }
        }
      }
    }
  }

  // row sums: the 16 tx lanes of a row hold partial sums over their own key columns
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int sh = 1; sh < 16; sh <<= 1) l[i] += __shfl_xor_sync(0xffffffffu, l[i], sh);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty + 16 * i;
    if (row >= tokens) continue;
    const float inv = 1.0f / l[i];
    const int64_t off = (static_cast<int64_t>(sample) * tokens + row) * out_ld + head * D;
#pragma unroll
    for (int jj = 0; jj < DV; ++jj) {
      float v[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = o[i][jj][e] * inv;
        lo[e] = v[e] - __bfloat162float(__float2bfloat16_rn(v[e]));
      }
      const int col = 4 * tx + 64 * jj;
      *reinterpret_cast<uint2*>(out_hi + off + col) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
      if (out_lo)
        *reinterpret_cast<uint2*>(out_lo + off + col) = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
    }
  }
}

template <int D>
static int launch_attention_f32(const float* qkv, int64_t qkv_ld, int n, int tokens, int heads, void* out_hi, void* out_lo,
                                int64_t out_ld, cudaStream_t stream) {
  using Cfg = AttnF32Cfg<D>;
  static PerDeviceFlag configured;
  const int cfg_dev = current_device();
  if (!configured.test(cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_f32: cudaFuncSetAttribute(%d bytes) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
    configured.set(cfg_dev);
  }
  const dim3 grid((tokens + kF32Tile - 1) / kF32Tile, heads, n);
  const float scale_log2e = static_cast<float>(1.4426950408889634 / sqrt(static_cast<double>(D)));
  attention_f32_kernel<D><<<grid, kF32Threads, Cfg::SMEM_BYTES, stream>>>(
      qkv, qkv_ld, tokens, heads, scale_log2e, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo),
      out_ld);
  FG_CHECK_LAUNCH("attention_f32");
  return FG_OK;
}

}  // namespace fg

extern "C" int fg_attention_f32(const float* qkv, int64_t qkv_ld, int32_t n, int32_t tokens, int32_t heads, int32_t d,
                                void* out_hi, void* out_lo, int64_t out_ld, fg_stream_t stream_) {
  using namespace fg;
  set_error("");
  FG_CHECK_ARG(qkv && out_hi, "attention_f32: null pointer");
  FG_CHECK_ARG(n > 0 && tokens > 0 && heads > 0 && heads <= 65535 && n <= 65535, "attention_f32: bad sizes n=%d tokens=%d heads=%d", n, tokens, heads);
  FG_CHECK_ARG(qkv_ld >= 3ll * heads * d && qkv_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0,
               "attention_f32: qkv_ld %lld too small or rows not 16-byte aligned", (long long)qkv_ld);
  FG_CHECK_ARG(out_ld >= static_cast<int64_t>(heads) * d && out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_lo) & 7) == 0,
               "attention_f32: out_ld %lld too small or rows not 8-byte aligned", (long long)out_ld);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (d) {
    case 64:
      return launch_attention_f32<64>(qkv, qkv_ld, n, tokens, heads, out_hi, out_lo, out_ld, stream);
    case 128:
      return launch_attention_f32<128>(qkv, qkv_ld, n, tokens, heads, out_hi, out_lo, out_ld, stream);
    case 192:
      return launch_attention_f32<192>(qkv, qkv_ld, n, tokens, heads, out_hi, out_lo, out_ld, stream);
    case 256:
      return launch_attention_f32<256>(qkv, qkv_ld, n, tokens, heads, out_hi, out_lo, out_ld, stream);
    default:
      set_error("attention_f32: head dimension %d not supported (64, 128, 192, 256)", d);
      return FG_ERR_UNSUPPORTED;
  }
}
