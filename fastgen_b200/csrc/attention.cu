// Flash attention for sm_100a: softmax(q.k^T * scale) v per (sample, head), fp32 softmax, bf16 operands.
//   TMA stages Q once and K/V tiles in a 2-deep ring; S = Q.K^T and O += P.V run on tcgen05 with S (double
//   buffered) and O accumulators in TMEM; the softmax warps own one query row per thread (tcgen05.ld 32x32b),
//   keep running max / sum in fp32, write P as the bf16 A operand into 128B-swizzled shared memory, and only
//   rescale O when the running max grew by more than 2^8 (lazy rescale).
//   V is consumed straight from its row-major [token][head][d] storage as an MN-major B operand - no transpose pass.
//
// One kernel serves every attention of the hot path:
//   * the UNet self-attention over a packed channels-last qkv tensor (fg_attention);
//   * the Wan / DiT self-attention over separate q / k / v views, where the keys are the concatenation of up to two
//     segments - the KV-cache prefix and the current chunk - so the reference's torch.cat([k_cached, key]) per layer
//     per call (network_causal.py:372-381) never materialises;
//   * the block-causal mask of the full-sequence causal forward (network_causal.py:131-196): a query of chunk c sees
//     the keys up to the end of chunk c;
//   * cross-attention to the 512 text tokens (Tq != Tkv).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = softmax/epilogue.
//
// Replaces: fastgen/networks/EDM/network.py:160-168, 290-296 (AttentionOp + the two einsums, which materialise the
// [N*heads, T, T] weights in fp32) and diffusers' dispatch_attention_fn / F.scaled_dot_product_attention as called
// from fastgen/networks/Wan/network_causal.py:312-320, 350-358, 404-412, 446-454.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kAttnThreads = 192;
constexpr int kQRows = 128;

template <int DP, int BKV>
struct AttnCfg {
  static constexpr int DC = DP / 64;                    // 64-wide head-dim chunks
  static constexpr int Q_BYTES = DC * kQRows * 128;
  static constexpr int KV_TILE_BYTES = DC * BKV * 128;  // K (or V) of one stage
  static constexpr int STAGES = 2;
  static constexpr int P_BYTES = (BKV / 64) * kQRows * 128;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * 2 * KV_TILE_BYTES + P_BYTES + 1024 + 256;
  static constexpr int TMEM_O = 2 * BKV;
  static constexpr int TMEM_COLS_RAW = 2 * BKV + DP;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_RAW <= 512, "TMEM budget");
};

struct AttnKParams {
  CUtensorMap tmQ;     // dims (d, H, Tq, B), box (64, 1, 128, 1)
  CUtensorMap tmK[2];  // dims (d, H, len_s, B), box (64, 1, BKV, 1)
  CUtensorMap tmV[2];
  int32_t B, Tq, H, d;
  int32_t kv_len[2];
  int32_t mask_first, mask_chunk;  // block-causal mask (tokens of the first chunk / of every later chunk); 0 = none
  float scale_log2;                // log2(e) * scale
  const int32_t* kv_len0_per_sample;  // optional [B]: per-sample key count of segment 0 (prefix attention mask)
  __nv_bfloat16* out;
  int64_t out_ld, out_bs;
};

// keys a query token may see under the block-causal mask: up to the end of its own chunk
__device__ __forceinline__ int causal_limit(const AttnKParams& p, int tok) {
  if (p.mask_chunk == 0) return 0x7fffffff;
  if (tok < p.mask_first) return p.mask_first;
  return p.mask_first + ((tok - p.mask_first) / p.mask_chunk + 1) * p.mask_chunk;
}

template <int DP, int BKV>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<DP, BKV>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;  // stage s: K at s*2*KV_TILE_BYTES, V right after
  uint8_t* sP = sKV + Cfg::STAGES * 2 * Cfg::KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = (p.Tq + kQRows - 1) / kQRows;
  const int qt = blockIdx.x % q_tiles;
  const int head = (blockIdx.x / q_tiles) % p.H;
  const int img = blockIdx.x / (q_tiles * p.H);
  const int q0 = qt * kQRows;
  // key tiles: segment 0 then segment 1; under the block-causal mask (single segment) the tile loop stops at the
  // limit of the last query row of this tile
  int len0 = p.kv_len[0];
  if (p.kv_len0_per_sample) len0 = max(0, min(len0, p.kv_len0_per_sample[img]));
  if (p.mask_chunk) len0 = min(len0, causal_limit(p, min(q0 + kQRows, p.Tq) - 1));
  const int n0 = (len0 + BKV - 1) / BKV;
  const int n1 = (p.kv_len[1] + BKV - 1) / BKV;
  const int nkv = n0 + n1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK[0]);
    tma_prefetch_desc(&p.tmV[0]);
    if (n1) {
      tma_prefetch_desc(&p.tmK[1]);
      tma_prefetch_desc(&p.tmV[1]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&s_full[i], 1);
      }
      mbar_init(p_full, 4);
      mbar_init(pv_done, 1);
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
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
      for (int dc = 0; dc < Cfg::DC; ++dc) tma_load_4d(sQ + dc * kQRows * 128, &p.tmQ, q_full, dc * 64, head, q0, img);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        const int seg = j >= n0;
        const int row0 = (seg ? j - n0 : j) * BKV;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1u);
        uint8_t* sk = sKV + s * 2 * Cfg::KV_TILE_BYTES;
        uint8_t* sv = sk + Cfg::KV_TILE_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_TILE_BYTES);
        for (int dc = 0; dc < Cfg::DC; ++dc) {
          tma_load_4d(sk + dc * BKV * 128, &p.tmK[seg], &kv_full[s], dc * 64, head, row0, img);
          tma_load_4d(sv + dc * BKV * 128, &p.tmV[seg], &kv_full[s], dc * 64, head, row0, img);
        }
      }
    }
  } else if (warp == 1) {
    {  // converged issuer warp, one elected lane issues (see umma_bf16_ss_elect)
      const uint32_t elected = elect_one_sync();
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kQRows, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQRows, DP, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * Cfg::KV_TILE_BYTES);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(s * BKV);
#pragma unroll
        for (int dc = 0; dc < Cfg::DC; ++dc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_elect(elected, d_tmem, umma_desc_sw128(q_addr + dc * kQRows * 128 + k * 32),
                         umma_desc_sw128(k_addr + dc * BKV * 128 + k * 32), idesc_qk, (dc | k) != 0 ? 1u : 0u);
        umma_commit_elect(elected, &s_full[s]);
      };
      mbar_wait(q_full, 0);
      if (nkv) issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_qk(j + 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int s = j & 1;
        const uint32_t v_addr = smem_u32(sKV + s * 2 * Cfg::KV_TILE_BYTES + Cfg::KV_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_bf16_ss_elect(elected, tmem_base + Cfg::TMEM_O, umma_desc_sw128(p_addr + (kk >> 2) * kQRows * 128 + (kk & 3) * 32),
                       umma_desc_sw128(v_addr + kk * 2048, BKV * 128), idesc_pv, (j | kk) != 0 ? 1u : 0u);
        umma_commit_elect(elected, &kv_empty[s]);
        umma_commit_elect(elected, pv_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tok = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int limit = causal_limit(p, tok);
    float m_run = -INFINITY;  // running max in scaled (log2) units
    float l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) {
        uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c]);
        tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(b * BKV + c), chunk);
      }
      tmem_ld_wait();
      // columns >= kv_valid are padding (end of a segment) or masked (beyond this row's causal limit)
      const int seg = j >= n0;
      const int col0 = (seg ? j - n0 : j) * BKV;
      int kv_valid = (seg ? p.kv_len[1] : len0) - col0;
      if (p.mask_chunk) kv_valid = min(kv_valid, limit - col0);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKV; ++c) {
        float s = __uint_as_float(sv[c]) * p.scale_log2;
        s = (c < kv_valid) ? s : -INFINITY;
        sv[c] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P buffer consumed, O stable
        tc_fence_after();
        const bool need = mx > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_run;
          const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
#pragma unroll 1
          for (int c = 0; c < DP; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_O + c), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_O + c), o);
          }
          tmem_st_wait();
        }
      } else {
        m_run = mx;
      }
      // a row whose every key so far is masked keeps m_run = -inf: its probabilities are 0, not exp2(-inf + inf)
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      float lsum = 0.f;
      uint8_t* prow = sP + row * 128;
#pragma unroll
      for (int c8 = 0; c8 < BKV / 8; ++c8) {
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pv[i] = exp2f(__uint_as_float(sv[c8 * 8 + i]) - m_use);
          lsum += pv[i];
        }
        uint4 u;
        u.x = pack_bf16x2(pv[0], pv[1]), u.y = pack_bf16x2(pv[2], pv[3]);
        u.z = pack_bf16x2(pv[4], pv[5]), u.w = pack_bf16x2(pv[6], pv[7]);
        const int sub = c8 >> 3, chunk = c8 & 7;
        *reinterpret_cast<uint4*>(prow + sub * (kQRows * 128) + ((chunk ^ (row & 7)) << 4)) = u;
      }
      l_run += lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 ----
    if (nkv) {
      mbar_wait(pv_done, (nkv - 1) & 1);
      tc_fence_after();
    }
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = p.out + static_cast<int64_t>(img) * p.out_bs + static_cast<int64_t>(tok) * p.out_ld + head * p.d;
#pragma unroll 1
    for (int c = 0; c < DP; c += 32) {
      uint32_t o[32];
      if (nkv) {
        tmem_ld_32x32(tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_O + c), o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (tok < p.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (c + i < p.d) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + i) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DP, int BKV>
static int launch_attention(AttnKParams& kp, const fg_flash_attn_params* p, cudaStream_t stream) {
  using Cfg = AttnCfg<DP, BKV>;
  static PerDeviceFlag configured;
  const int cfg_dev = current_device();
  if (!configured.test(cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<DP, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("flash_attn: cudaFuncSetAttribute(%d B) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
    configured.set(cfg_dev);
  }
  {
    const uint64_t dims[4] = {(uint64_t)kp.d, (uint64_t)kp.H, (uint64_t)kp.Tq, (uint64_t)kp.B};
    const uint64_t strides[4] = {2, (uint64_t)kp.d * 2, (uint64_t)p->q_ld * 2, (uint64_t)p->q_bs * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)kQRows, 1};
    int r = encode_tmap_bf16(&kp.tmQ, p->q, 4, dims, strides, box);
    if (r != FG_OK) return r;
  }
  for (int s = 0; s < 2; ++s) {
    // an empty segment still gets a valid descriptor (of segment 0) so the kernel parameter is well formed
    const int src = kp.kv_len[s] > 0 ? s : 0;
    const uint64_t dims[4] = {(uint64_t)kp.d, (uint64_t)kp.H, (uint64_t)(kp.kv_len[src] > 0 ? kp.kv_len[src] : 1),
                              (uint64_t)kp.B};
    const uint64_t strides[4] = {2, (uint64_t)kp.d * 2, (uint64_t)p->kv_ld[src] * 2, (uint64_t)p->kv_bs[src] * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)BKV, 1};
    int r = encode_tmap_bf16(&kp.tmK[s], p->k[src], 4, dims, strides, box);
    if (r != FG_OK) return r;
    r = encode_tmap_bf16(&kp.tmV[s], p->v[src], 4, dims, strides, box);
    if (r != FG_OK) return r;
  }
  const int q_tiles = (kp.Tq + kQRows - 1) / kQRows;
  const int grid = kp.B * kp.H * q_tiles;
  attention_kernel<DP, BKV><<<grid, kAttnThreads, Cfg::SMEM_BYTES, stream>>>(kp);
  FG_CHECK_LAUNCH("flash_attn");
  return FG_OK;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int flash_attn_v2(const fg_flash_attn_params* p, float scale, cudaStream_t stream);  // attention2.cu
size_t flash_attn_v2_workspace(const fg_flash_attn_params* p);

}  // namespace fg

extern "C" int fg_flash_attn(const fg_flash_attn_params* p, fg_stream_t stream_) {
  using namespace fg;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FG_CHECK_ARG(p && p->q && p->out && p->k[0] && p->v[0], "flash_attn: null pointer");
  FG_CHECK_ARG(p->batch > 0 && p->tq > 0 && p->heads > 0 && p->d > 0, "flash_attn: empty problem");
  FG_CHECK_ARG(p->d % 8 == 0 && p->d <= 256, "flash_attn: head dim %d must be a multiple of 8 and <= 256", p->d);
  FG_CHECK_ARG(p->kv_len[0] >= 0 && p->kv_len[1] >= 0 && p->kv_len[0] + p->kv_len[1] > 0,
               "flash_attn: bad key lengths %d + %d", p->kv_len[0], p->kv_len[1]);
  const int64_t hd = static_cast<int64_t>(p->heads) * p->d;
  FG_CHECK_ARG(p->q_ld >= hd && p->q_ld % 8 == 0 && p->q_bs % 8 == 0, "flash_attn: q_ld %lld / q_bs %lld too small or unaligned",
               (long long)p->q_ld, (long long)p->q_bs);
  FG_CHECK_ARG(p->out_ld >= hd && p->out_ld % 8 == 0 && p->out_bs % 8 == 0, "flash_attn: out_ld %lld too small / unaligned",
               (long long)p->out_ld);
  FG_CHECK_ARG(al16(p->q) && al16(p->out), "flash_attn: pointers must be 16-byte aligned");
  for (int s = 0; s < 2; ++s) {
    if (p->kv_len[s] == 0) continue;
    FG_CHECK_ARG(p->k[s] && p->v[s] && al16(p->k[s]) && al16(p->v[s]), "flash_attn: segment %d k/v null or unaligned", s);
    FG_CHECK_ARG(p->kv_ld[s] >= hd && p->kv_ld[s] % 8 == 0 && p->kv_bs[s] % 8 == 0,
                 "flash_attn: segment %d kv_ld %lld / kv_bs %lld too small or unaligned", s, (long long)p->kv_ld[s],
                 (long long)p->kv_bs[s]);
  }
  FG_CHECK_ARG(p->mask_chunk >= 0 && p->mask_first >= 0, "flash_attn: negative mask extents");
  if (p->mask_chunk > 0)
    FG_CHECK_ARG(p->kv_len[1] == 0 && p->kv_len[0] == p->tq && p->mask_first > 0,
                 "flash_attn: the block-causal mask needs one key segment of the query length and mask_first > 0");
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = p->batch, kp.Tq = p->tq, kp.H = p->heads, kp.d = p->d;
  kp.kv_len[0] = p->kv_len[0], kp.kv_len[1] = p->kv_len[1];
  if (kp.kv_len[0] == 0) {  // keep segment 0 the non-empty one
    FG_CHECK_ARG(false, "flash_attn: segment 0 must be non-empty (put a lone segment first)");
  }
  kp.mask_first = p->mask_chunk > 0 ? p->mask_first : 0;
  kp.mask_chunk = p->mask_chunk;
  kp.kv_len0_per_sample = p->kv_len0_per_sample;
  if (p->kv_len0_per_sample)
    FG_CHECK_ARG((reinterpret_cast<uintptr_t>(p->kv_len0_per_sample) & 3) == 0 && p->mask_chunk == 0,
                 "flash_attn: kv_len0_per_sample must be 4-byte aligned and is not combined with the block-causal mask");
  const float scale = p->scale > 0.f ? p->scale : 1.0f / sqrtf(static_cast<float>(p->d));
  {
    // long sequences: two query tiles per CTA, P in tensor memory (attention2.cu)
    static const bool allow_v2 = !(getenv("FG_ATTN_V2") && getenv("FG_ATTN_V2")[0] == '0');
    if (allow_v2 && p->d <= 128 && p->tq > 128) return flash_attn_v2(p, scale, stream);
  }
  kp.scale_log2 = 1.4426950408889634f * scale;
  kp.out = static_cast<__nv_bfloat16*>(p->out);
  kp.out_ld = p->out_ld;
  kp.out_bs = p->out_bs;
  if (p->d <= 64) return launch_attention<64, 128>(kp, p, stream);
  if (p->d <= 128) return launch_attention<128, 128>(kp, p, stream);
  return launch_attention<256, 64>(kp, p, stream);
}

extern "C" int64_t fg_flash_attn_workspace_size(const fg_flash_attn_params* p) {
  if (!p || p->d > 128 || p->tq <= 128 || p->mask_chunk > 0) return 0;
  return static_cast<int64_t>(fg::flash_attn_v2_workspace(p));
}

// packed channels-last qkv of the UNet blocks: [N, T, 3*heads*d] with channel order (q|k|v, head, d)
extern "C" int fg_attention(const void* qkv, int64_t qkv_ld, int32_t n, int32_t tokens, int32_t heads, int32_t d,
                            void* out, int64_t out_ld, fg_stream_t stream) {
  fg::set_error("");
  if (!qkv || !out) {
    fg::set_error("attention: null pointer");
    return FG_ERR_INVALID;
  }
  if (qkv_ld < 3ll * heads * d) {
    fg::set_error("attention: qkv_ld %lld too small / unaligned", (long long)qkv_ld);
    return FG_ERR_INVALID;
  }
  fg_flash_attn_params p;
  memset(&p, 0, sizeof(p));
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  p.q = base, p.q_ld = qkv_ld, p.q_bs = qkv_ld * tokens;
  p.k[0] = base + static_cast<int64_t>(heads) * d, p.v[0] = base + 2ll * heads * d;
  p.kv_ld[0] = qkv_ld, p.kv_bs[0] = qkv_ld * tokens, p.kv_len[0] = tokens;
  p.batch = n, p.heads = heads, p.d = d, p.tq = tokens;
  p.out = out, p.out_ld = out_ld, p.out_bs = out_ld * tokens;
  return fg_flash_attn(&p, stream);
}
