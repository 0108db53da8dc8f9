// Flash attention, second generation, for long sequences (head dim <= 128, more than 128 queries):
//   * one CTA owns TWO 128-row query tiles; two softmax warpgroups (one per tile) alternate with the single MMA
//     issuer so that the tensor pipe computes S = Q.K^T / O += P.V of one tile while the other tile's softmax runs;
//   * S (fp32) and O (fp32) accumulators live in TMEM; P is written back as bf16 OVER the S columns by the softmax
//     threads (tcgen05.st) and consumed by the P.V product straight from tensor memory (A operand in TMEM) -- no
//     shared-memory round trip for P;
//   * exp2: the log2(e)/sqrt(d) scale is folded into one FFMA per element, half of the exponentials run on the MUFU
//     unit and half as a degree-3 polynomial on the FMA pipe (Cody-Waite range reduction), since at 16 MUFU/clk/SM
//     the exponential, not the tensor pipe, would otherwise bound the kernel;
//   * O is rescaled lazily (only when the running max grows by more than 2^8);
//   * O leaves through 128B-swizzled shared memory (the Q buffers, free by then) and TMA stores.
// Same arguments and semantics as the first-generation kernel in attention.cu (two key segments, block-causal mask,
// ragged lengths); fg_flash_attn picks between them.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = softmax of query
// tile 0, warps 6..9 = softmax of query tile 1 (TMEM lane quarter = warp_idx % 4).
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace fg {

constexpr int kA2BKV = 128;

#ifdef FG_ATTN_TIMING
// Experiment build only (profiles/experiments/attn_phase_timing.py): clocks per softmax phase of one CTA, summed over its
// key tiles -- [tile t][wait S, TMEM read, max (+ rescale), exp + P stores, store wait + publish, tiles].
__device__ unsigned long long g_attn_timing[2][8];
// [tile][key tile 2..5][P published, issuer saw it, issuer done issuing, next S seen by the softmax warp] (SM clock)
__device__ long long g_attn_trace[2][4][4];
#define FG_TSTAMP(var) const long long var = clock64()
#define FG_TSET(var) var = clock64()
#else
#define FG_TSTAMP(var)
#define FG_TSET(var)
#endif

template <int DP>
struct Attn2Cfg {
  static constexpr int DC = DP / 64;
  static constexpr int Q_TILE_BYTES = DC * 128 * 128;
  static constexpr int KV_TILE_BYTES = DC * kA2BKV * 128;
  static constexpr int STAGES = DP == 128 ? 2 : 4;
  static constexpr int XCH_BYTES = 3 * 2 * 2 * 128 * 4;  // WG == 2: row max (two parities) and row sum of the column halves
  // BX > 0 (row sum taken by the P.V product): a 16-key slice of a "ones atom" -- 16 rows x 128 B, 1.0 in column 0 --
  // that every P.V MMA reads as V's columns [64, 80) through its descriptor's leading-dimension offset
  static constexpr int ONES_BYTES = 2048;
  static constexpr int SMEM_BYTES = 2 * Q_TILE_BYTES + STAGES * 2 * KV_TILE_BYTES + ONES_BYTES + 1024 + 256 + XCH_BYTES;
  static constexpr int TMEM_S = 0;        // S0 at 0, S1 at 128 (P aliases the first 64 columns of each)
  static constexpr int TMEM_O = 256;      // O0 at 256, O1 at 256 + DP
  static constexpr int TMEM_COLS = 512;
  static_assert(256 + 2 * DP <= 512, "TMEM budget");
};

struct Attn2KParams {
  CUtensorMap tmQ;     // dims (d, H, Tq, B), box (64, 1, 128, 1)
  CUtensorMap tmK[2];  // dims (d, H, len_s, B), box (64, 1, 128, 1)
  CUtensorMap tmV[2];
  CUtensorMap tmO;     // dims (d, H, Tq, B) of the output, box (64, 1, 128, 1)
  int32_t B, Tq, H, d;
  int32_t kv_len[2];
  int32_t mask_first, mask_chunk;
  float scale_log2;
  // split-KV: CTA (.., split) handles an equal share of the key tiles and writes a partial result -- its own
  // normalised O (bf16, through tmO, which then maps the workspace [B*H][Tq][nsplit][DP]) and (max, sum) per row --
  // that attn2_merge_kernel combines.  nsplit == 1: the final output is written directly.
  int32_t nsplit;
  float2* part_ml;  // [B*H][Tq][nsplit]
  const int32_t* kv_len0_per_sample;  // optional [B]: per-sample key count of segment 0 (prefix attention mask)
};

__device__ __forceinline__ int causal_limit2(const Attn2KParams& p, int tok) {
  if (p.mask_chunk == 0) return 0x7fffffff;
  if (tok < p.mask_first) return p.mask_first;
  return p.mask_first + ((tok - p.mask_first) / p.mask_chunk + 1) * p.mask_chunk;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2) and the 3-input max of sm_100: the softmax inner loop is bound by
//      instruction issue and by the 16/clk MUFU unit, so every pair of elements shares its FMA-pipe instructions ----
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float exp2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair in ONE MUFU issue: the arguments are rounded to bf16 first and the result arrives packed as bf16x2 -- the
// format P is stored in (profiles/attn_r01.md: what that rounding costs in accuracy)
__device__ __forceinline__ uint32_t exp2_bf16x2(float x0, float x1) {
  uint32_t xb, y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xb) : "f"(x1), "f"(x0));
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(xb));
  return y;
}
// 2^x for a pair, on the FMA pipe: x = n + f with f in [-0.5, 0.5] (adding 1.5 * 2^23 rounds to the nearest integer
// and leaves n in the low mantissa bits), 2^f by a degree-3 polynomial (max relative error 1.8e-4; the result is
// rounded to bf16 anyway), 2^n by adding n << 23 to the exponent field -- (bits(x + magic) << 23) is exactly n << 23.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t xr = fadd2(x2, magic);
  const uint64_t tr = fadd2(xr, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(tr, pack2(-1.0f, -1.0f), x2);
  uint64_t pl = ffma2(f, pack2(0.05460262f, 0.05460262f), pack2(0.24192413f, 0.24192413f));
  pl = ffma2(pl, f, pack2(0.69331648f, 0.69331648f));
  pl = ffma2(pl, f, pack2(1.0f, 1.0f));
  float p0, p1, r0, r1;
  unpack2(pl, p0, p1);
  unpack2(xr, r0, r1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// HALF (experiment, off by default -- see flash_attn_v2): the key tile is processed as two 64-key halves with their
// own S/P columns and barriers.  While the softmax warpgroup of a query tile works on one half, the tensor pipe runs
// O += P.V of the other half and the next S = Q.K^T into the columns that product has released, so the softmax does
// not idle for the MMA round trip between key tiles (23 % of its time in the ncu capture of the full-tile pipeline).
//
// WG (softmax warps per TMEM lane quarter and query tile): with WG == 1 only two softmax warps share a scheduler and the
// inner loop is bound by dependent-issue latency (issue slots 37-40 % used, no pipe saturated: profiles/attn_r01.md).
// WG == 2 runs 16 softmax warps (576 threads): the two warps of a lane quarter split the 128 keys of an S tile into
// column halves (a warp may only touch its own 32 TMEM lanes, so the split has to be along the columns), exchange their
// row maxima through shared memory once per key tile (64-thread named barrier per quarter pair), keep partial row sums
// and share the O columns in the rare rescale and in the epilogue.
//
// BX (d <= 64 only; pairs out of every 8 whose exponential is the packed bf16 MUFU form): the softmax warps of the d = 64
// problems are bound by their own instruction stream (profiles/attn_phase_timing_r02.txt: 1 850 of the 3 270 clocks of a key
// tile are the exp + pack loop), so (a) those pairs cost FFMA2 + CVT + MUFU for two keys and (b) the row sum leaves the
// warps altogether: V gets a 65th column of ones (read from a constant shared-memory slice through the descriptor's
// leading-dimension offset, N = 80), so column 64 of the O accumulator IS the row sum of the bf16-rounded weights -- the
// same weights the numerator uses -- and the lazy rescale covers it like any other O column.
//
// PP (ping-pong): the softmax warps of the two query tiles that share a scheduler (warps w and w + 4) pass a token through
// a pair of 64-thread named barriers so that their exp loops never overlap: the 16 / clk MUFU unit is what the two loops
// compete for, and a warp that has it to itself finishes in half the time while its partner waits for S, reads it and
// takes the row max.
//
// TP (two passes over S): the one-pass loop keeps all 128 scores of a row in registers from the TMEM read to the last
// exponential (~170 live registers, small spills, little room for the scheduler to overlap independent pairs).  TP reads S
// from tensor memory twice in 32-column chunks -- once for the row max, once for the exponentials -- double-buffered, so
// the loops run at ~70 live registers; a TMEM read is ~40 clocks per chunk and overlaps the arithmetic of the previous one.
//
// SEP (d <= 64: P in tensor-memory columns of its own).  With P written over S, S of key tile j + 1 can only be computed
// after P.V of tile j has consumed P, so every softmax warp waits an MMA round trip (~720 of the ~3 270 clocks of a key tile
// at d = 64: profiles/attn_phase_timing_r02.txt) between publishing P_j and seeing S_{j+1}.  At d <= 64 tensor memory has
// room for S0 S1 | O0 O1 | P0 P1 (128 + 128 + 64 + 64 + 64 + 64 = 512 columns): the softmax warps hand S back as soon as
// they have READ it (s_free), the issuer computes S_{j+1} under the softmax of tile j, and P.V_j follows when P_j is
// published -- the next S is already there.
template <int DP, int POLY, int HALF, int WG = 1, int BX = 0, int PP = 0, int TP = 0, int SEP = 0>  // POLY: pairs out of every 8 whose exponential runs on the FMA pipe
__global__ void __launch_bounds__(64 + 256 * WG, 1) attention2_kernel(const __grid_constant__ Attn2KParams p) {
  static_assert(SEP == 0 || (DP == 64 && HALF == 0 && WG == 1 && BX == 0 && TP == 0), "separate P columns: d <= 64, one-pass loop");
  constexpr int TMEM_P = 384;   // SEP: P of query tile t at TMEM_P + 64 t
  static_assert(TP == 0 || (HALF == 0 && WG == 1), "two-pass softmax: full tiles, WG = 1");
  static_assert(PP == 0 || (HALF == 0 && WG == 1), "ping-pong: full tiles, WG = 1");
  static_assert(WG == 1 || (WG == 2 && HALF == 0), "two warps per lane quarter: full-tile pipeline only");
  static_assert(BX == 0 || (DP == 64 && HALF == 0 && WG == 1 && POLY + BX <= 8 && BX >= -1), "ones column: d <= 64, full tiles, WG = 1");
  using Cfg = Attn2Cfg<DP>;
  constexpr bool ONES = BX != 0;           // BX = -1: the ones column alone (every exponential in fp32)
  constexpr int OW = DP + (ONES ? 16 : 0);   // accumulator columns per query tile (O, then the row sum)
  static_assert(Cfg::TMEM_O + 2 * OW <= 512, "TMEM budget");
  constexpr int BKV = kA2BKV;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // tile t at t * Q_TILE_BYTES
  uint8_t* sKV = sQ + 2 * Cfg::Q_TILE_BYTES;        // stage s: K at s*2*KV_TILE_BYTES, V right after
  uint8_t* sOnes = sKV + S * 2 * Cfg::KV_TILE_BYTES;   // 1024-aligned like every tile in front of it
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::ONES_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;        // [S]
  uint64_t* kv_empty = kv_full + S;    // [S]
  uint64_t* v_full = kv_empty + S;     // [S]  half-tile pipeline: K and V slots are filled / released separately
  uint64_t* v_empty = v_full + S;      // [S]  (kv_full / kv_empty then track K alone)
  uint64_t* s_full = v_empty + S;      // [2 tiles][2 halves] (index t * 2 + h; the full-tile pipeline uses h = 0)
  uint64_t* p_full = s_full + 4;       // [2][2]
  uint64_t* pv_done = p_full + 4;      // [2][2]
  uint64_t* s_free = pv_done + 4;      // [2 tiles] SEP: the softmax warps have read S (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [3][2 tiles][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_blocks = (p.Tq + 255) / 256;
  const int split = blockIdx.x % p.nsplit;
  const int qb = (blockIdx.x / p.nsplit) % q_blocks;
  const int head = (blockIdx.x / (p.nsplit * q_blocks)) % p.H;
  const int img = blockIdx.x / (p.nsplit * q_blocks * p.H);
  const int q0 = qb * 256;
  int len0 = p.kv_len[0];
  if (p.kv_len0_per_sample) len0 = max(0, min(len0, p.kv_len0_per_sample[img]));
  if (p.mask_chunk) len0 = min(len0, causal_limit2(p, min(q0 + 256, p.Tq) - 1));
  const int n0 = (len0 + BKV - 1) / BKV;
  const int n1 = (p.kv_len[1] + BKV - 1) / BKV;
  // this CTA's share [jb, jb + nkv) of the n0 + n1 key tiles
  const int jb = static_cast<int>(static_cast<int64_t>(n0 + n1) * split / p.nsplit);
  const int nkv = static_cast<int>(static_cast<int64_t>(n0 + n1) * (split + 1) / p.nsplit) - jb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK[0]);
    tma_prefetch_desc(&p.tmV[0]);
    tma_prefetch_desc(&p.tmO);
    if (n1) {
      tma_prefetch_desc(&p.tmK[1]);
      tma_prefetch_desc(&p.tmV[1]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < S; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4 * WG);
        mbar_init(&pv_done[i], 1);
      }
      mbar_init(&s_free[0], 4);
      mbar_init(&s_free[1], 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if constexpr (ONES) {
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {   // one 16-byte chunk each: row r holds 1.0 (bf16) in logical column 0,
      const int i = threadIdx.x - 64, r = i >> 3, ch = i & 7;   // i.e. in chunk (0 ^ (r & 7)) of the 128B-swizzled row
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (ch == (r & 7)) u.x = 0x3f80u;
      *reinterpret_cast<uint4*>(sOnes + i * 16) = u;
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int dc = 0; dc < Cfg::DC; ++dc)
          tma_load_4d(sQ + t * Cfg::Q_TILE_BYTES + dc * 128 * 128, &p.tmQ, q_full, dc * 64, head, q0 + t * 128, img);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int seg = (jb + j) >= n0;
        const int row0 = (seg ? jb + j - n0 : jb + j) * BKV;
        uint8_t* sk = sKV + stage * 2 * Cfg::KV_TILE_BYTES;
        uint8_t* sv = sk + Cfg::KV_TILE_BYTES;
        if constexpr (HALF) {
          // K_j is free once the S products of tile j - S have run (early in the previous iteration), V_j only after
          // its P.V products: separate barriers let K_{j+1} arrive a full iteration before its first reader
          mbar_wait(&kv_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&kv_full[stage], Cfg::KV_TILE_BYTES);
          for (int dc = 0; dc < Cfg::DC; ++dc)
            tma_load_4d(sk + dc * BKV * 128, &p.tmK[seg], &kv_full[stage], dc * 64, head, row0, img);
          mbar_wait(&v_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&v_full[stage], Cfg::KV_TILE_BYTES);
          for (int dc = 0; dc < Cfg::DC; ++dc)
            tma_load_4d(sv + dc * BKV * 128, &p.tmV[seg], &v_full[stage], dc * 64, head, row0, img);
        } else {
          mbar_wait(&kv_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&kv_full[stage], 2 * Cfg::KV_TILE_BYTES);
          for (int dc = 0; dc < Cfg::DC; ++dc) {
            tma_load_4d(sk + dc * BKV * 128, &p.tmK[seg], &kv_full[stage], dc * 64, head, row0, img);
            tma_load_4d(sv + dc * BKV * 128, &p.tmV[seg], &kv_full[stage], dc * 64, head, row0, img);
          }
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp runs the loop (converged: uniform operands, no per-MMA broadcast loop); one elected lane issues
    if (nkv > 0) {
      const uint32_t elected = elect_one_sync();
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, OW, 0, 1);  // B (= V [+ the ones column]) is MN-major
      const uint32_t ones_addr = smem_u32(sOnes);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int t, int stage) {
        const uint32_t k_addr = smem_u32(sKV + stage * 2 * Cfg::KV_TILE_BYTES);
        const uint32_t qa = q_addr + t * Cfg::Q_TILE_BYTES;
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(Cfg::TMEM_S + t * 128);
#pragma unroll
        for (int dc = 0; dc < Cfg::DC; ++dc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_elect(elected, d_tmem, umma_desc_sw128(qa + dc * 128 * 128 + k * 32),
                         umma_desc_sw128(k_addr + dc * BKV * 128 + k * 32), idesc_qk, (dc | k) != 0 ? 1u : 0u);
        umma_commit_elect(elected, &s_full[2 * t]);
      };
      auto issue_pv = [&](int t, int stage, int j) {
        const uint32_t v_addr = smem_u32(sKV + stage * 2 * Cfg::KV_TILE_BYTES + Cfg::KV_TILE_BYTES);
        const uint32_t o_tmem = tmem_base + static_cast<uint32_t>(Cfg::TMEM_O + t * OW);
        const uint32_t p_tmem = tmem_base + static_cast<uint32_t>(SEP ? TMEM_P + t * 64 : Cfg::TMEM_S + t * 128);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // second 64-column chunk of B: the next d chunk of V (d = 128) | the ones slice, the same 16 rows for every kk
          const uint32_t lbo = ONES ? ones_addr - (v_addr + kk * 2048) : static_cast<uint32_t>(BKV * 128);
          umma_bf16_ts_elect(elected, o_tmem, p_tmem + kk * 8, umma_desc_sw128(v_addr + kk * 2048, lbo), idesc_pv,
                       (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit_elect(elected, &pv_done[2 * t]);
      };
      // half-tile pipeline: per key tile j and half h, O_t += P_th.V_j[h] then S_th <- Q_t.K_{j+1}[h]^T for both tiles
      constexpr uint32_t idesc_qk_h = umma_idesc_bf16(128, 64, 0, 0);
      auto issue_qk_h = [&](int t, int h, int stage) {
        const uint32_t k_addr = smem_u32(sKV + stage * 2 * Cfg::KV_TILE_BYTES) + h * (64 * 128);
        const uint32_t qa = q_addr + t * Cfg::Q_TILE_BYTES;
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(Cfg::TMEM_S + t * 128 + h * 64);
#pragma unroll
        for (int dc = 0; dc < Cfg::DC; ++dc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_elect(elected, d_tmem, umma_desc_sw128(qa + dc * 128 * 128 + k * 32),
                         umma_desc_sw128(k_addr + dc * BKV * 128 + k * 32), idesc_qk_h, (dc | k) != 0 ? 1u : 0u);
        umma_commit_elect(elected, &s_full[2 * t + h]);
      };
      auto issue_pv_h = [&](int t, int h, int stage, int j) {
        const uint32_t v_addr = smem_u32(sKV + stage * 2 * Cfg::KV_TILE_BYTES + Cfg::KV_TILE_BYTES);
        const uint32_t o_tmem = tmem_base + static_cast<uint32_t>(Cfg::TMEM_O + t * DP);
        const uint32_t p_tmem = tmem_base + static_cast<uint32_t>(Cfg::TMEM_S + t * 128 + h * 64);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ts_elect(elected, o_tmem, p_tmem + kk * 8, umma_desc_sw128(v_addr + (h * 4 + kk) * 2048, BKV * 128), idesc_pv,
                       (j | h | kk) != 0 ? 1u : 0u);
        umma_commit_elect(elected, &pv_done[2 * t + h]);
      };
      if constexpr (HALF) {
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_qk_h(0, 0, 0);
        issue_qk_h(1, 0, 0);
        issue_qk_h(0, 1, 0);
        issue_qk_h(1, 1, 0);
        umma_commit_elect(elected, &kv_empty[0]);  // K_0 has no further reader
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < nkv; ++j) {
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == S) {
            nstage = 0;
            nphase ^= 1u;
          }
          const bool more = j + 1 < nkv;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (h == 0 && t == 0) mbar_wait(&v_full[stage], phase);
              mbar_wait(&p_full[2 * t + h], j & 1);
              tc_fence_after();
              issue_pv_h(t, h, stage, j);
              if (h == 1 && t == 1) umma_commit_elect(elected, &v_empty[stage]);  // every reader of V_j has been issued
              if (more) {
                if (h == 0 && t == 0) {
                  mbar_wait(&kv_full[nstage], nphase);
                  tc_fence_after();
                }
                issue_qk_h(t, h, nstage);  // overwrites P_th: issued after the product that reads it
                if (h == 1 && t == 1) umma_commit_elect(elected, &kv_empty[nstage]);  // every reader of K_{j+1} has been issued
              }
            }
          }
          stage = nstage;
          phase = nphase;
        }
      } else {
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == S) {
          nstage = 0;
          nphase ^= 1u;
        }
        const bool more = j + 1 < nkv;
        if constexpr (SEP) {
          // S_{j+1} of both tiles as soon as their softmax warps have read S_j; the P.V products when P_j is published
          if (more) {
            mbar_wait(&kv_full[nstage], nphase);
            mbar_wait(&s_free[0], j & 1);
            tc_fence_after();
            issue_qk(0, nstage);
            mbar_wait(&s_free[1], j & 1);
            tc_fence_after();
            issue_qk(1, nstage);
          }
          mbar_wait(&p_full[0], j & 1);
          tc_fence_after();
          issue_pv(0, stage, j);
          mbar_wait(&p_full[2], j & 1);
          tc_fence_after();
          issue_pv(1, stage, j);
          umma_commit_elect(elected, &kv_empty[stage]);
          stage = nstage;
          phase = nphase;
          continue;
        }
        // query tile 0: O0 += P0.V_j, then S0 <- Q0.K_{j+1}^T (overwrites P0: issued after the product that reads it)
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
#ifdef FG_ATTN_TIMING
        const bool trc = blockIdx.x == gridDim.x / 2 && j >= 2 && j < 6;
        if (trc) g_attn_trace[0][j - 2][1] = clock64();
#endif
        issue_pv(0, stage, j);
        if (more) {
          mbar_wait(&kv_full[nstage], nphase);
          tc_fence_after();
          issue_qk(0, nstage);
        }
#ifdef FG_ATTN_TIMING
        if (trc) g_attn_trace[0][j - 2][2] = clock64();
#endif
        // query tile 1
        mbar_wait(&p_full[2], j & 1);
        tc_fence_after();
#ifdef FG_ATTN_TIMING
        if (trc) g_attn_trace[1][j - 2][1] = clock64();
#endif
        issue_pv(1, stage, j);
        umma_commit_elect(elected, &kv_empty[stage]);  // every reader of K_j / V_j has been issued before this commit
        if (more) issue_qk(1, nstage);
#ifdef FG_ATTN_TIMING
        if (trc) g_attn_trace[1][j - 2][2] = clock64();
#endif
        stage = nstage;
        phase = nphase;
      }
      }  // !HALF
    }
  } else if constexpr (WG == 2) {
    // ===================== softmax: two warps per lane quarter, each one column half of the S tile =====================
    constexpr int HK = BKV / 2;              // keys per warp and key tile
    constexpr int OC = DP / 2;               // O columns per warp (rescale, epilogue)
    const int sw = warp - 2;
    const int t = sw >> 3;                   // query tile
    const int hh = (sw >> 2) & 1;            // column half
    const int q = warp & 3;                  // TMEM lane quarter
    const int row = q * 32 + lane;
    const int tok = q0 + t * 128 + row;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_S + t * 128);
    const uint32_t o_addr = tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_O + t * DP + hh * OC);
    const uint32_t pair_bar = 1u + static_cast<uint32_t>(t * 4 + q);
    const int limit = causal_limit2(p, tok);
    const float sc = p.scale_log2;
    float m_run = -INFINITY;  // running max in scaled (log2) units: identical in both warps of a pair
    float l_run = 0.f;        // partial row sum over this warp's keys
    // Register budget: 18 warps are allocated as 20, i.e. 96 registers per thread -- a whole 64-key half row (the WG == 1
    // scheme scaled down) spills.  So the row is handled as two 32-key chunks in two passes: pass 1 takes the maximum of
    // both chunks and keeps only the second one; after the exchange pass 2 re-reads the first chunk from tensor memory
    // (nobody has written there yet: P of keys k lands on column k / 2, i.e. this warp's first chunk is only ever
    // overwritten by its own P, and the partner's P -- columns 32..63 when it owns keys 64..127 -- falls on the SECOND
    // chunk of the warp that owns keys 0..63, which that warp holds in registers since before the barrier).
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[2 * t], j & 1);
      tc_fence_after();
      const uint32_t sh_addr = s_addr + static_cast<uint32_t>(hh * HK);
      uint32_t sa[32], sb[32];
      tmem_ld_32x32(sh_addr, sa);
      tmem_ld_32x32(sh_addr + 32, sb);
      tmem_ld_wait();
      const int seg = (jb + j) >= n0;
      const int col0 = (seg ? jb + j - n0 : jb + j) * BKV + hh * HK;
      int kv_valid = (seg ? p.kv_len[1] : len0) - col0;
      if (p.mask_chunk) kv_valid = min(kv_valid, limit - col0);
      const bool ragged = kv_valid < HK;  // ragged end of a segment / causal limit inside this half (rare)
      if (ragged) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c >= kv_valid) sa[c] = 0xff800000u;  // -inf
          if (c + 32 >= kv_valid) sb[c] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        mx0 = fmax3(mx0, __uint_as_float(sa[c]), __uint_as_float(sa[c + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sa[c + 2]), __uint_as_float(sa[c + 3]));
        mx2 = fmax3(mx2, __uint_as_float(sa[c + 4]), __uint_as_float(sa[c + 5]));
        mx3 = fmax3(mx3, __uint_as_float(sa[c + 6]), __uint_as_float(sa[c + 7]));
      }
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        mx0 = fmax3(mx0, __uint_as_float(sb[c]), __uint_as_float(sb[c + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sb[c + 2]), __uint_as_float(sb[c + 3]));
        mx2 = fmax3(mx2, __uint_as_float(sb[c + 4]), __uint_as_float(sb[c + 5]));
        mx3 = fmax3(mx3, __uint_as_float(sb[c + 6]), __uint_as_float(sb[c + 7]));
      }
      float mx = fmax3(fmaxf(mx0, mx1), mx2, mx3);
      // row max of the whole tile: exchange with the partner warp.  The barrier also orders this warp's S loads before
      // the partner's P stores.  Slots alternate with j, so a write two tiles later cannot overtake the partner's read
      // (one barrier per tile in between).
      float* xs = xch + ((j & 1) * 2 + t) * 256;
      xs[hh * 128 + row] = mx;
      named_bar_sync(pair_bar, 64);
      tmem_ld_32x32(sh_addr, sa);   // pass 2: the first chunk again (in flight during the max bookkeeping below)
      mx = fmaxf(mx, xs[(hh ^ 1) * 128 + row]);
      mx *= sc;  // scale > 0: max commutes with the scaling
      if (j > 0) {
        const bool need = mx > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {   // same rows, same values: both warps of the pair take the same branch
          mbar_wait(&pv_done[2 * t], (j - 1) & 1);  // O stable until this tile's P is published
          tc_fence_after();
          const float m_new = need ? mx : m_run;
          const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
          tmem_ld_wait();
#pragma unroll 1
          for (int c = 0; c < OC; c += 16) {   // this warp's half of the O columns
            uint32_t o[16];
            tmem_ld_32x16(o_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(o_addr + c, o);
          }
          tmem_st_wait();
        }
      } else {
        m_run = mx;
      }
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(neg_m, neg_m);
      uint64_t lsum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
      tmem_ld_wait();
      if (ragged) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c >= kv_valid) sa[c] = 0xff800000u;
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t(&sv)[32] = ch ? sb : sa;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sc2, negm2);
          float e0, e1;
          if (((i >> 1) & 7) < POLY) {
            exp2_poly2(x2, e0, e1);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            e0 = exp2_mufu(x0), e1 = exp2_mufu(x1);
          }
          lsum2[(i >> 1) & 3] = fadd2(lsum2[(i >> 1) & 3], pack2(e0, e1));
          pk[i >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x16(s_addr + ((hh * HK + ch * 32) >> 1), pk);  // P of key k lives at column k / 2 of this tile's S
      }
      float lsum0, lsum1;
      unpack2(fadd2(fadd2(lsum2[0], lsum2[1]), fadd2(lsum2[2], lsum2[3])), lsum0, lsum1);
      l_run += lsum0 + lsum1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t]);
    }
    // ---- epilogue: row sum of both halves, then O / l -> bf16 -> swizzled smem (this tile's Q buffer) -> TMA store ----
    float* xl = xch + (4 + t) * 256;
    xl[hh * 128 + row] = l_run;
    named_bar_sync(pair_bar, 64);
    l_run += xl[(hh ^ 1) * 128 + row];
    if (nkv) {
      mbar_wait(&pv_done[2 * t], (nkv - 1) & 1);  // the last product issued for this tile
      tc_fence_after();
    }
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    uint8_t* so = sQ + t * Cfg::Q_TILE_BYTES;
#pragma unroll 1
    for (int c0 = 0; c0 < OC; c0 += 32) {
      const int c = hh * OC + c0;   // column within the head
      uint32_t o[32];
      if (nkv) {
        tmem_ld_32x32(o_addr + c0, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      uint8_t* rowp = so + (c >> 6) * (128 * 128) + row * 128;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        const int chunk = ((c & 63) + i) >> 3;
        *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) = u;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(9 + t, 256);
    if (hh == 0 && p.nsplit > 1 && tok < p.Tq)
      p.part_ml[(static_cast<int64_t>(img * p.H + head) * p.Tq + tok) * p.nsplit + split] = make_float2(m_run, l_run);
    if (hh == 0 && q == 0 && lane == 0 && q0 + t * 128 < p.Tq) {
      for (int dc = 0; dc < Cfg::DC; ++dc) {
        if (p.nsplit > 1)
          tma_store_4d(&p.tmO, so + dc * 128 * 128, dc * 64, split, q0 + t * 128, img * p.H + head);
        else if (dc * 64 < p.d)
          tma_store_4d(&p.tmO, so + dc * 128 * 128, dc * 64, head, q0 + t * 128, img);
      }
      tma_store_commit();
      tma_store_wait<0>();
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int t = (warp - 2) >> 2;  // query tile of this warpgroup
    const int q = warp & 3;         // TMEM lane quarter
    const int row = q * 32 + lane;
    const int tok = q0 + t * 128 + row;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_S + t * 128);
    const uint32_t o_addr = tmem_base + lane_base + static_cast<uint32_t>(Cfg::TMEM_O + t * OW);
    const uint32_t p_addr = tmem_base + lane_base + static_cast<uint32_t>(TMEM_P + t * 64);   // SEP only
    const int limit = causal_limit2(p, tok);
    const float sc = p.scale_log2;
    float m_run = -INFINITY;  // running max in scaled (log2) units
    float l_run = 0.f;
    if constexpr (HALF) {
      constexpr int HK = BKV / 2;  // keys per half
      for (int i = 0; i < 2 * nkv; ++i) {
        const int j = i >> 1, h = i & 1;
        mbar_wait(&s_full[2 * t + h], j & 1);
        tc_fence_after();
        const uint32_t sh_addr = s_addr + static_cast<uint32_t>(h * HK);
        uint32_t sv[HK];
#pragma unroll
        for (int c = 0; c < HK; c += 32) {
          uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c]);
          tmem_ld_32x32(sh_addr + c, chunk);
        }
        tmem_ld_wait();
        const int seg = (jb + j) >= n0;
        const int col0 = (seg ? jb + j - n0 : jb + j) * BKV + h * HK;
        int kv_valid = (seg ? p.kv_len[1] : len0) - col0;
        if (p.mask_chunk) kv_valid = min(kv_valid, limit - col0);
        if (kv_valid < HK) {  // ragged end of a segment / causal limit inside this half (rare)
#pragma unroll
          for (int c = 0; c < HK; ++c)
            if (c >= kv_valid) sv[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < HK; c += 8) {
          mx0 = fmax3(mx0, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
          mx2 = fmax3(mx2, __uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5]));
          mx3 = fmax3(mx3, __uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7]));
        }
        float mx = fmax3(fmaxf(mx0, mx1), mx2, mx3);
        mx *= sc;  // scale > 0: max commutes with the scaling
        if (i > 0) {
          const bool need = mx > m_run + 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            // O must be stable: wait for the product of the previous half (the last one issued for this tile)
            mbar_wait(&pv_done[2 * t + (h ^ 1)], (h ? j : j - 1) & 1);
            tc_fence_after();
            const float m_new = need ? mx : m_run;
            const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
            m_run = m_new;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < DP; c += 32) {
              uint32_t o[32];
              tmem_ld_32x32(o_addr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st_32x32(o_addr + c, o);
            }
            tmem_st_wait();
          }
        } else {
          m_run = mx;
        }
        const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
        const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(neg_m, neg_m);
        uint64_t lsum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < HK; c += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[c + k]), __uint_as_float(sv[c + k + 1])), sc2, negm2);
            float e0, e1;
            if (((k >> 1) & 7) < POLY) {
              exp2_poly2(x2, e0, e1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              e0 = exp2_mufu(x0), e1 = exp2_mufu(x1);
            }
            lsum2[(k >> 1) & 3] = fadd2(lsum2[(k >> 1) & 3], pack2(e0, e1));
            pk[k >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st_32x16(sh_addr + (c >> 1), pk);  // P_th: bf16 pairs over the first 32 columns of this half's S
        }
        float lsum0, lsum1;
        unpack2(fadd2(fadd2(lsum2[0], lsum2[1]), fadd2(lsum2[2], lsum2[3])), lsum0, lsum1);
        l_run += lsum0 + lsum1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + h]);
      }
    } else {
#ifdef FG_ATTN_TIMING
    long long tacc[5] = {0, 0, 0, 0, 0};
#endif
    for (int j = 0; j < nkv; ++j) {
      FG_TSTAMP(ts0);
      mbar_wait(&s_full[2 * t], j & 1);
      tc_fence_after();
      FG_TSTAMP(ts1);
#ifdef FG_ATTN_TIMING
      long long ts2 = 0, ts3 = 0;
#endif
      if constexpr (TP) {
        const int seg = (jb + j) >= n0;
        const int col0 = (seg ? jb + j - n0 : jb + j) * BKV;
        int kv_valid = (seg ? p.kv_len[1] : len0) - col0;
        if (p.mask_chunk) kv_valid = min(kv_valid, limit - col0);
        const bool ragged = kv_valid < BKV;   // ragged end of a segment / causal limit inside this tile (rare)
        // ---- pass 1: row max, 32 columns at a time, the next chunk's TMEM read in flight ----
        uint32_t ca[32], cb[32];
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        auto max_chunk = [&](uint32_t (&sv)[32], int c) {
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i >= kv_valid) sv[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
            mx2 = fmax3(mx2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
            mx3 = fmax3(mx3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
          }
        };
        tmem_ld_32x32(s_addr, ca);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 32, cb);
        max_chunk(ca, 0);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 64, ca);
        max_chunk(cb, 32);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 96, cb);
        max_chunk(ca, 64);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr, ca);        // pass 2, first chunk: in flight during the max bookkeeping
        max_chunk(cb, 96);
        FG_TSET(ts2);
        float mx = fmax3(fmaxf(mx0, mx1), mx2, mx3);
        mx *= sc;  // scale > 0: max commutes with the scaling
        if (j > 0) {
          const bool need = mx > m_run + 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(&pv_done[2 * t], (j - 1) & 1);  // O stable until this tile's P is published
            tc_fence_after();
            const float m_new = need ? mx : m_run;
            const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
            m_run = m_new;
            l_run *= alpha;
            tmem_ld_wait();
#pragma unroll 1
            for (int c = 0; c < OW; c += 16) {
              uint32_t o[16];
              tmem_ld_32x16(o_addr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x16(o_addr + c, o);
            }
            tmem_st_wait();
          }
        } else {
          m_run = mx;
        }
        FG_TSET(ts3);
        if constexpr (PP) {
          if (t == 1 && j == 0) named_bar_arrive(3 + q, 64);
          named_bar_sync(3 + 4 * t + q, 64);
        }
        const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
        const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(neg_m, neg_m);
        uint64_t lsum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
        // ---- pass 2: P = exp2(s * scale - m) as bf16 over the S columns (P of chunk c lands in columns already consumed) ----
        auto exp_chunk = [&](uint32_t (&sv)[32], int c) {
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i >= kv_valid) sv[i] = 0xff800000u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sc2, negm2);
            float e0, e1;
            if (((i >> 1) & 7) < POLY) {
              exp2_poly2(x2, e0, e1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              e0 = exp2_mufu(x0), e1 = exp2_mufu(x1);
            }
            if constexpr (!ONES) lsum2[(i >> 1) & 3] = fadd2(lsum2[(i >> 1) & 3], pack2(e0, e1));
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st_32x16(s_addr + (c >> 1), pk);
        };
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 32, cb);
        exp_chunk(ca, 0);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 64, ca);
        exp_chunk(cb, 32);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 96, cb);
        exp_chunk(ca, 64);
        tmem_ld_wait();
        exp_chunk(cb, 96);
        if constexpr (!ONES) {
          float lsum0, lsum1;
          unpack2(fadd2(fadd2(lsum2[0], lsum2[1]), fadd2(lsum2[2], lsum2[3])), lsum0, lsum1);
          l_run += lsum0 + lsum1;
        }
      } else {
      uint32_t sv[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) {
        uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c]);
        tmem_ld_32x32(s_addr + c, chunk);
      }
      tmem_ld_wait();
      if constexpr (SEP) {   // S is in registers: hand its columns back, the issuer computes S of the next key tile now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
      }
      FG_TSET(ts2);
      const int seg = (jb + j) >= n0;
      const int col0 = (seg ? jb + j - n0 : jb + j) * BKV;
      int kv_valid = (seg ? p.kv_len[1] : len0) - col0;
      if (p.mask_chunk) kv_valid = min(kv_valid, limit - col0);
      if (kv_valid < BKV) {  // ragged end of a segment / causal limit inside this tile (rare)
#pragma unroll
        for (int c = 0; c < BKV; ++c)
          if (c >= kv_valid) sv[c] = 0xff800000u;  // -inf
      }
      // four independent FMNMX3 chains (16 deep each instead of one of 64): only two softmax warps share a scheduler, so
      // a serial dependency chain shows as issue bubbles
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKV; c += 8) {
        mx0 = fmax3(mx0, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
        mx2 = fmax3(mx2, __uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5]));
        mx3 = fmax3(mx3, __uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7]));
      }
      float mx = fmax3(fmaxf(mx0, mx1), mx2, mx3);
      mx *= sc;  // scale > 0: max commutes with the scaling
      if (j > 0) {
        const bool need = mx > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&pv_done[2 * t], (j - 1) & 1);  // O stable until this tile's P is published
          tc_fence_after();
          const float m_new = need ? mx : m_run;
          const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
#pragma unroll 1
          for (int c = 0; c < DP; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(o_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(o_addr + c, o);
          }
          if constexpr (ONES) {   // the row sum lives in column DP of the accumulator
            uint32_t o[16];
            tmem_ld_32x16(o_addr + DP, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(o_addr + DP, o);
          }
          tmem_st_wait();
        }
      } else {
        m_run = mx;
      }
      FG_TSET(ts3);
      if constexpr (PP) {   // wait for the exp token of this scheduler's warp pair (tile 0 holds it first)
        if (t == 1 && j == 0) named_bar_arrive(3 + q, 64);
        named_bar_sync(3 + 4 * t + q, 64);
      }
      if constexpr (SEP) {
        if (j > 0) {   // P_{j-1} must have been consumed by its P.V product before P_j overwrites it (long done by now)
          mbar_wait(&pv_done[2 * t], (j - 1) & 1);
          tc_fence_after();
        }
      }
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(neg_m, neg_m);
      uint64_t lsum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};  // independent chains
      // P = exp2(s * scale - m) as bf16, written over the S columns: 32 keys -> 16 packed columns per store.
      // kPolyOf8 of every 8 pairs take the polynomial, the rest the MUFU unit (balances MUFU time against issue slots)
#pragma unroll
      for (int c = 0; c < BKV; c += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[c + i]), __uint_as_float(sv[c + i + 1])), sc2, negm2);
          if (BX > 0 && ((i >> 1) & 7) >= 8 - BX) {   // packed bf16 exponential: two keys per MUFU issue, no F2FP
            float x0, x1;
            unpack2(x2, x0, x1);
            pk[i >> 1] = exp2_bf16x2(x0, x1);
            continue;
          }
          float e0, e1;
          if (((i >> 1) & 7) < POLY) {
            exp2_poly2(x2, e0, e1);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            e0 = exp2_mufu(x0), e1 = exp2_mufu(x1);
          }
          if constexpr (!ONES) lsum2[(i >> 1) & 3] = fadd2(lsum2[(i >> 1) & 3], pack2(e0, e1));
          pk[i >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x16((SEP ? p_addr : s_addr) + (c >> 1), pk);
      }
      if constexpr (!ONES) {
        float lsum0, lsum1;
        unpack2(fadd2(fadd2(lsum2[0], lsum2[1]), fadd2(lsum2[2], lsum2[3])), lsum0, lsum1);
        l_run += lsum0 + lsum1;
      }
      }
      FG_TSTAMP(ts4);
      if constexpr (PP) {   // hand the token to the partner (which needs it for its tile j, or j + 1 if this is tile 1)
        if (t == 0 || j + 1 < nkv) named_bar_arrive(3 + 4 * (t ^ 1) + q, 64);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t]);
#ifdef FG_ATTN_TIMING
      const long long ts5 = clock64();
      tacc[0] += ts1 - ts0, tacc[1] += ts2 - ts1, tacc[2] += ts3 - ts2, tacc[3] += ts4 - ts3, tacc[4] += ts5 - ts4;
      if (blockIdx.x == gridDim.x / 2 && q == 0 && lane == 0) {
        if (j >= 2 && j < 6) g_attn_trace[t][j - 2][0] = ts5;
        if (j >= 3 && j < 7) g_attn_trace[t][j - 3][3] = ts1;
      }
#endif
    }
#ifdef FG_ATTN_TIMING
    if (blockIdx.x == gridDim.x / 2 && q == 0 && lane == 0) {
      for (int i = 0; i < 5; ++i) g_attn_timing[t][i] = tacc[i];
      g_attn_timing[t][5] = nkv;
    }
#endif
    }  // !HALF
    // ---- epilogue: O / l -> bf16 -> swizzled smem (this tile's Q buffer) -> TMA store ----
    if (nkv) {
      mbar_wait(&pv_done[2 * t + HALF], (nkv - 1) & 1);  // the last product issued for this tile
      tc_fence_after();
    }
    if constexpr (ONES) {
      if (nkv) {   // column DP of the accumulator: sum of the bf16 weights the numerator was built from
        uint32_t o[16];
        tmem_ld_32x16(o_addr + DP, o);
        tmem_ld_wait();
        l_run = __uint_as_float(o[0]);
      }
    }
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    uint8_t* so = sQ + t * Cfg::Q_TILE_BYTES;
#pragma unroll 1
    for (int c = 0; c < DP; c += 32) {
      uint32_t o[32];
      if (nkv) {
        tmem_ld_32x32(o_addr + c, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      uint8_t* rowp = so + (c >> 6) * (128 * 128) + row * 128;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        const int chunk = ((c & 63) + i) >> 3;
        *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) = u;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1 + t, 128);
    if (p.nsplit > 1 && tok < p.Tq)
      p.part_ml[(static_cast<int64_t>(img * p.H + head) * p.Tq + tok) * p.nsplit + split] = make_float2(m_run, l_run);
    if (q == 0 && lane == 0 && q0 + t * 128 < p.Tq) {
      for (int dc = 0; dc < Cfg::DC; ++dc) {
        if (p.nsplit > 1)
          tma_store_4d(&p.tmO, so + dc * 128 * 128, dc * 64, split, q0 + t * 128, img * p.H + head);
        else if (dc * 64 < p.d)
          tma_store_4d(&p.tmO, so + dc * 128 * 128, dc * 64, head, q0 + t * 128, img);
      }
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// out[b, tok, h, :] = sum_s w_s O_s / sum_s w_s,  w_s = l_s 2^(m_s - max m): one warp per (b, h, tok)
template <int DP>
__global__ void __launch_bounds__(256) attn2_merge_kernel(const __nv_bfloat16* __restrict__ part_o,
                                                          const float2* __restrict__ part_ml, int B, int H, int Tq,
                                                          int d, int nsplit, __nv_bfloat16* __restrict__ out,
                                                          int64_t out_ld, int64_t out_bs) {
  constexpr int EPL = DP / 32;  // elements per lane
  const int lane = threadIdx.x & 31;
  const int64_t rows = static_cast<int64_t>(B) * H * Tq;
  for (int64_t r = blockIdx.x * 8ll + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8ll) {
    const int tok = static_cast<int>(r % Tq);
    const int bh = static_cast<int>(r / Tq);
    const float2* ml = part_ml + r * nsplit;
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, ml[s].x);
    float acc[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
    float wsum = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float2 v = ml[s];
      if (v.y <= 0.f) continue;
      const float w = v.y * exp2f(v.x - m);
      wsum += w;
      const __nv_bfloat16* po = part_o + (r * nsplit + s) * DP + lane * EPL;
      if constexpr (EPL == 4) {
        const uint2 u = *reinterpret_cast<const uint2*>(po);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        acc[0] = fmaf(w, a.x, acc[0]), acc[1] = fmaf(w, a.y, acc[1]), acc[2] = fmaf(w, b.x, acc[2]), acc[3] = fmaf(w, b.y, acc[3]);
      } else {
        const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(po));
        acc[0] = fmaf(w, a.x, acc[0]), acc[1] = fmaf(w, a.y, acc[1]);
      }
    }
    const float inv = wsum > 0.f ? 1.0f / wsum : 0.f;
    const int c = lane * EPL;
    if (c < d) {
      __nv_bfloat16* o = out + static_cast<int64_t>(bh / H) * out_bs + static_cast<int64_t>(tok) * out_ld + (bh % H) * d + c;
      if constexpr (EPL == 4) {
        uint2 u;
        u.x = pack_bf16x2(acc[0] * inv, acc[1] * inv), u.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
        *reinterpret_cast<uint2*>(o) = u;
      } else {
        *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(acc[0] * inv, acc[1] * inv);
      }
    }
  }
}

// number of KV splits: minimise (waves of CTAs over the SMs) x (key tiles per CTA + fixed per-CTA cost, in key-tile
// units: Q load, pipeline fill, partial-result epilogue); one more tile-time for the merge pass when splitting
static int pick_nsplit(int ctas, int kv_tiles, int num_sms) {
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 8 && ns <= kv_tiles; ++ns) {
    const int64_t waves = (static_cast<int64_t>(ctas) * ns + num_sms - 1) / num_sms;
    const double cost = static_cast<double>(waves) * (static_cast<double>(kv_tiles) / ns + 5.0) + (ns > 1 ? 2.0 : 0.0);
    if (cost < best_cost * 0.97) {  // split only for a clear win
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

size_t flash_attn_v2_workspace(const fg_flash_attn_params* p) {
  const int dp = p->d <= 64 ? 64 : 128;
  const size_t rows = static_cast<size_t>(p->batch) * p->heads * p->tq;
  // up to 8 splits: bf16 O + (m, l); slack for the 256-byte alignment of (m, l) and the 128-byte alignment of O
  return rows * 8 * (static_cast<size_t>(dp) * 2 + 8) + 256 + 128;
}

template <int DP, int POLY, int HALF = 0, int WG = 1, int BX = 0, int PP = 0, int TP = 0, int SEP = 0>
static int launch_attention2(const fg_flash_attn_params* p, float scale, cudaStream_t stream) {
  using Cfg = Attn2Cfg<DP>;
  static PerDeviceFlag configured;
  const int cfg_dev = current_device();
  if (!configured.test(cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(attention2_kernel<DP, POLY, HALF, WG, BX, PP, TP, SEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("flash_attn: cudaFuncSetAttribute(%d B) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return FG_ERR_CUDA;
    }
    configured.set(cfg_dev);
  }
  Attn2KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = p->batch, kp.Tq = p->tq, kp.H = p->heads, kp.d = p->d;
  kp.kv_len[0] = p->kv_len[0], kp.kv_len[1] = p->kv_len[1];
  kp.mask_first = p->mask_chunk > 0 ? p->mask_first : 0;
  kp.mask_chunk = p->mask_chunk;
  kp.scale_log2 = 1.4426950408889634f * scale;
  kp.kv_len0_per_sample = p->kv_len0_per_sample;
  const uint32_t box[4] = {64, 1, 128, 1};
  const int q_blocks = (kp.Tq + 255) / 256;
  // split the key range over several CTAs when the (sample, head, query block) grid alone leaves SMs idle or ends in
  // a mostly empty wave; needs the caller's workspace, and is not combined with the causal mask (uneven key counts)
  int nsplit = 1;
  __nv_bfloat16* part_o = nullptr;
  if (p->workspace && static_cast<size_t>(p->workspace_bytes) >= flash_attn_v2_workspace(p) && p->mask_chunk == 0) {
    const int num_sms = device_sm_count(cfg_dev);
    static const int force = getenv("FG_ATTN_SPLIT") ? atoi(getenv("FG_ATTN_SPLIT")) : 0;
    const int tiles = (kp.kv_len[0] + kA2BKV - 1) / kA2BKV + (kp.kv_len[1] + kA2BKV - 1) / kA2BKV;
    nsplit = force > 0 ? (force < tiles ? force : tiles) : pick_nsplit(kp.B * kp.H * q_blocks, tiles, num_sms);
    if (nsplit > 8) nsplit = 8;
    if (nsplit > 1) {
      const size_t rows = static_cast<size_t>(kp.B) * kp.H * kp.Tq;
      kp.part_ml = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(p->workspace) + 255) & ~uintptr_t(255));
      part_o = reinterpret_cast<__nv_bfloat16*>(kp.part_ml + rows * nsplit);
      part_o = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(part_o) + 127) & ~uintptr_t(127));
    }
  }
  kp.nsplit = nsplit;
  {
    const uint64_t dims[4] = {(uint64_t)kp.d, (uint64_t)kp.H, (uint64_t)kp.Tq, (uint64_t)kp.B};
    const uint64_t strides[4] = {2, (uint64_t)kp.d * 2, (uint64_t)p->q_ld * 2, (uint64_t)p->q_bs * 2};
    int r = encode_tmap_bf16(&kp.tmQ, p->q, 4, dims, strides, box);
    if (r != FG_OK) return r;
    if (nsplit > 1) {
      const uint64_t pdims[4] = {(uint64_t)DP, (uint64_t)nsplit, (uint64_t)kp.Tq, (uint64_t)kp.B * kp.H};
      const uint64_t pstrides[4] = {2, (uint64_t)DP * 2, (uint64_t)DP * nsplit * 2, (uint64_t)DP * nsplit * kp.Tq * 2};
      r = encode_tmap_bf16(&kp.tmO, part_o, 4, pdims, pstrides, box);
    } else {
      const uint64_t ostrides[4] = {2, (uint64_t)kp.d * 2, (uint64_t)p->out_ld * 2, (uint64_t)p->out_bs * 2};
      r = encode_tmap_bf16(&kp.tmO, p->out, 4, dims, ostrides, box);
    }
    if (r != FG_OK) return r;
  }
  for (int s = 0; s < 2; ++s) {
    const int src = kp.kv_len[s] > 0 ? s : 0;
    const uint64_t dims[4] = {(uint64_t)kp.d, (uint64_t)kp.H, (uint64_t)kp.kv_len[src], (uint64_t)kp.B};
    const uint64_t strides[4] = {2, (uint64_t)kp.d * 2, (uint64_t)p->kv_ld[src] * 2, (uint64_t)p->kv_bs[src] * 2};
    int r = encode_tmap_bf16(&kp.tmK[s], p->k[src], 4, dims, strides, box);
    if (r != FG_OK) return r;
    r = encode_tmap_bf16(&kp.tmV[s], p->v[src], 4, dims, strides, box);
    if (r != FG_OK) return r;
  }
  const int grid = kp.B * kp.H * q_blocks * nsplit;
  attention2_kernel<DP, POLY, HALF, WG, BX, PP, TP, SEP><<<grid, 64 + 256 * WG, Cfg::SMEM_BYTES, stream>>>(kp);
  FG_CHECK_LAUNCH("flash_attn");
  if (nsplit > 1) {
    const int64_t rows = static_cast<int64_t>(kp.B) * kp.H * kp.Tq;
    const int mgrid = static_cast<int>(rows / 8 + 1 < 148 * 8 ? rows / 8 + 1 : 148 * 8);
    attn2_merge_kernel<DP><<<mgrid, 256, 0, stream>>>(part_o, kp.part_ml, kp.B, kp.H, kp.Tq, kp.d, nsplit,
                                                      static_cast<__nv_bfloat16*>(p->out), p->out_ld, p->out_bs);
    FG_CHECK_LAUNCH("flash_attn_merge");
  }
  return FG_OK;
}

// called by fg_flash_attn after argument validation
int flash_attn_v2(const fg_flash_attn_params* p, float scale, cudaStream_t stream) {
  // share of the exponentials taken by the FMA-pipe polynomial (pairs out of 8; FG_ATTN_POLY=0|2|3|4 overrides).  Since
  // the MMA issuer runs converged (profiles/attn_r02.md) the d = 128 kernel is bound by the tensor pipe and the fewest
  // instructions win: all exponentials on the MUFU unit (+3..4 % over 2/8); at d = 64 the softmax warps are the bound
  // and 2/8 is still the best split (535 vs 522 TFLOP/s at 0/8, 507 at 3/8 on the EDM 32x32 shape).
  static const int poly_env = getenv("FG_ATTN_POLY") ? atoi(getenv("FG_ATTN_POLY")) : -1;
  const int poly = poly_env >= 0 ? poly_env : (p->d <= 64 ? 2 : 0);
  // half-tile pipeline: opt-in experiment (FG_ATTN_HALF=1).  Measured on B200 it is SLOWER at d = 128 (816 vs 1017 TFLOP/s
  // on the Wan chunk-6 shape: the N = 64 S products run at half the tensor-pipe efficiency) and equal at d = 64, where
  // the softmax warps are bound by dependent-issue latency, not by the wait for S (profiles/attn_r01.md).
  static const bool half = getenv("FG_ATTN_HALF") && getenv("FG_ATTN_HALF")[0] == '1';
  if (half && poly == 2) {
    if (p->d <= 64) return launch_attention2<64, 2, 1>(p, scale, stream);
    return launch_attention2<128, 2, 1>(p, scale, stream);
  }
  // softmax warps per lane quarter.  WG = 2 (FG_ATTN_WG=2; 16 softmax warps, column-split rows) is an experiment that is
  // correct (same tests) but measured SLOWER on B200: 970 vs 1008 TFLOP/s on the Wan chunk-6 shape, 431 vs 494 at d = 64
  // (profiles/attn_r02.md) -- at 96 registers per thread the second TMEM read of half of S, the per-tile pair barrier and
  // the halved per-thread instruction-level parallelism cost more than the extra warps per scheduler hide.
  static const int wg = getenv("FG_ATTN_WG") ? atoi(getenv("FG_ATTN_WG")) : 1;
  if (wg == 2 && (poly == 2 || poly == 0 || poly == 4)) {
    if (p->d <= 64) {
      if (poly == 0) return launch_attention2<64, 0, 0, 2>(p, scale, stream);
      if (poly == 4) return launch_attention2<64, 4, 0, 2>(p, scale, stream);
      return launch_attention2<64, 2, 0, 2>(p, scale, stream);
    }
    if (poly == 0) return launch_attention2<128, 0, 0, 2>(p, scale, stream);
    if (poly == 4) return launch_attention2<128, 4, 0, 2>(p, scale, stream);
    return launch_attention2<128, 2, 0, 2>(p, scale, stream);
  }
  // d <= 64: row sum through a ones column of V and packed bf16 exponentials (see attention2_kernel).  FG_ATTN_BX = -1
  // (ones column alone) | 0 (off) | 2 | 4 | 8 eighths of the exponentials in the packed form
  static const int bx_env = getenv("FG_ATTN_BX") ? atoi(getenv("FG_ATTN_BX")) : 0;
  // d <= 64: P in its own tensor-memory columns, S of the next key tile computed under the softmax (see attention2_kernel)
  // Default at d <= 64 since the end of round 2 (FG_ATTN_SEP=0 restores the aliased layout): 516 -> 551 TFLOP/s on the EDM
  // 32x32 shape with every exponential on the MUFU unit (the polynomial share no longer pays once S never has to be waited
  // for: 519 with 2 of 8, 438 with 4 of 8 -- profiles/attn_separate_p_r02.txt).
  static const int sep_env = getenv("FG_ATTN_SEP") ? atoi(getenv("FG_ATTN_SEP")) : 1;
  static const bool other_experiment = getenv("FG_ATTN_TP") || getenv("FG_ATTN_PP") || getenv("FG_ATTN_BX") ||
                                       getenv("FG_ATTN_WG") || getenv("FG_ATTN_HALF");
  if (sep_env && p->d <= 64 && !other_experiment) {
    if (poly_env == 2) return launch_attention2<64, 2, 0, 1, 0, 0, 0, 1>(p, scale, stream);
    if (poly_env == 4) return launch_attention2<64, 4, 0, 1, 0, 0, 0, 1>(p, scale, stream);
    return launch_attention2<64, 0, 0, 1, 0, 0, 0, 1>(p, scale, stream);
  }
  static const int tp_env = getenv("FG_ATTN_TP") ? atoi(getenv("FG_ATTN_TP")) : 0;
  if (tp_env) {
    if (p->d <= 64) {
      if (poly == 0) return launch_attention2<64, 0, 0, 1, 0, 0, 1>(p, scale, stream);
      if (poly == 4) return launch_attention2<64, 4, 0, 1, 0, 0, 1>(p, scale, stream);
      return launch_attention2<64, 2, 0, 1, 0, 0, 1>(p, scale, stream);
    }
    if (poly == 2) return launch_attention2<128, 2, 0, 1, 0, 0, 1>(p, scale, stream);
    return launch_attention2<128, 0, 0, 1, 0, 0, 1>(p, scale, stream);
  }
  static const int pp_env = getenv("FG_ATTN_PP") ? atoi(getenv("FG_ATTN_PP")) : 0;
  if (p->d <= 64 && pp_env) {
    if (bx_env == -1) return launch_attention2<64, 2, 0, 1, -1, 1>(p, scale, stream);
    if (poly == 0) return launch_attention2<64, 0, 0, 1, 0, 1>(p, scale, stream);
    if (poly == 4) return launch_attention2<64, 4, 0, 1, 0, 1>(p, scale, stream);
    return launch_attention2<64, 2, 0, 1, 0, 1>(p, scale, stream);
  }
  if (p->d > 64 && pp_env) {
    if (poly == 2) return launch_attention2<128, 2, 0, 1, 0, 1>(p, scale, stream);
    return launch_attention2<128, 0, 0, 1, 0, 1>(p, scale, stream);
  }
  if (p->d <= 64 && bx_env != 0) {
    if (bx_env == -1) return launch_attention2<64, 2, 0, 1, -1>(p, scale, stream);
    if (bx_env == 2) return launch_attention2<64, 2, 0, 1, 2>(p, scale, stream);
    if (bx_env == 4) return launch_attention2<64, 2, 0, 1, 4>(p, scale, stream);
    if (bx_env == 40) return launch_attention2<64, 0, 0, 1, 4>(p, scale, stream);
    if (bx_env == 8) return launch_attention2<64, 0, 0, 1, 8>(p, scale, stream);
  }
  if (p->d <= 64) {
    if (poly == 0) return launch_attention2<64, 0>(p, scale, stream);
    if (poly == 3) return launch_attention2<64, 3>(p, scale, stream);
    if (poly == 4) return launch_attention2<64, 4>(p, scale, stream);
    return launch_attention2<64, 2>(p, scale, stream);
  }
  if (poly == 0) return launch_attention2<128, 0>(p, scale, stream);
  if (poly == 3) return launch_attention2<128, 3>(p, scale, stream);
  if (poly == 4) return launch_attention2<128, 4>(p, scale, stream);
  return launch_attention2<128, 2>(p, scale, stream);
}

}  // namespace fg

#ifdef FG_ATTN_TIMING
extern "C" int fg_debug_attn_timing(unsigned long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fg::g_attn_timing, sizeof(fg::g_attn_timing)));
}
extern "C" int fg_debug_attn_trace(long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fg::g_attn_trace, sizeof(fg::g_attn_trace)));
}
#endif
