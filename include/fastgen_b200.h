/*
 * fastgen_b200 -- C ABI of the B200 (sm_100a) kernels behind the FastGen few-step student sampler.
 *
 * The reference (linnanwang/FastGen) is pure Python/PyTorch and has no FFI of its own; the "operator
 * API" of its hot path is two Python contracts: the network forward
 * `FastGenNetwork.forward(x_t, t, condition, ..., fwd_pred_type)` (fastgen/networks/network.py:156-208)
 * and the sampler `FastGenModel.generator_fn / _student_sample_loop` (fastgen/methods/model.py:315-420).
 * Every entry point below replaces the PyTorch/ATen/cuDNN call(s) the reference makes at the cited lines;
 * `fastgen_b200/` (Python) re-assembles them behind those two contracts and binds this file with ctypes.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless stated otherwise;
 *   - activations are channels-last: a "pixel row" is `C` contiguous elements, consecutive pixels of a
 *     [N,H,W,C] tensor are `ld` elements apart (ld >= C lets a tensor live inside a wider concat buffer);
 *   - every call is asynchronous on `stream` (a cudaStream_t), allocates nothing and is CUDA-graph capturable; the only
 *     process-wide state is a cached driver entry point, the one-time shared-memory attribute of each kernel, and the
 *     fg_sampler_vec switch;
 *   - return value: FG_OK (0) or a negative fg_status; fg_last_error() gives a thread-local message.
 */
#ifndef FASTGEN_B200_H_
#define FASTGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fg_stream_t; /* cudaStream_t */

typedef enum { FG_OK = 0, FG_ERR_INVALID = -1, FG_ERR_CUDA = -2, FG_ERR_UNSUPPORTED = -3 } fg_status;
typedef enum { FG_F32 = 0, FG_BF16 = 1, FG_F64 = 2 } fg_dtype;
typedef enum { FG_RES_NONE = 0, FG_RES_SAME = 1, FG_RES_UP2 = 2, FG_RES_DOWN2 = 3 } fg_resample;
/* FG_ACT_GEGLU: the accumulator columns are (hidden_j, gate_j) pairs (the caller interleaves the weight rows) and the
 * output has cout / 2 columns, out[p, j] = v[p, 2j] * gelu_erf(v[p, 2j+1]) -- diffusers GEGLU of the SDXL feed-forward. */
typedef enum { FG_ACT_NONE = 0, FG_ACT_GELU_TANH = 1, FG_ACT_SILU = 2, FG_ACT_GEGLU = 3 } fg_activation;

const char* fg_last_error(void);
int fg_version(void);
/* number of kernels this library has launched on the calling thread since the last reset (bench bookkeeping) */
int64_t fg_launch_count(void);
void fg_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (TMA-staged, TMEM accumulators).
 * Replaces F.conv2d 3x3/1x1 (fastgen/networks/EDM/network.py:109-123), Linear (:47-51) and the
 * bias / embedding / skip-add / skip_scale elementwise tail of UNetBlock.forward (:276-298).
 *
 *   v[p, co]   = sum_s sum_tap sum_ci  A_s[p + tap, ci] * W_s[co, tap, ci]  + bias[co] + rowvec[n(p), co]
 *   out[p, co] = out_scale * ( gate[p / gate_rows, co] * act(v[p, co]) + resample(res)[p, co] )
 *
 * Each K segment s has its own bf16 activation A_s (same N,H,W; zero padding 1 for 3x3) and bf16 weight
 * W_s laid out [cout][taps][cin]; several segments accumulate into one TMEM tile, which is how a block's
 * conv1 + 1x1 skip projection, and the bf16x3 split-precision ("fp32 path") products, become ONE launch.
 * H=W=1 turns it into a plain row-major GEMM  out[n, co] = A[n, :] . W[co, :].
 * ---------------------------------------------------------------------------------------------------- */
#define FG_MAX_SEG 6

typedef struct {
  const void* a;  /* bf16 [N,H,W,cin], pixel stride a_ld elements (a_ld % 8 == 0, 16-byte aligned base) */
  int64_t a_ld;
  const void* w;  /* bf16 [cout][taps][cin] contiguous */
  int32_t cin;    /* % 8 == 0 */
  int32_t taps;   /* 1 (1x1 / linear) or 9 (3x3, pad 1, tap = kh*3+kw) */
} fg_conv_seg;

typedef struct {
  int32_t n, h, w;      /* output == input spatial size (stride-1 conv); w, h powers of two or multiples of 128 */
  int32_t cout;
  int32_t nseg;
  fg_conv_seg seg[FG_MAX_SEG];
  const float* bias;    /* [cout] fp32 or NULL */
  const float* rowvec;  /* [n][rowvec_ld] fp32 per-sample additive vector or NULL (UNetBlock `x.add_(params)`) */
  int64_t rowvec_ld;
  int32_t act;          /* fg_activation applied to v: FFN GELU(tanh) of the Wan/DiT blocks (Wan/network.py:147-149) */
  const float* gate;    /* [ceil(pixels / gate_rows)][gate_ld] fp32 multiplicative AdaLN gate or NULL: per sample  */
  int64_t gate_ld;      /* (gate_rows = tokens per sample) or per frame (gate_rows = tokens per frame,            */
  int32_t gate_rows;    /* Wan/network_causal.py:467-548); pixels are counted row-major over [n][h][w]            */
  const void* res;      /* residual input or NULL; FG_RES_UP2 reads a (h/2,w/2) tensor (nearest), FG_RES_DOWN2 */
  int32_t res_dtype;    /* averages 2x2 of a (2h,2w) tensor: the kernel-0 skip of up/down blocks (:253-262)     */
  int32_t res_mode;
  int64_t res_ld;
  float out_scale;      /* skip_scale (:287,298) */
  void* out;            /* [N,H,W,cout] fp32 or bf16, pixel stride out_ld */
  int32_t out_dtype;
  int64_t out_ld;
  double* stats;        /* optional: accumulate GroupNorm statistics of `out` (fp32 only) into stats[n][g][{sum,sumsq}], */
  int32_t stats_groups; /* g = (stats_ch0 + co) / stats_cg -- the fg_gn_stats pass of the consumer fused into the      */
  int32_t stats_cg;     /* producer (the caller zeroes the buffer, e.g. fg_zero, before the launch)                     */
  int32_t stats_ch0;
  double* stats2;        /* optional second statistics target [n][stats2_groups][2]: a tensor that is also a skip connection */
  int32_t stats2_groups; /* feeds the decoder GroupNorm that reads it at channel offset stats2_ch0 of a concatenation       */
  int32_t stats2_cg;     /* (torch.cat([x, skip]), EDM/network.py:559-561); with stats_ch0 on the producer of x, that norm  */
  int32_t stats2_ch0;    /* needs no fg_gn_stats pass of its own                                                            */
} fg_conv_params;

int fg_conv_gemm(const fg_conv_params* p, fg_stream_t stream);

/* The same operator split into "validate + choose the tile + encode the TMA descriptors" and "launch", for callers that
 * replay a fixed launch list (every network forward here is one: the reference re-dispatches F.conv2d / F.linear per call,
 * fastgen/networks/EDM/network.py:93-126): fg_conv_gemm re-encodes up to 2 nseg + 2 CUtensorMaps on EVERY call, the
 * prepared form does it once per plan entry.  `blob` is caller-owned host memory, 64-byte aligned, of at least
 * fg_conv_gemm_prepared_size() bytes; it stays valid for as long as the tensors named in `p` keep their addresses, on
 * the device that was current in fg_conv_gemm_prepare.  No allocation, no hidden state. */
int64_t fg_conv_gemm_prepared_size(void);
int fg_conv_gemm_prepare(const fg_conv_params* p, void* blob, int64_t blob_bytes);
int fg_conv_gemm_run(const void* blob, fg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * GroupNorm (+ per-sample scale/shift, + SiLU, + 2x resample) -- replaces torch group_norm + silu +
 * addcmul + the depthwise resample convs (fastgen/networks/EDM/network.py:141-149, 114-121, 276-283).
 * ---------------------------------------------------------------------------------------------------- */
/* stats[n][g] = {sum, sum of squares} (fp64) over the hw * (c/groups) elements of group g of sample n. */
int fg_gn_stats(const void* x, int32_t x_dtype, int64_t x_ld, int32_t n, int32_t hw, int32_t c, int32_t groups,
                double* stats, fg_stream_t stream);

typedef struct {
  const void* x;        /* [N,H,W,C] fp32 or bf16 */
  int32_t x_dtype;
  int64_t x_ld;
  int32_t n, h, w, c, groups;
  const double* stats;  /* from fg_gn_stats */
  float eps;
  const float* gamma;   /* [c] */
  const float* beta;    /* [c] */
  const float* scale;   /* [n][mod_ld] or NULL: y = shift + gn(x) * (1 + scale)  (adaptive_scale, :279-281) */
  const float* shift;   /* [n][mod_ld] or NULL */
  int64_t mod_ld;
  int32_t silu;         /* apply SiLU after the affine */
  int32_t resample;     /* FG_RES_NONE / FG_RES_UP2 (nearest, output 2h x 2w) / FG_RES_DOWN2 (2x2 mean, h/2 x w/2) */
  void* y;              /* bf16 normalised output (the next conv's A operand), pixel stride y_ld */
  int64_t y_ld;
  void* y_lo;           /* optional bf16 low part: y + y_lo carries ~16 mantissa bits (fp32 path) or NULL */
  void* raw;            /* optional bf16 copy of x itself under the same resample (skip 1x1 operand) or NULL */
  int64_t raw_ld;
  void* raw_lo;         /* optional low part of raw or NULL */
  void* raw_f32;        /* FG_RES_DOWN2 only, optional: the 2x2 mean of x itself in fp32 -- the resampled identity skip of a   */
  int64_t raw_f32_ld;   /* down block (EDM/network.py:114-121, 286-288), which the conv behind then adds as FG_RES_SAME     */
} fg_gn_apply_params;

int fg_gn_apply(const fg_gn_apply_params* p, fg_stream_t stream);

/* y = bf16(x) for a channels-last [n,h,w,c] tensor (fp32 or bf16), resample = FG_RES_NONE, or FG_RES_UP2 for the
 * nearest 2x upsampling of diffusers Upsample2D (F.interpolate(scale_factor=2, mode="nearest"); output [n,2h,2w,c]):
 * the raw conv operands of the SDXL UNet's down/upsamplers (UNet2DConditionModel, wrapped by the reference at
 * fastgen/networks/SDXL/network.py:210-230 and driven by fastgen/networks/SD15/network.py:152-324). */
int fg_cast_resample(const void* x, int32_t x_dtype, int64_t x_ld, int32_t n, int32_t h, int32_t w, int32_t c,
                     int32_t resample, void* y, int64_t y_ld, fg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * UNet self-attention (softmax(q.k/sqrt(d)) v per head, fp32 softmax) -- replaces AttentionOp + the two
 * einsums (fastgen/networks/EDM/network.py:160-168, 290-296).  qkv is the 1x1-conv output in channels-last
 * form [N, T, 3*heads*d] with channel order (q|k|v, head, d) (the wrapper permutes the conv's rows once).
 * ---------------------------------------------------------------------------------------------------- */
int fg_attention(const void* qkv /* bf16 */, int64_t qkv_ld, int32_t n, int32_t tokens, int32_t heads, int32_t d,
                 void* out /* bf16 [N,T,heads*d] */, int64_t out_ld, fg_stream_t stream);

/* The same attention evaluated entirely in fp32 on the CUDA cores (no bf16 rounding of q, k, v, the softmax weights or
 * the output), for the split-precision "fp32 path" of the EDM networks: the reference's AttentionOp is fp32 whatever the
 * activation dtype (fastgen/networks/EDM/network.py:160-168).  qkv is fp32 [N, T, 3*heads*d] (q|k|v, head, d), rows
 * 16-byte aligned; the output is the bf16 pair out_hi + out_lo (~16 mantissa bits; out_lo may be NULL), the operand
 * format of a three-segment fg_conv_gemm projection.  d in {64, 128, 192, 256}. */
int fg_attention_f32(const float* qkv, int64_t qkv_ld, int32_t n, int32_t tokens, int32_t heads, int32_t d,
                     void* out_hi /* bf16 [N,T,heads*d] */, void* out_lo, int64_t out_ld, fg_stream_t stream);

/* General flash attention (bf16 operands, fp32 softmax): out[b, t, h, :] = softmax(q.k^T * scale) v.
 * q / k / v are separate strided views [batch][tokens][heads*d] (row stride *_ld, batch stride *_bs, in elements);
 * the keys are the concatenation of up to two segments -- KV-cache prefix then current chunk -- which replaces the
 * per-layer torch.cat of fastgen/networks/Wan/network_causal.py:372-381; tq != key length gives cross-attention
 * (:343-358).  mask_chunk > 0 applies the block-causal mask of :131-196: a query of chunk c (first chunk
 * mask_first tokens, later chunks mask_chunk tokens) sees the keys up to the end of chunk c. */
typedef struct {
  const void* q;        /* bf16 */
  int64_t q_ld, q_bs;
  const void* k[2];     /* segment 0 must be non-empty; segment 1 optional (kv_len[1] == 0) */
  const void* v[2];
  int64_t kv_ld[2], kv_bs[2];
  int32_t kv_len[2];
  int32_t batch, heads, d, tq;
  float scale;          /* <= 0: 1/sqrt(d) */
  int32_t mask_first, mask_chunk;
  void* out;            /* bf16 [batch][tq][heads*d] */
  int64_t out_ld, out_bs;
  void* workspace;      /* optional scratch of fg_flash_attn_workspace_size() bytes: lets the library split the key  */
  int64_t workspace_bytes; /* range over several CTAs (partial results + merge) when the grid would leave SMs idle */
  const int32_t* kv_len0_per_sample; /* optional device array [batch]: sample b attends to the first                    */
                        /* min(kv_len[0], kv_len0_per_sample[b]) keys of segment 0 -- a prefix `encoder_attention_mask`   */
                        /* (padding after the prompt's tokens) of the SD 1.5 UNet, fastgen/networks/SD15/network.py:142-145: */
                        /* the reference adds -10000 to the masked logits, whose softmax weight underflows to exactly 0  */
} fg_flash_attn_params;

int fg_flash_attn(const fg_flash_attn_params* p, fg_stream_t stream);
/* bytes of scratch the split-KV path may use for this problem (0 if it never splits it) */
int64_t fg_flash_attn_workspace_size(const fg_flash_attn_params* p);

/* ------------------------------------------------------------------------------------------------------
 * Row kernels of the transformer blocks (Wan / DiT), all HBM-bound, bf16 rows in and out, fp32 statistics.
 * ---------------------------------------------------------------------------------------------------- */
/* y = LayerNorm(x) * (1 + scale[g]) + shift[g], g = row / rows_per_group  (AdaLN modulate, per sample or per frame:
 * fastgen/networks/Wan/network.py:139,150, network_causal.py:503-507,538-542), or LayerNorm(x) * gamma + beta
 * (norm2, Wan/network.py:144); LayerNorm in fp32 without rounding in between (diffusers FP32LayerNorm). */
typedef struct {
  const void* x;        /* [rows][x_ld], bf16 or fp32 (x_dtype) */
  int64_t x_ld;
  int32_t rows, d;
  float eps;
  const float* gamma;   /* [d] or NULL */
  const float* beta;
  const float* scale;   /* [groups][mod_ld] fp32 or NULL */
  const float* shift;
  int64_t mod_ld;
  int32_t rows_per_group;
  void* y;              /* bf16 [rows][y_ld] */
  int64_t y_ld;
  int32_t x_dtype;      /* FG_BF16 (Wan / DiT token rows) or FG_F32 (fp32 token stream of the SDXL transformer blocks) */
} fg_ln_mod_params;

int fg_ln_modulate(const fg_ln_mod_params* p, fg_stream_t stream);

/* In place: q <- RoPE(RMSNorm(q) * wq), k <- RoPE(RMSNorm(k) * wk); the RMS is taken over all heads*d channels of a
 * row (qk_norm = "rms_norm_across_heads"), the rotary embedding rotates interleaved pairs (x[2i], x[2i+1]) with the
 * 3-axis tables of WanRotaryPosEmbed: pair i of a head uses the temporal row (frame0 + f) for i < nt, the height row
 * h for i < nt + nh and the width row w otherwise, token t = (f * pph + h) * ppw + w
 * (network_causal.py:79-128, 262-291).  rope_cos == NULL: norm only (cross-attention).  With k_cache / v_cache the
 * roped k and the raw v are also written to row (cache_row0 + t) of the caches (store_kv, :365-368). */
typedef struct {
  void* q;              /* bf16 [rows][q_ld] or NULL */
  void* k;              /* bf16 [rows][k_ld] or NULL */
  const void* v;        /* bf16 [rows][v_ld], only read for the cache append */
  int64_t q_ld, k_ld, v_ld;
  int32_t rows, heads, d;
  float eps;
  const void* wq;       /* bf16 [heads*d] */
  const void* wk;
  const float* rope_cos; /* fp32 [rope_max_pos][d/2] */
  const float* rope_sin;
  int32_t rope_max_pos, nt, nh;
  int32_t rows_per_batch, ppf, pph, ppw, frame0;
  int32_t frame_rows;   /* rows per frame in this buffer (0: pph*ppw) and in-frame index of the first one: a        */
  int32_t tok0;         /* sequence-parallel rank holds tokens [tok0, tok0 + frame_rows) of every frame             */
  void* k_cache;        /* bf16 [batch][cache_bs], row stride cache_ld, or NULL */
  void* v_cache;
  int64_t cache_ld, cache_bs;
  int32_t cache_row0;
} fg_qk_norm_rope_params;

int fg_qk_norm_rope(const fg_qk_norm_rope_params* p, fg_stream_t stream);

/* latent [B,C,F,H,W] (fp32 | bf16) -> bf16 token rows [B*F*(H/ph)*(W/pw)][a_ld], column (c*ph + i)*pw + j: the
 * operand of the Conv3d patch embedding written as a GEMM (Wan/network.py:300-301). */
int fg_patchify(const void* x, int32_t x_dtype, int32_t b, int32_t c, int32_t f, int32_t h, int32_t w, int32_t ph,
                int32_t pw, void* a, int64_t a_ld, fg_stream_t stream);
/* token rows [..][y_ld] with column ((i*pw + j)*C + c) -> latent [B,C,F,H,W] (Wan/network.py:262-266). */
int fg_unpatchify(const void* y, int64_t y_ld, int32_t b, int32_t c, int32_t f, int32_t h, int32_t w, int32_t ph,
                  int32_t pw, void* out, int32_t out_dtype, fg_stream_t stream);
/* 4-D strided copy of items of item_bytes (a multiple of 16) bytes; strides in BYTES (multiples of 16): the pack /
 * unpack permutations around the Ulysses all-to-all (tokens <-> heads) and the head-sharded KV-cache append. */
int fg_strided_copy(const void* src, void* dst, const int32_t* n, const int64_t* src_strides,
                    const int64_t* dst_strides, int32_t item_bytes, fg_stream_t stream);
/* ------------------------------------------------------------------------------------------------------
 * Sequence-parallel (Ulysses) exchange over NVLink peer memory: pack + all-to-all as one kernel, no NCCL.
 * Replaces the all_to_all_single pair (plus its pack / unpack copies) around the self-attention of the Wan blocks under
 * sequence parallelism (fastgen/networks/Wan/network.py:106-153, network_causal.py:199-464).
 * fg_peer_put copies, for every destination q, the 4-D item set at src + q * src_peer_stride into dst[q] + dst_offset
 * (dst[q] / flags[q] are the device-mapped addresses of rank q's buffer / flag array: CUDA IPC or symmetric memory; the
 * own rank is a destination like any other), then writes its put epoch to flags[q][rank] of every rank.  fg_peer_wait
 * spins (on the device, stream-ordered) until all `peers` local flags reached the local wait epoch.  `state` is 4
 * zero-initialised uint64 of local device memory per channel {put epoch, CTA counter, wait epoch, timeout flag}; all
 * ranks issue the same put / wait sequence per channel, which keeps the epochs in lockstep without host involvement.
 * ---------------------------------------------------------------------------------------------------- */
#define FG_MAX_PEERS 8

typedef struct {
  const void* src;
  void* dst[FG_MAX_PEERS];
  int32_t peers, rank;
  int32_t n[4];             /* extents of the item set of ONE destination */
  int64_t src_strides[4];   /* bytes, multiples of 16 */
  int64_t dst_strides[4];
  int64_t src_peer_stride;  /* bytes between the item sets of consecutive destinations */
  int64_t dst_offset;       /* bytes: where this rank's contribution starts in every destination buffer */
  int32_t item_bytes;       /* multiple of 16 */
  uint64_t* flags[FG_MAX_PEERS];
  uint64_t* state;
  const uint64_t* wait_flags; /* optional: this rank's own flag array of the channel -- the put then also WAITS (like      */
  int64_t wait_timeout_ns;    /* fg_peer_wait) before the kernel ends: the exchange is one launch instead of two           */
} fg_peer_put_params;

int fg_peer_put(const fg_peer_put_params* p, fg_stream_t stream);
/* timeout_ns > 0: give up after that long without the flags arriving and set state[3] = 1 (nothing hangs) */
int fg_peer_wait(const uint64_t* flags, int32_t peers, uint64_t* state, int64_t timeout_ns, fg_stream_t stream);

/* GEGLU feed-forward gate of the SDXL transformer blocks (diffusers GEGLU, used by UNet2DConditionModel which the
 * reference wraps at fastgen/networks/SDXL/network.py:210-230): y[r][j] = h[r][j] * gelu_erf(h[r][inner + j]), bf16. */
int fg_geglu(const void* h, int64_t h_ld, int32_t rows, int32_t inner, void* y, int64_t y_ld, fg_stream_t stream);
/* out[i][g][:] = a[i][:] + b[g][:], fp32: scale_shift_table of every block + timestep projection of every frame. */
int fg_bcast_add(const float* a, const float* b, int32_t na, int32_t nb, int32_t d, float* out, fg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Sampler-step kernels (fastgen/methods/model.py:346-370, noise_schedule.py:72-88,425-449,544-574,
 * EDM/network.py:755-805).  x tensors are the caller's NCHW latents; `f` is the network's raw output in
 * channels-last fp32 with pixel stride f_ld.  Per-sample coefficients are fp64 [n] device arrays computed
 * by the host mirror of the noise schedule, so they are bit-identical to the reference's.
 * ---------------------------------------------------------------------------------------------------- */
/* xin[n,h,w,c_pad] (bf16, zero padded channels; optional low part) = cast(c_in[n]) * x[n,c,h,w]  (:771-773) */
int fg_edm_input(const void* x, int32_t x_dtype, int32_t n, int32_t c, int32_t hw, const double* c_in,
                 void* xin, void* xin_lo, int32_t c_pad, fg_stream_t stream);

/* x0 = cast(c_skip)*x_t + cast(c_out)*F in x dtype (:800-805);  if eps != NULL also
 * x_next = cast_x( double(x0)*alpha_next + double(eps)*sigma_next )  (noise_schedule.py:441-449).
 * eps == NULL and ode != 0: eps is derived as (x_t - x0*alpha_cur)/clamp(sigma_cur) (:544-574) first. */
typedef struct {
  const void* f;        /* fp32 [N,H,W,f_ld] raw network output */
  int64_t f_ld;
  const void* x_t;      /* NCHW, x_dtype */
  const void* eps;      /* NCHW, x_dtype, or NULL */
  int32_t x_dtype;
  int32_t n, c, hw;
  const double* c_skip; /* [n] or NULL (drop_precond output): x0 = F */
  const double* c_out;  /* [n] */
  int32_t ode;
  const double* alpha_cur;  /* [n], ode only */
  const double* sigma_cur;  /* [n], ode only (already non-zero-clamped by the host) */
  const double* alpha_next; /* [n] or NULL: no re-noising (last step) */
  const double* sigma_next; /* [n] */
  void* x0;             /* NCHW out */
  void* x_next;         /* NCHW out or NULL */
} fg_edm_step_params;

int fg_edm_step(const fg_edm_step_params* p, fg_stream_t stream);

/* out[i] = cast( double(a[i])*ca[k] + double(b[i])*cb[k] ), k = (i / outer) * mid + (i / inner) % mid; b may be NULL.
 * Per-sample coefficients: outer = elements per sample, mid = 1, inner = outer.  Per-frame coefficients of a
 * [B,C,T,H,W] latent (t of shape [B,1,T,1,1], causvid.py:55-56): outer = C*T*H*W, mid = T, inner = H*W.
 * Covers latents() (noise*sigma), forward_process (x*alpha + eps*sigma), flow_to_x0 (x - v*t), the
 * MeanFlow update (x - dt*u) -- all "fp64 arithmetic, one rounding" in the reference. */
int fg_axpby(const void* a, const void* b, int32_t dtype, int64_t total, int64_t outer, int64_t mid, int64_t inner,
             const double* ca, const double* cb, void* out, fg_stream_t stream);

/* (a - b*cb[k]) / cd[k] in fp64, one rounding: x0_to_eps / eps_to_x0 (noise_schedule.py:544-608). */
int fg_sub_div(const void* a, const void* b, int32_t dtype, int64_t total, int64_t outer, int64_t mid, int64_t inner,
               const double* cb, const double* cd, void* out, fg_stream_t stream);

/* Select the vectorised forms of fg_edm_step / fg_axpby / fg_sub_div (4 pixels x all channels, resp. 16 bytes, per thread
 * and iteration; same arithmetic, bit-identical results) where shapes and alignment allow: enable = 1 / 0, or -1 to only
 * query.  Returns the previous setting.  Default: on (the environment variable FASTGEN_B200_SAMPLER_VEC=0 turns
 * them off); the scalar forms remain for shapes the vector forms cannot take and for the equality tests. */
int fg_sampler_vec(int enable);

/* cudaMemsetAsync(ptr, 0, bytes) as a plan step (statistics arenas) */
int fg_zero(void* ptr, int64_t bytes, fg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Hand-off either side of the sampler: the arithmetic the reference's VAE / text-encoder WRAPPERS do
 * around the external networks, one streaming pass each, rounded like the reference's eager ops (bit-identical).
 * ---------------------------------------------------------------------------------------------------- */
/* x, out: [outer][c][inner] (fp32 | bf16); a, b: [c] fp32 holding values already rounded to x's dtype.
 * mode 0: (x / a) + b   -- WanVideoEncoder.decode: a = 1 / std, b = mean (fastgen/networks/Wan/network.py:515-528)
 * mode 1: (x - b) * a   -- WanVideoEncoder.encode (:496-513)
 * mode 2: x * a         -- SDXLImageEncoder.encode / .decode: the VAE scaling factor or its reciprocal, c = 1
 *                          (fastgen/networks/SDXL/network.py:141-149)
 * every op rounds to x's dtype, the result is cast to out's dtype (the reference's .to(vae.dtype)). */
int fg_latent_affine(const void* x, int32_t x_dtype, void* out, int32_t out_dtype, const float* a, const float* b,
                     int64_t outer, int32_t c, int64_t inner, int32_t mode, fg_stream_t stream);
/* x <- min(max(x, lo), hi) in place, NaN kept: the .clip_(-1.0, 1.0) behind both VAE decoders (Wan/network.py:527,
 * SDXL/network.py:147) */
int fg_clamp(void* x, int32_t dtype, int64_t total, float lo, float hi, fg_stream_t stream);
/* out[s, t, :] = t < len[s] ? x[s, t, :] : 0 for x [b][rows][d] (row stride x_ld, batch stride x_bs, elements), out
 * contiguous: WanTextEncoder.encode's "u[:v] padded with zeros to 512 rows" + the cast to `precision`
 * (Wan/network.py:463-470).  len: device int64 [b]. */
int fg_zero_tail_rows(const void* x, int32_t x_dtype, int64_t x_ld, int64_t x_bs, const int64_t* len, int32_t b,
                      int32_t rows, int32_t d, void* out, int32_t out_dtype, fg_stream_t stream);

/* small helpers used by the embedding path */
/* PositionalEmbedding (:306-319): out[n, :] = [cos(t*f) | sin(t*f)] (optionally swapped), fp32 */
int fg_pos_embedding(const float* t, int32_t n, int32_t channels, int32_t max_positions, int32_t endpoint,
                     int32_t swap_sin_cos, float* out, int64_t out_ld, fg_stream_t stream);
/* y = act(x + add) split into bf16 hi/lo;  act: 0 none, 1 SiLU.  x fp32 [n][c] */
int fg_split_bf16(const float* x, int64_t x_ld, const float* add, int64_t add_ld, int32_t n, int32_t c, int32_t act,
                  void* hi, void* lo, int64_t out_ld, float* y_f32, int64_t y_ld, fg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTGEN_B200_H_ */
