/* svdpp.h — C ABI of libsvdpp.so: the sm_100a kernels behind the SVD denoising-step hot path.
 *
 * The reference (inai17ibar/video-diffusion-pipeline-parallel) has no FFI: its boundary for this
 * path is the Python call  unet(sample, timestep, encoder_hidden_states, added_time_ids)  at
 * src/models/svd_unet.py:389-395,400-406,416-422 plus the elementwise step maths at :382 and
 * :427-439.  Each entry point below replaces the library kernels (cuDNN/cuBLAS/SDPA/ATen) that
 * those lines reach; the comment on each one names the reference lines it stands in for.
 *
 * Conventions
 *  - every pointer is a CUDA device pointer owned by the caller (a torch tensor); fp16 unless
 *    stated; the library never allocates, frees or retains memory past stream completion;
 *  - every call enqueues on `stream` (a cudaStream_t passed as void*), never synchronises, and is
 *    CUDA-graph capturable;
 *  - return 0 on success, negative on error; svdpp_last_error() returns a thread-local message;
 *  - activations are channels-last: a [B*F, H, W, C] tensor is the row-major matrix [M = B*F*H*W, C].
 */
#ifndef SVDPP_H_
#define SVDPP_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVDPP_ABI_VERSION 1
#define SVDPP_MAX_TAPS 9

typedef void* svdpp_stream;

int svdpp_abi_version(void);
const char* svdpp_last_error(void);
/* compute capability and SM count of the current device; <0 if no usable device */
int svdpp_device_info(int* sm_major, int* sm_minor, int* num_sms);
/* Tuning switches of the kernels (process-wide; initial values come from the environment when the library is
 * loaded: SVDPP_NO_TMA_STORE=1 -> "tma_store" = 0, SVDPP_NO_TMA_R1=1 -> "tma_r1" = 0, SVDPP_PDL=1 -> "pdl" = 1).  They change HOW a result is
 * produced, never the result (tests/test_gpu_kernels.py runs both settings against the oracle):
 *   "tma_store"  1: GEMM output tiles leave through TMA tensor stores; 0: per-thread copy loop
 *   "tma_r1"     1 (with tma_store): the residual tile R1 reaches the epilogue through TMA tensor loads; 0: per-thread loads
 *   "epi_dma"    1: GEMMs with 256/128-wide tiles and K <= 64 * "epi_dma_max_kb" (default 5) run their epilogue I/O on a
 *                   dedicated DMA lane with two staging tiles; 2: wherever possible; 0: never (SVDPP_EPI_DMA, SVDPP_EPI_DMA_MAX_KB)
 *   "two_prod"   1 (default): GEMMs with K >= 64 * "two_prod_min_kb" (default 6) that do not use the DMA lane fill their smem
 *                   ring from TWO producer warps taking alternate k-blocks (one issuing thread sustains a k-block per
 *                   ~500-590 clocks whatever its size - tools/ubench/fill.cu - which capped the 128 / 160 / 256-wide
 *                   tiles); 0: one producer warp (SVDPP_TWO_PROD)
 *   "splitk"     1: impl 6 splits the tiles of a short last wave along K when the descriptor carries splitk_ws; 0: never
 *                   ("splitk_min_kb": fewest 64-wide k-blocks a slice may get, default 4; "splitk_min_total_kb": only for
 *                   K / 64 >= this, default 64)
 *   "r1_prefetch_max_kb"  the GEMM producer warp prefetches the residual tile into L2 for K / 64 <= this (default 5; 0: never)
 *   "fmha_stagger"  SM clocks by which the second query tile of the two-tile FMHA starts behind the first (default 900; SVDPP_FMHA_STAGGER)
 *   "reverse"    traversal direction of the NEXT GEMM / LayerNorm / GroupNorm launches: 1 = from the last rows to the first (the end
 *                   of a tensor a forward producer has just written is still in L2); GroupNorm: 1 = statistics from the end / apply
 *                   forward, 2 = statistics forward / apply from the end; 0 (default) = forward.  A GEMM whose last wave is split
 *                   along K keeps the forward order.
 *   "zigzag"     1: svdpp_unet_* flip that direction from producer to consumer along the network (default 0: measured -0.4 ms
 *                   of 93.9 ms per step, inside the noise; SVDPP_ZIGZAG)
 *   "pdl"        1: kernels are launched with programmatic stream serialisation (the prologue of kernel N+1
 *                   overlaps the tail of kernel N; every kernel waits on griddepcontrol before touching memory)
 * set returns 0, or -1 for an unknown key; get returns the value, or -1 for an unknown key.
 * No counterpart in the reference (its kernels are cuDNN/cuBLAS/ATen picks made by torch). */
int svdpp_set_tuning(const char* key, int value);
int svdpp_get_tuning(const char* key);

/* ---------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution with fused epilogue (tcgen05 + TMEM + TMA).
 *   acc[m, n] = sum_k A[m, k] * Wt[n, k]                         (fp16 inputs, fp32 accumulate)
 *   y[m, n]   = alpha * (acc + bias[n] + rowvec[rv(m), n]) + beta1 * R1[m, n] + beta2 * R2[m, n]
 *   geglu: Wt rows are interleaved per tile as [half value | half gate] (80 + 80 for 160-wide tiles, 128 + 128 for impl 3);
 *          D[m, j] = y_value[m, j] * gelu(y_gate[m, j]),  D has N/2 columns.
 * A is either a plain matrix (optionally the K-concatenation [A | A2], split at K1), or, in conv
 * mode, the channels-last activation [cB, cF, cH, cW, cC] read through `ntaps` shifted windows
 * (tap t contributes K-slice [t*cC, (t+1)*cC); out-of-range pixels/frames read as zero); with conv_stride = 2 the
 * windows are strided (the down-sampling Conv2d 3x3 stride 2 pad 1), still without materialising im2col.
 * Replaces: nn.Linear / Conv2d 3x3 / Conv3d (3,1,1) / 1x1 shortcut / GEGLU inside
 * UNetSpatioTemporalConditionModel (called at svd_unet.py:389-395).
 * Requirements: K % 64 == 0, K1 % 64 == 0, N a multiple of the tile width (pad Wt), 16-byte aligned rows.
 * -------------------------------------------------------------------------------------------*/
typedef struct svdpp_gemm_desc {
  int32_t M, N, K;
  const void* A;   int64_t lda;          /* row pitch in elements */
  const void* A2;  int64_t lda2; int32_t K1;   /* optional second source for k >= K1 */
  int32_t conv;                           /* 0: matrix A; 1: shifted-window activation */
  int32_t cB, cF, cH, cW, cC;
  int32_t ntaps;
  int8_t  taps[SVDPP_MAX_TAPS][4];        /* (dw, dh, df, 0) per tap */
  const void* Wt;  int64_t ldw;           /* [N, K], K contiguous */
  const void* bias;                       /* [N] or NULL */
  const void* rowvec; int64_t rv_ld;      /* [rows, N] or NULL; row = ((m / rv_hw) / rv_div) % rv_mod */
  int32_t rv_hw, rv_div, rv_mod;
  const void* R1;  int64_t ldr1; float beta1;
  const void* R2;  int64_t ldr2; float beta2;
  float alpha;
  int32_t geglu;
  void* D;         int64_t ldd;
  int32_t n_store;                        /* store columns n < n_store (after geglu halving) */
  /* strided convolution (conv mode): output pixel (h, w) reads input (h*conv_stride + dh, w*conv_stride + dw).
   * cH, cW above are then the OUTPUT extent (cB*cF*cH*cW == M) and cHin, cWin the input extent.
   * conv_stride 0 or 1: plain convolution, cHin/cWin ignored. */
  int32_t conv_stride, cHin, cWin;
  /* sub-pixel output map (conv mode): out_up = 2 stores GEMM row (img, h, w) of the cH x cW grid at pixel
   * (2h + out_up_y, 2w + out_up_x) of a (2 cH) x (2 cW) output image (D has cB*cF*4*cH*cW rows).  Used to run
   * "nearest-neighbour 2x upsample + Conv2d 3x3" as four 2x2-tap convolutions on the low-resolution input, one per
   * output parity, with pre-summed weights: 4/9 of the FLOPs and no upsampled tensor.  0 or 1: off. */
  int32_t out_up, out_up_y, out_up_x;
  /* optional scratch for the split-K tail of impl 6 (256x320 pair tiles): when the last wave of tiles would keep fewer
   * than half of the CTA pairs busy, its tiles are split along K over the idle pairs and reduced through this buffer
   * (fp32, slice order: deterministic).  >= 4096 + 74 * 327680 bytes covers every shape; its first 4096 bytes are
   * counters that must be zero before the first use (the kernel re-arms them).  NULL: no split. */
  void* splitk_ws; int64_t splitk_ws_bytes;
} svdpp_gemm_desc;

/* impl selects the tile shape of the tcgen05 kernel (Wt must be padded to a multiple of the tile's N):
 *   0 = one CTA per 128x160 tile           4 = one CTA per 128x128 tile (no GEGLU)
 *   2 = CTA pair (cta_group::2), 256x160   3 = CTA pair, 256x256 (GEGLU rows interleaved [128 value | 128 gate])
 *   6 = CTA pair, 256x320 as two N=160 MMAs per k-step, halves rotating through three TMEM buffers (no GEGLU)
 *   5 = impl 3 with 8 instead of 16 GEGLU epilogue warps (A/B measurements)
 *   7 = CTA pair, 256x128 (no GEGLU): impl 4's width, half the B-operand traffic per SM
 *   1 = plain CUDA-core kernel (slow; bring-up cross-check) */
int svdpp_gemm_f16(const svdpp_gemm_desc* d, int impl, svdpp_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Fused transformer feed-forward (C <= 320): y = epilogue( GEGLU(X W1^T + b1) W2^T + b2 ) in ONE kernel - the [M, 4C]
 * gated intermediate stays in tensor memory (tcgen05.mma with the A operand in TMEM), never in HBM.
 *   W1 [8C, C], b1 [8C]: the GEGLU projection with rows interleaved per 128 as [64 value | 64 gate] (value rows j, gate
 *                        rows 4C + j of the diffusers weight, for j = 64 t .. 64 t + 63 in tile t)
 *   W2 [w2_rows >= C, 4C] (row pitch ldw2), b2 [C]: ff.net.2 as stored by diffusers
 *   y[m, n] = alpha * (acc + b2[n] + rowvec[rv(m), n]) + beta1 * R1[m, n] + beta2 * R2[m, n]      (as svdpp_gemm_f16)
 * fp16 roundings as in the two-kernel path (projection, gelu(gate) and their product each rounded to fp16).
 * Replaces: FeedForward(GEGLU) of BasicTransformerBlock / TemporalBasicTransformerBlock inside the UNet called at
 * reference src/models/svd_unet.py:389-395 (two svdpp_gemm_f16 calls otherwise).
 * -------------------------------------------------------------------------------------------*/
typedef struct svdpp_ff_desc {
  int32_t M, C;
  const void* X;   int64_t ldx;           /* [M, C] */
  const void* W1;  const void* b1;        /* [8C, C] interleaved, [8C] interleaved */
  const void* W2;  int64_t ldw2; int32_t w2_rows;
  const void* b2;
  const void* rowvec; int64_t rv_ld;      /* [rows, C] or NULL; row = ((m / rv_hw) / rv_div) % rv_mod */
  int32_t rv_hw, rv_div, rv_mod;
  const void* R1;  int64_t ldr1; float beta1;
  const void* R2;  int64_t ldr2; float beta2;
  float alpha;
  void* D;         int64_t ldd;           /* [M, C] */
} svdpp_ff_desc;
int svdpp_ff_geglu_f16(const svdpp_ff_desc* d, svdpp_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Spatial self-attention over S = H*W tokens per image, head_dim 64, no mask (tcgen05 FMHA).
 * q/k/v are column blocks of one [n_img*S, ld] matrix: head h of q at columns q_off + 64*h, etc.
 * Replaces: attn1 of BasicTransformerBlock (SDPA) inside the UNet.
 * -------------------------------------------------------------------------------------------*/
typedef struct svdpp_attn_desc {
  const void* qkv; int64_t ld;
  int32_t q_off, k_off, v_off;
  void* out;       int64_t ldo;           /* [n_img*S, heads*64] */
  int32_t n_img, S, heads;
  float scale;                            /* 1/sqrt(64) */
} svdpp_attn_desc;
/* impl: 0 = one 128-query tile per CTA, two CTAs per SM (short sequences); 2 = two query tiles per CTA, P in tensor memory
 * (round 1); 3 = 2 with two threads per row; 4 = 2 with the quarter-pipelined softmax (row maximum of the next 32 keys
 * computed under the exponentials of the current 32); 5 / 6 = 4 with every 8th / 4th group of four exponentials evaluated
 * as an FMA-pipe polynomial; 7 = two query tiles per CTA whose exponential phases alternate through named barriers
 * ("ping-pong", the default for S >= 1024 inside svdpp_unet_*; "fmha_handover" = batches of 16 exponentials before the end
 * of a turn at which the partner warp is released); 8 = 7 with two threads per row (16 softmax warps, the two half-row
 * warps of a tile take their turn together; measured slower, kept selectable); 1 = plain CUDA-core kernel (bring-up
 * cross-check).
 * "fmha_stagger" (svdpp_set_tuning): SM clocks by which query tile 1 of impl 2..6 starts behind tile 0. */
int svdpp_attn_spatial_f16(const svdpp_attn_desc* d, int impl, svdpp_stream stream);
/* Debug: device buffer (>= 2 * ceil(S/128) * 8 uint32) that impl 2..6 fill with clock stamps of the softmax phases of two
 * warps of their first CTA (tools/attn_trace.py); NULL switches it off.  No counterpart in the reference. */
int svdpp_debug_attn_trace(void* dev_buffer);

/* Temporal self-attention: for every (batch, pixel, head) a sequence of F <= 32 frames.
 * Token (b, f, p) is row (b*F + f)*HW + p.  Replaces attn1 of TemporalBasicTransformerBlock. */
int svdpp_attn_temporal_f16(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off,
                            void* out, int64_t ldo, int32_t B, int32_t F, int32_t HW, int32_t heads,
                            float scale, svdpp_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Bandwidth kernels.
 * -------------------------------------------------------------------------------------------*/
/* GroupNorm(32 groups) [+ SiLU] over channels-last input that may be the channel concatenation
 * [x1 (C1) | x2 (C2)] (the up-block skip cat); statistics per (image, group), or per
 * (frames_per_stat consecutive images, group) for the temporal ResBlock's 5-D GroupNorm.
 * out is [n_img*HW, C1+C2].  workspace >= svdpp_groupnorm_workspace_bytes(), its first 32 KB ZERO-FILLED before the first
 * use (it holds arrival counters that every call leaves at zero again).  Deterministic: fixed summation order.
 * Replaces: nn.GroupNorm + SiLU in ResnetBlock2D / TemporalResnetBlock / transformer norm / conv_norm_out. */
size_t svdpp_groupnorm_workspace_bytes(int32_t n_img, int32_t HW);
int svdpp_groupnorm_silu(const void* x1, int32_t C1, const void* x2, int32_t C2, const void* gamma,
                         const void* beta, void* out, int32_t n_img, int32_t HW, int32_t frames_per_stat,
                         float eps, int32_t apply_silu, void* workspace, size_t workspace_bytes,
                         svdpp_stream stream);

/* LayerNorm over C of (x[m, :] + addvec[((m / add_hw) % add_mod), :]) with affine; addvec may be NULL.
 * Replaces: nn.LayerNorm in (Temporal)BasicTransformerBlock; the add is the frame-position embedding. */
int svdpp_layernorm(const void* x, int64_t ldx, const void* addvec, int32_t add_hw, int32_t add_mod,
                    const void* gamma, const void* beta, void* out, int64_t ldo, int32_t M, int32_t C,
                    float eps, svdpp_stream stream);

/* y[r, n] = act_out( sum_k act_in(x[r, k] + x_add[r, k]) * W[n, k] + bias[n] ) for a handful of rows
 * (embedding MLPs, time_emb_proj, 1-token cross-attention value path). x_add may be NULL (same pitch
 * as x). act: 0 none, 1 SiLU. */
int svdpp_linear_small(const void* x, const void* x_add, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* y,
                       int64_t ldy, int32_t R, int32_t N, int32_t K, int32_t act_in, int32_t act_out,
                       svdpp_stream stream);

/* n_groups independent small linears in one launch: for group g (a device array of descriptors),
 *   y[r, y_off + n] = sum_k x[r, x_off + k] * W[n, k] + bias[n],  n < N, k < K  (K, x_off multiples of 8).
 * Used for the output projections of the 1-token cross-attention of all transformer blocks at once
 * (attn2.to_out of BasicTransformerBlock / TemporalBasicTransformerBlock). max_n = max over groups of N. */
typedef struct svdpp_small_group {
  const void* W;      /* device, [N, K] fp16, K contiguous */
  const void* bias;   /* device, [N] fp16 or NULL */
  int32_t x_off, y_off, N, K;
} svdpp_small_group;
int svdpp_linear_small_grouped(const void* x, int64_t ldx, const svdpp_small_group* groups_dev, int32_t n_groups,
                               int32_t max_n, void* y, int64_t ldy, int32_t R, svdpp_stream stream);

/* Sinusoidal embedding [cos | sin] (flip_sin_to_cos, shift 0) of n_vals scalars into [n_vals, dim] fp16.
 * src_kind: 0 = fp32 device array, 1 = fp16 device array, 2 = the integers (i % src_mod). */
int svdpp_sinusoid_embed(const void* src, int32_t src_kind, int32_t src_mod, int32_t n_vals, int32_t dim,
                         void* out, svdpp_stream stream);

/* Nearest-neighbour 2x upsample, channels-last (Upsample2D before its conv). */
int svdpp_upsample2x_nhwc(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                          svdpp_stream stream);

/* Gather shifted windows into a matrix: out[m_out, t*C + c] = x[b, f+df, h*stride+dh, w*stride+dw, c]
 * (zero outside), columns [ntaps*C, ldo) zero-filled.  Used for stride-2 down-sample convs, conv_in
 * (C = 8) and image widths the TMA window path does not tile. */
int svdpp_im2col_nhwc(const void* x, void* out, int64_t ldo, int32_t B, int32_t F, int32_t H, int32_t W,
                      int32_t C, int32_t Ho, int32_t Wo, int32_t stride, int32_t ntaps,
                      const int8_t* taps /* [ntaps][4] host */, svdpp_stream stream);

/* Build the UNet's channels-last input [B*F, H, W, C0+C1]:
 *   out[.., c] = fp16(float(src0[b, f, c, h, w]) / in_div)  for c < C0,  src1[...] for c >= C0.
 * Element strides (in elements) are given per source so both [B,C,F,H,W] (svd_unet.py:382-388:
 * scale_model_input, cat, permute) and [B,F,C,H,W] (the UNet operator's `sample`) can be read. */
int svdpp_pack_unet_input(const void* src0, int64_t s0_b, int64_t s0_f, int64_t s0_c, int32_t C0, float in_div,
                          const void* src1, int64_t s1_b, int64_t s1_f, int64_t s1_c, int32_t C1, void* out,
                          int32_t out_bfchw /* 0: channels-last, 1: [B,F,C0+C1,H,W] */,
                          int32_t B, int32_t F, int32_t H, int32_t W, svdpp_stream stream);

/* Channels-last [B*F, H, W, C] -> [B, F, C, H, W] (the UNet operator's return layout). */
int svdpp_nhwc_to_bfchw(const void* x, void* out, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                        svdpp_stream stream);

/* Classifier-free-guidance combine + v-prediction Euler update, svd_unet.py:410-411,425-439:
 *   v      = v_cond == NULL ? v_a : fp16(v_a + fp16(gs[f] * fp16(v_cond - v_a)))        (fp16 ops)
 *   x0     = v * c_v + x / c_x ;  d = (x - x0) / sigma ;  out = fp16(x + d * dt)          (fp32 ops)
 * with c_v = -sigma/sqrt(sigma^2+1), c_x = sigma^2+1 computed by the caller in fp32.
 * latent/out are [B, C, F, H, W]; v_* are channels-last [B, F, H, W, C] if v_nhwc else [B, F, C, H, W]. */
int svdpp_euler_vpred_step(const void* latent, const void* v_a, const void* v_cond, const void* gs,
                           int32_t v_nhwc, float c_v, float c_x, float sigma, float dt, void* out,
                           int32_t B, int32_t C, int32_t F, int32_t H, int32_t W, svdpp_stream stream);

/* Stage-to-stage latent handoff over peer-mapped memory (NVLink), replacing the dist.send / dist.recv pair of reference
 * src/pipeline/pipeline.py:75-84.  The producer's LAST local step writes its result straight into a buffer that lives on
 * the next stage's GPU (`out` of svdpp_euler_vpred_step_signal / svdpp_unet_step_handoff is that peer-mapped pointer) and
 * the same kernel raises the consumer's flag when all of its stores are visible; the consumer's stream waits for the flag
 * with svdpp_flag_wait before its first kernel touches the buffer, and hands the slot back with svdpp_flag_set on the
 * producer's acknowledge flag.  No NCCL kernel, no lock-step between ranks, no extra copy.
 *   done_counter  LOCAL device uint32, zero before the first use (the kernel re-arms it)
 *   ready_flag    uint32 in the CONSUMER's memory (peer-mapped); set to flag_value by a release store at system scope
 * svdpp_flag_wait: one thread spins (acquire, system scope) until *flag == value, then stores reset_to if reset != 0; after
 * timeout_s seconds (<= 0: 600) it prints a message and traps, so a protocol bug becomes a CUDA error, not a hung GPU. */
typedef struct svdpp_handoff {
  void* done_counter;
  void* ready_flag;
  uint32_t flag_value;
} svdpp_handoff;
int svdpp_euler_vpred_step_signal(const void* latent, const void* v_a, const void* v_cond, const void* gs,
                                  int32_t v_nhwc, float c_v, float c_x, float sigma, float dt, void* out,
                                  int32_t B, int32_t C, int32_t F, int32_t H, int32_t W, const svdpp_handoff* ho /* or NULL */,
                                  svdpp_stream stream);
int svdpp_flag_wait(void* flag, uint32_t value, int32_t reset, uint32_t reset_to, int32_t timeout_s, svdpp_stream stream);
int svdpp_flag_set(void* flag, uint32_t value, svdpp_stream stream);

/* ---------------------------------------------------------------------------------------------
 * VAE front / back end (AutoencoderKLTemporalDecoder; reference scripts/generate_video_demo.py:119-143 encode, :154-195
 * decode).  Its convolutions and linears are svdpp_gemm_f16 calls; these are the pieces in between.
 * -------------------------------------------------------------------------------------------*/
/* In-place softmax over each row of an fp16 matrix [rows, n] (row pitch ld): x <- softmax(scale * x), fp32 maths; columns
 * >= n_valid (0: all n) are padding keys: ignored and written as 0.  The single-head, head_dim-512 attention of the VAE mid
 * blocks and the 257-token attention of the CLIP image encoder = GEMM (Q K^T) -> this -> GEMM (P V). */
int svdpp_softmax_rows(void* x, int64_t ld, int32_t rows, int32_t n, int32_t n_valid, float scale, svdpp_stream stream);
/* Whole attention of a short sequence in one launch: softmax(scale * Q K^T) V for every (image, head), S <= 512 keys held in
 * shared memory, fp32 maths on fp16 inputs.  Replaces the CLIPAttention of the image encoder (transformers
 * CLIPVisionModelWithProjection: 257 tokens, 16 heads of width 80; reference scripts/generate_video_demo.py:112-117), which as
 * a per-head GEMM -> softmax -> transpose -> GEMM chain cost 2048 launches per image.  qkv is one matrix (row pitch ld) with
 * head h of q / k / v at columns {q,k,v}_off + h * head_stride; every image owns S_pad rows of which the first S are tokens.
 * out[row, h * out_head_stride + d]: d < head_dim the result, head_dim <= d < out_head_stride zeros; rows of padding tokens
 * zeros.  head_dim % 8 == 0, <= 128; out_head_stride <= 128. */
int svdpp_attn_small_f16(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, int32_t head_stride,
                         void* out, int64_t ldo, int32_t out_head_stride, int32_t n_img, int32_t S, int32_t S_pad,
                         int32_t heads, int32_t head_dim, float scale, svdpp_stream stream);
/* out[c, r] = in[r, c] for an fp16 matrix [R, C] (V -> V^T, the K-major B operand of the P V product). */
int svdpp_transpose_f16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t R, int32_t C, svdpp_stream stream);
/* TemporalDecoder.time_conv_out: Conv3d (3,1,1) over the frames of a 3-channel video, reading the channels-last output of
 * conv_out [B, F, HW, x_channels >= 3] (the first 3 channels; conv_out is stored 4 wide so that its rows are 8-byte
 * aligned) and writing the caller's [B*F, 3, HW] (fp32 if out_fp32 else fp16).  w is [3, 3, 3] = (co, ci, kt). */
int svdpp_time_conv_out(const void* x, int32_t x_channels, const void* w, const void* bias, void* out, int32_t out_fp32,
                        int32_t B, int32_t F, int64_t HW, svdpp_stream stream);

/* Output format of the image -> video run (reference scripts/generate_video_demo.py:198-222: frames -> uint8 -> file).
 * frames: one video [3, F, H, W] in [-1, 1], fp32 (is_fp32) or fp16, element (c, f, y, x) at c * stride_c + f * stride_f +
 * y * W + x (the permuted view decode_latents returns: stride_c = H * W, stride_f = 3 * H * W).  rgb (or null): packed bytes
 * [F, H, W, 3] = ((v + 1) / 2 * 255).clamp(0, 255) truncated, bit-identical to the reference's torch expression.  idx (or
 * null): [F, H, W] indices r * 42 + g * 6 + b into the fixed 6 x 7 x 6 colour cube (levels k * 255 / 5, k * 255 / 6,
 * k * 255 / 5), with a 4 x 4 ordered dither when dither != 0 - a GIF frame that needs no palette search on the host.
 * W % 4 == 0, strides % 4 == 0. */
int svdpp_frames_to_bytes(const void* frames, int32_t is_fp32, int64_t stride_c, int64_t stride_f, int32_t F, int32_t H,
                          int32_t W, void* rgb, void* idx, int32_t dither, svdpp_stream stream);

/* DummyUNet step (reference src/models/dummy_unet.py:37-59), fp32, [B, C, F, H, W]:
 *   out = x + tanh_scale * conv3d(silu(conv3d(x, w1, b1)), w2, b2) + layernorm_C(x) */
int svdpp_dummy_unet_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* ln_g, const float* ln_b, float ln_eps, float tanh_scale, float* hidden_ws,
                          float* out, int32_t B, int32_t C, int32_t Ch, int32_t F, int32_t H, int32_t W,
                          svdpp_stream stream);

/* ---------------------------------------------------------------------------------------------
 * The whole UNet operator, and one whole denoising step, behind a handle.
 *
 * Replaces, in one call each:
 *   svdpp_unet_forward       unet(sample, timestep, encoder_hidden_states, added_time_ids, return_dict=False)[0]
 *                            - diffusers UNetSpatioTemporalConditionModel.forward as the reference calls it at
 *                            src/models/svd_unet.py:389-395,400-406,416-422 (boundary B2 of SURVEY section 8b)
 *   svdpp_unet_forward_nhwc  the same with channels-last input / output (what the fused step uses internally)
 *   svdpp_unet_step          StableVideoUNet.forward(latent, step), src/models/svd_unet.py:351-439: scale_model_input + cat +
 *                            permute, the UNet (both classifier-free-guidance branches as one batch of 2), the guidance
 *                            combine and the v-prediction Euler update
 * The handle owns the packed weights (device memory allocated in svdpp_unet_load_weights, freed in svdpp_unet_destroy).
 * Activations live in a caller-provided workspace of at least svdpp_unet_workspace_bytes(B, F, H, W) bytes (enough for
 * batch B without and with guidance); its contents need not be preserved between calls.  The forward never allocates,
 * never synchronises and is CUDA-graph capturable - except the FIRST call with a new frame count F, which computes and
 * caches the frame-position embeddings (one cudaMalloc per transformer block): run one eager call per F first.
 * One handle per device and host thread; not thread-safe.
 * -------------------------------------------------------------------------------------------*/
#define SVDPP_UNET_MAX_LEVELS 8
typedef struct svdpp_unet svdpp_unet;          /* opaque */

/* Architecture (diffusers unet/config.json) + GroupNorm eps per block class + kernel choices. */
typedef struct svdpp_unet_config {
  int32_t in_channels, out_channels;           /* 8, 4 */
  int32_t n_levels;                            /* 4 */
  int32_t block_out_channels[SVDPP_UNET_MAX_LEVELS];     /* 320, 640, 1280, 1280 */
  int32_t down_attn[SVDPP_UNET_MAX_LEVELS];              /* 1, 1, 1, 0: CrossAttnDownBlockSpatioTemporal vs DownBlockSpatioTemporal */
  int32_t num_attention_heads[SVDPP_UNET_MAX_LEVELS];    /* 5, 10, 20, 20 (head_dim 64) */
  int32_t layers_per_block;                    /* 2 */
  int32_t cross_attention_dim;                 /* 1024 */
  int32_t addition_time_embed_dim;             /* 256 */
  int32_t projection_class_embeddings_input_dim;  /* 768 */
  float eps_down_attn, eps_down, eps_mid, eps_up, eps_transformer, eps_out;   /* 1e-6, 1e-5, 1e-5, 1e-6, 1e-6, 1e-5 */
  int32_t gemm_impl;                           /* 3: choose the tile shape per GEMM (default); else force svdpp_gemm_f16's impl */
  int32_t attn_impl;                           /* -1: by sequence length (default); else force svdpp_attn_spatial_f16's impl */
  int32_t attn_impl_long;                      /* impl for S >= 1024 when attn_impl < 0 (0: library default) */
} svdpp_unet_config;

/* One entry of a diffusers-layout state_dict: `name` is the diffusers key ("down_blocks.0.resnets.0.spatial_res_block.
 * conv1.weight", ...), `data` a contiguous DEVICE tensor; dtype 0 = fp16 (all weights), 1 = fp32 (accepted for the
 * time_mixer.mix_factor scalars). */
typedef struct svdpp_tensor_desc {
  const char* name;
  const void* data;
  int32_t ndim;
  int64_t shape[5];
  int32_t dtype;
} svdpp_tensor_desc;

int svdpp_unet_create(svdpp_unet** out, const svdpp_unet_config* cfg);
/* Repack the state_dict into kernel layouts (conv filters tap-major, q/k/v fused, GEGLU rows interleaved per tile, the
 * up-sampling convs as four pre-summed 2x2-tap parity filters, every N padded to the tile width).  The caller's tensors
 * are only read during the call (it synchronises the device once). */
int svdpp_unet_load_weights(svdpp_unet* u, const svdpp_tensor_desc* tensors, int n);
size_t svdpp_unet_weight_bytes(const svdpp_unet* u);
/* 0 on error (svdpp_last_error) */
size_t svdpp_unet_workspace_bytes(svdpp_unet* u, int B, int F, int H, int W);
/* sample [B, F, in_channels, H, W] -> out [B, F, out_channels, H, W]; enc [B, 1, cross_attention_dim]; added_time_ids [B, 3] */
int svdpp_unet_forward(svdpp_unet* u, const void* sample, float timestep, const void* enc, const void* added_time_ids,
                       void* out, void* workspace, size_t workspace_bytes, int B, int F, int H, int W, svdpp_stream stream);
/* x_in [B*F*H*W, in_channels] -> out [B*F*H*W, out_channels], both channels-last */
int svdpp_unet_forward_nhwc(svdpp_unet* u, const void* x_in, float timestep, const void* enc, const void* added_time_ids,
                            void* out, void* workspace, size_t workspace_bytes, int B, int F, int H, int W, svdpp_stream stream);
/* One denoising step.  latent / image_latents / uncond_image_latents / out: [B, C, F, H, W] with C = out_channels.
 * uncond_image_latents == NULL: no guidance, enc [B, 1, D], added_time_ids [B, 3].  Otherwise classifier-free guidance as
 * one batch of 2B: enc [2B, 1, D] and added_time_ids [2B, 3] hold the unconditional half first, gs [F] is the per-frame
 * guidance scale.  in_div = sqrt(sigma^2 + 1), c_v = -sigma / sqrt(sigma^2 + 1), c_x = sigma^2 + 1, dt = sigma_next - sigma
 * are computed by the caller in fp32 (svdpp_euler_vpred_step has the formulas); timestep = 0.25 * ln(sigma). */
int svdpp_unet_step(svdpp_unet* u, const void* latent, const void* image_latents, const void* uncond_image_latents,
                    const void* enc, const void* added_time_ids, const void* gs, float timestep, float in_div, float c_v,
                    float c_x, float sigma, float dt, void* out, void* workspace, size_t workspace_bytes, int B, int F, int H,
                    int W, svdpp_stream stream);
/* svdpp_unet_step whose result goes to `out` AND raises a flag when it is complete (see svdpp_handoff): `out` is then the
 * next stage's peer-mapped receive slot.  ho == NULL: identical to svdpp_unet_step. */
int svdpp_unet_step_handoff(svdpp_unet* u, const void* latent, const void* image_latents, const void* uncond_image_latents,
                            const void* enc, const void* added_time_ids, const void* gs, float timestep, float in_div,
                            float c_v, float c_x, float sigma, float dt, void* out, void* workspace, size_t workspace_bytes,
                            int B, int F, int H, int W, const svdpp_handoff* ho, svdpp_stream stream);
/* kernels launched by the last forward / step on this handle (bench.py's gpu_launches) */
long long svdpp_unet_last_launches(const svdpp_unet* u);
void svdpp_unet_destroy(svdpp_unet* u);

#ifdef __cplusplus
}
#endif
#endif /* SVDPP_H_ */
