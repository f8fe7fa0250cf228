/* b200ldm.h -- C-ABI of libb200ldm.so: hand-written sm_100a kernels for the LoRA-adapted AudioLDM
 * UNet denoising loop.
 *
 * The reference (2025-comprehensive-design/AudioLDM-with-LoRA) is pure Python and owns no FFI; every
 * FLOP of this path runs inside diffusers / peft / torch ATen.  Each entry point below therefore
 * cites the reference CALL SITE that reaches the replaced ATen op(s) and the third-party routine
 * that issues them (diffusers 0.32.2 / peft 0.13.2, pinned at /root/reference/requirements.txt:24,90).
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * stated otherwise; the caller owns every buffer; nothing allocates, nothing synchronises; launches
 * go to `stream` (a cudaStream_t passed as void*).  Activations are NHWC bf16 ([N, H, W, C], C
 * innermost) -- a [B, S, C] token matrix is the same memory.  Return 0 on success, a negative
 * B200_ERR_* code otherwise; b200_last_error() gives the message of the calling thread's last
 * failure.
 */
#ifndef B200LDM_H
#define B200LDM_H
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_DRIVER (-3)
#define B200_ERR_UNSUPPORTED (-4)

int b200_version(void);
const char* b200_last_error(void);
/* Profiling aid: with env B200_GEMM_DEBUG=4 b200_conv_gemm's CTA 0 records its SM cycle counter at fixed
 * points (kernel entry, after the prologue, first operands landed, last MMA committed, epilogue done, exit);
 * this copies the first n (<= 32) records of the most recent launch to host memory. */
int b200_debug_timeline(unsigned long long* host_out, int n);
/* TMA descriptors (CUtensorMap) are cached process-wide by the value of their inputs (pointer, shape, strides, box), so
 * eager callers that re-issue the same launches on the same buffers -- the autograd seam of the fine-tuning step, the
 * attention-processor seam -- do not re-encode six descriptors per GEMM launch.  Returns the number of cache hits so far. */
long b200_tmap_cache_hits(void);

/* Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (TMA-fed, TMEM accumulator).
 *   out[pix, n] = epi( sum_seg sum_tap sum_c A_seg[pix @ tap, c] * wpacked[n, k(seg,tap,c)] )
 * a0 [nb,h,w,c0] carries ntaps (1 or 9: 3x3, pad 1) taps; a1/a2 (nullable, c1/c2 channels) are extra
 * 1-tap K segments: the ResNet 1x1 conv_shortcut over cat([h, skip]) and the rank-r LoRA branch
 * [x | x.A^T] . [W | s.B]^T ride in the same accumulator.  wpacked is bf16 [n_pad, K] (K-major,
 * K = ntaps*c0 + c1 + c2).  Epilogue: + bias[n] + rowvec[image, n] (timestep/class embedding
 * projection) + residual[pix, n]; geglu: tile columns are [values | gates] and out = v * gelu(g).
 * stride == 2 keeps only even (h, w) pixels (Downsample2D).  out is bf16 or fp32, leading dim out_ld.
 * ksplit > 1 (small-M, long-K layers): K is split over ksplit CTAs per tile, fp32 partial sums go to
 * `workspace` (ksplit * nb*h*w * n_pad floats) and a second kernel reduces them in fixed order and
 * applies the epilogue (deterministic).  cta_pair: 0 never, 1 when the k-loop is long enough to pay for it, 2 always (if the shape allows):
 * clusters of two CTAs compute 256 x block_n tiles with
 * tcgen05.mma.cta_group::2 (each CTA stages half of the weight tile).  block_n: tile width, multiple of 64 (32 allowed for fp32 output).
 * Replaces F.conv2d / F.linear (+ peft lora.Linear.forward, + GEGLU, + residual adds) under
 * UNet2DConditionModel.forward: /root/reference/script/train/train_audioldm_lora.py:539-546,
 * /root/reference/script/inference/generate_audio.py:47-52 (LoRA config :21-29), /root/reference/app.py:14. */
int b200_conv_gemm(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h, int w,
                   int ntaps, int stride, const void* wpacked, int n_pad, int n_valid, const float* bias,
                   const float* rowvec, int rowvec_ld, const void* residual, int res_ld, void* out, int out_ld,
                   int out_fp32, int geglu, int block_n, int max_ctas, int ksplit, float* workspace, int cta_pair,
                   void* stream);

/* b200_conv_gemm (bf16 output, stride 1, no split-K, no GEGLU) whose epilogue ALSO leaves the partial GroupNorm statistics
 * of the tensor it stores, for the one-pass b200_groupnorm_apply that consumes it: gn_stat fp32
 * [nb, slabs, n_valid / 4, 2] = per image, per 32-pixel slab of a tile and per 4-channel unit the (sum, sum of squares) of
 * the stored bf16 values; slabs = b200_gn_stat_slabs(nb, h, w) (0: geometry unsupported).  Fixed slots, no atomics:
 * deterministic.  Every GroupNorm input of the UNet is the output of such a layer (ResnetBlock2D.conv1 / conv2,
 * Transformer2DModel.proj_out, Upsample2D.conv, conv_in -- diffusers, via train_audioldm_lora.py:539-546), so
 * F.group_norm's statistics pass over the activation disappears. */
int b200_conv_gemm_gnstat(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h, int w,
                          int ntaps, const void* wpacked, int n_pad, int n_valid, const float* bias, const float* rowvec,
                          int rowvec_ld, const void* residual, int res_ld, void* out, int out_ld, int block_n,
                          int max_ctas, int cta_pair, float* gn_stat, void* stream);
int b200_gn_stat_slabs(int nb, int h, int w);

/* out[m, n] = sum_k a[m, k] * b[n, k] with BOTH operands bf16 K-major activations produced earlier in the stream (b200_conv_gemm's
 * kernel with the weight-operand producer made to wait for the previous grid): S = Q K^T and O = P V of the VAE decoder's
 * single-head, head_dim-512 attention (AutoencoderKL.decode, train_audioldm_lora.py:370 / app.py:14).  b must be readable for
 * b_rows rows (multiple of block_n, >= n_valid); out bf16 or fp32, leading dim out_ld. */
int b200_gemm_nt(const void* a, int m, int k, const void* b, int b_rows, int n_valid, void* out, int out_ld, int out_fp32,
                 int block_n, int cta_pair, void* stream);

/* 3x3 stride-2 convolution with asymmetric padding -- diffusers Downsample2D(padding=0) of the VAE ENCODER
 * (`vae.encode(batch["log_mel_spec"])`, /root/reference/script/train/train_audioldm_lora.py:495): F.pad(x, (0, 1, 0, 1)) then
 * Conv2d(k 3, s 2, p 0), i.e. out[ho, wo] = sum_{kh, kw} x[2 ho + kh, 2 wo + kw] W[kh, kw] with zeros past the bottom / right
 * edge.  x bf16 NHWC [nb, h, w, c]; wpacked as for b200_conv_gemm (tap-major); out bf16 [nb, (h-2)/2+1, (w-2)/2+1, n_valid]. */
int b200_conv3x3_s2_pad01(const void* x, int c, int nb, int h, int w, const void* wpacked, int n_pad, int n_valid,
                          const float* bias, void* out, int out_ld, int block_n, int cta_pair, void* stream);

/* 1-D convolution over time-major bf16 activations x [nb, len, c] -- the layers of the HiFi-GAN vocoder
 * (transformers SpeechT5HifiGan, loaded at train_audioldm_lora.py:371 and run at the end of AudioLDMPipeline.__call__:
 * /root/reference/app.py:14, generate_audio.py:47-52): nn.Conv1d with dilation, and -- one launch per output phase --
 * nn.ConvTranspose1d.  Same tcgen05 implicit-GEMM kernel as b200_conv_gemm with a 1-D tap walk:
 *   out[n, q, :] = act( bias + sum_{t < ntaps} x[n, q + dh0 + t * dh_step, :] . W_t^T + unact(residual[n, q, :]) ),  q < m_rows
 * rows outside [0, len) read as zero (the TMA fill is the convolution's zero padding).  wpacked bf16 [n_pad, ntaps * c],
 * tap-major; c a multiple of 64.
 *   Conv1d(k, dilation d, padding d (k - 1) / 2): dh0 = -d (k - 1) / 2, dh_step = d, m_rows = len, taps W[:, :, t].
 *   Phase phi of ConvTranspose1d(k, stride s, padding p): a = (phi + p) % s, b = (phi + p) / s, taps W[:, :, s t + a]^T,
 *   dh0 = b, dh_step = -1, m_rows = ceil((len_out - phi) / s), out = y + phi * c_out, out_ld = s * c_out,
 *   out_batch_stride = len_out * c_out (0: m_rows * out_ld).
 * act: LeakyReLU with slope act_slope in [0, 1] (1 = identity, 0 = ReLU) or, act_tanh = 1, tanh, or, act_tanh = 2, the
 * exact-erf GELU (with ntaps = 1 this is nn.Linear + activation: the layers of the CLAP text encoder, audioldm_with_lora_b200/clap.py).  HiFi-GAN's residual blocks need
 * both x and leaky_relu(x); only y = leaky_relu(x, s) is stored and the residual read recovers x = min(y, y / s):
 * res_neg_gain = 1 / s (1: the residual is stored as it is).  out bf16, or fp32 (out_fp32, contiguous). */
int b200_conv1d(const void* x, int c, int nb, int len, int ntaps, int dh0, int dh_step, int m_rows, const void* wpacked,
                int n_pad, int n_valid, const float* bias, const void* residual, int res_ld, float res_neg_gain, void* out,
                int out_ld, long out_batch_stride, int out_fp32, float act_slope, int act_tanh, int block_n, int cta_pair,
                void* stream);

/* y = leaky_relu((x0 + x1 + x2) / 3, out_slope) over n bf16 elements, where the three inputs are stored as
 * a_i = leaky_relu(x_i, in_slope): the mean over the three residual blocks of a HiFi-GAN upsampling stage and the
 * activation in front of the next layer (SpeechT5HifiGan.forward). */
int b200_lrelu_mean3(const void* a0, const void* a1, const void* a2, long n, float in_slope, float out_slope, void* y,
                     void* stream);

/* fp32 -> bf16 copy (the VAE decoder's fp32 log-mel as the vocoder's bf16 input). */
int b200_f32_to_bf16(const float* x, long n, void* y, void* stream);

/* Linear layer with the rank-r LoRA branch computed inside the kernel (peft lora.Linear, unmerged; LoRA config at
 * generate_audio.py:21-29, train_audioldm_lora.py:378-385):  out = x W^T + (x A^T)(s B)^T (+ bias + residual).
 * Phase 0 of every tile runs T = x A^T as a second, narrow tcgen05.mma into spare TMEM columns; the epilogue warps turn
 * it into a bf16 K-major shared-memory tile while the base k-blocks stream, and the last k-block multiplies that tile
 * with s.B -- the down-projection never leaves the SM and costs no launch of its own.  x bf16 [m, c]; wpacked bf16
 * [n_pad, c + 64] = [W | s.B in 64 padded columns]; lora_down bf16 [64, c] (lora_rows valid stacked lora_A rows, rest
 * zero); t_out (nullable) bf16 [m, 64]: copy of T for the backward pass.  block_n <= 192. */
int b200_linear_lora(const void* x, int c, int m, const void* wpacked, int n_pad, int n_valid, const float* bias,
                     const void* residual, int res_ld, void* out, int out_ld, int block_n, int max_ctas,
                     const void* lora_down, int lora_rows, void* t_out, void* stream);

/* Linear layer consuming LayerNorm(x) with the LayerNorm folded into the GEMM (BasicTransformerBlock norm1/2/3 ->
 * attn.to_q/k/v, ff.net.0.proj; diffusers via train_audioldm_lora.py:539-546), optionally with the in-kernel LoRA branch
 * and the GEGLU epilogue:   out = LN(x) W^T (+ (LN(x) A^T)(s B)^T) (+ residual).
 * The GEMM runs on the raw rows x with gamma folded into the packed weights (wpacked = [gamma o W | s.B]); the epilogue
 * finishes the normalisation per row:  LN(x) W^T = rstd (x (gamma o W)^T - mu ln_g) + bias,  ln_g[n] = sum_c gamma_c W[n,c],
 * bias[n] = sum_c beta_c W[n,c] (+ the layer's own bias); (mu, rstd) of a row come from ln_stats fp32 [m, c / 64, 2], the
 * per-chunk (sum, sum of squares) that the kernel which produced x left behind (b200_linear_stats).  With LoRA: lora_down = gamma o A stacked [64, c], ln_ga / ln_ba fp32 [64].
 * Removes the LayerNorm launch and the normalised activation's round trip through HBM. */
int b200_linear_ln(const void* x, int c, int m, const void* wpacked, int n_pad, int n_valid, const float* bias,
                   const float* ln_g, float ln_eps, const float* ln_stats, const void* residual, int res_ld, void* out,
                   int out_ld, int geglu, int block_n, int max_ctas, const void* lora_down, int lora_rows,
                   const float* ln_ga, const float* ln_ba, void* stream);

/* Producer side of that fusion: a linear layer [m, c] -> [m, n] (optional K segment a1 / c1, or the in-kernel LoRA
 * branch) whose epilogue also writes stat_out fp32 [m, n_valid / 64, 2]: per row and 64-column chunk the (sum, sum of
 * squares) of the bf16 values it stores -- fixed slots, no atomics.  b200_linear_ln reads them as ln_stats. */
int b200_linear_stats(const void* x, int c, const void* a1, int c1, int m, const void* wpacked, int n_pad, int n_valid,
                      const float* bias, const void* residual, int res_ld, void* out, int out_ld, int block_n,
                      int max_ctas, const void* lora_down, int lora_rows, float* stat_out, void* stream);

/* Launch-shape hint for the calling thread: the number of SMs the following launches should size themselves for
 * (0 = all).  Used when independent sub-batch chains run concurrently on parallel streams. */
int b200_set_sm_budget(int n);

/* GroupNorm (+ optional SiLU) over one or two NHWC sources (cat along C is never materialised).
 * One launch: a thread-block cluster per image exchanges the group statistics through DSMEM.
 * Replaces F.group_norm + F.silu in ResnetBlock2D / Transformer2DModel.norm / conv_norm_out
 * (diffusers, via train_audioldm_lora.py:539-546). */
int b200_groupnorm_silu(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                        const float* gamma, const float* beta, float eps, int silu, void* y,
                        void* stream);

/* One-pass GroupNorm (+SiLU) over cat(x0, x1): the group statistics are merged (fp64) from the partials the producers
 * left (b200_conv_gemm_gnstat: st0 [nb, slabs0, c0 / 4, 2], st1 [nb, slabs1, c1 / 4, 2]), then the tensor is streamed
 * once.  No cluster, no second read.  Same result contract as b200_groupnorm_silu. */
int b200_groupnorm_apply(const void* x0, int c0, const float* st0, int slabs0, const void* x1, int c1, const float* st1,
                         int slabs1, int nb, int hw, int groups, const float* gamma, const float* beta, float eps,
                         int silu, void* y, void* stream);

/* Row softmax p[r, :cols] = softmax(scale * s[r, :cols]) (fp32 in, bf16 out, columns cols..cols_pad-1 zeroed so that p
 * can be the K-padded A operand of the following P.V GEMM).  With two b200_conv_gemm launches (S = Q K^T, O = P V) this is
 * the single-head, head_dim-512 attention of the VAE decoder's mid block (AutoencoderKL.decode inside
 * AudioLDMPipeline.__call__: /root/reference/app.py:14; VAE loaded at train_audioldm_lora.py:370). */
int b200_softmax_rows(const float* s, int rows, int cols, int cols_pad, long ld_s, void* p, long ld_p, float scale,
                      void* stream);

/* LayerNorm over the last dim of a [m, c] bf16 matrix.  Replaces F.layer_norm in
 * BasicTransformerBlock.norm1/2/3 (diffusers, via train_audioldm_lora.py:539-546). */
int b200_layernorm(const void* x, int m, int c, const float* gamma, const float* beta, float eps, void* y,
                   void* stream);

/* Token embedding of a RoBERTa-style text encoder: y[row, :] = LayerNorm(word[ids[row]] + type0 + pos[pos_ids[row]]) as bf16
 * [m, c]; word fp32 [vocab, c], pos fp32 [npos, c], type0 fp32 [c] (token type 0), ids / pos_ids int32 [m] (clamped to the
 * tables), c % 4 == 0.  Replaces ClapTextEmbeddings.forward (transformers) under ClapTextModelWithProjection, the prompt
 * encoder of the reference: /root/reference/script/train/train_audioldm_lora.py:368-369, :513-524; inside
 * AudioLDMPipeline._encode_prompt for app.py:14 / generate_audio.py:47-52. */
int b200_embed_layernorm(const int* ids, const int* pos_ids, int m, int c, int vocab, int npos, const float* word,
                         const float* pos, const float* type0, const float* gamma, const float* beta, float eps, void* y,
                         void* stream);

/* Fused multi-head self-attention softmax(Q K^T * scale) V over the latent sequence.
 * qkv: bf16 [batch, seq, 3*heads*head_dim] = [Q | K | V] columns, head-major inside each;
 * out: bf16 [batch, seq, heads*head_dim].  variant: 0 (debug knob for descriptor conventions).
 * Replaces F.scaled_dot_product_attention inside AttnProcessor2_0.__call__ (diffusers attention
 * processor API; LoRA targets at train_audioldm_lora.py:378-385, generate_audio.py:21-29). */
int b200_attention(const void* qkv, void* out, int batch, int seq, int heads, int head_dim, float scale,
                   int variant, void* stream);

/* Timestep / class embedding (K12): emb = cat([time_embedding(sinusoid(t)), class_embedding(labels)]).
 * t_steps: device float table; step_ptr: device int (nullable => index 0); t index = *step_ptr when
 * per_sample == 0, else t_steps[b].  Writes emb fp32 [nb, 2*ted] (nullable) and silu(emb) bf16
 * [nb, 2*ted].  Weights fp32, TRANSPOSED to row-major [in, out] (w1t [tproj, ted], w2t [ted, ted],
 * wct [class_in, ted]) so the kernel's loads coalesce.
 * Replaces Timesteps + TimestepEmbedding + class_embedding in UNet2DConditionModel.forward
 * (class_labels=prompt_embeds: train_audioldm_lora.py:543). */
int b200_time_class_embed(const float* t_steps, const int* step_ptr, int per_sample, const float* labels, int nb,
                          int tproj, int ted, int class_in, const float* w1t, const float* b1, const float* w2t,
                          const float* b2, const float* wct, const float* bc, float* emb, void* silu_emb,
                          void* stream);

/* Layout helpers.  nchw fp32 [nb, c, hw] -> nhwc bf16 [nb, hw, c_pad] (only the first c channels are
 * written); nhwc fp32 [nb, hw, c] -> nchw fp32; nearest-neighbour resize of NHWC bf16
 * (src index = floor(dst * in / out), F.interpolate(mode="nearest") in Upsample2D). */
int b200_pack_nchw_to_nhwc(const float* x, int nb, int c, int hw, int c_pad, void* y, void* stream);
int b200_unpack_nhwc_to_nchw(const float* x, int nb, int c, int hw, float* y, void* stream);
int b200_upsample_nearest(const void* x, int nb, int h, int w, int c, int ho, int wo, void* y, void* stream);

/* Sampler step (K13): classifier-free-guidance combine + DDIM / PNDM(PLMS) latent update, fused.
 *   e = e_u + g (e_t - e_u)            (eps fp32 NHWC [2*nb, hw, c]: first half uncond; g <= 1: [nb,...] only cond)
 *   ehat = sum_i w[i] * {e, hist[h1], hist[h2], hist[h3]}      x' = a * x_base + b * ehat
 * with per-step rows of `table` (8 floats: a, b, w0..w3, flags, hist slots; see sampler.cu) indexed by
 * *step_ptr, which the kernel increments.  x (fp32 NHWC state) is updated in place and the next
 * CFG-duplicated UNet input is written as bf16 NHWC with c_pad channels.
 * Replaces noise_pred chunk/guidance + DDIMScheduler.step inside AudioLDMPipeline.__call__
 * (/root/reference/app.py:14, generate_audio.py:47-52; scheduler class pinned train_audioldm_lora.py:367). */
int b200_sampler_step(const float* eps, float* x, float* x_saved, float* hist, const float* table, int* step_ptr,
                      float guidance, int do_cfg, int nb, int hw, int c, int c_pad, void* xin_next, void* stream);

/* DDIMScheduler.add_noise (train_audioldm_lora.py:504): out = sa[b]*x0 + sb[b]*noise, fp32 NCHW in,
 * also writes the bf16 NHWC c_pad-channel UNet input. */
int b200_add_noise(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, int nb, int c,
                   int hw, float* out_nchw, void* stream);

/* LoRA fine-tuning tail (train_audioldm_lora.py:549-565): fused multi-tensor AdamW over a flat fp32
 * arena; and sum((pred-target)^2) partial reduction for F.mse_loss. */
int b200_adamw_flat(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, void* stream);
int b200_mse_partial(const float* pred, const float* target, long n, float* out_sum, void* stream);
/* b200_adamw_flat with the per-step scalars in device memory: hyper_dev = {lr, 1 - beta1^t, 1 - beta2^t, grad_scale}
 * (so a captured CUDA graph of the training step stays valid while the LR schedule advances). */
int b200_adamw_flat_dev(float* param, const float* grad, float* m, float* v, long n, const float* hyper_dev,
                        float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fine-tuning step: the forward forms that keep what the backward pass needs, and the
 * backward kernels.  Together they replace torch autograd under `accelerator.backward(loss)`
 * (/root/reference/script/train/train_audioldm_lora.py:557) for the frozen-base / LoRA-only setting of
 * train_audioldm_lora.py:373-385: activation gradients (dgrad) + rank-r weight gradients.  dgrad of every
 * conv / linear layer is b200_conv_gemm itself with transposed, tap-flipped packed weights.
 * ------------------------------------------------------------------------------------------------ */

/* b200_attention that also writes lse fp32 [batch, heads, seq]: log2-domain log-sum-exp of the scaled scores. */
int b200_attention_lse(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int head_dim,
                       float scale, void* stream);

/* Flash-attention backward on tcgen05: d[Q|K|V] (bf16 [batch, seq, 3C], same fused layout as qkv) from
 * qkv, o = attention output, dout = its gradient (bf16 [batch, seq, C]) and lse.  delta fp32 [batch, heads,
 * seq] is scratch (rowsum(dout * o), written here).  Two launches of one kernel (a key-block pass for dK/dV and a
 * query-block pass for dQ): every output element has one writer, no atomics.  head_dim 32/48/64/80/96.
 * Replaces the autograd backward of F.scaled_dot_product_attention. */
int b200_attention_bwd(const void* qkv, const void* o, const void* dout, const float* lse, float* delta, void* dqkv,
                       int batch, int seq, int heads, int head_dim, float scale, void* stream);

/* b200_groupnorm_silu that also writes stats fp32 [nb, groups, 2] = (mean, rstd). */
int b200_groupnorm_silu_stats(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                              const float* gamma, const float* beta, float eps, int silu, void* y, float* stats,
                              void* stream);

/* d/dx of SiLU(GroupNorm(cat(x0, x1))) (silu = 0: GroupNorm only): dx0 [nb*hw, c0], dx1 [nb*hw, c1] (nullable: the
 * skip's gradient is not needed) from dy [nb*hw, c0+c1]; dres (nullable, row stride res_ld, concatenated channel c at
 * column c) is added -- the residual / shortcut path of ResnetBlock2D and Transformer2DModel.  Frozen affine: no
 * dgamma / dbeta.  Replaces the autograd backward of F.group_norm + F.silu. */
int b200_groupnorm_silu_bwd(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                            const float* gamma, const float* beta, const float* stats, int silu, const void* dy,
                            const void* dres, int res_ld, void* dx0, void* dx1, void* stream);

/* d/dx of LayerNorm over [m, c] (+ dres, the residual-stream gradient).  Statistics are recomputed from x. */
int b200_layernorm_bwd(const void* x, const void* dy, int m, int c, const float* gamma, float eps, const void* dres,
                       void* dx, void* stream);

/* GEGLU with the pre-activation kept: h bf16 [m, 2f] = [values | gates] -> out [m, f] = v * gelu_erf(g);
 * backward: dh [m, 2f] from dout [m, f] and h.  (Inference fuses GEGLU into the GEMM epilogue instead.) */
int b200_geglu_fwd(const void* h, long m, int f, void* out, void* stream);
int b200_geglu_bwd(const void* h, const void* dout, long m, int f, void* dh, void* stream);

/* LoRA weight gradients (peft lora.Linear: y = base(x) + s * B A x): for each of n <= 8 descriptors
 *   out[c*ldc + j*ldj] += scale * sum_{row < m} U[row, c] * V[row, j]        (c < C, j < r <= 32)
 * dB [Cout, r] = s dY^T T (U = dY, V = T = x A^T, ldc = r, ldj = 1);  dA [r, Cin] = dT^T x (U = x, V = dT = s dY B,
 * ldc = 1, ldj = Cin).  `out` points into the flat fp32 LoRA-gradient arena that NCCL all-reduces (DDP, C1).
 * descs_host: HOST array of struct { const void* u; const void* v; float* out; int ldu, ldv, C, r, ldc, ldj;
 * float scale; } (u, v bf16 device pointers, leading dims in elements). */
int b200_lora_wgrad(const void* descs_host, int n, int m, void* stream);

/* Data movement of the backward walk: z[n, 2i, 2j, :] = dy[n, i, j, :] else 0 (the stride-2 Downsample2D dgrad
 * becomes a stride-1 dgrad over z [nb, h, w, c]); gradient of the nearest-neighbour resize; y += x (bf16). */
int b200_zero_insert(const void* dy, int nb, int h, int w, int c, void* z, void* stream);
int b200_upsample_nearest_bwd(const void* dy, int nb, int h, int w, int c, int ho, int wo, void* dx, void* stream);
int b200_add_bf16(void* y, const void* x, long n, void* stream);

/* F.mse_loss(pred, noise, "mean") and its gradient (train_audioldm_lora.py:549): *loss_sum += sum (pred - noise)^2;
 * deps bf16 [nb*hw, c_pad] = 2 (pred - noise) * inv_count in columns 0..7, zero elsewhere.  pred fp32 NHWC
 * [nb, hw, 8], noise fp32 NCHW [nb, 8, hw]. */
int b200_mse_grad(const float* pred_nhwc, const float* noise_nchw, int nb, int hw, int c_pad, float inv_count,
                  float* loss_sum, void* deps, void* stream);

/* After optimizer.step(): rewrite the bf16 packed LoRA operands from the flat fp32 parameter arena.  descs_dev: DEVICE
 * array of n records struct { void* dst; long long src_off; int dst_ld, src_rows, src_cols, transpose; float scale;
 * int pad; }: dst block = scale * (transpose ? src^T : src), src = flat + src_off as [src_rows, src_cols]. */
int b200_lora_refresh(const void* descs_dev, int n, const float* flat, void* stream);

#ifdef __cplusplus
}
#endif
#endif
