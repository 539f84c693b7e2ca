/* mic_b200 — C-ABI of the B200-native captioning hot path.
 *
 * The reference (gchhablani/multilingual-image-captioning) has NO native/FFI boundary of its own: the
 * arithmetic is emitted by XLA from flax.linen modules.  Each entry point below therefore names the
 * reference site (file:line under /root/reference) or the upstream module (HF-PT twin under
 * transformers/models/...) whose computation it replaces; SURVEY.md §8(a) row ids in brackets.
 *
 * Conventions (valid for XLA-FFI custom-call handlers and CUDA-graph capture):
 *   - plain pointers + sizes, all pointers are DEVICE pointers unless noted; `stream` is a cudaStream_t
 *   - every function only ENQUEUES work on `stream`: no allocation, no synchronisation, no default stream
 *   - returns 0 on success, non-zero on error; mic_last_error() gives a thread-local message
 *   - bf16 = __nv_bfloat16 bits; activations are row-major [rows, features]; Dense kernels are Flax
 *     layout (in, out); LayerNorm/bias/stat vectors are fp32
 */
#ifndef MIC_B200_H_
#define MIC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MIC_B200_ABI_VERSION 2

#define MIC_ACT_NONE 0
#define MIC_ACT_GELU 1       /* exact erf gelu: FlaxMBartDecoderLayer / ViT MLP */
#define MIC_ACT_QUICK_GELU 2 /* x*sigmoid(1.702x): FlaxCLIPMLP */

const char* mic_last_error(void);
int mic_abi_version(void);

/* Launch behaviour of the HBM-bound (non-GEMM) entry points called afterwards on this thread (-1 = leave unchanged).
 * programmatic_dependent_launch (default 0): kernels are launched with
 *   cudaLaunchAttributeProgrammaticStreamSerialization and call griddepcontrol.wait before touching any buffer,
 *   so their prologue overlaps the previous kernel's tail.  Measured on B200 inside CUDA graphs: no gain
 *   (profiles/r01_pdl_microbench.txt), hence off; the tcgen05 GEMMs never use it.
 * gemm_b_static: reserved (ignored).
 * gemm_sm_margin (default 0): the persistent tcgen05 GEMMs launched afterwards use #SMs - margin CTAs.  One such CTA
 *   owns a whole SM (~200 KB of shared memory), so a concurrent NCCL all-reduce (lax.pmean, main.py:698) otherwise
 *   only makes progress between GEMM kernels; data-parallel training sets a margin for the backward segment that
 *   runs under the gradient all-reduce. */
int mic_launch_options(int programmatic_dependent_launch, int gemm_b_static, int gemm_sm_margin);

/* ---- dense contraction ------------------------------------------------------------------------
 * D[M,N] = act(A[M,K] * B[N,K]^T + bias) + residual        (tcgen05/TMEM, TMA-fed, bf16 in, fp32 acc)
 * a_mn_major / b_mn_major = 0: operand stored [rows, K] (K contiguous); 1: stored [K, rows].
 * Replaces every flax.linen.Dense / Conv on the path: FlaxCLIPAttention / FlaxCLIPMLP [E3,E4],
 * visual_projection modeling_clip_vision_mbart.py:53-59,90 [E6], FlaxMBartAttention / FFN [D2-D4],
 * patch conv as GEMM [E1], and all their dgrad / wgrad contractions under jax.value_and_grad
 * main.py:696 [L2].  D is bf16 (d_is_f32=0) or fp32; accumulate=1 adds into an fp32 D.
 * D2 (optional, bf16, ld = ldd) receives the pre-activation.  block_n/group_m/split_k = 0 -> auto
 * (split_k > 1 cuts K into slices combined by TMA reduce-add; fp32 outputs only).
 * act < 0 (= -MIC_ACT_*): D = (A * B^T) o act'(U) with U (bf16 [M, ldr], the saved pre-activation of the layer
 * below) passed as `residual`: the dgrad contraction of fc2 fused with the activation backward of fc1
 * (K-major operands, bf16 output, no bias / D2 / dropout).
 * drop_seed (device u32, null = off): flax.linen.Dropout(rate=drop_p) on the activation output BEFORE the
 * residual add — FlaxMBartDecoderLayer `hidden_states = residual + dropout(...)`; the mask is a counter
 * hash of (*drop_seed + drop_site, element index), regenerated identically by mic_act_bwd_colsum. */
int mic_gemm_bf16(void* stream, int a_mn_major, int b_mn_major, const void* A, long long lda, const void* B,
                  long long ldb, int M, int N, int K, void* D, long long ldd, int d_is_f32, int accumulate,
                  const float* bias, int act, void* D2, const void* residual, long long ldr, int block_n,
                  int group_m, int split_k, const unsigned int* drop_seed, unsigned int drop_site, float drop_p);

/* ---- tied lm_head fused with log-softmax / label-smoothed CE ------------------------------------
 * modeling_clip_vision_mbart.py:170-178 [H1] + loss_fn main.py:658-680 [L1]; fp32 logits never reach HBM.
 * stats: per row and per 128-column slab: max, sum exp(z-max), sum z; and z[label].
 * Partial arrays are [mic_lm_head_num_partials(V), M]. */
int mic_lm_head_num_partials(int vocab);
int mic_lm_head_ce_stats(void* stream, const void* H, long long ldh, const void* E, long long lde,
                         const float* bias, const int* labels, int M, int V, int K, float* pmax, float* psum,
                         float* psumz, float* zlabel, void* logits_out /* optional bf16 [M, ldl] */, long long ldl);
/* finalize: lse[M], per-row loss, row_w = mask/sum(mask), out[0] = loss, out[1] = sum(mask) */
int mic_ce_finalize(void* stream, const float* pmax, const float* psum, const float* psumz, const float* zlabel,
                    const int* mask, int num_partials, int M, int V, float label_smoothing, float* lse,
                    float* row_loss, float* row_w, float* out);
/* backward [L2], no recompute: rewrites the bf16 logits left by mic_lm_head_ce_stats IN PLACE as
 * dlogits = (softmax - soft_labels) * row_w (columns >= V -> 0) and writes dbias[V] = column sums
 * (gradient of final_logits_bias).  workspace: mic_ce_softmax_bwd_workspace_floats; counters as below. */
long long mic_ce_softmax_bwd_workspace_floats(int M, long long ld);
int mic_ce_softmax_bwd(void* stream, void* logits_inout, long long ld, const int* labels, const float* lse,
                       const float* row_w, float conf, float low, int M, int V, float* dbias, float* workspace,
                       unsigned int* counters);
/* backward [L2], recompute variant (no stored logits): dlogits = (softmax - soft_labels) * row_w, bf16 [M, ldd],
 * ldd % 256 == 0, pad cols = 0 */
int mic_lm_head_ce_grad(void* stream, const void* H, long long ldh, const void* E, long long lde,
                        const float* bias, const int* labels, const float* lse, const float* row_w, float conf,
                        float low, int M, int V, int K, void* dlogits, long long ldd);
/* decode-time lm_head for greedy / beam search [G4,G5]: each CTA keeps a running log-softmax partial and a
 * running top-8 for its rows across all of its vocabulary tiles; partial arrays are
 * [mic_lm_head_search_num_partials(M), M(, 8)] */
int mic_lm_head_search_num_partials(int M);
/* upper_val / upper_idx (optional, [M]): second pass for 5..8 beams - only (logit, token) pairs ranking strictly
 * after the row's pair are considered (pass the 8th best of the first pass, mic_search_merge last_val/last_idx);
 * the partial statistics of such a pass are meaningless.
 * active (optional): device flag of the search loop's while_loop condition (generation_clip_vision_utils.py:798-820);
 * 0 = the loop has ended, the kernel returns without touching its outputs.
 * gumbel_key (optional, HOST pointer to 2 uint32): `_sample` generation_clip_vision_utils.py:537-663 — Gumbel noise
 * of jax.random.categorical's threefry2x32 stream for this key is added to every logit, so candidate 0 of a row is
 * the sampled token (the log-softmax partials are meaningless then). */
int mic_lm_head_search(void* stream, const void* H, long long ldh, const void* E, long long lde,
                       const float* bias, int mask_token, int M, int V, int K, float* pmax, float* psum,
                       float* cand_val, int* cand_idx, const float* upper_val, const int* upper_idx,
                       const int* active, const unsigned int* gumbel_key);

/* packed-operand variant of mic_lm_head_search: H and E are given as K-major tile images (mic_pack_kmajor_tiles:
 * H with tile_rows = 128 - the persistent decoder step writes it directly -, E with tile_rows = 256), so that a
 * pipeline stage is two contiguous bulk copies instead of 384 TMA box rows (the TMA unit's row rate bounds the
 * unpacked kernel at ~2.5 TB/s of the 512 MB table).  Same outputs as mic_lm_head_search. */
long long mic_pack_kmajor_tiles_bytes(long long rows, int K, int tile_rows);
int mic_pack_kmajor_tiles(void* stream, const void* src, long long ld, long long rows, int K, int tile_rows, void* out);
int mic_lm_head_search_packed(void* stream, const void* h_tiles, const void* e_tiles, const float* bias,
                              int mask_token, int M, int V, int K, float* pmax, float* psum, float* cand_val,
                              int* cand_idx, const float* upper_val, const int* upper_idx, const int* active,
                              const unsigned int* gumbel_key);

/* ---- normalisation / embedding / elementwise (HBM-bound, vectorised, warp-shuffle reductions) ----
 * flax.linen.LayerNorm (fp32 statistics, var = E[x^2]-E[x]^2) as used by FlaxCLIPEncoderLayer,
 * pre_layrnorm, FlaxMBartDecoderLayer, layernorm_embedding, layer_norm [E2,E5,D5,D6]. x,y bf16 [M,d]. */
int mic_layernorm_fwd(void* stream, const void* x, const float* gamma, const float* beta, float eps, void* y,
                      float* mean, float* rstd, int M, int d);
/* decode-time fusion [G4]: x += acc + bias; y = LayerNorm(x); acc = 0.  `acc` (fp32 [M,d]) is the split-K
 * accumulator the preceding out_proj / fc2 GEMM reduce-added into (mic_gemm_bf16 with accumulate=1). */
int mic_residual_ln_fwd(void* stream, float* acc, const float* bias, void* x, const float* gamma, const float* beta,
                        float eps, void* y, int M, int d);
/* workspaces (floats) for the two reductions below; `counters`: >= 1024 uint32 that the caller zero-initialises
 * ONCE - every kernel hands them back zeroed (ticket counters of the "last CTA reduces" scheme). */
long long mic_layernorm_bwd_workspace_floats(int M, int d);
long long mic_colsum_workspace_floats(int M, int N);
/* dx = dres + LN'(dy); dgamma/dbeta written (=) (skipped when dgamma is null) */
int mic_layernorm_bwd(void* stream, const void* dy, const void* x, const float* gamma, const float* mean,
                      const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, float* workspace,
                      unsigned int* counters, int M, int d);
/* dU = dY * act'(U) (bf16, skipped for MIC_ACT_NONE) and dbias (=|+=) column sums of the result (skipped when
 * dbias is null).  Backward of Dense bias + ACT2FN [L2]. */
int mic_act_bwd_colsum(void* stream, const void* dY, long long ldy, const void* U, long long ldu, int act, void* dU,
                       long long lddu, float* dbias, int accumulate, float* workspace, unsigned int* counters, int M,
                       int N, const unsigned int* drop_seed, unsigned int drop_site, float drop_p);
/* FlaxMBartDecoder embedding [D1]: shared[id]*scale + embed_positions[pos+offset] -> emb -> layernorm_embedding.
 * pos_ids null -> position = row % pos_mod (arange(T), modeling_clip_vision_mbart.py:490-494). */
int mic_embed_ln_fwd(void* stream, const int* ids, const int* pos_ids, int pos_mod, int pos_offset,
                     const void* table, const void* pos_table, float scale, const float* gamma, const float* beta,
                     float eps, void* emb, void* y, float* mean, float* rstd, int M, int d,
                     const unsigned int* drop_seed, unsigned int drop_site, float drop_p);
/* backward of the lookup: d_table[id] += d_emb*scale (fp32 atomics onto the tied embedding gradient);
 * d_pos_rows[t] (=) sum_b d_emb[b,t] */
int mic_embed_bwd(void* stream, const int* ids, const void* d_emb, float scale, float* d_table, float* d_pos_rows,
                  int B, int T, int d, int hot_id /* id summed before the atomics (pad token), -1 = none */);
int mic_batch_sum(void* stream, const void* x, int B, int T, int d, float* out, long long out_ld);
/* FlaxCLIPVisionEmbeddings / ViT embeddings [E1,V1]: conv(stride=kernel=patch) == GEMM over patches.
 * patchify: fp32 pixels (NHWC, or NCHW per modeling_vit_bart.py:445) -> bf16 [B*g*g, p*p*3] in (kh,kw,c)
 * order; trunc_int reproduces the int32 cast of encode() modeling_clip_vision_mbart.py:330. */
int mic_patchify(void* stream, const float* pixels, void* out, int B, int image_size, int patch, int channel_first,
                 int trunc_int);
/* Input hand-off [SURVEY 8f-2]: uint8 pixels as the data pipeline holds them after Resize/CenterCrop
 * (main.py:171-172); ConvertImageDtype(float) = x/255 and Normalize(mean, std) (main.py:173-174) are applied here,
 * fused with the patch gather.  mean3 / std3: HOST pointers to 3 floats (copied into the launch). */
int mic_patchify_u8(void* stream, const unsigned char* pixels, void* out, int B, int image_size, int patch,
                    int channel_first, int trunc_int, const float* mean3, const float* std3);
/* Input hand-off, first half [SURVEY 8f-2]: Transform of main.py:165-172 / evaluation.py:33-41 on the raw uint8 CHW
 * images of torchvision.io.read_image (main.py:224-226): Resize([S], BICUBIC) (shorter edge -> S; torchvision's tensor
 * path without antialias = ATen upsample_bicubic2d(align_corners=False), clamp, round-half-even) + CenterCrop(S), a
 * whole batch of differently sized images per launch.  blob: the images back to back, uint8 [3,H_i,W_i] each;
 * desc: int64 [n,8] = {byte offset in blob, H, W, resized H', resized W', crop top, crop left, 0} (the two integer
 * formulas of torchvision are evaluated on the host: mic_b200/transforms.py); out: uint8 [n,S,S,3], or [n,3,S,S] with
 * channel_first (the flax_vit_bart input layout). */
int mic_resize_crop_u8(void* stream, const unsigned char* blob, const long long* desc, int n, int S, int channel_first,
                       unsigned char* out);
int mic_vit_embed_ln_fwd(void* stream, const void* patch_out, const float* patch_bias, const void* cls,
                         const void* pos, const float* gamma, const float* beta, float eps, int use_ln, void* emb,
                         void* y, float* mean, float* rstd, int B, int S, int d);
int mic_drop_cls_rows(void* stream, const void* d_emb, void* out, int B, int S, int d);
/* optax.adamw main.py:629-635 + TrainState.apply_gradients :701 [O1]; flat fp32 p/m/v/g, bf16 shadow.
 * The step's scalars travel BY VALUE as kernel arguments (bias_corr1 = 1/(1-b1^t), bias_corr2 = 1/(1-b2^t),
 * grad_scale = 1/world for the pmean of main.py:698): a host that runs ahead of the stream can never change
 * them under an already enqueued update. */
int mic_adamw(void* stream, float* p, float* m, float* v, const float* g, void* shadow_bf16, long long n, float lr,
              float b1, float b2, float eps, float weight_decay, float bias_corr1, float bias_corr2,
              float grad_scale);
int mic_cast_f32_to_bf16(void* stream, const float* in, void* out, long long n);

/* ---- attention ------------------------------------------------------------------------------------
 * FlaxCLIPAttention / FlaxMBartAttention / FlaxViTSelfAttention core [E3,D2,D3]: softmax((q/sqrt(hd)) k^T + mask) v;
 * Tq,Tk <= 256, head_dim 64.  Q/K/V/O are strided views [B*T, ld] with head h at
 * column h*64.  key_mask: int [B,Tk] (1 = keep) or null; causal: key j <= query i. lse: [B,H,Tq]. */
int mic_attention_fwd(void* stream, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                      long long ldv, void* O, long long ldo, float* lse, const int* key_mask, int causal, int B,
                      int H, int Tq, int Tk, int head_dim, float scale);
int mic_attention_bwd(void* stream, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                      long long ldv, const void* O, long long ldo, const void* dO, long long lddo, const float* lse,
                      const int* key_mask, int causal, void* dQ, long long lddq, void* dK, long long lddk, void* dV,
                      long long lddv, int B, int H, int Tq, int Tk, int head_dim, float scale);
/* A/B switch for the two entry points above (process-wide): 0 (default) = the row-tiled kernels (one warp per 16-row
 * tile, cp.async operands, any Tq, Tk <= 256; the backward is ONE kernel: dQ pass then dK/dV pass per warp); 1 = the
 * one-CTA-per-(batch, head) kernels where they apply (Tq, Tk <= 64; longer sequences still take the row-tiled
 * kernels); 2 = row-tiled with the backward as a dQ and a dK/dV kernel. */
int mic_attention_impl(int impl);

/* cached 1-token attention of decode() modeling_clip_vision_mbart.py:519-651 [G4].  Cache element
 * (row,pos,h,d) at ((row*cache_len+pos)*ldkv + h*64 + d).  ancestors [R,cache_len]: cache row holding
 * position j of row r's beam history (replaces the cache gather of generation_...:945-953); null ->
 * kv row = r / rows_per_kv (cross-attention over the image's visual tokens). */
int mic_decode_attention(void* stream, const void* q, long long ldq, const void* k_cache, const void* v_cache,
                         long long ldkv, const int* ancestors, int cache_len, int n_keys, int rows_per_kv, void* o,
                         long long ldo, int R, int H, int head_dim, float scale);

/* ---- persistent decoder step (cached decode) -----------------------------------------------
 * One launch runs all layers of FlaxMBartDecoderLayer (pre-LN mBART; modeling_clip_vision_mbart.py:519-651 with
 * past_key_values) for one new token per row:  self_attn_layer_norm of layer 0, then per layer  q|k|v projection
 * (k|v written straight into position `pos` of the layer's cache) -> cached self-attention through the beam
 * ancestor table -> out_proj -> residual + encoder_attn_layer_norm -> cross-attention query -> attention over the
 * image's visual K/V -> out_proj -> residual + final_layer_norm -> fc1 + activation -> fc2 -> residual + the NEXT
 * block's LayerNorm (the decoder's layer_norm after the last layer).
 * One CTA per SM stays resident; the ops are phases separated by a grid barrier, and the producer streams the
 * next phase's weights while the current one drains.  Weights are consumed from a re-packed copy (8 KB tiles that
 * are the shared-memory image of the tcgen05 operand; mic_decoder_pack_weights, to be re-run whenever the
 * parameters change).  Caller provides (and keeps alive) every buffer.
 *   before the call:  buffers.x = residual stream after layernorm_embedding (row-major bf16 [R, d])
 *   after  the call:  buffers.h_out = hidden state after the decoder layer_norm (row-major) -> lm_head input */
typedef struct {
  const float* ln_sa_g; const float* ln_sa_b;      /* self_attn_layer_norm                      */
  const void* sa_qkv_w; const float* sa_qkv_b;     /* [d, 3d] bf16 (in, out), [3d]              */
  const void* sa_o_w;   const float* sa_o_b;       /* [d, d], [d]                               */
  const float* ln_ca_g; const float* ln_ca_b;      /* encoder_attn_layer_norm                   */
  const void* ca_q_w;   const float* ca_q_b;
  const void* ca_o_w;   const float* ca_o_b;
  const float* ln_f_g;  const float* ln_f_b;       /* final_layer_norm (before the FFN)         */
  const void* fc1_w;    const float* fc1_b;        /* [d, ffn], [ffn]                           */
  const void* fc2_w;    const float* fc2_b;        /* [ffn, d], [d]                             */
  void* self_kv;                                   /* bf16, R * cache_len * 2d elements, HEAD-MAJOR:
                                                      [row][head][K plane | V plane][pos][64], 16-byte chunk c of a
                                                      (pos) row stored at c ^ (pos & 7); private to the fused step */
  const void* enc_k;    const void* enc_v;         /* visual K / V of this layer, row pitch ld_enc */
} mic_decoder_layer_t;
typedef struct {
  void* x; void* q;                                /* bf16 [R, d] row-major                     */
  void* a_tiles; void* o_tiles;                    /* bf16 [ceil(R/128)*128, d]  tile-image scratch, 1024-B aligned */
  void* g_tiles;                                   /* bf16 [ceil(R/128)*128, ffn] tile-image scratch            */
  float* acc; float* q_acc;                        /* fp32 [R, d], zero before the first step (kept zero by the kernel) */
  const int* ancestors;                            /* [R, cache_len] beam ancestor table or NULL */
  void* h_out;                                     /* bf16 [R, d] row-major result              */
  const float* ln_out_g; const float* ln_out_b;    /* decoder layer_norm                        */
  void* h_out_tiles;                               /* optional: result as a 128-row tile image instead (input of
                                                      mic_lm_head_search_packed); h_out is then not written */
  const void* cross_kv_tiles;                      /* optional: visual K|V re-packed by mic_decoder_pack_cross_kv
                                                      ([image][layer][head][16 KB]): one bulk copy per attention item */
} mic_decoder_buffers_t;
long long mic_decoder_plan_bytes(int num_layers);
long long mic_decoder_packed_bytes(int num_layers, int d_model, int ffn_dim);
/* re-pack the six projection kernels of every layer into `packed` (1024-byte aligned, mic_decoder_packed_bytes).
 * Capturable; generate() runs it once per call so that updated parameters are always picked up. */
int mic_decoder_pack_weights(void* stream, const mic_decoder_layer_t* layers, int num_layers, int d_model,
                             int ffn_dim, void* packed);
/* visual K|V [B*S, ld] (layer l: K at column l*2d, V at l*2d + d) -> per (image, layer, head) a contiguous 16 KB
 * shared-memory stage image (S <= 64 keys, zero padded).  Capturable; once per generate() call. */
long long mic_decoder_cross_kv_tiles_bytes(int B, int num_layers, int heads);
int mic_decoder_pack_cross_kv(void* stream, const void* enc_kv, long long ld, int B, int S, int num_layers, int heads,
                              int d_model, void* out);
/* builds the phase table on the host and copies it to plan_dev (128-byte aligned device buffer of
 * mic_decoder_plan_bytes).  Not capturable into a CUDA graph (host staging): call once per buffer set. */
int mic_decoder_plan_init(void* stream, void* plan_dev, const mic_decoder_layer_t* layers, int num_layers,
                          const mic_decoder_buffers_t* buffers, const void* packed, int R, int d_model, int heads,
                          int ffn_dim, int cache_len, int enc_tokens, int rows_per_image, long long ld_enc, int act,
                          float eps);
/* sync_counter: 256 uint32 (1 KB), zero before the first call (barrier state; the kernel keeps it consistent).
 * Capturable.
 * phase_times (optional, NULL = off): [1 + 11 * num_layers, #SMs] (+256 trace slots) %globaltimer stamp of each
 * CTA's arrival at the end of each phase (profiling aid: tools/profile_decoder_step.py).
 * active (optional): device flag of the search loop's while_loop condition (generation_clip_vision_utils.py:798-820,
 * written by mic_beam_cond / mic_greedy_cond); 0 = the loop has ended and the step is skipped.
 * opts: bit 0 = additionally fence generic->async on the writer side of every phase (diagnostic; the readers'
 * fence after their acquire is what orders the proxies; results identical). */
int mic_decoder_step(void* stream, const void* plan_dev, int num_layers, int R, int pos, unsigned int* sync_counter,
                     unsigned long long* phase_times, const int* active, int opts);

/* ---- fp32 verification path (fp32 storage, fp32 SIMT arithmetic; forward + loss only) ------------------------
 * For the parity bar of BASELINE configs[0] (B = 8, fp32: logits within 1e-3 relative, loss within 1e-4 of the
 * reference restatement).  Row-major fp32 everywhere; B of the GEMM is [K, N] (Flax kernel, b_is_nk = 0) or
 * [N, K] (the embedding table used as tied lm_head, b_is_nk = 1).  Small shapes only. */
int mic_f32_gemm(void* stream, const float* A, long long lda, const float* B, long long ldb, int b_is_nk, int M, int N,
                 int K, const float* bias, int act, const float* residual, long long ldr, float* D, long long ldd);
int mic_f32_layernorm(void* stream, const float* x, const float* gamma, const float* beta, float eps, float* y, int M,
                      int d);
int mic_f32_attention(void* stream, const float* Q, long long ldq, const float* K, long long ldk, const float* V,
                      long long ldv, float* O, long long ldo, const int* key_mask, int causal, int B, int H, int Tq,
                      int Tk, int head_dim, float scale);
int mic_f32_embed(void* stream, const int* ids, const float* table, float scale, const float* pos_table, int pos_offset,
                  int T, float* out, int M, int d);
int mic_f32_patchify(void* stream, const float* pixels, float* out, int B, int image_size, int patch, int channel_first,
                     int trunc_int);
int mic_f32_vit_embed(void* stream, const float* patch_out, const float* patch_bias, const float* cls, const float* pos,
                      float* out, int B, int S, int d);
int mic_f32_ce_rows(void* stream, const float* logits, long long ld, const int* labels, int M, int V,
                    float label_smoothing, float* row_loss, float* lse);

/* profiling aid: n grid barriers of the kind mic_decoder_step uses (tools/microbench_barrier.py) */
int mic_barrier_bench(void* stream, unsigned int* sync_counter, int n, int variant);

/* ---- search steps (generation_clip_vision_utils.py) ---------------------------------------------------
 * merge the lm_head_search slab partials: per row log-softmax normaliser + top-8 (log-prob, token),
 * written to row_lp / row_tok [R, ld_out] at columns col_off .. col_off+7.  lse_given = 1: second pass (ranks
 * 9..16, 5..8 beams) - the normaliser is READ from row_max_logsum (written by the first pass).  last_val / last_idx
 * (optional, [R]): raw logit and token of the 8th best = the upper bound handed to the second search pass. */
int mic_search_merge(void* stream, const float* pmax, const float* psum, const float* cand_val, const int* cand_idx,
                     int num_partials, int R, float* row_lp, int* row_tok, float* row_max_logsum, int ld_out,
                     int col_off, int lse_given, float* last_val, int* last_idx);
/* beam_search_body_fn :822-966 steps 2-8 [G5]; forced_token >= 0 applies ForcedBOS/ForcedEOS [G2] */
int mic_beam_step(void* stream, const float* row_lp, const int* row_tok, int forced_token, int B, int K, int L,
                  int V, int cur_len, int eos_token_id, int early_stopping, float length_penalty, int* running_seq,
                  float* running_scores, int* sequences, float* scores, int* finished, int* ancestors,
                  int* next_token, int* active, int cand_per_row);   /* row_lp/row_tok are [B*K, cand_per_row]: 8 (K <= 4) or 16 (K <= 8) */
/* beam_search_cond_fn :798-820 [G6] for the next iteration; writes *active */
int mic_beam_cond(void* stream, const float* running_scores, const float* scores, const int* finished, int B, int K,
                  int cur_len, int max_length, float length_penalty, int early_stopping, int* active);
/* epilogue :978-990 */
int mic_beam_finalize(void* stream, const int* sequences, const float* scores, const int* finished,
                      const int* running_seq, const float* running_scores, int B, int K, int L, int* out_seq,
                      float* out_scores);
/* greedy_search_body_fn :489-522 / cond :480-487 [G7] */
int mic_greedy_step(void* stream, const int* row_tok, int forced_token, int R, int L, int cur_len, int eos, int pad,
                    int* sequences, int* finished, int* next_token, int* active);
int mic_greedy_cond(void* stream, const int* finished, int R, int cur_len, int max_length, int* active);

#ifdef __cplusplus
}
#endif
#endif /* MIC_B200_H_ */
