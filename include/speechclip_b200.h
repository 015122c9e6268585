/*
 * speechclip_b200.h — C ABI of libspeechclip_b200.so (sm_100a CUDA kernels for the SpeechCLIP
 * speech–image contrastive forward / training step).
 *
 * The reference (atosystem/SpeechCLIP) is pure Python and owns no kernels or FFI: every GPU
 * operation on its hot path is a PyTorch library call.  This header is therefore
 * the boundary a maintainer would bind INSTEAD of those calls; each entry point cites the reference
 * call site (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes stub and where each call goes in the reference's modules.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - every function returns 0 on success, a negative SCB_E* code otherwise; scb_last_error()
 *     returns a thread-local message for the last failure on the calling thread;
 *   - nothing here allocates or frees caller-visible memory; scratch is passed in by the caller;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous;
 *   - 16-bit tensors are SCB_F16 (IEEE half: forward activations and weights) or SCB_BF16
 *     (gradient operands); accumulation, normalisation statistics, softmax and the loss are fp32.
 */
#ifndef SPEECHCLIP_B200_H
#define SPEECHCLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_ABI_VERSION 4

enum { SCB_OK = 0, SCB_EINVAL = -1, SCB_ECUDA = -2, SCB_EUNSUPPORTED = -3 };
enum { SCB_F32 = 0, SCB_F16 = 1, SCB_BF16 = 2 };
enum { SCB_ACT_NONE = 0, SCB_ACT_GELU_ERF = 1, SCB_ACT_QUICK_GELU = 2 };

int scb_abi_version(void);
const char* scb_last_error(void);
/* Number of kernels this library has launched on the calling process since load (bench.py's gpu_launches). */
int64_t scb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores (TMA -> smem -> UMMA -> TMEM -> fused epilogue).
 *   out[b, m, g*out_group_cols + n] = epilogue( sum_k A[b, m(+tap), k] * B[g, n, k] )
 *   epilogue(x) = act(alpha * x + bias[g*out_group_cols + n]) + residual[same index as out]
 * Replaces every nn.Linear / Conv1d-as-GEMM / patchify on the path:
 *   fairseq q/k/v/out_proj, fc1, fc2, post_extract_proj (driven from speech_encoder_plus.py:49-53,84-85),
 *   fairseq conv1..6 and pos_conv (speech_encoder_plus.py:75,35), CLIP conv1 / in_proj / out_proj /
 *   c_fc / c_proj / proj (clip_official.py:209), nn.TransformerEncoderLayer linears
 *   (kw_modules/TransformerModels.py:64-75), linear_proj (kwClip.py:1105-1106), and their dgrad/wgrad.
 *
 * A is a 16-bit (or fp32, see ab_format) tensor viewed as [batch][a_rows][a_inner] (strides in elements, a_inner contiguous).
 * K is walked in 128-byte blocks (64 16-bit elements; 32 for fp32); block kb reads A columns
 *   a_col0 + g*a_group_cols + (kb % kb_per_tap)*64 ... +63   of row   m + (kb / kb_per_tap) * tap_row_shift
 * which expresses a plain GEMM (kb_per_tap = K/64, shift 0), a strided Conv1d over channel-last
 * activations (pairs of input frames viewed as one row) and the grouped positional conv (one tap per
 * block, shift 1) without an im2col buffer.  Rows/columns outside the tensor read as zero.
 * B is [groups][n][k] 16-bit (torch Linear layout: k contiguous).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scb_gemm_args {
  const void* a;
  int64_t a_inner, a_rows, a_row_stride, a_batch_stride;
  int32_t batch, m_per_batch;
  int32_t kb_per_tap, tap_row_shift, a_col0, a_group_cols;
  const void* b;
  int64_t b_row_stride, b_group_stride;
  int32_t n, k, groups;
  void* out;            /* [batch][m_per_batch][ldc] */
  int32_t out_dtype;    /* SCB_F32 | SCB_F16 | SCB_BF16 */
  int32_t out_group_cols;
  int64_t ldc, out_batch_stride;
  void* out2;           /* optional second copy of the result (same layout), may be NULL */
  int32_t out2_dtype;
  int32_t ab_format;    /* SCB_F16 | SCB_BF16 | SCB_F32: format of A and B (SCB_F32 operands are multiplied as TF32) */
  const float* bias;    /* optional, fp32, indexed by output column */
  const void* residual; /* optional; indexed like out unless residual_ld != 0 */
  int32_t residual_dtype;
  int32_t act;
  float alpha;
  int64_t residual_ld;           /* 0: residual shares out's ldc / batch stride (then the two below are ignored) */
  int64_t residual_batch_stride; /* may be 0 with residual_ld != 0: one [m_per_batch][n] table broadcast over the batch */
  void* workspace;               /* optional: scb_gemm_workspace_bytes() of device memory, ZEROED once by the caller and then left to
                                    the library, private to the stream the call runs on.  With it, a GEMM whose last wave of tiles
                                    would leave most SMs idle splits those tiles along K over all SMs (stream-K tail); NULL: never */
  int64_t workspace_bytes;
} scb_gemm_args;

int scb_gemm(const scb_gemm_args* args, void* stream);
int64_t scb_gemm_workspace_bytes(void);

/* ------------------------------------------------------------------------------------------------
 * fp32 SIMT contraction for the few hundred rows of the trainable head (CLS row of the parallel branch and
 * its gradients): C[m,n] = alpha * sum_k A(m,k) B(n,k) + beta * C[m,n] with A(m,k) = a[m*a_rs + k*a_cs],
 * B(n,k) = b[n*b_rs + k*b_cs] (arbitrary strides: no transposed copies for dgrad / wgrad).
 * Replaces nn.TransformerEncoderLayer's out_proj / linear1 / linear2 on row 0 (TransformerModels.py:64-81),
 * linear_proj (kwClip.py:1105-1106) and their autograd backward.
 * ---------------------------------------------------------------------------------------------- */
int scb_sgemm(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t ldc,
              int32_t M, int32_t N, int32_t K, float alpha, float beta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head attention forward, softmax(scale * Q K^T + mask) V, fp32 softmax, 16-bit operands.
 * q/k/v/o: [batch][T][heads*head_dim] views with row stride *_ld and batch stride *_bs (elements); head h lives at
 * columns h*head_dim.  kv_len (nullable, int32 [batch]) = number of valid keys per batch entry (key-padding mask:
 * fairseq self_attn_padding_mask, speech_encoder_plus.py:49-53; get_keypadding_mask, data_utils.py:4-20);
 * causal != 0 adds CLIP's text mask (clip_official.py:257).  head_dim in {16,32,64,96,128}.
 * ---------------------------------------------------------------------------------------------- */
int scb_attention_fwd(const void* q, const void* k, const void* v, void* o, int32_t fmt, int64_t q_ld, int64_t k_ld, int64_t v_ld,
                      int64_t o_ld, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, const int32_t* kv_len, int32_t batch,
                      int32_t heads, int32_t head_dim, int32_t Tq, int32_t Tk, float scale, int32_t causal, void* stream);

/* Single-query attention of the parallel branch: only output row 0 ([CLS]) of the branch is consumed
 * (kwClip.py:1103) and its query is the same learned vector for every utterance, so per (utterance, head) ONE
 * query attends over all keys.  q fp32 [heads*head_dim] (unscaled; scale applied inside); kv 16-bit [batch][Tk][kv_ld]
 * with K at column k_off + h*head_dim and V at v_off + h*head_dim.  probs fp32 [batch][heads][Tk] is saved for backward.
 * bwd writes dkv (16-bit, same layout as kv; rows >= kv_len zero) and ACCUMULATES dq (fp32 [heads*head_dim]).
 * drop_p > 0: attention dropout of nn.MultiheadAttention in train mode (TransformerModels.py:64-72, dropout 0.1): the context
 * is sum_j p_j m_j v_j with m_j in {0, 1/(1-drop_p)} drawn from (rng_state, rng_site, element (b*heads+h)*Tk + j) — see
 * scb_dropout_mask; probs keeps the UNdropped p and bwd regenerates m from the same rng_state. */
int scb_cls_attention_fwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_bs, int32_t k_off, int32_t v_off,
                          const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t Tk, float scale, float* probs,
                          float* ctx32, void* ctx16, int32_t ctx16_fmt, float drop_p, const int64_t* rng_state, int32_t rng_site,
                          void* stream);
int scb_cls_attention_bwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_bs, int32_t k_off, int32_t v_off,
                          const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t Tk, float scale,
                          const float* probs, const float* dctx, void* dkv, int32_t dkv_fmt, float* dq_part, float drop_p,
                          const int64_t* rng_state, int32_t rng_site, void* stream);

/* Dropout of the trainable branch (nn.TransformerEncoderLayer / nn.MultiheadAttention in train mode, p = 0.1:
 * TransformerModels.py:55-75,110-117; spchclp_p.yaml:27).  Decisions are a pure function of (seed, step, site, element):
 * Philox-4x32-10 with key = seed and counter = (element >> 2, site, step), 32-bit lane element & 3, dropped when the draw is
 * below p * 2^32, survivors scaled by 1/(1-p).  rng_state is a device int64[2] = {seed, step}; scb_rng_advance does step += 1
 * (once per training step, graph-capturable).  The backward pass regenerates the forward's masks from a copy of rng_state.
 *   scb_dropout_mask: mask[i] = 0 or 1/(1-p) for element i of `site` (tests hand these masks to the CPU oracle);
 *   scb_dropout_rows: y[i] = x[i] * mask[i] (in place allowed) — forward on activations, backward on their gradients. */
int scb_rng_advance(int64_t* rng_state, void* stream);
int scb_dropout_mask(const int64_t* rng_state, int32_t site, float p, float* mask, int64_t n, void* stream);
int scb_dropout_rows(const float* x, float* y, int64_t n, float p, const int64_t* rng_state, int32_t site, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Waveform front end.
 * scb_frame_lengths: per-utterance integer bookkeeping done on the device (the reference does it with B host syncs
 *   and python loops: speech_encoder_plus.py:539-552,602-611; fairseq forward_padding_mask):
 *     crop_len = min(wav_len, max_audio_len) (max_audio_len <= 0: no crop), crop_off = floor(u * (wav_len - crop_len + 1))
 *     (u in [0,1): the random-crop draw of audio_transforms.py:5-23; u == NULL -> offset 0),
 *     valid_frames = min(T, ceil(crop_len / (tw_out / T)))   -- a frame is padding iff ALL its samples are padding,
 *     feat_len = min(T, round_half_even(crop_len / rate)).
 * scb_wav_prepare: out[b, i] = i < crop_len[b] ? wav[b, crop_off[b] + i] : 0 for i < tw_out; normalize != 0 applies the
 *   per-utterance F.layer_norm(wav, wav.shape) of preprocess_input (speech_encoder_plus.py:507-508; eps 1e-5) first
 *   (stats_scratch: 2*batch floats, only read when normalize != 0).
 * scb_conv0_groupnorm_gelu: fairseq ConvFeatureExtractionModel layer 0 of HuBERT-base: Conv1d(1,512,k=10,s=5) ->
 *   GroupNorm(512 groups = per channel over time, eps) -> GELU(erf); out is channel-last 16-bit [batch][n_frames][512].
 *   scratch >= scb_conv0_scratch_bytes(batch).
 * scb_conv0_layernorm_gelu: HuBERT-large variant (extractor_mode=layer_norm): conv -> LayerNorm over the 512 channels
 *   of each frame -> GELU (statistics from the Gram matrix of the weights, normalisation folded into the tensor-core operands);
 *   scratch >= scb_conv0_scratch_bytes(batch) as well (only the first 512 bytes are used).
 * ---------------------------------------------------------------------------------------------- */
int scb_frame_lengths(const int64_t* wav_len, int32_t batch, int64_t tw_out, int32_t max_audio_len, int32_t n_frames, int32_t rate,
                      const float* u, int32_t* crop_off, int32_t* crop_len, int32_t* valid_frames, int32_t* feat_len,
                      int64_t* feat_len64, void* stream);
/* out[i] = clamp(in[i] + add, 0, clamp_max): key-padding lengths of the branch (audio_len + 1 for the [CLS] slot,
 * kwClip.py:1096-1099; get_keypadding_mask, data_utils.py:4-20) as the int32 valid-key counts the attention kernels take. */
int scb_lengths_to_i32(const int64_t* in, int32_t n, int32_t add, int32_t clamp_max, int32_t* out, void* stream);
int scb_wav_prepare(const float* wav, int64_t wav_ld, int32_t batch, const int32_t* crop_off, const int32_t* crop_len, int64_t tw_out,
                    int32_t normalize, float* stats_scratch, float* out, int64_t out_ld, void* stream);
int64_t scb_conv0_scratch_bytes(int32_t batch);
int scb_conv0_groupnorm_gelu(const float* wav, int64_t wav_ld, int32_t batch, int32_t n_samples, const float* w, const float* conv_bias,
                             const float* gamma, const float* beta, float eps, void* out, int32_t out_fmt, int64_t out_batch_stride,
                             void* scratch, int64_t scratch_bytes, void* stream);
int scb_conv0_layernorm_gelu(const float* wav, int64_t wav_ld, int32_t batch, int32_t n_samples, const float* w, const float* conv_bias,
                             const float* gamma, const float* beta, float eps, void* out, int32_t out_fmt, int64_t out_batch_stride,
                             void* scratch, int64_t scratch_bytes, void* stream);
/* Zero padded frames of x in place (speech_encoder_plus.py:32-33) and write the 16-bit, group-padded (channels per group ->
 * 64), time-padded copy that the positional-conv GEMM walks tap by tap (speech_encoder_plus.py:35).  EVERY element of
 * xpad[batch][rows_pad][groups * 64] is written (zero rows before pad_left and after pad_left + T, zero channels cpg..63 of each
 * group), so the caller's buffer needs no initialisation. */
int scb_posconv_pack(float* x, const int32_t* valid_frames, void* xpad, int32_t fmt, int32_t batch, int32_t T, int32_t D, int32_t groups,
                     int32_t pad_left, int32_t rows_pad, void* stream);
/* CLIP visual.conv1 input (kernel = stride = P, clip_official.py:209): patches[b*G*G + gy*G + gx][c*P*P + py*P + px]. */
int scb_patchify(const float* img, void* out, int32_t fmt, int32_t batch, int32_t C, int32_t H, int32_t W, int32_t P, int32_t ldk,
                 void* stream);
/* out[b, :] = a[:] + a2[:] for b < nb (class_embedding + positional_embedding[0]; [CLS] prepend of kwClip.py:1093-1094). */
int scb_broadcast_row(const float* a, const float* a2, void* out, int32_t out_dtype, int64_t out_stride, int32_t nb, int32_t d,
                      void* stream);
int scb_cast_rows(const void* in, int32_t in_dtype, int64_t in_ld, void* out, int32_t out_dtype, int64_t out_ld, int64_t rows, int32_t cols,
                  void* stream);
int scb_transpose(const void* in, int32_t in_dtype, int64_t in_ld, void* out, int32_t out_dtype, int64_t out_ld, int32_t rows,
                  int32_t cols, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row kernels (HBM bound).
 * scb_layernorm_fwd: torch / fairseq / CLIP LayerNorm over the last dim (fp32 statistics), optional fused activation
 *   on the result (HuBERT-large conv blocks: LN -> GELU); optional fp32 and 16-bit outputs; stats = [rows][2] (mean, rstd).
 * scb_layernorm_bwd: dx, and dgamma/dbeta ACCUMULATED (+=).
 * scb_l2norm_*: x / ||x|| (kwClip.py:1436,1451-1453).
 * scb_weighted_sum_*: softmax(w)-weighted sum of the L hidden states (weighted_sum.py:26-45; normalize != 0 applies the
 *   parameter-free LayerNorm of :41-42 first).  h = [L] slabs of [rows][d] (h_dtype: SCB_F32, or SCB_F16 for the
 *   16-bit hidden states of the post-LN tower) at layer_stride (elements).  The 16-bit output can be
 *   scattered into the branch source buffer: row r of utterance b goes to out16 + b*out16_batch_stride + (out16_row0 + r)*d.
 *   bwd reads dout the same way and ACCUMULATES grad_scale * dL/dw into grad_logits.
 * scb_rows_bias_act: y = act(x + bias + res) on [rows][d] fp32 (res row stride res_ld; 0 = one row broadcast);
 *   pre (nullable) receives x + bias + res before the activation (saved for backward).
 * scb_gelu_bwd: dx = dy * gelu'(pre).
 * scb_column_sum: out[c] (+)= sum_r in[r*ld + c]  (bias gradients; beta 0 or 1).
 * ---------------------------------------------------------------------------------------------- */
int scb_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, float* y32, void* y16, int32_t y16_fmt,
                      float* stats, int64_t rows, int32_t d, int64_t x_ld, int64_t y_ld, float eps, int32_t act, void* stream);
int scb_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                      int64_t rows, int32_t d, void* stream);
int scb_l2norm_fwd(const float* x, float* y, float* norms, int32_t rows, int32_t d, void* stream);
int scb_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int32_t rows, int32_t d, void* stream);
int scb_weighted_sum_fwd(const void* h, int32_t h_dtype, int64_t layer_stride, const float* w_logits, int32_t L, int32_t normalize, float* out32,
                         void* out16, int32_t out16_fmt, int64_t rows, int32_t d, int32_t rows_per_batch, int64_t out16_batch_stride,
                         int64_t out16_row0, void* stream);
int scb_weighted_sum_bwd(const void* h, int32_t h_dtype, int64_t layer_stride, const float* w_logits, int32_t L, int32_t normalize, const float* dout,
                         int64_t rows, int32_t d, int32_t rows_per_batch, int64_t dout_batch_stride, int64_t dout_row0,
                         float* scratch_L, float* grad_logits, float grad_scale, void* stream);
int scb_rows_bias_act(const float* x, int64_t x_ld, const float* bias, const float* res, int64_t res_ld, int32_t act, float* pre, float* y,
                      int64_t y_ld, int64_t rows, int32_t d, void* stream);
int scb_gelu_bwd(const float* dy, const float* pre, float* dx, int64_t n, void* stream);
int scb_column_sum(const void* in, int32_t in_dtype, int64_t ld, int64_t rows, int32_t cols, float* out, float beta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked symmetric InfoNCE, forward + backward (avssl/module/losses.py:185-245; kwClip.py:1248-1297).
 *   logits = A B^T * mult (mult = exp(*log_mult) if log_mult else fixed_mult), diagonal -= margin;
 *   neg_ij = (id_i != id_j) | (i == j & !dcl)    (ids NULL: i != j, plus the diagonal unless dcl);
 *   loss = [a2b] mean_i(-l_ii + log sum_j e^{l_ij} neg_ij) + [b2a] (columns) ; halved when both.
 * No max-subtraction, exactly like the reference.  B is not capped at MAX_EYE=256 (losses.py:126).
 * phase 1 = forward (logits, row/column sums and loss; the logits stay in scratch), 2 = backward from the scratch a phase-1
 * call left (consumes it), 3 = both.  Gradients (each nullable): dA = s * dloss/dA, dB, and dlog_mult (ACCUMULATED), with
 * s = upstream * (upstream_dev ? *upstream_dev : 1) — upstream_dev is autograd's incoming dloss on the device (no host sync).
 * logits_out nullable [B][B].
 * ---------------------------------------------------------------------------------------------- */
int64_t scb_infonce_scratch_bytes(int32_t B); /* enough for any D <= 1024 */
int scb_infonce(const float* feat_a, const float* feat_b, const int64_t* ids, int32_t B, int32_t D, const float* log_mult, float fixed_mult,
                float margin, int32_t dcl, int32_t a2b, int32_t b2a, int32_t phase, float* loss, float* logits_out, float upstream,
                const float* upstream_dev, float* dA, float* dB, float* dlog_mult, void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step over ONE flat fp32 buffer: Lightning's gradient_clip_val global-norm clip (spchclp_p.yaml:108) fused with
 * torch.optim.Adam (kwClip.py:666-694: L2 weight decay folded into the gradient, bias-corrected moments), refreshing the
 * optional 16-bit weight copies.  step counts from 1; max_norm <= 0 disables clipping; sumsq_scratch = 1 double.
 * ---------------------------------------------------------------------------------------------- */
int scb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double* sumsq_scratch, float grad_scale, float max_norm, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, void* p_f16, void* p_bf16, void* stream);

/* Retrieval (avssl/module/retrieval.py:45-121 without the argsort + python row loops): for query row i,
 * best = max_j { score[i][j] : cand_ids[j] == answers[i] }, rank[i] = #{ j : score[i][j] > best } (so the answer is inside
 * the top k of a descending sort iff rank[i] < k; rank = cols when no candidate matches), top1[i] = argmax_j score[i][j]
 * (lowest index on ties).  score fp32 [rows][ld]. */
int scb_retrieval_rank(const float* score, int64_t ld, int32_t rows, int32_t cols, const int64_t* cand_ids, const int64_t* answers,
                       int32_t* rank, int32_t* top1, void* stream);

/* ---- input side and pooling (SURVEY.md 8 rows f3 / b) -------------------------------------------------------------------- */

/* openai CLIP's image transform after the resize / crop (ToTensor + Normalize; the transform the reference's datasets and
 * ClipModel.prep_image apply per image on the host, clip_official.py:151-164): uint8 [batch][H][W][3] -> fp32 [batch][3][H][W],
 * (x / 255 - mean[c]) / std[c].  The host ships 1 byte per sample instead of 4. */
int scb_image_normalize(const uint8_t* img_hwc, int32_t batch, int32_t H, int32_t W, const float* mean3, const float* std3, float* out_chw,
                        void* stream);
/* collate_general's pad_sequence (avssl/data/collate_function.py:30-31) on the device: row b is packed[offsets[b] .. + lens[b]),
 * out is [batch][tmax] zero padded (every element written). */
int scb_pad_rows(const float* packed, const int64_t* offsets, const int64_t* lens, int32_t batch, int64_t tmax, float* out, void* stream);
/* MeanPoolingLayer (avssl/module/pooling.py:40-60): out[b][d] = mean over the first lens[b] frames of x [batch][T][D] (lens NULL:
 * all T).  bwd: dx[b][t][d] = dout[b][d] / lens[b] for t < lens[b], else 0. */
int scb_masked_mean_fwd(const float* x, const int64_t* lens, int32_t batch, int32_t T, int32_t D, float* out, void* stream);
int scb_masked_mean_bwd(const float* dout, const int64_t* lens, int32_t batch, int32_t T, int32_t D, float* dx, void* stream);
/* AttentivePoolingLayer.forward (pooling.py:335-390) after the two alignment GEMMs: align [batch][TA][TB] = A^T U B (before the
 * tanh), mask additive [batch][TA][TB] or NULL, A [batch][dA][TA], B [batch][dB][TB]:
 *   s = tanh(align) + mask; scoreA = softmax_TA(max_TB s); scoreB = softmax_TB(max_TA s); outA = A scoreA; outB = B scoreB. */
int scb_attentive_pool_fwd(const float* align, const float* mask, const float* A, const float* B, int32_t batch, int32_t TA, int32_t TB,
                           int32_t dA, int32_t dB, float* outA, float* outB, void* stream);
/* AttentivePoolingLayer.cal_batch_embedding (pooling.py:262-333): y = softmax over TA of tanh(x [batch][TA][N]) + mask [batch][TA]. */
int scb_tanh_softmax_dim1(const float* x, const float* mask, int32_t batch, int32_t TA, int32_t N, float* y, void* stream);
/* MLPLayers (avssl/module/projections.py:6-29): ReLU forward / backward on fp32 rows. */
int scb_relu_fwd(const float* x, float* y, int64_t n, void* stream);
int scb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);

/* ---- cascaded branch (avssl/model/kwClip.py:857-916 KW_CascadedBranch.extract_hidden_states / forward) ------------------- */

/* MultiheadAttentionAndNorm (avssl/module/kw_modules/TransformerModels.py:11-60) with the keyword [CLS] vectors as queries
 * (kwClip.py:866-872): nq <= 8 learned queries shared by every utterance attend over that utterance's frames.
 * q fp32 [nq][heads*head_dim] (already projected, unscaled); kv 16-bit [batch][Tk][kv_ld], K of head h at k_off + h*head_dim,
 * V at v_off + h*head_dim; kv_len[b] = valid rows (NULL = all) -- the key_padding_mask of kwClip.py:870.
 * probs fp32 [batch][heads][nq][Tk] (saved for the backward; zero beyond kv_len), ctx fp32 [batch][nq][heads*head_dim].
 * drop_p > 0: attention dropout (train mode, TransformerModels.py:110-117) on element ((b*heads+h)*nq + k)*Tk + j of rng_site,
 * as in scb_cls_attention_fwd. */
int scb_mq_attention_fwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_batch_stride, int32_t k_off, int32_t v_off,
                         const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t nq, int32_t Tk, float scale,
                         float* probs, float* ctx, float drop_p, const int64_t* rng_state, int32_t rng_site, void* stream);
/* Backward: dkv (16-bit, layout of kv; rows >= kv_len zeroed; only the K and V column ranges are written), dq_part fp32
 * [batch][nq][heads*head_dim] = every utterance's contribution to dq (sum over the batch with scb_column_sum: no atomics,
 * bit-reproducible). */
int scb_mq_attention_bwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_batch_stride, int32_t k_off, int32_t v_off,
                         const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t nq, int32_t Tk, float scale,
                         const float* probs, const float* dctx, void* dkv, int32_t dkv_fmt, float* dq_part, float drop_p,
                         const int64_t* rng_state, int32_t rng_site, void* stream);

/* Kw_BatchNorm, eachKw + parallel (avssl/module/speechclip_c_modules/kw_bn.py:96-125): x fp32 [batch][n_kw][d] is viewed as
 * BatchNorm1d over d*n_kw features with feature index f = dim * n_kw + kw (the permute/reshape of kw_bn.py:116-118).
 * training != 0: batch statistics (biased variance), running stats updated with `momentum` (unbiased variance), save_mean /
 * save_rstd [n_kw*d] kept for the backward; training == 0: running statistics. gamma/beta/running_* are indexed by f. */
int scb_batchnorm_fwd(const float* x, float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* save_mean, float* save_rstd, int32_t batch, int32_t n_kw, int32_t d, float eps, float momentum,
                      int32_t training, void* stream);
int scb_batchnorm_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                      float* dgamma, float* dbeta, int32_t batch, int32_t n_kw, int32_t d, void* stream);

/* Cosine similarity against the vocabulary + SimpleVectorQuantizer (kwClip.py:884-896; my_vector_quantizer.py:66-125).
 * dots fp32 [rows][ld] holds kw @ E^T on entry and the masked cosine scores on exit (cos = dot / max(|kw||E_v|, 1e-8); the ids
 * in mask_ids -> -inf, kwClip.py:886-888 / my_vector_quantizer.py:77-80); idx[r] = argmax_v (lowest index on ties);
 * stats fp32 [rows][4] = {max cos, sum exp((cos-max)/temp), sum exp(cos-max), |kw_r|}.  The forward VALUE of
 * subword_prob is the one-hot of idx (hard straight-through), so keywords = E[idx] (scb_keyword_embed).
 * kw == emb_norm == NULL: the rows already hold final scores (SimpleVectorQuantizer used on its own); stats[..][3] = 0. */
int scb_vq_forward(float* dots, const float* kw, const float* emb_norm, int32_t rows, int32_t vocab, int32_t d, int64_t ld,
                   const int32_t* mask_ids, int32_t n_mask, float temp, int64_t* idx, float* stats, void* stream);
/* g fp32 [rows][ld] = d loss / d subword_prob (= dkeywords @ E^T) on entry, d loss / d cos on exit (softmax(cos/temp) Jacobian,
 * the straight-through path of my_vector_quantizer.py:104-110); t2[r] = <dcos_r, cos_r>. */
int scb_vq_backward(float* g, const float* cos, int32_t rows, int32_t vocab, int64_t ld, const float* stats, float temp, float* t2,
                    void* stream);
/* Cosine-similarity backward wrt the keyword rows: dkw = t1 / |kw| - t2 * kw / |kw|^2 with t1 = dcos @ (E / |E|). */
int scb_cosine_bwd_rows(const float* t1, const float* t2, const float* kw, const float* stats, float* dkw, int32_t rows, int32_t d,
                        void* stream);
/* Logging statistics of my_vector_quantizer.py:84-118: hist[v] += [idx == v], avg[v] += softmax(cos)[v] (both ACCUMULATE over
 * rows; zero first), ent[r] = -sum_v p log(p + 1e-9). */
int scb_vq_diagnostics(const float* cos, int32_t rows, int32_t vocab, int64_t ld, const float* stats, const int64_t* idx, float* hist,
                       float* avg, float* ent, void* stream);
/* Text-tower input of ClipModel.encode_keywords (avssl/module/clip_official.py:220-268): x0 fp32 [batch][n_kw+2][d] =
 * {E[sot], E[idx[b][0..n_kw)], E[eot]} + positional rows 0..n_kw+1; keywords fp32 [batch][n_kw][d] = E[idx] (may be NULL). */
int scb_keyword_embed(const float* emb, const float* pos, const int64_t* idx, int64_t sot, int64_t eot, int32_t batch, int32_t n_kw, int32_t d,
                      float* x0, float* keywords, void* stream);
/* ClipModel.encode_text input (avssl/module/clip_official.py:211-218 -> openai CLIP.encode_text): x fp32 [batch][L][d] = E[tokens] + positional rows. */
int scb_token_embed(const float* emb, const float* pos, const int64_t* tokens, int32_t batch, int32_t L, int32_t d, int64_t vocab, float* x,
                    void* stream);
/* out[b] = src[b][row[b]] (the `x[arange(B), text.argmax(-1)]` pick of the EOT position in openai CLIP.encode_text). */
int scb_gather_rows(const float* src, const int64_t* row, int32_t batch, int32_t L, int32_t d, float* out, void* stream);
/* dQ/dK/dV of softmax attention over short sequences (the n_kw+2 live positions of the causal CLIP text tower):
 * qkv 16-bit [batch][L][3*heads*head_dim], dctx fp32 [batch][L][heads*head_dim], dqkv fp32 (layout of qkv). L <= 128. */
int scb_attention_small_bwd(const void* qkv, int32_t fmt, const float* dctx, float* dqkv, int32_t batch, int32_t L, int32_t heads,
                            int32_t head_dim, float scale, int32_t causal, void* stream);
/* Row softmax for attention with one wide head (MultiheadAttentionAndNorm has nhead = 1, head_dim = d_model, which the
 * flash kernels do not cover): out[r][c] = softmax(s[r][0..len))[c] for c < len[r / rows_per_batch] (len NULL = cols), zero for
 * len <= c < out_cols.  s fp32 [rows][ld] (already scaled), out 16-bit [rows][out_ld]. */
int scb_softmax_rows(const float* s, int64_t ld, int64_t rows, int32_t rows_per_batch, const int32_t* len, int32_t cols, void* out, int32_t fmt,
                     int64_t out_ld, int32_t out_cols, void* stream);
/* Split fp32 rows [rows][cols] (row stride src_ld) into TF32-exact halves hi / lo = x - hi, written as dst [rows][3*cols]:
 * role 0 (left operand) = [hi | lo | hi], role 1 (right operand) = [hi | hi | lo].  One SCB_F32 scb_gemm over k = 3*cols of a
 * role-0 and a role-1 buffer yields the product to ~2^-21 relative -- used for the keyword-vs-vocabulary scores whose argmax
 * selects the token (kwClip.py:890-911), where plain TF32 rounding could move the index. */
int scb_split_tf32(const float* src, int64_t src_ld, float* dst, int64_t rows, int32_t cols, int32_t role, void* stream);
/* out = act(pre) on 16-bit rows (n even), and dx = dy * act'(pre) with a 16-bit pre-activation. */
int scb_act16_fwd(const void* pre, int32_t fmt, int32_t act, void* out, int64_t n, void* stream);
int scb_act_bwd(const float* dy, const void* pre, int32_t fmt, int32_t act, float* dx, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHCLIP_B200_H */
