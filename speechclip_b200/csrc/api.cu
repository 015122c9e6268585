// C ABI entry points + host-side plumbing (error strings, tensor-map encoding, device queries).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "ops.cuh"

namespace scb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SCB_ECUDA;
}

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle128) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    SCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    SCB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, SCB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  // The element type only matters for OOB fill / arithmetic: 16-bit formats move as raw 2-byte words, fp32/tf32 as 4-byte words.
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_CHECK(r == CUDA_SUCCESS, SCB_ECUDA,
            "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] box [%u,%u,%u] base %p", (int)r,
            rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), (unsigned long long)(rank > 2 ? d[2] : 0),
            (unsigned long long)s[0], (unsigned long long)(rank > 2 ? s[1] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, base);
  return SCB_OK;
}

}  // namespace scb

extern "C" {

#define ST static_cast<cudaStream_t>(stream)
typedef long long ll;

int scb_abi_version(void) { return SCB_ABI_VERSION; }
const char* scb_last_error(void) { return scb::g_err; }
int64_t scb_launch_count(void) { return scb::g_launches.load(); }

int scb_gemm(const scb_gemm_args* args, void* stream) {
  if (!args) {
    scb::set_error("scb_gemm: args is NULL");
    return SCB_EINVAL;
  }
  return scb::gemm(*args, ST);
}

int64_t scb_gemm_workspace_bytes(void) { return scb::gemm_workspace_bytes(); }

int scb_sgemm(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t ldc, int32_t M,
              int32_t N, int32_t K, float alpha, float beta, void* stream) {
  return scb::sgemm(a, a_rs, a_cs, b, b_rs, b_cs, c, ldc, M, N, K, alpha, beta, ST);
}

int scb_attention_fwd(const void* q, const void* k, const void* v, void* o, int32_t fmt, int64_t q_ld, int64_t k_ld, int64_t v_ld,
                      int64_t o_ld, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, const int32_t* kv_len, int32_t batch,
                      int32_t heads, int32_t head_dim, int32_t Tq, int32_t Tk, float scale, int32_t causal, void* stream) {
  return scb::attention_fwd(q, k, v, o, fmt, q_ld, k_ld, v_ld, o_ld, q_bs, k_bs, v_bs, o_bs, kv_len, batch, heads, head_dim, Tq, Tk, scale,
                            causal, ST);
}
int scb_cls_attention_fwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_bs, int32_t k_off, int32_t v_off,
                          const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t Tk, float scale, float* probs,
                          float* ctx32, void* ctx16, int32_t ctx16_fmt, float drop_p, const int64_t* rng_state, int32_t rng_site,
                          void* stream) {
  return scb::cls_attention_fwd(q, kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, batch, heads, head_dim, Tk, scale, probs, ctx32, ctx16,
                                ctx16_fmt, drop_p, (const long long*)rng_state, rng_site, ST);
}
int scb_cls_attention_bwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_bs, int32_t k_off, int32_t v_off,
                          const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t Tk, float scale,
                          const float* probs, const float* dctx, void* dkv, int32_t dkv_fmt, float* dq, float drop_p,
                          const int64_t* rng_state, int32_t rng_site, void* stream) {
  return scb::cls_attention_bwd(q, kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, batch, heads, head_dim, Tk, scale, probs, dctx, dkv,
                                dkv_fmt, dq, drop_p, (const long long*)rng_state, rng_site, ST);
}

int scb_frame_lengths(const int64_t* wav_len, int32_t batch, int64_t tw_out, int32_t max_audio_len, int32_t n_frames, int32_t rate,
                      const float* u, int32_t* crop_off, int32_t* crop_len, int32_t* valid_frames, int32_t* feat_len,
                      int64_t* feat_len64, void* stream) {
  return scb::frame_lengths(reinterpret_cast<const ll*>(wav_len), batch, tw_out, max_audio_len, n_frames, rate, u, crop_off, crop_len,
                            valid_frames, feat_len, reinterpret_cast<ll*>(feat_len64), ST);
}
int scb_lengths_to_i32(const int64_t* in, int32_t n, int32_t add, int32_t clamp_max, int32_t* out, void* stream) {
  return scb::lengths_to_i32(reinterpret_cast<const ll*>(in), n, add, clamp_max, out, ST);
}
int scb_wav_prepare(const float* wav, int64_t wav_ld, int32_t batch, const int32_t* crop_off, const int32_t* crop_len, int64_t tw_out,
                    int32_t normalize, float* stats_scratch, float* out, int64_t out_ld, void* stream) {
  return scb::wav_prepare(wav, wav_ld, batch, crop_off, crop_len, tw_out, normalize, stats_scratch, out, out_ld, ST);
}
int64_t scb_conv0_scratch_bytes(int32_t batch) { return scb::conv0_scratch_bytes(batch); }
int scb_conv0_groupnorm_gelu(const float* wav, int64_t wav_ld, int32_t batch, int32_t n_samples, const float* w, const float* conv_bias,
                             const float* gamma, const float* beta, float eps, void* out, int32_t out_fmt, int64_t out_batch_stride,
                             void* scratch, int64_t scratch_bytes, void* stream) {
  return scb::conv0_groupnorm_gelu(wav, wav_ld, batch, n_samples, w, conv_bias, gamma, beta, eps, out, out_fmt, out_batch_stride, scratch,
                                   scratch_bytes, ST);
}
int scb_conv0_layernorm_gelu(const float* wav, int64_t wav_ld, int32_t batch, int32_t n_samples, const float* w, const float* conv_bias,
                             const float* gamma, const float* beta, float eps, void* out, int32_t out_fmt, int64_t out_batch_stride,
                             void* scratch, int64_t scratch_bytes, void* stream) {
  return scb::conv0_layernorm_gelu(wav, wav_ld, batch, n_samples, w, conv_bias, gamma, beta, eps, out, out_fmt, out_batch_stride, scratch,
                                   scratch_bytes, ST);
}
int scb_posconv_pack(float* x, const int32_t* valid_frames, void* xpad, int32_t fmt, int32_t batch, int32_t T, int32_t D, int32_t groups,
                     int32_t pad_left, int32_t rows_pad, void* stream) {
  return scb::posconv_pack(x, valid_frames, xpad, fmt, batch, T, D, groups, pad_left, rows_pad, ST);
}
int scb_patchify(const float* img, void* out, int32_t fmt, int32_t batch, int32_t C, int32_t H, int32_t W, int32_t P, int32_t ldk,
                 void* stream) {
  return scb::patchify(img, out, fmt, batch, C, H, W, P, ldk, ST);
}
int scb_broadcast_row(const float* a, const float* a2, void* out, int32_t out_dtype, int64_t out_stride, int32_t nb, int32_t d,
                      void* stream) {
  return scb::broadcast_row(a, a2, out, out_dtype, out_stride, nb, d, ST);
}
int scb_cast_rows(const void* in, int32_t in_dtype, int64_t in_ld, void* out, int32_t out_dtype, int64_t out_ld, int64_t rows, int32_t cols,
                  void* stream) {
  return scb::cast_rows(in, in_dtype, in_ld, out, out_dtype, out_ld, rows, cols, ST);
}
int scb_transpose(const void* in, int32_t in_dtype, int64_t in_ld, void* out, int32_t out_dtype, int64_t out_ld, int32_t rows,
                  int32_t cols, void* stream) {
  return scb::transpose(in, in_dtype, in_ld, out, out_dtype, out_ld, rows, cols, ST);
}

int scb_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, float* y32, void* y16, int32_t y16_fmt,
                      float* stats, int64_t rows, int32_t d, int64_t x_ld, int64_t y_ld, float eps, int32_t act, void* stream) {
  return scb::layernorm_fwd(x, x_dtype, gamma, beta, y32, y16, y16_fmt, stats, rows, d, x_ld, y_ld, eps, act, ST);
}
int scb_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                      int64_t rows, int32_t d, void* stream) {
  return scb::layernorm_bwd(dy, x, stats, gamma, dx, dgamma, dbeta, rows, d, ST);
}
int scb_l2norm_fwd(const float* x, float* y, float* norms, int32_t rows, int32_t d, void* stream) {
  return scb::l2norm_fwd(x, y, norms, rows, d, ST);
}
int scb_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int32_t rows, int32_t d, void* stream) {
  return scb::l2norm_bwd(dy, y, norms, dx, rows, d, ST);
}
int scb_weighted_sum_fwd(const void* h, int32_t h_dtype, int64_t layer_stride, const float* w_logits, int32_t L, int32_t normalize, float* out32,
                         void* out16, int32_t out16_fmt, int64_t rows, int32_t d, int32_t rows_per_batch, int64_t out16_batch_stride,
                         int64_t out16_row0, void* stream) {
  return scb::weighted_sum_fwd(h, h_dtype, layer_stride, w_logits, L, normalize, out32, out16, out16_fmt, rows, d, rows_per_batch,
                               out16_batch_stride, out16_row0, ST);
}
int scb_weighted_sum_bwd(const void* h, int32_t h_dtype, int64_t layer_stride, const float* w_logits, int32_t L, int32_t normalize, const float* dout,
                         int64_t rows, int32_t d, int32_t rows_per_batch, int64_t dout_batch_stride, int64_t dout_row0,
                         float* scratch_L, float* grad_logits, float grad_scale, void* stream) {
  return scb::weighted_sum_bwd(h, h_dtype, layer_stride, w_logits, L, normalize, dout, rows, d, rows_per_batch, dout_batch_stride, dout_row0,
                               scratch_L, grad_logits, grad_scale, ST);
}
int scb_rows_bias_act(const float* x, int64_t x_ld, const float* bias, const float* res, int64_t res_ld, int32_t act, float* pre, float* y,
                      int64_t y_ld, int64_t rows, int32_t d, void* stream) {
  return scb::rows_bias_act(x, x_ld, bias, res, res_ld, act, pre, y, y_ld, rows, d, ST);
}
int scb_gelu_bwd(const float* dy, const float* pre, float* dx, int64_t n, void* stream) { return scb::gelu_bwd(dy, pre, dx, n, ST); }
int scb_image_normalize(const uint8_t* img_hwc, int32_t batch, int32_t H, int32_t W, const float* mean3, const float* std3, float* out_chw,
                        void* stream) {
  return scb::image_normalize(img_hwc, batch, H, W, mean3, std3, out_chw, ST);
}
int scb_pad_rows(const float* packed, const int64_t* offsets, const int64_t* lens, int32_t batch, int64_t tmax, float* out, void* stream) {
  return scb::pad_rows(packed, (const long long*)offsets, (const long long*)lens, batch, tmax, out, ST);
}
int scb_masked_mean_fwd(const float* x, const int64_t* lens, int32_t batch, int32_t T, int32_t D, float* out, void* stream) {
  return scb::masked_mean_fwd(x, (const long long*)lens, batch, T, D, out, ST);
}
int scb_masked_mean_bwd(const float* dout, const int64_t* lens, int32_t batch, int32_t T, int32_t D, float* dx, void* stream) {
  return scb::masked_mean_bwd(dout, (const long long*)lens, batch, T, D, dx, ST);
}
int scb_attentive_pool_fwd(const float* align, const float* mask, const float* A, const float* B, int32_t batch, int32_t TA, int32_t TB,
                           int32_t dA, int32_t dB, float* outA, float* outB, void* stream) {
  return scb::attentive_pool_fwd(align, mask, A, B, batch, TA, TB, dA, dB, outA, outB, ST);
}
int scb_tanh_softmax_dim1(const float* x, const float* mask, int32_t batch, int32_t TA, int32_t N, float* y, void* stream) {
  return scb::tanh_softmax_dim1(x, mask, batch, TA, N, y, ST);
}
int scb_relu_fwd(const float* x, float* y, int64_t n, void* stream) { return scb::relu_fwd(x, y, n, ST); }
int scb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) { return scb::relu_bwd(dy, y, dx, n, ST); }
int scb_rng_advance(int64_t* rng_state, void* stream) { return scb::rng_advance((long long*)rng_state, ST); }
int scb_dropout_mask(const int64_t* rng_state, int32_t site, float p, float* mask, int64_t n, void* stream) {
  return scb::dropout_mask((const long long*)rng_state, site, p, mask, n, ST);
}
int scb_dropout_rows(const float* x, float* y, int64_t n, float p, const int64_t* rng_state, int32_t site, void* stream) {
  return scb::dropout_rows(x, y, n, p, (const long long*)rng_state, site, ST);
}
int scb_column_sum(const void* in, int32_t in_dtype, int64_t ld, int64_t rows, int32_t cols, float* out, float beta, void* stream) {
  return scb::column_sum(in, in_dtype, ld, rows, cols, out, beta, ST);
}

int64_t scb_infonce_scratch_bytes(int32_t B) { return scb::infonce_scratch_bytes(B); }
int scb_infonce(const float* feat_a, const float* feat_b, const int64_t* ids, int32_t B, int32_t D, const float* log_mult, float fixed_mult,
                float margin, int32_t dcl, int32_t a2b, int32_t b2a, int32_t phase, float* loss, float* logits_out, float upstream,
                const float* upstream_dev, float* dA, float* dB, float* dlog_mult, void* scratch, int64_t scratch_bytes, void* stream) {
  return scb::infonce(feat_a, feat_b, reinterpret_cast<const ll*>(ids), B, D, log_mult, fixed_mult, margin, dcl, a2b, b2a, phase, loss,
                      logits_out, upstream, upstream_dev, dA, dB, dlog_mult, scratch, scratch_bytes, ST);
}

int scb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double* sumsq_scratch, float grad_scale, float max_norm, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, void* p_f16, void* p_bf16, void* stream) {
  return scb::adam_step(p, g, m, v, n, sumsq_scratch, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, step, p_f16, p_bf16, ST);
}

int scb_retrieval_rank(const float* score, int64_t ld, int32_t rows, int32_t cols, const int64_t* cand_ids, const int64_t* answers,
                       int32_t* rank, int32_t* top1, void* stream) {
  return scb::retrieval_rank(score, ld, rows, cols, reinterpret_cast<const ll*>(cand_ids), reinterpret_cast<const ll*>(answers), rank, top1,
                             ST);
}

int scb_mq_attention_fwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_batch_stride, int32_t k_off, int32_t v_off,
                         const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t nq, int32_t Tk, float scale,
                         float* probs, float* ctx, float drop_p, const int64_t* rng_state, int32_t rng_site, void* stream) {
  return scb::mq_attention_fwd(q, kv, kv_fmt, kv_ld, kv_batch_stride, k_off, v_off, kv_len, batch, heads, head_dim, nq, Tk, scale, probs, ctx,
                               drop_p, (const long long*)rng_state, rng_site, ST);
}
int scb_mq_attention_bwd(const float* q, const void* kv, int32_t kv_fmt, int64_t kv_ld, int64_t kv_batch_stride, int32_t k_off, int32_t v_off,
                         const int32_t* kv_len, int32_t batch, int32_t heads, int32_t head_dim, int32_t nq, int32_t Tk, float scale,
                         const float* probs, const float* dctx, void* dkv, int32_t dkv_fmt, float* dq, float drop_p,
                         const int64_t* rng_state, int32_t rng_site, void* stream) {
  return scb::mq_attention_bwd(q, kv, kv_fmt, kv_ld, kv_batch_stride, k_off, v_off, kv_len, batch, heads, head_dim, nq, Tk, scale, probs, dctx,
                               dkv, dkv_fmt, dq, drop_p, (const long long*)rng_state, rng_site, ST);
}
int scb_batchnorm_fwd(const float* x, float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* save_mean, float* save_rstd, int32_t batch, int32_t n_kw, int32_t d, float eps, float momentum,
                      int32_t training, void* stream) {
  return scb::batchnorm_fwd(x, y, gamma, beta, running_mean, running_var, save_mean, save_rstd, batch, n_kw, d, eps, momentum, training, ST);
}
int scb_batchnorm_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                      float* dgamma, float* dbeta, int32_t batch, int32_t n_kw, int32_t d, void* stream) {
  return scb::batchnorm_bwd(dy, x, gamma, save_mean, save_rstd, dx, dgamma, dbeta, batch, n_kw, d, ST);
}
int scb_vq_forward(float* dots, const float* kw, const float* emb_norm, int32_t rows, int32_t vocab, int32_t d, int64_t ld,
                   const int32_t* mask_ids, int32_t n_mask, float temp, int64_t* idx, float* stats, void* stream) {
  return scb::vq_forward(dots, kw, emb_norm, rows, vocab, d, ld, mask_ids, n_mask, temp, reinterpret_cast<ll*>(idx), stats, ST);
}
int scb_vq_backward(float* g, const float* cos, int32_t rows, int32_t vocab, int64_t ld, const float* stats, float temp, float* t2,
                    void* stream) {
  return scb::vq_backward(g, cos, rows, vocab, ld, stats, temp, t2, ST);
}
int scb_cosine_bwd_rows(const float* t1, const float* t2, const float* kw, const float* stats, float* dkw, int32_t rows, int32_t d,
                        void* stream) {
  return scb::cosine_bwd_rows(t1, t2, kw, stats, dkw, rows, d, ST);
}
int scb_vq_diagnostics(const float* cos, int32_t rows, int32_t vocab, int64_t ld, const float* stats, const int64_t* idx, float* hist,
                       float* avg, float* ent, void* stream) {
  return scb::vq_diagnostics(cos, rows, vocab, ld, stats, reinterpret_cast<const ll*>(idx), hist, avg, ent, ST);
}
int scb_keyword_embed(const float* emb, const float* pos, const int64_t* idx, int64_t sot, int64_t eot, int32_t batch, int32_t n_kw, int32_t d,
                      float* x0, float* keywords, void* stream) {
  return scb::keyword_embed(emb, pos, reinterpret_cast<const ll*>(idx), sot, eot, batch, n_kw, d, x0, keywords, ST);
}
int scb_attention_small_bwd(const void* qkv, int32_t fmt, const float* dctx, float* dqkv, int32_t batch, int32_t L, int32_t heads,
                            int32_t head_dim, float scale, int32_t causal, void* stream) {
  return scb::attention_small_bwd(qkv, fmt, dctx, dqkv, batch, L, heads, head_dim, scale, causal, ST);
}
int scb_token_embed(const float* emb, const float* pos, const int64_t* tokens, int32_t batch, int32_t L, int32_t d, int64_t vocab, float* x,
                    void* stream) {
  return scb::token_embed(emb, pos, reinterpret_cast<const ll*>(tokens), batch, L, d, vocab, x, ST);
}
int scb_gather_rows(const float* src, const int64_t* row, int32_t batch, int32_t L, int32_t d, float* out, void* stream) {
  return scb::gather_rows(src, reinterpret_cast<const ll*>(row), batch, L, d, out, ST);
}
int scb_softmax_rows(const float* s, int64_t ld, int64_t rows, int32_t rows_per_batch, const int32_t* len, int32_t cols, void* out, int32_t fmt,
                     int64_t out_ld, int32_t out_cols, void* stream) {
  return scb::softmax_rows(s, ld, rows, rows_per_batch, len, cols, out, fmt, out_ld, out_cols, ST);
}
int scb_split_tf32(const float* src, int64_t src_ld, float* dst, int64_t rows, int32_t cols, int32_t role, void* stream) {
  return scb::split_tf32(src, src_ld, dst, rows, cols, role, ST);
}
int scb_act16_fwd(const void* pre, int32_t fmt, int32_t act, void* out, int64_t n, void* stream) {
  return scb::act16_fwd(pre, fmt, act, out, n, ST);
}
int scb_act_bwd(const float* dy, const void* pre, int32_t fmt, int32_t act, float* dx, int64_t n, void* stream) {
  return scb::act_bwd(dy, pre, fmt, act, dx, n, ST);
}

}  // extern "C"
