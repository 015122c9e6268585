// Internal prototypes: one host launcher per op, implemented next to its kernels.
#pragma once
#include "common.cuh"

namespace scb {
// gemm_tcgen05.cu
int gemm(const scb_gemm_args& a, cudaStream_t stream);
long long gemm_workspace_bytes();
// loss.cu
int sgemm(const float* a, long long a_rs, long long a_cs, const float* b, long long b_rs, long long b_cs, float* c, long long ldc, int M, int N,
          int K, float alpha, float beta, cudaStream_t st);
long long infonce_scratch_bytes(int B);
int infonce(const float* feat_a, const float* feat_b, const long long* ids, int B, int D, const float* log_mult, float fixed_mult,
            float margin, int dcl, int a2b, int b2a, int phase, float* loss, float* logits_out, float upstream, const float* upstream_dev,
            float* dA, float* dB, float* dlog_mult, void* scratch, long long scratch_bytes, cudaStream_t st);
// attention.cu
int attention_fwd(const void* q, const void* k, const void* v, void* o, int fmt, long long q_ld, long long k_ld, long long v_ld,
                  long long o_ld, long long q_bs, long long k_bs, long long v_bs, long long o_bs, const int* kv_len, int batch, int heads,
                  int head_dim, int Tq, int Tk, float scale, int causal, cudaStream_t st);
// attention_tc.cu (tcgen05 path; SCB_EUNSUPPORTED = shape outside its envelope, use the mma.sync kernel)
int attention_fwd_tc(const void* q, const void* k, const void* v, void* o, int fmt, long long q_ld, long long k_ld, long long v_ld,
                     long long o_ld, long long q_bs, long long k_bs, long long v_bs, long long o_bs, const int* kv_len, int batch, int heads,
                     int head_dim, int Tq, int Tk, float scale, int causal, cudaStream_t st);
int cls_attention_fwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off,
                      const int* kv_len, int batch, int heads, int head_dim, int Tk, float scale, float* probs, float* ctx32, void* ctx16,
                      int ctx16_fmt, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st);
int cls_attention_bwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off,
                      const int* kv_len, int batch, int heads, int head_dim, int Tk, float scale, const float* probs, const float* dctx,
                      void* dkv, int dkv_fmt, float* dq, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st);
// frontend.cu
long long conv0_scratch_bytes(int batch);
int conv0_groupnorm_gelu(const float* wav, long long wav_ld, int batch, int n_samples, const float* w, const float* conv_bias,
                         const float* gamma, const float* beta, float eps, void* out, int out_fmt, long long out_batch_stride,
                         void* scratch, long long scratch_bytes, cudaStream_t st);
int conv0_layernorm_gelu(const float* wav, long long wav_ld, int batch, int n_samples, const float* w, const float* conv_bias,
                         const float* gamma, const float* beta, float eps, void* out, int out_fmt, long long out_batch_stride,
                         void* scratch, long long scratch_bytes, cudaStream_t st);
int posconv_pack(float* x, const int* valid_frames, void* xpad, int fmt, int batch, int T, int D, int groups, int pad_left, int rows_pad,
                 cudaStream_t st);
int patchify(const float* img, void* out, int fmt, int batch, int C, int H, int W, int P, int ldk, cudaStream_t st);
int broadcast_row(const float* a, const float* a2, void* out, int out_dtype, long long out_stride, int nb, int d, cudaStream_t st);
int cast_rows(const void* in, int in_dtype, long long in_ld, void* out, int out_dtype, long long out_ld, long long rows, int cols,
              cudaStream_t st);
int transpose(const void* in, int in_dtype, long long in_ld, void* out, int out_dtype, long long out_ld, int rows, int cols,
              cudaStream_t st);
// norm.cu
int layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float* y32, void* y16, int y16_fmt,
                  float* stats, long long rows, int d, long long x_ld, long long y_ld, float eps, int act, cudaStream_t st);
int layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                  long long rows, int d, cudaStream_t st);
int l2norm_fwd(const float* x, float* y, float* norms, int rows, int d, cudaStream_t st);
int l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int d, cudaStream_t st);
int weighted_sum_fwd(const void* h, int h_dtype, long long layer_stride, const float* w_logits, int L, int normalize, float* out32, void* out16,
                     int out16_fmt, long long rows, int d, int rows_per_batch, long long out16_batch_stride, long long out16_row0,
                     cudaStream_t st);
int weighted_sum_bwd(const void* h, int h_dtype, long long layer_stride, const float* w_logits, int L, int normalize, const float* dout,
                     long long rows, int d, int rows_per_batch, long long dout_batch_stride, long long dout_row0, float* scratch_L,
                     float* grad_logits, float grad_scale, cudaStream_t st);
// optim.cu
int adam_step(float* p, const float* g, float* m, float* v, long long n, double* sumsq_scratch, float grad_scale, float max_norm, float lr,
              float beta1, float beta2, float eps, float weight_decay, int step, void* p_f16, void* p_bf16, cudaStream_t st);
// misc.cu
int frame_lengths(const long long* wav_len, int batch, long long tw_out, int max_audio_len, int n_frames, int rate, const float* u,
                  int* crop_off, int* crop_len, int* valid_frames, int* feat_len, long long* feat_len64, cudaStream_t st);
int lengths_to_i32(const long long* in, int n, int add, int clamp_max, int* out, cudaStream_t st);
int wav_prepare(const float* wav, long long wav_ld, int batch, const int* crop_off, const int* crop_len, long long tw_out, int normalize,
                float* stats_scratch, float* out, long long out_ld, cudaStream_t st);
int rows_bias_act(const float* x, long long x_ld, const float* bias, const float* res, long long res_ld, int act, float* pre, float* y,
                  long long y_ld, long long rows, int d, cudaStream_t st);
int gelu_bwd(const float* dy, const float* pre, float* dx, long long n, cudaStream_t st);
int rng_advance(long long* state, cudaStream_t st);
int dropout_mask(const long long* state, int site, float p, float* mask, long long n, cudaStream_t st);
int dropout_rows(const float* x, float* y, long long n, float p, const long long* state, int site, cudaStream_t st);
int column_sum(const void* in, int in_dtype, long long ld, long long rows, int cols, float* out, float beta, cudaStream_t st);
int retrieval_rank(const float* score, long long ld, int rows, int cols, const long long* cand_ids, const long long* answers, int* rank,
                   int* top1, cudaStream_t st);
// cascaded.cu
int mq_attention_fwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off, const int* kv_len,
                     int batch, int heads, int head_dim, int nq, int Tk, float scale, float* probs, float* ctx, float drop_p,
                     const long long* rng_state, int rng_site, cudaStream_t st);
int mq_attention_bwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off, const int* kv_len,
                     int batch, int heads, int head_dim, int nq, int Tk, float scale, const float* probs, const float* dctx, void* dkv,
                     int dkv_fmt, float* dq, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st);
int batchnorm_fwd(const float* x, float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  float* save_mean, float* save_rstd, int B, int NK, int D, float eps, float momentum, int training, cudaStream_t st);
int batchnorm_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                  float* dgamma, float* dbeta, int B, int NK, int D, cudaStream_t st);
int vq_forward(float* dots, const float* kw, const float* emb_norm, int R, int V, int D, long long ld, const int* mask_ids, int n_mask,
               float temp, long long* idx, float* stats, cudaStream_t st);
int vq_backward(float* g, const float* cos, int R, int V, long long ld, const float* stats, float temp, float* t2, cudaStream_t st);
int cosine_bwd_rows(const float* t1, const float* t2, const float* kw, const float* stats, float* dkw, int R, int D, cudaStream_t st);
int vq_diagnostics(const float* cos, int R, int V, long long ld, const float* stats, const long long* idx, float* hist, float* avg, float* ent,
                   cudaStream_t st);
int keyword_embed(const float* emb, const float* pos, const long long* idx, long long sot, long long eot, int B, int K, int D, float* x0,
                  float* keywords, cudaStream_t st);
int attention_small_bwd(const void* qkv, int fmt, const float* dctx, float* dqkv, int batch, int L, int heads, int head_dim, float scale,
                        int causal, cudaStream_t st);
int token_embed(const float* emb, const float* pos, const long long* tokens, int B, int L, int D, long long vocab, float* x, cudaStream_t st);
int gather_rows(const float* src, const long long* row, int B, int L, int D, float* out, cudaStream_t st);
int softmax_rows(const float* s, long long ld, long long rows, int rows_per_batch, const int* len, int cols, void* out, int fmt,
                 long long out_ld, int out_cols, cudaStream_t st);
int split_tf32(const float* src, long long src_ld, float* dst, long long rows, int cols, int role, cudaStream_t st);
int act16_fwd(const void* pre, int fmt, int act, void* out, long long n, cudaStream_t st);
int act_bwd(const float* dy, const void* pre, int fmt, int act, float* dx, long long n, cudaStream_t st);
// pooling.cu
int image_normalize(const uint8_t* img, int batch, int H, int W, const float* mean3, const float* std3, float* out, cudaStream_t st);
int pad_rows(const float* packed, const long long* offsets, const long long* lens, int batch, long long tmax, float* out, cudaStream_t st);
int masked_mean_fwd(const float* x, const long long* lens, int batch, int T, int D, float* out, cudaStream_t st);
int masked_mean_bwd(const float* dout, const long long* lens, int batch, int T, int D, float* dx, cudaStream_t st);
int attentive_pool_fwd(const float* align, const float* mask, const float* A, const float* Bm, int batch, int TA, int TB, int dA, int dB,
                       float* outA, float* outB, cudaStream_t st);
int tanh_softmax_dim1(const float* x, const float* mask, int batch, int TA, int N, float* y, cudaStream_t st);
int relu_fwd(const float* x, float* y, long long n, cudaStream_t st);
int relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st);
}  // namespace scb
