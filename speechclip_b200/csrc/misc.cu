// Small HBM / latency-bound helpers around the towers:
//   frame_lengths   per-utterance crop / frame-count bookkeeping on the device (no host syncs)
//   wav_prepare     crop + zero-pad (+ optional per-utterance normalisation) of the waveform batch
//   rows_bias_act   y = act(x + bias + residual) on a few hundred fp32 rows (trainable head, CLS row only)
//   gelu_bwd        dx = dy * gelu'(pre)
//   column_sum      bias gradients
//   retrieval_rank  rank of the best matching candidate per query row + top-1 index (recall@k)
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

__global__ void frame_lengths_kernel(const long long* __restrict__ wav_len, int batch, long long tw_out, int max_audio_len, int n_frames,
                                     int rate, const float* __restrict__ u, int* __restrict__ crop_off, int* __restrict__ crop_len,
                                     int* __restrict__ valid_frames, int* __restrict__ feat_len, long long* __restrict__ feat_len64) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  long long len = wav_len ? wav_len[b] : tw_out;
  if (len < 0) len = 0;
  long long cl = len;
  if (max_audio_len > 0 && cl > max_audio_len) cl = max_audio_len;
  if (cl > tw_out) cl = tw_out;
  long long off = 0;
  if (u && len > cl) {
    off = (long long)(u[b] * (float)(len - cl + 1));
    if (off > len - cl) off = len - cl;
    if (off < 0) off = 0;
  }
  if (crop_off) crop_off[b] = (int)off;
  if (crop_len) crop_len[b] = (int)cl;
  if (valid_frames) {
    // fairseq forward_padding_mask: drop tw_out % T trailing samples, view [T, chunk]; frame is padding iff all samples are.
    const long long chunk = n_frames > 0 ? tw_out / n_frames : 1;
    long long vf = chunk > 0 ? (cl + chunk - 1) / chunk : n_frames;
    if (vf > n_frames) vf = n_frames;
    valid_frames[b] = (int)vf;
  }
  // python round(len / rate): round-half-to-even
  const long long q = cl / rate, r = cl % rate;
  long long fl = q;
  if (2 * r > rate || (2 * r == rate && (q & 1))) fl = q + 1;
  if (fl > n_frames) fl = n_frames;
  if (feat_len) feat_len[b] = (int)fl;
  if (feat_len64) feat_len64[b] = fl;
}

__global__ void lengths_to_i32_kernel(const long long* __restrict__ in, int n, int add, int clamp_max, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = in[i] + add;
  if (v < 0) v = 0;
  if (v > clamp_max) v = clamp_max;
  out[i] = (int)v;
}

// One block per (utterance, chunk): plain crop copy.
__global__ void __launch_bounds__(256) wav_crop_kernel(const float* __restrict__ wav, long long wav_ld, const int* __restrict__ crop_off,
                                                       const int* __restrict__ crop_len, long long tw_out, const float2* __restrict__ stats,
                                                       float* __restrict__ out, long long out_ld) {
  const int b = blockIdx.y;
  const long long off = crop_off ? crop_off[b] : 0;
  const long long len = crop_len ? crop_len[b] : tw_out;
  float mean = 0.f, rstd = 1.f;
  if (stats) {
    mean = stats[b].x;
    rstd = stats[b].y;
  }
  const float* x = wav + (long long)b * wav_ld + off;
  float* o = out + (long long)b * out_ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tw_out; i += (long long)gridDim.x * blockDim.x)
    o[i] = i < len ? (x[i] - mean) * rstd : 0.f;
}

// F.layer_norm(wav, wav.shape): mean / biased variance over the (cropped) utterance, eps 1e-5.  One block per utterance.
__global__ void __launch_bounds__(1024) wav_stats_kernel(const float* __restrict__ wav, long long wav_ld, const int* __restrict__ crop_off,
                                                         const int* __restrict__ crop_len, long long tw_out, float2* __restrict__ stats) {
  const int b = blockIdx.x;
  const long long off = crop_off ? crop_off[b] : 0;
  const long long len = crop_len ? crop_len[b] : tw_out;
  const float* x = wav + (long long)b * wav_ld + off;
  __shared__ double red[32];
  __shared__ double s_mean;
  double s = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) s += (double)x[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    s_mean = len > 0 ? t / (double)len : 0.0;
  }
  __syncthreads();
  const double mean = s_mean;
  double v = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    v += d * d;
  }
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    const double var = len > 0 ? t / (double)len : 0.0;
    stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
  }
}

__global__ void __launch_bounds__(256) rows_bias_act_kernel(const float* __restrict__ x, long long x_ld, const float* __restrict__ bias,
                                                            const float* __restrict__ res, long long res_ld, int act,
                                                            float* __restrict__ pre, float* __restrict__ y, long long y_ld, long long rows,
                                                            int d) {
  const long long total = rows * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i % d);
    float v = x[r * x_ld + c];
    if (bias) v += bias[c];
    if (res) v += res[r * res_ld + c];
    if (pre) pre[r * (long long)d + c] = v;
    if (act == SCB_ACT_GELU_ERF) v = gelu_erf(v);
    else if (act == SCB_ACT_QUICK_GELU) v = quick_gelu(v);
    y[r * y_ld + c] = v;
  }
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx,
                                                       long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = pre[i];
    // exact erf-GELU derivative with a full-precision exp (gradients feed Adam's second moment)
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    dx[i] = dy[i] * (cdf + x * pdf);
  }
}

// ---- dropout of the trainable branch (TransformerModels.py:55-75,110-117: nn.TransformerEncoderLayer / nn.MultiheadAttention, p = 0.1)
__global__ void rng_advance_kernel(long long* state) { state[1] += 1; }

__global__ void __launch_bounds__(256) dropout_mask_kernel(const long long* __restrict__ state, int site, float p, float* __restrict__ mask,
                                                           long long n) {
  const DropoutRng rng(state, site, p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) mask[i] = rng.scale(i);
}

__global__ void __launch_bounds__(256) dropout_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p,
                                                           const long long* __restrict__ state, int site) {
  const DropoutRng rng(state, site, p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = x[i] * rng.scale(i);
}

// out[c] = beta*out[c] + sum_r in[r*ld + c].  Block = 32 columns x 8 row lanes; rows strided over blockIdx.y; atomics across y.
template <typename T>
__global__ void __launch_bounds__(256) column_sum_kernel(const T* __restrict__ in, long long ld, long long rows, int cols,
                                                         float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (long long r = (long long)blockIdx.y * 8 + ry; r < rows; r += (long long)gridDim.y * 8) s += (float)in[r * ld + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(&out[c], t);
  }
}
__global__ void scale_vec_kernel(float* __restrict__ x, int n, float beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = beta == 0.f ? 0.f : x[i] * beta;
}

// One warp per query row.
__global__ void __launch_bounds__(256) retrieval_rank_kernel(const float* __restrict__ score, long long ld, int rows, int cols,
                                                             const long long* __restrict__ cand_ids, const long long* __restrict__ answers,
                                                             int* __restrict__ rank, int* __restrict__ top1) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= rows) return;
  const float* s = score + (long long)i * ld;
  const long long ans = answers ? answers[i] : 0;
  float best = -INFINITY;
  int has = 0;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = lane; j < cols; j += 32) {
    const float v = s[j];
    if (answers && cand_ids[j] == ans) {
      best = fmaxf(best, v);
      has = 1;
    }
    if (v > mx || (v == mx && j < mi)) {
      mx = v;
      mi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    has |= __shfl_xor_sync(0xffffffffu, has, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int omi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (omx > mx || (omx == mx && omi < mi)) {
      mx = omx;
      mi = omi;
    }
  }
  if (top1 && lane == 0) top1[i] = mi;
  if (rank) {
    int cnt = 0;
    if (has)
      for (int j = lane; j < cols; j += 32) cnt += s[j] > best ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) rank[i] = has ? cnt : cols;
  }
}

}  // namespace

int frame_lengths(const long long* wav_len, int batch, long long tw_out, int max_audio_len, int n_frames, int rate, const float* u,
                  int* crop_off, int* crop_len, int* valid_frames, int* feat_len, long long* feat_len64, cudaStream_t st) {
  SCB_CHECK(rate > 0 && n_frames >= 0 && tw_out >= 0, SCB_EINVAL, "scb_frame_lengths: bad sizes");
  if (batch == 0) return SCB_OK;
  frame_lengths_kernel<<<(batch + 127) / 128, 128, 0, st>>>(wav_len, batch, tw_out, max_audio_len, n_frames, rate, u, crop_off, crop_len,
                                                           valid_frames, feat_len, feat_len64);
  note_launch();
  SCB_LAUNCH_OK("frame_lengths");
  return SCB_OK;
}

int lengths_to_i32(const long long* in, int n, int add, int clamp_max, int* out, cudaStream_t st) {
  SCB_CHECK(in && out, SCB_EINVAL, "scb_lengths_to_i32: null operand");
  if (n == 0) return SCB_OK;
  lengths_to_i32_kernel<<<(n + 127) / 128, 128, 0, st>>>(in, n, add, clamp_max, out);
  note_launch();
  SCB_LAUNCH_OK("lengths_to_i32");
  return SCB_OK;
}

int wav_prepare(const float* wav, long long wav_ld, int batch, const int* crop_off, const int* crop_len, long long tw_out, int normalize,
                float* stats_scratch, float* out, long long out_ld, cudaStream_t st) {
  SCB_CHECK(wav && out, SCB_EINVAL, "scb_wav_prepare: null operand");
  SCB_CHECK(batch <= 65535, SCB_EUNSUPPORTED, "scb_wav_prepare: batch exceeds grid limits");
  if (batch == 0 || tw_out == 0) return SCB_OK;
  float2* stats = nullptr;
  if (normalize) {
    SCB_CHECK(stats_scratch, SCB_EINVAL, "scb_wav_prepare: normalize needs stats_scratch (2*batch floats)");
    stats = reinterpret_cast<float2*>(stats_scratch);
    wav_stats_kernel<<<batch, 1024, 0, st>>>(wav, wav_ld, crop_off, crop_len, tw_out, stats);
    note_launch();
    SCB_LAUNCH_OK("wav_stats");
  }
  int chunks = (int)((tw_out + 4095) / 4096);
  if (chunks > 64) chunks = 64;
  wav_crop_kernel<<<dim3(chunks, batch), 256, 0, st>>>(wav, wav_ld, crop_off, crop_len, tw_out, stats, out, out_ld);
  note_launch();
  SCB_LAUNCH_OK("wav_crop");
  return SCB_OK;
}

int rows_bias_act(const float* x, long long x_ld, const float* bias, const float* res, long long res_ld, int act, float* pre, float* y,
                  long long y_ld, long long rows, int d, cudaStream_t st) {
  SCB_CHECK(x && y, SCB_EINVAL, "scb_rows_bias_act: null operand");
  if (rows == 0 || d == 0) return SCB_OK;
  long long blocks = (rows * d + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  rows_bias_act_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, x_ld, bias, res, res_ld, act, pre, y, y_ld, rows, d);
  note_launch();
  SCB_LAUNCH_OK("rows_bias_act");
  return SCB_OK;
}

int gelu_bwd(const float* dy, const float* pre, float* dx, long long n, cudaStream_t st) {
  SCB_CHECK(dy && pre && dx, SCB_EINVAL, "scb_gelu_bwd: null operand");
  if (n == 0) return SCB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  gelu_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(dy, pre, dx, n);
  note_launch();
  SCB_LAUNCH_OK("gelu_bwd");
  return SCB_OK;
}

int rng_advance(long long* state, cudaStream_t st) {
  SCB_CHECK(state, SCB_EINVAL, "scb_rng_advance: null state");
  rng_advance_kernel<<<1, 1, 0, st>>>(state);
  note_launch();
  SCB_LAUNCH_OK("rng_advance");
  return SCB_OK;
}

int dropout_mask(const long long* state, int site, float p, float* mask, long long n, cudaStream_t st) {
  SCB_CHECK(state && mask, SCB_EINVAL, "scb_dropout_mask: null operand");
  SCB_CHECK(p >= 0.f && p < 1.f, SCB_EINVAL, "scb_dropout_mask: p = %f outside [0, 1)", p);
  if (n == 0) return SCB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  dropout_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(state, site, p, mask, n);
  note_launch();
  SCB_LAUNCH_OK("dropout_mask");
  return SCB_OK;
}

int dropout_rows(const float* x, float* y, long long n, float p, const long long* state, int site, cudaStream_t st) {
  SCB_CHECK(x && y && state, SCB_EINVAL, "scb_dropout_rows: null operand");
  SCB_CHECK(p >= 0.f && p < 1.f, SCB_EINVAL, "scb_dropout_rows: p = %f outside [0, 1)", p);
  if (n == 0) return SCB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  dropout_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, n, p, state, site);
  note_launch();
  SCB_LAUNCH_OK("dropout_rows");
  return SCB_OK;
}

int column_sum(const void* in, int in_dtype, long long ld, long long rows, int cols, float* out, float beta, cudaStream_t st) {
  SCB_CHECK(in && out, SCB_EINVAL, "scb_column_sum: null operand");
  if (cols == 0) return SCB_OK;
  if (beta != 1.f) {
    scale_vec_kernel<<<(cols + 255) / 256, 256, 0, st>>>(out, cols, beta);
    note_launch();
  }
  if (rows == 0) return SCB_OK;
  long long gy = (rows + 63) / 64;
  if (gy > 1024) gy = 1024;
  const dim3 grid((cols + 31) / 32, (unsigned)gy);
  if (in_dtype == SCB_F32) column_sum_kernel<float><<<grid, 256, 0, st>>>((const float*)in, ld, rows, cols, out);
  else if (in_dtype == SCB_F16) column_sum_kernel<__half><<<grid, 256, 0, st>>>((const __half*)in, ld, rows, cols, out);
  else column_sum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, ld, rows, cols, out);
  note_launch();
  SCB_LAUNCH_OK("column_sum");
  return SCB_OK;
}

int retrieval_rank(const float* score, long long ld, int rows, int cols, const long long* cand_ids, const long long* answers, int* rank,
                   int* top1, cudaStream_t st) {
  SCB_CHECK(score && (rank || top1), SCB_EINVAL, "scb_retrieval_rank: null operand");
  SCB_CHECK(!rank || (cand_ids && answers), SCB_EINVAL, "scb_retrieval_rank: rank needs cand_ids and answers");
  if (rows == 0) return SCB_OK;
  retrieval_rank_kernel<<<(rows + 7) / 8, 256, 0, st>>>(score, ld, rows, cols, cand_ids, rank ? answers : nullptr, rank, top1);
  note_launch();
  SCB_LAUNCH_OK("retrieval_rank");
  return SCB_OK;
}

}  // namespace scb
