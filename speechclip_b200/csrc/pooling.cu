// Input-side and pooling kernels around the hot path (SURVEY.md §8 rows f3 and b): all HBM-bound, coalesced, one pass.
//   image_normalize       uint8 HWC -> fp32 CHW, (x / 255 - mean) / std   (openai CLIP _transform: ToTensor + Normalize, as driven
//                         by clip_official.py:151-164 and the datasets' image_transform)
//   pad_rows              ragged rows packed back to back -> zero-padded [B, Tmax] (collate_general's pad_sequence,
//                         collate_function.py:30-31, on the device: the host ships sum(len) samples, not B * Tmax)
//   masked_mean_fwd/bwd   MeanPoolingLayer (pooling.py:40-60): mean over the first len[b] frames
//   attentive_pool_fwd    AttentivePoolingLayer.forward (pooling.py:335-390): tanh(alignment) + mask, row / column max, two softmaxes,
//                         two weighted sums, one CTA per pair
//   tanh_softmax_dim1     AttentivePoolingLayer.cal_batch_embedding (pooling.py:262-333): softmax over the A-sequence axis
//   relu_fwd / relu_bwd   MLPLayers (projections.py:6-29)
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

__global__ void __launch_bounds__(256) image_normalize_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int H, int W,
                                                              float m0, float m1, float m2, float s0, float s1, float s2) {
  // one thread per pixel: 3 interleaved bytes in, one float into each of the 3 planes (plane writes are coalesced over x)
  const long long hw = (long long)H * W;
  const long long b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* px = img + (b * hw + i) * 3;
    float* o = out + b * 3 * hw + i;
    o[0] = ((float)px[0] * (1.f / 255.f) - m0) * s0;
    o[hw] = ((float)px[1] * (1.f / 255.f) - m1) * s1;
    o[2 * hw] = ((float)px[2] * (1.f / 255.f) - m2) * s2;
  }
}

__global__ void __launch_bounds__(256) pad_rows_kernel(const float* __restrict__ packed, const long long* __restrict__ offsets,
                                                       const long long* __restrict__ lens, long long tmax, float* __restrict__ out) {
  const long long b = blockIdx.y;
  const long long off = offsets[b], len = lens[b];
  float* o = out + b * tmax;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tmax; i += (long long)gridDim.x * blockDim.x)
    o[i] = i < len ? packed[off + i] : 0.f;
}

// x [B][T][D] -> out [B][D]; a thread per (b, d), frames walked in order (coalesced over d)
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ lens, int T, int D,
                                                              float* __restrict__ out) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  long long n = lens ? lens[b] : T;
  n = n < 0 ? 0 : (n > T ? T : n);
  const float* xb = x + (long long)b * T * D + d;
  float s = 0.f;
  for (long long t = 0; t < n; ++t) s += xb[t * D];
  out[(long long)b * D + d] = s / (float)n;   // n == 0: NaN, as torch's mean of an empty slice
}

__global__ void __launch_bounds__(256) masked_mean_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ lens, int T, int D,
                                                              float* __restrict__ dx) {
  const int b = blockIdx.y;
  long long n = lens ? lens[b] : T;
  n = n < 0 ? 0 : (n > T ? T : n);
  const float inv = n > 0 ? 1.f / (float)n : 0.f;
  const long long total = (long long)T * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / D;
    const int d = (int)(i % D);
    dx[(long long)b * total + i] = t < n ? dout[(long long)b * D + d] * inv : 0.f;
  }
}

__device__ __forceinline__ float block_max_256(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// One CTA per pair b.  align [B][TA][TB] (A^T U B before the tanh), mask additive [B][TA][TB] or NULL, A [B][dA][TA], Bm [B][dB][TB].
// shared: scoreA [TA] | scoreB [TB] | 8 floats
__global__ void __launch_bounds__(256) attentive_pool_fwd_kernel(const float* __restrict__ align, const float* __restrict__ mask,
                                                                 const float* __restrict__ A, const float* __restrict__ Bm, int TA, int TB,
                                                                 int dA, int dB, float* __restrict__ outA, float* __restrict__ outB) {
  extern __shared__ float sm[];
  float* sa = sm;        // max over the B axis per A position, then softmax
  float* sb = sa + TA;   // max over the A axis per B position, then softmax
  float* red = sb + TB;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const float* al = align + (long long)b * TA * TB;
  const float* mk = mask ? mask + (long long)b * TA * TB : nullptr;
  // row maxima: a warp per A position (coalesced over the B axis)
  for (int i = warp; i < TA; i += nwarp) {
    float m = -INFINITY;
    for (int j = lane; j < TB; j += 32) m = fmaxf(m, tanhf(al[(long long)i * TB + j]) + (mk ? mk[(long long)i * TB + j] : 0.f));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sa[i] = m;
  }
  // column maxima: a thread per B position (coalesced over threads)
  for (int j = tid; j < TB; j += blockDim.x) {
    float m = -INFINITY;
    for (int i = 0; i < TA; ++i) m = fmaxf(m, tanhf(al[(long long)i * TB + j]) + (mk ? mk[(long long)i * TB + j] : 0.f));
    sb[j] = m;
  }
  __syncthreads();
  // softmax over each score vector
  for (int which = 0; which < 2; ++which) {
    float* s = which ? sb : sa;
    const int n = which ? TB : TA;
    float m = -INFINITY;
    for (int i = tid; i < n; i += blockDim.x) m = fmaxf(m, s[i]);
    m = block_max_256(m, red);
    float sum = 0.f;
    for (int i = tid; i < n; i += blockDim.x) {
      const float e = expf(s[i] - m);
      s[i] = e;
      sum += e;
    }
    sum = block_sum_256(sum, red);
    const float inv = 1.f / sum;
    for (int i = tid; i < n; i += blockDim.x) s[i] *= inv;
    __syncthreads();
  }
  // weighted sums: a warp per feature row (rows of A / B are contiguous over the sequence axis)
  for (int d = warp; d < dA + dB; d += nwarp) {
    const bool isA = d < dA;
    const float* row = isA ? A + ((long long)b * dA + d) * TA : Bm + ((long long)b * dB + (d - dA)) * TB;
    const float* s = isA ? sa : sb;
    const int n = isA ? TA : TB;
    float acc = 0.f;
    for (int t = lane; t < n; t += 32) acc += row[t] * s[t];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (isA) outA[(long long)b * dA + d] = acc;
      else outB[(long long)b * dB + (d - dA)] = acc;
    }
  }
}

// x [B][TA][N] -> softmax over TA of tanh(x) (+ mask [B][TA], broadcast over N); a thread per (b, n) column, coalesced over n
__global__ void __launch_bounds__(256) tanh_softmax_dim1_kernel(const float* __restrict__ x, const float* __restrict__ mask, int TA, int N,
                                                                float* __restrict__ y) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* xb = x + (long long)b * TA * N + n;
  float* yb = y + (long long)b * TA * N + n;
  const float* mk = mask ? mask + (long long)b * TA : nullptr;
  float m = -INFINITY;
  for (int t = 0; t < TA; ++t) m = fmaxf(m, tanhf(xb[(long long)t * N]) + (mk ? mk[t] : 0.f));
  float sum = 0.f;
  for (int t = 0; t < TA; ++t) {
    const float e = expf(tanhf(xb[(long long)t * N]) + (mk ? mk[t] : 0.f) - m);
    yb[(long long)t * N] = e;
    sum += e;
  }
  const float inv = 1.f / sum;
  for (int t = 0; t < TA; ++t) yb[(long long)t * N] *= inv;
}

__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fmaxf(x[i], 0.f);
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

inline unsigned grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

int image_normalize(const uint8_t* img, int batch, int H, int W, const float* mean3, const float* std3, float* out, cudaStream_t st) {
  SCB_CHECK(img && out && mean3 && std3, SCB_EINVAL, "scb_image_normalize: null operand");
  SCB_CHECK(batch <= 65535, SCB_EUNSUPPORTED, "scb_image_normalize: batch exceeds the grid limit");
  if (batch == 0 || H == 0 || W == 0) return SCB_OK;
  const dim3 grid(grid_for((long long)H * W), (unsigned)batch);
  image_normalize_kernel<<<grid, 256, 0, st>>>(img, out, H, W, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2]);
  note_launch();
  SCB_LAUNCH_OK("image_normalize");
  return SCB_OK;
}

int pad_rows(const float* packed, const long long* offsets, const long long* lens, int batch, long long tmax, float* out, cudaStream_t st) {
  SCB_CHECK(packed && offsets && lens && out, SCB_EINVAL, "scb_pad_rows: null operand");
  SCB_CHECK(batch <= 65535, SCB_EUNSUPPORTED, "scb_pad_rows: batch exceeds the grid limit");
  if (batch == 0 || tmax == 0) return SCB_OK;
  pad_rows_kernel<<<dim3(grid_for(tmax), (unsigned)batch), 256, 0, st>>>(packed, offsets, lens, tmax, out);
  note_launch();
  SCB_LAUNCH_OK("pad_rows");
  return SCB_OK;
}

int masked_mean_fwd(const float* x, const long long* lens, int batch, int T, int D, float* out, cudaStream_t st) {
  SCB_CHECK(x && out, SCB_EINVAL, "scb_masked_mean_fwd: null operand");
  if (batch == 0 || D == 0) return SCB_OK;
  masked_mean_fwd_kernel<<<dim3((unsigned)((D + 255) / 256), (unsigned)batch), 256, 0, st>>>(x, lens, T, D, out);
  note_launch();
  SCB_LAUNCH_OK("masked_mean_fwd");
  return SCB_OK;
}

int masked_mean_bwd(const float* dout, const long long* lens, int batch, int T, int D, float* dx, cudaStream_t st) {
  SCB_CHECK(dout && dx, SCB_EINVAL, "scb_masked_mean_bwd: null operand");
  if (batch == 0 || D == 0 || T == 0) return SCB_OK;
  masked_mean_bwd_kernel<<<dim3(grid_for((long long)T * D), (unsigned)batch), 256, 0, st>>>(dout, lens, T, D, dx);
  note_launch();
  SCB_LAUNCH_OK("masked_mean_bwd");
  return SCB_OK;
}

int attentive_pool_fwd(const float* align, const float* mask, const float* A, const float* Bm, int batch, int TA, int TB, int dA, int dB,
                       float* outA, float* outB, cudaStream_t st) {
  SCB_CHECK(align && A && Bm && outA && outB, SCB_EINVAL, "scb_attentive_pool_fwd: null operand");
  const size_t smem = (size_t)(TA + TB + 8) * sizeof(float);
  SCB_CHECK(smem <= 48 * 1024, SCB_EUNSUPPORTED, "scb_attentive_pool_fwd: TA + TB = %d too long", TA + TB);
  if (batch == 0) return SCB_OK;
  attentive_pool_fwd_kernel<<<(unsigned)batch, 256, smem, st>>>(align, mask, A, Bm, TA, TB, dA, dB, outA, outB);
  note_launch();
  SCB_LAUNCH_OK("attentive_pool_fwd");
  return SCB_OK;
}

int tanh_softmax_dim1(const float* x, const float* mask, int batch, int TA, int N, float* y, cudaStream_t st) {
  SCB_CHECK(x && y, SCB_EINVAL, "scb_tanh_softmax_dim1: null operand");
  if (batch == 0 || N == 0) return SCB_OK;
  tanh_softmax_dim1_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)batch), 256, 0, st>>>(x, mask, TA, N, y);
  note_launch();
  SCB_LAUNCH_OK("tanh_softmax_dim1");
  return SCB_OK;
}

int relu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
  SCB_CHECK(x && y, SCB_EINVAL, "scb_relu_fwd: null operand");
  if (n == 0) return SCB_OK;
  relu_fwd_kernel<<<grid_for(n), 256, 0, st>>>(x, y, n);
  note_launch();
  SCB_LAUNCH_OK("relu_fwd");
  return SCB_OK;
}

int relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st) {
  SCB_CHECK(dy && y && dx, SCB_EINVAL, "scb_relu_bwd: null operand");
  if (n == 0) return SCB_OK;
  relu_bwd_kernel<<<grid_for(n), 256, 0, st>>>(dy, y, dx, n);
  note_launch();
  SCB_LAUNCH_OK("relu_bwd");
  return SCB_OK;
}

}  // namespace scb
