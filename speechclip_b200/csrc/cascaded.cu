// Kernels of the cascaded branch (KW_CascadedBranch, avssl/model/kwClip.py:857-916):
//   mq_attention_fwd/bwd   NQ learned queries (the keyword [CLS] vectors, identical for every utterance) attending over the
//                          K/V of all frames; heads x head_dim up to 1 x 1024 (MultiheadAttentionAndNorm uses ONE head)
//   batchnorm_fwd/bwd      Kw_BatchNorm eachKw/parallel: BatchNorm1d over (dim, keyword) features with batch statistics
//   vq_forward/backward    cosine similarity against the (reduced) CLIP vocabulary + SimpleVectorQuantizer: mask special ids,
//                          argmax, straight-through softmax(x / temp); backward folds the softmax and cosine Jacobians
//   vq_diagnostics         code / probability perplexity and per-keyword entropy (logging only)
//   keyword_embed          text-tower input rows: [SOT], the K selected token embeddings, [EOT] (+ positional embedding)
//   attention_small_bwd    dQ, dK, dV of softmax attention for short causal sequences (the K+2 = 10 live text positions)
//   act16_fwd / act_bwd    activation on 16-bit pre-activations and its derivative
//   split_tf32             fp32 rows -> (hi, lo) TF32 halves laid out so that ONE kind::tf32 GEMM over 3x the contraction length
//                          returns the fp32-accurate product (the argmax over the vocabulary must not move)
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[w]);
  return r;
}

// ------------------------------------------------------------------------------------------------ multi-query attention
// One CTA (8 warps) per (utterance, head).  q fp32 [NQ][heads*hd] (unscaled).  kv 16-bit [B][Tk][ld]: K at k_off + h*hd, V at
// v_off + h*hd.  probs fp32 [B][heads][NQ][Tk]; ctx fp32 [B][NQ][heads*hd].
// Row phases (scores = K q, dP = V dctx): a WARP per key row -- the 32 lanes read consecutive 16-byte chunks of the row (one
// coalesced 512-byte request), each accumulates its slice of all NQ dot products against the query block held in shared
// memory, then the NQ partials are reduced with shuffles.  Column phases (ctx = P V, dq = dS K): a thread per pair of
// output dims, looping over the keys, so that every key row is again read as one contiguous segment per warp.
constexpr int kMaxNQ = 8;
constexpr int kMqThreads = 256;

// out[k * out_ld + j] = scale_out * <row_j, vec_k> for j in [0, len): rows 16-bit at base + j*ld, vecs fp32 in shared memory [nq][hd]
__device__ __forceinline__ void mq_row_dots(const uint16_t* __restrict__ base, long long ld, int fmt, int len, int hd, int nq,
                                            const float* __restrict__ vecs, float* __restrict__ out, int out_ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < len; j += nwarps) {
    const uint16_t* row = base + (long long)j * ld;
    float dot[kMaxNQ];
#pragma unroll
    for (int k = 0; k < kMaxNQ; ++k) dot[k] = 0.f;
    for (int c = lane; c < hd / 8; c += 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + c * 8);
      const float2 a = unpack16(fmt, u.x), b = unpack16(fmt, u.y), cc = unpack16(fmt, u.z), d = unpack16(fmt, u.w);
#pragma unroll
      for (int k = 0; k < kMaxNQ; ++k) {
        if (k < nq) {
          const float4 q0 = *reinterpret_cast<const float4*>(vecs + k * hd + c * 8);
          const float4 q1 = *reinterpret_cast<const float4*>(vecs + k * hd + c * 8 + 4);
          dot[k] += a.x * q0.x + a.y * q0.y + b.x * q0.z + b.y * q0.w + cc.x * q1.x + cc.y * q1.y + d.x * q1.z + d.y * q1.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxNQ; ++k) {
      if (k < nq) {
        const float v = warp_sum(dot[k]);
        if (lane == 0) out[k * out_ld + j] = v;
      }
    }
  }
}

// acc[k] (a float2 = two adjacent dims) = sum_j w[k * w_ld + j] * row_j[pair]  for the pair of dims owned by this thread
__device__ __forceinline__ void mq_col_sums(const uint16_t* __restrict__ col, long long ld, int fmt, int len, int nq,
                                            const float* __restrict__ w, int w_ld, float2* acc) {
#pragma unroll
  for (int k = 0; k < kMaxNQ; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int j = 0; j < len; ++j) {
    const float2 vv = unpack16(fmt, *reinterpret_cast<const uint32_t*>(col + (long long)j * ld));
#pragma unroll
    for (int k = 0; k < kMaxNQ; ++k) {
      if (k < nq) {
        const float pj = w[k * w_ld + j];
        acc[k].x += pj * vv.x;
        acc[k].y += pj * vv.y;
      }
    }
  }
}

__global__ void __launch_bounds__(kMqThreads) mq_attention_fwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                      long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                      const int* __restrict__ kv_len, int Tk, int heads, int hd, int nq,
                                                                      float scale, float* __restrict__ probs, float* __restrict__ ctx,
                                                                      float drop_p, const long long* __restrict__ rng_state, int rng_site) {
  extern __shared__ __align__(16) float sm[];  // [nq][hd] q (scaled) | [nq][Tk] scores/probs | [8] scratch
  float* sq = sm;
  float* sc = sq + nq * hd;
  float* red = sc + nq * Tk;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * hd;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs + h * hd;
  for (int i = tid; i < nq * hd; i += kMqThreads) sq[i] = q[(i / hd) * D + h * hd + (i % hd)] * scale;
  __syncthreads();
  mq_row_dots(base + k_off, kv_ld, kv_fmt, len, hd, nq, sq, sc, Tk);
  __syncthreads();
  for (int k = 0; k < nq; ++k) {
    float mx = -INFINITY;
    for (int j = tid; j < len; j += kMqThreads) mx = fmaxf(mx, sc[k * Tk + j]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = tid; j < len; j += kMqThreads) {
      const float e = __expf(sc[k * Tk + j] - mx);
      sc[k * Tk + j] = e;
      sum += e;
    }
    sum = block_sum(sum, red);
    const float inv = 1.f / sum;
    float* po = probs + (((long long)b * heads + h) * nq + k) * Tk;
    const DropoutRng rng(rng_state, rng_site, drop_p);
    for (int j = tid; j < Tk; j += kMqThreads) {
      const float pj = j < len ? sc[k * Tk + j] * inv : 0.f;
      // attention dropout (nn.MultiheadAttention, train mode): probs keeps the UNdropped p, the context uses p_j * m_j
      if (j < len) sc[k * Tk + j] = drop_p > 0.f ? pj * rng.scale((((unsigned long long)b * heads + h) * nq + k) * Tk + j) : pj;
      po[j] = pj;
    }
  }
  __syncthreads();
  for (int pr = tid; pr < hd / 2; pr += kMqThreads) {
    float2 acc[kMaxNQ];
    mq_col_sums(base + v_off + pr * 2, kv_ld, kv_fmt, len, nq, sc, Tk, acc);
#pragma unroll
    for (int k = 0; k < kMaxNQ; ++k)
      if (k < nq) *reinterpret_cast<float2*>(ctx + ((long long)b * nq + k) * D + h * hd + pr * 2) = acc[k];
  }
}

// dctx fp32 [B][NQ][D].  dkv 16-bit (same layout as kv; rows >= len zero).  dq_part fp32 [B][NQ][D]: this utterance's
// contribution to dq (the caller sums over B -- no atomics, deterministic).
__global__ void __launch_bounds__(kMqThreads) mq_attention_bwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                      long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                      const int* __restrict__ kv_len, int Tk, int heads, int hd, int nq,
                                                                      float scale, const float* __restrict__ probs,
                                                                      const float* __restrict__ dctx, uint16_t* __restrict__ dkv, int dkv_fmt,
                                                                      float* __restrict__ dq_part, float drop_p,
                                                                      const long long* __restrict__ rng_state, int rng_site) {
  extern __shared__ __align__(16) float sm[];  // [nq][hd] q | [nq][hd] dctx | [nq][Tk] ds | [nq][Tk] p | [nq][Tk] p*m (dropout only) | [8]
  float* sq = sm;
  float* sdc = sq + nq * hd;
  float* sds = sdc + nq * hd;
  float* sp = sds + nq * Tk;
  float* spm = drop_p > 0.f ? sp + nq * Tk : sp;   // p_j * m_j: what multiplied V in the forward
  float* red = spm + nq * Tk;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * hd;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs + h * hd;
  uint16_t* dbase = dkv + (long long)b * kv_bs + h * hd;
  for (int i = tid; i < nq * hd; i += kMqThreads) {
    const int k = i / hd, d = i % hd;
    sq[i] = q[k * D + h * hd + d];
    sdc[i] = dctx[((long long)b * nq + k) * D + h * hd + d];
  }
  for (int i = tid; i < nq * Tk; i += kMqThreads) sp[i] = probs[((long long)b * heads + h) * nq * Tk + i];
  __syncthreads();
  mq_row_dots(base + v_off, kv_ld, kv_fmt, len, hd, nq, sdc, sds, Tk);   // dP[k][j] = <dctx_k, v_j>
  __syncthreads();
  if (drop_p > 0.f) {   // regenerate the forward's mask: d ctx / d p_j = m_j v_j, and dV sees p_j m_j
    const DropoutRng rng(rng_state, rng_site, drop_p);
    for (int i = tid; i < nq * Tk; i += kMqThreads) {
      const int k = i / Tk, j = i % Tk;
      const float mj = j < len ? rng.scale((((unsigned long long)b * heads + h) * nq + k) * Tk + j) : 0.f;
      sds[i] *= mj;
      spm[i] = sp[i] * mj;
    }
    __syncthreads();
  }
  for (int k = 0; k < nq; ++k) {
    float dot = 0.f;
    for (int j = tid; j < len; j += kMqThreads) dot += sp[k * Tk + j] * sds[k * Tk + j];
    dot = block_sum(dot, red);
    for (int j = tid; j < len; j += kMqThreads) sds[k * Tk + j] = sp[k * Tk + j] * (sds[k * Tk + j] - dot) * scale;
  }
  __syncthreads();
  // dK_j = sum_k ds[k][j] q_k ; dV_j = sum_k p[k][j] dctx_k : a 16-byte chunk of one key row per thread per step
  const int CH = hd / 8;
  for (int idx = tid; idx < Tk * CH; idx += kMqThreads) {
    const int j = idx / CH, c = idx % CH;
    float gk[8], gv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gk[i] = gv[i] = 0.f;
    if (j < len) {
      for (int k = 0; k < nq; ++k) {
        const float ds = sds[k * Tk + j], pj = spm[k * Tk + j];
        const float4 q0 = *reinterpret_cast<const float4*>(sq + k * hd + c * 8), q1 = *reinterpret_cast<const float4*>(sq + k * hd + c * 8 + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(sdc + k * hd + c * 8), g1 = *reinterpret_cast<const float4*>(sdc + k * hd + c * 8 + 4);
        gk[0] += ds * q0.x; gk[1] += ds * q0.y; gk[2] += ds * q0.z; gk[3] += ds * q0.w;
        gk[4] += ds * q1.x; gk[5] += ds * q1.y; gk[6] += ds * q1.z; gk[7] += ds * q1.w;
        gv[0] += pj * g0.x; gv[1] += pj * g0.y; gv[2] += pj * g0.z; gv[3] += pj * g0.w;
        gv[4] += pj * g1.x; gv[5] += pj * g1.y; gv[6] += pj * g1.z; gv[7] += pj * g1.w;
      }
    }
    uint4 uk, uv;
    uk.x = pack16(dkv_fmt, gk[0], gk[1]); uk.y = pack16(dkv_fmt, gk[2], gk[3]); uk.z = pack16(dkv_fmt, gk[4], gk[5]); uk.w = pack16(dkv_fmt, gk[6], gk[7]);
    uv.x = pack16(dkv_fmt, gv[0], gv[1]); uv.y = pack16(dkv_fmt, gv[2], gv[3]); uv.z = pack16(dkv_fmt, gv[4], gv[5]); uv.w = pack16(dkv_fmt, gv[6], gv[7]);
    *reinterpret_cast<uint4*>(dbase + (long long)j * kv_ld + k_off + c * 8) = uk;
    *reinterpret_cast<uint4*>(dbase + (long long)j * kv_ld + v_off + c * 8) = uv;
  }
  // dq_k (this utterance) = sum_j ds[k][j] k_j
  for (int pr = tid; pr < hd / 2; pr += kMqThreads) {
    float2 acc[kMaxNQ];
    mq_col_sums(base + k_off + pr * 2, kv_ld, kv_fmt, len, nq, sds, Tk, acc);
#pragma unroll
    for (int k = 0; k < kMaxNQ; ++k)
      if (k < nq) *reinterpret_cast<float2*>(dq_part + ((long long)b * nq + k) * D + h * hd + pr * 2) = acc[k];
  }
}

// ------------------------------------------------------------------------------------------------ keyword BatchNorm
// x, y fp32 [B][NK][D]; BatchNorm1d feature f = dim * NK + kw (kw_bn.py:116-123).  One thread per feature.
__global__ void batchnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                                     float* __restrict__ save_mean, float* __restrict__ save_rstd, int B, int NK, int D, float eps,
                                     float momentum, int training) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // t = kw * D + dim (coalesced over dim)
  if (t >= NK * D) return;
  const int kw = t / D, dim = t % D;
  const int f = dim * NK + kw;
  float mean, rstd;
  if (training) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[(long long)b * NK * D + t];
    mean = s / B;
    float v = 0.f;
    for (int b = 0; b < B; ++b) {
      const float dlt = x[(long long)b * NK * D + t] - mean;
      v += dlt * dlt;
    }
    const float var = v / B;
    rstd = rsqrtf(var + eps);
    if (running_mean) {
      running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * mean;
      running_var[f] = (1.f - momentum) * running_var[f] + momentum * (B > 1 ? v / (B - 1) : var);
    }
    if (save_mean) {
      save_mean[t] = mean;
      save_rstd[t] = rstd;
    }
  } else {
    mean = running_mean[f];
    rstd = rsqrtf(running_var[f] + eps);
  }
  const float g = gamma[f], be = beta[f];
  for (int b = 0; b < B; ++b) y[(long long)b * NK * D + t] = (x[(long long)b * NK * D + t] - mean) * rstd * g + be;
}

__global__ void batchnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ save_mean, const float* __restrict__ save_rstd, float* __restrict__ dx,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int NK, int D) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NK * D) return;
  const int kw = t / D, dim = t % D;
  const int f = dim * NK + kw;
  const float mean = save_mean[t], rstd = save_rstd[t], g = gamma[f];
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dy[(long long)b * NK * D + t];
    s1 += d;
    s2 += d * (x[(long long)b * NK * D + t] - mean) * rstd;
  }
  if (dgamma) dgamma[f] = s2;
  if (dbeta) dbeta[f] = s1;
  const float a = g * rstd / B;
  for (int b = 0; b < B; ++b) {
    const float xh = (x[(long long)b * NK * D + t] - mean) * rstd;
    dx[(long long)b * NK * D + t] = a * (B * dy[(long long)b * NK * D + t] - s1 - xh * s2);
  }
}

// ------------------------------------------------------------------------------------------------ cosine + vector quantiser
// One CTA per row r (= utterance x keyword).  dots[r][v] = <kw_r, E_v> on entry; cos on exit (masked ids = -inf).
//   cos = dot / max(|kw_r| |E_v|, 1e-8)   (F.cosine_similarity);  idx[r] = argmax (lowest index on ties)
//   stats[r] = {max cos, sum_v exp((cos - max) / temp), sum_v exp(cos - max), |kw_r|}
__global__ void __launch_bounds__(256) vq_forward_kernel(float* __restrict__ dots, const float* __restrict__ kw, const float* __restrict__ emb_norm,
                                                         int V, int D, long long ld, const int* __restrict__ mask_ids, int n_mask,
                                                         float temp, long long* __restrict__ idx, float* __restrict__ stats) {
  __shared__ float red[8];
  __shared__ int redi[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  float kn = 0.f;
  if (kw) {  // kw == NULL: the rows already hold final scores (standalone quantiser)
    float ss = 0.f;
    for (int d = tid; d < D; d += 256) {
      const float v = kw[(long long)r * D + d];
      ss += v * v;
    }
    kn = sqrtf(block_sum(ss, red));
  }
  float* row = dots + (long long)r * ld;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int v = tid; v < V; v += 256) {
    float c = kw ? row[v] / fmaxf(kn * emb_norm[v], 1e-8f) : row[v];
    for (int m = 0; m < n_mask; ++m)
      if (mask_ids[m] == v) c = -INFINITY;
    row[v] = c;
    if (c > mx) {
      mx = c;
      mi = v;
    }
  }
  // block argmax (value, lowest index)
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) {
      mx = om;
      mi = oi;
    }
  }
  __syncthreads();
  if (lane == 0) {
    red[warp] = mx;
    redi[warp] = mi;
  }
  __syncthreads();
  mx = red[0];
  mi = redi[0];
  for (int w = 1; w < 8; ++w)
    if (red[w] > mx || (red[w] == mx && redi[w] < mi)) {
      mx = red[w];
      mi = redi[w];
    }
  float st = 0.f, s1 = 0.f;
  const float it = 1.f / temp;
  for (int v = tid; v < V; v += 256) {
    const float c = row[v];
    st += __expf((c - mx) * it);
    s1 += __expf(c - mx);
  }
  st = block_sum(st, red);
  s1 = block_sum(s1, red);
  if (tid == 0) {
    idx[r] = mi;
    stats[r * 4 + 0] = mx;
    stats[r * 4 + 1] = st;
    stats[r * 4 + 2] = s1;
    stats[r * 4 + 3] = kn;
  }
}

// g[r][v] = d loss / d subword_prob on entry (= dkeywords @ E^T); on exit g = d loss / d cos:
//   p = exp((cos - max)/temp) / sum_t ;  dcos = p (g - <p, g>) / temp.   t2[r] = <dcos, cos> (for the cosine backward).
__global__ void __launch_bounds__(256) vq_backward_kernel(float* __restrict__ g, const float* __restrict__ cos, int V, long long ld,
                                                          const float* __restrict__ stats, float temp, float* __restrict__ t2) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float mx = stats[r * 4], inv_s = 1.f / stats[r * 4 + 1], it = 1.f / temp;
  float* gr = g + (long long)r * ld;
  const float* cr = cos + (long long)r * ld;
  float dot = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float c = cr[v];
    const float p = c == -INFINITY ? 0.f : __expf((c - mx) * it) * inv_s;
    dot += p * gr[v];
  }
  dot = block_sum(dot, red);
  float acc = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float c = cr[v];
    float d = 0.f;
    if (c != -INFINITY) {
      const float p = __expf((c - mx) * it) * inv_s;
      d = p * (gr[v] - dot) * it;
      acc += d * c;
    }
    gr[v] = d;
  }
  acc = block_sum(acc, red);
  if (tid == 0) t2[r] = acc;
}

// d kw[r][:] = t1[r][:] / |kw_r| - t2[r] * kw[r][:] / |kw_r|^2     (t1 = dcos @ (E / |E|))
__global__ void cosine_bwd_rows_kernel(const float* __restrict__ t1, const float* __restrict__ t2, const float* __restrict__ kw,
                                       const float* __restrict__ stats, float* __restrict__ dkw, int R, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * D) return;
  const int r = (int)(i / D);
  const float kn = fmaxf(stats[r * 4 + 3], 1e-20f);
  dkw[i] = t1[i] / kn - t2[r] * kw[i] / (kn * kn);
}

// Logging-only statistics of SimpleVectorQuantizer.forward (my_vector_quantizer.py:84-118):
//   hist[v] += 1 at the argmax, ent[r] = -sum_v p log(p + 1e-9) (row kernel); avg[v] = sum_r softmax(cos)[r][v] (column kernel)
__global__ void __launch_bounds__(256) vq_row_stats_kernel(const float* __restrict__ cos, int V, long long ld, const float* __restrict__ stats,
                                                           const long long* __restrict__ idx, float* __restrict__ hist, float* __restrict__ ent) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float mx = stats[r * 4], inv_s1 = 1.f / stats[r * 4 + 2];
  const float* cr = cos + (long long)r * ld;
  float e = 0.f;
  for (int v = tid; v < V; v += 256) {
    const float c = cr[v];
    const float p = c == -INFINITY ? 0.f : __expf(c - mx) * inv_s1;
    e -= p * __logf(p + 1e-9f);
  }
  e = block_sum(e, red);
  if (tid == 0) {
    ent[r] = e;
    atomicAdd(&hist[idx[r]], 1.0f);
  }
}

// 32 columns x 8 row lanes per CTA; every row of a 32-column strip is one 128-byte read
__global__ void __launch_bounds__(256) vq_col_avg_kernel(const float* __restrict__ cos, int R, int V, long long ld, const float* __restrict__ stats,
                                                         float* __restrict__ avg) {
  __shared__ float part[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int v = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (v < V) {
    for (int r = ry; r < R; r += 8) {
      const float c = cos[(long long)r * ld + v];
      if (c != -INFINITY) acc += __expf(c - stats[r * 4]) / stats[r * 4 + 2];
    }
  }
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && v < V) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][cx];
    avg[v] += t;
  }
}

// x0[b][0] = E[sot] + pos[0]; x0[b][1+k] = E[idx[b][k]] + pos[1+k]; x0[b][K+1] = E[eot] + pos[K+1]; keywords[b][k] = E[idx[b][k]]
__global__ void keyword_embed_kernel(const float* __restrict__ emb, const float* __restrict__ pos, const long long* __restrict__ idx,
                                     long long sot, long long eot, int B, int K, int D, float* __restrict__ x0, float* __restrict__ keywords) {
  const int L = K + 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * L * D) return;
  const int d = (int)(i % D);
  const int l = (int)((i / D) % L);
  const int b = (int)(i / ((long long)D * L));
  long long tok;
  if (l == 0) tok = sot;
  else if (l == K + 1) tok = eot;
  else tok = idx[(long long)b * K + (l - 1)];
  const float e = emb[tok * D + d];
  x0[i] = e + pos[l * D + d];
  if (l >= 1 && l <= K && keywords) keywords[((long long)b * K + (l - 1)) * D + d] = e;
}

// x[b][l] = E[tokens[b][l]] + pos[l]   (ClipModel.encode_text, clip_official.py:211-218)
__global__ void token_embed_kernel(const float* __restrict__ emb, const float* __restrict__ pos, const long long* __restrict__ tokens,
                                   long long n_rows, int L, int D, long long vocab, float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * D) return;
  const int d = (int)(i % D);
  const long long r = i / D;
  long long tok = tokens[r];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  x[i] = emb[tok * D + d] + pos[(r % L) * D + d];
}

// out[b][:] = src[b][row[b]][:]
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ row, int B, int L, int D, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int b = (int)(i / D), d = (int)(i % D);
  long long r = row[b];
  r = r < 0 ? 0 : (r >= L ? L - 1 : r);
  out[i] = src[((long long)b * L + r) * D + d];
}

// ------------------------------------------------------------------------------------------------ small attention backward
// qkv 16-bit [B][L][3*heads*hd] (q | k | v), dctx fp32 [B][L][heads*hd] -> dqkv fp32 (same layout as qkv).  L <= 32, hd <= 64.
// One CTA (128 threads) per (b, h): S, P, dP, dS in shared memory, fp32.
__global__ void __launch_bounds__(128) attention_small_bwd_kernel(const uint16_t* __restrict__ qkv, int fmt, const float* __restrict__ dctx,
                                                                  float* __restrict__ dqkv, int L, int heads, int hd, float scale, int causal) {
  extern __shared__ float sm[];  // q, k, v, do : [L][hd] each | P, dS : [L][L]
  float* sq = sm;
  float* sk = sq + L * hd;
  float* sv = sk + L * hd;
  float* sdo = sv + L * hd;
  float* sP = sdo + L * hd;
  float* sdS = sP + L * L;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads, tid = threadIdx.x;
  const int D = heads * hd;
  for (int i = tid; i < L * hd; i += 128) {
    const int l = i / hd, d = i % hd;
    const long long row = ((long long)b * L + l) * 3 * D + h * hd + d;
    if (fmt == SCB_BF16) {
      sq[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[row]);
      sk[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[row + D]);
      sv[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(qkv)[row + 2 * D]);
    } else {
      sq[i] = __half2float(reinterpret_cast<const __half*>(qkv)[row]);
      sk[i] = __half2float(reinterpret_cast<const __half*>(qkv)[row + D]);
      sv[i] = __half2float(reinterpret_cast<const __half*>(qkv)[row + 2 * D]);
    }
    sdo[i] = dctx[((long long)b * L + l) * D + h * hd + d];
  }
  __syncthreads();
  for (int i = tid; i < L * L; i += 128) {  // S and dP
    const int r = i / L, c = i % L;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) {
      s += sq[r * hd + d] * sk[c * hd + d];
      dp += sdo[r * hd + d] * sv[c * hd + d];
    }
    sP[i] = (causal && c > r) ? -INFINITY : s * scale;
    sdS[i] = dp;
  }
  __syncthreads();
  if (tid < L) {  // row softmax and dS = P (dP - <P, dP>) * scale
    const int r = tid;
    float mx = -INFINITY;
    for (int c = 0; c < L; ++c) mx = fmaxf(mx, sP[r * L + c]);
    float sum = 0.f;
    for (int c = 0; c < L; ++c) {
      const float e = __expf(sP[r * L + c] - mx);
      sP[r * L + c] = e;
      sum += e;
    }
    float dot = 0.f;
    for (int c = 0; c < L; ++c) {
      sP[r * L + c] /= sum;
      dot += sP[r * L + c] * sdS[r * L + c];
    }
    for (int c = 0; c < L; ++c) sdS[r * L + c] = sP[r * L + c] * (sdS[r * L + c] - dot) * scale;
  }
  __syncthreads();
  for (int i = tid; i < L * hd; i += 128) {
    const int l = i / hd, d = i % hd;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int c = 0; c < L; ++c) {
      dq += sdS[l * L + c] * sk[c * hd + d];
      dk += sdS[c * L + l] * sq[c * hd + d];
      dv += sP[c * L + l] * sdo[c * hd + d];
    }
    const long long row = ((long long)b * L + l) * 3 * D + h * hd + d;
    dqkv[row] = dq;
    dqkv[row + D] = dk;
    dqkv[row + 2 * D] = dv;
  }
}

__global__ void act16_fwd_kernel(const uint16_t* __restrict__ pre, int fmt, int act, uint16_t* __restrict__ out, long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    float2 v = unpack16(fmt, reinterpret_cast<const uint32_t*>(pre)[i]);
    if (act == SCB_ACT_GELU_ERF) {
      v.x = gelu_erf(v.x);
      v.y = gelu_erf(v.y);
    } else if (act == SCB_ACT_QUICK_GELU) {
      v.x = quick_gelu(v.x);
      v.y = quick_gelu(v.y);
    }
    reinterpret_cast<uint32_t*>(out)[i] = pack16(fmt, v.x, v.y);
  }
}

// dx = dy * act'(pre), pre 16-bit
__global__ void act_bwd_kernel(const float* __restrict__ dy, const uint16_t* __restrict__ pre, int fmt, int act, float* __restrict__ dx,
                               long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = fmt == SCB_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pre)[i])
                                    : __half2float(reinterpret_cast<const __half*>(pre)[i]);
    float d = 1.f;
    if (act == SCB_ACT_GELU_ERF) {
      d = gelu_erf_grad(x);
    } else if (act == SCB_ACT_QUICK_GELU) {
      const float s = 1.f / (1.f + expf(-1.702f * x));
      d = s * (1.f + 1.702f * x * (1.f - s));
    }
    dx[i] = dy[i] * d;
  }
}

// Row softmax over the first len[b] columns (a warp per row); 16-bit output, zero beyond len up to out_cols.
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long ld, long long rows, int rows_per_batch, const int* __restrict__ len,
                                    int cols, uint16_t* __restrict__ out, int fmt, long long out_ld, int out_cols) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int n = len ? min(len[r / rows_per_batch], cols) : cols;
  const float* x = s + r * ld;
  float mx = -INFINITY;
  for (int c = lane; c < n; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += __expf(x[c] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int c = lane; c < out_cols; c += 32) {
    const float p = c < n ? __expf(x[c] - mx) * inv : 0.f;
    if (fmt == SCB_BF16) reinterpret_cast<__nv_bfloat16*>(out)[r * out_ld + c] = __float2bfloat16(p);
    else reinterpret_cast<__half*>(out)[r * out_ld + c] = __float2half(p);
  }
}

// hi = x with the 13 low mantissa bits cleared (exactly what kind::tf32 reads), lo = x - hi (exact in fp32).
// role 0 (left operand):  dst row = [hi | lo | hi];  role 1 (right operand): dst row = [hi | hi | lo]
//   => sum over 3*cols of left*right = hi*hi + lo*hi + hi*lo  (the lo*lo term, ~2^-22 relative, is dropped)
__global__ void split_tf32_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst, long long rows, int cols, int role) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    const float x = src[r * src_ld + c];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const float lo = x - hi;
    float* d = dst + r * 3 * cols + c;
    d[0] = hi;
    d[cols] = role == 0 ? lo : hi;
    d[2 * cols] = role == 0 ? hi : lo;
  }
}

}  // namespace

#define SCB_SMEM_ATTR(kernel, bytes)                                                                           \
  do {                                                                                                         \
    static size_t cfg_ = 0;                                                                                    \
    if ((size_t)(bytes) > 48 * 1024 && (size_t)(bytes) > cfg_) {                                              \
      SCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));       \
      cfg_ = (bytes);                                                                                          \
    }                                                                                                          \
  } while (0)

int mq_attention_fwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off, const int* kv_len,
                     int batch, int heads, int head_dim, int nq, int Tk, float scale, float* probs, float* ctx, float drop_p,
                     const long long* rng_state, int rng_site, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && ctx, SCB_EINVAL, "scb_mq_attention_fwd: null operand");
  SCB_CHECK(drop_p == 0.f || (drop_p > 0.f && drop_p < 1.f && rng_state), SCB_EINVAL, "scb_mq_attention_fwd: dropout needs p in [0,1) and an rng_state");
  SCB_CHECK(nq >= 1 && nq <= kMaxNQ, SCB_EUNSUPPORTED, "scb_mq_attention_fwd: nq=%d out of [1,%d]", nq, kMaxNQ);
  SCB_CHECK(head_dim % 8 == 0 && kv_ld % 8 == 0 && kv_bs % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0, SCB_EINVAL,
            "scb_mq_attention_fwd: head_dim / strides / offsets must be multiples of 8");
  if (batch == 0) return SCB_OK;
  const size_t smem = (size_t)(nq * Tk + nq * head_dim + 8) * sizeof(float);
  SCB_CHECK(smem <= 200 * 1024, SCB_EUNSUPPORTED, "scb_mq_attention_fwd: nq*Tk + nq*head_dim too large");
  SCB_SMEM_ATTR(mq_attention_fwd_kernel, smem);
  mq_attention_fwd_kernel<<<batch * heads, kMqThreads, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads,
                                                           head_dim, nq, scale, probs, ctx, drop_p, rng_state, rng_site);
  note_launch();
  SCB_LAUNCH_OK("mq_attention_fwd");
  return SCB_OK;
}

int mq_attention_bwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off, const int* kv_len,
                     int batch, int heads, int head_dim, int nq, int Tk, float scale, const float* probs, const float* dctx, void* dkv,
                     int dkv_fmt, float* dq, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && dctx && dkv && dq, SCB_EINVAL, "scb_mq_attention_bwd: null operand");
  SCB_CHECK(drop_p == 0.f || (drop_p > 0.f && drop_p < 1.f && rng_state), SCB_EINVAL, "scb_mq_attention_bwd: dropout needs p in [0,1) and an rng_state");
  SCB_CHECK(nq >= 1 && nq <= kMaxNQ, SCB_EUNSUPPORTED, "scb_mq_attention_bwd: nq=%d out of [1,%d]", nq, kMaxNQ);
  SCB_CHECK(head_dim % 8 == 0 && kv_ld % 8 == 0 && kv_bs % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0, SCB_EINVAL,
            "scb_mq_attention_bwd: head_dim / strides / offsets must be multiples of 8");
  if (batch == 0) return SCB_OK;
  const size_t smem = (size_t)((drop_p > 0.f ? 3 : 2) * nq * Tk + 2 * nq * head_dim + 8) * sizeof(float);
  SCB_CHECK(smem <= 200 * 1024, SCB_EUNSUPPORTED, "scb_mq_attention_bwd: nq*Tk + nq*head_dim too large");
  SCB_SMEM_ATTR(mq_attention_bwd_kernel, smem);
  mq_attention_bwd_kernel<<<batch * heads, kMqThreads, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads,
                                                           head_dim, nq, scale, probs, dctx, (uint16_t*)dkv, dkv_fmt, dq, drop_p, rng_state, rng_site);
  note_launch();
  SCB_LAUNCH_OK("mq_attention_bwd");
  return SCB_OK;
}

int batchnorm_fwd(const float* x, float* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  float* save_mean, float* save_rstd, int B, int NK, int D, float eps, float momentum, int training, cudaStream_t st) {
  SCB_CHECK(x && y && gamma && beta, SCB_EINVAL, "scb_batchnorm_fwd: null operand");
  SCB_CHECK(training || (running_mean && running_var), SCB_EINVAL, "scb_batchnorm_fwd: eval mode needs running statistics");
  if (B == 0) return SCB_OK;
  batchnorm_fwd_kernel<<<(NK * D + 127) / 128, 128, 0, st>>>(x, y, gamma, beta, running_mean, running_var, save_mean, save_rstd, B, NK, D, eps,
                                                            momentum, training);
  note_launch();
  SCB_LAUNCH_OK("batchnorm_fwd");
  return SCB_OK;
}

int batchnorm_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                  float* dgamma, float* dbeta, int B, int NK, int D, cudaStream_t st) {
  SCB_CHECK(dy && x && gamma && save_mean && save_rstd && dx, SCB_EINVAL, "scb_batchnorm_bwd: null operand");
  if (B == 0) return SCB_OK;
  batchnorm_bwd_kernel<<<(NK * D + 127) / 128, 128, 0, st>>>(dy, x, gamma, save_mean, save_rstd, dx, dgamma, dbeta, B, NK, D);
  note_launch();
  SCB_LAUNCH_OK("batchnorm_bwd");
  return SCB_OK;
}

int vq_forward(float* dots, const float* kw, const float* emb_norm, int R, int V, int D, long long ld, const int* mask_ids, int n_mask,
               float temp, long long* idx, float* stats, cudaStream_t st) {
  SCB_CHECK(dots && idx && stats && ((kw == nullptr) == (emb_norm == nullptr)), SCB_EINVAL, "scb_vq_forward: null operand");
  SCB_CHECK(temp > 0.f, SCB_EINVAL, "scb_vq_forward: temperature must be positive");
  if (R == 0) return SCB_OK;
  vq_forward_kernel<<<R, 256, 0, st>>>(dots, kw, emb_norm, V, D, ld, mask_ids, n_mask, temp, idx, stats);
  note_launch();
  SCB_LAUNCH_OK("vq_forward");
  return SCB_OK;
}

int vq_backward(float* g, const float* cos, int R, int V, long long ld, const float* stats, float temp, float* t2, cudaStream_t st) {
  SCB_CHECK(g && cos && stats && t2, SCB_EINVAL, "scb_vq_backward: null operand");
  if (R == 0) return SCB_OK;
  vq_backward_kernel<<<R, 256, 0, st>>>(g, cos, V, ld, stats, temp, t2);
  note_launch();
  SCB_LAUNCH_OK("vq_backward");
  return SCB_OK;
}

int cosine_bwd_rows(const float* t1, const float* t2, const float* kw, const float* stats, float* dkw, int R, int D, cudaStream_t st) {
  SCB_CHECK(t1 && t2 && kw && stats && dkw, SCB_EINVAL, "scb_cosine_bwd_rows: null operand");
  if (R == 0) return SCB_OK;
  cosine_bwd_rows_kernel<<<(unsigned)(((long long)R * D + 255) / 256), 256, 0, st>>>(t1, t2, kw, stats, dkw, R, D);
  note_launch();
  SCB_LAUNCH_OK("cosine_bwd_rows");
  return SCB_OK;
}

int vq_diagnostics(const float* cos, int R, int V, long long ld, const float* stats, const long long* idx, float* hist, float* avg, float* ent,
                   cudaStream_t st) {
  SCB_CHECK(cos && stats && idx && hist && avg && ent, SCB_EINVAL, "scb_vq_diagnostics: null operand");
  if (R == 0) return SCB_OK;
  vq_row_stats_kernel<<<R, 256, 0, st>>>(cos, V, ld, stats, idx, hist, ent);
  note_launch();
  vq_col_avg_kernel<<<(V + 31) / 32, 256, 0, st>>>(cos, R, V, ld, stats, avg);
  note_launch();
  SCB_LAUNCH_OK("vq_diagnostics");
  return SCB_OK;
}

int keyword_embed(const float* emb, const float* pos, const long long* idx, long long sot, long long eot, int B, int K, int D, float* x0,
                  float* keywords, cudaStream_t st) {
  SCB_CHECK(emb && pos && idx && x0, SCB_EINVAL, "scb_keyword_embed: null operand");
  if (B == 0) return SCB_OK;
  const long long n = (long long)B * (K + 2) * D;
  keyword_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emb, pos, idx, sot, eot, B, K, D, x0, keywords);
  note_launch();
  SCB_LAUNCH_OK("keyword_embed");
  return SCB_OK;
}

int token_embed(const float* emb, const float* pos, const long long* tokens, int B, int L, int D, long long vocab, float* x, cudaStream_t st) {
  SCB_CHECK(emb && pos && tokens && x, SCB_EINVAL, "scb_token_embed: null operand");
  if (B == 0) return SCB_OK;
  const long long n = (long long)B * L * D;
  token_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emb, pos, tokens, (long long)B * L, L, D, vocab, x);
  note_launch();
  SCB_LAUNCH_OK("token_embed");
  return SCB_OK;
}

int gather_rows(const float* src, const long long* row, int B, int L, int D, float* out, cudaStream_t st) {
  SCB_CHECK(src && row && out, SCB_EINVAL, "scb_gather_rows: null operand");
  if (B == 0) return SCB_OK;
  gather_rows_kernel<<<(unsigned)(((long long)B * D + 255) / 256), 256, 0, st>>>(src, row, B, L, D, out);
  note_launch();
  SCB_LAUNCH_OK("gather_rows");
  return SCB_OK;
}

int attention_small_bwd(const void* qkv, int fmt, const float* dctx, float* dqkv, int batch, int L, int heads, int head_dim, float scale,
                        int causal, cudaStream_t st) {
  SCB_CHECK(qkv && dctx && dqkv, SCB_EINVAL, "scb_attention_small_bwd: null operand");
  SCB_CHECK(L >= 1 && L <= 128 && head_dim <= 128, SCB_EUNSUPPORTED, "scb_attention_small_bwd: L=%d / head_dim=%d too large", L, head_dim);
  if (batch == 0) return SCB_OK;
  const size_t smem = (size_t)(4 * L * head_dim + 2 * L * L) * sizeof(float);
  SCB_CHECK(smem <= 200 * 1024, SCB_EUNSUPPORTED, "scb_attention_small_bwd: shared memory");
  SCB_SMEM_ATTR(attention_small_bwd_kernel, smem);
  attention_small_bwd_kernel<<<batch * heads, 128, smem, st>>>((const uint16_t*)qkv, fmt, dctx, dqkv, L, heads, head_dim, scale, causal);
  note_launch();
  SCB_LAUNCH_OK("attention_small_bwd");
  return SCB_OK;
}

int softmax_rows(const float* s, long long ld, long long rows, int rows_per_batch, const int* len, int cols, void* out, int fmt,
                 long long out_ld, int out_cols, cudaStream_t st) {
  SCB_CHECK(s && out && rows_per_batch > 0 && (fmt == SCB_F16 || fmt == SCB_BF16), SCB_EINVAL, "scb_softmax_rows: bad operand");
  if (rows == 0) return SCB_OK;
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(s, ld, rows, rows_per_batch, len, cols, (uint16_t*)out, fmt, out_ld, out_cols);
  note_launch();
  SCB_LAUNCH_OK("softmax_rows");
  return SCB_OK;
}

int split_tf32(const float* src, long long src_ld, float* dst, long long rows, int cols, int role, cudaStream_t st) {
  SCB_CHECK(src && dst && (role == 0 || role == 1), SCB_EINVAL, "scb_split_tf32: null operand or bad role");
  if (rows == 0) return SCB_OK;
  long long blocks = (rows * cols + 255) / 256;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, src_ld, dst, rows, cols, role);
  note_launch();
  SCB_LAUNCH_OK("split_tf32");
  return SCB_OK;
}

int act16_fwd(const void* pre, int fmt, int act, void* out, long long n, cudaStream_t st) {
  SCB_CHECK(pre && out && n % 2 == 0, SCB_EINVAL, "scb_act16_fwd: null operand or odd length");
  if (n == 0) return SCB_OK;
  long long blocks = (n / 2 + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  act16_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint16_t*)pre, fmt, act, (uint16_t*)out, n / 2);
  note_launch();
  SCB_LAUNCH_OK("act16_fwd");
  return SCB_OK;
}

int act_bwd(const float* dy, const void* pre, int fmt, int act, float* dx, long long n, cudaStream_t st) {
  SCB_CHECK(dy && pre && dx, SCB_EINVAL, "scb_act_bwd: null operand");
  if (n == 0) return SCB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  act_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(dy, (const uint16_t*)pre, fmt, act, dx, n);
  note_launch();
  SCB_LAUNCH_OK("act_bwd");
  return SCB_OK;
}

}  // namespace scb
