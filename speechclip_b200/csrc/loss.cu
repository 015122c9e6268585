// Masked symmetric InfoNCE (avssl/module/losses.py:185-245) forward + backward in fp32, and the small fp32 SIMT GEMM it uses.
//
//   logits = A B^T * mult (- margin on the diagonal)
//   neg_ij = (id_i != id_j) | (i == j & !dcl)           (ids == NULL: i != j, | diagonal unless dcl)
//   loss   = [a2b] mean_i(-l_ii + log sum_j e^{l_ij} neg_ij) + [b2a] mean_j(-l_jj + log sum_i e^{l_ij} neg_ij)   (/2 if both)
//   G      = dloss/dlogits ;  dA = mult G B ;  dB = mult G^T A ;  dlog_mult = sum G (l + margin I)
//
// No max-subtraction, exactly like the reference (|logit| <= 1/0.07 keeps exp in fp32 range).  B is not capped at 256
// (the reference's MAX_EYE buffer, losses.py:126, raises IndexError beyond that; semantics here are eye(B)).
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

// C[m,n] = alpha * sum_k A(m,k) B(n,k) + beta * C[m,n]; A(m,k) = a[m*a_rs + k*a_cs], B(n,k) = b[n*b_rs + k*b_cs].
constexpr int TM = 64, TN = 64, TK = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ a, long long a_rs, long long a_cs, const float* __restrict__ b,
                                                    long long b_rs, long long b_cs, float* __restrict__ c, long long ldc, int M, int N, int K,
                                                    float alpha, float beta) {
  __shared__ float sa[TK][TM + 4];
  __shared__ float sb[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int mm, kk;
      if (a_cs == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      const int m = m0 + mm, k = k0 + kk;
      sa[kk][mm] = (m < M && k < K) ? a[m * a_rs + k * a_cs] : 0.f;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int nn, kk;
      if (b_cs == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      const int n = n0 + nn, k = k0 + kk;
      sb[kk][nn] = (n < N && k < K) ? b[n * b_rs + k * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) {
        float* p = c + (long long)m * ldc + n;
        *p = alpha * acc[i][j] + (beta != 0.f ? beta * *p : 0.f);
      }
    }
  }
}

// Skinny case M <= 8 (the [CLS] query projection and its gradient: one row against a d x d matrix).  The tiled kernel above
// walks K in 16-wide steps with two block barriers each on a dozen blocks (~100 us for 1 x 768 x 768); here every output
// column is a coalesced dot product.  KCONTIG: B rows are contiguous in k (b_cs == 1), a warp per column, lanes stride k.
// Otherwise B is contiguous in n (b_rs == 1): a block owns 32 columns (lane = column) and its 32 warps split k.
template <bool KCONTIG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sgemm_skinny_kernel(const float* __restrict__ a, long long a_rs, long long a_cs,
                                                           const float* __restrict__ b, long long b_rs, long long b_cs,
                                                           float* __restrict__ c, long long ldc, int M, int N, int K, float alpha, float beta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  if (KCONTIG) {
    const int n = blockIdx.x * WARPS + warp;
    if (n >= N) return;
    const float* br = b + (long long)n * b_rs;
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      const float bv = br[k];
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (m < M) acc[m] = fmaf(a[m * a_rs + k * a_cs], bv, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < M) {
        const float s = warp_sum(acc[m]);
        if (lane == 0) {
          float* p = c + (long long)m * ldc + n;
          *p = alpha * s + (beta != 0.f ? beta * *p : 0.f);
        }
      }
    }
  } else {
    __shared__ float red[WARPS][8][33];
    const int n = blockIdx.x * 32 + lane;
    if (n < N) {
#pragma unroll 4
      for (int k = warp; k < K; k += WARPS) {
        const float bv = b[(long long)k * b_cs + n];
#pragma unroll
        for (int m = 0; m < 8; ++m)
          if (m < M) acc[m] = fmaf(a[m * a_rs + k * a_cs], bv, acc[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) red[warp][m][lane] = acc[m];
    __syncthreads();
    if (warp < M && n < N) {  // warp m sums the WARPS partials of row m
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s += red[w][warp][lane];
      float* p = c + (long long)warp * ldc + n;
      *p = alpha * s + (beta != 0.f ? beta * *p : 0.f);
    }
  }
}

struct NceCfg {
  int B;
  const long long* ids;
  const float* log_mult;  // nullable: learnable temperature parameter (multiplier = exp(*log_mult))
  float fixed_mult;
  float margin;
  int dcl, a2b, b2a;
};

__device__ __forceinline__ bool neg_mask(const NceCfg& c, int i, int j) {
  if (i == j) return !c.dcl;
  return c.ids ? (c.ids[i] != c.ids[j]) : true;
}

// logits *= mult (sgemm wrote raw cosines), diagonal -= margin.
__global__ void nce_scale_kernel(float* __restrict__ l, NceCfg c) {
  const float mult = c.log_mult ? __expf(*c.log_mult) : c.fixed_mult;
  const long long total = (long long)c.B * c.B;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / c.B), j = (int)(idx % c.B);
    float v = l[idx] * mult;
    if (i == j) v -= c.margin;
    l[idx] = v;
  }
}

// rowsum[i] = sum_j e^{l_ij} neg_ij (one warp per row)
__global__ void __launch_bounds__(256) nce_rowsum_kernel(const float* __restrict__ l, NceCfg c, float* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= c.B) return;
  float s = 0.f;
  for (int j = lane; j < c.B; j += 32)
    if (neg_mask(c, i, j)) s += __expf(l[(long long)i * c.B + j]);
  s = warp_sum(s);
  if (lane == 0) rowsum[i] = s;
}
// colsum[j] = sum_i e^{l_ij} neg_ij (thread per column, coalesced across the warp; rows split over blockIdx.y + atomics)
__global__ void __launch_bounds__(256) nce_colsum_kernel(const float* __restrict__ l, NceCfg c, float* __restrict__ colsum, int rows_per_block) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c.B) return;
  const int i0 = blockIdx.y * rows_per_block, i1 = min(c.B, i0 + rows_per_block);
  float s = 0.f;
  for (int i = i0; i < i1; ++i)
    if (neg_mask(c, i, j)) s += __expf(l[(long long)i * c.B + j]);
  atomicAdd(&colsum[j], s);
}

__global__ void __launch_bounds__(1024) nce_loss_kernel(const float* __restrict__ l, NceCfg c, const float* __restrict__ rowsum,
                                                        const float* __restrict__ colsum, float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < c.B; i += blockDim.x) {
    const float pos = l[(long long)i * c.B + i];
    if (c.a2b) s += -pos + __logf(rowsum[i]);
    if (c.b2a) s += -pos + __logf(colsum[i]);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss = v / c.B / ((c.a2b && c.b2a) ? 2.f : 1.f);
  }
}

// l <- G in place; dlog_mult += sum G (l + margin I)
__global__ void __launch_bounds__(256) nce_grad_kernel(float* __restrict__ l, NceCfg c, const float* __restrict__ rowsum,
                                                       const float* __restrict__ colsum, float upstream, const float* __restrict__ upstream_dev,
                                                       float* __restrict__ dlog_mult) {
  const float gs = upstream * (upstream_dev ? *upstream_dev : 1.f) / c.B / ((c.a2b && c.b2a) ? 2.f : 1.f);
  const long long total = (long long)c.B * c.B;
  float acc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / c.B), j = (int)(idx % c.B);
    const float lv = l[idx];
    float g = 0.f;
    if (neg_mask(c, i, j)) {
      const float e = __expf(lv);
      if (c.a2b) g += e / rowsum[i];
      if (c.b2a) g += e / colsum[j];
    }
    if (i == j) g -= (float)(c.a2b + c.b2a);
    g *= gs;
    l[idx] = g;
    acc += g * (lv + (i == j ? c.margin : 0.f));
  }
  if (dlog_mult) {
    acc = warp_sum(acc);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
      atomicAdd(dlog_mult, v);
    }
  }
}

__global__ void scale_by_mult_kernel(float* __restrict__ x, long long n, const float* __restrict__ log_mult, float fixed_mult) {
  const float mult = log_mult ? __expf(*log_mult) : fixed_mult;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= mult;
}

}  // namespace

int sgemm(const float* a, long long a_rs, long long a_cs, const float* b, long long b_rs, long long b_cs, float* c, long long ldc, int M, int N,
          int K, float alpha, float beta, cudaStream_t st) {
  SCB_CHECK(a && b && c, SCB_EINVAL, "scb_sgemm: null operand");
  if (M == 0 || N == 0) return SCB_OK;
  if (M <= 8 && K >= 64 && b_cs == 1)
    sgemm_skinny_kernel<true, 8><<<(N + 7) / 8, 256, 0, st>>>(a, a_rs, a_cs, b, b_rs, b_cs, c, ldc, M, N, K, alpha, beta);
  else if (M <= 8 && K >= 64 && b_rs == 1)
    sgemm_skinny_kernel<false, 32><<<(N + 31) / 32, 1024, 0, st>>>(a, a_rs, a_cs, b, b_rs, b_cs, c, ldc, M, N, K, alpha, beta);
  else
    sgemm_kernel<<<dim3((N + TN - 1) / TN, (M + TM - 1) / TM), 256, 0, st>>>(a, a_rs, a_cs, b, b_rs, b_cs, c, ldc, M, N, K, alpha, beta);
  note_launch();
  SCB_LAUNCH_OK("sgemm");
  return SCB_OK;
}

// Above this batch the three B x B x D contractions (25.8 GFLOP each at B=4096, D=768: BASELINE config 5) run on the tensor
// cores as TF32 (fp32 operands, fp32 accumulate) instead of the fp32 SIMT kernel; below it the SIMT kernel keeps the loss and
// its gradients at fp32-exact parity with the reference fixtures.
constexpr int kNceTensorMinB = 1024;
static bool nce_tensor_path(int B, int D) { return B >= kNceTensorMinB && B % 8 == 0 && D % 8 == 0; }

// scratch: logits [B*B] | rowsum [B] | colsum [B] | pad ; tensor path adds G^T [B*B] and the two transposed feature matrices
long long infonce_scratch_bytes2(int B, int D) {
  long long fl = (long long)B * B + 2LL * B + 64;
  if (nce_tensor_path(B, D)) fl += (long long)B * B + 2LL * B * D;
  return fl * (long long)sizeof(float);
}
long long infonce_scratch_bytes(int B) { return infonce_scratch_bytes2(B, 1024); }  // D-independent upper bound for D <= 1024

static int tf32_gemm(const float* a, long long a_ld, const float* b, long long b_ld, float* c, long long ldc, int M, int N, int K,
                     cudaStream_t st) {
  scb_gemm_args g{};
  g.a = a; g.a_inner = K; g.a_rows = M; g.a_row_stride = a_ld; g.batch = 1; g.m_per_batch = M;
  g.kb_per_tap = (K + 31) / 32; g.b = b; g.b_row_stride = b_ld; g.n = N; g.k = K; g.groups = 1;
  g.out = c; g.out_dtype = SCB_F32; g.ldc = ldc; g.ab_format = SCB_F32; g.alpha = 1.f;
  return gemm(g, st);
}

int infonce(const float* feat_a, const float* feat_b, const long long* ids, int B, int D, const float* log_mult, float fixed_mult,
            float margin, int dcl, int a2b, int b2a, int phase, float* loss, float* logits_out, float upstream, const float* upstream_dev,
            float* dA, float* dB, float* dlog_mult, void* scratch, long long scratch_bytes, cudaStream_t st) {
  SCB_CHECK(feat_a && feat_b && scratch, SCB_EINVAL, "scb_infonce: null operand");
  SCB_CHECK(a2b || b2a, SCB_EINVAL, "scb_infonce: a2b and b2a cannot both be off");  // losses.py:154
  SCB_CHECK(B > 0 && D > 0, SCB_EINVAL, "scb_infonce: empty batch");
  SCB_CHECK(phase >= 1 && phase <= 3, SCB_EINVAL, "scb_infonce: phase must be 1 (forward), 2 (backward) or 3 (both)");
  SCB_CHECK(!(phase & 1) || loss, SCB_EINVAL, "scb_infonce: forward phase needs loss");
  SCB_CHECK(scratch_bytes >= infonce_scratch_bytes2(B, D), SCB_EINVAL, "scb_infonce: scratch too small (%lld < %lld)", scratch_bytes,
            infonce_scratch_bytes2(B, D));
  float* l = reinterpret_cast<float*>(scratch);
  float* rowsum = l + (long long)B * B;
  float* colsum = rowsum + B;
  NceCfg c{B, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a};
  const long long total = (long long)B * B;
  unsigned eb = (unsigned)((total + 255) / 256);
  if (eb > 8u * num_sms()) eb = 8u * num_sms();
  int e;
  if (phase & 1) {
    e = nce_tensor_path(B, D) ? tf32_gemm(feat_a, D, feat_b, D, l, B, B, B, D, st) : sgemm(feat_a, D, 1, feat_b, D, 1, l, B, B, B, D, 1.f, 0.f, st);
    if (e) return e;
    nce_scale_kernel<<<eb, 256, 0, st>>>(l, c);
    note_launch();
    if (logits_out) SCB_CUDA(cudaMemcpyAsync(logits_out, l, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SCB_CUDA(cudaMemsetAsync(colsum, 0, B * sizeof(float), st));
    nce_rowsum_kernel<<<(B + 7) / 8, 256, 0, st>>>(l, c, rowsum);
    note_launch();
    const int rpb = 64;
    nce_colsum_kernel<<<dim3((B + 255) / 256, (B + rpb - 1) / rpb), 256, 0, st>>>(l, c, colsum, rpb);
    note_launch();
    nce_loss_kernel<<<1, 1024, 0, st>>>(l, c, rowsum, colsum, loss);
    note_launch();
    SCB_LAUNCH_OK("infonce_fwd");
  }
  if ((phase & 2) && (dA || dB || dlog_mult)) {
    // consumes the logits (turned into dloss/dlogits in place): a second backward needs a new forward
    nce_grad_kernel<<<eb, 256, 0, st>>>(l, c, rowsum, colsum, upstream, upstream_dev, dlog_mult);
    note_launch();
    SCB_LAUNCH_OK("infonce_grad");
    unsigned sb = (unsigned)(((long long)B * D + 255) / 256);
    if (sb > 8u * num_sms()) sb = 8u * num_sms();
    const bool tens = nce_tensor_path(B, D);
    float* gt = colsum + B + 64;            // tensor path: G^T, then B^T / A^T ([D][B])
    float* ft = gt + (long long)B * B;
    if (dA) {  // dA = mult * G B
      if (tens) {
        e = transpose(feat_b, SCB_F32, D, ft, SCB_F32, B, B, D, st);
        if (!e) e = tf32_gemm(l, B, ft, B, dA, D, B, D, B, st);
      } else {
        e = sgemm(l, B, 1, feat_b, 1, D, dA, D, B, D, B, 1.f, 0.f, st);
      }
      if (e) return e;
      scale_by_mult_kernel<<<sb, 256, 0, st>>>(dA, (long long)B * D, log_mult, fixed_mult);
      note_launch();
    }
    if (dB) {  // dB = mult * G^T A
      if (tens) {
        e = transpose(l, SCB_F32, B, gt, SCB_F32, B, B, B, st);
        if (!e) e = transpose(feat_a, SCB_F32, D, ft + (long long)B * D, SCB_F32, B, B, D, st);
        if (!e) e = tf32_gemm(gt, B, ft + (long long)B * D, B, dB, D, B, D, B, st);
      } else {
        e = sgemm(l, 1, B, feat_a, 1, D, dB, D, B, D, B, 1.f, 0.f, st);
      }
      if (e) return e;
      scale_by_mult_kernel<<<sb, 256, 0, st>>>(dB, (long long)B * D, log_mult, fixed_mult);
      note_launch();
    }
    SCB_LAUNCH_OK("infonce_bwd");
  }
  return SCB_OK;
}

}  // namespace scb
