// Optimizer step for the trainable head: global-norm gradient clipping (Lightning gradient_clip_val, config
// spchclp_p.yaml:108) fused with torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected
// moments; kwClip.py:666-694) over ONE flat fp32 buffer, plus refresh of the 16-bit weight copies the GEMMs read.
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)g[i];
    s += v * v;
  }
  s = warp_sum_d(s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
    atomicAdd(out, v);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const double* __restrict__ sumsq, float grad_scale,
                                                   float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                                                   float bc1, float bc2, __half* __restrict__ p_f16, __nv_bfloat16* __restrict__ p_bf16) {
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = (float)sqrt(*sumsq) * grad_scale;  // norm of the unscaled gradient
    const float c = max_norm / (total + 1e-6f);
    coef *= fminf(c, 1.0f);
  }
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * coef + weight_decay * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi;
    if (p_f16) p_f16[i] = __float2half_rn(pi);
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

}  // namespace

int adam_step(float* p, const float* g, float* m, float* v, long long n, double* sumsq_scratch, float grad_scale, float max_norm, float lr,
              float beta1, float beta2, float eps, float weight_decay, int step, void* p_f16, void* p_bf16, cudaStream_t st) {
  SCB_CHECK(p && g && m && v && sumsq_scratch, SCB_EINVAL, "scb_adam_step: null operand");
  SCB_CHECK(step >= 1, SCB_EINVAL, "scb_adam_step: step counts from 1");
  if (n == 0) return SCB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  SCB_CUDA(cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), st));
  if (max_norm > 0.f) {
    sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n, sumsq_scratch);
    note_launch();
  }
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, sumsq_scratch, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, bc1,
                                               bc2, (__half*)p_f16, (__nv_bfloat16*)p_bf16);
  note_launch();
  SCB_LAUNCH_OK("adam_step");
  return SCB_OK;
}

}  // namespace scb
