// HBM-bound row kernels: LayerNorm fwd/bwd, L2-normalise fwd/bwd, layer weighted-sum fwd/bwd.
// One warp per row, 128-bit accesses, statistics in fp32 registers (two-pass on the cached row).
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

constexpr int kMaxVec = 8;  // per-lane float4 slots: rows up to 32*4*8 = 1024 elements stay in registers

template <typename TIn>
__device__ __forceinline__ float4 load4(const TIn* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = H16<SCB_F16>::unpack(u.x), b = H16<SCB_F16>::unpack(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = H16<SCB_BF16>::unpack(u.x), b = H16<SCB_BF16>::unpack(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4_16(void* p, int fmt, float4 v) {
  uint2 u;
  u.x = pack16(fmt, v.x, v.y);
  u.y = pack16(fmt, v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// y = (x - mean) * rstd * gamma + beta ; optional fp32 and 16-bit outputs; optional saved mean/rstd (for backward).
template <typename TIn>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y32,
                                                            void* __restrict__ y16, int y16_fmt, float* __restrict__ stats,
                                                            long long rows, int d, long long x_ld, long long y_ld, float eps, int act) {
  griddep_wait();               // programmatic dependent launch (see launch_pdl)
  griddep_launch_dependents();  // the next kernel (a GEMM) may set up its barriers / TMEM while the last wave of rows runs
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TIn* xr = x + row * x_ld;
  const int nvec = d >> 2;
  float4 v[kMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = load4<TIn>(xr + c * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      ss += a * a + b * b + cc * cc + dd * dd;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / d + eps);
  if (stats && lane == 0) {
    stats[row * 2] = mean;
    stats[row * 2 + 1] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      if (gamma) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w;
      }
      if (beta) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (act == SCB_ACT_GELU_ERF) {
        if (y32) {  // fp32 consumers get the erff form; 16-bit-only outputs (HuBERT-large conv blocks) the fit of gelu_h16
          o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
        } else {
          upk2(gelu_h16_x2(pk2(o.x, o.y)), o.x, o.y);
          upk2(gelu_h16_x2(pk2(o.z, o.w)), o.z, o.w);
        }
      } else if (act == SCB_ACT_QUICK_GELU) {
        o.x = quick_gelu(o.x); o.y = quick_gelu(o.y); o.z = quick_gelu(o.z); o.w = quick_gelu(o.w);
      }
      if (y32) *reinterpret_cast<float4*>(y32 + row * y_ld + c * 4) = o;
      if (y16) store4_16(reinterpret_cast<uint16_t*>(y16) + row * y_ld + c * 4, y16_fmt, o);
    }
  }
}

// 16-bit rows of up to 512 elements, 16-bit output only (the HuBERT-large conv blocks: LayerNorm over 512 channels + GELU, in
// place on B x T x 512 activations — 5.2 M rows per 256 utterances).  The generic kernel above moves 8 bytes per lane per load,
// re-reads gamma / beta for every row and runs one row per warp: 0.44 of the copy bandwidth.  Here a warp walks rows with a grid
// stride, keeps its gamma / beta slice in registers, and moves 16 bytes per lane per access.
template <int FMT>
__global__ void __launch_bounds__(256) layernorm16_rows_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, uint16_t* __restrict__ y, long long rows, int d,
                                                               long long x_ld, long long y_ld, float eps, int act) {
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int nv = d >> 3;   // 16-byte vectors per row (<= 64)
  const bool has1 = lane + 32 < nv, has0 = lane < nv;
  float g[2][8], b[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + i * 32) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = i == 0 ? has0 : has1;
      g[i][j] = (ok && gamma) ? gamma[c + j] : 1.f;
      b[i][j] = (ok && beta) ? beta[c + j] : 0.f;
    }
  }
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const uint16_t* xr = x + row * x_ld;
    uint4 u[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
    if (has0) u[0] = *reinterpret_cast<const uint4*>(xr + lane * 8);
    if (has1) u[1] = *reinterpret_cast<const uint4*>(xr + (lane + 32) * 8);
    float v[2][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = H16<FMT>::unpack(w[j]);
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        s += f.x + f.y;
      }
    }
    const float mean = warp_sum(s) / d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i == 0 ? has0 : has1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += (v[i][j] - mean) * (v[i][j] - mean);
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / d + eps);
    uint16_t* yr = y + row * y_ld;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i == 0 ? has0 : has1) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o0 = fmaf((v[i][2 * j] - mean) * rstd, g[i][2 * j], b[i][2 * j]);
          float o1 = fmaf((v[i][2 * j + 1] - mean) * rstd, g[i][2 * j + 1], b[i][2 * j + 1]);
          if (act == SCB_ACT_GELU_ERF) upk2(gelu_h16_x2(pk2(o0, o1)), o0, o1);
          else if (act == SCB_ACT_QUICK_GELU) upk2(quick_gelu_x2(pk2(o0, o1)), o0, o1);
          w[j] = H16<FMT>::pack(o0, o1);
        }
        *reinterpret_cast<uint4*>(yr + (lane + i * 32) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// LayerNorm backward for fp32 rows: dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));
// dgamma/dbeta accumulated per block in shared memory, then one atomicAdd per column per block.
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            float* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, long long rows, int d) {
  extern __shared__ float sacc[];  // [2][d]
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float4 xh[kMaxVec], g[kMaxVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 xv = *reinterpret_cast<const float4*>(x + row * d + c * 4);
        const float4 dv = *reinterpret_cast<const float4*>(dy + row * d + c * 4);
        const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        if (dgamma) {
          atomicAdd(&sacc[c * 4 + 0], dv.x * xh[i].x); atomicAdd(&sacc[c * 4 + 1], dv.y * xh[i].y);
          atomicAdd(&sacc[c * 4 + 2], dv.z * xh[i].z); atomicAdd(&sacc[c * 4 + 3], dv.w * xh[i].w);
          atomicAdd(&sacc[d + c * 4 + 0], dv.x); atomicAdd(&sacc[d + c * 4 + 1], dv.y);
          atomicAdd(&sacc[d + c * 4 + 2], dv.z); atomicAdd(&sacc[d + c * 4 + 3], dv.w);
        }
        g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (g[i].x - s1 - xh[i].x * s2);
        o.y = rstd * (g[i].y - s1 - xh[i].y * s2);
        o.z = rstd * (g[i].z - s1 - xh[i].z * s2);
        o.w = rstd * (g[i].w - s1 - xh[i].w * s2);
        *reinterpret_cast<float4*>(dx + row * d + c * 4) = o;
      }
    }
  }
  __syncthreads();
  if (dgamma)
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(&dgamma[i], sacc[i]);
      atomicAdd(&dbeta[i], sacc[d + i]);
    }
}

// y = x / ||x||_2 per row (kwClip.py:1436,1451-1453).  norms saved for backward.
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ norms, int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[(long long)row * d + c];
    ss += v * v;
  }
  const float nrm = sqrtf(warp_sum(ss));
  if (norms && lane == 0) norms[row] = nrm;
  const float inv = 1.0f / nrm;
  for (int c = lane; c < d; c += 32) y[(long long)row * d + c] = x[(long long)row * d + c] * inv;
}

// dx = (dy - y * <y, dy>) / ||x||
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                         const float* __restrict__ norms, float* __restrict__ dx, int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < d; c += 32) dot += y[(long long)row * d + c] * dy[(long long)row * d + c];
  dot = warp_sum(dot);
  const float inv = 1.0f / norms[row];
  for (int c = lane; c < d; c += 32)
    dx[(long long)row * d + c] = (dy[(long long)row * d + c] - y[(long long)row * d + c] * dot) * inv;
}

// Hidden-state rows are read in 16-byte units whatever their type: 4 fp32 or 8 fp16 elements per lane per load (with 8-byte
// fp16 loads the kernels had half the bytes in flight and ran at 3 TB/s instead of the 6.3 TB/s of the fp32 read).
template <typename TH> struct HVec;
template <> struct HVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
};
template <> struct HVec<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = H16<SCB_F16>::unpack(u.x), b = H16<SCB_F16>::unpack(u.y), c = H16<SCB_F16>::unpack(u.z), d = H16<SCB_F16>::unpack(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
};
constexpr int kMaxElems = kMaxVec * 4;  // elements of a row per lane

// out = sum_l softmax(w)_l * h_l   (weighted_sum.py:38-43), h = [L][rows][d] (fp32, or the fp16 hidden states the post-LN
// tower keeps: half the bytes of the 13-layer read) with layer stride.
// NORMALIZE: parameter-free LayerNorm over d on every h_l first (weighted_sum.py:41-42).  One warp per row.
template <bool NORMALIZE, typename TH>
__global__ void __launch_bounds__(256) weighted_sum_fwd_kernel(const TH* __restrict__ h, long long layer_stride,
                                                               const float* __restrict__ w_logits, int L,
                                                               float* __restrict__ out32, void* __restrict__ out16, int out16_fmt,
                                                               long long rows, int d, int rows_per_batch,
                                                               long long out16_batch_stride, long long out16_row0) {
  constexpr int V = HVec<TH>::N, NV = kMaxElems / V;
  __shared__ float sw[64];
  if (threadIdx.x < 32) {
    float v = threadIdx.x < L ? w_logits[threadIdx.x] : -INFINITY;
    float v2 = threadIdx.x + 32 < L ? w_logits[threadIdx.x + 32] : -INFINITY;
    const float m = warp_max(fmaxf(v, v2));
    const float e = threadIdx.x < L ? __expf(v - m) : 0.f, e2 = threadIdx.x + 32 < L ? __expf(v2 - m) : 0.f;
    const float s = warp_sum(e + e2);
    sw[threadIdx.x] = e / s;
    sw[threadIdx.x + 32] = e2 / s;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = d / V;
  float acc[NV][V];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < V; ++j) acc[i][j] = 0.f;
#pragma unroll 2  // two layers' loads in flight per lane: ~50 KB per SM, what the HBM latency needs
  for (int l = 0; l < L; ++l) {
    const TH* hr = h + (long long)l * layer_stride + row * d;
    const float wl = sw[l];
    float v[NV][V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        HVec<TH>::load(hr + c * V, v[i]);
        if (NORMALIZE) {
#pragma unroll
          for (int j = 0; j < V; ++j) s += v[i][j];
        }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (NORMALIZE) {
      mean = warp_sum(s) / d;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
#pragma unroll
          for (int j = 0; j < V; ++j) ss += (v[i][j] - mean) * (v[i][j] - mean);
        }
      }
      rstd = rsqrtf(warp_sum(ss) / d + 1e-5f);
    }
    const float a = wl * rstd, b = -wl * rstd * mean;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < V; ++j) acc[i][j] += a * v[i][j] + b;
      }
    }
  }
  const long long bidx = row / rows_per_batch, r = row % rows_per_batch;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const float4 o = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        if (out32) *reinterpret_cast<float4*>(out32 + row * d + c * V + j) = o;
        if (out16)
          store4_16(reinterpret_cast<uint16_t*>(out16) + bidx * out16_batch_stride + (out16_row0 + r) * d + c * V + j, out16_fmt, o);
      }
    }
  }
}

// ---- un-normalised weighted sum (the base configurations): one thread per 16-byte chunk of a row.
// The row-per-warp kernels above keep a whole row in registers because the parameter-free LayerNorm needs its statistics; without
// it nothing couples the columns, and the kernel is a pure 13-stream read.  What bounds it then is bytes in flight: the
// row-per-warp form (80 registers, 24 warps per SM, 2 layers unrolled) reached 3.6 TB/s.  Here every thread owns ONE 16-byte
// column chunk, issues kFlatUnroll independent 16-byte loads (one per layer) before it touches any of them, and keeps only
// 4 / 8 accumulators, so ~32 warps per SM hold 128 B each in flight.
constexpr int kFlatUnroll = 8;
template <typename TH> __device__ __forceinline__ void decode16(const uint4& u, float (&v)[HVec<TH>::N]);
template <> __device__ __forceinline__ void decode16<float>(const uint4& u, float (&v)[4]) {
  v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
}
template <> __device__ __forceinline__ void decode16<__half>(const uint4& u, float (&v)[8]) {
  const float2 a = H16<SCB_F16>::unpack(u.x), b = H16<SCB_F16>::unpack(u.y), c = H16<SCB_F16>::unpack(u.z), d = H16<SCB_F16>::unpack(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 ld_stream16(const void* p) {  // read-once data: non-coherent path, no L1 allocation
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}

template <typename TH>
__global__ void __launch_bounds__(256) weighted_sum_fwd_flat_kernel(const TH* __restrict__ h, long long layer_stride,
                                                                    const float* __restrict__ w_logits, int L,
                                                                    float* __restrict__ out32, void* __restrict__ out16, int out16_fmt,
                                                                    long long rows, int d, int rows_per_batch,
                                                                    long long out16_batch_stride, long long out16_row0) {
  constexpr int V = HVec<TH>::N;
  __shared__ float sw[64];
  if (threadIdx.x < 32) {
    float v = threadIdx.x < L ? w_logits[threadIdx.x] : -INFINITY;
    float v2 = threadIdx.x + 32 < L ? w_logits[threadIdx.x + 32] : -INFINITY;
    const float m = warp_max(fmaxf(v, v2));
    const float e = threadIdx.x < L ? __expf(v - m) : 0.f, e2 = threadIdx.x + 32 < L ? __expf(v2 - m) : 0.f;
    const float s = warp_sum(e + e2);
    sw[threadIdx.x] = e / s;
    sw[threadIdx.x + 32] = e2 / s;
  }
  __syncthreads();
  const int nvec = d / V;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int c = (int)(idx - row * nvec);
  const TH* p = h + row * d + c * V;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  for (int l0 = 0; l0 < L; l0 += kFlatUnroll) {
    uint4 raw[kFlatUnroll];
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u)
      if (l0 + u < L) raw[u] = ld_stream16(p + (long long)(l0 + u) * layer_stride);
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) {
      if (l0 + u < L) {
        float v[V];
        decode16<TH>(raw[u], v);
        const float wl = sw[l0 + u];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(wl, v[j], acc[j]);
      }
    }
  }
  const long long bidx = row / rows_per_batch, r = row % rows_per_batch;
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    if (out32) *reinterpret_cast<float4*>(out32 + row * d + c * V + j) = o;
    if (out16) store4_16(reinterpret_cast<uint16_t*>(out16) + bidx * out16_batch_stride + (out16_row0 + r) * d + c * V + j, out16_fmt, o);
  }
}

template <typename TH, int LMAX>
__global__ void __launch_bounds__(256, LMAX <= 16 ? 3 : 2) weighted_sum_bwd_flat_kernel(const TH* __restrict__ h, long long layer_stride, int L,
                                                                       const float* __restrict__ dout, long long rows, int d,
                                                                       int rows_per_batch, long long dout_batch_stride, long long dout_row0,
                                                                       float* __restrict__ dw_raw) {
  constexpr int V = HVec<TH>::N;
  __shared__ float sacc[64];
  if (threadIdx.x < 64) sacc[threadIdx.x] = 0.f;
  __syncthreads();
  const int nvec = d / V;
  const long long total = rows * nvec;
  float part[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) part[l] = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / nvec;
    const int c = (int)(idx - row * nvec);
    const long long bidx = row / rows_per_batch, r = row % rows_per_batch;
    const float* dr = dout + bidx * dout_batch_stride + (dout_row0 + r) * d + c * V;
    float g[V];
#pragma unroll
    for (int j = 0; j < V; j += 4) {
      const uint4 f = ld_stream16(dr + j);
      g[j] = __uint_as_float(f.x); g[j + 1] = __uint_as_float(f.y); g[j + 2] = __uint_as_float(f.z); g[j + 3] = __uint_as_float(f.w);
    }
    const TH* p = h + row * d + c * V;
#pragma unroll
    for (int l0 = 0; l0 < LMAX; l0 += kFlatUnroll) {
      uint4 raw[kFlatUnroll];
#pragma unroll
      for (int u = 0; u < kFlatUnroll; ++u)
        if (l0 + u < L) raw[u] = ld_stream16(p + (long long)(l0 + u) * layer_stride);
#pragma unroll
      for (int u = 0; u < kFlatUnroll; ++u) {
        if (l0 + u < L) {
          float v[V];
          decode16<TH>(raw[u], v);
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < V; ++j) dot = fmaf(g[j], v[j], dot);
          part[l0 + u] += dot;
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      const float s = warp_sum(part[l]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[l], s);
    }
  }
  __syncthreads();
  if (threadIdx.x < L) atomicAdd(&dw_raw[threadIdx.x], sacc[threadIdx.x]);
}

// dw_l = sum_{r,c} dout[r,c] * h_l[r,c]  (h optionally LayerNorm'ed first); then softmax backward into dlogits.
// dout rows may live inside a larger per-batch buffer (branch input gradient): row r of batch b is at
// dout + b*dout_batch_stride + (dout_row0 + r)*d.
template <bool NORMALIZE, typename TH, int LMAX>
__global__ void __launch_bounds__(256, 2) weighted_sum_bwd_kernel(const TH* __restrict__ h, long long layer_stride, int L,
                                                               const float* __restrict__ dout, long long rows, int d,
                                                               int rows_per_batch, long long dout_batch_stride, long long dout_row0,
                                                               float* __restrict__ dw_raw) {
  constexpr int V = HVec<TH>::N, NV = kMaxElems / V;
  __shared__ float sacc[64];
  if (threadIdx.x < 64) sacc[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = d / V;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  float part[LMAX];  // L <= LMAX (16 or 32) handled in registers per warp; larger L unsupported (checked on host)
#pragma unroll
  for (int l = 0; l < LMAX; ++l) part[l] = 0.f;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const long long bidx = row / rows_per_batch, r = row % rows_per_batch;
    const float* dr = dout + bidx * dout_batch_stride + (dout_row0 + r) * d;
    float g[NV][V];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < V; j += 4) {
          const float4 f = *reinterpret_cast<const float4*>(dr + c * V + j);
          g[i][j] = f.x; g[i][j + 1] = f.y; g[i][j + 2] = f.z; g[i][j + 3] = f.w;
        }
      }
    }
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < L) {
        const TH* hr = h + (long long)l * layer_stride + row * d;
        float v[NV][V];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) {
            HVec<TH>::load(hr + c * V, v[i]);
            if (NORMALIZE) {
#pragma unroll
              for (int j = 0; j < V; ++j) s += v[i][j];
            }
          }
        }
        float mean = 0.f, rstd = 1.f;
        if (NORMALIZE) {
          mean = warp_sum(s) / d;
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
#pragma unroll
              for (int j = 0; j < V; ++j) ss += (v[i][j] - mean) * (v[i][j] - mean);
            }
          }
          rstd = rsqrtf(warp_sum(ss) / d + 1e-5f);
        }
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) {
#pragma unroll
            for (int j = 0; j < V; ++j) dot += g[i][j] * (v[i][j] - mean);
          }
        }
        part[l] += dot * rstd;
      }
    }
  }
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      const float s = warp_sum(part[l]);
      if (lane == 0) atomicAdd(&sacc[l], s);
    }
  }
  __syncthreads();
  if (threadIdx.x < L) atomicAdd(&dw_raw[threadIdx.x], sacc[threadIdx.x]);
}

// dlogit_l = w_l * (dw_l - sum_k w_k dw_k), accumulated into grad (+=).
__global__ void softmax_bwd_small_kernel(const float* __restrict__ logits, const float* __restrict__ dw, float* __restrict__ grad,
                                         int L, float scale) {
  __shared__ float w[64];
  __shared__ float dot;
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int i = 0; i < L; ++i) m = fmaxf(m, logits[i]);
    float s = 0.f;
    for (int i = 0; i < L; ++i) {
      w[i] = __expf(logits[i] - m);
      s += w[i];
    }
    float dd = 0.f;
    for (int i = 0; i < L; ++i) {
      w[i] /= s;
      dd += w[i] * dw[i];
    }
    dot = dd;
  }
  __syncthreads();
  if (threadIdx.x < L) grad[threadIdx.x] += scale * w[threadIdx.x] * (dw[threadIdx.x] - dot);
}

}  // namespace

int layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float* y32, void* y16, int y16_fmt,
                  float* stats, long long rows, int d, long long x_ld, long long y_ld, float eps, int act, cudaStream_t st) {
  SCB_CHECK(x && (y32 || y16) && rows >= 0, SCB_EINVAL, "scb_layernorm: null operand");
  SCB_CHECK(d % 4 == 0 && d <= 32 * 4 * kMaxVec, SCB_EUNSUPPORTED, "scb_layernorm: d=%d must be a multiple of 4 and <= %d", d, 32 * 4 * kMaxVec);
  SCB_CHECK(x_ld % 4 == 0 && y_ld % 4 == 0, SCB_EINVAL, "scb_layernorm: leading dimensions must be multiples of 4");
  if (rows == 0) return SCB_OK;
  const int wpb = 8;
  if (x_dtype != SCB_F32 && !y32 && y16 && y16_fmt == x_dtype && !stats && d % 8 == 0 && d <= 512 && x_ld % 8 == 0 && y_ld % 8 == 0 && rows >= 4096 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y16)) & 15) == 0) {
    long long blocks = (rows + wpb - 1) / wpb;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    if (x_dtype == SCB_F16)
      SCB_CUDA(launch_pdl(layernorm16_rows_kernel<SCB_F16>, dim3((unsigned)blocks), wpb * 32, 0, st, (const uint16_t*)x, gamma, beta, (uint16_t*)y16, rows, d, x_ld, y_ld, eps, act));
    else
      SCB_CUDA(launch_pdl(layernorm16_rows_kernel<SCB_BF16>, dim3((unsigned)blocks), wpb * 32, 0, st, (const uint16_t*)x, gamma, beta, (uint16_t*)y16, rows, d, x_ld, y_ld, eps, act));
    note_launch();
    SCB_LAUNCH_OK("layernorm16_rows");
    return SCB_OK;
  }
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (x_dtype == SCB_F32)
    SCB_CUDA(launch_pdl(layernorm_fwd_kernel<float>, dim3(grid), wpb * 32, 0, st, (const float*)x, gamma, beta, y32, y16, y16_fmt, stats, rows, d, x_ld, y_ld, eps, act));
  else if (x_dtype == SCB_F16)
    SCB_CUDA(launch_pdl(layernorm_fwd_kernel<__half>, dim3(grid), wpb * 32, 0, st, (const __half*)x, gamma, beta, y32, y16, y16_fmt, stats, rows, d, x_ld, y_ld, eps, act));
  else
    SCB_CUDA(launch_pdl(layernorm_fwd_kernel<__nv_bfloat16>, dim3(grid), wpb * 32, 0, st, (const __nv_bfloat16*)x, gamma, beta, y32, y16, y16_fmt, stats, rows, d, x_ld, y_ld, eps, act));
  note_launch();
  SCB_LAUNCH_OK("layernorm_fwd");
  return SCB_OK;
}

int layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                  long long rows, int d, cudaStream_t st) {
  SCB_CHECK(dy && x && stats && dx, SCB_EINVAL, "scb_layernorm_bwd: null operand");
  SCB_CHECK((dgamma == nullptr) == (dbeta == nullptr), SCB_EINVAL, "scb_layernorm_bwd: dgamma and dbeta go together");
  SCB_CHECK(d % 4 == 0 && d <= 32 * 4 * kMaxVec, SCB_EUNSUPPORTED, "scb_layernorm_bwd: unsupported d=%d", d);
  if (rows == 0) return SCB_OK;
  const int wpb = 8;
  long long blocks = (rows + wpb - 1) / wpb;
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  layernorm_bwd_kernel<<<(unsigned)blocks, wpb * 32, 2 * d * sizeof(float), st>>>(dy, x, stats, gamma, dx, dgamma, dbeta, rows, d);
  note_launch();
  SCB_LAUNCH_OK("layernorm_bwd");
  return SCB_OK;
}

int l2norm_fwd(const float* x, float* y, float* norms, int rows, int d, cudaStream_t st) {
  SCB_CHECK(x && y, SCB_EINVAL, "scb_l2norm: null operand");
  if (rows == 0) return SCB_OK;
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, y, norms, rows, d);
  note_launch();
  SCB_LAUNCH_OK("l2norm_fwd");
  return SCB_OK;
}

int l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int d, cudaStream_t st) {
  SCB_CHECK(dy && y && norms && dx, SCB_EINVAL, "scb_l2norm_bwd: null operand");
  if (rows == 0) return SCB_OK;
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(dy, y, norms, dx, rows, d);
  note_launch();
  SCB_LAUNCH_OK("l2norm_bwd");
  return SCB_OK;
}

int weighted_sum_fwd(const void* h, int h_dtype, long long layer_stride, const float* w_logits, int L, int normalize, float* out32, void* out16,
                     int out16_fmt, long long rows, int d, int rows_per_batch, long long out16_batch_stride, long long out16_row0,
                     cudaStream_t st) {
  SCB_CHECK(h && w_logits && (out32 || out16), SCB_EINVAL, "scb_weighted_sum: null operand");
  SCB_CHECK(h_dtype == SCB_F32 || h_dtype == SCB_F16, SCB_EUNSUPPORTED, "scb_weighted_sum: hidden states must be fp32 or fp16");
  SCB_CHECK(L >= 1 && L <= 64, SCB_EUNSUPPORTED, "scb_weighted_sum: L=%d out of range [1,64]", L);
  SCB_CHECK(d % (h_dtype == SCB_F16 ? 8 : 4) == 0 && d <= 32 * 4 * kMaxVec, SCB_EUNSUPPORTED, "scb_weighted_sum: unsupported d=%d", d);
  if (rows == 0) return SCB_OK;
  if (rows_per_batch <= 0) rows_per_batch = (int)rows;
  const unsigned grid = (unsigned)((rows + 7) / 8);
#define SCB_WS_FWD(NORM, TH) \
  weighted_sum_fwd_kernel<NORM, TH><<<grid, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, w_logits, L, out32, out16, out16_fmt, \
                                                          rows, d, rows_per_batch, out16_batch_stride, out16_row0)
  const long long chunks = rows * (d / (h_dtype == SCB_F16 ? 8 : 4));
  const unsigned fgrid = (unsigned)((chunks + 255) / 256);
#define SCB_WS_FLAT(TH) \
  weighted_sum_fwd_flat_kernel<TH><<<fgrid, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, w_logits, L, out32, out16, out16_fmt, \
                                                          rows, d, rows_per_batch, out16_batch_stride, out16_row0)
  if (h_dtype == SCB_F16) {
    if (normalize) SCB_WS_FWD(true, __half); else SCB_WS_FLAT(__half);
  } else {
    if (normalize) SCB_WS_FWD(true, float); else SCB_WS_FLAT(float);
  }
#undef SCB_WS_FLAT
#undef SCB_WS_FWD
  note_launch();
  SCB_LAUNCH_OK("weighted_sum_fwd");
  return SCB_OK;
}

int weighted_sum_bwd(const void* h, int h_dtype, long long layer_stride, const float* w_logits, int L, int normalize, const float* dout,
                     long long rows, int d, int rows_per_batch, long long dout_batch_stride, long long dout_row0, float* scratch_L,
                     float* grad_logits, float grad_scale, cudaStream_t st) {
  SCB_CHECK(h && w_logits && dout && scratch_L && grad_logits, SCB_EINVAL, "scb_weighted_sum_bwd: null operand");
  SCB_CHECK(h_dtype == SCB_F32 || h_dtype == SCB_F16, SCB_EUNSUPPORTED, "scb_weighted_sum_bwd: hidden states must be fp32 or fp16");
  SCB_CHECK(L >= 1 && L <= 32, SCB_EUNSUPPORTED, "scb_weighted_sum_bwd: L=%d out of range [1,32]", L);
  SCB_CHECK(d % (h_dtype == SCB_F16 ? 8 : 4) == 0 && d <= 32 * 4 * kMaxVec, SCB_EUNSUPPORTED, "scb_weighted_sum_bwd: unsupported d=%d", d);
  if (rows_per_batch <= 0) rows_per_batch = (int)rows;
  SCB_CUDA(cudaMemsetAsync(scratch_L, 0, L * sizeof(float), st));
  if (rows > 0) {
    long long blocks = (rows + 7) / 8;
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
#define SCB_WS_BWD(NORM, TH)                                                                                                       \
  do {                                                                                                                             \
    if (L <= 16)                                                                                                                   \
      weighted_sum_bwd_kernel<NORM, TH, 16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, L, dout, rows, d, \
                                                                              rows_per_batch, dout_batch_stride, dout_row0, scratch_L); \
    else                                                                                                                           \
      weighted_sum_bwd_kernel<NORM, TH, 32><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, L, dout, rows, d, \
                                                                              rows_per_batch, dout_batch_stride, dout_row0, scratch_L); \
  } while (0)
    const long long chunks = rows * (d / (h_dtype == SCB_F16 ? 8 : 4));
    long long fblocks = (chunks + 255) / 256;
    const long long per_sm = L <= 16 ? 3 : 2;  // resident blocks (see the kernel's launch bounds)
    if (fblocks > per_sm * num_sms()) fblocks = per_sm * num_sms();
#define SCB_WS_BWD_FLAT(TH)                                                                                                        \
  do {                                                                                                                             \
    if (L <= 16)                                                                                                                   \
      weighted_sum_bwd_flat_kernel<TH, 16><<<(unsigned)fblocks, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, L, dout, rows, d, \
                                                                              rows_per_batch, dout_batch_stride, dout_row0, scratch_L); \
    else                                                                                                                           \
      weighted_sum_bwd_flat_kernel<TH, 32><<<(unsigned)fblocks, 256, 0, st>>>(static_cast<const TH*>(h), layer_stride, L, dout, rows, d, \
                                                                              rows_per_batch, dout_batch_stride, dout_row0, scratch_L); \
  } while (0)
    if (h_dtype == SCB_F16) {
      if (normalize) SCB_WS_BWD(true, __half); else SCB_WS_BWD_FLAT(__half);
    } else {
      if (normalize) SCB_WS_BWD(true, float); else SCB_WS_BWD_FLAT(float);
    }
#undef SCB_WS_BWD_FLAT
#undef SCB_WS_BWD
    note_launch();
    SCB_LAUNCH_OK("weighted_sum_bwd");
  }
  softmax_bwd_small_kernel<<<1, 64, 0, st>>>(w_logits, scratch_L, grad_logits, L, grad_scale);
  note_launch();
  SCB_LAUNCH_OK("softmax_bwd_small");
  return SCB_OK;
}

}  // namespace scb
