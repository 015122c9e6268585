// Front-end kernels (HBM / CUDA-core bound):
//   * HuBERT conv0 (C_in = 1, k = 10, s = 5) + GroupNorm-over-time + GELU   [fairseq ConvFeatureExtractionModel layer 0]
//       - statistics from the 10x10 windowed autocorrelation of the waveform (no second pass over 512 channels)
//       - apply pass recomputes the conv and writes channel-last f16 activations for the conv1 GEMM
//   * pos-conv input packing (zero padded frames, regroup 16 x 48 -> 16 x 64 channels, zero time padding)
//   * CLIP patchify (im2col for kernel = stride = P) and row broadcast helpers
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

constexpr int kTaps = 10;
constexpr int kStride = 5;
constexpr int kPairs = kTaps * (kTaps + 1) / 2;  // 55 upper-triangular autocorrelation entries
constexpr int kStatVals = kPairs + kTaps;        // + 10 plain sums

// acc[b][0..54] = sum_t x[5t+k] x[5t+k'] (k<=k'), acc[b][55..64] = sum_t x[5t+k]   over all n_frames of utterance b.
__global__ void __launch_bounds__(256) conv0_autocorr_kernel(const float* __restrict__ wav, long long wav_ld, int n_frames,
                                                             double* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* x = wav + (long long)b * wav_ld;
  float part[kStatVals];
#pragma unroll
  for (int i = 0; i < kStatVals; ++i) part[i] = 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_frames; t += gridDim.x * blockDim.x) {
    float v[kTaps];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) v[k] = __ldg(x + (long long)t * kStride + k);
    int idx = 0;
#pragma unroll
    for (int k = 0; k < kTaps; ++k)
#pragma unroll
      for (int k2 = k; k2 < kTaps; ++k2) part[idx++] += v[k] * v[k2];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) part[kPairs + k] += v[k];
  }
  __shared__ double red[8][kStatVals];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kStatVals; ++i) {
    const double s = warp_sum_d((double)part[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < kStatVals) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    atomicAdd(&acc[(long long)b * kStatVals + threadIdx.x], s);
  }
}

// Per (utterance, channel): mean / biased variance of the conv output from the autocorrelation, folded with the
// GroupNorm affine into y = conv * scale + shift.
__global__ void conv0_finalize_kernel(const double* __restrict__ acc, const float* __restrict__ w, const float* __restrict__ conv_bias,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, int n_frames, int channels,
                                      float eps, float2* __restrict__ scale_shift) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  const double* a = acc + (long long)b * kStatVals;
  double wk[kTaps];
#pragma unroll
  for (int k = 0; k < kTaps; ++k) wk[k] = (double)w[c * kTaps + k];
  double s1 = 0.0, s2 = 0.0;
  int idx = 0;
#pragma unroll
  for (int k = 0; k < kTaps; ++k) {
    s1 += wk[k] * a[kPairs + k];
#pragma unroll
    for (int k2 = k; k2 < kTaps; ++k2) {
      const double r = a[idx++];
      s2 += (k == k2 ? 1.0 : 2.0) * wk[k] * wk[k2] * r;
    }
  }
  const double mean_nb = s1 / n_frames;
  double var = s2 / n_frames - mean_nb * mean_nb;
  if (var < 0.0) var = 0.0;
  const double cb = conv_bias ? (double)conv_bias[c] : 0.0;
  const double mean = mean_nb + cb;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
  // y = ((conv_nb + cb) - mean) * rstd * g + be  with conv_nb the bias-free conv
  scale_shift[(long long)b * channels + c] = make_float2((float)(rstd * g), (float)(be - mean_nb * rstd * g));
}

// out[b, t, c] = gelu(conv0(wav)[b, c, t] * scale[b,c] + shift[b,c])   channel-last, 16-bit.
// The 10-tap convolution runs on the tensor cores (mma.sync m16n8k16, fp16 operands like the reference's autocast conv1d,
// fp32 accumulate): A = 16 frames x 16 "taps" (10 waveform samples, then two constant-1 columns, then zeros), B = the
// channel's 10 weights pre-multiplied by the GroupNorm scale plus the GroupNorm shift split into an fp16 hi/lo pair on the two
// constant columns — so the accumulator IS scale*conv + shift and the epilogue is only GELU + pack + store.  That leaves
// ~16 instructions per output element instead of ~32 on the fp32 SIMT path (the kernel evaluates 2.7 G activations per step).
// Block = 8 warps = 2 frame halves x 4 channel groups of 128; a warp keeps its 16 B-fragments (128 channels) in registers
// and walks frame tiles of 16.  The fragment column -> channel map is chosen so that a store instruction writes whole 32-byte
// sectors: column j of n-block n is channel (j/4)*64 + (n/4)*16 + ((j/2)%2)*8 + (n%4)*2 + (j%2), so after four n-blocks a lane
// holds 8 contiguous channels (16 bytes) and its neighbour (lane ^ 1) the next 8.
constexpr int kFramesPerBlock = 128;
template <int FMT>
__device__ __forceinline__ uint32_t gelu_pack(float a, float b) {  // two activations on the packed-fp32 path, then one 16-bit pair
  float lo, hi;
  upk2(gelu_h16_x2(pk2(a, b)), lo, hi);
  return H16<FMT>::pack(lo, hi);
}
template <int FPB>  // frames per block: the B fragments (weights x GroupNorm scale, 2 dependent global reads) are built once per block
__global__ void __launch_bounds__(256) conv0_apply_kernel(const float* __restrict__ wav, long long wav_ld, const float* __restrict__ w,
                                                          const float2* __restrict__ scale_shift, void* __restrict__ out, int out_fmt,
                                                          int n_frames, long long out_batch_stride, int channels) {
  __shared__ float xs[FPB * kStride + kTaps + 6];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FPB;
  const int nt = min(FPB, n_frames - t0);
  const int nsamp = (nt - 1) * kStride + kTaps;
  const float* x = wav + (long long)b * wav_ld + (long long)t0 * kStride;
  for (int i = threadIdx.x; i < FPB * kStride + kTaps + 6; i += blockDim.x) xs[i] = i < nsamp ? __ldg(x + i) : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int cgrp = warp & 3, fhalf = warp >> 2;
  const bool bf = out_fmt == SCB_BF16;
  // ---- B fragments: 16 n-blocks x {taps t4*2, t4*2+1 | taps t4*2+8, t4*2+9} of channel ch(nb, g)
  uint32_t bfrag[16][2];
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    const int ch = cgrp * 128 + (g >> 2) * 64 + (nb >> 2) * 16 + ((g >> 1) & 1) * 8 + (nb & 3) * 2 + (g & 1);
    const float2 ss = scale_shift[(long long)b * channels + ch];
    const float* wc = w + ch * kTaps;
    bfrag[nb][0] = H16<SCB_F16>::pack(wc[t4 * 2] * ss.x, wc[t4 * 2 + 1] * ss.x);
    if (t4 == 0) {
      bfrag[nb][1] = H16<SCB_F16>::pack(wc[8] * ss.x, wc[9] * ss.x);
    } else if (t4 == 1) {
      const float hi = __half2float(__float2half_rn(ss.y));
      bfrag[nb][1] = H16<SCB_F16>::pack(hi, ss.y - hi);  // shift = hi + lo on the two constant-1 columns
    } else {
      bfrag[nb][1] = 0u;
    }
  }
  __syncthreads();
  uint16_t* obase = reinterpret_cast<uint16_t*>(out) + (long long)b * out_batch_stride + (long long)t0 * channels + cgrp * 128 +
                    (t4 >> 1) * 64 + (t4 & 1) * 8;
  for (int ft = fhalf; ft * 16 < nt; ft += 2) {
    const int f0 = ft * 16;
    // ---- A fragment: rows g / g+8 of the tile, columns (t4*2, +1) and (t4*2+8, +9)
    const float* xa = xs + (f0 + g) * kStride + t4 * 2;
    const float* xb = xa + 8 * kStride;
    uint32_t a[4];
    a[0] = H16<SCB_F16>::pack(xa[0], xa[1]);
    a[1] = H16<SCB_F16>::pack(xb[0], xb[1]);
    if (t4 == 0) {
      a[2] = H16<SCB_F16>::pack(xa[8], xa[9]);
      a[3] = H16<SCB_F16>::pack(xb[8], xb[9]);
    } else if (t4 == 1) {
      a[2] = a[3] = 0x3C003C00u;  // (1.0, 1.0)
    } else {
      a[2] = a[3] = 0u;
    }
    const bool ok_a = f0 + g < nt, ok_b = f0 + g + 8 < nt;
    uint16_t* oa = obase + (long long)(f0 + g) * channels;
    uint16_t* ob = oa + 8LL * channels;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {  // 4 n-blocks -> 8 channels (16 bytes) per row
      float c[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfrag[q4 * 4 + j][0]), "r"(bfrag[q4 * 4 + j][1]));
      }
      uint4 ua, ub;
      if (bf) {
        ua.x = gelu_pack<SCB_BF16>(c[0][0], c[0][1]); ua.y = gelu_pack<SCB_BF16>(c[1][0], c[1][1]);
        ua.z = gelu_pack<SCB_BF16>(c[2][0], c[2][1]); ua.w = gelu_pack<SCB_BF16>(c[3][0], c[3][1]);
        ub.x = gelu_pack<SCB_BF16>(c[0][2], c[0][3]); ub.y = gelu_pack<SCB_BF16>(c[1][2], c[1][3]);
        ub.z = gelu_pack<SCB_BF16>(c[2][2], c[2][3]); ub.w = gelu_pack<SCB_BF16>(c[3][2], c[3][3]);
      } else {
        ua.x = gelu_pack<SCB_F16>(c[0][0], c[0][1]); ua.y = gelu_pack<SCB_F16>(c[1][0], c[1][1]);
        ua.z = gelu_pack<SCB_F16>(c[2][0], c[2][1]); ua.w = gelu_pack<SCB_F16>(c[3][0], c[3][1]);
        ub.x = gelu_pack<SCB_F16>(c[0][2], c[0][3]); ub.y = gelu_pack<SCB_F16>(c[1][2], c[1][3]);
        ub.z = gelu_pack<SCB_F16>(c[2][2], c[2][3]); ub.w = gelu_pack<SCB_F16>(c[3][2], c[3][3]);
      }
      if (ok_a) *reinterpret_cast<uint4*>(oa + q4 * 16) = ua;
      if (ok_b) *reinterpret_cast<uint4*>(ob + q4 * 16) = ub;
    }
  }
}

// ---- HuBERT-large conv0 block (conv -> LayerNorm over the 512 channels of each frame -> GELU) on the tensor cores.
// The LayerNorm statistics of a frame need no pass over its 512 conv outputs: with o[f,c] = sum_k w[c,k] x[f,k] + b[c],
//   mean_f = wbar . x_f + mean(b),   E_c[o^2] = x_f^T G x_f + 2 (W^T b / C) . x_f + mean(b^2),   G = W^T W / C  (10 x 10),
// so 77 channel sums of the weights (conv0_ln_gram_kernel, once per call) turn them into ~75 FMAs per frame.  The normalisation
// then folds into the MMA operands like the GroupNorm variant above — here the per-FRAME factors ride on the A rows and the
// per-CHANNEL factors on the B columns:
//   A row f    = [ rstd_f x[f,0..9] | m_hi, m_lo | 1, 1 | rstd_f, rstd_f ]          m = -mean_f rstd_f (fp16 hi/lo pair)
//   B column c = [ g_c w[c,0..9]    | g_c,  g_c  | beta_c hi, lo | (g_c b_c) hi, lo ]
// and the accumulator is g_c (o - mean_f) rstd_f + beta_c: the epilogue is GELU + pack + store, 16 bytes per lane per store.
constexpr int kGramVals = kPairs + 2 * kTaps + 2;  // G (55 upper-triangular), wbar (10), W^T b / C (10), mean(b), mean(b^2)
__global__ void __launch_bounds__(32) conv0_ln_gram_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias, int channels,
                                                           float* __restrict__ gram) {
  const int i = blockIdx.x, lane = threadIdx.x;
  int k = 0, k2 = 0, kind = 0;  // kind 0: G[k][k2], 1: wbar[k], 2: (W^T b)[k], 3: mean b, 4: mean b^2
  if (i < kPairs) {
    int idx = i;
    while (idx >= kTaps - k) { idx -= kTaps - k; ++k; }
    k2 = k + idx;
  } else if (i < kPairs + kTaps) { kind = 1; k = i - kPairs; }
  else if (i < kPairs + 2 * kTaps) { kind = 2; k = i - kPairs - kTaps; }
  else kind = 3 + (i - kPairs - 2 * kTaps);
  float s = 0.f;
  for (int c = lane; c < channels; c += 32) {
    const float b = conv_bias ? conv_bias[c] : 0.f;
    const float wk = w[c * kTaps + k];
    s += kind == 0 ? wk * w[c * kTaps + k2] : kind == 1 ? wk : kind == 2 ? wk * b : kind == 3 ? b : b * b;
  }
  s = warp_sum(s);
  if (lane == 0) gram[i] = s / channels;
}

__global__ void __launch_bounds__(256) conv0_ln_apply_kernel(const float* __restrict__ wav, long long wav_ld, const float* __restrict__ w,
                                                             const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ gram, float eps,
                                                             void* __restrict__ out, int out_fmt, int n_frames,
                                                             long long out_batch_stride, int channels) {
  __shared__ float xs[kFramesPerBlock * kStride + kTaps + 6];
  __shared__ float sg[kGramVals];
  __shared__ float2 fstat[kFramesPerBlock];  // (rstd_f, -mean_f rstd_f)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerBlock;
  const int nt = min(kFramesPerBlock, n_frames - t0);
  const int nsamp = (nt - 1) * kStride + kTaps;
  const float* x = wav + (long long)b * wav_ld + (long long)t0 * kStride;
  for (int i = threadIdx.x; i < kFramesPerBlock * kStride + kTaps + 6; i += blockDim.x) xs[i] = i < nsamp ? __ldg(x + i) : 0.f;
  if (threadIdx.x < kGramVals) sg[threadIdx.x] = gram[threadIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int cgrp = warp & 3, fhalf = warp >> 2;
  const bool bf = out_fmt == SCB_BF16;
  // ---- B fragments: 16 n-blocks x {k = t4*2, t4*2+1 | k = t4*2+8, t4*2+9} of channel ch(nb, g)
  uint32_t bfrag[16][2];
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    const int ch = cgrp * 128 + (g >> 2) * 64 + (nb >> 2) * 16 + ((g >> 1) & 1) * 8 + (nb & 3) * 2 + (g & 1);  // see conv0_apply_kernel
    const float gm = gamma ? gamma[ch] : 1.f;
    const float* wc = w + ch * kTaps;
    bfrag[nb][0] = H16<SCB_F16>::pack(wc[t4 * 2] * gm, wc[t4 * 2 + 1] * gm);
    if (t4 == 0) {
      bfrag[nb][1] = H16<SCB_F16>::pack(wc[8] * gm, wc[9] * gm);
    } else if (t4 == 1) {
      bfrag[nb][1] = H16<SCB_F16>::pack(gm, gm);
    } else {
      const float v = t4 == 2 ? (beta ? beta[ch] : 0.f) : (conv_bias ? gm * conv_bias[ch] : 0.f);
      const float hi = __half2float(__float2half_rn(v));
      bfrag[nb][1] = H16<SCB_F16>::pack(hi, v - hi);
    }
  }
  __syncthreads();
  // ---- per-frame LayerNorm statistics from the 10 samples of the frame
  if (threadIdx.x < kFramesPerBlock) {
    const float* xf = xs + threadIdx.x * kStride;
    float xv[kTaps];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) xv[k] = xf[k];
    float mean = sg[kPairs + 2 * kTaps], e2 = sg[kPairs + 2 * kTaps + 1];
    int idx = 0;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
      mean = fmaf(sg[kPairs + k], xv[k], mean);
      e2 = fmaf(2.f * sg[kPairs + kTaps + k], xv[k], e2);
#pragma unroll
      for (int k2 = k; k2 < kTaps; ++k2) e2 = fmaf((k == k2 ? 1.f : 2.f) * sg[idx++], xv[k] * xv[k2], e2);
    }
    const float rstd = rsqrtf(fmaxf(e2 - mean * mean, 0.f) + eps);
    fstat[threadIdx.x] = make_float2(rstd, -mean * rstd);
  }
  __syncthreads();
  uint16_t* obase = reinterpret_cast<uint16_t*>(out) + (long long)b * out_batch_stride + (long long)t0 * channels + cgrp * 128 +
                    (t4 >> 1) * 64 + (t4 & 1) * 8;
  for (int ft = fhalf; ft * 16 < nt; ft += 2) {
    const int f0 = ft * 16;
    // ---- A fragment: rows g / g+8 of the tile, columns (t4*2, +1) and (t4*2+8, +9)
    const float2 sa = fstat[f0 + g], sb = fstat[f0 + g + 8];
    const float* xa = xs + (f0 + g) * kStride + t4 * 2;
    const float* xb = xa + 8 * kStride;
    uint32_t a[4];
    a[0] = H16<SCB_F16>::pack(xa[0] * sa.x, xa[1] * sa.x);
    a[1] = H16<SCB_F16>::pack(xb[0] * sb.x, xb[1] * sb.x);
    if (t4 == 0) {
      a[2] = H16<SCB_F16>::pack(xa[8] * sa.x, xa[9] * sa.x);
      a[3] = H16<SCB_F16>::pack(xb[8] * sb.x, xb[9] * sb.x);
    } else if (t4 == 1) {
      const float ha = __half2float(__float2half_rn(sa.y)), hb = __half2float(__float2half_rn(sb.y));
      a[2] = H16<SCB_F16>::pack(ha, sa.y - ha);
      a[3] = H16<SCB_F16>::pack(hb, sb.y - hb);
    } else if (t4 == 2) {
      a[2] = a[3] = 0x3C003C00u;  // (1.0, 1.0)
    } else {
      a[2] = H16<SCB_F16>::pack(sa.x, sa.x);
      a[3] = H16<SCB_F16>::pack(sb.x, sb.x);
    }
    const bool ok_a = f0 + g < nt, ok_b = f0 + g + 8 < nt;
    uint16_t* oa = obase + (long long)(f0 + g) * channels;
    uint16_t* ob = oa + 8LL * channels;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {  // 4 n-blocks -> 8 channels (16 bytes) per row
      float c[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfrag[q4 * 4 + j][0]), "r"(bfrag[q4 * 4 + j][1]));
      }
      uint4 ua, ub;
      if (bf) {
        ua.x = gelu_pack<SCB_BF16>(c[0][0], c[0][1]); ua.y = gelu_pack<SCB_BF16>(c[1][0], c[1][1]);
        ua.z = gelu_pack<SCB_BF16>(c[2][0], c[2][1]); ua.w = gelu_pack<SCB_BF16>(c[3][0], c[3][1]);
        ub.x = gelu_pack<SCB_BF16>(c[0][2], c[0][3]); ub.y = gelu_pack<SCB_BF16>(c[1][2], c[1][3]);
        ub.z = gelu_pack<SCB_BF16>(c[2][2], c[2][3]); ub.w = gelu_pack<SCB_BF16>(c[3][2], c[3][3]);
      } else {
        ua.x = gelu_pack<SCB_F16>(c[0][0], c[0][1]); ua.y = gelu_pack<SCB_F16>(c[1][0], c[1][1]);
        ua.z = gelu_pack<SCB_F16>(c[2][0], c[2][1]); ua.w = gelu_pack<SCB_F16>(c[3][0], c[3][1]);
        ub.x = gelu_pack<SCB_F16>(c[0][2], c[0][3]); ub.y = gelu_pack<SCB_F16>(c[1][2], c[1][3]);
        ub.z = gelu_pack<SCB_F16>(c[2][2], c[2][3]); ub.w = gelu_pack<SCB_F16>(c[3][2], c[3][3]);
      }
      if (ok_a) *reinterpret_cast<uint4*>(oa + q4 * 16) = ua;
      if (ok_b) *reinterpret_cast<uint4*>(ob + q4 * 16) = ub;
    }
  }
}

// x fp32 [B, T, D] (post_extract_proj output): zero the frames t >= valid[b] in place (speech_encoder_plus.py:32-33) and
// write the 16-bit, group-padded, time-padded copy the positional-conv GEMM reads:
//   xpad[b, pad_left + t, g*64 + c] = x[b, t, g*cpg + c]  (c < cpg);   everything else stays zero (buffer pre-zeroed once).
__global__ void __launch_bounds__(256) posconv_pack_kernel(float* __restrict__ x, const int* __restrict__ valid, void* __restrict__ xpad,
                                                           int fmt, long long rows, int T, int D, int groups, int cpg, int pad_left,
                                                           int rows_pad) {
  // One warp per row of the PADDED buffer: the kernel writes every element of xpad (frame rows, the zero rows on both sides and the
  // zero channels cpg..63 of each group), so the buffer needs no zero-initialisation and can be shared by every (batch, T).
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = (int)(row / rows_pad), tp = (int)(row % rows_pad);
  const int t = tp - pad_left;
  uint16_t* pr = reinterpret_cast<uint16_t*>(xpad) + row * (groups * 64);
  if (t < 0 || t >= T) {
    for (int c8 = lane; c8 * 8 < groups * 64; c8 += 32) *reinterpret_cast<uint4*>(pr + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float* xr = x + ((long long)b * T + t) * D;
  const bool keep = valid == nullptr || t < valid[b];
  for (int c2 = lane; c2 * 2 < D; c2 += 32) {
    const int c = c2 * 2;
    float2 v = *reinterpret_cast<float2*>(xr + c);
    if (!keep) {
      v = make_float2(0.f, 0.f);
      *reinterpret_cast<float2*>(xr + c) = v;
    }
    const int g = c / cpg, ci = c % cpg;  // cpg is even, so the pair stays inside one group
    *reinterpret_cast<uint32_t*>(pr + g * 64 + ci) = pack16(fmt, v.x, v.y);
  }
  if (cpg < 64) {
    const int padc = (64 - cpg) / 2;  // zero channel pairs per group
    for (int i = lane; i < groups * padc; i += 32) *reinterpret_cast<uint32_t*>(pr + (i / padc) * 64 + cpg + (i % padc) * 2) = 0u;
  }
}

// patches[b*G*G + gy*G + gx, c*P*P + py*P + px] = image[b, c, gy*P+py, gx*P+px]  (K padded with zeros to ldk)
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, void* __restrict__ out, int fmt, int C, int H,
                                                       int W, int P, int G, int ldk) {
  const long long prow = blockIdx.x;  // patch row index
  const int b = (int)(prow / (G * G)), gi = (int)(prow % (G * G));
  const int gy = gi / G, gx = gi % G;
  const int K = C * P * P;
  uint16_t* o = reinterpret_cast<uint16_t*>(out) + prow * ldk;
  for (int i2 = threadIdx.x; i2 * 2 < ldk; i2 += blockDim.x) {
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = i2 * 2 + j;
      if (idx < K) {
        const int c = idx / (P * P), r = idx % (P * P), py = r / P, px = r % P;
        v[j] = __ldg(img + (((long long)b * C + c) * H + gy * P + py) * W + gx * P + px);
      } else {
        v[j] = 0.f;
      }
    }
    *reinterpret_cast<uint32_t*>(o + i2 * 2) = pack16(fmt, v[0], v[1]);
  }
}

// out[b, :] = a[:] + (a2 ? a2[:] : 0)   for b in [0, nb); out row stride given; fp32 or 16-bit output.
__global__ void broadcast_row_kernel(const float* __restrict__ a, const float* __restrict__ a2, void* __restrict__ out, int out_dtype,
                                     long long out_stride, int d) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float v = a[c] + (a2 ? a2[c] : 0.f);
  if (out_dtype == SCB_F32)
    reinterpret_cast<float*>(out)[(long long)b * out_stride + c] = v;
  else if (out_dtype == SCB_F16)
    reinterpret_cast<__half*>(out)[(long long)b * out_stride + c] = __float2half_rn(v);
  else
    reinterpret_cast<__nv_bfloat16*>(out)[(long long)b * out_stride + c] = __float2bfloat16_rn(v);
}

// Generic strided 2-D cast/copy: out[r, c] = (T_out) in[r, c]; used for small layout fix-ups (e.g. fp32 -> 16-bit rows).
__global__ void cast_rows_kernel(const void* __restrict__ in, int in_dtype, long long in_ld, void* __restrict__ out, int out_dtype,
                                 long long out_ld, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    float v;
    if (in_dtype == SCB_F32) v = reinterpret_cast<const float*>(in)[r * in_ld + c];
    else if (in_dtype == SCB_F16) v = __half2float(reinterpret_cast<const __half*>(in)[r * in_ld + c]);
    else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[r * in_ld + c]);
    if (out_dtype == SCB_F32) reinterpret_cast<float*>(out)[r * out_ld + c] = v;
    else if (out_dtype == SCB_F16) reinterpret_cast<__half*>(out)[r * out_ld + c] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(out)[r * out_ld + c] = __float2bfloat16_rn(v);
  }
}

// out[c, r] = (T_out) in[r, c] through a padded 32x32 shared tile (coalesced both ways).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_kernel(const TI* __restrict__ in, long long in_ld, TO* __restrict__ out, long long out_ld,
                                                        int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = (float)in[(long long)r * in_ld + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(long long)c * out_ld + r] = (TO)tile[threadIdx.x][j];
  }
}

// 16-bit -> 16-bit transpose (the bf16 operands of the head's wgrad GEMMs: [B*(T+1), 2d] and [B*(T+1), d] per step) with 16-byte
// global accesses on both sides.  A 64 x 64 tile lives in shared memory as 64 rows of 32 words (two adjacent columns per word)
// + 1 pad word.  Load: 8 lanes read one 128-byte input row.  Store: a thread gathers word w of 8 consecutive rows, splits the
// low / high halves into the 16-byte pieces of output rows 2w and 2w+1, and 8 lanes (8 row groups) write one full 128-byte
// output row — the element-wise 32 x 32 version moved 64 bytes per warp instruction (0.4 ms per step, 31 % of the HBM rate).
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, long long in_ld, uint16_t* __restrict__ out,
                                                          long long out_ld, int rows, int cols) {
  __shared__ uint32_t tile[64][33];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int row = (t >> 3) + 32 * pass, chunk = t & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + row < rows && c0 + chunk * 8 < cols) v = *reinterpret_cast<const uint4*>(in + (long long)(r0 + row) * in_ld + c0 + chunk * 8);
    uint32_t* d = &tile[row][chunk * 4];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int g = t & 7, w = t >> 3;   // 8 input rows r0 + 8g .. +7; input columns c0 + 2w, c0 + 2w + 1
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = tile[g * 8 + i][w];
  uint4 lo, hi;
  lo.x = __byte_perm(x[0], x[1], 0x5410); lo.y = __byte_perm(x[2], x[3], 0x5410); lo.z = __byte_perm(x[4], x[5], 0x5410); lo.w = __byte_perm(x[6], x[7], 0x5410);
  hi.x = __byte_perm(x[0], x[1], 0x7632); hi.y = __byte_perm(x[2], x[3], 0x7632); hi.z = __byte_perm(x[4], x[5], 0x7632); hi.w = __byte_perm(x[6], x[7], 0x7632);
  const int r = r0 + g * 8, c = c0 + 2 * w;
  if (r < rows) {   // rows % 8 == 0 (checked by the launcher): a group of 8 is inside or outside
    if (c < cols) *reinterpret_cast<uint4*>(out + (long long)c * out_ld + r) = lo;
    if (c + 1 < cols) *reinterpret_cast<uint4*>(out + (long long)(c + 1) * out_ld + r) = hi;
  }
}

}  // namespace

int conv0_groupnorm_gelu(const float* wav, long long wav_ld, int batch, int n_samples, const float* w, const float* conv_bias,
                         const float* gamma, const float* beta, float eps, void* out, int out_fmt, long long out_batch_stride,
                         void* scratch, long long scratch_bytes, cudaStream_t st) {
  const int channels = 512;
  SCB_CHECK(wav && w && out && scratch, SCB_EINVAL, "scb_conv0_groupnorm_gelu: null operand");
  SCB_CHECK(out_fmt == SCB_F16 || out_fmt == SCB_BF16, SCB_EINVAL, "scb_conv0_groupnorm_gelu: 16-bit output required");
  SCB_CHECK(n_samples >= kTaps, SCB_EINVAL, "scb_conv0_groupnorm_gelu: utterance shorter than the conv kernel");
  const int n_frames = (n_samples - kTaps) / kStride + 1;
  const long long need = conv0_scratch_bytes(batch);
  SCB_CHECK(batch <= 65535, SCB_EUNSUPPORTED, "scb_conv0_groupnorm_gelu: batch exceeds grid limits");
  SCB_CHECK(scratch_bytes >= need, SCB_EINVAL, "scb_conv0_groupnorm_gelu: scratch too small (%lld < %lld)", scratch_bytes, need);
  if (batch == 0) return SCB_OK;
  double* acc = reinterpret_cast<double*>(scratch);
  float2* ss = reinterpret_cast<float2*>(acc + (long long)batch * kStatVals);
  SCB_CUDA(cudaMemsetAsync(acc, 0, (size_t)batch * kStatVals * sizeof(double), st));
  int chunks = (n_frames + 2559) / 2560;
  if (chunks < 1) chunks = 1;
  conv0_autocorr_kernel<<<dim3(chunks, batch), 256, 0, st>>>(wav, wav_ld, n_frames, acc);
  note_launch();
  SCB_LAUNCH_OK("conv0_autocorr");
  conv0_finalize_kernel<<<dim3((channels + 127) / 128, batch), 128, 0, st>>>(acc, w, conv_bias, gamma, beta, n_frames, channels, eps, ss);
  note_launch();
  SCB_LAUNCH_OK("conv0_finalize");
  static const int fpb = [] { const char* e = getenv("SCB_CONV0_FPB"); return e ? atoi(e) : 256; }();  // measured at 256 x 102400 samples: 128 -> 1.74 ms, 256 -> 1.58 ms, 512 -> 1.56 ms
  if (fpb == 512)
    conv0_apply_kernel<512><<<dim3((n_frames + 511) / 512, batch), 256, 0, st>>>(wav, wav_ld, w, ss, out, out_fmt,
                                                                                n_frames, out_batch_stride, channels);
  else if (fpb != 128)
    conv0_apply_kernel<256><<<dim3((n_frames + 255) / 256, batch), 256, 0, st>>>(wav, wav_ld, w, ss, out, out_fmt,
                                                                                n_frames, out_batch_stride, channels);
  else
  conv0_apply_kernel<128><<<dim3((n_frames + kFramesPerBlock - 1) / kFramesPerBlock, batch), 256, 0, st>>>(wav, wav_ld, w, ss, out, out_fmt,
                                                                                                     n_frames, out_batch_stride, channels);
  note_launch();
  SCB_LAUNCH_OK("conv0_apply");
  return SCB_OK;
}


long long conv0_scratch_bytes(int batch) {
  return (long long)batch * kStatVals * (long long)sizeof(double) + (long long)batch * 512 * (long long)sizeof(float2);
}

int conv0_layernorm_gelu(const float* wav, long long wav_ld, int batch, int n_samples, const float* w, const float* conv_bias,
                         const float* gamma, const float* beta, float eps, void* out, int out_fmt, long long out_batch_stride,
                         void* scratch, long long scratch_bytes, cudaStream_t st) {
  SCB_CHECK(wav && w && out && scratch, SCB_EINVAL, "scb_conv0_layernorm_gelu: null operand");
  SCB_CHECK(out_fmt == SCB_F16 || out_fmt == SCB_BF16, SCB_EINVAL, "scb_conv0_layernorm_gelu: 16-bit output required");
  SCB_CHECK(n_samples >= kTaps, SCB_EINVAL, "scb_conv0_layernorm_gelu: utterance shorter than the conv kernel");
  SCB_CHECK(batch <= 65535, SCB_EUNSUPPORTED, "scb_conv0_layernorm_gelu: batch exceeds grid limits");
  SCB_CHECK(scratch_bytes >= (long long)(kGramVals * sizeof(float)), SCB_EINVAL, "scb_conv0_layernorm_gelu: scratch too small (%lld < %d)",
            scratch_bytes, (int)(kGramVals * sizeof(float)));
  if (batch == 0) return SCB_OK;
  const int n_frames = (n_samples - kTaps) / kStride + 1;
  float* gram = reinterpret_cast<float*>(scratch);
  conv0_ln_gram_kernel<<<kGramVals, 32, 0, st>>>(w, conv_bias, 512, gram);
  note_launch();
  SCB_LAUNCH_OK("conv0_ln_gram");
  conv0_ln_apply_kernel<<<dim3((n_frames + kFramesPerBlock - 1) / kFramesPerBlock, batch), 256, 0, st>>>(
      wav, wav_ld, w, conv_bias, gamma, beta, gram, eps, out, out_fmt, n_frames, out_batch_stride, 512);
  note_launch();
  SCB_LAUNCH_OK("conv0_layernorm_gelu");
  return SCB_OK;
}

int posconv_pack(float* x, const int* valid_frames, void* xpad, int fmt, int batch, int T, int D, int groups, int pad_left, int rows_pad,
                 cudaStream_t st) {
  SCB_CHECK(x && xpad, SCB_EINVAL, "scb_posconv_pack: null operand");
  SCB_CHECK(D % groups == 0 && (D / groups) % 2 == 0 && D / groups <= 64, SCB_EUNSUPPORTED,
            "scb_posconv_pack: channels per group (%d) must be even and <= 64", D / groups);
  SCB_CHECK(rows_pad >= pad_left + T, SCB_EINVAL, "scb_posconv_pack: rows_pad too small");
  const long long rows = (long long)batch * rows_pad;
  if (rows == 0) return SCB_OK;
  posconv_pack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, valid_frames, xpad, fmt, rows, T, D, groups, D / groups, pad_left,
                                                                 rows_pad);
  note_launch();
  SCB_LAUNCH_OK("posconv_pack");
  return SCB_OK;
}

int patchify(const float* img, void* out, int fmt, int batch, int C, int H, int W, int P, int ldk, cudaStream_t st) {
  SCB_CHECK(img && out, SCB_EINVAL, "scb_patchify: null operand");
  SCB_CHECK(H % P == 0 && W % P == 0 && H == W, SCB_EINVAL, "scb_patchify: image %dx%d not divisible by patch %d", H, W, P);
  SCB_CHECK(ldk >= C * P * P && ldk % 8 == 0, SCB_EINVAL, "scb_patchify: ldk must be >= C*P*P and a multiple of 8");
  const int G = H / P;
  if (batch == 0) return SCB_OK;
  patchify_kernel<<<(unsigned)((long long)batch * G * G), 256, 0, st>>>(img, out, fmt, C, H, W, P, G, ldk);
  note_launch();
  SCB_LAUNCH_OK("patchify");
  return SCB_OK;
}

int broadcast_row(const float* a, const float* a2, void* out, int out_dtype, long long out_stride, int nb, int d, cudaStream_t st) {
  SCB_CHECK(a && out, SCB_EINVAL, "scb_broadcast_row: null operand");
  if (nb == 0 || d == 0) return SCB_OK;
  broadcast_row_kernel<<<dim3((d + 255) / 256, nb), 256, 0, st>>>(a, a2, out, out_dtype, out_stride, d);
  note_launch();
  SCB_LAUNCH_OK("broadcast_row");
  return SCB_OK;
}

int cast_rows(const void* in, int in_dtype, long long in_ld, void* out, int out_dtype, long long out_ld, long long rows, int cols,
              cudaStream_t st) {
  SCB_CHECK(in && out, SCB_EINVAL, "scb_cast_rows: null operand");
  if (rows == 0 || cols == 0) return SCB_OK;
  long long blocks = (rows * cols + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  cast_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, in_dtype, in_ld, out, out_dtype, out_ld, rows, cols);
  note_launch();
  SCB_LAUNCH_OK("cast_rows");
  return SCB_OK;
}

template <typename TI>
static int transpose_out(const TI* in, long long in_ld, void* out, int out_dtype, long long out_ld, int rows, int cols, cudaStream_t st) {
  const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  if (out_dtype == SCB_F32) transpose_kernel<TI, float><<<grid, block, 0, st>>>(in, in_ld, (float*)out, out_ld, rows, cols);
  else if (out_dtype == SCB_F16) transpose_kernel<TI, __half><<<grid, block, 0, st>>>(in, in_ld, (__half*)out, out_ld, rows, cols);
  else transpose_kernel<TI, __nv_bfloat16><<<grid, block, 0, st>>>(in, in_ld, (__nv_bfloat16*)out, out_ld, rows, cols);
  note_launch();
  SCB_LAUNCH_OK("transpose");
  return SCB_OK;
}

int transpose(const void* in, int in_dtype, long long in_ld, void* out, int out_dtype, long long out_ld, int rows, int cols,
              cudaStream_t st) {
  SCB_CHECK(in && out, SCB_EINVAL, "scb_transpose: null operand");
  if (rows == 0 || cols == 0) return SCB_OK;
  SCB_CHECK((rows + 31) / 32 <= 65535, SCB_EUNSUPPORTED, "scb_transpose: too many rows (%d)", rows);
  if (in_dtype == out_dtype && in_dtype != SCB_F32 && rows % 8 == 0 && cols % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    transpose16_kernel<<<dim3((cols + 63) / 64, (rows + 63) / 64), 256, 0, st>>>(static_cast<const uint16_t*>(in), in_ld, static_cast<uint16_t*>(out),
                                                                                  out_ld, rows, cols);
    note_launch();
    SCB_LAUNCH_OK("transpose16");
    return SCB_OK;
  }
  if (in_dtype == SCB_F32) return transpose_out((const float*)in, in_ld, out, out_dtype, out_ld, rows, cols, st);
  if (in_dtype == SCB_F16) return transpose_out((const __half*)in, in_ld, out, out_dtype, out_ld, rows, cols, st);
  return transpose_out((const __nv_bfloat16*)in, in_ld, out, out_dtype, out_ld, rows, cols, st);
}

}  // namespace scb
