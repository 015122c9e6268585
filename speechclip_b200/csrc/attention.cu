// Attention kernels.
//   attention_fwd : softmax(Q K^T * scale + mask) V for short sequences (T <= ~512: HuBERT 319, branch 320/327, ViT 50/257,
//                   text 77); key-padding as a per-batch valid length, optional causal mask.  Flash-style single pass:
//                   64 query rows per CTA (4 warps x m16), K/V streamed through a cp.async double buffer in 64-key tiles,
//                   online softmax in fp32 registers, tensor-core mma.sync m16n8k16 (round-1 kernel; the tcgen05/TMEM
//                   version is the planned replacement).
//   cls_attention_fwd/bwd : the parallel branch only consumes output row 0 (kwClip.py:1103), and row 0 of its input is
//                   the same [CLS] vector for every utterance, so attention reduces to ONE query per (utterance, head)
//                   against all keys: an HBM-bound kernel that reads K,V once.  Backward yields dK, dV and dq.
#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int BQ = 64;   // query rows per CTA
constexpr int BKV = 64;  // keys per tile

struct AttnParams {
  const uint16_t* q;
  const uint16_t* k;
  const uint16_t* v;
  uint16_t* o;
  long long q_ld, k_ld, v_ld, o_ld;                 // row strides (elements)
  long long q_bs, k_bs, v_bs, o_bs;                 // batch strides (elements)
  const int* kv_len;                                // per batch valid keys (nullable -> Tk)
  int Tq, Tk, heads;
  float scale_log2;                                 // softmax scale * log2(e)
  int causal;
};

// smem row pitch in 16-byte chunks: 8 for HD=64, 16 for HD=96/128; chunk index swizzled with (row & 7) on its low 3 bits.
template <int HD>
struct Cfg {
  static constexpr int CHUNKS = HD / 8;
  static constexpr int PITCH_CHUNKS = HD == 64 ? 8 : 16;
  static constexpr int TILE_ELEMS = 64 * PITCH_CHUNKS * 8;
};

template <int HD>
__device__ __forceinline__ uint16_t* tile_ptr(uint16_t* base, int row, int chunk) {
  const int sw = (chunk & ~7) | ((chunk ^ row) & 7);
  return base + (row * Cfg<HD>::PITCH_CHUNKS + sw) * 8;
}

template <int HD>
__device__ __forceinline__ void load_tile(uint16_t* smem, const uint16_t* g, long long ld, int row0, int nrows_valid) {
  constexpr int CH = Cfg<HD>::CHUNKS;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx % CH;
    const bool ok = row0 + r < nrows_valid;
    cp_async16(tile_ptr<HD>(smem, r, c), g + (long long)(ok ? row0 + r : 0) * ld + c * 8, ok);
  }
}

template <int HD, bool BF16>
__global__ void __launch_bounds__(128) attention_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint16_t* sQ = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* sK = sQ + Cfg<HD>::TILE_ELEMS;       // 2 buffers
  uint16_t* sV = sK + 2 * Cfg<HD>::TILE_ELEMS;   // 2 buffers

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = qt * BQ;

  const uint16_t* Q = p.q + (long long)b * p.q_bs + h * HD;
  const uint16_t* K = p.k + (long long)b * p.k_bs + h * HD;
  const uint16_t* V = p.v + (long long)b * p.v_bs + h * HD;

  int kv_len = p.kv_len ? min(p.kv_len[b], p.Tk) : p.Tk;
  int kv_end = kv_len;
  if (p.causal) kv_end = min(kv_end, q0 + BQ);
  const int n_tiles = (kv_end + BKV - 1) / BKV;

  load_tile<HD>(sQ, Q, p.q_ld, q0, p.Tq);
  load_tile<HD>(sK, K, p.k_ld, 0, kv_end);
  load_tile<HD>(sV, V, p.v_ld, 0, kv_end);
  cp_async_commit();

  constexpr int KS = HD / 16;  // k-steps over the head dim
  uint32_t qf[KS][4];
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;

  for (int tile = 0; tile < n_tiles; ++tile) {
    cp_async_wait<0>();
    __syncthreads();
    if (tile + 1 < n_tiles) {  // prefetch next K/V tile into the other buffer
      load_tile<HD>(sK + ((tile + 1) & 1) * Cfg<HD>::TILE_ELEMS, K, p.k_ld, (tile + 1) * BKV, kv_end);
      load_tile<HD>(sV + ((tile + 1) & 1) * Cfg<HD>::TILE_ELEMS, V, p.v_ld, (tile + 1) * BKV, kv_end);
      cp_async_commit();
    }
    if (tile == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(qf[ks], tile_ptr<HD>(sQ, warp * 16 + (lane & 15), ks * 2 + (lane >> 4)));
    }
    const uint16_t* sKt = sK + (tile & 1) * Cfg<HD>::TILE_ELEMS;
    const uint16_t* sVt = sV + (tile & 1) * Cfg<HD>::TILE_ELEMS;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {
        uint32_t kb[4];
        ldmatrix_x4(kb, tile_ptr<HD>(const_cast<uint16_t*>(sKt), nb2 * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
        mma16816<BF16>(s[nb2 * 2], qf[ks], kb[0], kb[1]);
        mma16816<BF16>(s[nb2 * 2 + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // ---- mask (boundary / causal tiles only) + online softmax.  The softmax scale is folded into the exponent:
    //      p = 2^(s * scale_log2 - m * scale_log2), with the running max kept on the raw scores (scale > 0).
    const int key0 = tile * BKV;
    if (key0 + BKV > kv_len || p.causal) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = key0 + nb * 8 + t4 * 2 + (j & 1);
          const int row = (j < 2) ? row_a : row_b;
          const bool ok = key < kv_len && (!p.causal || key <= row);
          if (!ok) s[nb][j] = -INFINITY;
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
    float corr[2], m_off[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = ex2_approx((m_run[r] - m_use) * p.scale_log2);  // m_run = -inf -> 0
      m_off[r] = -m_use * p.scale_log2;
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for 4 k16 steps over the 64 keys
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = ex2_approx(fmaf(s[nb][0], p.scale_log2, m_off[0])), p1 = ex2_approx(fmaf(s[nb][1], p.scale_log2, m_off[0]));
      const float p2 = ex2_approx(fmaf(s[nb][2], p.scale_log2, m_off[1])), p3 = ex2_approx(fmaf(s[nb][3], p.scale_log2, m_off[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const uint32_t lo = BF16 ? H16<SCB_BF16>::pack(p0, p1) : H16<SCB_F16>::pack(p0, p1);
      const uint32_t hi = BF16 ? H16<SCB_BF16>::pack(p2, p3) : H16<SCB_F16>::pack(p2, p3);
      // C fragment of n-block nb -> A fragment regs of k-step nb/2: even nb -> a0 (row g), a1 (row g+8); odd nb -> a2, a3
      pf[nb >> 1][(nb & 1) * 2 + 0] = lo;
      pf[nb >> 1][(nb & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int db = 0; db < HD / 8; ++db) {
      o_acc[db][0] *= corr[0]; o_acc[db][1] *= corr[0];
      o_acc[db][2] *= corr[1]; o_acc[db][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks2 = 0; ks2 < 4; ++ks2) {
#pragma unroll
      for (int db2 = 0; db2 < HD / 16; ++db2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, tile_ptr<HD>(const_cast<uint16_t*>(sVt), ks2 * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), db2 * 2 + (lane >> 4)));
        mma16816<BF16>(o_acc[db2 * 2], pf[ks2], vb[0], vb[1]);
        mma16816<BF16>(o_acc[db2 * 2 + 1], pf[ks2], vb[2], vb[3]);
      }
    }
  }
  // ---- finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv_a = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv_b = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  uint16_t* O = p.o + (long long)b * p.o_bs + h * HD;
#pragma unroll
  for (int db = 0; db < HD / 8; ++db) {
    const int col = db * 8 + t4 * 2;
    if (row_a < p.Tq) {
      const uint32_t u = BF16 ? H16<SCB_BF16>::pack(o_acc[db][0] * inv_a, o_acc[db][1] * inv_a) : H16<SCB_F16>::pack(o_acc[db][0] * inv_a, o_acc[db][1] * inv_a);
      *reinterpret_cast<uint32_t*>(O + (long long)row_a * p.o_ld + col) = u;
    }
    if (row_b < p.Tq) {
      const uint32_t u = BF16 ? H16<SCB_BF16>::pack(o_acc[db][2] * inv_b, o_acc[db][3] * inv_b) : H16<SCB_F16>::pack(o_acc[db][2] * inv_b, o_acc[db][3] * inv_b);
      *reinterpret_cast<uint32_t*>(O + (long long)row_b * p.o_ld + col) = u;
    }
  }
}

template <int HD, bool BF16>
int launch_attn(const AttnParams& p, int batch, cudaStream_t st) {
  constexpr int smem = 5 * Cfg<HD>::TILE_ELEMS * 2;
  static bool configured = false;
  if (!configured) {
    SCB_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<HD, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const dim3 grid((p.Tq + BQ - 1) / BQ, p.heads, batch);
  attention_fwd_kernel<HD, BF16><<<grid, 128, smem, st>>>(p);
  note_launch();
  SCB_LAUNCH_OK("attention_fwd");
  return SCB_OK;
}

// ------------------------------------------------------------------------------------------------ single-query attention
// One warp per (batch, head).  q fp32 [heads*hd] shared by all batches.  kv 16-bit [B, Tk, ld] with K at column k_off + h*hd
// and V at v_off + h*hd.  probs saved [B, heads, Tk] fp32 for backward.  ctx fp32 [B, heads*hd] (+ optional 16-bit copy).
template <int HD>
__global__ void __launch_bounds__(128) cls_attention_fwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                const int* __restrict__ kv_len, int Tk, int heads, float scale,
                                                                float* __restrict__ probs, float* __restrict__ ctx32,
                                                                uint16_t* __restrict__ ctx16, int ctx16_fmt, int batch) {
  extern __shared__ float sprob[];  // [warps][Tk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= batch * heads) return;
  const int b = idx / heads, h = idx % heads;
  float* pr = sprob + warp * Tk;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs;
  // q slice in registers: each lane holds HD/32 pairs? use strided ownership: lane owns dims {2*lane + 64*i}
  constexpr int PAIRS = (HD + 63) / 64;
  float2 qv[PAIRS];
#pragma unroll
  for (int i = 0; i < PAIRS; ++i) {
    const int d = 2 * lane + 64 * i;
    qv[i] = d < HD ? make_float2(q[h * HD + d] * scale, q[h * HD + d + 1] * scale) : make_float2(0.f, 0.f);
  }
  float mx = -INFINITY;
  for (int j = 0; j < len; ++j) {
    const uint16_t* kr = base + (long long)j * kv_ld + k_off + h * HD;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int d = 2 * lane + 64 * i;
      if (d < HD) {
        const float2 kk = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(kr + d));
        dot += qv[i].x * kk.x + qv[i].y * kk.y;
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) pr[j] = dot;
    mx = fmaxf(mx, dot);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < len; j += 32) {
    const float e = __expf(pr[j] - mx);
    pr[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __syncwarp();
  float2 acc[PAIRS];
#pragma unroll
  for (int i = 0; i < PAIRS; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int j = 0; j < len; ++j) {
    const float pj = pr[j] * inv;
    const uint16_t* vr = base + (long long)j * kv_ld + v_off + h * HD;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int d = 2 * lane + 64 * i;
      if (d < HD) {
        const float2 vv = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(vr + d));
        acc[i].x += pj * vv.x;
        acc[i].y += pj * vv.y;
      }
    }
  }
  float* po = probs + ((long long)b * heads + h) * Tk;
  for (int j = lane; j < Tk; j += 32) po[j] = j < len ? pr[j] * inv : 0.f;
#pragma unroll
  for (int i = 0; i < PAIRS; ++i) {
    const int d = 2 * lane + 64 * i;
    if (d < HD) {
      const long long off = (long long)b * heads * HD + h * HD + d;
      if (ctx32) *reinterpret_cast<float2*>(ctx32 + off) = acc[i];
      if (ctx16) *reinterpret_cast<uint32_t*>(ctx16 + off) = pack16(ctx16_fmt, acc[i].x, acc[i].y);
    }
  }
}

// Backward of the single-query attention.  dctx fp32 [B, heads*hd].  Writes dKV 16-bit [B, Tk, ld] (K grads at k_off, V grads
// at v_off; rows >= len are zero) and accumulates dq (unscaled q gradient, fp32 [heads*hd]) with atomics.
template <int HD>
__global__ void __launch_bounds__(128) cls_attention_bwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                const int* __restrict__ kv_len, int Tk, int heads, float scale,
                                                                const float* __restrict__ probs, const float* __restrict__ dctx,
                                                                uint16_t* __restrict__ dkv, int dkv_fmt, float* __restrict__ dq, int batch) {
  extern __shared__ float sds[];  // [warps][Tk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= batch * heads) return;
  const int b = idx / heads, h = idx % heads;
  float* ds = sds + warp * Tk;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs;
  uint16_t* dbase = dkv + (long long)b * kv_bs;
  const float* pr = probs + ((long long)b * heads + h) * Tk;
  constexpr int PAIRS = (HD + 63) / 64;
  float2 qv[PAIRS], dc[PAIRS], dqa[PAIRS];
#pragma unroll
  for (int i = 0; i < PAIRS; ++i) {
    const int d = 2 * lane + 64 * i;
    const bool ok = d < HD;
    qv[i] = ok ? make_float2(q[h * HD + d], q[h * HD + d + 1]) : make_float2(0.f, 0.f);
    dc[i] = ok ? *reinterpret_cast<const float2*>(dctx + (long long)b * heads * HD + h * HD + d) : make_float2(0.f, 0.f);
    dqa[i] = make_float2(0.f, 0.f);
  }
  // dp_j = <dctx, v_j>;  dot = sum_j p_j dp_j
  float dot = 0.f;
  for (int j = 0; j < len; ++j) {
    const uint16_t* vr = base + (long long)j * kv_ld + v_off + h * HD;
    float dp = 0.f;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int d = 2 * lane + 64 * i;
      if (d < HD) {
        const float2 vv = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(vr + d));
        dp += dc[i].x * vv.x + dc[i].y * vv.y;
      }
    }
    dp = warp_sum(dp);
    if (lane == 0) ds[j] = dp;
    dot += pr[j] * dp;
  }
  __syncwarp();
  for (int j = 0; j < Tk; ++j) {
    uint16_t* dkr = dbase + (long long)j * kv_ld + k_off + h * HD;
    uint16_t* dvr = dbase + (long long)j * kv_ld + v_off + h * HD;
    if (j < len) {
      const float pj = pr[j];
      const float dsj = pj * (ds[j] - dot) * scale;  // d(score_j) * scale: score = scale * <q, k_j>
      const uint16_t* kr = base + (long long)j * kv_ld + k_off + h * HD;
#pragma unroll
      for (int i = 0; i < PAIRS; ++i) {
        const int d = 2 * lane + 64 * i;
        if (d < HD) {
          const float2 kk = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(kr + d));
          dqa[i].x += dsj * kk.x;
          dqa[i].y += dsj * kk.y;
          *reinterpret_cast<uint32_t*>(dkr + d) = pack16(dkv_fmt, dsj * qv[i].x, dsj * qv[i].y);
          *reinterpret_cast<uint32_t*>(dvr + d) = pack16(dkv_fmt, pj * dc[i].x, pj * dc[i].y);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < PAIRS; ++i) {
        const int d = 2 * lane + 64 * i;
        if (d < HD) {
          *reinterpret_cast<uint32_t*>(dkr + d) = 0u;
          *reinterpret_cast<uint32_t*>(dvr + d) = 0u;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PAIRS; ++i) {
    const int d = 2 * lane + 64 * i;
    if (d < HD) {
      atomicAdd(&dq[h * HD + d], dqa[i].x);
      atomicAdd(&dq[h * HD + d + 1], dqa[i].y);
    }
  }
}

}  // namespace

int attention_fwd(const void* q, const void* k, const void* v, void* o, int fmt, long long q_ld, long long k_ld, long long v_ld,
                  long long o_ld, long long q_bs, long long k_bs, long long v_bs, long long o_bs, const int* kv_len, int batch, int heads,
                  int head_dim, int Tq, int Tk, float scale, int causal, cudaStream_t st) {
  SCB_CHECK(q && k && v && o, SCB_EINVAL, "scb_attention_fwd: null operand");
  SCB_CHECK(fmt == SCB_F16 || fmt == SCB_BF16, SCB_EINVAL, "scb_attention_fwd: 16-bit operands required");
  SCB_CHECK((q_ld | k_ld | v_ld | o_ld | q_bs | k_bs | v_bs | o_bs) % 8 == 0, SCB_EINVAL, "scb_attention_fwd: strides must be multiples of 8");
  SCB_CHECK(heads <= 65535 && batch <= 65535, SCB_EUNSUPPORTED, "scb_attention_fwd: batch/heads exceed grid limits");
  if (batch == 0 || Tq == 0) return SCB_OK;
  AttnParams p;
  p.q = (const uint16_t*)q; p.k = (const uint16_t*)k; p.v = (const uint16_t*)v; p.o = (uint16_t*)o;
  p.q_ld = q_ld; p.k_ld = k_ld; p.v_ld = v_ld; p.o_ld = o_ld;
  p.q_bs = q_bs; p.k_bs = k_bs; p.v_bs = v_bs; p.o_bs = o_bs;
  p.kv_len = kv_len; p.Tq = Tq; p.Tk = Tk; p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  const bool bf = fmt == SCB_BF16;
  switch (head_dim) {
    case 64: return bf ? launch_attn<64, true>(p, batch, st) : launch_attn<64, false>(p, batch, st);
    case 96: return bf ? launch_attn<96, true>(p, batch, st) : launch_attn<96, false>(p, batch, st);
    case 128: return bf ? launch_attn<128, true>(p, batch, st) : launch_attn<128, false>(p, batch, st);
    case 16: return bf ? launch_attn<16, true>(p, batch, st) : launch_attn<16, false>(p, batch, st);
    case 32: return bf ? launch_attn<32, true>(p, batch, st) : launch_attn<32, false>(p, batch, st);
  }
  SCB_CHECK(false, SCB_EUNSUPPORTED, "scb_attention_fwd: head_dim %d not in {16,32,64,96,128}", head_dim);
}

#define SCB_HD_SWITCH(HDV, CALL)                                                        \
  switch (HDV) {                                                                        \
    case 8: { constexpr int HD_ = 8; CALL; break; }                                     \
    case 16: { constexpr int HD_ = 16; CALL; break; }                                   \
    case 32: { constexpr int HD_ = 32; CALL; break; }                                   \
    case 64: { constexpr int HD_ = 64; CALL; break; }                                   \
    case 96: { constexpr int HD_ = 96; CALL; break; }                                   \
    case 128: { constexpr int HD_ = 128; CALL; break; }                                 \
    default: SCB_CHECK(false, SCB_EUNSUPPORTED, "cls_attention: head_dim %d unsupported", HDV); \
  }

int cls_attention_fwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off,
                      const int* kv_len, int batch, int heads, int head_dim, int Tk, float scale, float* probs, float* ctx32, void* ctx16,
                      int ctx16_fmt, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && (ctx32 || ctx16), SCB_EINVAL, "scb_cls_attention_fwd: null operand");
  if (batch == 0) return SCB_OK;
  const int warps = 4;
  const size_t smem = (size_t)warps * Tk * sizeof(float);
  SCB_CHECK(smem <= 48 * 1024, SCB_EUNSUPPORTED, "scb_cls_attention_fwd: Tk=%d too long", Tk);
  const unsigned grid = (unsigned)((batch * heads + warps - 1) / warps);
  SCB_HD_SWITCH(head_dim, (cls_attention_fwd_kernel<HD_><<<grid, warps * 32, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads, scale, probs, ctx32, (uint16_t*)ctx16, ctx16_fmt, batch)));
  note_launch();
  SCB_LAUNCH_OK("cls_attention_fwd");
  return SCB_OK;
}

int cls_attention_bwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off,
                      const int* kv_len, int batch, int heads, int head_dim, int Tk, float scale, const float* probs, const float* dctx,
                      void* dkv, int dkv_fmt, float* dq, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && dctx && dkv && dq, SCB_EINVAL, "scb_cls_attention_bwd: null operand");
  if (batch == 0) return SCB_OK;
  const int warps = 4;
  const size_t smem = (size_t)warps * Tk * sizeof(float);
  SCB_CHECK(smem <= 48 * 1024, SCB_EUNSUPPORTED, "scb_cls_attention_bwd: Tk=%d too long", Tk);
  const unsigned grid = (unsigned)((batch * heads + warps - 1) / warps);
  SCB_HD_SWITCH(head_dim, (cls_attention_bwd_kernel<HD_><<<grid, warps * 32, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads, scale, probs, dctx, (uint16_t*)dkv, dkv_fmt, dq, batch)));
  note_launch();
  SCB_LAUNCH_OK("cls_attention_bwd");
  return SCB_OK;
}

}  // namespace scb
