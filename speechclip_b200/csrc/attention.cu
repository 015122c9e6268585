// Attention kernels.
//   attention_fwd : softmax(Q K^T * scale + mask) V for short sequences (T <= ~512: HuBERT 319, branch 320/327, ViT 50/257,
//                   text 77); key-padding as a per-batch valid length, optional causal mask.  Flash-style single pass:
//                   64 query rows per CTA (4 warps x m16), K/V streamed through a cp.async double buffer in 64-key tiles,
//                   online softmax in fp32 registers, tensor-core mma.sync m16n8k16 (round-1 kernel; the tcgen05/TMEM
//                   version is the planned replacement).
//   cls_attention_fwd/bwd : the parallel branch only consumes output row 0 (kwClip.py:1103), and row 0 of its input is
//                   the same [CLS] vector for every utterance, so attention reduces to ONE query per (utterance, head)
//                   against all keys: an HBM-bound kernel that reads K,V once.  Backward yields dK, dV and dq.
#include <cstdlib>

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
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
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

// Same copy with every address computed once per CTA: when 128 threads tile the 64 x CH chunk grid in whole rows
// (128 % CH == 0) thread t owns chunk column c0 = t % CH of rows r0 + i * (128 / CH); both the global pointer and the
// swizzled shared offset then advance by constants.
template <int HD>
struct TileLoader {
  static constexpr int CH = Cfg<HD>::CHUNKS;
  static constexpr bool FAST = (128 % CH) == 0;
  static constexpr int RSTEP = FAST ? 128 / CH : 0;
  static constexpr int NI = FAST ? CH / 2 : 0;  // 64 * CH / 128 chunks per thread (CH = 2 -> 1)
  const uint16_t* g0;   // global pointer of this thread's first chunk at tile row 0
  long long step;       // elements between this thread's consecutive rows (RSTEP * ld)
  long long tile_step;  // elements per 64-row tile
  uint32_t s0;          // shared byte offset of the first chunk inside a tile buffer
  int r0;
  __device__ __forceinline__ void init(const uint16_t* g, long long ld) {
    r0 = threadIdx.x / CH;
    const int c0 = threadIdx.x % CH;
    g0 = g + (long long)r0 * ld + c0 * 8;
    step = (long long)RSTEP * ld;
    tile_step = 64 * ld;
    s0 = (uint32_t)((r0 * Cfg<HD>::PITCH_CHUNKS + ((c0 & ~7) | ((c0 ^ r0) & 7))) * 16);
  }
  // rows row0 .. row0+63 of the tensor -> tile buffer at shared address sbase; rows >= nrows_valid are zero-filled
  __device__ __forceinline__ void load(uint32_t sbase, int tile_idx, int nrows_valid) const {
    const uint16_t* gp = g0 + (long long)tile_idx * tile_step;
    const int row0 = tile_idx * 64 + r0;
#pragma unroll
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) {
      const bool ok = row0 + i * RSTEP < nrows_valid;
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + s0 + (uint32_t)(i * RSTEP * Cfg<HD>::PITCH_CHUNKS * 16)),
                   "l"(ok ? gp + (long long)i * step : g0), "r"(sz) : "memory");
    }
  }
};

template <int HD, bool BF16>
__global__ void __launch_bounds__(128, HD <= 64 ? 4 : 2) attention_fwd_kernel(const AttnParams p) {
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

  constexpr int TILE_BYTES = Cfg<HD>::TILE_ELEMS * 2;
  constexpr int PITCH_B = Cfg<HD>::PITCH_CHUNKS * 16;  // bytes per tile row
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
  TileLoader<HD> ldk, ldv;
  if (TileLoader<HD>::FAST) {
    TileLoader<HD> ldq;
    ldq.init(Q + (long long)q0 * p.q_ld, p.q_ld);
    ldk.init(K, p.k_ld);
    ldv.init(V, p.v_ld);
    ldq.load(sQ_a, 0, p.Tq - q0);
    ldk.load(sK_a, 0, kv_end);
    ldv.load(sV_a, 0, kv_end);
  } else {
    load_tile<HD>(sQ, Q, p.q_ld, q0, p.Tq);
    load_tile<HD>(sK, K, p.k_ld, 0, kv_end);
    load_tile<HD>(sV, V, p.v_ld, 0, kv_end);
  }
  cp_async_commit();

  // ldmatrix addresses: (row, 16-byte chunk) of a tile lives at row*PITCH_B + ((chunk & ~7) | ((chunk ^ row) & 7)) * 16.  For
  // a lane, row & 7 and the low bit of the chunk are fixed, and the unrolled loops add EVEN chunk bases / row bases that are
  // multiples of 8, so the swizzled chunk is (base & ~7) | ((base & 7) ^ x_lane): four per-lane offsets + immediates.
  const int rk_l = (lane & 7) + ((lane >> 4) << 3), xk = ((lane >> 3) & 1) ^ (lane & 7);                    // K (non-transposed)
  const int rv_l = (lane & 7) + (((lane >> 3) & 1) << 3), xv = (lane >> 4) ^ (lane & 7);                    // V (transposed)
  uint32_t kxo[4], vxo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    kxo[j] = (uint32_t)(rk_l * PITCH_B + (((2 * j) ^ xk) << 4));
    vxo[j] = (uint32_t)(rv_l * PITCH_B + (((2 * j) ^ xv) << 4));
  }

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
      if (TileLoader<HD>::FAST) {
        ldk.load(sK_a + ((tile + 1) & 1) * TILE_BYTES, tile + 1, kv_end);
        ldv.load(sV_a + ((tile + 1) & 1) * TILE_BYTES, tile + 1, kv_end);
      } else {
        load_tile<HD>(sK + ((tile + 1) & 1) * Cfg<HD>::TILE_ELEMS, K, p.k_ld, (tile + 1) * BKV, kv_end);
        load_tile<HD>(sV + ((tile + 1) & 1) * Cfg<HD>::TILE_ELEMS, V, p.v_ld, (tile + 1) * BKV, kv_end);
      }
      cp_async_commit();
    }
    if (tile == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(qf[ks], tile_ptr<HD>(sQ, warp * 16 + (lane & 15), ks * 2 + (lane >> 4)));
    }
    uint32_t ka[4], va[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ka[j] = sK_a + (tile & 1) * TILE_BYTES + kxo[j];
      va[j] = sV_a + (tile & 1) * TILE_BYTES + vxo[j];
    }

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
        ldsm_x4(kb, ka[((ks * 2) & 7) >> 1] + (uint32_t)(nb2 * 16 * PITCH_B + ((ks * 2) & ~7) * 16));
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
        ldsm_x4_trans(vb, va[((db2 * 2) & 7) >> 1] + (uint32_t)(ks2 * 16 * PITCH_B + ((db2 * 2) & ~7) * 16));
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
// One CTA (128 threads) per (utterance, head).  q fp32 [heads*hd] shared by all utterances.  kv 16-bit [B, Tk, ld] with K at
// column k_off + h*hd and V at v_off + h*hd.  probs saved [B, heads, Tk] fp32 for backward.  ctx fp32 [B, heads*hd].
//   scores : one KEY per thread (a full hd-long dot product from 16-byte loads; no per-key warp reduction)
//   softmax: block reduction over the <= Tk scores held in shared memory
//   context: one pair of output dims per thread, keys split over 128 / (hd/2) thread groups, coalesced row reads
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // red may still be read from a previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <int HD>
__global__ void __launch_bounds__(128) cls_attention_fwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                const int* __restrict__ kv_len, int Tk, int heads, float scale,
                                                                float* __restrict__ probs, float* __restrict__ ctx32,
                                                                uint16_t* __restrict__ ctx16, int ctx16_fmt, int batch, float drop_p,
                                                                const long long* __restrict__ rng_state, int rng_site) {
  extern __shared__ float sm[];  // [Tk] scores/probs | [HD] q | [128 x 2] context partials | [8] reduction scratch
  float* sc = sm;
  float* sq = sc + Tk;
  float* part = sq + HD;
  float* red = part + 256;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs + h * HD;
  for (int d = tid; d < HD; d += 128) sq[d] = q[h * HD + d] * scale;
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < len; j += 128) {
    const uint16_t* kr = base + (long long)j * kv_ld + k_off;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(kr + c * 8);
      const float2 a = unpack16(kv_fmt, u.x), bb = unpack16(kv_fmt, u.y), cc = unpack16(kv_fmt, u.z), dd = unpack16(kv_fmt, u.w);
      const float* qq = sq + c * 8;
      dot += a.x * qq[0] + a.y * qq[1] + bb.x * qq[2] + bb.y * qq[3] + cc.x * qq[4] + cc.y * qq[5] + dd.x * qq[6] + dd.y * qq[7];
    }
    sc[j] = dot;
    mx = fmaxf(mx, dot);
  }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int j = tid; j < len; j += 128) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);  // (its barriers also publish sc[])
  const float inv = 1.f / sum;
  float* po = probs + ((long long)b * heads + h) * Tk;
  for (int j = tid; j < Tk; j += 128) po[j] = j < len ? sc[j] * inv : 0.f;   // the UNdropped probabilities (softmax backward)
  if (drop_p > 0.f) {  // attention dropout (nn.MultiheadAttention, train mode): the context uses p_j * m_j, m_j in {0, 1/(1-p)}
    const DropoutRng rng(rng_state, rng_site, drop_p);
    for (int j = tid; j < len; j += 128) sc[j] *= rng.scale(((unsigned long long)b * heads + h) * Tk + j);
    __syncthreads();
  }
  constexpr int P = HD / 2;                    // dim pairs
  constexpr int G = P >= 128 ? 1 : 128 / P;    // key groups
  const int pr = tid % P, grp = tid / P;
  float2 acc = make_float2(0.f, 0.f);
  if (grp < G) {
    const uint16_t* vr = base + v_off + pr * 2;
#pragma unroll 4
    for (int j = grp; j < len; j += G) {
      const float2 vv = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(vr + (long long)j * kv_ld));
      const float pj = sc[j];
      acc.x += pj * vv.x;
      acc.y += pj * vv.y;
    }
  }
  part[tid * 2] = acc.x;
  part[tid * 2 + 1] = acc.y;
  __syncthreads();
  if (tid < P) {
    float2 r = make_float2(0.f, 0.f);
    for (int g2 = 0; g2 < G; ++g2) {
      r.x += part[(g2 * P + tid) * 2];
      r.y += part[(g2 * P + tid) * 2 + 1];
    }
    r.x *= inv;
    r.y *= inv;
    const long long off = (long long)b * heads * HD + h * HD + tid * 2;
    if (ctx32) *reinterpret_cast<float2*>(ctx32 + off) = r;
    if (ctx16) *reinterpret_cast<uint32_t*>(ctx16 + off) = pack16(ctx16_fmt, r.x, r.y);
  }
}

// Backward of the single-query attention.  dctx fp32 [B, heads*hd].  Writes dKV 16-bit [B, Tk, ld] (K grads at k_off, V grads
// at v_off; rows >= len are zero) and accumulates dq (gradient of the UNSCALED q, fp32 [heads*hd]) with atomics.
template <int HD>
__global__ void __launch_bounds__(128) cls_attention_bwd_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kv, int kv_fmt,
                                                                long long kv_ld, long long kv_bs, int k_off, int v_off,
                                                                const int* __restrict__ kv_len, int Tk, int heads, float scale,
                                                                const float* __restrict__ probs, const float* __restrict__ dctx,
                                                                uint16_t* __restrict__ dkv, int dkv_fmt, float* __restrict__ dq, int batch,
                                                                float drop_p, const long long* __restrict__ rng_state, int rng_site) {
  extern __shared__ float sm[];  // [Tk] ds | [Tk] p | [Tk] p*m | [HD] q | [HD] dctx | [256] partials | [8] scratch
  float* sds = sm;
  float* sp = sds + Tk;
  float* spm = sp + Tk;   // p_j * m_j (dropout mask regenerated from the RNG state of the forward; = p_j without dropout)
  float* sq = spm + Tk;
  float* sdc = sq + HD;
  float* part = sdc + HD;
  float* red = part + 256;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const uint16_t* base = kv + (long long)b * kv_bs + h * HD;
  uint16_t* dbase = dkv + (long long)b * kv_bs + h * HD;
  const DropoutRng rng(rng_state, rng_site, drop_p);
  const float* pr_g = probs + ((long long)b * heads + h) * Tk;
  for (int d = tid; d < HD; d += 128) {
    sq[d] = q[h * HD + d];
    sdc[d] = dctx[(long long)b * heads * HD + h * HD + d];
  }
  __syncthreads();
  // dp_j = <dctx, v_j> (one key per thread);  dot = sum_j p_j dp_j
  float dot = 0.f;
  for (int j = tid; j < len; j += 128) {
    const uint16_t* vr = base + (long long)j * kv_ld + v_off;
    float dp = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(vr + c * 8);
      const float2 a = unpack16(kv_fmt, u.x), bb = unpack16(kv_fmt, u.y), cc = unpack16(kv_fmt, u.z), dd = unpack16(kv_fmt, u.w);
      const float* g = sdc + c * 8;
      dp += a.x * g[0] + a.y * g[1] + bb.x * g[2] + bb.y * g[3] + cc.x * g[4] + cc.y * g[5] + dd.x * g[6] + dd.y * g[7];
    }
    const float pj = pr_g[j];
    const float mj = drop_p > 0.f ? rng.scale(((unsigned long long)b * heads + h) * Tk + j) : 1.f;
    dp *= mj;            // d ctx / d p_j = m_j v_j
    sp[j] = pj;
    spm[j] = pj * mj;
    sds[j] = dp;
    dot += pj * dp;
  }
  dot = block_reduce(dot, red, false);
  for (int j = tid; j < len; j += 128) sds[j] = sp[j] * (sds[j] - dot) * scale;  // d(score_j) * scale: score = scale * <q, k_j>
  __syncthreads();
  // dK_j = ds_j * q, dV_j = p_j * dctx : one 16-byte chunk (8 dims) of one key row per thread per step, coalesced
  constexpr int CH = HD / 8;
  for (int idx = tid; idx < Tk * CH; idx += 128) {
    const int j = idx / CH, c = idx % CH;
    uint4 uk = make_uint4(0u, 0u, 0u, 0u), uv = uk;
    if (j < len) {
      const float ds = sds[j], pj = spm[j];
      const float* qq = sq + c * 8;
      const float* g = sdc + c * 8;
      uk.x = pack16(dkv_fmt, ds * qq[0], ds * qq[1]); uk.y = pack16(dkv_fmt, ds * qq[2], ds * qq[3]);
      uk.z = pack16(dkv_fmt, ds * qq[4], ds * qq[5]); uk.w = pack16(dkv_fmt, ds * qq[6], ds * qq[7]);
      uv.x = pack16(dkv_fmt, pj * g[0], pj * g[1]); uv.y = pack16(dkv_fmt, pj * g[2], pj * g[3]);
      uv.z = pack16(dkv_fmt, pj * g[4], pj * g[5]); uv.w = pack16(dkv_fmt, pj * g[6], pj * g[7]);
    }
    *reinterpret_cast<uint4*>(dbase + (long long)j * kv_ld + k_off + c * 8) = uk;
    *reinterpret_cast<uint4*>(dbase + (long long)j * kv_ld + v_off + c * 8) = uv;
  }
  // dq += sum_j ds_j k_j : a pair of dims per thread, keys split over thread groups
  constexpr int P = HD / 2;
  constexpr int G = P >= 128 ? 1 : 128 / P;
  const int pr = tid % P, grp = tid / P;
  float2 acc = make_float2(0.f, 0.f);
  if (grp < G) {
    const uint16_t* kr = base + k_off + pr * 2;
#pragma unroll 4
    for (int j = grp; j < len; j += G) {
      const float2 kk = unpack16(kv_fmt, *reinterpret_cast<const uint32_t*>(kr + (long long)j * kv_ld));
      const float ds = sds[j];
      acc.x += ds * kk.x;
      acc.y += ds * kk.y;
    }
  }
  part[tid * 2] = acc.x;
  part[tid * 2 + 1] = acc.y;
  __syncthreads();
  if (tid < P) {
    float2 r = make_float2(0.f, 0.f);
    for (int g2 = 0; g2 < G; ++g2) {
      r.x += part[(g2 * P + tid) * 2];
      r.y += part[(g2 * P + tid) * 2 + 1];
    }
    atomicAdd(&dq[h * HD + tid * 2], r.x);
    atomicAdd(&dq[h * HD + tid * 2 + 1], r.y);
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
  {  // tcgen05 / TMEM kernel for head_dim 64 and <= 320 keys (all tower attentions); SCB_ATTN_TC=0 forces the mma.sync kernel
    static const int tc_env = [] { const char* e = getenv("SCB_ATTN_TC"); return e ? atoi(e) : 1; }();
    if (tc_env) {
      const int e = attention_fwd_tc(q, k, v, o, fmt, q_ld, k_ld, v_ld, o_ld, q_bs, k_bs, v_bs, o_bs, kv_len, batch, heads, head_dim, Tq, Tk,
                                     scale, causal, st);
      if (e != SCB_EUNSUPPORTED) return e;
    }
  }
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
                      int ctx16_fmt, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && (ctx32 || ctx16), SCB_EINVAL, "scb_cls_attention_fwd: null operand");
  SCB_CHECK(drop_p == 0.f || (drop_p > 0.f && drop_p < 1.f && rng_state), SCB_EINVAL, "scb_cls_attention_fwd: dropout needs p in [0,1) and an rng_state");
  if (batch == 0) return SCB_OK;
  const size_t smem = (size_t)(Tk + head_dim + 256 + 8) * sizeof(float);
  SCB_CHECK(smem <= 48 * 1024, SCB_EUNSUPPORTED, "scb_cls_attention_fwd: Tk=%d too long", Tk);
  SCB_CHECK(kv_ld % 8 == 0 && kv_bs % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && head_dim % 8 == 0, SCB_EINVAL,
            "scb_cls_attention_fwd: kv strides / offsets / head_dim must be multiples of 8 elements");
  const unsigned grid = (unsigned)(batch * heads);
  SCB_HD_SWITCH(head_dim, (cls_attention_fwd_kernel<HD_><<<grid, 128, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads, scale, probs, ctx32, (uint16_t*)ctx16, ctx16_fmt, batch, drop_p, rng_state, rng_site)));
  note_launch();
  SCB_LAUNCH_OK("cls_attention_fwd");
  return SCB_OK;
}

int cls_attention_bwd(const float* q, const void* kv, int kv_fmt, long long kv_ld, long long kv_bs, int k_off, int v_off,
                      const int* kv_len, int batch, int heads, int head_dim, int Tk, float scale, const float* probs, const float* dctx,
                      void* dkv, int dkv_fmt, float* dq, float drop_p, const long long* rng_state, int rng_site, cudaStream_t st) {
  SCB_CHECK(q && kv && probs && dctx && dkv && dq, SCB_EINVAL, "scb_cls_attention_bwd: null operand");
  SCB_CHECK(drop_p == 0.f || (drop_p > 0.f && drop_p < 1.f && rng_state), SCB_EINVAL, "scb_cls_attention_bwd: dropout needs p in [0,1) and an rng_state");
  if (batch == 0) return SCB_OK;
  const size_t smem = (size_t)(3 * Tk + 2 * head_dim + 256 + 8) * sizeof(float);
  SCB_CHECK(smem <= 48 * 1024, SCB_EUNSUPPORTED, "scb_cls_attention_bwd: Tk=%d too long", Tk);
  SCB_CHECK(kv_ld % 8 == 0 && kv_bs % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && head_dim % 8 == 0, SCB_EINVAL,
            "scb_cls_attention_bwd: kv strides / offsets / head_dim must be multiples of 8 elements");
  const unsigned grid = (unsigned)(batch * heads);
  SCB_HD_SWITCH(head_dim, (cls_attention_bwd_kernel<HD_><<<grid, 128, smem, st>>>(q, (const uint16_t*)kv, kv_fmt, kv_ld, kv_bs, k_off, v_off, kv_len, Tk, heads, scale, probs, dctx, (uint16_t*)dkv, dkv_fmt, dq, batch, drop_p, rng_state, rng_site)));
  note_launch();
  SCB_LAUNCH_OK("cls_attention_bwd");
  return SCB_OK;
}

}  // namespace scb
