// Attention on the 5th-generation tensor cores (tcgen05 + TMEM) for head_dim 64 and up to 320 keys — every attention of the
// frozen towers (HuBERT 319 frames, CLIP ViT 50 / 257 tokens, CLIP text 77 tokens).
//
// One persistent CTA per SM walks (utterance, head) items.  Per item the producer thread TMA-loads K [NK x 64], V [NK x 64]
// and all Q tiles [128 x 64] straight out of the fused QKV activation (three 3-D tensor maps), 128B-swizzled.  Per 128-query
// tile:
//   MMA thread      S = Q K^T           tcgen05.mma 128 x NK x 64 (two instructions of N = NK/2 when NK > 256) -> TMEM cols [0, NK)
//   16 softmax warps row max / exp2 / row sum straight from TMEM (tcgen05.ld; the whole key axis of a row is resident, so there is
//                   no online rescaling), key-padding / causal masks, P (fp16/bf16) -> shared memory in the K-major swizzled
//                   layout the next MMA reads as its A operand
//   MMA thread      O = P V             tcgen05.mma 128 x 64 x NK, V read as the MN-major B operand directly from its row-major
//                   tile (no transposed copy of V anywhere) -> TMEM cols [320, 384)
//   16 softmax warps O / rowsum -> 16-bit -> global
// The S MMAs of tile i+1 are issued right behind the PV MMAs of tile i, so they run while the softmax warps drain O.
#include <cstdlib>

#include "common.cuh"
#include "ops.cuh"

namespace scb {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int MAX_NK = 320;
constexpr int MAX_QT = 3;                       // query tiles resident per item (Tq <= 384)
constexpr int O_COL = 320;                      // TMEM column of the O accumulator
constexpr int kSoftmaxWarps = 16;                // 4 per TMEM lane quarter, each owning NK/4 key columns
constexpr int kThreads = (2 + kSoftmaxWarps) * 32;  // warp 0: TMA producer (+ TMEM alloc), warp 1: MMA issuer, then the softmax warps
constexpr int K_BYTES = MAX_NK * 128;           // 40 KB
constexpr int V_BYTES = MAX_NK * 128;           // 40 KB (5 key blocks of 64 rows x 128 B)
constexpr int Q_BYTES = MAX_QT * BQ * 128;      // 48 KB
constexpr int P_BYTES = (MAX_NK / 64) * BQ * 128;  // 80 KB (5 k-blocks of [128 x 64])
constexpr int SMEM_BYTES = K_BYTES + V_BYTES + Q_BYTES + P_BYTES + 2 * 4 * BQ * 4 + 256 + 1024;  // 256: mbarriers + TMEM slot

struct AttnTcParams {
  uint16_t* o;
  long long o_ld, o_bs;
  const int* kv_len;
  int batch, heads, Tq, Tk, NK, n_qt;
  float scale_log2;
  int causal, bf16;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
__device__ __forceinline__ uint32_t idesc_pv(int fmt) { return umma_idesc_f16(BQ, HD, fmt) | (1u << 16); }

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + K_BYTES;
  uint8_t* sQ = sV + V_BYTES;
  uint8_t* sP = sQ + Q_BYTES;
  float* red_max = reinterpret_cast<float*>(sP + P_BYTES);  // [4 column parts][128 rows]
  float* red_sum = red_max + 4 * BQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_sum + 4 * BQ);
  uint64_t* kq_full = bars + 0;    // K and every Q tile of the item have landed
  uint64_t* item_empty = bars + 1; // every PV MMA of the item has retired: V may be overwritten
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint64_t* o_empty = bars + 5;
  uint64_t* v_full = bars + 6;     // V of the item has landed
  uint64_t* kq_empty = bars + 7;   // every S MMA of the item has retired: K and Q may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kq_full, 1);
    mbar_init(item_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(kq_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, kSoftmaxWarps);
    mbar_init(o_full, 1);
    mbar_init(o_empty, kSoftmaxWarps);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();               // programmatic dependent launch: global memory is only touched below
  griddep_launch_dependents();
  const int n_items = p.batch * p.heads;
  const int NK = p.NK;
  const int nkb = NK / 64;  // key blocks of the PV contraction

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ producer
    // K and Q of the NEXT item are loaded as soon as the last S = Q K^T of the current item has retired (kq_empty), i.e. while
    // its last query tile is still in the softmax / P V / drain phases; V follows once the last P V has retired (item_empty).
    // Without this split every item started with a ~2 us load bubble (14 % of the kernel at 319 frames).
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.heads, h = item % p.heads;
      mbar_wait(kq_empty, ph ^ 1u);
      mbar_expect_tx(kq_full, (uint32_t)(NK * 128 + p.n_qt * BQ * 128));
      if (NK <= 256) {
        tma_load_3d(sK, &tmK, kq_full, h * HD, 0, b);
      } else {
        tma_load_3d(sK, &tmK, kq_full, h * HD, 0, b);
        tma_load_3d(sK + (NK / 2) * 128, &tmK, kq_full, h * HD, NK / 2, b);
      }
      for (int qt = 0; qt < p.n_qt; ++qt) tma_load_3d(sQ + qt * BQ * 128, &tmQ, kq_full, h * HD, qt * BQ, b);
      mbar_wait(item_empty, ph ^ 1u);
      mbar_expect_tx(v_full, (uint32_t)(NK * 128));
      for (int kb = 0; kb < nkb; ++kb) tma_load_3d(sV + kb * 64 * 128, &tmV, v_full, h * HD, kb * 64, b);
      ph ^= 1u;
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const int fmt = BF16 ? 1 : 0;
    const int n1 = NK <= 256 ? NK : NK / 2;
    const uint32_t idesc_s = umma_idesc_f16(BQ, n1, fmt);
    const uint32_t idesc_o = idesc_pv(fmt);
    uint32_t ph_item = 0, ph_p = 0, ph_oe = 0;
    auto issue_s = [&](int qt) {
      const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sQ + qt * BQ * 128));
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        tc_mma_f16(tmem_base, a_desc + (uint64_t)(2 * k), umma_desc_kmajor_sw128(smem_u32(sK)) + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
        if (NK > 256)
          tc_mma_f16(tmem_base + (uint32_t)n1, a_desc + (uint64_t)(2 * k),
                     umma_desc_kmajor_sw128(smem_u32(sK + n1 * 128)) + (uint64_t)(2 * k), idesc_s, (uint32_t)(k != 0));
      }
      tc_commit(s_full);
    };
    // S(0) of an item is issued as soon as its K / Q have landed AND the S columns are free, i.e. right behind the P V MMAs of
    // the previous item's last tile (same slot S(qt+1) takes inside an item).
    bool first = true;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      if (first) {
        mbar_wait(kq_full, ph_item);
        tc_fence_after();
        issue_s(0);
        if (p.n_qt == 1) tc_commit(kq_empty);
        first = false;
      }
      const bool has_next = item + (int)gridDim.x < n_items;
      for (int qt = 0; qt < p.n_qt; ++qt) {
        mbar_wait(p_full, ph_p);       // P(qt) is in shared memory and S(qt) has been read
        ph_p ^= 1u;
        tc_fence_after();
        // S of the next tile goes first: the softmax warps start on it while the PV MMAs below are still running
        if (qt + 1 < p.n_qt) {
          issue_s(qt + 1);
          if (qt + 2 == p.n_qt) tc_commit(kq_empty);  // the item's last S: K / Q are free once it retires
        }
        if (qt == 0) {
          mbar_wait(v_full, ph_item);
          tc_fence_after();
        }
        mbar_wait(o_empty, ph_oe ^ 1u);  // O of the previous tile has been read out of TMEM
        ph_oe ^= 1u;
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sP + kb * BQ * 128));
          const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sV + kb * 64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 keys per instruction: A advances 32 B inside its row, B (MN-major) by 16 rows
            tc_mma_f16(tmem_base + O_COL, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(k * ((16 * 128) >> 4)), idesc_o,
                       (uint32_t)((kb | k) != 0));
        }
        tc_commit(o_full);
        if (qt + 1 == p.n_qt) {
          tc_commit(item_empty);
          if (has_next) {  // next item's first S behind this item's last P V
            mbar_wait(kq_full, ph_item ^ 1u);
            tc_fence_after();
            issue_s(0);
            if (p.n_qt == 1) tc_commit(kq_empty);
          }
        }
      }
      ph_item ^= 1u;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ softmax + output
    const int q = warp & 3, part = (warp - 2) >> 2;  // TMEM lane quarter; which quarter of the key columns
    const int r = q * 32 + lane;                       // row of the query tile owned by this lane
    const int part_cols = NK / 4;                      // multiple of 16
    const int nchunk = part_cols / 16;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ph_s = 0, ph_o = 0;
    // O / rowsum of one query tile -> global (its P V MMAs run while the softmax of the NEXT tile is in flight)
    auto drain_o = [&](int b, int h, int row_g, float inv) {
      mbar_wait(o_full, ph_o);
      ph_o ^= 1u;
      tc_fence_after();
      uint32_t ov[16];
      tmem_ld_32x16(lane_addr + (uint32_t)(O_COL + part * 16), ov);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (row_g < p.Tq) {
        uint16_t* dst = p.o + (long long)b * p.o_bs + (long long)row_g * p.o_ld + h * HD + part * 16;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint4 u;
          uint32_t* uu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            uu[i] = H16<BF16 ? SCB_BF16 : SCB_F16>::pack(__uint_as_float(ov[t * 8 + 2 * i]) * inv, __uint_as_float(ov[t * 8 + 2 * i + 1]) * inv);
          *reinterpret_cast<uint4*>(dst + t * 8) = u;
        }
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.heads, h = item % p.heads;
      const int kvl = p.kv_len ? min(p.kv_len[b], p.Tk) : p.Tk;
      float inv_prev = 0.f;
      for (int qt = 0; qt < p.n_qt; ++qt) {
        const int row_g = qt * BQ + r;      // query index
        const int key_hi = p.causal ? min(kvl, row_g + 1) : kvl;  // keys [0, key_hi) are visible to this row
        mbar_wait(s_full, ph_s);
        ph_s ^= 1u;
        tc_fence_after();
        // ---- pass 1: row max over this warp's quarter of the keys (masking only on chunks that straddle key_hi).
        //      (Keeping the 80 S values of a lane in registers instead of re-reading TMEM in pass 2 was measured SLOWER:
        //      408 vs 310 us per B=256 layer call — the kernel is bound by barrier / MMA / MUFU latency chains, not by TMEM
        //      read bandwidth: ncu shows tensor pipe 14 %, XU 28 %, LDTM 3 % busy.)
        float mx = -INFINITY;
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[16];
          const int key0 = part * part_cols + c * 16;
          tmem_ld_32x16(lane_addr + (uint32_t)key0, v);
          tmem_ld_wait();
          if (__all_sync(0xffffffffu, key0 + 16 <= key_hi)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (key0 + i < key_hi) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        red_max[part * BQ + r] = mx;
        named_bar_sync(1 + q, 128);
        mx = fmaxf(fmaxf(red_max[r], red_max[BQ + r]), fmaxf(red_max[2 * BQ + r], red_max[3 * BQ + r]));
        const float m_off = (mx == -INFINITY) ? 0.f : -mx * p.scale_log2;
        // ---- the previous tile's O: its PV MMAs read P, which pass 2 below overwrites
        if (qt > 0) drain_o(b, h, row_g - BQ, inv_prev);
        // ---- pass 2: p = 2^(s*scale - m*scale), row sum, P -> shared memory (K-major, 128B-swizzled A operand)
        float sum = 0.f;
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[16];
          const int key0 = part * part_cols + c * 16;
          tmem_ld_32x16(lane_addr + (uint32_t)key0, v);
          tmem_ld_wait();
          float pv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, m_off));
          if (!__all_sync(0xffffffffu, key0 + 16 <= key_hi)) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (key0 + i >= key_hi) pv[i] = 0.f;
          }
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sum += pv[2 * i] + pv[2 * i + 1];
            pk[i] = H16<BF16 ? SCB_BF16 : SCB_F16>::pack(pv[2 * i], pv[2 * i + 1]);
          }
          uint8_t* prow = sP + (key0 >> 6) * (BQ * 128) + r * 128;
          const int c8 = (key0 & 63) >> 3;  // first 16-byte unit of this chunk inside the 128-byte row (0, 2, 4 or 6)
          *reinterpret_cast<uint4*>(prow + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(prow + (((c8 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        red_sum[part * BQ + r] = sum;
        tc_fence_before();
        fence_proxy_async();  // generic-proxy writes of P become visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        named_bar_sync(1 + q, 128);
        sum = (red_sum[r] + red_sum[BQ + r]) + (red_sum[2 * BQ + r] + red_sum[3 * BQ + r]);
        inv_prev = sum > 0.f ? 1.f / sum : 0.f;
      }
      drain_o(b, h, (p.n_qt - 1) * BQ + r, inv_prev);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Returns SCB_EUNSUPPORTED (without setting an error) when the shape is outside this kernel's envelope: the caller then uses the
// mma.sync kernel.
int attention_fwd_tc(const void* q, const void* k, const void* v, void* o, int fmt, long long q_ld, long long k_ld, long long v_ld,
                     long long o_ld, long long q_bs, long long k_bs, long long v_bs, long long o_bs, const int* kv_len, int batch, int heads,
                     int head_dim, int Tq, int Tk, float scale, int causal, cudaStream_t st) {
  if (head_dim != HD || Tk > MAX_NK || Tq > MAX_QT * BQ || Tk < 1 || Tq < 1) return SCB_EUNSUPPORTED;
  // short sequences (CLIP ViT-B/32: 50 tokens, text: 77) leave the 128-row MMA tiles mostly empty: measured faster on the
  // mma.sync kernel (64-row tiles, 4-5 CTAs per SM); SCB_ATTN_TC=2 forces this kernel for every supported shape
  static const int force = [] { const char* e = getenv("SCB_ATTN_TC"); return e ? atoi(e) : 1; }();
  if (Tk < 192 && force != 2) return SCB_EUNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) != 0) return SCB_EUNSUPPORTED;
  AttnTcParams p{};
  p.o = (uint16_t*)o;
  p.o_ld = o_ld;
  p.o_bs = o_bs;
  p.kv_len = kv_len;
  p.batch = batch;
  p.heads = heads;
  p.Tq = Tq;
  p.Tk = Tk;
  p.NK = (Tk + 63) / 64 * 64;
  p.n_qt = (Tq + BQ - 1) / BQ;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.bf16 = fmt == SCB_BF16;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t dq[3] = {(uint64_t)heads * HD, (uint64_t)Tq, (uint64_t)batch};
  const uint64_t dk[3] = {(uint64_t)heads * HD, (uint64_t)Tk, (uint64_t)batch};
  const uint64_t sq[2] = {(uint64_t)q_ld * 2, (uint64_t)q_bs * 2}, sk[2] = {(uint64_t)k_ld * 2, (uint64_t)k_bs * 2},
                 sv[2] = {(uint64_t)v_ld * 2, (uint64_t)v_bs * 2};
  const uint32_t bq[3] = {64, BQ, 1};
  const uint32_t bk[3] = {64, (uint32_t)(p.NK <= 256 ? p.NK : p.NK / 2), 1};
  const uint32_t bv[3] = {64, 64, 1};
  int e = make_tmap(&tmQ, q, 2, 3, dq, sq, bq, 1);
  if (!e) e = make_tmap(&tmK, k, 2, 3, dk, sk, bk, 1);
  if (!e) e = make_tmap(&tmV, v, 2, 3, dk, sv, bv, 1);
  if (e) return e;
  static bool configured = false;
  if (!configured) {
    SCB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SCB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int items = batch * heads;
  const int grid = items < num_sms() ? items : num_sms();
  if (p.bf16) SCB_CUDA(launch_pdl(attention_tc_kernel<true>, dim3((unsigned)grid), kThreads, SMEM_BYTES, st, tmQ, tmK, tmV, p));
  else SCB_CUDA(launch_pdl(attention_tc_kernel<false>, dim3((unsigned)grid), kThreads, SMEM_BYTES, st, tmQ, tmK, tmV, p));
  note_launch();
  SCB_LAUNCH_OK("attention_tc");
  return SCB_OK;
}

}  // namespace scb
