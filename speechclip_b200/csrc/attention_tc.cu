// Attention on the 5th-generation tensor cores (tcgen05 + TMEM) for head_dim 64 and up to 320 keys — every attention of the
// frozen towers (HuBERT 319 frames, CLIP ViT 50 / 257 tokens, CLIP text 77 tokens).
//
// One persistent CTA per SM walks (utterance, head) items.  Per item the producer thread TMA-loads K [NK x 64], V [NK x 64]
// and all Q tiles [128 x 64] straight out of the fused QKV activation (three 3-D tensor maps), 128B-swizzled.
//
// A 128-query tile is processed as TWO key halves of H = NK / 2 columns, each with its own S accumulator in TMEM
// (columns [0, H) and [H, 2H)); O lives in columns [320, 384):
//   MMA thread       S_h = Q K_h^T        tcgen05.mma 128 x H x 64                          -> s_full[h]
//   16 softmax warps two passes over their H/4-column slice of S_h straight from TMEM (tcgen05.ld): row max (exchanged between the
//                    4 warps of a lane quarter), then exp2 (scale / offset on packed fma.f32x2) and 16-bit P into the K-major
//                    128B-swizzled shared-memory operand                                       -> s_free[h], p_full[h]
//   MMA thread       O (+)= P_h V_h        tcgen05.mma 128 x 64 x H, V as the MN-major B operand (no transposed copy of V)
//                    ... as ONE 128 x 80 x 16 MMA per key step: the B operand's second 64-column atom (descriptor LBO) is a tile
//                    of ones, so columns [64, 80) of the accumulator are the ROW SUMS of exactly the 16-bit P that multiplies
//                    V — the softmax warps neither add up their exponentials nor exchange partial sums (2 of their ~5
//                    instructions per element), and P is read from shared memory once
// so that while the softmax warps work on one half the tensor pipe computes S of the other half / of the next query tile and the
// P V of the previous half: the softmax warps never wait for an MMA in steady state, and the kernel runs at the rate of its
// exponentials (MUFU) instead of the sum of all phases (round 1: one 320-column S, every phase serial, 17 % tensor pipe).
// The two halves share one running row maximum: half 1 keeps the maximum of half 0 unless its own maximum is more than 2^8 above
// it (P <= 256 is harmless in 16 bits; softmax is shift-invariant), in which case — rare — the warps rescale O and the row
// sums in place (tcgen05.ld / st) before P V of half 1 is issued.
// O of a tile is drained (divided by the row sum, packed, stored) inside the NEXT tile, also across items.  Query quarters
// (32 rows) that lie entirely beyond Tq skip all softmax work (T = 319: 2 of the 12 quarters of an item).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ops.cuh"

namespace scb {
#ifdef SCB_ATTN_TRACE
__device__ long long g_attn_trace[64 * 16];
__device__ long long g_attn_trace_w[32 * 16 * 4];   // [tile][softmax warp][s_full passed / p_full arrive, per half]
#define ATTN_TRACE(cond, tile, slot) do { if ((cond) && blockIdx.x == 0 && (tile) < 64 && (tile) >= 0) g_attn_trace[(tile) * 16 + (slot)] = clock64(); } while (0)
#define ATTN_TRACE_W(tile, w, slot) do { if (blockIdx.x == 0 && (tile) < 32 && (threadIdx.x & 31) == 0) g_attn_trace_w[((tile) * 16 + (w)) * 4 + (slot)] = clock64(); } while (0)
extern "C" int scb_debug_attn_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(g_attn_trace));
}
extern "C" int scb_debug_attn_trace_w(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_attn_trace_w, sizeof(g_attn_trace_w));
}
#else
#define ATTN_TRACE(cond, tile, slot) do { } while (0)
#define ATTN_TRACE_W(tile, w, slot) do { } while (0)
#endif
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int MAX_NK = 320;
constexpr int MAX_QT = 3;                       // query tiles resident per item (Tq <= 384)
constexpr int O_COL = 320;                      // TMEM column of the O accumulator
constexpr int L_COL = O_COL + HD;               // TMEM columns [384, 400): row sums of P (every column holds the same sum)
constexpr int ONES_BYTES = 64 * 128;            // second MN atom of the P V B operand: 64 key rows x 128 B of ones (all elements equal, so
                                                // the swizzle does not matter); reached from every V block through the descriptor's LBO
constexpr int kSoftmaxWarps = 16;                // 4 per TMEM lane quarter, each owning H/4 key columns of a half
// Warp roles.  A warp's scheduler is warp % 4 and so is the TMEM lane quarter it may read, so the 4 softmax warps of a row quarter
// share one scheduler with each other and with whatever control warp has the same index mod 4.  Measured (per-warp trace): the
// quarter that shares its scheduler with the MMA-issuing warp runs ~800 clocks per tile (11 %) behind the others — the 28
// tcgen05.mma / commit dispatches and barrier polls of a tile are not free for their scheduler — and the slowest quarter sets the
// pace (P V waits for all 16 warps).  The control warps therefore sit on schedulers 2 and 3: the quarters 2 and 3 have no rows in
// the last, partial query tile of an item (T = 319: 63 rows), i.e. a third less softmax work than quarters 0 and 1.
constexpr int kMmaWarp = 18, kProducerWarp = 19;   // warps 16 and 17 are idle fillers (exit at once)
constexpr int kThreads = 20 * 32;   // (registers are allocated per 4 warps: 18 warps get the same 96-register cap as 20)
constexpr int K_BYTES = MAX_NK * 128;           // 40 KB
constexpr int V_BYTES = MAX_NK * 128;           // 40 KB (5 key blocks of 64 rows x 128 B)
constexpr int Q_BYTES = MAX_QT * BQ * 128;      // 48 KB
constexpr int P_BYTES = (MAX_NK / 64) * BQ * 128;  // 80 KB (5 k-blocks of [128 x 64]; half h owns the keys [h H, h H + H))
constexpr int RED_FLOATS = 2 * 4 * BQ;          // row max of half 0 / half 1: [4 column parts][128 rows] each
constexpr int COLD_BYTES = kSoftmaxWarps * 32 * 8;    // per softmax thread: the pending tile's output row pointer (bit 63: quarter is live)
constexpr int SMEM_BYTES = K_BYTES + V_BYTES + Q_BYTES + P_BYTES + ONES_BYTES + RED_FLOATS * 4 + COLD_BYTES + 256 + 1024;  // 256: mbarriers + TMEM slot
constexpr float kRescaleLog2 = 8.f;             // half 1 keeps half 0's row maximum unless its own is > 2^8 above it

struct AttnTcParams {
  uint16_t* o;
  long long o_ld, o_bs;
  const int* kv_len;
  int batch, heads, Tq, Tk, NK, n_qt;
  float scale_log2;
  int causal, bf16;
};

// Bounded wait WITHOUT a printf: a call in the hot loops makes the compiler keep loop state in stack slots, and with the shared-memory
// carve-out at its maximum the L1 behind the stack is ~28 KB — the reloads then cost L2 latency on the MMA issue path.
__device__ __forceinline__ bool try_wait_hint(uint64_t* bar, uint32_t parity) {
  // suspend-time hint (ns): the hardware parks the thread until the phase completes (or the hint expires) instead of returning
  // early — a parked waiter takes no issue slots from the softmax warps that share its scheduler
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  if (try_wait_hint(bar, parity)) return;
  const long long t0 = clock64();
  while (!try_wait_hint(bar, parity))
    if (clock64() - t0 > 8000000000LL) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
// (N = 64 head dims + 16 columns of row sums)
__device__ __forceinline__ uint32_t idesc_pv(int fmt) { return umma_idesc_f16(BQ, HD + 16, fmt) | (1u << 16); }
// MN-major SWIZZLE_128B operand whose MN extent exceeds one 64-element atom: LBO = byte distance to the next atom along MN
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = umma_desc_kmajor_sw128(smem_addr) & ~((uint64_t)0x3FFF << 16);
  return d | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16);
}

// NCH = NK / 64 (1..5): compile-time so that every S slice has a static register footprint
template <bool BF16, int NCH>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const AttnTcParams p) {
  constexpr int NK = NCH * 64;
  constexpr int H = NK / 2;       // keys per half: 32 .. 160 (multiple of 32)
  constexpr int CW = H / 4;       // key columns of a half owned by one softmax warp: 8 .. 40
  constexpr int NCK = CW / 8;     // ... in chunks of 8 (one 16-byte unit of a P row)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + K_BYTES;
  uint8_t* sQ = sV + V_BYTES;
  uint8_t* sP = sQ + Q_BYTES;
  uint8_t* sOnes = sP + P_BYTES;                             // 1024-aligned (every region above is a multiple of 1 KB)
  float* red_max = reinterpret_cast<float*>(sOnes + ONES_BYTES);  // [half][4 column parts][128 rows]
  uint2* cold = reinterpret_cast<uint2*>(red_max + 2 * 4 * BQ);   // [softmax thread]
  uint64_t* bars = reinterpret_cast<uint64_t*>(cold + kSoftmaxWarps * 32);
  uint64_t* kq_full = bars + 0;     // K and every Q tile of the item have landed
  uint64_t* kq_empty = bars + 1;    // every S MMA of the item has retired: K and Q may be overwritten
  uint64_t* v_full = bars + 2;      // V of the item has landed
  uint64_t* item_empty = bars + 3;  // every PV MMA of the item has retired: V may be overwritten
  uint64_t* s_full = bars + 4;      // [2] S_h of the tile is in TMEM
  uint64_t* s_free = bars + 6;      // [2] every softmax warp holds its slice of S_h in registers
  uint64_t* p_full = bars + 8;      // [2] P_h of the tile is in shared memory
  uint64_t* pv_done = bars + 10;    // [2] P_h V_h (and every earlier MMA) has retired; pv_done[1] = O of the tile is complete
  uint64_t* o_empty = bars + 12;    // O of the tile has been read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == kMmaWarp && lane == 0) {
    mbar_init(kq_full, 1);
    mbar_init(kq_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(item_empty, 1);
    for (int h = 0; h < 2; ++h) {
      mbar_init(s_full + h, 1);
      mbar_init(s_free + h, kSoftmaxWarps);
      mbar_init(p_full + h, kSoftmaxWarps);
      mbar_init(pv_done + h, 1);
    }
    mbar_init(o_empty, kSoftmaxWarps);
    mbar_fence_init();
  }
  if (warp == kProducerWarp) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += kThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = BF16 ? 0x3F803F80u : 0x3C003C00u;
  fence_proxy_async();  // the tensor core reads the tile through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();               // programmatic dependent launch: global memory is only touched below
  griddep_launch_dependents();
  const int n_items = p.batch * p.heads;

  if (warp == kProducerWarp && elect_one_sync()) {
    // ------------------------------------------------------------------ producer
    // K and Q of the NEXT item are loaded as soon as the last S of the current item has retired (kq_empty), i.e. while its last
    // query tile is still in the softmax / P V phases; V follows once the last P V has retired (item_empty).
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.heads, h = item % p.heads;
      wait_bar(kq_empty, ph ^ 1u);
      mbar_expect_tx(kq_full, (uint32_t)(NK * 128 + p.n_qt * BQ * 128));
      tma_load_3d(sK, &tmK, kq_full, h * HD, 0, b);
      tma_load_3d(sK + H * 128, &tmK, kq_full, h * HD, H, b);
      for (int qt = 0; qt < p.n_qt; ++qt) tma_load_3d(sQ + qt * BQ * 128, &tmQ, kq_full, h * HD, qt * BQ, b);
      wait_bar(item_empty, ph ^ 1u);
      mbar_expect_tx(v_full, (uint32_t)(NK * 128));
      for (int kb = 0; kb < NCH; ++kb) tma_load_3d(sV + kb * 64 * 128, &tmV, v_full, h * HD, kb * 64, b);
      ph ^= 1u;
    }
  } else if (warp == kMmaWarp && elect_one_sync()) {
    // ------------------------------------------------------------------ MMA issuer
    const int fmt = BF16 ? 1 : 0;
    const uint32_t idesc_s = umma_idesc_f16(BQ, H, fmt);
    const uint32_t idesc_o = idesc_pv(fmt);
    // S_h of query tile qt: 4 k-steps of 16 over the head dimension
    // (descriptor low words — address field + LBO — are precomputed; the high word is a constant: see gemm_tcgen05.cu)
    constexpr uint32_t kHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    auto join = [](uint32_t lo, uint32_t hi) __attribute__((always_inline)) {
      uint64_t d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
      return d;
    };
    const uint32_t q_lo0 = ((smem_u32(sQ) >> 4) & 0x3FFFu) | (1u << 16), k_lo0 = ((smem_u32(sK) >> 4) & 0x3FFFu) | (1u << 16);
    auto issue_s = [&](int qt, auto hc) __attribute__((always_inline)) {
      constexpr int h = decltype(hc)::value;
      const uint32_t a_lo = q_lo0 + (uint32_t)qt * (uint32_t)((BQ * 128) >> 4);
      const uint32_t b_lo = k_lo0 + (uint32_t)((h * H * 128) >> 4);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        tc_mma_f16(tmem_base + (uint32_t)(h * H), join(a_lo + (uint32_t)(2 * k), kHi), join(b_lo + (uint32_t)(2 * k), kHi), idesc_s, (uint32_t)(k != 0));
      tc_commit(s_full + h);
    };
    // O (+)= P_h V_h: H / 16 k-steps of 16 keys; key kk lives in block kk / 64 of P (K-major, 32 B per step inside the 128-byte
    // row) and of V (MN-major: 16 rows per step)
    // (fully unrolled with a compile-time half: every descriptor is base + constant.  The issuing thread shares its scheduler with
    // four busy softmax warps, so each dependent integer instruction in front of an MMA costs several issue rounds: with the
    // descriptors recomputed per step the 10 MMAs of a half took ~1900 clocks to ISSUE — 190 per MMA against ~40 of tensor time.)
    const uint32_t p_lo0 = ((smem_u32(sP) >> 4) & 0x3FFFu) | (1u << 16);
    // V block kb, key step s: address field (sV + kb * 8 KB + s * 2 KB) >> 4, LBO field (sOnes - (sV + kb * 8 KB)) >> 4: both are the
    // block-0 value plus / minus compile-time constants, so the low word is ONE add per MMA
    const uint32_t v_lo0 = ((smem_u32(sV) >> 4) & 0x3FFFu) | ((((smem_u32(sOnes) - smem_u32(sV)) >> 4) & 0x3FFFu) << 16);
    auto issue_pv = [&](auto hc) __attribute__((always_inline)) {
      constexpr int h = decltype(hc)::value;
#pragma unroll
      for (int j = 0; j < H / 16; ++j) {
        const int kk = h * H + j * 16;   // compile-time after unrolling
        const int blk = kk >> 6, step = (kk & 63) >> 4;
        const uint32_t a_lo = p_lo0 + (uint32_t)(((blk * (BQ * 128)) + ((kk & 63) * 2)) >> 4);
        const uint32_t b_lo = v_lo0 + (uint32_t)(((blk * (64 * 128)) + step * (16 * 128)) >> 4) - ((uint32_t)((blk * (64 * 128)) >> 4) << 16);
        tc_mma_f16(tmem_base + O_COL, join(a_lo, kHi), join(b_lo, kHi), idesc_o, (uint32_t)((h | j) != 0));   // columns 64..79: row sums of the same P
      }
      tc_commit(pv_done + h);
    };
    uint32_t ph_item = 0, pt = 0;  // parity of the item / of the query tile (every per-tile barrier completes once per tile)
    int trace_tile = 0;
    (void)trace_tile;
    wait_bar(kq_full, ph_item);
    tc_fence_after();
    issue_s(0, std::integral_constant<int, 0>{});
    issue_s(0, std::integral_constant<int, 1>{});
    if (p.n_qt == 1) tc_commit(kq_empty);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const bool has_next = item + (int)gridDim.x < n_items;
      for (int qt = 0; qt < p.n_qt; ++qt) {
        const bool same_item = qt + 1 < p.n_qt;
        const bool next = same_item || has_next;
        const int nqt = same_item ? qt + 1 : 0;
        // ---- half 0: P V, then S of the next tile's half 0 into the accumulator the softmax warps released long ago
        wait_bar(p_full + 0, pt);
        if (qt == 0) wait_bar(v_full, ph_item);
        wait_bar(o_empty, pt ^ 1u);   // O of the previous tile has been read out of TMEM
        tc_fence_after();
        ATTN_TRACE(true, trace_tile, 15);
        issue_pv(std::integral_constant<int, 0>{});
        ATTN_TRACE(true, trace_tile, 10);
        if (next) {
          wait_bar(s_free + 0, pt);
          if (!same_item) wait_bar(kq_full, ph_item ^ 1u);
          tc_fence_after();
          issue_s(nqt, std::integral_constant<int, 0>{});
        }
        ATTN_TRACE(true, trace_tile, 11);
        // ---- half 1
        wait_bar(p_full + 1, pt);
        tc_fence_after();
        issue_pv(std::integral_constant<int, 1>{});
        ATTN_TRACE(true, trace_tile, 12);
        if (!same_item) tc_commit(item_empty);
        if (next) {
          wait_bar(s_free + 1, pt);
          tc_fence_after();
          issue_s(nqt, std::integral_constant<int, 1>{});
          if (same_item ? (qt + 2 == p.n_qt) : (p.n_qt == 1)) tc_commit(kq_empty);  // the item's last S: K / Q are free once it retires
        }
        ATTN_TRACE(true, trace_tile, 13);
        ++trace_tile;
        pt ^= 1u;
      }
      ph_item ^= 1u;
    }
  } else if (warp < kSoftmaxWarps) {
    // ------------------------------------------------------------------ softmax + output
    const int q = warp & 3, part = warp >> 2;  // TMEM lane quarter; which quarter of a half's key columns
    const int r = q * 32 + lane;                                // row of the query tile owned by this lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float c2 = p.scale_log2;
    uint8_t* const p_row = sP + r * 128;   // this lane's row inside a P block
    const int rx = r & 7;                   // its 128B-swizzle phase
    uint32_t pt = 0;
    int trace_tile = 0;
    (void)trace_tile;
    const bool tr = (warp == 0 || warp == 1) && lane == 0;   // (q = 0, part = 0) and (q = 1, part = 0)
    const int trace_off = warp == 1 ? 32 : 0;
    (void)trace_off;
    (void)tr;
    // the tile whose O is still to be drained: always the previous tile of this CTA, possibly of the previous item.  Its (item,
    // tile, 1 / row sum) live in SHARED memory, not in registers: the softmax loop is at the register limit, and what the
    // compiler spills goes to the stack, whose L1 is almost entirely carved out for shared memory here (an L2 round trip in
    // front of a barrier wait); a 16-byte LDS per tile is cheap and predictable.
    uint2* my_cold = cold + threadIdx.x;
    bool have_pend = false;
    // O / rowsum of the previous query tile -> global (its last P V was issued when the previous tile's P_1 was complete)
    auto drain_o = [&]() __attribute__((always_inline)) {
      wait_bar(pv_done + 1, pt ^ 1u);
      tc_fence_after();
      ATTN_TRACE(tr, trace_tile + trace_off, 4);
      const uint2 pc = *my_cold;   // output pointer of the quarter's first row (48 bits), valid rows (bits 56..61), live (bit 63)
      uint16_t* const qbase = reinterpret_cast<uint16_t*>(((unsigned long long)(pc.y & 0x00FFFFFFu) << 32) | pc.x);
      const bool pend_live = (pc.y >> 31) != 0;
      const int pend_rows = (int)((pc.y >> 24) & 63u);
      uint32_t ov[16], lv[8];
      if (pend_live) {
        tmem_ld_32x16(lane_addr + (uint32_t)(O_COL + part * 16), ov);
        tmem_ld_32x8(lane_addr + (uint32_t)L_COL, lv);
        tmem_ld_wait();
      }
      ATTN_TRACE(tr, trace_tile + trace_off, 14);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (pend_live) {
        // Row-per-lane stores would touch 32 different 32-byte sectors per warp instruction (measured: the drain of a full tile
        // took ~1300 clocks, ~400 when only two quarters were live — LSU wavefronts).  The quarter's 32 x 128 B go through a
        // swizzled staging tile instead: the last P block, which the finished P V no longer reads and which this quarter's
        // warps only rewrite (exp pass of half 1) behind their next named barrier; 8 lanes then write one full 128-byte row.
        const float sum = __uint_as_float(lv[0]);
        const float pend_inv = sum > 0.f ? 1.f / sum : 0.f;   // no visible key: zeros, like the masked softmax of the reference's padding rows
        uint8_t* const stage = sP + 4 * (BQ * 128);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint4 u;
          uint32_t* uu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            uu[i] = H16<BF16 ? SCB_BF16 : SCB_F16>::pack(__uint_as_float(ov[t * 8 + 2 * i]) * pend_inv, __uint_as_float(ov[t * 8 + 2 * i + 1]) * pend_inv);
          *reinterpret_cast<uint4*>(stage + r * 128 + (((2 * part + t) ^ rx) << 4)) = u;
        }
        named_bar_sync(1 + q, 128);
        const int tq = part * 32 + lane, unit = tq & 7;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int row = k * 16 + (tq >> 3);
          if (row < pend_rows)
            *reinterpret_cast<uint4*>(qbase + (long long)row * p.o_ld + unit * 8) =
                *reinterpret_cast<const uint4*>(stage + (q * 32 + row) * 128 + ((unit ^ (row & 7)) << 4));
        }
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int kvl = p.kv_len ? min(p.kv_len[item / p.heads], p.Tk) : p.Tk;
      for (int qt = 0; qt < p.n_qt; ++qt) {
        const int row_g = qt * BQ + r;      // query index
        const bool live = qt * BQ + q * 32 < p.Tq;  // warp-uniform (and uniform over the 4 warps of this lane quarter)
        const int key_hi = p.causal ? min(kvl, row_g + 1) : kvl;  // keys [0, key_hi) are visible to this row
        int k_all = key_hi, k_any = key_hi;   // warp-uniform: keys below k_all are visible to every row of the warp, keys from k_any on to none
        if (p.causal) {
          k_all = __reduce_min_sync(0xffffffffu, key_hi);
          k_any = __reduce_max_sync(0xffffffffu, key_hi);
        }
        float m_run = -INFINITY;   // running row maximum (raw score units)
        ATTN_TRACE(tr, trace_tile + trace_off, 7);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          wait_bar(s_full + hf, pt);
          tc_fence_after();
          ATTN_TRACE(tr, trace_tile + trace_off, hf * 5 + 0);
          ATTN_TRACE_W(trace_tile, warp, hf * 2);
          const int col0 = hf * H + part * CW;   // first key column of this warp's slice (= TMEM column of S)
          if (live) {
            // ---- pass 1: row max of this half straight from TMEM (nothing is kept: holding the slice in registers from here to
            //      the exp pass pushes the loop over the register limit, and the spills / rematerialisation cost more than the
            //      second TMEM read), masking only on chunks that straddle key_hi; exchanged between the 4 warps of the quarter
            // (requesting the whole slice — 5 chunks, 40 registers — before one wait spills: 231 -> 400 us.  With the shared-memory
            //  carve-out at its maximum a spilled byte costs an L2 round trip.)
            float mx = -INFINITY;
#pragma unroll
            for (int c0 = 0; c0 < NCK; c0 += 2) {
              if (col0 + c0 * 8 < k_any) {   // (uniform) at least one row sees a key of this pair of chunks
                uint32_t v[2][8];
                tmem_ld_32x8(lane_addr + (uint32_t)(col0 + c0 * 8), v[0]);
                if (c0 + 1 < NCK) tmem_ld_32x8(lane_addr + (uint32_t)(col0 + c0 * 8 + 8), v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  if (c0 + cc < NCK) {
                    const int key0 = col0 + (c0 + cc) * 8;
                    if (key0 + 8 <= k_all) {
                      const float a = fmaxf(fmaxf(__uint_as_float(v[cc][0]), __uint_as_float(v[cc][1])), __uint_as_float(v[cc][2]));
                      const float b = fmaxf(fmaxf(__uint_as_float(v[cc][3]), __uint_as_float(v[cc][4])), __uint_as_float(v[cc][5]));
                      mx = fmaxf(fmaxf(mx, a), b);
                      mx = fmaxf(fmaxf(mx, __uint_as_float(v[cc][6])), __uint_as_float(v[cc][7]));   // 4 three-input FMNMX3 per 8 scores
                    } else {
#pragma unroll
                      for (int i = 0; i < 8; ++i)
                        if (key0 + i < key_hi) mx = fmaxf(mx, __uint_as_float(v[cc][i]));
                    }
                  }
                }
              }
            }
            float* rm = red_max + hf * 4 * BQ;
            rm[part * BQ + r] = mx;
            named_bar_sync(1 + q, 128);
            mx = fmaxf(fmaxf(rm[r], rm[BQ + r]), fmaxf(rm[2 * BQ + r], rm[3 * BQ + r]));
            ATTN_TRACE(tr, trace_tile + trace_off, hf * 5 + 1);
            if (hf == 0) {
              m_run = mx;
              // ---- the previous tile's O (the P V MMAs that produced it also read the P rows the stores below overwrite)
              if (have_pend) drain_o();
              ATTN_TRACE(tr, trace_tile + trace_off, 2);
            } else {
              // ---- half 1 keeps the running maximum unless its own is more than 2^8 above it; otherwise (rare) O and the partial
              //      row sum are rescaled in place.  The 4 warps of a quarter see the same rows, so they all take the same branch.
              const bool need = mx * c2 > m_run * c2 + kRescaleLog2;
              if (__any_sync(0xffffffffu, need)) {
                wait_bar(pv_done + 0, pt);      // P_0 V_0 of THIS tile has retired: O holds it
                tc_fence_after();
                const float alpha = need ? ex2_approx((m_run - mx) * c2) : 1.f;   // m_run = -inf (no visible key in half 0): 0
                uint32_t ov[16];
                tmem_ld_32x16(lane_addr + (uint32_t)(O_COL + part * 16), ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                tmem_st_32x16(lane_addr + (uint32_t)(O_COL + part * 16), ov);
                if (part == 0) {   // the row sums of half 0 live in TMEM too
                  tmem_ld_32x16(lane_addr + (uint32_t)L_COL, ov);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                  tmem_st_32x16(lane_addr + (uint32_t)L_COL, ov);
                }
                tmem_st_wait();
                tc_fence_before();
                if (need) m_run = mx;
              }
            }
            // ---- pass 2: p = 2^((s - m) * scale), row sum, P -> shared memory (K-major, 128B-swizzled A operand of P V).  A row
            //      of P is 5 blocks of 8 16-byte units; this warp's chunk c is unit u0 + c of the row.
            const float m_off = (m_run == -INFINITY) ? 0.f : -m_run * c2;
            const uint64_t c22 = pk2(c2, c2), mo2 = pk2(m_off, m_off);
            const int u0 = col0 >> 3;
#pragma unroll
            for (int c0 = 0; c0 < NCK; c0 += 2) {
              uint32_t v[2][8];
              tmem_ld_32x8(lane_addr + (uint32_t)(col0 + c0 * 8), v[0]);
              if (c0 + 1 < NCK) tmem_ld_32x8(lane_addr + (uint32_t)(col0 + c0 * 8 + 8), v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                if (c0 + cc < NCK) {
                  const int key0 = col0 + (c0 + cc) * 8;
                  uint32_t pk[4];
                  if (key0 + 8 <= k_all) {   // (uniform) every row sees all 8 keys: 4 packed FMAs, 8 exponentials, 4 packs
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      float a, b;
                      upk2(fma2(pk2(__uint_as_float(v[cc][2 * i]), __uint_as_float(v[cc][2 * i + 1])), c22, mo2), a, b);
                      pk[i] = H16<BF16 ? SCB_BF16 : SCB_F16>::pack(ex2_approx(a), ex2_approx(b));
                    }
                  } else if (key0 >= k_any) {
                    pk[0] = pk[1] = pk[2] = pk[3] = 0u;
                  } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const float a = key0 + 2 * i < key_hi ? ex2_approx(fmaf(__uint_as_float(v[cc][2 * i]), c2, m_off)) : 0.f;
                      const float b = key0 + 2 * i + 1 < key_hi ? ex2_approx(fmaf(__uint_as_float(v[cc][2 * i + 1]), c2, m_off)) : 0.f;
                      pk[i] = H16<BF16 ? SCB_BF16 : SCB_F16>::pack(a, b);
                    }
                  }
                  const int u = u0 + c0 + cc;
                  *reinterpret_cast<uint4*>(p_row + ((u >> 3) << 14) + (((u & 7) ^ rx) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
              }
            }
            tc_fence_before();
            fence_proxy_async();  // generic-proxy writes of P become visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(s_free + hf);   // S_hf has been read for the last time: the MMA thread may overwrite the accumulator
              mbar_arrive(p_full + hf);
            }
            ATTN_TRACE(tr, trace_tile + trace_off, hf * 5 + 3);
            ATTN_TRACE_W(trace_tile, warp, hf * 2 + 1);
          } else {
            // rows beyond Tq: nothing to compute (the MMA rows they would feed are never read); keep the barrier protocol going
            if (hf == 0 && have_pend) drain_o();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(s_free + hf);
              mbar_arrive(p_full + hf);
            }
          }
        }
        ATTN_TRACE(tr, trace_tile + trace_off, 9);
        ++trace_tile;
        {
          const int row0 = qt * BQ + q * 32;   // first query of this quarter
          const unsigned long long dst = (unsigned long long)(p.o + (long long)(item / p.heads) * p.o_bs + (long long)row0 * p.o_ld + (item % p.heads) * HD);
          const uint32_t rows = (uint32_t)max(0, min(32, p.Tq - row0));
          *my_cold = make_uint2((uint32_t)dst, ((uint32_t)(dst >> 32) & 0x00FFFFFFu) | (rows << 24) | (live ? 0x80000000u : 0u));
        }
        have_pend = true;
        pt ^= 1u;
      }
    }
    if (have_pend) drain_o();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp) {
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
  const uint32_t bk[3] = {64, (uint32_t)(p.NK / 2), 1};   // K is loaded as the two key halves
  const uint32_t bv[3] = {64, 64, 1};
  int e = make_tmap(&tmQ, q, 2, 3, dq, sq, bq, 1);
  if (!e) e = make_tmap(&tmK, k, 2, 3, dk, sk, bk, 1);
  if (!e) e = make_tmap(&tmV, v, 2, 3, dk, sv, bv, 1);
  if (e) return e;
  const int items = batch * heads;
  const int grid = items < num_sms() ? items : num_sms();
  static bool configured[2][6] = {};   // per kernel instantiation: the opt-in to > 48 KB of dynamic shared memory
  const int nch = p.NK / 64;
  auto launch = [&](auto kernel) -> int {
    if (!configured[p.bf16][nch]) {
      SCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      configured[p.bf16][nch] = true;
    }
    SCB_CUDA(launch_pdl(kernel, dim3((unsigned)grid), kThreads, SMEM_BYTES, st, tmQ, tmK, tmV, p));
    return SCB_OK;
  };
  if (p.bf16)
    e = nch == 1 ? launch(attention_tc_kernel<true, 1>) : nch == 2 ? launch(attention_tc_kernel<true, 2>) : nch == 3 ? launch(attention_tc_kernel<true, 3>)
        : nch == 4 ? launch(attention_tc_kernel<true, 4>) : launch(attention_tc_kernel<true, 5>);
  else
    e = nch == 1 ? launch(attention_tc_kernel<false, 1>) : nch == 2 ? launch(attention_tc_kernel<false, 2>) : nch == 3 ? launch(attention_tc_kernel<false, 3>)
        : nch == 4 ? launch(attention_tc_kernel<false, 4>) : launch(attention_tc_kernel<false, 5>);
  if (e) return e;
  note_launch();
  SCB_LAUNCH_OK("attention_tc");
  return SCB_OK;
}

}  // namespace scb
