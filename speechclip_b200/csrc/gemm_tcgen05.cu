// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 (1 thread)  TMA producer : cp.async.bulk.tensor A/B tiles -> 128B-swizzled smem ring
//   warp 1 (1 thread)  MMA issuer   : tcgen05.mma 128 x N x 16, fp32 accumulators in TMEM (2 buffers)
//   warp 2             TMEM allocator
//   warps 4..          epilogue     : tcgen05.ld -> alpha/bias/activation/residual -> global stores (16 warps for BN >= 128)
//
// The two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
// See include/speechclip_b200.h (scb_gemm) for the operand model (plain / strided-conv / grouped tap walk).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ops.cuh"

namespace scb {

namespace {

constexpr int BM = 128;
constexpr int kSlabTaps = 4;  // slab mode: taps (k-blocks) per pipeline stage — amortises the single-thread barrier/issue latency
constexpr int BK = 64;  // 16-bit elements per k-block (one 128-byte swizzle row); fp32/tf32 operands use 32
template <int BN> struct EpiCfg {
  static constexpr int WARPS = BN >= 128 ? 16 : 8;          // 4 (or 2) warps per TMEM lane quarter
  static constexpr int COLS = BN / (WARPS / 4);              // accumulator columns per warp: 64 / 32 / 32
  static constexpr int THREADS = (4 + WARPS) * 32;
  static constexpr int PATCH_BYTES = WARPS * 32 * 16 * 4;   // transpose patches of the register-store epilogue (2 KB per warp)
  static constexpr int TMA_BYTES = WARPS * 32 * 128;        // TMA-store epilogue: 32 rows x 128 B (64 16-bit columns) per warp
};
// Epilogues whose output is 16-bit with no residual (QKV, fc1, the conv stack, CLIP c_fc), and the fp32 + residual epilogues
// of the CTA-pair kernel (out-proj, fc2), leave through TMA stores: see epilogue_tile_tma / epilogue_tile_tma_res.  They give
// up one pipeline stage for the 64 KB of staging.
// (fp32 output + residual only on CTA pairs.  A 16-bit residual is read row-per-lane: 137 -> 127 us on the HuBERT out-proj.  An
// fp32 residual read that way costs more L1 wavefronts than the transpose patch it replaces (156 -> 182 us), so it arrives by
// TMA instead, INTO the staging tile the result leaves from: see epilogue_tile_tma_res.)
__host__ __device__ constexpr bool tma_out(bool pair, int BN, int ACT, int RES, int ODT) {
  return BN == 256 && ACT >= 0 && ((RES == 3 && ODT == SCB_F16) || (pair && ODT == SCB_F32 && (RES == SCB_F16 || RES == SCB_F32)));
}

// stream-K workspace (scb_gemm_args.workspace): arrival counters, then one [2][128 x 256] fp32 slot per CTA pair
constexpr int kSkFlagBytes = 4096;
inline bool stream_k_applies(long long tiles, int pairs, int k_blocks) {
  const long long rem = tiles % pairs;
  // worth it when the last wave is at most 80 % full, and every pair's share of it is at least 2 k-blocks
  return rem > 0 && rem * 5 <= (long long)pairs * 4 && rem * k_blocks >= 2LL * pairs;
}

// The MMA-issuing thread runs a long chain of dependent scalar instructions between tcgen05.mma issues; measured on the attention
// kernel, every instruction in front of an MMA costs ~6-10 clocks, and on the 128 x 48 pos-conv tiles (96 clocks of tensor time
// per tap) rebuilding two 64-bit descriptors per tap made the issue thread the bottleneck (~425 clocks per tap).  The shared-
// memory descriptors therefore live as a 32-bit low word (address field + LBO) that is advanced by constants, and a constant
// high word (SBO, version, swizzle).  Shared addresses are < 256 KB, so the 14-bit address field never carries.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B

struct GemmParams {
  int batch, m_tiles_per_batch, n_tiles, groups, num_tiles;
  int m_per_batch, n, k_blocks;
  int kb_per_tap, tap_row_shift, a_col0, a_group_cols;
  int umma_n;
  int bk;  // elements per k-block: 64 (16-bit) or 32 (tf32)
  int slab;       // 1: one-tap-per-k-block walk with row shift 1 (positional conv): A is loaded ONCE per tile as a 256-row slab
  int slab_stages, slab_sub_bytes;  // slab mode: B ring of slab_stages stages, each slab_taps sub-tiles of slab_sub_bytes
  // slab mode, several M tiles per work item (pos-conv: T = 319 is 3 tiles): the slab holds slab_mt * 128 + taps - 1 rows, loaded as
  // slab_parts boxes of slab_box_rows rows; every tap's weight sub-tile then feeds slab_mt x 4 MMAs instead of 4 — the weights
  // of a group (786 KB at 48 x 8192) are streamed from L2 once per utterance instead of once per 128 frames
  int slab_mt, slab_taps, slab_parts, slab_box_rows, slab_bytes;
  int tile_m;  // rows of M covered by one tile index of the single-CTA kernel (BM, or BM * slab_mt)
  uint32_t tx_bytes;
  void* out;
  void* out2;
  const float* bias;
  const void* residual;
  int out_dtype, out2_dtype, residual_dtype, act, ab_fmt;  // ab_fmt: 0 f16, 1 bf16, 2 tf32 (UMMA format codes)
  float alpha;
  long long ldc, out_batch_stride;
  long long res_ld, res_batch_stride;
  int out_group_cols;
  // stream-K tail of the CTA-pair kernel (sk_rem == 0: plain tile loop): every pair first takes sk_full whole tiles, then the
  // sk_rem tiles of the last, partial wave are cut along K into one contiguous range of k-blocks per pair
  int sk_full, sk_rem;
  float* sk_part;  // [pairs][2 CTAs][128 x 256] fp32 partial accumulators of the ranges that do not start a tile
  int* sk_flags;   // [sk_rem][2 CTAs] arrival counters (zero between launches: the owner resets them)
};

struct TileCoord {
  int g, b, m0, n0;
};

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int rest = tile / p.n_tiles;
  const int mt = rest % p.m_tiles_per_batch;
  rest /= p.m_tiles_per_batch;
  t.b = rest % p.batch;
  t.g = rest / p.batch;
  t.m0 = mt * p.tile_m;
  t.n0 = nt * BN;
  return t;
}

// 4 consecutive output elements of one row (8 B for 16-bit, 16 B for fp32).
__device__ __forceinline__ void store4(void* base, int dtype, long long off, const float (&x)[4]) {
  if (dtype == SCB_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = make_float4(x[0], x[1], x[2], x[3]);
  } else {
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + off) = make_uint2(pack16(dtype, x[0], x[1]), pack16(dtype, x[2], x[3]));
  }
}

// One output tile of the epilogue, executed by every epilogue warp of a CTA.  tfull_bar: the CTA-local "accumulator ready"
// barrier; tempty_addr: shared::cluster address of the "accumulator drained" barrier of the CTA that issues the MMAs (the CTA
// itself, or the pair leader in cta_group::2 mode).
//
// tcgen05.ld (32x32b) hands each lane one ROW of the accumulator.  Writing global memory in that layout costs one LSU
// wavefront per lane per instruction (32 different cache lines): measured 20 us per 128x256 fp32+residual tile against
// 3.5 us of MMA.  So every 32x16 chunk is transposed through a private, XOR-swizzled (bank-conflict-free) 2 KB shared
// patch: afterwards 4 lanes cover one 64-byte row segment and a warp instruction touches 8 rows instead of 32.  The
// residual loads of a chunk are issued before the TMEM load (out may alias residual, so the compiler cannot hoist them).
// ACT / RES / ODT >= 0 pin the activation, the residual dtype (3 = no residual) and the output dtype at compile time (and imply
// out2 == NULL) for the shapes that dominate the step; -1 keeps the runtime switch.
// tmem_col0: first TMEM column of the tile's accumulator; release: this is the last tile read from the accumulator buffer (the
// buffer holds several tiles in the multi-tile slab mode), so the warp hands it back after its last load.
template <int BN, int ACT, int RES, int ODT>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const TileCoord& t, uint32_t tmem_base, int tmem_col0, uint32_t acc_phase,
                                              uint64_t* tfull_bar, uint32_t tempty_addr, float* stage, int warp, int lane, bool release = true) {
  constexpr int COLS = EpiCfg<BN>::COLS;
  const int q = warp & 3;            // TMEM lane quarter this warp may touch
  const int part = (warp - 4) >> 2;  // which slice of the tile's columns
  float4* stg = reinterpret_cast<float4*>(stage + (warp - 4) * (32 * 16));
  const int prow = lane >> 2, u = lane & 3;  // after the transpose: row (within a pass of 8) and 4-column unit of this lane
  const int act = ACT >= 0 ? ACT : p.act;
  const bool has_res = RES >= 0 ? (RES != 3) : (p.residual != nullptr);
  const int res_dt = (RES >= 0 && RES != 3) ? (RES & 15) : p.residual_dtype;   // (RES = 16 + dtype: the same, kept off the TMA epilogue)
  const int out_dt = ODT >= 0 ? ODT : p.out_dtype;
  void* const out2 = ODT >= 0 ? nullptr : p.out2;
  {
      const int m_base = t.m0 + q * 32;
      const int col_base = t.n0 + part * COLS;
      const long long gcol = (long long)t.g * p.out_group_cols;
      const long long out_base = (long long)t.b * p.out_batch_stride + (long long)m_base * p.ldc + gcol;
      const long long res_base = (long long)t.b * p.res_batch_stride + (long long)m_base * p.res_ld + gcol;
      int nchunks = 0;
      if (col_base < p.n && m_base < p.m_per_batch) nchunks = min(COLS / 16, (p.n - col_base + 15) / 16);
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      if (nchunks == 0 && release) {
        tc_fence_before();
        if (lane == 0) mbar_arrive_cluster(tempty_addr);
      }
      for (int ch = 0; ch < nchunks; ++ch) {
        const int col = col_base + ch * 16 + u * 4;  // first of this lane's 4 columns
        const bool col_ok = col < p.n;               // n % 8 == 0: a 4-column unit is entirely inside or outside
        // ---- residual + bias: everything this chunk needs from global memory, in flight at once
        uint4 rr[4];
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
          if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol + col));
          if (has_res) {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const int r = ps * 8 + prow;
              if (m_base + r < p.m_per_batch) {
                const long long off = res_base + (long long)r * p.res_ld + col;
                if (res_dt == SCB_F32) {
                  rr[ps] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.residual) + off));
                } else {
                  const uint2 h2 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.residual) + off));
                  rr[ps] = make_uint4(h2.x, h2.y, 0u, 0u);
                }
              }
            }
          }
        }
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tmem_col0 + part * COLS + ch * 16), v);
        tmem_ld_wait();
        if (ch == nchunks - 1 && release) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
          tc_fence_before();
          if (lane == 0) mbar_arrive_cluster(tempty_addr);
        }
        __syncwarp();  // the previous chunk's reads of the patch are done
#pragma unroll
        for (int k = 0; k < 4; ++k)  // row = lane; 16-byte unit k lands at k ^ ((lane >> 1) & 3): conflict-free both ways
          stg[lane * 4 + (k ^ ((lane >> 1) & 3))] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]),
                                                               __uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3]));
        __syncwarp();
        if (col_ok) {
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int r = ps * 8 + prow;
            if (m_base + r < p.m_per_batch) {
              const float4 a = stg[r * 4 + (u ^ ((r >> 1) & 3))];
              float x[4] = {fmaf(p.alpha, a.x, bias4.x), fmaf(p.alpha, a.y, bias4.y), fmaf(p.alpha, a.z, bias4.z),
                            fmaf(p.alpha, a.w, bias4.w)};
              if (act == SCB_ACT_GELU_ERF) {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = gelu_fast(x[i]);
              } else if (act == SCB_ACT_QUICK_GELU) {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = quick_gelu(x[i]);
              }
              if (has_res) {
                if (res_dt == SCB_F32) {
                  x[0] += __uint_as_float(rr[ps].x); x[1] += __uint_as_float(rr[ps].y);
                  x[2] += __uint_as_float(rr[ps].z); x[3] += __uint_as_float(rr[ps].w);
                } else {
                  const float2 f0 = unpack16(res_dt, rr[ps].x), f1 = unpack16(res_dt, rr[ps].y);
                  x[0] += f0.x; x[1] += f0.y; x[2] += f1.x; x[3] += f1.y;
                }
              }
              const long long off = out_base + (long long)r * p.ldc + col;
              store4(p.out, out_dt, off, x);
              if (out2) store4(out2, p.out2_dtype, off, x);
            }
          }
        }
      }
  }
}


// TMA-store epilogue (16-bit output, bias + activation, no residual).  tcgen05.ld hands lane r ROW r of the accumulator; the
// lane converts its 64 columns to 16-bit and writes them as eight 16-byte units into row r of a [32 rows][128 B] staging tile
// in the 128B-swizzle pattern (unit c of row r at slot c ^ (r & 7): the 8 lanes of a quarter-warp hit 8 different slots, so
// the stores are bank-conflict-free), and one elected lane hands the tile to cp.async.bulk.tensor: no transpose pass, no
// per-lane global addressing or predicates (the tensor map clips the M / N edges), about half the instructions of the
// register-store path — which is what bounds the K = 768 shapes (fc1: 22 instructions per output measured, 15 of them GELU).
template <int ACT>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tmO, const TileCoord& t, uint32_t tmem_base,
                                                  int acc, uint32_t acc_phase, uint64_t* tfull_bar, uint32_t tempty_addr,
                                                  uint8_t* stage_all, int warp, int lane) {
  constexpr int BN = 256, COLS = EpiCfg<BN>::COLS;
  static_assert(COLS == 64, "one 128-byte staging row per accumulator row");
  const int q = warp & 3;
  const int part = (warp - 4) >> 2;
  const int m_base = t.m0 + q * 32;
  const int col_base = t.n0 + part * COLS;
  const uint32_t stg = smem_u32(stage_all + (warp - 4) * (32 * 128)) + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  const bool live = col_base < p.n && m_base < p.m_per_batch;
  const bool has_bias = p.bias != nullptr;
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  if (lane == 0) bulk_wait_read0();  // the previous tile's store has finished reading this warp's staging tile
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    if (live) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * COLS + h * 32), v);
    tmem_ld_wait();
    if (h == 1) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // 8 columns -> one 16-byte unit
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (has_bias) {  // columns past n (clipped by the store) read the last in-range unit instead: n % 8 == 0
          const float* bp = p.bias + min(col_base + h * 32 + k * 8, p.n - 8);
          b0 = __ldg(reinterpret_cast<const float4*>(bp));
          b1 = __ldg(reinterpret_cast<const float4*>(bp + 4));
        }
        const uint64_t al2 = pk2(p.alpha, p.alpha);
        uint64_t x[4] = {fma2(al2, pk2(__uint_as_float(v[8 * k + 0]), __uint_as_float(v[8 * k + 1])), pk2(b0.x, b0.y)),
                         fma2(al2, pk2(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])), pk2(b0.z, b0.w)),
                         fma2(al2, pk2(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])), pk2(b1.x, b1.y)),
                         fma2(al2, pk2(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])), pk2(b1.z, b1.w))};
        uint32_t h16[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ACT == SCB_ACT_GELU_ERF) x[i] = gelu_h16_x2(x[i]);
          else if (ACT == SCB_ACT_QUICK_GELU) x[i] = quick_gelu_x2(x[i]);
          float lo, hi;
          upk2(x[i], lo, hi);
          h16[i] = H16<SCB_F16>::pack(lo, hi);
        }
        const uint32_t unit = (uint32_t)(h * 4 + k);
        st_shared_v4(stg + ((unit ^ sw) << 4), h16[0], h16[1], h16[2], h16[3]);
      }
    }
  }
  if (live) {
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmO, stg, col_base, m_base, t.b);
      bulk_commit();
    }
  }
}

// TMA-store epilogue for fp32 output = alpha * acc + bias + residual (out-proj, fc2).  The register-store path above spends a
// 128 x 256 tile's epilogue in serial shared-memory round trips (transpose patch: STS -> LDS, ~1 instruction per cycle per SM
// measured, 10 us per tile against 4 us of MMAs at K = 768).  Here lane r keeps ROW r: it reads its own residual row segment
// (32 columns: 64 contiguous bytes of fp16 or 128 of fp32, issued before the accumulator wait), adds, writes the 32 fp32
// columns as eight swizzled 16-byte units of a [32 rows][128 B] staging tile and one lane issues the TMA store; a warp's 64
// columns are two such rounds through the same 4 KB tile (the second waits for the first store's shared-memory reads).
// fp32 residual (pre-LN towers: CLIP ViT, HuBERT-large): the warp's 32 x 32 residual box is TMA-loaded into the staging tile
// itself (same box and swizzle as the store), each lane adds its accumulator row in place (8 conflict-free 16-byte units) and
// the tile leaves by TMA again — no per-lane global addressing on either side.  rbar / rphase: the warp's own mbarrier for
// the residual loads and its running parity.
template <int RES>
__device__ __forceinline__ void epilogue_tile_tma_res(const GemmParams& p, const CUtensorMap* tmO, const CUtensorMap* tmR, const TileCoord& t,
                                                      uint32_t tmem_base, int acc, uint32_t acc_phase, uint64_t* tfull_bar, uint32_t tempty_addr,
                                                      uint8_t* stage_all, uint64_t* rbar, uint32_t& rphase, int warp, int lane) {
  constexpr int BN = 256, COLS = EpiCfg<BN>::COLS;
  static_assert(COLS == 64, "two 32-column rounds per warp");
  if constexpr (RES == SCB_F32) {
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    const int m_base = t.m0 + q * 32;
    const int col_base = t.n0 + part * COLS;
    uint8_t* const tile = stage_all + (warp - 4) * (32 * 128);
    const uint32_t stg = smem_u32(tile) + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const bool live = col_base < p.n && m_base < p.m_per_batch;
    const bool has_bias = p.bias != nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (live && lane == 0) {
        bulk_wait_read0();  // the previous store (last tile / first round) has finished reading the staging tile
        mbar_expect_tx(rbar, 32u * 128u);
        tma_load_3d(tile, tmR, rbar, col_base + h * 32, m_base, t.b);   // rows / columns past the edges arrive as zeros
      }
      if (h == 0) {
        mbar_wait(tfull_bar, acc_phase);
        tc_fence_after();
      }
      uint32_t v[32];
      if (live) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * COLS + h * 32), v);
      tmem_ld_wait();
      if (h == 1) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
        tc_fence_before();
        if (lane == 0) mbar_arrive_cluster(tempty_addr);
      }
      if (live) {
        mbar_wait(rbar, rphase);
        rphase ^= 1u;
        const uint64_t al2 = pk2(p.alpha, p.alpha);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 4 columns -> one 16-byte unit of fp32
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + min(col_base + h * 32 + k * 4, p.n - 4)));
          const uint32_t addr = stg + (((uint32_t)k ^ sw) << 4);
          uint32_t r0, r1, r2, r3;
          ld_shared_v4(addr, r0, r1, r2, r3);
          const uint64_t x0 = add2(fma2(al2, pk2(__uint_as_float(v[4 * k + 0]), __uint_as_float(v[4 * k + 1])), pk2(b0.x, b0.y)),
                                   pk2(__uint_as_float(r0), __uint_as_float(r1)));
          const uint64_t x1 = add2(fma2(al2, pk2(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])), pk2(b0.z, b0.w)),
                                   pk2(__uint_as_float(r2), __uint_as_float(r3)));
          float o0, o1, o2, o3;
          upk2(x0, o0, o1);
          upk2(x1, o2, o3);
          st_shared_v4(addr, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(tmO, stg, col_base + h * 32, m_base, t.b);
          bulk_commit();
        }
      }
    }
    return;
  }
  constexpr int RW = RES == SCB_F32 ? 8 : 4;  // 16-byte units of residual per lane per round
  const int q = warp & 3;
  const int part = (warp - 4) >> 2;
  const int m_base = t.m0 + q * 32;
  const int col_base = t.n0 + part * COLS;
  const int row = m_base + lane;
  const uint32_t stg = smem_u32(stage_all + (warp - 4) * (32 * 128)) + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  const bool live = col_base < p.n && m_base < p.m_per_batch;
  const bool row_ok = row < p.m_per_batch;
  const bool has_bias = p.bias != nullptr;
  const long long res_row = (long long)t.b * p.res_batch_stride + (long long)row * p.res_ld;
  uint4 rr[RW];
  auto load_res = [&](int h) {  // n % 8 == 0: an 8-column group is entirely inside or outside (outside is clipped by the store)
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const int col = col_base + h * 32 + i * (32 / RW);
      rr[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && col < p.n) {
        if (RES == SCB_F32) rr[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.residual) + res_row + col));
        else rr[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + res_row + col));
      }
    }
  };
  if (live) load_res(0);
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    if (live) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * COLS + h * 32), v);
    tmem_ld_wait();
    if (h == 1) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    if (lane == 0) bulk_wait_read0();  // the previous store (last tile / first round) has finished reading the staging tile
    __syncwarp();
    if (live) {
      const uint64_t al2 = pk2(p.alpha, p.alpha);
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // 4 columns -> one 16-byte unit of fp32
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + min(col_base + h * 32 + k * 4, p.n - 4)));
        float r0, r1, r2, r3;
        if (RES == SCB_F32) {
          r0 = __uint_as_float(rr[k].x); r1 = __uint_as_float(rr[k].y); r2 = __uint_as_float(rr[k].z); r3 = __uint_as_float(rr[k].w);
        } else {
          const uint4 u4 = rr[k >> 1];
          const float2 f0 = H16<RES == SCB_BF16 ? SCB_BF16 : SCB_F16>::unpack((k & 1) ? u4.z : u4.x);
          const float2 f1 = H16<RES == SCB_BF16 ? SCB_BF16 : SCB_F16>::unpack((k & 1) ? u4.w : u4.y);
          r0 = f0.x; r1 = f0.y; r2 = f1.x; r3 = f1.y;
        }
        const uint64_t x0 = add2(fma2(al2, pk2(__uint_as_float(v[4 * k + 0]), __uint_as_float(v[4 * k + 1])), pk2(b0.x, b0.y)), pk2(r0, r1));
        const uint64_t x1 = add2(fma2(al2, pk2(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])), pk2(b0.z, b0.w)), pk2(r2, r3));
        float o0, o1, o2, o3;
        upk2(x0, o0, o1);
        upk2(x1, o2, o3);
        st_shared_v4(stg + (((uint32_t)k ^ sw) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
      }
      if (h == 0) load_res(1);  // the second round's residual is in flight during the first round's store
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(tmO, stg - (uint32_t)lane * 128u, col_base + h * 32, m_base, t.b);
        bulk_commit();
      }
    }
  }
}

template <int BN, int STAGES, int ACT, int RES, int ODT>
__global__ void __launch_bounds__(EpiCfg<BN>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr bool TMAO = tma_out(false, BN, ACT, RES, ODT);
  constexpr int EPI_BYTES = TMAO ? EpiCfg<BN>::TMA_BYTES : EpiCfg<BN>::PATCH_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset arithmetic keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* epi = sB + STAGES * B_BYTES;  // 1024-aligned: transpose patches or TMA-store staging tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* sfull = tempty + 2;   // slab mode: A slab landed / A slab free (2 slabs carved out of the unused A ring)
  uint64_t* sempty = sfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 2);
  constexpr int SLAB_BYTES = 256 * BK * 2;  // 256 rows x 128 B
  static_assert(BN > 64 || STAGES * A_BYTES >= 2 * SLAB_BYTES, "the A ring must hold two slabs");
  float* stage = reinterpret_cast<float*>(epi);  // [epilogue warp][32 rows x 16 cols] transpose patches

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMAO) tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EpiCfg<BN>::WARPS);
      mbar_init(&sfull[s], 1);
      mbar_init(&sempty[s], 1);
    }
    mbar_fence_init();
  }
  const bool wide_tmem = BN == 64 && p.slab_mt > 1;   // 2 buffers x 256 columns
  if (warp == 2) {
    if (wide_tmem) tmem_alloc<512>(tmem_slot);
    else tmem_alloc<2 * BN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run while the
  // previous kernel of the stream drains; global memory is only touched below this point.
  griddep_wait();
  griddep_launch_dependents();  // a following row kernel (LayerNorm) may become resident beside this CTA and wait for its data

  if (warp == 0 && elect_one_sync()) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int sl = 0;
    uint32_t sl_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<BN>(p, tile);
      const int a_c0 = p.a_col0 + t.g * p.a_group_cols;
      if (p.slab) {  // rows m0 .. m0+255 of this group's 64 columns: every tap's A tile is a row-shifted window of it
        mbar_wait(&sempty[sl], sl_phase ^ 1u);
        mbar_expect_tx(&sfull[sl], (uint32_t)p.slab_bytes);
        for (int part = 0; part < p.slab_parts; ++part)
          tma_load_3d(sA + sl * p.slab_bytes + part * (p.slab_box_rows * 128), &tmA, &sfull[sl], a_c0, t.m0 + part * p.slab_box_rows, t.b);
        sl ^= 1;
        if (sl == 0) sl_phase ^= 1u;
      }
      if (p.slab) {
        // B ring for the slab walk: kSlabTaps consecutive taps per stage, carved out of [sA + 2 slabs, end of the B ring)
        uint8_t* sBs = sA + 2 * p.slab_bytes;
        for (int kb0 = 0; kb0 < p.k_blocks; kb0 += p.slab_taps) {
          const int ntap = min(p.slab_taps, p.k_blocks - kb0);
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], (uint32_t)(ntap * p.slab_sub_bytes));
          uint8_t* dst = sBs + stage * (p.slab_taps * p.slab_sub_bytes);
          int b_col = kb0 * p.bk;
          for (int tp = 0; tp < ntap; ++tp, dst += p.slab_sub_bytes, b_col += p.bk) tma_load_3d(dst, &tmB, &full[stage], b_col, t.n0, t.g);
          if (++stage == p.slab_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        continue;
      }
      // (coordinates advance incrementally: an integer division per k-block on this single thread costs more than the TMA issue)
      int a_col = a_c0, a_row = t.m0, b_col = 0, kin = 0;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_expect_tx(&full[stage], p.tx_bytes);
        tma_load_3d(sA + stage * A_BYTES, &tmA, &full[stage], a_col, a_row, t.b);
        tma_load_3d(sB + stage * B_BYTES, &tmB, &full[stage], b_col, t.n0, t.g);
        b_col += p.bk;
        a_col += p.bk;
        if (++kin == p.kb_per_tap) {   // next tap: back to the group's first column, tap_row_shift rows further
          kin = 0;
          a_col = a_c0;
          a_row += p.tap_row_shift;
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_f16(BM, p.umma_n, p.ab_fmt);
    const bool tf32 = p.ab_fmt == 2;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sl = 0;
    uint32_t sl_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * (wide_tmem ? 256 : BN));
      if (p.slab) {
        mbar_wait(&sfull[sl], sl_phase);
        tc_fence_after();
        // Tap kb reads rows kb .. kb+127 of the slab: the A descriptor's start address advances by kb rows (128 B each).
        // The 128B swizzle XORs the 16-byte chunk index with ABSOLUTE shared-address bits [7:9] (in TMA and in the MMA's
        // operand fetch alike), so a row-shifted window of a 1024-aligned slab needs no descriptor base offset (measured:
        // base offset (kb & 7) gives wrong results, 0 is bit-identical to reloading every tap).
        uint32_t a_lo = desc_lo(smem_u32(sA + sl * p.slab_bytes));   // + 8 (one 128-byte row) per tap; + 1024 per M tile
        const uint32_t b_lo0 = desc_lo(smem_u32(sA + 2 * p.slab_bytes));
        const uint32_t sub4 = (uint32_t)p.slab_sub_bytes >> 4;
        uint32_t accf = 0;   // 0 during the first tap (every tile's first MMA overwrites its accumulator)
        for (int kb0 = 0; kb0 < p.k_blocks; kb0 += p.slab_taps) {
          const int ntap = min(p.slab_taps, p.k_blocks - kb0);
          uint32_t b_lo = b_lo0 + (uint32_t)(stage * p.slab_taps) * sub4;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int tp = 0; tp < kSlabTaps; ++tp) {
            if (tp < ntap) {
#pragma unroll
              for (int mt = 0; mt < 3; ++mt) {
                if (mt < p.slab_mt) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    tc_mma_f16(d_tmem + (uint32_t)(mt * BN), desc_join(a_lo + (uint32_t)(mt * 1024 + 2 * k), kDescHi),
                               desc_join(b_lo + (uint32_t)(2 * k), kDescHi), idesc, k == 0 ? accf : 1u);
                }
              }
              accf = 1;
              a_lo += 8;
              b_lo += sub4;
            }
          }
          tc_commit(&empty[stage]);
          if (++stage == p.slab_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      {
        const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
        uint32_t accf = 0;
        for (int kb = 0; kb < (p.slab ? 0 : p.k_blocks); ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (uint32_t)(A_BYTES >> 4), b_lo = b_lo0 + (uint32_t)stage * (uint32_t)(B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (tf32) tc_mma_tf32(d_tmem, desc_join(a_lo + (uint32_t)(2 * k), kDescHi), desc_join(b_lo + (uint32_t)(2 * k), kDescHi), idesc, accf);
            else tc_mma_f16(d_tmem, desc_join(a_lo + (uint32_t)(2 * k), kDescHi), desc_join(b_lo + (uint32_t)(2 * k), kDescHi), idesc, accf);
            accf = 1;
          }
          tc_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.slab) {
        tc_commit(&sempty[sl]);  // every MMA that read this slab has retired
        sl ^= 1;
        if (sl == 0) sl_phase ^= 1u;
      }
      tc_commit(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (see epilogue_tile)
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t rphase_unused = 0;   // (the TMA-loaded fp32 residual exists on CTA pairs only)
    (void)rphase_unused;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<BN>(p, tile);
      if constexpr (TMAO && ODT == SCB_F32) epilogue_tile_tma_res<RES>(p, &tmO, &tmO, t, tmem_base, acc, acc_phase, &tfull[acc], smem_u32(&tempty[acc]), epi, nullptr, rphase_unused, warp, lane);
      else if constexpr (TMAO) epilogue_tile_tma<ACT>(p, &tmO, t, tmem_base, acc, acc_phase, &tfull[acc], smem_u32(&tempty[acc]), epi, warp, lane);
      else if (BN == 64 && p.slab_mt > 1) {   // several M tiles per accumulator buffer (256 columns per buffer, 64 per tile)
        for (int mt = 0; mt < p.slab_mt; ++mt) {
          TileCoord tt = t;
          tt.m0 = t.m0 + mt * BM;
          epilogue_tile<BN, ACT, RES, ODT>(p, tt, tmem_base, acc * 256 + mt * BN, acc_phase, &tfull[acc], smem_u32(&tempty[acc]), stage, warp, lane,
                                           mt == p.slab_mt - 1);
        }
      } else epilogue_tile<BN, ACT, RES, ODT>(p, t, tmem_base, acc * BN, acc_phase, &tfull[acc], smem_u32(&tempty[acc]), stage, warp, lane);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (TMAO && lane == 0) bulk_wait_all();  // the staging tiles must outlive the last store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (wide_tmem) tmem_dealloc<512>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ cta_group::2 variant
// A CTA pair (one cluster of 2 on a TPC) computes a 256 x 256 output tile: each CTA stages its own 128 A rows and HALF of the
// B tile (128 of the 256 N rows) per k-block, the leader issues tcgen05.mma.cta_group::2 (M = 256 across the pair, both CTAs'
// B halves are read by the pair's tensor cores), and each CTA's TMEM receives its 128 x 256 accumulator.  Per output tile the
// pair pulls (128 + 128) x 64 operand elements per CTA instead of (128 + 256): a third less L2 -> SM traffic per flop (the
// 1-CTA kernel saturates at ~1.1 PFLOP/s on the big shapes), and the 32 KB stages make the TMA ring 6 deep instead of 4.
//   full[s]   (leader)  : 1 arrival (leader's expect_tx) + 64 KB of transaction bytes from BOTH CTAs' TMA loads
//   empty[s]  (each CTA): tcgen05.commit multicast from the leader -> that CTA's producer may refill the stage
//   tfull[a]  (each CTA): tcgen05.commit multicast -> that CTA's epilogue may read its accumulator half
//   tempty[a] (leader)  : 2 x 16 epilogue warps (local + remote arrivals) -> the leader may overwrite the accumulators
constexpr int BN2 = 256;
__host__ __device__ constexpr int stages2(bool tma_epilogue) { return tma_epilogue ? 5 : 6; }

__device__ __forceinline__ TileCoord decode_tile2(const GemmParams& p, int tile, int rank) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int rest = tile / p.n_tiles;
  const int mt = rest % p.m_tiles_per_batch;  // 256-row super tiles
  rest /= p.m_tiles_per_batch;
  t.b = rest % p.batch;
  t.g = rest / p.batch;
  t.m0 = mt * 256 + rank * BM;
  t.n0 = nt * BN2;
  return t;
}

// ---- stream-K tail.  A GEMM whose tile count is not a multiple of the number of CTA pairs leaves most SMs idle in its last
// wave (HuBERT out-proj / fc2 at 32 utterances per GPU: 120 tiles on 74 pairs = 1.6 waves; CLIP ViT-B/32 fc2 at 256 images:
// 150 tiles = 2.03 waves).  Here the tiles of that last wave are cut along K instead: pair w gets the k-block units
// [w * UR / W, (w + 1) * UR / W) of the UR = sk_rem * k_blocks units, i.e. a piece of one tile and possibly the head of the
// next.  The pair whose range STARTS a tile owns it: the others dump their fp32 accumulators to a workspace slot and bump
// the tile's counter; the owner adds the partial sums into its TMEM accumulator (tcgen05.ld / add / tcgen05.st) and then runs
// the ordinary epilogue unchanged.  All pairs reach their ranges at the same time (after their whole tiles) and all are
// resident, so the owner's wait is short and cannot deadlock.
template <bool SK>
__device__ __forceinline__ bool sk_item(const GemmParams& p, int worker, int W, int i, int& tile, int& kb0, int& kb1) {
  kb0 = 0;
  kb1 = p.k_blocks;
  if (!SK) {
    tile = worker + i * W;
    return tile < p.num_tiles;
  }
  if (i < p.sk_full) {
    tile = worker + i * W;
    return true;
  }
  const int UR = p.sk_rem * p.k_blocks;
  const int u0 = (int)((long long)worker * UR / W), u1 = (int)((long long)(worker + 1) * UR / W);
  if (u1 <= u0) return false;
  const int r = u0 / p.k_blocks;
  if (i == p.sk_full) {
    tile = p.sk_full * W + r;
    kb0 = u0 - r * p.k_blocks;
    kb1 = min(p.k_blocks, kb0 + (u1 - u0));
    return true;
  }
  if (i == p.sk_full + 1 && u1 > (r + 1) * p.k_blocks) {
    tile = p.sk_full * W + r + 1;
    kb1 = u1 - (r + 1) * p.k_blocks;
    return true;
  }
  return false;
}
// pairs other than the owner that hold a piece of remainder tile r
__device__ __forceinline__ int sk_partners(const GemmParams& p, int worker, int W, int r) {
  const int UR = p.sk_rem * p.k_blocks;
  const long long u_last = (long long)(r + 1) * p.k_blocks - 1;
  return (int)(((u_last + 1) * W - 1) / UR) - worker;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EpiCfg<256>::WARPS * 32) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Slot layout: [epilogue warp][16-column chunk][16-byte unit][lane] float4 — every warp store / load is 512 contiguous bytes.
__device__ __forceinline__ void sk_dump(float* slot, int* flag, uint32_t tmem_base, int acc, uint32_t acc_phase, uint64_t* tfull_bar,
                                        uint32_t tempty_addr, int warp, int lane) {
  constexpr int BN = 256, COLS = EpiCfg<BN>::COLS;
  const int q = warp & 3, part = (warp - 4) >> 2;
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  float4* dst = reinterpret_cast<float4*>(slot) + (size_t)(warp - 4) * (COLS / 16) * 4 * 32 + lane;
#pragma unroll
  for (int c = 0; c < COLS / 16; ++c) {
    uint32_t v[16];
    tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * COLS + c * 16), v);
    tmem_ld_wait();
    if (c == COLS / 16 - 1) {  // accumulator fully read by this warp: hand the buffer back to the MMA warp
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      __stcg(dst + (c * 4 + k) * 32, make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                                                 __uint_as_float(v[4 * k + 3])));
  }
  epi_bar_sync();  // CTA-scope order of every warp's stores before the gpu-scope release below (cumulative)
  if (warp == 4 && lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flag) : "memory");
}
// slots: the workspace slot of the first partner (pairs worker + 1 ...), slot_stride floats apart
__device__ __forceinline__ void sk_reduce(const float* slots, size_t slot_stride, int partners, int* flag, uint32_t tmem_base, int acc,
                                          uint32_t acc_phase, uint64_t* tfull_bar, int warp, int lane) {
  constexpr int BN = 256, COLS = EpiCfg<BN>::COLS;
  const int q = warp & 3, part = (warp - 4) >> 2;
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  if (lane == 0) {
    while (ld_acquire_gpu(flag) < partners) __nanosleep(64);
  }
  __syncwarp();
  const float4* src = reinterpret_cast<const float4*>(slots) + (size_t)(warp - 4) * (COLS / 16) * 4 * 32 + lane;
  // two 32-column halves; the partial sums of a half (8 x 16 bytes per lane per partner) are all requested before any is used
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    float4 s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = __ldcg(src + (h * 8 + i) * 32);
    for (int pw = 1; pw < partners; ++pw) {
      const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)pw * slot_stride);
      float4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = __ldcg(s4 + (h * 8 + i) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i].x += t[i].x; s[i].y += t[i].y; s[i].z += t[i].z; s[i].w += t[i].w;
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * COLS + (h * 2 + c) * 16);
      uint32_t v[16];
      tmem_ld_32x16(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 a = s[c * 4 + k];
        v[4 * k] = __float_as_uint(__uint_as_float(v[4 * k]) + a.x);
        v[4 * k + 1] = __float_as_uint(__uint_as_float(v[4 * k + 1]) + a.y);
        v[4 * k + 2] = __float_as_uint(__uint_as_float(v[4 * k + 2]) + a.z);
        v[4 * k + 3] = __float_as_uint(__uint_as_float(v[4 * k + 3]) + a.w);
      }
      tmem_st_32x16(taddr, v);
    }
  }
  tmem_st_wait();
  epi_bar_sync();  // every epilogue warp is past its wait: the counter can go back to zero for the next launch
  if (warp == 4 && lane == 0) *reinterpret_cast<volatile int*>(flag) = 0;
}

template <int ACT, int RES, int ODT, bool SK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EpiCfg<BN2>::THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = (BN2 / 2) * BK * 2;
  constexpr bool TMAO = tma_out(true, BN2, ACT, RES, ODT);
  constexpr int STAGES = stages2(TMAO);
  constexpr int EPI_BYTES = TMAO ? EpiCfg<BN2>::TMA_BYTES : EpiCfg<BN2>::PATCH_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* epi = sB + STAGES * B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rbar = tempty + 2;   // [16] one per epilogue warp: its TMA-loaded residual box has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + EpiCfg<BN2>::WARPS);
  float* stage = reinterpret_cast<float*>(epi);
  constexpr bool TMAR = TMAO && ODT == SCB_F32 && RES == SCB_F32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMAO) tma_prefetch_desc(&tmO);
    if (TMAR) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * EpiCfg<BN2>::WARPS);
    }
    for (int s = 0; s < EpiCfg<BN2>::WARPS; ++s) mbar_init(&rbar[s], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_2sm<2 * BN2>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / cross-CTA TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // programmatic dependent launch: global memory is only touched below
  griddep_launch_dependents();

  if (warp == 0 && elect_one_sync()) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage_i = 0;
    uint32_t phase = 0;
    const uint32_t full_leader0 = mapa_u32(smem_u32(&full[0]), 0);   // the leader's full[] barriers, 8 bytes apart
    int tile, kb0, kb1;
    for (int it = 0; sk_item<SK>(p, pair, num_pairs, it, tile, kb0, kb1); ++it) {
      const TileCoord t = decode_tile2(p, tile, (int)rank);
      const int a_c0 = p.a_col0 + t.g * p.a_group_cols;
      int kin = kb0 % p.kb_per_tap;
      int a_row = t.m0 + (kb0 / p.kb_per_tap) * p.tap_row_shift, a_col = a_c0 + kin * p.bk, b_col = kb0 * p.bk;
      const int b_row = t.n0 + (int)rank * (BN2 / 2);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage_i], phase ^ 1u);
        const uint32_t full_leader = full_leader0 + (uint32_t)stage_i * 8u;
        if (leader) mbar_expect_tx(&full[stage_i], 2u * (uint32_t)(A_BYTES + B_BYTES));
        tma_load_3d_2sm(sA + stage_i * A_BYTES, &tmA, full_leader, a_col, a_row, t.b);
        tma_load_3d_2sm(sB + stage_i * B_BYTES, &tmB, full_leader, b_col, b_row, t.g);
        b_col += p.bk;
        a_col += p.bk;
        if (++kin == p.kb_per_tap) {
          kin = 0;
          a_col = a_c0;
          a_row += p.tap_row_shift;
        }
        if (++stage_i == STAGES) {
          stage_i = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && leader && elect_one_sync()) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    const uint32_t idesc = umma_idesc_f16(256, BN2, p.ab_fmt);
    const bool tf32 = p.ab_fmt == 2;
    const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
    int stage_i = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile, kb0, kb1;
    for (int it = 0; sk_item<SK>(p, pair, num_pairs, it, tile, kb0, kb1); ++it) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN2);
      uint32_t accf = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage_i], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage_i * (uint32_t)(A_BYTES >> 4), b_lo = b_lo0 + (uint32_t)stage_i * (uint32_t)(B_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (tf32) tc_mma_tf32_2sm(d_tmem, desc_join(a_lo + (uint32_t)(2 * k), kDescHi), desc_join(b_lo + (uint32_t)(2 * k), kDescHi), idesc, accf);
          else tc_mma_f16_2sm(d_tmem, desc_join(a_lo + (uint32_t)(2 * k), kDescHi), desc_join(b_lo + (uint32_t)(2 * k), kDescHi), idesc, accf);
          accf = 1;
        }
        tc_commit_2sm(&empty[stage_i], 3);
        if (++stage_i == STAGES) {
          stage_i = 0;
          phase ^= 1u;
        }
      }
      tc_commit_2sm(&tfull[acc], 3);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs; drains report to the leader)
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr size_t kSlot = (size_t)BM * BN2;  // floats per CTA slot; a pair's two slots are adjacent
    uint32_t rphase = 0;
    int tile, kb0, kb1;
    for (int it = 0; sk_item<SK>(p, pair, num_pairs, it, tile, kb0, kb1); ++it) {
      const TileCoord t = decode_tile2(p, tile, (int)rank);
      if (SK && kb0 != 0) {  // a range that does not start its tile: partial sums to the workspace, the owner finishes the tile
        const int r = tile - p.sk_full * num_pairs;
        sk_dump(p.sk_part + ((size_t)pair * 2 + rank) * kSlot, p.sk_flags + r * 2 + (int)rank, tmem_base, acc, acc_phase, &tfull[acc],
                mapa_u32(smem_u32(&tempty[acc]), 0), warp, lane);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        continue;
      }
      if (SK && kb1 != p.k_blocks) {  // owner of a tile whose tail other pairs computed
        const int r = tile - p.sk_full * num_pairs;
        sk_reduce(p.sk_part + ((size_t)(pair + 1) * 2 + rank) * kSlot, 2 * kSlot, sk_partners(p, pair, num_pairs, r),
                  p.sk_flags + r * 2 + (int)rank, tmem_base, acc, acc_phase, &tfull[acc], warp, lane);
      }
      if constexpr (TMAO && ODT == SCB_F32) epilogue_tile_tma_res<RES>(p, &tmO, &tmR, t, tmem_base, acc, acc_phase, &tfull[acc], mapa_u32(smem_u32(&tempty[acc]), 0), epi, &rbar[warp - 4], rphase, warp, lane);
      else if constexpr (TMAO) epilogue_tile_tma<ACT>(p, &tmO, t, tmem_base, acc, acc_phase, &tfull[acc], mapa_u32(smem_u32(&tempty[acc]), 0), epi, warp, lane);
      else epilogue_tile<BN2, ACT, RES, ODT>(p, t, tmem_base, acc * BN2, acc_phase, &tfull[acc], mapa_u32(smem_u32(&tempty[acc]), 0), stage, warp, lane);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (TMAO && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its shared memory until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<2 * BN2>(tmem_base);
  }
}

template <int ACT, int RES, int ODT, bool SK>
int launch2_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const GemmParams& p,
                 cudaStream_t stream) {
  constexpr bool TMAO = tma_out(true, BN2, ACT, RES, ODT);
  constexpr int smem_bytes = stages2(TMAO) * (BM * BK * 2 + (BN2 / 2) * BK * 2) + 1024 + 512 +
                             (TMAO ? EpiCfg<BN2>::TMA_BYTES : EpiCfg<BN2>::PATCH_BYTES);
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static bool configured = false;
  if (!configured) {
    SCB_CUDA(cudaFuncSetAttribute(gemm2_tcgen05_kernel<ACT, RES, ODT, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  int pairs = num_sms() / 2;
  if (p.num_tiles < pairs && !SK) pairs = p.num_tiles;
  SCB_CUDA(launch_pdl(gemm2_tcgen05_kernel<ACT, RES, ODT, SK>, dim3((unsigned)(2 * pairs)), EpiCfg<BN2>::THREADS, smem_bytes, stream, tmA, tmB, tmO, tmR, p));
  note_launch();
  SCB_LAUNCH_OK("gemm2_tcgen05");
  return SCB_OK;
}
// the stream-K variant is a separate instantiation: its dump / reduce paths cost the plain kernel neither registers nor branches
template <int ACT, int RES, int ODT>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, const GemmParams& p,
            cudaStream_t stream) {
  return p.sk_rem ? launch2_impl<ACT, RES, ODT, true>(tmA, tmB, tmO, tmR, p, stream) : launch2_impl<ACT, RES, ODT, false>(tmA, tmB, tmO, tmR, p, stream);
}

template <int BN, int STAGES, int ACT = -1, int RES = -1, int ODT = -1>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmParams& p, cudaStream_t stream) {
  constexpr bool TMAO = tma_out(false, BN, ACT, RES, ODT);
  constexpr int smem_bytes = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256 + (TMAO ? EpiCfg<BN>::TMA_BYTES : EpiCfg<BN>::PATCH_BYTES);
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static bool configured = false;
  if (!configured) {
    SCB_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES, ACT, RES, ODT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  SCB_CUDA(launch_pdl(gemm_tcgen05_kernel<BN, STAGES, ACT, RES, ODT>, dim3((unsigned)grid), EpiCfg<BN>::THREADS, smem_bytes, stream, tmA, tmB, tmO, p));
  note_launch();
  SCB_LAUNCH_OK("gemm_tcgen05");
  return SCB_OK;
}

}  // namespace

long long gemm_workspace_bytes() { return kSkFlagBytes + (long long)(num_sms() / 2) * 2 * BM * 256 * (long long)sizeof(float); }

int gemm(const scb_gemm_args& a, cudaStream_t stream) {
  SCB_CHECK(a.a && a.b && a.out, SCB_EINVAL, "scb_gemm: null operand");
  SCB_CHECK(a.ab_format == SCB_F16 || a.ab_format == SCB_BF16 || a.ab_format == SCB_F32, SCB_EINVAL,
            "scb_gemm: ab_format must be F16, BF16 or F32 (fp32 operands run as TF32)");
  const int eb = a.ab_format == SCB_F32 ? 4 : 2;  // operand element bytes
  const int al = 16 / eb;                          // elements per 16 bytes
  SCB_CHECK(a.n > 0 && a.k > 0 && a.batch > 0 && a.m_per_batch > 0 && a.groups > 0, SCB_EINVAL, "scb_gemm: empty problem");
  SCB_CHECK(a.n % 8 == 0, SCB_EINVAL, "scb_gemm: n (%d) must be a multiple of 8", a.n);
  SCB_CHECK(a.ldc % 8 == 0 && a.out_group_cols % 8 == 0 && a.out_batch_stride % 8 == 0, SCB_EINVAL,
            "scb_gemm: output strides must be multiples of 8 elements");
  SCB_CHECK(a.a_row_stride % al == 0 && a.b_row_stride % al == 0 && a.a_batch_stride % al == 0 && a.b_group_stride % al == 0,
            SCB_EINVAL, "scb_gemm: operand strides must be multiples of 16 bytes");
  SCB_CHECK((reinterpret_cast<uintptr_t>(a.a) | reinterpret_cast<uintptr_t>(a.b) | reinterpret_cast<uintptr_t>(a.out) |
             reinterpret_cast<uintptr_t>(a.out2) | reinterpret_cast<uintptr_t>(a.residual) |
             reinterpret_cast<uintptr_t>(a.bias)) % 16 == 0,
            SCB_EINVAL, "scb_gemm: pointers must be 16-byte aligned");
  SCB_CHECK(a.kb_per_tap > 0, SCB_EINVAL, "scb_gemm: kb_per_tap must be positive");
  SCB_CHECK(a.groups == 1 || a.out_group_cols >= a.n, SCB_EINVAL, "scb_gemm: out_group_cols < n");

  // Tile shape.  256-wide tiles (CTA pairs: 256 x 256) carry the large problems.  When M is small (CLIP ViT and the head at 32-64
  // pairs per GPU: strong scaling) the choice is between few large tiles on part of the SMs and many small tiles on all of
  // them; small tiles pull (128 + BN) x 128 B per k-block for 128 x BN x 64 MACs, so they are bound by the L2 -> SM path, and
  // a tile count just above the SM count (156 tiles of 128 x 64 for M = 1600, N = 768) costs a second wave.  The loop below
  // estimates each candidate with a two-term model (MMA clocks vs operand bytes over min(per-SM, chip / active SMs) bandwidth).
  static const int two_env = [] { const char* e = getenv("SCB_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  static const int force_bn = [] { const char* e = getenv("SCB_GEMM_FORCE_BN"); return e ? atoi(e) : 0; }();      // tuning sweeps
  static const int force_two = [] { const char* e = getenv("SCB_GEMM_FORCE_2CTA"); return e ? atoi(e) : -1; }();
  const long long mt1 = (long long)a.batch * a.groups * ((a.m_per_batch + BM - 1) / BM);
  const long long mt2 = (long long)a.batch * a.groups * ((a.m_per_batch + 255) / 256);
  // (tap-walk conv GEMMs: pairs measured +5 % on the long layers (10239 / 5119 rows per utterance), -5 % at 639 rows, where the
  // 256-row super tiles pad 17 %)
  const bool two_ok = two_env != 0 && a.n > 128 && a.m_per_batch >= 256 && (a.tap_row_shift == 0 || a.m_per_batch >= 2048);
  const int bn_max = a.n > 128 ? 256 : (a.n > 64 ? 128 : 64);
  static const int sk_env = [] { const char* e = getenv("SCB_GEMM_STREAMK"); return e ? atoi(e) : 1; }();
  const bool sk_ok = sk_env != 0 && a.workspace != nullptr && a.workspace_bytes >= gemm_workspace_bytes();
  int bn = bn_max;
  bool two = false;
  {
    const int kb = (a.k + (BK * 2 / eb) - 1) / (BK * 2 / eb);
    const int sms = num_sms();
    double best = 1e30;
    for (int cand = 0; cand < 4; ++cand) {  // 0..2: single CTA with BN = 64 / 128 / 256; 3: CTA pair 256 x 256
      const int cbn = cand == 3 ? 256 : (64 << cand);
      if (cand < 3 && cbn > bn_max) continue;        // never wider than the padded N
      if (cand == 3 && !two_ok) continue;
      const long long tiles = (cand == 3 ? mt2 : mt1) * ((a.n + cbn - 1) / cbn);
      const int slots = cand == 3 ? sms / 2 : sms;
      double waves = (double)((tiles + slots - 1) / slots);
      double active = (double)(tiles < slots ? tiles : slots) * (cand == 3 ? 2 : 1);
      if (cand == 3 && sk_ok && stream_k_applies(tiles, slots, kb)) {  // the last wave is cut along K over all pairs
        waves = (double)(tiles / slots) + (double)(tiles % slots) / slots + 0.35;  // + dump / reduce of the split tiles
        active = sms;
      }
      const double bw = fmin(64.0, 8500.0 / active);                                   // bytes per clock per CTA
      const double bytes = (cand == 3 ? (BM + 128) : (BM + cbn)) * 128.0;             // per CTA per k-block
      const double mma = cbn * 2.0;                                                    // 128 x BN x 64 MACs at 4096 MAC / clk
      // (64-wide tiles measured ~1.6x slower per k-block than their byte count predicts)
      const double t = 2000.0 + waves * (kb * fmax(mma, (cbn == 64 ? 1.6 : 1.0) * bytes / bw) + 600.0) + cbn * 8.0;  // fill + main loops + last epilogue
      if (t < best) {
        best = t;
        bn = cbn;
        two = cand == 3;
      }
    }
    if (force_bn) bn = force_bn;
    if (force_two >= 0) two = force_two != 0 && a.m_per_batch >= 256 && (a.tap_row_shift == 0 || force_two == 2);
    if (two) bn = 256;
  }
  GemmParams p{};
  p.batch = a.batch;
  p.m_tiles_per_batch = two ? (a.m_per_batch + 255) / 256 : (a.m_per_batch + BM - 1) / BM;
  p.n_tiles = (a.n + bn - 1) / bn;
  p.groups = a.groups;
  p.num_tiles = p.groups * p.batch * p.m_tiles_per_batch * p.n_tiles;
  p.m_per_batch = a.m_per_batch;
  p.n = a.n;
  p.bk = BK * 2 / eb;
  p.k_blocks = (a.k + p.bk - 1) / p.bk;
  p.kb_per_tap = a.kb_per_tap;
  p.tap_row_shift = a.tap_row_shift;
  p.a_col0 = a.a_col0;
  p.a_group_cols = a.a_group_cols;
  if (two && sk_ok && stream_k_applies(p.num_tiles, num_sms() / 2, p.k_blocks)) {
    const int W = num_sms() / 2;
    p.sk_full = p.num_tiles / W;
    p.sk_rem = p.num_tiles % W;
    p.sk_flags = static_cast<int*>(a.workspace);
    p.sk_part = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + kSkFlagBytes);
  }
  static const int dbg_env = [] { const char* e = getenv("SCB_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg_env)
    fprintf(stderr, "scb_gemm m=%d x%d n=%d k=%d: bn=%d pair=%d tiles=%d stream-k full=%d rem=%d\n", a.m_per_batch, a.batch, a.n, a.k, bn,
            (int)two, p.num_tiles, p.sk_full, p.sk_rem);
  p.umma_n = two ? 128 : ((a.n < bn ? a.n : bn) + 15) / 16 * 16;  // rows of the B box (cta_group::2: half of the 256-wide tile per CTA)
  p.slab = (!two && eb == 2 && bn == 64 && a.kb_per_tap == 1 && a.tap_row_shift == 1 && p.k_blocks + BM - 1 <= 256) ? 1 : 0;
  p.slab_sub_bytes = p.umma_n * BK * 2;  // 48 x 128 B = 6 KB (base) / 8 KB (large): multiples of the 1024-byte swizzle atom
  if (p.slab && p.slab_sub_bytes % 1024 != 0) p.slab = 0;
  p.tile_m = BM;
  p.slab_mt = 1;
  p.slab_taps = kSlabTaps;
  p.slab_parts = 1;
  p.slab_box_rows = 256;
  static const int mt_env = [] { const char* e = getenv("SCB_GEMM_SLAB_MT"); return e ? atoi(e) : 3; }();
  if (p.slab && p.m_tiles_per_batch >= 2 && p.m_tiles_per_batch <= mt_env && p.m_tiles_per_batch <= 3) {
    // every M tile of a batch entry in one work item: slab of mt * 128 + taps - 1 rows in two boxes
    p.slab_mt = p.m_tiles_per_batch;
    const int rows = p.slab_mt * BM + p.k_blocks - 1;
    p.slab_parts = 2;
    p.slab_box_rows = ((rows + 1) / 2 + 7) / 8 * 8;
    p.tile_m = BM * p.slab_mt;
    p.m_tiles_per_batch = 1;
    p.num_tiles = p.groups * p.batch * p.n_tiles;
  }
  p.slab_bytes = p.slab_parts * p.slab_box_rows * BK * 2;
  {
    const int ring = 8 * (BM * BK * 2 + 64 * BK * 2);   // the BN = 64 kernel's A + B rings (8 stages)
    const int region = ring - 2 * p.slab_bytes;          // what the two slabs leave for the weight sub-tiles
    if (p.slab && region < 3 * kSlabTaps * p.slab_sub_bytes) p.slab_taps = 2;   // fewer taps per stage rather than a 2-deep ring
    p.slab_stages = 8;
    while (p.slab && p.slab_stages > 1 && p.slab_stages * p.slab_taps * p.slab_sub_bytes > region) --p.slab_stages;
  }
  p.tx_bytes = (uint32_t)(BM * BK * 2 + p.umma_n * BK * 2);
  p.out = a.out;
  p.out2 = a.out2;
  p.bias = a.bias;
  p.residual = a.residual;
  p.out_dtype = a.out_dtype;
  p.out2_dtype = a.out2_dtype;
  p.residual_dtype = a.residual_dtype;
  p.act = a.act;
  p.ab_fmt = a.ab_format == SCB_F32 ? 2 : (a.ab_format == SCB_BF16 ? 1 : 0);
  p.alpha = a.alpha;
  p.ldc = a.ldc;
  p.out_batch_stride = a.out_batch_stride;
  p.out_group_cols = a.out_group_cols;
  p.res_ld = a.residual_ld ? a.residual_ld : a.ldc;
  p.res_batch_stride = a.residual_ld ? a.residual_batch_stride : a.out_batch_stride;
  SCB_CHECK(p.res_ld % 8 == 0 && p.res_batch_stride % 8 == 0, SCB_EINVAL, "scb_gemm: residual strides must be multiples of 8 elements");

  CUtensorMap tmA, tmB;
  {
    const uint64_t rows = (uint64_t)a.a_rows;
    const uint64_t bstride = a.a_batch_stride ? (uint64_t)a.a_batch_stride : rows * (uint64_t)a.a_row_stride;
    const uint64_t dims[3] = {(uint64_t)a.a_inner, rows, (uint64_t)a.batch};
    const uint64_t strides[2] = {(uint64_t)a.a_row_stride * eb, bstride * eb};
    const uint32_t box[3] = {(uint32_t)p.bk, (uint32_t)(p.slab ? p.slab_box_rows : BM), 1};
    int e = make_tmap(&tmA, a.a, eb, 3, dims, strides, box, 1);
    if (e) return e;
  }
  {
    const uint64_t gstride = a.b_group_stride ? (uint64_t)a.b_group_stride : (uint64_t)a.n * (uint64_t)a.b_row_stride;
    const uint64_t dims[3] = {(uint64_t)a.k, (uint64_t)a.n, (uint64_t)a.groups};
    const uint64_t strides[2] = {(uint64_t)a.b_row_stride * eb, gstride * eb};
    const uint32_t box[3] = {(uint32_t)p.bk, (uint32_t)p.umma_n, 1};
    int e = make_tmap(&tmB, a.b, eb, 3, dims, strides, box, 1);
    if (e) return e;
  }
  // epilogue specialisations for the shapes that carry the step (compile-time activation / residual / output type)
  int mode = 0;
  if (!a.out2 && bn == 256) {
    const bool nores = a.residual == nullptr;
    if (a.act == SCB_ACT_NONE && nores && a.out_dtype == SCB_F16) mode = 1;                                                    // QKV, K/V
    else if (a.act == SCB_ACT_NONE && !nores && a.residual_dtype == SCB_F32 && a.out_dtype == SCB_F32) mode = 2;               // out-proj, fc2
    else if (a.act == SCB_ACT_GELU_ERF && nores && a.out_dtype == SCB_F16) mode = 3;                                            // fc1, conv1..6
    else if (a.act == SCB_ACT_QUICK_GELU && nores && a.out_dtype == SCB_F16) mode = 4;                                          // CLIP c_fc
    else if (a.act == SCB_ACT_NONE && !nores && a.residual_dtype == SCB_F16 && a.out_dtype == SCB_F32) mode = 5;               // out-proj, fc2 (16-bit residual stream)
  }
  // modes 1, 3, 4 store through TMA: [batch][rows][cols] view of the output, 32-row x 64-column boxes in the 128B swizzle
  static const int tma_env = [] { const char* e = getenv("SCB_GEMM_TMA_STORE"); return e ? atoi(e) : 1; }();
  if (mode != 0 && (a.groups != 1 || tma_env == 0)) mode = 0;   // (the generic register-store epilogue handles everything)
  // fp32 residual by TMA (mode 2 on pairs) needs a residual laid out like the output: a [m_per_batch][n] table broadcast over the
  // batch (stride 0: the patch-embedding GEMM's positional table) stays on the register-store path
  if (mode == 2 && two && a.batch > 1 && a.residual_ld != 0 && a.residual_batch_stride == 0) mode = 0;
  CUtensorMap tmO = tmA, tmR = tmA;
  if (mode == 2 && two) {
    const uint64_t dims[3] = {(uint64_t)a.n, (uint64_t)a.m_per_batch, (uint64_t)a.batch};
    const uint64_t bstride = p.res_batch_stride ? (uint64_t)p.res_batch_stride : (uint64_t)a.m_per_batch * (uint64_t)p.res_ld;
    const uint64_t strides[2] = {(uint64_t)p.res_ld * 4, bstride * 4};
    const uint32_t box[3] = {32, 32, 1};
    int e = make_tmap(&tmR, a.residual, 4, 3, dims, strides, box, 1);
    if (e) return e;
  }
  if (mode == 1 || mode == 3 || mode == 4 || ((mode == 5 || mode == 2) && two)) {  // 32-row boxes of 128 bytes: 64 16-bit or 32 fp32 columns
    const int ob = a.out_dtype == SCB_F32 ? 4 : 2;
    const uint64_t dims[3] = {(uint64_t)a.n, (uint64_t)a.m_per_batch, (uint64_t)a.batch};
    const uint64_t bstride = a.out_batch_stride ? (uint64_t)a.out_batch_stride : (uint64_t)a.m_per_batch * (uint64_t)a.ldc;
    const uint64_t strides[2] = {(uint64_t)a.ldc * ob, bstride * ob};
    const uint32_t box[3] = {(uint32_t)(128 / ob), 32, 1};
    int e = make_tmap(&tmO, a.out, ob, 3, dims, strides, box, 1);
    if (e) return e;
  }
  if (two) {
    switch (mode) {
      case 1: return launch2<SCB_ACT_NONE, 3, SCB_F16>(tmA, tmB, tmO, tmR, p, stream);
      // fp32 residual: by TMA through the staging tile for short K, where the epilogue bounds the tile (ViT out-proj 38.8 -> 32.4 us,
      // HuBERT-large out-proj 61.5 -> 58.0); for long K the main loop bounds it and the 6th pipeline stage is worth more than
      // the epilogue (fc2 at K = 3072 / 4096 measured 5-6 % slower on the TMA path), so those keep the register-store epilogue
      case 2:
        return p.k_blocks <= 24 ? launch2<SCB_ACT_NONE, SCB_F32, SCB_F32>(tmA, tmB, tmO, tmR, p, stream)
                                : launch2<SCB_ACT_NONE, 16 + SCB_F32, SCB_F32>(tmA, tmB, tmO, tmR, p, stream);
      case 3: return launch2<SCB_ACT_GELU_ERF, 3, SCB_F16>(tmA, tmB, tmO, tmR, p, stream);
      case 4: return launch2<SCB_ACT_QUICK_GELU, 3, SCB_F16>(tmA, tmB, tmO, tmR, p, stream);
      case 5: return launch2<SCB_ACT_NONE, SCB_F16, SCB_F32>(tmA, tmB, tmO, tmR, p, stream);
      default: return launch2<-1, -1, -1>(tmA, tmB, tmO, tmR, p, stream);
    }
  }
  if (bn == 256) {
    switch (mode) {
      case 1: return launch<256, 3, SCB_ACT_NONE, 3, SCB_F16>(tmA, tmB, tmO, p, stream);
      case 2: return launch<256, 4, SCB_ACT_NONE, SCB_F32, SCB_F32>(tmA, tmB, tmO, p, stream);
      case 3: return launch<256, 3, SCB_ACT_GELU_ERF, 3, SCB_F16>(tmA, tmB, tmO, p, stream);
      case 4: return launch<256, 3, SCB_ACT_QUICK_GELU, 3, SCB_F16>(tmA, tmB, tmO, p, stream);
      case 5: return launch<256, 4, SCB_ACT_NONE, SCB_F16, SCB_F32>(tmA, tmB, tmO, p, stream);
      default: return launch<256, 4>(tmA, tmB, tmO, p, stream);
    }
  }
  if (bn == 128) return launch<128, 6>(tmA, tmB, tmO, p, stream);
  return launch<64, 8>(tmA, tmB, tmO, p, stream);
}

}  // namespace scb
