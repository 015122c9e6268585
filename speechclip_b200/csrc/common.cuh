// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers, 16-bit
// conversions, warp reductions, error plumbing.  No CUTLASS: everything is inline PTX.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include "../../include/speechclip_b200.h"

namespace scb {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define SCB_CHECK(cond, code, ...)        \
  do {                                    \
    if (!(cond)) {                        \
      ::scb::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)
#define SCB_CUDA(call)                                   \
  do {                                                   \
    int _e = ::scb::check_cuda((call), #call);           \
    if (_e) return _e;                                   \
  } while (0)
#define SCB_LAUNCH_OK(name) SCB_CUDA(cudaGetLastError())

int num_sms();
void note_launch();  // bumps the per-process kernel-launch counter (scb_launch_count)
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1; 2- or 4-byte elements.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle128);

// Kernel launch with programmatic stream serialization (SCB_PDL=0 disables): the kernel may start its prologue while the
// previous kernel of the stream is still draining and blocks in griddep_wait() until that kernel's memory is visible.
template <typename K, typename... Args>
inline cudaError_t launch_pdl(K kernel, dim3 grid, int block, int smem, cudaStream_t stream, Args... args) {
  static const int pdl = [] { const char* e = getenv("SCB_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---------------------------------------------------------------- 16-bit formats
// SCB_F16 = IEEE half (forward activations / weights), SCB_BF16 = bfloat16 (gradient operands).
template <int FMT> struct H16;
template <> struct H16<SCB_F16> {
  using T = __half;
  using T2 = __half2;
  __device__ static __forceinline__ T from(float x) { return __float2half_rn(x); }
  __device__ static __forceinline__ float to(T x) { return __half2float(x); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
};
template <> struct H16<SCB_BF16> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ T from(float x) { return __float2bfloat16_rn(x); }
  __device__ static __forceinline__ float to(T x) { return __bfloat162float(x); }
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
};

__device__ __forceinline__ uint32_t pack16(int fmt, float a, float b) {
  return fmt == SCB_BF16 ? H16<SCB_BF16>::pack(a, b) : H16<SCB_F16>::pack(a, b);
}
__device__ __forceinline__ float2 unpack16(int fmt, uint32_t u) {
  return fmt == SCB_BF16 ? H16<SCB_BF16>::unpack(u) : H16<SCB_F16>::unpack(u);
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// erf-GELU through the Abramowitz-Stegun 7.1.26 rational approximation of erfc (|error| < 5e-7 absolute on the GELU value):
// 2 MUFU + ~13 FP32 instructions, branch-free — for epilogues that evaluate billions of activations per step.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float h = p * t * ex2_approx(-1.4426950408889634f * z * z);  // Phi(-|x|)
  return x * (x >= 0.f ? 1.0f - h : h);
}
// erf-GELU for 16-bit outputs: x * sigmoid(x * (c0 + c1 s + c2 s^2)), s = min(x^2, 50) — a minimax fit of Phi(x) in that
// form (|error| <= 2.6e-5 absolute on the GELU value over the whole real line, an order below the half-ulp of an fp16 activation
// of ordinary size; the clamp keeps the negative c2 from turning the polynomial over beyond |x| = 7, where Phi is 0 / 1 to
// 1e-12): 6 FP32 + 2 MUFU instructions against 13 + 2 for gelu_fast.  The constants carry the -log2(e) of the exponential.
__device__ __forceinline__ float gelu_h16(float x) {
  const float s = fminf(x * x, 50.0f);
  float p = fmaf(-0.0007030335785217694f * -1.4426950408889634f, s, 0.07401129204959998f * -1.4426950408889634f);
  p = fmaf(p, s, 1.5950157685602808f * -1.4426950408889634f);
  return x * rcp_approx(1.0f + ex2_approx(p * x));
}
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two lanes of arithmetic per issue slot) for the GEMM epilogues.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// gelu_h16 on a pair: 5 packed FP32 + 2 FMNMX + 4 MUFU for two activations.
__device__ __forceinline__ uint64_t gelu_h16_x2(uint64_t x) {
  constexpr float kL = -1.4426950408889634f;
  float s0, s1;
  upk2(mul2(x, x), s0, s1);
  const uint64_t s = pk2(fminf(s0, 50.0f), fminf(s1, 50.0f));
  uint64_t p = fma2(s, pk2(-0.0007030335785217694f * kL, -0.0007030335785217694f * kL),
                    pk2(0.07401129204959998f * kL, 0.07401129204959998f * kL));
  p = fma2(p, s, pk2(1.5950157685602808f * kL, 1.5950157685602808f * kL));
  float t0, t1;
  upk2(mul2(p, x), t0, t1);
  float d0, d1;
  upk2(add2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(1.0f, 1.0f)), d0, d1);
  return mul2(x, pk2(rcp_approx(d0), rcp_approx(d1)));
}
__device__ __forceinline__ uint64_t quick_gelu_x2(uint64_t x) {
  constexpr float kQ = -1.702f * 1.4426950408889634f;
  float t0, t1;
  upk2(mul2(x, pk2(kQ, kQ)), t0, t1);
  float d0, d1;
  upk2(add2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(1.0f, 1.0f)), d0, d1);
  return mul2(x, pk2(rcp_approx(d0), rcp_approx(d1)));
}
__device__ __forceinline__ float quick_gelu(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * x)); }

// ---------------------------------------------------------------- counter-based RNG for dropout
// Philox-4x32-10 (Salmon et al., SC'11).  A dropout decision is a pure function of (seed, step, site, element index), so the
// backward pass regenerates the forward's mask instead of storing it, and a test can materialise the same mask for the oracle.
//   rng_state (device int64[2]) = {seed, step};  counter = (index >> 2 [64 bit], site, step), key = seed, lane = index & 3.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct DropoutRng {
  uint2 key;
  uint32_t site, step, threshold;
  float keep_scale;
  // p in [0, 1): an element is dropped when its 32-bit draw is below p * 2^32; survivors are scaled by 1 / (1 - p)
  __device__ __forceinline__ DropoutRng(const long long* state, int site_, float p) {
    const unsigned long long seed = state ? (unsigned long long)state[0] : 0ull;   // NULL: caller runs with p = 0 and never draws
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    step = state ? (uint32_t)state[1] : 0u;
    site = (uint32_t)site_;
    threshold = (uint32_t)fminf(p * 4294967296.f, 4294967040.f);
    keep_scale = 1.f / (1.f - p);
  }
  __device__ __forceinline__ float scale(unsigned long long idx) const {
    const unsigned long long blk = idx >> 2;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), site, step), key);
    const uint32_t lane = (uint32_t)idx & 3u;
    const uint32_t u = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    return u >= threshold ? keep_scale : 0.f;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive on a barrier given by its shared::cluster address (own CTA or, after mapa, a peer CTA of the cluster).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a system-dependent time when the phase is still open: a thread that
// polls SEVERAL barriers must use this form, or an arrival on one barrier goes unnoticed while it sleeps on another).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (→ cudaErrorLaunchFailure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s at 1.9 GHz
      printf("scb: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// Programmatic dependent launch: block until the kernels this launch depends on have completed and flushed (a no-op when the
// kernel was launched without the programmatic-serialization attribute).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ... and let the next kernel of the stream start launching (its own griddep_wait still waits for this grid to complete).
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// cta_group::2 TMA load: the data lands in THIS CTA's shared memory, the transaction bytes are signalled on an mbarrier given by
// its shared::cluster address (the pair leader's "full" barrier).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA store of a staged shared-memory tile (bulk-group completion): the writer threads fence the async proxy, one lane issues.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// arrives (once all prior MMAs of this thread retire) on the barrier at this CTA-relative address in every CTA of cta_mask
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T : M = 256 across the CTA pair
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Register re-allocation between warpgroups (4 consecutive warps; every thread of the warpgroup must execute the instruction):
// the data-movement warpgroup hands registers to the compute warpgroups.  N: multiple of 8 in [24, 256].
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // arrives on bar when all prior MMAs of this thread retire
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 16-bit inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with fp32 operands read as TF32 (K = 8 per instruction): the fp32 rows of the trainable head need no 16-bit copies.
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// registers -> TMEM, 32 lanes x 16 consecutive fp32 columns (thread t writes row lane base + t); tmem_st_wait() before the data is
// handed to another agent.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B (what a TMA box {64 x rows} of 16-bit
// elements with CU_TENSOR_MAP_SWIZZLE_128B writes): 8-row atoms of 1024 B, SBO = 1024 B, descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                 // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// One lane of a fully converged warp (elect.sync): the compiler then knows the code it guards runs on a single thread and issues
// tcgen05.mma / commit directly, instead of wrapping each of them in an "elect / issue / any lane left?" loop as it does under
// `lane == 0` (5 extra instructions per MMA on the issuing thread).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// kind::f16 instruction descriptor: fp32 accumulator, A/B both K-major, format 0 = f16 / 1 = bf16.
// (kind::tf32 uses the same layout with format 2 = tf32.)
__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n, int fmt) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= (uint32_t)fmt << 7;      // a_format
  d |= (uint32_t)fmt << 10;     // b_format
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

}  // namespace scb
