/*
 * speechclip_b200.h — C ABI of libspeechclip_b200.so (sm_100a CUDA kernels for the SpeechCLIP
 * speech–image contrastive forward / training step).
 *
 * The reference (atosystem/SpeechCLIP) is pure Python and owns no kernels or FFI: every GPU
 * operation on its hot path is a PyTorch library call.  This header is therefore
 * the boundary a maintainer would bind INSTEAD of those calls; each entry point cites the reference
 * call site (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes stub and where each call goes in the reference's modules.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - every function returns 0 on success, a negative SCB_E* code otherwise; scb_last_error()
 *     returns a thread-local message for the last failure on the calling thread;
 *   - nothing here allocates or frees caller-visible memory; scratch is passed in by the caller;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous;
 *   - 16-bit tensors are SCB_F16 (IEEE half: forward activations and weights) or SCB_BF16
 *     (gradient operands); accumulation, normalisation statistics, softmax and the loss are fp32.
 */
#ifndef SPEECHCLIP_B200_H
#define SPEECHCLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_ABI_VERSION 1

enum { SCB_OK = 0, SCB_EINVAL = -1, SCB_ECUDA = -2, SCB_EUNSUPPORTED = -3 };
enum { SCB_F32 = 0, SCB_F16 = 1, SCB_BF16 = 2 };
enum { SCB_ACT_NONE = 0, SCB_ACT_GELU_ERF = 1, SCB_ACT_QUICK_GELU = 2 };

int scb_abi_version(void);
const char* scb_last_error(void);
/* Number of kernels this library has launched on the calling process since load (bench.py's gpu_launches). */
int64_t scb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores (TMA -> smem -> UMMA -> TMEM -> fused epilogue).
 *   out[b, m, g*out_group_cols + n] = epilogue( sum_k A[b, m(+tap), k] * B[g, n, k] )
 *   epilogue(x) = act(alpha * x + bias[g*out_group_cols + n]) + residual[same index as out]
 * Replaces every nn.Linear / Conv1d-as-GEMM / patchify on the path:
 *   fairseq q/k/v/out_proj, fc1, fc2, post_extract_proj (driven from speech_encoder_plus.py:49-53,84-85),
 *   fairseq conv1..6 and pos_conv (speech_encoder_plus.py:75,35), CLIP conv1 / in_proj / out_proj /
 *   c_fc / c_proj / proj (clip_official.py:209), nn.TransformerEncoderLayer linears
 *   (kw_modules/TransformerModels.py:64-75), linear_proj (kwClip.py:1105-1106), and their dgrad/wgrad.
 *
 * A is a 16-bit tensor viewed as [batch][a_rows][a_inner] (strides in elements, a_inner contiguous).
 * K is walked in 64-element blocks; block kb reads A columns
 *   a_col0 + g*a_group_cols + (kb % kb_per_tap)*64 ... +63   of row   m + (kb / kb_per_tap) * tap_row_shift
 * which expresses a plain GEMM (kb_per_tap = K/64, shift 0), a strided Conv1d over channel-last
 * activations (pairs of input frames viewed as one row) and the grouped positional conv (one tap per
 * block, shift 1) without an im2col buffer.  Rows/columns outside the tensor read as zero.
 * B is [groups][n][k] 16-bit (torch Linear layout: k contiguous).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scb_gemm_args {
  const void* a;
  int64_t a_inner, a_rows, a_row_stride, a_batch_stride;
  int32_t batch, m_per_batch;
  int32_t kb_per_tap, tap_row_shift, a_col0, a_group_cols;
  const void* b;
  int64_t b_row_stride, b_group_stride;
  int32_t n, k, groups;
  void* out;            /* [batch][m_per_batch][ldc] */
  int32_t out_dtype;    /* SCB_F32 | SCB_F16 | SCB_BF16 */
  int32_t out_group_cols;
  int64_t ldc, out_batch_stride;
  void* out2;           /* optional second copy of the result (same layout), may be NULL */
  int32_t out2_dtype;
  int32_t ab_format;    /* SCB_F16 | SCB_BF16: format of A and B */
  const float* bias;    /* optional, fp32, indexed by output column */
  const void* residual; /* optional, same layout as out */
  int32_t residual_dtype;
  int32_t act;
  float alpha;
} scb_gemm_args;

int scb_gemm(const scb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHCLIP_B200_H */
