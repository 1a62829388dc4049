/*
 * libtcavp — C ABI of the B200-native (sm_100a) forward hot path of the Traffic-Context-Augmented
 * Vehicle Trajectory Prediction model.
 *
 * The reference (imjaegyun/Traffic-Context-Augmented-...-Multimodal-LLM) has no FFI/plugin layer: its
 * boundary is the Python class `MultiModalTrajectoryModel` (reference scripts/train.py:847-964).  This
 * header is the C boundary *beneath* our Python mirror of that class; each entry point names the
 * reference call site (file:line, relative to the reference root; "HF:" = transformers
 * models/llama/modeling_llama.py) whose arithmetic it replaces.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.  The caller owns every buffer;
 *     the library allocates no device memory and keeps no state besides a per-process cache of the
 *     driver entry point used to encode TMA descriptors.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     returns 0 on success or a negative TCAVP_ERR_* code; tcavp_last_error() gives the thread-local text.
 *   - dtype codes: TCAVP_F32 (fp32 storage, fp32 SIMT math) / TCAVP_BF16 (bf16 storage, fp32 accumulate).
 *   - Matrices are row-major; "ld*" are leading dimensions in ELEMENTS.  Weights keep nn.Linear's
 *     [out_features, in_features] layout so state_dict tensors are consumed without transposition.
 */
#ifndef TCAVP_H_
#define TCAVP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tcavp_stream_t; /* cudaStream_t */

enum { TCAVP_F32 = 0, TCAVP_BF16 = 1 };
enum { TCAVP_OK = 0, TCAVP_ERR_ARG = -1, TCAVP_ERR_CUDA = -2, TCAVP_ERR_UNSUPPORTED = -3 };
enum { TCAVP_ACT_NONE = 0, TCAVP_ACT_RELU = 1, TCAVP_ACT_SWIGLU = 2 };

/* ---- library ------------------------------------------------------------------------------- */
const char* tcavp_last_error(void);
int tcavp_version(void);
/* Fills SM count and compute capability of the current device. */
int tcavp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- dense contraction with fused epilogue --------------------------------------------------
 * out[m', n] = act( scale * sum_k A[m,k] * W[n,k] + bias[n] + LoRA(m,n) ) + residual[m', n]
 *
 * Replaces every nn.Linear / F.linear on the path: HF LlamaAttention q/k/v/o_proj (HF:251-289) and
 * LlamaMLP gate/up/down (HF:182), peft lora.Linear (y = Wx + (alpha/r) B(A x), train.py:432-440),
 * nn.MultiheadAttention / nn.Transformer* in/out projections and FFNs (train.py:358-359, 402-406,
 * 663-670, 754), q_proj (train.py:521), lane_fc / post_mlp (train.py:784-790).
 *
 *   in_dtype   TCAVP_BF16: A and W are bf16, TMA-fed tcgen05.mma with TMEM accumulators (fp32);
 *              requires K % 8 == 0 and 16-byte aligned A/W rows.
 *              TCAVP_F32 : A and W are fp32, SIMT FFMA kernel (exact fp32 accumulate, no TF32).
 *   act        RELU, or SWIGLU: W rows are interleaved (gate_0, up_0, gate_1, up_1, ...), N counts the
 *              interleaved rows, the output has N/2 columns: silu(gate_j) * up_j  (HF:190).
 *   LoRA       lora_r > 0: adds sum_k lora_t[m, toff + k] * lora_b[n, k] for columns n inside
 *              [seg_begin[i], seg_end[i]) (i = 0,1; toff = seg_toff[i]).  lora_t = x·A^T (fp32,
 *              leading dim lora_ldt) comes from a skinny GEMM over the same x; lora_b = (alpha/r)·B
 *              as fp32 [N, r].
 *   row remap  remap_gi > 0: m' = (m / remap_gi) * remap_go + (m % remap_gi) + remap_off (writes the
 *              16 image-token rows of each scene straight into the fused (B, L, H) buffer,
 *              train.py:521-528).  Otherwise m' = m.
 *   residual   optional [M', N] tensor added after the activation; may alias `out`.
 */
typedef struct tcavp_gemm_args {
  int M, N, K;
  const void* A; int lda;
  const void* W; int ldw;
  int in_dtype;
  void* out; int ldo; int out_dtype;
  const float* bias;
  const void* residual; int ldr; int res_dtype;
  int act;
  float scale;               /* 0 is treated as 1 */
  const float* lora_t; int lora_ldt; int lora_r; const float* lora_b;
  int seg_begin[2], seg_end[2], seg_toff[2];
  int remap_gi, remap_go, remap_off;
} tcavp_gemm_args;

int tcavp_gemm(const tcavp_gemm_args* args, tcavp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TCAVP_H_ */
