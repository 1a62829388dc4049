// Token cross-entropy of the stage-1 (CausalLM) objective: HF ForCausalLMLoss behind LlamaForCausalLM.forward(labels=...) — logits
// upcast to fp32, targets shifted by one, ignore_index -100, mean over the labelled positions (HF:487-491; reference
// scripts/check_generation.py:131-151, scripts/train.py:533-547).  The caller keeps only the labelled rows, projects them through
// lm_head in row chunks (tcavp_gemm, fp32 logits) and hands each chunk to this kernel, which does forward AND backward of the loss:
//     loss_sum += logsumexp(row) - row[target]          grad[row, :] = (softmax(row) - onehot(target)) * scale
// One CTA per row, three sweeps over the row (max, sum of exponentials, gradient); the row (<= 0.5 MB at a 128k vocabulary) stays in L2
// between the sweeps, so HBM sees the logits once and the gradient once.
#include "common.cuh"

namespace tcavp {

__device__ __forceinline__ float block_reduce(float v, float* scratch, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = scratch[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = is_max ? fmaxf(t, scratch[w]) : t + scratch[w];
  return t;
}

__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ targets,
                                                      float* __restrict__ loss_sum, void* __restrict__ grad, long long ldg, int grad_dtype, int V,
                                                      float scale) {
  __shared__ float scratch[8];
  const float* row = logits + (size_t)blockIdx.x * ld;
  const long long tgt = targets[blockIdx.x];
  const bool vec = (V % 4 == 0) && (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) % 16 == 0);
  float m = -INFINITY;
  if (vec) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int i = threadIdx.x; i < V / 4; i += blockDim.x) {
      const float4 v = r4[i];
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, row[i]);
  }
  m = block_reduce(m, scratch, true);
  float s = 0.f;
  if (vec) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int i = threadIdx.x; i < V / 4; i += blockDim.x) {
      const float4 v = r4[i];
      s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    }
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) s += __expf(row[i] - m);
  }
  s = block_reduce(s, scratch, false);
  const float lse = m + logf(s);
  if (threadIdx.x == 0 && tgt >= 0 && tgt < V) atomicAdd(loss_sum, lse - row[tgt]);
  if (!grad) return;
  if (grad_dtype == TCAVP_BF16) {
    __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(grad) + (size_t)blockIdx.x * ldg;
    if (vec && ldg % 4 == 0 && reinterpret_cast<uintptr_t>(grad) % 8 == 0) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      for (int i = threadIdx.x; i < V / 4; i += blockDim.x) {
        const float4 v = r4[i];
        const int j = i * 4;
        const __nv_bfloat162 a = __floats2bfloat162_rn((__expf(v.x - lse) - (j == tgt ? 1.f : 0.f)) * scale, (__expf(v.y - lse) - (j + 1 == tgt ? 1.f : 0.f)) * scale);
        const __nv_bfloat162 b = __floats2bfloat162_rn((__expf(v.z - lse) - (j + 2 == tgt ? 1.f : 0.f)) * scale, (__expf(v.w - lse) - (j + 3 == tgt ? 1.f : 0.f)) * scale);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&a);
        u.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(g + j) = u;
      }
    } else {
      for (int i = threadIdx.x; i < V; i += blockDim.x) g[i] = __float2bfloat16_rn((__expf(row[i] - lse) - (i == tgt ? 1.f : 0.f)) * scale);
    }
  } else {
    float* g = reinterpret_cast<float*>(grad) + (size_t)blockIdx.x * ldg;
    for (int i = threadIdx.x; i < V; i += blockDim.x) g[i] = (expf(row[i] - lse) - (i == tgt ? 1.f : 0.f)) * scale;
  }
}

}  // namespace tcavp

extern "C" int tcavp_ce_loss(const float* logits, long long ld, const long long* targets, float* loss_sum, void* grad, long long ldg, int grad_dtype,
                             long long rows, int V, float scale, tcavp_stream_t stream) {
  using namespace tcavp;
  TCAVP_REQUIRE(rows >= 0 && V > 0 && ld >= V && (!grad || ldg >= V), "tcavp_ce_loss: bad shape");
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(logits && targets && loss_sum && rows < (1ll << 31), "tcavp_ce_loss: null pointer / too many rows");
  TCAVP_REQUIRE(!grad || grad_dtype == TCAVP_F32 || grad_dtype == TCAVP_BF16, "tcavp_ce_loss: bad grad dtype");
  ce_loss_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, ld, targets, loss_sum, grad, ldg, grad_dtype, V, scale);
  return check_launch("ce_loss_kernel");
}
