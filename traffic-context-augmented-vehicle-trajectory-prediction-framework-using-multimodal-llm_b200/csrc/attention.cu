// tcavp_attention: softmax(scale * q.k^T + mask) . v for strided (b, t, h, d) views.
//
//   generic path : one warp per (batch, head, query row); lanes stride over the head dimension, keys are
//                  processed four at a time (four interleaved warp-shuffle reductions), online softmax in
//                  fp32.  Handles every head_dim on the path (16 ... 2048), causal and key-padding masks.
//   flash path   : bf16, head_dim 64/128 (the LLM self-attention, HF:251-289) — see attention_tc.cu.
//   wide path    : bf16, few queries against a wide head (LTSF cross-attention) — see attention_x.cu.
#include "common.cuh"

namespace tcavp {

int attention_tc_launch(const tcavp_attn_args& a, cudaStream_t stream);   // attention_tc.cu; returns 1 if not applicable
int attention_x_launch(const tcavp_attn_args& a, cudaStream_t stream);    // attention_x.cu;  returns 1 if not applicable

template <typename T, int NE>   // NE = ceil(dh / 32) head-dim elements per lane
__global__ void __launch_bounds__(128) attn_warp_kernel(tcavp_attn_args a) {
  const int lane = threadIdx.x & 31;
  const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)a.B * a.H * a.Tq;
  if (w >= total) return;
  const int i = (int)(w % a.Tq);
  const int h = (int)((w / a.Tq) % a.H);
  const int b = (int)(w / ((long long)a.Tq * a.H));
  const int hk = h / (a.H / a.Hkv);
  const T* q = reinterpret_cast<const T*>(a.q) + (size_t)b * a.q_sb + (size_t)i * a.q_st + (size_t)h * a.dh;
  const T* k = reinterpret_cast<const T*>(a.k) + (size_t)b * a.k_sb + (size_t)hk * a.dh;
  const T* v = reinterpret_cast<const T*>(a.v) + (size_t)b * a.v_sb + (size_t)hk * a.dh;
  const int32_t* km = a.key_mask ? a.key_mask + (size_t)b * a.Tk : nullptr;

  float qr[NE], acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int d = lane + 32 * e;
    qr[e] = d < a.dh ? Cvt<T>::to_f(q[d]) * a.scale : 0.f;
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int kend = a.causal ? i + 1 : a.Tk;
  for (int j0 = 0; j0 < kend; j0 += 4) {
    float s[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      ok[u] = j < kend && (!km || km[j] != 0);
      float p = 0.f;
      if (ok[u]) {
        const T* kj = k + (size_t)j * a.k_st;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int d = lane + 32 * e;
          if (d < a.dh) p = fmaf(qr[e], Cvt<T>::to_f(kj[d]), p);
        }
      }
      s[u] = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
    float mn = m;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) mn = fmaxf(mn, s[u]);
    if (mn == -INFINITY) continue;   // nothing attendable so far (warp-uniform)
    const float corr = __expf(m - mn);   // m == -inf -> 0
    float p[4];
    float ps = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      p[u] = ok[u] ? __expf(s[u] - mn) : 0.f;
      ps += p[u];
    }
    l = l * corr + ps;
#pragma unroll
    for (int e = 0; e < NE; ++e) acc[e] *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (ok[u]) {
        const T* vj = v + (size_t)(j0 + u) * a.v_st;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int d = lane + 32 * e;
          if (d < a.dh) acc[e] = fmaf(p[u], Cvt<T>::to_f(vj[d]), acc[e]);
        }
      }
    }
    m = mn;
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  T* o = reinterpret_cast<T*>(a.out) + (size_t)b * a.o_sb + (size_t)i * a.o_st + (size_t)h * a.dh;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int d = lane + 32 * e;
    if (d < a.dh) o[d] = Cvt<T>::from_f(acc[e] * inv);
  }
}

template <typename T>
static int launch_warp(const tcavp_attn_args& a, cudaStream_t stream) {
  const long long total = (long long)a.B * a.H * a.Tq;
  const int grid = (int)((total + 3) / 4);
  const int ne = (a.dh + 31) / 32;
#define CASE(N)                                              \
  if (ne <= N) {                                             \
    attn_warp_kernel<T, N><<<grid, 128, 0, stream>>>(a);     \
    return check_launch("attn_warp_kernel");                 \
  }
  CASE(1) CASE(2) CASE(3) CASE(4) CASE(8) CASE(12) CASE(16) CASE(32) CASE(64)
#undef CASE
  return fail_arg("tcavp_attention: head_dim %d > 2048 unsupported", a.dh);
}

}  // namespace tcavp

extern "C" int tcavp_attention(const tcavp_attn_args* a, tcavp_stream_t stream_) {
  using namespace tcavp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TCAVP_REQUIRE(a != nullptr, "tcavp_attention: null args");
  TCAVP_REQUIRE(a->B >= 0 && a->H > 0 && a->Hkv > 0 && a->H % a->Hkv == 0 && a->Tq >= 0 && a->Tk >= 0 && a->dh > 0,
                "tcavp_attention: bad shape B=%d H=%d Hkv=%d Tq=%d Tk=%d dh=%d", a->B, a->H, a->Hkv, a->Tq, a->Tk, a->dh);
  if (a->B == 0 || a->Tq == 0) return TCAVP_OK;
  TCAVP_REQUIRE(a->q && a->k && a->v && a->out, "tcavp_attention: null tensor");
  TCAVP_REQUIRE(!a->causal || a->Tq == a->Tk, "tcavp_attention: causal needs Tq == Tk");
  TCAVP_REQUIRE(a->dtype == TCAVP_F32 || a->dtype == TCAVP_BF16, "tcavp_attention: bad dtype");
  if (a->dtype == TCAVP_BF16) {
    int rc = a->dh > 128 ? attention_x_launch(*a, stream) : attention_tc_launch(*a, stream);
    if (rc <= 0) return rc;
    return launch_warp<__nv_bfloat16>(*a, stream);
  }
  return launch_warp<float>(*a, stream);
}
