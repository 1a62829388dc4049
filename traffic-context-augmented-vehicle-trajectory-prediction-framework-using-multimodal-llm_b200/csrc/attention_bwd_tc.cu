// Tensor-core attention backward (bf16, head_dim 16/32/64/96/128, Tq, Tk <= 256, H == Hkv): the gradient of the LLM
// self-attention (HF:251-289) and of the small nn.MultiheadAttention cores, used by the fine-tune step.
//
// One CTA per (batch, head).  Q, K, V and dO of the head are staged ONCE in shared memory (cp.async, padded rows); no score,
// probability or gradient tile ever touches shared or global memory, and there are no atomics:
//   pass B (query-outer, a warp owns 16-row query slabs, paired from both ends for causal balance)
//       sweep 1: S = Q K^T -> row max m_i and 1 / row sum (online softmax statistics), D_i = dO_i . O_i
//       sweep 2: S again -> P = exp2(S - m_i) / l_i;  dP = dO V^T;  dS = scale * P o (dP - D_i);  dQ += dS K   (registers)
//   pass A (key-outer, a warp owns 16-key slabs): for every query slab
//       S^T = K_J Q_I^T, dP^T = V_J dO_I^T  (computed transposed so that P^T / dS^T come out in the A-operand layout)
//       dV_J += P^T dO_I;  dK_J += dS^T Q_I                                                                 (registers)
// mma.sync m16n8k16 bf16 -> fp32.  The scores are recomputed three times instead of being stored: the contraction work is
// tiny next to the GEMMs of the step, while stored probabilities would cost B*H*L*L*2 bytes of HBM per layer.
#include "common.cuh"

namespace tcavp {
namespace fb {

constexpr int PAD = 8;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// KB = keys per block of the query-outer pass (48 when that pads Tk to fewer dead keys, e.g. L = 144 = 3 x 48; 64 otherwise).
template <int DH, int MAXT, int MINB, int KB>
__global__ void __launch_bounds__(MAXT, MINB) attn_bwd_tc_kernel(tcavp_attn_args a, const __nv_bfloat16* __restrict__ dout, long long do_sb,
                                                                 long long do_st, __nv_bfloat16* __restrict__ dq, long long dq_sb,
                                                                 long long dq_st, void* __restrict__ dk, long long dk_sb, long long dk_st,
                                                                 void* __restrict__ dv, long long dv_sb, long long dv_st, int dkv_bf16,
                                                                 int tq_pad, int tk_pad) {
  constexpr int LDS = DH + PAD;
  constexpr int KS = DH / 16;
  constexpr int NT = DH / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sdO = sQ + (size_t)tq_pad * LDS;
  __nv_bfloat16* sK = sdO + (size_t)tq_pad * LDS;
  __nv_bfloat16* sV = sK + (size_t)tk_pad * LDS;
  float* sM = reinterpret_cast<float*>(sV + (size_t)tk_pad * LDS);   // [tq_pad] row max of the scaled base-2 logits (0 if none)
  float* sIL = sM + tq_pad;                                           // [tq_pad] 1 / row sum (0 for empty / padded rows)
  float* sD = sIL + tq_pad;                                           // [tq_pad] dO_i . O_i
  int* sMask = reinterpret_cast<int*>(sD + tq_pad);                   // [tk_pad]

  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const __nv_bfloat16* gq = reinterpret_cast<const __nv_bfloat16*>(a.q) + (size_t)b * a.q_sb + (size_t)h * DH;
  const __nv_bfloat16* gk = reinterpret_cast<const __nv_bfloat16*>(a.k) + (size_t)b * a.k_sb + (size_t)h * DH;
  const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(a.v) + (size_t)b * a.v_sb + (size_t)h * DH;
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(a.out) + (size_t)b * a.o_sb + (size_t)h * DH;
  const __nv_bfloat16* gdo = dout + (size_t)b * do_sb + (size_t)h * DH;

  // ---- stage Q, dO (tq_pad rows) and K, V (tk_pad rows), zero-filled beyond the real extents ----
  {
    constexpr int VPR = DH / 8;
    const uint32_t sQ_a = (uint32_t)__cvta_generic_to_shared(sQ), sdO_a = (uint32_t)__cvta_generic_to_shared(sdO);
    const uint32_t sK_a = (uint32_t)__cvta_generic_to_shared(sK), sV_a = (uint32_t)__cvta_generic_to_shared(sV);
    for (int i = threadIdx.x; i < tq_pad * VPR; i += blockDim.x) {
      const int r = i / VPR, c = (i % VPR) * 8;
      const int sz = r < a.Tq ? 16 : 0;
      const size_t rr = r < a.Tq ? r : 0;
      const uint32_t off = (uint32_t)((r * LDS + c) * 2);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sQ_a + off), "l"(gq + rr * a.q_st + c), "r"(sz) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdO_a + off), "l"(gdo + rr * do_st + c), "r"(sz) : "memory");
    }
    for (int i = threadIdx.x; i < tk_pad * VPR; i += blockDim.x) {
      const int r = i / VPR, c = (i % VPR) * 8;
      const int sz = r < a.Tk ? 16 : 0;
      const size_t rr = r < a.Tk ? r : 0;
      const uint32_t off = (uint32_t)((r * LDS + c) * 2);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sK_a + off), "l"(gk + rr * a.k_st + c), "r"(sz) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sV_a + off), "l"(gv + rr * a.v_st + c), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int j = threadIdx.x; j < tk_pad; j += blockDim.x) sMask[j] = (j < a.Tk) && (!a.key_mask || a.key_mask[(size_t)b * a.Tk + j] != 0);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ), sdO_u = (uint32_t)__cvta_generic_to_shared(sdO);
  const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK), sV_u = (uint32_t)__cvta_generic_to_shared(sV);
  // ldmatrix lane offsets: A operand (16 rows x 16 k, row-major), B operand from [n][k] storage, B operand from [k][n] storage
  const int a_row = lane & 15, a_col = (lane >> 4) * 8;
  const int bn_row = (lane & 7) + ((lane >> 4) << 3), bn_col = ((lane >> 3) & 1) * 8;
  const int bk_row = (lane & 7) + (((lane >> 3) & 1) << 3), bk_col = (lane >> 4) * 8;
  const float sl2 = a.scale * 1.4426950408889634f;
  // train-mode dropout (O = P_d V, P_d = keep ? P / (1 - p_drop) : 0): dP = f o (dO V^T) with f = keep ? 1 / (1 - p_drop) : 0,
  // D_i = dO_i . O_i is unchanged, dV uses P_d.  The mask is regenerated from (seed, site, ((b*H + h)*Tq + i)*Tk + j).
  const bool drop = a.drop_thresh != 0;
  const uint32_t dkey = drop ? drop_key(a.drop_seed, a.drop_site) : 0u;
  const unsigned long long dbase = (unsigned long long)blockIdx.x * a.Tq * (unsigned long long)a.Tk;

  // =========================== pass B: statistics, D and dQ (query-outer) ===========================
  const int nqs = (a.Tq + 15) / 16;
  for (int idx = warp; idx < (nqs + 1) / 2; idx += nwarps) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int slab = pass == 0 ? nqs - 1 - idx : idx;
      if (pass == 1 && slab == nqs - 1 - idx) continue;
      const int row0 = slab * 16;
      const int r_lo = row0 + g, r_hi = row0 + g + 8;
      uint32_t qf[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(sQ_u + (uint32_t)(((row0 + a_row) * LDS + ks * 16 + a_col) * 2), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      const int k_end = a.causal ? min(a.Tk, row0 + 16) : a.Tk;
      // ---- sweep 1: softmax statistics ----
      float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
      for (int kb = 0; kb < k_end; kb += KB) {
        float s[KB / 8][4];
#pragma unroll
        for (int n = 0; n < KB / 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int np = 0; np < KB / 16; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sK_u + (uint32_t)(((kb + np * 16 + bn_row) * LDS + ks * 16 + bn_col) * 2), b0, b1, b2, b3);
            mma16816(s[2 * np], qf[ks], b0, b1);
            mma16816(s[2 * np + 1], qf[ks], b2, b3);
          }
        }
        float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
        for (int n = 0; n < KB / 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = kb + n * 8 + t4 * 2 + e;
            const bool okj = sMask[j] != 0;
            s[n][e] = (okj && (!a.causal || j <= r_lo)) ? s[n][e] * sl2 : -INFINITY;
            s[n][2 + e] = (okj && (!a.causal || j <= r_hi)) ? s[n][2 + e] * sl2 : -INFINITY;
          }
          mx_lo = fmaxf(mx_lo, fmaxf(s[n][0], s[n][1]));
          mx_hi = fmaxf(mx_hi, fmaxf(s[n][2], s[n][3]));
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float ref_lo = mx_lo == -INFINITY ? 0.f : mx_lo, ref_hi = mx_hi == -INFINITY ? 0.f : mx_hi;
        l_lo *= ex2(m_lo - ref_lo);
        l_hi *= ex2(m_hi - ref_hi);
        m_lo = mx_lo;
        m_hi = mx_hi;
#pragma unroll
        for (int n = 0; n < KB / 8; ++n) {
          l_lo += ex2(s[n][0] - ref_lo) + ex2(s[n][1] - ref_lo);
          l_hi += ex2(s[n][2] - ref_hi) + ex2(s[n][3] - ref_hi);
        }
      }
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
      const float rm_lo = m_lo == -INFINITY ? 0.f : m_lo, rm_hi = m_hi == -INFINITY ? 0.f : m_hi;
      const float il_lo = (l_lo > 0.f && r_lo < a.Tq) ? 1.f / l_lo : 0.f, il_hi = (l_hi > 0.f && r_hi < a.Tq) ? 1.f / l_hi : 0.f;
      // ---- D_i = dO_i . O_i ----
      float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int c = n * 8 + t4 * 2;
        if (r_lo < a.Tq) {
          const float2 o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(go + (size_t)r_lo * a.o_st + c));
          const float2 d2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sdO + r_lo * LDS + c));
          d_lo += o2.x * d2.x + o2.y * d2.y;
        }
        if (r_hi < a.Tq) {
          const float2 o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(go + (size_t)r_hi * a.o_st + c));
          const float2 d2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sdO + r_hi * LDS + c));
          d_hi += o2.x * d2.x + o2.y * d2.y;
        }
      }
      d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1);
      d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
      d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1);
      d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
      if (t4 == 0) {
        sM[r_lo] = rm_lo; sIL[r_lo] = il_lo; sD[r_lo] = d_lo;
        sM[r_hi] = rm_hi; sIL[r_hi] = il_hi; sD[r_hi] = d_hi;
      }
      // ---- sweep 2: dQ ----
      float acc[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      for (int kb = 0; kb < k_end; kb += KB) {
        float s[KB / 8][4], dp[KB / 8][4];
#pragma unroll
        for (int n = 0; n < KB / 8; ++n) {
          s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t df[4];
          ldsm_x4(sdO_u + (uint32_t)(((row0 + a_row) * LDS + ks * 16 + a_col) * 2), df[0], df[1], df[2], df[3]);
#pragma unroll
          for (int np = 0; np < KB / 16; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sK_u + (uint32_t)(((kb + np * 16 + bn_row) * LDS + ks * 16 + bn_col) * 2), b0, b1, b2, b3);
            mma16816(s[2 * np], qf[ks], b0, b1);
            mma16816(s[2 * np + 1], qf[ks], b2, b3);
            ldsm_x4(sV_u + (uint32_t)(((kb + np * 16 + bn_row) * LDS + ks * 16 + bn_col) * 2), b0, b1, b2, b3);
            mma16816(dp[2 * np], df, b0, b1);
            mma16816(dp[2 * np + 1], df, b2, b3);
          }
        }
        uint32_t pf[KB / 16][4];
#pragma unroll
        for (int n = 0; n < KB / 8; ++n) {
          float ds[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = kb + n * 8 + t4 * 2 + e;
            const bool okj = sMask[j] != 0;
            const float p_lo = (okj && (!a.causal || j <= r_lo)) ? ex2(s[n][e] * sl2 - rm_lo) * il_lo : 0.f;
            const float p_hi = (okj && (!a.causal || j <= r_hi)) ? ex2(s[n][2 + e] * sl2 - rm_hi) * il_hi : 0.f;
            float g_lo = dp[n][e], g_hi = dp[n][2 + e];
            if (drop) {
              g_lo = drop_keep(dkey, dbase + (unsigned long long)r_lo * a.Tk + (unsigned)j, a.drop_thresh) ? g_lo * a.drop_scale : 0.f;
              g_hi = drop_keep(dkey, dbase + (unsigned long long)r_hi * a.Tk + (unsigned)j, a.drop_thresh) ? g_hi * a.drop_scale : 0.f;
            }
            ds[e] = a.scale * p_lo * (g_lo - d_lo);
            ds[2 + e] = a.scale * p_hi * (g_hi - d_hi);
          }
          pf[n >> 1][(n & 1) * 2] = pack2(ds[0], ds[1]);
          pf[n >> 1][(n & 1) * 2 + 1] = pack2(ds[2], ds[3]);
        }
#pragma unroll
        for (int kk = 0; kk < KB / 16; ++kk) {
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(sK_u + (uint32_t)(((kb + kk * 16 + bk_row) * LDS + np * 16 + bk_col) * 2), b0, b1, b2, b3);
            mma16816(acc[2 * np], pf[kk], b0, b1);
            mma16816(acc[2 * np + 1], pf[kk], b2, b3);
          }
        }
      }
      __nv_bfloat16* gdq = dq + (size_t)b * dq_sb + (size_t)h * DH;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int c = n * 8 + t4 * 2;
        if (r_lo < a.Tq) *reinterpret_cast<uint32_t*>(gdq + (size_t)r_lo * dq_st + c) = pack2(acc[n][0], acc[n][1]);
        if (r_hi < a.Tq) *reinterpret_cast<uint32_t*>(gdq + (size_t)r_hi * dq_st + c) = pack2(acc[n][2], acc[n][3]);
      }
    }
  }
  __syncthreads();

  // =========================== pass A: dK, dV (key-outer) ===========================
  const int nks = (a.Tk + 15) / 16;
  for (int idx = warp; idx < (nks + 1) / 2; idx += nwarps) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int slab = pass == 0 ? idx : nks - 1 - idx;      // causal: key slab 0 sees every query slab, the last one only one
      if (pass == 1 && slab == idx) continue;
      const int j0 = slab * 16;
      const int j_lo = j0 + g, j_hi = j0 + g + 8;
      const bool ok_lo = sMask[j_lo] != 0, ok_hi = sMask[j_hi] != 0;
      float ak[NT][4], av[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        ak[n][0] = ak[n][1] = ak[n][2] = ak[n][3] = 0.f;
        av[n][0] = av[n][1] = av[n][2] = av[n][3] = 0.f;
      }
      const int i_begin = a.causal ? j0 : 0;                 // queries i < j0 see none of this slab's keys
      for (int i0 = i_begin; i0 < tq_pad; i0 += 16) {
        float st[2][4], dpt[2][4];                           // S^T and dP^T: rows = keys (lo / hi), cols = queries i0 + n*8 + 2*t4 + e
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
          dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t kf[4], vf[4], b0, b1, b2, b3;
          ldsm_x4(sK_u + (uint32_t)(((j0 + a_row) * LDS + ks * 16 + a_col) * 2), kf[0], kf[1], kf[2], kf[3]);
          ldsm_x4(sV_u + (uint32_t)(((j0 + a_row) * LDS + ks * 16 + a_col) * 2), vf[0], vf[1], vf[2], vf[3]);
          ldsm_x4(sQ_u + (uint32_t)(((i0 + bn_row) * LDS + ks * 16 + bn_col) * 2), b0, b1, b2, b3);
          mma16816(st[0], kf, b0, b1);
          mma16816(st[1], kf, b2, b3);
          ldsm_x4(sdO_u + (uint32_t)(((i0 + bn_row) * LDS + ks * 16 + bn_col) * 2), b0, b1, b2, b3);
          mma16816(dpt[0], vf, b0, b1);
          mma16816(dpt[1], vf, b2, b3);
        }
        uint32_t pT[4], dsT[4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = i0 + n * 8 + t4 * 2 + e;
            const float m = sM[i], il = sIL[i], D = sD[i];
            p[e] = (ok_lo && (!a.causal || j_lo <= i)) ? ex2(st[n][e] * sl2 - m) * il : 0.f;
            p[2 + e] = (ok_hi && (!a.causal || j_hi <= i)) ? ex2(st[n][2 + e] * sl2 - m) * il : 0.f;
            float g_lo = dpt[n][e], g_hi = dpt[n][2 + e];
            float f_lo = 1.f, f_hi = 1.f;
            if (drop) {
              const unsigned long long irow = dbase + (unsigned long long)i * a.Tk;
              f_lo = drop_keep(dkey, irow + (unsigned)j_lo, a.drop_thresh) ? a.drop_scale : 0.f;
              f_hi = drop_keep(dkey, irow + (unsigned)j_hi, a.drop_thresh) ? a.drop_scale : 0.f;
              g_lo *= f_lo;
              g_hi *= f_hi;
            }
            ds[e] = a.scale * p[e] * (g_lo - D);
            ds[2 + e] = a.scale * p[2 + e] * (g_hi - D);
            p[e] *= f_lo;            // dV_J += P_d^T dO
            p[2 + e] *= f_hi;
          }
          pT[n * 2] = pack2(p[0], p[1]);
          pT[n * 2 + 1] = pack2(p[2], p[3]);
          dsT[n * 2] = pack2(ds[0], ds[1]);
          dsT[n * 2 + 1] = pack2(ds[2], ds[3]);
        }
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sdO_u + (uint32_t)(((i0 + bk_row) * LDS + np * 16 + bk_col) * 2), b0, b1, b2, b3);
          mma16816(av[2 * np], pT, b0, b1);
          mma16816(av[2 * np + 1], pT, b2, b3);
          ldsm_x4_t(sQ_u + (uint32_t)(((i0 + bk_row) * LDS + np * 16 + bk_col) * 2), b0, b1, b2, b3);
          mma16816(ak[2 * np], dsT, b0, b1);
          mma16816(ak[2 * np + 1], dsT, b2, b3);
        }
      }
      if (dkv_bf16) {     // this CTA owns every key row of its head: the activation dtype is written directly (no fp32 round trip)
        __nv_bfloat16* gdk = reinterpret_cast<__nv_bfloat16*>(dk) + (size_t)b * dk_sb + (size_t)h * DH;
        __nv_bfloat16* gdv = reinterpret_cast<__nv_bfloat16*>(dv) + (size_t)b * dv_sb + (size_t)h * DH;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int c = n * 8 + t4 * 2;
          if (j_lo < a.Tk) {
            *reinterpret_cast<uint32_t*>(gdk + (size_t)j_lo * dk_st + c) = pack2(ak[n][0], ak[n][1]);
            *reinterpret_cast<uint32_t*>(gdv + (size_t)j_lo * dv_st + c) = pack2(av[n][0], av[n][1]);
          }
          if (j_hi < a.Tk) {
            *reinterpret_cast<uint32_t*>(gdk + (size_t)j_hi * dk_st + c) = pack2(ak[n][2], ak[n][3]);
            *reinterpret_cast<uint32_t*>(gdv + (size_t)j_hi * dv_st + c) = pack2(av[n][2], av[n][3]);
          }
        }
      } else {
        float* gdk = reinterpret_cast<float*>(dk) + (size_t)b * dk_sb + (size_t)h * DH;
        float* gdv = reinterpret_cast<float*>(dv) + (size_t)b * dv_sb + (size_t)h * DH;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int c = n * 8 + t4 * 2;
          if (j_lo < a.Tk) {
            *reinterpret_cast<float2*>(gdk + (size_t)j_lo * dk_st + c) = make_float2(ak[n][0], ak[n][1]);
            *reinterpret_cast<float2*>(gdv + (size_t)j_lo * dv_st + c) = make_float2(av[n][0], av[n][1]);
          }
          if (j_hi < a.Tk) {
            *reinterpret_cast<float2*>(gdk + (size_t)j_hi * dk_st + c) = make_float2(ak[n][2], ak[n][3]);
            *reinterpret_cast<float2*>(gdv + (size_t)j_hi * dv_st + c) = make_float2(av[n][2], av[n][3]);
          }
        }
      }
    }
  }
}

}  // namespace fb

// Returns 1 when the shape is not covered (caller uses the generic SIMT kernel), <= 0 otherwise.
int attention_bwd_tc_launch(const tcavp_attn_args& a, const void* dout, long long do_sb, long long do_st, void* dq, long long dq_sb,
                            long long dq_st, void* dk, long long dk_sb, long long dk_st, void* dv, long long dv_sb, long long dv_st,
                            int dkv_dtype, cudaStream_t stream) {
  if (a.dtype != TCAVP_BF16 || a.out == nullptr || a.H != a.Hkv) return 1;
  if (!(a.dh == 16 || a.dh == 32 || a.dh == 64 || a.dh == 96 || a.dh == 128) || a.Tk > 256 || a.Tq > 256) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(dout) || reinterpret_cast<uintptr_t>(a.out) % 4 || reinterpret_cast<uintptr_t>(dq) % 4 ||
      reinterpret_cast<uintptr_t>(dk) % 8 || reinterpret_cast<uintptr_t>(dv) % 8)
    return 1;
  if (a.q_sb % 8 || a.q_st % 8 || a.k_sb % 8 || a.k_st % 8 || a.v_sb % 8 || a.v_st % 8 || do_sb % 8 || do_st % 8 || a.o_sb % 2 || a.o_st % 2 ||
      dq_sb % 2 || dq_st % 2 || dk_sb % 2 || dk_st % 2 || dv_sb % 2 || dv_st % 2)
    return 1;
  const int tq_pad = (a.Tq + 15) / 16 * 16;
  auto pad_to = [](int n, int m) { return (n + m - 1) / m * m; };
  const int kb = pad_to(a.Tk, 48) < pad_to(a.Tk, 64) ? 48 : 64;
  const int tk_pad = pad_to(a.Tk, kb);
  const int nqs = (a.Tq + 15) / 16, nks = (a.Tk + 15) / 16;
  const int slabs = nqs > nks ? nqs : nks;
  int warps = (slabs + 1) / 2;
  if (warps > 8) warps = 8;
  const size_t smem = (size_t)2 * (tq_pad + tk_pad) * (a.dh + fb::PAD) * 2 + (size_t)3 * tq_pad * 4 + (size_t)tk_pad * 4;
  if (smem > 220 * 1024) return 1;
  const dim3 grid(a.B * a.H);
#define TCAVP_BWD_KB(DH, MAXT, MINB, KB_)                                                                                                   \
  do {                                                                                                                                      \
    TCAVP_CUDA(cudaFuncSetAttribute(fb::attn_bwd_tc_kernel<DH, MAXT, MINB, KB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    fb::attn_bwd_tc_kernel<DH, MAXT, MINB, KB_><<<grid, warps * 32, smem, stream>>>(                                                        \
        a, reinterpret_cast<const __nv_bfloat16*>(dout), do_sb, do_st, reinterpret_cast<__nv_bfloat16*>(dq), dq_sb, dq_st, dk, dk_sb, dk_st, \
        dv, dv_sb, dv_st, dkv_dtype == TCAVP_BF16 ? 1 : 0, tq_pad, tk_pad);                                                                  \
  } while (0)
#define TCAVP_BWD(DH, MAXT, MINB)                      \
  do {                                                 \
    if (kb == 48) TCAVP_BWD_KB(DH, MAXT, MINB, 48);    \
    else TCAVP_BWD_KB(DH, MAXT, MINB, 64);             \
  } while (0)
  if (a.dh == 64) {
    if (warps <= 5) TCAVP_BWD(64, 160, 2);
    else TCAVP_BWD(64, 256, 1);
  } else if (a.dh == 128) {
    TCAVP_BWD(128, 256, 1);
  } else if (a.dh == 96) {
    TCAVP_BWD(96, 256, 1);
  } else if (a.dh == 32) {
    TCAVP_BWD(32, 256, 2);
  } else {
    TCAVP_BWD(16, 256, 2);
  }
#undef TCAVP_BWD
#undef TCAVP_BWD_KB
  return check_launch("attn_bwd_tc_kernel");
}

}  // namespace tcavp
