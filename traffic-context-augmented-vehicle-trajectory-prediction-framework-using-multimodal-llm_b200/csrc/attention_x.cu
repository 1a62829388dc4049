// Tensor-core attention for FEW queries against a wide head: the LTSF cross-attention
// (reference scripts/train.py:793-798: T_out <= 64 queries, L keys, 2 heads of width H/2 = 384 ... 2048).
//
// The generic warp-per-query kernel re-reads all of K and V for every query row; here one CTA owns a
// (batch, head) pair and streams K, then V, through shared memory exactly once in 64-wide head-dim chunks
// (cp.async double buffering), so the kernel is bound by the single HBM pass over K and V:
//   phase 1   S[QROWS, Tk] = sum over chunks  Q[:, chunk] . K[:, chunk]^T     (mma.sync m16n8k16, fp32 accum, registers)
//   softmax   in registers (+ a QROWS x KG exchange of partial maxima / sums), unnormalised bf16 probabilities in shared memory
//   phase 2   O[:, chunk] = P . V[:, chunk] for every chunk, written straight to global memory
#include "common.cuh"

namespace tcavp {
namespace xa {

constexpr int CH = 64;           // head-dim chunk
constexpr int LDC = CH + 8;      // padded chunk row (elements): conflict-free ldmatrix
constexpr int THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// QROWS = padded query rows per CTA (32: T_out <= 32, 64: T_out <= 64, e.g. the 50-step horizon); MAXP = 16-key pairs per
// warp in phase 1.  Eight warps = (QROWS/16 m16 tiles) x (KG key groups).
template <int QROWS, int MAXP>
__global__ void __launch_bounds__(THREADS, 2) attn_x_kernel(tcavp_attn_args a, int tkp) {
  constexpr int MT = QROWS / 16;          // m16 tiles
  constexpr int KG = 8 / MT;              // key groups (phase 1) = 16-dim column groups (phase 2)
  constexpr int NPW = (CH / 16) / KG;     // 16-dim pairs of a chunk per warp in phase 2
  extern __shared__ __align__(16) uint8_t smem[];
  // stage[2] : (QROWS + tkp) x LDC bf16 (phase 1: Q rows then K rows; phase 2: V rows) | P : QROWS x (tkp+8) bf16 (unnormalised)
  // | red : QROWS x KG fp32 partial row maxima, then partial row sums
  const int stage_elems = (QROWS + tkp) * LDC;
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sP = stage + 2 * (size_t)stage_elems;
  const int ldp = tkp + 8;
  float* sRed = reinterpret_cast<float*>(sP + (size_t)QROWS * ldp);
  const uint32_t stage_u = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t sP_u = (uint32_t)__cvta_generic_to_shared(sP);

  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int hk = h / (a.H / a.Hkv);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* gq = reinterpret_cast<const __nv_bfloat16*>(a.q) + (size_t)b * a.q_sb + (size_t)h * a.dh;
  const __nv_bfloat16* gk = reinterpret_cast<const __nv_bfloat16*>(a.k) + (size_t)b * a.k_sb + (size_t)hk * a.dh;
  const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(a.v) + (size_t)b * a.v_sb + (size_t)hk * a.dh;
  const int nchunks = a.dh / CH;

  auto load_qk = [&](int c, int buf) {
    const uint32_t base = stage_u + (uint32_t)buf * stage_elems * 2;
    for (int i = tid; i < (QROWS + tkp) * (CH / 8); i += THREADS) {
      const int r = i >> 3, v = (i & 7) * 8;
      const bool isq = r < QROWS;
      const int rr = isq ? r : r - QROWS;
      const bool valid = isq ? rr < a.Tq : rr < a.Tk;
      const __nv_bfloat16* src = isq ? gq + (size_t)rr * a.q_st + c * CH + v : gk + (size_t)rr * a.k_st + c * CH + v;
      cp_async16(base + (uint32_t)(r * LDC + v) * 2, valid ? src : gq, valid);
    }
  };
  auto load_v = [&](int c, int buf) {
    const uint32_t base = stage_u + (uint32_t)buf * stage_elems * 2;
    for (int i = tid; i < tkp * (CH / 8); i += THREADS) {
      const int r = i >> 3, v = (i & 7) * 8;
      const bool valid = r < a.Tk;
      cp_async16(base + (uint32_t)(r * LDC + v) * 2, valid ? gv + (size_t)r * a.v_st + c * CH + v : gv, valid);
    }
  };

  // ---------------- phase 1: S = Q . K^T ----------------
  const int mt = warp % MT, kg = warp / MT;                  // m16 tile, key group
  const int npairs_all = tkp / 16;
  const int base_p = npairs_all / KG, rem_p = npairs_all % KG;
  const int my_pairs = base_p + (kg < rem_p ? 1 : 0);
  const int pair0 = kg * base_p + min(kg, rem_p);
  const bool m_active = mt * 16 < a.Tq;
  float acc[MAXP * 2][4];
#pragma unroll
  for (int i = 0; i < MAXP * 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;   // A operand (row-major)
  const int k_row = (lane & 7) + ((lane >> 4) << 3), k_col = ((lane >> 3) & 1) * 8;      // B operand from K rows
  const int v_row = (lane & 7) + (((lane >> 3) & 1) << 3), v_col = (lane >> 4) * 8;      // B operand from V rows (trans)

  load_qk(0, 0);
  cp_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_qk(c + 1, (c + 1) & 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (m_active) {
      const uint32_t sq = stage_u + (uint32_t)(c & 1) * stage_elems * 2;
      const uint32_t sk = sq + QROWS * LDC * 2;
#pragma unroll
      for (int ks = 0; ks < CH / 16; ++ks) {
        uint32_t af[4];
        ldsm_x4(sq + (uint32_t)((mt * 16 + a_row) * LDC + ks * 16 + a_col) * 2, af[0], af[1], af[2], af[3]);
#pragma unroll
        for (int p = 0; p < MAXP; ++p) {
          if (p < my_pairs) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sk + (uint32_t)(((pair0 + p) * 16 + k_row) * LDC + ks * 16 + k_col) * 2, b0, b1, b2, b3);
            mma16816(acc[2 * p], af, b0, b1);
            mma16816(acc[2 * p + 1], af, b2, b3);
          }
        }
      }
    }
    __syncthreads();
  }
  // prefetch the first V chunk while the softmax runs (both stage buffers are free now)
  load_v(0, 0);
  cp_commit();
  // ---------------- softmax in registers: scores never leave the warp that produced them ----------------
  // Each (m tile, key group) warp reduces its own keys; the KG partial maxima / sums of a row meet in shared memory.
  // P is stored UNNORMALISED (exp2(s - rowmax) in [0, 1]); phase 2 divides O by the row sum.
  const float sl2 = a.scale * 1.4426950408889634f;
  const int32_t* km = a.key_mask ? a.key_mask + (size_t)b * a.Tk : nullptr;
  const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p < my_pairs) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = (pair0 + p) * 16 + t * 8 + t4 * 2 + e;
          const bool ok = col < a.Tk && (!km || km[col] != 0);
          acc[2 * p + t][e] = ok ? acc[2 * p + t][e] * sl2 : -INFINITY;
          acc[2 * p + t][2 + e] = ok ? acc[2 * p + t][2 + e] * sl2 : -INFINITY;
          mx_lo = fmaxf(mx_lo, acc[2 * p + t][e]);
          mx_hi = fmaxf(mx_hi, acc[2 * p + t][2 + e]);
        }
      }
    }
  }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  if (t4 == 0) {
    sRed[r_lo * KG + kg] = mx_lo;
    sRed[r_hi * KG + kg] = mx_hi;
  }
  __syncthreads();
  float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
  for (int k = 0; k < KG; ++k) {
    m_lo = fmaxf(m_lo, sRed[r_lo * KG + k]);
    m_hi = fmaxf(m_hi, sRed[r_hi * KG + k]);
  }
  const float ref_lo = m_lo == -INFINITY ? 0.f : m_lo, ref_hi = m_hi == -INFINITY ? 0.f : m_hi;   // fully masked row: P = 0
  __syncthreads();                                           // every warp has read the maxima: sRed is reused for the sums
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p < my_pairs) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int col = (pair0 + p) * 16 + t * 8 + t4 * 2;
        const float p0 = exp2f(acc[2 * p + t][0] - ref_lo), p1 = exp2f(acc[2 * p + t][1] - ref_lo);
        const float p2 = exp2f(acc[2 * p + t][2] - ref_hi), p3 = exp2f(acc[2 * p + t][3] - ref_hi);
        const uint32_t lo = pack2(p0, p1), hi = pack2(p2, p3);
        // the row sum is taken over the ROUNDED probabilities phase 2 multiplies with
        sum_lo += __uint_as_float(lo << 16) + __uint_as_float(lo & 0xffff0000u);
        sum_hi += __uint_as_float(hi << 16) + __uint_as_float(hi & 0xffff0000u);
        *reinterpret_cast<uint32_t*>(sP + (size_t)r_lo * ldp + col) = lo;
        *reinterpret_cast<uint32_t*>(sP + (size_t)r_hi * ldp + col) = hi;
      }
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
  if (t4 == 0) {
    sRed[r_lo * KG + kg] = sum_lo;
    sRed[r_hi * KG + kg] = sum_hi;
  }
  __syncthreads();
  float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
  for (int k = 0; k < KG; ++k) {
    l_lo += sRed[r_lo * KG + k];
    l_hi += sRed[r_hi * KG + k];
  }
  const float i_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, i_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  // ---------------- phase 2: O[:, chunk] = P . V[:, chunk] ----------------
  __nv_bfloat16* go = reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)b * a.o_sb + (size_t)h * a.dh;
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_v(c + 1, (c + 1) & 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (m_active) {
      const uint32_t sv = stage_u + (uint32_t)(c & 1) * stage_elems * 2;
      float o[NPW][2][4];
#pragma unroll
      for (int i = 0; i < NPW; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) o[i][j][0] = o[i][j][1] = o[i][j][2] = o[i][j][3] = 0.f;
      for (int kk = 0; kk < npairs_all; ++kk) {
        uint32_t af[4];
        ldsm_x4(sP_u + (uint32_t)((mt * 16 + a_row) * ldp + kk * 16 + a_col) * 2, af[0], af[1], af[2], af[3]);
#pragma unroll
        for (int i = 0; i < NPW; ++i) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sv + (uint32_t)((kk * 16 + v_row) * LDC + (kg * NPW + i) * 16 + v_col) * 2, b0, b1, b2, b3);
          mma16816(o[i][0], af, b0, b1);
          mma16816(o[i][1], af, b2, b3);
        }
      }
#pragma unroll
      for (int i = 0; i < NPW; ++i) {
        const int col = c * CH + (kg * NPW + i) * 16 + t4 * 2;
        if (r_lo < a.Tq) {
          *reinterpret_cast<uint32_t*>(go + (size_t)r_lo * a.o_st + col) = pack2(o[i][0][0] * i_lo, o[i][0][1] * i_lo);
          *reinterpret_cast<uint32_t*>(go + (size_t)r_lo * a.o_st + col + 8) = pack2(o[i][1][0] * i_lo, o[i][1][1] * i_lo);
        }
        if (r_hi < a.Tq) {
          *reinterpret_cast<uint32_t*>(go + (size_t)r_hi * a.o_st + col) = pack2(o[i][0][2] * i_hi, o[i][0][3] * i_hi);
          *reinterpret_cast<uint32_t*>(go + (size_t)r_hi * a.o_st + col + 8) = pack2(o[i][1][2] * i_hi, o[i][1][3] * i_hi);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace xa

// Returns 1 when the shape is not covered (caller falls back to the generic kernel), <= 0 otherwise.
int attention_x_launch(const tcavp_attn_args& a, cudaStream_t stream) {
  if (a.causal || a.Tq > 64 || a.Tk > 256 || a.Tk < 1 || a.dh % xa::CH != 0) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.out)) return 1;
  if (a.q_sb % 8 || a.q_st % 8 || a.k_sb % 8 || a.k_st % 8 || a.v_sb % 8 || a.v_st % 8 || a.o_sb % 2 || a.o_st % 2) return 1;
  const int tkp = (a.Tk + 15) / 16 * 16;
  const int qrows = a.Tq <= 32 ? 32 : 64;
  const size_t smem = (size_t)2 * (qrows + tkp) * xa::LDC * 2 + (size_t)qrows * (tkp + 8) * 2 + (size_t)qrows * 4 * 4;
  if (qrows == 32) {       // 2 m tiles x 4 key groups: <= 4 pairs (64 keys) per warp
    TCAVP_CUDA(cudaFuncSetAttribute(xa::attn_x_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xa::attn_x_kernel<32, 4><<<a.B * a.H, xa::THREADS, smem, stream>>>(a, tkp);
  } else {                 // 4 m tiles x 2 key groups: <= 8 pairs (128 keys) per warp
    TCAVP_CUDA(cudaFuncSetAttribute(xa::attn_x_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xa::attn_x_kernel<64, 8><<<a.B * a.H, xa::THREADS, smem, stream>>>(a, tkp);
  }
  return check_launch("attn_x_kernel");
}

}  // namespace tcavp
