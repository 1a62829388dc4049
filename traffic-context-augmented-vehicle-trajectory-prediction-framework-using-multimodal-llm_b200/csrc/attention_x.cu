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
  const bool drop = a.drop_thresh != 0;       // dropout on the probabilities (train mode), mask index ((b*H + h)*Tq + i)*Tk + j
  const uint32_t dkey = drop ? drop_key(a.drop_seed, a.drop_site) : 0u;
  const unsigned long long dbase = (unsigned long long)blockIdx.x * a.Tq * (unsigned long long)a.Tk;
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
        uint32_t lo = pack2(p0, p1), hi = pack2(p2, p3);
        // the row sum is taken over the ROUNDED probabilities phase 2 multiplies with
        sum_lo += __uint_as_float(lo << 16) + __uint_as_float(lo & 0xffff0000u);
        sum_hi += __uint_as_float(hi << 16) + __uint_as_float(hi & 0xffff0000u);
        if (drop) {      // train mode: the P.V product sees keep ? p / (1 - p_drop) : 0 (row sums above stay undropped)
          const unsigned long long i_lo_ = dbase + (unsigned long long)r_lo * a.Tk + (unsigned)col, i_hi_ = dbase + (unsigned long long)r_hi * a.Tk + (unsigned)col;
          lo = pack2(drop_keep(dkey, i_lo_, a.drop_thresh) ? p0 * a.drop_scale : 0.f, drop_keep(dkey, i_lo_ + 1, a.drop_thresh) ? p1 * a.drop_scale : 0.f);
          hi = pack2(drop_keep(dkey, i_hi_, a.drop_thresh) ? p2 * a.drop_scale : 0.f, drop_keep(dkey, i_hi_ + 1, a.drop_thresh) ? p3 * a.drop_scale : 0.f);
        }
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

// ------------------------------------------------------------------------------------------------
// Backward of the few-query / wide-head attention (fine-tune step of the LTSF cross-attention, train.py:793-798).
// One CTA owns a (batch, head): every dq / dk / dv row of that head is produced here, so outputs are plain stores.
//   phase 1   S = Q K^T and dP = dO V^T, both accumulated in registers over 64-wide head-dim chunks (Q, dO, K, V staged once)
//   softmax   P = softmax(scale S + mask);  D_i = sum_j P_ij dP_ij;  dS = scale P o (dP - D)   (registers + a QROWS x KG exchange)
//             P and dS go to shared memory in bf16
//   phase 2   per chunk (Q, dO, K staged again):  dQ_c = dS K_c,   dK_c = dS^T Q_c,   dV_c = P^T dO_c
// ------------------------------------------------------------------------------------------------
template <int QROWS, int MAXP>
__global__ void __launch_bounds__(THREADS) attn_x_bwd_kernel(tcavp_attn_args a, const __nv_bfloat16* __restrict__ dout, long long do_sb, long long do_st,
                                                             __nv_bfloat16* __restrict__ dq, long long dq_sb, long long dq_st, void* __restrict__ dk,
                                                             long long dk_sb, long long dk_st, void* __restrict__ dv, long long dv_sb, long long dv_st,
                                                             int dkv_bf16, int tkp) {
  constexpr int MT = QROWS / 16;
  constexpr int KG = 8 / MT;
  constexpr int NPW = (CH / 16) / KG;
  extern __shared__ __align__(16) uint8_t smem[];
  // stage: Q rows | dO rows | K rows | V rows, each LDC wide;  P, dS : QROWS x (tkp + 8) bf16;  red : QROWS x KG x 2 fp32
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sdO = sQ + QROWS * LDC;
  __nv_bfloat16* sK = sdO + QROWS * LDC;
  __nv_bfloat16* sV = sK + (size_t)tkp * LDC;
  const int ldp = tkp + 8;
  __nv_bfloat16* sP = sV + (size_t)tkp * LDC;
  __nv_bfloat16* sdS = sP + (size_t)QROWS * ldp;
  float* sRed = reinterpret_cast<float*>(sdS + (size_t)QROWS * ldp);
  const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ), sdO_u = (uint32_t)__cvta_generic_to_shared(sdO);
  const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK), sV_u = (uint32_t)__cvta_generic_to_shared(sV);
  const uint32_t sP_u = (uint32_t)__cvta_generic_to_shared(sP), sdS_u = (uint32_t)__cvta_generic_to_shared(sdS);

  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* gq = reinterpret_cast<const __nv_bfloat16*>(a.q) + (size_t)b * a.q_sb + (size_t)h * a.dh;
  const __nv_bfloat16* gk = reinterpret_cast<const __nv_bfloat16*>(a.k) + (size_t)b * a.k_sb + (size_t)h * a.dh;
  const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(a.v) + (size_t)b * a.v_sb + (size_t)h * a.dh;
  const __nv_bfloat16* gdo = dout + (size_t)b * do_sb + (size_t)h * a.dh;
  const int nchunks = a.dh / CH;

  auto load_chunk = [&](int c, bool with_v) {
    for (int i = tid; i < QROWS * (CH / 8); i += THREADS) {
      const int r = i >> 3, v = (i & 7) * 8;
      const bool ok = r < a.Tq;
      cp_async16(sQ_u + (uint32_t)(r * LDC + v) * 2, ok ? gq + (size_t)r * a.q_st + c * CH + v : gq, ok);
      cp_async16(sdO_u + (uint32_t)(r * LDC + v) * 2, ok ? gdo + (size_t)r * do_st + c * CH + v : gdo, ok);
    }
    for (int i = tid; i < tkp * (CH / 8); i += THREADS) {
      const int r = i >> 3, v = (i & 7) * 8;
      const bool ok = r < a.Tk;
      cp_async16(sK_u + (uint32_t)(r * LDC + v) * 2, ok ? gk + (size_t)r * a.k_st + c * CH + v : gk, ok);
      if (with_v) cp_async16(sV_u + (uint32_t)(r * LDC + v) * 2, ok ? gv + (size_t)r * a.v_st + c * CH + v : gv, ok);
    }
    cp_commit();
  };

  // ---------------- phase 1 ----------------
  const int mt = warp % MT, kg = warp / MT;
  const int npairs_all = tkp / 16;
  const int base_p = npairs_all / KG, rem_p = npairs_all % KG;
  const int my_pairs = base_p + (kg < rem_p ? 1 : 0);
  const int pair0 = kg * base_p + min(kg, rem_p);
  float acs[MAXP * 2][4], acd[MAXP * 2][4];
#pragma unroll
  for (int i = 0; i < MAXP * 2; ++i) {
    acs[i][0] = acs[i][1] = acs[i][2] = acs[i][3] = 0.f;
    acd[i][0] = acd[i][1] = acd[i][2] = acd[i][3] = 0.f;
  }
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;   // A operand, row-major source
  const int k_row = (lane & 7) + ((lane >> 4) << 3), k_col = ((lane >> 3) & 1) * 8;      // B operand from [key][dim] rows
  const int v_row = (lane & 7) + (((lane >> 3) & 1) << 3), v_col = (lane >> 4) * 8;      // B operand, transposed source
  const int t_m = (lane & 7) + ((lane >> 4) << 3), t_n = ((lane >> 3) & 1) * 8;          // A operand, transposed source
  for (int c = 0; c < nchunks; ++c) {
    load_chunk(c, true);
    cp_wait<0>();
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < CH / 16; ++ks) {
      uint32_t aq[4], ad[4];
      ldsm_x4(sQ_u + (uint32_t)((mt * 16 + a_row) * LDC + ks * 16 + a_col) * 2, aq[0], aq[1], aq[2], aq[3]);
      ldsm_x4(sdO_u + (uint32_t)((mt * 16 + a_row) * LDC + ks * 16 + a_col) * 2, ad[0], ad[1], ad[2], ad[3]);
#pragma unroll
      for (int p = 0; p < MAXP; ++p) {
        if (p < my_pairs) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sK_u + (uint32_t)(((pair0 + p) * 16 + k_row) * LDC + ks * 16 + k_col) * 2, b0, b1, b2, b3);
          mma16816(acs[2 * p], aq, b0, b1);
          mma16816(acs[2 * p + 1], aq, b2, b3);
          ldsm_x4(sV_u + (uint32_t)(((pair0 + p) * 16 + k_row) * LDC + ks * 16 + k_col) * 2, b0, b1, b2, b3);
          mma16816(acd[2 * p], ad, b0, b1);
          mma16816(acd[2 * p + 1], ad, b2, b3);
        }
      }
    }
    __syncthreads();
  }
  // ---------------- softmax statistics, D, dS ----------------
  const float sl2 = a.scale * 1.4426950408889634f;
  const int32_t* km = a.key_mask ? a.key_mask + (size_t)b * a.Tk : nullptr;
  const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
  const bool drop = a.drop_thresh != 0;       // dropout on the probabilities (train mode), mask index ((b*H + h)*Tq + i)*Tk + j
  const uint32_t dkey = drop ? drop_key(a.drop_seed, a.drop_site) : 0u;
  const unsigned long long dbase = (unsigned long long)blockIdx.x * a.Tq * (unsigned long long)a.Tk;
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
          acs[2 * p + t][e] = ok ? acs[2 * p + t][e] * sl2 : -INFINITY;
          acs[2 * p + t][2 + e] = ok ? acs[2 * p + t][2 + e] * sl2 : -INFINITY;
          mx_lo = fmaxf(mx_lo, acs[2 * p + t][e]);
          mx_hi = fmaxf(mx_hi, acs[2 * p + t][2 + e]);
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
  const float ref_lo = m_lo == -INFINITY ? 0.f : m_lo, ref_hi = m_hi == -INFINITY ? 0.f : m_hi;
  __syncthreads();
  if (drop) {     // dP = f o (dO V^T), f = keep ? 1 / (1 - p_drop) : 0
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
      if (p < my_pairs) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const unsigned col = (unsigned)((pair0 + p) * 16 + t * 8 + t4 * 2 + e);
            acd[2 * p + t][e] = drop_keep(dkey, dbase + (unsigned long long)r_lo * a.Tk + col, a.drop_thresh) ? acd[2 * p + t][e] * a.drop_scale : 0.f;
            acd[2 * p + t][2 + e] = drop_keep(dkey, dbase + (unsigned long long)r_hi * a.Tk + col, a.drop_thresh) ? acd[2 * p + t][2 + e] * a.drop_scale : 0.f;
          }
        }
      }
    }
  }
  // unnormalised probabilities replace the scores; partial (sum p, sum p * dP) per row
  float l_lo = 0.f, l_hi = 0.f, pd_lo = 0.f, pd_hi = 0.f;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p < my_pairs) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pl = exp2f(acs[2 * p + t][e] - ref_lo), ph = exp2f(acs[2 * p + t][2 + e] - ref_hi);
          acs[2 * p + t][e] = pl;
          acs[2 * p + t][2 + e] = ph;
          l_lo += pl;
          l_hi += ph;
          pd_lo += pl * acd[2 * p + t][e];
          pd_hi += ph * acd[2 * p + t][2 + e];
        }
      }
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, o);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, o);
    pd_lo += __shfl_xor_sync(0xffffffffu, pd_lo, o);
    pd_hi += __shfl_xor_sync(0xffffffffu, pd_hi, o);
  }
  float* sRed2 = sRed + QROWS * KG;
  if (t4 == 0) {
    sRed[r_lo * KG + kg] = l_lo;
    sRed[r_hi * KG + kg] = l_hi;
    sRed2[r_lo * KG + kg] = pd_lo;
    sRed2[r_hi * KG + kg] = pd_hi;
  }
  __syncthreads();
  float L_lo = 0.f, L_hi = 0.f, D_lo = 0.f, D_hi = 0.f;
#pragma unroll
  for (int k = 0; k < KG; ++k) {
    L_lo += sRed[r_lo * KG + k];
    L_hi += sRed[r_hi * KG + k];
    D_lo += sRed2[r_lo * KG + k];
    D_hi += sRed2[r_hi * KG + k];
  }
  const float i_lo = L_lo > 0.f ? 1.f / L_lo : 0.f, i_hi = L_hi > 0.f ? 1.f / L_hi : 0.f;
  D_lo *= i_lo;      // D_i = sum_j P_ij dP_ij
  D_hi *= i_hi;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p < my_pairs) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int col = (pair0 + p) * 16 + t * 8 + t4 * 2;
        const float p0 = acs[2 * p + t][0] * i_lo, p1 = acs[2 * p + t][1] * i_lo;
        const float p2 = acs[2 * p + t][2] * i_hi, p3 = acs[2 * p + t][3] * i_hi;
        if (drop) {     // dV = P_d^T dO
          const unsigned long long i_lo_ = dbase + (unsigned long long)r_lo * a.Tk + (unsigned)col, i_hi_ = dbase + (unsigned long long)r_hi * a.Tk + (unsigned)col;
          *reinterpret_cast<uint32_t*>(sP + (size_t)r_lo * ldp + col) =
              pack2(drop_keep(dkey, i_lo_, a.drop_thresh) ? p0 * a.drop_scale : 0.f, drop_keep(dkey, i_lo_ + 1, a.drop_thresh) ? p1 * a.drop_scale : 0.f);
          *reinterpret_cast<uint32_t*>(sP + (size_t)r_hi * ldp + col) =
              pack2(drop_keep(dkey, i_hi_, a.drop_thresh) ? p2 * a.drop_scale : 0.f, drop_keep(dkey, i_hi_ + 1, a.drop_thresh) ? p3 * a.drop_scale : 0.f);
        } else {
          *reinterpret_cast<uint32_t*>(sP + (size_t)r_lo * ldp + col) = pack2(p0, p1);
          *reinterpret_cast<uint32_t*>(sP + (size_t)r_hi * ldp + col) = pack2(p2, p3);
        }
        *reinterpret_cast<uint32_t*>(sdS + (size_t)r_lo * ldp + col) =
            pack2(a.scale * p0 * (acd[2 * p + t][0] - D_lo), a.scale * p1 * (acd[2 * p + t][1] - D_lo));
        *reinterpret_cast<uint32_t*>(sdS + (size_t)r_hi * ldp + col) =
            pack2(a.scale * p2 * (acd[2 * p + t][2] - D_hi), a.scale * p3 * (acd[2 * p + t][3] - D_hi));
      }
    }
  }
  __syncthreads();
  // ---------------- phase 2 ----------------
  __nv_bfloat16* gdq = dq + (size_t)b * dq_sb + (size_t)h * a.dh;
  const int kv_units = npairs_all * (CH / 16);      // (16-key tile, 16-dim pair) units per output (dK or dV)
  for (int c = 0; c < nchunks; ++c) {
    load_chunk(c, false);
    cp_wait<0>();
    __syncthreads();
    {   // dQ_c = dS . K_c : this warp's m tile, NPW 16-dim pairs
      float o[NPW][2][4];
#pragma unroll
      for (int i = 0; i < NPW; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) o[i][j][0] = o[i][j][1] = o[i][j][2] = o[i][j][3] = 0.f;
      for (int kk = 0; kk < npairs_all; ++kk) {
        uint32_t af[4];
        ldsm_x4(sdS_u + (uint32_t)((mt * 16 + a_row) * ldp + kk * 16 + a_col) * 2, af[0], af[1], af[2], af[3]);
#pragma unroll
        for (int i = 0; i < NPW; ++i) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sK_u + (uint32_t)((kk * 16 + v_row) * LDC + (kg * NPW + i) * 16 + v_col) * 2, b0, b1, b2, b3);
          mma16816(o[i][0], af, b0, b1);
          mma16816(o[i][1], af, b2, b3);
        }
      }
#pragma unroll
      for (int i = 0; i < NPW; ++i) {
        const int col = c * CH + (kg * NPW + i) * 16 + t4 * 2;
        if (r_lo < a.Tq) {
          *reinterpret_cast<uint32_t*>(gdq + (size_t)r_lo * dq_st + col) = pack2(o[i][0][0], o[i][0][1]);
          *reinterpret_cast<uint32_t*>(gdq + (size_t)r_lo * dq_st + col + 8) = pack2(o[i][1][0], o[i][1][1]);
        }
        if (r_hi < a.Tq) {
          *reinterpret_cast<uint32_t*>(gdq + (size_t)r_hi * dq_st + col) = pack2(o[i][0][2], o[i][0][3]);
          *reinterpret_cast<uint32_t*>(gdq + (size_t)r_hi * dq_st + col + 8) = pack2(o[i][1][2], o[i][1][3]);
        }
      }
    }
    // dK_c = dS^T . Q_c and dV_c = P^T . dO_c : (16-key tile, 16-dim pair) units dealt round-robin to the warps
    for (int u = warp; u < 2 * kv_units; u += 8) {
      const bool is_v = u >= kv_units;
      const int uu = is_v ? u - kv_units : u;
      const int kt = uu / (CH / 16), np = uu % (CH / 16);
      const uint32_t sA = is_v ? sP_u : sdS_u;      // [q][key]: read transposed -> A[key][q]
      const uint32_t sB = is_v ? sdO_u : sQ_u;      // [q][dim]: contraction over q
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int qs = 0; qs < MT; ++qs) {
        uint32_t af[4], b0, b1, b2, b3;
        ldsm_x4_t(sA + (uint32_t)((qs * 16 + t_m) * ldp + kt * 16 + t_n) * 2, af[0], af[1], af[2], af[3]);
        ldsm_x4_t(sB + (uint32_t)((qs * 16 + v_row) * LDC + np * 16 + v_col) * 2, b0, b1, b2, b3);
        mma16816(o0, af, b0, b1);
        mma16816(o1, af, b2, b3);
      }
      const int j_lo = kt * 16 + g, j_hi = j_lo + 8;
      const int col = c * CH + np * 16 + t4 * 2;
      void* dst = is_v ? dv : dk;
      const long long sb = is_v ? dv_sb : dk_sb, st = is_v ? dv_st : dk_st;
      if (dkv_bf16) {
        __nv_bfloat16* gd = reinterpret_cast<__nv_bfloat16*>(dst) + (size_t)b * sb + (size_t)h * a.dh;
        if (j_lo < a.Tk) {
          *reinterpret_cast<uint32_t*>(gd + (size_t)j_lo * st + col) = pack2(o0[0], o0[1]);
          *reinterpret_cast<uint32_t*>(gd + (size_t)j_lo * st + col + 8) = pack2(o1[0], o1[1]);
        }
        if (j_hi < a.Tk) {
          *reinterpret_cast<uint32_t*>(gd + (size_t)j_hi * st + col) = pack2(o0[2], o0[3]);
          *reinterpret_cast<uint32_t*>(gd + (size_t)j_hi * st + col + 8) = pack2(o1[2], o1[3]);
        }
      } else {
        float* gd = reinterpret_cast<float*>(dst) + (size_t)b * sb + (size_t)h * a.dh;
        if (j_lo < a.Tk) {
          *reinterpret_cast<float2*>(gd + (size_t)j_lo * st + col) = make_float2(o0[0], o0[1]);
          *reinterpret_cast<float2*>(gd + (size_t)j_lo * st + col + 8) = make_float2(o1[0], o1[1]);
        }
        if (j_hi < a.Tk) {
          *reinterpret_cast<float2*>(gd + (size_t)j_hi * st + col) = make_float2(o0[2], o0[3]);
          *reinterpret_cast<float2*>(gd + (size_t)j_hi * st + col + 8) = make_float2(o1[2], o1[3]);
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

// Backward launcher for tcavp_attention_bwd_owned: returns 1 when the shape is not covered.
int attention_x_bwd_launch(const tcavp_attn_args& a, const void* dout, long long do_sb, long long do_st, void* dq, long long dq_sb, long long dq_st,
                           void* dk, long long dk_sb, long long dk_st, void* dv, long long dv_sb, long long dv_st, int dkv_dtype, cudaStream_t stream) {
  if (a.dtype != TCAVP_BF16 || a.causal || a.H != a.Hkv || a.Tq > 64 || a.Tk > 256 || a.Tk < 1 || a.dh % xa::CH != 0) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(dout) || reinterpret_cast<uintptr_t>(dq) % 4 || reinterpret_cast<uintptr_t>(dk) % 8 ||
      reinterpret_cast<uintptr_t>(dv) % 8)
    return 1;
  if (a.q_sb % 8 || a.q_st % 8 || a.k_sb % 8 || a.k_st % 8 || a.v_sb % 8 || a.v_st % 8 || do_sb % 8 || do_st % 8 || dq_sb % 2 || dq_st % 2 ||
      dk_sb % 2 || dk_st % 2 || dv_sb % 2 || dv_st % 2)
    return 1;
  const int tkp = (a.Tk + 15) / 16 * 16;
  const int qrows = a.Tq <= 32 ? 32 : 64;
  const size_t smem = (size_t)(2 * qrows + 2 * tkp) * xa::LDC * 2 + (size_t)2 * qrows * (tkp + 8) * 2 + (size_t)qrows * 4 * 2 * 4;
  if (smem > 220 * 1024) return 1;
  const int bf = dkv_dtype == TCAVP_BF16 ? 1 : 0;
  const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(dout);
  __nv_bfloat16* d_q = reinterpret_cast<__nv_bfloat16*>(dq);
  if (qrows == 32) {
    TCAVP_CUDA(cudaFuncSetAttribute(xa::attn_x_bwd_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xa::attn_x_bwd_kernel<32, 4><<<a.B * a.H, xa::THREADS, smem, stream>>>(a, d_o, do_sb, do_st, d_q, dq_sb, dq_st, dk, dk_sb, dk_st, dv, dv_sb, dv_st,
                                                                          bf, tkp);
  } else {
    TCAVP_CUDA(cudaFuncSetAttribute(xa::attn_x_bwd_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xa::attn_x_bwd_kernel<64, 8><<<a.B * a.H, xa::THREADS, smem, stream>>>(a, d_o, do_sb, do_st, d_q, dq_sb, dq_st, dk, dk_sb, dk_st, dv, dv_sb, dv_st,
                                                                          bf, tkp);
  }
  return check_launch("attn_x_bwd_kernel");
}

}  // namespace tcavp
