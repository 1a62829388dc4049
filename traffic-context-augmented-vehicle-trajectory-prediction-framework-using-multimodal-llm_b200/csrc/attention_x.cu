// Tensor-core attention for FEW queries against a wide head: the LTSF cross-attention
// (reference scripts/train.py:793-798: T_out <= 32 queries, L keys, 2 heads of width H/2 = 384 ... 2048).
//
// The generic warp-per-query kernel re-reads all of K and V for every query row; here one CTA owns a
// (batch, head) pair and streams K, then V, through shared memory exactly once in 64-wide head-dim chunks
// (cp.async double buffering), so the kernel is bound by the single HBM pass over K and V:
//   phase 1   S[32, Tk]   = sum over chunks  Q[:, chunk] . K[:, chunk]^T      (mma.sync m16n8k16, fp32 accum)
//   softmax   P = softmax(scale * S + key mask), fp32 statistics, bf16 probabilities kept in shared memory
//   phase 2   O[:, chunk] = P . V[:, chunk] for every chunk, written straight to global memory
#include "common.cuh"

namespace tcavp {
namespace xa {

constexpr int QROWS = 32;        // query rows per CTA (two m16 tiles), zero padded
constexpr int CH = 64;           // head-dim chunk
constexpr int LDC = CH + 8;      // padded chunk row (elements): conflict-free ldmatrix
constexpr int THREADS = 256;
constexpr int MAX_PAIRS = 4;     // 16-key pairs per warp in phase 1 (Tk <= 256)

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

__global__ void __launch_bounds__(THREADS, 2) attn_x_kernel(tcavp_attn_args a, int tkp) {
  extern __shared__ __align__(16) uint8_t smem[];
  // stage[2] : (QROWS + tkp) x LDC bf16 (phase 1: Q rows then K rows; phase 2: V rows) | S : QROWS x (tkp+4) fp32
  // | P : QROWS x (tkp+8) bf16
  const int stage_elems = (QROWS + tkp) * LDC;
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(smem);
  float* sS = reinterpret_cast<float*>(stage + 2 * (size_t)stage_elems);
  const int lds = tkp + 4;
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(sS + (size_t)QROWS * lds);
  const int ldp = tkp + 8;
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
  const int mt = warp & 1, kg = warp >> 1;                  // m16 tile, key group
  const int npairs_all = tkp / 16;
  const int base_p = npairs_all / 4, rem_p = npairs_all % 4;
  const int my_pairs = base_p + (kg < rem_p ? 1 : 0);
  const int pair0 = kg * base_p + min(kg, rem_p);
  const bool m_active = mt * 16 < a.Tq;
  float acc[MAX_PAIRS * 2][4];
#pragma unroll
  for (int i = 0; i < MAX_PAIRS * 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
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
        for (int p = 0; p < MAX_PAIRS; ++p) {
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
  // scores -> shared memory (fp32, pre-scaled for a base-2 softmax)
  const float sl2 = a.scale * 1.4426950408889634f;
  if (m_active) {
#pragma unroll
    for (int p = 0; p < MAX_PAIRS; ++p) {
      if (p < my_pairs) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int col = (pair0 + p) * 16 + t * 8 + t4 * 2;
          float* r0 = sS + (size_t)(mt * 16 + g) * lds + col;
          float* r1 = sS + (size_t)(mt * 16 + g + 8) * lds + col;
          r0[0] = acc[2 * p + t][0] * sl2; r0[1] = acc[2 * p + t][1] * sl2;
          r1[0] = acc[2 * p + t][2] * sl2; r1[1] = acc[2 * p + t][3] * sl2;
        }
      }
    }
  }
  __syncthreads();
  // ---------------- softmax: warp w owns rows 4w .. 4w+3 ----------------
  const int32_t* km = a.key_mask ? a.key_mask + (size_t)b * a.Tk : nullptr;
  for (int rr = 0; rr < QROWS / 8; ++rr) {
    const int row = warp * (QROWS / 8) + rr;
    __nv_bfloat16* prow = sP + (size_t)row * ldp;
    if (row >= a.Tq) {
      for (int j = lane; j < tkp; j += 32) prow[j] = __float2bfloat16_rn(0.f);
      continue;
    }
    const float* srow = sS + (size_t)row * lds;
    float mx = -INFINITY;
    for (int j = lane; j < a.Tk; j += 32)
      if (!km || km[j] != 0) mx = fmaxf(mx, srow[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < a.Tk; j += 32)
      if (!km || km[j] != 0) sum += exp2f(srow[j] - mx);
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int j = lane; j < tkp; j += 32) {
      const bool ok = j < a.Tk && (!km || km[j] != 0) && sum > 0.f;
      prow[j] = __float2bfloat16_rn(ok ? exp2f(srow[j] - mx) * inv : 0.f);
    }
  }
  __syncthreads();
  // ---------------- phase 2: O[:, chunk] = P . V[:, chunk] ----------------
  const int np = warp >> 1;                                  // 16-dim pair inside the 64-dim chunk
  const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
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
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kk = 0; kk < npairs_all; ++kk) {
        uint32_t af[4], b0, b1, b2, b3;
        ldsm_x4(sP_u + (uint32_t)((mt * 16 + a_row) * ldp + kk * 16 + a_col) * 2, af[0], af[1], af[2], af[3]);
        ldsm_x4_t(sv + (uint32_t)((kk * 16 + v_row) * LDC + np * 16 + v_col) * 2, b0, b1, b2, b3);
        mma16816(o0, af, b0, b1);
        mma16816(o1, af, b2, b3);
      }
      const int col = c * CH + np * 16 + t4 * 2;
      if (r_lo < a.Tq) {
        *reinterpret_cast<uint32_t*>(go + (size_t)r_lo * a.o_st + col) = pack2(o0[0], o0[1]);
        *reinterpret_cast<uint32_t*>(go + (size_t)r_lo * a.o_st + col + 8) = pack2(o1[0], o1[1]);
      }
      if (r_hi < a.Tq) {
        *reinterpret_cast<uint32_t*>(go + (size_t)r_hi * a.o_st + col) = pack2(o0[2], o0[3]);
        *reinterpret_cast<uint32_t*>(go + (size_t)r_hi * a.o_st + col + 8) = pack2(o1[2], o1[3]);
      }
    }
    __syncthreads();
  }
}

}  // namespace xa

// Returns 1 when the shape is not covered (caller falls back to the generic kernel), <= 0 otherwise.
int attention_x_launch(const tcavp_attn_args& a, cudaStream_t stream) {
  if (a.causal || a.Tq > xa::QROWS || a.Tk > 256 || a.Tk < 1 || a.dh % xa::CH != 0) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.out)) return 1;
  if (a.q_sb % 8 || a.q_st % 8 || a.k_sb % 8 || a.k_st % 8 || a.v_sb % 8 || a.v_st % 8 || a.o_sb % 2 || a.o_st % 2) return 1;
  const int tkp = (a.Tk + 15) / 16 * 16;
  const size_t smem = (size_t)2 * (xa::QROWS + tkp) * xa::LDC * 2 + (size_t)xa::QROWS * (tkp + 4) * 4 + (size_t)xa::QROWS * (tkp + 8) * 2;
  TCAVP_CUDA(cudaFuncSetAttribute(xa::attn_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  xa::attn_x_kernel<<<a.B * a.H, xa::THREADS, smem, stream>>>(a, tkp);
  return check_launch("attn_x_kernel");
}

}  // namespace tcavp
