// Fused feed-forward sub-layer of a d_model = 64 post-norm transformer encoder layer on tcgen05 / TMEM / TMA:
//     out = LayerNorm(x + linear2(relu(linear1(x))))            (torch nn.TransformerEncoderLayer, reference scripts/train.py:358 —
// the lane-polygon encoder: d_model 64, dim_feedforward 2048, 64 points per scene).  As two GEMMs the [rows, 2048] hidden activation
// makes a round trip through HBM (1 GB written + 1 GB read per layer at 4096 scenes) and both GEMMs are bound by it (K = 64 / N = 64).
// Here the hidden activation never leaves the SM:
//   warp 0      TMA producer: x tiles (two 128-row blocks per CTA step) and, per 128-unit hidden chunk, one stage of
//               { W1 rows [128 c, 128 c + 128) (K-major, 16 KB), W2 columns of the same units (two 64 x 64 boxes, 16 KB) }
//   warp 1      one thread issues tcgen05.mma:  S_k = X_k . W1_c^T (M 128, N 128, K 64) and O_k += P_k . W2_c^T (M 128, N 64, K 128) with the
//               A operand P_k read from TENSOR MEMORY; the two row blocks k = A, B share every weight stage and alternate on the pipe
//   warps 4-7   block A, one thread per row: tcgen05.ld S -> + b1 -> ReLU -> bf16 -> tcgen05.st P; after the last chunk
//               O + b2 + x -> LayerNorm over the 64 features (in the thread) -> bf16 -> one 128-byte row store
//   warps 8-11  block B
// TMEM: S_A | S_B (128 fp32 columns each) | P_A | P_B (64 columns of packed bf16) | O_A | O_B (64) = 512 columns.
// What bounds it: every fp32 score is read out of TMEM once (tcgen05.ld: ~64 B / clk / SM) — 256 rows x F x 4 B = 2 MB per CTA step,
// 32 K cycles, against 16 K cycles of MMA work; measured 177 us per launch at 262 144 rows = 1.5 x that floor.  Sixteen epilogue warps
// (two per lane quarter and block) were tried and are no faster (192 us): the reads, not the per-warp chain, are the limit.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace tcavp {
namespace ffn {

constexpr int THREADS = 384;
constexpr int E = 64;                 // d_model
constexpr int NC = 128;               // hidden units per chunk
constexpr int MAX_STAGES = 6;
constexpr uint32_t X_BYTES = 128u * 128u;          // one 128-row block of x (64 bf16 = 128 bytes per row)
constexpr uint32_t W1_BYTES = (uint32_t)NC * 128u;
constexpr uint32_t W2_BYTES = 2u * 64u * 128u;
constexpr uint32_t STAGE_BYTES = W1_BYTES + W2_BYTES;
constexpr uint32_t COL_S = 0, COL_P = 256, COL_O = 384;   // + 128 k / 64 k / 64 k for block k

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major shared-memory operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {   // A from TMEM
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(r[i]));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct Geo {
  long long M;
  int F, nch, stages;
  const float *b1, *b2, *ln_w, *ln_b;
  float eps;
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
};

// barrier block (8 bytes each): w_full[MAX_STAGES], w_empty[MAX_STAGES], x_full[2], x_empty[2], then per block k:
// s_full, s_free, p_ready, p_free, o_full, o_free
constexpr uint32_t BAR_WFULL = 0, BAR_WEMPTY = 8 * MAX_STAGES, BAR_XFULL = 16 * MAX_STAGES, BAR_XEMPTY = BAR_XFULL + 16, BAR_BLK = BAR_XEMPTY + 16;
constexpr uint32_t BK_SFULL = 0, BK_SFREE = 8, BK_PREADY = 16, BK_PFREE = 24, BK_OFULL = 32, BK_OFREE = 40, BK_BYTES = 48;
constexpr uint32_t BAR_TMEM = BAR_BLK + 2 * BK_BYTES, BAR_BYTES = BAR_TMEM + 16;

// 64 score columns of a row -> + b1 -> ReLU -> 32 packed bf16 pairs
__device__ __forceinline__ void relu_pack64(uint32_t srow, const float* bias, uint32_t (&pk)[32], bool last_of_chunk, uint32_t sfree_bar, int lane) {
  uint32_t r[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_ld16_issue(srow + i * 16, r[i]);
  tmem_ld_wait();
  if (last_of_chunk) {          // every score column of the chunk is in registers: the MMA warp may overwrite S with the next chunk
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(sfree_bar);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    tmem_ld_fence(r[i]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(bias + i * 16 + q * 4);     // warp-uniform address: broadcast
      pk[i * 8 + q * 2] = pack2(fmaxf(__uint_as_float(r[i][q * 4]) + b.x, 0.f), fmaxf(__uint_as_float(r[i][q * 4 + 1]) + b.y, 0.f));
      pk[i * 8 + q * 2 + 1] = pack2(fmaxf(__uint_as_float(r[i][q * 4 + 2]) + b.z, 0.f), fmaxf(__uint_as_float(r[i][q * 4 + 3]) + b.w, 0.f));
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
ffn64_ln_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w1, const __grid_constant__ CUtensorMap tma_w2, Geo g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sm_x = smem_base + g.stages * STAGE_BYTES;           // 2 buffers x 2 blocks x 16 KB
  const uint32_t bars = sm_x + 4u * X_BYTES;
  uint8_t* gen = smem_raw + (bars - smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + BAR_TMEM);
  float* s_b1 = reinterpret_cast<float*>(gen + BAR_BYTES);            // [F] then b2, ln_w, ln_b [64] each
  float* s_v = s_b1 + g.F;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_pairs = (g.M + 255) / 256;
  const int my_pairs = (long long)blockIdx.x < n_pairs ? (int)((n_pairs - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

  for (int i = threadIdx.x; i < g.F; i += THREADS) s_b1[i] = __ldg(g.b1 + i);
  if (threadIdx.x < E) {
    s_v[threadIdx.x] = __ldg(g.b2 + threadIdx.x);
    s_v[E + threadIdx.x] = __ldg(g.ln_w + threadIdx.x);
    s_v[2 * E + threadIdx.x] = __ldg(g.ln_b + threadIdx.x);
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w2)) : "memory");
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bars + BAR_WFULL + 8 * s, 1);
      mbar_init(bars + BAR_WEMPTY + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + BAR_XFULL + 8 * b, 1);
      mbar_init(bars + BAR_XEMPTY + 8 * b, 1);
      const uint32_t bb = bars + BAR_BLK + b * BK_BYTES;
      mbar_init(bb + BK_SFULL, 1);
      mbar_init(bb + BK_SFREE, 4);
      mbar_init(bb + BK_PREADY, 4);
      mbar_init(bb + BK_PFREE, 1);
      mbar_init(bb + BK_OFULL, 1);
      mbar_init(bb + BK_OFREE, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + BAR_TMEM), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t use = 0;
      for (int n = 0; n < my_pairs; ++n) {
        const long long row0 = ((long long)blockIdx.x + (long long)n * gridDim.x) * 256;
        const int xb = n & 1;
        mbar_wait(bars + BAR_XEMPTY + 8 * xb, (((uint32_t)n >> 1) & 1u) ^ 1u);
        const uint32_t xfull = bars + BAR_XFULL + 8 * xb;
        mbar_expect_tx(xfull, 2u * X_BYTES);
        tma_load_2d(sm_x + (uint32_t)xb * 2u * X_BYTES, &tma_x, xfull, 0, (int)row0);
        tma_load_2d(sm_x + (uint32_t)xb * 2u * X_BYTES + X_BYTES, &tma_x, xfull, 0, (int)row0 + 128);
        for (int c = 0; c < g.nch; ++c) {
          mbar_wait(bars + BAR_WEMPTY + 8 * s, (use & 1u) ^ 1u);
          const uint32_t full = bars + BAR_WFULL + 8 * s, st = smem_base + s * STAGE_BYTES;
          mbar_expect_tx(full, STAGE_BYTES);
          tma_load_2d(st, &tma_w1, full, 0, c * NC);
          tma_load_2d(st + W1_BYTES, &tma_w2, full, c * NC, 0);
          tma_load_2d(st + W1_BYTES + W2_BYTES / 2, &tma_w2, full, c * NC + 64, 0);
          if (++s == g.stages) { s = 0; ++use; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_s = idesc_f16(128, NC), id_o = idesc_f16(128, E);
      int s = 0;
      uint32_t use = 0, q = 0;          // q: running chunk counter (both blocks advance together)
      auto issue_s = [&](int k, uint32_t sx, uint32_t sw1) {
#pragma unroll
        for (int kk = 0; kk < E / 16; ++kk)
          umma_ss(tmem_base + COL_S + k * NC, desc_kmajor(sx + (uint32_t)k * X_BYTES + kk * 32u), desc_kmajor(sw1 + kk * 32u), id_s, kk != 0);
        umma_commit(bars + BAR_BLK + k * BK_BYTES + BK_SFULL);
      };
      for (int n = 0; n < my_pairs; ++n) {
        const int xb = n & 1;
        const uint32_t sx = sm_x + (uint32_t)xb * 2u * X_BYTES;
        mbar_wait(bars + BAR_XFULL + 8 * xb, ((uint32_t)n >> 1) & 1u);
        mbar_wait(bars + BAR_WFULL + 8 * s, use & 1u);
        tc_fence_after();
        for (int k = 0; k < 2; ++k) {
          if (q > 0) {                  // the epilogue has read this block's previous score chunk
            mbar_wait(bars + BAR_BLK + k * BK_BYTES + BK_SFREE, (q - 1) & 1u);
            tc_fence_after();
          }
          issue_s(k, sx, smem_base + s * STAGE_BYTES);
        }
        for (int c = 0; c < g.nch; ++c, ++q) {
          const uint32_t st = smem_base + s * STAGE_BYTES;
          int sn = s + 1;
          uint32_t usen = use;
          if (sn == g.stages) { sn = 0; ++usen; }
          // the next chunk's score products go out as soon as the epilogue warps have READ this chunk's scores (not once they have
          // finished with them): S(c + 1) runs on the pipe under the bias / ReLU / pack work of chunk c, so a block's epilogues run
          // back to back instead of waiting a product latency per chunk
          if (c + 1 < g.nch) {
            mbar_wait(bars + BAR_WFULL + 8 * sn, usen & 1u);
            for (int k = 0; k < 2; ++k) {
              mbar_wait(bars + BAR_BLK + k * BK_BYTES + BK_SFREE, q & 1u);
              tc_fence_after();
              issue_s(k, sx, smem_base + sn * STAGE_BYTES);
            }
          }
          for (int k = 0; k < 2; ++k) {
            const uint32_t bb = bars + BAR_BLK + k * BK_BYTES;
            mbar_wait(bb + BK_PREADY, q & 1u);
            if (c == 0 && n > 0) mbar_wait(bb + BK_OFREE, ((uint32_t)n - 1) & 1u);     // the previous tile's output has been drained
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < NC / 16; ++kk)
              umma_ts(tmem_base + COL_O + k * E, tmem_base + COL_P + k * (NC / 2) + kk * 8,
                      desc_kmajor(st + W1_BYTES + (uint32_t)(kk >> 2) * (W2_BYTES / 2) + (uint32_t)(kk & 3) * 32u), id_o, (c | kk) != 0);
            umma_commit(bb + BK_PFREE);
            if (c == g.nch - 1) umma_commit(bb + BK_OFULL);
          }
          umma_commit(bars + BAR_WEMPTY + 8 * s);      // every MMA reading this stage (S of this chunk, O of this chunk) has been issued
          s = sn; use = usen;
        }
        umma_commit(bars + BAR_XEMPTY + 8 * xb);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: block k = (warp - 4) / 4, TMEM lane quarter wq = warp % 4 =====================
    const int k = (warp - 4) >> 2, wq = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t bb = bars + BAR_BLK + k * BK_BYTES;
    const uint32_t tS = lane_base + COL_S + k * NC, tP = lane_base + COL_P + k * (NC / 2), tO = lane_base + COL_O + k * E;
    uint32_t q = 0;
    for (int n = 0; n < my_pairs; ++n) {
      const long long row = ((long long)blockIdx.x + (long long)n * gridDim.x) * 256 + k * 128 + wq * 32 + lane;
      for (int c = 0; c < g.nch; ++c, ++q) {
        mbar_wait(bb + BK_SFULL, q & 1u);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pk[32];
          relu_pack64(tS + h * 64, s_b1 + c * NC + h * 64, pk, h == 1, bb + BK_SFREE, lane);
          if (h == 0 && q > 0) {        // the product that read the previous chunk's activations must have retired before they are overwritten
            mbar_wait(bb + BK_PFREE, (q - 1) & 1u);
            tc_fence_after();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w[8] = {pk[i * 8], pk[i * 8 + 1], pk[i * 8 + 2], pk[i * 8 + 3], pk[i * 8 + 4], pk[i * 8 + 5], pk[i * 8 + 6], pk[i * 8 + 7]};
            tmem_st8(tP + h * 32 + i * 8, w);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bb + BK_PREADY);
      }
      // ---- output row: O + b2 + x -> LayerNorm -> bf16
      mbar_wait(bb + BK_OFULL, (uint32_t)n & 1u);
      tc_fence_after();
      uint32_t r[4][16];
#pragma unroll
      for (int i = 0; i < 4; ++i) tmem_ld16_issue(tO + i * 16, r[i]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bb + BK_OFREE);
      if (row < g.M) {
        float v[E];
        const uint4* xp = reinterpret_cast<const uint4*>(g.x + (size_t)row * E);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tmem_ld_fence(r[i]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 u = __ldg(xp + i * 2 + h);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = i * 16 + h * 8 + e * 2;
              // the sub-layer output is rounded to bf16 before the norm, like every activation that crosses a sub-layer boundary in
              // the bf16 mode: the kernel then agrees with the two-GEMM + LayerNorm path up to the fp32 accumulation order
              const uint32_t yy = pack2(__uint_as_float(r[i][h * 8 + e * 2]) + s_v[j] + __uint_as_float(w[e] << 16),
                                        __uint_as_float(r[i][h * 8 + e * 2 + 1]) + s_v[j + 1] + __uint_as_float(w[e] & 0xffff0000u));
              v[j] = __uint_as_float(yy << 16);
              v[j + 1] = __uint_as_float(yy & 0xffff0000u);
            }
          }
        }
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < E; ++j) mean += v[j];
        mean *= 1.f / E;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < E; ++j) var = fmaf(v[j] - mean, v[j] - mean, var);
        const float rstd = rsqrtf(var * (1.f / E) + g.eps);
        uint4* op = reinterpret_cast<uint4*>(g.out + (size_t)row * E);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = i * 8 + e * 2;
            w[e] = pack2((v[j] - mean) * rstd * s_v[E + j] + s_v[2 * E + j], (v[j + 1] - mean) * rstd * s_v[E + j + 1] + s_v[2 * E + j + 1]);
          }
          op[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
// [rows, cols] bf16 row-major view (row stride = cols); box = box_rows x 64 columns, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

}  // namespace ffn
}  // namespace tcavp

extern "C" int tcavp_ffn64_ln(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* ln_w, const float* ln_b,
                              float eps, void* out, long long M, int F, tcavp_stream_t stream) {
  using namespace tcavp;
  using namespace tcavp::ffn;
  TCAVP_REQUIRE(M >= 0 && F >= NC && F % NC == 0 && F <= 8192, "tcavp_ffn64_ln: dim_feedforward must be a multiple of 128 in [128, 8192] (got %d)", F);
  if (M == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && w1 && b1 && w2 && b2 && ln_w && ln_b && out, "tcavp_ffn64_ln: null pointer");
  TCAVP_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w1) % 16 == 0 && reinterpret_cast<uintptr_t>(w2) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "tcavp_ffn64_ln: x, w1, w2, out must be 16-byte aligned");
  TCAVP_REQUIRE(M + 256 < 0x7fffffffLL, "tcavp_ffn64_ln: too many rows");
  Geo g;
  g.M = M; g.F = F; g.nch = F / NC;
  g.b1 = b1; g.b2 = b2; g.ln_w = ln_w; g.ln_b = ln_b; g.eps = eps;
  g.x = reinterpret_cast<const __nv_bfloat16*>(x);
  g.out = reinterpret_cast<__nv_bfloat16*>(out);
  const size_t fixed = 4u * X_BYTES + BAR_BYTES + (size_t)(F + 3 * E) * sizeof(float) + 1024;
  int stages = (int)((227u * 1024u - fixed) / STAGE_BYTES);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  TCAVP_REQUIRE(stages >= 2, "tcavp_ffn64_ln: shared memory budget");
  g.stages = stages;
  CUtensorMap mx, mw1, mw2;
  if (make_map(&mx, x, M, E, 128) || make_map(&mw1, w1, F, E, NC) || make_map(&mw2, w2, E, F, 64)) {
    set_error("tcavp_ffn64_ln: cuTensorMapEncodeTiled failed");
    return TCAVP_ERR_CUDA;
  }
  const size_t smem = (size_t)stages * STAGE_BYTES + fixed;
  const long long pairs = (M + 255) / 256;
  const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
  TCAVP_CUDA(cudaFuncSetAttribute(ffn64_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ffn64_ln_kernel<<<grid, THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mx, mw1, mw2, g);
  return check_launch("ffn64_ln_kernel");
}
