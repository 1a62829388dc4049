// tcgen05 cross-attention for FEW queries against ONE wide shared key/value head: the absorbed LTSF fusion cross-attention
// (reference scripts/train.py:793-798 after engine.py: Engine._absorb_cross — keys = values = X, the backbone output (B, Tk, dh) read in
// place; H = 2 query heads of width dh = hidden size; T_out <= 64 queries per head).
//
// The mma.sync kernel (attention_x.cu) is tensor-bound on this shape (dh = 768: 57 MFLOP per scene, 130 TF/s); on the 5th-gen tensor
// cores the op falls back to its HBM floor (Q' and X read once, Z written once).  Persistent, warp-specialised, one CTA per SM looping
// over scenes; both heads of a scene are stacked into ONE M = 128 tile (rows [64 h, 64 h + Tq) = head h):
//   warp 0       TMA producer.  Phase 1 of a scene: dh / 64 stages of { Q' rows of both heads (2 boxes of 64 x 64), X (Tk16 x 64) };
//                phase 2: dh / 128 stages of { X[:, 128-column chunk] as two Tk16 x 64 boxes } (second pass over X: L2 hits).
//   warp 1       one thread issues tcgen05.mma:  S[128, Tk16] += Q'_chunk . X_chunk^T (both operands K-major from shared memory), then per
//                128-column chunk O = P . X_chunk with A = P read from TENSOR MEMORY (bf16, aliasing S) and X as an MN-major operand
//                (no transposed copy), double-buffered accumulators.
//   warps 4-7    softmax, one thread per score row (tcgen05.ld -> exp2 -> tcgen05.st of bf16 P over the scores), then drain columns
//                [0, 64) of every O chunk (x 1 / row sum -> bf16 -> 32-byte stores)
//   warps 8-11   drain columns [64, 128) of every O chunk
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace tcavp {
namespace xt {

constexpr int THREADS = 384;
constexpr int BOXC = 64;          // bf16 columns per TMA box = 128 bytes = one swizzle row
constexpr int QROWS = 64;         // padded query rows per head
constexpr int NCOL = 128;         // output columns per P.V chunk
constexpr int MAX_STAGES = 8;
constexpr uint32_t Q_BYTES = 2u * QROWS * 128u;    // both heads' 64 x 64 boxes of one k-chunk

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
// Shared-memory operand descriptors (SWIZZLE_128B, version 1): see attention_tm.cu for the layouts.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

struct Geo {
  int B, Tq, Tk, Tk16, dh;
  int stages;
  uint32_t box_bytes, stage_bytes;    // one [Tk16 x 64] box of X; one ring stage
  uint32_t o_col;                     // TMEM: S / P at column 0, the two O buffers at o_col and o_col + 128
  float sl2;                          // softmax scale * log2(e)
  __nv_bfloat16* out; long long o_sb, o_st;
};

// barrier block behind the stage ring (8 bytes each)
constexpr uint32_t BAR_FULL = 0, BAR_EMPTY = 8 * MAX_STAGES, BAR_SFULL = 16 * MAX_STAGES, BAR_PREADY = BAR_SFULL + 8, BAR_OFULL = BAR_PREADY + 8,
                   BAR_OFREE = BAR_OFULL + 16, BAR_TMEM = BAR_OFREE + 16, BAR_ROWSUM = BAR_TMEM + 16 /* 2 x 128 floats */, BAR_BYTES = BAR_ROWSUM + 1024;

// Softmax of one score row per thread (TMEM lane = row): row maximum, then exp2 / row sum, the bf16 probabilities overwrite the score
// columns (P aliases S: unit u's probabilities land in columns [8u, 8u + 8), which the sweep has already consumed).  Batches of three
// 16-column units per tcgen05.wait::ld.  Returns the sum of the ROUNDED probabilities (what P.V multiplies with).
__device__ __noinline__ float softmax_row(const Geo& g, uint32_t srow) {
  const int nu = g.Tk16 >> 4;
  float m = -INFINITY;
  for (int u0 = 0; u0 < nu; u0 += 3) {
    uint32_t r[3][16];
#pragma unroll
    for (int i = 0; i < 3; ++i) tmem_ld16_issue(srow + (u0 + i) * 16, r[i]);     // columns past Tk16 stay inside the allocation and are skipped
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (u0 + i < nu) {
        tmem_ld_fence(r[i]);
        const int c0 = (u0 + i) * 16;
        if (c0 + 16 <= g.Tk) {
#pragma unroll
          for (int c = 0; c < 16; ++c) m = fmaxf(m, __uint_as_float(r[i][c]));
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) m = fmaxf(m, c0 + c < g.Tk ? __uint_as_float(r[i][c]) : -INFINITY);
        }
      }
    }
  }
  const float mref = m == -INFINITY ? 0.f : m * g.sl2;
  float l = 0.f;
  for (int u0 = 0; u0 < nu; u0 += 3) {
    uint32_t r[3][16];
#pragma unroll
    for (int i = 0; i < 3; ++i) tmem_ld16_issue(srow + (u0 + i) * 16, r[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (u0 + i < nu) {
        tmem_ld_fence(r[i]);
        const int c0 = (u0 + i) * 16;
        uint32_t pk[8];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float p0 = ex2(fmaf(__uint_as_float(r[i][c]), g.sl2, -mref)), p1 = ex2(fmaf(__uint_as_float(r[i][c + 1]), g.sl2, -mref));
          if (c0 + 16 > g.Tk) {
            p0 = c0 + c < g.Tk ? p0 : 0.f;
            p1 = c0 + c + 1 < g.Tk ? p1 : 0.f;
          }
          const uint32_t w = pack2(p0, p1);
          l += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
          pk[c >> 1] = w;
        }
        tmem_st8(srow + (u0 + i) * 8, pk);
      }
    }
  }
  tmem_st_wait();
  tc_fence_before();
  return l;
}

// Drains 64 columns of one output row per thread: O x 1 / row sum -> bf16 -> global (128 contiguous bytes).
__device__ __forceinline__ void drain64(uint32_t orow, float inv, __nv_bfloat16* op, bool live) {
  uint32_t r[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_ld16_issue(orow + i * 16, r[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    tmem_ld_fence(r[i]);
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = pack2(__uint_as_float(r[i][2 * e]) * inv, __uint_as_float(r[i][2 * e + 1]) * inv);
    if (live) stg256(op + i * 16, w);
  }
  tc_fence_before();
}

__global__ void __launch_bounds__(THREADS, 1)
attn_xt_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_x, Geo g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + g.stages * g.stage_bytes;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (bars + BAR_TMEM - smem_u32(smem_raw)));
  float* rowsum = reinterpret_cast<float*>(smem_raw + (bars + BAR_ROWSUM - smem_u32(smem_raw)));     // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_items = blockIdx.x < (unsigned)g.B ? (g.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int kch = g.dh / BOXC;          // phase-1 stages per scene
  const int nch = g.dh / NCOL;          // phase-2 stages (output chunks) per scene
  const int nk16 = g.Tk16 >> 4;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bars + BAR_FULL + 8 * s, 1);
      mbar_init(bars + BAR_EMPTY + 8 * s, 1);
    }
    mbar_init(bars + BAR_SFULL, 1);
    mbar_init(bars + BAR_PREADY, 4);
    for (int k = 0; k < 2; ++k) {
      mbar_init(bars + BAR_OFULL + 8 * k, 1);
      mbar_init(bars + BAR_OFREE + 8 * k, 8);
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
      for (int n = 0; n < my_items; ++n) {
        const int b = (int)blockIdx.x + n * (int)gridDim.x;
        for (int c = 0; c < kch; ++c) {
          mbar_wait(bars + BAR_EMPTY + 8 * s, (use & 1u) ^ 1u);
          const uint32_t full = bars + BAR_FULL + 8 * s, st = smem_base + s * g.stage_bytes;
          mbar_expect_tx(full, Q_BYTES + g.box_bytes);
          tma_load_2d(st, &tma_q, full, c * BOXC, b * g.Tq);                       // head 0 -> tile rows [0, 64)
          tma_load_2d(st + Q_BYTES / 2, &tma_q, full, g.dh + c * BOXC, b * g.Tq);  // head 1 -> tile rows [64, 128)
          tma_load_2d(st + Q_BYTES, &tma_x, full, c * BOXC, b * g.Tk);
          if (++s == g.stages) { s = 0; ++use; }
        }
        for (int j = 0; j < nch; ++j) {
          mbar_wait(bars + BAR_EMPTY + 8 * s, (use & 1u) ^ 1u);
          const uint32_t full = bars + BAR_FULL + 8 * s, st = smem_base + s * g.stage_bytes;
          mbar_expect_tx(full, 2u * g.box_bytes);
          tma_load_2d(st, &tma_x, full, j * NCOL, b * g.Tk);
          tma_load_2d(st + g.box_bytes, &tma_x, full, j * NCOL + BOXC, b * g.Tk);
          if (++s == g.stages) { s = 0; ++use; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // tcgen05.mma executes in issue order: scene n + 1's score MMAs are issued after scene n's P.V products, so the S / P columns are
    // never overwritten before the products that read P have retired.
    if (lane == 0) {
      const uint32_t id_s = idesc_f16(128, g.Tk16, 0), id_pv = idesc_f16(128, NCOL, 1);
      int s = 0;
      uint32_t use = 0, q = 0;          // q: running output-chunk counter (accumulator buffer q & 1)
      for (int n = 0; n < my_items; ++n) {
        for (int c = 0; c < kch; ++c) {
          mbar_wait(bars + BAR_FULL + 8 * s, use & 1u);
          tc_fence_after();
          const uint32_t sq = smem_base + s * g.stage_bytes, sx = sq + Q_BYTES;
#pragma unroll
          for (int kk = 0; kk < BOXC / 16; ++kk)
            umma_ss(tmem_base, desc_kmajor(sq + kk * 32u), desc_kmajor(sx + kk * 32u), id_s, (c | kk) != 0);
          umma_commit(bars + BAR_EMPTY + 8 * s);
          if (++s == g.stages) { s = 0; ++use; }
        }
        umma_commit(bars + BAR_SFULL);
        mbar_wait(bars + BAR_PREADY, (uint32_t)n & 1u);
        tc_fence_after();
        for (int j = 0; j < nch; ++j, ++q) {
          const uint32_t buf = q & 1u;
          mbar_wait(bars + BAR_FULL + 8 * s, use & 1u);
          mbar_wait(bars + BAR_OFREE + 8 * buf, ((q >> 1) & 1u) ^ 1u);      // this accumulator buffer's previous chunk has been drained
          tc_fence_after();
          const uint32_t sv = smem_base + s * g.stage_bytes;
          const uint32_t tO = tmem_base + g.o_col + buf * NCOL;
          for (int kk = 0; kk < nk16; ++kk)
            umma_ts(tO, tmem_base + kk * 8, desc_mnmajor(sv + (uint32_t)kk * 2048u, g.box_bytes), id_pv, kk != 0);
          umma_commit(bars + BAR_OFULL + 8 * buf);
          umma_commit(bars + BAR_EMPTY + 8 * s);
          if (++s == g.stages) { s = 0; ++use; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax (warps 4-7) / output drain (warps 4-11) =====================
    const int wq = warp & 3, half = (warp - 4) >> 2;            // TMEM lane quarter; which 64 columns of a chunk this warp drains
    const int row = wq * 32 + lane;                             // tile row: head row >> 6, query row & 63
    const int h = row >> 6, t = row & 63;
    const bool live = t < g.Tq;
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t q = 0;
    for (int n = 0; n < my_items; ++n) {
      const int b = (int)blockIdx.x + n * (int)gridDim.x;
      float l = 0.f;
      if (half == 0) {
        mbar_wait(bars + BAR_SFULL, (uint32_t)n & 1u);
        tc_fence_after();
        l = softmax_row(g, lane_base);
        rowsum[(n & 1) * 128 + row] = l;
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BAR_PREADY);
      }
      __nv_bfloat16* op = g.out + (size_t)b * g.o_sb + (size_t)t * g.o_st + (size_t)h * g.dh + half * 64;
      float inv = 0.f;
      for (int j = 0; j < nch; ++j, ++q) {
        const uint32_t buf = q & 1u;
        mbar_wait(bars + BAR_OFULL + 8 * buf, (q >> 1) & 1u);
        tc_fence_after();
        if (j == 0) {
          if (half != 0) l = rowsum[(n & 1) * 128 + row];       // published before the softmax warps' arrival that released these MMAs
          inv = l > 0.f ? 1.f / l : 0.f;
        }
        drain64(lane_base + g.o_col + buf * NCOL + half * 64, inv, op + j * NCOL, live);
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BAR_OFREE + 8 * buf);
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
// [rows, cols] bf16 view with row stride ld; box = box_rows x 64 columns, 128-byte swizzle, zero fill out of bounds.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BOXC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

}  // namespace xt

// Returns 1 when the shape is not covered (caller falls back to the mma.sync kernel of attention_x.cu), <= 0 otherwise.
int attention_xt_launch(const tcavp_attn_args& a, cudaStream_t stream) {
  using namespace xt;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TCAVP_ATTNX_TCGEN05");      // 0: keep the mma.sync kernel (A/B runs)
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return 1;
  if (a.dtype != TCAVP_BF16 || a.causal || a.key_mask || a.drop_thresh != 0) return 1;
  // two query heads on one shared key / value head whose keys ARE its values (same rows of the same tensor)
  if (a.H != 2 || a.Hkv != 1 || a.k != a.v || a.k_sb != a.v_sb || a.k_st != a.v_st) return 1;
  if (a.Tq < 1 || a.Tq > QROWS || a.Tk < 1 || a.Tk > 256 || a.dh % NCOL != 0 || a.dh < NCOL || a.B < 1) return 1;
  // rows of consecutive scenes must be contiguous (one 2-D tensor map per operand) and TMA-addressable
  if (a.q_sb != (long long)a.Tq * a.q_st || a.k_sb != (long long)a.Tk * a.k_st) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || reinterpret_cast<uintptr_t>(a.out) % 32 || a.q_st % 8 || a.k_st % 8 || a.o_st % 16 || a.o_sb % 16) return 1;
  if ((long long)a.B * a.Tq > 0x7fffffffLL || (long long)a.B * a.Tk > 0x7fffffffLL) return 1;
  Geo g;
  g.B = a.B; g.Tq = a.Tq; g.Tk = a.Tk; g.Tk16 = (a.Tk + 15) / 16 * 16; g.dh = a.dh;
  g.box_bytes = (uint32_t)g.Tk16 * 128u;
  const uint32_t p1 = Q_BYTES + g.box_bytes, p2 = 2u * g.box_bytes;
  g.stage_bytes = ((p1 > p2 ? p1 : p2) + 1023u) & ~1023u;
  g.o_col = (uint32_t)((g.Tk16 + 31) / 32 * 32);
  if (g.o_col + 2u * NCOL > 512u) return 1;
  int stages = (int)((225u * 1024u - BAR_BYTES - 1024u) / g.stage_bytes);
  if (stages < 2) return 1;
  g.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
  g.sl2 = a.scale * 1.4426950408889634f;
  g.out = reinterpret_cast<__nv_bfloat16*>(a.out); g.o_sb = a.o_sb; g.o_st = a.o_st;
  CUtensorMap mq, mx;
  if (make_map(&mq, a.q, (long long)a.B * a.Tq, (long long)a.H * a.dh, a.q_st, QROWS) ||
      make_map(&mx, a.k, (long long)a.B * a.Tk, a.dh, a.k_st, g.Tk16))
    return 1;
  // at least 118 KB so that a second CTA can never become co-resident on an SM (each CTA allocates all 512 TMEM columns)
  size_t smem = (size_t)g.stages * g.stage_bytes + BAR_BYTES + 1024;
  if (smem < 118 * 1024) smem = 118 * 1024;
  const int grid = a.B < sm_count() ? a.B : sm_count();
  TCAVP_CUDA(cudaFuncSetAttribute(attn_xt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_xt_kernel<<<grid, THREADS, smem, stream>>>(mq, mx, g);
  return check_launch("attn_xt_kernel");
}

}  // namespace tcavp
