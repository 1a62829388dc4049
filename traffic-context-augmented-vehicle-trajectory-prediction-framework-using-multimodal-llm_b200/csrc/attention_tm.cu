// tcgen05 flash attention for the LLM self-attention (HF:251-289 with the causal + key-padding mask of HF:399; GQA):
// bf16, causal, head_dim 64 / 128, 128 <= L <= 256 (the fused sequence is 16 image + 128 text tokens = 144).
//
// Persistent, warp-specialised, one CTA per SM looping over (scene, head) items:
//   warp 0      TMA producer: Q, K, V tiles of the item ([L16 rows] x 64-column boxes, 128-byte swizzle) straight out of the packed
//               [rows, (nh + 2 nkv) dh] QKV activation into a 2-stage shared-memory ring (the next item lands under this item's math)
//   warp 1      one thread issues tcgen05.mma:  S = Q K^T (both operands K-major from shared memory, fp32 accumulators in TMEM) and
//               O = P V  (A = P read from TENSOR MEMORY, B = V as an MN-major shared-memory operand: no transposed copy of V)
//   warps 4-7   softmax of tile A, one thread per query row: tcgen05.ld the row's scores, mask, exp2, row sum, and write the bf16
//               probabilities back over the scores with tcgen05.st (P aliases S in TMEM); later drain O, scale by 1 / row sum, store
//   warps 8-11  the same for tile B
// L = 144 does not fit a 128-row UMMA tile, and padding to 256 would waste a whole second tile.  The causal structure gives a cheap
// split instead: tile A = the LAST 128 query rows [L16-128, L16) against all L16 keys, tile B = the first L16-128 rows (16 for L = 144),
// which can only see the first L16-128 keys — an M = 128, N = 16 MMA with 16 live rows.  TMEM: S_A (L16 cols) + S_B + O_A + O_B (dh each).
// Scores / probabilities never leave the SM; Q, K, V are read once and O written once (the HBM floor of the op).
#include <cuda.h>
#include <mutex>
#include "common.cuh"

namespace tcavp {
namespace tm {

constexpr int THREADS = 384;      // 12 warps: 0 TMA, 1 MMA (+ TMEM owner), 2-3 idle, 4-7 softmax A, 8-11 softmax B
constexpr int BOXC = 64;          // bf16 columns per TMA box = 128 bytes = one swizzle row

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

// Shared-memory operand descriptors (cute::UMMA::SmemDescriptor, SWIZZLE_128B, version 1).
//   K-major  (Q, K):  rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); the k-step advances the start address by 32 bytes.
//   MN-major (V as the B operand of P.V, N = head dim contiguous, K = keys): canonical ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in
//            elements — 64 contiguous head-dim elements (128 bytes) per key row, 8-key groups 1024 bytes apart (SBO), the next 64
//            head-dim elements LBO bytes away (the second 64-column TMA box of a 128-wide head).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D fp32, A / B bf16, A K-major, B K-major or MN-major (bit 16).
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
  int B, H, Hkv, L, L16, NB;          // NB = L16 - 128: rows / keys of tile B (0: no tile B)
  int stages;
  uint32_t box_bytes, stage_bytes;    // one [L16 x 64] box; Q + K + V of one item
  uint32_t sb_col, oa_col, ob_col;    // TMEM columns of S_B, O_A, O_B (S_A sits at 0)
  float sl2;                          // softmax scale * log2(e)
  const int32_t* key_mask;
  __nv_bfloat16* out; long long o_sb, o_st;
};

// One softmax thread = one query row of a tile.  `q`: the row's query index, `ncols`: key columns of the tile (multiple of 16),
// `s_col` / `o_col`: TMEM columns of the tile's scores and output.
template <int DH>
__device__ __forceinline__ void softmax_tile(const Geo& g, uint32_t tmem_lane_base, uint32_t s_col, uint32_t o_col, int q, int ncols, int warp_kmax,
                                             const uint32_t* s_mask, uint32_t s_full, uint32_t p_ready, uint32_t o_full, uint32_t o_free,
                                             uint32_t phase, int b, int h, int lane) {
  // keys this warp can ever see (causal): units of 16 columns, warp-uniform
  const int nu = min(ncols, (warp_kmax + 16) & ~15) >> 4;
  const int nu_all = ncols >> 4;
  mbar_wait(s_full, phase);
  tc_fence_after();
  const uint32_t srow = tmem_lane_base + s_col;
  float m = -INFINITY;
  for (int u = 0; u < nu; ++u) {
    uint32_t r[16];
    tmem_ld16(srow + u * 16, r);
    const uint32_t bits = s_mask[u >> 1] >> ((u & 1) * 16);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int j = u * 16 + c;
      const bool ok = ((bits >> c) & 1u) && j <= q;
      m = fmaxf(m, ok ? __uint_as_float(r[c]) : -INFINITY);
    }
  }
  const float mref = m == -INFINITY ? 0.f : m * g.sl2;
  float l = 0.f;
  for (int u = 0; u < nu; ++u) {
    uint32_t r[16], pk[8];
    tmem_ld16(srow + u * 16, r);
    const uint32_t bits = s_mask[u >> 1] >> ((u & 1) * 16);
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const int j = u * 16 + c;
      const float p0 = (((bits >> c) & 1u) && j <= q) ? ex2(fmaf(__uint_as_float(r[c]), g.sl2, -mref)) : 0.f;
      const float p1 = (((bits >> (c + 1)) & 1u) && j + 1 <= q) ? ex2(fmaf(__uint_as_float(r[c + 1]), g.sl2, -mref)) : 0.f;
      const uint32_t w = pack2(p0, p1);
      // the row sum is taken over the ROUNDED probabilities the P.V product multiplies with
      l += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
      pk[c >> 1] = w;
    }
    tmem_st8(srow + u * 8, pk);      // P (bf16 pairs) overwrites score columns that were already consumed: 8u + 8 <= 16u + 16
  }
  {
    const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (int u = nu; u < nu_all; ++u) tmem_st8(srow + u * 8, z);    // keys beyond the causal bound of this warp's rows
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(p_ready);
  // ---- O = P V is on its way: drain it, normalise, store ----
  mbar_wait(o_full, phase);
  tc_fence_after();
  const float inv = l > 0.f ? 1.f / l : 0.f;
  __nv_bfloat16* op = g.out + (size_t)b * g.o_sb + (size_t)q * g.o_st + (size_t)h * DH;
  const bool live = q < g.L;
#pragma unroll
  for (int c0 = 0; c0 < DH; c0 += 16) {
    uint32_t r[16], w[8];
    tmem_ld16(tmem_lane_base + o_col + c0, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = pack2(__uint_as_float(r[2 * e]) * inv, __uint_as_float(r[2 * e + 1]) * inv);
    if (live) stg256(op + c0, w);
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(o_free);
}

template <int DH>
__global__ void __launch_bounds__(THREADS, 1)
attn_tm_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v, Geo g) {
  constexpr int NBOX = DH / BOXC;      // 64-column boxes per operand
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t s_mask[12][8];   // per warp: key-valid bits of the current item
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + g.stages * g.stage_bytes;
  const uint32_t qkv_full = bars, qkv_empty = bars + 16;             // [2] each
  const uint32_t sa_full = bars + 32, sb_full = bars + 40, pa_ready = bars + 48, pb_ready = bars + 56;
  const uint32_t oa_full = bars + 64, ob_full = bars + 72, oa_free = bars + 80, ob_free = bars + 88;
  const uint32_t tmem_slot = bars + 96;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbw = (g.NB + 31) >> 5;                                   // softmax-B warps with live rows
  const int n_items = g.B * g.H;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_v)) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(qkv_full + 8 * s, 1);
      mbar_init(qkv_empty + 8 * s, 1);
    }
    mbar_init(sa_full, 1);
    mbar_init(sb_full, 1);
    mbar_init(pa_ready, 4);
    mbar_init(pb_ready, nbw > 0 ? nbw : 1);
    mbar_init(oa_full, 1);
    mbar_init(ob_full, 1);
    mbar_init(oa_free, 4);
    mbar_init(ob_free, nbw > 0 ? nbw : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const int s = n % g.stages;
        const uint32_t use = (uint32_t)(n / g.stages);
        const int b = it / g.H, h = it % g.H, hk = h / (g.H / g.Hkv);
        mbar_wait(qkv_empty + 8 * s, (use & 1u) ^ 1u);
        mbar_expect_tx(qkv_full + 8 * s, g.stage_bytes);
        const uint32_t st = smem_base + s * g.stage_bytes;
#pragma unroll
        for (int c = 0; c < NBOX; ++c) {
          tma_load_2d(st + c * g.box_bytes, &tma_q, qkv_full + 8 * s, h * DH + c * BOXC, b * g.L);
          tma_load_2d(st + (NBOX + c) * g.box_bytes, &tma_k, qkv_full + 8 * s, hk * DH + c * BOXC, b * g.L);
          tma_load_2d(st + (2 * NBOX + c) * g.box_bytes, &tma_v, qkv_full + 8 * s, hk * DH + c * BOXC, b * g.L);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_sa = idesc_f16(128, g.L16, 0), id_sb = idesc_f16(128, g.NB > 0 ? g.NB : 16, 0), id_pv = idesc_f16(128, DH, 1);
      const int qa0 = g.L16 - 128;      // first query row of tile A
      int n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const int s = n % g.stages;
        const uint32_t use = (uint32_t)(n / g.stages), ph = (uint32_t)(n & 1);
        const uint32_t sq = smem_base + s * g.stage_bytes, sk = sq + NBOX * g.box_bytes, sv = sk + NBOX * g.box_bytes;
        mbar_wait(qkv_full + 8 * s, use & 1u);
        tc_fence_after();
        // S_A = Q[qa0 .. qa0+128) . K^T over all L16 keys.  The previous item's P.V products (issued by this thread, executed in
        // order) have consumed the columns these accumulators overwrite.
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint32_t boff = (uint32_t)(k >> 2) * g.box_bytes + (uint32_t)(k & 3) * 32u;
          umma_ss(tmem_base, desc_kmajor(sq + boff + (uint32_t)qa0 * 128u), desc_kmajor(sk + boff), id_sa, k != 0);
        }
        umma_commit(sa_full);
        if (g.NB > 0) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            const uint32_t boff = (uint32_t)(k >> 2) * g.box_bytes + (uint32_t)(k & 3) * 32u;
            umma_ss(tmem_base + g.sb_col, desc_kmajor(sq + boff), desc_kmajor(sk + boff), id_sb, k != 0);
          }
          umma_commit(sb_full);
          // O_B = P_B . V[0 .. NB)
          mbar_wait(pb_ready, ph);
          mbar_wait(ob_free, ph ^ 1u);
          tc_fence_after();
          for (int kk = 0; kk < g.NB / 16; ++kk)
            umma_ts(tmem_base + g.ob_col, tmem_base + g.sb_col + kk * 8, desc_mnmajor(sv + (uint32_t)kk * 2048u, g.box_bytes), id_pv, kk != 0);
          umma_commit(ob_full);
        }
        // O_A = P_A . V
        mbar_wait(pa_ready, ph);
        mbar_wait(oa_free, ph ^ 1u);
        tc_fence_after();
        for (int kk = 0; kk < g.L16 / 16; ++kk)
          umma_ts(tmem_base + g.oa_col, tmem_base + kk * 8, desc_mnmajor(sv + (uint32_t)kk * 2048u, g.box_bytes), id_pv, kk != 0);
        umma_commit(oa_full);
        umma_commit(qkv_empty + 8 * s);      // every MMA that reads this stage has retired when this arrives
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output warps =====================
    const bool tile_a = warp < 8;
    const int wq = warp & 3;                                  // TMEM lane quarter of this warp
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int qa0 = g.L16 - 128;
    if (tile_a || wq < nbw) {
      int n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const int b = it / g.H, h = it % g.H;
        // key-valid bits of this scene (HF:399: padding keys are masked for every query)
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const int j = w * 32 + lane;
          const bool ok = j < g.L && (!g.key_mask || g.key_mask[(size_t)b * g.L + j] != 0);
          const uint32_t bits = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) s_mask[warp][w] = bits;
        }
        __syncwarp();
        const uint32_t ph = (uint32_t)(n & 1);
        if (tile_a) {
          const int q = qa0 + wq * 32 + lane;
          softmax_tile<DH>(g, lane_base, 0u, g.oa_col, q, g.L16, qa0 + wq * 32 + 31, s_mask[warp], sa_full, pa_ready, oa_full, oa_free, ph, b, h, lane);
        } else {
          const int q = wq * 32 + lane;
          // rows >= NB of tile B are not part of the problem: q is pushed past L so nothing is stored for them
          softmax_tile<DH>(g, lane_base, g.sb_col, g.ob_col, q < g.NB ? q : g.L + q, g.NB, min(wq * 32 + 31, g.NB - 1), s_mask[warp], sb_full,
                           pb_ready, ob_full, ob_free, ph, b, h, lane);
        }
        __syncwarp();
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
static int make_map(CUtensorMap* map, const void* base, long long rows, int cols, long long ld, int box_rows) {
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

}  // namespace tm

// Returns 1 when the shape is not covered (caller falls back to the mma.sync flash kernel), <= 0 otherwise.
int attention_tm_launch(const tcavp_attn_args& a, cudaStream_t stream) {
  using namespace tm;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TCAVP_ATTN_TCGEN05");      // 0: keep the mma.sync kernel (A/B runs)
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return 1;
  if (a.dtype != TCAVP_BF16 || !a.causal || a.Tq != a.Tk || a.drop_thresh != 0) return 1;
  if (!(a.dh == 64 || a.dh == 128) || a.Tq < 128 || a.Tq > 256 || a.B < 1) return 1;
  const int L = a.Tq, L16 = (L + 15) / 16 * 16, NB = L16 - 128;
  // rows of consecutive scenes must be contiguous (one 2-D tensor map per operand) and TMA-addressable
  if (a.q_sb != (long long)L * a.q_st || a.k_sb != (long long)L * a.k_st || a.v_sb != (long long)L * a.v_st) return 1;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v) || reinterpret_cast<uintptr_t>(a.out) % 32 || a.q_st % 8 || a.k_st % 8 || a.v_st % 8 || a.o_st % 16 ||
      a.o_sb % 16)
    return 1;
  Geo g;
  g.B = a.B; g.H = a.H; g.Hkv = a.Hkv; g.L = L; g.L16 = L16; g.NB = NB;
  g.box_bytes = (uint32_t)L16 * 128u;
  g.stage_bytes = 3u * (uint32_t)(a.dh / BOXC) * g.box_bytes;
  g.stages = 2u * g.stage_bytes + 2048u <= 225u * 1024u ? 2 : 1;
  if ((size_t)g.stages * g.stage_bytes + 2048 > 225 * 1024) return 1;
  g.sb_col = (uint32_t)((L16 + 31) / 32 * 32);
  g.oa_col = g.sb_col + (uint32_t)(NB > 0 ? (NB + 31) / 32 * 32 : 0);
  g.ob_col = g.oa_col + (uint32_t)a.dh;
  if (g.ob_col + (uint32_t)(NB > 0 ? a.dh : 0) > 512u) return 1;
  g.sl2 = a.scale * 1.4426950408889634f;
  g.key_mask = a.key_mask;
  g.out = reinterpret_cast<__nv_bfloat16*>(a.out); g.o_sb = a.o_sb; g.o_st = a.o_st;
  CUtensorMap mq, mk, mv;
  const long long rows = (long long)a.B * L;
  if (make_map(&mq, a.q, rows, a.H * a.dh, a.q_st, L16) || make_map(&mk, a.k, rows, a.Hkv * a.dh, a.k_st, L16) ||
      make_map(&mv, a.v, rows, a.Hkv * a.dh, a.v_st, L16))
    return 1;
  // at least 118 KB so that a second CTA can never become co-resident on an SM (each CTA allocates all 512 TMEM columns)
  size_t smem = (size_t)g.stages * g.stage_bytes + 256 + 1024;
  if (smem < 118 * 1024) smem = 118 * 1024;
  const int items = a.B * a.H;
  const int grid = items < sm_count() ? items : sm_count();
  if (a.dh == 64) {
    TCAVP_CUDA(cudaFuncSetAttribute(attn_tm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_tm_kernel<64><<<grid, THREADS, smem, stream>>>(mq, mk, mv, g);
  } else {
    TCAVP_CUDA(cudaFuncSetAttribute(attn_tm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_tm_kernel<128><<<grid, THREADS, smem, stream>>>(mq, mk, mv, g);
  }
  return check_launch("attn_tm_kernel");
}

}  // namespace tcavp
