// Tensor-core weight gradient of a bf16 linear map, no transposed copies:
//     out[n][k] += sum_m  Y[m][n] * s[m] * X[m][k]        (Y = dL/d(output) [M, N], X = layer input [M, K], s optional)
// (autograd of F.linear / peft lora.Linear in the reference loop, scripts/im_kim_train_GRN.py:1039).
//
// Both operands stay in their row-major [M, *] activation layout: the contraction runs over ROWS, so the mma.sync fragments are
// gathered with ldmatrix.trans (A = Y^T tile, B = X tile).  A CTA owns a 128 (n) x BK (k) output tile and one slice of the M
// rows (grid.z), streams 32-row slabs of Y and X through a cp.async double buffer, and adds its fp32 partial tile to the
// zero-initialised output with atomics.  The per-row factor s (the RMSNorm rstd of the folded-norm LoRA form) is applied to
// the narrow X slab on its way into shared memory.  Rank-r LoRA gradients (K = targets x r <= 32) are bound by the single
// HBM pass over Y; wide layers replace transpose + transpose + GEMM.
#include "common.cuh"

namespace tcavp {
namespace dwtc {

constexpr int BN = 128;      // output rows (n) per CTA: 8 warps x 16
constexpr int MS = 32;       // contraction rows per slab
constexpr int PAD = 8;       // bf16 elements of smem row padding (conflict-free ldmatrix)
constexpr int THREADS = 256;

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// DROP: the rows of Y carry a per-target dropout mask (peft lora_dropout: target t = output column block [t r, (t + 1) r) sees
// keep_t(m N + n) ? Y[m][n] : 0 — include/tcavp.h: tcavp_lora_da_drop).  The masks are regenerated on the A fragments in registers, one
// masked copy of the fragment per target, so the masked activations never exist in memory.
struct DropSpec {
  const uint32_t* seed;
  uint32_t site[4], thresh[4];
  int r, n;
};

template <int BK, bool DROP>
__global__ void __launch_bounds__(THREADS) dw_tc_kernel(const __nv_bfloat16* __restrict__ Y, int ldy, const __nv_bfloat16* __restrict__ X, int ldx,
                                                        const float* __restrict__ scale, float* __restrict__ out, int ldo, long long M, int N, int K,
                                                        int m_per_block, DropSpec ds) {
  constexpr int LDY = BN + PAD, LDX = BK + PAD;
  constexpr int YV = MS * BN / 8 / THREADS;                 // 16-byte vectors of the Y slab per thread (2)
  constexpr int XV = (MS * BK / 8 + THREADS - 1) / THREADS;   // ... of the X slab (1 for BK <= 64, 2 for BK = 128)
  __shared__ __align__(16) __nv_bfloat16 sY[2][MS * LDY];
  __shared__ __align__(16) __nv_bfloat16 sX[2][MS * LDX];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN, k0 = blockIdx.x * BK;
  const long long m0 = (long long)blockIdx.z * m_per_block;
  long long m1 = m0 + m_per_block;
  if (m1 > M) m1 = M;
  const uint32_t sY_u = (uint32_t)__cvta_generic_to_shared(&sY[0][0]), sX_u = (uint32_t)__cvta_generic_to_shared(&sX[0][0]);

  auto load_y = [&](long long mb, int buf) {
#pragma unroll
    for (int i = 0; i < YV; ++i) {
      const int e = tid + i * THREADS;
      const int r = e / (BN / 8), c = (e % (BN / 8)) * 8;
      const bool ok = mb + r < m1 && n0 + c < N;
      cp_async16(sY_u + (uint32_t)((buf * MS * LDY + r * LDY + c) * 2), ok ? Y + (size_t)(mb + r) * ldy + n0 + c : Y, ok);
    }
  };
  // X slab: through registers when a row factor has to be applied, cp.async otherwise
  uint4 xr[XV];
  auto fetch_x = [&](long long mb, int buf) {
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      const int e = tid + i * THREADS;
      const int r = e / (BK / 8), c = (e % (BK / 8)) * 8;
      if (e >= MS * BK / 8) continue;
      const bool ok = mb + r < m1 && k0 + c < K;
      if (scale) {
        xr[i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          xr[i] = *reinterpret_cast<const uint4*>(X + (size_t)(mb + r) * ldx + k0 + c);
          const float s = __ldg(scale + mb + r);
          uint32_t* w = reinterpret_cast<uint32_t*>(&xr[i]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(w[q] << 16) * s, __uint_as_float(w[q] & 0xffff0000u) * s);
            w[q] = *reinterpret_cast<const uint32_t*>(&v);
          }
        }
      } else {
        cp_async16(sX_u + (uint32_t)((buf * MS * LDX + r * LDX + c) * 2), ok ? X + (size_t)(mb + r) * ldx + k0 + c : X, ok);
      }
    }
  };
  auto commit_x = [&](int buf) {
    if (!scale) return;
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      const int e = tid + i * THREADS;
      const int r = e / (BK / 8), c = (e % (BK / 8)) * 8;
      if (e < MS * BK / 8) *reinterpret_cast<uint4*>(&sX[buf][r * LDX + c]) = xr[i];
    }
  };

  float acc[BK / 8][4];
#pragma unroll
  for (int i = 0; i < BK / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // ldmatrix.trans lane offsets.  A (Y^T tile 16 n x 16 m from sY[m][n]): matrices (n lo/hi) x (m lo/hi) in fragment order a0..a3
  const int a_m = (lane & 7) + ((lane >> 4) << 3), a_n = ((lane >> 3) & 1) * 8;
  // B (X tile 16 m x 16 k from sX[m][k]): registers = (k lo: m lo, m hi), (k hi: m lo, m hi)
  const int b_m = (lane & 7) + (((lane >> 3) & 1) << 3), b_k = (lane >> 4) * 8;

  const int nslabs = (int)((m1 - m0 + MS - 1) / MS);
  if (nslabs > 0) {
    load_y(m0, 0);
    fetch_x(m0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    commit_x(0);
  }
  for (int s = 0; s < nslabs; ++s) {
    const int buf = s & 1;
    if (s + 1 < nslabs) {
      load_y(m0 + (long long)(s + 1) * MS, buf ^ 1);
      fetch_x(m0 + (long long)(s + 1) * MS, buf ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int ms = 0; ms < MS / 16; ++ms) {
      uint32_t af[4];
      ldsm_x4_t(sY_u + (uint32_t)((buf * MS * LDY + (ms * 16 + a_m) * LDY + warp * 16 + a_n) * 2), af[0], af[1], af[2], af[3]);
      if (DROP) {
        uint32_t bq[BK / 16][4];
#pragma unroll
        for (int kp = 0; kp < BK / 16; ++kp)
          ldsm_x4_t(sX_u + (uint32_t)((buf * MS * LDX + (ms * 16 + b_m) * LDX + kp * 16 + b_k) * 2), bq[kp][0], bq[kp][1], bq[kp][2], bq[kp][3]);
        // fragment element (register i, half e): n = n0 + 16 warp + g + 8 (i & 1),  m = slab row 16 ms + 2 t4 + e + 8 (i >> 1)
        const unsigned long long mrow = (unsigned long long)(m0 + (long long)s * MS + ms * 16 + (lane & 3) * 2);
        const unsigned int ncol = (unsigned int)(n0 + warp * 16 + (lane >> 2));
        int cur = -1;
        uint32_t mf[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < BK / 8; ++i) {
          if (k0 + i * 8 < K) {
            const int t = (k0 + i * 8) / ds.r;
            if (t != cur) {
              cur = t;
              const uint32_t key = drop_key(ds.seed, ds.site[t]), th = ds.thresh[t];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const unsigned long long idx = (mrow + (unsigned)((q >> 1) * 8)) * (unsigned long long)N + ncol + (unsigned)((q & 1) * 8);
                const uint32_t keep = (drop_keep(key, idx, th) ? 0x0000FFFFu : 0u) | (drop_keep(key, idx + (unsigned long long)N, th) ? 0xFFFF0000u : 0u);
                mf[q] = af[q] & keep;
              }
            }
            mma16816(acc[i], mf, bq[i >> 1][(i & 1) * 2], bq[i >> 1][(i & 1) * 2 + 1]);
          }
        }
      } else {
#pragma unroll
        for (int kp = 0; kp < BK / 16; ++kp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sX_u + (uint32_t)((buf * MS * LDX + (ms * 16 + b_m) * LDX + kp * 16 + b_k) * 2), b0, b1, b2, b3);
          mma16816(acc[2 * kp], af, b0, b1);
          mma16816(acc[2 * kp + 1], af, b2, b3);
        }
      }
    }
    __syncthreads();                 // everyone is done with `buf`; the scaled X slab of s+1 can be published into buf^1
    if (s + 1 < nslabs) commit_x(buf ^ 1);
  }
  const int g = lane >> 2, t4 = lane & 3;
  const int n_lo = n0 + warp * 16 + g, n_hi = n_lo + 8;
#pragma unroll
  for (int i = 0; i < BK / 8; ++i) {
    const int k = k0 + i * 8 + t4 * 2;
    if (k < K) {     // K % 8 == 0: k + 1 < K as well
      if (n_lo < N) {
        atomicAdd(out + (size_t)n_lo * ldo + k, acc[i][0]);
        atomicAdd(out + (size_t)n_lo * ldo + k + 1, acc[i][1]);
      }
      if (n_hi < N) {
        atomicAdd(out + (size_t)n_hi * ldo + k, acc[i][2]);
        atomicAdd(out + (size_t)n_hi * ldo + k + 1, acc[i][3]);
      }
    }
  }
}

}  // namespace dwtc

static int dw_tc_launch_impl(const void* Y, int ldy, const void* X, int ldx, const float* scale, float* out, int ldo, long long M, int N, int K,
                             const dwtc::DropSpec* ds, cudaStream_t stream) {
  if (N % 8 || K % 8 || ldy % 8 || ldx % 8 || reinterpret_cast<uintptr_t>(Y) % 16 || reinterpret_cast<uintptr_t>(X) % 16) return 1;
  const int bk = K <= 32 ? 32 : (K <= 64 || K % 128 ? 64 : 128);
  const int tn = (N + dwtc::BN - 1) / dwtc::BN, tk = (K + bk - 1) / bk;
  long long splits = ((long long)sm_count() * 3) / ((long long)tn * tk);
  if (splits < 1) splits = 1;
  long long mper = (M + splits - 1) / splits;
  mper = (mper + dwtc::MS - 1) / dwtc::MS * dwtc::MS;
  if (mper < 4 * dwtc::MS) mper = 4 * dwtc::MS;
  const long long nz = (M + mper - 1) / mper;
  if (nz > 65535 || tn > 65535) return 1;
  const dim3 grid(tk, tn, (unsigned)nz);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(Y);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(X);
  if (ds) {
    if (bk == 128) return 1;
    if (bk == 32) dwtc::dw_tc_kernel<32, true><<<grid, dwtc::THREADS, 0, stream>>>(y, ldy, x, ldx, scale, out, ldo, M, N, K, (int)mper, *ds);
    else dwtc::dw_tc_kernel<64, true><<<grid, dwtc::THREADS, 0, stream>>>(y, ldy, x, ldx, scale, out, ldo, M, N, K, (int)mper, *ds);
    return check_launch("dw_tc_drop_kernel");
  }
  dwtc::DropSpec none = {};
  if (bk == 32) dwtc::dw_tc_kernel<32, false><<<grid, dwtc::THREADS, 0, stream>>>(y, ldy, x, ldx, scale, out, ldo, M, N, K, (int)mper, none);
  else if (bk == 64) dwtc::dw_tc_kernel<64, false><<<grid, dwtc::THREADS, 0, stream>>>(y, ldy, x, ldx, scale, out, ldo, M, N, K, (int)mper, none);
  else dwtc::dw_tc_kernel<128, false><<<grid, dwtc::THREADS, 0, stream>>>(y, ldy, x, ldx, scale, out, ldo, M, N, K, (int)mper, none);
  return check_launch("dw_tc_kernel");
}

// Returns 1 when the operands are not covered (caller uses the FFMA kernel), <= 0 otherwise.
int dw_tc_launch(const void* Y, int ldy, const void* X, int ldx, const float* scale, float* out, int ldo, long long M, int N, int K,
                 cudaStream_t stream) {
  return dw_tc_launch_impl(Y, ldy, X, ldx, scale, out, ldo, M, N, K, nullptr, stream);
}

}  // namespace tcavp

// peft lora.Linear in train mode, gradient of lora_A (include/tcavp.h).
extern "C" int tcavp_lora_da_drop(const void* x, int ldx, const void* dT, int lddt, const float* row_scale, float* out, int ldo, long long M, int H,
                                  int r, int n_targets, const uint32_t* seed, const uint32_t* sites, const uint32_t* thresh, tcavp_stream_t stream) {
  using namespace tcavp;
  TCAVP_REQUIRE(M >= 0 && H > 0 && (r == 8 || r == 16) && n_targets >= 1 && n_targets <= 4 && ldo >= r * n_targets,
                "tcavp_lora_da_drop: bad shape (H=%d r=%d targets=%d)", H, r, n_targets);
  if (M == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && dT && out && seed && sites && thresh, "tcavp_lora_da_drop: null pointer");
  dwtc::DropSpec ds = {};
  ds.seed = seed; ds.r = r; ds.n = n_targets;
  for (int t = 0; t < n_targets; ++t) { ds.site[t] = sites[t]; ds.thresh[t] = thresh[t]; }
  const int rc = dw_tc_launch_impl(x, ldx, dT, lddt, row_scale, out, ldo, M, H, r * n_targets, &ds, reinterpret_cast<cudaStream_t>(stream));
  if (rc > 0) return fail_arg("tcavp_lora_da_drop: operands must be bf16 with 16-byte aligned rows (H %% 8, ldx %% 8, lddt %% 8)");
  return rc;
}
