// peft lora.Linear in train() mode (reference scripts/train.py:432-440, lora_dropout = 0.1): every target module (q_proj, k_proj,
// v_proj) computes lora_A(dropout(x)) with its OWN mask of the same normalised input.  Done literally that is one masked copy of the
// [rows, H] residual stream + one skinny GEMM per target in the forward pass, and again per target for each of the two gradients —
// about ten extra passes over the residual stream per decoder layer.  The masks are a pure function of (seed, site, element)
// (common.cuh: drop_keep), so these kernels regenerate them on the operand fragments in registers instead:
//   tcavp_lora_a_drop    forward:   T[m, j]  = sum_h keep_{t(j)}(m H + h) x[m, h] A[j, h]                (one pass over x, mma.sync)
//   tcavp_lora_dx_drop   backward:  dx[m, h] += sum_t keep_t(m H + h) sum_{j in t} dT[m, j] A[j, h]        (one pass over dx)
//   tcavp_lora_da_drop   backward:  dA[h, j] += sum_m s[m] keep_{t(j)}(m H + h) x[m, h] dT[m, j]           (dw_tc.cu, masked fragments)
// with t(j) = j / r.  The masked activations never exist in memory.
#include "common.cuh"

namespace tcavp {
namespace ld {

constexpr int TM = 64;            // rows per CTA (4 warps x 16)
constexpr int KC = 64;            // contraction columns per stage
constexpr int LDS = KC + 8;       // padded shared-memory row (bf16 elements): conflict-free ldmatrix / fragment loads
constexpr int THREADS = 128;
constexpr int MAX_NL = 64;        // targets x r

struct Spec {
  const uint32_t* seed;
  uint32_t site[4], thresh[4];
  int r, n;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Forward.  A CTA owns 64 rows; x and A stream through a cp.async double buffer in 64-column stages; per 16-column step a warp loads
// its 16 x 16 fragment of x once and multiplies one MASKED copy of it per target with that target's r / 8 n-tiles of A.
__global__ void __launch_bounds__(THREADS) lora_a_drop_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ A, int lda,
                                                              __nv_bfloat16* __restrict__ out, int ldo, long long M, int H, Spec sp) {
  __shared__ __align__(16) __nv_bfloat16 sx[2][TM * LDS];
  __shared__ __align__(16) __nv_bfloat16 sa[2][MAX_NL * LDS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const long long m0 = (long long)blockIdx.x * TM;
  const int NL = sp.r * sp.n, ntiles = NL >> 3, tiles_per_t = sp.r >> 3;
  const uint32_t sx_u = (uint32_t)__cvta_generic_to_shared(&sx[0][0]), sa_u = (uint32_t)__cvta_generic_to_shared(&sa[0][0]);
  auto load = [&](int kc, int buf) {
#pragma unroll
    for (int i = 0; i < TM * (KC / 8) / THREADS; ++i) {
      const int e = tid + i * THREADS, r = e >> 3, c = (e & 7) * 8;
      const bool ok = m0 + r < M;
      cp_async16(sx_u + (uint32_t)((buf * TM * LDS + r * LDS + c) * 2), ok ? x + (size_t)(m0 + r) * ldx + kc * KC + c : x, ok);
    }
    for (int e = tid; e < NL * (KC / 8); e += THREADS) {
      const int r = e >> 3, c = (e & 7) * 8;
      cp_async16(sa_u + (uint32_t)((buf * MAX_NL * LDS + r * LDS + c) * 2), A + (size_t)r * lda + kc * KC + c, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t key[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) key[t] = t < sp.n ? drop_key(sp.seed, sp.site[t]) : 0u;
  float acc[MAX_NL / 8][4];
#pragma unroll
  for (int i = 0; i < MAX_NL / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;
  // this thread's fragment rows / first column: registers (0, 1) = rows (g, g + 8) of columns 2 t4 + {0, 1}; (2, 3) = the same rows, + 8 columns
  const unsigned long long row_lo = (unsigned long long)(m0 + warp * 16 + g) * (unsigned long long)H;
  const unsigned long long row_hi = row_lo + 8ull * (unsigned long long)H;
  const int nk = H / KC;
  load(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) {
      load(kc + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < KC / 16; ++ks) {
      uint32_t af[4];
      ldsm_x4(sx_u + (uint32_t)((buf * TM * LDS + (warp * 16 + a_row) * LDS + ks * 16 + a_col) * 2), af[0], af[1], af[2], af[3]);
      const unsigned int h0 = (unsigned int)(kc * KC + ks * 16 + t4 * 2);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < sp.n) {
          const uint32_t th = sp.thresh[t];
          uint32_t mf[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned long long idx = ((q & 1) ? row_hi : row_lo) + h0 + (unsigned)((q >> 1) * 8);
            const uint32_t keep = (drop_keep(key[t], idx, th) ? 0x0000FFFFu : 0u) | (drop_keep(key[t], idx + 1, th) ? 0xFFFF0000u : 0u);
            mf[q] = af[q] & keep;
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < tiles_per_t) {
              const int nt = t * tiles_per_t + j;
              const __nv_bfloat16* bp = &sa[buf][(nt * 8 + g) * LDS + ks * 16 + t4 * 2];
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(bp), b1 = *reinterpret_cast<const uint32_t*>(bp + 8);
              // acc index must be a compile-time constant for the accumulators to stay in registers
              if (tiles_per_t == 1) mma16816(acc[t], mf, b0, b1);
              else mma16816(acc[2 * t + j], mf, b0, b1);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  const long long r_lo = m0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
  for (int i = 0; i < MAX_NL / 8; ++i) {
    if (i < ntiles) {
      const int c = i * 8 + t4 * 2;
      if (r_lo < M) *reinterpret_cast<__nv_bfloat162*>(out + (size_t)r_lo * ldo + c) = __floats2bfloat162_rn(acc[i][0], acc[i][1]);
      if (r_hi < M) *reinterpret_cast<__nv_bfloat162*>(out + (size_t)r_hi * ldo + c) = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
    }
  }
}

// Backward w.r.t. the input.  A thread owns 8 consecutive columns of RW = 4 rows: per target the rank-r product dT_t . A_t of those
// 32 elements is formed in registers (A rows read once per four rows, dT broadcast loads), masked, and added to dx in place.
// Measured (1.1 TB/s on the 768 shape, 0.74 TB/s at 7B): instruction-issue bound, about half mask hash and half unpack + FMA.  Tried and
// not faster: two rows per thread at three CTAs per SM (occupancy is not the limit), and a keep-bit bitmap handed over from the forward
// kernel so that this kernel and the lora_A-gradient kernel read bits instead of hashing (bit-exact, 0.5 % on the 7B step, -2 % on the
// 768 step: the forward kernel pays for the bit packing what the gradient kernels save).
constexpr int RW = 4;
__global__ void __launch_bounds__(256, 2) lora_dx_drop_kernel(const __nv_bfloat16* __restrict__ dT, int lddt, const __nv_bfloat16* __restrict__ A, int lda,
                                                           __nv_bfloat16* __restrict__ dx, int lddx, long long M, int H, Spec sp) {
  const int hv = H >> 3;
  const long long groups = (M + RW - 1) / RW;
  uint32_t key[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) key[t] = t < sp.n ? drop_key(sp.seed, sp.site[t]) : 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < groups * hv; i += (long long)gridDim.x * blockDim.x) {
    const long long m0 = (i / hv) * RW;
    const int h = (int)(i % hv) * 8;
    float acc[RW][8];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
#pragma unroll 1
    for (int t = 0; t < sp.n; ++t) {
      float tmp[RW][8];
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) tmp[r][e] = 0.f;
#pragma unroll 1
      for (int j0 = 0; j0 < sp.r; j0 += 8) {
        uint4 dv[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const long long m = m0 + r < M ? m0 + r : M - 1;
          dv[r] = __ldg(reinterpret_cast<const uint4*>(dT + (size_t)m * lddt + t * sp.r + j0));
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint4 av = __ldg(reinterpret_cast<const uint4*>(A + (size_t)(t * sp.r + j0 + jj) * lda + h));
          const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
          float a[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            a[2 * e] = __uint_as_float(aw[e] << 16);
            a[2 * e + 1] = __uint_as_float(aw[e] & 0xffff0000u);
          }
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            const uint32_t dw[4] = {dv[r].x, dv[r].y, dv[r].z, dv[r].w};
            const uint32_t w = dw[jj >> 1];
            const float d = (jj & 1) ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
#pragma unroll
            for (int e = 0; e < 8; ++e) tmp[r][e] = fmaf(d, a[e], tmp[r][e]);
          }
        }
      }
      const uint32_t th = t == 0 ? sp.thresh[0] : t == 1 ? sp.thresh[1] : t == 2 ? sp.thresh[2] : sp.thresh[3];
      const uint32_t kt = t == 0 ? key[0] : t == 1 ? key[1] : t == 2 ? key[2] : key[3];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const unsigned long long base = (unsigned long long)(m0 + r) * (unsigned long long)H + (unsigned)h;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] += drop_keep(kt, base + e, th) ? tmp[r][e] : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (m0 + r < M) {
        uint4* p = reinterpret_cast<uint4*>(dx + (size_t)(m0 + r) * lddx + h);
        uint4 u = *p;
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(w[e] << 16) + acc[r][2 * e], __uint_as_float(w[e] & 0xffff0000u) + acc[r][2 * e + 1]);
          w[e] = *reinterpret_cast<const uint32_t*>(&v);
        }
        *p = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

static int fill_spec(Spec& sp, int r, int n, const uint32_t* seed, const uint32_t* sites, const uint32_t* thresh) {
  sp.seed = seed; sp.r = r; sp.n = n;
  for (int t = 0; t < 4; ++t) {
    sp.site[t] = t < n ? sites[t] : 0u;
    sp.thresh[t] = t < n ? thresh[t] : 0u;
  }
  return 0;
}

}  // namespace ld
}  // namespace tcavp

using namespace tcavp;

extern "C" int tcavp_lora_a_drop(const void* x, int ldx, const void* A, int lda, void* out, int ldo, long long M, int H, int r, int n_targets,
                                 const uint32_t* seed, const uint32_t* sites, const uint32_t* thresh, tcavp_stream_t stream) {
  TCAVP_REQUIRE(M >= 0 && H > 0 && H % ld::KC == 0 && (r == 8 || r == 16) && n_targets >= 1 && n_targets <= 4 && ldx >= H && lda >= H &&
                    ldo >= r * n_targets,
                "tcavp_lora_a_drop: bad shape (H=%d must be a multiple of 64, r=%d in {8, 16}, targets=%d <= 4)", H, r, n_targets);
  if (M == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && A && out && seed && sites && thresh, "tcavp_lora_a_drop: null pointer");
  TCAVP_REQUIRE(ldx % 8 == 0 && lda % 8 == 0 && ldo % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(A) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(out) % 4 == 0,
                "tcavp_lora_a_drop: x / A rows must be 16-byte aligned (ldx %% 8, lda %% 8), out 4-byte aligned with an even ldo");
  ld::Spec sp;
  ld::fill_spec(sp, r, n_targets, seed, sites, thresh);
  const long long grid = (M + ld::TM - 1) / ld::TM;
  TCAVP_REQUIRE(grid <= 0x7fffffffLL, "tcavp_lora_a_drop: too many rows");
  ld::lora_a_drop_kernel<<<(unsigned)grid, ld::THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<__nv_bfloat16*>(out), ldo, M, H, sp);
  return check_launch("lora_a_drop_kernel");
}

extern "C" int tcavp_lora_dx_drop(const void* dT, int lddt, const void* A, int lda, void* dx, int lddx, long long M, int H, int r, int n_targets,
                                  const uint32_t* seed, const uint32_t* sites, const uint32_t* thresh, tcavp_stream_t stream) {
  TCAVP_REQUIRE(M >= 0 && H > 0 && H % 8 == 0 && (r == 8 || r == 16) && n_targets >= 1 && n_targets <= 4 && lddt >= r * n_targets && lda >= H &&
                    lddx >= H,
                "tcavp_lora_dx_drop: bad shape (H=%d r=%d targets=%d)", H, r, n_targets);
  if (M == 0) return TCAVP_OK;
  TCAVP_REQUIRE(dT && A && dx && seed && sites && thresh, "tcavp_lora_dx_drop: null pointer");
  TCAVP_REQUIRE(lddt % 8 == 0 && lda % 8 == 0 && lddx % 8 == 0 && reinterpret_cast<uintptr_t>(dT) % 16 == 0 && reinterpret_cast<uintptr_t>(A) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(dx) % 16 == 0,
                "tcavp_lora_dx_drop: rows of dT / A / dx must be 16-byte aligned");
  ld::Spec sp;
  ld::fill_spec(sp, r, n_targets, seed, sites, thresh);
  const long long work = ((M + ld::RW - 1) / ld::RW) * (H / 8);
  long long grid = (work + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  ld::lora_dx_drop_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dT), lddt, reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<__nv_bfloat16*>(dx), lddx, M, H, sp);
  return check_launch("lora_dx_drop_kernel");
}
