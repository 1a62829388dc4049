// Temporal (LTSF) kernels: token_proj + NLinear encoder + positional add, NLinear decoder (+lane adjust),
// and the fusion head (LayerNorm -> MLP -> out_proj -> last-position residual) fused with the
// de-normalised loss / ADE / FDE reduction.  All fp32 math; weights are pre-permuted at pack time so the
// channel index is the fastest-varying one and every global access is coalesced over channels.
#include "common.cuh"

namespace tcavp {

constexpr int MAX_T_IN = 64;

// reference scripts/train.py:837 (token_proj), 701-716 (NLinear encoder), 839 (+pos_encoding)
//   we : [T_in(t), T_in(s), C]   be, pos : [T_in, C]   wt : [C, F]   enc : (B, T_in, C)
__global__ void __launch_bounds__(256) ltsf_encode_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                          const float* __restrict__ bt, const float* __restrict__ we,
                                                          const float* __restrict__ be, const float* __restrict__ pos,
                                                          void* __restrict__ enc, int out_dtype, int B, int F, int C, int T) {
  const long long total = (long long)B * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / C;
    float xp[MAX_T_IN];
#pragma unroll 1
    for (int s = 0; s < T; ++s) {
      float v = __ldg(bt + c);
      for (int f = 0; f < F; ++f) v = fmaf(__ldg(wt + c * F + f), __ldg(x + ((size_t)b * F + f) * T + s), v);
      xp[s] = v;
    }
    const float last = xp[T - 1];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float acc = __ldg(be + (size_t)t * C + c);
      for (int s = 0; s < T; ++s) acc = fmaf(__ldg(we + ((size_t)t * T + s) * C + c), xp[s] - last, acc);
      acc += last + __ldg(pos + (size_t)t * C + c);
      store_from_f(enc, ((size_t)b * T + t) * C + c, out_dtype, acc);
    }
  }
}

// reference scripts/train.py:769-785.  wd : [T_out, T_in, C], bd : [T_out, C]; enc (B,T_in,C); dec (B,T_out,C)
__global__ void __launch_bounds__(256) nlinear_decode_kernel(const void* __restrict__ enc, int enc_dtype, const float* __restrict__ wd,
                                                             const float* __restrict__ bd, const void* __restrict__ adj, int adj_dtype,
                                                             void* __restrict__ dec, int out_dtype, int B, int C, int T, int To) {
  const long long total = (long long)B * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / C;
    float e[MAX_T_IN];
#pragma unroll 1
    for (int s = 0; s < T; ++s) e[s] = load_as_f(enc, ((size_t)b * T + s) * C + c, enc_dtype);
    const float last = e[T - 1];
#pragma unroll 1
    for (int t = 0; t < To; ++t) {
      float acc = __ldg(bd + (size_t)t * C + c);
      for (int s = 0; s < T; ++s) acc = fmaf(__ldg(wd + ((size_t)t * T + s) * C + c), e[s] - last, acc);
      acc += last;
      const size_t o = ((size_t)b * To + t) * C + c;
      if (adj) acc += load_as_f(adj, o, adj_dtype);
      store_from_f(dec, o, out_dtype, acc);
    }
  }
}

// Register-tiled variants for a compile-time T_in: a thread owns one channel of NB scenes, so the history (and its projection)
// lives in registers and every weight load (coalesced over channels, L1/L2 resident) feeds NB FMAs instead of one.
constexpr int NB = 4;

template <int T>
__global__ void __launch_bounds__(256) ltsf_encode_tiled_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                                const float* __restrict__ bt, const float* __restrict__ we,
                                                                const float* __restrict__ be, const float* __restrict__ pos,
                                                                void* __restrict__ enc, int out_dtype, int B, int F, int C) {
  const int c = threadIdx.x % C;
  const int slot = threadIdx.x / C, slots = blockDim.x / C;
  const long long b0 = ((long long)blockIdx.x * slots + slot) * NB;
  if (b0 >= B) return;
  float xp[NB][T];
  const float btc = __ldg(bt + c);
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const long long b = b0 + j < B ? b0 + j : B - 1;
#pragma unroll
    for (int s = 0; s < T; ++s) xp[j][s] = btc;
    for (int f = 0; f < F; ++f) {
      const float w = __ldg(wt + c * F + f);
#pragma unroll
      for (int s = 0; s < T; ++s) xp[j][s] = fmaf(w, __ldg(x + ((size_t)b * F + f) * T + s), xp[j][s]);
    }
  }
  float last[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    last[j] = xp[j][T - 1];
#pragma unroll
    for (int s = 0; s < T; ++s) xp[j][s] -= last[j];
  }
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const float base = __ldg(be + (size_t)t * C + c) + __ldg(pos + (size_t)t * C + c);
    float acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = base + last[j];
#pragma unroll
    for (int s = 0; s < T; ++s) {
      const float w = __ldg(we + ((size_t)t * T + s) * C + c);
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] = fmaf(w, xp[j][s], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (b0 + j < B) store_from_f(enc, ((size_t)(b0 + j) * T + t) * C + c, out_dtype, acc[j]);
  }
}

// (two scenes per thread and three resident CTAs per SM: with four scenes the 115 registers per thread left 16 warps per SM, and the kernel was
// bound by the latency of its own weight / lane-adjust loads: 22 % issue utilisation, 9 % of the DRAM bandwidth)
constexpr int NBD = 2;
template <int T>
__global__ void __launch_bounds__(256, 3) nlinear_decode_tiled_kernel(const void* __restrict__ enc, int enc_dtype, const float* __restrict__ wd,
                                                                   const float* __restrict__ bd, const void* __restrict__ adj, int adj_dtype,
                                                                   void* __restrict__ dec, int out_dtype, int B, int C, int To) {
  const int c = threadIdx.x % C;
  const int slot = threadIdx.x / C, slots = blockDim.x / C;
  const long long b0 = ((long long)blockIdx.x * slots + slot) * NBD;
  if (b0 >= B) return;
  float e[NBD][T], last[NBD];
#pragma unroll
  for (int j = 0; j < NBD; ++j) {
    const long long b = b0 + j < B ? b0 + j : B - 1;
#pragma unroll
    for (int s = 0; s < T; ++s) e[j][s] = load_as_f(enc, ((size_t)b * T + s) * C + c, enc_dtype);
    last[j] = e[j][T - 1];
#pragma unroll
    for (int s = 0; s < T; ++s) e[j][s] -= last[j];
  }
  // grid.y splits the horizon: more resident warps to hide the L2 / DRAM latency of the weight, lane-adjust and output traffic
  const int t_per = (To + gridDim.y - 1) / gridDim.y;
  const int t_lo = blockIdx.y * t_per, t_hi = min(To, t_lo + t_per);
#pragma unroll 2
  for (int t = t_lo; t < t_hi; ++t) {
    const float base = __ldg(bd + (size_t)t * C + c);
    float acc[NBD];
#pragma unroll
    for (int j = 0; j < NBD; ++j) acc[j] = base + last[j];
#pragma unroll
    for (int s = 0; s < T; ++s) {
      const float w = __ldg(wd + ((size_t)t * T + s) * C + c);
#pragma unroll
      for (int j = 0; j < NBD; ++j) acc[j] = fmaf(w, e[j][s], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NBD; ++j) {
      if (b0 + j < B) {
        const size_t o = ((size_t)(b0 + j) * To + t) * C + c;
        float v = acc[j];
        if (adj) v += load_as_f(adj, o, adj_dtype);
        store_from_f(dec, o, out_dtype, v);
      }
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

// Per-scene error statistics shared by fusion_head and traj_metrics.  Thread `t` owns time step t.
//   sq_x, sq_y: squared de-normalised errors (MSE numerators, train.py:954-961)
//   dist      : sqrt(dx^2 + dy^2)   (train.py:1318-1320)
__device__ __forceinline__ void scene_error(float px, float py, float gx, float gy, const float* ns, float& sqx, float& sqy, float& dist) {
  const float rx = ns[1] - ns[0], ry = ns[3] - ns[2];
  const float dx = (px * rx + ns[0]) - (gx * rx + ns[0]);
  const float dy = (py * ry + ns[2]) - (gy * ry + ns[2]);
  sqx = dx * dx;
  sqy = dy * dy;
  dist = sqrtf(dx * dx + dy * dy);
}

// reference scripts/train.py:801-805 (fusion_layer, out_proj), 941-943 (+last input position), 945-962, 1302-1322.
// One block per scene (grid-stride); warp w handles time steps 4w .. 4w+3, 4w+32 .., ...; C <= 128, C % 32 == 0.
// w1t / w2t are the transposed 64x64 weights ([in][out]) staged in shared memory: lanes read consecutive
// outputs (conflict-free) while the input activation is a broadcast.
template <int CPL>  // channels per lane = C / 32
__global__ void __launch_bounds__(256) fusion_head_kernel(const void* __restrict__ fused, int in_dtype, const float* __restrict__ ln_w,
                                                          const float* __restrict__ ln_b, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, const float* __restrict__ wo,
                                                          const float* __restrict__ bo, const float* __restrict__ x,
                                                          float* __restrict__ decoded, const float* __restrict__ y,
                                                          const float* __restrict__ norm_stat, float* __restrict__ metrics,
                                                          float* __restrict__ per_scene, int B, int T_in, int T_out) {
  constexpr int C = CPL * 32;
  extern __shared__ float sm[];
  float* w1t = sm;                 // [C][C] : w1t[i*C + o] = w1[o*C + i]
  float* w2t = w1t + C * C;
  float* act = w2t + C * C;        // [8 warps][C][4 rows]
  float* red = act + 8 * 4 * C;    // [8][4]
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    const int o = i / C, k = i % C;
    w1t[k * C + o] = __ldg(w1 + i);
    w2t[k * C + o] = __ldg(w2 + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Each warp takes RW = 4 time steps at once: the activations of the four rows sit interleaved in shared memory ([k][4], one
  // 16-byte broadcast read per k), so one pass over a weight matrix feeds 4 x CPL FMAs per two weight reads.
  constexpr int RW = 4;
  float4* my4 = reinterpret_cast<float4*>(act + warp * RW * C);     // [C] x (4 rows)
  float* my = act + warp * RW * C;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float s_sqx = 0.f, s_sqy = 0.f, s_dist = 0.f, s_fde = 0.f;
    for (int t0 = warp * RW; t0 < T_out; t0 += 8 * RW) {
      float v[RW][CPL];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const int t = t0 + r < T_out ? t0 + r : T_out - 1;          // tail rows recompute the last step (never stored)
        const size_t row = (size_t)b * T_out + t;
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < CPL; ++e) {
          v[r][e] = load_as_f(fused, row * C + lane + 32 * e, in_dtype);
          sum += v[r][e];
        }
        const float mean = warp_sum(sum) / C;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < CPL; ++e) q += (v[r][e] - mean) * (v[r][e] - mean);
        const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
#pragma unroll
        for (int e = 0; e < CPL; ++e) {
          const int c = lane + 32 * e;
          my[c * RW + r] = (v[r][e] - mean) * rstd * __ldg(ln_w + c) + __ldg(ln_b + c);
        }
      }
      __syncwarp();
      float h[RW][CPL];
#pragma unroll
      for (int e = 0; e < CPL; ++e) {
        const float bb = __ldg(b1 + lane + 32 * e);
#pragma unroll
        for (int r = 0; r < RW; ++r) h[r][e] = bb;
      }
      for (int k = 0; k < C; ++k) {
        const float4 a = my4[k];
#pragma unroll
        for (int e = 0; e < CPL; ++e) {
          const float w = w1t[k * C + lane + 32 * e];
          h[0][e] = fmaf(w, a.x, h[0][e]);
          h[1][e] = fmaf(w, a.y, h[1][e]);
          h[2][e] = fmaf(w, a.z, h[2][e]);
          h[3][e] = fmaf(w, a.w, h[3][e]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < CPL; ++e)
#pragma unroll
        for (int r = 0; r < RW; ++r) my[(lane + 32 * e) * RW + r] = fmaxf(h[r][e], 0.f);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < CPL; ++e) {
        const float bb = __ldg(b2 + lane + 32 * e);
#pragma unroll
        for (int r = 0; r < RW; ++r) h[r][e] = bb;
      }
      for (int k = 0; k < C; ++k) {
        const float4 a = my4[k];
#pragma unroll
        for (int e = 0; e < CPL; ++e) {
          const float w = w2t[k * C + lane + 32 * e];
          h[0][e] = fmaf(w, a.x, h[0][e]);
          h[1][e] = fmaf(w, a.y, h[1][e]);
          h[2][e] = fmaf(w, a.z, h[2][e]);
          h[3][e] = fmaf(w, a.w, h[3][e]);
        }
      }
      __syncwarp();
      const float xl0 = __ldg(x + ((size_t)b * 2 + 0) * T_in + T_in - 1), xl1 = __ldg(x + ((size_t)b * 2 + 1) * T_in + T_in - 1);
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int e = 0; e < CPL; ++e) {
          const int c = lane + 32 * e;
          o0 = fmaf(__ldg(wo + c), h[r][e], o0);
          o1 = fmaf(__ldg(wo + C + c), h[r][e], o1);
        }
        o0 = warp_sum(o0) + __ldg(bo) + xl0;
        o1 = warp_sum(o1) + __ldg(bo + 1) + xl1;
        const int t = t0 + r;
        if (lane == 0 && t < T_out) {
          decoded[((size_t)b * 2 + 0) * T_out + t] = o0;
          decoded[((size_t)b * 2 + 1) * T_out + t] = o1;
          if (y) {
            float sqx, sqy, dist;
            scene_error(o0, o1, y[((size_t)b * 2 + 0) * T_out + t], y[((size_t)b * 2 + 1) * T_out + t], norm_stat + (size_t)b * 4, sqx, sqy, dist);
            s_sqx += sqx; s_sqy += sqy; s_dist += dist;
            if (t == T_out - 1) s_fde = dist;
          }
        }
      }
    }
    if (y) {   // uniform across the block
      if (lane == 0) {
        red[warp * 4 + 0] = s_sqx; red[warp * 4 + 1] = s_sqy; red[warp * 4 + 2] = s_dist; red[warp * 4 + 3] = s_fde;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w * 4 + threadIdx.x];
        if (threadIdx.x == 2) t /= (float)T_out;                    // ADE_b
        if (threadIdx.x >= 2 && per_scene) per_scene[(size_t)b * 2 + threadIdx.x - 2] = t;
        atomicAdd(metrics + threadIdx.x, t);
      } else if (threadIdx.x == 4) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w * 4] + red[w * 4 + 1];
        atomicAdd(metrics + 4, t / ((float)B * (float)T_out));      // MSE_x + MSE_y (train.py:959-961)
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core form of the fusion head for d_model = 64 (the bf16 compute mode; the FFMA kernel above stays the exact fp32 path).
// The two 64 x 64 fusion_layer linears are batched as mma.sync m16n8k16 tiles over (scene, step) rows with the weights resident in
// shared memory; fp32-class accuracy comes from the split-bf16 form (x = hi + lo, three products hi.hi + lo.hi + hi.lo, fp32
// accumulation: relative error ~2^-16).  A warp owns a 16-row tile from the fused-feature load to the ADE / FDE contribution:
// LayerNorm on the A-fragment layout (a row lives in one quad: two shuffles), linear 1 + ReLU, the accumulator fragments re-packed
// in registers as the next A operand (no shared-memory round trip), linear 2, out_proj as a quad reduction, de-normalised errors
// into per-scene shared-memory accumulators.  A CTA owns GROUP consecutive scenes, so a scene's ADE / FDE never crosses CTAs.
// ------------------------------------------------------------------------------------------------
namespace fh {
constexpr int C = 64, LDW = C + 8, GROUP = 8;
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (a, b) -> packed bf16 pair of the high parts and of the residuals
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// acc[nt][.] += A . W^T for the 16 x 64 tile held as split fragments; W (hi / lo) in shared memory as [out][in] rows of LDW elements
__device__ __forceinline__ void linear64(float (&acc)[8][4], const uint32_t (&ahi)[4][4], const uint32_t (&alo)[4][4], const __nv_bfloat16* whi,
                                         const __nv_bfloat16* wlo, int g, int t4) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int o = (nt * 8 + g) * LDW + kt * 16 + t4 * 2;
      const uint32_t h0 = *reinterpret_cast<const uint32_t*>(whi + o), h1 = *reinterpret_cast<const uint32_t*>(whi + o + 8);
      const uint32_t l0 = *reinterpret_cast<const uint32_t*>(wlo + o), l1 = *reinterpret_cast<const uint32_t*>(wlo + o + 8);
      mma16816(acc[nt], ahi[kt], h0, h1);
      mma16816(acc[nt], alo[kt], h0, h1);
      mma16816(acc[nt], ahi[kt], l0, l1);
    }
  }
}
}  // namespace fh

__global__ void __launch_bounds__(256) fusion_head_tc_kernel(const void* __restrict__ fused, int in_dtype, const float* __restrict__ ln_w,
                                                             const float* __restrict__ ln_b, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, const float* __restrict__ w2,
                                                             const float* __restrict__ b2, const float* __restrict__ wo,
                                                             const float* __restrict__ bo, const float* __restrict__ x,
                                                             float* __restrict__ decoded, const float* __restrict__ y,
                                                             const float* __restrict__ norm_stat, float* __restrict__ metrics,
                                                             float* __restrict__ per_scene, int B, int T_in, int T_out) {
  using namespace fh;
  __shared__ __align__(16) __nv_bfloat16 sw[4][C * LDW];      // w1 hi, w1 lo, w2 hi, w2 lo
  __shared__ float sv[6][C];                                   // ln_w, ln_b, b1, b2, wo[0], wo[1]
  __shared__ float sacc[GROUP][4];                             // per scene of the group: sum sq_x, sum sq_y, sum dist, fde
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    const int o = i / C, k = i % C;
    uint32_t hi, lo;
    split2(__ldg(w1 + i), 0.f, hi, lo);
    sw[0][o * LDW + k] = __ushort_as_bfloat16((unsigned short)(hi & 0xffffu));
    sw[1][o * LDW + k] = __ushort_as_bfloat16((unsigned short)(lo & 0xffffu));
    split2(__ldg(w2 + i), 0.f, hi, lo);
    sw[2][o * LDW + k] = __ushort_as_bfloat16((unsigned short)(hi & 0xffffu));
    sw[3][o * LDW + k] = __ushort_as_bfloat16((unsigned short)(lo & 0xffffu));
  }
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sv[0][c] = __ldg(ln_w + c); sv[1][c] = __ldg(ln_b + c); sv[2][c] = __ldg(b1 + c); sv[3][c] = __ldg(b2 + c);
    sv[4][c] = __ldg(wo + c); sv[5][c] = __ldg(wo + C + c);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const float bo0 = __ldg(bo), bo1 = __ldg(bo + 1);
  const int ngroups = (B + GROUP - 1) / GROUP;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b_lo = grp * GROUP, nb = min(GROUP, B - b_lo);
    const int rows = nb * T_out;
    __syncthreads();                       // weights staged (first pass) / the previous group's accumulators consumed
    if (threadIdx.x < GROUP * 4) (&sacc[0][0])[threadIdx.x] = 0.f;
    __syncthreads();
    for (int tile = warp; tile * 16 < rows; tile += 8) {
      const int r0 = tile * 16 + g, r1 = r0 + 8;                   // rows of the group this thread's fragments hold
      // ---- load + LayerNorm (train.py:759-764 norm of fusion_layer) on the A-fragment layout
      float v[2][16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = min(h ? r1 : r0, rows - 1);                  // tail rows recompute the last row (never stored)
        const size_t base = ((size_t)b_lo * T_out + r) * C;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = kt * 16 + q * 8 + t4 * 2;
            if (in_dtype == TCAVP_F32) {
              const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(fused) + base + c);
              v[h][kt * 4 + q * 2] = f.x; v[h][kt * 4 + q * 2 + 1] = f.y;
            } else {
              const __nv_bfloat162 f = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(fused) + base + c);
              v[h][kt * 4 + q * 2] = __low2float(f); v[h][kt * 4 + q * 2 + 1] = __high2float(f);
            }
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) sum += v[h][e];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / C);
        float qs = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) qs += (v[h][e] - mean) * (v[h][e] - mean);
        qs += __shfl_xor_sync(0xffffffffu, qs, 1);
        qs += __shfl_xor_sync(0xffffffffu, qs, 2);
        const float rstd = rsqrtf(qs * (1.f / C) + 1e-5f);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = kt * 16 + q * 8 + t4 * 2 + e;
              v[h][kt * 4 + q * 2 + e] = (v[h][kt * 4 + q * 2 + e] - mean) * rstd * sv[0][c] + sv[1][c];
            }
      }
      uint32_t ahi[4][4], alo[4][4];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        // fragment registers: (row g, cols 2 t4 + {0, 1}), (row g + 8, same), (row g, cols + 8), (row g + 8, cols + 8)
        split2(v[0][kt * 4], v[0][kt * 4 + 1], ahi[kt][0], alo[kt][0]);
        split2(v[1][kt * 4], v[1][kt * 4 + 1], ahi[kt][1], alo[kt][1]);
        split2(v[0][kt * 4 + 2], v[0][kt * 4 + 3], ahi[kt][2], alo[kt][2]);
        split2(v[1][kt * 4 + 2], v[1][kt * 4 + 3], ahi[kt][3], alo[kt][3]);
      }
      // ---- linear 1 + ReLU (train.py:801 fusion_layer[0..1])
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] = acc[nt][2] = sv[2][nt * 8 + t4 * 2];
        acc[nt][1] = acc[nt][3] = sv[2][nt * 8 + t4 * 2 + 1];
      }
      linear64(acc, ahi, alo, sw[0], sw[1], g, t4);
      // accumulator tiles (2 kt, 2 kt + 1) are exactly the A fragment of k-step kt of the next product
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        split2(fmaxf(acc[2 * kt][0], 0.f), fmaxf(acc[2 * kt][1], 0.f), ahi[kt][0], alo[kt][0]);
        split2(fmaxf(acc[2 * kt][2], 0.f), fmaxf(acc[2 * kt][3], 0.f), ahi[kt][1], alo[kt][1]);
        split2(fmaxf(acc[2 * kt + 1][0], 0.f), fmaxf(acc[2 * kt + 1][1], 0.f), ahi[kt][2], alo[kt][2]);
        split2(fmaxf(acc[2 * kt + 1][2], 0.f), fmaxf(acc[2 * kt + 1][3], 0.f), ahi[kt][3], alo[kt][3]);
      }
      // ---- linear 2 (fusion_layer[2]) and out_proj (train.py:803) as a quad reduction
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] = acc[nt][2] = sv[3][nt * 8 + t4 * 2];
        acc[nt][1] = acc[nt][3] = sv[3][nt * 8 + t4 * 2 + 1];
      }
      linear64(acc, ahi, alo, sw[2], sw[3], g, t4);
      float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};       // [row half][coordinate]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + t4 * 2;
        const float wx0 = sv[4][c], wx1 = sv[4][c + 1], wy0 = sv[5][c], wy1 = sv[5][c + 1];
        o[0][0] = fmaf(wx0, acc[nt][0], fmaf(wx1, acc[nt][1], o[0][0]));
        o[0][1] = fmaf(wy0, acc[nt][0], fmaf(wy1, acc[nt][1], o[0][1]));
        o[1][0] = fmaf(wx0, acc[nt][2], fmaf(wx1, acc[nt][3], o[1][0]));
        o[1][1] = fmaf(wy0, acc[nt][2], fmaf(wy1, acc[nt][3], o[1][1]));
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          o[h][f] += __shfl_xor_sync(0xffffffffu, o[h][f], 1);
          o[h][f] += __shfl_xor_sync(0xffffffffu, o[h][f], 2);
        }
      // ---- + last observed position (train.py:941-943), store, error statistics (train.py:945-962, 1302-1322): lanes t4 = 0, 1 take one row each
      if (t4 < 2) {
        const int r = t4 ? r1 : r0;
        if (r < rows) {
          const int sb = r / T_out, t = r - sb * T_out, b = b_lo + sb;
          const float px = o[t4][0] + bo0 + __ldg(x + ((size_t)b * 2 + 0) * T_in + T_in - 1);
          const float py = o[t4][1] + bo1 + __ldg(x + ((size_t)b * 2 + 1) * T_in + T_in - 1);
          decoded[((size_t)b * 2 + 0) * T_out + t] = px;
          decoded[((size_t)b * 2 + 1) * T_out + t] = py;
          if (y) {
            float sqx, sqy, dist;
            scene_error(px, py, y[((size_t)b * 2 + 0) * T_out + t], y[((size_t)b * 2 + 1) * T_out + t], norm_stat + (size_t)b * 4, sqx, sqy, dist);
            atomicAdd(&sacc[sb][0], sqx);
            atomicAdd(&sacc[sb][1], sqy);
            atomicAdd(&sacc[sb][2], dist);
            if (t == T_out - 1) sacc[sb][3] = dist;
          }
        }
      }
    }
    if (y) {   // uniform across the block
      __syncthreads();
      if (warp == 0) {
        float sqx = 0.f, sqy = 0.f, ade = 0.f, fde = 0.f;
        if (lane < nb) {
          sqx = sacc[lane][0]; sqy = sacc[lane][1]; ade = sacc[lane][2] / (float)T_out; fde = sacc[lane][3];
          if (per_scene) { per_scene[(size_t)(b_lo + lane) * 2] = ade; per_scene[(size_t)(b_lo + lane) * 2 + 1] = fde; }
        }
        sqx = warp_sum(sqx); sqy = warp_sum(sqy); ade = warp_sum(ade); fde = warp_sum(fde);
        if (lane == 0) {
          atomicAdd(metrics + 0, sqx); atomicAdd(metrics + 1, sqy); atomicAdd(metrics + 2, ade); atomicAdd(metrics + 3, fde);
          atomicAdd(metrics + 4, (sqx + sqy) / ((float)B * (float)T_out));      // MSE_x + MSE_y (train.py:959-961)
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128) traj_metrics_kernel(const float* __restrict__ decoded, const float* __restrict__ y,
                                                           const float* __restrict__ norm_stat, float* __restrict__ metrics,
                                                           float* __restrict__ per_scene, int B, int T_out) {
  __shared__ float scratch[8];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float sqx = 0.f, sqy = 0.f, dist = 0.f, fde = 0.f;
    for (int t = threadIdx.x; t < T_out; t += blockDim.x) {
      float a, c, d;
      scene_error(decoded[((size_t)b * 2) * T_out + t], decoded[((size_t)b * 2 + 1) * T_out + t], y[((size_t)b * 2) * T_out + t],
                  y[((size_t)b * 2 + 1) * T_out + t], norm_stat + (size_t)b * 4, a, c, d);
      sqx += a; sqy += c; dist += d;
      if (t == T_out - 1) fde = d;
    }
    sqx = block_sum(sqx, scratch);
    sqy = block_sum(sqy, scratch);
    dist = block_sum(dist, scratch);
    fde = block_sum(fde, scratch);
    if (threadIdx.x == 0) {
      const float ade = dist / (float)T_out;
      if (per_scene) { per_scene[(size_t)b * 2] = ade; per_scene[(size_t)b * 2 + 1] = fde; }
      atomicAdd(metrics + 0, sqx); atomicAdd(metrics + 1, sqy); atomicAdd(metrics + 2, ade); atomicAdd(metrics + 3, fde);
      atomicAdd(metrics + 4, (sqx + sqy) / ((float)B * (float)T_out));
    }
  }
}


// Best-of-K candidate reduction (reference scripts/test.py:1336-1368, seed_fix_train.py:1183-1243): K candidate trajectories per scene
// -> de-normalise, per-candidate ADE / FDE / RMSE, minimum over the candidates, batch sums.  One warp per (scene, candidate) pass,
// one block per scene; candidates (B, K, 2, T), y (B, 2, T), norm_stat (B, 4) = (min_x, max_x, min_y, max_y).
__global__ void __launch_bounds__(256) best_of_k_kernel(const float* __restrict__ cand, const float* __restrict__ y, const float* __restrict__ ns,
                                                        float* __restrict__ per_scene, float* __restrict__ totals, int B, int K, int T) {
  __shared__ float best[8][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float min_x = ns[(size_t)b * 4], max_x = ns[(size_t)b * 4 + 1], min_y = ns[(size_t)b * 4 + 2], max_y = ns[(size_t)b * 4 + 3];
    const float rx = max_x - min_x, ry = max_y - min_y;
    float m_ade = INFINITY, m_fde = INFINITY, m_rmse = INFINITY;
    for (int k = warp; k < K; k += 8) {
      const float* c = cand + ((size_t)b * K + k) * 2 * T;
      float sd = 0.f, sq = 0.f, last = 0.f;
      for (int t = lane; t < T; t += 32) {
        const float dx = (c[t] * rx + min_x) - (y[(size_t)b * 2 * T + t] * rx + min_x);
        const float dy = (c[T + t] * ry + min_y) - (y[(size_t)b * 2 * T + T + t] * ry + min_y);
        const float e2 = dx * dx + dy * dy;
        const float e = sqrtf(e2);
        sd += e;
        sq += e2;
        if (t == T - 1) last = e;
      }
      sd = warp_sum(sd);
      sq = warp_sum(sq);
      last = warp_sum(last);
      m_ade = fminf(m_ade, sd / (float)T);
      m_fde = fminf(m_fde, last);
      m_rmse = fminf(m_rmse, sqrtf(sq / (2.f * (float)T)));
    }
    __syncthreads();
    if (lane == 0) { best[warp][0] = m_ade; best[warp][1] = m_fde; best[warp][2] = m_rmse; }
    __syncthreads();
    if (threadIdx.x < 3) {
      float v = INFINITY;
      for (int w = 0; w < 8; ++w) v = fminf(v, best[w][threadIdx.x]);
      if (per_scene) per_scene[(size_t)b * 3 + threadIdx.x] = v;
      if (totals) atomicAdd(totals + threadIdx.x, v);
    }
  }
}

}  // namespace tcavp

using namespace tcavp;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define DT_OK(d) ((d) == TCAVP_F32 || (d) == TCAVP_BF16)

extern "C" int tcavp_ltsf_encode(const float* x, const float* wt, const float* bt, const float* we, const float* be,
                                 const float* pos, void* enc, int out_dtype, int B, int F, int C, int T_in, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && F > 0 && C > 0 && T_in > 0 && T_in <= MAX_T_IN, "tcavp_ltsf_encode: bad shape (T_in=%d, max %d)", T_in, MAX_T_IN);
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && wt && bt && we && be && pos && enc && DT_OK(out_dtype), "tcavp_ltsf_encode: bad pointer/dtype");
  if (C <= 256 && 256 % C == 0) {      // register-tiled kernels for the history lengths the reference's drivers use
    const int per_block = (256 / C) * NB;
    const int grid = (B + per_block - 1) / per_block;
#define TCAVP_ENC(T)                                                                                                          \
  case T:                                                                                                                      \
    ltsf_encode_tiled_kernel<T><<<grid, 256, 0, STREAM(stream)>>>(x, wt, bt, we, be, pos, enc, out_dtype, B, F, C);              \
    return check_launch("ltsf_encode_kernel")
    switch (T_in) {
      TCAVP_ENC(6);
      TCAVP_ENC(10);
      TCAVP_ENC(15);
      TCAVP_ENC(18);
      TCAVP_ENC(20);
      default: break;
    }
#undef TCAVP_ENC
  }
  const long long total = (long long)B * C;
  ltsf_encode_kernel<<<(int)((total + 127) / 128), 128, 0, STREAM(stream)>>>(x, wt, bt, we, be, pos, enc, out_dtype, B, F, C, T_in);
  return check_launch("ltsf_encode_kernel");
}

extern "C" int tcavp_nlinear_decode(const void* enc, int enc_dtype, const float* wd, const float* bd, const void* lane_adj,
                                    int adj_dtype, void* dec, int out_dtype, int B, int C, int T_in, int T_out, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && C > 0 && T_in > 0 && T_in <= MAX_T_IN && T_out > 0, "tcavp_nlinear_decode: bad shape");
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(enc && wd && bd && dec && DT_OK(enc_dtype) && DT_OK(out_dtype) && (!lane_adj || DT_OK(adj_dtype)), "tcavp_nlinear_decode: bad pointer/dtype");
  if (C <= 256 && 256 % C == 0) {
    const int per_block = (256 / C) * NBD;
    const int grid = (B + per_block - 1) / per_block;
    int ty = (sm_count() * 9 + grid - 1) / grid;          // aim at ~3 waves of three resident blocks per SM
    ty = ty < 1 ? 1 : (ty > (T_out + 3) / 4 ? (T_out + 3) / 4 : ty);
#define TCAVP_DEC(T)                                                                                                                          \
  case T:                                                                                                                                      \
    nlinear_decode_tiled_kernel<T><<<dim3(grid, ty), 256, 0, STREAM(stream)>>>(enc, enc_dtype, wd, bd, lane_adj, adj_dtype, dec, out_dtype, B, C, T_out);  \
    return check_launch("nlinear_decode_kernel")
    switch (T_in) {
      TCAVP_DEC(6);
      TCAVP_DEC(10);
      TCAVP_DEC(15);
      TCAVP_DEC(18);
      TCAVP_DEC(20);
      default: break;
    }
#undef TCAVP_DEC
  }
  const long long total = (long long)B * C;
  nlinear_decode_kernel<<<(int)((total + 127) / 128), 128, 0, STREAM(stream)>>>(enc, enc_dtype, wd, bd, lane_adj, adj_dtype, dec, out_dtype, B, C, T_in, T_out);
  return check_launch("nlinear_decode_kernel");
}

extern "C" int tcavp_fusion_head(const void* fused, int in_dtype, const float* ln_w, const float* ln_b, const float* w1,
                                 const float* b1, const float* w2, const float* b2, const float* wo, const float* bo,
                                 const float* x, float* decoded, const float* y, const float* norm_stat, float* metrics,
                                 float* per_scene, int B, int C, int T_in, int T_out, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && T_in > 0 && T_out > 0, "tcavp_fusion_head: bad shape");
  TCAVP_REQUIRE(C == 32 || C == 64 || C == 128, "tcavp_fusion_head: d_model must be 32, 64 or 128 (got %d)", C);
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(fused && ln_w && ln_b && w1 && b1 && w2 && b2 && wo && bo && x && decoded && DT_OK(in_dtype), "tcavp_fusion_head: bad pointer/dtype");
  TCAVP_REQUIRE(!y || (norm_stat && metrics), "tcavp_fusion_head: y needs norm_stat and metrics");
  const size_t smem = (size_t)(2 * C * C + 8 * 4 * C + 32) * sizeof(float);
  const int grid = B < sm_count() * 5 ? B : sm_count() * 5;      // 41 KB of shared memory per block: five resident blocks per SM
#define LAUNCH(CPL)                                                                                                            \
  do {                                                                                                                         \
    TCAVP_CUDA(cudaFuncSetAttribute(fusion_head_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    fusion_head_kernel<CPL><<<grid, 256, smem, STREAM(stream)>>>(fused, in_dtype, ln_w, ln_b, w1, b1, w2, b2, wo, bo, x, decoded, y, \
                                                                 norm_stat, metrics, per_scene, B, T_in, T_out);               \
  } while (0)
  if (C == 32) LAUNCH(1);
  else if (C == 64) LAUNCH(2);
  else LAUNCH(4);
#undef LAUNCH
  return check_launch("fusion_head_kernel");
}

extern "C" int tcavp_fusion_head_tc(const void* fused, int in_dtype, const float* ln_w, const float* ln_b, const float* w1,
                                    const float* b1, const float* w2, const float* b2, const float* wo, const float* bo,
                                    const float* x, float* decoded, const float* y, const float* norm_stat, float* metrics,
                                    float* per_scene, int B, int C, int T_in, int T_out, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && T_in > 0 && T_out > 0, "tcavp_fusion_head_tc: bad shape");
  TCAVP_REQUIRE(C == 64, "tcavp_fusion_head_tc: d_model must be 64 (got %d); tcavp_fusion_head covers 32 / 64 / 128", C);
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(fused && ln_w && ln_b && w1 && b1 && w2 && b2 && wo && bo && x && decoded && DT_OK(in_dtype), "tcavp_fusion_head_tc: bad pointer/dtype");
  TCAVP_REQUIRE(!y || (norm_stat && metrics), "tcavp_fusion_head_tc: y needs norm_stat and metrics");
  TCAVP_REQUIRE(reinterpret_cast<uintptr_t>(fused) % 8 == 0, "tcavp_fusion_head_tc: fused must be 8-byte aligned");
  const int groups = (B + fh::GROUP - 1) / fh::GROUP;
  const int grid = groups < sm_count() * 4 ? groups : sm_count() * 4;
  fusion_head_tc_kernel<<<grid, 256, 0, STREAM(stream)>>>(fused, in_dtype, ln_w, ln_b, w1, b1, w2, b2, wo, bo, x, decoded, y, norm_stat, metrics,
                                                          per_scene, B, T_in, T_out);
  return check_launch("fusion_head_tc_kernel");
}

extern "C" int tcavp_traj_metrics(const float* decoded, const float* y, const float* norm_stat, float* metrics, float* per_scene,
                                  int B, int T_out, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && T_out > 0, "tcavp_traj_metrics: bad shape");
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(decoded && y && norm_stat && metrics, "tcavp_traj_metrics: null pointer");
  const int grid = B < sm_count() * 8 ? B : sm_count() * 8;
  traj_metrics_kernel<<<grid, 128, 0, STREAM(stream)>>>(decoded, y, norm_stat, metrics, per_scene, B, T_out);
  return check_launch("traj_metrics_kernel");
}

extern "C" int tcavp_best_of_k(const float* candidates, const float* y, const float* norm_stat, float* per_scene, float* totals, int B, int K,
                               int T_out, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && K > 0 && T_out > 0, "tcavp_best_of_k: bad shape");
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(candidates && y && norm_stat && (per_scene || totals), "tcavp_best_of_k: null pointer");
  const int grid = B < sm_count() * 8 ? B : sm_count() * 8;
  best_of_k_kernel<<<grid, 256, 0, STREAM(stream)>>>(candidates, y, norm_stat, per_scene, totals, B, K, T_out);
  return check_launch("best_of_k_kernel");
}
