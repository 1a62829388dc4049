// LayerNorm (+residual, +row remap, +row vector) and RMSNorm.  Bandwidth-bound: one warp per row,
// 16-byte vector accesses, warp-shuffle reductions, fp32 statistics for both storage dtypes.
#include "common.cuh"

namespace tcavp {

// Load / store 8 consecutive elements (16 bytes for bf16, 32 bytes for fp32).
__device__ __forceinline__ void load8(const void* base, size_t idx, int dtype, float (&v)[8]) {
  if (dtype == TCAVP_BF16) {
    uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __bfloat1622float2(h[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store8(void* base, size_t idx, int dtype, const float (&v)[8]) {
  if (dtype == TCAVP_BF16) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// VEC = true: cols % 8 == 0 and all row starts are 16B (bf16) / 32B (fp32) aligned.
template <bool VEC>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* x, const void* res,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        void* out, int rows, int cols, float eps, int in_dtype,
                                                        int out_dtype, int gi, int go, int off, const float* __restrict__ rowvec) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = (size_t)row * cols;
  float s = 0.f;
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, in_dtype, v);
      if (res) {
        float r[8];
        load8(res, base + c, in_dtype, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[e];
    }
  } else {
    for (int c = lane; c < cols; c += 32) s += load_as_f(x, base + c, in_dtype) + (res ? load_as_f(res, base + c, in_dtype) : 0.f);
  }
  const float mean = warp_sum(s) / cols;
  float q = 0.f;
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, in_dtype, v);
      if (res) {
        float r[8];
        load8(res, base + c, in_dtype, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) q += (v[e] - mean) * (v[e] - mean);
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      float v = load_as_f(x, base + c, in_dtype) + (res ? load_as_f(res, base + c, in_dtype) : 0.f);
      q += (v - mean) * (v - mean);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / cols + eps);
  const size_t obase = (size_t)remap_row(gi, go, off, row) * cols;
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, in_dtype, v);
      if (res) {
        float r[8];
        load8(res, base + c, in_dtype, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = (v[e] - mean) * rstd * __ldg(w + c + e) + __ldg(b + c + e);
        if (rowvec) v[e] += __ldg(rowvec + c + e);
      }
      store8(out, obase + c, out_dtype, v);
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      float v = load_as_f(x, base + c, in_dtype) + (res ? load_as_f(res, base + c, in_dtype) : 0.f);
      v = (v - mean) * rstd * __ldg(w + c) + __ldg(b + c);
      if (rowvec) v += __ldg(rowvec + c);
      store_from_f(out, obase + c, out_dtype, v);
    }
  }
}

// Register-resident LayerNorm for rows of <= G * NV * 8 elements: G lanes share a row (32 / G rows per warp), every lane
// keeps NV 8-element vectors, so the row is read ONCE and narrow rows (d_model = 64: G = 8) still use all 32 lanes.
template <int G, int NV>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const void* __restrict__ x, const void* __restrict__ res,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            void* __restrict__ out, int rows, int cols, float eps, int in_dtype,
                                                            int out_dtype, int gi, int go, int off, const float* __restrict__ rowvec) {
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane % G;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / G;
  const bool live = row < rows;
  const size_t base = (size_t)(live ? row : 0) * cols;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * G + sub) * 8;
    if (live && c < cols) {
      load8(x, base + c, in_dtype, v[i]);
      if (res) {
        float r[8];
        load8(res, base + c, in_dtype, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] += r[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if ((i * G + sub) * 8 < cols) {
#pragma unroll
      for (int e = 0; e < 8; ++e) q += (v[i][e] - mean) * (v[i][e] - mean);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / cols + eps);
  if (!live) return;
  const size_t obase = (size_t)remap_row(gi, go, off, row) * cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * G + sub) * 8;
    if (c < cols) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o8[e] = (v[i][e] - mean) * rstd * ww[e] + bb[e];
        if (rowvec) o8[e] += __ldg(rowvec + c + e);
      }
      store8(out, obase + c, out_dtype, o8);
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                      void* __restrict__ out, int rows, int cols, int ldi, int ldo, float eps,
                                                      int in_dtype, int out_dtype) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = (size_t)row * ldi, obase = (size_t)row * ldo;
  float q = 0.f;
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, in_dtype, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) q += v[e] * v[e];
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      float v = load_as_f(x, base + c, in_dtype);
      q += v * v;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / cols + eps);
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, in_dtype, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = __ldg(w + c + e) * (v[e] * rstd);
      }
      store8(out, obase + c, out_dtype, v);
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      store_from_f(out, obase + c, out_dtype, __ldg(w + c) * (load_as_f(x, base + c, in_dtype) * rstd));
    }
  }
}

// rstd[r] = rsqrt(mean(x[r,:]^2) + eps): one warp per row, 16-byte loads
template <bool VEC>
__global__ void __launch_bounds__(256) row_rstd_kernel(const void* __restrict__ x, int ldx, int rows, int cols, float eps, int dtype,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = (size_t)row * ldx;
  float q = 0.f;
  if (VEC) {
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      load8(x, base + c, dtype, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) q += v[e] * v[e];
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      const float v = load_as_f(x, base + c, dtype);
      q += v * v;
    }
  }
  q = warp_sum(q);
  if (lane == 0) out[row] = rsqrtf(q / cols + eps);
}

static bool aligned_rows(const void* p, int ld, int dtype) {
  const size_t es = dtype == TCAVP_BF16 ? 2 : 4;
  return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (8 * es) == 0 && ((size_t)ld * es) % (8 * es) == 0);
}

}  // namespace tcavp

namespace tcavp {
// Register-resident RMSNorm for rows of <= 1024 elements: the row is read once (NV 8-element vectors per lane).
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_reg_kernel(const void* __restrict__ x, const float* __restrict__ w, void* __restrict__ out, int rows,
                                                          int cols, int ldi, int ldo, float eps, int in_dtype, int out_dtype) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = (size_t)row * ldi, obase = (size_t)row * ldo;
  float v[NV][8];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < cols) {
      load8(x, base + c, in_dtype, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) q += v[i][e] * v[i][e];
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / cols + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < cols) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float o8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] = v[i][e] * rstd * ww[e];
      store8(out, obase + c, out_dtype, o8);
    }
  }
}
}  // namespace tcavp

extern "C" int tcavp_layernorm(const void* x, const void* residual, const float* w, const float* b, void* out, int rows,
                               int cols, float eps, int in_dtype, int out_dtype, int remap_gi, int remap_go, int remap_off,
                               const float* rowvec, tcavp_stream_t stream_) {
  using namespace tcavp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TCAVP_REQUIRE(rows >= 0 && cols > 0, "tcavp_layernorm: bad shape rows=%d cols=%d", rows, cols);
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && w && b && out, "tcavp_layernorm: null pointer");
  TCAVP_REQUIRE((in_dtype | 1) == 1 && (out_dtype | 1) == 1, "tcavp_layernorm: bad dtype");
  const bool vec = cols % 8 == 0 && aligned_rows(x, cols, in_dtype) && aligned_rows(residual, cols, in_dtype) &&
                   aligned_rows(out, cols, out_dtype);
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  const bool wb_al = reinterpret_cast<uintptr_t>(w) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0;
  if (vec && wb_al && cols <= 1024) {
#define TCAVP_LN_REG(G, NV)                                                                                                         \
  do {                                                                                                                              \
    const int rpb = wpb * (32 / G);                                                                                                 \
    layernorm_reg_kernel<G, NV><<<(rows + rpb - 1) / rpb, wpb * 32, 0, stream>>>(x, residual, w, b, out, rows, cols, eps, in_dtype, \
                                                                               out_dtype, remap_gi, remap_go, remap_off, rowvec); \
  } while (0)
    if (cols <= 64) TCAVP_LN_REG(8, 1);
    else if (cols <= 128) TCAVP_LN_REG(16, 1);
    else if (cols <= 256) TCAVP_LN_REG(32, 1);
    else if (cols <= 512) TCAVP_LN_REG(32, 2);
    else if (cols <= 768) TCAVP_LN_REG(32, 3);
    else TCAVP_LN_REG(32, 4);
#undef TCAVP_LN_REG
  } else if (vec)
    layernorm_kernel<true><<<grid, wpb * 32, 0, stream>>>(x, residual, w, b, out, rows, cols, eps, in_dtype, out_dtype, remap_gi, remap_go, remap_off, rowvec);
  else
    layernorm_kernel<false><<<grid, wpb * 32, 0, stream>>>(x, residual, w, b, out, rows, cols, eps, in_dtype, out_dtype, remap_gi, remap_go, remap_off, rowvec);
  return check_launch("layernorm_kernel");
}

extern "C" int tcavp_row_rstd(const void* x, int ldx, int rows, int cols, float eps, int dtype, float* out, tcavp_stream_t stream_) {
  using namespace tcavp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TCAVP_REQUIRE(rows >= 0 && cols > 0 && ldx >= cols, "tcavp_row_rstd: bad shape rows=%d cols=%d ldx=%d", rows, cols, ldx);
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && out && (dtype | 1) == 1, "tcavp_row_rstd: bad pointer/dtype");
  const int grid = (rows + 7) / 8;
  if (cols % 8 == 0 && aligned_rows(x, ldx, dtype))
    row_rstd_kernel<true><<<grid, 256, 0, stream>>>(x, ldx, rows, cols, eps, dtype, out);
  else
    row_rstd_kernel<false><<<grid, 256, 0, stream>>>(x, ldx, rows, cols, eps, dtype, out);
  return check_launch("row_rstd_kernel");
}

extern "C" int tcavp_rmsnorm(const void* x, int ldi, const float* w, void* out, int rows, int cols, int ldo, float eps, int in_dtype,
                             int out_dtype, tcavp_stream_t stream_) {
  using namespace tcavp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TCAVP_REQUIRE(rows >= 0 && cols > 0 && ldo >= cols && ldi >= cols, "tcavp_rmsnorm: bad shape rows=%d cols=%d ldi=%d ldo=%d", rows, cols, ldi, ldo);
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && w && out, "tcavp_rmsnorm: null pointer");
  TCAVP_REQUIRE((in_dtype | 1) == 1 && (out_dtype | 1) == 1, "tcavp_rmsnorm: bad dtype");
  const bool vec = cols % 8 == 0 && aligned_rows(x, ldi, in_dtype) && aligned_rows(out, ldo, out_dtype);
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  if (vec && cols <= 1024 && reinterpret_cast<uintptr_t>(w) % 16 == 0) {
    if (cols <= 256) rmsnorm_reg_kernel<1><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
    else if (cols <= 512) rmsnorm_reg_kernel<2><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
    else if (cols <= 768) rmsnorm_reg_kernel<3><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
    else rmsnorm_reg_kernel<4><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
  } else if (vec)
    rmsnorm_kernel<true><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
  else
    rmsnorm_kernel<false><<<grid, wpb * 32, 0, stream>>>(x, w, out, rows, cols, ldi, ldo, eps, in_dtype, out_dtype);
  return check_launch("rmsnorm_kernel");
}
