// Small bandwidth-bound helpers: RoPE (in place), RoPE table, text-embedding gather into the fused
// sequence buffer, row-vector add with row remap, dtype cast, polygon input projection, masked mean.
#include "common.cuh"

namespace tcavp {

__global__ void rope_table_kernel(float* __restrict__ cs, const float* __restrict__ inv_freq, int L, int half, int layout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * half) return;
  const int pos = i / half, j = i % half;
  // HF:131-134: angle = pos * inv_freq in fp32 (inv_freq comes from the host, computed exactly as HF:86-88 does)
  const float ang = (float)pos * __ldg(inv_freq + j);
  // layout 0: [L][dh/2][2]   layout 1: [dh/4][L][4] (pairs 2q, 2q+1 of position pos share one float4)
  const size_t o = layout == 0 ? 2 * (size_t)i : ((size_t)(j >> 1) * L + pos) * 4 + (j & 1) * 2;
  cs[o] = cosf(ang);
  cs[o + 1] = sinf(ang);
}

// One thread per (row, head, 8 consecutive pair indices); q heads then k heads are adjacent in the row.
template <typename T>
__global__ void rope_kernel(T* __restrict__ qkv, int rows, int L, int ld, int heads, int dh, const float* __restrict__ cs) {
  const int half = dh >> 1;
  const long long total = (long long)rows * heads * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % half);
    const int h = (int)((i / half) % heads);
    const long long r = i / ((long long)half * heads);
    const int pos = (int)(r % L);
    const float c = __ldg(cs + 2 * ((size_t)pos * half + j)), s = __ldg(cs + 2 * ((size_t)pos * half + j) + 1);
    T* p = qkv + (size_t)r * ld + (size_t)h * dh;
    const float x1 = Cvt<T>::to_f(p[j]), x2 = Cvt<T>::to_f(p[j + half]);
    p[j] = Cvt<T>::from_f(x1 * c - x2 * s);          // q*cos + rotate_half(q)*sin, first half: -q[j+half]
    p[j + half] = Cvt<T>::from_f(x2 * c + x1 * s);   // second half: +q[j]
  }
}

__global__ void embed_text_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ amask, const void* __restrict__ embed,
                                  int embed_dtype, const float* __restrict__ text_mod, void* __restrict__ fused, int fused_dtype,
                                  int32_t* __restrict__ mask_out, int B, int L_text, int n_img, int H, int vocab) {
  const int L = n_img + L_text;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);   // over B * L
  if (row >= (long long)B * L) return;
  const int b = (int)(row / L), t = (int)(row % L);
  if (t < n_img) {
    if (lane == 0 && mask_out) mask_out[row] = 1;
    return;
  }
  const int j = t - n_img;
  if (lane == 0 && mask_out) mask_out[row] = amask ? (amask[(size_t)b * L_text + j] != 0) : 1;
  long long id = ids[(size_t)b * L_text + j];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int c = lane; c < H; c += 32) {
    const float v = load_as_f(embed, (size_t)id * H + c, embed_dtype) + __ldg(text_mod + c);
    store_from_f(fused, (size_t)row * H + c, fused_dtype, v);
  }
}

__global__ void add_rowvec_kernel(const void* __restrict__ x, const float* __restrict__ rowvec, void* __restrict__ out, int rows,
                                  int cols, int in_dtype, int out_dtype, int gi, int go, int off) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = load_as_f(x, i, in_dtype) + (rowvec ? __ldg(rowvec + c) : 0.f);
    store_from_f(out, (size_t)remap_row(gi, go, off, r) * cols + c, out_dtype, v);
  }
}

__global__ void cast_kernel(const void* __restrict__ in, int ldi, int in_dtype, void* __restrict__ out, int ldo, int out_dtype,
                            int rows, int cols, int in_row_mod) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    const long long ri = in_row_mod > 0 ? r % in_row_mod : r;
    store_from_f(out, (size_t)r * ldo + c, out_dtype, load_as_f(in, (size_t)ri * ldi + c, in_dtype));
  }
}

// 8 elements per thread (16-byte bf16 / 32-byte fp32 accesses): cols % 8 == 0 and 16-byte aligned rows on both sides.
__global__ void __launch_bounds__(256) cast_vec_kernel(const void* __restrict__ in, int ldi, int in_dtype, void* __restrict__ out, int ldo,
                                                       int out_dtype, int rows, int cols, int in_row_mod) {
  const int oct = cols >> 3;
  const long long total = (long long)rows * oct;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / oct;
    const int c = (int)(i % oct) * 8;
    const long long ri = in_row_mod > 0 ? r % in_row_mod : r;
    float v[8];
    if (in_dtype == TCAVP_BF16) {
      const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)ri * ldi + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[2 * e] = __uint_as_float(w[e] << 16);
        v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      }
    } else {
      const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + (size_t)ri * ldi + c);
      const float4 a = p4[0], b = p4[1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (out_dtype == TCAVP_BF16) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)r * ldo + c) = u;
    } else {
      float4* p4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)r * ldo + c);
      p4[0] = make_float4(v[0], v[1], v[2], v[3]);
      p4[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// Counter-based dropout (include/tcavp.h): out = [out +] [residual +] keep(r * cols + c) ? in * scale : 0.  Eight columns per thread
// on the vector path (16-byte bf16 / 32-byte fp32 accesses); the mask is recomputed from (seed, site, index), never stored.
template <int V>
__global__ void __launch_bounds__(256) dropout_kernel(const void* in, int ldi, int in_dtype, const void* __restrict__ res, int ldr,
                                                      int res_dtype, void* out, int ldo, int out_dtype, long long rows, int cols,
                                                      const uint32_t* __restrict__ seed, uint32_t site, uint32_t thresh, float scale, int accumulate) {
  const uint32_t key = drop_key(seed, site);
  const int per = cols / V;
  const long long total = rows * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per;
    const int c = (int)(i % per) * V;
    float v[V], o[V];
    if (V == 8) {
      if (in_dtype == TCAVP_BF16) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)r * ldi + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[2 * e] = __uint_as_float(w[e] << 16);
          v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
      } else {
        const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + (size_t)r * ldi + c);
        const float4 a = p4[0], b = p4[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
    } else {
      v[0] = load_as_f(in, (size_t)r * ldi + c, in_dtype);
    }
    const unsigned long long base = (unsigned long long)r * (unsigned long long)cols + (unsigned long long)c;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      o[e] = drop_keep(key, base + e, thresh) ? v[e] * scale : 0.f;
      if (res) o[e] += load_as_f(res, (size_t)r * ldr + c + e, res_dtype);
      if (accumulate) o[e] += load_as_f(out, (size_t)r * ldo + c + e, out_dtype);
    }
    if (V == 8 && out_dtype == TCAVP_BF16) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)r * ldo + c) = u;
    } else if (V == 8) {
      float4* p4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)r * ldo + c);
      p4[0] = make_float4(o[0], o[1], o[2], o[3]);
      p4[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
      store_from_f(out, (size_t)r * ldo + c, out_dtype, o[0]);
    }
  }
}

__global__ void poly_embed_kernel(const float* __restrict__ poly, const int32_t* __restrict__ len, const float* __restrict__ w,
                                  const float* __restrict__ bias, const float* __restrict__ pos, void* __restrict__ out,
                                  int out_dtype, int32_t* __restrict__ key_mask, int B, int P, int D) {
  const long long total = (long long)B * P * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long bp = i / D;
    const int p = (int)(bp % P), b = (int)(bp / P);
    const float px = __ldg(poly + bp * 2), py = __ldg(poly + bp * 2 + 1);
    const float v = fmaf(__ldg(w + 2 * d), px, fmaf(__ldg(w + 2 * d + 1), py, __ldg(bias + d))) + __ldg(pos + (size_t)p * D + d);
    store_from_f(out, i, out_dtype, v);
    if (d == 0 && key_mask) key_mask[bp] = p < __ldg(len + b);
  }
}

// fp32 output, D a multiple of 4: four features per thread — one 8-byte point load shared by the D / 4 threads of a point, 16-byte
// weight / bias / position-table loads (L1-resident: 2D + D + P D floats), one 16-byte store.  The kernel is a pure 4 D-byte-per-point
// write stream; the element-per-thread form above spent its time on 64-bit index arithmetic and 4-byte stores.
__global__ void __launch_bounds__(256) poly_embed_vec_kernel(const float* __restrict__ poly, const int32_t* __restrict__ len, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ pos, float* __restrict__ out,
                                                             int32_t* __restrict__ key_mask, int B, int P, int D) {
  const int dq = D >> 2;
  const unsigned total = (unsigned)B * (unsigned)P * (unsigned)dq;      // the host checks B P D < 2^31
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned bp = i / (unsigned)dq;
    const int d = (int)(i - bp * (unsigned)dq) << 2;
    const int b = (int)(bp / (unsigned)P), p = (int)(bp - (unsigned)b * (unsigned)P);
    const float2 pt = __ldg(reinterpret_cast<const float2*>(poly) + bp);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + 2 * d)), w1 = __ldg(reinterpret_cast<const float4*>(w + 2 * d + 4));
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias + d));
    const float4 ps = __ldg(reinterpret_cast<const float4*>(pos + (size_t)p * D + d));
    float4 o;      // same operation order as the scalar kernel: fma(wx, px, fma(wy, py, bias)) + pos
    o.x = fmaf(w0.x, pt.x, fmaf(w0.y, pt.y, bs.x)) + ps.x;
    o.y = fmaf(w0.z, pt.x, fmaf(w0.w, pt.y, bs.y)) + ps.y;
    o.z = fmaf(w1.x, pt.x, fmaf(w1.y, pt.y, bs.z)) + ps.z;
    o.w = fmaf(w1.z, pt.x, fmaf(w1.w, pt.y, bs.w)) + ps.w;
    reinterpret_cast<float4*>(out)[i] = o;
    if (d == 0 && key_mask) key_mask[bp] = p < __ldg(len + b);
  }
}

__global__ void masked_mean_kernel(const void* __restrict__ x, int in_dtype, const int32_t* __restrict__ len, void* __restrict__ out,
                                   int out_dtype, int B, int P, int D) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int d = (int)(i % D), b = (int)(i / D);
  int n = __ldg(len + b);
  n = n < 0 ? 0 : (n > P ? P : n);
  float s = 0.f;
  for (int p = 0; p < n; ++p) s += load_as_f(x, ((size_t)b * P + p) * D + d, in_dtype);
  store_from_f(out, i, out_dtype, n > 0 ? s / (float)n : 0.f);
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tcavp

namespace tcavp {
// fp32 -> three bf16 column blocks [hi | hi | lo] with hi = bf16(x), lo = bf16(x - hi): against weights packed [hi | lo | hi]
// one bf16 tensor-core GEMM over 3K columns returns x_hi w_hi + x_hi w_lo + x_lo w_hi, i.e. the fp32 product to ~2^-16 relative
// (fp32 accumulation).  Used for the few-MFLOP fp32 temporal / fusion layers in bf16 compute mode; the exact-fp32 parity mode keeps FFMA.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ out, int ldo, long long rows,
                                                     int cols) {
  const int q = cols >> 2;
  const long long total = rows * q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q;
    const int c = (int)(i % q) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = __float2bfloat16_rn(f[e]);
      lo[e] = __float2bfloat16_rn(f[e] - __bfloat162float(hi[e]));
    }
    __nv_bfloat16* o = out + (size_t)r * ldo + c;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(o + cols) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(o + 2 * cols) = *reinterpret_cast<const uint2*>(lo);
  }
}
}  // namespace tcavp

using namespace tcavp;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define DT_OK(d) ((d) == TCAVP_F32 || (d) == TCAVP_BF16)

extern "C" int tcavp_rope_table(float* cos_sin, const float* inv_freq, int L, int dh, int layout, tcavp_stream_t stream) {
  TCAVP_REQUIRE(cos_sin && inv_freq && L > 0 && dh > 0 && dh % 2 == 0, "tcavp_rope_table: bad args L=%d dh=%d", L, dh);
  TCAVP_REQUIRE(layout == 0 || (layout == 1 && dh % 4 == 0), "tcavp_rope_table: bad layout %d for dh=%d", layout, dh);
  const int n = L * (dh / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, STREAM(stream)>>>(cos_sin, inv_freq, L, dh / 2, layout);
  return check_launch("rope_table_kernel");
}

extern "C" int tcavp_rope(void* qkv, int rows, int L, int ld, int n_q_heads, int n_k_heads, int dh, const float* cos_sin,
                          int dtype, tcavp_stream_t stream) {
  TCAVP_REQUIRE(rows >= 0 && L > 0 && dh > 0 && dh % 2 == 0 && n_q_heads > 0 && n_k_heads >= 0, "tcavp_rope: bad shape");
  TCAVP_REQUIRE(ld >= (n_q_heads + n_k_heads) * dh, "tcavp_rope: ld %d too small", ld);
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(qkv && cos_sin && DT_OK(dtype), "tcavp_rope: bad pointer/dtype");
  const int heads = n_q_heads + n_k_heads;
  const long long total = (long long)rows * heads * (dh / 2);
  if (dtype == TCAVP_BF16)
    rope_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<__nv_bfloat16*>(qkv), rows, L, ld, heads, dh, cos_sin);
  else
    rope_kernel<float><<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<float*>(qkv), rows, L, ld, heads, dh, cos_sin);
  return check_launch("rope_kernel");
}

extern "C" int tcavp_embed_text(const int64_t* ids, const int64_t* attn_mask, const void* embed, int embed_dtype,
                                const float* text_mod, void* fused, int fused_dtype, int32_t* mask_out, int B, int L_text,
                                int n_img, int H, int vocab, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && L_text >= 0 && n_img >= 0 && H > 0 && vocab > 0, "tcavp_embed_text: bad shape");
  if (B == 0 || n_img + L_text == 0) return TCAVP_OK;
  TCAVP_REQUIRE((ids || L_text == 0) && embed && text_mod && fused && DT_OK(embed_dtype) && DT_OK(fused_dtype), "tcavp_embed_text: bad pointer/dtype");
  const long long rows = (long long)B * (n_img + L_text);
  embed_text_kernel<<<(int)((rows + 7) / 8), 256, 0, STREAM(stream)>>>(ids, attn_mask, embed, embed_dtype, text_mod, fused, fused_dtype,
                                                                        mask_out, B, L_text, n_img, H, vocab);
  return check_launch("embed_text_kernel");
}

extern "C" int tcavp_add_rowvec(const void* x, const float* rowvec, void* out, int rows, int cols, int in_dtype, int out_dtype,
                                int remap_gi, int remap_go, int remap_off, tcavp_stream_t stream) {
  TCAVP_REQUIRE(rows >= 0 && cols > 0, "tcavp_add_rowvec: bad shape");
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && out && DT_OK(in_dtype) && DT_OK(out_dtype), "tcavp_add_rowvec: bad pointer/dtype");
  add_rowvec_kernel<<<grid_for((long long)rows * cols, 256), 256, 0, STREAM(stream)>>>(x, rowvec, out, rows, cols, in_dtype, out_dtype, remap_gi, remap_go, remap_off);
  return check_launch("add_rowvec_kernel");
}

extern "C" int tcavp_cast(const void* in, int ldi, int in_dtype, void* out, int ldo, int out_dtype, int rows, int cols,
                          int in_row_mod, tcavp_stream_t stream) {
  TCAVP_REQUIRE(rows >= 0 && cols > 0 && ldi >= cols && ldo >= cols, "tcavp_cast: bad shape");
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(in && out && DT_OK(in_dtype) && DT_OK(out_dtype), "tcavp_cast: bad pointer/dtype");
  auto al = [](const void* p, int ld, int dtype) {
    const size_t esz = dtype == TCAVP_BF16 ? 2 : 4;
    return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ((size_t)ld * esz) % 16 == 0;
  };
  if (cols % 8 == 0 && al(in, ldi, in_dtype) && al(out, ldo, out_dtype)) {
    cast_vec_kernel<<<grid_for((long long)rows * (cols / 8), 256), 256, 0, STREAM(stream)>>>(in, ldi, in_dtype, out, ldo, out_dtype, rows, cols,
                                                                                          in_row_mod);
    return check_launch("cast_kernel");
  }
  cast_kernel<<<grid_for((long long)rows * cols, 256), 256, 0, STREAM(stream)>>>(in, ldi, in_dtype, out, ldo, out_dtype, rows, cols, in_row_mod);
  return check_launch("cast_kernel");
}

extern "C" int tcavp_dropout(const void* in, int ldi, int in_dtype, const void* residual, int ldr, int res_dtype, void* out, int ldo, int out_dtype,
                             long long rows, int cols, const uint32_t* seed, uint32_t site, uint32_t thresh, float scale, int accumulate,
                             tcavp_stream_t stream) {
  TCAVP_REQUIRE(rows >= 0 && cols > 0 && ldi >= cols && ldo >= cols && (!residual || ldr >= cols), "tcavp_dropout: bad shape");
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(in && out && seed && DT_OK(in_dtype) && DT_OK(out_dtype) && (!residual || DT_OK(res_dtype)), "tcavp_dropout: bad pointer/dtype");
  auto al = [](const void* p, int ld, int dtype) {
    const size_t esz = dtype == TCAVP_BF16 ? 2 : 4;
    return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ((size_t)ld * esz) % 16 == 0;
  };
  if (cols % 8 == 0 && al(in, ldi, in_dtype) && al(out, ldo, out_dtype)) {
    dropout_kernel<8><<<grid_for(rows * (cols / 8), 256), 256, 0, STREAM(stream)>>>(in, ldi, in_dtype, residual, ldr, res_dtype, out, ldo, out_dtype,
                                                                                  rows, cols, seed, site, thresh, scale, accumulate);
  } else {
    dropout_kernel<1><<<grid_for(rows * cols, 256), 256, 0, STREAM(stream)>>>(in, ldi, in_dtype, residual, ldr, res_dtype, out, ldo, out_dtype, rows,
                                                                            cols, seed, site, thresh, scale, accumulate);
  }
  return check_launch("dropout_kernel");
}

extern "C" int tcavp_poly_embed(const float* polygon, const int32_t* len, const float* w, const float* bias, const float* pos,
                                void* out, int out_dtype, int32_t* key_mask, int B, int P, int D, tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && P > 0 && D > 0, "tcavp_poly_embed: bad shape");
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(polygon && len && w && bias && pos && out && DT_OK(out_dtype), "tcavp_poly_embed: bad pointer/dtype");
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  if (out_dtype == TCAVP_F32 && D % 4 == 0 && (long long)B * P * D < (1ll << 31) && al16(w) && al16(bias) && al16(pos) && al16(out) &&
      reinterpret_cast<uintptr_t>(polygon) % 8 == 0) {
    poly_embed_vec_kernel<<<grid_for((long long)B * P * (D / 4), 256), 256, 0, STREAM(stream)>>>(polygon, len, w, bias, pos, reinterpret_cast<float*>(out),
                                                                                              key_mask, B, P, D);
    return check_launch("poly_embed_kernel");
  }
  poly_embed_kernel<<<grid_for((long long)B * P * D, 256), 256, 0, STREAM(stream)>>>(polygon, len, w, bias, pos, out, out_dtype, key_mask, B, P, D);
  return check_launch("poly_embed_kernel");
}

extern "C" int tcavp_masked_mean(const void* x, int in_dtype, const int32_t* len, void* out, int out_dtype, int B, int P, int D,
                                 tcavp_stream_t stream) {
  TCAVP_REQUIRE(B >= 0 && P > 0 && D > 0, "tcavp_masked_mean: bad shape");
  if (B == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && len && out && DT_OK(in_dtype) && DT_OK(out_dtype), "tcavp_masked_mean: bad pointer/dtype");
  const long long n = (long long)B * D;
  masked_mean_kernel<<<(int)((n + 255) / 256), 256, 0, STREAM(stream)>>>(x, in_dtype, len, out, out_dtype, B, P, D);
  return check_launch("masked_mean_kernel");
}

extern "C" int tcavp_split_bf16x3(const float* x, int ldx, void* out, int ldo, long long rows, int cols, tcavp_stream_t stream) {
  TCAVP_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && ldx >= cols && ldx % 4 == 0 && ldo >= 3 * cols && ldo % 4 == 0,
                "tcavp_split_bf16x3: bad shape (cols=%d must be a multiple of 4, ldx=%d, ldo=%d)", cols, ldx, ldo);
  if (rows == 0) return TCAVP_OK;
  TCAVP_REQUIRE(x && out && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0, "tcavp_split_bf16x3: bad pointer");
  split3_kernel<<<grid_for(rows * (cols / 4), 256), 256, 0, STREAM(stream)>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, cols);
  return check_launch("split3_kernel");
}
