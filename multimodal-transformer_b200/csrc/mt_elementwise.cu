// HBM-bound row-wise / element-wise kernels: LayerNorm (reference semantics: unbiased std, eps on std),
// dropout-gradient, casts, activation backward, MSE loss, Adam, transposed weight packs.
// All of them are one pass over their operands with 16-byte vector accesses; warp-shuffle reductions.
#include <type_traits>
#include "mt_ops.cuh"

namespace {

constexpr int LN_WARPS = 8;
// sqrt.approx.f32: maximum relative error 2^-23, no slow-path call (the IEEE sqrtf's range check + subroutine are ~12 instructions a row)
__device__ __forceinline__ float ln_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, NCH float4 chunks per lane (d = NCH * 128).
//   y = a * (x - mean) / (std_unbiased + eps) + b                    MFT/multiTransformer.py:88-91
// ------------------------------------------------------------------------------------------------------
//   DRAW: every warp also draws its share of an attention dropout's keep bits (mt_dropbits.cuh), 32 consecutive words per chunk, the
//   chunks spread over the row iterations so that the hash arithmetic sits between the loads of the next row and their first use
template <typename TY, int NCH, bool DRAW>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(int M, const float* __restrict__ x, const float* __restrict__ a,
                                                               const float* __restrict__ b, float eps, TY* __restrict__ y, size_t pstride,
                                                               const __grid_constant__ MtBitsArgs ba, uint32_t* __restrict__ bits) {
  constexpr int d = NCH * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mt_pdl_gate();
  MtBitsKeys bk;
  uint32_t chunk = 0, n_chunks = 0, chunk_stride = 1, cpi = 0;
  if (DRAW) {
    bk = mt_bits_resolve(ba);
    n_chunks = (bk.n + 31u) >> 5;
    chunk_stride = gridDim.x * gridDim.y * LN_WARPS;
    chunk = (blockIdx.y * gridDim.x + blockIdx.x) * LN_WARPS + warp;
    const uint32_t row_iters = (uint32_t)M * gridDim.y;      // row iterations of the whole launch
    cpi = (n_chunks + row_iters - 1) / row_iters;
  }
  auto draw_chunk = [&]() {
    const uint32_t idx = chunk * 32u + lane;
    if (idx < bk.n) bits[idx] = mt_bits_word(ba, bk, idx);
    chunk += chunk_stride;
  };
  // grouped launch (gridDim.y = modality stacks): group g owns rows [g*M, (g+1)*M) and the parameters at a + g*pstride
  x += (size_t)blockIdx.y * M * d; y += (size_t)blockIdx.y * M * d; a += blockIdx.y * pstride; b += blockIdx.y * pstride;
  float4 av[NCH], bv[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    av[c] = ld4(a + c * 128 + lane * 4);
    bv[c] = ld4(b + c * 128 + lane * 4);
  }
  // software pipeline: the next TWO rows (d <= 256; one beyond) are in flight while this one goes through its two dependent warp
  // reductions -- 32 warps per SM with one 1 KB row each are too few bytes for the HBM latency
  const int stride = gridDim.x * LN_WARPS;
  auto load_row = [&](float4* nv, int row) {
    if (row < M) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) nv[c] = ld4(x + (size_t)row * d + c * 128 + lane * 4);
    }
  };
  auto do_row = [&](float4* nv, int row, int next) {
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = nv[c];
    load_row(nv, next);
    if (DRAW) {
      for (uint32_t k = 0; k < cpi && chunk < n_chunks; ++k) draw_chunk();
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
      q += (v[c].x * v[c].x + v[c].y * v[c].y) + (v[c].z * v[c].z + v[c].w * v[c].w);
    }
    const float sd = ln_sqrt(warp_sum(q) * (1.0f / (d - 1)));
    const float inv = __fdividef(1.0f, sd + eps);      // 2 ulp: the IEEE division's slow path is ~15 of the row's instructions
    TY* yr = y + (size_t)row * d;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float4 o;
      o.x = av[c].x * v[c].x * inv + bv[c].x;
      o.y = av[c].y * v[c].y * inv + bv[c].y;
      o.z = av[c].z * v[c].z * inv + bv[c].z;
      o.w = av[c].w * v[c].w * inv + bv[c].w;
      st4(yr + c * 128 + lane * 4, o);
    }
  };
  int row = blockIdx.x * LN_WARPS + warp;
  constexpr bool DEEP = NCH <= 2;
  float4 n0[NCH];
  load_row(n0, row);
  if (DEEP) {
    float4 n1[NCH];
    load_row(n1, row + stride);
    for (; row < M; row += 2 * stride) {
      do_row(n0, row, row + 2 * stride);
      if (row + stride < M) do_row(n1, row + stride, row + 3 * stride);
    }
  } else {
    for (; row < M; row += stride) do_row(n0, row, row + stride);
  }
  if (DRAW) {      // whatever the row loop left (warps without rows, more chunks than row iterations)
    while (chunk < n_chunks) draw_chunk();
  }
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm backward (checked against autograd in tests/test_oracle_golden.py):
//   g = dy*a;  dx = (g - mean(g) - xc * sum(g*xc) / ((d-1) * s * (s+eps))) / (s+eps)  [+ dres]
//   da += sum_rows dy * xhat ; db += sum_rows dy        (register partials -> smem -> one atomic per column per CTA)
// ------------------------------------------------------------------------------------------------------
//   NEXT: the same pass also emits what the layer below consumes first -- nx_out = dx * dropout_factor(site, row*d + col)
//   in the operand dtype (the gradient through that sublayer's output dropout) and nx_db += colsum(nx_out) (the bias
//   gradient of the Linear that produced the dropped tensor) -- instead of a drop_grad pass + a colsum pass over dx.
//   GM: dtype of the residual-stream gradient, 0 = dres and dx fp32, 1 = both bf16 (inside a bf16-mode stack: it is an activation
//   gradient like dy, and half of this HBM-bound kernel's bytes), 2 = dres bf16, dx fp32 (the stack's bottom: dx leaves the library)
template <int GM> struct LnG { typedef float R; typedef float X; };
template <> struct LnG<1> { typedef bf16 R; typedef bf16 X; };
template <> struct LnG<2> { typedef bf16 R; typedef float X; };
template <typename TY, int NCH, bool NEXT, int GM>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(int M, const float* __restrict__ x, const float* __restrict__ a, float eps,
                                                               const TY* __restrict__ dy, const typename LnG<GM>::R* __restrict__ dres,
                                                               typename LnG<GM>::X* __restrict__ dx, float* __restrict__ da, float* __restrict__ db,
                                                               TY* __restrict__ nx_out, float* __restrict__ nx_db, DropGroups nx_drops,
                                                               size_t pstride) {
  constexpr int d = NCH * 128;
  __shared__ float s_da[d], s_db[d], s_dn[NEXT ? d : 1];
  mt_pdl_gate();
  const DropCfg nx_drop = mt_drop_resolve(nx_drops.d[blockIdx.y]);
  {   // grouped launch: see ln_fwd_kernel; gradients of parameters are laid out like the parameters
    const size_t ro = (size_t)blockIdx.y * M * d, po = blockIdx.y * pstride;
    x += ro; dy += ro; dx += ro; a += po; da += po; db += po;
    if (dres) dres += ro;
    if (NEXT) { nx_out += ro; nx_db += po; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d; i += blockDim.x) { s_da[i] = 0.f; s_db[i] = 0.f; if (NEXT) s_dn[i] = 0.f; }
  __syncthreads();
  float4 av[NCH], pa[NCH], pb[NCH], pn[NEXT ? NCH : 1];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    av[c] = ld4(a + c * 128 + lane * 4);
    pa[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NEXT) pn[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // software pipeline (d <= 256): TWO rows' x / dy / dres are in flight per warp while a row goes through its three dependent warp reductions
  // (16 warps per SM at 110-120 registers: one row in flight is 2-2.5 KB per warp, too few bytes for the HBM latency once the
  // residual-stream gradient is bf16)
  const int stride = gridDim.x * LN_WARPS;
  struct RowBuf {
    float4 v[NCH];
    typename Raw4<TY>::type g[NCH];
    typename Raw4<typename LnG<GM>::R>::type r[NCH];
  };
  auto load_row = [&](RowBuf& rb, int row) {
    if (row < M) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        rb.v[c] = ld4(x + (size_t)row * d + c * 128 + lane * 4);
        rb.g[c] = ld4raw(dy + (size_t)row * d + c * 128 + lane * 4);
        if (dres) rb.r[c] = ld4raw(dres + (size_t)row * d + c * 128 + lane * 4);
      }
    }
  };
  // consumes rb (row `row`), refills it with row `next` before the arithmetic starts
  auto do_row = [&](RowBuf& rb, int row, int next) {
    float4 v[NCH], g[NCH], r[NCH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) { v[c] = rb.v[c]; g[c] = cvt4(rb.g[c]); r[c] = dres ? cvt4(rb.r[c]) : make_float4(0.f, 0.f, 0.f, 0.f); }
    load_row(rb, next);
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
      q += (v[c].x * v[c].x + v[c].y * v[c].y) + (v[c].z * v[c].z + v[c].w * v[c].w);
    }
    const float sd = ln_sqrt(warp_sum(q) * (1.0f / (d - 1)));
    const float inv = __fdividef(1.0f, sd + eps);      // 2 ulp: the IEEE division's slow path is ~15 of the row's instructions
    float gs = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      // parameter-gradient partials use the raw dy
      pb[c].x += g[c].x; pb[c].y += g[c].y; pb[c].z += g[c].z; pb[c].w += g[c].w;
      pa[c].x += g[c].x * v[c].x * inv; pa[c].y += g[c].y * v[c].y * inv;
      pa[c].z += g[c].z * v[c].z * inv; pa[c].w += g[c].w * v[c].w * inv;
      g[c].x *= av[c].x; g[c].y *= av[c].y; g[c].z *= av[c].z; g[c].w *= av[c].w;
      gs += (g[c].x + g[c].y) + (g[c].z + g[c].w);
      dot += (g[c].x * v[c].x + g[c].y * v[c].y) + (g[c].z * v[c].z + g[c].w * v[c].w);
    }
    const float gm = warp_sum(gs) * (1.0f / d);
    dot = warp_sum(dot);
    const float k = __fdividef(dot, (float)(d - 1) * sd * (sd + eps));
    typename LnG<GM>::X* dxr = dx + (size_t)row * d;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float4 o;
      o.x = (g[c].x - gm - v[c].x * k) * inv;
      o.y = (g[c].y - gm - v[c].y * k) * inv;
      o.z = (g[c].z - gm - v[c].z * k) * inv;
      o.w = (g[c].w - gm - v[c].w * k) * inv;
      o.x += r[c].x; o.y += r[c].y; o.z += r[c].z; o.w += r[c].w;
      st4(dxr + c * 128 + lane * 4, o);
      if (NEXT) {
        const uint64_t idx = (uint64_t)row * d + c * 128 + lane * 4;
        float f[4];
        mt_drop_quad(nx_drop, idx, f);
        o.x *= f[0]; o.y *= f[1]; o.z *= f[2]; o.w *= f[3];
        st4(nx_out + (size_t)row * d + c * 128 + lane * 4, o);
        pn[c].x += o.x; pn[c].y += o.y; pn[c].z += o.z; pn[c].w += o.w;
      }
    }
  };
  int row = blockIdx.x * LN_WARPS + warp;
  constexpr bool DEEP = NCH <= 2;      // wider rows carry enough bytes per row (and have no registers for a second buffer)
  RowBuf b0;
  load_row(b0, row);
  if (DEEP) {
    RowBuf b1;
    load_row(b1, row + stride);
    for (; row < M; row += 2 * stride) {
      do_row(b0, row, row + 2 * stride);
      if (row + stride < M) do_row(b1, row + stride, row + 3 * stride);
    }
  } else {
    for (; row < M; row += stride) do_row(b0, row, row + stride);
  }
  if (NEXT) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int col = c * 128 + lane * 4;
      atomicAdd(&s_dn[col + 0], pn[c].x); atomicAdd(&s_dn[col + 1], pn[c].y);
      atomicAdd(&s_dn[col + 2], pn[c].z); atomicAdd(&s_dn[col + 3], pn[c].w);
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int col = c * 128 + lane * 4;
    atomicAdd(&s_da[col + 0], pa[c].x); atomicAdd(&s_da[col + 1], pa[c].y);
    atomicAdd(&s_da[col + 2], pa[c].z); atomicAdd(&s_da[col + 3], pa[c].w);
    atomicAdd(&s_db[col + 0], pb[c].x); atomicAdd(&s_db[col + 1], pb[c].y);
    atomicAdd(&s_db[col + 2], pb[c].z); atomicAdd(&s_db[col + 3], pb[c].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(da + i, s_da[i]);
    atomicAdd(db + i, s_db[i]);
    if (NEXT) atomicAdd(nx_db + i, s_dn[i]);
  }
}

template <typename TO>
__global__ void drop_grad_kernel(size_t n4, const float* __restrict__ g, TO* __restrict__ out, DropCfg drop_in) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = ld4(g + i * 4);
    float f[4];
    mt_drop_quad(drop, (uint64_t)i * 4, f);
    v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
    st4(out + i * 4, v);
  }
}

// out = x + y * dropout_factor   (SublayerConnection residual, stand-alone path)
__global__ void residual_dropout_kernel(size_t n, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                        DropCfg drop_in) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (x ? x[i] : 0.f) + y[i] * mt_drop_factor(drop, i);
}

template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ src, int lds, TD* __restrict__ dst, int ldd, int rows, int cols, int dcols, DropCfg drop_in) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  size_t n = (size_t)rows * dcols;                            // destination columns [cols, dcols) are zeroed, [dcols, ldd) left alone
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / dcols), c = (int)(i % dcols);
    float v = 0.f;
    if (c < cols) v = to_f(src[(size_t)r * lds + c]) * mt_drop_factor(drop, (uint64_t)r * cols + c);
    dst[(size_t)r * ldd + c] = from_f<TD>(v);
  }
}

// vectorised cast2d: 4 columns per thread (cols, dcols, lds, ldd multiples of 4; 16-byte aligned rows); pad columns [cols, dcols) are zeroed
template <typename TS, typename TD>
__global__ void cast2d_vec_kernel(const TS* __restrict__ src, int lds, TD* __restrict__ dst, int ldd, int rows, int cols, int dcols, DropCfg drop_in) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  const int qpr = dcols >> 2;                                 // quads written per destination row
  const size_t n = (size_t)rows * qpr;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / qpr), c = (int)(i - (size_t)r * qpr) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) {
      v = ld4(src + (size_t)r * lds + c);
      if (drop.thresh != 0u) {
        float f[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = mt_drop_factor(drop, (uint64_t)r * cols + c + k);
        v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
      }
    }
    st4(dst + (size_t)r * ldd + c, v);
  }
}

template <typename TG, typename TY, typename TZ>
__global__ void act_bwd_kernel(size_t n, int N, const TG* __restrict__ dy, const TY* __restrict__ y, int act,
                               const float* __restrict__ rowmask, TZ* __restrict__ dz) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float g = to_f(dy[i]);
    if (rowmask) g *= rowmask[i / N];
    if (act == MT_ACT_RELU) g = to_f(y[i]) > 0.f ? g : 0.f;
    else if (act == MT_ACT_TANH) { float t = to_f(y[i]); g *= (1.0f - t * t); }
    dz[i] = from_f<TZ>(g);
  }
}

// The same gradient with the bias gradient fused in: db[n] += sum_m dz[m][n] (fp32 values, before dz is rounded), four columns per
// thread, 8 row phases per CTA, four rows in flight per thread -- replaces an act_bwd pass + a column-sum pass over dz (N % 4 == 0).
template <typename TG, typename TY, typename TZ>
__global__ void __launch_bounds__(256) act_bwd_colsum_kernel(int M, int N, const TG* __restrict__ dy, const TY* __restrict__ y, int act,
                                                             const float* __restrict__ rowmask, TZ* __restrict__ dz, float* __restrict__ db,
                                                             int rows_per_block) {
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + tx * 4;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  auto one = [&](int m, float4 g, float4 t) {
    if (rowmask) { const float r = rowmask[m]; g.x *= r; g.y *= r; g.z *= r; g.w *= r; }
    if (act == MT_ACT_RELU) { g.x = t.x > 0.f ? g.x : 0.f; g.y = t.y > 0.f ? g.y : 0.f; g.z = t.z > 0.f ? g.z : 0.f; g.w = t.w > 0.f ? g.w : 0.f; }
    else if (act == MT_ACT_TANH) { g.x *= 1.0f - t.x * t.x; g.y *= 1.0f - t.y * t.y; g.z *= 1.0f - t.z * t.z; g.w *= 1.0f - t.w * t.w; }
    st4(dz + (size_t)m * N + n, g);
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
  };
  if (n < N) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = m0 + ty;
    for (; m + 56 < m1; m += 64) {
      float4 g[8], t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[k] = ld4(dy + (size_t)(m + 8 * k) * N + n);
        t[k] = act != MT_ACT_NONE ? ld4(y + (size_t)(m + 8 * k) * N + n) : zero;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) one(m + 8 * k, g[k], t[k]);
    }
    for (; m < m1; m += 8) one(m, ld4(dy + (size_t)m * N + n), act != MT_ACT_NONE ? ld4(y + (size_t)m * N + n) : zero);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int q = 1; q < 8; ++q) { float4 o = red[q][tx]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    atomicAdd(db + n, s.x); atomicAdd(db + n + 1, s.y); atomicAdd(db + n + 2, s.z); atomicAdd(db + n + 3, s.w);
  }
}

struct TransposeJobs { TransposeJob j[16]; };
template <typename TD>
__global__ void transpose_pack_kernel(TransposeJobs jobs) {
  const TransposeJob jb = jobs.j[blockIdx.y];
  const int n = jb.R * jb.C;
  TD* dst = reinterpret_cast<TD*>(jb.dst);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int c = e / jb.R, r = e % jb.R;
    dst[(size_t)c * jb.ldd + r] = from_f<TD>(jb.src[(size_t)r * jb.C + c]);
  }
}

__global__ void cast_f2b_kernel(const float* __restrict__ s, bf16* __restrict__ d, size_t n) {
  size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) st4(d + i * 4, ld4(s + i * 4));
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2bfloat16(s[i]);
}
__global__ void cast_b2f_kernel(const bf16* __restrict__ s, float* __restrict__ d, size_t n) {
  size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) st4(d + i * 4, ld4(s + i * 4));
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __bfloat162float(s[i]);
}

__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, size_t n, float inv_norm,
                           const float* __restrict__ inv_norm_dev, float* __restrict__ loss, float* __restrict__ dpred) {
  if (inv_norm_dev) inv_norm = __ldg(inv_norm_dev);
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float e = pred[i] - target[i];
    acc += e * e;
    if (dpred) dpred[i] = 2.0f * e * inv_norm;
  }
  acc = warp_sum(acc);
  __shared__ float s[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = acc;
  __syncthreads();
  if (w == 0) {
    acc = lane < (blockDim.x >> 5) ? s[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(loss, acc * inv_norm);
  }
}

// torch.optim.Adam (L2 weight decay added to the gradient, bias-corrected), MFT/train.py:557
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            const float* __restrict__ lr_dev, const long long* __restrict__ step_dev, bf16* __restrict__ p_lp) {
  if (step_dev) {                        // captured-graph mode: the step count (and lr) live in device memory
    const float t = (float)__ldg(step_dev);
    bc1 = 1.0f - powf(b1, t);
    bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  }
  if (lr_dev) lr = __ldg(lr_dev);
  const float step = lr / bc1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] + wd * pi;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step * (mi / denom);
    p[i] = pi;
    if (p_lp) p_lp[i] = __float2bfloat16(pi);
  }
}

inline int ew_grid(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  size_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename TY>
int ln_fwd_dispatch(int M, int d, const float* x, const float* a, const float* b, float eps, TY* y, cudaStream_t st, int G, size_t pstride,
                    const MtBitsJob* draw) {
  int gx = (M + LN_WARPS - 1) / LN_WARPS;
  const int cap4 = 148 * 4 / (g_mt_tune[MT_TUNE_LN_SHARE] > 1 ? g_mt_tune[MT_TUNE_LN_SHARE] : 1) / G;
  if (gx > cap4) gx = cap4 > 0 ? cap4 : 1;
  const dim3 grid((unsigned)gx, (unsigned)G);
  mt_prof_work(0.0, (double)G * M * d * (4.0 + sizeof(TY)));
  MtBitsArgs ba = {};
  uint32_t* bits = nullptr;
  if (draw) { ba = draw->a; bits = draw->bits; }
#define MT_LNF(NCH)                                                                                                                      \
  do {                                                                                                                                   \
    if (draw) MT_CUDA(mt_launch_dep(MT_PDL_LN_FWD, ln_fwd_kernel<TY, NCH, true>, grid, dim3(LN_WARPS * 32), 0, st, M, x, a, b, eps, y, pstride, ba, bits)); \
    else MT_CUDA(mt_launch_dep(MT_PDL_LN_FWD, ln_fwd_kernel<TY, NCH, false>, grid, dim3(LN_WARPS * 32), 0, st, M, x, a, b, eps, y, pstride, ba, bits));     \
  } while (0)
  switch (d / 128) {
    case 1: MT_LNF(1); break;
    case 2: MT_LNF(2); break;
    case 3: MT_LNF(3); break;
    case 4: MT_LNF(4); break;
    case 6: MT_LNF(6); break;
    case 8: MT_LNF(8); break;
    default: return MT_ERR_UNSUPPORTED;
  }
#undef MT_LNF
  MT_LAUNCH_CHECK();
  return MT_OK;
}

template <typename TY, bool NEXT>
int ln_bwd_dispatch(int M, int d, const float* x, const float* a, float eps, const TY* dy, const void* dres, void* dx, int gm, float* da,
                    float* db, TY* nx_out, float* nx_db, const DropGroups& nx_drop, cudaStream_t st, int G, size_t pstride) {
  int gx = (M + LN_WARPS - 1) / LN_WARPS;
  const int cap2 = 148 * 2 / (g_mt_tune[MT_TUNE_LN_SHARE] > 1 ? g_mt_tune[MT_TUNE_LN_SHARE] : 1) / G;
  if (gx > cap2) gx = cap2 > 0 ? cap2 : 1;
  const dim3 grid((unsigned)gx, (unsigned)G);
  const double gr_b = gm == 0 ? 4.0 : 2.0, gx_b = gm == 1 ? 2.0 : 4.0;
  mt_prof_work(0.0, (double)G * M * d * (4.0 + gx_b + sizeof(TY) + (dres ? gr_b : 0.0) + (NEXT ? sizeof(TY) : 0.0)));
  if (gm != 0 && !std::is_same<TY, bf16>::value) return MT_ERR_ARG;      // a bf16 gradient stream exists in bf16 mode only
#define MT_LNB_G(NCH, GMV)                                                                                                                \
  MT_CUDA(mt_launch_dep(MT_PDL_LN_BWD, ln_bwd_kernel<TY, NCH, NEXT, GMV>, grid, dim3(LN_WARPS * 32), 0, st, M, x, a, eps, dy,                             \
                        (const typename LnG<GMV>::R*)dres, (typename LnG<GMV>::X*)dx, da, db, nx_out, nx_db, nx_drop, pstride))
#define MT_LNB(NCH)                                                                          \
  do {                                                                                       \
    if constexpr (std::is_same<TY, bf16>::value) {                                           \
      if (gm == 1) { MT_LNB_G(NCH, 1); break; }                                              \
      if (gm == 2) { MT_LNB_G(NCH, 2); break; }                                              \
    }                                                                                        \
    MT_LNB_G(NCH, 0);                                                                        \
  } while (0)
  switch (d / 128) {
    case 1: MT_LNB(1); break;
    case 2: MT_LNB(2); break;
    case 3: MT_LNB(3); break;
    case 4: MT_LNB(4); break;
    case 6: MT_LNB(6); break;
    case 8: MT_LNB(8); break;
    default: return MT_ERR_UNSUPPORTED;
  }
#undef MT_LNB
#undef MT_LNB_G
  MT_LAUNCH_CHECK();
  return MT_OK;
}

}  // namespace

int mt_ln_fwd_run(int M, int d, const float* x, const float* a, const float* b, float eps, void* y, bool y_bf16, cudaStream_t st, int G,
                  size_t pstride, const MtBitsJob* draw) {
  if (M <= 0 || d <= 0 || d % 128 != 0 || d > 1024 || !x || !a || !b || !y || G < 1 || G > MT_LN_MAX_GROUPS) return MT_ERR_ARG;
  if (draw && !draw->bits) return MT_ERR_ARG;
  if (y_bf16) return ln_fwd_dispatch<bf16>(M, d, x, a, b, eps, (bf16*)y, st, G, pstride, draw);
  return ln_fwd_dispatch<float>(M, d, x, a, b, eps, (float*)y, st, G, pstride, draw);
}

int mt_ln_bwd_run(int M, int d, const float* x, const float* a, float eps, const void* dy, bool dy_bf16, const void* dres, void* dx,
                  float* da, float* db, cudaStream_t st, const LnBwdNext* nx, int G, size_t pstride, const DropCfg* nx_drops, int gmode) {
  if (M <= 0 || d <= 0 || d % 128 != 0 || d > 1024 || !x || !a || !dy || !dx || !da || !db || G < 1 || G > MT_LN_MAX_GROUPS) return MT_ERR_ARG;
  if (gmode < 0 || gmode > 2 || (gmode != 0 && !dy_bf16)) return MT_ERR_ARG;
  DropGroups dg;
  for (int i = 0; i < MT_LN_MAX_GROUPS; ++i) dg.d[i] = mt_make_drop(0.f, 0, 0);
  if (nx && nx->out) {
    if (!nx->dbias || nx->out == dy) return MT_ERR_ARG;
    for (int i = 0; i < G; ++i) dg.d[i] = nx_drops ? nx_drops[i] : nx->drop;
    if (dy_bf16) return ln_bwd_dispatch<bf16, true>(M, d, x, a, eps, (const bf16*)dy, dres, dx, gmode, da, db, (bf16*)nx->out, nx->dbias, dg, st, G, pstride);
    return ln_bwd_dispatch<float, true>(M, d, x, a, eps, (const float*)dy, dres, dx, gmode, da, db, (float*)nx->out, nx->dbias, dg, st, G, pstride);
  }
  if (dy_bf16) return ln_bwd_dispatch<bf16, false>(M, d, x, a, eps, (const bf16*)dy, dres, dx, gmode, da, db, nullptr, nullptr, dg, st, G, pstride);
  return ln_bwd_dispatch<float, false>(M, d, x, a, eps, (const float*)dy, dres, dx, gmode, da, db, nullptr, nullptr, dg, st, G, pstride);
}

int mt_drop_grad_run(int M, int N, const float* g, void* out, bool out_bf16, DropCfg drop, cudaStream_t st) {
  size_t n = (size_t)M * N;
  if (n == 0 || n % 4 != 0) return MT_ERR_ARG;
  mt_prof_work(0.0, (double)n * (4.0 + (out_bf16 ? 2.0 : 4.0)));
  if (out_bf16) drop_grad_kernel<bf16><<<ew_grid(n / 4, 256), 256, 0, st>>>(n / 4, g, (bf16*)out, drop);
  else drop_grad_kernel<float><<<ew_grid(n / 4, 256), 256, 0, st>>>(n / 4, g, (float*)out, drop);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_cast2d_run(const void* src, bool src_bf16, int lds, void* dst, bool dst_bf16, int ldd, int rows, int cols, DropCfg drop,
                  cudaStream_t st, int dcols) {
  if (dcols < 0) dcols = ldd;                                 // default: zero the whole row tail (a padded staging buffer)
  if (rows <= 0 || cols <= 0 || ldd < dcols || dcols < cols || lds < cols) return MT_ERR_ARG;
  size_t n = (size_t)rows * dcols;
  const size_t ses = src_bf16 ? 2 : 4, des = dst_bf16 ? 2 : 4;
  if (cols % 4 == 0 && dcols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src % (4 * ses)) == 0 && ((uintptr_t)dst % (4 * des)) == 0) {
    const int gridv = ew_grid(n / 4, 256);
    mt_prof_work(0.0, (double)rows * cols * ses + (double)rows * dcols * des);
    if (!src_bf16 && dst_bf16) cast2d_vec_kernel<float, bf16><<<gridv, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, rows, cols, dcols, drop);
    else if (!src_bf16 && !dst_bf16) cast2d_vec_kernel<float, float><<<gridv, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, cols, dcols, drop);
    else if (src_bf16 && dst_bf16) cast2d_vec_kernel<bf16, bf16><<<gridv, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, rows, cols, dcols, drop);
    else cast2d_vec_kernel<bf16, float><<<gridv, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, rows, cols, dcols, drop);
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  int grid = ew_grid(n, 256);
  if (!src_bf16 && dst_bf16) cast2d_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, rows, cols, dcols, drop);
  else if (!src_bf16 && !dst_bf16) cast2d_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, cols, dcols, drop);
  else if (src_bf16 && dst_bf16) cast2d_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, rows, cols, dcols, drop);
  else cast2d_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, rows, cols, dcols, drop);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_act_bwd_run(int M, int N, const void* dy, bool dy_bf16, const void* y, bool y_bf16, int act, const float* rowmask, void* dz,
                   bool dz_bf16, cudaStream_t st, float* db, bool* db_done) {
  size_t n = (size_t)M * N;
  if (n == 0) return MT_ERR_ARG;
  if (act != MT_ACT_NONE && !y) return MT_ERR_ARG;
  if (db_done) *db_done = false;
  if (db && db_done && N % 4 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dz & 15) == 0 && (!y || ((uintptr_t)y & 15) == 0)) {
    // fused bias gradient (overwrites db): one pass instead of act_bwd + colsum
    MT_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
    const int gx = (N + 127) / 128;
    int gy = (148 * 4 + gx - 1) / gx;      // four CTAs per SM: every CTA ends in 128 same-address atomics, twice as many cost more than they hide
    int rpb = ((M + gy - 1) / gy + 63) / 64 * 64;
    if (rpb < 64) rpb = 64;
    const dim3 gridv(gx, (M + rpb - 1) / rpb);
    mt_prof_work(0.0, (double)n * ((dy_bf16 ? 2.0 : 4.0) + (act != MT_ACT_NONE ? (y_bf16 ? 2.0 : 4.0) : 0.0) + (dz_bf16 ? 2.0 : 4.0)));
#define MT_ABC(TG, TYY, TZ) act_bwd_colsum_kernel<TG, TYY, TZ><<<gridv, 256, 0, st>>>(M, N, (const TG*)dy, (const TYY*)y, act, rowmask, (TZ*)dz, db, rpb)
    if (dy_bf16) {
      if (y_bf16) { if (dz_bf16) MT_ABC(bf16, bf16, bf16); else MT_ABC(bf16, bf16, float); }
      else { if (dz_bf16) MT_ABC(bf16, float, bf16); else MT_ABC(bf16, float, float); }
    } else {
      if (y_bf16) { if (dz_bf16) MT_ABC(float, bf16, bf16); else MT_ABC(float, bf16, float); }
      else { if (dz_bf16) MT_ABC(float, float, bf16); else MT_ABC(float, float, float); }
    }
#undef MT_ABC
    MT_LAUNCH_CHECK();
    *db_done = true;
    return MT_OK;
  }
  int grid = ew_grid(n, 256);
#define MT_AB(TG, TYY, TZ) act_bwd_kernel<TG, TYY, TZ><<<grid, 256, 0, st>>>(n, N, (const TG*)dy, (const TYY*)y, act, rowmask, (TZ*)dz)
  if (dy_bf16) {
    if (y_bf16) { if (dz_bf16) MT_AB(bf16, bf16, bf16); else MT_AB(bf16, bf16, float); }
    else { if (dz_bf16) MT_AB(bf16, float, bf16); else MT_AB(bf16, float, float); }
  } else {
    if (y_bf16) { if (dz_bf16) MT_AB(float, bf16, bf16); else MT_AB(float, bf16, float); }
    else { if (dz_bf16) MT_AB(float, float, bf16); else MT_AB(float, float, float); }
  }
#undef MT_AB
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_transpose_pack_run(const TransposeJob* jobs, int n_jobs, bool dst_bf16, cudaStream_t st) {
  if (n_jobs <= 0 || n_jobs > 16) return MT_ERR_ARG;
  TransposeJobs J;
  int mx = 0;
  for (int i = 0; i < n_jobs; ++i) { J.j[i] = jobs[i]; mx = max(mx, jobs[i].R * jobs[i].C); }
  dim3 grid(min((mx + 255) / 256, 64), n_jobs);
  if (dst_bf16) transpose_pack_kernel<bf16><<<grid, 256, 0, st>>>(J);
  else transpose_pack_kernel<float><<<grid, 256, 0, st>>>(J);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

extern "C" {

int mt_layernorm_fwd(int dtype, int M, int d, const float* x, const float* a_2, const float* b_2, float eps, void* y, int y_f32,
                     void* stream) {
  return mt_ln_fwd_run(M, d, x, a_2, b_2, eps, y, dtype == MT_BF16 && !y_f32, (cudaStream_t)stream);
}

int mt_layernorm_bwd(int dtype, int M, int d, const float* x, const float* a_2, float eps, const void* dy, int dy_f32,
                     const float* dres, float* dx, float* da, float* db, void* stream) {
  return mt_ln_bwd_run(M, d, x, a_2, eps, dy, dtype == MT_BF16 && !dy_f32, dres, dx, da, db, (cudaStream_t)stream);
}

int mt_residual_dropout_fwd(const float* x, const float* y, float* out, size_t n, float p, uint64_t seed, uint32_t site, void* stream) {
  if (!y || !out || n == 0) return MT_ERR_ARG;
  residual_dropout_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, y, out, mt_make_drop(p, seed, site));
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_dropout_bwd(const float* g, float* out, size_t n, float p, uint64_t seed, uint32_t site, void* stream) {
  if (!g || !out || n == 0) return MT_ERR_ARG;
  residual_dropout_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(n, nullptr, g, out, mt_make_drop(p, seed, site));
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  if (!src || !dst) return MT_ERR_ARG;
  if (n == 0) return MT_OK;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 7)) return MT_ERR_ALIGN;
  cast_f2b_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_cast_bf16_to_f32(const void* src, float* dst, size_t n, void* stream) {
  if (!src || !dst) return MT_ERR_ARG;
  if (n == 0) return MT_OK;
  if (((uintptr_t)src & 7) || ((uintptr_t)dst & 15)) return MT_ERR_ALIGN;
  cast_b2f_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_mse_loss_fwd_bwd(const float* pred, const float* target, size_t n, float inv_norm, float* loss, float* dpred, void* stream) {
  if (!pred || !target || !loss || n == 0) return MT_ERR_ARG;
  int grid = ew_grid(n, 256);
  if (grid > 148) grid = 148;
  mse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, n, inv_norm, nullptr, loss, dpred);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_mse_loss_fwd_bwd_dev(const float* pred, const float* target, size_t n, const float* inv_norm_dev, float* loss, float* dpred,
                            void* stream) {
  if (!pred || !target || !loss || !inv_norm_dev || n == 0) return MT_ERR_ARG;
  int grid = ew_grid(n, 256);
  if (grid > 148) grid = 148;
  mse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, n, 0.f, inv_norm_dev, loss, dpred);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, void* stream) {
  if (!p || !g || !m || !v || step < 1) return MT_ERR_ARG;
  if (n == 0) return MT_OK;
  float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), nullptr,
                                                                   nullptr, nullptr);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

int mt_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, const float* lr_dev, float lr, float beta1, float beta2,
                     float eps, float weight_decay, const int64_t* step_dev, void* p_lp, void* stream) {
  if (!p || !g || !m || !v || !step_dev) return MT_ERR_ARG;
  if (n == 0) return MT_OK;
  adam_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, lr_dev,
                                                                   reinterpret_cast<const long long*>(step_dev), reinterpret_cast<bf16*>(p_lp));
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

}  // extern "C"
