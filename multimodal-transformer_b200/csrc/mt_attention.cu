// Fused multi-head self-attention core: scores, the reference's QUERY-ROW mask, softmax, dropout, P.V and
// head merge in one kernel; no [B,h,T,T] tensor ever reaches HBM.       attention() MFT/multiTransformer.py:22-34
//
// Mask semantics: mask[b,i] == 0 overwrites the WHOLE score row i with -1e9, so a
// padded query attends uniformly (1/T) to ALL T keys, and valid queries attend to padded keys.
//
// Layout: qkv [B,T,3d] with q | k | v along the last dim; head hd owns columns hd*DK..hd*DK+DK-1 of each.
// This file is the FFMA (fp32-accumulate, exact-softmax) engine used by both dtypes; one thread owns one
// query row (forward, dQ) or one key row (dK/dV), the opposite operand streams through shared memory
// in tiles and is read as warp-wide broadcasts.
#include "mt_ops.cuh"

namespace {

constexpr int AT_THREADS = 128;
constexpr int AT_TILE_ELEMS = 2048;   // floats per staged operand tile (tile rows = 2048 / DK)

template <typename T, int DK>
__device__ __forceinline__ void load_row(float* dst, const T* src) {
#pragma unroll
  for (int c = 0; c < DK; c += 4) {
    float4 f = ld4(src + c);
    dst[c] = f.x; dst[c + 1] = f.y; dst[c + 2] = f.z; dst[c + 3] = f.w;
  }
}

// stage `rows` rows (starting at row j0, clipped at T) of one head's operand into smem[rows][DK]
template <typename T, int DK>
__device__ __forceinline__ void stage_tile(float* s, const T* base /* (b, row 0, column offset applied) */, int ld, int j0, int rows,
                                           int T_) {
  constexpr int V = DK / 4;
  for (int e = threadIdx.x; e < rows * V; e += AT_THREADS) {
    int r = e / V, c4 = (e % V) * 4;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + r < T_) f = ld4(base + (size_t)(j0 + r) * ld + c4);
    *reinterpret_cast<float4*>(s + r * DK + c4) = f;
  }
}

template <typename T, int DK>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(int B, int T_, int d, int h, const T* __restrict__ qkv,
                                                              const float* __restrict__ mask, T* __restrict__ out,
                                                              float* __restrict__ lse, DropCfg drop_in, float scale,
                                                              const int* __restrict__ klen) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int KT = AT_TILE_ELEMS / DK;
  __shared__ __align__(16) float Ks[KT * DK];
  __shared__ __align__(16) float Vs[KT * DK];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int i = blockIdx.x * AT_THREADS + threadIdx.x;
  const bool active = i < T_;
  const int ld = 3 * d;
  const T* qb = qkv + (size_t)b * T_ * ld + hd * DK;
  // ragged inference: narrative b only HAS its first klen[b] windows -- keys beyond them do not exist (mt_b200.h: key_len)
  const int Tk = klen ? max(1, min(klen[b], T_)) : T_;
  float q[DK], o[DK];
  bool masked = false;
  if (active) {
    load_row<T, DK>(q, qb + (size_t)i * ld);
    masked = mask != nullptr && mask[(size_t)b * T_ + i] == 0.f;
  }
#pragma unroll
  for (int c = 0; c < DK; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  const uint64_t drop_row = (uint64_t)((size_t)b * h + hd) * T_ + (uint64_t)(active ? i : 0);      // flat (b, head, query) row
  const uint32_t P2 = (uint32_t)(T_ + 1) >> 1;
  for (int j0 = 0; j0 < Tk; j0 += KT) {
    __syncthreads();
    stage_tile<T, DK>(Ks, qb + d, ld, j0, KT, T_);
    stage_tile<T, DK>(Vs, qb + 2 * d, ld, j0, KT, T_);
    __syncthreads();
    if (!active) continue;
    const int jn = min(KT, Tk - j0);
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * DK);
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 kv = kr[c];
        s = fmaf(q[4 * c], kv.x, s); s = fmaf(q[4 * c + 1], kv.y, s);
        s = fmaf(q[4 * c + 2], kv.z, s); s = fmaf(q[4 * c + 3], kv.w, s);
      }
      s = masked ? -1e9f : s * scale;
      if (s > m) {
        float corr = expf(m - s);
        l *= corr;
#pragma unroll
        for (int c = 0; c < DK; ++c) o[c] *= corr;
        m = s;
      }
      float p = expf(s - m);
      l += p;
      p *= mt_attn_drop_factor(drop, drop_row, P2, (uint32_t)(j0 + j));
      const float4* vr = reinterpret_cast<const float4*>(Vs + j * DK);
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 vv = vr[c];
        o[4 * c] = fmaf(p, vv.x, o[4 * c]); o[4 * c + 1] = fmaf(p, vv.y, o[4 * c + 1]);
        o[4 * c + 2] = fmaf(p, vv.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, vv.w, o[4 * c + 3]);
      }
    }
  }
  if (!active) return;
  const float inv = 1.0f / l;
  T* orow = out + ((size_t)b * T_ + i) * d + hd * DK;
#pragma unroll
  for (int c = 0; c < DK; c += 4) st4(orow + c, make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv));
  // masked rows: store the lse of a row of zeros so that backward can use p = exp(0 - lse) = 1/T exactly
  if (lse) lse[((size_t)b * h + hd) * T_ + i] = masked ? logf((float)T_) : m + logf(l);
}

// dQ and D = rowsum(dO * O):  thread = query row
template <typename T, int DK>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dq_kernel(int B, int T_, int d, int h, const T* __restrict__ qkv,
                                                                 const float* __restrict__ mask, const T* __restrict__ out,
                                                                 const float* __restrict__ lse, const T* __restrict__ dout,
                                                                 T* __restrict__ dqkv, float* __restrict__ Dws, DropCfg drop_in,
                                                                 float scale) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int KT = AT_TILE_ELEMS / DK;
  __shared__ __align__(16) float Ks[KT * DK];
  __shared__ __align__(16) float Vs[KT * DK];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int i = blockIdx.x * AT_THREADS + threadIdx.x;
  const bool active = i < T_;
  const int ld = 3 * d;
  const T* qb = qkv + (size_t)b * T_ * ld + hd * DK;
  float q[DK], go[DK], dq[DK];
  bool masked = false;
  float D = 0.f, L = 0.f;
  if (active) {
    load_row<T, DK>(q, qb + (size_t)i * ld);
    load_row<T, DK>(go, dout + ((size_t)b * T_ + i) * d + hd * DK);
    float ov[DK];
    load_row<T, DK>(ov, out + ((size_t)b * T_ + i) * d + hd * DK);
#pragma unroll
    for (int c = 0; c < DK; ++c) D = fmaf(go[c], ov[c], D);
    masked = mask != nullptr && mask[(size_t)b * T_ + i] == 0.f;
    L = lse[((size_t)b * h + hd) * T_ + i];
    Dws[((size_t)b * h + hd) * T_ + i] = D;
  }
#pragma unroll
  for (int c = 0; c < DK; ++c) dq[c] = 0.f;
  const uint64_t drop_row = (uint64_t)((size_t)b * h + hd) * T_ + (uint64_t)(active ? i : 0);      // flat (b, head, query) row
  const uint32_t P2 = (uint32_t)(T_ + 1) >> 1;
  for (int j0 = 0; j0 < T_; j0 += KT) {
    __syncthreads();
    stage_tile<T, DK>(Ks, qb + d, ld, j0, KT, T_);
    stage_tile<T, DK>(Vs, qb + 2 * d, ld, j0, KT, T_);
    __syncthreads();
    if (!active || masked) continue;     // masked_fill blocks the score gradient of the whole row
    const int jn = min(KT, T_ - j0);
    for (int j = 0; j < jn; ++j) {
      float s = 0.f, dp = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * DK);
      const float4* vr = reinterpret_cast<const float4*>(Vs + j * DK);
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 kv = kr[c], vv = vr[c];
        s = fmaf(q[4 * c], kv.x, s); s = fmaf(q[4 * c + 1], kv.y, s);
        s = fmaf(q[4 * c + 2], kv.z, s); s = fmaf(q[4 * c + 3], kv.w, s);
        dp = fmaf(go[4 * c], vv.x, dp); dp = fmaf(go[4 * c + 1], vv.y, dp);
        dp = fmaf(go[4 * c + 2], vv.z, dp); dp = fmaf(go[4 * c + 3], vv.w, dp);
      }
      float p = expf(s * scale - L);
      dp *= mt_attn_drop_factor(drop, drop_row, P2, (uint32_t)(j0 + j));
      float ds = p * (dp - D) * scale;
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 kv = kr[c];
        dq[4 * c] = fmaf(ds, kv.x, dq[4 * c]); dq[4 * c + 1] = fmaf(ds, kv.y, dq[4 * c + 1]);
        dq[4 * c + 2] = fmaf(ds, kv.z, dq[4 * c + 2]); dq[4 * c + 3] = fmaf(ds, kv.w, dq[4 * c + 3]);
      }
    }
  }
  if (!active) return;
  T* r = dqkv + ((size_t)b * T_ + i) * ld + hd * DK;
#pragma unroll
  for (int c = 0; c < DK; c += 4) st4(r + c, make_float4(dq[c], dq[c + 1], dq[c + 2], dq[c + 3]));
}

// dK, dV: thread = key row; queries (q, dO, lse, D, mask) stream through shared memory
template <typename T, int DK>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dkv_kernel(int B, int T_, int d, int h, const T* __restrict__ qkv,
                                                                  const float* __restrict__ mask, const float* __restrict__ lse,
                                                                  const T* __restrict__ dout, T* __restrict__ dqkv,
                                                                  const float* __restrict__ Dws, DropCfg drop_in, float scale) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int QT = AT_TILE_ELEMS / DK;
  __shared__ __align__(16) float Qs[QT * DK];
  __shared__ __align__(16) float Gs[QT * DK];
  __shared__ float Ls[QT], Ds[QT], Ms[QT];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int j = blockIdx.x * AT_THREADS + threadIdx.x;
  const bool active = j < T_;
  const int ld = 3 * d;
  const T* qb = qkv + (size_t)b * T_ * ld + hd * DK;
  float k[DK], v[DK], dk[DK], dv[DK];
  if (active) {
    load_row<T, DK>(k, qb + d + (size_t)j * ld);
    load_row<T, DK>(v, qb + 2 * d + (size_t)j * ld);
  }
#pragma unroll
  for (int c = 0; c < DK; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
  const size_t bh = (size_t)b * h + hd;
  for (int i0 = 0; i0 < T_; i0 += QT) {
    __syncthreads();
    stage_tile<T, DK>(Qs, qb, ld, i0, QT, T_);
    stage_tile<T, DK>(Gs, dout + (size_t)b * T_ * d + hd * DK, d, i0, QT, T_);
    for (int e = threadIdx.x; e < QT; e += AT_THREADS) {
      int i = i0 + e;
      bool ok = i < T_;
      Ls[e] = ok ? lse[bh * T_ + i] : 0.f;
      Ds[e] = ok ? Dws[bh * T_ + i] : 0.f;
      Ms[e] = ok ? ((mask != nullptr && mask[(size_t)b * T_ + i] == 0.f) ? 0.f : 1.f) : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int in = min(QT, T_ - i0);
    for (int e = 0; e < in; ++e) {
      const float4* qr = reinterpret_cast<const float4*>(Qs + e * DK);
      const float4* gr = reinterpret_cast<const float4*>(Gs + e * DK);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 qv = qr[c], gv = gr[c];
        s = fmaf(qv.x, k[4 * c], s); s = fmaf(qv.y, k[4 * c + 1], s);
        s = fmaf(qv.z, k[4 * c + 2], s); s = fmaf(qv.w, k[4 * c + 3], s);
        dp = fmaf(gv.x, v[4 * c], dp); dp = fmaf(gv.y, v[4 * c + 1], dp);
        dp = fmaf(gv.z, v[4 * c + 2], dp); dp = fmaf(gv.w, v[4 * c + 3], dp);
      }
      const float valid = Ms[e];
      // masked query rows: scores are a constant row -> p = 1/T (lse holds log T), and no score gradient
      float p = expf((valid != 0.f ? s * scale : 0.f) - Ls[e]);
      float f = mt_attn_drop_factor(drop, (uint64_t)(bh * T_ + (size_t)(i0 + e)), (uint32_t)(T_ + 1) >> 1, (uint32_t)j);
      float pd = p * f;
      float ds = valid * p * (dp * f - Ds[e]) * scale;
#pragma unroll
      for (int c = 0; c < DK / 4; ++c) {
        float4 qv = qr[c], gv = gr[c];
        dv[4 * c] = fmaf(pd, gv.x, dv[4 * c]); dv[4 * c + 1] = fmaf(pd, gv.y, dv[4 * c + 1]);
        dv[4 * c + 2] = fmaf(pd, gv.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(pd, gv.w, dv[4 * c + 3]);
        dk[4 * c] = fmaf(ds, qv.x, dk[4 * c]); dk[4 * c + 1] = fmaf(ds, qv.y, dk[4 * c + 1]);
        dk[4 * c + 2] = fmaf(ds, qv.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(ds, qv.w, dk[4 * c + 3]);
      }
    }
  }
  if (!active) return;
  T* rk = dqkv + ((size_t)b * T_ + j) * ld + d + hd * DK;
  T* rv = dqkv + ((size_t)b * T_ + j) * ld + 2 * d + hd * DK;
#pragma unroll
  for (int c = 0; c < DK; c += 4) {
    st4(rk + c, make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]));
    st4(rv + c, make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]));
  }
}

// p_attn materialisation (debug / MultiHeadedAttention.attn): one warp per (b, head, query row)
template <typename T>
__global__ void attn_probs_kernel(int B, int T_, int d, int h, int dk, const T* __restrict__ qkv, const float* __restrict__ mask,
                                  float* __restrict__ probs, float scale) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (size_t)B * h * T_) return;
  const int i = (int)(row % T_), hd = (int)((row / T_) % h), b = (int)(row / ((size_t)T_ * h));
  const int ld = 3 * d;
  const T* qb = qkv + (size_t)b * T_ * ld + hd * dk;
  const bool masked = mask != nullptr && mask[(size_t)b * T_ + i] == 0.f;
  float* pr = probs + row * T_;
  float mx = -INFINITY;
  for (int j = lane; j < T_; j += 32) {
    float s = 0.f;
    for (int c = 0; c < dk; ++c) s = fmaf(to_f(qb[(size_t)i * ld + c]), to_f(qb[(size_t)j * ld + d + c]), s);
    s = masked ? -1e9f : s * scale;
    pr[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < T_; j += 32) { float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < T_; j += 32) pr[j] *= inv;
}

template <typename T>
int fwd_dispatch(int B, int T_, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                 const int* klen) {
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  dim3 grid((T_ + AT_THREADS - 1) / AT_THREADS, h, B);
  mt_prof_work(4.0 * B * (double)T_ * T_ * d, (double)B * T_ * d * 4.0 * sizeof(T));
  switch (dk) {
    case 16: attn_fwd_kernel<T, 16><<<grid, AT_THREADS, 0, st>>>(B, T_, d, h, (const T*)qkv, mask, (T*)out, lse, drop, scale, klen); break;
    case 32: attn_fwd_kernel<T, 32><<<grid, AT_THREADS, 0, st>>>(B, T_, d, h, (const T*)qkv, mask, (T*)out, lse, drop, scale, klen); break;
    case 64: attn_fwd_kernel<T, 64><<<grid, AT_THREADS, 0, st>>>(B, T_, d, h, (const T*)qkv, mask, (T*)out, lse, drop, scale, klen); break;
    default: return MT_ERR_UNSUPPORTED;
  }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

template <typename T>
int bwd_dispatch(int B, int T_, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                 void* dqkv, DropCfg drop, float* Dws, cudaStream_t st) {
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  dim3 grid((T_ + AT_THREADS - 1) / AT_THREADS, h, B);
#define MT_BWD(DK)                                                                                                             \
  mt_prof_work(6.0 * B * (double)T_ * T_ * d, (double)B * T_ * d * 6.0 * sizeof(T));                                            \
  attn_bwd_dq_kernel<T, DK><<<grid, AT_THREADS, 0, st>>>(B, T_, d, h, (const T*)qkv, mask, (const T*)out, lse, (const T*)dout,  \
                                                         (T*)dqkv, Dws, drop, scale);                                          \
  MT_LAUNCH_CHECK();                                                                                                           \
  mt_prof_work(8.0 * B * (double)T_ * T_ * d, (double)B * T_ * d * 6.0 * sizeof(T));                                            \
  attn_bwd_dkv_kernel<T, DK><<<grid, AT_THREADS, 0, st>>>(B, T_, d, h, (const T*)qkv, mask, lse, (const T*)dout, (T*)dqkv, Dws, \
                                                          drop, scale);                                                        \
  MT_LAUNCH_CHECK();
  switch (dk) {
    case 16: { MT_BWD(16) } break;
    case 32: { MT_BWD(32) } break;
    case 64: { MT_BWD(64) } break;
    default: return MT_ERR_UNSUPPORTED;
  }
#undef MT_BWD
  return MT_OK;
}

int check_shape(int B, int T_, int d, int h) {
  if (B <= 0 || T_ <= 0 || d <= 0 || h <= 0 || d % h != 0) return MT_ERR_ARG;
  if (B > 65535 || h > 65535) return MT_ERR_ARG;
  if (d % 4 != 0) return MT_ERR_ALIGN;
  return MT_OK;
}

}  // namespace

static int g_attn_force_ffma = 0;

// G stacks back to back in one call: one launch on the tcgen05 engine, one call per stack otherwise
int mt_attn_group_fwd_run(int dtype, int G, int B, int T_, int d, int h, const void* qkv, const float* mask, void* out, float* lse,
                          const DropCfg* drops, cudaStream_t st, const int* klen, const uint32_t* dbits) {
  MT_TRY(check_shape(B, T_, d, h));
  if (!qkv || !out || G < 1 || G > 4) return MT_ERR_ARG;
  if ((G > 1 || dbits) && dtype == MT_BF16 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() && mt_attn_tc_supported(B, T_, d, h) &&
      !(((uintptr_t)qkv | (uintptr_t)out) & 15)) {
    const int rc = mt_attn_tc_fwd_run(B, T_, d, h, qkv, mask, out, lse, drops[0], st, klen, G, drops, dbits);
    if (rc != MT_ERR_UNSUPPORTED) return rc;
  }
  const size_t es = dtype == MT_BF16 ? 2 : 4, M = (size_t)B * T_;
  for (int g = 0; g < G; ++g)
    MT_TRY(mt_attn_fwd_run(dtype, B, T_, d, h, (const char*)qkv + g * M * 3 * d * es, mask, (char*)out + g * M * d * es,
                           lse ? lse + (size_t)g * B * h * T_ : nullptr, drops[g], st, klen));
  return MT_OK;
}

bool mt_attn_group_bwd_uses_tc(int dtype, int G, int B, int T_, int d, int h, const void* qkv, const void* out, const void* dout,
                               const void* dqkv, const float* Dws, const float* dbias) {
  return G >= 1 && G <= 4 && dtype == MT_BF16 && dbias && h <= 8 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() &&
         mt_attn_tc_supported(B, T_, d, h) && (long long)G * B * T_ <= 0x7fffffffLL / (3LL * d) &&
         !(((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv | (uintptr_t)Dws) & 15);
}

int mt_attn_group_bwd_run(int dtype, int G, int B, int T_, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                          const void* dout, void* dqkv, const DropCfg* drops, float* Dws, cudaStream_t st, float* dbias, size_t dbias_gstride,
                          int d_ready, const uint32_t* dbits) {
  MT_TRY(check_shape(B, T_, d, h));
  if (!qkv || !out || !lse || !dout || !dqkv || !Dws || G < 1 || G > 4) return MT_ERR_ARG;
  if ((G > 1 || d_ready || dbits) && mt_attn_group_bwd_uses_tc(dtype, G, B, T_, d, h, qkv, out, dout, dqkv, Dws, dbias)) {
    const int rc = mt_attn_tc_bwd_run(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drops[0], dbias, Dws, st, G, drops, dbias_gstride, d_ready, dbits);
    if (rc != MT_ERR_UNSUPPORTED || d_ready) return rc;
  }
  if (d_ready) return MT_ERR_UNSUPPORTED;      // the caller skipped the full preparation: only the tcgen05 engine can continue
  const size_t es = dtype == MT_BF16 ? 2 : 4, M = (size_t)B * T_;
  for (int g = 0; g < G; ++g)
    MT_TRY(mt_attn_bwd_run(dtype, B, T_, d, h, (const char*)qkv + g * M * 3 * d * es, mask, (const char*)out + g * M * d * es,
                           lse + (size_t)g * B * h * T_, (const char*)dout + g * M * d * es, (char*)dqkv + g * M * 3 * d * es, drops[g],
                           Dws + (size_t)g * mt_attn_bwd_ws_floats(B, T_, h), st, dbias ? dbias + g * dbias_gstride : nullptr));
  return MT_OK;
}

int mt_attn_fwd_run(int dtype, int B, int T_, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop,
                    cudaStream_t st, const int* klen) {
  MT_TRY(check_shape(B, T_, d, h));
  if (!qkv || !out) return MT_ERR_ARG;
  if (dtype == MT_BF16 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() && mt_attn_tc_supported(B, T_, d, h) &&
      !(((uintptr_t)qkv | (uintptr_t)out) & 15))
    return mt_attn_tc_fwd_run(B, T_, d, h, qkv, mask, out, lse, drop, st, klen);
  if (dtype == MT_BF16 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() && !klen && T_ > 128 && mt_attn_flash_supported(B, T_, d, h) &&
      !(((uintptr_t)qkv | (uintptr_t)out) & 15))
    return mt_attn_flash_fwd_run(B, T_, d, h, qkv, mask, out, lse, drop, st);
  if (dtype == MT_BF16 && !g_attn_force_ffma && mt_attn_mma_supported(B, T_, d, h)) return mt_attn_mma_fwd_run(B, T_, d, h, qkv, mask, out, lse, drop, st, klen);
  if (dtype == MT_BF16) return fwd_dispatch<bf16>(B, T_, d, h, qkv, mask, out, lse, drop, st, klen);
  return fwd_dispatch<float>(B, T_, d, h, qkv, mask, out, lse, drop, st, klen);
}

int mt_attn_bwd_run(int dtype, int B, int T_, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                    const void* dout, void* dqkv, DropCfg drop, float* Dws, cudaStream_t st, float* dbias) {
  MT_TRY(check_shape(B, T_, d, h));
  if (!qkv || !out || !lse || !dout || !dqkv || !Dws) return MT_ERR_ARG;
  bool done = false;
  // tcgen05 / TMEM engine: T <= 128, 32-wide heads; it also produces the QKV bias gradient unless h > 8
  if (dtype == MT_BF16 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() && mt_attn_tc_supported(B, T_, d, h) &&
      !(((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv | (uintptr_t)Dws) & 15)) {
    float* db = h <= 8 ? dbias : nullptr;
    MT_TRY(mt_attn_tc_bwd_run(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drop, db, Dws, st));
    done = db != nullptr;
  } else if (dtype == MT_BF16 && !g_attn_force_ffma && !g_mt_attn_no_tc && !mt_attn_force_tiled_on() && T_ > 128 && mt_attn_flash_supported(B, T_, d, h) &&
             !(((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv | (uintptr_t)Dws) & 15)) {
    MT_TRY(mt_attn_flash_bwd_run(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drop, Dws, st));      // long sequences, 64-wide heads: tcgen05 flash
  } else if (dtype == MT_BF16 && !g_attn_force_ffma && mt_attn_mma_supported(B, T_, d, h))
    MT_TRY(mt_attn_mma_bwd_run(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drop, st, dbias, &done));
  else if (dtype == MT_BF16) MT_TRY(bwd_dispatch<bf16>(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drop, Dws, st));
  else MT_TRY(bwd_dispatch<float>(B, T_, d, h, qkv, mask, out, lse, dout, dqkv, drop, Dws, st));
  if (dbias && !done) MT_TRY(mt_colsum_run(dtype == MT_BF16, B * T_, 3 * d, dqkv, 3 * d, dbias, 1, st));
  return MT_OK;
}

extern "C" {

/* test hook: run bf16 attention on the FFMA engine instead of the tensor-core engine; returns the previous setting */
int mt_attention_force_ffma(int on) { int old = g_attn_force_ffma; g_attn_force_ffma = on; return old; }

int mt_attention_fwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, float p_drop,
                     uint64_t seed, uint32_t site, void* stream) {
  return mt_attn_fwd_run(dtype, B, T, d, h, qkv, mask, out, lse, mt_make_drop(p_drop, seed, site), (cudaStream_t)stream);
}

size_t mt_attention_bwd_ws_bytes(int B, int T, int h) { return mt_attn_bwd_ws_floats(B, T, h) * sizeof(float); }

int mt_attention_ragged_fwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, const int* key_len, void* out,
                            void* stream) {
  if (!qkv || !out || !key_len || B <= 0 || T <= 0 || d <= 0 || h <= 0 || d % h != 0) return MT_ERR_ARG;
  return mt_attn_fwd_run(dtype, B, T, d, h, qkv, mask, out, nullptr, mt_make_drop(0.f, 0, 0), (cudaStream_t)stream, key_len);
}

int mt_attention_bwd(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                     const void* dout, void* dqkv, float p_drop, uint64_t seed, uint32_t site, void* ws, size_t ws_bytes,
                     void* stream) {
  if (!ws || ws_bytes < mt_attention_bwd_ws_bytes(B, T, h)) return MT_ERR_WS;
  return mt_attn_bwd_run(dtype, B, T, d, h, qkv, mask, out, lse, dout, dqkv, mt_make_drop(p_drop, seed, site), (float*)ws,
                         (cudaStream_t)stream);
}

int mt_attention_probs(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, float* probs, void* stream) {
  MT_TRY(check_shape(B, T, d, h));
  if (!qkv || !probs) return MT_ERR_ARG;
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  size_t rows = (size_t)B * h * T;
  int grid = (int)((rows + 7) / 8);
  if (dtype == MT_BF16) attn_probs_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, T, d, h, dk, (const bf16*)qkv, mask, probs, scale);
  else attn_probs_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(B, T, d, h, dk, (const float*)qkv, mask, probs, scale);
  MT_LAUNCH_CHECK_S((cudaStream_t)stream);
  return MT_OK;
}

}  // extern "C"
