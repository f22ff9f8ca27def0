// Tensor-core attention core for bf16 mode (forward + backward), flash-style: no [T,T] tensor leaves the SM.
// Same semantics as mt_attention.cu (the fp32 FFMA engine): query-ROW mask (-1e9 over the whole row), softmax,
// dropout on the probabilities (counter-based, element index ((b*h+hd)*T + i)*T + j), P.V, head merge.
//
// Tiles are d_k = 16/32/64 wide and T <= a few hundred long, far below the 128x128xK footprint of a tcgen05 tile, so
// the contractions run on warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate): a warp owns 16 query rows
// (forward, dQ) or 16 key rows (dK/dV); the opposite operand streams through shared memory in 64-row tiles and is
// fetched with ldmatrix (transposed where the contraction runs over the row index).  Softmax statistics, the
// probabilities and all accumulators stay in registers; P / dS are re-packed from accumulator layout to A-operand
// layout without touching shared memory.
#include "mt_mma.cuh"

namespace {

using namespace mtmma;

constexpr int TILE = 64;          // rows of the streamed operand per shared-memory tile
constexpr int WARPS = 4;          // 16 owned rows each -> 64 owned rows per CTA
constexpr int THREADS = WARPS * 32;

// stage TILE rows [row0, row0 + TILE) of a [rows x DK] global operand into smem[TILE][DK + 8] (zero beyond `rows`)
template <int DK>
__device__ __forceinline__ void stage(bf16* s, const bf16* base, int ld, int row0, int rows) {
  constexpr int LD = DK + 8, V = DK / 8;
  for (int e = threadIdx.x; e < TILE * V; e += THREADS) {
    const int r = e / V, c = (e % V) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < rows) v = *reinterpret_cast<const uint4*>(base + (size_t)(row0 + r) * ld + c);
    *reinterpret_cast<uint4*>(s + r * LD + c) = v;
  }
}

// acc[nt] (nt = 0..7: 8-column tiles over the TILE streamed rows) = A[16 x DK] . S[TILE x DK]^T
template <int DK>
__device__ __forceinline__ void mma_a_bt(float (*acc)[4], const uint32_t (*a)[4], const bf16* s, int lane) {
  constexpr int LD = DK + 8;
#pragma unroll
  for (int nt = 0; nt < TILE / 8; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int kp = 0; kp < DK / 32; ++kp) {       // one ldmatrix.x4 covers two k16 steps
      uint32_t b[4];
      ldsm4(b, s + (nt * 8 + (lane & 7)) * LD + kp * 32 + (lane >> 3) * 8);
      mma16816(acc[nt], a[2 * kp], b[0], b[1]);
      mma16816(acc[nt], a[2 * kp + 1], b[2], b[3]);
    }
    if (DK % 32 != 0) {                           // DK == 16: a single k16 step
      uint32_t b[4];
      ldsm4(b, s + (nt * 8 + (lane & 7)) * LD + ((lane >> 3) & 1) * 8);
      mma16816(acc[nt], a[DK / 32 * 2], b[0], b[1]);
    }
  }
}

// o[DK/8] += P[16 x TILE] . S[TILE x DK]   with P given as packed A fragments pa[TILE/16][4]
template <int DK>
__device__ __forceinline__ void mma_p_s(float (*o)[4], const uint32_t (*pa)[4], const bf16* s, int lane) {
  constexpr int LD = DK + 8;
#pragma unroll
  for (int kk = 0; kk < TILE / 16; ++kk) {
#pragma unroll
    for (int nd = 0; nd < DK / 16; ++nd) {
      uint32_t b[4];
      const int mi = lane >> 3, r = lane & 7;
      ldsm4t(b, s + (kk * 16 + (mi & 1) * 8 + r) * LD + nd * 16 + (mi >> 1) * 8);
      mma16816(o[2 * nd], pa[kk], b[0], b[1]);
      mma16816(o[2 * nd + 1], pa[kk], b[2], b[3]);
    }
  }
}

template <int DK>
__global__ void __launch_bounds__(THREADS) attn_mma_fwd_kernel(int B, int T, int d, int h, const bf16* __restrict__ qkv,
                                                               const float* __restrict__ mask, bf16* __restrict__ out,
                                                               float* __restrict__ lse, DropCfg drop_in, float scale,
                                                               const int* __restrict__ klen) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int LD = DK + 8;
  __shared__ __align__(16) bf16 Ks[TILE * LD];
  __shared__ __align__(16) bf16 Vs[TILE * LD];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int Tk = klen ? max(1, min(klen[b], T)) : T;      // ragged inference: keys beyond the narrative's own length do not exist
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * d;
  const bf16* qb = qkv + (size_t)b * T * ld + hd * DK;
  const int i0 = blockIdx.x * (WARPS * 16) + warp * 16;
  const int r0 = i0 + (lane >> 2), r1 = r0 + 8;
  uint32_t qa[DK / 16][4];
  load_a_frags<DK>(qa, qb, ld, i0, T, lane);
  const bool mk0 = mask != nullptr && r0 < T && mask[(size_t)b * T + r0] == 0.f;
  const bool mk1 = mask != nullptr && r1 < T && mask[(size_t)b * T + r1] == 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[DK / 8][4];
#pragma unroll
  for (int i = 0; i < DK / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const uint64_t bh = (uint64_t)b * h + hd;
  const uint64_t drow0 = bh * T + (uint64_t)min(r0, T - 1), drow1 = bh * T + (uint64_t)min(r1, T - 1);
  const uint32_t P2 = (uint32_t)(T + 1) >> 1;

  for (int j0 = 0; j0 < Tk; j0 += TILE) {
    __syncthreads();
    stage<DK>(Ks, qb + d, ld, j0, T);
    stage<DK>(Vs, qb + 2 * d, ld, j0, T);
    __syncthreads();
    float s[TILE / 8][4];
    mma_a_bt<DK>(s, qa, Ks, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < TILE / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + nt * 8 + 2 * (lane & 3) + e;
        const bool in = j < Tk;
        float v0 = mk0 ? -1e9f : s[nt][e] * scale, v1 = mk1 ? -1e9f : s[nt][2 + e] * scale;
        v0 = in ? v0 : -INFINITY; v1 = in ? v1 : -INFINITY;
        s[nt][e] = v0; s[nt][2 + e] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float c0 = __expf(m0 - n0), c1 = __expf(m1 - n1);
    m0 = n0; m1 = n1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    uint32_t pa[TILE / 16][4];
#pragma unroll
    for (int nt = 0; nt < TILE / 8; ++nt) {
      float p[4], f[4];
      mt_attn_drop_pair(drop, drow0, P2, (uint32_t)(j0 + nt * 8 + 2 * (lane & 3)), f[0], f[1]);
      mt_attn_drop_pair(drop, drow1, P2, (uint32_t)(j0 + nt * 8 + 2 * (lane & 3)), f[2], f[3]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = __expf(s[nt][e] - m0), p1 = __expf(s[nt][2 + e] - m1);
        l0 += p0; l1 += p1;
        p[e] = p0 * f[e];
        p[2 + e] = p1 * f[2 + e];
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack2(p[0], p[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(p[2], p[3]);
    }
    mma_p_s<DK>(o, pa, Vs, lane);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const int c = 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < DK / 8; ++i) {
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r0) * d + hd * DK + i * 8 + c) = pack2(o[i][0] * inv0, o[i][1] * inv0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r1) * d + hd * DK + i * 8 + c) = pack2(o[i][2] * inv1, o[i][3] * inv1);
  }
  if (lse && (lane & 3) == 0) {
    if (r0 < T) lse[bh * T + r0] = mk0 ? logf((float)T) : m0 + logf(l0);
    if (r1 < T) lse[bh * T + r1] = mk1 ? logf((float)T) : m1 + logf(l1);
  }
}

// Backward.  Phase A: the CTA's 64 query rows -> dQ (streams K, V).  Phase B: the CTA's 64 key rows -> dK, dV
// (streams Q, dO and recomputes D = rowsum(dO * O) for the streamed queries).
template <int DK>
__global__ void __launch_bounds__(THREADS) attn_mma_bwd_kernel(int B, int T, int d, int h, const bf16* __restrict__ qkv,
                                                               const float* __restrict__ mask, const bf16* __restrict__ out,
                                                               const float* __restrict__ lse, const bf16* __restrict__ dout,
                                                               bf16* __restrict__ dqkv, DropCfg drop_in, float scale) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int LD = DK + 8;
  __shared__ __align__(16) bf16 S0[TILE * LD];
  __shared__ __align__(16) bf16 S1[TILE * LD];
  __shared__ float Ls[TILE], Ds[TILE], Vd[TILE];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * d;
  const bf16* qb = qkv + (size_t)b * T * ld + hd * DK;
  const bf16* gob = dout + (size_t)b * T * d + hd * DK;
  const bf16* ob = out + (size_t)b * T * d + hd * DK;
  const uint64_t bh = (uint64_t)b * h + hd;
  const int w0 = blockIdx.x * (WARPS * 16) + warp * 16;       // first owned row of this warp (query in A, key in B)
  const int r0 = w0 + (lane >> 2), r1 = r0 + 8;
  const int c = 2 * (lane & 3);

  // ---------------- phase A: dQ -------------------------------------------------------------------------
  {
    uint32_t qa[DK / 16][4], ga[DK / 16][4];
    load_a_frags<DK>(qa, qb, ld, w0, T, lane);
    load_a_frags<DK>(ga, gob, d, w0, T, lane);
    // D = rowsum(dO * O) of the two owned rows: each lane holds 2*DK/8... elements of the row; reduce over the quad
    float D0 = 0.f, D1 = 0.f;
    {
      uint32_t oa[DK / 16][4];
      load_a_frags<DK>(oa, ob, d, w0, T, lane);
#pragma unroll
      for (int ks = 0; ks < DK / 16; ++ks) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 g2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&ga[ks][q]));
          float2 o2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&oa[ks][q]));
          const float t = g2.x * o2.x + g2.y * o2.y;
          if (q & 1) D1 += t; else D0 += t;
        }
      }
      D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
      D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
    }
    const bool v0 = r0 < T && !(mask != nullptr && mask[(size_t)b * T + r0] == 0.f);
    const bool v1 = r1 < T && !(mask != nullptr && mask[(size_t)b * T + r1] == 0.f);
    const float L0 = r0 < T ? lse[bh * T + r0] : 0.f, L1 = r1 < T ? lse[bh * T + r1] : 0.f;
    const uint64_t drow0 = bh * T + (uint64_t)min(r0, T - 1), drow1 = bh * T + (uint64_t)min(r1, T - 1);
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    float dq[DK / 8][4];
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int j0 = 0; j0 < T; j0 += TILE) {
      __syncthreads();
      stage<DK>(S0, qb + d, ld, j0, T);          // K tile
      stage<DK>(S1, qb + 2 * d, ld, j0, T);      // V tile
      __syncthreads();
      float s[TILE / 8][4], dp[TILE / 8][4];
      mma_a_bt<DK>(s, qa, S0, lane);
      mma_a_bt<DK>(dp, ga, S1, lane);
      uint32_t da[TILE / 16][4];
#pragma unroll
      for (int nt = 0; nt < TILE / 8; ++nt) {
        float ds[4], fa[2], fb[2];
        mt_attn_drop_pair(drop, drow0, P2, (uint32_t)(j0 + nt * 8 + c), fa[0], fa[1]);
        mt_attn_drop_pair(drop, drow1, P2, (uint32_t)(j0 + nt * 8 + c), fb[0], fb[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = j0 + nt * 8 + c + e;
          const bool in = j < T;
          const float p0 = __expf(s[nt][e] * scale - L0), p1 = __expf(s[nt][2 + e] * scale - L1);
          const float f0 = fa[e], f1 = fb[e];
          ds[e] = (v0 && in) ? p0 * (dp[nt][e] * f0 - D0) * scale : 0.f;
          ds[2 + e] = (v1 && in) ? p1 * (dp[nt][2 + e] * f1 - D1) * scale : 0.f;
        }
        da[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]);
        da[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_p_s<DK>(dq, da, S0, lane);             // dQ += dS . K
    }
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r0) * ld + hd * DK + i * 8 + c) = pack2(dq[i][0], dq[i][1]);
      if (r1 < T) *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r1) * ld + hd * DK + i * 8 + c) = pack2(dq[i][2], dq[i][3]);
    }
  }

  // ---------------- phase B: dK, dV (owned rows are KEYS; everything is the transposed problem) -------------
  {
    uint32_t ka[DK / 16][4], va[DK / 16][4];
    load_a_frags<DK>(ka, qb + d, ld, w0, T, lane);
    load_a_frags<DK>(va, qb + 2 * d, ld, w0, T, lane);
    float dk[DK / 8][4], dv[DK / 8][4];
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    for (int i0 = 0; i0 < T; i0 += TILE) {
      __syncthreads();
      stage<DK>(S0, qb, ld, i0, T);              // Q tile
      stage<DK>(S1, gob, d, i0, T);              // dO tile
      for (int e = threadIdx.x; e < TILE; e += THREADS) {
        const int i = i0 + e;
        float Dv = 0.f, Lv = 0.f, vv = 0.f;
        if (i < T) {
          const bf16* go = gob + (size_t)i * d;
          const bf16* oo = ob + (size_t)i * d;
#pragma unroll
          for (int q = 0; q < DK; q += 4) {
            float4 g4 = ld4(go + q), o4 = ld4(oo + q);
            Dv += g4.x * o4.x + g4.y * o4.y + g4.z * o4.z + g4.w * o4.w;
          }
          Lv = lse[bh * T + i];
          vv = (mask != nullptr && mask[(size_t)b * T + i] == 0.f) ? 0.f : 1.f;
        }
        Ds[e] = Dv; Ls[e] = Lv; Vd[e] = i < T ? (vv != 0.f ? 1.f : -1.f) : 0.f;     // 1 valid, -1 masked query, 0 out of range
      }
      __syncthreads();
      float s[TILE / 8][4], dp[TILE / 8][4];
      mma_a_bt<DK>(s, ka, S0, lane);             // S^T tile: rows = keys, cols = queries
      mma_a_bt<DK>(dp, va, S1, lane);            // dP^T
      uint32_t pa[TILE / 16][4], da[TILE / 16][4];
#pragma unroll
      for (int nt = 0; nt < TILE / 8; ++nt) {
        float pd[4], ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = nt * 8 + c + e;           // query column inside the tile
          const int i = i0 + ci;
          const float vd = Vd[ci], Lq = Ls[ci], Dq = Ds[ci];
          const uint64_t drow = bh * T + (uint64_t)min(i, T - 1);
          // masked query rows are constant rows: p = exp(0 - log T) = 1/T, and they carry no score gradient
          const float p0 = vd == 0.f ? 0.f : __expf((vd > 0.f ? s[nt][e] * scale : 0.f) - Lq);
          const float p1 = vd == 0.f ? 0.f : __expf((vd > 0.f ? s[nt][2 + e] * scale : 0.f) - Lq);
          const float f0 = mt_attn_drop_factor(drop, drow, (uint32_t)(T + 1) >> 1, (uint32_t)min(r0, T - 1));
          const float f1 = mt_attn_drop_factor(drop, drow, (uint32_t)(T + 1) >> 1, (uint32_t)min(r1, T - 1));
          pd[e] = p0 * f0; pd[2 + e] = p1 * f1;
          ds[e] = vd > 0.f ? p0 * (dp[nt][e] * f0 - Dq) * scale : 0.f;
          ds[2 + e] = vd > 0.f ? p1 * (dp[nt][2 + e] * f1 - Dq) * scale : 0.f;
        }
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack2(pd[0], pd[1]); pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(pd[2], pd[3]);
        da[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]); da[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_p_s<DK>(dv, pa, S1, lane);             // dV += (P^T . drop) dO
      mma_p_s<DK>(dk, da, S0, lane);             // dK += dS^T Q
    }
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) {
      if (r0 < T) {
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r0) * ld + d + hd * DK + i * 8 + c) = pack2(dk[i][0], dk[i][1]);
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r0) * ld + 2 * d + hd * DK + i * 8 + c) = pack2(dv[i][0], dv[i][1]);
      }
      if (r1 < T) {
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r1) * ld + d + hd * DK + i * 8 + c) = pack2(dk[i][2], dk[i][3]);
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r1) * ld + 2 * d + hd * DK + i * 8 + c) = pack2(dv[i][2], dv[i][3]);
      }
    }
  }
}

}  // namespace

bool mt_attn_mma_supported(int B, int T, int d, int h) {
  if (d % h != 0) return false;
  const int dk = d / h;
  return (dk == 16 || dk == 32 || dk == 64) && d % 8 == 0 && B <= 65535 && h <= 65535;
}

static int g_force_tiled = 0;
extern "C" int mt_attention_force_tiled(int on) { const int old = g_force_tiled; g_force_tiled = on; return old; }
bool mt_attn_force_tiled_on() { return g_force_tiled != 0; }

int mt_attn_mma_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                        const int* klen) {
  if (!g_force_tiled && mt_attn128_supported(B, T, d, h)) return mt_attn128_fwd_run(B, T, d, h, qkv, mask, out, lse, drop, st, klen);
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  dim3 grid((T + WARPS * 16 - 1) / (WARPS * 16), h, B);
  mt_prof_work(4.0 * B * (double)T * T * d, (double)B * T * d * 4.0 * 2.0);
  switch (dk) {
    case 16: attn_mma_fwd_kernel<16><<<grid, THREADS, 0, st>>>(B, T, d, h, (const bf16*)qkv, mask, (bf16*)out, lse, drop, scale, klen); break;
    case 32: attn_mma_fwd_kernel<32><<<grid, THREADS, 0, st>>>(B, T, d, h, (const bf16*)qkv, mask, (bf16*)out, lse, drop, scale, klen); break;
    case 64: attn_mma_fwd_kernel<64><<<grid, THREADS, 0, st>>>(B, T, d, h, (const bf16*)qkv, mask, (bf16*)out, lse, drop, scale, klen); break;
    default: return MT_ERR_UNSUPPORTED;
  }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_attn_mma_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                        void* dqkv, DropCfg drop, cudaStream_t st, float* dbias, bool* dbias_done) {
  if (dbias_done) *dbias_done = false;
  if (!g_force_tiled && mt_attn128_supported(B, T, d, h)) {
    if (dbias_done) *dbias_done = dbias != nullptr;
    return mt_attn128_bwd_run(B, T, d, h, qkv, mask, out, lse, dout, dqkv, drop, st, dbias);
  }
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  dim3 grid((T + WARPS * 16 - 1) / (WARPS * 16), h, B);
  mt_prof_work(14.0 * B * (double)T * T * d, (double)B * T * d * 9.0 * 2.0);
#define MT_GO(DK)                                                                                                                   \
  attn_mma_bwd_kernel<DK><<<grid, THREADS, 0, st>>>(B, T, d, h, (const bf16*)qkv, mask, (const bf16*)out, lse, (const bf16*)dout, \
                                                    (bf16*)dqkv, drop, scale)
  switch (dk) {
    case 16: MT_GO(16); break;
    case 32: MT_GO(32); break;
    case 64: MT_GO(64); break;
    default: return MT_ERR_UNSUPPORTED;
  }
#undef MT_GO
  MT_LAUNCH_CHECK();
  return MT_OK;
}
