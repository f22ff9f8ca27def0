// Attention core for sequences of up to 128 windows (the SEND narratives are 105-125 windows; BASELINE configs 1-3
// use T = 128): ONE CTA owns a whole (narrative, head) pair, so K / V / Q / dO are staged in shared memory exactly once
// and nothing is recomputed.  Same semantics as mt_attention.cu: query-ROW mask (the whole row becomes uniform 1/T),
// softmax, pair-hash dropout on the probabilities, P.V, head merge.
//
//   forward : warp w owns query rows 16w..16w+15: S = Q K^T for all keys stays in registers (mma.sync m16n8k16),
//             one-pass softmax in the exp2 domain, P repacked from accumulator to A-operand layout in registers.
//   backward: phase 1 (warp = 16 query rows) computes P and dS chunk by chunk (16 keys), accumulates dQ = dS K and
//             leaves P.drop and dS in shared memory as bf16 [query][key]; phase 2 (warp = 16 keys) contracts over the
//             queries: dV = (P.drop)^T dO and dK = dS^T Q with transposed ldmatrix loads.  5 GEMMs, 1 exp and half a
//             hash per score instead of the 7 / 2 / 2 of the tiled general kernel (mt_attention_mma.cu, any T).
#include "mt_mma.cuh"

namespace {

using namespace mtmma;

constexpr int TMAX = 128;
constexpr int NW = 8;                 // warps per CTA, 16 query (or key) rows each
constexpr int NT = NW * 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// stage rows [0, T) of a [T x DK] global operand into smem[TMAX][DK + 8]; rows >= T are zero
template <int DK>
__device__ __forceinline__ void stage_all(bf16* s, const bf16* base, int ld, int T) {
  constexpr int LD = DK + 8, V = DK / 8;
  for (int e = threadIdx.x; e < TMAX * V; e += NT) {
    const int r = e / V, c = (e % V) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < T) v = *reinterpret_cast<const uint4*>(base + (size_t)r * ld + c);
    *reinterpret_cast<uint4*>(s + r * LD + c) = v;
  }
}

// 2^x on the SFU (ex2.approx: 2 ulp, flushes denormals; -inf -> 0): this kernel is the bf16 path
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// dropout factors of keys j, j + 1 (j even) of the query row whose pair-index base is row * ceil(T / 2) (hoisted by the caller)
__device__ __forceinline__ void drop_pair_at(const DropCfg& d, uint64_t pair_base, uint32_t j, float& f0, float& f1) {
  if (d.thresh == 0u) { f0 = f1 = 1.0f; return; }
  const uint32_t bits = mt_draw32(d, pair_base + (uint64_t)(j >> 1));
  const uint32_t t16 = d.thresh >> 16;
  f0 = (bits & 0xFFFFu) >= t16 ? d.scale : 0.0f;
  f1 = (bits >> 16) >= t16 ? d.scale : 0.0f;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;                  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous version of stage_all: rows [0, T) of a [T x DK] global operand -> smem[TMAX][DK + 8], rows >= T zero-filled
template <int DK>
__device__ __forceinline__ void stage_all_async(bf16* s, const bf16* base, int ld, int T) {
  constexpr int LD = DK + 8, V = DK / 8;
  for (int e = threadIdx.x; e < TMAX * V; e += NT) {
    const int r = e / V, c = (e % V) * 8;
    cp_async16(s + r * LD + c, base + (size_t)min(r, T - 1) * ld + c, r < T);
  }
}

template <int DK>
struct FwdSmem {
  static constexpr int LD = DK + 8;
  static constexpr size_t BYTES = (size_t)(2 * 3 * TMAX * LD) * sizeof(bf16);      // double-buffered Q | K | V
};

// Persistent forward: a CTA walks (narrative, head) items; the next item's Q / K / V tiles are in flight (cp.async) while the
// current one is computed, so no global-load latency is exposed after the first item.
template <int DK, bool FULL>      // FULL: T == 128, no bound tests
__global__ void __launch_bounds__(NT, 2) attn128_fwd_kernel(int B, int T, int d, int h, const bf16* __restrict__ qkv,
                                                            const float* __restrict__ mask, bf16* __restrict__ out,
                                                            float* __restrict__ lse, DropCfg drop_in, float scale,
                                                            const int* __restrict__ klen) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int LD = DK + 8;
  extern __shared__ __align__(16) unsigned char fwd_smem[];
  bf16* bufs = reinterpret_cast<bf16*>(fwd_smem);                 // [2][Q | K | V][TMAX * LD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * d;
  const int n_items = B * h;
  const int i0 = warp * 16;
  {
    const bf16* qb = qkv + (size_t)(blockIdx.x / h) * T * ld + (blockIdx.x % h) * DK;
    stage_all_async<DK>(bufs, qb, ld, T);
    stage_all_async<DK>(bufs + TMAX * LD, qb + d, ld, T);
    stage_all_async<DK>(bufs + 2 * TMAX * LD, qb + 2 * d, ld, T);
    cp_async_commit();
  }
  int cur = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, cur ^= 1) {
  const int b = item / h, hd = item % h;
  {
    const int nxt = item + gridDim.x;
    if (nxt < n_items) {
      const bf16* qn = qkv + (size_t)(nxt / h) * T * ld + (nxt % h) * DK;
      bf16* nb = bufs + (cur ^ 1) * 3 * TMAX * LD;
      stage_all_async<DK>(nb, qn, ld, T);
      stage_all_async<DK>(nb + TMAX * LD, qn + d, ld, T);
      stage_all_async<DK>(nb + 2 * TMAX * LD, qn + 2 * d, ld, T);
    }
    cp_async_commit();
  }
  cp_async_wait<1>();
  __syncthreads();
  const bf16* Qs = bufs + cur * 3 * TMAX * LD;
  const bf16* Ks = Qs + TMAX * LD;
  const bf16* Vs = Ks + TMAX * LD;
  if (FULL || i0 < T) {
  uint32_t qa[DK / 16][4];
#pragma unroll
  for (int ks = 0; ks < DK / 16; ++ks) ldsm_a(qa[ks], Qs + i0 * LD + ks * 16, LD, lane);
  const int r0 = i0 + (lane >> 2), r1 = r0 + 8, c = 2 * (lane & 3);
  // masked query rows: every score becomes the same constant, i.e. scale 0 (reference: masked_fill(-1e9) over the row)
  const float rs0 = (mask != nullptr && r0 < T && mask[(size_t)b * T + r0] == 0.f) ? 0.f : scale * LOG2E;
  const float rs1 = (mask != nullptr && r1 < T && mask[(size_t)b * T + r1] == 0.f) ? 0.f : scale * LOG2E;
  // ragged inference (klen, !FULL only): keys beyond the narrative's own length do not exist
  const int Tk = (!FULL && klen != nullptr) ? max(1, min(klen[b], T)) : T;
  const int nkt = FULL ? TMAX / 8 : (Tk + 7) >> 3;    // 8-key tiles that hold at least one valid key
  float s[TMAX / 8][4];
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < TMAX / 8; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    if (nt < nkt) {
#pragma unroll
      for (int kp = 0; kp < DK / 32; ++kp) {
        uint32_t bb[4];
        ldsm4(bb, Ks + (nt * 8 + (lane & 7)) * LD + kp * 32 + (lane >> 3) * 8);
        mma16816(s[nt], qa[2 * kp], bb[0], bb[1]);
        mma16816(s[nt], qa[2 * kp + 1], bb[2], bb[3]);
      }
      if (DK % 32 != 0) {
        uint32_t bb[4];
        ldsm4(bb, Ks + (nt * 8 + (lane & 7)) * LD + (DK / 32) * 32 + ((lane >> 3) & 1) * 8);
        mma16816(s[nt], qa[DK / 32 * 2], bb[0], bb[1]);
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {          // keys beyond T never contribute (their probability is exactly 0)
      const bool in = FULL || nt * 8 + c + e < Tk;
      s[nt][e] = in ? s[nt][e] * rs0 : -INFINITY;
      s[nt][2 + e] = in ? s[nt][2 + e] * rs1 : -INFINITY;
      mx0 = fmaxf(mx0, s[nt][e]); mx1 = fmaxf(mx1, s[nt][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const uint64_t bh = (uint64_t)b * h + hd;
  const uint32_t P2 = (uint32_t)(T + 1) >> 1;
  const uint64_t dbase0 = (bh * T + (uint64_t)min(r0, T - 1)) * P2, dbase1 = (bh * T + (uint64_t)min(r1, T - 1)) * P2;
  float l0 = 0.f, l1 = 0.f;
  float o[DK / 8][4];
#pragma unroll
  for (int i = 0; i < DK / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < TMAX / 16; ++kk) {
    if (FULL || kk * 16 < Tk) {
      uint32_t pa[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = 2 * kk + hf;
        const float p0 = fast_exp2(s[nt][0] - mx0), p1 = fast_exp2(s[nt][1] - mx0), p2 = fast_exp2(s[nt][2] - mx1), p3 = fast_exp2(s[nt][3] - mx1);
        l0 += p0 + p1; l1 += p2 + p3;
        float f0, f1, f2, f3;
        drop_pair_at(drop, dbase0, (uint32_t)(nt * 8 + c), f0, f1);
        drop_pair_at(drop, dbase1, (uint32_t)(nt * 8 + c), f2, f3);
        pa[hf * 2 + 0] = pack2(p0 * f0, p1 * f1);
        pa[hf * 2 + 1] = pack2(p2 * f2, p3 * f3);
      }
#pragma unroll
      for (int nd = 0; nd < DK / 16; ++nd) {
        uint32_t bb[4];
        ldsm_b2(bb, Vs + kk * 16 * LD + nd * 16, LD, lane);
        mma16816(o[2 * nd], pa, bb[0], bb[1]);
        mma16816(o[2 * nd + 1], pa, bb[2], bb[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
  for (int i = 0; i < DK / 8; ++i) {
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r0) * d + hd * DK + i * 8 + c) = pack2(o[i][0] * inv0, o[i][1] * inv0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r1) * d + hd * DK + i * 8 + c) = pack2(o[i][2] * inv1, o[i][3] * inv1);
  }
  if (lse && (lane & 3) == 0) {      // natural-log LSE of the scaled scores (a masked row gives log T)
    if (r0 < T) lse[bh * T + r0] = (mx0 + log2f(l0)) * LN2;
    if (r1 < T) lse[bh * T + r1] = (mx1 + log2f(l1)) * LN2;
  }
  }                                  // i0 < T
  __syncthreads();                   // every warp is done with this buffer before the stage after next overwrites it
  }                                  // items
  cp_async_wait<0>();
}

// column sums over the 16 rows of an accumulator fragment set acc[DK/8][4] (rows lane/4 and lane/4 + 8, columns
// i*8 + 2*(lane%4) + {0,1}) added into s[DK]; rows beyond T hold exact zeros
template <int DK>
__device__ __forceinline__ void colsum_frag(float* s, const float (*acc)[4], int lane) {
#pragma unroll
  for (int i = 0; i < DK / 8; ++i) {
    float a = acc[i][0] + acc[i][2], b = acc[i][1] + acc[i][3];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane < 4) { atomicAdd(s + i * 8 + 2 * lane, a); atomicAdd(s + i * 8 + 2 * lane + 1, b); }
  }
}

template <int DK>
struct BwdSmem {
  static constexpr int LD = DK + 8, LP = TMAX + 8;
  static constexpr size_t BYTES = (size_t)(4 * TMAX * LD + 2 * TMAX * LP) * sizeof(bf16);
};

template <int DK, bool FULL>
__global__ void __launch_bounds__(NT, DK <= 32 ? 2 : 1) attn128_bwd_kernel(int B, int T, int d, int h, const bf16* __restrict__ qkv,
                                                                           const float* __restrict__ mask, const bf16* __restrict__ out,
                                                                           const float* __restrict__ lse, const bf16* __restrict__ dout,
                                                                           bf16* __restrict__ dqkv, DropCfg drop_in, float scale,
                                                                           float* __restrict__ dbias) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  constexpr int LD = BwdSmem<DK>::LD, LP = BwdSmem<DK>::LP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_cs[3 * DK];          // column sums of this head's dQ | dK | dV slabs (bias gradient of the QKV projection)
  if (dbias) for (int i = threadIdx.x; i < 3 * DK; i += NT) s_cs[i] = 0.f;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + TMAX * LD;
  bf16* Vs = Ks + TMAX * LD;
  bf16* Gs = Vs + TMAX * LD;            // dO
  bf16* Ps = Gs + TMAX * LD;            // P . dropout      [query][key]
  bf16* Ds = Ps + TMAX * LP;            // dS               [query][key]
  // Persistent: CTA c owns head c % h and walks the narratives c / h, c / h + gridDim.x / h, ... (gridDim.x % h == 0), so the
  // bias-gradient column sums accumulate in shared memory and leave with ONE atomic per column per CTA.  Loads are
  // staggered half a phase ahead, all on single buffers: K / V of the next item land during phase 2 (phase 1 is their only
  // reader), Q / dO of this item land during phase 1 (phase 2 is their only shared-memory reader; phase 1 takes its own 16
  // query rows of Q / dO / O straight from global memory as register fragments).
  const int hd = blockIdx.x % h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * d;
  const int bstep = gridDim.x / h;
  const int w0 = warp * 16;
  const int r0 = w0 + (lane >> 2), r1 = r0 + 8, c = 2 * (lane & 3);
  const int nchunk = FULL ? TMAX / 16 : (T + 15) >> 4;      // 16-row chunks holding at least one valid row
  {
    const bf16* qb = qkv + (size_t)(blockIdx.x / h) * T * ld + hd * DK;
    if ((int)blockIdx.x / h < B) { stage_all_async<DK>(Ks, qb + d, ld, T); stage_all_async<DK>(Vs, qb + 2 * d, ld, T); }
    cp_async_commit();
  }
  for (int b = blockIdx.x / h; b < B; b += bstep) {
  const bf16* qb = qkv + (size_t)b * T * ld + hd * DK;
  const bf16* gob = dout + (size_t)b * T * d + hd * DK;
  const bf16* ob = out + (size_t)b * T * d + hd * DK;
  const uint64_t bh = (uint64_t)b * h + hd;
  stage_all_async<DK>(Qs, qb, ld, T);
  stage_all_async<DK>(Gs, gob, d, T);
  cp_async_commit();
  // own rows of Q / dO as A fragments, and D = rowsum(dO * O) of the two owned query rows
  uint32_t qa[DK / 16][4], ga[DK / 16][4];
  float D0 = 0.f, D1 = 0.f;
  if (FULL || w0 < T) {
    uint32_t oa[DK / 16][4];
    load_a_frags<DK>(qa, qb, ld, w0, T, lane);
    load_a_frags<DK>(ga, gob, d, w0, T, lane);
    load_a_frags<DK>(oa, ob, d, w0, T, lane);
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 g2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&ga[ks][q]));
        const float2 o2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&oa[ks][q]));
        const float t = g2.x * o2.x + g2.y * o2.y;
        if (q & 1) D1 += t; else D0 += t;
      }
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  }
  cp_async_wait<1>();                          // K / V of this item have landed (Q / dO may still be in flight)
  __syncthreads();

  // ---------------- phase 1: this warp's 16 query rows -> P, dS (shared memory) and dQ --------------------
  if (FULL || w0 < T) {
    const bool in0 = FULL || r0 < T, in1 = FULL || r1 < T;
    const bool mk0 = mask != nullptr && in0 && mask[(size_t)b * T + r0] == 0.f;
    const bool mk1 = mask != nullptr && in1 && mask[(size_t)b * T + r1] == 0.f;
    const float rs0 = mk0 ? 0.f : scale * LOG2E, rs1 = mk1 ? 0.f : scale * LOG2E;
    // masked rows carry no score gradient (masked_fill blocks it) but still feed dV with P = 1/T
    const float gs0 = (in0 && !mk0) ? scale : 0.f, gs1 = (in1 && !mk1) ? scale : 0.f;
    const float L0 = in0 ? lse[bh * T + r0] * LOG2E : 0.f, L1 = in1 ? lse[bh * T + r1] * LOG2E : 0.f;
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    const uint64_t dbase0 = (bh * T + (uint64_t)min(r0, T - 1)) * P2, dbase1 = (bh * T + (uint64_t)min(r1, T - 1)) * P2;
    float dq[DK / 8][4];
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kc = 0; kc < nchunk; ++kc) {
      uint32_t da[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int j = kc * 16 + hf * 8;          // first key of this 8-key tile
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kp = 0; kp < DK / 32; ++kp) {
          uint32_t bk[4], bv[4];
          ldsm4(bk, Ks + (j + (lane & 7)) * LD + kp * 32 + (lane >> 3) * 8);
          ldsm4(bv, Vs + (j + (lane & 7)) * LD + kp * 32 + (lane >> 3) * 8);
          mma16816(s, qa[2 * kp], bk[0], bk[1]); mma16816(s, qa[2 * kp + 1], bk[2], bk[3]);
          mma16816(dp, ga[2 * kp], bv[0], bv[1]); mma16816(dp, ga[2 * kp + 1], bv[2], bv[3]);
        }
        if (DK % 32 != 0) {
          uint32_t bk[4], bv[4];
          ldsm4(bk, Ks + (j + (lane & 7)) * LD + (DK / 32) * 32 + ((lane >> 3) & 1) * 8);
          ldsm4(bv, Vs + (j + (lane & 7)) * LD + (DK / 32) * 32 + ((lane >> 3) & 1) * 8);
          mma16816(s, qa[DK / 32 * 2], bk[0], bk[1]);
          mma16816(dp, ga[DK / 32 * 2], bv[0], bv[1]);
        }
        float f0, f1, f2, f3;
        drop_pair_at(drop, dbase0, (uint32_t)(j + c), f0, f1);
        drop_pair_at(drop, dbase1, (uint32_t)(j + c), f2, f3);
        const bool k0 = FULL || j + c < T, k1 = FULL || j + c + 1 < T;
        const float p0 = (in0 && k0) ? fast_exp2(s[0] * rs0 - L0) : 0.f, p1 = (in0 && k1) ? fast_exp2(s[1] * rs0 - L0) : 0.f;
        const float p2 = (in1 && k0) ? fast_exp2(s[2] * rs1 - L1) : 0.f, p3 = (in1 && k1) ? fast_exp2(s[3] * rs1 - L1) : 0.f;
        const uint32_t pd01 = pack2(p0 * f0, p1 * f1), pd23 = pack2(p2 * f2, p3 * f3);
        const uint32_t ds01 = pack2(p0 * (dp[0] * f0 - D0) * gs0, p1 * (dp[1] * f1 - D0) * gs0);
        const uint32_t ds23 = pack2(p2 * (dp[2] * f2 - D1) * gs1, p3 * (dp[3] * f3 - D1) * gs1);
        *reinterpret_cast<uint32_t*>(Ps + r0 * LP + j + c) = pd01;
        *reinterpret_cast<uint32_t*>(Ps + r1 * LP + j + c) = pd23;
        *reinterpret_cast<uint32_t*>(Ds + r0 * LP + j + c) = ds01;
        *reinterpret_cast<uint32_t*>(Ds + r1 * LP + j + c) = ds23;
        da[hf * 2 + 0] = ds01; da[hf * 2 + 1] = ds23;
      }
#pragma unroll
      for (int nd = 0; nd < DK / 16; ++nd) {       // dQ += dS[:, chunk] K[chunk, :]
        uint32_t bb[4];
        ldsm_b2(bb, Ks + kc * 16 * LD + nd * 16, LD, lane);
        mma16816(dq[2 * nd], da, bb[0], bb[1]);
        mma16816(dq[2 * nd + 1], da, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) {
      if (in0) *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r0) * ld + hd * DK + i * 8 + c) = pack2(dq[i][0], dq[i][1]);
      if (in1) *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * T + r1) * ld + hd * DK + i * 8 + c) = pack2(dq[i][2], dq[i][3]);
    }
    if (dbias) colsum_frag<DK>(s_cs, dq, lane);
  }
  cp_async_wait<0>();                          // Q / dO tiles of this item
  __syncthreads();
  if (b + bstep < B) {                         // K / V are dead until the next item's phase 1: fetch them now
    const bf16* qn = qkv + (size_t)(b + bstep) * T * ld + hd * DK;
    stage_all_async<DK>(Ks, qn + d, ld, T);
    stage_all_async<DK>(Vs, qn + 2 * d, ld, T);
  }
  cp_async_commit();

  // ---------------- phase 2: this warp's 16 keys -> dV = (P.drop)^T dO, dK = dS^T Q -------------------------
  if (FULL || w0 < T) {
    float dk[DK / 8][4], dv[DK / 8][4];
#pragma unroll
    for (int i = 0; i < DK / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    for (int qc = 0; qc < nchunk; ++qc) {         // contraction over 16 queries at a time
      uint32_t pt[4], dt[4];
      ldsm_at(pt, Ps + qc * 16 * LP + w0, LP, lane);
      ldsm_at(dt, Ds + qc * 16 * LP + w0, LP, lane);
#pragma unroll
      for (int nd = 0; nd < DK / 16; ++nd) {
        uint32_t bg[4], bq[4];
        ldsm_b2(bg, Gs + qc * 16 * LD + nd * 16, LD, lane);
        ldsm_b2(bq, Qs + qc * 16 * LD + nd * 16, LD, lane);
        mma16816(dv[2 * nd], pt, bg[0], bg[1]); mma16816(dv[2 * nd + 1], pt, bg[2], bg[3]);
        mma16816(dk[2 * nd], dt, bq[0], bq[1]); mma16816(dk[2 * nd + 1], dt, bq[2], bq[3]);
      }
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
    if (dbias) { colsum_frag<DK>(s_cs + DK, dk, lane); colsum_frag<DK>(s_cs + 2 * DK, dv, lane); }
  }
  __syncthreads();                             // Q / dO / P / dS tiles are free for the next item
  }                                            // items
  cp_async_wait<0>();
  if (dbias) for (int i = threadIdx.x; i < 3 * DK; i += NT) atomicAdd(dbias + (i / DK) * d + hd * DK + (i % DK), s_cs[i]);
}

template <int DK>
int launch_bwd(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout, void* dqkv,
               DropCfg drop, float scale, float* dbias, cudaStream_t st) {
  static MtPerDeviceOnce once;
  if (once.first()) {
    MT_CUDA(cudaFuncSetAttribute(attn128_bwd_kernel<DK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdSmem<DK>::BYTES));
    MT_CUDA(cudaFuncSetAttribute(attn128_bwd_kernel<DK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdSmem<DK>::BYTES));
  }
  const int per_sm = DK <= 32 ? 2 : 1;
  int grid = (per_sm * 148 / (g_mt_tune[MT_TUNE_ATTN_SHARE] > 1 ? g_mt_tune[MT_TUNE_ATTN_SHARE] : 1) / h) * h;      // a multiple of h: every CTA keeps one head
  if (grid < h) grid = h;
  if (grid > B * h) grid = B * h;
  if (T == TMAX)
    attn128_bwd_kernel<DK, true><<<grid, NT, BwdSmem<DK>::BYTES, st>>>(B, T, d, h, (const bf16*)qkv, mask, (const bf16*)out, lse, (const bf16*)dout,
                                                                         (bf16*)dqkv, drop, scale, dbias);
  else
    attn128_bwd_kernel<DK, false><<<grid, NT, BwdSmem<DK>::BYTES, st>>>(B, T, d, h, (const bf16*)qkv, mask, (const bf16*)out, lse, (const bf16*)dout,
                                                                          (bf16*)dqkv, drop, scale, dbias);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

}  // namespace

bool mt_attn128_supported(int B, int T, int d, int h) {
  if (d % h != 0 || T < 1 || T > TMAX) return false;
  const int dk = d / h;
  return (dk == 16 || dk == 32 || dk == 64) && d % 8 == 0 && (long long)B * h <= 0x7fffffffLL;
}

int mt_attn128_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                       const int* klen) {
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  mt_prof_work(4.0 * B * (double)T * T * d, (double)B * T * d * 4.0 * 2.0);
  const int slots = 2 * 148 / (g_mt_tune[MT_TUNE_ATTN_SHARE] > 1 ? g_mt_tune[MT_TUNE_ATTN_SHARE] : 1);
  const int grid = B * h < slots ? B * h : slots;
#define MT_FWD(DK)                                                                                                                     \
  {                                                                                                                                    \
    static MtPerDeviceOnce once;                                                                                                       \
    if (once.first()) {                                                                                                                \
      MT_CUDA(cudaFuncSetAttribute(attn128_fwd_kernel<DK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdSmem<DK>::BYTES));  \
      MT_CUDA(cudaFuncSetAttribute(attn128_fwd_kernel<DK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdSmem<DK>::BYTES)); \
    }                                                                                                                                  \
    if (T == TMAX && !klen) attn128_fwd_kernel<DK, true><<<grid, NT, FwdSmem<DK>::BYTES, st>>>(B, T, d, h, (const bf16*)qkv, mask, (bf16*)out, lse, drop, scale, nullptr);  \
    else attn128_fwd_kernel<DK, false><<<grid, NT, FwdSmem<DK>::BYTES, st>>>(B, T, d, h, (const bf16*)qkv, mask, (bf16*)out, lse, drop, scale, klen);          \
  }
  switch (dk) {
    case 16: MT_FWD(16) break;
    case 32: MT_FWD(32) break;
    case 64: MT_FWD(64) break;
    default: return MT_ERR_UNSUPPORTED;
  }
#undef MT_FWD
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_attn128_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                       void* dqkv, DropCfg drop, cudaStream_t st, float* dbias) {
  const int dk = d / h;
  const float scale = 1.0f / sqrtf((float)dk);
  mt_prof_work(10.0 * B * (double)T * T * d, (double)B * T * d * 9.0 * 2.0);
  switch (dk) {
    case 16: return launch_bwd<16>(B, T, d, h, qkv, mask, out, lse, dout, dqkv, drop, scale, dbias, st);
    case 32: return launch_bwd<32>(B, T, d, h, qkv, mask, out, lse, dout, dqkv, drop, scale, dbias, st);
    case 64: return launch_bwd<64>(B, T, d, h, qkv, mask, out, lse, dout, dqkv, drop, scale, dbias, st);
    default: return MT_ERR_UNSUPPORTED;
  }
}
