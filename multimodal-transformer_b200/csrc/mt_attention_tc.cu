// tcgen05 / TMEM attention core for sequences of up to 128 windows and 32-wide heads (the MFT / SFT / B2 encoder stacks: d = 256,
// h = 8; SEND narratives are 105-125 windows, BASELINE configs 1-3 use T = 128).  Semantics of attention() in
// MFT/multiTransformer.py:22-34 as called by MultiHeadedAttention.forward (:47-65): query-ROW mask (the whole row becomes uniform),
// fill value -1e9, live padded keys, softmax, dropout on the probabilities, P.V, heads merged in place.
//
// One narrative is exactly one UMMA M tile (128 rows), so a work item is (narrative, PAIR of heads): the pair's Q / K / V slabs are
// three TMA boxes [128 rows x 64 columns] (128-byte swizzle) of the packed qkv activation, and one 4-warp group per head works
// with one query row (forward) or one key row (backward) per thread -- no cross-lane reductions anywhere.
//
//   forward : S = Q K^T (tcgen05.mma M128 N128 K32, accumulator in TMEM) -> tcgen05.ld, softmax in the exp2 domain, pair-hash dropout
//             -> P (bf16) written back over S in TMEM -> O = P V with P as the TMEM A operand (M128 N32 K128) -> tcgen05.ld, 1/l, store.
//   backward: transposed formulation so that both big contractions over the queries take their A operand from TMEM:
//             S^T = K Q^T and dP^T = V dO^T (lane = key) -> P^T, dS^T (bf16, TMEM) ; dV = P^T dO, dK = dS^T Q (TS form);
//             dS also goes to shared memory once (the thread's row is an MN-major A operand) for dQ = dS K.
//             The per-query scalars (log-sum-exp, D = rowsum(dO . O), mask) arrive as a small TMA bulk copy per item.
// Roles per CTA (one per SM, persistent over items): warps 0-3 / 4-7 = head 0 / head 1 of the pair, warp 8 = TMA producer,
// warp 9 = tcgen05.mma issuer (one elected lane) and TMEM owner.
#include "mt_ops.cuh"
#include "mt_tcgen05.cuh"

namespace {

using namespace tc5;

constexpr int TM = 128;                      // rows of a tile = max T
constexpr int HD = 32;                       // head width
constexpr int TILE_BYTES = TM * 128;         // [128 rows x 64 bf16], 128-byte swizzle
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int NT = 320;

int g_variant = 0;       // test hook (mt_attention_tc_variant): bit 0 = P through shared memory (SS form), bit 1 = 64-wide P.V tile

struct FwdArgs {
  int B, T, d, h, n_items, variant;
  float scale_log2;
  const float* mask;
  bf16* out;
  float* lse;
  const int* klen;
  DropCfg drop;
};

constexpr int FWD_STAGES = 3;
constexpr int FWD_STAGE_BYTES = 3 * TILE_BYTES;
constexpr int FWD_P_BYTES = 2 * TM * TM * 2;                                   // variant bit 0 only: P of both heads, K-major A operand
constexpr int FWD_SMEM = FWD_STAGES * FWD_STAGE_BYTES + FWD_P_BYTES + 256 + 1024;

// dropout factors of keys j, j + 1 (j even) of the query row whose pair-index base is row * ceil(T / 2)
__device__ __forceinline__ void drop_pair_at(const DropCfg& d, uint64_t pair_base, uint32_t j, float& f0, float& f1) {
  if (d.thresh == 0u) { f0 = f1 = 1.0f; return; }
  const uint32_t bits = mt_draw32(d, pair_base + (uint64_t)(j >> 1));
  const uint32_t t16 = d.thresh >> 16;
  f0 = (bits & 0xFFFFu) >= t16 ? d.scale : 0.0f;
  f1 = (bits >> 16) >= t16 ? d.scale : 0.0f;
}

template <bool FULL>
__global__ void __launch_bounds__(NT, 1) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* p_smem = smem + FWD_STAGES * FWD_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + FWD_P_BYTES);
  uint64_t* full = bars;                    // [FWD_STAGES] TMA -> MMA
  uint64_t* empty = bars + FWD_STAGES;      // [FWD_STAGES] MMA -> TMA
  uint64_t* s_full = empty + FWD_STAGES;    // [2] S of head w is in TMEM
  uint64_t* p_ready = s_full + 2;           // [2] P of head w is in place (128 arrivals)
  uint64_t* o_full = p_ready + 2;           // [2] O of head w is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp_count = a.h >> 1;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_ready[w], 128); mbar_init(&o_full[w], 1); }
    mbar_init_fence();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S of head w at w * 128 (P overlays its first 64 columns), O of head w at 256 + w * 64
  const bool p_via_smem = (a.variant & 1) != 0, wide_pv = (a.variant & 2) != 0;

  if (warp == 8) {
    // ===== TMA producer: Q | K | V boxes of the item's head pair =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int b = item / hp_count, hp = item % hp_count;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sb = stage_base + stage * FWD_STAGE_BYTES;
        mbar_expect_tx(&full[stage], FWD_STAGE_BYTES);
#pragma unroll
        for (int w = 0; w < 3; ++w) tma_load_2d(sb + w * TILE_BYTES, &map_qkv, w * a.d + hp * 64, b * a.T, &full[stage]);
        if (++stage == FWD_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(TM, TM, 0, 0);
      const uint32_t idesc_o = wide_pv ? make_idesc(TM, 64, 0, 1) : make_idesc(TM, HD, 0, 1);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)it & 1u;
        const uint32_t sq = smem_u32(stage_base + stage * FWD_STAGE_BYTES), sk = sq + TILE_BYTES, sv = sk + TILE_BYTES;
        mbar_wait(&full[stage], phase);
        fence_after();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          // S_w overwrites P_w of the previous item: wait until that item's P.V MMAs have completed
          if (it > 0) { mbar_wait(&o_full[w], par ^ 1); fence_after(); }
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            mma_ss(tmem_base + (uint32_t)(w * 128), make_desc(sq + 64 * w + 32 * ks, 16, 1024), make_desc(sk + 64 * w + 32 * ks, 16, 1024),
                   idesc_s, ks > 0);
          commit(&s_full[w]);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          mbar_wait(&p_ready[w], par);
          fence_after();
          const uint32_t d_o = tmem_base + (uint32_t)(256 + w * 64);
#pragma unroll
          for (int ks = 0; ks < TM / 16; ++ks) {
            const uint64_t bd = make_desc(sv + (wide_pv ? 0 : 64 * w) + 2048 * ks, 8192, 1024);
            if (p_via_smem)
              mma_ss(d_o, make_desc(smem_u32(p_smem) + w * (TM * TM * 2) + (ks >> 2) * TILE_BYTES + 32 * (ks & 3), 16, 1024), bd, idesc_o, ks > 0);
            else
              mma_ts(d_o, tmem_base + (uint32_t)(w * 128 + 8 * ks), bd, idesc_o, ks > 0);
          }
          commit(&o_full[w]);
        }
        commit(&empty[stage]);
        if (++stage == FWD_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== softmax / epilogue: warp group w owns head 2 hp + w, thread = query row =====
    const DropCfg drop = mt_drop_resolve(a.drop);
    const int w = warp >> 2, r = threadIdx.x & 127;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + (uint32_t)(w * 128), t_o = t_lane + (uint32_t)(256 + w * 64 + (wide_pv ? w * HD : 0));
    const uint32_t P2 = (uint32_t)(a.T + 1) >> 1;
    int it = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int b = item / hp_count, hd = 2 * (item % hp_count) + w;
      const bool row_ok = FULL || r < a.T;
      const bool masked = a.mask != nullptr && row_ok && a.mask[(size_t)b * a.T + r] == 0.f;
      const float rs = masked ? 0.f : a.scale_log2;        // masked query rows: every score becomes the same constant
      const int Tk = (!FULL && a.klen != nullptr) ? max(1, min(a.klen[b], a.T)) : a.T;
      const uint64_t bh = (uint64_t)b * a.h + hd;
      const uint64_t dbase = (bh * a.T + (uint64_t)min(r, a.T - 1)) * P2;
      mbar_wait(&s_full[w], par);
      fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ld32(t_s + (uint32_t)(c * 32), v);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (FULL || c * 32 + i < Tk) ? __uint_as_float(v[i]) * rs : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      float l = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], pk[16];
        ld32(t_s + (uint32_t)(c * 32), v);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int j = c * 32 + i;
          const float p0 = (FULL || j < Tk) ? ex2(__uint_as_float(v[i]) * rs - mx) : 0.f;
          const float p1 = (FULL || j + 1 < Tk) ? ex2(__uint_as_float(v[i + 1]) * rs - mx) : 0.f;
          l += p0 + p1;
          float f0, f1;
          drop_pair_at(drop, dbase, (uint32_t)j, f0, f1);
          pk[i >> 1] = pack_bf2(p0 * f0, p1 * f1);
        }
        if (p_via_smem) {         // K-major A operand: row r, keys c*32 .. c*32+31 = four 16-byte chunks of k-block c / 2
          uint8_t* pb = p_smem + w * (TM * TM * 2) + (c >> 1) * TILE_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(pb + sw128_off(r, (c & 1) * 4 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        } else {
          st16(t_s + (uint32_t)(c * 16), pk);
        }
      }
      if (p_via_smem) fence_proxy_async(); else st_wait();
      fence_before();
      mbar_arrive(&p_ready[w]);
      const float inv = 1.0f / l;
      if (a.lse != nullptr && row_ok) a.lse[bh * a.T + r] = (mx + log2f(l)) * LN2;      // natural-log LSE of the scaled scores
      mbar_wait(&o_full[w], par);
      fence_after();
      uint32_t o[32];
      ld32(t_o, o);
      ld_wait();
      if (row_ok) {
        bf16* op = a.out + ((size_t)b * a.T + r) * a.d + hd * HD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
          u.y = pack_bf2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
          u.z = pack_bf2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
          u.w = pack_bf2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(op + 8 * q) = u;
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------------------------
struct BwdArgs {
  int B, T, d, h, n_items, variant;
  const float* aux;       // [B][h][4][T]: lse * log2e | D = rowsum(dO . O) | score scale * log2e (0: masked row) | score-gradient scale (0: masked)
  bf16* dqkv;
  float* dbias;           // optional fp32 [3d], accumulated
  DropCfg drop;
};

constexpr int BWD_STAGES = 2;
constexpr int BWD_AUX_BYTES = 2 * 4 * TM * 4;                           // two heads x four per-query vectors
constexpr int BWD_STAGE_BYTES = 4 * TILE_BYTES + BWD_AUX_BYTES;         // Q | K | V | dO | aux
constexpr int BWD_DS_BYTES = TM * TM * 2;                               // dS of one head as an MN-major A operand
constexpr int BWD_CS_FLOATS = 4 * 2 * 96 * 2;                           // column sums: up to 8 head pairs are folded modulo 4 below -> sized for h <= 16
constexpr int BWD_SMEM = BWD_STAGES * BWD_STAGE_BYTES + 2 * BWD_DS_BYTES + 256 + 1024;

// aux[b][hd][k][q], see BwdArgs; one thread per (b, hd, q), q fastest
__global__ void attn_tc_prep_kernel(int B, int T, int d, int h, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                    const float* __restrict__ lse, const float* __restrict__ mask, float* __restrict__ aux, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * T) return;
  const int q = (int)(idx % T);
  const long long bh = idx / T;
  const int b = (int)(bh / h), hd = (int)(bh % h);
  const size_t row = (size_t)b * T + q;
  const uint4* o4 = reinterpret_cast<const uint4*>(out + row * d + hd * HD);
  const uint4* g4 = reinterpret_cast<const uint4*>(dout + row * d + hd * HD);
  float D = 0.f;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const uint4 o = o4[i], g = g4[i];
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
      const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[k]));
      D += of.x * gf.x + of.y * gf.y;
    }
  }
  const bool masked = mask != nullptr && mask[row] == 0.f;
  float* a = aux + (size_t)bh * 4 * T;
  a[q] = lse[bh * T + q] * LOG2E;
  a[T + q] = D;
  a[2 * T + q] = masked ? 0.f : scale * LOG2E;      // masked query rows: constant scores (uniform P) ...
  a[3 * T + q] = masked ? 0.f : scale;              // ... and no score gradient (masked_fill blocks it)
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

template <bool FULL>
__global__ void __launch_bounds__(NT, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_cs[4 * 2 * 96];                  // [head pair % 4][w][dQ | dK | dV][32]: bias-gradient column sums of this CTA
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* ds_smem = smem + BWD_STAGES * BWD_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + 2 * BWD_DS_BYTES);
  uint64_t* full = bars;                    // [2] TMA -> MMA / compute
  uint64_t* empty = bars + BWD_STAGES;      // [2] MMA -> TMA
  uint64_t* s_full = empty + BWD_STAGES;    // [2] S^T and dP^T of head w are in TMEM
  uint64_t* p_ready = s_full + 2;           // [2] P^T / dS^T (TMEM) and dS (smem) of head w are in place (128 arrivals)
  uint64_t* g_full = p_ready + 2;           // [2] dV, dK, dQ of head w are in TMEM
  uint64_t* g_read = g_full + 2;            // [2] ... and have been read out (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_read + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp_count = a.h >> 1;
  const int T = a.T;
  for (int i = threadIdx.x; i < 4 * 2 * 96; i += NT) s_cs[i] = 0.f;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_ready[w], 128); mbar_init(&g_full[w], 1); mbar_init(&g_read[w], 128); }
    mbar_init_fence();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = ((int)blockIdx.x < a.n_items) ? (a.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // TMEM columns of head w (base w * 256): [0,128) S^T -> P^T packed in [0,64), dV in [64,96), dK in [96,128);
  //                                        [128,256) dP^T -> dS^T packed in [128,192), dQ in [192,224)

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int b = item / hp_count, hp = item % hp_count;
        const int stage = it & 1;
        mbar_wait(&empty[stage], (((uint32_t)it >> 1) & 1u) ^ 1u);
        uint8_t* sb = stage_base + stage * BWD_STAGE_BYTES;
        mbar_expect_tx(&full[stage], (uint32_t)(4 * TILE_BYTES + 32 * T));
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_2d(sb + k * TILE_BYTES, &map_qkv, k * a.d + hp * 64, b * T, &full[stage]);
        tma_load_2d(sb + 3 * TILE_BYTES, &map_do, hp * 64, b * T, &full[stage]);
        bulk_load(sb + 4 * TILE_BYTES, a.aux + ((size_t)b * a.h + 2 * hp) * 4 * T, (uint32_t)(32 * T), &full[stage]);
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer: two independent per-head state machines, polled =====
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(TM, TM, 0, 0);
      const uint32_t idesc_ts = make_idesc(TM, HD, 0, 1);      // A in TMEM (K-major), B MN-major
      const uint32_t idesc_dq = make_idesc(TM, HD, 1, 1);      // A = dS from shared memory, MN-major
      int s_it[2] = {0, 0}, g_it[2] = {0, 0};
      uint32_t spins = 0;
      while (g_it[0] < n_my || g_it[1] < n_my) {
        bool progressed = false;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          // --- S^T_w = K Q^T and dP^T_w = V dO^T of item s_it[w] ---
          if (s_it[w] < n_my) {
            const int it = s_it[w];
            const int stage = it & 1;
            bool ok = mbar_test(&full[stage], ((uint32_t)it >> 1) & 1u);
            if (ok && it > 0) ok = mbar_test(&g_read[w], (uint32_t)(it - 1) & 1u);     // the accumulators of the previous item are drained
            if (ok && w == 1 && it == 0) ok = g_it[0] > 0;                            // start head 1 half a period behind head 0
            if (ok) {
              fence_after();
              const uint32_t sq = smem_u32(stage_base + stage * BWD_STAGE_BYTES), sk = sq + TILE_BYTES, sv = sk + TILE_BYTES, sg = sv + TILE_BYTES;
              const uint32_t tw = tmem_base + (uint32_t)(w * 256);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                mma_ss(tw, make_desc(sk + 64 * w + 32 * ks, 16, 1024), make_desc(sq + 64 * w + 32 * ks, 16, 1024), idesc_s, ks > 0);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                mma_ss(tw + 128, make_desc(sv + 64 * w + 32 * ks, 16, 1024), make_desc(sg + 64 * w + 32 * ks, 16, 1024), idesc_s, ks > 0);
              commit(&s_full[w]);
              ++s_it[w];
              progressed = true;
            }
          }
          // --- dV_w = P^T dO, dK_w = dS^T Q, dQ_w = dS K of item g_it[w] ---
          if (g_it[w] < s_it[w]) {
            const int it = g_it[w];
            if (mbar_test(&p_ready[w], (uint32_t)it & 1u)) {
              fence_after();
              const int stage = it & 1;
              const uint32_t sq = smem_u32(stage_base + stage * BWD_STAGE_BYTES), sk = sq + TILE_BYTES, sg = sk + 2 * TILE_BYTES;
              const uint32_t sds = smem_u32(ds_smem + w * BWD_DS_BYTES);
              const uint32_t tw = tmem_base + (uint32_t)(w * 256);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks)
                mma_ts(tw + 64, tw + 8 * ks, make_desc(sg + 64 * w + 2048 * ks, 8192, 1024), idesc_ts, ks > 0);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks)
                mma_ts(tw + 96, tw + 128 + 8 * ks, make_desc(sq + 64 * w + 2048 * ks, 8192, 1024), idesc_ts, ks > 0);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks)
                mma_ss(tw + 192, make_desc(sds + 2048 * ks, TILE_BYTES, 1024), make_desc(sk + 64 * w + 2048 * ks, 8192, 1024), idesc_dq, ks > 0);
              commit(&g_full[w]);
              ++g_it[w];
              if (g_it[w ^ 1] >= g_it[w]) commit(&empty[stage]);      // both heads of this item are done with the stage
              progressed = true;
            }
          }
        }
        if (progressed) spins = 0;
        else if (++spins > SPIN_LIMIT) __trap();
      }
    }
  } else {
    // ===== compute: warp group w owns head 2 hp + w, thread = key row j =====
    const DropCfg drop = mt_drop_resolve(a.drop);
    const int w = warp >> 2, j = threadIdx.x & 127;
    const uint32_t tw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(w * 256);
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    const uint32_t t16 = drop.thresh >> 16;
    const bool key_ok = FULL || j < T;
    const uint32_t odd = (uint32_t)j & 1u;
    uint8_t* dsb = ds_smem + w * BWD_DS_BYTES;
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = item / hp_count, hp = item % hp_count, hd = 2 * hp + w;
      const int stage = it & 1;
      const float* ax = reinterpret_cast<const float*>(stage_base + stage * BWD_STAGE_BYTES + 4 * TILE_BYTES) + w * 4 * T;
      const uint64_t bh = (uint64_t)b * a.h + hd;
      mbar_wait(&full[stage], ((uint32_t)it >> 1) & 1u);       // aux vectors (the MMA warp waits on the same phase for the tiles)
      mbar_wait(&s_full[w], par);
      fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32], dp[32], pkp[16], pks[16];
        ld32(tw + (uint32_t)(c * 32), s);
        ld32(tw + (uint32_t)(128 + c * 32), dp);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int q = c * 32 + i;
          const float2 L = *reinterpret_cast<const float2*>(ax + q), D = *reinterpret_cast<const float2*>(ax + T + q);
          const float2 rs = *reinterpret_cast<const float2*>(ax + 2 * T + q), gs = *reinterpret_cast<const float2*>(ax + 3 * T + q);
          const bool ok0 = key_ok && (FULL || q < T), ok1 = key_ok && (FULL || q + 1 < T);
          const float p0 = ok0 ? ex2(__uint_as_float(s[i]) * rs.x - L.x) : 0.f;
          const float p1 = ok1 ? ex2(__uint_as_float(s[i + 1]) * rs.y - L.y) : 0.f;
          float f0 = 1.f, f1 = 1.f;
          if (drop.thresh != 0u) {
            // the 32-bit draw of (query, key pair j >> 1) serves keys j and j ^ 1: this lane draws for query q + (j & 1), its
            // neighbour for the other query of the pair, and they swap
            const uint64_t qa = (uint64_t)min(q + (int)odd, T - 1);
            const uint32_t mine = mt_draw32(drop, (bh * T + qa) * P2 + (uint64_t)(j >> 1));
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const uint32_t b0 = odd ? other : mine, b1 = odd ? mine : other;
            f0 = ((odd ? (b0 >> 16) : (b0 & 0xFFFFu)) >= t16) ? drop.scale : 0.f;
            f1 = ((odd ? (b1 >> 16) : (b1 & 0xFFFFu)) >= t16) ? drop.scale : 0.f;
          }
          pkp[i >> 1] = pack_bf2(p0 * f0, p1 * f1);
          pks[i >> 1] = pack_bf2(p0 * (__uint_as_float(dp[i]) * f0 - D.x) * gs.x, p1 * (__uint_as_float(dp[i + 1]) * f1 - D.y) * gs.y);
        }
        st16(tw + (uint32_t)(c * 16), pkp);
        st16(tw + (uint32_t)(128 + c * 16), pks);
        // dS as the MN-major A operand of dQ = dS K: k row = key j, 64 queries per 128-byte row, two 64-query blocks
        uint8_t* db = dsb + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<uint4*>(db + sw128_off(j, (c & 1) * 4 + qd)) = make_uint4(pks[4 * qd], pks[4 * qd + 1], pks[4 * qd + 2], pks[4 * qd + 3]);
      }
      st_wait();
      fence_proxy_async();
      fence_before();
      mbar_arrive(&p_ready[w]);
      // ---- gradients: dV | dK rows = keys, dQ rows = queries; all rows beyond T are exact zeros ----
      mbar_wait(&g_full[w], par);
      fence_after();
      uint32_t gv[32], gk[32], gq[32];
      ld32(tw + 64, gv);
      ld32(tw + 96, gk);
      ld32(tw + 192, gq);
      ld_wait();
      fence_before();
      mbar_arrive(&g_read[w]);
      if (key_ok) {
        bf16* gp = a.dqkv + ((size_t)b * T + j) * (3 * a.d) + hd * HD;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint32_t* src = k == 0 ? gq : (k == 1 ? gk : gv);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            uint4 u;
            u.x = pack_bf2(__uint_as_float(src[8 * qd]), __uint_as_float(src[8 * qd + 1]));
            u.y = pack_bf2(__uint_as_float(src[8 * qd + 2]), __uint_as_float(src[8 * qd + 3]));
            u.z = pack_bf2(__uint_as_float(src[8 * qd + 4]), __uint_as_float(src[8 * qd + 5]));
            u.w = pack_bf2(__uint_as_float(src[8 * qd + 6]), __uint_as_float(src[8 * qd + 7]));
            *reinterpret_cast<uint4*>(gp + (size_t)k * a.d + 8 * qd) = u;
          }
        }
      }
      if (a.dbias != nullptr) {
        // column sums over the warp's 32 rows: butterfly that halves the number of live columns per lane at every step
        float v[96];
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(gq[i]); v[32 + i] = __uint_as_float(gk[i]); v[64 + i] = __uint_as_float(gv[i]); }
        int col = 0;
#pragma unroll
        for (int step = 0; step < 5; ++step) {
          const int n2 = 48 >> step;                // 48, 24, 12, 6, 3
          const int m = 16 >> step;
          const bool up = (lane & m) != 0;
#pragma unroll
          for (int i = 0; i < n2; ++i) {
            const float send = up ? v[i] : v[i + n2], keep = up ? v[i + n2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
          }
          col += up ? n2 : 0;
        }
        float* cs = s_cs + ((hp & 3) * 2 + w) * 96;
#pragma unroll
        for (int i = 0; i < 3; ++i) atomicAdd(cs + col + i, v[i]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (a.dbias != nullptr) {
    // s_cs[(hp % 4)][w][k][c] -> dbias[k * d + (2 hp + w) * 32 + c]; with more than 4 head pairs several pairs share a slot only if
    // h > 8, which mt_attn_tc_bwd_run excludes when dbias is requested
    for (int i = threadIdx.x; i < hp_count * 2 * 96; i += NT) {
      const int hp = i / 192, w = (i / 96) & 1, k = (i % 96) / 32, c = i % 32;
      const float val = s_cs[i];
      if (val != 0.f) atomicAdd(a.dbias + (size_t)k * a.d + (2 * hp + w) * HD + c, val);
    }
  }
}

// one-time (per device) opt-in to the large dynamic shared-memory carve-out
template <typename K>
int set_smem_attr(K kernel, int bytes, bool (&done)[16]) {
  int dev = 0;
  MT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || !done[dev]) {
    MT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 16) done[dev] = true;
  }
  return MT_OK;
}

int num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace

bool mt_attn_tc_supported(int B, int T, int d, int h) {
  if (d % h != 0 || d / h != HD || (h & 1) || T < 1 || T > TM) return false;
  if ((long long)B * T > 0x7fffffffLL / (3LL * d)) return false;
  return d % 64 == 0;
}

int mt_attn_tc_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                       const int* klen) {
  if (!mt_attn_tc_supported(B, T, d, h)) return MT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return MT_ERR_ALIGN;
  CUtensorMap map;
  MT_TRY(make_map_2d(&map, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, TM));
  FwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h; a.n_items = B * (h / 2); a.variant = g_variant;
  a.scale_log2 = LOG2E / sqrtf((float)HD);
  a.mask = mask; a.out = (bf16*)out; a.lse = lse; a.klen = klen; a.drop = drop;
  static bool attr_full[16] = {}, attr_part[16] = {};
  const int sms = num_sms();
  const int grid = a.n_items < sms ? a.n_items : sms;
  mt_prof_work(4.0 * B * (double)T * T * d, (double)B * T * d * 4.0 * 2.0);
  if (T == TM && !klen) {
    MT_TRY(set_smem_attr(attn_tc_fwd_kernel<true>, FWD_SMEM, attr_full));
    attn_tc_fwd_kernel<true><<<grid, NT, FWD_SMEM, st>>>(map, a);
  } else {
    MT_TRY(set_smem_attr(attn_tc_fwd_kernel<false>, FWD_SMEM, attr_part));
    attn_tc_fwd_kernel<false><<<grid, NT, FWD_SMEM, st>>>(map, a);
  }
  MT_LAUNCH_CHECK();
  return MT_OK;
}



int mt_attn_tc_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                       void* dqkv, DropCfg drop, float* dbias, float* aux, cudaStream_t st) {
  if (!mt_attn_tc_supported(B, T, d, h) || (T & 1)) return MT_ERR_UNSUPPORTED;      // T even: 16-byte granularity of the aux bulk copy
  if (dbias != nullptr && h > 8) return MT_ERR_UNSUPPORTED;
  if (!aux || ((uintptr_t)aux & 15) || ((uintptr_t)qkv & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)dqkv & 15) || ((uintptr_t)out & 15))
    return MT_ERR_ALIGN;
  const float scale = 1.0f / sqrtf((float)HD);
  {
    const long long n = (long long)B * h * T;
    attn_tc_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, T, d, h, (const bf16*)out, (const bf16*)dout, lse, mask, aux, scale);
    MT_LAUNCH_CHECK();
  }
  CUtensorMap map_qkv, map_do;
  MT_TRY(make_map_2d(&map_qkv, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, TM));
  MT_TRY(make_map_2d(&map_do, dout, (uint64_t)d, (uint64_t)B * T, (uint64_t)d, 64, TM));
  BwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h; a.n_items = B * (h / 2); a.variant = g_variant;
  a.aux = aux; a.dqkv = (bf16*)dqkv; a.dbias = dbias; a.drop = drop;
  static bool attr_full[16] = {}, attr_part[16] = {};
  const int sms = num_sms();
  const int grid = a.n_items < sms ? a.n_items : sms;
  mt_prof_work(10.0 * B * (double)T * T * d, (double)B * T * d * 8.0 * 2.0);
  if (T == TM) {
    MT_TRY(set_smem_attr(attn_tc_bwd_kernel<true>, BWD_SMEM, attr_full));
    attn_tc_bwd_kernel<true><<<grid, NT, BWD_SMEM, st>>>(map_qkv, map_do, a);
  } else {
    MT_TRY(set_smem_attr(attn_tc_bwd_kernel<false>, BWD_SMEM, attr_part));
    attn_tc_bwd_kernel<false><<<grid, NT, BWD_SMEM, st>>>(map_qkv, map_do, a);
  }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int g_mt_attn_no_tc = 0;

extern "C" {
int mt_attention_force_no_tc(int on) { int old = g_mt_attn_no_tc; g_mt_attn_no_tc = on; return old; }
size_t mt_attention_tc_bwd_ws_bytes(int B, int T, int h) { return mt_attn_bwd_ws_floats(B, T, h) * sizeof(float); }
int mt_attention_tc_bwd(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                        void* dqkv, float p_drop, uint64_t seed, uint32_t site, float* dbias, void* ws, size_t ws_bytes, void* stream) {
  if (!qkv || !out || !lse || !dout || !dqkv) return MT_ERR_ARG;
  if (!ws || ws_bytes < mt_attn_bwd_ws_floats(B, T, h) * sizeof(float)) return MT_ERR_WS;
  return mt_attn_tc_bwd_run(B, T, d, h, qkv, mask, out, lse, dout, dqkv, mt_make_drop(p_drop, seed, site), dbias, (float*)ws,
                            (cudaStream_t)stream);
}
/* test hook: variant bits of the tcgen05 attention kernels (see g_variant); returns the previous value */
int mt_attention_tc_variant(int v) { int old = g_variant; g_variant = v; return old; }
/* direct entry to the tcgen05 forward (tests / probes); MT_ERR_UNSUPPORTED outside its envelope */
int mt_attention_tc_fwd(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, float p_drop, uint64_t seed,
                        uint32_t site, const int* key_len, void* stream) {
  return mt_attn_tc_fwd_run(B, T, d, h, qkv, mask, out, lse, mt_make_drop(p_drop, seed, site), (cudaStream_t)stream, key_len);
}
}
