// tcgen05 / TMEM flash attention for LONG sequences and 64-wide heads (BASELINE configuration 5: d = 512, h = 8, T = 4096), bf16.
// Semantics of attention() in MFT/multiTransformer.py:22-34 as called by MultiHeadedAttention.forward (:47-65): query-ROW mask (the whole
// row becomes uniform over all T keys), fill value -1e9, live padded keys, softmax, dropout on the probabilities, P.V, heads merged in
// place; nothing [T, T] ever leaves the SM.
//
// Forward.  A work item is (narrative, head, PAIR of 128-query tiles); each query tile has two softmax warp groups (thread = query row,
// the groups split the key columns) and both tiles share every K / V tile of the item (one TMA box [128 x 64] each per key tile,
// 3-stage ring).  Two passes over the keys instead of an online
// rescale of the TMEM accumulator:
//   pass 1: S = Q K^T (tcgen05.mma M128 N128 K64 into TMEM) -> tcgen05.ld -> running row maximum (no exponentials);
//   pass 2: S again -> p = exp2(s * scale - max) with the FINAL maximum, row sum, pair-hash dropout, P (bf16) into its own TMEM columns as the
//           A operand -> O += P V (M128 N64 K128, accumulating across key tiles, never rescaled).
// The extra Q K^T pass costs a third more MMA work on a tensor pipe that idles anyway -- the per-probability instruction stream (exp2,
// dropout hash, conversions) is the bound -- and removes the correction warp / conditional-rescale machinery altogether.
// S, P and O live in disjoint TMEM columns (2 x 128 + 2 x 64 + 2 x 64 = 512), so the only hazards are explicit: S may be overwritten
// once both column halves have read it, P once the previous P.V product has completed (pv_done).
#include "mt_ops.cuh"
#include "mt_tcgen05.cuh"

namespace {

using namespace tc5;

constexpr int TQ = 128;                     // queries per warp group / keys per tile
constexpr int HDF = 64;                     // head width
constexpr int TILE_B = TQ * 128;            // [128 rows x 64 bf16], 128-byte swizzle
constexpr int NST = 3;                      // K / V ring stages
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

struct FlashFwdArgs {
  int B, T, d, h, n_qp, n_kt, n_items;      // n_qp = query-tile pairs per (b, h), n_kt = key tiles
  float scale_log2;
  const float* mask;
  bf16* out;
  float* lse;
  DropCfg drop;
};

constexpr int FF_NT = 576;                  // warps 0-15 softmax (query tile w = warp / 8, key-column half hf = (warp / 4) % 2), 16: TMA, 17: MMA + TMEM
constexpr int FF_SMEM = 2 * TILE_B + NST * 2 * TILE_B + 512 + 2 * 2 * 2 * TQ * 4 + 1024;

__global__ void __launch_bounds__(FF_NT, 1) attn_flash_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ FlashFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                  // two query tiles
  uint8_t* kv_s = smem + 2 * TILE_B;                    // ring: stage s = K at s * 2 * TILE_B, V behind it
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_s + NST * 2 * TILE_B);
  uint64_t* full = bars;                    // [NST] TMA -> MMA
  uint64_t* empty = bars + NST;             // [NST] MMA -> TMA
  uint64_t* q_full = empty + NST;           // [1]
  uint64_t* q_free = q_full + 1;            // [1]   all S MMAs of the item are complete
  uint64_t* s_full = q_free + 1;            // [2]   S of query tile w is in TMEM
  uint64_t* s_free = s_full + 2;            // [2]   pass 1: S has been read (256 arrivals)
  uint64_t* p_ready = s_free + 2;           // [2]   pass 2: S has been read and P is in place (256 arrivals)
  uint64_t* pv_done = p_ready + 2;          // [2]   the P.V product of the key tile is complete: P may be overwritten
  uint64_t* o_full = pv_done + 2;           // [2]   all P.V products of the item are complete
  uint64_t* o_read = o_full + 2;            // [2]   O has been read out (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_read + 2);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);      // [tile][half][row] partial row maxima
  float* xsum = xmax + 2 * 2 * TQ;                                                       // ... and partial row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(q_full, 1); mbar_init(q_free, 1);
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1); mbar_init(&s_free[w], 256); mbar_init(&p_ready[w], 256); mbar_init(&pv_done[w], 1); mbar_init(&o_full[w], 1);
      mbar_init(&o_read[w], 256);
    }
    mbar_init_fence();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_kt = a.n_kt;
  // TMEM columns: S of query tile w at w * 128 | O at 256 + w * 64 | P (bf16, packed two per column) at 384 + w * 64

  if (warp == 16) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t fill = 0;                     // ring uses so far
      int it = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        const int qp = item % a.n_qp, bh = item / a.n_qp, hd = bh % a.h, b = bh / a.h;
        const int row0 = b * a.T;
        if (it > 0) mbar_wait(q_free, (uint32_t)(it - 1) & 1u);
        mbar_expect_tx(q_full, 2 * TILE_B);
        tma_load_2d(q_s, &map_qkv, hd * HDF, row0 + qp * 2 * TQ, q_full);
        tma_load_2d(q_s + TILE_B, &map_qkv, hd * HDF, row0 + qp * 2 * TQ + TQ, q_full);
        for (int pass = 0; pass < 2; ++pass) {
          for (int kt = 0; kt < n_kt; ++kt, ++fill) {
            const int stage = (int)(fill % NST);
            mbar_wait(&empty[stage], ((fill / NST) & 1u) ^ 1u);
            uint8_t* sb = kv_s + stage * 2 * TILE_B;
            mbar_expect_tx(&full[stage], pass ? 2 * TILE_B : TILE_B);
            tma_load_2d(sb, &map_qkv, a.d + hd * HDF, row0 + kt * TQ, &full[stage]);
            if (pass) tma_load_2d(sb + TILE_B, &map_qkv, 2 * a.d + hd * HDF, row0 + kt * TQ, &full[stage]);
          }
        }
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(TQ, TQ, 0, 0);
      const uint32_t idesc_o = make_idesc(TQ, HDF, 0, 1);
      const uint32_t q_u32 = smem_u32(q_s), kv_u32 = smem_u32(kv_s);
      uint32_t use = 0;                      // ring uses consumed so far
      uint32_t n_sfree[2] = {0, 0}, n_pready[2] = {0, 0};      // completed waits on those barriers (phase counters)
      int it = 0;
      auto issue_s = [&](int w, uint32_t sk) {
        const uint64_t dq = make_desc(q_u32 + (uint32_t)(w * TILE_B), 16, 1024), dk = make_desc(sk, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < HDF / 16; ++ks) mma_ss(tmem_base + (uint32_t)(w * 128), dq + (uint64_t)(2 * ks), dk + (uint64_t)(2 * ks), idesc_s, ks > 0);
        commit(&s_full[w]);
      };
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        mbar_wait(q_full, (uint32_t)it & 1u);
        fence_after();
        // ---- pass 1: S only; S_w of the next key tile waits until the softmax warps have read the previous one.  (The first S of an
        //      item needs no wait: the softmax warps delivered the last P of the previous item only after reading its S.) ----
        for (int kt = 0; kt < n_kt; ++kt, ++use) {
          const int stage = (int)(use % NST);
          mbar_wait(&full[stage], (use / NST) & 1u);
          fence_after();
          const uint32_t sk = kv_u32 + (uint32_t)(stage * 2 * TILE_B);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (kt > 0) { mbar_wait(&s_free[w], n_sfree[w] & 1u); ++n_sfree[w]; fence_after(); }
            issue_s(w, sk);
          }
          commit(&empty[stage]);
        }
        // ---- pass 2: S, then P.V once the softmax warps delivered P (which also says that S has been read); S of kt + 1 is issued
        //      right behind P.V of kt ----
#pragma unroll
        for (int w = 0; w < 2; ++w) { mbar_wait(&s_free[w], n_sfree[w] & 1u); ++n_sfree[w]; }      // the last pass-1 S has been read
        fence_after();
        {
          const int stage = (int)(use % NST);
          mbar_wait(&full[stage], (use / NST) & 1u);
          fence_after();
          const uint32_t sk = kv_u32 + (uint32_t)(stage * 2 * TILE_B);
          issue_s(0, sk);
          issue_s(1, sk);
        }
        for (int kt = 0; kt < n_kt; ++kt, ++use) {
          const int stage = (int)(use % NST);
          const uint32_t sv = kv_u32 + (uint32_t)(stage * 2 * TILE_B + TILE_B);
          const bool more = kt + 1 < n_kt;
          uint32_t sk_next = 0;
          if (more) {
            const int st2 = (int)((use + 1) % NST);
            mbar_wait(&full[st2], ((use + 1) / NST) & 1u);
            fence_after();
            sk_next = kv_u32 + (uint32_t)(st2 * 2 * TILE_B);
          }
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (kt == 0 && it > 0) { mbar_wait(&o_read[w], (uint32_t)(it - 1) & 1u); fence_after(); }      // O of the previous item has been read out
            mbar_wait(&p_ready[w], n_pready[w] & 1u); ++n_pready[w];
            fence_after();
            const uint64_t dv = make_desc(sv, 8192, 1024);
#pragma unroll
            for (int ks = 0; ks < TQ / 16; ++ks)      // A = P in TMEM (8 columns per 16 keys), B = V as an MN-major operand (n = head column)
              mma_ts(tmem_base + (uint32_t)(256 + w * HDF), tmem_base + (uint32_t)(384 + w * 64 + 8 * ks), dv + (uint64_t)(128 * ks), idesc_o,
                     (kt > 0 || ks > 0) ? 1u : 0u);
            commit(&pv_done[w]);
            if (more) issue_s(w, sk_next);
            else commit(&o_full[w]);
          }
          commit(&empty[stage]);
          if (!more) commit(q_free);
        }
      }
    }
  } else {
    // ===== softmax / epilogue: 16 warps.  Query tile w = warp / 8; the two warp groups of a tile split its key COLUMNS (half hf takes
    //       keys 64 hf .. 64 hf + 63 of every key tile, and head columns 32 hf .. of O); thread = query row.  Row maxima (after pass
    //       1) and row sums (at the end) of the two halves meet in shared memory. =====
    const DropCfg drop = mt_drop_resolve(a.drop);
    const int w = warp >> 3, hf = (warp >> 2) & 1, r = threadIdx.x & 127;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + (uint32_t)(w * 128), t_o = t_lane + (uint32_t)(256 + w * HDF), t_p = t_lane + (uint32_t)(384 + w * 64);
    const uint32_t P2 = (uint32_t)(a.T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    float* my_max = xmax + (w * 2 + hf) * TQ + r;
    float* my_sum = xsum + (w * 2 + hf) * TQ + r;
    const float* ot_max = xmax + (w * 2 + (hf ^ 1)) * TQ + r;
    const float* ot_sum = xsum + (w * 2 + (hf ^ 1)) * TQ + r;
    uint32_t n_sfull = 0, n_pv = 0;
    int it = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      const int qp = item % a.n_qp, bh = item / a.n_qp, hd = bh % a.h, b = bh / a.h;
      const int q = qp * 2 * TQ + w * TQ + r;           // query index inside the narrative
      const bool row_ok = q < a.T;
      const bool masked = a.mask != nullptr && row_ok && a.mask[(size_t)b * a.T + q] == 0.f;
      const float rs = masked ? 0.f : a.scale_log2;     // masked query rows: every score becomes the same constant
      // ---- pass 1: row maximum of the raw scores over this half's key columns ----
      float mraw = -INFINITY;
      for (int kt = 0; kt < n_kt; ++kt) {
        mbar_wait(&s_full[w], n_sfull & 1u); ++n_sfull;
        fence_after();
        const int kleft = a.T - kt * TQ;                // keys of this tile that exist
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
          uint32_t v[32];
          ld32(t_s + (uint32_t)(c * 32), v);
          ld_wait();
          if (kleft >= TQ) {                            // interior key tile (all but possibly the last): no per-key test
#pragma unroll
            for (int i = 0; i < 32; i += 2) mraw = fmaxf(mraw, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, (c * 32 + i < kleft) ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
        fence_before();
        mbar_arrive(&s_free[w]);
      }
      *my_max = mraw;
      bar_sync(1 + w, 256);
      mraw = fmaxf(mraw, *ot_max);
      const float mx = masked ? 0.f : mraw * rs;         // rs > 0 for live rows: max(s * rs) = rs * max(s)
      const uint64_t rs2 = pk2(rs, rs), nmx2 = pk2(-mx, -mx);
      uint64_t l2 = pk2(0.f, 0.f);
      const uint64_t drow = ((uint64_t)bh * (uint64_t)a.T + (uint64_t)min(q, a.T - 1)) * (uint64_t)P2;      // pair-index base of this row
      // ---- pass 2: probabilities with the final maximum, P -> its own TMEM columns, O accumulates in TMEM ----
      for (int kt = 0; kt < n_kt; ++kt) {
        mbar_wait(&s_full[w], n_sfull & 1u); ++n_sfull;
        fence_after();
        const int kleft = a.T - kt * TQ;
        const uint64_t pb64 = drow + (uint64_t)(kt * (TQ / 2));
        const uint32_t plo = (uint32_t)pb64;
        const uint32_t phi = (uint32_t)(pb64 >> 32) * 0xC2B2AE35u;      // constant inside the tile unless the low word wraps (handled below)
        const bool wraps = plo > 0xFFFFFFFFu - 64u;
        const bool full_tile = kleft >= TQ;             // warp-uniform: the compiler hoists it out of the unrolled pair loop
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
          uint32_t v[32], pk[16];
          ld32(t_s + (uint32_t)(c * 32), v);
          ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int j = c * 32 + i;
            float p0, p1;
            upk2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), rs2, nmx2), p0, p1);
            if (full_tile) { p0 = ex2(p0); p1 = ex2(p1); }
            else { p0 = j < kleft ? ex2(p0) : 0.f; p1 = j + 1 < kleft ? ex2(p1) : 0.f; }
            l2 = add2(l2, pk2(p0, p1));
            if (dropping) {        // one draw per pair of keys: low half -> key j, high half -> key j + 1; the keep scale is applied with 1 / l
              const uint32_t off = (uint32_t)(c * 16 + (i >> 1));
              uint32_t bits;
              if (!wraps) bits = mt_mix32((plo + off) ^ drop.key ^ phi);
              else bits = mt_draw32(drop, pb64 + (uint64_t)off);
              p0 = (bits << 16) >= thr_hi ? p0 : 0.f;
              p1 = bits >= thr_hi ? p1 : 0.f;
            }
            pk[i >> 1] = pack_bf2(p0, p1);
          }
          if (c == 2 * hf && (kt > 0 || it > 0)) {      // the previous P.V product has read P (completion number n_pv - 1 of pv_done)
            mbar_wait(&pv_done[w], (n_pv - 1) & 1u);
            fence_after();
          }
          st16(t_p + (uint32_t)(c * 16), pk);
        }
        ++n_pv;
        st_wait();
        fence_before();
        mbar_arrive(&p_ready[w]);
      }
      float l0, l1;
      upk2(l2, l0, l1);
      float l = l0 + l1;
      *my_sum = l;
      bar_sync(1 + w, 256);
      l += *ot_sum;
      const float inv = drop.scale / l;
      if (hf == 0 && a.lse != nullptr && row_ok) a.lse[(size_t)bh * a.T + q] = (mx + log2f(l)) * LN2_F;      // natural-log LSE of the scaled scores
      mbar_wait(&o_full[w], (uint32_t)it & 1u);
      fence_after();
      uint32_t o[32];
      ld32(t_o + (uint32_t)(hf * 32), o);
      ld_wait();
      fence_before();
      mbar_arrive(&o_read[w]);
      if (row_ok) {
        bf16* op = a.out + ((size_t)b * a.T + q) * a.d + hd * HDF + hf * 32;
        const uint64_t inv2 = pk2(inv, inv);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            upk2(mul2(pk2(__uint_as_float(o[8 * g + 2 * k]), __uint_as_float(o[8 * g + 2 * k + 1])), inv2), f[2 * k], f[2 * k + 1]);
          uint4 u;
          u.x = pack_bf2(f[0], f[1]); u.y = pack_bf2(f[2], f[3]); u.z = pack_bf2(f[4], f[5]); u.w = pack_bf2(f[6], f[7]);
          *reinterpret_cast<uint4*>(op + 8 * g) = u;
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 17) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------------------------
// A work item is (narrative, head, key tile j): K_j / V_j stay in shared memory and dK_j / dV_j accumulate in TMEM while the query tiles
// i = 0 .. T/128 - 1 stream through a TMA ring (Q_i, dO_i and the four per-query vectors lse*log2e | D | score scale | gradient scale).
// Per block (i, j), in the transposed formulation of mt_attention_tc.cu (lane = key):
//   S^T = K_j Q_i^T, dP^T = V_j dO_i^T (tcgen05.mma into TMEM)  ->  16 compute warps (warp group g: queries 32 g .. 32 g + 31, thread =
//   key row) recompute P^T = exp2(S^T * scale - lse), apply the pair-hash dropout mask, form dS^T = P^T (dP^T * keep - D) * scale
//   ->  P^T (bf16) goes to its OWN TMEM columns, dS^T to shared memory (double-buffered), so S^T / dP^T of block i + 1 are issued as soon
//   as the compute warps have read block i, in front of the second round of block i:
//   dV_j += P^T dO_i (A from TMEM), dK_j += dS^T Q_i, dQ_i = dS K_j (A from shared memory, K-major / MN-major views of the same tile).
// dQ_i is a partial sum over the key tiles: it leaves with fp32 vector reductions into a [B*T, d] accumulation buffer (the classic
// flash-attention backward); a last pass rounds it into dqkv.  dK_j / dV_j are stored once per item.
constexpr int NSB = 3;                                   // Q / dO / aux ring stages
constexpr int FB_NT = 576;                               // warps 0-15 compute, 16 TMA, 17 MMA issue + TMEM
constexpr int FB_AUX = 4 * TQ * 4;                       // four per-query vectors
constexpr int FB_STAGE = 2 * TILE_B + FB_AUX;
constexpr int FB_DS = 2 * TILE_B;                        // dS^T of one block: [128 keys x 128 queries] bf16 as two 64-query tiles
constexpr int FB_SMEM = 2 * TILE_B + NSB * FB_STAGE + 2 * FB_DS + 512 + 1024;

struct FlashBwdArgs {
  int B, T, d, h, n_t, n_items, t_stride;  // n_t = tiles of 128 (queries and keys), t_stride = row stride of aux
  const float* aux;                        // [B][h][4][t_stride]
  float* dq_acc;                           // fp32 [B*T, d], zero-initialised
  bf16* dqkv;
  DropCfg drop;
};

__global__ void __launch_bounds__(FB_NT, 1)
attn_flash_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ FlashBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* kv_s = smem;                                  // K_j | V_j
  uint8_t* ring = smem + 2 * TILE_B;                     // stage: Q_i | dO_i | aux
  uint8_t* ds_s = ring + NSB * FB_STAGE;                 // two dS^T buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + 2 * FB_DS);
  uint64_t* full = bars;                    // [NSB]
  uint64_t* empty = bars + NSB;             // [NSB]
  uint64_t* kv_full = empty + NSB;          // [1]
  uint64_t* kv_free = kv_full + 1;          // [1]
  uint64_t* s_full = kv_free + 1;           // [1]  S^T, dP^T of the block are in TMEM
  uint64_t* s_free = s_full + 1;            // [1]  ... and have been read (512 arrivals)
  uint64_t* p_ready = s_free + 1;           // [1]  P^T (TMEM) and dS^T (shared) are in place (512 arrivals)
  uint64_t* pt_free = p_ready + 1;          // [1]  dV MMAs of the block are complete: P^T may be overwritten
  uint64_t* ds_free = pt_free + 1;          // [2]  dK / dQ MMAs that read this dS^T buffer are complete
  uint64_t* g_full = ds_free + 2;           // [1]  dQ of the block is in TMEM
  uint64_t* dq_free = g_full + 1;           // [1]  ... and has been drained (512 arrivals)
  uint64_t* acc_full = dq_free + 1;         // [1]  dK, dV of the item are final
  uint64_t* acc_read = acc_full + 1;        // [1]  ... and have been stored (512 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_read + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_t = a.n_t, T = a.T;
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    for (int s = 0; s < NSB; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(kv_full, 1); mbar_init(kv_free, 1); mbar_init(s_full, 1); mbar_init(s_free, 512); mbar_init(p_ready, 512);
    mbar_init(pt_free, 1); mbar_init(&ds_free[0], 1); mbar_init(&ds_free[1], 1); mbar_init(g_full, 1); mbar_init(dq_free, 512);
    mbar_init(acc_full, 1); mbar_init(acc_read, 512);
    mbar_init_fence();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S^T [0,128) | dP^T [128,256) | P^T packed bf16 [256,320) | dV [320,384) | dK [384,448) | dQ [448,512)
  const int n_my = ((int)blockIdx.x < a.n_items) ? (a.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 16) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t fill = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int jt = item % n_t, bh = item / n_t, hd = bh % a.h, b = bh / a.h;
        const int row0 = b * T;
        if (it > 0) mbar_wait(kv_free, (uint32_t)(it - 1) & 1u);
        mbar_expect_tx(kv_full, 2 * TILE_B);
        tma_load_2d(kv_s, &map_qkv, a.d + hd * HDF, row0 + jt * TQ, kv_full);
        tma_load_2d(kv_s + TILE_B, &map_qkv, 2 * a.d + hd * HDF, row0 + jt * TQ, kv_full);
        for (int i = 0; i < n_t; ++i, ++fill) {
          const int stage = (int)(fill % NSB);
          mbar_wait(&empty[stage], ((fill / NSB) & 1u) ^ 1u);
          uint8_t* sb = ring + stage * FB_STAGE;
          mbar_expect_tx(&full[stage], (uint32_t)FB_STAGE);
          tma_load_2d(sb, &map_qkv, hd * HDF, row0 + i * TQ, &full[stage]);
          tma_load_2d(sb + TILE_B, &map_do, hd * HDF, row0 + i * TQ, &full[stage]);
          const float* ax = a.aux + (size_t)bh * 4 * a.t_stride + (size_t)i * TQ;
#pragma unroll
          for (int k = 0; k < 4; ++k) bulk_load(sb + 2 * TILE_B + k * TQ * 4, ax + (size_t)k * a.t_stride, TQ * 4, &full[stage]);
        }
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(TQ, TQ, 0, 0);       // S^T / dP^T: both operands K-major
      const uint32_t idesc_ts = make_idesc(TQ, HDF, 0, 1);     // dV (A in TMEM) and dK (A K-major in shared memory), B MN-major
      const uint32_t idesc_dq = make_idesc(TQ, HDF, 1, 1);     // dQ: A = dS from shared memory, MN-major
      const uint32_t kv_u32 = smem_u32(kv_s), ring_u32 = smem_u32(ring), ds_u32 = smem_u32(ds_s);
      uint32_t use = 0, blk = 0;            // ring uses / blocks issued so far (S^T side)
      uint32_t blk2 = 0, use2 = 0;          // blocks whose second round has been issued
      auto round2 = [&](bool first_of_item) {
        const int stage = (int)(use2 % NSB);
        const uint32_t sq = ring_u32 + (uint32_t)(stage * FB_STAGE), sg = sq + TILE_B;
        const uint32_t sds = ds_u32 + (uint32_t)((blk2 & 1u) * FB_DS);
        mbar_wait(p_ready, blk2 & 1u);
        fence_after();
        const uint64_t dgm = make_desc(sg, 8192, 1024), dqm = make_desc(sq, 8192, 1024), dkm = make_desc(kv_u32, 8192, 1024);
#pragma unroll
        for (int ks = 0; ks < TQ / 16; ++ks)      // dV += P^T dO: A = P^T in TMEM (8 columns per 16 queries)
          mma_ts(tmem_base + 320, tmem_base + (uint32_t)(256 + 8 * ks), dgm + (uint64_t)(128 * ks), idesc_ts, (!first_of_item || ks > 0) ? 1u : 0u);
        commit(pt_free);
#pragma unroll
        for (int ks = 0; ks < TQ / 16; ++ks) {    // dK += dS^T Q: A K-major over the queries, two 64-query tiles
          const uint64_t da = make_desc(sds + (uint32_t)((ks >> 2) * TILE_B + (ks & 3) * 32), 16, 1024);
          mma_ss(tmem_base + 384, da, dqm + (uint64_t)(128 * ks), idesc_ts, (!first_of_item || ks > 0) ? 1u : 0u);
        }
        if (blk2 > 0) { mbar_wait(dq_free, (blk2 - 1) & 1u); fence_after(); }
        {
          const uint64_t dsm = make_desc(sds, TILE_B, 1024);
#pragma unroll
          for (int ks = 0; ks < TQ / 16; ++ks) mma_ss(tmem_base + 448, dsm + (uint64_t)(128 * ks), dkm + (uint64_t)(128 * ks), idesc_dq, ks > 0);
        }
        commit(g_full);
        commit(&ds_free[blk2 & 1u]);
        commit(&empty[stage]);
        ++blk2; ++use2;
      };
      for (int it = 0; it < n_my; ++it) {
        mbar_wait(kv_full, (uint32_t)it & 1u);
        fence_after();
        if (it > 0) { mbar_wait(acc_read, (uint32_t)(it - 1) & 1u); fence_after(); }
        const uint64_t dk0 = make_desc(kv_u32, 16, 1024), dv0 = make_desc(kv_u32 + TILE_B, 16, 1024);
        for (int i = 0; i < n_t; ++i, ++use, ++blk) {
          const int stage = (int)(use % NSB);
          mbar_wait(&full[stage], (use / NSB) & 1u);
          fence_after();
          if (blk > 0) { mbar_wait(s_free, (blk - 1) & 1u); fence_after(); }
          const uint32_t sq = ring_u32 + (uint32_t)(stage * FB_STAGE), sg = sq + TILE_B;
          const uint64_t dq0 = make_desc(sq, 16, 1024), dg0 = make_desc(sg, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < HDF / 16; ++ks) mma_ss(tmem_base, dk0 + (uint64_t)(2 * ks), dq0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
#pragma unroll
          for (int ks = 0; ks < HDF / 16; ++ks) mma_ss(tmem_base + 128, dv0 + (uint64_t)(2 * ks), dg0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
          commit(s_full);
          if (i > 0) round2(i == 1);
        }
        round2(n_t == 1);
        commit(kv_free);
        commit(acc_full);
      }
    }
  } else {
    // ===== compute: warp group g owns the queries [32 g, 32 g + 32) of every block, thread = key row j =====
    const DropCfg drop = mt_drop_resolve(a.drop);
    const int g = warp >> 2, j = threadIdx.x & 127;
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    const uint32_t odd = (uint32_t)j & 1u;
    const uint64_t ds2 = pk2(drop.scale, drop.scale);
    uint32_t fillc = 0, blk = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int jt = item % n_t, bh = item / n_t, hd = bh % a.h, b = bh / a.h;
      const int jg = jt * TQ + j;                      // key index inside the narrative
      const bool key_ok = jg < T;
      for (int i = 0; i < n_t; ++i, ++fillc, ++blk) {
        const int stage = (int)(fillc % NSB);
        const float* ax = reinterpret_cast<const float*>(ring + stage * FB_STAGE + 2 * TILE_B);
        uint8_t* dsb = ds_s + (blk & 1u) * FB_DS + (g >> 1) * TILE_B;      // this group's 64-query tile of the dS^T operand
        // pair index of (query q, key pair jg >> 1) = (bh T + q) P2 + (jg >> 1); this lane draws for the queries q + (j & 1)
        const uint64_t p64 = ((uint64_t)bh * (uint64_t)T + (uint64_t)(i * TQ + g * 32) + odd) * (uint64_t)P2 + (uint64_t)(jg >> 1);
        const uint32_t plo = (uint32_t)p64, phi = (uint32_t)(p64 >> 32) * 0xC2B2AE35u;
        const bool wraps = plo > 0xFFFFFFFFu - 32u * P2;
        const bool interior = (jt + 1) * TQ <= T && (i + 1) * TQ <= T;      // every key and query of the block exists (CTA-uniform)
        mbar_wait(&full[stage], (fillc / NSB) & 1u);         // per-query vectors (the MMA warp waits on the same phase for the tiles)
        mbar_wait(s_full, blk & 1u);
        fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int q0 = g * 32 + cc * 16;
          uint32_t s[16], dp[16], pkp[8], pks[8];
          ld16(tl + (uint32_t)q0, s);
          ld16(tl + (uint32_t)(128 + q0), dp);
          ld_wait();
          if (cc == 1) { fence_before(); mbar_arrive(s_free); }      // this thread's last read of S^T / dP^T
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const int q = q0 + e;
            const float4 L = *reinterpret_cast<const float4*>(ax + q), D = *reinterpret_cast<const float4*>(ax + TQ + q);
            const float4 rs = *reinterpret_cast<const float4*>(ax + 2 * TQ + q), gs = *reinterpret_cast<const float4*>(ax + 3 * TQ + q);
            const float Dk[4] = {D.x, D.y, D.z, D.w};
            float p[4], t[4];
            upk2(fma2(pk2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), pk2(rs.x, rs.y), pk2(-L.x, -L.y)), p[0], p[1]);
            upk2(fma2(pk2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), pk2(rs.z, rs.w), pk2(-L.z, -L.w)), p[2], p[3]);
            if (interior) {
#pragma unroll
              for (int k = 0; k < 4; ++k) p[k] = ex2(p[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) p[k] = (key_ok && i * TQ + q + k < T) ? ex2(p[k]) : 0.f;
            }
            // t = dP . keep-scale - D ; without a kept draw the probability's gradient is -D
            upk2(fma2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), ds2, pk2(-D.x, -D.y)), t[0], t[1]);
            upk2(fma2(pk2(__uint_as_float(dp[e + 2]), __uint_as_float(dp[e + 3])), ds2, pk2(-D.z, -D.w)), t[2], t[3]);
            float pd[4] = {p[0], p[1], p[2], p[3]};
            if (dropping) {
              // the 32-bit draw of (query, key pair jg >> 1) serves keys jg and jg ^ 1: this lane draws for query q + (j & 1), its
              // neighbour for the other query of the pair, and they swap
#pragma unroll
              for (int k = 0; k < 4; k += 2) {
                const uint32_t off = (uint32_t)(cc * 16 + e + k) * P2;
                const uint32_t mine = wraps ? mt_draw32(drop, p64 + (uint64_t)off) : mt_mix32((plo + off) ^ drop.key ^ phi);
                const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                const uint32_t b0 = odd ? other : mine, b1 = odd ? mine : other;
                const bool k0 = odd ? (b0 >= thr_hi) : ((b0 << 16) >= thr_hi), k1 = odd ? (b1 >= thr_hi) : ((b1 << 16) >= thr_hi);
                pd[k] = k0 ? p[k] : 0.f;
                t[k] = k0 ? t[k] : -Dk[k];
                pd[k + 1] = k1 ? p[k + 1] : 0.f;
                t[k + 1] = k1 ? t[k + 1] : -Dk[k + 1];
              }
            }
            float ev[4];
            upk2(mul2(mul2(pk2(p[0], p[1]), pk2(gs.x, gs.y)), pk2(t[0], t[1])), ev[0], ev[1]);
            upk2(mul2(mul2(pk2(p[2], p[3]), pk2(gs.z, gs.w)), pk2(t[2], t[3])), ev[2], ev[3]);
            pkp[e >> 1] = pack_bf2(pd[0], pd[1]);
            pkp[(e >> 1) + 1] = pack_bf2(pd[2], pd[3]);
            pks[e >> 1] = pack_bf2(ev[0], ev[1]);
            pks[(e >> 1) + 1] = pack_bf2(ev[2], ev[3]);
          }
          if (cc == 0) {      // the block's first writes: the previous block's dV MMAs have read P^T, and the dK / dQ MMAs of block
                              // blk - 2 have read this dS^T buffer
            if (blk > 0) mbar_wait(pt_free, (blk - 1) & 1u);
            if (blk > 1) mbar_wait(&ds_free[blk & 1u], ((blk >> 1) - 1) & 1u);
            fence_after();
          }
          st8(tl + (uint32_t)(256 + g * 16 + cc * 8), pkp);
          // dS^T row of this key: 16 queries = two 16-byte chunks of the 64-query tile (g >> 1), chunks (g & 1) * 4 + cc * 2 ..
          *reinterpret_cast<uint4*>(dsb + sw128_off(j, (g & 1) * 4 + cc * 2)) = make_uint4(pks[0], pks[1], pks[2], pks[3]);
          *reinterpret_cast<uint4*>(dsb + sw128_off(j, (g & 1) * 4 + cc * 2 + 1)) = make_uint4(pks[4], pks[5], pks[6], pks[7]);
        }
        st_wait();
        fence_proxy_async();
        fence_before();
        mbar_arrive(p_ready);
        // ---- drain dQ of the PREVIOUS block (its second round was issued behind this block's S^T): rows = queries, this group's
        //      16 columns, fp32 vector reductions into the accumulation buffer ----
        if (i > 0) {
          mbar_wait(g_full, (blk - 1) & 1u);
          fence_after();
          uint32_t dq[16];
          ld16(tl + (uint32_t)(448 + g * 16), dq);
          ld_wait();
          fence_before();
          mbar_arrive(dq_free);
          const int qg = (i - 1) * TQ + j;
          if (qg < T) {
            float* dst = a.dq_acc + ((size_t)b * T + qg) * a.d + hd * HDF + g * 16;
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * v4), "f"(__uint_as_float(dq[4 * v4])),
                           "f"(__uint_as_float(dq[4 * v4 + 1])), "f"(__uint_as_float(dq[4 * v4 + 2])), "f"(__uint_as_float(dq[4 * v4 + 3]))
                           : "memory");
          }
        }
      }
      // ---- item tail: dQ of the last block, then dK / dV of the key tile ----
      {
        mbar_wait(g_full, (blk - 1) & 1u);
        fence_after();
        uint32_t dq[16];
        ld16(tl + (uint32_t)(448 + g * 16), dq);
        ld_wait();
        fence_before();
        mbar_arrive(dq_free);
        const int qg = (n_t - 1) * TQ + j;
        if (qg < T) {
          float* dst = a.dq_acc + ((size_t)b * T + qg) * a.d + hd * HDF + g * 16;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * v4), "f"(__uint_as_float(dq[4 * v4])),
                         "f"(__uint_as_float(dq[4 * v4 + 1])), "f"(__uint_as_float(dq[4 * v4 + 2])), "f"(__uint_as_float(dq[4 * v4 + 3]))
                         : "memory");
        }
      }
      mbar_wait(acc_full, (uint32_t)it & 1u);
      fence_after();
      uint32_t gv[16], gk[16];
      ld16(tl + (uint32_t)(320 + g * 16), gv);
      ld16(tl + (uint32_t)(384 + g * 16), gk);
      ld_wait();
      fence_before();
      mbar_arrive(acc_read);
      if (key_ok) {
        bf16* gp = a.dqkv + ((size_t)b * T + jg) * (3 * a.d) + hd * HDF + g * 16;
        float fv[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) upk2(mul2(pk2(__uint_as_float(gv[e]), __uint_as_float(gv[e + 1])), ds2), fv[e], fv[e + 1]);      // dV still lacks the keep scale
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint4 u;
          u.x = pack_bf2(__uint_as_float(gk[8 * hh]), __uint_as_float(gk[8 * hh + 1])); u.y = pack_bf2(__uint_as_float(gk[8 * hh + 2]), __uint_as_float(gk[8 * hh + 3]));
          u.z = pack_bf2(__uint_as_float(gk[8 * hh + 4]), __uint_as_float(gk[8 * hh + 5])); u.w = pack_bf2(__uint_as_float(gk[8 * hh + 6]), __uint_as_float(gk[8 * hh + 7]));
          *reinterpret_cast<uint4*>(gp + a.d + 8 * hh) = u;
          uint4 w;
          w.x = pack_bf2(fv[8 * hh], fv[8 * hh + 1]); w.y = pack_bf2(fv[8 * hh + 2], fv[8 * hh + 3]);
          w.z = pack_bf2(fv[8 * hh + 4], fv[8 * hh + 5]); w.w = pack_bf2(fv[8 * hh + 6], fv[8 * hh + 7]);
          *reinterpret_cast<uint4*>(gp + 2 * a.d + 8 * hh) = w;
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 17) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// per-query scalars of the backward for any T: aux[b][hd][k][q] with row stride t_stride; one warp per token row (coalesced reads of
// out / dout), D closes inside the head's 8-lane group
__global__ void __launch_bounds__(256) flash_prep_kernel(int B, int T, int d, int h, int t_stride, const bf16* __restrict__ out,
                                                         const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                         const float* __restrict__ mask, float* __restrict__ aux, float scale) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * T;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const int b = (int)(row / T), q = (int)(row % T);
    const bool masked = mask != nullptr && mask[row] == 0.f;
    for (int c0 = 0; c0 < d; c0 += 256) {
      const int col = c0 + lane * 8;
      const bool valid = col < d;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      const uint4 o = valid ? *reinterpret_cast<const uint4*>(out + (size_t)row * d + col) : zero;
      const uint4 gq = valid ? *reinterpret_cast<const uint4*>(dout + (size_t)row * d + col) : zero;
      const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, gw[4] = {gq.x, gq.y, gq.z, gq.w};
      float D = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
        const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[k]));
        D += of.x * gf.x + of.y * gf.y;
      }
#pragma unroll
      for (int s = 1; s < HDF / 8; s <<= 1) D += __shfl_xor_sync(0xffffffffu, D, s);
      if (valid && (lane & (HDF / 8 - 1)) == 0) {
        const size_t bh = (size_t)b * h + col / HDF;
        float* ap = aux + bh * 4 * t_stride;
        ap[q] = lse[bh * T + q] * LOG2E_F;
        ap[t_stride + q] = D;
        ap[2 * t_stride + q] = masked ? 0.f : scale * LOG2E_F;
        ap[3 * t_stride + q] = masked ? 0.f : scale;
      }
    }
  }
}

// dqkv[:, 0:d] = bf16(dq_acc)
__global__ void flash_dq_finish_kernel(long long rows, int d, const float* __restrict__ acc, bf16* __restrict__ dqkv) {
  const long long n4 = rows * (d / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (d / 4);
    const int c = (int)(i % (d / 4)) * 4;
    st4(dqkv + r * 3 * d + c, ld4(acc + r * d + c));
  }
}

int num_sms_f() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace

bool mt_attn_flash_supported(int B, int T, int d, int h) {
  if (d % h != 0 || d / h != HDF || T < 1 || B < 1) return false;
  if ((long long)B * T > 0x7fffffffLL / (3LL * d)) return false;
  return d % 64 == 0;
}

int mt_attn_flash_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st) {
  if (!mt_attn_flash_supported(B, T, d, h)) return MT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return MT_ERR_ALIGN;
  CUtensorMap map;
  MT_TRY(make_map_2d(&map, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, TQ));
  FlashFwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h;
  a.n_qp = (T + 2 * TQ - 1) / (2 * TQ);
  a.n_kt = (T + TQ - 1) / TQ;
  a.n_items = B * h * a.n_qp;
  a.scale_log2 = LOG2E_F / sqrtf((float)HDF);
  a.mask = mask; a.out = (bf16*)out; a.lse = lse; a.drop = drop;
  static MtPerDeviceOnce once;
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(attn_flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM));
  const int sms = num_sms_f();
  const int grid = a.n_items < sms ? a.n_items : sms;
  mt_prof_work(4.0 * B * (double)T * T * d, (double)B * T * d * 4.0 * 2.0);
  attn_flash_fwd_kernel<<<grid, FF_NT, FF_SMEM, st>>>(map, a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

// Dws layout: per-query scalars [B][h][4][Tpad] then the fp32 dQ accumulation buffer [B*T, d]
int mt_attn_flash_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                          void* dqkv, DropCfg drop, float* ws, cudaStream_t st) {
  if (!mt_attn_flash_supported(B, T, d, h)) return MT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)dqkv & 15) || ((uintptr_t)ws & 15)) return MT_ERR_ALIGN;
  const int n_t = (T + TQ - 1) / TQ, tpad = n_t * TQ;
  float* aux = ws;
  float* dq_acc = ws + (size_t)B * h * 4 * tpad;
  MT_CUDA(cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)B * T * d, st));
  {
    const long long rows = (long long)B * T;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    flash_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, T, d, h, tpad, (const bf16*)out, (const bf16*)dout, lse, mask, aux, 1.0f / sqrtf((float)HDF));
    MT_LAUNCH_CHECK();
  }
  CUtensorMap map_qkv, map_do;
  MT_TRY(make_map_2d(&map_qkv, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, TQ));
  MT_TRY(make_map_2d(&map_do, dout, (uint64_t)d, (uint64_t)B * T, (uint64_t)d, 64, TQ));
  FlashBwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h; a.n_t = n_t; a.n_items = B * h * n_t; a.t_stride = tpad;
  a.aux = aux; a.dq_acc = dq_acc; a.dqkv = (bf16*)dqkv; a.drop = drop;
  static MtPerDeviceOnce once;
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(attn_flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
  const int sms = num_sms_f();
  const int grid = a.n_items < sms ? a.n_items : sms;
  mt_prof_work(10.0 * B * (double)T * T * d, (double)B * T * d * 8.0 * 2.0);
  attn_flash_bwd_kernel<<<grid, FB_NT, FB_SMEM, st>>>(map_qkv, map_do, a);
  MT_LAUNCH_CHECK();
  {
    const long long n4 = (long long)B * T * (d / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    flash_dq_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>((long long)B * T, d, dq_acc, (bf16*)dqkv);
    MT_LAUNCH_CHECK();
  }
  return MT_OK;
}
