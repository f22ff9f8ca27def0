// tcgen05 / TMEM flash attention for LONG sequences and 64-wide heads (BASELINE configuration 5: d = 512, h = 8, T = 4096), bf16.
// Semantics of attention() in MFT/multiTransformer.py:22-34 as called by MultiHeadedAttention.forward (:47-65): query-ROW mask (the whole
// row becomes uniform over all T keys), fill value -1e9, live padded keys, softmax, dropout on the probabilities, P.V, heads merged in
// place; nothing [T, T] ever leaves the SM.
//
// Forward.  A work item is (narrative, head, PAIR of 128-query tiles); the two warp groups of the CTA own one query tile each and share
// every K / V tile of the item (one TMA box [128 x 64] each per key tile, 3-stage ring).  Two passes over the keys instead of an online
// rescale of the TMEM accumulator:
//   pass 1: S = Q K^T (tcgen05.mma M128 N128 K64 into TMEM) -> tcgen05.ld -> running row maximum (no exponentials);
//   pass 2: S again -> p = exp2(s * scale - max) with the FINAL maximum, row sum, pair-hash dropout, P (bf16) written back over S as the
//           TMEM A operand -> O += P V (M128 N64 K128, accumulating across key tiles, never rescaled).
// The extra Q K^T pass costs a third more MMA work on a tensor pipe that idles anyway -- the per-probability instruction stream (exp2,
// dropout hash, conversions) is the bound -- and removes the correction warp / conditional-rescale machinery altogether.
// One thread issues every MMA in program order; the tensor pipe executes them in that order, so S of key tile kt + 1 (issued after the
// P.V product of kt) cannot overwrite P before it has been read.
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

constexpr int FF_NT = 320;                  // warps 0-3 / 4-7: query tile 0 / 1, warp 8: TMA, warp 9: MMA issue + TMEM
constexpr int FF_SMEM = 2 * TILE_B + NST * 2 * TILE_B + 512 + 1024;

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
  uint64_t* s_free = s_full + 2;            // [2]   pass 1: S has been read (128 arrivals)
  uint64_t* p_ready = s_free + 2;           // [2]   pass 2: P is in place (128 arrivals)
  uint64_t* o_full = p_ready + 2;           // [2]   all P.V products of the item are complete
  uint64_t* o_read = o_full + 2;            // [2]   O has been read out (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_read + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(q_full, 1); mbar_init(q_free, 1);
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1); mbar_init(&s_free[w], 128); mbar_init(&p_ready[w], 128); mbar_init(&o_full[w], 1); mbar_init(&o_read[w], 128);
    }
    mbar_init_fence();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_kt = a.n_kt;
  // TMEM: S of query tile w at w * 128 (P packed over its first 64 columns), O at 256 + w * 64

  if (warp == 8) {
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
  } else if (warp == 9) {
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
        // ---- pass 1: S only; S_w of the next key tile waits until the softmax warps have read the previous one ----
        for (int kt = 0; kt < n_kt; ++kt, ++use) {
          const int stage = (int)(use % NST);
          mbar_wait(&full[stage], (use / NST) & 1u);
          fence_after();
          const uint32_t sk = kv_u32 + (uint32_t)(stage * 2 * TILE_B);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (kt > 0 || it > 0) {
              // first S of an item: the previous item's last P.V read P from this region -- in program order before this MMA; its O read
              // is not needed here.  Within pass 1 wait for the read of the previous S.
              if (kt > 0) { mbar_wait(&s_free[w], n_sfree[w] & 1u); ++n_sfree[w]; fence_after(); }
            }
            issue_s(w, sk);
          }
          commit(&empty[stage]);
        }
        // ---- pass 2: S, then P.V once the softmax warps delivered P; S of kt + 1 is issued right behind P.V of kt ----
        // the last pass-1 S must have been read before it is overwritten
#pragma unroll
        for (int w = 0; w < 2; ++w) { mbar_wait(&s_free[w], n_sfree[w] & 1u); ++n_sfree[w]; }
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
              mma_ts(tmem_base + (uint32_t)(256 + w * HDF), tmem_base + (uint32_t)(w * 128 + 8 * ks), dv + (uint64_t)(128 * ks), idesc_o,
                     (kt > 0 || ks > 0) ? 1u : 0u);
            if (more) issue_s(w, sk_next);
            else commit(&o_full[w]);
          }
          commit(&empty[stage]);
          if (!more) commit(q_free);
        }
      }
    }
  } else {
    // ===== softmax / epilogue: warp group w owns query tile w of the pair, thread = query row =====
    const DropCfg drop = mt_drop_resolve(a.drop);
    const int w = warp >> 2, r = threadIdx.x & 127;
    const uint32_t t_s = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(w * 128);
    const uint32_t t_o = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 + w * HDF);
    const uint32_t P2 = (uint32_t)(a.T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    uint32_t n_sfull = 0;
    int it = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      const int qp = item % a.n_qp, bh = item / a.n_qp, hd = bh % a.h, b = bh / a.h;
      const int q = qp * 2 * TQ + w * TQ + r;           // query index inside the narrative
      const bool row_ok = q < a.T;
      const bool masked = a.mask != nullptr && row_ok && a.mask[(size_t)b * a.T + q] == 0.f;
      const float rs = masked ? 0.f : a.scale_log2;     // masked query rows: every score becomes the same constant
      // ---- pass 1: row maximum of the raw scores ----
      float mraw = -INFINITY;
      for (int kt = 0; kt < n_kt; ++kt) {
        mbar_wait(&s_full[w], n_sfull & 1u); ++n_sfull;
        fence_after();
        const int kleft = a.T - kt * TQ;                // keys of this tile that exist
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          ld32(t_s + (uint32_t)(c * 32), v);
          ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, (c * 32 + i < kleft) ? __uint_as_float(v[i]) : -INFINITY);
        }
        fence_before();
        mbar_arrive(&s_free[w]);
      }
      const float mx = masked ? 0.f : mraw * rs;         // rs > 0 for live rows: max(s * rs) = rs * max(s)
      const uint64_t rs2 = pk2(rs, rs), nmx2 = pk2(-mx, -mx);
      uint64_t l2 = pk2(0.f, 0.f);
      const uint64_t drow = ((uint64_t)bh * (uint64_t)a.T + (uint64_t)min(q, a.T - 1)) * (uint64_t)P2;      // pair-index base of this row
      // ---- pass 2: probabilities with the final maximum, P -> TMEM, O accumulates in TMEM ----
      for (int kt = 0; kt < n_kt; ++kt) {
        mbar_wait(&s_full[w], n_sfull & 1u); ++n_sfull;
        fence_after();
        const int kleft = a.T - kt * TQ;
        const uint64_t pb64 = drow + (uint64_t)(kt * (TQ / 2));
        const uint32_t plo = (uint32_t)pb64;
        const uint32_t phi = (uint32_t)(pb64 >> 32) * 0xC2B2AE35u;      // constant inside the tile unless the low word wraps (handled below)
        const bool wraps = plo > 0xFFFFFFFFu - 64u;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32], pk[16];
          ld32(t_s + (uint32_t)(c * 32), v);
          ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int j = c * 32 + i;
            float p0, p1;
            upk2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), rs2, nmx2), p0, p1);
            p0 = j < kleft ? ex2(p0) : 0.f;
            p1 = j + 1 < kleft ? ex2(p1) : 0.f;
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
          st16(t_s + (uint32_t)(c * 16), pk);
        }
        st_wait();
        fence_before();
        mbar_arrive(&p_ready[w]);
      }
      float l0, l1;
      upk2(l2, l0, l1);
      const float l = l0 + l1;
      const float inv = drop.scale / l;
      if (a.lse != nullptr && row_ok) a.lse[(size_t)bh * a.T + q] = (mx + log2f(l)) * LN2_F;      // natural-log LSE of the scaled scores
      mbar_wait(&o_full[w], (uint32_t)it & 1u);
      fence_after();
      uint32_t o[64];
      ld32(t_o, o);
      ld32(t_o + 32, o + 32);
      ld_wait();
      fence_before();
      mbar_arrive(&o_read[w]);
      if (row_ok) {
        bf16* op = a.out + ((size_t)b * a.T + q) * a.d + hd * HDF;
        const uint64_t inv2 = pk2(inv, inv);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
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
  if (warp == 9) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
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
