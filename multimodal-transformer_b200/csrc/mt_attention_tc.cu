// tcgen05 / TMEM attention core for sequences of up to 128 windows and 32-wide heads (the MFT / SFT / B2 encoder stacks: d = 256,
// h = 8; SEND narratives are 105-125 windows, BASELINE configs 1-3 use T = 128).  Semantics of attention() in
// MFT/multiTransformer.py:22-34 as called by MultiHeadedAttention.forward (:47-65): query-ROW mask (the whole row becomes uniform),
// fill value -1e9, live padded keys, softmax, dropout on the probabilities, P.V, heads merged in place.
//
// One narrative is exactly one UMMA M tile (128 rows), so a work item is (narrative, PAIR of heads): the pair's Q / K / V slabs are
// three TMA boxes [128 rows x 64 columns] (128-byte swizzle) of the packed qkv activation, and the compute warps work with one
// query row (forward) or one key row (backward) per thread -- no cross-lane reductions in the main loops.
//
//   forward : S = Q K^T (tcgen05.mma M128 N128 K32, accumulator in TMEM) -> tcgen05.ld, softmax in the exp2 domain, pair-hash dropout
//             -> P (bf16) written back over S in TMEM -> O = P V with P as the TMEM A operand (M128 N32 K128) -> tcgen05.ld, scale / l,
//             store.  Two CTAs per SM (one warp group per head each) so that one CTA's softmax covers the other's MMA / TMA latency.
//   backward: transposed formulation so that both big contractions over the queries take their A operand from TMEM:
//             S^T = K Q^T and dP^T = V dO^T (lane = key) -> P^T, dS^T (bf16, TMEM) ; dV = P^T dO, dK = dS^T Q (TS form);
//             dS also goes to shared memory once (the thread's row is an MN-major A operand) for dQ = dS K.
//             The per-query scalars (log-sum-exp, D = rowsum(dO . O), mask) arrive as a TMA bulk copy per item.  Two warp groups per
//             head split the query columns (16 compute warps per SM); the QKV bias gradient (column sums of dQ | dK | dV) is reduced
//             with a halving butterfly and leaves with one atomic per column per CTA.
// The instruction stream of both kernels is dominated by the per-probability work (exp2, dropout hash, conversions), not by the MMAs:
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2), the dropout scale folded into the row normaliser, 32-bit pair indices.
#include "mt_dropbits.cuh"
#include "mt_ops.cuh"
#include "mt_tcgen05.cuh"

namespace {

using namespace tc5;

constexpr int TM = 128;                      // rows of a tile = max T
constexpr int HD = 32;                       // head width
constexpr int TILE_BYTES = TM * 128;         // [128 rows x 64 bf16], 128-byte swizzle
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------------------
// Grouped launches (G modality stacks back to back, B narratives each): a CTA is pinned to ONE group -- blockIdx % G -- and walks that
// group's (narrative, head pair) items, so the dropout key, the mask / key-length rows (shared by the groups) and the bias-gradient block
// are fixed per CTA; global narrative index g * B + b addresses qkv / out / lse / aux, the LOCAL b the mask and the dropout stream.
constexpr int MAXG = MT_BITS_MAXG;
struct FwdArgs {
  int B, T, d, h, G;      // B = narratives per group
  float scale_log2;
  const float* mask;
  bf16* out;
  float* lse;
  const int* klen;
  const uint32_t* dbits;   // optional precomputed keep bits (attn_tc_dropbits_kernel): [G*B][h][4][128] words, bit i of word (c, q) = key 32 c + i of query q
  DropCfg drop[MAXG];
  int early_trigger;       // programmatic launch: let the next kernel be scheduled as soon as this grid is resident (mt_tune 15 bit 0 clears it)
};
struct ItemWalk {          // items of this CTA: global item = base + it * step, it < n
  int grp, base, step, n;
};
__device__ __forceinline__ ItemWalk item_walk(int G, int B, int hp_count) {
  ItemWalk w;
  const int items_g = B * hp_count, cpg = (int)gridDim.x / G, rank = (int)blockIdx.x / G;
  w.grp = (int)blockIdx.x % G;
  w.base = w.grp * items_g + rank;
  w.step = cpg;
  w.n = rank < items_g ? (items_g - 1 - rank) / cpg + 1 : 0;
  return w;
}

constexpr int FWD_NT = 320;                  // warps 0-3 / 4-7: head 0 / 1 of the pair, warp 8: TMA, warp 9: MMA issue + TMEM
constexpr int FWD_STAGES = 2;
constexpr int FWD_STAGE_BYTES = 3 * TILE_BYTES;
constexpr int FWD_SMEM = FWD_STAGES * FWD_STAGE_BYTES + 256 + 1024;     // 99.6 KB: two CTAs per SM
constexpr int FWD_TMEM = 256;                // head w: S at w * 128 (P packed over its first 64 columns), O at w * 128 + 64

template <bool FULL, bool BITS>
__global__ void __launch_bounds__(FWD_NT, 2) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_STAGES * FWD_STAGE_BYTES);
  uint64_t* full = bars;                    // [FWD_STAGES] TMA -> MMA
  uint64_t* empty = bars + FWD_STAGES;      // [FWD_STAGES] MMA -> TMA
  uint64_t* s_full = empty + FWD_STAGES;    // [2] S of head w is in TMEM
  uint64_t* p_ready = s_full + 2;           // [2] P of head w is in place (128 arrivals)
  uint64_t* o_full = p_ready + 2;           // [2] O of head w is in TMEM
  uint64_t* o_read = o_full + 2;            // [2] ... and has been read out (128 arrivals): S of the next item may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_read + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp_count = a.h >> 1;
  const ItemWalk iw = item_walk(a.G, a.B, hp_count);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_ready[w], 128); mbar_init(&o_full[w], 1); mbar_init(&o_read[w], 128); }
    mbar_init_fence();
  }
  if (warp == 9) tmem_alloc<FWD_TMEM>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  mt_pdl_wait();      // everything above touches no global memory
  if (a.early_trigger) mt_pdl_trigger();

  if (warp == 8) {
    // ===== TMA producer: Q | K | V boxes of the item's head pair =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iw.n; ++it) {
        const int item = iw.base + it * iw.step;
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
      const uint32_t idesc_o = make_idesc(TM, HD, 0, 1);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iw.n; ++it) {
        const uint32_t par = (uint32_t)it & 1u;
        const uint32_t sq = smem_u32(stage_base + stage * FWD_STAGE_BYTES), sk = sq + TILE_BYTES, sv = sk + TILE_BYTES;
        mbar_wait(&full[stage], phase);
        fence_after();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          // S_w overwrites P_w and O_w of the previous item: its P.V MMAs are complete and O has been read once o_read flips
          if (it > 0) { mbar_wait(&o_read[w], par ^ 1); fence_after(); }
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
#pragma unroll
          for (int ks = 0; ks < TM / 16; ++ks)      // A = P in TMEM (8 columns per 16 keys), B = V as an MN-major operand (n = head column)
            mma_ts(tmem_base + (uint32_t)(w * 128 + 64), tmem_base + (uint32_t)(w * 128 + 8 * ks), make_desc(sv + 64 * w + 2048 * ks, 8192, 1024),
                   idesc_o, ks > 0);
          commit(&o_full[w]);
        }
        commit(&empty[stage]);
        if (++stage == FWD_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== softmax / epilogue: warp group w owns head 2 hp + w, thread = query row =====
    const DropCfg drop = mt_drop_resolve(a.drop[iw.grp]);
    const int w = warp >> 2, r = threadIdx.x & 127;
    const uint32_t t_s = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(w * 128);
    const uint32_t P2 = (uint32_t)(a.T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    for (int it = 0; it < iw.n; ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int item = iw.base + it * iw.step;
      const int b = item / hp_count, hd = 2 * (item % hp_count) + w;      // b: global narrative (group * B + local)
      const int bl = b - iw.grp * a.B;
      const bool row_ok = FULL || r < a.T;
      const bool masked = a.mask != nullptr && row_ok && a.mask[(size_t)bl * a.T + r] == 0.f;
      const float rs = masked ? 0.f : a.scale_log2;        // masked query rows: every score becomes the same constant
      const int Tk = (!FULL && a.klen != nullptr) ? max(1, min(a.klen[bl], a.T)) : a.T;
      const uint32_t bh = (uint32_t)(b * a.h + hd);
      const uint32_t dbase = ((uint32_t)(bl * a.h + hd) * (uint32_t)a.T + (uint32_t)min(r, a.T - 1)) * P2;      // pair-index base of this row, group-local (fits: see host check)
      uint32_t kw0 = 0u, kw1 = 0u, kw2 = 0u, kw3 = 0u;      // BITS: this row's 128 keep bits, drawn once per step by attn_tc_dropbits_kernel
      if (BITS && dropping) {
        const uint32_t* kp = a.dbits + (size_t)bh * 512 + r;
        kw0 = __ldg(kp); kw1 = __ldg(kp + 128); kw2 = __ldg(kp + 256); kw3 = __ldg(kp + 384);
      }
      mbar_wait(&s_full[w], par);
      fence_after();
      float mraw = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ld32(t_s + (uint32_t)(c * 32), v);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float s0 = (FULL || c * 32 + i < Tk) ? __uint_as_float(v[i]) : -INFINITY;
          const float s1 = (FULL || c * 32 + i + 1 < Tk) ? __uint_as_float(v[i + 1]) : -INFINITY;
          mraw = fmaxf(mraw, fmaxf(s0, s1));
        }
      }
      const float mx = masked ? 0.f : mraw * rs;           // rs > 0 for live rows: max(s * rs) = rs * max(s)
      const uint64_t rs2 = pk2(rs, rs), nmx2 = pk2(-mx, -mx);
      uint64_t l2 = pk2(0.f, 0.f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], pk[16];
        ld32(t_s + (uint32_t)(c * 32), v);
        ld_wait();
        const uint32_t pbase = dbase + (uint32_t)(c * 16);
        const uint32_t kw = c == 0 ? kw0 : (c == 1 ? kw1 : (c == 2 ? kw2 : kw3));
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int j = c * 32 + i;
          float p0, p1;
          upk2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), rs2, nmx2), p0, p1);
          p0 = (FULL || j < Tk) ? ex2(p0) : 0.f;
          p1 = (FULL || j + 1 < Tk) ? ex2(p1) : 0.f;
          l2 = add2(l2, pk2(p0, p1));
          if (dropping) {        // one draw per pair of keys: low half -> key j, high half -> key j + 1; the keep scale is applied with 1 / l
            if (BITS) {
              p0 = (kw & (1u << i)) ? p0 : 0.f;
              p1 = (kw & (2u << i)) ? p1 : 0.f;
            } else {
              const uint32_t bits = mt_mix32((pbase + (uint32_t)(i >> 1)) ^ drop.key);
              p0 = (bits << 16) >= thr_hi ? p0 : 0.f;
              p1 = bits >= thr_hi ? p1 : 0.f;
            }
          }
          pk[i >> 1] = pack_bf2(p0, p1);
        }
        st16(t_s + (uint32_t)(c * 16), pk);
      }
      st_wait();
      fence_before();
      mbar_arrive(&p_ready[w]);
      float l0, l1;
      upk2(l2, l0, l1);
      const float l = l0 + l1;
      const float inv = drop.scale / l;
      if (a.lse != nullptr && row_ok) a.lse[(size_t)bh * a.T + r] = (mx + log2f(l)) * LN2;      // natural-log LSE of the scaled scores
      mbar_wait(&o_full[w], par);
      fence_after();
      uint32_t o[32];
      ld32(t_s + 64, o);
      ld_wait();
      fence_before();
      mbar_arrive(&o_read[w]);
      if (row_ok) {
        bf16* op = a.out + ((size_t)b * a.T + r) * a.d + hd * HD;
        const uint64_t inv2 = pk2(inv, inv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            upk2(mul2(pk2(__uint_as_float(o[8 * q + 2 * k]), __uint_as_float(o[8 * q + 2 * k + 1])), inv2), f[2 * k], f[2 * k + 1]);
          uint4 u;
          u.x = pack_bf2(f[0], f[1]); u.y = pack_bf2(f[2], f[3]); u.z = pack_bf2(f[4], f[5]); u.w = pack_bf2(f[6], f[7]);
          *reinterpret_cast<uint4*>(op + 8 * q) = u;
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) {
    fence_after();
    tmem_dealloc<FWD_TMEM>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------------------------
struct BwdArgs {
  int B, T, d, h, G;      // B = narratives per group
  const float* aux;       // [G*B][h][4][128]: lse * log2e | D = rowsum(dO . O) | score scale * log2e (0: masked row) | score-gradient scale (0: masked)
  bf16* dqkv;
  float* dbias;           // optional fp32 [3d] per group (dbias_gstride floats apart), accumulated
  size_t dbias_gstride;
  const uint32_t* dbits;  // optional precomputed keep bits, see FwdArgs
  DropCfg drop[MAXG];
};

constexpr int BWD_NT = 640;                                             // warps 0-15 compute, 16 TMA, 17 MMA issue + TMEM, 18-19 idle (they
                                                                            // complete the fifth warpgroup for setmaxnreg, second formulation)
constexpr int BWD1_NT = 576;                                            // first formulation (A/B): 18 warps
constexpr int BWD_REGS_COMPUTE = 112, BWD_REGS_SPECIAL = 32;               // balanced inside the CTA's own allocation: 20 x 96 = 16 x 112 + 4 x 32
constexpr int BWD_STAGES = 2;
constexpr int BWD_AUX_BYTES = 2 * 4 * TM * 4;                           // two heads x four per-query vectors
constexpr int BWD_BITS_BYTES = 2 * 4 * TM * 4;                          // two heads x 128 keep bits per query (optional)
constexpr int BWD_STAGE_BYTES = 4 * TILE_BYTES + BWD_AUX_BYTES + BWD_BITS_BYTES;      // Q | K | V | dO | aux | keep bits
constexpr int BWD_DS_BYTES = TM * TM * 2;                               // dS of one head as an MN-major A operand
constexpr int BWD_SMEM = BWD_STAGES * BWD_STAGE_BYTES + 2 * BWD_DS_BYTES + 256 + 1024;

// aux[b][hd][k][q] (row stride 128 whatever T is), see BwdArgs.  A block takes 32 consecutive token rows: each warp reads four rows'
// 512 bytes of `out` and `dout` as 16-byte pieces (fully coalesced; lane l holds 8 columns of head l / (HD / 8)), the per-head dot
// product closes with shuffles inside the head's lane group and lands in shared memory [head][row]; then the block writes the four
// per-query vectors with q (the row) fastest, i.e. as full 128-byte segments.
constexpr int PREP_ROWS = 32, PREP_MAXH = 32;
__global__ void __launch_bounds__(256) attn_tc_prep_kernel(int B, int Bg, int T, int d, int h, const bf16* __restrict__ out,
                                                           const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                           const float* __restrict__ mask, float* __restrict__ aux, float scale) {
  constexpr int LPH = HD / 8;                       // lanes per head
  __shared__ float sD[PREP_MAXH][PREP_ROWS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)B * T;           // B = all narratives of the launch, Bg = per group (the mask is shared by the groups)
  for (long long r0 = (long long)blockIdx.x * PREP_ROWS; r0 < rows; r0 += (long long)gridDim.x * PREP_ROWS) {
    for (int rr = warp; rr < PREP_ROWS; rr += 8) {
      const long long row = r0 + rr;
      if (row >= rows) continue;
      for (int c0 = 0; c0 < d; c0 += 256) {         // 32 lanes x 8 columns per pass; d % 32 == 0, so a head's lane group is all-in or all-out
        const int col = c0 + lane * 8;
        const bool valid = col < d;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const uint4 o = valid ? *reinterpret_cast<const uint4*>(out + (size_t)row * d + col) : zero;
        const uint4 g = valid ? *reinterpret_cast<const uint4*>(dout + (size_t)row * d + col) : zero;
        const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, gw[4] = {g.x, g.y, g.z, g.w};
        float D = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
          const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[k]));
          D += of.x * gf.x + of.y * gf.y;
        }
#pragma unroll
        for (int s = 1; s < LPH; s <<= 1) D += __shfl_xor_sync(0xffffffffu, D, s);
        if (valid && (lane & (LPH - 1)) == 0) sD[col / HD][rr] = D;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h * PREP_ROWS; i += 256) {
      const int hd = i / PREP_ROWS, rr = i % PREP_ROWS;
      const long long row = r0 + rr;
      if (row >= rows) continue;
      const int b = (int)(row / T), q = (int)(row % T);
      const bool masked = mask != nullptr && mask[(size_t)(b % Bg) * T + q] == 0.f;
      const size_t bh = (size_t)b * h + hd;
      float* a = aux + bh * 4 * TM;
      a[q] = lse[bh * T + q] * LOG2E;
      a[TM + q] = sD[hd][rr];
      a[2 * TM + q] = masked ? 0.f : scale * LOG2E;      // masked query rows: constant scores (uniform P) ...
      a[3 * TM + q] = masked ? 0.f : scale;              // ... and no score gradient (masked_fill blocks it)
    }
    __syncthreads();
  }
}

// the same without D (row 1 of aux already holds it: the output projection's input-gradient GEMM wrote it from its epilogue, see
// mt_gemm_rs.cu R_ATTD): one thread per (b, hd, q), q fastest -- 4 bytes read, 12 written per query
__global__ void attn_tc_prep_light_kernel(int B, int Bg, int T, int h, const float* __restrict__ lse, const float* __restrict__ mask,
                                          float* __restrict__ aux, float scale) {
  mt_pdl_gate();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * T) return;
  const int q = (int)(idx % T);
  const long long bh = idx / T;
  const int b = (int)(bh / h);
  const bool masked = mask != nullptr && mask[(size_t)(b % Bg) * T + q] == 0.f;
  float* a = aux + (size_t)bh * 4 * TM;
  a[q] = lse[bh * T + q] * LOG2E;
  a[2 * TM + q] = masked ? 0.f : scale * LOG2E;
  a[3 * TM + q] = masked ? 0.f : scale;
}

// TMEM columns of head w (base w * 256).  Warp group hf of the head owns the queries [64 hf, 64 hf + 64):
//   [0,128)   S^T  (fp32)  -> P^T  packed bf16 at [0,32) (hf 0) and [64,96) (hf 1): each group overwrites columns it has already read
//   [128,256) dP^T (fp32)  -> dS^T packed bf16 at [128,160) and [192,224)
//   accumulators of the second round in the gaps: dV [32,64), dK [96,128), dQ [160,192)
template <bool FULL>
__global__ void __launch_bounds__(BWD1_NT, 1)      // 18 warps are allocated as 20: 96 registers per thread is the ceiling
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_cs[4 * 2 * 96];                  // [head pair][w][dQ | dK | dV][32]: bias-gradient column sums of this CTA
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* ds_smem = smem + BWD_STAGES * BWD_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + 2 * BWD_DS_BYTES);
  uint64_t* full = bars;                    // [2] TMA -> MMA / compute
  uint64_t* empty = bars + BWD_STAGES;      // [2] MMA -> TMA
  uint64_t* s_full = empty + BWD_STAGES;    // [2] S^T and dP^T of head w are in TMEM
  uint64_t* p_ready = s_full + 2;           // [2] P^T / dS^T (TMEM) and dS (smem) of head w are in place (256 arrivals)
  uint64_t* g_full = p_ready + 2;           // [2] dV, dK, dQ of head w are in TMEM
  uint64_t* g_read = g_full + 2;            // [2] ... and have been read out (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_read + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp_count = a.h >> 1;
  const int T = a.T;
  for (int i = threadIdx.x; i < 4 * 2 * 96; i += BWD1_NT) s_cs[i] = 0.f;
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_ready[w], 256); mbar_init(&g_full[w], 1); mbar_init(&g_read[w], 256); }
    mbar_init_fence();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const ItemWalk iw = item_walk(a.G, a.B, hp_count);
  const int n_my = iw.n;

  if (warp == 16) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int item = iw.base + it * iw.step;
        const int b = item / hp_count, hp = item % hp_count;
        const int stage = it & 1;
        mbar_wait(&empty[stage], (((uint32_t)it >> 1) & 1u) ^ 1u);
        uint8_t* sb = stage_base + stage * BWD_STAGE_BYTES;
        mbar_expect_tx(&full[stage], (uint32_t)(4 * TILE_BYTES + BWD_AUX_BYTES));
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_2d(sb + k * TILE_BYTES, &map_qkv, k * a.d + hp * 64, b * T, &full[stage]);
        tma_load_2d(sb + 3 * TILE_BYTES, &map_do, hp * 64, b * T, &full[stage]);
        bulk_load(sb + 4 * TILE_BYTES, a.aux + ((size_t)b * a.h + 2 * hp) * 4 * TM, (uint32_t)BWD_AUX_BYTES, &full[stage]);
      }
    }
  } else if (warp == 17) {
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
              // one thread issues everything: build each operand's descriptor once, advance its 14-bit address field by a constant per k-step
              const uint64_t dk0 = make_desc(sk + 64 * w, 16, 1024), dq0 = make_desc(sq + 64 * w, 16, 1024);
              const uint64_t dv0 = make_desc(sv + 64 * w, 16, 1024), dg0 = make_desc(sg + 64 * w, 16, 1024);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks) mma_ss(tw, dk0 + (uint64_t)(2 * ks), dq0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks) mma_ss(tw + 128, dv0 + (uint64_t)(2 * ks), dg0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
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
              const uint64_t dgm = make_desc(sg + 64 * w, 8192, 1024), dqm = make_desc(sq + 64 * w, 8192, 1024);
              const uint64_t dsm = make_desc(sds, TILE_BYTES, 1024), dkm = make_desc(sk + 64 * w, 8192, 1024);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks)      // packed A columns of 16 queries: [8 ks, 8 ks + 8) for ks < 4, 64 + ... beyond
                mma_ts(tw + 32, tw + (ks < 4 ? 8 * ks : 32 + 8 * ks), dgm + (uint64_t)(128 * ks), idesc_ts, ks > 0);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks)
                mma_ts(tw + 96, tw + 128 + (ks < 4 ? 8 * ks : 32 + 8 * ks), dqm + (uint64_t)(128 * ks), idesc_ts, ks > 0);
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks) mma_ss(tw + 160, dsm + (uint64_t)(128 * ks), dkm + (uint64_t)(128 * ks), idesc_dq, ks > 0);
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
    // ===== compute: warp group g = 2 w + hf: head 2 hp + w, queries [64 hf, 64 hf + 64), thread = key row j =====
    const DropCfg drop = mt_drop_resolve(a.drop[iw.grp]);
    const int g = warp >> 2, w = g >> 1, hf = g & 1, j = threadIdx.x & 127;
    const uint32_t tw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(w * 256);
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    const bool key_ok = FULL || j < T;
    const uint32_t odd = (uint32_t)j & 1u;
    uint8_t* dsb = ds_smem + w * BWD_DS_BYTES + hf * TILE_BYTES;      // this group's 64-query block of the MN-major dS operand
    const uint64_t ds2 = pk2(drop.scale, drop.scale);
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int item = iw.base + it * iw.step;
      const int b = item / hp_count, hp = item % hp_count, hd = 2 * hp + w;      // b: global narrative (group * B + local)
      const int stage = it & 1;
      const float* ax = reinterpret_cast<const float*>(stage_base + stage * BWD_STAGE_BYTES + 4 * TILE_BYTES) + w * 4 * TM;
      // pair index of (query q, key pair j >> 1) = (bh T + q) P2 + (j >> 1); this lane draws for the queries q + (j & 1)
      const uint32_t bh = (uint32_t)((b - iw.grp * a.B) * a.h + hd);      // group-local: the dropout stream of a stack does not depend on its group slot
      const uint32_t pidx0 = (bh * (uint32_t)T + (uint32_t)(hf * 64) + odd) * P2 + (uint32_t)(j >> 1);
      mbar_wait(&full[stage], ((uint32_t)it >> 1) & 1u);       // aux vectors (the MMA warp waits on the same phase for the tiles)
      mbar_wait(&s_full[w], par);
      fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int q0 = hf * 64 + cc * 16;
        uint32_t s[16], dp[16], pkp[8], pks[8];
        ld16(tw + (uint32_t)q0, s);
        ld16(tw + (uint32_t)(128 + q0), dp);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int q = q0 + i;
          const float4 L = *reinterpret_cast<const float4*>(ax + q), D = *reinterpret_cast<const float4*>(ax + TM + q);
          const float4 rs = *reinterpret_cast<const float4*>(ax + 2 * TM + q), gs = *reinterpret_cast<const float4*>(ax + 3 * TM + q);
          const float Dk[4] = {D.x, D.y, D.z, D.w};
          float p[4], t[4];
          upk2(fma2(pk2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), pk2(rs.x, rs.y), pk2(-L.x, -L.y)), p[0], p[1]);
          upk2(fma2(pk2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), pk2(rs.z, rs.w), pk2(-L.z, -L.w)), p[2], p[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k) p[k] = (key_ok && (FULL || q + k < T)) ? ex2(p[k]) : 0.f;
          // t = dP . keep-scale - D ; without a kept draw the probability's gradient is -D
          upk2(fma2(pk2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), ds2, pk2(-D.x, -D.y)), t[0], t[1]);
          upk2(fma2(pk2(__uint_as_float(dp[i + 2]), __uint_as_float(dp[i + 3])), ds2, pk2(-D.z, -D.w)), t[2], t[3]);
          float pd[4] = {p[0], p[1], p[2], p[3]};
          if (dropping) {
            // the 32-bit draw of (query, key pair j >> 1) serves keys j and j ^ 1: this lane draws for query q + (j & 1), its
            // neighbour for the other query of the pair, and they swap
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
              const uint32_t mine = mt_mix32((pidx0 + (uint32_t)(cc * 16 + i + k) * P2) ^ drop.key);
              const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
              const uint32_t b0 = odd ? other : mine, b1 = odd ? mine : other;
              const bool k0 = odd ? (b0 >= thr_hi) : ((b0 << 16) >= thr_hi), k1 = odd ? (b1 >= thr_hi) : ((b1 << 16) >= thr_hi);
              pd[k] = k0 ? p[k] : 0.f;
              t[k] = k0 ? t[k] : -Dk[k];
              pd[k + 1] = k1 ? p[k + 1] : 0.f;
              t[k + 1] = k1 ? t[k + 1] : -Dk[k + 1];
            }
          }
          float e[4];
          upk2(mul2(mul2(pk2(p[0], p[1]), pk2(gs.x, gs.y)), pk2(t[0], t[1])), e[0], e[1]);
          upk2(mul2(mul2(pk2(p[2], p[3]), pk2(gs.z, gs.w)), pk2(t[2], t[3])), e[2], e[3]);
          pkp[i >> 1] = pack_bf2(pd[0], pd[1]);
          pkp[(i >> 1) + 1] = pack_bf2(pd[2], pd[3]);
          pks[i >> 1] = pack_bf2(e[0], e[1]);
          pks[(i >> 1) + 1] = pack_bf2(e[2], e[3]);
        }
        st8(tw + (uint32_t)(hf * 64 + cc * 8), pkp);
        st8(tw + (uint32_t)(128 + hf * 64 + cc * 8), pks);
        // dS as the MN-major A operand of dQ = dS K: k row = key j, 64 queries per 128-byte row
        *reinterpret_cast<uint4*>(dsb + sw128_off(j, cc * 2)) = make_uint4(pks[0], pks[1], pks[2], pks[3]);
        *reinterpret_cast<uint4*>(dsb + sw128_off(j, cc * 2 + 1)) = make_uint4(pks[4], pks[5], pks[6], pks[7]);
      }
      st_wait();
      fence_proxy_async();
      fence_before();
      mbar_arrive(&p_ready[w]);
      // ---- gradients: dV | dK rows = keys, dQ rows = queries; all rows beyond T are exact zeros.  Group 0 of the head takes
      //      dQ and the first half of dK, group 1 dV (which still lacks the keep scale) and the second half of dK ----
      mbar_wait(&g_full[w], par);
      fence_after();
      uint32_t gm[32], gk[16];
      ld32(tw + (hf ? 32u : 160u), gm);
      ld16(tw + (uint32_t)(96 + hf * 16), gk);
      ld_wait();
      fence_before();
      mbar_arrive(&g_read[w]);
      float v[48];
      {
        const uint64_t sc2 = hf ? ds2 : pk2(1.f, 1.f);
#pragma unroll
        for (int i = 0; i < 32; i += 2) upk2(mul2(pk2(__uint_as_float(gm[i]), __uint_as_float(gm[i + 1])), sc2), v[i], v[i + 1]);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[32 + i] = __uint_as_float(gk[i]);
      }
      if (key_ok) {
        bf16* gp = a.dqkv + ((size_t)b * T + j) * (3 * a.d) + hd * HD;
        bf16* gmain = gp + (hf ? 2 * a.d : 0);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint4 u;
          u.x = pack_bf2(v[8 * qd], v[8 * qd + 1]); u.y = pack_bf2(v[8 * qd + 2], v[8 * qd + 3]);
          u.z = pack_bf2(v[8 * qd + 4], v[8 * qd + 5]); u.w = pack_bf2(v[8 * qd + 6], v[8 * qd + 7]);
          __stcs(reinterpret_cast<uint4*>(gmain + 8 * qd), u);      // streaming store: leaves the few KB of L1 beside 200 KB of shared memory alone
        }
        bf16* gkp = gp + a.d + hf * 16;
#pragma unroll
        for (int qd = 0; qd < 2; ++qd) {
          uint4 u;
          u.x = pack_bf2(v[32 + 8 * qd], v[33 + 8 * qd]); u.y = pack_bf2(v[34 + 8 * qd], v[35 + 8 * qd]);
          u.z = pack_bf2(v[36 + 8 * qd], v[37 + 8 * qd]); u.w = pack_bf2(v[38 + 8 * qd], v[39 + 8 * qd]);
          __stcs(reinterpret_cast<uint4*>(gkp + 8 * qd), u);
        }
      }
      if (a.dbias != nullptr) {
        // column sums over the warp's 32 rows: a butterfly that halves the number of live columns per lane at every step
        // (the lane index is re-read here: keeping the kernel-wide `lane` alive across this 48-value block made ptxas spill it)
        int lane;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        int col = 0;
#pragma unroll
        for (int step = 0; step < 4; ++step) {
          const int n2 = 24 >> step;                // 24, 12, 6, 3
          const int m = 16 >> step;
          const bool up = (lane & m) != 0;
#pragma unroll
          for (int i = 0; i < n2; ++i) {
            const float send = up ? v[i] : v[i + n2], keep = up ? v[i + n2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
          }
          col += up ? n2 : 0;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
        if ((lane & 1) == 0) {
          float* cs = s_cs + ((hp & 3) * 2 + w) * 96;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int cidx = col + i;      // < 32: column of dQ (hf 0) / dV (hf 1); else column cidx - 32 + 16 hf of dK
            atomicAdd(cs + (cidx < 32 ? (hf ? 64 : 0) + cidx : 32 + 16 * hf + (cidx - 32)), v[i]);
          }
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
  if (a.dbias != nullptr) {
    // s_cs[hp][w][k][c] -> dbias[k * d + (2 hp + w) * 32 + c]   (h <= 8: at most 4 head pairs)
    for (int i = threadIdx.x; i < hp_count * 2 * 96; i += BWD1_NT) {
      const int hp = i / 192, w = (i / 96) & 1, k = (i % 96) / 32, c = i % 32;
      const float val = s_cs[i];
      if (val != 0.f) atomicAdd(a.dbias + iw.grp * a.dbias_gstride + (size_t)k * a.d + (2 * hp + w) * HD + c, val);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// backward, second formulation (default): the block pipeline of the long-sequence kernel (mt_attention_flash.cu) for T <= 128.
// An item is still a (narrative, head pair) whose four TMA boxes share a ring stage, but its two heads are two BLOCKS of one pipeline:
// all 16 compute warps work on the same head (warp group g: queries 32 g .. 32 g + 31, thread = key row), P^T goes to its own TMEM
// columns and dS^T to a double-buffered shared-memory tile (K-major A operand of dK, MN-major A operand of dQ), so S^T / dP^T of the
// next block are issued in front of the gradient MMAs of this one and the compute warps never wait for them; dQ / dK / dV of a block
// are drained (8 columns per thread) while the next block's gradient MMAs run.
// TMEM columns: S^T [0,128) | dP^T [128,256) | P^T packed bf16 [256,320) | dV [320,352) | dK [352,384) | dQ [384,416)
template <bool FULL, bool BITS>
__global__ void __launch_bounds__(BWD_NT, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_cs[4 * 4 * 2 * 96];              // [warp & 3][head pair][w][dQ | dK | dV][32]: bias-gradient column sums of this CTA, one slice per
                                                      // row quarter so that every address has ONE writing warp (plain adds, no CAS loops)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* ds_smem = smem + BWD_STAGES * BWD_STAGE_BYTES;      // two dS^T buffers (block parity)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + 2 * BWD_DS_BYTES);
  uint64_t* full = bars;                    // [2] TMA -> MMA / compute
  uint64_t* empty = bars + BWD_STAGES;      // [2] MMA -> TMA
  uint64_t* s_full = empty + BWD_STAGES;    // [1] S^T and dP^T of the block are in TMEM
  uint64_t* s_free = s_full + 1;            // [1] ... and have been read (512 arrivals)
  uint64_t* p_ready = s_free + 1;           // [1] P^T (TMEM) and dS^T (shared) of the block are in place (512 arrivals)
  uint64_t* pt_free = p_ready + 1;          // [1] the dV MMAs of the block are complete: P^T may be overwritten
  uint64_t* ds_free = pt_free + 1;          // [2] the dK / dQ MMAs that read this dS^T buffer are complete
  uint64_t* g_full = ds_free + 2;           // [1] dV, dK, dQ of the block are in TMEM
  uint64_t* g_read = g_full + 1;            // [1] ... and have been read out (512 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_read + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp_count = a.h >> 1;
  const int T = a.T;
  for (int i = threadIdx.x; i < 4 * 4 * 2 * 96; i += BWD_NT) s_cs[i] = 0.f;
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 512); mbar_init(p_ready, 512); mbar_init(pt_free, 1);
    mbar_init(&ds_free[0], 1); mbar_init(&ds_free[1], 1); mbar_init(g_full, 1); mbar_init(g_read, 512);
    mbar_init_fence();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  mt_pdl_gate();      // everything above touches no global memory
  // register re-allocation between the warpgroups: the 16 compute warps are register-starved at the launch ceiling of 96 (20 resident
  // warps), the TMA / MMA issuing threads need a fraction of that
  // (the setmaxnreg instructions head the role branches below: the register budget of a region follows the one that dominates it)
  const ItemWalk iw = item_walk(a.G, a.B, hp_count);
  const int n_my = iw.n;

  if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BWD_REGS_SPECIAL));
  if (warp == 16) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int item = iw.base + it * iw.step;
        const int b = item / hp_count, hp = item % hp_count;
        const int stage = it & 1;
        mbar_wait(&empty[stage], (((uint32_t)it >> 1) & 1u) ^ 1u);
        uint8_t* sb = stage_base + stage * BWD_STAGE_BYTES;
        mbar_expect_tx(&full[stage], (uint32_t)(4 * TILE_BYTES + BWD_AUX_BYTES + (BITS ? BWD_BITS_BYTES : 0)));
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_2d(sb + k * TILE_BYTES, &map_qkv, k * a.d + hp * 64, b * T, &full[stage]);
        tma_load_2d(sb + 3 * TILE_BYTES, &map_do, hp * 64, b * T, &full[stage]);
        bulk_load(sb + 4 * TILE_BYTES, a.aux + ((size_t)b * a.h + 2 * hp) * 4 * TM, (uint32_t)BWD_AUX_BYTES, &full[stage]);
        if (BITS) bulk_load(sb + 4 * TILE_BYTES + BWD_AUX_BYTES, a.dbits + ((size_t)b * a.h + 2 * hp) * 4 * TM, (uint32_t)BWD_BITS_BYTES, &full[stage]);
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(TM, TM, 0, 0);
      const uint32_t idesc_ts = make_idesc(TM, HD, 0, 1);      // dV: A in TMEM; dK: A K-major in shared memory; B MN-major
      const uint32_t idesc_dq = make_idesc(TM, HD, 1, 1);      // dQ: A = dS from shared memory, MN-major
      const uint32_t st_u32 = smem_u32(stage_base), ds_u32 = smem_u32(ds_smem);
      const int n_blk = 2 * n_my;
      auto round2 = [&](int pb) {           // gradient MMAs of block pb
        const int stage = (pb >> 1) & 1, w = pb & 1;
        const uint32_t sq = st_u32 + (uint32_t)(stage * BWD_STAGE_BYTES), sk = sq + TILE_BYTES, sg = sk + 2 * TILE_BYTES;
        const uint32_t sds = ds_u32 + (uint32_t)((pb & 1) * BWD_DS_BYTES);
        mbar_wait(p_ready, (uint32_t)pb & 1u);
        if (pb > 0) mbar_wait(g_read, (uint32_t)(pb - 1) & 1u);      // dV / dK / dQ of the previous block have been read out
        fence_after();
        const uint64_t dgm = make_desc(sg + 64 * w, 8192, 1024), dqm = make_desc(sq + 64 * w, 8192, 1024), dkm = make_desc(sk + 64 * w, 8192, 1024);
#pragma unroll
        for (int ks = 0; ks < TM / 16; ++ks)      // dV = P^T dO: A = P^T in TMEM (8 columns per 16 queries)
          mma_ts(tmem_base + 320, tmem_base + (uint32_t)(256 + 8 * ks), dgm + (uint64_t)(128 * ks), idesc_ts, ks > 0);
        commit(pt_free);
#pragma unroll
        for (int ks = 0; ks < TM / 16; ++ks)      // dK = dS^T Q: A K-major over the queries, two 64-query tiles
          mma_ss(tmem_base + 352, make_desc(sds + (uint32_t)((ks >> 2) * TILE_BYTES + (ks & 3) * 32), 16, 1024), dqm + (uint64_t)(128 * ks), idesc_ts, ks > 0);
        {
          const uint64_t dsm = make_desc(sds, TILE_BYTES, 1024);
#pragma unroll
          for (int ks = 0; ks < TM / 16; ++ks) mma_ss(tmem_base + 384, dsm + (uint64_t)(128 * ks), dkm + (uint64_t)(128 * ks), idesc_dq, ks > 0);
        }
        commit(g_full);
        commit(&ds_free[pb & 1]);
        if (w == 1) commit(&empty[stage]);      // both heads of the item are done with the stage
      };
      for (int blk = 0; blk < n_blk; ++blk) {
        const int it = blk >> 1, w = blk & 1, stage = it & 1;
        if (w == 0) { mbar_wait(&full[stage], ((uint32_t)it >> 1) & 1u); }
        if (blk > 0) mbar_wait(s_free, (uint32_t)(blk - 1) & 1u);
        fence_after();
        const uint32_t sq = st_u32 + (uint32_t)(stage * BWD_STAGE_BYTES), sk = sq + TILE_BYTES, sv = sk + TILE_BYTES, sg = sv + TILE_BYTES;
        const uint64_t dk0 = make_desc(sk + 64 * w, 16, 1024), dq0 = make_desc(sq + 64 * w, 16, 1024);
        const uint64_t dv0 = make_desc(sv + 64 * w, 16, 1024), dg0 = make_desc(sg + 64 * w, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) mma_ss(tmem_base, dk0 + (uint64_t)(2 * ks), dq0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) mma_ss(tmem_base + 128, dv0 + (uint64_t)(2 * ks), dg0 + (uint64_t)(2 * ks), idesc_s, ks > 0);
        commit(s_full);
        if (blk > 0) round2(blk - 1);
      }
      if (n_blk > 0) round2(n_blk - 1);
    }
  } else if (warp < 16) {
    // ===== compute: warp group g owns the queries [32 g, 32 g + 32) of every block, thread = key row j =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BWD_REGS_COMPUTE));
    const DropCfg drop = mt_drop_resolve(a.drop[iw.grp]);
    const int g = warp >> 2, j = threadIdx.x & 127;
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t P2 = (uint32_t)(T + 1) >> 1;
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    const bool dropping = drop.thresh != 0u;
    const bool key_ok = FULL || j < T;
    const uint32_t odd = (uint32_t)j & 1u;
    const uint64_t ds2 = pk2(drop.scale, drop.scale);
    const int n_blk = 2 * n_my;
    // read out dQ / dK / dV of block pb (8 columns of each per thread), store them, add their column sums to the bias gradient
    auto drain = [&](int pb) {
      const int it = pb >> 1, w = pb & 1;
      const int item = iw.base + it * iw.step;
      const int b = item / hp_count, hp = item % hp_count, hd = 2 * hp + w;
      mbar_wait(g_full, (uint32_t)pb & 1u);
      fence_after();
      uint32_t rq[8], rk[8], rv[8];
      ld8(tl + (uint32_t)(384 + g * 8), rq);
      ld8(tl + (uint32_t)(352 + g * 8), rk);
      ld8(tl + (uint32_t)(320 + g * 8), rv);
      ld_wait();
      fence_before();
      mbar_arrive(g_read);
      float v[24];
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = __uint_as_float(rq[i]); v[8 + i] = __uint_as_float(rk[i]); }
#pragma unroll
      for (int i = 0; i < 8; i += 2) upk2(mul2(pk2(__uint_as_float(rv[i]), __uint_as_float(rv[i + 1])), ds2), v[16 + i], v[17 + i]);      // dV still lacks the keep scale
      // pack first, reduce, store LAST: the 16-byte stores of a warp go to 32 different rows and drain slowly, and whatever overwrites
      // their source registers waits for them -- behind the butterfly nothing does until the next block's TMEM loads
      uint4 u[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        u[k].x = pack_bf2(v[8 * k], v[8 * k + 1]); u[k].y = pack_bf2(v[8 * k + 2], v[8 * k + 3]);
        u[k].z = pack_bf2(v[8 * k + 4], v[8 * k + 5]); u[k].w = pack_bf2(v[8 * k + 6], v[8 * k + 7]);
      }
      if (a.dbias != nullptr) {
        // column sums over the warp's 32 rows: a butterfly that halves the number of live columns per lane at every step
        // (lane = j & 31: reading %laneid costs an S2R round trip per block)
        int col = 0;
#pragma unroll
        for (int step = 0; step < 3; ++step) {
          const int n2 = 12 >> step;                // 12, 6, 3
          const int m = 16 >> step;
          const bool up = (j & m) != 0;
#pragma unroll
          for (int i = 0; i < n2; ++i) {
            const float send = up ? v[i] : v[i + n2], keep = up ? v[i + n2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
          }
          col += up ? n2 : 0;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
          v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
        }
        if ((j & 3) == 0) {
          float* cs = s_cs + (((warp & 3) * 4 + (hp & 3)) * 2 + w) * 96;      // this warp is the only writer of its columns in this slice
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int cidx = col + i;      // 0..23: [dQ | dK | dV] x 8 columns of this warp group
            cs[(cidx >> 3) * 32 + g * 8 + (cidx & 7)] += v[i];
          }
        }
      }
      if (key_ok) {           // row j is query j of dQ and key j of dK / dV; rows beyond T are exact zeros
        bf16* gp = a.dqkv + ((size_t)b * T + j) * (3 * a.d) + hd * HD + g * 8;
#pragma unroll
        for (int k = 0; k < 3; ++k) __stcs(reinterpret_cast<uint4*>(gp + (size_t)k * a.d), u[k]);
      }
    };
    for (int blk = 0; blk < n_blk; ++blk) {
      const int it = blk >> 1, w = blk & 1, stage = it & 1;
      const int item = iw.base + it * iw.step;
      const int b = item / hp_count, hp = item % hp_count, hd = 2 * hp + w;
      const float* ax = reinterpret_cast<const float*>(stage_base + stage * BWD_STAGE_BYTES + 4 * TILE_BYTES) + w * 4 * TM;
      // BITS: keep bits of this head, word (c, q) = keys 32 c .. 32 c + 31 of query q; this thread's key is bit (j & 31) of word c = j / 32
      const uint32_t* kbits = reinterpret_cast<const uint32_t*>(stage_base + stage * BWD_STAGE_BYTES + 4 * TILE_BYTES + BWD_AUX_BYTES) + w * 4 * TM + (j >> 5) * TM;
      const uint32_t kbit = 1u << (j & 31);
      uint8_t* dsb = ds_smem + (blk & 1) * BWD_DS_BYTES + (g >> 1) * TILE_BYTES;      // this group's 64-query tile of the dS^T operand
      // pair index of (query q, key pair j >> 1) = (bh T + q) P2 + (j >> 1), group-local bh; this lane draws for the queries q + (j & 1)
      const uint32_t bh = (uint32_t)((b - iw.grp * a.B) * a.h + hd);
      const uint32_t pidx0 = (bh * (uint32_t)T + (uint32_t)(g * 32) + odd) * P2 + (uint32_t)(j >> 1);
      if (w == 0) mbar_wait(&full[stage], ((uint32_t)it >> 1) & 1u);       // per-query vectors (the MMA warp waits on the same phase for the tiles)
      mbar_wait(s_full, (uint32_t)blk & 1u);
      fence_after();
      // The block's 32 queries go through in two halves of 16 (cc).  The TMEM loads of the second half are issued as soon as the first
      // half's arithmetic has consumed its registers, so their latency passes under the barrier waits and the P^T / dS^T stores of the
      // first half instead of in front of the second half's arithmetic.
      uint32_t s[16], dp[16], pkp[8], pks[8];
      auto load_half = [&](int cc) {
        ld16(tl + (uint32_t)(g * 32 + cc * 16), s);
        ld16(tl + (uint32_t)(128 + g * 32 + cc * 16), dp);
      };
      auto math_half = [&](int cc) {
        const int q0 = g * 32 + cc * 16;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
            const int q = q0 + e;
            const float4 L = *reinterpret_cast<const float4*>(ax + q), D = *reinterpret_cast<const float4*>(ax + TM + q);
            const float4 rs = *reinterpret_cast<const float4*>(ax + 2 * TM + q), gs = *reinterpret_cast<const float4*>(ax + 3 * TM + q);
            const float Dk[4] = {D.x, D.y, D.z, D.w};
            float p[4], t[4];
            upk2(fma2(pk2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), pk2(rs.x, rs.y), pk2(-L.x, -L.y)), p[0], p[1]);
            upk2(fma2(pk2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), pk2(rs.z, rs.w), pk2(-L.z, -L.w)), p[2], p[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = (key_ok && (FULL || q + k < T)) ? ex2(p[k]) : 0.f;
            // t = dP . keep-scale - D ; without a kept draw the probability's gradient is -D
            upk2(fma2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), ds2, pk2(-D.x, -D.y)), t[0], t[1]);
            upk2(fma2(pk2(__uint_as_float(dp[e + 2]), __uint_as_float(dp[e + 3])), ds2, pk2(-D.z, -D.w)), t[2], t[3]);
            float pd[4] = {p[0], p[1], p[2], p[3]};
            if (BITS || dropping) {           // keep bits exist only when the launch drops: no run-time branch around their loads
              if (BITS) {
                const uint4 kw = *reinterpret_cast<const uint4*>(kbits + q);
                const uint32_t kq[4] = {kw.x, kw.y, kw.z, kw.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const bool keep = (kq[k] & kbit) != 0u;
                  pd[k] = keep ? p[k] : 0.f;
                  t[k] = keep ? t[k] : -Dk[k];
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                  const uint32_t mine = mt_mix32((pidx0 + (uint32_t)(cc * 16 + e + k) * P2) ^ drop.key);
                  const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                  const uint32_t b0 = odd ? other : mine, b1 = odd ? mine : other;
                  const bool k0 = odd ? (b0 >= thr_hi) : ((b0 << 16) >= thr_hi), k1 = odd ? (b1 >= thr_hi) : ((b1 << 16) >= thr_hi);
                  pd[k] = k0 ? p[k] : 0.f;
                  t[k] = k0 ? t[k] : -Dk[k];
                  pd[k + 1] = k1 ? p[k + 1] : 0.f;
                  t[k + 1] = k1 ? t[k + 1] : -Dk[k + 1];
                }
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
      };
      auto store_half = [&](int cc) {
        st8(tl + (uint32_t)(256 + g * 16 + cc * 8), pkp);
        *reinterpret_cast<uint4*>(dsb + sw128_off(j, (g & 1) * 4 + cc * 2)) = make_uint4(pks[0], pks[1], pks[2], pks[3]);
        *reinterpret_cast<uint4*>(dsb + sw128_off(j, (g & 1) * 4 + cc * 2 + 1)) = make_uint4(pks[4], pks[5], pks[6], pks[7]);
      };
      load_half(0);
      ld_wait();
      math_half(0);
      load_half(1);
      // first writes of the block: the previous block's dV MMAs have read P^T, the dK / dQ MMAs of block blk - 2 this dS^T buffer
      if (blk > 0) mbar_wait(pt_free, (uint32_t)(blk - 1) & 1u);
      if (blk > 1) mbar_wait(&ds_free[blk & 1], (uint32_t)((blk >> 1) - 1) & 1u);
      fence_after();
      store_half(0);
      ld_wait();
      fence_before();
      mbar_arrive(s_free);      // this thread's last read of S^T / dP^T
      math_half(1);
      store_half(1);
      st_wait();
      fence_proxy_async();
      fence_before();
      mbar_arrive(p_ready);
      if (blk > 0) drain(blk - 1);      // its gradient MMAs were issued behind this block's S^T
    }
    if (n_blk > 0) drain(n_blk - 1);
  }
  fence_before();
  __syncthreads();
  if (warp == 17) {
    fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (a.dbias != nullptr) {
    for (int i = threadIdx.x; i < hp_count * 2 * 96; i += BWD_NT) {
      const int hp = i / 192, w = (i / 96) & 1, k = (i % 96) / 32, c = i % 32;
      const float val = (s_cs[i] + s_cs[4 * 2 * 96 + i]) + (s_cs[2 * 4 * 2 * 96 + i] + s_cs[3 * 4 * 2 * 96 + i]);
      if (val != 0.f) atomicAdd(a.dbias + iw.grp * a.dbias_gstride + (size_t)k * a.d + (2 * hp + w) * HD + c, val);
    }
  }
}

// Keep bits of the attention-probability dropout for a whole launch, drawn ONCE per step and layer (forward and backward read the same
// words): bits[bh][c][q], bit i of word (c, q) = key 32 c + i of query q, from exactly the pair-hash draws the kernels make themselves
// (pair index ((b_local h + hd) T + q) P2 + (j >> 1), low half -> even key).  One thread per word: 16 draws.
__global__ void __launch_bounds__(256) attn_tc_dropbits_kernel(const __grid_constant__ MtBitsArgs a, uint32_t* __restrict__ bits) {
  mt_pdl_gate();
  const MtBitsKeys k = mt_bits_resolve(a);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < k.n; idx += gridDim.x * blockDim.x) bits[idx] = mt_bits_word(a, k, idx);
}

// one-time (per device) opt-in to the large dynamic shared-memory carve-out
template <typename K>
int set_smem_attr(K kernel, int bytes, MtPerDeviceOnce& once) {
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
  if ((unsigned long long)B * h * T * ((T + 1) / 2) >= 0xffffffffULL) return false;      // 32-bit dropout pair indices
  return d % 64 == 0;
}

size_t mt_attn_tc_dropbits_words(int G, int B, int h) { return (size_t)G * B * h * 4 * TM; }

int mt_attn_tc_dropbits_job(int G, int B, int T, int h, const DropCfg* drops, uint32_t* bits, MtBitsJob* job) {
  if (G < 1 || G > MAXG || !drops || !bits || !job || T < 1 || T > TM || (unsigned long long)G * B * h * 512ull >= 0xffffffffull) return MT_ERR_ARG;
  job->a.B = B; job->a.T = T; job->a.h = h; job->a.G = G;
  for (int i = 0; i < MAXG; ++i) job->a.drop[i] = drops[i < G ? i : 0];
  job->bits = bits;
  return MT_OK;
}

int mt_attn_tc_dropbits_run(int G, int B, int T, int h, const DropCfg* drops, uint32_t* bits, cudaStream_t st) {
  if (G < 1 || G > MAXG || !drops || !bits || T < 1 || T > TM || (unsigned long long)G * B * h * 512ull >= 0xffffffffull) return MT_ERR_ARG;
  MtBitsArgs a;
  a.B = B; a.T = T; a.h = h; a.G = G;
  for (int i = 0; i < MAXG; ++i) a.drop[i] = drops[i < G ? i : 0];
  const size_t n = mt_attn_tc_dropbits_words(G, B, h);
  mt_prof_work(0.0, (double)n * 4.0);
  const size_t cap = (size_t)num_sms() * 8;
  MT_CUDA(mt_launch_dep(MT_PDL_MISC, attn_tc_dropbits_kernel, dim3((unsigned)((n + 255) / 256 < cap ? (n + 255) / 256 : cap)), dim3(256), 0, st, a, bits));
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_attn_tc_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                       const int* klen, int G, const DropCfg* drops, const uint32_t* dbits) {
  if (!mt_attn_tc_supported(B, T, d, h) || G < 1 || G > MAXG || (long long)G * B * T > 0x7fffffffLL / (3LL * d)) return MT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return MT_ERR_ALIGN;
  CUtensorMap map;
  MT_TRY(make_map_2d(&map, qkv, (uint64_t)3 * d, (uint64_t)G * B * T, (uint64_t)3 * d, 64, TM));
  FwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h; a.G = G;
  a.scale_log2 = LOG2E / sqrtf((float)HD);
  a.mask = mask; a.out = (bf16*)out; a.lse = lse; a.klen = klen; a.dbits = dbits;
  a.early_trigger = (g_mt_tune[15] & 1) ? 0 : 1;
  for (int i = 0; i < MAXG; ++i) a.drop[i] = drops && i < G ? drops[i] : drop;
  static MtPerDeviceOnce attr_full, attr_part;
  const int slots = 2 * num_sms(), n_items = B * (h / 2);
  const int cpg = n_items < slots / G ? n_items : slots / G;      // CTAs per group
  const int grid = cpg * G;
  mt_prof_work(4.0 * G * B * (double)T * T * d, (double)G * B * T * d * 4.0 * 2.0);
  static MtPerDeviceOnce attr_full_b, attr_part_b;
  if (T == TM && !klen) {
    if (dbits) { MT_TRY(set_smem_attr(attn_tc_fwd_kernel<true, true>, FWD_SMEM, attr_full_b)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_FWD, attn_tc_fwd_kernel<true, true>, dim3(grid), dim3(FWD_NT), FWD_SMEM, st, map, a)); }
    else { MT_TRY(set_smem_attr(attn_tc_fwd_kernel<true, false>, FWD_SMEM, attr_full)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_FWD, attn_tc_fwd_kernel<true, false>, dim3(grid), dim3(FWD_NT), FWD_SMEM, st, map, a)); }
  } else {
    if (dbits) { MT_TRY(set_smem_attr(attn_tc_fwd_kernel<false, true>, FWD_SMEM, attr_part_b)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_FWD, attn_tc_fwd_kernel<false, true>, dim3(grid), dim3(FWD_NT), FWD_SMEM, st, map, a)); }
    else { MT_TRY(set_smem_attr(attn_tc_fwd_kernel<false, false>, FWD_SMEM, attr_part)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_FWD, attn_tc_fwd_kernel<false, false>, dim3(grid), dim3(FWD_NT), FWD_SMEM, st, map, a)); }
  }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_attn_tc_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                       void* dqkv, DropCfg drop, float* dbias, float* aux, cudaStream_t st, int G, const DropCfg* drops, size_t dbias_gstride,
                       int d_ready, const uint32_t* dbits) {
  if (!mt_attn_tc_supported(B, T, d, h) || G < 1 || G > MAXG || (long long)G * B * T > 0x7fffffffLL / (3LL * d)) return MT_ERR_UNSUPPORTED;
  if (dbias != nullptr && h > 8) return MT_ERR_UNSUPPORTED;
  if (!aux || ((uintptr_t)aux & 15) || ((uintptr_t)qkv & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)dqkv & 15) || ((uintptr_t)out & 15))
    return MT_ERR_ALIGN;
  const float scale = 1.0f / sqrtf((float)HD);
  if (d_ready == 2) {
    // all four rows of the per-query scalars are the caller's (mt_gemm_rs.cu R_ATTD with attd_lse): no preparation launch
  } else if (d_ready) {
    const long long n = (long long)G * B * h * T;
    MT_CUDA(mt_launch_dep(MT_PDL_MISC, attn_tc_prep_light_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, G * B, B, T, h, lse, mask, aux, scale));
    MT_LAUNCH_CHECK();
  } else {
    const long long rows = (long long)G * B * T;
    if (h > PREP_MAXH) return MT_ERR_UNSUPPORTED;
    long long blocks = (rows + PREP_ROWS - 1) / PREP_ROWS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    attn_tc_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(G * B, B, T, d, h, (const bf16*)out, (const bf16*)dout, lse, mask, aux, scale);
    MT_LAUNCH_CHECK();
  }
  CUtensorMap map_qkv, map_do;
  MT_TRY(make_map_2d(&map_qkv, qkv, (uint64_t)3 * d, (uint64_t)G * B * T, (uint64_t)3 * d, 64, TM));
  MT_TRY(make_map_2d(&map_do, dout, (uint64_t)d, (uint64_t)G * B * T, (uint64_t)d, 64, TM));
  BwdArgs a;
  a.B = B; a.T = T; a.d = d; a.h = h; a.G = G;
  a.aux = aux; a.dqkv = (bf16*)dqkv; a.dbias = dbias; a.dbias_gstride = dbias_gstride; a.dbits = dbits;
  for (int i = 0; i < MAXG; ++i) a.drop[i] = drops && i < G ? drops[i] : drop;
  if (a.drop[0].thresh == 0u) { a.dbits = nullptr; dbits = nullptr; }      // keep bits mean something only when the launch drops
  static MtPerDeviceOnce attr_full, attr_part;
  const int sms = num_sms(), n_items = B * (h / 2);
  const int cpg = n_items < sms / G ? n_items : sms / G;
  const int grid = cpg * G;
  mt_prof_work(10.0 * G * B * (double)T * T * d, (double)G * B * T * d * 8.0 * 2.0);
  static MtPerDeviceOnce attr2_full, attr2_part;
  if (g_mt_tune[4] == 3) {                    // A/B hook: the first formulation (two heads side by side, P^T / dS^T over S^T / dP^T)
    if (T == TM) {
      MT_TRY(set_smem_attr(attn_tc_bwd_kernel<true>, BWD_SMEM, attr_full));
      attn_tc_bwd_kernel<true><<<grid, BWD1_NT, BWD_SMEM, st>>>(map_qkv, map_do, a);
    } else {
      MT_TRY(set_smem_attr(attn_tc_bwd_kernel<false>, BWD_SMEM, attr_part));
      attn_tc_bwd_kernel<false><<<grid, BWD1_NT, BWD_SMEM, st>>>(map_qkv, map_do, a);
    }
  } else if (T == TM) {
    static MtPerDeviceOnce attr2_full_b;
    if (dbits) { MT_TRY(set_smem_attr(attn_tc_bwd2_kernel<true, true>, BWD_SMEM, attr2_full_b)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_BWD, attn_tc_bwd2_kernel<true, true>, dim3(grid), dim3(BWD_NT), BWD_SMEM, st, map_qkv, map_do, a)); }
    else { MT_TRY(set_smem_attr(attn_tc_bwd2_kernel<true, false>, BWD_SMEM, attr2_full)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_BWD, attn_tc_bwd2_kernel<true, false>, dim3(grid), dim3(BWD_NT), BWD_SMEM, st, map_qkv, map_do, a)); }
  } else {
    static MtPerDeviceOnce attr2_part_b;
    if (dbits) { MT_TRY(set_smem_attr(attn_tc_bwd2_kernel<false, true>, BWD_SMEM, attr2_part_b)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_BWD, attn_tc_bwd2_kernel<false, true>, dim3(grid), dim3(BWD_NT), BWD_SMEM, st, map_qkv, map_do, a)); }
    else { MT_TRY(set_smem_attr(attn_tc_bwd2_kernel<false, false>, BWD_SMEM, attr2_part)); MT_CUDA(mt_launch_dep(MT_PDL_ATTN_BWD, attn_tc_bwd2_kernel<false, false>, dim3(grid), dim3(BWD_NT), BWD_SMEM, st, map_qkv, map_do, a)); }
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
/* direct entry to the tcgen05 forward (tests / probes); MT_ERR_UNSUPPORTED outside its envelope */
int mt_attention_tc_fwd(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, float p_drop, uint64_t seed,
                        uint32_t site, const int* key_len, void* stream) {
  return mt_attn_tc_fwd_run(B, T, d, h, qkv, mask, out, lse, mt_make_drop(p_drop, seed, site), (cudaStream_t)stream, key_len);
}
}
