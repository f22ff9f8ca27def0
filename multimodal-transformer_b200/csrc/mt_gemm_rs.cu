// Row-stream tcgen05 GEMM for the encoder's projections (sm_100a): weight-RESIDENT, grouped over the modality stacks, TMA in and TMA out.
//
// Every projection of the encoder stacks (MFT/multiTransformer.py:19-20 FFN, :47-65 Q/K/V/out projections, and their input
// gradients) multiplies a tall activation matrix [G*Mg, K] by a small per-stack weight matrix (K, N <= 768): arithmetic intensity
// is below the B200 ridge for all of them, so the roof is HBM -- and, one step earlier, the ~2x-HBM L2 -> SM fabric: a kernel that
// re-fetches its weight slice for every row tile spends that budget on weights.  Here a CTA is pinned to ONE (stack, column slice)
// pair for its whole life:
//   * the weight slice [BN x K] of that pair is loaded ONCE into shared memory (<= 128 KB) and stays;
//   * the CTA streams 128-row activation tiles of its stack through a TMA ring of 16 KB k-blocks (one elected thread issues
//     tcgen05.mma M128 x BN x 16 into a double-buffered TMEM accumulator, so tile i+1's MMAs overlap tile i's epilogue);
//   * the epilogue (two warp groups, alternating 64-column pairs of chunks of the same tile) works ROW-PER-THREAD straight out of TMEM
//     (tcgen05.ld 32 columns at a time: lane = row): bias / ReLU / pair-hash
//     dropout / ReLU-dropout gate / fp32 residual are applied in registers, the result is written into a 128-byte-swizzled
//     staging box and leaves with ONE TMA store per box; residual and gate tiles arrive the same way (TMA boxes, own producer warp,
//     own ring) -- no thread ever issues a global load or store, so HBM sees only full 128-byte lines;
//   * column sums (bias gradients) use a halving butterfly over the warp's 32 rows, one shared-memory atomic per column per warp;
//   * with N == BN == 256 (the model width) a thread owns a COMPLETE output row, so the LayerNorm that follows the projection in
//     the reference (SublayerConnection, MFT/multiTransformer.py:93-104 -> LayerNorm :81-91) is fused: the fp32 row goes back to TMEM,
//     shifted single-pass moments give mean / unbiased std, and a second sweep emits the normalised bf16 operand of the next GEMM.
// Grouping: the three modality stacks have identical shapes, so their rows are laid out back to back ([G*Mg, .]) and one launch
// serves all of them; only the weight / bias / dropout-key / LayerNorm-gain pointers are per group.
#include "mt_gemm_rs.cuh"
#include "mt_tcgen05.cuh"

namespace {

using namespace tc5;

// ---- thread-block-cluster helpers of the LayerNorm that spans a CTA pair (R_LNX) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t raddr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int BOX = BM * 128;            // every staged box is [128 rows x 128 bytes] = 16 KB
constexpr int NT = 384;                  // warp 0: A / weight TMA, warp 1: MMA issue + TMEM, warp 2: residual / gate TMA, warps 4-7 / 8-11: epilogue
constexpr int EPI_THREADS = 128;         // per epilogue warp group
constexpr int MAXG = MT_RS_MAX_GROUPS;
constexpr int SMEM_LIMIT = 227 * 1024;

enum : uint32_t { R_BIAS = 1u, R_RELU = 2u, R_DROP = 4u, R_GATE = 8u, R_RES = 16u, R_CF32 = 32u, R_COLSUM = 64u, R_LN = 128u, R_ATTD = 256u, R_LNX = 512u };

struct RsMaps {
  CUtensorMap a, c, r, g, ln;
  CUtensorMap b[MAXG];
};

struct RsArgs {
  int G, N, tiles_n, tiles_per_group, rows_per_group, cnt, b_mn;
  float gate_scale, ln_eps;
  const float* bias[MAXG];
  float* colsum[MAXG];
  const float* ln_a[MAXG];
  const float* ln_b[MAXG];
  DropCfg drop[MAXG];
  float* attd_aux; int attd_T;      // R_ATTD: see RsDesc
  const float* attd_lse; const float* attd_mask; float attd_scale;   // R_ATTD, optional: the other three per-query scalars
  unsigned long long* trace;      // debug (mt_gemm_rs_trace): clock64 stamps of CTA 0, 16 words per tile, first 32 tiles
};

template <int BN, int KB, uint32_t F>
struct Cfg {
  static constexpr int W_BYTES = KB * BN * 128;
  static constexpr int AUX_BYTES = 1024 + 4 * BN * 4 + 4096 + 2048;   // barriers | bias | column sums | LayerNorm a_2, b_2 | row moments | the peer CTA's moments (R_LNX)
  static constexpr int BOXES = (SMEM_LIMIT - 1024 - AUX_BYTES - W_BYTES) / BOX;      // 16 KB boxes left beside the resident weights
  static constexpr bool TIGHT = BOXES < 8;                        // the 96 / 128 KB weight slices
  static constexpr int NSO = (TIGHT || (F & (R_RES | R_GATE | R_ATTD))) ? 1 : 2;   // output staging boxes PER epilogue warp group
  static constexpr int NSR = (F & R_RES) ? 4 : 0;                 // residual ring (fp32 [128 x 32] boxes)
  static constexpr int NSG = (F & (R_GATE | R_ATTD)) ? 2 : 0;     // gate / attention-output ring (bf16 [128 x 64] boxes)
  static constexpr int FREE = BOXES - 2 * NSO - NSR - NSG;
  static constexpr int NSA = FREE > 8 ? 8 : FREE;                 // activation ring
  static constexpr int TOTAL = 1024 + W_BYTES + (NSA + 2 * NSO + NSR + NSG) * BOX + AUX_BYTES;
  static constexpr int CHUNKS = BN / 32;
  static constexpr int PAIRS = BN / 64;
  static_assert(NSA >= 2, "no room for the activation ring");
  static_assert(2 * BN <= 512, "two accumulators must fit in TMEM");
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);      // allocation: a power of two
};

template <int BN, int KB, uint32_t F>
__global__ void __launch_bounds__(NT, 1) gemm_rs_kernel(const __grid_constant__ RsMaps maps, const __grid_constant__ RsArgs g) {
  using C = Cfg<BN, KB, F>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w_s = smem;
  uint8_t* a_ring = w_s + C::W_BYTES;
  uint8_t* o_stage = a_ring + C::NSA * BOX;
  uint8_t* r_ring = o_stage + 2 * C::NSO * BOX;
  uint8_t* g_ring = r_ring + C::NSR * BOX;
  uint8_t* aux = g_ring + C::NSG * BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* full_a = bars;                 // [8]
  uint64_t* empty_a = bars + 8;            // [8]
  uint64_t* acc_full = bars + 16;          // [2]
  uint64_t* acc_empty = bars + 18;         // [2]
  uint64_t* w_full = bars + 20;            // [1]
  uint64_t* full_r = bars + 24;            // [4]
  uint64_t* empty_r = bars + 28;           // [4]
  uint64_t* full_g = bars + 32;            // [4]
  uint64_t* empty_g = bars + 36;           // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  float* bias_s = reinterpret_cast<float*>(aux + 1024);
  float* cs_s = bias_s + BN;
  float* lna_s = cs_s + BN;
  float* lnb_s = lna_s + BN;
  float2* mom_s = reinterpret_cast<float2*>(lnb_s + BN);      // [tile parity][warp group][row] (mean, M2) of a row's column half
  float2* mom_x = mom_s + 2 * 2 * BM;                         // R_LNX: [tile parity][row] (mean, M2) of the PEER CTA's BN columns, written by the peer
  uint64_t* x_full = bars + 44;                               // [2] R_LNX: the peer's moments of this tile parity have arrived (128 remote arrivals)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's (group, column slice) pair and its share of the group's row tiles
  const int P = g.G * g.tiles_n;
  const int pair = (int)blockIdx.x % P, rank = (int)blockIdx.x / P;
  const int grp = pair / g.tiles_n, tn = pair % g.tiles_n;
  const int n0 = tn * BN;
  const int n_tiles = rank < g.tiles_per_group ? (g.tiles_per_group - rank + g.cnt - 1) / g.cnt : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a); tma_prefetch_desc(&maps.b[grp]); tma_prefetch_desc(&maps.c);
    for (int s = 0; s < C::NSA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    mbar_init(w_full, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&full_r[s], 1); mbar_init(&empty_r[s], 4); mbar_init(&full_g[s], 1); mbar_init(&empty_g[s], 4); }
    if (F & R_LNX) { mbar_init(&x_full[0], EPI_THREADS); mbar_init(&x_full[1], EPI_THREADS); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  if (F & R_LNX) cluster_sync_all();      // the peer's barriers exist before anything is sent to them
  const uint32_t tmem_base = *tmem_slot;
  mt_pdl_gate();
  // (the producer / MMA warps start at once; bias, column-sum and LayerNorm-gain staging is the epilogue warps' own business, below)

  if (warp == 0) {
    // ===== weight slice once, then the activation ring =====
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(w_full, (uint32_t)C::W_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        uint8_t* sb = w_s + kb * (BN * 128);
        if (g.b_mn) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &maps.b[grp], n0 + 64 * j, kb * BK, w_full);
        } else {
          tma_load_2d(sb, &maps.b[grp], kb * BK, n0, w_full);
        }
      }
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < n_tiles; ++i) {
        const int row0 = grp * g.rows_per_group + (rank + i * g.cnt) * BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_a[stage], phase ^ 1);
          if (g.trace && blockIdx.x == 0 && i < 32 && kb < 4) g.trace[i * 16 + kb] = clock64();
          mbar_expect_tx(&full_a[stage], BOX);
          tma_load_2d(a_ring + stage * BOX, &maps.a, kb * BK, row0, &full_a[stage]);
          if (++stage == C::NSA) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && n_tiles > 0) {
      const uint32_t b_mn = (uint32_t)g.b_mn;
      const uint32_t idesc = make_idesc(BM, BN, 0, (int)b_mn);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lbo = 1u << 16, b_lbo = (b_mn ? (uint32_t)((BK * 128) >> 4) : 1u) << 16;
      const uint32_t a_kstep = 2u, b_kstep = b_mn ? (uint32_t)((16 * 128) >> 4) : 2u;
      const uint32_t ring_u32 = smem_u32(a_ring), w_u32 = smem_u32(w_s);
      mbar_wait(w_full, 0);
      fence_after();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_tiles; ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_after();
        if (g.trace && blockIdx.x == 0 && it < 32) g.trace[it * 16 + 8] = clock64();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0u;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_a[stage], phase);
          fence_after();
          if (g.trace && blockIdx.x == 0 && it < 32 && kb < 4) g.trace[it * 16 + 4 + kb] = clock64();
          const uint32_t sa = ring_u32 + (uint32_t)(stage * BOX), sb = w_u32 + (uint32_t)(kb * BN * 128);
          const uint32_t a_lo = a_lbo | ((sa >> 4) & 0x3FFFu), b_lo = b_lbo | ((sb >> 4) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)k * a_kstep);
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)k * b_kstep);
            mma_ss(tmem_d, ad, bd, idesc, accum);
            accum = 1u;
          }
          commit(&empty_a[stage]);
          if (kb == KB - 1) commit(&acc_full[acc]);
          if (++stage == C::NSA) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===== residual / gate boxes, in the order the epilogue consumes them =====
    if ((F & (R_RES | R_GATE | R_ATTD)) && lane == 0) {
      int rs = 0, gs = 0; uint32_t rph = 0, gph = 0;
      for (int i = 0; i < n_tiles; ++i) {
        const int row0 = grp * g.rows_per_group + (rank + i * g.cnt) * BM;
        for (int c = 0; c < C::CHUNKS; ++c) {
          if ((F & (R_GATE | R_ATTD)) && (c & 1) == 0) {
            mbar_wait(&empty_g[gs], gph ^ 1);
            mbar_expect_tx(&full_g[gs], BOX);
            tma_load_2d(g_ring + gs * BOX, &maps.g, n0 + 32 * c, row0, &full_g[gs]);
            if (++gs == (C::NSG ? C::NSG : 1)) { gs = 0; gph ^= 1; }
          }
          if (F & R_RES) {
            mbar_wait(&empty_r[rs], rph ^ 1);
            mbar_expect_tx(&full_r[rs], BOX);
            tma_load_2d(r_ring + rs * BOX, &maps.r, n0 + 32 * c, row0, &full_r[rs]);
            if (++rs == (C::NSR ? C::NSR : 1)) { rs = 0; rph ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: two warp groups; thread = one row of the tile (TMEM lane), 32 columns per pass.  Warp group w owns the chunk
    // pairs (64 columns) pr with pr % 2 == w of every tile, its own staging box(es), named barrier and TMA-store thread. =====
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool storer = (threadIdx.x & (EPI_THREADS - 1)) == 0;
    const int bar_id = 1 + wg;
    uint8_t* my_stage = o_stage + wg * C::NSO * BOX;
    DropCfg drop = g.drop[grp];
    if (F & R_DROP) drop = mt_drop_resolve(drop);
    const uint32_t thr_hi = (drop.thresh >> 16) << 16;
    int ob = 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int i = (int)threadIdx.x - 4 * 32; i < BN; i += 2 * EPI_THREADS) {
      bias_s[i] = (F & R_BIAS) ? g.bias[grp][n0 + i] : 0.f;
      cs_s[i] = 0.f;
      if (F & R_LN) { lna_s[i] = g.ln_a[grp][n0 + i]; lnb_s[i] = g.ln_b[grp][n0 + i]; }
    }
    bar_sync(3, 2 * EPI_THREADS);
    for (int it = 0; it < n_tiles; ++it) {
      const int tm = rank + it * g.cnt;
      const int row0 = grp * g.rows_per_group + tm * BM;
      const int acc = it & 1;
      // R_ATTD with attd_lse: the scalars of this warp group's first column pair are fetched in front of the accumulator wait
      float pf_lse0 = 0.f, pf_lse1 = 0.f, pf_mask = 1.f;
      if ((F & R_ATTD) && g.attd_lse != nullptr && tm * BM + row < g.rows_per_group) {
        const int m = tm * BM + row, bl = m / g.attd_T, tq = m - bl * g.attd_T;
        const size_t bh = ((size_t)grp * (g.rows_per_group / g.attd_T) + bl) * (size_t)(g.N >> 5) + (size_t)((n0 >> 5) + 2 * wg);
        pf_lse0 = __ldg(g.attd_lse + bh * (size_t)g.attd_T + tq);
        pf_lse1 = __ldg(g.attd_lse + (bh + 1) * (size_t)g.attd_T + tq);
        if (g.attd_mask != nullptr) pf_mask = __ldg(g.attd_mask + (size_t)bl * g.attd_T + tq);
      }
      mbar_wait(&acc_full[acc], (uint32_t)(it >> 1) & 1u);
      fence_after();
      const bool tr = g.trace && blockIdx.x == 0 && it < 32 && storer;
      if (tr) g.trace[it * 16 + 9 + 2 * wg] = clock64();
      const uint32_t tacc = lane_base + (uint32_t)(acc * BN);
      float s1 = 0.f, s2 = 0.f, x0 = 0.f;
#pragma unroll 1
      for (int pr = wg; pr < C::PAIRS; pr += 2) {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * pr + cc;
          const bool box_start = (F & R_CF32) ? true : cc == 0;
          const bool box_end = (F & R_CF32) ? true : cc == 1;
          if (box_start) {          // the staging box about to be written must have been read out by its previous store
            if (storer) bulk_wait_read<C::NSO - 1>();
            bar_sync(bar_id, EPI_THREADS);
          }
          uint32_t v[32];
          ld32(tacc + (uint32_t)(c * 32), v);
          ld_wait();
          float o[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + 4 * j);
            o[4 * j] = __uint_as_float(v[4 * j]) + b4.x; o[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            o[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; o[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          if (F & R_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          if (F & R_DROP) {
            if (drop.thresh != 0u) {
              // element index (m local to the group) * N + n: even, so a pair never straddles two rows; the pair index fits 32 bits
              // (rows per group <= 2^21, see mt_gemm_rs_supported), where mt_draw32 reduces to mix32(pair ^ key)
              const uint32_t p0 = (uint32_t)(((uint64_t)(tm * BM + row) * (uint64_t)g.N + (uint64_t)(n0 + c * 32)) >> 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t bits = mt_mix32((p0 + (uint32_t)j) ^ drop.key);
                o[2 * j] = (bits << 16) >= thr_hi ? o[2 * j] * drop.scale : 0.f;      // low half -> even element, high half -> odd
                o[2 * j + 1] = bits >= thr_hi ? o[2 * j + 1] * drop.scale : 0.f;
              }
            }
          }
          if (F & R_GATE) {        // gate box use number: one per chunk pair, consumed by this warp group only
            const int ug = it * C::PAIRS + pr, gs = ug % (C::NSG ? C::NSG : 1);
            if (cc == 0) mbar_wait(&full_g[gs], (uint32_t)(ug / (C::NSG ? C::NSG : 1)) & 1u);
            const uint8_t* gb = g_ring + gs * BOX;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 gw = *reinterpret_cast<const uint4*>(gb + sw128_off(row, cc * 4 + j));
              const uint32_t w4[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[k]));
                o[8 * j + 2 * k] = g2.x > 0.f ? o[8 * j + 2 * k] * g.gate_scale : 0.f;
                o[8 * j + 2 * k + 1] = g2.y > 0.f ? o[8 * j + 2 * k + 1] * g.gate_scale : 0.f;
              }
            }
            if (cc == 1) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty_g[gs]);
            }
          }
          if (F & R_ATTD) {        // D = rowsum over this head's 32 columns of dO . O (attention backward), O from the bf16 box ring
            const int ug = it * C::PAIRS + pr, gs = ug % (C::NSG ? C::NSG : 1);
            if (cc == 0) mbar_wait(&full_g[gs], (uint32_t)(ug / (C::NSG ? C::NSG : 1)) & 1u);
            const uint8_t* gb = g_ring + gs * BOX;
            // the other three per-query scalars of the attention backward (attd_lse given): their loads fly under the D sum
            const int m = tm * BM + row;                          // row inside the group
            const bool live = m < g.rows_per_group;
            const int bl = m / g.attd_T, q = m - bl * g.attd_T;
            const size_t bh = ((size_t)grp * (g.rows_per_group / g.attd_T) + bl) * (size_t)(g.N >> 5) + (size_t)((n0 >> 5) + c);
            float lse_q = cc ? pf_lse1 : pf_lse0, mask_q = pf_mask;
            if (g.attd_lse != nullptr && live && pr != wg) lse_q = __ldg(g.attd_lse + bh * (size_t)g.attd_T + q);      // later pairs (BN > 128): in place
            float D = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 gw = *reinterpret_cast<const uint4*>(gb + sw128_off(row, cc * 4 + j));
              const uint32_t w4[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[k]));
                D = fmaf(o[8 * j + 2 * k], g2.x, D);
                D = fmaf(o[8 * j + 2 * k + 1], g2.y, D);
              }
            }
            if (cc == 1) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty_g[gs]);
            }
            if (live) {
              float* ax = g.attd_aux + bh * 4 * 128 + q;
              ax[128] = D;
              if (g.attd_lse != nullptr) {                        // what attn_tc_prep_light_kernel would write (mt_attention_tc.cu)
                constexpr float LOG2E = 1.4426950408889634f;
                const bool masked = mask_q == 0.f;
                ax[0] = lse_q * LOG2E;
                ax[2 * 128] = masked ? 0.f : g.attd_scale * LOG2E;
                ax[3 * 128] = masked ? 0.f : g.attd_scale;
              }
            }
          }
          if (F & R_RES) {         // residual box use number: one per chunk
            const int ur = it * C::CHUNKS + c, rs = ur % (C::NSR ? C::NSR : 1);
            mbar_wait(&full_r[rs], (uint32_t)(ur / (C::NSR ? C::NSR : 1)) & 1u);
            const uint8_t* rb = r_ring + rs * BOX;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r4 = *reinterpret_cast<const float4*>(rb + sw128_off(row, j));
              o[4 * j] += r4.x; o[4 * j + 1] += r4.y; o[4 * j + 2] += r4.z; o[4 * j + 3] += r4.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_r[rs]);
          }
          if (F & R_LN) {          // keep the finished fp32 row in TMEM for the normalising sweep; shifted single-pass moments
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(o[j]);
            st16(tacc + (uint32_t)(c * 32), w);
            st16(tacc + (uint32_t)(c * 32 + 16), w + 16);
            if (pr == wg && cc == 0) x0 = o[0];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float t = o[j] - x0; s1 += t; s2 = fmaf(t, t, s2); }
          }
          uint8_t* ob_s = my_stage + ob * BOX;
          if (F & R_CF32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(ob_s + sw128_off(row, j)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              pk.x = pack_bf2(o[8 * j], o[8 * j + 1]); pk.y = pack_bf2(o[8 * j + 2], o[8 * j + 3]);
              pk.z = pack_bf2(o[8 * j + 4], o[8 * j + 5]); pk.w = pack_bf2(o[8 * j + 6], o[8 * j + 7]);
              *reinterpret_cast<uint4*>(ob_s + sw128_off(row, cc * 4 + j)) = pk;
            }
          }
          if (F & R_COLSUM) {      // halving butterfly over the warp's 32 rows: lane l ends with the sum of column l
#pragma unroll
            for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                const float send = up ? o[i] : o[i + n], keep = up ? o[i + n] : o[i];
                o[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            atomicAdd(cs_s + c * 32 + lane, o[0]);
          }
          if (box_end) {
            fence_proxy_async();
            bar_sync(bar_id, EPI_THREADS);
            if (storer) {
              tma_store_2d(&maps.c, ob_s, (F & R_CF32) ? n0 + 32 * c : n0 + 64 * pr, row0);
              bulk_commit();
            }
            if (++ob == C::NSO) ob = 0;
          }
        }
      }
      if (F & R_LN) {
        // this warp group's half of the row -> (mean, M2); the two halves are merged through shared memory (Chan et al.)
        constexpr float half_n = (float)(BN / 2);
        const float m_w = x0 + s1 / half_n;
        const float M2_w = fmaxf(s2 - s1 * s1 / half_n, 0.f);
        float2* mom = mom_s + (it & 1) * 2 * BM;
        mom[wg * BM + row] = make_float2(m_w, M2_w);
        st_wait();
        bar_sync(3, 2 * EPI_THREADS);
        const float2 other = mom[(wg ^ 1) * BM + row];
        const float dm = m_w - other.x;
        float mean = 0.5f * (m_w + other.x);
        float M2 = M2_w + other.y + dm * dm * (half_n * 0.5f);
        float var = M2 * (1.0f / (float)(BN - 1));
        if (F & R_LNX) {
          // the row continues in the peer CTA of the pair (the other 128-column slice of the same row tile): exchange (mean, M2) of the
          // BN columns through distributed shared memory -- warp group 0 writes into the peer's mom_x and arrives on the peer's barrier
          const uint32_t peer = cluster_ctarank() ^ 1u;
          if (wg == 0) {
            st_cluster_f2(mapa_u32(smem_u32(&mom_x[(it & 1) * BM + row]), peer), mean, M2);
            mbar_arrive_cluster(mapa_u32(smem_u32(&x_full[it & 1]), peer));
          }
          mbar_wait_cluster(&x_full[it & 1], (uint32_t)(it >> 1) & 1u);
          const float2 px = mom_x[(it & 1) * BM + row];
          const float dx = mean - px.x;
          M2 = M2 + px.y + dx * dx * ((float)BN * 0.5f);
          mean = 0.5f * (mean + px.x);
          var = M2 * (1.0f / (float)(2 * BN - 1));
        }
        const float inv = 1.0f / (sqrtf(var) + g.ln_eps);
#pragma unroll 1
        for (int pr = wg; pr < C::PAIRS; pr += 2) {
          if (storer) bulk_wait_read<C::NSO - 1>();
          bar_sync(bar_id, EPI_THREADS);
          uint8_t* ob_s = my_stage + ob * BOX;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * pr + cc;
            uint32_t v[32];
            ld32(tacc + (uint32_t)(c * 32), v);
            ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float y[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int col = c * 32 + 8 * j + k;
                y[k] = fmaf(lna_s[col] * inv, __uint_as_float(v[8 * j + k]) - mean, lnb_s[col]);
              }
              uint4 pk;
              pk.x = pack_bf2(y[0], y[1]); pk.y = pack_bf2(y[2], y[3]); pk.z = pack_bf2(y[4], y[5]); pk.w = pack_bf2(y[6], y[7]);
              *reinterpret_cast<uint4*>(ob_s + sw128_off(row, cc * 4 + j)) = pk;
            }
          }
          fence_proxy_async();
          bar_sync(bar_id, EPI_THREADS);
          if (storer) {
            tma_store_2d(&maps.ln, ob_s, n0 + 64 * pr, row0);
            bulk_commit();
          }
          if (++ob == C::NSO) ob = 0;
        }
      }
      if (tr) g.trace[it * 16 + 10 + 2 * wg] = clock64();
      // every TMEM read of this tile by this warp is done: hand the accumulator back to the MMA warp
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
    if (storer) bulk_wait_all<0>();
  }
  fence_before();
  __syncthreads();
  if (F & R_LNX) cluster_sync_all();      // neither CTA of the pair leaves while the other may still write into its shared memory
  if (F & R_COLSUM) {
    for (int i = threadIdx.x; i < BN; i += NT) atomicAdd(g.colsum[grp] + n0 + i, cs_s[i]);
  }
  if (warp == 1) {
    fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

unsigned long long* g_rs_trace = nullptr;

int num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

template <int BN, int KB, uint32_t F>
int launch(const RsDesc& d, cudaStream_t st) {
  using C = Cfg<BN, KB, F>;
  static_assert(C::TOTAL <= SMEM_LIMIT, "shared memory budget");
  RsMaps maps;
  RsArgs g;
  const size_t rows = (size_t)d.G * d.Mg;
  MT_TRY(make_map_2d(&maps.a, d.A, (uint64_t)d.K, rows, (uint64_t)d.lda, BK, BM));
  if (d.c_f32) MT_TRY(make_map_2d(&maps.c, d.C, (uint64_t)d.N, rows, (uint64_t)d.ldc, 32, BM, 4));
  else MT_TRY(make_map_2d(&maps.c, d.C, (uint64_t)d.N, rows, (uint64_t)d.ldc, 64, BM));
  maps.r = maps.c; maps.g = maps.c; maps.ln = maps.c;
  if (F & R_RES) MT_TRY(make_map_2d(&maps.r, d.residual, (uint64_t)d.N, rows, (uint64_t)d.ldr, 32, BM, 4));
  if (F & R_GATE) MT_TRY(make_map_2d(&maps.g, d.gate, (uint64_t)d.N, rows, (uint64_t)d.ldg, 64, BM));
  if (F & R_ATTD) MT_TRY(make_map_2d(&maps.g, d.attd_src, (uint64_t)d.N, rows, (uint64_t)d.attd_ld, 64, BM));
  if (F & R_LN) MT_TRY(make_map_2d(&maps.ln, d.ln_out, (uint64_t)d.N, rows, (uint64_t)d.ld_ln, 64, BM));
  for (int i = 0; i < MAXG; ++i) {
    const int s = i < d.G ? i : 0;
    if (d.b_kmajor) MT_TRY(make_map_2d(&maps.b[i], d.B[s], (uint64_t)d.K, (uint64_t)d.N, (uint64_t)d.ldb, BK, BN));
    else MT_TRY(make_map_2d(&maps.b[i], d.B[s], (uint64_t)d.N, (uint64_t)d.K, (uint64_t)d.ldb, 64, BK));
    g.bias[i] = d.bias[s]; g.colsum[i] = d.colsum[s]; g.ln_a[i] = d.ln_a[s]; g.ln_b[i] = d.ln_b[s]; g.drop[i] = d.drop[s];
  }
  g.G = d.G; g.N = d.N; g.tiles_n = d.N / BN;
  g.tiles_per_group = (d.Mg + BM - 1) / BM;
  g.rows_per_group = d.Mg;
  g.b_mn = d.b_kmajor ? 0 : 1;
  g.gate_scale = d.gate_scale; g.ln_eps = d.ln_eps;
  g.attd_aux = d.attd_aux; g.attd_T = d.attd_T > 0 ? d.attd_T : 1;
  g.attd_lse = d.attd_lse; g.attd_mask = d.attd_mask; g.attd_scale = d.attd_scale;
  g.trace = g_rs_trace;
  const int P = d.G * g.tiles_n;
  int cnt = num_sms() / P;
  if (cnt < 1) cnt = 1;
  if (cnt > g.tiles_per_group) cnt = g.tiles_per_group;
  g.cnt = cnt;
  static MtPerDeviceOnce once;
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(gemm_rs_kernel<BN, KB, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cnt * P)); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = C::TOTAL; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = mt_pdl_enabled(st, MT_PDL_GEMM_RS);
  cfg.attrs = at; cfg.numAttrs = 1;
  if (F & R_LNX) {      // the two column slices of a row tile are adjacent CTAs: one cluster
    if (g.tiles_n != 2) return MT_ERR_UNSUPPORTED;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  MT_CUDA(cudaLaunchKernelEx(&cfg, gemm_rs_kernel<BN, KB, F>, maps, g));
  MT_LAUNCH_CHECK();
  return MT_OK;
}

uint32_t features(const RsDesc& d) {
  uint32_t f = 0;
  if (d.bias[0]) f |= R_BIAS;
  if (d.act == MT_ACT_RELU) f |= R_RELU;
  if (d.drop[0].thresh != 0u) f |= R_DROP;
  if (d.gate) f |= R_GATE;
  if (d.residual) f |= R_RES;
  if (d.c_f32) f |= R_CF32;
  if (d.colsum[0]) f |= R_COLSUM;
  if (d.ln_out) f |= (d.N == 256 && d.K == 256) ? (R_LN | R_LNX) : R_LN;      // K = 256: two 128-column slices, the LayerNorm spans the CTA pair
  if (d.attd_aux) f |= R_ATTD;
  return f;
}

// the instantiations of the encoder path: (N, K, feature set) -> (BN, KB)
int dispatch(const RsDesc& d, cudaStream_t st, bool probe_only) {
  const uint32_t f = features(d);
#define RS_PROBE(N_, K_, BN_, F_) \
  if (d.N == (N_) && d.K == (K_) && f == (uint32_t)(F_)) return probe_only ? MT_OK : launch<BN_, (K_) / 64, (uint32_t)(F_)>(d, st)
  // forward
  if (g_mt_tune[4] == 1) { RS_PROBE(768, 256, 128, R_BIAS); }              // experiment: 6 narrow slices, deep ring
  if (g_mt_tune[4] == 2) { RS_PROBE(768, 256, 192, R_BIAS); }              // experiment: 4 slices of 192 columns
  RS_PROBE(768, 256, 256, R_BIAS);                                         // QKV projection
  RS_PROBE(256, 256, 128, R_BIAS | R_DROP | R_RES | R_CF32);               // output projection, train
  RS_PROBE(256, 256, 128, R_BIAS | R_RES | R_CF32);                        // output projection, eval
  RS_PROBE(256, 256, 128, R_BIAS | R_DROP | R_RES | R_CF32 | R_LN | R_LNX);   // output projection + LayerNorm 2 across a CTA pair, train
  RS_PROBE(256, 256, 128, R_BIAS | R_RES | R_CF32 | R_LN | R_LNX);            // ... eval
  RS_PROBE(128, 256, 128, R_BIAS | R_RELU | R_DROP);                       // FFN w_1, train
  RS_PROBE(128, 256, 128, R_BIAS | R_RELU);                                // FFN w_1, eval
  RS_PROBE(256, 128, 256, R_BIAS | R_DROP | R_RES | R_CF32);               // FFN w_2, train
  RS_PROBE(256, 128, 256, R_BIAS | R_RES | R_CF32);                        // FFN w_2, eval
  RS_PROBE(256, 128, 256, R_BIAS | R_DROP | R_RES | R_CF32 | R_LN);        // FFN w_2 + next LayerNorm, train
  RS_PROBE(256, 128, 256, R_BIAS | R_RES | R_CF32 | R_LN);                 // FFN w_2 + next LayerNorm, eval
  // input gradients (weights read transposed in place)
  RS_PROBE(128, 256, 128, R_GATE | R_COLSUM);                              // d hidden = (d out . w_2) gated by relu' / dropout, + d b_1
  RS_PROBE(256, 128, 256, 0u);                                             // d LN2-out = d hidden . w_1
  RS_PROBE(256, 256, 256, 0u);                                             // d att = d out . w_o
  RS_PROBE(256, 256, 128, R_ATTD);                                         // ... that also leaves D = rowsum(d att . att) per head for the attention backward
#undef RS_PROBE
  return MT_ERR_UNSUPPORTED;
}

}  // namespace

bool mt_gemm_rs_supported(const RsDesc& d) {
  if (d.G < 1 || d.G > MAXG || d.Mg <= 0 || !d.A || !d.C) return false;
  if (d.G > 1 && d.Mg % BM != 0) return false;
  if ((size_t)d.G * d.Mg > 0x7fffffffull / 1024) return false;
  for (int i = 0; i < d.G; ++i) {
    if (!d.B[i] || ((uintptr_t)d.B[i] & 15)) return false;
    if ((d.bias[0] != nullptr) != (d.bias[i] != nullptr) || (d.colsum[0] != nullptr) != (d.colsum[i] != nullptr)) return false;
    if ((d.drop[0].thresh != 0u) != (d.drop[i].thresh != 0u)) return false;
    if (d.ln_out && (!d.ln_a[i] || !d.ln_b[i])) return false;
  }
  if (((uintptr_t)d.A & 15) || ((uintptr_t)d.C & 15) || d.lda % 8 != 0 || d.ldb % 8 != 0) return false;
  if (d.ldc % (d.c_f32 ? 4 : 8) != 0) return false;
  if (d.residual && (((uintptr_t)d.residual & 15) || d.ldr % 4 != 0)) return false;
  if (d.gate && (((uintptr_t)d.gate & 15) || d.ldg % 8 != 0)) return false;
  if (d.ln_out && (((uintptr_t)d.ln_out & 15) || d.ld_ln % 8 != 0 || !d.c_f32)) return false;
  if (d.attd_aux && (!d.attd_src || ((uintptr_t)d.attd_src & 15) || d.attd_ld % 8 != 0 || d.attd_T <= 0 || d.attd_T > 128 || d.Mg % d.attd_T != 0 ||
                     d.gate || d.c_f32))
    return false;
  return dispatch(d, nullptr, true) == MT_OK;
}

int mt_gemm_rs_run(const RsDesc& d, cudaStream_t st) {
  if (!mt_gemm_rs_supported(d)) return MT_ERR_UNSUPPORTED;
  if (g_mt_prof_on) {
    char tag[48];
    snprintf(tag, sizeof(tag), "rs g%d m%d n%d k%d %c%s", d.G, d.Mg, d.N, d.K, d.b_kmajor ? 'K' : 'M', d.ln_out ? " +ln" : "");
    mt_prof_tag(tag);
    const double rows = (double)d.G * d.Mg;
    mt_prof_work(2.0 * rows * d.N * d.K, rows * d.K * 2.0 + (double)d.G * d.N * d.K * 2.0 + rows * d.N * (d.c_f32 ? 4.0 : 2.0) +
                                             (d.residual ? 4.0 * rows * d.N : 0.0) + (d.gate ? 2.0 * rows * d.N : 0.0) +
                                             (d.ln_out ? 2.0 * rows * d.N : 0.0));
  }
  return dispatch(d, st, false);
}

extern "C" {

/* debug hook: CTA 0 of every row-stream GEMM writes clock64 stamps (16 x uint64 per tile, first 32 tiles: TMA issue of k-blocks 0-3,
 * their arrival as seen by the MMA thread, accumulator free, epilogue start / end of the two warp groups) into dev_buf (>= 4 KB) */
int mt_gemm_rs_trace(void* dev_buf) { g_rs_trace = (unsigned long long*)dev_buf; return MT_OK; }

/* Probe / test entry of the row-stream engine: G groups of Mg rows; B = G weight matrices back to back ([N,K] K-major when b_kmajor,
 * else [K,N]); bias / colsum G x N back to back (may be NULL); drop_p > 0 applies output dropout with per-group sites site + 512 * g;
 * ln_a / ln_b G x N (LayerNorm fused when ln_out != NULL). */
int mt_gemm_rs(int G, int Mg, int N, int K, const void* A, const void* B, int b_kmajor, void* C, int c_f32, const float* bias, int act,
               float drop_p, uint64_t seed, uint32_t site, const void* gate, float gate_scale, const float* residual, float* colsum,
               void* ln_out, const float* ln_a, const float* ln_b, void* stream) {
  if (G < 1 || G > MAXG) return MT_ERR_ARG;
  RsDesc d;
  d.G = G; d.Mg = Mg; d.N = N; d.K = K;
  d.A = A; d.lda = K;
  d.ldb = b_kmajor ? K : N; d.b_kmajor = b_kmajor != 0;
  d.C = C; d.ldc = N; d.c_f32 = c_f32 != 0;
  d.act = act;
  d.gate = gate; d.ldg = N; d.gate_scale = gate_scale;
  d.residual = residual; d.ldr = N;
  d.ln_out = ln_out; d.ld_ln = N;
  for (int i = 0; i < G; ++i) {
    d.B[i] = (const bf16*)B + (size_t)i * N * K;
    d.bias[i] = bias ? bias + (size_t)i * N : nullptr;
    d.colsum[i] = colsum ? colsum + (size_t)i * N : nullptr;
    d.drop[i] = mt_make_drop(drop_p, seed, site + 512u * (uint32_t)i);
    d.ln_a[i] = ln_a ? ln_a + (size_t)i * N : nullptr;
    d.ln_b[i] = ln_b ? ln_b + (size_t)i * N : nullptr;
  }
  return mt_gemm_rs_run(d, (cudaStream_t)stream);
}

}  // extern "C"
