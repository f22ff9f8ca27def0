// tcgen05 GEMM engine (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   C[m,n] = epilogue( sum_k A(m,k) B(n,k) )      A, B each K-major or MN-major (see mt_gemm.cuh)
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0  : TMA producer   -- cp.async.bulk.tensor 2-D boxes (128-byte swizzle) into a shared-memory ring
//   warp 1  : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16) from shared-memory
//                               descriptors; tcgen05.commit releases ring slots and publishes accumulators
//   warps 2-9: epilogue      -- each owns a TMEM lane quarter x a column half: tcgen05.ld the fp32 accumulator
//                               (double-buffered in TMEM so the next tile's MMAs overlap this tile's epilogue),
//                               release it, transpose 32-column passes through shared memory, then stream rows: the
//                               residual / gate loads of a pass are issued before any is consumed, and bias /
//                               activation / dropout / gate / residual / row-mask are applied on fully coalesced
//                               row segments, 8 columns per lane (vector fp32 reductions for split-K wgrad work items).
// Two ring modes:
//   streaming      : every k-block brings an A box and a B box (5 stages of 32 KB) -- long K, split-K wgrads.
//   weight-resident: K <= 256 (every projection of the encoder forward and most dgrads): the CTA is pinned to ONE column
//                    slice of the output, loads that slice of the weight matrix (<= 64 KB) once, keeps it in shared memory
//                    for all its row tiles, and the ring (6 stages of 16 KB) carries activations only.  This halves
//                    the L2 -> SM traffic of these HBM-bound GEMMs and doubles the look-ahead of the ring.
// Out-of-bounds rows / columns / k are zero-filled by TMA and predicated in the epilogue.
#include <cuda.h>

#include "mt_gemm.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per k-block = one 128-byte swizzle row
// epilogue warps = 4 TMEM lane quarters x (2 | 4) column slices, each slice drained in passes of PW columns through a staging
// buffer.  The 256-wide tile uses 16 warps (the epilogue, not the MMA issue loop, bounded it with 8) and 16-column passes so
// that the staging still fits next to the 128 KB resident weight slice.
// MT = 128-row subtiles of a CTA tile (1 or 2).  MT = 2: a 256-row tile whose two tcgen05.mma per k-step share the B operand -- halves
// the operand re-reads of the L2-bound shapes (split-K weight gradients, the long-K input gradient of the QKV projection).
template <int BN, int MT = 1> struct EpiCfg {
  static constexpr int EW = (BN > 128 && MT == 1) ? 16 : 8;
  static constexpr int PW = (BN > 128 || MT == 2) ? 16 : 32;
};
constexpr int KB_RES = 4;              // weight-resident mode: at most this many k-blocks (K <= 256)
constexpr int MAX_STAGES = 8;
constexpr uint32_t SPIN_LIMIT = 1u << 27;

struct TcMapsB { CUtensorMap m[4]; };      // B operand per row group (one entry unless TcArgs.mgroups > 1)

struct TcArgs {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_total, kb_per;
  int a_mn, b_mn;                      // operand major-ness (1 = MN-major)
  void* C; int ldc; int c_f32; int atomic;
  GemmEpi epi;
  int gate_bf16;
  int mgroups, tiles_m_group;          // row groups (input gradients of several stacks): rows of group g = tiles [g * tiles_m_group, +tiles_m_group) use B map g
  int groups, kb_group;                // grouped split-K (wgrads of several modality stacks): group g contracts k-blocks [g*kb_group, +kb_total)
  long long c_gstride;                 // ... into C + g * c_gstride (fp32 elements)
  unsigned long long* trace;           // debug: per-tile clock64 stamps of CTA 0 (mt_gemm_debug_trace), else null
};

// ---- PTX wrappers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();      // never hang the GPU: a protocol bug becomes a launch failure
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // LayoutType::SWIZZLE_128B
  return d;
}

// OCC = CTAs per SM.  OCC = 2 (two independent pipelines per SM, streaming ring of 2 stages each): the single TMA / MMA issuing
// threads and the epilogue warps are all latency-bound, so a second resident CTA fills their bubbles.
template <int BN, int OCC, int MT = 1>
struct Smem {
  static constexpr int A_BYTES = MT * BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int RING_BYTES = OCC == 2 ? 2 * (A_BYTES + B_BYTES) : (MT == 2 ? (BN > 128 ? 192 : 144) : (BN > 128 ? 176 : 160)) * 1024;   // resident weights + ring, or ring only
  static constexpr int NS_STREAM = RING_BYTES / (A_BYTES + B_BYTES);                    // 6 (BN = 64) / 5 (128) / 3 (256)
  static constexpr int NS_RES = OCC == 2 ? 1 : (RING_BYTES - KB_RES * B_BYTES) / A_BYTES;   // 8 / 6 / 3 (unused with OCC = 2)
  static constexpr int EW = EpiCfg<BN, MT>::EW, PW = EpiCfg<BN, MT>::PW;
  static constexpr int NACC = 2 * MT * BN <= 512 ? 2 : 1;                               // accumulator buffers in TMEM
  static constexpr int TMEM_COLS = NACC * MT * BN <= 128 ? 128 : (NACC * MT * BN <= 256 ? 256 : 512);
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int EPI_BYTES = EW * 32 * (PW + 4) * 4;                              // per epilogue warp: 32 rows x (PW + 4) floats
  static constexpr int BAR_BYTES = 256;
  static constexpr int CS_COLS = 1024;                                                  // epi.colsum: per-CTA column accumulators (N <= CS_COLS)
  static constexpr int TOTAL = RING_BYTES + EPI_BYTES + BAR_BYTES + CS_COLS * 4 + 1024 /* alignment slack */;
  static_assert(NS_STREAM <= MAX_STAGES && NS_RES <= MAX_STAGES, "barrier arrays");
};

// Epilogue feature bits.  The hot shapes of the encoder / MFN path get their own instantiation (every test below folds at
// compile time); F_GENERIC keeps all features behind run-time tests of the argument block.
enum : uint32_t {
  F_BIAS = 1u, F_RELU = 2u, F_TANH = 4u, F_DROP = 8u, F_GATE = 16u, F_RES = 32u, F_ROWMASK = 64u, F_CF32 = 128u, F_ATOMIC = 256u,
  F_COLSUM = 512u, F_ALPHA = 1024u, F_EDGE = 2048u, F_RUNTIME = 4096u
};
constexpr uint32_t F_GENERIC = 0x1FFFu;
#define HAS(bit, rt) (((F) & (bit)) != 0u && ((((F) & F_RUNTIME) == 0u) || (rt)))

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// work item i of this CTA -> (m tile, n tile, split); false when the CTA is done.  Enumerated identically by all roles.
template <bool RES>
__device__ __forceinline__ bool get_work(const TcArgs& g, int i, int& tm, int& tn, int& split, int& grp) {
  grp = 0;
  if (RES) {                     // pinned to column slice blockIdx.x % tiles_n; row tiles dealt among the CTAs of that slice
    tn = (int)blockIdx.x % g.tiles_n;
    const int rank = (int)blockIdx.x / g.tiles_n, cnt = ((int)gridDim.x - tn + g.tiles_n - 1) / g.tiles_n;
    tm = rank + i * cnt;
    split = 0;
    return tm < g.tiles_m;
  }
  const int tiles = g.tiles_m * g.tiles_n, per = tiles * g.splits;
  int w = (int)blockIdx.x + i * (int)gridDim.x;
  if (w >= per * g.groups) return false;
  grp = w / per;
  w -= grp * per;
  const int tile = w % tiles;
  split = w / tiles;
  tm = tile / g.tiles_n; tn = tile % g.tiles_n;
  return true;
}

template <int BN, uint32_t F, bool RES, int OCC, int MT>
__global__ void __launch_bounds__(Smem<BN, OCC, MT>::THREADS, OCC)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ TcMapsB maps_b, const __grid_constant__ TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1 KB aligned, still a shared-space pointer
  using S = Smem<BN, OCC, MT>;
  static_assert(!(RES && OCC == 2), "weight-resident mode needs the whole SM");
  static_assert(!(RES && MT == 2), "weight-resident mode uses 128-row tiles");
  constexpr int TM_ROWS = MT * BM;
  constexpr int NS = RES ? S::NS_RES : S::NS_STREAM;
  constexpr int SLOT = RES ? S::A_BYTES : S::A_BYTES + S::B_BYTES;
  uint8_t* res_b = smem;                                            // RES: resident weight slice, KB_RES boxes
  uint8_t* ring = RES ? smem + KB_RES * S::B_BYTES : smem;
  float* epi_stage = reinterpret_cast<float*>(smem + S::RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::RING_BYTES + S::EPI_BYTES);
  uint64_t* full = bars;                         // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty = bars + MAX_STAGES;           // [MAX_STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * MAX_STAGES;    // [2]           MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2]           epilogue -> MMA
  uint64_t* b_full = acc_empty + 2;              // [1]           resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  float* cs_smem = reinterpret_cast<float*>(smem + S::RING_BYTES + S::EPI_BYTES + S::BAR_BYTES);
  if (HAS(F_COLSUM, g.epi.colsum != nullptr)) for (int i = threadIdx.x; i < g.N; i += blockDim.x) cs_smem[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps_b.m[0]) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], S::EW); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(S::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run while the
  // previous kernel of the stream is still draining; from here on global memory is touched, so wait for it to complete.
  mt_pdl_gate();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int a_mn = g.a_mn, b_mn = g.b_mn, kb_per = g.kb_per, kb_total = g.kb_total;
      const bool tracing = g.trace != nullptr && blockIdx.x == 0;
      int tm, tn, split, grp;
      if (RES && get_work<RES>(g, 0, tm, tn, split, grp)) {          // the weight slice of this CTA's column tile, once
        mbar_expect_tx(b_full, (uint32_t)(kb_total * S::B_BYTES));
        for (int kb = 0; kb < kb_total; ++kb) {
          uint8_t* sb = res_b + kb * S::B_BYTES;
          if (b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &maps_b.m[0], tn * BN + 64 * j, kb * BK, b_full);
          } else {
            tma_load_2d(sb, &maps_b.m[0], kb * BK, tn * BN, b_full);
          }
        }
      }
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; get_work<RES>(g, i, tm, tn, split, grp); ++i) {
        const int m0 = tm * TM_ROWS, n0 = tn * BN;
        const CUtensorMap* map_b = &maps_b.m[g.mgroups > 1 ? tm / g.tiles_m_group : 0];
        const int kbase = grp * g.kb_group;
        const int kb0 = kbase + split * kb_per, kb1 = min(kbase + kb_total, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (tracing && kb == kb0 && i < 64) g.trace[i * 8 + 0] = clock64();
          if (tracing && i < 16 && kb - kb0 < 4) g.trace[512 + (i * 4 + kb - kb0) * 2] = clock64();
          uint8_t* sa = ring + stage * SLOT;
          mbar_expect_tx(&full[stage], SLOT);
          const int k0 = kb * BK;
          if (a_mn) {
#pragma unroll
            for (int j = 0; j < TM_ROWS / 64; ++j) tma_load_2d(sa + j * (BK * 128), &map_a, m0 + 64 * j, k0, &full[stage]);
          } else {
#pragma unroll
            for (int j = 0; j < MT; ++j) tma_load_2d(sa + j * (BM * BK * 2), &map_a, k0, m0 + BM * j, &full[stage]);
          }
          if (!RES) {
            uint8_t* sb = sa + S::A_BYTES;
            if (b_mn) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), map_b, n0 + 64 * j, k0, &full[stage]);
            } else {
              tma_load_2d(sb, map_b, k0, n0, &full[stage]);
            }
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // ONE thread issues everything, so every instruction here is exposed latency: all loop-invariant descriptor fields are
    // built once (a shared-memory descriptor only differs in its 14-bit start-address field), the per-k advance is one add
    if (lane == 0) {
      const uint32_t a_mn = (uint32_t)g.a_mn, b_mn = (uint32_t)g.b_mn;
      const int kb_per = g.kb_per, kb_total = g.kb_total;
      const bool tracing = g.trace != nullptr && blockIdx.x == 0;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // descriptor = [start addr >> 4 : 14][LBO >> 4 : 14 @16] | hi: [SBO >> 4 : 14][version 1 @14][SWIZZLE_128B = 2 @29]
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lbo = (a_mn ? (uint32_t)((BK * 128) >> 4) : 1u) << 16, b_lbo = (b_mn ? (uint32_t)((BK * 128) >> 4) : 1u) << 16;
      const uint32_t a_kstep = a_mn ? (uint32_t)((16 * 128) >> 4) : 2u, b_kstep = b_mn ? (uint32_t)((16 * 128) >> 4) : 2u;   // 16 k-elements, in 16-byte units
      const uint32_t ring_u32 = smem_u32(ring), resb_u32 = smem_u32(res_b);
      int stage = 0; uint32_t phase = 0;
      int tm, tn, split, grp;
      if (RES && get_work<RES>(g, 0, tm, tn, split, grp)) { mbar_wait(b_full, 0); tc_fence_after(); }
      for (int it = 0; get_work<RES>(g, it, tm, tn, split, grp); ++it) {
        const int kbase = grp * g.kb_group;
        const int kb0 = kbase + split * kb_per, kb1 = min(kbase + kb_total, kb0 + kb_per);
        const int acc = it % S::NACC;
        const uint32_t acc_phase = (uint32_t)(it / S::NACC) & 1;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MT * BN);
        if (tracing && it < 64) g.trace[it * 8 + 1] = clock64();
        uint32_t accum = 0u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (tracing && it < 64 && kb == kb0) g.trace[it * 8 + 2] = clock64();
          if (tracing && it < 64 && kb == kb1 - 1) g.trace[it * 8 + 3] = clock64();
          if (tracing && it < 16 && kb - kb0 < 4) g.trace[512 + (it * 4 + kb - kb0) * 2 + 1] = clock64();
          const uint32_t sa = ring_u32 + (uint32_t)(stage * SLOT);
          const uint32_t sb = RES ? resb_u32 + (uint32_t)((kb - kbase) * S::B_BYTES) : sa + (uint32_t)S::A_BYTES;
          const uint32_t a_lo = a_lbo | ((sa >> 4) & 0x3FFFu), b_lo = b_lbo | ((sb >> 4) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)k * b_kstep);
#pragma unroll
            for (int j = 0; j < MT; ++j) {      // the subtiles' A blocks are (BM * BK * 2) bytes apart in either layout; they share B
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * ((BM * BK * 2) >> 4)) + (uint32_t)k * a_kstep);
              tc_mma(tmem_d + (uint32_t)(j * BN), ad, bd, idesc, accum);
            }
            accum = 1u;
          }
          tc_commit(&empty[stage]);                 // ring slot is free once these MMAs have read it
          if (kb == kb1 - 1) tc_commit(&acc_full[acc]);
          if (tracing && it < 16 && kb - kb0 < 4) g.trace[640 + it * 4 + kb - kb0] = clock64();
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps: warp w owns TMEM lane quarter (w % 4) and column slice (w - 2) / 4 of every tile =====
    constexpr int PW = S::PW;                  // columns per pass through the staging buffer
    constexpr int HN = BN / (S::EW / 4);       // columns per epilogue warp
    constexpr int NP = HN / PW;
    constexpr int LDS = PW + 4;                // staged row stride in floats (16-byte aligned)
    constexpr int LPR = PW / 8;                // lanes per output row: each lane owns 8 consecutive columns
    constexpr int RPI = 32 / LPR;              // rows covered by one warp-wide access
    constexpr int ITERS = 32 / RPI;
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stg = epi_stage + (warp - 2) * 32 * LDS;
    const GemmEpi& e = g.epi;
    const bool f_atomic = HAS(F_ATOMIC, g.atomic), f_cf32 = HAS(F_CF32, g.c_f32), f_bias = HAS(F_BIAS, e.bias != nullptr);
    const bool f_res = HAS(F_RES, e.residual != nullptr), f_gate = HAS(F_GATE, e.gate != nullptr), f_rm = HAS(F_ROWMASK, e.rowmask != nullptr);
    const bool f_colsum = HAS(F_COLSUM, e.colsum != nullptr), f_relu = HAS(F_RELU, e.act == MT_ACT_RELU), f_tanh = HAS(F_TANH, e.act == MT_ACT_TANH);
    const bool f_alpha = HAS(F_ALPHA, e.alpha != 1.0f);
    DropCfg edrop = mt_drop_resolve(e.drop);
    const bool f_drop = HAS(F_DROP, edrop.thresh != 0u);
    const int lr = lane / LPR, lc = (lane % LPR) * 8;
    int tm, tn, split, grp;
    for (int it = 0; get_work<RES>(g, it, tm, tn, split, grp); ++it) {
      const int m0 = tm * TM_ROWS;
      const int acc = it % S::NACC;
      const uint32_t acc_phase = (uint32_t)(it / S::NACC) & 1;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const bool tr = g.trace && blockIdx.x == 0 && it < 64 && warp == 2 && lane == 0;
      if (tr) g.trace[it * 8 + 4] = clock64();
#pragma unroll 1
      for (int pp = 0; pp < MT * NP; ++pp) {
        const int sub = pp / NP, p = pp - sub * NP;      // 128-row subtile, column pass
        const int row_base = m0 + sub * BM + q * 32;
        uint32_t v[PW];
        if constexpr (PW == 32) tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + sub) * BN + half * HN + p * PW), v);
        else tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + sub) * BN + half * HN + p * PW), v);
        if (pp == MT * NP - 1) {      // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
          if (tr) g.trace[it * 8 + 5] = clock64();
        }
        __syncwarp();                             // previous pass's reads of `stg` are complete
#pragma unroll
        for (int j = 0; j < PW / 4; ++j)
          *reinterpret_cast<float4*>(stg + lane * LDS + j * 4) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int n = tn * BN + half * HN + p * PW + lc;
        bool hi_ok = true;                        // columns n + 4 .. n + 7 exist (N % 4 == 0, so a quad is all-or-nothing)
        if (F & F_EDGE) {
          if (n >= g.N || row_base >= g.M) continue;
          hi_ok = n + 4 < g.N;
        }
        if (f_atomic) {                           // split-K partial sums: vector reductions into the zero-initialised fp32 C
#pragma unroll
          for (int i = 0; i < ITERS; ++i) {
            const int r = i * RPI + lr, m = row_base + r;
            if ((F & F_EDGE) && m >= g.M) continue;
            float4 a0 = *reinterpret_cast<const float4*>(stg + r * LDS + lc), a1 = *reinterpret_cast<const float4*>(stg + r * LDS + lc + 4);
            if (f_alpha) { a0.x *= e.alpha; a0.y *= e.alpha; a0.z *= e.alpha; a0.w *= e.alpha; a1.x *= e.alpha; a1.y *= e.alpha; a1.z *= e.alpha; a1.w *= e.alpha; }
            float* cp = reinterpret_cast<float*>(g.C) + (size_t)grp * g.c_gstride + (size_t)m * g.ldc + n;
            red_add_v4(cp, a0.x, a0.y, a0.z, a0.w);
            if (hi_ok) red_add_v4(cp + 4, a1.x, a1.y, a1.z, a1.w);
          }
          continue;
        }
        float bias8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) bias8[k] = 0.f;
        if (f_bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n);
          bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
          if (hi_ok) {
            const float4 b1 = *reinterpret_cast<const float4*>(e.bias + n + 4);
            bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
          }
        }
        // issue every global load of this pass first (residual / gate / row mask of the lane's ITERS rows) ...
        float4 res[ITERS][2];
        uint4 gt[ITERS];
        float rm[ITERS];
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int m = row_base + i * RPI + lr;
          res[i][0] = res[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          gt[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);      // bf16 ones: gate open
          rm[i] = 1.f;
          if (!(F & F_EDGE) || m < g.M) {
            if (f_res) {
              res[i][0] = *reinterpret_cast<const float4*>(e.residual + (size_t)m * e.ldr + n);
              if (hi_ok) res[i][1] = *reinterpret_cast<const float4*>(e.residual + (size_t)m * e.ldr + n + 4);
            }
            if (f_gate) {
              const bf16* gp = reinterpret_cast<const bf16*>(e.gate) + (size_t)m * e.ldg + n;
              if (hi_ok && (e.ldg & 7) == 0) gt[i] = *reinterpret_cast<const uint4*>(gp);
              else {
                const uint2 g0 = *reinterpret_cast<const uint2*>(gp);
                gt[i].x = g0.x; gt[i].y = g0.y;
                if (hi_ok) { const uint2 g1 = *reinterpret_cast<const uint2*>(gp + 4); gt[i].z = g1.x; gt[i].w = g1.y; }
              }
            }
            if (f_rm) rm[i] = e.rowmask[m];
          }
        }
        float cs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] = 0.f;
        // ... then stream the rows: 8 columns per lane, 16-byte (bf16) / 2 x 16-byte (fp32) stores
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int r = i * RPI + lr, m = row_base + r;
          if ((F & F_EDGE) && m >= g.M) continue;
          const float4 a0 = *reinterpret_cast<const float4*>(stg + r * LDS + lc), a1 = *reinterpret_cast<const float4*>(stg + r * LDS + lc + 4);
          float o[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          if (f_alpha) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] *= e.alpha;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += bias8[k];
          if (f_relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
          } else if (f_tanh) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = tanhf(o[k]);
          }
          if (f_drop) {            // N % 4 == 0 and n % 4 == 0: the element index is a multiple of 4
            float f[8];
            const uint64_t idx = (uint64_t)m * (uint64_t)g.N + (uint64_t)n;
            mt_drop_quad(edrop, idx, f);
            mt_drop_quad(edrop, idx + 4, f + 4);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] *= f[k];
          }
          if (f_gate) {
            const uint32_t gw[4] = {gt[i].x, gt[i].y, gt[i].z, gt[i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[k]));
              o[2 * k] = g2.x > 0.f ? o[2 * k] * e.gate_scale : 0.f;
              o[2 * k + 1] = g2.y > 0.f ? o[2 * k + 1] * e.gate_scale : 0.f;
            }
          }
          if (f_res) {
            o[0] += res[i][0].x; o[1] += res[i][0].y; o[2] += res[i][0].z; o[3] += res[i][0].w;
            o[4] += res[i][1].x; o[5] += res[i][1].y; o[6] += res[i][1].z; o[7] += res[i][1].w;
          }
          if (f_rm) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] *= rm[i];
          }
          if (f_cf32) {
            float* cp = reinterpret_cast<float*>(g.C) + (size_t)m * g.ldc + n;
            st4(cp, make_float4(o[0], o[1], o[2], o[3]));
            if (hi_ok) st4(cp + 4, make_float4(o[4], o[5], o[6], o[7]));
          } else {
            bf16* cp = reinterpret_cast<bf16*>(g.C) + (size_t)m * g.ldc + n;
            if (hi_ok && (g.ldc & 7) == 0) {
              uint4 pk;
              pk.x = pack_bf2(o[0], o[1]); pk.y = pack_bf2(o[2], o[3]); pk.z = pack_bf2(o[4], o[5]); pk.w = pack_bf2(o[6], o[7]);
              *reinterpret_cast<uint4*>(cp) = pk;
            } else {
              st4(cp, make_float4(o[0], o[1], o[2], o[3]));
              if (hi_ok) st4(cp + 4, make_float4(o[4], o[5], o[6], o[7]));
            }
          }
          if (f_colsum) {
#pragma unroll
            for (int k = 0; k < 8; ++k) cs[k] += o[k];
          }
        }
        if (f_colsum) {        // lanes lc, lc + LPR, ... hold the same columns (and took the same `continue` decisions above)
          const unsigned am = __activemask();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) cs[k] += __shfl_xor_sync(am, cs[k], o);
          }
          if (lr == 0) {       // shared-memory accumulation across this CTA's tiles and warps; one global atomic per column at exit
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < 4 || hi_ok) atomicAdd(cs_smem + n + k, cs[k]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (HAS(F_COLSUM, g.epi.colsum != nullptr)) for (int i = threadIdx.x; i < g.N; i += blockDim.x) atomicAdd(g.epi.colsum + i, cs_smem[i]);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// operand(r, k): K-major  -> global tensor {K (inner), rows}, box {64, box_rows}
//                MN-major -> global tensor {rows (inner), K},  box {64, 64}
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld, bool kmajor, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MT_ERR_UNSUPPORTED;
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (kmajor) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = BK; box[1] = (cuuint32_t)box_rows; }
  else { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = 64; box[1] = BK; }
  strides[0] = (cuuint64_t)ld * 2;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "cuTensorMapEncodeTiled failed with CUresult %d (rows %d K %d ld %d kmajor %d)", (int)r, rows, K,
             ld, (int)kmajor);
    return MT_ERR_CUDA;
  }
  return MT_OK;
}

int num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

unsigned long long* g_trace = nullptr;
int g_tc_mode = 2;       // 0 = 256-wide tiles where they apply, 1 = one CTA per SM everywhere, 2 = never use the 256-wide tile (default: measured fastest)

template <int BN, uint32_t F, bool RES, int OCC = 1, int MT = 1>
int launch_inst(const CUtensorMap& ma, const TcMapsB& mb, const TcArgs& g, int grid, cudaStream_t st) {
  static MtPerDeviceOnce once;
  static_assert(Smem<BN, OCC, MT>::TOTAL <= 227 * 1024, "shared memory budget");
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, F, RES, OCC, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN, OCC, MT>::TOTAL));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Smem<BN, OCC, MT>::THREADS); cfg.dynamicSmemBytes = Smem<BN, OCC, MT>::TOTAL; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = mt_pdl_enabled(st, MT_PDL_GEMM_TC);
  cfg.attrs = at; cfg.numAttrs = 1;
  MT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, F, RES, OCC, MT>, ma, mb, g));
  MT_LAUNCH_CHECK();
  return MT_OK;
}

// the feature set a descriptor needs
template <int BN, int MT = 1>
uint32_t needed_features(const GemmDesc& d, const TcArgs& g) {
  uint32_t f = 0;
  if (g.atomic) f |= F_ATOMIC;
  if (d.epi.bias) f |= F_BIAS;
  if (d.epi.act == MT_ACT_RELU) f |= F_RELU;
  if (d.epi.act == MT_ACT_TANH) f |= F_TANH;
  if (d.epi.drop.thresh != 0u) f |= F_DROP;
  if (d.epi.gate) f |= F_GATE;
  if (d.epi.residual) f |= F_RES;
  if (d.epi.rowmask) f |= F_ROWMASK;
  if (d.c_f32) f |= F_CF32;
  if (d.epi.colsum) f |= F_COLSUM;
  if (d.epi.alpha != 1.0f) f |= F_ALPHA;
  if (d.M % (MT * BM) != 0 || d.N % BN != 0 || d.ldc % 8 != 0 || (d.epi.gate && d.epi.ldg % 8 != 0)) f |= F_EDGE;
  return f;
}

template <int BN, int OCC, int MT = 1>
int launch_tc(const GemmDesc& d, cudaStream_t st) {
  TcArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.tiles_m = (d.M + MT * BM - 1) / (MT * BM);
  g.tiles_n = (d.N + BN - 1) / BN;
  g.kb_total = (d.K + BK - 1) / BK;
  int splits = 1;
  const int groups = d.groups > 1 ? d.groups : 1;
  if (d.split_k > 1) {                       // wgrad-style: few output tiles, long K -> one work item per SM
    splits = OCC * num_sms() / (g.tiles_m * g.tiles_n * groups);
    if (splits < 1) splits = 1;
  }
  if (splits > g.kb_total) splits = g.kb_total;
  g.kb_per = (g.kb_total + splits - 1) / splits;
  g.splits = (g.kb_total + g.kb_per - 1) / g.kb_per;
  g.a_mn = d.a_kmajor ? 0 : 1;
  g.b_mn = d.b_kmajor ? 0 : 1;
  g.C = d.C; g.ldc = d.ldc; g.c_f32 = d.c_f32 ? 1 : 0;
  g.atomic = d.split_k > 1 ? 1 : 0;
  g.epi = d.epi;
  g.gate_bf16 = 1;
  g.groups = groups; g.kb_group = groups > 1 ? d.K / BK : 0; g.c_gstride = d.c_gstride;
  g.trace = g_trace;
  const int mgroups = d.mgroups > 1 ? d.mgroups : 1;
  if (mgroups > 1 && (mgroups > 4 || d.M % mgroups != 0 || (d.M / mgroups) % (MT * BM) != 0)) return MT_ERR_UNSUPPORTED;
  g.mgroups = mgroups; g.tiles_m_group = g.tiles_m / mgroups;
  CUtensorMap ma;
  TcMapsB mb;
  memset(&mb, 0, sizeof(mb));
  MT_TRY(make_map(&ma, d.A, d.M, d.K * groups, d.lda, d.a_kmajor, BM));
  for (int i = 0; i < mgroups; ++i)
    MT_TRY(make_map(&mb.m[i], (const bf16*)d.B + (size_t)i * (size_t)d.b_gstride, d.N, d.K * groups, d.ldb, d.b_kmajor, BN));
  const int n_work = g.tiles_m * g.tiles_n * g.splits * groups;
  // g_tc_share > 1: launch only 1/share of the resident CTA slots so that GEMMs of concurrent streams (the three modality stacks)
  // co-reside on every SM instead of queueing behind each other (mt_tune)
  int slots = OCC * num_sms();
  if (g_mt_tune[MT_TUNE_GEMM_SHARE] > 1 && d.split_k <= 1) slots = slots / g_mt_tune[MT_TUNE_GEMM_SHARE];
  const int grid = n_work < slots ? n_work : slots;
  const uint32_t f = needed_features<BN, MT>(d, g);
  // weight-resident mode: short K, no split, and at least one CTA per column slice
  const bool res = OCC == 1 && g.splits == 1 && g.kb_total <= KB_RES && grid >= g.tiles_n && mgroups == 1;
  if constexpr (MT == 2) {      // 256-row tiles: the L2-bound shapes only (see mt_gemm_tc_run)
    switch (f) {
      case F_ATOMIC | F_CF32: return launch_inst<BN, F_ATOMIC | F_CF32, false, 1, 2>(ma, mb, g, grid, st);
      case F_ATOMIC | F_CF32 | F_EDGE: return launch_inst<BN, F_ATOMIC | F_CF32 | F_EDGE, false, 1, 2>(ma, mb, g, grid, st);
      case 0u: return launch_inst<BN, 0u, false, 1, 2>(ma, mb, g, grid, st);
      case F_EDGE: return launch_inst<BN, F_EDGE, false, 1, 2>(ma, mb, g, grid, st);
      default: return MT_ERR_UNSUPPORTED;
    }
  } else if constexpr (BN >= 128) {
    // instantiations of the encoder / MFN hot path (see mt_encoder.cu): exact feature-set matches only
#define MT_INST(FEAT)                                                                       \
  case (FEAT):                                                                              \
    if constexpr (OCC == 2) return launch_inst<BN, (FEAT), false, 2>(ma, mb, g, grid, st);  \
    else return res ? launch_inst<BN, (FEAT), true>(ma, mb, g, grid, st) : launch_inst<BN, (FEAT), false>(ma, mb, g, grid, st)
    switch (f) {
      MT_INST(F_BIAS);                                   // QKV projection
      MT_INST(F_BIAS | F_DROP | F_RES | F_CF32);         // out-proj / FFN2, train
      MT_INST(F_BIAS | F_RES | F_CF32);                  // out-proj / FFN2, eval
      MT_INST(F_BIAS | F_RELU | F_DROP);                 // FFN1, train
      MT_INST(F_BIAS | F_RELU);                          // FFN1, eval
      MT_INST(0u);                                       // dgrads
      MT_INST(F_GATE | F_COLSUM);                        // dgrad through relu + dropout
      case F_ATOMIC | F_CF32: return launch_inst<BN, F_ATOMIC | F_CF32, false, OCC>(ma, mb, g, grid, st);                    // wgrads (full tiles)
      case F_ATOMIC | F_CF32 | F_EDGE: return launch_inst<BN, F_ATOMIC | F_CF32 | F_EDGE, false, OCC>(ma, mb, g, grid, st);  // wgrads (ragged tiles)
      default: break;
    }
#undef MT_INST
  }
  if constexpr (MT == 2) return MT_ERR_UNSUPPORTED;
  else if constexpr (OCC == 2) return launch_inst<BN, F_GENERIC, false, 2>(ma, mb, g, grid, st);
  else return res ? launch_inst<BN, F_GENERIC, true>(ma, mb, g, grid, st) : launch_inst<BN, F_GENERIC, false>(ma, mb, g, grid, st);
}

}  // namespace

bool mt_gemm_tc_supported(const GemmDesc& d) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || !d.A || !d.B || !d.C) return false;
  if (d.N % 4 != 0 || d.ldc % 4 != 0) return false;
  if (d.lda % 8 != 0 || d.ldb % 8 != 0) return false;
  if (((uintptr_t)d.A & 15) || ((uintptr_t)d.B & 15) || ((uintptr_t)d.C & 15)) return false;
  if (d.epi.accumulate) return false;
  if (d.groups > 1 && (d.split_k <= 1 || d.a_kmajor || d.b_kmajor || d.K % BK != 0)) return false;      // grouping: split-K wgrad form only
  if (d.mgroups > 1 && (d.groups > 1 || d.split_k > 1 || d.mgroups > 4 || d.M % d.mgroups != 0 || (d.b_gstride & 7) || d.epi.bias || d.epi.colsum ||
                        d.epi.drop.thresh != 0u))
    return false;                                         // row groups: one weight matrix per group, nothing else per group
  if (d.split_k > 1 && !d.c_f32) return false;
  if (d.epi.colsum && (d.split_k > 1 || d.N > 1024)) return false;
  if (d.epi.bias && ((uintptr_t)d.epi.bias & 15)) return false;
  if (d.epi.residual && (((uintptr_t)d.epi.residual & 15) || d.epi.ldr % 4 != 0)) return false;
  if (d.epi.gate && (((uintptr_t)d.epi.gate & 7) || d.epi.ldg % 4 != 0)) return false;
  // MN-major operands are fetched in 64-wide boxes along their contiguous dimension: the global extent must
  // allow 16-byte aligned rows (ld % 8 above) -- nothing else; K-major likewise.
  return true;
}

void mt_gemm_tc_set_trace(unsigned long long* p) { g_trace = p; }

int mt_gemm_tc_run(const GemmDesc& d, cudaStream_t st) {
  if (!mt_gemm_tc_supported(d)) return MT_ERR_UNSUPPORTED;
  // 256-wide tiles (tcgen05.mma N = 256: half as many MMA / TMA issues per output, the single issuing threads are the
  // bottleneck of short-K GEMMs) whenever the weight-resident mode applies
  const int kb_total = (d.K + BK - 1) / BK;
  (void)kb_total;
  const bool wide_ok = d.N >= 256 && d.split_k <= 1 && kb_total <= KB_RES;
  if (g_tc_mode == 1) {                                 // one CTA per SM everywhere
    if (wide_ok) return launch_tc<256, 1>(d, st);
    if (d.N > 64) return launch_tc<128, 1>(d, st);
    return launch_tc<64, 1>(d, st);
  }
  // mode 0: 256-wide weight-resident tiles (one CTA per SM, 16 epilogue warps) where they apply -- measured 5 % SLOWER on the MFT
  // step than mode 2 (8.84 vs 9.26 ms: 768 wide tiles quantise badly over 148 SMs and a lone CTA cannot hide its own fill
  // latency), so the default is mode 2: split-K wgrads on fewer, longer splits, everything else two CTAs per SM with a
  // streaming ring
  if (g_tc_mode == 0 && wide_ok) return launch_tc<256, 1>(d, st);
  // L2-bound shapes: every 128 x 128 tile re-reads both operand blocks, and the SM <-> L2 fabric (about 1.8 x HBM) is what these kernels
  // saturate.  256-row tiles (two MMAs per k-step sharing B) and 256-wide tiles halve those re-reads: split-K weight gradients, and the
  // long-K input gradient of the QKV projection (K = 768: the weight matrix cannot stay resident, mt_gemm_rs.cu does not take it)
  if (!g_mt_tune[MT_TUNE_NO_BIG_TILES]) {
    const bool plain = !d.epi.bias && d.epi.act == MT_ACT_NONE && d.epi.drop.thresh == 0u && !d.epi.gate && !d.epi.residual && !d.epi.rowmask &&
                       !d.epi.colsum && d.epi.alpha == 1.0f && d.ldc % 8 == 0;
    // (measured: the big tile pays off for the [768, 256] weight gradient of the QKV projection only -- 58 -> 55 us for three stacks; the
    // smaller gradients are bound by the burst of fp32 atomics at the end of their few, long work items and lose with fewer, larger tiles)
    if (d.split_k > 1 && plain && d.N % 256 == 0 && (long long)d.M * d.N >= 768LL * 256 && d.M % 256 == 0) {
      const int rc = launch_tc<256, 1, 2>(d, st);
      if (rc != MT_ERR_UNSUPPORTED) return rc;
    }
    if (d.split_k <= 1 && plain && !d.c_f32 && d.N == 256 && d.K >= 512 && d.M >= 256 * 64 && d.a_kmajor) {
      const int rc = launch_tc<256, 1, 2>(d, st);
      if (rc != MT_ERR_UNSUPPORTED) return rc;
    }
  }
  if (d.split_k > 1 && d.N > 64) return launch_tc<128, 1>(d, st);
  if (d.N > 64) return launch_tc<128, 2>(d, st);
  return launch_tc<64, 1>(d, st);
}

int mt_gemm_tc_set_mode(int mode) { int old = g_tc_mode; g_tc_mode = mode; return old; }
