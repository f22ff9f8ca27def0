// tcgen05 GEMM engine (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   C[m,n] = epilogue( sum_k A(m,k) B(n,k) )      A, B each K-major or MN-major (see mt_gemm.cuh)
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0  : TMA producer   -- cp.async.bulk.tensor 2-D boxes (128-byte swizzle) into a 4-stage shared-memory ring
//   warp 1  : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16) from shared-memory
//                               descriptors; tcgen05.commit releases ring slots and publishes accumulators
//   warps 2-9: epilogue      -- each owns a TMEM lane quarter x a column half: tcgen05.ld the fp32 accumulator
//                               (double-buffered in TMEM so the next tile's MMAs overlap this tile's epilogue),
//                               release it, transpose through shared memory, then stream rows: the residual /
//                               gate loads of four row groups are issued back to back before any is consumed, and
//                               bias / activation / dropout / gate / residual / row-mask are applied on fully
//                               coalesced 128-256-byte row segments (fp32 atomics for split-K wgrad work items).
// Work items are (tile_m, tile_n, k_split) triples enumerated identically by all three roles.
// Out-of-bounds rows / columns / k are zero-filled by TMA and predicated in the epilogue.
#include <cuda.h>

#include "mt_gemm.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 4;
// epilogue warps per CTA: 4 per TMEM lane quarter for the 128-wide tile (the epilogue is issue-bound; more warps hide its
// latencies), 2 per quarter for the 64-wide one; plus the TMA warp and the MMA warp
template <int BN> struct EpiWarps { static constexpr int value = BN >= 128 ? 16 : 8; };
constexpr uint32_t SPIN_LIMIT = 1u << 27;

struct TcArgs {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_total, kb_per;
  int a_mn, b_mn;                      // operand major-ness (1 = MN-major)
  void* C; int ldc; int c_f32; int atomic;
  GemmEpi epi;
  int gate_bf16;
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

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EW = EpiWarps<BN>::value;
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int EPI_BYTES = EW * 32 * (BN / (EW / 4) + 4) * 4;      // per epilogue warp: 32 rows x (its columns + 4) floats
  static constexpr int BAR_BYTES = 256;
  static constexpr int CS_COLS = 1024;                                      // epi.colsum: per-CTA column accumulators (N <= CS_COLS)
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + CS_COLS * 4 + 1024 /* alignment slack */;
};

template <int BN>
__global__ void __launch_bounds__(Smem<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = Smem<BN>;
  uint8_t* stage_base = smem;
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::EPI_BYTES);
  uint64_t* full = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES; // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;     // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* cs_smem = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::EPI_BYTES + S::BAR_BYTES);
  if (g.epi.colsum) for (int i = threadIdx.x; i < g.N; i += blockDim.x) cs_smem[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = g.tiles_m * g.tiles_n * g.splits;
  const int tiles = g.tiles_m * g.tiles_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], S::EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int tile = w % tiles, split = w / tiles;
        const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN;
        const int kb0 = split * g.kb_per, kb1 = min(g.kb_total, kb0 + g.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_expect_tx(&full[stage], S::STAGE_BYTES);
          const int k0 = kb * BK;
          if (g.a_mn) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &map_a, m0 + 64 * j, k0, &full[stage]);
          } else {
            tma_load_2d(sa, &map_a, k0, m0, &full[stage]);
          }
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &map_b, n0 + 64 * j, k0, &full[stage]);
          } else {
            tma_load_2d(sb, &map_b, k0, n0, &full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int split = w / tiles;
        const int kb0 = split * g.kb_per, kb1 = min(g.kb_total, kb0 + g.kb_per);
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = g.a_mn ? make_desc(sa + k * 16 * 128, BK * 128, 1024) : make_desc(sa + k * 32, 16, 1024);
            const uint64_t bd = g.b_mn ? make_desc(sb + k * 16 * 128, BK * 128, 1024) : make_desc(sb + k * 32, 16, 1024);
            tc_mma(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);                 // ring slot is free once these MMAs have read it
          if (kb == kb1 - 1) tc_commit(&acc_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps: warp w owns TMEM lane quarter (w % 4) and column slice (w - 2) / 4 of every tile =====
    constexpr int HN = BN / (S::EW / 4);       // columns per epilogue warp
    constexpr int LDS = HN + 4;                // staged row stride in floats (16-byte aligned, conflict-free both ways)
    constexpr int LPR = HN / 4;                // lanes per output row (16-byte column quads)
    constexpr int RPI = 32 / LPR;              // rows covered by one warp-wide access
    constexpr int ITERS = 32 / RPI;
    constexpr int U = 4;                       // row-iterations whose global loads are issued back to back
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stg = epi_stage + (warp - 2) * 32 * LDS;
    const GemmEpi& e = g.epi;
    const DropCfg edrop = mt_drop_resolve(e.drop);
    const int lr = lane / LPR, lc = (lane % LPR) * 4;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int tile = w % tiles;
      const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN + half * HN;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int row_base = m0 + q * 32;
      __syncwarp();                             // previous tile's reads of `stg` are complete
#pragma unroll
      for (int c0 = 0; c0 < HN; c0 += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HN + c0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * LDS + c0 + j * 4) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      // all TMEM reads of this accumulator are done: hand it back to the MMA warp before touching global memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      const int n = n0 + lc;
      if (n >= g.N || row_base >= g.M) continue;
      if (g.atomic) {
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
          const int r = i * RPI + lr, m = row_base + r;
          if (m >= g.M) continue;
          const float4 a4 = *reinterpret_cast<const float4*>(stg + r * LDS + lc);
          float* cp = reinterpret_cast<float*>(g.C) + (size_t)m * g.ldc + n;
          atomicAdd(cp, a4.x * e.alpha); atomicAdd(cp + 1, a4.y * e.alpha); atomicAdd(cp + 2, a4.z * e.alpha); atomicAdd(cp + 3, a4.w * e.alpha);
        }
        continue;
      }
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias) bias4 = *reinterpret_cast<const float4*>(e.bias + n);
      float cs[4] = {0.f, 0.f, 0.f, 0.f};       // this lane's column partial sums (epi.colsum)
#pragma unroll 1
      for (int i0 = 0; i0 < ITERS; i0 += U) {
        float4 res[U], gt[U];
        float rm[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {            // issue every global load of this batch first
          const int m = row_base + (i0 + u) * RPI + lr;
          res[u] = make_float4(0.f, 0.f, 0.f, 0.f); gt[u] = make_float4(1.f, 1.f, 1.f, 1.f); rm[u] = 1.f;
          if (m < g.M) {
            if (e.residual) res[u] = *reinterpret_cast<const float4*>(e.residual + (size_t)m * e.ldr + n);
            if (e.gate) gt[u] = ld4(reinterpret_cast<const bf16*>(e.gate) + (size_t)m * e.ldg + n);
            if (e.rowmask) rm[u] = e.rowmask[m];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = (i0 + u) * RPI + lr, m = row_base + r;
          if (m >= g.M) continue;
          const float4 a4 = *reinterpret_cast<const float4*>(stg + r * LDS + lc);
          float o[4] = {a4.x * e.alpha + bias4.x, a4.y * e.alpha + bias4.y, a4.z * e.alpha + bias4.z, a4.w * e.alpha + bias4.w};
          if (e.act == MT_ACT_RELU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.f);
          } else if (e.act == MT_ACT_TANH) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = tanhf(o[k]);
          }
          if (edrop.thresh != 0u) {
            float f[4];        // N % 4 == 0 and n % 4 == 0: idx is a multiple of 4
            mt_drop_quad(edrop, (uint64_t)m * (uint64_t)g.N + (uint64_t)n, f);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] *= f[k];
          }
          if (e.gate) {
            o[0] = gt[u].x > 0.f ? o[0] * e.gate_scale : 0.f; o[1] = gt[u].y > 0.f ? o[1] * e.gate_scale : 0.f;
            o[2] = gt[u].z > 0.f ? o[2] * e.gate_scale : 0.f; o[3] = gt[u].w > 0.f ? o[3] * e.gate_scale : 0.f;
          }
          o[0] = (o[0] + res[u].x) * rm[u]; o[1] = (o[1] + res[u].y) * rm[u];
          o[2] = (o[2] + res[u].z) * rm[u]; o[3] = (o[3] + res[u].w) * rm[u];
          if (g.c_f32) st4(reinterpret_cast<float*>(g.C) + (size_t)m * g.ldc + n, make_float4(o[0], o[1], o[2], o[3]));
          else st4(reinterpret_cast<bf16*>(g.C) + (size_t)m * g.ldc + n, make_float4(o[0], o[1], o[2], o[3]));
          cs[0] += o[0]; cs[1] += o[1]; cs[2] += o[2]; cs[3] += o[3];
        }
      }
      if (e.colsum) {        // lanes lc, lc + LPR, ... hold the same columns (and took the same `continue` decisions above)
        const unsigned am = __activemask();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) cs[k] += __shfl_xor_sync(am, cs[k], o);
        }
        if (lr == 0) {       // shared-memory accumulation across this CTA's tiles and warps; one global atomic per column at exit
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(cs_smem + n + k, cs[k]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g.epi.colsum) for (int i = threadIdx.x; i < g.N; i += blockDim.x) atomicAdd(g.epi.colsum + i, cs_smem[i]);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
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
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch_tc(const GemmDesc& d, cudaStream_t st) {
  TcArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.tiles_m = (d.M + BM - 1) / BM;
  g.tiles_n = (d.N + BN - 1) / BN;
  g.kb_total = (d.K + BK - 1) / BK;
  int splits = 1;
  if (d.split_k > 1) {                       // wgrad-style: few output tiles, long K -> one work item per SM
    splits = num_sms() / (g.tiles_m * g.tiles_n);
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
  CUtensorMap ma, mb;
  MT_TRY(make_map(&ma, d.A, d.M, d.K, d.lda, d.a_kmajor, BM));
  MT_TRY(make_map(&mb, d.B, d.N, d.K, d.ldb, d.b_kmajor, BN));
  static bool attr_set = false;
  if (!attr_set) {
    MT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL));
    attr_set = true;
  }
  const int n_work = g.tiles_m * g.tiles_n * g.splits;
  const int grid = n_work < num_sms() ? n_work : num_sms();
  gemm_tc_kernel<BN><<<grid, Smem<BN>::THREADS, Smem<BN>::TOTAL, st>>>(ma, mb, g);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

}  // namespace

bool mt_gemm_tc_supported(const GemmDesc& d) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || !d.A || !d.B || !d.C) return false;
  if (d.N % 4 != 0 || d.ldc % 4 != 0) return false;
  if (d.lda % 8 != 0 || d.ldb % 8 != 0) return false;
  if (((uintptr_t)d.A & 15) || ((uintptr_t)d.B & 15) || ((uintptr_t)d.C & 15)) return false;
  if (d.epi.accumulate) return false;
  if (d.split_k > 1 && !d.c_f32) return false;
  if (d.epi.colsum && (d.split_k > 1 || d.N > 1024)) return false;
  if (d.epi.bias && ((uintptr_t)d.epi.bias & 15)) return false;
  if (d.epi.residual && (((uintptr_t)d.epi.residual & 15) || d.epi.ldr % 4 != 0)) return false;
  if (d.epi.gate && (((uintptr_t)d.epi.gate & 7) || d.epi.ldg % 4 != 0)) return false;
  // MN-major operands are fetched in 64-wide boxes along their contiguous dimension: the global extent must
  // allow 16-byte aligned rows (ld % 8 above) -- nothing else; K-major likewise.
  return true;
}

int mt_gemm_tc_run(const GemmDesc& d, cudaStream_t st) {
  if (!mt_gemm_tc_supported(d)) return MT_ERR_UNSUPPORTED;
  if (d.N > 64) return launch_tc<128>(d, st);
  return launch_tc<64>(d, st);
}
