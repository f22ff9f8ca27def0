// Building blocks of the persistent recurrence kernels (MFN, LSTM decoder): a CTA owns BT narratives, keeps
// activations feature-major [feature][BT] in shared memory and applies small dense layers.
//
// Weight feed (STREAM = true): every in-loop weight matrix is consumed in a fixed order each step, so a dedicated
// producer warp streams them from L2 through a ring of shared-memory slots with 1-D bulk async copies
// (cp.async.bulk + mbarrier complete_tx), running ahead of the 16 consumer warps across layer and step boundaries:
// the dependent chain of a step never waits on an L2 round trip.  STREAM = false is the generic path (any
// alignment): weights are read straight from global memory with coalesced vector loads.
#pragma once
#include "mt_ops.cuh"

namespace mtrec {

constexpr int BT = 4;                 // narratives per CTA
constexpr int NTHREADS = 512;         // consumer threads (16 warps); a streaming kernel adds one producer warp
constexpr int VN = 4;                 // outputs per thread and per weight load
constexpr int PART_FLOATS = NTHREADS * VN * BT;   // K-split partial sums: P * N * BT <= NTHREADS * VN * BT
constexpr int CH_BYTES = 12288;       // bytes per ring slot
constexpr int NSLOT = 5;
constexpr int MAX_STREAM_LAYERS = 20;
constexpr uint32_t STREAM_SPIN_LIMIT = 1u << 27;

struct StreamLayer { const void* ptr; const void* ptr_special; int K, N; };    // ptr_special: used in the special step if non-null
struct StreamTable { StreamLayer L[MAX_STREAM_LAYERS]; int n; };

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = s_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > STREAM_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src),
               "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

// per-thread view of the weight ring (identical progression in the producer and in every consumer thread)
struct WRing {
  uint8_t* slots;
  uint64_t* full;
  uint64_t* empty;
  int slot;
  uint32_t phase;
  __device__ __forceinline__ void advance() { if (++slot == NSLOT) { slot = 0; phase ^= 1u; } }
};

template <typename WT>
__device__ __forceinline__ int rows_per_chunk(int N) { return CH_BYTES / (N * (int)sizeof(WT)); }

// ---- producer warp: one lane streams every layer of every step --------------------------------------------
template <typename WT>
__device__ __forceinline__ void stream_producer(const StreamTable& tab, int n_steps, int special_step, WRing r) {
  for (int step = 0; step < n_steps; ++step) {
    for (int l = 0; l < tab.n; ++l) {
      const StreamLayer& L = tab.L[l];
      const uint8_t* src = reinterpret_cast<const uint8_t*>((step == special_step && L.ptr_special) ? L.ptr_special : L.ptr);
      const int R = rows_per_chunk<WT>(L.N);
      const size_t row_bytes = (size_t)L.N * sizeof(WT);
      for (int k0 = 0; k0 < L.K; k0 += R) {
        const int rows = min(R, L.K - k0);
        sbar_wait(&r.empty[r.slot], r.phase ^ 1u);
        const uint32_t bytes = (uint32_t)(rows * row_bytes);
        sbar_expect_tx(&r.full[r.slot], bytes);
        bulk_g2s(r.slots + (size_t)r.slot * CH_BYTES, src + (size_t)k0 * row_bytes, bytes, &r.full[r.slot]);
        r.advance();
      }
    }
  }
}

__device__ __forceinline__ void load_w4(const float* p, float* w) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
__device__ __forceinline__ void load_w4(const bf16* p, float* w) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

// cross-slice reduction + epilogue shared by both feeds; ends with a CTA barrier
template <typename Epi>
__device__ __forceinline__ void dense_finish(float (*acc)[BT], int P, int p, int g, int N, float* part, Epi epi) {
  const int tid = threadIdx.x;
  if (P > 1) {
    if (p < P) {
#pragma unroll
      for (int i = 0; i < VN; ++i)
        *reinterpret_cast<float4*>(part + ((size_t)p * N + g * VN + i) * BT) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    cta_sync();
    for (int n = tid; n < N; n += NTHREADS) {
      float4 s = *reinterpret_cast<const float4*>(part + (size_t)n * BT);
      for (int q = 1; q < P; ++q) {
        const float4 o = *reinterpret_cast<const float4*>(part + ((size_t)q * N + n) * BT);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
      float a4[BT] = {s.x, s.y, s.z, s.w};
      epi(n, a4);
    }
  } else if (p < P) {
#pragma unroll
    for (int i = 0; i < VN; ++i) epi(g * VN + i, acc[i]);
  }
  cta_sync();
}

// ---- in-CTA dense layer ------------------------------------------------------------------------------
// out[n][b] = sum_k Wt[k*N + n] * xs[k][b],  n in [0,N), b in [0,BT).   Consecutive n are contiguous in Wt.
// Each thread owns VN consecutive outputs over one of P interleaved slices of the K range; the P partial sums
// meet in `part` (PART_FLOATS floats) in a fixed order (deterministic), and epi(n, acc[BT]) runs exactly once per n.
// All NTHREADS consumer threads must call this; on return every epi() write is visible to all of them.
template <bool STREAM, typename WT, typename Epi>
__device__ __forceinline__ void dense(WRing& ring, const WT* __restrict__ Wt, int K, int N, const float* __restrict__ xs, float* part,
                                      Epi epi) {
  const int tid = threadIdx.x;
  if (STREAM || ((N % VN == 0) && ((reinterpret_cast<uintptr_t>(Wt) & 15) == 0) && (N / VN <= NTHREADS))) {
    const int NG = N / VN;
    int P = NTHREADS / NG;
    if (P > 16) P = 16;
    if (P > K) P = K;
    const int p = tid / NG, g = tid - p * NG;
    float acc[VN][BT];
#pragma unroll
    for (int i = 0; i < VN; ++i)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[i][b] = 0.f;
    if (STREAM) {
      const int R = rows_per_chunk<WT>(N);
      for (int k0 = 0; k0 < K; k0 += R) {
        const int rows = min(R, K - k0);
        sbar_wait(&ring.full[ring.slot], ring.phase);
        if (p < P) {
          const WT* w = reinterpret_cast<const WT*>(ring.slots + (size_t)ring.slot * CH_BYTES) + g * VN;
#pragma unroll 4
          for (int r = p; r < rows; r += P) {
            float wv[VN];
            load_w4(w + (size_t)r * N, wv);
            const float4 x = *reinterpret_cast<const float4*>(xs + (k0 + r) * BT);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
              acc[i][0] = fmaf(wv[i], x.x, acc[i][0]); acc[i][1] = fmaf(wv[i], x.y, acc[i][1]);
              acc[i][2] = fmaf(wv[i], x.z, acc[i][2]); acc[i][3] = fmaf(wv[i], x.w, acc[i][3]);
            }
          }
        }
        __syncwarp();
        if ((tid & 31) == 0) sbar_arrive(&ring.empty[ring.slot]);      // one arrival per consumer warp
        ring.advance();
      }
    } else if (p < P) {
      const WT* w = Wt + g * VN;
#pragma unroll 4
      for (int k = p; k < K; k += P) {
        float wv[VN];
        load_w4(w + (size_t)k * N, wv);
        const float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          acc[i][0] = fmaf(wv[i], x.x, acc[i][0]); acc[i][1] = fmaf(wv[i], x.y, acc[i][1]);
          acc[i][2] = fmaf(wv[i], x.z, acc[i][2]); acc[i][3] = fmaf(wv[i], x.w, acc[i][3]);
        }
      }
    }
    dense_finish(acc, P, p, g, N, part, epi);
    return;
  }
  // generic path (any N / alignment): one output per thread, scalar weight loads
  for (int n = tid; n < N; n += NTHREADS) {
    float acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.f;
    const WT* w = Wt + n;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float wv = to_f(w[(size_t)k * N]);
      const float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
      acc[0] = fmaf(wv, x.x, acc[0]); acc[1] = fmaf(wv, x.y, acc[1]);
      acc[2] = fmaf(wv, x.z, acc[2]); acc[3] = fmaf(wv, x.w, acc[3]);
    }
    epi(n, acc);
  }
  cta_sync();
}

// can this table be streamed?  (16-byte aligned sources, rows that are multiples of 16 bytes, rows that fit a slot)
template <typename WT>
inline bool stream_table_ok(const StreamTable& t) {
  for (int i = 0; i < t.n; ++i) {
    const StreamLayer& L = t.L[i];
    const size_t rb = (size_t)L.N * sizeof(WT);
    if (L.N % VN != 0 || L.N / VN > NTHREADS || rb % 16 != 0 || rb > CH_BYTES) return false;
    if ((reinterpret_cast<uintptr_t>(L.ptr) & 15) || (reinterpret_cast<uintptr_t>(L.ptr_special) & 15)) return false;
  }
  return t.n > 0;
}

// shared-memory bytes of the ring + its barriers (placed at the END of the dynamic shared memory block, 128-aligned)
constexpr size_t RING_BYTES = (size_t)NSLOT * CH_BYTES + 128;

__device__ __forceinline__ WRing ring_setup(uint8_t* base /* 128-byte aligned */, bool init_thread) {
  WRing r;
  r.slots = base + 128;
  r.full = reinterpret_cast<uint64_t*>(base);
  r.empty = r.full + NSLOT;
  r.slot = 0; r.phase = 0;
  if (init_thread) {
    for (int s = 0; s < NSLOT; ++s) { sbar_init(&r.full[s], 1); sbar_init(&r.empty[s], NTHREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  return r;
}

// copy a [w][BT] shared buffer to BT global rows (row r of sample b at base + row_b*w)
__device__ __forceinline__ void stash_rows(float* __restrict__ g, int w, const float* __restrict__ s, const long long* rows, int nb) {
  if (!g) return;
  for (int e = threadIdx.x; e < nb * w; e += NTHREADS) {
    int b = e / w, f = e - b * w;
    g[rows[b] * w + f] = s[f * BT + b];
  }
}
__device__ __forceinline__ void load_rows(float* __restrict__ s, int w, const float* __restrict__ g, const long long* rows, int nb) {
  for (int e = threadIdx.x; e < BT * w; e += NTHREADS) {
    int b = e / w, f = e - b * w;
    s[f * BT + b] = b < nb ? g[rows[b] * w + f] : 0.f;
  }
}

// raise a kernel's dynamic shared-memory limit (only when it has to grow; keyed by the kernel's address)
template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return MT_ERR_UNSUPPORTED;
  static const void* fn[64];
  static size_t granted[64];
  static int n = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  int i = 0;
  while (i < n && fn[i] != key) ++i;
  if (i == n) {
    if (n == 64) { MT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); return MT_OK; }
    fn[n] = key; granted[n] = 0; ++n;
  }
  if (bytes > granted[i]) {
    MT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    granted[i] = bytes;
  }
  return MT_OK;
}

}  // namespace mtrec
