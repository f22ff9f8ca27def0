// Tensor-core versions of the MFN recurrence kernels (bf16 mode), MFT/multiTransformer.py:200-235 and their reverse-time
// counterparts.  Same argument blocks and stash layouts as the FFMA kernels in mt_mfn.cu (which remain the fp32 / any-size
// path).
//
// A recurrence step is a chain of tiny dense layers on the state of a few narratives, so its cost is pure latency: the FFMA
// kernels spend ~5.8 us per step on K-split partial sums that meet in shared memory behind two CTA barriers per layer.  Here
//   * a CTA owns NB = 8 narratives -- exactly the n extent of mma.sync.m16n8k16 -- and 8 warps;
//   * every in-loop weight matrix is the A operand ([out, in] row-major, the PyTorch layout) and lives in REGISTERS as
//     pre-loaded fragments for the whole sequence (each warp owns 16-row slices of the outputs);
//   * the activations are the B operand: bf16 [narrative][feature] in shared memory, two 32-bit loads per k-step;
//   * the fp32 recurrent state (mem / c / dmem / dc) lives in the accumulator-fragment registers of the thread that owns
//     that (feature, narrative) pair, so the element-wise updates need no shared memory at all;
//   * per-step inputs are row segments of the stash: one warp issues one 1-D bulk async copy (cp.async.bulk + mbarrier
//     complete_tx) per (segment, narrative) FEED_DEPTH - 1 steps ahead into a shared-memory ring.  (Prefetching into registers
//     does not work: loads of different steps end up on the same hardware scoreboard, so the consumer of step t waits for
//     the loads of step t + 3 as well -- measured: the full DRAM latency was exposed in every step.)
//   * outputs leave with plain global stores.
// One layer = <= 8 dependent mma per warp + one CTA barrier.
#include "mt_mfn.cuh"
#include "mt_mma.cuh"
#include "mt_recurrent.cuh"
#include "mt_tcgen05.cuh"

// timing experiment (mt_tune key 8, bit 5): clock64 stamps of warp 0 / lane 0 of one CTA, 8 per step
__device__ unsigned long long g_rec_trace[4][128][8];
__device__ unsigned long long g_rec_trace2[8][128][8];      // [compute warp][step][stamp] of the second-cut LSTM forward

namespace {

using namespace mtmma;

__device__ __forceinline__ unsigned long long clk_after(float dep) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");
  return t;
}
#define REC_STAMP(KERN, STEP, SLOT, DEP)                                                                          \
  do {                                                                                                            \
    if (trace && (STEP) < 128) g_rec_trace[KERN][STEP][SLOT] = clk_after(DEP);                                      \
  } while (0)

constexpr int NB = 8;            // narratives per CTA
constexpr int NW = 8;            // warps per CTA
constexpr int NTH = NW * 32;
constexpr int LDK = 128 + 8;     // shared-memory row stride (bf16 elements) of a [NB][128] activation tile: conflict-free 32-bit reads

// A-operand fragment of a 16 x 16 block of a matrix given element-wise: a[0] = (r0, c..c+1), a[1] = (r1, ..), a[2] = (r0, c+8..), a[3] = (r1, c+8..)
template <typename Get>
__device__ __forceinline__ void frag_a(uint32_t* a, int row0, int k0, int lane, Get get) {
  const int r0 = row0 + (lane >> 2), r1 = r0 + 8, c = k0 + 2 * (lane & 3);
  a[0] = pack2(get(r0, c), get(r0, c + 1));
  a[1] = pack2(get(r1, c), get(r1, c + 1));
  a[2] = pack2(get(r0, c + 8), get(r0, c + 9));
  a[3] = pack2(get(r1, c + 8), get(r1, c + 9));
}
// B-operand fragment (k-step ks) of an activation tile S[NB][ld] (bf16, feature index = k)
__device__ __forceinline__ void frag_b(uint32_t& b0, uint32_t& b1, const bf16* S, int ld, int k0, int lane) {
  const bf16* p = S + (lane >> 2) * ld + k0 + 2 * (lane & 3);
  b0 = *reinterpret_cast<const uint32_t*>(p);
  b1 = *reinterpret_cast<const uint32_t*>(p + 8);
}
__device__ __forceinline__ float bf(const bf16* p) { return __bfloat162float(*p); }

// ---- per-step input feed: contiguous per-(narrative, step) spans of [M, ld] stash tensors -> shared-memory ring -----------
constexpr int FEED_MAXSEG = 6;
constexpr int FEED_PAD = 16;       // bytes between the narratives of a staged segment: spreads them over the banks
struct FeedSeg { const char* base; long long row_bytes; int bytes; int soff; };
struct Feed {
  FeedSeg s[FEED_MAXSEG];
  int n, stage_bytes, tx_bytes;
  __device__ void add(const void* base, long long row_bytes, int bytes) {
    s[n].base = reinterpret_cast<const char*>(base); s[n].row_bytes = row_bytes; s[n].bytes = bytes; s[n].soff = stage_bytes;
    stage_bytes += NB * (bytes + FEED_PAD); tx_bytes += NB * bytes; ++n;
  }
};
// called by ALL lanes of one warp, after a CTA barrier that follows the last read of `stage`
__device__ __forceinline__ void feed_issue(const Feed& f, char* stage, uint64_t* bar, int b0, int B, long long sb, long long st, int t, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (lane == 0) mtrec::sbar_expect_tx(bar, (uint32_t)f.tx_bytes);
  __syncwarp();
  for (int c = lane; c < f.n * NB; c += 32) {
    const int sg = c / NB, n = c - sg * NB;
    const FeedSeg& x = f.s[sg];
    const long long row = (long long)min(b0 + n, B - 1) * sb + (long long)t * st;
    mtrec::bulk_g2s(stage + x.soff + n * (x.bytes + FEED_PAD), x.base + row * x.row_bytes, (uint32_t)x.bytes, bar);
  }
}
// dedicated producer warp (warp NW of a NTH + 32 thread CTA): keeps the ring FD - 1 steps ahead of the consumers, throttled by
// the `empty` barriers (one arrival per consumer warp per step).  step_time(i) = time index of the i-th step.
template <typename TimeOf>
__device__ __forceinline__ void feed_producer(const Feed& f, char* stages, uint64_t* full, uint64_t* empty, int n_stage, int b0, int B,
                                              long long sb, long long st, int T, int lane, TimeOf time_of) {
  for (int i = 0; i < T; ++i) {
    const int slot = i % n_stage;
    if (i >= n_stage) mtrec::sbar_wait(&empty[slot], (uint32_t)(i / n_stage - 1) & 1u);
    feed_issue(f, stages + (size_t)slot * f.stage_bytes, &full[slot], b0, B, sb, st, time_of(i), lane);
  }
}
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); }
__device__ __forceinline__ void release_slot(uint64_t* empty, int lane) {
  __syncwarp();
  if (lane == 0) mtrec::sbar_arrive(empty);
}
// element f of narrative n of segment sg of a staged step
template <typename T>
__device__ __forceinline__ T feed_at(const Feed& f, const char* stage, int sg, int n, int e) {
  return *reinterpret_cast<const T*>(stage + f.s[sg].soff + n * (f.s[sg].bytes + FEED_PAD) + e * (int)sizeof(T));
}
// bf16-mode transcendental functions: hardware approximations (ex2 / tanh units, ~1e-3 relative), far inside the bf16 budget
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ======================================================================================================
// memory recurrence, forward:  gh = drop(relu(gpre_t + Wm mem_{t-1}));  g{1,2} = sigmoid(W2{1,2} gh{1,2} + b);
//                              mem_t = g1 mem_{t-1} + g2 cHat_t                       (MEM = 2G = 128)
// thread (warp w, gid = lane / 4, q = lane % 4) owns features f0 = 16 w + gid, f1 = f0 + 8 of narratives n0 = 2 q, n1 = n0 + 1
// in every layer: value index v = 2 * (row half) + (narrative parity), the mma accumulator order.
// ======================================================================================================
constexpr int FD = 4;            // stages of the input ring (FD - 1 steps of look-ahead)

template <bool TRAIN>
__global__ void __launch_bounds__(NTH + 32, 1) mem_fwd_mma_kernel(const __grid_constant__ MemArgs a) {
  __shared__ __align__(16) bf16 memS[NB * LDK];
  __shared__ __align__(16) bf16 ghS[NB * LDK];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ Feed feed;
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int MEM = 128, G = 64, G2 = 128, M2 = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const int f0 = warp * 16 + gid, f1 = f0 + 8;
  const int ldw1 = 2 * a.Hs + MEM;
  if (threadIdx.x == 0) {
    feed.n = 0; feed.stage_bytes = 0; feed.tx_bytes = 0;
    feed.add(a.gpre, (long long)G2 * sizeof(float), G2 * (int)sizeof(float));          // 0: gate pre-activations (batched part)
    feed.add(a.chat, (long long)MEM * sizeof(float), MEM * (int)sizeof(float));        // 1: cHat
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- weights -> A fragments ----
  uint32_t A1[8][4], A2a[4][4], A2b[4][4];
  float bias1[2] = {0.f, 0.f}, bias2[2] = {0.f, 0.f};
  if (warp < NW) {
    bias1[0] = a.g1_fc2_b[f0]; bias1[1] = a.g1_fc2_b[f1]; bias2[0] = a.g2_fc2_b[f0]; bias2[1] = a.g2_fc2_b[f1];
    const bf16* w1 = reinterpret_cast<const bf16*>(warp < 4 ? a.g1_fc1_w : a.g2_fc1_w) + 2 * a.Hs;      // mem columns of gamma{1,2}_fc1
    const int rb = (warp & 3) * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) frag_a(A1[ks], rb, ks * 16, lane, [&](int r, int c) { return bf(w1 + (size_t)r * ldw1 + c); });
    const bf16* v1 = reinterpret_cast<const bf16*>(a.g1_fc2_w);
    const bf16* v2 = reinterpret_cast<const bf16*>(a.g2_fc2_w);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      frag_a(A2a[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return bf(v1 + (size_t)r * G + c); });
      frag_a(A2b[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return bf(v2 + (size_t)r * G + c); });
    }
  }
  const DropCfg drop = mt_drop_resolve(warp < 4 ? a.drop_g1 : a.drop_g2);
  const int jg0 = f0 & 63, jg1 = f1 & 63;           // index inside the gate's own [T,B,G] dropout tensor
  // ---- narratives of this thread ----
  const int nn[2] = {2 * q, 2 * q + 1};
  bool val[2]; long long rbase[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) { val[p] = b0 + nn[p] < a.B; rbase[p] = (long long)min(b0 + nn[p], a.B - 1) * a.sb; }
  const int ff[2] = {f0, f1};
  for (int e = threadIdx.x; e < NB * LDK; e += NTH) { memS[e] = __float2bfloat16(0.f); ghS[e] = __float2bfloat16(0.f); }
  float mem[4] = {0.f, 0.f, 0.f, 0.f};
  bf16* gh_op = reinterpret_cast<bf16*>(a.gh_op);
  bf16* memprev_op = reinterpret_cast<bf16*>(a.memprev_op);
  bf16* last_op = reinterpret_cast<bf16*>(a.last_op);
  const int LW = a.Hs + MEM;
  __syncthreads();
  char* stages = reinterpret_cast<char*>(feed_smem);
  const int SBY = feed.stage_bytes;
  if (warp == NW) {
    feed_producer(feed, stages, full, empty, FD, b0, a.B, a.sb, a.st, a.T, lane, [](int i) { return i; });
    return;
  }

  {
    {
      for (int t = 0; t < a.T; ++t) {
      if (!(a.dbg & 4)) mtrec::sbar_wait(&full[t % FD], (uint32_t)(t / FD) & 1u);
      const char* sg = stages + (t % FD) * SBY;
      // ---- layer 1 ----
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t bb0, bb1;
        frag_b(bb0, bb1, memS, LDK, ks * 16, lane);
        mma16816(acc[ks & 1], A1[ks], bb0, bb1);
      }
      float gh[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float x = fmaxf(acc[0][v] + acc[1][v] + feed_at<float>(feed, sg, 0, nn[v & 1], ff[v >> 1]), 0.f);
        // element index of the gate's [T,B,G] tensor (oracle/mt_oracle.py:_drop_t)
        x *= mt_drop_factor(drop, ((uint64_t)t * a.B + (uint64_t)(b0 + nn[v & 1])) * (uint64_t)G + (uint64_t)((v >> 1) ? jg1 : jg0));
        gh[v] = x;
        ghS[nn[v & 1] * LDK + ff[v >> 1]] = __float2bfloat16(x);
      }
      consumers_sync();
      // ---- layer 2: gamma1 from hidden[0:64), gamma2 from hidden[64:128) ----
      float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t x0, x1, y0, y1;
        frag_b(x0, x1, ghS, LDK, ks * 16, lane);
        frag_b(y0, y1, ghS, LDK, G + ks * 16, lane);
        mma16816(c1, A2a[ks], x0, x1);
        mma16816(c2, A2b[ks], y0, y1);
      }
      float mp[4], g1[4], g2[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        g1[v] = fsig(c1[v] + bias1[v >> 1]);
        g2[v] = fsig(c2[v] + bias2[v >> 1]);
        mp[v] = mem[v];
        mem[v] = g1[v] * mp[v] + g2[v] * feed_at<float>(feed, sg, 1, nn[v & 1], ff[v >> 1]);
        memS[nn[v & 1] * LDK + ff[v >> 1]] = __float2bfloat16(mem[v]);
      }
      // ---- outputs of this step, then refill the prefetch slot ----
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (!val[v & 1] || (a.dbg & 2)) continue;
        const long long row = rbase[v & 1] + (long long)t * a.st;
        const int f = ff[v >> 1];
        if (TRAIN) {
          gh_op[row * G2 + f] = __float2bfloat16(gh[v]);
          a.gm[row * M2 + f] = g1[v];
          a.gm[row * M2 + MEM + f] = g2[v];
          memprev_op[row * MEM + f] = __float2bfloat16(mp[v]);
        }
        last_op[row * LW + a.Hs + f] = __float2bfloat16(mem[v]);
      }
      release_slot(&empty[t % FD], lane);
      consumers_sync();
      }
    }
  }
  if (a.mem_last) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (val[v & 1]) a.mem_last[(size_t)(b0 + nn[v & 1]) * MEM + ff[v >> 1]] = mem[v];
  }
}

// ======================================================================================================
// memory recurrence, backward (reverse time).  g = dmem + d(mem_t from the head);
//   dzg1 = g mem_{t-1} g1 (1 - g1);  dzg2 = g cHat g2 (1 - g2);  dzchat = g g2 (1 - cHat^2);  dmem = g g1
//   dgh = [gh > 0] sc * (W21^T dzg1 | W22^T dzg2);     dmem += Wm^T dgh
// ======================================================================================================

__global__ void __launch_bounds__(NTH + 32, 1) mem_bwd_mma_kernel(const __grid_constant__ MemArgs a) {
  __shared__ __align__(16) bf16 dz1S[NB * LDK];
  __shared__ __align__(16) bf16 dz2S[NB * LDK];
  __shared__ __align__(16) bf16 dghS[NB * LDK];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ Feed feed;
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int MEM = 128, G = 64, G2 = 128, M2 = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const int f0 = warp * 16 + gid, f1 = f0 + 8;
  const int ldw1 = 2 * a.Hs + MEM;
  if (threadIdx.x == 0) {
    feed.n = 0; feed.stage_bytes = 0; feed.tx_bytes = 0;
    feed.add(a.dlast + a.Hs, (long long)(a.Hs + MEM) * sizeof(float), MEM * (int)sizeof(float));      // 0: d mem_t from the head
    feed.add(a.gm, (long long)M2 * sizeof(float), M2 * (int)sizeof(float));                            // 1: gamma1 | gamma2
    feed.add(a.chat, (long long)MEM * sizeof(float), MEM * (int)sizeof(float));                        // 2: cHat
    feed.add(a.memprev_op, (long long)MEM * sizeof(bf16), MEM * (int)sizeof(bf16));                    // 3: mem_{t-1}
    feed.add(a.gh_op, (long long)G2 * sizeof(bf16), G2 * (int)sizeof(bf16));                           // 4: gamma hidden
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // A3: rows = hidden index n (warp's 16), k = mem feature: n < G: gamma1_fc2[k][n], else gamma2_fc2[k][n - G]
  // A4: rows = mem feature (warp's 16), k = hidden index: k < G: gamma1_fc1[k][2Hs + row], else gamma2_fc1[k - G][2Hs + row]
  uint32_t A3[8][4], A4[8][4];
  if (warp < NW) {
    const bf16* v = reinterpret_cast<const bf16*>(warp < 4 ? a.g1_fc2_w : a.g2_fc2_w);
    const int nb = (warp & 3) * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) frag_a(A3[ks], nb, ks * 16, lane, [&](int r, int c) { return bf(v + (size_t)c * G + r); });
    const bf16* w1 = reinterpret_cast<const bf16*>(a.g1_fc1_w) + 2 * a.Hs;
    const bf16* w2 = reinterpret_cast<const bf16*>(a.g2_fc1_w) + 2 * a.Hs;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      frag_a(A4[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return c < G ? bf(w1 + (size_t)c * ldw1 + r) : bf(w2 + (size_t)(c - G) * ldw1 + r); });
  }
  const int nn[2] = {2 * q, 2 * q + 1};
  bool val[2]; long long rbase[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) { val[p] = b0 + nn[p] < a.B; rbase[p] = (long long)min(b0 + nn[p], a.B - 1) * a.sb; }
  const int ff[2] = {f0, f1};
  const int LW = a.Hs + MEM;
  const bf16* memprev_op = reinterpret_cast<const bf16*>(a.memprev_op);
  const bf16* gh_op = reinterpret_cast<const bf16*>(a.gh_op);
  bf16* dzg_op = reinterpret_cast<bf16*>(a.dzg_op);
  bf16* dzchat_op = reinterpret_cast<bf16*>(a.dzchat_op);
  bf16* dgh_op = reinterpret_cast<bf16*>(a.dgh_op);
  const float sc_g = a.drop_g1.scale;
  float dmem[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  char* stages = reinterpret_cast<char*>(feed_smem);
  const int SBY = feed.stage_bytes;
  if (warp == NW) {
    const int T = a.T;
    feed_producer(feed, stages, full, empty, FD, b0, a.B, a.sb, a.st, a.T, lane, [T](int i) { return T - 1 - i; });
    return;
  }

  {
    {
      for (int i = 0; i < a.T; ++i) {
      const int t = a.T - 1 - i;
      if (!(a.dbg & 4)) mtrec::sbar_wait(&full[i % FD], (uint32_t)(i / FD) & 1u);
      const char* sg = stages + (i % FD) * SBY;
      struct { float dl[4], g1[4], g2[4], ch[4], mp[4], gh[4]; } x;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int n = nn[v & 1], f = ff[v >> 1];
        x.dl[v] = feed_at<float>(feed, sg, 0, n, f);
        x.g1[v] = feed_at<float>(feed, sg, 1, n, f);
        x.g2[v] = feed_at<float>(feed, sg, 1, n, MEM + f);
        x.ch[v] = feed_at<float>(feed, sg, 2, n, f);
        x.mp[v] = __bfloat162float(feed_at<bf16>(feed, sg, 3, n, f));
        x.gh[v] = __bfloat162float(feed_at<bf16>(feed, sg, 4, n, f));
      }
      release_slot(&empty[i % FD], lane);
      float ghv[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float g = dmem[v] + x.dl[v];
        const float d1 = g * x.mp[v] * x.g1[v] * (1.f - x.g1[v]);
        const float d2 = g * x.ch[v] * x.g2[v] * (1.f - x.g2[v]);
        const float dc = g * x.g2[v] * (1.f - x.ch[v] * x.ch[v]);
        dmem[v] = g * x.g1[v];
        ghv[v] = x.gh[v];
        const int so = nn[v & 1] * LDK + ff[v >> 1];
        dz1S[so] = __float2bfloat16(d1);
        dz2S[so] = __float2bfloat16(d2);
        if (val[v & 1] && !(a.dbg & 2)) {
          const long long row = rbase[v & 1] + (long long)t * a.st;
          dzg_op[row * M2 + ff[v >> 1]] = __float2bfloat16(d1);
          dzg_op[row * M2 + MEM + ff[v >> 1]] = __float2bfloat16(d2);
          dzchat_op[row * MEM + ff[v >> 1]] = __float2bfloat16(dc);
        }
      }
      consumers_sync();
      // ---- d hidden (hidden index = this thread's f0 / f1: warps 0-3 gate 1, warps 4-7 gate 2) ----
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const bf16* src = warp < 4 ? dz1S : dz2S;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t bb0, bb1;
        frag_b(bb0, bb1, src, LDK, ks * 16, lane);
        mma16816(acc[ks & 1], A3[ks], bb0, bb1);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float dg = ghv[v] > 0.f ? (acc[0][v] + acc[1][v]) * sc_g : 0.f;
        dghS[nn[v & 1] * LDK + ff[v >> 1]] = __float2bfloat16(dg);
        if (val[v & 1] && !(a.dbg & 2)) dgh_op[(rbase[v & 1] + (long long)t * a.st) * G2 + ff[v >> 1]] = __float2bfloat16(dg);
      }
      consumers_sync();
      // ---- d mem_{t-1} += gamma_fc1[:, 2H:]^T d hidden ----
      float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t bb0, bb1;
        frag_b(bb0, bb1, dghS, LDK, ks * 16, lane);
        mma16816(ac2[ks & 1], A4[ks], bb0, bb1);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) dmem[v] += ac2[0][v] + ac2[1][v];
      }
    }
  }
}

// ======================================================================================================
// LSTM recurrence, forward (LSTHM cell, MFT/multiTransformer.py:208): z = zx_t + b_hh + W_hh h_{t-1}, gates (i, f, g, o).
// One CTA per (8 narratives, modality).  The 4H gate rows are permuted so that one 16-row mma tile holds the four gates of
// four hidden units (tile row r: gate r / 4, unit 4 * tile + r % 4): the thread pair (lane, lane ^ 16) then holds i, g and
// f, o of the same unit, activates its own two gates, swaps them with one shuffle each, and each of the two finishes one
// narrative of that unit -- c_t and h_t never leave registers (h_t also goes to shared memory as the next step's B operand).
// H <= 96, H % 4 == 0 (the defaults are 48 / 88 / 88): tiles 4 * (warp + 8 i), i < 3; K padded to 96 with zero fragments.
// ======================================================================================================
constexpr int LMT = 3;            // 16-row tiles per warp
constexpr int LKS = 6;            // k-steps (K = 96)
constexpr int LDH = 96 + 8;       // h tile row stride: conflict-free fragment reads

template <bool TRAIN>
__global__ void __launch_bounds__(NTH + 32, 1) lstm_fwd_mma_kernel(const __grid_constant__ LstmArgs a) {
  __shared__ __align__(16) bf16 hS[2][NB * LDH];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ Feed feed;
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  if (threadIdx.x == 0) {
    feed.n = 0; feed.stage_bytes = 0; feed.tx_bytes = 0;
    feed.add(a.gates + 4 * hoff, (long long)4 * Hs * sizeof(float), 4 * H * (int)sizeof(float));      // 0: zx = x W_ih^T + b_ih, [gate][unit]
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool lower = gid < 4;                       // holds i (row gid) and g (row gid + 8); the upper half holds f and o
  const int gate0 = lower ? 0 : 1, gate1 = gate0 + 2;
  const bf16* W = reinterpret_cast<const bf16*>(a.w_hh[m]);
  uint32_t A[LMT][LKS][4];
  int unit[LMT];
  float bz[LMT][2];                                 // b_hh of this thread's two gate rows
#pragma unroll
  for (int i = 0; i < LMT; ++i) {
    const int mt = warp + NW * i;
    unit[i] = 4 * mt + (gid & 3);
    if (warp >= NW) continue;
    auto get = [&](int r, int c) {
      const int u = 4 * mt + (r & 3);
      return (u < H && c < H) ? bf(W + (size_t)((r >> 2) * H + u) * H + c) : 0.f;
    };
#pragma unroll
    for (int ks = 0; ks < LKS; ++ks) frag_a(A[i][ks], 0, ks * 16, lane, get);
    const bool on = unit[i] < H;
    bz[i][0] = on ? a.b_hh[m][gate0 * H + unit[i]] : 0.f;
    bz[i][1] = on ? a.b_hh[m][gate1 * H + unit[i]] : 0.f;
  }
  const int nn[2] = {2 * q, 2 * q + 1};
  bool val[2]; long long rbase[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) { val[p] = b0 + nn[p] < a.B; rbase[p] = (long long)min(b0 + nn[p], a.B - 1) * a.sb; }
  // the narrative (of its two) this thread finishes: n0 for the lower half, n1 for the upper half
  const int n_mine = lower ? nn[0] : nn[1];
  const bool val_mine = lower ? val[0] : val[1];
  const long long rb_mine = lower ? rbase[0] : rbase[1];
  for (int e = threadIdx.x; e < 2 * NB * LDH; e += NTH) hS[0][e] = __float2bfloat16(0.f);
  float c[LMT] = {0.f, 0.f, 0.f}, h[LMT] = {0.f, 0.f, 0.f};
  bf16* cstar_op = reinterpret_cast<bf16*>(a.cstar_op);
  bf16* last_op = reinterpret_cast<bf16*>(a.last_op);
  bf16* hprev_op = reinterpret_cast<bf16*>(a.hprev_op);
  const int LW = Hs + a.MEM;
  __syncthreads();
  char* stages = reinterpret_cast<char*>(feed_smem);
  const int SBY = feed.stage_bytes;
  if (warp == NW) {
    feed_producer(feed, stages, full, empty, FD, b0, a.B, a.sb, a.st, a.T, lane, [](int i) { return i; });
    return;
  }

  {
    {
      const bool trace = (a.dbg & 32) && blockIdx.x == 0 && m == a.n_mods - 1 && threadIdx.x == 0;
      for (int t = 0; t < a.T; ++t) {
      REC_STAMP(0, t, 0, 0.f);
      const bf16* hin = hS[t & 1];
      bf16* hout = hS[(t & 1) ^ 1];
      float acc[LMT][4];
#pragma unroll
      for (int i = 0; i < LMT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < LKS; ++ks) {
        if (ks * 16 >= H || (a.dbg & 8)) continue;
        uint32_t bb0, bb1;
        frag_b(bb0, bb1, hin, LDH, ks * 16, lane);
#pragma unroll
        for (int i = 0; i < LMT; ++i) mma16816(acc[i], A[i][ks], bb0, bb1);
      }
      REC_STAMP(0, t, 1, 0.f);
      if (!(a.dbg & 4)) mtrec::sbar_wait(&full[t % FD], (uint32_t)(t / FD) & 1u);
      REC_STAMP(0, t, 2, 0.f);
      REC_STAMP(0, t, 3, acc[0][0] + acc[1][0] + acc[2][0]);
      const char* sg = stages + (t % FD) * SBY;
#pragma unroll
      for (int i = 0; i < LMT; ++i) {
        if (4 * (warp + NW * i) >= H) continue;                  // warp-uniform: this tile does not exist for this modality
        float g[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float z = acc[i][v] + feed_at<float>(feed, sg, 0, nn[v & 1], ((v >> 1) ? gate1 : gate0) * H + min(unit[i], H - 1)) + bz[i][v >> 1];
          // lower half: row gid = i (sigmoid), row gid + 8 = g (tanh); upper half: f and o (both sigmoid)
          g[v] = (lower && (v >> 1)) ? ftanh(z) : fsig(z);
        }
        if (TRAIN && unit[i] < H && !(a.dbg & 1)) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (val[v & 1])
              a.gates[(rbase[v & 1] + (long long)t * a.st) * (4 * Hs) + 4 * hoff + ((v >> 1) ? gate1 : gate0) * H + unit[i]] = g[v];
        }
        // swap: the lower thread sends its narrative-n1 pair (i, g), the upper thread its narrative-n0 pair (f, o)
        const float s0 = lower ? g[1] : g[0], s1 = lower ? g[3] : g[2];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
        const float gi = lower ? g[0] : r0, gg = lower ? g[2] : r1, gf = lower ? r0 : g[1], go = lower ? r1 : g[3];
        const float cp = c[i], hp = h[i];
        const float cn = gf * cp + gi * gg;
        const float hn = go * ftanh(cn);
        if (i == 0) REC_STAMP(0, t, 4, hn);
        if (i == LMT - 1) REC_STAMP(0, t, 5, hn);
        c[i] = cn; h[i] = hn;
        if (unit[i] < H) {
          hout[n_mine * LDH + unit[i]] = __float2bfloat16(hn);
          if (val_mine && !(a.dbg & 2)) {
            const long long row = rb_mine + (long long)t * a.st;
            const int col = hoff + unit[i];
            a.cstar[row * (2 * Hs) + col] = cp;
            a.cstar[row * (2 * Hs) + Hs + col] = cn;
            if (cstar_op) { cstar_op[row * (2 * Hs) + col] = __float2bfloat16(cp); cstar_op[row * (2 * Hs) + Hs + col] = __float2bfloat16(cn); }
            last_op[row * LW + col] = __float2bfloat16(hn);
            if (TRAIN) hprev_op[row * Hs + col] = __float2bfloat16(hp);
          }
        }
      }
      REC_STAMP(0, t, 6, 0.f);
      release_slot(&empty[t % FD], lane);
      REC_STAMP(0, t, 7, 0.f);
      consumers_sync();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LMT; ++i) {
    if (unit[i] < H && val_mine) {
      if (a.h_last) a.h_last[(size_t)(b0 + n_mine) * Hs + hoff + unit[i]] = h[i];
      if (a.c_last) a.c_last[(size_t)(b0 + n_mine) * Hs + hoff + unit[i]] = c[i];
    }
  }
}

// ======================================================================================================
// LSTM recurrence, forward, second cut.  What the clock trace of the kernel above showed (profiles/r02_mfn_rec_trace.txt): of
// 5 400 clocks per step 2 500 were the compute warps' own scattered 4-byte global stores (30 per thread and step), 1 900 the three
// tiles' element-wise code running one after the other (address arithmetic through a shared-memory descriptor, 64-bit index
// math per store), 400 the 18 mma of a warp (mma.sync issues at ~11 clocks per SMSP: a throughput floor, not a latency chain).
// Here
//   * the compute warps never touch global memory: gate activations and cell states go to a double-buffered staging tile in
//     shared memory, h_t into a 3-deep ring that is also the next step's B operand; FOUR STORER WARPS copy step t out while step
//     t + 1 computes -- every task is one (tensor, narrative, <= 32 chunks of 16 bytes) span from a table built once per CTA, so a
//     step leaves as ~72 fully coalesced 16-byte-per-lane stores instead of 7 680 scalar ones; bf16 operand copies of the cell
//     state are converted by the storers;
//   * zx_t + b_hh is the INITIAL VALUE of the accumulators (loaded before the step's barrier), the k range is split over two
//     accumulators per tile (chains of 3), the feed barrier of step t + 1 is waited for at the end of step t;
//   * all shared-memory offsets are per-thread constants in registers.
// Hand-off: named barriers 2 / 3 (compute -> storers: slot staged; bar.arrive by the compute warps, bar.sync by the storers) and
// 4 / 5 (storers -> compute: slot drained, two steps later).
// ======================================================================================================
#define REC_STAMP2(STEP, SLOT, DEP)                                                                             \
  do {                                                                                                            \
    if (trace && (STEP) < 128) g_rec_trace2[threadIdx.x >> 5][STEP][SLOT] = clk_after(DEP);                         \
  } while (0)
constexpr int NSW = 3;                              // storer warps
constexpr int NST = NSW * 32;
constexpr int NTH2 = (NW + 1 + NSW) * 32;           // compute | producer | storers
constexpr int NSYNC2 = (NW + NSW) * 32;             // participants of the staged / drained barriers
constexpr int HRING = 3;

__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float lds_f32(const char* p) { return *reinterpret_cast<const float*>(p); }
__device__ __forceinline__ void stg16(unsigned long long gaddr, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// sigmoid through the tanh unit: one MUFU instead of ex2 + rcp, no divergent branch between the gate kinds (abs. error ~3e-4)
__device__ __forceinline__ float act_tanh(float z, float pre, float post_mul, float post_add) { return fmaf(ftanh(z * pre), post_mul, post_add); }

// One storer thread's fixed list of 16-byte chunks of one output tensor: chunk c of narrative n lives at staging offset
// n * sstride + c * cbytes (cbytes = 16, or 32 where the storer converts fp32 -> bf16) and goes to dst0 + (row_n + t * st) * pitch + 16 c.
template <int MAXJ>
struct StoreJobs {
  uint32_t src[MAXJ];
  long long dst[MAXJ];       // byte offset from the tensor base at t = 0; < 0: no job
  unsigned long long base;
  long long inc;
  __device__ __forceinline__ void build(void* tensor, long long pitch, long long col, int chunks, int sstride, int soff, int cbytes, int b0, int B,
                                        long long sb, long long st, int tid, int nst = NST) {
    base = reinterpret_cast<unsigned long long>(tensor);
    inc = st * pitch;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int idx = tid + j * nst;
      const int n = idx / chunks, c = idx - n * chunks;
      const bool on = tensor != nullptr && n < NB && b0 + n < B;
      src[j] = (uint32_t)(soff + n * sstride + c * cbytes);
      dst[j] = on ? (long long)(b0 + n) * sb * pitch + col + 16ll * c : -1ll;
    }
  }
};

template <bool TRAIN, int NT, int KS>
__device__ __forceinline__ void lstm_fwd2_compute(const LstmArgs& a, int m, bf16 (*hS)[NB * LDH], const char* stages, char* stag, int SBY, int FST,
                                                  int GST, int CST, int SLOT, uint64_t* full, uint64_t* empty, bool trace) {
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const bool lower = gid < 4;                        // holds i (row gid) and g (row gid + 8); the upper half holds f and o
  const int gate0 = lower ? 0 : 1, gate1 = gate0 + 2;
  const bf16* W = reinterpret_cast<const bf16*>(a.w_hh[m]);
  uint32_t A[NT][KS][4];
  int zo[NT], so_g[NT], so_c[NT], so_h[NT];          // byte offsets: feed (n0, gate0), staging gates / c, h ring element
  float bz[NT][2];
  const int nn0 = 2 * q, n_mine = lower ? nn0 : nn0 + 1;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int mt = warp + NW * i;
    const int unit = 4 * mt + (gid & 3);
    auto get = [&](int r, int c) {
      const int u = 4 * mt + (r & 3);
      return (u < H && c < H) ? bf(W + (size_t)((r >> 2) * H + u) * H + c) : 0.f;
    };
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) frag_a(A[i][ks], 0, ks * 16, lane, get);
    bz[i][0] = a.b_hh[m][gate0 * H + unit];
    bz[i][1] = a.b_hh[m][gate1 * H + unit];
    zo[i] = nn0 * FST + (gate0 * H + unit) * 4;
    so_g[i] = n_mine * GST + unit * 4;
    so_c[i] = NB * GST + n_mine * CST + unit * 4;
    so_h[i] = (n_mine * LDH + unit) * 2;
  }
  // activation constants of this thread's two accumulator rows: sigmoid(z) = 0.5 + 0.5 tanh(z / 2); row 1 of the lower half is the g gate
  const float pre1 = lower ? 1.f : 0.5f, mul1 = lower ? 1.f : 0.5f, add1 = lower ? 0.f : 0.5f;
  float c[NT], h[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) c[i] = h[i] = 0.f;
  // ldmatrix.x4 source row of this lane: matrix j = lane / 8 covers k = 8 j .. 8 j + 7 of a 32-wide k block, row = narrative lane % 8
  const uint32_t hs0 = mtrec::s_u32(&hS[0][0]) + (uint32_t)(((lane & 7) * LDH + (lane >> 3) * 8) * 2);
  int r0 = 0;
  mtrec::sbar_wait(&full[0], 0u);
  for (int t = 0; t < a.T; ++t) {
    REC_STAMP2(t, 0, 0.f);
    const int slot = t & 1, r1 = r0 == HRING - 1 ? 0 : r0 + 1;
    const char* sg = stages + (t % FD) * SBY;
    float accA[NT][4], accB[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        accA[i][v] = lds_f32(sg + zo[i] + (v & 1) * FST + (v >> 1) * 8 * H) + bz[i][v >> 1];
        accB[i][v] = 0.f;
      }
    }
    consumers_sync();                                // h_{t-1} of every warp is in ring slot r0
    const uint32_t hin = hs0 + (uint32_t)(r0 * NB * LDH * 2);
#pragma unroll
    for (int kp = 0; kp < (KS + 1) / 2; ++kp) {
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(hin + kp * 64));
#pragma unroll
      for (int i = 0; i < NT; ++i) mma16816(accA[i], A[i][2 * kp], b[0], b[1]);
      if (2 * kp + 1 < KS) {
#pragma unroll
        for (int i = 0; i < NT; ++i) mma16816(accB[i], A[i][2 * kp + 1], b[2], b[3]);
      }
    }
    REC_STAMP2(t, 1, 0.f);
    release_slot(&empty[t % FD], lane);
    if (t + 1 < a.T) mtrec::sbar_wait(&full[(t + 1) % FD], (uint32_t)((t + 1) / FD) & 1u);      // long complete: overlaps the mma
    if (t >= 2) nb_sync(4 + slot, NSYNC2);           // the storers are done with this staging slot and with ring slot r1
    REC_STAMP2(t, 2, accA[0][0] + accB[0][0]);
    char* sbase = stag + slot * SLOT;
    char* hout = reinterpret_cast<char*>(hS[r1]);
    float g[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      g[i][0] = act_tanh(accA[i][0] + accB[i][0], 0.5f, 0.5f, 0.5f);
      g[i][1] = act_tanh(accA[i][1] + accB[i][1], 0.5f, 0.5f, 0.5f);
      g[i][2] = act_tanh(accA[i][2] + accB[i][2], pre1, mul1, add1);
      g[i][3] = act_tanh(accA[i][3] + accB[i][3], pre1, mul1, add1);
    }
    REC_STAMP2(t, 3, g[0][0] + g[NT - 1][3]);
    float gi[NT], gf[NT], gg[NT], go[NT], cp[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      // swap: the lower thread sends its narrative-n1 pair (i, g), the upper thread its narrative-n0 pair (f, o)
      const float s0 = lower ? g[i][1] : g[i][0], s1 = lower ? g[i][3] : g[i][2];
      const float x0 = __shfl_xor_sync(0xffffffffu, s0, 16), x1 = __shfl_xor_sync(0xffffffffu, s1, 16);
      gi[i] = lower ? g[i][0] : x0; gg[i] = lower ? g[i][2] : x1; gf[i] = lower ? x0 : g[i][1]; go[i] = lower ? x1 : g[i][3];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      cp[i] = c[i];
      c[i] = gf[i] * cp[i] + gi[i] * gg[i];
      h[i] = go[i] * ftanh(c[i]);
    }
    REC_STAMP2(t, 4, h[0]);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      *reinterpret_cast<bf16*>(hout + so_h[i]) = __float2bfloat16(h[i]);
      if (!(a.dbg & 1)) {
        float* sc = reinterpret_cast<float*>(sbase + so_c[i]);
        sc[0] = cp[i]; sc[H] = c[i];
        if (TRAIN) {
          float* sgt = reinterpret_cast<float*>(sbase + so_g[i]);
          sgt[0] = gi[i]; sgt[H] = gf[i]; sgt[2 * H] = gg[i]; sgt[3 * H] = go[i];
        }
      }
    }
    REC_STAMP2(t, 5, 0.f);
    nb_arrive(2 + slot, NSYNC2);                     // staged: the storers take step t from here
    REC_STAMP2(t, 6, 0.f);
    r0 = r1;
  }
  const bool val_mine = b0 + n_mine < a.B;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int unit = 4 * (warp + NW * i) + (gid & 3);
    if (val_mine) {
      if (a.h_last) a.h_last[(size_t)(b0 + n_mine) * Hs + hoff + unit] = h[i];
      if (a.c_last) a.c_last[(size_t)(b0 + n_mine) * Hs + hoff + unit] = c[i];
    }
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(NTH2, 1) lstm_fwd_mma2_kernel(const __grid_constant__ LstmArgs a) {
  __shared__ __align__(16) bf16 hS[HRING][NB * LDH];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ Feed feed;
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * NB;
  const int LW = Hs + a.MEM;
  const int FST = 16 * H + FEED_PAD;                 // feed: bytes between narratives of a staged step ([gate][unit] fp32 rows)
  const int GST = TRAIN ? 16 * H + 16 : 0;           // staging: gate activations [gate][unit] fp32 per narrative
  const int CST = 8 * H + 16;                        //          c_{t-1} | c_t fp32 per narrative
  const int SLOT = NB * (GST + CST);
  char* stages = reinterpret_cast<char*>(feed_smem);
  char* stag = stages + (size_t)FD * NB * FST;
  if (threadIdx.x == 0) {
    feed.n = 0; feed.stage_bytes = 0; feed.tx_bytes = 0;
    feed.add(a.gates + 4 * hoff, (long long)4 * Hs * sizeof(float), 4 * H * (int)sizeof(float));      // 0: zx = x W_ih^T + b_ih, [gate][unit]
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = threadIdx.x; e < HRING * NB * LDH; e += NTH2) hS[0][e] = __float2bfloat16(0.f);
  __syncthreads();
  const int SBY = feed.stage_bytes;
  if (warp == NW) {
    feed_producer(feed, stages, full, empty, FD, b0, a.B, a.sb, a.st, a.T, lane, [](int i) { return i; });
    return;
  }
  if (warp > NW) {                                   // ---- storers ----
    const int tid = threadIdx.x - (NW + 1) * 32;
    // per-thread job lists (H <= 96: at most 8 * 96 / 96 gate chunks, 8 * 24 / 96 state chunks, ... per thread)
    StoreJobs<8> jg; StoreJobs<2> jc0, jc1; StoreJobs<1> jv0, jv1, jh, jp;
    jg.build(TRAIN ? a.gates : nullptr, 16ll * Hs, 16ll * hoff, H, GST, 0, 16, b0, a.B, a.sb, a.st, tid);
    jc0.build(a.cstar, 8ll * Hs, 4ll * hoff, H / 4, CST, NB * GST, 16, b0, a.B, a.sb, a.st, tid);
    jc1.build(a.cstar, 8ll * Hs, 4ll * (Hs + hoff), H / 4, CST, NB * GST + 4 * H, 16, b0, a.B, a.sb, a.st, tid);
    jv0.build(a.cstar_op, 4ll * Hs, 2ll * hoff, H / 8, CST, NB * GST, 32, b0, a.B, a.sb, a.st, tid);
    jv1.build(a.cstar_op, 4ll * Hs, 2ll * (Hs + hoff), H / 8, CST, NB * GST + 4 * H, 32, b0, a.B, a.sb, a.st, tid);
    jh.build(a.last_op, 2ll * LW, 2ll * hoff, H / 8, LDH * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    jp.build(TRAIN ? a.hprev_op : nullptr, 2ll * Hs, 2ll * hoff, H / 8, LDH * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    const uint32_t stag_s = mtrec::s_u32(stag), hs_s = mtrec::s_u32(&hS[0][0]);
    int r0 = 0;                                      // ring slot of h_{t-1}
    for (int t = 0; t < a.T; ++t) {
      const int slot = t & 1, r1 = r0 == HRING - 1 ? 0 : r0 + 1;
      const uint32_t sb_s = stag_s + slot * SLOT, hn_s = hs_s + r1 * NB * LDH * 2, hp_s = hs_s + r0 * NB * LDH * 2;
      nb_sync(2 + slot, NSYNC2);
      if (!(a.dbg & 2)) {
        uint4 vg[8], vc[4], vv[2], vh[2];
#pragma unroll
        for (int j = 0; j < 8; ++j) if (jg.dst[j] >= 0) vg[j] = lds16(sb_s + jg.src[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (jc0.dst[j] >= 0) vc[j] = lds16(sb_s + jc0.src[j]);
          if (jc1.dst[j] >= 0) vc[2 + j] = lds16(sb_s + jc1.src[j]);
        }
        if (jh.dst[0] >= 0) vh[0] = lds16(hn_s + jh.src[0]);
        if (jp.dst[0] >= 0) vh[1] = lds16(hp_s + jp.src[0]);
        if (jv0.dst[0] >= 0) {
          const uint4 x = lds16(sb_s + jv0.src[0]), y = lds16(sb_s + jv0.src[0] + 16);
          vv[0] = make_uint4(pack2(__uint_as_float(x.x), __uint_as_float(x.y)), pack2(__uint_as_float(x.z), __uint_as_float(x.w)),
                             pack2(__uint_as_float(y.x), __uint_as_float(y.y)), pack2(__uint_as_float(y.z), __uint_as_float(y.w)));
        }
        if (jv1.dst[0] >= 0) {
          const uint4 x = lds16(sb_s + jv1.src[0]), y = lds16(sb_s + jv1.src[0] + 16);
          vv[1] = make_uint4(pack2(__uint_as_float(x.x), __uint_as_float(x.y)), pack2(__uint_as_float(x.z), __uint_as_float(x.w)),
                             pack2(__uint_as_float(y.x), __uint_as_float(y.y)), pack2(__uint_as_float(y.z), __uint_as_float(y.w)));
        }
        const unsigned long long tg = jg.base + (unsigned long long)(t * jg.inc), tc = jc0.base + (unsigned long long)(t * jc0.inc);
        const unsigned long long tv = jv0.base + (unsigned long long)(t * jv0.inc), th = jh.base + (unsigned long long)(t * jh.inc);
        const unsigned long long tp = jp.base + (unsigned long long)(t * jp.inc);
#pragma unroll
        for (int j = 0; j < 8; ++j) if (jg.dst[j] >= 0) stg16(tg + jg.dst[j], vg[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (jc0.dst[j] >= 0) stg16(tc + jc0.dst[j], vc[j]);
          if (jc1.dst[j] >= 0) stg16(tc + jc1.dst[j], vc[2 + j]);
        }
        if (jh.dst[0] >= 0) stg16(th + jh.dst[0], vh[0]);
        if (jp.dst[0] >= 0) stg16(tp + jp.dst[0], vh[1]);
        if (jv0.dst[0] >= 0) stg16(tv + jv0.dst[0], vv[0]);
        if (jv1.dst[0] >= 0) stg16(tv + jv1.dst[0], vv[1]);
      }
      if (t + 2 < a.T) nb_arrive(4 + slot, NSYNC2);
      r0 = r1;
    }
    return;
  }
  // ---- compute warps: the loop is instantiated per (tiles of this warp, k-steps) ----
  const bool trace = (a.dbg & 32) && !(a.dbg & 16) && blockIdx.x == 0 && m == a.n_mods - 1 && lane == 0;
  const int nt = (H / 4 - warp + NW - 1) / NW;       // 16-row tiles 4 * (warp + 8 i) < H
#define LSTM2_GO(NT_, KS_) lstm_fwd2_compute<TRAIN, NT_, KS_>(a, m, hS, stages, stag, SBY, FST, GST, CST, SLOT, full, empty, trace)
  if (H <= 48) {
    if (nt == 2) LSTM2_GO(2, 3); else LSTM2_GO(1, 3);
  } else {
    if (nt == 3) LSTM2_GO(3, 6); else if (nt == 2) LSTM2_GO(2, 6); else LSTM2_GO(1, 6);
  }
#undef LSTM2_GO
}

// ======================================================================================================
// LSTM recurrence, backward (reverse time).  Warp w < ceil(H / 16) owns hidden units 16 w .. 16 w + 15 in BOTH roles: the
// element-wise cell backward of those units (dc in registers) and the rows of dh_{t-1} = W_hh^T dz (A = W_hh^T with
// k = gate * 96 + unit, 24 k-steps, fragments in registers), so dh never leaves the accumulator registers either.
// ======================================================================================================
constexpr int BKS = 24;            // k-steps over k = gate * 96 + unit
constexpr int LDZ = 4 * 96 + 8;    // dz tile row stride

__global__ void __launch_bounds__(NTH + 32, 1) lstm_bwd_mma_kernel(const __grid_constant__ LstmArgs a) {
  __shared__ __align__(16) bf16 dzS[2][NB * LDZ];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ Feed feed;
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  if (threadIdx.x == 0) {
    const int hb = H * (int)sizeof(float);
    feed.n = 0; feed.stage_bytes = 0; feed.tx_bytes = 0;
    feed.add(a.gates + 4 * hoff, (long long)4 * Hs * sizeof(float), 4 * hb);                           // 0: gate activations [gate][unit]
    feed.add(a.cstar + hoff, (long long)2 * Hs * sizeof(float), hb);                                   // 1: c_{t-1}
    feed.add(a.cstar + Hs + hoff, (long long)2 * Hs * sizeof(float), hb);                              // 2: c_t
    feed.add(a.dcstar + hoff, (long long)2 * Hs * sizeof(float), hb);                                  // 3: d cStar (prev half)
    feed.add(a.dcstar + Hs + hoff, (long long)2 * Hs * sizeof(float), hb);                             // 4: d cStar (new half)
    feed.add(a.dlast + hoff, (long long)(Hs + a.MEM) * sizeof(float), hb);                             // 5: d h_t from the head
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool active = warp < NW && warp * 16 < H;    // warp-uniform
  const bf16* W = reinterpret_cast<const bf16*>(a.w_hh[m]);
  uint32_t A[BKS][4];
  if (active) {
    auto get = [&](int r, int k) {                   // r = output unit, k = gate * 96 + unit
      const int gate = k / 96, u = k - gate * 96;
      return (r < H && u < H) ? bf(W + (size_t)(gate * H + u) * H + r) : 0.f;
    };
#pragma unroll
    for (int ks = 0; ks < BKS; ++ks) frag_a(A[ks], warp * 16, ks * 16, lane, get);
  }
  const int uu[2] = {warp * 16 + gid, warp * 16 + gid + 8};
  const int nn[2] = {2 * q, 2 * q + 1};
  bool val[2]; long long rbase[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) { val[p] = b0 + nn[p] < a.B; rbase[p] = (long long)min(b0 + nn[p], a.B - 1) * a.sb; }
  for (int e = threadIdx.x; e < 2 * NB * LDZ; e += NTH) dzS[0][e] = __float2bfloat16(0.f);
  const int LW = Hs + a.MEM;
  bf16* dz_op = reinterpret_cast<bf16*>(a.dz_op);
  float dh[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  char* stages = reinterpret_cast<char*>(feed_smem);
  const int SBY = feed.stage_bytes;
  if (warp == NW) {
    const int T = a.T;
    feed_producer(feed, stages, full, empty, FD, b0, a.B, a.sb, a.st, a.T, lane, [T](int i) { return T - 1 - i; });
    return;
  }

  {
    {
      const bool trace = (a.dbg & 32) && blockIdx.x == 0 && m == a.n_mods - 1 && threadIdx.x == 0;
      for (int i = 0; i < a.T; ++i) {
      const int t = a.T - 1 - i;
      REC_STAMP(1, i, 0, 0.f);
      if (!(a.dbg & 4)) mtrec::sbar_wait(&full[i % FD], (uint32_t)(i / FD) & 1u);
      REC_STAMP(1, i, 1, 0.f);
      const char* sg = stages + (i % FD) * SBY;
      bf16* dzo = dzS[i & 1];
      if (active) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int u = uu[v >> 1], n = nn[v & 1];
          if (u >= H) continue;
          const float gi = feed_at<float>(feed, sg, 0, n, u), gf = feed_at<float>(feed, sg, 0, n, H + u);
          const float gg = feed_at<float>(feed, sg, 0, n, 2 * H + u), go = feed_at<float>(feed, sg, 0, n, 3 * H + u);
          const float c_prev = feed_at<float>(feed, sg, 1, n, u), c_new = feed_at<float>(feed, sg, 2, n, u);
          const float tc = ftanh(c_new);
          const float dhv = dh[v] + feed_at<float>(feed, sg, 5, n, u);
          const float dcv = dc[v] + feed_at<float>(feed, sg, 4, n, u) + dhv * go * (1.f - tc * tc);
          const float zi = dcv * gg * gi * (1.f - gi);
          const float zf = dcv * c_prev * gf * (1.f - gf);
          const float zg = dcv * gi * (1.f - gg * gg);
          const float zo = dhv * tc * go * (1.f - go);
          dc[v] = feed_at<float>(feed, sg, 3, n, u) + dcv * gf;         // gradient wrt c_{t-1}
          bf16* zr = dzo + nn[v & 1] * LDZ + u;
          zr[0] = __float2bfloat16(zi); zr[96] = __float2bfloat16(zf); zr[192] = __float2bfloat16(zg); zr[288] = __float2bfloat16(zo);
          if (val[v & 1] && !(a.dbg & 2)) {
            bf16* zo_g = dz_op + (rbase[v & 1] + (long long)t * a.st) * (4 * Hs) + 4 * hoff + u;
            zo_g[0] = __float2bfloat16(zi); zo_g[H] = __float2bfloat16(zf); zo_g[2 * H] = __float2bfloat16(zg); zo_g[3 * H] = __float2bfloat16(zo);
          }
        }
      }
      REC_STAMP(1, i, 2, dc[0] + dc[3]);
      release_slot(&empty[i % FD], lane);
      REC_STAMP(1, i, 3, 0.f);
      consumers_sync();
      REC_STAMP(1, i, 4, 0.f);
      if (active) {                                  // dh_{t-1} = W_hh^T dz
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < BKS; ++ks) {
          if ((ks % 6) * 16 >= H) continue;          // zero-padded unit columns of this gate block
          uint32_t bb0, bb1;
          frag_b(bb0, bb1, dzo, LDZ, ks * 16, lane);
          mma16816(acc[ks & 3], A[ks], bb0, bb1);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) dh[v] = (acc[0][v] + acc[1][v]) + (acc[2][v] + acc[3][v]);
      }
      REC_STAMP(1, i, 5, dh[0] + dh[3]);
      }
    }
  }
}

// ======================================================================================================
// LSTM recurrence, backward, second cut.  The first cut spent its step waiting for its input feed: six stash tensors x eight
// narratives = 48 small bulk copies per step, and the copy engine issues ~80 clocks apiece (3 900 clocks per step measured, 1 400-1 700
// of them on the `full` barrier).  Here
//   * the feed is SIX 3-D TMA boxes per step ({columns, 1 step, 8 narratives} of the [B, T, C] view of each stash tensor; boxes are 4
//     floats wider than the modality's span so that the 8 narrative rows land on different banks; narratives beyond B are zero-filled
//     by the copy engine); c_{t-1} is not fetched at all -- it is c_t of the NEXT ring stage;
//   * 8 warps, 255 registers: warps 0 .. ceil(H / 16) - 1 own 16 hidden units each in both roles (cell backward, rows of
//     dh_{t-1} = W_hh^T dz with k = gate * H + unit, fragments in registers), warp 6 copies dz_t (bf16, already laid out like its
//     [M, 4 Hs] row segment) to global memory with 16-byte stores while the next step computes, one lane of warp 7 feeds the ring.
// ======================================================================================================
struct RecMaps { CUtensorMap m[MT_MAX_MODS][4]; };       // per modality: gate activations, cStar (c_t half), d cStar, d h
constexpr int BFD = 4;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(mtrec::s_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int KS2>
__device__ __forceinline__ void lstm_bwd2_compute(const LstmArgs& a, int m, const char* stages, char* dzS, int STAGE, int GB, int HB, int DZS,
                                                  int ncw, uint64_t* full, uint64_t* empty) {
  static_assert(KS2 % 2 == 0, "k-steps are consumed in ldmatrix.x4 pairs");
  const int H = a.H[m];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const bf16* W = reinterpret_cast<const bf16*>(a.w_hh[m]);
  uint32_t A[KS2][4];
  {
    auto get = [&](int r, int k) { return (r < H && k < 4 * H) ? bf(W + (size_t)k * H + r) : 0.f; };      // A = W_hh^T, k = gate * H + unit
#pragma unroll
    for (int ks = 0; ks < KS2; ++ks) frag_a(A[ks], warp * 16, ks * 16, lane, get);
  }
  const int RG = (2 * H + 4) * 4, RH = (H + 4) * 4;   // row strides of the gate boxes / the H-wide boxes
  int og[4], oh[4], od[4];
  bool on[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int u = warp * 16 + gid + 8 * (v >> 1), n = 2 * q + (v & 1);
    on[v] = u < H;
    og[v] = n * RG + u * 4; oh[v] = n * RH + u * 4; od[v] = n * DZS + u * 2;
  }
  const int nsync = (ncw + 1) * 32;
  const uint32_t dz_s = mtrec::s_u32(dzS) + (uint32_t)((lane & 7) * DZS + (lane >> 3) * 16);
  float dh[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
  mtrec::sbar_wait(&full[0], 0u);
  for (int i = 0; i < a.T; ++i) {
    const int slot = i & 1;
    const char* sg = stages + (i % BFD) * STAGE;
    const bool has_prev = i + 1 < a.T;
    if (has_prev) mtrec::sbar_wait(&full[(i + 1) % BFD], (uint32_t)((i + 1) / BFD) & 1u);
    const char* sgn = stages + ((i + 1) % BFD) * STAGE;
    float zi[4], zf[4], zg[4], zo[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float gi = lds_f32(sg + og[v]), gf = lds_f32(sg + og[v] + 4 * H);
      const float gg = lds_f32(sg + GB + og[v]), go = lds_f32(sg + GB + og[v] + 4 * H);
      const float c_new = lds_f32(sg + 2 * GB + oh[v]);
      const float c_prev = has_prev ? lds_f32(sgn + 2 * GB + oh[v]) : 0.f;
      const float dcp = lds_f32(sg + 2 * GB + HB + oh[v]), dcn = lds_f32(sg + 2 * GB + 2 * HB + oh[v]);
      const float dl = lds_f32(sg + 2 * GB + 3 * HB + oh[v]);
      const float tc = ftanh(c_new);
      const float dhv = dh[v] + dl;
      const float dcv = dc[v] + dcn + dhv * go * (1.f - tc * tc);
      zi[v] = dcv * gg * gi * (1.f - gi);
      zf[v] = dcv * c_prev * gf * (1.f - gf);
      zg[v] = dcv * gi * (1.f - gg * gg);
      zo[v] = dhv * tc * go * (1.f - go);
      dc[v] = dcp + dcv * gf;                          // gradient wrt c_{t-1}
    }
    release_slot(&empty[i % BFD], lane);
    if (i >= 2) nb_sync(4 + slot, nsync);              // the storer is done with this dz slot
    char* dzo = dzS + slot * NB * DZS;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (!on[v]) continue;
      bf16* zr = reinterpret_cast<bf16*>(dzo + od[v]);
      zr[0] = __float2bfloat16(zi[v]); zr[H] = __float2bfloat16(zf[v]); zr[2 * H] = __float2bfloat16(zg[v]); zr[3 * H] = __float2bfloat16(zo[v]);
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
    nb_arrive(2 + slot, nsync);                        // staged: the storer takes dz_t from here
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const uint32_t src = dz_s + (uint32_t)(slot * NB * DZS);
#pragma unroll
    for (int kp = 0; kp < KS2 / 2; ++kp) {
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(src + kp * 64));
      mma16816(acc[(2 * kp) & 3], A[2 * kp], b[0], b[1]);
      mma16816(acc[(2 * kp + 1) & 3], A[2 * kp + 1], b[2], b[3]);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) dh[v] = (acc[0][v] + acc[1][v]) + (acc[2][v] + acc[3][v]);
  }
}

__global__ void __launch_bounds__(NTH, 1) lstm_bwd_mma2_kernel(const __grid_constant__ LstmArgs a, const __grid_constant__ RecMaps maps) {
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ __align__(8) uint64_t full[BFD], empty[BFD];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * NB;
  const int ncw = (H + 15) / 16;                       // compute warps (H <= 96: at most 6)
  const int ks2 = H <= 48 ? 12 : (H <= 88 ? 22 : 24);
  const int GB = NB * (2 * H + 4) * 4, HB = NB * (H + 4) * 4;
  const int STAGE = 2 * GB + 4 * HB;
  const int DZS = ks2 * 32 + 16;                       // dz row: k = gate * H + unit, bf16; 16 bytes of padding: conflict-free ldmatrix
  char* stages = reinterpret_cast<char*>(feed_smem);
  char* dzS = stages + (size_t)BFD * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < BFD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], ncw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = threadIdx.x; e < 2 * NB * DZS / 4; e += NTH) reinterpret_cast<uint32_t*>(dzS)[e] = 0u;
  __syncthreads();
  if (warp == NW - 1) {                                // ---- producer ----
    if (lane == 0) {
      const CUtensorMap* mg = &maps.m[m][0]; const CUtensorMap* mc = &maps.m[m][1];
      const CUtensorMap* md = &maps.m[m][2]; const CUtensorMap* ml = &maps.m[m][3];
      for (int i = 0; i < a.T; ++i) {
        const int slot = i % BFD, t = a.T - 1 - i;
        if (i >= BFD) mtrec::sbar_wait(&empty[slot], (uint32_t)(i / BFD - 1) & 1u);
        mtrec::sbar_expect_tx(&full[slot], (uint32_t)STAGE);
        const uint32_t d = mtrec::s_u32(stages + (size_t)slot * STAGE);
        tma_load_3d(d, mg, 0, t, b0, &full[slot]);                        // gates i | f
        tma_load_3d(d + GB, mg, 2 * H, t, b0, &full[slot]);               // gates g | o
        tma_load_3d(d + 2 * GB, mc, 0, t, b0, &full[slot]);               // c_t
        tma_load_3d(d + 2 * GB + HB, md, 0, t, b0, &full[slot]);          // d cStar, c_{t-1} half
        tma_load_3d(d + 2 * GB + 2 * HB, md, Hs, t, b0, &full[slot]);     // d cStar, c_t half
        tma_load_3d(d + 2 * GB + 3 * HB, ml, 0, t, b0, &full[slot]);      // d h_t from the head
      }
    }
    return;
  }
  if (warp == NW - 2) {                                // ---- storer: dz_t -> [M, 4 Hs] bf16 ----
    StoreJobs<12> jz;
    jz.build(a.dz_op, 8ll * Hs, 8ll * hoff, H / 2, DZS, 0, 16, b0, a.B, a.sb, a.st, lane, 32);
    const uint32_t dz_s = mtrec::s_u32(dzS);
    const int nsync = (ncw + 1) * 32;
    for (int i = 0; i < a.T; ++i) {
      const int slot = i & 1;
      const long long t = a.T - 1 - i;
      nb_sync(2 + slot, nsync);
      uint4 v[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) if (jz.dst[j] >= 0) v[j] = lds16(dz_s + slot * NB * DZS + jz.src[j]);
      const unsigned long long tz = jz.base + (unsigned long long)(t * jz.inc);
#pragma unroll
      for (int j = 0; j < 12; ++j) if (jz.dst[j] >= 0) stg16(tz + jz.dst[j], v[j]);
      if (i + 2 < a.T) nb_arrive(4 + slot, nsync);
    }
    return;
  }
  if (warp >= ncw) return;
  if (ks2 == 12) lstm_bwd2_compute<12>(a, m, stages, dzS, STAGE, GB, HB, DZS, ncw, full, empty);
  else if (ks2 == 22) lstm_bwd2_compute<22>(a, m, stages, dzS, STAGE, GB, HB, DZS, ncw, full, empty);
  else lstm_bwd2_compute<24>(a, m, stages, dzS, STAGE, GB, HB, DZS, ncw, full, empty);
}

// ======================================================================================================
// Memory recurrence, second cut (same ideas as the LSTM kernels above): 3-D TMA boxes feed the ring (2 per step forward, 6 per step
// backward instead of 16 / 40 bulk copies), storer warps move the step's outputs out of shared-memory staging with 16-byte stores,
// mem_t lives in a 3-deep bf16 ring that is both the next step's B operand and the source of the mem_{t-1} / mem_t stash rows,
// sigmoids on the tanh unit, the dropout factors of a step are drawn before the step's first barrier.
// ======================================================================================================
struct MemMaps { CUtensorMap m[6]; };
constexpr int MBOX = (128 + 4) * 4 * NB;             // bytes of an fp32 box {132, 1, 8}
constexpr int MBOXH = (128 + 8) * 2 * NB;            // bytes of a bf16 box {136, 1, 8}
constexpr int MROW = (128 + 4) * 4, MROWH = (128 + 8) * 2;
constexpr int GMS = 256 * 4 + 16;                    // staging row of gamma1 | gamma2 (fp32)

template <bool TRAIN>
__global__ void __launch_bounds__(NTH2, 1) mem_fwd_mma2_kernel(const __grid_constant__ MemArgs a, const __grid_constant__ MemMaps maps) {
  __shared__ __align__(16) bf16 memS[HRING][NB * LDK];
  __shared__ __align__(16) bf16 ghS[2][NB * LDK];
  __shared__ __align__(16) unsigned char gmS[2][NB * GMS];
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int MEM = 128, G = 64, G2 = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const int LW = a.Hs + MEM;
  char* stages = reinterpret_cast<char*>(feed_smem);
  constexpr int STAGE = 2 * MBOX;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = threadIdx.x; e < HRING * NB * LDK; e += NTH2) memS[0][e] = __float2bfloat16(0.f);
  for (int e = threadIdx.x; e < 2 * NB * LDK; e += NTH2) ghS[0][e] = __float2bfloat16(0.f);
  __syncthreads();
  if (warp == NW) {                                  // ---- producer ----
    if (lane == 0) {
      for (int t = 0; t < a.T; ++t) {
        const int slot = t % FD;
        if (t >= FD) mtrec::sbar_wait(&empty[slot], (uint32_t)(t / FD - 1) & 1u);
        mtrec::sbar_expect_tx(&full[slot], (uint32_t)STAGE);
        const uint32_t d = mtrec::s_u32(stages + (size_t)slot * STAGE);
        tma_load_3d(d, &maps.m[0], 0, t, b0, &full[slot]);                // gate pre-activations (batched part)
        tma_load_3d(d + MBOX, &maps.m[1], 0, t, b0, &full[slot]);         // cHat
      }
    }
    return;
  }
  if (warp > NW) {                                   // ---- storers ----
    const int tid = threadIdx.x - (NW + 1) * 32;
    StoreJobs<6> jgm; StoreJobs<2> jgh, jmp, jml;
    jgm.build(TRAIN ? a.gm : nullptr, 4ll * 2 * MEM, 0, 64, GMS, 0, 16, b0, a.B, a.sb, a.st, tid);
    jgh.build(TRAIN ? a.gh_op : nullptr, 2ll * G2, 0, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    jmp.build(TRAIN ? a.memprev_op : nullptr, 2ll * MEM, 0, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    jml.build(a.last_op, 2ll * LW, 2ll * a.Hs, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    const uint32_t gm_s = mtrec::s_u32(&gmS[0][0]), gh_s = mtrec::s_u32(&ghS[0][0]), mem_s = mtrec::s_u32(&memS[0][0]);
    int r0 = 0;
    for (int t = 0; t < a.T; ++t) {
      const int slot = t & 1, r1 = r0 == HRING - 1 ? 0 : r0 + 1;
      nb_sync(2 + slot, NSYNC2);
      if (a.dbg & 2) { if (t + 2 < a.T) nb_arrive(4 + slot, NSYNC2); r0 = r1; continue; }
      uint4 vg[6], vh[2], vp[2], vl[2];
#pragma unroll
      for (int j = 0; j < 6; ++j) if (jgm.dst[j] >= 0) vg[j] = lds16(gm_s + slot * NB * GMS + jgm.src[j]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (jgh.dst[j] >= 0) vh[j] = lds16(gh_s + slot * NB * LDK * 2 + jgh.src[j]);
        if (jmp.dst[j] >= 0) vp[j] = lds16(mem_s + r0 * NB * LDK * 2 + jmp.src[j]);
        if (jml.dst[j] >= 0) vl[j] = lds16(mem_s + r1 * NB * LDK * 2 + jml.src[j]);
      }
      const unsigned long long tg = jgm.base + (unsigned long long)(t * jgm.inc), th = jgh.base + (unsigned long long)(t * jgh.inc);
      const unsigned long long tp = jmp.base + (unsigned long long)(t * jmp.inc), tl = jml.base + (unsigned long long)(t * jml.inc);
#pragma unroll
      for (int j = 0; j < 6; ++j) if (jgm.dst[j] >= 0) stg16(tg + jgm.dst[j], vg[j]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (jgh.dst[j] >= 0) stg16(th + jgh.dst[j], vh[j]);
        if (jmp.dst[j] >= 0) stg16(tp + jmp.dst[j], vp[j]);
        if (jml.dst[j] >= 0) stg16(tl + jml.dst[j], vl[j]);
      }
      if (t + 2 < a.T) nb_arrive(4 + slot, NSYNC2);
      r0 = r1;
    }
    return;
  }
  // ---- compute warps: warp w owns features f0 = 16 w + gid, f1 = f0 + 8 of narratives n0 = 2 q, n1 = n0 + 1 in every layer ----
  const int f0 = warp * 16 + gid, f1 = f0 + 8;
  const int ldw1 = 2 * a.Hs + MEM;
  uint32_t A1[8][4], A2a[4][4], A2b[4][4];
  {
    const bf16* w1 = reinterpret_cast<const bf16*>(warp < 4 ? a.g1_fc1_w : a.g2_fc1_w) + 2 * a.Hs;      // mem columns of gamma{1,2}_fc1
    const int rb = (warp & 3) * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) frag_a(A1[ks], rb, ks * 16, lane, [&](int r, int c) { return bf(w1 + (size_t)r * ldw1 + c); });
    const bf16* v1 = reinterpret_cast<const bf16*>(a.g1_fc2_w);
    const bf16* v2 = reinterpret_cast<const bf16*>(a.g2_fc2_w);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      frag_a(A2a[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return bf(v1 + (size_t)r * G + c); });
      frag_a(A2b[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return bf(v2 + (size_t)r * G + c); });
    }
  }
  const float bias1[2] = {a.g1_fc2_b[f0], a.g1_fc2_b[f1]}, bias2[2] = {a.g2_fc2_b[f0], a.g2_fc2_b[f1]};
  const DropCfg drop = mt_drop_resolve(warp < 4 ? a.drop_g1 : a.drop_g2);
  const int jg[2] = {f0 & 63, f1 & 63};              // index inside the gate's own [T,B,G] dropout tensor
  int of[4], ob[4], og[4];                           // byte offsets: fp32 feed box, bf16 [n][LDK] tiles, gamma staging
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int n = 2 * q + (v & 1), f = (v >> 1) ? f1 : f0;
    of[v] = n * MROW + f * 4; ob[v] = (n * LDK + f) * 2; og[v] = n * GMS + f * 4;
  }
  const uint32_t lm_off = (uint32_t)(((lane & 7) * LDK + (lane >> 3) * 8) * 2);      // ldmatrix.x4 row of this lane inside a [n][LDK] tile
  const uint32_t mem_s = mtrec::s_u32(&memS[0][0]) + lm_off, gh_s = mtrec::s_u32(&ghS[0][0]) + lm_off;
  float mem[4] = {0.f, 0.f, 0.f, 0.f};
  int r0 = 0;
  const bool trace = (a.dbg & 32) && (a.dbg & 16) && blockIdx.x == 0 && lane == 0;
  mtrec::sbar_wait(&full[0], 0u);
  for (int t = 0; t < a.T; ++t) {
    const int slot = t & 1, r1 = r0 == HRING - 1 ? 0 : r0 + 1;
    REC_STAMP2(t, 0, 0.f);
    const char* sg = stages + (t % FD) * STAGE;
    float acc[2][4], df[4], ch[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      acc[0][v] = lds_f32(sg + of[v]); acc[1][v] = 0.f;
      ch[v] = lds_f32(sg + MBOX + of[v]);
      // element index of the gate's [T,B,G] tensor (oracle/mt_oracle.py:_drop_t)
      df[v] = mt_drop_factor(drop, ((uint64_t)t * a.B + (uint64_t)(b0 + 2 * q + (v & 1))) * (uint64_t)G + (uint64_t)jg[v >> 1]);
    }
    REC_STAMP2(t, 1, df[0] + df[3] + acc[0][0]);
    consumers_sync();                                // mem_{t-1} of every warp is in ring slot r0
    // ---- layer 1: gh = drop(relu(gpre_t + Wm mem_{t-1})) ----
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(mem_s + r0 * NB * LDK * 2 + kp * 64));
      mma16816(acc[0], A1[2 * kp], b[0], b[1]);
      mma16816(acc[1], A1[2 * kp + 1], b[2], b[3]);
    }
    REC_STAMP2(t, 2, 0.f);
    release_slot(&empty[t % FD], lane);
    if (t + 1 < a.T) mtrec::sbar_wait(&full[(t + 1) % FD], (uint32_t)((t + 1) / FD) & 1u);      // long complete: overlaps the mma
    if (t >= 2) nb_sync(4 + slot, NSYNC2);           // the storers are done with this staging slot and with ring slot r1
    REC_STAMP2(t, 3, acc[0][0] + acc[1][0]);
    char* gho = reinterpret_cast<char*>(ghS[slot]);
#pragma unroll
    for (int v = 0; v < 4; ++v) *reinterpret_cast<bf16*>(gho + ob[v]) = __float2bfloat16(fmaxf(acc[0][v] + acc[1][v], 0.f) * df[v]);
    REC_STAMP2(t, 4, 0.f);
    consumers_sync();
    // ---- layer 2: gamma1 from hidden[0:64), gamma2 from hidden[64:128) ----
    float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t x[4], y[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(gh_s + slot * NB * LDK * 2 + kp * 64));
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]) : "r"(gh_s + slot * NB * LDK * 2 + 128 + kp * 64));
      mma16816(c1, A2a[2 * kp], x[0], x[1]); mma16816(c2, A2b[2 * kp], y[0], y[1]);
      mma16816(c1, A2a[2 * kp + 1], x[2], x[3]); mma16816(c2, A2b[2 * kp + 1], y[2], y[3]);
    }
    REC_STAMP2(t, 5, c1[0] + c2[0]);
    char* mo = reinterpret_cast<char*>(memS[r1]);
    char* go = reinterpret_cast<char*>(gmS[slot]);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float g1 = act_tanh(c1[v] + bias1[v >> 1], 0.5f, 0.5f, 0.5f), g2 = act_tanh(c2[v] + bias2[v >> 1], 0.5f, 0.5f, 0.5f);
      mem[v] = g1 * mem[v] + g2 * ch[v];
      *reinterpret_cast<bf16*>(mo + ob[v]) = __float2bfloat16(mem[v]);
      if (TRAIN) { *reinterpret_cast<float*>(go + og[v]) = g1; *reinterpret_cast<float*>(go + og[v] + MEM * 4) = g2; }
    }
    REC_STAMP2(t, 6, mem[0]);
    nb_arrive(2 + slot, NSYNC2);                     // staged: the storers take step t from here
    REC_STAMP2(t, 7, 0.f);
    r0 = r1;
  }
  if (a.mem_last) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int n = 2 * q + (v & 1);
      if (b0 + n < a.B) a.mem_last[(size_t)(b0 + n) * MEM + ((v >> 1) ? f1 : f0)] = mem[v];
    }
  }
}

// backward:  g = dmem + d(mem_t from the head);  dzg1 = g mem_{t-1} g1 (1 - g1);  dzg2 = g cHat g2 (1 - g2);  dzchat = g g2 (1 - cHat^2);
//            dmem = g g1;  dgh = [gh > 0] sc * (W21^T dzg1 | W22^T dzg2);  dmem += Wm^T dgh
__global__ void __launch_bounds__(NTH2, 1) mem_bwd_mma2_kernel(const __grid_constant__ MemArgs a, const __grid_constant__ MemMaps maps) {
  __shared__ __align__(16) bf16 dzS[2][2][NB * LDK];       // [slot][gate]: dzg1 / dzg2, B operands and the dzg stash rows
  __shared__ __align__(16) bf16 dghS[2][NB * LDK];
  __shared__ __align__(16) bf16 dcS[2][NB * LDK];          // dzchat stash rows
  extern __shared__ __align__(128) unsigned char feed_smem[];
  __shared__ __align__(8) uint64_t full[FD], empty[FD];
  const int MEM = 128, G = 64, G2 = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  char* stages = reinterpret_cast<char*>(feed_smem);
  constexpr int STAGE = 4 * MBOX + 2 * MBOXH;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FD; ++i) { mtrec::sbar_init(&full[i], 1); mtrec::sbar_init(&empty[i], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = threadIdx.x; e < 4 * NB * LDK; e += NTH2) dzS[0][0][e] = __float2bfloat16(0.f);
  for (int e = threadIdx.x; e < 2 * NB * LDK; e += NTH2) { dghS[0][e] = __float2bfloat16(0.f); dcS[0][e] = __float2bfloat16(0.f); }
  __syncthreads();
  if (warp == NW) {                                  // ---- producer ----
    if (lane == 0) {
      for (int i = 0; i < a.T; ++i) {
        const int slot = i % FD, t = a.T - 1 - i;
        if (i >= FD) mtrec::sbar_wait(&empty[slot], (uint32_t)(i / FD - 1) & 1u);
        mtrec::sbar_expect_tx(&full[slot], (uint32_t)STAGE);
        const uint32_t d = mtrec::s_u32(stages + (size_t)slot * STAGE);
        tma_load_3d(d, &maps.m[0], 0, t, b0, &full[slot]);                        // d mem_t from the head
        tma_load_3d(d + MBOX, &maps.m[1], 0, t, b0, &full[slot]);                 // gamma1
        tma_load_3d(d + 2 * MBOX, &maps.m[1], MEM, t, b0, &full[slot]);           // gamma2
        tma_load_3d(d + 3 * MBOX, &maps.m[2], 0, t, b0, &full[slot]);             // cHat
        tma_load_3d(d + 4 * MBOX, &maps.m[3], 0, t, b0, &full[slot]);             // mem_{t-1} (bf16)
        tma_load_3d(d + 4 * MBOX + MBOXH, &maps.m[4], 0, t, b0, &full[slot]);     // gamma hidden (bf16)
      }
    }
    return;
  }
  if (warp > NW) {                                   // ---- storers ----
    const int tid = threadIdx.x - (NW + 1) * 32;
    StoreJobs<2> j1, j2, jc, jh;
    j1.build(a.dzg_op, 2ll * 2 * MEM, 0, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    j2.build(a.dzg_op, 2ll * 2 * MEM, 2ll * MEM, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    jc.build(a.dzchat_op, 2ll * MEM, 0, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    jh.build(a.dgh_op, 2ll * G2, 0, 16, LDK * 2, 0, 16, b0, a.B, a.sb, a.st, tid);
    const uint32_t dz_s = mtrec::s_u32(&dzS[0][0][0]), dgh_s = mtrec::s_u32(&dghS[0][0]), dc_s = mtrec::s_u32(&dcS[0][0]);
    for (int i = 0; i < a.T; ++i) {
      const int slot = i & 1;
      const long long t = a.T - 1 - i;
      nb_sync(2 + slot, NSYNC2);
      uint4 v1[2], v2[2], vc[2], vh[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j1.dst[j] >= 0) v1[j] = lds16(dz_s + (slot * 2) * NB * LDK * 2 + j1.src[j]);
        if (j2.dst[j] >= 0) v2[j] = lds16(dz_s + (slot * 2 + 1) * NB * LDK * 2 + j2.src[j]);
        if (jc.dst[j] >= 0) vc[j] = lds16(dc_s + slot * NB * LDK * 2 + jc.src[j]);
        if (jh.dst[j] >= 0) vh[j] = lds16(dgh_s + slot * NB * LDK * 2 + jh.src[j]);
      }
      const unsigned long long t1 = j1.base + (unsigned long long)(t * j1.inc), tc = jc.base + (unsigned long long)(t * jc.inc);
      const unsigned long long th = jh.base + (unsigned long long)(t * jh.inc);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j1.dst[j] >= 0) stg16(t1 + j1.dst[j], v1[j]);
        if (j2.dst[j] >= 0) stg16(t1 + j2.dst[j], v2[j]);
        if (jc.dst[j] >= 0) stg16(tc + jc.dst[j], vc[j]);
        if (jh.dst[j] >= 0) stg16(th + jh.dst[j], vh[j]);
      }
      if (i + 2 < a.T) nb_arrive(4 + slot, NSYNC2);
    }
    return;
  }
  // ---- compute warps ----
  const int f0 = warp * 16 + gid, f1 = f0 + 8;
  const int ldw1 = 2 * a.Hs + MEM;
  // A3: rows = hidden index n (warp's 16), k = mem feature: n < G: gamma1_fc2[k][n], else gamma2_fc2[k][n - G]
  // A4: rows = mem feature (warp's 16), k = hidden index: k < G: gamma1_fc1[k][2Hs + row], else gamma2_fc1[k - G][2Hs + row]
  uint32_t A3[8][4], A4[8][4];
  {
    const bf16* v = reinterpret_cast<const bf16*>(warp < 4 ? a.g1_fc2_w : a.g2_fc2_w);
    const int nb = (warp & 3) * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) frag_a(A3[ks], nb, ks * 16, lane, [&](int r, int c) { return bf(v + (size_t)c * G + r); });
    const bf16* w1 = reinterpret_cast<const bf16*>(a.g1_fc1_w) + 2 * a.Hs;
    const bf16* w2 = reinterpret_cast<const bf16*>(a.g2_fc1_w) + 2 * a.Hs;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      frag_a(A4[ks], warp * 16, ks * 16, lane, [&](int r, int c) { return c < G ? bf(w1 + (size_t)c * ldw1 + r) : bf(w2 + (size_t)(c - G) * ldw1 + r); });
  }
  int of[4], oh[4], ob[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int n = 2 * q + (v & 1), f = (v >> 1) ? f1 : f0;
    of[v] = n * MROW + f * 4; oh[v] = n * MROWH + f * 2; ob[v] = (n * LDK + f) * 2;
  }
  const uint32_t lm_off = (uint32_t)(((lane & 7) * LDK + (lane >> 3) * 8) * 2);
  const uint32_t dz_s = mtrec::s_u32(&dzS[0][0][0]) + lm_off, dgh_s = mtrec::s_u32(&dghS[0][0]) + lm_off;
  const float sc_g = a.drop_g1.scale;
  float dmem[4] = {0.f, 0.f, 0.f, 0.f};
  mtrec::sbar_wait(&full[0], 0u);
  for (int i = 0; i < a.T; ++i) {
    const int slot = i & 1;
    const char* sg = stages + (i % FD) * STAGE;
    float d1[4], d2[4], dcv[4], ghv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float dl = lds_f32(sg + of[v]), g1 = lds_f32(sg + MBOX + of[v]), g2 = lds_f32(sg + 2 * MBOX + of[v]), ch = lds_f32(sg + 3 * MBOX + of[v]);
      const float mp = __bfloat162float(*reinterpret_cast<const bf16*>(sg + 4 * MBOX + oh[v]));
      ghv[v] = __bfloat162float(*reinterpret_cast<const bf16*>(sg + 4 * MBOX + MBOXH + oh[v]));
      const float g = dmem[v] + dl;
      d1[v] = g * mp * g1 * (1.f - g1);
      d2[v] = g * ch * g2 * (1.f - g2);
      dcv[v] = g * g2 * (1.f - ch * ch);
      dmem[v] = g * g1;
    }
    release_slot(&empty[i % FD], lane);
    if (i + 1 < a.T) mtrec::sbar_wait(&full[(i + 1) % FD], (uint32_t)((i + 1) / FD) & 1u);
    if (i >= 2) nb_sync(4 + slot, NSYNC2);           // the storers are done with this slot's dz / dgh / dzchat tiles
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      *reinterpret_cast<bf16*>(reinterpret_cast<char*>(dzS[slot][0]) + ob[v]) = __float2bfloat16(d1[v]);
      *reinterpret_cast<bf16*>(reinterpret_cast<char*>(dzS[slot][1]) + ob[v]) = __float2bfloat16(d2[v]);
      *reinterpret_cast<bf16*>(reinterpret_cast<char*>(dcS[slot]) + ob[v]) = __float2bfloat16(dcv[v]);
    }
    consumers_sync();
    // ---- d hidden (hidden index = this thread's f0 / f1: warps 0-3 gate 1, warps 4-7 gate 2) ----
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const uint32_t src = dz_s + (uint32_t)((slot * 2 + (warp < 4 ? 0 : 1)) * NB * LDK * 2);
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(src + kp * 64));
      mma16816(acc[0], A3[2 * kp], b[0], b[1]);
      mma16816(acc[1], A3[2 * kp + 1], b[2], b[3]);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v)
      *reinterpret_cast<bf16*>(reinterpret_cast<char*>(dghS[slot]) + ob[v]) = __float2bfloat16(ghv[v] > 0.f ? (acc[0][v] + acc[1][v]) * sc_g : 0.f);
    consumers_sync();
    nb_arrive(2 + slot, NSYNC2);                     // staged: dzg, dzchat and dgh of step t
    // ---- d mem_{t-1} += gamma_fc1[:, 2H:]^T d hidden ----
    float ac2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const uint32_t sr2 = dgh_s + (uint32_t)(slot * NB * LDK * 2);
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(sr2 + kp * 64));
      mma16816(ac2[0], A4[2 * kp], b[0], b[1]);
      mma16816(ac2[1], A4[2 * kp + 1], b[2], b[3]);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) dmem[v] += ac2[0][v] + ac2[1][v];
  }
}

}  // namespace

// fp32 [B, T, cols] view of a stash tensor (row (b, t) at (b * sb + t * st) * pitch) as a 3-D tensor map, box {box_cols, 1, NB}
static int make_map_rows(CUtensorMap* map, const void* base, long long cols, long long pitch, int T, int B, long long sb, long long st, int box_cols,
                         int esize = 4) {
  tc5::EncodeTiledFn enc = tc5::get_encode();
  if (!enc) return MT_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)(st * pitch * esize), (cuuint64_t)(sb * pitch * esize)};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, 1u, (cuuint32_t)NB}, estr[3] = {1, 1, 1};
  CUresult r = enc(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "cuTensorMapEncodeTiled (recurrence feed) failed with CUresult %d", (int)r);
    return MT_ERR_CUDA;
  }
  return MT_OK;
}

extern "C" int mt_mfn_rec_trace2(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_rec_trace2, sizeof(unsigned long long) * 8 * 128 * 8) == cudaSuccess ? MT_OK : MT_ERR_CUDA;
}
extern "C" int mt_mfn_rec_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_rec_trace, sizeof(unsigned long long) * 4 * 128 * 8) == cudaSuccess ? MT_OK : MT_ERR_CUDA;
}

// dynamic shared memory of a feed ring whose per-narrative spans are `spans` bytes in total over `nseg` segments
static size_t feed_bytes(int spans, int nseg) { return (size_t)FD * NB * (size_t)(spans + nseg * FEED_PAD); }

bool mt_mfn_mma_mem_supported(const MemArgs& a) { return a.MEM == 128 && a.G == 64 && (a.Hs % 4) == 0; }

int mt_mfn_mma_mem_fwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
  if (a.Hs % 8 == 0 && !(a.dbg & 64)) {
    MemMaps maps;
    memset(&maps, 0, sizeof(maps));
    MT_TRY(make_map_rows(&maps.m[0], a.gpre, 128, 128, a.T, a.B, a.sb, a.st, 132));
    MT_TRY(make_map_rows(&maps.m[1], a.chat, 128, 128, a.T, a.B, a.sb, a.st, 132));
    const size_t smem2 = (size_t)FD * 2 * MBOX;
    if (a.training) { MT_TRY(mtrec::set_smem(mem_fwd_mma2_kernel<true>, smem2)); mem_fwd_mma2_kernel<true><<<grid, NTH2, smem2, st>>>(a, maps); }
    else { MT_TRY(mtrec::set_smem(mem_fwd_mma2_kernel<false>, smem2)); mem_fwd_mma2_kernel<false><<<grid, NTH2, smem2, st>>>(a, maps); }
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  const size_t smem = feed_bytes(2 * 128 * 4, 2);
  if (a.training) { MT_TRY(mtrec::set_smem(mem_fwd_mma_kernel<true>, smem)); mem_fwd_mma_kernel<true><<<grid, NTH + 32, smem, st>>>(a); }
  else { MT_TRY(mtrec::set_smem(mem_fwd_mma_kernel<false>, smem)); mem_fwd_mma_kernel<false><<<grid, NTH + 32, smem, st>>>(a); }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_mma_mem_bwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
  if (a.Hs % 8 == 0 && !(a.dbg & 64)) {
    MemMaps maps;
    memset(&maps, 0, sizeof(maps));
    MT_TRY(make_map_rows(&maps.m[0], a.dlast + a.Hs, 128, a.Hs + 128, a.T, a.B, a.sb, a.st, 132));
    MT_TRY(make_map_rows(&maps.m[1], a.gm, 256, 256, a.T, a.B, a.sb, a.st, 132));
    MT_TRY(make_map_rows(&maps.m[2], a.chat, 128, 128, a.T, a.B, a.sb, a.st, 132));
    MT_TRY(make_map_rows(&maps.m[3], a.memprev_op, 128, 128, a.T, a.B, a.sb, a.st, 136, 2));
    MT_TRY(make_map_rows(&maps.m[4], a.gh_op, 128, 128, a.T, a.B, a.sb, a.st, 136, 2));
    const size_t smem2 = (size_t)FD * (4 * MBOX + 2 * MBOXH);
    MT_TRY(mtrec::set_smem(mem_bwd_mma2_kernel, smem2));
    mem_bwd_mma2_kernel<<<grid, NTH2, smem2, st>>>(a, maps);
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  const size_t smem = feed_bytes(128 * 4 + 256 * 4 + 128 * 4 + 128 * 2 + 128 * 2, 5);
  MT_TRY(mtrec::set_smem(mem_bwd_mma_kernel, smem));
  mem_bwd_mma_kernel<<<grid, NTH + 32, smem, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

bool mt_mfn_mma_lstm_supported(const LstmArgs& a) {
  for (int m = 0; m < a.n_mods; ++m)
    if (a.H[m] > 96 || a.H[m] % 4 != 0 || a.hoff[m] % 4 != 0) return false;      // 16-byte aligned spans for the bulk copies
  return a.Hs % 4 == 0 && a.MEM % 4 == 0;
}

// the storer-warp kernels move 16-byte chunks of bf16 rows: every span must start and end on a 16-byte boundary
static bool lstm_v2_ok(const LstmArgs& a) {
  for (int m = 0; m < a.n_mods; ++m)
    if (a.H[m] % 8 != 0 || a.hoff[m] % 8 != 0 || a.H[m] < 32) return false;      // >= 8 tiles: every compute warp owns one
  return a.Hs % 8 == 0 && a.MEM % 8 == 0 && !(a.dbg & 64);
}

int mt_mfn_mma_lstm_fwd(const LstmArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + NB - 1) / NB, a.n_mods);
  int Hmax = 0;
  for (int m = 0; m < a.n_mods; ++m) Hmax = a.H[m] > Hmax ? a.H[m] : Hmax;
  if (lstm_v2_ok(a)) {
    const size_t smem2 = (size_t)FD * NB * (16 * Hmax + FEED_PAD) + 2 * (size_t)NB * ((a.training ? 16 * Hmax + 16 : 0) + 8 * Hmax + 16);
    if (a.training) { MT_TRY(mtrec::set_smem(lstm_fwd_mma2_kernel<true>, smem2)); lstm_fwd_mma2_kernel<true><<<grid, NTH2, smem2, st>>>(a); }
    else { MT_TRY(mtrec::set_smem(lstm_fwd_mma2_kernel<false>, smem2)); lstm_fwd_mma2_kernel<false><<<grid, NTH2, smem2, st>>>(a); }
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  const size_t smem = feed_bytes(4 * Hmax * 4, 1);
  if (a.training) { MT_TRY(mtrec::set_smem(lstm_fwd_mma_kernel<true>, smem)); lstm_fwd_mma_kernel<true><<<grid, NTH + 32, smem, st>>>(a); }
  else { MT_TRY(mtrec::set_smem(lstm_fwd_mma_kernel<false>, smem)); lstm_fwd_mma_kernel<false><<<grid, NTH + 32, smem, st>>>(a); }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_mma_lstm_bwd(const LstmArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + NB - 1) / NB, a.n_mods);
  int Hmax = 0;
  for (int m = 0; m < a.n_mods; ++m) Hmax = a.H[m] > Hmax ? a.H[m] : Hmax;
  if (lstm_v2_ok(a)) {
    RecMaps maps;
    memset(&maps, 0, sizeof(maps));
    const long long Hs = a.Hs;
    for (int m = 0; m < a.n_mods; ++m) {
      const int H = a.H[m], ho = a.hoff[m];
      MT_TRY(make_map_rows(&maps.m[m][0], a.gates + 4 * ho, 4 * Hs - 4 * ho, 4 * Hs, a.T, a.B, a.sb, a.st, 2 * H + 4));
      MT_TRY(make_map_rows(&maps.m[m][1], a.cstar + Hs + ho, Hs - ho, 2 * Hs, a.T, a.B, a.sb, a.st, H + 4));
      MT_TRY(make_map_rows(&maps.m[m][2], a.dcstar + ho, 2 * Hs - ho, 2 * Hs, a.T, a.B, a.sb, a.st, H + 4));
      MT_TRY(make_map_rows(&maps.m[m][3], a.dlast + ho, Hs + a.MEM - ho, Hs + a.MEM, a.T, a.B, a.sb, a.st, H + 4));
    }
    const int ks2 = Hmax <= 48 ? 12 : (Hmax <= 88 ? 22 : 24);
    const size_t smem2 = (size_t)BFD * (2 * NB * (2 * Hmax + 4) * 4 + 4 * NB * (Hmax + 4) * 4) + 2 * (size_t)NB * (ks2 * 32 + 16);
    MT_TRY(mtrec::set_smem(lstm_bwd_mma2_kernel, smem2));
    lstm_bwd_mma2_kernel<<<grid, NTH, smem2, st>>>(a, maps);
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  const size_t smem = feed_bytes(9 * Hmax * 4, 6);
  MT_TRY(mtrec::set_smem(lstm_bwd_mma_kernel, smem));
  lstm_bwd_mma_kernel<<<grid, NTH + 32, smem, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
