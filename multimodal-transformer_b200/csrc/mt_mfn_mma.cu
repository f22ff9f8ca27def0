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

namespace {

using namespace mtmma;

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
      mtrec::sbar_wait(&full[t % FD], (uint32_t)(t / FD) & 1u);
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
        if (!val[v & 1]) continue;
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
      mtrec::sbar_wait(&full[i % FD], (uint32_t)(i / FD) & 1u);
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
        if (val[v & 1]) {
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
        if (val[v & 1]) dgh_op[(rbase[v & 1] + (long long)t * a.st) * G2 + ff[v >> 1]] = __float2bfloat16(dg);
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
      for (int t = 0; t < a.T; ++t) {
      const bf16* hin = hS[t & 1];
      bf16* hout = hS[(t & 1) ^ 1];
      float acc[LMT][4];
#pragma unroll
      for (int i = 0; i < LMT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < LKS; ++ks) {
        if (ks * 16 >= H) continue;
        uint32_t bb0, bb1;
        frag_b(bb0, bb1, hin, LDH, ks * 16, lane);
#pragma unroll
        for (int i = 0; i < LMT; ++i) mma16816(acc[i], A[i][ks], bb0, bb1);
      }
      mtrec::sbar_wait(&full[t % FD], (uint32_t)(t / FD) & 1u);
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
        if (TRAIN && unit[i] < H) {
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
        c[i] = cn; h[i] = hn;
        if (unit[i] < H) {
          hout[n_mine * LDH + unit[i]] = __float2bfloat16(hn);
          if (val_mine) {
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
      release_slot(&empty[t % FD], lane);
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
      for (int i = 0; i < a.T; ++i) {
      const int t = a.T - 1 - i;
      mtrec::sbar_wait(&full[i % FD], (uint32_t)(i / FD) & 1u);
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
          if (val[v & 1]) {
            bf16* zo_g = dz_op + (rbase[v & 1] + (long long)t * a.st) * (4 * Hs) + 4 * hoff + u;
            zo_g[0] = __float2bfloat16(zi); zo_g[H] = __float2bfloat16(zf); zo_g[2 * H] = __float2bfloat16(zg); zo_g[3 * H] = __float2bfloat16(zo);
          }
        }
      }
      release_slot(&empty[i % FD], lane);
      consumers_sync();
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
      }
    }
  }
}

}  // namespace

// dynamic shared memory of a feed ring whose per-narrative spans are `spans` bytes in total over `nseg` segments
static size_t feed_bytes(int spans, int nseg) { return (size_t)FD * NB * (size_t)(spans + nseg * FEED_PAD); }

bool mt_mfn_mma_mem_supported(const MemArgs& a) { return a.MEM == 128 && a.G == 64 && (a.Hs % 4) == 0; }

int mt_mfn_mma_mem_fwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
  const size_t smem = feed_bytes(2 * 128 * 4, 2);
  if (a.training) { MT_TRY(mtrec::set_smem(mem_fwd_mma_kernel<true>, smem)); mem_fwd_mma_kernel<true><<<grid, NTH + 32, smem, st>>>(a); }
  else { MT_TRY(mtrec::set_smem(mem_fwd_mma_kernel<false>, smem)); mem_fwd_mma_kernel<false><<<grid, NTH + 32, smem, st>>>(a); }
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_mma_mem_bwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
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

int mt_mfn_mma_lstm_fwd(const LstmArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + NB - 1) / NB, a.n_mods);
  int Hmax = 0;
  for (int m = 0; m < a.n_mods; ++m) Hmax = a.H[m] > Hmax ? a.H[m] : Hmax;
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
  const size_t smem = feed_bytes(9 * Hmax * 4, 6);
  MT_TRY(mtrec::set_smem(lstm_bwd_mma_kernel, smem));
  lstm_bwd_mma_kernel<<<grid, NTH + 32, smem, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
