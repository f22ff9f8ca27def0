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
//   * per-step inputs are prefetched DEPTH steps ahead into registers, outputs leave with plain global stores.
// One layer = <= 8 dependent mma per warp + one CTA barrier.
#include "mt_mfn.cuh"
#include "mt_mma.cuh"

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
constexpr int DEPTH = 4;         // steps of input prefetch held in registers

template <bool TRAIN>
__global__ void __launch_bounds__(NTH, 1) mem_fwd_mma_kernel(const __grid_constant__ MemArgs a) {
  __shared__ __align__(16) bf16 memS[NB * LDK];
  __shared__ __align__(16) bf16 ghS[NB * LDK];
  const int MEM = 128, G = 64, G2 = 128, M2 = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const int f0 = warp * 16 + gid, f1 = f0 + 8;
  const int ldw1 = 2 * a.Hs + MEM;
  // ---- weights -> A fragments ----
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
  const int jg0 = f0 & 63, jg1 = f1 & 63;           // index inside the gate's own [T,B,G] dropout tensor
  // ---- narratives of this thread ----
  const int nn[2] = {2 * q, 2 * q + 1};
  bool val[2]; long long rbase[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) { val[p] = b0 + nn[p] < a.B; rbase[p] = (long long)min(b0 + nn[p], a.B - 1) * a.sb; }
  const int ff[2] = {f0, f1};
  for (int e = threadIdx.x; e < NB * LDK; e += NTH) { memS[e] = __float2bfloat16(0.f); ghS[e] = __float2bfloat16(0.f); }
  float mem[4] = {0.f, 0.f, 0.f, 0.f};
  float gp[DEPTH][4], ch[DEPTH][4];
  auto fetch = [&](int t, float* g, float* c) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long row = rbase[v & 1] + (long long)t * a.st;
      g[v] = a.gpre[row * G2 + ff[v >> 1]];
      c[v] = a.chat[row * MEM + ff[v >> 1]];
    }
  };
#pragma unroll
  for (int i = 0; i < DEPTH; ++i) if (i < a.T) fetch(i, gp[i], ch[i]);
  bf16* gh_op = reinterpret_cast<bf16*>(a.gh_op);
  bf16* memprev_op = reinterpret_cast<bf16*>(a.memprev_op);
  bf16* last_op = reinterpret_cast<bf16*>(a.last_op);
  const int LW = a.Hs + MEM;
  __syncthreads();

  for (int t0 = 0; t0 < a.T; t0 += DEPTH) {
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) {
      const int t = t0 + s;
      if (t >= a.T) break;
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
        float x = fmaxf(acc[0][v] + acc[1][v] + gp[s][v], 0.f);
        // element index of the gate's [T,B,G] tensor (oracle/mt_oracle.py:_drop_t)
        x *= mt_drop_factor(drop, ((uint64_t)t * a.B + (uint64_t)(b0 + nn[v & 1])) * (uint64_t)G + (uint64_t)((v >> 1) ? jg1 : jg0));
        gh[v] = x;
        ghS[nn[v & 1] * LDK + ff[v >> 1]] = __float2bfloat16(x);
      }
      __syncthreads();
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
        mem[v] = g1[v] * mp[v] + g2[v] * ch[s][v];
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
      if (t + DEPTH < a.T) fetch(t + DEPTH, gp[s], ch[s]);
      __syncthreads();
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
constexpr int BDEPTH = 2;

__global__ void __launch_bounds__(NTH, 1) mem_bwd_mma_kernel(const __grid_constant__ MemArgs a) {
  __shared__ __align__(16) bf16 dz1S[NB * LDK];
  __shared__ __align__(16) bf16 dz2S[NB * LDK];
  __shared__ __align__(16) bf16 dghS[NB * LDK];
  const int MEM = 128, G = 64, G2 = 128, M2 = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
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
  struct In { float dl[4], g1[4], g2[4], ch[4], mp[4], gh[4]; };
  In in[BDEPTH];
  auto fetch = [&](int t, In& x) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long row = rbase[v & 1] + (long long)t * a.st;
      const int f = ff[v >> 1];
      x.dl[v] = a.dlast[row * LW + a.Hs + f];
      x.g1[v] = a.gm[row * M2 + f];
      x.g2[v] = a.gm[row * M2 + MEM + f];
      x.ch[v] = a.chat[row * MEM + f];
      x.mp[v] = bf(memprev_op + row * MEM + f);
      x.gh[v] = bf(gh_op + row * G2 + f);
    }
  };
#pragma unroll
  for (int i = 0; i < BDEPTH; ++i) if (i < a.T) fetch(a.T - 1 - i, in[i]);

  for (int i0 = 0; i0 < a.T; i0 += BDEPTH) {
#pragma unroll
    for (int s = 0; s < BDEPTH; ++s) {
      const int i = i0 + s;
      if (i >= a.T) break;
      const int t = a.T - 1 - i;
      const In& x = in[s];
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
      if (i + BDEPTH < a.T) fetch(t - BDEPTH, in[s]);
      __syncthreads();
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
      __syncthreads();
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
constexpr int LDEPTH = 3;

template <bool TRAIN>
__global__ void __launch_bounds__(NTH, 1) lstm_fwd_mma_kernel(const __grid_constant__ LstmArgs a) {
  __shared__ __align__(16) bf16 hS[2][NB * LDH];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
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
  float zx[LDEPTH][LMT][4];
  auto fetch = [&](int t, float (*z)[4]) {
#pragma unroll
    for (int i = 0; i < LMT; ++i) {
      if (unit[i] >= H) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const long long row = rbase[v & 1] + (long long)t * a.st;
        z[i][v] = a.gates[row * (4 * Hs) + 4 * hoff + ((v >> 1) ? gate1 : gate0) * H + unit[i]];
      }
    }
  };
#pragma unroll
  for (int i = 0; i < LDEPTH; ++i) if (i < a.T) fetch(i, zx[i]);
  bf16* cstar_op = reinterpret_cast<bf16*>(a.cstar_op);
  bf16* last_op = reinterpret_cast<bf16*>(a.last_op);
  bf16* hprev_op = reinterpret_cast<bf16*>(a.hprev_op);
  const int LW = Hs + a.MEM;
  __syncthreads();

  for (int t0 = 0; t0 < a.T; t0 += LDEPTH) {
#pragma unroll
    for (int s = 0; s < LDEPTH; ++s) {
      const int t = t0 + s;
      if (t >= a.T) break;
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
#pragma unroll
      for (int i = 0; i < LMT; ++i) {
        if (4 * (warp + NW * i) >= H) continue;                  // warp-uniform: this tile does not exist for this modality
        float g[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float z = acc[i][v] + zx[s][i][v] + bz[i][v >> 1];
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
      if (t + LDEPTH < a.T) fetch(t + LDEPTH, zx[s]);
      __syncthreads();
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
constexpr int LBDEPTH = 2;

__global__ void __launch_bounds__(NTH, 1) lstm_bwd_mma_kernel(const __grid_constant__ LstmArgs a) {
  __shared__ __align__(16) bf16 dzS[2][NB * LDZ];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], Hs = a.Hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * NB;
  const bool active = warp * 16 < H;                 // warp-uniform
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
  struct In { float gi[4], gf[4], gg[4], go[4], cp[4], cn[4], dcp[4], dcn[4], dhd[4]; };
  In in[LBDEPTH];
  auto fetch = [&](int t, In& x) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int u = uu[v >> 1];
      if (!active || u >= H) continue;
      const long long row = rbase[v & 1] + (long long)t * a.st;
      const float* gt = a.gates + row * (4 * Hs) + 4 * hoff + u;
      x.gi[v] = gt[0]; x.gf[v] = gt[H]; x.gg[v] = gt[2 * H]; x.go[v] = gt[3 * H];
      x.cp[v] = a.cstar[row * (2 * Hs) + hoff + u];
      x.cn[v] = a.cstar[row * (2 * Hs) + Hs + hoff + u];
      x.dcp[v] = a.dcstar[row * (2 * Hs) + hoff + u];
      x.dcn[v] = a.dcstar[row * (2 * Hs) + Hs + hoff + u];
      x.dhd[v] = a.dlast[row * LW + hoff + u];
    }
  };
#pragma unroll
  for (int i = 0; i < LBDEPTH; ++i) if (i < a.T) fetch(a.T - 1 - i, in[i]);
  __syncthreads();

  for (int i0 = 0; i0 < a.T; i0 += LBDEPTH) {
#pragma unroll
    for (int s = 0; s < LBDEPTH; ++s) {
      const int i = i0 + s;
      if (i >= a.T) break;
      const int t = a.T - 1 - i;
      bf16* dzo = dzS[i & 1];
      if (active) {
        const In& x = in[s];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int u = uu[v >> 1];
          if (u >= H) continue;
          const float tc = ftanh(x.cn[v]);
          const float dhv = dh[v] + x.dhd[v];
          const float dcv = dc[v] + x.dcn[v] + dhv * x.go[v] * (1.f - tc * tc);
          const float zi = dcv * x.gg[v] * x.gi[v] * (1.f - x.gi[v]);
          const float zf = dcv * x.cp[v] * x.gf[v] * (1.f - x.gf[v]);
          const float zg = dcv * x.gi[v] * (1.f - x.gg[v] * x.gg[v]);
          const float zo = dhv * tc * x.go[v] * (1.f - x.go[v]);
          dc[v] = x.dcp[v] + dcv * x.gf[v];                          // gradient wrt c_{t-1}
          bf16* zr = dzo + nn[v & 1] * LDZ + u;
          zr[0] = __float2bfloat16(zi); zr[96] = __float2bfloat16(zf); zr[192] = __float2bfloat16(zg); zr[288] = __float2bfloat16(zo);
          if (val[v & 1]) {
            bf16* zo_g = dz_op + (rbase[v & 1] + (long long)t * a.st) * (4 * Hs) + 4 * hoff + u;
            zo_g[0] = __float2bfloat16(zi); zo_g[H] = __float2bfloat16(zf); zo_g[2 * H] = __float2bfloat16(zg); zo_g[3 * H] = __float2bfloat16(zo);
          }
        }
        if (i + LBDEPTH < a.T) fetch(t - LBDEPTH, in[s]);
      }
      __syncthreads();
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

}  // namespace

bool mt_mfn_mma_mem_supported(const MemArgs& a) { return a.MEM == 128 && a.G == 64 && (a.Hs % 2) == 0; }

int mt_mfn_mma_mem_fwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
  if (a.training) mem_fwd_mma_kernel<true><<<grid, NTH, 0, st>>>(a);
  else mem_fwd_mma_kernel<false><<<grid, NTH, 0, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_mma_mem_bwd(const MemArgs& a, cudaStream_t st) {
  const int grid = (a.B + NB - 1) / NB;
  mem_bwd_mma_kernel<<<grid, NTH, 0, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

bool mt_mfn_mma_lstm_supported(const LstmArgs& a) {
  for (int m = 0; m < a.n_mods; ++m)
    if (a.H[m] > 96 || a.H[m] % 4 != 0) return false;
  return true;
}

int mt_mfn_mma_lstm_fwd(const LstmArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + NB - 1) / NB, a.n_mods);
  if (a.training) lstm_fwd_mma_kernel<true><<<grid, NTH, 0, st>>>(a);
  else lstm_fwd_mma_kernel<false><<<grid, NTH, 0, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_mma_lstm_bwd(const LstmArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + NB - 1) / NB, a.n_mods);
  lstm_bwd_mma_kernel<<<grid, NTH, 0, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
