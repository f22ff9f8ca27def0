// Tensor-core forward of the step-wise LSTM decoder (SFT/multiTransformer.py:465-483, MFT/multiTransformer.py:357-375) for bf16 mode.
//
// The recurrent matrix of the decoder is [4E, E] = [1024, 256]: 512 KB in bf16 -- too big for one SM's registers or shared memory, which
// is why the FFMA kernel (mt_lstm_head.cu) re-streams it from L2 at every step and spends ~29 us per step.  Here a thread-block CLUSTER
// of 8 CTAs owns a tile of narratives for the whole sequence:
//   * CTA r of the cluster owns hidden units [32 r, 32 r + 32), i.e. 128 gate rows, and keeps that [128, 256] slice of
//     W_ih[:, :E] + W_hh as mma.sync A fragments in REGISTERS (64 per thread; rows permuted so that one 16-row tile holds i / f / g / o of
//     4 units and a single shuffle pair finishes a cell, as in mt_mfn_mma.cu);
//   * h_{t-1} (bf16, [narrative][unit]) is the B operand in shared memory; a step is 16 dependent k-steps of mma.sync.m16n8k16;
//   * the hoisted input projection zx_t (one GEMM over all T*B rows before the kernel) arrives through a ring of bulk async copies
//     issued FD - 1 steps ahead by a producer warp;
//   * the new h_t of the CTA's 32 units is written straight into the shared memory of ALL 8 CTAs (st.shared::cluster), followed by one
//     cluster barrier per step -- the only synchronisation of the recurrence;
//   * the MLP head (Linear -> ReLU -> Linear, per step independent of the recurrence) is hoisted OUT of the loop: one tcgen05 GEMM over
//     all rows of the stored h_t plus a row-wise dot product.
// Step 0 applies W_hh only (o_{-1} = 0 while h_{-1} = dec_h0); from step 1 on o_{t-1} = h_{t-1}, so the two matrices are pre-summed.
#include "mt_mma.cuh"
#include "mt_recurrent.cuh"

namespace {

using namespace mtmma;

constexpr int CL = 8;              // CTAs per cluster
constexpr int EH = 256;            // hidden width this kernel is built for
constexpr int UC = EH / CL;        // hidden units per CTA
constexpr int NWC = UC / 4;        // compute warps (4 units each)
constexpr int LDH = EH + 8;        // row stride (bf16) of the [narrative][unit] operand: conflict-free 32-bit fragment reads
constexpr int FDH = 4;             // stages of the zx ring
constexpr int KS = EH / 16;

struct ClArgs {
  const float* w_ih; const float* w_hh; const float* b_hh; const float* h0; const float* c0;
  float* gates;                    // [M, 4E]: in = zx (enc projection + b_ih); out (training) = post-activation gates
  float* hprev; float* oprev; float* cprev; float* hcur;      // training stash, fp32 [M, E]
  bf16* hall;                      // [M, E] h_t, operand of the hoisted head
  int B, T;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NTL = 8-narrative n-tiles per cluster
template <int NTL, bool TRAIN>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__((NWC + 1) * 32, 1) head_fwd_cluster_kernel(const __grid_constant__ ClArgs a) {
  constexpr int NB = 8 * NTL;
  constexpr int STAGE_FLOATS = NB * 4 * UC;
  extern __shared__ __align__(128) unsigned char dsm[];
  bf16* hS = reinterpret_cast<bf16*>(dsm);                                    // [2][NB][LDH]
  float* zring = reinterpret_cast<float*>(dsm + 2 * NB * LDH * sizeof(bf16));   // [FDH][NB][4][UC]
  __shared__ __align__(8) uint64_t full[FDH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
  const int rank = (int)cluster_rank();
  const int b0 = ((int)blockIdx.x / CL) * NB;
  const int E = EH;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FDH; ++i) mtrec::sbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // h_{-1} = dec_h0 for every narrative: each CTA fills its own copy of the operand
  for (int e = threadIdx.x; e < NB * EH; e += (NWC + 1) * 32) {
    const int n = e / EH, u = e % EH;
    hS[n * LDH + u] = __float2bfloat16(a.h0[u]);
  }
  __syncthreads();
  cluster_sync_all();                       // every CTA's barriers and operand are in place before anybody writes remotely

  if (warp == NWC) {
    // ===== producer: zx rows of this CTA's units, FDH - 1 steps ahead; takes part in the per-step cluster barrier =====
    auto issue = [&](int t) {
      const int slot = t % FDH;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) mtrec::sbar_expect_tx(&full[slot], (uint32_t)(STAGE_FLOATS * sizeof(float)));
      __syncwarp();
      for (int c = lane; c < NB * 4; c += 32) {
        const int n = c >> 2, gate = c & 3;
        const long long row = (long long)min(b0 + n, a.B - 1) * a.T + t;
        mtrec::bulk_g2s(zring + (size_t)slot * STAGE_FLOATS + (n * 4 + gate) * UC, a.gates + row * (4 * E) + gate * E + rank * UC,
                        (uint32_t)(UC * sizeof(float)), &full[slot]);
      }
    };
    for (int t = 0; t < FDH - 1 && t < a.T; ++t) issue(t);
    for (int t = 0; t < a.T; ++t) {
      if (t + FDH - 1 < a.T) issue(t + FDH - 1);      // its slot was last read in step t - 1, i.e. before the previous cluster barrier
      cluster_sync_all();
    }
    return;
  }

  // ===== compute warps: warp w owns units 4 w .. 4 w + 3 of this CTA =====
  const bool lower = gid < 4;                       // rows gid (gate i) and gid + 8 (gate g); the upper half holds f and o
  const int gate0 = lower ? 0 : 1, gate1 = gate0 + 2;
  const int ul = 4 * warp + (gid & 3);              // unit inside the CTA
  const int ug = rank * UC + ul;                    // global hidden unit
  uint32_t A[KS][4];
  auto load_weights = [&](bool with_ih) {
    // tile row r (0..15): gate r >> 2, unit 4 w + (r & 3); this thread holds rows gid and gid + 8
    const int n0 = (gid >> 2) * E + rank * UC + 4 * warp + (gid & 3);
    const int n1 = ((gid + 8) >> 2) * E + rank * UC + 4 * warp + (gid & 3);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c = ks * 16 + 2 * q + 8 * hh;
        float2 v0 = *reinterpret_cast<const float2*>(a.w_hh + (size_t)n0 * E + c);
        float2 v1 = *reinterpret_cast<const float2*>(a.w_hh + (size_t)n1 * E + c);
        if (with_ih) {
          const float2 i0 = *reinterpret_cast<const float2*>(a.w_ih + (size_t)n0 * 2 * E + c);
          const float2 i1 = *reinterpret_cast<const float2*>(a.w_ih + (size_t)n1 * 2 * E + c);
          v0.x += i0.x; v0.y += i0.y; v1.x += i1.x; v1.y += i1.y;
        }
        A[ks][2 * hh] = pack2(v0.x, v0.y);
        A[ks][2 * hh + 1] = pack2(v1.x, v1.y);
      }
    }
  };
  load_weights(false);
  const float bz0 = a.b_hh[gate0 * E + ug], bz1 = a.b_hh[gate1 * E + ug];
  float c[NTL], h[NTL];
#pragma unroll
  for (int i = 0; i < NTL; ++i) { c[i] = a.c0[ug]; h[i] = a.h0[ug]; }
  const uint32_t hs_u32 = mtrec::s_u32(hS);
  uint32_t peer[CL];
#pragma unroll
  for (int r = 0; r < CL; ++r) peer[r] = map_to_cta(hs_u32, (uint32_t)r);

  for (int t = 0; t < a.T; ++t) {
    const bf16* hin = hS + (t & 1) * NB * LDH;
    const uint32_t out_off = (uint32_t)(((t & 1) ^ 1) * NB * LDH * (int)sizeof(bf16));
    // two independent accumulator chains per n-tile (even / odd k-steps): halves the dependent mma chain of the step
    float acc[NTL][4], acb[NTL][4];
#pragma unroll
    for (int i = 0; i < NTL; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = acb[i][0] = acb[i][1] = acb[i][2] = acb[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int i = 0; i < NTL; ++i) {
        const bf16* p = hin + (i * 8 + gid) * LDH + ks * 16 + 2 * q;
        mma16816((ks & 1) ? acb[i] : acc[i], A[ks], *reinterpret_cast<const uint32_t*>(p), *reinterpret_cast<const uint32_t*>(p + 8));
      }
    }
#pragma unroll
    for (int i = 0; i < NTL; ++i) { acc[i][0] += acb[i][0]; acc[i][1] += acb[i][1]; acc[i][2] += acb[i][2]; acc[i][3] += acb[i][3]; }
    mtrec::sbar_wait(&full[t % FDH], (uint32_t)(t / FDH) & 1u);
    const float* zs = zring + (size_t)(t % FDH) * STAGE_FLOATS;
#pragma unroll
    for (int i = 0; i < NTL; ++i) {
      float g[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int n = i * 8 + 2 * q + (v & 1);
        const float z = acc[i][v] + zs[(n * 4 + ((v >> 1) ? gate1 : gate0)) * UC + ul] + ((v >> 1) ? bz1 : bz0);
        g[v] = (lower && (v >> 1)) ? ftanh(z) : fsig(z);
      }
      // swap: the lower thread sends its narrative-(2q+1) pair (i, g), the upper thread its narrative-2q pair (f, o)
      const float s0 = lower ? g[1] : g[0], s1 = lower ? g[3] : g[2];
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
      const float gi = lower ? g[0] : r0, gg = lower ? g[2] : r1, gf = lower ? r0 : g[1], go = lower ? r1 : g[3];
      const float cp = c[i], hp = h[i];
      const float cn = gf * cp + gi * gg;
      const float hn = go * ftanh(cn);
      c[i] = cn; h[i] = hn;
      const int nm = i * 8 + (lower ? 2 * q : 2 * q + 1);        // the narrative this thread finishes
      const __nv_bfloat16 hb = __float2bfloat16(hn);
      const unsigned short hbits = *reinterpret_cast<const unsigned short*>(&hb);
      const uint32_t off = out_off + (uint32_t)((nm * LDH + ug) * (int)sizeof(bf16));
#pragma unroll
      for (int r = 0; r < CL; ++r) st_cluster_u16(peer[r] + off, hbits);
      if (b0 + nm < a.B) {
        const size_t row = (size_t)(b0 + nm) * a.T + t;
        a.hall[row * E + ug] = hb;
        if (TRAIN) {
          float* gr = a.gates + row * (4 * E) + ug;
          gr[0] = gi; gr[E] = gf; gr[2 * E] = gg; gr[3 * E] = go;
          a.hprev[row * E + ug] = hp;
          a.oprev[row * E + ug] = t == 0 ? 0.f : hp;
          a.cprev[row * E + ug] = cp;
          a.hcur[row * E + ug] = hn;
        }
      }
    }
    if (t == 0) load_weights(true);                 // from step 1 on: (W_ih[:, :E] + W_hh) h_{t-1}
    cluster_sync_all();
  }
}

// out[r] = (oh[r, :] . w2 + b2) * mask[r]: one warp per row
__global__ void head_out_kernel(int M, int Hd, const float* __restrict__ oh, const float* __restrict__ w2, const float* __restrict__ b2,
                                const float* __restrict__ mask, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += gridDim.x * (blockDim.x >> 5)) {
    float acc = 0.f;
    for (int j = lane; j < Hd; j += 32) acc = fmaf(oh[(size_t)r * Hd + j], w2[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float y = acc + b2[0];
      if (mask) y *= mask[r];
      out[r] = y;
    }
  }
}

template <int NTL, bool TRAIN>
int launch_cluster(const ClArgs& a, cudaStream_t st) {
  constexpr int NB = 8 * NTL;
  const size_t smem = (size_t)2 * NB * LDH * sizeof(bf16) + (size_t)FDH * NB * 4 * UC * sizeof(float);
  static MtPerDeviceOnce once;
  if (once.first()) MT_CUDA(cudaFuncSetAttribute(head_fwd_cluster_kernel<NTL, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int clusters = (a.B + NB - 1) / NB;
  // algorithmic work of the recurrence: one [4E x E] mat-vec per narrative and step; per (narrative, step) the pre-projected gates come
  // in (4E fp32) and h_t goes out (E bf16), plus the training stash (gates back out and four E-wide fp32 states).  A latency chain of T
  // dependent steps: neither figure is a roof it could reach.
  mt_prof_work(2.0 * a.B * (double)a.T * 4.0 * EH * EH,
               (double)a.B * a.T * (4.0 * EH * 4.0 + EH * 2.0 + (TRAIN ? 4.0 * EH * 4.0 + 4.0 * EH * 4.0 : 0.0)));
  head_fwd_cluster_kernel<NTL, TRAIN><<<clusters * CL, (NWC + 1) * 32, smem, st>>>(a);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

}  // namespace

bool mt_lstm_head_mma_supported(int E, int Hd) { return E == EH && Hd > 0; }

// recurrence (all steps) on the cluster kernel; zx must already sit in `gates`
int mt_lstm_head_mma_recurrence(int B, int T, bool training, const float* w_ih, const float* w_hh, const float* b_hh, const float* h0,
                                const float* c0, float* gates, float* hprev, float* oprev, float* cprev, float* hcur, void* hall,
                                cudaStream_t st) {
  ClArgs a;
  a.w_ih = w_ih; a.w_hh = w_hh; a.b_hh = b_hh; a.h0 = h0; a.c0 = c0;
  a.gates = gates; a.hprev = hprev; a.oprev = oprev; a.cprev = cprev; a.hcur = hcur; a.hall = (bf16*)hall;
  a.B = B; a.T = T;
  // narratives per cluster: as few as keeps all clusters resident at once (8 CTAs each on 148 SMs)
  const int max_clusters = 148 / CL;
  if ((B + 7) / 8 <= max_clusters) return training ? launch_cluster<1, true>(a, st) : launch_cluster<1, false>(a, st);
  if ((B + 15) / 16 <= max_clusters) return training ? launch_cluster<2, true>(a, st) : launch_cluster<2, false>(a, st);
  return training ? launch_cluster<4, true>(a, st) : launch_cluster<4, false>(a, st);
}

int mt_lstm_head_out_run(int M, int Hd, const float* oh, const float* w2, const float* b2, const float* mask, float* out, cudaStream_t st) {
  int grid = (M + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  head_out_kernel<<<grid, 256, 0, st>>>(M, Hd, oh, w2, b2, mask, out);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
