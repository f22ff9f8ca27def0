// Building blocks of the persistent recurrence kernels (MFN, LSTM decoder): a CTA owns BT narratives, keeps
// activations feature-major [feature][BT] in shared memory and applies small dense layers whose weights stream
// from L2 with coalesced 16-byte loads.
#pragma once
#include "mt_ops.cuh"

namespace mtrec {

constexpr int BT = 4;            // narratives per CTA
constexpr int NTHREADS = 512;
constexpr int PART_FLOATS = NTHREADS * 8 * BT;   // K-split partial sums: P * N * BT <= NTHREADS * VN * BT

template <typename WT> struct WVec;
template <> struct WVec<float> {
  static constexpr int VN = 4;
  __device__ static __forceinline__ void load(const float* p, float* w) {
    float4 v = *reinterpret_cast<const float4*>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
};
template <> struct WVec<bf16> {
  static constexpr int VN = 8;
  __device__ static __forceinline__ void load(const bf16* p, float* w) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
    float2 c = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.z));
    float2 d = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.w));
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y; w[6] = d.x; w[7] = d.y;
  }
};

// ---- in-CTA dense layer ------------------------------------------------------------------------------
// out[n][b] = sum_k Wt[k*ldw + n] * xs[k][b],  n in [0,N), b in [0,BT).   Consecutive n are contiguous in Wt.
// Each thread owns VN consecutive outputs (one 16-byte weight load per k feeds VN*BT FMAs) over one of P slices of
// the K range, so the dependent-load chain per layer is K/P long; the P partial sums meet in `part`
// (PART_FLOATS floats of shared memory) and epi(n, acc[BT]) runs exactly once per n.
// All threads of the CTA must call this; on return every epi() write is visible to the whole CTA.
template <typename WT, typename Epi>
__device__ __forceinline__ void dense(const WT* __restrict__ Wt, int ldw, int K, int N, const float* __restrict__ xs, float* part, Epi epi) {
  constexpr int VN = WVec<WT>::VN;
  const int tid = threadIdx.x;
  const bool vec_ok = (N % VN == 0) && (ldw % VN == 0) && ((reinterpret_cast<uintptr_t>(Wt) & 15) == 0) && (N / VN <= NTHREADS);
  if (vec_ok) {
    const int NG = N / VN;
    int P = NTHREADS / NG;
    if (P > 16) P = 16;          // deeper splits only lengthen the serial partial-sum pass
    if (P > K) P = K;
    const int kc = (K + P - 1) / P;
    const int p = tid / NG, g = tid - p * NG;
    float acc[VN][BT];
#pragma unroll
    for (int i = 0; i < VN; ++i)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[i][b] = 0.f;
    if (p < P) {
      const int k0 = p * kc, k1 = min(K, k0 + kc);
      const WT* w = Wt + (size_t)k0 * ldw + g * VN;
#pragma unroll 4
      for (int k = k0; k < k1; ++k, w += ldw) {
        float wv[VN];
        WVec<WT>::load(w, wv);
        const float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          acc[i][0] = fmaf(wv[i], x.x, acc[i][0]); acc[i][1] = fmaf(wv[i], x.y, acc[i][1]);
          acc[i][2] = fmaf(wv[i], x.z, acc[i][2]); acc[i][3] = fmaf(wv[i], x.w, acc[i][3]);
        }
      }
      if (P > 1) {
#pragma unroll
        for (int i = 0; i < VN; ++i)
          *reinterpret_cast<float4*>(part + ((size_t)p * N + g * VN + i) * BT) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      } else {
#pragma unroll
        for (int i = 0; i < VN; ++i) epi(g * VN + i, acc[i]);
      }
    }
    if (P > 1) {
      __syncthreads();
      for (int n = tid; n < N; n += NTHREADS) {
        float4 s = *reinterpret_cast<const float4*>(part + (size_t)n * BT);
        for (int q = 1; q < P; ++q) {
          const float4 o = *reinterpret_cast<const float4*>(part + ((size_t)q * N + n) * BT);
          s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        }
        float a4[BT] = {s.x, s.y, s.z, s.w};
        epi(n, a4);
      }
    }
    __syncthreads();
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
      const float wv = to_f(w[(size_t)k * ldw]);
      const float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
      acc[0] = fmaf(wv, x.x, acc[0]); acc[1] = fmaf(wv, x.y, acc[1]);
      acc[2] = fmaf(wv, x.z, acc[2]); acc[3] = fmaf(wv, x.w, acc[3]);
    }
    epi(n, acc);
  }
  __syncthreads();
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

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return MT_ERR_UNSUPPORTED;
  MT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MT_OK;
}

}  // namespace mtrec
