// Building blocks of the persistent recurrence kernels (MFN, LSTM decoder): a CTA owns BT narratives, keeps
// activations feature-major [feature][BT] in shared memory and applies small dense layers whose weights stream
// from L2 with coalesced loads.
#pragma once
#include "mt_ops.cuh"

namespace mtrec {

constexpr int BT = 4;            // narratives per CTA
constexpr int NTHREADS = 256;

// ---- in-CTA dense layer ------------------------------------------------------------------------------
// out[n][b] = sum_k Wt[k*ldw + n] * xs[k][b],  n in [0,N), b in [0,BT).   Wt is "K-major over rows": consecutive n are
// contiguous, so a warp's weight loads are coalesced.  When N is small the K range is split over thread groups and
// the partial sums are combined through `part` (needs (NTHREADS) * BT floats).  epi(n, acc) runs exactly once per n.
// Caller must __syncthreads() before reading anything epi wrote.
template <typename WT, typename Epi>
__device__ __forceinline__ void dense(const WT* __restrict__ Wt, int ldw, int K, int N, const float* __restrict__ xs, float* part, Epi epi) {
  const int tid = threadIdx.x;
  int npad = (N + 31) & ~31;
  int P = 1;
  while (P * 2 * npad <= NTHREADS) P *= 2;
  if (P == 1) {
    for (int n = tid; n < N; n += NTHREADS) {
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[b] = 0.f;
      const WT* w = Wt + n;
#pragma unroll 8
      for (int k = 0; k < K; ++k) {
        float wv = to_f(w[(size_t)k * ldw]);
        float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
        acc[0] = fmaf(wv, x.x, acc[0]); acc[1] = fmaf(wv, x.y, acc[1]);
        acc[2] = fmaf(wv, x.z, acc[2]); acc[3] = fmaf(wv, x.w, acc[3]);
      }
      epi(n, acc);
    }
  } else {
    const int p = tid / npad, n = tid % npad;
    const int kc = (K + P - 1) / P;
    const int k0 = p * kc, k1 = min(K, k0 + kc);
    float acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.f;
    if (p < P && n < N) {
      const WT* w = Wt + n;
#pragma unroll 8
      for (int k = k0; k < k1; ++k) {
        float wv = to_f(w[(size_t)k * ldw]);
        float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
        acc[0] = fmaf(wv, x.x, acc[0]); acc[1] = fmaf(wv, x.y, acc[1]);
        acc[2] = fmaf(wv, x.z, acc[2]); acc[3] = fmaf(wv, x.w, acc[3]);
      }
    }
    if (p > 0 && p < P) *reinterpret_cast<float4*>(part + (size_t)tid * BT) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (p == 0 && n < N) {
      for (int q = 1; q < P; ++q) {
        float4 o = *reinterpret_cast<const float4*>(part + (size_t)(q * npad + n) * BT);
        acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
      }
      epi(n, acc);
    }
    __syncthreads();      // `part` may be rewritten by the next dense() call
  }
}

// copy a [w][BT] shared buffer to BT global rows (row r of sample b at base + row_b*w)
__device__ __forceinline__ void stash_rows(float* __restrict__ g, int w, const float* __restrict__ s, const long long* rows, int nb) {
  if (!g) return;
  for (int e = threadIdx.x; e < nb * w; e += NTHREADS) {
    int b = e / w, f = e % w;
    g[rows[b] * w + f] = s[f * BT + b];
  }
}
__device__ __forceinline__ void load_rows(float* __restrict__ s, int w, const float* __restrict__ g, const long long* rows, int nb) {
  for (int e = threadIdx.x; e < BT * w; e += NTHREADS) {
    int b = e / w, f = e % w;
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
