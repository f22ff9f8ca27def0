// Internal GEMM interface shared by every op in the library.
//   C[m,n] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
// A(m,k) = a_kmajor ? A[m*lda + k] : A[k*lda + m];   B(n,k) = b_kmajor ? B[n*ldb + k] : B[k*ldb + n]
// so  forward  y = x W^T      : A = x  (k-major), B = W  (k-major)
//     dgrad    dx = dy W      : A = dy (k-major), B = W  (mn-major: W[N_out,K_in] read as B(n=k_in, k=n_out))
//     wgrad    dW = dy^T x    : A = dy (mn-major), B = x (mn-major), contraction over tokens
// Engines: FFMA SIMT (fp32 mode, and any shape the tensor path does not take) and tcgen05 (bf16 mode).
#pragma once
#include "mt_common.cuh"

struct GemmEpi {
  const float* bias = nullptr;     // [N]
  int act = MT_ACT_NONE;
  DropCfg drop = {0u, 1.0f, 0ull, 0u, nullptr};   // output dropout, element index m*N + n
  const void* gate = nullptr;      // operand-dtype [M, ldg]: out *= (gate > 0 ? gate_scale : 0)   (relu/dropout backward)
  int ldg = 0;
  float gate_scale = 1.0f;
  const float* residual = nullptr; // fp32 [M, ldr], added after activation/dropout
  int ldr = 0;
  const float* rowmask = nullptr;  // fp32 [M], multiplies the final value
  float alpha = 1.0f;
  int accumulate = 0;              // C += result (non-atomic)
  float* colsum = nullptr;         // fp32 [N], ACCUMULATED: column sums of the final output (bias gradient of the layer whose
                                   // pre-activation gradient this GEMM produces); not with split_k
};

struct GemmDesc {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr; int lda = 0; bool a_kmajor = true;
  const void* B = nullptr; int ldb = 0; bool b_kmajor = true;
  void* C = nullptr; int ldc = 0; bool c_f32 = false;
  int split_k = 1;                 // >1: partial sums are atomically added into fp32 C (caller zeroes C)
  int groups = 1;                  // >1 (tcgen05 split-K wgrad form only, both operands mn-major): group g contracts rows [g*K, (g+1)*K) of
  long long c_gstride = 0;         //     A / B into C + g * c_gstride -- the weight gradients of several modality stacks in one launch
  int mgroups = 1;                 // >1 (tcgen05 engine only): the M rows are mgroups equal blocks (modality stacks back to back), block g
  long long b_gstride = 0;         //     multiplies by the weight matrix at B + g * b_gstride elements; no bias / colsum / dropout
  GemmEpi epi;
};

int mt_gemm_run(int dtype, const GemmDesc& g, cudaStream_t st);
int mt_gemm_simt_run(int dtype, const GemmDesc& g, cudaStream_t st);
// returns MT_ERR_UNSUPPORTED when the shape/alignment is outside the tensor-core kernel's envelope
int mt_gemm_tc_run(const GemmDesc& g, cudaStream_t st);
bool mt_gemm_tc_supported(const GemmDesc& g);

// column sums: out[n] (+)= sum_m X[m*ldx + n] * (gate ? ...)   -- bias gradients
int mt_colsum_run(int x_is_bf16, int M, int N, const void* X, int ldx, float* out, int accumulate, cudaStream_t st);

// several column-sum jobs over tensors with the same number of rows in one launch (outputs are ACCUMULATED)
#define MT_COLSUM_MAX_JOBS 16
struct ColsumJob { const void* X; int ldx; int N; float* out; };
int mt_colsum_multi_run(int x_is_bf16, int M, const ColsumJob* jobs, int n_jobs, cudaStream_t st);
