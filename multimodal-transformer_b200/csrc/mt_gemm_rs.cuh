// Row-stream GEMM engine (mt_gemm_rs.cu): weight-resident, grouped, TMA in / TMA out.  See the header of mt_gemm_rs.cu.
//   C[g*Mg + m, n] = epilogue( sum_k A[g*Mg + m, k] * B_g(n, k) )        g = 0 .. G-1 (one group = one modality stack)
// A: bf16 [G*Mg, K] K-major, contiguous across groups (lda elements per row).  B_g: that group's weight matrix, bf16, either
// K-major [N, K] (forward projections) or MN-major [K, N] (dgrads: the forward weight read transposed in place).
// C / residual / gate are contiguous across groups like A.
#pragma once
#include "mt_gemm.cuh"

#define MT_RS_MAX_GROUPS 4

struct RsDesc {
  int G = 1;                 // groups
  int Mg = 0;                // rows per group (G > 1: a multiple of 128)
  int N = 0, K = 0;          // per-group problem: N % 64 == 0, K % 64 == 0
  const void* A = nullptr; int lda = 0;
  const void* B[MT_RS_MAX_GROUPS] = {}; int ldb = 0; bool b_kmajor = true;
  void* C = nullptr; int ldc = 0; bool c_f32 = false;
  const float* bias[MT_RS_MAX_GROUPS] = {};         // [N] per group (all or none)
  int act = MT_ACT_NONE;                            // NONE | RELU
  DropCfg drop[MT_RS_MAX_GROUPS] = {};              // output dropout of group g, element index m*N + n with m LOCAL to the group
  const void* gate = nullptr; int ldg = 0; float gate_scale = 1.0f;   // bf16 [G*Mg, ldg]: out = gate > 0 ? out * gate_scale : 0
  const float* residual = nullptr; int ldr = 0;     // fp32 [G*Mg, ldr], added last
  float* colsum[MT_RS_MAX_GROUPS] = {};             // fp32 [N] per group, ACCUMULATED column sums of the final output
  // fused LayerNorm of the (row-complete, N == 256) fp32 output: second output ln_out = a_2 * (c - mean) / (std + eps) + b_2 in bf16
  void* ln_out = nullptr; int ld_ln = 0;
  const float* ln_a[MT_RS_MAX_GROUPS] = {};
  const float* ln_b[MT_RS_MAX_GROUPS] = {};
  float ln_eps = 1e-6f;
  // attention-backward helper (input gradient of the output projection, N == d): D[b, head, q] = sum over the head's 32 columns of
  // C (fp32, before rounding) * attd_src (the attention output, bf16 [G*Mg, N]) -> attd_aux[((b * h + head) * 4 + 1) * 128 + q]
  // with b = row / attd_T (global narrative index), q = row % attd_T, h = N / 32: row 1 of mt_attention_tc.cu's per-query scalars
  const void* attd_src = nullptr; int attd_ld = 0;
  float* attd_aux = nullptr; int attd_T = 0;
  // optional: with attd_lse (fp32 [G*B, h, attd_T], natural log) the same epilogue also writes rows 0 / 2 / 3 of the scalars -- lse * log2e,
  // score scale * log2e and score-gradient scale (both 0 where attd_mask[b LOCAL to the group, q] == 0; attd_mask may be null) --
  // so that the attention backward needs no preparation launch at all
  const float* attd_lse = nullptr; const float* attd_mask = nullptr; float attd_scale = 0.f;
};

bool mt_gemm_rs_supported(const RsDesc& d);
int mt_gemm_rs_run(const RsDesc& d, cudaStream_t st);
