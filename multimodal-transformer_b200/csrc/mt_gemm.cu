// GEMM dispatcher: picks the tcgen05 engine (bf16 operands, shapes inside its envelope) or the FFMA engine.
#include "mt_gemm.cuh"

#ifndef MT_NO_TC
#define MT_HAVE_TC 1
#else
#define MT_HAVE_TC 0
#endif

static int g_force_simt = 0;
void mt_gemm_tc_set_trace(unsigned long long* p);      // mt_gemm_tc.cu
int mt_gemm_tc_set_mode(int mode);

int mt_gemm_run(int dtype, const GemmDesc& g, cudaStream_t st) {
  if (g_mt_prof_on) {
    const double es = dtype == MT_BF16 ? 2.0 : 4.0, cs = (g.c_f32 || dtype == MT_F32) ? 4.0 : 2.0;
    char tag[40];
    snprintf(tag, sizeof(tag), "m%d n%d k%d %c%c%s", g.M, g.N, g.K, g.a_kmajor ? 'K' : 'M', g.b_kmajor ? 'K' : 'M', g.split_k > 1 ? " splitk" : "");
    mt_prof_tag(tag);
    mt_prof_work(2.0 * g.M * (double)g.N * g.K, ((double)g.M * g.K + (double)g.N * g.K) * es + (double)g.M * g.N * cs +
                                                    (g.epi.residual ? 4.0 * g.M * g.N : 0.0) + (g.epi.gate ? es * g.M * g.N : 0.0));
  }
#if MT_HAVE_TC
  if (dtype == MT_BF16 && !g_force_simt && mt_gemm_tc_supported(g)) return mt_gemm_tc_run(g, st);
#endif
  if (g.groups > 1 || g.mgroups > 1) return MT_ERR_UNSUPPORTED;      // grouped forms exist on the tcgen05 engine only: the caller launches per group
  if (g.epi.colsum) {        // the FFMA engine has no fused column sum: run it, then one column-sum pass over C
    if (g.split_k > 1) return MT_ERR_ARG;
    GemmDesc g2 = g;
    g2.epi.colsum = nullptr;
    MT_TRY(mt_gemm_simt_run(dtype, g2, st));
    return mt_colsum_run(dtype == MT_BF16 && !g.c_f32, g.M, g.N, g.C, g.ldc, g.epi.colsum, 1, st);
  }
  return mt_gemm_simt_run(dtype, g, st);
}

extern "C" {

int mt_gemm(int dtype, int M, int N, int K, const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, void* C,
            int ldc, int c_f32, const float* bias, int act, int split_k_atomic, void* stream) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_kmajor = a_kmajor != 0;
  g.B = B; g.ldb = ldb; g.b_kmajor = b_kmajor != 0;
  g.C = C; g.ldc = ldc; g.c_f32 = (c_f32 != 0) || dtype == MT_F32;
  g.split_k = split_k_atomic > 1 ? split_k_atomic : 1;
  g.epi.bias = bias; g.epi.act = act;
  return mt_gemm_run(dtype, g, (cudaStream_t)stream);
}

int mt_gemm_engine(int dtype, int M, int N, int K, int a_kmajor, int b_kmajor) {
#if MT_HAVE_TC
  if (dtype != MT_BF16 || g_force_simt) return 0;
  GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.a_kmajor = a_kmajor != 0; g.b_kmajor = b_kmajor != 0;
  g.lda = a_kmajor ? K : M; g.ldb = b_kmajor ? K : N; g.ldc = N;
  g.A = g.B = (const void*)0x1000; g.C = (void*)0x1000;
  return mt_gemm_tc_supported(g) ? 1 : 0;
#else
  return 0;
#endif
}

/* debug hook: CTA 0 of every tcgen05 GEMM writes per-tile clock64 stamps (8 words per tile, 64 tiles) into `dev_buf`; NULL = off */
int mt_gemm_tc_mode(int mode) { return mt_gemm_tc_set_mode(mode); }
int mt_gemm_debug_trace(void* dev_buf) { mt_gemm_tc_set_trace((unsigned long long*)dev_buf); return MT_OK; }

/* test hook: route every GEMM through the FFMA engine (used to A/B the tensor-core engine) */
int mt_gemm_force_simt(int on) { int old = g_force_simt; g_force_simt = on; return old; }

}  // extern "C"
