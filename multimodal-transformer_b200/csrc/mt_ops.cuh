// Internal (non-ABI) entry points shared between the translation units of libmt_b200.
#pragma once
#include "mt_gemm.cuh"
#include "mt_dropbits.cuh"

// ---- workspace carving: the same sequence of take() calls yields the same offsets in fwd and bwd ------
struct WsCarver {
  char* base;
  size_t off = 0;
  explicit WsCarver(void* p) : base(reinterpret_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = mt_align_up(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return r;
  }
  void* take_bytes(size_t nbytes) { return take<char>(nbytes); }
  size_t total() const { return mt_align_up(off, 256); }
};

static inline size_t mt_esize(int dtype) { return dtype == MT_BF16 ? 2 : 4; }

// ---- element-wise / row-wise kernels (mt_elementwise.cu) ---------------------------------------------
// G > 1: grouped launch over G modality stacks -- x / y are [G*M, d] (rows of group g at g*M), a / b of group g at a + g*pstride floats
#define MT_LN_MAX_GROUPS 4
struct DropGroups { DropCfg d[MT_LN_MAX_GROUPS]; };
// draw != nullptr: the same launch also draws the keep bits of an attention dropout (mt_dropbits.cuh) -- the hash work of the
// ALU-bound draw runs under the memory latency of the HBM-bound norm instead of as a kernel of its own
int mt_ln_fwd_run(int M, int d, const float* x, const float* a, const float* b, float eps, void* y, bool y_bf16, cudaStream_t st, int G = 1,
                  size_t pstride = 0, const MtBitsJob* draw = nullptr);
// optional second output of the LayerNorm backward: out = dx * dropout_factor(drop, row*d + col) in the operand dtype
// (same dtype as dy) and dbias[d] += colsum(out) -- what the sublayer below needs first (see ln_bwd_kernel)
struct LnBwdNext { void* out; float* dbias; DropCfg drop; };
// G > 1: grouped like mt_ln_fwd_run (dy / dres / dx / nx->out rows of group g at g*M; a / da / db / nx->dbias at + g*pstride;
// nx_drops[g] = that group's dropout stream, element index local to the group)
// gmode: dtype of the residual-stream gradient -- 0: dres / dx fp32; 1: both bf16; 2: dres bf16, dx fp32 (1 and 2 with bf16 dy only)
int mt_ln_bwd_run(int M, int d, const float* x, const float* a, float eps, const void* dy, bool dy_bf16, const void* dres,
                  void* dx, float* da, float* db, cudaStream_t st, const LnBwdNext* nx = nullptr, int G = 1, size_t pstride = 0,
                  const DropCfg* nx_drops = nullptr, int gmode = 0);
// out[M,N] (bf16 or f32) = g[M,N] (f32) * dropout_factor(site, m*N+n)      (gradient through an output dropout)
int mt_drop_grad_run(int M, int N, const float* g, void* out, bool out_bf16, DropCfg drop, cudaStream_t st);
// 2-D cast with zero padding / optional input dropout: dst[r, c] = c < cols ? src[r*lds + c] * drop(r*cols + c) : 0
int mt_cast2d_run(const void* src, bool src_bf16, int lds, void* dst, bool dst_bf16, int ldd, int rows, int cols, DropCfg drop,
                  cudaStream_t st, int dcols = -1);
// dz = dy * act'(y) * rowmask   (y = post-activation output)
// db + db_done given: the bias gradient db[n] = sum_m dz[m][n] is fused into the pass where the shape allows (N % 4 == 0, 16-byte aligned
// operands) and *db_done says whether it was
int mt_act_bwd_run(int M, int N, const void* dy, bool dy_bf16, const void* y, bool y_bf16, int act, const float* rowmask, void* dz,
                   bool dz_bf16, cudaStream_t st, float* db = nullptr, bool* db_done = nullptr);
// batched transposes for the MFN forward weight pack: dst[c*ldd + r] = src[r*C + c]
struct TransposeJob { const float* src; void* dst; int R, C, ldd; };
int mt_transpose_pack_run(const TransposeJob* jobs, int n_jobs, bool dst_bf16, cudaStream_t st);

// ---- tensor-core attention engine for bf16 (mt_attention_mma.cu) --------------------------------------------
bool mt_attn_mma_supported(int B, int T, int d, int h);
// klen (optional, int [B]): ragged INFERENCE -- narrative b only has its first klen[b] windows, keys beyond them are excluded
int mt_attn_mma_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                        const int* klen = nullptr);
// dbias (optional, fp32 [3d], ACCUMULATED): column sums of dqkv = the QKV projection's bias gradient; *dbias_done tells
// whether the kernel that ran produced it (otherwise the caller runs a column-sum pass)
int mt_attn_mma_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                        void* dqkv, DropCfg drop, cudaStream_t st, float* dbias = nullptr, bool* dbias_done = nullptr);

// ---- whole-head-per-CTA attention for T <= 128 (mt_attention_t128.cu), bf16 -----------------------------------
bool mt_attn128_supported(int B, int T, int d, int h);
int mt_attn128_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                       const int* klen = nullptr);
int mt_attn128_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                       void* dqkv, DropCfg drop, cudaStream_t st, float* dbias = nullptr);

// ---- tcgen05 / TMEM attention for T <= 128, 32-wide heads, even head count (mt_attention_tc.cu), bf16 ---------------------
extern int g_mt_attn_no_tc;
bool mt_attn_force_tiled_on();
bool mt_attn_tc_supported(int B, int T, int d, int h);
// G > 1: G modality stacks back to back (B narratives each; qkv / out / lse hold G*B narratives, mask / klen are shared), drops[g] = the
// dropout stream of group g (pair indices local to the group)
int mt_attn_tc_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st,
                       const int* klen = nullptr, int G = 1, const DropCfg* drops = nullptr, const uint32_t* dbits = nullptr);
// keep bits of the probability dropout of one launch (G * B * h * 512 words), drawn once and read by both the forward and the backward kernel
size_t mt_attn_tc_dropbits_words(int G, int B, int h);
int mt_attn_tc_dropbits_run(int G, int B, int T, int h, const DropCfg* drops, uint32_t* bits, cudaStream_t st);
// the same draw as a job description for a kernel that draws on the side (mt_ln_fwd_run)
int mt_attn_tc_dropbits_job(int G, int B, int T, int h, const DropCfg* drops, uint32_t* bits, MtBitsJob* job);
// aux: fp32 workspace of mt_attn_bwd_ws_floats(B, T, h) floats (per-query scalars); dbias as in mt_attn_mma_bwd_run (h <= 8)
// G > 1 as in mt_attn_tc_fwd_run; aux holds G * mt_attn_bwd_ws_floats(B, T, h) floats, dbias of group g at dbias + g * dbias_gstride
int mt_attn_tc_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                       void* dqkv, DropCfg drop, float* dbias, float* aux, cudaStream_t st, int G = 1, const DropCfg* drops = nullptr,
                       size_t dbias_gstride = 0, int d_ready = 0, const uint32_t* dbits = nullptr);      // d_ready: 1 = aux row 1 (D) was written by the caller, 2 = all four rows were
// ---- tcgen05 / TMEM flash attention for long sequences, 64-wide heads (mt_attention_flash.cu), bf16 ------------------------
bool mt_attn_flash_supported(int B, int T, int d, int h);
int mt_attn_flash_fwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop, cudaStream_t st);
// ws: mt_attn_bwd_ws_floats(B, T, h) floats; the QKV bias gradient is NOT produced (the caller column-sums dqkv)
int mt_attn_flash_bwd_run(int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse, const void* dout,
                          void* dqkv, DropCfg drop, float* ws, cudaStream_t st);
// workspace of any attention backward (Dws of mt_attn_bwd_run), in floats
// (per-query scalars [B][h][4][T rounded up to 128], plus -- long sequences -- the fp32 dQ accumulation buffer [B*T, 64 h] of the flash backward)
static inline size_t mt_attn_bwd_ws_floats(int B, int T, int h) {
  const size_t tpad = ((size_t)T + 127) / 128 * 128;
  return 4 * (size_t)B * tpad * (size_t)h + 64 + (T > 128 ? (size_t)B * (size_t)T * 64 * (size_t)h : 0);
}

// ---- attention (mt_attention.cu) ---------------------------------------------------------------------
int mt_attn_fwd_run(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse, DropCfg drop,
                    cudaStream_t st, const int* klen = nullptr);
int mt_attn_bwd_run(int dtype, int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                    const void* dout, void* dqkv, DropCfg drop, float* Dws, cudaStream_t st, float* dbias = nullptr);
// G modality stacks back to back (B narratives each; mask / klen shared; drops[g] = stack g's dropout stream; Dws: G * mt_attn_bwd_ws_floats;
// dbias of stack g at dbias + g * dbias_gstride): one launch on the tcgen05 engine, one call per stack on the other engines
int mt_attn_group_fwd_run(int dtype, int G, int B, int T, int d, int h, const void* qkv, const float* mask, void* out, float* lse,
                          const DropCfg* drops, cudaStream_t st, const int* klen = nullptr, const uint32_t* dbits = nullptr);
int mt_attn_group_bwd_run(int dtype, int G, int B, int T, int d, int h, const void* qkv, const float* mask, const void* out, const float* lse,
                          const void* dout, void* dqkv, const DropCfg* drops, float* Dws, cudaStream_t st, float* dbias, size_t dbias_gstride,
                          int d_ready = 0, const uint32_t* dbits = nullptr);
// true when mt_attn_group_bwd_run will take the tcgen05 engine for these arguments (then the caller may pre-fill D = rowsum(dout . out)
// into Dws[((b * h + head) * 4 + 1) * 128 + q] and pass d_ready)
bool mt_attn_group_bwd_uses_tc(int dtype, int G, int B, int T, int d, int h, const void* qkv, const void* out, const void* dout,
                               const void* dqkv, const float* Dws, const float* dbias);
