// Encoder stack (Encoder.forward MFT/multiTransformer.py:73-76 over EncoderLayer :106-116) and the stand-alone
// Linear (embed / fusion layers), forward and backward, as sequences of GEMMs with fused epilogues, fused
// attention and row-wise LayerNorm kernels.  The residual stream stays fp32; GEMM operands are `dtype`.
#include "mt_ops.cuh"
#include "mt_gemm_rs.cuh"

namespace {

struct EncParams {
  // offsets (floats) inside one layer's block of the flat layout documented in include/mt_b200.h
  size_t w_qkv, b_qkv, w_o, b_o, w_1, b_1, w_2, b_2, ln1_a, ln1_b, ln2_a, ln2_b, layer_stride;
  size_t lnf_a, lnf_b, total;
};

EncParams enc_params(int d, int dff, int n_layers) {
  EncParams p;
  size_t o = 0;
  p.w_qkv = o; o += (size_t)3 * d * d;
  p.b_qkv = o; o += (size_t)3 * d;
  p.w_o = o; o += (size_t)d * d;
  p.b_o = o; o += d;
  p.w_1 = o; o += (size_t)dff * d;
  p.b_1 = o; o += dff;
  p.w_2 = o; o += (size_t)d * dff;
  p.b_2 = o; o += d;
  p.ln1_a = o; o += d;
  p.ln1_b = o; o += d;
  p.ln2_a = o; o += d;
  p.ln2_b = o; o += d;
  p.layer_stride = o;
  p.lnf_a = o * n_layers;
  p.lnf_b = p.lnf_a + d;
  p.total = p.lnf_b + d;
  return p;
}

// Activations kept per layer for backward (training) or reused across layers (inference).
struct LayerBufs {
  float* x_out;   // fp32 [M,d] residual stream after the layer
  void* u;        // LN1(x)            [M,d]   dtype
  void* qkv;      //                   [M,3d]  dtype
  float* lse;     //                   [B,h,T]
  void* att;      // merged heads      [M,d]   dtype
  float* xp;      // x + attn sublayer [M,d]   fp32
  void* v;        // LN2(xp)           [M,d]   dtype
  void* hid;      // relu/dropout FFN hidden [M,dff] dtype
  uint32_t* dbits; // keep bits of the attention-probability dropout (tcgen05 attention, training): drawn once, read by forward and backward
};

struct EncWs {
  LayerBufs L[64];
  // backward scratch
  float* g0; float* g1;      // fp32 [M,d] gradient of the residual stream (ping-pong)
  void* dact;                // [M,d] dtype (dropout-applied sublayer gradient; later du / dv)
  void* dact2;               // [M,d] dtype
  void* dhid;                // [M,dff] dtype
  void* dqkv;                // [M,3d] dtype
  float* Dws;                // [B,h,T]
  size_t bytes;
};

// A call serves G modality stacks of identical shape (G = 1: the plain single-stack entry points).  Every activation buffer holds the
// stacks back to back ([G*M, .], stack g at rows g*M), the parameter / gradient blocks of consecutive stacks are `pstride` floats apart.
struct Groups {
  int G = 1;
  size_t pstride = 0;
  uint64_t seed[MT_RS_MAX_GROUPS] = {};
  int stack_id[MT_RS_MAX_GROUPS] = {};
  bool grouped = false;      // came in through mt_encoder_group_*: the only entry that takes part in the overlapped all-reduce
};

int carve(const MtEncoderCfg& c, int G, void* ws, EncWs& w) {
  if (c.n_layers < 1 || c.n_layers > 64) return MT_ERR_ARG;
  const size_t M = (size_t)G * c.B * c.T, d = c.d, es = mt_esize(c.dtype);
  WsCarver k(ws);
  const int sets = c.training ? c.n_layers : 1;
  for (int s = 0; s < sets; ++s) {
    LayerBufs& b = w.L[s];
    b.x_out = k.take<float>(M * d);
    b.u = k.take_bytes(M * d * es);
    b.qkv = k.take_bytes(M * 3 * d * es);
    b.lse = k.take<float>((size_t)G * c.B * c.h * c.T);
    b.att = k.take_bytes(M * d * es);
    b.xp = k.take<float>(M * d);
    b.v = k.take_bytes(M * d * es);
    b.hid = k.take_bytes(M * c.dff * es);
    b.dbits = (c.training && c.dtype == MT_BF16 && mt_attn_tc_supported(c.B, c.T, c.d, c.h)) ? k.take<uint32_t>(mt_attn_tc_dropbits_words(G, c.B, c.h)) : nullptr;
  }
  float* x_alt = c.training ? nullptr : k.take<float>(M * d);   // inference ping-pong of the residual stream
  for (int l = sets; l < c.n_layers; ++l) {
    w.L[l] = w.L[0];
    if ((l & 1) && x_alt) w.L[l].x_out = x_alt;
  }
  if (c.training) {
    w.g0 = k.take<float>(M * d);
    w.g1 = k.take<float>(M * d);
    w.dact = k.take_bytes(M * d * es);
    w.dact2 = k.take_bytes(M * d * es);
    w.dhid = k.take_bytes(M * c.dff * es);
    w.dqkv = k.take_bytes(M * 3 * d * es);
    w.Dws = k.take<float>((size_t)G * mt_attn_bwd_ws_floats(c.B, c.T, c.h));
  }
  w.bytes = k.total();
  return MT_OK;
}

// the stack's kernels share the SMs with `share - 1` concurrent streams for the duration of one C call (mt_tune knobs)
struct GridShareScope {
  int g0, g2;
  explicit GridShareScope(int share) : g0(g_mt_tune[MT_TUNE_GEMM_SHARE]), g2(g_mt_tune[MT_TUNE_LN_SHARE]) {
    if (share > 1) { g_mt_tune[MT_TUNE_GEMM_SHARE] = share; g_mt_tune[MT_TUNE_LN_SHARE] = share; }
  }
  ~GridShareScope() { g_mt_tune[MT_TUNE_GEMM_SHARE] = g0; g_mt_tune[MT_TUNE_LN_SHARE] = g2; }
};

int check_cfg(const MtEncoderCfg* c) {
  if (!c) return MT_ERR_ARG;
  if (c->B <= 0 || c->T <= 0 || c->d <= 0 || c->h <= 0 || c->dff <= 0 || c->n_layers <= 0) return MT_ERR_ARG;
  if (c->d % c->h != 0 || c->d % 128 != 0 || c->d > 1024 || c->dff % 4 != 0) return MT_ERR_ARG;
  if (c->dtype != MT_F32 && c->dtype != MT_BF16) return MT_ERR_ARG;
  if ((size_t)c->B * c->T > 0x7fffffffull / (3 * (size_t)c->d)) return MT_ERR_ARG;   // 32-bit row*ld products inside the GEMMs
  return MT_OK;
}

inline const void* wptr(const MtEncoderCfg& c, const float* params, const void* params_lp, size_t off) {
  return c.dtype == MT_BF16 ? (const void*)((const bf16*)params_lp + off) : (const void*)(params + off);
}

// y = epi(x W^T): A = x [M,K] k-major, B = W [N,K] k-major
GemmDesc fwd_gemm(int M, int N, int K, const void* x, const void* W, void* y, bool y_f32) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.A = x; g.lda = K; g.a_kmajor = true;
  g.B = W; g.ldb = K; g.b_kmajor = true;
  g.C = y; g.ldc = N; g.c_f32 = y_f32;
  return g;
}
// dx = dy W: A = dy [M,Nout] k-major, B = W [Nout,Kin] read as B(n = kin, k = nout) -> mn-major
GemmDesc dgrad_gemm(int M, int Nout, int Kin, const void* dy, const void* W, void* dx, bool dx_f32) {
  GemmDesc g;
  g.M = M; g.N = Kin; g.K = Nout;
  g.A = dy; g.lda = Nout; g.a_kmajor = true;
  g.B = W; g.ldb = Kin; g.b_kmajor = false;
  g.C = dx; g.ldc = Kin; g.c_f32 = dx_f32;
  return g;
}
// dW[Nout,Kin] = dy^T x : contraction over the M tokens, both operands mn-major, split-K with fp32 atomics
GemmDesc wgrad_gemm(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw, int groups = 1) {
  GemmDesc g;
  g.M = Nout; g.N = Kin; g.K = M;
  g.A = dy; g.lda = ldy; g.a_kmajor = false;
  g.B = x; g.ldb = ldx; g.b_kmajor = false;
  g.C = dW; g.ldc = ldw; g.c_f32 = true;
  size_t tiles = (size_t)((Nout + 63) / 64) * ((Kin + 63) / 64) * groups;
  int split = (int)((148 * 4 + tiles - 1) / tiles);
  int max_split = (M + 255) / 256;
  if (split > max_split) split = max_split;
  g.split_k = split < 2 ? 2 : split;          // always the atomic epilogue: dW is zero-initialised by the caller
  return g;
}

// One projection of the stack(s):  C[g] = epi(A[g] . W_g) for every group -- through the weight-resident row-stream engine
// (mt_gemm_rs.cu, ONE launch for all groups) when the shape is one of its instantiations, else one streaming GEMM per group.
struct Proj {
  const MtEncoderCfg& c;
  const Groups& gr;
  const float* params; const void* params_lp;
  float* grads;
  cudaStream_t st;
  int M;                                      // rows per group

  // w_off / b_off: offsets inside a group's parameter block.  fwd: W [N,K]; dgrad: W [K(out),N(in)] read transposed.
  // drop_k >= 0: output dropout site k of `layer`; colsum_off: bias-gradient block (grads) receiving the column sums;
  // ln_* (fwd only): fused LayerNorm of the fp32 output with the gains at ln_a_off / ln_b_off -> ln_out (bf16)
  int run(bool dgrad, int N, int K, const void* A, size_t w_off, void* C, bool c_f32, long long b_off, int act, int layer, int drop_k,
          const void* gate, float gate_scale, const float* residual, long long colsum_off, void* ln_out = nullptr, size_t ln_a_off = 0,
          size_t ln_b_off = 0) const {
    const bool lp = c.dtype == MT_BF16;
    const float p = drop_k >= 0 ? c.p_drop : 0.f;
    if (lp) {
      RsDesc r;
      r.G = gr.G; r.Mg = M; r.N = N; r.K = K;
      r.A = A; r.lda = K;
      r.ldb = dgrad ? N : K; r.b_kmajor = !dgrad;
      r.C = C; r.ldc = N; r.c_f32 = c_f32;
      r.act = act;
      r.gate = gate; r.ldg = N; r.gate_scale = gate_scale;
      r.residual = residual; r.ldr = N;
      r.ln_out = ln_out; r.ld_ln = N; r.ln_eps = 1e-6f;
      for (int g = 0; g < gr.G; ++g) {
        r.B[g] = (const bf16*)params_lp + g * gr.pstride + w_off;
        r.bias[g] = b_off >= 0 ? params + g * gr.pstride + b_off : nullptr;
        r.drop[g] = mt_make_drop(p, gr.seed[g], mt_enc_site(gr.stack_id[g], layer, drop_k < 0 ? 0 : drop_k));
        r.colsum[g] = colsum_off >= 0 ? grads + g * gr.pstride + colsum_off : nullptr;
        r.ln_a[g] = params + g * gr.pstride + ln_a_off;
        r.ln_b[g] = params + g * gr.pstride + ln_b_off;
      }
      if (!g_mt_tune[MT_TUNE_NO_RS] && mt_gemm_rs_supported(r)) return mt_gemm_rs_run(r, st);
    }
    const size_t es = mt_esize(c.dtype);
    // a projection the row-stream engine does not take (the long-K input gradient of the QKV projection) with nothing per group but its
    // weight matrix: ONE streaming launch over the stacked rows, tile -> weight matrix of its stack (GemmDesc.mgroups)
    if (lp && gr.G > 1 && b_off < 0 && act == MT_ACT_NONE && p == 0.f && colsum_off < 0 && !ln_out && !g_mt_tune[MT_TUNE_NO_MGROUPS]) {
      const void* W0 = wptr(c, params, params_lp, w_off);
      GemmDesc d = dgrad ? dgrad_gemm(gr.G * M, K, N, A, W0, C, c_f32) : fwd_gemm(gr.G * M, N, K, A, W0, C, c_f32);
      d.mgroups = gr.G; d.b_gstride = (long long)gr.pstride;
      if (gate) { d.epi.gate = gate; d.epi.ldg = N; d.epi.gate_scale = gate_scale; }
      if (residual) { d.epi.residual = residual; d.epi.ldr = N; }
      const int rc = mt_gemm_run(c.dtype, d, st);
      if (rc != MT_ERR_UNSUPPORTED) return rc;
    }
    for (int g = 0; g < gr.G; ++g) {
      const size_t ro = (size_t)g * M;
      const void* Ag = (const char*)A + ro * K * es;
      void* Cg = (char*)C + ro * N * (c_f32 ? 4 : es);
      const void* W = wptr(c, params + g * gr.pstride, lp ? (const void*)((const bf16*)params_lp + g * gr.pstride) : nullptr, w_off);
      GemmDesc d = dgrad ? dgrad_gemm(M, K, N, Ag, W, Cg, c_f32 || !lp) : fwd_gemm(M, N, K, Ag, W, Cg, c_f32 || !lp);
      if (b_off >= 0) d.epi.bias = params + g * gr.pstride + b_off;
      d.epi.act = act;
      d.epi.drop = mt_make_drop(p, gr.seed[g], mt_enc_site(gr.stack_id[g], layer, drop_k < 0 ? 0 : drop_k));
      if (gate) { d.epi.gate = (const char*)gate + ro * N * es; d.epi.ldg = N; d.epi.gate_scale = gate_scale; }
      if (residual) { d.epi.residual = residual + ro * N; d.epi.ldr = N; }
      if (colsum_off >= 0) d.epi.colsum = grads + g * gr.pstride + colsum_off;
      MT_TRY(mt_gemm_run(c.dtype, d, st));
    }
    if (ln_out)
      MT_TRY(mt_ln_fwd_run(M, N, (const float*)C, params + ln_a_off, params + ln_b_off, 1e-6f, ln_out, lp, st, gr.G, gr.pstride));
    return MT_OK;
  }

  // dW[g] = dy[g]^T x[g] for every group, one launch on the tcgen05 engine when the grouped split-K form applies
  int wgrad(int Nout, int Kin, const void* dy, const void* x, size_t w_off) const {
    const size_t es = mt_esize(c.dtype);
    if (gr.G > 1 && c.dtype == MT_BF16 && M % 64 == 0) {
      GemmDesc d = wgrad_gemm(M, Nout, Kin, dy, Nout, x, Kin, grads + w_off, Kin, gr.G);
      d.groups = gr.G; d.c_gstride = (long long)gr.pstride;
      const int rc = mt_gemm_run(c.dtype, d, st);
      if (rc != MT_ERR_UNSUPPORTED) return rc;
    }
    for (int g = 0; g < gr.G; ++g) {
      const size_t ro = (size_t)g * M;
      MT_TRY(mt_gemm_run(c.dtype, wgrad_gemm(M, Nout, Kin, (const char*)dy + ro * Nout * es, Nout, (const char*)x + ro * Kin * es, Kin,
                                              grads + g * gr.pstride + w_off, Kin), st));
    }
    return MT_OK;
  }
};

int encoder_fwd_impl(const MtEncoderCfg& c, const Groups& gr, const float* params, const void* params_lp, const float* x, const float* mask,
                     void* y, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!params || !x || !y || !ws || (c.dtype == MT_BF16 && !params_lp)) return MT_ERR_ARG;
  if (c.key_len && (c.training || c.p_drop > 0.f)) return MT_ERR_UNSUPPORTED;      // ragged batches: inference only
  EncWs w;
  MT_TRY(carve(c, gr.G, ws, w));
  if (ws_bytes < w.bytes) return MT_ERR_WS;
  const int M = c.B * c.T, d = c.d, dff = c.dff, G = gr.G;
  const bool lp = c.dtype == MT_BF16;
  const size_t es = mt_esize(c.dtype);
  const EncParams P = enc_params(d, dff, c.n_layers);
  const float* xin = x;
  GridShareScope share(c.grid_share);
  const Proj pj{c, gr, params, params_lp, nullptr, st, M};
  // the LayerNorm that follows a layer (the next layer's first norm, or the final norm) rides in the epilogue of that layer's last
  // GEMM whenever its output has the operand dtype and the row-stream engine takes the shape
  const bool y_lp = lp && !c.y_f32;
  bool u_ready = false;
  // Keep bits of the attention-probability dropout (tcgen05 engine, training): drawn once per layer, right before its attention; the
  // backward kernel reads the same words.  (Measured: drawing all layers up front on a side stream buys nothing -- the row-stream GEMMs
  // and LayerNorm kernels it would run under own every register of their SMs, so the ALU-bound draw kernel cannot co-reside.)
  const bool use_bits = c.training && c.p_drop > 0.f && !c.key_len && !g_mt_tune[MT_TUNE_NO_DROPBITS] && w.L[0].dbits != nullptr;
  // ... so the draw rides inside an HBM-bound kernel instead: the stand-alone LayerNorm pass that precedes an attention in program order
  // (layer 0's first norm, every layer's second norm for the NEXT layer) hashes that attention's keep bits between its row loads and
  // their first use (ln_fwd_kernel<.., DRAW>); only an attention without such a pass in front of it gets the draw kernel
  const bool ln_draws = use_bits && !g_mt_tune[MT_TUNE_NO_LNDRAW];
  bool bits_ready = false;
  auto attn_drops = [&](int layer, DropCfg* ad) {
    for (int g = 0; g < G; ++g) ad[g] = mt_make_drop(c.p_drop, gr.seed[g], mt_enc_site(gr.stack_id[g], layer, MT_SITE_ATTN_P));
  };
  MtBitsJob job;
  auto draw_job = [&](int layer) -> const MtBitsJob* {
    if (!ln_draws) return nullptr;
    DropCfg ad[MT_RS_MAX_GROUPS];
    attn_drops(layer, ad);
    if (mt_attn_tc_dropbits_job(G, c.B, c.T, c.h, ad, w.L[layer].dbits, &job) != MT_OK) return nullptr;
    bits_ready = true;
    return &job;
  };
  for (int l = 0; l < c.n_layers; ++l) {
    const size_t base = P.layer_stride * l;
    LayerBufs& b = w.L[l];
    const bool last = l == c.n_layers - 1;
    // sublayer 0: x + dropout(self_attn(LN(x)))
    if (!u_ready) MT_TRY(mt_ln_fwd_run(M, d, xin, params + base + P.ln1_a, params + base + P.ln1_b, 1e-6f, b.u, lp, st, G, gr.pstride, draw_job(l)));
    MT_TRY(pj.run(false, 3 * d, d, b.u, base + P.w_qkv, b.qkv, !lp, (long long)(base + P.b_qkv), MT_ACT_NONE, l, -1, nullptr, 1.f, nullptr, -1));
    {
      DropCfg ad[MT_RS_MAX_GROUPS];
      attn_drops(l, ad);
      const uint32_t* bits = use_bits ? b.dbits : nullptr;
      if (bits && !bits_ready) MT_TRY(mt_attn_tc_dropbits_run(G, c.B, c.T, c.h, ad, b.dbits, st));
      bits_ready = false;
      MT_TRY(mt_attn_group_fwd_run(c.dtype, G, c.B, c.T, d, c.h, b.qkv, mask, b.att, b.lse, ad, st, c.key_len, bits));
    }
    // mt_tune(11, 1): ... with the sublayer-1 LayerNorm in its epilogue (the two 128-column slices of a row tile are a CTA pair that exchanges
    // the row moments through distributed shared memory, mt_gemm_rs.cu R_LNX).  Opt-in: measured 78.7 us per three stacks against 49.7 + 32
    // for projection + LayerNorm pass -- the per-tile exchange and the second sweep sit on the epilogue, which already paces this kernel.
    const bool ln2_fused = lp && !g_mt_tune[MT_TUNE_NO_LNFUSE] && g_mt_tune[MT_TUNE_LNX];
    MT_TRY(pj.run(false, d, d, b.att, base + P.w_o, b.xp, true, (long long)(base + P.b_o), MT_ACT_NONE, l, MT_SITE_SUB0, nullptr, 1.f, xin, -1,
                  ln2_fused ? b.v : nullptr, base + P.ln2_a, base + P.ln2_b));
    // sublayer 1: x + dropout(w_2(dropout(relu(w_1(LN(x))))))
    if (!ln2_fused) MT_TRY(mt_ln_fwd_run(M, d, b.xp, params + base + P.ln2_a, params + base + P.ln2_b, 1e-6f, b.v, lp, st, G, gr.pstride,
                                         last ? nullptr : draw_job(l + 1)));
    MT_TRY(pj.run(false, dff, d, b.v, base + P.w_1, b.hid, !lp, (long long)(base + P.b_1), MT_ACT_RELU, l, MT_SITE_FFN_H, nullptr, 1.f, nullptr, -1));
    void* ln_out = nullptr;
    size_t la = 0, lb = 0;
    if (lp && !g_mt_tune[MT_TUNE_NO_LNFUSE] && (!last || y_lp)) {
      ln_out = last ? y : w.L[l + 1].u;
      la = last ? P.lnf_a : base + P.layer_stride + P.ln1_a;
      lb = last ? P.lnf_b : base + P.layer_stride + P.ln1_b;
    }
    MT_TRY(pj.run(false, d, dff, b.hid, base + P.w_2, b.x_out, true, (long long)(base + P.b_2), MT_ACT_NONE, l, MT_SITE_SUB1, nullptr, 1.f, b.xp, -1,
                  ln_out, la, lb));
    u_ready = ln_out != nullptr;
    xin = b.x_out;
  }
  if (u_ready) return MT_OK;      // the final norm left with the last FFN GEMM
  return mt_ln_fwd_run(M, d, xin, params + P.lnf_a, params + P.lnf_b, 1e-6f, y, y_lp, st, G, gr.pstride);
}

int encoder_bwd_impl(const MtEncoderCfg& c, const Groups& gr, const float* params, const void* params_lp, const float* x, const float* mask,
                     const void* dy, float* dx, float* grads, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!c.training) return MT_ERR_ARG;
  if (!params || !x || !dy || !dx || !grads || !ws || (c.dtype == MT_BF16 && !params_lp)) return MT_ERR_ARG;
  EncWs w;
  MT_TRY(carve(c, gr.G, ws, w));
  if (ws_bytes < w.bytes) return MT_ERR_WS;
  const int M = c.B * c.T, d = c.d, dff = c.dff, G = gr.G;
  const bool lp = c.dtype == MT_BF16;
  const size_t es = mt_esize(c.dtype);
  const EncParams P = enc_params(d, dff, c.n_layers);
  const float p = c.p_drop;
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  GridShareScope share(c.grid_share);
  const Proj pj{c, gr, params, params_lp, grads, st, M};
  for (int g = 0; g < G; ++g) MT_CUDA(cudaMemsetAsync(grads + g * gr.pstride, 0, sizeof(float) * P.total, st));

  // The residual-stream gradient dL/dx_l between the sublayers stays fp32.  mt_tune(13, 1) carries it in bf16 inside a bf16-mode stack
  // (ln_bwd_kernel GM = 1 / 2; only the stack's own dx leaves in fp32): 12 instead of 16 bytes per element of the LayerNorm backward, but
  // measured slower on B200 (1.00 ms vs 0.94 ms per step over the 13 launches: the kernel is bound by its memory-instruction rate, not
  // by bytes, and 8-byte accesses carry half as much per instruction) -- kept as an A/B switch with its parity test.
  const bool dy_lp = lp && !c.y_f32;
  const bool g_lp = lp && dy_lp && g_mt_tune[MT_TUNE_BF16_GSTREAM];
  float* g_cur = w.g0;
  float* g_nxt = w.g1;
  const float* x_last = w.L[c.n_layers - 1].x_out;
  DropCfg drops[MT_RS_MAX_GROUPS];
  auto site_drops = [&](int layer, int k) {
    for (int g = 0; g < G; ++g) drops[g] = mt_make_drop(p, gr.seed[g], mt_enc_site(gr.stack_id[g], layer, k));
    return drops[0];
  };
  // every LayerNorm backward also emits the dropped operand-dtype gradient the sublayer below starts from, and that
  // sublayer's output-bias gradient (LnBwdNext), so no separate dropout-gradient / column-sum passes run over [M,d]
  {
    const size_t bl = P.layer_stride * (c.n_layers - 1);
    const DropCfg dr = site_drops(c.n_layers - 1, MT_SITE_SUB1);
    LnBwdNext nx{w.dact, grads + bl + P.b_2, dr};
    if (dy_lp == lp) {
      MT_TRY(mt_ln_bwd_run(M, d, x_last, params + P.lnf_a, 1e-6f, dy, dy_lp, nullptr, g_cur, grads + P.lnf_a, grads + P.lnf_b, st, &nx, G,
                           gr.pstride, drops, g_lp ? 1 : 0));
    } else {      // fp32 dy in bf16 mode: the fused second output has dy's dtype, so take the two-pass route once
      MT_TRY(mt_ln_bwd_run(M, d, x_last, params + P.lnf_a, 1e-6f, dy, dy_lp, nullptr, g_cur, grads + P.lnf_a, grads + P.lnf_b, st, nullptr, G,
                           gr.pstride));
      for (int g = 0; g < G; ++g) {
        const size_t ro = (size_t)g * M * d;
        MT_TRY(mt_drop_grad_run(M, d, g_cur + ro, (char*)w.dact + ro * es, lp, drops[g], st));
        MT_TRY(mt_colsum_run(lp, M, d, (char*)w.dact + ro * es, d, grads + g * gr.pstride + bl + P.b_2, 1, st));
      }
    }
  }

  const int ar_split = gr.grouped ? mt_comm_overlap_split(c.n_layers) : -1;
  for (int l = c.n_layers - 1; l >= 0; --l) {
    const size_t base = P.layer_stride * l;
    LayerBufs& b = w.L[l];
    const float* x_l = l == 0 ? x : w.L[l - 1].x_out;
    // ---- FFN sublayer: x_out = xp + drop(w_2 hid + b_2); w.dact = drop' . g_cur and db_2 are already there ----------
    MT_TRY(pj.wgrad(d, dff, w.dact, b.hid, base + P.w_2));
    // relu' and the hidden dropout mask in one test (gate), db_1 = colsum(dhid) from the epilogue
    MT_TRY(pj.run(true, dff, d, w.dact, base + P.w_2, w.dhid, !lp, -1, MT_ACT_NONE, l, -1, b.hid, keep_scale, nullptr, (long long)(base + P.b_1)));
    MT_TRY(pj.wgrad(dff, d, w.dhid, b.v, base + P.w_1));
    MT_TRY(pj.run(true, d, dff, w.dhid, base + P.w_1, w.dact2, !lp, -1, MT_ACT_NONE, l, -1, nullptr, 1.f, nullptr, -1));
    {
      LnBwdNext nx{w.dact, grads + base + P.b_o, site_drops(l, MT_SITE_SUB0)};
      MT_TRY(mt_ln_bwd_run(M, d, b.xp, params + base + P.ln2_a, 1e-6f, w.dact2, lp, g_cur, g_nxt, grads + base + P.ln2_a, grads + base + P.ln2_b, st,
                           &nx, G, gr.pstride, drops, g_lp ? 1 : 0));
    }
    // ---- attention sublayer: xp = x + drop(att w_o + b_o); w.dact = drop' . g_nxt and db_o are already there ---------
    MT_TRY(pj.wgrad(d, d, w.dact, b.att, base + P.w_o));
    // d att = d out . w_o; when the tcgen05 attention backward follows, the same GEMM leaves D = rowsum(d att . att) per head in the
    // attention kernel's per-query scalars (no separate pass over att and d att)
    int d_ready = 0;
    if (lp && !g_mt_tune[MT_TUNE_NO_RS] && d == 32 * c.h &&
        mt_attn_group_bwd_uses_tc(c.dtype, G, c.B, c.T, d, c.h, b.qkv, b.att, w.dact2, w.dqkv, w.Dws, grads + base + P.b_qkv)) {
      RsDesc r;
      r.G = G; r.Mg = M; r.N = d; r.K = d;
      r.A = w.dact; r.lda = d; r.ldb = d; r.b_kmajor = false;
      r.C = w.dact2; r.ldc = d; r.c_f32 = false;
      r.attd_src = b.att; r.attd_ld = d; r.attd_aux = w.Dws; r.attd_T = c.T;
      const bool all_rows = !(g_mt_tune[MT_TUNE_PDL_DEBUG] & 2);      // the epilogue writes the other three per-query scalars as well (no light pass)
      if (all_rows) { r.attd_lse = b.lse; r.attd_mask = mask; r.attd_scale = 1.0f / sqrtf(32.0f); }
      for (int g = 0; g < G; ++g) { r.B[g] = (const bf16*)params_lp + g * gr.pstride + base + P.w_o; r.drop[g] = mt_make_drop(0.f, 0, 0); }
      if (mt_gemm_rs_supported(r)) { MT_TRY(mt_gemm_rs_run(r, st)); d_ready = all_rows ? 2 : 1; }
    }
    if (!d_ready) MT_TRY(pj.run(true, d, d, w.dact, base + P.w_o, w.dact2, !lp, -1, MT_ACT_NONE, l, -1, nullptr, 1.f, nullptr, -1));
    {
      DropCfg ad[MT_RS_MAX_GROUPS];
      for (int g = 0; g < G; ++g) ad[g] = mt_make_drop(p, gr.seed[g], mt_enc_site(gr.stack_id[g], l, MT_SITE_ATTN_P));
      const uint32_t* bits = (b.dbits && p > 0.f && !c.key_len && !g_mt_tune[MT_TUNE_NO_DROPBITS]) ? b.dbits : nullptr;      // drawn by the forward
      MT_TRY(mt_attn_group_bwd_run(c.dtype, G, c.B, c.T, d, c.h, b.qkv, mask, b.att, b.lse, w.dact2, w.dqkv, ad, w.Dws, st,
                                   grads + base + P.b_qkv, gr.pstride, d_ready, bits));
    }
    MT_TRY(pj.wgrad(3 * d, d, w.dqkv, b.u, base + P.w_qkv));
    MT_TRY(pj.run(true, d, 3 * d, w.dqkv, base + P.w_qkv, w.dact2, !lp, -1, MT_ACT_NONE, l, -1, nullptr, 1.f, nullptr, -1));
    float* out = l == 0 ? dx : g_cur;
    LnBwdNext nx{nullptr, nullptr, mt_make_drop(0.f, 0, 0)};
    if (l > 0) nx = LnBwdNext{w.dact, grads + base - P.layer_stride + P.b_2, site_drops(l - 1, MT_SITE_SUB1)};
    MT_TRY(mt_ln_bwd_run(M, d, x_l, params + base + P.ln1_a, 1e-6f, w.dact2, lp, g_nxt, out, grads + base + P.ln1_a, grads + base + P.ln1_b, st, &nx,
                         G, gr.pstride, drops, g_lp ? (l == 0 ? 2 : 1) : 0));
    // g_cur now holds dL/dx_l (g_nxt is free again)
    // data-parallel training: every gradient of layers >= l and of the final norm is enqueued -- their all-reduce may start now, on the
    // communication stream, under the backward of layers l - 1 .. 0 (mt_comm_overlap_arm)
    if (gr.grouped && l == ar_split) MT_TRY(mt_comm_overlap_fire(grads, gr.pstride, G, base, P.total, st));
  }
  return MT_OK;
}

Groups single_group(const MtEncoderCfg& c) {
  Groups g;
  g.G = 1; g.pstride = 0; g.seed[0] = c.seed; g.stack_id[0] = c.stack_id;
  return g;
}

int make_groups(const MtEncoderCfg& c, int G, const uint64_t* seeds, const int* stack_ids, size_t pstride, Groups& g) {
  if (G < 1 || G > MT_RS_MAX_GROUPS) return MT_ERR_ARG;
  if (G > 1 && pstride < enc_params(c.d, c.dff, c.n_layers).total) return MT_ERR_ARG;
  if (G > 1 && pstride % 8 != 0) return MT_ERR_ALIGN;      // bf16 weight blocks of every group stay 16-byte aligned (TMA)
  g.G = G; g.pstride = G > 1 ? pstride : 0;
  for (int i = 0; i < G; ++i) { g.seed[i] = seeds ? seeds[i] : c.seed; g.stack_id[i] = stack_ids ? stack_ids[i] : c.stack_id + i; }
  return MT_OK;
}

}  // namespace

// shared with mt_mfn.cu
GemmDesc mt_wgrad_desc(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw) {
  return wgrad_gemm(M, Nout, Kin, dy, ldy, x, ldx, dW, ldw);
}

extern "C" {

size_t mt_encoder_param_count(int d, int dff, int n_layers) { return enc_params(d, dff, n_layers).total; }

size_t mt_encoder_ws_bytes(const MtEncoderCfg* cfg) {
  if (check_cfg(cfg) != MT_OK) return 0;
  EncWs w;
  if (carve(*cfg, 1, nullptr, w) != MT_OK) return 0;
  return w.bytes;
}

size_t mt_encoder_group_ws_bytes(const MtEncoderCfg* cfg, int n_stacks) {
  if (check_cfg(cfg) != MT_OK || n_stacks < 1 || n_stacks > MT_RS_MAX_GROUPS) return 0;
  if ((size_t)n_stacks * cfg->B * cfg->T > 0x7fffffffull / (3 * (size_t)cfg->d)) return 0;
  EncWs w;
  if (carve(*cfg, n_stacks, nullptr, w) != MT_OK) return 0;
  return w.bytes;
}

int mt_encoder_fwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x, const float* mask, void* y,
                   void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  return encoder_fwd_impl(*cfg, single_group(*cfg), params, params_lp, x, mask, y, ws, ws_bytes, (cudaStream_t)stream);
}

int mt_encoder_bwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x, const float* mask,
                   const void* dy, float* dx, float* grads, void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  return encoder_bwd_impl(*cfg, single_group(*cfg), params, params_lp, x, mask, dy, dx, grads, ws, ws_bytes, (cudaStream_t)stream);
}

int mt_encoder_group_fwd(const MtEncoderCfg* cfg, int n_stacks, const uint64_t* seeds, const int* stack_ids, const float* params,
                         const void* params_lp, size_t param_stride, const float* x, const float* mask, void* y, void* ws, size_t ws_bytes,
                         void* stream) {
  MT_TRY(check_cfg(cfg));
  if (mt_encoder_group_ws_bytes(cfg, n_stacks) == 0) return MT_ERR_ARG;
  Groups g;
  MT_TRY(make_groups(*cfg, n_stacks, seeds, stack_ids, param_stride, g));
  return encoder_fwd_impl(*cfg, g, params, params_lp, x, mask, y, ws, ws_bytes, (cudaStream_t)stream);
}

int mt_encoder_group_bwd(const MtEncoderCfg* cfg, int n_stacks, const uint64_t* seeds, const int* stack_ids, const float* params,
                         const void* params_lp, size_t param_stride, const float* x, const float* mask, const void* dy, float* dx, float* grads,
                         void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  if (mt_encoder_group_ws_bytes(cfg, n_stacks) == 0) return MT_ERR_ARG;
  Groups g;
  MT_TRY(make_groups(*cfg, n_stacks, seeds, stack_ids, param_stride, g));
  g.grouped = true;
  return encoder_bwd_impl(*cfg, g, params, params_lp, x, mask, dy, dx, grads, ws, ws_bytes, (cudaStream_t)stream);
}

int mt_encoder_stack_fwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x, const float* mask, void* y,
                         void* ws, size_t ws_bytes, void* stream) {
  return mt_encoder_fwd(cfg, params, params_lp, x, mask, y, ws, ws_bytes, stream);
}
int mt_encoder_stack_bwd(const MtEncoderCfg* cfg, const float* params, const void* params_lp, const float* x, const float* mask,
                         const void* dy, float* dx, float* grads, void* ws, size_t ws_bytes, void* stream) {
  return mt_encoder_bwd(cfg, params, params_lp, x, mask, dy, dx, grads, ws, ws_bytes, stream);
}

// ------------------------------------------------------------------------------------------------------
// Linear.  W is always the fp32 master weight [N,K]; in bf16 mode it is cast (and K-padded to a multiple of 8
// so rows are 16-byte aligned for TMA) into the workspace on every call -- these matrices are tiny.
// ------------------------------------------------------------------------------------------------------
static inline int lin_kp(int dtype, int K) { return dtype == MT_BF16 ? (K + 7) / 8 * 8 : K; }
static inline bool lin_stage_x(int dtype, int K, int x_f32, float in_drop_p) {
  return in_drop_p > 0.f || (dtype == MT_BF16 && ((K % 8) != 0 || x_f32));
}

size_t mt_linear_ws_bytes(int dtype, int M, int N, int K, int x_f32, float in_drop_p) {
  const int Kp = lin_kp(dtype, K);
  WsCarver k(nullptr);
  if (lin_stage_x(dtype, K, x_f32, in_drop_p)) k.take_bytes((size_t)M * Kp * mt_esize(dtype));
  if (dtype == MT_BF16) k.take_bytes((size_t)N * Kp * 2);
  k.take_bytes(256);
  return k.total();
}

int mt_linear_fwd(int dtype, int M, int N, int K, const void* x, int x_f32, const float* W, const float* b, void* y, int y_f32,
                  int act, const float* rowmask, float in_drop_p, uint64_t seed, uint32_t site, void* ws, size_t ws_bytes,
                  void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !x || !W || !y) return MT_ERR_ARG;
  if (dtype != MT_F32 && dtype != MT_BF16) return MT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool lp = dtype == MT_BF16;
  if (!ws || ws_bytes < mt_linear_ws_bytes(dtype, M, N, K, x_f32, in_drop_p)) return MT_ERR_WS;
  const int Kp = lin_kp(dtype, K);
  WsCarver k(ws);
  const void* xa = x; const void* wa = W;
  if (lin_stage_x(dtype, K, x_f32, in_drop_p)) {
    void* xs = k.take_bytes((size_t)M * Kp * mt_esize(dtype));
    MT_TRY(mt_cast2d_run(x, lp && !x_f32, K, xs, lp, Kp, M, K, mt_make_drop(in_drop_p, seed, site), st));
    xa = xs;
  }
  if (lp) {
    void* wsb = k.take_bytes((size_t)N * Kp * 2);
    MT_TRY(mt_cast2d_run(W, false, K, wsb, true, Kp, N, K, mt_make_drop(0.f, 0, 0), st));
    wa = wsb;
  }
  GemmDesc g = fwd_gemm(M, N, Kp, xa, wa, y, !lp || y_f32);
  g.epi.bias = b; g.epi.act = act; g.epi.rowmask = rowmask;
  return mt_gemm_run(dtype, g, st);
}

/* Modality concat (SFT/models.py:136-138, B2-Trans/models.py:130-132: torch.cat(outputs, dim=2) in front of the fusion / embed Linear):
 * dst[:, off_s : off_s + width_s] = src_s, converting to dst's dtype on the way -- the concatenated operand of the following GEMM is
 * assembled by the library's own strided cast kernel (one launch per modality), no framework copy kernel runs. */
int mt_concat_fwd(int M, int n_src, const void* const* srcs, const int* widths, const int* src_f32, void* dst, int ld_dst, int dst_f32,
                  void* stream) {
  if (M <= 0 || n_src <= 0 || n_src > 8 || !srcs || !widths || !src_f32 || !dst) return MT_ERR_ARG;
  size_t off = 0;
  for (int s = 0; s < n_src; ++s) {
    if (!srcs[s] || widths[s] <= 0 || off + widths[s] > (size_t)ld_dst) return MT_ERR_ARG;
    MT_TRY(mt_cast2d_run(srcs[s], !src_f32[s], widths[s], (char*)dst + off * (dst_f32 ? 4 : 2), !dst_f32, ld_dst, M, widths[s],
                         mt_make_drop(0.f, 0, 0), (cudaStream_t)stream, widths[s]));
    off += widths[s];
  }
  return MT_OK;
}
/* gradient of the concat: dsrc_s = ddst[:, off_s : off_s + width_s] (contiguous per modality, in dsrc's dtype) */
int mt_concat_bwd(int M, int n_src, void* const* dsrcs, const int* widths, const int* dsrc_f32, const void* ddst, int ld, int ddst_f32,
                  void* stream) {
  if (M <= 0 || n_src <= 0 || n_src > 8 || !dsrcs || !widths || !dsrc_f32 || !ddst) return MT_ERR_ARG;
  size_t off = 0;
  for (int s = 0; s < n_src; ++s) {
    if (widths[s] <= 0 || off + widths[s] > (size_t)ld) return MT_ERR_ARG;
    if (dsrcs[s])
      MT_TRY(mt_cast2d_run((const char*)ddst + off * (ddst_f32 ? 4 : 2), !ddst_f32, ld, dsrcs[s], !dsrc_f32[s], widths[s], M, widths[s],
                           mt_make_drop(0.f, 0, 0), (cudaStream_t)stream));
    off += widths[s];
  }
  return MT_OK;
}

size_t mt_linear_bwd_ws_bytes(int dtype, int M, int N, int K, int x_f32, float in_drop_p) {
  WsCarver k(nullptr);
  k.take_bytes((size_t)M * N * mt_esize(dtype));
  k.take_bytes((size_t)M * lin_kp(dtype, K) * mt_esize(dtype));
  if (dtype == MT_BF16) k.take_bytes((size_t)N * K * 2);
  return k.total();
}

/* y: the forward output (needed when act != NONE), stored fp32 when y_f32 else `dtype`; dy likewise per dy_f32. */
int mt_linear_bwd(int dtype, int M, int N, int K, const void* x, int x_f32, const float* W, const void* y, int y_f32, const void* dy,
                  int dy_f32, int act, const float* rowmask, float in_drop_p, uint64_t seed, uint32_t site, void* dx, float* dW,
                  float* db, void* ws, size_t ws_bytes, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !x || !W || !dy || !dW) return MT_ERR_ARG;
  if (dtype != MT_F32 && dtype != MT_BF16) return MT_ERR_ARG;
  if (act != MT_ACT_NONE && !y) return MT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool lp = dtype == MT_BF16;
  if (!ws || ws_bytes < mt_linear_bwd_ws_bytes(dtype, M, N, K, x_f32, in_drop_p)) return MT_ERR_WS;
  WsCarver k(ws);
  // the staged input keeps 16-byte aligned rows (K padded to a multiple of 8 in bf16 mode, zero columns) so that the weight
  // gradient stays on the tcgen05 engine for any K (the 300-wide word vectors): TMA clips the box at the logical width K
  const int Kp = lin_kp(dtype, K);
  void* dz = k.take_bytes((size_t)M * N * mt_esize(dtype));
  void* xs = k.take_bytes((size_t)M * Kp * mt_esize(dtype));
  void* wl = lp ? k.take_bytes((size_t)N * K * 2) : nullptr;
  const bool dy_lp = lp && !dy_f32;
  const void* dzp = dy;
  bool db_done = false;
  if (act != MT_ACT_NONE || rowmask != nullptr || (lp && dy_f32)) {
    MT_TRY(mt_act_bwd_run(M, N, dy, dy_lp, y, lp && !y_f32, act, rowmask, dz, lp, st, db, &db_done));
    dzp = dz;
  }
  const void* xa = x;
  int ldx = K;
  if (in_drop_p > 0.f || (lp && (x_f32 || Kp != K))) {
    MT_TRY(mt_cast2d_run(x, lp && !x_f32, K, xs, lp, Kp, M, K, mt_make_drop(in_drop_p, seed, site), st));
    xa = xs; ldx = Kp;
  }
  MT_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
  MT_TRY(mt_gemm_run(dtype, wgrad_gemm(M, N, K, dzp, N, xa, ldx, dW, K), st));
  if (db && !db_done) MT_TRY(mt_colsum_run(lp, M, N, dzp, N, db, 0, st));
  if (dx) {
    const void* wa = W;
    if (lp) {
      MT_TRY(mt_cast2d_run(W, false, K, wl, true, K, N, K, mt_make_drop(0.f, 0, 0), st));
      wa = wl;
    }
    GemmDesc g = dgrad_gemm(M, N, K, dzp, wa, dx, !lp);
    g.epi.drop = mt_make_drop(in_drop_p, seed, site);      // element index m*K + k == the forward's input-dropout index
    MT_TRY(mt_gemm_run(dtype, g, st));
  }
  return MT_OK;
}

}  // extern "C"
