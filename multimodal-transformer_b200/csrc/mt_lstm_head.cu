// Step-wise LSTM decoder with output feedback + MLP head: SFT/multiTransformer.py:465-483 (NLPTransformer) and
// MFT/multiTransformer.py:357-375 (UniTransformer).  The reference calls nn.LSTM with a length-1 sequence T times,
// feeding i_t = [o_{t-1} || enc_t] where o_{t-1} is the previous hidden state (zeros at t = 0, while the initial
// hidden state itself is the learnable dec_h0).  Here:
//   * the enc half of the input projection is hoisted into one GEMM over all T*B rows,
//   * one persistent kernel runs the recurrence; for t >= 1 the o_{t-1} and h_{t-1} operands are the same vector,
//     so the two recurrent matrices are pre-summed (W_ih[:, :E] + W_hh) and applied once,
//   * backward is a reverse-time persistent kernel + batched wgrad GEMMs.
#include "mt_recurrent.cuh"

GemmDesc mt_wgrad_desc(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw);
// tensor-core / cluster forward (mt_lstm_head_mma.cu), bf16 mode
bool mt_lstm_head_mma_supported(int E, int Hd);
int mt_lstm_head_mma_recurrence(int B, int T, bool training, const float* w_ih, const float* w_hh, const float* b_hh, const float* h0,
                                const float* c0, float* gates, float* hprev, float* oprev, float* cprev, float* hcur, void* hall,
                                cudaStream_t st);
int mt_lstm_head_out_run(int M, int Hd, const float* oh, const float* w2, const float* b2, const float* mask, float* out, cudaStream_t st);
static int g_head_force_ffma = 0;

namespace {

using namespace mtrec;

struct HOff {
  size_t w_ih, w_hh, b_ih, b_hh, h0, c0, w0, b0, w2, b2, total;
};
HOff offsets(int E, int Hd) {
  HOff o;
  size_t p = 0;
  o.w_ih = p; p += (size_t)4 * E * 2 * E;
  o.w_hh = p; p += (size_t)4 * E * E;
  o.b_ih = p; p += 4 * E;
  o.b_hh = p; p += 4 * E;
  o.h0 = p; p += E;
  o.c0 = p; p += E;
  o.w0 = p; p += (size_t)Hd * E;
  o.b0 = p; p += Hd;
  o.w2 = p; p += Hd;
  o.b2 = p; p += 1;
  o.total = p;
  return o;
}

struct HStash {
  void* t_hh;      // [E][4E]  W_hh^T                         (WT)
  void* t_sum;     // [E][4E]  (W_ih[:, :E] + W_hh)^T         (WT)
  void* t_w0;      // [E][Hd]  out.0.weight^T                 (WT)
  void* r_sum;     // [4E][E]  W_ih[:, :E] + W_hh  row-major  (WT, backward)
  float* gates;    // [M,4E]
  float* hprev;    // [M,E]   h_{t-1} (dec_h0 at t = 0)
  float* oprev;    // [M,E]   o_{t-1} (zeros at t = 0)
  float* cprev;    // [M,E]
  float* hcur;     // [M,E]
  float* oh;       // [M,Hd]
  float* dz;       // [M,4E]
  float* doh;      // [M,Hd]
  float* dyv;      // [M]
  void* hall;      // [M,E] bf16 h_t: operand of the hoisted head (tensor-core forward)
  float* oh_fwd;   // [M,Hd] relu(head layer 0): = oh when training, scratch otherwise
  size_t bytes;
};

void carve(const MtLstmHeadCfg& c, void* ws, HStash& s) {
  const size_t M = (size_t)c.B * c.T, E = c.E, Hd = c.Hd, es = mt_esize(c.dtype);
  WsCarver k(ws);
  s.t_hh = k.take_bytes(4 * E * E * es);
  s.t_sum = k.take_bytes(4 * E * E * es);
  s.t_w0 = k.take_bytes(E * Hd * es);
  s.r_sum = k.take_bytes(4 * E * E * es);
  s.gates = k.take<float>(M * 4 * E);
  if (c.training) {
    s.hprev = k.take<float>(M * E); s.oprev = k.take<float>(M * E); s.cprev = k.take<float>(M * E);
    s.hcur = k.take<float>(M * E); s.oh = k.take<float>(M * Hd);
    s.dz = k.take<float>(M * 4 * E); s.doh = k.take<float>(M * Hd); s.dyv = k.take<float>(M);
  } else {
    s.hprev = s.oprev = s.cprev = s.hcur = s.oh = s.dz = s.doh = s.dyv = nullptr;
  }
  s.hall = k.take_bytes(M * E * 2);
  s.oh_fwd = c.training ? s.oh : k.take<float>(M * Hd);
  s.bytes = k.total();
}

struct HArgs {
  HOff O;
  HStash S;
  const float* params;
  const void* params_w;   // backward: flat params in WT
  const float* mask;
  float* out;
  const float* dout;
  float* grads;
  int B, T, E, Hd;
  StreamTable tab;
};

// pack: t_hh = W_hh^T, t_sum = (W_a + W_hh)^T, r_sum = W_a + W_hh, t_w0 = W_0^T
template <typename WT>
__global__ void head_pack_kernel(HArgs a) {
  const int E = a.E, Hd = a.Hd;
  const float* w_ih = a.params + a.O.w_ih;
  const float* w_hh = a.params + a.O.w_hh;
  const float* w0 = a.params + a.O.w0;
  WT* t_hh = (WT*)a.S.t_hh; WT* t_sum = (WT*)a.S.t_sum; WT* r_sum = (WT*)a.S.r_sum; WT* t_w0 = (WT*)a.S.t_w0;
  const int n1 = 4 * E * E;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n1 + E * Hd; e += gridDim.x * blockDim.x) {
    if (e < n1) {
      const int k = e / (4 * E), n = e % (4 * E);      // destination [k][n]
      const float hh = w_hh[(size_t)n * E + k], aa = w_ih[(size_t)n * 2 * E + k];
      t_hh[e] = from_f<WT>(hh);
      t_sum[e] = from_f<WT>(hh + aa);
      r_sum[(size_t)n * E + k] = from_f<WT>(hh + aa);
    } else {
      const int q = e - n1;
      const int k = q / Hd, n = q % Hd;
      t_w0[q] = from_f<WT>(w0[(size_t)n * E + k]);
    }
  }
}

template <bool STREAM, typename WT>
__global__ void __launch_bounds__(NTHREADS + 32, 1) head_fwd_kernel(const __grid_constant__ HArgs a) {
  extern __shared__ __align__(128) float smem[];
  const int E = a.E, Hd = a.Hd;
  float* h = smem + (STREAM ? RING_BYTES / sizeof(float) : 0); float* c = h + E * BT; float* z = c + E * BT; float* oh = z + 4 * E * BT; float* zero = oh + Hd * BT;
  float* part = zero + E * BT;
  __shared__ long long rows[BT];
  // weight ring (STREAM): barriers + slots live at the start of the dynamic shared memory block
  WRing ring = ring_setup(reinterpret_cast<uint8_t*>(smem), STREAM && threadIdx.x == 0);
  if (STREAM) {
    __syncthreads();                              // the only CTA-wide barrier that includes the producer warp
    if (threadIdx.x >= NTHREADS) {
      if (threadIdx.x == NTHREADS) stream_producer<WT>(a.tab, a.T, 0, ring);
      return;
    }
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  const float* P = a.params;
  for (int e = tid; e < E * BT; e += NTHREADS) {
    h[e] = P[a.O.h0 + e / BT]; c[e] = P[a.O.c0 + e / BT]; zero[e] = 0.f;
  }
  cta_sync();
  for (int t = 0; t < a.T; ++t) {
    if (tid < BT) rows[tid] = (long long)(b0 + min(tid, nb - 1)) * a.T + t;
    cta_sync();
    load_rows(z, 4 * E, a.S.gates, rows, nb);
    stash_rows(a.S.hprev, E, h, rows, nb);
    stash_rows(a.S.oprev, E, t == 0 ? zero : h, rows, nb);
    stash_rows(a.S.cprev, E, c, rows, nb);
    cta_sync();
    const WT* Wr = reinterpret_cast<const WT*>(t == 0 ? a.S.t_hh : a.S.t_sum);
    dense<STREAM, WT>(ring, Wr, E, 4 * E, h, part, [&](int n, float* acc) {
      const float bias = P[a.O.b_hh + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) z[n * BT + b] += acc[b] + bias;
    });
    cta_sync();
    for (int e = tid; e < E * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      float gi = sigmoidf_(z[(0 * E + j) * BT + b]), gf = sigmoidf_(z[(1 * E + j) * BT + b]);
      float gg = tanhf(z[(2 * E + j) * BT + b]), go = sigmoidf_(z[(3 * E + j) * BT + b]);
      float cn = gf * c[e] + gi * gg;
      z[(0 * E + j) * BT + b] = gi; z[(1 * E + j) * BT + b] = gf; z[(2 * E + j) * BT + b] = gg; z[(3 * E + j) * BT + b] = go;
      c[e] = cn;
      h[e] = go * tanhf(cn);
    }
    cta_sync();
    stash_rows(a.S.hcur ? a.S.gates : nullptr, 4 * E, z, rows, nb);
    stash_rows(a.S.hcur, E, h, rows, nb);
    dense<STREAM, WT>(ring, reinterpret_cast<const WT*>(a.S.t_w0), E, Hd, h, part, [&](int n, float* acc) {
      const float bias = P[a.O.b0 + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) oh[n * BT + b] = fmaxf(acc[b] + bias, 0.f);
    });
    cta_sync();
    stash_rows(a.S.oh, Hd, oh, rows, nb);
    if (warp < BT) {
      float acc = 0.f;
      for (int j = lane; j < Hd; j += 32) acc = fmaf(oh[j * BT + warp], P[a.O.w2 + j], acc);
      acc = warp_sum(acc);
      if (lane == 0 && warp < nb) {
        const size_t r = (size_t)(b0 + warp) * a.T + t;
        float y = acc + P[a.O.b2];
        if (a.mask) y *= a.mask[r];
        a.out[r] = y;
      }
    }
    cta_sync();
  }
}

template <bool STREAM, typename WT>
__global__ void __launch_bounds__(NTHREADS + 32, 1) head_bwd_kernel(const __grid_constant__ HArgs a) {
  extern __shared__ __align__(128) float smem[];
  const int E = a.E, Hd = a.Hd;
  float* dh = smem + (STREAM ? RING_BYTES / sizeof(float) : 0); float* dc = dh + E * BT; float* dhp = dc + E * BT; float* gates = dhp + E * BT; float* dz = gates + 4 * E * BT;
  float* cprev = dz + 4 * E * BT; float* oh = cprev + E * BT; float* doh = oh + Hd * BT; float* part = doh + Hd * BT;
  __shared__ long long rows[BT];
  __shared__ float dyv[BT];
  // weight ring (STREAM): barriers + slots live at the start of the dynamic shared memory block
  WRing ring = ring_setup(reinterpret_cast<uint8_t*>(smem), STREAM && threadIdx.x == 0);
  if (STREAM) {
    __syncthreads();                              // the only CTA-wide barrier that includes the producer warp
    if (threadIdx.x >= NTHREADS) {
      if (threadIdx.x == NTHREADS) stream_producer<WT>(a.tab, a.T, a.T - 1, ring);
      return;
    }
  }
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  const float* P = a.params;
  const WT* W = reinterpret_cast<const WT*>(a.params_w);
  for (int e = tid; e < E * BT; e += NTHREADS) { dh[e] = 0.f; dc[e] = 0.f; }
  cta_sync();
  for (int t = a.T - 1; t >= 0; --t) {
    if (tid < BT) {
      const size_t r = (size_t)(b0 + min(tid, nb - 1)) * a.T + t;
      rows[tid] = (long long)r;
      float g = tid < nb ? a.dout[r] : 0.f;
      if (a.mask) g *= a.mask[r];
      dyv[tid] = g;
      if (tid < nb) a.S.dyv[r] = g;
    }
    cta_sync();
    load_rows(gates, 4 * E, a.S.gates, rows, nb);
    load_rows(cprev, E, a.S.cprev, rows, nb);
    load_rows(oh, Hd, a.S.oh, rows, nb);
    cta_sync();
    for (int e = tid; e < Hd * BT; e += NTHREADS) doh[e] = oh[e] > 0.f ? dyv[e % BT] * P[a.O.w2 + e / BT] : 0.f;
    cta_sync();
    stash_rows(a.S.doh, Hd, doh, rows, nb);
    dense<STREAM, WT>(ring, W + a.O.w0, Hd, E, doh, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dh[n * BT + b] += acc[b];
    });
    cta_sync();
    for (int e = tid; e < E * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const float gi = gates[(0 * E + j) * BT + b], gf = gates[(1 * E + j) * BT + b];
      const float gg = gates[(2 * E + j) * BT + b], go = gates[(3 * E + j) * BT + b];
      const float cn = gf * cprev[e] + gi * gg;
      const float tc = tanhf(cn);
      const float dhv = dh[e];
      const float dcv = dc[e] + dhv * go * (1.f - tc * tc);
      dz[(0 * E + j) * BT + b] = dcv * gg * gi * (1.f - gi);
      dz[(1 * E + j) * BT + b] = dcv * cprev[e] * gf * (1.f - gf);
      dz[(2 * E + j) * BT + b] = dcv * gi * (1.f - gg * gg);
      dz[(3 * E + j) * BT + b] = dhv * tc * go * (1.f - go);
      dc[e] = dcv * gf;
    }
    cta_sync();
    stash_rows(a.S.dz, 4 * E, dz, rows, nb);
    // gradient wrt h_{t-1}: through W_hh always, and through W_ih[:, :E] (as o_{t-1}) for t >= 1
    const WT* Wr = t == 0 ? W + a.O.w_hh : reinterpret_cast<const WT*>(a.S.r_sum);
    dense<STREAM, WT>(ring, Wr, 4 * E, E, dz, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dhp[n * BT + b] = acc[b];
    });
    cta_sync();
    for (int e = tid; e < E * BT; e += NTHREADS) dh[e] = dhp[e];
    cta_sync();
  }
  // dec_h0 / dec_c0 are broadcast over the batch (:465-466): their gradient is the batch sum of the t = 0 carries
  for (int j = tid; j < E; j += NTHREADS) {
    float sh = 0.f, sc = 0.f;
    for (int b = 0; b < nb; ++b) { sh += dh[j * BT + b]; sc += dc[j * BT + b]; }
    atomicAdd(a.grads + a.O.h0 + j, sh);
    atomicAdd(a.grads + a.O.c0 + j, sc);
  }
}

int check_cfg(const MtLstmHeadCfg* c) {
  if (!c || c->B <= 0 || c->T <= 0 || c->E <= 0 || c->Hd <= 0 || c->E % 4 != 0 || c->Hd % 4 != 0) return MT_ERR_ARG;
  if (c->dtype != MT_F32 && c->dtype != MT_BF16) return MT_ERR_ARG;
  return MT_OK;
}

}  // namespace

extern "C" {

size_t mt_lstm_head_param_count(const MtLstmHeadCfg* cfg) { return check_cfg(cfg) == MT_OK ? offsets(cfg->E, cfg->Hd).total : 0; }

size_t mt_lstm_head_ws_bytes(const MtLstmHeadCfg* cfg) {
  if (check_cfg(cfg) != MT_OK) return 0;
  HStash s;
  carve(*cfg, nullptr, s);
  return s.bytes;
}

int mt_lstm_head_fwd(const MtLstmHeadCfg* cfg, const float* params, const void* params_lp, const void* enc, const float* mask,
                     float* out, void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  const MtLstmHeadCfg& c = *cfg;
  const bool lp = c.dtype == MT_BF16;
  if (!params || !enc || !out || !ws || (lp && !params_lp)) return MT_ERR_ARG;
  HArgs a;
  a.O = offsets(c.E, c.Hd);
  carve(c, ws, a.S);
  if (ws_bytes < a.S.bytes) return MT_ERR_WS;
  cudaStream_t st = (cudaStream_t)stream;
  a.params = params; a.params_w = nullptr; a.mask = mask; a.out = out; a.dout = nullptr; a.grads = nullptr;
  a.B = c.B; a.T = c.T; a.E = c.E; a.Hd = c.Hd;
  const int M = c.B * c.T, E = c.E;
  if (lp) head_pack_kernel<bf16><<<148, 256, 0, st>>>(a); else head_pack_kernel<float><<<148, 256, 0, st>>>(a);
  MT_LAUNCH_CHECK();
  // hoisted enc half of the input projection: gates = enc W_ih[:, E:]^T + b_ih
  GemmDesc g;
  g.M = M; g.N = 4 * E; g.K = E;
  g.A = enc; g.lda = E; g.a_kmajor = true;
  g.B = lp ? (const void*)((const bf16*)params_lp + a.O.w_ih + E) : (const void*)(params + a.O.w_ih + E);
  g.ldb = 2 * E; g.b_kmajor = true;
  g.C = a.S.gates; g.ldc = 4 * E; g.c_f32 = true;
  g.epi.bias = params + a.O.b_ih;
  MT_TRY(mt_gemm_run(c.dtype, g, st));
  if (lp && !g_head_force_ffma && mt_lstm_head_mma_supported(E, c.Hd)) {
    // bf16 mode: recurrence on the cluster / mma.sync kernel, MLP head hoisted into one GEMM + a row-wise dot product
    MT_TRY(mt_lstm_head_mma_recurrence(c.B, c.T, c.training != 0, params + a.O.w_ih, params + a.O.w_hh, params + a.O.b_hh, params + a.O.h0,
                                       params + a.O.c0, a.S.gates, a.S.hprev, a.S.oprev, a.S.cprev, a.S.hcur, a.S.hall, st));
    GemmDesc h;
    h.M = M; h.N = c.Hd; h.K = E;
    h.A = a.S.hall; h.lda = E; h.a_kmajor = true;
    h.B = (const bf16*)params_lp + a.O.w0; h.ldb = E; h.b_kmajor = true;
    h.C = a.S.oh_fwd; h.ldc = c.Hd; h.c_f32 = true;
    h.epi.bias = params + a.O.b0; h.epi.act = MT_ACT_RELU;
    MT_TRY(mt_gemm_run(c.dtype, h, st));
    return mt_lstm_head_out_run(M, c.Hd, a.S.oh_fwd, params + a.O.w2, params + a.O.b2, mask, out, st);
  }
  const size_t smem = ((size_t)(E * 3 + 4 * E + c.Hd) * BT + (size_t)PART_FLOATS) * sizeof(float) + RING_BYTES;
  const int grid = (c.B + BT - 1) / BT;
  a.tab.n = 2;
  a.tab.L[0] = StreamLayer{a.S.t_sum, a.S.t_hh, E, 4 * E};          // step 0 applies W_hh only (o_{-1} = 0)
  a.tab.L[1] = StreamLayer{a.S.t_w0, nullptr, E, c.Hd};
#define MT_HEAD_LAUNCH(KERNEL, WT_)                                                       \
  do {                                                                                     \
    if (stream_table_ok<WT_>(a.tab)) {                                                     \
      MT_TRY(set_smem(KERNEL<true, WT_>, smem));                                           \
      KERNEL<true, WT_><<<grid, NTHREADS + 32, smem, st>>>(a);                             \
    } else {                                                                               \
      MT_TRY(set_smem(KERNEL<false, WT_>, smem));                                          \
      KERNEL<false, WT_><<<grid, NTHREADS, smem, st>>>(a);                                 \
    }                                                                                      \
  } while (0)
  if (lp) MT_HEAD_LAUNCH(head_fwd_kernel, bf16); else MT_HEAD_LAUNCH(head_fwd_kernel, float);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_lstm_head_bwd(const MtLstmHeadCfg* cfg, const float* params, const void* params_lp, const void* enc, const float* mask,
                     const float* dout, void* denc, float* grads, void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  const MtLstmHeadCfg& c = *cfg;
  const bool lp = c.dtype == MT_BF16;
  if (!c.training) return MT_ERR_ARG;
  if (!params || !enc || !dout || !grads || !ws || (lp && !params_lp)) return MT_ERR_ARG;
  HArgs a;
  a.O = offsets(c.E, c.Hd);
  carve(c, ws, a.S);
  if (ws_bytes < a.S.bytes) return MT_ERR_WS;
  cudaStream_t st = (cudaStream_t)stream;
  a.params = params; a.params_w = lp ? params_lp : (const void*)params; a.mask = mask; a.out = nullptr; a.dout = dout; a.grads = grads;
  a.B = c.B; a.T = c.T; a.E = c.E; a.Hd = c.Hd;
  const int M = c.B * c.T, E = c.E, Hd = c.Hd;
  MT_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * a.O.total, st));
  const size_t smem = ((size_t)(E * 4 + 8 * E + 2 * Hd) * BT + (size_t)PART_FLOATS) * sizeof(float) + RING_BYTES;
  const int grid = (c.B + BT - 1) / BT;
  {
    const size_t wsz = mt_esize(c.dtype);
    const char* pw = reinterpret_cast<const char*>(a.params_w);
    a.tab.n = 2;
    a.tab.L[0] = StreamLayer{pw + a.O.w0 * wsz, nullptr, Hd, E};
    a.tab.L[1] = StreamLayer{a.S.r_sum, pw + a.O.w_hh * wsz, 4 * E, E};    // last iteration (t = 0): W_hh only
  }
  if (lp) MT_HEAD_LAUNCH(head_bwd_kernel, bf16); else MT_HEAD_LAUNCH(head_bwd_kernel, float);
  MT_LAUNCH_CHECK();
  float* G = grads;
  const HStash& S = a.S;
  // the enc operand of the input-projection wgrad is `dtype`; the fp32 engine needs fp32 operands on both sides
  const float* encf = (const float*)enc;
  if (lp) {
    // reuse the (dead) oh-gradient-free region: hcur is still needed, so stage enc into dz's sibling buffer cprev
    MT_TRY(mt_cast2d_run(enc, true, E, S.cprev, false, E, M, E, mt_make_drop(0.f, 0, 0), st));
    encf = S.cprev;
  }
  MT_TRY(mt_gemm_run(MT_F32, mt_wgrad_desc(M, 4 * E, E, S.dz, 4 * E, S.oprev, E, G + a.O.w_ih, 2 * E), st));
  MT_TRY(mt_gemm_run(MT_F32, mt_wgrad_desc(M, 4 * E, E, S.dz, 4 * E, encf, E, G + a.O.w_ih + E, 2 * E), st));
  MT_TRY(mt_gemm_run(MT_F32, mt_wgrad_desc(M, 4 * E, E, S.dz, 4 * E, S.hprev, E, G + a.O.w_hh, E), st));
  MT_TRY(mt_colsum_run(0, M, 4 * E, S.dz, 4 * E, G + a.O.b_ih, 1, st));
  MT_TRY(mt_colsum_run(0, M, 4 * E, S.dz, 4 * E, G + a.O.b_hh, 1, st));
  MT_TRY(mt_gemm_run(MT_F32, mt_wgrad_desc(M, Hd, E, S.doh, Hd, S.hcur, E, G + a.O.w0, E), st));
  MT_TRY(mt_colsum_run(0, M, Hd, S.doh, Hd, G + a.O.b0, 1, st));
  MT_TRY(mt_gemm_run(MT_F32, mt_wgrad_desc(M, 1, Hd, S.dyv, 1, S.oh, Hd, G + a.O.w2, Hd), st));
  MT_TRY(mt_colsum_run(0, M, 1, S.dyv, 1, G + a.O.b2, 1, st));
  if (denc) {
    GemmDesc g;
    g.M = M; g.N = E; g.K = 4 * E;
    g.A = S.dz; g.lda = 4 * E; g.a_kmajor = true;
    g.B = params + a.O.w_ih + E; g.ldb = 2 * E; g.b_kmajor = false;
    if (lp) {
      g.C = S.hprev; g.ldc = E; g.c_f32 = true;          // hprev is dead after its wgrad above
      MT_TRY(mt_gemm_run(MT_F32, g, st));
      MT_TRY(mt_cast2d_run(S.hprev, false, E, denc, true, E, M, E, mt_make_drop(0.f, 0, 0), st));
    } else {
      g.C = denc; g.ldc = E; g.c_f32 = true;
      MT_TRY(mt_gemm_run(MT_F32, g, st));
    }
  }
  return MT_OK;
}

/* test hook: keep the bf16 decoder forward on the FFMA kernel (A/B against the cluster / tensor-core forward); returns the previous setting */
int mt_lstm_head_force_ffma(int on) { int old = g_head_force_ffma; g_head_force_ffma = on; return old; }

}  // extern "C"
