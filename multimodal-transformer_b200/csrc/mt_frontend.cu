// Window front-end of MultiCNNTransformer: per modality CNN (Conv1d over the K vectors of a window -> global max-pool) -> Highway ->
// dropout, for ALL B*T windows of a batch at once.  Reference: CNN MFT/models.py:57-79, Highway :27-55, the per-narrative python loop
// :117-132 (B x mods iterations per forward).
//
// Conv1d(D -> E, kernel k) over the K vectors of one window is a GEMM whose A rows are the k consecutive vectors x[j..j+k-1] -- which
// are CONTIGUOUS in the [n_win, K, D] input.  So A is the flat [R = n_win*K, D] matrix read with row stride D and row LENGTH k*D
// (overlapping rows; TMA tensor maps and the FFMA engine both take row stride < row length), no im2col copy.  Rows whose window would
// straddle two windows (j > K-k) are computed and ignored by the pool.
//   fwd:  xs = bf16(x) (bf16 mode) -> Y[R,E] = xs_windows Wp^T -> c = max_j Y + b (argmax kept) -> PG = c [Wproj;Wgate]^T + b
//         -> out = dropout(g*p + (1-g)*c)
//   bwd:  dPG, dc_direct from (dout, PG, c) -> dHW = dPG^T c, db = colsum(dPG), dc = dc_direct + dPG HW -> dconv_b = colsum(dc)
//         -> dY = scatter(dc at argmax) -> dWp = dY^T xs_windows -> conv weight layout [E,D,k]
#include "mt_ops.cuh"

GemmDesc mt_wgrad_desc(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw);

namespace {

struct FrontDims {
  int R;        // n_win * K rows of the flat input
  int L;        // K - k + 1 conv positions per window
  int Mg;       // R - (k - 1): GEMM rows (the last k-1 rows have no complete window)
  int ldA;      // row stride of the GEMM's A operand (D, padded to 8 in bf16 mode)
  int Kg;       // k * ldA
  int Ey;       // E rounded up (8 in bf16 mode, 4 in fp32 mode): leading dimension of Y / c-staged / half of PG
  bool lp;
};

FrontDims front_dims(const MtWindowCnnCfg& c) {
  FrontDims d;
  d.lp = c.dtype == MT_BF16;
  d.R = c.n_win * c.K;
  d.L = c.K - c.k + 1;
  d.Mg = d.R - (c.k - 1);
  d.ldA = d.lp ? (c.D + 7) / 8 * 8 : c.D;
  d.Kg = c.k * d.ldA;
  d.Ey = d.lp ? (c.E + 7) / 8 * 8 : (c.E + 3) / 4 * 4;
  return d;
}

struct FrontWs {
  void* xs;       // [R, ldA] bf16 staged input (bf16 mode only)
  void* wp;       // [E, Kg] packed conv weight, operand dtype
  void* Y;        // [R, Ey] conv outputs (forward) / their gradient (backward), operand dtype
  float* c;       // [n_win, E] pooled CNN output (+ bias), fp32
  void* cs;       // [n_win, Ey] the same in the operand dtype, zero padded (A operand of the Highway GEMM); fp32 mode: padded copy
  unsigned short* arg;   // [n_win, E] arg-max position
  void* hw;       // [2*Ey, Ey] stacked Highway weights (projection rows 0.., gate rows Ey..), operand dtype, zero padded
  float* hb;      // [2*Ey] stacked biases
  void* pg;       // [n_win, 2*Ey] projection | gate pre-activations, operand dtype
  // backward only
  void* dpg;      // [n_win, 2*Ey]
  float* dc;      // [n_win, E] direct path gradient
  float* dct;     // [n_win, E] total gradient of c
  float* dhw;     // [2*Ey, Ey]
  float* dwp;     // [E, Kg]
  size_t bytes;
};

void carve(const MtWindowCnnCfg& c, const FrontDims& d, void* ws, FrontWs& w) {
  const size_t es = mt_esize(c.dtype);
  WsCarver k(ws);
  const bool conv = (c.stages & 1) != 0, hwy = (c.stages & 2) != 0;
  w.xs = (conv && d.lp) ? k.take_bytes((size_t)d.R * d.ldA * 2) : nullptr;
  w.wp = conv ? k.take_bytes((size_t)c.E * d.Kg * es) : nullptr;
  w.Y = conv ? k.take_bytes((size_t)d.R * d.Ey * es) : nullptr;
  w.c = conv ? k.take<float>((size_t)c.n_win * c.E) : nullptr;
  w.arg = conv ? k.take<unsigned short>((size_t)c.n_win * c.E) : nullptr;
  w.cs = hwy ? k.take_bytes((size_t)c.n_win * d.Ey * es) : nullptr;
  w.hw = hwy ? k.take_bytes((size_t)2 * d.Ey * d.Ey * es) : nullptr;
  w.hb = hwy ? k.take<float>((size_t)2 * d.Ey) : nullptr;
  w.pg = hwy ? k.take_bytes((size_t)c.n_win * 2 * d.Ey * es) : nullptr;
  if (c.training) {
    w.dpg = hwy ? k.take_bytes((size_t)c.n_win * 2 * d.Ey * es) : nullptr;
    w.dc = hwy ? k.take<float>((size_t)c.n_win * c.E) : nullptr;
    w.dct = hwy ? k.take<float>((size_t)c.n_win * c.E) : nullptr;
    w.dhw = hwy ? k.take<float>((size_t)2 * d.Ey * d.Ey) : nullptr;
    w.dwp = conv ? k.take<float>((size_t)c.E * d.Kg) : nullptr;
  } else {
    w.dpg = nullptr; w.dc = w.dct = w.dhw = w.dwp = nullptr;
  }
  w.bytes = k.total();
}

int check_cfg(const MtWindowCnnCfg* c) {
  if (!c) return MT_ERR_ARG;
  if (c->dtype != MT_F32 && c->dtype != MT_BF16) return MT_ERR_ARG;
  if (c->n_win <= 0 || c->E <= 0 || (c->stages & 3) == 0 || (c->stages & ~3)) return MT_ERR_ARG;
  if (c->stages & 1) {
    if (c->K <= 0 || c->D <= 0 || c->k <= 0 || c->k > c->K || c->K > 65535) return MT_ERR_ARG;
    const FrontDims d = front_dims(*c);
    const size_t widest = (size_t)(d.Kg > d.Ey ? d.Kg : d.Ey);
    if ((size_t)d.R * widest > 0x7fffffffull) return MT_ERR_ARG;      // 32-bit row*ld products inside the GEMMs
  }
  if ((size_t)c->n_win * 2 * (size_t)(c->E + 8) > 0x7fffffffull) return MT_ERR_ARG;
  if (c->dropout_p < 0.f || c->dropout_p >= 1.f) return MT_ERR_ARG;
  return MT_OK;
}

// ---- conv weight [E, D, k]  <->  GEMM layout [E, k*ldA] (zero padded) ------------------------------------------------------------
template <typename T>
__global__ void conv_w_pack_kernel(const float* __restrict__ w, T* __restrict__ wp, int E, int D, int k, int ldA) {
  const int Kg = k * ldA;
  const size_t n = (size_t)E * Kg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / Kg), r = (int)(i % Kg), j = r / ldA, dd = r % ldA;
    wp[i] = from_f<T>(dd < D ? w[((size_t)f * D + dd) * k + j] : 0.f);
  }
}
__global__ void conv_w_unpack_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int E, int D, int k, int ldA) {
  const size_t n = (size_t)E * D * k;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % k), dd = (int)((i / k) % D), f = (int)(i / ((size_t)k * D));
    dw[i] = dwp[(size_t)f * k * ldA + (size_t)j * ldA + dd];
  }
}

// ---- stacked Highway operands -------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void hw_pack_kernel(const float* __restrict__ wproj, const float* __restrict__ bproj, const float* __restrict__ wgate,
                               const float* __restrict__ bgate, T* __restrict__ hw, float* __restrict__ hb, int E, int Ey) {
  const size_t n = (size_t)2 * Ey * Ey;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / Ey), col = (int)(i % Ey), half = row / Ey, f = row % Ey;
    float v = 0.f;
    if (f < E && col < E) v = (half ? wgate : wproj)[(size_t)f * E + col];
    hw[i] = from_f<T>(v);
    if (col == 0) hb[row] = f < E ? (half ? bgate : bproj)[f] : 0.f;
  }
}

// ---- global max-pool over the L conv positions of a window (+ bias), 4 features per thread ----------------------------------------
// Y [R, Ey] operand dtype; c [n_win, E] fp32; cs [n_win, Ey] operand dtype zero padded (may be null); arg [n_win, E]
template <typename T>
__global__ void pool_fwd_kernel(const T* __restrict__ Y, const float* __restrict__ bias, float* __restrict__ c, T* __restrict__ cs,
                                unsigned short* __restrict__ arg, int n_win, int K, int L, int E, int Ey) {
  const int G = Ey / 4;
  const size_t n = (size_t)n_win * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i / G), f0 = (int)(i % G) * 4;
    const T* y = Y + (size_t)w * K * Ey + f0;
    float4 best = ld4(y);
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
    for (int j = 1; j < L; ++j) {
      const float4 v = ld4(y + (size_t)j * Ey);
      if (v.x > best.x) { best.x = v.x; a0 = j; }
      if (v.y > best.y) { best.y = v.y; a1 = j; }
      if (v.z > best.z) { best.z = v.z; a2 = j; }
      if (v.w > best.w) { best.w = v.w; a3 = j; }
    }
    float o[4] = {best.x, best.y, best.z, best.w};
    const int a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = f0 + e;
      if (f < E) {
        o[e] += bias[f];
        c[(size_t)w * E + f] = o[e];
        arg[(size_t)w * E + f] = (unsigned short)a[e];
      } else {
        o[e] = 0.f;
      }
    }
    if (cs) st4(cs + (size_t)w * Ey + f0, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// dY[(w*K + j), f] = (j == arg[w,f]) ? dc[w,f] : 0 for every row of the window (rows past the last conv position and padded columns: 0)
template <typename T>
__global__ void pool_bwd_kernel(const float* __restrict__ dc, const unsigned short* __restrict__ arg, T* __restrict__ dY, int n_win, int K,
                                int E, int Ey) {
  const int G = Ey / 4;
  const size_t n = (size_t)n_win * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i / G), f0 = (int)(i % G) * 4;
    float g[4];
    int a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = f0 + e;
      g[e] = f < E ? dc[(size_t)w * E + f] : 0.f;
      a[e] = f < E ? (int)arg[(size_t)w * E + f] : -1;
    }
    T* y = dY + (size_t)w * K * Ey + f0;
    for (int j = 0; j < K; ++j)
      st4(y + (size_t)j * Ey, make_float4(a[0] == j ? g[0] : 0.f, a[1] == j ? g[1] : 0.f, a[2] == j ? g[2] : 0.f, a[3] == j ? g[3] : 0.f));
  }
}

// cs[w, :] = operand-dtype copy of c[w, :E], zero padded to Ey (Highway-only stage: c comes from the caller)
template <typename T>
__global__ void stage_c_kernel(const float* __restrict__ c, T* __restrict__ cs, int n_win, int E, int Ey) {
  const size_t n = (size_t)n_win * Ey;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i / Ey), f = (int)(i % Ey);
    cs[i] = from_f<T>(f < E ? c[(size_t)w * E + f] : 0.f);
  }
}

// ---- Highway combine: out = dropout(g * p + (1 - g) * c),  g = sigmoid(gate)   (MFT/models.py:51-54, dropout :129) --------------------
template <typename T>
__global__ void highway_fwd_kernel(const T* __restrict__ pg, const float* __restrict__ c, float* __restrict__ out, int n_win, int E, int Ey,
                                   DropCfg drop) {
  drop = mt_drop_resolve(drop);
  const int G = Ey / 4;
  const size_t n = (size_t)n_win * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i / G), f0 = (int)(i % G) * 4;
    const float4 p4 = ld4(pg + (size_t)w * 2 * Ey + f0), g4 = ld4(pg + (size_t)w * 2 * Ey + Ey + f0);
    const float p[4] = {p4.x, p4.y, p4.z, p4.w}, gp[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = f0 + e;
      if (f < E) {
        const size_t idx = (size_t)w * E + f;
        const float cv = c[idx], g = sigmoidf_(gp[e]);
        out[idx] = (g * p[e] + (1.f - g) * cv) * mt_drop_factor(drop, idx);
      }
    }
  }
}

// d = dout * dropout factor;  dproj = d * g;  dgate_pre = d * (p - c) * g * (1 - g);  dc_direct = d * (1 - g)
template <typename T>
__global__ void highway_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ pg, const float* __restrict__ c, T* __restrict__ dpg,
                                   float* __restrict__ dc, int n_win, int E, int Ey, DropCfg drop) {
  drop = mt_drop_resolve(drop);
  const int G = Ey / 4;
  const size_t n = (size_t)n_win * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i / G), f0 = (int)(i % G) * 4;
    const float4 p4 = ld4(pg + (size_t)w * 2 * Ey + f0), g4 = ld4(pg + (size_t)w * 2 * Ey + Ey + f0);
    const float p[4] = {p4.x, p4.y, p4.z, p4.w}, gp[4] = {g4.x, g4.y, g4.z, g4.w};
    float dp[4], dg[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = f0 + e;
      dp[e] = dg[e] = 0.f;
      if (f < E) {
        const size_t idx = (size_t)w * E + f;
        const float d = dout[idx] * mt_drop_factor(drop, idx), cv = c[idx], g = sigmoidf_(gp[e]);
        dp[e] = d * g;
        dg[e] = d * (p[e] - cv) * g * (1.f - g);
        dc[idx] = d * (1.f - g);
      }
    }
    st4(dpg + (size_t)w * 2 * Ey + f0, make_float4(dp[0], dp[1], dp[2], dp[3]));
    st4(dpg + (size_t)w * 2 * Ey + Ey + f0, make_float4(dg[0], dg[1], dg[2], dg[3]));
  }
}

inline int fe_grid(size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = 148 * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
int front_fwd(const MtWindowCnnCfg& c, const FrontDims& d, FrontWs& w, const float* x, const float* conv_w, const float* conv_b,
              const float* wproj, const float* bproj, const float* wgate, const float* bgate, float* out, cudaStream_t st) {
  const float* cfeat = x;                                     // Highway-only stage: x IS the pooled feature matrix [n_win, E]
  if (c.stages & 1) {
    const void* A = x;
    if (d.lp) {
      MT_TRY(mt_cast2d_run(x, false, c.D, w.xs, true, d.ldA, d.R, c.D, mt_make_drop(0.f, 0, 0), st));
      A = w.xs;
    }
    conv_w_pack_kernel<T><<<fe_grid((size_t)c.E * d.Kg), 256, 0, st>>>(conv_w, (T*)w.wp, c.E, c.D, c.k, d.ldA);
    MT_LAUNCH_CHECK();
    GemmDesc g;
    g.M = d.Mg; g.N = c.E; g.K = d.Kg;
    g.A = A; g.lda = d.ldA; g.a_kmajor = true;                // overlapping rows: stride ldA, length k * ldA
    g.B = w.wp; g.ldb = d.Kg; g.b_kmajor = true;
    g.C = w.Y; g.ldc = d.Ey; g.c_f32 = !d.lp;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    float* cdst = (c.stages & 2) ? w.c : out;
    mt_prof_work(0.0, (double)c.n_win * d.L * c.E * sizeof(T) + (double)c.n_win * c.E * 6.0);
    pool_fwd_kernel<T><<<fe_grid((size_t)c.n_win * (d.Ey / 4)), 256, 0, st>>>((const T*)w.Y, conv_b, cdst, (c.stages & 2) ? (T*)w.cs : nullptr,
                                                                              w.arg, c.n_win, c.K, d.L, c.E, d.Ey);
    MT_LAUNCH_CHECK();
    cfeat = w.c;
  } else {
    stage_c_kernel<T><<<fe_grid((size_t)c.n_win * d.Ey), 256, 0, st>>>(x, (T*)w.cs, c.n_win, c.E, d.Ey);
    MT_LAUNCH_CHECK();
  }
  if (c.stages & 2) {
    hw_pack_kernel<T><<<fe_grid((size_t)2 * d.Ey * d.Ey), 256, 0, st>>>(wproj, bproj, wgate, bgate, (T*)w.hw, w.hb, c.E, d.Ey);
    MT_LAUNCH_CHECK();
    GemmDesc g;
    g.M = c.n_win; g.N = 2 * d.Ey; g.K = d.Ey;
    g.A = w.cs; g.lda = d.Ey; g.a_kmajor = true;
    g.B = w.hw; g.ldb = d.Ey; g.b_kmajor = true;
    g.C = w.pg; g.ldc = 2 * d.Ey; g.c_f32 = !d.lp;
    g.epi.bias = w.hb;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    const DropCfg drop = mt_make_drop(c.dropout_p, c.seed, c.site);
    mt_prof_work(0.0, (double)c.n_win * c.E * (2.0 * sizeof(T) + 8.0));
    highway_fwd_kernel<T><<<fe_grid((size_t)c.n_win * (d.Ey / 4)), 256, 0, st>>>((const T*)w.pg, cfeat, out, c.n_win, c.E, d.Ey, drop);
    MT_LAUNCH_CHECK();
  }
  return MT_OK;
}

template <typename T>
int front_bwd(const MtWindowCnnCfg& c, const FrontDims& d, FrontWs& w, const float* x, const float* dout, float* dx, float* dconv_w,
              float* dconv_b, float* dwproj, float* dbproj, float* dwgate, float* dbgate, cudaStream_t st) {
  const float* dct = dout;                                    // conv-only stage: dout IS the gradient of the pooled features
  if (c.stages & 2) {
    const float* cfeat = (c.stages & 1) ? w.c : x;
    const DropCfg drop = mt_make_drop(c.dropout_p, c.seed, c.site);
    mt_prof_work(0.0, (double)c.n_win * c.E * (4.0 * sizeof(T) + 12.0));
    highway_bwd_kernel<T><<<fe_grid((size_t)c.n_win * (d.Ey / 4)), 256, 0, st>>>(dout, (const T*)w.pg, cfeat, (T*)w.dpg, w.dc, c.n_win, c.E,
                                                                                 d.Ey, drop);
    MT_LAUNCH_CHECK();
    // weight / bias gradients of the two Highway linears
    MT_CUDA(cudaMemsetAsync(w.dhw, 0, sizeof(float) * (size_t)2 * d.Ey * d.Ey, st));
    MT_TRY(mt_gemm_run(c.dtype, mt_wgrad_desc(c.n_win, 2 * d.Ey, d.Ey, w.dpg, 2 * d.Ey, w.cs, d.Ey, w.dhw, d.Ey), st));
    MT_TRY(mt_cast2d_run(w.dhw, false, d.Ey, dwproj, false, c.E, c.E, c.E, mt_make_drop(0.f, 0, 0), st));
    MT_TRY(mt_cast2d_run(w.dhw + (size_t)d.Ey * d.Ey, false, d.Ey, dwgate, false, c.E, c.E, c.E, mt_make_drop(0.f, 0, 0), st));
    MT_TRY(mt_colsum_run(d.lp, c.n_win, c.E, w.dpg, 2 * d.Ey, dbproj, 0, st));
    MT_TRY(mt_colsum_run(d.lp, c.n_win, c.E, (const char*)w.dpg + (size_t)d.Ey * sizeof(T), 2 * d.Ey, dbgate, 0, st));
    // dc_total = dc_direct + dPG [Wproj; Wgate]
    float* dst = (c.stages & 1) ? w.dct : dx;
    if (dst) {
      GemmDesc g;
      g.M = c.n_win; g.N = c.E; g.K = 2 * d.Ey;
      g.A = w.dpg; g.lda = 2 * d.Ey; g.a_kmajor = true;
      g.B = w.hw; g.ldb = d.Ey; g.b_kmajor = false;           // B(n = input feature, k = stacked output row) = hw[k * Ey + n]
      g.C = dst; g.ldc = c.E; g.c_f32 = true;
      g.epi.residual = w.dc; g.epi.ldr = c.E;
      MT_TRY(mt_gemm_run(c.dtype, g, st));
    }
    dct = dst;
  }
  if (c.stages & 1) {
    MT_TRY(mt_colsum_run(0, c.n_win, c.E, dct, c.E, dconv_b, 0, st));
    mt_prof_work(0.0, (double)d.R * d.Ey * sizeof(T) + (double)c.n_win * c.E * 6.0);
    pool_bwd_kernel<T><<<fe_grid((size_t)c.n_win * (d.Ey / 4)), 256, 0, st>>>(dct, w.arg, (T*)w.Y, c.n_win, c.K, c.E, d.Ey);
    MT_LAUNCH_CHECK();
    const void* A = d.lp ? (const void*)w.xs : (const void*)x;
    MT_CUDA(cudaMemsetAsync(w.dwp, 0, sizeof(float) * (size_t)c.E * d.Kg, st));
    MT_TRY(mt_gemm_run(c.dtype, mt_wgrad_desc(d.Mg, c.E, d.Kg, w.Y, d.Ey, A, d.ldA, w.dwp, d.Kg), st));
    conv_w_unpack_kernel<<<fe_grid((size_t)c.E * c.D * c.k), 256, 0, st>>>(w.dwp, dconv_w, c.E, c.D, c.k, d.ldA);
    MT_LAUNCH_CHECK();
  }
  return MT_OK;
}

}  // namespace

extern "C" {

size_t mt_window_cnn_ws_bytes(const MtWindowCnnCfg* cfg) {
  if (check_cfg(cfg) != MT_OK) return 0;
  FrontWs w;
  carve(*cfg, front_dims(*cfg), nullptr, w);
  return w.bytes;
}

int mt_window_cnn_fwd(const MtWindowCnnCfg* cfg, const float* x, const float* conv_w, const float* conv_b, const float* wproj,
                      const float* bproj, const float* wgate, const float* bgate, float* out, void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  const MtWindowCnnCfg& c = *cfg;
  if (!x || !out) return MT_ERR_ARG;
  if ((c.stages & 1) && (!conv_w || !conv_b)) return MT_ERR_ARG;
  if ((c.stages & 2) && (!wproj || !bproj || !wgate || !bgate)) return MT_ERR_ARG;
  const FrontDims d = front_dims(c);
  FrontWs w;
  carve(c, d, ws, w);
  if (!ws || ws_bytes < w.bytes) return MT_ERR_WS;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.lp) return front_fwd<bf16>(c, d, w, x, conv_w, conv_b, wproj, bproj, wgate, bgate, out, st);
  return front_fwd<float>(c, d, w, x, conv_w, conv_b, wproj, bproj, wgate, bgate, out, st);
}

int mt_window_cnn_bwd(const MtWindowCnnCfg* cfg, const float* x, const float* dout, float* dx, float* dconv_w, float* dconv_b,
                      float* dwproj, float* dbproj, float* dwgate, float* dbgate, void* ws, size_t ws_bytes, void* stream) {
  MT_TRY(check_cfg(cfg));
  const MtWindowCnnCfg& c = *cfg;
  if (!c.training || !x || !dout) return MT_ERR_ARG;
  if ((c.stages & 1) && (!dconv_w || !dconv_b)) return MT_ERR_ARG;
  if ((c.stages & 2) && (!dwproj || !dbproj || !dwgate || !dbgate)) return MT_ERR_ARG;
  const FrontDims d = front_dims(c);
  FrontWs w;
  carve(c, d, ws, w);
  if (!ws || ws_bytes < w.bytes) return MT_ERR_WS;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.lp) return front_bwd<bf16>(c, d, w, x, dout, dx, dconv_w, dconv_b, dwproj, dbproj, dwgate, dbgate, st);
  return front_bwd<float>(c, d, w, x, dout, dx, dconv_w, dconv_b, dwproj, dbproj, dwgate, dbgate, st);
}

}  // extern "C"
