// Memory Fusion Network (MFN.forward MFT/multiTransformer.py:181-248), forward and backward.
//
//  1. the LSTM input projections of all T steps are hoisted into one GEMM per modality        (K13, hoisted part)
//  2. ONE persistent kernel runs the whole time recurrence: a CTA owns a tile of BT narratives, keeps
//     h / c / mem and every per-step activation in shared memory, and walks t = 0..T-1 without leaving the SM;
//     the in-loop weights (transposed pack, L2-resident) stream through the FMA pipes                (K13-K18)
//  3. backward is a second persistent kernel walking t = T-1..0 that carries dh / dc / dmem in shared memory and
//     writes the pre-activation gradients of every in-loop layer; all weight gradients are then batched
//     wgrad GEMMs over the T*B rows (no per-step atomics).
//
// Activations inside a CTA are stored feature-major [feature][BT] so one 16-byte shared load feeds BT FMAs.
// Stash row order follows the inputs: row(b,t) = b*sb + t*st with (sb,st) = (T,1) for [B,T,D] inputs and (1,B)
// for the reference's permuted [T,B,D] views.
#include "mt_recurrent.cuh"

GemmDesc mt_wgrad_desc(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw);

namespace {

using namespace mtrec;

struct LinOff { size_t w, b; };

struct Dims {
  int n_mods;
  int D[MT_MAX_MODS], H[MT_MAX_MODS], hoff[MT_MAX_MODS];
  int Hs, MEM, A1, A2, G, O;
  // fp32 flat parameter offsets
  size_t w_ih[MT_MAX_MODS], w_hh[MT_MAX_MODS], b_ih[MT_MAX_MODS], b_hh[MT_MAX_MODS];
  LinOff att1_fc1, att1_fc2, att2_fc1, att2_fc2, g1_fc1, g1_fc2, g2_fc1, g2_fc2, out_fc1, out_fc2;
  size_t total;
  // transposed forward pack offsets (elements)
  size_t t_hh[MT_MAX_MODS], t_att1_fc1, t_att1_fc2, t_att2_fc1, t_att2_fc2, t_g_fc1, t_g1_fc2, t_g2_fc2, t_out_fc1, t_total;
};

int make_dims(const MtMfnCfg& c, Dims& D) {
  if (c.n_mods < 1 || c.n_mods > MT_MAX_MODS || c.B <= 0 || c.T <= 0) return MT_ERR_ARG;
  D.n_mods = c.n_mods;
  D.Hs = 0;
  for (int m = 0; m < c.n_mods; ++m) {
    if (c.in_dim[m] <= 0 || c.hid[m] <= 0 || c.hid[m] % 4 != 0 || c.in_dim[m] % 4 != 0) return MT_ERR_ARG;
    D.D[m] = c.in_dim[m]; D.H[m] = c.hid[m]; D.hoff[m] = D.Hs; D.Hs += c.hid[m];
  }
  D.MEM = c.mem_dim; D.A1 = c.h_att1; D.A2 = c.h_att2; D.G = c.h_gamma; D.O = c.h_out;
  if (D.MEM <= 0 || D.A1 <= 0 || D.A2 <= 0 || D.G <= 0 || D.O <= 0) return MT_ERR_ARG;
  if ((D.MEM | D.A1 | D.A2 | D.G | D.O) % 4 != 0) return MT_ERR_ARG;
  size_t o = 0;
  for (int m = 0; m < c.n_mods; ++m) {
    D.w_ih[m] = o; o += (size_t)4 * D.H[m] * D.D[m];
    D.w_hh[m] = o; o += (size_t)4 * D.H[m] * D.H[m];
    D.b_ih[m] = o; o += 4 * D.H[m];
    D.b_hh[m] = o; o += 4 * D.H[m];
  }
  auto lin = [&](LinOff& l, int out, int in) { l.w = o; o += (size_t)out * in; l.b = o; o += out; };
  const int H2 = 2 * D.Hs;
  lin(D.att1_fc1, D.A1, H2); lin(D.att1_fc2, H2, D.A1);
  lin(D.att2_fc1, D.A2, H2); lin(D.att2_fc2, D.MEM, D.A2);
  lin(D.g1_fc1, D.G, H2 + D.MEM); lin(D.g1_fc2, D.MEM, D.G);
  lin(D.g2_fc1, D.G, H2 + D.MEM); lin(D.g2_fc2, D.MEM, D.G);
  lin(D.out_fc1, D.O, D.Hs + D.MEM); lin(D.out_fc2, 1, D.O);
  D.total = o;
  size_t t = 0;
  for (int m = 0; m < c.n_mods; ++m) { D.t_hh[m] = t; t += (size_t)4 * D.H[m] * D.H[m]; }
  D.t_att1_fc1 = t; t += (size_t)H2 * D.A1;
  D.t_att1_fc2 = t; t += (size_t)D.A1 * H2;
  D.t_att2_fc1 = t; t += (size_t)H2 * D.A2;
  D.t_att2_fc2 = t; t += (size_t)D.A2 * D.MEM;
  D.t_g_fc1 = t; t += (size_t)(H2 + D.MEM) * 2 * D.G;
  D.t_g1_fc2 = t; t += (size_t)D.G * D.MEM;
  D.t_g2_fc2 = t; t += (size_t)D.G * D.MEM;
  D.t_out_fc1 = t; t += (size_t)(D.Hs + D.MEM) * D.O;
  D.t_total = t;
  return MT_OK;
}

// global stash (fp32, one row per (b,t)); widths in floats
struct Stash {
  float* gates;   // [M,4Hs]  zx before the recurrence; post-activation i,f,g,o after it (per modality: i|f|g|o blocks)
  float* hprev;   // [M,Hs]
  float* cstar;   // [M,2Hs]  c_{t-1} || c_t
  float* a1;      // [M,A1]
  float* att;     // [M,2Hs]  softmax output
  float* both;    // [M,2Hs+MEM]  attended || mem_{t-1}
  float* a2;      // [M,A2]
  float* chat;    // [M,MEM]
  float* gh;      // [M,2G]   gamma1 | gamma2 hidden (post relu, post dropout)
  float* gm;      // [M,2MEM] gamma1 | gamma2
  float* last;    // [M,Hs+MEM]   h_t || mem_t
  float* oh;      // [M,O]    out hidden (post relu, post dropout)
  // backward: pre-activation gradients
  float* dz_lstm; // [M,4Hs]
  float* dlogit;  // [M,2Hs]
  float* da1;     // [M,A1]
  float* dzchat;  // [M,MEM]
  float* da2;     // [M,A2]
  float* dzg;     // [M,2MEM]
  float* dgh;     // [M,2G]
  float* dzoh;    // [M,O]
  float* dyv;     // [M]
  float* xf[MT_MAX_MODS];   // fp32 copies of bf16 inputs
  void* tpack;    // transposed forward weights
  size_t bytes;
};

void carve(const MtMfnCfg& c, const Dims& D, void* ws, Stash& s) {
  const size_t M = (size_t)c.B * c.T;
  const int H2 = 2 * D.Hs;
  WsCarver k(ws);
  s.tpack = k.take_bytes(D.t_total * mt_esize(c.dtype));
  s.gates = k.take<float>(M * 4 * D.Hs);
  s.last = k.take<float>(M * (D.Hs + D.MEM));
  for (int m = 0; m < MT_MAX_MODS; ++m) s.xf[m] = nullptr;
  if (c.dtype == MT_BF16)
    for (int m = 0; m < D.n_mods; ++m) s.xf[m] = k.take<float>(M * D.D[m]);
  if (c.training) {
    s.hprev = k.take<float>(M * D.Hs);
    s.cstar = k.take<float>(M * H2);
    s.a1 = k.take<float>(M * D.A1);
    s.att = k.take<float>(M * H2);
    s.both = k.take<float>(M * (H2 + D.MEM));
    s.a2 = k.take<float>(M * D.A2);
    s.chat = k.take<float>(M * D.MEM);
    s.gh = k.take<float>(M * 2 * D.G);
    s.gm = k.take<float>(M * 2 * D.MEM);
    s.oh = k.take<float>(M * D.O);
    s.dz_lstm = k.take<float>(M * 4 * D.Hs);
    s.dlogit = k.take<float>(M * H2);
    s.da1 = k.take<float>(M * D.A1);
    s.dzchat = k.take<float>(M * D.MEM);
    s.da2 = k.take<float>(M * D.A2);
    s.dzg = k.take<float>(M * 2 * D.MEM);
    s.dgh = k.take<float>(M * 2 * D.G);
    s.dzoh = k.take<float>(M * D.O);
    s.dyv = k.take<float>(M);
  } else {
    s.hprev = s.cstar = s.a1 = s.att = s.both = s.a2 = s.chat = s.gh = s.gm = s.oh = nullptr;
    s.dz_lstm = s.dlogit = s.da1 = s.dzchat = s.da2 = s.dzg = s.dgh = s.dzoh = s.dyv = nullptr;
  }
  s.bytes = k.total();
}

struct KArgs {
  Dims D;
  Stash S;
  const float* params;     // fp32 flat
  const void* wlp;         // forward: transposed pack (WT); backward: flat params in WT
  const float* mask;       // [B,T] or null
  float* out;              // forward: [B,T]
  const float* dout;       // backward: [B,T]
  float* h_last; float* c_last; float* mem_last;
  int B, T;
  long long sb, st;        // row(b,t) = b*sb + t*st
  int training;
  DropCfg drop_g1, drop_g2, drop_out;
  StreamTable tab;         // in-loop weight blocks in consumption order
};

struct SmemFwd {
  float *h, *c, *mem, *z, *cstar, *a1, *att, *both, *a2, *chat, *gh, *gm, *last, *oh, *part;
};

__device__ __forceinline__ int mod_of(const Dims& D, int j) {
  int m = 0;
#pragma unroll
  for (int q = 1; q < MT_MAX_MODS; ++q) if (q < D.n_mods && j >= D.hoff[q]) m = q;
  return m;
}

template <bool STREAM, typename WT>
__global__ void __launch_bounds__(NTHREADS + 32, 1) mfn_fwd_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) float smem[];
  const Dims& D = a.D;
  const int Hs = D.Hs, H2 = 2 * D.Hs, MEM = D.MEM, A1 = D.A1, A2 = D.A2, G = D.G, O = D.O;
  SmemFwd s;
  {
    float* p = smem + (STREAM ? RING_BYTES / sizeof(float) : 0);
    s.h = p; p += Hs * BT; s.c = p; p += Hs * BT; s.mem = p; p += MEM * BT; s.z = p; p += 4 * Hs * BT;
    s.cstar = p; p += H2 * BT; s.a1 = p; p += A1 * BT; s.att = p; p += H2 * BT; s.both = p; p += (H2 + MEM) * BT;
    s.a2 = p; p += A2 * BT; s.chat = p; p += MEM * BT; s.gh = p; p += 2 * G * BT; s.gm = p; p += 2 * MEM * BT;
    s.last = p; p += (Hs + MEM) * BT; s.oh = p; p += O * BT; s.part = p;
  }
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
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, a.B - b0);
  const WT* TP = reinterpret_cast<const WT*>(a.wlp);
  const float* P = a.params;

  for (int e = tid; e < Hs * BT; e += NTHREADS) { s.h[e] = 0.f; s.c[e] = 0.f; }
  for (int e = tid; e < MEM * BT; e += NTHREADS) s.mem[e] = 0.f;
  cta_sync();

  for (int t = 0; t < a.T; ++t) {
    if (tid < BT) rows[tid] = (long long)(b0 + min(tid, nb - 1)) * a.sb + (long long)t * a.st;
    cta_sync();
    // ---- z = zx (hoisted x-projection + both biases) + W_hh h_{t-1} -------------------------------
    load_rows(s.z, 4 * Hs, a.S.gates, rows, nb);
    stash_rows(a.S.hprev, Hs, s.h, rows, nb);
    for (int e = tid; e < Hs * BT; e += NTHREADS) s.cstar[e] = s.c[e];          // c_{t-1} half
    cta_sync();
    for (int m = 0; m < D.n_mods; ++m) {
      const int H = D.H[m];
      float* zm = s.z + 4 * D.hoff[m] * BT;
      const float* bhh = P + D.b_hh[m];
      dense<STREAM, WT>(ring, TP + D.t_hh[m], H, 4 * H, s.h + D.hoff[m] * BT, s.part, [&](int n, float* acc) {
        const float bias = bhh[n];
#pragma unroll
        for (int b = 0; b < BT; ++b) zm[n * BT + b] += acc[b] + bias;
      });
    }
    cta_sync();
    // ---- LSTM gates: c_t = s(f) c + s(i) tanh(g); h_t = s(o) tanh(c_t) ----------------------------
    for (int e = tid; e < Hs * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const int m = mod_of(D, j);
      const int H = D.H[m], jj = j - D.hoff[m];
      float* zm = s.z + 4 * D.hoff[m] * BT;
      float gi = sigmoidf_(zm[(0 * H + jj) * BT + b]);
      float gf = sigmoidf_(zm[(1 * H + jj) * BT + b]);
      float gg = tanhf(zm[(2 * H + jj) * BT + b]);
      float go = sigmoidf_(zm[(3 * H + jj) * BT + b]);
      float cn = gf * s.c[e] + gi * gg;
      float hn = go * tanhf(cn);
      zm[(0 * H + jj) * BT + b] = gi; zm[(1 * H + jj) * BT + b] = gf;
      zm[(2 * H + jj) * BT + b] = gg; zm[(3 * H + jj) * BT + b] = go;
      s.c[e] = cn; s.h[e] = hn;
      s.cstar[(Hs + j) * BT + b] = cn;
      s.last[e] = hn;
    }
    cta_sync();
    if (a.training) { stash_rows(a.S.gates, 4 * Hs, s.z, rows, nb); stash_rows(a.S.cstar, H2, s.cstar, rows, nb); }
    // ---- delta-memory attention over cStar --------------------------------------------------------
    dense<STREAM, WT>(ring, TP + D.t_att1_fc1, H2, A1, s.cstar, s.part, [&](int n, float* acc) {
      const float bias = P[D.att1_fc1.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.a1[n * BT + b] = fmaxf(acc[b] + bias, 0.f);
    });
    cta_sync();
    dense<STREAM, WT>(ring, TP + D.t_att1_fc2, A1, H2, s.a1, s.part, [&](int n, float* acc) {
      const float bias = P[D.att1_fc2.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.att[n * BT + b] = acc[b] + bias;
    });
    cta_sync();
    // softmax over the 2Hs FEATURES of each narrative: warp b handles narrative b
    if (warp < BT) {
      float mx = -INFINITY;
      for (int f = lane; f < H2; f += 32) mx = fmaxf(mx, s.att[f * BT + warp]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int f = lane; f < H2; f += 32) { float e = expf(s.att[f * BT + warp] - mx); s.att[f * BT + warp] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int f = lane; f < H2; f += 32) {
        float p = s.att[f * BT + warp] * inv;
        s.att[f * BT + warp] = p;
        s.both[f * BT + warp] = p * s.cstar[f * BT + warp];          // attended
      }
    }
    for (int e = tid; e < MEM * BT; e += NTHREADS) s.both[H2 * BT + e] = s.mem[e];   // || mem_{t-1}
    cta_sync();
    if (a.training) {
      stash_rows(a.S.a1, A1, s.a1, rows, nb); stash_rows(a.S.att, H2, s.att, rows, nb);
      stash_rows(a.S.both, H2 + MEM, s.both, rows, nb);
    }
    // ---- cHat = tanh(att2(attended)) ; gamma hidden = relu(gamma_fc1(both)) -----------------------
    dense<STREAM, WT>(ring, TP + D.t_att2_fc1, H2, A2, s.both, s.part, [&](int n, float* acc) {
      const float bias = P[D.att2_fc1.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.a2[n * BT + b] = fmaxf(acc[b] + bias, 0.f);
    });
    dense<STREAM, WT>(ring, TP + D.t_g_fc1, H2 + MEM, 2 * G, s.both, s.part, [&](int n, float* acc) {
      const bool second = n >= G;
      const float bias = second ? P[D.g2_fc1.b + n - G] : P[D.g1_fc1.b + n];
      const DropCfg& dc = second ? a.drop_g2 : a.drop_g1;
      const int j = second ? n - G : n;
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = fmaxf(acc[b] + bias, 0.f);
        // element index of the [T,B,G] tensor (oracle/mt_oracle.py:_drop_t)
        v *= mt_drop_factor(dc, ((uint64_t)t * a.B + (uint64_t)(b0 + b)) * (uint64_t)G + (uint64_t)j);
        s.gh[n * BT + b] = v;
      }
    });
    cta_sync();
    dense<STREAM, WT>(ring, TP + D.t_att2_fc2, A2, MEM, s.a2, s.part, [&](int n, float* acc) {
      const float bias = P[D.att2_fc2.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.chat[n * BT + b] = tanhf(acc[b] + bias);
    });
    dense<STREAM, WT>(ring, TP + D.t_g1_fc2, G, MEM, s.gh, s.part, [&](int n, float* acc) {
      const float bias = P[D.g1_fc2.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.gm[n * BT + b] = sigmoidf_(acc[b] + bias);
    });
    dense<STREAM, WT>(ring, TP + D.t_g2_fc2, G, MEM, s.gh + G * BT, s.part, [&](int n, float* acc) {
      const float bias = P[D.g2_fc2.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) s.gm[(MEM + n) * BT + b] = sigmoidf_(acc[b] + bias);
    });
    cta_sync();
    // ---- mem_t = gamma1 * mem_{t-1} + gamma2 * cHat -----------------------------------------------
    for (int e = tid; e < MEM * BT; e += NTHREADS) {
      float mn = s.gm[e] * s.mem[e] + s.gm[MEM * BT + e] * s.chat[e];
      s.mem[e] = mn;
      s.last[Hs * BT + e] = mn;
    }
    cta_sync();
    if (a.training) {
      stash_rows(a.S.a2, A2, s.a2, rows, nb); stash_rows(a.S.chat, MEM, s.chat, rows, nb);
      stash_rows(a.S.gh, 2 * G, s.gh, rows, nb); stash_rows(a.S.gm, 2 * MEM, s.gm, rows, nb);
      stash_rows(a.S.last, Hs + MEM, s.last, rows, nb);
    }
    // ---- output head: y_t = out_fc2(drop(relu(out_fc1([h_t || mem_t])))) * mask ------------------------
    dense<STREAM, WT>(ring, TP + D.t_out_fc1, Hs + MEM, O, s.last, s.part, [&](int n, float* acc) {
      const float bias = P[D.out_fc1.b + n];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = fmaxf(acc[b] + bias, 0.f);
        v *= mt_drop_factor(a.drop_out, ((uint64_t)t * a.B + (uint64_t)(b0 + b)) * (uint64_t)O + (uint64_t)n);
        s.oh[n * BT + b] = v;
      }
    });
    cta_sync();
    if (a.training) stash_rows(a.S.oh, O, s.oh, rows, nb);
    if (warp < BT) {
      float acc = 0.f;
      for (int j = lane; j < O; j += 32) acc = fmaf(s.oh[j * BT + warp], P[D.out_fc2.w + j], acc);
      acc = warp_sum(acc);
      if (lane == 0 && warp < nb) {
        const int b = b0 + warp;
        float y = acc + P[D.out_fc2.b];
        if (a.mask) y *= a.mask[(size_t)b * a.T + t];
        a.out[(size_t)b * a.T + t] = y;
      }
    }
    cta_sync();
  }
  // final state (MFN.h / MFN.c / MFN.mem attributes of the reference module)
  for (int e = tid; e < nb * Hs; e += NTHREADS) {
    int b = e / Hs, j = e % Hs;
    if (a.h_last) a.h_last[(size_t)(b0 + b) * Hs + j] = s.h[j * BT + b];
    if (a.c_last) a.c_last[(size_t)(b0 + b) * Hs + j] = s.c[j * BT + b];
  }
  if (a.mem_last)
    for (int e = tid; e < nb * MEM; e += NTHREADS) { int b = e / MEM, j = e % MEM; a.mem_last[(size_t)(b0 + b) * MEM + j] = s.mem[j * BT + b]; }
}

// ------------------------------------------------------------------------------------------------------
// reverse-time kernel.  Weights are the row-major originals W[out][in] viewed as Wt with K = out, N = in.
// ------------------------------------------------------------------------------------------------------
template <bool STREAM, typename WT>
__global__ void __launch_bounds__(NTHREADS + 32, 1) mfn_bwd_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) float smem[];
  const Dims& D = a.D;
  const int Hs = D.Hs, H2 = 2 * D.Hs, MEM = D.MEM, A1 = D.A1, A2 = D.A2, G = D.G, O = D.O;
  float* p = smem + (STREAM ? RING_BYTES / sizeof(float) : 0);
  float* dh = p; p += Hs * BT;            // carries (gradient wrt h_t, c_t, mem_t arriving from step t+1)
  float* dc = p; p += Hs * BT;
  float* dmem = p; p += MEM * BT;
  float* dhp = p; p += Hs * BT;           // gradient wrt h_{t-1} produced at this step
  float* dzoh = p; p += O * BT;
  float* gm = p; p += 2 * MEM * BT;
  float* dzg = p; p += 2 * MEM * BT;
  float* dgh = p; p += 2 * G * BT;
  float* gh = p; p += 2 * G * BT;
  float* both = p; p += (H2 + MEM) * BT;
  float* dboth = p; p += (H2 + MEM) * BT;
  float* chat = p; p += MEM * BT;
  float* dzchat = p; p += MEM * BT;
  float* a2 = p; p += A2 * BT;
  float* da2 = p; p += A2 * BT;
  float* att = p; p += H2 * BT;
  float* cstar = p; p += H2 * BT;
  float* dlogit = p; p += H2 * BT;
  float* dcstar = p; p += H2 * BT;
  float* a1 = p; p += A1 * BT;
  float* da1 = p; p += A1 * BT;
  float* gates = p; p += 4 * Hs * BT;
  float* dz = p; p += 4 * Hs * BT;
  float* oh = p; p += O * BT;
  float* part = p;
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
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, a.B - b0);
  const WT* W = reinterpret_cast<const WT*>(a.wlp);
  const float* P = a.params;
  const float sc_g = a.drop_g1.scale, sc_o = a.drop_out.scale;

  for (int e = tid; e < Hs * BT; e += NTHREADS) { dh[e] = 0.f; dc[e] = 0.f; }
  for (int e = tid; e < MEM * BT; e += NTHREADS) dmem[e] = 0.f;
  cta_sync();

  for (int t = a.T - 1; t >= 0; --t) {
    if (tid < BT) {
      const int b = b0 + min(tid, nb - 1);
      rows[tid] = (long long)b * a.sb + (long long)t * a.st;
      float g = tid < nb ? a.dout[(size_t)b * a.T + t] : 0.f;
      if (a.mask) g *= a.mask[(size_t)b * a.T + t];
      dyv[tid] = g;
      if (tid < nb) a.S.dyv[rows[tid]] = g;
    }
    cta_sync();
    load_rows(oh, O, a.S.oh, rows, nb);
    load_rows(gm, 2 * MEM, a.S.gm, rows, nb);
    load_rows(gh, 2 * G, a.S.gh, rows, nb);
    load_rows(both, H2 + MEM, a.S.both, rows, nb);
    load_rows(chat, MEM, a.S.chat, rows, nb);
    load_rows(a2, A2, a.S.a2, rows, nb);
    load_rows(att, H2, a.S.att, rows, nb);
    load_rows(cstar, H2, a.S.cstar, rows, nb);
    load_rows(a1, A1, a.S.a1, rows, nb);
    load_rows(gates, 4 * Hs, a.S.gates, rows, nb);
    cta_sync();
    // ---- head: d_oh = dy * w_o2 gated by (oh > 0) -------------------------------------------------
    for (int e = tid; e < O * BT; e += NTHREADS) {
      int j = e / BT, b = e % BT;
      dzoh[e] = oh[e] > 0.f ? dyv[b] * P[D.out_fc2.w + j] * sc_o : 0.f;
    }
    cta_sync();
    stash_rows(a.S.dzoh, O, dzoh, rows, nb);
    dense<STREAM, WT>(ring, W + D.out_fc1.w, O, Hs + MEM, dzoh, part, [&](int n, float* acc) {
      float* dst = n < Hs ? dh + n * BT : dmem + (n - Hs) * BT;
#pragma unroll
      for (int b = 0; b < BT; ++b) dst[b] += acc[b];
    });
    cta_sync();
    // ---- mem_t = gm1*mem_prev + gm2*chat ----------------------------------------------------------
    for (int e = tid; e < MEM * BT; e += NTHREADS) {
      const float g = dmem[e], g1 = gm[e], g2 = gm[MEM * BT + e];
      const float mem_prev = both[H2 * BT + e], ch = chat[e];
      dzg[e] = g * mem_prev * g1 * (1.f - g1);
      dzg[MEM * BT + e] = g * ch * g2 * (1.f - g2);
      dzchat[e] = g * g2 * (1.f - ch * ch);
      dmem[e] = g * g1;                                            // direct path to mem_{t-1}
    }
    cta_sync();
    stash_rows(a.S.dzg, 2 * MEM, dzg, rows, nb);
    stash_rows(a.S.dzchat, MEM, dzchat, rows, nb);
    dense<STREAM, WT>(ring, W + D.g1_fc2.w, MEM, G, dzg, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dgh[n * BT + b] = gh[n * BT + b] > 0.f ? acc[b] * sc_g : 0.f;
    });
    dense<STREAM, WT>(ring, W + D.g2_fc2.w, MEM, G, dzg + MEM * BT, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dgh[(G + n) * BT + b] = gh[(G + n) * BT + b] > 0.f ? acc[b] * sc_g : 0.f;
    });
    dense<STREAM, WT>(ring, W + D.att2_fc2.w, MEM, A2, dzchat, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) da2[n * BT + b] = a2[n * BT + b] > 0.f ? acc[b] : 0.f;
    });
    cta_sync();
    stash_rows(a.S.dgh, 2 * G, dgh, rows, nb);
    stash_rows(a.S.da2, A2, da2, rows, nb);
    // ---- d both = gamma1_fc1^T dgh1 + gamma2_fc1^T dgh2 ; d attended += att2_fc1^T da2 ------------------
    dense<STREAM, WT>(ring, W + D.g1_fc1.w, G, H2 + MEM, dgh, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dboth[n * BT + b] = acc[b];
    });
    cta_sync();
    dense<STREAM, WT>(ring, W + D.g2_fc1.w, G, H2 + MEM, dgh + G * BT, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dboth[n * BT + b] += acc[b];
    });
    cta_sync();
    dense<STREAM, WT>(ring, W + D.att2_fc1.w, A2, H2, da2, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dboth[n * BT + b] += acc[b];
    });
    cta_sync();
    for (int e = tid; e < MEM * BT; e += NTHREADS) dmem[e] += dboth[H2 * BT + e];     // both = attended || mem_{t-1}
    // ---- attended = att * cstar ; att = softmax(logits) -------------------------------------------
    if (warp < BT) {
      float dot = 0.f;
      for (int f = lane; f < H2; f += 32) dot = fmaf(dboth[f * BT + warp] * cstar[f * BT + warp], att[f * BT + warp], dot);
      dot = warp_sum(dot);
      for (int f = lane; f < H2; f += 32) {
        const float da = dboth[f * BT + warp], pa = att[f * BT + warp], cs = cstar[f * BT + warp];
        dlogit[f * BT + warp] = pa * (da * cs - dot);
        dcstar[f * BT + warp] = da * pa;
      }
    }
    cta_sync();
    stash_rows(a.S.dlogit, H2, dlogit, rows, nb);
    dense<STREAM, WT>(ring, W + D.att1_fc2.w, H2, A1, dlogit, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) da1[n * BT + b] = a1[n * BT + b] > 0.f ? acc[b] : 0.f;
    });
    cta_sync();
    stash_rows(a.S.da1, A1, da1, rows, nb);
    dense<STREAM, WT>(ring, W + D.att1_fc1.w, A1, H2, da1, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dcstar[n * BT + b] += acc[b];
    });
    cta_sync();
    // ---- LSTM cell backward -----------------------------------------------------------------------
    for (int e = tid; e < Hs * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const int m = mod_of(D, j);
      const int H = D.H[m], jj = j - D.hoff[m];
      const float* gmod = gates + 4 * D.hoff[m] * BT;
      float* dzm = dz + 4 * D.hoff[m] * BT;
      const float gi = gmod[(0 * H + jj) * BT + b], gf = gmod[(1 * H + jj) * BT + b];
      const float gg = gmod[(2 * H + jj) * BT + b], go = gmod[(3 * H + jj) * BT + b];
      const float c_prev = cstar[j * BT + b], c_new = cstar[(Hs + j) * BT + b];
      const float tc = tanhf(c_new);
      const float dhv = dh[e];
      float dcv = dc[e] + dcstar[(Hs + j) * BT + b] + dhv * go * (1.f - tc * tc);
      dzm[(0 * H + jj) * BT + b] = dcv * gg * gi * (1.f - gi);
      dzm[(1 * H + jj) * BT + b] = dcv * c_prev * gf * (1.f - gf);
      dzm[(2 * H + jj) * BT + b] = dcv * gi * (1.f - gg * gg);
      dzm[(3 * H + jj) * BT + b] = dhv * tc * go * (1.f - go);
      dc[e] = dcstar[j * BT + b] + dcv * gf;                       // gradient wrt c_{t-1}
    }
    cta_sync();
    stash_rows(a.S.dz_lstm, 4 * Hs, dz, rows, nb);
    for (int m = 0; m < D.n_mods; ++m) {
      const int H = D.H[m];
      float* dst = dhp + D.hoff[m] * BT;
      dense<STREAM, WT>(ring, W + D.w_hh[m], 4 * H, H, dz + 4 * D.hoff[m] * BT, part, [&](int n, float* acc) {
#pragma unroll
        for (int b = 0; b < BT; ++b) dst[n * BT + b] = acc[b];
      });
    }
    cta_sync();
    for (int e = tid; e < Hs * BT; e += NTHREADS) dh[e] = dhp[e];
    cta_sync();
  }
}

size_t fwd_smem_floats(const Dims& D) {
  const int Hs = D.Hs, H2 = 2 * D.Hs;
  size_t f = (size_t)Hs * 2 + D.MEM + 4 * Hs + H2 + D.A1 + H2 + (H2 + D.MEM) + D.A2 + D.MEM + 2 * D.G + 2 * D.MEM + (Hs + D.MEM) + D.O;
  return f * BT + (size_t)PART_FLOATS + RING_BYTES / sizeof(float);
}
size_t bwd_smem_floats(const Dims& D) {
  const int Hs = D.Hs, H2 = 2 * D.Hs;
  size_t f = (size_t)Hs * 3 + D.MEM + D.O + 4 * D.MEM + 4 * D.G + 2 * (H2 + D.MEM) + 2 * D.MEM + 2 * D.A2 + 4 * H2 + 2 * D.A1 + 8 * Hs + D.O;
  return f * BT + (size_t)PART_FLOATS + RING_BYTES / sizeof(float);
}

int layout_strides(const MtMfnCfg& c, const Dims& D, const int64_t* stride_b, const int64_t* stride_t, long long& sb, long long& st) {
  // all modalities must share one dense row order: [B,T,D] (sb=T, st=1) or the reference's [T,B,D] (sb=1, st=B)
  bool bt = true, tb = true;
  for (int m = 0; m < D.n_mods; ++m) {
    const int64_t Dm = D.D[m];
    bt = bt && (c.T == 1 || stride_t[m] == Dm) && (c.B == 1 || stride_b[m] == Dm * c.T);
    tb = tb && (c.B == 1 || stride_b[m] == Dm) && (c.T == 1 || stride_t[m] == Dm * c.B);
  }
  if (bt) { sb = c.T; st = 1; return MT_OK; }
  if (tb) { sb = 1; st = c.B; return MT_OK; }
  return MT_ERR_ARG;
}

}  // namespace

extern "C" {

size_t mt_mfn_param_count(const MtMfnCfg* cfg) {
  Dims D;
  if (!cfg || make_dims(*cfg, D) != MT_OK) return 0;
  return D.total;
}

size_t mt_mfn_ws_bytes(const MtMfnCfg* cfg) {
  Dims D;
  if (!cfg || make_dims(*cfg, D) != MT_OK) return 0;
  Stash s;
  carve(*cfg, D, nullptr, s);
  return s.bytes;
}

int mt_mfn_fwd(const MtMfnCfg* cfg, const float* params, const void* params_lp, const void* const* x, const int64_t* stride_b,
               const int64_t* stride_t, const float* mask, float* out, float* h_last, float* c_last, float* mem_last, void* ws,
               size_t ws_bytes, void* stream) {
  if (!cfg || !params || !x || !stride_b || !stride_t || !out || !ws) return MT_ERR_ARG;
  const MtMfnCfg& c = *cfg;
  if (c.dtype != MT_F32 && c.dtype != MT_BF16) return MT_ERR_ARG;
  Dims D;
  MT_TRY(make_dims(c, D));
  Stash S;
  carve(c, D, ws, S);
  if (ws_bytes < S.bytes) return MT_ERR_WS;
  long long sb, stt;
  MT_TRY(layout_strides(c, D, stride_b, stride_t, sb, stt));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = c.B * c.T;
  const bool lp = c.dtype == MT_BF16;
  (void)params_lp;

  // transposed forward pack
  TransposeJob jobs[16];
  int nj = 0;
  auto tp = [&](size_t src, size_t dst, int R, int C, int ldd, int col_off) {
    jobs[nj].src = params + src;
    jobs[nj].dst = lp ? (void*)((bf16*)S.tpack + dst + col_off) : (void*)((float*)S.tpack + dst + col_off);
    jobs[nj].R = R; jobs[nj].C = C; jobs[nj].ldd = ldd; ++nj;
  };
  const int H2 = 2 * D.Hs;
  for (int m = 0; m < D.n_mods; ++m) tp(D.w_hh[m], D.t_hh[m], 4 * D.H[m], D.H[m], 4 * D.H[m], 0);
  tp(D.att1_fc1.w, D.t_att1_fc1, D.A1, H2, D.A1, 0);
  tp(D.att1_fc2.w, D.t_att1_fc2, H2, D.A1, H2, 0);
  tp(D.att2_fc1.w, D.t_att2_fc1, D.A2, H2, D.A2, 0);
  tp(D.att2_fc2.w, D.t_att2_fc2, D.MEM, D.A2, D.MEM, 0);
  tp(D.g1_fc1.w, D.t_g_fc1, D.G, H2 + D.MEM, 2 * D.G, 0);
  tp(D.g2_fc1.w, D.t_g_fc1, D.G, H2 + D.MEM, 2 * D.G, D.G);
  tp(D.g1_fc2.w, D.t_g1_fc2, D.MEM, D.G, D.MEM, 0);
  tp(D.g2_fc2.w, D.t_g2_fc2, D.MEM, D.G, D.MEM, 0);
  tp(D.out_fc1.w, D.t_out_fc1, D.O, D.Hs + D.MEM, D.O, 0);
  MT_TRY(mt_transpose_pack_run(jobs, nj, lp, st));

  // hoisted input projections: gates[:, 4*hoff_m : +4H_m] = x_m W_ih^T + b_ih  (+ b_hh added below)
  for (int m = 0; m < D.n_mods; ++m) {
    const void* xm = x[m];
    if (lp) {
      MT_TRY(mt_cast2d_run(x[m], true, D.D[m], S.xf[m], false, D.D[m], M, D.D[m], mt_make_drop(0.f, 0, 0), st));
      xm = S.xf[m];
    }
    GemmDesc g;
    g.M = M; g.N = 4 * D.H[m]; g.K = D.D[m];
    g.A = xm; g.lda = D.D[m]; g.a_kmajor = true;
    g.B = params + D.w_ih[m]; g.ldb = D.D[m]; g.b_kmajor = true;
    g.C = S.gates + 4 * D.hoff[m]; g.ldc = 4 * D.Hs; g.c_f32 = true;
    g.epi.bias = params + D.b_ih[m];
    MT_TRY(mt_gemm_run(MT_F32, g, st));      // b_hh is added inside the recurrence together with W_hh h
  }
  KArgs a;
  a.D = D; a.S = S; a.params = params; a.wlp = S.tpack; a.mask = mask; a.out = out; a.dout = nullptr;
  a.h_last = h_last; a.c_last = c_last; a.mem_last = mem_last;
  a.B = c.B; a.T = c.T; a.sb = sb; a.st = stt; a.training = c.training;
  const float pg = c.p_gamma, po = c.p_out;
  a.drop_g1 = mt_make_drop(pg, c.seed, MT_SITE_MFN_G1);
  a.drop_g2 = mt_make_drop(pg, c.seed, MT_SITE_MFN_G2);
  a.drop_out = mt_make_drop(po, c.seed, MT_SITE_MFN_OUT);
  const size_t smem = fwd_smem_floats(D) * sizeof(float);
  mt_prof_work(2.0 * (double)D.t_total * c.B * c.T, 0.0);
  const int grid = (c.B + BT - 1) / BT;
  // weight blocks in the order the kernel consumes them each step
  a.tab.n = 0;
  auto add = [&](size_t off, int K, int N) {
    StreamLayer& Lr = a.tab.L[a.tab.n++];
    Lr.ptr = lp ? (const void*)((const bf16*)S.tpack + off) : (const void*)((const float*)S.tpack + off);
    Lr.ptr_special = nullptr; Lr.K = K; Lr.N = N;
  };
  for (int m = 0; m < D.n_mods; ++m) add(D.t_hh[m], D.H[m], 4 * D.H[m]);
  add(D.t_att1_fc1, H2, D.A1); add(D.t_att1_fc2, D.A1, H2); add(D.t_att2_fc1, H2, D.A2); add(D.t_g_fc1, H2 + D.MEM, 2 * D.G);
  add(D.t_att2_fc2, D.A2, D.MEM); add(D.t_g1_fc2, D.G, D.MEM); add(D.t_g2_fc2, D.G, D.MEM); add(D.t_out_fc1, D.Hs + D.MEM, D.O);
#define MT_MFN_LAUNCH(KERNEL, WT_)                                                        \
  do {                                                                                     \
    if (stream_table_ok<WT_>(a.tab)) {                                                     \
      MT_TRY(set_smem(KERNEL<true, WT_>, smem));                                           \
      KERNEL<true, WT_><<<grid, NTHREADS + 32, smem, st>>>(a);                             \
    } else {                                                                               \
      MT_TRY(set_smem(KERNEL<false, WT_>, smem));                                          \
      KERNEL<false, WT_><<<grid, NTHREADS, smem, st>>>(a);                                 \
    }                                                                                      \
  } while (0)
  if (lp) MT_MFN_LAUNCH(mfn_fwd_kernel, bf16); else MT_MFN_LAUNCH(mfn_fwd_kernel, float);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_mfn_bwd(const MtMfnCfg* cfg, const float* params, const void* params_lp, const void* const* x, const int64_t* stride_b,
               const int64_t* stride_t, const float* mask, const float* dout, void* const* dx, float* grads, void* ws,
               size_t ws_bytes, void* stream) {
  if (!cfg || !params || !x || !stride_b || !stride_t || !dout || !grads || !ws) return MT_ERR_ARG;
  const MtMfnCfg& c = *cfg;
  if (!c.training) return MT_ERR_ARG;
  if (c.dtype != MT_F32 && c.dtype != MT_BF16) return MT_ERR_ARG;
  if (c.dtype == MT_BF16 && !params_lp) return MT_ERR_ARG;
  Dims D;
  MT_TRY(make_dims(c, D));
  Stash S;
  carve(c, D, ws, S);
  if (ws_bytes < S.bytes) return MT_ERR_WS;
  long long sb, stt;
  MT_TRY(layout_strides(c, D, stride_b, stride_t, sb, stt));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = c.B * c.T;
  const bool lp = c.dtype == MT_BF16;
  const int Hs = D.Hs, H2 = 2 * D.Hs, MEM = D.MEM;

  KArgs a;
  a.D = D; a.S = S; a.params = params; a.wlp = lp ? params_lp : (const void*)params; a.mask = mask; a.out = nullptr; a.dout = dout;
  a.h_last = a.c_last = a.mem_last = nullptr;
  a.B = c.B; a.T = c.T; a.sb = sb; a.st = stt; a.training = 1;
  a.drop_g1 = mt_make_drop(c.p_gamma, c.seed, MT_SITE_MFN_G1);
  a.drop_g2 = mt_make_drop(c.p_gamma, c.seed, MT_SITE_MFN_G2);
  a.drop_out = mt_make_drop(c.p_out, c.seed, MT_SITE_MFN_OUT);
  const size_t smem = bwd_smem_floats(D) * sizeof(float);
  mt_prof_work(2.0 * (double)D.t_total * c.B * c.T, 0.0);
  const int grid = (c.B + BT - 1) / BT;
  a.tab.n = 0;
  auto add = [&](size_t off, int K, int N) {
    StreamLayer& Lr = a.tab.L[a.tab.n++];
    Lr.ptr = lp ? (const void*)((const bf16*)params_lp + off) : (const void*)(params + off);
    Lr.ptr_special = nullptr; Lr.K = K; Lr.N = N;
  };
  // row-major originals W[out][in] consumed as K = out, N = in, in the order of the reverse-time kernel
  add(D.out_fc1.w, D.O, Hs + MEM); add(D.g1_fc2.w, MEM, D.G); add(D.g2_fc2.w, MEM, D.G); add(D.att2_fc2.w, MEM, D.A2);
  add(D.g1_fc1.w, D.G, H2 + MEM); add(D.g2_fc1.w, D.G, H2 + MEM); add(D.att2_fc1.w, D.A2, H2); add(D.att1_fc2.w, H2, D.A1);
  add(D.att1_fc1.w, D.A1, H2);
  for (int m = 0; m < D.n_mods; ++m) add(D.w_hh[m], 4 * D.H[m], D.H[m]);
  if (lp) MT_MFN_LAUNCH(mfn_bwd_kernel, bf16); else MT_MFN_LAUNCH(mfn_bwd_kernel, float);
  MT_LAUNCH_CHECK();

  // ---- batched weight gradients over all T*B rows ----------------------------------------------------
  MT_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * D.total, st));
  auto wg = [&](const float* dz, int ldz, int Nout, const float* xin, int ldx, int Kin, size_t w_off, int ldw) -> int {
    return mt_gemm_run(MT_F32, mt_wgrad_desc(M, Nout, Kin, dz, ldz, xin, ldx, grads + w_off, ldw), st);
  };
  auto bg = [&](const float* dz, int ldz, int Nout, size_t b_off) -> int {
    return mt_colsum_run(0, M, Nout, dz, ldz, grads + b_off, 1, st);
  };
  for (int m = 0; m < D.n_mods; ++m) {
    const float* dzm = S.dz_lstm + 4 * D.hoff[m];
    const float* xm = lp ? S.xf[m] : (const float*)x[m];
    MT_TRY(wg(dzm, 4 * Hs, 4 * D.H[m], xm, D.D[m], D.D[m], D.w_ih[m], D.D[m]));
    MT_TRY(wg(dzm, 4 * Hs, 4 * D.H[m], S.hprev + D.hoff[m], Hs, D.H[m], D.w_hh[m], D.H[m]));
    MT_TRY(bg(dzm, 4 * Hs, 4 * D.H[m], D.b_ih[m]));
    MT_TRY(bg(dzm, 4 * Hs, 4 * D.H[m], D.b_hh[m]));
    if (dx && dx[m]) {
      GemmDesc g;
      g.M = M; g.N = D.D[m]; g.K = 4 * D.H[m];
      g.A = dzm; g.lda = 4 * Hs; g.a_kmajor = true;
      g.B = params + D.w_ih[m]; g.ldb = D.D[m]; g.b_kmajor = false;
      if (lp) {
        // fp32 result into the (no longer needed) fp32 input copy, then one cast to the bf16 gradient
        g.C = S.xf[m]; g.ldc = D.D[m]; g.c_f32 = true;
        MT_TRY(mt_gemm_run(MT_F32, g, st));
        MT_TRY(mt_cast2d_run(S.xf[m], false, D.D[m], dx[m], true, D.D[m], M, D.D[m], mt_make_drop(0.f, 0, 0), st));
      } else {
        g.C = dx[m]; g.ldc = D.D[m]; g.c_f32 = true;
        MT_TRY(mt_gemm_run(MT_F32, g, st));
      }
    }
  }
  MT_TRY(wg(S.da1, D.A1, D.A1, S.cstar, H2, H2, D.att1_fc1.w, H2));          MT_TRY(bg(S.da1, D.A1, D.A1, D.att1_fc1.b));
  MT_TRY(wg(S.dlogit, H2, H2, S.a1, D.A1, D.A1, D.att1_fc2.w, D.A1));        MT_TRY(bg(S.dlogit, H2, H2, D.att1_fc2.b));
  MT_TRY(wg(S.da2, D.A2, D.A2, S.both, H2 + MEM, H2, D.att2_fc1.w, H2));     MT_TRY(bg(S.da2, D.A2, D.A2, D.att2_fc1.b));
  MT_TRY(wg(S.dzchat, MEM, MEM, S.a2, D.A2, D.A2, D.att2_fc2.w, D.A2));      MT_TRY(bg(S.dzchat, MEM, MEM, D.att2_fc2.b));
  MT_TRY(wg(S.dgh, 2 * D.G, D.G, S.both, H2 + MEM, H2 + MEM, D.g1_fc1.w, H2 + MEM));          MT_TRY(bg(S.dgh, 2 * D.G, D.G, D.g1_fc1.b));
  MT_TRY(wg(S.dgh + D.G, 2 * D.G, D.G, S.both, H2 + MEM, H2 + MEM, D.g2_fc1.w, H2 + MEM));    MT_TRY(bg(S.dgh + D.G, 2 * D.G, D.G, D.g2_fc1.b));
  MT_TRY(wg(S.dzg, 2 * MEM, MEM, S.gh, 2 * D.G, D.G, D.g1_fc2.w, D.G));                       MT_TRY(bg(S.dzg, 2 * MEM, MEM, D.g1_fc2.b));
  MT_TRY(wg(S.dzg + MEM, 2 * MEM, MEM, S.gh + D.G, 2 * D.G, D.G, D.g2_fc2.w, D.G));           MT_TRY(bg(S.dzg + MEM, 2 * MEM, MEM, D.g2_fc2.b));
  MT_TRY(wg(S.dzoh, D.O, D.O, S.last, Hs + MEM, Hs + MEM, D.out_fc1.w, Hs + MEM));            MT_TRY(bg(S.dzoh, D.O, D.O, D.out_fc1.b));
  MT_TRY(wg(S.dyv, 1, 1, S.oh, D.O, D.O, D.out_fc2.w, D.O));                                  MT_TRY(bg(S.dyv, 1, 1, D.out_fc2.b));
  return MT_OK;
}

}  // extern "C"
