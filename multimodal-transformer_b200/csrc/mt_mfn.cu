// Memory Fusion Network (MFN.forward MFT/multiTransformer.py:181-248), forward and backward.
//
// The reference walks one python loop over t that does everything.  Its data dependences are much thinner than that
// loop: the LSTHM cells (:208) only see x_t and their own (h, c); the delta-memory attention (:210-220) is a pure
// function of cStar_t = [c_{t-1} || c_t]; only the gamma gates + memory update (:221-224) carry `mem`.  So the work
// is re-cut B200-first into
//
//   F1  zx      = x_m W_ih^T + b_ih                     one GEMM per modality over all T*B rows       (tcgen05)
//   F2  LSTM recurrence, one persistent CTA per (narrative tile, modality): W_hh of that modality lives in shared
//       memory for the whole sequence, h / c never leave the SM, zx rows are prefetched with cp.async 3 steps ahead
//   F3  a1 = relu(att1_fc1 cStar), logits = att1_fc2 a1, att = softmax_features(logits), attended = att * cStar,
//       a2 = relu(att2_fc1 attended), cHat = tanh(att2_fc2 a2), gpre = gamma{1,2}_fc1[:, :2H] attended + b
//                                                        batched GEMMs with fused epilogues over all rows (tcgen05)
//   F4  memory recurrence, one persistent CTA per narrative tile: gh = drop(relu(gpre_t + gamma_fc1[:, 2H:] mem)),
//       gamma = sigmoid(gamma_fc2 gh), mem = gamma1 mem + gamma2 cHat_t; 32 k weights resident in shared memory
//   F5  head: relu(out_fc1 [h || mem]) as one GEMM, then dropout / out_fc2 / mask row-wise
//
// and backward mirrors it: head -> reverse-time memory recurrence (carries dmem) -> batched dgrads through the
// attention block -> reverse-time LSTM recurrence (carries dh, dc) -> batched wgrad GEMMs over all T*B rows.
// The sequential part shrinks from 847 kflop/token to 207 kflop/token and never streams a weight from L2.
//
// Inside a recurrence CTA activations are feature-major [feature][BT] in shared memory so one 16-byte load feeds BT
// FMAs.  Stash row order follows the inputs: row(b,t) = b*sb + t*st with (sb,st) = (T,1) for [B,T,D] inputs and
// (1,B) for the reference's permuted [T,B,D] views.  ST is the GEMM operand type (float / bf16): tensors that only
// feed GEMMs are kept in ST, recurrent state and everything entering a transcendental stays fp32.
#include <algorithm>
#include "mt_recurrent.cuh"
#include "mt_mfn.cuh"

GemmDesc mt_wgrad_desc(int M, int Nout, int Kin, const void* dy, int ldy, const void* x, int ldx, float* dW, int ldw);

namespace {

using namespace mtrec;

constexpr int NSTAGE = 4;       // cp.async row ring: step t + NSTAGE - 1 is in flight while step t is consumed
constexpr int MAX_SEG = 6;

struct LinOff { size_t w, b; };

struct Dims {
  int n_mods;
  int D[MT_MAX_MODS], H[MT_MAX_MODS], hoff[MT_MAX_MODS];
  int Hs, MEM, A1, A2, G, O;
  // fp32 flat parameter offsets
  size_t w_ih[MT_MAX_MODS], w_hh[MT_MAX_MODS], b_ih[MT_MAX_MODS], b_hh[MT_MAX_MODS];
  LinOff att1_fc1, att1_fc2, att2_fc1, att2_fc2, g1_fc1, g1_fc2, g2_fc1, g2_fc2, out_fc1, out_fc2;
  size_t total;
};

int make_dims(const MtMfnCfg& c, Dims& D) {
  if (c.n_mods < 1 || c.n_mods > MT_MAX_MODS || c.B <= 0 || c.T <= 0) return MT_ERR_ARG;
  D.n_mods = c.n_mods;
  D.Hs = 0;
  const int q = c.dtype == MT_BF16 ? 8 : 4;      // 16-byte row segments for cp.async and vector loads
  for (int m = 0; m < c.n_mods; ++m) {
    if (c.in_dim[m] <= 0 || c.hid[m] <= 0 || c.hid[m] % q != 0 || c.in_dim[m] % 4 != 0) return MT_ERR_ARG;
    D.D[m] = c.in_dim[m]; D.H[m] = c.hid[m]; D.hoff[m] = D.Hs; D.Hs += c.hid[m];
  }
  D.MEM = c.mem_dim; D.A1 = c.h_att1; D.A2 = c.h_att2; D.G = c.h_gamma; D.O = c.h_out;
  if (D.MEM <= 0 || D.A1 <= 0 || D.A2 <= 0 || D.G <= 0 || D.O <= 0) return MT_ERR_ARG;
  if ((D.A1 | D.A2 | D.O) % 4 != 0 || D.MEM % q != 0 || D.G % 4 != 0) return MT_ERR_ARG;
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
  return MT_OK;
}

// global stash, one row per (b,t).  "op" = GEMM operand dtype (ST).
struct Stash {
  float* gates;     // [M,4Hs] fp32  zx before the LSTM recurrence; post-activation i,f,g,o after it (per modality i|f|g|o)
  float* cstar;     // [M,2Hs] fp32  c_{t-1} || c_t
  void* cstar_op;   // [M,2Hs] op    (aliases cstar in fp32 mode)
  void* last_op;    // [M,Hs+MEM] op h_t || mem_t
  void* a1_op;      // [M,A1]  op
  float* att;       // [M,2Hs] fp32  logits, then the softmax output
  void* attd_op;    // [M,2Hs] op    attended = att * cStar
  void* a2_op;      // [M,A2]  op
  float* chat;      // [M,MEM] fp32
  float* gpre;      // [M,2G]  fp32  gamma{1,2}_fc1 applied to `attended` (+ bias); the mem part is added in the recurrence
  float* pre;       // [M,O]   fp32  relu(out_fc1 last + b)
  // training only
  void* hprev_op;   // [M,Hs]  op
  void* gh_op;      // [M,2G]  op    gamma1 | gamma2 hidden (post relu, post dropout)
  float* gm;        // [M,2MEM] fp32 gamma1 | gamma2
  void* memprev_op; // [M,MEM] op
  void* oh_op;      // [M,O]   op    out hidden (post relu, post dropout)
  // backward
  float* dlast;     // [M,Hs+MEM] fp32
  void* dzoh_op;    // [M,O]   op
  void* dzg_op;     // [M,2MEM] op
  void* dzchat_op;  // [M,MEM] op
  void* dgh_op;     // [M,2G]  op
  void* da2_op;     // [M,A2]  op
  float* datt;      // [M,2Hs] fp32  d attended, then d cStar (in place)
  void* dlogit_op;  // [M,2Hs] op
  void* da1_op;     // [M,A1]  op
  void* dz_op;      // [M,4Hs] op    LSTM pre-activation gradients
  size_t bytes;
};

void carve(const MtMfnCfg& c, const Dims& D, void* ws, Stash& s) {
  const size_t M = (size_t)c.B * c.T, es = mt_esize(c.dtype);
  const int H2 = 2 * D.Hs;
  const bool lp = c.dtype == MT_BF16;
  WsCarver k(ws);
  s.gates = k.take<float>(M * 4 * D.Hs);
  s.cstar = k.take<float>(M * H2);
  s.cstar_op = lp ? k.take_bytes(M * H2 * es) : (void*)s.cstar;
  s.last_op = k.take_bytes(M * (D.Hs + D.MEM) * es);
  s.a1_op = k.take_bytes(M * D.A1 * es);
  s.att = k.take<float>(M * H2);
  s.attd_op = k.take_bytes(M * H2 * es);
  s.a2_op = k.take_bytes(M * D.A2 * es);
  s.chat = k.take<float>(M * D.MEM);
  s.gpre = k.take<float>(M * 2 * D.G);
  s.pre = k.take<float>(M * D.O);
  if (c.training) {
    s.hprev_op = k.take_bytes(M * D.Hs * es);
    s.gh_op = k.take_bytes(M * 2 * D.G * es);
    s.gm = k.take<float>(M * 2 * D.MEM);
    s.memprev_op = k.take_bytes(M * D.MEM * es);
    s.oh_op = k.take_bytes(M * D.O * es);
    s.dlast = k.take<float>(M * (D.Hs + D.MEM));
    s.dzoh_op = k.take_bytes(M * D.O * es);
    s.dzg_op = k.take_bytes(M * 2 * D.MEM * es);
    s.dzchat_op = k.take_bytes(M * D.MEM * es);
    s.dgh_op = k.take_bytes(M * 2 * D.G * es);
    s.da2_op = k.take_bytes(M * D.A2 * es);
    s.datt = k.take<float>(M * H2);
    s.dlogit_op = k.take_bytes(M * H2 * es);
    s.da1_op = k.take_bytes(M * D.A1 * es);
    s.dz_op = k.take_bytes(M * 4 * D.Hs * es);
  } else {
    s.hprev_op = s.gh_op = s.memprev_op = s.oh_op = nullptr;
    s.gm = nullptr;
    s.dlast = s.datt = nullptr;
    s.dzoh_op = s.dzg_op = s.dzchat_op = s.dgh_op = s.da2_op = s.dlogit_op = s.da1_op = s.dz_op = nullptr;
  }
  s.bytes = k.total();
}

// ======================================================================================================
// device helpers
// ======================================================================================================
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one per-step row segment of a [M, ld] global tensor: `bytes` bytes starting at column byte offset folded into base
struct Seg { const char* base; long long row_bytes; int bytes; int soff; };
struct SegTab {
  Seg s[MAX_SEG];
  int n, stage_bytes, chunks;
  __device__ void add(const void* base, long long row_bytes, int bytes) {
    s[n].base = reinterpret_cast<const char*>(base); s[n].row_bytes = row_bytes; s[n].bytes = bytes; s[n].soff = stage_bytes;
    stage_bytes += BT * bytes; chunks += BT * (bytes >> 4); ++n;
  }
};

// issue the 16-byte async copies of step t for every narrative of the tile; staged layout: segment -> [b][bytes]
__device__ __forceinline__ void prefetch_rows(const SegTab& tab, char* stage, int b0, int nb, long long sb, long long st, int t) {
  for (int c = threadIdx.x; c < tab.chunks; c += NTHREADS) {
    int s = 0, cc = c;
    while (s < tab.n - 1 && cc >= BT * (tab.s[s].bytes >> 4)) { cc -= BT * (tab.s[s].bytes >> 4); ++s; }
    const Seg& sg = tab.s[s];
    const int per = sg.bytes >> 4;
    const int b = cc / per, ch = cc - b * per;
    const long long row = (long long)(b0 + min(b, nb - 1)) * sb + (long long)t * st;
    cp16(stage + sg.soff + b * sg.bytes + ch * 16, sg.base + row * sg.row_bytes + ch * 16);
  }
}

// in-CTA dense layer on shared-memory-resident weights:  out[n][b] = sum_k Wt[k*N + n] * xs(g)[k*BT + b]
// Thread (p, g) owns the VN outputs n = g*VN.. of K-slice p; xs_of(g) picks the activation block of that output group
// (lets two independent layers share one pass).  epi(n, acc[BT]) runs exactly once per n; ends with a CTA barrier.
template <typename WT, typename XsOf, typename Epi>
__device__ __forceinline__ void dense_s(const WT* __restrict__ Wt, int K, int N, XsOf xs_of, float* part, Epi epi) {
  const int tid = threadIdx.x;
  const int NG = N / VN;
  int P = NTHREADS / NG;
  if (P > 16) P = 16;
  if (P > K) P = K;
  const int p = tid / NG, g = tid - p * NG;
  float acc[VN][BT];
#pragma unroll
  for (int i = 0; i < VN; ++i)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[i][b] = 0.f;
  if (p < P) {
    const WT* w = Wt + g * VN;
    const float* xs = xs_of(g);
#pragma unroll 4
    for (int k = p; k < K; k += P) {
      float wv[VN];
      load_w4(w + (size_t)k * N, wv);
      const float4 x = *reinterpret_cast<const float4*>(xs + k * BT);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        acc[i][0] = fmaf(wv[i], x.x, acc[i][0]); acc[i][1] = fmaf(wv[i], x.y, acc[i][1]);
        acc[i][2] = fmaf(wv[i], x.z, acc[i][2]); acc[i][3] = fmaf(wv[i], x.w, acc[i][3]);
      }
    }
  }
  if (P == 1) cta_sync();                  // the epilogue may overwrite buffers other threads are still reading
  dense_finish(acc, P, p, g, N, part, epi);
}

__host__ __device__ inline size_t part_floats(int K, int N) {
  const int NG = N / VN;
  int P = NTHREADS / NG;
  if (P > 16) P = 16;
  if (P > K) P = K;
  return P > 1 ? (size_t)P * N * BT : 0;
}

__host__ __device__ inline size_t smax(size_t a, size_t b) { return a > b ? a : b; }

template <typename ST>
__device__ __forceinline__ void st_op(ST* p, float v) { *p = from_f<ST>(v); }

// ======================================================================================================
// F2 / B4: LSTM recurrences (one CTA per narrative tile and modality)
// ======================================================================================================
template <typename WT, typename ST>
__global__ void __launch_bounds__(NTHREADS, 1) mfn_lstm_fwd_kernel(const __grid_constant__ LstmArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], G4 = 4 * H, Hs = a.Hs;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  // ---- shared memory carve ----
  char* sp = reinterpret_cast<char*>(smem_raw);
  char* stages = sp; sp += (size_t)NSTAGE * BT * G4 * sizeof(float);
  float* h = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* c = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* hp = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* cp = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* z = reinterpret_cast<float*>(sp); sp += G4 * BT * sizeof(float);
  float* part = reinterpret_cast<float*>(sp); sp += part_floats(H, G4) * sizeof(float);
  WT* Wt = reinterpret_cast<WT*>(sp);                       // [H][4H]: Wt[k*4H + n] = W_hh[n][k]
  __shared__ long long rows[2][BT];
  __shared__ SegTab tab;
  if (tid == 0) {
    tab.n = 0; tab.stage_bytes = 0; tab.chunks = 0;
    tab.add(a.gates + 4 * hoff, (long long)4 * Hs * sizeof(float), G4 * (int)sizeof(float));
  }
  const WT* W = reinterpret_cast<const WT*>(a.w_hh[m]);
  for (int e = tid; e < G4 * H; e += NTHREADS) { const int n = e / H, k = e - n * H; Wt[k * G4 + n] = W[e]; }
  for (int e = tid; e < H * BT; e += NTHREADS) { h[e] = 0.f; c[e] = 0.f; }
  __syncthreads();
  const int stage_bytes = tab.stage_bytes;
  for (int i = 0; i < NSTAGE - 1; ++i) {
    if (i < a.T) prefetch_rows(tab, stages + (size_t)i * stage_bytes, b0, nb, a.sb, a.st, i);
    cp_commit();
  }
  const float* bhh = a.b_hh[m];
  ST* cstar_op = reinterpret_cast<ST*>(a.cstar_op);
  ST* last_op = reinterpret_cast<ST*>(a.last_op);
  ST* hprev_op = reinterpret_cast<ST*>(a.hprev_op);
  const int LW = Hs + a.MEM;

  for (int t = 0; t < a.T; ++t) {
    long long* rw = rows[t & 1];
    if (tid < BT) rw[tid] = (long long)(b0 + min(tid, nb - 1)) * a.sb + (long long)t * a.st;
    cp_wait<NSTAGE - 2>();
    cta_sync();
    if (t + NSTAGE - 1 < a.T) prefetch_rows(tab, stages + (size_t)((t + NSTAGE - 1) % NSTAGE) * stage_bytes, b0, nb, a.sb, a.st, t + NSTAGE - 1);
    cp_commit();
    const float* zs = reinterpret_cast<const float*>(stages + (size_t)(t % NSTAGE) * stage_bytes);      // [b][4H]
    // ---- z = zx (hoisted x-projection + b_ih) + b_hh + W_hh h_{t-1} ----
    dense_s<WT>(Wt, H, G4, [&](int) { return h; }, part, [&](int n, float* acc) {
      const float bias = bhh[n];
#pragma unroll
      for (int b = 0; b < BT; ++b) z[n * BT + b] = acc[b] + bias + zs[b * G4 + n];
    });
    // ---- gates: c_t = s(f) c + s(i) tanh(g); h_t = s(o) tanh(c_t) ----
    for (int e = tid; e < H * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const float gi = sigmoidf_(z[(0 * H + j) * BT + b]);
      const float gf = sigmoidf_(z[(1 * H + j) * BT + b]);
      const float gg = tanhf(z[(2 * H + j) * BT + b]);
      const float go = sigmoidf_(z[(3 * H + j) * BT + b]);
      const float co = c[e];
      const float cn = gf * co + gi * gg;
      const float hn = go * tanhf(cn);
      z[(0 * H + j) * BT + b] = gi; z[(1 * H + j) * BT + b] = gf;
      z[(2 * H + j) * BT + b] = gg; z[(3 * H + j) * BT + b] = go;
      hp[e] = h[e]; cp[e] = co;
      c[e] = cn; h[e] = hn;
    }
    cta_sync();
    // ---- coalesced stores (the next step's first barrier orders them against the next overwrite) ----
    if (a.training) {
      for (int e = tid; e < nb * G4; e += NTHREADS) {
        const int b = e / G4, f = e - b * G4;
        a.gates[rw[b] * (4 * Hs) + 4 * hoff + f] = z[f * BT + b];
      }
      for (int e = tid; e < nb * H; e += NTHREADS) {
        const int b = e / H, j = e - b * H;
        st_op(hprev_op + rw[b] * Hs + hoff + j, hp[j * BT + b]);
      }
    }
    for (int e = tid; e < nb * H; e += NTHREADS) {
      const int b = e / H, j = e - b * H;
      const float cpv = cp[j * BT + b], cnv = c[j * BT + b];
      a.cstar[rw[b] * (2 * Hs) + hoff + j] = cpv;
      a.cstar[rw[b] * (2 * Hs) + Hs + hoff + j] = cnv;
      if (cstar_op) { st_op(cstar_op + rw[b] * (2 * Hs) + hoff + j, cpv); st_op(cstar_op + rw[b] * (2 * Hs) + Hs + hoff + j, cnv); }
      st_op(last_op + rw[b] * LW + hoff + j, h[j * BT + b]);
    }
  }
  cp_wait<0>();
  for (int e = tid; e < nb * H; e += NTHREADS) {
    const int b = e / H, j = e % H;
    if (a.h_last) a.h_last[(size_t)(b0 + b) * Hs + hoff + j] = h[j * BT + b];
    if (a.c_last) a.c_last[(size_t)(b0 + b) * Hs + hoff + j] = c[j * BT + b];
  }
}

template <typename WT, typename ST>
__global__ void __launch_bounds__(NTHREADS, 1) mfn_lstm_bwd_kernel(const __grid_constant__ LstmArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int m = blockIdx.y;
  const int H = a.H[m], hoff = a.hoff[m], G4 = 4 * H, Hs = a.Hs;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  char* sp = reinterpret_cast<char*>(smem_raw);
  char* stages = sp; sp += (size_t)NSTAGE * BT * (G4 + 5 * H) * sizeof(float);
  float* dh = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* dc = reinterpret_cast<float*>(sp); sp += H * BT * sizeof(float);
  float* dz = reinterpret_cast<float*>(sp); sp += G4 * BT * sizeof(float);
  float* part = reinterpret_cast<float*>(sp); sp += part_floats(G4, H) * sizeof(float);
  WT* Ws = reinterpret_cast<WT*>(sp);                       // [4H][H] row-major original: K = gate row, N = hidden unit
  __shared__ long long rows[2][BT];
  __shared__ SegTab tab;
  if (tid == 0) {
    tab.n = 0; tab.stage_bytes = 0; tab.chunks = 0;
    const int hb = H * (int)sizeof(float);
    tab.add(a.gates + 4 * hoff, (long long)4 * Hs * sizeof(float), 4 * hb);            // 0: gate activations
    tab.add(a.cstar + hoff, (long long)2 * Hs * sizeof(float), hb);                    // 1: c_{t-1}
    tab.add(a.cstar + Hs + hoff, (long long)2 * Hs * sizeof(float), hb);               // 2: c_t
    tab.add(a.dcstar + hoff, (long long)2 * Hs * sizeof(float), hb);                   // 3: d cStar (prev half)
    tab.add(a.dcstar + Hs + hoff, (long long)2 * Hs * sizeof(float), hb);              // 4: d cStar (new half)
    tab.add(a.dlast + hoff, (long long)(Hs + a.MEM) * sizeof(float), hb);              // 5: d h_t from the head
  }
  const WT* W = reinterpret_cast<const WT*>(a.w_hh[m]);
  for (int e = tid; e < G4 * H; e += NTHREADS) Ws[e] = W[e];
  for (int e = tid; e < H * BT; e += NTHREADS) { dh[e] = 0.f; dc[e] = 0.f; }
  __syncthreads();
  const int stage_bytes = tab.stage_bytes;
  for (int i = 0; i < NSTAGE - 1; ++i) {
    if (i < a.T) prefetch_rows(tab, stages + (size_t)i * stage_bytes, b0, nb, a.sb, a.st, a.T - 1 - i);
    cp_commit();
  }
  ST* dz_op = reinterpret_cast<ST*>(a.dz_op);

  for (int i = 0; i < a.T; ++i) {
    const int t = a.T - 1 - i;
    long long* rw = rows[i & 1];
    if (tid < BT) rw[tid] = (long long)(b0 + min(tid, nb - 1)) * a.sb + (long long)t * a.st;
    cp_wait<NSTAGE - 2>();
    cta_sync();
    if (i + NSTAGE - 1 < a.T)
      prefetch_rows(tab, stages + (size_t)((i + NSTAGE - 1) % NSTAGE) * stage_bytes, b0, nb, a.sb, a.st, t - (NSTAGE - 1));
    cp_commit();
    const char* sg = stages + (size_t)(i % NSTAGE) * stage_bytes;
    const float* gat = reinterpret_cast<const float*>(sg + tab.s[0].soff);     // [b][4H]
    const float* cpv = reinterpret_cast<const float*>(sg + tab.s[1].soff);     // [b][H]
    const float* cnv = reinterpret_cast<const float*>(sg + tab.s[2].soff);
    const float* dcp = reinterpret_cast<const float*>(sg + tab.s[3].soff);
    const float* dcn = reinterpret_cast<const float*>(sg + tab.s[4].soff);
    const float* dhd = reinterpret_cast<const float*>(sg + tab.s[5].soff);
    // ---- LSTM cell backward ----
    for (int e = tid; e < H * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const float gi = gat[b * G4 + 0 * H + j], gf = gat[b * G4 + 1 * H + j];
      const float gg = gat[b * G4 + 2 * H + j], go = gat[b * G4 + 3 * H + j];
      const float c_prev = cpv[b * H + j], c_new = cnv[b * H + j];
      const float tc = tanhf(c_new);
      const float dhv = dh[e] + dhd[b * H + j];
      const float dcv = dc[e] + dcn[b * H + j] + dhv * go * (1.f - tc * tc);
      dz[(0 * H + j) * BT + b] = dcv * gg * gi * (1.f - gi);
      dz[(1 * H + j) * BT + b] = dcv * c_prev * gf * (1.f - gf);
      dz[(2 * H + j) * BT + b] = dcv * gi * (1.f - gg * gg);
      dz[(3 * H + j) * BT + b] = dhv * tc * go * (1.f - go);
      dc[e] = dcp[b * H + j] + dcv * gf;                          // gradient wrt c_{t-1}
    }
    cta_sync();
    // ---- d h_{t-1} = W_hh^T dz ----
    dense_s<WT>(Ws, G4, H, [&](int) { return dz; }, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dh[n * BT + b] = acc[b];
    });
    for (int e = tid; e < nb * G4; e += NTHREADS) {
      const int b = e / G4, f = e - b * G4;
      st_op(dz_op + rw[b] * (4 * Hs) + 4 * hoff + f, dz[f * BT + b]);
    }
  }
  cp_wait<0>();
}

// ======================================================================================================
// F4 / B2: memory recurrences (one CTA per narrative tile)
// ======================================================================================================
template <typename WT, typename ST>
__global__ void __launch_bounds__(NTHREADS, 1) mfn_mem_fwd_kernel(const __grid_constant__ MemArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int MEM = a.MEM, G = a.G, G2 = 2 * a.G, M2 = 2 * a.MEM, H2 = 2 * a.Hs, Hs = a.Hs;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  char* sp = reinterpret_cast<char*>(smem_raw);
  char* stages = sp; sp += (size_t)NSTAGE * BT * (G2 + MEM) * sizeof(float);
  float* mem = reinterpret_cast<float*>(sp); sp += MEM * BT * sizeof(float);
  float* mp = reinterpret_cast<float*>(sp); sp += MEM * BT * sizeof(float);
  float* gh = reinterpret_cast<float*>(sp); sp += G2 * BT * sizeof(float);
  float* gm = reinterpret_cast<float*>(sp); sp += M2 * BT * sizeof(float);
  float* part = reinterpret_cast<float*>(sp); sp += smax(part_floats(MEM, G2), part_floats(G, M2)) * sizeof(float);
  WT* Wm = reinterpret_cast<WT*>(sp); sp += (size_t)MEM * G2 * sizeof(WT);       // [MEM][2G]: mem part of gamma{1,2}_fc1, transposed
  WT* W2 = reinterpret_cast<WT*>(sp);                                            // [G][2MEM]: gamma1_fc2^T | gamma2_fc2^T
  __shared__ long long rows[2][BT];
  __shared__ SegTab tab;
  if (tid == 0) {
    tab.n = 0; tab.stage_bytes = 0; tab.chunks = 0;
    tab.add(a.gpre, (long long)G2 * sizeof(float), G2 * (int)sizeof(float));
    tab.add(a.chat, (long long)MEM * sizeof(float), MEM * (int)sizeof(float));
  }
  {
    const WT* w1 = reinterpret_cast<const WT*>(a.g1_fc1_w);
    const WT* w2 = reinterpret_cast<const WT*>(a.g2_fc1_w);
    const int ld = H2 + MEM;
    for (int e = tid; e < G2 * MEM; e += NTHREADS) {
      const int n = e / MEM, k = e - n * MEM;
      Wm[k * G2 + n] = n < G ? w1[(size_t)n * ld + H2 + k] : w2[(size_t)(n - G) * ld + H2 + k];
    }
    const WT* v1 = reinterpret_cast<const WT*>(a.g1_fc2_w);
    const WT* v2 = reinterpret_cast<const WT*>(a.g2_fc2_w);
    for (int e = tid; e < M2 * G; e += NTHREADS) {
      const int n = e / G, k = e - n * G;
      W2[k * M2 + n] = n < MEM ? v1[(size_t)n * G + k] : v2[(size_t)(n - MEM) * G + k];
    }
  }
  for (int e = tid; e < MEM * BT; e += NTHREADS) mem[e] = 0.f;
  __syncthreads();
  const int stage_bytes = tab.stage_bytes;
  for (int i = 0; i < NSTAGE - 1; ++i) {
    if (i < a.T) prefetch_rows(tab, stages + (size_t)i * stage_bytes, b0, nb, a.sb, a.st, i);
    cp_commit();
  }
  const DropCfg drop_g1 = mt_drop_resolve(a.drop_g1), drop_g2 = mt_drop_resolve(a.drop_g2);
  ST* gh_op = reinterpret_cast<ST*>(a.gh_op);
  ST* memprev_op = reinterpret_cast<ST*>(a.memprev_op);
  ST* last_op = reinterpret_cast<ST*>(a.last_op);
  const int LW = Hs + MEM;

  for (int t = 0; t < a.T; ++t) {
    long long* rw = rows[t & 1];
    if (tid < BT) rw[tid] = (long long)(b0 + min(tid, nb - 1)) * a.sb + (long long)t * a.st;
    cp_wait<NSTAGE - 2>();
    cta_sync();
    if (t + NSTAGE - 1 < a.T) prefetch_rows(tab, stages + (size_t)((t + NSTAGE - 1) % NSTAGE) * stage_bytes, b0, nb, a.sb, a.st, t + NSTAGE - 1);
    cp_commit();
    const char* sg = stages + (size_t)(t % NSTAGE) * stage_bytes;
    const float* gp = reinterpret_cast<const float*>(sg + tab.s[0].soff);      // [b][2G]
    const float* ch = reinterpret_cast<const float*>(sg + tab.s[1].soff);      // [b][MEM]
    // ---- gamma hidden = drop(relu(gpre_t + W_gm mem_{t-1})) ----
    dense_s<WT>(Wm, MEM, G2, [&](int) { return mem; }, part, [&](int n, float* acc) {
      const bool second = n >= G;
      const DropCfg& dc = second ? drop_g2 : drop_g1;
      const int j = second ? n - G : n;
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = fmaxf(acc[b] + gp[b * G2 + n], 0.f);
        // element index of the [T,B,G] tensor (oracle/mt_oracle.py:_drop_t)
        v *= mt_drop_factor(dc, ((uint64_t)t * a.B + (uint64_t)(b0 + b)) * (uint64_t)G + (uint64_t)j);
        gh[n * BT + b] = v;
      }
    });
    // ---- gamma{1,2} = sigmoid(gamma{1,2}_fc2 hidden) in one pass: outputs n < MEM read gh1, the rest gh2 ----
    dense_s<WT>(W2, G, M2, [&](int g) { return g * VN < MEM ? gh : gh + G * BT; }, part, [&](int n, float* acc) {
      const float bias = n < MEM ? a.g1_fc2_b[n] : a.g2_fc2_b[n - MEM];
#pragma unroll
      for (int b = 0; b < BT; ++b) gm[n * BT + b] = sigmoidf_(acc[b] + bias);
    });
    // ---- mem_t = gamma1 * mem_{t-1} + gamma2 * cHat_t ----
    for (int e = tid; e < MEM * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const float mo = mem[e];
      mp[e] = mo;
      mem[e] = gm[e] * mo + gm[MEM * BT + e] * ch[b * MEM + j];
    }
    cta_sync();
    if (a.training) {
      for (int e = tid; e < nb * G2; e += NTHREADS) { const int b = e / G2, f = e - b * G2; st_op(gh_op + rw[b] * G2 + f, gh[f * BT + b]); }
      for (int e = tid; e < nb * M2; e += NTHREADS) { const int b = e / M2, f = e - b * M2; a.gm[rw[b] * M2 + f] = gm[f * BT + b]; }
      for (int e = tid; e < nb * MEM; e += NTHREADS) { const int b = e / MEM, f = e - b * MEM; st_op(memprev_op + rw[b] * MEM + f, mp[f * BT + b]); }
    }
    for (int e = tid; e < nb * MEM; e += NTHREADS) { const int b = e / MEM, f = e - b * MEM; st_op(last_op + rw[b] * LW + Hs + f, mem[f * BT + b]); }
  }
  cp_wait<0>();
  if (a.mem_last)
    for (int e = tid; e < nb * MEM; e += NTHREADS) { const int b = e / MEM, j = e % MEM; a.mem_last[(size_t)(b0 + b) * MEM + j] = mem[j * BT + b]; }
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return to_f(*p); }

template <typename WT, typename ST>
__global__ void __launch_bounds__(NTHREADS, 1) mfn_mem_bwd_kernel(const __grid_constant__ MemArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int MEM = a.MEM, G = a.G, G2 = 2 * a.G, M2 = 2 * a.MEM, H2 = 2 * a.Hs, Hs = a.Hs;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT, nb = min(BT, a.B - b0);
  char* sp = reinterpret_cast<char*>(smem_raw);
  const int seg_bytes = (MEM + M2 + MEM) * (int)sizeof(float) + (MEM + G2) * (int)sizeof(ST);
  char* stages = sp; sp += (size_t)NSTAGE * BT * seg_bytes;
  float* dmem = reinterpret_cast<float*>(sp); sp += MEM * BT * sizeof(float);
  float* dzg = reinterpret_cast<float*>(sp); sp += M2 * BT * sizeof(float);
  float* dzchat = reinterpret_cast<float*>(sp); sp += MEM * BT * sizeof(float);
  float* dgh = reinterpret_cast<float*>(sp); sp += G2 * BT * sizeof(float);
  float* part = reinterpret_cast<float*>(sp); sp += smax(part_floats(MEM, G2), part_floats(G2, MEM)) * sizeof(float);
  WT* Wb2 = reinterpret_cast<WT*>(sp); sp += (size_t)MEM * G2 * sizeof(WT);      // [MEM][2G]: row k = gamma1_fc2[k][:] | gamma2_fc2[k][:]
  WT* Wbm = reinterpret_cast<WT*>(sp);                                           // [2G][MEM]: mem columns of gamma1_fc1 ; gamma2_fc1
  __shared__ long long rows[2][BT];
  __shared__ SegTab tab;
  if (tid == 0) {
    tab.n = 0; tab.stage_bytes = 0; tab.chunks = 0;
    tab.add(a.dlast + Hs, (long long)(Hs + MEM) * sizeof(float), MEM * (int)sizeof(float));                      // 0: d mem_t from the head
    tab.add(a.gm, (long long)M2 * sizeof(float), M2 * (int)sizeof(float));                                       // 1: gamma1 | gamma2
    tab.add(a.chat, (long long)MEM * sizeof(float), MEM * (int)sizeof(float));                                   // 2: cHat
    tab.add(a.memprev_op, (long long)MEM * sizeof(ST), MEM * (int)sizeof(ST));                                   // 3: mem_{t-1}
    tab.add(a.gh_op, (long long)G2 * sizeof(ST), G2 * (int)sizeof(ST));                                          // 4: gamma hidden
  }
  {
    const WT* v1 = reinterpret_cast<const WT*>(a.g1_fc2_w);
    const WT* v2 = reinterpret_cast<const WT*>(a.g2_fc2_w);
    for (int e = tid; e < MEM * G2; e += NTHREADS) {
      const int k = e / G2, n = e - k * G2;
      Wb2[e] = n < G ? v1[(size_t)k * G + n] : v2[(size_t)k * G + n - G];
    }
    const WT* w1 = reinterpret_cast<const WT*>(a.g1_fc1_w);
    const WT* w2 = reinterpret_cast<const WT*>(a.g2_fc1_w);
    const int ld = H2 + MEM;
    for (int e = tid; e < G2 * MEM; e += NTHREADS) {
      const int k = e / MEM, n = e - k * MEM;
      Wbm[e] = k < G ? w1[(size_t)k * ld + H2 + n] : w2[(size_t)(k - G) * ld + H2 + n];
    }
  }
  for (int e = tid; e < MEM * BT; e += NTHREADS) dmem[e] = 0.f;
  __syncthreads();
  const int stage_bytes = tab.stage_bytes;
  for (int i = 0; i < NSTAGE - 1; ++i) {
    if (i < a.T) prefetch_rows(tab, stages + (size_t)i * stage_bytes, b0, nb, a.sb, a.st, a.T - 1 - i);
    cp_commit();
  }
  ST* dzg_op = reinterpret_cast<ST*>(a.dzg_op);
  ST* dzchat_op = reinterpret_cast<ST*>(a.dzchat_op);
  ST* dgh_op = reinterpret_cast<ST*>(a.dgh_op);
  const float sc_g = a.drop_g1.scale;

  for (int i = 0; i < a.T; ++i) {
    const int t = a.T - 1 - i;
    long long* rw = rows[i & 1];
    if (tid < BT) rw[tid] = (long long)(b0 + min(tid, nb - 1)) * a.sb + (long long)t * a.st;
    cp_wait<NSTAGE - 2>();
    cta_sync();
    if (i + NSTAGE - 1 < a.T)
      prefetch_rows(tab, stages + (size_t)((i + NSTAGE - 1) % NSTAGE) * stage_bytes, b0, nb, a.sb, a.st, t - (NSTAGE - 1));
    cp_commit();
    const char* sg = stages + (size_t)(i % NSTAGE) * stage_bytes;
    const float* dlm = reinterpret_cast<const float*>(sg + tab.s[0].soff);     // [b][MEM]
    const float* gmm = reinterpret_cast<const float*>(sg + tab.s[1].soff);     // [b][2MEM]
    const float* chs = reinterpret_cast<const float*>(sg + tab.s[2].soff);     // [b][MEM]
    const ST* mps = reinterpret_cast<const ST*>(sg + tab.s[3].soff);           // [b][MEM]
    const ST* ghs = reinterpret_cast<const ST*>(sg + tab.s[4].soff);           // [b][2G]
    // ---- mem_t = g1 * mem_{t-1} + g2 * cHat ----
    for (int e = tid; e < MEM * BT; e += NTHREADS) {
      const int j = e / BT, b = e % BT;
      const float g = dmem[e] + dlm[b * MEM + j];
      const float g1 = gmm[b * M2 + j], g2 = gmm[b * M2 + MEM + j];
      const float mprev = ldf(mps + b * MEM + j), ch = chs[b * MEM + j];
      dzg[e] = g * mprev * g1 * (1.f - g1);
      dzg[MEM * BT + e] = g * ch * g2 * (1.f - g2);
      dzchat[e] = g * g2 * (1.f - ch * ch);
      dmem[e] = g * g1;                                            // direct path to mem_{t-1}
    }
    cta_sync();
    // ---- d gamma hidden (both gates in one pass), gated by relu / dropout ----
    dense_s<WT>(Wb2, MEM, G2, [&](int g) { return g * VN < G ? dzg : dzg + MEM * BT; }, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dgh[n * BT + b] = ldf(ghs + b * G2 + n) > 0.f ? acc[b] * sc_g : 0.f;
    });
    // ---- d mem_{t-1} += gamma_fc1[:, 2H:]^T d hidden ----
    dense_s<WT>(Wbm, G2, MEM, [&](int) { return dgh; }, part, [&](int n, float* acc) {
#pragma unroll
      for (int b = 0; b < BT; ++b) dmem[n * BT + b] += acc[b];
    });
    for (int e = tid; e < nb * M2; e += NTHREADS) { const int b = e / M2, f = e - b * M2; st_op(dzg_op + rw[b] * M2 + f, dzg[f * BT + b]); }
    for (int e = tid; e < nb * MEM; e += NTHREADS) { const int b = e / MEM, f = e - b * MEM; st_op(dzchat_op + rw[b] * MEM + f, dzchat[f * BT + b]); }
    for (int e = tid; e < nb * G2; e += NTHREADS) { const int b = e / G2, f = e - b * G2; st_op(dgh_op + rw[b] * G2 + f, dgh[f * BT + b]); }
  }
  cp_wait<0>();
}

// ======================================================================================================
// row-wise kernels of the batched part (one warp per (b,t) row)
// ======================================================================================================
// att = softmax over the 2H FEATURES of a row (MFT/multiTransformer.py:218), attended = att * cStar (:219)
template <typename ST>
__global__ void __launch_bounds__(256) mfn_softmax_attend_fwd_kernel(int M, int W, float* __restrict__ att, const float* __restrict__ cstar,
                                                                      ST* __restrict__ attd) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float* ar = att + row * W;
  const float* cr = cstar + row * W;
  ST* dr = attd + row * W;
  float mx = -INFINITY;
  for (int f = lane * 4; f < W; f += 128) { const float4 v = ld4(ar + f); mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int f = lane * 4; f < W; f += 128) { const float4 v = ld4(ar + f); sum += expf(v.x - mx) + expf(v.y - mx) + expf(v.z - mx) + expf(v.w - mx); }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int f = lane * 4; f < W; f += 128) {
    const float4 v = ld4(ar + f), cv = ld4(cr + f);
    const float4 p = make_float4(expf(v.x - mx) * inv, expf(v.y - mx) * inv, expf(v.z - mx) * inv, expf(v.w - mx) * inv);
    st4(ar + f, p);
    st4(dr + f, make_float4(p.x * cv.x, p.y * cv.y, p.z * cv.z, p.w * cv.w));
  }
}

// in: datt = d attended.  out: dlogit = att * (datt * cStar - sum(datt * cStar * att)),  datt <- d cStar = datt * att
template <typename ST>
__global__ void __launch_bounds__(256) mfn_softmax_attend_bwd_kernel(int M, int W, const float* __restrict__ att, const float* __restrict__ cstar,
                                                                      float* __restrict__ datt, ST* __restrict__ dlogit) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* ar = att + row * W;
  const float* cr = cstar + row * W;
  float* gr = datt + row * W;
  ST* lr = dlogit + row * W;
  float dot = 0.f;
  for (int f = lane * 4; f < W; f += 128) {
    const float4 p = ld4(ar + f), cv = ld4(cr + f), g = ld4(gr + f);
    dot += g.x * cv.x * p.x + g.y * cv.y * p.y + g.z * cv.z * p.z + g.w * cv.w * p.w;
  }
  dot = warp_sum(dot);
  for (int f = lane * 4; f < W; f += 128) {
    const float4 p = ld4(ar + f), cv = ld4(cr + f), g = ld4(gr + f);
    st4(lr + f, make_float4(p.x * (g.x * cv.x - dot), p.y * (g.y * cv.y - dot), p.z * (g.z * cv.z - dot), p.w * (g.w * cv.w - dot)));
    st4(gr + f, make_float4(g.x * p.x, g.y * p.y, g.z * p.z, g.w * p.w));
  }
}

// head: y = out_fc2(drop(pre)) * mask, pre = relu(out_fc1 [h || mem] + b) from the GEMM   (MFT/multiTransformer.py:239-246,310)
template <typename ST>
__global__ void __launch_bounds__(256) mfn_head_fwd_kernel(int B, int T, long long sb, long long st, int O, const float* __restrict__ pre,
                                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                                            const float* __restrict__ mask, DropCfg drop_in, ST* __restrict__ oh,
                                                            float* __restrict__ out) {
  const DropCfg drop = mt_drop_resolve(drop_in);
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // i = b*T + t
  if (i >= (long long)B * T) return;
  const int b = (int)(i / T), t = (int)(i - (long long)b * T);
  const long long row = (long long)b * sb + (long long)t * st;
  float acc = 0.f;
  for (int n = lane; n < O; n += 32) {
    float v = pre[row * O + n];
    v *= mt_drop_factor(drop, ((uint64_t)t * B + (uint64_t)b) * (uint64_t)O + (uint64_t)n);      // [T,B,O] element index
    if (oh) st_op(oh + row * O + n, v);
    acc = fmaf(v, w2[n], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float y = acc + b2[0];
    if (mask) y *= mask[i];
    out[i] = y;
  }
}

// The same with LPR = O / 4 lanes per row (O = 32 / 64 / 128): 16-byte loads of pre, 32 / LPR rows per warp instruction, the dot product
// closed by a butterfly inside the row's lane group (the one-row-per-warp kernel above moves 64-256 bytes per warp instruction).
template <typename ST, int LPR>
__global__ void __launch_bounds__(256) mfn_head_fwd_vec_kernel(int B, int T, long long sb, long long st, const float* __restrict__ pre,
                                                                const float* __restrict__ w2, const float* __restrict__ b2,
                                                                const float* __restrict__ mask, DropCfg drop_in, ST* __restrict__ oh,
                                                                float* __restrict__ out) {
  constexpr int O = 4 * LPR, RPW = 32 / LPR;
  const DropCfg drop = mt_drop_resolve(drop_in);
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsub = lane / LPR;
  const float4 w = ld4(w2 + sub * 4);
  const float bias = b2[0];
  const long long total = (long long)B * T;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * RPW;
  for (long long i0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; i0 < total; i0 += wstride) {      // warp-uniform trip count
    const long long i = i0 + rsub;
    const bool ok = i < total;
    float acc = 0.f;
    if (ok) {
      const int b = (int)(i / T), t = (int)(i - (long long)b * T);
      const long long row = (long long)b * sb + (long long)t * st;
      float4 v = ld4(pre + row * O + sub * 4);
      const uint64_t e0 = ((uint64_t)t * B + (uint64_t)b) * (uint64_t)O + (uint64_t)(sub * 4);      // [T,B,O] element index (even)
      float f0, f1, f2, f3;
      mt_drop_pair(drop, e0, f0, f1);
      mt_drop_pair(drop, e0 + 2, f2, f3);
      v.x *= f0; v.y *= f1; v.z *= f2; v.w *= f3;
      if (oh) st4(oh + row * O + sub * 4, v);
      acc = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, v.w * w.w)));
    }
#pragma unroll
    for (int m = LPR / 2; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (ok && sub == 0) {
      float y = acc + bias;
      if (mask) y *= mask[i];
      out[i] = y;
    }
  }
}

// dzoh = (oh > 0) * dy * w2 * keep_scale; dW_o2 += sum dy * oh; db_o2 += sum dy   (dy = dout * mask)
template <typename ST>
__global__ void __launch_bounds__(256) mfn_head_bwd_kernel(int B, int T, long long sb, long long st, int O, const float* __restrict__ dout,
                                                            const float* __restrict__ mask, const ST* __restrict__ oh,
                                                            const float* __restrict__ w2, float scale, ST* __restrict__ dzoh,
                                                            float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ float red[];                 // [O + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = threadIdx.x; e <= O; e += blockDim.x) red[e] = 0.f;
  __syncthreads();
  const long long total = (long long)B * T;
  float dbs = 0.f;
  for (int n0 = 0; n0 < O; n0 += 32) {
    const int n = n0 + lane;
    float dws = 0.f;
    const float wn = n < O ? w2[n] : 0.f;
    for (long long i = (long long)blockIdx.x * nw + warp; i < total; i += (long long)gridDim.x * nw) {
      const int b = (int)(i / T), t = (int)(i - (long long)b * T);
      const long long row = (long long)b * sb + (long long)t * st;
      float dy = dout[i];
      if (mask) dy *= mask[i];
      if (n0 == 0 && lane == 0) dbs += dy;
      if (n < O) {
        const float o = to_f(oh[row * O + n]);
        dws = fmaf(dy, o, dws);
        st_op(dzoh + row * O + n, o > 0.f ? dy * wn * scale : 0.f);
      }
    }
    if (n < O) atomicAdd(&red[n], dws);
  }
  if (lane == 0) atomicAdd(&red[O], dbs);
  __syncthreads();
  for (int e = threadIdx.x; e < O; e += blockDim.x) atomicAdd(&dw2[e], red[e]);
  if (threadIdx.x == 0) atomicAdd(db2, red[O]);
}

// The same, 16 bytes per thread (O % E == 0, E = 8 bf16 / 4 fp32 columns): O / E adjacent lanes cover a row, so a warp instruction moves
// 512 contiguous bytes of oh / dzoh instead of 64, every thread keeps its E columns of dW_o2 in registers over its rows, and dy is
// loaded once per row (not once per 32-column block).  The one-row-per-warp kernel above ran at 0.45 TB/s at M = 262 144.
template <typename ST>
__global__ void __launch_bounds__(256) mfn_head_bwd_vec_kernel(int B, int T, long long sb, long long st, int O, const float* __restrict__ dout,
                                                                const float* __restrict__ mask, const ST* __restrict__ oh,
                                                                const float* __restrict__ w2, float scale, ST* __restrict__ dzoh,
                                                                float* __restrict__ dw2, float* __restrict__ db2) {
  constexpr int E = 16 / (int)sizeof(ST);
  extern __shared__ float red[];                 // [O + 1]
  for (int e = threadIdx.x; e <= O; e += blockDim.x) red[e] = 0.f;
  __syncthreads();
  const int cg = O / E;                          // column groups per row
  const int total = B * T;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = tid % cg;                        // this thread's column group, the same for all its rows (the stride is a multiple of cg)
  const int rows_per_pass = (gridDim.x * blockDim.x) / cg;
  float wv[E], dws[E];
#pragma unroll
  for (int k = 0; k < E; ++k) { wv[k] = w2[c * E + k] * scale; dws[k] = 0.f; }
  float dbs = 0.f;
  if (tid < rows_per_pass * cg) {
    for (int i = tid / cg; i < total; i += rows_per_pass) {
      const int b = i / T, t = i - b * T;
      const long long row = (long long)b * sb + (long long)t * st;
      float dy = dout[i];
      if (mask) dy *= mask[i];
      if (c == 0) dbs += dy;
      const ST* src = oh + row * O + c * E;
      ST* dst = dzoh + row * O + c * E;
      float o[E], z[E];
      if (E == 8) {
        const float4 lo = ld4(src), hi = ld4(src + 4);
        o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = lo.w; o[E - 4] = hi.x; o[E - 3] = hi.y; o[E - 2] = hi.z; o[E - 1] = hi.w;
      } else {
        const float4 lo = ld4(src);
        o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = lo.w;
      }
#pragma unroll
      for (int k = 0; k < E; ++k) {
        dws[k] = fmaf(dy, o[k], dws[k]);
        z[k] = o[k] > 0.f ? dy * wv[k] : 0.f;
      }
      st4(dst, make_float4(z[0], z[1], z[2], z[3]));
      if (E == 8) st4(dst + 4, make_float4(z[E - 4], z[E - 3], z[E - 2], z[E - 1]));
    }
  }
#pragma unroll
  for (int k = 0; k < E; ++k) atomicAdd(&red[c * E + k], dws[k]);
  if (c == 0) atomicAdd(&red[O], dbs);
  __syncthreads();
  for (int e = threadIdx.x; e < O; e += blockDim.x) atomicAdd(&dw2[e], red[e]);
  if (threadIdx.x == 0) atomicAdd(db2, red[O]);
}

// ======================================================================================================
// host side
// ======================================================================================================
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

size_t lstm_fwd_smem(int H, size_t ws) {
  const int G4 = 4 * H;
  return (size_t)NSTAGE * BT * G4 * 4 + (size_t)4 * H * BT * 4 + (size_t)G4 * BT * 4 + part_floats(H, G4) * 4 + (size_t)G4 * H * ws;
}
size_t lstm_bwd_smem(int H, size_t ws) {
  const int G4 = 4 * H;
  return (size_t)NSTAGE * BT * (G4 + 5 * H) * 4 + (size_t)2 * H * BT * 4 + (size_t)G4 * BT * 4 + part_floats(G4, H) * 4 + (size_t)G4 * H * ws;
}
size_t mem_fwd_smem(const Dims& D, size_t ws) {
  const int MEM = D.MEM, G2 = 2 * D.G, M2 = 2 * D.MEM;
  const size_t part = part_floats(MEM, G2) > part_floats(D.G, M2) ? part_floats(MEM, G2) : part_floats(D.G, M2);
  return (size_t)NSTAGE * BT * (G2 + MEM) * 4 + (size_t)(2 * MEM + G2 + M2) * BT * 4 + part * 4 + ((size_t)MEM * G2 + (size_t)D.G * M2) * ws;
}
size_t mem_bwd_smem(const Dims& D, size_t ws, size_t es) {
  const int MEM = D.MEM, G2 = 2 * D.G, M2 = 2 * D.MEM;
  const size_t part = part_floats(MEM, G2) > part_floats(G2, MEM) ? part_floats(MEM, G2) : part_floats(G2, MEM);
  const size_t seg = (size_t)(MEM + M2 + MEM) * 4 + (size_t)(MEM + G2) * es;
  return (size_t)NSTAGE * BT * seg + (size_t)(MEM + M2 + MEM + G2) * BT * 4 + part * 4 + ((size_t)MEM * G2 + (size_t)G2 * MEM) * ws;
}

inline const void* wsel(bool lp, const float* params, const void* params_lp, size_t off) {
  return lp ? (const void*)((const bf16*)params_lp + off) : (const void*)(params + off);
}
inline void* op_off(bool lp, void* p, size_t off) { return lp ? (void*)((bf16*)p + off) : (void*)((float*)p + off); }
inline const void* op_off(bool lp, const void* p, size_t off) { return lp ? (const void*)((const bf16*)p + off) : (const void*)((const float*)p + off); }

// y[M,N] (ldc) = act(A[M,K] (lda) W[N,K]^T (ldb) + bias)
GemmDesc lin_fwd(int M, int N, int K, const void* A, int lda, const void* W, int ldb, void* C, int ldc, bool c_f32, const float* bias, int act) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_kmajor = true;
  g.B = W; g.ldb = ldb; g.b_kmajor = true;
  g.C = C; g.ldc = ldc; g.c_f32 = c_f32;
  g.epi.bias = bias; g.epi.act = act;
  return g;
}
// dx[M,Kin] (ldc) = dy[M,Nout] (lda) W[Nout, :Kin] (ldb)
GemmDesc lin_dgrad(int M, int Nout, int Kin, const void* dy, int lda, const void* W, int ldb, void* C, int ldc, bool c_f32) {
  GemmDesc g;
  g.M = M; g.N = Kin; g.K = Nout;
  g.A = dy; g.lda = lda; g.a_kmajor = true;
  g.B = W; g.ldb = ldb; g.b_kmajor = false;
  g.C = C; g.ldc = ldc; g.c_f32 = c_f32;
  return g;
}

}  // namespace

static int g_mfn_force_ffma = 0;

extern "C" {

/* test hook: run the bf16 recurrences on the FFMA kernels instead of the tensor-core ones; returns the previous setting */
int mt_mfn_force_ffma(int on) { int old = g_mfn_force_ffma; g_mfn_force_ffma = on; return old; }

size_t mt_mfn_param_count(const MtMfnCfg* cfg) {
  Dims D;
  if (!cfg || make_dims(*cfg, D) != MT_OK) return 0;
  return D.total;
}

size_t mt_mfn_ws_bytes(const MtMfnCfg* cfg) {
  Dims D;
  if (!cfg || make_dims(*cfg, D) != MT_OK) return 0;
  if (cfg->dtype != MT_F32 && cfg->dtype != MT_BF16) return 0;
  const size_t ws = mt_esize(cfg->dtype);
  for (int m = 0; m < D.n_mods; ++m)
    if (lstm_fwd_smem(D.H[m], ws) > 227 * 1024 || lstm_bwd_smem(D.H[m], ws) > 227 * 1024) return 0;
  if (mem_fwd_smem(D, ws) > 227 * 1024 || mem_bwd_smem(D, ws, ws) > 227 * 1024) return 0;
  Stash s;
  carve(*cfg, D, nullptr, s);
  return s.bytes;
}


#define MT_REC_LAUNCH(KERNEL, GRID, SMEM, ARGS)                                                        \
  do {                                                                                                 \
    if (lp) { MT_TRY(set_smem(KERNEL<bf16, bf16>, SMEM)); KERNEL<bf16, bf16><<<GRID, NTHREADS, SMEM, st>>>(ARGS); } \
    else { MT_TRY(set_smem(KERNEL<float, float>, SMEM)); KERNEL<float, float><<<GRID, NTHREADS, SMEM, st>>>(ARGS); } \
    MT_LAUNCH_CHECK();                                                                                 \
  } while (0)

static void fill_lstm_args(LstmArgs& a, const MtMfnCfg& c, const Dims& D, const Stash& S, const float* params, const void* params_lp,
                           long long sb, long long stt) {
  const bool lp = c.dtype == MT_BF16;
  a.B = c.B; a.T = c.T; a.n_mods = D.n_mods; a.sb = sb; a.st = stt;
  a.Hs = D.Hs; a.MEM = D.MEM;
  for (int m = 0; m < D.n_mods; ++m) {
    a.H[m] = D.H[m]; a.hoff[m] = D.hoff[m];
    a.w_hh[m] = wsel(lp, params, params_lp, D.w_hh[m]);
    a.b_hh[m] = params + D.b_hh[m];
  }
  a.gates = S.gates; a.cstar = S.cstar; a.cstar_op = lp ? S.cstar_op : nullptr; a.last_op = S.last_op; a.hprev_op = S.hprev_op;
  a.h_last = a.c_last = nullptr;
  a.training = c.training;
  a.dlast = S.dlast; a.dcstar = S.datt; a.dz_op = S.dz_op;
  a.dbg = g_mt_tune[MT_TUNE_REC_DEBUG];
}

static void fill_mem_args(MemArgs& a, const MtMfnCfg& c, const Dims& D, const Stash& S, const float* params, const void* params_lp,
                          long long sb, long long stt) {
  const bool lp = c.dtype == MT_BF16;
  a.B = c.B; a.T = c.T; a.sb = sb; a.st = stt;
  a.Hs = D.Hs; a.MEM = D.MEM; a.G = D.G;
  a.g1_fc1_w = wsel(lp, params, params_lp, D.g1_fc1.w); a.g2_fc1_w = wsel(lp, params, params_lp, D.g2_fc1.w);
  a.g1_fc2_w = wsel(lp, params, params_lp, D.g1_fc2.w); a.g2_fc2_w = wsel(lp, params, params_lp, D.g2_fc2.w);
  a.g1_fc2_b = params + D.g1_fc2.b; a.g2_fc2_b = params + D.g2_fc2.b;
  a.gpre = S.gpre; a.chat = S.chat; a.gh_op = S.gh_op; a.gm = S.gm; a.memprev_op = S.memprev_op; a.last_op = S.last_op;
  a.mem_last = nullptr;
  a.drop_g1 = mt_make_drop(c.p_gamma, c.seed, MT_SITE_MFN_G1);
  a.drop_g2 = mt_make_drop(c.p_gamma, c.seed, MT_SITE_MFN_G2);
  a.training = c.training;
  a.dlast = S.dlast; a.dzg_op = S.dzg_op; a.dzchat_op = S.dzchat_op; a.dgh_op = S.dgh_op;
  a.dbg = g_mt_tune[MT_TUNE_REC_DEBUG];
}

int mt_mfn_fwd(const MtMfnCfg* cfg, const float* params, const void* params_lp, const void* const* x, const int64_t* stride_b,
               const int64_t* stride_t, const float* mask, float* out, float* h_last, float* c_last, float* mem_last, void* ws,
               size_t ws_bytes, void* stream) {
  if (!cfg || !params || !x || !stride_b || !stride_t || !out || !ws) return MT_ERR_ARG;
  const MtMfnCfg& c = *cfg;
  if (c.dtype != MT_F32 && c.dtype != MT_BF16) return MT_ERR_ARG;
  if (c.dtype == MT_BF16 && !params_lp) return MT_ERR_ARG;
  Dims D;
  MT_TRY(make_dims(c, D));
  if (mt_mfn_ws_bytes(cfg) == 0) return MT_ERR_UNSUPPORTED;
  Stash S;
  carve(c, D, ws, S);
  if (ws_bytes < S.bytes) return MT_ERR_WS;
  long long sb, stt;
  MT_TRY(layout_strides(c, D, stride_b, stride_t, sb, stt));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = c.B * c.T;
  const bool lp = c.dtype == MT_BF16;
  const size_t wsz = mt_esize(c.dtype);
  const int Hs = D.Hs, H2 = 2 * D.Hs, MEM = D.MEM, G = D.G;
  const int tiles = (c.B + BT - 1) / BT;
  auto W = [&](size_t off) { return wsel(lp, params, params_lp, off); };

  // F1: hoisted input projections  gates[:, 4*hoff_m : +4H_m] = x_m W_ih^T + b_ih      (b_hh is added in the recurrence)
  for (int m = 0; m < D.n_mods; ++m)
    MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, 4 * D.H[m], D.D[m], x[m], D.D[m], W(D.w_ih[m]), D.D[m], S.gates + 4 * D.hoff[m], 4 * Hs, true,
                                        params + D.b_ih[m], MT_ACT_NONE), st));
  // F2: LSTM recurrence
  {
    LstmArgs a;
    fill_lstm_args(a, c, D, S, params, params_lp, sb, stt);
    a.h_last = h_last; a.c_last = c_last;
    size_t smem = 0;
    int Hq = 0;
    for (int m = 0; m < D.n_mods; ++m) { smem = smax(smem, lstm_fwd_smem(D.H[m], wsz)); Hq += D.H[m] * D.H[m]; }
    mt_prof_work(2.0 * 4.0 * Hq * (double)M, 4.0 * M * (double)(4 * Hs * 2 + 2 * H2 + Hs));
    if (lp && !g_mfn_force_ffma && mt_mfn_mma_lstm_supported(a)) MT_TRY(mt_mfn_mma_lstm_fwd(a, st));
    else MT_REC_LAUNCH(mfn_lstm_fwd_kernel, dim3(tiles, D.n_mods), smem, a);
  }
  // F3: delta-memory attention block, batched over all rows
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, D.A1, H2, S.cstar_op, H2, W(D.att1_fc1.w), H2, S.a1_op, D.A1, !lp, params + D.att1_fc1.b, MT_ACT_RELU), st));
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, H2, D.A1, S.a1_op, D.A1, W(D.att1_fc2.w), D.A1, S.att, H2, true, params + D.att1_fc2.b, MT_ACT_NONE), st));
  {
    const int wpb = 8;
    mt_prof_work(0.0, (double)M * H2 * (4.0 * 3 + wsz));
    if (lp) mfn_softmax_attend_fwd_kernel<bf16><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(M, H2, S.att, S.cstar, (bf16*)S.attd_op);
    else mfn_softmax_attend_fwd_kernel<float><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(M, H2, S.att, S.cstar, (float*)S.attd_op);
    MT_LAUNCH_CHECK();
  }
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, D.A2, H2, S.attd_op, H2, W(D.att2_fc1.w), H2, S.a2_op, D.A2, !lp, params + D.att2_fc1.b, MT_ACT_RELU), st));
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, MEM, D.A2, S.a2_op, D.A2, W(D.att2_fc2.w), D.A2, S.chat, MEM, true, params + D.att2_fc2.b, MT_ACT_TANH), st));
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, G, H2, S.attd_op, H2, W(D.g1_fc1.w), H2 + MEM, S.gpre, 2 * G, true, params + D.g1_fc1.b, MT_ACT_NONE), st));
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, G, H2, S.attd_op, H2, W(D.g2_fc1.w), H2 + MEM, S.gpre + G, 2 * G, true, params + D.g2_fc1.b, MT_ACT_NONE), st));
  // F4: memory recurrence
  {
    MemArgs a;
    fill_mem_args(a, c, D, S, params, params_lp, sb, stt);
    a.mem_last = mem_last;
    const size_t smem = mem_fwd_smem(D, wsz);
    mt_prof_work(2.0 * (2.0 * G * MEM + 2.0 * G * MEM) * (double)M, 4.0 * M * (double)(2 * G + MEM + 2 * MEM + MEM));
    if (lp && !g_mfn_force_ffma && mt_mfn_mma_mem_supported(a)) MT_TRY(mt_mfn_mma_mem_fwd(a, st));
    else MT_REC_LAUNCH(mfn_mem_fwd_kernel, dim3(tiles), smem, a);
  }
  // F5: head
  MT_TRY(mt_gemm_run(c.dtype, lin_fwd(M, D.O, Hs + MEM, S.last_op, Hs + MEM, W(D.out_fc1.w), Hs + MEM, S.pre, D.O, true, params + D.out_fc1.b, MT_ACT_RELU), st));
  {
    const int wpb = 8;
    const DropCfg dr = mt_make_drop(c.p_out, c.seed, MT_SITE_MFN_OUT);
    mt_prof_work(0.0, (double)M * D.O * (4.0 + wsz));
#define MT_HEADF(ST_, LPR_) mfn_head_fwd_vec_kernel<ST_, LPR_><<<gridv, 256, 0, st>>>(c.B, c.T, sb, stt, S.pre, params + D.out_fc2.w, params + D.out_fc2.b, mask, dr, (ST_*)S.oh_op, out)
    const bool vec_ok = (D.O == 32 || D.O == 64 || D.O == 128) && (((uintptr_t)S.pre | (uintptr_t)S.oh_op | (uintptr_t)(params + D.out_fc2.w)) & 15) == 0;
    if (vec_ok) {
      const int rpw = 128 / D.O, rows_per_cta = 8 * rpw;
      const int gridv = (int)std::min<long long>(((long long)M + rows_per_cta - 1) / rows_per_cta, 148 * 8);
      if (lp) { if (D.O == 32) MT_HEADF(bf16, 8); else if (D.O == 64) MT_HEADF(bf16, 16); else MT_HEADF(bf16, 32); }
      else { if (D.O == 32) MT_HEADF(float, 8); else if (D.O == 64) MT_HEADF(float, 16); else MT_HEADF(float, 32); }
    } else if (lp) mfn_head_fwd_kernel<bf16><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(c.B, c.T, sb, stt, D.O, S.pre, params + D.out_fc2.w, params + D.out_fc2.b, mask, dr, (bf16*)S.oh_op, out);
    else mfn_head_fwd_kernel<float><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(c.B, c.T, sb, stt, D.O, S.pre, params + D.out_fc2.w, params + D.out_fc2.b, mask, dr, (float*)S.oh_op, out);
#undef MT_HEADF
    MT_LAUNCH_CHECK();
  }
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
  if (mt_mfn_ws_bytes(cfg) == 0) return MT_ERR_UNSUPPORTED;
  Stash S;
  carve(c, D, ws, S);
  if (ws_bytes < S.bytes) return MT_ERR_WS;
  long long sb, stt;
  MT_TRY(layout_strides(c, D, stride_b, stride_t, sb, stt));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = c.B * c.T;
  const bool lp = c.dtype == MT_BF16;
  const size_t wsz = mt_esize(c.dtype);
  const int Hs = D.Hs, H2 = 2 * D.Hs, MEM = D.MEM, G = D.G;
  const int tiles = (c.B + BT - 1) / BT;
  auto W = [&](size_t off) { return wsel(lp, params, params_lp, off); };

  // cfg->bwd_phase: 0 = everything; 1 = B1-B4 and the input gradients (what the encoder stacks wait for); 2 = only the batched
  // weight / bias gradients B5 (independent of everything downstream: the caller may run it on another stream after phase 1)
  const int phase = c.bwd_phase;
  if (phase < 0 || phase > 2) return MT_ERR_ARG;
  if (phase != 2) {
  MT_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * D.total, st));
  // B1: head
  {
    const DropCfg dr = mt_make_drop(c.p_out, c.seed, MT_SITE_MFN_OUT);
    const int grid = 296;
    const size_t sm = (D.O + 1) * sizeof(float);
    mt_prof_work(0.0, (double)M * D.O * 2.0 * wsz);
    const int E = lp ? 8 : 4;
    if (D.O % E == 0 && D.O / E <= 256 && (long long)c.B * c.T < 0x7fffffffLL && (((uintptr_t)S.oh_op | (uintptr_t)S.dzoh_op) & 15) == 0) {
      const int gridv = 148 * 4;                 // 256 threads, 16 bytes each per pass: every thread keeps one column group
      if (lp) mfn_head_bwd_vec_kernel<bf16><<<gridv, 256, sm, st>>>(c.B, c.T, sb, stt, D.O, dout, mask, (const bf16*)S.oh_op, params + D.out_fc2.w, dr.scale, (bf16*)S.dzoh_op, grads + D.out_fc2.w, grads + D.out_fc2.b);
      else mfn_head_bwd_vec_kernel<float><<<gridv, 256, sm, st>>>(c.B, c.T, sb, stt, D.O, dout, mask, (const float*)S.oh_op, params + D.out_fc2.w, dr.scale, (float*)S.dzoh_op, grads + D.out_fc2.w, grads + D.out_fc2.b);
      MT_LAUNCH_CHECK();
    } else {
    if (lp) mfn_head_bwd_kernel<bf16><<<grid, 256, sm, st>>>(c.B, c.T, sb, stt, D.O, dout, mask, (const bf16*)S.oh_op, params + D.out_fc2.w, dr.scale, (bf16*)S.dzoh_op, grads + D.out_fc2.w, grads + D.out_fc2.b);
    else mfn_head_bwd_kernel<float><<<grid, 256, sm, st>>>(c.B, c.T, sb, stt, D.O, dout, mask, (const float*)S.oh_op, params + D.out_fc2.w, dr.scale, (float*)S.dzoh_op, grads + D.out_fc2.w, grads + D.out_fc2.b);
    MT_LAUNCH_CHECK();
    }
  }
  MT_TRY(mt_gemm_run(c.dtype, lin_dgrad(M, D.O, Hs + MEM, S.dzoh_op, D.O, W(D.out_fc1.w), Hs + MEM, S.dlast, Hs + MEM, true), st));
  // B2: reverse-time memory recurrence
  {
    MemArgs a;
    fill_mem_args(a, c, D, S, params, params_lp, sb, stt);
    const size_t smem = mem_bwd_smem(D, wsz, wsz);
    mt_prof_work(2.0 * (2.0 * G * MEM + 2.0 * G * MEM) * (double)M, 4.0 * M * (double)(6 * MEM + 2 * G) + wsz * M * (double)(4 * MEM + 4 * G));
    if (lp && !g_mfn_force_ffma && mt_mfn_mma_mem_supported(a)) MT_TRY(mt_mfn_mma_mem_bwd(a, st));
    else MT_REC_LAUNCH(mfn_mem_bwd_kernel, dim3(tiles), smem, a);
  }
  // B3: batched dgrads through the attention block
  {
    GemmDesc g = lin_dgrad(M, MEM, D.A2, S.dzchat_op, MEM, W(D.att2_fc2.w), D.A2, S.da2_op, D.A2, !lp);
    g.epi.gate = S.a2_op; g.epi.ldg = D.A2;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    MT_TRY(mt_gemm_run(c.dtype, lin_dgrad(M, D.A2, H2, S.da2_op, D.A2, W(D.att2_fc1.w), H2, S.datt, H2, true), st));
    g = lin_dgrad(M, G, H2, S.dgh_op, 2 * G, W(D.g1_fc1.w), H2 + MEM, S.datt, H2, true);
    g.epi.residual = S.datt; g.epi.ldr = H2;                       // in-place accumulate
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    g = lin_dgrad(M, G, H2, op_off(lp, (const void*)S.dgh_op, G), 2 * G, W(D.g2_fc1.w), H2 + MEM, S.datt, H2, true);
    g.epi.residual = S.datt; g.epi.ldr = H2;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    const int wpb = 8;
    mt_prof_work(0.0, (double)M * H2 * (4.0 * 4 + wsz));
    if (lp) mfn_softmax_attend_bwd_kernel<bf16><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(M, H2, S.att, S.cstar, S.datt, (bf16*)S.dlogit_op);
    else mfn_softmax_attend_bwd_kernel<float><<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(M, H2, S.att, S.cstar, S.datt, (float*)S.dlogit_op);
    MT_LAUNCH_CHECK();
    g = lin_dgrad(M, H2, D.A1, S.dlogit_op, H2, W(D.att1_fc2.w), D.A1, S.da1_op, D.A1, !lp);
    g.epi.gate = S.a1_op; g.epi.ldg = D.A1;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
    g = lin_dgrad(M, D.A1, H2, S.da1_op, D.A1, W(D.att1_fc1.w), H2, S.datt, H2, true);
    g.epi.residual = S.datt; g.epi.ldr = H2;
    MT_TRY(mt_gemm_run(c.dtype, g, st));
  }
  // B4: reverse-time LSTM recurrence
  {
    LstmArgs a;
    fill_lstm_args(a, c, D, S, params, params_lp, sb, stt);
    size_t smem = 0;
    int Hq = 0;
    for (int m = 0; m < D.n_mods; ++m) { smem = smax(smem, lstm_bwd_smem(D.H[m], wsz)); Hq += D.H[m] * D.H[m]; }
    mt_prof_work(2.0 * 4.0 * Hq * (double)M, 4.0 * M * (double)(4 * Hs + 2 * H2 + H2 + Hs) + wsz * M * 4.0 * Hs);
    if (lp && !g_mfn_force_ffma && mt_mfn_mma_lstm_supported(a)) MT_TRY(mt_mfn_mma_lstm_bwd(a, st));
    else MT_REC_LAUNCH(mfn_lstm_bwd_kernel, dim3(tiles, D.n_mods), smem, a);
  }
  for (int m = 0; m < D.n_mods; ++m) {          // input gradients first: the encoder stacks' backward starts from them
    const void* dzm = op_off(lp, (const void*)S.dz_op, 4 * D.hoff[m]);
    if (dx && dx[m]) MT_TRY(mt_gemm_run(c.dtype, lin_dgrad(M, 4 * D.H[m], D.D[m], dzm, 4 * Hs, W(D.w_ih[m]), D.D[m], dx[m], D.D[m], !lp), st));
  }
  }                                             // phase != 2
  if (phase == 1) return MT_OK;
  // B5: batched weight gradients over all T*B rows
  auto wg = [&](const void* dz, int ldz, int Nout, const void* xin, int ldx, int Kin, size_t w_off, int ldw) -> int {
    return mt_gemm_run(c.dtype, mt_wgrad_desc(M, Nout, Kin, dz, ldz, xin, ldx, grads + w_off, ldw), st);
  };
  // bias gradients = column sums of the pre-activation gradients: collected and issued as ONE launch at the end
  ColsumJob cjobs[MT_COLSUM_MAX_JOBS];
  int n_cjobs = 0;
  auto bg = [&](const void* dz, int ldz, int Nout, size_t b_off) -> int {
    if (n_cjobs == MT_COLSUM_MAX_JOBS) return MT_ERR_ARG;
    cjobs[n_cjobs++] = ColsumJob{dz, ldz, Nout, grads + b_off};
    return MT_OK;
  };
  for (int m = 0; m < D.n_mods; ++m) {
    const void* dzm = op_off(lp, (const void*)S.dz_op, 4 * D.hoff[m]);
    MT_TRY(wg(dzm, 4 * Hs, 4 * D.H[m], x[m], D.D[m], D.D[m], D.w_ih[m], D.D[m]));
    MT_TRY(wg(dzm, 4 * Hs, 4 * D.H[m], op_off(lp, (const void*)S.hprev_op, D.hoff[m]), Hs, D.H[m], D.w_hh[m], D.H[m]));
    MT_TRY(bg(dzm, 4 * Hs, 4 * D.H[m], D.b_ih[m]));
  }
  MT_TRY(wg(S.da1_op, D.A1, D.A1, S.cstar_op, H2, H2, D.att1_fc1.w, H2));             MT_TRY(bg(S.da1_op, D.A1, D.A1, D.att1_fc1.b));
  MT_TRY(wg(S.dlogit_op, H2, H2, S.a1_op, D.A1, D.A1, D.att1_fc2.w, D.A1));           MT_TRY(bg(S.dlogit_op, H2, H2, D.att1_fc2.b));
  MT_TRY(wg(S.da2_op, D.A2, D.A2, S.attd_op, H2, H2, D.att2_fc1.w, H2));              MT_TRY(bg(S.da2_op, D.A2, D.A2, D.att2_fc1.b));
  MT_TRY(wg(S.dzchat_op, MEM, MEM, S.a2_op, D.A2, D.A2, D.att2_fc2.w, D.A2));         MT_TRY(bg(S.dzchat_op, MEM, MEM, D.att2_fc2.b));
  const void* dgh2 = op_off(lp, (const void*)S.dgh_op, G);
  MT_TRY(wg(S.dgh_op, 2 * G, G, S.attd_op, H2, H2, D.g1_fc1.w, H2 + MEM));
  MT_TRY(wg(S.dgh_op, 2 * G, G, S.memprev_op, MEM, MEM, D.g1_fc1.w + H2, H2 + MEM));  MT_TRY(bg(S.dgh_op, 2 * G, G, D.g1_fc1.b));
  MT_TRY(wg(dgh2, 2 * G, G, S.attd_op, H2, H2, D.g2_fc1.w, H2 + MEM));
  MT_TRY(wg(dgh2, 2 * G, G, S.memprev_op, MEM, MEM, D.g2_fc1.w + H2, H2 + MEM));      MT_TRY(bg(dgh2, 2 * G, G, D.g2_fc1.b));
  const void* dzg2 = op_off(lp, (const void*)S.dzg_op, MEM);
  const void* gh2 = op_off(lp, (const void*)S.gh_op, G);
  MT_TRY(wg(S.dzg_op, 2 * MEM, MEM, S.gh_op, 2 * G, G, D.g1_fc2.w, G));               MT_TRY(bg(S.dzg_op, 2 * MEM, MEM, D.g1_fc2.b));
  MT_TRY(wg(dzg2, 2 * MEM, MEM, gh2, 2 * G, G, D.g2_fc2.w, G));                       MT_TRY(bg(dzg2, 2 * MEM, MEM, D.g2_fc2.b));
  MT_TRY(wg(S.dzoh_op, D.O, D.O, S.last_op, Hs + MEM, Hs + MEM, D.out_fc1.w, Hs + MEM));   MT_TRY(bg(S.dzoh_op, D.O, D.O, D.out_fc1.b));
  MT_TRY(mt_colsum_multi_run(lp, M, cjobs, n_cjobs, st));
  for (int m = 0; m < D.n_mods; ++m)            // b_hh enters the gates exactly like b_ih: same gradient
    MT_CUDA(cudaMemcpyAsync(grads + D.b_hh[m], grads + D.b_ih[m], sizeof(float) * 4 * D.H[m], cudaMemcpyDeviceToDevice, st));
  return MT_OK;
}

}  // extern "C"
