// FFMA (SIMT) GEMM: the fp32 parity-mode engine, and the engine for shapes outside the tcgen05 kernel's
// envelope (tiny N, unaligned leading dimensions).  64x64x16 tiles, 256 threads, 4x4 outputs per thread,
// operands staged through shared memory as [k][row] so the inner product reads are conflict-free float4s.
#include "mt_gemm.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

struct SimtArgs {
  int M, N, K;
  const void* A; int lda;
  const void* B; int ldb;
  void* C; int ldc;
  int k_chunk;     // K range per blockIdx.z
  int split_k;
  GemmEpi epi;
};

// load a [rows x BK] operand tile into smem[k][row]; operand element (r,k) lives at
// KMAJ ? p[r*ld + k] : p[k*ld + r].  VEC = 4 requires 4-element alignment of the contiguous direction.
template <typename T, bool KMAJ, int VEC>
__device__ __forceinline__ void load_tile(float (*s)[BM + PAD], const T* __restrict__ p, int ld, int row0, int rows,
                                          int k0, int kend, int tid) {
  if (KMAJ) {
    // thread -> (row = tid/4, 4 consecutive k starting at (tid%4)*4)
    int r = tid >> 2, kk = (tid & 3) * 4;
    int gr = row0 + r, gk = k0 + kk;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < rows) {
      const T* q = p + (size_t)gr * ld + gk;
      if (VEC == 4) {
        if (gk < kend) { float4 f = ld4(q); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (gk + j < kend) v[j] = to_f(q[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s[kk + j][r] = v[j];
  } else {
    // thread -> (k = tid/16, 4 consecutive rows starting at (tid%16)*4)
    int kk = tid >> 4, r = (tid & 15) * 4;
    int gk = k0 + kk, gr = row0 + r;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gk < kend) {
      const T* q = p + (size_t)gk * ld + gr;
      if (VEC == 4) {
        if (gr < rows) { float4 f = ld4(q); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (gr + j < rows) v[j] = to_f(q[j]);
      }
    }
    *reinterpret_cast<float4*>(&s[kk][r]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <typename TI, typename TC, bool AK, bool BKM, int VEC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtArgs g) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  const TI* A = reinterpret_cast<const TI*>(g.A);
  const TI* B = reinterpret_cast<const TI*>(g.B);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    load_tile<TI, AK, VEC>(As, A, g.lda, m0, g.M, k0, kend, tid);
    load_tile<TI, BKM, VEC>(Bs, B, g.ldb, n0, g.N, k0, kend, tid);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const GemmEpi& e = g.epi;
  const DropCfg edrop = mt_drop_resolve(e.drop);
  TC* C = reinterpret_cast<TC*>(g.C);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    float rm = e.rowmask ? e.rowmask[m] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * e.alpha;
      if (g.split_k > 1) {
        atomicAdd(reinterpret_cast<float*>(g.C) + (size_t)m * g.ldc + n, v);
        continue;
      }
      if (e.bias) v += e.bias[n];
      if (e.act == MT_ACT_RELU) v = fmaxf(v, 0.f);
      else if (e.act == MT_ACT_TANH) v = tanhf(v);
      v *= mt_drop_factor(edrop, (uint64_t)m * (uint64_t)g.N + (uint64_t)n);
      if (e.gate) {
        float gt = to_f(reinterpret_cast<const TI*>(e.gate)[(size_t)m * e.ldg + n]);
        v = gt > 0.f ? v * e.gate_scale : 0.f;
      }
      if (e.residual) v += e.residual[(size_t)m * e.ldr + n];
      v *= rm;
      size_t o = (size_t)m * g.ldc + n;
      if (e.accumulate) v += to_f(C[o]);
      C[o] = from_f<TC>(v);
    }
  }
}

template <typename TI, typename TC>
int launch(const GemmDesc& d, cudaStream_t st) {
  SimtArgs a;
  a.M = d.M; a.N = d.N; a.K = d.K;
  a.A = d.A; a.lda = d.lda; a.B = d.B; a.ldb = d.ldb; a.C = d.C; a.ldc = d.ldc;
  a.epi = d.epi;
  int split = d.split_k < 1 ? 1 : d.split_k;
  int kc = (d.K + split - 1) / split;
  kc = (kc + BK - 1) / BK * BK;
  split = (d.K + kc - 1) / kc;
  a.k_chunk = kc;
  a.split_k = d.split_k > 1 ? 2 : 1;     // any value > 1 selects the atomic epilogue, even if only one chunk remains
  dim3 grid((d.N + BN - 1) / BN, (d.M + BM - 1) / BM, split);
  // vector path needs 4-element alignment of base pointers and leading dims (both operands)
  size_t es = sizeof(TI);
  bool vec = ((uintptr_t)d.A % (4 * es) == 0) && ((uintptr_t)d.B % (4 * es) == 0) && (d.lda % 4 == 0) && (d.ldb % 4 == 0);
  // K-major tiles vectorise along k: every 4-group must be fully inside [kbeg,kend) -> K % 4 == 0
  if (d.a_kmajor || d.b_kmajor) vec = vec && (d.K % 4 == 0);
  // MN-major tiles vectorise along rows: rows % 4 == 0
  if (!d.a_kmajor) vec = vec && (d.M % 4 == 0);
  if (!d.b_kmajor) vec = vec && (d.N % 4 == 0);
#define MT_SIMT_GO(AK, BKM)                                                                  \
  do {                                                                                       \
    if (vec) gemm_simt_kernel<TI, TC, AK, BKM, 4><<<grid, 256, 0, st>>>(a);                  \
    else gemm_simt_kernel<TI, TC, AK, BKM, 1><<<grid, 256, 0, st>>>(a);                      \
  } while (0)
  if (d.a_kmajor && d.b_kmajor) MT_SIMT_GO(true, true);
  else if (d.a_kmajor && !d.b_kmajor) MT_SIMT_GO(true, false);
  else if (!d.a_kmajor && d.b_kmajor) MT_SIMT_GO(false, true);
  else MT_SIMT_GO(false, false);
#undef MT_SIMT_GO
  MT_LAUNCH_CHECK();
  return MT_OK;
}

__global__ void colsum_kernel_f32(int M, int N, const float* __restrict__ X, int ldx, float* __restrict__ out, int rows_per_block) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += X[(size_t)m * ldx + n];
  atomicAdd(out + n, s);
}
__global__ void colsum_kernel_bf16(int M, int N, const bf16* __restrict__ X, int ldx, float* __restrict__ out, int rows_per_block) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += __bfloat162float(X[(size_t)m * ldx + n]);
  atomicAdd(out + n, s);
}

// vectorised column sums: 32 column-quads x 8 row lanes per CTA, rows strided over the grid's y dimension
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(int M, int N, const T* __restrict__ X, int ldx, float* __restrict__ out,
                                                          int rows_per_block) {
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + tx * 4;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    int m = m0 + ty;
    for (; m + 24 < m1; m += 32) {            // 4 independent 16-byte loads in flight
      float4 a = ld4(X + (size_t)m * ldx + n), b = ld4(X + (size_t)(m + 8) * ldx + n);
      float4 c = ld4(X + (size_t)(m + 16) * ldx + n), d = ld4(X + (size_t)(m + 24) * ldx + n);
      s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; m < m1; m += 8) {
      float4 a = ld4(X + (size_t)m * ldx + n);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int q = 1; q < 8; ++q) { float4 o = red[q][tx]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    atomicAdd(out + n, s.x); atomicAdd(out + n + 1, s.y); atomicAdd(out + n + 2, s.z); atomicAdd(out + n + 3, s.w);
  }
}

// several column-sum jobs of the same row count in ONE launch (blockIdx.z = job): the MFN backward has a dozen small bias
// gradients whose individual launches cost more than their memory traffic
struct ColsumJobs { ColsumJob j[MT_COLSUM_MAX_JOBS]; };
template <typename T>
__global__ void __launch_bounds__(256) colsum_multi_kernel(int M, ColsumJobs jobs, int rows_per_block) {
  __shared__ float4 red[8][32];
  const ColsumJob jb = jobs.j[blockIdx.z];
  const T* X = reinterpret_cast<const T*>(jb.X);
  const int N = jb.N, ldx = jb.ldx;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + tx * 4;
  if (blockIdx.x * 128 >= N) return;                      // CTA-uniform
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    int m = m0 + ty;
    for (; m + 56 < m1; m += 64) {      // eight 8 / 16-byte loads in flight per thread: the kernel is latency-bound below that
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ld4(X + (size_t)(m + 8 * k) * ldx + n);
      s.x += ((v[0].x + v[1].x) + (v[2].x + v[3].x)) + ((v[4].x + v[5].x) + (v[6].x + v[7].x));
      s.y += ((v[0].y + v[1].y) + (v[2].y + v[3].y)) + ((v[4].y + v[5].y) + (v[6].y + v[7].y));
      s.z += ((v[0].z + v[1].z) + (v[2].z + v[3].z)) + ((v[4].z + v[5].z) + (v[6].z + v[7].z));
      s.w += ((v[0].w + v[1].w) + (v[2].w + v[3].w)) + ((v[4].w + v[5].w) + (v[6].w + v[7].w));
    }
    for (; m + 24 < m1; m += 32) {
      float4 a = ld4(X + (size_t)m * ldx + n), b = ld4(X + (size_t)(m + 8) * ldx + n);
      float4 c = ld4(X + (size_t)(m + 16) * ldx + n), d = ld4(X + (size_t)(m + 24) * ldx + n);
      s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; m < m1; m += 8) {
      float4 a = ld4(X + (size_t)m * ldx + n);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int q = 1; q < 8; ++q) { float4 o = red[q][tx]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    atomicAdd(jb.out + n, s.x); atomicAdd(jb.out + n + 1, s.y); atomicAdd(jb.out + n + 2, s.z); atomicAdd(jb.out + n + 3, s.w);
  }
}

}  // namespace

// out_j[n] += sum_m X_j[m * ldx_j + n] for up to MT_COLSUM_MAX_JOBS tensors with M rows each (N_j % 4 == 0, 16-byte aligned rows)
int mt_colsum_multi_run(int x_is_bf16, int M, const ColsumJob* jobs, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return MT_OK;
  if (n_jobs > MT_COLSUM_MAX_JOBS) return MT_ERR_ARG;
  const size_t es = x_is_bf16 ? 2 : 4;
  ColsumJobs J;
  int maxN = 0, col_blocks = 0;
  double bytes = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const ColsumJob& b = jobs[i];
    if (b.N % 4 != 0 || b.ldx % 4 != 0 || ((uintptr_t)b.X % (4 * es)) != 0) {       // odd shape: one job at a time
      for (int k = 0; k < n_jobs; ++k) MT_TRY(mt_colsum_run(x_is_bf16, M, jobs[k].N, jobs[k].X, jobs[k].ldx, jobs[k].out, 1, st));
      return MT_OK;
    }
    J.j[i] = b;
    maxN = b.N > maxN ? b.N : maxN;
    col_blocks += (b.N + 127) / 128;
    bytes += (double)M * b.N * es;
  }
  const int gx = (maxN + 127) / 128;
  // CTAs beyond a job's own width exit at once: size the row split by the column blocks that do work (8 CTAs of 256 threads per SM)
  int gy = (148 * 8 + col_blocks - 1) / col_blocks;
  if (gy < 1) gy = 1;
  int rpb = ((M + gy - 1) / gy + 63) / 64 * 64;
  if (rpb < 64) rpb = 64;
  dim3 grid(gx, (M + rpb - 1) / rpb, n_jobs);
  mt_prof_work(0.0, bytes);
  if (x_is_bf16) colsum_multi_kernel<bf16><<<grid, 256, 0, st>>>(M, J, rpb);
  else colsum_multi_kernel<float><<<grid, 256, 0, st>>>(M, J, rpb);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

int mt_gemm_simt_run(int dtype, const GemmDesc& d, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || !d.A || !d.B || !d.C) return MT_ERR_ARG;
  if (d.split_k > 1 && !(d.c_f32 || dtype == MT_F32)) return MT_ERR_ARG;
  if (dtype == MT_F32) return launch<float, float>(d, st);
  if (d.c_f32) return launch<bf16, float>(d, st);
  return launch<bf16, bf16>(d, st);
}

int mt_colsum_run(int x_is_bf16, int M, int N, const void* X, int ldx, float* out, int accumulate, cudaStream_t st) {
  if (!accumulate) MT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st));
  mt_prof_work(0.0, (double)M * N * (x_is_bf16 ? 2.0 : 4.0));
  const size_t es = x_is_bf16 ? 2 : 4;
  if (N % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)X % (4 * es)) == 0) {
    const int gx = (N + 127) / 128;
    int gy = (148 * 4 + gx - 1) / gx;
    int rpbv = ((M + gy - 1) / gy + 31) / 32 * 32;
    if (rpbv < 64) rpbv = 64;
    dim3 gridv(gx, (M + rpbv - 1) / rpbv);
    if (x_is_bf16) colsum_vec_kernel<bf16><<<gridv, 256, 0, st>>>(M, N, (const bf16*)X, ldx, out, rpbv);
    else colsum_vec_kernel<float><<<gridv, 256, 0, st>>>(M, N, (const float*)X, ldx, out, rpbv);
    MT_LAUNCH_CHECK();
    return MT_OK;
  }
  int rpb = 256;
  // enough row blocks to fill the machine without drowning in atomics
  while ((size_t)((M + rpb - 1) / rpb) * ((N + 127) / 128) > 148 * 8 && rpb < 4096) rpb *= 2;
  dim3 grid((N + 127) / 128, (M + rpb - 1) / rpb);
  if (x_is_bf16) colsum_kernel_bf16<<<grid, 128, 0, st>>>(M, N, (const bf16*)X, ldx, out, rpb);
  else colsum_kernel_f32<<<grid, 128, 0, st>>>(M, N, (const float*)X, ldx, out, rpb);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
