// Batched evaluation metrics: per-narrative CCC / Pearson r / squared error of a padded batch in one launch.
// eval_ccc MFT/train.py:42-50 (population variances, np.cov(bias=True)); the reference evaluates one narrative per forward
// (batch_size=1, MFT/train.py:169,218) and computes these on the host.
#include "mt_common.cuh"

namespace {

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();                 // sh may still be read from the previous reduction
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

// one CTA per narrative; two passes (means, then central moments) in fp64
__global__ void ccc_kernel(const float* __restrict__ pred, const float* __restrict__ target, const int* __restrict__ lengths, int T,
                           double* __restrict__ ccc, double* __restrict__ pearson, double* __restrict__ sq_err) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  int n = lengths[b];
  n = n < 0 ? 0 : (n > T ? T : n);
  const float* p = pred + (size_t)b * T;
  const float* t = target + (size_t)b * T;
  double sp = 0.0, st = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { sp += (double)p[i]; st += (double)t[i]; }
  sp = block_sum(sp, sh);
  st = block_sum(st, sh);
  const double inv = n > 0 ? 1.0 / (double)n : 0.0;
  const double pm = sp * inv, tm = st * inv;
  double vp = 0.0, vt = 0.0, cv = 0.0, se = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dp = (double)p[i] - pm, dt = (double)t[i] - tm, e = (double)p[i] - (double)t[i];
    vp += dp * dp; vt += dt * dt; cv += dp * dt; se += e * e;
  }
  vp = block_sum(vp, sh) * inv;
  vt = block_sum(vt, sh) * inv;
  cv = block_sum(cv, sh) * inv;
  se = block_sum(se, sh);
  if (threadIdx.x == 0) {
    ccc[b] = 2.0 * cv / (vt + vp + (pm - tm) * (pm - tm));
    if (pearson) pearson[b] = cv / sqrt(vp * vt);
    if (sq_err) atomicAdd(sq_err, se);
  }
}

}  // namespace

extern "C" int mt_ccc_batched(const float* pred, const float* target, const int* lengths, int B, int T, double* ccc, double* pearson,
                              double* sq_err, void* stream) {
  if (!pred || !target || !lengths || !ccc || B <= 0 || T <= 0) return MT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (sq_err) MT_CUDA(cudaMemsetAsync(sq_err, 0, sizeof(double), st));
  ccc_kernel<<<B, 256, 0, st>>>(pred, target, lengths, T, ccc, pearson, sq_err);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// GPU-side batcher: the reference assembles every batch with torch.tensor(nested python lists) and a sort by
// length (generateTrainBatch / generateInputChunkHelper, MFT/train.py:59-108).  Here the padded corpus lives on the device once and
// a batch is an index gather: dst[b, :prefix] = src[idx[b], :prefix] (the first T_batch windows of narrative idx[b]).
// ---------------------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void gather_prefix_kernel(const float* __restrict__ src, size_t row_stride, const int* __restrict__ idx, int B, size_t prefix,
                                     float* __restrict__ dst, int vec) {
  // grid.y = narrative of the batch; grid.x strides over its prefix
  const int b = blockIdx.y;
  const float* s = src + (size_t)idx[b] * row_stride;
  float* d = dst + (size_t)b * prefix;
  if (vec) {
    const size_t n4 = prefix / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
      reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < prefix; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
  }
}

// mask[b, t] = t < lengths[b]   (MFT/train.py:103-106)
__global__ void length_mask_kernel(const int* __restrict__ lengths, int B, int T, float* __restrict__ mask) {
  const size_t n = (size_t)B * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    mask[i] = (int)(i % T) < lengths[i / T] ? 1.f : 0.f;
}

}  // namespace

extern "C" int mt_batch_gather(const float* src, size_t row_stride, const int* idx, int B, size_t prefix, float* dst, void* stream) {
  if (!src || !idx || !dst || B <= 0 || prefix == 0 || prefix > row_stride) return MT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (prefix % 4 == 0 && row_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) ? 1 : 0;
  const size_t work = vec ? prefix / 4 : prefix;
  size_t gx = (work + 255) / 256;
  const size_t cap = (size_t)(148 * 8 + B - 1) / B;          // ~8 CTAs per SM over the whole batch
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  mt_prof_work(0.0, 8.0 * (double)B * (double)prefix);
  gather_prefix_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(src, row_stride, idx, B, prefix, dst, vec);
  MT_LAUNCH_CHECK();
  return MT_OK;
}

extern "C" int mt_length_mask(const int* lengths, int B, int T, float* mask, void* stream) {
  if (!lengths || !mask || B <= 0 || T <= 0) return MT_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  size_t g = ((size_t)B * T + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  length_mask_kernel<<<(unsigned)g, 256, 0, st>>>(lengths, B, T, mask);
  MT_LAUNCH_CHECK();
  return MT_OK;
}
