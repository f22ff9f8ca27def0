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
