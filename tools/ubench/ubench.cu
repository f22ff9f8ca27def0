// Instruction-level micro-benchmarks for the recurrence kernels: mma.sync m16n8k16 bf16 and MUFU latency / throughput per SM
// sub-partition on sm_100a.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }

// CHAINS independent accumulators per warp, N mma per chain
template <int CHAINS>
__global__ void k_mma(long long* out, int iters, float* sink) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  float c[CHAINS][4] = {};
  __syncthreads();
  const long long t0 = clk();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma16816(c[j], a, 0x3f803f80u, 0x3f803f80u);
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  asm volatile("" ::"f"(s));
  const long long t1 = clk();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 12345.f) *sink = s;
}

template <int OP, int CHAINS>
__global__ void k_mufu(long long* out, int iters, float* sink) {
  float x[CHAINS];
  for (int j = 0; j < CHAINS; ++j) x[j] = 0.001f * (threadIdx.x + j + 1);
  __syncthreads();
  const long long t0 = clk();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
    }
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) s += x[j];
  asm volatile("" ::"f"(s));
  const long long t1 = clk();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 12345.f) *sink = s;
}

__global__ void k_shfl(long long* out, int iters, float* sink) {
  float x = threadIdx.x;
  __syncthreads();
  const long long t0 = clk();
  for (int i = 0; i < iters; ++i) x = __shfl_xor_sync(0xffffffffu, x, 16) + 1.f;
  asm volatile("" ::"f"(x));
  const long long t1 = clk();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (x == 12345.f) *sink = x;
}

template <typename K>
static void run(const char* name, K kern, int threads, int iters, int per_iter) {
  long long* d; float* sink;
  cudaMalloc(&d, 8 * 8); cudaMalloc(&sink, 4);
  kern<<<1, threads>>>(d, iters, sink);
  kern<<<1, threads>>>(d, iters, sink);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const int warps = threads / 32;
  printf("%-34s warps %2d (per SMSP %.1f): %7.1f clk per op per warp, %6.2f clk per warp-instruction per SMSP\n", name, warps, warps / 4.0,
         (double)h / (iters * (double)per_iter), (double)h / (iters * (double)per_iter) / (warps / 4.0 > 1 ? warps / 4.0 : 1));
  cudaFree(d); cudaFree(sink);
}

int main() {
  const int it = 2000;
  for (int threads : {32, 128, 256, 512}) {
    run("mma16816 1 chain", k_mma<1>, threads, it, 1);
    run("mma16816 3 chains", k_mma<3>, threads, it, 3);
    run("mma16816 6 chains", k_mma<6>, threads, it, 6);
  }
  for (int threads : {32, 256, 512}) {
    run("ex2 1 chain", k_mufu<0, 1>, threads, it, 1);
    run("ex2 8 chains", k_mufu<0, 8>, threads, it, 8);
    run("rcp 8 chains", k_mufu<1, 8>, threads, it, 8);
    run("tanh 1 chain", k_mufu<2, 1>, threads, it, 1);
    run("tanh 8 chains", k_mufu<2, 8>, threads, it, 8);
  }
  run("shfl+fadd chain", k_shfl, 32, it, 1);
  run("shfl+fadd chain", k_shfl, 256, it, 1);
  return 0;
}
