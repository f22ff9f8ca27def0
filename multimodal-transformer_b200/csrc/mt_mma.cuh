// Warp-level tensor-core helpers (mma.sync.m16n8k16 bf16, ldmatrix) shared by the attention and recurrence kernels.
#pragma once
#include "mt_ops.cuh"

namespace mtmma {

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t* r, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// A-operand fragments of a [16 x DK] row-major global operand (rows r0 = row0 + lane/4 and r0 + 8), zero beyond `rows`
template <int DK>
__device__ __forceinline__ void load_a_frags(uint32_t (*a)[4], const bf16* base, int ld, int row0, int rows, int lane) {
  const int r0 = row0 + (lane >> 2), r1 = r0 + 8, c = 2 * (lane & 3);
#pragma unroll
  for (int ks = 0; ks < DK / 16; ++ks) {
    a[ks][0] = r0 < rows ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + ks * 16 + c) : 0u;
    a[ks][1] = r1 < rows ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + ks * 16 + c) : 0u;
    a[ks][2] = r0 < rows ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + ks * 16 + 8 + c) : 0u;
    a[ks][3] = r1 < rows ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + ks * 16 + 8 + c) : 0u;
  }
}


// A-operand fragment (16 rows x 16 k) of a row-major shared-memory tile: s points at (row0, k0), ld in elements
__device__ __forceinline__ void ldsm_a(uint32_t* a, const bf16* s, int ld, int lane) {
  ldsm4(a, s + (lane & 15) * ld + (lane >> 4) * 8);
}
// A-operand fragment of the TRANSPOSE of a row-major tile: result rows = source columns c0.., k = source rows r0..
// (s points at (r0, c0)); used for P^T . dO style contractions over the row index
__device__ __forceinline__ void ldsm_at(uint32_t* a, const bf16* s, int ld, int lane) {
  ldsm4t(a, s + ((lane & 7) + ((lane >> 4) & 1) * 8) * ld + ((lane >> 3) & 1) * 8);
}
// two B-operand fragments (k16 x n8 each, n = c0..c0+7 and c0+8..c0+15) of a row-major [k][n] tile: s points at (k0, c0)
__device__ __forceinline__ void ldsm_b2(uint32_t* b, const bf16* s, int ld, int lane) {
  ldsm4t(b, s + (((lane >> 3) & 1) * 8 + (lane & 7)) * ld + (lane >> 4) * 8);
}
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}

}  // namespace mtmma
