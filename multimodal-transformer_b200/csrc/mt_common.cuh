// Shared device/host helpers for libmt_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mt_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing ----------------------------------------------------------------------------------
extern char g_mt_cuda_err[512];
extern unsigned long long g_mt_launches;
int mt_set_cuda_error(cudaError_t e, const char* file, int line);
#define MT_CUDA(x)                                                        \
  do {                                                                    \
    cudaError_t _e = (x);                                                 \
    if (_e != cudaSuccess) return mt_set_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)
// per-launch profiling (mt_prof_* in include/mt_b200.h): when enabled, an event is recorded after every launch
extern int g_mt_prof_on;
void mt_prof_record(const char* func, int line, cudaStream_t st);
void mt_prof_work(double flops, double bytes);      // annotates the NEXT launch with its algorithmic work
void mt_prof_tag(const char* tag);                   // ... and with a shape tag appended to the site name
#define MT_LAUNCH_CHECK_S(stream_)                                        \
  do {                                                                    \
    ++g_mt_launches;                                                      \
    cudaError_t _e = cudaGetLastError();                                  \
    if (_e != cudaSuccess) return mt_set_cuda_error(_e, __FILE__, __LINE__); \
    if (g_mt_prof_on) mt_prof_record(__func__, __LINE__, (stream_));      \
  } while (0)
#define MT_LAUNCH_CHECK() MT_LAUNCH_CHECK_S(st)
#define MT_TRY(x)            \
  do {                       \
    int _r = (x);            \
    if (_r != MT_OK) return _r; \
  } while (0)

// tuning knobs (mt_tune): [0] GEMM grid share, [1] attention grid share, [2] LayerNorm grid share -- a share of s launches 1/s of the
// resident CTA slots so that kernels of concurrent streams (the three modality stacks) co-reside instead of queueing
extern int g_mt_tune[16];
// overlapped gradient all-reduce of the grouped encoder backward (mt_comm.cu)
int mt_comm_overlap_split(int n_layers);
int mt_comm_overlap_fire(float* grads, size_t pstride, int G, size_t tail_off, size_t total, cudaStream_t st);
#define MT_TUNE_GEMM_SHARE 0
#define MT_TUNE_ATTN_SHARE 1
#define MT_TUNE_LN_SHARE 2
#define MT_TUNE_PDL 3          // [3] programmatic dependent launch of the encoder-chain kernels (mt_launch_dep below): 0 off (default), 1 eager launches only, 2 always
#define MT_TUNE_NO_RS 5        // [5] != 0: the encoder's projections skip the row-stream engine (A/B against the streaming engine)
#define MT_TUNE_NO_BIG_TILES 7 // [7] != 0: no 256-row / 256-wide tiles for the L2-bound GEMMs (split-K wgrads, long-K dgrad)
#define MT_TUNE_REC_DEBUG 8    // [8] MFN recurrences: bit 5 clock trace of one CTA, bit 6 first-cut kernels (A/B), bits 0-3 timing experiments of the first cut
#define MT_TUNE_NO_DROPBITS 9  // [9] != 0: the tcgen05 attention kernels draw their dropout hashes themselves instead of reading precomputed keep bits
#define MT_TUNE_NO_MGROUPS 10  // [10] != 0: streaming GEMMs of the stacked modality rows launch per stack instead of once with row groups
#define MT_TUNE_LNX 11         // [11] != 0: the output projection fuses the sublayer-1 LayerNorm across a CTA pair (opt-in: measured no faster than the pass)
#define MT_TUNE_NO_LNDRAW 12   // [12] != 0: the attention keep bits come from the stand-alone draw kernel instead of riding in a LayerNorm forward pass
#define MT_TUNE_BF16_GSTREAM 13 // [13] != 0: bf16 mode carries the residual-stream gradient between the sublayers in bf16 (opt-in: 12 instead of 16 bytes per element, but measured SLOWER -- the LayerNorm backward is bound by its memory-instruction rate, and 8-byte accesses halve the bytes per instruction)
#define MT_TUNE_PDL_MASK 14    // [14] kernel families that may launch programmatically (MT_PDL_* bits below; default all)
#define MT_TUNE_PDL_DEBUG 15   // [15] bit 0: the attention forward does not trigger its dependents early (bisecting aid, see mt_pdl_enabled); bit 1: the attention backward keeps its light preparation launch (A/B: by default the output projection's input-gradient GEMM writes all four per-query scalars)
#define MT_TUNE_NO_LNFUSE 6    // [6] != 0: no LayerNorm fused into the FFN output projection's epilogue

// one-time-per-DEVICE guard for cudaFuncSetAttribute-style opt-ins (a process may drive several GPUs): true the first time the
// current device asks
struct MtPerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

static inline size_t mt_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch (mt_tune key MT_TUNE_PDL) -------------------------------------------
// A kernel launched through mt_launch_dep may become resident while the previous kernel of the stream is still draining (its grid
// launch latency, barrier init, TMEM allocation and descriptor prefetch overlap that tail).  Contract of every such kernel:
// mt_pdl_gate() sits in front of its FIRST global-memory access (read or write): it waits until the previous kernel has completed and
// flushed, then lets the next kernel of the stream be scheduled.  Gate after wait keeps the look-ahead at exactly one kernel.  Kernels
// launched with plain <<< >>> serialise fully on both sides, so the two kinds mix freely.
__device__ __forceinline__ void mt_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mt_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mt_pdl_gate() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// MT_TUNE_PDL: 0 = never (default), 1 = launches outside stream capture only, 2 = always; MT_TUNE_PDL_MASK selects the kernel families.
// OFF by default, for two measured reasons (B200, MFT-VAL train step, 184 kernels): (a) a captured graph gets slower with it (6.32 ->
// 6.38 ms: a graph's kernel-to-kernel edges already cost less than the programmatic hand-shake), only the eager step gains (7.27 ->
// 7.05 ms); (b) tools/fwd_determinism.py found the bf16 train-mode forward at B = 40 run-to-run NON-deterministic (|d pred| up to 3e-2)
// whenever the tcgen05 attention forward AND the output-projection GEMM behind it both launch programmatically -- each family alone, and
// every other combination tried, reproduces bit for bit.  Bisected further: launching only that GEMM shape plainly, or letting the
// attention forward NOT trigger early (mt_tune(15, 1): its dependents launch at its completion), restores determinism; a proxy fence
// behind the wait, a 20 us sleep behind the GEMM's gate and a 20 us sleep in front of the attention's exit do not.  Every kernel here
// executes the gate before its first global access, so the chain should be safe by the documented semantics of griddepcontrol.wait;
// until that pair is understood the switch is an experiment.
enum { MT_PDL_LN_FWD = 1, MT_PDL_LN_BWD = 2, MT_PDL_GEMM_RS = 4, MT_PDL_GEMM_TC = 8, MT_PDL_ATTN_FWD = 16, MT_PDL_ATTN_BWD = 32, MT_PDL_MISC = 64 };
static inline int mt_pdl_enabled(cudaStream_t st, int family) {
  if ((g_mt_tune[MT_TUNE_PDL_MASK] & family) == 0) return 0;
  const int mode = g_mt_tune[MT_TUNE_PDL];
  if (mode != 1) return mode ? 1 : 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return cs == cudaStreamCaptureStatusNone ? 1 : 0;
}
template <typename... KA, typename... A>
static inline cudaError_t mt_launch_dep(int family, void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = mt_pdl_enabled(st, family);
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

// ---- typed loads / stores ----------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }

// 4 consecutive elements <-> float4
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
// raw (unconverted) 4-element loads: keep prefetched registers free of dependent convert instructions
template <typename T> struct Raw4;
template <> struct Raw4<float> { typedef float4 type; };
template <> struct Raw4<bf16> { typedef uint2 type; };
__device__ __forceinline__ float4 ld4raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint2 ld4raw(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ float4 cvt4(float4 v) { return v; }
__device__ __forceinline__ float4 cvt4(uint2 u) {
  float2 fa = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x)), fb = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---- warp reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- counter-based dropout generator (mirrored by oracle/dropout_rng.py) -----------------------------
// One 32-bit draw decides TWO adjacent elements (16 bits each, threshold = (p * 2^32) >> 16): the per-element hash was the
// largest single instruction cost of every epilogue that applies dropout.
//   key            = mix32(seed_lo ^ 0x9E3779B9 * (site + 1)) + seed_hi * 0x85EBCA6B          (once per thread)
//   draw32(pair)   = mix32(pair_lo ^ key ^ pair_hi * 0xC2B2AE35)
//   element e      -> pair e >> 1, low half when e is even;   keep = half >= t16
__host__ __device__ __forceinline__ uint32_t mt_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
static inline uint32_t mt_drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t <= 0) return 0u;
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)t;
}
struct DropCfg {
  uint32_t thresh;   // 0 => dropout disabled; the kernels compare 16-bit halves against thresh >> 16
  float scale;       // 1/(1-p)
  uint64_t seed;
  uint32_t site;
  const unsigned long long* seed_off;   // optional device word added to `seed` (lets a captured CUDA graph draw fresh masks per replay)
  uint32_t key;      // filled by mt_drop_resolve
};
extern const unsigned long long* g_mt_seed_offset_ptr;   // set by mt_set_seed_offset_ptr (mt_api.cu)
static inline DropCfg mt_make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  d.thresh = (p > 0.f) ? mt_drop_threshold(p) : 0u;
  d.scale = (p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
  d.seed = seed;
  d.site = site;
  d.seed_off = (p > 0.f) ? g_mt_seed_offset_ptr : nullptr;
  d.key = 0u;
  return d;
}
// fold the optional device-side seed offset into the seed and derive the site key ONCE per thread, at kernel entry
__device__ __forceinline__ DropCfg mt_drop_resolve(DropCfg d) {
  if (d.thresh != 0u) {
    if (d.seed_off) d.seed += (uint64_t)__ldg(d.seed_off);
    d.key = mt_mix32((uint32_t)d.seed ^ (0x9E3779B9u * (d.site + 1u))) + (uint32_t)(d.seed >> 32) * 0x85EBCA6Bu;
  }
  d.seed_off = nullptr;
  return d;
}
__device__ __forceinline__ uint32_t mt_draw32(const DropCfg& d, uint64_t pair) {
  return mt_mix32((uint32_t)pair ^ d.key ^ ((uint32_t)(pair >> 32) * 0xC2B2AE35u));
}
__device__ __forceinline__ float mt_drop_factor(const DropCfg& d, uint64_t idx) {
  if (d.thresh == 0u) return 1.0f;
  const uint32_t bits = mt_draw32(d, idx >> 1);
  const uint32_t v = (idx & 1ull) ? (bits >> 16) : (bits & 0xFFFFu);
  return v >= (d.thresh >> 16) ? d.scale : 0.0f;
}
// elements idx (even) and idx + 1
__device__ __forceinline__ void mt_drop_pair(const DropCfg& d, uint64_t idx, float& f0, float& f1) {
  if (d.thresh == 0u) { f0 = f1 = 1.0f; return; }
  const uint32_t bits = mt_draw32(d, idx >> 1);
  const uint32_t t16 = d.thresh >> 16;
  f0 = (bits & 0xFFFFu) >= t16 ? d.scale : 0.0f;
  f1 = (bits >> 16) >= t16 ? d.scale : 0.0f;
}
// elements idx .. idx + 3 (idx a multiple of 2)
__device__ __forceinline__ void mt_drop_quad(const DropCfg& d, uint64_t idx, float* f) {
  mt_drop_pair(d, idx, f[0], f[1]);
  mt_drop_pair(d, idx + 2, f[2], f[3]);
}

// Attention-probability dropout (site MT_SITE_ATTN_P): pair index = row * ceil(T/2) + (j >> 1) so a pair never straddles
// two query rows.  row = flat (batch, head, query) index, j = key index.   Mirrored by oracle/dropout_rng.py:attn_keep_mask.
__device__ __forceinline__ uint32_t mt_attn_drop_bits(const DropCfg& d, uint64_t row, uint32_t P2, uint32_t j) {
  return mt_draw32(d, row * (uint64_t)P2 + (uint64_t)(j >> 1));
}
__device__ __forceinline__ float mt_attn_drop_factor(const DropCfg& d, uint64_t row, uint32_t P2, uint32_t j) {
  if (d.thresh == 0u) return 1.0f;
  const uint32_t bits = mt_attn_drop_bits(d, row, P2, j);
  const uint32_t v = (j & 1u) ? (bits >> 16) : (bits & 0xFFFFu);
  return v >= (d.thresh >> 16) ? d.scale : 0.0f;
}
// both keys of the pair (j even): f[0] for key j, f[1] for key j + 1
__device__ __forceinline__ void mt_attn_drop_pair(const DropCfg& d, uint64_t row, uint32_t P2, uint32_t j, float& f0, float& f1) {
  if (d.thresh == 0u) { f0 = f1 = 1.0f; return; }
  const uint32_t bits = mt_attn_drop_bits(d, row, P2, j);
  const uint32_t t16 = d.thresh >> 16;
  f0 = (bits & 0xFFFFu) >= t16 ? d.scale : 0.0f;
  f1 = (bits >> 16) >= t16 ? d.scale : 0.0f;
}

// dropout site ids (oracle/dropout_rng.py)
#define MT_SITE_ATTN_P 0
#define MT_SITE_SUB0 1
#define MT_SITE_FFN_H 2
#define MT_SITE_SUB1 3
#define MT_SITE_MFN_G1 0x4000u
#define MT_SITE_MFN_G2 0x4001u
#define MT_SITE_MFN_OUT 0x4002u
static inline uint32_t mt_enc_site(int stack, int layer, int k) { return (uint32_t)((stack * 64 + layer) * 8 + k); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
