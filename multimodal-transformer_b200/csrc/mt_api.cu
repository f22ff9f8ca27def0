// Library-level entry points of libmt_b200: error text, device probe, launch counter.
#include <stdlib.h>

#include "mt_common.cuh"

char g_mt_cuda_err[512] = "";
unsigned long long g_mt_launches = 0ull;
const unsigned long long* g_mt_seed_offset_ptr = nullptr;

namespace {
__global__ void spin_kernel(long long ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); } while ((long long)(t1 - t0) < ns);
}
}  // namespace

int mt_set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
  return MT_ERR_CUDA;
}

// ---- per-launch profiler ---------------------------------------------------------------------------------
int g_mt_prof_on = 0;
namespace {
struct ProfRec { char name[96]; double flops, bytes; cudaEvent_t ev; };
ProfRec* g_recs = nullptr;
int g_prof_cap = 0, g_prof_n = 0;
cudaEvent_t g_prof_base = nullptr;
double g_pending_flops = 0.0, g_pending_bytes = 0.0;
char g_pending_tag[40] = "";
}  // namespace

void mt_prof_tag(const char* tag) {
  if (g_mt_prof_on) snprintf(g_pending_tag, sizeof(g_pending_tag), "%s", tag);
}

void mt_prof_work(double flops, double bytes) {
  if (!g_mt_prof_on) return;
  g_pending_flops = flops;
  g_pending_bytes = bytes;
}

void mt_prof_record(const char* func, int line, cudaStream_t st) {
  if (g_prof_n >= g_prof_cap) { g_pending_flops = g_pending_bytes = 0.0; return; }
  ProfRec& r = g_recs[g_prof_n];
  snprintf(r.name, sizeof(r.name), g_pending_tag[0] ? "%s:%d %s" : "%s:%d", func, line, g_pending_tag);
  g_pending_tag[0] = 0;
  r.flops = g_pending_flops; r.bytes = g_pending_bytes;
  g_pending_flops = g_pending_bytes = 0.0;
  if (cudaEventRecord(r.ev, st) == cudaSuccess) ++g_prof_n;
}

extern "C" {

int mt_prof_start(int max_records, void* stream) {
  if (max_records <= 0) return MT_ERR_ARG;
  if (max_records > g_prof_cap) {
    ProfRec* n = (ProfRec*)realloc(g_recs, sizeof(ProfRec) * (size_t)max_records);
    if (!n) return MT_ERR_ARG;
    g_recs = n;
    for (int i = g_prof_cap; i < max_records; ++i) MT_CUDA(cudaEventCreate(&g_recs[i].ev));
    g_prof_cap = max_records;
  }
  if (!g_prof_base) MT_CUDA(cudaEventCreate(&g_prof_base));
  g_prof_n = 0;
  MT_CUDA(cudaEventRecord(g_prof_base, (cudaStream_t)stream));
  g_mt_prof_on = 1;
  return MT_OK;
}

int mt_prof_stop(void) {
  g_mt_prof_on = 0;
  return g_prof_n;
}

int mt_prof_get(int i, char* name, int name_cap, float* ms, double* flops, double* bytes) {
  if (i < 0 || i >= g_prof_n || !name || name_cap <= 0 || !ms) return MT_ERR_ARG;
  MT_CUDA(cudaEventSynchronize(g_recs[i].ev));
  MT_CUDA(cudaEventElapsedTime(ms, i == 0 ? g_prof_base : g_recs[i - 1].ev, g_recs[i].ev));
  snprintf(name, (size_t)name_cap, "%s", g_recs[i].name);
  if (flops) *flops = g_recs[i].flops;
  if (bytes) *bytes = g_recs[i].bytes;
  return MT_OK;
}

int mt_set_seed_offset_ptr(const uint64_t* dev_ptr) {
  g_mt_seed_offset_ptr = reinterpret_cast<const unsigned long long*>(dev_ptr);
  return MT_OK;
}

int mt_spin(float ms, void* stream) {
  if (ms < 0.f || ms > 2000.f) return MT_ERR_ARG;
  spin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long)(ms * 1e6f));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mt_set_cuda_error(e, __FILE__, __LINE__);
  return MT_OK;
}

const char* mt_error_string(int code) {
  switch (code) {
    case MT_OK: return "ok";
    case MT_ERR_ARG: return "bad argument (shape / null pointer / unsupported size)";
    case MT_ERR_ALIGN: return "pointer or leading dimension not aligned as required";
    case MT_ERR_CUDA: return "CUDA runtime error (see mt_last_cuda_error)";
    case MT_ERR_WS: return "workspace too small";
    case MT_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error code";
  }
}

const char* mt_last_cuda_error(void) { return g_mt_cuda_err; }

int mt_version(void) { return 100; }

/* tuning knobs (see include/mt_b200.h: mt_tune) */
int g_mt_tune[16] = {1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0};
int mt_tune(int key, int value) {
  if (key < 0 || key >= 16) return -1;
  int old = g_mt_tune[key];
  g_mt_tune[key] = value;
  return old;
}

uint64_t mt_launch_count(void) { return (uint64_t)g_mt_launches; }

int mt_check_device(int dev) {
  cudaDeviceProp p;
  MT_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "device %d is sm_%d%d; libmt_b200 is built for sm_100a only", dev, p.major, p.minor);
    return MT_ERR_UNSUPPORTED;
  }
  return MT_OK;
}

}  // extern "C"
