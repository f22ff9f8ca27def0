// Gradient all-reduce of the data-parallel training path: ONE grouped NCCL all-reduce (SUM, fp32) over all flat
// gradient arenas of a model, enqueued on the caller's stream (so it can be captured into the step graph).
//
// NCCL is not linked: the functions are resolved at run time from the libnccl.so.2 that the host process has already loaded
// (PyTorch bundles it), which keeps libmt_b200.so loadable on a box without NCCL for the single-GPU path.  The communicator is
// created here from a unique id that the host broadcasts with whatever it already has (torch.distributed in training.py).
#include <dlfcn.h>

#include "mt_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                    // ncclSuccess == 0
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // already in the process (torch)?
    if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (n.h) {
      *(void**)&n.GetUniqueId = dlsym(n.h, "ncclGetUniqueId");
      *(void**)&n.CommInitRank = dlsym(n.h, "ncclCommInitRank");
      *(void**)&n.CommDestroy = dlsym(n.h, "ncclCommDestroy");
      *(void**)&n.AllReduce = dlsym(n.h, "ncclAllReduce");
      *(void**)&n.GroupStart = dlsym(n.h, "ncclGroupStart");
      *(void**)&n.GroupEnd = dlsym(n.h, "ncclGroupEnd");
      *(void**)&n.GetErrorString = dlsym(n.h, "ncclGetErrorString");
      n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GroupStart && n.GroupEnd;
    }
  }
  return n;
}

int nccl_fail(const Nccl& n, ncclResult_t r, const char* what) {
  snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "NCCL %s failed: %s", what, n.GetErrorString ? n.GetErrorString(r) : "error");
  return MT_ERR_CUDA;
}

// ---- overlapped all-reduce: the tail layers of the grouped encoder backward are reduced on a side stream while the head layers still run ----
struct Overlap {
  void* comm = nullptr;
  int split = -1;
  bool armed = false, fired = false;
  cudaStream_t cs = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float* ptr[MT_MAX_MODS];
  size_t cnt[MT_MAX_MODS];
  int n = 0;
};
Overlap g_ov;

}  // namespace

// called by mt_encoder_group_bwd once the gradients of layers >= split (and of the final norm) of all G stacks are enqueued on `st`:
// grads + g * pstride + [tail_off(split), total) are complete.  Returns the split layer through *split_out when armed (else -1).
int mt_comm_overlap_split(int n_layers) {
  if (!g_ov.armed || n_layers < 2) return -1;
  int s = g_ov.split;
  if (s < 0) s = n_layers / 3;
  if (s < 1) s = 1;
  if (s > n_layers - 1) s = n_layers - 1;
  return s;
}
int mt_comm_overlap_fire(float* grads, size_t pstride, int G, size_t tail_off, size_t total, cudaStream_t st) {
  Nccl& n = nccl();
  if (!g_ov.armed || !n.ok || G > MT_MAX_MODS || tail_off >= total) return MT_OK;
  g_ov.armed = false;
  MT_CUDA(cudaEventRecord(g_ov.ev_fork, st));
  MT_CUDA(cudaStreamWaitEvent(g_ov.cs, g_ov.ev_fork, 0));
  ncclResult_t r = n.GroupStart();
  if (r != 0) return nccl_fail(n, r, "ncclGroupStart");
  for (int g = 0; g < G; ++g) {
    g_ov.ptr[g] = grads + (size_t)g * pstride + tail_off;
    g_ov.cnt[g] = total - tail_off;
    r = n.AllReduce(g_ov.ptr[g], g_ov.ptr[g], g_ov.cnt[g], kNcclFloat32, kNcclSum, (ncclComm_t)g_ov.comm, g_ov.cs);
    if (r != 0) { n.GroupEnd(); return nccl_fail(n, r, "ncclAllReduce"); }
  }
  r = n.GroupEnd();
  if (r != 0) return nccl_fail(n, r, "ncclGroupEnd");
  g_ov.n = G;
  g_ov.fired = true;
  MT_CUDA(cudaEventRecord(g_ov.ev_join, g_ov.cs));
  return MT_OK;
}

extern "C" {

/* Arm the overlapped gradient all-reduce for the NEXT mt_encoder_group_bwd call of this process: as soon as that call has enqueued the
 * weight gradients of layers >= split_layer (split_layer < 0: n_layers / 3) and of the final norm, they are all-reduced (SUM, fp32) on the
 * library's communication stream while the remaining layers' backward still runs on the caller's stream. */
int mt_comm_overlap_arm(void* comm, int split_layer) {
  if (!nccl().ok) return MT_ERR_UNSUPPORTED;
  if (!comm) return MT_ERR_ARG;
  if (!g_ov.cs) {
    MT_CUDA(cudaStreamCreateWithFlags(&g_ov.cs, cudaStreamNonBlocking));
    MT_CUDA(cudaEventCreateWithFlags(&g_ov.ev_fork, cudaEventDisableTiming));
    MT_CUDA(cudaEventCreateWithFlags(&g_ov.ev_join, cudaEventDisableTiming));
  }
  g_ov.comm = comm; g_ov.split = split_layer; g_ov.armed = true;
  return MT_OK;
}
/* Make `stream` wait for the overlapped all-reduce (if one was started since the last join) and report which ranges it covered, so the
 * caller reduces only the rest (mt_allreduce_grads).  ptrs / counts: room for MT_MAX_MODS entries; *n_ranges = 0 when nothing was started. */
int mt_comm_overlap_join(void* stream, float** ptrs, size_t* counts, int* n_ranges) {
  if (!ptrs || !counts || !n_ranges) return MT_ERR_ARG;
  *n_ranges = 0;
  if (!g_ov.fired) return MT_OK;
  g_ov.fired = false;
  MT_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_ov.ev_join, 0));
  for (int g = 0; g < g_ov.n; ++g) { ptrs[g] = g_ov.ptr[g]; counts[g] = g_ov.cnt[g]; }
  *n_ranges = g_ov.n;
  return MT_OK;
}

int mt_comm_available(void) { return nccl().ok ? 1 : 0; }

/* rank 0: fill id[128]; the host broadcasts it to every rank */
int mt_comm_unique_id(char* id) {
  Nccl& n = nccl();
  if (!n.ok) return MT_ERR_UNSUPPORTED;
  if (!id) return MT_ERR_ARG;
  ncclUniqueId u;
  ncclResult_t r = n.GetUniqueId(&u);
  if (r != 0) return nccl_fail(n, r, "ncclGetUniqueId");
  memcpy(id, u.internal, 128);
  return MT_OK;
}

/* collective: every rank calls it with the same id; the current CUDA device is the rank's GPU */
int mt_comm_init(const char* id, int rank, int world, void** comm) {
  Nccl& n = nccl();
  if (!n.ok) return MT_ERR_UNSUPPORTED;
  if (!id || !comm || rank < 0 || rank >= world) return MT_ERR_ARG;
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = n.CommInitRank(&c, world, u, rank);
  if (r != 0) return nccl_fail(n, r, "ncclCommInitRank");
  *comm = c;
  return MT_OK;
}

int mt_comm_destroy(void* comm) {
  Nccl& n = nccl();
  if (!n.ok) return MT_ERR_UNSUPPORTED;
  if (!comm) return MT_OK;
  ncclResult_t r = n.CommDestroy((ncclComm_t)comm);
  return r == 0 ? MT_OK : nccl_fail(n, r, "ncclCommDestroy");
}

/* In-place SUM all-reduce of n fp32 device buffers (the flat gradient arenas; orphan parameters are never packed) as one NCCL
 * group on `stream`. */
int mt_allreduce_grads(void* comm, float* const* bufs, const size_t* counts, int n_bufs, void* stream) {
  Nccl& n = nccl();
  if (!n.ok) return MT_ERR_UNSUPPORTED;
  if (!comm || (n_bufs > 0 && (!bufs || !counts)) || n_bufs < 0) return MT_ERR_ARG;
  ncclResult_t r = n.GroupStart();
  if (r != 0) return nccl_fail(n, r, "ncclGroupStart");
  for (int i = 0; i < n_bufs; ++i) {
    if (!bufs[i] || counts[i] == 0) continue;
    r = n.AllReduce(bufs[i], bufs[i], counts[i], kNcclFloat32, kNcclSum, (ncclComm_t)comm, (cudaStream_t)stream);
    if (r != 0) { n.GroupEnd(); return nccl_fail(n, r, "ncclAllReduce"); }
  }
  r = n.GroupEnd();
  return r == 0 ? MT_OK : nccl_fail(n, r, "ncclGroupEnd");
}

}  // extern "C"
