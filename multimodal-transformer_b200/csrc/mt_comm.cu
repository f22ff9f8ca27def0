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

}  // namespace

extern "C" {

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
