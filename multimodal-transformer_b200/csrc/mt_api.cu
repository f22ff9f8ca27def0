// Library-level entry points of libmt_b200: error text, device probe, launch counter.
#include "mt_common.cuh"

char g_mt_cuda_err[512] = "";
unsigned long long g_mt_launches = 0ull;

int mt_set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_mt_cuda_err, sizeof(g_mt_cuda_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
  return MT_ERR_CUDA;
}

extern "C" {

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
