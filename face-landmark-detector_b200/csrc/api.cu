// api.cu — handle life-cycle and error reporting for libfld_sm100.so (the library owns no scratch memory: every
// buffer, including decode partials, is caller-provided so that calls on different streams never share state)
#include <stdarg.h>
#include "common.cuh"

std::atomic<uint64_t> g_fld_launches{0};
static thread_local char g_err[1024] = "";

void fld_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int fld_abi_version(void) { return FLD_ABI_VERSION; }
extern "C" const char* fld_last_error(void) { return g_err; }
extern "C" uint64_t fld_launch_count(void) { return g_fld_launches.load(); }

extern "C" int fld_create(int device, fld_handle** out) {
  if (!out) { fld_set_error("fld_create: null out"); return FLD_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fld_set_error("fld_create: no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    return FLD_ERR_NODEVICE;
  }
  if (device < 0 || device >= n) { fld_set_error("fld_create: device %d out of range [0,%d)", device, n); return FLD_ERR_INVALID; }
  cudaDeviceProp prop;
  FLD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    fld_set_error("fld_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return FLD_ERR_NODEVICE;
  }
  FLD_CUDA(cudaSetDevice(device));
  fld_handle* h = new fld_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) h->encode_tiled = fn;
  else (void)cudaGetLastError();
  *out = h;
  return FLD_OK;
}

extern "C" void fld_destroy(fld_handle* h) {
  if (!h) return;
  delete h;
}
